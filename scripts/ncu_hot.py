"""List the hottest SASS instructions of an ncu report's source page with their stall reasons.
usage: ncu -i rep --page source --csv > src.csv ; python scripts/ncu_hot.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ci = {c: i for i, c in enumerate(hdr)}
stalls = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
tot = 0; tot_exec = 0
L = []
for k, r in enumerate(rows[h + 1:]):
    if len(r) < len(hdr) or r[0] == "Address": continue
    try: v = int(r[ci["Warp Stall Sampling (All Samples)"]]); ex = int(r[ci["Instructions Executed"]])
    except ValueError: continue
    tot += v; tot_exec += ex
    L.append((v, k, r))
print("total samples", tot, "warp-instructions executed", tot_exec)
agg = {s: 0 for s in stalls}
for v, k, r in L:
    for s in stalls:
        try: agg[s] += int(r[ci[s]])
        except ValueError: pass
print("by reason:", {k: v for k, v in sorted(agg.items(), key=lambda t: -t[1]) if v})
for v, k, r in sorted(L, key=lambda t: -t[0])[:n]:
    top = sorted(((int(r[ci[s]] or 0), s) for s in stalls), reverse=True)[:3]
    print("%7d  #%4d  exec=%9s  %-70s %s" % (v, k, r[ci["Instructions Executed"]], r[ci["Source"]].strip()[:70], [(s[6:], c) for c, s in top if c]))
