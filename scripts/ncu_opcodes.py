"""Dynamic opcode histogram of a kernel from an ncu report's source page (warp-level executed instructions per SASS mnemonic).
usage: ncu -i rep --page source --csv > src.csv ; python scripts/ncu_opcodes.py src.csv [N]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ci = {c: i for i, c in enumerate(hdr)}
tot = collections.Counter()
for r in rows[h + 1:]:
    if len(r) < len(hdr) or r[0] == "Address":
        continue
    try:
        ex = int(r[ci["Instructions Executed"]])
    except ValueError:
        continue
    src = re.sub(r"^@!?U?P\d+\s+", "", r[ci["Source"]].strip())
    op = ".".join(src.split()[0].split(".")[:2]) if src else "?"
    tot[op] += ex
total = sum(tot.values())
print("warp instructions executed: %d" % total)
for op, c in tot.most_common(n):
    print("%-22s %12d  %5.1f %%" % (op, c, 100.0 * c / total))
