"""ncu raw CSV of scripts/all_kernels_probe.py -> per-kernel table (time, DRAM bytes, GB/s, fraction of the measured peak).
usage: ncu -i rep --page raw --csv | python scripts/kernel_table.py [peak_GBps] > profiles/xxx.md"""
import csv, sys
peak = float(sys.argv[1]) if len(sys.argv) > 1 else 6567.7
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
sc = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
def val(r, k):
    return float(r[ix[k]]) * sc.get(units[ix[k]], 1.0)
agg = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("s2pb::", "")
    t = val(r, "gpu__time_duration.sum"); b = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    grid = r[ix["launch__grid_size"]]; regs = r[ix["launch__registers_per_thread"]]
    key = (name, grid)
    a = agg.setdefault(key, [0, 0.0, 0.0, regs, 0.0])
    a[0] += 1; a[1] += t; a[2] += b
    a[4] += float(r[ix["smsp__issue_active.avg.pct_of_peak_sustained_active"]])
print("| kernel | grid | launches | avg time | DRAM bytes / launch | DRAM GB/s | of measured %.0f GB/s | issue active | regs |" % peak)
print("|---|---|---|---|---|---|---|---|---|")
for (name, grid), (n, t, b, regs, ia) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    gbs = b / t / 1e9 if t else 0
    print("| `%s` | %s | %d | %.1f us | %.2f MB | %.0f | %.2f | %.0f %% | %s |" % (name[:60], grid, n, t / n * 1e6, b / n / 1e6, gbs, gbs / peak, ia / n, regs))
