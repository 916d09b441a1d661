"""How ragged are the label ranges of mgm_multi's levels?  Runs the CPU oracle on a C3-like tile with its range
dump hook (ORC_DUMP_RANGES) and compares, per mgm_call, the work of a dense slab (hull rounded to the slab widths
the engine has), of a fully ragged volume (per-pixel widths rounded to 32 labels) and of a lock-step band of 16
scanlines that only processes, at every step, the 32-label chunks its 16 pixels need.  CPU only.
usage: OMP_NUM_THREADS=1 python scripts/range_width_analysis.py [h w dmin dmax seed]"""
import glob
import os
import re
import sys
import tempfile

out = tempfile.mkdtemp(prefix="s2pb_ranges_")
os.environ["ORC_DUMP_RANGES"] = out
os.environ.setdefault("OMP_NUM_THREADS", "1")        # the dump counter is not thread safe
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import oracle as O
from s2p_b200.synth import make_pair

a = [int(x) for x in sys.argv[1:6]] + [532, 768, -128, 127, 5][len(sys.argv) - 1:]
h, w, dmin, dmax, seed = a
ref, sec, _ = make_pair(h, w, dmin, dmax, seed=seed)
d, _, _ = O.port.mgm_multi(ref, sec, dmin, dmax, O.mgm_multi_params())
print("tile %dx%d, %d labels: %.1f %% valid" % (w, h, dmax - dmin + 1, 100 * np.isfinite(d).mean()))


def lock_step(pw):            # scanlines = rows of pw: bands of 16, every step costs the widest of the band's pixels
    tot = 0
    for b in range(0, pw.shape[0], 16):
        blk = pw[b:b + 16]
        tot += int(blk.max(axis=0).sum()) * blk.shape[0]
    return tot


tot = np.zeros(3)
for f in sorted(glob.glob(os.path.join(out, "ranges_*.bin"))):
    m = re.search(r"ranges_(\d+)_z(\d)_(\d+)x(\d+)", f)
    idx, z, ww, hh = (int(m.group(k)) for k in range(1, 5))
    r = np.fromfile(f, np.int32)
    n = ww * hh
    lo, hi = r[:n].reshape(hh, ww), r[n:].reshape(hh, ww)
    wd = hi - lo + 1
    pw = (wd + 31) // 32 * 32
    hull = int(hi.max() - lo.min() + 1)
    dp = next((32 * k for k in (1, 2, 3, 4, 5, 6, 8, 12, 16) if 32 * k >= hull), hull)
    ls = 0.5 * (lock_step(pw) + lock_step(pw.T))
    print("call %2d zoom %d %4dx%-4d hull %4d -> slab %4d | mean width %6.1f | wider than 128: %5.1f %% | Mvoxel: dense %6.1f  ragged %6.1f  "
          "lock-step chunks %6.1f" % (idx, z, ww, hh, hull, dp, wd.mean(), 100 * (wd > 128).mean(), n * dp / 1e6, pw.sum() / 1e6, ls / 1e6))
    tot += (n * dp, pw.sum(), ls)
print("total Mvoxel: dense %.1f  ragged %.1f (%.2fx less)  lock-step chunks %.1f (%.2fx less)" % (
    tot[0] / 1e6, tot[1] / 1e6, tot[0] / tot[1], tot[2] / 1e6, tot[0] / tot[2]))
