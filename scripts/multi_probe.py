"""mgm_multi on the GPU with a watchdog: prints timings and mismatch counts against the oracle port."""
import faulthandler, os, sys, time
faulthandler.dump_traceback_later(int(os.environ.get("WATCHDOG", 50)), exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from s2p_b200.engine import Engine, default_params
from s2p_b200.synth import make_pair
from oracle import oracle as O
eng = Engine(0)
cases = [((120, 160), -20, 20, 1, dict(subpix=1, remove_small_cc=0)), ((120, 160), -20, 20, 1, dict(subpix=1)), ((120, 160), -20, 20, 1, dict()),
         ((230, 260), -40, 40, 3, dict(scales=3))]
if os.environ.get("BIG"): cases.append(((532, 768), -128, 127, 5, dict()))
for (h, w), dmin, dmax, seed, kw in cases:
    ref, sec, gt = make_pair(h, w, dmin, dmax, seed=seed)
    print("case", (h, w), kw, flush=True)
    t = time.perf_counter(); out = eng.mgm(ref, sec, dmin, dmax, default_params("mgm_multi", **kw), want_right=True); tg = time.perf_counter() - t
    t = time.perf_counter(); out = eng.mgm(ref, sec, dmin, dmax, default_params("mgm_multi", **kw), want_right=True); tg2 = time.perf_counter() - t
    t = time.perf_counter(); d, c, dr = O.port.mgm_multi(ref, sec, dmin, dmax, O.mgm_multi_params(**kw)); to = time.perf_counter() - t
    eq = lambda a, b: int((~((a == b) | (np.isnan(a) & np.isnan(b)))).sum())
    both = np.isfinite(d) & np.isfinite(out["disp"])
    print("  gpu %.1f ms (first %.1f)  oracle %.2f s | mismatch disp %d conf %d dispR %d of %d | nan-diff %d  max|d| %.4f  >0.25px: %d" % (
        tg2 * 1e3, tg * 1e3, to, eq(out["disp"], d), eq(out["conf"], c), eq(out["disp_right"], dr), d.size,
        int((np.isnan(d) != np.isnan(out["disp"])).sum()), float(np.abs(d - out["disp"])[both].max()) if both.any() else 0, int((np.abs(d - out["disp"])[both] > 0.25).sum())), flush=True)
