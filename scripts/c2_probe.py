"""One 1024x1024x128 tile (BASELINE config 2): stage timings + full-size parity against the oracle port."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from s2p_b200.engine import Engine, default_params
from s2p_b200.synth import make_pair

h = w = int(os.environ.get("N", 1024)); dmin, dmax = -64, 63
ref, sec, gt = make_pair(h, w, dmin, dmax, seed=0, nan_border=float(os.environ.get("NANB", 0)))
eng = Engine(0)
p = default_params("mgm")
for it in range(4):
    t = time.perf_counter(); out = eng.mgm(ref, sec, dmin, dmax, p); dt = time.perf_counter() - t
    print("iter", it, "host wall %.1f ms" % (dt * 1e3), {k: round(v, 3) for k, v in eng.last_timings().items()}, flush=True)
print("valid %.3f  median |err| vs gt %.3f" % (np.isfinite(out["disp"]).mean(), np.nanmedian(np.abs(out["disp"] - gt))))
if os.environ.get("PARITY", "1") == "1":
    from oracle import oracle as O
    t = time.perf_counter(); d, c, dr = O.port.mgm(ref, sec, dmin, dmax, O.mgm_params()); print("oracle port %.1f s" % (time.perf_counter() - t))
    eq = lambda a, b: int((~((a == b) | (np.isnan(a) & np.isnan(b)))).sum())
    print("mismatch disp", eq(out["disp"], d), "conf", eq(out["conf"], c), "mask", int((out["mask"] != O.port.rejection_mask(d, ref, sec)).sum()))
