"""First contact of the experimental chunk-skipping aggregation (S2PB_CHUNKED=1) with a GPU: mgm_multi on small
tiles against the oracle, with a matcher timeout so that a deadlock is drained through the abort flag.
usage: S2PB_CHUNKED=1 S2PB_CHUNKED_MIN_DP=0 python scripts/chunked_probe.py   (MIN_DP=0: also on the narrow slabs of the small
cases; once with S2PB_CHUNKED=0 for the timing of the dense kernels; BIG=only COMPARE=/tmp/x for the 768x532x256 tile, dense first)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import oracle as O
from s2p_b200.engine import Engine, default_params
from s2p_b200.synth import make_pair

eng = Engine(0)
cases = [((120, 160), -20, 20, 1, dict(subpix=1, remove_small_cc=0)), ((120, 160), -20, 20, 1, dict()), ((230, 260), -40, 40, 3, dict(scales=3))]
if os.environ.get("BIG"):
    cases = [((532, 768), -128, 127, 5, dict())] if os.environ.get("BIG") == "only" else cases + [((532, 768), -128, 127, 5, dict())]
eq = lambda a, b: int((~((a == b) | (np.isnan(a) & np.isnan(b)))).sum())
for (h, w), dmin, dmax, seed, kw in cases:
    ref, sec, _ = make_pair(h, w, dmin, dmax, seed=seed)
    p = default_params("mgm_multi", timeout_ms=15000, **kw)
    try:
        out = eng.mgm(ref, sec, dmin, dmax, p, want_right=True)
        t = time.perf_counter(); out = eng.mgm(ref, sec, dmin, dmax, p, want_right=True); dt = time.perf_counter() - t
    except Exception as e:
        print("case", (h, w), kw, "FAILED:", e, flush=True)
        break
    ref_file = os.environ.get("COMPARE")       # compare with (or save for) another run instead of running the CPU oracle
    if ref_file:
        tag = "%s_%dx%d_%d.npz" % (ref_file, w, h, len(kw))
        if os.path.exists(tag):
            z = np.load(tag)
            d, c, dr = z["d"], z["c"], z["dr"]
        else:
            np.savez(tag, d=out["disp"], c=out["conf"], dr=out["disp_right"])
            print("case", (h, w), kw, "saved, %.1f ms" % (dt * 1e3), flush=True)
            continue
    else:
        d, c, dr = O.port.mgm_multi(ref, sec, dmin, dmax, O.mgm_multi_params(**kw))
    print("case", (h, w), kw, "chunked" if os.environ.get("S2PB_CHUNKED") == "1" else "dense", "%.1f ms | mismatch disp %d conf %d dispR %d of %d" % (
        dt * 1e3, eq(out["disp"], d), eq(out["conf"], c), eq(out["disp_right"], dr), d.size), flush=True)
