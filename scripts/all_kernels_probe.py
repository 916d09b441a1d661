"""Runs every CUDA kernel of the library once on representative sizes (for an ncu launch list / full capture)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from s2p_b200.engine import Engine, default_params
from s2p_b200.synth import make_pair, _blur

eng = Engine(0)
# matcher, BASELINE config 2 and a C3-like mgm_multi tile
ref, sec, _ = make_pair(1024, 1024, -64, 63, seed=0)
for _ in range(2):
    out = eng.mgm(ref, sec, -64, 63, default_params("mgm"))
# the same tile with 5 % no-data strips: the DCT round trip of the matched image, per-view slab widths
refn, secn, _ = make_pair(1024, 1024, -64, 63, seed=1, nan_border=0.05)
eng.mgm(refn, secn, -64, 63, default_params("mgm"), want_pkr=True)
# a range beyond 512 labels: the chunk-skipping kernels
rw, sw_, _ = make_pair(64, 700, -300, 299, seed=2)
eng.mgm(rw, sw_, -300, 299, default_params("mgm"))
# general flavour: another distance, NCC, census with -wl / -wr weights
wl = np.maximum(np.random.default_rng(1).uniform(0, 1, ref.shape) ** 2, 0.1).astype(np.float32)
eng.mgm(ref, sec, -64, 63, default_params("mgm", cost="btad"))
eng.mgm(ref, sec, -64, 63, default_params("mgm", cost="ncc"))
eng.mgm(ref, sec, -64, 63, default_params("mgm"), weights=(wl, wl))
r3, s3, _ = make_pair(532, 768, -128, 127, seed=5)
for _ in range(2):
    eng.mgm(r3, s3, -128, 127, default_params("mgm_multi"))
# rectification warp: a 1024x1024 output from a rotated, slightly shrinking homography (anti-aliasing branch)
rng = np.random.default_rng(0)
src = _blur(rng.integers(0, 4096, size=(1500, 1500)).astype(np.float64)).astype(np.float32)
th, s = 1.36, 0.9885
H = np.array([[np.cos(th) * s, -np.sin(th) * s, 1100.0], [np.sin(th) * s, np.cos(th) * s, -350.0], [0, 0, 1.0]])
for _ in range(2):
    eng.homography(src, H, 1024, 1024)
eng.homography(src, np.array([[1.0, 0.02, -100.0], [-0.01, 1.0, -80.0], [0, 0, 1.0]]), 1024, 1024)
# next rows: mask erosion, n-view merge, triangulation
eng.erode_mask(out["mask"], 2)
eng.merge_n([out["disp"], out["disp"] + 0.5], [0.0, 0.5], "average_if_close", 3)
import test_triangulation as T
from s2p_b200.triangulation import disp_to_lonlatalt
H1, H2, bbx = T._geometry()
disp = rng.normal(0, 6, (512, 512)).astype(np.float32)
disp_to_lonlatalt(disp, np.ones((512, 512), np.float32), np.ones((111, 151), np.float32), H1, H2, T._rpc(1, False), T._rpc(2, False), bbx, engine=eng)
print("done", eng.kernel_launches(), "launches")
