import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from s2p_b200.engine import Engine
from s2p_b200.synth import _blur
eng = Engine(0)
rng = np.random.default_rng(0)
src = _blur(rng.integers(0, 4096, size=(1500, 1500)).astype(np.float64)).astype(np.float32)
th, s = 1.36, 0.9885
for name, H in [("aa", np.array([[np.cos(th) * s, -np.sin(th) * s, 1100.0], [np.sin(th) * s, np.cos(th) * s, -350.0], [0, 0, 1.0]])),
                ("plain", np.array([[1.0, 0.02, -100.0], [-0.01, 1.0, -80.0], [0, 0, 1.0]]))]:
    for it in range(3):
        t = time.perf_counter(); out = eng.homography(src, H, 1024, 1024); dt = time.perf_counter() - t
    print(name, "%.2f ms (host wall, H2D+D2H included)" % (dt * 1e3), "nan %.3f" % np.isnan(out).mean(), flush=True)
