import faulthandler, os, sys
faulthandler.dump_traceback_later(30, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from s2p_b200.engine import Engine
from oracle import oracle as O
eng = Engine(0)
rng = np.random.default_rng(0)
for (h, w) in [(60, 80), (120, 160), (300, 500)]:
    a = (rng.integers(0, 4, size=(h, w)) * 6).astype(np.float32) + rng.normal(0, 0.5, (h, w)).astype(np.float32)
    a[rng.random((h, w)) < 0.15] = np.nan
    print("cc", (h, w), flush=True)
    g = eng.remove_small_cc(a, 25)
    o = O.port.remove_small_cc(a, 25)
    print("  mismatch", int((~((g == o) | (np.isnan(g) & np.isnan(o)))).sum()), "removed", int(np.isnan(o).sum() - np.isnan(a).sum()), flush=True)
