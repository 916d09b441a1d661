"""One BASELINE configs[2] tile (768x532, 256 labels, mgm_multi) on its own: the workload for an ncu launch list of the pyramid.
Usage: ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/c3_probe.py [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from s2p_b200.engine import Engine, default_params
from s2p_b200.synth import make_pair

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
W, H, dmin, dmax = 768, 532, -128, 127
eng = Engine(0)
p = default_params("mgm_multi")
ref, sec = make_pair(H, W, dmin, dmax, seed=1000)[:2]
dev = torch.device("cuda:0")
a, b = torch.from_numpy(ref).to(dev), torch.from_numpy(sec).to(dev)
o = [torch.empty((H, W), dtype=torch.float32, device=dev) for _ in range(2)] + [torch.empty((H, W), dtype=torch.uint8, device=dev)]
st = torch.cuda.Stream(device=dev)
for k in range(1 + reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.mgm_device(0, a.data_ptr(), b.data_ptr(), W, H, dmin, dmax, p, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), 0, stream=st.cuda_stream)
    torch.cuda.synchronize(); print("tile %d: %.2f ms" % (k, 1e3 * (time.perf_counter() - t0)), "valid", float(torch.isfinite(o[0]).float().mean()))
