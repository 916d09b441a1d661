"""How long does the host take to enqueue a tile (s2pb_mgm_device) when the GPU queue is empty, and what is the
device throughput for 1, 2, 4, 8 tiles in flight?  Usage: python scripts/enqueue_probe.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from s2p_b200.engine import Engine, default_params
from s2p_b200.synth import make_pair

dev = torch.device("cuda", 0)
eng = Engine(0)
H = W = 1024
dmin, dmax = -64, 63
B = 8
pairs = [make_pair(H, W, dmin, dmax, seed=t)[:2] for t in range(B)]
d_ref = [torch.from_numpy(r).to(dev) for r, _ in pairs]
d_sec = [torch.from_numpy(s).to(dev) for _, s in pairs]
d_disp = [torch.empty((H, W), dtype=torch.float32, device=dev) for _ in range(B)]
d_conf = [torch.empty((H, W), dtype=torch.float32, device=dev) for _ in range(B)]
d_mask = [torch.empty((H, W), dtype=torch.uint8, device=dev) for _ in range(B)]
p = default_params("mgm")
for nslots in (1, 2, 3, 4, 8):
    eng.reserve(nslots, W, H, 128)
    streams = [torch.cuda.Stream(device=dev) for _ in range(nslots)]

    def run(n):
        for t in range(n):
            sl = t % nslots
            k = t % B
            eng.mgm_device(sl, d_ref[k].data_ptr(), d_sec[k].data_ptr(), W, H, dmin, dmax, p, d_disp[k].data_ptr(),
                           d_conf[k].data_ptr(), d_mask[k].data_ptr(), 0, nodata_hint=0, stream=streams[sl].cuda_stream)
    run(2 * nslots)
    torch.cuda.synchronize()
    N = 48
    t0 = time.perf_counter()
    run(N)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("slots %d: host enqueue %.3f ms/tile, total %.3f ms/tile -> %.1f Mpix/s" % (
        nslots, (t1 - t0) / N * 1e3, (t2 - t0) / N * 1e3, N * H * W / (t2 - t0) / 1e6), flush=True)
