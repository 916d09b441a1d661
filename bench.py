#!/usr/bin/env python
"""Benchmark of the stereo-matching hot path (BASELINE.json: disparity Mpix/s and cost-volume
Gvoxel/s on 1024x1024 rectified tiles, disp_range 128, census 5x5 + 8-path MGM, `mgm` flags of s2p).

    python bench.py --gpus N --steps K --warmup W              # our CUDA engine
    python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU `mgm` on host cores

A "step" = one pass of the hot path over a batch of `--tiles` synthetic rectified pairs per GPU.
`value`  : whole-job Mpix/s with the pairs already resident in HBM (device pointers through
           s2pb_mgm_device, CUDA-event timed on the launching streams, max over ranks).
`e2e`    : the same metric through the reference-facing C-ABI call with HOST buffers
           (s2pb_mgm_batch from page-locked host buffers: H2D, kernels, D2H every step).
`roofline`: the 8-path aggregation kernel, algorithmic bytes (12 B per voxel
           per path = 192 B per left-reference voxel for the two views) / its CUDA-event duration.
`cpu_baseline`: the reference's own `mgm` binary (oracle/_ref, built from the reference sources)
           on a bounded sample of the same workload, on this box's host cores.
Tiles are independent: with N GPUs each rank processes its own tiles, no data-path collective.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "disparity_throughput_1024x1024x128_mgm"
UNIT = "Mpix/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tiles", type=int, default=16, help="tiles per GPU per step")
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--dmin", type=int, default=-64)
    ap.add_argument("--dmax", type=int, default=63)
    ap.add_argument("--slots", type=int, default=8, help="tiles in flight per GPU (one 8.5 GiB workspace each at the default shape)")
    ap.add_argument("--cpu-procs", type=int, default=0, help="concurrent reference processes (0 = calibrate: all host cores, 1/2, 1/4)")
    ap.add_argument("--cpu-rows", type=int, default=32, help="rows of the CPU sample strips")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--nan-border", type=float, default=0.0,
                    help="fraction of the tile width turned into no-data strips (0 = the BASELINE workload; > 0 exercises the "
                         "no-data sentinel range, which can widen the right view's slab: DESIGN.md, limits)")
    return ap.parse_args()


def config(a, world):
    return {
        "workload": "%ssingle %dx%d rectified tile, disp_range=%d, census5x5 + 8-path MGM "
                    "(s2p algo 'mgm': TSGM=3, P1=8, P2=32, vfit, MEDIAN=1, LR check), %d tiles per GPU per step" % (
                        "BASELINE configs[1]: " if (a.size, a.dmin, a.dmax) == (1024, -64, 63) else "", a.size, a.size,
                        a.dmax - a.dmin + 1, a.tiles),
        "tile": [a.size, a.size], "dmin": a.dmin, "dmax": a.dmax, "labels": a.dmax - a.dmin + 1,
        "tiles_per_gpu_per_step": a.tiles, "tiles_in_flight": a.slots, "parallelism": "tile-shard x%d" % world,
        "nan_border": a.nan_border,
        "l2": "per-tile working set %.1f GiB (8 float path volumes per view) >> 126 MB L2; inputs differ per tile" % (
            2 * 8 * 4.0 * a.size * a.size * (32 * ((a.dmax - a.dmin + 32) // 32)) / 2 ** 30 * 1.0625),
    }


# ----------------------------------------------------------------------------- clocks

class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def window(self, t0, t1):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        sm, reasons = [], set()
        for t, line in self.rows:
            if t < t0 or t > t1:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                out["sm_max_mhz"] = float(f[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
        out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out

    def stop(self):
        if self.proc:
            self.proc.terminate()


# ----------------------------------------------------------------------------- reference CPU arm

def cpu_reference_step(strips, dmin, dmax, procs):
    """Run `procs` single-thread reference `mgm` processes concurrently (s2p's deployment mode:
    one process per tile, OMP_NUM_THREADS=1, s2p/config.py:46), one strip each.  -> seconds."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    P = O.mgm_params()
    t0 = time.perf_counter()
    with ThreadPoolExecutor(procs) as ex:
        list(ex.map(lambda rs: O.run_ref(rs[0], rs[1], dmin, dmax, P, threads=1), strips))
    return time.perf_counter() - t0


def make_strips(a, procs, base_tiles=None):
    from s2p_b200.synth import make_pair
    rows = min(a.cpu_rows, a.size)
    tiles = base_tiles or [make_pair(a.size, a.size, a.dmin, a.dmax, seed=s)[:2] for s in range(min(procs, 4))]
    strips = []
    for k in range(procs):
        ref, sec = tiles[k % len(tiles)]
        r0 = (k // len(tiles) * rows) % max(1, a.size - rows + 1)
        strips.append((ref[r0:r0 + rows], sec[r0:r0 + rows]))
    return strips, rows


def cpu_proc_candidates(a):
    """Process counts to try: the reference is memory-bound with many concurrent single-thread processes, so
    'all host threads' is not always its fastest deployment -- the best of (all, 1/2, 1/4 of the cores) is used."""
    n = os.cpu_count() or 1
    if a.cpu_procs > 0:
        return [a.cpu_procs]
    return sorted({n, max(1, n // 2), max(1, n // 4)}, reverse=True)


def cpu_best(a, base_tiles=None):
    """-> (Mpix/s, procs, rows, seconds, strips) of the fastest candidate, one step each."""
    best = None
    for procs in cpu_proc_candidates(a):
        strips, rows = make_strips(a, procs, base_tiles=base_tiles)
        secs = cpu_reference_step(strips, a.dmin, a.dmax, procs)
        mpix = procs * rows * a.size / secs / 1e6
        if best is None or mpix > best[0]:
            best = (mpix, procs, rows, secs, strips)
    return best


def run_reference(a, rank, world):
    if rank != 0:
        return
    from oracle import oracle as O
    if not O.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/mgm not built (make -C oracle ref)"}))
        return
    _, procs, rows, _, strips = cpu_best(a)          # calibration doubles as warm-up
    for _ in range(max(0, a.warmup - len(cpu_proc_candidates(a)))):
        cpu_reference_step(strips, a.dmin, a.dmax, procs)
    t = [cpu_reference_step(strips, a.dmin, a.dmax, procs) for _ in range(a.steps)]
    total = sum(t)
    pix = procs * rows * a.size * a.steps
    val = pix / total / 1e6
    D = a.dmax - a.dmin + 1
    sample = "%d strips of %dx%d px, %d labels, one single-thread reference `mgm` process each (OMP_NUM_THREADS=1), PFM I/O included" % (
        procs, a.size, rows, D)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config(a, world), "gvoxel_per_s": val * D / 1e3,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": procs, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- our arm

def run_ours(a, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from s2p_b200.engine import Engine, default_params
    from s2p_b200.synth import make_pair

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    W = H = a.size
    D = a.dmax - a.dmin + 1
    B = a.tiles
    p = default_params("mgm")
    eng = Engine(local_rank)
    nslots = max(1, min(a.slots, B))
    eng.reserve(nslots, W, H, D)

    # synthetic rectified pairs, distinct per tile and per rank (seed = global tile id)
    pairs = [make_pair(H, W, a.dmin, a.dmax, seed=rank * B + t, nan_border=a.nan_border)[:2] for t in range(B)]
    d_ref = [torch.from_numpy(r).to(dev) for r, _ in pairs]
    d_sec = [torch.from_numpy(s).to(dev) for _, s in pairs]
    d_disp = [torch.empty((H, W), dtype=torch.float32, device=dev) for _ in range(B)]
    d_conf = [torch.empty((H, W), dtype=torch.float32, device=dev) for _ in range(B)]
    d_mask = [torch.empty((H, W), dtype=torch.uint8, device=dev) for _ in range(B)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(nslots)]
    main = torch.cuda.current_stream(dev)

    # The tile queue of a step is spread over `nslots` workspaces, each with its own stream; consecutive steps keep
    # the queue full (a workspace's stream orders its own tiles), the streams are joined where the timed region ends.
    def fork():
        ev = torch.cuda.Event()
        ev.record(main)
        for st in streams:
            st.wait_event(ev)

    def join():
        for st in streams:
            ev = torch.cuda.Event()
            ev.record(st)
            main.wait_event(ev)

    def device_step():
        for t in range(B):
            sl = t % nslots
            eng.mgm_device(sl, d_ref[t].data_ptr(), d_sec[t].data_ptr(), W, H, a.dmin, a.dmax, p, d_disp[t].data_ptr(),
                           d_conf[t].data_ptr(), d_mask[t].data_ptr(), 0, nodata_hint=(2 if a.nan_border > 0 else 0), stream=streams[sl].cuda_stream)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank) if rank == 0 else None

    # ---- device-resident arm
    fork()
    for _ in range(max(3, a.warmup)):
        device_step()
    join()
    barrier()
    l0 = eng.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.perf_counter()
    e0.record(main)
    fork()
    for _ in range(a.steps):
        device_step()
    join()
    e1.record(main)
    barrier()
    tw1 = time.perf_counter()
    launches = eng.kernel_launches() - l0
    ms = max_over_ranks(e0.elapsed_time(e1))
    # stage durations of the LAST tile of each workspace while the tiles overlap (diagnostic: how much the
    # memory-bound WTA stretches when it shares the SMs with the next tile's aggregation)
    overlapped = {}
    try:
        per_slot = [eng.last_timings(s) for s in range(len(streams))]
        overlapped = {k: float(np.mean([t[k] for t in per_slot])) for k in per_slot[0]}
    except Exception:
        pass
    pix_step = world * B * W * H
    value = pix_step * a.steps / (ms * 1e-3) / 1e6

    # checksum of the device results (and a sanity check that the engine produced disparities)
    valid = float(torch.isfinite(d_disp[0]).float().mean().item())

    # ---- end-to-end arm: host buffers through the C ABI
    # host buffers are page-locked (torch pinned tensors viewed as numpy): the library DMAs from / to them directly
    keep = []

    def pinned(arr):
        t = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
        keep.append(t)
        return t.numpy()
    refs = [pinned(r) for r, _ in pairs]
    secs = [pinned(s) for _, s in pairs]
    outs = ([pinned(np.empty((H, W), np.float32)) for _ in pairs], [pinned(np.empty((H, W), np.float32)) for _ in pairs],
            [pinned(np.empty((H, W), np.uint8)) for _ in pairs])
    for _ in range(max(1, min(3, a.warmup))):
        eng.mgm_batch(refs, secs, a.dmin, a.dmax, p, out=outs)
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        out = eng.mgm_batch(refs, secs, a.dmin, a.dmax, p, out=outs)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e = pix_step * a.steps / e2e_s / 1e6
    h2d = B * 2 * W * H * 4
    d2h = B * (2 * W * H * 4 + W * H)

    # ---- roofline of the dominant kernel: serial launches on one stream, the library's own CUDA events
    agg_ms, tot_ms, stage = [], [], {}
    for k in range(max(3, min(a.steps, 10))):
        t = k % B
        eng.mgm_device(0, d_ref[t].data_ptr(), d_sec[t].data_ptr(), W, H, a.dmin, a.dmax, p, d_disp[t].data_ptr(),
                       d_conf[t].data_ptr(), d_mask[t].data_ptr(), 0, nodata_hint=(2 if a.nan_border > 0 else 0), stream=streams[0].cuda_stream)
        streams[0].synchronize()
        tm = eng.last_timings(0)
        agg_ms.append(tm["aggregate"])
        tot_ms.append(tm["total"])
        stage = tm
    agg = float(np.mean(agg_ms[1:]))
    alg_bytes = 192.0 * W * H * D          # 12 B/voxel/path x 8 paths x 2 views
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (agg * 1e-3) / 1e9
    traffic = None
    if (W, H, D) == (1024, 1024, 128):      # the committed ncu capture is of this configuration
        try:
            traffic = float(json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["aggregate_kernel"]["bytes"])
        except (OSError, ValueError, KeyError):
            traffic = None
    roofline = {"bound": "hbm", "kernel": "aggregate_kernel (8 passes x 2 views, one persistent launch)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                "traffic": traffic, "traffic_source": "profiles/traffic.json (ncu --set full, dram__bytes_read+write per launch)",
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": agg,
                "tile_ms_serial": float(np.mean(tot_ms[1:])), "stage_ms": stage,
                "stage_ms_overlapped": overlapped, "how": "serial single-tile launches after the timed region, CUDA events recorded by the library on the launching stream"}

    clocks = None
    if sampler:
        clocks = sampler.window(tw0, tw1)
        sampler.stop()

    # ---- CPU baseline beside it (rank 0, N=1 only): the reference binary on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        from oracle import oracle as O
        if O.have_ref():
            mpix, procs, rows, secs_cpu, _ = cpu_best(a, base_tiles=pairs[:4])
            cpu = {"value": mpix, "unit": UNIT, "cores": procs, "kind": "reference",
                   "sample": "%d strips of %dx%d px, %d labels, one single-thread reference `mgm` process each "
                             "(OMP_NUM_THREADS=1, the way s2p deploys it), %.1f s wall; fastest of %s concurrent processes on %d host threads"
                             % (procs, W, rows, D, secs_cpu, cpu_proc_candidates(a), os.cpu_count() or 1)}
        else:
            t0 = time.perf_counter()
            O.port.mgm(pairs[0][0][:64], pairs[0][1][:64], a.dmin, a.dmax, O.mgm_params())
            dt = time.perf_counter() - t0
            cpu = {"value": 64 * W / dt / 1e6, "unit": UNIT, "cores": min(8, os.cpu_count() or 1), "kind": "port",
                   "sample": "one %dx64 strip through oracle/liboracle.so (OpenMP over the 8 passes)" % W}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup),
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config(a, world), "gvoxel_per_s": value * D / 1e3,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "timer": "host wall clock around the synchronised region, max over ranks; page-locked host buffers on both sides",
                    "gvoxel_per_s": e2e * D / 1e3},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "valid_fraction_tile0": valid,
        }
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if a.impl == "reference":
        run_reference(a, rank, world)
    else:
        run_ours(a, rank, world, local_rank)


if __name__ == "__main__":
    main()
