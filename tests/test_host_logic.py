"""CPU tests of the host side: the drop-in's range / parameter / error logic, the C-ABI library
(loads, exports every declared symbol, refuses to compute without a GPU), raster I/O, tile sharding."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from s2p_b200 import _lib
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "s2pb200.h")).read()
    declared = sorted(set(re.findall(r"\b(s2pb_[a-z_0-9]+)\s*\(", hdr)))
    assert declared == _lib.exported_symbols(), "ctypes table and header disagree"
    for name in declared:
        assert hasattr(L, name), name
    assert L.s2pb_version() == 101


def test_no_cpu_fallback():
    from s2p_b200 import _lib
    from s2p_b200.engine import Engine, S2pbError
    if _lib.lib().s2pb_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(S2pbError) as e:
        Engine(0)
    assert "no CPU path" in str(e.value)


def test_default_params_mirror_s2p_flags():
    from s2p_b200.engine import default_params
    p = default_params("mgm")            # s2p/block_matching.py:155-186
    assert (p.tsgm, p.median, p.census_win, p.ndir, p.refine, p.scales, p.subpix) == (3, 1, 5, 8, 1, -1, 1)
    assert (p.P1, p.P2, p.lr_mode, p.lr_tau, p.mindiff, p.remove_small_cc) == (8.0, 32.0, 1, 1.0, -1.0, 0)
    q = default_params("mgm_multi")      # s2p/block_matching.py:269-308
    assert (q.tsgm, q.median, q.scales, q.subpix, q.remove_small_cc) == (4, 0, 6, 2, 25)
    r = default_params("mgm_multi_lsd")  # s2p/block_matching.py:191-266: P1 = 12, P2 = 48, MEDIAN = 1, SUBPIX = 2, -S 6
    assert (r.tsgm, r.median, r.scales, r.subpix, r.remove_small_cc, r.P1, r.P2, r.cost) == (4, 1, 6, 2, 25, 12.0, 48.0, 0)
    assert default_params("mgm", cost="ncc").cost == 3 and default_params("mgm", cost=4).cost == 4
    with pytest.raises(ValueError):
        default_params("mgm", cost="l1")
    with pytest.raises(Exception):
        default_params("sgbm")


def test_params_struct_matches_the_header():
    """ctypes mirror and C struct must agree field for field (order, width): parse the header's struct."""
    import ctypes
    import re
    from s2p_b200 import _lib
    src = open(os.path.join(ROOT, "include", "s2pb200.h")).read()
    body = re.search(r"typedef struct s2pb_mgm_params \{(.*?)\} s2pb_mgm_params;", src, re.S).group(1)
    fields = re.findall(r"^\s*(int32_t|float)\s+(\w+);", body, re.M)
    assert [n for _, n in fields] == [n for n, _ in _lib.MgmParams._fields_]
    assert [t for t, _ in fields] == ["int32_t" if t is ctypes.c_int32 else "float" for _, t in _lib.MgmParams._fields_]
    assert ctypes.sizeof(_lib.MgmParams) == 4 * len(fields)
    enum = re.search(r"enum \{ (S2PB_COST_CENSUS.*?)\};", src, re.S).group(1)
    names = [x.strip().split(" ")[0] for x in enum.split(",") if x.strip()]
    assert [n[len("S2PB_COST_"):].lower() for n in names[:-1]] == list(_lib.COSTS) and names[-1] == "S2PB_COST_COUNT"


def test_disparity_bounds_and_errors():
    from s2p_b200 import block_matching as bm
    assert bm.disparity_bounds(1000, -3.5, 7.2) == (-4, 8)
    assert bm.disparity_bounds(100, -200, 300) == (0, 100)           # clamped around the centre
    assert bm.disparity_bounds(100, None, None) == (None, None)
    with pytest.raises(bm.MaxDisparityRangeError):                    # tests/block_matching_test.py:23-36 in the reference
        bm.disparity_bounds(1024, -100, 100, max_disp_range=10)
    assert bm.confidence_path("/x/rectified_disp.tif") == "/x/rectified_disp_confidence.tif"
    assert bm.confidence_path("/x/rectified_disp.tif", "mgm_multi_lsd") == "/x/rectified_disp.tif.confidence.tif"   # :238
    with pytest.raises(NotImplementedError):
        bm.compute_disparity_map("a", "b", "c", "d", "sgbm")


def test_matcher_params_follow_cfg():
    from s2p_b200 import block_matching as bm
    from s2p_b200.config import cfg
    old = dict(cfg)
    try:
        cfg.update(census_ncc_win=3, mgm_nb_directions=4, mgm_leftright_threshold=2.0, stereo_regularity_multiplier=2.0,
                   stereo_speckle_filter=7)
        p = bm.matcher_params("mgm", 600)
        assert (p.census_win, p.ndir, p.lr_tau, p.P1, p.P2, p.timeout_ms) == (3, 4, 2.0, 8.0, 32.0, 600000)
        q = bm.matcher_params("mgm_multi", 5)
        assert (q.P1, q.P2, q.remove_small_cc, q.timeout_ms) == (16.0, 64.0, 7, 5000)
        r = bm.matcher_params("mgm_multi_lsd", 0)
        assert (r.P1, r.P2, r.remove_small_cc, r.median, r.timeout_ms) == (24.0, 96.0, 7, 1, 0)
    finally:
        cfg.clear()
        cfg.update(old)


def test_raster_roundtrip(tmp_path):
    from s2p_b200 import rasterio_compat as rio
    a = np.random.default_rng(0).normal(size=(17, 23)).astype(np.float32)
    a[3, 4] = np.nan
    p = str(tmp_path / "x.tif")
    rio.write_float_tiff(p, a)
    b = rio.read_band(p)
    assert b.dtype == np.float32 and np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])
    assert rio.image_size(p) == (23, 17)
    m = (a > 0).astype(np.uint8)
    q = str(tmp_path / "m.png")
    rio.write_mask_png(q, m)
    assert np.array_equal(rio.read_band(q).astype(np.uint8), m)


def test_multiband_raster_roundtrip(tmp_path):
    """the colour image of s2p/__init__.py:276 has several bands; band order survives the TIFF and PNG paths"""
    from s2p_b200 import rasterio_compat as rio
    rng = np.random.default_rng(0)
    bands = [rng.uniform(0, 255, (20, 30)).astype(np.float32) for _ in range(3)]
    bands[1][3, 4] = np.nan
    p = str(tmp_path / "c.tif")
    rio.write_float_tiff_bands(p, bands)
    assert rio.band_count(p) == 3 and rio.image_size(p) == (30, 20)
    for k in range(3):
        assert np.array_equal(rio.read_window(p, 2, 3, 10, 8, k + 1), bands[k][3:11, 2:12], equal_nan=True)
    assert np.array_equal(rio.read_band(p), bands[0])
    from PIL import Image
    rgb = rng.uniform(0, 255, (20, 30, 3)).astype(np.uint8)
    q = str(tmp_path / "c.png")
    Image.fromarray(rgb).save(q)
    assert rio.band_count(q) == 3
    assert np.array_equal(rio.read_window(q, 0, 0, 30, 20, 3), rgb[..., 2].astype(np.float32))


def test_synth_is_deterministic_and_shaped():
    from s2p_b200.synth import make_pair
    a, b, d = make_pair(40, 60, -8, 7, seed=3, nan_border=0.1)
    a2, b2, d2 = make_pair(40, 60, -8, 7, seed=3, nan_border=0.1)
    assert a.shape == (40, 60) and a.dtype == np.float32
    assert np.array_equal(np.isnan(a), np.isnan(a2)) and np.array_equal(d, d2)
    assert np.isnan(a).any() and np.isnan(b).any() and d.min() >= -8 and d.max() <= 7


def test_shard_covers_everything_once():
    from s2p_b200.tiles import shard
    for n, w in [(1521, 8), (7, 8), (64, 4), (0, 2), (5, 1)]:
        got = [i for r in range(w) for i in shard(n, r, w)]
        assert got == list(range(n))
        sizes = [len(shard(n, r, w)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1
    assert len(shard(1521, 0, 8)) == 191 and len(shard(1521, 7, 8)) == 190


def test_div3_trick_is_exact_on_a_dense_sample():
    """x/3 via two fmas (agg_kernel.cuh div3_exact) == IEEE division.  Exhaustive over all 2^31
    non-negative floats takes ~20 s of CPU (scripts/check_div3.c); here every 61st bit pattern plus
    all exponent boundaries."""
    src = os.path.join(ROOT, "scripts", "check_div3.c")
    exe = os.path.join(ROOT, "scripts", "check_div3")
    subprocess.run(["gcc", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-o", exe, src, "-lm"], check=True)
    out = subprocess.run([exe, "61"], check=True, capture_output=True, text=True).stdout
    assert "mismatches: 0" in out, out


def test_gloo_two_ranks_shard_and_gather(tmp_path):
    """world_size 2 on CPU: each rank takes its shard of a tile list, 'processes' it and the
    checksums are gathered on rank 0 (the N>1 path of bench.py, minus the GPU)."""
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys\n"
        "sys.path.insert(0, %r)\n"
        "import numpy as np, torch.distributed as dist\n"
        "from s2p_b200.tiles import shard, checksum, gather_checksums\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        "n = 11\n"
        "mine = shard(n, r, w)\n"
        "local = {i: checksum(np.full((4, 4), i, np.float32)) for i in mine}\n"
        "allc = gather_checksums(local, n, r, w)\n"
        "if r == 0:\n"
        "    want = [checksum(np.full((4, 4), i, np.float32)) for i in range(n)]\n"
        "    assert allc == want, (allc, want)\n"
        "    print('GATHER_OK', len(mine))\n"
        "dist.destroy_process_group()\n" % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29617")
    out = subprocess.run(["python", "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                          "127.0.0.1", "--master-port", "29617", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert "GATHER_OK 6" in out.stdout, out.stdout + out.stderr


def test_gloo_two_ranks_dynamic_queue(tmp_path):
    """world_size 2 on CPU: the dynamic tile queue of bench.py's strong-scaling configuration (BASELINE configs[3]) hands
    every tile to exactly one rank, also when one rank is slower."""
    script = tmp_path / "q.py"
    script.write_text(
        "import os, sys, time\n"
        "sys.path.insert(0, %r)\n"
        "import numpy as np, torch, torch.distributed as dist\n"
        "from s2p_b200.tiles import DynamicQueue\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        "n = 37\n"
        "q = DynamicQueue(n, 4, w)\n"
        "mine = []\n"
        "while True:\n"
        "    ids = q.next()\n"
        "    if not ids: break\n"
        "    mine += ids\n"
        "    time.sleep(0.02 if r == 0 else 0.002)\n"
        "t = torch.zeros(n, dtype=torch.int64)\n"
        "t[mine] = 1\n"
        "dist.all_reduce(t)\n"
        "assert bool((t == 1).all()), t\n"
        "print('QUEUE_OK rank %%d took %%d' %% (r, len(mine)))\n"
        "dist.destroy_process_group()\n" % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29618")
    out = subprocess.run(["python", "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                          "127.0.0.1", "--master-port", "29618", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert out.stdout.count("QUEUE_OK") == 2, out.stdout + out.stderr
    from s2p_b200.tiles import DynamicQueue
    q = DynamicQueue(10, 4)
    assert [q.next(), q.next(), q.next(), q.next()] == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9], []]


def test_public_header_is_plain_c(tmp_path):
    """include/s2pb200.h is the C ABI: it must compile as C99 on its own (no C++, no CUDA, no torch types)."""
    src = tmp_path / "hdr.c"
    src.write_text('#include "s2pb200.h"\nint main(void) { s2pb_mgm_params p; s2pb_rpc r; (void)p; (void)r; return S2PB_OK; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                        "-o", str(tmp_path / "hdr.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
