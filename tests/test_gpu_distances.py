"""GPU parity of the matcher's wider option surface: the other distances of the
reference's `-t` table (ad, sd, ncc, btad, btsd; mgm_costvolume.h:186-197) and the -wl / -wr regularity weights
that algo == 'mgm_multi_lsd' passes (s2p/block_matching.py:191-266).  These run through the "general" flavour
of the kernels (float32 cost slab, per-pixel weights).  Everything is compared bit for bit with the CPU oracle,
which is itself pinned against the reference binary for each of these options (tests/test_oracle.py), and with
golden outputs of that binary."""
import os
import sys

import numpy as np
import pytest

from s2p_b200.synth import make_pair
from util import nmismatch, same

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))


def _weights(shape, seed, ones=0.6):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 255, shape)
    w = np.maximum(((255 - x) / 255) ** 2, 0.1).astype(np.float32)
    w[rng.random(shape) < ones] = 1.0
    return w


@pytest.mark.parametrize("cost,win", [("ad", 5), ("sd", 5), ("btad", 5), ("btsd", 5), ("ncc", 3), ("ncc", 5), ("ncc", 7),
                                      ("census", 5), ("census", 3), ("census", 7)])
def test_costvolume_distances(engine, oracle, cost, win):
    h, w, dmin, dmax = 31, 120, -14, 21
    ref, sec, _ = make_pair(h, w, dmin, dmax, seed=3)
    lo = np.full((h, w), dmin, np.int32)
    hi = np.full((h, w), dmax, np.int32)
    lo[5:9, 10:40] = dmin + 1            # ragged per-pixel ranges
    hi[5:9, 10:40] = dmin + 3
    D = dmax - dmin + 1
    Cg = engine.costvolume(ref, sec, lo, hi, dmin, D, win, cost=cost)
    Co = oracle.port.costvolume(ref, sec, lo, hi, dmin, D, win, cost=cost)
    assert same(Cg, Co), "%d voxels differ" % nmismatch(Cg, Co)


@pytest.mark.parametrize("tsgm", [1, 2, 3, 4])
@pytest.mark.parametrize("ndir,weighted", [(8, True), (4, True), (8, False)])
def test_aggregate_general(engine, oracle, tsgm, ndir, weighted):
    """float costs that f16 cannot hold, penalties whose products with the weights are inexact (12, 48)"""
    h, w, dmin, dmax = 45, 70, -10, 13
    ref, sec, _ = make_pair(h, w, dmin, dmax, seed=5)
    lo = np.full((h, w), dmin, np.int32)
    hi = np.full((h, w), dmax, np.int32)
    lo[20:24, 30:50] = dmin
    hi[20:24, 30:50] = dmin + 1
    D = dmax - dmin + 1
    C = oracle.port.costvolume(ref, sec, lo, hi, dmin, D, 5, cost="btad")
    wgt = _weights((h, w), 11) if weighted else None
    So, do, co, fo = oracle.port.aggregate(C, lo, hi, dmin, 12.0, 48.0, ndir, tsgm, weights=wgt)
    Sg, dg, cg, fg = engine.aggregate(C, lo, hi, dmin, 12.0, 48.0, ndir, tsgm, weights=wgt, general=True)
    assert same(dg, do), "integer WTA index differs at %d pixels" % nmismatch(dg, do)
    assert same(fg, fo), "consensus differs at %d pixels" % nmismatch(fg, fo)
    assert same(Sg, So), "aggregated volume differs at %d voxels" % nmismatch(Sg, So)
    assert same(cg, co)


@pytest.mark.parametrize("cost,kw", [("ad", {}), ("sd", {"tsgm": 4}), ("ncc", {}), ("ncc", {"census_win": 7, "ndir": 4}),
                                     ("btad", {"tsgm": 2}), ("btsd", {"refine": 2})])
def test_mgm_distances(engine, oracle, cost, kw):
    from s2p_b200.engine import default_params
    h, w, dmin, dmax = 50, 132, -40, 30           # D = 71 -> 3 labels per lane
    ref, sec, _ = make_pair(h, w, dmin, dmax, seed=31)
    out = engine.mgm(ref, sec, dmin, dmax, default_params("mgm", cost=cost, **kw), want_right=True)
    d, c, dr = oracle.port.mgm(ref, sec, dmin, dmax, oracle.mgm_params(cost=oracle.COSTS.index(cost), **kw))
    assert same(out["disp"], d), "disparity differs at %d px" % nmismatch(out["disp"], d)
    assert same(out["conf"], c)
    assert same(out["disp_right"], dr)


@pytest.mark.parametrize("shape,dmin,dmax,cost,weighted", [
    ((30, 200), -90, 70, "ad", False),      # D=161 -> 6 labels per lane, barrier every second step
    ((24, 260), -128, 127, "btad", True),   # D=256 -> 8
    ((20, 400), -150, 150, "census", True), # D=301 -> 12
    ((18, 520), -250, 249, "sd", True),     # D=500 -> 16
])
def test_general_flavour_wide_volumes(engine, oracle, shape, dmin, dmax, cost, weighted):
    from s2p_b200.engine import default_params
    h, w = shape
    ref, sec, _ = make_pair(h, w, dmin, dmax, seed=91)
    wts = (_weights((h, w), 7), _weights((h, w), 8)) if weighted else None
    out = engine.mgm(ref, sec, dmin, dmax, default_params("mgm", cost=cost, P1=12.0, P2=48.0), want_right=True, weights=wts)
    d, c, dr = oracle.port.mgm(ref, sec, dmin, dmax, oracle.mgm_params(cost=oracle.COSTS.index(cost), P1=12.0, P2=48.0),
                               *(wts or (None, None)))
    assert same(out["disp"], d), "disparity differs at %d px" % nmismatch(out["disp"], d)
    assert same(out["conf"], c)
    assert same(out["disp_right"], dr)


@pytest.mark.parametrize("kw", [dict(), dict(P1=12.0, P2=48.0), dict(P1=12.0, P2=48.0, tsgm=4), dict(P1=5.5, P2=41.0, tsgm=2, ndir=4),
                                dict(P1=12.0, P2=48.0, cost="ad"), dict(census_win=7, P1=12.0, P2=48.0)])
def test_mgm_weighted(engine, oracle, kw):
    from s2p_b200.engine import default_params
    h, w, dmin, dmax = 57, 100, -12, 17
    ref, sec, _ = make_pair(h, w, dmin, dmax, seed=41)
    wl, wr = _weights((h, w), 1), _weights((h, w), 2)
    okw = dict(kw)
    if "cost" in okw:
        okw["cost"] = oracle.COSTS.index(okw["cost"])
    out = engine.mgm(ref, sec, dmin, dmax, default_params("mgm", **kw), want_right=True, weights=(wl, wr))
    d, c, dr = oracle.port.mgm(ref, sec, dmin, dmax, oracle.mgm_params(**okw), wl, wr)
    assert same(out["disp"], d), "disparity differs at %d px" % nmismatch(out["disp"], d)
    assert same(out["conf"], c)
    assert same(out["disp_right"], dr)
    # unit weights are the unweighted matcher, bit for bit
    ones = np.ones((h, w), np.float32)
    a = engine.mgm(ref, sec, dmin, dmax, default_params("mgm", **kw), weights=(ones, ones))
    b = engine.mgm(ref, sec, dmin, dmax, default_params("mgm", **kw))
    assert same(a["disp"], b["disp"]) and same(a["conf"], b["conf"])


def test_mgm_multi_lsd(engine, oracle):
    """mgm_multi with the flags and weights of algo == 'mgm_multi_lsd'; ZOOM=1 levels bit-exact, see test_mgm_multi"""
    from s2p_b200.engine import default_params
    h, w, dmin, dmax = 130, 190, -21, 16
    ref, sec, _ = make_pair(h, w, dmin, dmax, seed=51)
    wl, wr = _weights((h, w), 3, 0.75), _weights((h, w), 4, 0.75)
    for subpix in (1, 2):
        out = engine.mgm(ref, sec, dmin, dmax, default_params("mgm_multi_lsd", subpix=subpix), want_right=True, weights=(wl, wr))
        d, c, dr = oracle.port.mgm_multi(ref, sec, dmin, dmax,
                                         oracle.mgm_multi_params(P1=12.0, P2=48.0, median=1, subpix=subpix), wl, wr)
        assert same(out["conf"], c), "consensus differs at %d px" % nmismatch(out["conf"], c)
        if subpix == 1:
            assert same(out["disp"], d), "disparity differs at %d px" % nmismatch(out["disp"], d)
            assert same(out["disp_right"], dr)
        else:
            both = np.isfinite(d) & np.isfinite(out["disp"])
            assert (np.isnan(d) != np.isnan(out["disp"])).mean() < 2e-3
            assert (np.abs(d[both] - out["disp"][both]) > 0.25).mean() < 2e-3


def test_mgm_multi_distance(engine, oracle):
    from s2p_b200.engine import default_params
    h, w, dmin, dmax = 120, 170, -15, 18
    ref, sec, _ = make_pair(h, w, dmin, dmax, seed=61)
    out = engine.mgm(ref, sec, dmin, dmax, default_params("mgm_multi", subpix=1, cost="ad"), want_right=True)
    d, c, dr = oracle.port.mgm_multi(ref, sec, dmin, dmax, oracle.mgm_multi_params(subpix=1, cost=1))
    assert same(out["disp"], d), "disparity differs at %d px" % nmismatch(out["disp"], d)
    assert same(out["conf"], c)
    assert same(out["disp_right"], dr)


@pytest.mark.parametrize("name", ["ncc5", "ad_w", "btsd_t4", "lsd"])
def test_against_reference_golden_vectors(engine, name):
    """tests/golden/*.npz are outputs of the unmodified reference binary (tests/golden/make_golden.py)."""
    import make_golden as G
    from s2p_b200.engine import default_params
    ref, sec, dmin, dmax, kw = G.inputs(name)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    algo, kw = G.split_kw(kw)
    out = engine.mgm(ref, sec, dmin, dmax, default_params(algo, **kw), want_right=True, weights=G.weights_for(name))
    assert np.array_equal(out["conf"].astype(np.uint8), g["conf"])
    if name == "lsd":        # half-pixel pass: held to the tolerance (see test_gpu_parity.py::test_mgm_multi)
        both = np.isfinite(g["disp"]) & np.isfinite(out["disp"])
        assert (np.isnan(g["disp"]) != np.isnan(out["disp"])).mean() < 2e-3
        assert (np.abs(g["disp"][both] - out["disp"][both]) > 0.25).mean() < 2e-3
        return
    assert same(out["disp"], g["disp"]), "%d px differ from the reference" % nmismatch(out["disp"], g["disp"])
    assert same(out["disp_right"], g["dispR"])


def test_dropin_mgm_multi_lsd(engine, oracle, tmp_path, monkeypatch):
    """compute_disparity_map(algo='mgm_multi_lsd'): the file contract of s2p/block_matching.py:191-266 (the
    confidence goes to disp + '.confidence.tif'); the LSD weight maps come from a stand-in for the host pipeline."""
    from s2p_b200 import block_matching as bm, rasterio_compat as rio
    h, w, dmin, dmax = 110, 160, -14, 11
    ref, sec, _ = make_pair(h, w, dmin, dmax, seed=71)
    im1, im2 = str(tmp_path / "rectified_ref.tif"), str(tmp_path / "rectified_sec.tif")
    disp, mask = str(tmp_path / "rectified_disp.tif"), str(tmp_path / "rectified_mask.png")
    rio.write_float_tiff(im1, ref)
    rio.write_float_tiff(im2, sec)
    wmap = {im1: _weights((h, w), 5, 0.8), im2: _weights((h, w), 6, 0.8)}
    monkeypatch.setattr(bm, "lsd_weight_map", lambda path: wmap[path])
    assert bm.compute_disparity_map(im1, im2, disp, mask, "mgm_multi_lsd", dmin, dmax) is None
    d, c, _ = oracle.port.mgm_multi(ref, sec, dmin, dmax, oracle.mgm_multi_params(P1=12.0, P2=48.0, median=1),
                                    wmap[im1], wmap[im2])
    assert same(rio.read_band(disp + ".confidence.tif"), c)
    got = rio.read_band(disp)
    both = np.isfinite(d) & np.isfinite(got)
    assert (np.isnan(d) != np.isnan(got)).mean() < 2e-3 and (np.abs(d[both] - got[both]) > 0.25).mean() < 2e-3
    assert set(np.unique(rio.read_band(mask))) <= {0, 1}
