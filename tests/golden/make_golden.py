"""Regenerates tests/golden/*.npz: outputs of the UNMODIFIED reference binary oracle/_ref/mgm
(built by oracle/Makefile from /root/reference/3rdparty/mgm_multi) at OMP_NUM_THREADS=1 on seeded
synthetic pairs.  Inputs are not stored: they are regenerated from the seed by s2p_b200.synth.
Run from the repo root in the build container:   python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O                      # noqa: E402
from s2p_b200.synth import make_pair                # noqa: E402

# name -> (h, w, dmin, dmax, seed, nan_border, which images get the NaN strips, matcher overrides)
CASES = {
    "plain":      (64, 96, -12, 11, 101, 0.0, "", {}),
    "wide":       (40, 150, -64, 63, 102, 0.0, "", {}),
    "nan_ref":    (56, 88, -10, 9, 103, 0.07, "ref", {}),
    "tsgm4_o4":   (48, 80, -8, 12, 104, 0.0, "", {"tsgm": 4, "ndir": 4}),
    "census3":    (48, 80, -8, 12, 105, 0.0, "", {"census_win": 3}),
    "nan_both":   (56, 88, -10, 9, 106, 0.07, "both", {}),   # depends on the DCT round trip of shift(): oracle-only
    # `mgm_multi` with the flags of s2p/block_matching.py:269-308 (-S 6, SUBPIX=2, REMOVESMALLCC=25, TSGM=4)
    "multi":      (120, 168, -18, 21, 107, 0.0, "", {"_algo": "mgm_multi"}),
    "multi_s1":   (110, 150, -16, 12, 108, 0.0, "", {"_algo": "mgm_multi", "subpix": 1}),
    # the other distances of `-t` and the -wl / -wr regularity weights
    "ncc5":       (48, 80, -8, 12, 109, 0.0, "", {"cost": 3}),
    "ad_w":       (48, 80, -8, 12, 110, 0.0, "", {"cost": 1, "_weights": True, "P1": 12.0, "P2": 48.0}),
    "btsd_t4":    (40, 72, -9, 7, 111, 0.0, "", {"cost": 5, "tsgm": 4}),
    # what algo == 'mgm_multi_lsd' runs (s2p/block_matching.py:191-266): mgm_multi, P1=12, P2=48, MEDIAN=1, weights
    "lsd":        (120, 168, -18, 21, 112, 0.0, "", {"_algo": "mgm_multi", "_weights": True, "P1": 12.0, "P2": 48.0, "median": 1}),
}


def weights_for(name):
    """(wl, wr) for the cases that pass -wl / -wr, else None: mostly 1 with low values on a random 30 % of the
    pixels, the value range of the LSD segment maps s2p builds (`255 x - 255 / 2 pow 0.1 fmax`)."""
    h, w, _, _, seed, _, _, kw = CASES[name]
    if not kw.get("_weights"):
        return None
    rng = np.random.default_rng(seed + 1000)
    out = []
    for _ in range(2):
        x = rng.uniform(0, 255, (h, w))
        wt = np.maximum(((255 - x) / 255) ** 2, 0.1).astype(np.float32)
        wt[rng.random((h, w)) < 0.7] = 1.0
        out.append(wt)
    return tuple(out)


def split_kw(kw):
    """-> (algo, oracle-parameter overrides) of a case's kw (drops the '_' keys)."""
    kw = dict(kw)
    algo = kw.pop("_algo", "mgm")
    kw.pop("_weights", None)
    return algo, kw


def inputs(name):
    h, w, dmin, dmax, seed, nb, which, kw = CASES[name]
    ref, sec, _ = make_pair(h, w, dmin, dmax, seed=seed, nan_border=nb)
    if which == "ref":
        sec = make_pair(h, w, dmin, dmax, seed=seed)[1]
    return ref, sec, dmin, dmax, kw


def main():
    assert O.have_ref(), "build oracle/_ref first: make -C oracle ref"
    out = os.path.dirname(os.path.abspath(__file__))
    for name in CASES:
        ref, sec, dmin, dmax, kw = inputs(name)
        algo, kw = split_kw(kw)
        mk = O.mgm_multi_params if algo == "mgm_multi" else O.mgm_params
        wts = weights_for(name) or (None, None)
        r = O.run_ref(ref, sec, dmin, dmax, mk(dct_shift=1, **kw), threads=1, wl=wts[0], wr=wts[1])
        np.savez_compressed(os.path.join(out, name + ".npz"), disp=r["disp"], conf=r["conf"].astype(np.uint8),
                            dispR=r["dispR"])
        print(name, ref.shape, "valid %.3f" % np.isfinite(r["disp"]).mean())


if __name__ == "__main__":
    main()
