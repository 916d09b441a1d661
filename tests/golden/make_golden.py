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
}


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
        kw = dict(kw)
        mk = O.mgm_multi_params if kw.pop("_algo", "mgm") == "mgm_multi" else O.mgm_params
        r = O.run_ref(ref, sec, dmin, dmax, mk(dct_shift=1, **kw), threads=1)
        np.savez_compressed(os.path.join(out, name + ".npz"), disp=r["disp"], conf=r["conf"].astype(np.uint8),
                            dispR=r["dispR"])
        print(name, ref.shape, "valid %.3f" % np.isfinite(r["disp"]).mean())


if __name__ == "__main__":
    main()
