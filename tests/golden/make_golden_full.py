"""Regenerates tests/golden/full_c2.json: the UNMODIFIED reference binary (oracle/_ref/mgm, OMP_NUM_THREADS=1) on full
BASELINE configs[1] tiles -- 1024 x 1024, 128 labels, s2p's `mgm` flags -- one without and one with 5 % no-data strips in
both images.  The rasters (12 MB per tile) are not stored: the golden is one blake2b digest per block of 32 rows of the
disparity, confidence and right-disparity rasters, plus the count of valid pixels per block, so that a mismatch can be
located.  Inputs are regenerated from the seed by s2p_b200.synth.  ~4 minutes per tile on one core.
    python tests/golden/make_golden_full.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CASES = {"c2_plain": dict(h=1024, w=1024, dmin=-64, dmax=63, seed=0, nan_border=0.0),
         "c2_nodata": dict(h=1024, w=1024, dmin=-64, dmax=63, seed=1, nan_border=0.05)}
BLOCK = 32


def digests(a):
    """-> list of hex digests, one per block of BLOCK rows (NaN payloads normalised)"""
    a = np.ascontiguousarray(a, dtype=np.float32)
    a = np.where(np.isnan(a), np.float32(np.nan), a)
    return [hashlib.blake2b(a[r:r + BLOCK].tobytes(), digest_size=8).hexdigest() for r in range(0, a.shape[0], BLOCK)]


def main():
    from oracle import oracle as O
    from s2p_b200.synth import make_pair
    assert O.have_ref(), "build oracle/_ref first: make -C oracle ref"
    out = {}
    for name, c in CASES.items():
        ref, sec, _ = make_pair(c["h"], c["w"], c["dmin"], c["dmax"], seed=c["seed"], nan_border=c["nan_border"])
        r = O.run_ref(ref, sec, c["dmin"], c["dmax"], O.mgm_params(), threads=1)
        out[name] = dict(c, block_rows=BLOCK, seconds=round(r["seconds"], 1),
                         disp=digests(r["disp"]), conf=digests(r["conf"]), dispR=digests(r["dispR"]),
                         valid=[int(np.isfinite(r["disp"][k:k + BLOCK]).sum()) for k in range(0, c["h"], BLOCK)])
        print(name, "%.0f s, valid %.3f" % (r["seconds"], np.isfinite(r["disp"]).mean()), flush=True)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "full_c2.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
