"""How far do two legitimate builds of the REFERENCE differ on tiles with no-data in both images?

The reference sends the matched image through a DCT -> iDCT even for a shift of 0 (mgm_costvolume.cc:23-60).  On pixels that
are exactly 0 (no-data, NaN -> 0, main_mgm.cc:172-173) the round trip leaves rounding noise of either sign instead of 0, and
the census transform then compares noise with noise.  The result depends on the summation order of the DCT, i.e. on the fftw
build.  This script runs the unmodified reference matcher linked against three DCTs that differ only in summation order
(oracle/_ref/mgm, mgm_alt1, mgm_alt2; oracle/Makefile) and the identity the engine uses (the oracle port with dct_shift=0),
and prints the pairwise spread.  CPU only.  usage: python scripts/nodata_spread.py [out.md]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import oracle as O
from s2p_b200.synth import make_pair

CASES = [("nan_both golden (56x88, 7 % strips)", 56, 88, -10, 9, 106, 0.07),
         ("200x300, 5 % strips", 200, 300, -24, 23, 7, 0.05),
         ("256x384, 5 % strips", 256, 384, -32, 31, 11, 0.05),
         ("200x300, 2 % strips", 200, 300, -24, 23, 13, 0.02)]


def stats(a, b):
    """(pixels that differ, of them by > 0.25 px, NaN-mask mismatches)"""
    na, nb = np.isnan(a), np.isnan(b)
    both = ~na & ~nb
    d = np.abs(np.where(both, a - b, 0))
    return int(((a != b) & ~(na & nb)).sum()), int((d > 0.25).sum()), int((na != nb).sum())


def main():
    rows = []
    for name, h, w, dmin, dmax, seed, nb in CASES:
        ref, sec, _ = make_pair(h, w, dmin, dmax, seed=seed, nan_border=nb)
        p = O.mgm_params()
        runs = {b: O.run_ref(ref, sec, dmin, dmax, p, binary=b) for b in ("mgm", "mgm_alt1", "mgm_alt2")}
        d, c, dr = O.port.mgm(ref, sec, dmin, dmax, O.mgm_params(dct_shift=0))
        runs["identity"] = dict(disp=d, conf=c, dispR=dr)
        n = h * w
        for a, b in (("mgm", "mgm_alt1"), ("mgm", "mgm_alt2"), ("mgm_alt1", "mgm_alt2"), ("mgm", "identity"), ("mgm_alt1", "identity"), ("mgm_alt2", "identity")):
            sl = stats(runs[a]["disp"], runs[b]["disp"])
            sr = stats(runs[a]["dispR"], runs[b]["dispR"])
            cf = int((runs[a]["conf"] != runs[b]["conf"]).sum())
            rows.append((name, n, a, b, sl, sr, cf))
    lines = ["| tile | pair of builds | left disp differs | > 0.25 px | NaN mask | right disp differs | > 0.25 px | NaN mask | confidence differs |",
             "|---|---|---|---|---|---|---|---|---|"]
    for name, n, a, b, sl, sr, cf in rows:
        lines.append("| %s | %s vs %s | %d (%.2f %%) | %d | %d | %d (%.2f %%) | %d | %d | %d (%.2f %%) |" % (
            name, a, b, sl[0], 100. * sl[0] / n, sl[1], sl[2], sr[0], 100. * sr[0] / n, sr[1], sr[2], cf, 100. * cf / n))
    txt = "\n".join(lines)
    print(txt)
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            f.write("# Reference-vs-reference spread on no-data tiles (scripts/nodata_spread.py)\n\n"
                    "`mgm` = unmodified reference + table DCT; `mgm_alt1` = same products summed from the far end; `mgm_alt2` = 80-bit\n"
                    "accumulator; `identity` = what the engine computes (no round trip).  Counts are pixels of the tile.\n\n" + txt + "\n")


if __name__ == "__main__":
    main()
