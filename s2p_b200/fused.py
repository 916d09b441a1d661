"""Opt-in fusion of pipeline steps 3 and 4: rectify the pair and match it in ONE call, the
rectified images staying on the device between the two.

In the reference, `rectification_pair` (s2p/__init__.py:147-155) writes `rectified_ref.tif` / `rectified_sec.tif` through two
`homography` subprocesses, and `stereo_matching` (s2p/__init__.py:184-190) reads them back in another worker process to
run `mgm`.  The two steps run in different `multiprocessing.Pool`s (s2p/parallel.py:80), so device memory cannot be handed
from one to the other without changing s2p/__init__.py; the drop-ins of `common.py` and `block_matching.py` therefore keep the
file hand-off.  A pipeline that calls both steps from the same worker can use `rectify_and_match` instead: same files out
(rectified pair, disparity, confidence, mask), bit-identical to the two-step path, one TIFF decode and two host<->device
round trips fewer per tile.  INTEGRATION.md shows the ten lines of s2p/__init__.py this replaces.
"""
import subprocess

import numpy as np

from . import _lib, rasterio_compat as rio
from .block_matching import _NATIVE, confidence_path, disparity_bounds, matcher_params
from .common import needed_roi
from .engine import S2pbError, get_engine


def _crop(im, H, w, h):
    """the part of image file `im` the warp needs (s2p_b200.common.image_apply_homography does the same), crop-compensated H"""
    H = np.asarray(H, dtype=np.float64).reshape(3, 3)
    sw, sh = rio.image_size(im)
    x, y, rw, rh = needed_roi(H, w, h)
    if x < 0:
        rw += x
        x = 0
    if y < 0:
        rh += y
        y = 0
    rw, rh = min(rw, sw - x), min(rh, sh - y)
    if rw <= 0 or rh <= 0:
        raise subprocess.CalledProcessError(1, ["s2pb200:homography", im], output="ERROR: empty roi")
    return rio.read_window(im, x, y, rw, rh), H @ np.array([[1, 0, x], [0, 1, y], [0, 0, 1]], dtype=np.float64)


def rectify_and_match(out1, out2, disp, mask, im1, im2, H1, H2, w, h, algo, disp_min, disp_max, timeout=600, max_disp_range=None):
    """image_apply_homography(out1, im1, H1, w, h); image_apply_homography(out2, im2, H2, w, h);
    compute_disparity_map(out1, out2, disp, mask, algo, disp_min, disp_max, timeout, max_disp_range) -- in one device call."""
    if algo not in _NATIVE or algo == "mgm_multi_lsd":
        raise NotImplementedError("the fused path serves algo 'mgm' and 'mgm_multi'")
    disp_min, disp_max = disparity_bounds(w, disp_min, disp_max, max_disp_range)
    src1, H1c = _crop(im1, H1, w, h)
    src2, H2c = _crop(im2, H2, w, h)
    cmd = ["s2pb200:rectify+%s" % algo, im1, im2, disp]
    print("\nRUN: %s" % " ".join(cmd))
    try:
        out = get_engine().rectify_match(src1, H1c, src2, H2c, w, h, disp_min, disp_max, matcher_params(algo, timeout))
    except S2pbError as e:
        if e.code == _lib.ERR_TIMEOUT:
            raise subprocess.TimeoutExpired(cmd, timeout) from e
        raise subprocess.CalledProcessError(-e.code, cmd, output=str(e)) from e
    rio.write_float_tiff(out1, out["rect1"])
    rio.write_float_tiff(out2, out["rect2"])
    rio.write_float_tiff(disp, out["disp"])
    rio.write_float_tiff(confidence_path(disp, algo), out["conf"])
    rio.write_mask_png(mask, out["mask"])
