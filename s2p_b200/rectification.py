"""Boundary #2: ``s2p.rectification.rectify_pair``.

Everything in ``rectify_pair`` except the two image warps is 3x3 algebra on at most a few hundred points
(s2p/rectification.py:281-378: virtual matches from the RPCs, affine fundamental matrix, rectifying
similarities, disparity range) and needs s2p's own geometry modules (rpcm, estimation, rpc_utils); it stays
the reference's numpy code.  Only ``common.image_apply_homography`` (:379-380) is heavy, and
``s2p_b200.common`` replaces it.  ``rectify_pair`` below therefore is the reference's function running with the
B200 warp installed; it exists so that callers can import the boundary from one place.
"""
from . import common


def rectify_pair(im1, im2, rpc1, rpc2, x, y, w, h, out1, out2, A=None, sift_matches=None, method="rpc",
                 hmargin=0, vmargin=0):
    """Same signature and return value (H1, H2, disp_min, disp_max) as s2p.rectification.rectify_pair."""
    try:
        import s2p.rectification as original
    except Exception as e:  # pragma: no cover - depends on the installation
        raise NotImplementedError("rectify_pair's geometry lives in the s2p package (rpcm, estimation, rpc_utils), "
                                  "which is not importable here; s2p_b200.common.image_apply_homography is the "
                                  "part this engine replaces") from e
    common.install()
    return original.rectify_pair(im1, im2, rpc1, rpc2, x, y, w, h, out1, out2, A, sift_matches, method, hmargin, vmargin)


def install():
    """Patch an importable s2p: both boundaries (matcher and warp) and the two "next" rows already served
    (mask erosion, n-view merge) go to the B200 engine."""
    from . import block_matching, fusion, masking, triangulation
    common.install()
    block_matching.install()
    masking.install()
    fusion.install()
    triangulation.install()
