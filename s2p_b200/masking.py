"""Drop-in for ``s2p.masking.erosion`` (s2p/masking.py:87-97).

The reference erodes the rejection mask in place right after the matcher (s2p/__init__.py:190) by spawning
``morsi diskR erosion``; here it is one small kernel through ``s2pb_erode_mask``.  Same signature, same file
contract (8-bit PNG, 0 = rejected), and as in the reference nothing happens for a radius below 2.
"""
from . import rasterio_compat as rio
from .engine import get_engine


def erosion(out, msk, radius):
    if radius >= 2:
        print("\nRUN: s2pb200:morsi disk%d erosion %s %s" % (int(radius), msk, out))
        m = rio.read_band(msk)
        rio.write_mask_png(out, get_engine().erode_mask((m != 0).astype("uint8") if m.max() > 1 else m.astype("uint8"), int(radius)))


def install():
    import s2p.masking as original
    if not hasattr(original, "_s2pb_original_erosion"):
        original._s2pb_original_erosion = original.erosion
    original.erosion = erosion
    return original
