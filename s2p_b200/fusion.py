"""Drop-in for ``s2p.fusion.merge_n`` (s2p/fusion.py:25-68).

Same signature and file contract: the n input rasters are read, merged on the GPU
(``s2pb_merge_n``: per pixel in float64, ``average_if_close`` or a nan-aware numpy reducer, plus
the mean offset) and the result replaces the band of a copy of the first input, so that the
geo-referencing metadata is kept exactly as the reference keeps it.
"""
import shutil

from . import rasterio_compat as rio
from .engine import get_engine


def merge_n(output, inputs, offsets, averaging="average_if_close", threshold=1):
    assert len(inputs) == len(offsets)
    if not inputs:
        return
    rasters = [rio.read_band(p) for p in inputs]
    avg = get_engine().merge_n(rasters, offsets, averaging, threshold)
    shutil.copy(inputs[0], output)          # keeps the metadata of the first input (s2p/fusion.py:64)
    rio.overwrite_band(output, avg)


def install():
    import s2p.fusion as original
    if not hasattr(original, "_s2pb_original_merge_n"):
        original._s2pb_original_merge_n = original.merge_n
    original.merge_n = merge_n
    return original
