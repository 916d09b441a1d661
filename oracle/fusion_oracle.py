"""TEST INFRASTRUCTURE ONLY -- numpy restatement of s2p.fusion.merge_n's arithmetic (s2p/fusion.py:16-68).

`merge_ref` runs the REFERENCE's own `average_if_close`, extracted from /root/reference/s2p/fusion.py when that
tree is present (the module itself cannot be imported here: it needs rasterio), through np.apply_along_axis
exactly as merge_n does; `merge_port` is the vectorised restatement used where the reference is absent.
"""
import ast
import os
import warnings

import numpy as np

REF = "/root/reference/s2p/fusion.py"


def reference_average_if_close():
    """-> the reference's function object, or None when /root/reference is absent."""
    if not os.path.exists(REF):
        return None
    tree = ast.parse(open(REF).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "average_if_close"][0]
    ns = {"np": np}
    exec(compile(ast.Module([fn], []), REF, "exec"), ns)
    return ns["average_if_close"]


def merge_ref(rasters, offsets, threshold):
    f = reference_average_if_close()
    x = np.stack([np.asarray(r, np.float32).astype(np.float64) - o for r, o in zip(rasters, offsets)], axis=2)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        avg = np.apply_along_axis(f, 2, x, threshold)
    avg += np.mean(offsets)
    return avg.astype("float32")


def merge_port(rasters, offsets, averaging="average_if_close", threshold=1):
    x = np.stack([np.asarray(r, np.float32).astype(np.float64) - o for r, o in zip(rasters, offsets)], axis=2)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if averaging == "average_if_close":
            spread = np.nanmax(x, axis=2) - np.nanmin(x, axis=2)
            avg = np.where(spread > threshold, np.nan, np.nanmedian(x, axis=2))
        else:
            avg = getattr(np, averaging.split(".")[1])(x, axis=2)
    avg = avg + np.mean(offsets)
    return avg.astype("float32")
