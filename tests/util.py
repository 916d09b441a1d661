import numpy as np


def same(a, b):
    """bit-equality with NaN == NaN"""
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype.kind == "f":
        return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))
    return bool(np.array_equal(a, b))


def nmismatch(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype.kind == "f":
        return int((~((a == b) | (np.isnan(a) & np.isnan(b)))).sum())
    return int((a != b).sum())
