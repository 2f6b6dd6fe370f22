"""Oracle: min-max quantisation helpers (``/root/reference/src/imgcompressionmps/utils/filetools.py:7-39``).

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  PINNED: the reference's
``filetools.py`` is importable here and was executed to produce
``tests/golden/quantise.npz``.
"""
from __future__ import annotations

import numpy as np


def get_num_bits(dtype):
    """``filetools.py:7-17``."""
    dtype = np.dtype(dtype)
    if np.issubdtype(dtype, np.integer):
        return int(np.iinfo(dtype).bits)
    if np.issubdtype(dtype, np.floating):
        return int(np.finfo(dtype).bits)
    raise ValueError(f"Unsupported dtype {dtype!r}")


def scale_to_dtype(array, dtype=np.uint8):
    """``filetools.py:20-26``: (x - min) / max(x - min) * iinfo.max, cast by truncation."""
    shifted = np.asarray(array) - np.min(array)
    unit = shifted / np.max(shifted)
    return (unit * np.iinfo(dtype).max).astype(dtype)


def scale_back(array, arr_min, arr_max, dtype=np.uint8):
    """``filetools.py:29-39``."""
    unit = np.asarray(array) / np.iinfo(dtype).max
    return unit * (arr_max - arr_min) + arr_min
