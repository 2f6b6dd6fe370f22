"""Quantisation helpers with the reference's names and semantics
(``utils/filetools.py:7-39``).  These host versions operate on numpy arrays, as in
the reference; ``NDMPS.compress_to_dtype`` uses the device kernels
(``ndmps_quantize`` / ``ndmps_dequantize``) on the cores instead.
The reference's file-system helpers (``filetools.py:42-123``) are out of scope.
"""
from __future__ import annotations

import numpy as np


def get_num_bits(dtype) -> int:
    dtype = np.dtype(dtype)
    if np.issubdtype(dtype, np.integer):
        return np.iinfo(dtype).bits
    if np.issubdtype(dtype, np.floating):
        return np.finfo(dtype).bits
    raise ValueError(f"Unsupported dtype {dtype!r}")


def scale_to_dtype(array: np.ndarray, dtype=np.uint8) -> np.ndarray:
    """Min-max normalise to [0, iinfo(dtype).max]; the cast truncates."""
    shifted = array - np.min(array)
    return (shifted / np.max(shifted) * np.iinfo(dtype).max).astype(dtype)


def scale_back(array: np.ndarray, arr_min: float, arr_max: float, dtype=np.uint8) -> np.ndarray:
    return array / np.iinfo(dtype).max * (arr_max - arr_min) + arr_min
