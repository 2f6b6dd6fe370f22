"""Quantisation helpers with the reference's names and semantics
(``utils/filetools.py:7-39``).  These host versions operate on numpy arrays, as in
the reference; ``NDMPS.compress_to_dtype`` uses the device kernels
(``ndmps_quantize`` / ``ndmps_dequantize``) on the cores instead.
The reference's data-set helpers (``filetools.py:42-123``: central slices, file discovery, result
merging, shapes) follow below with the same names, arguments and return values; they are host-side
bookkeeping around the device path.
"""
from __future__ import annotations

import json
import os
from pathlib import Path

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


# ---------------------------------------------------------------------------------------------
# data-set helpers (host side)
# ---------------------------------------------------------------------------------------------
def mri_to_slices(data_list, bitsize_list=None):
    """The three central 2-D slices (one per axis) of every 3-D volume (``filetools.py:42-71``).
    Non-3-D entries are skipped with a message; a missing ``bitsize_list`` means 16 bits each.
    Always returns ``(slices, bits)``."""
    slices, bits = [], []
    for index, volume in enumerate(data_list):
        if volume.ndim != 3:
            print(f"Skipping non-3D volume at index {index} with shape {volume.shape}")
            continue
        centre = [n // 2 for n in volume.shape]
        slices += [volume[centre[0]], volume[:, centre[1]], volume[:, :, centre[2]]]
        bits += [bitsize_list[index] if bitsize_list else 16] * 3
    return slices, bits


def find_project_root(marker: str = "src") -> Path:
    """First ancestor of this file that contains ``marker`` (``filetools.py:74-83``)."""
    here = Path(__file__).resolve()
    for candidate in (here, *here.parents):
        if (candidate / marker).exists():
            return candidate
    raise FileNotFoundError(f"Could not find directory containing '{marker}'")


def find_specific_files(directory_path, file_extension=None):
    """Every file below ``directory_path`` (``os.walk`` order), optionally only those ending in
    ``file_extension`` (``filetools.py:86-95``)."""
    return [os.path.join(root, name)
            for root, _, names in os.walk(directory_path)
            for name in names
            if file_extension is None or name.endswith(file_extension)]


def combine_jsons(input_file_1, input_file_2, output_file):
    """Merge two result files (``filetools.py:98-116``): list-valued entries are concatenated, except
    ``cutoff_list`` / ``mode`` and non-list entries, which come from the first file."""
    with open(input_file_1, "r") as f:
        first = json.load(f)
    with open(input_file_2, "r") as f:
        second = json.load(f)
    merged = {key: value if key in ("cutoff_list", "mode") or not isinstance(value, list) else value + second[key]
              for key, value in first.items()}
    with open(output_file, "w") as f:
        json.dump(merged, f, indent=4)
    print("Combined JSON created successfully!")


def get_shapes(data_list):
    """Shape of every array of the list (``filetools.py:119-123``)."""
    return [np.shape(data) for data in data_list]
