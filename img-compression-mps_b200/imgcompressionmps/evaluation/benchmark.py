"""The reference's benchmark module on device-resident data.

Mirrors ``evaluation/benchmark.py`` with the same names, arguments, result keys and post-processing:
``load_tensors`` (:16-55, in ``loader.py``), ``conv_to_mps`` / ``conv_to_tensors`` (:58-100),
``compress_list`` (:103-118), ``benchmark_metric`` (:121-146), ``run_benchmark`` (:149-194) and
``run_full_benchmark`` (:197-242, same JSON keys), so the reference's callers (``main.py``, the
notebook) can import them unchanged.  What differs is how the work
is issued: the original tensors are put on the device once, every (tensor, cutoff) pair is
reconstructed ONCE and that reconstruction serves both SSIM and PSNR (the reference calls
``to_tensor()`` per metric), nothing but scalars comes back to the host, and the tensors of a
list are independent, so they run several at a time (``batch.VolumePipeline``); uploads are
prefetched through pinned buffers (``loader.prefetch_to_device``).  Plotting stays out.

The module keeps the reference module's GLOBAL names as well (``nib``, ``NDMPS``, ``get_num_bits``, ``find_specific_files``,
``get_shapes``, ``mri_to_slices``, ``compute_ssim_by_dim``, ``compute_psnr``, ``compute_overlap``) and every function reaches
them through the module, so code that replaces them - the reference's own ``tests/evaluation/test_benchmark.py`` does, with
a stand-in ``NDMPS`` class - gets the reference's plain loop on whatever objects it supplies; the device path is taken
when the items are this package's ``NDMPS`` objects."""
from __future__ import annotations

import json
from copy import deepcopy
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np

from ..core.ndmps import NDMPS
from ..utils.filetools import find_specific_files, get_num_bits, get_shapes, mri_to_slices
from ..utils.metrics import compute_overlap, compute_psnr, compute_ssim_by_dim
from . import loader as _loader
from .loader import conv_to_mps_streamed

nib = _loader.nifti_module()             # nibabel when installed, else the NIfTI-1 reader of loader.py behind ``nib.load``
_DeviceNDMPS = NDMPS                     # the class whose objects take the device path below

_METRICS = ("ssim", "compression_ratio", "bond_dims", "psnr", "fidelity", "storage", "gzip_bytes", "gzip_ratio")


def load_tensors(files, ending, shape=None):
    """``(data_list, bitsize_list)`` of a list of ``.npz`` (key ``"sequence"``) or NIfTI ``.gz`` files, optionally
    cropped to ``shape = (B, H, W)`` (``benchmark.py:16-55``)."""
    return _loader.load_tensors(files, ending, shape, nib=nib, num_bits=get_num_bits)


def conv_to_mps(data_list, mode="DCT", *, max_bond=None, cutoff: float = 1e-10):
    """``NDMPS.from_tensor(data, norm=False, mode=mode)`` for every tensor (``benchmark.py:58-79``), one progress line
    each; with this package's class the host-to-device copies run ahead of the encoding."""
    if len(data_list) == 0:
        return []
    if NDMPS is _DeviceNDMPS:
        return conv_to_mps_streamed(data_list, mode, max_bond=max_bond, cutoff=cutoff)
    extras = {}
    if max_bond is not None:
        extras["max_bond"] = max_bond
    if cutoff != 1e-10:
        extras["cutoff"] = cutoff
    mps_list = []
    for index, data in enumerate(data_list):
        print(f"Converting file {index + 1}/{len(data_list)}")
        mps_list.append(NDMPS.from_tensor(data, norm=False, mode=mode, **extras))
    return mps_list


def conv_to_tensors(mps_list):
    """``mps.to_tensor()`` for every item (``benchmark.py:82-100``)."""
    data_list = []
    for index, mps in enumerate(mps_list):
        print(f"Converting file {index + 1}/{len(mps_list)}")
        data_list.append(mps.to_tensor())
    return data_list


def compress_list(mps_list, compression_factors):
    """``mps.compress(compression_factors)`` for every item, in place (``benchmark.py:103-118``)."""
    if compression_factors is None:
        raise ValueError("compression_factors must not be None")
    for mps in mps_list:
        mps.compress(compression_factors)


def _reconstruction(mps):
    """On the device when the object can keep it there, else ``to_tensor()`` as the reference calls it."""
    return mps.to_tensor_device() if hasattr(mps, "to_tensor_device") else mps.to_tensor()


def _one_metric(mps, ref, metric: str, dtype, rec=None):
    if metric == "compression_ratio":
        return mps.compression_ratio()
    if metric == "storage":
        return mps.get_storage_space(dtype)
    if metric == "gzip_bytes":
        return mps.get_bytesize_on_disk(dtype=dtype)
    if metric == "gzip_ratio":
        return mps.compression_ratio_on_disk(dtype=dtype, replace=True)       # replace=True as the reference does
    if metric == "ssim":
        return compute_ssim_by_dim(_reconstruction(mps) if rec is None else rec, ref)
    if metric == "psnr":
        return compute_psnr(_reconstruction(mps) if rec is None else rec, ref)
    if metric == "bond_dims":
        return mps.bond_sizes()
    if metric == "shape":
        return ref.shape
    if metric == "fidelity":
        return compute_overlap(mps, ref)
    raise ValueError(f"Unsupported metric: {metric}")


def benchmark_metric(mps_list, reference_list=None, metric: str = "compression_ratio", dtype=np.uint16):
    """One metric for every item of ``mps_list`` (``benchmark.py:121-146``)."""
    if metric not in _METRICS + ("shape",):
        raise ValueError(f"Unsupported metric: {metric}")
    if reference_list and len(reference_list) != len(mps_list):
        raise IndexError("Length mismatch: reference_list and mps_list must have the same length.")
    return [_one_metric(mps, reference_list[i] if reference_list else None, metric, dtype) for i, mps in enumerate(mps_list)]


def _to_device(x):
    import torch
    if isinstance(x, torch.Tensor):
        return x if x.is_cuda else x.cuda()
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _level(item, dtype):
    """All metrics of one tensor at the current compression level; one reconstruction shared by SSIM and PSNR."""
    mps, original, reference_mps = item
    rec = mps.to_tensor_device()
    out = {"ssim": float(compute_ssim_by_dim(rec, original)), "psnr": float(compute_psnr(rec, original))}
    del rec
    for name in ("compression_ratio", "bond_dims", "fidelity", "storage", "gzip_bytes", "gzip_ratio"):   # the reference's order
        out[name] = _one_metric(mps, reference_mps if name == "fidelity" else None, name, dtype)
    return out


def run_benchmark(mps_list, original_tensors_list, cutoff_list, *, workers: Optional[int] = None, dtype=np.uint16) -> Dict:
    """Metrics of every tensor before compression and after each cutoff of ``cutoff_list`` (cumulative, in place),
    in the reference's result layout (``benchmark.py:149-194``): per metric an array ``[tensor][level]``
    (``bond_dims`` stays ``[level][tensor]``)."""
    if len(original_tensors_list) != len(mps_list):
        raise IndexError("Length mismatch: reference_list and mps_list must have the same length.")
    if not (len(mps_list) and all(isinstance(m, _DeviceNDMPS) for m in mps_list)):
        return _run_benchmark_plain(mps_list, original_tensors_list, cutoff_list, dtype)
    originals = [_to_device(t) for t in original_tensors_list]
    original_mps_list = deepcopy(mps_list)
    results: Dict[str, List] = {name: [] for name in _METRICS}
    pipe = None
    if workers is None:
        workers = min(4, len(mps_list))
    if workers > 1:
        from ..batch import VolumePipeline
        pipe = VolumePipeline(workers=workers)
    try:
        items = list(zip(mps_list, originals, original_mps_list))

        def measure():
            rows = pipe.map(lambda it: _level(it, dtype), items) if pipe else [_level(it, dtype) for it in items]
            for name in _METRICS:
                results[name].append([row[name] for row in rows])

        measure()
        for i, cutoff in enumerate(cutoff_list):
            print(f"Status: {100 * (i + 1) / len(cutoff_list):.2f}% - Cutoff: {cutoff}")
            if pipe:
                pipe.map(lambda mps: mps.compress(cutoff), mps_list)
            else:
                compress_list(mps_list, cutoff)
            measure()
    finally:
        if pipe:
            pipe.close()
    return {name: _layout(name, levels) for name, levels in results.items()}


def _run_benchmark_plain(mps_list, original_tensors_list, cutoff_list, dtype):
    """The reference's own loop (``benchmark.py:149-194``) for objects that are not this package's ``NDMPS`` (a stand-in
    class, another implementation of the same interface) and for empty lists: every metric through ``benchmark_metric``
    before compression and after each cutoff, same order of calls, same result layout."""
    original_mps_list = deepcopy(mps_list)
    references = {"ssim": original_tensors_list, "psnr": original_tensors_list, "fidelity": original_mps_list}
    results: Dict[str, List] = {name: [] for name in _METRICS}

    def measure():
        for name in _METRICS:
            results[name].append(benchmark_metric(mps_list, references.get(name), metric=name, dtype=dtype))

    measure()
    for i, cutoff in enumerate(cutoff_list):
        print(f"Status: {100 * (i + 1) / len(cutoff_list):.2f}% - Cutoff: {cutoff}")
        compress_list(mps_list, cutoff)
        measure()
    return {name: _layout(name, levels) for name, levels in results.items()}


def _layout(name: str, levels: List[List]):
    """Result layout of the reference: scalar metrics become an array indexed [tensor][level]; the bond
    dimensions (ragged lists) stay [level][tensor]; an empty tensor list gives an empty result."""
    if len(levels) == 0 or len(levels[0]) == 0:
        return []
    if name == "bond_dims":
        return levels
    return np.asarray(levels).transpose()


def run_full_benchmark(dataset_path, cutoff_list, result_file, datatype="MRI", mode="DCT", start=0, end=-1, ending=".gz",
                       shape=None):
    """Data set -> result JSON (``benchmark.py:197-242``): find the files, load (and optionally crop / slice) them,
    build the MPS list, run the cutoff sweep and write ``datatype, mode, files, cutoff_list, bitsize_list, shapes``
    plus one list per metric.  Relative result paths land under ``src/evaluation/results`` as in the reference."""
    result_path = Path(result_file)
    if not result_path.is_absolute() and not str(result_path).startswith("src/evaluation/results"):
        result_path = Path("src/evaluation/results") / result_path
    files = find_specific_files(Path(dataset_path), ending)
    files = files[start:] if end == -1 else files[start:end]
    if not files:
        raise FileNotFoundError(f"No files with extension {ending} found in {dataset_path}")
    data_list, bitsize_list = load_tensors(files, ending, shape)
    if datatype == "MRI_Slice":
        data_list, bitsize_list = mri_to_slices(data_list, bitsize_list)
    mps_list = conv_to_mps(data_list, mode)
    print("Starting benchmark...")
    metrics = run_benchmark(mps_list, data_list, cutoff_list)
    print(f"Saving results to {result_path}")
    result_dict = {"datatype": datatype, "mode": mode, "files": files, "cutoff_list": np.asarray(cutoff_list).tolist(),
                   "bitsize_list": bitsize_list, "shapes": get_shapes(data_list)}
    result_dict.update({key: value.tolist() if hasattr(value, "tolist") else value for key, value in metrics.items()})
    result_path.parent.mkdir(parents=True, exist_ok=True)
    with open(result_path, "w") as f:
        json.dump(result_dict, f, indent=2)
