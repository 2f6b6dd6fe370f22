"""SSIM / PSNR / fidelity with the reference's names, argument order and error
behaviour (``utils/metrics.py:11-160``), computed on the B200.

Arguments may be numpy arrays (uploaded) or CUDA torch tensors (used in place).
The arithmetic is float64 on the device whatever the storage dtype, which is what
the reference computes (its inputs are float64).  Note the reference's order
sensitivity, kept here: the SECOND argument is the one clipped at 0, and PSNR's
peak is ``max`` of the FIRST argument.

``compute_mean_std`` (``metrics.py:163-202``) is host-side curve aggregation on result
dictionaries (numpy only).
"""
from __future__ import annotations

from typing import List

import numpy as np

from .. import _ops


def _torch():
    import torch
    return torch


def _pair_on_device(a, b):
    """Two same-shape arrays -> contiguous CUDA tensors of one float dtype."""
    torch = _torch()

    def up(x):
        if isinstance(x, torch.Tensor):
            t = x if x.is_cuda else x.cuda()
        else:
            arr = np.asarray(x)
            if arr.dtype not in (np.float32, np.float64):
                arr = arr.astype(np.float64)
            t = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
        if t.dtype not in (torch.float32, torch.float64):
            t = t.to(torch.float64)
        return t.contiguous()

    ta, tb = up(a), up(b)
    if ta.dtype != tb.dtype:
        ta, tb = ta.to(torch.float64), tb.to(torch.float64)
    return ta, tb


def _shape(x):
    return tuple(int(s) for s in x.shape)


def compute_ssim_2d(original, compressed) -> float:
    """SSIM of two 2-D arrays: second argument clipped at 0, joint data range, window
    min(7, smallest side) forced odd (``metrics.py:11-32``)."""
    if _shape(original) != _shape(compressed):
        raise ValueError("Input images must have the same dimensions.")
    a, b = _pair_on_device(original, compressed)
    if a.ndim != 2:
        raise ValueError(f"compute_ssim_2d expects 2-D arrays, got {a.ndim}-D")
    return _ops.ssim(a, b)


def ssim_3d_axis(original, compressed, axis: int = 0) -> List[float]:
    """Slice-wise SSIM scores along one axis of a 3-D volume (``metrics.py:35-65``)."""
    if _shape(original) != _shape(compressed):
        raise ValueError("Shape mismatch between 3D arrays.")
    ndim = len(_shape(original))
    if axis >= ndim or axis < -ndim:
        raise ValueError(f"Invalid axis {axis} for 3D SSIM.")
    if axis < 0:
        return []          # the reference validates negative axes but never branches on them
    a, b = _pair_on_device(original, compressed)
    return [float(v) for v in _ops.ssim_slices(a, b, axis)]


def avg_ssim_3d(original, compressed) -> float:
    """Mean over the three axes of the mean slice SSIM (``metrics.py:68-85``)."""
    if _shape(original) != _shape(compressed):
        raise ValueError("Shape mismatch between 3D volumes.")
    a, b = _pair_on_device(original, compressed)
    return _ops.ssim(a, b)


def avg_ssim_4d(original, compressed) -> float:
    """Mean over last-axis frames of the 3-D average SSIM (``metrics.py:88-105``)."""
    if _shape(original) != _shape(compressed):
        raise ValueError("Shape mismatch between 4D volumes.")
    a, b = _pair_on_device(original, compressed)
    return _ops.ssim(a, b)


def compute_ssim_by_dim(a, b) -> float:
    """Dispatch on dimensionality (``metrics.py:108-129``)."""
    ndim = len(_shape(a))
    if ndim == 4:
        return avg_ssim_4d(a, b)
    if ndim == 3:
        return avg_ssim_3d(a, b)
    if ndim == 2:
        return compute_ssim_2d(a, b)
    raise ValueError(f"Unsupported tensor dimension for SSIM: {ndim}")


def compute_psnr(original, compressed) -> float:
    """10 log10(max(original)^2 / MSE); inf at MSE 0 (``metrics.py:132-146``)."""
    a, b = _pair_on_device(original, compressed)
    if a.shape != b.shape:
        raise ValueError("Shape mismatch between PSNR operands.")
    sq, peak = _ops.psnr_terms(a, b)
    mse = sq / a.numel()
    if mse == 0:
        return np.inf
    return float(10 * np.log10((peak ** 2) / mse))


def compute_overlap(mps1, mps2) -> float:
    """Normalised overlap <a|b> / (|a| |b|) of two NDMPS objects (``metrics.py:149-160``)."""
    return (mps1.mps @ mps2.mps) / (mps1.norm_value * mps2.norm_value)


def _is_prime(n: int) -> bool:
    n = int(n)
    if n < 2:
        return False
    if n < 4:
        return True
    if n % 2 == 0:
        return False
    f = 3
    while f * f <= n:
        if n % f == 0:
            return False
        f += 2
    return True


def compute_mean_std(dict, num_common_points, key_x="compressionratio_list_disk", key_y="ssim_list"):
    """Mean and standard deviation of the ``key_y`` curves over a common compression-FACTOR grid
    (``metrics.py:163-202``): x is ``1 / dict[key_x]`` per sample, the grid spans the range every sample
    covers (from the largest first factor to the smallest last factor), samples whose shape consists of
    primes only are left out, values outside a sample's own range are NaN (linear interpolation without
    extrapolation).  Returns ``(mean, std, grid)``; ``(nan, nan, grid)`` when no sample qualifies."""
    ratios = np.array(dict[key_x])
    values = np.array(dict[key_y])
    shapes = np.array(dict["shapes"])
    factors = 1 / ratios
    grid = np.linspace(np.max(factors[:, 0]), np.min(factors[:, -1]), num_common_points)
    curves = []
    for x, y, shape in zip(factors, values, shapes):
        if all(_is_prime(extent) for extent in np.atleast_1d(shape)):
            continue
        order = np.argsort(x, kind="stable")
        curves.append(np.interp(grid, x[order], y[order], left=np.nan, right=np.nan))
    if not curves:
        return np.nan, np.nan, grid
    return np.mean(curves, axis=0), np.std(curves, axis=0), np.array(grid)
