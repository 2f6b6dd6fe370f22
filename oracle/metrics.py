"""Oracle: SSIM / PSNR / fidelity exactly as the reference defines them.

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  Follows
``/root/reference/src/imgcompressionmps/utils/metrics.py:11-160``; the
scikit-image 0.24.0 ``structural_similarity`` call at ``metrics.py:32`` is
restated from the library's published algorithm -
**parity unpinned** for SSIM (scikit-image is absent here and the reference's
own metric tests never import the package's metrics).
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import uniform_filter

from . import mps as _mps


def structural_similarity(im1, im2, data_range, win_size=7):
    """skimage.metrics.structural_similarity, uniform-window branch (A.5)."""
    im1 = np.asarray(im1)
    im2 = np.asarray(im2)
    if im1.shape != im2.shape:
        raise ValueError("Input images must have the same dimensions.")
    if np.any(np.asarray(im1.shape) - win_size < 0):
        raise ValueError("win_size exceeds image extent.")
    if win_size % 2 != 1:
        raise ValueError("Window size must be odd.")
    ftype = np.float32 if (im1.dtype == np.float32 and im2.dtype == np.float32) else np.float64
    x = im1.astype(ftype, copy=False)
    y = im2.astype(ftype, copy=False)
    npix = win_size ** x.ndim
    cov_norm = npix / (npix - 1.0)
    ux = uniform_filter(x, size=win_size)
    uy = uniform_filter(y, size=win_size)
    uxx = uniform_filter(x * x, size=win_size)
    uyy = uniform_filter(y * y, size=win_size)
    uxy = uniform_filter(x * y, size=win_size)
    vx = cov_norm * (uxx - ux * ux)
    vy = cov_norm * (uyy - uy * uy)
    vxy = cov_norm * (uxy - ux * uy)
    c1 = (0.01 * data_range) ** 2
    c2 = (0.03 * data_range) ** 2
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
    pad = (win_size - 1) // 2
    inner = s[tuple(slice(pad, n - pad) for n in s.shape)]
    return float(inner.mean(dtype=np.float64))


def compute_ssim_2d(original, compressed):
    """``metrics.py:11-32``: clip the 2nd argument at 0, joint data range,
    window min(7, smallest side) forced odd."""
    original = np.asarray(original)
    compressed = np.clip(np.asarray(compressed), 0, None)
    data_range = max(original.max(), compressed.max()) - min(original.min(), compressed.min())
    win = min(7, min(original.shape))
    if win % 2 == 0:
        win -= 1
    return structural_similarity(original, compressed, data_range=data_range, win_size=win)


def ssim_3d_axis(original, compressed, axis=0):
    """``metrics.py:35-65`` (negative axes validate but yield an empty list)."""
    if original.shape != compressed.shape:
        raise ValueError("Shape mismatch between 3D arrays.")
    if axis >= original.ndim or axis < -original.ndim:
        raise ValueError(f"Invalid axis {axis} for 3D SSIM.")
    comp = np.clip(compressed, 0, None)
    out = []
    for i in range(original.shape[axis]):
        if axis == 0:
            out.append(compute_ssim_2d(original[i], comp[i]))
        elif axis == 1:
            out.append(compute_ssim_2d(original[:, i, :], comp[:, i, :]))
        elif axis == 2:
            out.append(compute_ssim_2d(original[:, :, i], comp[:, :, i]))
    return out


def avg_ssim_3d(original, compressed):
    """``metrics.py:68-85``."""
    if original.shape != compressed.shape:
        raise ValueError("Shape mismatch between 3D volumes.")
    return float(np.mean([np.mean(ssim_3d_axis(original, compressed, ax)) for ax in range(3)]))


def avg_ssim_4d(original, compressed):
    """``metrics.py:88-105``."""
    if original.shape != compressed.shape:
        raise ValueError("Shape mismatch between 4D volumes.")
    return float(np.mean([avg_ssim_3d(original[..., t], compressed[..., t]) for t in range(original.shape[-1])]))


def compute_ssim_by_dim(a, b):
    """``metrics.py:108-129``."""
    if a.ndim == 4:
        return avg_ssim_4d(a, b)
    if a.ndim == 3:
        return avg_ssim_3d(a, b)
    if a.ndim == 2:
        return compute_ssim_2d(a, b)
    raise ValueError(f"Unsupported tensor dimension for SSIM: {a.ndim}")


def compute_psnr(original, compressed):
    """``metrics.py:132-146``: peak is max(original) (sign-sensitive), inf at MSE 0."""
    original = np.asarray(original, dtype=np.float64)
    compressed = np.asarray(compressed, dtype=np.float64)
    mse = np.mean((original - compressed) ** 2)
    if mse == 0:
        return np.inf
    return float(10 * np.log10((np.max(original) ** 2) / mse))


def compute_overlap(cores1, norm1, cores2, norm2):
    """``metrics.py:149-160`` on core lists: <a|b> / (|a| |b|)."""
    return _mps.overlap(cores1, cores2) / (norm1 * norm2)
