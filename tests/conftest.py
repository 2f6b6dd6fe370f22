"""Test configuration: import paths, the ``gpu`` marker, shared fixtures.

``-m "not gpu"``: oracle vs the reference's golden vectors, host-side planning, C-ABI
symbol table (no device work).  ``-m gpu``: parity of the CUDA path against the oracle,
called through the C ABI.  Nothing here reads /root/reference at run time.
"""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
PKG = ROOT / "img-compression-mps_b200"
for p in (str(ROOT), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) device; run with -m gpu")


@pytest.fixture(scope="session")
def golden_encoding():
    return np.load(GOLDEN / "encoding.npz")


@pytest.fixture(scope="session")
def golden_quantise():
    return np.load(GOLDEN / "quantise.npz")


@pytest.fixture(scope="session")
def golden_dct():
    return np.load(GOLDEN / "dct.npz")


def phantom(shape, seed, noise=0.02, background=0.0):
    """Synthetic 'MRI': nested ellipsoids, smoothed edges, exact-zero background, small noise
    inside the object only."""
    rng = np.random.default_rng(seed)
    grids = np.meshgrid(*[np.linspace(-1, 1, n) for n in shape], indexing="ij")
    vol = np.zeros(shape, dtype=np.float64)
    for k in range(4):
        centre = rng.uniform(-0.25, 0.25, size=len(shape))
        radii = rng.uniform(0.35, 0.8, size=len(shape)) * (1.0 - 0.18 * k)
        r2 = sum(((g - c) / r) ** 2 for g, c, r in zip(grids, centre, radii))
        vol += (0.25 + 0.1 * k) / (1.0 + np.exp(np.clip((r2 - 1.0) * 12.0, -60.0, 60.0)))
    vol += 0.05 * np.prod([np.cos(3.0 * g + k) for k, g in enumerate(grids)], axis=0) * (vol > 0.05)
    vol += noise * rng.random(shape) * (vol > 0.05)
    vol[vol < 0.02] = 0.0
    if background:
        # faint texture everywhere: no slice is constant, so no slice has data range 0 (where the
        # reference's SSIM is 0/0)
        vol += background * rng.random(shape)
    return vol
