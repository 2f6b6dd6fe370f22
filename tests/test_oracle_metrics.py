"""Oracle metrics, DCT and quantisation: definitions restated from
utils/metrics.py:11-160 / SURVEY Appendix A.5-A.6, pinned where the reference's own
code could be executed (DCT via scipy.fftpack, quantise via utils/filetools.py)."""
import numpy as np
import pytest
from scipy.fft import dct, idct

from oracle import metrics as OM
from oracle import quantise as OQ


def _ssim_loops(x, y, R, w=7):
    """Direct definition, no filter library: window means over every interior pixel."""
    p = (w - 1) // 2
    n = w * w
    cov = n / (n - 1.0)
    c1, c2 = (0.01 * R) ** 2, (0.03 * R) ** 2
    vals = []
    for i in range(p, x.shape[0] - p):
        for j in range(p, x.shape[1] - p):
            a = x[i - p:i + p + 1, j - p:j + p + 1]
            b = y[i - p:i + p + 1, j - p:j + p + 1]
            ux, uy = a.mean(), b.mean()
            vx = cov * ((a * a).mean() - ux * ux)
            vy = cov * ((b * b).mean() - uy * uy)
            vxy = cov * ((a * b).mean() - ux * uy)
            vals.append(((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2)))
    return float(np.mean(vals))


def test_structural_similarity_matches_definition():
    rng = np.random.default_rng(0)
    x = rng.random((19, 23))
    y = x + 0.1 * rng.standard_normal(x.shape)
    R = max(x.max(), y.max()) - min(x.min(), y.min())
    assert abs(OM.structural_similarity(x, y, R) - _ssim_loops(x, y, R)) < 1e-12
    assert abs(OM.structural_similarity(x, y, R, win_size=5) - _ssim_loops(x, y, R, 5)) < 1e-12


def test_ssim_identity_and_errors():
    rng = np.random.default_rng(1)
    x = rng.random((16, 16))
    assert OM.compute_ssim_2d(x, x) == pytest.approx(1.0)
    with pytest.raises(ValueError):
        OM.structural_similarity(x, x[:8], 1.0)
    with pytest.raises(ValueError):
        OM.ssim_3d_axis(np.zeros((4, 4, 4)), np.zeros((4, 4, 5)))
    with pytest.raises(ValueError):
        OM.ssim_3d_axis(np.zeros((8, 8, 8)), np.zeros((8, 8, 8)), axis=3)
    with pytest.raises(ValueError):
        OM.compute_ssim_by_dim(np.zeros(5), np.zeros(5))
    assert OM.ssim_3d_axis(rng.random((8, 8, 8)), rng.random((8, 8, 8)), axis=-1) == []


def test_clip_and_argument_order():
    rng = np.random.default_rng(2)
    a = rng.random((12, 14))
    b = a - 0.5
    # only the second argument is clipped at 0 (metrics.py:23)
    assert OM.compute_ssim_2d(a, b) == pytest.approx(OM.compute_ssim_2d(a, np.clip(b, 0, None)))
    assert OM.compute_ssim_2d(b, a) != pytest.approx(OM.compute_ssim_2d(np.clip(b, 0, None), a))


def test_small_window():
    rng = np.random.default_rng(3)
    a, b = rng.random((5, 30)), rng.random((5, 30))
    R = max(a.max(), b.max()) - min(a.min(), b.min())
    assert OM.compute_ssim_2d(a, b) == pytest.approx(_ssim_loops(a, b, R, 5))
    a, b = rng.random((6, 30)), rng.random((6, 30))        # min(7, 6) = 6 -> forced odd -> 5
    R = max(a.max(), b.max()) - min(a.min(), b.min())
    assert OM.compute_ssim_2d(a, b) == pytest.approx(_ssim_loops(a, b, R, 5))


def test_3d_4d_composition():
    rng = np.random.default_rng(4)
    a = rng.random((9, 10, 11))
    b = np.clip(a + 0.05 * rng.standard_normal(a.shape), -0.1, None)
    per_axis = [np.mean(OM.ssim_3d_axis(a, b, ax)) for ax in range(3)]
    assert OM.avg_ssim_3d(a, b) == pytest.approx(np.mean(per_axis))
    a4 = rng.random((8, 9, 10, 3))
    b4 = a4 + 0.05 * rng.standard_normal(a4.shape)
    assert OM.avg_ssim_4d(a4, b4) == pytest.approx(np.mean([OM.avg_ssim_3d(a4[..., t], b4[..., t]) for t in range(3)]))
    assert OM.compute_ssim_by_dim(a4, b4) == OM.avg_ssim_4d(a4, b4)


def test_psnr():
    a = np.array([[0.0, 1.0], [2.0, 4.0]])
    assert OM.compute_psnr(a, a) == np.inf
    b = a + 0.5
    assert OM.compute_psnr(a, b) == pytest.approx(10 * np.log10(16.0 / 0.25))
    # peak is max of the FIRST argument
    assert OM.compute_psnr(b, a) == pytest.approx(10 * np.log10(4.5 ** 2 / 0.25))


def test_dct_matches_scipy_fftpack_fixture(golden_dct):
    g = golden_dct
    for i in range(5):
        x = g[f"in{i}"]
        assert np.allclose(dct(x, type=2, axis=-1, norm="ortho"), g[f"dct{i}"], atol=1e-13)
        assert np.allclose(idct(x, type=2, axis=-1, norm="ortho"), g[f"idct{i}"], atol=1e-13)
        n = x.shape[-1]
        k = np.arange(n)[:, None]
        c = np.sqrt(2.0 / n) * np.cos(np.pi * (2 * np.arange(n)[None, :] + 1) * k / (2 * n))
        c[0] /= np.sqrt(2.0)
        assert np.allclose(x @ c.T, g[f"dct{i}"], atol=1e-12)
        assert np.allclose(x @ c, g[f"idct{i}"], atol=1e-12)


def test_quantise_fixture(golden_quantise):
    g = golden_quantise
    assert [OQ.get_num_bits(d) for d in (np.uint8, np.uint16, np.int32, np.float32, np.float64)] == list(g["bits"])
    for i in range(3):
        a = g[f"in{i}"]
        for dt in (np.uint8, np.uint16):
            name = np.dtype(dt).name
            q = OQ.scale_to_dtype(a, dt)
            assert np.array_equal(q, g[f"q{i}_{name}"])
            assert np.array_equal(OQ.scale_back(q, a.min(), a.max(), dt), g[f"back{i}_{name}"])
