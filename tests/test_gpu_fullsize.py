"""Full BASELINE sizes on the B200 through size-independent properties: permutation round trip
and checksums, norm preservation, monotone fidelity in chi, left-canonical cores, reconstruction
error == discarded weight, batch sharding through run_sharded, DCT mode round trip.  The oracle
comparison at these sizes (15 s to 2 min of host time per case) lives in
``test_gpu_baseline_parity.py``."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def volume256():
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    from bench import synthetic_volume
    return torch.from_numpy(synthetic_volume((256, 256, 256), 2026)).cuda()


def test_permutation_properties_256(volume256):
    from imgcompressionmps import _ops
    dense = _ops.encode(volume256)
    assert list(dense.shape) == [8] * 8
    # a permutation: same multiset (sum, sum of squares, max) and exact inverse
    assert float(dense.double().sum()) == pytest.approx(float(volume256.double().sum()), rel=1e-12)
    assert float(dense.max()) == float(volume256.max())
    assert _ops.sumsq(dense) == pytest.approx(_ops.sumsq(volume256), rel=1e-13)
    assert torch.equal(_ops.decode(dense, (256, 256, 256)), volume256)
    # Morton structure: site l of voxel (x, y, z) is 4*x_l + 2*y_l + z_l (most significant bits first)
    idx = (37, 201, 150)
    sites = [(((idx[0] >> (7 - l)) & 1) << 2) | (((idx[1] >> (7 - l)) & 1) << 1) | ((idx[2] >> (7 - l)) & 1) for l in range(8)]
    assert float(dense[tuple(sites)]) == float(volume256[idx])
    # the gather kernel and the tiled kernel agree bit for bit
    from imgcompressionmps import _native
    ctx = _native.context()
    ctx.set_option("permute_path", 2)
    try:
        assert torch.equal(_ops.encode(volume256), dense)
    finally:
        ctx.set_option("permute_path", 0)


def test_chi_sweep_properties_256(volume256):
    """configs[1]: chi sweep 8..128 on the 256^3 volume."""
    from imgcompressionmps import _ops
    from imgcompressionmps.core.ndmps import NDMPS
    from imgcompressionmps.utils.metrics import compute_overlap, compute_psnr, compute_ssim_by_dim
    norm_x = math.sqrt(_ops.sumsq(volume256))
    full = None
    last_fid, last_psnr = 0.0, -1.0
    for chi in (8, 16, 32, 64, 128):
        obj = NDMPS.from_tensor(volume256, max_bond=chi)
        bonds = obj.bond_sizes()
        assert max(bonds) <= chi and bonds[0] == min(8, chi)
        # renorm=2 keeps the Frobenius norm of the state
        assert obj.norm_value == pytest.approx(norm_x, rel=1e-6)
        # left-canonical: every core but the last is an isometry
        for c in obj.mps.cores[:-1]:
            m = c.reshape(-1, c.shape[-1]).double()
            g = (m.T @ m).cpu().numpy()
            assert np.allclose(g, np.eye(g.shape[0]), atol=5e-6)
        rec = obj.to_tensor_device()
        if full is None:
            full = NDMPS.from_tensor(volume256, max_bond=128)
        fid = compute_overlap(obj, full)
        psnr = compute_psnr(rec, volume256)
        assert 0.0 < fid <= 1.0 + 1e-6 and fid >= last_fid - 1e-6 and psnr >= last_psnr - 1e-6
        last_fid, last_psnr = fid, psnr
        # <rec|x> / (|rec||x|) equals the cosine between reconstruction and data
        cos = float((rec.double() * volume256.double()).sum()) / (math.sqrt(_ops.sumsq(rec)) * norm_x)
        err = math.sqrt(_ops.psnr_terms(rec, volume256)[0]) / norm_x
        # both have the same Frobenius norm (renorm), so |rec - x|^2 = 2 |x|^2 (1 - cos)
        assert cos > 0.95 and err == pytest.approx(math.sqrt(max(2.0 * (1.0 - cos), 0.0)), abs=2e-3)
    ssim = compute_ssim_by_dim(rec, volume256)
    assert 0.5 < ssim <= 1.0


def test_sharded_batch_and_dct(volume256):
    from imgcompressionmps.distributed import compress_and_score, run_sharded
    vols = [volume256[96:160, 96:160, 96:160].contiguous() * (1.0 + 0.1 * i) for i in range(3)]      # central crop: the corners are background
    res = run_sharded(vols, lambda i, v: compress_and_score(i, v, max_bond=16))
    assert [r["index"] for r in res] == [0, 1, 2]
    assert all(r["bond_dims"] == res[0]["bond_dims"] for r in res)            # scaling the data does not change ranks
    assert all(abs(r["fidelity"] - res[0]["fidelity"]) < 1e-6 for r in res)
    from imgcompressionmps.core.ndmps import NDMPS
    v = volume256[64:192, 64:192, 80:176].contiguous()
    d = NDMPS.from_tensor(v, mode="DCT")
    rec = d.to_tensor_device()
    # the default rsum2 cutoff (1e-10 of the weight per bond) costs ~1e-5 relative per bond
    assert float(torch.linalg.vector_norm(rec - v) / torch.linalg.vector_norm(v)) < 1e-4


def test_fmri_like_4d():
    """configs[3] shape family (one subject, fewer frames): 4-D encode, truncation, 4-D SSIM."""
    from imgcompressionmps.core.ndmps import NDMPS
    from imgcompressionmps.utils.metrics import compute_ssim_by_dim
    gx, gy, gz = torch.meshgrid(torch.linspace(-1, 1, 64), torch.linspace(-1, 1, 64), torch.linspace(-1, 1, 32), indexing="ij")
    base = (torch.exp(-2.0 * (gx ** 2 + 1.5 * gy ** 2 + 0.7 * gz ** 2)) + 0.2 * torch.cos(3 * gx) * torch.cos(2 * gy)).cuda()
    t = torch.linspace(0, 6.28, 100, device="cuda").reshape(1, 1, 1, -1)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = (base[..., None] * (1.0 + 0.05 * torch.sin(t) + 0.03 * torch.cos(3 * t) * gx.cuda()[..., None])
         + 0.005 * torch.rand((64, 64, 32, 100), device="cuda", generator=g)).float().contiguous()
    errs = []
    for chi in (8, 32):
        obj = NDMPS.from_tensor(x, max_bond=chi)
        assert int(np.prod(obj.mps.site_dims)) == x.numel() and max(obj.bond_sizes()) <= chi
        rec = obj.to_tensor_device()
        errs.append(float(torch.linalg.vector_norm(rec - x) / torch.linalg.vector_norm(x)))
    print("fmri-like rel errors", errs)
    assert errs[1] < errs[0] < 0.5 and errs[1] < 0.1
    s = compute_ssim_by_dim(rec, x)
    assert 0.3 < s <= 1.0
