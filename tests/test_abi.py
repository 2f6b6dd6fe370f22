"""The C-ABI library loads without a GPU, exports every symbol include/ndmps.h
declares, and its host-side permutation plan reproduces the reference's encoding map.
No device work happens here."""
import re
from pathlib import Path

import numpy as np
import pytest

from imgcompressionmps import _native as N
from imgcompressionmps.utils.core import gen_encoding_map, get_factorlist

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    text = (ROOT / "include" / "ndmps.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ndmps_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    declared = _declared()
    assert len(declared) >= 25
    assert sorted(N.PROTOTYPES) == declared


def test_library_exports_every_declared_symbol():
    lib = N.load_library()
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.ndmps_version() >= 100
    assert isinstance(lib.ndmps_last_error(), bytes)


def test_bad_arguments_report_errors():
    with pytest.raises(ValueError):
        N.Plan((4, 6), np.array([[2, 2], [2, 2]]))          # factors do not multiply to the extents
    with pytest.raises(ValueError):
        N.Plan((4, 0), np.array([[4, 1]]))


@pytest.mark.parametrize("shape", [(8, 9), (4, 6), (12, 18, 10), (16, 16, 8, 20), (30, 40, 50), (7,), (1, 4), (64, 64)])
def test_plan_offsets_match_encoding_map(shape):
    """Destination->source offsets the kernels will use == the reference's map."""
    factors, _ = get_factorlist(shape)
    plan = N.Plan(shape, factors)
    dims, enc = gen_encoding_map(shape)
    assert plan.site_dims == [int(d) for d in dims]
    total = int(np.prod(shape))
    # flat site index of every voxel (volume order)
    site_flat = np.zeros(shape, dtype=np.int64)
    for lvl in range(len(dims)):
        site_flat = site_flat * int(dims[lvl]) + enc[lvl]
    site_flat = site_flat.reshape(-1)
    enc_src = plan.debug_offsets(False, 0, total)            # dst = site order, src = volume offset
    assert np.array_equal(site_flat[enc_src], np.arange(total))
    dec_src = plan.debug_offsets(True, 0, total)             # dst = volume order, src = site offset
    assert np.array_equal(dec_src, site_flat)
    # partial ranges
    if total > 10:
        assert np.array_equal(plan.debug_offsets(False, 5, 4), enc_src[5:9])


def test_compute_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from imgcompressionmps.core.ndmps import NDMPS
    with pytest.raises(RuntimeError):
        NDMPS.from_tensor(np.zeros((8, 8), dtype=np.float32))


@pytest.mark.parametrize("shape", [(8, 9), (12, 18, 10), (16, 16, 8, 20), (30, 40, 50), (256, 128), (512, 680), (64, 64, 64),
                                   (32, 16, 8, 20), (6, 10, 15)])
def test_tiled_permutation_tables_match_oracle(shape):
    """The shared-memory tiling (source runs -> skewed slots -> destination runs) reproduces the
    oracle's permutation when its tables are walked on the host exactly as the kernel walks them."""
    from oracle import encoding as OE
    factors, _ = get_factorlist(shape)
    plan = N.Plan(shape, factors)
    ramp = np.arange(int(np.prod(shape)), dtype=np.int32).reshape(shape)
    want = OE.encode(ramp).reshape(-1)
    enc, dec = plan.tile_info(False), plan.tile_info(True)
    assert enc["tiled"] and dec["tiled"]
    assert enc["tile"] <= 8192 and enc["tile"] * enc["tiles"] == ramp.size
    assert np.array_equal(plan.apply_tiled_host(False, ramp), want)
    assert np.array_equal(plan.apply_tiled_host(True, want), ramp.reshape(-1))
