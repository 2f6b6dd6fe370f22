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


@pytest.mark.parametrize("shape", [(64, 64, 64), (128, 128, 128), (256, 256), (256, 128), (512, 512), (32, 32, 32),
                                   (16, 16, 16, 16), (1024, 16), (4096, 4), (32, 16, 8, 32), (64, 64, 32, 32)])
def test_bit_permutation_plan_matches_oracle(shape):
    """Power-of-two shapes take the register bit-permutation kernel: its thread -> offset function and its 2 x 2
    register transposition, walked on the host exactly as the kernel runs them, reproduce the oracle's permutation."""
    from oracle import encoding as OE
    factors, _ = get_factorlist(shape)
    plan = N.Plan(shape, factors)
    ramp = np.arange(int(np.prod(shape)), dtype=np.int32).reshape(shape)
    want = OE.encode(ramp).reshape(-1)
    for inverse in (False, True):
        info = plan.bit_info(inverse)
        assert info["bits"] and (1 << info["nbits"]) == ramp.size and info["ctas"] * 8192 == ramp.size
        assert info["dst_run"] >= 32 and info["src_run"] >= 32          # whole 128-byte lines on both sides
    assert np.array_equal(plan.apply_bits_host(False, ramp), want)
    assert np.array_equal(plan.apply_bits_host(True, want), ramp.reshape(-1))


@pytest.mark.parametrize("shape", [(8, 9), (30, 40, 50), (512, 680), (8, 8, 8), (16, 1024), (64, 64, 32, 400)])
def test_bit_permutation_plan_declines_other_shapes(shape):
    """Non-power-of-two factors, volumes below one CTA's 8192 elements and low-bit patterns the register
    transposition does not cover keep the tiled kernel."""
    factors, _ = get_factorlist(shape)
    plan = N.Plan(shape, factors)
    assert not plan.bit_info(False)["bits"] and not plan.bit_info(True)["bits"]
    assert plan.tile_info(False)["tiled"]


def test_bit_permutation_plan_random_power_of_two_shapes():
    """400 random power-of-two shapes (2-D ... 5-D, 2^13 ... 2^21 elements, axes of extent 1 and 2 included): wherever the
    planner accepts a shape, the kernel's thread -> offset function and register transposition reproduce the oracle's
    permutation in both directions, both directions are accepted together, and every warp covers whole 128-byte lines."""
    from oracle import encoding as OE
    rng = np.random.default_rng(123)
    seen, accepted = set(), 0
    while len(seen) < 400:
        exps = rng.integers(0, 9, size=int(rng.integers(2, 6)))
        shape = tuple(int(2 ** e) for e in exps)
        if not 13 <= exps.sum() <= 21 or shape in seen:
            continue
        seen.add(shape)
        factors, _ = get_factorlist(shape)
        plan = N.Plan(shape, factors)
        enc, dec = plan.bit_info(False), plan.bit_info(True)
        assert enc["bits"] == dec["bits"], shape
        if not enc["bits"]:
            continue
        accepted += 1
        ramp = np.arange(int(np.prod(shape)), dtype=np.int32).reshape(shape)
        want = OE.encode(ramp).reshape(-1)
        assert np.array_equal(plan.apply_bits_host(False, ramp), want), shape
        assert np.array_equal(plan.apply_bits_host(True, want), ramp.reshape(-1)), shape
        assert min(enc["dst_run"], enc["src_run"], dec["dst_run"], dec["src_run"]) >= 32, shape
    assert accepted >= 50
