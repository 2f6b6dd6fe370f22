"""Generate tests/golden/*.npz by EXECUTING the reference's own importable modules.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference's ``utils/core.py`` and ``utils/filetools.py`` import with numpy +
sympy; ``core/ndmps.py`` and ``utils/metrics.py`` do not (quimb / scikit-image
absent), so only encoding, quantise and scipy's DCT are pinned this way.
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, "/root/reference/src")
from imgcompressionmps.utils import core as ref_core            # noqa: E402
from imgcompressionmps.utils import filetools as ref_ft         # noqa: E402
from scipy.fftpack import dct, idct                              # noqa: E402  (what core/ndmps.py:5 imports)

OUT = Path(__file__).resolve().parent

SMALL = [(8, 9), (4, 6), (3, 3), (1, 4), (7,), (8,), (12, 18, 10), (6, 10, 15), (4, 4, 4, 4), (16, 16, 8, 20)]
LARGE = [(256, 128), (30, 40, 50), (512, 680), (8, 512, 680), (256, 256), (64, 64, 64), (64, 64, 32, 400),
         (1920, 1080, 64)]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def scatter_encode(tensor):
    """core/ndmps.py:57-71 verbatim in behaviour: scatter through the reference's map."""
    qubit_size, enc = ref_core.gen_encoding_map(tuple(int(s) for s in tensor.shape))
    enc = np.moveaxis(enc, 0, -1)
    out = np.empty(tuple(qubit_size), dtype=tensor.dtype)
    k = enc.shape[-1]
    flat = enc.reshape(-1, k).astype(int)
    out[tuple(flat[:, d] for d in range(k))] = tensor.flatten()
    return out


def main():
    enc = {}
    for shape in SMALL + LARGE:
        key = "x".join(map(str, shape))
        fac, prod = ref_core.get_factorlist(shape)
        enc[f"{key}/factors"] = fac
        enc[f"{key}/prod"] = prod
        n = int(np.prod(shape))
        if n <= 2_800_000 * 4 and shape not in [(64, 64, 32, 400), (1920, 1080, 64)]:
            q, m = ref_core.gen_encoding_map(shape)
            enc[f"{key}/qubit_sizes"] = q
            if shape in SMALL:
                enc[f"{key}/map"] = m
            else:
                enc[f"{key}/map_sha256"] = np.array(sha(m.astype(np.int64)))
            # the permuted payload itself: int32 ramp 0..n-1 scattered by the reference
            ramp = np.arange(n, dtype=np.int32).reshape(shape)
            dense = scatter_encode(ramp)
            if shape in SMALL:
                enc[f"{key}/encoded_ramp"] = dense
            else:
                enc[f"{key}/encoded_ramp_sha256"] = np.array(sha(dense))
    np.savez_compressed(OUT / "encoding.npz", **enc)

    rng = np.random.default_rng(7)
    qz = {}
    for i, shp in enumerate([(17,), (5, 8, 3), (34, 20, 9)]):
        a = rng.standard_normal(shp) * (i + 1)
        qz[f"in{i}"] = a
        for dt in (np.uint8, np.uint16):
            q = ref_ft.scale_to_dtype(a, dt)
            qz[f"q{i}_{np.dtype(dt).name}"] = q
            qz[f"back{i}_{np.dtype(dt).name}"] = ref_ft.scale_back(q, a.min(), a.max(), dt)
    qz["bits"] = np.array([ref_ft.get_num_bits(d) for d in (np.uint8, np.uint16, np.int32, np.float32, np.float64)])
    np.savez_compressed(OUT / "quantise.npz", **qz)

    dc = {}
    for i, shp in enumerate([(6, 16), (3, 5, 12), (2, 3, 4, 25), (4, 400), (3, 680)]):
        a = rng.random(shp)
        dc[f"in{i}"] = a
        dc[f"dct{i}"] = dct(a, norm="ortho")
        dc[f"idct{i}"] = idct(a, norm="ortho")
    np.savez_compressed(OUT / "dct.npz", **dc)
    print("wrote", [p.name for p in OUT.glob("*.npz")])


if __name__ == "__main__":
    main()
