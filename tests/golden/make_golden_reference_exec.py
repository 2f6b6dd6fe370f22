"""Generate tests/golden/reference_class.npz, reference_metrics.npz and reference_benchmark.npz by EXECUTING the
reference's own ``core/ndmps.py``, ``utils/metrics.py`` and ``evaluation/benchmark.py``.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_reference_exec.py

Those two modules import quimb and scikit-image, which are not installed here and are not vendored by the reference.
This script registers stand-ins for exactly the library surface the two files touch
(``quimb.tensor.MatrixProductState.from_dense / .arrays / @ / ^ ... / [i] / .sites / iteration / bond_sizes``,
``quimb.tensor.tensor_compress_bond``, ``skimage.metrics.structural_similarity``) built on ``oracle.mps`` and
``oracle.metrics.structural_similarity``, then runs the REFERENCE SOURCE on seeded inputs.  What the fixtures pin is
therefore everything the reference's own code does AROUND the third-party cores: the scatter / gather through its
encoding map, the norm and DCT options, boundary list, norm value, the compress loop and its refreshes, element
counts and ratios, quantisation with the stored boundaries, gzip sizes, storage figures, the printed lines of
``continuous_compress``; and for the metrics the clipping, data range, window choice, slice order, axis / frame
averaging, dispatch, PSNR and ``compute_mean_std``.  What they do NOT pin is the inside of the TT-SVD, the bond
compression, the overlap and the SSIM window formula - those are the oracle's restatement on both sides
(parity unpinned, see oracle/__init__.py).
"""
import contextlib
import io
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import metrics as OM            # noqa: E402
from oracle import mps as M                 # noqa: E402

OUT = Path(__file__).resolve().parent


# ---- stand-in for the part of quimb.tensor the reference touches --------------------------------
class _Site:
    """``mps[i]`` / an item of ``for t in mps``: a view of core i that follows in-place replacement."""

    def __init__(self, mps, i):
        self.mps, self.i = mps, i

    @property
    def data(self):
        return self.mps._cores[self.i]

    @property
    def size(self):
        return self.data.size


class _Dense:
    """``mps ^ ...``: the contracted tensor with named indices k0..k{L-1}."""

    def __init__(self, data):
        self.data = data
        self.inds = tuple(f"k{i}" for i in range(data.ndim))

    def moveindex(self, name, pos, inplace=False):
        cur = self.inds.index(name)
        data = np.moveaxis(self.data, cur, pos)
        inds = list(self.inds)
        inds.insert(pos, inds.pop(cur))
        if inplace:
            self.data, self.inds = data, tuple(inds)
            return self
        out = _Dense(data)
        out.inds = tuple(inds)
        return out


class MatrixProductState:
    def __init__(self, cores):
        self._cores = list(cores)

    @classmethod
    def from_dense(cls, psi, dims):
        return cls(M.tt_svd(psi, list(dims)))

    @property
    def arrays(self):
        return tuple(self._cores)

    @property
    def sites(self):
        return tuple(range(len(self._cores)))

    def __getitem__(self, i):
        return _Site(self, i)

    def __iter__(self):
        return iter(_Site(self, i) for i in range(len(self._cores)))

    def __matmul__(self, other):
        return M.overlap(self._cores, other._cores)

    def __xor__(self, what):
        assert what is Ellipsis
        return _Dense(M.contract_dense(self._cores))

    def bond_sizes(self):
        return M.bond_sizes(self._cores)

    def show(self):
        print("MPS bonds", self.bond_sizes())


def tensor_compress_bond(t1, t2, cutoff=1e-10, cutoff_mode="rel", **kw):
    mps, i = t1.mps, t2.i
    last = i == len(mps._cores) - 1
    a, b, _ = M.compress_bond(t1.data, t2.data, first=(t1.i == 0), last=last, cutoff=cutoff, cutoff_mode=cutoff_mode)
    mps._cores[t1.i], mps._cores[i] = a, b


def install_stand_ins():
    quimb = types.ModuleType("quimb")
    qtn = types.ModuleType("quimb.tensor")
    qtn.MatrixProductState = MatrixProductState
    qtn.tensor_compress_bond = tensor_compress_bond
    quimb.tensor = qtn
    sys.modules["quimb"], sys.modules["quimb.tensor"] = quimb, qtn
    skimage = types.ModuleType("skimage")
    skm = types.ModuleType("skimage.metrics")
    skm.structural_similarity = lambda a, b, data_range=None, win_size=7: OM.structural_similarity(a, b, data_range, win_size)
    skimage.metrics = skm
    sys.modules["skimage"], sys.modules["skimage.metrics"] = skimage, skm
    nib = types.ModuleType("nibabel")          # evaluation/benchmark.py imports it at the top; the loop below never loads a file
    nib.load = lambda path: (_ for _ in ()).throw(RuntimeError("nibabel stand-in: no files are read here"))
    sys.modules["nibabel"] = nib


def smooth(shape, seed):
    rng = np.random.default_rng(seed)
    grids = np.meshgrid(*[np.linspace(-1, 1, n) for n in shape], indexing="ij")
    r2 = sum((g - 0.1 * k) ** 2 * (1 + 0.3 * k) for k, g in enumerate(grids))
    return np.exp(-2.5 * r2) + 0.3 * np.prod([np.cos(2.0 * g + k) for k, g in enumerate(grids)], axis=0) + 0.01 * rng.random(shape)


CLASS_CASES = {
    # name: (shape, generator, norm, mode, compress cutoff)
    "rand2d": ((24, 36), "random", False, "Std", 0.1),
    "rand3d_dct": ((12, 18, 10), "random", False, "DCT", 0.1),
    "rand3d_norm": ((8, 16, 6), "random", True, "Std", 0.2),
    "smooth3d": ((16, 16, 16), "smooth", False, "Std", 0.05),
    "smooth4d_dct_norm": ((8, 6, 4, 10), "smooth", True, "DCT", 0.02),
}


def class_input(name):
    shape, gen, _, _, _ = CLASS_CASES[name]
    seed = 2025 + sorted(CLASS_CASES).index(name)
    return np.random.default_rng(seed).random(shape) if gen == "random" else smooth(shape, seed)


def flat(arrs):
    return np.concatenate([np.ravel(a) for a in arrs])


def main():
    install_stand_ins()
    sys.path.insert(0, "/root/reference/src")
    from imgcompressionmps.core.ndmps import NDMPS as RefNDMPS          # the reference's class, unmodified
    from imgcompressionmps.utils import metrics as ref_metrics           # the reference's metrics, unmodified

    out = {}
    for name, (shape, _, norm, mode, cutoff) in CLASS_CASES.items():
        x = class_input(name)
        g = RefNDMPS.from_tensor(x.copy(), norm=norm, mode=mode)
        out[f"{name}/input"] = x
        out[f"{name}/options"] = np.array([str(norm), mode, repr(cutoff)])
        out[f"{name}/qubit_size"] = np.asarray(g.qubit_size)
        out[f"{name}/bonds0"] = np.asarray(g.bond_sizes())
        out[f"{name}/boundary0"] = np.asarray(g.boundary_list)
        out[f"{name}/norm0"] = np.float64(g.norm_value)
        out[f"{name}/ratio0"] = np.float64(g.compression_ratio())
        out[f"{name}/elements0"] = np.int64(g.number_elements_in_MPS())
        out[f"{name}/tensor0"] = g.to_tensor()
        out[f"{name}/storage0"] = np.float64(g.get_storage_space(np.uint16))
        g.compress(cutoff)
        out[f"{name}/bonds1"] = np.asarray(g.bond_sizes())
        out[f"{name}/boundary1"] = np.asarray(g.boundary_list)
        out[f"{name}/norm1"] = np.float64(g.norm_value)
        out[f"{name}/ratio1"] = np.float64(g.compression_ratio())
        out[f"{name}/tensor1"] = g.to_tensor()
        out[f"{name}/fidelity1"] = np.float64(ref_metrics.compute_overlap(g, RefNDMPS.from_tensor(x.copy(), norm=norm, mode=mode)))
        for dt in (np.uint16, np.uint8):
            tag = np.dtype(dt).name
            out[f"{name}/ints_{tag}"] = flat(g.compress_to_dtype(dt))
            out[f"{name}/gzip_{tag}"] = np.int64(g.get_bytesize_on_disk(dt))
            out[f"{name}/disk_ratio_{tag}"] = np.float64(g.compression_ratio_on_disk(dt))
            out[f"{name}/storage_{tag}"] = np.float64(g.get_storage_space(dt))
        # quantise in place (core/ndmps.py:201-206), then everything that is refreshed
        g.compress_to_dtype(np.uint8, replace=True)
        out[f"{name}/boundary2"] = np.asarray(g.boundary_list)
        out[f"{name}/norm2"] = np.float64(g.norm_value)
        out[f"{name}/tensor2"] = g.to_tensor()
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            g.continuous_compress(2 * cutoff)
        out[f"{name}/printed"] = np.array(buf.getvalue())
        out[f"{name}/bonds3"] = np.asarray(g.bond_sizes())
        out[f"{name}/tensor3"] = g.to_tensor()
        # fidelity of the compressed state against a fresh one (utils/metrics.py:149-160)
        fresh = RefNDMPS.from_tensor(x.copy(), norm=norm, mode=mode)
        out[f"{name}/fidelity"] = np.float64(ref_metrics.compute_overlap(g, fresh))
    np.savez_compressed(OUT / "reference_class.npz", **out)

    rng = np.random.default_rng(99)
    met = {}
    pairs = {"img": (40, 52), "small": (5, 9), "tiny": (4, 4), "vol": (12, 10, 9), "thin": (6, 20, 3), "series": (8, 9, 7, 4)}
    for name, shape in pairs.items():
        a = smooth(shape, 7 + len(shape)) + 0.2
        b = a + 0.05 * rng.standard_normal(shape)
        b[tuple(0 for _ in shape)] = -0.7                      # exercises the clip of the second argument at 0
        met[f"{name}/a"], met[f"{name}/b"] = a, b
        met[f"{name}/ssim"] = np.float64(ref_metrics.compute_ssim_by_dim(a, b))
        met[f"{name}/ssim_swapped"] = np.float64(ref_metrics.compute_ssim_by_dim(b, a))
        met[f"{name}/psnr"] = np.float64(ref_metrics.compute_psnr(a, b))
        met[f"{name}/psnr_same"] = np.float64(ref_metrics.compute_psnr(a, a))
        if len(shape) == 3:
            for ax in range(3):
                met[f"{name}/axis{ax}"] = np.asarray(ref_metrics.ssim_3d_axis(a, b, ax))
            met[f"{name}/axis_neg"] = np.asarray(ref_metrics.ssim_3d_axis(a, b, -1), dtype=np.float64)
    # compute_mean_std (utils/metrics.py:163-202): three samples, one of them with an all-prime shape
    curves = {"compressionratio_list_disk": [[0.9, 0.5, 0.2, 0.05], [0.8, 0.4, 0.25, 0.04], [0.7, 0.3, 0.1, 0.02]],
              "ssim_list": [[0.99, 0.9, 0.7, 0.4], [0.98, 0.85, 0.75, 0.3], [0.5, 0.4, 0.3, 0.2]],
              "shapes": [[64, 64, 32], [40, 52, 9], [7, 11, 13]]}
    mean, std, grid = ref_metrics.compute_mean_std(curves, 6)
    met["mean_std/x"] = np.asarray(curves["compressionratio_list_disk"])
    met["mean_std/y"] = np.asarray(curves["ssim_list"])
    met["mean_std/shapes"] = np.asarray(curves["shapes"])
    met["mean_std/mean"], met["mean_std/std"], met["mean_std/grid"] = np.asarray(mean), np.asarray(std), np.asarray(grid)
    primes = {"compressionratio_list_disk": curves["compressionratio_list_disk"][:1], "ssim_list": curves["ssim_list"][:1], "shapes": [[7, 11, 13]]}
    m2, s2, g2 = ref_metrics.compute_mean_std(primes, 4)
    met["mean_std/all_prime_is_nan"] = np.array([np.isnan(m2) and np.isnan(s2)])
    met["mean_std/all_prime_grid"] = np.asarray(g2)
    np.savez_compressed(OUT / "reference_metrics.npz", **met)

    # the reference's cutoff sweep (evaluation/benchmark.py:149-194) on three small volumes: every metric before compression
    # and after each cutoff, including the in-place uint16 quantisation its gzip_ratio metric performs at every level
    from imgcompressionmps.evaluation import benchmark as ref_bm
    vols = [smooth((18, 12, 8), 300 + i) + 0.1 for i in range(3)]
    cutoffs = np.array([0.02, 0.1])
    with contextlib.redirect_stdout(io.StringIO()) as printed:
        res = ref_bm.run_benchmark(ref_bm.conv_to_mps(vols, mode="Std"), vols, cutoffs)
    bench = {"volumes": np.stack(vols), "cutoffs": cutoffs, "printed": np.array(printed.getvalue())}
    for key, value in res.items():
        if key == "bond_dims":
            bench["bond_dims"] = np.array([[b for b in level] for level in value])      # [level][tensor][bond]
        else:
            bench[key] = np.asarray(value)
    np.savez_compressed(OUT / "reference_benchmark.npz", **bench)
    print("wrote", OUT / "reference_class.npz", OUT / "reference_metrics.npz", OUT / "reference_benchmark.npz")


if __name__ == "__main__":
    main()
