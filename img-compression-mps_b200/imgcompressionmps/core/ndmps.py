"""``NDMPS`` - the reference's class (``core/ndmps.py:11-277``) on the B200 device path.

Same constructor, methods, attribute names, argument meaning and error behaviour
as the reference; positional signatures are unchanged and the options the BASELINE
configs need are keyword-only with reference defaults:

* ``from_tensor(tensor, norm=False, mode="Std", *, max_bond=None, cutoff=1e-10,
  dtype=None, device=None)``  - quimb's ``from_dense(**split_opts)`` at
  ``core/ndmps.py:74`` already accepts ``max_bond`` / ``cutoff``; rule
  ``n = min(n_by_cutoff, max_bond)``.
* ``compress(cutoff, *, max_bond=None)`` - likewise for ``tensor_compress_bond``.

``dtype=None`` follows the input: float32 input stays float32 (the B200 fast path the
BASELINE configs name), everything else is promoted to float64 exactly as the
reference does at ``core/ndmps.py:56``.  Inputs may be numpy arrays (copied to the
device) or CUDA torch tensors (used in place, never aliased by the MPS).

Every numeric step is a call into ``libndmps_sm100.so``; there is no CPU path.
"""
from __future__ import annotations

import gzip
import io

import numpy as np

from .. import _ops
from ..utils.core import gen_encoding_map
from ..utils.filetools import get_num_bits
from .mps import DeviceMPS


def _torch():
    import torch
    return torch


def _to_device(tensor, dtype=None, device=None):
    """numpy / torch input -> contiguous CUDA tensor of the working dtype (always a copy
    or a fresh upload: the MPS never aliases the caller's data, as ``astype`` guarantees
    at ``core/ndmps.py:56``)."""
    torch = _torch()
    if isinstance(tensor, torch.Tensor):
        work = dtype or (torch.float32 if tensor.dtype == torch.float32 else torch.float64)
        dev = device or (tensor.device if tensor.is_cuda else torch.device("cuda"))
        return tensor.to(device=dev, dtype=work).contiguous()
    arr = np.asarray(tensor)
    work = dtype or (torch.float32 if arr.dtype == np.float32 else torch.float64)
    np_work = np.float32 if work == torch.float32 else np.float64
    host = np.ascontiguousarray(arr, dtype=np_work)
    return torch.from_numpy(host).to(device or "cuda", non_blocking=False)


class NDMPS:
    """Class for storing and compressing N-dimensional tensors using MPS (device-resident)."""

    def __init__(self, mps=None, qubit_size=None, encoding_map=None, boundary_list=None, norm: bool = True,
                 norm_value=None, mode: str = "Std", dim: int = None):
        self.qubit_size = qubit_size
        self._encoding_map = encoding_map
        self._shape = None
        self.mps = mps
        self.dim = dim
        self.norm = norm
        self.norm_value = norm_value
        self.mode = mode
        self.boundary_list = np.array(boundary_list)          # core/ndmps.py:34 (wraps None too)

    # The reference stores the (*shape, L) int64 map eagerly (8*L bytes per voxel, 9.7 GB at
    # 512^3).  Nothing on the device path needs it, so it is built on first access only.
    @property
    def encoding_map(self):
        if self._encoding_map is None and self._shape is not None:
            _, enc = gen_encoding_map(self._shape)
            self._encoding_map = np.moveaxis(enc, 0, -1)
        return self._encoding_map

    @encoding_map.setter
    def encoding_map(self, value):
        self._encoding_map = value

    @property
    def shape(self):
        return self._shape

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_tensor(cls, tensor, norm: bool = False, mode: str = "Std", *, max_bond=None, cutoff: float = 1e-10,
                    dtype=None, device=None) -> "NDMPS":
        """N-D array -> MPS: (normalise) -> (DCT) -> site-order permutation -> TT-SVD
        (``core/ndmps.py:36-78``)."""
        vol = _to_device(tensor, dtype, device)
        shape = tuple(int(s) for s in vol.shape)
        plan = _ops.plan_for(shape)                      # raises ValueError like gen_encoding_map
        scale = 1.0
        if norm:
            scale = 1.0 / float(np.sqrt(_ops.sumsq(vol)))                 # core/ndmps.py:60-61
        if mode == "DCT":
            vol = _ops.dct_last_axis(vol, inverse=False)                  # core/ndmps.py:62-63
        dense = _ops.encode(vol, scale)                                   # core/ndmps.py:66-71
        del vol
        cores, _, svals = _ops.ttsvd(dense, plan.site_dims, cutoff=cutoff, cutoff_mode="rsum2",
                                     max_bond=max_bond)                  # core/ndmps.py:74
        del dense
        obj = cls(DeviceMPS(cores, plan.site_dims), np.array(plan.site_dims), None, None, norm, None, mode, len(shape))
        obj._shape = shape
        obj.singular_values = svals
        obj.update_boundary_list()                                        # core/ndmps.py:75
        obj.update_norm()                                                 # core/ndmps.py:76
        return obj

    # ------------------------------------------------------------------ bookkeeping
    def update_boundary_list(self):
        """Recompute min/max boundaries for each MPS tensor (``core/ndmps.py:80-82``)."""
        self.boundary_list = _ops.minmax(self.mps.cores)

    def update_norm(self):
        """Update stored norm of the current MPS (``core/ndmps.py:84-86``)."""
        self.norm_value = np.sqrt(self.mps @ self.mps)

    def compression_ratio(self):
        """MPS elements / original tensor elements (``core/ndmps.py:88-92``)."""
        return self.number_elements_in_MPS() / np.prod(self.qubit_size)

    def number_elements_in_MPS(self) -> int:
        return sum(t.size for t in self.mps)

    def bond_sizes(self):
        return self.mps.bond_sizes()

    def show(self):
        self.mps.show()

    # ------------------------------------------------------------------ truncation
    def compress(self, cutoff: float, *, max_bond=None):
        """Pairwise bond truncation, bonds left to right, relative cutoff, no
        canonicalisation in between (``core/ndmps.py:94-108``)."""
        cores = self.mps.cores
        L = len(cores)
        svals = []
        for i in range(1, L):
            t1, t2 = cores[i - 1], cores[i]
            a = t1.reshape(-1, t1.shape[-1])
            b = t2.reshape(t2.shape[0], -1)
            a_new, b_new, s = _ops.compress_bond(a, b, cutoff, "rel", max_bond)
            n = a_new.shape[1]
            cores[i - 1] = a_new.view(*t1.shape[:-1], n)
            cores[i] = b_new.view(n, *t2.shape[1:])
            svals.append(s)
        self.singular_values = svals
        self.update_boundary_list()
        self.update_norm()

    def continuous_compress(self, cutoff: float, print_ratio: bool = True):
        """20 compress steps with cutoffs linspace(0, 1, 20) * cutoff (``core/ndmps.py:110-125``)."""
        for c in np.linspace(0, 1, 20) * cutoff:
            self.compress(c)
            if print_ratio:
                print(f"Compression ratio at {c}: {self.compression_ratio()}")

    # ------------------------------------------------------------------ reconstruction
    def to_tensor_device(self):
        """Reconstruction as a CUDA tensor of the original shape (no host copy)."""
        dense = _ops.contract_dense(self.mps.cores)                       # core/ndmps.py:140-142
        rec = _ops.decode(dense, self._shape)                             # core/ndmps.py:144-148
        if self.mode == "Std":
            return rec
        if self.mode == "DCT":
            return _ops.dct_last_axis(rec, inverse=True)                  # core/ndmps.py:153
        return None

    def to_tensor(self) -> np.ndarray:
        """MPS back to a numpy array (with inverse DCT in "DCT" mode); any other mode
        returns None as in the reference (``core/ndmps.py:131-153``)."""
        rec = self.to_tensor_device()
        return None if rec is None else rec.cpu().numpy()

    # ------------------------------------------------------------------ core access
    def replace_tensordata(self, tensorlist):
        """Shape-asserting in-place overwrite of every core (``core/ndmps.py:163-175``)."""
        arrays = self.mps.arrays
        for i in range(len(arrays)):
            assert arrays[i].shape == tuple(tensorlist[i].shape)
            arrays[i][:] = tensorlist[i]
        self.update_boundary_list()
        self.update_norm()

    def return_tensors_data(self):
        return [t for t in self.mps.arrays]

    # ------------------------------------------------------------------ quantisation / size
    def _quantised_device(self, dtype):
        """[uint core on device], following scale_to_dtype (``filetools.py:20-26``) with each
        core's own current min/max."""
        code = _ops.int_code(dtype)              # any np.iinfo dtype, like the reference; ValueError otherwise
        bounds = _ops.minmax(self.mps.cores)
        return [_ops.quantize(c, lo, hi, code) for c, (lo, hi) in zip(self.mps.cores, bounds)], code

    def compress_to_dtype(self, dtype=np.uint16, replace: bool = False):
        """Integer-truncate each core (``core/ndmps.py:182-207``).  De-quantisation uses the
        STORED ``boundary_list`` like the reference (stale if cores were edited by hand)."""
        q_dev, bits = self._quantised_device(dtype)
        if replace:
            scaled_back = [_ops.dequantize(q, float(b[0]), float(b[1]), bits, c.dtype)
                           for q, b, c in zip(q_dev, self.boundary_list, self.mps.cores)]
            self.replace_tensordata(scaled_back)
        return [_ops.to_host_int(q, dtype) for q in q_dev]

    def get_bytesize_on_disk(self, dtype=np.uint16, replace: bool = False) -> int:
        """Sum of gzip sizes of the quantised cores (``core/ndmps.py:209-234``); the deflate
        runs on the host as in the reference."""
        total_bytes = 0
        for arr in self.compress_to_dtype(dtype, replace):
            buf = io.BytesIO()
            with gzip.GzipFile(fileobj=buf, mode="wb") as gz:
                gz.write(arr.tobytes())
            total_bytes += len(buf.getvalue())
        return total_bytes

    def compression_ratio_on_disk(self, dtype=np.uint16, replace: bool = False) -> float:
        original_size = np.prod(self.qubit_size) * get_num_bits(dtype) / 8.0
        return self.get_bytesize_on_disk(dtype, replace) / original_size

    def get_storage_space(self, dtype=np.uint16, verbose: bool = False) -> float:
        size_bytes = self.number_elements_in_MPS() * get_num_bits(dtype) / 8
        if verbose:
            print(f"The storage space is approximately: {size_bytes / 1024:.2f} KB")
        return size_bytes
