"""Oracle: the reference's ``NDMPS`` class in float64 numpy (CPU baseline "port").

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  Mirrors
``/root/reference/src/imgcompressionmps/core/ndmps.py:11-277`` method by method
(cited below) on top of ``oracle.encoding`` / ``oracle.mps`` /
``oracle.quantise``.  The keyword-only ``max_bond`` / ``cutoff`` options are
the extensions SURVEY.md section 8(b) defines through quimb's own
``n = min(n_by_cutoff, max_bond)`` rule; with their defaults the behaviour is
the reference's.  **Parity unpinned** for everything that goes through quimb.
"""
from __future__ import annotations

import gzip
import io

import numpy as np
from scipy.fft import dct, idct

from . import encoding as enc
from . import mps as M
from . import quantise as Q


class OracleNDMPS:
    def __init__(self, cores, qubit_size, shape, norm, norm_value, mode, boundary_list):
        self.cores = cores
        self.qubit_size = qubit_size
        self.shape = tuple(shape)
        self.dim = len(shape)
        self.norm = norm
        self.norm_value = norm_value
        self.mode = mode
        self.boundary_list = np.array(boundary_list)

    # core/ndmps.py:36-78
    @classmethod
    def from_tensor(cls, tensor, norm=False, mode="Std", *, max_bond=None, cutoff=1e-10):
        tensor = np.asarray(tensor).astype(np.float64)
        shape = tuple(int(s) for s in tensor.shape)
        if norm:
            tensor = tensor / np.linalg.norm(tensor)
        if mode == "DCT":
            tensor = dct(tensor, type=2, axis=-1, norm="ortho")
        dense = enc.encode(tensor)
        dims = list(dense.shape)
        cores, svals = M.tt_svd(dense, dims, cutoff=cutoff, max_bond=max_bond, return_svals=True)
        obj = cls(cores, np.array(dims), shape, norm, None, mode, [[c.min(), c.max()] for c in cores])
        obj.singular_values = svals
        obj.update_norm()
        return obj

    # core/ndmps.py:80-86
    def update_boundary_list(self):
        self.boundary_list = np.array([[c.min(), c.max()] for c in self.cores])

    def update_norm(self):
        self.norm_value = float(np.sqrt(M.overlap(self.cores, self.cores)))

    # core/ndmps.py:88-92, 127-129, 159-161
    def number_elements_in_MPS(self):
        return M.num_elements(self.cores)

    def compression_ratio(self):
        return self.number_elements_in_MPS() / np.prod(self.qubit_size)

    def bond_sizes(self):
        return M.bond_sizes(self.cores)

    # core/ndmps.py:94-125
    def compress(self, cutoff, *, max_bond=None):
        self.cores, self.last_svals = M.compress_all(self.cores, cutoff, max_bond=max_bond)
        self.update_boundary_list()
        self.update_norm()

    def continuous_compress(self, cutoff, print_ratio=True):
        for c in np.linspace(0, 1, 20) * cutoff:
            self.compress(c)
            if print_ratio:
                print(f"Compression ratio at {c}: {self.compression_ratio()}")

    # core/ndmps.py:131-153
    def to_tensor(self):
        rec = enc.decode(M.contract_dense(self.cores), self.shape)
        if self.mode == "Std":
            return rec
        if self.mode == "DCT":
            return idct(rec, type=2, axis=-1, norm="ortho")
        return None

    # core/ndmps.py:163-180
    def replace_tensordata(self, tensorlist):
        for i in range(len(self.cores)):
            assert self.cores[i].shape == tensorlist[i].shape
            self.cores[i][...] = tensorlist[i]
        self.update_boundary_list()
        self.update_norm()

    def return_tensors_data(self):
        return list(self.cores)

    # core/ndmps.py:182-277
    def compress_to_dtype(self, dtype=np.uint16, replace=False):
        ints = [Q.scale_to_dtype(c, dtype) for c in self.cores]
        if replace:
            self.replace_tensordata([Q.scale_back(t, b[0], b[1], dtype) for t, b in zip(ints, self.boundary_list)])
        return ints

    def get_bytesize_on_disk(self, dtype=np.uint16, replace=False):
        total = 0
        for arr in self.compress_to_dtype(dtype, replace):
            buf = io.BytesIO()
            with gzip.GzipFile(fileobj=buf, mode="wb") as gz:
                gz.write(arr.tobytes())
            total += len(buf.getvalue())
        return total

    def compression_ratio_on_disk(self, dtype=np.uint16, replace=False):
        original = np.prod(self.qubit_size) * Q.get_num_bits(dtype) / 8.0
        return self.get_bytesize_on_disk(dtype, replace) / original

    def get_storage_space(self, dtype=np.uint16, verbose=False):
        size = self.number_elements_in_MPS() * Q.get_num_bits(dtype) / 8
        if verbose:
            print(f"The storage space is approximately: {size / 1024:.2f} KB")
        return size
