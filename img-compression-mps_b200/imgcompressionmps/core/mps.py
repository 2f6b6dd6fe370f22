"""Device-resident stand-in for the ``quimb.tensor.MatrixProductState`` object that
leaks through ``NDMPS.mps`` in the reference.

Only the surface the reference and its tests touch is provided:
``mps.arrays`` (ordered, mutable, write-through), ``mps @ other``, ``mps.sites``,
``mps[i]``, iteration over site tensors with ``.size``, ``mps ^ ...`` (an object with
``.inds``, ``.moveindex`` and ``.data``), ``mps.bond_sizes()``, ``mps.show()`` and
``copy.deepcopy``.  Cores live on the GPU as torch tensors laid out
(left bond, physical, right bond); every contraction goes through the C ABI.
"""
from __future__ import annotations

import numpy as np

from .. import _ops


class CoreArray:
    """numpy-flavoured, write-through handle on one core (or a slice of it)."""

    __array_priority__ = 1000

    def __init__(self, tensor):
        self._t = tensor

    # -- metadata ------------------------------------------------------------------
    @property
    def shape(self):
        return tuple(int(s) for s in self._t.shape)

    @property
    def ndim(self):
        return self._t.ndim

    @property
    def size(self):
        return int(self._t.numel())

    @property
    def dtype(self):
        return np.dtype(str(self._t.dtype).replace("torch.", ""))

    @property
    def device_tensor(self):
        return self._t

    # -- numpy interop ---------------------------------------------------------------
    def __array__(self, dtype=None, copy=None):
        arr = self._t.detach().cpu().numpy()
        return arr.astype(dtype) if dtype is not None else arr

    def tobytes(self):
        return np.asarray(self).tobytes()

    def __len__(self):
        return int(self._t.shape[0])

    def __repr__(self):
        return f"CoreArray(shape={self.shape}, dtype={self.dtype}, device='{self._t.device}')"

    # -- write-through indexing --------------------------------------------------------
    def _coerce(self, value):
        import torch
        if isinstance(value, CoreArray):
            return value._t
        if isinstance(value, torch.Tensor):
            return value.to(device=self._t.device, dtype=self._t.dtype)
        if np.isscalar(value):
            return value
        return torch.as_tensor(np.asarray(value), device=self._t.device).to(self._t.dtype)

    def __getitem__(self, idx):
        out = self._t[idx]
        return CoreArray(out) if out.ndim > 0 else out.item()

    def __setitem__(self, idx, value):
        value = self._coerce(value)
        target = self._t[idx]
        if hasattr(value, "data_ptr") and value.data_ptr() == target.data_ptr() and value.shape == target.shape:
            return  # `a[:] *= 10` assigns the view back onto itself
        self._t[idx] = value

    def __imul__(self, other):
        self._t.mul_(self._coerce(other))
        return self

    def __itruediv__(self, other):
        self._t.div_(self._coerce(other))
        return self

    def __iadd__(self, other):
        self._t.add_(self._coerce(other))
        return self

    def __isub__(self, other):
        self._t.sub_(self._coerce(other))
        return self


class SiteTensor:
    """What ``mps[i]`` / iteration yields: one core with quimb-like attributes."""

    def __init__(self, mps, i):
        self._mps, self._i = mps, i

    @property
    def data(self):
        return CoreArray(self._mps.cores[self._i])

    @property
    def shape(self):
        return tuple(int(s) for s in self._mps.cores[self._i].shape)

    @property
    def size(self):
        return int(self._mps.cores[self._i].numel())

    @property
    def inds(self):
        i, L = self._i, self._mps.L
        phys = f"k{i}"
        if L == 1:
            return (phys,)
        if i == 0:
            return (phys, "b0")
        if i == L - 1:
            return (f"b{i - 1}", phys)
        return (f"b{i - 1}", phys, f"b{i}")


class DenseResult:
    """Result of ``mps ^ ...``: the fully contracted tensor with named indices."""

    def __init__(self, dense_device, inds):
        self._dense = dense_device
        self.inds = tuple(inds)

    def moveindex(self, name, position, inplace=False):
        order = list(self.inds)
        src = order.index(name)
        order.insert(position, order.pop(src))
        perm = [self.inds.index(n) for n in order]
        moved = self._dense.permute(perm)
        if inplace:
            self._dense, self.inds = moved, tuple(order)
            return self
        return DenseResult(moved, order)

    @property
    def data(self):
        return self._dense.contiguous().cpu().numpy()

    @property
    def device_tensor(self):
        return self._dense

    @property
    def shape(self):
        return tuple(int(s) for s in self._dense.shape)


class DeviceMPS:
    """Matrix product state held as a list of device tensors."""

    def __init__(self, cores, site_dims):
        self.cores = list(cores)
        self.site_dims = [int(d) for d in site_dims]

    # -- structure -------------------------------------------------------------------
    @property
    def L(self):
        return len(self.cores)

    nsites = L

    @property
    def sites(self):
        return tuple(range(self.L))

    @property
    def arrays(self):
        return tuple(CoreArray(c) for c in self.cores)

    def __len__(self):
        return self.L

    def __getitem__(self, i):
        if i < 0:
            i += self.L
        if not 0 <= i < self.L:
            raise IndexError(i)
        return SiteTensor(self, i)

    def __iter__(self):
        return (SiteTensor(self, i) for i in range(self.L))

    def bond_sizes(self):
        return _ops._ranks_of(self.cores)

    @property
    def dtype(self):
        return self.cores[0].dtype

    def copy(self):
        return DeviceMPS([c.clone() for c in self.cores], self.site_dims)

    def __deepcopy__(self, memo):
        return self.copy()

    # -- contractions (C ABI) ------------------------------------------------------------
    def __matmul__(self, other):
        if not isinstance(other, DeviceMPS):
            return NotImplemented
        return _ops.overlap(self.cores, other.cores)

    def __xor__(self, other):
        if other is not Ellipsis:
            return NotImplemented
        return DenseResult(_ops.contract_dense(self.cores), [f"k{i}" for i in range(self.L)])

    def to_dense_device(self):
        return _ops.contract_dense(self.cores)

    # -- display -----------------------------------------------------------------------
    def show(self):
        bonds = self.bond_sizes()
        top = " ".join(f"{b}" for b in bonds)
        line = "─".join("●" for _ in range(self.L))
        legs = " ".join("│" for _ in range(self.L))
        print(f" {top}\n{line}\n{legs}")

    def __repr__(self):
        return f"DeviceMPS(L={self.L}, site_dims={self.site_dims}, bonds={self.bond_sizes()}, dtype={self.dtype})"
