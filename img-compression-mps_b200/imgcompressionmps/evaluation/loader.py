"""Data-set ingestion for the device path.

``load_tensors`` has the reference's signature and return value (``evaluation/benchmark.py:16-55``):
``.npz`` archives carry the tensor under the key ``"sequence"``; ``.gz`` files are NIfTI volumes.
The reference reads NIfTI with nibabel (``img.get_fdata()`` / ``img.header.get_data_dtype()``); nibabel
is used here when it is installed, otherwise the small NIfTI-1 reader below returns the same two things
(scaled float64 data in file axis order, stored dtype).

``prefetch_to_device`` is the piece the reference has no counterpart for: the next tensors are
converted and staged in pinned host memory by a helper thread and copied to the device on a side
stream while the current one is being encoded, so the sweep never waits for PCIe.
"""
from __future__ import annotations

import gzip
import queue
import struct
import threading
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from ..utils.filetools import get_num_bits

_NIFTI_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16,
                 768: np.uint32, 1024: np.int64, 1280: np.uint64}


def read_nifti(path) -> Tuple[np.ndarray, np.dtype]:
    """(float64 data with the header's intensity scaling applied, stored dtype) of a single-file NIfTI-1
    image, plain or gzip-compressed - what ``nib.load(path).get_fdata()`` and ``.header.get_data_dtype()`` give."""
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        raw = f.read()
    if len(raw) < 352:
        raise ValueError(f"{path}: too short for a NIfTI-1 header")
    for endian in ("<", ">"):
        if struct.unpack(endian + "i", raw[0:4])[0] == 348:
            break
    else:
        raise ValueError(f"{path}: not a NIfTI-1 file (sizeof_hdr != 348)")
    if raw[344:348] != b"n+1\0":
        raise ValueError(f"{path}: only single-file NIfTI-1 (magic 'n+1') is supported, found {raw[344:348]!r}")
    dim = struct.unpack(endian + "8h", raw[40:56])
    datatype = struct.unpack(endian + "h", raw[70:72])[0]
    vox_offset = struct.unpack(endian + "f", raw[108:112])[0]
    slope, inter = struct.unpack(endian + "2f", raw[112:120])
    if datatype not in _NIFTI_DTYPES:
        raise ValueError(f"{path}: unsupported NIfTI datatype code {datatype}")
    stored = np.dtype(_NIFTI_DTYPES[datatype]).newbyteorder(endian)
    shape = tuple(int(d) for d in dim[1:1 + dim[0]])
    count = int(np.prod(shape))
    start = int(vox_offset)
    data = np.frombuffer(raw, dtype=stored, count=count, offset=start).reshape(shape, order="F").astype(np.float64)
    if slope not in (0.0, 1.0) or inter != 0.0:
        if slope != 0.0 and np.isfinite(slope) and np.isfinite(inter):
            data = data * float(slope) + float(inter)
    return data, np.dtype(_NIFTI_DTYPES[datatype])


class _NiftiHeader:
    def __init__(self, stored):
        self._stored = stored

    def get_data_dtype(self):
        return self._stored


class _NiftiImage:
    """The two things the reference asks of a nibabel image (``evaluation/benchmark.py:41-43``)."""

    def __init__(self, path):
        self._data, stored = read_nifti(path)
        self.header = _NiftiHeader(stored)

    def get_fdata(self):
        return self._data


class _NiftiModule:
    """Stands where the reference's module-level ``nib`` stands when nibabel is not installed: ``nib.load(path)``."""

    @staticmethod
    def load(path):
        return _NiftiImage(path)


def nifti_module():
    """nibabel when it is installed, else the NIfTI-1 reader above behind the same ``load`` call."""
    try:
        import nibabel
        return nibabel
    except ImportError:
        return _NiftiModule()


def load_tensors(files, ending, shape=None, *, nib=None, num_bits=None):
    """Load every file of ``files`` (``evaluation/benchmark.py:16-55``): returns ``(data_list, bitsize_list)``.
    ``shape = (B, H, W)`` crops the first three axes; ``ending`` must end in ``.gz`` or ``.npz``.  ``nib`` / ``num_bits``
    let ``evaluation.benchmark`` hand in its own module-level ``nib`` and ``get_num_bits`` (the names the reference's
    module has, and the ones its tests replace)."""
    if not (ending.endswith(".gz") or ending.endswith(".npz")):
        raise ValueError(f"Unsupported file extension: {ending}")
    nib = nib if nib is not None else nifti_module()
    num_bits = num_bits if num_bits is not None else get_num_bits
    crop = tuple(slice(None, n) for n in shape) if shape else None
    data_list, bitsize_list = [], []
    for index, path in enumerate(files):
        print(f"Loading file {index + 1}/{len(files)}")
        if ending.endswith(".gz"):
            img = nib.load(path)
            data, stored = img.get_fdata(), img.header.get_data_dtype()
        else:
            with np.load(path) as archive:
                data = archive["sequence"]
                stored = data.dtype
        if crop:
            data = data[crop]
        data_list.append(data)
        bitsize_list.append(num_bits(stored))
    return data_list, bitsize_list


def prefetch_to_device(arrays: Iterable, depth: int = 2, dtype=None, device: Optional[int] = None) -> Iterator:
    """Yield the arrays of ``arrays`` as CUDA tensors, ``depth`` uploads ahead of the consumer.

    A helper thread makes each array contiguous in the working dtype (float32 stays float32, everything
    else becomes float64, as ``NDMPS.from_tensor`` promotes), stages it in a pinned buffer (``depth + 1``
    buffers per distinct size, recycled) and enqueues the host-to-device copy on its own stream; the
    consumer's current stream waits for the copy's event only.  ``arrays`` may be a generator that reads
    files lazily, so loading, uploading and encoding overlap."""
    import torch
    if not torch.cuda.is_available():
        from .. import _native
        raise _native.NativeError("prefetch_to_device needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
    dev = torch.cuda.current_device() if device is None else int(device)
    copy_stream = torch.cuda.Stream(device=dev)
    ready: "queue.Queue" = queue.Queue(maxsize=max(1, depth))
    pools: dict = {}
    free = threading.Semaphore(depth + 1)

    def producer():
        try:
            torch.cuda.set_device(dev)
            for arr in arrays:
                arr = np.asarray(arr)
                work = np.float32 if (dtype is None and arr.dtype == np.float32) or dtype == np.float32 else np.float64
                free.acquire()
                key = (arr.shape, work)
                stash = pools.setdefault(key, [])
                host = stash.pop() if stash else torch.empty(arr.shape, dtype=torch.float32 if work == np.float32 else torch.float64).pin_memory()
                np.copyto(host.numpy(), arr, casting="unsafe")
                with torch.cuda.stream(copy_stream):
                    on_device = host.to(f"cuda:{dev}", non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(copy_stream)
                ready.put((on_device, done, host, key))
            ready.put(None)
        except BaseException as exc:      # noqa: BLE001 - re-raised in the consumer
            ready.put(exc)

    thread = threading.Thread(target=producer, name="ndmps-prefetch", daemon=True)
    thread.start()
    while True:
        item = ready.get()
        if item is None:
            break
        if isinstance(item, BaseException):
            raise item
        on_device, done, host, key = item
        torch.cuda.current_stream(dev).wait_event(done)
        on_device.record_stream(torch.cuda.current_stream(dev))
        done.synchronize()                 # the pinned buffer may be refilled once the copy has left it
        pools[key].append(host)
        free.release()
        yield on_device
    thread.join()


def conv_to_mps_streamed(data_list: Sequence, mode: str = "DCT", *, max_bond=None, cutoff: float = 1e-10, depth: int = 2) -> List:
    """``[NDMPS.from_tensor(x, norm=False, mode=mode) for x in data_list]`` with the uploads prefetched."""
    from ..core.ndmps import NDMPS
    out = []
    for index, volume in enumerate(prefetch_to_device(data_list, depth=depth)):
        print(f"Converting file {index + 1}/{len(data_list)}")
        out.append(NDMPS.from_tensor(volume, norm=False, mode=mode, max_bond=max_bond, cutoff=cutoff))
    return out
