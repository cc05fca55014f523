"""Raw dgCMatrix files for SHARP_unlimited3 and the double-buffered reader behind its streamed path.

The reference reads every part with ``readRDS(allfiles[i])`` (R/SHARP_unlimited3.R:105) and frees it after use
(:124-125); RDS decoding stays in R.  For streaming straight into the GPU pipeline the parts are kept in the container
``SHCSC001`` (include/sharp_b200.h documents the layout: 64-byte header, then the slots ``p``, ``i``, ``x``), which the
native reader (``sharp_csc_file_read``: pread from several threads into PINNED buffers) loads while the previous batch
of parts is being uploaded and clustered.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import threading

import numpy as np

from . import _lib

MAGIC = b"SHCSC001"
SUFFIX = ".csc"


def _pad64(b: int) -> int:
    return (b + 63) & ~63


def write_csc(path, m: int, n: int, colptr, rowidx, val) -> None:
    """write one part (dgCMatrix slots p / i / x of an m x n matrix) as an SHCSC001 file"""
    colptr = np.ascontiguousarray(colptr, dtype=np.int64)
    rowidx = np.ascontiguousarray(rowidx, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float64)
    nnz = int(colptr[-1])
    if len(colptr) != n + 1 or len(rowidx) != nnz or len(val) != nnz:
        raise ValueError("write_csc: slot lengths do not match the dimensions")
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<iiqq", int(m), 0, int(n), nnz) + b"\0" * 32)
        for a in (colptr, rowidx):
            b = a.tobytes()
            f.write(b + b"\0" * (_pad64(len(b)) - len(b)))
        f.write(val.tobytes())


def file_info(path) -> tuple[int, int, int]:
    """(m, n, nnz) from the header (no data is read)"""
    m, n, nnz = C.c_int(), C.c_int64(), C.c_int64()
    _lib._check(_lib.load().sharp_csc_file_info(os.fsencode(path), C.byref(m), C.byref(n), C.byref(nnz)))
    return m.value, n.value, nnz.value


class PinnedBuffer:
    """``nbytes`` of page-locked host memory (sharp_host_alloc) viewed through numpy"""

    def __init__(self, nbytes: int):
        self._p = C.c_void_p()
        self.nbytes = max(int(nbytes), 64)
        _lib._check(_lib.load().sharp_host_alloc(C.byref(self._p), C.c_size_t(self.nbytes)))
        self._raw = (C.c_ubyte * self.nbytes).from_address(self._p.value)

    def view(self, dtype, count, offset=0) -> np.ndarray:
        return np.frombuffer(self._raw, dtype=dtype, count=int(count), offset=int(offset))

    def close(self):
        if self._p:
            self._raw = None
            _lib.load().sharp_host_free(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PartSlot:
    """pinned room for one part: colptr | rowidx | val"""

    def __init__(self, n_max: int, nnz_max: int, pinned=True):
        self.n_max, self.nnz_max = int(n_max), int(nnz_max)
        self.o_ri = _pad64((self.n_max + 1) * 8)
        self.o_v = self.o_ri + _pad64(self.nnz_max * 4)
        total = self.o_v + self.nnz_max * 8
        self.buf = PinnedBuffer(total) if pinned else None
        self._np = None if pinned else np.zeros(total, dtype=np.uint8)

    def _view(self, dtype, count, offset):
        if self.buf is not None:
            return self.buf.view(dtype, count, offset)
        return np.frombuffer(self._np, dtype=dtype, count=int(count), offset=int(offset))

    def load(self, path, threads=4):
        m, n, nnz = file_info(path)
        if n > self.n_max or nnz > self.nnz_max:
            raise ValueError(f"{path}: {n} cells / {nnz} non-zeros do not fit the slot ({self.n_max} / {self.nnz_max})")
        cp, ri, v = self._view(np.int64, n + 1, 0), self._view(np.int32, nnz, self.o_ri), self._view(np.float64, nnz, self.o_v)
        _lib._check(_lib.load().sharp_csc_file_read(os.fsencode(path), cp.ctypes.data_as(C.POINTER(C.c_int64)),
                                                    ri.ctypes.data_as(C.POINTER(C.c_int32)),
                                                    v.ctypes.data_as(C.POINTER(C.c_double)), int(threads)))
        return m, n, (cp, ri, v)

    def close(self):
        if self.buf is not None:
            self.buf.close()


# Page-locking memory costs about a second per 2 GB, far more than reading a part: the slots of a finished run are kept
# for the next one (up to _POOL_BYTES) instead of being released, like the library's device workspaces.
_POOL_BYTES = 24 << 30
_pool: list = []
_pool_lock = threading.Lock()


def _acquire_slot(n_max, nnz_max, pinned) -> PartSlot:
    with _pool_lock:
        for k, s in enumerate(_pool):
            if (s.buf is not None) == bool(pinned) and s.n_max >= n_max and s.nnz_max >= nnz_max and s.nnz_max <= 2 * nnz_max + 4096:
                return _pool.pop(k)
    # 6 % slack, so that parts of slightly different sizes reuse the same slots
    return PartSlot(n_max + n_max // 16 + 64, nnz_max + nnz_max // 16 + 1024, pinned)


def _release_slot(slot: PartSlot) -> None:
    with _pool_lock:
        held = sum(s.o_v + s.nnz_max * 8 for s in _pool)
        if held + slot.o_v + slot.nnz_max * 8 <= _POOL_BYTES:
            _pool.append(slot)
            return
    slot.close()


def release_pool() -> None:
    """give the pooled page-locked slots back to the system"""
    with _pool_lock:
        slots = list(_pool)
        _pool.clear()
    for s in slots:
        s.close()


class BatchReader:
    """Reads the files of batch b + 1 on a host thread (the native reader releases the GIL) while batch b is in the GPU
    pipeline: two sets of ``batch`` pinned slots."""

    def __init__(self, paths, infos, batch: int, threads=4, pinned=True):
        self.paths, self.infos, self.batch, self.threads = list(paths), list(infos), max(1, int(batch)), threads
        n_max = max(i[1] for i in infos)
        nnz_max = max(i[2] for i in infos)
        self.sets = [[_acquire_slot(n_max, nnz_max, pinned) for _ in range(min(self.batch, len(paths)))] for _ in range(2)]
        self.nb = (len(paths) + self.batch - 1) // self.batch
        self._thread = None
        self._result = None
        self._error = None

    def _read(self, b):
        try:
            out = []
            for j, i in enumerate(range(b * self.batch, min(len(self.paths), (b + 1) * self.batch))):
                out.append(self.sets[b & 1][j].load(self.paths[i], self.threads))
            self._result = out
        except Exception as ex:  # surfaced by the consumer
            self._error = ex

    def __iter__(self):
        if self.nb == 0:
            return
        self._read(0)
        for b in range(self.nb):
            if self._thread is not None:
                self._thread.join()
                self._thread = None
            if self._error is not None:
                raise self._error
            cur = self._result
            if b + 1 < self.nb:  # the other set of slots is free: the batch that used it has been consumed
                self._thread = threading.Thread(target=self._read, args=(b + 1,), daemon=True)
                self._thread.start()
            yield b, cur

    def close(self):
        if self._thread is not None:
            self._thread.join()
        for s in self.sets:
            for slot in s:
                _release_slot(slot)
        self.sets = []
