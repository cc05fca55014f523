"""One-process-per-GPU plumbing without PyTorch: the communicator lives behind the C ABI (``sharp_comm_*``: NCCL over
NVLink, loaded by libsharpb200.so at run time); this module only carries the 128-byte NCCL unique id from rank 0 to the
other ranks -- a TCP rendezvous on ``MASTER_ADDR`` (what an R host would do with ``socketConnection``) -- and wraps the
host-buffer collectives the sharded drivers use.

The path shards naturally: parts and cell blocks are independent until the meta-clustering steps, so
the only exchanges are allgathers of labels, cluster counts and reduced-space rows / centroids.

Environment (what ``torchrun`` / ``python -m torch.distributed.run`` exports; any launcher may set the same variables):
``RANK``, ``WORLD_SIZE``, ``LOCAL_RANK``, ``MASTER_ADDR``, ``MASTER_PORT``.  The rendezvous listens on
``MASTER_PORT + 1 ...`` (the launcher's own store owns ``MASTER_PORT``).
"""
from __future__ import annotations

import ctypes as C
import os
import pickle
import socket
import struct
import time

import numpy as np

from . import _lib

_MAGIC = b"SHARPB200"
_PORT_SPAN = 16


def _recv_exact(sock, n: int) -> bytes:
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(n - len(buf))
        if not chunk:
            raise ConnectionError("rendezvous: peer closed the connection")
        buf += chunk
    return bytes(buf)


def exchange_from_root(payload: bytes | None, rank: int, world: int, addr: str, port: int, timeout: float = 300.0,
                       token: bytes = b"") -> bytes:
    """Rank 0 hands ``payload`` to every other rank over TCP (star topology, one short-lived connection per rank).
    ``token`` distinguishes concurrent jobs that share a host and a port range."""
    if world <= 1:
        return payload or b""
    hello = _MAGIC + struct.pack("<ii", world, len(token)) + token
    if rank == 0:
        srv = None
        for off in range(_PORT_SPAN):
            try:
                srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
                srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
                srv.bind((addr if addr not in ("localhost",) else "127.0.0.1", port + off))
                break
            except OSError:
                srv.close()
                srv = None
        if srv is None:
            raise OSError(f"rendezvous: no free port in {port}..{port + _PORT_SPAN - 1}")
        srv.listen(world)
        srv.settimeout(timeout)
        served = set()
        try:
            while len(served) < world - 1:
                conn, _ = srv.accept()
                with conn:
                    conn.settimeout(timeout)
                    try:
                        got = _recv_exact(conn, len(hello) + 4)
                    except (ConnectionError, socket.timeout):
                        continue
                    if got[:len(hello)] != hello:
                        continue  # somebody else's job probing the port range
                    r = struct.unpack("<i", got[len(hello):])[0]
                    conn.sendall(struct.pack("<q", len(payload)) + payload)
                    served.add(r)
        finally:
            srv.close()
        return payload
    deadline = time.time() + timeout
    off = 0
    while True:
        try:
            with socket.create_connection((addr, port + off), timeout=3.0) as s:
                s.settimeout(3.0)   # rank 0 answers at once; a silent listener on this port is somebody else's (try the next one)
                s.sendall(hello + struct.pack("<i", rank))
                n = struct.unpack("<q", _recv_exact(s, 8))[0]
                return _recv_exact(s, n)
        except (OSError, ConnectionError, struct.error):
            if time.time() > deadline:
                raise TimeoutError(f"rendezvous: rank {rank} could not reach rank 0 at {addr}:{port}..{port + _PORT_SPAN - 1}")
            off = (off + 1) % _PORT_SPAN
            if off == 0:
                time.sleep(0.05)


class NcclComm:
    """The communicator of one rank: ``sharp_comm_init`` on the context's device, host-buffer collectives through it.
    Same interface as the torch-based test communicator (rank, world, barrier, bcast_obj, allgather_parts, max_float),
    so the drivers of ``sharp_b200.api`` take either."""

    def __init__(self, ctx, rank: int, world: int, unique_id: bytes):
        self.ctx, self.rank, self.world = ctx, int(rank), int(world)
        self.backend = "nccl"
        uid = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        _lib._check(_lib.load().sharp_comm_init(ctx._h, uid, self.rank, self.world))

    def close(self):
        if self.ctx is not None and self.ctx._h:
            _lib.load().sharp_comm_destroy(self.ctx._h)
        self.ctx = None

    # -- collectives on host buffers -----------------------------------------------------------------
    def allgather_arrays(self, payload: np.ndarray) -> list[np.ndarray]:
        """payload: a uint8 array (any length) per rank -> the list of every rank's array (views into ONE receive
        buffer; segments start on 8-byte boundaries)"""
        lib = _lib.load()
        payload = np.ascontiguousarray(payload, dtype=np.uint8)
        padded = (len(payload) + 7) & ~7
        one = np.array([len(payload)], dtype=np.int64)
        sizes = np.zeros(self.world, dtype=np.int64)
        eight = np.full(self.world, 8, dtype=np.int64)
        _lib._check(lib.sharp_comm_allgatherv(self.ctx._h, one.ctypes.data_as(C.c_void_p), eight.ctypes.data_as(C.POINTER(C.c_int64)),
                                              sizes.ctypes.data_as(C.c_void_p)))
        seg = (sizes + 7) & ~7
        total = int(seg.sum())
        if padded != len(payload) or padded == 0:
            send = np.zeros(max(padded, 8), dtype=np.uint8)
            send[:len(payload)] = payload
        else:
            send = payload
        recv = np.empty(max(total, 8), dtype=np.uint8)
        _lib._check(lib.sharp_comm_allgatherv(self.ctx._h, send.ctypes.data_as(C.c_void_p), seg.ctypes.data_as(C.POINTER(C.c_int64)),
                                              recv.ctypes.data_as(C.c_void_p)))
        out, off = [], 0
        for r in range(self.world):
            out.append(recv[off:off + int(sizes[r])])
            off += int(seg[r])
        return out

    def allgather_bytes(self, payload: bytes) -> list[bytes]:
        return [a.tobytes() for a in self.allgather_arrays(np.frombuffer(payload, dtype=np.uint8))]

    def barrier(self):
        _lib._check(_lib.load().sharp_comm_barrier(self.ctx._h))

    def bcast_obj(self, obj, src=0):
        blob = pickle.dumps(obj) if self.rank == src else b""
        n = np.array([len(blob)], dtype=np.int64)
        lib = _lib.load()
        _lib._check(lib.sharp_comm_bcast(self.ctx._h, n.ctypes.data_as(C.c_void_p), C.c_int64(8), int(src)))
        buf = np.frombuffer(blob, dtype=np.uint8).copy() if self.rank == src else np.zeros(int(n[0]), dtype=np.uint8)
        if int(n[0]):
            _lib._check(lib.sharp_comm_bcast(self.ctx._h, buf.ctypes.data_as(C.c_void_p), C.c_int64(int(n[0])), int(src)))
        return pickle.loads(buf.tobytes())

    def allgather_parts(self, mine: dict, nparts: int) -> list:
        """every rank contributes the arrays of the parts it owns ({part index: ndarray}); returns the list of all
        ``nparts`` arrays on every rank.  Fixed binary layout per array: index, dtype code, ndim, shape, bytes."""
        return unpack_parts(self.allgather_arrays(pack_parts(mine)), nparts)

    def max_float(self, x: float) -> float:
        vals = [struct.unpack("<d", b)[0] for b in self.allgather_bytes(struct.pack("<d", float(x)))]
        return max(vals)


_DT = {"<f8": 0, "<i4": 1, "<i8": 2, "|u1": 3, "<f4": 4}
_DT_INV = {v: k for k, v in _DT.items()}


def pack_parts(mine: dict) -> np.ndarray:
    """{index: ndarray} -> one uint8 array.  Per array: int32 index, dtype code, ndim, 0; int64 shape[ndim]; the data,
    padded to 8 bytes -- so every array starts on an 8-byte boundary of the blob and unpacking needs no copy."""
    pieces = []
    for i in sorted(mine):
        a = np.ascontiguousarray(mine[i])
        head = struct.pack("<iiii", int(i), _DT[a.dtype.str], a.ndim, 0) + struct.pack(f"<{a.ndim}q", *a.shape)
        pieces.append(np.frombuffer(head, dtype=np.uint8))
        pieces.append(a.reshape(-1).view(np.uint8))
        tail = (-a.nbytes) & 7
        if tail:
            pieces.append(np.zeros(tail, dtype=np.uint8))
    return np.concatenate(pieces) if pieces else np.zeros(0, dtype=np.uint8)


def unpack_parts(blobs: list, nparts: int) -> list:
    """the arrays of all ranks (views into the blobs, which stay alive with them)"""
    out = [None] * nparts
    for blob in blobs:
        blob = np.frombuffer(blob, dtype=np.uint8) if isinstance(blob, (bytes, bytearray)) else blob
        off = 0
        while off < len(blob):
            i, code, nd, _ = struct.unpack_from("<iiii", blob, off)
            off += 16
            shape = struct.unpack_from(f"<{nd}q", blob, off)
            off += 8 * nd
            dt = np.dtype(_DT_INV[code])
            cnt = int(np.prod(shape)) if nd else 1
            nb = cnt * dt.itemsize
            seg = blob[off:off + nb]
            out[i] = (seg.view(dt) if seg.ctypes.data % dt.itemsize == 0 else np.frombuffer(seg.tobytes(), dtype=dt)).reshape(shape)
            off += (nb + 7) & ~7
    missing = [i for i, a in enumerate(out) if a is None]
    if missing:
        raise RuntimeError(f"allgather_parts: no rank contributed parts {missing}")
    return out


def init_from_env(ctx=None) -> NcclComm | None:
    """RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT -> a communicator on ``ctx`` (default: the module-level
    context of device LOCAL_RANK); None for a single process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return None
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if ctx is None:
        from . import api
        api.set_devices(local)
        ctx = api.get_context(local)
    addr = os.environ.get("MASTER_ADDR", "127.0.0.1")
    port = int(os.environ.get("MASTER_PORT", "29500")) + 1
    uid = None
    if rank == 0:
        buf = (C.c_ubyte * 128)()
        _lib._check(_lib.load().sharp_comm_unique_id(buf, 128))
        uid = bytes(buf)
    token = os.environ.get("TORCHELASTIC_RUN_ID", "").encode()[:32]
    uid = exchange_from_root(uid, rank, world, addr, port, token=token)
    return NcclComm(ctx, rank, world, uid)
