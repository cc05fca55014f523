"""TEST INFRASTRUCTURE: a torch.distributed (gloo) stand-in for sharp_b200.comm.NcclComm, so that the N > 1 HOST logic of the
sharded drivers (which parts / blocks a rank owns, what is exchanged, that every rank returns the full result) runs on a
machine without GPUs: world_size 2, backend "gloo".  Same interface as NcclComm (rank, world, barrier, bcast_obj,
allgather_parts, max_float).  The product package does not import torch: its communicator is NCCL behind the C ABI
(sharp_comm_*, sharp_b200/comm.py)."""
from __future__ import annotations

import os

import numpy as np


class Comm:
    """Thin wrapper over an initialised ``torch.distributed`` process group."""

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (call init_from_env() first)")
        self._torch, self._dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.backend = dist.get_backend()
        self.device = torch.device("cuda", device if device is not None else torch.cuda.current_device()) \
            if self.backend == "nccl" else torch.device("cpu")

    def barrier(self):
        self._dist.barrier()

    def bcast_obj(self, obj, src=0):
        box = [obj]
        self._dist.broadcast_object_list(box, src=src, device=self.device)
        return box[0]

    def allgather_bytes(self, payload: bytes) -> list[bytes]:
        """variable-length allgather: sizes first, then one padded all_gather_into_tensor"""
        torch, dist = self._torch, self._dist
        n = torch.tensor([len(payload)], dtype=torch.int64, device=self.device)
        sizes = torch.empty(self.world, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(sizes, n)
        sizes = sizes.cpu().tolist()
        mx = max(max(sizes), 1)
        buf = torch.zeros(mx, dtype=torch.uint8)
        if len(payload):
            buf[:len(payload)] = torch.frombuffer(bytearray(payload), dtype=torch.uint8)
        buf = buf.to(self.device)
        out = torch.empty(self.world * mx, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(out, buf)
        out = out.cpu().numpy()
        return [out[r * mx:r * mx + sizes[r]].tobytes() for r in range(self.world)]

    def allgather_parts(self, mine: dict, nparts: int) -> list:
        """every rank contributes the arrays of the parts it owns ({part index: ndarray}); returns the list of
        all ``nparts`` arrays on every rank."""
        idx = sorted(mine)
        metas = [(i, mine[i].dtype.str, mine[i].shape) for i in idx]
        blob = b"".join(np.ascontiguousarray(mine[i]).tobytes() for i in idx)
        all_meta = [None] * self.world
        self._dist.all_gather_object(all_meta, metas)
        blobs = self.allgather_bytes(blob)
        out = [None] * nparts
        for r in range(self.world):
            off = 0
            for i, dt, shape in all_meta[r]:
                cnt = int(np.prod(shape)) * np.dtype(dt).itemsize
                out[i] = np.frombuffer(blobs[r], dtype=np.dtype(dt), count=int(np.prod(shape)), offset=off).reshape(shape).copy()
                off += cnt
        missing = [i for i, a in enumerate(out) if a is None]
        if missing:
            raise RuntimeError(f"allgather_parts: no rank contributed parts {missing}")
        return out

    def max_float(self, x: float) -> float:
        t = self._torch.tensor([x], dtype=self._torch.float64, device=self.device)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX)
        return float(t.item())


def init_from_env(backend: str | None = None) -> Comm | None:
    """Initialise from torchrun's environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*); None when single."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return None
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=int(os.environ["RANK"]), world_size=world,
                                **({"device_id": torch.device("cuda", local)} if backend == "nccl" else {}))
    return Comm(local if backend == "nccl" else None)
