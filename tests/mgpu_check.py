"""Multi-GPU check of the block-sharded paths (run under torchrun / any launcher that sets RANK, WORLD_SIZE, LOCAL_RANK,
MASTER_ADDR, MASTER_PORT; one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py

Every rank runs the SAME calls with a communicator (NCCL behind the C ABI, sharp_b200.comm) and then the plain
single-GPU call; the label vectors must be identical: (1) SHARP() -> SHARP_large with the cell blocks dealt over the
ranks, shuffled (n < 1e5, every rank holds the matrix) and un-shuffled (n >= 1e5, host data: only the rank's columns
are uploaded); (2) SHARP_unlimited with whole parts dealt round-robin and the left-over part block-sharded."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (HERE, os.path.dirname(HERE)):
    if p not in sys.path:
        sys.path.insert(0, p)

import synth  # noqa: E402
from sharp_b200 import api, comm as sharp_comm  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    api.set_devices(local)
    ctx = api.get_context(local)
    comm = sharp_comm.init_from_env(ctx)
    assert comm is not None, "run with WORLD_SIZE > 1"
    rank, world = comm.rank, comm.world
    ok = True

    def same(a, b, what):
        nonlocal ok
        eq = bool(np.array_equal(a, b))
        hashes = comm.allgather_bytes(np.ascontiguousarray(a, dtype=np.int32).tobytes()[:64])
        agree = len(set(hashes)) == 1
        if rank == 0:
            print(f"[mgpu] {what}: sharded == single-GPU: {eq}; ranks agree: {agree}; clusters: {len(np.unique(a))}", flush=True)
        ok = ok and eq and agree

    # (1a) shuffled: 12 000 cells, 6 blocks of 2000
    x, _ = synth.make_expression(800, 12000, n_types=5, seed=3, kind="umi", zero_frac=0.8, sep=2.0, frac=0.4)
    csc = synth.to_csc(x) + (x.shape,)
    a = api.SHARP(csc, exp_type="UMI", ensize_K=3, rN_seed=7, logflag=False, forview=True, ctx=ctx, comm=comm)
    b = api.SHARP(csc, exp_type="UMI", ensize_K=3, rN_seed=7, logflag=False, forview=True, ctx=ctx)
    same(a["pred_clusters"], b["pred_clusters"], "SHARP_large, shuffled, 6 blocks")
    assert np.array_equal(a["viE"], b["viE"])
    # (1b) un-shuffled: 100 400 cells (>= 1e5), 51 blocks, host CSC: each rank uploads only its columns
    x, _ = synth.make_expression(400, 100400, n_types=6, seed=4, kind="umi", zero_frac=0.7, sep=2.0, frac=0.4)
    csc = synth.to_csc(x) + (x.shape,)
    a = api.SHARP(csc, exp_type="UMI", ensize_K=2, rN_seed=7, logflag=False, forview=False, ctx=ctx, comm=comm)
    b = api.SHARP(csc, exp_type="UMI", ensize_K=2, rN_seed=7, logflag=False, forview=False, ctx=ctx)
    same(a["pred_clusters"], b["pred_clusters"], "SHARP_large, un-shuffled, 51 blocks, column slices")
    # (2) SHARP_unlimited: 2 * world + 1 parts -> the last one is block-sharded
    sizes = [10000 + 150 * i for i in range(2 * world + 1)]
    x, _ = synth.make_expression(700, sum(sizes), n_types=5, seed=21, kind="umi", zero_frac=0.8, sep=2.0, frac=0.4)
    parts, o = [], 0
    for n in sizes:
        parts.append(synth.to_csc(np.asfortranarray(x[:, o:o + n])) + ((700, n),))
        o += n
    a = api.SHARP_unlimited(parts, viewflag=False, rN_seed=31, ensize_K=3, exp_type="UMI", ctx=ctx, comm=comm)
    b = api.SHARP_unlimited(parts, viewflag=False, rN_seed=31, ensize_K=3, exp_type="UMI", ctx=ctx)
    same(a["pred_clusters"], b["pred_clusters"], f"SHARP_unlimited, {len(sizes)} parts on {world} ranks, last part block-sharded")
    comm.barrier()
    comm.close()
    if rank == 0:
        print("[mgpu] ALL OK" if ok else "[mgpu] MISMATCH", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
