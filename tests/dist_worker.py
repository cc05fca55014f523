"""worker of tests/test_dist_gloo.py: one process per rank, gloo backend, CPU only"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (HERE, os.path.dirname(HERE)):
    if p not in sys.path:
        sys.path.insert(0, p)


def main(rank, world, port, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from sharp_b200 import api
    import torchcomm as dist
    comm = dist.init_from_env("gloo")
    assert comm.rank == rank and comm.world == world
    # collectives
    mine = {i: np.full((i + 1, 3), float(i)) for i in range(5) if i % world == rank}
    parts = comm.allgather_parts(mine, 5)
    assert [p.shape for p in parts] == [(i + 1, 3) for i in range(5)] and all(np.all(parts[i] == i) for i in range(5))
    ints = comm.allgather_parts({i: np.arange(i, dtype=np.int32) for i in range(5) if i % world == rank}, 5)
    assert all(np.array_equal(ints[i], np.arange(i)) and ints[i].dtype == np.int32 for i in range(5))
    assert comm.bcast_obj({"a": rank} if rank == 0 else None, 0) == {"a": 0}
    assert comm.max_float(float(rank)) == world - 1
    # the sharded driver: parts dealt round-robin, centroids + labels exchanged, every rank gets the full result
    import synth
    from fakectx import FakeContext
    x, _ = synth.make_expression(600, 5 * 240, n_types=3, seed=9, sep=2.5, frac=0.5)
    plist = [np.asfortranarray(x[:, i * 240:(i + 1) * 240]) for i in range(5)]
    ctx = FakeContext()
    r = api.SHARP_unlimited(plist, viewflag=True, rN_seed=5, ctx=ctx, comm=comm, n_streams=1)
    assert len(ctx.calls) == len([i for i in range(5) if i % world == rank])   # only the own parts were clustered
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), pred=r["pred_clusters"], vie=r["viE"], k=r["N.pred_clusters"])
    comm.barrier()
    import torch.distributed as td
    td.destroy_process_group()


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
