"""sharp_b200/comm.py without a GPU: the TCP rendezvous that carries the NCCL unique id from rank 0 to the other ranks
(three processes on 127.0.0.1, a busy first port, a foreign client probing the port range) and the fixed binary layout of
the part exchange.  The collectives themselves are NCCL behind the C ABI and run in the -m gpu / multi-GPU bench legs."""
import multiprocessing as mp
import os
import socket

import numpy as np

from sharp_b200 import comm


def _worker(rank, world, port, q, token):
    payload = bytes(range(128)) if rank == 0 else None
    got = comm.exchange_from_root(payload, rank, world, "127.0.0.1", port, timeout=60.0, token=token)
    q.put((rank, got))


def _free_port_block(span=20):
    while True:
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        p = s.getsockname()[1]
        s.close()
        if p + span < 65000:
            return p


def test_rendezvous_three_ranks_with_a_busy_port_and_a_stranger():
    port = _free_port_block()
    blocker = socket.socket()                      # the launcher's own store sits on the first port
    blocker.bind(("127.0.0.1", port))
    blocker.listen(1)
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_worker, args=(r, 3, port, q, b"job-A")) for r in range(3)]
    for p in procs:
        p.start()
    # a client of ANOTHER job (different token) must be ignored by rank 0
    try:
        comm.exchange_from_root(None, 1, 3, "127.0.0.1", port, timeout=1.5, token=b"job-B")
        raise AssertionError("a foreign job received the id")
    except TimeoutError:
        pass
    res = dict(q.get(timeout=90) for _ in range(3))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    blocker.close()
    assert res[0] == res[1] == res[2] == bytes(range(128))


def test_part_exchange_layout_round_trip():
    mine0 = {0: np.arange(12, dtype=np.int32).reshape(3, 4), 2: np.linspace(0, 1, 5)}
    mine1 = {1: np.array([7], dtype=np.int64), 3: np.zeros((0, 6))}
    out = comm.unpack_parts([comm.pack_parts(mine0), comm.pack_parts(mine1)], 4)
    assert np.array_equal(out[0], mine0[0]) and out[0].dtype == np.int32
    assert np.array_equal(out[1], mine1[1]) and out[1].dtype == np.int64
    assert np.array_equal(out[2], mine0[2]) and out[3].shape == (0, 6)
    try:
        comm.unpack_parts([comm.pack_parts(mine0)], 4)
        raise AssertionError("missing parts must be reported")
    except RuntimeError:
        pass


def test_single_process_has_no_communicator(monkeypatch):
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    assert comm.init_from_env() is None
