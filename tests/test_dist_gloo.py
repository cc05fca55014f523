"""The N > 1 path on CPU: world_size 2, gloo backend (127.0.0.1 rendezvous).  Collectives of tests/torchcomm.py (the gloo stand-in of sharp_b200.comm.NcclComm) and
the sharded SHARP_unlimited driver (compute answered by the oracle through tests/fakectx.py): both ranks must return
the result of the single-process run."""
import os
import socket
import subprocess
import sys

import numpy as np

import synth
from fakectx import FakeContext
from sharp_b200 import api

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_world_size_2_gloo(tmp_path):
    port = free_port()
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "dist_worker.py"), str(r), "2", str(port), str(tmp_path)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=600)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    x, truth = synth.make_expression(600, 5 * 240, n_types=3, seed=9, sep=2.5, frac=0.5)
    plist = [np.asfortranarray(x[:, i * 240:(i + 1) * 240]) for i in range(5)]
    ref = api.SHARP_unlimited(plist, viewflag=True, rN_seed=5, ctx=FakeContext(), n_streams=1)
    for r in range(2):
        z = np.load(os.path.join(tmp_path, f"rank{r}.npz"))
        assert np.array_equal(z["pred"], ref["pred_clusters"]) and int(z["k"]) == ref["N.pred_clusters"]
        assert np.array_equal(z["vie"], ref["viE"])
    assert synth.ari(ref["pred_clusters"], truth) > 0.9
