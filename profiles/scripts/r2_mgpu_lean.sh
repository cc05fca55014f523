#!/bin/bash
# bash profiles/scripts/r2_mgpu_lean.sh TAG NGPU [extra bench args]: one bench run with the host-side trace on
mkdir -p gpurun_out
T=$1; N=$2; shift; shift
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(SHARP_B200_TRACE=1 timeout 400 $TR --master-port 29521 bench.py --gpus $N --steps 3 --warmup 2 "$@" > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err)
python - gpurun_out/${T}_bench.json <<'PY'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith("{"):
        d=json.loads(line)
        print("N",d["n_gpus"],"value",round(d["value"]),"ms",round(d["ms_per_step"],1),"e2e",round(d["e2e"]["value"]),round(d["e2e"]["ms_per_step"],1),"hash",d["result"]["label_sha1_16"],d["result"]["ranks_agree"],d["config"].get("host_affinity_rank0"), "h2d", d["e2e"].get("h2d_gbs_measured"))
PY
grep "trace py\] ranM\|trace py\] fused" gpurun_out/${T}_bench.err | tail -n 6
