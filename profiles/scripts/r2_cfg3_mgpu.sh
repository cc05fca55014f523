#!/bin/bash
# bash profiles/scripts/r2_cfg3_mgpu.sh TAG NGPU: BASELINE config 3 (one 100 000-cell matrix, K = 15, cell blocks dealt over the ranks)
mkdir -p gpurun_out
T=$1; N=$2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(timeout 300 $TR --master-port 29541 bench.py --gpus $N --workload cfg3 --steps 3 --warmup 2 > gpurun_out/${T}_cfg3.json 2> gpurun_out/${T}_cfg3.err)
python - gpurun_out/${T}_cfg3.json <<'PY'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith("{"):
        d=json.loads(line)
        print("cfg3 N",d["n_gpus"],"value",round(d["value"]),"ms",round(d["ms_per_step"],1),"e2e",round(d["e2e"]["value"]),round(d["e2e"]["ms_per_step"],1),"hash",d["result"]["label_sha1_16"],d["result"]["ranks_agree"])
PY
tail -n 2 gpurun_out/${T}_cfg3.err
