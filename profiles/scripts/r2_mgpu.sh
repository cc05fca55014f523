#!/bin/bash
# scratch multi-GPU experiment driver (git-ignored): bash tests/_exp_mgpu.sh TAG NGPU
mkdir -p gpurun_out
T=$1; N=$2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(timeout 300 $TR --master-port 29511 tests/mgpu_check.py > gpurun_out/${T}_mgpu.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_mgpu.log)
grep "mgpu\|rc=" gpurun_out/${T}_mgpu.log | tail -n 6
(timeout 400 $TR --master-port 29521 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err)
(SHARP_B200_TRACE=1 timeout 300 $TR --master-port 29531 bench.py --gpus $N --steps 2 --warmup 2 --no-serial-profile > gpurun_out/${T}_trace.json 2> gpurun_out/${T}_trace.txt)
python - gpurun_out/${T}_bench.json <<'PY'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith("{"):
        d=json.loads(line)
        print("N",d["n_gpus"],"value",round(d["value"]),"ms",round(d["ms_per_step"],1),"e2e",round(d["e2e"]["value"]),round(d["e2e"]["ms_per_step"],1),"hash",d["result"]["label_sha1_16"],d["result"]["ranks_agree"])
PY
