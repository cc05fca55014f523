#!/bin/bash
# bash profiles/scripts/r2_check.sh TAG -- GPU tests + smoke + the default bench line
mkdir -p gpurun_out
T=$1
(timeout 600 python -m pytest tests -m gpu -q --durations=3 > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gpu.log)
tail -n 2 gpurun_out/${T}_pytest_gpu.log
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1); tail -n 1 gpurun_out/${T}_smoke.log
(timeout 500 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err)
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
print(round(d["value"]), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],1), d["parity"]["equal"], d["result"]["label_sha1_16"], {k:round(v["ms_per_step"],1) for k,v in list(d["kernels"].items())[:8]})
PY
