#!/bin/bash
# bash profiles/scripts/r2_grouping_small.sh TAG -- rank-sized workloads (3 and 7 parts), more lanes
mkdir -p gpurun_out
T=$1
summ() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open(f))
    print(f, round(d["value"]), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],1), d["result"].get("label_sha1_16"))
except Exception as e:
    print(f, "FAILED", e)
PY
}
for V in "3 0 0" "3 1 3" "3 2 3" "7 0 0" "7 1 3" "7 2 3" "7 3 3"; do
set -- $V
(timeout 120 python bench.py --parts $1 --group $2 --lanes $3 --steps 3 --warmup 2 --no-cpu-baseline --no-serial-profile > gpurun_out/${T}_p$1_g$2_l$3.json 2> /dev/null); summ gpurun_out/${T}_p$1_g$2_l$3.json
done
