#!/bin/bash
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
for V in "0 0" "2 2" "1 2" "2 3" "3 2"; do
set -- $V
(timeout 300 python bench.py --group $1 --lanes $2 --steps 2 --warmup 2 --no-cpu-baseline --no-serial-profile > gpurun_out/${T}_g$1_l$2.json 2> /dev/null); summ gpurun_out/${T}_g$1_l$2.json
done
