#!/bin/bash
mkdir -p gpurun_out
T=$1
summ() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open(f))
    print(f, round(d["value"]), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],1), d["config"].get("partition","")[:60], d["result"].get("label_sha1_16"))
except Exception as e:
    print(f, "FAILED", e)
PY
}
for P in 3 4; do
for V in "0 0" "$P 1" "1 2" "2 2"; do
set -- $V
(timeout 200 python bench.py --parts $P --group $1 --lanes $2 --steps 3 --warmup 2 --no-cpu-baseline --no-serial-profile > gpurun_out/${T}_p${P}_g$1_l$2.json 2> /dev/null); summ gpurun_out/${T}_p${P}_g$1_l$2.json
done; done
(timeout 500 python bench.py --workload cfg5 --steps 2 --warmup 2 > gpurun_out/${T}_cfg5.json 2> gpurun_out/${T}_cfg5.err); summ gpurun_out/${T}_cfg5.json; tail -n 3 gpurun_out/${T}_cfg5.err
