#!/bin/bash
mkdir -p gpurun_out
T=$1
(timeout 700 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log)
tail -n 3 gpurun_out/${T}_pytest.log
summ() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open(f))
    print(f, round(d["value"]), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"]), {k:round(v["ms_per_step"],1) for k,v in list((d.get("kernels") or {}).items())[:7]}, "parity", (d.get("parity") or {}).get("equal"), d["result"].get("label_sha1_16"))
except Exception as e:
    print(f, "FAILED", e)
PY
}
(timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err); summ gpurun_out/${T}_bench.json
(SHARP_RP_UNCOND=1 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/${T}_bench_unc.json 2> gpurun_out/${T}_bench_unc.err); summ gpurun_out/${T}_bench_unc.json
(timeout 500 python bench.py --workload cfg5 --steps 2 --warmup 1 > gpurun_out/${T}_cfg5.json 2> gpurun_out/${T}_cfg5.err); summ gpurun_out/${T}_cfg5.json; tail -n 3 gpurun_out/${T}_cfg5.err
(timeout 300 python bench.py --workload cfg3 --steps 2 --warmup 2 > gpurun_out/${T}_cfg3.json 2> gpurun_out/${T}_cfg3.err); summ gpurun_out/${T}_cfg3.json; tail -n 2 gpurun_out/${T}_cfg3.err
df -h /dev/shm /tmp | tail -n 2
