#!/bin/bash
# scratch (git-ignored): tests + bench + variants + ncu captures; results in gpurun_out/
mkdir -p gpurun_out
T=$1
(timeout 700 python -m pytest tests -m gpu -q -x --durations=3 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log)
tail -n 3 gpurun_out/${T}_pytest.log
summ() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open(f))
    print(f, round(d["value"]), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"]), {k:round(v["ms_per_step"],1) for k,v in list(d["kernels"].items())[:6]}, "parity", (d.get("parity") or {}).get("equal"), d["result"].get("label_sha1_16"))
except Exception as e:
    print(f, "FAILED", e)
PY
}
(timeout 400 python bench.py --steps 3 --warmup 2 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err); summ gpurun_out/${T}_bench.json
(SHARP_TRI_THREADS=512 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/${T}_bench_tri512.json 2> /dev/null); summ gpurun_out/${T}_bench_tri512.json
# ncu: full metric set of the four big kernels on the first group of the full-shape workload (2 parts: launches of 125 problems)
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rp_project_v3|hclust_tri|corrdist_kernel|sweep_nested" -c 8 -o gpurun_out/${T}_top python bench.py --parts 2 --group 1 --steps 1 --warmup 1 --no-cpu-baseline --no-serial-profile > gpurun_out/${T}_ncu.log 2>&1)
ncu -i gpurun_out/${T}_top.ncu-rep --page raw --csv > gpurun_out/${T}_raw.csv 2>/dev/null
for k in hclust_tri sweep_nested; do ncu -i gpurun_out/${T}_top.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:$k --launch-count 1 > gpurun_out/${T}_src_$k.csv 2>/dev/null; done
rm -f gpurun_out/${T}_top.ncu-rep
ls -la gpurun_out | grep ${T}
