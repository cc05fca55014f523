#!/bin/bash
# bash profiles/scripts/r2_final_n1.sh TAG -- the single-GPU evidence of the round: tests, bench line, reference arm,
# ncu launch list, ncu --set full of the four big kernels, compute-sanitizer, config 2.  Results in gpurun_out/.
mkdir -p gpurun_out
T=$1
(timeout 600 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gpu.log)
tail -n 2 gpurun_out/${T}_pytest_gpu.log
(timeout 500 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err); tail -c 300 gpurun_out/${T}_bench.err
(timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> /dev/null)
(timeout 200 python bench.py --workload cfg2 --steps 3 --warmup 2 > gpurun_out/${T}_cfg2.json 2> /dev/null)
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${T}_ncu_launches_parts8.csv python bench.py --parts 8 --steps 1 --warmup 1 --no-cpu-baseline --no-serial-profile > /dev/null 2>&1)
(timeout 500 ncu --set full --clock-control none --import-source on -k regex:"rp_project_v3|hclust_tri|corrdist_kernel|sweep_nested" -c 8 -o gpurun_out/${T}_top python bench.py --parts 2 --group 1 --steps 1 --warmup 1 --no-cpu-baseline --no-serial-profile > gpurun_out/${T}_ncu.log 2>&1)
ncu -i gpurun_out/${T}_top.ncu-rep --page raw --csv > gpurun_out/${T}_ncu_full_raw.csv 2>/dev/null
ncu -i gpurun_out/${T}_top.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:rp_project_v3 --launch-count 1 > gpurun_out/${T}_src_rp.csv 2>/dev/null
rm -f gpurun_out/${T}_top.ncu-rep
(timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_stages.py -q -x -k "round_parallel or count_classes or record_overflow or opt_hclust_feature or smetac" > gpurun_out/${T}_sanitizer_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_sanitizer_memcheck.log)
tail -n 4 gpurun_out/${T}_sanitizer_memcheck.log
(timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_stages.py -q -x -k "opt_hclust_round_parallel and not ties" > gpurun_out/${T}_sanitizer_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_sanitizer_racecheck.log)
tail -n 4 gpurun_out/${T}_sanitizer_racecheck.log
python - <<PY
import json
for f in ("gpurun_out/${T}_bench.json","gpurun_out/${T}_bench_reference.json","gpurun_out/${T}_cfg2.json"):
    try:
        d=json.load(open(f)); print(f, round(d["value"]), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"]), (d.get("parity") or {}).get("equal"), (d.get("roofline") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f,"FAILED",e)
PY
ls -la gpurun_out | grep ${T}
