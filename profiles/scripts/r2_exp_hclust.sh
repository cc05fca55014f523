#!/bin/bash
# hclust occupancy experiment: serial-profile kernel times per configuration (8 parts)
for cfg in "512 1" "512 2" "1024 1" "1024 2"; do
  set -- $cfg
  SHARP_TRI_THREADS=$1 SHARP_WAVE_PER_SM=$2 timeout 300 python bench.py --parts 8 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/exp_tri_$1_$2.json 2> gpurun_out/exp_tri_$1_$2.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/exp_tri_$1_$2.json"))
    k=d["kernels"]
    print("threads=$1 wave_per_sm=$2 ms_per_step=%.1f serial=%.1f hclust=%.1f corrdist=%.1f sweep=%.1f rp=%.1f" % (d["ms_per_step"], d["roofline"]["serial_step_ms"], k["hclust"]["ms_per_step"], k["corrdist"]["ms_per_step"], k["sweep_nested"]["ms_per_step"], k["rp_project"]["ms_per_step"]))
except Exception as e: print("$1 $2 failed", e)
PY
done
SHARP_HCLUST_FULL=1 timeout 300 python bench.py --parts 8 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/exp_full.json 2> gpurun_out/exp_full.err
python -c "
import json
d=json.load(open('gpurun_out/exp_full.json')); k=d['kernels']
print('old rnn kernel: ms_per_step=%.1f serial=%.1f hclust=%.1f' % (d['ms_per_step'], d['roofline']['serial_step_ms'], k['hclust']['ms_per_step']))"
