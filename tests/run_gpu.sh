#!/bin/bash
# Runs the GPU test groups in separate processes (a hung kernel only loses its own group). Usage: tests/run_gpu.sh [outdir]
out=${1:-gpurun_out}
mkdir -p $out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $out/gpu.txt 2>&1
for grp in "rp_project" "corrdist" "hclust" "opt_hclust or getrowcolor" "wmetac" "smetac"; do
  name=$(echo $grp | tr ' ' '_')
  timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -q -k "$grp" --no-header -p no:cacheprovider > $out/stage_$name.log 2>&1
  echo "[$grp] exit $?" | tee -a $out/summary.txt
  tail -n 3 $out/stage_$name.log | tee -a $out/summary.txt
done
timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --no-header -p no:cacheprovider > $out/pipeline.log 2>&1
echo "[pipeline] exit $?" | tee -a $out/summary.txt
tail -n 3 $out/pipeline.log | tee -a $out/summary.txt
