#!/bin/bash
out=gpurun_out/e14; mkdir -p $out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 900 python -m pytest tests/test_gpu_stages.py -m gpu -q --no-header -p no:cacheprovider -x -k "hclust or getrowcolor" > $out/stages.log 2>&1
tail -3 $out/stages.log
timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --no-header -p no:cacheprovider -x > $out/pipeline.log 2>&1
tail -3 $out/pipeline.log
for cfg in "4 1" "4 2"; do set -- $cfg
  timeout 600 python bench.py --steps 2 --warmup 2 --group $1 --lanes $2 --streams 1 --no-cpu-baseline > $out/b_g$1_l$2.json 2> $out/b_g$1_l$2.err
  tail -c 300 $out/b_g$1_l$2.err
done
