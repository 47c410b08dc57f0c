#!/bin/bash
# A/B of kernel variants: stage times at the bench configuration
python tools/stage_times.py --tag split_default > gpurun_out/ab1.log 2>gpurun_out/ab1.err
GRT_GEO_SPLIT=0 python tools/stage_times.py --tag fused >> gpurun_out/ab1.log 2>>gpurun_out/ab1.err
for v in gu128x4 gu256x3 gu512x1 gu128x3 gu256x2m256x4 trcp; do
  GRT_B200_LIB=_variants/lib$v.so python tools/stage_times.py --tag $v >> gpurun_out/ab1.log 2>>gpurun_out/ab1.err
done
cat gpurun_out/ab1.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
