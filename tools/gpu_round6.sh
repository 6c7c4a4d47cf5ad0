#!/bin/bash
TAG=${1:-r1k}
mkdir -p gpurun_out
for v in 0 1 2 3 4 5; do
  STEPS_B200_S1R2_VARIANT=$v python tools/topo_bench.py s1r2nl:200000 2>/dev/null | grep "^{" | sed "s/^{/{\"variant\": $v, /" >> gpurun_out/${TAG}_s1r2_variants.txt
done
cat gpurun_out/${TAG}_s1r2_variants.txt
