#!/bin/bash
# Round-1 session-5 GPU call 3 (last of the round's budget): ncu --set full of the FP32 action-reaction kernel, FP32 bench line at N=2M.
TAG=${1:-r1ac}
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 ))s] $*" | tee -a $O/${TAG}_timeline.txt; }
stamp "ncu --set full, one launch of the FP32 sym kernel at N=400k"
timeout 35 ncu --set full --clock-control none --import-source on -k regex:force_r3_f32_sym -s 1 -c 1 -o $O/${TAG}_sym_f32_n400k \
    python tools/ncu_f32_sym.py 400000 > $O/${TAG}_ncu_full.out 2>&1
tail -2 $O/${TAG}_ncu_full.out | cut -c1-300
stamp "FP32 bench line, C5 shape at N=2M"
timeout 42 python bench.py --config c5 --n 2000000 --steps 2 --warmup 3 --no-cpu > $O/${TAG}_bench_c5shape_n2m.json 2> $O/${TAG}_bench_c5shape_n2m.err
cut -c1-400 $O/${TAG}_bench_c5shape_n2m.json; tail -2 $O/${TAG}_bench_c5shape_n2m.err
stamp "done"
