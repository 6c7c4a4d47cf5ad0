#!/bin/bash
# round-1 session-3 GPU call: full gpu test tier, issue-cost microbenchmarks, kernel-shape sweep, ncu full capture, bench
TAG=${1:-r1f}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_gpu_tests.log
tail -3 gpurun_out/${TAG}_gpu_tests.log
./build/ubench2_fp64 > gpurun_out/${TAG}_ubench2_fp64.txt 2>&1
cat gpurun_out/${TAG}_ubench2_fp64.txt
python tools/sweep_f64.py 400000 0,7,10,16,17,19,1 > gpurun_out/${TAG}_variant_sweep_n400k.txt 2>&1
cat gpurun_out/${TAG}_variant_sweep_n400k.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:force_r3 -s 1 -c 1 -f -o gpurun_out/${TAG}_pair_n400k \
    python bench.py --steps 1 --warmup 1 --n 400000 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_c2_1gpu.json 2> gpurun_out/${TAG}_bench_c2_1gpu.err
cat gpurun_out/${TAG}_bench_c2_1gpu.json
tail -5 gpurun_out/${TAG}_bench_c2_1gpu.err
ls -la gpurun_out
