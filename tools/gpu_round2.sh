#!/bin/bash
# second-generation GPU round: tests, microbenchmarks, short ncu full capture at a development size
TAG=${1:-r1b}
mkdir -p gpurun_out
if [ "${2:-tests}" = "tests" ]; then
  python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "softening lengths\|Mmin =\|^\.\.\.done\|finished on MPI\|^$\|KDK Leapfrog\|Calculating Forces\|Timestep wall" > gpurun_out/${TAG}_gpu_tests.log
  tail -5 gpurun_out/${TAG}_gpu_tests.log
fi
if [ -x build/ubench_fp64 ]; then ./build/ubench_fp64 > gpurun_out/${TAG}_ubench_fp64.txt 2>&1; cat gpurun_out/${TAG}_ubench_fp64.txt; fi
python bench.py --steps 3 --warmup 3 --n 400000 --no-cpu > gpurun_out/${TAG}_bench_n400k.json 2> gpurun_out/${TAG}_bench_n400k.err
cat gpurun_out/${TAG}_bench_n400k.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:force_r3 -s 1 -c 1 -f -o gpurun_out/${TAG}_pair_n400k \
    python bench.py --steps 1 --warmup 1 --n 400000 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out
