#!/bin/bash
# run on the GPU box (via gpurun): gpu test tier, 1-GPU bench, ncu launch list + one full capture of the pair kernel.
# usage: tools/gpu_round.sh <tag> [tests|notests] [ncu|noncu]
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
if [ "${2:-tests}" = "tests" ]; then
  python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "softening lengths\|Mmin =\|^\.\.\.done\|finished on MPI\|^$\|KDK Leapfrog\|Calculating Forces\|Timestep wall" > gpurun_out/${TAG}_gpu_tests.log
  tail -5 gpurun_out/${TAG}_gpu_tests.log
fi
python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> /dev/null
cat gpurun_out/${TAG}_bench_ref.json
if [ "${3:-ncu}" = "ncu" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:force_r3 -s 1 -c 1 -f -o gpurun_out/${TAG}_pair \
      python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
  ls -la gpurun_out
fi
