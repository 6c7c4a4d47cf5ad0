#!/bin/bash
# run on the GPU box: the gpu test tier, log kept under gpurun_out/
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "softening lengths\|Mmin =\|^\.\.\.done\|finished on MPI\|^$\|KDK Leapfrog\|Calculating Forces\|Timestep wall" | tee gpurun_out/gpu_tests.log | tail -60
