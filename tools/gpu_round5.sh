#!/bin/bash
TAG=${1:-r1i}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "s1r2 or periodic or kdk or shim" 2>&1 | tail -8 > gpurun_out/${TAG}_gpu_tests_s1r2.log
cat gpurun_out/${TAG}_gpu_tests_s1r2.log
python tools/topo_bench.py s1r2nl:200000,s1r2nl:400000 > gpurun_out/${TAG}_topo_bench.txt 2> gpurun_out/${TAG}_topo_bench.err
cat gpurun_out/${TAG}_topo_bench.txt; tail -3 gpurun_out/${TAG}_topo_bench.err
