#!/bin/bash
TAG=${1:-r1h}
mkdir -p gpurun_out
python tools/sweep_f64.py 400000 0,20,21 > gpurun_out/${TAG}_variant_sweep_n400k.txt 2>&1
cat gpurun_out/${TAG}_variant_sweep_n400k.txt
python tools/topo_bench.py c1,r3f32:1000000,s1r2nl:200000,s1r2:200000,t3:40 > gpurun_out/${TAG}_topo_bench.txt 2> gpurun_out/${TAG}_topo_bench.err
cat gpurun_out/${TAG}_topo_bench.txt; tail -3 gpurun_out/${TAG}_topo_bench.err
python -c "
import steps_b200 as sb
print('fp64 peak burst', sb.fma_peak(0, 8), 'sustained', sb.fma_peak_sustained(0, 8, 2.0))
print('fp32 peak burst', sb.fma_peak(0, 4), 'sustained', sb.fma_peak_sustained(0, 4, 2.0))
" > gpurun_out/${TAG}_peaks.txt 2>&1
cat gpurun_out/${TAG}_peaks.txt
