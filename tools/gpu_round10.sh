#!/bin/bash
# Round-1 session-4 final 1-GPU evidence run with the production defaults (action-reaction path, shape 3).
TAG=${1:-r1w}
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 ))s] $*" | tee -a $O/${TAG}_timeline.txt; }
stamp "pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q --timeout 400 > $O/${TAG}_gpu_tests.log 2>&1
echo "rc=$?" >> $O/${TAG}_gpu_tests.log
tail -4 $O/${TAG}_gpu_tests.log
stamp "smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log
stamp "bench reference arm"
timeout 420 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref_c2.json 2> $O/${TAG}_bench_ref_c2.err
cut -c1-400 $O/${TAG}_bench_ref_c2.json
stamp "bench"
timeout 420 python bench.py --steps 3 --warmup 3 > $O/${TAG}_bench_c2_1gpu.json 2> $O/${TAG}_bench_c2_1gpu.err
cut -c1-600 $O/${TAG}_bench_c2_1gpu.json; tail -3 $O/${TAG}_bench_c2_1gpu.err
stamp "one-sided bench for the record (STEPS_B200_SYM=0)"
STEPS_B200_SYM=0 timeout 420 python bench.py --steps 2 --warmup 3 --no-cpu > $O/${TAG}_bench_c2_1gpu_onesided.json 2> $O/${TAG}_bench_c2_1gpu_onesided.err
cut -c1-300 $O/${TAG}_bench_c2_1gpu_onesided.json
stamp "ncu launch list of the bench command"
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/${TAG}_ncu_launches_bench_c2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $O/${TAG}_ncu_launches_bench.out 2>&1
tail -2 $O/${TAG}_ncu_launches_bench.out | cut -c1-300
stamp "ncu dram traffic + pipe counters, every pass of one force evaluation at C2"
timeout 420 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed.sum \
    --clock-control none -k regex:"force_r3_f64|reduce_sym" -c 16 --csv --log-file $O/${TAG}_ncu_pair_c2.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu > $O/${TAG}_ncu_pair_c2.out 2>&1
tail -2 $O/${TAG}_ncu_pair_c2.out | cut -c1-300
stamp "ncu --set full, one launch of the sym kernel at N=400k"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:force_r3_f64_sym -s 1 -c 1 -o $O/${TAG}_sym_n400k \
    python bench.py --steps 1 --warmup 1 --n 400000 --no-cpu > $O/${TAG}_ncu_full.out 2>&1
tail -2 $O/${TAG}_ncu_full.out | cut -c1-300
stamp "other configurations (C1 shape, FP32, S^1xR^2, T^3): pair-kernel throughput"
timeout 300 python tools/topo_bench.py c1,r3f32:2000000,s1r2nl:400000,s1r2:200000,t3:64 > $O/${TAG}_topo_bench.txt 2> $O/${TAG}_topo_bench.err
cut -c1-330 $O/${TAG}_topo_bench.txt; tail -2 $O/${TAG}_topo_bench.err
stamp "done"
