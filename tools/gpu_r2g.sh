#!/bin/bash
# round 2, call g: T^3 action-reaction kernel with a 2-tile window (more L1 for the gather): shapes; C2 DRAM traffic of one evaluation; launch list
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
for v in 5 3 4 0; do
  STEPS_B200_GEN_SYM_VARIANT=$v timeout 300 python tools/topo_bench.py t3:64,t3:48 2>&1 | grep "^{" | cut -c1-330
done | tee $O/r2g_t3_window2_sweep.txt
timeout 300 python tools/topo_bench.py s1r2:200000 2>&1 | grep "^{" | cut -c1-330 | tee $O/r2g_s1r2_lookup.txt
timeout 200 python -m pytest tests/test_gpu_generic_sym.py -m gpu -q -x --timeout 300 2>&1 | tail -2
# C2: DRAM bytes + pipe activity of every launch of ONE force evaluation (the second: skip the initial one)
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active \
   --clock-control none -k 'regex:force_r3_f64|reduce_sym' -s 2 -c 2 --csv --log-file $O/r2g_ncu_pair_c2.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-parity --no-refcuda --no-e2e > $O/r2g_ncu_pair_c2.out 2>&1
grep -c force_r3 $O/r2g_ncu_pair_c2.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2g_ncu_launches_bench_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-parity --no-refcuda > $O/r2g_ncu_launches.out 2>&1
# --set full of the headline kernel at N = 400k (one launch = one evaluation)
SWEEP_SYM=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:force_r3_f64_sym -s 1 -c 1 -o $O/r2g_sym_n400k python tools/sweep_f64.py child 400000 > $O/r2g_ncu_full_sym.out 2>&1
ls -la $O/r2g*
