#!/bin/bash
# Round-1 session-4 GPU call 2: action-reaction path is now the default.  Full GPU tests, bench, ncu evidence, tuning sweep.
TAG=${1:-r1t}
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 ))s] $*" | tee -a $O/${TAG}_timeline.txt; }

stamp "pytest -m gpu (default = action-reaction path)"
timeout 900 python -m pytest tests -m gpu -x -q --timeout 400 > $O/${TAG}_gpu_tests.log 2>&1
echo "rc=$?" >> $O/${TAG}_gpu_tests.log
tail -6 $O/${TAG}_gpu_tests.log

stamp "smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log

stamp "bench"
timeout 420 python bench.py --steps 3 --warmup 3 > $O/${TAG}_bench_c2_1gpu.json 2> $O/${TAG}_bench_c2_1gpu.err
cat $O/${TAG}_bench_c2_1gpu.json | cut -c1-1800; tail -3 $O/${TAG}_bench_c2_1gpu.err

stamp "sweep N=400k: sym variants and chunk counts"
( for v in 0 1 2 3; do SWEEP_SYM=1 STEPS_B200_SYM_VARIANT=$v python tools/sweep_f64.py 400000 0; done
  for c in 28 112 224; do echo "chunks=$c"; STEPS_B200_SYM_CHUNKS=$c SWEEP_SYM=1 python tools/sweep_f64.py 400000 0; done ) > $O/${TAG}_sym_sweep_n400k.txt 2>&1
cut -c1-330 $O/${TAG}_sym_sweep_n400k.txt

stamp "sweep N=2M: sym variants 0 2 3 and chunk counts"
( for v in 0 2 3; do SWEEP_SYM=1 STEPS_B200_SYM_VARIANT=$v python tools/sweep_f64.py 2000000 0; done
  for c in 112 224; do echo "chunks=$c"; STEPS_B200_SYM_CHUNKS=$c SWEEP_SYM=1 python tools/sweep_f64.py 2000000 0; done ) > $O/${TAG}_sym_sweep_n2m.txt 2>&1
cut -c1-330 $O/${TAG}_sym_sweep_n2m.txt

stamp "ncu launch list of the bench command"
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/${TAG}_ncu_launches_bench_c2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $O/${TAG}_ncu_launches_bench.out 2>&1
tail -2 $O/${TAG}_ncu_launches_bench.out | cut -c1-300

stamp "ncu dram traffic + pipe counters of the pair kernel at C2 (one force evaluation = all passes)"
timeout 420 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed.sum \
    --clock-control none -k regex:"force_r3_f64|reduce_sym" -c 64 --csv --log-file $O/${TAG}_ncu_pair_c2.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu > $O/${TAG}_ncu_pair_c2.out 2>&1
tail -2 $O/${TAG}_ncu_pair_c2.out | cut -c1-300

stamp "ncu --set full, one launch of the sym kernel at N=400k"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:force_r3_f64_sym -s 1 -c 1 -o $O/${TAG}_sym_n400k \
    python bench.py --steps 1 --warmup 1 --n 400000 --no-cpu > $O/${TAG}_ncu_full.out 2>&1
tail -2 $O/${TAG}_ncu_full.out | cut -c1-300
ls -la $O/${TAG}_sym_n400k.ncu-rep
stamp "done"
