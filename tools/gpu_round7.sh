#!/bin/bash
# Round-1 session-4 GPU call: verify the action-reaction path, then measure the mode that will ship.
# Everything lands in gpurun_out/${TAG}_*; every step has its own timeout.
TAG=${1:-r1s}
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 ))s] $*" | tee -a $O/${TAG}_timeline.txt; }

stamp "sym_debug"
timeout 240 python tools/sym_debug.py > $O/${TAG}_sym_debug.txt 2>&1
cat $O/${TAG}_sym_debug.txt | cut -c1-600

stamp "pytest test_gpu_sym (no -x)"
timeout 600 python -m pytest tests/test_gpu_sym.py -m gpu -q -s --timeout 240 > $O/${TAG}_gpu_tests_sym.log 2>&1
SYM_RC=$?
tail -25 $O/${TAG}_gpu_tests_sym.log

stamp "pytest -m gpu (default mode, without the sym file)"
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 --deselect tests/test_gpu_sym.py > $O/${TAG}_gpu_tests_default.log 2>&1
DEF_RC=$?
tail -5 $O/${TAG}_gpu_tests_default.log

stamp "pytest -m gpu with STEPS_B200_SYM=1 (R^3 FP64 whole-range calls take the action-reaction path)"
STEPS_B200_SYM=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "r3 or ragged or coincident or deterministic or momentum or resident or single_particle or full_size or zoom" > $O/${TAG}_gpu_tests_symenv.log 2>&1
SYMENV_RC=$?
tail -8 $O/${TAG}_gpu_tests_symenv.log

stamp "variant sweep N=400k"
( python tools/sweep_f64.py 400000 0
  for v in 0 1 2; do SWEEP_SYM=1 STEPS_B200_SYM_VARIANT=$v python tools/sweep_f64.py 400000 0; done ) > $O/${TAG}_variant_sweep_n400k.txt 2>&1
cat $O/${TAG}_variant_sweep_n400k.txt | cut -c1-400

MODE=0
if [ $SYM_RC -eq 0 ] && [ $SYMENV_RC -eq 0 ]; then MODE=1; fi
stamp "sym tests rc=$SYM_RC symenv rc=$SYMENV_RC default rc=$DEF_RC -> bench mode STEPS_B200_SYM=$MODE"

stamp "bench (mode $MODE)"
STEPS_B200_SYM=$MODE timeout 420 python bench.py --steps 3 --warmup 3 > $O/${TAG}_bench_c2_1gpu_sym${MODE}.json 2> $O/${TAG}_bench_c2_1gpu_sym${MODE}.err
cat $O/${TAG}_bench_c2_1gpu_sym${MODE}.json | cut -c1-1500; tail -3 $O/${TAG}_bench_c2_1gpu_sym${MODE}.err

stamp "ncu launch list of the bench command (mode $MODE)"
STEPS_B200_SYM=$MODE timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/${TAG}_ncu_launches_bench_c2_sym${MODE}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $O/${TAG}_ncu_launches_bench.out 2>&1
tail -2 $O/${TAG}_ncu_launches_bench.out | cut -c1-300

stamp "ncu dram traffic + pipe counters of the pair kernel at C2 (mode $MODE)"
STEPS_B200_SYM=$MODE timeout 420 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed.sum \
    --clock-control none -k regex:force_r3_f64 -c 24 --csv --log-file $O/${TAG}_ncu_pair_c2_sym${MODE}.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu > $O/${TAG}_ncu_pair_c2.out 2>&1
tail -2 $O/${TAG}_ncu_pair_c2.out | cut -c1-300

if [ $MODE -eq 1 ]; then
  stamp "ncu --set full, one pass of the sym kernel at N=400k"
  STEPS_B200_SYM=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:force_r3_f64_sym -s 1 -c 1 -o $O/${TAG}_sym_n400k \
      python bench.py --steps 1 --warmup 1 --n 400000 --no-cpu > $O/${TAG}_ncu_full.out 2>&1
  tail -2 $O/${TAG}_ncu_full.out | cut -c1-300
  stamp "one-sided bench for the record"
  STEPS_B200_SYM=0 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu > $O/${TAG}_bench_c2_1gpu_sym0.json 2> $O/${TAG}_bench_c2_1gpu_sym0.err
  cat $O/${TAG}_bench_c2_1gpu_sym0.json | cut -c1-600
fi
stamp "done"
