#!/bin/bash
# GPU script for the start of the next round: verify and measure what round 1 had to leave unverified (its GPU budget ran out).
# usage: bash tools/gpu_next_round.sh <tag> [part ...]      parts (default: all, in this order):
#   verify   every gated test file (STEPS_B200_EXPERIMENTAL=1): S^1xR^2 / table-lookup action-reaction kernels (exact + lean), glass-making
#            mode, asynchronous snapshot, C1 golden run, the reference's own CUDA kernels as a second oracle            (~2 min)
#   measure  sweeps: S^1xR^2 and T^3 / S^1xR^2-lookup action-reaction shapes against the one-sided kernels; the reference's CUDA
#            kernels timed against ours on the same rows                                                                   (~5 min)
#   profile  ncu --set full: T^3 one-sided and action-reaction kernels (L1 wavefronts per load), FP32 action-reaction kernel  (~4 min)
#   final    full regression + bench on the same box                                                                       (~2 min)
#   c3full, c4full (only when named) pair-kernel throughput at the full C3 / C4 sizes, one-sided then action-reaction    (~15 / ~8 min)
#   c5full   (only when named) one KDK step at the full C5 size, N = 16.7M FP32                                            (~8 min)
TAG=${1:-r2a}
shift
PARTS=${*:-verify measure profile final}
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 ))s] $*" | tee -a $O/${TAG}_timeline.txt; }
gated() {  # gated <logname> <grep pattern> <pytest args...>
  local log=$O/${TAG}_$1.log pat=$2; shift 2
  STEPS_B200_EXPERIMENTAL=1 timeout 300 python -m pytest "$@" -m gpu -q -s --timeout 150 > $log 2>&1
  echo "rc=$?" >> $log; grep -E "$pat|passed|failed|rc=|Error|assert" $log | cut -c1-260 | tail -16
}

for part in $PARTS; do case $part in
verify)
  stamp "verify: S^1xR^2 action-reaction kernel"
  gated s1r2_sym_tests "s1r2 sym" tests/test_gpu_s1r2_sym.py
  stamp "verify: table-lookup topologies (T^3, S^1xR^2 lookup), exact arithmetic"
  gated generic_sym_tests " sym" tests/test_gpu_generic_sym.py
  stamp "verify: same with the lean T^3 arithmetic (shape 3)"
  STEPS_B200_GEN_SYM_VARIANT=3 gated generic_sym_tests_lean " sym" tests/test_gpu_generic_sym.py -k "t3 or multi_rank or softened or deterministic or kdk"
  stamp "verify: glass-making mode (engine vs the CPU port, drop-in glass build vs the reference's)"
  gated glass_tests "^glass" tests/test_gpu_glass.py
  stamp "verify: asynchronous ASCII snapshot"
  gated snapshot_test "snapshot" tests/test_gpu_snapshot.py
  stamp "verify: BASELINE configs[0] (C1) against the reference's own run"
  gated c1_test "^C1" tests/test_c1_config.py
  stamp "verify: group in spatial order returns the caller's order"
  gated spatial_order_test "spatial order" tests/test_spatial_order.py
  stamp "verify: the drop-in executable StePS_b200 against the reference executable (same parameter file and ASCII IC)"
  gated whole_binary_test "whole binary" tests/test_whole_binary.py
  stamp "verify: ours against the reference's own CUDA kernels (forces_cuda.cu for sm_100a)"
  gated reference_cuda_tests "ours vs" tests/test_gpu_reference_cuda.py
  ;;
measure)
  stamp "measure: S^1xR^2 NOLOOKUP N=400k: one-sided, then the four action-reaction shapes"
  timeout 120 python tools/topo_bench.py s1r2nl:400000 > $O/${TAG}_s1r2_sweep.txt 2>&1
  for v in 0 1 2 3; do
    STEPS_B200_S1R2_SYM=1 STEPS_B200_S1R2_SYM_VARIANT=$v timeout 120 python tools/topo_bench.py s1r2nl:400000 >> $O/${TAG}_s1r2_sweep.txt 2>&1
  done
  cut -c1-300 $O/${TAG}_s1r2_sweep.txt
  stamp "measure: T^3 64^3 and S^1xR^2 lookup N=200k: one-sided, then the action-reaction shapes (3-5 = lean T^3 arithmetic)"
  timeout 200 python tools/topo_bench.py t3:64,s1r2:200000 > $O/${TAG}_generic_sweep.txt 2>&1
  for v in 0 1 2 3 4 5; do
    STEPS_B200_GEN_SYM=1 STEPS_B200_GEN_SYM_VARIANT=$v timeout 200 python tools/topo_bench.py t3:64,s1r2:200000 >> $O/${TAG}_generic_sweep.txt 2>&1
  done
  cut -c1-300 $O/${TAG}_generic_sweep.txt
  stamp "measure: T^3 48^3 in lattice / random / cell-sorted particle order (tools/t3_wavefront_model.py predicts 1 : 1/6 : 1), one-sided then action-reaction"
  timeout 300 python tools/topo_bench.py t3:48,t3:48:random,t3:48:cellsort > $O/${TAG}_t3_order.txt 2>&1
  STEPS_B200_GEN_SYM=1 STEPS_B200_GEN_SYM_VARIANT=3 timeout 300 python tools/topo_bench.py t3:48,t3:48:random,t3:48:cellsort >> $O/${TAG}_t3_order.txt 2>&1
  cut -c1-260 $O/${TAG}_t3_order.txt
  stamp "measure: C2 with a 48 GB j-side row buffer (3 passes instead of 8; tools/pass_model.py: 0.988 -> 0.995 of ideal)"
  STEPS_B200_SYM_GPART_MB=49152 timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu > $O/${TAG}_bench_c2_gpart48g.json 2> $O/${TAG}_bench_c2_gpart48g.err
  cut -c1-200 $O/${TAG}_bench_c2_gpart48g.json
  stamp "measure: the reference's own CUDA path timed against ours on the same rows"
  for cs in "c2 65536" "t3:64 32768" "s1r2nl:400000 32768"; do timeout 300 python tools/bench_ref_cuda.py $cs 2>&1 | tail -1 | cut -c1-600; done | tee $O/${TAG}_reference_cuda_bench.txt
  ;;
profile)
  stamp "profile: ncu --set full of the T^3 kernels at 48^3 (one-sided: L1 wavefronts per load, DESIGN 3.4; action-reaction: broadcast hypothesis)"
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:force_generic_kernel -s 1 -c 1 -o $O/${TAG}_t3_onesided_48 \
      python tools/topo_bench.py t3:48 > $O/${TAG}_ncu_t3_onesided.out 2>&1
  STEPS_B200_GEN_SYM=1 STEPS_B200_GEN_SYM_VARIANT=3 timeout 240 ncu --set full --clock-control none --import-source on -k regex:force_generic_sym -s 1 -c 1 -o $O/${TAG}_t3_sym_48 \
      python tools/topo_bench.py t3:48 > $O/${TAG}_ncu_t3_sym.out 2>&1
  tail -1 $O/${TAG}_ncu_t3_onesided.out | cut -c1-200; tail -1 $O/${TAG}_ncu_t3_sym.out | cut -c1-200
  stamp "profile: ncu --set full, one launch of the FP32 action-reaction kernel at N=400k"
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:force_r3_f32_sym -s 1 -c 1 -o $O/${TAG}_sym_f32_n400k \
      python tools/ncu_f32_sym.py 400000 > $O/${TAG}_ncu_full_f32.out 2>&1
  tail -2 $O/${TAG}_ncu_full_f32.out | cut -c1-300
  ;;
final)
  stamp "final: full regression"
  timeout 600 python -m pytest tests -m gpu -q --timeout 300 > $O/${TAG}_gpu_tests.log 2>&1
  echo "rc=$?" >> $O/${TAG}_gpu_tests.log; tail -4 $O/${TAG}_gpu_tests.log
  stamp "final: bench"
  timeout 300 python bench.py --steps 3 --warmup 3 > $O/${TAG}_bench_c2_1gpu.json 2> $O/${TAG}_bench_c2_1gpu.err
  cut -c1-260 $O/${TAG}_bench_c2_1gpu.json; tail -2 $O/${TAG}_bench_c2_1gpu.err
  ;;
c5full)
  # not in the default list: BASELINE configs[4] at its real size (N = 16 777 216 FP32, 2.8e14 interactions, about 100 s per evaluation)
  stamp "c5full: one timed KDK step at the full C5 size (warm-up 1: a development number, not a bench line)"
  timeout 1500 python bench.py --config c5 --steps 1 --warmup 1 --no-cpu > $O/${TAG}_bench_c5_full.json 2> $O/${TAG}_bench_c5_full.err
  cut -c1-300 $O/${TAG}_bench_c5_full.json; tail -3 $O/${TAG}_bench_c5_full.err
  ;;
c4full)
  # not in the default list: BASELINE configs[3] at its real size (S^1xR^2 NOLOOKUP, N = 4 194 304), one-sided then action-reaction
  stamp "c4full: pair-kernel throughput at the full C4 size"
  timeout 900 python tools/topo_bench.py s1r2nl:4194304 2>&1 | tail -1 | cut -c1-400 | tee $O/${TAG}_c4_full.txt
  STEPS_B200_S1R2_SYM=1 timeout 900 python tools/topo_bench.py s1r2nl:4194304 2>&1 | tail -1 | cut -c1-400 | tee -a $O/${TAG}_c4_full.txt
  ;;
c3full)
  # not in the default list: BASELINE configs[2] at its real size (T^3, 128^3), one-sided then action-reaction (lean arithmetic)
  stamp "c3full: pair-kernel throughput at the full C3 size"
  timeout 1500 python tools/topo_bench.py t3:128 2>&1 | tail -1 | cut -c1-400 | tee $O/${TAG}_c3_full.txt
  STEPS_B200_GEN_SYM=1 STEPS_B200_GEN_SYM_VARIANT=3 timeout 1500 python tools/topo_bench.py t3:128 2>&1 | tail -1 | cut -c1-400 | tee -a $O/${TAG}_c3_full.txt
  ;;
*) echo "unknown part $part";;
esac; done
stamp "done"
