#!/bin/bash
# First GPU call of the next round: verify and measure what round 1 had to leave unverified (its GPU budget ran out):
#   1. the S^1xR^2 action-reaction kernel (pair_s1r2_sym.cuh, opt-in): tests, then every shape against the one-sided kernel
#   1b. the action-reaction kernel of the table-lookup topologies (pair_generic_sym.cuh, opt-in, never run): tests + sweep
#   1c. glass-making mode of the KDK step (never run): tests
#   1d. the reference's own CUDA kernels on the same GPU (second oracle + the kernel to beat)
#   2. ncu --set full of the FP32 action-reaction kernel (the round-1 capture was cut off mid-replay)
#   3. full regression + bench on the same box
TAG=${1:-r2a}
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 ))s] $*" | tee -a $O/${TAG}_timeline.txt; }
stamp "S^1xR^2 action-reaction kernel: tests"
STEPS_B200_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_s1r2_sym.py -m gpu -q -s --timeout 120 > $O/${TAG}_s1r2_sym_tests.log 2>&1
echo "rc=$?" >> $O/${TAG}_s1r2_sym_tests.log; grep -E "s1r2 sym|passed|failed|rc=|Error|assert" $O/${TAG}_s1r2_sym_tests.log | cut -c1-260 | tail -20
stamp "S^1xR^2 NOLOOKUP N=400k: one-sided, then the four action-reaction shapes"
timeout 120 python tools/topo_bench.py s1r2nl:400000 > $O/${TAG}_s1r2_sweep.txt 2>&1
for v in 0 1 2 3; do
  STEPS_B200_S1R2_SYM=1 STEPS_B200_S1R2_SYM_VARIANT=$v timeout 120 python tools/topo_bench.py s1r2nl:400000 >> $O/${TAG}_s1r2_sweep.txt 2>&1
done
cut -c1-300 $O/${TAG}_s1r2_sweep.txt
stamp "table-lookup topologies (T^3, S^1xR^2 lookup): action-reaction kernel tests"
STEPS_B200_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_generic_sym.py -m gpu -q -s --timeout 120 > $O/${TAG}_generic_sym_tests.log 2>&1
echo "rc=$?" >> $O/${TAG}_generic_sym_tests.log; grep -E " sym|passed|failed|rc=|Error|assert" $O/${TAG}_generic_sym_tests.log | cut -c1-260 | tail -20
stamp "same tests with the lean T^3 arithmetic (shape 3)"
STEPS_B200_EXPERIMENTAL=1 STEPS_B200_GEN_SYM_VARIANT=3 timeout 300 python -m pytest tests/test_gpu_generic_sym.py -m gpu -q -s --timeout 120 -k "t3 or multi_rank or softened or deterministic or kdk" > $O/${TAG}_generic_sym_tests_lean.log 2>&1
echo "rc=$?" >> $O/${TAG}_generic_sym_tests_lean.log; grep -E " sym|passed|failed|rc=|Error|assert" $O/${TAG}_generic_sym_tests_lean.log | cut -c1-260 | tail -20
stamp "T^3 64^3 and S^1xR^2 lookup N=200k: one-sided, then the action-reaction shapes"
timeout 200 python tools/topo_bench.py t3:64,s1r2:200000 > $O/${TAG}_generic_sweep.txt 2>&1
for v in 0 1 2 3 4 5; do
  STEPS_B200_GEN_SYM=1 STEPS_B200_GEN_SYM_VARIANT=$v timeout 200 python tools/topo_bench.py t3:64,s1r2:200000 >> $O/${TAG}_generic_sweep.txt 2>&1
done
cut -c1-300 $O/${TAG}_generic_sweep.txt
stamp "ncu --set full of the T^3 kernels at 48^3: one-sided (L1 wavefronts per load: the bound of DESIGN 3.4) and action-reaction (broadcast hypothesis)"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:force_generic_kernel -s 1 -c 1 -o $O/${TAG}_t3_onesided_48 \
    python tools/topo_bench.py t3:48 > $O/${TAG}_ncu_t3_onesided.out 2>&1
STEPS_B200_GEN_SYM=1 STEPS_B200_GEN_SYM_VARIANT=3 timeout 240 ncu --set full --clock-control none --import-source on -k regex:force_generic_sym -s 1 -c 1 -o $O/${TAG}_t3_sym_48 \
    python tools/topo_bench.py t3:48 > $O/${TAG}_ncu_t3_sym.out 2>&1
tail -1 $O/${TAG}_ncu_t3_onesided.out | cut -c1-200; tail -1 $O/${TAG}_ncu_t3_sym.out | cut -c1-200
stamp "the reference's own CUDA path (forces_cuda.cu for sm_100a): second oracle, then timed against ours on the same rows"
STEPS_B200_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_gpu_reference_cuda.py -m gpu -q -s --timeout 150 2>&1 | grep -E "ours vs|passed|failed|Error" | cut -c1-200 | tee $O/${TAG}_reference_cuda_tests.log
for cs in "c2 65536" "t3:64 32768" "s1r2nl:400000 32768"; do timeout 300 python tools/bench_ref_cuda.py $cs 2>&1 | tail -1 | cut -c1-600; done | tee $O/${TAG}_reference_cuda_bench.txt
stamp "BASELINE configs[0] (C1: N=32768, force evaluation + 10 KDK steps) against the reference's own run (golden fixture)"
STEPS_B200_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_c1_config.py -m gpu -q -s --timeout 100 2>&1 | grep -E "^C1|passed|failed|Error|assert" | cut -c1-260 | tee $O/${TAG}_c1_test.log
stamp "asynchronous ASCII snapshot of the resident engines"
STEPS_B200_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_gpu_snapshot.py -m gpu -q --timeout 100 2>&1 | tail -3 | tee $O/${TAG}_snapshot_test.log
stamp "glass-making mode (glass_kernels.cuh, never run): engine vs the CPU port, drop-in glass build vs the reference's"
STEPS_B200_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_gpu_glass.py -m gpu -q -s --timeout 120 > $O/${TAG}_glass_tests.log 2>&1
echo "rc=$?" >> $O/${TAG}_glass_tests.log; grep -E "^glass|passed|failed|rc=|Error|assert" $O/${TAG}_glass_tests.log | cut -c1-260 | tail -12
stamp "ncu --set full, one launch of the FP32 action-reaction kernel at N=400k"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:force_r3_f32_sym -s 1 -c 1 -o $O/${TAG}_sym_f32_n400k \
    python tools/ncu_f32_sym.py 400000 > $O/${TAG}_ncu_full_f32.out 2>&1
tail -2 $O/${TAG}_ncu_full_f32.out | cut -c1-300
stamp "full regression"
timeout 600 python -m pytest tests -m gpu -q --timeout 300 > $O/${TAG}_gpu_tests.log 2>&1
echo "rc=$?" >> $O/${TAG}_gpu_tests.log; tail -4 $O/${TAG}_gpu_tests.log
stamp "bench"
timeout 300 python bench.py --steps 3 --warmup 3 > $O/${TAG}_bench_c2_1gpu.json 2> $O/${TAG}_bench_c2_1gpu.err
cut -c1-260 $O/${TAG}_bench_c2_1gpu.json; tail -2 $O/${TAG}_bench_c2_1gpu.err
stamp "done"
