#!/bin/bash
# round 2, call f: T^3 kernel reading aligned row copies with 128-bit loads: tests, throughput, ncu (both kernels at 32^3)
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_generic_sym.py tests/test_gpu_parity.py tests/test_abi.py -m gpu -q -x --timeout 300 -k "not two_million" > $O/r2f_t3_tests.log 2>&1; echo rc=$? >> $O/r2f_t3_tests.log; tail -6 $O/r2f_t3_tests.log
for v in 5 3; do
  STEPS_B200_GEN_SYM_VARIANT=$v timeout 300 python tools/topo_bench.py t3:64,t3:48 2>&1 | grep "^{" | cut -c1-330
done | tee $O/r2f_t3_aligned_sweep.txt
STEPS_B200_GEN_SYM=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:force_generic_kernel -s 1 -c 1 -o $O/r2f_t3_onesided_32 python tools/topo_bench.py t3:32 > $O/r2f_ncu_t3_onesided.out 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:force_generic_sym -s 1 -c 1 -o $O/r2f_t3_sym_32 python tools/topo_bench.py t3:32 > $O/r2f_ncu_t3_sym.out 2>&1
ls -la $O/r2f*.ncu-rep; tail -3 $O/r2f_ncu_t3_onesided.out | cut -c1-200
