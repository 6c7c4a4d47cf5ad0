#!/bin/bash
# round 2, call e: T^3 kernel reading the z-window table with 128-bit loads: tests, throughput, ncu of the one-sided and the action-reaction kernel
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_generic_sym.py tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "not two_million" > $O/r2e_t3_tests.log 2>&1; echo rc=$? >> $O/r2e_t3_tests.log; tail -6 $O/r2e_t3_tests.log
for v in 5 3 4; do
  STEPS_B200_GEN_SYM_VARIANT=$v timeout 300 python tools/topo_bench.py t3:64,t3:48 2>&1 | grep "^{" | cut -c1-330
done | tee $O/r2e_t3_zwin_sweep.txt
STEPS_B200_GEN_SYM=0 timeout 240 ncu --set full --clock-control none --import-source on -k regex:force_generic_kernel -s 1 -c 1 -o $O/r2e_t3_onesided_48 python tools/topo_bench.py t3:48 > $O/r2e_ncu_t3_onesided.out 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k regex:force_generic_sym -s 1 -c 1 -o $O/r2e_t3_sym_48 python tools/topo_bench.py t3:48 > $O/r2e_ncu_t3_sym.out 2>&1
ls -la $O/*.ncu-rep
