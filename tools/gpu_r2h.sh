#!/bin/bash
# round 2, call h: T^3 action-reaction kernel with software prefetch of the next pair's table rows; S^1xR^2 lookup window
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
for v in 6 7 8 5; do
  STEPS_B200_GEN_SYM_VARIANT=$v timeout 300 python tools/topo_bench.py t3:64,t3:48 2>&1 | grep "^{" | cut -c1-330
done | tee $O/r2h_t3_prefetch_sweep.txt
timeout 300 python tools/topo_bench.py s1r2:200000 2>&1 | grep "^{" | cut -c1-330 | tee $O/r2h_s1r2_lookup.txt
for v in 6 7; do STEPS_B200_GEN_SYM_VARIANT=$v timeout 200 python -m pytest tests/test_gpu_generic_sym.py -m gpu -q -x --timeout 300 2>&1 | tail -2; done | tee $O/r2h_tests.txt
STEPS_B200_GEN_SYM_VARIANT=6 timeout 300 ncu --set full --clock-control none --import-source on -k regex:force_generic_sym -s 1 -c 1 -o $O/r2h_t3_sym_prefetch_48 python tools/topo_bench.py t3:48 > $O/r2h_ncu_t3_sym.out 2>&1
