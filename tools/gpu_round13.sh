#!/bin/bash
# Round-1 session-5 GPU call: FP32 action-reaction kernel (tests + shape sweep at N=2M against the one-sided kernel), full regression.
TAG=${1:-r1aa}
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 ))s] $*" | tee -a $O/${TAG}_timeline.txt; }
stamp "FP32 action-reaction tests"
timeout 90 python -m pytest tests/test_gpu_sym_f32.py -m gpu -q -s --timeout 60 > $O/${TAG}_sym_f32_tests.log 2>&1
echo "rc=$?" >> $O/${TAG}_sym_f32_tests.log; grep -E "fp32|passed|failed|rc=|Error|assert" $O/${TAG}_sym_f32_tests.log | cut -c1-260 | tail -14
stamp "FP32 sweep at N=2M"
timeout 75 python tools/sweep_f32_sym.py 2000000 one,0,1,2 > $O/${TAG}_sym_f32_sweep_n2m.txt 2>&1
cut -c1-330 $O/${TAG}_sym_f32_sweep_n2m.txt
stamp "full regression (other files)"
timeout 120 python -m pytest tests -m gpu -q --timeout 100 --deselect tests/test_gpu_sym_f32.py > $O/${TAG}_gpu_tests.log 2>&1
echo "rc=$?" >> $O/${TAG}_gpu_tests.log; tail -4 $O/${TAG}_gpu_tests.log
stamp "done"
