#!/bin/bash
# Round-1 session-5 GPU call 2: full regression with the FP32 action-reaction path on by default.
TAG=${1:-r1ab}
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 ))s] $*" | tee -a $O/${TAG}_timeline.txt; }
stamp "full regression"
timeout 150 python -m pytest tests -m gpu -q --timeout 100 > $O/${TAG}_gpu_tests.log 2>&1
echo "rc=$?" >> $O/${TAG}_gpu_tests.log; tail -5 $O/${TAG}_gpu_tests.log
stamp "FP32 statistics"
timeout 40 python -m pytest tests/test_gpu_parity.py::test_r3_f32_vs_oracle tests/test_gpu_sym_f32.py::test_sym_f32_zoom_geometry_vs_truth_reference_and_one_sided -m gpu -q -s --timeout 30 2>&1 | grep -E "fp32|passed|failed" | cut -c1-400 | tee $O/${TAG}_f32_stats.txt
stamp "done"
