#!/bin/bash
# Round-1 session-4 GPU call: tuned T^3 kernel (opt-in) against the exact-branch kernel, plus a full regression after the shape-table trim.
TAG=${1:-r1y}
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 ))s] $*" | tee -a $O/${TAG}_timeline.txt; }
stamp "T^3 tests, exact-branch kernel (default)"
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -s --timeout 300 -k "t3" > $O/${TAG}_t3_tests_generic.log 2>&1
echo "rc=$?" >> $O/${TAG}_t3_tests_generic.log; grep -E "t3|passed|failed|rc=" $O/${TAG}_t3_tests_generic.log | cut -c1-200 | tail -12
for v in 0 1 2 3; do
  stamp "T^3 tests, tuned kernel shape $v"
  STEPS_B200_T3_TUNED=1 STEPS_B200_T3_VARIANT=$v timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -s --timeout 300 -k "t3" > $O/${TAG}_t3_tests_tuned_v$v.log 2>&1
  echo "rc=$?" >> $O/${TAG}_t3_tests_tuned_v$v.log; grep -E "t3|passed|failed|rc=|Error|error" $O/${TAG}_t3_tests_tuned_v$v.log | cut -c1-200 | tail -12
done
stamp "T^3 throughput: exact-branch vs tuned shapes (48^3 and 64^3)"
( echo "generic"; timeout 200 python tools/topo_bench.py t3:48 2>/dev/null | grep "^{"
  for v in 0 1 2 3; do echo "tuned shape $v"; STEPS_B200_T3_TUNED=1 STEPS_B200_T3_VARIANT=$v timeout 200 python tools/topo_bench.py t3:48,t3:64 2>/dev/null | grep "^{"; done ) > $O/${TAG}_t3_bench.txt 2>&1
cut -c1-260 $O/${TAG}_t3_bench.txt
stamp "full regression (default settings)"
timeout 900 python -m pytest tests -m gpu -x -q --timeout 400 > $O/${TAG}_gpu_tests.log 2>&1
echo "rc=$?" >> $O/${TAG}_gpu_tests.log; tail -4 $O/${TAG}_gpu_tests.log
stamp "bench (short)"
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu > $O/${TAG}_bench_c2_1gpu.json 2> $O/${TAG}_bench_c2_1gpu.err
cut -c1-260 $O/${TAG}_bench_c2_1gpu.json; tail -2 $O/${TAG}_bench_c2_1gpu.err
stamp "done"
