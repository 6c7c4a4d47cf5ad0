#!/bin/bash
# round 2, call c: superblock / window restructure of the action-reaction kernels: tests, then C2 throughput for several superblock sizes
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
tools/_bin/ubench_loads > $O/r2c_ubench_loads.txt 2>&1; cat $O/r2c_ubench_loads.txt
timeout 600 python -m pytest tests/test_gpu_sym.py tests/test_gpu_sym_f32.py tests/test_gpu_s1r2_sym.py tests/test_gpu_generic_sym.py -m gpu -q -x --timeout 300 > $O/r2c_sym_tests.log 2>&1; echo rc=$? >> $O/r2c_sym_tests.log; tail -15 $O/r2c_sym_tests.log
for sb in 8 1 4 16; do
  STEPS_B200_SYM_SB=$sb timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu > $O/r2c_bench_c2_sb$sb.json 2> $O/r2c_bench_c2_sb$sb.err
  python - <<PY
import json
try:
    d=json.load(open("$O/r2c_bench_c2_sb$sb.json")); print("sb=$sb", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["launch_shape"])
except Exception as ex: print("sb=$sb failed", ex)
PY
done
