#!/bin/bash
# round 2, call d: occupancy shapes of the headline kernel; the extended bench.py (configs c3/c4/c5 at development sizes, parity block, reference CUDA leg)
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
for v in 0 4 5 6; do
  SWEEP_SYM=1 STEPS_B200_SYM_VARIANT=$v timeout 120 python tools/sweep_f64.py child 400000 2>&1 | tail -1 | cut -c1-300
done | tee $O/r2d_sym_shape_sweep_n400k.txt
run() { # run <tag> <args...>
  local tag=$1; shift
  timeout 600 python bench.py "$@" > $O/r2d_bench_$tag.json 2> $O/r2d_bench_$tag.err; echo "rc=$?" >> $O/r2d_bench_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("$O/r2d_bench_$tag.json"))
    print("$tag", "value %.4g" % d["value"], "ms %.1f" % d["ms_per_step"], "frac %.3f (%s)" % (d["roofline"]["frac"], d["roofline"]["bound"]), "e2e %.4g" % d["e2e"]["value"],
          "parity", d["parity"] and (d["parity"].get("passed"), d["parity"].get("max_dF_over_sum_abs_fij")), "refcuda", d["reference_cuda"] and d["reference_cuda"].get("value"), "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"])
except Exception as ex:
    print("$tag failed", ex); import subprocess; print(open("$O/r2d_bench_$tag.err").read()[-1500:])
PY
}
run c2 --steps 2 --warmup 3
run c3_n48 --config c3 --n 48 --steps 2 --warmup 3
run c4_n400k --config c4 --n 400000 --steps 2 --warmup 3
run c5_n2m --config c5 --n 2000000 --steps 2 --warmup 3
run c1 --config c1 --steps 5 --warmup 3
