#!/bin/bash
# Round-1 session-4 GPU call 3: re-verify the action-reaction kernel after the staging/unroll changes, sweep its shapes.
TAG=${1:-r1u}
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 ))s] $*" | tee -a $O/${TAG}_timeline.txt; }
stamp "pytest sym tests for every candidate shape"
for v in 1 2 3; do
  STEPS_B200_SYM_VARIANT=$v timeout 300 python -m pytest tests/test_gpu_sym.py -m gpu -x -q --timeout 200 2>&1 | tail -2 | sed "s/^/variant $v: /" >> $O/${TAG}_gpu_tests_sym_variants.log
done
cat $O/${TAG}_gpu_tests_sym_variants.log
stamp "pytest -m gpu (default)"
timeout 900 python -m pytest tests -m gpu -x -q --timeout 400 > $O/${TAG}_gpu_tests.log 2>&1
echo "rc=$?" >> $O/${TAG}_gpu_tests.log
tail -4 $O/${TAG}_gpu_tests.log
stamp "sweep N=400k"
( for v in 0 1 2 3; do SWEEP_SYM=1 STEPS_B200_SYM_VARIANT=$v python tools/sweep_f64.py 400000 0; done ) > $O/${TAG}_sym_sweep_n400k.txt 2>&1
cut -c1-200 $O/${TAG}_sym_sweep_n400k.txt
stamp "sweep N=2M"
( for v in 0 1 2 3; do SWEEP_SYM=1 STEPS_B200_SYM_VARIANT=$v python tools/sweep_f64.py 2000000 0; done ) > $O/${TAG}_sym_sweep_n2m.txt 2>&1
cut -c1-200 $O/${TAG}_sym_sweep_n2m.txt
BEST=$(python - <<PY
import json
best=None
for l in open("$O/${TAG}_sym_sweep_n2m.txt"):
    if l.startswith("{"):
        d=json.loads(l)
        if best is None or d["ms"]<best[0]: best=(d["ms"], d["sym_variant"])
print(best[1] if best else 0)
PY
)
stamp "best shape at N=2M: $BEST -> bench"
STEPS_B200_SYM_VARIANT=$BEST timeout 420 python bench.py --steps 3 --warmup 3 > $O/${TAG}_bench_c2_1gpu_v${BEST}.json 2> $O/${TAG}_bench_c2_1gpu.err
cut -c1-700 $O/${TAG}_bench_c2_1gpu_v${BEST}.json; tail -2 $O/${TAG}_bench_c2_1gpu.err
stamp "done"
