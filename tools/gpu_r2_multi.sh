#!/bin/bash
# round 2: N-GPU box: multi-GPU tests, C2 at N GPUs (bench line with parity), C5 at full size at N GPUs, the stateless multi-GPU call
# usage: gpu_r2_multi.sh <N> [parts]   parts: tests c2 c5 multi (default: all)
N=$1; shift; PARTS=${*:-tests c2 c5 multi}
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1].split('/')[-1], "value %.4g" % d["value"], "ms %.1f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "e2e", d["e2e"] and "%.4g" % d["e2e"]["value"],
          "parity", d["parity"] and (d["parity"].get("passed"), d["parity"].get("max_dF_over_sum_abs_fij")), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as ex:
    print(sys.argv[1], "failed:", ex)
PY
}
for part in $PARTS; do case $part in
tests)
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 > $O/r2m${N}_multi_tests.log 2>&1; echo "rc=$?" >> $O/r2m${N}_multi_tests.log; tail -4 $O/r2m${N}_multi_tests.log ;;
c2)
  timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 > $O/r2m${N}_bench_c2.json 2> $O/r2m${N}_bench_c2.err; echo "rc=$?" >> $O/r2m${N}_bench_c2.err
  show $O/r2m${N}_bench_c2.json; tail -2 $O/r2m${N}_bench_c2.err | cut -c1-300 ;;
c5)
  timeout 2400 $TR bench.py --gpus $N --config c5 --steps 1 --warmup 3 --no-e2e > $O/r2m${N}_bench_c5_full.json 2> $O/r2m${N}_bench_c5_full.err; echo "rc=$?" >> $O/r2m${N}_bench_c5_full.err
  show $O/r2m${N}_bench_c5_full.json; tail -2 $O/r2m${N}_bench_c5_full.err | cut -c1-300 ;;
multi)
  timeout 600 python tools/bench_forces_multi.py $N 2>&1 | tail -1 | tee $O/r2m${N}_forces_multi.txt
  STEPS_B200_MULTI_ONESIDED=1 timeout 600 python tools/bench_forces_multi.py $N 2>&1 | tail -1 | tee -a $O/r2m${N}_forces_multi.txt ;;
esac; done
