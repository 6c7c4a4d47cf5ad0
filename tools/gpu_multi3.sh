#!/bin/bash
# 2-GPU call: the multi-GPU action-reaction path (ring assignment of block pairs + one all-reduce per evaluation)
TAG=${1:-r1v}
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 ))s] $*" | tee -a $O/${TAG}_timeline.txt; }
stamp "pytest test_gpu_multi"
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 300 > $O/${TAG}_gpu_multi_tests.log 2>&1
echo "rc=$?" >> $O/${TAG}_gpu_multi_tests.log
tail -5 $O/${TAG}_gpu_multi_tests.log
stamp "bench 2 GPUs"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > $O/${TAG}_bench_c2_2gpu.json 2> $O/${TAG}_bench_c2_2gpu.err
cut -c1-900 $O/${TAG}_bench_c2_2gpu.json; tail -4 $O/${TAG}_bench_c2_2gpu.err
stamp "done"
