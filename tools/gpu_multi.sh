#!/bin/bash
# multi-GPU round: NG = number of GPUs of the box
NG=${1:-2}; TAG=${2:-r1m}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5
for n in 1 $NG; do
  if [ $n = 1 ]; then
    python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_g1.json 2> gpurun_out/${TAG}_bench_g1.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_g$n.json 2> gpurun_out/${TAG}_bench_g$n.err
  fi
  cat gpurun_out/${TAG}_bench_g$n.json; tail -3 gpurun_out/${TAG}_bench_g$n.err
done
