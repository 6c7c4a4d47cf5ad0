#!/bin/bash
# short multi-GPU check: device-restore test + N-GPU bench with phase timestamps; everything under its own timeout
NG=${1:-2}; TAG=${2:-r1o}
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "restore" 2>&1 | tail -3
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_g$NG.json 2> gpurun_out/${TAG}_bench_g$NG.err
cat gpurun_out/${TAG}_bench_g$NG.json; grep "bench\]\|Error\|error" gpurun_out/${TAG}_bench_g$NG.err | tail -20
