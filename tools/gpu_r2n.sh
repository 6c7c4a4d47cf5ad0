#!/bin/bash
# round 2, call n: launch planning with >= 128 waves per pass: tests, C2 on one GPU, per-rank pair-phase times of 8- and 4-GPU jobs played on one GPU
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
STEPS_B200_POISON=1 timeout 400 python -m pytest tests/test_gpu_sym.py tests/test_gpu_sym_f32.py tests/test_gpu_s1r2_sym.py tests/test_gpu_generic_sym.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-refcuda > $O/r2n_bench_c2.json 2> $O/r2n_bench_c2.err
python -c "import json; d=json.load(open('$O/r2n_bench_c2.json')); print('c2', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['launch_shape'], d['parity']['passed'], d['e2e']['value'])"
timeout 300 python tools/rankplay_time.py c2 8 0,3,7 2>&1 | grep "^{" | tee $O/r2n_rankplay_c2_p8.txt
timeout 300 python tools/rankplay_time.py c2 2 0 2>&1 | grep "^{" | tee -a $O/r2n_rankplay_c2_p8.txt
timeout 400 python tools/rankplay_time.py c5 8 0 2>&1 | grep "^{" | tee $O/r2n_rankplay_c5.txt
