#!/bin/bash
# round 2, call p: final regression (full pytest -m gpu with durations), default bench + reference arm, ncu --set full of the FP32 action-reaction kernel, smoke
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q --timeout 400 --durations=15 > $O/r2p_gpu_tests.log 2>&1; echo "rc=$?" >> $O/r2p_gpu_tests.log; grep -E "passed|failed|rc=|^[0-9.]+s " $O/r2p_gpu_tests.log | tail -22
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > $O/r2p_bench_c2_default.json 2> $O/r2p_bench_c2_default.err; echo "rc=$?"; python -c "import json; d=json.load(open('$O/r2p_bench_c2_default.json')); print('c2 default', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['parity']['passed'], d['reference_cuda'] and d['reference_cuda'].get('value'), d['cpu_baseline'] and d['cpu_baseline']['value'], d['roofline']['traffic'] and d['roofline']['traffic']['dram_bytes_per_evaluation'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2p_bench_ref_c2.json 2>/dev/null; cut -c1-200 $O/r2p_bench_ref_c2.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:force_r3_f32_sym -s 1 -c 1 -o $O/r2p_sym_f32_n400k python tools/ncu_f32_sym.py 400000 > $O/r2p_ncu_full_f32.out 2>&1; tail -1 $O/r2p_ncu_full_f32.out | cut -c1-200
