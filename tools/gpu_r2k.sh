#!/bin/bash
# round 2, call k: small-n multi-rank debugging on one GPU; C2 DRAM traffic with the L2 keep hint on the i-side round trip
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 120 python tools/debug_rankplay.py 6001 2 2>&1 | grep -v "^Calc\|Mmin\|done\|^$" | tee $O/r2k_rankplay_6001.txt
timeout 120 python tools/debug_rankplay.py 6001 2 f32 2>&1 | grep -v "^Calc\|Mmin\|done\|^$" | tail -4
timeout 300 python -m pytest tests/test_gpu_sym.py -m gpu -q -x --timeout 300 2>&1 | tail -2
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active \
   --clock-control none -k 'regex:force_r3_f64|reduce_sym' -s 2 -c 2 --csv --log-file $O/r2k_ncu_pair_c2_keep.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-parity --no-refcuda --no-e2e > $O/r2k_ncu_pair_c2.out 2>&1
python tools/ncu_traffic.py $O/r2k_ncu_pair_c2_keep.csv "C2, L2 evict_last hint on the i-side round trip" "r2k" | grep -E "dram_bytes_per_launch|pair_kernel_dram|pair_kernel_ms"
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-refcuda --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2 keep-hint', d['value'], d['ms_per_step'], d['roofline']['frac'])"
