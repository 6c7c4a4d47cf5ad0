#!/bin/bash
# round 2, call i: S^1xR^2 lookup after the stencil fix; BASELINE configs[3] (C4) at full size
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 300 python tools/topo_bench.py s1r2:200000 2>&1 | grep "^{" | cut -c1-330 | tee $O/r2i_s1r2_lookup.txt
timeout 200 python -m pytest tests/test_gpu_generic_sym.py tests/test_spatial_order.py -m gpu -q -x --timeout 300 2>&1 | tail -2
timeout 1200 python bench.py --config c4 --steps 1 --warmup 3 > $O/r2i_bench_c4_full.json 2> $O/r2i_bench_c4_full.err; echo "rc=$?" >> $O/r2i_bench_c4_full.err
tail -4 $O/r2i_bench_c4_full.err | cut -c1-300; cut -c1-400 $O/r2i_bench_c4_full.json
