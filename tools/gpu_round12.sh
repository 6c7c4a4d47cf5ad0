#!/bin/bash
# Round-1 session-4 GPU call: GPU T^3 Ewald table builder against the reference builder, then the full regression + bench of the final tree.
TAG=${1:-r1z}
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 ))s] $*" | tee -a $O/${TAG}_timeline.txt; }
stamp "T^3 Ewald table builder tests"
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -s --timeout 300 -k "ewald_table or gpu_built_table or radial_force_table" > $O/${TAG}_table_tests.log 2>&1
echo "rc=$?" >> $O/${TAG}_table_tests.log; grep -E "T\^3|t3_|passed|failed|rc=|Error" $O/${TAG}_table_tests.log | cut -c1-220 | tail -12
stamp "table build time 63^3 / 127^3 / 255^3 on the GPU"
timeout 200 python - > $O/${TAG}_table_build_times.txt 2>&1 <<'PY'
import time, numpy as np
import steps_b200 as sb
from steps_b200 import _lib
lib = _lib.load()
for ip in (2, 3, 4, 2, 3, 4):
    d = sb.t3_ewald_defaults(ip, 100.0); n = d["ngrid"]
    tab = np.empty(n * n * n * 3)
    t0 = time.perf_counter()
    _lib.check(lib.steps_b200_t3_ewald_table_f64(n, 100.0, d["alpha"], d["rel_cut"], d["rec_cut"], tab.ctypes.data, 0))
    print(f"T^3 IS_PERIODIC={ip}: {n}^3 table in {time.perf_counter() - t0:.3f} s (call incl. alloc + D2H), checksum {np.abs(tab).sum():.12e}", flush=True)
for ip in (2, 3, 4):
    d = sb.s1r2_ewald_defaults(ip, 100.0, 500.0)
    tab = np.empty(d["nrho"] * d["nz"] * 2)
    t0 = time.perf_counter()
    _lib.check(lib.steps_b200_s1r2_ewald_table_f64(d["nrho"], d["nz"], d["rho_max"], 100.0, d["alpha"], d["nmax"], d["mmax"], tab.ctypes.data, 0))
    print(f"S1R2 IS_PERIODIC={ip}: {d['nrho']}x{d['nz']} table in {time.perf_counter() - t0:.3f} s, checksum {np.abs(tab).sum():.12e}", flush=True)
PY
cat $O/${TAG}_table_build_times.txt
stamp "full regression"
timeout 900 python -m pytest tests -m gpu -x -q --timeout 400 > $O/${TAG}_gpu_tests.log 2>&1
echo "rc=$?" >> $O/${TAG}_gpu_tests.log; tail -4 $O/${TAG}_gpu_tests.log
stamp "smoke + bench (short)"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu > $O/${TAG}_bench_c2_1gpu.json 2> $O/${TAG}_bench_c2_1gpu.err
cut -c1-260 $O/${TAG}_bench_c2_1gpu.json; tail -2 $O/${TAG}_bench_c2_1gpu.err
stamp "done"
