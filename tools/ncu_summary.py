#!/usr/bin/env python
"""development aid: the metrics of an `ncu --set full` report that the profiles/ summaries quote, one per line.
usage: ncu_summary.py report.ncu-rep [extra_metric_regex]  > profiles/<name>_summary.txt   (runs `ncu -i ... --page raw --csv` here)"""
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(dram__bytes_(read|write)\.sum$|gpu__time_duration\.sum$|launch__(registers_per_thread|waves_per_multiprocessor|occupancy_limit.*)$|"
    r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$|l1tex__data_pipe_lsu_wavefronts(_mem_shared)?\.sum$|l1tex__t_sector_hit_rate\.pct$|"
    r"l1tex__t_(requests|sectors)_pipe_lsu_mem_global_op_ld\.sum$|lts__t_sector_hit_rate\.pct$|lts__t_bytes\.sum$|"
    r"sm__cycles_elapsed\.avg(\.per_second)?$|sm__inst_executed_pipe_(alu|fma|fmaheavy|fp64|lsu|xu)\.avg\.pct_of_peak_sustained_active$|"
    r"sm__pipe_(fp64|fma)_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)$|sm__throughput\.avg\.pct_of_peak_sustained_elapsed$|"
    r"sm__warps_active\.avg\.pct_of_peak_sustained_active$|smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio$|"
    r"smsp__inst_executed\.sum$|smsp__issue_active\.avg\.pct_of_peak_sustained_active$)")


def main():
    rep = sys.argv[1]
    extra = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(head, r))
        print(f"{'Kernel Name':<120}{d.get('Kernel Name', '')}")
        print(f"{'Block Size':<120}{d.get('Block Size', '')}")
        print(f"{'Grid Size':<120}{d.get('Grid Size', '')}")
        for k, u in sorted(zip(head, units)):
            if KEEP.match(k) or (extra and extra.search(k)):
                print(f"{k:<101}{u:<19}{d[k]}")
        print()


if __name__ == "__main__":
    main()
