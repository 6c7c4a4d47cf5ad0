#!/usr/bin/env python
"""development aid: DRAM traffic and pipe activity of one force evaluation (pair kernel passes + row reductions) from an ncu CSV.
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed,\
smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k 'regex:force_r3_f64|reduce_sym' -s <skip> -c <launches of ONE evaluation> \
        --csv --log-file x.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-parity --no-refcuda --no-e2e
    python tools/ncu_traffic.py x.csv "<config text>" "<source text>" > profiles/pair_kernel_traffic.json   (bench.py reads it for roofline.traffic)"""
import csv
import json
import sys


def main():
    path, config, source = sys.argv[1], sys.argv[2], sys.argv[3]
    lines = [ln for ln in open(path) if ln.startswith('"')]
    rows = list(csv.DictReader(lines))
    per = {}
    order = []
    for r in rows:
        key = r["ID"]
        if key not in per:
            per[key] = {"kernel": r["Kernel Name"]}
            order.append(key)
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3, "second": 1e3, "msecond": 1.0,
                 "usecond": 1e-3, "nsecond": 1e-6}.get(unit, 1.0)
        per[key][r["Metric Name"]] = v * scale
    out = {"config": config, "source": source, "launches": []}
    tot = {"pair_r": 0.0, "pair_w": 0.0, "red_r": 0.0, "red_w": 0.0, "pair_ms": 0.0, "red_ms": 0.0, "fp64w": 0.0}
    for k in order:
        d = per[k]
        name = d["kernel"]
        is_pair = "force_" in name
        rd, wr, ms = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0), d.get("gpu__time_duration.sum", 0.0)
        fp = d.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0)
        out["launches"].append({"kernel": name[:60], "ms": ms, "dram_read_bytes": rd, "dram_write_bytes": wr, "fp64_pipe_active_pct": fp,
                                "issue_active_pct": d.get("smsp__issue_active.avg.pct_of_peak_sustained_active")})
        if is_pair:
            tot["pair_r"] += rd; tot["pair_w"] += wr; tot["pair_ms"] += ms; tot["fp64w"] += fp * ms
        else:
            tot["red_r"] += rd; tot["red_w"] += wr; tot["red_ms"] += ms
    out.update({
        "passes": sum(1 for k in order if "force_" in per[k]["kernel"]),
        "dram_bytes_per_launch": tot["pair_r"] + tot["pair_w"] + tot["red_r"] + tot["red_w"],
        "pair_kernel_dram_read_bytes": tot["pair_r"], "pair_kernel_dram_write_bytes": tot["pair_w"],
        "reduce_sym_dram_read_bytes": tot["red_r"], "reduce_sym_dram_write_bytes": tot["red_w"],
        "pair_kernel_ms_under_ncu": tot["pair_ms"], "reduce_sym_ms_under_ncu": tot["red_ms"],
        "fp64_pipe_active_pct_time_weighted": tot["fp64w"] / tot["pair_ms"] if tot["pair_ms"] else None,
        "note": "'launch' = one force evaluation = every pass of the pair kernel + its row reductions",
    })
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
