#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json: pair interactions/s and s/step of the
direct-summation KDK step at N = 2,000,000 FP64 (configs[1], compactified R^3 zoom-in, mass-dependent
softening), at 1/2/4/8 B200, next to the reference's OpenMP CPU direct sum timed on this box's cores.

    python bench.py --gpus N --steps K --warmup W          # our arm (one process per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K --warmup W    # the reference's CPU path, rank 0 only

One "step" = one KDK step of the resident engine: kick(h/2)+drift(h) -> position all-gather (NCCL, N>1)
-> force evaluation of all N i-particles against all N j-particles -> kick(h/2)+errmax reduction, i.e.
N^2 pair evaluations (self pair included, as the reference evaluates it).  Rank 0 prints ONE JSON line.

What the keys mean here (see DESIGN.md "Measurement"):
  value     whole-job pairs/s with the particle state resident in HBM, CUDA events on the engine's stream
            around exactly K steps, max over ranks.
  e2e       the same metric through the reference-facing C-ABI call steps_b200_forces_f64() with HOST
            (pinned) buffers: H2D of x, M, s and D2H of F are inside the timed region, every step.
  roofline  the pair kernel against the FP64 FMA pipe (this path is FP64-pipe bound, not HBM or tensor):
            achieved = 20 flop x pair INTERACTIONS per launch (the reference's count, n_i x N) / CUDA-event duration of
            the pair phase (the action-reaction path delivers two interactions per evaluation, so its achieved
            figure can exceed what 15 instructions per directed pair allow),
            peak = DFMA microbenchmark measured live on the same GPU (MEASURED_PEAKS.json has no FP64 entry).
  cpu_baseline  oracle/_ref (the unmodified reference, kind "reference") or the plain-C port (kind "port")
            on a bounded i-subrange of the same workload, all host threads.
PyTorch is used only for plumbing: torch.distributed rendezvous/barrier and pinned host buffers.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOP_PER_PAIR = 20  # the reference's literal far-field arithmetic, sqrt/div counted as 1
METRIC = "pair_interactions_per_s"
UNIT = "pairs/s"


_T0 = time.perf_counter()


def log(*a):
    print(f"[{time.perf_counter() - _T0:7.1f}s]", *a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------ workload
def make_ic(args):
    import numpy as np

    from steps_b200 import ic

    if args.config == "c2":
        if args.n and args.n != 2_000_000:
            # development sizes: same geometry (122 tan-spaced shells holding 85.4 % of the particles), scaled counts
            c = ic.compactified_r3(args.n, 224, max(1, int(0.854 * args.n / 122)), 20242, name=f"C2-shaped compactified R^3 N={args.n}")
        else:
            c = ic.config_c2()
    elif args.config == "c1":
        c = ic.config_c1()
    elif args.config == "c5":
        if args.n and args.n != 16_777_216:
            # development sizes of the single-precision configuration: same geometry, scaled counts
            c = ic.compactified_r3(args.n, 224, max(1, int(0.854 * args.n / 122)), 20245, np.float32, name=f"C5-shaped compactified R^3 FP32 N={args.n}")
        else:
            c = ic.config_c5()
    else:
        raise SystemExit(f"unknown --config {args.config}")
    assert c.x.dtype == (np.float64 if args.config != "c5" else np.float32)
    return c


def workload_config(c, world, symmetric=False):
    g = c.g
    return {
        "evaluation": ("action-reaction: every unordered pair evaluated once and applied to both particles (N(N+1)/2 evaluations "
                       "deliver the N^2 interactions the reference evaluates one by one)") if symmetric else
                      "one-sided: N^2 directed pair evaluations, as the reference",
        "workload": f"{c.name}: one KDK step (kick+drift, position all-gather, N^2 direct-sum force, kick+errmax), "
                    "mass-dependent pairwise softening, comoving LCDM background term",
        "baseline_config": "configs[1]" if "C2" in c.name else c.name.split()[0],
        "n_particles": int(g.N),
        "pairs_per_step": int(g.N) * int(g.N),
        "topology": "R3",
        "flop_per_pair": FLOP_PER_PAIR,
        "parallelism": f"i-partition over {world} GPU(s), full j replica per GPU" + (", NCCL all-gather of positions per step" if world > 1 else ""),
        "l2": "no flush needed: the packed j-stream (64 B x N = 128 MB at N=2M) exceeds the 126 MB L2 and is re-streamed by every CTA wave",
    }


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """samples SM clock and throttle reasons of one GPU during the timed region (NVML; nvidia-smi fallback)"""

    def __init__(self, device_index: int, uuid: str | None = None, period: float = 0.1):
        self.samples, self.reasons, self.power = [], set(), []
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        self._period = period
        self._h = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByUUID(uuid) if uuid else pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception as ex:  # noqa: BLE001
            log(f"[bench] NVML unavailable ({ex}); clocks not sampled")
            self._h = None

    def _loop(self):
        nv = self._nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(self._period)

    def start(self):
        if self._h is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self) -> dict:
        self._stop.set()
        if self._thr:
            self._thr.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_min_mhz": s[0], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "power_w_max": max(self.power) if self.power else None}


# ------------------------------------------------------------------------------------------------ CPU legs
def cpu_forces_fn(c):
    """-> (kind, callable(lo, hi) -> F, cores): the reference itself if its build travelled, else the plain-C port"""
    from oracle import pyport, pyref

    g = c.g
    cores = os.cpu_count() or 1
    key = (g.topology, 8 if g.REAL.__name__ == "float64" else 4)
    variant = pyref.VARIANT.get(key)
    if variant and pyref.available(variant):
        r = pyref.Reference(variant)
        r.configure(g)
        return "reference", (lambda lo, hi: r.forces(c.x, lo, hi, cores)), cores
    pyport.load()
    return "port", (lambda lo, hi: pyport.forces(g, c.x, lo, hi, cores)), cores


def cpu_sample(c, target_s: float):
    """time the CPU direct sum on a bounded, contiguous i-subrange sized for ~target_s seconds"""
    g = c.g
    kind, fn, cores = cpu_forces_fn(c)
    lo = g.N // 2  # middle of the load (shell particles in the zoom geometry; every i costs N pairs anyway)
    n_i = min(64, g.N - lo)
    t0 = time.perf_counter()
    fn(lo, lo + n_i - 1)
    t = time.perf_counter() - t0
    rate = n_i * g.N / max(t, 1e-9)
    n_i = int(max(cores, min(g.N - lo, target_s * rate / g.N)))
    t0 = time.perf_counter()
    fn(lo, lo + n_i - 1)
    t = time.perf_counter() - t0
    return {"value": n_i * g.N / t, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"i in [{lo}, {lo + n_i - 1}] ({n_i} rows) x all N={g.N} j, {t:.2f} s wall, OpenMP {cores} threads; "
                      f"extrapolated full step = {g.N * g.N / (n_i * g.N / t):.0f} s"}, n_i, lo, fn


def run_reference(args, out_fd):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = make_ic(args)
    g = c.g
    total = max(1, args.steps + args.warmup)
    per_step = args.ref_seconds if args.ref_seconds > 0 else min(10.0, 150.0 / total)
    base, n_i, lo, fn = cpu_sample(c, per_step)
    for _ in range(args.warmup):
        fn(lo, lo + n_i - 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn(lo, lo + n_i - 1)
    t = time.perf_counter() - t0
    value = args.steps * n_i * g.N / t
    base["value"] = value
    base["sample"] = f"each step = i in [{lo}, {lo + n_i - 1}] ({n_i} rows) x all N={g.N} j, OpenMP {base['cores']} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * g.N * g.N / value, "ms_per_sample_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64" if g.REAL.__name__ == "float64" else "f32",
        "data": "synthetic", "config": workload_config(c, 1), "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "ms_per_step is the full N^2 step extrapolated from the sampled rows (every i costs exactly N pairs)",
    }
    os.write(out_fd, (json.dumps(line) + "\n").encode())


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args, out_fd):
    import numpy as np
    import torch
    import torch.distributed as dist

    import steps_b200 as sb
    from steps_b200 import ranks

    rank, world, local = ranks.env_rank()
    if world != args.gpus:
        log(f"[bench] WORLD_SIZE={world} but --gpus {args.gpus}: using WORLD_SIZE")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: steps_b200 has no CPU path")
    torch.cuda.set_device(local)
    os.environ["STEPS_B200_DEVICE"] = str(local)  # the stateless C-ABI calls (e2e leg) run on this rank's GPU
    dev = f"cuda:{local}"
    if world > 1:
        import datetime

        log(f"[bench] rank {rank}/{world}: init_process_group(nccl) ...")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
        log(f"[bench] rank {rank}: process group up")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        return ranks.reduce_scalar(dist, world, v, "max", dev)

    def sum_over_ranks(v: float) -> float:
        return ranks.reduce_scalar(dist, world, v, "sum", dev)

    c = make_ic(args)  # same seeds on every rank -> bit-identical arrays
    g = c.g
    N = g.N
    rb = 8 if g.REAL == np.float64 else 4
    eng = sb.Engine(g, local)
    if world > 1:
        eng.comm_init(ranks.share_unique_id(dist, rank, world, sb.Engine.nccl_unique_id), rank, world)
        log(f"[bench] rank {rank}: engine NCCL communicator up")
    eng.upload(c.x, c.v)
    eng.forces()
    h = eng.calculate_init_h()
    h = min(max(h, g.h_min), g.h_max)
    log(f"[bench] rank {rank}: initial forces done, h0 = {h:.4e}")

    # FP pipe peak, measured live on this GPU (burst = kernel alone; sustained = 2 s back to back)
    peak_burst, implied_mhz = sb.fma_peak(local, rb)
    peak_sust = sb.fma_peak_sustained(local, rb, 2.0)

    for _ in range(args.warmup):
        eng.step(h)
        h = eng.next_h()
    uuid = None
    try:
        uuid = "GPU-" + str(torch.cuda.get_device_properties(local).uuid)
    except Exception:  # noqa: BLE001
        pass
    sampler = ClockSampler(local, uuid)
    launches0 = eng.launch_count()
    pair_ms, force_ms = [], []
    barrier()
    sampler.start()
    eng.mark(0)
    for _ in range(args.steps):
        eng.step(h)  # returns after the errmax D2H of this step (the next h depends on it, as in main.cc:1834)
        h = eng.next_h()
        pair_ms.append(eng.pair_kernel_ms())
        force_ms.append(eng.timings()[0])
    eng.mark(1)
    ms_total = eng.elapsed_ms(0, 1)
    barrier()
    clocks = sampler.stop()
    launches = eng.launch_count() - launches0
    ms_total = max_over_ranks(ms_total)
    launches = int(sum_over_ranks(float(launches)))
    ms_step = ms_total / args.steps
    value = N * float(N) / (ms_step * 1e-3)

    # roofline of the dominant kernel (the pair kernel) on this rank
    n_i = eng.i_hi - eng.i_lo
    pk_ms = sum(pair_ms) / len(pair_ms)
    pairs_per_launch = float(n_i) * N
    achieved_tf = FLOP_PER_PAIR * pairs_per_launch / (pk_ms * 1e-3) / 1e12
    shape = eng.launch_shape(eng.i_lo, eng.i_hi - 1)
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "pair_kernel_traffic.json")
    if os.path.exists(tr_path) and world == 1 and args.config == "c2" and not args.n:
        try:
            traffic = json.load(open(tr_path)).get("dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            traffic = None
    symmetric = eng.symmetric
    roofline = {
        "bound": "fp64_pipe" if rb == 8 else "fp32_pipe",
        "kernel": ("force_r3_f64_sym_kernel" if symmetric else "force_r3_f64_kernel") if rb == 8 else
                  ("force_r3_f32_sym_kernel" if symmetric else "force_r3_f32_kernel"),
        "fp64_instr_per_interaction": (10 if symmetric else 15) if rb == 8 else None,
        "fp32_instr_per_interaction": None if rb == 8 else (9.5 if symmetric else 14),
        "achieved": achieved_tf, "peak": peak_sust, "unit": "TFLOP/s", "frac": achieved_tf / peak_sust,
        "peak_burst": peak_burst, "frac_of_burst": achieved_tf / peak_burst,
        "peak_source": "DFMA/FFMA microbenchmark (steps_b200_fma_peak_sustained: 2 s back to back; burst = best single launch) "
                       "measured live on this GPU; MEASURED_PEAKS.json has no FP64/FP32 CUDA-core entry",
        "flop_per_pair": FLOP_PER_PAIR, "pairs_per_launch": pairs_per_launch, "kernel_ms": pk_ms,
        "kernel_share_of_step": pk_ms / ms_step,
        "traffic": traffic, "algorithmic_bytes_per_launch": 64.0 * N * shape["ctas"] / max(1, shape["j_chunks"]) + 24.0 * n_i * shape["j_chunks"],
        "launch_shape": shape,
    }

    # e2e: the reference-facing stateless C-ABI call with host buffers, every step H2D(x,M,s) + D2H(F)
    lo, hi = eng.i_lo, eng.i_hi - 1
    xh = torch.from_numpy(c.x).pin_memory()
    mh = torch.from_numpy(np.ascontiguousarray(g.M)).pin_memory()
    sh = torch.from_numpy(np.ascontiguousarray(g.SOFT_LENGTH)).pin_memory()
    Fh = torch.empty(3 * (hi - lo + 1), dtype=xh.dtype).pin_memory()
    lib = sb._lib.load()
    p = g.cparams()
    import ctypes as C

    fn = lib.steps_b200_forces_f64 if rb == 8 else lib.steps_b200_forces_f32

    if world == 1:
        # the drop-in call of the reference's forces(): stateless, everything crosses the host boundary every call
        def e2e_call():
            rc = fn(C.byref(p), xh.data_ptr(), mh.data_ptr(), sh.data_ptr(), Fh.data_ptr(), lo, hi)
            if rc != 0:
                raise SystemExit("e2e: " + lib.steps_b200_last_error().decode())

        e2e_desc = (f"steps_b200_forces_{'f64' if rb == 8 else 'f32'}(params, x, M, soft, F, id_min, id_max) with pinned host buffers; wall clock around K "
                    "synchronous calls, max over ranks")
        h2d = (3 * N + 2 * N) * rb
    else:
        # one process per GPU: the collective force call of the resident engine with HOST positions in and HOST forces out
        # (upload_x = H2D of the full x replica, forces = all ranks together, download_forces = D2H of the owned rows);
        # a stateless per-rank sub-range call cannot exchange the j-side sums of the action-reaction evaluation
        xnp, Fnp = xh.numpy(), Fh.numpy()

        def e2e_call():
            eng.upload_x(xnp)
            eng.forces()
            rc = lib.steps_b200_engine_download_forces(eng._h, Fnp.ctypes.data, lo, hi)
            if rc != 0:
                raise SystemExit("e2e: " + lib.steps_b200_last_error().decode())

        e2e_desc = ("Engine.upload_x(x) + Engine.forces() [collective] + Engine.download_forces(own rows) with pinned host buffers; "
                    "wall clock around K synchronous calls, max over ranks")
        h2d = 3 * N * rb

    log(f"[bench] rank {rank}: timed steps done ({ms_step:.1f} ms/step), e2e leg ...")
    e2e_call()  # first call: creates the cached engine of the stateless path / warms the collective path
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_call()  # synchronous: returns after the D2H of F
    t_e2e = time.perf_counter() - t0
    barrier()
    t_e2e = max_over_ranks(t_e2e)
    e2e_value = args.steps * N * float(N) / t_e2e
    d2h = 3 * (hi - lo + 1) * rb
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(sum_over_ranks(float(h2d))),
           "d2h_bytes_per_step": int(sum_over_ranks(float(d2h))), "ms_per_step": 1e3 * t_e2e / args.steps, "call": e2e_desc}
    launches_e2e = 4 * args.steps  # pack + pair + reduce + tile_smax per call (lower bound: the action-reaction path adds one row reduction per pass)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpu = cpu_sample(c, args.cpu_seconds)[0]
        except Exception as ex:  # noqa: BLE001
            log(f"[bench] cpu_baseline failed: {ex}")
    eng.close()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "s_per_step": ms_step * 1e-3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64" if rb == 8 else "f32", "data": "synthetic", "config": workload_config(c, world, symmetric),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "gpu_launches_e2e": launches_e2e,
            "clocks": clocks, "tflops_20flop": FLOP_PER_PAIR * value / 1e12,
            "frac_of_fp_peak_whole_job": FLOP_PER_PAIR * value / 1e12 / (peak_sust * world),
            "implied_fma_clock_mhz": implied_mhz,
        }
        os.write(out_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c5"])
    ap.add_argument("--n", type=int, default=0, help="override N of config c2 / c5 (development only; the judged run uses the default)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline sample size in seconds")
    ap.add_argument("--ref-seconds", type=float, default=0.0, help="--impl reference: CPU seconds per sampled step (0 = auto, <= 10 s)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.steps < 1:
        raise SystemExit("--steps must be >= 1")
    # keep stdout clean for the ONE JSON line: everything else (including the reference's own printf) goes to stderr
    sys.stdout.flush()
    out_fd = os.dup(1)
    os.dup2(2, 1)
    try:
        if args.impl == "reference":
            run_reference(args, out_fd)
        else:
            run_ours(args, out_fd)
    except BaseException as ex:  # noqa: BLE001
        # under torchrun a rank that raises must die at once (no destructor may block on a collective), so that the
        # launcher tears the other ranks down instead of letting them wait for the NCCL timeout
        import traceback

        traceback.print_exc()
        sys.stderr.flush()
        if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not isinstance(ex, SystemExit):
            os._exit(1)
        raise


if __name__ == "__main__":
    main()
