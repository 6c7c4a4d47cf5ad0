"""CPU: the N>1 host logic with world_size 2 and 3 over gloo (no GPU, no NCCL), and the reference arm of
bench.py under a multi-rank launch (rank 0 alone works and prints, the others exit 0 silently)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def torchrun(nproc, port, script, *args, timeout=300):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), script, *args]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("world,port", [(2, 29631), (3, 29632)])
def test_partition_exchange_and_reductions_over_gloo(world, port):
    r = torchrun(world, port, os.path.join(ROOT, "tests", "_gloo_worker.py"))
    assert r.returncode == 0, r.stderr[-2000:]
    assert f"GLOO_WORKER_OK {world}" in r.stdout


def test_reference_arm_under_two_ranks_prints_one_line_from_rank0():
    r = torchrun(2, 29633, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                 "--config", "c1", "--ref-seconds", "0.2")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "pair_interactions_per_s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
