"""N > 1 host-side paths on CPU (gloo, world size 2 and 3): the bucket-sharded query algorithm simulated rank for rank
over torch.distributed, and the reference arm of bench.py under torchrun (rank 0 alone works and prints)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def torchrun(nproc, port, script, *args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), script, *args]
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("world,port", [(2, 29531), (3, 29532)])
def test_bucket_sharded_algorithm_over_gloo(world, port):
    r = torchrun(world, port, os.path.join(ROOT, "tests", "gloo_shard_sim.py"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count(": ok") == 2 and "MISMATCH" not in r.stdout


def test_reference_arm_under_torchrun_prints_one_line_from_rank0():
    r = torchrun(2, 29533, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                 "--rows", "3000", "--queries", "40", "--max-node-size", "64", "--trees", "2", "--dim", "64")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["unit"] == "queries/s"
