"""2-GPU tests: engine data-parallel step == single-rank oracle on the concatenated batch, replicas bit-identical --
once with the gradients exchanged through NVLink peer memory inside the Adam kernel (the default), once with the
NCCL collectives in the step graph (OSRL_DP=nccl)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_equivalence_nccl(lib_built):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29613", os.path.join(ROOT, "tests", "dp_worker.py"), "bc", "bcql", "bearl", "cpq",
           "pipe:bcql", "pipe:bearl", "cdt"]
    env = dict(os.environ, OSRL_EXPECT_DP="peer")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    _keep_log("peer", r.stdout + ("" if r.returncode == 0 else "\n--- stderr ---\n" + r.stderr[-6000:]))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def _keep_log(tag, text):
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f"dp_2gpu_{tag}.log"), "w") as f:
        f.write(text)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_equivalence_nccl_collectives(lib_built):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29614", os.path.join(ROOT, "tests", "dp_worker.py"), "bcql", "pipe:bcql"]
    env = dict(os.environ, OSRL_DP="nccl", OSRL_EXPECT_DP="nccl")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    _keep_log("nccl", r.stdout + ("" if r.returncode == 0 else "\n--- stderr ---\n" + r.stderr[-6000:]))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
