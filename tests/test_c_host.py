"""examples/c_host/bcql_step.c: the C ABI driven from plain C (gcc, no Python, no torch).  On a machine without a GPU
the program gets as far as the parameter table (osrl_plan) and the loud failure of osrl_engine_create (exit code 3); on
a B200 it takes three BCQ-Lag steps and prints the statistics."""
import os
import re
import shutil
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _parse_stats(stdout: str) -> dict:
    """the `loss/<name> = <float>` lines of the program's output"""
    out = {}
    for ln in stdout.splitlines():
        m = re.fullmatch(r"(loss/\S+) = (\S+)", ln.strip())
        if m:
            out[m.group(1)] = float(m.group(2))
    return out


def test_parse_stats_lines():
    text = ("ABI 3: 122 state_dict tensors, 1563648 parameters; first = actor.pi.0.weight [256 x 10]\n"
            "loss/loss_vae = 0.281065\nloss/critic_loss = 9.4e-02\nloss/qc_penalty = 0\nlaunches per step: 47\n")
    assert _parse_stats(text) == {"loss/loss_vae": 0.281065, "loss/critic_loss": 0.094, "loss/qc_penalty": 0.0}


def _build(tmp_path, lib_built):
    exe = str(tmp_path / "bcql_step")
    libdir = os.path.join(ROOT, "osrl_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "c_host", "bcql_step.c"), "-L", libdir, "-losrl_b200",
                    f"-Wl,-rpath,{libdir}", "-o", exe], check=True)
    return exe


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_c_host_links_and_fails_loudly_without_gpu(tmp_path, lib_built):
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_c_host_steps_on_gpu")
    r = subprocess.run([_build(tmp_path, lib_built)], capture_output=True, text=True)
    assert r.returncode == 3, r.stdout + r.stderr
    assert "122 state_dict tensors" in r.stdout and "no CUDA device" in r.stderr


@pytest.mark.gpu
@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_c_host_steps_on_gpu(tmp_path, lib_built):
    r = subprocess.run([_build(tmp_path, lib_built)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    stats = _parse_stats(r.stdout)
    assert set(stats) >= {"loss/loss_vae", "loss/critic_loss", "loss/cost_critic_loss", "loss/actor_loss"}, r.stdout
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "c_host_bcql_step.log"), "w") as f:
        f.write(r.stdout)
