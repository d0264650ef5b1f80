"""The single-process multi-GPU host driver (rls_multi_* of include/rls_b200.h, rlshaders_b200/host/rls_driver --gpus N):
contiguous index ranges of one synthetic stream per device, the sweep's spp range sharded over the devices with the
NCCL all-reduce and CUDA graph behind the C ABI.  The 2-device checks run when the box has two GPUs; the partition
arithmetic itself is covered on the CPU (tests/test_host_logic.py, gloo world size 2)."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "rlshaders_b200", "host", "rls_driver")

pytestmark = pytest.mark.gpu


def run(*args):
    r = subprocess.run([EXE] + [str(a) for a in args], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])     # NCCL prints its version banner first


def n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("workload", ["dielectric", "disney", "skin"])
@pytest.mark.parametrize("policy", ["fast", "tolerant"])
def test_one_device_through_the_multi_driver(workload, policy):
    d = run("--gpus", 1, "--policy", policy, "--reps", 3, workload, 20)
    assert d["gpus"] == 1 and d["samples_per_s"] > 1e9 and len(d["check"]) == 1 and d["check"][0] > 0


def test_sweep_one_device_uses_the_graph():
    d = run("--gpus", 1, "--reps", 3, "sweep")
    assert d["graph_replays"] >= 3 and d["samples_per_s"] > 1e10


def test_two_devices_partition_the_stream_and_reduce_the_sweep():
    if n_gpus() < 2:
        pytest.skip("needs two GPUs")
    one = run("--gpus", 1, "--reps", 2, "dielectric", 22)
    two = run("--gpus", 2, "--reps", 2, "dielectric", 22)
    # device 0 of the 2-GPU run works on the same index range as the single device: same estimate, bit for bit
    assert two["check"][0] == one["check"][0] and two["check"][1] != one["check"][0]
    assert two["samples_per_s"] > 1.6 * one["samples_per_s"]
    s1 = run("--gpus", 1, "--reps", 2, "sweep")
    s2 = run("--gpus", 2, "--reps", 2, "sweep")
    assert s2["graph_replays"] >= 2
    for c in s2["check"]:                       # every device holds the full table after the all-reduce
        assert abs(c - s1["check"][0]) <= 1e-12 * abs(s1["check"][0])
    t2 = run("--gpus", 2, "--policy", "tolerant", "--reps", 2, "sweep")
    assert abs(t2["check"][0] - s1["check"][0]) <= 1e-5 * abs(s1["check"][0])
