"""The C++ host driver (rlshaders_b200/host): builds against the C ABI alone, refuses to run
without a device (CPU), and drives all four workloads through host buffers (GPU)."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "rlshaders_b200", "host", "rls_driver")


@pytest.fixture(scope="module")
def driver():
    import sys
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build_cuda()
    g.build_host_driver()
    return EXE


def test_driver_builds_and_has_no_cpu_path(driver):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([driver, "ggx", "10"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU path" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("workload,lo,hi", [("ggx", 0.5, 1.01), ("dielectric", 0.0, 0.6), ("disney", 0.0, 1.5), ("skin", 0.0, 1e9)])
def test_driver_runs_every_workload(driver, workload, lo, hi):
    r = subprocess.run([driver, workload, "20"], capture_output=True, text=True, check=True)
    d = json.loads(r.stdout)
    assert d["workload"] == workload and d["samples"] == 1 << 20
    assert d["host_to_host_samples_per_s"] > 1e7
    assert lo <= d["mean_f_over_pdf"] <= hi
    assert d["nodes"] == ["rlGgx", "rlDisney", "rlSkin"]


@pytest.mark.gpu
@pytest.mark.parametrize("workload", ["ggx", "dielectric", "disney"])
def test_driver_compact_frames_agree_with_full_frames(driver, workload):
    """`<workload>_q`: the same batch with unit quaternions uploaded and the frames decoded on the device (rls_*_hostq,
    rls_host.hpp quaternion_from_frame).  The decoded frames differ from U, V, N by one rounding (<= 4e-7), so the
    white-furnace style estimate over 2^20 samples agrees to 1e-3 relative."""
    a = json.loads(subprocess.run([driver, workload, "20"], capture_output=True, text=True, check=True).stdout)
    b = json.loads(subprocess.run([driver, workload + "_q", "20"], capture_output=True, text=True, check=True).stdout)
    assert b["workload"] == workload + "_q" and b["samples"] == a["samples"]
    assert abs(a["mean_f_over_pdf"] - b["mean_f_over_pdf"]) <= 1e-3 * abs(a["mean_f_over_pdf"]) + 1e-6
