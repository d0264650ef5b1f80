"""The product's device libm (rlshaders_b200/csrc/rls_libm.cuh) against the HOST C library.

CPU: the same __host__ __device__ source is compiled with g++ and compared bit for bit with
the host libm (tests/native/libm_check.cpp).  The default run strides through the binary32
bit patterns so the suite stays fast; RLS_LIBM_EXHAUSTIVE=1 walks all 2^32 arguments of every
univariate function and 2^30 argument pairs of the bivariate ones (about a minute on 8 cores;
result recorded in DESIGN.md: 0 mismatches).
GPU: the compiled DEVICE code is compared with the host libm on the path's argument ranges.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native", "libm_check.cpp")
EXE = os.path.join(ROOT, "tests", "native", "libm_check")


@pytest.fixture(scope="module")
def checker():
    deps = [SRC, os.path.join(ROOT, "rlshaders_b200", "csrc", "rls_libm.cuh")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fopenmp",
                        "-o", EXE, SRC, "-lm"], check=True)
    return EXE


@pytest.mark.parametrize("fn", ["sincos", "cos", "tan", "atan", "acos", "exp", "log", "atan2", "pow", "pow5", "pow5unit", "explean", "sincoslean", "loglean"])
def test_device_libm_source_matches_host_libm(checker, fn):
    stride = "1" if os.environ.get("RLS_LIBM_EXHAUSTIVE") else ("61" if fn in ("atan2", "pow", "pow5") else "253")
    out = subprocess.run([checker, fn, stride], check=True, capture_output=True, text=True).stdout.split()
    assert out[0] == fn and int(out[2]) > 1 << 22
    assert int(out[4]) == 0, " ".join(out)


def test_host_libm_probe_reports_ok_on_the_pinned_image():
    """parity.host_libm_status (RLS_HOST_LIBM_CHECK): this image's glibc 2.39 is the libm the device port reproduces, so
    the probe must say "ok" here; on another libm it says "degraded" and the GPU parity tests relax bit-identity."""
    import parity
    status, why = parity.host_libm_status()
    assert status == "ok", why


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _host(libm, name, *cols):
    f = getattr(libm, name)
    f.restype, f.argtypes = ctypes.c_float, [ctypes.c_float] * len(cols)
    return np.array([f(*[float(v) for v in row]) for row in zip(*cols)], dtype=np.float32)


@pytest.mark.gpu
def test_compiled_device_libm_matches_host_libm():
    import torch
    from rlshaders_b200 import api
    libm = ctypes.CDLL("libm.so.6")
    ctx = api.Context(0)
    n, m = 1 << 20, 1 << 16          # evaluated on the device / checked against ctypes calls
    rng = np.random.default_rng(7)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(ctx.device)   # noqa: E731

    angles = rng.uniform(-2 * np.pi, 2 * np.pi, n).astype(np.float32)
    unit = rng.uniform(-1, 1, n).astype(np.float32)
    cases = [("tanf", np.abs(angles) / 2), ("atanf", rng.standard_cauchy(n).astype(np.float32)), ("acosf", unit),
             ("expf", rng.uniform(-130, 5, n).astype(np.float32)), ("logf", rng.uniform(1e-30, 4, n).astype(np.float32))]
    for name, x in cases:
        got = ctx.debug_libm(name, dev(x)).cpu().numpy()
        assert np.array_equal(_bits(got[:m]), _bits(_host(libm, name, x[:m]))), name

    s, c = ctx.debug_libm("sincosf", dev(angles))
    s, c = s.cpu().numpy(), c.cpu().numpy()
    fs = libm.sincosf
    fs.restype = None
    fs.argtypes = [ctypes.c_float, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]
    hs, hc = ctypes.c_float(), ctypes.c_float()
    for i in range(1 << 14):
        fs(float(angles[i]), ctypes.byref(hs), ctypes.byref(hc))
        assert _bits(np.float32(hs.value)) == _bits(s[i]) and _bits(np.float32(hc.value)) == _bits(c[i])

    y, x = unit, rng.uniform(-1, 1, n).astype(np.float32)
    got = ctx.debug_libm("atan2f", dev(y), dev(x)).cpu().numpy()
    assert np.array_equal(_bits(got[:m]), _bits(_host(libm, "atan2f", y[:m], x[:m])))
    base, ex = np.abs(unit), rng.uniform(0, 1, n).astype(np.float32)
    ex[::2] = 5.0
    got = ctx.debug_libm("powf", dev(base), dev(ex)).cpu().numpy()
    assert np.array_equal(_bits(got[:m]), _bits(_host(libm, "powf", base[:m], ex[:m])))
    ctx.close()


@pytest.mark.gpu
def test_fast_policy_equals_exact_policy_exhaustively():
    """rls_fp.cuh / rls_libm.cuh: for EVERY binary32 argument (all 2^32 bit patterns of the univariate
    operations; 2^32 numerators / angles against a set of second operands for the bivariate ones) the
    guard-free fast-policy sequence either leaves the operand tracker unsatisfied (the sample is then
    re-run with the guarded operators) or returns the bits of the exact policy."""
    from rlshaders_b200 import api
    ctx = api.Context(0)
    try:
        full = 1 << 32
        for name in ("sqrt", "rcp", "tanf", "acosf", "div3", "expf", "sincosf"):
            ok, bad, rerun = ctx.debug_policy_check(name, 0, full)
            assert ok + rerun == full and bad == 0, (name, ok, bad, rerun)
            assert ok > (1 << 28), (name, ok)          # the window is not vacuous
        seconds = [1.0, -1.0, 3.0, 0.3, -0.7071068, 1e-4, 0.9999999, 1.0000001, 2.0 ** -30, 2.0 ** 40, 1e-20, -123456.7,
                   float(np.float32(np.pi)), 0.5, 1.5, 7.0]
        for name in ("div", "rdiv", "div_pz", "div_shared", "atan2f_yx", "atan2f_xy"):
            total_ok = 0
            for b in seconds:
                if name == "div_pz" and b <= 0.0:
                    continue
                ok, bad, rerun = ctx.debug_policy_check(name, 0, full // 16, stride=16, b=b)
                assert bad == 0, (name, b, ok, bad, rerun)
                total_ok += ok
            assert total_ok > (1 << 28), (name, total_ok)
    finally:
        ctx.close()
