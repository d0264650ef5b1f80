"""GPU: the kernels that were built, verified bit-exact and MEASURED SLOWER (rlshaders_b200/csrc/experiments/, built as
librls_b200_experiments.so by __graft_entry__.build_experiments / tools/build_experiments.sh): they must still give the
product's bits.  Skipped when that library has not been built; the product library contains none of this code."""
import os

import numpy as np
import pytest
import torch

import oracle_lib as ol
import parity
from rlshaders_b200 import _abi as abi
from rlshaders_b200 import _lib
from test_gpu_parity import FRAC_EXACT, FRAC_LOOSE, FRAC_TOL, N, _adversarial_shading, _pick, check, dev  # noqa: F401

pytestmark = pytest.mark.gpu
EXP = _lib.EXPERIMENTS_LIB_PATH


@pytest.fixture(scope="module", autouse=True)
def experiments_library():
    """The experiments library is not part of build(): (re)build it when it is missing or older than the sources (nvcc
    is on the GPU box too), skip when that is not possible -- a stale build would lack newer C-ABI symbols."""
    import __graft_entry__ as g
    try:
        g.build_experiments()
    except Exception as e:        # noqa: BLE001
        pytest.skip(f"experiments library cannot be built here: {e}")


@pytest.fixture(scope="module")
def ctx():
    from rlshaders_b200 import api
    c = api.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module", params=["reference", "port"])
def orc(request):
    o = ol.load_ref() if request.param == "reference" else ol.load_port()
    if o is None:
        pytest.skip("reference library not shipped")
    return o


def test_disney_lobe_partition_is_invisible(ctx, monkeypatch):
    """The CTA-level partition by specular lobe (k_disney_sample_eval_pdf<.., kLobeSort>) only changes which thread
    works on which sample: ragged sizes (partial last CTA, single sample), uniform and per-sample clearcoat, and the
    same bits as the default context (partition off)."""
    from rlshaders_b200 import api
    port = ol.load_port()
    monkeypatch.setenv("RLS_DISNEY_LOBE_SORT", "1")     # off by default (measured slower); read at rls_init
    plain, ctx = ctx, api.Context(0, lib_path=EXP)
    plain_owner = ctx
    monkeypatch.delenv("RLS_DISNEY_LOBE_SORT")
    kinds = dict(wi_s="dir", f_s="rel", pdf_s="rel", wi_d="dir", f_d="rel", pdf_d="rel", flags="flags")
    try:
        for n in (1, 31, 255, 257, 1000, 65536 + 77):
            sg, kw, u = parity.disney_inputs(n, seed=0x5EED0300 + n)
            for uniform_clearcoat in (None, 0.0, 1.0):
                if uniform_clearcoat is not None:
                    kw = dict(kw, clearcoat=uniform_clearcoat)
                cpu = port.disney_sample_eval_pdf(sg, abi.disney_params(**kw), *u)
                outs = []
                for c in (ctx, plain):
                    smp = api.DisneySampler(c, api.ShadingBatch.from_numpy(sg, c.device), **parity.params_to_dev(kw, c.device))
                    outs.append(smp.sampleEvalPdf(*[dev(t, c) for t in u]))
                    c.synchronize()
                for k in outs[0]:
                    assert torch.equal(outs[0][k].view(torch.int32), outs[1][k].view(torch.int32)), (n, uniform_clearcoat, k)
                st = parity.summarize(outs[0], cpu, kinds)
                assert st["flags"]["mismatches"] == 0, (n, uniform_clearcoat)
                for k in ("wi_s", "f_s", "pdf_s", "wi_d", "f_d", "pdf_d"):
                    assert st[k]["bit_exact"] >= FRAC_EXACT, (n, uniform_clearcoat, k, st[k])
    finally:
        plain_owner.close()


def _masked(run):
    """The packed experiment predates RLS_FLAG_SLOPE_EARLY_OUT and does not report it: compare the other bits."""
    _, gpu, cpu, _ = run
    kinds = dict(fresnel="rel", wi_r="dir", f_r="rel", pdf_r="rel", wi_t="dir", f_t="rel", weight_t="rel", flags="flags")
    gpu = dict(gpu, flags=gpu["flags"] & ~abi.FLAG_SLOPE_EARLY_OUT)
    cpu = dict(cpu, flags=cpu["flags"] & np.uint32(~abi.FLAG_SLOPE_EARLY_OUT & 0xffffffff))
    return parity.summarize(gpu, cpu, kinds)


def test_packed_kernel_is_bit_exact(orc, monkeypatch):
    """The two-samples-per-thread f32x2 kernel (rls_packed.cuh; off by default, RLS_PACKED=1): same
    bits as the oracle and as the scalar kernel, for even / odd batch sizes (tail lane), uniform
    and per-sample parameters, and operands that force the exact re-run of a pair."""
    from rlshaders_b200 import api
    monkeypatch.setenv("RLS_PACKED", "1")
    c = api.Context(0, lib_path=EXP)
    try:
        check(_masked(parity.run_ggx_dielectric(c, orc, N, aniso=True)), f"packed dielectric vs {orc.kind}")
        check(_masked(parity.run_ggx_dielectric(c, orc, 100003)), f"packed dielectric, odd n, vs {orc.kind}")
        assert c.fallback_count(reset=True) > 0
        n = 1 << 18
        sg = _adversarial_shading(n, 77)
        rx, ry = ol.hash_uniform(n, 77, 0), ol.hash_uniform(n, 77, 1)
        kw = dict(specularRoughness=_pick(n, 77, 2, [0.0, 1e-3, 0.05, 0.3, 1.0]), ior=_pick(n, 77, 3, [1.0, 0.47, 1.5, 2.5]))
        cpu = orc.ggx_dielectric(sg, abi.ggx_params(**kw), rx, ry)
        s = api.GgxSampler(c, api.ShadingBatch.from_numpy(sg, c.device), **parity.params_to_dev(kw, c.device))
        gpu = s.dielectricSampleEvalPdf(dev(rx, c), dev(ry, c))
        c.synchronize()
        check(_masked((None, gpu, cpu, None)), "packed dielectric, adversarial operands")
    finally:
        c.close()


# ------------------------------------ persistent TMA-staged kernels (rls_tile.cuh) vs plain kernels
def _tma_cases(c, n, seed, adversarial):
    """Outputs of every fused entry point that has a TMA-staged form, for one context."""
    from rlshaders_b200 import api
    out = {}
    if adversarial:
        sg = _adversarial_shading(n, seed)
        rough = _pick(n, seed, 2, [0.0, 1e-3, 0.01, 0.05, 0.3, 1.0, 1.0, 0.7])
        ior = _pick(n, seed, 3, [1.0, 1.0, 0.47, 1e-4, 1.5, 2.5, 1.33, 1.0001])
        aniso = _pick(n, seed, 4, [0.0, 0.0, 1.0, 0.5])
        rx, ry = ol.hash_uniform(n, seed, 0), ol.hash_uniform(n, seed, 1)
    else:
        sg, kw, rx, ry = parity.ggx_dielectric_inputs(n, seed, aniso=True)
        rough, ior, aniso = kw["specularRoughness"], kw["ior"], kw["anisotropic"]
    dsg = api.ShadingBatch.from_numpy(sg, c.device)
    drx, dry = dev(rx, c), dev(ry, c)
    g = api.GgxSampler(c, dsg, KsColor=(1.0, 0.5, 0.25), specularRoughness=dev(rough, c), ior=dev(ior, c),
                       anisotropic=dev(aniso, c))
    out["dielectric"] = g.dielectricSampleEvalPdf(drx, dry)
    gu = api.GgxSampler(c, dsg, KsColor=(1.0, 0.5, 0.25), specularRoughness=0.3, ior=1.5, anisotropic=0.25)
    out["dielectric_uniform"] = gu.dielectricSampleEvalPdf(drx, dry)
    c.synchronize()
    return {k: {kk: vv.cpu().numpy() for kk, vv in v.items()} for k, v in out.items()}


@pytest.mark.parametrize("n,adversarial", [(256, False), (256 * 7 + 37, False), (100003, True), (1 << 20, False)])
def test_tma_staged_kernels_match_plain_kernels(monkeypatch, n, adversarial):
    """The persistent TMA-staged forms (default) and the plain one-LDG-per-array kernels (RLS_TMA=0)
    give the same bits, for whole tiles, a ragged tail, uniform parameters and operands that force
    the exact re-run (which reloads its inputs from global memory)."""
    from rlshaders_b200 import api
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("RLS_TMA", mode)
        c = api.Context(0, lib_path=EXP)
        try:
            res[mode] = _tma_cases(c, n, 0x7A11 + n, adversarial)
            fb = c.fallback_count(reset=True)
            if adversarial:
                assert fb > 0
        finally:
            c.close()
    for case, outs in res["1"].items():
        for k, a in outs.items():
            b = res["0"][case][k]
            au, bu = a.view(np.uint32), b.view(np.uint32)
            bad = (au != bu)
            if a.dtype == np.float32:
                bad &= ~(np.isnan(a) & np.isnan(b))
            assert not bad.any(), f"{case}.{k}: {int(bad.sum())} / {bad.size} elements differ (TMA vs plain), n={n}"
