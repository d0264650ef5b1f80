"""Host-side mirror of the reference's sampler interface over the C ABI.

The reference hands Arnold three callbacks per sampler object -- evalSample(rx, ry),
evalBrdf(indir), evalPdf(indir) (src/rlGgx.h:97-127, src/rlDisney.cpp:109-152) -- on a
sampler constructed per shading point from AtShaderGlobals + node parameters.  Here the
same names operate on BATCHES: a sampler object is constructed from a `ShadingBatch`
(one shading frame + view vector per sample) and node parameters that are either
uniform scalars or per-sample arrays, and every call processes the whole batch on the
GPU through librls_b200.so.

torch is used for device memory and streams only.  Tensors on a CUDA device go to the
device entry points (asynchronous on the context's stream); pinned CPU tensors go to the
`*_host` entry points (chunked H2D -> kernel -> D2H pipeline, synchronous).
"""
import ctypes as C

import torch

from . import _abi as abi
from ._lib import load


class RlsError(RuntimeError):
    pass


def _check(ctx_handle, rc, lib):
    if rc != abi.RLS_OK:
        msg = lib.rls_last_error_string(ctx_handle)
        raise RlsError(f"rls error {rc}: {msg.decode() if msg else ''}")


class Context:
    """Owns one rls_context bound to a CUDA device and a stream (default: torch's current
    stream on that device, so torch.cuda.Event timing sees the kernels)."""

    def __init__(self, device=0, stream=None, lib_path=None):
        self.lib = load(lib_path)
        self.index = int(device if not isinstance(device, torch.device) else (device.index or 0))
        handle = C.c_void_p()
        if stream is None and torch.cuda.is_available():
            stream = torch.cuda.current_stream(self.index)
        self.stream = stream
        # torch's default stream has handle 0, which rls_init reads as "create a private
        # stream"; name the legacy default stream explicitly (cudaStreamLegacy == 0x1) so that
        # torch tensors, torch.cuda.Event timing and the kernels share one stream.
        raw = C.c_void_p(stream.cuda_stream or 1) if stream is not None else None
        rc = self.lib.rls_init(self.index, raw, C.byref(handle))
        if rc != abi.RLS_OK:
            msg = self.lib.rls_last_error_string(None)
            raise RlsError(f"rls_init failed ({rc}): {msg.decode() if msg else ''}")
        self.handle = handle
        self.device = torch.device("cuda", self.index)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.rls_shutdown(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        _check(self.handle, self.lib.rls_synchronize(self.handle), self.lib)

    @property
    def kernel_launches(self):
        return int(self.lib.rls_kernel_launch_count(self.handle))

    def set_arith_policy(self, policy):
        """"fast" (default: guard-free IEEE sequences + exact re-run of out-of-window samples) or "exact" (guarded
        operators only): same bits either way (csrc/rls_fp.cuh).  "tolerant" (opt-in): the fused units to a stated
        tolerance with bit-exact flags (csrc/rls_tol.cuh, include/rls_b200.h RLS_ARITH_TOLERANT)."""
        code = {"fast": 0, "exact": 1, "tolerant": 2}[policy]
        _check(self.handle, self.lib.rls_set_arith_policy(self.handle, code), self.lib)

    def fallback_count(self, reset=False):
        """Samples the fused kernels re-ran with the guarded operators (synchronises)."""
        import ctypes
        v = ctypes.c_uint64(0)
        _check(self.handle, self.lib.rls_fallback_count(self.handle, ctypes.byref(v), int(bool(reset))), self.lib)
        return int(v.value)

    # ---- allocation helpers -------------------------------------------------
    def empty(self, *shape, dtype=torch.float32, like=None):
        if like is not None and like.device.type == "cpu":
            return torch.empty(*shape, dtype=dtype, pin_memory=True)
        return torch.empty(*shape, dtype=dtype, device=self.device)

    # ---- synthetic workload generators --------------------------------------
    def synth_uniform(self, n, seed, stream, first_index=0, lo=0.0, hi=1.0, out=None):
        out = self.empty(n) if out is None else out
        _check(self.handle, self.lib.rls_synth_uniform(self.handle, n, seed, stream, first_index,
                                                       lo, hi, out.data_ptr()), self.lib)
        return out

    def synth_shading(self, n, seed, first_index=0, cos_lo=0.02, cos_hi=1.0, backfacing_fraction=0.0):
        sg = ShadingBatch(self.empty(3, n), self.empty(3, n), self.empty(3, n), self.empty(3, n),
                          self.empty(n, dtype=torch.uint8) if backfacing_fraction > 0 else None)
        _check(self.handle, self.lib.rls_synth_shading(self.handle, n, seed, first_index, cos_lo, cos_hi,
                                                       backfacing_fraction, C.byref(sg.struct)), self.lib)
        return sg

    # ---- diagnostics ------------------------------------------------------------
    LIBM_FUNCTIONS = dict(sincosf=0, tanf=1, atanf=2, acosf=3, expf=4, logf=5, atan2f=6, powf=7)

    def debug_libm(self, name, a, b=None):
        """Element-wise evaluation of the library's own device libm (for parity tests)."""
        fn = self.LIBM_FUNCTIONS[name]
        out0 = torch.empty_like(a)
        out1 = torch.empty_like(a) if fn == 0 else None
        _check(self.handle, self.lib.rls_debug_libm(self.handle, fn, a.numel(), a.data_ptr(),
                                                    b.data_ptr() if b is not None else None, out0.data_ptr(),
                                                    out1.data_ptr() if out1 is not None else None), self.lib)
        return (out0, out1) if fn == 0 else out0

    POLICY_CHECKS = dict(sqrt=0, rcp=1, div=2, tanf=3, acosf=4, atan2f_yx=5, atan2f_xy=6, div_pz=7, rdiv=8, div3=9, div_shared=10, expf=11, sincosf=12)

    def debug_policy_check(self, name, first_bits, count, stride=1, b=1.0):
        """Fast-policy == exact-policy over the binary32 bit patterns first_bits + k*stride, k < count.
        Returns (accepted by the operand tracker, mismatches among them, sent to the exact re-run)."""
        counts = torch.zeros(3, dtype=torch.int64, device=self.device)
        _check(self.handle, self.lib.rls_debug_policy_check(self.handle, self.POLICY_CHECKS[name], int(first_bits) & 0xffffffff,
                                                            int(count), int(stride), float(b), counts.data_ptr()), self.lib)
        self.synchronize()
        return tuple(int(v) for v in counts.cpu().tolist())

    # ---- albedo sweep ----------------------------------------------------------
    def albedo_sweep(self, grid, seed, spp_begin, spp_end, out=None):
        cells = grid.n_rough * grid.n_cos * grid.n_ior
        if out is None:
            out = torch.empty(cells, abi.SWEEP_VALUES_PER_CELL, dtype=torch.float64, device=self.device)
        _check(self.handle, self.lib.rls_albedo_sweep(self.handle, C.byref(grid), seed, spp_begin, spp_end,
                                                      out.data_ptr()), self.lib)
        return out


def _placed(t, name, ctx, host):
    """Every pointer handed to the C ABI must live where the entry point reads it: on the context's device for the
    device forms, in host memory for the *_host forms (a wrong placement is an illegal-address fault, i.e. a dead CUDA
    context, not an exception)."""
    if ctx is None:
        return
    if host:
        if t.device.type != "cpu":
            raise ValueError(f"{name}: the host-buffer form needs a CPU (pinned) tensor, got {t.device}")
    elif t.device != ctx.device:
        raise ValueError(f"{name}: expected a tensor on {ctx.device}, got {t.device} "
                         "(only the fused *SampleEvalPdf calls have a host-buffer form)")


def _f32rows(t, name, n=None, ctx=None, host=False):
    if t.dtype != torch.float32 or t.dim() != 2 or t.shape[0] != 3 or not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous float32 tensor of shape [3, n]")
    if n is not None and t.shape[1] != n:
        raise ValueError(f"{name}: expected {n} samples, got {t.shape[1]}")
    _placed(t, name, ctx, host)
    return (t[0], t[1], t[2])


def _f32(t, name, n=None, ctx=None, host=False):
    if t.dtype != torch.float32 or t.dim() != 1 or not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous float32 tensor of shape [n]")
    if n is not None and t.shape[0] != n:
        raise ValueError(f"{name}: expected {n} samples, got {t.shape[0]}")
    _placed(t, name, ctx, host)
    return t


def _i32(t, name, n, ctx=None, host=False):
    if t.dtype != torch.int32 or t.dim() != 1 or not t.is_contiguous() or t.shape[0] != n:
        raise ValueError(f"{name}: expected a contiguous int32 tensor of shape [{n}]")
    _placed(t, name, ctx, host)
    return t


def _params_placed(params_kw, n, ctx, host):
    """Node-parameter tensors (per-sample arrays) of a sampler: length n, float32, placed like the shading batch."""
    for k, v in params_kw.items():
        for j, t in enumerate(v if isinstance(v, (tuple, list)) else (v,)):
            if isinstance(t, torch.Tensor):
                _f32(t, f"{k}[{j}]" if isinstance(v, (tuple, list)) else k, n, ctx, host)


class ShadingBatch:
    """The AtShaderGlobals fields the path reads, one per sample: frame (U, V, N = sg->Nf),
    view direction wo = -sg->Rd, optional `backfacing` (sg->N == -sg->Nf).  [3, n] tensors."""

    def __init__(self, U, V, N, wo, backfacing=None):
        self.U, self.V, self.N, self.wo, self.backfacing = U, V, N, wo, backfacing
        self.n = n = U.shape[1] if U.dim() == 2 else -1
        self.on_host = U.device.type == "cpu"
        for name, t in (("V", V), ("N", N), ("wo", wo), ("backfacing", backfacing)):
            if t is not None and t.device != U.device:
                raise ValueError(f"ShadingBatch: {name} is on {t.device}, U on {U.device}")
        if backfacing is not None and (backfacing.dtype != torch.uint8 or backfacing.dim() != 1 or backfacing.shape[0] != n
                                       or not backfacing.is_contiguous()):
            raise ValueError(f"backfacing: expected a contiguous uint8 tensor of shape [{n}]")
        self.struct = abi.shading(_f32rows(U, "U", n), _f32rows(V, "V", n), _f32rows(N, "N", n),
                                  _f32rows(wo, "wo", n), backfacing)

    def placed(self, ctx, host=False):
        """Raises unless the batch lives where the entry point reads it (see _placed)."""
        _placed(self.U, "ShadingBatch", ctx, host)
        return self

    @classmethod
    def from_numpy(cls, sg, device=None, pin=False):
        """From the dict-of-arrays form used by the tests (keys Ux..woz, backfacing)."""
        import numpy as np

        def mk(prefix):
            t = torch.from_numpy(np.stack([sg[prefix + c] for c in "xyz"]))
            if device is not None:
                return t.to(device)
            return t.pin_memory() if pin else t
        bf = sg.get("backfacing")
        if bf is not None:
            bf = torch.from_numpy(bf)
            bf = bf.to(device) if device is not None else (bf.pin_memory() if pin else bf)
        return cls(mk("U"), mk("V"), mk("N"), mk("wo"), bf)


class QuatShadingBatch:
    """Compact shading batch for the host-buffer forms (include/rls_b200.h rls_shading_quat_soa): a unit quaternion
    q = (x, y, z, w) [4, n] in place of the nine floats of U, V, N -- PCIe bounds those forms, and this cuts the upload
    from 65 to 45 B per rlGgx sample.  The frame is DEFINED by the header's binary32 decode (`decode(ctx)` runs it on
    the device; the test suite restates it in numpy).  CPU (pinned) tensors only."""

    def __init__(self, q, wo, backfacing=None):
        self.q, self.wo, self.backfacing = q, wo, backfacing
        if q.dtype != torch.float32 or q.dim() != 2 or q.shape[0] != 4 or not q.is_contiguous():
            raise ValueError("q: expected a contiguous float32 tensor of shape [4, n]")
        self.n = n = q.shape[1]
        self.on_host = True
        self.quat = True
        for name, t in (("q", q), ("wo", wo), ("backfacing", backfacing)):
            if t is not None and t.device.type != "cpu":
                raise ValueError(f"QuatShadingBatch: {name} must be a CPU (pinned) tensor; decode(ctx) gives the device frame")
        if backfacing is not None and (backfacing.dtype != torch.uint8 or backfacing.dim() != 1 or backfacing.shape[0] != n
                                       or not backfacing.is_contiguous()):
            raise ValueError(f"backfacing: expected a contiguous uint8 tensor of shape [{n}]")
        self.struct = abi.shading_quat((q[0], q[1], q[2], q[3]), _f32rows(wo, "wo", n), backfacing)

    def placed(self, ctx, host=False):
        if not host:
            raise ValueError("QuatShadingBatch: only the fused *SampleEvalPdf calls (host-buffer form) take compact frames")
        return self

    def decode(self, ctx):
        """-> ShadingBatch on ctx.device with U, V, N decoded by rls_frame_from_quaternion."""
        q = self.q.to(ctx.device)
        U, V, N = ctx.empty(3, self.n), ctx.empty(3, self.n), ctx.empty(3, self.n)
        _check(ctx.handle, ctx.lib.rls_frame_from_quaternion(ctx.handle, self.n, q[0].data_ptr(), q[1].data_ptr(),
                                                             q[2].data_ptr(), q[3].data_ptr(), abi.vec3((U[0], U[1], U[2])),
                                                             abi.vec3((V[0], V[1], V[2])), abi.vec3((N[0], N[1], N[2]))), ctx.lib)
        ctx.synchronize()
        return ShadingBatch(U, V, N, self.wo.to(ctx.device),
                            self.backfacing.to(ctx.device) if self.backfacing is not None else None)


def _host_form(lib, name, sg):
    """The host-buffer entry point of a fused unit: *_hostq for compact frames, *_host otherwise."""
    return getattr(lib, name + ("_hostq" if getattr(sg, "quat", False) else "_host"))


def _param(v):
    """Node-parameter value -> what _abi.param1/param3 accept (scalars stay uniform)."""
    return v


class GgxSampler:
    """Batched rls::GgxSampler (src/rlGgx.h:92-375).  Constructor arguments follow the
    reference ctor (sg, specColor, ior, roughness, anisotropic) under their node-parameter
    names (src/rlGgx.cpp:172-186); each is a scalar or a per-sample tensor."""

    def __init__(self, ctx, sg, KsColor=(1.0, 1.0, 1.0), ior=1.0, specularRoughness=0.0, anisotropic=0.0,
                 normal_sampler=abi.GGX_SAMPLER_VNDF, **ignored_node_params):
        self.ctx, self.sg = ctx, sg
        _params_placed(dict(KsColor=KsColor, ior=ior, specularRoughness=specularRoughness, anisotropic=anisotropic,
                            **ignored_node_params), sg.n, ctx, sg.on_host)
        self.params = abi.ggx_params(KsColor=KsColor, ior=ior, specularRoughness=specularRoughness,
                                     anisotropic=anisotropic, normal_sampler=normal_sampler,
                                     **ignored_node_params)

    def evalSample(self, rx, ry, want_fresnel=True):
        n, c = self.sg.n, self.ctx
        self.sg.placed(c)
        wi = c.empty(3, n)
        F = c.empty(n) if want_fresnel else None
        _check(c.handle, c.lib.rls_ggx_eval_sample(
            c.handle, n, C.byref(self.sg.struct), C.byref(self.params), _f32(rx, "rx", n, c).data_ptr(),
            _f32(ry, "ry", n, c).data_ptr(), abi.vec3(_f32rows(wi, "wi")), F.data_ptr() if F is not None else None), c.lib)
        return wi, F

    def evalLightSample(self, Ld, Li, light_pdf, rx=None, ry=None, Li_at_l=None, pdf_at_l=None):
        """AiEvaluateLightSample-shaped MIS of one light sample (src/rlGgx.h:167-170); the BRDF
        half is included when the light's radiance / pdf along L = evalSample(rx, ry) are given."""
        return _eval_light_sample(self.ctx.lib.rls_ggx_evaluate_light_sample, self, (), Ld, Li, light_pdf,
                                  rx, ry, Li_at_l, pdf_at_l)

    def evalBrdf(self, wi):
        n, c = self.sg.n, self.ctx
        self.sg.placed(c)
        f = c.empty(3, n)
        _check(c.handle, c.lib.rls_ggx_eval_brdf(c.handle, n, C.byref(self.sg.struct), C.byref(self.params),
                                                 abi.vec3(_f32rows(wi, "wi", n, c)), abi.vec3(_f32rows(f, "f"))), c.lib)
        return f

    def evalPdf(self, wi):
        n, c = self.sg.n, self.ctx
        self.sg.placed(c)
        pdf = c.empty(n)
        _check(c.handle, c.lib.rls_ggx_eval_pdf(c.handle, n, C.byref(self.sg.struct), C.byref(self.params),
                                                abi.vec3(_f32rows(wi, "wi", n, c)), pdf.data_ptr()), c.lib)
        return pdf

    def getRefractDirection(self, m):
        """getRefractDirection(m, V) (src/rlGgx.h:277-291) for a microfacet normal m: (wi, flags); wi is the zero
        vector and flags carries RLS_FLAG_TIR where refraction fails."""
        n, c = self.sg.n, self.ctx
        self.sg.placed(c)
        wi, fl = c.empty(3, n), c.empty(n, dtype=torch.int32)
        _check(c.handle, c.lib.rls_ggx_refract_direction(c.handle, n, C.byref(self.sg.struct), C.byref(self.params),
                                                         abi.vec3(_f32rows(m, "m", n, c)), abi.vec3(_f32rows(wi, "wi")),
                                                         fl.data_ptr()), c.lib)
        return wi, fl

    def refraction(self, wi):
        """refraction(V, wi, N) (src/rlGgx.h:316-328): the BTDF at a transmitted direction."""
        n, c = self.sg.n, self.ctx
        self.sg.placed(c)
        ft = c.empty(n)
        _check(c.handle, c.lib.rls_ggx_eval_btdf(c.handle, n, C.byref(self.sg.struct), C.byref(self.params),
                                                 abi.vec3(_f32rows(wi, "wi", n, c)), ft.data_ptr()), c.lib)
        return ft

    def getSampleWeight(self, wi, m):
        """getSampleWeight(V, wi, m) (src/rlGgx.h:294-301)."""
        n, c = self.sg.n, self.ctx
        self.sg.placed(c)
        w = c.empty(n)
        _check(c.handle, c.lib.rls_ggx_sample_weight(c.handle, n, C.byref(self.sg.struct), C.byref(self.params),
                                                     abi.vec3(_f32rows(wi, "wi", n, c)), abi.vec3(_f32rows(m, "m", n, c)),
                                                     w.data_ptr()), c.lib)
        return w

    def alloc_out(self, like, want_fresnel=True):
        n, c = self.sg.n, self.ctx
        return dict(wi=c.empty(3, n, like=like), f=c.empty(3, n, like=like), pdf=c.empty(n, like=like),
                    fresnel=c.empty(n, like=like) if want_fresnel else None,
                    flags=c.empty(n, dtype=torch.int32, like=like))

    def sampleEvalPdf(self, rx, ry, out=None, want_fresnel=True, chunk=0):
        """The fused unit of work: ctor + evalSample + evalBrdf(L) + evalPdf(L)."""
        n, c = self.sg.n, self.ctx
        out = self.alloc_out(rx, want_fresnel) if out is None else out
        h = self.sg.on_host
        self.sg.placed(c, h)
        o = abi.BsdfOut(abi.vec3(_f32rows(out["wi"], "wi", n, c, h)), abi.vec3(_f32rows(out["f"], "f", n, c, h)),
                        _f32(out["pdf"], "pdf", n, c, h).data_ptr(),
                        _f32(out["fresnel"], "fresnel", n, c, h).data_ptr() if out.get("fresnel") is not None else None,
                        _i32(out["flags"], "flags", n, c, h).data_ptr())
        args = (c.handle, n, C.byref(self.sg.struct), C.byref(self.params), _f32(rx, "rx", n, c, h).data_ptr(),
                _f32(ry, "ry", n, c, h).data_ptr(), C.byref(o))
        if self.sg.on_host:
            _check(c.handle, _host_form(c.lib, "rls_ggx_sample_eval_pdf", self.sg)(*args, chunk), c.lib)
        else:
            _check(c.handle, c.lib.rls_ggx_sample_eval_pdf(*args), c.lib)
        return out

    def alloc_dielectric_out(self, like):
        n, c = self.sg.n, self.ctx
        e = lambda *s, **k: c.empty(*s, like=like, **k)   # noqa: E731
        return dict(fresnel=e(n), wi_r=e(3, n), f_r=e(n), pdf_r=e(n), wi_t=e(3, n), f_t=e(n), weight_t=e(n),
                    flags=e(n, dtype=torch.int32))

    def dielectricSampleEvalPdf(self, rx, ry, out=None, chunk=0):
        """Rough dielectric (Walter'07): one visible-normal sample, reflection AND
        refraction branches (src/rlGgx.h:228-243, 277-328)."""
        n, c = self.sg.n, self.ctx
        out = self.alloc_dielectric_out(rx) if out is None else out
        h = self.sg.on_host
        self.sg.placed(c, h)
        sc = lambda k: _f32(out[k], k, n, c, h).data_ptr()   # noqa: E731
        o = abi.GgxDielectricOut(sc("fresnel"), abi.vec3(_f32rows(out["wi_r"], "wi_r", n, c, h)), sc("f_r"), sc("pdf_r"),
                                 abi.vec3(_f32rows(out["wi_t"], "wi_t", n, c, h)), sc("f_t"), sc("weight_t"),
                                 _i32(out["flags"], "flags", n, c, h).data_ptr())
        args = (c.handle, n, C.byref(self.sg.struct), C.byref(self.params), _f32(rx, "rx", n, c, h).data_ptr(),
                _f32(ry, "ry", n, c, h).data_ptr(), C.byref(o))
        if self.sg.on_host:
            _check(c.handle, _host_form(c.lib, "rls_ggx_dielectric_sample_eval_pdf", self.sg)(*args, chunk), c.lib)
        else:
            _check(c.handle, c.lib.rls_ggx_dielectric_sample_eval_pdf(*args), c.lib)
        return out


def _light(direction, radiance, pdf, n, ctx):
    d = _f32rows(direction, "Ld", n, ctx) if direction is not None else None
    return abi.light_sample(d, _f32rows(radiance, "Li", n, ctx), _f32(pdf, "light pdf", n, ctx))


def _eval_light_sample(fn, sampler, head_args, Ld, Li, light_pdf, rx, ry, Li_at_l, pdf_at_l):
    """Shared body of {GgxSampler,DisneySampler}.evalLightSample (include/rls_b200.h, f3)."""
    n, c = sampler.sg.n, sampler.ctx
    sampler.sg.placed(c)
    light = _light(Ld, Li, light_pdf, n, c)
    at_l = None
    if Li_at_l is not None:
        at_l = _light(None, Li_at_l, pdf_at_l, n, c)
    rgb = c.empty(3, n)
    wl, wb = c.empty(n), c.empty(n)
    _check(c.handle, fn(c.handle, n, C.byref(sampler.sg.struct), C.byref(sampler.params), *head_args, C.byref(light),
                        _f32(rx, "rx", n, c).data_ptr() if rx is not None else None,
                        _f32(ry, "ry", n, c).data_ptr() if ry is not None else None,
                        C.byref(at_l) if at_l is not None else None,
                        abi.vec3(_f32rows(rgb, "rgb")), wl.data_ptr(), wb.data_ptr()), c.lib)
    return dict(rgb=rgb, w_light=wl, w_brdf=wb)


class SampleWriter:
    """rls::SampleWriter (src/rlUtil.h:44-171) for one shading point of a sampler: writeRadiance
    fills the lat-long BRDF image, writeSample paints the sample scatter over it (green; red
    below the horizon).  `image` is [3, height, width] float32 in the writer's B, G, R plane
    order; save() writes the scanline OpenEXR file the reference's writer produces (HALF
    channels B, G, R), or a .npy."""

    def __init__(self, ctx, width, height, outpath=""):
        self.ctx, self.width, self.height, self.outpath = ctx, int(width), int(height), outpath
        self.image = torch.zeros(3, self.height, self.width, dtype=torch.float32, device=ctx.device)
        self._scratch = torch.empty(self.height * self.width, dtype=torch.int32, device=ctx.device)
        self.missing = torch.zeros(1, dtype=torch.int32, device=ctx.device)

    def _node(self, brdf):
        if isinstance(brdf, GgxSampler):
            return abi.NODE_GGX, 0
        return abi.NODE_DISNEY, brdf.sample_type

    def writeRadiance(self, brdf, point=0):
        c = self.ctx
        node, st = self._node(brdf)
        brdf.sg.placed(c)
        if not 0 <= int(point) < brdf.sg.n:
            raise ValueError("point index out of range")
        _check(c.handle, c.lib.rls_sample_writer_radiance(c.handle, node, C.byref(brdf.sg.struct),
                                                          C.cast(C.byref(brdf.params), C.c_void_p), point, st,
                                                          self.width, self.height, self.image.data_ptr()), c.lib)
        return self.image

    def writeSample(self, brdf, rx, ry, point=0):
        c = self.ctx
        node, st = self._node(brdf)
        n = rx.shape[0]
        brdf.sg.placed(c)
        _check(c.handle, c.lib.rls_sample_writer_scatter(c.handle, node, C.byref(brdf.sg.struct),
                                                         C.cast(C.byref(brdf.params), C.c_void_p), point, st, n,
                                                         _f32(rx, "rx", n, c).data_ptr(), _f32(ry, "ry", n, c).data_ptr(),
                                                         self.width, self.height, self.image.data_ptr(),
                                                         self._scratch.data_ptr(), self.missing.data_ptr()), c.lib)
        return self.image

    def save(self, path=None):
        from . import exr
        path = path or self.outpath
        img = self.image.cpu().numpy()
        if path.endswith(".npy"):
            import numpy as np
            np.save(path, img)
        else:
            exr.write_scanline_exr(path, img, ("B", "G", "R"), half=True)
        return path


class DisneySampler:
    """Batched DisneySampler (src/rlDisney.cpp:105-602).  Node parameters by name
    (src/rlDisney.cpp:606-610); setSampleType mirrors :194."""

    def __init__(self, ctx, sg, **node_params):
        self.ctx, self.sg = ctx, sg
        _params_placed(node_params, sg.n, ctx, sg.on_host)
        self.params = abi.disney_params(**node_params)
        self.sample_type = abi.RLS_RAY_GLOSSY

    def setSampleType(self, sample_type):
        if sample_type not in (abi.RLS_RAY_DIFFUSE, abi.RLS_RAY_GLOSSY):
            raise ValueError("sample type must be RLS_RAY_DIFFUSE or RLS_RAY_GLOSSY")
        self.sample_type = sample_type

    def evalLightSample(self, Ld, Li, light_pdf, rx=None, ry=None, Li_at_l=None, pdf_at_l=None):
        """evalDiffuseLightSample / evalSpecularLightSample (src/rlDisney.cpp:266-277) for the
        current sample type, as a two-sample power-heuristic MIS (include/rls_b200.h, f3)."""
        return _eval_light_sample(self.ctx.lib.rls_disney_evaluate_light_sample, self, (self.sample_type,),
                                  Ld, Li, light_pdf, rx, ry, Li_at_l, pdf_at_l)

    def evalSample(self, rx, ry):
        n, c = self.sg.n, self.ctx
        self.sg.placed(c)
        wi = c.empty(3, n)
        flags = c.empty(n, dtype=torch.int32)
        _check(c.handle, c.lib.rls_disney_eval_sample(
            c.handle, n, C.byref(self.sg.struct), C.byref(self.params), self.sample_type,
            _f32(rx, "rx", n, c).data_ptr(), _f32(ry, "ry", n, c).data_ptr(), abi.vec3(_f32rows(wi, "wi")),
            flags.data_ptr()), c.lib)
        return wi, flags

    def evalBrdf(self, wi):
        n, c = self.sg.n, self.ctx
        self.sg.placed(c)
        f = c.empty(3, n)
        _check(c.handle, c.lib.rls_disney_eval_brdf(c.handle, n, C.byref(self.sg.struct), C.byref(self.params),
                                                    self.sample_type, abi.vec3(_f32rows(wi, "wi", n, c)),
                                                    abi.vec3(_f32rows(f, "f"))), c.lib)
        return f

    def evalPdf(self, wi):
        n, c = self.sg.n, self.ctx
        self.sg.placed(c)
        pdf = c.empty(n)
        _check(c.handle, c.lib.rls_disney_eval_pdf(c.handle, n, C.byref(self.sg.struct), C.byref(self.params),
                                                   self.sample_type, abi.vec3(_f32rows(wi, "wi", n, c)),
                                                   pdf.data_ptr()), c.lib)
        return pdf

    def alloc_out(self, like):
        n, c = self.sg.n, self.ctx
        e = lambda *s, **k: c.empty(*s, like=like, **k)   # noqa: E731
        return dict(wi_s=e(3, n), f_s=e(3, n), pdf_s=e(n), wi_d=e(3, n), f_d=e(3, n), pdf_d=e(n),
                    flags=e(n, dtype=torch.int32))

    def sampleEvalPdf(self, rx_s, ry_s, rx_d, ry_d, out=None, chunk=0):
        """Fused: ctor + glossy triple on (rx_s, ry_s) + diffuse triple on (rx_d, ry_d)."""
        n, c = self.sg.n, self.ctx
        out = self.alloc_out(rx_s) if out is None else out
        h = self.sg.on_host
        self.sg.placed(c, h)
        v3 = lambda k: abi.vec3(_f32rows(out[k], k, n, c, h))   # noqa: E731
        o = abi.DisneyOut(v3("wi_s"), v3("f_s"), _f32(out["pdf_s"], "pdf_s", n, c, h).data_ptr(), v3("wi_d"), v3("f_d"),
                          _f32(out["pdf_d"], "pdf_d", n, c, h).data_ptr(), _i32(out["flags"], "flags", n, c, h).data_ptr())
        args = (c.handle, n, C.byref(self.sg.struct), C.byref(self.params), _f32(rx_s, "rx_s", n, c, h).data_ptr(),
                _f32(ry_s, "ry_s", n, c, h).data_ptr(), _f32(rx_d, "rx_d", n, c, h).data_ptr(),
                _f32(ry_d, "ry_d", n, c, h).data_ptr(), C.byref(o))
        if self.sg.on_host:
            _check(c.handle, _host_form(c.lib, "rls_disney_sample_eval_pdf", self.sg)(*args, chunk), c.lib)
        else:
            _check(c.handle, c.lib.rls_disney_sample_eval_pdf(*args), c.lib)
        return out


class NDProfile:
    """Batched rls::NDProfile (src/rlSss.h:27-61, src/rlSss.cpp:20-106)."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.state = None

    def setDistance(self, dist, albedo):
        n, c = dist.shape[1], self.ctx
        self.n = n
        self.state = dict(distance=c.empty(3, n), C1=c.empty(3, n), C2=c.empty(3, n), max_radius=c.empty(n))
        self._struct = abi.NdProfileSoA(abi.vec3(_f32rows(self.state["distance"], "distance")),
                                        abi.vec3(_f32rows(self.state["C1"], "C1")),
                                        abi.vec3(_f32rows(self.state["C2"], "C2")),
                                        self.state["max_radius"].data_ptr())
        _check(c.handle, c.lib.rls_ndprofile_set_distance(c.handle, n, abi.vec3(_f32rows(dist, "dist", n, c)),
                                                          abi.vec3(_f32rows(albedo, "albedo", n, c)),
                                                          C.byref(self._struct)), c.lib)
        return self.state

    def maxRadius(self):
        return self.state["max_radius"]

    def getRadius(self, rx):
        c = self.ctx
        r, fl = c.empty(self.n), c.empty(self.n, dtype=torch.int32)
        _check(c.handle, c.lib.rls_ndprofile_get_radius(c.handle, self.n, C.byref(self._struct),
                                                        _f32(rx, "rx", self.n, c).data_ptr(), r.data_ptr(),
                                                        fl.data_ptr()), c.lib)
        return r, fl

    def getPdf(self, r):
        c = self.ctx
        pdf = c.empty(self.n)
        _check(c.handle, c.lib.rls_ndprofile_get_pdf(c.handle, self.n, C.byref(self._struct),
                                                     _f32(r, "r", self.n, c).data_ptr(), pdf.data_ptr()), c.lib)
        return pdf

    def evalProfile(self, r):
        c = self.ctx
        rd = c.empty(3, self.n)
        _check(c.handle, c.lib.rls_ndprofile_eval_profile(c.handle, self.n, C.byref(self._struct),
                                                          _f32(r, "r", self.n, c).data_ptr(),
                                                          abi.vec3(_f32rows(rd, "rd"))), c.lib)
        return rd


class GaussianProfile:
    """`rls::GaussianProfile` (src/rlSss.h:63-97), the alternative `Profile` argument of SssSampler;
    same method names as the reference class.  Single channel: setDistance reads dist.x only."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.state = None

    def setDistance(self, dist, albedo):
        n, c = dist.shape[1], self.ctx
        self.n = n
        self.state = dict(variance=c.empty(n), max_radius=c.empty(n), norm=c.empty(n))
        self._struct = abi.GaussProfileSoA(self.state["variance"].data_ptr(), self.state["max_radius"].data_ptr(),
                                           self.state["norm"].data_ptr())
        _check(c.handle, c.lib.rls_gaussprofile_set_distance(c.handle, n, abi.vec3(_f32rows(dist, "dist", n, c)),
                                                             abi.vec3(_f32rows(albedo, "albedo", n, c)),
                                                             C.byref(self._struct)), c.lib)
        return self.state

    def maxRadius(self):
        return self.state["max_radius"]

    def _unary(self, fn, x, name):
        c = self.ctx
        out = c.empty(self.n)
        _check(c.handle, fn(c.handle, self.n, C.byref(self._struct), _f32(x, name, self.n, c).data_ptr(),
                            out.data_ptr()), c.lib)
        return out

    def getRadius(self, rx):
        return self._unary(self.ctx.lib.rls_gaussprofile_get_radius, rx, "rx")

    def getPdf(self, r):
        return self._unary(self.ctx.lib.rls_gaussprofile_get_pdf, r, "r")

    def evalProfile(self, r):
        return self._unary(self.ctx.lib.rls_gaussprofile_eval_profile, r, "r")

    @staticmethod
    def sampleEvalPdf(ctx, dist_x, rx):
        """Fused unit: setDistance + getRadius + getPdf + evalProfile, one launch."""
        n = rx.shape[0]
        out = dict(r=ctx.empty(n), pdf=ctx.empty(n), Rd=ctx.empty(n))
        _check(ctx.handle, ctx.lib.rls_gaussprofile_sample_eval_pdf(
            ctx.handle, n, _f32(dist_x, "dist_x", n, ctx).data_ptr(), _f32(rx, "rx", n, ctx).data_ptr(),
            out["r"].data_ptr(), out["pdf"].data_ptr(), out["Rd"].data_ptr()), ctx.lib)
        return out


class SkinProfile:
    """The rlSkin diffusion-profile unit (src/rlSkin.cpp:234-241 + NDProfile): node
    parameters by name (src/rlSkin.cpp:109-131)."""

    def __init__(self, ctx, n, **node_params):
        self.ctx, self.n = ctx, n
        self._node_params = node_params
        self.params = abi.skin_params(**node_params)

    def alloc_out(self, like):
        c, n = self.ctx, self.n
        return dict(r=c.empty(n, like=like), pdf=c.empty(n, like=like), Rd=c.empty(3, n, like=like),
                    flags=c.empty(n, dtype=torch.int32, like=like))

    def sampleEvalPdf(self, rx, out=None, chunk=0):
        c, n = self.ctx, self.n
        out = self.alloc_out(rx) if out is None else out
        h = rx.device.type == "cpu"
        _params_placed(self._node_params, n, c, h)
        o = abi.ProfileOut(_f32(out["r"], "r", n, c, h).data_ptr(), _f32(out["pdf"], "pdf", n, c, h).data_ptr(),
                           abi.vec3(_f32rows(out["Rd"], "Rd", n, c, h)), _i32(out["flags"], "flags", n, c, h).data_ptr())
        args = (c.handle, n, C.byref(self.params), _f32(rx, "rx", n, c, h).data_ptr(), C.byref(o))
        if h:
            _check(c.handle, c.lib.rls_skin_profile_sample_eval_pdf_host(*args, chunk), c.lib)
        else:
            _check(c.handle, c.lib.rls_skin_profile_sample_eval_pdf(*args), c.lib)
        return out

    def getProbeRay(self, sg, rx, ry):
        """SssSampler::getProbeRay (src/rlSss.h:487-533) for one (rx, ry) per sample; origins are
        relative to the shading point."""
        c, n = self.ctx, self.n
        out = dict(r=c.empty(n), origin=c.empty(3, n), dir=c.empty(3, n), maxdist=c.empty(n),
                   flags=c.empty(n, dtype=torch.int32))
        o = abi.ProbeOut(out["r"].data_ptr(), abi.vec3(_f32rows(out["origin"], "origin")),
                         abi.vec3(_f32rows(out["dir"], "dir")), out["maxdist"].data_ptr(), out["flags"].data_ptr())
        if sg.n != n:
            raise ValueError(f"ShadingBatch has {sg.n} samples, the profile {n}")
        sg.placed(c)
        _params_placed(self._node_params, n, c, False)
        _check(c.handle, c.lib.rls_skin_probe_ray(c.handle, n, C.byref(sg.struct), C.byref(self.params),
                                                  _f32(rx, "rx", n, c).data_ptr(), _f32(ry, "ry", n, c).data_ptr(),
                                                  C.byref(o)), c.lib)
        return out

    def probeMisPdf(self, sg, disp, hit_normal):
        """The 3-axis MIS pdf of a probe hit (src/rlSss.h:252-263)."""
        c, n = self.ctx, self.n
        pdf = c.empty(n)
        if sg.n != n:
            raise ValueError(f"ShadingBatch has {sg.n} samples, the profile {n}")
        sg.placed(c)
        _params_placed(self._node_params, n, c, False)
        _check(c.handle, c.lib.rls_skin_probe_mis_pdf(c.handle, n, C.byref(sg.struct), C.byref(self.params),
                                                      abi.vec3(_f32rows(disp, "disp", n, c)),
                                                      abi.vec3(_f32rows(hit_normal, "hit_normal", n, c)),
                                                      pdf.data_ptr()), c.lib)
        return pdf

    def glossyLayers(self, sg, k, rx_sheen, ry_sheen, rx_specular, ry_specular, li_sheen=None, li_specular=None):
        """rlSkin's sheen + specular layers with K BRDF samples per shading point and the
        average-Fresnel hand-off to the SSS weight (src/rlSkin.cpp:184-238).  Sample arrays are
        sample-major [k * n]; li_* are optional [3, k * n] incoming radiances."""
        c, n = self.ctx, self.n
        if int(k) < 1 or sg.n != n:
            raise ValueError("glossyLayers: k must be at least 1 and the ShadingBatch must hold one entry per shading point")
        sg.placed(c)
        _params_placed(self._node_params, n, c, False)
        for nm, t in (("rx_sheen", rx_sheen), ("ry_sheen", ry_sheen), ("rx_specular", rx_specular), ("ry_specular", ry_specular)):
            _f32(t, nm, n * k, c)
        out = dict(sheen=c.empty(3, n), specular=c.empty(3, n), sheen_fresnel=c.empty(n), specular_fresnel=c.empty(n),
                   sss_weight=c.empty(n), flags=c.empty(n, dtype=torch.int32))
        o = abi.SkinLayersOut(abi.vec3(_f32rows(out["sheen"], "sheen")), abi.vec3(_f32rows(out["specular"], "specular")),
                              out["sheen_fresnel"].data_ptr(), out["specular_fresnel"].data_ptr(),
                              out["sss_weight"].data_ptr(), out["flags"].data_ptr())
        la = abi.vec3(_f32rows(li_sheen, "li_sheen", n * k, c)) if li_sheen is not None else abi.vec3(None)
        lb = abi.vec3(_f32rows(li_specular, "li_specular", n * k, c)) if li_specular is not None else abi.vec3(None)
        _check(c.handle, c.lib.rls_skin_glossy_layers(c.handle, n, k, C.byref(sg.struct), C.byref(self.params),
                                                      rx_sheen.data_ptr(), ry_sheen.data_ptr(), rx_specular.data_ptr(),
                                                      ry_specular.data_ptr(), la, lb, C.byref(o)), c.lib)
        return out

    def layerWeights(self, avg_fresnel_sheen, avg_fresnel_specular):
        c, n = self.ctx, self.n
        a, b = c.empty(n), c.empty(n)
        _params_placed(self._node_params, n, c, False)
        _check(c.handle, c.lib.rls_skin_layer_weights(c.handle, n, C.byref(self.params),
                                                      _f32(avg_fresnel_sheen, "avg_fresnel_sheen", n, c).data_ptr(),
                                                      _f32(avg_fresnel_specular, "avg_fresnel_specular", n, c).data_ptr(),
                                                      a.data_ptr(), b.data_ptr()), c.lib)
        return a, b
