// rls_kernel_args.cuh -- kernel argument blocks (device-side mirrors of the ABI structs of include/rls_b200.h)
// shared by the translation units of the library: rls_b200.cu (bit-exact policies, C ABI) and rls_tol.cu (the
// tolerance policy, compiled with FMA contraction).
#pragma once
#include "rls_math.cuh"
#include "rls_disney.cuh"
#include "rls_profile.cuh"

namespace rls {

struct GgxParamsDev { P3 ks; P1 rough, ior, aniso; int ndf; };
struct BsdfOutDev { V3 wi, f; float *pdf, *fresnel; uint32_t *flags; };
struct DielectricOutDev { float *fresnel; V3 wi_r; float *f_r, *pdf_r; V3 wi_t; float *f_t, *weight_t; uint32_t *flags; };
struct DisneyOutDev { V3 wi_s, f_s; float *pdf_s; V3 wi_d, f_d; float *pdf_d; uint32_t *flags; };
struct ProfileOutDev { float *r, *pdf; V3 Rd; uint32_t *flags; };

} // namespace rls
