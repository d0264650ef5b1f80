// rls_libm.cuh -- binary32 transcendentals whose results follow the HOST libm.
//
// The reference's hot path calls sqrtf, sincosf, atan2f, acosf, tanf, powf, logf, expf
// from the host C library (glibc 2.39 on the build image).  Visible-normal sampling is
// ill-conditioned (SURVEY.md 7 "Hard parts"): a 1-ulp difference in one of these results
// moves the sampled direction by more than 1e-6 in a few percent of samples.  CUDA's
// binary32 libdevice versions are 1-4 ulp, so they cannot be used where the value feeds a
// sampled direction or a discrete decision.
//
// v1 policy: evaluate in binary64 (CUDA libdevice, <= 2 ulp in double) and round once to
// binary32.  That gives the correctly rounded binary32 result except with probability
// ~2^-27 per call, and glibc's own binary32 functions are correctly rounded in the large
// majority of calls.
#pragma once
#include "rls_math.cuh"

namespace rlm {

RLS_DEV void sincosf_(float x, float *s, float *c)
{
    double ds, dc;
    sincos((double)x, &ds, &dc);
    *s = (float)ds;
    *c = (float)dc;
}
RLS_DEV float tanf_(float x)   { return (float)tan((double)x); }
RLS_DEV float acosf_(float x)  { return (float)acos((double)x); }
RLS_DEV float atan2f_(float y, float x) { return (float)atan2((double)y, (double)x); }
RLS_DEV float expf_(float x)   { return (float)exp((double)x); }
RLS_DEV float logf_(float x)   { return (float)log((double)x); }
RLS_DEV float powf_(float x, float y) { return (float)pow((double)x, (double)y); }
// powf(x, 5.0f) for x in [0, 1]: three exact-ish binary64 products, one rounding.
RLS_DEV float pow5f_(float x)
{
    double d = (double)x;
    double d2 = d * d;
    return (float)(d2 * d2 * d);
}

} // namespace rlm
