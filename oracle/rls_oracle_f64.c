/*
 * oracle/rls_oracle_f64.c -- TEST INFRASTRUCTURE ONLY (kind "port-f64", parity unpinned BY DESIGN: it is not an oracle of
 * the reference's bits but a yardstick for its rounding noise).
 *
 * The plain-C restatement oracle/rls_oracle.c re-typed to binary64: the SAME source text is included below with
 * `float` spelled `double` and every libm call replaced by its binary64 form, so each arithmetic step of the
 * reference's algorithm (paths and lines as cited in rls_oracle.c) is carried with 29 more bits.  The float literals
 * keep their binary32 values (they are the reference's constants).  What it is for: tests/test_tolerant_policy.py and
 * tests/hunts/tol_vs_f64.py measure, sample by sample,
 *      |reference (binary32) - this|      the reference's own rounding noise, and
 *      |RLS_ARITH_TOLERANT   - this|      the tolerance policy's error,
 * which is how the repository states what "within tolerance of the reference" can mean for an ill-conditioned sampler.
 *
 * ABI: the structs of include/rls_b200.h keep their binary32 arrays (inputs are read, outputs rounded, as float);
 * BARE `float *` parameters of the oracle_* functions become `double *` here, and every exported name gets the prefix
 * f64_ (oracle/_f64_rename.h, generated from oracle_api.h by the Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include "oracle_api.h"            /* ABI structs and prototypes with the real `float` */
#include "_f64_rename.h"
#define float   double
#define sqrtf   sqrt
#define sinf    sin
#define cosf    cos
#define tanf    tan
#define atan2f  atan2
#define atanf   atan
#define acosf   acos
#define asinf   asin
#define powf    pow
#define logf    log
#define expf    exp
#define fabsf   fabs
#define floorf  floor
#define fmaxf   fmax
#define fminf   fmin
#include "rls_oracle.c"
