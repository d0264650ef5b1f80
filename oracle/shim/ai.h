/*
 * oracle/shim/ai.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A minimal stand-in for the proprietary Arnold 4.2 SDK header <ai.h>, just big
 * enough that the reference sources under /root/reference/src compile UNMODIFIED
 * (see oracle/Makefile).  Only the inline math/types the BSDF hot path touches have
 * real bodies; every renderer service (lights, tracing, samplers, AOVs, node
 * registry) is a declaration whose body in shim_stubs.cpp aborts -- the oracle never
 * reaches them.
 *
 * Because the Arnold SDK cannot be consulted offline, the definitions below are
 * NORMATIVE for this repository's parity claims.  Each open choice:
 *
 *   AI_EPSILON = 1e-4f, AI_BIG = 1e12f, AI_PI... = float literals.
 *   SQR/ABS/MIN/MAX/CLAMP  obvious macros; MAX(a,b) = (a > b) ? a : b.
 *   LERP(t,a,b)            = (1-t)*a + b*t   (usage LERP(FH, F0, white), rlDisney.cpp:341)
 *   LINEARSTEP(lo,hi,t)    = clamp((t-lo)/(hi-lo), 0, 1)   (usage rlSss.h:33)
 *   SGN(x)                 = int: -1, 0, +1 (three-way; parity inputs avoid exact 0)
 *   AiV3Dot                = a.x*b.x + a.y*b.y + a.z*b.z   (left to right, no FMA)
 *   AiV3Length             = sqrtf(dot(a,a))
 *   AiV3Normalize          = a * (1/len) (reciprocal then 3 multiplies); zero stays zero
 *   AiV3RotateToFrame(a,u,v,w): a = a.x*u + a.y*v + a.z*w, per component left to right
 *   AiV3IsZero / AiColorIsZero = exact == 0 on every component
 *   AiColorIsSmall         = all |c| < AI_EPSILON
 *   AiBuildLocalFramePolar = proprietary.  The harness always supplies explicit frames
 *                            (rls_shim_set_frame); without one a documented polar
 *                            construction is used.
 *   AiM4Frame / AiM4VectorByMatrixMult: columns of the 3x3 block are (u,v,w) so that
 *                            vector-by-matrix maps WORLD -> LOCAL, as rlSss.h:160,253
 *                            assumes ("mWorldToLocalMat").
 *   vector/colour operators: component-wise, scalar on either side, one IEEE op each.
 */
#ifndef RLS_ORACLE_SHIM_AI_H
#define RLS_ORACLE_SHIM_AI_H

#include <cmath>
#include <math.h>   /* libstdc++ wrapper: brings the float overloads of exp/log/pow/sqrt into the global
                       namespace, as <math.h> does on the author's MSVC toolchain, so the unsuffixed
                       calls in rlSss.cpp/rlDisney.cpp stay binary32 */
#include <cstdint>
#include <cstddef>
#include <cstring>

/* ---------------------------------------------------------------- constants */
#define AI_VERSION "4.2.11.0-shim"
#define AI_PI          3.14159265358979323846f
#define AI_PITIMES2    6.28318530717958647692f
#define AI_PIOVER2     1.57079632679489661923f
#define AI_ONEOVERPI   0.31830988618379067154f
#define AI_ONEOVER2PI  0.15915494309189533577f
#define AI_EPSILON     1.0e-4f
#define AI_BIG         1.0e12f

#define AI_RAY_UNDEFINED  0x00
#define AI_RAY_CAMERA     0x01
#define AI_RAY_SHADOW     0x02
#define AI_RAY_REFLECTED  0x04
#define AI_RAY_REFRACTED  0x08
#define AI_RAY_SUBSURFACE 0x10
#define AI_RAY_DIFFUSE    0x20
#define AI_RAY_GLOSSY     0x40

#define AI_TYPE_RGB     5
#define AI_NODE_SHADER  0x0010

/* ------------------------------------------------------------------- macros */
#ifdef MIN
#undef MIN
#endif
#ifdef MAX
#undef MAX
#endif
#define SQR(a)        ((a) * (a))
#define ABS(a)        (((a) < 0) ? -(a) : (a))
#define MIN(a, b)     (((a) < (b)) ? (a) : (b))
#define MAX(a, b)     (((a) > (b)) ? (a) : (b))
#define CLAMP(v, lo, hi) (((v) < (lo)) ? (lo) : (((v) > (hi)) ? (hi) : (v)))
#define SGN(a)        (((a) < 0) ? -1 : (((a) > 0) ? 1 : 0))

typedef uint8_t  AtByte;
typedef uint16_t AtUInt16;
typedef uint32_t AtUInt32;

/* -------------------------------------------------------------------- types */
struct AtVector2 { float x, y; };

struct AtVector {
    float x, y, z;
    float       &operator[](unsigned i)       { return (&x)[i]; }
    const float &operator[](unsigned i) const { return (&x)[i]; }
    AtVector &operator*=(float s)             { x *= s; y *= s; z *= s; return *this; }
    AtVector &operator*=(const AtVector &o)   { x *= o.x; y *= o.y; z *= o.z; return *this; }
    AtVector &operator+=(const AtVector &o)   { x += o.x; y += o.y; z += o.z; return *this; }
    AtVector &operator-=(const AtVector &o)   { x -= o.x; y -= o.y; z -= o.z; return *this; }
};
typedef AtVector AtPoint;

inline AtVector rls_shim_v3(float x, float y, float z) { AtVector v; v.x = x; v.y = y; v.z = z; return v; }
inline AtVector operator+(const AtVector &a, const AtVector &b) { return rls_shim_v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline AtVector operator-(const AtVector &a, const AtVector &b) { return rls_shim_v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline AtVector operator-(const AtVector &a)                    { return rls_shim_v3(-a.x, -a.y, -a.z); }
inline AtVector operator*(const AtVector &a, float s)           { return rls_shim_v3(a.x * s, a.y * s, a.z * s); }
inline AtVector operator*(float s, const AtVector &a)           { return rls_shim_v3(a.x * s, a.y * s, a.z * s); }
inline AtVector operator*(const AtVector &a, const AtVector &b) { return rls_shim_v3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline AtVector operator/(const AtVector &a, float s)           { return rls_shim_v3(a.x / s, a.y / s, a.z / s); }
inline bool operator==(const AtVector &a, const AtVector &b)    { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline bool operator!=(const AtVector &a, const AtVector &b)    { return !(a == b); }

struct AtColor {
    float r, g, b;
    float       &operator[](unsigned i)       { return (&r)[i]; }
    const float &operator[](unsigned i) const { return (&r)[i]; }
    AtColor &operator*=(float s)            { r *= s; g *= s; b *= s; return *this; }
    AtColor &operator*=(const AtColor &o)   { r *= o.r; g *= o.g; b *= o.b; return *this; }
    AtColor &operator+=(const AtColor &o)   { r += o.r; g += o.g; b += o.b; return *this; }
    AtColor &operator-=(const AtColor &o)   { r -= o.r; g -= o.g; b -= o.b; return *this; }
};
typedef AtColor AtRGB;

inline AtColor rls_shim_rgb(float r, float g, float b) { AtColor c; c.r = r; c.g = g; c.b = b; return c; }
inline AtColor operator+(const AtColor &a, const AtColor &b) { return rls_shim_rgb(a.r + b.r, a.g + b.g, a.b + b.b); }
inline AtColor operator+(const AtColor &a, float s)          { return rls_shim_rgb(a.r + s, a.g + s, a.b + s); }
inline AtColor operator+(float s, const AtColor &a)          { return rls_shim_rgb(a.r + s, a.g + s, a.b + s); }
inline AtColor operator-(const AtColor &a, const AtColor &b) { return rls_shim_rgb(a.r - b.r, a.g - b.g, a.b - b.b); }
inline AtColor operator*(const AtColor &a, float s)          { return rls_shim_rgb(a.r * s, a.g * s, a.b * s); }
inline AtColor operator*(float s, const AtColor &a)          { return rls_shim_rgb(a.r * s, a.g * s, a.b * s); }
inline AtColor operator*(const AtColor &a, const AtColor &b) { return rls_shim_rgb(a.r * b.r, a.g * b.g, a.b * b.b); }
inline AtColor operator/(const AtColor &a, float s)          { return rls_shim_rgb(a.r / s, a.g / s, a.b / s); }
inline bool operator==(const AtColor &a, const AtColor &b)   { return a.r == b.r && a.g == b.g && a.b == b.b; }

static const AtVector AI_V3_ZERO   = { 0.0f, 0.0f, 0.0f };
static const AtColor  AI_RGB_BLACK = { 0.0f, 0.0f, 0.0f };
static const AtColor  AI_RGB_WHITE = { 1.0f, 1.0f, 1.0f };
static const AtColor  AI_RGB_RED   = { 1.0f, 0.0f, 0.0f };
static const AtColor  AI_RGB_GREEN = { 0.0f, 1.0f, 0.0f };

typedef float AtMatrix[4][4];

/* LERP / LINEARSTEP as overload sets so they work on floats and colours alike. */
template <typename T> inline T LERP(float t, const T &a, const T &b) { return (1.0f - t) * a + b * t; }
inline float LINEARSTEP(float lo, float hi, float t) { float v = (t - lo) / (hi - lo); return CLAMP(v, 0.0f, 1.0f); }

inline float fast_exp(float x) { return expf(x); }
inline bool  AiIsFinite(float x) { return std::isfinite(x); }

/* ------------------------------------------------------------- vector maths */
inline void  AiV3Create(AtVector &v, float x, float y, float z) { v.x = x; v.y = y; v.z = z; }
inline float AiV3Dot(const AtVector &a, const AtVector &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float AiV3Length(const AtVector &a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
inline float AiV3Dist(const AtVector &a, const AtVector &b) { return AiV3Length(a - b); }
inline AtVector AiV3Normalize(const AtVector &a)
{
    float len = AiV3Length(a);
    if (len != 0.0f) {
        float inv = 1.0f / len;
        return rls_shim_v3(a.x * inv, a.y * inv, a.z * inv);
    }
    return a;
}
inline AtVector AiV3Cross(const AtVector &a, const AtVector &b)
{
    return rls_shim_v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline bool AiV3IsZero(const AtVector &a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }
inline bool AiV3isZero(const AtVector &a) { return AiV3IsZero(a); }
inline bool AiV3Exists(const AtVector &a) { return std::isfinite(a.x) && std::isfinite(a.y) && std::isfinite(a.z); }
inline void AiV3RotateToFrame(AtVector &a, const AtVector &u, const AtVector &v, const AtVector &w)
{
    float x = a.x * u.x + a.y * v.x + a.z * w.x;
    float y = a.x * u.y + a.y * v.y + a.z * w.y;
    float z = a.x * u.z + a.y * v.z + a.z * w.z;
    a.x = x; a.y = y; a.z = z;
}
inline bool AiColorIsZero(const AtColor &c)  { return c.r == 0.0f && c.g == 0.0f && c.b == 0.0f; }
inline bool AiColorIsSmall(const AtColor &c) { return ABS(c.r) < AI_EPSILON && ABS(c.g) < AI_EPSILON && ABS(c.b) < AI_EPSILON; }
inline AtColor AiColorClamp(const AtColor &c, float lo, float hi)
{
    return rls_shim_rgb(CLAMP(c.r, lo, hi), CLAMP(c.g, lo, hi), CLAMP(c.b, lo, hi));
}

inline void AiM4Frame(AtMatrix m, const AtPoint *o, const AtVector *u, const AtVector *v, const AtVector *w)
{
    for (int i = 0; i < 3; i++) {
        m[i][0] = (*u)[i]; m[i][1] = (*v)[i]; m[i][2] = (*w)[i]; m[i][3] = 0.0f;
    }
    m[3][0] = -AiV3Dot(*o, *u); m[3][1] = -AiV3Dot(*o, *v); m[3][2] = -AiV3Dot(*o, *w); m[3][3] = 1.0f;
}
inline void AiM4VectorByMatrixMult(AtVector *out, const AtMatrix m, const AtVector *in)
{
    float x = in->x * m[0][0] + in->y * m[1][0] + in->z * m[2][0];
    float y = in->x * m[0][1] + in->y * m[1][1] + in->z * m[2][1];
    float z = in->x * m[0][2] + in->y * m[1][2] + in->z * m[2][2];
    out->x = x; out->y = y; out->z = z;
}

/* Explicit-frame hook: the harness supplies U,V for the next AiBuildLocalFramePolar
 * call(s) on this thread (north_star: "the bench harness feeds kernels explicit
 * shading frames").  Defined in shim_stubs.cpp. */
void rls_shim_set_frame(const AtVector *u, const AtVector *v);
void rls_shim_clear_frame();
void AiBuildLocalFramePolar(AtVector *u, AtVector *v, const AtVector *n);
void AiBuildLocalFrameShirley(AtVector *u, AtVector *v, const AtVector *n);

/* ------------------------------------------------------- renderer data types */
struct AtNode;
struct AtNodeEntry;
struct AtList;
struct AtMetaDataStore;
struct AtSampler;
struct AtSamplerIterator;
struct AtNodeMethods;
union  AtParamValue { float FLT; int INT; void *PTR; };

class AtString {
public:
    AtString() : mStr("") {}
    explicit AtString(const char *s) : mStr(s) {}
    const char *c_str() const { return mStr; }
private:
    const char *mStr;
};

struct AtShaderGlobals {
    AtPoint   P, Po, Ro;
    AtVector  Rd, N, Nf, Ng, Ngf, Ns;
    AtVector  dPdu, dPdv, dPdx, dPdy, dNdx, dNdy;
    double    Rl;
    float     area, bu, bv;
    AtUInt32  fi;
    AtUInt16  Rt;
    AtByte    Rr, Rr_refr, Rr_diff, Rr_gloss;
    bool      fhemi;
    AtNode   *Op, *shader, *Lp;
    AtShaderGlobals *psg;
    union { AtRGB RGB; float FLT; } out;
    AtRGB     out_opacity;
};

struct AtRay {
    AtPoint  origin;
    AtVector dir;
    double   maxdist;
};

struct AtScrSample {
    AtColor color;
    AtColor opacity;
};

struct AtNodeLib {
    int                  node_type;
    AtByte               output_type;
    const char          *name;
    const AtNodeMethods *methods;
    char                 version[64];
};

typedef AtVector (*AtBRDFEvalSampleFunc)(const void *brdf_data, float rx, float ry);
typedef AtColor  (*AtBRDFEvalBrdfFunc)(const void *brdf_data, const AtVector *indir);
typedef float    (*AtBRDFEvalPdfFunc)(const void *brdf_data, const AtVector *indir);

struct AtNodeMethods {
    void (*Parameters)(AtList *, AtMetaDataStore *);
    void (*Initialize)(AtNode *, AtParamValue *);
    void (*Update)(AtNode *, AtParamValue *);
    void (*Finish)(AtNode *);
    void (*Evaluate)(AtNode *, AtShaderGlobals *);
};

/* ------------------------------------------------- node / shader boilerplate */
#define AI_SHADER_NODE_EXPORT_METHODS(tag)                                   \
    static void Parameters(AtList *params, AtMetaDataStore *mds);            \
    static void Initialize(AtNode *node, AtParamValue *params);              \
    static void Update(AtNode *node, AtParamValue *params);                  \
    static void Finish(AtNode *node);                                        \
    static void Evaluate(AtNode *node, AtShaderGlobals *sg);                 \
    static AtNodeMethods tag##_methods_impl = { Parameters, Initialize, Update, Finish, Evaluate }; \
    AtNodeMethods *tag = &tag##_methods_impl;

#define node_parameters  static void Parameters(AtList *params, AtMetaDataStore *mds)
#define node_initialize  static void Initialize(AtNode *node, AtParamValue *params)
#define node_update      static void Update(AtNode *node, AtParamValue *params)
#define node_finish      static void Finish(AtNode *node)
#define shader_evaluate  static void Evaluate(AtNode *node, AtShaderGlobals *sg)
#define node_loader      extern "C" bool NodeLoader(int i, AtNodeLib *node)

void AiNodeParamFlt(AtList *, const char *, float);
void AiNodeParamRGB(AtList *, const char *, float, float, float);
void AiNodeParamVec(AtList *, const char *, float, float, float);
void AiNodeParamStr(AtList *, const char *, const char *);
void AiNodeParamBool(AtList *, const char *, bool);
#define AiParameterFLT(n, d)        AiNodeParamFlt(params, n, d)
#define AiParameterFlt(n, d)        AiNodeParamFlt(params, n, d)
#define AiParameterRGB(n, r, g, b)  AiNodeParamRGB(params, n, r, g, b)
#define AiParameterVec(n, x, y, z)  AiNodeParamVec(params, n, x, y, z)
#define AiParameterSTR(n, d)        AiNodeParamStr(params, n, d)
#define AiParameterBool(n, d)       AiNodeParamBool(params, n, d)
void AiMetaDataSetInt(AtMetaDataStore *, const char *, const char *, int);
void AiMetaDataSetFlt(AtMetaDataStore *, const char *, const char *, float);
void AiMetaDataSetBool(AtMetaDataStore *, const char *, const char *, bool);

/* Parameter evaluation: the oracle driver installs a per-thread table indexed by the
 * reference's own parameter enum (rlDisney.cpp:24-46). */
float       AiShaderEvalParamFuncFlt(AtShaderGlobals *, const AtNode *, int);
AtRGB       AiShaderEvalParamFuncRGB(AtShaderGlobals *, const AtNode *, int);
AtVector    AiShaderEvalParamFuncVec(AtShaderGlobals *, const AtNode *, int);
const char *AiShaderEvalParamFuncStr(AtShaderGlobals *, const AtNode *, int);
#define AiShaderEvalParamFlt(p) AiShaderEvalParamFuncFlt(sg, node, p)
#define AiShaderEvalParamRGB(p) AiShaderEvalParamFuncRGB(sg, node, p)
#define AiShaderEvalParamVec(p) AiShaderEvalParamFuncVec(sg, node, p)
#define AiShaderEvalParamStr(p) AiShaderEvalParamFuncStr(sg, node, p)
void rls_shim_set_param_table(const float *flt_table /* [64][3] */);

/* ------------------------------ renderer services: declarations, abort stubs */
AtShaderGlobals *AiShaderGlobals();
bool   AiShaderGlobalsApplyOpacity(AtShaderGlobals *, AtRGB);
void  *AiShaderGlobalsQuickAlloc(AtShaderGlobals *, size_t);
void   AiShaderGlobalsSetTraceSet(AtShaderGlobals *, AtString, bool);
void   AiShaderGlobalsUnsetTraceSet(AtShaderGlobals *);
AtNode *AiUniverseGetOptions();
int    AiNodeGetInt(const AtNode *, const char *);
bool   AiNodeGetBool(const AtNode *, AtString);
const char *AiNodeGetStr(const AtNode *, const char *);
const char *AiNodeGetStrAtString(const AtNode *, AtString);
void  *AiNodeGetPtr(const AtNode *, const char *);
AtNode *AiNodeLookUpByName(const char *);
const AtNodeEntry *AiNodeGetNodeEntry(const AtNode *);
void   AiNodeSetLocalData(AtNode *, void *);
void  *AiNodeGetLocalData(const AtNode *);
AtSampler *AiSampler(int, int);
void   AiSamplerDestroy(AtSampler *);
AtSamplerIterator *AiSamplerIterator(const AtSampler *, const AtShaderGlobals *);
bool   AiSamplerGetSample(AtSamplerIterator *, float *);
float  AiSamplerGetSampleInvCount(const AtSamplerIterator *);
int    AiSamplerGetSampleCount(const AtSamplerIterator *);
void   AiLightsPrepare(AtShaderGlobals *);
bool   AiLightsGetSample(AtShaderGlobals *);
bool   AiLightGetAffectDiffuse(const AtNode *);
bool   AiLightGetAffectSpecular(const AtNode *);
float  AiLightGetDiffuse(const AtNode *);
AtColor AiEvaluateLightSample(AtShaderGlobals *, const void *, AtBRDFEvalSampleFunc, AtBRDFEvalBrdfFunc, AtBRDFEvalPdfFunc);
AtColor AiBRDFIntegrate(AtShaderGlobals *, const void *, AtBRDFEvalSampleFunc, AtBRDFEvalBrdfFunc, AtBRDFEvalPdfFunc, AtUInt16);
void  *AiOrenNayarMISCreateData(const AtShaderGlobals *, float);
AtVector AiOrenNayarMISSample(const void *, float, float);
AtColor  AiOrenNayarMISBRDF(const void *, const AtVector *);
float    AiOrenNayarMISPDF(const void *, const AtVector *);
void   AiMakeRay(AtRay *, AtUInt32, const AtPoint *, const AtVector *, double, const AtShaderGlobals *);
bool   AiRefractRay(AtRay *, const AtVector *, float, float, AtShaderGlobals *);
void   AiReflectRay(AtRay *, const AtVector *, const AtShaderGlobals *);
bool   AiTrace(const AtRay *, AtScrSample *);
void   AiTraceBackground(const AtRay *, AtScrSample *);
bool   AiTraceProbe(const AtRay *, AtShaderGlobals *);
float  AiFresnelWeight(AtVector, AtVector, float);
bool   AiAOVSetRGB(AtShaderGlobals *, const char *, AtRGB);
bool   AiStateSetMsgInt(const char *, int);
bool   AiStateGetMsgInt(const char *, int *);
bool   AiStateSetMsgPtr(const char *, void *);
bool   AiStateGetMsgPtr(const char *, void **);
void   AiMsgInfo(const char *, ...);
void   AiMsgWarning(const char *, ...);

#endif /* RLS_ORACLE_SHIM_AI_H */
