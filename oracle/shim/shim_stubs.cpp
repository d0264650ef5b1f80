/*
 * oracle/shim/shim_stubs.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Bodies for the few Arnold services the BSDF hot path really uses (frame build,
 * parameter evaluation) plus aborting stubs for every renderer service the node-glue
 * halves of the reference TUs name but the oracle never calls.
 */
#include <cstdio>
#include <cstdlib>
#include <cstdarg>
#include <ai.h>

static thread_local const AtVector *tlsFrameU = nullptr;
static thread_local const AtVector *tlsFrameV = nullptr;
static thread_local const float    *tlsParamTable = nullptr;   /* [64][3] */

void rls_shim_set_frame(const AtVector *u, const AtVector *v) { tlsFrameU = u; tlsFrameV = v; }
void rls_shim_clear_frame() { tlsFrameU = tlsFrameV = nullptr; }
void rls_shim_set_param_table(const float *t) { tlsParamTable = t; }

/* Documented fallback when no explicit frame was supplied: polar construction around
 * the normal (u along increasing azimuth, v = n x u). Arnold's own body is proprietary. */
void AiBuildLocalFramePolar(AtVector *u, AtVector *v, const AtVector *n)
{
    if (tlsFrameU && tlsFrameV) {
        *u = *tlsFrameU;
        *v = *tlsFrameV;
        return;
    }
    if (n->x == 0.0f && n->y == 0.0f) {
        AiV3Create(*u, 1.0f, 0.0f, 0.0f);
    } else {
        AtVector t;
        AiV3Create(t, -n->y, n->x, 0.0f);
        *u = AiV3Normalize(t);
    }
    *v = AiV3Cross(*n, *u);
}
void AiBuildLocalFrameShirley(AtVector *u, AtVector *v, const AtVector *n) { AiBuildLocalFramePolar(u, v, n); }

float AiShaderEvalParamFuncFlt(AtShaderGlobals *, const AtNode *, int p) { return tlsParamTable[p * 3]; }
AtRGB AiShaderEvalParamFuncRGB(AtShaderGlobals *, const AtNode *, int p)
{
    return rls_shim_rgb(tlsParamTable[p * 3], tlsParamTable[p * 3 + 1], tlsParamTable[p * 3 + 2]);
}
AtVector AiShaderEvalParamFuncVec(AtShaderGlobals *, const AtNode *, int p)
{
    return rls_shim_v3(tlsParamTable[p * 3], tlsParamTable[p * 3 + 1], tlsParamTable[p * 3 + 2]);
}

[[noreturn]] static void unreachable_service(const char *name)
{
    std::fprintf(stderr, "[oracle shim] renderer service %s is outside the BSDF hot path\n", name);
    std::abort();
}
#define STUB(sig, name) sig { unreachable_service(name); }

STUB(const char *AiShaderEvalParamFuncStr(AtShaderGlobals *, const AtNode *, int), "AiShaderEvalParamStr")
void AiNodeParamFlt(AtList *, const char *, float) {}
void AiNodeParamRGB(AtList *, const char *, float, float, float) {}
void AiNodeParamVec(AtList *, const char *, float, float, float) {}
void AiNodeParamStr(AtList *, const char *, const char *) {}
void AiNodeParamBool(AtList *, const char *, bool) {}
void AiMetaDataSetInt(AtMetaDataStore *, const char *, const char *, int) {}
void AiMetaDataSetFlt(AtMetaDataStore *, const char *, const char *, float) {}
void AiMetaDataSetBool(AtMetaDataStore *, const char *, const char *, bool) {}

STUB(AtShaderGlobals *AiShaderGlobals(), "AiShaderGlobals")
STUB(bool AiShaderGlobalsApplyOpacity(AtShaderGlobals *, AtRGB), "AiShaderGlobalsApplyOpacity")
STUB(void *AiShaderGlobalsQuickAlloc(AtShaderGlobals *, size_t), "AiShaderGlobalsQuickAlloc")
STUB(void AiShaderGlobalsSetTraceSet(AtShaderGlobals *, AtString, bool), "AiShaderGlobalsSetTraceSet")
STUB(void AiShaderGlobalsUnsetTraceSet(AtShaderGlobals *), "AiShaderGlobalsUnsetTraceSet")
STUB(AtNode *AiUniverseGetOptions(), "AiUniverseGetOptions")
STUB(int AiNodeGetInt(const AtNode *, const char *), "AiNodeGetInt")
STUB(bool AiNodeGetBool(const AtNode *, AtString), "AiNodeGetBool")
STUB(const char *AiNodeGetStr(const AtNode *, const char *), "AiNodeGetStr")
STUB(const char *AiNodeGetStrAtString(const AtNode *, AtString), "AiNodeGetStrAtString")
STUB(void *AiNodeGetPtr(const AtNode *, const char *), "AiNodeGetPtr")
STUB(AtNode *AiNodeLookUpByName(const char *), "AiNodeLookUpByName")
STUB(const AtNodeEntry *AiNodeGetNodeEntry(const AtNode *), "AiNodeGetNodeEntry")
STUB(void AiNodeSetLocalData(AtNode *, void *), "AiNodeSetLocalData")
STUB(void *AiNodeGetLocalData(const AtNode *), "AiNodeGetLocalData")
STUB(AtSampler *AiSampler(int, int), "AiSampler")
void AiSamplerDestroy(AtSampler *) {}
STUB(AtSamplerIterator *AiSamplerIterator(const AtSampler *, const AtShaderGlobals *), "AiSamplerIterator")
STUB(bool AiSamplerGetSample(AtSamplerIterator *, float *), "AiSamplerGetSample")
STUB(float AiSamplerGetSampleInvCount(const AtSamplerIterator *), "AiSamplerGetSampleInvCount")
STUB(int AiSamplerGetSampleCount(const AtSamplerIterator *), "AiSamplerGetSampleCount")
STUB(void AiLightsPrepare(AtShaderGlobals *), "AiLightsPrepare")
STUB(bool AiLightsGetSample(AtShaderGlobals *), "AiLightsGetSample")
STUB(bool AiLightGetAffectDiffuse(const AtNode *), "AiLightGetAffectDiffuse")
STUB(bool AiLightGetAffectSpecular(const AtNode *), "AiLightGetAffectSpecular")
STUB(float AiLightGetDiffuse(const AtNode *), "AiLightGetDiffuse")
STUB(AtColor AiEvaluateLightSample(AtShaderGlobals *, const void *, AtBRDFEvalSampleFunc, AtBRDFEvalBrdfFunc, AtBRDFEvalPdfFunc), "AiEvaluateLightSample")
STUB(AtColor AiBRDFIntegrate(AtShaderGlobals *, const void *, AtBRDFEvalSampleFunc, AtBRDFEvalBrdfFunc, AtBRDFEvalPdfFunc, AtUInt16), "AiBRDFIntegrate")
STUB(void *AiOrenNayarMISCreateData(const AtShaderGlobals *, float), "AiOrenNayarMISCreateData")
STUB(AtVector AiOrenNayarMISSample(const void *, float, float), "AiOrenNayarMISSample")
STUB(AtColor AiOrenNayarMISBRDF(const void *, const AtVector *), "AiOrenNayarMISBRDF")
STUB(float AiOrenNayarMISPDF(const void *, const AtVector *), "AiOrenNayarMISPDF")
STUB(void AiMakeRay(AtRay *, AtUInt32, const AtPoint *, const AtVector *, double, const AtShaderGlobals *), "AiMakeRay")
STUB(bool AiRefractRay(AtRay *, const AtVector *, float, float, AtShaderGlobals *), "AiRefractRay")
STUB(void AiReflectRay(AtRay *, const AtVector *, const AtShaderGlobals *), "AiReflectRay")
STUB(bool AiTrace(const AtRay *, AtScrSample *), "AiTrace")
STUB(void AiTraceBackground(const AtRay *, AtScrSample *), "AiTraceBackground")
STUB(bool AiTraceProbe(const AtRay *, AtShaderGlobals *), "AiTraceProbe")
STUB(float AiFresnelWeight(AtVector, AtVector, float), "AiFresnelWeight")
STUB(bool AiAOVSetRGB(AtShaderGlobals *, const char *, AtRGB), "AiAOVSetRGB")
STUB(bool AiStateSetMsgInt(const char *, int), "AiStateSetMsgInt")
STUB(bool AiStateGetMsgInt(const char *, int *), "AiStateGetMsgInt")
STUB(bool AiStateSetMsgPtr(const char *, void *), "AiStateSetMsgPtr")
STUB(bool AiStateGetMsgPtr(const char *, void **), "AiStateGetMsgPtr")
void AiMsgInfo(const char *, ...) {}
void AiMsgWarning(const char *, ...) {}
