// rls_tol_launch.cuh -- interface between the two translation units of librls_b200.so:
//   rls_tol.cu   the tolerance-policy kernels (rls_tol.cuh), compiled WITH FMA contraction;
//   rls_b200.cu  everything else (bit-exact policies, the exact re-run of the samples the tolerance kernels list,
//                the C ABI), compiled with -fmad=false.
// A tolerance kernel writes every sample's outputs and appends the index of each sample whose band tracker fired
// (rls_tol.cuh Bands) to `list`; rls_b200.cu then launches k_*_rerun over that list with the bit-exact policy.  When the
// list is full the kernel marks the sample instead by storing kRerunSentinel in its flags word, and the re-run kernel
// scans the flags of the whole batch (no sample is ever lost; the list capacity only decides which path is taken).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "rls_kernel_args.cuh"
#include "rls_sweep.cuh"

namespace rls {
namespace tol {

constexpr uint32_t kRerunSentinel = 0xffffffffu;    // not a valid flags word (bits 14-15 and 30-31 are never set)

struct Worklist {
    uint32_t *list;     // device: sample indices to re-run
    unsigned *count;    // device: number appended (may exceed cap), zeroed by the caller on the launch stream
    uint32_t  cap;
};

cudaError_t launch_ggx_sample_eval_pdf(cudaStream_t st, size_t n, const ShadingSoA &sg, const GgxParamsDev &p, const float *rx,
                                       const float *ry, const BsdfOutDev &o, const Worklist &wl);
cudaError_t launch_ggx_dielectric(cudaStream_t st, size_t n, const ShadingSoA &sg, const GgxParamsDev &p, const float *rx,
                                  const float *ry, const DielectricOutDev &o, const Worklist &wl);
cudaError_t launch_disney(cudaStream_t st, size_t n, const ShadingSoA &sg, const DisneyParamsDev &p, bool all_arrays,
                          const float *rx_s, const float *ry_s, const float *rx_d, const float *ry_d, const DisneyOutDev &o,
                          const Worklist &wl);
// The sweep's re-run list: (cell, k) of every band sample.  The tolerance kernel leaves such a sample OUT of its sums and
// k_albedo_sweep_rerun (rls_b200.cu) adds its bit-exact contribution; when the list is full the kernel keeps its own
// tolerance-policy value for the sample instead (nothing is lost, the counts may then differ on a band sample).
struct SweepWorklist { uint2 *list; unsigned *count; uint32_t cap; };
cudaError_t launch_albedo_sweep(cudaStream_t st, const SweepGridDev &g, uint32_t n_cells, uint64_t seed, uint32_t k0, uint32_t k1,
                                double *table, const SweepWorklist &wl);
cudaError_t launch_skin_profile(cudaStream_t st, size_t n, const SkinParamsDev &p, const float *rx, const ProfileOutDev &o,
                                const Worklist &wl);

} // namespace tol
} // namespace rls
