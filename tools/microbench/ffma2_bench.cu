// Microbenchmark: issue rate of scalar FP32 ops vs packed f32x2 ops on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ffma2_bench ffma2_bench.cu
// NOTE (ptxas 12.9): a mul.rn.f32x2 whose result feeds an add.rn.f32x2 / sub.rn.f32x2 is
// CONTRACTED into one FFMA2 even with --fmad=false (tools/microbench/f32x2_exact_test.cu shows
// the result is the fused one).  The packed modes below therefore keep multiplies and adds on
// SEPARATE accumulators so that the SASS really contains FMUL2 / FADD2 (checked with cuobjdump).
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

__device__ __forceinline__ unsigned long long pk(float a, float b)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float sum2(unsigned long long v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }

template <int MODE> __global__ void k(float *out, float seed)
{
    float s = seed + threadIdx.x * 1e-7f;
    float r = 0;
    if (MODE == 0) {            // scalar fma.rn
        float a[ILP];
        for (int j = 0; j < ILP; j++) a[j] = s + j;
        for (int i = 0; i < ITERS; i++)
#pragma unroll
            for (int j = 0; j < ILP; j++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[j]) : "f"(s), "f"(seed));
        for (int j = 0; j < ILP; j++) r += a[j];
    } else if (MODE == 1) {     // packed fma.rn.f32x2
        unsigned long long a[ILP], b = pk(s, s), c = pk(seed, seed);
        for (int j = 0; j < ILP; j++) a[j] = pk(s + j, s - j);
        for (int i = 0; i < ITERS; i++)
#pragma unroll
            for (int j = 0; j < ILP; j++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[j]) : "l"(b), "l"(c));
        for (int j = 0; j < ILP; j++) r += sum2(a[j]);
    } else if (MODE == 2) {     // scalar mul.rn / add.rn on separate accumulators
        float a[ILP];
        for (int j = 0; j < ILP; j++) a[j] = s + j;
        for (int i = 0; i < ITERS; i++)
#pragma unroll
            for (int j = 0; j < ILP; j += 2) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[j]) : "f"(s)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[j + 1]) : "f"(seed)); }
        for (int j = 0; j < ILP; j++) r += a[j];
    } else if (MODE == 3) {     // packed mul.rn.f32x2 / add.rn.f32x2 on separate accumulators
        unsigned long long a[ILP], b = pk(s, s), c = pk(seed, seed);
        for (int j = 0; j < ILP; j++) a[j] = pk(s + j, s - j);
        for (int i = 0; i < ITERS; i++)
#pragma unroll
            for (int j = 0; j < ILP; j += 2) { asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(a[j]) : "l"(b)); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[j + 1]) : "l"(c)); }
        for (int j = 0; j < ILP; j++) r += sum2(a[j]);
    } else if (MODE == 4) {     // packed mul.rn.f32x2 only
        unsigned long long a[ILP], b = pk(s, s);
        for (int j = 0; j < ILP; j++) a[j] = pk(s + j, s - j);
        for (int i = 0; i < ITERS; i++)
#pragma unroll
            for (int j = 0; j < ILP; j++) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(a[j]) : "l"(b));
        for (int j = 0; j < ILP; j++) r += sum2(a[j]);
    } else {                    // packed add.rn.f32x2 only
        unsigned long long a[ILP], c = pk(seed, seed);
        for (int j = 0; j < ILP; j++) a[j] = pk(s + j, s - j);
        for (int i = 0; i < ITERS; i++)
#pragma unroll
            for (int j = 0; j < ILP; j++) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[j]) : "l"(c));
        for (int j = 0; j < ILP; j++) r += sum2(a[j]);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE> void run(const char *name, int lanes, int flops_per_lane_op)
{
    float *out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(out, 1.0f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(out, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warp_inst = 148.0 * 8 * (256 / 32) * (double)ITERS * ILP;
    double flops = warp_inst * 32 * lanes * flops_per_lane_op;
    printf("%-44s %7.3f ms  %8.1f G warp-inst/s  %7.2f TFLOP/s\n", name, ms, warp_inst / ms * 1e-6, flops / ms * 1e-9);
    cudaFree(out);
}

int main()
{
    run<0>("fma.rn.f32                          (FFMA)", 1, 2);
    run<1>("fma.rn.f32x2                        (FFMA2)", 2, 2);
    run<2>("mul.rn.f32 | add.rn.f32             (FMUL, FADD)", 1, 1);
    run<3>("mul.rn.f32x2 | add.rn.f32x2         (FMUL2, FADD2)", 2, 1);
    run<4>("mul.rn.f32x2                        (FMUL2)", 2, 1);
    run<5>("add.rn.f32x2                        (FADD2)", 2, 1);
    return 0;
}
