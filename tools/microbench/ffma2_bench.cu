// Microbenchmark: issue/throughput of scalar FP32 ops vs packed f32x2 ops on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
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
__device__ __forceinline__ float lo(unsigned long long v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }

template <int MODE> __global__ void k(float *out, float seed)
{
    float s = seed + threadIdx.x * 1e-7f;
    if (MODE == 0) {            // scalar fma.rn
        float a[ILP];
        for (int j = 0; j < ILP; j++) a[j] = s + j;
        for (int i = 0; i < ITERS; i++)
#pragma unroll
            for (int j = 0; j < ILP; j++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[j]) : "f"(s), "f"(seed));
        float r = 0; for (int j = 0; j < ILP; j++) r += a[j];
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    } else if (MODE == 1) {     // packed fma.rn.f32x2
        unsigned long long a[ILP], b = pk(s, s), c = pk(seed, seed);
        for (int j = 0; j < ILP; j++) a[j] = pk(s + j, s - j);
        for (int i = 0; i < ITERS; i++)
#pragma unroll
            for (int j = 0; j < ILP; j++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[j]) : "l"(b), "l"(c));
        float r = 0; for (int j = 0; j < ILP; j++) r += lo(a[j]);
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    } else if (MODE == 2) {     // scalar mul.rn + add.rn alternating
        float a[ILP];
        for (int j = 0; j < ILP; j++) a[j] = s + j;
        for (int i = 0; i < ITERS; i++)
#pragma unroll
            for (int j = 0; j < ILP; j++) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[j]) : "f"(s)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[j]) : "f"(seed)); }
        float r = 0; for (int j = 0; j < ILP; j++) r += a[j];
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    } else {                    // packed mul.rn.f32x2 + add.rn.f32x2 alternating
        unsigned long long a[ILP], b = pk(s, s), c = pk(seed, seed);
        for (int j = 0; j < ILP; j++) a[j] = pk(s + j, s - j);
        for (int i = 0; i < ITERS; i++)
#pragma unroll
            for (int j = 0; j < ILP; j++) { asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(a[j]) : "l"(b)); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[j]) : "l"(c)); }
        float r = 0; for (int j = 0; j < ILP; j++) r += lo(a[j]);
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    }
}

template <int MODE> void run(const char *name, int ops_per_iter, int flops_per_op)
{
    float *out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(out, 1.0f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 10; r++) k<MODE><<<148 * 8, 256>>>(out, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    double winst = (double)148 * 8 * 8 * ITERS * ILP * ops_per_iter;       // warp instructions
    printf("%-28s %8.3f ms  %7.1f G warp-inst/s  %7.2f TFLOP/s\n", name, ms, winst / ms / 1e6,
           winst * 32 * flops_per_op / ms / 1e9);
    cudaFree(out);
}

int main()
{
    run<0>("fma.rn.f32", 1, 2);
    run<1>("fma.rn.f32x2", 1, 4);
    run<2>("mul.rn.f32 + add.rn.f32", 2, 1);
    run<3>("mul.rn.f32x2 + add.rn.f32x2", 2, 2);
    return 0;
}
