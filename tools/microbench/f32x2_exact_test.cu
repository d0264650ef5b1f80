#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
struct f2p { unsigned long long v; };
__device__ __forceinline__ f2p pack(float a, float b){ f2p r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(f2p p, float &a, float &b){ asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p.v)); }
__device__ __forceinline__ f2p mul2(f2p a, f2p b){ f2p r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2p add2(f2p a, f2p b){ f2p r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2p fma2(f2p a, f2p b, f2p c){ f2p r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ uint32_t hash(uint32_t x){ x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__global__ void k(unsigned long long *mism_unfused, unsigned long long *mism_fused, int iters){
  uint32_t i = blockIdx.x*blockDim.x+threadIdx.x;
  unsigned long long m1=0, m2=0;
  for (int it=0; it<iters; it++){
    uint32_t h = hash(i*977u + it*0x9e3779b9u);
    float a = __uint_as_float(0x3f000000u | (hash(h) & 0x7fffffu));     // [0.5,1)
    float b = __uint_as_float(0x3f000000u | (hash(h+1) & 0x7fffffu));
    float c = -__uint_as_float(0x3e800000u | (hash(h+2) & 0x7fffffu));  // negative: cancellation exposes fusion
    float d = __uint_as_float(0x3f000000u | (hash(h+3) & 0x7fffffu));
    float e = __uint_as_float(0x3f000000u | (hash(h+4) & 0x7fffffu));
    float f = -__uint_as_float(0x3e800000u | (hash(h+5) & 0x7fffffu));
    float r1 = __fadd_rn(__fmul_rn(a,b), c), r2 = __fadd_rn(__fmul_rn(d,e), f);
    float g1 = __fmaf_rn(a,b,c), g2 = __fmaf_rn(d,e,f);
    float p1, p2; unpack(add2(mul2(pack(a,d), pack(b,e)), pack(c,f)), p1, p2);
    float q1, q2; unpack(fma2(pack(a,d), pack(b,e), pack(c,f)), q1, q2);
    m1 += (__float_as_uint(p1) != __float_as_uint(r1)) + (__float_as_uint(p2) != __float_as_uint(r2));
    m2 += (__float_as_uint(q1) != __float_as_uint(g1)) + (__float_as_uint(q2) != __float_as_uint(g2));
    if (it == 0 && i == 0) printf("fused differs from unfused here? %d\n", __float_as_uint(g1) != __float_as_uint(r1));
  }
  atomicAdd(mism_unfused, m1); atomicAdd(mism_fused, m2);
}
int main(){
  unsigned long long *d; cudaMalloc(&d, 16); cudaMemset(d, 0, 16);
  k<<<1024,256>>>(d, d+1, 1000);
  unsigned long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("packed mul2+add2 vs scalar unfused: %llu mismatches; packed fma2 vs scalar fma: %llu mismatches (of %llu)\n", h[0], h[1], 2ull*1024*256*1000);
  return 0;
}
