// rls_tile.cuh -- SoA tile staging through shared memory with the TMA bulk-copy engine
// (cp.async.bulk, SASS UBLKCP), for the persistent forms of the fused kernels.
//
// Why: the fused kernels are instruction-issue bound (profiles/).  With one LDG / STG per
// component array, 29 memory instructions per rough-dielectric sample drag ~75 further issue slots
// of 64-bit address arithmetic and pointer loads behind them (one LDC.64 + IMAD.WIDE, or
// IADD3 + IADD3.X, per array), and every warp starts by waiting for its loads.  Here ONE warp per
// CTA hands whole tiles to the copy engine: kTile consecutive samples of every input array are
// copied global -> shared with one bulk copy per array (completion on an mbarrier), the threads
// read their sample with LDS at immediate offsets from one base register, write their results
// with STS, and one bulk copy per output array drains the tile shared -> global.  The CTA is
// persistent (grid = resident CTAs, a multiple of the SM count): the loads of its NEXT tile are
// issued as soon as every thread has its inputs in registers, a whole tile's compute ahead.
//
// Requirements checked at the launch site: every array pointer 16-byte aligned (cp.async.bulk);
// the kernel processes whole tiles, the launch site sends the ragged tail to the plain kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rls {
namespace tile {

constexpr int kTile = 256;                  // samples per tile == threads per CTA
constexpr int kSlotBytes = kTile * 4;       // one component array of one tile
constexpr int kMaxIn = 32, kMaxOut = 16;

// Per-launch description of the arrays (a __grid_constant__ kernel parameter: lane l of the
// copy warp reads entry l through the constant bank).  A null entry is an absent array (uniform
// parameter / optional output).  elem = bytes per sample (4, or 1 for the backfacing bytes).
struct Arrays {
    const void *in[kMaxIn];
    void *out[kMaxOut];
    uint8_t in_elem[kMaxIn];
    uint32_t in_bytes_per_tile;             // sum over present inputs of kTile * elem
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// The per-CTA pipeline.  Shared memory: kIn input slots, then kOut output slots, of kSlotBytes
// (slot numbers are compile-time constants, so that LDS / STS take immediate offsets).
// Protocol per tile (all threads call every method, in this order):
//   wait_inputs();  <LDS own sample>  inputs_consumed(next_tile);
//   <compute>
//   begin_store();  <STS own results>  end_store(tile);
template <int kIn, int kOut>
struct Pipe {
    static constexpr int kSmemBytes = (kIn + kOut) * kSlotBytes;
    const Arrays *a;
    unsigned char *sin, *sout;
    uint64_t *bar;
    uint32_t parity;

    __device__ __forceinline__ void init(const Arrays *arrays, unsigned char *smem, uint64_t *mbar)
    {
        a = arrays; sin = smem; sout = smem + kIn * kSlotBytes; bar = mbar; parity = 0;
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
    }
    // copy warp: one bulk copy per present input array of tile `t`
    __device__ __forceinline__ void issue_loads(uint32_t t)
    {
        if (threadIdx.x < 32) {
            if (threadIdx.x == 0) mbar_expect_tx(bar, a->in_bytes_per_tile);
            __syncwarp();
            const int l = threadIdx.x;
            if (l < kIn) {
                const char *src = (const char *)a->in[l];
                const uint32_t bytes = kTile * a->in_elem[l];
                if (src) bulk_g2s(sin + l * kSlotBytes, src + (size_t)t * bytes, bytes, bar);
            }
        }
    }
    __device__ __forceinline__ void wait_inputs() { mbar_wait(bar, parity); parity ^= 1u; }
    // every thread holds its inputs in registers: the input slots are free for tile `next`
    __device__ __forceinline__ void inputs_consumed(uint32_t next, uint32_t n_tiles)
    {
        __syncthreads();
        if (next < n_tiles) issue_loads(next);
    }
    __device__ __forceinline__ void begin_store()
    {
        if (threadIdx.x < (unsigned)kOut) bulk_wait_read0();     // the previous tile's drains have read sout
        __syncthreads();
    }
    __device__ __forceinline__ void end_store(uint32_t t)
    {
        fence_async_smem();                 // generic-proxy STS visible to the async proxy
        __syncthreads();
        const int l = threadIdx.x;
        if (l < kOut) {
            char *dst = (char *)a->out[l];
            if (dst) bulk_s2g(dst + (size_t)t * kSlotBytes, sout + l * kSlotBytes, kSlotBytes);
            bulk_commit();
        }
    }
    __device__ __forceinline__ void finish()
    {
        if (threadIdx.x < (unsigned)kOut) bulk_wait_all();
    }
    __device__ __forceinline__ const float *in_slot(int k) const { return reinterpret_cast<const float *>(sin + k * kSlotBytes); }
    __device__ __forceinline__ float *out_slot(int k) const { return reinterpret_cast<float *>(sout + k * kSlotBytes); }
};

} // namespace tile
} // namespace rls
