// rxm_device.cuh -- sm_100a device building blocks for the patch-local query engine.
//
// B200-native counterparts of the reference's device helpers
// (/root/reference/include/rxmesh/kernels/loader.cuh:16-46 load_async,
//  kernels/collective.cuh:12-54 cub_block_exclusive_sum,
//  kernels/rxmesh_queries.cuh:16-107 block_mat_transpose,
//  iterator.cuh:117-143 Iterator::operator[]):
//   * patch sections arrive in shared memory through TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx) issued by ONE thread, instead of
//     per-thread LDGSTS;
//   * the block-wide exclusive sum is a hand-written shuffle scan (no CUB);
//   * the CSR transpose takes ONE shared-memory atomic per non-zero (the
//     returned rank stays in a register) instead of the reference's two CAS-loop
//     passes over 16-bit counters, and can sort each list so the neighbour order
//     is deterministic (the reference's order is a race);
//   * ribbon -> owner resolution is a direct table lookup, not a cuckoo probe.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "patch_layout.h"

namespace rxm {
namespace dev {

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// an mbarrier object must be invalidated before its storage is initialised again (PTX: mbarrier.init on a live
// object is undefined); used by code that runs several load phases per kernel through one static barrier
__device__ __forceinline__ void mbar_inval(uint64_t* bar)
{
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// make the barrier initialisation visible to the async (TMA) proxy
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// generic-proxy writes to shared memory -> visible to subsequent async-proxy ops
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}

// TMA bulk copy global -> shared (1-D), completion counted in bytes on `bar`.
// dst, src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// TMA bulk copy shared -> global (1-D) + group commit / wait
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_all_read()
{
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// streaming (no L1 allocate) global accesses for data touched once
__device__ __forceinline__ float ldg_stream(const float* p)
{
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// ------------------------------------------------------------------ packed fp32x2 arithmetic
// sm_100a executes two fp32 operations per issued instruction (SASS FFMA2 / FADD2 / FMUL2, PTX *.f32x2).
// The FP32 pipe time is that of two scalar instructions, but only ONE issue slot is spent, and integer / LSU
// instructions issue in the shadow (scripts/microbench/ffma2.cu: 8 FFMA2 + 8 IADD take the cycles of 8 FFMA2
// alone, 16 FFMA + 8 IADD take the sum).  The fan kernels are issue-bound, so they process TWO vertices per
// thread with every arithmetic instruction packed across the pair.  Each lane is an ordinary IEEE fp32
// operation: results are bit-identical to the scalar code.
struct f2
{
    unsigned long long v;
};
__device__ __forceinline__ f2 pk(float lo, float hi)
{
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk(f2 a, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
}
__device__ __forceinline__ f2 neg2(f2 a)  // folded by ptxas into the consumer's operand modifier
{
    float lo, hi;
    upk(a, lo, hi);
    return pk(-lo, -hi);
}
__device__ __forceinline__ f2 add2(f2 a, f2 b)
{
    f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b)
{
    f2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b)
{
    f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c)
{
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}

// ------------------------------------------------------------------ shared memory carve-up
// Bump allocator over dynamic shared memory; every allocation is 16-byte aligned
// (TMA destinations need it; the reference aligns to 8, shmem_allocator.cuh:17).
struct Smem
{
    uint8_t* base;
    uint32_t used;
    __device__ __forceinline__ explicit Smem(uint8_t* b) : base(b), used(0) {}
    template <typename T>
    __device__ __forceinline__ T* alloc(uint32_t count)
    {
        T* p = reinterpret_cast<T*>(base + used);
        used += (count * (uint32_t)sizeof(T) + 15u) & ~15u;
        return p;
    }
};

// ------------------------------------------------------------------ block exclusive scan
// In-place exclusive prefix sum of a[0..n) in shared memory; a[n] receives the
// total. warp_tmp: >= 33 u32 of shared scratch. Must be called by all BT threads;
// contains the barriers that make `a` consistent on entry and on exit.
template <int BT>
__device__ __forceinline__ void block_exclusive_scan(uint32_t* a, uint32_t n, uint32_t* warp_tmp)
{
    constexpr int  NW   = BT / 32;
    const uint32_t tid  = threadIdx.x;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    const uint32_t ipt  = (n + BT - 1) / BT;
    const uint32_t beg  = min(tid * ipt, n), end = min(beg + ipt, n);
    __syncthreads();
    uint32_t sum = 0;
    for (uint32_t i = beg; i < end; ++i)
        sum += a[i];
    uint32_t inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (uint32_t)d) inc += t;
    }
    if (lane == 31) warp_tmp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w  = lane < NW ? warp_tmp[lane] : 0;
        uint32_t wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= (uint32_t)d) wi += t;
        }
        if (lane < NW) warp_tmp[lane] = wi - w;
        if (lane == NW - 1) warp_tmp[32] = wi;
    }
    __syncthreads();
    uint32_t run = warp_tmp[warp] + inc - sum;
    for (uint32_t i = beg; i < end; ++i) {
        uint32_t c = a[i];
        a[i]       = run;
        run += c;
    }
    if (tid == 0) a[n] = warp_tmp[32];
    __syncthreads();
}

// ------------------------------------------------------------------ CSR transpose
// Builds, in shared memory, the CSR of the transpose of a sparse incidence with
// `nnz` non-zeros: non-zero i sits in column col(i) and contributes value
// val(i). One shared atomic per non-zero; its return value (the rank inside the
// column) is kept in a register until the offsets are known.
// off: u32[ncols+1] (zeroed here), out: u16[nnz]. Requires nnz <= KMAX*BT.
template <int BT, int KMAX, typename ColFn, typename ValFn>
__device__ __forceinline__ void csr_transpose(uint32_t  nnz,
                                              uint32_t  ncols,
                                              uint32_t* off,
                                              uint16_t* out,
                                              uint32_t* warp_tmp,
                                              ColFn     col,
                                              ValFn     val)
{
    const uint32_t tid = threadIdx.x;
    for (uint32_t i = tid; i <= ncols; i += BT)
        off[i] = 0;
    __syncthreads();
    uint16_t rank[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        const uint32_t i = tid + k * BT;
        if (i < nnz) rank[k] = (uint16_t)atomicAdd(&off[col(i)], 1u);
    }
    block_exclusive_scan<BT>(off, ncols, warp_tmp);
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        const uint32_t i = tid + k * BT;
        if (i < nnz) out[off[col(i)] + rank[k]] = (uint16_t)val(i);
    }
    __syncthreads();
}

// Sort every list ascending (deterministic neighbour order). One thread per list.
template <int BT>
__device__ __forceinline__ void csr_sort_lists(uint32_t n_lists, const uint32_t* off, uint16_t* val)
{
    for (uint32_t c = threadIdx.x; c < n_lists; c += BT) {
        const uint32_t b = off[c], e = off[c + 1];
        for (uint32_t i = b + 1; i < e; ++i) {
            const uint16_t x = val[i];
            uint32_t       j = i;
            while (j > b && val[j - 1] > x) {
                val[j] = val[j - 1];
                --j;
            }
            val[j] = x;
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------ owner resolution
struct OwnerTable
{
    const uint32_t*   own;    // shared: packed owner of local id (n_owned + i)
    const StashEntry* stash;  // shared
    uint32_t          n_owned;
    uint32_t          patch;      // this patch id
    uint32_t          slot_base;  // this patch's slot base for the element type
    uint32_t          type;

    __device__ __forceinline__ uint64_t handle(uint32_t lid) const
    {
        if (lid < n_owned) return ((uint64_t)patch << 32) | lid;
        if (lid == 0xFFFFu) return INVALID64_;  // empty slot of a fixed-width result (EVDiamond / EE on a boundary)
        const uint32_t o = own[lid - n_owned];
        return ((uint64_t)stash[o >> 16].patch << 32) | (o & 0xFFFFu);
    }
    __device__ __forceinline__ uint32_t slot(uint32_t lid) const
    {
        if (lid < n_owned) return slot_base + lid;
        const uint32_t o = own[lid - n_owned];
        return stash[o >> 16].slot_base[type] + (o & 0xFFFFu);
    }
};

}  // namespace dev
}  // namespace rxm
