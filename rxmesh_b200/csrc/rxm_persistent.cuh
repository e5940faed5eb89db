// rxm_persistent.cuh -- persistent, software-pipelined patch kernels (sm_100a).
//
// The block-per-patch kernels expose a dependent latency chain per patch
// (descriptor load -> TMA of the sections -> ribbon gather -> compute) that only
// occupancy can hide (ncu, profiles/r01_*: 53-65 % issue-active, 28-34 % DRAM).
// Here every CTA is persistent (grid = #SMs x resident CTAs) and walks its
// patches through a 4-deep pipeline; patch j is touched at four consecutive
// "ticks":
//   tick j   : thread 0 TMA-loads the 128-byte PatchDesc into a shared ring
//   tick j+1 : thread 0 reads it and issues the TMA bulk copies of the sections
//              and of the patch's owned attribute slice into stage j % 4
//   tick j+2 : all threads (after the stage's mbarrier) run W::pre: convert /
//              transpose in shared memory and issue the ribbon gathers as
//              cp.async (LDGSTS) straight into shared memory
//   tick j+3 : cp.async.wait_group 1 + ONE __syncthreads, then W::compute
// so that while patch j is computed the loads of j+1, j+2 and the descriptor of
// j+3 are in flight.  Results are written straight to global memory.
//
// A worker W provides: Args, Layout (stage carve-up, host-computed from the mesh
// maxima), issue() [thread 0], pre() and compute() [all threads].
#pragma once
#include "rxmesh_b200/rxm_device.cuh"

namespace rxm {
namespace dev {

constexpr int PIPE_STAGES = 4;
constexpr int DESC_RING   = 8;

__device__ __forceinline__ void cp_async_4(void* dst_smem, const void* src_gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit()
{
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <class W, int BT>
__global__ void __launch_bounds__(BT) k_persistent(MeshView mv, typename W::Args args, typename W::Layout lay)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t                      full_bar[PIPE_STAGES];
    __shared__ uint64_t                      desc_bar[DESC_RING];
    __shared__ __align__(16) PatchDesc       s_desc[DESC_RING];

    const uint32_t P = mv.num_patches, stride = gridDim.x, first = blockIdx.x;
    const uint32_t n_local = first < P ? (P - first + stride - 1) / stride : 0u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < PIPE_STAGES; ++i)
            mbar_init(&full_bar[i], 1);
        for (int i = 0; i < DESC_RING; ++i)
            mbar_init(&desc_bar[i], 1);
        fence_mbar_init();
    }
    __syncthreads();

    for (uint32_t t = 0; t < n_local + 3; ++t) {
        // ---- producer (thread 0) ----
        if (threadIdx.x == 0) {
            if (t < n_local) {  // descriptor of patch t
                const uint32_t s = t % DESC_RING;
                mbar_arrive_expect_tx(&desc_bar[s], (uint32_t)sizeof(PatchDesc));
                bulk_g2s(&s_desc[s], mv.desc + (first + (uint64_t)t * stride), (uint32_t)sizeof(PatchDesc), &desc_bar[s]);
            }
            if (t >= 1 && t - 1 < n_local) {  // sections of patch t-1
                const uint32_t j = t - 1, s = j % DESC_RING;
                mbar_wait(&desc_bar[s], (j / DESC_RING) & 1u);
                const PatchDesc& d = s_desc[s];
                W::issue(d, mv.topo + d.topo_off, args, lay, smem_raw + (j % PIPE_STAGES) * lay.stage_bytes,
                         &full_bar[j % PIPE_STAGES]);
            }
        }
        // ---- pre: patch t-2 ----
        if (t >= 2 && t - 2 < n_local) {
            const uint32_t j = t - 2;
            mbar_wait(&desc_bar[j % DESC_RING], (j / DESC_RING) & 1u);
            mbar_wait(&full_bar[j % PIPE_STAGES], (j / PIPE_STAGES) & 1u);
            W::pre(s_desc[j % DESC_RING], args, lay, smem_raw + (j % PIPE_STAGES) * lay.stage_bytes);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        // ---- compute: patch t-3 ----
        if (t >= 3 && t - 3 < n_local) {
            const uint32_t j = t - 3;
            W::compute(s_desc[j % DESC_RING], args, lay, smem_raw + (j % PIPE_STAGES) * lay.stage_bytes);
        }
    }
}

}  // namespace dev
}  // namespace rxm
