// rxmesh/launch_box.h -- LaunchBox<blockThreads>: what RXMeshStatic::prepare_launch_box fills in and what a user passes
// to <<<...>>> or RXMeshStatic::run_kernel.  Member names are the reference's public contract
// (include/rxmesh/launch_box.h:12-18); everything else here is ours.
#pragma once
#include <cstddef>
#include <cstdint>

namespace rxmesh {

template <uint32_t blockThreads>
struct LaunchBox
{
    static_assert(blockThreads % 32 == 0 && blockThreads >= 32 && blockThreads <= 1024, "block size must be whole warps");

    // grid: one block per patch of the (local shard of the) mesh
    uint32_t blocks = 0;
    // block: fixed by the template argument, as in the reference
    const uint32_t num_threads = blockThreads;
    // shared memory: what the queried ops need per block (dynamic) and what the kernel declares itself (static)
    size_t smem_bytes_dyn    = 0;
    size_t smem_bytes_static = 0;
    // from cudaFuncGetAttributes on the user's kernel: occupancy hints only
    uint32_t num_registers_per_thread = 0;
    size_t   local_mem_per_thread     = 0;

    // resident blocks per SM this configuration allows on a B200 SM (228 KB shared memory, 64 K registers, 2048 threads)
    uint32_t blocks_per_sm() const
    {
        const size_t   smem = smem_bytes_dyn + smem_bytes_static + 1024;
        const uint32_t by_smem = (uint32_t)((228u * 1024u) / (smem ? smem : 1));
        const uint32_t by_regs = num_registers_per_thread ? 65536u / (num_registers_per_thread * blockThreads) : 32u;
        const uint32_t by_thr  = 2048u / blockThreads;
        uint32_t       b       = by_smem < by_regs ? by_smem : by_regs;
        b                      = b < by_thr ? b : by_thr;
        return b < 32u ? b : 32u;
    }
};

}  // namespace rxmesh
