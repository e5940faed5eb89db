// rxmesh/launch_box.h -- LaunchBox (include/rxmesh/launch_box.h:12-18)
#pragma once
#include <stddef.h>
#include <stdint.h>
namespace rxmesh {
template <uint32_t blockThreads>
struct LaunchBox
{
    uint32_t       blocks = 0, num_registers_per_thread = 0;
    size_t         smem_bytes_dyn = 0, smem_bytes_static = 0;
    size_t         local_mem_per_thread = 0;
    const uint32_t num_threads = blockThreads;
};
}  // namespace rxmesh
