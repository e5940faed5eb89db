// rxmesh/kernels/shmem_allocator.cuh -- ShmemAllocator (include/rxmesh/kernels/shmem_allocator.cuh:15-122):
// bump allocator over the dynamic shared memory of the block; 16-byte aligned (TMA destinations).
#pragma once
#include "rxmesh_b200/rxm_device.cuh"
namespace rxmesh {
extern __shared__ __align__(128) uint8_t SHMEM_START[];
struct ShmemAllocator
{
    __device__ ShmemAllocator() : m_sm(SHMEM_START) {}
    template <typename T>
    __device__ T* alloc(uint32_t count) { return m_sm.template alloc<T>(count); }
    __device__ char* alloc(uint32_t num_bytes) { return (char*)m_sm.template alloc<uint8_t>(num_bytes); }
    __device__ void  dealloc(uint32_t num_bytes) { m_sm.used -= (num_bytes + 15u) & ~15u; }
    template <typename T>
    __device__ void     dealloc(uint32_t count) { dealloc(count * (uint32_t)sizeof(T)); }
    __device__ uint32_t get_allocated_size_bytes() const { return m_sm.used; }
    rxm::dev::Smem m_sm;
};
}  // namespace rxmesh
