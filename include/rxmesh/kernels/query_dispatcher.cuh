// rxmesh/kernels/query_dispatcher.cuh -- the free-function forms of the query API
// (include/rxmesh/kernels/query_dispatcher.cuh:181-565): query_block_dispatcher<op, blockThreads>(...) in its four
// overloads and higher_query_block_dispatcher<op, blockThreads>(context, src_id, lambda) for queries on elements
// that are not assigned one-per-thread by the patch (k-ring traversals such as apps/Filtering).
#pragma once
#include "rxmesh/query.h"

namespace rxmesh {
// (block, shrd_alloc, context, compute_op, compute_active_set, oriented)   query_dispatcher.cuh:320-342
template <Op op, uint32_t blockThreads, typename computeT, typename activeSetT>
__device__ __inline__ void query_block_dispatcher(cooperative_groups::thread_block& block, ShmemAllocator& shrd_alloc,
                                                  const Context& context, computeT compute_op,
                                                  activeSetT compute_active_set, const bool oriented = false)
{
    if (blockIdx.x >= context.get_num_patches()) return;
    Query<blockThreads> query(context);
    query.template dispatch<op>(block, shrd_alloc, compute_op, compute_active_set, oriented);
}
// (context, compute_op, compute_active_set, oriented)   query_dispatcher.cuh:346-366
template <Op op, uint32_t blockThreads, typename computeT, typename activeSetT>
__device__ __inline__ void query_block_dispatcher(const Context& context, computeT compute_op, activeSetT compute_active_set,
                                                  const bool oriented = false)
{
    auto           block = cooperative_groups::this_thread_block();
    ShmemAllocator shrd_alloc;
    query_block_dispatcher<op, blockThreads>(block, shrd_alloc, context, compute_op, compute_active_set, oriented);
}
// (block, shrd_alloc, context, compute_op, oriented)   query_dispatcher.cuh:381-402
template <Op op, uint32_t blockThreads, typename computeT>
__device__ __inline__ void query_block_dispatcher(cooperative_groups::thread_block& block, ShmemAllocator& shrd_alloc,
                                                  const Context& context, computeT compute_op, const bool oriented = false)
{
    using InH = typename InputHandle<op>::type;
    query_block_dispatcher<op, blockThreads>(block, shrd_alloc, context, compute_op, [](InH) { return true; }, oriented);
}
// (context, compute_op, oriented)   query_dispatcher.cuh:407-421
template <Op op, uint32_t blockThreads, typename computeT>
__device__ __inline__ void query_block_dispatcher(const Context& context, computeT compute_op, const bool oriented = false)
{
    using InH = typename InputHandle<op>::type;
    query_block_dispatcher<op, blockThreads>(context, compute_op, [](InH) { return true; }, oriented);
}

// higher_query_block_dispatcher (query_dispatcher.cuh:445-565): every thread brings its own source handle (any patch;
// HandleT() = not participating); the block visits the distinct patches among them, answers the query of each and
// hands every thread the iterator of its element.  Called by the whole block.  The reference finds the distinct
// patches with a block radix sort + discontinuity flags; here each round takes the smallest patch id not yet served
// (one shared atomicMin), which needs no sort storage and visits the same set.
template <Op op, uint32_t blockThreads, typename computeT, typename HandleT>
__device__ __inline__ void higher_query_block_dispatcher(const Context& context, const HandleT src_id, computeT compute_op,
                                                         const bool oriented = false)
{
    __shared__ uint32_t s_next_patch;
    const bool          valid = src_id.is_valid();
    const uint32_t      mine  = valid ? src_id.patch_id() : INVALID32;
    bool                served = !valid;
    ShmemAllocator      shrd_alloc;
    int                 round = 0;
    while (true) {
        if (threadIdx.x == 0) s_next_patch = INVALID32;
        __syncthreads();
        if (!served) atomicMin(&s_next_patch, mine);
        __syncthreads();
        const uint32_t p = s_next_patch;
        __syncthreads();  // everyone has read s_next_patch before the next round resets it
        if (p == INVALID32) break;
        Query<blockThreads> query(context, p);
        query.template dispatch_src<op>(shrd_alloc, !served && mine == p, src_id, compute_op, oriented, round++);
        if (mine == p) served = true;
    }
    if (round) Query<blockThreads>::end_rounds();
}
}  // namespace rxmesh
