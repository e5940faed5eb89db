// rxmesh/kernels/for_each.cuh -- device-side for_each<Op::V | Op::E | Op::F, blockThreads>(context, lambda)
// (include/rxmesh/kernels/for_each.cuh:108-168): the block applies the lambda to every OWNED element of its patch.
#pragma once
#include "rxmesh/context.h"
#include "rxmesh/handle.h"
#include "rxmesh/types.h"

namespace rxmesh {
template <Op op, uint32_t blockThreads, typename computeT>
__device__ __inline__ void for_each(const Context& context, computeT compute_op)
{
    static_assert(op == Op::V || op == Op::E || op == Op::F,
                  "for_each() only accepts unary operator for its template parameter i.e., Op::V, Op::E, or Op::F");
    using HandleT = std::conditional_t<op == Op::V, VertexHandle, std::conditional_t<op == Op::E, EdgeHandle, FaceHandle>>;
    const uint32_t p_id = blockIdx.x;
    if (p_id < context.get_num_patches()) {
        const rxm::PatchDesc* d = context.view.desc + p_id;
        const uint32_t        n = d->n_owned[HandleT::elem], pid = d->patch_id;
        for (uint32_t i = threadIdx.x; i < n; i += blockThreads) {
            HandleT h(pid, typename HandleT::LocalT((uint16_t)i));  // an lvalue: user lambdas may take HandleT&
            compute_op(h);
        }
    }
    __syncthreads();
}
}  // namespace rxmesh
