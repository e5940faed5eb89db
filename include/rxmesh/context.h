// rxmesh/context.h -- Context: the by-value kernel argument (include/rxmesh/context.h:15-441). Here it is
// the B200 MeshView (patch descriptors + topology blob) -- 80 bytes, no pointer chasing.
#pragma once
#include "rxmesh/handle.h"
namespace rxmesh {
class Context
{
   public:
    rxm::MeshView view{};
    __host__ __device__ uint32_t get_num_patches() const { return view.num_patches; }
    __host__ __device__ uint32_t get_num_vertices() const { return view.num_elems[rxm::ELEM_V]; }
    __host__ __device__ uint32_t get_num_edges() const { return view.num_elems[rxm::ELEM_E]; }
    __host__ __device__ uint32_t get_num_faces() const { return view.num_elems[rxm::ELEM_F]; }
    // linear_id (context.h:275-290) of an OWNER handle: prefix[patch] + local
    template <typename HandleT>
    __device__ uint32_t linear_id(HandleT h) const
    {
        return view.desc[h.patch_id()].lin_base[HandleT::elem] + h.local_id();
    }
};
}  // namespace rxmesh
