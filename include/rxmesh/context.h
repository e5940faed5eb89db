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
    // get_owner_handle (context.h:219-270): a (patch, local id) naming a not-owned copy -> the owner's handle, through
    // the patch's direct owner table in the topology blob (two dependent global loads, no hash probe)
    template <typename HandleT>
    __device__ HandleT get_owner_handle(const HandleT handle) const
    {
        const rxm::PatchDesc& d   = view.desc[handle.patch_id()];
        const uint32_t        lid = handle.local_id(), t = HandleT::elem;
        if (lid < d.n_owned[t]) return handle;
        const uint8_t*  blob = view.topo + d.topo_off;
        const uint32_t  o    = reinterpret_cast<const uint32_t*>(blob + d.off_own(t))[lid - d.n_owned[t]];
        const uint32_t  pid  = reinterpret_cast<const rxm::StashEntry*>(blob + d.off_stash())[o >> 16].patch;
        return HandleT(pid, typename HandleT::LocalT((uint16_t)(o & 0xFFFFu)));
    }
    // get_handle (context.h:330-352): the inverse of linear_id -- binary search over the per-patch prefix
    template <typename HandleT>
    __device__ HandleT get_handle(const uint32_t i) const
    {
        uint32_t lo = 0, hi = view.num_patches;  // last patch whose prefix <= i
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) / 2;
            if (view.desc[mid].lin_base[HandleT::elem] <= i)
                lo = mid;
            else
                hi = mid;
        }
        return HandleT(view.desc[lo].patch_id, typename HandleT::LocalT((uint16_t)(i - view.desc[lo].lin_base[HandleT::elem])));
    }
    // linear_id (context.h:275-290) of an OWNER handle: prefix[patch] + local
    template <typename HandleT>
    __device__ uint32_t linear_id(HandleT h) const
    {
        return view.desc[h.patch_id()].lin_base[HandleT::elem] + h.local_id();
    }
};
}  // namespace rxmesh
