// mesh_builder.h -- host-side mesh build for the static path.
//
// B200-first replacement of RXMesh::build / build_supporting_structures /
// build_single_patch_ltog / build_single_patch_topology / populate_patch_stash
// (/root/reference/include/rxmesh/rxmesh.cpp:290-448,518-681,754-996,1096-1137)
// and of the patcher's host passes extract_ribbons / assign_patch
// (patcher/patcher.cu:640-757), written over flat arrays so that it scales
// linearly in the number of faces (the reference is O(P*(V+E))).
#pragma once
#include <cstdint>
#include <memory>
#include <utility>
#include <string>
#include <vector>

#include "rxmesh_b200/patch_layout.h"

namespace rxm {

// std::vector storage whose resize() leaves trivially constructible elements uninitialised: the 5.6 GB patch store of a
// 100 M-face mesh is then zero-filled by all cores instead of one
template <typename T>
struct NoInitAlloc : std::allocator<T>
{
    template <typename U>
    struct rebind
    {
        using other = NoInitAlloc<U>;
    };
    template <typename U, typename... A>
    void construct(U* p, A&&... a)
    {
        if constexpr (sizeof...(A) == 0)
            ::new ((void*)p) U;
        else
            ::new ((void*)p) U(std::forward<A>(a)...);
    }
};
using ByteBuf = std::vector<uint8_t, NoInitAlloc<uint8_t>>;
using U32Buf  = std::vector<uint32_t, NoInitAlloc<uint32_t>>;  // the element-count-sized id maps: filled by all cores

struct HostMesh
{
    // ---- global sizes / input statistics (rxmesh.h getters) ----
    uint32_t num_elems[3]   = {0, 0, 0};  // V, E, F
    uint32_t num_patches    = 0;
    uint32_t patch_size     = 512;
    uint32_t max_valence    = 0;
    uint32_t max_edge_incident_faces = 0;
    uint32_t max_face_adjacent_faces = 0;
    bool     is_closed        = true;
    bool     is_edge_manifold = true;
    uint32_t max_per_patch[3]       = {0, 0, 0};  // ribbon included
    uint32_t max_owned_per_patch[3] = {0, 0, 0};
    uint32_t max_not_owned[3]       = {0, 0, 0};
    uint32_t max_stash              = 0;
    uint32_t num_slots[3]           = {0, 0, 0};
    uint64_t total_local[3]         = {0, 0, 0};  // sum over patches of n[t] (ribbon stats)
    bool     packed                 = false;  // rank-annotated patch format (patch_layout.h)
    bool     fans                   = false;  // one-ring fans stored (manifold, consistently oriented input)
    uint32_t max_fan_total          = 0;
    bool     ring2                  = false;  // ring-2 extension stored (patch_layout.h, FLAG_RING2)
    uint32_t max_ext = 0, max_r2 = 0, max_r2_total = 0;
    double   build_seconds          = 0;
    double   patcher_seconds        = 0;
    uint32_t lloyd_runs             = 0;  // assignment passes of the built-in patcher (0: caller-supplied patches)

    // ---- the patch store ----
    std::vector<PatchDesc> desc;
    ByteBuf                topo;          // all patch blobs back to back (zero-filled in parallel by the builder)
    std::vector<uint32_t>  slot_base[3];  // [P+1]
    std::vector<uint32_t>  lin_base[3];   // [P+1]

    // ---- id maps ----
    U32Buf                ltog[3];      // concatenated local->global, per patch (owned first)
    std::vector<uint64_t> ltog_off[3];  // [P+1]
    U32Buf                slot_to_global[3];  // [num_slots], INVALID32_ for padding slots
    U32Buf                global_to_slot[3];  // [num_elems]
    std::vector<uint32_t> elem_patch[3];      // [num_elems] owner patch of every element

    // global edges (kept for host-side queries such as get_edge_id)
    U32Buf ev;  // 2*E: (larger id, smaller id)
    U32Buf fe;  // 3*F: global edge ids
};

struct BuildOptions
{
    uint32_t patch_size   = 512;
    int      num_threads  = 0;     // 0 = omp default
    bool     keep_ltog    = true;  // false: drop ltog after the build (large meshes)
    bool     verbose      = false;
    uint32_t lloyd_iters  = 5;
    bool     force_wide   = false;  // never use the packed format (tests of the atomic path)
    bool     no_fans      = false;  // never store one-ring fans (tests of the generic kernels)
    uint32_t ring_depth   = 2;      // rings around the owned vertices whose vertices carry their complete one-ring
    bool     reorder_patches = true;  // renumber the Lloyd patches in breadth-first (locality) order; a caller-supplied
                                      // face->patch array is always honoured as is
    bool     no_ring2     = false;  // skip the ring-2 extension (meshes that never run a k-ring consumer: saves build time
                                    // and ~3 bytes per face of patch store)
};

// Global edge numbering identical to the reference (first appearance while
// scanning faces, rxmesh.cpp:589-611) computed with counting sorts instead of a
// hash map. ev: 2*E (max id, min id); fe: 3*F. Returns the number of edges.
uint32_t build_edges(const uint32_t* fv, uint32_t nf, uint32_t nv, U32Buf& ev, U32Buf& fe);

// Deterministic Lloyd clustering of faces over the face-adjacency graph; the
// role of patcher::Patcher::run_lloyd (patcher/patcher.cu:828-987).
void patcher_lloyd(const uint32_t* fe, uint32_t nf, uint32_t ne, uint32_t patch_size,
                   uint32_t lloyd_iters, std::vector<uint32_t>& face_patch, uint32_t& num_patches,
                   uint32_t* lloyd_runs = nullptr);  // lloyd_runs: assignment passes run (Patcher::get_num_lloyd_run)

// The outer Lloyd loop of patcher_lloyd on the current CUDA device (rxm_patcher_gpu.cu); false = not run, use the host passes.
bool patcher_lloyd_gpu(const std::vector<uint32_t>& ff_off, const std::vector<uint32_t>& ff_val, uint32_t nf, uint32_t patch_size,
                       uint32_t lloyd_iters, std::vector<uint32_t>& seeds, std::vector<uint32_t>& face_patch,
                       std::vector<uint32_t>& queue, std::vector<uint32_t>& psize, int* n_assign);

// The built-in patcher alone: face -> patch (Lloyd + locality ordering of the ids), no patch store.
std::string compute_face_patch(const uint32_t* fv, uint32_t nf, const BuildOptions& opt, std::vector<uint32_t>& face_patch,
                               uint32_t& num_patches);

// Builds everything. face_patch may be null (run the Lloyd patcher) or a
// user-supplied face->patch assignment (the analogue of the reference's
// patcher_file constructor argument, rxmesh_static.h:61-66). Returns "" on
// success, an error message otherwise.
std::string build_mesh(const uint32_t* fv, uint32_t nf, const uint32_t* face_patch,
                       const BuildOptions& opt, HostMesh& out);

}  // namespace rxm
