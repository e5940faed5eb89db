// rxm_kernels.h -- host-callable launchers of the fixed-function sm_100a kernels.
#pragma once
#include <cuda_runtime.h>

#include "rxmesh_b200/patch_layout.h"

namespace rxm {

struct KernelLimits
{
    uint32_t max_n[3];
    uint32_t max_owned[3];
    uint32_t max_not_owned[3];
    uint32_t max_stash;
    uint32_t max_face_adjacent_faces;
    uint32_t max_fan_total;
    uint32_t max_ext, max_r2, max_r2_total;  // ring-2 extension (patch_layout.h)
};

// Every launcher returns cudaSuccess or the launch error; `err` (may be null)
// receives a human-readable reason when the configuration is unsupported.
cudaError_t launch_query_store(int op, const MeshView& mv, const KernelLimits& lim, AttrView<uint64_t> in,
                               AttrView<uint64_t> out, cudaStream_t stream, const char** err);

cudaError_t launch_query_consume(int op, const MeshView& mv, const KernelLimits& lim, AttrView<float> in,
                                 AttrView<float> out, cudaStream_t stream, const char** err);

cudaError_t launch_vertex_normals(const MeshView& mv, const KernelLimits& lim, const float* coords_aos,
                                  float* normals_aos, int unit_face_normals, cudaStream_t stream,
                                  const char** err);

// Laplacian step fused with the ribbon (halo) exchange of a patch-sharded mesh: the kernel that produces the new
// positions also stores the rows mirrored on neighbouring GPUs straight into their ghost slots (NVLink P2P stores
// through cudaIpc-mapped pointers) and, when its last block is done, raises a flag on every neighbour; blocks whose
// patch reads ghost slots wait for the neighbours' flags of the previous step.  No NCCL call, no host synchronisation.
struct FusedHaloView
{
    const uint32_t*  push_off;     // [P + 1] over the mesh's patch slots (index = patch index in the descriptor array)
    const uint2*     push;         // {local vertex id | neighbour index << 16, slot on the neighbour}
    float* const*    peer_out;     // [npeers] base of the OUT attribute in each neighbour's memory
    uint32_t* const* peer_flag;    // [npeers] address, in the neighbour's memory, of its flags[me]
    const uint8_t*   reads_ghost;  // [P] the patch has ribbon elements owned by a ghost patch
    uint32_t*        flags;        // [npeers] raised by the neighbours (step count they have finished)
    uint32_t*        done_ctr;     // blocks of this launch that are past their ghost reads and pushes
    uint32_t         npeers, first, step;
    uint32_t         shift;          // block b works on patch (b + shift) % #patches
    uint32_t         n_sync_blocks;  // patches of the active range that read ghost slots or have rows to push
};
cudaError_t launch_laplacian_step_fused(const MeshView& mv, const KernelLimits& lim, const float* x_in_aos, float* x_out_aos,
                                        double lr, const FusedHaloView& fh, cudaStream_t stream, const char** err);

cudaError_t launch_laplacian_step(const MeshView& mv, const KernelLimits& lim, const float* x_in_aos,
                                  float* x_out_aos, double lr, cudaStream_t stream, const char** err);

// materialise a query as a CSR over attribute slots (patch-grouped): csr_off[num_slots(src)+1] (the last
// entry is written by the caller), csr_val[nnz] = owner slots; patch_nnz_off[P] = first entry of every patch
cudaError_t launch_query_csr(int op, const MeshView& mv, const KernelLimits& lim, const uint32_t* patch_nnz_off,
                             uint32_t* csr_off, uint32_t* csr_val, cudaStream_t stream, const char** err);

cudaError_t launch_bilateral_step(const uint32_t* csr_off, const uint32_t* csr_val, uint32_t num_slots, const float* x_aos,
                                  const float* normals_aos, float* x_out_aos, uint32_t* overflow_flag, cudaStream_t stream);

// bilateral filtering, one block per patch, normals fused (the default), followed by the compacted cross-patch pass over
// the vertices the patches deferred.  flags: 3 x u32 -- [0] = a neighbourhood exceeded 80 vertices, [1] += deferred
// vertices, [2] / [3] work-list fill of even / odd iterations (all zero before iteration 0); work: 20 bytes per vertex slot
// (worst case: every vertex deferred); csr_* = the VV CSR over slots
cudaError_t launch_bilateral_patch(const MeshView& mv, const KernelLimits& lim, const uint32_t* csr_off, const uint32_t* csr_val,
                                   const float* x_aos, float* x_out_aos, uint32_t* flags, void* work, uint32_t iteration,
                                   cudaStream_t stream, const char** err);

cudaError_t launch_boundary_vertices(const MeshView& mv, const KernelLimits& lim, uint32_t* flag_per_slot,
                                     cudaStream_t stream, const char** err);

// out_slot[slot*nattr + a] = in_global[global_of_slot*nattr + a] (AoS), 0 for padding slots
// Where the attribute storage of a range of patches lives, for the layout-aware copy kernels: slot_base / lin_base
// point at the first patch of the range ([num_patches + 1] entries each, absolute values).
struct SlotMap
{
    const uint32_t* slot_base;
    const uint32_t* lin_base;
    uint32_t        num_patches;
    uint32_t        num_slots;  // of the whole mesh (AoS / AoSoA strides)
    uint32_t        num_elems;  // of the whole mesh (SoA column stride)
};

// 32-bit attributes: copy between layouts over the same elements
cudaError_t launch_relayout(const void* src, void* dst, const SlotMap& sm, uint32_t nattr, uint32_t layout_src,
                            uint32_t layout_dst, cudaStream_t stream);
cudaError_t launch_permute_to_slots(const void* in_global, void* out_slots, const uint32_t* slot_to_global, const SlotMap& sm,
                                    uint32_t elem_bytes, uint32_t nattr, uint32_t layout, cudaStream_t stream);
cudaError_t launch_permute_to_global(const void* in_slots, void* out_global, const uint32_t* slot_to_global, const SlotMap& sm,
                                     uint32_t elem_bytes, uint32_t nattr, uint32_t layout, cudaStream_t stream);

// halo exchange helpers (AoS attributes, rows of row_words 32-bit words)
cudaError_t launch_slot_rows(bool gather, void* attr, const uint32_t* idx, uint64_t n, uint32_t row_words, void* buf,
                             cudaStream_t stream);
cudaError_t launch_push_rows(const void* local, const uint32_t* local_idx, void* remote, const uint32_t* remote_idx,
                             uint64_t n, uint32_t row_words, cudaStream_t stream);

// reductions over owned elements; scratch: (grid + 1) x 16 bytes, result (double value, u64 handle) at scratch[grid]
cudaError_t launch_reduce(const MeshView& mv, AttrView<float> a, AttrView<float> b, int elem, int kind, uint32_t attr_id,
                          void* scratch, uint32_t grid, cudaStream_t stream);

cudaError_t launch_fill(void* data, uint64_t count, uint32_t elem_bytes, const void* value, cudaStream_t stream);

// ---- mean-curvature flow by matrix-free CG (rxm_mcf.cu; apps/MCF/mcf_cg_mat_free.h, matrix/cg_mat_free_attr_solver.h) ----
// The solver's scalars live on the device: the last block of every kernel reduces the per-block partial sums in block order
// and does the scalar step (alpha, beta, convergence, iteration count), so an iteration needs no host synchronisation.
struct McfState
{
    double   dot_sp, delta_new, delta_old, start, final_res;
    float    alpha, beta;
    uint32_t iters;      // completed iterations (IterativeSolver::m_iter_taken)
    uint32_t converged;  // is_converged() held (or the start residual is exactly 0)
    uint32_t ctr;        // blocks that have published their partial sum in the running kernel
    uint32_t pad[3];
};
struct McfBuffers
{
    const uint32_t* fan_base;  // [P] first entry of every patch's slice of W (multiples of 4 entries)
    float*          W;         // time_step * max(0, cotangent weight) per fan entry (null for the uniform Laplacian)
    float*          diag;      // [num_slots(V)] 1 / v_weight + sum of the edge weights
    float *         R, *S, *X; // [num_slots(V)][3], AoS
    float*          P[2];      // search direction, double-buffered (iteration k reads P[k & 1], writes P[(k + 1) & 1])
    double*         partials;  // [partials_split + mcf_update_grid()]
    McfState*       state;
    uint32_t        partials_split;  // = number of patches
};
uint32_t    mcf_update_grid();
// precond: Jacobi-preconditioned CG (the MCF app's pcg_mat_free: Z = R / diag, delta = <R, Z>)
cudaError_t launch_mcf_setup(const MeshView& mv, const KernelLimits& lim, const float* x0_aos, const McfBuffers& B, bool uniform,
                             bool precond, float time_step, cudaStream_t stream, const char** err);
// iteration `it` (0-based): mat-vec kernel + update kernel
cudaError_t launch_mcf_iteration(const MeshView& mv, const KernelLimits& lim, const McfBuffers& B, uint32_t it, bool uniform,
                                 bool precond, float time_step, float tol_abs, float tol_rel, uint32_t max_iter, cudaStream_t stream,
                                 const char** err);
void count_launches(uint64_t n);  // adds to launch_counter() (kernels launched from other translation units)

// number of kernels launched by this library since load (bench.py "gpu_launches")
uint64_t launch_counter();

}  // namespace rxm
