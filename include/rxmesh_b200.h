/*
 * rxmesh_b200.h -- the C ABI of librxmesh_b200.so (B200 / sm_100a).
 *
 * The reference (owensgroup/RXMesh) has no FFI: its boundary is a C++ template
 * API compiled into the user's .cu.  This header is the thin
 * C-ABI layer that the drop-in C++ shim (include/rxmesh/ *.h, same class names
 * as the reference) and the Python host mirror (rxmesh_b200/) call.  Every
 * entry point names the reference interface it replaces; paths are relative to
 * /root/reference/include/rxmesh unless stated otherwise.
 *
 * Conventions: plain pointers and sizes only; every function that can fail
 * returns an int status (0 = RXM_OK) and records a thread-local message readable
 * through rxm_last_error(); `stream` is a cudaStream_t passed as void* (NULL =
 * default stream).  The reference's CUDA_ERROR macro logs and exit()s
 * (util/macros.h:77-89); this ABI returns the error instead, the C++ shim
 * restores the reference behaviour.
 */
#ifndef RXMESH_B200_H
#define RXMESH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RXM_OK 0
#define RXM_ERR_INVALID 1   /* bad argument */
#define RXM_ERR_CUDA 2      /* CUDA runtime / driver error (no device, launch failure ...) */
#define RXM_ERR_UNSUPPORTED 3

/* element types and query ops: numeric values of the reference's Op enum (types.h:113-129) */
enum { RXM_V = 0, RXM_E = 1, RXM_F = 2 };
enum { RXM_OP_VV = 3, RXM_OP_VE = 4, RXM_OP_VF = 5, RXM_OP_FV = 6, RXM_OP_FE = 7, RXM_OP_FF = 8,
       RXM_OP_EV = 9, RXM_OP_EE = 10, RXM_OP_EF = 11, RXM_OP_EVDIAMOND = 12 };
/* locationT (types.h:51-58) and layoutT (types.h:84-90) */
enum { RXM_HOST = 0x01, RXM_DEVICE = 0x02, RXM_LOCATION_ALL = 0x0F };
enum { RXM_AOS = 0, RXM_AOSOA = 1, RXM_SOA = 2 };

typedef struct rxm_mesh rxm_mesh; /* RXMeshStatic (rxmesh_static.h:37) */
typedef struct rxm_attr rxm_attr; /* Attribute<T,HandleT> (attribute.h:56) */

const char* rxm_last_error(void);
const char* rxm_version(void);

/* rx_init(device) (rxmesh.h:23-30): select the device, verify it is sm_100. */
int rxm_init(int device);

/* ---------------------------------------------------------------- mesh ---- */
/* RXMeshStatic(fv, patcher_file, patch_size, ...) (rxmesh_static.h:61-100; rxmesh.cpp:89-225).
 * fv: 3*num_faces zero-based vertex ids.  face_patch: NULL -> run the built-in Lloyd patcher
 * (patcher/patcher.cu:828-987); otherwise a face -> patch assignment to honour, the analogue of
 * constructing from a saved `patcher_file`.  Host-only work: succeeds without a GPU. */
int rxm_mesh_create(const uint32_t* fv, uint32_t num_faces, const uint32_t* face_patch, uint32_t patch_size,
                    int num_threads, rxm_mesh** out);
/* same, with build flags: sections of the patch store a mesh will never need can be left out */
#define RXM_BUILD_NO_RING2 1u /* no ring-2 extension (rxm_bilateral_filter then runs its cross-patch path for every vertex
                                 whose neighbourhood leaves the patch); saves build time and ~3 bytes per face */
#define RXM_BUILD_NO_PATCH_REORDER 2u /* keep the Lloyd patcher's seed-order patch ids (default: renumbered breadth-first over the
                                        patch graph from a peripheral patch, so that contiguous id ranges are compact slabs) */
int rxm_mesh_create_ex(const uint32_t* fv, uint32_t num_faces, const uint32_t* face_patch, uint32_t patch_size,
                       int num_threads, uint32_t flags, rxm_mesh** out);
/* RXMesh::build_device (rxmesh.cpp:1139-1615): upload the patch store to the current device. */
int rxm_mesh_to_device(rxm_mesh* m);
/* Release host-side helper arrays of a large mesh once it is on the device (local->global lists, global edge
 * arrays, the host copy of the patch store): rxm_mesh_patch / rxm_mesh_edges / rxm_query_csr / halo planning
 * are no longer available afterwards; attributes, kernels and the slot<->global maps keep working. */
int rxm_mesh_compact(rxm_mesh* m);
/* ~RXMesh (rxmesh.cpp:227-288) */
void rxm_mesh_destroy(rxm_mesh* m);

/* getters (rxmesh.h:44-399). `what` keys: */
enum {
    RXM_INFO_NUM_VERTICES = 0, RXM_INFO_NUM_EDGES = 1, RXM_INFO_NUM_FACES = 2, RXM_INFO_NUM_PATCHES = 3,
    RXM_INFO_PATCH_SIZE = 4, RXM_INFO_MAX_VALENCE = 5, RXM_INFO_MAX_EDGE_INCIDENT_FACES = 6,
    RXM_INFO_MAX_FACE_ADJACENT_FACES = 7, RXM_INFO_IS_CLOSED = 8, RXM_INFO_IS_EDGE_MANIFOLD = 9,
    RXM_INFO_MAX_VERTICES_PER_PATCH = 10, RXM_INFO_MAX_EDGES_PER_PATCH = 11, RXM_INFO_MAX_FACES_PER_PATCH = 12,
    RXM_INFO_NUM_SLOTS_V = 13, RXM_INFO_NUM_SLOTS_E = 14, RXM_INFO_NUM_SLOTS_F = 15,
    RXM_INFO_TOPO_BYTES = 16, RXM_INFO_TOTAL_LOCAL_V = 17, RXM_INFO_TOTAL_LOCAL_E = 18,
    RXM_INFO_TOTAL_LOCAL_F = 19, RXM_INFO_MAX_STASH = 20, RXM_INFO_ON_DEVICE = 21, RXM_INFO_PACKED = 22,
    RXM_INFO_FANS = 23, RXM_INFO_RING2 = 24,
    RXM_INFO_LLOYD_RUNS = 25,     /* Patcher::get_num_lloyd_run: assignment passes of the built-in patcher */
    RXM_INFO_NUM_COMPONENTS = 26  /* RXMesh::get_num_components (computed on first request; 0 after rxm_mesh_compact) */
};
uint64_t rxm_mesh_info(const rxm_mesh* m, int what);
double   rxm_mesh_build_seconds(const rxm_mesh* m, int patcher_only);

/* per-patch host view (PatchInfo, patch_info.h:24-92 + m_h_patches_ltog_*, rxmesh.h:598-640).
 * Pointers stay valid until rxm_mesh_destroy. */
typedef struct {
    uint32_t        patch_id;
    uint32_t        n[3];        /* #V,#E,#F incl. ribbon */
    uint32_t        n_owned[3];
    uint32_t        slot_base[3];
    uint32_t        lin_base[3];
    /* packed != 0: entries are rank-annotated (include/rxmesh_b200/patch_layout.h): EV/FV entry = id | rank << 11,
     * FE entry = dir | edge << 1 | rank << 12; otherwise plain 16-bit ids */
    uint32_t        packed;
    const uint16_t* ev;          /* 2*n[E]: local (larger-id vertex, smaller-id vertex) */
    const uint16_t* fe;          /* 3*n[F]: (local edge << 1) | dir */
    const uint16_t* fv;          /* 3*n[F] */
    const uint16_t* voff_e;      /* n[V]+1: start of every local vertex's edge list (VE / VV) */
    const uint16_t* voff_f;      /* n[V]+1: start of every local vertex's face list (VF) */
    const uint16_t* eoff_f;      /* n[E]+1: start of every local edge's face list (EF) */
    /* one-ring fans of the owned vertices (NULL when the mesh has none): fan_off[n_owned[V]+1], bit 15 =
     * closed fan, low 15 bits = start in fan_v; fan_v = neighbour local vertex ids in oriented order */
    const uint16_t* fan_off;
    const uint16_t* fan_v;
    const uint16_t* fan_f;       /* fan_f[i] = local face between fan_v[i] and fan_v[i+1] (0xFFFF at the end of an open fan) */
    uint32_t        fan_total;
    const uint32_t* owner[3];    /* n[t]-n_owned[t]: (stash slot << 16) | local id in owner */
    const uint32_t* stash;       /* 4*n_stash u32: patch, slot base V, E, F */
    uint32_t        n_stash;
    const uint32_t* ltog[3];     /* n[t] global ids (owned first, each half ascending) */
    /* stored adjacency rows of the OWNED faces / edges (NULL unless the input is edge-manifold): ff = 3 per face, the
     * local faces across edges 0, 1, 2 compacted to the front; ef = 2 per edge, its local faces ascending; 0xFFFF = none */
    const uint16_t* ff;
    const uint16_t* ef;
    /* fan_e[i] = local edge between the fan's vertex and fan_v[i] (NULL without fans): VE in oriented order */
    const uint16_t* fan_e;
    /* ring extension (NULL / 0 when built with RXM_BUILD_NO_RING2): the complete one-ring of every NOT-OWNED vertex within
     * two rings of an owned one, as extended local ids: < n[V] = a vertex of the patch, n[V] + k = ext vertex k (not held by
     * the patch), whose owner record is ext_owner[k] (same encoding as owner[]).
     * r2_idx[n[V] - n_owned[V] + n_ext] (index = extended id - n_owned[V]) = ring index or 0xFFFF;
     * ring r = r2_val[r2_off[r] .. r2_off[r + 1]) */
    const uint16_t* r2_idx;
    const uint16_t* r2_off;
    const uint16_t* r2_val;
    const uint32_t* ext_owner;
    uint32_t        n_r2, n_ext, r2_total;
} rxm_patch_view;
int rxm_mesh_patch(const rxm_mesh* m, uint32_t patch, rxm_patch_view* out);

/* flat host id maps: slot -> global id (0xFFFFFFFF for padding), global -> slot, element -> owner patch
 * (map_to_global / linear_id / get_owner_handle, rxmesh_static.cu:669-685, context.h:218-331) */
const uint32_t* rxm_mesh_slot_to_global(const rxm_mesh* m, int elem);
const uint32_t* rxm_mesh_global_to_slot(const rxm_mesh* m, int elem);
const uint32_t* rxm_mesh_elem_patch(const rxm_mesh* m, int elem);
const uint32_t* rxm_mesh_slot_base(const rxm_mesh* m, int elem); /* [num_patches+1] */
const uint32_t* rxm_mesh_lin_base(const rxm_mesh* m, int elem);  /* [num_patches+1] */
/* global edge list in the reference's numbering: 2*E (larger id, smaller id); 3*F edge ids */
const uint32_t* rxm_mesh_edges(const rxm_mesh* m);
const uint32_t* rxm_mesh_face_edges(const rxm_mesh* m);

/* device copy of rxm_mesh_slot_base (what Attribute::operator() indexes on the device) */
const uint32_t* rxm_mesh_device_slot_base(const rxm_mesh* m, int elem);
/* device copy of rxm_mesh_lin_base: the SoA (tensor) layout indexes data[attr * N + lin_base[patch] + local id],
 * the reference's m_d_linear_id_element_prefix (attribute.h:406-421,623-630) */
const uint32_t* rxm_mesh_device_lin_base(const rxm_mesh* m, int elem);
/* get_context() (rxmesh_static.h): copies the by-value kernel argument (rxm::MeshView of
 * include/rxmesh_b200/patch_layout.h, the counterpart of Context, context.h:15-441) into out_view;
 * out_bytes must equal sizeof(rxm::MeshView). Used by the C++ header shim (include/rxmesh/). */
int rxm_mesh_view(const rxm_mesh* m, void* out_view, uint32_t out_bytes);

/* prepare_launch_box / calc_shared_memory (rxmesh_static.inl:443-841): dynamic shared memory bytes and
 * grid size the query kernel for `op` uses on this mesh. */
int rxm_mesh_launch_box(const rxm_mesh* m, int op, uint32_t* blocks, uint32_t* threads, uint32_t* smem_bytes);

/* ----------------------------------------------------------- attributes ---- */
/* add_{vertex,edge,face}_attribute<T>(name, n, location, layout) (rxmesh_static.h:608-806; attribute.cu:30-83,
 * 531-590). elem_bytes in {1,2,4,8}. Storage holds owned elements only: num_slots(elem) * num_attr values in slot
 * order for AoS / AoSoA (a patch's slots are padded to a multiple of 4); SoA is the reference's tensor layout, a
 * gap-free column-major num_elements x num_attr matrix indexed by linear id (attribute.h:249-261,406-421). */
int rxm_attr_create(rxm_mesh* m, int elem, uint32_t elem_bytes, uint32_t num_attr, int location, int layout,
                    rxm_attr** out);
void     rxm_attr_destroy(rxm_attr* a);                                /* remove_attribute */
int      rxm_attr_release(rxm_attr* a, int location);                  /* Attribute::release(location), attribute.cu:375-390: frees only that side */
void*    rxm_attr_data(rxm_attr* a, int location);                     /* Attribute::data(location) */
uint64_t rxm_attr_count(const rxm_attr* a);                            /* values stored: storage_size(), attribute.h:249 */
int      rxm_attr_reset(rxm_attr* a, const void* value, int location, void* stream); /* attribute.cu:306-357 */
/* attribute.cu:325-373: HOST <-> DEVICE; a target side that is not allocated yet is allocated first */
int      rxm_attr_move(rxm_attr* a, int source, int target, void* stream);
/* attribute.cu:392-500: source / target are location MASKS; every (source side, target side) pair whose bits are set is
 * copied (host->host, device->device, device->host, host->device) */
int      rxm_attr_copy_from(rxm_attr* dst, rxm_attr* src, int source, int target, void* stream);
/* add_vertex_attribute(Verts, name): fill from / read back to an array in GLOBAL element order
 * ([num_elems][num_attr], AoS), the role of rxmesh_static.inl:147-189 and of the
 * for_each_vertex(HOST, map_to_global) read-back loops of the apps. Host buffers; copies + a device
 * permutation kernel run on `stream`. */
int rxm_attr_upload_global(rxm_attr* a, const void* host_global, void* stream);
int rxm_attr_download_global(rxm_attr* a, void* host_global, void* stream);
/* same with DEVICE buffers in global order (no host copy) */
int rxm_attr_from_global_device(rxm_attr* a, const void* dev_global, void* stream);
int rxm_attr_to_global_device(rxm_attr* a, void* dev_global, void* stream);

/* ----------------------------------------------------- fixed-function hot path ---- */
/* The reference's query test kernel (tests/RXMesh_test/query_kernel.cuh:13-46 through
 * Query::dispatch, query.inl:107-156): in(h)=h, out(h,i)=iter[i] as 64-bit handles
 * (patch_id << 32 | local id, handle.h:19-41). in: 1 x u64 on the source element type; out: W x u64. */
int rxm_query_store(rxm_mesh* m, int op, rxm_attr* in, rxm_attr* out, void* stream);
/* roofline "consume" variant: out(s) = sum_i in(iter[i]); in: 1 x fp32 on the op's output element type,
 * out: 1 x fp32 on its source element type. */
int rxm_query_consume(rxm_mesh* m, int op, rxm_attr* in, rxm_attr* out, void* stream);
/* compute_vertex_normal (apps/VertexNormal/vertex_normal_kernel.cuh:10-43), coords/normals: 3 x fp32 vertex attributes
 * in any layout (AoS is the fast path; AoSoA / SoA run through an AoS stand-in -- same for the two calls below).
 * unit_face_normals != 0 -> the Filtering variant (apps/Filtering/filtering_rxmesh_kernel.cuh:15-46). */
int rxm_vertex_normals(rxm_mesh* m, rxm_attr* coords, rxm_attr* normals, int unit_face_normals, void* stream);
/* manual smoothing (apps/Smoothing/manual.h:86-104): `iters` Jacobi steps x <- x - lr * sum_u 2(x - x_u);
 * result in `out` (may not alias `in`). */
int rxm_laplacian_smooth(rxm_mesh* m, rxm_attr* in, rxm_attr* out, double lr, uint32_t iters, void* stream);
/* bilateral filtering (apps/Filtering/filtering_rxmesh.cuh:75-95): `iters` iterations of
 * unit-face vertex normals + bilateral_filtering; result in `out`. */
int rxm_bilateral_filter(rxm_mesh* m, rxm_attr* in, rxm_attr* out, uint32_t iters, void* stream);
/* vertex-iterations of the last rxm_bilateral_filter call whose neighbourhood walk left the patch (+ its ring-2 extension)
 * and ran on the cross-patch path instead (diagnostic; the role of the reference's higher_query_block_dispatcher rounds) */
uint64_t rxm_bilateral_deferred(const rxm_mesh* m);
/* Mean-curvature-flow smoothing, matrix-free conjugate gradients (apps/MCF/mcf_cg_mat_free.h:13-178 driving
 * CGMatFreeAttrSolver, include/rxmesh/matrix/cg_mat_free_attr_solver.h:45-125, over init_B / matvec of
 * apps/MCF/mcf_kernels.cuh:57-205): solves (M + time_step L) X = M X0 for the three coordinates together; L = cotangent
 * (use_uniform_laplace == 0) or uniform Laplacian, M = lumped mixed-Voronoi / valence mass. Stops like
 * IterativeSolver::is_converged (iterative_solver.h:57-63): <R,R> < tol_abs or <R,R> / <R0,R0> < tol_rel, or after max_iter
 * iterations. Reference defaults (apps/MCF/mcf.cu:19-23): time_step 10, tol_abs 1e-6, tol_rel 0, max_iter 100, uniform.
 * The mesh must be closed and edge-manifold (mcf_cg_mat_free.h:39-43, mcf.cu:106-109). coords / out: distinct 3 x fp32
 * vertex attributes. */
typedef struct {
    uint32_t iterations;      /* completed iterations (IterativeSolver::iter_taken) */
    uint32_t converged;       /* 1: the tolerance test held; 0: stopped at max_iter */
    float    start_residual;  /* <R0, R0> */
    float    final_residual;  /* <R, R> at the stop */
} rxm_mcf_info;
int rxm_mcf_solve(rxm_mesh* m, rxm_attr* coords, rxm_attr* out, float time_step, int use_uniform_laplace, uint32_t max_iter,
                  float tol_abs, float tol_rel, rxm_mcf_info* info, void* stream);
/* jacobi != 0: the Jacobi-preconditioned form, mcf_pcg_mat_free (apps/MCF/mcf_cg_mat_free.h:181-254): PCGMatFreeAttrSolver
 * (include/rxmesh/matrix/pcg_mat_free_attr_solver.h:40-140) with precond_matvec (apps/MCF/mcf_kernels.cuh:216-295, out = in /
 * diagonal); the residual the tolerances apply to is then <R, M^-1 R>. jacobi == 0 is rxm_mcf_solve. */
int rxm_mcf_solve_ex(rxm_mesh* m, rxm_attr* coords, rxm_attr* out, float time_step, int use_uniform_laplace, int jacobi,
                     uint32_t max_iter, float tol_abs, float tol_rel, rxm_mcf_info* info, void* stream);
/* Materialise a query as a CSR over attribute SLOTS on the device (cached in the mesh): off[num_slots(src)+1],
 * val[nnz] = owner slots of the neighbours, lists grouped by patch. The k-ring consumer (bilateral filtering)
 * traverses this instead of re-running whole-patch queries per foreign patch like the reference's
 * higher_query_block_dispatcher (kernels/query_dispatcher.cuh:445-565). Pointers are device pointers owned by m. */
int rxm_query_csr(rxm_mesh* m, int op, uint32_t** dev_off, uint32_t** dev_val, uint64_t* nnz, void* stream);
/* get_boundary_vertices (rxmesh_static.h; kernels/boundary.cuh:11-44): flag: 1 x u32 vertex attribute, 1 = boundary */
int rxm_boundary_vertices(rxm_mesh* m, rxm_attr* flag, void* stream);

/* host-buffer entry points (what bench.py times as `e2e`): H2D + kernel(s) + D2H inside the call,
 * arrays in GLOBAL vertex order, [V][3] fp32. */
int rxm_vertex_normals_host(rxm_mesh* m, const float* coords, float* normals, void* stream);
int rxm_laplacian_smooth_host(rxm_mesh* m, const float* coords, float* out, double lr, uint32_t iters, void* stream);
/* in: [num output-type elements] fp32, out: [num source-type elements] fp32, global order */
int rxm_query_consume_host(rxm_mesh* m, int op, const float* in, float* out, void* stream);

/* ------------------------------------------------------------- multi-GPU (new) ---- */
/* A rank's mesh = its own ("real") patches plus the ghost patches that own its ribbon elements. Kernels run
 * on the real range only; ghost patches' attribute slots are filled by the halo exchange. */
int rxm_mesh_set_active_patches(rxm_mesh* m, uint32_t first, uint32_t count);
/* sorted attribute slots, owned by patches OUTSIDE [first, first+count), that the patches inside reference
 * through their owner tables (= what a halo exchange must fill). *out is malloc'd: release with rxm_free. */
int  rxm_mesh_halo_slots(const rxm_mesh* m, int elem, uint32_t first, uint32_t count, uint32_t** out, uint64_t* n);
void rxm_free(void* p);
int  rxm_memcpy_d2h(void* host, const void* dev, uint64_t bytes); /* read back a device array the library owns */
/* pack / unpack rows of an AoS attribute through a device index list (device buffers) */
int rxm_attr_gather_slots(rxm_attr* a, const uint32_t* dev_idx, uint64_t n, void* dev_out, void* stream);
int rxm_attr_scatter_slots(rxm_attr* a, const uint32_t* dev_idx, uint64_t n, const void* dev_in, void* stream);
/* direct NVLink P2P halo push: remote_data[remote_idx[i]] = a[local_idx[i]] with remote_data a peer-mapped
 * pointer (rxm_ipc_open) to the neighbour rank's attribute storage */
int rxm_attr_push_slots(rxm_attr* a, const uint32_t* dev_local_idx, void* remote_data, const uint32_t* dev_remote_idx,
                        uint64_t n, void* stream);
/* Laplacian step fused with the halo exchange (new: the reference is single-GPU): the kernel stores the rows mirrored
 * on neighbouring ranks into their ghost slots over NVLink P2P and raises a flag there; patches that read ghost slots
 * wait for the neighbours' flags.  Setup: create (allocates the flag words), export / exchange rxm_fused_halo_flags and
 * both attributes over cudaIpc, then set.  See rxmesh_b200/distributed.py: FusedHalo. */
/* test hook: the chunk frontiers of the pipelined host-buffer calls (rxm_capi.cu: build_pipe_plan), host only */
int rxm_mesh_pipe_plan(rxm_mesh* m, uint32_t chunks, uint32_t* pb, uint64_t* up_hi, uint64_t* down_lo, uint32_t* need);
typedef struct rxm_fused_halo rxm_fused_halo;
int   rxm_fused_halo_create(rxm_mesh* m, uint32_t npeers, rxm_fused_halo** out);
void* rxm_fused_halo_flags(rxm_fused_halo* h);
/* blocks of a fused step that wait for neighbour flags or push rows (known after rxm_fused_halo_set) */
uint32_t rxm_fused_halo_sync_blocks(const rxm_fused_halo* h);
int   rxm_fused_halo_set(rxm_fused_halo* h, const uint32_t* push_off, const uint32_t* push_lid_peer, const uint32_t* push_slot,
                         uint64_t n_push, void* const* peer_attr_a, void* const* peer_attr_b, void* const* peer_flag);
void  rxm_fused_halo_destroy(rxm_fused_halo* h);
int   rxm_laplacian_smooth_fused(rxm_mesh* m, rxm_attr* in, rxm_attr* out, double lr, rxm_fused_halo* h, int out_is_b,
                                 uint32_t step, void* stream);
/* ---- single-process multi-GPU mode: one host process, one shard per device (rxmesh_b200/csrc/rxm_multi.cu) ----
 * What SURVEY.md 8(e) asks of RXMeshStatic ("device-list / shard options", rxmesh_static.h:61-100 has a single device): the
 * mesh is cut into contiguous patch-id ranges, one per device; every device holds its range plus two vertex-rings of ghost
 * faces; mirrored rows travel by direct NVLink peer stores from the kernel that computes them (k_laplacian_fan2<true>).
 * devices == NULL: host-only plan (no CUDA call; the compute entry points then fail with RXM_ERR_CUDA).
 * face_patch == NULL: built-in Lloyd patcher with locality-ordered patch ids. */
typedef struct rxm_multi rxm_multi;
int  rxm_multi_create(const uint32_t* fv, uint32_t num_faces, const uint32_t* face_patch, uint32_t patch_size, const int* devices,
                      int num_shards, int num_threads, rxm_multi** out);
void rxm_multi_destroy(rxm_multi* m);
/* what: 0 shards, 1 patches, 2 ghost vertex rows refreshed per exchange (all shards), 3 vertices, 4 faces; per shard:
 * 10 real patches, 11 faces held, 12 vertices owned by its real patches, 13 ghost rows received, 14 rows pushed, 15 neighbours */
uint64_t  rxm_multi_info(const rxm_multi* m, int what, int shard);
rxm_mesh* rxm_multi_shard_mesh(rxm_multi* m, int shard);
/* manual smoothing (apps/Smoothing/manual.h:86-104) over all devices: host buffers in GLOBAL vertex order, [V][3] fp32;
 * one fused compute + halo kernel per device and iteration, no host synchronisation inside the loop */
int rxm_multi_laplacian_smooth(rxm_multi* m, const float* coords, float* out, double lr, uint32_t iters);
/* compute_vertex_normal (apps/VertexNormal/vertex_normal_kernel.cuh:10-43) over all devices */
int rxm_multi_vertex_normals(rxm_multi* m, const float* coords, float* normals);

int rxm_ipc_export(void* dev_ptr, void* handle64);        /* cudaIpcGetMemHandle */
int rxm_ipc_open(const void* handle64, void** dev_ptr);   /* cudaIpcOpenMemHandle */
int rxm_ipc_close(void* dev_ptr);

/* ReduceHandle (reduce_handle.h:64-166, reduce_handle.cu:53-156) over the OWNED elements of fp32 attributes.
 * kind: 0 dot(a,b), 1 sum of squares (norm2 = sqrt), 2 sum, 3 min, 4 max, 5 arg-min, 6 arg-max.
 * attribute_id = 0xFFFFFFFF -> all attributes. out_handle (arg ops): 64-bit owner handle. */
int rxm_attr_reduce(rxm_attr* a, rxm_attr* b, int kind, uint32_t attribute_id, double* out_value, uint64_t* out_handle,
                    void* stream);

/* Saved patchings: the reference's Patcher::save / Patcher(filename) (patcher/patcher.h:154-182; the
 * `patcher_file` constructor argument, rxmesh_static.h:61-66), cereal PortableBinary layout. header = patch_size,
 * num_patches, num_vertices, num_edges, num_faces, num_seeds, max_num_patches, num_components, num_lloyd_run;
 * vec = face_patch, vertex_patch, edge_patch, patches_val, patches_offset, ribbon_ext_val, ribbon_ext_offset. */
typedef struct {
    uint32_t  header[9];
    uint32_t* vec[7];
    uint64_t  len[7];
    float     patching_time_ms;
} rxm_patcher_file;
int  rxm_patcher_file_read(const char* path, rxm_patcher_file* out); /* vec[0] is the face_patch for rxm_mesh_create */
void rxm_patcher_file_free(rxm_patcher_file* pf);
int  rxm_mesh_save_patcher_file(const rxm_mesh* m, const char* path);

/* Asynchronous host calls: with rxm_set_async(1) the *_host entry points return once their H2D copy, kernels
 * and D2H copy are ENQUEUED on `stream` (the host buffers must be pinned and stay valid); results are ready
 * after rxm_stream_sync(stream). Calls issued on different streams overlap their copies and kernels (every
 * entry point owns its device staging buffers). Default: synchronous. */
void rxm_set_async(int on);
int  rxm_stream_sync(void* stream);

/* kernels launched by this library so far (bench.py "gpu_launches") */
uint64_t rxm_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* RXMESH_B200_H */
