// rxm_capi.cu -- implementation of the C ABI declared in include/rxmesh_b200.h.
// No CPU fallback: every compute entry point needs the mesh on a CUDA device and
// fails with RXM_ERR_CUDA otherwise.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "../../include/rxmesh_b200.h"
#include "mesh_builder.h"
#include "rxm_kernels.h"

using namespace rxm;

static thread_local std::string g_err;

static int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}

#define CU(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(RXM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));            \
    } while (0)

struct rxm_mesh
{
    HostMesh     h;
    bool         on_device = false;
    PatchDesc*   d_desc    = nullptr;
    uint8_t*     d_topo    = nullptr;
    uint32_t*    d_slot_base[3] = {nullptr, nullptr, nullptr};
    uint32_t*    d_lin_base[3]  = {nullptr, nullptr, nullptr};  // [P+1] linear-id prefixes (SoA attribute indexing)
    uint32_t*    d_s2g[3]       = {nullptr, nullptr, nullptr};
    uint64_t     topo_bytes   = 0;
    uint32_t     active_first = 0, active_count = 0;  // patches the kernels run on (a shard's real patches)
    MeshView     view{};
    KernelLimits lim{};
    // scratch for the host-buffer entry points and multi-iteration drivers
    void*     d_stage[32]       = {};  // [0] generic; host entry points own dedicated in/out buffers
    size_t    d_stage_bytes[32] = {};
    rxm_attr* scratch[8]    = {};  // [0..3] drivers / host entry points, [4..7] AoS stand-ins for other layouts
    struct Csr
    {
        uint32_t *off = nullptr, *val = nullptr;
        uint64_t  nnz = 0;
    };
    Csr       csr[16];      // materialised queries (rxm_query_csr), indexed by op
    uint32_t* d_flag = nullptr;
    uint32_t* d_fan_base = nullptr;  // [P] prefix of the patches' fan entries, rounded up to 4 (rxm_mcf_solve: slices of W)
    uint64_t  fan_entries = 0;
    void*     d_mcf_ws = nullptr;  // workspace of rxm_mcf_solve, kept between solves (an MCF flow solves once per time step)
    uint64_t  mcf_ws_bytes = 0;
    uint64_t  bilateral_deferred = 0;  // vertices the last rxm_bilateral_filter call sent down the cross-patch path
    rxm_attr* scratch1[32]  = {};  // per query op: [2*op] input, [2*op+1] output of rxm_query_consume_host
    // ---- chunked upload / compute / download pipeline of the host-buffer entry points (see pipelined_host_call) ----
    struct PipePlan
    {
        uint32_t              K = 0;       // chunks (contiguous patch ranges); 0 = pipeline unavailable
        std::vector<uint32_t> pb;          // [K+1] patch boundaries
        std::vector<uint64_t> up_hi[3];    // [K] per element type: ids < up_hi[c] cover every element OWNED by patches < pb[c+1]
        std::vector<uint64_t> down_lo[3];  // [K+1] ids < down_lo[c] are owned by patches < pb[c] only; [K] = #elements
        std::vector<uint32_t> need;        // [K] chunk q reads ribbon values owned by patches of chunks <= need[q]
    } plan;
    struct Pipe
    {
        cudaStream_t             h2d = nullptr, d2h = nullptr;
        cudaEvent_t              start = nullptr, done = nullptr;
        std::vector<cudaEvent_t> up, comp;
    } pipe[32];
};

struct rxm_attr
{
    rxm_mesh* m;
    int       elem;
    uint32_t  elem_bytes, nattr;
    int       layout;
    uint64_t  count;  // T values in the storage: num_slots * nattr; SoA (tensor layout): num_elems * nattr
    void*     h = nullptr;
    void*     d = nullptr;
    bool      h_pinned = false;
};

template <typename T>
static AttrView<T> view_of(rxm_attr* a)
{
    AttrView<T> v;
    v.data      = (T*)a->d;
    v.slot_base = a->m->d_slot_base[a->elem];
    v.num_slots = a->m->h.num_slots[a->elem];
    v.nattr     = a->nattr;
    v.layout    = (uint32_t)a->layout;
    v.lin_base  = a->m->d_lin_base[a->elem];
    v.num_elems = a->m->h.num_elems[a->elem];
    return v;
}

// slot / linear-id prefixes of the patch range [p0, p0 + np) on the device
static SlotMap slot_map(const rxm_mesh* m, int t, uint32_t p0, uint32_t np)
{
    return SlotMap{m->d_slot_base[t] + p0, m->d_lin_base[t] + p0, np, m->h.num_slots[t], m->h.num_elems[t]};
}
static SlotMap slot_map(const rxm_mesh* m, int t)
{
    return slot_map(m, t, 0, m->h.num_patches);
}

extern "C" {

const char* rxm_last_error(void)
{
    return g_err.c_str();
}

// for the other translation units of the library (rxm_multi.cu): record an error, return its code
int rxm_set_last_error(int code, const char* msg)
{
    return fail(code, msg ? msg : "");
}

const char* rxm_version(void)
{
    return "rxmesh_b200 0.1 (sm_100a)";
}

int rxm_init(int device)
{
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(RXM_ERR_UNSUPPORTED, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                             "; this library is built for sm_100a only");
    return RXM_OK;
}

int rxm_mesh_create(const uint32_t* fv, uint32_t num_faces, const uint32_t* face_patch, uint32_t patch_size,
                    int num_threads, rxm_mesh** out)
{
    return rxm_mesh_create_ex(fv, num_faces, face_patch, patch_size, num_threads, getenv("RXM_NO_RING2") ? RXM_BUILD_NO_RING2 : 0u, out);
}

int rxm_mesh_create_ex(const uint32_t* fv, uint32_t num_faces, const uint32_t* face_patch, uint32_t patch_size,
                       int num_threads, uint32_t flags, rxm_mesh** out)
{
    if (!fv || !out) return fail(RXM_ERR_INVALID, "rxm_mesh_create: null argument");
    rxm_mesh* m = new (std::nothrow) rxm_mesh();
    if (!m) return fail(RXM_ERR_INVALID, "rxm_mesh_create: out of memory");
    BuildOptions opt;
    opt.patch_size  = patch_size ? patch_size : 512;
    opt.num_threads = num_threads;
    opt.verbose     = getenv("RXM_VERBOSE") != nullptr;
    opt.force_wide  = getenv("RXM_FORCE_WIDE") != nullptr;  // tests: exercise the atomic (wide-format) kernels
    opt.no_fans     = getenv("RXM_NO_FANS") != nullptr;     // tests: exercise the generic (transpose) kernels
    opt.no_ring2    = (flags & RXM_BUILD_NO_RING2) != 0;
    opt.reorder_patches = !(flags & RXM_BUILD_NO_PATCH_REORDER) && getenv("RXM_NO_PATCH_REORDER") == nullptr;
    std::string e;
    try {
        e = build_mesh(fv, num_faces, face_patch, opt, m->h);
    } catch (const std::exception& ex) {
        e = std::string("rxm_mesh_create: ") + ex.what();
    }
    if (!e.empty()) {
        delete m;
        return fail(RXM_ERR_INVALID, e);
    }
    for (int t = 0; t < 3; ++t) {
        m->lim.max_n[t]         = m->h.max_per_patch[t];
        m->lim.max_owned[t]     = m->h.max_owned_per_patch[t];
        m->lim.max_not_owned[t] = m->h.max_not_owned[t];
    }
    m->topo_bytes                  = m->h.topo.size();
    m->lim.max_stash               = m->h.max_stash;
    m->lim.max_face_adjacent_faces = m->h.max_face_adjacent_faces;
    m->lim.max_fan_total           = m->h.max_fan_total;
    m->lim.max_ext                 = m->h.max_ext;
    m->lim.max_r2                  = m->h.max_r2;
    m->lim.max_r2_total            = m->h.max_r2_total;
    *out                           = m;
    return RXM_OK;
}

// Frontiers that let a host-buffer call overlap its H2D copy, its kernels and its D2H copy: patches are cut into K
// contiguous chunks; chunk c's owned slots can be filled once the global-order input prefix [0, up_hi[c]) is on the
// device, its kernel can run once the chunks holding its ribbon owners (<= need[c]) are filled, and the output prefix
// [0, down_lo[c+1]) is final once chunks <= c are done.  Always correct; how much overlaps depends on how well patch
// order follows global element order (row-major tiles of a grid: almost perfectly; Lloyd patches of a scrambled mesh:
// the first piece is most of the array and the call degenerates to upload -> compute -> download).
static void build_pipe_plan(rxm_mesh* m, uint32_t K)
{
    const HostMesh& h = m->h;
    auto&           P = m->plan;
    P.K               = 0;
    if (K < 2 || h.num_patches < 2 * K || h.topo.empty()) return;
    if (m->active_count && m->active_count != h.num_patches) return;  // shards with ghost patches: plain path
    P.pb.resize(K + 1);
    for (uint32_t c = 0; c <= K; ++c)
        P.pb[c] = (uint32_t)((uint64_t)h.num_patches * c / K);
    for (int t = 0; t < 3; ++t) {
        std::vector<uint64_t> omin(h.num_patches, UINT64_MAX), omax(h.num_patches, 0);
#pragma omp parallel for schedule(static)
        for (int64_t p = 0; p < (int64_t)h.num_patches; ++p) {
            const uint32_t b = h.slot_base[t][p], no = h.desc[p].n_owned[t];
            for (uint32_t i = 0; i < no; ++i) {
                const uint64_t g = h.slot_to_global[t][b + i];
                omin[p] = std::min(omin[p], g), omax[p] = std::max(omax[p], g + 1);
            }
        }
        P.up_hi[t].assign(K, 0), P.down_lo[t].assign(K + 1, h.num_elems[t]);
        uint64_t run = 0;
        for (uint32_t c = 0; c < K; ++c) {
            for (uint32_t p = P.pb[c]; p < P.pb[c + 1]; ++p)
                run = std::max(run, omax[p]);
            P.up_hi[t][c] = run;
        }
        P.up_hi[t][K - 1] = h.num_elems[t];
        uint64_t lo = h.num_elems[t];
        for (uint32_t c = K; c-- > 0;) {
            for (uint32_t p = P.pb[c]; p < P.pb[c + 1]; ++p)
                lo = std::min(lo, omin[p]);
            P.down_lo[t][c] = lo;
        }
        P.down_lo[t][0] = 0;
    }
    P.need.assign(K, 0);
    for (uint32_t c = 0; c < K; ++c) {
        uint32_t mx = P.pb[c + 1] - 1;
        for (uint32_t p = P.pb[c]; p < P.pb[c + 1]; ++p) {
            const PatchDesc&  D  = h.desc[p];
            const StashEntry* st = reinterpret_cast<const StashEntry*>(h.topo.data() + D.topo_off + D.off_stash());
            for (uint32_t i = 0; i < D.n_stash; ++i)
                mx = std::max(mx, st[i].patch);
        }
        uint32_t q = c;
        while (q + 1 < K && mx >= P.pb[q + 1]) ++q;
        P.need[c] = q;
    }
    P.K = K;
}

// test hook: the pipeline frontiers for `chunks` chunks, computed on the host (no device needed).  Returns the number of
// chunks actually planned (0: mesh too small); pb[chunks+1], up_hi[3][chunks], down_lo[3][chunks+1], need[chunks].
int rxm_mesh_pipe_plan(rxm_mesh* m, uint32_t chunks, uint32_t* pb, uint64_t* up_hi, uint64_t* down_lo, uint32_t* need)
{
    if (!m || !pb || !up_hi || !down_lo || !need) return fail(RXM_ERR_INVALID, "rxm_mesh_pipe_plan: null argument");
    const auto saved = m->plan;
    build_pipe_plan(m, chunks);
    const uint32_t K = m->plan.K;
    if (K) {
        memcpy(pb, m->plan.pb.data(), 4 * (size_t)(K + 1));
        memcpy(need, m->plan.need.data(), 4 * (size_t)K);
        for (int t = 0; t < 3; ++t) {
            memcpy(up_hi + (size_t)t * K, m->plan.up_hi[t].data(), 8 * (size_t)K);
            memcpy(down_lo + (size_t)t * (K + 1), m->plan.down_lo[t].data(), 8 * (size_t)(K + 1));
        }
    }
    m->plan = saved;
    return (int)K;
}

int rxm_mesh_to_device(rxm_mesh* m)
{
    if (!m) return fail(RXM_ERR_INVALID, "rxm_mesh_to_device: null mesh");
    if (m->on_device) return RXM_OK;
    const HostMesh& h = m->h;
    CU(cudaMalloc(&m->d_desc, h.desc.size() * sizeof(PatchDesc)));
    CU(cudaMemcpy(m->d_desc, h.desc.data(), h.desc.size() * sizeof(PatchDesc), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&m->d_topo, h.topo.size()));
    CU(cudaMemcpy(m->d_topo, h.topo.data(), h.topo.size(), cudaMemcpyHostToDevice));
    for (int t = 0; t < 3; ++t) {
        CU(cudaMalloc(&m->d_slot_base[t], h.slot_base[t].size() * 4));
        CU(cudaMemcpy(m->d_slot_base[t], h.slot_base[t].data(), h.slot_base[t].size() * 4, cudaMemcpyHostToDevice));
        CU(cudaMalloc(&m->d_lin_base[t], h.lin_base[t].size() * 4));
        CU(cudaMemcpy(m->d_lin_base[t], h.lin_base[t].data(), h.lin_base[t].size() * 4, cudaMemcpyHostToDevice));
        CU(cudaMalloc(&m->d_s2g[t], std::max<size_t>(h.slot_to_global[t].size(), 1) * 4));
        CU(cudaMemcpy(m->d_s2g[t], h.slot_to_global[t].data(), h.slot_to_global[t].size() * 4, cudaMemcpyHostToDevice));
    }
    m->view.desc        = m->d_desc;
    m->view.topo        = m->d_topo;
    m->view.num_patches = h.num_patches;
    if (m->active_count) {
        m->view.desc        = m->d_desc + m->active_first;
        m->view.num_patches = m->active_count;
    }
    m->view.packed      = h.packed ? 1u : 0u;
    m->view.fans        = h.fans ? 1u : 0u;
    m->view.edge_manifold = h.max_edge_incident_faces <= 2 ? 1u : 0u;
    m->view.ring2         = h.ring2 ? 1u : 0u;
    for (int t = 0; t < 3; ++t) {
        m->view.num_slots[t]       = h.num_slots[t];
        m->view.num_elems[t]       = h.num_elems[t];
        m->view.patch_slot_base[t] = m->d_slot_base[t];
    }
    m->on_device = true;
    const char* env = getenv("RXM_PIPE_CHUNKS");
    build_pipe_plan(m, env ? (uint32_t)atoi(env) : 8u);
    return RXM_OK;
}

int rxm_mesh_compact(rxm_mesh* m)
{
    if (!m) return fail(RXM_ERR_INVALID, "rxm_mesh_compact: null mesh");
    HostMesh& h = m->h;
    for (int t = 0; t < 3; ++t) {
        U32Buf().swap(h.ltog[t]);
        std::vector<uint64_t>().swap(h.ltog_off[t]);
    }
    U32Buf().swap(h.ev);
    U32Buf().swap(h.fe);
    if (m->on_device) ByteBuf().swap(h.topo);  // the device holds the patch store
    return RXM_OK;
}

void rxm_mesh_destroy(rxm_mesh* m)
{
    if (!m) return;
    for (auto& pp : m->pipe) {
        for (auto e : pp.up) cudaEventDestroy(e);
        for (auto e : pp.comp) cudaEventDestroy(e);
        if (pp.start) cudaEventDestroy(pp.start);
        if (pp.done) cudaEventDestroy(pp.done);
        if (pp.h2d) cudaStreamDestroy(pp.h2d);
        if (pp.d2h) cudaStreamDestroy(pp.d2h);
    }
    for (auto*& s : m->scratch)
        if (s) {
            rxm_attr_destroy(s);
            s = nullptr;
        }
    for (auto*& s : m->scratch1)
        if (s) {
            rxm_attr_destroy(s);
            s = nullptr;
        }
    if (m->on_device) {
        cudaFree(m->d_desc);
        cudaFree(m->d_topo);
        for (int t = 0; t < 3; ++t) {
            cudaFree(m->d_slot_base[t]);
            cudaFree(m->d_lin_base[t]);
            cudaFree(m->d_s2g[t]);
        }
        for (void* b : m->d_stage)
            if (b) cudaFree(b);
        if (m->d_flag) cudaFree(m->d_flag);
        if (m->d_fan_base) cudaFree(m->d_fan_base);
        if (m->d_mcf_ws) cudaFree(m->d_mcf_ws);
        for (auto& c : m->csr) {
            if (c.off) cudaFree(c.off);
            if (c.val) cudaFree(c.val);
        }
    }
    delete m;
}


// connected components of the input mesh (union-find over the edge list; the reference counts them while patching)
static uint64_t count_components(const rxm::HostMesh& h)
{
    if (h.ev.empty()) return 0;
    std::vector<uint32_t> parent(h.num_elems[0]);
    for (uint32_t v = 0; v < h.num_elems[0]; ++v)
        parent[v] = v;
    auto find = [&](uint32_t v) {
        while (parent[v] != v) {
            parent[v] = parent[parent[v]];
            v         = parent[v];
        }
        return v;
    };
    for (size_t e = 0; e + 1 < h.ev.size(); e += 2) {
        const uint32_t a = find(h.ev[e]), b = find(h.ev[e + 1]);
        if (a != b) parent[std::max(a, b)] = std::min(a, b);
    }
    uint64_t n = 0;
    for (uint32_t v = 0; v < h.num_elems[0]; ++v)
        n += parent[v] == v;
    return n;
}

uint64_t rxm_mesh_info(const rxm_mesh* m, int what)
{
    if (!m) return 0;
    const HostMesh& h = m->h;
    switch (what) {
        case RXM_INFO_NUM_VERTICES: return h.num_elems[ELEM_V];
        case RXM_INFO_NUM_EDGES: return h.num_elems[ELEM_E];
        case RXM_INFO_NUM_FACES: return h.num_elems[ELEM_F];
        case RXM_INFO_NUM_PATCHES: return h.num_patches;
        case RXM_INFO_PATCH_SIZE: return h.patch_size;
        case RXM_INFO_MAX_VALENCE: return h.max_valence;
        case RXM_INFO_MAX_EDGE_INCIDENT_FACES: return h.max_edge_incident_faces;
        case RXM_INFO_MAX_FACE_ADJACENT_FACES: return h.max_face_adjacent_faces;
        case RXM_INFO_IS_CLOSED: return h.is_closed;
        case RXM_INFO_IS_EDGE_MANIFOLD: return h.is_edge_manifold;
        case RXM_INFO_MAX_VERTICES_PER_PATCH: return h.max_per_patch[ELEM_V];
        case RXM_INFO_MAX_EDGES_PER_PATCH: return h.max_per_patch[ELEM_E];
        case RXM_INFO_MAX_FACES_PER_PATCH: return h.max_per_patch[ELEM_F];
        case RXM_INFO_NUM_SLOTS_V: return h.num_slots[ELEM_V];
        case RXM_INFO_NUM_SLOTS_E: return h.num_slots[ELEM_E];
        case RXM_INFO_NUM_SLOTS_F: return h.num_slots[ELEM_F];
        case RXM_INFO_TOPO_BYTES: return m->topo_bytes;
        case RXM_INFO_TOTAL_LOCAL_V: return h.total_local[ELEM_V];
        case RXM_INFO_TOTAL_LOCAL_E: return h.total_local[ELEM_E];
        case RXM_INFO_TOTAL_LOCAL_F: return h.total_local[ELEM_F];
        case RXM_INFO_MAX_STASH: return h.max_stash;
        case RXM_INFO_ON_DEVICE: return m->on_device;
        case RXM_INFO_PACKED: return h.packed;
        case RXM_INFO_FANS: return h.fans;
        case RXM_INFO_RING2: return h.ring2;
        case RXM_INFO_LLOYD_RUNS: return h.lloyd_runs;
        case RXM_INFO_NUM_COMPONENTS: return count_components(h);
        default: return 0;
    }
}

double rxm_mesh_build_seconds(const rxm_mesh* m, int patcher_only)
{
    return m ? (patcher_only ? m->h.patcher_seconds : m->h.build_seconds) : 0.0;
}

int rxm_mesh_patch(const rxm_mesh* m, uint32_t p, rxm_patch_view* o)
{
    if (!m || !o || p >= m->h.num_patches) return fail(RXM_ERR_INVALID, "rxm_mesh_patch: bad argument");
    if (m->h.topo.empty()) return fail(RXM_ERR_INVALID, "rxm_mesh_patch: host patch store was released (rxm_mesh_compact)");
    const PatchDesc& D = m->h.desc[p];
    const uint8_t*   B = m->h.topo.data() + D.topo_off;
    o->patch_id        = D.patch_id;
    for (int t = 0; t < 3; ++t) {
        o->n[t]         = D.n[t];
        o->n_owned[t]   = D.n_owned[t];
        o->slot_base[t] = D.slot_base[t];
        o->lin_base[t]  = D.lin_base[t];
        o->owner[t]     = reinterpret_cast<const uint32_t*>(B + D.off_own(t));
        o->ltog[t]      = m->h.ltog[t].empty() ? nullptr : m->h.ltog[t].data() + m->h.ltog_off[t][p];
    }
    o->ev      = reinterpret_cast<const uint16_t*>(B + D.off_ev());
    o->fe      = reinterpret_cast<const uint16_t*>(B + D.off_fe());
    o->fv      = reinterpret_cast<const uint16_t*>(B + D.off_fv());
    o->voff_e  = reinterpret_cast<const uint16_t*>(B + D.off_voff_e());
    o->voff_f  = reinterpret_cast<const uint16_t*>(B + D.off_voff_f());
    o->eoff_f  = reinterpret_cast<const uint16_t*>(B + D.off_eoff_f());
    o->packed  = (D.flags & FLAG_PACKED) ? 1u : 0u;
    o->fan_off = (D.flags & FLAG_FANS) ? reinterpret_cast<const uint16_t*>(B + D.off_fanoff()) : nullptr;
    o->fan_v   = (D.flags & FLAG_FANS) ? reinterpret_cast<const uint16_t*>(B + D.off_fanv()) : nullptr;
    o->fan_f   = (D.flags & FLAG_FANS) ? reinterpret_cast<const uint16_t*>(B + D.off_fanf()) : nullptr;
    o->ff      = (D.flags & FLAG_FF) ? reinterpret_cast<const uint16_t*>(B + D.off_ff()) : nullptr;
    o->ef      = (D.flags & FLAG_FF) ? reinterpret_cast<const uint16_t*>(B + D.off_ef()) : nullptr;
    o->fan_total = D.fan_total;
    o->fan_e   = (D.flags & FLAG_FANS) ? reinterpret_cast<const uint16_t*>(B + D.off_fane()) : nullptr;
    const bool r2 = (D.flags & FLAG_RING2) != 0;
    o->r2_idx  = r2 ? reinterpret_cast<const uint16_t*>(B + D.o_r2idx) : nullptr;
    o->r2_off  = r2 ? reinterpret_cast<const uint16_t*>(B + D.o_r2off) : nullptr;
    o->r2_val  = r2 ? reinterpret_cast<const uint16_t*>(B + D.o_r2val) : nullptr;
    o->ext_owner = r2 ? reinterpret_cast<const uint32_t*>(B + D.o_ext) : nullptr;
    o->n_r2 = D.n_r2, o->n_ext = D.n_ext, o->r2_total = D.r2_total;
    o->stash   = reinterpret_cast<const uint32_t*>(B + D.off_stash());
    o->n_stash = D.n_stash;
    return RXM_OK;
}

const uint32_t* rxm_mesh_slot_to_global(const rxm_mesh* m, int t)
{
    return (m && t >= 0 && t < 3) ? m->h.slot_to_global[t].data() : nullptr;
}
const uint32_t* rxm_mesh_global_to_slot(const rxm_mesh* m, int t)
{
    return (m && t >= 0 && t < 3) ? m->h.global_to_slot[t].data() : nullptr;
}
const uint32_t* rxm_mesh_elem_patch(const rxm_mesh* m, int t)
{
    return (m && t >= 0 && t < 3) ? m->h.elem_patch[t].data() : nullptr;
}
const uint32_t* rxm_mesh_slot_base(const rxm_mesh* m, int t)
{
    return (m && t >= 0 && t < 3) ? m->h.slot_base[t].data() : nullptr;
}
const uint32_t* rxm_mesh_lin_base(const rxm_mesh* m, int t)
{
    return (m && t >= 0 && t < 3) ? m->h.lin_base[t].data() : nullptr;
}
// NULL once rxm_mesh_compact released the global edge arrays (callers check)
const uint32_t* rxm_mesh_edges(const rxm_mesh* m)
{
    return (m && !m->h.ev.empty()) ? m->h.ev.data() : nullptr;
}
const uint32_t* rxm_mesh_face_edges(const rxm_mesh* m)
{
    return (m && !m->h.fe.empty()) ? m->h.fe.data() : nullptr;
}

const uint32_t* rxm_mesh_device_slot_base(const rxm_mesh* m, int t)
{
    return (m && m->on_device && t >= 0 && t < 3) ? m->d_slot_base[t] : nullptr;
}
const uint32_t* rxm_mesh_device_lin_base(const rxm_mesh* m, int t)
{
    return (m && m->on_device && t >= 0 && t < 3) ? m->d_lin_base[t] : nullptr;
}

int rxm_mesh_view(const rxm_mesh* m, void* out_view, uint32_t out_bytes)
{
    if (!m || !out_view || out_bytes != sizeof(MeshView)) return fail(RXM_ERR_INVALID, "rxm_mesh_view: bad argument");
    if (!m->on_device) return fail(RXM_ERR_CUDA, "rxm_mesh_view: mesh is not on a CUDA device");
    memcpy(out_view, &m->view, sizeof(MeshView));
    return RXM_OK;
}

int rxm_mesh_launch_box(const rxm_mesh* m, int op, uint32_t* blocks, uint32_t* threads, uint32_t* smem_bytes)
{
    if (!m) return fail(RXM_ERR_INVALID, "rxm_mesh_launch_box: null mesh");
    auto           r16 = [](uint32_t x) { return (x + 15u) & ~15u; };
    const auto&    L   = m->lim;
    const uint32_t ev = r16(4 * L.max_n[ELEM_E]), f3 = r16(6 * L.max_n[ELEM_F]);
    uint32_t       b = 0;
    switch (op) {
        case RXM_OP_EV: b = ev; break;
        case RXM_OP_FV: case RXM_OP_FE: b = f3; break;
        case RXM_OP_VV: case RXM_OP_VE: b = 2 * ev + r16(4 * (L.max_n[ELEM_V] + 1)); break;
        case RXM_OP_VF: b = 2 * f3 + r16(4 * (L.max_n[ELEM_V] + 1)); break;
        case RXM_OP_EF: b = 2 * f3 + r16(4 * (L.max_n[ELEM_E] + 1)); break;
        case RXM_OP_FF: b = 2 * f3 + r16(4 * (L.max_n[ELEM_E] + 1)) + r16(4 * (L.max_n[ELEM_F] + 1)) + 2 * f3; break;
        case RXM_OP_EE: case RXM_OP_EVDIAMOND:  // rxmesh_static.inl:519-525: edge-manifold input only
            if (m->h.max_edge_incident_faces > 2)
                return fail(RXM_ERR_UNSUPPORTED, "Op::EVDiamond and Op::EE only work on edge-manifold meshes");
            b = f3 + r16(8 * L.max_n[ELEM_E]) + ev;
            break;
        default: return fail(RXM_ERR_INVALID, "rxm_mesh_launch_box: unknown op");
    }
    const uint32_t mo = std::max(L.max_not_owned[0], std::max(L.max_not_owned[1], L.max_not_owned[2]));
    b += r16(4 * mo) + 16 * L.max_stash;
    if (blocks) *blocks = m->h.num_patches;
    if (threads) *threads = 256;
    if (smem_bytes) *smem_bytes = b;
    return RXM_OK;
}

// ------------------------------------------------------------------ attributes
int rxm_attr_create(rxm_mesh* m, int elem, uint32_t elem_bytes, uint32_t nattr, int location, int layout,
                    rxm_attr** out)
{
    if (!m || !out || elem < 0 || elem > 2 || nattr == 0)
        return fail(RXM_ERR_INVALID, "rxm_attr_create: bad argument");
    if (elem_bytes != 1 && elem_bytes != 2 && elem_bytes != 4 && elem_bytes != 8)
        return fail(RXM_ERR_INVALID, "rxm_attr_create: elem_bytes must be 1, 2, 4 or 8");
    if (layout != RXM_AOS && layout != RXM_AOSOA && layout != RXM_SOA)
        return fail(RXM_ERR_INVALID, "rxm_attr_create: unknown layout");
    rxm_attr* a   = new rxm_attr();
    a->m          = m;
    a->elem       = elem;
    a->elem_bytes = elem_bytes;
    a->nattr      = nattr;
    a->layout     = layout;
    a->count      = (uint64_t)(layout == RXM_SOA ? m->h.num_elems[elem] : m->h.num_slots[elem]) * nattr;
    const size_t bytes = std::max<size_t>(a->count * elem_bytes, 16);
    if (location & RXM_DEVICE) {
        if (!m->on_device) {
            delete a;
            return fail(RXM_ERR_CUDA, "rxm_attr_create: DEVICE location requested but the mesh is not on a device");
        }
        cudaError_t e = cudaMalloc(&a->d, bytes);
        // defined contents: the reference's own tests read components they never wrote and expect zero
        // (test_attribute.cu:56-79).  The fill runs on the legacy default stream, which non-blocking streams (PyTorch's, the
        // pipeline's) do not wait for: it must be complete before the caller can launch anything that writes the attribute.
        if (e == cudaSuccess) e = cudaMemset(a->d, 0, bytes);
        if (e == cudaSuccess) e = cudaStreamSynchronize(0);
        if (e != cudaSuccess) {
            if (a->d) cudaFree(a->d);
            delete a;
            return fail(RXM_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
        }
    }
    if (location & RXM_HOST) {
        if (m->on_device && cudaMallocHost(&a->h, bytes) == cudaSuccess) {
            a->h_pinned = true;
        } else {
            cudaGetLastError();
            a->h = malloc(bytes);
        }
        if (!a->h) {
            rxm_attr_destroy(a);
            return fail(RXM_ERR_INVALID, "rxm_attr_create: host allocation failed");
        }
        memset(a->h, 0, bytes);
    }
    *out = a;
    return RXM_OK;
}

void rxm_attr_destroy(rxm_attr* a)
{
    if (!a) return;
    if (a->d) cudaFree(a->d);
    if (a->h) {
        if (a->h_pinned)
            cudaFreeHost(a->h);
        else
            free(a->h);
    }
    delete a;
}

void* rxm_attr_data(rxm_attr* a, int location)
{
    if (!a) return nullptr;
    return (location & RXM_DEVICE) ? a->d : a->h;
}

uint64_t rxm_attr_count(const rxm_attr* a)
{
    return a ? a->count : 0;
}

int rxm_attr_reset(rxm_attr* a, const void* value, int location, void* stream)
{
    if (!a || !value) return fail(RXM_ERR_INVALID, "rxm_attr_reset: null argument");
    if ((location & RXM_DEVICE) && a->d) {
        cudaError_t e = launch_fill(a->d, a->count, a->elem_bytes, value, (cudaStream_t)stream);
        if (e != cudaSuccess) return fail(RXM_ERR_CUDA, std::string("fill: ") + cudaGetErrorString(e));
    }
    if ((location & RXM_HOST) && a->h) {
        uint8_t* p = (uint8_t*)a->h;
        for (uint64_t i = 0; i < a->count; ++i)
            memcpy(p + i * a->elem_bytes, value, a->elem_bytes);
    }
    return RXM_OK;
}

// allocate the side(s) of `location` that the attribute does not have yet (Attribute::allocate, attribute.cu:531-590)
static int attr_allocate(rxm_attr* a, int location)
{
    rxm_mesh*    m     = a->m;
    const size_t bytes = std::max<size_t>(a->count * a->elem_bytes, 16);
    if ((location & RXM_DEVICE) && !a->d) {
        if (!m->on_device) return fail(RXM_ERR_CUDA, "attribute: DEVICE location requested but the mesh is not on a device");
        cudaError_t e = cudaMalloc(&a->d, bytes);
        if (e == cudaSuccess) e = cudaMemset(a->d, 0, bytes);
        if (e == cudaSuccess) e = cudaStreamSynchronize(0);
        if (e != cudaSuccess) {
            if (a->d) cudaFree(a->d);
            a->d = nullptr;
            return fail(RXM_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
        }
    }
    if ((location & RXM_HOST) && !a->h) {
        if (m->on_device && cudaMallocHost(&a->h, bytes) == cudaSuccess) {
            a->h_pinned = true;
        } else {
            cudaGetLastError();
            a->h        = malloc(bytes);
            a->h_pinned = false;
        }
        if (!a->h) return fail(RXM_ERR_INVALID, "attribute: host allocation failed");
        memset(a->h, 0, bytes);
    }
    return RXM_OK;
}

// Attribute::release(location) (attribute.cu:375-390): frees only the requested side(s)
int rxm_attr_release(rxm_attr* a, int location)
{
    if (!a) return fail(RXM_ERR_INVALID, "rxm_attr_release: null attribute");
    if ((location & RXM_DEVICE) && a->d) {
        cudaFree(a->d);
        a->d = nullptr;
    }
    if ((location & RXM_HOST) && a->h) {
        if (a->h_pinned)
            cudaFreeHost(a->h);
        else
            free(a->h);
        a->h = nullptr, a->h_pinned = false;
    }
    return RXM_OK;
}

int rxm_attr_move(rxm_attr* a, int source, int target, void* stream)
{
    if (!a) return fail(RXM_ERR_INVALID, "rxm_attr_move: null attribute");
    if (source == target) return RXM_OK;  // the reference warns and returns (attribute.cu:331-338)
    if (!((source == RXM_HOST && target == RXM_DEVICE) || (source == RXM_DEVICE && target == RXM_HOST)))
        return fail(RXM_ERR_INVALID, "rxm_attr_move: source/target must be HOST or DEVICE");
    if (!(source == RXM_HOST ? a->h : a->d)) return fail(RXM_ERR_INVALID, "rxm_attr_move: the source side is not allocated");
    int rc = attr_allocate(a, target);  // the reference allocates a missing target before moving (attribute.cu:348-353)
    if (rc) return rc;
    const size_t bytes = a->count * a->elem_bytes;
    if (source == RXM_HOST)
        CU(cudaMemcpyAsync(a->d, a->h, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    else {
        CU(cudaMemcpyAsync(a->h, a->d, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
        CU(cudaStreamSynchronize((cudaStream_t)stream));
    }
    return RXM_OK;
}

// Attribute::copy_from (attribute.cu:392-500): every (source side, target side) pair whose bits are set in the two flags is
// copied, in the reference's order host->host, device->device, device->host, host->device; e.g. (LOCATION_ALL,
// LOCATION_ALL) refreshes both copies, (HOST, LOCATION_ALL) fans the host copy out to both sides.
int rxm_attr_copy_from(rxm_attr* dst, rxm_attr* src, int source, int target, void* stream)
{
    if (!dst || !src) return fail(RXM_ERR_INVALID, "rxm_attr_copy_from: null attribute");
    if (dst->count != src->count || dst->elem_bytes != src->elem_bytes || dst->layout != src->layout ||
        dst->elem != src->elem)
        return fail(RXM_ERR_INVALID, "rxm_attr_copy_from: attributes differ in shape");
    if ((source & RXM_LOCATION_ALL) == RXM_LOCATION_ALL && (target & RXM_LOCATION_ALL) != RXM_LOCATION_ALL)
        return fail(RXM_ERR_INVALID, "rxm_attr_copy_from: invalid configuration (source LOCATION_ALL needs target LOCATION_ALL)");
    const size_t bytes = dst->count * dst->elem_bytes;
    cudaStream_t S     = (cudaStream_t)stream;
    int          done = 0, missing = 0;
    bool         d2h = false;
    auto pair = [&](int sbit, int tbit) -> int {
        if (!(source & sbit) || !(target & tbit)) return RXM_OK;
        void* s = sbit == RXM_DEVICE ? src->d : src->h;
        void* d = tbit == RXM_DEVICE ? dst->d : dst->h;
        if (!s || !d) {
            ++missing;
            return RXM_OK;
        }
        if (sbit == RXM_HOST && tbit == RXM_HOST)
            memcpy(d, s, bytes);
        else
            CU(cudaMemcpyAsync(d, s, bytes, cudaMemcpyDefault, S));
        d2h |= (sbit == RXM_DEVICE && tbit == RXM_HOST);
        ++done;
        return RXM_OK;
    };
    int rc;
    if ((rc = pair(RXM_HOST, RXM_HOST)) || (rc = pair(RXM_DEVICE, RXM_DEVICE)) || (rc = pair(RXM_DEVICE, RXM_HOST)) ||
        (rc = pair(RXM_HOST, RXM_DEVICE)))
        return rc;
    if (d2h) CU(cudaStreamSynchronize(S));
    if (!done) return fail(RXM_ERR_INVALID, missing ? "rxm_attr_copy_from: side not allocated" : "rxm_attr_copy_from: no location selected");
    return RXM_OK;
}

static int  g_async_host_calls = 0;  // rxm_set_async

static int ensure_stage(rxm_mesh* m, size_t bytes, int k = 0)
{
    if (m->d_stage_bytes[k] >= bytes) return RXM_OK;
    if (m->d_stage[k]) cudaFree(m->d_stage[k]);
    m->d_stage[k] = nullptr, m->d_stage_bytes[k] = 0;
    CU(cudaMalloc(&m->d_stage[k], bytes));
    m->d_stage_bytes[k] = bytes;
    return RXM_OK;
}

// upload / download through a dedicated staging buffer `k` (concurrent host calls on different streams
// must not share one); `sync` = wait for the result before returning
static int upload_via(rxm_attr* a, const void* host_global, void* stream, int k)
{
    rxm_mesh*    m     = a->m;
    const size_t bytes = (size_t)m->h.num_elems[a->elem] * a->nattr * a->elem_bytes;
    int          rc    = ensure_stage(m, bytes, k);
    if (rc) return rc;
    CU(cudaMemcpyAsync(m->d_stage[k], host_global, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return rxm_attr_from_global_device(a, m->d_stage[k], stream);
}
static int download_via(rxm_attr* a, void* host_global, void* stream, int k, bool sync)
{
    rxm_mesh*    m     = a->m;
    const size_t bytes = (size_t)m->h.num_elems[a->elem] * a->nattr * a->elem_bytes;
    int          rc    = ensure_stage(m, bytes, k);
    if (rc) return rc;
    if ((rc = rxm_attr_to_global_device(a, m->d_stage[k], stream))) return rc;
    CU(cudaMemcpyAsync(host_global, m->d_stage[k], bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    if (sync) CU(cudaStreamSynchronize((cudaStream_t)stream));
    return RXM_OK;
}

int rxm_attr_from_global_device(rxm_attr* a, const void* dev_global, void* stream)
{
    if (!a || !a->d || !dev_global) return fail(RXM_ERR_INVALID, "rxm_attr_from_global_device: bad argument");
    rxm_mesh* m = a->m;
    cudaError_t e = launch_permute_to_slots(dev_global, a->d, m->d_s2g[a->elem], slot_map(m, a->elem), a->elem_bytes,
                                            a->nattr, a->layout, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(RXM_ERR_CUDA, std::string("permute: ") + cudaGetErrorString(e));
    return RXM_OK;
}

int rxm_attr_to_global_device(rxm_attr* a, void* dev_global, void* stream)
{
    if (!a || !a->d || !dev_global) return fail(RXM_ERR_INVALID, "rxm_attr_to_global_device: bad argument");
    rxm_mesh* m = a->m;
    cudaError_t e = launch_permute_to_global(a->d, dev_global, m->d_s2g[a->elem], slot_map(m, a->elem), a->elem_bytes,
                                             a->nattr, a->layout, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(RXM_ERR_CUDA, std::string("permute: ") + cudaGetErrorString(e));
    return RXM_OK;
}

// host-side permutation for HOST-only attributes
static void host_permute(rxm_attr* a, void* global, bool to_slots)
{
    const HostMesh& h  = a->m->h;
    const int       t  = a->elem;
    const uint32_t  eb = a->elem_bytes, na = a->nattr;
    AttrView<uint8_t> v{nullptr, h.slot_base[t].data(), h.num_slots[t], na, (uint32_t)a->layout, h.lin_base[t].data(),
                        h.num_elems[t]};
    uint8_t* S = (uint8_t*)a->h;
    uint8_t* G = (uint8_t*)global;
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < (int64_t)h.num_patches; ++p) {
        const uint32_t b = h.slot_base[t][p], cap = h.slot_base[t][p + 1] - b, lb = h.lin_base[t][p];
        const uint32_t n = a->layout == RXM_SOA ? h.lin_base[t][p + 1] - lb : cap;  // SoA: no padding slots
        for (uint32_t lid = 0; lid < n; ++lid) {
            const uint32_t g = h.slot_to_global[t][b + lid];
            for (uint32_t k = 0; k < na; ++k) {
                const uint64_t si = v.index_known(b, cap, lb, lid, k);
                if (to_slots) {
                    if (g != INVALID32_)
                        memcpy(S + si * eb, G + ((uint64_t)g * na + k) * eb, eb);
                    else
                        memset(S + si * eb, 0, eb);
                } else if (g != INVALID32_)
                    memcpy(G + ((uint64_t)g * na + k) * eb, S + si * eb, eb);
            }
        }
    }
}

int rxm_attr_upload_global(rxm_attr* a, const void* host_global, void* stream)
{
    if (!a || !host_global) return fail(RXM_ERR_INVALID, "rxm_attr_upload_global: bad argument");
    if (a->h) host_permute(a, const_cast<void*>(host_global), true);
    if (a->d) {
        rxm_mesh*    m     = a->m;
        const size_t bytes = (size_t)m->h.num_elems[a->elem] * a->nattr * a->elem_bytes;
        (void)bytes;
        return upload_via(a, host_global, stream, 0);
    }
    return RXM_OK;
}

int rxm_attr_download_global(rxm_attr* a, void* host_global, void* stream)
{
    if (!a || !host_global) return fail(RXM_ERR_INVALID, "rxm_attr_download_global: bad argument");
    if (a->d) {
        rxm_mesh*    m     = a->m;
        const size_t bytes = (size_t)m->h.num_elems[a->elem] * a->nattr * a->elem_bytes;
        (void)bytes, (void)m;
        return download_via(a, host_global, stream, 0, true);
    }
    if (a->h) {
        host_permute(a, host_global, false);
        return RXM_OK;
    }
    return fail(RXM_ERR_INVALID, "rxm_attr_download_global: attribute has no storage");
}

// ------------------------------------------------------------------ kernels
static int check_dev(rxm_mesh* m, const char* who)
{
    if (!m) return fail(RXM_ERR_INVALID, std::string(who) + ": null mesh");
    if (!m->on_device)
        return fail(RXM_ERR_CUDA, std::string(who) + ": mesh is not on a CUDA device (call rxm_mesh_to_device); "
                                                     "there is no CPU fallback");
    return RXM_OK;
}

static int op_src(int op)
{
    switch (op) {
        case RXM_OP_VV: case RXM_OP_VE: case RXM_OP_VF: return RXM_V;
        case RXM_OP_EV: case RXM_OP_EF: case RXM_OP_EE: case RXM_OP_EVDIAMOND: return RXM_E;
        case RXM_OP_FV: case RXM_OP_FE: case RXM_OP_FF: return RXM_F;
        default: return -1;
    }
}
static int op_dst(int op)
{
    switch (op) {
        case RXM_OP_VV: case RXM_OP_EV: case RXM_OP_FV: case RXM_OP_EVDIAMOND: return RXM_V;
        case RXM_OP_VE: case RXM_OP_FE: case RXM_OP_EE: return RXM_E;
        case RXM_OP_VF: case RXM_OP_EF: case RXM_OP_FF: return RXM_F;
        default: return -1;
    }
}

static int kernel_status(cudaError_t e, const char* why, const char* who)
{
    if (e == cudaSuccess) return RXM_OK;
    if (why) return fail(RXM_ERR_UNSUPPORTED, std::string(who) + ": " + why);
    return fail(RXM_ERR_CUDA, std::string(who) + ": " + cudaGetErrorString(e));
}

int rxm_query_store(rxm_mesh* m, int op, rxm_attr* in, rxm_attr* out, void* stream)
{
    int rc = check_dev(m, "rxm_query_store");
    if (rc) return rc;
    if (op_src(op) < 0) return fail(RXM_ERR_INVALID, "rxm_query_store: unknown op");
    if (!in || !out || !in->d || !out->d || in->elem != op_src(op) || out->elem != op_src(op) ||
        in->elem_bytes != 8 || out->elem_bytes != 8 || in->nattr != 1)
        return fail(RXM_ERR_INVALID, "rxm_query_store: in/out must be device u64 attributes on the op's source type");
    const char* why = nullptr;
    cudaError_t e   = launch_query_store(op, m->view, m->lim, view_of<uint64_t>(in), view_of<uint64_t>(out),
                                         (cudaStream_t)stream, &why);
    return kernel_status(e, why, "rxm_query_store");
}

int rxm_query_consume(rxm_mesh* m, int op, rxm_attr* in, rxm_attr* out, void* stream)
{
    int rc = check_dev(m, "rxm_query_consume");
    if (rc) return rc;
    if (op_src(op) < 0) return fail(RXM_ERR_INVALID, "rxm_query_consume: unknown op");
    if (!in || !out || !in->d || !out->d || in->elem != op_dst(op) || out->elem != op_src(op) ||
        in->elem_bytes != 4 || out->elem_bytes != 4 || in->nattr != 1 || out->nattr != 1)
        return fail(RXM_ERR_INVALID, "rxm_query_consume: in = 1 x fp32 on the output type, out = 1 x fp32 on the source type");
    if (in->layout == RXM_SOA || out->layout == RXM_SOA)  // the kernels stream the patches' slot slices
        return fail(RXM_ERR_INVALID, "rxm_query_consume: SoA (tensor-layout) attributes are stored by linear id; use AoS or AoSoA");
    const char* why = nullptr;
    cudaError_t e   = launch_query_consume(op, m->view, m->lim, view_of<float>(in), view_of<float>(out),
                                           (cudaStream_t)stream, &why);
    return kernel_status(e, why, "rxm_query_consume");
}

static int get_scratch(rxm_mesh* m, int idx, rxm_attr** out)
{
    if (!m->scratch[idx]) {
        int rc = rxm_attr_create(m, RXM_V, 4, 3, RXM_DEVICE, RXM_AOS, &m->scratch[idx]);
        if (rc) return rc;
    }
    *out = m->scratch[idx];
    return RXM_OK;
}

// The fixed-function kernels read and write AoS xyz.  An attribute in another layout (the reference's default is
// AoSoA) is served through an AoS stand-in: relayout in, run, relayout out -- two extra streaming passes, same results.
static bool is_vec3(rxm_attr* a)
{
    return a && a->d && a->elem == RXM_V && a->elem_bytes == 4 && a->nattr == 3;
}
static int aos_standin(rxm_mesh* m, rxm_attr* a, int idx, bool load, void* stream, rxm_attr** out)
{
    if (a->layout == RXM_AOS) {
        *out = a;
        return RXM_OK;
    }
    int rc = get_scratch(m, idx, out);
    if (rc) return rc;
    if (load) {
        cudaError_t e = launch_relayout(a->d, (*out)->d, slot_map(m, RXM_V), 3, (uint32_t)a->layout, RXM_AOS, (cudaStream_t)stream);
        if (e != cudaSuccess) return fail(RXM_ERR_CUDA, std::string("relayout: ") + cudaGetErrorString(e));
    }
    return RXM_OK;
}
static int aos_writeback(rxm_mesh* m, rxm_attr* a, rxm_attr* standin, void* stream)
{
    if (a == standin) return RXM_OK;
    cudaError_t e = launch_relayout(standin->d, a->d, slot_map(m, RXM_V), 3, RXM_AOS, (uint32_t)a->layout, (cudaStream_t)stream);
    return e == cudaSuccess ? RXM_OK : fail(RXM_ERR_CUDA, std::string("relayout: ") + cudaGetErrorString(e));
}

static bool is_vec3_aos(rxm_attr* a)
{
    return a && a->d && a->elem == RXM_V && a->elem_bytes == 4 && a->nattr == 3 && a->layout == RXM_AOS;
}

int rxm_vertex_normals(rxm_mesh* m, rxm_attr* coords, rxm_attr* normals, int unit, void* stream)
{
    int rc = check_dev(m, "rxm_vertex_normals");
    if (rc) return rc;
    if (!is_vec3(coords) || !is_vec3(normals))
        return fail(RXM_ERR_INVALID, "rxm_vertex_normals: coords/normals must be device 3 x fp32 vertex attributes");
    rxm_attr *x, *n;
    if ((rc = aos_standin(m, coords, 4, true, stream, &x)) || (rc = aos_standin(m, normals, 5, false, stream, &n))) return rc;
    const char* why = nullptr;
    cudaError_t e   = launch_vertex_normals(m->view, m->lim, (const float*)x->d, (float*)n->d, unit, (cudaStream_t)stream, &why);
    if ((rc = kernel_status(e, why, "rxm_vertex_normals"))) return rc;
    return aos_writeback(m, normals, n, stream);
}

int rxm_laplacian_smooth(rxm_mesh* m, rxm_attr* in, rxm_attr* out, double lr, uint32_t iters, void* stream)
{
    int rc = check_dev(m, "rxm_laplacian_smooth");
    if (rc) return rc;
    if (!is_vec3(in) || !is_vec3(out) || in == out)
        return fail(RXM_ERR_INVALID, "rxm_laplacian_smooth: in/out must be distinct device 3 x fp32 vertex attributes");
    if (in->layout != RXM_AOS || out->layout != RXM_AOS) {
        rxm_attr *x, *y;
        if ((rc = aos_standin(m, in, 4, true, stream, &x)) || (rc = aos_standin(m, out, 5, false, stream, &y))) return rc;
        if ((rc = rxm_laplacian_smooth(m, x, y, lr, iters, stream))) return rc;
        return aos_writeback(m, out, y, stream);
    }
    if (iters == 0) return rxm_attr_copy_from(out, in, RXM_DEVICE, RXM_DEVICE, stream);
    rxm_attr* tmp = nullptr;
    if (iters > 1) {
        rc = get_scratch(m, 0, &tmp);
        if (rc) return rc;
    }
    const float* src = (const float*)in->d;
    for (uint32_t k = 1; k <= iters; ++k) {
        float*      dst = ((iters - k) % 2 == 0) ? (float*)out->d : (float*)tmp->d;
        const char* why = nullptr;
        cudaError_t e   = launch_laplacian_step(m->view, m->lim, src, dst, lr, (cudaStream_t)stream, &why);
        if (e != cudaSuccess) return kernel_status(e, why, "rxm_laplacian_smooth");
        src = dst;
    }
    return RXM_OK;
}

// Mean-curvature flow, matrix-free CG (apps/MCF/mcf_cg_mat_free.h:13-178 with CGMatFreeAttrSolver,
// matrix/cg_mat_free_attr_solver.h:45-125): see rxm_mcf.cu.
int rxm_mcf_solve(rxm_mesh* m, rxm_attr* coords, rxm_attr* out, float time_step, int use_uniform_laplace, uint32_t max_iter,
                  float tol_abs, float tol_rel, rxm_mcf_info* info, void* stream)
{
    return rxm_mcf_solve_ex(m, coords, out, time_step, use_uniform_laplace, 0, max_iter, tol_abs, tol_rel, info, stream);
}

// jacobi != 0: mcf_pcg_mat_free (apps/MCF/mcf_cg_mat_free.h:181-254): PCGMatFreeAttrSolver with precond_matvec
int rxm_mcf_solve_ex(rxm_mesh* m, rxm_attr* coords, rxm_attr* out, float time_step, int use_uniform_laplace, int jacobi,
                     uint32_t max_iter, float tol_abs, float tol_rel, rxm_mcf_info* info, void* stream)
{
    int rc = check_dev(m, "rxm_mcf_solve");
    if (rc) return rc;
    if (!is_vec3(coords) || !is_vec3(out) || coords == out)
        return fail(RXM_ERR_INVALID, "rxm_mcf_solve: coords/out must be distinct device 3 x fp32 vertex attributes");
    // mcf_cg_mat_free.h:39-43 ("only takes watertight/closed mesh without boundaries"), mcf.cu:106-109 (edge-manifold)
    if (!m->h.is_closed || !m->h.is_edge_manifold)
        return fail(RXM_ERR_INVALID, "rxm_mcf_solve: MCF needs a closed, edge-manifold mesh");
    if (!m->h.fans)
        return fail(RXM_ERR_UNSUPPORTED, "rxm_mcf_solve: the mesh stores no one-ring fans (inconsistently oriented input)");
    if (m->active_count && (m->active_first != 0 || m->active_count != m->h.num_patches))
        return fail(RXM_ERR_UNSUPPORTED, "rxm_mcf_solve: runs on a whole mesh, not on a shard's active patch range");
    if (coords->layout != RXM_AOS || out->layout != RXM_AOS) {
        rxm_attr *x, *y;
        if ((rc = aos_standin(m, coords, 4, true, stream, &x)) || (rc = aos_standin(m, out, 5, false, stream, &y))) return rc;
        if ((rc = rxm_mcf_solve_ex(m, x, y, time_step, use_uniform_laplace, jacobi, max_iter, tol_abs, tol_rel, info, stream))) return rc;
        return aos_writeback(m, out, y, stream);
    }
    cudaStream_t   st      = (cudaStream_t)stream;
    const bool     uniform = use_uniform_laplace != 0, precond = jacobi != 0;
    const uint32_t P       = m->h.num_patches;
    if (!m->d_fan_base) {  // where every patch's slice of the weight array starts
        std::vector<uint32_t> fb(P);
        uint64_t              acc = 0;
        for (uint32_t p = 0; p < P; ++p) {
            fb[p] = (uint32_t)acc;
            acc += (m->h.desc[p].fan_total + 3u) & ~3u;
        }
        if (acc > 0xFFFFFFF0ull) return fail(RXM_ERR_UNSUPPORTED, "rxm_mcf_solve: more than 2^32 fan entries");
        CU(cudaMalloc((void**)&m->d_fan_base, 4 * (size_t)std::max<uint32_t>(P, 1u)));
        CU(cudaMemcpy(m->d_fan_base, fb.data(), 4 * (size_t)P, cudaMemcpyHostToDevice));
        m->fan_entries = acc;
    }
    // one allocation: state | partials | diag | W | R | S | P0 | P1   (X is `out` itself)
    const uint64_t slots = m->h.num_slots[ELEM_V];
    auto           al    = [](uint64_t b) { return (b + 255ull) & ~255ull; };
    const uint64_t b_state = al(sizeof(McfState)), b_part = al(8ull * (P + mcf_update_grid())), b_diag = al(4ull * slots),
                   b_w = uniform ? 0ull : al(4ull * (m->fan_entries + 4ull)), b_vec = al(12ull * slots);
    const uint64_t need = b_state + b_part + b_diag + b_w + 4ull * b_vec;
    if (m->mcf_ws_bytes < need) {
        if (m->d_mcf_ws) cudaFree(m->d_mcf_ws);
        m->d_mcf_ws = nullptr, m->mcf_ws_bytes = 0;
        CU(cudaMalloc(&m->d_mcf_ws, need));
        m->mcf_ws_bytes = need;
    }
    uint8_t* base = (uint8_t*)m->d_mcf_ws;
    McfBuffers B{};
    uint8_t*   q       = base;
    B.state            = (McfState*)q, q += b_state;
    B.partials         = (double*)q, q += b_part;
    B.diag             = (float*)q, q += b_diag;
    B.W                = uniform ? nullptr : (float*)q, q += b_w;
    B.R                = (float*)q, q += b_vec;
    B.S                = (float*)q, q += b_vec;
    B.P[0]             = (float*)q, q += b_vec;
    B.P[1]             = (float*)q, q += b_vec;
    B.X                = (float*)out->d;
    B.fan_base         = m->d_fan_base;
    B.partials_split   = P;
    McfState    hs{};
    const char* why = nullptr;
    cudaError_t e   = cudaMemsetAsync(base, 0, b_state + b_part, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out->d, coords->d, 12ull * slots, cudaMemcpyDeviceToDevice, st);  // X = X0
    if (e == cudaSuccess) e = launch_mcf_setup(m->view, m->lim, (const float*)coords->d, B, uniform, precond, time_step, st, &why);
    // iterations are queued in batches with no host synchronisation inside; the device-side state says when to stop (a
    // converged solve turns the rest of a batch into kernels that return at their first instruction)
    const uint32_t batch = 8;
    uint32_t       it    = 0;
    while (e == cudaSuccess) {
        for (uint32_t k = 0; k < batch && it < max_iter && e == cudaSuccess; ++k, ++it)
            e = launch_mcf_iteration(m->view, m->lim, B, it, uniform, precond, time_step, tol_abs, tol_rel, max_iter, st, &why);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&hs, B.state, sizeof(McfState), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess || hs.converged || it >= max_iter) break;
    }
    if (e != cudaSuccess) return kernel_status(e, why, "rxm_mcf_solve");
    if (info) {
        info->iterations     = hs.iters;
        info->converged      = hs.converged;
        info->start_residual = (float)hs.start;
        info->final_residual = (float)hs.final_res;
    }
    return RXM_OK;
}

// number of list entries a patch contributes to the CSR of `op` (owned sources only), from the stored offsets
static uint32_t patch_nnz(const HostMesh& h, uint32_t p, int op)
{
    const PatchDesc& D = h.desc[p];
    const uint8_t*   B = h.topo.data() + D.topo_off;
    auto u16 = [&](uint32_t off, uint32_t i) { return (uint32_t) reinterpret_cast<const uint16_t*>(B + off)[i]; };
    switch (op) {
        case RXM_OP_VV: case RXM_OP_VE: return u16(D.off_voff_e(), D.n_owned[ELEM_V]);
        case RXM_OP_VF: return u16(D.off_voff_f(), D.n_owned[ELEM_V]);
        case RXM_OP_EF: return u16(D.off_eoff_f(), D.n_owned[ELEM_E]);
        case RXM_OP_EV: return 2u * D.n_owned[ELEM_E];
        case RXM_OP_FV: case RXM_OP_FE: return 3u * D.n_owned[ELEM_F];
        case RXM_OP_FF: {
            const uint16_t* fe = reinterpret_cast<const uint16_t*>(B + D.off_fe());
            const uint32_t  em = (D.flags & FLAG_PACKED) ? PK_ID_MASK : 0x7FFFu;
            uint32_t        k  = 0;
            for (uint32_t i = 0; i < 3u * D.n_owned[ELEM_F]; ++i) {
                const uint32_t e = (fe[i] >> 1) & em;
                k += u16(D.off_eoff_f(), e + 1) - u16(D.off_eoff_f(), e) - 1;
            }
            return k;
        }
        default: return 0;
    }
}

int rxm_query_csr(rxm_mesh* m, int op, uint32_t** dev_off, uint32_t** dev_val, uint64_t* nnz, void* stream)
{
    int rc = check_dev(m, "rxm_query_csr");
    if (rc) return rc;
    if (op_src(op) < 0 || !dev_off || !dev_val || !nnz) return fail(RXM_ERR_INVALID, "rxm_query_csr: bad argument");
    if (m->active_count && m->active_count != m->h.num_patches)
        return fail(RXM_ERR_UNSUPPORTED, "rxm_query_csr: not available on a shard (active patch range set)");
    if (op == RXM_OP_EE || op == RXM_OP_EVDIAMOND)
        return fail(RXM_ERR_UNSUPPORTED, "rxm_query_csr: Op::EE / Op::EVDiamond have no CSR form (fixed-width results; use rxm_query_store)");
    auto& C = m->csr[op];
    if (!C.off && m->h.topo.empty())
        return fail(RXM_ERR_INVALID, "rxm_query_csr: host patch store was released (rxm_mesh_compact)");
    if (!C.off) {
        const HostMesh&       h = m->h;
        std::vector<uint32_t> pno(h.num_patches + 1, 0);
        uint64_t              run = 0;
        for (uint32_t p = 0; p < h.num_patches; ++p) {
            pno[p] = (uint32_t)run;
            run += patch_nnz(h, p, op);
        }
        if (run > 0xFFFFFFFFull) return fail(RXM_ERR_UNSUPPORTED, "rxm_query_csr: more than 2^32 entries");
        pno[h.num_patches] = (uint32_t)run;
        const uint32_t ns  = h.num_slots[op_src(op)];
        // built into local pointers and published to the cache only when the kernel and the stream sync succeeded: a failed
        // build must not leave a half-initialised CSR behind for the next call to return
        uint32_t *d_pno = nullptr, *off = nullptr, *val = nullptr;
        auto      build = [&]() -> int {
            CU(cudaMalloc(&d_pno, pno.size() * 4));
            CU(cudaMemcpyAsync(d_pno, pno.data(), pno.size() * 4, cudaMemcpyHostToDevice, (cudaStream_t)stream));
            CU(cudaMalloc(&off, ((size_t)ns + 1) * 4));
            CU(cudaMalloc(&val, std::max<size_t>(run, 1) * 4));
            const char* why = nullptr;
            cudaError_t e   = launch_query_csr(op, m->view, m->lim, d_pno, off, val, (cudaStream_t)stream, &why);
            if (e != cudaSuccess) return kernel_status(e, why, "rxm_query_csr");
            const uint32_t total = (uint32_t)run;
            CU(cudaMemcpyAsync(off + ns, &total, 4, cudaMemcpyHostToDevice, (cudaStream_t)stream));
            CU(cudaStreamSynchronize((cudaStream_t)stream));
            return RXM_OK;
        };
        rc = build();
        if (d_pno) cudaFree(d_pno);
        if (rc) {
            if (off) cudaFree(off);
            if (val) cudaFree(val);
            return rc;
        }
        C.off = off, C.val = val, C.nnz = run;
    }
    *dev_off = C.off, *dev_val = C.val, *nnz = C.nnz;
    return RXM_OK;
}

int rxm_bilateral_filter(rxm_mesh* m, rxm_attr* in, rxm_attr* out, uint32_t iters, void* stream)
{
    int rc = check_dev(m, "rxm_bilateral_filter");
    if (rc) return rc;
    if (!is_vec3(in) || !is_vec3(out) || in == out)
        return fail(RXM_ERR_INVALID, "rxm_bilateral_filter: in/out must be distinct device 3 x fp32 vertex attributes");
    if (in->layout != RXM_AOS || out->layout != RXM_AOS) {
        rxm_attr *x, *y;
        if ((rc = aos_standin(m, in, 4, true, stream, &x)) || (rc = aos_standin(m, out, 5, false, stream, &y))) return rc;
        if ((rc = rxm_bilateral_filter(m, x, y, iters, stream))) return rc;
        return aos_writeback(m, out, y, stream);
    }
    if (iters == 0) return rxm_attr_copy_from(out, in, RXM_DEVICE, RXM_DEVICE, stream);
    uint32_t *off = nullptr, *val = nullptr;
    uint64_t  nnz = 0;
    if ((rc = rxm_query_csr(m, RXM_OP_VV, &off, &val, &nnz, stream))) return rc;
    rxm_attr *nrm = nullptr, *tmp = nullptr;
    // default: one patch-local kernel per iteration with the normals fused in (k_bilateral_patch); meshes without one-ring
    // fans (non-manifold / inconsistently oriented input) and RXM_BILATERAL_CSR=1 keep the two-kernel path over the CSR
    const bool patch_local = m->view.fans && !getenv("RXM_BILATERAL_CSR");
    if (!patch_local && (rc = get_scratch(m, 3, &nrm))) return rc;
    if (iters > 1 && (rc = get_scratch(m, 0, &tmp))) return rc;
    if (!m->d_flag) CU(cudaMalloc(&m->d_flag, 16));
    CU(cudaMemsetAsync(m->d_flag, 0, 16, (cudaStream_t)stream));
    if (patch_local && (rc = ensure_stage(m, 20ull * m->h.num_slots[ELEM_V] + 16, 31))) return rc;  // deferred work list
    rxm_attr* src = in;
    for (uint32_t k = 1; k <= iters; ++k) {
        rxm_attr* dst = ((iters - k) % 2 == 0) ? out : tmp;
        if (patch_local) {
            const char* why = nullptr;
            cudaError_t e   = launch_bilateral_patch(m->view, m->lim, off, val, (const float*)src->d, (float*)dst->d, m->d_flag,
                                                     m->d_stage[31], k - 1, (cudaStream_t)stream, &why);
            if ((rc = kernel_status(e, why, "rxm_bilateral_filter"))) return rc;
        } else {
            // filtering_rxmesh.cuh:75-95: vertex normals of the current positions, then the filter
            if ((rc = rxm_vertex_normals(m, src, nrm, 1, stream))) return rc;
            cudaError_t e = launch_bilateral_step(off, val, m->h.num_slots[ELEM_V], (const float*)src->d, (const float*)nrm->d,
                                                  (float*)dst->d, m->d_flag, (cudaStream_t)stream);
            if (e != cudaSuccess) return fail(RXM_ERR_CUDA, std::string("bilateral: ") + cudaGetErrorString(e));
        }
        src = dst;
    }
    uint32_t flag[2] = {0, 0};
    CU(cudaMemcpyAsync(flag, m->d_flag, 8, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    m->bilateral_deferred = flag[1];
    if (flag[0]) return fail(RXM_ERR_UNSUPPORTED, "rxm_bilateral_filter: a neighbourhood exceeded 80 vertices (maxVVSize of the reference)");
    return RXM_OK;
}

uint64_t rxm_bilateral_deferred(const rxm_mesh* m)
{
    return m ? m->bilateral_deferred : 0;
}

int rxm_boundary_vertices(rxm_mesh* m, rxm_attr* flag, void* stream)
{
    int rc = check_dev(m, "rxm_boundary_vertices");
    if (rc) return rc;
    if (!flag || !flag->d || flag->elem != RXM_V || flag->elem_bytes != 4 || flag->nattr != 1)
        return fail(RXM_ERR_INVALID, "rxm_boundary_vertices: flag must be a device 1 x u32 vertex attribute");
    if (flag->layout == RXM_SOA) {  // the kernel addresses flags by slot: run on a slot-ordered stand-in, copy by linear id
        if (!m->scratch[6] && (rc = rxm_attr_create(m, RXM_V, 4, 1, RXM_DEVICE, RXM_AOS, &m->scratch[6]))) return rc;
        if ((rc = rxm_boundary_vertices(m, m->scratch[6], stream))) return rc;
        cudaError_t e = launch_relayout(m->scratch[6]->d, flag->d, slot_map(m, RXM_V), 1, RXM_AOS, RXM_SOA, (cudaStream_t)stream);
        return e == cudaSuccess ? RXM_OK : fail(RXM_ERR_CUDA, std::string("relayout: ") + cudaGetErrorString(e));
    }
    const uint32_t zero = 0;
    rc                  = rxm_attr_reset(flag, &zero, RXM_DEVICE, stream);
    if (rc) return rc;
    const char* why = nullptr;
    cudaError_t e   = launch_boundary_vertices(m->view, m->lim, (uint32_t*)flag->d, (cudaStream_t)stream, &why);
    return kernel_status(e, why, "rxm_boundary_vertices");
}

// One host-buffer call as a K-stage pipeline (plan: build_pipe_plan): the input travels H2D in K global-order pieces on
// its own stream; as soon as a chunk of patches has its owned slots filled and its ribbon owners present, `launch`
// runs on that patch range (a MeshView whose descriptor pointer is offset) on the caller's stream; its results are
// scattered back to global order and leave D2H on a third stream while later pieces are still arriving.
// the pipeline covers the whole mesh: a shard that computes on a sub-range of its patches (ghost patches excluded,
// rxm_mesh_set_active_patches) keeps the plain upload -> kernel -> download path
static bool use_pipeline(const rxm_mesh* m)
{
    return m->plan.K && (!m->active_count || m->active_count == m->h.num_patches);
}

extern "C++" {
template <class Launch>
static int pipelined_host_call(rxm_mesh* m, rxm_attr* ain, rxm_attr* aout, const void* host_in, void* host_out, int k,
                               void* stream, bool sync, Launch launch)
{
    const auto&    P  = m->plan;
    const uint32_t K  = P.K;
    const int      ti = ain->elem, to = aout->elem;
    const size_t   row_i = (size_t)ain->nattr * ain->elem_bytes, row_o = (size_t)aout->nattr * aout->elem_bytes;
    int            rc;
    if ((rc = ensure_stage(m, (size_t)m->h.num_elems[ti] * row_i, k))) return rc;
    if ((rc = ensure_stage(m, (size_t)m->h.num_elems[to] * row_o, k + 1))) return rc;
    auto& pp = m->pipe[k];
    if (!pp.h2d) {
        CU(cudaStreamCreateWithFlags(&pp.h2d, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&pp.d2h, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&pp.start, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&pp.done, cudaEventDisableTiming));
        pp.up.resize(K), pp.comp.resize(K);
        for (uint32_t c = 0; c < K; ++c) {
            CU(cudaEventCreateWithFlags(&pp.up[c], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&pp.comp[c], cudaEventDisableTiming));
        }
    }
    cudaStream_t S = (cudaStream_t)stream;
    uint8_t *    din = (uint8_t*)m->d_stage[k], *dout = (uint8_t*)m->d_stage[k + 1];
    CU(cudaEventRecord(pp.start, S));  // the pipeline starts after the work already queued on the caller's stream
    CU(cudaStreamWaitEvent(pp.h2d, pp.start, 0));
    CU(cudaStreamWaitEvent(pp.d2h, pp.start, 0));
    for (uint32_t c = 0; c < K; ++c) {
        const uint64_t g0 = c ? P.up_hi[ti][c - 1] : 0, g1 = P.up_hi[ti][c];
        if (g1 > g0)
            CU(cudaMemcpyAsync(din + g0 * row_i, (const uint8_t*)host_in + g0 * row_i, (g1 - g0) * row_i,
                               cudaMemcpyHostToDevice, pp.h2d));
        CU(cudaEventRecord(pp.up[c], pp.h2d));
    }
    uint32_t q = 0;
    for (uint32_t c = 0; c < K; ++c) {
        CU(cudaStreamWaitEvent(S, pp.up[c], 0));
        cudaError_t e = launch_permute_to_slots(din, ain->d, m->d_s2g[ti], slot_map(m, ti, P.pb[c], P.pb[c + 1] - P.pb[c]),
                                                ain->elem_bytes, ain->nattr, ain->layout, S);
        if (e != cudaSuccess) return fail(RXM_ERR_CUDA, std::string("permute: ") + cudaGetErrorString(e));
        while (q < K && P.need[q] <= c) {
            MeshView v    = m->view;
            v.desc        = m->d_desc + P.pb[q];
            v.num_patches = P.pb[q + 1] - P.pb[q];
            if ((rc = launch(v, S))) return rc;
            e = launch_permute_to_global(aout->d, dout, m->d_s2g[to], slot_map(m, to, P.pb[q], P.pb[q + 1] - P.pb[q]),
                                         aout->elem_bytes, aout->nattr, aout->layout, S);
            if (e != cudaSuccess) return fail(RXM_ERR_CUDA, std::string("permute: ") + cudaGetErrorString(e));
            CU(cudaEventRecord(pp.comp[q], S));
            CU(cudaStreamWaitEvent(pp.d2h, pp.comp[q], 0));
            const uint64_t g0 = P.down_lo[to][q], g1 = P.down_lo[to][q + 1];
            if (g1 > g0)
                CU(cudaMemcpyAsync((uint8_t*)host_out + g0 * row_o, dout + g0 * row_o, (g1 - g0) * row_o,
                                   cudaMemcpyDeviceToHost, pp.d2h));
            ++q;
        }
    }
    CU(cudaEventRecord(pp.done, pp.d2h));
    CU(cudaStreamWaitEvent(S, pp.done, 0));  // a sync on the caller's stream covers the whole call
    if (sync) CU(cudaStreamSynchronize(S));
    return RXM_OK;
}
}  // extern "C++"

int rxm_vertex_normals_host(rxm_mesh* m, const float* coords, float* normals, void* stream)
{
    int rc = check_dev(m, "rxm_vertex_normals_host");
    if (rc) return rc;
    if (!coords || !normals) return fail(RXM_ERR_INVALID, "rxm_vertex_normals_host: null buffer");
    rxm_attr *x, *n;
    if ((rc = get_scratch(m, 1, &x)) || (rc = get_scratch(m, 2, &n))) return rc;
    if (use_pipeline(m))
        return pipelined_host_call(m, x, n, coords, normals, 2, stream, !g_async_host_calls,
                                   [&](const MeshView& v, cudaStream_t s) {
                                       const char* why = nullptr;
                                       cudaError_t e = launch_vertex_normals(v, m->lim, (const float*)x->d, (float*)n->d, 0, s, &why);
                                       return kernel_status(e, why, "rxm_vertex_normals_host");
                                   });
    if ((rc = upload_via(x, coords, stream, 2))) return rc;
    if ((rc = rxm_vertex_normals(m, x, n, 0, stream))) return rc;
    return download_via(n, normals, stream, 3, !g_async_host_calls);
}

int rxm_laplacian_smooth_host(rxm_mesh* m, const float* coords, float* out, double lr, uint32_t iters, void* stream)
{
    int rc = check_dev(m, "rxm_laplacian_smooth_host");
    if (rc) return rc;
    if (!coords || !out) return fail(RXM_ERR_INVALID, "rxm_laplacian_smooth_host: null buffer");
    rxm_attr *x, *y;
    if ((rc = get_scratch(m, 1, &x)) || (rc = get_scratch(m, 2, &y))) return rc;
    if ((rc = rxm_attr_upload_global(x, coords, stream))) return rc;
    if ((rc = rxm_laplacian_smooth(m, x, y, lr, iters, stream))) return rc;
    return rxm_attr_download_global(y, out, stream);
}

int rxm_query_consume_host(rxm_mesh* m, int op, const float* in, float* out, void* stream)
{
    int rc = check_dev(m, "rxm_query_consume_host");
    if (rc) return rc;
    if (!in || !out || op_src(op) < 0) return fail(RXM_ERR_INVALID, "rxm_query_consume_host: bad argument");
    // cached 1 x fp32 attributes per op (calls for different ops may be in flight on different streams)
    rxm_attr*& a = m->scratch1[2 * (op & 15)];
    rxm_attr*& b = m->scratch1[2 * (op & 15) + 1];
    if (!a && (rc = rxm_attr_create(m, op_dst(op), 4, 1, RXM_DEVICE, RXM_AOS, &a))) return rc;
    if (!b && (rc = rxm_attr_create(m, op_src(op), 4, 1, RXM_DEVICE, RXM_AOS, &b))) return rc;
    if (use_pipeline(m))
        return pipelined_host_call(m, a, b, in, out, 4 + 2 * (op & 15) % 28, stream, !g_async_host_calls,
                                   [&](const MeshView& v, cudaStream_t s) {
                                       const char* why = nullptr;
                                       cudaError_t e = launch_query_consume(op, v, m->lim, view_of<float>(a), view_of<float>(b), s, &why);
                                       return kernel_status(e, why, "rxm_query_consume_host");
                                   });
    if ((rc = upload_via(a, in, stream, 4 + 2 * (op & 15) % 28))) return rc;
    if ((rc = rxm_query_consume(m, op, a, b, stream))) return rc;
    return download_via(b, out, stream, 5 + 2 * (op & 15) % 28, !g_async_host_calls);
}

// ------------------------------------------------------------------ multi-GPU support
int rxm_mesh_set_active_patches(rxm_mesh* m, uint32_t first, uint32_t count)
{
    if (!m || (uint64_t)first + count > m->h.num_patches)
        return fail(RXM_ERR_INVALID, "rxm_mesh_set_active_patches: range outside the mesh");
    m->active_first = first;
    m->active_count = count;
    if (m->on_device) {
        m->view.desc        = m->d_desc + first;
        m->view.num_patches = count;
    }
    return RXM_OK;
}

int rxm_mesh_halo_slots(const rxm_mesh* m, int elem, uint32_t first, uint32_t count, uint32_t** out, uint64_t* n)
{
    if (!m || !out || !n || elem < 0 || elem > 2 || (uint64_t)first + count > m->h.num_patches)
        return fail(RXM_ERR_INVALID, "rxm_mesh_halo_slots: bad argument");
    const HostMesh&      h = m->h;
    if (h.topo.empty()) return fail(RXM_ERR_INVALID, "rxm_mesh_halo_slots: host patch store was released (rxm_mesh_compact)");
    std::vector<uint8_t> mark(h.num_slots[elem], 0);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t p = first; p < (int64_t)first + count; ++p) {
        const PatchDesc&  D   = h.desc[p];
        const uint8_t*    B   = h.topo.data() + D.topo_off;
        const uint32_t*   own = reinterpret_cast<const uint32_t*>(B + D.off_own(elem));
        const StashEntry* st  = reinterpret_cast<const StashEntry*>(B + D.off_stash());
        for (uint32_t i = 0; i < (uint32_t)(D.n[elem] - D.n_owned[elem]); ++i) {
            const uint32_t q = st[own[i] >> 16].patch;
            if (q < first || q >= first + count) mark[st[own[i] >> 16].slot_base[elem] + (own[i] & 0xFFFFu)] = 1;
        }
    }
    uint64_t k = 0;
    for (uint8_t b : mark)
        k += b;
    uint32_t* r = (uint32_t*)malloc(std::max<uint64_t>(k, 1) * sizeof(uint32_t));
    if (!r) return fail(RXM_ERR_INVALID, "rxm_mesh_halo_slots: out of memory");
    uint64_t j = 0;
    for (uint32_t s = 0; s < h.num_slots[elem]; ++s)
        if (mark[s]) r[j++] = s;
    *out = r;
    *n   = k;
    return RXM_OK;
}

int rxm_memcpy_d2h(void* host, const void* dev, uint64_t bytes)
{
    CU(cudaMemcpy(host, dev, bytes, cudaMemcpyDeviceToHost));
    return RXM_OK;
}

void rxm_free(void* p)
{
    free(p);
}

static int rows_check(rxm_attr* a, const char* who)
{
    if (!a || !a->d) return fail(RXM_ERR_INVALID, std::string(who) + ": attribute has no device storage");
    if (a->layout == RXM_SOA || (a->layout != RXM_AOS && a->nattr != 1))  // rows are addressed by slot
        return fail(RXM_ERR_INVALID, std::string(who) + ": AoS attributes only");
    if ((a->elem_bytes * a->nattr) % 4) return fail(RXM_ERR_INVALID, std::string(who) + ": row size must be a multiple of 4 bytes");
    return RXM_OK;
}

int rxm_attr_gather_slots(rxm_attr* a, const uint32_t* dev_idx, uint64_t n, void* dev_out, void* stream)
{
    int rc = rows_check(a, "rxm_attr_gather_slots");
    if (rc) return rc;
    cudaError_t e = launch_slot_rows(true, a->d, dev_idx, n, a->elem_bytes * a->nattr / 4, dev_out, (cudaStream_t)stream);
    return e == cudaSuccess ? RXM_OK : fail(RXM_ERR_CUDA, cudaGetErrorString(e));
}

int rxm_attr_scatter_slots(rxm_attr* a, const uint32_t* dev_idx, uint64_t n, const void* dev_in, void* stream)
{
    int rc = rows_check(a, "rxm_attr_scatter_slots");
    if (rc) return rc;
    cudaError_t e = launch_slot_rows(false, a->d, dev_idx, n, a->elem_bytes * a->nattr / 4, const_cast<void*>(dev_in),
                                     (cudaStream_t)stream);
    return e == cudaSuccess ? RXM_OK : fail(RXM_ERR_CUDA, cudaGetErrorString(e));
}

int rxm_attr_push_slots(rxm_attr* a, const uint32_t* dev_local_idx, void* remote_data, const uint32_t* dev_remote_idx,
                        uint64_t n, void* stream)
{
    int rc = rows_check(a, "rxm_attr_push_slots");
    if (rc) return rc;
    if (!remote_data) return fail(RXM_ERR_INVALID, "rxm_attr_push_slots: null remote pointer");
    cudaError_t e = launch_push_rows(a->d, dev_local_idx, remote_data, dev_remote_idx, n, a->elem_bytes * a->nattr / 4,
                                     (cudaStream_t)stream);
    return e == cudaSuccess ? RXM_OK : fail(RXM_ERR_CUDA, cudaGetErrorString(e));
}

// ---- compute + halo exchange in one kernel (FusedHaloView, rxm_kernels.h) ----
struct rxm_fused_halo
{
    rxm_mesh* m        = nullptr;
    uint32_t  npeers   = 0, n_sync_blocks = 0, shift = 0;
    std::vector<uint8_t> reads;  // host copy of d_reads
    uint32_t *d_flags = nullptr, *d_ctr = nullptr, *d_push_off = nullptr;
    uint2*    d_push  = nullptr;
    uint8_t*  d_reads = nullptr;
    float**   d_peer_attr[2] = {nullptr, nullptr};  // neighbour-side base of attribute A / attribute B
    uint32_t** d_peer_flag   = nullptr;
};

int rxm_fused_halo_create(rxm_mesh* m, uint32_t npeers, rxm_fused_halo** out)
{
    int rc = check_dev(m, "rxm_fused_halo_create");
    if (rc) return rc;
    if (!out || npeers == 0 || npeers > 64) return fail(RXM_ERR_INVALID, "rxm_fused_halo_create: bad argument");
    if (m->h.topo.empty()) return fail(RXM_ERR_INVALID, "rxm_fused_halo_create: host patch store was released (rxm_mesh_compact)");
    rxm_fused_halo* h = new rxm_fused_halo();
    h->m = m, h->npeers = npeers;
    CU(cudaMalloc(&h->d_flags, 4 * npeers));
    CU(cudaMemset(h->d_flags, 0, 4 * npeers));
    CU(cudaMalloc(&h->d_ctr, 4));
    CU(cudaMemset(h->d_ctr, 0, 4));
    // patches that read ghost slots: a stash entry names a patch outside the active (real) range
    const HostMesh&      H = m->h;
    const uint32_t       a0 = m->active_count ? m->active_first : 0, a1 = m->active_count ? a0 + m->active_count : H.num_patches;
    std::vector<uint8_t> reads(H.num_patches, 0);
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < (int64_t)H.num_patches; ++p) {
        const PatchDesc&  D  = H.desc[p];
        const StashEntry* st = reinterpret_cast<const StashEntry*>(H.topo.data() + D.topo_off + D.off_stash());
        for (uint32_t i = 0; i < D.n_stash; ++i)
            if (st[i].patch < a0 || st[i].patch >= a1) reads[p] = 1;
    }
    CU(cudaMalloc(&h->d_reads, std::max<size_t>(reads.size(), 1)));
    CU(cudaMemcpy(h->d_reads, reads.data(), reads.size(), cudaMemcpyHostToDevice));
    h->reads.swap(reads);
    *out = h;
    return RXM_OK;
}

void* rxm_fused_halo_flags(rxm_fused_halo* h)
{
    return h ? h->d_flags : nullptr;
}

uint32_t rxm_fused_halo_sync_blocks(const rxm_fused_halo* h)
{
    return h ? h->n_sync_blocks : 0u;
}

// push lists: for patch index p the entries [push_off[p], push_off[p+1]) of (local vertex id | neighbour << 16) and the
// slot on that neighbour; peer_attr_a / _b: neighbour-side base pointers (rxm_ipc_open) of the two attributes the
// iteration ping-pongs between; peer_flag: neighbour-side address of the flag word this rank raises
int rxm_fused_halo_set(rxm_fused_halo* h, const uint32_t* push_off, const uint32_t* push_lid_peer, const uint32_t* push_slot,
                       uint64_t n_push, void* const* peer_attr_a, void* const* peer_attr_b, void* const* peer_flag)
{
    if (!h || !push_off || !peer_attr_a || !peer_attr_b || !peer_flag) return fail(RXM_ERR_INVALID, "rxm_fused_halo_set: null argument");
    const uint32_t P = h->m->h.num_patches;
    if (push_off[P] != n_push) return fail(RXM_ERR_INVALID, "rxm_fused_halo_set: push_off does not end at n_push");
    uint32_t n_push_blocks = 0;
    h->n_sync_blocks       = 0;
    {   // blocks that check in at the step counter: the ACTIVE patches that read ghost slots or push rows
        const uint32_t a0 = h->m->active_count ? h->m->active_first : 0, cnt = h->m->active_count ? h->m->active_count : P;
        for (uint32_t p = 0; p < P; ++p) {
            const bool pushes = push_off[p + 1] > push_off[p];
            if (pushes && (p < a0 || p >= a0 + cnt))
                return fail(RXM_ERR_INVALID, "rxm_fused_halo_set: a patch outside the active range has rows to push");
            n_push_blocks += pushes ? 1u : 0u;
            if (p >= a0 && p < a0 + cnt && (pushes || h->reads[p])) ++h->n_sync_blocks;
        }
    }
    {   // rotation: start the launch at the first pushing patch of the upper half of the active range
        const uint32_t a0 = h->m->active_count ? h->m->active_first : 0, cnt = h->m->active_count ? h->m->active_count : P;
        h->shift = 0;
        for (uint32_t r = cnt / 2; r < cnt; ++r)
            if (push_off[a0 + r + 1] > push_off[a0 + r]) {
                h->shift = r;
                break;
            }
    }
    if (n_push_blocks == 0) return fail(RXM_ERR_INVALID, "rxm_fused_halo_set: nothing to push (no neighbour shares an element)");
    std::vector<uint2> e(std::max<uint64_t>(n_push, 1));
    for (uint64_t i = 0; i < n_push; ++i)
        e[i] = make_uint2(push_lid_peer[i], push_slot[i]);
    CU(cudaMalloc(&h->d_push_off, 4 * (size_t)(P + 1)));
    CU(cudaMemcpy(h->d_push_off, push_off, 4 * (size_t)(P + 1), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&h->d_push, sizeof(uint2) * e.size()));
    CU(cudaMemcpy(h->d_push, e.data(), sizeof(uint2) * e.size(), cudaMemcpyHostToDevice));
    const void* const* src[3] = {peer_attr_a, peer_attr_b, peer_flag};
    void**             dst[3] = {(void**)&h->d_peer_attr[0], (void**)&h->d_peer_attr[1], (void**)&h->d_peer_flag};
    for (int k = 0; k < 3; ++k) {
        CU(cudaMalloc(dst[k], sizeof(void*) * h->npeers));
        CU(cudaMemcpy(*dst[k], src[k], sizeof(void*) * h->npeers, cudaMemcpyHostToDevice));
    }
    return RXM_OK;
}

void rxm_fused_halo_destroy(rxm_fused_halo* h)
{
    if (!h) return;
    cudaFree(h->d_flags), cudaFree(h->d_ctr), cudaFree(h->d_push_off), cudaFree(h->d_push), cudaFree(h->d_reads);
    cudaFree(h->d_peer_attr[0]), cudaFree(h->d_peer_attr[1]), cudaFree(h->d_peer_flag);
    delete h;
}

// one Laplacian step in -> out with the halo push fused in; out_is_b selects which neighbour-side attribute `out` is;
// step = number of fused steps every rank has issued before this one (monotonic)
int rxm_laplacian_smooth_fused(rxm_mesh* m, rxm_attr* in, rxm_attr* out, double lr, rxm_fused_halo* h, int out_is_b,
                               uint32_t step, void* stream)
{
    int rc = check_dev(m, "rxm_laplacian_smooth_fused");
    if (rc) return rc;
    if (!h || !h->d_push_off || !is_vec3_aos(in) || !is_vec3_aos(out) || in == out)
        return fail(RXM_ERR_INVALID, "rxm_laplacian_smooth_fused: bad argument");
    FusedHaloView v;
    v.push_off = h->d_push_off, v.push = h->d_push, v.peer_out = h->d_peer_attr[out_is_b ? 1 : 0];
    v.peer_flag = h->d_peer_flag, v.reads_ghost = h->d_reads, v.flags = h->d_flags, v.done_ctr = h->d_ctr;
    v.npeers = h->npeers, v.first = m->active_count ? m->active_first : 0, v.step = step;
    v.n_sync_blocks = h->n_sync_blocks, v.shift = h->shift;
    const char* why = nullptr;
    cudaError_t e   = launch_laplacian_step_fused(m->view, m->lim, (const float*)in->d, (float*)out->d, lr, v,
                                                  (cudaStream_t)stream, &why);
    return kernel_status(e, why, "rxm_laplacian_smooth_fused");
}

int rxm_ipc_export(void* dev_ptr, void* handle64)
{
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, dev_ptr));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, 64);
    return RXM_OK;
}

int rxm_ipc_open(const void* handle64, void** dev_ptr)
{
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CU(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return RXM_OK;
}

int rxm_ipc_close(void* dev_ptr)
{
    CU(cudaIpcCloseMemHandle(dev_ptr));
    return RXM_OK;
}

// ------------------------------------------------------------------ reductions (ReduceHandle)
int rxm_attr_reduce(rxm_attr* a, rxm_attr* b, int kind, uint32_t attribute_id, double* out_value, uint64_t* out_handle,
                    void* stream)
{
    if (!a || !a->d || a->elem_bytes != 4) return fail(RXM_ERR_INVALID, "rxm_attr_reduce: needs a device fp32 attribute");
    if (kind < 0 || kind > 6) return fail(RXM_ERR_INVALID, "rxm_attr_reduce: unknown kind");
    if (kind == 0 && (!b || !b->d || b->elem != a->elem || b->nattr != a->nattr || b->layout != a->layout || b->elem_bytes != 4))
        return fail(RXM_ERR_INVALID, "rxm_attr_reduce: dot needs two attributes of the same shape");
    if (attribute_id != INVALID32_ && attribute_id >= a->nattr) return fail(RXM_ERR_INVALID, "rxm_attr_reduce: attribute_id out of range");
    rxm_mesh* m = a->m;
    int       rc = check_dev(m, "rxm_attr_reduce");
    if (rc) return rc;
    const uint32_t grid = std::min<uint32_t>(m->view.num_patches, 148u * 8u);
    if ((rc = ensure_stage(m, ((size_t)grid + 1) * 16, 1))) return rc;
    MeshView full = m->view;  // reductions cover every owned element of the (local) mesh
    cudaError_t e = launch_reduce(full, view_of<float>(a), kind == 0 ? view_of<float>(b) : view_of<float>(a), a->elem, kind,
                                  attribute_id, m->d_stage[1], grid, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(RXM_ERR_CUDA, std::string("reduce: ") + cudaGetErrorString(e));
    struct { double v; uint64_t h; } r;
    CU(cudaMemcpyAsync(&r, (char*)m->d_stage[1] + (size_t)grid * 16, 16, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    if (out_value) *out_value = r.v;
    if (out_handle) *out_handle = r.h;
    return RXM_OK;
}

// ------------------------------------------------------------------ saved patchings
// The reference's Patcher::serialize layout (patcher/patcher.h:162-182) in cereal's PortableBinary archive:
// 1 byte (1 = little endian), 9 x u32 scalars, 7 x (u64 length + u32[length]), 1 x f32.
int rxm_patcher_file_read(const char* path, rxm_patcher_file* out)
{
    if (!path || !out) return fail(RXM_ERR_INVALID, "rxm_patcher_file_read: null argument");
    memset(out, 0, sizeof(*out));
    FILE* f = fopen(path, "rb");
    if (!f) return fail(RXM_ERR_INVALID, std::string("rxm_patcher_file_read: cannot open ") + path);
    uint8_t endian = 0;
    bool    ok     = fread(&endian, 1, 1, f) == 1 && endian == 1;
    ok             = ok && fread(out->header, 4, 9, f) == 9;
    for (int i = 0; ok && i < 7; ++i) {
        uint64_t n = 0;
        ok         = fread(&n, 8, 1, f) == 1 && n < (1ull << 32);
        if (!ok) break;
        out->len[i] = n;
        out->vec[i] = (uint32_t*)malloc(std::max<uint64_t>(n, 1) * 4);
        ok          = out->vec[i] && fread(out->vec[i], 4, n, f) == n;
    }
    ok = ok && fread(&out->patching_time_ms, 4, 1, f) == 1;
    fclose(f);
    if (!ok) {
        rxm_patcher_file_free(out);
        return fail(RXM_ERR_INVALID, std::string("rxm_patcher_file_read: not a little-endian Patcher archive: ") + path);
    }
    return RXM_OK;
}

void rxm_patcher_file_free(rxm_patcher_file* pf)
{
    if (!pf) return;
    for (int i = 0; i < 7; ++i) {
        free(pf->vec[i]);
        pf->vec[i] = nullptr;
    }
}

int rxm_mesh_save_patcher_file(const rxm_mesh* m, const char* path)
{
    if (!m || !path) return fail(RXM_ERR_INVALID, "rxm_mesh_save_patcher_file: null argument");
    const HostMesh& h = m->h;
    if (h.ltog[ELEM_F].empty()) return fail(RXM_ERR_INVALID, "rxm_mesh_save_patcher_file: local->global lists were released");
    const uint32_t        P = h.num_patches;
    std::vector<uint32_t> pval, poff(P), rval, roff(P);
    for (uint32_t p = 0; p < P; ++p) {
        const uint32_t* l  = h.ltog[ELEM_F].data() + h.ltog_off[ELEM_F][p];
        const uint32_t  no = h.desc[p].n_owned[ELEM_F], n = h.desc[p].n[ELEM_F];
        pval.insert(pval.end(), l, l + no);
        rval.insert(rval.end(), l + no, l + n);
        poff[p] = (uint32_t)pval.size();  // the reference stores inclusive END offsets (rxmesh.cpp:760-768)
        roff[p] = (uint32_t)rval.size();
    }
    FILE* f = fopen(path, "wb");
    if (!f) return fail(RXM_ERR_INVALID, std::string("rxm_mesh_save_patcher_file: cannot open ") + path);
    const uint8_t  endian = 1;
    const uint32_t hdr[9] = {h.patch_size, P, h.num_elems[ELEM_V], h.num_elems[ELEM_E], h.num_elems[ELEM_F], P, P, 1, 1};
    fwrite(&endian, 1, 1, f);
    fwrite(hdr, 4, 9, f);
    const std::vector<uint32_t>* vecs[7] = {&h.elem_patch[ELEM_F], &h.elem_patch[ELEM_V], &h.elem_patch[ELEM_E], &pval, &poff, &rval, &roff};
    for (auto* v : vecs) {
        const uint64_t n = v->size();
        fwrite(&n, 8, 1, f);
        fwrite(v->data(), 4, n, f);
    }
    const float t = (float)(h.patcher_seconds * 1e3);
    fwrite(&t, 4, 1, f);
    fclose(f);
    return RXM_OK;
}

void rxm_set_async(int on)
{
    g_async_host_calls = on;
}

int rxm_stream_sync(void* stream)
{
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    return RXM_OK;
}

uint64_t rxm_launch_count(void)
{
    return launch_counter();
}

}  // extern "C"
