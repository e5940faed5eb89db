// rxm_multi.cu -- single-process multi-GPU mode behind the C ABI (include/rxmesh_b200.h: rxm_multi_*).
//
// New relative to the reference, which runs on one device (the survey asks for "device-list / shard options"
// on RXMeshStatic).  One host process drives every GPU of the box: the mesh is cut into contiguous patch-id ranges (one per
// device, balanced by faces), every device gets the faces of its range plus two vertex-rings of foreign faces ("ghost"
// patches: built like any patch, never computed on), and the rows a shard's real patches read from ghost slots are kept
// current by the owners: the Laplacian step that produces a mirrored row also stores it into the neighbour's ghost slot
// over NVLink and raises a flag there (k_laplacian_fan2<true>, the same kernel and protocol the one-process-per-GPU mode of
// rxmesh_b200/distributed.py uses through cudaIpc; here the peers' memory is addressed directly after
// cudaDeviceEnablePeerAccess).  Everything is planned on the host in this file -- shard construction, matching ghost slots with
// their owners by GLOBAL vertex id, push lists -- so the C++ drop-in (include/rxmesh/multi_gpu.h) needs no Python.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rxmesh_b200.h"
#include "mesh_builder.h"

using namespace rxm;

extern "C" int rxm_set_last_error(int code, const char* msg);  // rxm_capi.cu

namespace {
struct Shard
{
    int                   device = -1;
    rxm_mesh*             mesh   = nullptr;
    rxm_attr *            a = nullptr, *b = nullptr;
    rxm_fused_halo*       halo = nullptr;
    cudaStream_t          stream = nullptr;
    uint32_t              first = 0, count = 0;        // real patches (indices into the shard's patch array)
    std::vector<uint32_t> l2g_v, l2g_f;                // sorted global ids of the shard's vertices / faces
    std::vector<uint32_t> patch_global;                // shard patch q <-> global patch id
    std::vector<uint8_t>  real_v;                      // per local vertex: owned by a real patch of this shard
    std::vector<uint32_t> peers;                       // neighbour shards (send or receive), ascending
    std::vector<std::vector<uint32_t>> recv, send;     // per shard index: my ghost slots to fill / my slots to push (same order)
    std::vector<float>    stage;                       // host staging (local vertex order)
};
}  // namespace

struct rxm_multi
{
    std::vector<Shard>    sh;
    std::vector<uint32_t> bounds;  // [n + 1] global patch-id ranges
    uint32_t              num_vertices = 0, num_faces = 0, num_patches = 0;
    uint32_t              step = 0;  // fused steps issued so far (flag values, ping-pong parity)
    uint64_t              halo_elements = 0;
    bool                  on_device = false;
};

#define MFAIL(code, msg) return rxm_set_last_error(code, (std::string("rxm_multi: ") + (msg)).c_str())
#define MCK(call)                                      \
    do {                                               \
        int rc_ = (call);                              \
        if (rc_) return rc_; /* message already set */ \
    } while (0)
#define MCU(call)                                                                      \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) MFAIL(RXM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

extern "C" {

void rxm_multi_destroy(rxm_multi* M)
{
    if (!M) return;
    for (auto& s : M->sh) {
        if (s.device >= 0 && M->on_device) cudaSetDevice(s.device);
        if (s.halo) rxm_fused_halo_destroy(s.halo);
        if (s.a) rxm_attr_destroy(s.a);
        if (s.b) rxm_attr_destroy(s.b);
        if (s.stream) cudaStreamDestroy(s.stream);
        if (s.mesh) rxm_mesh_destroy(s.mesh);
    }
    delete M;
}

int rxm_multi_create(const uint32_t* fv, uint32_t num_faces, const uint32_t* face_patch, uint32_t patch_size, const int* devices,
                     int num_shards, int num_threads, rxm_multi** out)
{
    if (!fv || !out || num_shards < 1 || num_faces == 0) MFAIL(RXM_ERR_INVALID, "bad argument");
    patch_size = patch_size ? patch_size : 512;
    // ---- 1. face -> patch of the WHOLE mesh ----
    std::vector<uint32_t> fp;
    uint32_t              P = 0;
    if (face_patch) {
        fp.assign(face_patch, face_patch + num_faces);
        for (uint32_t f = 0; f < num_faces; ++f)
            P = std::max(P, fp[f] + 1);
    } else {
        BuildOptions opt;
        opt.patch_size = patch_size, opt.num_threads = num_threads;
        opt.reorder_patches = getenv("RXM_NO_PATCH_REORDER") == nullptr;  // as rxm_mesh_create_ex
        const std::string e = compute_face_patch(fv, num_faces, opt, fp, P);
        if (!e.empty()) MFAIL(RXM_ERR_INVALID, e);
    }
    uint32_t nv = 0;
    for (uint64_t i = 0; i < 3ull * num_faces; ++i)
        nv = std::max(nv, fv[i]);
    nv += 1;
    rxm_multi* M    = new rxm_multi();
    M->num_vertices = nv, M->num_faces = num_faces, M->num_patches = P;
    auto fail_free = [&](int rc) {
        rxm_multi_destroy(M);
        return rc;
    };
    // ---- 2. contiguous patch-id ranges, balanced by faces (rxmesh_b200/distributed.py: patch_ranges) ----
    {
        std::vector<uint64_t> cum((size_t)P + 1, 0);
        for (uint32_t f = 0; f < num_faces; ++f)
            cum[fp[f] + 1]++;
        for (uint32_t p = 0; p < P; ++p)
            cum[p + 1] += cum[p];
        M->bounds.assign((size_t)num_shards + 1, 0);
        for (int r = 1; r < num_shards; ++r) {
            const double target = (double)cum[P] * r / num_shards;
            M->bounds[r] = (uint32_t)(std::lower_bound(cum.begin(), cum.end(), target, [](uint64_t c, double t) { return (double)c < t; }) - cum.begin());
        }
        M->bounds[num_shards] = P;
        for (int r = 0; r < num_shards; ++r)
            if (M->bounds[r + 1] <= M->bounds[r]) {
                rxm_set_last_error(RXM_ERR_INVALID, "rxm_multi: fewer patches than shards (use a smaller patch_size or fewer devices)");
                return fail_free(RXM_ERR_INVALID);
            }
    }
    // ---- 3. shards: real faces + two vertex-rings, built and (if devices are given) uploaded ----
    M->sh.resize(num_shards);
    std::vector<uint32_t> g2l(nv);
    for (int r = 0; r < num_shards; ++r) {
        Shard&               S = M->sh[r];
        std::vector<uint8_t> selv(nv, 0), self(num_faces, 0);
        const uint32_t       b0 = M->bounds[r], b1 = M->bounds[r + 1];
#pragma omp parallel for schedule(static)
        for (int64_t f = 0; f < (int64_t)num_faces; ++f)
            self[f] = fp[f] >= b0 && fp[f] < b1;
        for (int ring = 0; ring < 2; ++ring) {
            std::fill(selv.begin(), selv.end(), 0);
            for (uint32_t f = 0; f < num_faces; ++f)
                if (self[f]) selv[fv[3ull * f]] = selv[fv[3ull * f + 1]] = selv[fv[3ull * f + 2]] = 1;
#pragma omp parallel for schedule(static)
            for (int64_t f = 0; f < (int64_t)num_faces; ++f)
                self[f] = selv[fv[3ull * f]] | selv[fv[3ull * f + 1]] | selv[fv[3ull * f + 2]];
        }
        std::fill(selv.begin(), selv.end(), 0);
        for (uint32_t f = 0; f < num_faces; ++f)
            if (self[f]) {
                S.l2g_f.push_back(f);
                selv[fv[3ull * f]] = selv[fv[3ull * f + 1]] = selv[fv[3ull * f + 2]] = 1;
            }
        for (uint32_t v = 0; v < nv; ++v)
            if (selv[v]) {
                g2l[v] = (uint32_t)S.l2g_v.size();
                S.l2g_v.push_back(v);
            }
        std::vector<uint32_t> lfv(3 * S.l2g_f.size()), lfp(S.l2g_f.size());
        for (size_t i = 0; i < S.l2g_f.size(); ++i) {
            const uint32_t f = S.l2g_f[i];
            lfp[i]           = fp[f];
            for (int j = 0; j < 3; ++j)
                lfv[3 * i + j] = g2l[fv[3ull * f + j]];
        }
        S.patch_global = lfp;
        std::sort(S.patch_global.begin(), S.patch_global.end());
        S.patch_global.erase(std::unique(S.patch_global.begin(), S.patch_global.end()), S.patch_global.end());
        if (devices) {
            S.device = devices[r];
            if (cudaSetDevice(S.device) != cudaSuccess) {
                rxm_set_last_error(RXM_ERR_CUDA, "rxm_multi: cudaSetDevice failed");
                return fail_free(RXM_ERR_CUDA);
            }
        }
        int rc = rxm_mesh_create_ex(lfv.data(), (uint32_t)lfp.size(), lfp.data(), patch_size, num_threads, RXM_BUILD_NO_RING2, &S.mesh);
        if (rc) return fail_free(rc);
        S.first = (uint32_t)(std::lower_bound(S.patch_global.begin(), S.patch_global.end(), b0) - S.patch_global.begin());
        S.count = (uint32_t)(std::lower_bound(S.patch_global.begin(), S.patch_global.end(), b1) - S.patch_global.begin()) - S.first;
        if ((rc = rxm_mesh_set_active_patches(S.mesh, S.first, S.count))) return fail_free(rc);
        const uint32_t* ep = rxm_mesh_elem_patch(S.mesh, RXM_V);
        S.real_v.resize(S.l2g_v.size());
        for (size_t v = 0; v < S.l2g_v.size(); ++v)
            S.real_v[v] = ep[v] >= S.first && ep[v] < S.first + S.count;
    }
    // ---- 4. halo plan: every ghost slot a shard's real patches read, matched with its owner by global vertex id ----
    for (int r = 0; r < num_shards; ++r)
        M->sh[r].recv.assign(num_shards, {}), M->sh[r].send.assign(num_shards, {});
    for (int r = 0; r < num_shards && num_shards > 1; ++r) {
        Shard&    S = M->sh[r];
        uint32_t* slots = nullptr;
        uint64_t  n     = 0;
        int       rc    = rxm_mesh_halo_slots(S.mesh, RXM_V, S.first, S.count, &slots, &n);
        if (rc) return fail_free(rc);
        const uint32_t *s2g = rxm_mesh_slot_to_global(S.mesh, RXM_V), *ep = rxm_mesh_elem_patch(S.mesh, RXM_V);
        for (uint64_t i = 0; i < n; ++i) {
            const uint32_t loc = s2g[slots[i]], g = S.l2g_v[loc], gp = S.patch_global[ep[loc]];
            const int      q   = (int)(std::upper_bound(M->bounds.begin(), M->bounds.end(), gp) - M->bounds.begin()) - 1;
            Shard&         Q   = M->sh[q];
            auto           it  = std::lower_bound(Q.l2g_v.begin(), Q.l2g_v.end(), g);
            if (q == r || it == Q.l2g_v.end() || *it != g || !Q.real_v[it - Q.l2g_v.begin()]) {
                rxm_free(slots);
                rxm_set_last_error(RXM_ERR_INVALID, "rxm_multi: internal error, a ghost vertex has no owner shard");
                return fail_free(RXM_ERR_INVALID);
            }
            S.recv[q].push_back(slots[i]);
            Q.send[r].push_back(rxm_mesh_global_to_slot(Q.mesh, RXM_V)[it - Q.l2g_v.begin()]);
        }
        M->halo_elements += n;
        rxm_free(slots);
    }
    for (int r = 0; r < num_shards; ++r)
        for (int q = 0; q < num_shards; ++q)
            if (!M->sh[r].recv[q].empty() || !M->sh[r].send[q].empty()) M->sh[r].peers.push_back((uint32_t)q);
    if (!devices) {  // host-only plan (tests without a GPU)
        *out = M;
        return RXM_OK;
    }
    // ---- 5. devices: upload, peer access, attributes, fused-halo state ----
    for (int r = 0; r < num_shards; ++r) {
        Shard& S = M->sh[r];
        if (cudaSetDevice(S.device) != cudaSuccess) {
            rxm_set_last_error(RXM_ERR_CUDA, "rxm_multi: cudaSetDevice failed");
            return fail_free(RXM_ERR_CUDA);
        }
        int rc = rxm_mesh_to_device(S.mesh);
        if (rc) return fail_free(rc);
        for (int q = 0; q < num_shards; ++q)
            if (q != r && M->sh[q].device != S.device) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, S.device, M->sh[q].device);
                if (!can) {
                    rxm_set_last_error(RXM_ERR_UNSUPPORTED, "rxm_multi: the devices have no peer access to each other");
                    return fail_free(RXM_ERR_UNSUPPORTED);
                }
                cudaError_t e = cudaDeviceEnablePeerAccess(M->sh[q].device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    rxm_set_last_error(RXM_ERR_CUDA, cudaGetErrorString(e));
                    return fail_free(RXM_ERR_CUDA);
                }
                cudaGetLastError();
            }
        if (cudaStreamCreateWithFlags(&S.stream, cudaStreamNonBlocking) != cudaSuccess) return fail_free(RXM_ERR_CUDA);
        if ((rc = rxm_attr_create(S.mesh, RXM_V, 4, 3, RXM_DEVICE, RXM_AOS, &S.a))) return fail_free(rc);
        if ((rc = rxm_attr_create(S.mesh, RXM_V, 4, 3, RXM_DEVICE, RXM_AOS, &S.b))) return fail_free(rc);
        if (!S.peers.empty() && (rc = rxm_fused_halo_create(S.mesh, (uint32_t)S.peers.size(), &S.halo))) return fail_free(rc);
    }
    M->on_device = true;
    for (int r = 0; r < num_shards; ++r) {
        Shard& S = M->sh[r];
        if (S.peers.empty()) continue;
        if (cudaSetDevice(S.device) != cudaSuccess) {
            rxm_set_last_error(RXM_ERR_CUDA, "rxm_multi: cudaSetDevice failed");
            return fail_free(RXM_ERR_CUDA);
        }
        const uint32_t  np = (uint32_t)rxm_mesh_info(S.mesh, RXM_INFO_NUM_PATCHES);
        const uint32_t* sb = rxm_mesh_slot_base(S.mesh, RXM_V);
        struct E
        {
            uint32_t patch, lid_peer, slot;
        };
        std::vector<E> ent;
        for (size_t k = 0; k < S.peers.size(); ++k) {
            const uint32_t q = S.peers[k];
            for (size_t i = 0; i < S.send[q].size(); ++i) {
                const uint32_t s = S.send[q][i];
                const uint32_t p = (uint32_t)(std::upper_bound(sb, sb + np + 1, s) - sb) - 1;
                ent.push_back({p, (s - sb[p]) | ((uint32_t)k << 16), M->sh[q].recv[r][i]});
            }
        }
        std::stable_sort(ent.begin(), ent.end(), [](const E& x, const E& y) { return x.patch < y.patch; });
        std::vector<uint32_t> off((size_t)np + 1, 0), lp(ent.size()), slot(ent.size());
        for (size_t i = 0; i < ent.size(); ++i)
            off[ent[i].patch + 1]++, lp[i] = ent[i].lid_peer, slot[i] = ent[i].slot;
        for (uint32_t p = 0; p < np; ++p)
            off[p + 1] += off[p];
        std::vector<void*> pa(S.peers.size()), pb(S.peers.size()), pf(S.peers.size());
        for (size_t k = 0; k < S.peers.size(); ++k) {
            Shard& Q = M->sh[S.peers[k]];
            pa[k]    = rxm_attr_data(Q.a, RXM_DEVICE);
            pb[k]    = rxm_attr_data(Q.b, RXM_DEVICE);
            const size_t me = std::lower_bound(Q.peers.begin(), Q.peers.end(), (uint32_t)r) - Q.peers.begin();
            pf[k]           = (uint32_t*)rxm_fused_halo_flags(Q.halo) + me;  // my flag word on that neighbour
        }
        int rc = rxm_fused_halo_set(S.halo, off.data(), lp.data(), slot.data(), ent.size(), pa.data(), pb.data(), pf.data());
        if (rc) return fail_free(rc);
        // Shards that SHARE a device (tests, debugging): a fused step's boundary blocks spin until the neighbour's previous
        // step has raised its flags.  If that neighbour runs on the same device and the spinning blocks could fill it, the
        // neighbour's kernel might never be scheduled.  Refuse such plans instead of risking a hang.
        bool shared = false;
        for (int q = 0; q < num_shards; ++q)
            shared |= q != r && M->sh[q].device == S.device;
        if (shared && rxm_fused_halo_sync_blocks(S.halo) > 592u) {
            rxm_set_last_error(RXM_ERR_UNSUPPORTED,
                               "rxm_multi: shards that share a device have too many boundary patches (more than 592 waiting blocks "
                               "per step could keep the neighbour's kernel off the device): give every shard its own device");
            return fail_free(RXM_ERR_UNSUPPORTED);
        }
    }
    *out = M;
    return RXM_OK;
}

uint64_t rxm_multi_info(const rxm_multi* M, int what, int shard)
{
    if (!M) return 0;
    switch (what) {
        case 0: return M->sh.size();
        case 1: return M->num_patches;
        case 2: return M->halo_elements;                  // ghost vertex rows refreshed per exchange, all shards
        case 3: return M->num_vertices;
        case 4: return M->num_faces;
        default: break;
    }
    if (shard < 0 || shard >= (int)M->sh.size()) return 0;
    const Shard& S = M->sh[shard];
    switch (what) {
        case 10: return S.count;                          // real patches
        case 11: return S.l2g_f.size();                   // faces held (real + ghost rings)
        case 12: { uint64_t n = 0; for (uint8_t b : S.real_v) n += b; return n; }  // vertices this shard is responsible for
        case 13: { uint64_t n = 0; for (auto& v : S.recv) n += v.size(); return n; }
        case 14: { uint64_t n = 0; for (auto& v : S.send) n += v.size(); return n; }
        case 15: return S.peers.size();
        default: return 0;
    }
}

rxm_mesh* rxm_multi_shard_mesh(rxm_multi* M, int shard)
{
    return (M && shard >= 0 && shard < (int)M->sh.size()) ? M->sh[shard].mesh : nullptr;
}

// scatter a global [V][3] array into every shard's attribute (ghost patches included: their rows ARE the halo)
static int multi_upload(rxm_multi* M, const float* coords, bool into_b)
{
    for (auto& S : M->sh) {
        MCU(cudaSetDevice(S.device));
        S.stage.resize(3 * S.l2g_v.size());
        for (size_t v = 0; v < S.l2g_v.size(); ++v)
            memcpy(&S.stage[3 * v], coords + 3ull * S.l2g_v[v], 12);
        MCK(rxm_attr_upload_global(into_b ? S.b : S.a, S.stage.data(), S.stream));
    }
    for (auto& S : M->sh) {
        MCU(cudaSetDevice(S.device));
        MCU(cudaStreamSynchronize(S.stream));
    }
    return RXM_OK;
}

static int multi_download(rxm_multi* M, bool from_b, float* out)
{
    for (auto& S : M->sh) {
        MCU(cudaSetDevice(S.device));
        S.stage.resize(3 * S.l2g_v.size());
        MCK(rxm_attr_download_global(from_b ? S.b : S.a, S.stage.data(), S.stream));  // synchronises the stream
        for (size_t v = 0; v < S.l2g_v.size(); ++v)
            if (S.real_v[v]) memcpy(out + 3ull * S.l2g_v[v], &S.stage[3 * v], 12);
    }
    return RXM_OK;
}

int rxm_multi_laplacian_smooth(rxm_multi* M, const float* coords, float* out, double lr, uint32_t iters)
{
    if (!M || !coords || !out) MFAIL(RXM_ERR_INVALID, "null argument");
    if (!M->on_device) MFAIL(RXM_ERR_CUDA, "built without devices (host-only plan); there is no CPU fallback");
    bool src_b = (M->step & 1u) != 0;  // the attribute the next fused step reads
    MCK(multi_upload(M, coords, src_b));
    if (M->sh.size() == 1) {
        Shard& S = M->sh[0];
        MCU(cudaSetDevice(S.device));
        if (iters) MCK(rxm_laplacian_smooth(S.mesh, src_b ? S.b : S.a, src_b ? S.a : S.b, lr, iters, S.stream));
        return multi_download(M, iters ? !src_b : src_b, out);
    }
    // every step is ONE kernel per device; the devices keep in step among themselves (flag words in each other's memory, at
    // most one step apart), so the host only queues launches: one host thread per shard queues that shard's whole series --
    // a single thread switching devices paid ~10 us per launch and capped 2 devices at 1.5x on a 10 M-face mesh
    const int         n     = (int)M->sh.size();
    const uint32_t    step0 = M->step;
    std::vector<int>         rcs((size_t)n, RXM_OK);
    std::vector<std::string> msgs((size_t)n);
    auto run_shard = [&](int r) {
        Shard& S = M->sh[r];
        bool   sb = src_b;
        if (cudaSetDevice(S.device) != cudaSuccess) {
            rcs[r] = RXM_ERR_CUDA, msgs[r] = "cudaSetDevice failed";
            return;
        }
        for (uint32_t it = 0; it < iters && rcs[r] == RXM_OK; ++it, sb = !sb) {
            rcs[r] = rxm_laplacian_smooth_fused(S.mesh, sb ? S.b : S.a, sb ? S.a : S.b, lr, S.halo, sb ? 0 : 1, step0 + it, S.stream);
            if (rcs[r]) msgs[r] = rxm_last_error();  // the message is thread-local: carry it to the caller's thread
        }
    };
    {
        // plain threads, not an OpenMP team: every shard MUST have its own launching thread (a thread that queued one shard's
        // whole series before another shard's first kernel would fill the launch queue with kernels that wait for each other)
        std::vector<std::thread> team;
        for (int r = 1; r < n; ++r)
            team.emplace_back(run_shard, r);
        run_shard(0);
        for (auto& t : team)
            t.join();
    }
    for (int r = 0; r < n; ++r) {
        Shard& S = M->sh[r];
        if (rcs[r] == RXM_OK && (cudaSetDevice(S.device) != cudaSuccess || cudaStreamSynchronize(S.stream) != cudaSuccess))
            rcs[r] = RXM_ERR_CUDA, msgs[r] = "stream synchronize failed";
    }
    for (int r = 0; r < n; ++r)
        if (rcs[r]) return rxm_set_last_error(rcs[r], msgs[r].c_str());
    M->step += iters;
    if (iters & 1u) src_b = !src_b;
    return multi_download(M, src_b, out);
}

int rxm_multi_vertex_normals(rxm_multi* M, const float* coords, float* normals)
{
    if (!M || !coords || !normals) MFAIL(RXM_ERR_INVALID, "null argument");
    if (!M->on_device) MFAIL(RXM_ERR_CUDA, "built without devices (host-only plan); there is no CPU fallback");
    const bool src_b = (M->step & 1u) != 0;
    MCK(multi_upload(M, coords, src_b));  // ghost rows included: no exchange needed for a one-shot kernel
    for (auto& S : M->sh) {
        MCU(cudaSetDevice(S.device));
        MCK(rxm_vertex_normals(S.mesh, src_b ? S.b : S.a, src_b ? S.a : S.b, 0, S.stream));
    }
    return multi_download(M, !src_b, normals);
}

}  // extern "C"
