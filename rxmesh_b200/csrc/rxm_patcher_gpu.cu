// rxm_patcher_gpu.cu -- the Lloyd patcher's passes on the GPU (round 2).
//
// Reference: include/rxmesh/patcher/patcher.cu:828-987 (run_lloyd: cluster seed propagation, interior / boundary detection,
// seed relocation, all as kernels) and patcher_kernel.cuh:57-108.  mesh_builder.cpp:patcher_lloyd keeps the definition -- a
// FIFO multi-source BFS from the seeds ("assign"), a BFS from the patch borders whose deepest face becomes the new seed
// ("recentre"), extra seeds in oversized patches -- and its level-synchronous multi-core form.  This file runs exactly that
// level-synchronous form on the device, so the face -> patch array is THE SAME, bit for bit, as the host's (tested):
//   assign:   frontier position i claims an unassigned neighbour g with atomicMin(claim[g], i); the next frontier is written in
//             (claimer position, neighbour-list index) order through a prefix sum -- the order the FIFO queue would have had;
//   recentre: distances to the border do not depend on the visiting order: compare-and-swap claims, appended in any order;
//             per patch the deepest face wins, the smallest face id among equals (one 64-bit atomicMax per face);
//   split:    per oversized patch the face farthest from the seed, the smallest id among equals (the same 64-bit key).
// One level = 3 small kernels + a device scan; the host reads one counter per level.  100 M faces: ~25 assign / recentre passes
// of ~30 levels each, a few milliseconds per pass, against 2-3 s per pass on 16 host threads.
#include <cuda_runtime.h>

#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mesh_builder.h"

namespace rxm {
namespace {
constexpr uint32_t INV = 0xFFFFFFFFu;
constexpr int      TB  = 256;

#define PCK(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) {                                                                           \
            if (getenv("RXM_VERBOSE")) fprintf(stderr, "[rxmesh_b200] gpu patcher: %s: %s\n", #call, cudaGetErrorString(e_)); \
            return false;                                                                                  \
        }                                                                                                  \
    } while (0)

__global__ void k_fill2(uint32_t* a, uint32_t* b, uint32_t n, uint32_t v)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        a[i] = v;
        if (b) b[i] = v;
    }
}

// seeds: the first seed index that names a face keeps it (duplicates are dropped later as empty patches)
__global__ void k_seed_claim(const uint32_t* seeds, uint32_t k0, uint32_t k1, uint32_t* claim)
{
    const uint32_t i = k0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k1) atomicMin(&claim[seeds[i]], i);
}
__global__ void k_seed_flag(const uint32_t* seeds, uint32_t k0, uint32_t k1, const uint32_t* claim, const uint32_t* face_patch,
                            uint32_t* flag)
{
    const uint32_t i = k0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k1) flag[i - k0] = (face_patch[seeds[i]] == INV && claim[seeds[i]] == i) ? 1u : 0u;
}
__global__ void k_seed_emit(const uint32_t* seeds, uint32_t k0, uint32_t k1, const uint32_t* flag, const uint32_t* offs,
                            uint32_t tail, uint32_t* face_patch, uint32_t* dist, uint32_t* queue, uint32_t* claim)
{
    const uint32_t i = k0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k1) return;
    const uint32_t f = seeds[i];
    if (flag[i - k0]) {
        face_patch[f] = i, dist[f] = 0, queue[tail + offs[i - k0]] = f;
    }
    claim[f] = INV;  // positions of the BFS levels are claimed in the same array
}

// ---- assign: one BFS level over the frontier queue[lo, hi) ----
__global__ void k_assign_claim(const uint32_t* __restrict__ ff_off, const uint32_t* __restrict__ ff_val, const uint32_t* __restrict__ queue,
                               uint32_t lo, uint32_t hi, const uint32_t* __restrict__ face_patch, uint32_t* claim)
{
    const uint32_t i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const uint32_t f = queue[i];
    for (uint32_t k = ff_off[f]; k < ff_off[f + 1]; ++k) {
        const uint32_t g = ff_val[k];
        if (face_patch[g] == INV) atomicMin(&claim[g], i);
    }
}
template <bool EMIT>
__global__ void k_assign_emit(const uint32_t* __restrict__ ff_off, const uint32_t* __restrict__ ff_val, uint32_t* queue, uint32_t lo,
                              uint32_t hi, const uint32_t* __restrict__ claim, uint32_t* cnt_or_offs, uint32_t* face_patch, uint32_t* dist)
{
    const uint32_t i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const uint32_t f = queue[i], b = ff_off[f], e = ff_off[f + 1];
    uint32_t       n = 0, w = EMIT ? hi + cnt_or_offs[i - lo] : 0u;
    const uint32_t pf = EMIT ? face_patch[f] : 0u, df = EMIT ? dist[f] + 1u : 0u;
    for (uint32_t k = b; k < e; ++k) {
        const uint32_t g = ff_val[k];
        if (claim[g] != i) continue;
        bool again = false;  // a face listed twice (two shared edges) is taken at its first occurrence
        for (uint32_t k2 = b; k2 < k; ++k2)
            again |= (ff_val[k2] == g);
        if (again) continue;
        if (EMIT) face_patch[g] = pf, dist[g] = df, queue[w++] = g;
        ++n;
    }
    if (!EMIT) cnt_or_offs[i - lo] = n;
}

// first face >= from that no seed reached (components without a seed)
__global__ void k_first_unassigned(const uint32_t* face_patch, uint32_t from, uint32_t nf, uint32_t* out)
{
    for (uint32_t f = from + blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x)
        if (face_patch[f] == INV) {
            atomicMin(out, f);
            return;  // this thread's later faces are larger
        }
}

__global__ void k_histogram(const uint32_t* face_patch, uint32_t nf, uint32_t* psize)
{
    for (uint32_t f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x)
        atomicAdd(&psize[face_patch[f]], 1u);
}
__global__ void k_remap(uint32_t* face_patch, uint32_t nf, const uint32_t* remap)
{
    for (uint32_t f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x)
        face_patch[f] = remap[face_patch[f]];
}

// ---- recentre ----
__global__ void k_border(const uint32_t* __restrict__ ff_off, const uint32_t* __restrict__ ff_val, const uint32_t* __restrict__ face_patch,
                         uint32_t nf, uint32_t* depth, uint32_t* queue, uint32_t* tail)
{
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    bool           border = false;
    if (f < nf) {
        const uint32_t p = face_patch[f];
        for (uint32_t k = ff_off[f]; k < ff_off[f + 1]; ++k)
            border |= (face_patch[ff_val[k]] != p);
        depth[f] = border ? 0u : INV;
    }
    // warp-aggregated append (the order inside a level does not matter here)
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, border);
    if (m) {
        const uint32_t lane = threadIdx.x & 31u, leader = __ffs(m) - 1;
        uint32_t       base = 0;
        if (lane == leader) base = atomicAdd(tail, __popc(m));
        base = __shfl_sync(0xFFFFFFFFu, base, leader);
        if (border) queue[base + __popc(m & ((1u << lane) - 1u))] = f;
    }
}
__global__ void k_depth_level(const uint32_t* __restrict__ ff_off, const uint32_t* __restrict__ ff_val, const uint32_t* __restrict__ face_patch,
                              uint32_t* queue, uint32_t lo, uint32_t hi, uint32_t* depth, uint32_t* tail)
{
    const uint32_t i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const uint32_t f = queue[i], p = face_patch[f], d1 = depth[f] + 1u;
    for (uint32_t k = ff_off[f]; k < ff_off[f + 1]; ++k) {
        const uint32_t g = ff_val[k];
        if (face_patch[g] != p) continue;
        if (depth[g] == INV && atomicCAS(&depth[g], INV, d1) == INV) queue[atomicAdd(tail, 1u)] = g;
    }
}
// per patch: the face with the largest value, the smallest face id among equals
__global__ void k_argmax_per_patch(const uint32_t* __restrict__ face_patch, const uint32_t* __restrict__ value, uint32_t nf,
                                   const uint32_t* __restrict__ psize, uint32_t only_above, unsigned long long* best)
{
    for (uint32_t f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x) {
        const uint32_t v = value[f], p = face_patch[f];
        if (v == INV) continue;                          // recentre: patch without border keeps its seed
        if (psize && psize[p] <= only_above) continue;   // split: oversized patches only
        const unsigned long long key = ((unsigned long long)v << 32) | (unsigned long long)(INV - f);
        if (key > best[p]) atomicMax(&best[p], key);
    }
}

struct DevBuf
{
    void* p = nullptr;
    ~DevBuf()
    {
        if (p) cudaFree(p);
    }
    template <class T>
    T* as()
    {
        return (T*)p;
    }
    bool alloc(size_t bytes) { return cudaMalloc(&p, std::max<size_t>(bytes, 16)) == cudaSuccess; }
};
inline uint32_t grid_for(uint64_t n)
{
    return (uint32_t)std::min<uint64_t>((n + TB - 1) / TB, 148u * 32u);
}
}  // namespace

// Runs the outer Lloyd loop of patcher_lloyd (everything between the initial seeds and the last-resort chop) on the current
// device.  In: face adjacency CSR, initial seeds.  Out (host): face_patch, queue (BFS order of the last assign), psize, seeds.
// false = not run (no device, out of memory, a mesh of very many components): the caller runs the host passes instead.
bool patcher_lloyd_gpu(const std::vector<uint32_t>& ff_off, const std::vector<uint32_t>& ff_val, uint32_t nf, uint32_t patch_size,
                       uint32_t lloyd_iters, std::vector<uint32_t>& seeds, std::vector<uint32_t>& face_patch,
                       std::vector<uint32_t>& queue, std::vector<uint32_t>& psize, int* n_assign_out)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return false;
    }
    size_t free_b = 0, total_b = 0;
    PCK(cudaMemGetInfo(&free_b, &total_b));
    const size_t need = 4ull * (ff_off.size() + ff_val.size() + 8ull * nf) + (64u << 20);
    if (need > free_b) return false;
    DevBuf d_off, d_val, d_fp, d_dist, d_claim, d_queue, d_depth, d_cnt, d_offs, d_seeds, d_psize, d_best, d_remap, d_tmp, d_ctr;
    uint32_t seed_cap = 0;
    if (!d_off.alloc(4 * ff_off.size()) || !d_val.alloc(4 * ff_val.size()) || !d_fp.alloc(4ull * nf) || !d_dist.alloc(4ull * nf) ||
        !d_claim.alloc(4ull * nf) || !d_queue.alloc(4ull * nf) || !d_depth.alloc(4ull * nf) || !d_cnt.alloc(4ull * nf) ||
        !d_offs.alloc(4ull * nf) || !d_ctr.alloc(16)) {
        cudaGetLastError();
        return false;
    }
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_cnt.as<uint32_t>(), d_offs.as<uint32_t>(), (int)nf);
    if (!d_tmp.alloc(tmp_bytes)) return false;
    PCK(cudaMemcpy(d_off.p, ff_off.data(), 4 * ff_off.size(), cudaMemcpyHostToDevice));
    PCK(cudaMemcpy(d_val.p, ff_val.data(), 4 * ff_val.size(), cudaMemcpyHostToDevice));
    const uint32_t *off = d_off.as<uint32_t>(), *val = d_val.as<uint32_t>();
    uint32_t *      fp = d_fp.as<uint32_t>(), *dist = d_dist.as<uint32_t>(), *claim = d_claim.as<uint32_t>(), *q = d_queue.as<uint32_t>(),
             *depth = d_depth.as<uint32_t>(), *cnt = d_cnt.as<uint32_t>(), *offs = d_offs.as<uint32_t>(), *ctr = d_ctr.as<uint32_t>();

    auto ensure_seed_cap = [&](uint32_t n) -> bool {
        if (n <= seed_cap) return true;
        seed_cap = std::max<uint32_t>(2 * n, 1024);
        for (DevBuf* b : {&d_seeds, &d_psize, &d_remap}) {
            if (b->p) cudaFree(b->p);
            b->p = nullptr;
            if (!b->alloc(4ull * seed_cap)) return false;
        }
        if (d_best.p) cudaFree(d_best.p);
        d_best.p = nullptr;
        return d_best.alloc(8ull * seed_cap);
    };
    auto scan = [&](uint32_t n) -> bool {
        size_t tb = tmp_bytes;
        return cub::DeviceScan::ExclusiveSum(d_tmp.p, tb, cnt, offs, (int)n) == cudaSuccess;
    };
    auto last_sum = [&](uint32_t n, uint32_t& total) -> bool {  // offs[n-1] + cnt[n-1]
        uint32_t a = 0, b = 0;
        if (cudaMemcpy(&a, offs + n - 1, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return false;
        if (cudaMemcpy(&b, cnt + n - 1, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return false;
        total = a + b;
        return true;
    };

    // enqueue seeds [k0, k1) (host vector already holds them) at the queue's tail
    auto push_seeds = [&](uint32_t k0, uint32_t k1, uint32_t& tail) -> bool {
        const uint32_t n = k1 - k0;
        if (!n) return true;
        k_seed_claim<<<(n + TB - 1) / TB, TB>>>(d_seeds.as<uint32_t>(), k0, k1, claim);
        k_seed_flag<<<(n + TB - 1) / TB, TB>>>(d_seeds.as<uint32_t>(), k0, k1, claim, fp, cnt);
        if (!scan(n)) return false;
        uint32_t total = 0;
        if (!last_sum(n, total)) return false;
        k_seed_emit<<<(n + TB - 1) / TB, TB>>>(d_seeds.as<uint32_t>(), k0, k1, cnt, offs, tail, fp, dist, q, claim);
        tail += total;
        return true;
    };

    int  n_assign = 0;
    auto assign   = [&]() -> bool {
        if (!ensure_seed_cap((uint32_t)seeds.size() + 64)) return false;
        PCK(cudaMemcpy(d_seeds.p, seeds.data(), 4 * seeds.size(), cudaMemcpyHostToDevice));
        k_fill2<<<grid_for(nf), TB>>>(fp, claim, nf, INV);
        uint32_t tail = 0, lo = 0, scan_from = 0;
        if (!push_seeds(0, (uint32_t)seeds.size(), tail)) return false;
        int extra_components = 0;
        while (true) {
            while (lo < tail) {
                const uint32_t hi = tail, n = hi - lo, g = (n + TB - 1) / TB;
                k_assign_claim<<<g, TB>>>(off, val, q, lo, hi, fp, claim);
                k_assign_emit<false><<<g, TB>>>(off, val, q, lo, hi, claim, cnt, fp, dist);
                if (!scan(n)) return false;
                uint32_t total = 0;
                if (!last_sum(n, total)) return false;
                if (total) k_assign_emit<true><<<g, TB>>>(off, val, q, lo, hi, claim, offs, fp, dist);
                lo = hi, tail += total;
            }
            if (tail == nf) break;
            // components no seed reached get a seed of their own (first unassigned face, ascending)
            if (++extra_components > 256) return false;  // a mesh of very many pieces: the host loop handles it better
            uint32_t first = INV;
            PCK(cudaMemcpy(ctr, &first, 4, cudaMemcpyHostToDevice));
            k_first_unassigned<<<grid_for(nf - scan_from), TB>>>(fp, scan_from, nf, ctr);
            PCK(cudaMemcpy(&first, ctr, 4, cudaMemcpyDeviceToHost));
            if (first == INV) return false;  // cannot happen (tail < nf)
            scan_from = first;
            seeds.push_back(first);
            if (!ensure_seed_cap((uint32_t)seeds.size() + 64)) return false;
            PCK(cudaMemcpy(d_seeds.p, seeds.data(), 4 * seeds.size(), cudaMemcpyHostToDevice));
            if (!push_seeds((uint32_t)seeds.size() - 1, (uint32_t)seeds.size(), tail)) return false;
        }
        // drop_empty_patches
        const uint32_t K = (uint32_t)seeds.size();
        PCK(cudaMemset(d_psize.p, 0, 4ull * K));
        k_histogram<<<grid_for(nf), TB>>>(fp, nf, d_psize.as<uint32_t>());
        psize.resize(K);
        PCK(cudaMemcpy(psize.data(), d_psize.p, 4ull * K, cudaMemcpyDeviceToHost));
        std::vector<uint32_t> remap(K);
        uint32_t              k = 0;
        for (uint32_t i = 0; i < K; ++i) {
            remap[i] = k;
            if (psize[i]) seeds[k] = seeds[i], psize[k] = psize[i], ++k;
        }
        if (k != K) {
            seeds.resize(k), psize.resize(k);
            PCK(cudaMemcpy(d_remap.p, remap.data(), 4ull * K, cudaMemcpyHostToDevice));
            k_remap<<<grid_for(nf), TB>>>(fp, nf, d_remap.as<uint32_t>());
            PCK(cudaMemcpy(d_psize.p, psize.data(), 4ull * k, cudaMemcpyHostToDevice));
        }
        ++n_assign;
        return true;
    };

    auto argmax = [&](const uint32_t* value, bool oversized_only, std::vector<unsigned long long>& best) -> bool {
        const uint32_t K = (uint32_t)seeds.size();
        PCK(cudaMemset(d_best.p, 0, 8ull * K));
        k_argmax_per_patch<<<grid_for(nf), TB>>>(fp, value, nf, oversized_only ? d_psize.as<uint32_t>() : nullptr, patch_size,
                                                 d_best.as<unsigned long long>());
        best.resize(K);
        PCK(cudaMemcpy(best.data(), d_best.p, 8ull * K, cudaMemcpyDeviceToHost));
        return true;
    };

    int  rc_recenter = 0;  // 0 unchanged, 1 changed, -1 error
    auto recenter    = [&]() -> int {
        uint32_t zero = 0;
        if (cudaMemcpy(ctr, &zero, 4, cudaMemcpyHostToDevice) != cudaSuccess) return -1;
        k_border<<<(nf + TB - 1) / TB, TB>>>(off, val, fp, nf, depth, q, ctr);
        uint32_t lo = 0, tail = 0;
        if (cudaMemcpy(&tail, ctr, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        while (lo < tail) {
            const uint32_t hi = tail;
            k_depth_level<<<(hi - lo + TB - 1) / TB, TB>>>(off, val, fp, q, lo, hi, depth, ctr);
            if (cudaMemcpy(&tail, ctr, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
            lo = hi;
        }
        std::vector<unsigned long long> best;
        if (!argmax(depth, false, best)) return -1;
        bool changed = false;
        for (uint32_t p = 0; p < seeds.size(); ++p)
            if (best[p]) {  // key 0 = no face with a depth (a face with depth 0 and id INV - 0... has key >= 1 unless f == INV)
                const uint32_t b = INV - (uint32_t)(best[p] & 0xFFFFFFFFull);
                if (b != seeds[p]) seeds[p] = b, changed = true;
            }
        return changed ? 1 : 0;
    };
    (void)rc_recenter;

    for (int outer = 0; outer < 64; ++outer) {
        if (!assign()) return false;
        for (uint32_t it = 0; it < lloyd_iters; ++it) {
            const int r = recenter();
            if (r < 0) return false;
            if (r == 0) break;
            if (!assign()) return false;
        }
        // split patches that are still too large: one more seed at the face farthest from the current seed
        bool over = false;
        for (uint32_t p = 0; p < psize.size(); ++p)
            over |= psize[p] > patch_size;
        if (!over) break;
        std::vector<unsigned long long> best;
        if (!argmax(dist, true, best)) return false;
        const uint32_t K0  = (uint32_t)seeds.size();
        bool           any = false;
        for (uint32_t p = 0; p < K0; ++p)
            if (psize[p] > patch_size && best[p]) {
                const uint32_t far = INV - (uint32_t)(best[p] & 0xFFFFFFFFull);
                if (far != seeds[p]) seeds.push_back(far), any = true;
            }
        if (!any) break;
    }
    if (!assign()) return false;
    face_patch.resize(nf), queue.resize(nf);
    PCK(cudaMemcpy(face_patch.data(), fp, 4ull * nf, cudaMemcpyDeviceToHost));
    PCK(cudaMemcpy(queue.data(), q, 4ull * nf, cudaMemcpyDeviceToHost));
    PCK(cudaDeviceSynchronize());
    if (n_assign_out) *n_assign_out = n_assign;
    return true;
}

}  // namespace rxm
