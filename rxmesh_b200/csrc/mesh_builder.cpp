// mesh_builder.cpp -- see mesh_builder.h.  Host C++17 + OpenMP, flat arrays only.
#include "mesh_builder.h"

#include <omp.h>

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iterator>
#include <numeric>

namespace rxm {

namespace {
double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch())
        .count();
}

// exclusive prefix sum in place over n+1 entries (entry n receives the total)
template <typename T, typename A>
void exclusive_scan(std::vector<T, A>& a)
{
    T run = 0;
    for (size_t i = 0; i < a.size(); ++i) {
        T c  = a[i];
        a[i] = run;
        run += c;
    }
}
}  // namespace

// ---------------------------------------------------------------------------
// Edge numbering (reference semantics: rxmesh.cpp:589-611, util/util.h:410-416)
// ---------------------------------------------------------------------------
uint32_t build_edges(const uint32_t* fv, uint32_t nf, uint32_t nv, U32Buf& ev, U32Buf& fe)
{
    const uint64_t H = 3ull * nf;
    fe.resize(H);  // uninitialised: every entry is written by the last pass
    if (nf == 0) {
        ev.clear();
        return 0;
    }
    // 1. bucket half-edges by their smaller endpoint
    U32Buf off((size_t)nv + 1);
#pragma omp parallel for schedule(static)
    for (int64_t v = 0; v <= (int64_t)nv; ++v)
        off[v] = 0;
#pragma omp parallel for schedule(static)
    for (int64_t h = 0; h < (int64_t)H; ++h) {
        const uint32_t f = (uint32_t)(h / 3), j = (uint32_t)(h % 3);
        const uint32_t a = fv[3ull * f + j], b = fv[3ull * f + (j + 1) % 3];
        const uint32_t lo = a < b ? a : b;
        __atomic_fetch_add(&off[lo], 1u, __ATOMIC_RELAXED);
    }
    exclusive_scan(off);
    struct HE
    {
        uint32_t hi, h;
    };
    std::vector<HE, NoInitAlloc<HE>> bucket(H);  // every slot is written by the scatter below
    U32Buf cur((size_t)nv);
#pragma omp parallel for schedule(static)
    for (int64_t v = 0; v < (int64_t)nv; ++v)
        cur[v] = off[v];
#pragma omp parallel for schedule(static)
    for (int64_t h = 0; h < (int64_t)H; ++h) {
        const uint32_t f = (uint32_t)(h / 3), j = (uint32_t)(h % 3);
        const uint32_t a = fv[3ull * f + j], b = fv[3ull * f + (j + 1) % 3];
        const uint32_t lo = a < b ? a : b, hi = a < b ? b : a;
        const uint32_t p  = __atomic_fetch_add(&cur[lo], 1u, __ATOMIC_RELAXED);
        bucket[p]         = {hi, (uint32_t)h};
    }
    // 2. inside a bucket, equal `hi` = same edge; its first half-edge (smallest h)
    //    decides the edge id.
    std::vector<uint32_t, NoInitAlloc<uint32_t>> rep(H);        // half-edge -> first half-edge of its edge (all written)
    std::vector<uint32_t, NoInitAlloc<uint32_t>> first(H + 1);  // 1 at the first half-edge of every edge
#pragma omp parallel for schedule(static)
    for (int64_t h = 0; h <= (int64_t)H; ++h)
        first[h] = 0;
#pragma omp parallel for schedule(dynamic, 4096)
    for (int64_t v = 0; v < (int64_t)nv; ++v) {
        HE* b = bucket.data() + off[v];
        HE* e = bucket.data() + off[v + 1];
        std::sort(b, e, [](const HE& x, const HE& y) { return x.hi != y.hi ? x.hi < y.hi : x.h < y.h; });
        for (HE* g = b; g < e;) {
            HE* g2 = g;
            while (g2 < e && g2->hi == g->hi) {
                rep[g2->h] = g->h;
                ++g2;
            }
            first[g->h] = 1;
            g           = g2;
        }
    }
    exclusive_scan(first);
    const uint32_t ne = first[H];
    ev.resize(2ull * ne);  // uninitialised: the first half-edge of every edge writes its pair
#pragma omp parallel for schedule(static)
    for (int64_t h = 0; h < (int64_t)H; ++h) {
        const uint32_t e = first[rep[h]];
        fe[h]            = e;
        if (rep[h] == (uint32_t)h) {
            const uint32_t f = (uint32_t)(h / 3), j = (uint32_t)(h % 3);
            const uint32_t a = fv[3ull * f + j], b = fv[3ull * f + (j + 1) % 3];
            ev[2ull * e]     = a < b ? b : a;
            ev[2ull * e + 1] = a < b ? a : b;
        }
    }
    return ne;
}

// ---------------------------------------------------------------------------
// Lloyd patcher
// ---------------------------------------------------------------------------
void patcher_lloyd(const uint32_t* fe, uint32_t nf, uint32_t ne, uint32_t patch_size,
                   uint32_t lloyd_iters, std::vector<uint32_t>& face_patch, uint32_t& num_patches, uint32_t* lloyd_runs)
{
    if (lloyd_runs) *lloyd_runs = 0;
    face_patch.assign(nf, INVALID32_);
    num_patches = 0;
    if (nf == 0) return;
    // edge -> faces CSR
    // (all cores: counts and cursors by relaxed atomics, then every edge's short list is sorted -- the ascending face order
    // the serial fill in face order produces; 100 M faces: 3.5 s -> under 1 s of the patcher's time)
    U32Buf ef_off((size_t)ne + 1), ef_val(3ull * nf);
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e <= (int64_t)ne; ++e)
        ef_off[e] = 0;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)(3ull * nf); ++i)
        __atomic_fetch_add(&ef_off[fe[i]], 1u, __ATOMIC_RELAXED);
    exclusive_scan(ef_off);
    {
        U32Buf cur((size_t)ne);
#pragma omp parallel for schedule(static)
        for (int64_t e = 0; e < (int64_t)ne; ++e)
            cur[e] = ef_off[e];
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < (int64_t)(3ull * nf); ++i)
            ef_val[__atomic_fetch_add(&cur[fe[i]], 1u, __ATOMIC_RELAXED)] = (uint32_t)(i / 3);
#pragma omp parallel for schedule(static)
        for (int64_t e = 0; e < (int64_t)ne; ++e) {
            uint32_t* b = &ef_val[ef_off[e]];
            const uint32_t n = ef_off[e + 1] - ef_off[e];
            if (n == 2) {
                if (b[1] < b[0]) std::swap(b[0], b[1]);
            } else if (n > 2) {
                std::sort(b, b + n);
            }
        }
    }
    // face -> adjacent faces CSR (one indirection per visit instead of face -> edge -> faces)
    std::vector<uint32_t> ff_off((size_t)nf + 1, 0);
#pragma omp parallel for schedule(static)
    for (int64_t f = 0; f < (int64_t)nf; ++f) {
        uint32_t k = 0;
        for (int j = 0; j < 3; ++j) {
            const uint32_t e = fe[3ull * f + j];
            k += ef_off[e + 1] - ef_off[e] - 1;
        }
        ff_off[f + 1] = k;
    }
    for (uint32_t f = 0; f < nf; ++f)
        ff_off[f + 1] += ff_off[f];
    std::vector<uint32_t> ff_val(ff_off[nf]);
#pragma omp parallel for schedule(static)
    for (int64_t f = 0; f < (int64_t)nf; ++f) {
        uint32_t w = ff_off[f];
        for (int j = 0; j < 3; ++j) {
            const uint32_t e = fe[3ull * f + j];
            for (uint32_t i = ef_off[e]; i < ef_off[e + 1]; ++i)
                if (ef_val[i] != (uint32_t)f) ff_val[w++] = ef_val[i];
        }
    }
    U32Buf().swap(ef_off);
    U32Buf().swap(ef_val);
    auto for_each_nbr = [&](uint32_t f, auto&& fn) {
        for (uint32_t i = ff_off[f]; i < ff_off[f + 1]; ++i)
            fn(ff_val[i]);
    };

    // start near the patch count the size bound ends up needing (the reference
    // converges to ~F/370 for patch_size 512)
    uint32_t K = std::max<uint32_t>(1, (uint32_t)((uint64_t)nf * 4 / (3ull * patch_size)) + 1);
    K          = std::min(K, nf);
    std::vector<uint32_t> seeds(K);
    for (uint32_t i = 0; i < K; ++i)
        seeds[i] = (uint32_t)(((2ull * i + 1) * nf) / (2ull * K));

    std::vector<uint32_t> queue(nf), dist(nf), psize;
    // RXM_PATCHER_SERIAL=1 keeps the single-threaded FIFO passes (the definition the parallel ones are tested against;
    // also what runs with fewer than three threads, where the level-synchronous form does not pay)
    const bool serial = getenv("RXM_PATCHER_SERIAL") != nullptr || (omp_get_max_threads() < 3 && !getenv("RXM_PATCHER_PARALLEL"));

    // compact away patches that received no face (duplicate seeds)
    auto drop_empty_patches = [&]() {
        psize.assign(seeds.size(), 0);
#pragma omp parallel
        {
            std::vector<uint32_t> mine(seeds.size(), 0);  // per-thread histogram, merged under a lock: counts only
#pragma omp for schedule(static) nowait
            for (int64_t f = 0; f < (int64_t)nf; ++f)
                mine[face_patch[f]]++;
#pragma omp critical
            for (size_t i = 0; i < mine.size(); ++i)
                psize[i] += mine[i];
        }
        std::vector<uint32_t> remap(seeds.size());
        uint32_t              k = 0;
        for (uint32_t i = 0; i < seeds.size(); ++i) {
            remap[i] = k;
            if (psize[i]) {
                seeds[k] = seeds[i];
                psize[k] = psize[i];
                ++k;
            }
        }
        if (k != seeds.size()) {
            seeds.resize(k);
            psize.resize(k);
#pragma omp parallel for schedule(static)
            for (int64_t f = 0; f < (int64_t)nf; ++f)
                face_patch[f] = remap[face_patch[f]];
        }
    };

    // Multi-source BFS from the seeds, level by level on all cores, with exactly the outcome of the FIFO version below:
    // an unvisited face goes to the frontier face with the smallest queue position among its neighbours (atomic min of
    // positions), and the next frontier is written in (claimer position, neighbour-list index) order through a prefix sum
    // -- the order the FIFO queue would have produced.  Same face_patch, dist and queue contents for any thread count.
    std::vector<uint32_t>              claim, emit_n((size_t)omp_get_max_threads(), 0);
    std::vector<std::vector<uint32_t>> emit_buf((size_t)omp_get_max_threads());
    auto assign_parallel = [&]() {
        claim.resize(nf);
#pragma omp parallel for schedule(static)
        for (int64_t f = 0; f < (int64_t)nf; ++f)
            face_patch[f] = INVALID32_, claim[f] = INVALID32_;
        uint32_t tail = 0;
        for (uint32_t i = 0; i < seeds.size(); ++i) {
            if (face_patch[seeds[i]] != INVALID32_) continue;  // duplicate seed: dropped below
            face_patch[seeds[i]] = i;
            dist[seeds[i]]       = 0;
            queue[tail++]        = seeds[i];
        }
        uint32_t lo = 0, scan = 0;
        while (true) {
            while (lo < tail) {
                const uint32_t hi = tail, n = hi - lo;
                std::fill(emit_n.begin(), emit_n.end(), 0u);
#pragma omp parallel if (n > 2048)
                {
                    // thread t owns the contiguous positions [i0, i1): concatenating the threads' outputs in thread order
                    // is the (claimer position, neighbour index) order
                    const int      t = omp_get_thread_num(), T = omp_get_num_threads();
                    const uint32_t i0 = lo + (uint32_t)((uint64_t)n * t / T), i1 = lo + (uint32_t)((uint64_t)n * (t + 1) / T);
                    for (uint32_t i = i0; i < i1; ++i) {
                        const uint32_t f = queue[i];
                        for (uint32_t k = ff_off[f]; k < ff_off[f + 1]; ++k) {
                            const uint32_t g = ff_val[k];
                            if (face_patch[g] != INVALID32_) continue;
                            uint32_t old = __atomic_load_n(&claim[g], __ATOMIC_RELAXED);
                            while (i < old && !__atomic_compare_exchange_n(&claim[g], &old, i, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
                            }
                        }
                    }
#pragma omp barrier
                    std::vector<uint32_t>& mine = emit_buf[t];
                    mine.clear();
                    for (uint32_t i = i0; i < i1; ++i) {
                        const uint32_t f = queue[i];
                        for (uint32_t k = ff_off[f]; k < ff_off[f + 1]; ++k) {
                            const uint32_t g = ff_val[k];
                            if (claim[g] != i) continue;
                            bool again = false;  // a face listed twice (two shared edges) is taken at its first occurrence
                            for (uint32_t k2 = ff_off[f]; k2 < k; ++k2)
                                again |= (ff_val[k2] == g);
                            if (again) continue;
                            face_patch[g] = face_patch[f];
                            dist[g]       = dist[f] + 1;
                            mine.push_back(g);
                        }
                    }
                    emit_n[t] = (uint32_t)mine.size();
#pragma omp barrier
                    uint32_t w = hi;
                    for (int u = 0; u < t; ++u)
                        w += emit_n[u];
                    std::copy(mine.begin(), mine.end(), queue.begin() + w);
                }
                lo = hi;
                for (uint32_t c : emit_n)
                    tail += c;
            }
            // components no seed reached get a seed of their own
            while (scan < nf && face_patch[scan] != INVALID32_)
                ++scan;
            if (scan == nf) break;
            seeds.push_back(scan);
            face_patch[scan] = (uint32_t)seeds.size() - 1;
            dist[scan]       = 0;
            queue[tail++]    = scan;
        }
        drop_empty_patches();
    };

    auto assign_serial = [&]() {
        std::fill(face_patch.begin(), face_patch.end(), INVALID32_);
        uint32_t head = 0, tail = 0;
        for (uint32_t i = 0; i < seeds.size(); ++i) {
            if (face_patch[seeds[i]] != INVALID32_) continue;  // duplicate seed: dropped below
            face_patch[seeds[i]] = i;
            dist[seeds[i]]       = 0;
            queue[tail++]        = seeds[i];
        }
        uint32_t scan = 0;
        while (true) {
            while (head < tail) {
                const uint32_t f = queue[head++];
                for_each_nbr(f, [&](uint32_t g) {
                    if (face_patch[g] == INVALID32_) {
                        face_patch[g] = face_patch[f];
                        dist[g]       = dist[f] + 1;
                        queue[tail++] = g;
                    }
                });
            }
            // components no seed reached get a seed of their own
            while (scan < nf && face_patch[scan] != INVALID32_)
                ++scan;
            if (scan == nf) break;
            seeds.push_back(scan);
            face_patch[scan] = (uint32_t)seeds.size() - 1;
            dist[scan]       = 0;
            queue[tail++]    = scan;
        }
        drop_empty_patches();
    };
    double t_assign = 0, t_recenter = 0;  // reported with RXM_VERBOSE
    auto   assign   = [&]() {
        const double t0 = now_s();
        if (serial)
            assign_serial();
        else
            assign_parallel();
        t_assign += now_s() - t0;
    };

    std::vector<uint32_t> depth(nf);
    // parallel form of recenter_serial: BFS distances do not depend on the visiting order, so every level claims its
    // faces with a compare-and-swap and appends them in any order; the per-patch arg-max keeps the smallest face id
    // among equals by merging per-thread results in thread (= ascending face) order
    auto recenter_parallel = [&]() -> bool {
        uint32_t tail = 0;
#pragma omp parallel for schedule(static)
        for (int64_t f = 0; f < (int64_t)nf; ++f) {
            bool border = false;
            for (uint32_t k = ff_off[f]; k < ff_off[f + 1]; ++k)
                border |= (face_patch[ff_val[k]] != face_patch[f]);
            depth[f] = border ? 0u : INVALID32_;
        }
        for (uint32_t f = 0; f < nf; ++f)
            if (depth[f] == 0) queue[tail++] = f;
        uint32_t lo = 0;
        while (lo < tail) {
            const uint32_t hi = tail;
#pragma omp parallel if (hi - lo > 2048)
            {
                std::vector<uint32_t> mine;  // this thread's share of the next level, appended with ONE reservation
#pragma omp for schedule(static) nowait
                for (int64_t i = lo; i < (int64_t)hi; ++i) {
                    const uint32_t f = queue[i], d1 = depth[f] + 1;
                    for (uint32_t k = ff_off[f]; k < ff_off[f + 1]; ++k) {
                        const uint32_t g = ff_val[k];
                        if (face_patch[g] != face_patch[f]) continue;
                        uint32_t expect = INVALID32_;
                        if (__atomic_load_n(&depth[g], __ATOMIC_RELAXED) == INVALID32_ &&
                            __atomic_compare_exchange_n(&depth[g], &expect, d1, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED))
                            mine.push_back(g);
                    }
                }
                if (!mine.empty()) {
                    const uint32_t w = __atomic_fetch_add(&tail, (uint32_t)mine.size(), __ATOMIC_RELAXED);
                    std::copy(mine.begin(), mine.end(), queue.begin() + w);
                }
            }
            lo = hi;
        }
        const int                          nt = omp_get_max_threads();
        std::vector<std::vector<uint32_t>> best_t(nt);
#pragma omp parallel num_threads(nt)
        {
            const int              t = omp_get_thread_num(), T = omp_get_num_threads();
            std::vector<uint32_t>& b = best_t[t];
            b.assign(seeds.size(), INVALID32_);
            const uint64_t f0 = (uint64_t)nf * t / T, f1 = (uint64_t)nf * (t + 1) / T;
            for (uint64_t f = f0; f < f1; ++f) {
                if (depth[f] == INVALID32_) continue;  // patch without border: keep its seed
                uint32_t& x = b[face_patch[f]];
                if (x == INVALID32_ || depth[f] > depth[x]) x = (uint32_t)f;
            }
        }
        bool changed = false;
        for (uint32_t p = 0; p < seeds.size(); ++p) {
            uint32_t b = INVALID32_;
            for (int t = 0; t < nt; ++t) {
                if (best_t[t].empty()) continue;  // fewer threads ran than asked for
                const uint32_t x = best_t[t][p];
                if (x != INVALID32_ && (b == INVALID32_ || depth[x] > depth[b])) b = x;
            }
            if (b != INVALID32_ && b != seeds[p]) {
                seeds[p] = b;
                changed  = true;
            }
        }
        return changed;
    };
    auto recenter_serial = [&]() -> bool {
        // distance of every face to its patch boundary; the deepest face becomes the seed
        uint32_t head = 0, tail = 0;
        std::fill(depth.begin(), depth.end(), INVALID32_);
        for (uint32_t f = 0; f < nf; ++f) {
            bool border = false;
            for_each_nbr(f, [&](uint32_t g) { border |= (face_patch[g] != face_patch[f]); });
            if (border) {
                depth[f]      = 0;
                queue[tail++] = f;
            }
        }
        while (head < tail) {
            const uint32_t f = queue[head++];
            for_each_nbr(f, [&](uint32_t g) {
                if (depth[g] == INVALID32_ && face_patch[g] == face_patch[f]) {
                    depth[g]      = depth[f] + 1;
                    queue[tail++] = g;
                }
            });
        }
        std::vector<uint32_t> best(seeds.size(), INVALID32_);
        for (uint32_t f = 0; f < nf; ++f) {
            if (depth[f] == INVALID32_) continue;  // patch without border: keep its seed
            uint32_t& b = best[face_patch[f]];
            if (b == INVALID32_ || depth[f] > depth[b]) b = f;
        }
        bool changed = false;
        for (uint32_t p = 0; p < seeds.size(); ++p)
            if (best[p] != INVALID32_ && best[p] != seeds[p]) {
                seeds[p] = best[p];
                changed  = true;
            }
        return changed;
    };
    auto recenter = [&]() -> bool {
        const double t0 = now_s();
        const bool   r  = serial ? recenter_serial() : recenter_parallel();
        t_recenter += now_s() - t0;
        return r;
    };

    int n_assign = 0, n_outer = 0;
    // the same passes on the GPU when there is one and the mesh is large enough to pay for the launches
    // (rxm_patcher_gpu.cu: identical face -> patch array); RXM_PATCHER_GPU=0 / 1 forces the choice
    bool on_gpu = false;
    {
        const char* g    = getenv("RXM_PATCHER_GPU");
        const bool  want = g ? atoi(g) != 0 : nf >= 500000u;
        if (want && !getenv("RXM_PATCHER_SERIAL")) {
            const std::vector<uint32_t> seeds0 = seeds;
            const double                t0     = now_s();
            on_gpu = patcher_lloyd_gpu(ff_off, ff_val, nf, patch_size, lloyd_iters, seeds, face_patch, queue, psize, &n_assign);
            if (!on_gpu) seeds = seeds0;
            t_assign += now_s() - t0;
        }
    }
    for (int outer = 0; outer < 64 && !on_gpu; ++outer) {
        ++n_outer;
        assign();
        for (uint32_t it = 0; it < lloyd_iters; ++it) {
            if (!recenter()) break;
            assign();
            ++n_assign;
        }
        // split patches that are still too large: one more seed at the face
        // farthest from the current seed
        std::vector<uint32_t> far(seeds.size(), INVALID32_);
        bool                  any = false;
        for (uint32_t f = 0; f < nf; ++f) {
            const uint32_t p = face_patch[f];
            if (psize[p] <= patch_size) continue;
            if (far[p] == INVALID32_ || dist[f] > dist[far[p]]) far[p] = f;
        }
        const uint32_t K0 = (uint32_t)seeds.size();
        for (uint32_t p = 0; p < K0; ++p)
            if (far[p] != INVALID32_ && far[p] != seeds[p]) {
                seeds.push_back(far[p]);
                any = true;
            }
        if (!any) break;
    }
    if (!on_gpu) assign();
    // last-resort guarantee of the size bound: chop oversized patches by BFS order
    {
        bool over = false;
        for (uint32_t p = 0; p < psize.size(); ++p)
            over |= psize[p] > patch_size;
        if (over) {
            std::vector<uint32_t> taken(psize.size(), 0), cur_id(psize.size());
            std::iota(cur_id.begin(), cur_id.end(), 0u);
            uint32_t next = (uint32_t)psize.size();
            // queue still holds the BFS order of the last assign()
            for (uint32_t i = 0; i < nf; ++i) {
                const uint32_t f = queue[i], p = face_patch[f];
                if (psize[p] <= patch_size) continue;
                if (taken[p] == patch_size) {
                    taken[p]  = 0;
                    cur_id[p] = next++;
                }
                ++taken[p];
                dist[f] = cur_id[p];  // reuse as new label
            }
            for (uint32_t f = 0; f < nf; ++f)
                if (psize[face_patch[f]] > patch_size) face_patch[f] = dist[f];
            seeds.resize(next);
        }
    }
    num_patches = (uint32_t)seeds.size();
    if (lloyd_runs) *lloyd_runs = (uint32_t)n_assign + (on_gpu ? 0u : (uint32_t)n_outer + 1u);  // every assign() pass
    if (getenv("RXM_VERBOSE"))
        fprintf(stderr, "[rxmesh_b200] lloyd: %d outer rounds, %d recentre+assign passes, %u patches (assign %.2fs, recentre %.2fs, %s)\n",
                n_outer, n_assign, num_patches, t_assign, t_recenter, on_gpu ? "gpu" : serial ? "serial" : "parallel");
}

// ---------------------------------------------------------------------------
// Locality ordering of the patch ids (the role of the reference's disabled Patcher::bfs, patcher/patcher.cu:583-638):
// Lloyd numbers patches by the position of their seed in the input face order, which says nothing about where the patch
// lies.  Patches are renumbered in breadth-first order over the patch adjacency graph, started from a pseudo-peripheral
// patch (the last patch a first sweep reaches), so that consecutive ids form layers that sweep across the mesh: a
// contiguous id range -- what a rank of the multi-GPU mode owns -- is then a slab with a short frontier, and neighbouring
// blocks of a launch touch neighbouring memory.  Deterministic: neighbours are visited in ascending id.
// ---------------------------------------------------------------------------
static void reorder_patches_bfs(const uint32_t* fe, uint32_t nf, uint32_t ne, std::vector<uint32_t>& fpatch, uint32_t P)
{
    // patch adjacency: an edge whose faces lie in different patches
    // first[e] = the smallest face on edge e; every other face of the edge is paired with it (all cores; the pair SET is what
    // counts: it is sorted and made unique below)
    U32Buf first((size_t)ne);
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < (int64_t)ne; ++e)
        first[e] = INVALID32_;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)(3ull * nf); ++i) {
        const uint32_t f   = (uint32_t)(i / 3);
        uint32_t       old = __atomic_load_n(&first[fe[i]], __ATOMIC_RELAXED);
        while (f < old && !__atomic_compare_exchange_n(&first[fe[i]], &old, f, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
        }
    }
    std::vector<uint64_t> pairs;
#pragma omp parallel
    {
        std::vector<uint64_t> mine;
#pragma omp for schedule(static) nowait
        for (int64_t i = 0; i < (int64_t)(3ull * nf); ++i) {
            const uint32_t f = (uint32_t)(i / 3), g = first[fe[i]];
            if (g == f) continue;
            const uint32_t a = fpatch[g], b = fpatch[f];
            if (a != b) mine.push_back(((uint64_t)a << 32) | b), mine.push_back(((uint64_t)b << 32) | a);
        }
        std::sort(mine.begin(), mine.end());
        mine.erase(std::unique(mine.begin(), mine.end()), mine.end());
#pragma omp critical
        pairs.insert(pairs.end(), mine.begin(), mine.end());
    }
    U32Buf().swap(first);
    std::sort(pairs.begin(), pairs.end());
    pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
    std::vector<uint32_t> off((size_t)P + 1, 0);
    for (uint64_t k : pairs)
        off[(k >> 32) + 1]++;
    for (uint32_t p = 0; p < P; ++p)
        off[p + 1] += off[p];
    auto bfs = [&](uint32_t start, std::vector<uint32_t>& order) {
        std::vector<uint8_t> seen(P, 0);
        order.clear();
        order.reserve(P);
        uint32_t scan = 0;
        for (uint32_t s = start;;) {
            seen[s] = 1;
            size_t head = order.size();
            order.push_back(s);
            while (head < order.size()) {
                const uint32_t p = order[head++];
                for (uint32_t i = off[p]; i < off[p + 1]; ++i) {
                    const uint32_t q = (uint32_t)(pairs[i] & 0xFFFFFFFFu);
                    if (!seen[q]) seen[q] = 1, order.push_back(q);
                }
            }
            while (scan < P && seen[scan]) ++scan;  // next component
            if (scan == P) break;
            s = scan;
        }
    };
    std::vector<uint32_t> order;
    bfs(0, order);
    // the last patch of the component of patch 0 (pseudo-peripheral): the first sweep's order ends with the other components
    // when there are any, so take the last patch REACHED from 0 by redoing the sweep's first phase
    uint32_t far = 0;
    {
        std::vector<uint8_t> seen(P, 0);
        std::vector<uint32_t> q(1, 0u);
        seen[0] = 1;
        for (size_t head = 0; head < q.size(); ++head)
            for (uint32_t i = off[q[head]]; i < off[q[head] + 1]; ++i) {
                const uint32_t t = (uint32_t)(pairs[i] & 0xFFFFFFFFu);
                if (!seen[t]) seen[t] = 1, q.push_back(t);
            }
        far = q.back();
    }
    bfs(far, order);
    std::vector<uint32_t> label(P);
    for (uint32_t i = 0; i < P; ++i)
        label[order[i]] = i;
#pragma omp parallel for schedule(static)
    for (int64_t f = 0; f < (int64_t)nf; ++f)
        fpatch[f] = label[fpatch[f]];
}

// face -> patch of the built-in patcher alone (edges + Lloyd + locality ordering), without building the patch store: what the
// multi-GPU mode needs to cut a mesh into shards before any shard is built
std::string compute_face_patch(const uint32_t* fv, uint32_t nf, const BuildOptions& opt, std::vector<uint32_t>& fpatch, uint32_t& P)
{
    if (opt.num_threads > 0) omp_set_num_threads(opt.num_threads);
    if (nf == 0) return "compute_face_patch: empty face list";
    uint32_t nv = 0;
    for (uint64_t i = 0; i < 3ull * nf; ++i)
        nv = std::max(nv, fv[i]);
    nv += 1;
    U32Buf ev, fe;
    const uint32_t ne = build_edges(fv, nf, nv, ev, fe);
    patcher_lloyd(fe.data(), nf, ne, opt.patch_size, opt.lloyd_iters, fpatch, P);
    if (opt.reorder_patches && P > 2) reorder_patches_bfs(fe.data(), nf, ne, fpatch, P);
    return "";
}

// ---------------------------------------------------------------------------
// build_mesh
// ---------------------------------------------------------------------------
std::string build_mesh(const uint32_t* fv, uint32_t nf, const uint32_t* face_patch_in,
                       const BuildOptions& opt, HostMesh& M)
{
    const double t_start = now_s();
    // RXM_VERBOSE: seconds per build phase
    const bool   verbose = getenv("RXM_VERBOSE") != nullptr;
    double       t_lap   = t_start;
    std::string  laps;
    auto lap = [&](const char* what) {
        if (!verbose) return;
        const double t = now_s();
        char         b[64];
        snprintf(b, sizeof(b), " %s %.2f", what, t - t_lap);
        laps += b;
        t_lap = t;
    };
    if (opt.num_threads > 0) omp_set_num_threads(opt.num_threads);
    if (nf == 0) return "build_mesh: empty face list";
    if (opt.patch_size == 0 || opt.patch_size > 16384) return "build_mesh: patch_size must be in [1, 16384]";
    M            = HostMesh();
    M.patch_size = opt.patch_size;

    // ---- sizes, degenerate-face check (rxmesh.cpp:590-597 requires triangles) ----
    uint32_t nv = 0;
    bool     degenerate = false;
#pragma omp parallel for reduction(max : nv) reduction(|| : degenerate) schedule(static)
    for (int64_t f = 0; f < (int64_t)nf; ++f) {
        const uint32_t a = fv[3 * f], b = fv[3 * f + 1], c = fv[3 * f + 2];
        nv         = std::max(nv, std::max(a, std::max(b, c)));
        degenerate = degenerate || (a == b || b == c || a == c);
    }
    if (degenerate) return "build_mesh: degenerate face (repeated vertex)";
    if (nv == INVALID32_) return "build_mesh: vertex id 0xFFFFFFFF is reserved";
    nv += 1;

    lap("check");
    // ---- global edges ----
    const uint32_t ne = build_edges(fv, nf, nv, M.ev, M.fe);
    M.num_elems[ELEM_V] = nv;
    M.num_elems[ELEM_E] = ne;
    M.num_elems[ELEM_F] = nf;
    const uint32_t* fe = M.fe.data();

    lap("edges");
    // ---- input statistics (rxmesh.cpp:560-650) ----
    std::vector<uint32_t> ef_cnt(ne, 0);
    for (uint64_t i = 0; i < 3ull * nf; ++i)
        ef_cnt[fe[i]]++;
    {
        std::vector<uint32_t> val(nv, 0);
        uint32_t max_ef = 0, max_val = 0;
        bool     closed = true, manifold = true;
        for (uint32_t e = 0; e < ne; ++e) {
            max_ef = std::max(max_ef, ef_cnt[e]);
            closed &= ef_cnt[e] >= 2;
            manifold &= ef_cnt[e] <= 2;
            max_val = std::max(max_val, ++val[M.ev[2ull * e]]);
            max_val = std::max(max_val, ++val[M.ev[2ull * e + 1]]);
        }
        uint32_t max_ff = 0;
#pragma omp parallel for reduction(max : max_ff) schedule(static)
        for (int64_t f = 0; f < (int64_t)nf; ++f)
            max_ff = std::max(max_ff, ef_cnt[fe[3 * f]] + ef_cnt[fe[3 * f + 1]] + ef_cnt[fe[3 * f + 2]] - 3);
        M.max_valence             = max_val;
        M.max_edge_incident_faces = max_ef;
        M.max_face_adjacent_faces = max_ff;
        M.is_closed               = closed;
        M.is_edge_manifold        = manifold;
    }
    std::vector<uint32_t>().swap(ef_cnt);

    lap("stats");
    // ---- face -> patch ----
    std::vector<uint32_t>& fpatch = M.elem_patch[ELEM_F];
    uint32_t               P      = 0;
    const double           t_p0   = now_s();
    if (face_patch_in) {
        // compact user labels, keeping their order
        uint32_t maxp = 0;
        for (uint32_t f = 0; f < nf; ++f) {
            if (face_patch_in[f] == INVALID32_) return "build_mesh: face_patch holds an invalid id";
            maxp = std::max(maxp, face_patch_in[f]);
        }
        std::vector<uint32_t> used((size_t)maxp + 2, 0);
        for (uint32_t f = 0; f < nf; ++f)
            used[face_patch_in[f]] = 1;
        exclusive_scan(used);
        P = used[(size_t)maxp + 1];
        fpatch.resize(nf);
#pragma omp parallel for schedule(static)
        for (int64_t f = 0; f < (int64_t)nf; ++f)
            fpatch[f] = used[face_patch_in[f]];
    } else {
        patcher_lloyd(fe, nf, ne, opt.patch_size, opt.lloyd_iters, fpatch, P, &M.lloyd_runs);
        if (opt.reorder_patches && P > 2) reorder_patches_bfs(fe, nf, ne, fpatch, P);
    }
    M.patcher_seconds = now_s() - t_p0;
    M.num_patches     = P;

    lap("patcher");
    // ---- vertex / edge owner = lowest patch id among the incident faces'
    //      patches (patcher/patcher.cu:730-756: first claim in patch order) ----
    std::vector<uint32_t>& vpatch = M.elem_patch[ELEM_V];
    std::vector<uint32_t>& epatch = M.elem_patch[ELEM_E];
    vpatch.assign(nv, INVALID32_);
    epatch.assign(ne, INVALID32_);
    auto atomic_min = [](uint32_t* addr, uint32_t v) {
        uint32_t old = __atomic_load_n(addr, __ATOMIC_RELAXED);
        while (v < old && !__atomic_compare_exchange_n(addr, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
        }
    };
#pragma omp parallel for schedule(static)
    for (int64_t f = 0; f < (int64_t)nf; ++f)
        for (int j = 0; j < 3; ++j) {
            atomic_min(&vpatch[fv[3 * f + j]], fpatch[f]);
            atomic_min(&epatch[fe[3 * f + j]], fpatch[f]);
        }
    for (uint32_t v = 0; v < nv; ++v)
        if (vpatch[v] == INVALID32_) return "build_mesh: isolated vertex id (not referenced by any face): " + std::to_string(v);

    lap("owners");
    // ---- faces of every patch (ascending id) and vertex -> faces CSR ----
    std::vector<uint32_t> pf_off((size_t)P + 1, 0), pf_val(nf);
    for (uint32_t f = 0; f < nf; ++f)
        pf_off[fpatch[f]]++;
    exclusive_scan(pf_off);
    {
        std::vector<uint32_t> cur(pf_off.begin(), pf_off.end() - 1);
        for (uint32_t f = 0; f < nf; ++f)
            pf_val[cur[fpatch[f]]++] = f;
    }
    for (uint32_t p = 0; p < P; ++p)
        if (pf_off[p + 1] - pf_off[p] > 16384)
            return "build_mesh: patch " + std::to_string(p) + " owns more than 16384 faces";
    std::vector<uint32_t> vf_off((size_t)nv + 1, 0), vf_val(3ull * nf);
    for (uint64_t i = 0; i < 3ull * nf; ++i)
        vf_off[fv[i]]++;
    exclusive_scan(vf_off);
    {
        std::vector<uint32_t> cur(vf_off.begin(), vf_off.end() - 1);
        for (uint32_t f = 0; f < nf; ++f)
            for (int j = 0; j < 3; ++j)
                vf_val[cur[fv[3ull * f + j]]++] = f;
    }

    lap("patch faces");
    // ---- phase A: per patch element lists (owned first, each half sorted by
    //      global id: rxmesh.cpp:845-869) and the neighbour-patch stash ----
    struct Tmp
    {
        std::vector<uint32_t> l[3];
        std::vector<uint32_t> stash;
        uint32_t              n_owned[3];
        // ring-2 extension: ext = global ids (ascending) of the vertices two rings out; r2_idx per not-owned vertex,
        // r2_off / r2_val = the stored rings as extended local ids
        std::vector<uint32_t> ext;
        std::vector<uint16_t> r2_idx, r2_off, r2_val;
    };
    std::vector<Tmp> tmp(P);
    std::string      err;
    const bool       ring2 = !opt.no_ring2;
#pragma omp parallel
    {
        std::vector<uint32_t> scratch;
#pragma omp for schedule(dynamic, 16)
        for (int64_t p = 0; p < (int64_t)P; ++p) {
            Tmp&            T  = tmp[p];
            const uint32_t* of = pf_val.data() + pf_off[p];
            const uint32_t  no = pf_off[p + 1] - pf_off[p];
            // ribbon = foreign faces sharing >= 1 vertex with an owned face
            // (patcher/patcher.cu:668-713)
            scratch.clear();
            for (uint32_t i = 0; i < no; ++i)
                for (int j = 0; j < 3; ++j) {
                    const uint32_t v = fv[3ull * of[i] + j];
                    for (uint32_t k = vf_off[v]; k < vf_off[v + 1]; ++k)
                        if (fpatch[vf_val[k]] != (uint32_t)p) scratch.push_back(vf_val[k]);
                }
            std::sort(scratch.begin(), scratch.end());
            scratch.erase(std::unique(scratch.begin(), scratch.end()), scratch.end());
            T.l[ELEM_F].assign(of, of + no);
            T.l[ELEM_F].insert(T.l[ELEM_F].end(), scratch.begin(), scratch.end());
            T.n_owned[ELEM_F] = no;
            // vertices and edges touched by the patch's faces
            for (int t = 0; t < 2; ++t) {
                const uint32_t*              src   = t == ELEM_V ? fv : fe;
                const std::vector<uint32_t>& owner = t == ELEM_V ? vpatch : epatch;
                scratch.clear();
                for (uint32_t f : T.l[ELEM_F])
                    for (int j = 0; j < 3; ++j)
                        scratch.push_back(src[3ull * f + j]);
                std::sort(scratch.begin(), scratch.end());
                scratch.erase(std::unique(scratch.begin(), scratch.end()), scratch.end());
                auto mid = std::stable_partition(scratch.begin(), scratch.end(),
                                                 [&](uint32_t g) { return owner[g] == (uint32_t)p; });
                T.n_owned[t] = (uint32_t)(mid - scratch.begin());
                T.l[t]       = scratch;
            }
            if (ring2) {
                // ---- ring extension: the COMPLETE one-ring of every not-owned vertex within `ring_depth` rings of an owned
                //      vertex.  Level-by-level over the global vertex -> faces map; vertices of those rings that the patch
                //      does not hold become "ext" vertices (extended local id n[V] + k, resolved through their owner) ----
                const auto&    LV  = T.l[ELEM_V];
                const uint32_t nov = T.n_owned[ELEM_V], nvp = (uint32_t)LV.size();
                auto lv = [&](uint32_t g) -> uint32_t {  // global vertex -> local id, INVALID32_ if not in the patch
                    auto b = vpatch[g] == (uint32_t)p ? LV.begin() : LV.begin() + nov;
                    auto e = vpatch[g] == (uint32_t)p ? LV.begin() + nov : LV.end();
                    auto it = std::lower_bound(b, e, g);
                    return (it != e && *it == g) ? (uint32_t)(it - LV.begin()) : INVALID32_;
                };
                auto ring_of = [&](uint32_t w, std::vector<uint32_t>& out) {  // sorted unique global ids of w's one-ring
                    out.clear();
                    for (uint32_t k = vf_off[w]; k < vf_off[w + 1]; ++k)
                        for (int j = 0; j < 3; ++j) {
                            const uint32_t g = fv[3ull * vf_val[k] + j];
                            if (g != w) out.push_back(g);
                        }
                    std::sort(out.begin(), out.end());
                    out.erase(std::unique(out.begin(), out.end()), out.end());
                };
                std::vector<uint32_t> ringed, frontier(LV.begin(), LV.begin() + nov), next_f, ring;  // global ids
                for (uint32_t depth = 0; depth < opt.ring_depth; ++depth) {
                    next_f.clear();
                    for (uint32_t g : frontier) {
                        ring_of(g, ring);
                        for (uint32_t u : ring)
                            if (vpatch[u] != (uint32_t)p) next_f.push_back(u);
                    }
                    std::sort(next_f.begin(), next_f.end());
                    next_f.erase(std::unique(next_f.begin(), next_f.end()), next_f.end());
                    // new = next_f \ ringed (both sorted)
                    std::vector<uint32_t> fresh;
                    std::set_difference(next_f.begin(), next_f.end(), ringed.begin(), ringed.end(), std::back_inserter(fresh));
                    std::vector<uint32_t> merged(ringed.size() + fresh.size());
                    std::merge(ringed.begin(), ringed.end(), fresh.begin(), fresh.end(), merged.begin());
                    ringed.swap(merged);
                    frontier.swap(fresh);
                }
                // rings of the ringed vertices; members the patch does not hold are ext vertices (ringed ext vertices too)
                std::vector<uint32_t> ring_g, ring_off(1, 0);
                T.ext.clear();
                for (uint32_t w : ringed) {
                    if (lv(w) == INVALID32_) T.ext.push_back(w);
                    ring_of(w, ring);
                    for (uint32_t g : ring) {
                        ring_g.push_back(g);
                        if (lv(g) == INVALID32_) T.ext.push_back(g);
                    }
                    ring_off.push_back((uint32_t)ring_g.size());
                }
                std::sort(T.ext.begin(), T.ext.end());
                T.ext.erase(std::unique(T.ext.begin(), T.ext.end()), T.ext.end());
                auto xid = [&](uint32_t g) -> uint32_t {  // extended local id
                    const uint32_t l = lv(g);
                    return l != INVALID32_ ? l : nvp + (uint32_t)(std::lower_bound(T.ext.begin(), T.ext.end(), g) - T.ext.begin());
                };
                if (nvp + T.ext.size() > 65535u || ring_g.size() > 65535u || ringed.size() >= 0xFFFFu) {
#pragma omp critical
                    err = "build_mesh: patch " + std::to_string(p) + " exceeds 65535 entries in its ring extension; use a smaller patch_size";
                    T.ext.clear(), T.r2_idx.clear();
                } else {
                    T.r2_idx.assign(nvp - nov + T.ext.size(), 0xFFFFu);
                    for (size_t r = 0; r < ringed.size(); ++r)
                        T.r2_idx[xid(ringed[r]) - nov] = (uint16_t)r;
                    T.r2_off.assign(ring_off.begin(), ring_off.end());
                    T.r2_val.resize(ring_g.size());
                    for (size_t i = 0; i < ring_g.size(); ++i)
                        T.r2_val[i] = (uint16_t)xid(ring_g[i]);
                }
            }
            // neighbour patches referenced by not-owned elements (and by the ext vertices)
            scratch.clear();
            for (uint32_t g : T.ext)
                scratch.push_back(vpatch[g]);
            for (int t = 0; t < 3; ++t)
                for (uint32_t i = T.n_owned[t]; i < T.l[t].size(); ++i)
                    scratch.push_back(M.elem_patch[t][T.l[t][i]]);
            std::sort(scratch.begin(), scratch.end());
            scratch.erase(std::unique(scratch.begin(), scratch.end()), scratch.end());
            T.stash = scratch;
        }
    }
    if (!err.empty()) return err;
    for (uint32_t p = 0; p < P; ++p)
        for (int t = 0; t < 3; ++t)
            if (tmp[p].l[t].size() > 65535u || tmp[p].stash.size() > 65535u)
                return "build_mesh: patch " + std::to_string(p) + " exceeds 65535 local elements; use a smaller patch_size";
    // the 16-bit encodings of the patch store: local face-edge entries are (edge << 1) | dir, the list offsets are u16
    // prefix sums over 2 nE and 3 nF entries
    for (uint32_t p = 0; p < P; ++p)
        if (tmp[p].l[ELEM_E].size() > 32767u || 3 * tmp[p].l[ELEM_F].size() > 65535u || 2 * tmp[p].l[ELEM_E].size() > 65535u)
            return "build_mesh: patch " + std::to_string(p) + " has " + std::to_string(tmp[p].l[ELEM_E].size()) + " edges / " +
                   std::to_string(tmp[p].l[ELEM_F].size()) + " faces with its ribbon: beyond the 16-bit local encodings (at most 32767 "
                   "edges and 21845 faces per patch); use a smaller patch_size";
    std::vector<uint32_t>().swap(vf_off);
    std::vector<uint32_t>().swap(vf_val);

    lap("lists");
    // ---- phase A2: one-ring fans of the owned vertices (local ids), if the input allows ----
    // For owned vertex v every incident face (v, a, b) (a cyclic rotation of its stored corner
    // order) is a directed link a -> b; the links must chain into ONE open or closed sequence.
    std::vector<std::vector<uint16_t>> fan_v(P), fan_off(P), fan_f(P), fan_e(P);
    bool fans_ok = !opt.no_fans && M.is_edge_manifold && M.max_valence < 4096;
    if (fans_ok) {
        int bad = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(| : bad)
        for (int64_t p = 0; p < (int64_t)P; ++p) {
            if (bad) continue;
            const Tmp&     T   = tmp[p];
            const uint32_t nov = T.n_owned[ELEM_V];
            auto lv = [&](uint32_t g) -> uint32_t {  // global vertex -> local id in this patch
                const auto& L = T.l[ELEM_V];
                if (vpatch[g] == (uint32_t)p) return (uint32_t)(std::lower_bound(L.begin(), L.begin() + nov, g) - L.begin());
                return (uint32_t)(std::lower_bound(L.begin() + nov, L.end(), g) - L.begin());
            };
            auto le = [&](uint32_t g) -> uint32_t {  // global edge -> local id in this patch
                const auto&    L   = T.l[ELEM_E];
                const uint32_t noe = T.n_owned[ELEM_E];
                if (epatch[g] == (uint32_t)p) return (uint32_t)(std::lower_bound(L.begin(), L.begin() + noe, g) - L.begin());
                return (uint32_t)(std::lower_bound(L.begin() + noe, L.end(), g) - L.begin());
            };
            // links grouped by owned vertex
            std::vector<uint32_t> cnt(nov + 1, 0);
            std::vector<std::array<uint16_t, 3>> lf(T.l[ELEM_F].size());
            for (size_t f = 0; f < T.l[ELEM_F].size(); ++f)
                for (int j = 0; j < 3; ++j) {
                    lf[f][j] = (uint16_t)lv(fv[3ull * T.l[ELEM_F][f] + j]);
                    if (lf[f][j] < nov) cnt[lf[f][j]]++;
                }
            std::vector<uint32_t> off(nov + 1, 0);
            for (uint32_t v = 0; v < nov; ++v)
                off[v + 1] = off[v] + cnt[v];
            // (first, second, face, edge to first, edge to second): corner j of a face sees its edge j towards corner j+1
            // and its edge j+2 coming back from corner j+2
            std::vector<std::array<uint16_t, 5>> links(off[nov]);
            std::vector<uint32_t>                cur(off.begin(), off.end() - 1);
            for (size_t f = 0; f < lf.size(); ++f)
                for (int j = 0; j < 3; ++j)
                    if (lf[f][j] < nov) {
                        const uint64_t gf = T.l[ELEM_F][f];
                        links[cur[lf[f][j]]++] = {lf[f][(j + 1) % 3], lf[f][(j + 2) % 3], (uint16_t)f,
                                                  (uint16_t)le(fe[3ull * gf + j]), (uint16_t)le(fe[3ull * gf + (j + 2) % 3])};
                    }
            auto& FO = fan_off[p];
            auto& FV = fan_v[p];
            auto& FF = fan_f[p];
            auto& FE = fan_e[p];
            FO.assign(nov + 1, 0);
            for (uint32_t v = 0; v < nov && !bad; ++v) {
                auto*          L = links.data() + off[v];
                const uint32_t k = off[v + 1] - off[v];
                // start: a link whose first vertex is nobody's second (open fan), else link 0
                uint32_t start = k, n_start = 0;
                for (uint32_t i = 0; i < k; ++i) {
                    bool has_pred = false;
                    for (uint32_t j = 0; j < k; ++j)
                        has_pred |= (L[j][1] == L[i][0]);
                    if (!has_pred) start = i, ++n_start;
                }
                const bool closed = (n_start == 0);
                if (n_start > 1 || k == 0) { bad = 1; break; }
                if (closed) {  // deterministic start: the link with the smallest first vertex
                    start = 0;
                    for (uint32_t i = 1; i < k; ++i)
                        if (L[i][0] < L[start][0]) start = i;
                }
                if (FV.size() + k + 1 > FAN_OFF_MASK) { bad = 1; break; }
                FO[v] = (uint16_t)(FV.size() | (closed ? FAN_CLOSED : 0));
                uint32_t curl = start, used = 0;
                FV.push_back(L[curl][0]);
                FF.push_back(L[curl][2]);
                FE.push_back(L[curl][3]);
                while (used < k) {
                    ++used;
                    const uint16_t nxt = L[curl][1];
                    if (used == k) {
                        if (closed) { if (nxt != L[start][0]) bad = 1; }
                        else { FV.push_back(nxt); FF.push_back(0xFFFFu); FE.push_back(L[curl][4]); }
                        break;
                    }
                    FV.push_back(nxt);
                    uint32_t found = k, nf_ = 0;
                    for (uint32_t j = 0; j < k; ++j)
                        if (L[j][0] == nxt) found = j, ++nf_;
                    if (nf_ != 1) { bad = 1; break; }
                    curl = found;
                    FF.push_back(L[curl][2]);
                    FE.push_back(L[curl][3]);
                }
            }
            FO[nov] = (uint16_t)FV.size();
        }
        fans_ok = !bad;
    }
    M.fans  = fans_ok;
    M.ring2 = ring2;

    lap("fans");
    // ---- prefixes: attribute slots (padded to 4), linear ids, ltog offsets, blob offsets ----
    M.desc.assign(P, PatchDesc());
    for (int t = 0; t < 3; ++t) {
        M.slot_base[t].assign((size_t)P + 1, 0);
        M.lin_base[t].assign((size_t)P + 1, 0);
        M.ltog_off[t].assign((size_t)P + 1, 0);
    }
    uint64_t topo_total = 0;
    for (uint32_t p = 0; p < P; ++p) {
        PatchDesc& D = M.desc[p];
        memset(&D, 0, sizeof(D));
        D.patch_id = p;
        for (int t = 0; t < 3; ++t) {
            D.n[t]              = (uint16_t)tmp[p].l[t].size();
            D.n_owned[t]        = (uint16_t)tmp[p].n_owned[t];
            D.slot_base[t]      = M.slot_base[t][p];
            D.lin_base[t]       = M.lin_base[t][p];
            M.slot_base[t][p + 1] = M.slot_base[t][p] + round_up(D.n_owned[t], 4);
            M.lin_base[t][p + 1]  = M.lin_base[t][p] + D.n_owned[t];
            M.ltog_off[t][p + 1]  = M.ltog_off[t][p] + D.n[t];
            M.max_per_patch[t]       = std::max<uint32_t>(M.max_per_patch[t], D.n[t]);
            M.max_owned_per_patch[t] = std::max<uint32_t>(M.max_owned_per_patch[t], D.n_owned[t]);
            M.max_not_owned[t]       = std::max<uint32_t>(M.max_not_owned[t], D.n[t] - D.n_owned[t]);
            M.total_local[t] += D.n[t];
        }
        D.n_stash    = (uint16_t)tmp[p].stash.size();
        D.flags      = (M.fans ? FLAG_FANS : 0) | (M.max_edge_incident_faces <= 2 ? FLAG_FF : 0) | (ring2 ? FLAG_RING2 : 0);
        if (M.fans) {  // all owned fans closed with six neighbours: the offsets are 6 v (patch_layout.h FLAG_UNIFORM6)
            const auto& FO  = fan_off[p];
            bool        u6  = D.n_owned[ELEM_V] > 0;
            for (uint32_t v = 0; v < D.n_owned[ELEM_V] && u6; ++v)
                u6 = FO[v] == (uint16_t)((6u * v) | FAN_CLOSED);
            u6 = u6 && FO[D.n_owned[ELEM_V]] == (uint16_t)(6u * D.n_owned[ELEM_V]);
            if (u6) D.flags |= FLAG_UNIFORM6;
        }
        if (ring2) {
            D.n_r2     = (uint16_t)(tmp[p].r2_off.empty() ? 0 : tmp[p].r2_off.size() - 1);
            D.n_ext    = (uint16_t)tmp[p].ext.size();
            D.r2_total = (uint32_t)tmp[p].r2_val.size();
            M.max_ext      = std::max<uint32_t>(M.max_ext, D.n_ext);
            M.max_r2       = std::max<uint32_t>(M.max_r2, D.n_r2);
            M.max_r2_total = std::max<uint32_t>(M.max_r2_total, D.r2_total);
        }
        D.fan_total  = M.fans ? (uint32_t)fan_v[p].size() : 0;
        M.max_fan_total = std::max<uint32_t>(M.max_fan_total, D.fan_total);
        M.max_stash  = std::max<uint32_t>(M.max_stash, D.n_stash);
        D.topo_off   = topo_total;
        D.compute_layout();
        topo_total += D.topo_bytes;
    }
    M.packed = !opt.force_wide && M.max_valence < PK_MAX_VRANK && M.max_edge_incident_faces <= PK_MAX_ERANK &&
               M.max_per_patch[0] <= PK_MAX_ELEMS && M.max_per_patch[1] <= PK_MAX_ELEMS &&
               M.max_per_patch[2] <= PK_MAX_ELEMS;
    if (M.packed)
        for (uint32_t p = 0; p < P; ++p)
            M.desc[p].flags |= FLAG_PACKED;
    for (int t = 0; t < 3; ++t) {
        M.num_slots[t] = M.slot_base[t][P];
        if (M.lin_base[t][P] != M.num_elems[t]) return "build_mesh: internal error, ownership does not partition the mesh";
        M.slot_to_global[t].resize(M.num_slots[t]);  // uninitialised: filled in parallel below
        M.global_to_slot[t].resize(M.num_elems[t]);
        uint32_t *s2g = M.slot_to_global[t].data(), *g2s = M.global_to_slot[t].data();
        const int64_t ns = M.num_slots[t], ng = M.num_elems[t];
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < std::max(ns, ng); ++i) {
            if (i < ns) s2g[i] = INVALID32_;
            if (i < ng) g2s[i] = INVALID32_;
        }
        M.ltog[t].resize(M.ltog_off[t][P]);
    }
    M.topo.resize(topo_total + 16);
    {
        const uint64_t nb = topo_total + 16, chunk = 1ull << 22;
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < (int64_t)((nb + chunk - 1) / chunk); ++c)
            memset(M.topo.data() + c * chunk, 0, (size_t)std::min<uint64_t>(chunk, nb - c * chunk));
    }

    lap("prefixes");
    // ---- phase B: id maps ----
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t p = 0; p < (int64_t)P; ++p)
        for (int t = 0; t < 3; ++t) {
            const auto& L = tmp[p].l[t];
            std::copy(L.begin(), L.end(), M.ltog[t].begin() + M.ltog_off[t][p]);
            for (uint32_t i = 0; i < tmp[p].n_owned[t]; ++i) {
                M.slot_to_global[t][M.slot_base[t][p] + i] = L[i];
                M.global_to_slot[t][L[i]]                  = M.slot_base[t][p] + i;
            }
        }

    lap("id maps");
    // ---- phase C: local topology (rxmesh.cpp:872-996), owner tables, stash ----
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t p = 0; p < (int64_t)P; ++p) {
        const Tmp&       T = tmp[p];
        const PatchDesc& D = M.desc[p];
        uint8_t*         B = M.topo.data() + D.topo_off;
        auto local_of = [&](int t, uint32_t g) -> uint32_t {
            if (M.elem_patch[t][g] == (uint32_t)p) return M.global_to_slot[t][g] - D.slot_base[t];
            auto b  = T.l[t].begin() + T.n_owned[t];
            auto it = std::lower_bound(b, T.l[t].end(), g);
            return (uint32_t)(it - T.l[t].begin());
        };
        uint16_t* lev = reinterpret_cast<uint16_t*>(B + D.off_ev());
        uint16_t* lfe = reinterpret_cast<uint16_t*>(B + D.off_fe());
        uint16_t* lfv = reinterpret_cast<uint16_t*>(B + D.off_fv());
        for (uint32_t e = 0; e < D.n[ELEM_E]; ++e) {
            const uint32_t g = T.l[ELEM_E][e];
            lev[2 * e]       = (uint16_t)local_of(ELEM_V, M.ev[2ull * g]);      // larger global id
            lev[2 * e + 1]   = (uint16_t)local_of(ELEM_V, M.ev[2ull * g + 1]);  // smaller global id
        }
        for (uint32_t f = 0; f < D.n[ELEM_F]; ++f) {
            const uint32_t g = T.l[ELEM_F][f];
            for (int j = 0; j < 3; ++j) {
                const uint32_t v0 = fv[3ull * g + j], v1 = fv[3ull * g + (j + 1) % 3];
                const uint32_t le = local_of(ELEM_E, fe[3ull * g + j]);
                const uint32_t dir = v0 > v1 ? 0u : 1u;  // 0 iff the corner vertex is the larger id
                lfe[3 * f + j]     = (uint16_t)((le << 1) | dir);
                lfv[3 * f + j]     = lev[2 * le + dir];  // == local id of v0
            }
        }
        // list offsets (always) and ranks (packed format only): rank = position of the
        // incidence inside its column's list, rows visited in ascending local id
        {
            uint16_t* voe = reinterpret_cast<uint16_t*>(B + D.off_voff_e());
            uint16_t* vof = reinterpret_cast<uint16_t*>(B + D.off_voff_f());
            uint16_t* eof = reinterpret_cast<uint16_t*>(B + D.off_eoff_f());
            const uint32_t nvp = D.n[ELEM_V], nep = D.n[ELEM_E], nfp = D.n[ELEM_F];
            auto annotate = [&](uint16_t* conn, uint32_t nnz, uint16_t* off, uint32_t ncols, uint32_t id_shift,
                                uint32_t rank_shift) {
                std::vector<uint16_t> cnt(ncols + 1, 0);
                for (uint32_t i = 0; i < nnz; ++i) {
                    const uint32_t c = (uint32_t)conn[i] >> id_shift;
                    const uint32_t r = cnt[c]++;
                    if (M.packed) conn[i] = (uint16_t)(conn[i] | (r << rank_shift));
                }
                uint32_t run = 0;
                for (uint32_t c = 0; c <= ncols; ++c) {
                    off[c] = (uint16_t)run;
                    run += c < ncols ? cnt[c] : 0;
                }
            };
            annotate(lev, 2 * nep, voe, nvp, 0, PK_ID_BITS);
            annotate(lfv, 3 * nfp, vof, nvp, 0, PK_ID_BITS);
            annotate(lfe, 3 * nfp, eof, nep, 1, PK_ID_BITS + 1);
        }
        if (D.flags & FLAG_FF) {
            // stored FF: the (at most two) faces of every local edge, then per owned face the other face across
            // edges 0, 1, 2, compacted to the front
            const uint32_t        nep = D.n[ELEM_E], em = M.packed ? PK_ID_MASK : 0x7FFFu;
            std::vector<uint16_t> ef2(2 * (size_t)nep, 0xFFFF);
            for (uint32_t f = 0; f < D.n[ELEM_F]; ++f)
                for (int j = 0; j < 3; ++j) {
                    const uint32_t e = ((uint32_t)lfe[3 * f + j] >> 1) & em;
                    ef2[2 * e + (ef2[2 * e] == 0xFFFF ? 0 : 1)] = (uint16_t)f;
                }
            uint16_t* ff = reinterpret_cast<uint16_t*>(B + D.off_ff());
            for (uint32_t f = 0; f < D.n_owned[ELEM_F]; ++f) {
                uint32_t k = 0;
                for (int j = 0; j < 3; ++j) {
                    const uint32_t e = ((uint32_t)lfe[3 * f + j] >> 1) & em;
                    const uint16_t o = ef2[2 * e] == f ? ef2[2 * e + 1] : ef2[2 * e];
                    if (o != 0xFFFF) ff[3 * f + k++] = o;
                }
                for (; k < 3; ++k)
                    ff[3 * f + k] = 0xFFFF;
            }
            // stored EF: the pairs of the owned edges as they are (faces were visited in ascending local id)
            memcpy(B + D.off_ef(), ef2.data(), 4 * (size_t)D.n_owned[ELEM_E]);
        }
        if (M.fans) {
            memcpy(B + D.off_fanoff(), fan_off[p].data(), fan_off[p].size() * 2);
            memcpy(B + D.off_fanv(), fan_v[p].data(), fan_v[p].size() * 2);
            memcpy(B + D.off_fanf(), fan_f[p].data(), fan_f[p].size() * 2);
            memcpy(B + D.off_fane(), fan_e[p].data(), fan_e[p].size() * 2);
        }
        if (D.flags & FLAG_RING2) {
            if (!T.r2_idx.empty()) memcpy(B + D.o_r2idx, T.r2_idx.data(), T.r2_idx.size() * 2);
            if (!T.r2_off.empty())
                memcpy(B + D.o_r2off, T.r2_off.data(), T.r2_off.size() * 2);
            if (!T.r2_val.empty()) memcpy(B + D.o_r2val, T.r2_val.data(), T.r2_val.size() * 2);
            uint32_t* eo = reinterpret_cast<uint32_t*>(B + D.o_ext);
            for (size_t k = 0; k < T.ext.size(); ++k) {
                const uint32_t g = T.ext[k], q = vpatch[g];
                const uint32_t s = (uint32_t)(std::lower_bound(T.stash.begin(), T.stash.end(), q) - T.stash.begin());
                eo[k]            = pack_owner(s, M.global_to_slot[ELEM_V][g] - M.slot_base[ELEM_V][q]);
            }
        }
        for (int t = 0; t < 3; ++t) {
            uint32_t* own = reinterpret_cast<uint32_t*>(B + D.off_own(t));
            for (uint32_t i = T.n_owned[t]; i < T.l[t].size(); ++i) {
                const uint32_t g = T.l[t][i];
                const uint32_t q = M.elem_patch[t][g];
                const uint32_t s = (uint32_t)(std::lower_bound(T.stash.begin(), T.stash.end(), q) - T.stash.begin());
                own[i - T.n_owned[t]] = pack_owner(s, M.global_to_slot[t][g] - M.slot_base[t][q]);
            }
        }
        StashEntry* st = reinterpret_cast<StashEntry*>(B + D.off_stash());
        for (uint32_t s = 0; s < T.stash.size(); ++s) {
            st[s].patch = T.stash[s];
            for (int t = 0; t < 3; ++t)
                st[s].slot_base[t] = M.slot_base[t][T.stash[s]];
        }
    }
    if (!opt.keep_ltog)
        for (int t = 0; t < 3; ++t) {
            U32Buf().swap(M.ltog[t]);
        }
    lap("topology");
    M.build_seconds = now_s() - t_start;
    if (verbose) fprintf(stderr, "[rxmesh_b200] build phases (s):%s\n", laps.c_str());
    if (opt.verbose)
        fprintf(stderr, "[rxmesh_b200] build: V=%u E=%u F=%u patches=%u (patcher %.2fs, total %.2fs)\n",
                nv, ne, nf, P, M.patcher_seconds, M.build_seconds);
    return "";
}

}  // namespace rxm
