// rxm_kernels.cu -- fixed-function sm_100a kernels of the static hot path.
//
//   k_query_store     the reference's query test kernel (tests/RXMesh_test/query_kernel.cuh:13-46):
//                     input(h) = h, output(h, i) = iter[i] as 64-bit owner handles
//   k_query_consume   roofline "consume" variant: out(s) = sum_i in(iter[i])
//   k_vertex_normals  apps/VertexNormal/vertex_normal_kernel.cuh:10-43 recast as owner-computes:
//                     no global atomics, no zero-fill, one coalesced store per patch
//   k_laplacian       apps/Smoothing/manual.h:86-104, the two kernels of an iteration fused
//   k_bilateral       apps/Filtering/filtering_rxmesh_kernel.cuh:426-548 on a materialised VV CSR
//   k_boundary        kernels/boundary.cuh:11-44
// Kernel families (chosen per mesh by the launchers at the bottom of this file):
//   *_fan2 / *_consume_fan<128>  one-ring fans: two vertices per thread with packed fp32x2 arithmetic (normals, Laplacian;
//                                k_laplacian_fan2<true> also pushes halo rows to peer GPUs), plain-read VV / VF consume
//   *_pk / <.., true>            rank-annotated ("packed") incidence: atomic-free transposes (no fans: non-manifold input)
//   <.., false>                  wide format: shared atomics + scan (patches beyond the packed limits)
//   k_persistent<Worker>         opt-in software-pipelined variants (RXM_PERSIST=1, rxm_persistent.cuh)
//
// One thread block per patch; every patch section and the patch's owned attribute
// slice arrive by TMA bulk copies under one mbarrier phase.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "rxm_kernels.h"
#include "rxm_persistent.cuh"
#include "rxmesh_b200/rxm_query.cuh"

namespace rxm {

static std::atomic<uint64_t> g_launches{0};  // launches may be queued from several host threads (rxm_multi.cu)
uint64_t        launch_counter()
{
    return g_launches;
}
void count_launches(uint64_t n)
{
    g_launches += n;
}

namespace {

using namespace dev;

constexpr int BT = 256;

// 1/x with one MUFU.RCP (max relative error 2^-23, i.e. within 1 ulp of the correctly rounded
// quotient the reference's n / (l_i + l_k) produces; tests hold the result to 1e-5 relative)
__device__ __forceinline__ float fast_rcp(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ PatchDesc load_desc(const PatchDesc* g)
{
    PatchDesc    d;
    const uint4* s = reinterpret_cast<const uint4*>(g);
    uint4*       t = reinterpret_cast<uint4*>(&d);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(PatchDesc) / 16); ++i)
        t[i] = __ldg(s + i);
    return d;
}

// --------------------------------------------------------------------------
// query -> store handles
// --------------------------------------------------------------------------
template <int OP, int KMAX, bool PACKED>
__global__ void __launch_bounds__(BT) k_query_store(MeshView mv, AttrView<uint64_t> in, AttrView<uint64_t> out)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t                      bar;
    __shared__ uint32_t                      warp_tmp[36];
    using Q = PatchQuery<OP, BT, KMAX, PACKED>;
    const uint32_t  p    = blockIdx.x;
    const PatchDesc d    = load_desc(mv.desc + p);
    const uint8_t*  blob = mv.topo + d.topo_off;
    Smem            sm(smem_raw);
    Q               q;
    q.plan(d, sm, true, false, mv.edge_manifold != 0);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(&bar, q.tx_bytes(d, true));
        q.issue(d, blob, &bar, true);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    const QueryResult r  = q.compute(d, warp_tmp, false, false);
    const OwnerTable  ot = q.owner_table(d);
    constexpr uint32_t S = OpTraits<OP>::src;
    const uint32_t sb_in = d.slot_base[S], cap_in = d.slot_cap(S), lb_in = d.lin_base[S];
    for (uint32_t s = threadIdx.x; s < r.n_src; s += BT) {
        in.data[in.index_known(sb_in, cap_in, lb_in, s, 0)] = ((uint64_t)d.patch_id << 32) | s;
        const uint32_t b = r.begin(s), n = min(r.size(s), out.nattr);
        for (uint32_t i = 0; i < n; ++i)
            out.data[out.index_known(sb_in, cap_in, lb_in, s, i)] = ot.handle(r.at(b + i));
    }
}

// --------------------------------------------------------------------------
// query -> consume (gather one fp32 per neighbour, write one fp32 per source)
// --------------------------------------------------------------------------
template <int OP, int KMAX, bool PACKED>
__global__ void __launch_bounds__(BT) k_query_consume(MeshView mv, const float* __restrict__ in, float* __restrict__ out)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t                      bar;
    __shared__ uint32_t                      warp_tmp[36];
    using Q = PatchQuery<OP, BT, KMAX, PACKED>;
    constexpr uint32_t S = OpTraits<OP>::src, D = OpTraits<OP>::dst;
    const uint32_t     p    = blockIdx.x;
    const PatchDesc    d    = load_desc(mv.desc + p);
    const uint8_t*     blob = mv.topo + d.topo_off;
    Smem               sm(smem_raw);
    Q                  q;
    q.plan(d, sm, true, false, mv.edge_manifold != 0);
    const uint32_t capD = d.slot_cap(D);
    float*         s_in = sm.alloc<float>(max((uint32_t)d.n[D], capD));
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(&bar, q.tx_bytes(d, true) + 4u * capD);
        q.issue(d, blob, &bar, true);
        if (capD) bulk_g2s(s_in, in + d.slot_base[D], 4u * capD, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    // ribbon values: gathered from their owners' slots
    const OwnerTable ot = q.owner_table(d);
    for (uint32_t i = d.n_owned[D] + threadIdx.x; i < d.n[D]; i += BT)
        s_in[i] = ldg_stream(in + ot.slot(i));
    const QueryResult r = q.compute(d, warp_tmp, false, true);  // compute() syncs before returning for CSR ops
    if (op_is_fixed<OP>()) __syncthreads();
    for (uint32_t s = threadIdx.x; s < r.n_src; s += BT) {
        const uint32_t b = r.begin(s), e = r.end(s);
        float          a = 0.f;
        for (uint32_t i = b; i < e; ++i)
            a += s_in[r.at(i)];
        out[d.slot_base[S] + s] = a;
    }
}

// --------------------------------------------------------------------------
// vertex normals (owner-computes)
// --------------------------------------------------------------------------
// UNIT = 0: Max-1999 weights (apps/VertexNormal); UNIT = 1: sum of unit face
// normals (apps/Filtering/filtering_rxmesh_kernel.cuh:15-46).
// ---- packed format: atomic-free, deterministic ----
// Phase 1, one thread per patch face: face normal n and the three corner weights;
// corner j's contribution n*w_j is STORED at position voff_f[v_j] + rank_j of a
// per-patch contribution list (the rank bits of the FV entry say where), so no
// two threads ever write the same address.  Phase 2, one thread per owned vertex:
// sum its contiguous segment in ascending face order.  Result: bit-reproducible.
template <int UNIT>
__global__ void __launch_bounds__(BT) k_vertex_normals_pk(MeshView mv, const float* __restrict__ x,
                                                          float* __restrict__ nrm)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t                      bar;
    const uint32_t  p    = blockIdx.x;
    const PatchDesc d    = load_desc(mv.desc + p);
    const uint8_t*  blob = mv.topo + d.topo_off;
    const uint32_t  nv = d.n[ELEM_V], nov = d.n_owned[ELEM_V], nf = d.n[ELEM_F];
    const uint32_t  cap = d.slot_cap(ELEM_V);
    const uint32_t  loff_bytes = round_up(2u * (nov + 1), 16);
    Smem            sm(smem_raw);
    uint16_t*       s_fv    = sm.alloc<uint16_t>(d.fe_bytes() / 2);
    uint16_t*       s_loff  = sm.alloc<uint16_t>(loff_bytes / 2);
    uint32_t*       s_own   = sm.alloc<uint32_t>(d.own_bytes(ELEM_V) / 4);
    StashEntry*     s_stash = sm.alloc<StashEntry>(d.n_stash);
    float*          s_xp    = sm.alloc<float>(3 * cap);  // owned coordinates as they arrive (packed xyz); reused for the result
    float4*         s_x     = sm.alloc<float4>(nv);      // all local vertices, one LDS.128 each
    float4*         s_c     = sm.alloc<float4>(3 * nf);  // contribution lists (only owned vertices' segments are used)
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(&bar, d.fe_bytes() + loff_bytes + d.own_bytes(ELEM_V) + d.stash_bytes() + 12u * cap);
        if (d.fe_bytes()) bulk_g2s(s_fv, blob + d.off_fv(), d.fe_bytes(), &bar);
        bulk_g2s(s_loff, blob + d.off_voff_f(), loff_bytes, &bar);
        if (d.own_bytes(ELEM_V)) bulk_g2s(s_own, blob + d.off_own(ELEM_V), d.own_bytes(ELEM_V), &bar);
        if (d.stash_bytes()) bulk_g2s(s_stash, blob + d.off_stash(), d.stash_bytes(), &bar);
        if (cap) bulk_g2s(s_xp, x + 3ull * d.slot_base[ELEM_V], 12u * cap, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    for (uint32_t i = threadIdx.x; i < nv; i += BT) {
        float4 q;
        if (i < nov) {
            q = make_float4(s_xp[3 * i], s_xp[3 * i + 1], s_xp[3 * i + 2], 0.f);
        } else {  // ribbon vertex: gather from the owner patch's slots
            const uint32_t o = s_own[i - nov];
            const float*   g = x + 3ull * ((uint64_t)s_stash[o >> 16].slot_base[ELEM_V] + (o & 0xFFFFu));
            q = make_float4(__ldg(g), __ldg(g + 1), __ldg(g + 2), 0.f);  // L1 keeps the row's sector between the three loads
        }
        s_x[i] = q;
    }
    __syncthreads();
    for (uint32_t f = threadIdx.x; f < nf; f += BT) {
        const uint32_t e0 = s_fv[3 * f], e1 = s_fv[3 * f + 1], e2 = s_fv[3 * f + 2];
        const uint32_t v0 = e0 & PK_ID_MASK, v1 = e1 & PK_ID_MASK, v2 = e2 & PK_ID_MASK;
        if (v0 >= nov && v1 >= nov && v2 >= nov) continue;  // touches no owned vertex
        const float4 p0 = s_x[v0], p1 = s_x[v1], p2 = s_x[v2];
        const float  ax = p1.x - p0.x, ay = p1.y - p0.y, az = p1.z - p0.z;
        const float  bx = p2.x - p0.x, by = p2.y - p0.y, bz = p2.z - p0.z;
        const float  nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
        float        w0, w1, w2;
        if (UNIT) {
            w0 = w1 = w2 = rsqrtf(nx * nx + ny * ny + nz * nz);
        } else {
            const float cx = p2.x - p1.x, cy = p2.y - p1.y, cz = p2.z - p1.z;
            const float l0 = ax * ax + ay * ay + az * az;  // |v0 v1|^2
            const float l1 = cx * cx + cy * cy + cz * cz;  // |v1 v2|^2
            const float l2 = bx * bx + by * by + bz * bz;  // |v2 v0|^2
            w0 = __frcp_rn(l0 + l2);
            w1 = __frcp_rn(l1 + l0);
            w2 = __frcp_rn(l2 + l1);
        }
        if (v0 < nov) s_c[s_loff[v0] + (e0 >> PK_ID_BITS)] = make_float4(nx * w0, ny * w0, nz * w0, 0.f);
        if (v1 < nov) s_c[s_loff[v1] + (e1 >> PK_ID_BITS)] = make_float4(nx * w1, ny * w1, nz * w1, 0.f);
        if (v2 < nov) s_c[s_loff[v2] + (e2 >> PK_ID_BITS)] = make_float4(nx * w2, ny * w2, nz * w2, 0.f);
    }
    __syncthreads();
    for (uint32_t v = threadIdx.x; v < cap; v += BT) {
        float sx = 0.f, sy = 0.f, sz = 0.f;
        if (v < nov) {
            const uint32_t b = s_loff[v], e = s_loff[v + 1];
            for (uint32_t i = b; i < e; ++i) {
                const float4 c = s_c[i];
                sx += c.x, sy += c.y, sz += c.z;
            }
        }
        s_xp[3 * v] = sx, s_xp[3 * v + 1] = sy, s_xp[3 * v + 2] = sz;
    }
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0 && cap) {
        bulk_s2g(nrm + 3ull * d.slot_base[ELEM_V], s_xp, 12u * cap);
        bulk_commit();
        bulk_wait_all_read();
    }
}

// ---- wide format: shared-memory float atomics (order not deterministic) ----
template <int UNIT>
__global__ void __launch_bounds__(BT) k_vertex_normals(MeshView mv, const float* __restrict__ x, float* __restrict__ nrm)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t                      bar;
    const uint32_t  p    = blockIdx.x;
    const PatchDesc d    = load_desc(mv.desc + p);
    const uint8_t*  blob = mv.topo + d.topo_off;
    const uint32_t  nv = d.n[ELEM_V], nov = d.n_owned[ELEM_V], nf = d.n[ELEM_F];
    const uint32_t  cap = d.slot_cap(ELEM_V);
    Smem            sm(smem_raw);
    uint16_t*       s_fv    = sm.alloc<uint16_t>(d.fe_bytes() / 2);
    uint32_t*       s_own   = sm.alloc<uint32_t>(d.own_bytes(ELEM_V) / 4);
    StashEntry*     s_stash = sm.alloc<StashEntry>(d.n_stash);
    float*          s_x     = sm.alloc<float>(3 * max(nv, cap));
    float*          s_n     = sm.alloc<float>(3 * cap);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(&bar, d.fe_bytes() + d.own_bytes(ELEM_V) + d.stash_bytes() + 12u * cap);
        if (d.fe_bytes()) bulk_g2s(s_fv, blob + d.off_fv(), d.fe_bytes(), &bar);
        if (d.own_bytes(ELEM_V)) bulk_g2s(s_own, blob + d.off_own(ELEM_V), d.own_bytes(ELEM_V), &bar);
        if (d.stash_bytes()) bulk_g2s(s_stash, blob + d.off_stash(), d.stash_bytes(), &bar);
        if (cap) bulk_g2s(s_x, x + 3ull * d.slot_base[ELEM_V], 12u * cap, &bar);
    }
    for (uint32_t i = threadIdx.x; i < 3 * cap; i += BT)
        s_n[i] = 0.f;
    __syncthreads();
    mbar_wait(&bar, 0);
    // ribbon vertices: gather their coordinates from the owner patches
    for (uint32_t i = threadIdx.x; i < nv - nov; i += BT) {
        const uint32_t o    = s_own[i];
        const uint64_t slot = (uint64_t)s_stash[o >> 16].slot_base[ELEM_V] + (o & 0xFFFFu);
        const float*   g    = x + 3ull * slot;
        s_x[3 * (nov + i) + 0] = __ldg(g + 0);
        s_x[3 * (nov + i) + 1] = __ldg(g + 1);
        s_x[3 * (nov + i) + 2] = __ldg(g + 2);
    }
    __syncthreads();
    for (uint32_t f = threadIdx.x; f < nf; f += BT) {
        const uint32_t v0 = s_fv[3 * f], v1 = s_fv[3 * f + 1], v2 = s_fv[3 * f + 2];
        if (v0 >= nov && v1 >= nov && v2 >= nov) continue;  // touches no owned vertex
        const float p0x = s_x[3 * v0], p0y = s_x[3 * v0 + 1], p0z = s_x[3 * v0 + 2];
        const float p1x = s_x[3 * v1], p1y = s_x[3 * v1 + 1], p1z = s_x[3 * v1 + 2];
        const float p2x = s_x[3 * v2], p2y = s_x[3 * v2 + 1], p2z = s_x[3 * v2 + 2];
        const float ax = p1x - p0x, ay = p1y - p0y, az = p1z - p0z;
        const float bx = p2x - p0x, by = p2y - p0y, bz = p2z - p0z;
        const float nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
        float       w0, w1, w2;
        if (UNIT) {
            w0 = w1 = w2 = rsqrtf(nx * nx + ny * ny + nz * nz);
        } else {
            const float cx = p2x - p1x, cy = p2y - p1y, cz = p2z - p1z;
            const float l0 = ax * ax + ay * ay + az * az;  // |v0 v1|^2
            const float l1 = cx * cx + cy * cy + cz * cz;  // |v1 v2|^2
            const float l2 = bx * bx + by * by + bz * bz;  // |v2 v0|^2
            w0 = __frcp_rn(l0 + l2);
            w1 = __frcp_rn(l1 + l0);
            w2 = __frcp_rn(l2 + l1);
        }
        if (v0 < nov) {
            atomicAdd(&s_n[3 * v0], nx * w0), atomicAdd(&s_n[3 * v0 + 1], ny * w0), atomicAdd(&s_n[3 * v0 + 2], nz * w0);
        }
        if (v1 < nov) {
            atomicAdd(&s_n[3 * v1], nx * w1), atomicAdd(&s_n[3 * v1 + 1], ny * w1), atomicAdd(&s_n[3 * v1 + 2], nz * w1);
        }
        if (v2 < nov) {
            atomicAdd(&s_n[3 * v2], nx * w2), atomicAdd(&s_n[3 * v2 + 1], ny * w2), atomicAdd(&s_n[3 * v2 + 2], nz * w2);
        }
    }
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0 && cap) {
        bulk_s2g(nrm + 3ull * d.slot_base[ELEM_V], s_n, 12u * cap);
        bulk_commit();
        bulk_wait_all_read();
    }
}

// --------------------------------------------------------------------------
// Laplacian smoothing step: x_out(v) = x(v) - lr * sum_u 2 (x(v) - x(u))
// --------------------------------------------------------------------------
template <int KMAX, bool PACKED>
__global__ void __launch_bounds__(BT) k_laplacian(MeshView mv, const float* __restrict__ x, float* __restrict__ xo, double lr)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t                      bar;
    __shared__ uint32_t                      warp_tmp[36];
    using Q = PatchQuery<OP_VV, BT, KMAX, PACKED>;
    const uint32_t  p    = blockIdx.x;
    const PatchDesc d    = load_desc(mv.desc + p);
    const uint8_t*  blob = mv.topo + d.topo_off;
    const uint32_t  nv = d.n[ELEM_V], nov = d.n_owned[ELEM_V];
    const uint32_t  cap = d.slot_cap(ELEM_V);
    Smem            sm(smem_raw);
    Q               q;
    q.plan(d, sm, true, false, mv.edge_manifold != 0);
    float* s_x = sm.alloc<float>(3 * max(nv, cap));
    float* s_o = sm.alloc<float>(3 * cap);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(&bar, q.tx_bytes(d, true) + 12u * cap);
        q.issue(d, blob, &bar, true);
        if (cap) bulk_g2s(s_x, x + 3ull * d.slot_base[ELEM_V], 12u * cap, &bar);
    }
    for (uint32_t i = 3 * nov + threadIdx.x; i < 3 * cap; i += BT)
        s_o[i] = 0.f;
    __syncthreads();
    mbar_wait(&bar, 0);
    const OwnerTable ot = q.owner_table(d);
    for (uint32_t i = nov + threadIdx.x; i < nv; i += BT) {
        const float* g = x + 3ull * ot.slot(i);
        s_x[3 * i + 0] = __ldg(g + 0);
        s_x[3 * i + 1] = __ldg(g + 1);
        s_x[3 * i + 2] = __ldg(g + 2);
    }
    const QueryResult r = q.compute(d, warp_tmp, false, true);
    for (uint32_t v = threadIdx.x; v < nov; v += BT) {
        const float    vx = s_x[3 * v], vy = s_x[3 * v + 1], vz = s_x[3 * v + 2];
        float          gx = 0.f, gy = 0.f, gz = 0.f;
        const uint32_t b = r.begin(v), n = r.size(v);
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t u = r.at(b + i);
            gx += 2.f * (vx - s_x[3 * u]);
            gy += 2.f * (vy - s_x[3 * u + 1]);
            gz += 2.f * (vz - s_x[3 * u + 2]);
        }
        // lr is a double in the reference (manual.h:37): the step is evaluated in fp64
        s_o[3 * v + 0] = (float)__dsub_rn((double)vx, __dmul_rn(lr, (double)gx));
        s_o[3 * v + 1] = (float)__dsub_rn((double)vy, __dmul_rn(lr, (double)gy));
        s_o[3 * v + 2] = (float)__dsub_rn((double)vz, __dmul_rn(lr, (double)gz));
    }
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0 && cap) {
        bulk_s2g(xo + 3ull * d.slot_base[ELEM_V], s_o, 12u * cap);
        bulk_commit();
        bulk_wait_all_read();
    }
}


// --------------------------------------------------------------------------
// one-ring-fan kernels: one thread per OWNED vertex, register accumulators
// --------------------------------------------------------------------------
// Loads the fan sections + owner table + the patch's coordinate slice, gathers the
// ribbon vertices' coordinates, and leaves every local vertex as one float4 in s_x.
struct FanPatch
{
    const uint16_t* s_fo;
    const uint16_t* s_fv;
    float*          s_xp;  // 3*cap floats: owned slice as it arrived; reused for the result
    float4*         s_x;
    uint32_t        nv, nov, cap;
    bool            u6;  // FLAG_UNIFORM6: every owned fan is closed with six neighbours, fan v starts at 6 v
};

__device__ __forceinline__ FanPatch fan_load(const MeshView& mv, const PatchDesc& d, const float* __restrict__ x,
                                             uint8_t* smem_raw, uint64_t* bar)
{
    const uint8_t* blob = mv.topo + d.topo_off;
    FanPatch       F;
    F.nv = d.n[ELEM_V], F.nov = d.n_owned[ELEM_V], F.cap = d.slot_cap(ELEM_V);
    Smem        sm(smem_raw);
    uint16_t*   s_fo    = sm.alloc<uint16_t>(d.fanoff_bytes() / 2);
    uint16_t*   s_fv    = sm.alloc<uint16_t>(d.fanv_bytes() / 2);
    uint32_t*   s_own   = sm.alloc<uint32_t>(d.own_bytes(ELEM_V) / 4);
    StashEntry* s_stash = sm.alloc<StashEntry>(d.n_stash);
    F.s_xp              = sm.alloc<float>(3 * F.cap);
    F.s_x               = sm.alloc<float4>(F.nv);
    F.s_fo = s_fo, F.s_fv = s_fv;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(bar, d.fanoff_bytes() + d.fanv_bytes() + d.own_bytes(ELEM_V) + d.stash_bytes() + 12u * F.cap);
        bulk_g2s(s_fo, blob + d.off_fanoff(), d.fanoff_bytes(), bar);
        if (d.fanv_bytes()) bulk_g2s(s_fv, blob + d.off_fanv(), d.fanv_bytes(), bar);
        if (d.own_bytes(ELEM_V)) bulk_g2s(s_own, blob + d.off_own(ELEM_V), d.own_bytes(ELEM_V), bar);
        if (d.stash_bytes()) bulk_g2s(s_stash, blob + d.off_stash(), d.stash_bytes(), bar);
        if (F.cap) bulk_g2s(F.s_xp, x + 3ull * d.slot_base[ELEM_V], 12u * F.cap, bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);
    for (uint32_t i = threadIdx.x; i < F.nv; i += BT) {
        float4 q;
        if (i < F.nov) {
            q = make_float4(F.s_xp[3 * i], F.s_xp[3 * i + 1], F.s_xp[3 * i + 2], 0.f);
        } else {
            const uint32_t o = s_own[i - F.nov];
            const float*   g = x + 3ull * ((uint64_t)s_stash[o >> 16].slot_base[ELEM_V] + (o & 0xFFFFu));
            q = make_float4(__ldg(g), __ldg(g + 1), __ldg(g + 2), 0.f);  // L1 keeps the row's sector between the three loads
        }
        F.s_x[i] = q;
    }
    __syncthreads();
    return F;
}

__device__ __forceinline__ void fan_store(const PatchDesc& d, const FanPatch& F, float* __restrict__ out)
{
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0 && F.cap) {
        bulk_s2g(out + 3ull * d.slot_base[ELEM_V], F.s_xp, 12u * F.cap);
        bulk_commit();
        bulk_wait_all_read();
    }
}

template <int UNIT>
__global__ void __launch_bounds__(BT, 8) k_vertex_normals_fan(MeshView mv, const float* __restrict__ x,
                                                           float* __restrict__ nrm)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t                      bar;
    const PatchDesc d = load_desc(mv.desc + blockIdx.x);
    const FanPatch  F = fan_load(mv, d, x, smem_raw, &bar);
    for (uint32_t v = threadIdx.x; v < F.cap; v += BT) {
        float sx = 0.f, sy = 0.f, sz = 0.f;
        if (v < F.nov) {
            const uint32_t o = F.s_fo[v], b = o & FAN_OFF_MASK, e = F.s_fo[v + 1] & FAN_OFF_MASK;
            const float4   p = F.s_x[v];
            // face (v, prev, cur): n = (prev - v) x (cur - v); corner weight 1 / (|prev-v|^2 + |cur-v|^2)
            auto face = [&](float px, float py, float pz, float pl, float cx, float cy, float cz, float cl) {
                const float nx = py * cz - pz * cy, ny = pz * cx - px * cz, nz = px * cy - py * cx;
                const float w  = UNIT ? rsqrtf(nx * nx + ny * ny + nz * nz) : fast_rcp(pl + cl);
                sx += nx * w, sy += ny * w, sz += nz * w;
            };
            if (o == (b | FAN_CLOSED) && e - b == 6) {
                // the regular case (closed fan of valence 6): straight-line code, all six neighbour loads in
                // flight together, no loop counter / branches
                float dx[6], dy[6], dz[6], dl[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    const float4 q = F.s_x[F.s_fv[b + k]];
                    dx[k] = q.x - p.x, dy[k] = q.y - p.y, dz[k] = q.z - p.z;
                    dl[k] = dx[k] * dx[k] + dy[k] * dy[k] + dz[k] * dz[k];
                }
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    const int j = (k + 1) % 6;
                    face(dx[k], dy[k], dz[k], dl[k], dx[j], dy[j], dz[j], dl[j]);
                }
            } else {
                float4      q  = F.s_x[F.s_fv[b]];
                const float d0x = q.x - p.x, d0y = q.y - p.y, d0z = q.z - p.z;
                const float l0 = d0x * d0x + d0y * d0y + d0z * d0z;
                float       px = d0x, py = d0y, pz = d0z, pl = l0;
                for (uint32_t i = b + 1; i < e; ++i) {
                    q = F.s_x[F.s_fv[i]];
                    const float cx = q.x - p.x, cy = q.y - p.y, cz = q.z - p.z;
                    const float cl = cx * cx + cy * cy + cz * cz;
                    face(px, py, pz, pl, cx, cy, cz, cl);
                    px = cx, py = cy, pz = cz, pl = cl;
                }
                if (o & FAN_CLOSED) face(px, py, pz, pl, d0x, d0y, d0z, l0);
            }
        }
        F.s_xp[3 * v] = sx, F.s_xp[3 * v + 1] = sy, F.s_xp[3 * v + 2] = sz;
    }
    fan_store(d, F, nrm);
}

// ---- two vertices per thread, packed fp32x2 arithmetic ----
// Evidence for the redesign (profiles/r01_ncu_full_tile32x16.csv, k_vertex_normals_fan; profiles/r01_vn_fan2_ncu.txt):
// the one-vertex-per-thread kernel sits at 74 % issue-active AND 78 % of the L1 data-pipe wavefront peak, DRAM at
// 57 %.  (1) Thread t owns vertices (t, t + BT2) and every subtraction / product / FMA of the pair is ONE FADD2 /
// FMUL2 / FFMA2 (dev::f2): 40 % fewer issued instructions.  (2) Shared-memory wavefronts are the remaining limiter,
// so nothing is staged twice: neighbours are read straight from the AoS slice the TMA delivered (3 LDS.32 at a
// 12-byte stride = 3 conflict-free wavefronts; the old float4 copy cost 4 per gather plus the conversion pass),
// ribbon vertices are gathered into the tail of the same array, fan ids are fetched two per LDS.32, and the result
// goes to its own buffer so no barrier separates compute from store.
#ifndef RXM_BT2
#define RXM_BT2 128
#endif
constexpr int BT2 = RXM_BT2;  // block size of the two-vertices-per-thread kernels.  Measured (gpurun r02u; 100 M-face grid VN /
                              // Laplacian ms, Lloyd-patched icosphere VN): 128 x 10 blocks 0.440 / 0.383 / 0.0574, 160 x 8 (two
                              // rounds instead of three for a 561-vertex tile) 0.461 / 0.402 / 0.0606, 192 x 6 0.519 / 0.473 /
                              // 0.0697, 96 x 13 0.473 / 0.438 / 0.0684: more, smaller blocks overlap their phases better
// resident blocks per SM the register allocation of the two-vertices-per-thread kernels aims at.  10 = 48 registers.
// 100 M-face grid, vertex normals / Laplacian ms per launch, Lloyd-patched 10 M icosphere normals (gpurun r02n):
//   8 blocks (64 registers) 0.463 / 0.391 / 0.0615    10 blocks (48) 0.441 / 0.383 / 0.0574
//   12 blocks (40 registers: spills, and 12 x 17.4 KB of shared memory leaves the 196 KB carve-out) 0.652 / 0.469 / 0.0846
#ifndef RXM_VN_MINB
#define RXM_VN_MINB 10
#endif

struct FanPatch2
{
    const uint16_t* s_fo;
    const uint16_t* s_fv;
    const float*    s_x;    // AoS xyz of every local vertex: [0, 3 nov) by TMA, [3 nov, 3 nv) gathered from the owners
    float*          s_out;  // 3*cap floats, bulk-stored to the patch's owned slice
    uint32_t        nv, nov, cap;
    bool            u6;  // FLAG_UNIFORM6: every owned fan is closed with six neighbours, fan v starts at 6 v
};

// COHERENT: the ribbon gathers may read ghost slots that a peer GPU wrote WHILE this kernel is running (fused halo
// exchange): they must not go through the non-coherent (.nc) path, which lies outside the PTX memory model and could
// serve a stale L1 sector; ld.global.cg is an ordinary (weak) load cached in L2 only, ordered by the acquire + barrier.
template <bool COHERENT = false, bool WITH_OUT = true>
__device__ __forceinline__ FanPatch2 fan_load2(const MeshView& mv, const PatchDesc& d, const float* __restrict__ x,
                                               uint8_t* smem_raw, uint64_t* bar)
{
    const uint8_t* blob = mv.topo + d.topo_off;
    FanPatch2      F;
    F.nv = d.n[ELEM_V], F.nov = d.n_owned[ELEM_V], F.cap = d.slot_cap(ELEM_V);
    Smem        sm(smem_raw);
    uint16_t*   s_fo    = sm.alloc<uint16_t>(d.fanoff_bytes() / 2);
    uint16_t*   s_fv    = sm.alloc<uint16_t>(d.fanv_bytes() / 2);
    uint32_t*   s_own   = sm.alloc<uint32_t>(d.own_bytes(ELEM_V) / 4);
    StashEntry* s_stash = sm.alloc<StashEntry>(d.n_stash);
    float*      s_x     = sm.alloc<float>(3 * max(F.nv, F.cap));
    F.s_out             = WITH_OUT ? sm.alloc<float>(3 * F.cap) : nullptr;
    F.s_fo = s_fo, F.s_fv = s_fv, F.s_x = s_x;
    const bool u6 = (d.flags & FLAG_UNIFORM6) != 0;
    F.u6          = u6;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        // FLAG_UNIFORM6: every owned fan is closed with six neighbours -- the offsets are written below, the section stays in HBM
        const uint32_t fo_bytes = u6 ? 0u : d.fanoff_bytes();
        mbar_arrive_expect_tx(bar, fo_bytes + d.fanv_bytes() + d.own_bytes(ELEM_V) + d.stash_bytes() + 12u * F.cap);
        if (fo_bytes) bulk_g2s(s_fo, blob + d.off_fanoff(), fo_bytes, bar);
        if (d.fanv_bytes()) bulk_g2s(s_fv, blob + d.off_fanv(), d.fanv_bytes(), bar);
        if (d.own_bytes(ELEM_V)) bulk_g2s(s_own, blob + d.off_own(ELEM_V), d.own_bytes(ELEM_V), bar);
        if (d.stash_bytes()) bulk_g2s(s_stash, blob + d.off_stash(), d.stash_bytes(), bar);
        if (F.cap) bulk_g2s(s_x, x + 3ull * d.slot_base[ELEM_V], 12u * F.cap, bar);
    }
    if (u6)
        for (uint32_t v = threadIdx.x; v <= F.nov; v += BT2)
            s_fo[v] = (uint16_t)(6u * v) | (v < F.nov ? FAN_CLOSED : (uint16_t)0);
    __syncthreads();
    mbar_wait(bar, 0);
    for (uint32_t i = F.nov + threadIdx.x; i < F.nv; i += BT2) {  // ribbon vertices: from their owners' slots
        const uint32_t o = s_own[i - F.nov];
        const float*   g = x + 3ull * ((uint64_t)s_stash[o >> 16].slot_base[ELEM_V] + (o & 0xFFFFu));
        // the three words of a row share a sector and neighbouring ribbon rows share lines: let L1 keep them between the loads
        const float    a = COHERENT ? __ldcg(g) : __ldg(g), b = COHERENT ? __ldcg(g + 1) : __ldg(g + 1),
                    c = COHERENT ? __ldcg(g + 2) : __ldg(g + 2);
        s_x[3 * i] = a, s_x[3 * i + 1] = b, s_x[3 * i + 2] = c;
    }
    __syncthreads();
    return F;
}

// one vertex, any valence, open or closed fan (scalar arithmetic)
template <int UNIT>
__device__ __forceinline__ void vn_one(const FanPatch2& F, uint32_t v, float& sx, float& sy, float& sz)
{
    const uint32_t o = F.s_fo[v], b = o & FAN_OFF_MASK, e = F.s_fo[v + 1] & FAN_OFF_MASK;
    const float    X = F.s_x[3 * v], Y = F.s_x[3 * v + 1], Z = F.s_x[3 * v + 2];
    sx = sy = sz = 0.f;
    if (b == e) return;
    auto face = [&](float px, float py, float pz, float pl, float cx, float cy, float cz, float cl) {
        const float nx = py * cz - pz * cy, ny = pz * cx - px * cz, nz = px * cy - py * cx;
        const float w  = UNIT ? rsqrtf(nx * nx + ny * ny + nz * nz) : fast_rcp(pl + cl);
        sx += nx * w, sy += ny * w, sz += nz * w;
    };
    const float* q   = F.s_x + 3u * F.s_fv[b];
    const float  d0x = q[0] - X, d0y = q[1] - Y, d0z = q[2] - Z;
    const float  l0  = d0x * d0x + d0y * d0y + d0z * d0z;
    float        px = d0x, py = d0y, pz = d0z, pl = l0;
    for (uint32_t i = b + 1; i < e; ++i) {
        q              = F.s_x + 3u * F.s_fv[i];
        const float cx = q[0] - X, cy = q[1] - Y, cz = q[2] - Z;
        const float cl = cx * cx + cy * cy + cz * cz;
        face(px, py, pz, pl, cx, cy, cz, cl);
        px = cx, py = cy, pz = cz, pl = cl;
    }
    if (o & FAN_CLOSED) face(px, py, pz, pl, d0x, d0y, d0z, l0);
}

// DIRECT: results go from registers straight to global memory (a warp's 32 rows are 384 contiguous bytes) instead of through
// a staging buffer + bulk store: 6 KB less shared memory per block
template <int UNIT, bool DIRECT>
__global__ void __launch_bounds__(BT2, RXM_VN_MINB) k_vertex_normals_fan2(MeshView mv, const float* __restrict__ x,
                                                              float* __restrict__ nrm)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t                      bar;
    const PatchDesc d = load_desc(mv.desc + blockIdx.x);
    const FanPatch2 F = fan_load2<false, !DIRECT>(mv, d, x, smem_raw, &bar);
    float* const    gout = nrm + 3ull * d.slot_base[ELEM_V];
    // thread t owns the pair (t, t + half), half = ceil(nov / 2): every owned vertex has a partner whatever the patch's size
    // (pairs (t, t + BT2) left the last nov mod 2 BT2 vertices of a patch on the scalar path: a third of a 384-vertex Lloyd
    // patch, profiles/r02i: 0.59-0.69 of the HBM peak on Lloyd patches against 0.85 on 561-vertex tiles)
    const uint32_t half = (F.nov + 1u) >> 1;
    for (uint32_t v = F.nov + threadIdx.x; v < F.cap; v += BT2) {  // slots past the owned vertices
        float* const o = DIRECT ? gout : F.s_out;
        o[3 * v] = o[3 * v + 1] = o[3 * v + 2] = 0.f;
    }
    for (uint32_t vA = threadIdx.x; vA < half; vA += BT2) {
        const uint32_t vB = vA + half;
        float          ax = 0.f, ay = 0.f, az = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
        bool           fast = false;
        uint32_t       bA = 0, bB = 0;
        if (vB < F.nov) {
            if (F.u6) {  // the whole patch is regular: no offsets to read, nothing to check
                bA = 6u * vA, bB = 6u * vB, fast = true;
            } else {
                const uint32_t oA = F.s_fo[vA], eA = F.s_fo[vA + 1] & FAN_OFF_MASK;
                const uint32_t oB = F.s_fo[vB], eB = F.s_fo[vB + 1] & FAN_OFF_MASK;
                bA = oA & FAN_OFF_MASK, bB = oB & FAN_OFF_MASK;
                // both fans closed with valence 6 (the regular case), ids readable as aligned 32-bit pairs
                fast = (oA & oB & FAN_CLOSED) && eA - bA == 6 && eB - bB == 6 && ((bA | bB) & 1u) == 0;
            }
        }
        if (fast) {
            // straight-line packed code: lane 0 of every f2 belongs to vertex A, lane 1 to vertex B
            const float *pa = F.s_x + 3u * vA, *pb = F.s_x + 3u * vB;
            const f2     PX = pk(pa[0], pb[0]), PY = pk(pa[1], pb[1]), PZ = pk(pa[2], pb[2]);
            const uint32_t* ia = reinterpret_cast<const uint32_t*>(F.s_fv + bA);
            const uint32_t* ib = reinterpret_cast<const uint32_t*>(F.s_fv + bB);
            f2              dx[6], dy[6], dz[6], dl[6];
#pragma unroll
            for (int k2 = 0; k2 < 3; ++k2) {
                const uint32_t wa = ia[k2], wb = ib[k2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int    k  = 2 * k2 + h;
                    const float* qa = F.s_x + 3u * (h ? wa >> 16 : wa & 0xFFFFu);
                    const float* qb = F.s_x + 3u * (h ? wb >> 16 : wb & 0xFFFFu);
                    dx[k] = sub2(pk(qa[0], qb[0]), PX);
                    dy[k] = sub2(pk(qa[1], qb[1]), PY);
                    dz[k] = sub2(pk(qa[2], qb[2]), PZ);
                    dl[k] = fma2(dz[k], dz[k], fma2(dy[k], dy[k], mul2(dx[k], dx[k])));
                }
            }
            f2 SX = pk(0.f, 0.f), SY = SX, SZ = SX;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const int j  = (k + 1) % 6;
                const f2  nx = fma2(dy[k], dz[j], neg2(mul2(dz[k], dy[j])));
                const f2  ny = fma2(dz[k], dx[j], neg2(mul2(dx[k], dz[j])));
                const f2  nz = fma2(dx[k], dy[j], neg2(mul2(dy[k], dx[j])));
                float     w0, w1;
                if (UNIT) {
                    upk(fma2(nz, nz, fma2(ny, ny, mul2(nx, nx))), w0, w1);
                    w0 = rsqrtf(w0), w1 = rsqrtf(w1);
                } else {
                    upk(add2(dl[k], dl[j]), w0, w1);
                    w0 = fast_rcp(w0), w1 = fast_rcp(w1);
                }
                const f2 W = pk(w0, w1);
                SX = fma2(nx, W, SX), SY = fma2(ny, W, SY), SZ = fma2(nz, W, SZ);
            }
            upk(SX, ax, bx), upk(SY, ay, by), upk(SZ, az, bz);
        } else {
            vn_one<UNIT>(F, vA, ax, ay, az);
            if (vB < F.nov) vn_one<UNIT>(F, vB, bx, by, bz);
        }
        float* const o = DIRECT ? gout : F.s_out;
        o[3 * vA] = ax, o[3 * vA + 1] = ay, o[3 * vA + 2] = az;
        if (vB < F.nov) o[3 * vB] = bx, o[3 * vB + 1] = by, o[3 * vB + 2] = bz;
    }
    if (DIRECT) return;
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0 && F.cap) {
        bulk_s2g(nrm + 3ull * d.slot_base[ELEM_V], F.s_out, 12u * F.cap);
        bulk_commit();
        bulk_wait_all_read();
    }
}

// Laplacian step (apps/Smoothing/manual.h:86-104) in the same two-vertices-per-thread form.  grad = sum 2 (x_v - x_u)
// is accumulated as 2 * sum (x_v - x_u): doubling is exact in binary floating point, so both give the same bits.
__device__ __forceinline__ void lap_finish(float px, float py, float pz, float gx, float gy, float gz, double lr, float* o)
{
    o[0] = (float)__dsub_rn((double)px, __dmul_rn(lr, (double)(2.f * gx)));
    o[1] = (float)__dsub_rn((double)py, __dmul_rn(lr, (double)(2.f * gy)));
    o[2] = (float)__dsub_rn((double)pz, __dmul_rn(lr, (double)(2.f * gz)));
}
__device__ __forceinline__ void lap_one(const FanPatch2& F, uint32_t v, double lr, float* out)
{
    const uint32_t b = F.s_fo[v] & FAN_OFF_MASK, e = F.s_fo[v + 1] & FAN_OFF_MASK;
    const float    X = F.s_x[3 * v], Y = F.s_x[3 * v + 1], Z = F.s_x[3 * v + 2];
    float          gx = 0.f, gy = 0.f, gz = 0.f;
    for (uint32_t i = b; i < e; ++i) {
        const float* q = F.s_x + 3u * F.s_fv[i];
        gx += X - q[0], gy += Y - q[1], gz += Z - q[2];
    }
    lap_finish(X, Y, Z, gx, gy, gz, lr, out + 3 * v);
}

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Flag protocol of the fused variant (two attribute buffers A / B, step s reads buf[s & 1] and writes buf[~s & 1], also
// into the neighbours' ghost slots of that buffer).  flag value s on a neighbour means "every row of my steps < s has
// landed in your ghost slots AND every block of mine that reads ghost slots has finished reading for steps < s".  The
// second half matters: the neighbour's step-s pushes overwrite the ghost slots my step s-1 blocks read, so the flag may
// only go up once ALL of them -- including patches that read ghost values but push nothing, e.g. a patch that touches the
// rank boundary only through vertices owned by lower-id patches -- are past their loads.  Every block that reads ghost
// slots or pushes rows therefore checks in at done_ctr (readers right after their gathers, pushers after their fenced
// stores) and the last one raises the flags.
// DIRECT: results go from registers straight to global memory (a warp's 32 rows are 384 contiguous bytes) instead of through
// a staging buffer + bulk store; 6 KB less shared memory per block keeps 8 resident blocks inside the 164 KB shared-memory
// carve-out: 100 M-face grid 0.459 -> 0.399 ms per step (0.80 -> 0.92 of the measured HBM peak).  The halo push of the fused
// variant then reads the few mirrored rows back from the block's own global writes (visible after the block barrier).
template <bool FUSED, bool DIRECT = false>
__global__ void __launch_bounds__(BT2, RXM_VN_MINB) k_laplacian_fan2(MeshView mv, const float* __restrict__ x, float* __restrict__ xo,
                                                         double lr, FusedHaloView fh)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t                      bar;
    // fused: blocks are rotated so that the patches with rows to push (the two ends of the rank's patch range) run
    // first and the neighbours' flags are raised early in the launch, not at its end
    const uint32_t  bidx = FUSED ? (blockIdx.x + fh.shift) % gridDim.x : blockIdx.x;
    const PatchDesc d    = load_desc(mv.desc + bidx);
    bool            syncs = false, pushes = false;
    if (FUSED) {
        // ghost slots hold what the neighbours pushed at the end of THEIR previous step: wait for their flags
        const uint32_t lp0 = fh.first + bidx;
        pushes = fh.push_off[lp0 + 1] > fh.push_off[lp0];
        syncs  = fh.reads_ghost[lp0] || pushes;
        if (syncs) {
            if (threadIdx.x < fh.npeers)
                while (ld_acquire_sys(fh.flags + threadIdx.x) < fh.step) {}
            __syncthreads();
        }
    }
    const FanPatch2 F = fan_load2<FUSED, !DIRECT>(mv, d, x, smem_raw, &bar);
    float* const    so = DIRECT ? xo + 3ull * d.slot_base[ELEM_V] : F.s_out;
    auto check_in = [&]() {  // thread 0 of a block whose ghost reads (and pushes) are complete
        if (atomicAdd(fh.done_ctr, 1u) == fh.n_sync_blocks - 1) {
            *fh.done_ctr = 0;
            __threadfence_system();
            for (uint32_t q = 0; q < fh.npeers; ++q)
                st_release_sys(fh.peer_flag[q], fh.step + 1);
        }
    };
    // fan_load2 ends with a barrier behind the ribbon gathers: a block that only READS ghost slots is done with them
    if (FUSED && syncs && !pushes && threadIdx.x == 0) check_in();
    const uint32_t half = (F.nov + 1u) >> 1;  // pairs (t, t + half): see k_vertex_normals_fan2
    for (uint32_t v = F.nov + threadIdx.x; v < F.cap; v += BT2)
        so[3 * v] = so[3 * v + 1] = so[3 * v + 2] = 0.f;
    for (uint32_t vA = threadIdx.x; vA < half; vA += BT2) {
        const uint32_t vB = vA + half;
        bool           fast = false;
        uint32_t       bA = 0, bB = 0;
        if (vB < F.nov) {
            if (F.u6) {
                bA = 6u * vA, bB = 6u * vB, fast = true;
            } else {
                const uint32_t eA = F.s_fo[vA + 1] & FAN_OFF_MASK, eB = F.s_fo[vB + 1] & FAN_OFF_MASK;
                bA = F.s_fo[vA] & FAN_OFF_MASK, bB = F.s_fo[vB] & FAN_OFF_MASK;
                fast = eA - bA == 6 && eB - bB == 6 && ((bA | bB) & 1u) == 0;
            }
        }
        if (fast) {
            const float *   pa = F.s_x + 3u * vA, *pb = F.s_x + 3u * vB;
            const f2        PX = pk(pa[0], pb[0]), PY = pk(pa[1], pb[1]), PZ = pk(pa[2], pb[2]);
            const uint32_t* ia = reinterpret_cast<const uint32_t*>(F.s_fv + bA);
            const uint32_t* ib = reinterpret_cast<const uint32_t*>(F.s_fv + bB);
            f2              GX = pk(0.f, 0.f), GY = GX, GZ = GX;
#pragma unroll
            for (int k2 = 0; k2 < 3; ++k2) {
                const uint32_t wa = ia[k2], wb = ib[k2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float* qa = F.s_x + 3u * (h ? wa >> 16 : wa & 0xFFFFu);
                    const float* qb = F.s_x + 3u * (h ? wb >> 16 : wb & 0xFFFFu);
                    GX = add2(GX, sub2(PX, pk(qa[0], qb[0])));
                    GY = add2(GY, sub2(PY, pk(qa[1], qb[1])));
                    GZ = add2(GZ, sub2(PZ, pk(qa[2], qb[2])));
                }
            }
            float gax, gbx, gay, gby, gaz, gbz;
            upk(GX, gax, gbx), upk(GY, gay, gby), upk(GZ, gaz, gbz);
            lap_finish(pa[0], pa[1], pa[2], gax, gay, gaz, lr, so + 3 * vA);
            lap_finish(pb[0], pb[1], pb[2], gbx, gby, gbz, lr, so + 3 * vB);
        } else {
            lap_one(F, vA, lr, so);
            if (vB < F.nov) lap_one(F, vB, lr, so);
        }
    }
    if (!DIRECT) {
        fence_proxy_async();
        __syncthreads();
        if (threadIdx.x == 0 && F.cap) {
            bulk_s2g(xo + 3ull * d.slot_base[ELEM_V], F.s_out, 12u * F.cap);
            bulk_commit();
            bulk_wait_all_read();
        }
    }
    if (FUSED && pushes) {
        if (DIRECT) __syncthreads();  // the block's own global writes are visible to all its threads
        // rows mirrored on other GPUs go straight into the neighbours' ghost slots (NVLink P2P stores); only the few
        // blocks that have such rows pay for the system-scope fence
        const uint32_t lp = fh.first + bidx;
        const uint32_t pb = fh.push_off[lp], pe = fh.push_off[lp + 1];
        for (uint32_t i = pb + threadIdx.x; i < pe; i += BT2) {
            const uint2  e   = fh.push[i];
            const float* src = so + 3u * (e.x & 0xFFFFu);
            float*       dst = fh.peer_out[e.x >> 16] + 3ull * e.y;
            if (DIRECT)
                dst[0] = __ldcg(src), dst[1] = __ldcg(src + 1), dst[2] = __ldcg(src + 2);
            else
                dst[0] = src[0], dst[1] = src[1], dst[2] = src[2];
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) check_in();
    }
}

__global__ void __launch_bounds__(BT) k_laplacian_fan(MeshView mv, const float* __restrict__ x, float* __restrict__ xo,
                                                      double lr)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t                      bar;
    const PatchDesc d = load_desc(mv.desc + blockIdx.x);
    const FanPatch  F = fan_load(mv, d, x, smem_raw, &bar);
    for (uint32_t v = threadIdx.x; v < F.cap; v += BT) {
        float ox = 0.f, oy = 0.f, oz = 0.f;
        if (v < F.nov) {
            const uint32_t b = F.s_fo[v] & FAN_OFF_MASK, e = F.s_fo[v + 1] & FAN_OFF_MASK;
            const float4   p = F.s_x[v];
            float          gx = 0.f, gy = 0.f, gz = 0.f;
            for (uint32_t i = b; i < e; ++i) {
                const float4 q = F.s_x[F.s_fv[i]];
                gx += 2.f * (p.x - q.x), gy += 2.f * (p.y - q.y), gz += 2.f * (p.z - q.z);
            }
            ox = (float)__dsub_rn((double)p.x, __dmul_rn(lr, (double)gx));
            oy = (float)__dsub_rn((double)p.y, __dmul_rn(lr, (double)gy));
            oz = (float)__dsub_rn((double)p.z, __dmul_rn(lr, (double)gz));
        }
        F.s_xp[3 * v] = ox, F.s_xp[3 * v + 1] = oy, F.s_xp[3 * v + 2] = oz;
    }
    fan_store(d, F, xo);
}

// sum of in[ids[b .. e)] in list order.  Six entries at an even offset (the regular vertex) are read as three LDS.32 and gathered
// without a loop: the scalar loop spent 60 % of the consume kernels' instructions on its own bookkeeping (profiles/r02q: 7
// instructions per neighbour, issue-active 65-72 % at 82-85 % of the DRAM peak); same additions in the same order.
__device__ __forceinline__ float gather_sum(const float* __restrict__ s_in, const uint16_t* __restrict__ ids, uint32_t b, uint32_t e)
{
    float a = 0.f;
    if (e - b == 6u && (b & 1u) == 0u) {
        const uint32_t* w  = reinterpret_cast<const uint32_t*>(ids + b);
        const uint32_t  w0 = w[0], w1 = w[1], w2 = w[2];
        a += s_in[w0 & 0xFFFFu];
        a += s_in[w0 >> 16];
        a += s_in[w1 & 0xFFFFu];
        a += s_in[w1 >> 16];
        a += s_in[w2 & 0xFFFFu];
        a += s_in[w2 >> 16];
        return a;
    }
    // any other list of up to 8 entries (mixed-valence meshes: most vertices): eight predicated steps instead of a loop whose
    // trip count differs from lane to lane -- no divergence, no loop bookkeeping; additions still in list order
    const uint32_t n = e - b;
    if (n <= 8u) {
#pragma unroll
        for (uint32_t k = 0; k < 8u; ++k)
            if (k < n) a += s_in[ids[b + k]];
        return a;
    }
    for (uint32_t i = b; i < e; ++i)
        a += s_in[ids[i]];
    return a;
}

// VF consume through the fan faces: out(v) = sum over the incident faces of in(f)
template <int BTC>
__global__ void __launch_bounds__(BTC, 2048 / BTC) k_vf_consume_fan(MeshView mv, const float* __restrict__ in, float* __restrict__ out)
{
    constexpr int BT = BTC;  // block size of the consume kernels (shadows the file-wide 256)
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t                      bar;
    const PatchDesc d    = load_desc(mv.desc + blockIdx.x);
    const uint8_t*  blob = mv.topo + d.topo_off;
    const uint32_t  nf = d.n[ELEM_F], nof = d.n_owned[ELEM_F], nov = d.n_owned[ELEM_V], cap = d.slot_cap(ELEM_F);
    Smem            sm(smem_raw);
    uint16_t*       s_fo    = sm.alloc<uint16_t>(d.fanoff_bytes() / 2);
    uint16_t*       s_ff    = sm.alloc<uint16_t>(d.fanf_bytes() / 2);
    uint32_t*       s_own   = sm.alloc<uint32_t>(d.own_bytes(ELEM_F) / 4);
    StashEntry*     s_stash = sm.alloc<StashEntry>(d.n_stash);
    float*          s_in    = sm.alloc<float>(max(nf, cap));
    const bool      u6      = (d.flags & FLAG_UNIFORM6) != 0;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
        const uint32_t fo_bytes = u6 ? 0u : d.fanoff_bytes();  // FLAG_UNIFORM6: offsets written below
        mbar_arrive_expect_tx(&bar, fo_bytes + d.fanf_bytes() + d.own_bytes(ELEM_F) + d.stash_bytes() + 4u * cap);
        if (fo_bytes) bulk_g2s(s_fo, blob + d.off_fanoff(), fo_bytes, &bar);
        if (d.fanf_bytes()) bulk_g2s(s_ff, blob + d.off_fanf(), d.fanf_bytes(), &bar);
        if (d.own_bytes(ELEM_F)) bulk_g2s(s_own, blob + d.off_own(ELEM_F), d.own_bytes(ELEM_F), &bar);
        if (d.stash_bytes()) bulk_g2s(s_stash, blob + d.off_stash(), d.stash_bytes(), &bar);
        if (cap) bulk_g2s(s_in, in + d.slot_base[ELEM_F], 4u * cap, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    for (uint32_t i = nof + threadIdx.x; i < nf; i += BT) {
        const uint32_t o = s_own[i - nof];
        s_in[i]          = ldg_stream(in + s_stash[o >> 16].slot_base[ELEM_F] + (o & 0xFFFFu));
    }
    __syncthreads();
    if (u6) {  // regular patch: fan v is entries [6 v, 6 v + 6)
        for (uint32_t v = threadIdx.x; v < nov; v += BT)
            out[d.slot_base[ELEM_V] + v] = gather_sum(s_in, s_ff, 6u * v, 6u * v + 6u);
        return;
    }
    for (uint32_t v = threadIdx.x; v < nov; v += BT) {
        const uint32_t o = s_fo[v], b = o & FAN_OFF_MASK;
        const uint32_t e = (s_fo[v + 1] & FAN_OFF_MASK) - ((o & FAN_CLOSED) ? 0u : 1u);  // open fan: last slot has no face
        out[d.slot_base[ELEM_V] + v] = gather_sum(s_in, s_ff, b, e);
    }
}

// VV consume through the fans: out(v) = sum over the one-ring of in(u)
template <int BTC>
__global__ void __launch_bounds__(BTC, 2048 / BTC) k_vv_consume_fan(MeshView mv, const float* __restrict__ in, float* __restrict__ out)
{
    constexpr int BT = BTC;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t                      bar;
    const PatchDesc d    = load_desc(mv.desc + blockIdx.x);
    const uint8_t*  blob = mv.topo + d.topo_off;
    const uint32_t  nv = d.n[ELEM_V], nov = d.n_owned[ELEM_V], cap = d.slot_cap(ELEM_V);
    Smem            sm(smem_raw);
    uint16_t*       s_fo    = sm.alloc<uint16_t>(d.fanoff_bytes() / 2);
    uint16_t*       s_fv    = sm.alloc<uint16_t>(d.fanv_bytes() / 2);
    uint32_t*       s_own   = sm.alloc<uint32_t>(d.own_bytes(ELEM_V) / 4);
    StashEntry*     s_stash = sm.alloc<StashEntry>(d.n_stash);
    float*          s_in    = sm.alloc<float>(max(nv, cap));
    const bool      u6      = (d.flags & FLAG_UNIFORM6) != 0;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
        const uint32_t fo_bytes = u6 ? 0u : d.fanoff_bytes();  // FLAG_UNIFORM6: offsets written below
        mbar_arrive_expect_tx(&bar, fo_bytes + d.fanv_bytes() + d.own_bytes(ELEM_V) + d.stash_bytes() + 4u * cap);
        if (fo_bytes) bulk_g2s(s_fo, blob + d.off_fanoff(), fo_bytes, &bar);
        if (d.fanv_bytes()) bulk_g2s(s_fv, blob + d.off_fanv(), d.fanv_bytes(), &bar);
        if (d.own_bytes(ELEM_V)) bulk_g2s(s_own, blob + d.off_own(ELEM_V), d.own_bytes(ELEM_V), &bar);
        if (d.stash_bytes()) bulk_g2s(s_stash, blob + d.off_stash(), d.stash_bytes(), &bar);
        if (cap) bulk_g2s(s_in, in + d.slot_base[ELEM_V], 4u * cap, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    for (uint32_t i = nov + threadIdx.x; i < nv; i += BT) {
        const uint32_t o = s_own[i - nov];
        s_in[i]          = ldg_stream(in + s_stash[o >> 16].slot_base[ELEM_V] + (o & 0xFFFFu));
    }
    __syncthreads();
    if (u6) {  // regular patch: fan v is entries [6 v, 6 v + 6)
        for (uint32_t v = threadIdx.x; v < nov; v += BT)
            out[d.slot_base[ELEM_V] + v] = gather_sum(s_in, s_fv, 6u * v, 6u * v + 6u);
        return;
    }
    for (uint32_t v = threadIdx.x; v < nov; v += BT) {
        const uint32_t b = s_fo[v] & FAN_OFF_MASK, e = s_fo[v + 1] & FAN_OFF_MASK;
        out[d.slot_base[ELEM_V] + v] = gather_sum(s_in, s_fv, b, e);
    }
}

// --------------------------------------------------------------------------
// persistent pipelined workers (rxm_persistent.cuh)
// --------------------------------------------------------------------------
struct FanLayout
{
    uint32_t stage_bytes, o_fo, o_fv, o_own, o_stash, o_xp, o_x4;
};

// MODE 0: Max-1999 vertex normals, 1: unit-face-normal sum, 2: Laplacian step
template <int MODE>
struct FanWorker
{
    struct Args
    {
        const float* x;
        float*       out;
        double       lr;
    };
    using Layout = FanLayout;

    static __device__ __forceinline__ void issue(const PatchDesc& d, const uint8_t* blob, const Args& a, const Layout& L,
                                                 uint8_t* st, uint64_t* bar)
    {
        const uint32_t cap = d.slot_cap(ELEM_V);
        mbar_arrive_expect_tx(bar, d.fanoff_bytes() + d.fanv_bytes() + d.own_bytes(ELEM_V) + d.stash_bytes() + 12u * cap);
        bulk_g2s(st + L.o_fo, blob + d.off_fanoff(), d.fanoff_bytes(), bar);
        if (d.fanv_bytes()) bulk_g2s(st + L.o_fv, blob + d.off_fanv(), d.fanv_bytes(), bar);
        if (d.own_bytes(ELEM_V)) bulk_g2s(st + L.o_own, blob + d.off_own(ELEM_V), d.own_bytes(ELEM_V), bar);
        if (d.stash_bytes()) bulk_g2s(st + L.o_stash, blob + d.off_stash(), d.stash_bytes(), bar);
        if (cap) bulk_g2s(st + L.o_xp, a.x + 3ull * d.slot_base[ELEM_V], 12u * cap, bar);
    }

    static __device__ __forceinline__ void pre(const PatchDesc& d, const Args& a, const Layout& L, uint8_t* st)
    {
        const uint32_t    nv = d.n[ELEM_V], nov = d.n_owned[ELEM_V];
        const uint32_t*   own   = reinterpret_cast<const uint32_t*>(st + L.o_own);
        const StashEntry* stash = reinterpret_cast<const StashEntry*>(st + L.o_stash);
        const float*      xp    = reinterpret_cast<const float*>(st + L.o_xp);
        float4*           x4    = reinterpret_cast<float4*>(st + L.o_x4);
        for (uint32_t i = threadIdx.x; i < nv; i += BT) {
            if (i < nov) {
                x4[i] = make_float4(xp[3 * i], xp[3 * i + 1], xp[3 * i + 2], 0.f);
            } else {  // ribbon vertex: asynchronous gather from the owner patch's slots
                const uint32_t o = own[i - nov];
                const float*   g = a.x + 3ull * ((uint64_t)stash[o >> 16].slot_base[ELEM_V] + (o & 0xFFFFu));
                float*         t = reinterpret_cast<float*>(&x4[i]);
                cp_async_4(t, g), cp_async_4(t + 1, g + 1), cp_async_4(t + 2, g + 2);
            }
        }
    }

    static __device__ __forceinline__ void compute(const PatchDesc& d, const Args& a, const Layout& L, uint8_t* st)
    {
        const uint32_t  nov = d.n_owned[ELEM_V];
        const uint16_t* fo  = reinterpret_cast<const uint16_t*>(st + L.o_fo);
        const uint16_t* fv  = reinterpret_cast<const uint16_t*>(st + L.o_fv);
        const float4*   x4  = reinterpret_cast<const float4*>(st + L.o_x4);
        float*          out = a.out + 3ull * d.slot_base[ELEM_V];
        for (uint32_t v = threadIdx.x; v < nov; v += BT) {
            const uint32_t o = fo[v], b = o & FAN_OFF_MASK, e = fo[v + 1] & FAN_OFF_MASK;
            const float4   p = x4[v];
            float          rx, ry, rz;
            if (MODE == 2) {
                float gx = 0.f, gy = 0.f, gz = 0.f;
                for (uint32_t i = b; i < e; ++i) {
                    const float4 q = x4[fv[i]];
                    gx += 2.f * (p.x - q.x), gy += 2.f * (p.y - q.y), gz += 2.f * (p.z - q.z);
                }
                rx = (float)__dsub_rn((double)p.x, __dmul_rn(a.lr, (double)gx));
                ry = (float)__dsub_rn((double)p.y, __dmul_rn(a.lr, (double)gy));
                rz = (float)__dsub_rn((double)p.z, __dmul_rn(a.lr, (double)gz));
            } else {
                float       sx = 0.f, sy = 0.f, sz = 0.f;
                float4      q  = x4[fv[b]];
                const float d0x = q.x - p.x, d0y = q.y - p.y, d0z = q.z - p.z;
                const float l0 = d0x * d0x + d0y * d0y + d0z * d0z;
                float       px = d0x, py = d0y, pz = d0z, pl = l0;
                auto face = [&](float cx, float cy, float cz, float cl) {
                    const float nx = py * cz - pz * cy, ny = pz * cx - px * cz, nz = px * cy - py * cx;
                    const float w  = MODE == 1 ? rsqrtf(nx * nx + ny * ny + nz * nz) : fast_rcp(pl + cl);
                    sx += nx * w, sy += ny * w, sz += nz * w;
                };
#pragma unroll 2
                for (uint32_t i = b + 1; i < e; ++i) {
                    q = x4[fv[i]];
                    const float cx = q.x - p.x, cy = q.y - p.y, cz = q.z - p.z;
                    const float cl = cx * cx + cy * cy + cz * cz;
                    face(cx, cy, cz, cl);
                    px = cx, py = cy, pz = cz, pl = cl;
                }
                if (o & FAN_CLOSED) face(d0x, d0y, d0z, l0);
                rx = sx, ry = sy, rz = sz;
            }
            out[3 * v] = rx, out[3 * v + 1] = ry, out[3 * v + 2] = rz;
        }
    }
};

struct ConsumeLayout
{
    uint32_t stage_bytes, o_a, o_b, o_own, o_stash, o_in, o_val;
};

// out(v) = sum of in(u) over the one-ring fan
struct VVFanConsumeWorker
{
    struct Args
    {
        const float* in;
        float*       out;
    };
    using Layout = ConsumeLayout;
    static __device__ __forceinline__ void issue(const PatchDesc& d, const uint8_t* blob, const Args& a, const Layout& L,
                                                 uint8_t* st, uint64_t* bar)
    {
        const uint32_t cap = d.slot_cap(ELEM_V);
        mbar_arrive_expect_tx(bar, d.fanoff_bytes() + d.fanv_bytes() + d.own_bytes(ELEM_V) + d.stash_bytes() + 4u * cap);
        bulk_g2s(st + L.o_a, blob + d.off_fanoff(), d.fanoff_bytes(), bar);
        if (d.fanv_bytes()) bulk_g2s(st + L.o_b, blob + d.off_fanv(), d.fanv_bytes(), bar);
        if (d.own_bytes(ELEM_V)) bulk_g2s(st + L.o_own, blob + d.off_own(ELEM_V), d.own_bytes(ELEM_V), bar);
        if (d.stash_bytes()) bulk_g2s(st + L.o_stash, blob + d.off_stash(), d.stash_bytes(), bar);
        if (cap) bulk_g2s(st + L.o_in, a.in + d.slot_base[ELEM_V], 4u * cap, bar);
    }
    static __device__ __forceinline__ void pre(const PatchDesc& d, const Args& a, const Layout& L, uint8_t* st)
    {
        const uint32_t    nv = d.n[ELEM_V], nov = d.n_owned[ELEM_V];
        const uint32_t*   own   = reinterpret_cast<const uint32_t*>(st + L.o_own);
        const StashEntry* stash = reinterpret_cast<const StashEntry*>(st + L.o_stash);
        float*            sin   = reinterpret_cast<float*>(st + L.o_in);
        for (uint32_t i = nov + threadIdx.x; i < nv; i += BT) {
            const uint32_t o = own[i - nov];
            cp_async_4(&sin[i], a.in + stash[o >> 16].slot_base[ELEM_V] + (o & 0xFFFFu));
        }
    }
    static __device__ __forceinline__ void compute(const PatchDesc& d, const Args& a, const Layout& L, uint8_t* st)
    {
        const uint32_t  nov = d.n_owned[ELEM_V];
        const uint16_t* fo  = reinterpret_cast<const uint16_t*>(st + L.o_a);
        const uint16_t* fv  = reinterpret_cast<const uint16_t*>(st + L.o_b);
        const float*    sin = reinterpret_cast<const float*>(st + L.o_in);
        for (uint32_t v = threadIdx.x; v < nov; v += BT) {
            const uint32_t b = fo[v] & FAN_OFF_MASK, e = fo[v + 1] & FAN_OFF_MASK;
            float          acc = 0.f;
            for (uint32_t i = b; i < e; ++i)
                acc += sin[fv[i]];
            a.out[d.slot_base[ELEM_V] + v] = acc;
        }
    }
};

// out(v) = sum of in(f) over VF(v); VF built in shared memory by the rank scatter (packed format)
struct VFConsumeWorker
{
    struct Args
    {
        const float* in;
        float*       out;
    };
    using Layout = ConsumeLayout;
    static __device__ __forceinline__ uint32_t loff_bytes(const PatchDesc& d) { return round_up(2u * (d.n_owned[ELEM_V] + 1u), 16); }
    static __device__ __forceinline__ void issue(const PatchDesc& d, const uint8_t* blob, const Args& a, const Layout& L,
                                                 uint8_t* st, uint64_t* bar)
    {
        const uint32_t cap = d.slot_cap(ELEM_F);
        mbar_arrive_expect_tx(bar, d.fe_bytes() + loff_bytes(d) + d.own_bytes(ELEM_F) + d.stash_bytes() + 4u * cap);
        if (d.fe_bytes()) bulk_g2s(st + L.o_a, blob + d.off_fv(), d.fe_bytes(), bar);
        bulk_g2s(st + L.o_b, blob + d.off_voff_f(), loff_bytes(d), bar);
        if (d.own_bytes(ELEM_F)) bulk_g2s(st + L.o_own, blob + d.off_own(ELEM_F), d.own_bytes(ELEM_F), bar);
        if (d.stash_bytes()) bulk_g2s(st + L.o_stash, blob + d.off_stash(), d.stash_bytes(), bar);
        if (cap) bulk_g2s(st + L.o_in, a.in + d.slot_base[ELEM_F], 4u * cap, bar);
    }
    static __device__ __forceinline__ void pre(const PatchDesc& d, const Args& a, const Layout& L, uint8_t* st)
    {
        const uint32_t    nf = d.n[ELEM_F], nof = d.n_owned[ELEM_F], nov = d.n_owned[ELEM_V];
        const uint16_t*   fv    = reinterpret_cast<const uint16_t*>(st + L.o_a);
        const uint16_t*   loff  = reinterpret_cast<const uint16_t*>(st + L.o_b);
        const uint32_t*   own   = reinterpret_cast<const uint32_t*>(st + L.o_own);
        const StashEntry* stash = reinterpret_cast<const StashEntry*>(st + L.o_stash);
        float*            sin   = reinterpret_cast<float*>(st + L.o_in);
        uint16_t*         val   = reinterpret_cast<uint16_t*>(st + L.o_val);
        for (uint32_t i = nof + threadIdx.x; i < nf; i += BT) {
            const uint32_t o = own[i - nof];
            cp_async_4(&sin[i], a.in + stash[o >> 16].slot_base[ELEM_F] + (o & 0xFFFFu));
        }
        for (uint32_t f = threadIdx.x; f < nf; f += BT) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const uint32_t e = fv[3 * f + j], c = e & PK_ID_MASK;
                if (c < nov) val[loff[c] + (e >> PK_ID_BITS)] = (uint16_t)f;
            }
        }
    }
    static __device__ __forceinline__ void compute(const PatchDesc& d, const Args& a, const Layout& L, uint8_t* st)
    {
        const uint32_t  nov  = d.n_owned[ELEM_V];
        const uint16_t* loff = reinterpret_cast<const uint16_t*>(st + L.o_b);
        const uint16_t* val  = reinterpret_cast<const uint16_t*>(st + L.o_val);
        const float*    sin  = reinterpret_cast<const float*>(st + L.o_in);
        for (uint32_t v = threadIdx.x; v < nov; v += BT) {
            const uint32_t b = loff[v], e = loff[v + 1];
            float          acc = 0.f;
            for (uint32_t i = b; i < e; ++i)
                acc += sin[val[i]];
            a.out[d.slot_base[ELEM_V] + v] = acc;
        }
    }
};

// --------------------------------------------------------------------------
// query -> materialised CSR in slot space (patch-grouped): off[slot], val = owner slots
// --------------------------------------------------------------------------
template <int OP, int KMAX, bool PACKED>
__global__ void __launch_bounds__(BT) k_query_csr(MeshView mv, const uint32_t* __restrict__ patch_nnz_off,
                                                  uint32_t* __restrict__ csr_off, uint32_t* __restrict__ csr_val)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t                      bar;
    __shared__ uint32_t                      warp_tmp[36];
    using Q = PatchQuery<OP, BT, KMAX, PACKED>;
    constexpr uint32_t S = OpTraits<OP>::src;
    const uint32_t  p    = blockIdx.x;
    const PatchDesc d    = load_desc(mv.desc + p);
    const uint8_t*  blob = mv.topo + d.topo_off;
    Smem            sm(smem_raw);
    Q               q;
    q.plan(d, sm, true, false);  // the CSR layout needs gap-free lists: FF stays on the generic (scan) path
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(&bar, q.tx_bytes(d, true));
        q.issue(d, blob, &bar, true);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    const QueryResult r    = q.compute(d, warp_tmp, false, true);
    const OwnerTable  ot   = q.owner_table(d);
    const uint32_t    base = patch_nnz_off[d.patch_id];
    const uint32_t    cap  = d.slot_cap(S), sb = d.slot_base[S];
    const uint32_t    tot  = r.n_src ? r.end(r.n_src - 1) : 0u;
    for (uint32_t s = threadIdx.x; s < cap; s += BT)
        csr_off[sb + s] = base + (s < r.n_src ? r.begin(s) : tot);  // padding slots: empty lists
    for (uint32_t s = threadIdx.x; s < r.n_src; s += BT) {
        const uint32_t b = r.begin(s), e = r.end(s);
        for (uint32_t i = b; i < e; ++i)
            csr_val[base + i] = ot.slot(r.at(i));
    }
}

// --------------------------------------------------------------------------
// bilateral mesh denoising on the materialised VV CSR (one thread per owned vertex slot)
// apps/Filtering/filtering_rxmesh_kernel.cuh:426-548 (+ 52-85, filtering_util.h:8-59)
// --------------------------------------------------------------------------
constexpr int BILATERAL_MAX_VV = 80;  // maxVVSize of the reference (filtering_rxmesh.cuh)

__global__ void __launch_bounds__(128) k_bilateral(const uint32_t* __restrict__ off, const uint32_t* __restrict__ val,
                                                   const float* __restrict__ x, const float* __restrict__ nrm,
                                                   float* __restrict__ xo, uint32_t num_slots, uint32_t* __restrict__ overflow)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= num_slots) return;
    const uint32_t b = off[s], e = off[s + 1];
    const float    px = x[3ull * s], py = x[3ull * s + 1], pz = x[3ull * s + 2];
    if (b == e) {  // padding slot
        xo[3ull * s] = px, xo[3ull * s + 1] = py, xo[3ull * s + 2] = pz;
        return;
    }
    float nx = nrm[3ull * s], ny = nrm[3ull * s + 1], nz = nrm[3ull * s + 2];
    {
        const float inv = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz);
        nx *= inv, ny *= inv, nz *= inv;
    }
    auto dist2 = [&](uint32_t u) {
        const float dx = x[3ull * u] - px, dy = x[3ull * u + 1] - py, dz = x[3ull * u + 2] - pz;
        return dx * dx + dy * dy + dz * dz;
    };
    float sc2 = 1e10f;
    for (uint32_t i = b; i < e; ++i)
        sc2 = fminf(sc2, dist2(val[i]));
    const float radius = 4.0f * sc2;
    uint32_t    vv[BILATERAL_MAX_VV];
    uint32_t    cnt = 1;
    vv[0]           = s;
    for (uint32_t head = 0; head < cnt; ++head) {
        const uint32_t w = vv[head];
        for (uint32_t i = off[w]; i < off[w + 1]; ++i) {
            const uint32_t u = val[i];
            if (u == s) continue;
            bool dup = false;
            for (uint32_t k = 0; k < cnt; ++k)
                dup |= (vv[k] == u);
            if (dup) continue;
            if (dist2(u) <= radius) {
                if (cnt < BILATERAL_MAX_VV)
                    vv[cnt++] = u;
                else
                    *overflow = 1u;  // the reference asserts here
            }
        }
    }
    float sum = 0.f, sum_sq = 0.f;
    for (uint32_t k = 0; k < cnt; ++k) {
        const uint32_t u = vv[k];
        float h = (x[3ull * u] - px) * nx + (x[3ull * u + 1] - py) * ny + (x[3ull * u + 2] - pz) * nz;
        h = fabsf(h);
        sum += h, sum_sq += h * h;
    }
    const float c   = (float)cnt;
    float       ss2 = sum_sq / c - (sum * sum) / (c * c);
    if (ss2 < 1.0e-20f) ss2 += 1.0e-20f;
    float num = 0.f, den = 0.f;
    for (uint32_t k = 0; k < cnt; ++k) {
        const uint32_t u  = vv[k];
        const float    qx = x[3ull * u] - px, qy = x[3ull * u + 1] - py, qz = x[3ull * u + 2] - pz;
        const float    t2 = qx * qx + qy * qy + qz * qz;
        const float    h  = qx * nx + qy * ny + qz * nz;
        const float    wc = expf(-0.5f * t2 / sc2), ws = expf(-0.5f * h * h / ss2);
        num += wc * ws * h, den += wc * ws;
    }
    const float k = num / den;
    xo[3ull * s] = px + nx * k, xo[3ull * s + 1] = py + ny * k, xo[3ull * s + 2] = pz + nz * k;
}

// --------------------------------------------------------------------------
// bilateral mesh denoising, PATCH-LOCAL (round 2; replaces the thread-per-slot walk over a global CSR above as the default)
// apps/Filtering/filtering_rxmesh_kernel.cuh:15-46 (unit-face vertex normals) + :426-548 (the filter) in ONE kernel.
//
// Evidence for the redesign (profiles/r01c_ncu_full.csv, k_bilateral): 82 % issue-active at 6 % DRAM, 3600 thread
// instructions per vertex, 48 % of them the O(cnt^2) duplicate scan over an 80-entry LOCAL-memory list, a separate normals
// kernel and a 28 B/vertex global CSR.  Here one block owns one patch:
//   * the fans, the ribbon / ext owner tables, the ring-2 extension and the patch's coordinate slice arrive by TMA bulk
//     copies; ribbon and ext coordinates are gathered once per patch; every k-ring walk then runs out of shared memory;
//   * the vertex normal (sum of unit face normals) falls out of the first pass over the vertex's own fan -- the filter needs
//     no other vertex's normal, so the normals kernel and its attribute round trip are gone;
//   * "seen" is a per-thread BITMAP over the patch's extended local ids (word-interleaved in shared memory: conflict-free),
//     so a duplicate visit costs a load and a test instead of a list scan, and rejected vertices are not measured twice;
//   * the accepted list lives in shared memory (no LDL); sigma_s' sums accumulate while vertices are accepted, and the two
//     exponentials of a weight are one: exp(-t^2 / 2 sigma_c^2) exp(-h^2 / 2 sigma_s^2) = exp(-(t^2/sigma_c^2 + h^2/sigma_s^2) / 2);
//   * rings: owned vertices -> their fan; ribbon vertices next to an owned vertex -> the ring-2 extension (complete ring in
//     extended local ids, patch_layout.h); a walk that has to expand any other vertex (outer ribbon, ext, or more than
//     BIL_FAST_CAP accepted) is DEFERRED: the block compacts those vertices and reruns them on the cross-patch path (slot
//     space, the materialised VV CSR, the reference's 80-entry cap), a few full warps instead of stragglers in every warp.
// Membership (dist^2 <= 4 sigma_c^2) is evaluated as fmaf(dz, dz, fmaf(dy, dy, dx * dx)) on fp32 differences in BOTH paths
// and in the oracle's fp32-membership mode, so the neighbourhoods are the same sets everywhere.
// --------------------------------------------------------------------------
constexpr int BIL_BT       = 128;  // block size of the cross-patch pass
constexpr int BIL_FAST_CAP = 16;   // accepted neighbours the fast path keeps per vertex before it defers

__device__ __forceinline__ float dist2f(float dx, float dy, float dz)
{
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

struct BilPatch
{
    const uint16_t *fo, *fv, *r2idx, *r2off, *r2val;
    const float4*   x;  // xyz of every extended local vertex, one LDS.128 each
    uint32_t        nx_, nov;  // nx_ = extended local vertices (patch + ext)
};

// pass over the vertex's own fan: unit vertex normal (normalised sum of unit face normals) and sigma_c^2 = min ring dist^2
__device__ __forceinline__ void bil_normal_sigma(const BilPatch& B, uint32_t v, float px, float py, float pz, float& nx,
                                                 float& ny, float& nz, float& sc2)
{
    const uint32_t o = B.fo[v], b = o & FAN_OFF_MASK, e = B.fo[v + 1] & FAN_OFF_MASK;
    nx = ny = nz = 0.f;
    float4      q   = B.x[B.fv[b]];
    const float d0x = q.x - px, d0y = q.y - py, d0z = q.z - pz;
    sc2             = dist2f(d0x, d0y, d0z);
    float ax = d0x, ay = d0y, az = d0z;
    auto  face = [&](float cx, float cy, float cz) {
        const float fx = ay * cz - az * cy, fy = az * cx - ax * cz, fz = ax * cy - ay * cx;
        const float w  = rsqrtf(fx * fx + fy * fy + fz * fz);
        nx += fx * w, ny += fy * w, nz += fz * w;
    };
    for (uint32_t i = b + 1; i < e; ++i) {
        q              = B.x[B.fv[i]];
        const float cx = q.x - px, cy = q.y - py, cz = q.z - pz;
        sc2            = fminf(sc2, dist2f(cx, cy, cz));
        face(cx, cy, cz);
        ax = cx, ay = cy, az = cz;
    }
    if (o & FAN_CLOSED) face(d0x, d0y, d0z);
    const float inv = rsqrtf(nx * nx + ny * ny + nz * nz);
    nx *= inv, ny *= inv, nz *= inv;
}

// fast path of one owned vertex; false = the walk left what the patch can answer (caller defers the vertex).
// ONE flattened loop over "visits" (the vertex's own ring, then the ring of every accepted vertex in turn): the lanes of a
// warp stay together until their TOTAL number of visits is used up, instead of diverging on every ring's length.
template <uint32_t BT>
__device__ __forceinline__ bool bil_fast(const BilPatch& B, uint32_t v, uint32_t* bm, uint32_t bm_words, uint16_t* lst, float* out)
{
    const uint32_t tid = threadIdx.x;
    const float4   P   = B.x[v];
    const float    px = P.x, py = P.y, pz = P.z;
    float          nx, ny, nz, sc2;
    bil_normal_sigma(B, v, px, py, pz, nx, ny, nz, sc2);
    const float radius = 4.0f * sc2;
    for (uint32_t w = 0; w < bm_words; ++w)
        bm[w * BT + tid] = 0u;
    bm[(v >> 5) * BT + tid] = 1u << (v & 31u);
    uint32_t        na = 0, head = 0;
    float           sum = 0.f, sum_sq = 0.f;
    const uint16_t* ring = B.fv;
    uint32_t        i = B.fo[v] & FAN_OFF_MASK, re = B.fo[v + 1] & FAN_OFF_MASK;
    while (true) {
        if (i == re) {  // next ring: the next accepted vertex
            if (head == na) break;
            const uint32_t w = lst[head * BT + tid];
            ++head;
            if (w < B.nov) {
                ring = B.fv, i = B.fo[w] & FAN_OFF_MASK, re = B.fo[w + 1] & FAN_OFF_MASK;
            } else {
                const uint32_t r = B.r2idx[w - B.nov];
                if (r == 0xFFFFu) return false;  // its ring is not in this patch
                ring = B.r2val, i = B.r2off[r], re = B.r2off[r + 1];
            }
            continue;
        }
        const uint32_t u    = ring[i++];
        uint32_t&      word = bm[(u >> 5) * BT + tid];
        const uint32_t bit = 1u << (u & 31u), old = word;
        if (old & bit) continue;  // seen before (accepted or rejected)
        word             = old | bit;
        const float4 q   = B.x[u];
        const float  cx = q.x - px, cy = q.y - py, cz = q.z - pz;
        if (dist2f(cx, cy, cz) > radius) continue;
        if (na == (uint32_t)BIL_FAST_CAP) return false;
        lst[na * BT + tid] = (uint16_t)u;
        ++na;
        const float h = fabsf(cx * nx + cy * ny + cz * nz);
        sum += h, sum_sq += h * h;
    }
    const float rc  = fast_rcp((float)(na + 1u));  // the vertex itself is the first member of its neighbourhood (h = 0, t = 0)
    const float m1  = sum * rc;
    float       ss2 = sum_sq * rc - m1 * m1;
    if (ss2 < 1.0e-20f) ss2 += 1.0e-20f;
    const float ic = -0.5f * fast_rcp(sc2), is = -0.5f * fast_rcp(ss2);
    float       num = 0.f, den = 1.f;
    for (uint32_t k = 0; k < na; ++k) {
        const float4 q  = B.x[lst[k * BT + tid]];
        const float  cx = q.x - px, cy = q.y - py, cz = q.z - pz;
        const float  t2 = dist2f(cx, cy, cz), h = cx * nx + cy * ny + cz * nz;
        const float  w  = __expf(t2 * ic + h * h * is);
        num += w * h, den += w;
    }
    const float kk = num * fast_rcp(den);
    out[0] = px + nx * kk, out[1] = py + ny * kk, out[2] = pz + nz * kk;
    return true;
}

// Tried and measured against this version on the 10 M-face torus (profiles/r02_bilateral_bt.txt, r02k / r02l): no "seen" bitmap
// (distance first, accepted-list scan for candidates inside the radius; 52 KB per block, 3-4 resident blocks) 0.454-0.52 ms;
// the same with ring 1 taken from the distances of the normal pass and the other rings measured two candidates per FFMA2
// 0.48 ms (ncu: 496 M warp instructions at 14.8 active lanes against 328 M at 22.0 here).  A vertex accepts ~10 neighbours and
// makes ~66 visits, 70 % of them duplicates: the O(1) bitmap probe beats any list scan, and the walk is bound by the number
// of visits, not by residency.
// A vertex the patch could not finish: its slot, unit normal and sigma_c^2 (already computed from its fan)
struct BilDeferred
{
    uint32_t slot;
    float    nx, ny, nz, sc2;
};

// cross-patch path: the DEFERRED vertices of all patches, compacted into one work list by k_bilateral_patch, one thread
// each with every warp full (inside the patch kernel they were a handful of stragglers per block whose dependent global
// loads -- CSR row, then coordinates -- nothing could hide).  Slot space, VV CSR, accepted list of interleaved u32 slots in
// shared memory, duplicate check by list scan, the reference's cap of 80 (filtering_rxmesh.cuh: maxVVSize).
__global__ void __launch_bounds__(BIL_BT) k_bilateral_deferred(const uint32_t* __restrict__ count, uint32_t* __restrict__ next_count,
                                                               const BilDeferred* __restrict__ work,
                                                               const uint32_t* __restrict__ off, const uint32_t* __restrict__ val,
                                                               const float* __restrict__ xg, float* __restrict__ xo,
                                                               uint32_t* __restrict__ flags)
{
    constexpr uint32_t R = BIL_BT;
    __shared__ uint32_t lst_all[BILATERAL_MAX_VV * BIL_BT];
    const uint32_t      n = *count, j = threadIdx.x;
    if (blockIdx.x == 0 && j == 0) {
        flags[1] += n;    // statistics: deferred vertices of the call
        *next_count = 0;  // the counter the NEXT iteration's patch kernel fills (the two alternate)
    }
    uint32_t*           lst = lst_all;
    for (uint32_t i = blockIdx.x * BIL_BT + j; i < n; i += gridDim.x * BIL_BT) {
        const BilDeferred w  = work[i];
        const float       px = xg[3ull * w.slot], py = xg[3ull * w.slot + 1], pz = xg[3ull * w.slot + 2];
        const float       nx = w.nx, ny = w.ny, nz = w.nz, sc2 = w.sc2, radius = 4.0f * sc2;
        uint32_t          cnt = 1;
        lst[j]                = w.slot;
        float sum = 0.f, sum_sq = 0.f;
        for (uint32_t head = 0; head < cnt; ++head) {
            const uint32_t c = lst[head * R + j], rb = off[c], re = off[c + 1];
            for (uint32_t k = rb; k < re; ++k) {
                const uint32_t u   = val[k];
                bool           dup = false;
                for (uint32_t q = 0; q < cnt; ++q)
                    dup |= (lst[q * R + j] == u);
                if (dup) continue;
                const float cx = xg[3ull * u] - px, cy = xg[3ull * u + 1] - py, cz = xg[3ull * u + 2] - pz;
                if (dist2f(cx, cy, cz) > radius) continue;
                if (cnt < (uint32_t)BILATERAL_MAX_VV) {
                    lst[cnt * R + j] = u;
                    ++cnt;
                    const float h = fabsf(cx * nx + cy * ny + cz * nz);
                    sum += h, sum_sq += h * h;
                } else
                    flags[0] = 1u;  // the reference asserts here
            }
        }
        const float c   = (float)cnt;
        float       ss2 = sum_sq / c - (sum * sum) / (c * c);
        if (ss2 < 1.0e-20f) ss2 += 1.0e-20f;
        const float ic = -0.5f / sc2, is = -0.5f / ss2;
        float       num = 0.f, den = 1.f;
        for (uint32_t q = 1; q < cnt; ++q) {
            const uint32_t u  = lst[q * R + j];
            const float    cx = xg[3ull * u] - px, cy = xg[3ull * u + 1] - py, cz = xg[3ull * u + 2] - pz;
            const float    t2 = dist2f(cx, cy, cz), h = cx * nx + cy * ny + cz * nz;
            const float    wgt = __expf(t2 * ic + h * h * is);
            num += wgt * h, den += wgt;
        }
        const float kk = num / den;
        xo[3ull * w.slot] = px + nx * kk, xo[3ull * w.slot + 1] = py + ny * kk, xo[3ull * w.slot + 2] = pz + nz * kk;
    }
}

// Block size = the patch's owned vertices split into equal rounds (chosen by the launcher): a fixed 512 left 14 of 16 warps
// waiting at the barrier while 2 finished the tail of a 561-vertex patch (profiles/r02e: barrier stall 1.8 per issue).
// (a compile-time block size: with blockDim.x as the stride of the interleaved bitmaps / lists the kernel was 25 % slower)
template <uint32_t BT>
__global__ void __launch_bounds__(BT) k_bilateral_patch(MeshView mv, const float* __restrict__ x, float* __restrict__ xo,
                                                        uint32_t bm_words, uint32_t* __restrict__ work_count,
                                                        BilDeferred* __restrict__ work)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t                      bar;
    __shared__ uint32_t                      s_ndef, s_wbase;
    const PatchDesc d    = load_desc(mv.desc + blockIdx.x);
    const uint8_t*  blob = mv.topo + d.topo_off;
    const bool      r2   = (d.flags & FLAG_RING2) != 0;
    const uint32_t  nv = d.n[ELEM_V], nov = d.n_owned[ELEM_V], cap = d.slot_cap(ELEM_V), next = r2 ? d.n_ext : 0u;
    Smem            sm(smem_raw);
    uint16_t*       s_fo    = sm.alloc<uint16_t>(d.fanoff_bytes() / 2);
    uint16_t*       s_fv    = sm.alloc<uint16_t>(d.fanv_bytes() / 2);
    uint32_t*       s_own   = sm.alloc<uint32_t>(d.own_bytes(ELEM_V) / 4);
    StashEntry*     s_stash = sm.alloc<StashEntry>(d.n_stash);
    uint16_t*       s_r2i   = sm.alloc<uint16_t>(r2 ? d.r2idx_bytes() / 2 : nv - nov + 8u);
    uint16_t*       s_r2o   = sm.alloc<uint16_t>(r2 ? d.r2off_bytes() / 2 : 8);
    uint16_t*       s_r2v   = sm.alloc<uint16_t>(r2 ? d.r2val_bytes() / 2 : 8);
    uint32_t*       s_ext   = sm.alloc<uint32_t>(r2 ? d.ext_bytes() / 4 : 4);
    float4*         s_x     = sm.alloc<float4>(nv + next);
    float*          s_out   = sm.alloc<float>(3 * cap);  // the owned slice as it arrives (AoS), later the results
    uint16_t*       s_def   = sm.alloc<uint16_t>(nov + 8u);
    uint32_t*       s_priv  = sm.alloc<uint32_t>((bm_words + BIL_FAST_CAP / 2) * BT);  // per-thread bitmaps, then the u16 lists
    if (threadIdx.x == 0) {
        s_ndef = 0;
        mbar_init(&bar, 1);
        fence_mbar_init();
        const uint32_t r2b = r2 ? d.r2idx_bytes() + d.r2off_bytes() + d.r2val_bytes() + d.ext_bytes() : 0u;
        mbar_arrive_expect_tx(&bar, d.fanoff_bytes() + d.fanv_bytes() + d.own_bytes(ELEM_V) + d.stash_bytes() + r2b + 12u * cap);
        bulk_g2s(s_fo, blob + d.off_fanoff(), d.fanoff_bytes(), &bar);
        if (d.fanv_bytes()) bulk_g2s(s_fv, blob + d.off_fanv(), d.fanv_bytes(), &bar);
        if (d.own_bytes(ELEM_V)) bulk_g2s(s_own, blob + d.off_own(ELEM_V), d.own_bytes(ELEM_V), &bar);
        if (d.stash_bytes()) bulk_g2s(s_stash, blob + d.off_stash(), d.stash_bytes(), &bar);
        if (r2) {
            if (d.r2idx_bytes()) bulk_g2s(s_r2i, blob + d.o_r2idx, d.r2idx_bytes(), &bar);
            bulk_g2s(s_r2o, blob + d.o_r2off, d.r2off_bytes(), &bar);
            if (d.r2val_bytes()) bulk_g2s(s_r2v, blob + d.o_r2val, d.r2val_bytes(), &bar);
            if (d.ext_bytes()) bulk_g2s(s_ext, blob + d.o_ext, d.ext_bytes(), &bar);
        }
        if (cap) bulk_g2s(s_out, x + 3ull * d.slot_base[ELEM_V], 12u * cap, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    // every extended local vertex as one float4: owned from the slice the TMA delivered, ribbon and ext vertices from
    // their owners' slots
    for (uint32_t i = threadIdx.x; i < nv + next; i += BT) {
        float4 q;
        if (i < nov) {
            q = make_float4(s_out[3 * i], s_out[3 * i + 1], s_out[3 * i + 2], 0.f);
        } else {
            const uint32_t o = i < nv ? s_own[i - nov] : s_ext[i - nv];
            const float*   g = x + 3ull * ((uint64_t)s_stash[o >> 16].slot_base[ELEM_V] + (o & 0xFFFFu));
            q = make_float4(__ldg(g), __ldg(g + 1), __ldg(g + 2), 0.f);  // L1 keeps the row's sector between the three loads
        }
        s_x[i] = q;
    }
    if (!r2)  // no ring extension: no not-owned vertex has a stored ring
        for (uint32_t i = threadIdx.x; i < nv - nov; i += BT)
            s_r2i[i] = 0xFFFFu;
    __syncthreads();
    BilPatch B;
    B.fo = s_fo, B.fv = s_fv, B.r2idx = s_r2i, B.r2off = s_r2o, B.r2val = s_r2v, B.x = s_x;
    B.nx_ = nv + next, B.nov = nov;
    uint32_t* bm  = s_priv;
    uint16_t* lst = reinterpret_cast<uint16_t*>(s_priv + bm_words * BT);
    for (uint32_t v = threadIdx.x; v < cap; v += BT) {
        float* const o = s_out + 3 * v;  // results go straight to the staging buffer (no local array behind a call)
        if (v < nov) {
            const uint32_t fb = s_fo[v] & FAN_OFF_MASK, fe = s_fo[v + 1] & FAN_OFF_MASK;
            if (fb == fe) {  // no neighbours: keeps its position
                o[0] = s_x[v].x, o[1] = s_x[v].y, o[2] = s_x[v].z;
            } else if (!bil_fast<BT>(B, v, bm, bm_words, lst, o)) {  // writes o only when it succeeds
                s_def[atomicAdd(&s_ndef, 1u)] = (uint16_t)v;
                o[0] = o[1] = o[2] = 0.f;
            }
        } else {
            o[0] = o[1] = o[2] = 0.f;
        }
    }
    __syncthreads();
    const uint32_t ndef = s_ndef;
    if (ndef) {
        // the vertices this patch could not finish go to the global work list of k_bilateral_deferred (one reservation per
        // block), with the normal and sigma_c^2 their fan gives; their rows of s_out are overwritten by that kernel
        if (threadIdx.x == 0) s_wbase = atomicAdd(work_count, ndef);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < ndef; i += BT) {
            const uint32_t v = s_def[i];
            BilDeferred    w;
            w.slot = d.slot_base[ELEM_V] + v;
            bil_normal_sigma(B, v, s_x[v].x, s_x[v].y, s_x[v].z, w.nx, w.ny, w.nz, w.sc2);
            work[s_wbase + i] = w;
        }
    }
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0 && cap) {
        bulk_s2g(xo + 3ull * d.slot_base[ELEM_V], s_out, 12u * cap);
        bulk_commit();
        bulk_wait_all_read();
    }
}

// --------------------------------------------------------------------------
// reductions over OWNED elements (ReduceHandle: reduce_handle.cu:53-156, kernels/reduce.cuh:43-191)
// kind: 0 dot, 1 sum of squares, 2 sum, 3 min, 4 max, 5 arg-min, 6 arg-max.  fp32 values, fp64 partials.
// --------------------------------------------------------------------------
struct RedPair
{
    double   v;
    uint64_t h;
};

__device__ __forceinline__ RedPair red_combine(int kind, RedPair a, RedPair b)
{
    if (kind <= 2) return RedPair{a.v + b.v, 0};
    const bool take_min = (kind == 3 || kind == 5);
    const bool b_better = take_min ? (b.v < a.v) : (b.v > a.v);
    if (b_better || (b.v == a.v && b.h < a.h)) return b;  // ties: smallest handle, deterministic
    return a;
}

__global__ void __launch_bounds__(256) k_reduce_stage1(MeshView mv, AttrView<float> a, AttrView<float> b, int elem, int kind,
                                                       uint32_t attr_id, RedPair* __restrict__ partial)
{
    __shared__ RedPair s_w[8];
    const double init = kind <= 2 ? 0.0 : ((kind == 3 || kind == 5) ? 1e300 : -1e300);
    RedPair      acc{init, INVALID64_};
    for (uint32_t p = blockIdx.x; p < mv.num_patches; p += gridDim.x) {
        const PatchDesc* d   = mv.desc + p;
        const uint32_t   no  = d->n_owned[elem], sb = d->slot_base[elem], cap = (no + 3u) & ~3u, pid = d->patch_id;
        const uint32_t   lb  = d->lin_base[elem];
        const uint32_t   na  = attr_id == INVALID32_ ? a.nattr : 1u;
        for (uint32_t i = threadIdx.x; i < no * na; i += blockDim.x) {
            const uint32_t lid = i % no, k = attr_id == INVALID32_ ? i / no : attr_id;
            const float    x = a.data[a.index_known(sb, cap, lb, lid, k)];
            RedPair        c;
            if (kind == 0)
                c = RedPair{(double)x * (double)b.data[b.index_known(sb, cap, lb, lid, k)], 0};
            else if (kind == 1)
                c = RedPair{(double)x * (double)x, 0};
            else
                c = RedPair{(double)x, ((uint64_t)pid << 32) | lid};
            acc = red_combine(kind, acc, c);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        RedPair t;
        t.v = __shfl_xor_sync(0xffffffffu, acc.v, o);
        t.h = __shfl_xor_sync(0xffffffffu, acc.h, o);
        acc = red_combine(kind, acc, t);
    }
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (uint32_t w = 1; w < blockDim.x / 32; ++w)
            acc = red_combine(kind, acc, s_w[w]);
        partial[blockIdx.x] = acc;
    }
}

__global__ void k_reduce_stage2(const RedPair* __restrict__ partial, uint32_t n, int kind, RedPair* __restrict__ out)
{
    RedPair acc = partial[0];
    for (uint32_t i = 1; i < n; ++i)  // n <= a few thousand: one thread, deterministic order
        acc = red_combine(kind, acc, partial[i]);
    *out = acc;
}

// --------------------------------------------------------------------------
// boundary vertices: an edge with one incident face marks its two vertices
// --------------------------------------------------------------------------
template <int KMAX, bool PACKED>
__global__ void __launch_bounds__(BT) k_boundary(MeshView mv, uint32_t* __restrict__ flag)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t                      bar;
    __shared__ uint32_t                      warp_tmp[36];
    using Q = PatchQuery<OP_EF, BT, KMAX, PACKED>;
    const uint32_t  p    = blockIdx.x;
    const PatchDesc d    = load_desc(mv.desc + p);
    const uint8_t*  blob = mv.topo + d.topo_off;
    Smem            sm(smem_raw);
    Q               q;
    q.plan(d, sm, false, false);
    uint16_t*   s_ev    = sm.alloc<uint16_t>(d.ev_bytes() / 2);
    uint32_t*   s_own   = sm.alloc<uint32_t>(d.own_bytes(ELEM_V) / 4);
    StashEntry* s_stash = sm.alloc<StashEntry>(d.n_stash);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(&bar, q.tx_bytes(d, false) + d.ev_bytes() + d.own_bytes(ELEM_V) + d.stash_bytes());
        q.issue(d, blob, &bar, false);
        if (d.ev_bytes()) bulk_g2s(s_ev, blob + d.off_ev(), d.ev_bytes(), &bar);
        if (d.own_bytes(ELEM_V)) bulk_g2s(s_own, blob + d.off_own(ELEM_V), d.own_bytes(ELEM_V), &bar);
        if (d.stash_bytes()) bulk_g2s(s_stash, blob + d.off_stash(), d.stash_bytes(), &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    const QueryResult r = q.compute(d, warp_tmp, false, false);
    OwnerTable        ot;
    ot.own = s_own, ot.stash = s_stash, ot.n_owned = d.n_owned[ELEM_V], ot.patch = d.patch_id;
    ot.slot_base = d.slot_base[ELEM_V], ot.type = ELEM_V;
    constexpr uint32_t vm = PACKED ? PK_ID_MASK : 0xFFFFu;
    for (uint32_t e = threadIdx.x; e < r.n_src; e += BT)
        if (r.size(e) == 1) {
            flag[ot.slot(s_ev[2 * e] & vm)]     = 1u;
            flag[ot.slot(s_ev[2 * e + 1] & vm)] = 1u;
        }
}

// --------------------------------------------------------------------------
// attribute helpers
// --------------------------------------------------------------------------
// same slots, another layout (AoS <-> AoSoA <-> SoA): lets the fixed-function kernels, which read AoS, serve attributes
// created with the reference's default layout
__global__ void k_relayout(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, const uint32_t* __restrict__ slot_base,
                           const uint32_t* __restrict__ lin_base, uint32_t num_patches, uint32_t num_slots, uint32_t num_elems,
                           uint32_t nattr, uint32_t layout_src, uint32_t layout_dst)
{
    AttrView<uint32_t> vs{nullptr, slot_base, num_slots, nattr, layout_src, lin_base, num_elems},
        vd{nullptr, slot_base, num_slots, nattr, layout_dst, lin_base, num_elems};
    const bool owned_only = layout_src == LAYOUT_SOA || layout_dst == LAYOUT_SOA;  // SoA has no padding slots
    for (uint32_t p = blockIdx.x; p < num_patches; p += gridDim.x) {
        const uint32_t b = slot_base[p], cap = slot_base[p + 1] - b, lb = lin_base[p];
        const uint32_t n = owned_only ? lin_base[p + 1] - lb : cap;
        for (uint32_t i = threadIdx.x; i < n * nattr; i += blockDim.x) {
            const uint32_t lid = i / nattr, a = i % nattr;
            dst[vd.index_known(b, cap, lb, lid, a)] = src[vs.index_known(b, cap, lb, lid, a)];
        }
        if (owned_only && layout_dst != LAYOUT_SOA)  // the source has no padding slots: define them in the copy
            for (uint32_t i = n * nattr + threadIdx.x; i < cap * nattr; i += blockDim.x)
                dst[vd.index_known(b, cap, lb, i / nattr, i % nattr)] = 0u;
    }
}

template <typename T, bool TO_SLOTS>
__global__ void k_permute(const T* __restrict__ src, T* __restrict__ dst, const uint32_t* __restrict__ slot_to_global,
                          const uint32_t* __restrict__ slot_base, const uint32_t* __restrict__ lin_base, uint32_t num_patches,
                          uint32_t num_slots, uint32_t num_elems, uint32_t nattr, uint32_t layout)
{
    for (uint32_t p = blockIdx.x; p < num_patches; p += gridDim.x) {
        const uint32_t b = slot_base[p], cap = slot_base[p + 1] - b;
        const uint32_t lb = lin_base[p], n = layout == LAYOUT_SOA ? lin_base[p + 1] - lb : cap;  // SoA: owned elements only
        for (uint32_t i = threadIdx.x; i < n * nattr; i += blockDim.x) {
            // enumerate in the slot-side storage order so the slot side is coalesced
            uint32_t lid, a;
            uint64_t sidx;
            if (layout == LAYOUT_AOS) {
                lid = i / nattr, a = i % nattr;
                sidx = (uint64_t)(b + lid) * nattr + a;
            } else if (layout == LAYOUT_SOA) {
                a = i / n, lid = i % n;
                sidx = (uint64_t)a * num_elems + lb + lid;
            } else {
                a = i / cap, lid = i % cap;
                sidx = (uint64_t)b * nattr + (uint64_t)a * cap + lid;
            }
            const uint32_t g = slot_to_global[b + lid];
            if (TO_SLOTS) {
                if (g != INVALID32_)
                    dst[sidx] = src[(uint64_t)g * nattr + a];
                else {
                    T z;
                    memset(&z, 0, sizeof(T));
                    dst[sidx] = z;
                }
            } else if (g != INVALID32_) {
                dst[(uint64_t)g * nattr + a] = src[sidx];
            }
        }
    }
}

// halo pack / unpack: rows of `row_words` 32-bit words, AoS attribute storage
template <bool GATHER>
__global__ void k_slot_rows(uint32_t* __restrict__ attr, const uint32_t* __restrict__ idx, uint64_t n, uint32_t row_words,
                            uint32_t* __restrict__ buf)
{
    const uint64_t total = n * row_words;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / row_words, w = i % row_words;
        const uint64_t a = (uint64_t)idx[r] * row_words + w;
        if (GATHER)
            buf[i] = attr[a];
        else
            attr[a] = buf[i];
    }
}

// direct NVLink P2P halo push: remote[remote_idx[r]] = local[local_idx[r]] (remote = peer-mapped pointer)
__global__ void k_push_rows(const uint32_t* __restrict__ local, const uint32_t* __restrict__ local_idx,
                            uint32_t* __restrict__ remote, const uint32_t* __restrict__ remote_idx, uint64_t n,
                            uint32_t row_words)
{
    const uint64_t total = n * row_words;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / row_words, w = i % row_words;
        remote[(uint64_t)remote_idx[r] * row_words + w] = local[(uint64_t)local_idx[r] * row_words + w];
    }
}

template <typename T>
__global__ void k_fill(T* __restrict__ data, uint64_t n, T v)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        data[i] = v;
}

template <typename K>
cudaError_t set_smem(K kernel, uint32_t bytes)
{
    if (bytes > 227u * 1024u) return cudaErrorInvalidValue;
    if (bytes > 48u * 1024u)
        return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    return cudaSuccess;
}

uint32_t r16(uint32_t x)
{
    return (x + 15u) & ~15u;
}

// KMAX needed so that one atomic per non-zero fits in registers: nnz <= KMAX*BT
int pick_kmax(uint32_t nnz)
{
    if (nnz <= 12u * BT) return 12;
    if (nnz <= 24u * BT) return 24;
    return 0;
}

}  // namespace

#define RXM_FAIL(msg)                 \
    do {                              \
        if (err) *err = msg;          \
        return cudaErrorInvalidValue; \
    } while (0)

#define RXM_LAUNCH_OP(KERNEL, KM, PK, ...)                                               \
    do {                                                                                 \
        switch (op) {                                                                    \
            case OP_VV: RXM_LAUNCH_ONE(KERNEL, OP_VV, KM, PK, __VA_ARGS__); break;       \
            case OP_VE: RXM_LAUNCH_ONE(KERNEL, OP_VE, KM, PK, __VA_ARGS__); break;       \
            case OP_VF: RXM_LAUNCH_ONE(KERNEL, OP_VF, KM, PK, __VA_ARGS__); break;       \
            case OP_EV: RXM_LAUNCH_ONE(KERNEL, OP_EV, KM, PK, __VA_ARGS__); break;       \
            case OP_EF: RXM_LAUNCH_ONE(KERNEL, OP_EF, KM, PK, __VA_ARGS__); break;       \
            case OP_FV: RXM_LAUNCH_ONE(KERNEL, OP_FV, KM, PK, __VA_ARGS__); break;       \
            case OP_FE: RXM_LAUNCH_ONE(KERNEL, OP_FE, KM, PK, __VA_ARGS__); break;       \
            case OP_FF: RXM_LAUNCH_ONE(KERNEL, OP_FF, KM, PK, __VA_ARGS__); break;       \
            RXM_EDGE4_CASES(KERNEL, KM, PK, __VA_ARGS__)                                 \
            default: RXM_FAIL("unsupported query op");                                   \
        }                                                                                \
    } while (0)

#define RXM_LAUNCH_ONE(KERNEL, OPV, KM, PK, ...)                                         \
    do {                                                                                 \
        using Q = dev::PatchQuery<OPV, BT, KM, PK>;                                      \
        smem    = Q::smem_bytes(lim.max_n, lim.max_not_owned, lim.max_stash, true,                                \
                                (OPV == OP_VV || OPV == OP_VE || OPV == OP_VF) ? fan_rows : (OPV == OP_EF ? stored_ef : stored_ff), \
                                fan_entries) + extra_smem(OPV);                                                \
        auto kern = KERNEL<OPV, KM, PK>;                                                 \
        e         = set_smem(kern, smem);                                                \
        if (e != cudaSuccess) RXM_FAIL("patch needs more shared memory than 227 KB");    \
        kern<<<mv.num_patches, BT, smem, stream>>>(__VA_ARGS__);                         \
    } while (0)


// The persistent, software-pipelined variants (rxm_persistent.cuh) are opt-in: measured on the
// 100 M-face grid they win only for the lightest kernel at 512-face patches (VV consume 0.46 ->
// 0.38 ms) and lose once patches hold 1024 faces, where block-per-patch kernels already run at
// 62-85 % of the HBM roofline (profiles/r01_tile_sweep.txt).
static bool use_persistent()
{
    return getenv("RXM_PERSIST") != nullptr;  // read per launch so tests can switch it
}

static int num_sms()
{
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// launch a persistent pipelined kernel: grid = #SMs x resident CTAs (<= #patches)
template <class W>
static cudaError_t launch_persistent(const MeshView& mv, const typename W::Args& a, const typename W::Layout& L,
                                     cudaStream_t stream)
{
    auto           kern = dev::k_persistent<W, BT>;
    const uint32_t smem = dev::PIPE_STAGES * L.stage_bytes;
    cudaError_t    e    = set_smem(kern, smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e          = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BT, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidValue;
    const uint32_t grid = std::min<uint32_t>(mv.num_patches, (uint32_t)(num_sms() * per_sm));
    kern<<<grid, BT, smem, stream>>>(mv, a, L);
    ++g_launches;
    return cudaGetLastError();
}

static FanLayout fan_layout(const KernelLimits& lim)
{
    FanLayout      L;
    const uint32_t capv = lim.max_owned[ELEM_V] + 4;
    uint32_t       o    = 0;
    L.o_fo = o, o += r16(2u * (lim.max_owned[ELEM_V] + 1) + 16);
    L.o_fv = o, o += r16(2u * lim.max_fan_total + 16);
    L.o_own = o, o += r16(4u * lim.max_not_owned[ELEM_V] + 16);
    L.o_stash = o, o += 16u * lim.max_stash + 16;
    L.o_xp = o, o += r16(12u * capv);
    L.o_x4 = o, o += 16u * lim.max_n[ELEM_V];
    L.stage_bytes = (o + 127u) & ~127u;
    return L;
}

// shared memory of the sections every fan kernel stages (offsets, neighbours, owner table, stash)
static uint32_t fan_smem(const KernelLimits& lim)
{
    return r16(2u * (lim.max_owned[ELEM_V] + 1) + 16) + r16(2u * lim.max_fan_total + 16) +
           r16(4u * lim.max_not_owned[ELEM_V]) + 16u * lim.max_stash;
}

static uint32_t max_nnz(const KernelLimits& lim)
{
    return std::max(2u * lim.max_n[ELEM_E], 3u * lim.max_n[ELEM_F]);
}

// EE / EVDiamond exist for the store kernel (and user lambdas through the headers) only
#define RXM_EDGE4_CASES(KERNEL, KM, PK, ...)                                              \
    case OP_EE: RXM_LAUNCH_ONE(KERNEL, OP_EE, KM, PK, __VA_ARGS__); break;               \
    case OP_EVDIAMOND: RXM_LAUNCH_ONE(KERNEL, OP_EVDIAMOND, KM, PK, __VA_ARGS__); break;
cudaError_t launch_query_store(int op, const MeshView& mv, const KernelLimits& lim, AttrView<uint64_t> in,
                               AttrView<uint64_t> out, cudaStream_t stream, const char** err)
{
    if ((op == OP_EE || op == OP_EVDIAMOND) && !mv.edge_manifold)
        RXM_FAIL("Op::EE / Op::EVDiamond only work on edge-manifold meshes (rxmesh_static.inl:519-525)");
    const int km = pick_kmax(max_nnz(lim));
    if (!km) RXM_FAIL("patch too large for the query kernels (nnz > 24*256)");
    if (op == OP_FF && lim.max_face_adjacent_faces > 6) RXM_FAIL("FF: more than 6 adjacent faces per face");
    uint32_t    smem = 0;
    cudaError_t e    = cudaSuccess;
    const uint32_t stored_ff  = mv.edge_manifold ? lim.max_owned[ELEM_F] : 0u;  // FF from the stored rows (plan(): ff3)
    const uint32_t stored_ef  = mv.edge_manifold ? lim.max_owned[ELEM_E] : 0u;  // EF from the stored pairs (plan(): ef3)
    const bool     fanq = mv.edge_manifold && mv.fans;                             // VV / VE / VF from the fans (plan(): fanq)
    const uint32_t fan_rows = fanq ? lim.max_owned[ELEM_V] : 0u, fan_entries = fanq ? lim.max_fan_total + 8u : 0u;
    auto        extra_smem = [&](int) { return 0u; };
    if (mv.packed)
        RXM_LAUNCH_OP(k_query_store, 1, true, mv, in, out);
    else if (km == 12)
        RXM_LAUNCH_OP(k_query_store, 12, false, mv, in, out);
    else
        RXM_LAUNCH_OP(k_query_store, 24, false, mv, in, out);
    ++g_launches;
    return cudaGetLastError();
}
#undef RXM_EDGE4_CASES
#define RXM_EDGE4_CASES(KERNEL, KM, PK, ...)

cudaError_t launch_query_consume(int op, const MeshView& mv, const KernelLimits& lim, AttrView<float> in,
                                 AttrView<float> out, cudaStream_t stream, const char** err)
{
    const int km = pick_kmax(max_nnz(lim));
    if (!km) RXM_FAIL("patch too large for the query kernels (nnz > 24*256)");
    if (in.nattr != 1 || out.nattr != 1) RXM_FAIL("query_consume needs single-component fp32 attributes");
    if (op == OP_VV && mv.fans && use_persistent()) {
        ConsumeLayout L;
        uint32_t      o = 0;
        L.o_a = o, o += r16(2u * (lim.max_owned[ELEM_V] + 1) + 16);
        L.o_b = o, o += r16(2u * lim.max_fan_total + 16);
        L.o_own = o, o += r16(4u * lim.max_not_owned[ELEM_V] + 16);
        L.o_stash = o, o += 16u * lim.max_stash + 16;
        L.o_in = o, o += r16(4u * (std::max(lim.max_n[ELEM_V], lim.max_owned[ELEM_V] + 4)));
        L.o_val = o;
        L.stage_bytes = (o + 127u) & ~127u;
        cudaError_t e = launch_persistent<VVFanConsumeWorker>(mv, VVFanConsumeWorker::Args{in.data, out.data}, L, stream);
        if (e == cudaErrorInvalidValue) RXM_FAIL("patch needs more shared memory than 227 KB");
        return e;
    }
    if (op == OP_VF && mv.packed && use_persistent()) {
        ConsumeLayout L;
        uint32_t      o = 0;
        L.o_a = o, o += r16(6u * lim.max_n[ELEM_F] + 16);
        L.o_b = o, o += r16(2u * (lim.max_owned[ELEM_V] + 1) + 16);
        L.o_own = o, o += r16(4u * lim.max_not_owned[ELEM_F] + 16);
        L.o_stash = o, o += 16u * lim.max_stash + 16;
        L.o_in = o, o += r16(4u * (std::max(lim.max_n[ELEM_F], lim.max_owned[ELEM_F] + 4)));
        L.o_val = o, o += r16(6u * lim.max_n[ELEM_F] + 16);
        L.stage_bytes = (o + 127u) & ~127u;
        cudaError_t e = launch_persistent<VFConsumeWorker>(mv, VFConsumeWorker::Args{in.data, out.data}, L, stream);
        if (e == cudaErrorInvalidValue) RXM_FAIL("patch needs more shared memory than 227 KB");
        return e;
    }
    if (op == OP_VF && mv.fans && !use_persistent()) {
        const uint32_t smem = r16(2u * (lim.max_owned[ELEM_V] + 1) + 16) + r16(2u * lim.max_fan_total + 16) +
                              r16(4u * lim.max_not_owned[ELEM_F]) + 16u * lim.max_stash +
                              r16(4u * (std::max(lim.max_n[ELEM_F], lim.max_owned[ELEM_F] + 4)));
        const bool small = smem <= 14000u && !getenv("RXM_CONSUME_BT256");  // 16 blocks of 128 threads fit one SM
        if (set_smem(k_vf_consume_fan<128>, smem) != cudaSuccess || set_smem(k_vf_consume_fan<256>, smem) != cudaSuccess)
            RXM_FAIL("patch needs more shared memory than 227 KB");
        if (small)
            k_vf_consume_fan<128><<<mv.num_patches, 128, smem, stream>>>(mv, in.data, out.data);
        else
            k_vf_consume_fan<256><<<mv.num_patches, 256, smem, stream>>>(mv, in.data, out.data);
        ++g_launches;
        return cudaGetLastError();
    }
    if (op == OP_VV && mv.fans) {
        const uint32_t smem = fan_smem(lim) + r16(4u * (std::max(lim.max_n[ELEM_V], lim.max_owned[ELEM_V] + 4)));
        const bool small = smem <= 14000u && !getenv("RXM_CONSUME_BT256");
        if (set_smem(k_vv_consume_fan<128>, smem) != cudaSuccess || set_smem(k_vv_consume_fan<256>, smem) != cudaSuccess)
            RXM_FAIL("patch needs more shared memory than 227 KB");
        if (small)
            k_vv_consume_fan<128><<<mv.num_patches, 128, smem, stream>>>(mv, in.data, out.data);
        else
            k_vv_consume_fan<256><<<mv.num_patches, 256, smem, stream>>>(mv, in.data, out.data);
        ++g_launches;
        return cudaGetLastError();
    }
    if (op == OP_FF && lim.max_face_adjacent_faces > 6) RXM_FAIL("FF: more than 6 adjacent faces per face");
    uint32_t    smem = 0;
    cudaError_t e    = cudaSuccess;
    const uint32_t stored_ff = mv.edge_manifold ? lim.max_owned[ELEM_F] : 0u;  // FF from the stored rows (plan(): ff3)
    const uint32_t stored_ef = mv.edge_manifold ? lim.max_owned[ELEM_E] : 0u;  // EF from the stored pairs (plan(): ef3)
    const bool     fanq = mv.edge_manifold && mv.fans;
    const uint32_t fan_rows = fanq ? lim.max_owned[ELEM_V] : 0u, fan_entries = fanq ? lim.max_fan_total + 8u : 0u;
    auto extra_smem = [&](int opv) {
        uint32_t dst = 0;
        switch (opv) {
            case OP_VV: case OP_EV: case OP_FV: dst = ELEM_V; break;
            case OP_VE: case OP_FE: dst = ELEM_E; break;
            default: dst = ELEM_F;
        }
        return r16(4u * (lim.max_n[dst] + 4u));
    };
    if (mv.packed)
        RXM_LAUNCH_OP(k_query_consume, 1, true, mv, in.data, out.data);
    else if (km == 12)
        RXM_LAUNCH_OP(k_query_consume, 12, false, mv, in.data, out.data);
    else
        RXM_LAUNCH_OP(k_query_consume, 24, false, mv, in.data, out.data);
    ++g_launches;
    return cudaGetLastError();
}

cudaError_t launch_vertex_normals(const MeshView& mv, const KernelLimits& lim, const float* x, float* n,
                                  int unit, cudaStream_t stream, const char** err)
{
    const uint32_t capv = lim.max_owned[ELEM_V] + 4;
    if (mv.fans && use_persistent()) {
        const FanLayout L = fan_layout(lim);
        cudaError_t     e;
        if (unit)
            e = launch_persistent<FanWorker<1>>(mv, FanWorker<1>::Args{x, n, 0.0}, L, stream);
        else
            e = launch_persistent<FanWorker<0>>(mv, FanWorker<0>::Args{x, n, 0.0}, L, stream);
        if (e == cudaErrorInvalidValue) RXM_FAIL("patch needs more shared memory than 227 KB");
        return e;
    }
    if (mv.fans && !getenv("RXM_VN_SCALAR")) {  // two vertices per thread, packed fp32x2 (the default)
        const uint32_t smem = fan_smem(lim) + r16(12u * std::max(lim.max_n[ELEM_V], capv)) + r16(12u * capv) + 64u;
        const bool  direct = getenv("RXM_STAGED_STORE") == nullptr;  // default: direct stores (staged: the round-1 form)
        const uint32_t sm2 = direct ? smem - r16(12u * capv) : smem;
        cudaError_t e = cudaSuccess;
#define RXM_VN(U, D)                                                            \
    do {                                                                        \
        e = set_smem(k_vertex_normals_fan2<U, D>, sm2);                         \
        if (e == cudaSuccess) k_vertex_normals_fan2<U, D><<<mv.num_patches, BT2, sm2, stream>>>(mv, x, n); \
    } while (0)
        if (unit && direct) RXM_VN(1, true);
        else if (unit) RXM_VN(1, false);
        else if (direct) RXM_VN(0, true);
        else RXM_VN(0, false);
#undef RXM_VN
        if (e != cudaSuccess) RXM_FAIL("patch needs more shared memory than 227 KB");
        ++g_launches;
        return cudaGetLastError();
    }
    if (mv.fans) {
        const uint32_t smem = fan_smem(lim) + r16(12u * capv) + 16u * lim.max_n[ELEM_V];
        cudaError_t e = unit ? set_smem(k_vertex_normals_fan<1>, smem) : set_smem(k_vertex_normals_fan<0>, smem);
        if (e != cudaSuccess) RXM_FAIL("patch needs more shared memory than 227 KB");
        if (unit)
            k_vertex_normals_fan<1><<<mv.num_patches, BT, smem, stream>>>(mv, x, n);
        else
            k_vertex_normals_fan<0><<<mv.num_patches, BT, smem, stream>>>(mv, x, n);
        ++g_launches;
        return cudaGetLastError();
    }
    if (mv.packed) {
        const uint32_t smem = r16(6u * lim.max_n[ELEM_F]) + r16(2u * (lim.max_owned[ELEM_V] + 1) + 16) +
                              r16(4u * lim.max_not_owned[ELEM_V]) + 16u * lim.max_stash + r16(12u * capv) +
                              16u * lim.max_n[ELEM_V] + 48u * lim.max_n[ELEM_F];
        cudaError_t e = unit ? set_smem(k_vertex_normals_pk<1>, smem) : set_smem(k_vertex_normals_pk<0>, smem);
        if (e != cudaSuccess) RXM_FAIL("patch needs more shared memory than 227 KB");
        if (unit)
            k_vertex_normals_pk<1><<<mv.num_patches, BT, smem, stream>>>(mv, x, n);
        else
            k_vertex_normals_pk<0><<<mv.num_patches, BT, smem, stream>>>(mv, x, n);
        ++g_launches;
        return cudaGetLastError();
    }
    const uint32_t smem = r16(6u * lim.max_n[ELEM_F]) + r16(4u * lim.max_not_owned[ELEM_V]) + 16u * lim.max_stash +
                          r16(12u * std::max(lim.max_n[ELEM_V], capv)) + r16(12u * capv);
    cudaError_t e = unit ? set_smem(k_vertex_normals<1>, smem) : set_smem(k_vertex_normals<0>, smem);
    if (e != cudaSuccess) RXM_FAIL("patch needs more shared memory than 227 KB");
    if (unit)
        k_vertex_normals<1><<<mv.num_patches, BT, smem, stream>>>(mv, x, n);
    else
        k_vertex_normals<0><<<mv.num_patches, BT, smem, stream>>>(mv, x, n);
    ++g_launches;
    return cudaGetLastError();
}

cudaError_t launch_laplacian_step(const MeshView& mv, const KernelLimits& lim, const float* x, float* xo, double lr,
                                  cudaStream_t stream, const char** err)
{
    const uint32_t capv = lim.max_owned[ELEM_V] + 4;
    if (mv.fans && use_persistent()) {
        cudaError_t e = launch_persistent<FanWorker<2>>(mv, FanWorker<2>::Args{x, xo, lr}, fan_layout(lim), stream);
        if (e == cudaErrorInvalidValue) RXM_FAIL("patch needs more shared memory than 227 KB");
        return e;
    }
    if (mv.fans && !getenv("RXM_VN_SCALAR")) {
        const uint32_t smem = fan_smem(lim) + r16(12u * std::max(lim.max_n[ELEM_V], capv)) + r16(12u * capv) + 64u;
        if (!getenv("RXM_STAGED_STORE")) {  // default: direct stores
            const uint32_t sm2 = smem - r16(12u * capv);
            if (set_smem(k_laplacian_fan2<false, true>, sm2) != cudaSuccess) RXM_FAIL("patch needs more shared memory than 227 KB");
            k_laplacian_fan2<false, true><<<mv.num_patches, BT2, sm2, stream>>>(mv, x, xo, lr, FusedHaloView{});
            ++g_launches;
            return cudaGetLastError();
        }
        if (set_smem(k_laplacian_fan2<false>, smem) != cudaSuccess) RXM_FAIL("patch needs more shared memory than 227 KB");
        k_laplacian_fan2<false><<<mv.num_patches, BT2, smem, stream>>>(mv, x, xo, lr, FusedHaloView{});
        ++g_launches;
        return cudaGetLastError();
    }
    if (mv.fans) {
        const uint32_t smem = fan_smem(lim) + r16(12u * capv) + 16u * lim.max_n[ELEM_V];
        if (set_smem(k_laplacian_fan, smem) != cudaSuccess) RXM_FAIL("patch needs more shared memory than 227 KB");
        k_laplacian_fan<<<mv.num_patches, BT, smem, stream>>>(mv, x, xo, lr);
        ++g_launches;
        return cudaGetLastError();
    }
    const int km = pick_kmax(2u * lim.max_n[ELEM_E]);
    if (!km) RXM_FAIL("patch too large for the query kernels (nnz > 24*256)");
    const uint32_t extra = r16(12u * std::max(lim.max_n[ELEM_V], capv)) + r16(12u * capv);
    uint32_t       smem;
    cudaError_t    e;
    if (mv.packed) {
        smem = dev::PatchQuery<OP_VV, BT, 1, true>::smem_bytes(lim.max_n, lim.max_not_owned, lim.max_stash, true) + extra;
        e    = set_smem(k_laplacian<1, true>, smem);
        if (e != cudaSuccess) RXM_FAIL("patch needs more shared memory than 227 KB");
        k_laplacian<1, true><<<mv.num_patches, BT, smem, stream>>>(mv, x, xo, lr);
    } else if (km == 12) {
        smem = dev::PatchQuery<OP_VV, BT, 12, false>::smem_bytes(lim.max_n, lim.max_not_owned, lim.max_stash, true) + extra;
        e    = set_smem(k_laplacian<12, false>, smem);
        if (e != cudaSuccess) RXM_FAIL("patch needs more shared memory than 227 KB");
        k_laplacian<12, false><<<mv.num_patches, BT, smem, stream>>>(mv, x, xo, lr);
    } else {
        smem = dev::PatchQuery<OP_VV, BT, 24, false>::smem_bytes(lim.max_n, lim.max_not_owned, lim.max_stash, true) + extra;
        e    = set_smem(k_laplacian<24, false>, smem);
        if (e != cudaSuccess) RXM_FAIL("patch needs more shared memory than 227 KB");
        k_laplacian<24, false><<<mv.num_patches, BT, smem, stream>>>(mv, x, xo, lr);
    }
    ++g_launches;
    return cudaGetLastError();
}

cudaError_t launch_laplacian_step_fused(const MeshView& mv, const KernelLimits& lim, const float* x, float* xo, double lr,
                                        const FusedHaloView& fh, cudaStream_t stream, const char** err)
{
    if (!mv.fans) RXM_FAIL("the fused Laplacian + halo kernel needs the one-ring fans (manifold, consistently oriented input)");
    if (fh.npeers > BT2) RXM_FAIL("too many neighbour ranks");
    const uint32_t capv = lim.max_owned[ELEM_V] + 4;
    uint32_t smem = fan_smem(lim) + r16(12u * std::max(lim.max_n[ELEM_V], capv)) + r16(12u * capv) + 64u;
    if (!getenv("RXM_STAGED_STORE")) {
        smem -= r16(12u * capv);
        if (set_smem(k_laplacian_fan2<true, true>, smem) != cudaSuccess) RXM_FAIL("patch needs more shared memory than 227 KB");
        k_laplacian_fan2<true, true><<<mv.num_patches, BT2, smem, stream>>>(mv, x, xo, lr, fh);
        ++g_launches;
        return cudaGetLastError();
    }
    if (set_smem(k_laplacian_fan2<true>, smem) != cudaSuccess) RXM_FAIL("patch needs more shared memory than 227 KB");
    k_laplacian_fan2<true><<<mv.num_patches, BT2, smem, stream>>>(mv, x, xo, lr, fh);
    ++g_launches;
    return cudaGetLastError();
}

cudaError_t launch_boundary_vertices(const MeshView& mv, const KernelLimits& lim, uint32_t* flag, cudaStream_t stream,
                                     const char** err)
{
    const int km = pick_kmax(3u * lim.max_n[ELEM_F]);
    if (!km) RXM_FAIL("patch too large for the query kernels (nnz > 24*256)");
    const uint32_t extra = r16(4u * lim.max_n[ELEM_E]) + r16(4u * lim.max_not_owned[ELEM_V]) + 16u * lim.max_stash;
    uint32_t       smem;
    cudaError_t    e;
    if (mv.packed) {
        smem = dev::PatchQuery<OP_EF, BT, 1, true>::smem_bytes(lim.max_n, lim.max_not_owned, lim.max_stash, false) + extra;
        e    = set_smem(k_boundary<1, true>, smem);
        if (e != cudaSuccess) RXM_FAIL("patch needs more shared memory than 227 KB");
        k_boundary<1, true><<<mv.num_patches, BT, smem, stream>>>(mv, flag);
    } else if (km == 12) {
        smem = dev::PatchQuery<OP_EF, BT, 12, false>::smem_bytes(lim.max_n, lim.max_not_owned, lim.max_stash, false) + extra;
        e    = set_smem(k_boundary<12, false>, smem);
        if (e != cudaSuccess) RXM_FAIL("patch needs more shared memory than 227 KB");
        k_boundary<12, false><<<mv.num_patches, BT, smem, stream>>>(mv, flag);
    } else {
        smem = dev::PatchQuery<OP_EF, BT, 24, false>::smem_bytes(lim.max_n, lim.max_not_owned, lim.max_stash, false) + extra;
        e    = set_smem(k_boundary<24, false>, smem);
        if (e != cudaSuccess) RXM_FAIL("patch needs more shared memory than 227 KB");
        k_boundary<24, false><<<mv.num_patches, BT, smem, stream>>>(mv, flag);
    }
    ++g_launches;
    return cudaGetLastError();
}

cudaError_t launch_query_csr(int op, const MeshView& mv, const KernelLimits& lim, const uint32_t* patch_nnz_off,
                             uint32_t* csr_off, uint32_t* csr_val, cudaStream_t stream, const char** err)
{
    const int km = pick_kmax(max_nnz(lim));
    if (!km) RXM_FAIL("patch too large for the query kernels (nnz > 24*256)");
    if (op == OP_FF && lim.max_face_adjacent_faces > 6) RXM_FAIL("FF: more than 6 adjacent faces per face");
    uint32_t    smem = 0;
    cudaError_t e    = cudaSuccess;
    const uint32_t stored_ff = 0, stored_ef = 0, fan_rows = 0, fan_entries = 0;  // k_query_csr plans the generic path (gap-free ascending lists)
    auto        extra_smem = [&](int) { return 0u; };
    if (mv.packed)
        RXM_LAUNCH_OP(k_query_csr, 1, true, mv, patch_nnz_off, csr_off, csr_val);
    else if (km == 12)
        RXM_LAUNCH_OP(k_query_csr, 12, false, mv, patch_nnz_off, csr_off, csr_val);
    else
        RXM_LAUNCH_OP(k_query_csr, 24, false, mv, patch_nnz_off, csr_off, csr_val);
    ++g_launches;
    return cudaGetLastError();
}

cudaError_t launch_bilateral_step(const uint32_t* csr_off, const uint32_t* csr_val, uint32_t num_slots, const float* x,
                                  const float* normals, float* xo, uint32_t* overflow_flag, cudaStream_t stream)
{
    if (num_slots == 0) return cudaSuccess;
    k_bilateral<<<(num_slots + 127) / 128, 128, 0, stream>>>(csr_off, csr_val, x, normals, xo, num_slots, overflow_flag);
    ++g_launches;
    return cudaGetLastError();
}

cudaError_t launch_bilateral_patch(const MeshView& mv, const KernelLimits& lim, const uint32_t* csr_off, const uint32_t* csr_val,
                                   const float* x, float* xo, uint32_t* flags, void* work, uint32_t iteration, cudaStream_t stream,
                                   const char** err)
{
    if (!mv.fans) RXM_FAIL("the patch-local bilateral kernel needs the one-ring fans");
    const uint32_t capv = lim.max_owned[ELEM_V] + 4, nvx = lim.max_n[ELEM_V] + lim.max_ext;
    const uint32_t bm_words = (nvx + 31u) / 32u;
    const uint32_t fixed = fan_smem(lim) + r16(2u * (lim.max_not_owned[ELEM_V] + lim.max_ext) + 32) + r16(2u * (lim.max_r2 + 1) + 16) +
                           r16(2u * lim.max_r2_total + 16) + r16(4u * lim.max_ext + 16) + 16u * nvx + r16(12u * capv) +
                           r16(2u * (lim.max_owned[ELEM_V] + 8)) + 64u;
    // block size: the compiled size that keeps most warps resident (per-thread bitmap + list storage grows with the block,
    // the staged patch does not); measured on the 10 M-face torus (561 owned vertices per patch, profiles/r02_bilateral_bt.txt):
    // 512 threads x 2 blocks = 32 warps 0.454 ms, 256 x 3 = 24 warps 0.470, 128 x 4 = 16 warps 0.553, 576 x 1 = 18 warps
    // 0.705 -- resident warps decide, idle warps of a half-empty second round cost nothing; larger block on a tie
    static const uint32_t sizes[] = {128, 192, 256, 288, 320, 384, 448, 512, 576};
    uint32_t    bt = 0, smem = 0, best_warps = 0;
    const char* force = getenv("RXM_BILATERAL_BT");  // experiment knob
    for (uint32_t t : sizes) {
        if (force && t != (uint32_t)atoi(force)) continue;
        if (!force && t > 128u && t >= 2u * lim.max_owned[ELEM_V]) break;  // more than half the block would never have a vertex
        const uint32_t sm = fixed + r16(4u * (bm_words + BIL_FAST_CAP / 2) * t);
        if (sm > 227u * 1024u) continue;
        const uint32_t blocks = std::min(227u * 1024u / (sm + 1024u), 2048u / t), warps = blocks * t / 32u;
        if (warps >= best_warps) best_warps = warps, bt = t, smem = sm;
    }
    if (!bt) RXM_FAIL("patch needs more shared memory than 227 KB");
    // flags: [0] overflow, [1] deferred vertices of the call, [2], [3] work-list fill of even / odd iterations
    uint32_t* cnt = flags + 2 + (iteration & 1u);
#define RXM_BIL(T)                                                                                                       \
    case T:                                                                                                              \
        if (set_smem(k_bilateral_patch<T>, smem) != cudaSuccess) RXM_FAIL("patch needs more shared memory than 227 KB");  \
        k_bilateral_patch<T><<<mv.num_patches, T, smem, stream>>>(mv, x, xo, bm_words, cnt, (BilDeferred*)work);          \
        break;
    switch (bt) {
        RXM_BIL(128) RXM_BIL(192) RXM_BIL(256) RXM_BIL(288) RXM_BIL(320) RXM_BIL(384) RXM_BIL(448) RXM_BIL(512) RXM_BIL(576)
        default: RXM_FAIL("internal: block size not compiled");
    }
#undef RXM_BIL
    k_bilateral_deferred<<<148 * 8, BIL_BT, 0, stream>>>(cnt, flags + 2 + ((iteration + 1u) & 1u), (const BilDeferred*)work, csr_off,
                                                          csr_val, x, xo, flags);
    g_launches += 2;
    return cudaGetLastError();
}

template <bool TO_SLOTS>
static cudaError_t permute_dispatch(const void* src, void* dst, const uint32_t* s2g, const SlotMap& sm, uint32_t elem_bytes,
                                    uint32_t nattr, uint32_t layout, cudaStream_t stream)
{
    const uint32_t grid = std::min<uint32_t>(sm.num_patches, 148u * 16u);
    if (grid == 0) return cudaSuccess;
#define RXM_PERMUTE(T)                                                                                              \
    k_permute<T, TO_SLOTS><<<grid, 128, 0, stream>>>((const T*)src, (T*)dst, s2g, sm.slot_base, sm.lin_base,        \
                                                     sm.num_patches, sm.num_slots, sm.num_elems, nattr, layout)
    if (elem_bytes == 4)
        RXM_PERMUTE(uint32_t);
    else if (elem_bytes == 8)
        RXM_PERMUTE(uint64_t);
    else if (elem_bytes == 2)
        RXM_PERMUTE(uint16_t);
    else if (elem_bytes == 1)
        RXM_PERMUTE(uint8_t);
    else
        return cudaErrorInvalidValue;
#undef RXM_PERMUTE
    ++g_launches;
    return cudaGetLastError();
}

cudaError_t launch_relayout(const void* src, void* dst, const SlotMap& sm, uint32_t nattr, uint32_t layout_src,
                            uint32_t layout_dst, cudaStream_t stream)
{
    const uint32_t grid = std::min<uint32_t>(sm.num_patches, 148u * 16u);
    if (grid == 0) return cudaSuccess;
    k_relayout<<<grid, 128, 0, stream>>>((const uint32_t*)src, (uint32_t*)dst, sm.slot_base, sm.lin_base, sm.num_patches,
                                         sm.num_slots, sm.num_elems, nattr, layout_src, layout_dst);
    ++g_launches;
    return cudaGetLastError();
}

cudaError_t launch_permute_to_slots(const void* in_global, void* out_slots, const uint32_t* s2g, const SlotMap& sm,
                                    uint32_t elem_bytes, uint32_t nattr, uint32_t layout, cudaStream_t stream)
{
    return permute_dispatch<true>(in_global, out_slots, s2g, sm, elem_bytes, nattr, layout, stream);
}

cudaError_t launch_permute_to_global(const void* in_slots, void* out_global, const uint32_t* s2g, const SlotMap& sm,
                                     uint32_t elem_bytes, uint32_t nattr, uint32_t layout, cudaStream_t stream)
{
    return permute_dispatch<false>(in_slots, out_global, s2g, sm, elem_bytes, nattr, layout, stream);
}

cudaError_t launch_slot_rows(bool gather, void* attr, const uint32_t* idx, uint64_t n, uint32_t row_words, void* buf,
                             cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    const uint32_t grid = (uint32_t)std::min<uint64_t>((n * row_words + 255) / 256, 148ull * 16);
    if (gather)
        k_slot_rows<true><<<grid, 256, 0, stream>>>((uint32_t*)attr, idx, n, row_words, (uint32_t*)buf);
    else
        k_slot_rows<false><<<grid, 256, 0, stream>>>((uint32_t*)attr, idx, n, row_words, (uint32_t*)buf);
    ++g_launches;
    return cudaGetLastError();
}

cudaError_t launch_push_rows(const void* local, const uint32_t* local_idx, void* remote, const uint32_t* remote_idx,
                             uint64_t n, uint32_t row_words, cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    const uint32_t grid = (uint32_t)std::min<uint64_t>((n * row_words + 255) / 256, 148ull * 16);
    k_push_rows<<<grid, 256, 0, stream>>>((const uint32_t*)local, local_idx, (uint32_t*)remote, remote_idx, n, row_words);
    ++g_launches;
    return cudaGetLastError();
}

cudaError_t launch_reduce(const MeshView& mv, AttrView<float> a, AttrView<float> b, int elem, int kind, uint32_t attr_id,
                          void* scratch /* (grid+1) * 16 bytes */, uint32_t grid, cudaStream_t stream)
{
    RedPair* part = reinterpret_cast<RedPair*>(scratch);
    k_reduce_stage1<<<grid, 256, 0, stream>>>(mv, a, b, elem, kind, attr_id, part);
    k_reduce_stage2<<<1, 1, 0, stream>>>(part, grid, kind, part + grid);
    g_launches += 2;
    return cudaGetLastError();
}

cudaError_t launch_fill(void* data, uint64_t count, uint32_t elem_bytes, const void* value, cudaStream_t stream)
{
    if (count == 0) return cudaSuccess;
    const uint32_t grid = (uint32_t)std::min<uint64_t>((count + 255) / 256, 148ull * 32);
    if (elem_bytes == 4) {
        uint32_t v;
        memcpy(&v, value, 4);
        k_fill<uint32_t><<<grid, 256, 0, stream>>>((uint32_t*)data, count, v);
    } else if (elem_bytes == 8) {
        uint64_t v;
        memcpy(&v, value, 8);
        k_fill<uint64_t><<<grid, 256, 0, stream>>>((uint64_t*)data, count, v);
    } else if (elem_bytes == 2) {
        uint16_t v;
        memcpy(&v, value, 2);
        k_fill<uint16_t><<<grid, 256, 0, stream>>>((uint16_t*)data, count, v);
    } else if (elem_bytes == 1) {
        uint8_t v;
        memcpy(&v, value, 1);
        k_fill<uint8_t><<<grid, 256, 0, stream>>>((uint8_t*)data, count, v);
    } else
        return cudaErrorInvalidValue;
    ++g_launches;
    return cudaGetLastError();
}

}  // namespace rxm
