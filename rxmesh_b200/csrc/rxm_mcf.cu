// rxm_mcf.cu -- mean-curvature-flow smoothing by matrix-free conjugate gradients, sm_100a.
//
// What it replaces (reference): apps/MCF/mcf_cg_mat_free.h:13-178 (the driver), apps/MCF/mcf_kernels.cuh:57-115 (init_B),
// :117-205 (matvec), include/rxmesh/matrix/cg_mat_free_attr_solver.h:45-125 (pre_solve / solve: S = A P, alpha, X += alpha P,
// R -= alpha S, delta, beta, P = R + beta P) and the two ReduceHandle calls of every iteration (reduce_handle.cu:53-156).
// The system is (M + dt L) X = M X0 with M the lumped (mixed-Voronoi or valence) mass and L the cotangent or uniform
// Laplacian; one right-hand side per coordinate, solved together (the reference's dot / norm2 run over all three).
//
// Design (B200-first, not a translation):
//   * The reference recomputes every cotangent weight and Voronoi area from the coordinates in EVERY mat-vec (about 900
//     flops per vertex against 60 bytes) and runs six kernels and two host synchronisations per CG iteration (mat-vec, dot,
//     two axpy, norm2, axpy).  The matrix does not change during a solve, so k_mcf_setup evaluates the weights ONCE, with
//     the reference's fp32 formulas in the reference's order, into a per-fan-entry array W (24 B per vertex, resident in
//     HBM) and a per-slot diagonal; the same kernel forms B, S0 = A X0, R0 = B - S0 and <R0, R0>.
//   * An iteration is TWO kernels and no host synchronisation:
//       k_mcf_matvec  (one block per patch)  P' = R + beta P for the patch's own AND ribbon vertices (the ribbon rows are
//                     gathered as (R, P) pairs from their owners' slots: the axpy that updates P is fused into the load of
//                     the mat-vec, P is double-buffered so the owners' writes cannot race the neighbours' reads),
//                     S = A P' from W / diag staged by TMA bulk copies, and the block's share of <S, P'>;
//       k_mcf_update  (grid-stride, float4)  X += alpha P', R -= alpha S, the block's share of <R, R>.
//     Both end in a last-block reduction of the per-block fp64 partial sums IN BLOCK ORDER (deterministic: the same bits
//     every run) that also does the scalar work of the solver on the device: alpha, beta, the convergence test of
//     IterativeSolver::is_converged (iterative_solver.h:57-63) and the iteration counter.  Once `converged` is set every
//     later kernel returns at its first instruction, so the host may queue iterations in batches and look at the state
//     only between batches; X is exactly what the reference's loop leaves (updated in the converging iteration).
//   * Traffic per iteration and vertex: mat-vec 24 (R, P) + 24 (W) + 12 (fan ids) + 4 (diag) + 24 (P', S) + ribbon rows,
//     update 48 + 24: about 165 B against about 230 B + the weight arithmetic for the six-kernel form.
#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "rxm_kernels.h"
#include "rxmesh_b200/rxm_device.cuh"

namespace rxm {
namespace {

using namespace dev;

constexpr int MBT = 256;  // threads per block of the setup and update kernels

__device__ __forceinline__ PatchDesc mcf_load_desc(const PatchDesc* g)
{
    PatchDesc    d;
    const uint4* s = reinterpret_cast<const uint4*>(g);
    uint4*       t = reinterpret_cast<uint4*>(&d);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(PatchDesc) / 16); ++i)
        t[i] = __ldg(s + i);
    return d;
}

// sum of `v` over the block, valid in thread 0.  Fixed shuffle tree + fixed order over the warps: deterministic.
template <int BT = MBT>
__device__ __forceinline__ double block_sum(double v, double* s_red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_down_sync(0xFFFFFFFFu, v, o);
    __syncthreads();  // s_red may still be read by a previous call
    if ((threadIdx.x & 31u) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double a = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < BT / 32; ++w)
            a += s_red[w];
    return a;
}

// Publish this block's partial sum; true (in every thread) in the block that arrives last.  The last block then sums
// partials[0 .. gridDim.x) in index order.
__device__ __forceinline__ bool publish_partial(double part, double* partials, uint32_t* ctr, uint32_t* s_flag)
{
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = part;
        __threadfence();
        *s_flag = atomicAdd(ctr, 1u) == gridDim.x - 1u ? 1u : 0u;
    }
    __syncthreads();
    return *s_flag != 0u;
}
template <int BT = MBT>
__device__ __forceinline__ double sum_partials(const double* partials, double* s_red)
{
    __threadfence();
    double a = 0.0;
    for (uint32_t i = threadIdx.x; i < gridDim.x; i += BT)
        a += __ldcg(partials + i);
    return block_sum<BT>(a, s_red);
}

// ---- the reference's geometry helpers in fp32 (include/rxmesh/geometry_util.cuh:52-58 tri_area, :107-113 clamp_cot,
//      :120-172 partial_voronoi_area, :178-206 edge_cotan_weight), same operations in the same order ----
struct V3
{
    float x, y, z;
};
__device__ __forceinline__ V3 sub(V3 a, V3 b)
{
    return V3{a.x - b.x, a.y - b.y, a.z - b.z};
}
__device__ __forceinline__ float dot(V3 a, V3 b)
{
    return a.x * b.x + a.y * b.y + a.z * b.z;
}
__device__ __forceinline__ float tri_area(V3 p0, V3 p1, V3 p2)
{
    const V3 u = sub(p1, p0), v = sub(p2, p0);
    const V3 c = V3{u.y * v.z - v.y * u.z, u.z * v.x - v.z * u.x, u.x * v.y - v.x * u.y};
    return 0.5f * sqrtf(dot(c, c));
}
__device__ __forceinline__ float clamp_cot(float v)
{
    const float bound = 19.1f;
    return (v < -bound) ? -bound : ((v > bound) ? bound : v);
}
constexpr float FLT_MIN_ = 1.17549435e-38f;  // std::numeric_limits<float>::min()
__device__ __forceinline__ float partial_voronoi_area(V3 p, V3 q, V3 r)
{
    const V3    pq = sub(q, p), qr = sub(r, q), pr = sub(r, p);
    const float area = tri_area(p, q, r);
    if (area <= FLT_MIN_) return -1.f;
    const float dotp = dot(pq, pr), dotq = -dot(qr, pq), dotr = dot(qr, pr);
    if (dotp < 0.f) return 0.25f * area;
    if (dotq < 0.f || dotr < 0.f) return 0.125f * area;
    const float cotq = clamp_cot(dotq / area), cotr = clamp_cot(dotr / area);
    return 0.125f * (dot(pr, pr) * cotq + dot(pq, pq) * cotr);
}
__device__ __forceinline__ float cot_part(V3 p, V3 r, V3 v)
{
    const V3    d0 = sub(p, v), d1 = sub(r, v);
    const float area = tri_area(p, r, v);
    return area > FLT_MIN_ ? clamp_cot(dot(d0, d1) / area) : 0.f;
}

// What every patch kernel stages: fan offsets, fan neighbours, the owner table of the ribbon vertices, the stash.
struct Staged
{
    uint16_t*   fo;
    uint16_t*   fv;
    uint32_t*   own;
    StashEntry* stash;
    uint32_t    bytes;  // what the bulk copies issued by stage_topology deliver
};
__device__ __forceinline__ Staged stage_alloc(Smem& sm, const PatchDesc& d)
{
    Staged s;
    s.fo    = sm.alloc<uint16_t>(d.fanoff_bytes() / 2);
    s.fv    = sm.alloc<uint16_t>(d.fanv_bytes() / 2 + 8u);
    s.own   = sm.alloc<uint32_t>(d.own_bytes(ELEM_V) / 4 + 4u);
    s.stash = sm.alloc<StashEntry>(d.n_stash + 1u);
    s.bytes = d.fanoff_bytes() + d.fanv_bytes() + d.own_bytes(ELEM_V) + d.stash_bytes();
    return s;
}
// thread 0 only, after arrive_expect_tx
__device__ __forceinline__ void stage_issue(const Staged& s, const PatchDesc& d, const uint8_t* blob, uint64_t* bar)
{
    if (d.fanoff_bytes()) bulk_g2s(s.fo, blob + d.off_fanoff(), d.fanoff_bytes(), bar);
    if (d.fanv_bytes()) bulk_g2s(s.fv, blob + d.off_fanv(), d.fanv_bytes(), bar);
    if (d.own_bytes(ELEM_V)) bulk_g2s(s.own, blob + d.off_own(ELEM_V), d.own_bytes(ELEM_V), bar);
    if (d.stash_bytes()) bulk_g2s(s.stash, blob + d.off_stash(), d.stash_bytes(), bar);
}

// --------------------------------------------------------------------------
// setup: W, diag, R0 = B - A X0, <R0, R0>           (init_B + the pre_solve mat-vec + init_PR + norm2, one pass)
// --------------------------------------------------------------------------
template <bool UNIFORM, bool PRECOND>
__global__ void __launch_bounds__(MBT) k_mcf_setup(MeshView mv, const float* __restrict__ x0, McfBuffers B, float dt)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t                      bar;
    __shared__ double                        s_red[MBT / 32];
    __shared__ uint32_t                      s_flag;
    const PatchDesc d    = mcf_load_desc(mv.desc + blockIdx.x);
    const uint8_t*  blob = mv.topo + d.topo_off;
    const uint32_t  nv = d.n[ELEM_V], nov = d.n_owned[ELEM_V], cap = d.slot_cap(ELEM_V);
    Smem            sm(smem_raw);
    const Staged    T   = stage_alloc(sm, d);
    float*          s_x = sm.alloc<float>(3u * max(nv, cap) + 4u);
    const uint64_t  g   = 3ull * d.slot_base[ELEM_V];
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(&bar, T.bytes + 12u * cap);
        stage_issue(T, d, blob, &bar);
        if (cap) bulk_g2s(s_x, x0 + g, 12u * cap, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    for (uint32_t i = nov + threadIdx.x; i < nv; i += MBT) {  // ribbon vertices: from their owners' slots
        const uint32_t o  = T.own[i - nov];
        const float*   gp = x0 + 3ull * ((uint64_t)T.stash[o >> 16].slot_base[ELEM_V] + (o & 0xFFFFu));
        s_x[3 * i] = __ldg(gp), s_x[3 * i + 1] = __ldg(gp + 1), s_x[3 * i + 2] = __ldg(gp + 2);
    }
    __syncthreads();
    auto at = [&](uint32_t u) { return V3{s_x[3 * u], s_x[3 * u + 1], s_x[3 * u + 2]}; };
    const uint32_t wb   = B.fan_base[blockIdx.x];
    double         part = 0.0;
    for (uint32_t v = threadIdx.x; v < cap; v += MBT) {
        float r0 = 0.f, r1 = 0.f, r2 = 0.f, dg = 0.f;
        if (v < nov) {
            const uint32_t b = T.fo[v] & FAN_OFF_MASK, e = T.fo[v + 1] & FAN_OFF_MASK;
            if (e > b) {
                const V3 p     = at(v);
                float    sum_e = 0.f, vw = 0.f, x = 0.f, y = 0.f, z = 0.f;
                V3       q = at(T.fv[e - 1]);  // iter.back()
                for (uint32_t i = b; i < e; ++i) {
                    const V3 r = at(T.fv[i]);
                    float    w;
                    if (UNIFORM) {
                        w = 1.f;
                    } else {
                        const V3 s = at(T.fv[i + 1 == e ? b : i + 1]);
                        w          = cot_part(p, r, q) + cot_part(p, r, s);  // edge_cotan_weight(p, r, q, s)
                        w          = (w >= 0.f ? 1.f : 0.f) * w;
                    }
                    w *= dt;
                    if (!UNIFORM) B.W[wb + i] = w;
                    sum_e += w;
                    x -= w * r.x, y -= w * r.y, z -= w * r.z;
                    if (UNIFORM) {
                        vw += 1.f;
                    } else {
                        const float ta = partial_voronoi_area(p, q, r);
                        vw += ta > 0.f ? ta : 0.f;
                        q = r;
                    }
                }
                // v_weight = 0.5 / v_weight (1.0 / for the uniform Laplacian), diag = 1.0 / v_weight + sum_e_weight: the
                // reference's double-precision literals make both a double division rounded to float
                const float vwt = (float)((UNIFORM ? 1.0 : 0.5) / (double)vw);
                dg              = (float)((1.0 / (double)vwt) + (double)sum_e);
                // init_B: X * valence (uniform), X / v_weight (cotangent)
                const float bx = UNIFORM ? p.x * vw : p.x / vwt, by = UNIFORM ? p.y * vw : p.y / vwt,
                            bz = UNIFORM ? p.z * vw : p.z / vwt;
                r0 = bx - (x + dg * p.x), r1 = by - (y + dg * p.y), r2 = bz - (z + dg * p.z);
            }
        }
        B.diag[d.slot_base[ELEM_V] + v] = dg;
        B.R[g + 3 * v] = r0, B.R[g + 3 * v + 1] = r1, B.R[g + 3 * v + 2] = r2;
        if (PRECOND) {  // delta = <R0, Z0>, Z0 = R0 / diag (precond_matvec, mcf_kernels.cuh:216-295)
            const float z0 = dg != 0.f ? r0 / dg : 0.f, z1 = dg != 0.f ? r1 / dg : 0.f, z2 = dg != 0.f ? r2 / dg : 0.f;
            part += (double)r0 * z0 + (double)r1 * z1 + (double)r2 * z2;
        } else {
            part += (double)r0 * r0 + (double)r1 * r1 + (double)r2 * r2;
        }
    }
    part = block_sum(part, s_red);
    if (publish_partial(part, B.partials, &B.state->ctr, &s_flag)) {
        const double a = fabs(sum_partials(B.partials, s_red));  // pcg_mat_free_attr_solver.h:61-62 takes |<R, P>|
        if (threadIdx.x == 0) {
            McfState* st  = B.state;
            st->ctr       = 0;
            st->start     = a;
            st->delta_new = a;
            st->delta_old = a;
            st->final_res = a;
            st->dot_sp    = 0.0;
            st->alpha     = 0.f;
            st->beta      = 0.f;  // first iteration: P = R (init_PR)
            st->iters     = 0;
            st->converged = a == 0.0 ? 1u : 0u;  // X0 already solves the system: nothing to iterate on (alpha would be 0 / 0)
        }
    }
}

// --------------------------------------------------------------------------
// iteration, first half: P' = R + beta P, S = A P', <S, P'>
// --------------------------------------------------------------------------
// BT: the launcher picks the block size that leaves the fewest idle threads in the last round over the owned vertices
// (561-vertex tiles: 3 x 192 instead of 256 + 256 + 49).  FIRST: iteration 0, P' = R (beta = 0; P is not read, so the
// solve needs no zeroed P buffer).
// WSTAGE = false (cotangent only): the weights are read from global memory where they are used (a thread's six are 24
// contiguous bytes, a warp's 768) instead of being staged: 13.5 KB less shared memory per 561-vertex tile, 8 resident blocks
// instead of 6.
template <bool UNIFORM, int BT, bool PRECOND, bool WSTAGE = true>
__global__ void __launch_bounds__(BT, 1536 / BT) k_mcf_matvec(MeshView mv, McfBuffers B, const float* __restrict__ Pold,
                                                   float* __restrict__ Pnew, float dt, int first)
{
    // the descriptor and the solver state are fetched together (one DRAM round trip at the head of the block, not two)
    const PatchDesc d    = mcf_load_desc(mv.desc + blockIdx.x);
    const uint32_t  conv = B.state->converged;
    const float     beta = B.state->beta;
    if (conv) return;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t                      bar;
    __shared__ double                        s_red[BT / 32];
    __shared__ uint32_t                      s_flag;
    const uint8_t*  blob = mv.topo + d.topo_off;
    const uint32_t  nv = d.n[ELEM_V], nov = d.n_owned[ELEM_V], cap = d.slot_cap(ELEM_V);
    const uint32_t  nw = (d.fan_total + 3u) & ~3u;  // entries of this patch's slice of W
    Smem            sm(smem_raw);
    const Staged    T    = stage_alloc(sm, d);
    float*          s_w  = sm.alloc<float>(UNIFORM || !WSTAGE ? 4u : nw + 4u);
    const float*    Wg   = UNIFORM ? nullptr : B.W + B.fan_base[blockIdx.x];
    float*          s_dg = sm.alloc<float>(cap + 4u);
    float*          s_p  = sm.alloc<float>(3u * max(nv, cap) + 4u);
    const uint64_t  g    = 3ull * d.slot_base[ELEM_V];
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(&bar, T.bytes + (UNIFORM || !WSTAGE ? 0u : 4u * nw) + 4u * cap);
        stage_issue(T, d, blob, &bar);
        if (!UNIFORM && WSTAGE && nw) bulk_g2s(s_w, Wg, 4u * nw, &bar);
        if (cap) bulk_g2s(s_dg, B.diag + d.slot_base[ELEM_V], 4u * cap, &bar);
    }
    __syncthreads();  // the barrier object is initialised for everyone who waits on it below
    // while the copies fly: the new search direction of the patch's own vertices (the owner writes it back).  The slice
    // is 3 * cap floats = 3 * cap / 4 float4 (slot caps are multiples of 4, slices 48-byte aligned); three load pairs per thread
    // are issued before the first store -- written as a plain load / store loop the compiler keeps every load behind the
    // previous store (R comes through a struct member, it may alias P') and the block pays one DRAM latency per step
    // (profiles/r02M: 14.8 us per block, long-scoreboard 18 per issue).
    {
        const float*  dgp = B.diag + d.slot_base[ELEM_V];  // PRECOND: Z = R / diag stands where R stands (P' = Z + beta P)
        const float4* R4 = reinterpret_cast<const float4*>(B.R + g);
        const float4* O4 = reinterpret_cast<const float4*>(Pold + g);
        float4*       N4 = reinterpret_cast<float4*>(Pnew + g);
        float4*       s4 = reinterpret_cast<float4*>(s_p);
        const uint32_t n4 = 3u * cap / 4u, lim = 3u * nov;
        for (uint32_t j0 = threadIdx.x; j0 < n4; j0 += 3u * BT) {
            float4 r[3], o[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const uint32_t j = j0 + k * BT;
                r[k] = o[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j < n4) {
                    r[k] = __ldg(R4 + j);
                    if (!first) o[k] = __ldg(O4 + j);
                    if (PRECOND) {
                        const uint32_t e  = 4u * j;  // element e belongs to vertex e / 3
                        const float    d0 = __ldg(dgp + e / 3u), d1 = __ldg(dgp + (e + 1u) / 3u), d2 = __ldg(dgp + (e + 2u) / 3u),
                                    d3 = __ldg(dgp + (e + 3u) / 3u);
                        r[k].x = d0 != 0.f ? r[k].x / d0 : 0.f, r[k].y = d1 != 0.f ? r[k].y / d1 : 0.f;
                        r[k].z = d2 != 0.f ? r[k].z / d2 : 0.f, r[k].w = d3 != 0.f ? r[k].w / d3 : 0.f;
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const uint32_t j = j0 + k * BT;
                if (j < n4) {
                    float4 pn;  // rows past the owned vertices stay zero
                    pn.x = 4u * j < lim ? (first ? r[k].x : __fmaf_rn(beta, o[k].x, r[k].x)) : 0.f;
                    pn.y = 4u * j + 1u < lim ? (first ? r[k].y : __fmaf_rn(beta, o[k].y, r[k].y)) : 0.f;
                    pn.z = 4u * j + 2u < lim ? (first ? r[k].z : __fmaf_rn(beta, o[k].z, r[k].z)) : 0.f;
                    pn.w = 4u * j + 3u < lim ? (first ? r[k].w : __fmaf_rn(beta, o[k].w, r[k].w)) : 0.f;
                    N4[j] = pn;
                    // shared memory: the OWNED rows only.  The floats from 3 * nov on belong to the ribbon rows, which other
                    // threads fill below without a barrier in between -- zeros written here would race with them
                    if (4u * j + 3u < lim) {
                        s4[j] = pn;
                    } else {
                        if (4u * j < lim) s_p[4u * j] = pn.x;
                        if (4u * j + 1u < lim) s_p[4u * j + 1u] = pn.y;
                        if (4u * j + 2u < lim) s_p[4u * j + 2u] = pn.z;
                    }
                }
            }
        }
    }
    mbar_wait(&bar, 0);
    // ribbon vertices: the SAME expression on the rows of their owners (bit-identical to what the owner stores)
    for (uint32_t i = nov + threadIdx.x; i < nv; i += BT) {
        const uint32_t o = T.own[i - nov];
        const uint64_t s = 3ull * ((uint64_t)T.stash[o >> 16].slot_base[ELEM_V] + (o & 0xFFFFu));
        float          r0 = __ldg(B.R + s), r1 = __ldg(B.R + s + 1), r2 = __ldg(B.R + s + 2);
        if (PRECOND) {
            const float dr = __ldg(B.diag + s / 3ull);
            r0 = dr != 0.f ? r0 / dr : 0.f, r1 = dr != 0.f ? r1 / dr : 0.f, r2 = dr != 0.f ? r2 / dr : 0.f;
        }
        if (first) {
            s_p[3 * i] = r0, s_p[3 * i + 1] = r1, s_p[3 * i + 2] = r2;
        } else {
            s_p[3 * i]     = __fmaf_rn(beta, __ldg(Pold + s), r0);
            s_p[3 * i + 1] = __fmaf_rn(beta, __ldg(Pold + s + 1), r1);
            s_p[3 * i + 2] = __fmaf_rn(beta, __ldg(Pold + s + 2), r2);
        }
    }
    __syncthreads();
    double part = 0.0;
    for (uint32_t v = threadIdx.x; v < cap; v += BT) {
        float ox = 0.f, oy = 0.f, oz = 0.f;
        if (v < nov) {
            const uint32_t b = T.fo[v] & FAN_OFF_MASK, e = T.fo[v + 1] & FAN_OFF_MASK;
            float          x = 0.f, y = 0.f, z = 0.f;
            if (e - b == 6u && (b & 1u) == 0u) {
                // the regular vertex: ids two per LDS.32, no loop bookkeeping, every load in flight at once; the same
                // subtractions in the same order as the loop below
                const uint32_t* w32 = reinterpret_cast<const uint32_t*>(T.fv + b);
                const uint32_t  i01 = w32[0], i23 = w32[1], i45 = w32[2];
                float           w0 = dt, w1 = dt, w2 = dt, w3 = dt, w4 = dt, w5 = dt;
                if (!UNIFORM) {
                    float2 a, c, f;
                    if (WSTAGE) {
                        const float2* wp = reinterpret_cast<const float2*>(s_w + b);
                        a = wp[0], c = wp[1], f = wp[2];
                    } else {
                        const float2* wp = reinterpret_cast<const float2*>(Wg + b);
                        a = __ldg(wp), c = __ldg(wp + 1), f = __ldg(wp + 2);
                    }
                    w0 = a.x, w1 = a.y, w2 = c.x, w3 = c.y, w4 = f.x, w5 = f.y;
                }
                const float *q0 = s_p + 3u * (i01 & 0xFFFFu), *q1 = s_p + 3u * (i01 >> 16), *q2 = s_p + 3u * (i23 & 0xFFFFu),
                            *q3 = s_p + 3u * (i23 >> 16), *q4 = s_p + 3u * (i45 & 0xFFFFu), *q5 = s_p + 3u * (i45 >> 16);
                x -= w0 * q0[0], y -= w0 * q0[1], z -= w0 * q0[2];
                x -= w1 * q1[0], y -= w1 * q1[1], z -= w1 * q1[2];
                x -= w2 * q2[0], y -= w2 * q2[1], z -= w2 * q2[2];
                x -= w3 * q3[0], y -= w3 * q3[1], z -= w3 * q3[2];
                x -= w4 * q4[0], y -= w4 * q4[1], z -= w4 * q4[2];
                x -= w5 * q5[0], y -= w5 * q5[1], z -= w5 * q5[2];
            } else {
                for (uint32_t i = b; i < e; ++i) {
                    const float  w = UNIFORM ? dt : (WSTAGE ? s_w[i] : __ldg(Wg + i));
                    const float* q = s_p + 3u * T.fv[i];
                    x -= w * q[0], y -= w * q[1], z -= w * q[2];
                }
            }
            const float dg = s_dg[v], px = s_p[3 * v], py = s_p[3 * v + 1], pz = s_p[3 * v + 2];
            ox = x + dg * px, oy = y + dg * py, oz = z + dg * pz;
            part += (double)ox * px + (double)oy * py + (double)oz * pz;
        }
        B.S[g + 3 * v] = ox, B.S[g + 3 * v + 1] = oy, B.S[g + 3 * v + 2] = oz;
    }
    part = block_sum<BT>(part, s_red);
    if (publish_partial(part, B.partials, &B.state->ctr, &s_flag)) {
        const double a = sum_partials<BT>(B.partials, s_red);
        if (threadIdx.x == 0) {
            B.state->ctr    = 0;
            B.state->dot_sp = a;
            // alpha = delta_new / <S, P> in the solver's type (float)
            B.state->alpha = (float)B.state->delta_new / (float)a;
        }
    }
}

// --------------------------------------------------------------------------
// iteration, second half: X += alpha P', R -= alpha S, <R, R>; then the scalar step of the solver
// --------------------------------------------------------------------------
template <bool PRECOND>
__global__ void __launch_bounds__(MBT) k_mcf_update(uint64_t n4, McfBuffers B, const float4* __restrict__ P, float tol_abs,
                                                    float tol_rel, uint32_t max_iter)
{
    if (B.state->converged || B.state->iters >= max_iter) return;
    __shared__ double   s_red[MBT / 32];
    __shared__ uint32_t s_flag;
    const float         alpha = B.state->alpha;
    float4* __restrict__       X = reinterpret_cast<float4*>(B.X);
    float4* __restrict__       R = reinterpret_cast<float4*>(B.R);
    const float4* __restrict__ S = reinterpret_cast<const float4*>(B.S);
    double                     part = 0.0;
    // axpy(X, P, alpha, 1): X = alpha P + X;  axpy(R, S, -alpha, 1): R = -alpha S + R
    // PRECOND: delta = <R, Z> with Z = R / diag; dg = the diagonals of the (at most two) vertices the four elements belong to
    auto step = [&](float4 p, float4 s, float4& x, float4& r, float4 dg) {
        x.x = __fmaf_rn(alpha, p.x, x.x), x.y = __fmaf_rn(alpha, p.y, x.y), x.z = __fmaf_rn(alpha, p.z, x.z),
        x.w = __fmaf_rn(alpha, p.w, x.w);
        r.x = __fmaf_rn(-alpha, s.x, r.x), r.y = __fmaf_rn(-alpha, s.y, r.y), r.z = __fmaf_rn(-alpha, s.z, r.z),
        r.w = __fmaf_rn(-alpha, s.w, r.w);
        if (PRECOND) {
            const float zx = dg.x != 0.f ? r.x / dg.x : 0.f, zy = dg.y != 0.f ? r.y / dg.y : 0.f,
                        zz = dg.z != 0.f ? r.z / dg.z : 0.f, zw = dg.w != 0.f ? r.w / dg.w : 0.f;
            part += (double)r.x * zx + (double)r.y * zy + (double)r.z * zz + (double)r.w * zw;
        } else {
            part += (double)r.x * r.x + (double)r.y * r.y + (double)r.z * r.z + (double)r.w * r.w;
        }
    };
    auto diag4 = [&](uint64_t i) {  // element 4 i + c belongs to slot (4 i + c) / 3
        if (!PRECOND) return make_float4(0.f, 0.f, 0.f, 0.f);
        const uint64_t e = 4ull * i;
        return make_float4(__ldg(B.diag + e / 3ull), __ldg(B.diag + (e + 1ull) / 3ull), __ldg(B.diag + (e + 2ull) / 3ull),
                           __ldg(B.diag + (e + 3ull) / 3ull));
    };
    // two elements per trip, all eight loads before the first store (see k_mcf_matvec on why)
    const uint64_t stride = (uint64_t)gridDim.x * MBT;
    uint64_t       i      = blockIdx.x * (uint64_t)MBT + threadIdx.x;
    for (; i + stride < n4; i += 2 * stride) {
        const uint64_t k  = i + stride;
        const float4   p0 = __ldg(P + i), s0 = __ldg(S + i), p1 = __ldg(P + k), s1 = __ldg(S + k);
        const float4   d0 = diag4(i), d1 = diag4(k);
        float4         x0 = X[i], r0 = R[i], x1 = X[k], r1 = R[k];
        step(p0, s0, x0, r0, d0);
        step(p1, s1, x1, r1, d1);
        X[i] = x0, R[i] = r0, X[k] = x1, R[k] = r1;
    }
    if (i < n4) {
        const float4 p0 = __ldg(P + i), s0 = __ldg(S + i);
        const float4 d0 = diag4(i);
        float4       x0 = X[i], r0 = R[i];
        step(p0, s0, x0, r0, d0);
        X[i] = x0, R[i] = r0;
    }
    part = block_sum(part, s_red);
    if (publish_partial(part, B.partials + B.partials_split, &B.state->ctr, &s_flag)) {
        const double a = sum_partials(B.partials + B.partials_split, s_red);
        if (threadIdx.x == 0) {
            McfState* st  = B.state;
            st->ctr       = 0;
            st->delta_old = st->delta_new;
            st->delta_new = a;
            st->final_res = a;
            // IterativeSolver::is_converged(start, current), iterative_solver.h:57-63, in the solver's type
            const float cur = (float)a, init = (float)st->start;
            if (cur < tol_abs || cur / init < tol_rel) {
                st->converged = 1u;  // m_iter_taken is NOT incremented by the converging iteration
            } else {
                st->beta = cur / (float)st->delta_old;
                st->iters += 1u;
            }
        }
    }
}

template <typename K>
cudaError_t mcf_set_smem(K kernel, uint32_t bytes)
{
    if (bytes > 227u * 1024u) return cudaErrorInvalidValue;
    if (bytes > 48u * 1024u) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    return cudaSuccess;
}
uint32_t a16(uint32_t x)
{
    return (x + 15u) & ~15u;
}
// fan offsets, fan neighbours (+8 entries), owner table (+4), stash (+1): stage_alloc
uint32_t staged_smem(const KernelLimits& lim)
{
    return a16(2u * (lim.max_owned[ELEM_V] + 1u) + 16u) + a16(2u * lim.max_fan_total + 32u) + a16(4u * lim.max_not_owned[ELEM_V] + 32u) +
           16u * (lim.max_stash + 1u);
}

}  // namespace

#define RXM_MCF_FAIL(msg)             \
    do {                              \
        if (err) *err = msg;          \
        return cudaErrorInvalidValue; \
    } while (0)

uint32_t mcf_update_grid()
{
    return 148u * 8u;
}

namespace {
template <bool U, bool PC, bool WS = true>
cudaError_t launch_mv(int bt, const MeshView& mv, const McfBuffers& B, const float* pold, float* pnew, float dt, int first,
                      uint32_t smem, cudaStream_t stream)
{
    cudaError_t e = cudaSuccess;
#define RXM_MCF_MV(BTV)                                                                                              \
    do {                                                                                                             \
        e = mcf_set_smem(k_mcf_matvec<U, BTV, PC, WS>, smem);                                                        \
        if (e == cudaSuccess) k_mcf_matvec<U, BTV, PC, WS><<<mv.num_patches, BTV, smem, stream>>>(mv, B, pold, pnew, dt, first); \
    } while (0)
    if (bt == 128) RXM_MCF_MV(128);
    else if (bt == 192) RXM_MCF_MV(192);
    else if (bt == 288) RXM_MCF_MV(288);
    else RXM_MCF_MV(256);
#undef RXM_MCF_MV
    return e;
}
}  // namespace

cudaError_t launch_mcf_setup(const MeshView& mv, const KernelLimits& lim, const float* x0, const McfBuffers& B, bool uniform,
                             bool precond, float dt, cudaStream_t stream, const char** err)
{
    if (!mv.fans) RXM_MCF_FAIL("the mesh stores no one-ring fans (non-manifold or inconsistently oriented input)");
    const uint32_t capv = lim.max_owned[ELEM_V] + 4u;
    const uint32_t smem = staged_smem(lim) + a16(12u * std::max(lim.max_n[ELEM_V], capv) + 16u) + 64u;
    cudaError_t    e;
#define RXM_MCF_SETUP(U, PC)                                                                        \
    do {                                                                                            \
        e = mcf_set_smem(k_mcf_setup<U, PC>, smem);                                                 \
        if (e == cudaSuccess) k_mcf_setup<U, PC><<<mv.num_patches, MBT, smem, stream>>>(mv, x0, B, dt); \
    } while (0)
    if (uniform && precond) RXM_MCF_SETUP(true, true);
    else if (uniform) RXM_MCF_SETUP(true, false);
    else if (precond) RXM_MCF_SETUP(false, true);
    else RXM_MCF_SETUP(false, false);
#undef RXM_MCF_SETUP
    if (e != cudaSuccess) RXM_MCF_FAIL("patch needs more shared memory than 227 KB");
    count_launches(1);
    return cudaGetLastError();
}

cudaError_t launch_mcf_iteration(const MeshView& mv, const KernelLimits& lim, const McfBuffers& B, uint32_t it, bool uniform,
                                 bool precond, float dt, float tol_abs, float tol_rel, uint32_t max_iter, cudaStream_t stream,
                                 const char** err)
{
    const uint32_t capv = lim.max_owned[ELEM_V] + 4u;
    // cotangent weights: read where they are used (default) or staged in shared memory (RXM_MCF_WSTAGE=1).  10 M-face torus,
    // mat-vec alone under ncu: staged 95.1 us (32.6 KB per block, 6 resident), global 82.5 us (19.1 KB, 8 resident)
    static const bool w_staged = [] { const char* f = getenv("RXM_MCF_WSTAGE"); return f ? atoi(f) != 0 : false; }();
    const bool        ws       = uniform || w_staged;
    const uint32_t smem = staged_smem(lim) + (uniform || !ws ? 16u : a16(4u * (lim.max_fan_total + 8u))) + a16(4u * (capv + 4u)) +
                          a16(12u * std::max(lim.max_n[ELEM_V], capv) + 16u) + 64u;
    // block size: the fewest idle threads in the last round over a full patch's owned vertices (ties: the larger block)
    int      bt   = 256;
    uint32_t idle = ~0u;
    for (int c : {256, 192, 128}) {  // (288 = two rounds over a 561-vertex tile: RXM_MCF_BT, measured)
        const uint32_t nov = std::max(lim.max_owned[ELEM_V], 1u), w = (nov + c - 1) / c * c - nov;
        if (w < idle) idle = w, bt = c;
    }
    if (const char* f = getenv("RXM_MCF_BT")) bt = atoi(f);
    const float* pold  = B.P[it & 1u];
    float*       pnew  = B.P[(it + 1u) & 1u];
    const int    first = it == 0;
    cudaError_t  e;
    if (uniform && precond) e = launch_mv<true, true>(bt, mv, B, pold, pnew, dt, first, smem, stream);
    else if (uniform) e = launch_mv<true, false>(bt, mv, B, pold, pnew, dt, first, smem, stream);
    else if (precond && ws) e = launch_mv<false, true>(bt, mv, B, pold, pnew, dt, first, smem, stream);
    else if (precond) e = launch_mv<false, true, false>(bt, mv, B, pold, pnew, dt, first, smem, stream);
    else if (ws) e = launch_mv<false, false>(bt, mv, B, pold, pnew, dt, first, smem, stream);
    else e = launch_mv<false, false, false>(bt, mv, B, pold, pnew, dt, first, smem, stream);
    if (e != cudaSuccess) RXM_MCF_FAIL("patch needs more shared memory than 227 KB");
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const uint64_t n4 = 3ull * mv.num_slots[ELEM_V] / 4ull;  // slot caps are multiples of 4: 3 * slots floats = n4 float4
    if (precond)
        k_mcf_update<true><<<mcf_update_grid(), MBT, 0, stream>>>(n4, B, reinterpret_cast<const float4*>(pnew), tol_abs, tol_rel, max_iter);
    else
        k_mcf_update<false><<<mcf_update_grid(), MBT, 0, stream>>>(n4, B, reinterpret_cast<const float4*>(pnew), tol_abs, tol_rel, max_iter);
    count_launches(2);
    return cudaGetLastError();
}

}  // namespace rxm
