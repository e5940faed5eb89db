/*
 * rxmesh_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, CPU restatement of the reference algorithms on the hot path named
 * by BASELINE.json:north_star (static connectivity queries + vertex normals +
 * Laplacian / bilateral smoothing).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product path (rxmesh_b200/) never links, imports or calls it.
 *
 * Every function cites the reference file:line (relative to /root/reference)
 * whose behaviour it restates.  Nothing here is copied from the reference: the
 * reference builds std::unordered_map / vector<vector<>> structures, this file
 * works on flat arrays with an open-addressing table.
 *
 * Pinning (see oracle/README.md and tests/test_oracle.py):
 *   - normals:   pinned bit-for-bit against the reference's own
 *                apps/VertexNormal/vertex_normal_ref.h compiled UNMODIFIED into
 *                oracle/_ref/ (fixtures tests/golden/<mesh>.npz: vn_ref) and against
 *                the SURVEY.md 8(c) checksums.
 *   - queries:   pinned by the reference's known-answer tests (cube.obj counts
 *                test_for_each.cu:25-29, bunnyhead.obj 98 boundary vertices
 *                test_boundary.cu:27) and the ground-truth construction of
 *                tests/RXMesh_test/rxmesh_test.h:65-339 which this restates.
 *   - laplacian: PARITY UNPINNED (the reference app has no check);
 *                apps/Smoothing/manual.h:86-104 is the spec.
 *   - bilateral: weakly pinned (the only reference check is abs 1e-2 against
 *                OpenMesh, apps/Filtering/filtering_rxmesh.cuh:114-125; OpenMesh
 *                8.1 is not in the tree).
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define RXO_INVALID32 0xFFFFFFFFu

/* ------------------------------------------------------------------------ */
/* Edge numbering.                                                           */
/* rxmesh.cpp:589-611: scan faces f = 0..F-1, corners (v, (v+1)%3); the key  */
/* is (max, min) (util/util.h:410-416); an edge id is its order of first     */
/* appearance.  ev_out[2e] = max id, ev_out[2e+1] = min id, which is the      */
/* order the EV ground truth uses (rxmesh_test.h:207-211).                    */
/* fe_out[3f+j] = id of edge (fv[3f+j], fv[3f+(j+1)%3]) (rxmesh_test.h:22-38). */
/* ------------------------------------------------------------------------ */
typedef struct
{
    uint64_t* keys;
    uint32_t* vals;
    uint64_t  mask;
} rxo_table;

static uint64_t rxo_mix(uint64_t x)
{
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

/* returns number of edges; *num_vertices = max vertex id + 1 (rxmesh.cpp:602,640) */
uint32_t rxo_build_edges(const uint32_t* fv,
                         uint32_t        nf,
                         uint32_t*       ev_out, /* capacity 2*3*nf */
                         uint32_t*       fe_out, /* 3*nf */
                         uint32_t*       num_vertices)
{
    uint64_t cap = 16;
    while (cap < (uint64_t)nf * 6u + 8u)
        cap <<= 1;
    rxo_table t;
    t.keys = (uint64_t*)malloc(cap * sizeof(uint64_t));
    t.vals = (uint32_t*)malloc(cap * sizeof(uint32_t));
    t.mask = cap - 1;
    memset(t.keys, 0xFF, cap * sizeof(uint64_t));

    uint32_t ne = 0, nv = 0;
    for (uint32_t f = 0; f < nf; ++f) {
        for (uint32_t j = 0; j < 3; ++j) {
            uint32_t v0 = fv[3 * (uint64_t)f + j];
            uint32_t v1 = fv[3 * (uint64_t)f + (j + 1) % 3];
            if (v0 > nv) nv = v0;
            if (v1 > nv) nv = v1;
            uint32_t hi  = v0 > v1 ? v0 : v1;
            uint32_t lo  = v0 > v1 ? v1 : v0;
            uint64_t key = ((uint64_t)hi << 32) | lo;
            uint64_t h   = rxo_mix(key) & t.mask;
            while (t.keys[h] != key && t.keys[h] != ~0ULL)
                h = (h + 1) & t.mask;
            if (t.keys[h] == ~0ULL) {
                t.keys[h]          = key;
                t.vals[h]          = ne;
                ev_out[2 * (uint64_t)ne]     = hi;
                ev_out[2 * (uint64_t)ne + 1] = lo;
                ++ne;
            }
            fe_out[3 * (uint64_t)f + j] = t.vals[h];
        }
    }
    free(t.keys);
    free(t.vals);
    *num_vertices = nf ? nv + 1 : 0;
    return ne;
}

/* ------------------------------------------------------------------------ */
/* Generic CSR transpose helper: rows of fixed degree `deg` with column ids    */
/* in `rows`; pushes row id r onto column rows[r*deg+k] in (r, k) scan order,  */
/* which is the push_back order of the reference ground truth.                */
/* ------------------------------------------------------------------------ */
static void rxo_transpose(const uint32_t* rows,
                          uint32_t        nrows,
                          uint32_t        deg,
                          uint32_t        ncols,
                          uint32_t*       off, /* ncols+1 */
                          uint32_t*       val) /* nrows*deg */
{
    memset(off, 0, ((size_t)ncols + 1) * sizeof(uint32_t));
    for (uint64_t i = 0; i < (uint64_t)nrows * deg; ++i)
        off[rows[i] + 1]++;
    for (uint32_t c = 0; c < ncols; ++c)
        off[c + 1] += off[c];
    uint32_t* cur = (uint32_t*)malloc(((size_t)ncols + 1) * sizeof(uint32_t));
    memcpy(cur, off, ((size_t)ncols + 1) * sizeof(uint32_t));
    for (uint32_t r = 0; r < nrows; ++r)
        for (uint32_t k = 0; k < deg; ++k)
            val[cur[rows[(uint64_t)r * deg + k]]++] = r;
    free(cur);
}

/* VE ground truth: rxmesh_test.h:137-160 (v_e[first].push(e); v_e[second].push(e)). */
void rxo_query_ve(const uint32_t* ev, uint32_t ne, uint32_t nv, uint32_t* off, uint32_t* val)
{
    rxo_transpose(ev, ne, 2, nv, off, val);
}

/* VV ground truth: rxmesh_test.h:65-83 (v_v[first].push(second) and vice versa). */
void rxo_query_vv(const uint32_t* ev, uint32_t ne, uint32_t nv, uint32_t* off, uint32_t* val)
{
    rxo_transpose(ev, ne, 2, nv, off, val);
    /* replace edge id by the other endpoint; entries of column v hold edges
     * incident to v. */
    for (uint32_t v = 0; v < nv; ++v)
        for (uint32_t i = off[v]; i < off[v + 1]; ++i) {
            uint32_t e = val[i];
            val[i]     = (ev[2 * (uint64_t)e] == v) ? ev[2 * (uint64_t)e + 1] : ev[2 * (uint64_t)e];
        }
}

/* VF ground truth: rxmesh_test.h:165-193. */
void rxo_query_vf(const uint32_t* fv, uint32_t nf, uint32_t nv, uint32_t* off, uint32_t* val)
{
    rxo_transpose(fv, nf, 3, nv, off, val);
}

/* EF ground truth: rxmesh_test.h:228-252 (faces in order, edges of FE). */
void rxo_query_ef(const uint32_t* fe, uint32_t nf, uint32_t ne, uint32_t* off, uint32_t* val)
{
    rxo_transpose(fe, nf, 3, ne, off, val);
}

/* FF ground truth: rxmesh_test.h:292-339: for every edge, every unordered pair
 * of incident faces (f0,f1) contributes f1 to f_f[f0] and f0 to f_f[f1]
 * (multiplicity kept: two faces sharing two edges list each other twice).
 * Returns nnz; val must have capacity for it (call with val==NULL to count). */
uint64_t rxo_query_ff(const uint32_t* fe, uint32_t nf, uint32_t ne, uint32_t* off, uint32_t* val)
{
    uint32_t* eoff = (uint32_t*)malloc(((size_t)ne + 1) * sizeof(uint32_t));
    uint32_t* eval = (uint32_t*)malloc((size_t)nf * 3 * sizeof(uint32_t) + 4);
    rxo_transpose(fe, nf, 3, ne, eoff, eval);
    memset(off, 0, ((size_t)nf + 1) * sizeof(uint32_t));
    for (uint32_t e = 0; e < ne; ++e) {
        uint32_t k = eoff[e + 1] - eoff[e];
        for (uint32_t i = eoff[e]; i < eoff[e + 1]; ++i)
            off[eval[i] + 1] += k - 1;
    }
    for (uint32_t f = 0; f < nf; ++f)
        off[f + 1] += off[f];
    uint64_t nnz = off[nf];
    if (val) {
        uint32_t* cur = (uint32_t*)malloc(((size_t)nf + 1) * sizeof(uint32_t));
        memcpy(cur, off, ((size_t)nf + 1) * sizeof(uint32_t));
        for (uint32_t e = 0; e < ne; ++e)
            for (uint32_t i = eoff[e]; i + 1 < eoff[e + 1]; ++i)
                for (uint32_t j = i + 1; j < eoff[e + 1]; ++j) {
                    uint32_t f0 = eval[i], f1 = eval[j];
                    val[cur[f0]++] = f1;
                    val[cur[f1]++] = f0;
                }
        free(cur);
    }
    free(eoff);
    free(eval);
    return nnz;
}

/* Boundary vertices: kernels/boundary.cuh:11-44 -- an edge with exactly one
 * incident face marks both its vertices.  Returns the count, flags[v] in {0,1}. */
uint32_t rxo_boundary_vertices(const uint32_t* ev, const uint32_t* fe, uint32_t nf, uint32_t ne,
                               uint32_t nv, uint8_t* flags)
{
    uint32_t* cnt = (uint32_t*)calloc(ne ? ne : 1, sizeof(uint32_t));
    for (uint64_t i = 0; i < (uint64_t)nf * 3; ++i)
        cnt[fe[i]]++;
    memset(flags, 0, nv);
    for (uint32_t e = 0; e < ne; ++e)
        if (cnt[e] == 1) {
            flags[ev[2 * (uint64_t)e]]     = 1;
            flags[ev[2 * (uint64_t)e + 1]] = 1;
        }
    uint32_t n = 0;
    for (uint32_t v = 0; v < nv; ++v)
        n += flags[v];
    free(cnt);
    return n;
}

/* Input statistics of rxmesh.cpp:560-650: max valence, max faces per edge,
 * max adjacent faces per face, closed, edge-manifold. out[5]. */
void rxo_input_stats(const uint32_t* ev, const uint32_t* fe, uint32_t nf, uint32_t ne, uint32_t nv,
                     uint32_t* out)
{
    uint32_t* cnt = (uint32_t*)calloc(ne ? ne : 1, sizeof(uint32_t));
    uint32_t* val = (uint32_t*)calloc(nv ? nv : 1, sizeof(uint32_t));
    for (uint64_t i = 0; i < (uint64_t)nf * 3; ++i)
        cnt[fe[i]]++;
    uint32_t max_val = 0, max_ef = 0, closed = 1, manifold = 1, max_ff = 0;
    for (uint32_t e = 0; e < ne; ++e) {
        if (cnt[e] > max_ef) max_ef = cnt[e];
        if (cnt[e] < 2) closed = 0;
        if (cnt[e] > 2) manifold = 0;
        if (++val[ev[2 * (uint64_t)e]] > max_val) max_val = val[ev[2 * (uint64_t)e]];
        if (++val[ev[2 * (uint64_t)e + 1]] > max_val) max_val = val[ev[2 * (uint64_t)e + 1]];
    }
    for (uint32_t f = 0; f < nf; ++f) {
        uint32_t k = 0;
        for (uint32_t j = 0; j < 3; ++j)
            k += cnt[fe[3 * (uint64_t)f + j]] - 1;
        if (k > max_ff) max_ff = k;
    }
    out[0] = max_val;
    out[1] = max_ef;
    out[2] = max_ff;
    out[3] = closed;
    out[4] = manifold;
    free(cnt);
    free(val);
}

/* ------------------------------------------------------------------------ */
/* Vertex normals, Max 1999 weights.                                         */
/* apps/VertexNormal/vertex_normal_ref.h:5-85: serial loop over faces in      */
/* input order; n = (v1-v0) x (v2-v0); squared edge lengths l0=|v0v1|^2,      */
/* l1=|v1v2|^2, l2=|v2v0|^2; corner i receives n / (l[i] + l[(i+2)%3]).       */
/* fp32 arithmetic in exactly the reference's operation order so the result   */
/* is bit-identical to oracle/_ref (compile without FMA contraction).         */
/* ------------------------------------------------------------------------ */
static float rxo_l2sq_f(const float* a, const float* b)
{
    float x0 = a[0] - b[0];
    float x1 = a[1] - b[1];
    float x2 = a[2] - b[2];
    return x0 * x0 + x1 * x1 + x2 * x2;
}

void rxo_vertex_normals_f32(const uint32_t* fv, uint32_t nf, const float* x, uint32_t nv, float* n)
{
    memset(n, 0, (size_t)nv * 3 * sizeof(float));
    for (uint32_t f = 0; f < nf; ++f) {
        const uint32_t* v  = fv + 3 * (uint64_t)f;
        const float *   p0 = x + 3 * (uint64_t)v[0], *p1 = x + 3 * (uint64_t)v[1],
                    *p2 = x + 3 * (uint64_t)v[2];
        float a0 = p1[0] - p0[0], a1 = p1[1] - p0[1], a2 = p1[2] - p0[2];
        float b0 = p2[0] - p0[0], b1 = p2[1] - p0[1], b2 = p2[2] - p0[2];
        float fn[3];
        fn[0] = a1 * b2 - a2 * b1;
        fn[1] = a2 * b0 - a0 * b2;
        fn[2] = a0 * b1 - a1 * b0;
        float l[3];
        l[0] = rxo_l2sq_f(p0, p1);
        l[1] = rxo_l2sq_f(p1, p2);
        l[2] = rxo_l2sq_f(p2, p0);
        for (uint32_t i = 0; i < 3; ++i) {
            uint32_t k = (i + 2) % 3;
            float*   o = n + 3 * (uint64_t)v[i];
            for (uint32_t c = 0; c < 3; ++c)
                o[c] += fn[c] / (l[i] + l[k]);
        }
    }
}

/* same formula in float64 from the same fp32 inputs: the tolerance yardstick
 * (SURVEY.md section 7 "hard parts": -use_fast_math in the reference build). */
void rxo_vertex_normals_f64(const uint32_t* fv, uint32_t nf, const float* x, uint32_t nv, double* n)
{
    memset(n, 0, (size_t)nv * 3 * sizeof(double));
    for (uint32_t f = 0; f < nf; ++f) {
        const uint32_t* v = fv + 3 * (uint64_t)f;
        double          p[3][3];
        for (int i = 0; i < 3; ++i)
            for (int c = 0; c < 3; ++c)
                p[i][c] = x[3 * (uint64_t)v[i] + c];
        double a[3], b[3], fn[3], l[3];
        for (int c = 0; c < 3; ++c) {
            a[c] = p[1][c] - p[0][c];
            b[c] = p[2][c] - p[0][c];
        }
        fn[0] = a[1] * b[2] - a[2] * b[1];
        fn[1] = a[2] * b[0] - a[0] * b[2];
        fn[2] = a[0] * b[1] - a[1] * b[0];
        for (int i = 0; i < 3; ++i) {
            int j = (i + 1) % 3;
            l[i]  = 0;
            for (int c = 0; c < 3; ++c)
                l[i] += (p[i][c] - p[j][c]) * (p[i][c] - p[j][c]);
        }
        for (int i = 0; i < 3; ++i) {
            int k = (i + 2) % 3;
            for (int c = 0; c < 3; ++c)
                n[3 * (uint64_t)v[i] + c] += fn[c] / (l[i] + l[k]);
        }
    }
}

/* Filtering's vertex normal: apps/Filtering/filtering_rxmesh_kernel.cuh:15-46
 * -- sum over incident faces of the NORMALISED face normal (not normalised at
 * the end; bilateral_filtering normalises when it reads it, :452). float64
 * accumulate from fp32 inputs when out64 != NULL, fp32 otherwise. */
void rxo_vertex_normals_unit_faces(const uint32_t* fv, uint32_t nf, const float* x, uint32_t nv,
                                   float* out32, double* out64)
{
    if (out32) memset(out32, 0, (size_t)nv * 3 * sizeof(float));
    if (out64) memset(out64, 0, (size_t)nv * 3 * sizeof(double));
    for (uint32_t f = 0; f < nf; ++f) {
        const uint32_t* v = fv + 3 * (uint64_t)f;
        if (out64) {
            double a[3], b[3], fn[3];
            for (int c = 0; c < 3; ++c) {
                a[c] = (double)x[3 * (uint64_t)v[1] + c] - x[3 * (uint64_t)v[0] + c];
                b[c] = (double)x[3 * (uint64_t)v[2] + c] - x[3 * (uint64_t)v[0] + c];
            }
            fn[0]    = a[1] * b[2] - a[2] * b[1];
            fn[1]    = a[2] * b[0] - a[0] * b[2];
            fn[2]    = a[0] * b[1] - a[1] * b[0];
            double s = 1.0 / sqrt(fn[0] * fn[0] + fn[1] * fn[1] + fn[2] * fn[2]);
            for (int i = 0; i < 3; ++i)
                for (int c = 0; c < 3; ++c)
                    out64[3 * (uint64_t)v[i] + c] += fn[c] * s;
        }
        if (out32) {
            float a[3], b[3], fn[3];
            for (int c = 0; c < 3; ++c) {
                a[c] = x[3 * (uint64_t)v[1] + c] - x[3 * (uint64_t)v[0] + c];
                b[c] = x[3 * (uint64_t)v[2] + c] - x[3 * (uint64_t)v[0] + c];
            }
            fn[0]   = a[1] * b[2] - a[2] * b[1];
            fn[1]   = a[2] * b[0] - a[0] * b[2];
            fn[2]   = a[0] * b[1] - a[1] * b[0];
            float s = 1.0f / sqrtf(fn[0] * fn[0] + fn[1] * fn[1] + fn[2] * fn[2]);
            for (int i = 0; i < 3; ++i)
                for (int c = 0; c < 3; ++c)
                    out32[3 * (uint64_t)v[i] + c] += fn[c] * s;
        }
    }
}

/* ------------------------------------------------------------------------ */
/* Laplacian smoothing ("manual" path).                                      */
/* apps/Smoothing/manual.h:86-104: grad(v,i) = sum_u 2*(x(v,i) - x(u,i)) in   */
/* fp32 (grad starts at 0, one += per neighbour); then x(v,i) -= lr*grad(v,i) */
/* with lr a double (manual.h:37), i.e. the update is evaluated in double and */
/* rounded to fp32.  All gradients are computed before any position moves     */
/* (two kernels), i.e. a Jacobi step.  vv_off/vv_val = VV adjacency.          */
/* ------------------------------------------------------------------------ */
void rxo_laplacian_step_f32(const uint32_t* vv_off, const uint32_t* vv_val, uint32_t nv,
                            const float* x_in, float* x_out, double lr)
{
    for (uint32_t v = 0; v < nv; ++v) {
        float g[3] = {0.f, 0.f, 0.f};
        for (uint32_t i = vv_off[v]; i < vv_off[v + 1]; ++i) {
            uint32_t u = vv_val[i];
            for (int c = 0; c < 3; ++c)
                g[c] += 2 * (x_in[3 * (uint64_t)v + c] - x_in[3 * (uint64_t)u + c]);
        }
        for (int c = 0; c < 3; ++c)
            x_out[3 * (uint64_t)v + c] = (float)((double)x_in[3 * (uint64_t)v + c] - lr * (double)g[c]);
    }
}

void rxo_laplacian_step_f64(const uint32_t* vv_off, const uint32_t* vv_val, uint32_t nv,
                            const double* x_in, double* x_out, double lr)
{
    for (uint32_t v = 0; v < nv; ++v) {
        double g[3] = {0, 0, 0};
        for (uint32_t i = vv_off[v]; i < vv_off[v + 1]; ++i) {
            uint32_t u = vv_val[i];
            for (int c = 0; c < 3; ++c)
                g[c] += 2 * (x_in[3 * (uint64_t)v + c] - x_in[3 * (uint64_t)u + c]);
        }
        for (int c = 0; c < 3; ++c)
            x_out[3 * (uint64_t)v + c] = x_in[3 * (uint64_t)v + c] - lr * g[c];
    }
}

/* ------------------------------------------------------------------------ */
/* Bilateral mesh denoising, one iteration.                                  */
/* Restates apps/Filtering/filtering_rxmesh_kernel.cuh:426-548 (+52-85 and    */
/* filtering_util.h:8-59), which is the GPU side of the app; the OpenMesh side */
/* (filtering_openmesh.h:73-207) computes the same quantities:                */
/*   n      = normalise(sum of unit face normals)                             */
/*   sc2    = min squared distance to a 1-ring neighbour                      */
/*   radius = 4*sc2 on squared distances (== 2*sigma_c on distances, :53)     */
/*   N      = {v} + breadth-first closure of neighbours with |q-v|^2<=radius, */
/*            visited in discovery order, expanded only from members of N     */
/*   ss2    = var(|<q-v,n>|) over N (population variance), +1e-20 if <1e-20   */
/*   v'     = v + n * sum(wc*ws*h)/sum(wc*ws), wc=exp(-t^2/(2 sc2)),          */
/*            ws=exp(-h^2/(2 ss2)), t=|q-v|, h=<q-v,n>                        */
/* max_nbrs mirrors maxVVSize (filtering_rxmesh.cuh: 80): exceeding it is an  */
/* assert in the reference; here the count is reported and the list clipped.  */
/* `normals` = output of rxo_vertex_normals_unit_faces (unnormalised sum).    */
/* Computed in float64 from fp32 inputs when use_f64 != 0, else in fp32.      */
/* use_f64 == 2: the MEMBERSHIP decisions (squared distances, sigma_c^2, the  */
/* <= 4 sigma_c^2 test) are made in fp32 exactly as a fp32 implementation of   */
/* glm::distance2 makes them -- differences in fp32, then                      */
/* fmaf(dz, dz, fmaf(dy, dy, dx * dx)) -- so that the neighbourhoods are the   */
/* same SETS as the GPU's; everything after membership stays in float64.       */
/* Returns the maximum neighbourhood size seen.                               */
/* ------------------------------------------------------------------------ */
static double rxo_dist2_f32(const float* a, const float* b)
{
    const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return (double)fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

static uint32_t rxo_bilateral_range(const uint32_t* vv_off, const uint32_t* vv_val, uint32_t v_begin, uint32_t v_end,
                                    const float* x, const double* normals, float* x_out,
                                    uint32_t max_nbrs, int use_f64)
{
    uint32_t* list = (uint32_t*)malloc((size_t)(max_nbrs + 1) * sizeof(uint32_t));
    uint32_t  worst = 0;
    for (uint32_t v = v_begin; v < v_end; ++v) {
        double p[3], n[3];
        for (int c = 0; c < 3; ++c) {
            p[c] = x[3 * (uint64_t)v + c];
            n[c] = normals[3 * (uint64_t)v + c];
        }
        double nl = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        for (int c = 0; c < 3; ++c)
            n[c] = use_f64 ? n[c] / nl : (double)((float)n[c] / (float)nl);
        double sc2 = 1e10;
        for (uint32_t i = vv_off[v]; i < vv_off[v + 1]; ++i) {
            uint32_t u = vv_val[i];
            double   d = 0;
            for (int c = 0; c < 3; ++c) {
                double t = (double)x[3 * (uint64_t)u + c] - p[c];
                d += t * t;
            }
            if (!use_f64) d = (float)d;
            if (use_f64 == 2) d = rxo_dist2_f32(x + 3 * (uint64_t)u, x + 3 * (uint64_t)v);
            if (d < sc2) sc2 = d;
        }
        double   radius = 4.0 * sc2;
        uint32_t cnt    = 0;
        list[cnt++]     = v;
        for (uint32_t head = 0; head < cnt; ++head) {
            uint32_t w = list[head];
            for (uint32_t i = vv_off[w]; i < vv_off[w + 1]; ++i) {
                uint32_t u = vv_val[i];
                if (u == v) continue;
                int dup = 0;
                for (uint32_t k = 0; k < cnt; ++k)
                    if (list[k] == u) {
                        dup = 1;
                        break;
                    }
                if (dup) continue;
                double d = 0;
                for (int c = 0; c < 3; ++c) {
                    double t = (double)x[3 * (uint64_t)u + c] - p[c];
                    d += t * t;
                }
                if (!use_f64) d = (float)d;
                if (use_f64 == 2) d = rxo_dist2_f32(x + 3 * (uint64_t)u, x + 3 * (uint64_t)v);
                if (d <= radius) {
                    if (cnt < max_nbrs) list[cnt] = u;
                    cnt++;
                    if (cnt > max_nbrs) cnt = max_nbrs, worst = max_nbrs + 1;
                }
            }
        }
        if (cnt > worst) worst = cnt;
        double sum = 0, sum_sq = 0;
        for (uint32_t k = 0; k < cnt; ++k) {
            double h = 0;
            for (int c = 0; c < 3; ++c)
                h += ((double)x[3 * (uint64_t)list[k] + c] - p[c]) * n[c];
            h = fabs(h);
            sum += h;
            sum_sq += h * h;
        }
        double cc  = (double)cnt;
        double ss2 = sum_sq / cc - (sum * sum) / (cc * cc);
        if (ss2 < 1.0e-20) ss2 += 1.0e-20;
        double num = 0, den = 0;
        for (uint32_t k = 0; k < cnt; ++k) {
            double t2 = 0, h = 0;
            for (int c = 0; c < 3; ++c) {
                double q = (double)x[3 * (uint64_t)list[k] + c] - p[c];
                t2 += q * q;
                h += q * n[c];
            }
            double wc = exp(-0.5 * t2 / sc2);
            double ws = exp(-0.5 * h * h / ss2);
            num += wc * ws * h;
            den += wc * ws;
        }
        for (int c = 0; c < 3; ++c)
            x_out[3 * (uint64_t)v + c] = (float)(p[c] + n[c] * (num / den));
    }
    free(list);
    return worst;
}

uint32_t rxo_bilateral_step(const uint32_t* vv_off, const uint32_t* vv_val, uint32_t nv,
                            const float* x, const double* normals, float* x_out,
                            uint32_t max_nbrs, int use_f64)
{
    return rxo_bilateral_range(vv_off, vv_val, 0, nv, x, normals, x_out, max_nbrs, use_f64);
}

/* the same iteration over `threads` OpenMP threads, vertices split statically as the reference's CPU side of the app     */
/* does (apps/Filtering/filtering_openmesh.h:112-116: omp parallel for schedule(static)); every vertex is independent, so  */
/* the result is the serial one bit for bit.  Used as the CPU baseline of bench_configs.py.                               */
uint32_t rxo_bilateral_step_mt(const uint32_t* vv_off, const uint32_t* vv_val, uint32_t nv,
                               const float* x, const double* normals, float* x_out,
                               uint32_t max_nbrs, int use_f64, int threads)
{
    uint32_t worst = 0;
    if (threads < 1) threads = 1;
#pragma omp parallel num_threads(threads) reduction(max : worst)
    {
        const int      t = omp_get_thread_num(), T = omp_get_num_threads();
        const uint32_t b = (uint32_t)((uint64_t)nv * t / T), e = (uint32_t)((uint64_t)nv * (t + 1) / T);
        const uint32_t w = rxo_bilateral_range(vv_off, vv_val, b, e, x, normals, x_out, max_nbrs, use_f64);
        if (w > worst) worst = w;
    }
    return worst;
}

/* manual smoothing step (rxo_laplacian_step_f32) over `threads` OpenMP threads; the reference app has no CPU side, the     */
/* split over vertices is the obvious one.  Bit-identical to the serial step.                                             */
void rxo_laplacian_step_f32_mt(const uint32_t* vv_off, const uint32_t* vv_val, uint32_t nv,
                               const float* x_in, float* x_out, double lr, int threads)
{
    if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int64_t v = 0; v < (int64_t)nv; ++v) {
        float g[3] = {0.f, 0.f, 0.f};
        for (uint32_t i = vv_off[v]; i < vv_off[v + 1]; ++i) {
            uint32_t u = vv_val[i];
            for (int c = 0; c < 3; ++c)
                g[c] += 2 * (x_in[3 * (uint64_t)v + c] - x_in[3 * (uint64_t)u + c]);
        }
        for (int c = 0; c < 3; ++c)
            x_out[3 * (uint64_t)v + c] = (float)((double)x_in[3 * (uint64_t)v + c] - lr * (double)g[c]);
    }
}

/* ------------------------------------------------------------------------ */
/* All-cores variants for the CPU baseline ("an OpenMP-over-faces  */
/* variant with per-thread accumulation"; the reference's own vertex_normal_ref  */
/* is serial, its Filtering CPU path uses `omp parallel for schedule(static)`     */
/* with omp_get_max_threads() threads, filtering.cu:45-46).  Deterministic: every */
/* thread accumulates a contiguous face range into its own array, the arrays are  */
/* summed in thread order.  scratch: threads * nv * 3 floats, caller-owned.        */
/* ------------------------------------------------------------------------ */
#ifdef _OPENMP
#include <omp.h>
#endif
int rxo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void rxo_vertex_normals_f32_mt(const uint32_t* fv, uint32_t nf, const float* x, uint32_t nv, float* n, float* scratch,
                               int threads)
{
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    uint32_t lo[256], hi[256]; /* vertex span [lo, hi) touched by each thread's face range */
#pragma omp parallel num_threads(threads)
    {
#ifdef _OPENMP
        const int t = omp_get_thread_num(), nt = omp_get_num_threads();
#else
        const int t = 0, nt = 1;
#endif
        float*         acc = scratch + (size_t)t * nv * 3;
        const uint32_t f0 = (uint32_t)((uint64_t)nf * t / nt), f1 = (uint32_t)((uint64_t)nf * (t + 1) / nt);
        uint32_t       vmin = nv, vmax = 0;
        for (uint64_t i = 3 * (uint64_t)f0; i < 3 * (uint64_t)f1; ++i) {
            if (fv[i] < vmin) vmin = fv[i];
            if (fv[i] + 1 > vmax) vmax = fv[i] + 1;
        }
        if (vmin > vmax) vmin = vmax = 0;
        lo[t] = vmin, hi[t] = vmax;
        memset(acc + 3 * (size_t)vmin, 0, (size_t)(vmax - vmin) * 3 * sizeof(float));
        for (uint32_t f = f0; f < f1; ++f) {
            const uint32_t* v  = fv + 3 * (uint64_t)f;
            const float *   p0 = x + 3 * (uint64_t)v[0], *p1 = x + 3 * (uint64_t)v[1], *p2 = x + 3 * (uint64_t)v[2];
            float a0 = p1[0] - p0[0], a1 = p1[1] - p0[1], a2 = p1[2] - p0[2];
            float b0 = p2[0] - p0[0], b1 = p2[1] - p0[1], b2 = p2[2] - p0[2];
            float fn[3] = {a1 * b2 - a2 * b1, a2 * b0 - a0 * b2, a0 * b1 - a1 * b0};
            float l[3]  = {rxo_l2sq_f(p0, p1), rxo_l2sq_f(p1, p2), rxo_l2sq_f(p2, p0)};
            for (uint32_t i = 0; i < 3; ++i) {
                float* o = acc + 3 * (uint64_t)v[i];
                for (uint32_t c = 0; c < 3; ++c)
                    o[c] += fn[c] / (l[i] + l[(i + 2) % 3]);
            }
        }
#pragma omp barrier
        const uint32_t v0 = (uint32_t)((uint64_t)nv * t / nt), v1 = (uint32_t)((uint64_t)nv * (t + 1) / nt);
        memset(n + 3 * (size_t)v0, 0, (size_t)(v1 - v0) * 3 * sizeof(float));
        for (int k = 0; k < nt; ++k) { /* thread order: deterministic */
            const uint32_t a = lo[k] > v0 ? lo[k] : v0, e = hi[k] < v1 ? hi[k] : v1;
            const float*   src = scratch + (size_t)k * nv * 3;
            for (uint64_t i = 3 * (uint64_t)a; i < 3 * (uint64_t)e && a < e; ++i)
                n[i] += src[i];
        }
    }
}
void rxo_consume_sum_f32_mt(const uint32_t* off, const uint32_t* val, uint32_t n_src, const float* in, float* out, int threads)
{
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int64_t s = 0; s < (int64_t)n_src; ++s) {
        float a = 0.f;
        for (uint32_t i = off[s]; i < off[s + 1]; ++i)
            a += in[val[i]];
        out[s] = a;
    }
}

/* consume-variant checksums used by the roofline kernels: out[v] = sum of
 * in[u] over the VV / VF lists, accumulated in float64 (order-free yardstick). */
void rxo_consume_sum(const uint32_t* off, const uint32_t* val, uint32_t n_src, const float* in,
                     double* out)
{
    for (uint32_t s = 0; s < n_src; ++s) {
        double a = 0;
        for (uint32_t i = off[s]; i < off[s + 1]; ++i)
            a += in[val[i]];
        out[s] = a;
    }
}

/* ------------------------------------------------------------------------ */
/* Oriented one-rings of a CLOSED, consistently oriented manifold mesh: for    */
/* vertex v the cyclic sequence of neighbours such that consecutive entries    */
/* (a, b) span a face (v, a, b); the contract of the reference's oriented VV    */
/* (kernels/rxmesh_queries.cuh:375-499; start and winding unspecified).         */
/* off[nv+1], val[3*nf]. Returns 0 on success, 1 if some ring is not a cycle.   */
/* ------------------------------------------------------------------------ */
int rxo_oriented_rings(const uint32_t* fv, uint32_t nf, uint32_t nv, uint32_t* off, uint32_t* val)
{
    /* next[(v, a)] = b for every face rotation (v, a, b): store per vertex as pairs */
    uint32_t* cnt = (uint32_t*)calloc((size_t)nv + 1, sizeof(uint32_t));
    for (uint64_t i = 0; i < (uint64_t)nf * 3; ++i)
        cnt[fv[i] + 1]++;
    off[0] = 0;
    for (uint32_t v = 0; v < nv; ++v)
        off[v + 1] = off[v] + cnt[v + 1];
    uint32_t* a = (uint32_t*)malloc((size_t)nf * 3 * sizeof(uint32_t));
    uint32_t* b = (uint32_t*)malloc((size_t)nf * 3 * sizeof(uint32_t));
    uint32_t* cur = (uint32_t*)malloc(((size_t)nv + 1) * sizeof(uint32_t));
    memcpy(cur, off, ((size_t)nv + 1) * sizeof(uint32_t));
    for (uint32_t f = 0; f < nf; ++f)
        for (int j = 0; j < 3; ++j) {
            uint32_t v = fv[3 * (uint64_t)f + j];
            a[cur[v]]   = fv[3 * (uint64_t)f + (j + 1) % 3];
            b[cur[v]++] = fv[3 * (uint64_t)f + (j + 2) % 3];
        }
    int bad = 0;
    for (uint32_t v = 0; v < nv && !bad; ++v) {
        uint32_t k = off[v + 1] - off[v], base = off[v];
        uint32_t curl = 0;
        for (uint32_t i = 0; i < k; ++i) {
            val[base + i] = a[base + curl];
            uint32_t nxt = b[base + curl], found = k;
            for (uint32_t j = 0; j < k; ++j)
                if (a[base + j] == nxt) found = j;
            if (found == k) { bad = 1; break; }
            curl = found;
        }
        if (!bad && curl != 0) bad = 1; /* must close the cycle */
    }
    free(cnt); free(a); free(b); free(cur);
    return bad;
}

static double rxo_tri_area(const double* p, const double* q, const double* r)
{
    double u[3] = {q[0] - p[0], q[1] - p[1], q[2] - p[2]}, w[3] = {r[0] - p[0], r[1] - p[1], r[2] - p[2]};
    double c[3] = {u[1] * w[2] - u[2] * w[1], u[2] * w[0] - u[0] * w[2], u[0] * w[1] - u[1] * w[0]};
    return 0.5 * sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
}
static double rxo_clamp_cot(double v)
{
    return v < -19.1 ? -19.1 : (v > 19.1 ? 19.1 : v);
}
/* geometry_util.cuh:120-172 */
static double rxo_partial_voronoi(const double* p, const double* q, const double* r)
{
    double pq[3], qr[3], pr[3];
    for (int c = 0; c < 3; ++c) pq[c] = q[c] - p[c], qr[c] = r[c] - q[c], pr[c] = r[c] - p[c];
    double area = rxo_tri_area(p, q, r);
    if (area <= 1.17549435e-38) return -1;
    double dotp = pq[0] * pr[0] + pq[1] * pr[1] + pq[2] * pr[2];
    double dotq = -(qr[0] * pq[0] + qr[1] * pq[1] + qr[2] * pq[2]);
    double dotr = qr[0] * pr[0] + qr[1] * pr[1] + qr[2] * pr[2];
    if (dotp < 0) return 0.25 * area;
    if (dotq < 0 || dotr < 0) return 0.125 * area;
    double cotq = rxo_clamp_cot(dotq / area), cotr = rxo_clamp_cot(dotr / area);
    double lpr = pr[0] * pr[0] + pr[1] * pr[1] + pr[2] * pr[2], lpq = pq[0] * pq[0] + pq[1] * pq[1] + pq[2] * pq[2];
    return 0.125 * (lpr * cotq + lpq * cotr);
}
/* geometry_util.cuh:178-206 */
static double rxo_cot_part(const double* p, const double* r, const double* v)
{
    double area = rxo_tri_area(p, r, v);
    if (area > 1.17549435e-38) {
        double d = (p[0] - v[0]) * (r[0] - v[0]) + (p[1] - v[1]) * (r[1] - v[1]) + (p[2] - v[2]) * (r[2] - v[2]);
        return rxo_clamp_cot(d / area);
    }
    return 0;
}

/* MCF matrix-free mat-vec, cotan weights (apps/MCF/mcf_kernels.cuh:117-205):
 * out(p) = (1/vw + sum_w) in(p) - sum_r w(p,r) in(r), w = max(0, cot_q + cot_s) * time_step,
 * vw = 0.5 / sum of positive partial Voronoi areas. rings = rxo_oriented_rings. float64 from fp32 inputs. */
/* scale (may be NULL): per vertex |diag| |in_p| + sum_i |w_i| |in_i| -- the magnitude of the terms whose DIFFERENCE the   */
/* result is.  A fp32 evaluation is accurate relative to this scale (backward error), not relative to |out|: for a smooth */
/* input the terms cancel to a few per cent of their size.                                                              */
void rxo_mcf_matvec_scaled(const uint32_t* off, const uint32_t* val, uint32_t nv, const float* X, const float* in,
                           double time_step, double* out, double* scale);
void rxo_mcf_matvec(const uint32_t* off, const uint32_t* val, uint32_t nv, const float* X, const float* in,
                    double time_step, double* out)
{
    rxo_mcf_matvec_scaled(off, val, nv, X, in, time_step, out, NULL);
}
void rxo_mcf_matvec_scaled(const uint32_t* off, const uint32_t* val, uint32_t nv, const float* X, const float* in,
                           double time_step, double* out, double* scale)
{
    for (uint32_t p = 0; p < nv; ++p) {
        double P[3] = {X[3 * (uint64_t)p], X[3 * (uint64_t)p + 1], X[3 * (uint64_t)p + 2]};
        uint32_t k = off[p + 1] - off[p];
        const uint32_t* ring = val + off[p];
        double sum_w = 0, vw = 0, x[3] = {0, 0, 0}, mag = 0;
        for (uint32_t v = 0; v < k; ++v) {
            uint32_t qi = ring[(v + k - 1) % k], ri = ring[v], si = ring[(v + 1) % k];
            double Q[3], R[3], S[3];
            for (int c = 0; c < 3; ++c)
                Q[c] = X[3 * (uint64_t)qi + c], R[c] = X[3 * (uint64_t)ri + c], S[c] = X[3 * (uint64_t)si + c];
            double w = rxo_cot_part(P, R, Q) + rxo_cot_part(P, R, S);
            if (w < 0) w = 0;
            w *= time_step;
            sum_w += w;
            for (int c = 0; c < 3; ++c)
                x[c] -= w * in[3 * (uint64_t)ri + c];
            mag += w * sqrt((double)in[3 * (uint64_t)ri] * in[3 * (uint64_t)ri] + (double)in[3 * (uint64_t)ri + 1] * in[3 * (uint64_t)ri + 1] +
                            (double)in[3 * (uint64_t)ri + 2] * in[3 * (uint64_t)ri + 2]);
            double ta = rxo_partial_voronoi(P, Q, R);
            vw += ta > 0 ? ta : 0;
        }
        vw = 0.5 / vw;
        double diag = 1.0 / vw + sum_w;
        for (int c = 0; c < 3; ++c)
            out[3 * (uint64_t)p + c] = x[c] + diag * in[3 * (uint64_t)p + c];
        if (scale)
            scale[p] = mag + fabs(diag) * sqrt((double)in[3 * (uint64_t)p] * in[3 * (uint64_t)p] + (double)in[3 * (uint64_t)p + 1] * in[3 * (uint64_t)p + 1] +
                                               (double)in[3 * (uint64_t)p + 2] * in[3 * (uint64_t)p + 2]);
    }
}

/* MCF solve, matrix-free CG: apps/MCF/mcf_cg_mat_free.h:13-178 (driver), mcf_kernels.cuh:57-115 (init_B: B = X * valence or
 * X / v_weight), :117-205 (matvec, both Laplacians), matrix/cg_mat_free_attr_solver.h:45-125 (pre_solve: S = A X, R = B - S,
 * P = R, delta = <R,R>; solve: S = A P, alpha = delta / <S,P>, X += alpha P, R -= alpha S, delta' = <R,R>, stop when
 * delta' < tol_abs or delta' / delta0 < tol_rel (iterative_solver.h:57-63) -- the converging iteration is not counted --
 * else beta = delta' / delta, P = R + beta P).  float64 from fp32 coordinates; the three coordinates share one alpha / beta
 * (the reference's dot / norm2 run over all attributes).  info: [0] iterations, [1] converged, [2] <R0,R0>, [3] final <R,R>.
 * PARITY: the reference app has no correctness check for the solve; pinned to tests/golden/ref_mcf.npz, the solves of the
 * reference's own init_B / matvec / precond_matvec kernels (compiled unmodified) run on a B200 under the drop-in CG / PCG solver
 * headers (tests/golden/make_golden_mcf.py), and by the property A X = B. */
static void rxo_mcf_weights(const uint32_t* off, const uint32_t* val, uint32_t nv, const float* X, double time_step, int uniform,
                            double* W, double* diag, double* mass)
{
    for (uint32_t p = 0; p < nv; ++p) {
        double P[3] = {X[3 * (uint64_t)p], X[3 * (uint64_t)p + 1], X[3 * (uint64_t)p + 2]};
        uint32_t k = off[p + 1] - off[p];
        const uint32_t* ring = val + off[p];
        double sum_w = 0, vw = 0;
        for (uint32_t v = 0; v < k; ++v) {
            uint32_t qi = ring[(v + k - 1) % k], ri = ring[v], si = ring[(v + 1) % k];
            double Q[3], R[3], S[3], w = 1;
            for (int c = 0; c < 3; ++c)
                Q[c] = X[3 * (uint64_t)qi + c], R[c] = X[3 * (uint64_t)ri + c], S[c] = X[3 * (uint64_t)si + c];
            if (!uniform) {
                w = rxo_cot_part(P, R, Q) + rxo_cot_part(P, R, S);
                if (w < 0) w = 0;
            }
            w *= time_step;
            W[off[p] + v] = w;
            sum_w += w;
            if (uniform) {
                vw += 1;
            } else {
                double ta = rxo_partial_voronoi(P, Q, R);
                vw += ta > 0 ? ta : 0;
            }
        }
        /* 1 / v_weight with v_weight = 1 / valence (uniform) or 0.5 / area (cotangent) */
        mass[p] = k ? (uniform ? vw : 2.0 * vw) : 0.0;
        diag[p] = mass[p] + sum_w;
    }
}
static void rxo_mcf_apply(const uint32_t* off, const uint32_t* val, uint32_t nv, const double* W, const double* diag,
                          const double* in, double* out)
{
    for (uint32_t p = 0; p < nv; ++p) {
        double x[3] = {0, 0, 0};
        for (uint32_t i = off[p]; i < off[p + 1]; ++i)
            for (int c = 0; c < 3; ++c)
                x[c] -= W[i] * in[3 * (uint64_t)val[i] + c];
        for (int c = 0; c < 3; ++c)
            out[3 * (uint64_t)p + c] = x[c] + diag[p] * in[3 * (uint64_t)p + c];
    }
}
static double rxo_dot3(const double* a, const double* b, uint64_t n)
{
    double s = 0;
    for (uint64_t i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}
/* residual (may be NULL): B - A out, the property the tests check.
 * precond != 0: the Jacobi-preconditioned form, apps/MCF/mcf_cg_mat_free.h:181-254 with precond_matvec (mcf_kernels.cuh:216-295:
 * out = in / diag) under matrix/pcg_mat_free_attr_solver.h:40-140: P = Z = R / diag, delta = |<R,Z>| at the start, <R,Z> after,
 * the same stopping rule applied to delta, P = Z + beta P. */
int rxo_mcf_solve_ex(const uint32_t* off, const uint32_t* val, uint32_t nv, const float* X0, double time_step, int uniform,
                     int precond, uint32_t max_iter, double tol_abs, double tol_rel, double* out, double* residual, double* info)
{
    const uint64_t n = 3 * (uint64_t)nv;
    double* W = (double*)malloc(sizeof(double) * (off[nv] + 1));
    double* buf = (double*)malloc(sizeof(double) * (2 * (uint64_t)nv + 5 * n));
    if (!W || !buf) { free(W); free(buf); return 1; }
    double *diag = buf, *mass = buf + nv, *B = mass + nv, *R = B + n, *P = R + n, *S = P + n, *Z = S + n;
    rxo_mcf_weights(off, val, nv, X0, time_step, uniform, W, diag, mass);
    for (uint64_t i = 0; i < n; ++i) out[i] = X0[i], B[i] = X0[i] * mass[i / 3];
    rxo_mcf_apply(off, val, nv, W, diag, out, S);
    for (uint64_t i = 0; i < n; ++i) {
        R[i] = B[i] - S[i];
        Z[i] = precond ? (diag[i / 3] != 0 ? R[i] / diag[i / 3] : 0.0) : R[i];
        P[i] = Z[i];
    }
    double delta_new = fabs(rxo_dot3(R, Z, n)), start = delta_new;
    uint32_t it = 0;
    int conv = start == 0.0;
    while (!conv && it < max_iter) {
        rxo_mcf_apply(off, val, nv, W, diag, P, S);
        double alpha = delta_new / rxo_dot3(S, P, n);
        for (uint64_t i = 0; i < n; ++i) out[i] += alpha * P[i], R[i] -= alpha * S[i];
        for (uint64_t i = 0; i < n; ++i) Z[i] = precond ? (diag[i / 3] != 0 ? R[i] / diag[i / 3] : 0.0) : R[i];
        double delta_old = delta_new;
        delta_new = rxo_dot3(R, Z, n);
        if (delta_new < tol_abs || delta_new / start < tol_rel) { conv = 1; break; }
        double beta = delta_new / delta_old;
        for (uint64_t i = 0; i < n; ++i) P[i] = Z[i] + beta * P[i];
        ++it;
    }
    if (residual) {
        rxo_mcf_apply(off, val, nv, W, diag, out, S);
        for (uint64_t i = 0; i < n; ++i) residual[i] = B[i] - S[i];
    }
    info[0] = it, info[1] = conv, info[2] = start, info[3] = delta_new;
    free(W); free(buf);
    return 0;
}
int rxo_mcf_solve(const uint32_t* off, const uint32_t* val, uint32_t nv, const float* X0, double time_step, int uniform,
                  uint32_t max_iter, double tol_abs, double tol_rel, double* out, double* residual, double* info)
{
    return rxo_mcf_solve_ex(off, val, nv, X0, time_step, uniform, 0, max_iter, tol_abs, tol_rel, out, residual, info);
}
/* B - A x for a candidate solution x (double): the size-independent check of a solve done elsewhere; also returns <B,B> */
int rxo_mcf_residual(const uint32_t* off, const uint32_t* val, uint32_t nv, const float* X0, double time_step, int uniform,
                     const double* x, double* residual, double* bb)
{
    const uint64_t n = 3 * (uint64_t)nv;
    double* W = (double*)malloc(sizeof(double) * (off[nv] + 1));
    double* buf = (double*)malloc(sizeof(double) * (2 * (uint64_t)nv));
    if (!W || !buf) { free(W); free(buf); return 1; }
    double *diag = buf, *mass = buf + nv;
    rxo_mcf_weights(off, val, nv, X0, time_step, uniform, W, diag, mass);
    rxo_mcf_apply(off, val, nv, W, diag, x, residual);
    double s = 0;
    for (uint64_t i = 0; i < n; ++i) {
        double b = X0[i] * mass[i / 3];
        residual[i] = b - residual[i];
        s += b * b;
    }
    *bb = s;
    free(W); free(buf);
    return 0;
}

/* Gaussian curvature accumulators (apps/GaussianCurvature/gaussian_curvature_kernel.cuh:10-69):
 * per face corner v: gcs(v) -= angle_v; amix(v) += mixed Voronoi area share. float64 from fp32 inputs. */
void rxo_gaussian_curvature(const uint32_t* fv, uint32_t nf, const float* X, uint32_t nv, double* gcs, double* amix)
{
    const double PI = 3.14159265358979323846;
    memset(gcs, 0, (size_t)nv * sizeof(double));
    memset(amix, 0, (size_t)nv * sizeof(double));
    for (uint32_t f = 0; f < nf; ++f) {
        const uint32_t* t = fv + 3 * (uint64_t)f;
        double c[3][3];
        for (int i = 0; i < 3; ++i)
            for (int k = 0; k < 3; ++k)
                c[i][k] = X[3 * (uint64_t)t[i] + k];
        double l[3], dt[3], u[3], w[3];
        for (int i = 0; i < 3; ++i) {
            int j = (i + 1) % 3;
            l[i] = 0;
            for (int k = 0; k < 3; ++k) l[i] += (c[i][k] - c[j][k]) * (c[i][k] - c[j][k]);
        }
        for (int k = 0; k < 3; ++k) u[k] = c[1][k] - c[0][k], w[k] = c[2][k] - c[0][k];
        double cr[3] = {u[1] * w[2] - u[2] * w[1], u[2] * w[0] - u[0] * w[2], u[0] * w[1] - u[1] * w[0]};
        double s = sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
        for (int i = 0; i < 3; ++i) {
            int j = (i + 1) % 3, m = (i + 2) % 3;
            dt[i] = 0;
            for (int k = 0; k < 3; ++k) dt[i] += (c[j][k] - c[i][k]) * (c[m][k] - c[i][k]);
        }
        double rads[3];
        int    ob = 0;
        for (int i = 0; i < 3; ++i) {
            rads[i] = atan2(s, dt[i]);
            if (rads[i] > PI * 0.5) ob = 1;
        }
        for (int v = 0; v < 3; ++v) {
            int v1 = (v + 1) % 3, v2 = (v + 2) % 3;
            if (ob)
                amix[t[v]] += (rads[v] > PI * 0.5) ? 0.25 * s : 0.125 * s;
            else
                amix[t[v]] += 0.125 * (l[v2] * (dt[v1] / s) + l[v] * (dt[v2] / s));
            gcs[t[v]] -= rads[v];
        }
    }
}
