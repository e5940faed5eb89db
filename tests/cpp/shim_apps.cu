// tests/cpp/shim_apps.cu -- "user code" written against the reference's C++ API surface
// (RXMeshStatic / Query::dispatch / for_each<Op> / for_each_vertex / VertexAttribute), compiled against
// the drop-in headers in include/rxmesh/ and linked to librxmesh_b200.so.  The kernels follow the shape
// of the reference apps (apps/VertexNormal/vertex_normal_kernel.cuh:10-43, apps/Smoothing/manual.h:86-104,
// tests/RXMesh_test/query_kernel.cuh:13-46) so that a reader can see the same call sites compile unchanged.
// Exposed through extern "C" so tests/test_gpu_shim.py can drive it with ctypes.
#include <memory>
#include <vector>

#include "rxmesh/attribute.h"
#include "rxmesh/geometry_util.cuh"
#include "rxmesh/kernels/query_dispatcher.cuh"
#include "rxmesh/matrix/cg_mat_free_attr_solver.h"
#include "rxmesh/matrix/pcg_mat_free_attr_solver.h"
#include "rxmesh/query.h"
#include "rxmesh/reduce_handle.h"
#include "rxmesh/rxmesh_static.h"
#include "rxmesh/util/report.h"

using namespace rxmesh;

// RXM_REFSRC (tests/cpp/Makefile, only where /root/reference exists): the kernels named below are not the restatements
// in this file but THE REFERENCE'S OWN SOURCE FILES, included unmodified from the reference tree and compiled against
// the drop-in headers; everything else (the drivers, the exported entry points) is shared, so tests/test_gpu_shim.py runs
// the same checks on them.  1: apps/VertexNormal/vertex_normal_kernel.cuh, apps/GaussianCurvature/
// gaussian_curvature_kernel.cuh, apps/MCF/mcf_kernels.cuh (the matrix-free mat-vec), tests/RXMesh_test/query_kernel.cuh,
// tests/RXMesh_test/higher_query.cuh.  2: apps/Filtering/filtering_rxmesh_kernel.cuh (its
// compute_vertex_normal has the VertexNormal app's name and signature, hence a second translation unit).
#ifndef RXM_REFSRC
#define RXM_REFSRC 0
#endif
#if RXM_REFSRC == 1
#include "gaussian_curvature_kernel.cuh"
#include "higher_query.cuh"
#include "mcf_kernels.cuh"
#include "query_kernel.cuh"
#include "vertex_normal_kernel.cuh"
#define user_vertex_normal compute_vertex_normal
#define user_gaussian_curvature compute_gaussian_curvature
#define user_query_kernel query_kernel
#define user_higher_query higher_query
// compile-only: the Geodesic app's PTP relaxation kernel (apps/Geodesic/geodesic_kernel.cuh:95-185, a VV consumer outside
// this repo's app list) instantiates against the drop-in headers as it stands
#include "geodesic_kernel.cuh"
template __global__ void relax_ptp_rxmesh<float, 256>(const rxmesh::Context, const rxmesh::VertexAttribute<float>,
                                                      rxmesh::VertexAttribute<float>, const rxmesh::VertexAttribute<float>,
                                                      const rxmesh::VertexAttribute<int>, const int, const int, int*, const float,
                                                      const float);
#elif RXM_REFSRC == 2
#include "filtering_rxmesh_kernel.cuh"
#define user_filter_vertex_normal compute_vertex_normal
#define user_bilateral_filtering bilateral_filtering
#endif

#if RXM_REFSRC != 1
template <typename T, uint32_t blockThreads>
__global__ static void user_vertex_normal(const Context context, VertexAttribute<T> coords, VertexAttribute<T> normals)
{
    auto vn_lambda = [&](FaceHandle face_id, VertexIterator& fv) {
        (void)face_id;
        vec3<T> c0 = coords.template to_glm<3>(fv[0]);
        vec3<T> c1 = coords.template to_glm<3>(fv[1]);
        vec3<T> c2 = coords.template to_glm<3>(fv[2]);
        vec3<T> n  = cross(c1 - c0, c2 - c0);
        vec3<T> l(glm::distance2(c0, c1), glm::distance2(c1, c2), glm::distance2(c2, c0));
        for (uint32_t v = 0; v < 3; ++v)
            for (uint32_t i = 0; i < 3; ++i)
                atomicAdd(&normals(fv[v], i), n[i] / (l[v] + l[(v + 2) % 3]));
    };
    auto                block = cooperative_groups::this_thread_block();
    Query<blockThreads> query(context);
    ShmemAllocator      shrd_alloc;
    query.template dispatch<Op::FV>(block, shrd_alloc, vn_lambda);
}
#endif

// the MCF matrix-free mat-vec with cotan weights (apps/MCF/mcf_kernels.cuh:117-205): oriented VV
#if RXM_REFSRC != 1
template <typename T, uint32_t blockThreads>
__global__ static void user_mcf_matvec(const Context context, const VertexAttribute<T> coords, const VertexAttribute<T> in,
                                       VertexAttribute<T> out, const T time_step)
{
    auto matvec_lambda = [&](VertexHandle& p_id, const VertexIterator& iter) {
        T            sum_e_weight(0), v_weight(0);
        vec3<T>      x(T(0));
        const vec3<T> p = coords.template to_glm<3>(p_id);
        VertexHandle q_id = iter.back();
        for (uint32_t v = 0; v < iter.size(); ++v) {
            VertexHandle r_id = iter[v];
            VertexHandle s_id = (v == iter.size() - 1) ? iter[0] : iter[v + 1];
            const vec3<T> r = coords.template to_glm<3>(r_id), q = coords.template to_glm<3>(q_id),
                          s = coords.template to_glm<3>(s_id);
            T e_weight = edge_cotan_weight(p, r, q, s);
            e_weight   = (static_cast<T>(e_weight >= 0.0)) * e_weight;
            e_weight *= time_step;
            sum_e_weight += e_weight;
            x[0] -= e_weight * in(r_id, 0);
            x[1] -= e_weight * in(r_id, 1);
            x[2] -= e_weight * in(r_id, 2);
            T tri = partial_voronoi_area(p, q, r);
            v_weight += (tri > 0) ? tri : 0;
            q_id = r_id;
        }
        v_weight     = 0.5 / v_weight;
        T diag       = ((1.0 / v_weight) + sum_e_weight);
        out(p_id, 0) = x[0] + diag * in(p_id, 0);
        out(p_id, 1) = x[1] + diag * in(p_id, 1);
        out(p_id, 2) = x[2] + diag * in(p_id, 2);
    };
    auto                block = cooperative_groups::this_thread_block();
    Query<blockThreads> query(context);
    ShmemAllocator      shrd_alloc;
    query.template dispatch<Op::VV>(block, shrd_alloc, matvec_lambda, [](VertexHandle) { return true; }, true);
}
#endif

// the Jacobi preconditioner of the MCF system (apps/MCF/mcf_kernels.cuh:216-295): out = in / diagonal of the mat-vec above
#if RXM_REFSRC != 1
template <typename T, uint32_t blockThreads>
__global__ static void user_mcf_precond(const Context context, const VertexAttribute<T> coords, const VertexAttribute<T> in,
                                        VertexAttribute<T> out, const T time_step)
{
    auto lambda = [&](VertexHandle& p_id, const VertexIterator& iter) {
        T             sum_e_weight(0), v_weight(0);
        const vec3<T> p    = coords.template to_glm<3>(p_id);
        VertexHandle  q_id = iter.back();
        for (uint32_t v = 0; v < iter.size(); ++v) {
            VertexHandle  r_id = iter[v];
            VertexHandle  s_id = (v == iter.size() - 1) ? iter[0] : iter[v + 1];
            const vec3<T> r = coords.template to_glm<3>(r_id), q = coords.template to_glm<3>(q_id),
                          s = coords.template to_glm<3>(s_id);
            T e_weight = edge_cotan_weight(p, r, q, s);
            e_weight   = (static_cast<T>(e_weight >= 0.0)) * e_weight;
            sum_e_weight += e_weight * time_step;
            T tri = partial_voronoi_area(p, q, r);
            v_weight += (tri > 0) ? tri : 0;
            q_id = r_id;
        }
        v_weight = 0.5 / v_weight;
        T diag   = ((1.0 / v_weight) + sum_e_weight);
        for (uint32_t i = 0; i < 3; ++i)
            out(p_id, i) = in(p_id, i) / diag;
    };
    auto                block = cooperative_groups::this_thread_block();
    Query<blockThreads> query(context);
    ShmemAllocator      shrd_alloc;
    query.template dispatch<Op::VV>(block, shrd_alloc, lambda, [](VertexHandle) { return true; }, true);
}
#endif

// Gaussian curvature accumulators (apps/GaussianCurvature/gaussian_curvature_kernel.cuh:10-69): FV + atomics
#if RXM_REFSRC != 1
template <typename T, uint32_t blockThreads>
__global__ static void user_gaussian_curvature(const Context context, VertexAttribute<T> coords, VertexAttribute<T> gcs,
                                               VertexAttribute<T> amix)
{
    auto gc_lambda = [&](FaceHandle, VertexIterator& fv) {
        const vec3<T> c0 = coords.template to_glm<3>(fv[0]), c1 = coords.template to_glm<3>(fv[1]),
                      c2 = coords.template to_glm<3>(fv[2]);
        vec3<T> l(glm::distance2(c0, c1), glm::distance2(c1, c2), glm::distance2(c2, c0));
        T       s = glm::length(glm::cross(c1 - c0, c2 - c0));
        vec3<T> c(glm::dot(c1 - c0, c2 - c0), glm::dot(c2 - c1, c0 - c1), glm::dot(c0 - c2, c1 - c2));
        vec3<T> rads(atan2(s, c[0]), atan2(s, c[1]), atan2(s, c[2]));
        const T half_pi = T(1.57079632679489661923);
        bool    is_ob   = false;
        for (int i = 0; i < 3; ++i)
            if (rads[i] > half_pi) is_ob = true;
        for (uint32_t v = 0; v < 3; ++v) {
            uint32_t v1 = (v + 1) % 3, v2 = (v + 2) % 3;
            if (is_ob)
                atomicAdd(&amix(fv[v]), rads[v] > half_pi ? T(0.25) * s : T(0.125) * s);
            else
                atomicAdd(&amix(fv[v]), T(0.125) * ((l[v2]) * (c[v1] / s) + (l[v]) * (c[v2] / s)));
            atomicAdd(&gcs(fv[v]), -rads[v]);
        }
    };
    auto                block = cooperative_groups::this_thread_block();
    Query<blockThreads> query(context);
    ShmemAllocator      shrd_alloc;
    query.template dispatch<Op::FV>(block, shrd_alloc, gc_lambda);
}
#endif

// ---- the Filtering app (apps/Filtering/filtering_rxmesh_kernel.cuh:15-85,426-548, filtering_util.h:31-76):
// unit-face-normal vertex normals, then per vertex a breadth-first k-ring gathered with the free-function
// query_block_dispatcher (first ring) and higher_query_block_dispatcher (every further ring), then the bilateral update
#if RXM_REFSRC != 2
template <typename T, uint32_t blockThreads>
__global__ static void user_filter_vertex_normal(const Context context, VertexAttribute<T> coords, VertexAttribute<T> normals)
{
    auto vn_lambda = [&](FaceHandle, VertexIterator& fv) {
        const vec3<T> c0 = coords.template to_glm<3>(fv[0]), c1 = coords.template to_glm<3>(fv[1]),
                      c2 = coords.template to_glm<3>(fv[2]);
        const vec3<T> n = glm::normalize(glm::cross(c1 - c0, c2 - c0));
        for (uint32_t v = 0; v < 3; ++v)
            for (uint32_t i = 0; i < 3; ++i)
                atomicAdd(&normals(fv[v], i), n[i]);
    };
    auto                block = cooperative_groups::this_thread_block();
    Query<blockThreads> query(context);
    ShmemAllocator      shrd_alloc;
    query.template dispatch<Op::FV>(block, shrd_alloc, vn_lambda);
}

template <typename T, typename S>
__device__ __forceinline__ bool user_linear_search(const T list[], const T item, const S end)
{
    for (S i = 0; i < end; ++i)
        if (list[i] == item) return true;
    return false;
}

template <typename T, uint32_t blockThreads, uint32_t maxVVSize>
__global__ static void user_bilateral_filtering(const Context context, VertexAttribute<T> input_coords,
                                                VertexAttribute<T> filtered_coords, VertexAttribute<T> vertex_normals)
{
    VertexHandle vv[maxVVSize];
    uint32_t     num_vv     = 0;
    T            sigma_c_sq = 0, radius = 0;
    vec3<T>      vertex, normal;
    VertexHandle v_id;

    auto first_ring = [&](VertexHandle& p_id, VertexIterator& iter) {
        v_id   = p_id;
        vertex = input_coords.template to_glm<3>(v_id);
        normal = glm::normalize(vertex_normals.template to_glm<3>(v_id));
        vv[0]  = v_id;
        ++num_vv;
        sigma_c_sq = 1e10;
        for (uint32_t v = 0; v < iter.size(); ++v) {
            const T len = glm::distance2(vertex, input_coords.template to_glm<3>(iter[v]));
            if (len < sigma_c_sq) sigma_c_sq = len;
        }
        radius = 4.0 * sigma_c_sq;
        for (uint32_t v = 0; v < iter.size(); ++v) {
            const VertexHandle vv_id = iter[v];
            if (glm::distance2(vertex, input_coords.template to_glm<3>(vv_id)) <= radius) vv[num_vv++] = vv_id;
        }
    };
    query_block_dispatcher<Op::VV, blockThreads>(context, first_ring);
    __syncthreads();

    uint32_t next_id = 1;
    while (true) {
        VertexHandle next_vertex;
        if (v_id.is_valid() && next_id < num_vv) next_vertex = vv[next_id];
        auto n_rings = [&](const VertexHandle& id, const VertexIterator& iter) {
            for (uint32_t i = 0; i < iter.size(); ++i) {
                const VertexHandle vvv_id = iter[i];
                if (vvv_id != v_id && !user_linear_search(vv, vvv_id, num_vv)) {
                    if (glm::distance2(input_coords.template to_glm<3>(vvv_id), vertex) <= radius && num_vv < maxVVSize)
                        vv[num_vv++] = vvv_id;
                }
            }
        };
        higher_query_block_dispatcher<Op::VV, blockThreads>(context, next_vertex, n_rings);
        const bool is_done = (next_id >= num_vv) || !v_id.is_valid();
        if (__syncthreads_and(is_done)) break;
        next_id++;
    }

    if (v_id.is_valid()) {
        // compute_sigma_s_sq (filtering_util.h:31-59) + compute_new_coordinates (filtering_rxmesh_kernel.cuh:52-85)
        T sum = 0, sum_sq = 0;
        for (uint32_t i = 0; i < num_vv; ++i) {
            T t = fabs(glm::dot(input_coords.template to_glm<3>(vv[i]) - vertex, normal));
            sum += t, sum_sq += t * t;
        }
        const T c          = static_cast<T>(num_vv);
        T       sigma_s_sq = (sum_sq / c) - ((sum * sum) / (c * c));
        sigma_s_sq         = (sigma_s_sq < 1.0e-20) ? (sigma_s_sq + 1.0e-20) : sigma_s_sq;
        T acc = 0, normalizer = 0;
        for (uint32_t i = 0; i < num_vv; ++i) {
            const vec3<T> q  = input_coords.template to_glm<3>(vv[i]) - vertex;
            const T       t  = glm::length(q), h = glm::dot(q, normal);
            const T       wc = exp(-0.5 * t * t / sigma_c_sq), ws = exp(-0.5 * h * h / sigma_s_sq);
            acc += wc * ws * h, normalizer += wc * ws;
        }
        vertex += normal * (acc / normalizer);
        filtered_coords(v_id, 0) = vertex[0], filtered_coords(v_id, 1) = vertex[1], filtered_coords(v_id, 2) = vertex[2];
    }
}
#endif

// ---- the split query API + valence + device for_each, in one user kernel:
//   deg_split(v) = iterator size from prologue / get_iterator;  deg_val(v) = vertex_valence;  face_cnt(v) = #incident faces
//   via run_compute on a second query;  touched(f) = 1 through for_each<Op::F>;  owner(v) = patch of get_owner_handle
template <uint32_t blockThreads>
__global__ static void user_split_api(const Context context, VertexAttribute<int> deg_split, VertexAttribute<int> deg_val,
                                      VertexAttribute<int> face_cnt, FaceAttribute<int> touched, VertexAttribute<int> owner_ok)
{
    auto                block = cooperative_groups::this_thread_block();
    Query<blockThreads> query(context);
    ShmemAllocator      shrd_alloc;
    query.compute_vertex_valence(block, shrd_alloc);
    query.template prologue<Op::VV>(block, shrd_alloc);
    const uint32_t nv_owned = query.get_patch_info().n_owned[0], nv = query.get_patch_info().n[0];
    for (uint32_t v = threadIdx.x; v < nv; v += blockThreads) {
        VertexHandle vh(query.get_patch_id(), LocalVertexT((uint16_t)v));
        if (v < nv_owned) {
            VertexIterator it = query.template get_iterator<VertexIterator>((uint16_t)v);
            deg_split(vh)     = it.size();
            deg_val(vh)       = query.vertex_valence(vh);
        } else {
            // a ribbon copy resolves to an owner handle that lives in another patch and is owned there
            const VertexHandle oh = context.get_owner_handle(vh);
            if (oh.patch_id() == vh.patch_id() || oh.local_id() >= context.view.desc[oh.patch_id()].n_owned[0])
                owner_ok(VertexHandle(query.get_patch_id(), LocalVertexT(0))) = 0;
        }
    }
    query.epilogue(block, shrd_alloc);
    query.template prologue<Op::VF>(block, shrd_alloc, [](VertexHandle) { return true; }, false, false);
    query.run_compute(block, [&](const VertexHandle& vh, const FaceIterator& it) { face_cnt(vh) = it.size(); });
    query.epilogue(block, shrd_alloc);
    for_each<Op::F, blockThreads>(context, [&](const FaceHandle fh) { touched(fh) = 1; });
}

// ---- tests/RXMesh_test/test_multi_queries.cu:10-92: vertex edge-length sums, once by an EV scatter with atomics and once
// by a primary VE query whose lambda reads a SECONDARY EV query through prologue / get_iterator(local) / epilogue
template <uint32_t blockThreads, typename T>
__global__ static void user_sum_edges_ev(const Context context, const VertexAttribute<T> coords, VertexAttribute<T> vertex_sum)
{
    auto sum_edges = [&](const EdgeHandle&, const VertexIterator& iter) {
        const T edge_len = glm::distance2(coords.template to_glm<3>(iter[0]), coords.template to_glm<3>(iter[1]));
        ::atomicAdd(&vertex_sum(iter[0]), edge_len);
        ::atomicAdd(&vertex_sum(iter[1]), edge_len);
    };
    auto                block = cooperative_groups::this_thread_block();
    Query<blockThreads> query(context);
    ShmemAllocator      shrd_alloc;
    query.template dispatch<Op::EV>(block, shrd_alloc, sum_edges);
}
template <uint32_t blockThreads, typename T>
__global__ static void user_sum_edges_multi_queries(const Context context, const VertexAttribute<T> coords,
                                                    VertexAttribute<T> vertex_sum)
{
    auto                block = cooperative_groups::this_thread_block();
    ShmemAllocator      shrd_alloc;
    Query<blockThreads> ev_query(context);
    ev_query.template prologue<Op::EV>(block, shrd_alloc);  // the secondary query
    auto sum_edges = [&](const VertexHandle& vertex, const EdgeIterator& eiter) {
        const vec3<T> p0 = coords.template to_glm<3>(vertex);
        for (uint16_t i = 0; i < eiter.size(); ++i) {
            VertexIterator     viter = ev_query.template get_iterator<VertexIterator>(eiter.local(i));
            const VertexHandle vh0(viter[0]), vh1(viter[1]);
            const vec3<T>      p1 = coords.template to_glm<3>(vertex != vh0 ? vh0 : vh1);
            vertex_sum(vertex) += glm::distance2(p0, p1);
        }
    };
    Query<blockThreads> ve_query(context);  // the primary query
    ve_query.template dispatch<Op::VE>(block, shrd_alloc, sum_edges);
    ev_query.epilogue(block, shrd_alloc);
}

// ---- tests/RXMesh_test/higher_query.cuh:15-90: 2-ring VV through query_block_dispatcher + higher_query_block_dispatcher
#if RXM_REFSRC != 1
template <uint32_t blockThreads, Op op>
__global__ static void user_higher_query(const Context context, VertexAttribute<VertexHandle> input,
                                         VertexAttribute<VertexHandle> output)
{
    VertexHandle thread_vertex;
    uint32_t     num_vv_1st_ring(0), num_vv(0);
    auto first_ring_lambda = [&](VertexHandle id, Iterator<VertexHandle>& iter) {
        num_vv_1st_ring = iter.size();
        num_vv          = num_vv_1st_ring;
        thread_vertex        = id;
        input(thread_vertex) = thread_vertex;
        for (uint32_t i = 0; i < iter.size(); ++i)
            output(thread_vertex, i) = iter[i];
    };
    query_block_dispatcher<op, blockThreads>(context, first_ring_lambda);
    uint32_t next_id = 0;
    while (true) {
        VertexHandle next_vertex;
        if (thread_vertex.is_valid() && next_id < num_vv_1st_ring) next_vertex = output(thread_vertex, next_id);
        auto higher_rings_lambda = [&](const VertexHandle&, const VertexIterator& iter) {
            for (uint32_t i = 0; i < iter.size(); ++i) {
                if (iter[i] != thread_vertex) {
                    bool duplicate = false;
                    for (uint32_t j = 0; j < num_vv; ++j)
                        if (iter[i] == output(thread_vertex, j)) {
                            duplicate = true;
                            break;
                        }
                    if (!duplicate) {
                        output(thread_vertex, num_vv) = iter[i];
                        num_vv++;
                    }
                }
            }
        };
        higher_query_block_dispatcher<op, blockThreads>(context, next_vertex, higher_rings_lambda);
        const bool is_done = (next_id >= num_vv_1st_ring) || !thread_vertex.is_valid();
        if (__syncthreads_and(is_done)) break;
        next_id++;
    }
}
#endif

// ---- unit tests of the device building blocks (the reference's Util.Scan / Util.BlockMatrixTranspose,
// tests/RXMesh_test/test_util.cu:160-358): block_exclusive_scan and csr_transpose from rxm_device.cuh, one block
template <int BT>
__global__ static void unit_block_scan(uint32_t* a, uint32_t n)
{
    extern __shared__ __align__(16) uint8_t raw[];
    uint32_t* s   = reinterpret_cast<uint32_t*>(raw);
    uint32_t* tmp = s + n + 1;
    for (uint32_t i = threadIdx.x; i < n; i += BT)
        s[i] = a[i];
    rxm::dev::block_exclusive_scan<BT>(s, n, tmp);
    for (uint32_t i = threadIdx.x; i <= n; i += BT)
        a[i] = s[i];
}
// src: [rows * deg] column ids; out: off[cols + 1], val[rows * deg] (row ids, each list sorted ascending)
template <int BT, int KMAX>
__global__ static void unit_block_transpose(const uint16_t* src, uint32_t rows, uint32_t cols, uint32_t deg, uint32_t* off,
                                            uint16_t* val)
{
    extern __shared__ __align__(16) uint8_t raw[];
    uint32_t*      s_off = reinterpret_cast<uint32_t*>(raw);
    uint32_t*      tmp   = s_off + cols + 1;
    uint16_t*      s_val = reinterpret_cast<uint16_t*>(tmp + 40);
    const uint32_t nnz   = rows * deg;
    rxm::dev::csr_transpose<BT, KMAX>(
        nnz, cols, s_off, s_val, tmp, [=](uint32_t i) { return (uint32_t)src[i]; }, [=](uint32_t i) { return i / deg; });
    rxm::dev::csr_sort_lists<BT>(cols, s_off, s_val);
    for (uint32_t i = threadIdx.x; i <= cols; i += BT)
        off[i] = s_off[i];
    for (uint32_t i = threadIdx.x; i < nnz; i += BT)
        val[i] = s_val[i];
}

// a kernel launched through run_kernel: out(v) = scale * valence(v)
template <uint32_t blockThreads>
__global__ static void user_scaled_valence(const Context context, VertexAttribute<float> out, float scale)
{
    auto lambda = [&](VertexHandle& vh, const VertexIterator& it) { out(vh) = scale * it.size(); };
    query_block_dispatcher<Op::VV, blockThreads>(context, lambda);
}

#if RXM_REFSRC != 1
template <uint32_t blockThreads, Op op, typename InH, typename OutH, typename InA, typename OutA>
__global__ static void user_query_kernel(const Context context, InA input, OutA output, const bool oriented)
{
    auto store = [&](const InH& id, const Iterator<OutH>& iter) {
        input(id) = id;
        for (uint32_t i = 0; i < iter.size(); ++i)
            output(id, i) = iter[i];
    };
    auto                block = cooperative_groups::this_thread_block();
    Query<blockThreads> query(context);
    ShmemAllocator      shrd_alloc;
    query.template dispatch<op>(block, shrd_alloc, store, [](InH) { return true; }, oriented);
}
#endif

static std::vector<std::vector<uint32_t>> to_faces(const uint32_t* fv, uint32_t nf)
{
    std::vector<std::vector<uint32_t>> F(nf, std::vector<uint32_t>(3));
    for (uint32_t f = 0; f < nf; ++f)
        for (int i = 0; i < 3; ++i)
            F[f][i] = fv[3 * f + i];
    return F;
}
static std::vector<std::vector<float>> to_verts(const float* x, uint32_t nv)
{
    std::vector<std::vector<float>> V(nv, std::vector<float>(3));
    for (uint32_t v = 0; v < nv; ++v)
        for (int i = 0; i < 3; ++i)
            V[v][i] = x[3 * v + i];
    return V;
}

template <typename H, typename L>
static void host_for_each(RXMeshStatic& rx, L f)
{
    if constexpr (std::is_same_v<H, VertexHandle>) rx.for_each_vertex(HOST, f, NULL, false);
    if constexpr (std::is_same_v<H, EdgeHandle>) rx.for_each_edge(HOST, f, NULL, false);
    if constexpr (std::is_same_v<H, FaceHandle>) rx.for_each_face(HOST, f, NULL, false);
}

// vertex valence through for_each<Op::VV> + a step through for_each_vertex(DEVICE) (both lambda paths)
// (nvcc: extended __device__ lambdas must not be defined inside extern "C" functions, so the apps are
// plain C++ functions and the extern "C" entry points at the bottom only forward)
static int app_valence(const uint32_t* fv, uint32_t nf, uint32_t patch_size, float* out_valence, float* out_plus1)
{
    rx_init(0);
    RXMeshStatic rx(to_faces(fv, nf), "", patch_size);
    auto val = *rx.add_vertex_attribute<float>("val", 1, LOCATION_ALL);
    auto one = *rx.add_vertex_attribute<float>("one", 1, LOCATION_ALL);
    val.reset(-1.f, DEVICE);
    one.reset(0.f, DEVICE);
    rx.for_each<Op::VV, 256>([=] __device__(const VertexHandle& vh, const VertexIterator& iter) mutable { val(vh) = iter.size(); });
    rx.for_each_vertex(DEVICE, [val, one] __device__(const VertexHandle& vh) { one(vh) = val(vh) + 1.f; });
    if (cudaDeviceSynchronize() != cudaSuccess) return 1;
    val.move(DEVICE, HOST);
    one.move(DEVICE, HOST);
    rx.for_each_vertex(HOST, [&](const VertexHandle& vh) {
        out_valence[rx.map_to_global(vh)] = val(vh);
        out_plus1[rx.map_to_global(vh)]   = one(vh);
    });
    return 0;
}

static int app_mcf_matvec(const uint32_t* fv, uint32_t nf, const float* x, const float* vin, uint32_t nv, uint32_t patch_size,
                          float time_step, float* out)
{
    rx_init(0);
    RXMeshStatic rx(to_faces(fv, nf), "", patch_size);
    constexpr uint32_t blockThreads = 256;
    auto coords = rx.add_vertex_attribute<float>(to_verts(x, nv), "coords");
    auto in     = rx.add_vertex_attribute<float>(to_verts(vin, nv), "in");
    auto res    = rx.add_vertex_attribute<float>("out", 3, LOCATION_ALL);
    LaunchBox<blockThreads> lb;
#if RXM_REFSRC == 1  // the reference's kernel: matvec(context, coords, in, out, use_uniform_laplace, time_step)
    rx.prepare_launch_box({Op::VV}, lb, (void*)matvec<float, blockThreads>, true);
    matvec<float, blockThreads><<<lb.blocks, lb.num_threads, lb.smem_bytes_dyn>>>(rx.get_context(), *coords, *in, *res, false, time_step);
#else
    rx.prepare_launch_box({Op::VV}, lb, (void*)user_mcf_matvec<float, blockThreads>, true);
    user_mcf_matvec<float, blockThreads><<<lb.blocks, lb.num_threads, lb.smem_bytes_dyn>>>(rx.get_context(), *coords, *in, *res, time_step);
#endif
    if (cudaDeviceSynchronize() != cudaSuccess) return 1;
    res->move(DEVICE, HOST);
    rx.for_each_vertex(HOST, [&](const VertexHandle& vh) {
        for (uint32_t i = 0; i < 3; ++i)
            out[rx.map_to_global(vh) * 3 + i] = (*res)(vh, i);
    }, NULL, false);
    return 0;
}

// The MCF app's solve (apps/MCF/mcf_cg_mat_free.h:13-178) as user code: init_B, then CGMatFreeAttrSolver (the drop-in
// header include/rxmesh/matrix/cg_mat_free_attr_solver.h) driving the matrix-free mat-vec kernel through run_kernel.  With
// RXM_REFSRC == 1 both kernels are the reference's own (apps/MCF/mcf_kernels.cuh, unmodified, either Laplacian); otherwise the
// restated cotangent mat-vec above, with B = M X0 taken from a mat-vec at time step 0.  info[4]: iterations, start residual,
// final residual, milliseconds of pre_solve + solve.
static int app_mcf_cg(const uint32_t* fv, uint32_t nf, const float* x, uint32_t nv, uint32_t patch_size, float time_step,
                      int uniform, int pcg, int max_iter, float tol_abs, float tol_rel, float* out, float* info)
{
    rx_init(0);
    RXMeshStatic rx(to_faces(fv, nf), "", patch_size);
    if (!rx.is_closed()) return 2;
    constexpr uint32_t blockThreads = 256;
    auto coords = rx.add_vertex_attribute<float>(to_verts(x, nv), "coords");
    auto B      = rx.add_vertex_attribute<float>("B", 3, DEVICE);
    auto X      = rx.add_vertex_attribute<float>("X", 3, LOCATION_ALL);
    B->reset(0.f, DEVICE);
    X->copy_from(*coords, DEVICE, DEVICE);
    LaunchBox<blockThreads> lb;
#if RXM_REFSRC == 1
    const bool uni = uniform != 0;
    rx.prepare_launch_box({Op::VV}, lb, (void*)init_B<float, blockThreads>, !uni);
    init_B<float, blockThreads><<<lb.blocks, lb.num_threads, lb.smem_bytes_dyn>>>(rx.get_context(), *X, *B, uni);
    rx.prepare_launch_box({Op::VV}, lb, (void*)matvec<float, blockThreads>, !uni);
    auto mat_vec = [&](const VertexAttribute<float>& in, VertexAttribute<float>& o, cudaStream_t stream) {
        rx.run_kernel(lb, matvec<float, blockThreads>, stream, *coords, in, o, uni, time_step);
    };
#else
    if (uniform) return 3;  // the restated kernel has the cotangent weights only
    rx.prepare_launch_box({Op::VV}, lb, (void*)user_mcf_matvec<float, blockThreads>, true);
    user_mcf_matvec<float, blockThreads><<<lb.blocks, lb.num_threads, lb.smem_bytes_dyn>>>(rx.get_context(), *coords, *X, *B, 0.f);
    auto mat_vec = [&](const VertexAttribute<float>& in, VertexAttribute<float>& o, cudaStream_t stream) {
        rx.run_kernel(lb, user_mcf_matvec<float, blockThreads>, stream, *coords, in, o, time_step);
    };
#endif
    // pcg: mcf_pcg_mat_free (mcf_cg_mat_free.h:181-254): the Jacobi kernel as the second std::function
    LaunchBox<blockThreads> plb;
#if RXM_REFSRC == 1
    rx.prepare_launch_box({Op::VV}, plb, (void*)precond_matvec<float, blockThreads>, !uni);
    auto precond = [&](const VertexAttribute<float>& in, VertexAttribute<float>& o, cudaStream_t stream) {
        rx.run_kernel(plb, precond_matvec<float, blockThreads>, stream, *coords, in, o, uni, time_step);
    };
#else
    rx.prepare_launch_box({Op::VV}, plb, (void*)user_mcf_precond<float, blockThreads>, true);
    auto precond = [&](const VertexAttribute<float>& in, VertexAttribute<float>& o, cudaStream_t stream) {
        rx.run_kernel(plb, user_mcf_precond<float, blockThreads>, stream, *coords, in, o, time_step);
    };
#endif
    if (cudaDeviceSynchronize() != cudaSuccess) return 1;
    std::unique_ptr<IterativeSolver<float, VertexAttribute<float>>> sp;  // one solver: both register the attributes CG:S / P / R
    if (pcg)
        sp.reset(new PCGMatFreeAttrSolver<float, VertexHandle>(rx, mat_vec, precond, 3, max_iter, tol_abs, tol_rel));
    else
        sp.reset(new CGMatFreeAttrSolver<float, VertexHandle>(rx, mat_vec, 3, max_iter, tol_abs, tol_rel));
    IterativeSolver<float, VertexAttribute<float>>& solver = *sp;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    cudaEventRecord(e0, NULL);
    solver.pre_solve(*B, *X, NULL);
    solver.solve(*B, *X, NULL);
    cudaEventRecord(e1, NULL);
    if (cudaDeviceSynchronize() != cudaSuccess) return 1;
    info[0] = (float)solver.iter_taken(), info[1] = solver.start_residual(), info[2] = solver.final_residual();
    cudaEventElapsedTime(&info[3], e0, e1);  // pre_solve + solve, ms (what the app reports as pre-solve + solve)
    cudaEventDestroy(e0), cudaEventDestroy(e1);
    X->move(DEVICE, HOST);
    rx.for_each_vertex(HOST, [&](const VertexHandle& vh) {
        for (uint32_t i = 0; i < 3; ++i)
            out[rx.map_to_global(vh) * 3 + i] = (*X)(vh, i);
    }, NULL, false);
    return 0;
}

static int app_gaussian_curvature(const uint32_t* fv, uint32_t nf, const float* x, uint32_t nv, uint32_t patch_size,
                                  float* out_gcs, float* out_amix)
{
    rx_init(0);
    RXMeshStatic rx(to_faces(fv, nf), "", patch_size);
    constexpr uint32_t blockThreads = 256;
    auto coords = rx.add_vertex_attribute<float>(to_verts(x, nv), "coords");
    auto gcs    = rx.add_vertex_attribute<float>("gcs", 1, LOCATION_ALL);
    auto amix   = rx.add_vertex_attribute<float>("amix", 1, LOCATION_ALL);
    gcs->reset(0, DEVICE);
    amix->reset(0, DEVICE);
    LaunchBox<blockThreads> lb;
    rx.prepare_launch_box({Op::FV}, lb, (void*)user_gaussian_curvature<float, blockThreads>);
    user_gaussian_curvature<float, blockThreads><<<lb.blocks, lb.num_threads, lb.smem_bytes_dyn>>>(rx.get_context(), *coords, *gcs, *amix);
    if (cudaDeviceSynchronize() != cudaSuccess) return 1;
    gcs->move(DEVICE, HOST);
    amix->move(DEVICE, HOST);
    rx.for_each_vertex(HOST, [&](const VertexHandle& vh) {
        out_gcs[rx.map_to_global(vh)]  = (*gcs)(vh);
        out_amix[rx.map_to_global(vh)] = (*amix)(vh);
    }, NULL, false);
    return 0;
}

// the query test (tests/RXMesh_test/test_queries.h:98-219) for one op; out_global: [num_src][width] global ids of
// the output handles (0xFFFFFFFF = invalid), rows in GLOBAL source order; returns -1 on a failed invariant
template <Op op, typename InH, typename OutH>
static int run_query(RXMeshStatic& rx, uint32_t width, bool oriented, uint32_t* out_global)
{
    constexpr uint32_t blockThreads = 256;
    auto input  = rx.add_attribute<InH, InH>("input", 1);
    auto output = rx.add_attribute<OutH, InH>("output", width);
    input->reset(InH(), DEVICE);
    output->reset(OutH(), DEVICE);
    LaunchBox<blockThreads> lb;
    auto kern = user_query_kernel<blockThreads, op, InH, OutH, Attribute<InH, InH>, Attribute<OutH, InH>>;
    rx.prepare_launch_box({op}, lb, (void*)kern, oriented);
    kern<<<lb.blocks, blockThreads, lb.smem_bytes_dyn>>>(rx.get_context(), *input, *output, oriented);
    if (cudaDeviceSynchronize() != cudaSuccess) return 1;
    input->move(DEVICE, HOST);
    output->move(DEVICE, HOST);
    int bad = 0;
    host_for_each<InH>(rx, [&](const InH& h) {
        if ((*input)(h) != h) bad = 1;
        const uint32_t g = rx.map_to_global(h);
        for (uint32_t i = 0; i < width; ++i) {
            OutH o = (*output)(h, i);
            out_global[(size_t)g * width + i] = o.is_valid() ? rx.map_to_global(o) : 0xFFFFFFFFu;
        }
    });
    rx.remove_attribute("input");
    rx.remove_attribute("output");
    return bad ? -1 : 0;
}

// the VertexNormal app (apps/VertexNormal/vertex_normal.cu:28-110), RXMesh path
static int app_vertex_normals(const uint32_t* fv, uint32_t nf, const float* x, uint32_t nv, uint32_t patch_size, float* out)
{
    rx_init(0);
    RXMeshStatic rx(to_faces(fv, nf), "", patch_size);
    constexpr uint32_t blockThreads = 256;
    auto coords    = rx.add_vertex_attribute<float>(to_verts(x, nv), "coordinates");
    auto v_normals = rx.add_vertex_attribute<float>("v_normals", 3, LOCATION_ALL);
    LaunchBox<blockThreads> launch_box;
    rx.prepare_launch_box({Op::FV}, launch_box, (void*)user_vertex_normal<float, blockThreads>);
    v_normals->reset(0, DEVICE);
    user_vertex_normal<float, blockThreads><<<launch_box.blocks, launch_box.num_threads, launch_box.smem_bytes_dyn>>>(
        rx.get_context(), *coords, *v_normals);
    if (cudaDeviceSynchronize() != cudaSuccess) return 1;
    v_normals->move(DEVICE, HOST);
    rx.for_each_vertex(HOST, [&](const VertexHandle& vh) {
        const uint32_t v_id = rx.map_to_global(vh);
        for (uint32_t i = 0; i < 3; ++i)
            out[v_id * 3 + i] = (*v_normals)(vh, i);
    });
    return 0;
}

// timing of the SAME user kernel (reference call sites, our headers): cudaEvent pair around nrun launches after one
// warm-up, output zeroed outside the timed region as the app does (apps/VertexNormal/vertex_normal.cu:69-85).
// face_patch (optional) replays a given patching, e.g. the reference's own.
static int app_time_vertex_normals(const uint32_t* fv, uint32_t nf, const float* x, uint32_t nv, const uint32_t* face_patch,
                                   uint32_t patch_size, int nrun, float* ms_out)
{
    rx_init(0);
    std::vector<uint32_t> fp;
    if (face_patch) fp.assign(face_patch, face_patch + nf);
    RXMeshStatic rx(fv, nf, fp, patch_size);
    constexpr uint32_t blockThreads = 256;
    auto coords    = rx.add_vertex_attribute<float>("coordinates", 3, LOCATION_ALL);
    auto v_normals = rx.add_vertex_attribute<float>("v_normals", 3, LOCATION_ALL);
    rx.for_each_vertex(HOST, [&](const VertexHandle& vh) {
        const uint32_t v_id = rx.map_to_global(vh);
        for (uint32_t i = 0; i < 3; ++i)
            (*coords)(vh, i) = x[v_id * 3 + i];
    }, NULL, false);
    coords->move(HOST, DEVICE);
    LaunchBox<blockThreads> launch_box;
    rx.prepare_launch_box({Op::FV}, launch_box, (void*)user_vertex_normal<float, blockThreads>);
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    float total = 0;
    for (int i = 0; i <= nrun; ++i) {
        v_normals->reset(0, DEVICE);
        cudaEventRecord(a);
        user_vertex_normal<float, blockThreads><<<launch_box.blocks, launch_box.num_threads, launch_box.smem_bytes_dyn>>>(
            rx.get_context(), *coords, *v_normals);
        cudaEventRecord(b);
        if (cudaEventSynchronize(b) != cudaSuccess) return 1;
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (i) total += ms;
    }
    *ms_out = total / nrun;
    return 0;
}

// host + device API surface that the apps above do not touch: OBJ constructor +
// get_input_vertex_coordinates, add_*_attribute_like, vector<T> attribute constructors, run_kernel (two overloads),
// Query::prologue / get_iterator / run_compute / epilogue, compute_vertex_valence, device for_each<Op::F>,
// Context::get_owner_handle (device) and RXMeshStatic::get_owner_handle (host), get_boundary_vertices, export_obj.
// out[0..nv) = valence from the split API, [nv..2nv) = vertex_valence, [2nv..3nv) = #incident faces,
// [3nv..4nv) = 3 * valence through run_kernel, [4nv..5nv) = boundary flag; returns a bit mask of failed host checks
static int app_api_surface(const char* obj_path, const char* export_path, uint32_t patch_size, float* out)
{
    rx_init(0);
    RXMeshStatic rx(std::string(obj_path), "", patch_size);
    constexpr uint32_t blockThreads = 256;
    const uint32_t nv = rx.get_num_vertices(), nf = rx.get_num_faces();
    int bad = 0;
    auto coords = rx.get_input_vertex_coordinates();
    auto a      = rx.add_vertex_attribute<int>("deg_split", 1, LOCATION_ALL);
    auto b      = rx.add_vertex_attribute_like<int>("deg_val", *a);
    auto c      = rx.add_vertex_attribute_like<int>("face_cnt", *a);
    auto ok     = rx.add_vertex_attribute_like<int>("owner_ok", *a);
    auto t      = rx.add_face_attribute<int>(std::vector<int>(nf, 0), "touched");
    auto sv     = rx.add_vertex_attribute<float>("scaled", 1, LOCATION_ALL);
    auto bd     = rx.add_vertex_attribute<int>("boundary", 1, LOCATION_ALL);
    if (b->get_num_attributes() != 1 || b->get_layout() != a->get_layout() || !rx.does_attribute_exist("face_cnt")) bad |= 1;
    a->reset(-1, DEVICE), b->reset(-1, DEVICE), c->reset(-1, DEVICE), ok->reset(1, DEVICE), sv->reset(0.f, DEVICE);
    LaunchBox<blockThreads> lb;
    rx.prepare_launch_box({Op::VV, Op::VF}, lb, (void*)user_split_api<blockThreads>, false, true);
    rx.run_kernel(lb, user_split_api<blockThreads>, *a, *b, *c, *t, *ok);
    rx.run_kernel<blockThreads>({Op::VV}, user_scaled_valence<blockThreads>, *sv, 3.0f);
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    rx.get_boundary_vertices(*bd);
    a->move(DEVICE, HOST), b->move(DEVICE, HOST), c->move(DEVICE, HOST), t->move(DEVICE, HOST), ok->move(DEVICE, HOST);
    sv->move(DEVICE, HOST);
    rx.for_each_vertex(HOST, [&](const VertexHandle& vh) {
        const uint32_t g = rx.map_to_global(vh);
        out[g] = (float)(*a)(vh), out[nv + g] = (float)(*b)(vh), out[2 * nv + g] = (float)(*c)(vh);
        out[3 * nv + g] = (*sv)(vh), out[4 * nv + g] = (float)(*bd)(vh);
        if ((*ok)(vh) != 1) bad |= 2;
        if (!(rx.get_owner_handle(vh) == vh)) bad |= 4;  // an owned handle is its own owner
    }, NULL, false);
    rx.for_each_face(HOST, [&](const FaceHandle& fh) { if ((*t)(fh) != 1) bad |= 8; }, NULL, false);
    rx.export_obj(export_path, *coords);
    if (rx.get_attribute_names().size() < 8) bad |= 16;
    return bad;
}

// TEST(RXMeshStatic, MultiQueries) (tests/RXMesh_test/test_multi_queries.cu:94-153): both sums per vertex, global order
static int app_multi_queries(const uint32_t* fv, uint32_t nf, const float* x, uint32_t nv, uint32_t patch_size, float* out_ev,
                             float* out_multi)
{
    rx_init(0);
    RXMeshStatic rx(to_faces(fv, nf), "", patch_size);
    constexpr uint32_t blockThreads = 320;
    auto coords = rx.add_vertex_attribute<float>(to_verts(x, nv), "coords");
    auto a      = rx.add_vertex_attribute<float>("sum_ev", 1, LOCATION_ALL);
    auto b      = rx.add_vertex_attribute<float>("sum_multi", 1, LOCATION_ALL);
    a->reset(0, DEVICE), b->reset(0, DEVICE);
    rx.run_kernel<blockThreads>({Op::EV}, user_sum_edges_ev<blockThreads, float>, *coords, *a);
    // two queries live at the same time: is_concurrent = true (test_multi_queries.cu:128-136)
    rx.run_kernel<blockThreads>(user_sum_edges_multi_queries<blockThreads, float>, {Op::EV, Op::VE}, false, false, true,
                                [](uint32_t, uint32_t, uint32_t) { return 0; }, NULL, *coords, *b);
    if (cudaDeviceSynchronize() != cudaSuccess) return 1;
    a->move(DEVICE, HOST), b->move(DEVICE, HOST);
    rx.for_each_vertex(HOST, [&](const VertexHandle& vh) {
        out_ev[rx.map_to_global(vh)]    = (*a)(vh);
        out_multi[rx.map_to_global(vh)] = (*b)(vh);
    }, NULL, false);
    return 0;
}

// TEST(RXMeshStatic, DISABLED_HigherQueries) (tests/RXMesh_test/test_higher_queries.cu:8-52): 2-ring VV; out_global is
// [nv][width] global ids (0xFFFFFFFF = none)
static int app_higher_query(const uint32_t* fv, uint32_t nf, uint32_t patch_size, uint32_t width, uint32_t* out_global)
{
    rx_init(0);
    RXMeshStatic rx(to_faces(fv, nf), "", patch_size);
    constexpr uint32_t blockThreads = 512;  // one vertex of the patch per thread (see app_filtering)
    auto input  = rx.add_vertex_attribute<VertexHandle>("input", 1);
    auto output = rx.add_vertex_attribute<VertexHandle>("output", width);
    input->reset(VertexHandle(), DEVICE);
    output->reset(VertexHandle(), DEVICE);
    LaunchBox<blockThreads> lb;
    rx.prepare_launch_box({Op::VV}, lb, (void*)user_higher_query<blockThreads, Op::VV>);
    user_higher_query<blockThreads, Op::VV><<<lb.blocks, blockThreads, lb.smem_bytes_dyn>>>(rx.get_context(), *input, *output);
    if (cudaDeviceSynchronize() != cudaSuccess) return 1;
    input->move(DEVICE, HOST), output->move(DEVICE, HOST);
    int bad = 0;
    rx.for_each_vertex(HOST, [&](const VertexHandle& vh) {
        if ((*input)(vh) != vh) bad = 1;
        for (uint32_t i = 0; i < width; ++i) {
            const VertexHandle o = (*output)(vh, i);
            out_global[(size_t)rx.map_to_global(vh) * width + i] = o.is_valid() ? rx.map_to_global(o) : 0xFFFFFFFFu;
        }
    }, NULL, false);
    return bad ? -1 : 0;
}

// TEST(RXMeshStatic, Indices) (tests/RXMesh_test/test_indices.cu:11-63): linear_id(handle) <-> get_handle(i) on the device
template <typename HandleT>
static int indices_round_trip(RXMeshStatic& rx)
{
    const uint32_t size = rx.get_num_elements<HandleT>();
    HandleT*       handles = nullptr;
    int*           d_bad   = nullptr;
    if (cudaMalloc((void**)&handles, sizeof(HandleT) * size) != cudaSuccess || cudaMalloc((void**)&d_bad, 4) != cudaSuccess) return 1;
    cudaMemset(handles, 0xFF, sizeof(HandleT) * size);
    cudaMemset(d_bad, 0, 4);
    auto ctx = rx.get_context();
    rx.for_each<HandleT>(DEVICE, [=] __device__(const HandleT h) { handles[ctx.template linear_id<HandleT>(h)] = h; });
    rx.for_each<HandleT>(DEVICE, [=] __device__(const HandleT h) {
        const uint32_t i = ctx.template linear_id<HandleT>(h);
        if (i >= size || ctx.template get_handle<HandleT>(i) != handles[i] || handles[i] != h) *d_bad = 1;
    });
    int bad = 1;
    if (cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost) != cudaSuccess) bad = 1;
    cudaFree(handles), cudaFree(d_bad);
    return bad;
}
static int app_indices(const uint32_t* fv, uint32_t nf, uint32_t patch_size)
{
    rx_init(0, 2);  // (device, log level) as in the reference's mains
    auto         faces = to_faces(fv, nf);
    RXMeshStatic rx(faces, "", patch_size, 1.0, 1.0, 0.8);  // the reference's full constructor signature
    int bad = indices_round_trip<VertexHandle>(rx) | (indices_round_trip<EdgeHandle>(rx) << 1) | (indices_round_trip<FaceHandle>(rx) << 2);
    // the host side of the same maps: map_to_local_*(linear_id(h)) == h, per-patch counts, get_edge_id, patch statistics
    rx.for_each_vertex(HOST, [&](const VertexHandle h) { if (rx.map_to_local_vertex(rx.linear_id(h)) != h) bad |= 8; }, NULL, false);
    rx.for_each_edge(HOST, [&](const EdgeHandle h) { if (rx.map_to_local_edge(rx.linear_id(h)) != h) bad |= 8; }, NULL, false);
    rx.for_each_face(HOST, [&](const FaceHandle h) { if (rx.map_to_local_face(rx.linear_id(h)) != h) bad |= 8; }, NULL, false);
    uint32_t ov = 0, oe = 0, of = 0, lf = 0;
    for (uint32_t p = 0; p < rx.get_num_patches(); ++p) {
        ov += rx.get_num_owned_vertices(p), oe += rx.get_num_owned_edges(p), of += rx.get_num_owned_faces(p), lf += rx.get_num_faces(p);
        if (rx.get_num_vertices(p) < rx.get_num_owned_vertices(p) || rx.get_element_prefix<FaceHandle>(HOST)[p + 1] != of) bad |= 16;
    }
    if (ov != rx.get_num_vertices() || oe != rx.get_num_edges() || of != rx.get_num_faces()) bad |= 16;
    uint32_t mn, mx, avg;
    rx.get_max_min_avg_patch_size(mn, mx, avg);
    if (mn > avg || avg > mx || mx != rx.get_per_patch_max_faces() || avg != (uint32_t)((float)lf / (float)rx.get_num_patches())) bad |= 32;
    if (std::fabs(rx.get_ribbon_overhead() - 100.0 * (double(lf) - nf) / nf) > 1e-9 || rx.get_max_num_patches() != rx.get_num_patches()) bad |= 32;
    // get_edge_id: every face's three vertex pairs name an edge, the numbering is the order of first appearance
    uint32_t next = 0;
    for (uint32_t f = 0; f < nf; ++f)
        for (int j = 0; j < 3; ++j) {
            const uint32_t e = rx.get_edge_id(fv[3 * f + j], fv[3 * f + (j + 1) % 3]);
            if (e == INVALID32 || e > next || e != rx.get_edge_id(fv[3 * f + (j + 1) % 3], fv[3 * f + j])) bad |= 64;
            if (e == next) ++next;
        }
    if (next != rx.get_num_edges() || rx.get_edge_id(fv[0], fv[0]) != INVALID32) bad |= 64;
    return bad;
}

// tests/RXMesh_test/test_attribute.cu restated as one user program: Norm2 / Dot / Reduce / ArgMax / CopyFrom /
// AddingAndRemoving / DefaultLayoutIsAoSoA / TrueSoAHostStorageIsColumnMajor / TrueSoADeviceWritesColumnMajor /
// ResetSetsAllComponents.  Returns a bit mask of the checks that FAILED (0 = all passed).
struct UserCustomMin
{
    template <typename T>
    __device__ __forceinline__ T operator()(const T& a, const T& b) const { return (b < a) ? b : a; }
};
static int app_attribute_tests(const uint32_t* fv, uint32_t nf, uint32_t patch_size)
{
    rx_init(0);
    RXMeshStatic   rx(to_faces(fv, nf), "", patch_size);
    int            failed = 0;
    const uint32_t n = rx.get_num_vertices();
    auto           near = [](double a, double b) { return std::fabs(a - b) <= 4e-7 * std::fabs(b); };  // EXPECT_FLOAT_EQ: 4 ulp
    {   // Norm2 (test_attribute.cu:56-79)
        auto        attr = rx.add_vertex_attribute<float>("v", 3, DEVICE);
        const float val  = 2.0f;
        auto        a    = *attr;
        rx.for_each_vertex(DEVICE, [a, val] __device__(const VertexHandle vh) { a(vh, 0) = val, a(vh, 1) = val, a(vh, 2) = val; });
        ReduceHandle reduce_handle(*attr);
        const float  out = reduce_handle.norm2(*attr);
        if (cudaDeviceSynchronize() != cudaSuccess || !near(out, std::sqrt(3.0 * val * val * n))) failed |= 1;
        // one component only: the test's own expectation, sqrt(val^2 * #vertices)
        if (!near(reduce_handle.norm2(*attr, 1), std::sqrt((double)val * val * n))) failed |= 1;
        rx.remove_attribute("v");
    }
    {   // Dot (test_attribute.cu:82-103)
        auto        v1 = rx.add_vertex_attribute<float>("v1", 3, DEVICE);
        auto        v2 = rx.add_vertex_attribute<float>("v2", 3, DEVICE);
        const float a1 = 2.0f, a2 = 3.0f;
        auto        x = *v1, y = *v2;
        rx.for_each_vertex(DEVICE, [x, y, a1, a2] __device__(const VertexHandle vh) {
            for (int c = 0; c < 3; ++c) x(vh, c) = a1, y(vh, c) = a2;
        });
        ReduceHandle reduce_handle(*v1);
        if (!near(reduce_handle.dot(*v1, *v2), 3.0 * a1 * a2 * n)) failed |= 2;
        if (!near(reduce_handle.dot(*v1, *v2, 2), (double)a1 * a2 * n)) failed |= 2;
    }
    {   // Reduce (test_attribute.cu:105-138): max over the edges of patch id * local id, cub::Max and a user functor
        auto attr = rx.add_edge_attribute<uint32_t>("e", 3, DEVICE);
        auto e    = *attr;
        rx.for_each_edge(DEVICE, [e] __device__(const EdgeHandle eh) {
            auto pl = eh.unpack();
            for (int c = 0; c < 3; ++c) e(eh, c) = pl.first * pl.second + 1u;
        });
        ReduceHandle   reduce_handle(*attr);
        const uint32_t out_max = reduce_handle.reduce(*attr, cub::Max(), 0u);
        const uint32_t out_min = reduce_handle.reduce(*attr, UserCustomMin(), 0xFFFFFFFFu);
        uint32_t       result = 0, result_min = 0xFFFFFFFFu;
        rx.for_each_edge(HOST, [&](const EdgeHandle eh) {
            auto pl    = eh.unpack();
            result     = std::max(result, pl.first * pl.second + 1u);
            result_min = std::min(result_min, pl.first * pl.second + 1u);
        }, NULL, false);
        if (cudaDeviceSynchronize() != cudaSuccess || out_max != result || out_min != result_min) failed |= 4;
    }
    {   // ArgMax (test_attribute.cu:141-182) and its mirror ArgMin
        auto        attr = *rx.add_vertex_attribute<float>("am", 1);
        const float val  = 2.0f;
        rx.for_each_vertex(DEVICE, [attr, val] __device__(const VertexHandle vh) { attr(vh) = val; });
        attr.move(DEVICE, HOST);
        const uint32_t chosen = n - 1, chosen_min = n / 2;
        VertexHandle   chosenHandle, chosenMinHandle;
        rx.for_each_vertex(HOST, [&](const VertexHandle vh) {
            if (rx.linear_id(vh) == chosen) attr(vh) = 10.f, chosenHandle = vh;
            if (rx.linear_id(vh) == chosen_min) attr(vh) = -3.f, chosenMinHandle = vh;
        }, NULL, false);
        attr.move(HOST, DEVICE);
        ReduceHandle reduce_handle(attr);
        auto         mx = reduce_handle.arg_max(attr);
        auto         mn = reduce_handle.arg_min(attr, 0);
        if (!chosenHandle.is_valid() || !(mx.key == chosenHandle) || mx.value != 10.f) failed |= 8;
        if (!(mn.key == chosenMinHandle) || mn.value != -3.f) failed |= 8;
    }
    {   // CopyFrom (test_attribute.cu:185-203)
        auto           f_device = rx.add_face_attribute<uint32_t>("d", 3, DEVICE);
        auto           f_host   = rx.add_face_attribute<uint32_t>("h", 3, HOST);
        const uint32_t val      = 99;
        auto           fd       = *f_device;
        rx.for_each_face(DEVICE, [fd, val] __device__(const FaceHandle fh) { fd(fh, 0) = val, fd(fh, 1) = val, fd(fh, 2) = val; });
        f_host->copy_from(*f_device, DEVICE, HOST);
        rx.for_each_face(HOST, [&](const FaceHandle fh) {
            for (int c = 0; c < 3; ++c)
                if ((*f_host)(fh, c) != val) failed |= 16;
        }, NULL, false);
    }
    {   // AddingAndRemoving (test_attribute.cu:205-226)
        auto vertex_attr = rx.add_vertex_attribute<float>("v_attr", 3, LOCATION_ALL);
        if (!rx.does_attribute_exist("v_attr")) failed |= 32;
        vertex_attr->move(HOST, DEVICE);
        if (cudaDeviceSynchronize() != cudaSuccess) failed |= 32;
        rx.remove_attribute("v_attr");
        if (rx.does_attribute_exist("v_attr")) failed |= 32;
    }
    {   // DefaultLayoutIsAoSoA (test_attribute.cu:228-238)
        auto attr = rx.add_vertex_attribute<float>("default_layout", 3, HOST);
        if (attr->get_layout() != AoSoA || layout_to_string(attr->get_layout()) != "AoSoA") failed |= 64;
    }
    {   // TrueSoAHostStorageIsColumnMajor (test_attribute.cu:240-279)
        auto attr = rx.add_vertex_attribute<float>("true_soa", 3, HOST, SoA);
        if (!attr->is_tensor_layout() || attr->storage_size() != size_t(n) * attr->get_num_attributes() || attr->rows() != n ||
            attr->cols() != 3)
            failed |= 128;
        float* data = attr->data(HOST);
        for (uint32_t c = 0; c < attr->get_num_attributes(); ++c)
            for (uint32_t i = 0; i < n; ++i)
                data[c * n + i] = float(c * 1000 + i);
        rx.for_each_vertex(HOST, [&](const VertexHandle vh) {
            const uint32_t row = rx.linear_id(vh);
            for (uint32_t c = 0; c < attr->get_num_attributes(); ++c)
                if ((*attr)(vh, c) != float(c * 1000 + row) || (*attr)(row, c) != float(c * 1000 + row)) failed |= 128;
        }, NULL, false);
        // operator()(row, col) names the same element in the slot-ordered layouts
        for (layoutT layout : {AoS, AoSoA}) {
            auto other = rx.add_vertex_attribute<float>("rows_" + layout_to_string(layout), 3, HOST, layout);
            rx.for_each_vertex(HOST, [&](const VertexHandle vh) {
                for (uint32_t c = 0; c < 3; ++c) (*other)(vh, c) = float(c * 1000 + rx.linear_id(vh));
            }, NULL, false);
            for (uint32_t row = 0; row < n; row += 7)
                for (uint32_t c = 0; c < 3; ++c)
                    if ((*other)(row, c) != float(c * 1000 + row)) failed |= 128;
        }
    }
    {   // TrueSoADeviceWritesColumnMajor (test_attribute.cu:281-303)
        auto attr = rx.add_vertex_attribute<float>("true_soa_device", 3, LOCATION_ALL, SoA);
        auto a    = *attr;
        rx.for_each_vertex(DEVICE, [=] __device__(const VertexHandle vh) mutable { a(vh, 0) = 11.0f, a(vh, 1) = 22.0f, a(vh, 2) = 33.0f; });
        if (cudaDeviceSynchronize() != cudaSuccess) failed |= 256;
        attr->move(DEVICE, HOST);
        const float* data = attr->data(HOST);
        for (uint32_t i = 0; i < n; ++i)
            if (data[i] != 11.0f || data[n + i] != 22.0f || data[2 * n + i] != 33.0f) failed |= 256;
    }
    {   // ResetSetsAllComponents (test_attribute.cu:305-329), host and device side
        for (layoutT layout : {AoS, AoSoA, SoA}) {
            auto attr = rx.add_vertex_attribute<float>("reset_" + layout_to_string(layout), 3, LOCATION_ALL, layout);
            attr->reset(5.0f, HOST);
            rx.for_each_vertex(HOST, [&](const VertexHandle vh) {
                for (uint32_t c = 0; c < attr->get_num_attributes(); ++c)
                    if ((*attr)(vh, c) != 5.0f) failed |= 512;
            }, NULL, false);
            attr->reset(7.0f, DEVICE);
            attr->move(DEVICE, HOST);
            rx.for_each_vertex(HOST, [&](const VertexHandle vh) {
                for (uint32_t c = 0; c < attr->get_num_attributes(); ++c)
                    if ((*attr)(vh, c) != 7.0f) failed |= 512;
            }, NULL, false);
        }
    }
    {   // the fixed-function path on a tensor-layout attribute: get_boundary_vertices writes flags by linear id
        auto flags = rx.add_vertex_attribute<int>("bd_soa", 1, LOCATION_ALL, SoA);
        auto ref   = rx.add_vertex_attribute<int>("bd_ref", 1, LOCATION_ALL, AoS);
        rx.get_boundary_vertices(*flags);
        rx.get_boundary_vertices(*ref);
        rx.for_each_vertex(HOST, [&](const VertexHandle vh) {
            if ((*flags)(vh) != (*ref)(vh) || flags->data(HOST)[rx.linear_id(vh)] != (*ref)(vh)) failed |= 1024;
        }, NULL, false);
    }
    return failed;
}

// TEST(RXMeshStatic, MultipleMeshes) + TEST(RXMeshStatic, Export) (tests/RXMesh_test/test_multiple_meshes.cu,
// test_export.cu): several OBJ files as one mesh with region labels, bounding_box / scale, export_obj / export_vtk.
// Returns a bit mask of failed checks.
static int app_multiple_meshes(const char* path_a, const char* path_b, const char* out_obj, const char* out_vtk)
{
    rx_init(0);
    int failed = 0;
    const std::string file_a = path_a, file_b = path_b;
    RXMeshStatic      ra(file_a), rb(file_b);  // the parts on their own, for the expected counts
    const uint32_t nva = ra.get_num_vertices(), nfa = ra.get_num_faces();
    std::vector<std::string> inputs = {path_a, path_b};
    RXMeshStatic             rx(inputs);
    if (rx.get_num_regions() != 2 || rx.get_num_vertices() != nva + rb.get_num_vertices() ||
        rx.get_num_faces() != nfa + rb.get_num_faces() || rx.get_num_edges() != ra.get_num_edges() + rb.get_num_edges())
        failed |= 1;
    auto x       = *rx.get_input_vertex_coordinates();
    auto v_label = *rx.get_vertex_region_label();
    auto f_label = *rx.get_face_region_label();
    auto e_label = *rx.get_edge_region_label();
    rx.for_each_vertex(HOST, [&](const VertexHandle vh) {
        if (v_label(vh) != (rx.map_to_global(vh) < nva ? 0 : 1)) failed |= 2;
    }, NULL, false);
    rx.for_each_face(HOST, [&](const FaceHandle fh) {
        if (f_label(fh) != (rx.map_to_global(fh) < nfa ? 0 : 1)) failed |= 4;
    }, NULL, false);
    // an edge lies in the region of its end vertices: EV on the device against the host copy of the edge labels
    auto bad = rx.add_edge_attribute<int>("bad", 1, LOCATION_ALL);
    bad->reset(0, DEVICE);
    auto badv = *bad;
    rx.for_each<Op::EV, 256>([=] __device__(const EdgeHandle eh, const VertexIterator& iter) {
        if (e_label(eh) != v_label(iter[0]) || e_label(eh) != v_label(iter[1])) badv(eh) = 1;
    });
    if (cudaDeviceSynchronize() != cudaSuccess) failed |= 8;
    bad->move(DEVICE, HOST);
    rx.for_each_edge(HOST, [&](const EdgeHandle eh) {
        if ((*bad)(eh) != 0 || (e_label(eh) != 0 && e_label(eh) != 1)) failed |= 8;
    }, NULL, false);
    // bounding_box, then the test's own move of region i along axis i % 3, then scale into the unit cube
    glm::vec3 lower, upper;
    rx.bounding_box(lower, upper);
    glm::vec3 lo2(1e30f), up2(-1e30f);
    rx.for_each_vertex(HOST, [&](const VertexHandle vh) {
        for (int i = 0; i < 3; ++i) lo2[i] = std::min(lo2[i], x(vh, i)), up2[i] = std::max(up2[i], x(vh, i));
    }, NULL, false);
    for (int i = 0; i < 3; ++i)
        if (lower[i] != lo2[i] || upper[i] != up2[i] || !(lower[i] < upper[i])) failed |= 16;
    const glm::vec3 bb = upper - lower;
    for (int i = 0; i < rx.get_num_regions(); ++i)
        rx.for_each_vertex(HOST, [&](const VertexHandle vh) {
            const int j = i % 3;
            if (v_label(vh) == i) x(vh, j) += 0.5f * (i + 1) * bb[j];
        }, NULL, false);
    rx.scale(glm::fvec3(0.f, 0.f, 0.f), glm::fvec3(1.f, 1.f, 1.f));
    rx.bounding_box(lower, upper);
    for (int i = 0; i < 3; ++i)
        if (lower[i] < -1e-5f || upper[i] > 1.f + 1e-5f) failed |= 32;
    if (std::max(upper[0], std::max(upper[1], upper[2])) < 0.999f) failed |= 32;  // the longest side fills the cube
    // Export: scalar / 2- / 3-component vertex and face attributes tied to the exported coordinates
    auto vs = *rx.add_vertex_attribute<float>("vScalar", 1);
    auto v2 = *rx.add_vertex_attribute<float>("vVector2", 2);
    auto v3 = *rx.add_vertex_attribute<float>("vVector3", 3);
    auto fs = *rx.add_face_attribute<float>("fScalar", 1);
    auto f3 = *rx.add_face_attribute<float>("fVector3", 3);
    rx.for_each_vertex(HOST, [&](const VertexHandle& vh) {
        vs(vh, 0) = 2.f * x(vh, 0);
        v2(vh, 0) = x(vh, 1), v2(vh, 1) = x(vh, 2);
        for (uint32_t i = 0; i < 3; ++i) v3(vh, i) = -x(vh, i);
    }, NULL, false);
    rx.for_each_face(HOST, [&](const FaceHandle& fh) {
        fs(fh, 0) = (float)f_label(fh);
        for (uint32_t i = 0; i < 3; ++i) f3(fh, i) = (float)(rx.linear_id(fh) + i);
    }, NULL, false);
    rx.export_obj(out_obj, x);
    rx.export_vtk(out_vtk, x, vs, v2, v3, fs, f3);
    if (cudaDeviceSynchronize() != cudaSuccess) failed |= 64;
    return failed;
}

// the Filtering driver loop (apps/Filtering/filtering_rxmesh.cuh:60-100)
static int app_filtering(const uint32_t* fv, uint32_t nf, const float* x, uint32_t nv, uint32_t patch_size, int num_iter,
                         float* out)
{
    rx_init(0);
    RXMeshStatic rx(to_faces(fv, nf), "", patch_size);
    // the kernel keeps ONE vertex's state per thread (as the reference's does), so a block must cover every owned vertex
    // of its patch: 512-face patches own up to ~300 vertices -> 512 threads (the reference app launches 256)
    constexpr uint32_t blockThreads = 512, maxVVSize = 80;
    auto coords   = rx.add_vertex_attribute<float>(to_verts(x, nv), "coords");
    auto filtered = rx.add_vertex_attribute<float>("filtered", 3, LOCATION_ALL);
    auto normals  = rx.add_vertex_attribute<float>("vn", 3, LOCATION_ALL);
    LaunchBox<blockThreads> lb_vn, lb_f;
    rx.prepare_launch_box({Op::FV}, lb_vn, (void*)user_filter_vertex_normal<float, blockThreads>);
    rx.prepare_launch_box({Op::VV}, lb_f, (void*)user_bilateral_filtering<float, blockThreads, maxVVSize>);
    VertexAttribute<float>* a = coords.get();
    VertexAttribute<float>* b = filtered.get();
    for (int it = 0; it < num_iter; ++it) {
        normals->reset(0, DEVICE);
        user_filter_vertex_normal<float, blockThreads><<<lb_vn.blocks, blockThreads, lb_vn.smem_bytes_dyn>>>(rx.get_context(), *a, *normals);
        user_bilateral_filtering<float, blockThreads, maxVVSize><<<lb_f.blocks, blockThreads, lb_f.smem_bytes_dyn>>>(
            rx.get_context(), *a, *b, *normals);
        std::swap(a, b);
    }
    if (cudaDeviceSynchronize() != cudaSuccess) return 1;
    a->move(DEVICE, HOST);
    rx.for_each_vertex(HOST, [&](const VertexHandle& vh) {
        const uint32_t v_id = rx.map_to_global(vh);
        for (uint32_t i = 0; i < 3; ++i)
            out[v_id * 3 + i] = (*a)(vh, i);
    }, NULL, false);
    return 0;
}

// manual smoothing (apps/Smoothing/manual.h:86-104): for_each<Op::VV> gradient + for_each_vertex(DEVICE) step
static int app_smoothing(const uint32_t* fv, uint32_t nf, const float* x, uint32_t nv, uint32_t patch_size, double lr,
                         int num_iter, int oriented, float* out)
{
    rx_init(0);
    RXMeshStatic rx(to_faces(fv, nf), "", patch_size);
    constexpr int blockThreads = 256;
    auto pos  = *rx.add_vertex_attribute<float>(to_verts(x, nv), "pos");
    auto grad = *rx.add_vertex_attribute<float>("grad", 3, LOCATION_ALL);
    const int cols = pos.get_num_attributes();
    for (int iter = 0; iter < num_iter; ++iter) {
        grad.reset(0, DEVICE);
        rx.for_each<Op::VV, blockThreads>(
            [=] __device__(const VertexHandle& vh, const VertexIterator& iter) mutable {
                for (int v = 0; v < iter.size(); ++v) {
                    const VertexHandle uh = iter[v];
                    for (int i = 0; i < cols; ++i)
                        grad(vh, i) += 2 * (pos(vh, i) - pos(uh, i));
                }
            },
            oriented != 0);
        rx.for_each_vertex(DEVICE, [grad, pos, lr, cols] __device__(const VertexHandle& vh) {
            for (int i = 0; i < cols; ++i)
                pos(vh, i) -= lr * grad(vh, i);
        });
    }
    if (cudaDeviceSynchronize() != cudaSuccess) return 1;
    pos.move(DEVICE, HOST);
    rx.for_each_vertex(HOST, [&](const VertexHandle& vh) {
        const uint32_t v_id = rx.map_to_global(vh);
        for (uint32_t i = 0; i < 3; ++i)
            out[v_id * 3 + i] = pos(vh, i);
    });
    return 0;
}

static int app_query(int op, const uint32_t* fv, uint32_t nf, uint32_t patch_size, uint32_t width, int oriented,
                     uint32_t* out_global)
{
    rx_init(0);
    RXMeshStatic rx(to_faces(fv, nf), "", patch_size);
    switch ((Op)op) {
        case Op::VV: return run_query<Op::VV, VertexHandle, VertexHandle>(rx, width, oriented, out_global);
        case Op::VE: return run_query<Op::VE, VertexHandle, EdgeHandle>(rx, width, oriented, out_global);
        case Op::VF: return run_query<Op::VF, VertexHandle, FaceHandle>(rx, width, oriented, out_global);
        case Op::EV: return run_query<Op::EV, EdgeHandle, VertexHandle>(rx, width, oriented, out_global);
        case Op::EF: return run_query<Op::EF, EdgeHandle, FaceHandle>(rx, width, oriented, out_global);
        case Op::FV: return run_query<Op::FV, FaceHandle, VertexHandle>(rx, width, oriented, out_global);
        case Op::FE: return run_query<Op::FE, FaceHandle, EdgeHandle>(rx, width, oriented, out_global);
        case Op::FF: return run_query<Op::FF, FaceHandle, FaceHandle>(rx, width, oriented, out_global);
        case Op::EVDiamond: return run_query<Op::EVDiamond, EdgeHandle, VertexHandle>(rx, width, oriented, out_global);
        case Op::EE: return run_query<Op::EE, EdgeHandle, EdgeHandle>(rx, width, oriented, out_global);
        default: return 2;
    }
}

// The record a reference app writes around its run (apps/VertexNormal/vertex_normal.cu:136-157: Report, command_line, device,
// system, model_data, add_member, TestData + add_test, write), as a user program on the drop-in headers.
static int app_report(const char* obj_path, const char* out_dir, uint32_t patch_size)
{
    rx_init(0);
    RXMeshStatic rx(std::string(obj_path), "", patch_size);
    Report       report("VertexNormal_RXMesh");
    char         a0[] = "shim_apps", a1[] = "-input", a2[] = "in.obj";
    char*        argv[] = {a0, a1, a2};
    report.command_line(3, argv);
    report.device();
    report.system();
    report.model_data("in.obj", rx);
    report.add_member("method", std::string("RXMesh"));
    report.add_member("num_run", 3);
    TestData td;
    td.test_name   = "VertexNormal";
    td.num_threads = 256;
    td.num_blocks  = (int32_t)rx.get_num_patches();
    td.dyn_smem    = 1024;
    td.time_ms     = {0.25f, 0.5f, 1.0f};
    td.passed      = {true, true, false};
    report.add_test(td);
    report.write(out_dir, "record.obj", false);  // -> <out_dir>/record.json
    CustomReport custom("Other");
    custom.model_data("in.obj", rx.get_num_vertices(), rx.get_num_faces());
    custom.write(out_dir, "custom", false);
    return 0;
}

extern "C" {
int shim_report(const char* obj_path, const char* out_dir, uint32_t patch_size)
{
    return app_report(obj_path, out_dir, patch_size);
}
int shim_vertex_normals(const uint32_t* fv, uint32_t nf, const float* x, uint32_t nv, uint32_t patch_size, float* out)
{
    return app_vertex_normals(fv, nf, x, nv, patch_size, out);
}
int shim_time_vertex_normals(const uint32_t* fv, uint32_t nf, const float* x, uint32_t nv, const uint32_t* face_patch,
                             uint32_t patch_size, int nrun, float* ms_out)
{
    return app_time_vertex_normals(fv, nf, x, nv, face_patch, patch_size, nrun, ms_out);
}
int shim_api_surface(const char* obj_path, const char* export_path, uint32_t patch_size, float* out)
{
    return app_api_surface(obj_path, export_path, patch_size, out);
}
int shim_multi_queries(const uint32_t* fv, uint32_t nf, const float* x, uint32_t nv, uint32_t patch_size, float* out_ev,
                       float* out_multi)
{
    return app_multi_queries(fv, nf, x, nv, patch_size, out_ev, out_multi);
}
int shim_higher_query(const uint32_t* fv, uint32_t nf, uint32_t patch_size, uint32_t width, uint32_t* out_global)
{
    return app_higher_query(fv, nf, patch_size, width, out_global);
}
int shim_indices(const uint32_t* fv, uint32_t nf, uint32_t patch_size)
{
    return app_indices(fv, nf, patch_size);
}
int shim_attribute_tests(const uint32_t* fv, uint32_t nf, uint32_t patch_size)
{
    return app_attribute_tests(fv, nf, patch_size);
}
int shim_multiple_meshes(const char* path_a, const char* path_b, const char* out_obj, const char* out_vtk)
{
    return app_multiple_meshes(path_a, path_b, out_obj, out_vtk);
}
int shim_unit_scan(uint32_t* host_a, uint32_t n)  // in place: a[0..n) -> exclusive prefix sums, a[n] = total
{
    uint32_t* d = nullptr;
    if (cudaMalloc((void**)&d, 4 * (size_t)(n + 1)) != cudaSuccess) return 1;
    cudaMemcpy(d, host_a, 4 * (size_t)n, cudaMemcpyHostToDevice);
    unit_block_scan<256><<<1, 256, 4 * (n + 1) + 4 * 40>>>(d, n);
    const int rc = cudaMemcpy(host_a, d, 4 * (size_t)(n + 1), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 1;
    cudaFree(d);
    return rc;
}
int shim_unit_transpose(const uint16_t* host_src, uint32_t rows, uint32_t cols, uint32_t deg, uint32_t* host_off, uint16_t* host_val)
{
    const uint32_t nnz = rows * deg;
    if (nnz > 12u * 256u) return 2;
    uint16_t *d_src = nullptr, *d_val = nullptr;
    uint32_t* d_off = nullptr;
    cudaMalloc((void**)&d_src, 2 * (size_t)nnz), cudaMalloc((void**)&d_val, 2 * (size_t)nnz), cudaMalloc((void**)&d_off, 4 * (size_t)(cols + 1));
    cudaMemcpy(d_src, host_src, 2 * (size_t)nnz, cudaMemcpyHostToDevice);
    unit_block_transpose<256, 12><<<1, 256, 4 * (cols + 1) + 4 * 40 + 2 * nnz + 16>>>(d_src, rows, cols, deg, d_off, d_val);
    int rc = cudaMemcpy(host_off, d_off, 4 * (size_t)(cols + 1), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 1;
    rc |= cudaMemcpy(host_val, d_val, 2 * (size_t)nnz, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 1;
    cudaFree(d_src), cudaFree(d_val), cudaFree(d_off);
    return rc;
}
int shim_filtering(const uint32_t* fv, uint32_t nf, const float* x, uint32_t nv, uint32_t patch_size, int num_iter, float* out)
{
    return app_filtering(fv, nf, x, nv, patch_size, num_iter, out);
}
int shim_smoothing(const uint32_t* fv, uint32_t nf, const float* x, uint32_t nv, uint32_t patch_size, double lr, int num_iter,
                   int oriented, float* out)
{
    return app_smoothing(fv, nf, x, nv, patch_size, lr, num_iter, oriented, out);
}
int shim_valence(const uint32_t* fv, uint32_t nf, uint32_t patch_size, float* out_valence, float* out_plus1)
{
    return app_valence(fv, nf, patch_size, out_valence, out_plus1);
}
int shim_mcf_matvec(const uint32_t* fv, uint32_t nf, const float* x, const float* vin, uint32_t nv, uint32_t patch_size,
                    float time_step, float* out)
{
    return app_mcf_matvec(fv, nf, x, vin, nv, patch_size, time_step, out);
}
int shim_mcf_cg(const uint32_t* fv, uint32_t nf, const float* x, uint32_t nv, uint32_t patch_size, float time_step, int uniform,
                int pcg, int max_iter, float tol_abs, float tol_rel, float* out, float* info)
{
    return app_mcf_cg(fv, nf, x, nv, patch_size, time_step, uniform, pcg, max_iter, tol_abs, tol_rel, out, info);
}
int shim_gaussian_curvature(const uint32_t* fv, uint32_t nf, const float* x, uint32_t nv, uint32_t patch_size, float* out_gcs,
                            float* out_amix)
{
    return app_gaussian_curvature(fv, nf, x, nv, patch_size, out_gcs, out_amix);
}
int shim_query(int op, const uint32_t* fv, uint32_t nf, uint32_t patch_size, uint32_t width, int oriented, uint32_t* out_global)
{
    return app_query(op, fv, nf, patch_size, width, oriented, out_global);
}
}  // extern "C"
