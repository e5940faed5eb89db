// oracle/_ref GPU driver -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Links the reference's OWN, UNMODIFIED hot-path sources (rxmesh.cpp, patcher/patcher.cu, lp_hashtable.cu,
// patch_stash.cu, hash_functions.cu, patch_info.cu, patch_lock.cu, patch_scheduler.cu, query.cu), compiled from
// where they lie under /root/reference (oracle/Makefile target `ref_gpu`), and runs the reference's
// Query<256>::dispatch<op> (query.inl:107-260) for the eight static ops on a mesh file, exactly as its test
// kernel does (tests/RXMesh_test/query_kernel.cuh:13-46), except that the results go to raw arrays instead of
// the reference's Attribute class (attribute.h needs the real Eigen / cuBLAS wrappers, which are not in the image).
//
// Output directory (consumed by tests/golden/make_golden_ref_gpu.py and bench_configs.py):
//   meta.json                  counts, per-op kernel time (cudaEvent, nrun launches after one warm-up -- the
//                              reference's own protocol, tests/RXMesh_test/test_queries.h:61-96), smem, occupancy
//   face_patch.u32             the reference patcher's face -> patch assignment
//   ltog_{v,e,f}.u32 / ltog_off_{v,e,f}.u32 / owned_{v,e,f}.u32   per-patch local->global maps, #owned per patch
//   q_<OP>.u32                 [num_src][width] GLOBAL ids of the output handles, rows in global source order,
//                              columns in the reference's iteration order, 0xFFFFFFFF = none
//   vn.f32                     vertex normals from the reference's FV lambda
//                              (apps/VertexNormal/vertex_normal_kernel.cuh:10-43) over raw AoSoA arrays
//   lap_1.f32, lap_5.f32       positions after 1 and 5 iterations of the reference's manual smoothing
//                              (apps/Smoothing/manual.h:86-104: the VV gradient lambda through Query<256>::dispatch<Op::VV>,
//                              then the step lambda through the reference's own detail::for_each_vertex kernel;
//                              learning rate 0.01 as a double, apps/Smoothing/smoothing.cu:17-18)
//
// usage: ref_gpu_queries <mesh.bin> <outdir> [patch_size=512] [dump=1] [nrun=100]
//   mesh.bin = u32 nv, u32 nf, u32 fv[3 nf], f32 x[3 nv]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "rxmesh/iterator.cuh"
#include "rxmesh/kernels/for_each.cuh"
#include "rxmesh/query.h"
#include "rxmesh/rxmesh.h"
#include "rxmesh/util/bitmask_util.h"

using namespace rxmesh;

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                           \
        }                                                                                      \
    } while (0)

// RXMeshStatic's constructor body minus the attribute container (rxmesh_static.cu:49-64)
struct RefMesh : public RXMesh
{
    RefMesh(std::vector<std::vector<uint32_t>>& fv, uint32_t patch_size) : RXMesh(patch_size)
    {
        this->init(fv, "", 1.0f, 1.0f, 0.8f);
    }
    ~RefMesh() {}
    const std::vector<std::vector<uint32_t>>& ltog(int t) const
    {
        return t == 0 ? m_h_patches_ltog_v : (t == 1 ? m_h_patches_ltog_e : m_h_patches_ltog_f);
    }
    const std::vector<uint16_t>& num_owned(int t) const
    {
        return t == 0 ? m_h_num_owned_v : (t == 1 ? m_h_num_owned_e : m_h_num_owned_f);
    }
    const uint32_t* owned_mask(uint32_t p, int t) const
    {
        const PatchInfo& pi = m_h_patches_info[p];
        return t == 0 ? pi.owned_mask_v : (t == 1 ? pi.owned_mask_e : pi.owned_mask_f);
    }
    std::vector<uint32_t>& face_patch() { return m_patcher->get_face_patch(); }
    uint32_t max_per_patch(int t) const
    {
        return t == 0 ? m_max_vertices_per_patch : (t == 1 ? m_max_edges_per_patch : m_max_faces_per_patch);
    }
    const PatchInfo* device_patches() const { return m_d_patches_info; }
    uint32_t max_valence() const { return m_input_max_valence; }
    uint32_t max_ef() const { return m_input_max_edge_incident_faces; }
    uint32_t max_ff() const { return m_input_max_face_adjacent_faces; }
};

// the reference's test kernel with raw-array sinks: row = patch * cap + local id of the source
template <uint32_t blockThreads, Op op, typename InH, typename OutH>
__global__ static void ref_query_kernel(const Context context, uint64_t* out, uint16_t* out_size, uint32_t width,
                                        uint32_t cap)
{
    auto store_lambda = [&](const InH& id, const Iterator<OutH>& iter) {
        const auto     pl  = id.unpack();
        const size_t   row = (size_t)pl.first * cap + pl.second;
        const uint32_t n   = iter.size();
        out_size[row]      = (uint16_t)n;
        for (uint32_t i = 0; i < n && i < width; ++i)
            out[row * width + i] = iter[i].unique_id();
    };
    auto                block = cooperative_groups::this_thread_block();
    Query<blockThreads> query(context);
    ShmemAllocator      shrd_alloc;
    query.template dispatch<op>(block, shrd_alloc, store_lambda, [](InH) { return true; }, false);
}

// apps/VertexNormal/vertex_normal_kernel.cuh:10-43 with the attribute replaced by raw per-patch AoSoA arrays
// (same address pattern as the reference's default layout: patch slab, attribute-major inside it)
struct RawAttr3
{
    float*   data;
    uint32_t cap;
    __device__ __forceinline__ float& operator()(const VertexHandle& h, uint32_t a) const
    {
        const auto pl = h.unpack();
        return data[((size_t)pl.first * 3 + a) * cap + pl.second];
    }
};
template <uint32_t blockThreads>
__global__ static void ref_vertex_normal_kernel(const Context context, RawAttr3 coords, RawAttr3 normals)
{
    auto vn_lambda = [&](FaceHandle face_id, VertexIterator& fv) {
        float c[3][3];
        for (int i = 0; i < 3; ++i)
            for (int k = 0; k < 3; ++k)
                c[i][k] = coords(fv[i], k);
        float e1[3] = {c[1][0] - c[0][0], c[1][1] - c[0][1], c[1][2] - c[0][2]};
        float e2[3] = {c[2][0] - c[0][0], c[2][1] - c[0][1], c[2][2] - c[0][2]};
        float n[3]  = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
        float l[3];
        for (int i = 0; i < 3; ++i) {
            const int j = (i + 1) % 3;
            l[i]        = 0;
            for (int k = 0; k < 3; ++k)
                l[i] += (c[i][k] - c[j][k]) * (c[i][k] - c[j][k]);
        }
        for (uint32_t v = 0; v < 3; ++v)
            for (uint32_t i = 0; i < 3; ++i)
                atomicAdd(&normals(fv[v], i), n[i] / (l[v] + l[(v + 2) % 3]));
    };
    auto                block = cooperative_groups::this_thread_block();
    Query<blockThreads> query(context);
    ShmemAllocator      shrd_alloc;
    query.template dispatch<Op::FV>(block, shrd_alloc, vn_lambda);
}

// apps/Smoothing/manual.h:86-95, the gradient lambda of the non-area branch, verbatim in form: grad(vh, i) += 2 * (pos(vh, i) -
// pos(uh, i)) over the VV iterator (the attribute is the raw AoSoA array; the DenseMatrix `grad` of the app is one too)
template <uint32_t blockThreads>
__global__ static void ref_smoothing_grad_kernel(const Context context, RawAttr3 pos, RawAttr3 grad)
{
    constexpr int cols = 3;
    auto          grad_lambda = [&](const VertexHandle& vh, const VertexIterator& iter) {
        for (int v = 0; v < iter.size(); ++v) {
            const VertexHandle uh = iter[v];
            for (int i = 0; i < cols; ++i) {
                grad(vh, i) += 2 * (pos(vh, i) - pos(uh, i));
            }
        }
    };
    auto                block = cooperative_groups::this_thread_block();
    Query<blockThreads> query(context);
    ShmemAllocator      shrd_alloc;
    query.template dispatch<Op::VV>(block, shrd_alloc, grad_lambda);
}
// apps/Smoothing/manual.h:97-103, the step lambda: pos(vh, i) -= lr * grad(vh, i) with lr a double; launched through the
// reference's own for_each_vertex kernel (kernels/for_each.cuh:28-40) exactly as RXMeshStatic::for_each_vertex(DEVICE) does
struct RefSmoothingStep
{
    RawAttr3 grad, pos;
    double   lr;
    __device__ void operator()(const VertexHandle& vh) const
    {
        for (int i = 0; i < 3; ++i)
            pos(vh, i) -= lr * grad(vh, i);
    }
};

// "consume" variant (the roofline kernels of bench.py): out(s) = sum_i in(iter[i]) over raw
// per-patch arrays, one fp32 per element (row = patch * cap + local id, the reference's slab addressing)
template <uint32_t blockThreads, Op op, typename InH, typename OutH>
__global__ static void ref_consume_kernel(const Context context, const float* in, uint32_t cap_in, float* out,
                                          uint32_t cap_out)
{
    auto sum_lambda = [&](const InH& id, const Iterator<OutH>& iter) {
        float a = 0.f;
        for (uint32_t i = 0; i < iter.size(); ++i) {
            const auto pl = iter[i].unpack();
            a += in[(size_t)pl.first * cap_in + pl.second];
        }
        const auto pl                            = id.unpack();
        out[(size_t)pl.first * cap_out + pl.second] = a;
    };
    auto                block = cooperative_groups::this_thread_block();
    Query<blockThreads> query(context);
    ShmemAllocator      shrd_alloc;
    query.template dispatch<op>(block, shrd_alloc, sum_lambda);
}

static void write_file(const std::string& path, const void* p, size_t bytes)
{
    FILE* f = fopen(path.c_str(), "wb");
    if (!f || fwrite(p, 1, bytes, f) != bytes) {
        fprintf(stderr, "cannot write %s\n", path.c_str());
        exit(3);
    }
    fclose(f);
}

struct OpRun
{
    const char* name;
    int         src, dst;
    uint32_t    width;
    double      ms;
    int         occupancy;
    int         regs;
    int         smem_static;
};

constexpr uint32_t BT = 256;
// Dynamic shared memory handed to every reference kernel.  The reference computes a per-op bound
// (rxmesh_static.inl:499-841, ~11-22 KB at patch size 512).  The kernels' residency is bounded by their registers
// (printed with the measured blocks/SM in meta.json), not by this value, as long as it stays below 227 KB / 4;
// REF_SMEM=<bytes> overrides it for a sensitivity check.
static uint32_t g_smem = 27 * 1024;

template <Op op, typename InH, typename OutH>
static void run_op(RefMesh& rx, OpRun& r, const std::string& outdir, bool dump, int nrun)
{
    auto           kern = ref_query_kernel<BT, op, InH, OutH>;
    const uint32_t P    = rx.get_num_patches();
    const uint32_t cap  = rx.max_per_patch(r.src);
    uint64_t*      d_out;
    uint16_t*      d_size;
    const size_t   rows = (size_t)P * cap;
    CK(cudaMalloc(&d_out, rows * r.width * sizeof(uint64_t)));
    CK(cudaMalloc(&d_size, rows * sizeof(uint16_t)));
    CK(cudaMemset(d_out, 0xFF, rows * r.width * sizeof(uint64_t)));
    CK(cudaMemset(d_size, 0, rows * sizeof(uint16_t)));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r.occupancy, kern, BT, g_smem));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, kern));
    r.regs = fa.numRegs, r.smem_static = (int)fa.sharedSizeBytes;
    kern<<<P, BT, g_smem>>>(rx.get_context(), d_out, d_size, r.width, cap);
    CK(cudaDeviceSynchronize());
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    CK(cudaEventRecord(a));
    for (int i = 0; i < nrun; ++i)
        kern<<<P, BT, g_smem>>>(rx.get_context(), d_out, d_size, r.width, cap);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    r.ms = ms / nrun;
    if (dump) {
        std::vector<uint64_t> h(rows * r.width);
        std::vector<uint16_t> hs(rows);
        CK(cudaMemcpy(h.data(), d_out, h.size() * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hs.data(), d_size, hs.size() * sizeof(uint16_t), cudaMemcpyDeviceToHost));
        const uint32_t        n_src = r.src == 0 ? rx.get_num_vertices() : (r.src == 1 ? rx.get_num_edges() : rx.get_num_faces());
        std::vector<uint32_t> g((size_t)n_src * r.width, 0xFFFFFFFFu);
        std::vector<uint8_t>  seen(n_src, 0);
        for (uint32_t p = 0; p < P; ++p) {
            const auto& ls = rx.ltog(r.src)[p];
            for (uint32_t l = 0; l < ls.size(); ++l) {
                if (!detail::is_owned((uint16_t)l, rx.owned_mask(p, r.src))) continue;
                const uint32_t gs = ls[l];
                if (seen[gs]++) {
                    fprintf(stderr, "%s: source %u owned twice\n", r.name, gs);
                    exit(4);
                }
                const size_t row = (size_t)p * cap + l;
                if (hs[row] > r.width) {
                    fprintf(stderr, "%s: row wider (%u) than %u\n", r.name, hs[row], r.width);
                    exit(4);
                }
                for (uint32_t i = 0; i < hs[row]; ++i) {
                    const uint64_t hd = h[row * r.width + i];
                    if (hd == INVALID64) continue;
                    const auto pl = detail::unpack(hd);
                    g[(size_t)gs * r.width + i] = rx.ltog(r.dst)[pl.first][pl.second];
                }
            }
        }
        for (uint32_t s = 0; s < n_src; ++s)
            if (!seen[s]) {
                fprintf(stderr, "%s: source %u never visited\n", r.name, s);
                exit(4);
            }
        write_file(outdir + "/q_" + r.name + ".u32", g.data(), g.size() * sizeof(uint32_t));
    }
    CK(cudaFree(d_out));
    CK(cudaFree(d_size));
}

int main(int argc, char** argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: %s mesh.bin outdir [patch_size] [dump] [nrun]\n", argv[0]);
        return 1;
    }
    const std::string outdir     = argv[2];
    const uint32_t    patch_size = argc > 3 ? (uint32_t)atoi(argv[3]) : 512;
    const bool        dump       = argc > 4 ? atoi(argv[4]) != 0 : true;
    const int         nrun       = argc > 5 ? atoi(argv[5]) : 100;
    if (getenv("REF_SMEM")) g_smem = (uint32_t)atoi(getenv("REF_SMEM"));
    FILE*             f          = fopen(argv[1], "rb");
    if (!f) {
        fprintf(stderr, "cannot open %s\n", argv[1]);
        return 1;
    }
    uint32_t hdr[2];
    if (fread(hdr, 4, 2, f) != 2) return 1;
    const uint32_t        nv = hdr[0], nf = hdr[1];
    std::vector<uint32_t> fvflat((size_t)nf * 3);
    std::vector<float>    x((size_t)nv * 3);
    if (fread(fvflat.data(), 4, fvflat.size(), f) != fvflat.size()) return 1;
    if (fread(x.data(), 4, x.size(), f) != x.size()) return 1;
    fclose(f);
    std::vector<std::vector<uint32_t>> fv(nf, std::vector<uint32_t>(3));
    for (uint32_t i = 0; i < nf; ++i)
        for (int k = 0; k < 3; ++k)
            fv[i][k] = fvflat[3 * (size_t)i + k];

    rx_init(0, spdlog::level::warn);
    CPUTimer build_timer;
    build_timer.start();
    RefMesh rx(fv, patch_size);
    build_timer.stop();
    const uint32_t P = rx.get_num_patches();

    OpRun runs[10] = {{"VV", 0, 0, rx.max_valence()}, {"VE", 0, 1, rx.max_valence()}, {"VF", 0, 2, rx.max_valence()},
                     {"EV", 1, 0, 2},          {"EF", 1, 2, rx.max_ef()}, {"FV", 2, 0, 3},
                     {"FE", 2, 1, 3},          {"FF", 2, 2, rx.max_ff()},
                     {"EVDiamond", 1, 0, 4},   {"EE", 1, 1, 4}};
    run_op<Op::VV, VertexHandle, VertexHandle>(rx, runs[0], outdir, dump, nrun);
    run_op<Op::VE, VertexHandle, EdgeHandle>(rx, runs[1], outdir, dump, nrun);
    run_op<Op::VF, VertexHandle, FaceHandle>(rx, runs[2], outdir, dump, nrun);
    run_op<Op::EV, EdgeHandle, VertexHandle>(rx, runs[3], outdir, dump, nrun);
    run_op<Op::EF, EdgeHandle, FaceHandle>(rx, runs[4], outdir, dump, nrun);
    run_op<Op::FV, FaceHandle, VertexHandle>(rx, runs[5], outdir, dump, nrun);
    run_op<Op::FE, FaceHandle, EdgeHandle>(rx, runs[6], outdir, dump, nrun);
    run_op<Op::FF, FaceHandle, FaceHandle>(rx, runs[7], outdir, dump, nrun);
    const bool edge4 = rx.max_ef() <= 2 && !getenv("REF_NO_EDGE4");  // EVDiamond / EE need an edge-manifold mesh
    if (edge4) {
        run_op<Op::EVDiamond, EdgeHandle, VertexHandle>(rx, runs[8], outdir, dump, nrun);
        run_op<Op::EE, EdgeHandle, EdgeHandle>(rx, runs[9], outdir, dump, nrun);
    }

    // vertex normals: raw AoSoA coords scattered from the global array through ltog (every local copy filled,
    // like the reference's attribute upload), normals zeroed outside the timed region (vertex_normal.cu:70-72)
    double vn_ms = 0;
    int    vn_occ = 0;
    {
        const uint32_t     cap = rx.max_per_patch(0);
        std::vector<float> hc((size_t)P * 3 * cap, 0.f);
        for (uint32_t p = 0; p < P; ++p) {
            const auto& ls = rx.ltog(0)[p];
            for (uint32_t l = 0; l < ls.size(); ++l)
                for (int a = 0; a < 3; ++a)
                    hc[((size_t)p * 3 + a) * cap + l] = x[3 * (size_t)ls[l] + a];
        }
        RawAttr3 coords{nullptr, cap}, normals{nullptr, cap};
        CK(cudaMalloc(&coords.data, hc.size() * 4));
        CK(cudaMalloc(&normals.data, hc.size() * 4));
        CK(cudaMemcpy(coords.data, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice));
        auto kern = ref_vertex_normal_kernel<BT>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_smem));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&vn_occ, kern, BT, g_smem));
        cudaEvent_t a, b;
        CK(cudaEventCreate(&a));
        CK(cudaEventCreate(&b));
        float total = 0;
        for (int i = 0; i <= nrun; ++i) {
            CK(cudaMemset(normals.data, 0, hc.size() * 4));
            CK(cudaEventRecord(a));
            kern<<<P, BT, g_smem>>>(rx.get_context(), coords, normals);
            CK(cudaEventRecord(b));
            CK(cudaEventSynchronize(b));
            float ms;
            CK(cudaEventElapsedTime(&ms, a, b));
            if (i > 0) total += ms;
        }
        vn_ms = total / nrun;
        if (dump) {
            std::vector<float> hn(hc.size());
            CK(cudaMemcpy(hn.data(), normals.data, hn.size() * 4, cudaMemcpyDeviceToHost));
            std::vector<float> g((size_t)nv * 3, 0.f);
            for (uint32_t p = 0; p < P; ++p) {
                const auto& ls = rx.ltog(0)[p];
                for (uint32_t l = 0; l < ls.size(); ++l)
                    if (detail::is_owned((uint16_t)l, rx.owned_mask(p, 0)))
                        for (int a = 0; a < 3; ++a)
                            g[3 * (size_t)ls[l] + a] = hn[((size_t)p * 3 + a) * cap + l];
            }
            write_file(outdir + "/vn.f32", g.data(), g.size() * 4);
        }
        CK(cudaFree(coords.data));
        CK(cudaFree(normals.data));
    }

    // manual smoothing: 5 iterations, positions after iteration 1 and 5 (lap_1.f32, lap_5.f32)
    double lap_ms = 0;
    {
        const uint32_t     cap = rx.max_per_patch(0);
        std::vector<float> hc((size_t)P * 3 * cap, 0.f);
        for (uint32_t p = 0; p < P; ++p) {
            const auto& ls = rx.ltog(0)[p];
            for (uint32_t l = 0; l < ls.size(); ++l)
                for (int a = 0; a < 3; ++a)
                    hc[((size_t)p * 3 + a) * cap + l] = x[3 * (size_t)ls[l] + a];
        }
        RawAttr3 pos{nullptr, cap}, grad{nullptr, cap};
        CK(cudaMalloc(&pos.data, hc.size() * 4));
        CK(cudaMalloc(&grad.data, hc.size() * 4));
        CK(cudaMemcpy(pos.data, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice));
        auto kern = ref_smoothing_grad_kernel<BT>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_smem));
        cudaEvent_t a, b;
        CK(cudaEventCreate(&a));
        CK(cudaEventCreate(&b));
        CK(cudaEventRecord(a));
        for (int it = 1; it <= 5; ++it) {
            CK(cudaMemset(grad.data, 0, hc.size() * 4));  // grad.reset(0, DEVICE), manual.h:38
            kern<<<P, BT, g_smem>>>(rx.get_context(), pos, grad);
            detail::for_each_vertex<<<P, 256>>>(P, rx.device_patches(), RefSmoothingStep{grad, pos, 0.01});
            if (dump && (it == 1 || it == 5)) {
                CK(cudaDeviceSynchronize());
                std::vector<float> hp(hc.size());
                CK(cudaMemcpy(hp.data(), pos.data, hp.size() * 4, cudaMemcpyDeviceToHost));
                std::vector<float> g((size_t)nv * 3, 0.f);
                for (uint32_t p = 0; p < P; ++p) {
                    const auto& ls = rx.ltog(0)[p];
                    for (uint32_t l = 0; l < ls.size(); ++l)
                        if (detail::is_owned((uint16_t)l, rx.owned_mask(p, 0)))
                            for (int c = 0; c < 3; ++c)
                                g[3 * (size_t)ls[l] + c] = hp[((size_t)p * 3 + c) * cap + l];
                }
                write_file(outdir + (it == 1 ? "/lap_1.f32" : "/lap_5.f32"), g.data(), g.size() * 4);
            }
        }
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        lap_ms = ms / 5;
        CK(cudaFree(pos.data));
        CK(cudaFree(grad.data));
    }

    // consume variants of VV and VF (inputs = 1.0 everywhere, so out = neighbour count: checked on the host)
    double cons_ms[2] = {0, 0};
    {
        const uint32_t capv = rx.max_per_patch(0), capf = rx.max_per_patch(2);
        float *        d_inv, *d_inf, *d_out;
        CK(cudaMalloc(&d_inv, (size_t)P * capv * 4));
        CK(cudaMalloc(&d_inf, (size_t)P * capf * 4));
        CK(cudaMalloc(&d_out, (size_t)P * capv * 4));
        std::vector<float> ones((size_t)P * std::max(capv, capf), 1.0f);
        CK(cudaMemcpy(d_inv, ones.data(), (size_t)P * capv * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_inf, ones.data(), (size_t)P * capf * 4, cudaMemcpyHostToDevice));
        auto kvv = ref_consume_kernel<BT, Op::VV, VertexHandle, VertexHandle>;
        auto kvf = ref_consume_kernel<BT, Op::VF, VertexHandle, FaceHandle>;
        CK(cudaFuncSetAttribute(kvv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_smem));
        CK(cudaFuncSetAttribute(kvf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_smem));
        cudaEvent_t a, b;
        CK(cudaEventCreate(&a));
        CK(cudaEventCreate(&b));
        for (int which = 0; which < 2; ++which) {
            auto launch = [&]() {
                if (which == 0)
                    kvv<<<P, BT, g_smem>>>(rx.get_context(), d_inv, capv, d_out, capv);
                else
                    kvf<<<P, BT, g_smem>>>(rx.get_context(), d_inf, capf, d_out, capv);
            };
            CK(cudaMemset(d_out, 0, (size_t)P * capv * 4));
            launch();
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(a));
            for (int i = 0; i < nrun; ++i)
                launch();
            CK(cudaEventRecord(b));
            CK(cudaEventSynchronize(b));
            float ms;
            CK(cudaEventElapsedTime(&ms, a, b));
            cons_ms[which] = ms / nrun;
            std::vector<float> ho((size_t)P * capv);
            CK(cudaMemcpy(ho.data(), d_out, ho.size() * 4, cudaMemcpyDeviceToHost));
            double total = 0;
            for (uint32_t p = 0; p < P; ++p)
                for (uint32_t l = 0; l < rx.ltog(0)[p].size(); ++l)
                    if (detail::is_owned((uint16_t)l, rx.owned_mask(p, 0))) total += ho[(size_t)p * capv + l];
            const double want = which == 0 ? 2.0 * rx.get_num_edges() : 3.0 * rx.get_num_faces();
            if (total != want) {
                fprintf(stderr, "consume %d: sum %.0f != %.0f\n", which, total, want);
                exit(4);
            }
        }
        CK(cudaFree(d_inv));
        CK(cudaFree(d_inf));
        CK(cudaFree(d_out));
    }

    if (dump) {
        write_file(outdir + "/face_patch.u32", rx.face_patch().data(), (size_t)nf * 4);
        const char* tn[3] = {"v", "e", "f"};
        for (int t = 0; t < 3; ++t) {
            std::vector<uint32_t> off(P + 1, 0), val, own(P);
            for (uint32_t p = 0; p < P; ++p) {
                const auto& ls = rx.ltog(t)[p];
                off[p + 1]     = off[p] + (uint32_t)ls.size();
                val.insert(val.end(), ls.begin(), ls.end());
                own[p] = rx.num_owned(t)[p];
            }
            write_file(outdir + "/ltog_off_" + tn[t] + ".u32", off.data(), off.size() * 4);
            write_file(outdir + "/ltog_" + tn[t] + ".u32", val.data(), val.size() * 4);
            write_file(outdir + "/owned_" + tn[t] + ".u32", own.data(), own.size() * 4);
        }
    }

    std::string js = "{";
    char        buf[512];
    snprintf(buf, sizeof buf,
             "\"nv\": %u, \"ne\": %u, \"nf\": %u, \"patches\": %u, \"patch_size\": %u, \"build_ms\": %.1f, "
             "\"max_v\": %u, \"max_e\": %u, \"max_f\": %u, \"smem_dyn\": %u, \"block\": %u, \"nrun\": %d, \"ops\": {",
             rx.get_num_vertices(), rx.get_num_edges(), rx.get_num_faces(), P, patch_size, build_timer.elapsed_millis(),
             rx.max_per_patch(0), rx.max_per_patch(1), rx.max_per_patch(2), g_smem, BT, nrun);
    js += buf;
    for (int i = 0; i < (edge4 ? 10 : 8); ++i) {
        snprintf(buf, sizeof buf,
                 "%s\"%s\": {\"ms\": %.6f, \"width\": %u, \"blocks_per_sm\": %d, \"regs\": %d, \"smem_static\": %d}",
                 i ? ", " : "", runs[i].name, runs[i].ms, runs[i].width, runs[i].occupancy, runs[i].regs, runs[i].smem_static);
        js += buf;
    }
    snprintf(buf, sizeof buf,
             "}, \"vertex_normals\": {\"ms\": %.6f, \"blocks_per_sm\": %d}, \"consume\": {\"VV\": %.6f, \"VF\": %.6f}, "
             "\"smoothing_ms_per_iteration\": %.6f}", vn_ms, vn_occ, cons_ms[0], cons_ms[1], lap_ms);
    js += buf;
    write_file(outdir + "/meta.json", js.data(), js.size());
    printf("%s\n", js.c_str());
    return 0;
}
