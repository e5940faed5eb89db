// oracle/_ref/ref_hardwired -- TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/README.md).
// The reference's flat-array GPU vertex-normal kernel, apps/VertexNormal/vertex_normal_hardwired.cuh:6-78 (one thread per
// face over a plain u32 face list, nine global float atomicAdds per face), included UNMODIFIED from /root/reference and
// launched exactly as its host wrapper does (:126-147: 256 threads, cudaMemset of the normals outside the timer, one
// event pair per launch).  SURVEY.md 2.3 names it "the baseline to beat on the same box".  Input: the bench's own grid
// generator (rxmesh_b200/meshio.py: grid()) restated here so the binary needs no file: vertex id = i*n + j,
// x = j, z = i, y = 0.05 sin(0.01 x) cos(0.013 z); per quad (a, b, c), (c, b, d) with a = idx, b = idx + n, c = idx + 1.
//   ref_hardwired <n> <num_run>   -> one JSON line
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "vertex_normal_hardwired.cuh"  // -I $(REF)/apps/VertexNormal, stand-ins for its two util includes in ref_shim/hardwired

#define CK(x)                                                                             \
    do {                                                                                  \
        cudaError_t e_ = (x);                                                             \
        if (e_ != cudaSuccess) {                                                          \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                      \
            return 1;                                                                     \
        }                                                                                 \
    } while (0)

int main(int argc, char** argv)
{
    const uint32_t n = argc > 1 ? (uint32_t)atoi(argv[1]) : 1000u, num_run = argc > 2 ? (uint32_t)atoi(argv[2]) : 10u;
    const uint64_t nv = (uint64_t)n * n, nf = 2ull * (n - 1) * (n - 1);
    if (nv > 0x7FFFFFFFull / 3 * 3 || nf * 3 > 0xFFFFFFFFull) {
        fprintf(stderr, "mesh too large for the kernel's 32-bit indexing\n");
        return 1;
    }
    std::vector<float>    hv(3 * nv);
    std::vector<uint32_t> hf(3 * nf);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i)
        for (uint32_t j = 0; j < n; ++j) {
            const uint64_t v = (uint64_t)i * n + j;
            const float    x = (float)j, z = (float)i;
            hv[3 * v]     = x;
            hv[3 * v + 1] = (float)(0.05 * std::sin(0.01 * (double)x) * std::cos(0.013 * (double)z));
            hv[3 * v + 2] = z;
        }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n - 1; ++i)
        for (uint32_t j = 0; j + 1 < n; ++j) {
            const uint64_t q = (uint64_t)i * (n - 1) + j;
            const uint32_t a = (uint32_t)(i * n + j), b = a + n, c = a + 1, d = a + n + 1;
            uint32_t*      t = hf.data() + 6 * q;
            t[0] = a, t[1] = b, t[2] = c, t[3] = c, t[4] = b, t[5] = d;
        }
    uint32_t* d_face = nullptr;
    float *   d_verts = nullptr, *d_normals = nullptr;
    CK(cudaMalloc(&d_face, hf.size() * 4));
    CK(cudaMalloc(&d_verts, hv.size() * 4));
    CK(cudaMalloc(&d_normals, hv.size() * 4));
    CK(cudaMemcpy(d_face, hf.data(), hf.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_verts, hv.data(), hv.size() * 4, cudaMemcpyHostToDevice));
    const uint32_t threads = 256, blocks = (uint32_t)((nf + threads - 1) / threads);
    cudaEvent_t    e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    double total = 0, best = 1e30;
    for (uint32_t it = 0; it < num_run + 1; ++it) {  // first launch = warm-up
        CK(cudaMemset(d_normals, 0, hv.size() * 4));
        CK(cudaEventRecord(e0));
        vertex_normal_hardwired_kernel<float><<<blocks, threads>>>((uint32_t)nf, (uint32_t)nv, d_face, d_verts, d_normals);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        CK(cudaGetLastError());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it) total += ms, best = ms < best ? ms : best;
    }
    std::vector<float> hn(12);  // a few normals as a sanity value (interior vertex n+1: y component dominates)
    CK(cudaMemcpy(hn.data(), d_normals + 3ull * (n + 1), 12 * 4, cudaMemcpyDeviceToHost));
    const double ms = total / num_run;
    printf("{\"kernel\": \"vertex_normal_hardwired_kernel<float> (reference, unmodified)\", \"grid_side\": %u, \"faces\": %llu, "
           "\"vertices\": %llu, \"num_run\": %u, \"hardwired_ms\": %.6f, \"hardwired_ms_best\": %.6f, \"faces_per_s\": %.6e, "
           "\"bytes_moved_model\": \"12 F face list + 9 gathers + 9 global atomics per face\", "
           "\"alg_gbs\": %.2f, \"sample_normal\": [%.6f, %.6f, %.6f]}\n",
           n, (unsigned long long)nf, (unsigned long long)nv, num_run, ms, best, nf / (ms * 1e-3), 24.0 * nf / (ms * 1e-3) / 1e9,
           hn[0], hn[1], hn[2]);
    return 0;
}
