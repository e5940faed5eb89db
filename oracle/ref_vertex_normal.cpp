// oracle/_ref wrapper -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
// Compiles the reference's own CPU vertex-normal implementation
// (/root/reference/apps/VertexNormal/vertex_normal_ref.h:5-85) from where it
// lies, unmodified, and exposes it through a flat-array C entry point so that
// tests/golden/make_golden.py and bench.py --impl reference can call it.
#include <cstdint>
#include <cstring>
#include <vector>

#include "vertex_normal_ref.h"  // found via -I/root/reference/apps/VertexNormal

extern "C" void ref_vertex_normal_f32(const uint32_t* fv, uint32_t nf, const float* x, uint32_t nv,
                                      float* out)
{
    std::vector<std::vector<uint32_t>> Faces(nf, std::vector<uint32_t>(3));
    std::vector<std::vector<float>>    Verts(nv, std::vector<float>(3));
    for (uint32_t f = 0; f < nf; ++f)
        for (int i = 0; i < 3; ++i)
            Faces[f][i] = fv[3 * (size_t)f + i];
    for (uint32_t v = 0; v < nv; ++v)
        for (int i = 0; i < 3; ++i)
            Verts[v][i] = x[3 * (size_t)v + i];
    std::vector<float> n(3 * (size_t)nv);
    vertex_normal_ref<float>(Faces, Verts, n);
    std::memcpy(out, n.data(), n.size() * sizeof(float));
}

// Timed variant: the vector<vector<>> conversion is setup (the reference app
// does it in import_obj, outside its timer); only the reference loop is timed.
#include <chrono>
extern "C" double ref_vertex_normal_f32_timed(const uint32_t* fv, uint32_t nf, const float* x,
                                              uint32_t nv, float* out, int repeats)
{
    std::vector<std::vector<uint32_t>> Faces(nf, std::vector<uint32_t>(3));
    std::vector<std::vector<float>>    Verts(nv, std::vector<float>(3));
    for (uint32_t f = 0; f < nf; ++f)
        for (int i = 0; i < 3; ++i)
            Faces[f][i] = fv[3 * (size_t)f + i];
    for (uint32_t v = 0; v < nv; ++v)
        for (int i = 0; i < 3; ++i)
            Verts[v][i] = x[3 * (size_t)v + i];
    std::vector<float> n(3 * (size_t)nv);
    auto               t0 = std::chrono::high_resolution_clock::now();
    for (int r = 0; r < repeats; ++r)
        vertex_normal_ref<float>(Faces, Verts, n);
    auto t1 = std::chrono::high_resolution_clock::now();
    std::memcpy(out, n.data(), n.size() * sizeof(float));
    return std::chrono::duration<double>(t1 - t0).count() / repeats;
}
