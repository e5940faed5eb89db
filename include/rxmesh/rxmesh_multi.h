// rxmesh_multi.h -- RXMeshStatic over several GPUs from ONE host process (no Python, no torch.distributed).
//
// New relative to the reference: RXMeshStatic's constructors take no device (rxmesh_static.h:61-100, one cudaSetDevice in
// rx_init); SURVEY.md 8(e) asks for "device-list / shard options".  The class below is that knob: the constructor takes the
// device list next to the reference's (fv, patch_size) arguments, cuts the mesh into one shard of contiguous patches per
// device (rxm_multi_create) and keeps the mirrored ("ribbon across devices") rows current from inside the compute kernel.
// Plain host C++: compiles with g++, links against librxmesh_b200.so.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../rxmesh_b200.h"

namespace rxmesh {

class RXMeshMultiGPU
{
   public:
    RXMeshMultiGPU(const RXMeshMultiGPU&) = delete;
    // devices: CUDA device ids, one shard each (an id may repeat: several shards on one device)
    RXMeshMultiGPU(const std::vector<std::vector<uint32_t>>& fv, const std::vector<int>& devices, const uint32_t patch_size = 512,
                   const std::vector<uint32_t>& face_patch = {})
    {
        std::vector<uint32_t> flat;
        flat.reserve(3 * fv.size());
        for (const auto& f : fv) {
            if (f.size() != 3) die("non-triangular faces are not supported");  // rxmesh.cpp:590-597
            flat.insert(flat.end(), f.begin(), f.end());
        }
        create(flat.data(), (uint32_t)fv.size(), devices.data(), (int)devices.size(), patch_size, face_patch);
    }
    RXMeshMultiGPU(const uint32_t* fv, uint32_t num_faces, const std::vector<int>& devices, const uint32_t patch_size = 512,
                   const std::vector<uint32_t>& face_patch = {})
    {
        create(fv, num_faces, devices.data(), (int)devices.size(), patch_size, face_patch);
    }
    // host-only plan of num_shards shards (no CUDA call): sizes and halo counts can be inspected, compute calls fail
    RXMeshMultiGPU(const uint32_t* fv, uint32_t num_faces, int num_shards, const uint32_t patch_size = 512)
    {
        create(fv, num_faces, nullptr, num_shards, patch_size, {});
    }
    ~RXMeshMultiGPU() { rxm_multi_destroy(m_multi); }

    uint32_t get_num_shards() const { return (uint32_t)rxm_multi_info(m_multi, 0, -1); }
    uint32_t get_num_patches() const { return (uint32_t)rxm_multi_info(m_multi, 1, -1); }
    uint32_t get_num_vertices() const { return (uint32_t)rxm_multi_info(m_multi, 3, -1); }
    uint32_t get_num_faces() const { return (uint32_t)rxm_multi_info(m_multi, 4, -1); }
    // rows that cross a device boundary per exchange (all shards)
    uint64_t get_num_mirrored_vertices() const { return rxm_multi_info(m_multi, 2, -1); }
    uint32_t get_num_patches(int shard) const { return (uint32_t)rxm_multi_info(m_multi, 10, shard); }
    uint32_t get_num_owned_vertices(int shard) const { return (uint32_t)rxm_multi_info(m_multi, 12, shard); }
    uint32_t get_num_ghost_vertices(int shard) const { return (uint32_t)rxm_multi_info(m_multi, 13, shard); }
    // the shard's own mesh (real + ghost patches) for anything the whole-mesh calls below do not cover
    rxm_mesh* get_shard_mesh(int shard) { return rxm_multi_shard_mesh(m_multi, shard); }

    // apps/Smoothing/manual.h:86-104 iterated over all devices; coords / result in GLOBAL vertex order, [V][3]
    std::vector<float> laplacian_smooth(const std::vector<float>& coords, double lr, uint32_t iters)
    {
        std::vector<float> out(coords.size());
        if (coords.size() != 3ull * get_num_vertices()) die("laplacian_smooth: coords must hold 3 floats per vertex");
        if (rxm_multi_laplacian_smooth(m_multi, coords.data(), out.data(), lr, iters)) die(rxm_last_error());
        return out;
    }
    // apps/VertexNormal/vertex_normal_kernel.cuh:10-43 over all devices
    std::vector<float> vertex_normals(const std::vector<float>& coords)
    {
        std::vector<float> out(coords.size());
        if (coords.size() != 3ull * get_num_vertices()) die("vertex_normals: coords must hold 3 floats per vertex");
        if (rxm_multi_vertex_normals(m_multi, coords.data(), out.data())) die(rxm_last_error());
        return out;
    }

   private:
    rxm_multi* m_multi = nullptr;
    static void die(const char* msg)
    {
        fprintf(stderr, "rxmesh_b200: %s\n", msg);
        exit(EXIT_FAILURE);
    }
    void create(const uint32_t* fv, uint32_t nf, const int* devices, int n, uint32_t patch_size, const std::vector<uint32_t>& face_patch)
    {
        if (!face_patch.empty() && face_patch.size() != nf) die("face_patch must hold one patch id per face");
        if (rxm_multi_create(fv, nf, face_patch.empty() ? nullptr : face_patch.data(), patch_size, devices, n, 0, &m_multi))
            die(rxm_last_error());
    }
};

}  // namespace rxmesh
