// A user program of include/rxmesh/rxmesh_multi.h (plain g++): a grid mesh over N shards.
//   multi_user plan N        host-only plan, prints "<shards> <patches> <vertices> <mirrored>"
//   multi_user run d0,d1,..  smooth + normals on the listed devices; compares with a one-shard run; prints "ok <max diff>"
#include <cmath>
#include <cstring>
#include <string>

#include "rxmesh/rxmesh_multi.h"

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    const int             nx = 90, ny = 61;
    std::vector<uint32_t> fv;
    std::vector<float>    x;
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            x.push_back((float)i), x.push_back((float)j), x.push_back(0.3f * std::sin(0.4f * i) * std::cos(0.3f * j));
            if (i + 1 < nx && j + 1 < ny) {
                const uint32_t a = j * nx + i, b = a + 1, c = a + nx, d = c + 1;
                fv.insert(fv.end(), {a, b, d, a, d, c});
            }
        }
    const uint32_t nf = (uint32_t)fv.size() / 3;
    if (!strcmp(argv[1], "plan")) {
        rxmesh::RXMeshMultiGPU rx(fv.data(), nf, atoi(argv[2]), 128);
        uint64_t               owned = 0;
        for (uint32_t s = 0; s < rx.get_num_shards(); ++s)
            owned += rx.get_num_owned_vertices((int)s);
        if (owned != rx.get_num_vertices()) return 1;
        printf("%u %u %u %llu\n", rx.get_num_shards(), rx.get_num_patches(), rx.get_num_vertices(),
               (unsigned long long)rx.get_num_mirrored_vertices());
        return 0;
    }
    std::vector<int> devs;
    for (char* t = strtok(argv[2], ","); t; t = strtok(nullptr, ","))
        devs.push_back(atoi(t));
    rxmesh::RXMeshMultiGPU rx(fv.data(), nf, devs, 128), one(fv.data(), nf, std::vector<int>{devs[0]}, 128);
    const auto             a = rx.laplacian_smooth(x, 0.01, 30), b = one.laplacian_smooth(x, 0.01, 30);
    const auto             na = rx.vertex_normals(x), nb = one.vertex_normals(x);
    float                  d = 0.f, dn = 0.f;
    for (size_t i = 0; i < a.size(); ++i)
        d = std::fmax(d, std::fabs(a[i] - b[i])), dn = std::fmax(dn, std::fabs(na[i] - nb[i]));
    if (d != 0.f || dn != 0.f || a == x) {
        printf("mismatch %g %g\n", d, dn);
        return 1;
    }
    printf("ok %g\n", d);
    return 0;
}
