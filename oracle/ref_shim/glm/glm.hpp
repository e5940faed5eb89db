// stand-in for glm 1.0.1 (see ../README.md): only the aggregate types rxmesh/types.h names
#pragma once
namespace glm {
enum qualifier { defaultp };
template <int N, typename T, qualifier Q = defaultp>
struct vec {
    T d[N];
    __host__ __device__ T&       operator[](int i) { return d[i]; }
    __host__ __device__ const T& operator[](int i) const { return d[i]; }
};
template <int N, int M, typename T, qualifier Q = defaultp>
struct mat { T d[N * M]; };
}  // namespace glm
