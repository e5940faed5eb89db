// rxmesh/types.h -- drop-in names of the reference's types.h (include/rxmesh/types.h:10-129) for the
// static hot path, implemented over librxmesh_b200.  Not a copy: only the public vocabulary is kept.
#pragma once
#include <stdint.h>
#include <cmath>
#include <string>

#include "rxmesh_b200/patch_layout.h"

#ifndef __CUDACC__
#error "the rxmesh_b200 C++ shim is device code: compile user sources with nvcc (-arch sm_100a --expt-extended-lambda)"
#endif

// minimal glm-compatible small vectors: the reference's device lambdas use glm::vec / cross / distance2 /
// normalize / length / dot (glm 1.0.1, apps/VertexNormal/vertex_normal_kernel.cuh:23-28)
namespace glm {
template <int N, typename T>
struct vec
{
    T v[N];
    __host__ __device__ vec() { for (int i = 0; i < N; ++i) v[i] = T(0); }
    __host__ __device__ explicit vec(T s) { for (int i = 0; i < N; ++i) v[i] = s; }
    __host__ __device__ vec(T a, T b) { static_assert(N == 2, ""); v[0] = a, v[1] = b; }
    __host__ __device__ vec(T a, T b, T c) { static_assert(N == 3, ""); v[0] = a, v[1] = b, v[2] = c; }
    __host__ __device__ T&       operator[](int i) { return v[i]; }
    __host__ __device__ const T& operator[](int i) const { return v[i]; }
    __host__ __device__ vec& operator+=(const vec& o) { for (int i = 0; i < N; ++i) v[i] += o.v[i]; return *this; }
    __host__ __device__ vec& operator-=(const vec& o) { for (int i = 0; i < N; ++i) v[i] -= o.v[i]; return *this; }
    __host__ __device__ vec& operator*=(T s) { for (int i = 0; i < N; ++i) v[i] *= s; return *this; }
    __host__ __device__ vec& operator/=(T s) { for (int i = 0; i < N; ++i) v[i] /= s; return *this; }
};
template <int N, typename T> __host__ __device__ inline vec<N, T> operator+(vec<N, T> a, const vec<N, T>& b) { return a += b; }
template <int N, typename T> __host__ __device__ inline vec<N, T> operator-(vec<N, T> a, const vec<N, T>& b) { return a -= b; }
template <int N, typename T> __host__ __device__ inline vec<N, T> operator*(vec<N, T> a, T s) { return a *= s; }
template <int N, typename T> __host__ __device__ inline vec<N, T> operator*(T s, vec<N, T> a) { return a *= s; }
template <int N, typename T> __host__ __device__ inline vec<N, T> operator/(vec<N, T> a, T s) { return a /= s; }
template <int N, typename T> __host__ __device__ inline T dot(const vec<N, T>& a, const vec<N, T>& b)
{
    T s = T(0);
    for (int i = 0; i < N; ++i) s += a[i] * b[i];
    return s;
}
template <typename T> __host__ __device__ inline vec<3, T> cross(const vec<3, T>& a, const vec<3, T>& b)
{
    return vec<3, T>(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
template <int N, typename T> __host__ __device__ inline T length2(const vec<N, T>& a) { return dot(a, a); }
template <int N, typename T> __host__ __device__ inline T length(const vec<N, T>& a) { return sqrt(dot(a, a)); }
template <int N, typename T> __host__ __device__ inline T distance2(const vec<N, T>& a, const vec<N, T>& b) { return length2(a - b); }
template <int N, typename T> __host__ __device__ inline T distance(const vec<N, T>& a, const vec<N, T>& b) { return length(a - b); }
template <int N, typename T> __host__ __device__ inline vec<N, T> normalize(const vec<N, T>& a) { return a / length(a); }
template <typename T> __host__ __device__ constexpr T pi() { return T(3.14159265358979323846264338327950288); }
template <typename T> __host__ __device__ constexpr T half_pi() { return T(1.57079632679489661923132169163975144); }
template <typename T> __host__ __device__ constexpr T two_pi() { return T(6.28318530717958647692528676655900576); }
template <typename T> __host__ __device__ inline T acos(T x) { return ::acos(x); }
using vec2  = vec<2, float>;
using vec3  = vec<3, float>;
using fvec3 = vec<3, float>;
using uvec3 = vec<3, uint32_t>;
}  // namespace glm

namespace rxmesh {
using rx_coord_t = float;  // types.h:10-14 (RX_USE_DOUBLE off)
template <typename T> using vec2 = glm::vec<2, T>;
template <typename T> using vec3 = glm::vec<3, T>;
using glm::cross; using glm::dot; using glm::normalize; using glm::length; using glm::length2;
using glm::distance; using glm::distance2;
template <int N, typename T> __host__ __device__ inline T dist2(const glm::vec<N, T>& a, const glm::vec<N, T>& b) { return glm::distance2(a, b); }

using locationT = uint32_t;
enum : locationT { LOCATION_NONE = 0x00, HOST = 0x01, DEVICE = 0x02, LOCATION_ALL = 0x0F };
using layoutT = uint32_t;
enum : layoutT { AoS = 0x00, AoSoA = 0x01, SoA = 0x02 };
inline std::string layout_to_string(const layoutT layout)  // types.h:94-108
{
    return layout == AoS ? "AoS" : (layout == AoSoA ? "AoSoA" : (layout == SoA ? "SoA" : ""));
}
inline std::string location_to_string(const locationT location)  // types.h:63-79
{
    return location == LOCATION_NONE ? "NONE" : (location == HOST ? "HOST" : (location == DEVICE ? "DEVICE" : (location == LOCATION_ALL ? "ALL" : "")));
}

enum class Op { INVALID = -1, V = 0, E = 1, F = 2, VV = 3, VE = 4, VF = 5, FV = 6, FE = 7, FF = 8, EV = 9, EE = 10, EF = 11, EVDiamond = 12 };

inline std::string op_to_string(Op op)
{
    switch (op) {
        case Op::VV: return "VV"; case Op::VE: return "VE"; case Op::VF: return "VF"; case Op::FV: return "FV";
        case Op::FE: return "FE"; case Op::FF: return "FF"; case Op::EV: return "EV"; case Op::EF: return "EF";
        case Op::V: return "V"; case Op::E: return "E"; case Op::F: return "F"; default: return "INVALID";
    }
}

#ifndef INVALID64
#define INVALID64 0xFFFFFFFFFFFFFFFFu
#define INVALID32 0xFFFFFFFFu
#define INVALID16 0xFFFFu
#endif
#ifndef DIVIDE_UP
#define DIVIDE_UP(a, b) (((a) + (b) - 1) / (b))
#endif
}  // namespace rxmesh
