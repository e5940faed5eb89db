// rxmesh/geometry_util.cuh -- the small geometry helpers user kernels of the reference call
// (include/rxmesh/geometry_util.cuh:40-206): tri_normal, tri_area, clamp_cot, partial_voronoi_area,
// edge_cotan_weight.  Same names, argument order and results; written against the shim's glm-like vec.
#pragma once
#include <limits>
#include "rxmesh/types.h"

namespace rxmesh {
template <typename T>
__host__ __device__ inline vec3<T> tri_normal(const vec3<T>& p0, const vec3<T>& p1, const vec3<T>& p2)
{
    return glm::normalize(glm::cross(p1 - p0, p2 - p0));
}
template <typename T>
__host__ __device__ inline T tri_area(const vec3<T>& p0, const vec3<T>& p1, const vec3<T>& p2)
{
    return T(0.5) * glm::length(glm::cross(p1 - p0, p2 - p0));
}
// clamp a cotangent as if the angle were in [3, 177] degrees (geometry_util.cuh:105-114)
template <typename T>
__host__ __device__ inline void clamp_cot(T& v)
{
    const T bound = T(19.1);
    v             = v < -bound ? -bound : (v > bound ? bound : v);
}
// partial (mixed) Voronoi area of centre p in triangle p->q->r (geometry_util.cuh:120-172)
template <typename T>
__host__ __device__ inline T partial_voronoi_area(const vec3<T>& p, const vec3<T>& q, const vec3<T>& r)
{
    const vec3<T> pq = q - p, qr = r - q, pr = r - p;
    const T       area = tri_area(p, q, r);
    if (area <= std::numeric_limits<T>::min()) return T(-1);
    const T dotp = glm::dot(pq, pr), dotq = -glm::dot(qr, pq), dotr = glm::dot(qr, pr);
    if (dotp < T(0)) return T(0.25) * area;
    if (dotq < T(0) || dotr < T(0)) return T(0.125) * area;
    T cotq = dotq / area, cotr = dotr / area;
    clamp_cot(cotq);
    clamp_cot(cotr);
    return T(0.125) * (glm::length2(pr) * cotq + glm::length2(pq) * cotr);
}
// cotangent weight of edge p-r whose diamond is closed by q and s (geometry_util.cuh:178-206)
template <typename T>
__host__ __device__ inline T edge_cotan_weight(const vec3<T>& p, const vec3<T>& r, const vec3<T>& q, const vec3<T>& s)
{
    auto partial = [&](const vec3<T>& v) -> T {
        const T area = tri_area(p, r, v);
        if (area > std::numeric_limits<T>::min()) {
            T c = glm::dot(p - v, r - v) / area;
            clamp_cot(c);
            return c;
        }
        return T(0);
    };
    return partial(q) + partial(s);
}
}  // namespace rxmesh
