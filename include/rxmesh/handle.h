// rxmesh/handle.h -- Vertex/Edge/FaceHandle: 64-bit (patch_id << 32 | local id), the reference's encoding
// (include/rxmesh/handle.h:19-41) so handles stored in attributes are interchangeable.
#pragma once
#include <utility>
#include "rxmesh/types.h"

namespace rxmesh {
struct LocalVertexT { uint16_t id; __host__ __device__ LocalVertexT(uint16_t i = INVALID16) : id(i) {} };
struct LocalEdgeT { uint16_t id; __host__ __device__ LocalEdgeT(uint16_t i = INVALID16) : id(i) {} };
struct LocalFaceT { uint16_t id; __host__ __device__ LocalFaceT(uint16_t i = INVALID16) : id(i) {} };

namespace detail {
template <typename LocalTT, uint32_t ELEM>
struct HandleBase
{
    using LocalT = LocalTT;
    static constexpr uint32_t elem = ELEM;
    constexpr __host__ __device__ HandleBase() : m_handle(INVALID64) {}
    explicit constexpr __host__ __device__ HandleBase(uint64_t h) : m_handle(h) {}
    __host__ __device__ HandleBase(uint32_t patch_id, LocalTT local) : m_handle(((uint64_t)patch_id << 32) | local.id) {}
    constexpr __host__ __device__ bool     is_valid() const { return m_handle != INVALID64; }
    constexpr __host__ __device__ uint64_t unique_id() const { return m_handle; }
    constexpr __host__ __device__ uint32_t patch_id() const { return (uint32_t)(m_handle >> 32); }
    constexpr __host__ __device__ uint16_t local_id() const { return (uint16_t)(m_handle & 0xFFFFu); }
    __host__ __device__ std::pair<uint32_t, uint16_t> unpack() const { return std::make_pair(patch_id(), local_id()); }
    uint64_t m_handle;
};
}  // namespace detail

struct VertexHandle : detail::HandleBase<LocalVertexT, rxm::ELEM_V>
{
    using Handle = VertexHandle;
    using detail::HandleBase<LocalVertexT, rxm::ELEM_V>::HandleBase;
    constexpr __host__ __device__ bool operator==(const VertexHandle& r) const { return m_handle == r.m_handle; }
    constexpr __host__ __device__ bool operator!=(const VertexHandle& r) const { return m_handle != r.m_handle; }
};
struct EdgeHandle : detail::HandleBase<LocalEdgeT, rxm::ELEM_E>
{
    using Handle = EdgeHandle;
    using detail::HandleBase<LocalEdgeT, rxm::ELEM_E>::HandleBase;
    constexpr __host__ __device__ bool operator==(const EdgeHandle& r) const { return m_handle == r.m_handle; }
    constexpr __host__ __device__ bool operator!=(const EdgeHandle& r) const { return m_handle != r.m_handle; }
};
struct FaceHandle : detail::HandleBase<LocalFaceT, rxm::ELEM_F>
{
    using Handle = FaceHandle;
    using detail::HandleBase<LocalFaceT, rxm::ELEM_F>::HandleBase;
    constexpr __host__ __device__ bool operator==(const FaceHandle& r) const { return m_handle == r.m_handle; }
    constexpr __host__ __device__ bool operator!=(const FaceHandle& r) const { return m_handle != r.m_handle; }
};

// InputHandle<op> / OutputHandle<op> traits (handle.h:506-669)
template <Op op> struct InputHandle;
template <Op op> struct OutputHandle;
#define RXM_OP_HANDLES(OPV, IN, OUT)                       \
    template <> struct InputHandle<OPV> { using type = IN; }; \
    template <> struct OutputHandle<OPV> { using type = OUT; };
RXM_OP_HANDLES(Op::VV, VertexHandle, VertexHandle)
RXM_OP_HANDLES(Op::VE, VertexHandle, EdgeHandle)
RXM_OP_HANDLES(Op::VF, VertexHandle, FaceHandle)
RXM_OP_HANDLES(Op::EV, EdgeHandle, VertexHandle)
RXM_OP_HANDLES(Op::EF, EdgeHandle, FaceHandle)
RXM_OP_HANDLES(Op::FV, FaceHandle, VertexHandle)
RXM_OP_HANDLES(Op::FE, FaceHandle, EdgeHandle)
RXM_OP_HANDLES(Op::FF, FaceHandle, FaceHandle)
RXM_OP_HANDLES(Op::EE, EdgeHandle, EdgeHandle)
RXM_OP_HANDLES(Op::EVDiamond, EdgeHandle, VertexHandle)
#undef RXM_OP_HANDLES
}  // namespace rxmesh
