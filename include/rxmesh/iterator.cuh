// rxmesh/iterator.cuh -- Iterator<HandleT> (include/rxmesh/iterator.cuh:50-194): a view over one source
// element's query result in shared memory; operator[] resolves a local id to its OWNER handle through the
// patch's direct owner table (no hash probe).
#pragma once
#include "rxmesh/handle.h"
#include "rxmesh_b200/rxm_query.cuh"
namespace rxmesh {
template <typename HandleT>
struct Iterator
{
    using LocalT = typename HandleT::LocalT;
    __device__ Iterator(const rxm::dev::QueryResult& r, const rxm::dev::OwnerTable& ot, uint32_t src)
        : m_r(r), m_ot(ot), m_begin(r.begin(src)), m_end(r.end(src)) {}
    __device__ uint16_t size() const { return (uint16_t)(m_end - m_begin); }
    __device__ HandleT  operator[](const uint16_t i) const
    {
        if (m_begin + i >= m_end) return HandleT();
        return HandleT(m_ot.handle(m_r.at(m_begin + i)));
    }
    __device__ uint16_t local(const uint16_t i) const { return m_begin + i >= m_end ? (uint16_t)INVALID16 : (uint16_t)m_r.at(m_begin + i); }
    __device__ HandleT  front() const { return (*this)[0]; }
    __device__ HandleT  back() const { return (*this)[size() - 1]; }

   private:
    const rxm::dev::QueryResult& m_r;
    const rxm::dev::OwnerTable&  m_ot;
    uint32_t                     m_begin, m_end;
};
using VertexIterator = Iterator<VertexHandle>;
using EdgeIterator   = Iterator<EdgeHandle>;
using FaceIterator   = Iterator<FaceHandle>;
}  // namespace rxmesh
