// rxmesh/attribute.h -- Attribute<T, HandleT> (include/rxmesh/attribute.h:56-731, attribute.cu:30-590) over the
// C ABI: storage for OWNED elements (AoS / AoSoA in slot order, include/rxmesh_b200/patch_layout.h; SoA = the
// reference's tensor layout, a gap-free column-major #elements x #attributes matrix over linear ids,
// attribute.h:249-261,406-421), host + device copies and the reference's operator()(handle, attr) on both sides.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../rxmesh_b200.h"
#include "rxmesh/handle.h"
#include "rxmesh/matrix/dense_matrix.h"

namespace rxmesh {
namespace detail {
inline void rxm_check(int rc)
{
    if (rc != RXM_OK) {  // CUDA_ERROR semantics of the reference: log and exit (util/macros.h:77-89)
        fprintf(stderr, "rxmesh_b200: %s\n", rxm_last_error());
        exit(EXIT_FAILURE);
    }
}
}  // namespace detail

class AttributeBase
{
   public:
    virtual const char* get_name() const = 0;
    virtual void        release(locationT location = LOCATION_ALL) = 0;
    virtual ~AttributeBase()                                       = default;
};

template <class T, typename HandleT>
class Attribute : public AttributeBase
{
   public:
    using Type       = T;
    using HandleType = HandleT;
    Attribute()      = default;

    // created by RXMeshStatic::add_*_attribute
    // Like the reference's Attribute (attribute.h:194) this object is a SHALLOW, trivially copyable view:
    // raw pointers only, so that copies captured by device lambdas / passed to kernels are plain memcpy.
    // The storage is owned by RXMeshStatic's container and freed by release().
    Attribute(rxm_mesh* mesh, const char* name, uint32_t num_attributes, locationT location, layoutT layout)
        : m_nattr(num_attributes), m_layout(layout), m_location(location)
    {
        static_assert(sizeof(T) == 1 || sizeof(T) == 2 || sizeof(T) == 4 || sizeof(T) == 8, "unsupported attribute type");
        m_name = strdup(name);
        rxm_attr* a = nullptr;
        detail::rxm_check(rxm_attr_create(mesh, HandleT::elem, sizeof(T), num_attributes, (int)location, (int)layout, &a));
        m_attr         = a;
        m_h            = (T*)rxm_attr_data(a, RXM_HOST);
        m_d            = (T*)rxm_attr_data(a, RXM_DEVICE);
        m_num_slots    = (uint32_t)rxm_mesh_info(mesh, RXM_INFO_NUM_SLOTS_V + HandleT::elem);
        m_num_elems    = (uint32_t)rxm_mesh_info(mesh, RXM_INFO_NUM_VERTICES + HandleT::elem);
        m_num_patches  = (uint32_t)rxm_mesh_info(mesh, RXM_INFO_NUM_PATCHES);
        m_storage      = rxm_attr_count(a);
        m_h_slot_base  = rxm_mesh_slot_base(mesh, HandleT::elem);
        m_d_slot_base  = rxm_mesh_device_slot_base(mesh, HandleT::elem);
        m_h_lin_base   = rxm_mesh_lin_base(mesh, HandleT::elem);
        m_d_lin_base   = rxm_mesh_device_lin_base(mesh, HandleT::elem);
    }
    Attribute(const Attribute&) = default;  // shallow, like the reference (attribute.h:194)

    const char* get_name() const override { return m_name; }
    __host__ __device__ uint32_t  get_num_attributes() const { return m_nattr; }
    __host__ __device__ layoutT   get_layout() const { return m_layout; }
    __host__ __device__ locationT get_allocated() const { return m_location; }
    __host__ __device__ bool      is_device_allocated() const { return (m_location & DEVICE) == DEVICE; }
    __host__ __device__ bool      is_host_allocated() const { return (m_location & HOST) == HOST; }
    __host__ __device__ T*        data(locationT location = DEVICE) const { return (location & DEVICE) ? m_d : m_h; }
    // number of T values in the raw allocation (attribute.h:249-252); SoA: rows() * cols()
    __host__ __device__ size_t storage_size() const { return (size_t)m_storage; }
    // true when the raw storage is one global column-major matrix (attribute.h:258-261)
    __host__ __device__ bool is_tensor_layout() const { return m_layout == SoA; }
    // rows = mesh elements of this type, cols = attributes per element (attribute.h:113-120)
    __host__ __device__ size_t   rows() const { return m_num_elems; }
    __host__ __device__ size_t   cols() const { return m_nattr; }
    __host__ __device__ uint32_t size() const { return m_num_elems; }
    __host__ __device__ uint32_t get_num_patches() const { return m_num_patches; }
    // owned elements of patch p (the reference's size(p), attribute.h:692) and the linear-id prefix behind it
    __host__ __device__ uint32_t size(const uint32_t p) const { return lin_base()[p + 1] - lin_base()[p]; }
    __host__ __device__ const uint32_t* lin_base(locationT location) const { return (location & DEVICE) ? m_d_lin_base : m_h_lin_base; }
    rxm_attr* c_handle() const { return m_attr; }

    void reset(const T value, locationT location, cudaStream_t stream = NULL) { detail::rxm_check(rxm_attr_reset(m_attr, &value, (int)location, stream)); }
    void move(locationT source, locationT target, cudaStream_t stream = NULL)
    {
        detail::rxm_check(rxm_attr_move(m_attr, (int)source, (int)target, stream));
        refresh_pointers();  // a missing target side was allocated (attribute.cu:348-353)
    }
    void copy_from(Attribute<T, HandleT>& source, locationT source_flag, locationT dst_flag, cudaStream_t stream = NULL)
    {
        detail::rxm_check(rxm_attr_copy_from(m_attr, source.m_attr, (int)source_flag, (int)dst_flag, stream));
    }
    // release(location) (attribute.cu:375-390): frees only the requested side(s); the handle and the name go with the
    // last side (shallow copies handed to kernels keep dangling pointers, exactly as in the reference)
    void release(locationT location = LOCATION_ALL) override
    {
        if (!m_attr) return;
        rxm_attr_release(m_attr, (int)location);
        refresh_pointers();
        if (!m_h && !m_d) {
            rxm_attr_destroy(m_attr);
            free(m_name);
            m_attr = nullptr, m_name = nullptr;
        }
    }

    // Attribute::operator()(handle, attr) (attribute.h:313-319,406-434)
    __host__ __device__ __forceinline__ T& operator()(const HandleT handle, const uint32_t attr = 0) const
    {
#ifdef __CUDA_ARCH__
        return m_d[index(m_d_slot_base, m_d_lin_base, handle.patch_id(), handle.local_id(), attr)];
#else
        return m_h[index(m_h_slot_base, m_h_lin_base, handle.patch_id(), handle.local_id(), attr)];
#endif
    }
    // operator()(i, j): element with LINEAR id i, attribute j (attribute.h:109-111), any layout
    __host__ __device__ __forceinline__ T& operator()(const size_t i, const size_t j = 0) const
    {
#ifdef __CUDA_ARCH__
        const uint32_t* lb = m_d_lin_base;
        T*              base = m_d;
        const uint32_t* sb = m_d_slot_base;
#else
        const uint32_t* lb = m_h_lin_base;
        T*              base = m_h;
        const uint32_t* sb = m_h_slot_base;
#endif
        if (m_layout == SoA) return base[j * m_num_elems + i];
        uint32_t lo = 0, hi = m_num_patches;  // the patch whose linear-id range holds i
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) / 2;
            if (lb[mid] <= i) lo = mid; else hi = mid;
        }
        return base[index(sb, lb, lo, (uint32_t)(i - lb[lo]), (uint32_t)j)];
    }
    // to_matrix / from_matrix (attribute.h:125-145): the HOST copy as a #elements x #attributes matrix whose row i is the
    // element with linear id i, and back; any layout
    template <int Order = 0>
    std::shared_ptr<DenseMatrix<T, Order>> to_matrix() const
    {
        auto mat = std::make_shared<DenseMatrix<T, Order>>(m_num_elems, m_nattr);
        for (uint32_t i = 0; i < m_num_elems; ++i)
            for (uint32_t j = 0; j < m_nattr; ++j)
                (*mat)(i, j) = (*this)((size_t)i, (size_t)j);
        return mat;
    }
    template <int Order>
    void from_matrix(DenseMatrix<T, Order>* mat)
    {
        for (uint32_t i = 0; i < m_num_elems; ++i)
            for (uint32_t j = 0; j < m_nattr; ++j)
                (*this)((size_t)i, (size_t)j) = (*mat)(i, j);
    }
    template <int N>
    __host__ __device__ __forceinline__ glm::vec<N, T> to_glm(const HandleT& handle) const
    {
        glm::vec<N, T> r;
        for (int i = 0; i < N; ++i) r[i] = (*this)(handle, i);
        return r;
    }
    template <int N>
    __host__ __device__ __forceinline__ void from_glm(const HandleT& handle, const glm::vec<N, T>& in) const
    {
        for (int i = 0; i < N; ++i) (*this)(handle, i) = in[i];
    }

   private:
    __host__ __device__ const uint32_t* lin_base() const
    {
#ifdef __CUDA_ARCH__
        return m_d_lin_base;
#else
        return m_h_lin_base;
#endif
    }
    __host__ __device__ __forceinline__ uint64_t index(const uint32_t* sb, const uint32_t* lb, uint32_t p, uint32_t lid,
                                                       uint32_t a) const
    {
        if (m_layout == SoA) return (uint64_t)a * m_num_elems + lb[p] + lid;  // attribute.h:406-421
        const uint32_t b = sb[p];
        if (m_layout == AoS) return (uint64_t)(b + lid) * m_nattr + a;
        return (uint64_t)b * m_nattr + (uint64_t)a * (sb[p + 1] - b) + lid;
    }
    char*                     m_name = nullptr;
    rxm_attr*                 m_attr = nullptr;
    void refresh_pointers()
    {
        m_h        = m_attr ? (T*)rxm_attr_data(m_attr, RXM_HOST) : nullptr;
        m_d        = m_attr ? (T*)rxm_attr_data(m_attr, RXM_DEVICE) : nullptr;
        m_location = (locationT)((m_h ? HOST : 0) | (m_d ? DEVICE : 0));
    }
    T *                       m_h = nullptr, *m_d = nullptr;
    const uint32_t *          m_h_slot_base = nullptr, *m_d_slot_base = nullptr;
    const uint32_t *          m_h_lin_base = nullptr, *m_d_lin_base = nullptr;
    uint64_t                  m_storage = 0;
    uint32_t                  m_num_slots = 0, m_num_elems = 0, m_num_patches = 0, m_nattr = 0;
    layoutT                   m_layout   = AoSoA;
    locationT                 m_location = LOCATION_NONE;
};

template <class T> using VertexAttribute = Attribute<T, VertexHandle>;
template <class T> using EdgeAttribute   = Attribute<T, EdgeHandle>;
template <class T> using FaceAttribute   = Attribute<T, FaceHandle>;
}  // namespace rxmesh
