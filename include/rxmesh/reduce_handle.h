// rxmesh/reduce_handle.h -- ReduceHandle<T, HandleT> (include/rxmesh/reduce_handle.h:22-204, reduce_handle.cu:53-156,
// kernels/reduce.cuh:43-191, arg_ops.h:7-45): dot / norm2 / arg_min / arg_max / generic reduce over the OWNED elements of
// an attribute, result returned on the host.  Header-only here: two small kernels (one block per group of patches, then
// one block over the partial results) templated on the value type and the reduction functor, so cub::Max / cub::Sum /
// user functors work for every attribute type like in the reference.  Sums for dot / norm2 are accumulated in double.
#pragma once
#include <cub/cub.cuh>

#include <limits>
#include <memory>

#include "rxmesh/attribute.h"

namespace rxmesh {
template <typename HandleT, typename T>
using KeyValuePair = cub::KeyValuePair<HandleT, T>;

namespace detail {
constexpr uint32_t reduce_block = 256;

// stage 1: block b folds map(handle, attribute id) over the owned elements of patches b, b + grid, ...
template <typename AccT, typename HandleT, typename MapF, typename RedF>
__global__ static void reduce_stage1(uint32_t num_patches, const uint32_t* __restrict__ lin_base, uint32_t num_attr,
                                     uint32_t attribute_id, AccT init, MapF map, RedF red, AccT* __restrict__ partial)
{
    using BlockReduce = cub::BlockReduce<AccT, reduce_block>;
    __shared__ typename BlockReduce::TempStorage tmp;
    AccT                                         acc = init;
    for (uint32_t p = blockIdx.x; p < num_patches; p += gridDim.x) {
        const uint32_t n  = lin_base[p + 1] - lin_base[p];
        const uint32_t na = attribute_id == INVALID32 ? num_attr : 1u;
        for (uint32_t i = threadIdx.x; i < n * na; i += reduce_block) {
            const uint32_t lid = i % n, a = attribute_id == INVALID32 ? i / n : attribute_id;
            acc = red(acc, map(HandleT(p, typename HandleT::LocalT((uint16_t)lid)), a));
        }
    }
    acc = BlockReduce(tmp).Reduce(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
// stage 2: one block over the partial results; the total lands behind them (partial[n])
template <typename AccT, typename RedF>
__global__ static void reduce_stage2(uint32_t n, AccT init, RedF red, AccT* __restrict__ partial)
{
    using BlockReduce = cub::BlockReduce<AccT, reduce_block>;
    __shared__ typename BlockReduce::TempStorage tmp;
    AccT                                         acc = init;
    for (uint32_t i = threadIdx.x; i < n; i += reduce_block)
        acc = red(acc, partial[i]);
    acc = BlockReduce(tmp).Reduce(acc, red);
    if (threadIdx.x == 0) partial[n] = acc;
}
struct SumD
{
    __device__ __forceinline__ double operator()(double a, double b) const { return a + b; }
};
// what stage 1 folds: functors rather than extended lambdas (nvcc does not allow those in private member functions)
template <typename T, typename HandleT>
struct DotMap
{
    Attribute<T, HandleT> x, y;
    __device__ __forceinline__ double operator()(const HandleT h, uint32_t a) const { return (double)x(h, a) * (double)y(h, a); }
};
template <typename T, typename HandleT>
struct ValueMap
{
    Attribute<T, HandleT> x;
    __device__ __forceinline__ T operator()(const HandleT h, uint32_t a) const { return x(h, a); }
};
template <typename T, typename HandleT>
struct KeyValueMap
{
    Attribute<T, HandleT> x;
    __device__ __forceinline__ KeyValuePair<HandleT, T> operator()(const HandleT h, uint32_t a) const
    {
        return KeyValuePair<HandleT, T>(h, x(h, a));
    }
};
// arg ops (arg_ops.h:12-45); ties go to the smaller handle so the result does not depend on the block schedule
template <typename HandleT, typename T>
struct ArgMaxOp
{
    constexpr T default_val() const { return std::numeric_limits<T>::lowest(); }
    __device__ __forceinline__ KeyValuePair<HandleT, T> operator()(const KeyValuePair<HandleT, T>& a,
                                                                   const KeyValuePair<HandleT, T>& b) const
    {
        return (b.value > a.value || (b.value == a.value && b.key.unique_id() < a.key.unique_id())) ? b : a;
    }
};
template <typename HandleT, typename T>
struct ArgMinOp
{
    constexpr T default_val() const { return std::numeric_limits<T>::max(); }
    __device__ __forceinline__ KeyValuePair<HandleT, T> operator()(const KeyValuePair<HandleT, T>& a,
                                                                   const KeyValuePair<HandleT, T>& b) const
    {
        return (b.value < a.value || (b.value == a.value && b.key.unique_id() < a.key.unique_id())) ? b : a;
    }
};
}  // namespace detail

template <typename T, typename HandleT>
class ReduceHandle
{
   public:
    using HandleType = HandleT;
    using Type       = T;
    using KeyValue   = KeyValuePair<HandleT, T>;

    ReduceHandle()                    = default;
    ReduceHandle(const ReduceHandle&) = default;
    // allocates the partial-result buffer used by every reduction (reduce_handle.h:40-59)
    ReduceHandle(const Attribute<T, HandleT>& attr) : ReduceHandle(attr.get_num_patches()) {}
    ReduceHandle(uint32_t num_patches) : m_num_patches(num_patches)
    {
        m_grid  = num_patches < 148u * 8u ? (num_patches ? num_patches : 1u) : 148u * 8u;
        void* d = nullptr;
        if (cudaMalloc(&d, ((size_t)m_grid + 1) * 32) != cudaSuccess) {
            fprintf(stderr, "rxmesh_b200: ReduceHandle: cudaMalloc failed\n");
            exit(EXIT_FAILURE);
        }
        m_partial = std::shared_ptr<void>(d, [](void* p) { cudaFree(p); });
    }

    T dot(const Attribute<T, HandleT>& attr1, const Attribute<T, HandleT>& attr2, uint32_t attribute_id = INVALID32,
          cudaStream_t stream = NULL)
    {
        check_device(attr1, "dot"), check_device(attr2, "dot");
        return (T)run<double>(attr1, attribute_id, 0.0, detail::DotMap<T, HandleT>{attr1, attr2}, detail::SumD(), stream);
    }
    T norm2(const Attribute<T, HandleT>& attr, uint32_t attribute_id = INVALID32, cudaStream_t stream = NULL)
    {
        check_device(attr, "norm2");
        return (T)std::sqrt(run<double>(attr, attribute_id, 0.0, detail::DotMap<T, HandleT>{attr, attr}, detail::SumD(), stream));
    }
    KeyValue arg_max(const Attribute<T, HandleT>& attr, uint32_t attribute_id = INVALID32, cudaStream_t stream = NULL)
    {
        return arg(attr, detail::ArgMaxOp<HandleT, T>(), attribute_id, stream);
    }
    KeyValue arg_min(const Attribute<T, HandleT>& attr, uint32_t attribute_id = INVALID32, cudaStream_t stream = NULL)
    {
        return arg(attr, detail::ArgMinOp<HandleT, T>(), attribute_id, stream);
    }
    // generic reduction with a CUB-style binary functor and its neutral element (reduce_handle.h:140-166)
    template <typename ReductionOp>
    T reduce(const Attribute<T, HandleT>& attr, ReductionOp reduction_op, T init, uint32_t attribute_id = INVALID32,
             cudaStream_t stream = NULL)
    {
        check_device(attr, "reduce");
        return run<T>(attr, attribute_id, init, detail::ValueMap<T, HandleT>{attr}, reduction_op, stream);
    }

   private:
    static void check_device(const Attribute<T, HandleT>& attr, const char* who)
    {
        if ((attr.get_allocated() & DEVICE) != DEVICE) {  // RXMESH_ERROR in the reference: log, the launch then faults
            fprintf(stderr, "rxmesh_b200: ReduceHandle::%s() input attribute should be allocated on the device\n", who);
            exit(EXIT_FAILURE);
        }
    }
    template <typename ArgOp>
    KeyValue arg(const Attribute<T, HandleT>& attr, ArgOp op, uint32_t attribute_id, cudaStream_t stream)
    {
        check_device(attr, "arg_min/arg_max");
        if (attribute_id == INVALID32 && attr.get_num_attributes() > 1) {
            fprintf(stderr, "rxmesh_b200: ReduceHandle::arg_min/arg_max need an attribute_id for multi-component attributes\n");
            exit(EXIT_FAILURE);
        }
        const uint32_t a0 = attribute_id == INVALID32 ? 0u : attribute_id;
        return run<KeyValue>(attr, a0, KeyValue(HandleT(), op.default_val()), detail::KeyValueMap<T, HandleT>{attr}, op, stream);
    }
    template <typename AccT, typename MapF, typename RedF>
    AccT run(const Attribute<T, HandleT>& attr, uint32_t attribute_id, AccT init, MapF map, RedF red, cudaStream_t stream)
    {
        static_assert(sizeof(AccT) <= 32, "partial-result slots hold 32 bytes");
        AccT* partial = reinterpret_cast<AccT*>(m_partial.get());
        detail::reduce_stage1<AccT, HandleT><<<m_grid, detail::reduce_block, 0, stream>>>(
            m_num_patches, attr.lin_base(DEVICE), attr.get_num_attributes(), attribute_id, init, map, red, partial);
        detail::reduce_stage2<AccT><<<1, detail::reduce_block, 0, stream>>>(m_grid, init, red, partial);
        AccT out = init;
        if (cudaMemcpyAsync(&out, partial + m_grid, sizeof(AccT), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
            cudaStreamSynchronize(stream) != cudaSuccess) {
            fprintf(stderr, "rxmesh_b200: ReduceHandle: %s\n", cudaGetErrorString(cudaGetLastError()));
            exit(EXIT_FAILURE);
        }
        return out;
    }
    uint32_t              m_num_patches = 0, m_grid = 0;
    std::shared_ptr<void> m_partial;
};
template <typename T, typename HandleT>
ReduceHandle(const Attribute<T, HandleT>&) -> ReduceHandle<T, HandleT>;
}  // namespace rxmesh
