// rxmesh/query.h -- Query<blockThreads> (include/rxmesh/query.h:15-137, query.inl:107-260): dispatch<op>
// loads the patch sections by TMA, builds the adjacency in shared memory and calls the user's device lambda
// (InputHandle, OutputIterator&) once per OWNED source element that passes the active-set predicate.
#pragma once
#include <cooperative_groups.h>
#include <cstdio>
#include <type_traits>
#include "rxmesh/context.h"
#include "rxmesh/iterator.cuh"
#include "rxmesh/kernels/shmem_allocator.cuh"

namespace rxmesh {
namespace detail {
// the one mbarrier every query of a kernel loads through (a function-scope __shared__ object: one per kernel)
__device__ __forceinline__ uint64_t* query_barrier()
{
    __shared__ uint64_t bar;
    return &bar;
}
// first / second parameter types of a compute lambda (the reference's FunctionTraits, util/meta.h)
template <typename T>
struct LambdaArgs : LambdaArgs<decltype(&T::operator())> {};
template <typename C, typename R, typename A0, typename A1>
struct LambdaArgs<R (C::*)(A0, A1) const>
{
    using Arg0 = std::remove_cv_t<std::remove_reference_t<A0>>;
    using Arg1 = std::remove_cv_t<std::remove_reference_t<A1>>;
};
template <typename C, typename R, typename A0, typename A1>
struct LambdaArgs<R (C::*)(A0, A1)>
{
    using Arg0 = std::remove_cv_t<std::remove_reference_t<A0>>;
    using Arg1 = std::remove_cv_t<std::remove_reference_t<A1>>;
};
}  // namespace detail
template <uint32_t blockThreads>
struct Query
{
    __device__ Query(const Context& context, const uint32_t pid = blockIdx.x)
        : m_ctx(context), m_pid(pid), m_desc(load(context.view.desc + pid)) {}
    Query(const Query&)            = delete;
    Query& operator=(const Query&) = delete;
    __device__ int                  get_patch_id() const { return (int)m_desc.patch_id; }
    __device__ const rxm::PatchDesc& get_patch_info() const { return m_desc; }

    template <Op op, typename computeT>
    __device__ void dispatch(cooperative_groups::thread_block& block, ShmemAllocator& shrd_alloc, computeT compute_op,
                             const bool oriented = false)
    {
        using InH = typename InputHandle<op>::type;
        dispatch<op>(block, shrd_alloc, compute_op, [](InH) { return true; }, oriented);
    }

    template <Op op, typename computeT, typename activeSetT>
    __device__ void dispatch(cooperative_groups::thread_block& block, ShmemAllocator& shrd_alloc, computeT compute_op,
                             activeSetT compute_active_set, const bool oriented = false, const bool allow_not_owned = false)
    {
        (void)block;
        if (m_ctx.view.packed)
            run<op, true>(shrd_alloc, compute_op, compute_active_set, oriented, allow_not_owned);
        else
            run<op, false>(shrd_alloc, compute_op, compute_active_set, oriented, allow_not_owned);
    }

    // ---- the split form of dispatch (query.h:83-127, query.inl:178-260): prologue builds the adjacency and keeps it in
    // shared memory, run_compute / get_iterator read it, epilogue frees it.  prologue's allow_not_owned defaults to
    // true as in the reference: lists exist for not-owned (ribbon) sources too.
    template <Op op>
    __device__ void prologue(cooperative_groups::thread_block& block, ShmemAllocator& shrd_alloc, const bool oriented = false,
                             const bool allow_not_owned = true)
    {
        using InH = typename InputHandle<op>::type;
        prologue<op>(block, shrd_alloc, [](InH) { return true; }, oriented, allow_not_owned);
    }
    template <Op op, typename activeSetT>
    __device__ void prologue(cooperative_groups::thread_block& block, ShmemAllocator& shrd_alloc, activeSetT compute_active_set,
                             const bool oriented = false, const bool allow_not_owned = true)
    {
        (void)block;
        using InH  = typename InputHandle<op>::type;
        m_op       = op;
        m_used0    = m_ctx.view.packed ? build<op, true>(shrd_alloc, oriented, allow_not_owned, m_r, m_ot)
                                       : build<op, false>(shrd_alloc, oriented, allow_not_owned, m_r, m_ot);
        // participant bitmask (query_dispatcher.cuh:27-177): owned sources that pass the predicate, plus every
        // not-owned source when allow_not_owned
        const uint32_t words = (m_r.n_src + 31u) / 32u;
        m_participant        = shrd_alloc.template alloc<uint32_t>(words ? words : 1u);
        // one source per lane, one ballot per word (blockThreads is a multiple of 32, so the 32 sources of a word always
        // sit in the lanes of one warp; every lane of the warp runs the same number of rounds)
        static_assert(blockThreads % 32u == 0, "Query<blockThreads>: whole warps");
        for (uint32_t s0 = threadIdx.x & ~31u; s0 < 32u * words; s0 += blockThreads) {
            const uint32_t s  = s0 + (threadIdx.x & 31u);
            const bool     in = s < m_r.n_src && (s >= m_desc.n_owned[InH::elem] ||
                                              compute_active_set(InH(m_desc.patch_id, typename InH::LocalT((uint16_t)s))));
            const uint32_t bits = __ballot_sync(0xFFFFFFFFu, in);
            if ((threadIdx.x & 31u) == 0) m_participant[s0 >> 5] = bits;
        }
        __syncthreads();
    }
    template <typename computeT>
    __device__ void run_compute(cooperative_groups::thread_block& block, computeT compute_op)
    {
        (void)block;
        using Traits = detail::LambdaArgs<computeT>;
        using InH    = typename Traits::Arg0;
        using ItT    = typename Traits::Arg1;
        for (uint32_t s = threadIdx.x; s < m_r.n_src; s += blockThreads) {
            if (!((m_participant[s >> 5] >> (s & 31u)) & 1u)) continue;
            InH h(m_desc.patch_id, typename InH::LocalT((uint16_t)s));
            ItT it(m_r, m_ot, s);
            compute_op(h, it);
        }
    }
    template <typename IteratorT>
    __device__ IteratorT get_iterator(uint16_t local_id) const
    {
        return IteratorT(m_r, m_ot, local_id);
    }
    __device__ void epilogue(cooperative_groups::thread_block& block, ShmemAllocator& shrd_alloc)
    {
        (void)block;
        release(shrd_alloc, m_used0);
        m_participant = nullptr;
    }
    // compute_vertex_valence / vertex_valence (query.inl:25-74): edges incident to every local vertex of the patch,
    // counted from EV in shared memory (stays allocated; the reference keeps it for the life of the kernel too)
    __device__ void compute_vertex_valence(cooperative_groups::thread_block& block, ShmemAllocator& shrd_alloc)
    {
        (void)block;
        const uint32_t nv = m_desc.n[rxm::ELEM_V], ne = m_desc.n[rxm::ELEM_E];
        m_valence         = shrd_alloc.template alloc<uint32_t>(nv ? nv : 1u);
        for (uint32_t v = threadIdx.x; v < nv; v += blockThreads)
            m_valence[v] = 0;
        __syncthreads();
        const uint32_t* ev   = reinterpret_cast<const uint32_t*>(m_ctx.view.topo + m_desc.topo_off + m_desc.off_ev());
        const uint32_t  mask = m_ctx.view.packed ? rxm::PK_ID_MASK : 0xFFFFu;
        for (uint32_t e = threadIdx.x; e < ne; e += blockThreads) {
            const uint32_t w = __ldg(ev + e);
            atomicAdd(&m_valence[w & mask], 1u);
            atomicAdd(&m_valence[(w >> 16) & mask], 1u);
        }
        __syncthreads();
    }
    __device__ uint16_t vertex_valence(uint16_t v) const { return (uint16_t)m_valence[v]; }
    __device__ uint16_t vertex_valence(VertexHandle vh) const { return (uint16_t)m_valence[vh.local_id()]; }

   private:
    static __device__ rxm::PatchDesc load(const rxm::PatchDesc* g)
    {
        rxm::PatchDesc d;
        const uint4*   s = reinterpret_cast<const uint4*>(g);
        uint4*         t = reinterpret_cast<uint4*>(&d);
#pragma unroll
        for (int i = 0; i < (int)(sizeof(rxm::PatchDesc) / 16); ++i)
            t[i] = __ldg(s + i);
        return d;
    }

    // builds the adjacency of this patch in shared memory (TMA loads + transpose / fans); leaves the result view
    // and the owner table; all threads of the block call it; returns the allocator mark to restore in release()
    // round < 0: a stand-alone query -- the barrier is initialised, used for one phase and invalidated (an mbarrier must
    // be invalidated before its storage is initialised again).  round >= 0: one of a series of queries of the same kernel
    // call site (higher_query_block_dispatcher): initialised in round 0, later rounds wait on the next phase parity, the
    // caller invalidates it after the last round (end_rounds).
    template <Op op, bool PACKED>
    __device__ uint32_t build(ShmemAllocator& sa, bool oriented, bool all_sources, rxm::dev::QueryResult& r,
                              rxm::dev::OwnerTable& ot, int round = -1)
    {
        using namespace rxm::dev;
        constexpr int OPV = (int)op;
        uint64_t*     barp = detail::query_barrier();
        uint64_t&     bar  = *barp;
        __shared__ uint32_t warp_tmp[36];
        const rxm::PatchDesc& d    = m_desc;
        const uint8_t*        blob = m_ctx.view.topo + d.topo_off;
        const uint32_t        used0 = sa.m_sm.used;
        // oriented VV / VE (orient_edges_around_vertices, kernels/rxmesh_queries.cuh:375-499,905-914): served by the stored
        // one-ring fans (fan_v / fan_e).  The reference can only orient on edge-manifold input and asserts inside its
        // kernels; here a mesh without fans (non-manifold or inconsistently oriented input, RXM_NO_FANS) or a request for
        // the lists of not-owned sources stops the kernel with a message instead of silently handing out unoriented lists.
        if ((op == Op::VV || op == Op::VE) && oriented && (!m_ctx.view.fans || all_sources)) {
            if (threadIdx.x == 0 && blockIdx.x == 0)
                printf("rxmesh: oriented %s needs %s\n", op == Op::VV ? "Op::VV" : "Op::VE",
                       all_sources ? "allow_not_owned = false (the one-ring fans cover owned vertices only)"
                                   : "an edge-manifold, consistently oriented input mesh (no one-ring fans were stored)");
            __syncthreads();
            __trap();
        }
        const bool use_fans = (op == Op::VV || op == Op::VE) && oriented;
        PatchQuery<OPV, (int)blockThreads, 12, PACKED> q;
        uint16_t *s_fo = nullptr, *s_fv = nullptr;
        constexpr uint32_t fan_dst = op == Op::VE ? rxm::ELEM_E : rxm::ELEM_V;  // fan_e names edges, fan_v vertices
        uint32_t* s_own = nullptr;
        rxm::StashEntry* s_stash = nullptr;
        if (use_fans) {
            s_fo    = sa.template alloc<uint16_t>(d.fanoff_bytes() / 2);
            s_fv    = sa.template alloc<uint16_t>(d.fanv_bytes() / 2);
            s_own   = sa.template alloc<uint32_t>(d.own_bytes(fan_dst) / 4);
            s_stash = sa.template alloc<rxm::StashEntry>(d.n_stash);
        } else {
            q.plan(d, sa.m_sm, true, all_sources, m_ctx.view.edge_manifold != 0);
        }
        if (threadIdx.x == 0) {
            if (round <= 0) {
                mbar_init(&bar, 1);
                fence_mbar_init();
            }
            if (use_fans) {
                mbar_arrive_expect_tx(&bar, d.fanoff_bytes() + d.fanv_bytes() + d.own_bytes(fan_dst) + d.stash_bytes());
                bulk_g2s(s_fo, blob + d.off_fanoff(), d.fanoff_bytes(), &bar);
                if (d.fanv_bytes()) bulk_g2s(s_fv, blob + (op == Op::VE ? d.off_fane() : d.off_fanv()), d.fanv_bytes(), &bar);
                if (d.own_bytes(fan_dst)) bulk_g2s(s_own, blob + d.off_own(fan_dst), d.own_bytes(fan_dst), &bar);
                if (d.stash_bytes()) bulk_g2s(s_stash, blob + d.off_stash(), d.stash_bytes(), &bar);
            } else {
                mbar_arrive_expect_tx(&bar, q.tx_bytes(d, true));
                q.issue(d, blob, &bar, true);
            }
        }
        __syncthreads();
        mbar_wait(&bar, round > 0 ? (uint32_t)(round & 1) : 0u);
        if (round < 0) {
            __syncthreads();
            if (threadIdx.x == 0) mbar_inval(&bar);  // a kernel may run several queries: the next one initialises it again
        }
        if (use_fans) {
            // fan_off entries carry the closed flag in bit 15: strip it once so the list bounds are plain
            for (uint32_t i = threadIdx.x; i <= d.n_owned[rxm::ELEM_V]; i += blockThreads)
                s_fo[i] &= rxm::FAN_OFF_MASK;
            __syncthreads();
            r.off16 = s_fo, r.off32 = nullptr, r.cnt = nullptr, r.end16 = nullptr, r.val = s_fv, r.stride = 0, r.shift = 0, r.mask = 0xFFFFu;
            r.n_src = d.n_owned[rxm::ELEM_V];
            ot.own = s_own, ot.stash = s_stash, ot.n_owned = d.n_owned[fan_dst], ot.patch = d.patch_id;
            ot.slot_base = d.slot_base[fan_dst], ot.type = fan_dst;
        } else {
            r = q.compute(d, warp_tmp, all_sources, false);
            if (op_is_fixed<OPV>()) __syncthreads();
            ot = q.owner_table(d);
        }
        return used0;
    }
    // epilogue: every thread is done with the result; release the query's shared memory (query.inl:74-91)
    __device__ void release(ShmemAllocator& sa, uint32_t used0)
    {
        rxm::dev::fence_proxy_async();  // generic-proxy accesses to this shared memory before the next query's TMA writes
        __syncthreads();
        sa.m_sm.used = used0;
    }

    template <Op op, bool PACKED, typename computeT, typename activeSetT>
    __device__ void run(ShmemAllocator& sa, computeT& compute_op, activeSetT& active, bool oriented, bool all_sources)
    {
        using InH  = typename InputHandle<op>::type;
        using OutH = typename OutputHandle<op>::type;
        rxm::dev::QueryResult r;
        rxm::dev::OwnerTable  ot;
        const uint32_t        used0 = build<op, PACKED>(sa, oriented, all_sources, r, ot);
        for (uint32_t s = threadIdx.x; s < r.n_src; s += blockThreads) {
            InH h(m_desc.patch_id, typename InH::LocalT((uint16_t)s));
            if (!active(h)) continue;
            Iterator<OutH> it(r, ot, s);
            compute_op(h, it);
        }
        release(sa, used0);
    }

   public:
    // The query of THIS patch answered for one source element per thread (the element may differ per thread, threads
    // without one pass has_src = false): the building block of higher_query_block_dispatcher
    // (kernels/query_dispatcher.cuh:445-565), called by the whole block.
    // after the last round of a series (build with round >= 0): every thread has passed its waits
    static __device__ void end_rounds()
    {
        __syncthreads();
        if (threadIdx.x == 0) rxm::dev::mbar_inval(detail::query_barrier());
    }
    template <Op op, typename computeT>
    __device__ void dispatch_src(ShmemAllocator& sa, bool has_src, typename InputHandle<op>::type src, computeT& compute_op,
                                 bool oriented, int round)
    {
        using OutH = typename OutputHandle<op>::type;
        rxm::dev::QueryResult r;
        rxm::dev::OwnerTable  ot;
        const uint32_t used0 = m_ctx.view.packed ? build<op, true>(sa, oriented, false, r, ot, round)
                                                 : build<op, false>(sa, oriented, false, r, ot, round);
        if (has_src && src.local_id() < r.n_src) {
            Iterator<OutH> it(r, ot, src.local_id());
            compute_op(src, it);
        }
        release(sa, used0);
    }

   private:
    const Context&       m_ctx;
    uint32_t             m_pid;
    const rxm::PatchDesc m_desc;
    // state between prologue and epilogue
    rxm::dev::QueryResult m_r{};
    rxm::dev::OwnerTable  m_ot{};
    uint32_t              m_used0       = 0;
    uint32_t*             m_participant = nullptr;
    uint32_t*             m_valence     = nullptr;
    Op                    m_op          = Op::INVALID;
};
}  // namespace rxmesh
