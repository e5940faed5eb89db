// rxm_query.cuh -- the eight patch-local static queries, computed in shared memory.
//
// B200-native counterpart of detail::query<blockThreads, op> and its helpers
// v_v / v_e / v_f / e_f / f_v / f_f
// (/root/reference/include/rxmesh/kernels/rxmesh_queries.cuh:501-1053) and of
// detail::query_block_dispatcher (kernels/query_dispatcher.cuh:27-177).
//
// One thread block handles one patch.  PatchQuery<OP> knows which sections of
// the patch blob the op needs, asks thread 0 to TMA them into shared memory
// under the caller's mbarrier (so the caller can batch its own loads, e.g. the
// patch's attribute slice, into the same phase), and then builds the adjacency:
//   EV, FV, FE : the stored rows themselves (fixed stride 2 / 3 / 3); only the
//                OWNED prefix of the rows is loaded;
//   VV, VE     : the stored one-ring fans (fan_v / fan_e: plain read, oriented order) when the mesh has them, else the
//                transpose of EV (VV stores the other endpoint directly);
//   VF         : transpose of FV;   EF : transpose of FE;
//   FF         : EF, then per owned face the faces across its three edges, in
//                edge order (the reference's order on manifold input,
//                rxmesh_queries.cuh:839-853).  Edge-manifold packed meshes take a pair
//                table instead (no CSR, no scan);
//   EVDiamond  : per edge [v0, w0, v1, w1]: its end vertices and the vertices opposite
//                to it in the face that traverses it v0->v1 (w0) / v1->v0 (w1);
//                fixed stride 4, 0xFFFF where a face is missing
//                (e_v_diamond, rxmesh_queries.cuh:198-264);
//   EE         : per edge, for the face on side dir = 0 / 1: [next edge, previous
//                edge] in that face's winding; fixed stride 4, 0xFFFF on a mesh
//                boundary (e_e_manifold, rxmesh_queries.cuh:267-342; the reference
//                indexes the previous edge with (cur - 1) % 3, which is -1 for
//                cur = 0 in C++ -- the intended (cur + 2) % 3 is used here).
// Two transposition paths:
//   PACKED : every incidence entry carries its rank inside the transposed list
//            and the list offsets are stored with the patch (patch_layout.h), so
//            the transpose is ONE pass of plain shared-memory stores
//            out[offset[col] + rank] = row -- no atomics, no scan, deterministic
//            ascending order;
//   wide   : one shared atomic per non-zero (rank kept in a register) + a
//            hand-written shuffle scan (optionally followed by a per-list sort).
// Only columns of OWNED source elements are built unless `all_sources` is set
// (the reference's allow_not_owned, query.inl:174-203).
#pragma once
#include "rxm_device.cuh"

namespace rxm {
namespace dev {

template <int OP>
struct OpTraits;
#define RXM_OP_TRAITS(OPV, SRC, DST, CONN)            \
    template <>                                       \
    struct OpTraits<OPV>                              \
    {                                                 \
        static constexpr uint32_t src  = SRC;         \
        static constexpr uint32_t dst  = DST;         \
        static constexpr int      conn = CONN; /* 0 = EV, 1 = FE, 2 = FV */ \
    };
RXM_OP_TRAITS(OP_VV, ELEM_V, ELEM_V, 0)
RXM_OP_TRAITS(OP_VE, ELEM_V, ELEM_E, 0)
RXM_OP_TRAITS(OP_VF, ELEM_V, ELEM_F, 2)
RXM_OP_TRAITS(OP_EV, ELEM_E, ELEM_V, 0)
RXM_OP_TRAITS(OP_EF, ELEM_E, ELEM_F, 1)
RXM_OP_TRAITS(OP_FV, ELEM_F, ELEM_V, 2)
RXM_OP_TRAITS(OP_FE, ELEM_F, ELEM_E, 1)
RXM_OP_TRAITS(OP_FF, ELEM_F, ELEM_F, 1)
RXM_OP_TRAITS(OP_EE, ELEM_E, ELEM_E, 1)         // the four edges of the two faces at an edge (edge-manifold input)
RXM_OP_TRAITS(OP_EVDIAMOND, ELEM_E, ELEM_V, 1)  // the two end vertices + the two opposite vertices
#undef RXM_OP_TRAITS

template <int OP>
__host__ __device__ constexpr bool op_is_edge4()
{
    return OP == OP_EE || OP == OP_EVDIAMOND;
}

template <int OP>
__host__ __device__ constexpr bool op_is_fixed()
{
    return OP == OP_EV || OP == OP_FV || OP == OP_FE;
}

// View over the result of a query for one patch (the data behind the
// reference's Iterator, iterator.cuh:50-194).
struct QueryResult
{
    const uint16_t* off16;   // CSR offsets (shared, packed path / stored offsets)
    const uint32_t* off32;   // CSR offsets (shared, atomic path and FF)
    const uint16_t* val;     // neighbour local ids (shared)
    uint32_t        stride;  // 2 / 3 for EV / FV, FE when both off pointers are null
    uint32_t        shift;   // 1 for FE (drops the direction bit), else 0
    uint32_t        mask;    // id mask applied after the shift
    uint32_t        n_src;   // number of source elements with a list
    const uint16_t* cnt;     // fixed stride with per-row fill count (FF on edge-manifold input), else null
    const uint16_t* end16;   // explicit list ends next to off16 (VF through the fan faces: open fans end one slot early)

    __device__ __forceinline__ uint32_t begin(uint32_t s) const
    {
        return off16 ? (uint32_t)off16[s] : (off32 ? off32[s] : s * stride);
    }
    __device__ __forceinline__ uint32_t end(uint32_t s) const
    {
        return end16 ? (uint32_t)end16[s]
                     : (off16 ? (uint32_t)off16[s + 1]
                              : (off32 ? off32[s + 1] : (cnt ? s * stride + (uint32_t)cnt[s] : (s + 1) * stride)));
    }
    __device__ __forceinline__ uint32_t size(uint32_t s) const { return end(s) - begin(s); }
    __device__ __forceinline__ uint32_t at(uint32_t pos) const { return ((uint32_t)val[pos] >> shift) & mask; }
};

template <int OP, int BT, int KMAX, bool PACKED>
struct PatchQuery
{
    using Tr = OpTraits<OP>;
    static constexpr uint32_t W       = Tr::conn == 0 ? 2 : 3;
    static constexpr uint32_t ID_MASK = PACKED ? PK_ID_MASK : 0xFFFFu;
    uint16_t*   s_conn;
    uint16_t*   s_loff;  // packed: stored list offsets
    uint32_t*   s_off;   // wide: counters / offsets
    uint16_t*   s_val;
    uint32_t*   s_off2;  // FF only
    uint16_t*   s_val2;  // FF only
    uint32_t*   s_own;
    StashEntry* s_stash;
    uint32_t    conn_bytes, loff_bytes;
    uint32_t    n_rows;  // rows of the connectivity section that get loaded
    uint32_t    n_cols;  // columns whose lists are built
    bool        ff2;     // FF on edge-manifold input, packed format: pair table instead of the EF CSR
    bool        ff3;     // FF read from the stored rows (patch_layout.h FLAG_FF): plain read + per-row count
    bool        ef3;     // EF read from the stored pairs of the owned edges (same flag): plain read + per-row count
    bool        fanq;    // VV / VE / VF read from the stored one-ring fans (FLAG_FANS): plain read, oriented order

    // shared-memory bytes this op needs for a patch with the given maxima
    // (host side; the role of calc_shared_memory, rxmesh_static.inl:498-841)
    // stored_rows != 0: every patch answers FF / EF from its stored rows (plan(): ff3 / ef3) and owns at most that many
    // faces / edges; the launch then reserves the rows + counts instead of the transposes' scratch (FF: 5 -> 8 resident
    // blocks per SM)
    // VV / VE / VF with fan_entries != 0 (every patch stores fans): fan offsets of at most stored_rows owned vertices + that
    // many fan entries (+ the list ends of VF) instead of the transposes' scratch
    __host__ static uint32_t smem_bytes(const uint32_t max_n[3], const uint32_t max_not_owned[3],
                                        uint32_t max_stash, bool with_owner, uint32_t stored_rows = 0, uint32_t fan_entries = 0)
    {
        auto           r16 = [](uint32_t x) { return (x + 15u) & ~15u; };
        uint32_t       b   = 0;
        if ((OP == OP_VV || OP == OP_VE || OP == OP_VF) && fan_entries) {
            b = r16(2 * (stored_rows + 1) + 16) + r16(2 * fan_entries + 16) + (OP == OP_VF ? r16(2 * (stored_rows + 1)) : 0u);
            if (with_owner) b += r16(4 * max_not_owned[Tr::dst]) + 16 * max_stash;
            return b;
        }
        if ((OP == OP_FF || OP == OP_EF) && stored_rows) {
            b = r16((OP == OP_FF ? 6 : 4) * stored_rows) + r16(2 * stored_rows);
            if (with_owner) b += r16(4 * max_not_owned[Tr::dst]) + 16 * max_stash;
            return b;
        }
        const uint32_t nr  = Tr::conn == 0 ? max_n[ELEM_E] : max_n[ELEM_F];
        b += r16(2 * W * nr);
        if (op_is_edge4<OP>()) {
            b += r16(8 * max_n[ELEM_E]) + (OP == OP_EVDIAMOND ? r16(4 * max_n[ELEM_E]) : 0u);
        } else if (!op_is_fixed<OP>()) {
            const uint32_t ncols = OP == OP_FF ? max_n[ELEM_E] : max_n[Tr::src];
            b += (PACKED ? r16(2 * (ncols + 1) + 16) : r16(4 * (ncols + 1))) + r16(2 * W * nr);
            if (OP == OP_FF) b += r16(4 * (max_n[ELEM_F] + 1)) + r16(2 * 3 * max_n[ELEM_F] * 2);
        }
        if (with_owner) b += r16(4 * max_not_owned[Tr::dst]) + 16 * max_stash;
        return b;
    }

    // carve shared memory (all threads, identical arithmetic)
    __device__ __forceinline__ void plan(const PatchDesc& d, Smem& sm, bool with_owner, bool all_sources,
                                         bool edge_manifold = false)
    {
        ef3  = false;
        fanq = (OP == OP_VV || OP == OP_VE || OP == OP_VF) && edge_manifold && (d.flags & FLAG_FANS) && !all_sources;
        if (fanq) {
            // sections: fan_off (u16 offsets, bit 15 = closed fan) and fan_v; `edge_manifold` doubles as "stored sections
            // may be used" (k_query_csr passes false: it needs the ascending-id order of the transposes)
            n_rows = 0, n_cols = d.n_owned[ELEM_V];
            // fan_v, fan_e and fan_f are parallel arrays of fan_total entries (same byte size)
            loff_bytes = d.fanoff_bytes(), conn_bytes = d.fanv_bytes();
            s_loff     = sm.alloc<uint16_t>(loff_bytes / 2);
            s_conn     = sm.alloc<uint16_t>(conn_bytes / 2);
            s_off = nullptr, s_val = nullptr, s_off2 = nullptr, s_val2 = nullptr, s_own = nullptr, s_stash = nullptr;
            if (OP == OP_VF) s_val = sm.alloc<uint16_t>(n_cols + 1u);  // list ends
            ff2 = ff3 = false;
            if (with_owner) {
                s_own   = sm.alloc<uint32_t>(d.own_bytes(Tr::dst) / 4);
                s_stash = sm.alloc<StashEntry>(d.n_stash);
            }
            return;
        }
        ff3 = OP == OP_FF && edge_manifold && (d.flags & FLAG_FF) && !all_sources;
        ef3 = OP == OP_EF && edge_manifold && (d.flags & FLAG_FF) && !all_sources;
        ff2 = OP == OP_FF && PACKED && edge_manifold && !ff3;
        if (ff3 || ef3) {
            n_rows = ff3 ? d.n_owned[ELEM_F] : d.n_owned[ELEM_E], n_cols = 0, loff_bytes = 0;
            conn_bytes = ff3 ? d.ff_bytes() : d.ef_bytes();
            s_conn     = sm.alloc<uint16_t>(conn_bytes / 2);
            s_val      = sm.alloc<uint16_t>(n_rows);  // per-row fill count
            s_loff = nullptr, s_off = nullptr, s_off2 = nullptr, s_val2 = nullptr, s_own = nullptr, s_stash = nullptr;
            if (with_owner) {
                s_own   = sm.alloc<uint32_t>(d.own_bytes(Tr::dst) / 4);
                s_stash = sm.alloc<StashEntry>(d.n_stash);
            }
            return;
        }
        const uint32_t rows_all = Tr::conn == 0 ? d.n[ELEM_E] : d.n[ELEM_F];
        n_rows = (op_is_fixed<OP>() && !all_sources) ? d.n_owned[Tr::src] : rows_all;
        if (OP == OP_FF || op_is_edge4<OP>()) n_rows = rows_all;
        conn_bytes = round_up(2u * W * n_rows, 16);
        s_conn     = sm.alloc<uint16_t>(conn_bytes / 2);
        s_loff = nullptr, s_off = nullptr, s_val = nullptr, s_off2 = nullptr, s_val2 = nullptr;
        loff_bytes = 0, n_cols = 0;
        if (op_is_edge4<OP>()) {
            n_cols = d.n[ELEM_E];
            s_val  = sm.alloc<uint16_t>(4u * n_cols);
            if (OP == OP_EVDIAMOND) s_val2 = sm.alloc<uint16_t>(d.ev_bytes() / 2);  // EV section
        } else if (!op_is_fixed<OP>()) {
            n_cols = OP == OP_FF ? d.n[ELEM_E] : (all_sources ? d.n[Tr::src] : d.n_owned[Tr::src]);
            if (ff2) {
                // (face, face) pair per edge + compacted 3-wide rows + per-row counts; no offsets, no scan
                s_off  = sm.alloc<uint32_t>(n_cols);
                const uint32_t nsrc = all_sources ? d.n[ELEM_F] : d.n_owned[ELEM_F];
                s_val2              = sm.alloc<uint16_t>(3u * nsrc);
                s_val               = sm.alloc<uint16_t>(nsrc);
            } else if (PACKED) {
                loff_bytes = round_up(2u * (n_cols + 1), 16);
                s_loff     = sm.alloc<uint16_t>(loff_bytes / 2);
            } else {
                s_off = sm.alloc<uint32_t>(n_cols + 1);
            }
            if (!ff2) s_val = sm.alloc<uint16_t>(W * n_rows);
            if (OP == OP_FF && !ff2) {
                s_off2 = sm.alloc<uint32_t>(d.n[ELEM_F] + 1);
                s_val2 = sm.alloc<uint16_t>(6u * d.n[ELEM_F]);
            }
        }
        s_own = nullptr, s_stash = nullptr;
        if (with_owner) {
            s_own   = sm.alloc<uint32_t>(d.own_bytes(Tr::dst) / 4);
            s_stash = sm.alloc<StashEntry>(d.n_stash);
        }
    }

    // bytes that issue() will put in flight
    __device__ __forceinline__ uint32_t tx_bytes(const PatchDesc& d, bool with_owner) const
    {
        return conn_bytes + loff_bytes + (OP == OP_EVDIAMOND ? d.ev_bytes() : 0u) +
               (with_owner ? d.own_bytes(Tr::dst) + d.stash_bytes() : 0u);
    }

    // thread 0 only, after mbar_arrive_expect_tx
    __device__ __forceinline__ void issue(const PatchDesc& d, const uint8_t* blob, uint64_t* bar, bool with_owner) const
    {
        if ((OP == OP_VV || OP == OP_VE || OP == OP_VF) && fanq) {
            if (conn_bytes)
                bulk_g2s(s_conn, blob + (OP == OP_VV ? d.off_fanv() : (OP == OP_VE ? d.off_fane() : d.off_fanf())), conn_bytes, bar);
            bulk_g2s(s_loff, blob + d.off_fanoff(), loff_bytes, bar);
            if (with_owner) {
                if (d.own_bytes(Tr::dst)) bulk_g2s(s_own, blob + d.off_own(Tr::dst), d.own_bytes(Tr::dst), bar);
                if (d.stash_bytes()) bulk_g2s(s_stash, blob + d.off_stash(), d.stash_bytes(), bar);
            }
            return;
        }
        const uint32_t o = (OP == OP_FF && ff3) ? d.off_ff()
                           : ((OP == OP_EF && ef3) ? d.off_ef() : (Tr::conn == 0 ? d.off_ev() : (Tr::conn == 1 ? d.off_fe() : d.off_fv())));
        if (conn_bytes) bulk_g2s(s_conn, blob + o, conn_bytes, bar);
        if (OP == OP_EVDIAMOND && d.ev_bytes()) bulk_g2s(s_val2, blob + d.off_ev(), d.ev_bytes(), bar);
        if (loff_bytes) {
            const uint32_t lo = (OP == OP_VV || OP == OP_VE) ? d.off_voff_e()
                                                              : (OP == OP_VF ? d.off_voff_f() : d.off_eoff_f());
            bulk_g2s(s_loff, blob + lo, loff_bytes, bar);
        }
        if (with_owner) {
            if (d.own_bytes(Tr::dst)) bulk_g2s(s_own, blob + d.off_own(Tr::dst), d.own_bytes(Tr::dst), bar);
            if (d.stash_bytes()) bulk_g2s(s_stash, blob + d.off_stash(), d.stash_bytes(), bar);
        }
    }

    __device__ __forceinline__ OwnerTable owner_table(const PatchDesc& d) const
    {
        OwnerTable t;
        t.own       = s_own;
        t.stash     = s_stash;
        t.n_owned   = d.n_owned[Tr::dst];
        t.patch     = d.patch_id;
        t.slot_base = d.slot_base[Tr::dst];
        t.type      = Tr::dst;
        return t;
    }

    // all threads, after the mbarrier wait. Ends with a block barrier for CSR ops.
    __device__ __forceinline__ QueryResult compute(const PatchDesc& d, uint32_t* warp_tmp, bool all_sources,
                                                   bool sorted)
    {
        QueryResult    r;
        const uint32_t lim = all_sources ? d.n[Tr::src] : d.n_owned[Tr::src];
        r.n_src = lim, r.shift = 0, r.stride = 0, r.mask = 0xFFFFu;
        r.off16 = nullptr, r.off32 = nullptr, r.cnt = nullptr, r.end16 = nullptr;
        const uint16_t* c = s_conn;
        if ((OP == OP_VV || OP == OP_VE) && fanq) {
            for (uint32_t i = threadIdx.x; i <= lim; i += BT)
                s_loff[i] &= FAN_OFF_MASK;  // drop the closed-fan flag: plain list bounds
            __syncthreads();
            r.off16 = s_loff, r.val = c, r.mask = 0xFFFFu;
            return r;
        }
        if (OP == OP_VF && fanq) {
            // fan_f[i] = face between fan vertices i and i + 1: an open fan has no face after its last vertex
            for (uint32_t v = threadIdx.x; v < lim; v += BT) {
                const uint32_t o = s_loff[v], b = o & FAN_OFF_MASK, e = s_loff[v + 1] & FAN_OFF_MASK;
                s_val[v]         = (uint16_t)((o & FAN_CLOSED) || e == b ? e : e - 1u);
            }
            __syncthreads();
            for (uint32_t i = threadIdx.x; i <= lim; i += BT)
                s_loff[i] &= FAN_OFF_MASK;
            __syncthreads();
            r.off16 = s_loff, r.end16 = s_val, r.val = c, r.mask = 0xFFFFu;
            return r;
        }
        if (OP == OP_FF && ff3) {
            for (uint32_t f = threadIdx.x; f < lim; f += BT)
                s_val[f] = (uint16_t)((c[3 * f] != 0xFFFFu) + (c[3 * f + 1] != 0xFFFFu) + (c[3 * f + 2] != 0xFFFFu));
            __syncthreads();
            r.val = c, r.stride = 3, r.cnt = s_val;
            return r;
        }
        if (OP == OP_EF && ef3) {
            for (uint32_t e = threadIdx.x; e < lim; e += BT)
                s_val[e] = (uint16_t)((c[2 * e] != 0xFFFFu) + (c[2 * e + 1] != 0xFFFFu));
            __syncthreads();
            r.val = c, r.stride = 2, r.cnt = s_val;
            return r;
        }
        if (op_is_edge4<OP>()) {
            const uint32_t em = PACKED ? PK_ID_MASK : 0x7FFFu;
            const uint32_t ne = n_cols;
            if (OP == OP_EVDIAMOND) {
                const uint32_t* ev2 = reinterpret_cast<const uint32_t*>(s_val2);
                uint32_t*       o2  = reinterpret_cast<uint32_t*>(s_val);
                for (uint32_t e = threadIdx.x; e < ne; e += BT) {
                    const uint32_t w = ev2[e];
                    o2[2 * e]     = (w & ID_MASK) | 0xFFFF0000u;
                    o2[2 * e + 1] = ((w >> 16) & ID_MASK) | 0xFFFF0000u;
                }
                __syncthreads();
                for (uint32_t f = threadIdx.x; f < n_rows; f += BT) {
                    uint32_t e[3], dir[3];
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const uint32_t a = c[3 * f + j];
                        e[j] = (a >> 1) & em, dir[j] = a & 1u;
                    }
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const int      j1 = (j + 1) % 3;
                        const uint16_t v  = s_val[4 * e[j] + 2 * dir[j]];  // first vertex of edge j in the face's winding
                        s_val[4 * e[j1] + 1 + 2 * dir[j1]] = v;           // = the vertex opposite to edge j + 1
                    }
                }
            } else {
                for (uint32_t i = threadIdx.x; i < 4u * ne; i += BT)
                    s_val[i] = 0xFFFFu;
                __syncthreads();
                for (uint32_t f = threadIdx.x; f < n_rows; f += BT) {
                    uint32_t e[3], dir[3];
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const uint32_t a = c[3 * f + j];
                        e[j] = (a >> 1) & em, dir[j] = a & 1u;
                    }
#pragma unroll
                    for (int cur = 0; cur < 3; ++cur) {
                        const unsigned short nxt = (unsigned short)e[(cur + 1) % 3], prv = (unsigned short)e[(cur + 2) % 3];
                        unsigned short*      side = reinterpret_cast<unsigned short*>(s_val) + 4 * e[cur] + 2 * dir[cur];
                        if (atomicCAS(side, (unsigned short)0xFFFFu, nxt) != 0xFFFFu) {
                            // both faces traverse the edge the same way (inconsistent orientation): take the other side
                            side = reinterpret_cast<unsigned short*>(s_val) + 4 * e[cur] + 2 * (dir[cur] ^ 1u);
                            atomicCAS(side, (unsigned short)0xFFFFu, nxt);
                        }
                        side[1] = prv;
                    }
                }
            }
            __syncthreads();
            r.val = s_val, r.stride = 4;
            return r;
        }
        if (OP == OP_FF && ff2) {
            // every edge has at most two faces: slot (2 e + rank) of the pair table names them; the face across
            // edge j of f is the other entry of the pair (reference order: edge 0, 1, 2, boundary edges skipped)
            constexpr uint32_t rk = PK_ID_BITS + 1;
            for (uint32_t e = threadIdx.x; e < n_cols; e += BT)
                s_off[e] = 0xFFFFFFFFu;
            __syncthreads();
            uint16_t* pair = reinterpret_cast<uint16_t*>(s_off);
            for (uint32_t f = threadIdx.x; f < n_rows; f += BT) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const uint32_t a = c[3 * f + j];
                    pair[2u * ((a >> 1) & PK_ID_MASK) + ((a >> rk) & 1u)] = (uint16_t)f;
                }
            }
            __syncthreads();
            for (uint32_t f = threadIdx.x; f < lim; f += BT) {
                uint32_t k = 0;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const uint32_t w = s_off[(c[3 * f + j] >> 1) & PK_ID_MASK];
                    const uint32_t o = (w & 0xFFFFu) == f ? (w >> 16) : (w & 0xFFFFu);
                    if (o != 0xFFFFu) s_val2[3 * f + k++] = (uint16_t)o;
                }
                s_val[f] = (uint16_t)k;
            }
            __syncthreads();
            r.val = s_val2, r.stride = 3, r.cnt = s_val;
            return r;
        }
        if (OP == OP_EV) {
            r.val = c, r.stride = 2, r.mask = ID_MASK;
        } else if (OP == OP_FV) {
            r.val = c, r.stride = 3, r.mask = ID_MASK;
        } else if (OP == OP_FE) {
            r.val = c, r.stride = 3, r.shift = 1, r.mask = PACKED ? PK_ID_MASK : 0x7FFFu;
        } else if (OP == OP_VV || OP == OP_VE) {
            if (PACKED) {
                // one thread per edge: both endpoints in one 32-bit read
                const uint32_t* c2 = reinterpret_cast<const uint32_t*>(c);
                for (uint32_t e = threadIdx.x; e < n_rows; e += BT) {
                    const uint32_t w  = c2[e];
                    const uint32_t a = w & 0xFFFFu, b = w >> 16;
                    const uint32_t va = a & PK_ID_MASK, vb = b & PK_ID_MASK;
                    if (va < n_cols) s_val[s_loff[va] + (a >> PK_ID_BITS)] = (uint16_t)(OP == OP_VV ? vb : e);
                    if (vb < n_cols) s_val[s_loff[vb] + (b >> PK_ID_BITS)] = (uint16_t)(OP == OP_VV ? va : e);
                }
                __syncthreads();
            } else {
                transpose_atomic(2u * n_rows, warp_tmp, [c](uint32_t i) { return (uint32_t)c[i]; },
                                 [c](uint32_t i) { return OP == OP_VV ? (uint32_t)c[i ^ 1u] : (i >> 1); });
                if (sorted) csr_sort_lists<BT>(n_cols, s_off, s_val);
            }
            r.off16 = PACKED ? s_loff : nullptr, r.off32 = PACKED ? nullptr : s_off, r.val = s_val;
        } else {  // VF, EF, FF: transpose of a 3-wide face array
            constexpr uint32_t sh = OP == OP_VF ? 0u : 1u;                          // FE: drop dir bit
            constexpr uint32_t rk = OP == OP_VF ? PK_ID_BITS : PK_ID_BITS + 1;       // rank position
            if (PACKED) {
                for (uint32_t f = threadIdx.x; f < n_rows; f += BT) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const uint32_t a  = c[3 * f + j];
                        const uint32_t cc = (a >> sh) & PK_ID_MASK;
                        if (cc < n_cols) s_val[s_loff[cc] + (a >> rk)] = (uint16_t)f;
                    }
                }
                __syncthreads();
            } else {
                transpose_atomic(3u * n_rows, warp_tmp, [c](uint32_t i) { return (uint32_t)c[i] >> sh; },
                                 [](uint32_t i) { return i / 3u; });
                if (sorted && OP != OP_FF) csr_sort_lists<BT>(n_cols, s_off, s_val);
            }
            r.off16 = PACKED ? s_loff : nullptr, r.off32 = PACKED ? nullptr : s_off, r.val = s_val;
            if (OP == OP_FF) {
                // faces across the three edges of every source face, in edge order
                const QueryResult ef = r;
                const uint32_t    nf = lim;
                const uint32_t    em = PACKED ? PK_ID_MASK : 0x7FFFu;
                for (uint32_t f = threadIdx.x; f <= nf; f += BT) {
                    uint32_t k = 0;
                    if (f < nf)
                        for (int j = 0; j < 3; ++j)
                            k += ef.size((c[3 * f + j] >> 1) & em) - 1;
                    s_off2[f] = k;
                }
                block_exclusive_scan<BT>(s_off2, nf, warp_tmp);
                for (uint32_t f = threadIdx.x; f < nf; f += BT) {
                    uint32_t w = s_off2[f];
                    for (int j = 0; j < 3; ++j) {
                        const uint32_t e = (c[3 * f + j] >> 1) & em;
                        for (uint32_t i = ef.begin(e); i < ef.end(e); ++i)
                            if (s_val[i] != f) s_val2[w++] = s_val[i];
                    }
                }
                __syncthreads();
                r.off16 = nullptr, r.off32 = s_off2, r.val = s_val2;
            }
        }
        return r;
    }

   private:
    // wide path: one shared atomic per non-zero, rank kept in a register
    template <typename ColFn, typename ValFn>
    __device__ __forceinline__ void transpose_atomic(uint32_t nnz, uint32_t* warp_tmp, ColFn col, ValFn val)
    {
        const uint32_t tid = threadIdx.x;
        const uint32_t lim = n_cols;
        for (uint32_t i = tid; i <= lim; i += BT)
            s_off[i] = 0;
        __syncthreads();
        uint16_t rank[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            const uint32_t i = tid + k * BT;
            if (i < nnz) {
                const uint32_t cc = col(i);
                if (cc < lim) rank[k] = (uint16_t)atomicAdd(&s_off[cc], 1u);
            }
        }
        block_exclusive_scan<BT>(s_off, lim, warp_tmp);
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            const uint32_t i = tid + k * BT;
            if (i < nnz) {
                const uint32_t cc = col(i);
                if (cc < lim) s_val[s_off[cc] + rank[k]] = (uint16_t)val(i);
            }
        }
        __syncthreads();
    }
};

}  // namespace dev
}  // namespace rxm
