// patch_layout.h -- the B200 patch record (host + device POD).
//
// Replaces the reference's PatchInfo / PatchStash / LPHashTable / Context bundle
// (/root/reference/include/rxmesh/patch_info.h:24-92, patch_stash.h:12-67,
// lp_hashtable.h:46-290, context.h:15-441) for the STATIC path.
//
// Design (see DESIGN.md "Data layout in HBM"):
//  * one 160-byte PatchDesc per patch in a dense array: counts, slot bases and the
//    PRECOMPUTED byte offset of every section, so a block issues its TMA bulk
//    copies without any layout arithmetic;
//  * one contiguous, 16-byte-aligned "topology blob" per patch holding, in this
//    order, EV (2 u16 / edge), FE (3 u16 / face, bit0 = direction, as the
//    reference rxmesh.cpp:941-983), FV (3 u16 / face, derived = FE o EV, stored so
//    that FV/VF consumers read 6 instead of 12 bytes per face), the three
//    not-owned -> owner tables and the neighbour-patch stash.  Every section
//    starts on a 16-byte boundary and is padded to a multiple of 16 bytes so a
//    single cp.async.bulk moves it into shared memory;
//  * RANK-ANNOTATED INCIDENCE (the "packed" format, used whenever every patch has
//    <= 2048 elements per type, vertex valence < 32 and <= 16 faces per edge):
//    the spare high bits of every 16-bit EV / FV / FE entry carry the RANK of
//    that incidence inside the transposed list it belongs to (the k-th edge of
//    its vertex, the k-th face of its vertex, the k-th face of its edge), and
//    three small u16 offset arrays (VE/VV, VF, EF list starts) follow FV.  One
//    array therefore answers a query in both directions: mask the rank bits to
//    read EV/FV/FE, or scatter row ids to offset[col] + rank to obtain
//    VE/VV/VF/EF in shared memory with no atomics, no scan and a deterministic
//    (ascending row id) neighbour order -- the reference builds the transpose
//    with two passes of CAS-emulated 16-bit shared atomics in racy order
//    (kernels/rxmesh_queries.cuh:16-107), the same way the reference already
//    hides the edge direction in bit 0 of FE (rxmesh.cpp:941-983).  Meshes that
//    exceed the limits use the "wide" format (plain ids) and the atomic path;
//  * ONE-RING FANS (present when every owned vertex of every patch has a single,
//    consistently oriented, edge-manifold fan): per owned vertex the cyclic
//    (or, on a mesh boundary, open) sequence of its neighbour vertices, faces
//    lying between consecutive entries.  2 bytes per half-edge -- the same HBM
//    footprint as EV -- but it answers VV (and oriented VV, the reference's
//    orient_edges_around_vertices, kernels/rxmesh_queries.cuh:375-499) by a plain
//    read, and lets vertex normals / Laplacian run one thread per OWNED vertex
//    with register accumulators: no transpose, no atomics, no ribbon-face work.
//    A parallel array fan_f names the face between consecutive fan vertices, so
//    VF is a plain read as well;
//  * local ids are owned-first, each half sorted by global id (the reference's
//    numbering, rxmesh.cpp:845-869), so the owned / active bitmasks of the
//    reference collapse to a prefix [0, n_owned) and the not-owned -> owner
//    lookup is a DIRECT table indexed by (local id - n_owned) instead of a
//    cuckoo hash probe;
//  * attribute storage holds OWNED elements only.  Patch p owns the slot range
//    [slot_base(p), slot_base(p) + round_up4(n_owned(p))); slot_base is a
//    multiple of 4 so that a patch's owned slice of a 1/2/3/4-component fp32
//    attribute is 16-byte aligned and TMA-loadable.  The stash carries the
//    neighbour patches' slot bases, so resolving a ribbon element to the address
//    of its value needs no dependent global load.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RXM_HD __host__ __device__ __forceinline__
#else
#define RXM_HD inline
#endif

namespace rxm {

enum : uint32_t
{
    ELEM_V = 0,
    ELEM_E = 1,
    ELEM_F = 2
};

// Same numeric values as the reference's Op enum (types.h:113-129).
enum : int
{
    OP_V  = 0,
    OP_E  = 1,
    OP_F  = 2,
    OP_VV = 3,
    OP_VE = 4,
    OP_VF = 5,
    OP_FV = 6,
    OP_FE = 7,
    OP_FF = 8,
    OP_EV = 9,
    OP_EE = 10,
    OP_EF = 11,
    OP_EVDIAMOND = 12
};

constexpr uint32_t INVALID32_ = 0xFFFFFFFFu;

// packed (rank-annotated) entry formats
constexpr uint32_t PK_ID_BITS   = 11;      // local id bits of EV / FV entries and of the edge id in FE
constexpr uint32_t PK_ID_MASK   = 0x7FFu;
constexpr uint32_t PK_MAX_ELEMS = 2048;    // per element type and patch
constexpr uint32_t PK_MAX_VRANK = 32;      // ranks of vertex lists: 5 bits (entry >> 11)
constexpr uint32_t PK_MAX_ERANK = 16;      // ranks of edge lists: 4 bits (FE entry >> 12)
constexpr uint16_t FLAG_PACKED  = 1;
constexpr uint16_t FLAG_FANS    = 2;
constexpr uint16_t FLAG_FF      = 4;       // bit 2: stored FF rows of the owned faces (edge-manifold input)
constexpr uint16_t FLAG_RING2   = 8;       // bit 3: full one-rings of the ribbon vertices next to owned vertices (k-ring consumers)
constexpr uint16_t FLAG_UNIFORM6 = 16;     // bit 4: EVERY owned vertex has a closed fan of six: fan_off[v] = 6 v | FAN_CLOSED, so the
                                           //        fan kernels synthesise the offsets instead of loading the section (2 B / vertex)
constexpr uint16_t FAN_CLOSED   = 0x8000;  // bit 15 of a fan_off entry: the fan of this vertex is closed
constexpr uint16_t FAN_OFF_MASK = 0x7FFF;
constexpr uint64_t INVALID64_ = 0xFFFFFFFFFFFFFFFFull;

RXM_HD uint32_t round_up(uint32_t x, uint32_t m)
{
    return (x + m - 1) / m * m;
}

// one entry of the neighbour-patch stash (16 bytes)
struct StashEntry
{
    uint32_t patch;      // neighbour patch id
    uint32_t slot_base[3];  // its attribute slot base for V, E, F
};

// owner record of a not-owned local element: (stash slot << 16) | local id in owner
RXM_HD uint32_t pack_owner(uint32_t stash_slot, uint32_t owner_local)
{
    return (stash_slot << 16) | owner_local;
}

struct alignas(16) PatchDesc
{
    uint64_t topo_off;      // byte offset of this patch's blob in the topology buffer (16B aligned)
    uint32_t topo_bytes;    // total blob bytes (multiple of 16)
    uint32_t patch_id;      // global patch id (differs from the local index on a sharded mesh)
    uint16_t n[3];          // #V, #E, #F in the patch, ribbon included
    uint16_t n_owned[3];    // #owned V, E, F (local ids [0, n_owned) are owned)
    uint16_t n_stash;       // neighbour patches referenced by the owner tables
    uint16_t flags;         // bit 0: packed (rank-annotated) format, bit 1: one-ring fans present
    uint32_t slot_base[3];  // attribute slot base for V, E, F (multiple of 4)
    uint32_t fan_total;     // entries of the fan neighbour array
    uint32_t lin_base[3];   // gap-free linear-id prefix (reference Context::linear_id, context.h:275-290)
    uint32_t pad0;
    // ---- section byte offsets inside the blob (all multiples of 16), precomputed by the builder so a
    //      kernel spends no instructions on layout arithmetic. Order: EV, FE, FV, voff_e, voff_f, eoff_f,
    //      fan_off, fan_v, owner V/E/F, stash ----
    uint32_t o_fe, o_fv, o_voff_e, o_voff_f, o_eoff_f, o_fanoff, o_fanv, o_own[3], o_stash;
    uint32_t o_fanf;        // fan faces: fan_f[i] = local face between fan_v[i] and fan_v[i+1] (0xFFFF: none)
    uint32_t o_ff;          // stored FF: 3 u16 per OWNED face, the faces across edge 0, 1, 2 compacted to the front, 0xFFFF after
    uint32_t o_ef;          // stored EF: 2 u16 per OWNED edge, its (at most two) faces in ascending local id, 0xFFFF after
    uint32_t pad1[2];
    // ---- sections appended in round 2 (after EF, so every older offset keeps its meaning) ----
    uint32_t o_fane;        // fan edges: fan_e[i] = local edge between the fan's vertex and fan_v[i] (VE as a plain read, oriented VE)
    // RING EXTENSION (FLAG_RING2): a k-ring walk that starts at an owned vertex may have to expand a vertex whose one-ring
    // is only partly inside the patch (a ribbon vertex) or that the patch does not hold at all.  For every NOT-OWNED vertex
    // within ring_depth (default 2) rings of an owned one the builder stores the COMPLETE ring as ids in an extended local
    // space: [0, n[V]) = the patch's vertices, n[V] + k = the k-th "ext" vertex (beyond the ribbon), resolved through
    // ext_own like a ribbon vertex.
    uint32_t o_r2idx;       // u16[n[V] - n_owned[V] + n_ext]: ring index of the not-owned / ext vertex, 0xFFFF = ring not stored
    uint32_t o_r2off;       // u16[n_r2 + 1]
    uint32_t o_r2val;       // u16[r2_total] extended local ids
    uint32_t o_ext;         // u32[n_ext] owner records (stash slot << 16 | local id in owner)
    uint16_t n_r2, n_ext;
    uint32_t r2_total;
    uint32_t pad2;

    RXM_HD uint32_t ev_bytes() const { return o_fe; }
    RXM_HD uint32_t fe_bytes() const { return o_fv - o_fe; }
    RXM_HD uint32_t own_bytes(uint32_t t) const { return (t == 2 ? o_stash : o_own[t + 1]) - o_own[t]; }
    RXM_HD uint32_t off_ev() const { return 0; }
    RXM_HD uint32_t off_fe() const { return o_fe; }
    RXM_HD uint32_t off_fv() const { return o_fv; }
    // list-offset arrays (u16, one entry per local column + 1): VE/VV, VF, EF
    RXM_HD uint32_t voff_bytes() const { return o_voff_f - o_voff_e; }
    RXM_HD uint32_t eoff_bytes() const { return o_fanoff - o_eoff_f; }
    RXM_HD uint32_t off_voff_e() const { return o_voff_e; }
    RXM_HD uint32_t off_voff_f() const { return o_voff_f; }
    RXM_HD uint32_t off_eoff_f() const { return o_eoff_f; }
    // one-ring fans of the owned vertices: fan_off[nov+1] (bit 15 = closed fan), fan_v[fan_total]
    RXM_HD uint32_t fanoff_bytes() const { return o_fanv - o_fanoff; }
    RXM_HD uint32_t fanv_bytes() const { return o_fanf - o_fanv; }
    RXM_HD uint32_t fanf_bytes() const { return o_own[0] - o_fanf; }
    RXM_HD uint32_t off_fanf() const { return o_fanf; }
    RXM_HD uint32_t off_fanoff() const { return o_fanoff; }
    RXM_HD uint32_t off_fanv() const { return o_fanv; }
    RXM_HD uint32_t off_own(uint32_t t) const { return o_own[t]; }
    RXM_HD uint32_t off_stash() const { return o_stash; }
    RXM_HD uint32_t stash_bytes() const { return 16u * n_stash; }
    RXM_HD uint32_t off_ff() const { return o_ff; }
    RXM_HD uint32_t ff_bytes() const { return (flags & FLAG_FF) ? round_up(6u * n_owned[ELEM_F], 16) : 0u; }
    RXM_HD uint32_t off_ef() const { return o_ef; }
    RXM_HD uint32_t ef_bytes() const { return (flags & FLAG_FF) ? round_up(4u * n_owned[ELEM_E], 16) : 0u; }
    RXM_HD uint32_t slot_cap(uint32_t t) const { return (n_owned[t] + 3u) & ~3u; }
    RXM_HD uint32_t off_fane() const { return o_fane; }
    RXM_HD uint32_t fane_bytes() const { return o_r2idx - o_fane; }
    RXM_HD uint32_t r2idx_bytes() const { return o_r2off - o_r2idx; }
    RXM_HD uint32_t r2off_bytes() const { return o_r2val - o_r2off; }
    RXM_HD uint32_t r2val_bytes() const { return o_ext - o_r2val; }
    RXM_HD uint32_t ext_bytes() const { return (flags & FLAG_RING2) ? round_up(4u * n_ext, 16) : 0u; }

    // builder: lay the sections out from the counts / flags already stored in this record
    inline void compute_layout()
    {
        uint32_t o = 0;
        o += round_up(4u * n[ELEM_E], 16);
        o_fe = o;
        o += round_up(6u * n[ELEM_F], 16);
        o_fv = o;
        o += round_up(6u * n[ELEM_F], 16);
        o_voff_e = o;
        o += round_up(2u * (n[ELEM_V] + 1u), 16);
        o_voff_f = o;
        o += round_up(2u * (n[ELEM_V] + 1u), 16);
        o_eoff_f = o;
        o += round_up(2u * (n[ELEM_E] + 1u), 16);
        o_fanoff = o;
        o += (flags & 2) ? round_up(2u * (n_owned[ELEM_V] + 1u), 16) : 0u;
        o_fanv = o;
        o += (flags & 2) ? round_up(2u * fan_total, 16) : 0u;
        o_fanf = o;
        o += (flags & 2) ? round_up(2u * fan_total, 16) : 0u;
        for (int t = 0; t < 3; ++t) {
            o_own[t] = o;
            o += round_up(4u * (uint32_t)(n[t] - n_owned[t]), 16);
        }
        o_stash    = o;
        o_ff       = o + 16u * n_stash;
        o_ef       = o_ff + ff_bytes();
        o          = o_ef + ef_bytes();
        o_fane     = o;
        o += (flags & FLAG_FANS) ? round_up(2u * fan_total, 16) : 0u;
        const bool r2 = (flags & FLAG_RING2) != 0;
        o_r2idx = o;
        o += r2 ? round_up(2u * ((uint32_t)(n[ELEM_V] - n_owned[ELEM_V]) + n_ext), 16) : 0u;
        o_r2off = o;
        o += r2 ? round_up(2u * (n_r2 + 1u), 16) : 0u;
        o_r2val = o;
        o += r2 ? round_up(2u * r2_total, 16) : 0u;
        o_ext = o;
        o += r2 ? round_up(4u * n_ext, 16) : 0u;
        topo_bytes = o;
    }
};
static_assert(sizeof(PatchDesc) == 160, "PatchDesc must be 160 bytes");

// By-value kernel argument: the static-path equivalent of the reference Context.
struct MeshView
{
    const PatchDesc* desc;   // [num_patches]
    const uint8_t*   topo;   // topology blob
    uint32_t         num_patches;
    uint32_t         num_slots[3];  // attribute slots per element type (sum of slot caps)
    uint32_t         num_elems[3];  // #V, #E, #F of the (local shard of the) mesh
    const uint32_t*  patch_slot_base[3];  // [num_patches+1] per type: slot base of every patch
    uint32_t         packed;              // 1: every patch uses the rank-annotated format
    uint32_t         fans;                // 1: every patch stores the one-ring fans of its owned vertices
    uint32_t         edge_manifold;       // 1: no edge of the input has more than two incident faces
    uint32_t         ring2;               // 1: every patch stores the ring-2 extension (FLAG_RING2)
};

// Attribute layouts: numeric values of the reference's layoutT (types.h:84-90).
enum : uint32_t
{
    LAYOUT_AOS   = 0,  // data[(slot) * nattr + a]
    LAYOUT_AOSOA = 1,  // data[slot_base(p) * nattr + a * cap(p) + lid]  (patch-local SoA; reference default)
    LAYOUT_SOA   = 2   // data[a * num_elems + linear_id]: the reference's tensor layout (attribute.h:406-421), a gap-free
                       // column-major #elements x nattr matrix indexed by linear id -- no padding slots
};

// Device/host view of an attribute (shallow, by value into kernels like the
// reference's Attribute copy, attribute.h:194).
template <typename T>
struct AttrView
{
    T*              data;
    const uint32_t* slot_base;  // [num_patches+1] for this element type
    uint32_t        num_slots;
    uint32_t        nattr;
    uint32_t        layout;
    const uint32_t* lin_base;   // [num_patches+1] linear-id prefix of this element type (SoA only)
    uint32_t        num_elems;  // elements of this type (SoA only)

    RXM_HD uint64_t index(uint32_t patch, uint32_t lid, uint32_t a) const
    {
        if (layout == LAYOUT_SOA) return (uint64_t)a * num_elems + lin_base[patch] + lid;
        const uint32_t b = slot_base[patch];
        if (layout == LAYOUT_AOS) return (uint64_t)(b + lid) * nattr + a;
        const uint32_t cap = slot_base[patch + 1] - b;
        return (uint64_t)b * nattr + (uint64_t)a * cap + lid;
    }
    // same, when the caller already knows the patch's slot base, capacity and linear-id base (no global load).
    // SoA has storage for the n_owned elements of a patch only: callers never pass a padding slot (lid >= n_owned).
    RXM_HD uint64_t index_known(uint32_t b, uint32_t cap, uint32_t lb, uint32_t lid, uint32_t a) const
    {
        if (layout == LAYOUT_AOS) return (uint64_t)(b + lid) * nattr + a;
        if (layout == LAYOUT_SOA) return (uint64_t)a * num_elems + lb + lid;
        return (uint64_t)b * nattr + (uint64_t)a * cap + lid;
    }
};

}  // namespace rxm
