"""Multi-GPU mode: patches sharded across ranks + ribbon (halo) exchange.

New relative to the reference (single device).  One process per GPU; every rank
builds only ITS part of the mesh:

  * patches are split into contiguous id ranges, one per rank ("real" patches);
  * a rank's local mesh = the faces of its real patches plus two vertex-rings of foreign faces,
    which makes (i) every ribbon element of a real patch present locally and (ii) the ownership
    of those ribbon elements (lowest incident patch id, patcher/patcher.cu:730-756) identical to
    the global mesh.  The foreign faces form "ghost" patches: built like any patch, never
    computed on (rxm_mesh_set_active_patches), their attribute slots are the halo mirrors;
  * before a kernel that reads neighbours of freshly written data, `HaloExchange.exchange(attr)`
    fills the ghost slots a rank's real patches reference with the owners' values, matched by
    GLOBAL element id.  Transport: packed buffers + NCCL send/recv (torch.distributed), or one
    direct NVLink P2P store kernel per neighbour into the peer's attribute storage
    (rxm_attr_push_slots over cudaIpc-mapped pointers).  With the gloo backend (CPU tests) the
    same plan runs on HOST attributes -- that path exists to test the plan, not to compute.
"""
import ctypes as C

import numpy as np

from . import meshio
from ._lib import check, lib, u32p
from .mesh import AoS, DEVICE, HOST, RXMeshStatic, SoA, _stream_ptr


def patch_ranges(face_patch, world):
    """Contiguous patch-id ranges per rank, balanced by face count. Returns bounds[world+1]."""
    cnt = np.bincount(np.asarray(face_patch, dtype=np.int64))
    cum = np.concatenate([[0], np.cumsum(cnt)])
    targets = cum[-1] * np.arange(1, world) / world
    inner = np.searchsorted(cum, targets, side="left")
    return np.concatenate([[0], inner, [cnt.shape[0]]]).astype(np.int64)


def shard_faces(fv, face_patch, rank, world, rings=2):
    """Generic (any mesh): the faces a rank needs = its real faces + `rings` vertex-rings.
    Returns dict(fv=local faces with compact vertex ids, face_patch=GLOBAL patch ids of those faces,
    l2g_v, l2g_f = sorted global ids, bounds = patch_ranges)."""
    fv = np.ascontiguousarray(fv, dtype=np.uint32).reshape(-1, 3)
    fp = np.ascontiguousarray(face_patch, dtype=np.uint32)
    bounds = patch_ranges(fp, world)
    sel = (fp >= bounds[rank]) & (fp < bounds[rank + 1])
    nv = int(fv.max()) + 1
    for _ in range(rings):
        mark = np.zeros(nv, dtype=bool)
        mark[fv[sel].reshape(-1)] = True
        sel = mark[fv].any(axis=1)
    l2g_f = np.nonzero(sel)[0].astype(np.uint32)
    l2g_v = np.unique(fv[l2g_f])
    return dict(fv=np.searchsorted(l2g_v, fv[l2g_f]).astype(np.uint32), face_patch=fp[l2g_f],
                l2g_v=l2g_v.astype(np.uint64), l2g_f=l2g_f.astype(np.uint64), bounds=bounds)


def grid_slab(nx, ny_global, tile, tile_i, rank, world, dx=1.0):
    """Analytic shard of meshio.grid(nx, ny_global): rank r owns a contiguous range of tile rows and
    generates only its slab plus one ghost tile row on each side (>= 2 vertex-rings since tile_i >= 2).
    Global vertex id = local id + row0 * nx; global face id = local id + 2*(nx-1)*row0."""
    n_tile_rows = (ny_global - 1 + tile_i - 1) // tile_i
    ntj = (nx - 1 + tile - 1) // tile
    tb = np.round(np.linspace(0, n_tile_rows, world + 1)).astype(np.int64)  # tile-row bounds per rank
    t0, t1 = int(tb[rank]), int(tb[rank + 1])
    h0, h1 = max(t0 - 1, 0), min(t1 + 1, n_tile_rows)  # with ghost tile rows
    row0, row1 = h0 * tile_i, min(h1 * tile_i, ny_global - 1)  # quad rows [row0, row1)
    V, F = meshio.grid(nx, row1 - row0 + 1, dx)
    V[:, 2] += np.float32(dx * row0)
    xg, zg = V[:, 0].astype(np.float64), V[:, 2].astype(np.float64)
    V[:, 1] = (0.05 * np.sin(0.01 * xg) * np.cos(0.013 * zg)).astype(np.float32)  # same field as the global grid
    fp_local = meshio.grid_face_tiles(nx, row1 - row0 + 1, tile, tile_i)  # local tile ids, row-major
    fp_global = fp_local + np.uint32(h0 * ntj)
    nv_loc = V.shape[0]
    return dict(fv=F, verts=V, face_patch=fp_global,
                l2g_v=np.arange(nv_loc, dtype=np.uint64) + np.uint64(row0 * nx),
                l2g_f=np.arange(F.shape[0], dtype=np.uint64) + np.uint64(2 * (nx - 1) * row0),
                bounds=tb * ntj)


class ShardedMesh:
    """A rank's part of a patch-sharded mesh."""

    def __init__(self, shard, rank, world, patch_size=512, device=True, num_threads=0, ring2=None):
        self.rank, self.world = rank, world
        self.l2g = {0: np.asarray(shard["l2g_v"], dtype=np.uint64), 2: np.asarray(shard["l2g_f"], dtype=np.uint64)}
        self.bounds = np.asarray(shard["bounds"], dtype=np.int64)
        self.verts = shard.get("verts")
        gp = np.asarray(shard["face_patch"], dtype=np.uint32)
        self.patch_global = np.unique(gp)  # local patch q <-> global patch patch_global[q]
        self.mesh = RXMeshStatic(shard["fv"], face_patch=gp, patch_size=patch_size, device=device,
                                 num_threads=num_threads, ring2=ring2)
        g0, g1 = self.bounds[rank], self.bounds[rank + 1]
        self.first = int(np.searchsorted(self.patch_global, g0))
        self.count = int(np.searchsorted(self.patch_global, g1)) - self.first
        check(lib().rxm_mesh_set_active_patches(self.mesh._h, self.first, self.count))
        # elements owned by real patches (what this rank is responsible for)
        self._real_owned = {}

    def real_owned_mask(self, elem):
        """bool per LOCAL element: owned by one of this rank's real patches."""
        if elem not in self._real_owned:
            ep = self.mesh.elem_patch(elem)
            self._real_owned[elem] = (ep >= self.first) & (ep < self.first + self.count)
        return self._real_owned[elem]

    def halo_slots(self, elem):
        out, n = u32p(), C.c_uint64()
        check(lib().rxm_mesh_halo_slots(self.mesh._h, int(elem), self.first, self.count, C.byref(out), C.byref(n)))
        arr = np.ctypeslib.as_array(out, shape=(max(n.value, 1),))[:n.value].copy()
        lib().rxm_free(out)
        return arr


class HaloExchange:
    """Halo plan for one element type (0 = vertices, 2 = faces) of a ShardedMesh."""

    def __init__(self, sm, elem, group=None):
        import torch.distributed as dist
        self.sm, self.elem, self.dist, self.group = sm, int(elem), dist, group
        m = sm.mesh
        need_slots = sm.halo_slots(elem)                       # my ghost slots to fill
        loc = m.slot_to_global(elem)[need_slots]               # local element ids
        gids = sm.l2g[elem][loc]                                # global element ids
        owner_gp = sm.patch_global[m.elem_patch(elem)[loc]]     # global owner patch -> rank
        owner_rank = np.searchsorted(sm.bounds, owner_gp, side="right") - 1
        assert not np.any(owner_rank == sm.rank), "a needed halo element is owned by this rank's ghost copy only"
        req = {int(r): (gids[owner_rank == r], need_slots[owner_rank == r]) for r in np.unique(owner_rank)}
        # tell every owner which global ids I need (setup-time object collective)
        all_req = [None] * sm.world
        dist.all_gather_object(all_req, {r: v[0] for r, v in req.items()}, group=group)
        self.recv = {r: v[1].astype(np.int32) for r, v in req.items()}   # peer -> my slots (in request order)
        self.send = {}
        g2s, real = m.global_to_slot(elem), sm.real_owned_mask(elem)
        for peer, wants in enumerate(all_req):
            ids = wants.get(sm.rank) if wants else None
            if ids is None or len(ids) == 0:
                continue
            li = np.searchsorted(sm.l2g[elem], ids)
            assert np.array_equal(sm.l2g[elem][li], ids), "peer asked for an element this rank does not hold"
            assert real[li].all(), "peer asked for an element this rank does not own"
            self.send[peer] = g2s[li].astype(np.int32)
        self.bytes_per_exchange_row = None
        self._dev = None

    def halo_elements(self):
        return int(sum(len(v) for v in self.recv.values()))

    # -------------------------------------------------------------- device path (NCCL send/recv)
    def _device_state(self, attr):
        import torch
        if self._dev is None:
            dev = torch.device("cuda", torch.cuda.current_device())
            self._dev = dict(
                send_idx={p: torch.from_numpy(v).to(dev) for p, v in self.send.items()},
                recv_idx={p: torch.from_numpy(v).to(dev) for p, v in self.recv.items()}, bufs={})
        key = (attr.num_attributes, attr.dtype.itemsize)
        if key not in self._dev["bufs"]:
            import torch
            dev = torch.device("cuda", torch.cuda.current_device())
            row_bytes = attr.num_attributes * attr.dtype.itemsize
            assert row_bytes % 4 == 0, "halo rows travel as 32-bit words: the row size must be a multiple of 4 bytes"
            words = row_bytes // 4
            self._dev["bufs"][key] = (
                {p: torch.empty((len(v), words), dtype=torch.int32, device=dev) for p, v in self.send.items()},
                {p: torch.empty((len(v), words), dtype=torch.int32, device=dev) for p, v in self.recv.items()})
        return self._dev, self._dev["bufs"][key]

    def exchange(self, attr, stream=None):
        """Fill the ghost slots of `attr` (AoS) with the owners' current values."""
        assert attr.elem == self.elem and (attr.layout == AoS or (attr.num_attributes == 1 and attr.layout != SoA)), \
            "halo rows are addressed by slot: an AoS (or single-component AoSoA) attribute of the plan's element type"
        if attr.location & DEVICE and self.dist.get_backend(self.group) == "nccl":
            return self._exchange_device(attr, stream)
        return self._exchange_host(attr)

    def _exchange_device(self, attr, stream):
        import torch
        st, (sbuf, rbuf) = self._device_state(attr)
        stream = stream if stream is not None else torch.cuda.current_stream()
        sp = _stream_ptr(stream)
        # gather, NCCL send/recv and scatter are all ordered on `stream`: torch's collectives order against the CURRENT
        # stream, so it is made current for the duration of the exchange
        with torch.cuda.stream(stream):
            for p, idx in st["send_idx"].items():
                check(lib().rxm_attr_gather_slots(attr._h, C.c_void_p(idx.data_ptr()), idx.numel(),
                                                  C.c_void_p(sbuf[p].data_ptr()), sp))
            ops = []
            for p in sorted(set(sbuf) | set(rbuf)):
                if p in sbuf:
                    ops.append(self.dist.P2POp(self.dist.isend, sbuf[p], p, group=self.group))
                if p in rbuf:
                    ops.append(self.dist.P2POp(self.dist.irecv, rbuf[p], p, group=self.group))
            if ops:
                for w in self.dist.batch_isend_irecv(ops):
                    w.wait()
            for p, idx in st["recv_idx"].items():
                check(lib().rxm_attr_scatter_slots(attr._h, C.c_void_p(idx.data_ptr()), idx.numel(),
                                                   C.c_void_p(rbuf[p].data_ptr()), sp))

    def _exchange_host(self, attr):
        import torch
        na = attr.num_attributes
        h = attr.host_array().reshape(-1, na)
        sbuf = {p: torch.from_numpy(np.ascontiguousarray(h[idx])) for p, idx in self.send.items()}
        rbuf = {p: torch.empty((len(idx), na), dtype=sbuf[p].dtype if p in sbuf else
                               torch.from_numpy(h[:1]).dtype) for p, idx in self.recv.items()}
        ops = []
        for p in sorted(set(sbuf) | set(rbuf)):
            if p in sbuf:
                ops.append(self.dist.P2POp(self.dist.isend, sbuf[p], p, group=self.group))
            if p in rbuf:
                ops.append(self.dist.P2POp(self.dist.irecv, rbuf[p], p, group=self.group))
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()
        for p, idx in self.recv.items():
            h[idx] = rbuf[p].numpy()

    # -------------------------------------------------------------- direct NVLink P2P stores
    def bind_p2p(self, attr):
        """Map every neighbour's storage of the SAME attribute (created in the same order on all ranks)
        through cudaIpc, and learn the neighbour-side slot of every row this rank sends."""
        import torch
        dist = self.dist
        handle = (C.c_uint8 * 64)()
        check(lib().rxm_ipc_export(C.c_void_p(attr.data_ptr(DEVICE)), handle))
        infos = [None] * self.sm.world
        dist.all_gather_object(infos, dict(handle=bytes(handle), recv={p: v for p, v in self.recv.items()}),
                               group=self.group)
        dev = torch.device("cuda", torch.cuda.current_device())
        peers = {}
        for p, idx in self.send.items():
            ptr = C.c_void_p()
            hb = (C.c_uint8 * 64).from_buffer_copy(infos[p]["handle"])
            check(lib().rxm_ipc_open(hb, C.byref(ptr)))
            remote_slots = infos[p]["recv"][self.sm.rank]  # peer's ghost slots, in the order I send
            assert len(remote_slots) == len(idx)
            peers[p] = (ptr, torch.from_numpy(idx).to(dev), torch.from_numpy(remote_slots.astype(np.int32)).to(dev))
        return peers

    def exchange_p2p(self, attr, peers, stream=None, barrier=True):
        """One store kernel per neighbour writes my owned boundary rows straight into the neighbour's
        ghost slots over NVLink; the barrier makes the pushes visible before anyone reads."""
        import torch
        stream = stream if stream is not None else torch.cuda.current_stream()
        sp = _stream_ptr(stream)
        for p, (ptr, lidx, ridx) in peers.items():
            check(lib().rxm_attr_push_slots(attr._h, C.c_void_p(lidx.data_ptr()), ptr, C.c_void_p(ridx.data_ptr()),
                                            lidx.numel(), sp))
        if barrier:
            stream.synchronize()  # the stream the pushes were issued on
            self.dist.barrier(group=self.group)


class FusedHalo:
    """Laplacian smoothing with the ribbon exchange fused into the compute kernel (rxm_laplacian_smooth_fused).

    Every step is ONE kernel per rank: it computes the new positions, stores the rows that are mirrored on neighbouring
    ranks straight into their ghost slots over NVLink (P2P stores through cudaIpc-mapped pointers), and its last block
    raises a flag word on every neighbour; blocks whose patch reads ghost slots wait on the flags the neighbours raised
    at the end of their previous step.  No NCCL call and no host synchronisation inside the iteration.  The two vertex
    attributes the iteration ping-pongs between must have been created in the same order on every rank.
    `plan(hx)` (host only, no GPU needed) is what the gloo tests check."""

    @staticmethod
    def plan(hx):
        """push lists from a vertex HaloExchange: per local patch the (local vertex id, neighbour index) pairs this rank
        sends, in the neighbours' request order; -> (peers, push_off[P+1], patch, lid, peer_index, position in request)"""
        m = hx.sm.mesh
        sb = m.slot_base(0).astype(np.int64)
        peers = sorted(hx.send)
        pat, lid, pidx, pos = [], [], [], []
        for k, p in enumerate(peers):
            slots = hx.send[p].astype(np.int64)
            pp = np.searchsorted(sb, slots, side="right") - 1
            pat.append(pp), lid.append(slots - sb[pp]), pidx.append(np.full(len(slots), k)), pos.append(np.arange(len(slots)))
        pat, lid, pidx, pos = (np.concatenate(a) if a else np.zeros(0, np.int64) for a in (pat, lid, pidx, pos))
        order = np.argsort(pat, kind="stable")
        pat, lid, pidx, pos = pat[order], lid[order], pidx[order], pos[order]
        off = np.zeros(m.get_num_patches() + 1, dtype=np.uint32)
        np.add.at(off, pat + 1, 1)
        return peers, np.cumsum(off, dtype=np.uint32), pat, lid, pidx, pos

    def __init__(self, hx, attr_a, attr_b):
        import torch  # noqa: F401
        assert hx.elem == 0
        self.hx, self.mesh, self.a, self.b, self.step = hx, hx.sm.mesh, attr_a, attr_b, 0
        dist, sm = hx.dist, hx.sm
        peers, off, pat, lid, pidx, pos = self.plan(hx)
        assert set(peers) == set(hx.recv), "halo neighbourhoods must be symmetric"
        self.peers = peers
        h = C.c_void_p()
        check(lib().rxm_fused_halo_create(self.mesh._h, len(peers), C.byref(h)))
        self._h = h

        def export(ptr):
            buf = (C.c_uint8 * 64)()
            check(lib().rxm_ipc_export(C.c_void_p(ptr), buf))
            return bytes(buf)

        infos = [None] * sm.world
        dist.all_gather_object(infos, dict(a=export(attr_a.data_ptr(DEVICE)), b=export(attr_b.data_ptr(DEVICE)),
                                           flags=export(lib().rxm_fused_halo_flags(h)), peers=peers,
                                           recv={p: v for p, v in hx.recv.items()}), group=hx.group)

        def open_(raw):
            ptr = C.c_void_p()
            check(lib().rxm_ipc_open((C.c_uint8 * 64).from_buffer_copy(raw), C.byref(ptr)))
            return ptr.value

        pa = (C.c_void_p * len(peers))(*[open_(infos[p]["a"]) for p in peers])
        pb = (C.c_void_p * len(peers))(*[open_(infos[p]["b"]) for p in peers])
        # my flag word on neighbour p = its flags[index of me among ITS neighbours]
        pf = (C.c_void_p * len(peers))(*[open_(infos[p]["flags"]) + 4 * infos[p]["peers"].index(sm.rank) for p in peers])
        # slot on the neighbour of every row I push = its ghost slots in the order it requested them
        slot = np.empty(len(pat), dtype=np.uint32)
        for k, p in enumerate(peers):
            sel = pidx == k
            slot[sel] = np.asarray(infos[p]["recv"][sm.rank], dtype=np.uint32)[pos[sel]]
        lp = (lid.astype(np.uint32) | (pidx.astype(np.uint32) << np.uint32(16))).astype(np.uint32)
        check(lib().rxm_fused_halo_set(h, off.ctypes.data_as(C.c_void_p), lp.ctypes.data_as(C.c_void_p),
                                       slot.ctypes.data_as(C.c_void_p), len(lp), pa, pb, pf))
        self._keep = (pa, pb, pf)
        dist.barrier(group=hx.group)

    def buffers(self):
        """(attribute the next fused step reads, attribute it writes)"""
        return (self.a, self.b) if self.step % 2 == 0 else (self.b, self.a)

    def smooth(self, lr, iters, stream=None):
        """iters fused steps starting from attribute A if an even number of steps has been run, else B; returns the
        attribute that holds the result.  Ghost slots of the starting attribute must be current (HaloExchange.exchange
        before the first call)."""
        src, dst = (self.a, self.b) if self.step % 2 == 0 else (self.b, self.a)
        for _ in range(iters):
            check(lib().rxm_laplacian_smooth_fused(self.mesh._h, src._h, dst._h, float(lr), self._h,
                                                   int(dst is self.b), self.step, _stream_ptr(stream)))
            self.step += 1
            src, dst = dst, src
        return src

    def __del__(self):
        if getattr(self, "_h", None):
            lib().rxm_fused_halo_destroy(self._h)
            self._h = None
