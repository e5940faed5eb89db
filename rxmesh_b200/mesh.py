"""Host-side mirror of the reference's RXMeshStatic / Attribute interface for the hot path.

Same names, argument meaning and error behaviour as
/root/reference/include/rxmesh/rxmesh_static.h:37-1245 and attribute.h:56-731, over the C ABI
(include/rxmesh_b200.h).  Python is plumbing here: all compute happens in librxmesh_b200.so's
sm_100a kernels; nothing in this module computes mesh results on the CPU.
"""
import ctypes as C
import enum

import numpy as np

from . import meshio
from ._lib import PatcherFile, PatchView, RXMeshError, check, lib

HOST, DEVICE, LOCATION_ALL = 0x01, 0x02, 0x0F  # types.h:51-58
AoS, AoSoA, SoA = 0, 1, 2  # types.h:84-90
INVALID64 = 0xFFFFFFFFFFFFFFFF


class Op(enum.IntEnum):  # types.h:113-129
    V = 0
    E = 1
    F = 2
    VV = 3
    VE = 4
    VF = 5
    FV = 6
    FE = 7
    FF = 8
    EV = 9
    EE = 10
    EF = 11
    EVDiamond = 12


_SRC = {Op.VV: 0, Op.VE: 0, Op.VF: 0, Op.EV: 1, Op.EF: 1, Op.FV: 2, Op.FE: 2, Op.FF: 2, Op.EE: 1, Op.EVDiamond: 1}
_DST = {Op.VV: 0, Op.VE: 1, Op.VF: 2, Op.EV: 0, Op.EF: 2, Op.FV: 0, Op.FE: 1, Op.FF: 2, Op.EE: 1, Op.EVDiamond: 0}


def _stream_ptr(stream):
    if stream is None:
        return None
    if isinstance(stream, int):
        return C.c_void_p(stream)
    return C.c_void_p(stream.cuda_stream)  # torch.cuda.Stream


def rx_init(device=0):
    """rx_init (rxmesh.h:23-30)."""
    check(lib().rxm_init(int(device)))


_PF_FIELDS = ["face_patch", "vertex_patch", "edge_patch", "patches_val", "patches_offset", "ribbon_ext_val",
              "ribbon_ext_offset"]
_PF_HEADER = ["patch_size", "num_patches", "num_vertices", "num_edges", "num_faces", "num_seeds",
              "max_num_patches", "num_components", "num_lloyd_run"]


def load_patcher_file(path):
    """Read a patching saved by the reference (Patcher::save, patcher/patcher.h:154-182) or by
    RXMeshStatic.save_patcher_file: dict of header scalars + uint32 arrays."""
    pf = PatcherFile()
    check(lib().rxm_patcher_file_read(str(path).encode(), C.byref(pf)))
    out = {k: int(pf.header[i]) for i, k in enumerate(_PF_HEADER)}
    for i, k in enumerate(_PF_FIELDS):
        n = int(pf.len[i])
        out[k] = np.ctypeslib.as_array(pf.vec[i], shape=(max(n, 1),))[:n].copy()
    out["patching_time_ms"] = float(pf.patching_time_ms)
    lib().rxm_patcher_file_free(C.byref(pf))
    return out


class Attribute:
    """Attribute<T, HandleT> (attribute.h:56-731): storage for OWNED elements; AoS / AoSoA in slot order, SoA as the
    reference's tensor layout (column-major #elements x #attributes over linear ids, attribute.h:249-261,406-421)."""

    def __init__(self, mesh, elem, dtype, num_attributes=1, location=LOCATION_ALL, layout=AoSoA,
                 name=""):
        self.mesh, self.elem, self.name = mesh, int(elem), name
        self.dtype = np.dtype(dtype)
        self.num_attributes, self.layout = int(num_attributes), int(layout)
        self.location = int(location)
        h = C.c_void_p()
        check(lib().rxm_attr_create(mesh._h, self.elem, self.dtype.itemsize, self.num_attributes,
                                    self.location, self.layout, C.byref(h)))
        self._h = h
        self._keep = mesh  # the mesh must outlive its attributes

    def release(self, location=None):
        """release() frees the attribute; release(HOST | DEVICE) only that side (attribute.cu:375-390)."""
        if not getattr(self, "_h", None):
            return
        if location is None or (int(location) & 0x03) == 0x03:
            lib().rxm_attr_destroy(self._h)
            self._h = None
        else:
            check(lib().rxm_attr_release(self._h, int(location)))
            self.location &= ~int(location)

    def __del__(self):
        self.release()

    def get_num_attributes(self):
        return self.num_attributes

    def count(self):
        return int(lib().rxm_attr_count(self._h))

    def data_ptr(self, location=DEVICE):
        return lib().rxm_attr_data(self._h, int(location))

    def host_array(self):
        """numpy view of the HOST copy in storage order (flat; see index())."""
        p = lib().rxm_attr_data(self._h, HOST)
        if not p:
            raise RXMeshError("attribute has no HOST allocation")
        buf = (C.c_uint8 * (self.count() * self.dtype.itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=self.dtype)

    def reset(self, value, location=DEVICE, stream=None):
        v = np.asarray(value, dtype=self.dtype).reshape(1)
        check(lib().rxm_attr_reset(self._h, v.ctypes.data_as(C.c_void_p), int(location),
                                   _stream_ptr(stream)))

    def move(self, source, target, stream=None):
        check(lib().rxm_attr_move(self._h, int(source), int(target), _stream_ptr(stream)))

    def copy_from(self, source_attr, source=DEVICE, target=DEVICE, stream=None):
        check(lib().rxm_attr_copy_from(self._h, source_attr._h, int(source), int(target),
                                       _stream_ptr(stream)))

    # ---- global-order views (the apps' for_each_vertex(HOST, map_to_global) loops) ----
    def from_global(self, arr, stream=None):
        n = self.mesh._num(self.elem)
        a = np.ascontiguousarray(arr, dtype=self.dtype).reshape(n, self.num_attributes)
        check(lib().rxm_attr_upload_global(self._h, a.ctypes.data_as(C.c_void_p), _stream_ptr(stream)))

    def to_global(self, stream=None):
        n = self.mesh._num(self.elem)
        out = np.empty((n, self.num_attributes), dtype=self.dtype)
        check(lib().rxm_attr_download_global(self._h, out.ctypes.data_as(C.c_void_p),
                                             _stream_ptr(stream)))
        return out

    # ---- ReduceHandle (reduce_handle.h:64-166) ----
    def _reduce(self, kind, other=None, attribute_id=None, stream=None):
        v, h = C.c_double(), C.c_uint64()
        check(lib().rxm_attr_reduce(self._h, other._h if other is not None else None, kind,
                                    0xFFFFFFFF if attribute_id is None else int(attribute_id), C.byref(v), C.byref(h),
                                    _stream_ptr(stream)))
        return v.value, h.value

    def dot(self, other, attribute_id=None, stream=None):
        return self._reduce(0, other, attribute_id, stream)[0]

    def norm2(self, attribute_id=None, stream=None):
        return self._reduce(1, None, attribute_id, stream)[0] ** 0.5

    def reduce(self, op="sum", attribute_id=None, stream=None):
        return self._reduce({"sum": 2, "min": 3, "max": 4}[op], None, attribute_id, stream)[0]

    def arg_max(self, attribute_id=0, stream=None):
        v, h = self._reduce(6, None, attribute_id, stream)
        return h, v

    def arg_min(self, attribute_id=0, stream=None):
        v, h = self._reduce(5, None, attribute_id, stream)
        return h, v

    def index(self, patch, lid, attr=0):
        """flat storage index of (handle, attr): Attribute::operator() (attribute.h:406-434)."""
        sb = self.mesh.slot_base(self.elem)
        b, cap = int(sb[patch]), int(sb[patch + 1] - sb[patch])
        if self.layout == AoS:
            return (b + lid) * self.num_attributes + attr
        if self.layout == SoA:  # the reference's tensor layout: column-major over linear ids, no padding slots
            return attr * self.mesh._num(self.elem) + int(self.mesh.lin_base(self.elem)[patch]) + lid
        return b * self.num_attributes + attr * cap + lid


class RXMeshStatic:
    """RXMeshStatic (rxmesh_static.h:37): build from an OBJ path or a face array.

    patcher_file's role (a saved patching) is played by `face_patch`, an explicit face->patch array.
    """

    def __init__(self, faces_or_path, face_patch=None, patch_size=512, num_threads=0, device=True,
                 verts=None, patcher_file=None, ring2=None):
        if isinstance(faces_or_path, str):
            verts, faces = meshio.import_obj(faces_or_path)
        else:
            faces = faces_or_path
        self._fv = np.ascontiguousarray(faces, dtype=np.uint32).reshape(-1, 3)
        self._verts = None if verts is None else np.ascontiguousarray(verts, dtype=np.float32)
        if patcher_file is not None:  # RXMeshStatic(fv, patcher_file) (rxmesh_static.h:61-66)
            pf = load_patcher_file(patcher_file)
            if pf["num_faces"] != self._fv.shape[0]:
                raise RXMeshError("patcher_file was saved for a mesh with a different number of faces")
            face_patch, patch_size = pf["face_patch"], pf["patch_size"]
        fp = None
        if face_patch is not None:
            fp = np.ascontiguousarray(face_patch, dtype=np.uint32)
            if fp.shape[0] != self._fv.shape[0]:
                raise RXMeshError("face_patch must have one entry per face")
        h = C.c_void_p()
        # ring2=False: leave out the ring-2 extension (RXM_BUILD_NO_RING2) -- meshes that never run a k-ring consumer
        import os
        flags = 1 if (ring2 is False or (ring2 is None and os.environ.get("RXM_NO_RING2"))) else 0
        check(lib().rxm_mesh_create_ex(self._fv.ctypes.data_as(C.c_void_p), self._fv.shape[0],
                                       None if fp is None else fp.ctypes.data_as(C.c_void_p),
                                       int(patch_size), int(num_threads), flags, C.byref(h)))
        self._h = h
        self._attrs = {}
        if device:
            check(lib().rxm_mesh_to_device(self._h))

    def __del__(self):
        if getattr(self, "_h", None):
            for a in list(getattr(self, "_attrs", {}).values()):
                a.release()
            lib().rxm_mesh_destroy(self._h)
            self._h = None

    # ---- getters (rxmesh.h:44-399) ----
    def _info(self, k):
        return int(lib().rxm_mesh_info(self._h, k))

    def _num(self, elem):
        return self._info(elem)

    def get_num_vertices(self):
        return self._info(0)

    def get_num_edges(self):
        return self._info(1)

    def get_num_faces(self):
        return self._info(2)

    def get_num_patches(self):
        return self._info(3)

    def get_patch_size(self):
        return self._info(4)

    def get_input_max_valence(self):
        return self._info(5)

    def get_input_max_edge_incident_faces(self):
        return self._info(6)

    def get_input_max_face_adjacent_faces(self):
        return self._info(7)

    def is_closed(self):
        return bool(self._info(8))

    def is_edge_manifold(self):
        return bool(self._info(9))

    def get_per_patch_max_vertices(self):
        return self._info(10)

    def get_per_patch_max_edges(self):
        return self._info(11)

    def get_per_patch_max_faces(self):
        return self._info(12)

    def num_slots(self, elem):
        return self._info(13 + int(elem))

    def topo_bytes(self):
        return self._info(16)

    def total_local(self, elem):
        return self._info(17 + int(elem))

    def save_patcher_file(self, path):
        """RXMesh::save (rxmesh.h:326-329): write the patching in the reference's Patcher archive format."""
        check(lib().rxm_mesh_save_patcher_file(self._h, str(path).encode()))

    def compact(self):
        """rxm_mesh_compact: free host-side helper arrays of a large mesh that already lives on the device."""
        check(lib().rxm_mesh_compact(self._h))

    def is_packed(self):
        """True when the patch store uses the rank-annotated (atomic-free) format."""
        return bool(self._info(22))

    def has_ring2(self):
        """True when the patches store the ring-2 extension (complete rings of the ribbon vertices next to owned ones)."""
        return bool(self._info(24))

    def has_fans(self):
        """True when the patches store the oriented one-ring fans of their owned vertices."""
        return bool(self._info(23))

    def ribbon_overhead(self):
        """ribbon faces / F (patcher/patcher.h:139-142)."""
        return self.total_local(2) / self.get_num_faces() - 1.0

    def build_seconds(self, patcher_only=False):
        return float(lib().rxm_mesh_build_seconds(self._h, int(patcher_only)))

    def _arr(self, fn, elem, n):
        p = fn(self._h, int(elem)) if elem is not None else fn(self._h)
        return np.ctypeslib.as_array(p, shape=(n,)) if n else np.zeros(0, np.uint32)

    def slot_to_global(self, elem):
        return self._arr(lib().rxm_mesh_slot_to_global, elem, self.num_slots(elem))

    def global_to_slot(self, elem):
        return self._arr(lib().rxm_mesh_global_to_slot, elem, self._num(elem))

    def elem_patch(self, elem):
        return self._arr(lib().rxm_mesh_elem_patch, elem, self._num(elem))

    def slot_base(self, elem):
        return self._arr(lib().rxm_mesh_slot_base, elem, self.get_num_patches() + 1)

    def lin_base(self, elem):
        return self._arr(lib().rxm_mesh_lin_base, elem, self.get_num_patches() + 1)

    def edges(self):
        if not lib().rxm_mesh_edges(self._h):
            raise RXMeshError("the global edge arrays were released (rxm_mesh_compact)")
        return self._arr(lib().rxm_mesh_edges, None, 2 * self.get_num_edges()).reshape(-1, 2)

    def face_edges(self):
        if not lib().rxm_mesh_face_edges(self._h):
            raise RXMeshError("the global edge arrays were released (rxm_mesh_compact)")
        return self._arr(lib().rxm_mesh_face_edges, None, 3 * self.get_num_faces()).reshape(-1, 3)

    def patch(self, p):
        """host view of one patch: dict of numpy arrays (PatchInfo + ltog)."""
        v = PatchView()
        check(lib().rxm_mesh_patch(self._h, int(p), C.byref(v)))
        n, no = list(v.n), list(v.n_owned)

        def arr(ptr, cnt):
            return np.ctypeslib.as_array(ptr, shape=(cnt,)).copy() if cnt else np.zeros(0, np.uint32)

        ev_raw, fe_raw, fv_raw = arr(v.ev, 2 * n[1]), arr(v.fe, 3 * n[2]), arr(v.fv, 3 * n[2])
        pk = bool(v.packed)
        idm, fem = (0x7FF, 0xFFF) if pk else (0xFFFF, 0xFFFF)
        return dict(patch_id=int(v.patch_id), n=n, n_owned=no, slot_base=list(v.slot_base),
                    lin_base=list(v.lin_base), packed=pk,
                    ev=(ev_raw & idm).reshape(-1, 2), fe=(fe_raw & fem).reshape(-1, 3),
                    fv=(fv_raw & idm).reshape(-1, 3),
                    ev_rank=(ev_raw >> 11).reshape(-1, 2) if pk else None,
                    fv_rank=(fv_raw >> 11).reshape(-1, 3) if pk else None,
                    fe_rank=(fe_raw >> 12).reshape(-1, 3) if pk else None,
                    voff_e=arr(v.voff_e, n[0] + 1), voff_f=arr(v.voff_f, n[0] + 1),
                    eoff_f=arr(v.eoff_f, n[1] + 1),
                    fan_off=arr(v.fan_off, no[0] + 1) if v.fan_off else None,
                    fan_v=arr(v.fan_v, v.fan_total) if v.fan_off else None,
                    fan_f=arr(v.fan_f, v.fan_total) if v.fan_off else None,
                    ff=arr(v.ff, 3 * no[2]).reshape(-1, 3) if v.ff else None,
                    ef=arr(v.ef, 2 * no[1]).reshape(-1, 2) if v.ef else None,
                    fan_e=arr(v.fan_e, v.fan_total) if v.fan_e else None,
                    r2_idx=arr(v.r2_idx, n[0] - no[0] + v.n_ext) if v.r2_idx else None,
                    r2_off=arr(v.r2_off, v.n_r2 + 1) if v.r2_idx else None,
                    r2_val=arr(v.r2_val, v.r2_total) if v.r2_idx else None,
                    ext_owner=arr(v.ext_owner, v.n_ext) if v.r2_idx else None,
                    owner=[arr(v.owner[t], n[t] - no[t]) for t in range(3)],
                    stash=arr(v.stash, 4 * v.n_stash).reshape(-1, 4),
                    ltog=[arr(v.ltog[t], n[t]) for t in range(3)])

    def launch_box(self, op):
        """prepare_launch_box (rxmesh_static.inl:443-496): (blocks, threads, dynamic smem bytes)."""
        b, t, s = C.c_uint32(), C.c_uint32(), C.c_uint32()
        check(lib().rxm_mesh_launch_box(self._h, int(op), C.byref(b), C.byref(t), C.byref(s)))
        return b.value, t.value, s.value

    # ---- handles ----
    def map_to_global(self, elem, handles):
        """map_to_global (rxmesh_static.cu:669-685): 64-bit owner handles -> global ids."""
        h = np.asarray(handles, dtype=np.uint64)
        patch = (h >> np.uint64(32)).astype(np.int64)
        lid = (h & np.uint64(0xFFFF)).astype(np.int64)
        valid = h != np.uint64(INVALID64)
        sb = self.slot_base(elem).astype(np.int64)
        s2g = self.slot_to_global(elem)
        out = np.full(h.shape, 0xFFFFFFFF, dtype=np.uint32)
        idx = sb[np.where(valid, patch, 0)] + np.where(valid, lid, 0)
        out[valid] = s2g[idx[valid]]
        return out

    def linear_id(self, elem, handles):
        """Context::linear_id (context.h:275-290): prefix[patch] + local for owner handles."""
        h = np.asarray(handles, dtype=np.uint64)
        return (self.lin_base(elem).astype(np.int64)[(h >> np.uint64(32)).astype(np.int64)] +
                (h & np.uint64(0xFFFF)).astype(np.int64))

    # ---- attributes (rxmesh_static.h:608-806) ----
    def _add(self, elem, name, dtype, n, location, layout):
        if name in self._attrs:
            raise RXMeshError(f"attribute {name} already exists")
        a = Attribute(self, elem, dtype, n, location, layout, name)
        self._attrs[name] = a
        return a

    def add_vertex_attribute(self, name, dtype=np.float32, num_attributes=1, location=LOCATION_ALL,
                             layout=AoSoA, values=None):
        a = self._add(0, name, dtype, num_attributes if values is None else np.asarray(values).shape[1],
                      location, layout)
        if values is not None:
            a.from_global(values)
        return a

    def add_edge_attribute(self, name, dtype=np.float32, num_attributes=1, location=LOCATION_ALL,
                           layout=AoSoA):
        return self._add(1, name, dtype, num_attributes, location, layout)

    def add_face_attribute(self, name, dtype=np.float32, num_attributes=1, location=LOCATION_ALL,
                           layout=AoSoA):
        return self._add(2, name, dtype, num_attributes, location, layout)

    def does_attribute_exist(self, name):
        return name in self._attrs

    def remove_attribute(self, name):
        a = self._attrs.pop(name, None)
        if a is None:
            raise RXMeshError(f"attribute {name} does not exist")
        a.release()

    # ---- for_each (rxmesh_static.h:205-379), HOST side: yields owner handles ----
    def for_each(self, elem, fn):
        sb, lb = self.slot_base(elem), self.lin_base(elem)
        for p in range(self.get_num_patches()):
            for lid in range(int(lb[p + 1] - lb[p])):
                fn((p << 32) | lid)

    def for_each_vertex(self, location, fn):
        assert location & HOST, "device lambdas live in the C++ header API (include/rxmesh/)"
        self.for_each(0, fn)

    def for_each_edge(self, location, fn):
        assert location & HOST
        self.for_each(1, fn)

    def for_each_face(self, location, fn):
        assert location & HOST
        self.for_each(2, fn)

    # ---- fixed-function hot path ----
    def query_store(self, op, inp, out, stream=None):
        check(lib().rxm_query_store(self._h, int(op), inp._h, out._h, _stream_ptr(stream)))

    def query_consume(self, op, inp, out, stream=None):
        check(lib().rxm_query_consume(self._h, int(op), inp._h, out._h, _stream_ptr(stream)))

    def vertex_normals(self, coords, normals, unit_face_normals=False, stream=None):
        check(lib().rxm_vertex_normals(self._h, coords._h, normals._h, int(unit_face_normals),
                                       _stream_ptr(stream)))

    def laplacian_smooth(self, inp, out, lr, iters=1, stream=None):
        check(lib().rxm_laplacian_smooth(self._h, inp._h, out._h, float(lr), int(iters),
                                         _stream_ptr(stream)))

    def bilateral_filter(self, inp, out, iters=1, stream=None):
        check(lib().rxm_bilateral_filter(self._h, inp._h, out._h, int(iters), _stream_ptr(stream)))

    def mcf_solve(self, coords, out, time_step=10.0, use_uniform_laplace=True, max_iter=100, tol_abs=1e-6, tol_rel=0.0,
                  stream=None, precondition=False):
        """Mean-curvature flow by matrix-free CG (apps/MCF/mcf_cg_mat_free.h; defaults of apps/MCF/mcf.cu:19-23);
        precondition: the Jacobi-preconditioned solver of the same file (mcf_pcg_mat_free).
        Returns dict(iterations, converged, start_residual, final_residual) like the solver's getters."""
        buf = np.zeros(4, dtype=np.uint32)
        check(lib().rxm_mcf_solve_ex(self._h, coords._h, out._h, float(time_step), int(bool(use_uniform_laplace)),
                                     int(bool(precondition)), int(max_iter), float(tol_abs), float(tol_rel), buf.ctypes.data,
                                     _stream_ptr(stream)))
        f = buf.view(np.float32)
        return dict(iterations=int(buf[0]), converged=bool(buf[1]), start_residual=float(f[2]), final_residual=float(f[3]))

    def bilateral_deferred(self):
        """vertex-iterations of the last bilateral_filter call that ran on the cross-patch path"""
        return int(lib().rxm_bilateral_deferred(self._h))

    def query_csr(self, op, stream=None):
        """Materialised query in slot space, downloaded: (off[num_slots+1], val[nnz]) numpy arrays."""
        off, val, nnz = C.c_void_p(), C.c_void_p(), C.c_uint64()
        check(lib().rxm_query_csr(self._h, int(op), C.byref(off), C.byref(val), C.byref(nnz), _stream_ptr(stream)))
        ns = self.num_slots(_SRC[Op(op)])
        h_off = np.empty(ns + 1, dtype=np.uint32)
        h_val = np.empty(max(nnz.value, 1), dtype=np.uint32)
        check(lib().rxm_memcpy_d2h(h_off.ctypes.data_as(C.c_void_p), off, 4 * (ns + 1)))
        check(lib().rxm_memcpy_d2h(h_val.ctypes.data_as(C.c_void_p), val, 4 * nnz.value))
        return h_off, h_val[:nnz.value]

    def boundary_vertices(self, flag, stream=None):
        check(lib().rxm_boundary_vertices(self._h, flag._h, _stream_ptr(stream)))

    def vertex_normals_host(self, coords, out=None, stream=None):
        x = np.ascontiguousarray(coords, dtype=np.float32).reshape(-1, 3)
        if out is None:
            out = np.empty_like(x)
        check(lib().rxm_vertex_normals_host(self._h, x.ctypes.data_as(C.c_void_p),
                                            out.ctypes.data_as(C.c_void_p), _stream_ptr(stream)))
        return out

    def laplacian_smooth_host(self, coords, lr, iters=1, out=None, stream=None):
        x = np.ascontiguousarray(coords, dtype=np.float32).reshape(-1, 3)
        if out is None:
            out = np.empty_like(x)
        check(lib().rxm_laplacian_smooth_host(self._h, x.ctypes.data_as(C.c_void_p),
                                              out.ctypes.data_as(C.c_void_p), float(lr), int(iters),
                                              _stream_ptr(stream)))
        return out

    def query_consume_host(self, op, values, out=None, stream=None):
        op = Op(op)
        v = np.ascontiguousarray(values, dtype=np.float32).reshape(-1)
        if out is None:
            out = np.empty(self._num(_SRC[op]), dtype=np.float32)
        check(lib().rxm_query_consume_host(self._h, int(op), v.ctypes.data_as(C.c_void_p),
                                           out.ctypes.data_as(C.c_void_p), _stream_ptr(stream)))
        return out

    # raw-pointer variants (pinned torch tensors in bench.py): no numpy conversion
    def vertex_normals_host_ptr(self, in_ptr, out_ptr, stream=None):
        check(lib().rxm_vertex_normals_host(self._h, C.c_void_p(in_ptr), C.c_void_p(out_ptr),
                                            _stream_ptr(stream)))

    def query_consume_host_ptr(self, op, in_ptr, out_ptr, stream=None):
        check(lib().rxm_query_consume_host(self._h, int(op), C.c_void_p(in_ptr), C.c_void_p(out_ptr),
                                           _stream_ptr(stream)))

    # ---- helper used by the tests: run a query and return per-source global neighbour lists ----
    def query_global(self, op, width=None, stream=None, layout=AoSoA):
        op = Op(op)
        src, dst = _SRC[op], _DST[op]
        if width is None:
            width = {Op.EV: 2, Op.FV: 3, Op.FE: 3, Op.EE: 4, Op.EVDiamond: 4,
                     Op.EF: self.get_input_max_edge_incident_faces(),
                     Op.FF: self.get_input_max_face_adjacent_faces() + 2}.get(
                         op, self.get_input_max_valence())
        inp = Attribute(self, src, np.uint64, 1, LOCATION_ALL, layout)
        out = Attribute(self, src, np.uint64, width, LOCATION_ALL, layout)
        inp.reset(INVALID64, DEVICE, stream)
        out.reset(INVALID64, DEVICE, stream)
        self.query_store(op, inp, out, stream)
        inp.move(DEVICE, HOST, stream)
        out.move(DEVICE, HOST, stream)
        return inp, out, src, dst


def set_async(on):
    """rxm_set_async: *_host entry points return after enqueueing; sync the stream before reading results."""
    lib().rxm_set_async(int(bool(on)))


def launch_count():
    return int(lib().rxm_launch_count())
