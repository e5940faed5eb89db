"""ctypes/numpy front-end of oracle/rxmesh_oracle.c and oracle/_ref (TEST INFRASTRUCTURE ONLY).

Each wrapper names the C function it calls; the C functions cite the reference
file:line they restate.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

u32p = C.POINTER(C.c_uint32)
f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)
u8p = C.POINTER(C.c_uint8)


def build(force=False):
    """Compile librxmesh_oracle.so (and _ref/ when /root/reference exists)."""
    so = os.path.join(_HERE, "librxmesh_oracle.so")
    src = os.path.join(_HERE, "rxmesh_oracle.c")
    need = force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src)
    ref_so = os.path.join(_HERE, "_ref", "libvn_ref.so")
    if os.path.isdir("/root/reference/apps/VertexNormal") and not os.path.exists(ref_so):
        need = True
    if need:
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        build()
        _LIB = C.CDLL(os.path.join(_HERE, "librxmesh_oracle.so"))
        _LIB.rxo_build_edges.restype = C.c_uint32
        _LIB.rxo_query_ff.restype = C.c_uint64
        _LIB.rxo_boundary_vertices.restype = C.c_uint32
        _LIB.rxo_bilateral_step.restype = C.c_uint32
    return _LIB


def ref_lib():
    """The reference's own vertex_normal_ref.h compiled unmodified (None if absent)."""
    global _REF
    if _REF is None:
        build()
        p = os.path.join(_HERE, "_ref", "libvn_ref.so")
        if not os.path.exists(p):
            return None
        _REF = C.CDLL(p)
        _REF.ref_vertex_normal_f32_timed.restype = C.c_double
    return _REF


def _p(a, t):
    return a.ctypes.data_as(t)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Topology:
    """Global edge numbering + FE of a triangle soup (rxo_build_edges)."""

    def __init__(self, fv):
        self.fv = _u32(fv).reshape(-1, 3)
        nf = self.fv.shape[0]
        ev = np.empty((max(3 * nf, 1), 2), dtype=np.uint32)
        self.fe = np.empty((nf, 3), dtype=np.uint32)
        nv = C.c_uint32(0)
        ne = lib().rxo_build_edges(_p(self.fv, u32p), nf, _p(ev, u32p), _p(self.fe, u32p),
                                   C.byref(nv))
        self.ev = np.ascontiguousarray(ev[:ne])
        self.nf, self.ne, self.nv = nf, int(ne), int(nv.value)

    # --- the eight queries, as CSR (offsets, values) over global ids ---
    def _csr(self, fn, rows, nrows, ncols, nnz):
        off = np.empty(ncols + 1, dtype=np.uint32)
        val = np.empty(max(nnz, 1), dtype=np.uint32)
        fn(_p(rows, u32p), nrows, ncols, _p(off, u32p), _p(val, u32p))
        return off, val[:nnz]

    def query(self, op):
        op = op.upper()
        L = lib()
        if op == "VV":
            return self._csr(L.rxo_query_vv, self.ev, self.ne, self.nv, 2 * self.ne)
        if op == "VE":
            return self._csr(L.rxo_query_ve, self.ev, self.ne, self.nv, 2 * self.ne)
        if op == "VF":
            return self._csr(L.rxo_query_vf, self.fv, self.nf, self.nv, 3 * self.nf)
        if op == "EF":
            return self._csr(L.rxo_query_ef, self.fe, self.nf, self.ne, 3 * self.nf)
        if op == "EV":
            return np.arange(0, 2 * self.ne + 1, 2, dtype=np.uint32), self.ev.reshape(-1).copy()
        if op == "FV":
            return np.arange(0, 3 * self.nf + 1, 3, dtype=np.uint32), self.fv.reshape(-1).copy()
        if op == "FE":
            return np.arange(0, 3 * self.nf + 1, 3, dtype=np.uint32), self.fe.reshape(-1).copy()
        if op == "FF":
            off = np.empty(self.nf + 1, dtype=np.uint32)
            nnz = L.rxo_query_ff(_p(self.fe, u32p), self.nf, self.ne, _p(off, u32p), None)
            val = np.empty(max(int(nnz), 1), dtype=np.uint32)
            L.rxo_query_ff(_p(self.fe, u32p), self.nf, self.ne, _p(off, u32p), _p(val, u32p))
            return off, val[:int(nnz)]
        raise ValueError(op)

    def edge_dirs(self):
        """dir[f, j] = 0 if face f traverses its j-th edge from ev[e, 0] to ev[e, 1], else 1 (rxmesh.cpp:941-944)."""
        return (self.ev[self.fe, 0] != self.fv).astype(np.uint32)

    def ev_diamond(self):
        """e_v_diamond (kernels/rxmesh_queries.cuh:198-264): [ne, 4] = [v0, w0, v1, w1]; w_d = the vertex opposite to
        the edge in the face that traverses it with direction d; 0xFFFFFFFF where that face is missing."""
        out = np.full((self.ne, 4), 0xFFFFFFFF, dtype=np.uint32)
        out[:, 0], out[:, 2] = self.ev[:, 0], self.ev[:, 1]
        d = self.edge_dirs()
        for j in range(3):
            out[self.fe[:, j], 1 + 2 * d[:, j]] = self.fv[:, (j + 2) % 3]
        return out

    def ee(self):
        """e_e_manifold (kernels/rxmesh_queries.cuh:267-342) for consistently oriented edge-manifold input:
        [ne, 4] = for the face on side dir = 0 and 1: [next edge, previous edge] in that face's winding."""
        out = np.full((self.ne, 4), 0xFFFFFFFF, dtype=np.uint32)
        d = self.edge_dirs()
        for j in range(3):
            out[self.fe[:, j], 2 * d[:, j]] = self.fe[:, (j + 1) % 3]
            out[self.fe[:, j], 2 * d[:, j] + 1] = self.fe[:, (j + 2) % 3]
        return out

    def boundary_vertices(self):
        flags = np.zeros(max(self.nv, 1), dtype=np.uint8)
        n = lib().rxo_boundary_vertices(_p(self.ev, u32p), _p(self.fe, u32p), self.nf, self.ne,
                                        self.nv, _p(flags, u8p))
        return int(n), flags[:self.nv].astype(bool)

    def stats(self):
        out = np.zeros(5, dtype=np.uint32)
        lib().rxo_input_stats(_p(self.ev, u32p), _p(self.fe, u32p), self.nf, self.ne, self.nv,
                              _p(out, u32p))
        return dict(max_valence=int(out[0]), max_edge_incident_faces=int(out[1]),
                    max_face_adjacent_faces=int(out[2]), is_closed=bool(out[3]),
                    is_edge_manifold=bool(out[4]))


def vertex_normals(fv, x, dtype=np.float32):
    """rxo_vertex_normals_f32 / _f64 (apps/VertexNormal/vertex_normal_ref.h:5-85)."""
    fv = _u32(fv).reshape(-1, 3)
    x = _f32(x).reshape(-1, 3)
    if dtype == np.float32:
        n = np.empty_like(x)
        lib().rxo_vertex_normals_f32(_p(fv, u32p), fv.shape[0], _p(x, f32p), x.shape[0], _p(n, f32p))
    else:
        n = np.empty(x.shape, dtype=np.float64)
        lib().rxo_vertex_normals_f64(_p(fv, u32p), fv.shape[0], _p(x, f32p), x.shape[0], _p(n, f64p))
    return n


def ref_vertex_normals(fv, x, repeats=0):
    """The reference's own loop from oracle/_ref. repeats>0 -> (normals, seconds per run)."""
    R = ref_lib()
    if R is None:
        raise RuntimeError("oracle/_ref/libvn_ref.so not built (needs /root/reference)")
    fv = _u32(fv).reshape(-1, 3)
    x = _f32(x).reshape(-1, 3)
    n = np.empty_like(x)
    if repeats:
        t = R.ref_vertex_normal_f32_timed(_p(fv, u32p), fv.shape[0], _p(x, f32p), x.shape[0],
                                          _p(n, f32p), int(repeats))
        return n, float(t)
    R.ref_vertex_normal_f32(_p(fv, u32p), fv.shape[0], _p(x, f32p), x.shape[0], _p(n, f32p))
    return n


def vertex_normals_unit_faces(fv, x, dtype=np.float64):
    """rxo_vertex_normals_unit_faces (apps/Filtering/filtering_rxmesh_kernel.cuh:15-46)."""
    fv = _u32(fv).reshape(-1, 3)
    x = _f32(x).reshape(-1, 3)
    if dtype == np.float32:
        n = np.empty_like(x)
        lib().rxo_vertex_normals_unit_faces(_p(fv, u32p), fv.shape[0], _p(x, f32p), x.shape[0],
                                            _p(n, f32p), None)
    else:
        n = np.empty(x.shape, dtype=np.float64)
        lib().rxo_vertex_normals_unit_faces(_p(fv, u32p), fv.shape[0], _p(x, f32p), x.shape[0],
                                            None, _p(n, f64p))
    return n


def laplacian_step(vv, x, lr, dtype=np.float32):
    """rxo_laplacian_step_f32/_f64 (apps/Smoothing/manual.h:86-104)."""
    off, val = vv
    if dtype == np.float32:
        x = _f32(x).reshape(-1, 3)
        out = np.empty_like(x)
        lib().rxo_laplacian_step_f32(_p(off, u32p), _p(val, u32p), x.shape[0], _p(x, f32p),
                                     _p(out, f32p), C.c_double(lr))
    else:
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 3)
        out = np.empty_like(x)
        lib().rxo_laplacian_step_f64(_p(off, u32p), _p(val, u32p), x.shape[0], _p(x, f64p),
                                     _p(out, f64p), C.c_double(lr))
    return out


def bilateral_step(vv, fv, x, max_nbrs=80, use_f64=True):
    """One bilateral iteration = unit-face normals + rxo_bilateral_step
    (apps/Filtering/filtering_rxmesh.cuh:75-95). Returns (x_new, max neighbourhood).
    use_f64 = 2: neighbourhood membership decided in fp32 (the decisions a fp32 implementation makes), the rest in float64."""
    off, val = vv
    x = _f32(x).reshape(-1, 3)
    n = vertex_normals_unit_faces(fv, x, np.float64)
    out = np.empty_like(x)
    worst = lib().rxo_bilateral_step(_p(off, u32p), _p(val, u32p), x.shape[0], _p(x, f32p),
                                     _p(n, f64p), _p(out, f32p), int(max_nbrs), int(use_f64))
    return out, int(worst)


def bilateral_step_mt(vv, fv, x, threads, max_nbrs=80, use_f64=2):
    """rxo_bilateral_step_mt: the same iteration, vertices split statically over `threads` OpenMP threads (the reference's CPU
    side of the app: filtering_openmesh.h:112-116).  Returns (x_new, seconds of the filter step alone, normals excluded)."""
    import time
    off, val = vv
    x = _f32(x).reshape(-1, 3)
    n = vertex_normals_unit_faces(fv, x, np.float64)
    out = np.empty_like(x)
    fn = lib().rxo_bilateral_step_mt
    fn.restype = C.c_uint32
    t = time.perf_counter()
    fn(_p(off, u32p), _p(val, u32p), x.shape[0], _p(x, f32p), _p(n, f64p), _p(out, f32p), int(max_nbrs), int(use_f64), int(threads))
    return out, time.perf_counter() - t


def laplacian_step_mt(vv, x, lr, threads):
    """rxo_laplacian_step_f32_mt -> (x_new, seconds)."""
    import time
    off, val = vv
    x = _f32(x).reshape(-1, 3)
    out = np.empty_like(x)
    t = time.perf_counter()
    lib().rxo_laplacian_step_f32_mt(_p(off, u32p), _p(val, u32p), x.shape[0], _p(x, f32p), _p(out, f32p), C.c_double(lr), int(threads))
    return out, time.perf_counter() - t


def consume_sum(csr, src):
    """rxo_consume_sum: out[s] = sum_{t in list(s)} src[t] in float64."""
    off, val = csr
    src = _f32(src)
    out = np.empty(off.shape[0] - 1, dtype=np.float64)
    lib().rxo_consume_sum(_p(off, u32p), _p(val, u32p), off.shape[0] - 1, _p(src, f32p),
                          _p(out, f64p))
    return out


def max_threads():
    return int(lib().rxo_max_threads())


def vertex_normals_mt(fv, x, threads, repeats=1):
    """rxo_vertex_normals_f32_mt: all-cores port (per-thread accumulators). -> (normals, seconds per run)."""
    import time
    fv = _u32(fv).reshape(-1, 3)
    x = _f32(x).reshape(-1, 3)
    n = np.empty_like(x)
    scratch = np.empty((threads, x.shape[0], 3), dtype=np.float32)
    t0 = time.perf_counter()
    for _ in range(max(1, repeats)):
        lib().rxo_vertex_normals_f32_mt(_p(fv, u32p), fv.shape[0], _p(x, f32p), x.shape[0], _p(n, f32p),
                                        _p(scratch, f32p), int(threads))
    return n, (time.perf_counter() - t0) / max(1, repeats)


def consume_sum_mt(csr, src, threads):
    """rxo_consume_sum_f32_mt: out[s] = sum_{t in list(s)} src[t], fp32, rows split over `threads`."""
    off, val = csr
    src = _f32(src)
    out = np.empty(off.shape[0] - 1, dtype=np.float32)
    lib().rxo_consume_sum_f32_mt(_p(off, u32p), _p(val, u32p), off.shape[0] - 1, _p(src, f32p), _p(out, f32p),
                                 int(threads))
    return out


def csr_to_sets(csr):
    """list of sorted tuples (multiset per source element) for set-equality checks."""
    off, val = csr
    return [tuple(sorted(val[off[i]:off[i + 1]].tolist())) for i in range(off.shape[0] - 1)]


def oriented_rings(fv, nv):
    """rxo_oriented_rings: cyclic one-rings of a closed, consistently oriented manifold mesh (CSR)."""
    fv = _u32(fv).reshape(-1, 3)
    off = np.empty(nv + 1, dtype=np.uint32)
    val = np.empty(3 * fv.shape[0], dtype=np.uint32)
    bad = lib().rxo_oriented_rings(_p(fv, u32p), fv.shape[0], nv, _p(off, u32p), _p(val, u32p))
    if bad:
        raise ValueError("mesh is not closed / consistently oriented")
    return off, val


def mcf_matvec(rings, X, vec_in, time_step, with_scale=False):
    """rxo_mcf_matvec (apps/MCF/mcf_kernels.cuh:117-205), float64.  with_scale: also the per-vertex magnitude of the terms the
    result is a difference of (|diag| |in_p| + sum |w_i| |in_i|), the scale a fp32 evaluation is accurate against."""
    off, val = rings
    X, vin = _f32(X).reshape(-1, 3), _f32(vec_in).reshape(-1, 3)
    out = np.empty(X.shape, dtype=np.float64)
    scale = np.empty(X.shape[0], dtype=np.float64)
    lib().rxo_mcf_matvec_scaled(_p(off, u32p), _p(val, u32p), X.shape[0], _p(X, f32p), _p(vin, f32p), C.c_double(time_step),
                                _p(out, f64p), _p(scale, f64p))
    return (out, scale) if with_scale else out


def mcf_solve(rings, X0, time_step=10.0, uniform=True, max_iter=100, tol_abs=1e-6, tol_rel=0.0, with_residual=False,
              precond=False):
    """rxo_mcf_solve_ex (apps/MCF/mcf_cg_mat_free.h + matrix/cg_mat_free_attr_solver.h:45-125; precond: the Jacobi form of
    pcg_mat_free_attr_solver.h:40-140 with precond_matvec), float64 CG.
    Returns (X, info) with info = dict(iterations, converged, start_residual, final_residual) [, residual B - A X]."""
    off, val = rings
    X0 = _f32(X0).reshape(-1, 3)
    out = np.empty(X0.shape, dtype=np.float64)
    res = np.empty(X0.shape, dtype=np.float64)
    info = np.zeros(4, dtype=np.float64)
    rc = lib().rxo_mcf_solve_ex(_p(off, u32p), _p(val, u32p), X0.shape[0], _p(X0, f32p), C.c_double(time_step), int(bool(uniform)),
                                int(bool(precond)), int(max_iter), C.c_double(tol_abs), C.c_double(tol_rel), _p(out, f64p),
                                _p(res, f64p), _p(info, f64p))
    if rc:
        raise MemoryError("rxo_mcf_solve")
    d = dict(iterations=int(info[0]), converged=bool(info[1]), start_residual=float(info[2]), final_residual=float(info[3]))
    return (out, d, res) if with_residual else (out, d)


def mcf_residual(rings, X0, X, time_step=10.0, uniform=True):
    """B - A X (float64) for a candidate solution X of the MCF system built from X0, and <B, B>."""
    off, val = rings
    X0 = _f32(X0).reshape(-1, 3)
    X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, 3)
    res = np.empty(X0.shape, dtype=np.float64)
    bb = C.c_double(0)
    rc = lib().rxo_mcf_residual(_p(off, u32p), _p(val, u32p), X0.shape[0], _p(X0, f32p), C.c_double(time_step), int(bool(uniform)),
                                _p(X, f64p), _p(res, f64p), C.byref(bb))
    if rc:
        raise MemoryError("rxo_mcf_residual")
    return res, bb.value


def gaussian_curvature(fv, X):
    """rxo_gaussian_curvature (apps/GaussianCurvature/gaussian_curvature_kernel.cuh:10-69): (gcs, amix), float64."""
    fv, X = _u32(fv).reshape(-1, 3), _f32(X).reshape(-1, 3)
    g, a = np.empty(X.shape[0]), np.empty(X.shape[0])
    lib().rxo_gaussian_curvature(_p(fv, u32p), fv.shape[0], _p(X, f32p), X.shape[0], _p(g, f64p), _p(a, f64p))
    return g, a
