"""Single-process multi-GPU mode (rxm_multi_* of the C ABI, rxmesh_b200/csrc/rxm_multi.cu): one host process, one shard of
the mesh per device, ribbon rows pushed over NVLink by the kernel that computes them.  The planning (shards, ghost rings,
owner matching, push lists) happens in the library; this module is only the ctypes front-end.  The one-process-per-GPU
form over torch.distributed lives in rxmesh_b200/distributed.py."""
import ctypes as C

import numpy as np

from ._lib import check, lib


class RXMeshMulti:
    """devices: list of CUDA device ids, or an int n with device=False for a host-only plan of n shards (tests)."""

    def __init__(self, faces, devices, face_patch=None, patch_size=512, num_threads=0, device=True):
        self._fv = np.ascontiguousarray(faces, dtype=np.uint32).reshape(-1, 3)
        fp = None if face_patch is None else np.ascontiguousarray(face_patch, dtype=np.uint32)
        if device:
            devs = np.ascontiguousarray(list(devices), dtype=np.int32)
            n, dptr = len(devs), devs.ctypes.data_as(C.c_void_p)
        else:
            n, dptr = int(devices), None
        h = C.c_void_p()
        check(lib().rxm_multi_create(self._fv.ctypes.data_as(C.c_void_p), self._fv.shape[0],
                                     None if fp is None else fp.ctypes.data_as(C.c_void_p), int(patch_size), dptr, n,
                                     int(num_threads), C.byref(h)))
        self._h, self.num_shards = h, n

    def __del__(self):
        if getattr(self, "_h", None):
            lib().rxm_multi_destroy(self._h)
            self._h = None

    def info(self, what, shard=-1):
        return int(lib().rxm_multi_info(self._h, int(what), int(shard)))

    def halo_elements(self):
        return self.info(2)

    def get_num_vertices(self):
        return self.info(3)

    def laplacian_smooth(self, coords, lr, iters=1):
        x = np.ascontiguousarray(coords, dtype=np.float32).reshape(-1, 3)
        out = np.empty_like(x)
        check(lib().rxm_multi_laplacian_smooth(self._h, x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                               float(lr), int(iters)))
        return out

    def vertex_normals(self, coords):
        x = np.ascontiguousarray(coords, dtype=np.float32).reshape(-1, 3)
        out = np.empty_like(x)
        check(lib().rxm_multi_vertex_normals(self._h, x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
        return out
