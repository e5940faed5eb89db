"""OBJ reader and procedural mesh generators (host side, numpy).

import_obj mirrors the contract of the reference's util/import_obj.h (positions
as float32 [V,3], triangles as uint32 [F,3], zero-based).  The generators follow
SURVEY.md section 8(d): create_plane semantics (geometry_factory.h:13-75) for
the grid, a class-I geodesic icosphere and a periodic quad torus.
"""
import numpy as np


def import_obj(path):
    verts, faces = [], []
    with open(path, "r") as fh:
        for line in fh:
            if line.startswith("v "):
                p = line.split()
                verts.append((float(p[1]), float(p[2]), float(p[3])))
            elif line.startswith("f "):
                p = line.split()[1:]
                idx = [int(t.split("/")[0]) for t in p]
                idx = [i - 1 if i > 0 else len(verts) + i for i in idx]
                for k in range(1, len(idx) - 1):  # fan-triangulate like rapidobj
                    faces.append((idx[0], idx[k], idx[k + 1]))
    return (np.asarray(verts, dtype=np.float32).reshape(-1, 3),
            np.asarray(faces, dtype=np.uint32).reshape(-1, 3))


def grid(nx, ny, dx=1.0, height=True):
    """create_plane(plane=1) semantics: vertex id = i*nx + j, x = dx*j, z = dx*i; per quad the
    triangles (a,b,c),(c,b,d) with a=idx, b=idx+nx, c=idx+1, d=idx+nx+1.
    height=True adds y = 0.05*sin(0.01 x)*cos(0.013 z)."""
    j = np.arange(nx, dtype=np.float32)
    i = np.arange(ny, dtype=np.float32)
    x = np.broadcast_to((dx * j)[None, :], (ny, nx))
    z = np.broadcast_to((dx * i)[:, None], (ny, nx))
    if height:
        y = (0.05 * np.sin(0.01 * x.astype(np.float64)) * np.cos(0.013 * z.astype(np.float64)))
    else:
        y = np.zeros((ny, nx))
    verts = np.stack([x, y.astype(np.float32), z], axis=-1).reshape(-1, 3).astype(np.float32)
    idx = (np.arange(ny - 1, dtype=np.uint32)[:, None] * np.uint32(nx) +
           np.arange(nx - 1, dtype=np.uint32)[None, :]).reshape(-1)
    a, b, c, d = idx, idx + np.uint32(nx), idx + np.uint32(1), idx + np.uint32(nx + 1)
    faces = np.empty((idx.shape[0], 2, 3), dtype=np.uint32)
    faces[:, 0, 0], faces[:, 0, 1], faces[:, 0, 2] = a, b, c
    faces[:, 1, 0], faces[:, 1, 1], faces[:, 1, 2] = c, b, d
    return verts, faces.reshape(-1, 3)


def grid_random_diagonals(n, seed=1):
    """an n x n grid (same vertices and height field as grid()) whose quads are split along a RANDOM diagonal: a manifold,
    consistently oriented mesh of mixed valence (4..8, mean 6) -- what real inputs look like to the valence-6 fast paths"""
    verts, _ = grid(n, n)
    idx = (np.arange(n - 1, dtype=np.uint32)[:, None] * np.uint32(n) + np.arange(n - 1, dtype=np.uint32)[None, :]).reshape(-1)
    a, b, c, d = idx, idx + np.uint32(n), idx + np.uint32(1), idx + np.uint32(n + 1)
    flip = np.random.RandomState(seed).rand(idx.shape[0]) < 0.5
    faces = np.empty((idx.shape[0], 2, 3), dtype=np.uint32)
    # (a,b,c),(c,b,d) or, the other diagonal with the same orientation, (a,b,d),(a,d,c)
    faces[:, 0, 0], faces[:, 0, 1], faces[:, 0, 2] = a, b, np.where(flip, d, c)
    faces[:, 1, 0], faces[:, 1, 1], faces[:, 1, 2] = np.where(flip, a, c), np.where(flip, d, b), np.where(flip, c, d)
    return verts, faces.reshape(-1, 3)


def grid_face_tiles(nx, ny, tile, tile_i=None):
    """Analytic patching of grid(): face -> patch id by tile (columns) x tile_i (rows) quad blocks
    (2*tile*tile_i faces per full patch), the role of the reference's dead
    Patcher::grid (patcher/patcher.cu:181-224)."""
    tile_i = tile if tile_i is None else tile_i
    qi = np.arange(ny - 1, dtype=np.uint32)[:, None] // np.uint32(tile_i)
    qj = np.arange(nx - 1, dtype=np.uint32)[None, :] // np.uint32(tile)
    ntj = (nx - 1 + tile - 1) // tile
    pid = (qi * np.uint32(ntj) + qj).reshape(-1)
    return np.repeat(pid, 2).astype(np.uint32)


def torus(nu, nv, R=1.0, r=0.35, noise=0.0, seed=12345):
    """Periodic nu x nv quad torus split into 2 triangles per quad."""
    u = np.arange(nu, dtype=np.float64) * (2 * np.pi / nu)
    v = np.arange(nv, dtype=np.float64) * (2 * np.pi / nv)
    U, V = np.meshgrid(u, v, indexing="ij")
    x = (R + r * np.cos(V)) * np.cos(U)
    y = (R + r * np.cos(V)) * np.sin(U)
    z = r * np.sin(V)
    verts = np.stack([x, y, z], axis=-1).reshape(-1, 3)
    if noise > 0:
        nrm = np.stack([np.cos(V) * np.cos(U), np.cos(V) * np.sin(U), np.sin(V)], -1).reshape(-1, 3)
        mean_edge = 0.5 * (2 * np.pi * R / nu + 2 * np.pi * r / nv)
        rng = np.random.RandomState(seed)
        verts = verts + nrm * (noise * mean_edge * (2 * rng.rand(verts.shape[0], 1) - 1))
    i = np.arange(nu, dtype=np.uint32)[:, None]
    j = np.arange(nv, dtype=np.uint32)[None, :]
    i1, j1 = (i + 1) % np.uint32(nu), (j + 1) % np.uint32(nv)
    a = (i * np.uint32(nv) + j).reshape(-1)
    b = (i1 * np.uint32(nv) + j).reshape(-1)
    c = (i * np.uint32(nv) + j1).reshape(-1)
    d = (i1 * np.uint32(nv) + j1).reshape(-1)
    faces = np.empty((a.shape[0], 2, 3), dtype=np.uint32)
    faces[:, 0, 0], faces[:, 0, 1], faces[:, 0, 2] = a, b, c
    faces[:, 1, 0], faces[:, 1, 1], faces[:, 1, 2] = c, b, d
    return verts.astype(np.float32), faces.reshape(-1, 3)


def torus_face_tiles(nu, nv, tile):
    qi = np.arange(nu, dtype=np.uint32)[:, None] // np.uint32(tile)
    qj = np.arange(nv, dtype=np.uint32)[None, :] // np.uint32(tile)
    ntj = (nv + tile - 1) // tile
    pid = (qi * np.uint32(ntj) + qj).reshape(-1)
    return np.repeat(pid, 2).astype(np.uint32)


_ICO_V = None


def _icosahedron():
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t],
                  [0, -1, -t], [0, 1, -t], [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]],
                 dtype=np.float64)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9],
                  [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2],
                  [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10],
                  [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    return v / np.linalg.norm(v, axis=1, keepdims=True), f


def icosphere(nu):
    """Class-I geodesic sphere: every icosahedron face split into nu^2 triangles;
    F = 20 nu^2, V = 10 nu^2 + 2, closed, edge-manifold, consistently oriented."""
    V0, F0 = _icosahedron()
    # unique ids: 12 corners, then (nu-1) per edge for 30 edges, then interior per face
    edges = {}
    verts = [V0]
    nvert = 12
    for f in F0:
        for k in range(3):
            a, b = int(f[k]), int(f[(k + 1) % 3])
            key = (min(a, b), max(a, b))
            if key not in edges:
                edges[key] = nvert
                if nu > 1:
                    t = (np.arange(1, nu, dtype=np.float64) / nu)[:, None]
                    verts.append(V0[key[0]] * (1 - t) + V0[key[1]] * t)
                nvert += nu - 1
    faces = []
    for f in F0:
        A, B, Cc = int(f[0]), int(f[1]), int(f[2])
        # barycentric lattice id(i,j): i steps toward B, j toward C, i+j<=nu
        ids = -np.ones((nu + 1, nu + 1), dtype=np.int64)

        def edge_id(p, q, s):  # s-th lattice point from p to q (1..nu-1)
            key = (min(p, q), max(p, q))
            base = edges[key]
            return base + (s - 1 if p == key[0] else nu - 1 - s)

        ii, jj = np.meshgrid(np.arange(nu + 1), np.arange(nu + 1), indexing="ij")
        interior = (ii > 0) & (jj > 0) & (ii + jj < nu)
        n_int = int(interior.sum())
        if n_int:
            ids[interior] = nvert + np.arange(n_int)
            bi = ii[interior][:, None] / nu
            bj = jj[interior][:, None] / nu
            verts.append(V0[A] * (1 - bi - bj) + V0[B] * bi + V0[Cc] * bj)
            nvert += n_int
        ids[0, 0], ids[nu, 0], ids[0, nu] = A, B, Cc
        for s in range(1, nu):
            ids[s, 0] = edge_id(A, B, s)
            ids[0, s] = edge_id(A, Cc, s)
            ids[nu - s, s] = edge_id(B, Cc, s)
        i, j = np.meshgrid(np.arange(nu), np.arange(nu), indexing="ij")
        m = (i + j) < nu
        up = np.stack([ids[i[m], j[m]], ids[i[m] + 1, j[m]], ids[i[m], j[m] + 1]], -1)
        m2 = (i + j) < nu - 1
        dn = np.stack([ids[i[m2] + 1, j[m2]], ids[i[m2] + 1, j[m2] + 1], ids[i[m2], j[m2] + 1]], -1)
        faces.append(up)
        faces.append(dn)
    V = np.concatenate(verts, 0)
    V = V / np.linalg.norm(V, axis=1, keepdims=True)
    return V.astype(np.float32), np.concatenate(faces, 0).astype(np.uint32)
