"""Deterministic tet-mesh inputs for the PBDServer substep path (no RNG, no files).

* ``kuhn_grid(n)``  -- the synthetic "tetrahedralised cube" of BASELINE.json configs 2-5
  (SURVEY.md 8(d)): Kuhn/Freudenthal 6-tet split of an n^3-cell cube.
* ``build_edges(tets)`` -- unique tet edges in the order the Unity client produces them
  (reference ``Assets/Scripts/Softbody/PBDRemoteSoftBody.cs:253-285``: per tet the pairs
  ab, ac, ad, bc, bd, cd are inserted as (min,max) into a HashSet whose enumeration order,
  with no removals, is insertion order).  Bit-exact against the ``edgeIds`` stored in the
  reference's committed ``.asset`` meshes (tests/test_mesh_cpu.py, golden fixtures).
* ``place_body`` -- rigid placement of body-local vertices in world space; PBDServer takes
  world positions (``PBDRemoteSoftBody.cs:139-161``).
"""
from __future__ import annotations

import itertools
import math

import numpy as np

__all__ = ["kuhn_grid", "kuhn_counts", "build_edges", "place_body", "rotation_zx", "pin_top_layer"]


def kuhn_counts(n: int) -> tuple[int, int, int]:
    """(V, E, T) of ``kuhn_grid(n)``: V=(n+1)^3, T=6n^3, E=3n m^2 + 3n^2 m + n^3 (m=n+1)."""
    m = n + 1
    return m ** 3, 3 * n * m * m + 3 * n * n * m + n ** 3, 6 * n ** 3


def rotation_zx(deg_z: float = 20.0, deg_x: float = 10.0) -> np.ndarray:
    """Rz(deg_z) @ Rx(deg_x) as float64 3x3."""
    cz, sz = math.cos(math.radians(deg_z)), math.sin(math.radians(deg_z))
    cx, sx = math.cos(math.radians(deg_x)), math.sin(math.radians(deg_x))
    rz = np.array([[cz, -sz, 0.0], [sz, cz, 0.0], [0.0, 0.0, 1.0]])
    rx = np.array([[1.0, 0.0, 0.0], [0.0, cx, -sx], [0.0, sx, cx]])
    return rz @ rx


def place_body(local: np.ndarray, rot: np.ndarray | None = None, lowest_y: float | None = 0.25,
               translate=(0.0, 0.0, 0.0)) -> np.ndarray:
    """Rotate (float64), translate, optionally lift so min y == lowest_y; returns float32 [V,3]."""
    p = np.asarray(local, dtype=np.float64)
    if rot is not None:
        p = p @ np.asarray(rot, dtype=np.float64).T
    p = p + np.asarray(translate, dtype=np.float64)[None, :]
    if lowest_y is not None:
        p[:, 1] += lowest_y - p[:, 1].min()
    return np.ascontiguousarray(p.astype(np.float32))


def kuhn_grid(n: int, rot: np.ndarray | None = None, lowest_y: float | None = 0.25,
              size: float = 1.0, with_edges: bool = True):
    """Kuhn 6-tet split of an n^3-cell cube of edge ``size``.

    Vertex id = (k*m + j)*m + i with m = n+1 and position (i,j,k)*size/n; cells visited in
    k, j, i order; inside a cell the 6 axis permutations in lexicographic order, each giving
    the tet (c, c+e_p0, c+e_p0+e_p1, c+e_p0+e_p1+e_p2); vertices 1 and 2 are swapped when the
    signed volume is negative, so every tet is positively oriented.  The default placement is
    SURVEY.md 8(d)'s: rotate Rz20.Rx10, lowest vertex at y = 0.25.

    Returns (x0 float32 [V,3], tets uint32 [T,4], edges uint32 [E,2] or None).
    """
    if n < 1:
        raise ValueError("n must be >= 1")
    m = n + 1
    idx = np.arange(m, dtype=np.float64) * (size / n)
    kk, jj, ii = np.meshgrid(idx, idx, idx, indexing="ij")  # vertex id = (k*m+j)*m+i
    local = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], axis=1)

    ck, cj, ci = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    base = ((ck * m + cj) * m + ci).ravel().astype(np.int64)  # cell origin vertex, k,j,i order
    stride = np.array([1, m, m * m], dtype=np.int64)           # step along x(i), y(j), z(k)
    tets = np.empty((base.size, 6, 4), dtype=np.int64)
    for q, perm in enumerate(itertools.permutations(range(3))):  # lexicographic
        v0 = base
        v1 = v0 + stride[perm[0]]
        v2 = v1 + stride[perm[1]]
        v3 = v2 + stride[perm[2]]
        # orientation of the unrotated tet: sign of det[e_p0, e_p0+e_p1, e_p0+e_p1+e_p2] = sign(perm)
        sign = np.linalg.det(np.eye(3)[list(perm)])
        if sign < 0:
            v1, v2 = v2, v1
        tets[:, q, 0], tets[:, q, 1], tets[:, q, 2], tets[:, q, 3] = v0, v1, v2, v3
    tets = np.ascontiguousarray(tets.reshape(-1, 4).astype(np.uint32))

    if rot is None:
        rot = rotation_zx()
    x0 = place_body(local, rot=rot, lowest_y=lowest_y)
    edges = build_edges(tets) if with_edges else None
    return x0, tets, edges


def build_edges(tets: np.ndarray) -> np.ndarray:
    """Unique (min,max) tet edges in first-seen order, pairs ab, ac, ad, bc, bd, cd per tet."""
    t = np.asarray(tets).reshape(-1, 4).astype(np.int64)
    if t.shape[0] == 0:
        return np.zeros((0, 2), dtype=np.uint32)
    pa = t[:, [0, 0, 0, 1, 1, 2]].ravel()
    pb = t[:, [1, 2, 3, 2, 3, 3]].ravel()
    lo = np.minimum(pa, pb)
    hi = np.maximum(pa, pb)
    span = int(hi.max()) + 1
    key = lo * span + hi
    _, first = np.unique(key, return_index=True)
    first.sort()
    return np.ascontiguousarray(np.stack([lo[first], hi[first]], axis=1).astype(np.uint32))


def pin_top_layer(local: np.ndarray, eps: float = 1e-4) -> np.ndarray:
    """Indices with |y - max y| <= eps on the body-local vertices
    (reference ``PBDRemoteSoftBody.cs:163-183``)."""
    y = np.asarray(local, dtype=np.float32)[:, 1]
    return np.nonzero(np.abs(y - y.max()) <= np.float32(eps))[0].astype(np.uint32)
