"""Headless asset front-end: everything the Unity client does between a tet mesh
asset and the MSG_INIT payload, so that any committed mesh can be stepped without Unity.

Mirrors, in the reference (paths relative to /root/reference/Assets/Scripts/Softbody):

===============================  =============================================================
here                             reference
===============================  =============================================================
``load_tet_asset``               the serialized ``SoftBodyTetMeshAsset`` (SoftBodyTetMeshAsset.cs:8-30):
                                 Unity YAML, ``vertices:`` as ``- {x: .., y: .., z: ..}`` lines, ``tetIds`` /
                                 ``edgeIds`` / ``surfaceTriIds`` as one line of little-endian int32 hex each
``orient_tets_positive``         ``OrientTetsPositive``     SoftBodyTetMeshAsset.cs:83-102 (swap b <-> c when
                                 the signed volume, :104-110, is negative)
``build_edges_and_surface``      ``BuildEdgesAndSurface``   SoftBodyTetMeshAsset.cs:139-203 ==
                                 PBDRemoteSoftBody.cs:247-316: unique edges in first-seen order (pairs ab, ac,
                                 ad, bc, bd, cd), boundary faces (seen once) wound outward
``meshgen.pin_top_layer``        ``BuildPinnedTopLayer``    PBDRemoteSoftBody.cs:163-183
``world_positions``              ``BuildWorldPositions``    PBDRemoteSoftBody.cs:139-161 (local -> world)
``init_payload``                 ``SendInit``               PBDRemoteWorld.cs:278-349
===============================  =============================================================

Known-answer check (tests/test_assets_cpu.py): rebuilding edges and surface triangles from the ``tetIds``
of every committed reference asset reproduces the asset's stored ``edgeIds`` / ``surfaceTriIds`` exactly.
"""
from __future__ import annotations

import re
from dataclasses import dataclass

import numpy as np

from . import meshgen


@dataclass
class TetMesh:
    vertices: np.ndarray      # float32 [V,3], body-local
    tets: np.ndarray          # uint32 [T,4]
    edges: np.ndarray         # uint32 [E,2]
    surface: np.ndarray       # uint32 [S,3] boundary triangles, outward winding


_VERT = re.compile(r"^\s*- \{x: ([^,]+), y: ([^,]+), z: ([^}]+)\}")


def load_tet_asset(path: str) -> TetMesh:
    verts, ids = [], {}
    in_vertices = False
    with open(path, "r") as f:
        for line in f:
            s = line.strip()
            if s.startswith("vertices:"):
                in_vertices = True
                continue
            m = _VERT.match(line) if in_vertices else None
            if m:
                verts.append([float(m.group(1)), float(m.group(2)), float(m.group(3))])
                continue
            if in_vertices and s and not s.startswith("-"):
                in_vertices = False
            for key in ("tetIds", "edgeIds", "surfaceTriIds"):
                if s.startswith(key + ":"):
                    hexs = s.split(":", 1)[1].strip()
                    ids[key] = np.frombuffer(bytes.fromhex(hexs), dtype="<i4").astype(np.uint32)
    for key in ("tetIds", "edgeIds", "surfaceTriIds"):
        if key not in ids:
            raise ValueError(f"{path}: no '{key}:' line -- not a SoftBodyTetMeshAsset")
    v = np.asarray(verts, dtype=np.float32).reshape(-1, 3)
    mesh = TetMesh(v, ids["tetIds"].reshape(-1, 4), ids["edgeIds"].reshape(-1, 2), ids["surfaceTriIds"].reshape(-1, 3))
    for name, a in (("tetIds", mesh.tets), ("edgeIds", mesh.edges), ("surfaceTriIds", mesh.surface)):
        if a.size and int(a.max()) >= len(v):
            raise ValueError(f"{path}: {name} refers to vertex {int(a.max())} of {len(v)}")
    return mesh


def save_tet_asset(path: str, mesh: TetMesh, name: str = "Generated_Tet"):
    """Write the same YAML layout (used for fixtures; the reference creates these from the editor,
    SoftBodyTetMeshAsset.cs:33-81)."""
    def hexline(a):
        return np.ascontiguousarray(a, dtype="<i4").tobytes().hex()
    with open(path, "w") as f:
        f.write("%YAML 1.1\n%TAG !u! tag:unity3d.com,2011:\n--- !u!114 &11400000\nMonoBehaviour:\n")
        f.write(f"  m_Name: {name}\n  vertices:\n")
        for x, y, z in np.asarray(mesh.vertices, dtype=np.float32):
            f.write(f"  - {{x: {float(x)!r}, y: {float(y)!r}, z: {float(z)!r}}}\n")
        f.write(f"  tetIds: {hexline(mesh.tets)}\n  edgeIds: {hexline(mesh.edges)}\n  surfaceTriIds: {hexline(mesh.surface)}\n")


def signed_volumes(vertices: np.ndarray, tets: np.ndarray) -> np.ndarray:
    """TetSignedVolume (SoftBodyTetMeshAsset.cs:104-110) in float32, like the C# floats."""
    v = np.asarray(vertices, dtype=np.float32)
    t = np.asarray(tets, dtype=np.int64).reshape(-1, 4)
    a, b, c = v[t[:, 1]] - v[t[:, 0]], v[t[:, 2]] - v[t[:, 0]], v[t[:, 3]] - v[t[:, 0]]
    return (np.einsum("ij,ij->i", np.cross(a, b).astype(np.float32), c).astype(np.float32) / np.float32(6.0)).astype(np.float32)


def orient_tets_positive(vertices: np.ndarray, tets: np.ndarray) -> np.ndarray:
    t = np.array(tets, dtype=np.uint32).reshape(-1, 4)
    neg = signed_volumes(vertices, t) < 0
    t[neg, 1], t[neg, 2] = t[neg, 2].copy(), t[neg, 1].copy()
    return t


def build_edges_and_surface(vertices: np.ndarray, tets: np.ndarray):
    """(edges uint32 [E,2], surface uint32 [S,3]).  Edges: (min, max) keys in first-seen order.  Faces:
    per tet (a,b,c | d), (a,d,b | c), (a,c,d | b), (b,d,c | a); a face seen exactly once is boundary; it
    keeps the vertex order of its first sighting, flipped when its normal points at the opposite vertex."""
    v = np.asarray(vertices, dtype=np.float32)
    t = np.asarray(tets, dtype=np.int64).reshape(-1, 4)
    edges = meshgen.build_edges(t)
    if len(t) == 0:
        return edges, np.zeros((0, 3), np.uint32)
    faces = np.stack([t[:, [0, 1, 2, 3]], t[:, [0, 3, 1, 2]], t[:, [0, 2, 3, 1]], t[:, [1, 3, 2, 0]]], axis=1).reshape(-1, 4)
    key = np.sort(faces[:, :3], axis=1)
    span = int(t.max()) + 1
    k = (key[:, 0] * span + key[:, 1]) * span + key[:, 2]
    _, first, count = np.unique(k, return_index=True, return_counts=True)
    first = np.sort(first[count == 1])                       # dictionary enumeration = insertion order
    f = faces[first]
    p0, p1, p2, po = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]], v[f[:, 3]]
    s = np.einsum("ij,ij->i", np.cross(p1 - p0, p2 - p0).astype(np.float32), po - p0)
    flip = s > 0
    f[flip, 1], f[flip, 2] = f[flip, 2].copy(), f[flip, 1].copy()
    return edges, np.ascontiguousarray(f[:, :3].astype(np.uint32))


def world_positions(local: np.ndarray, position=(0.0, 0.0, 0.0), rotation: np.ndarray | None = None, scale=1.0) -> np.ndarray:
    """transform.TransformPoint on every local vertex: scale, rotate, translate (PBDRemoteSoftBody.cs:152-160)."""
    p = np.asarray(local, dtype=np.float64) * np.asarray(scale, dtype=np.float64)
    if rotation is not None:
        p = p @ np.asarray(rotation, dtype=np.float64).T
    return np.ascontiguousarray((p + np.asarray(position, dtype=np.float64)[None, :]).astype(np.float32))


def init_payload(mesh: TetMesh, params, position=(0.0, 0.0, 0.0), rotation=None, scale=1.0, pin_top: bool = False) -> bytes:
    """The MSG_INIT payload the Unity client would send for this asset (PBDRemoteWorld.cs:278-349)."""
    from . import capi
    x0 = world_positions(mesh.vertices, position, rotation, scale)
    pinned = meshgen.pin_top_layer(mesh.vertices) if pin_top else None
    return capi.pack_init_payload(params, x0, mesh.edges, mesh.tets, pinned)
