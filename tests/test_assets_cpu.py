"""Headless asset front-end (cs121-softbodysim_b200/assets.py) against the reference's
committed meshes: tests/golden/mesh_*.npz hold vertices / tetIds / edgeIds / surfaceTriIds parsed from
Assets/SoftBody/Generated/*_Tet.asset by tests/golden/make_golden.py.  Known answers: rebuilding edges
and boundary triangles from tetIds alone reproduces what the reference's BuildEdgesAndSurface stored."""
import glob
import importlib
import os

import numpy as np
import pytest

MESHES = ["icosphere", "bunny", "icosphere001", "default"]
ASSET_DIR = "/root/reference/Assets/SoftBody/Generated"
ASSET_OF = {"default": "default_Tet.asset", "icosphere001": "Icosphere.001_Tet.asset", "bunny": "Bunny-LowPoly_Tet.asset",
            "icosphere": "Icosphere_Tet.asset"}


@pytest.fixture(scope="module")
def assets(pkg):
    return importlib.import_module("cs121-softbodysim_b200.assets")


@pytest.mark.parametrize("mesh", MESHES)
def test_edges_and_surface_rebuilt_from_tets_match_the_reference_asset(mesh, assets, golden):
    m = golden(f"mesh_{mesh}.npz")
    edges, surface = assets.build_edges_and_surface(m["vertices"], m["tets"])
    assert np.array_equal(edges, m["edges"])                      # first-seen order of the HashSet, SoftBodyTetMeshAsset.cs:143-175
    assert np.array_equal(surface, m["surface"])                  # boundary faces, insertion order, outward winding (:177-203)
    assert (assets.signed_volumes(m["vertices"], m["tets"]) > 0).all()
    assert np.array_equal(assets.orient_tets_positive(m["vertices"], m["tets"]), m["tets"])   # already positive: untouched


def test_orient_tets_positive_flips_negative_tets(assets, meshgen):
    x0, tets, _ = meshgen.kuhn_grid(3, rot=np.eye(3), lowest_y=None)
    bad = tets.copy()
    bad[::3, [1, 2]] = bad[::3, [2, 1]]                           # flip every third tet
    assert (assets.signed_volumes(x0, bad)[::3] < 0).all()
    fixed = assets.orient_tets_positive(x0, bad)
    assert np.array_equal(fixed, tets) and (assets.signed_volumes(x0, fixed) > 0).all()
    # the boundary of a cube split into tets: 2 triangles per boundary cell face, every edge of the surface shared by two triangles
    edges, surf = assets.build_edges_and_surface(x0, tets)
    assert len(surf) == 6 * 3 * 3 * 2 and len(edges) == len(meshgen.build_edges(tets))
    und = np.sort(np.concatenate([surf[:, [0, 1]], surf[:, [1, 2]], surf[:, [2, 0]]]), axis=1)
    _, cnt = np.unique(und, axis=0, return_counts=True)
    assert (cnt == 2).all()                                       # closed surface
    c = x0.mean(0)                                                # outward: normal . (centroid of triangle - body centre) > 0
    n = np.cross(x0[surf[:, 1]] - x0[surf[:, 0]], x0[surf[:, 2]] - x0[surf[:, 0]])
    assert (np.einsum("ij,ij->i", n, x0[surf].mean(1) - c) > 0).all()


def test_asset_round_trip_and_init_payload(assets, capi, meshgen, golden, tmp_path):
    m = golden("mesh_icosphere.npz")
    mesh = assets.TetMesh(m["vertices"], m["tets"], m["edges"], m["surface"])
    p = str(tmp_path / "Icosphere_Tet.asset")
    assets.save_tet_asset(p, mesh)
    back = assets.load_tet_asset(p)
    for a, b in ((back.vertices, mesh.vertices), (back.tets, mesh.tets), (back.edges, mesh.edges), (back.surface, mesh.surface)):
        assert np.array_equal(a, b)
    prm = capi.SolverParams.default(substeps=10)
    pay = assets.init_payload(back, prm, position=(0.0, 2.0, 0.0), pin_top=True)
    pins = meshgen.pin_top_layer(mesh.vertices)
    assert len(pay) == capi.lib().pbd_init_payload_size(len(mesh.vertices), len(mesh.edges), len(mesh.tets), len(pins))
    V, E, T = np.frombuffer(pay[:12], "<u4")
    assert (V, E, T) == (len(mesh.vertices), len(mesh.edges), len(mesh.tets))
    x0 = np.frombuffer(pay[64 + 4 * len(pins):64 + 4 * len(pins) + 12 * V], "<f4").reshape(-1, 3)
    assert np.allclose(x0, mesh.vertices + np.array([0, 2, 0], np.float32))
    with open(p, "a") as f:
        f.write("  tetIds: 0000\n")
    with pytest.raises(ValueError):
        assets.load_tet_asset(p)                                  # hex that is not a whole number of int32 / indices out of range


@pytest.mark.parametrize("mesh", MESHES)
def test_reference_assets_parse_to_the_committed_fixtures(mesh, assets, golden):
    path = os.path.join(ASSET_DIR, ASSET_OF[mesh])
    if not os.path.exists(path):
        pytest.skip("the reference tree is not on this machine")
    a = assets.load_tet_asset(path)
    m = golden(f"mesh_{mesh}.npz")
    assert np.array_equal(a.vertices, m["vertices"]) and np.array_equal(a.tets, m["tets"])
    assert np.array_equal(a.edges, m["edges"]) and np.array_equal(a.surface, m["surface"])
    # the auto-generated copies of the asset (SoftBodyTetMeshAsset.cs:57-75) parse too, and for each of them the
    # stored edges / surface are what BuildEdgesAndSurface derives from its own tets
    stem = ASSET_OF[mesh][:-len("_Tet.asset")]
    for q in sorted(glob.glob(os.path.join(ASSET_DIR, stem + "_Tet*.asset")))[:4]:
        d = assets.load_tet_asset(q)
        e, s_ = assets.build_edges_and_surface(d.vertices, d.tets)
        assert d.vertices.shape == a.vertices.shape and np.array_equal(e, d.edges) and np.array_equal(s_, d.surface)
