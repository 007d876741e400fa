"""Indexing known answers (P0): the edge builder reproduces the edgeIds stored in the reference's
committed Unity assets exactly; the synthetic Kuhn grid has the stated counts and is a valid mesh."""
import numpy as np
import pytest


@pytest.mark.parametrize("name,V,T,E", [("default", 8613, 26070, 41488), ("icosphere001", 2562, 7175, 12296),
                                        ("bunny", 276, 798, 1297), ("icosphere", 162, 425, 746)])
def test_edge_builder_matches_asset_edges_bit_exact(name, V, T, E, golden, meshgen):
    m = golden(f"mesh_{name}.npz")
    assert m["vertices"].shape == (V, 3) and m["tets"].shape == (T, 4) and m["edges"].shape == (E, 2)
    assert np.array_equal(meshgen.build_edges(m["tets"]), m["edges"])


@pytest.mark.parametrize("n", [1, 2, 5, 10])
def test_kuhn_grid_counts_orientation_conformity(n, meshgen):
    x0, tets, edges = meshgen.kuhn_grid(n)
    V, E, T = meshgen.kuhn_counts(n)
    assert x0.shape == (V, 3) and tets.shape == (T, 4) and edges.shape == (E, 2)
    p = x0.astype(np.float64)
    a, b, c, d = (p[tets[:, k]] for k in range(4))
    vol = np.einsum("ij,ij->i", np.cross(b - a, c - a), d - a) / 6.0
    assert (vol > 0).all()
    np.testing.assert_allclose(vol.sum(), 1.0, rtol=1e-5)
    np.testing.assert_allclose(x0[:, 1].min(), 0.25, atol=1e-6)
    # conforming: every face belongs to 1 (boundary) or 2 tets; boundary faces = 12 n^2
    faces = np.sort(np.concatenate([tets[:, [0, 1, 2]], tets[:, [0, 1, 3]], tets[:, [0, 2, 3]], tets[:, [1, 2, 3]]]), axis=1)
    _, cnt = np.unique(faces, axis=0, return_counts=True)
    assert cnt.max() == 2 and (cnt == 1).sum() == 12 * n * n
    if n >= 2:
        val = np.bincount(tets.ravel(), minlength=V)
        assert val.max() == 24                        # interior tet valence
        assert np.bincount(edges.ravel(), minlength=V).max() == 14


def test_kuhn_headline_sizes(meshgen):
    assert meshgen.kuhn_counts(26) == (19683, 129194, 105456)       # config 2
    assert meshgen.kuhn_counts(56) == (185193, 1257704, 1053696)    # config 3 (headline)
    assert meshgen.kuhn_counts(10) == (1331, 7930, 6000)            # config 4 body
    assert meshgen.kuhn_counts(175) == (5451776, 37791775, 32156250)  # config 5


def test_pin_top_layer(meshgen, golden):
    v = golden("mesh_icosphere.npz")["vertices"]
    pins = meshgen.pin_top_layer(v)
    assert np.array_equal(pins, golden("ref_icosphere_pinned.npz")["pinned"]) and pins.size >= 1
