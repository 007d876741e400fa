#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference, IN THE BUILD CONTAINER (needs /root/reference).

Run:  python tests/golden/make_golden.py            (all fixtures; ~2-3 min of CPU)
      python tests/golden/make_golden.py --quick    (skip the 1000-frame config-1 run)

Outputs (committed; the GPU box has no /root/reference):
  mesh_<name>.npz      vertices (body-local float32 [V,3]), tets uint32 [T,4], edges uint32 [E,2],
                       surface uint32 [S,3] parsed from the reference's committed Unity assets
                       Assets/SoftBody/Generated/<name>_Tet.asset.  Input DATA,
                       not source code.
  ref_<case>.npz       outputs of the UNMODIFIED reference (oracle/_ref/libpbdref.so =
                       CProgram/src/Sim.cpp compiled in place) stepped headless on those inputs:
                       inverse masses, rest values, positions / lambdas after given frame counts.
  ref_config1_1000.npz BASELINE.json config 1: default mesh, 10 substeps, 1000 frames:
                       positions at frames 1,10,100,1000 and the P3 residual metrics.
"""
from __future__ import annotations

import argparse
import importlib
import os
import re
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

pkg = importlib.import_module("cs121-softbodysim_b200")
mg = pkg.meshgen

ASSETS = "/root/reference/Assets/SoftBody/Generated"
MESHES = {"default": "default_Tet.asset", "icosphere001": "Icosphere.001_Tet.asset",
          "bunny": "Bunny-LowPoly_Tet.asset", "icosphere": "Icosphere_Tet.asset"}


def parse_asset(path):
    """Unity YAML SoftBodyTetMeshAsset: `vertices:` list of {x,y,z}; tetIds/edgeIds/surfaceTriIds
    are one line each of little-endian int32 hex (reference SoftBodyTetMeshAsset.cs fields)."""
    verts, ids = [], {}
    vre = re.compile(r"^\s*- \{x: ([^,]+), y: ([^,]+), z: ([^}]+)\}")
    with open(path, "r") as f:
        for line in f:
            mm = vre.match(line)
            if mm:
                verts.append([float(mm.group(1)), float(mm.group(2)), float(mm.group(3))])
                continue
            for key in ("tetIds", "edgeIds", "surfaceTriIds"):
                if line.startswith("  " + key + ":"):
                    hexs = line.split(":", 1)[1].strip()
                    ids[key] = np.frombuffer(bytes.fromhex(hexs), dtype="<i4").astype(np.uint32)
    return (np.asarray(verts, dtype=np.float32), ids["tetIds"].reshape(-1, 4),
            ids["edgeIds"].reshape(-1, 2), ids["surfaceTriIds"].reshape(-1, 3))


def config1_placement(local):
    """translate so min y = 1.0 (drops 1 unit onto y=0), identity rot."""
    return mg.place_body(local, rot=None, lowest_y=1.0)


def run_reference(x0, edges, tets, params, frames_at, pinned=None, dt=1.0 / 60.0):
    ref = po.Oracle(params, x0, edges, tets, pinned=pinned, kind="reference")
    out = {"w": ref.get(po.GET_W), "edge_rest": ref.get(po.GET_EDGE_REST), "tet_rest": ref.get(po.GET_TET_REST)}
    done = 0
    for fr in frames_at:
        ref.step(dt, fr - done)
        done = fr
        out[f"pos_{fr}"] = ref.positions()
        out[f"vel_{fr}"] = ref.get(po.GET_V)
        out[f"elam_{fr}"] = ref.get(po.GET_EDGE_LAMBDA)
    ref.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    po.build()
    assert po.have("reference"), "needs /root/reference to build oracle/_ref"

    meshes = {}
    for name, fn in MESHES.items():
        v, t, e, s = parse_asset(os.path.join(ASSETS, fn))
        meshes[name] = (v, t, e, s)
        np.savez_compressed(os.path.join(HERE, f"mesh_{name}.npz"), vertices=v, tets=t, edges=e, surface=s)
        print(f"mesh_{name}: V={len(v)} T={len(t)} E={len(e)} S={len(s)}")

    # short-horizon reference outputs on every asset mesh (config-1 placement, substeps=10)
    for name, (v, t, e, s) in meshes.items():
        p = po.Params.default(substeps=10)
        x0 = config1_placement(v)
        frames = (1, 10) if name == "default" else (1, 10, 100)
        out = run_reference(x0, e, t, p, frames)
        np.savez_compressed(os.path.join(HERE, f"ref_{name}.npz"), x0=x0, frames=np.array(frames),
                            params=np.frombuffer(bytes(p), dtype=np.uint8), **out)
        print(f"ref_{name}: frames {frames}")

    # pinned variant (top layer pinned, PBDRemoteSoftBody.cs:163-183) + nonzero volume compliance,
    # gravity with x/z components, ground off: exercises the other parameter paths.
    v, t, e, s = meshes["icosphere"]
    pins = mg.pin_top_layer(v)
    p = po.Params.default(substeps=4, iterations=3, volumeCompliance=1e-6, gx=0.5, gz=-0.25, groundEnabled=0)
    out = run_reference(config1_placement(v), e, t, p, (1, 10, 60), pinned=pins)
    np.savez_compressed(os.path.join(HERE, "ref_icosphere_pinned.npz"), x0=config1_placement(v), pinned=pins,
                        frames=np.array((1, 10, 60)), params=np.frombuffer(bytes(p), dtype=np.uint8), **out)

    # synthetic Kuhn grids (generator is ours; outputs are the reference's)
    for n, frames in ((3, (1, 10, 100)), (6, (1, 10, 100))):
        x0, t, e = mg.kuhn_grid(n)
        p = po.Params.default(substeps=10)
        out = run_reference(x0, e, t, p, frames)
        np.savez_compressed(os.path.join(HERE, f"ref_kuhn{n}.npz"), x0=x0, frames=np.array(frames),
                            params=np.frombuffer(bytes(p), dtype=np.uint8), **out)
        print(f"ref_kuhn{n}")

    if not args.quick:
        v, t, e, s = meshes["default"]
        x0 = config1_placement(v)
        p = po.Params.default(substeps=10)
        t0 = time.time()
        ref = po.Oracle(p, x0, e, t, kind="reference")
        out, done = {}, 0
        for fr in (1, 10, 100, 1000):
            ref.step(1.0 / 60.0, fr - done)
            done = fr
            pos = ref.positions()
            out[f"pos_{fr}"] = pos
            r = po.residuals(pos, x0, e, t)
            out[f"res_{fr}"] = np.array([r["edge_rms"], r["vol_rel"], r["tet_vol_rms"], r["min_y_dynamic"], float(r["finite"])])
            print(f"config1 frame {fr}: {r}  ({time.time() - t0:.1f}s)")
        st = ref.stats()
        out["ref_ms_per_frame"] = np.array([st["totalMs"] / 1000.0])
        np.savez_compressed(os.path.join(HERE, "ref_config1_1000.npz"), x0=x0,
                            params=np.frombuffer(bytes(p), dtype=np.uint8), **out)


if __name__ == "__main__":
    main()
