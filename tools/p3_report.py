"""P3 on the GPU box: 1000 frames of BASELINE config 1 in every mode bench.py can run, the measured
windowed residuals and their RANK among the 12 constraint orders of the unmodified reference
(tests/golden/ref_config1_p3_window.npz).  Writes profiles/p3_residuals.json.

    gpurun -- 'python tools/p3_report.py > gpurun_out/p3_residuals.json'   (then copy to profiles/)
"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.build()
pkg = importlib.import_module("cs121-softbodysim_b200")
capi, mg = pkg.capi, pkg.meshgen
from oracle import pyoracle as po  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "ref_config1_p3_window.npz"))
window = [int(f) for f in g["window"]]
ref = g["residuals"].mean(axis=1)[:, :3]          # [12 orders, (edge_rms, vol_rel, tet_vol_rms)]
m = np.load(os.path.join(ROOT, "tests", "golden", "mesh_default.npz"))
x0, edges, tets = mg.place_body(m["vertices"], lowest_y=1.0), m["edges"], m["tets"]
MODES = [("stream/strict/exact", capi.BACKEND_STREAM, capi.ORDER_STRICT, 0),
         ("tile/strict/exact", capi.BACKEND_TILE, capi.ORDER_STRICT, 0),
         ("tile/interleaved/exact", capi.BACKEND_TILE, capi.ORDER_INTERLEAVED, 0),
         ("tile/interleaved/exact+tagged", capi.BACKEND_TILE, capi.ORDER_INTERLEAVED, capi.FLAG_TAGGED_HANDOVER),
         ("tile/interleaved/fast+tagged", capi.BACKEND_TILE, capi.ORDER_INTERLEAVED, capi.FLAG_TAGGED_HANDOVER | capi.FLAG_FAST_ARITH),
         ("tile/riding/exact", capi.BACKEND_TILE, capi.ORDER_RIDING, 0),
         ("tile/riding/fast+tagged", capi.BACKEND_TILE, capi.ORDER_RIDING, capi.FLAG_TAGGED_HANDOVER | capi.FLAG_FAST_ARITH)]
out = {"what": "BASELINE config 1 (default_Tet, 10 substeps x 6 iterations, dt = 1/60), 1000 frames; residuals averaged over "
               "frames 800..1000 (every 10th); rank = 1 + number of the reference's 12 constraint orders (original + 11 seeded "
               "permutations, unmodified Sim.cpp) with a SMALLER windowed mean; 13 = worse than all of them",
       "keys": ["edge_rms", "vol_rel", "tet_vol_rms"],
       "reference_orders": {"min": ref.min(0).tolist(), "median": np.median(ref, 0).tolist(), "max": ref.max(0).tolist(),
                            "original_order": ref[0].tolist()},
       "modes": {}}
for name, backend, order, flags in MODES:
    with capi.Body(capi.SolverParams.default(substeps=10), x0, edges, tets, device=0,
                   options=capi.Options(backend=backend, order_mode=order, flags=flags)) as b:
        done, rows, ok = 0, [], True
        for fr in window:
            b.step_async(1 / 60, fr - done)
            b.sync()
            done = fr
            r = po.residuals(b.read_positions(), x0, edges, tets)
            ok &= bool(r["finite"] and r["min_y_dynamic"] >= -1e-6)
            rows.append([r["edge_rms"], r["vol_rel"], r["tet_vol_rms"]])
        mean = np.mean(rows, axis=0)
        out["modes"][name] = {"backend": b.name(), "windowed_mean": mean.tolist(),
                              "rank_among_12_reference_orders": [int(1 + (ref[:, k] < mean[k]).sum()) for k in range(3)],
                              "ratio_to_reference_worst": (mean / ref.max(0)).tolist(), "finite_and_above_ground": ok,
                              "passes_1.10x_worst_bound": bool((mean <= 1.10 * ref.max(0)).all())}
    print(name, out["modes"][name], file=sys.stderr)
print(json.dumps(out, indent=1))
