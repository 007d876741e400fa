"""Soak: many frames / many re-creations, looking for hangs or drift between two runs of the same
schedule (the point-to-point tile sync must make the result independent of timing)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("cs121-softbodysim_b200")
capi, mg = pkg.capi, pkg.meshgen
x0, tets, edges = mg.kuhn_grid(40)       # 384k tets: 144 tiles per phase, every SM busy
prm = capi.SolverParams.default(substeps=10)
res = []
t0 = time.time()
for rep in range(3):
    b = capi.Body(prm, x0, edges, tets, device=0, options=capi.Options(backend=2, order_mode=1))
    b.step_async(1 / 60, 600)
    ms = b.sync()
    res.append(b.read_positions())
    print(f"rep {rep}: 600 frames in {ms/1e3:.2f} s device time, min y {res[-1][:,1].min():.4f}, finite {np.isfinite(res[-1]).all()}", flush=True)
    b.close()
print("identical across runs:", all(np.array_equal(res[0], r) for r in res[1:]), f"wall {time.time()-t0:.1f} s")
m = np.load(os.path.join(ROOT, "tests", "golden", "mesh_default.npz"))
xd = mg.place_body(m["vertices"], lowest_y=1.0)
outs = []
for rep in range(2):
    b = capi.Body(capi.SolverParams.default(substeps=10), xd, m["edges"], m["tets"], device=0, options=capi.Options(backend=2, order_mode=1))
    b.step_async(1 / 60, 2000); b.sync(); outs.append(b.read_positions()); b.close()
print("default mesh 2000 frames identical across runs:", np.array_equal(outs[0], outs[1]), "min y", outs[0][:,1].min())
