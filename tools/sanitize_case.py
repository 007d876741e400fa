"""Small end-to-end case for compute-sanitizer: tile backend (strict + interleaved, several tiles per
phase, point-to-point tile sync), batch backend and stream backend, a few frames each, checked
against the oracle."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("cs121-softbodysim_b200")
capi, mg = pkg.capi, pkg.meshgen
from oracle import pyoracle as po
x0, tets, edges = mg.kuhn_grid(7)
prm = dict(substeps=2, iterations=3)
for name, opt in (("tile/interleaved", capi.Options(backend=2, order_mode=1, tile_vertices=120)),
                  ("tile/strict", capi.Options(backend=2, order_mode=0, tile_vertices=200)),
                  ("stream", capi.Options(backend=1, flags=capi.FLAG_NO_GRAPH))):
    b = capi.Body(capi.SolverParams.default(**prm), x0, edges, tets, device=0, options=opt)
    o = po.Oracle(po.Params.default(**prm), x0, edges, tets, kind="port")
    o.permute_constraints(*b.schedule_order())
    seq = b.schedule_sequence()
    for _ in range(3):
        b.step(1 / 60)
        o.step_sequence(1 / 60, seq)
    print(name, "bit-exact:", np.array_equal(b.read_positions(), o.positions()), b.info()["tiles"], flush=True)
    b.close()
bodies = [(x0, edges, tets), (mg.kuhn_grid(4)[0], mg.kuhn_grid(4)[2], mg.kuhn_grid(4)[1])]
bt = capi.Batch(capi.SolverParams.default(**prm), bodies, device=0)
bt.step_async(1 / 60, 3); bt.sync()
o = po.Oracle(po.Params.default(**prm), x0, edges, tets, kind="port"); o.permute_constraints(*bt.schedule_order(0)); o.step(1 / 60, 3)
print("batch bit-exact:", np.array_equal(bt.body_positions(0, bt.read_positions()), o.positions()), flush=True)
bt.close()
