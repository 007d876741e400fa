"""Schedule statistics of the tile planner, host only (no GPU): colour steps per tile visit, step
sizes, per-vertex loads, per-visit work spread.  What the planner changes of this round were
judged by before going to the GPU.

    python tools/plan_report.py [kuhn_n=56] [--strict] [--partitions K]
    PBD_PLAN_DEBUG=1 python tools/plan_report.py 80        # + wall time of every planning stage
"""
import argparse
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
capi = importlib.import_module("cs121-softbodysim_b200.capi")
meshgen = importlib.import_module("cs121-softbodysim_b200.meshgen")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("n", nargs="?", type=int, default=56)
    ap.add_argument("--strict", action="store_true")
    ap.add_argument("--riding", action="store_true")
    ap.add_argument("--partitions", type=int, default=0)
    ap.add_argument("--tiles-per-sm", type=int, default=0)
    ap.add_argument("--block-threads", type=int, default=0)
    ap.add_argument("--tile-vertices", type=int, default=0)
    ap.add_argument("--no-tagged", action="store_true")
    a = ap.parse_args()
    x0, tets, edges = meshgen.kuhn_grid(a.n)
    om = capi.ORDER_STRICT if a.strict else capi.ORDER_RIDING if a.riding else capi.ORDER_INTERLEAVED
    t0 = time.time()
    p = capi.Plan(x0, edges, tets, capi.Options(backend=capi.BACKEND_TILE, order_mode=om, partitions=a.partitions,
                                             tiles_per_sm=a.tiles_per_sm, block_threads=a.block_threads, tile_vertices=a.tile_vertices,
                                             flags=0 if a.no_tagged else capi.FLAG_TAGGED_HANDOVER))
    i = p.info()
    print(f"Kuhn n={a.n}: V={len(x0)} E={len(edges)} T={len(tets)}  order={'strict' if a.strict else 'interleaved'}  "
          f"tile visits/iteration={i['tiles']}  planned in {time.time() - t0:.1f} s")
    (eph, etl, eco), (tph, ttl, tco) = p.slots(False), p.slots(True)
    nco = int(max(eco.max(initial=0), tco.max(initial=0))) + 1
    ke, ce = np.unique(etl.astype(np.int64) * nco + eco, return_counts=True)
    kt, ct = np.unique(ttl.astype(np.int64) * nco + tco, return_counts=True)
    allk = np.union1d(ke, kt)
    ne = np.zeros(len(allk), int)
    nt = np.zeros(len(allk), int)
    ne[np.searchsorted(allk, ke)] = ce
    nt[np.searchsorted(allk, kt)] = ct
    visit = allk // nco
    uv, steps = np.unique(visit, return_counts=True)
    mixed = bool(((ne > 0) & (nt > 0)).any())
    print(f"colour steps per visit: mean {steps.mean():.2f}  histogram { {int(k): int(v) for k, v in zip(*np.unique(steps, return_counts=True))} }"
          f"  ({'mixed edge+tet steps' if mixed else 'edge steps and tet steps apart'})")
    print(f"per step: edges mean {ne[ne > 0].mean():.0f} max {ne.max()}, tets mean {nt[nt > 0].mean():.0f} max {nt.max()}, "
          f"warps mean {(((ne + 31) // 32) + ((nt + 31) // 32)).mean():.1f}")
    K = int(max(eph.max(initial=0), tph.max(initial=0))) + 1
    load = np.zeros((len(x0), K), dtype=np.int32)
    for j in range(2):
        np.add.at(load, (edges[:, j].astype(np.int64), eph.astype(np.int64)), 1)
    for j in range(4):
        np.add.at(load, (tets[:, j].astype(np.int64), tph.astype(np.int64)), 1)
    print(f"constraints per (vertex, phase): histogram {np.bincount(load.ravel()).tolist()}")
    # cost model fitted to the step trace: ~570 cycles per step + 0.5 per edge + 1.7 per tet
    cost = np.zeros(len(uv))
    np.add.at(cost, np.searchsorted(uv, visit), 570 + 0.5 * ne + 1.7 * nt)
    print(f"modelled sweep cycles per visit: mean {cost.mean():.0f}  min {cost.min():.0f}  max {cost.max():.0f}")


if __name__ == "__main__":
    main()
