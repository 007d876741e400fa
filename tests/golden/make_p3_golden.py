#!/usr/bin/env python
"""P3 golden: how the REFERENCE's own 1000-frame residuals on BASELINE config 1 vary with the
constraint order.  Trajectories are chaotic after contact, so "no worse than the
reference" is judged against the reference's own spread: the unmodified reference
(oracle/_ref) is run on the original order and on 11 seeded random permutations of the edge and
tet arrays; residuals are averaged over frames 800..1000 (every 10th frame).

Run in the build container:  python tests/golden/make_p3_golden.py   (~5 min on 6 cores)
Writes tests/golden/ref_config1_p3_window.npz.
"""
import importlib
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

WINDOW = list(range(800, 1001, 10))
KEYS = ("edge_rms", "vol_rel", "tet_vol_rms", "min_y_dynamic")


def run(seed):
    from oracle import pyoracle as po
    mg = importlib.import_module("cs121-softbodysim_b200").meshgen
    m = np.load(os.path.join(HERE, "mesh_default.npz"))
    x0 = mg.place_body(m["vertices"], lowest_y=1.0)
    e, t = m["edges"], m["tets"]
    ref = po.Oracle(po.Params.default(substeps=10), x0, e, t, kind="reference")
    if seed:
        rng = np.random.RandomState(seed)
        ref.permute_constraints(rng.permutation(len(e)).astype(np.uint32), rng.permutation(len(t)).astype(np.uint32))
    rows, done = [], 0
    for fr in WINDOW:
        ref.step(1 / 60, fr - done)
        done = fr
        r = po.residuals(ref.positions(), x0, e, t)
        assert r["finite"]
        rows.append([r[k] for k in KEYS])
    return np.array(rows)


if __name__ == "__main__":
    with ProcessPoolExecutor(6) as ex:
        res = list(ex.map(run, list(range(12))))
    res = np.stack(res)                      # [order, frame, metric]
    np.savez_compressed(os.path.join(HERE, "ref_config1_p3_window.npz"), window=np.array(WINDOW),
                        keys=np.array(KEYS), residuals=res)
    print("mean over window per order:\n", res.mean(1))
    print("max over window per order:\n", res.max(1))
