#!/usr/bin/env python3
"""Host only: fingerprints (PBD_PLAN_DEBUG, FNV-1a over every array of the plan) of the tile planner's output for a set
of meshes and options.  `python tools/plan_fingerprints.py > before.txt`, change the planner, run again, diff: how a
planner change that claims to be result-neutral (threading, data structures) is checked before any GPU run."""
import importlib, os, re, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [
    ("kuhn56 headline (interleaved, tagged, fast)", "kuhn:56", dict(backend=2, order_mode=1, flags=3, plan_sms=148)),
    ("kuhn56 exact", "kuhn:56", dict(backend=2, order_mode=1, flags=1, plan_sms=148)),
    ("kuhn26 config 2", "kuhn:26", dict(backend=2, order_mode=1, flags=3, plan_sms=148)),
    ("kuhn26 strict", "kuhn:26", dict(backend=2, order_mode=0, flags=0, plan_sms=148)),
    ("kuhn26 riding", "kuhn:26", dict(backend=2, order_mode=2, flags=3, plan_sms=148)),
    ("kuhn12 small tiles", "kuhn:12", dict(backend=2, order_mode=1, flags=3, tile_vertices=300, plan_sms=148)),
    ("kuhn12 3 partitions", "kuhn:12", dict(backend=2, order_mode=1, flags=0, tile_vertices=150, partitions=3, plan_sms=148)),
    ("kuhn40 sharded x2", "kuhn:40", dict(backend=2, order_mode=1, flags=3, plan_sms=296, shard_world=2, shard_rank=1)),
    ("kuhn80 2.9M tets (no placement search)", "kuhn:80", dict(backend=2, order_mode=1, flags=3, plan_sms=148)),
    ("kuhn20 lanes 4 strict", "kuhn:20", dict(backend=2, order_mode=0, flags=0, lanes_per_tet=4, plan_sms=148)),
]


def child(mesh, opts):
    sys.path.insert(0, ROOT)
    pkg = importlib.import_module("cs121-softbodysim_b200")
    capi, mg = pkg.capi, pkg.meshgen
    kind, n = mesh.split(":")
    x0, tets, edges = mg.kuhn_grid(int(n))
    o = {k: v for k, v in opts.items()}
    fl = o.pop("flags", 0)
    flags = (capi.FLAG_TAGGED_HANDOVER if fl & 1 else 0) | (capi.FLAG_FAST_ARITH if fl & 2 else 0)
    t0 = time.perf_counter()
    p = capi.Plan(x0, edges, tets, options=capi.Options(flags=flags, **o))
    print("PLANMS %.0f" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)
    p.close()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(sys.argv[2], eval(sys.argv[3]))
        sys.exit(0)
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    for name, mesh, opts in CASES:
        if only and only not in name:
            continue
        r = subprocess.run([sys.executable, __file__, "--child", mesh, repr(opts)], capture_output=True, text=True,
                           env=dict(os.environ, PBD_PLAN_DEBUG="1"))
        fp = re.findall(r"fingerprint ([0-9a-f]+)", r.stderr)
        ms = re.findall(r"PLANMS (\d+)", r.stderr)
        print(f"{name:45s} {fp[-1] if fp else 'FAILED: ' + r.stderr[-200:]}", file=sys.stdout, flush=True)
        print(f"   {name}: {ms[-1] if ms else '?'} ms", file=sys.stderr, flush=True)
