#!/usr/bin/env python
"""bench.py -- the PBDServer substep hot path on B200 vs. the reference CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload headline|config2|config1|batch4096] [--backend auto|stream|tile]

A "step" is one frame = one ``pbd_step(dt=1/60)`` = `substeps` XPBD substeps of the workload.
Default workload = BASELINE.json configs[2], the configuration the metric is quoted on
("1M tets"): Kuhn n=56 grid (V=185,193 E=1,257,704 T=1,053,696), 20 substeps x 6 iterations,
edge + volume constraints, ground plane.  With --gpus N (launched by torchrun, one rank per GPU)
every rank steps its own body of that size -- independent bodies, no data-path collective,
"scaling": "weak" -- and `value` = substeps of all ranks / max-over-ranks device time.

Printed (rank 0, ONE JSON line): metric/value/unit per BASELINE.json, `roofline` (algorithmic
bytes per frame / CUDA-event frame time, against MEASURED_PEAKS.json), `cpu_baseline` (the
reference's own CPU path timed on this box's host cores on a bounded sample), `e2e` (the same
metric through the reference-facing stepper API with HOST buffers: step + positions D2H into
pinned memory inside the timed region), `gpu_launches`, `clocks`.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "pbd_substeps_per_s"
UNIT = "substeps/s"
DT = 1.0 / 60.0

WORKLOADS = {
    # name: (kuhn n or asset, substeps, iterations, description)
    "headline": dict(kuhn=56, substeps=20, iterations=6,
                     desc="BASELINE configs[2]: synthetic 1M-tet Kuhn cube grid n=56 (V=185193 E=1257704 T=1053696), "
                          "20 substeps/frame x 6 iterations, edge+volume constraints, ground plane"),
    "config2": dict(kuhn=26, substeps=10, iterations=6,
                    desc="BASELINE configs[1]: synthetic 100k-tet Kuhn cube n=26 (V=19683 E=129194 T=105456), "
                         "10 substeps x 6 iterations, ground plane"),
    "config1": dict(asset="default", substeps=10, iterations=6,
                    desc="BASELINE configs[0]: default Assets/SoftBody tet mesh (V=8613 E=41488 T=26070), 10 substeps x 6 iterations"),
    "small": dict(kuhn=10, substeps=10, iterations=6, desc="one 6k-tet body (config-4 body), smoke-sized"),
    "big8m": dict(kuhn=112, substeps=10, iterations=6,
                  desc="scaling probe: 8.4M-tet Kuhn cube n=112 (V=1442897 E=9947504 T=8429568), 10 substeps x 6 iterations, 1 GPU"),
    "big32m": dict(kuhn=175, substeps=10, iterations=6,
                   desc="BASELINE configs[4] mesh on ONE GPU: 32M-tet Kuhn cube n=175 (V=5451776 E=37791775 T=32156250), "
                        "10 substeps x 6 iterations (no domain decomposition)"),
    "batch4096": dict(kuhn=10, bodies=4096, substeps=10, iterations=6,
                      desc="BASELINE configs[3]: batch of 4096 independent 6,000-tet bodies (Kuhn n=10, V=1331 E=7930 "
                           "T=6000 each, per-body rotation), 10 substeps x 6 iterations; bodies sharded across the GPUs"),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_profile(workload: str, backend: str) -> dict:
    """Per-launch figures of the frame kernel from the committed `ncu --set full` capture of this workload
    (profiles/traffic.json: dram_bytes = dram__bytes_read.sum + dram__bytes_write.sum, warp_instructions =
    smsp__inst_executed.sum, smem_wavefronts = l1tex__data_pipe_lsu_wavefronts_mem_shared.sum); {} if never captured."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            v = json.load(f).get(f"{workload}:{backend}")
    except Exception:
        return {}
    if v is None:
        return {}
    return v if isinstance(v, dict) else {"dram_bytes": v}


def ncu_traffic(workload: str, backend: str):
    return ncu_profile(workload, backend).get("dram_bytes")


def make_workload(name, mg):
    import numpy as np
    w = WORKLOADS[name]
    if "kuhn" in w:
        x0, tets, edges = mg.kuhn_grid(w["kuhn"])
    else:
        m = np.load(os.path.join(ROOT, "tests", "golden", f"mesh_{w['asset']}.npz"))
        x0, tets, edges = mg.place_body(m["vertices"], lowest_y=1.0), m["tets"], m["edges"]
    return x0, edges, tets, w


class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power))
        return out


def cpu_reference_run(x0, edges, tets, w, sample_substeps, frames, warm, threads):
    """Time the reference's own CPU path (oracle/_ref = unmodified Sim.cpp; else the C port) on a
    bounded sample: `frames` steps of `sample_substeps` substeps at the workload's substep dt."""
    from oracle import pyoracle as po
    kind = "reference" if po.have("reference") else "port"
    if kind == "port":
        po.build()
    prm = po.Params.default(substeps=sample_substeps, iterations=w["iterations"])
    dt = DT * sample_substeps / w["substeps"]            # same substep dt as the full workload
    ora = po.Oracle(prm, x0, edges, tets, kind=kind, threads=threads if kind == "reference" else 0)
    for _ in range(warm):
        ora.step(dt)
    t0 = time.perf_counter()
    for _ in range(frames):
        ora.step(dt)
    sec = time.perf_counter() - t0
    st = ora.stats()
    ora.close()
    return dict(kind=kind, seconds=sec, substeps=frames * sample_substeps, stats=st, name=ora.name())


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    pkg = importlib.import_module("cs121-softbodysim_b200")
    x0, edges, tets, w = make_workload(args.workload, pkg.meshgen)
    ncpu = os.cpu_count() or 1
    # each step = a bounded sample: 2 substeps of the full mesh (~0.9 s at 1M tets)
    sample = 2 if len(tets) > 200000 else w["substeps"]
    r = cpu_reference_run(x0, edges, tets, w, sample, args.steps, args.warmup, threads=ncpu)
    val = r["substeps"] / r["seconds"]
    unit = UNIT
    if "bodies" in w:
        # batch workload: the reference steps ONE body per process; a substep of the whole batch costs
        # `bodies` body-substeps on one core (bodies are independent, so P cores would give P x this)
        val, unit = val / w["bodies"], "batch-" + UNIT
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "iterations": w["iterations"], "substeps_per_frame": w["substeps"],
                   "step": f"{sample} substeps of the full mesh (bounded sample of one {w['substeps']}-substep frame)"},
        "tet_constraints_per_s": val * len(tets) * w["iterations"],
        "cpu_baseline": {"value": val, "unit": unit, "cores": 1, "kind": r["kind"],
                         "threads_offered": ncpu,
                         "sample": f"{args.steps} x {sample} substeps after {args.warmup} warm-up steps; "
                                   f"ParallelStepper(threads={ncpu}) -- its constraint sweeps are serial "
                                   "(Sim.cpp:334-337), so the hot loop uses 1 core"},
        "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


def run_batch(args, ctx=None, emit_line=True):
    """--workload batch4096 (BASELINE configs[3]): 4096 independent 6k-tet bodies, sharded across the
    ranks (contiguous body ranges), one kernel per frame per GPU, no collective on the data
    path.  Total work is fixed -> "scaling": "strong".  A substep here = one substep of ALL bodies.
    `value` = fast arithmetic, the bit-exact mode under `alt` (--arith).  With `ctx` (called from main for
    the N > 1 sub-record) the process group of the caller is used and the line is returned, not printed."""
    if ctx is None:
        import numpy as np
        import torch

        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
        torch.cuda.set_device(local)
        dist = None
        if world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        import __graft_entry__ as ge
        if rank == 0:
            ge.build()
        if dist:
            dist.barrier()
        pkg = ge.package()
    else:
        np, torch, dist, rank, world, local, pkg = (ctx[k] for k in ("np", "torch", "dist", "rank", "world", "local", "pkg"))
    arith = ["fast", "exact"] if args.arith == "both" else [args.arith]
    lines = [_run_batch_mode(args, np, torch, dist, rank, world, local, pkg, m, cpu_baseline=(i == 0 and ctx is None))
             for i, m in enumerate(arith)]
    line = lines[0]
    if rank == 0:
        line["config"]["arithmetic"] = ARITH_DOC[arith[0]]
        if len(lines) > 1:
            o = lines[1]
            line["alt"] = {arith[1]: {"value": o["value"], "unit": o["unit"], "ms_per_step": o["ms_per_step"],
                                      "roofline_frac": o["roofline"]["frac"], "e2e": o["e2e"]["value"],
                                      "arithmetic": ARITH_DOC[arith[1]], "sane": o["sane"]}}
        if emit_line:
            emit(line)
    if dist and ctx is None:
        dist.destroy_process_group()
    return line if ctx is not None else 0


def _run_batch_mode(args, np, torch, dist, rank, world, local, pkg, mode, cpu_baseline):
    capi, mg = pkg.capi, pkg.meshgen
    w = WORKLOADS["batch4096"]
    nb, S, I = w["bodies"], w["substeps"], w["iterations"]
    local_xyz, tets, edges = mg.kuhn_grid(w["kuhn"], rot=np.eye(3), lowest_y=None)
    mine = pkg.shard.body_slice(nb, world, rank)
    bodies = []
    for b in mine:       # deterministic per-body orientation and drop height: trajectories differ
        rot = mg.rotation_zx(7.0 * (b % 47), 3.0 * (b % 29))
        bodies.append((mg.place_body(local_xyz, rot=rot, lowest_y=0.25 + 0.001 * (b % 13)), edges, tets))
    V, E, T = len(local_xyz), len(edges), len(tets)
    t0 = time.perf_counter()
    bopt = capi.Options(flags=capi.FLAG_FAST_ARITH if mode == "fast" else 0, lanes_per_tet=args.lanes)
    batch = capi.Batch(capi.SolverParams.default(substeps=S, iterations=I), bodies, device=local, options=bopt)
    init_ms = (time.perf_counter() - t0) * 1e3
    info = batch.info()
    host_pos = torch.empty(3 * V * len(mine), dtype=torch.float32).pin_memory()

    def sync_all():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    if args.preroll:                                  # untimed: past free fall and first contact
        batch.step_async(DT, args.preroll)
        batch.sync()
    for _ in range(args.warmup):
        batch.step_async(DT, 1)
        batch.sync()
    sampler = ClockSampler(local)
    sync_all()
    sampler.start()
    dev_ms = []
    for _ in range(args.steps):                      # working set (~0.9 GB per GPU at N=1) >> L2: no flush needed
        batch.step_async(DT, 1)
        dev_ms.append(batch.sync())
    sync_all()
    total_ms = sum(dev_ms)
    e0 = time.perf_counter()
    for _ in range(args.steps):
        batch.step(DT)
        batch.read_positions(out_ptr=host_pos.data_ptr())
    e2e_s = time.perf_counter() - e0
    clocks = sampler.stop()
    total_ms, e2e_s = pkg.shard.reduce_max([total_ms, e2e_s], dist, "cuda")
    assert sum(pkg.shard.gather_counts(len(mine), dist, "cuda")) == nb
    pos = host_pos.numpy().reshape(-1, 3)
    sane = bool(np.isfinite(pos).all() and pos[:, 1].min() >= -1e-5)
    batch.close()
    if rank != 0:
        return None
    frame_ms = total_ms / args.steps
    value = args.steps * S / (total_ms * 1e-3)                      # substeps of the WHOLE batch per second
    bytes_sub = nb * (104 * V + I * (20 * E + 28 * T + 84 * V))
    achieved = bytes_sub * S / (frame_ms * 1e-3) / 1e9
    peak, peak_src = load_peaks()
    name = "b200-batch-fast" if mode == "fast" else "b200-batch"
    line = {
        "metric": METRIC, "value": value, "unit": "batch-" + UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": frame_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "bodies": nb, "bodies_per_gpu": len(mine), "V": V, "E": E, "T": T,
                   "substeps_per_frame": S, "iterations": I, "dt": DT, "backend": name,
                   "order_mode": "strict", "parallelism": f"{world} GPU(s), bodies sharded, no collective",
                   "preroll_frames": args.preroll,
                   "l2": "not flushed: per-GPU working set %.2f GB >> L2" % (info["device_bytes"] / 1e9)},
        "body_substeps_per_s": value * nb, "tet_constraints_per_s": value * nb * T * I,
        "frame_ms": {"min": min(dev_ms), "median": statistics.median(dev_ms), "max": max(dev_ms), "n": len(dev_ms)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak * world, "unit": "GB/s",
                     "frac": achieved / (peak * world), "traffic": ncu_traffic("batch4096", name),
                     "peak_source": peak_src + (" x %d GPUs" % world if world > 1 else ""),
                     "kernel": "batch_frame_kernel, 1 launch per frame per GPU",
                     "algorithmic_bytes_per_substep": bytes_sub},
        "e2e": {"value": args.steps * S / e2e_s, "unit": "batch-" + UNIT, "h2d_bytes_per_step": 52,
                "d2h_bytes_per_step": 12 * V * len(mine), "ms_per_step": 1e3 * e2e_s / args.steps,
                "api": "pbd_batch_step + pbd_batch_read_positions -> pinned host buffer"},
        "gpu_launches": args.steps * world, "clocks": clocks, "init_ms": init_ms, "plan_ms": info["plan_ms"],
        "schedule": {k: info[k] for k in ("edge_colors", "tet_colors", "tiles", "grid_blocks", "block_threads", "lanes_per_tet")},
        "sane": sane,
    }
    if cpu_baseline and not args.no_cpu_baseline and world == 1:
        r = cpu_reference_run(bodies[0][0], edges, tets, w, S, 8, 1, threads=0)   # 8 frames of ONE body
        per_body = r["substeps"] / r["seconds"]
        line["cpu_baseline"] = {"value": per_body / nb, "unit": "batch-" + UNIT, "cores": 1, "kind": r["kind"],
                                "sample": f"8 frames x {S} substeps of ONE body on 1 of {os.cpu_count()} host cores "
                                          f"({per_body:.1f} body-substeps/s), divided by {nb} bodies"}
    return line


_JSON_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner there)
    must not interleave: everything written to fd 1 from here on goes to stderr, and emit() writes the
    result line to the real stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="headline", choices=sorted(WORKLOADS))
    ap.add_argument("--backend", default="auto", choices=["auto", "stream", "tile"])
    ap.add_argument("--order", default="interleaved", choices=["strict", "interleaved", "riding"],
                    help="interleaved (default): a tile visit projects its edges then its tets, bit-exact vs the "
                         "oracle's sequence sweep; strict: all edges then all tets per iteration, bit-exact vs the "
                         "unmodified reference run on the permuted arrays")
    ap.add_argument("--block-threads", type=int, default=0)
    ap.add_argument("--tile-vertices", type=int, default=0)
    ap.add_argument("--lanes", type=int, default=0, help="tile backend: lanes per tet (0 = auto, 1, 2, 4)")
    ap.add_argument("--partitions", type=int, default=0, help="tile backend: shifted partitions (0 = auto)")
    ap.add_argument("--tiles-per-sm", type=int, default=0, help="tile backend: resident tiles (CTAs) per SM (0 = auto)")
    ap.add_argument("--shard", action="store_true",
                    help="with --gpus N > 1: ONE body of the workload spread over the N GPUs (tiles of other ranks' "
                         "vertices are read/written in place over NVLink; strong scaling) instead of one body per GPU")
    ap.add_argument("--fast", action="store_true", help="shorthand for --arith fast")
    ap.add_argument("--tagged", action="store_true", help="(default now; kept for older scripts)")
    ap.add_argument("--arith", default="both", choices=["both", "fast", "exact"],
                    help="both (default): `value` = fast arithmetic (tolerance-validated), the bit-exact mode under `alt`")
    ap.add_argument("--no-tagged", action="store_true", help="release/acquire done counters instead of the tagged hand-over")
    ap.add_argument("--preroll", type=int, default=30, help="untimed frames before warm-up (past first ground contact)")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--no-splits", action="store_true", help="N > 1: skip the batch4096 / sharded-body sub-records")
    ap.add_argument("--no-extra", action="store_true", help="N = 1: skip the config1 / config2 sub-records")
    ap.add_argument("--split-body", default="big8m", choices=["big8m", "big32m"],
                    help="N > 1: the body of the sharded sub-record (big32m = BASELINE configs[4]; planning it takes minutes)")
    ap.add_argument("--plan-sms", type=int, default=0, help="plan for this many SMs (0 = the device's SM count x ranks of a sharded body)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed frames")
    args = ap.parse_args()
    if args.backend == "stream":
        args.order = "strict"          # the per-colour stream backend only has the reference's edges-then-tets order
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.fast:
        args.arith = "fast"

    if args.workload == "batch4096" and args.impl != "reference":
        return run_batch(args)
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if dist:
        dist.barrier()
    pkg = ge.package()
    ctx = dict(np=np, torch=torch, dist=dist, rank=rank, world=world, local=local, pkg=pkg)

    sharded = args.shard and world > 1
    arith = ["fast", "exact"] if args.arith == "both" else [args.arith]
    if args.backend == "stream" or sharded:
        arith = ["exact"] if args.arith == "both" else arith[:1]     # stream backend / sharded body: one mode
    results = {}
    for mode in arith:
        results[mode] = measure_body(ctx, args, args.workload, mode, sharded, cpu_baseline=(mode == arith[0]))
    line = None
    if rank == 0:
        line = results[arith[0]]
        line["config"]["arithmetic"] = ARITH_DOC[arith[0]]
        if len(arith) > 1:
            o = results[arith[1]]
            line["alt"] = {arith[1]: {"value": o["value"], "unit": o["unit"], "ms_per_step": o["ms_per_step"],
                                      "roofline_frac": o["roofline"]["frac"], "e2e": o["e2e"]["value"], "backend": o["config"]["backend"],
                                      "frame_ms": o["frame_ms"], "arithmetic": ARITH_DOC[arith[1]], "sane": o["sane"]}}
    # ---- the smaller BASELINE configs next to the headline (N = 1): configs[0] (default mesh) and configs[1] (100k tets)
    if world == 1 and args.workload == "headline" and not args.no_extra and args.backend != "stream":
        extra = {}
        for wl in ("config1", "config2"):
            try:
                r = measure_body(ctx, args, wl, arith[0], False, cpu_baseline=False, steps=max(3, args.steps // 2), sustained=False)
                extra[wl] = {"value": r["value"], "unit": r["unit"], "ms_per_step": r["ms_per_step"], "roofline_frac": r["roofline"]["frac"],
                             "e2e": r["e2e"]["value"], "backend": r["config"]["backend"], "workload": r["config"]["workload"],
                             "schedule": r["schedule"], "sane": r["sane"], "arithmetic": arith[0]}
            except Exception as e:
                extra[wl] = {"error": repr(e)}
        line["extra"] = extra
    # ---- the two splits BASELINE.json's north_star names, measured in the same launch when N > 1
    if world > 1 and not sharded and args.workload == "headline" and not args.no_splits:
        sub = {}
        try:
            sub["batch4096"] = run_batch(args, ctx=ctx, emit_line=False)
        except Exception as e:                                        # a sub-record must never take the headline line down
            sub["batch4096"] = {"error": repr(e)}
        try:
            sub[args.split_body] = measure_body(ctx, args, args.split_body, "fast" if args.arith != "exact" else "exact", True,
                                                cpu_baseline=False, steps=max(3, args.steps // 2), sustained=False, check_small=True)
        except Exception as e:
            sub[args.split_body] = {"error": repr(e)}
        if rank == 0:
            line["splits"] = sub
    if rank == 0:
        emit(line)
    if dist:
        dist.destroy_process_group()
    return 0


ARITH_DOC = {
    "fast": "PBD_FLAG_FAST_ARITH: FFMA + SFU rcp/rsqrt forms of the projections, validated by tolerance (P2 <= 1e-4 rel. RMS after 10 "
            "frames, P3 residuals after 1000 frames; tests/test_parity_gpu.py::test_fast_arith_*), not bit for bit; with zero volume "
            "compliance (alpha = 0, the reference default) the tet multipliers never enter a correction and are not carried "
            "(PBD_ARRAY_TET_LAMBDA stays as it was; positions bit-identical to carrying them, test_fast_resident_blocks_and_inert_multipliers)",
    "exact": "IEEE binary32 without FMA in the reference's evaluation order: BIT-EXACT against the oracle replaying the disclosed order",
}


def measure_body(ctx, args, workload, mode, sharded, cpu_baseline, steps=None, sustained=True, check_small=False):
    """One body of `workload` per rank (or ONE body across all ranks when `sharded`): device-timed K frames,
    the end-to-end leg through the stepper API, a sustained sample; returns the JSON line (rank 0) or None."""
    np, torch, dist, rank, world, local, pkg = (ctx[k] for k in ("np", "torch", "dist", "rank", "world", "local", "pkg"))
    capi, mg = pkg.capi, pkg.meshgen
    steps = steps or args.steps
    x0, edges, tets, w = make_workload(workload, mg)
    V, E, T = len(x0), len(edges), len(tets)
    S, I = w["substeps"], w["iterations"]
    prm = capi.SolverParams.default(substeps=S, iterations=I)
    tagged = (not args.no_tagged) and args.backend != "stream"
    flags = (capi.FLAG_FAST_ARITH if mode == "fast" else 0) | (capi.FLAG_TAGGED_HANDOVER if tagged else 0)
    opt = capi.Options(backend={"auto": 0, "stream": 1, "tile": 2}[args.backend],
                       order_mode={"strict": 0, "interleaved": 1, "riding": 2}[args.order],
                       block_threads=args.block_threads, tile_vertices=args.tile_vertices,
                       lanes_per_tet=args.lanes, partitions=args.partitions, tiles_per_sm=args.tiles_per_sm, flags=flags,
                       plan_sms=args.plan_sms)

    t0 = time.perf_counter()
    check = None
    if sharded:
        if check_small:
            check = shard_identity_check(ctx, opt)
        sb = capi.ShardedBody(prm, x0, edges, tets, rank, world, dist, device=local, options=opt)
        body, stepper, state = sb.body, None, None
    else:
        stepper = capi.CudaStepper(device=local, options=opt)
        state = capi.PBDState(prm, x0, edges, tets)
        body = stepper._bind(state)                  # MSG_INIT: plan + upload (not in the timed region)
    init_ms = (time.perf_counter() - t0) * 1e3
    info = body.info()

    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    host_pos = torch.empty(3 * V, dtype=torch.float32).pin_memory()

    def sync_all():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(v):
        if not dist:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- pre-roll (untimed): past free fall and first contact, so that every timed frame is steady-state
    # contact (the body starts 0.25 above the ground; the exact arithmetic has a fast path for exactly
    # satisfied constraints that only free fall takes)
    if args.preroll:
        body.step_async(DT, args.preroll)
        body.sync()
    for _ in range(args.warmup):
        body.step_async(DT, 1)
        body.sync()
    # ---- device-timed region: K frames, inputs resident in HBM, CUDA events on the launching stream
    sampler = ClockSampler(local)
    sync_all()
    sampler.start()
    dev_ms = []
    wall0 = time.perf_counter()
    for _ in range(steps):
        if flush is not None:
            flush.zero_()                            # evict L2 (126 MB) between timed frames; not timed
            torch.cuda.synchronize()
        body.step_async(DT, 1)
        dev_ms.append(body.sync())                   # ms between the library's events around the frame
    wall_ms = (time.perf_counter() - wall0) * 1e3
    sync_all()
    total_ms = max_over_ranks(sum(dev_ms))

    # ---- end-to-end through the stepper API with host buffers (step + pack D2H into pinned memory)
    stats = capi.StepStats()

    def e2e_step():
        if sharded:                                  # every rank reads back its own part of the body
            body.step_async(DT, 1)
            body.sync()
        else:
            stepper.step(state, DT, stats)
        body.read_positions(out_ptr=host_pos.data_ptr())

    for _ in range(2):
        e2e_step()
    sync_all()
    e0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    e2e_s = max_over_ranks(time.perf_counter() - e0)
    clocks = sampler.stop()

    # ---- sustained sample: >= 2 s of back-to-back frames (no flush, no host work in between), own clocks record
    sus = None
    if sustained and not args.no_sustained:
        per = max(1, int(0.25 / max(1e-4, (total_ms / steps) * 1e-3)))       # ~0.25 s of frames per sync
        s2 = ClockSampler(local)
        sync_all()
        s2.start()
        ms_sum, frames, t_s = 0.0, 0, time.perf_counter()
        while time.perf_counter() - t_s < 2.0:
            body.step_async(DT, per)
            ms_sum += body.sync()
            frames += per
        sus_clk = s2.stop()
        ms_sum = max_over_ranks(ms_sum)
        sus = {"seconds": ms_sum * 1e-3, "frames": frames, "ms_per_step": ms_sum / frames,
               "value": (1 if sharded else world) * frames * S / (ms_sum * 1e-3), "clocks": sus_clk}
    body.read_positions(out_ptr=host_pos.data_ptr())
    pos = host_pos.numpy().reshape(-1, 3)
    if sharded:
        sane = bool(np.isfinite(pos[sb.owner == rank]).all() and pos[sb.owner == rank][:, 1].min() >= -1e-5)
    else:
        sane = bool(np.isfinite(pos).all() and pos[:, 1].min() >= -1e-5)
    name = body.name()
    if stepper:
        stepper.close()
    else:
        sb.close()
    if rank != 0:
        return None

    bodies = 1 if sharded else world
    value = bodies * steps * S / (total_ms * 1e-3)
    bytes_sub = info["algorithmic_bytes_per_substep"]
    frame_ms = total_ms / steps
    achieved = bodies * bytes_sub * S / (frame_ms * 1e-3) / 1e9
    peak, peak_src = load_peaks()
    peak *= world
    prof = ncu_profile(workload, name)
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup,
        "ms_per_step": frame_ms, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "V": V, "E": E, "T": T, "substeps_per_frame": S, "iterations": I,
                   "dt": DT, "backend": name, "order_mode": args.order, "hand_over": "tagged 128-bit {x,y,z,tag} words" if "tagged" in name else "release/acquire done counters",
                   "lanes_per_tet": info.get("lanes_per_tet"), "partitions": info.get("partitions"),
                   "parallelism": ("1 GPU" if world == 1 else
                                   f"ONE body across {world} GPUs: tiles read/write other ranks' vertices in place over NVLink "
                                   "(peer memory, CUDA IPC; system-scope tagged words or release/acquire counters); no NCCL on the data path"
                                   if sharded else f"{world} independent bodies, one per GPU, no collective"),
                   "l2": "not flushed" if flush is None else "flushed between timed frames (256 MiB memset, untimed)",
                   "preroll_frames": args.preroll, "working_set_bytes": info["device_bytes"]},
        "tet_constraints_per_s": value * T * I,
        "substeps_per_s_per_gpu": value / world,
        "frame_ms": {"min": min(dev_ms), "median": statistics.median(dev_ms), "max": max(dev_ms), "n": len(dev_ms),
                     "note": "device ms of every timed frame on this rank (CUDA events on the library's stream)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": prof.get("dram_bytes"), "peak_source": peak_src,
                     "kernel": "whole frame = %d launches of %s" % (info["launches_per_frame"], name),
                     "algorithmic_bytes_per_substep": bytes_sub, "frac_of_nominal_8TBs": achieved / 8000.0},
        "e2e": {"value": bodies * steps * S / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 52,
                "d2h_bytes_per_step": 12 * V, "ms_per_step": 1e3 * e2e_s / steps,
                "api": "CudaStepper.step(state, dt) + pack_positions -> pinned host buffer (C ABI pbd_step + pbd_read_positions)"},
        "gpu_launches": steps * info["launches_per_frame"],
        "clocks": clocks,
        "init_ms": init_ms, "plan_ms": info["plan_ms"], "wall_ms_timed_region": wall_ms,
        "schedule": {k: info[k] for k in ("edge_colors", "tet_colors", "edge_phases", "tet_phases", "tiles", "partitions",
                                          "lanes_per_tet", "launches_per_frame", "grid_blocks", "block_threads",
                                          "gather_wavefronts_permille")},
        "sane": sane,
    }
    sm_count = torch.cuda.get_device_properties(local).multi_processor_count
    per_sm = max(1, -(-info["grid_blocks"] // sm_count))                  # resident CTAs per SM
    sms_used = max(1, info["grid_blocks"] // per_sm)
    if prof.get("warp_instructions"):
        # second roofline: issue slots.  warp-instructions per frame (ncu smsp__inst_executed.sum of the committed capture
        # of this kernel) / (SMs that run CTAs x 4 schedulers x cycles of the frame at the sampled SM clock)
        cyc = frame_ms * 1e-3 * sm_mhz * 1e6
        line["roofline"]["issue_slot_frac"] = prof["warp_instructions"] / (sms_used * 4 * cyc)
        line["roofline"]["issue_slot_source"] = prof.get("source")
    if prof.get("smem_wavefronts"):
        # third: the shared-memory data pipe, one wavefront per cycle and SM (what bounds the colour sweeps)
        cyc = frame_ms * 1e-3 * sm_mhz * 1e6
        line["roofline"]["smem_pipe_frac"] = prof["smem_wavefronts"] / (sms_used * cyc)
        line["roofline"]["sms_used"] = sms_used
    if sus:
        line["sustained"] = sus
    if check is not None:
        line["bit_identity_check"] = check
    if cpu_baseline and not args.no_cpu_baseline and world == 1:
        sample = 2 if T > 200000 else S
        frames = 3 if T > 200000 else max(1, int(2e8 / max(1, (20 * E + 45 * T) * I * sample)))
        r = cpu_reference_run(x0, edges, tets, w, sample, frames, 1, threads=0)
        cv = r["substeps"] / r["seconds"]
        line["cpu_baseline"] = {"value": cv, "unit": UNIT, "cores": 1, "kind": r["kind"],
                                "sample": f"{frames} x {sample} substeps (same mesh, same substep dt) after 1 warm-up; "
                                          f"SerialStepper on 1 of {os.cpu_count()} host cores",
                                "solve_fraction": r["stats"]["solveMs"] / max(r["stats"]["totalMs"], 1e-9)}
    return line


def shard_identity_check(ctx, opt):
    """Spot check run before a sharded measurement: a small body (Kuhn n=16) stepped 3 frames as ONE body
    across the ranks must be bit-identical to rank 0's single-GPU run of the same plan (plan_sms)."""
    np, torch, dist, rank, world, local, pkg = (ctx[k] for k in ("np", "torch", "dist", "rank", "world", "local", "pkg"))
    capi, mg = pkg.capi, pkg.meshgen
    import ctypes as C
    x0, tets, edges = mg.kuhn_grid(16)
    prm = capi.SolverParams.default(substeps=4)
    o = capi.Options()
    C.memmove(C.byref(o), C.byref(opt), C.sizeof(opt))
    o.tile_vertices, o.plan_sms = 160, 16 * world
    sb = capi.ShardedBody(prm, x0, edges, tets, rank, world, dist, device=local, options=o)
    for _ in range(3):
        sb.step_async(1 / 60)
        sb.sync()
    got = sb.read_positions()
    sb.close()
    same = None
    if rank == 0:
        o1 = capi.Options()
        C.memmove(C.byref(o1), C.byref(o), C.sizeof(o))
        o1.shard_world, o1.shard_rank = 0, 0
        with capi.Body(prm, x0, edges, tets, device=local, options=o1) as b:
            for _ in range(3):
                b.step(1 / 60)
            same = bool(np.array_equal(b.read_positions(), got))
    if dist:
        dist.barrier()
    return {"body": "Kuhn n=16, 3 frames, ONE body across the ranks vs rank 0 alone on the same plan", "bit_identical": same}


if __name__ == "__main__":
    sys.exit(main())
