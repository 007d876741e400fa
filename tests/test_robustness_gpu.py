"""Robustness of the tile backend's cross-CTA synchronisation:
32-bit iteration counters / tags that wrap, and waits that are bounded instead of hanging the GPU."""
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("order,flags", [("interleaved", 0), ("strict", 0), ("interleaved", 4), ("riding", 0), ("riding", 4)])
def test_iteration_counters_wrap_safely(order, flags, capi, po, meshgen, monkeypatch):
    """The per-tile done counters (and the hand-over tags) count iterations across frames and never
    reset.  PBD_DEBUG_ITERBASE starts them 100 iterations below 2^32: 10 frames of 4 x 6 iterations
    cross the wrap; the waits compare by signed distance / equality, so the result stays bit-exact."""
    monkeypatch.setenv("PBD_DEBUG_ITERBASE", str(2**32 - 100))
    x0, tets, edges = meshgen.kuhn_grid(8)
    om = {"strict": capi.ORDER_STRICT, "interleaved": capi.ORDER_INTERLEAVED, "riding": capi.ORDER_RIDING}[order]
    body = capi.Body(capi.SolverParams.default(substeps=4), x0, edges, tets, device=0,
                     options=capi.Options(backend=capi.BACKEND_TILE, order_mode=om, tile_vertices=120, flags=flags))
    ora = po.Oracle(po.Params.default(substeps=4), x0, edges, tets, kind="port")
    ora.permute_constraints(*body.schedule_order())
    seq = body.schedule_sequence()
    for fr in range(10):
        body.step(1 / 60)
        ora.step_sequence(1 / 60, seq)
        assert np.array_equal(body.read_positions(), ora.positions()), f"frame {fr} (wrap happens in frame 4)"
    body.close()


def test_unlaunched_peer_rank_times_out_instead_of_hanging(capi, meshgen, monkeypatch):
    """One body over two ranks (both on device 0 here), only rank 0 launched: its tiles wait for tiles
    rank 1 never runs.  The wait gives up after PBD_SPIN_LIMIT_MS and pbd_sync reports PBD_ERR_CUDA
    -- the GPU is not left spinning until reset."""
    monkeypatch.setenv("PBD_SPIN_LIMIT_MS", "250")
    x0, tets, edges = meshgen.kuhn_grid(8)
    opt = capi.Options(backend=capi.BACKEND_TILE, order_mode=capi.ORDER_INTERLEAVED, tile_vertices=150, plan_sms=8)
    shards = capi.sharded_bodies_one_process(capi.SolverParams.default(substeps=2), x0, edges, tets, devices=[0, 0], options=opt)
    t0 = time.time()
    shards[0].step_async(1 / 60, 1)
    with pytest.raises(capi.PBDError) as e:
        shards[0].sync()
    assert e.value.code == capi.PBD_ERR_CUDA and "waited longer" in str(e.value)
    assert time.time() - t0 < 20.0
    # the device is still usable: a fresh single-GPU body steps fine
    with capi.Body(capi.SolverParams.default(substeps=2), x0, edges, tets, device=0) as b:
        b.step(1 / 60)
        assert np.isfinite(b.read_positions()).all()
    for s in shards:
        s.close()


@pytest.mark.parametrize("flags", [0, 4], ids=["counters", "tagged"])
@pytest.mark.parametrize("mesh,order", [("kuhn12", "interleaved"), ("kuhn12", "strict"), ("icosphere001", "riding")])
def test_two_rank_shard_on_one_gpu_bit_exact(mesh, order, flags, capi, po, meshgen, golden):
    """The sharded-body path (tiles of rank r read / write vertices owned by rank 1-r in place, done
    counters at system scope) exercised on ONE GPU: both ranks' cooperative kernels are small enough to
    be co-resident on device 0 and are launched on their own streams.  Bit-identical to the single-handle
    run of the same schedule and to the oracle (tests/test_shard_gpu.py is the 2-GPU version)."""
    if mesh.startswith("kuhn"):
        x0, tets, edges = meshgen.kuhn_grid(int(mesh[4:]))
    else:
        m = golden(f"mesh_{mesh}.npz")
        x0, edges, tets = meshgen.place_body(m["vertices"], lowest_y=1.0), m["edges"], m["tets"]
    om = {"strict": capi.ORDER_STRICT, "interleaved": capi.ORDER_INTERLEAVED, "riding": capi.ORDER_RIDING}[order]
    opt = capi.Options(backend=capi.BACKEND_TILE, order_mode=om, tile_vertices=0 if flags & 4 else 150, plan_sms=8, flags=flags)   # (the tagged hand-over needs the regular partitions: every phase covers every vertex)
    prm = capi.SolverParams.default(substeps=4)
    single = capi.Body(prm, x0, edges, tets, device=0, options=opt)
    shards = capi.sharded_bodies_one_process(prm, x0, edges, tets, devices=[0, 0], options=opt)
    if flags & 4 and mesh.startswith("kuhn"):
        assert "tagged" in shards[0].name() and "tagged" in single.name()   # (a plan with residual phases falls back to the counters)
    owner = capi.shard_owner(shards[0])
    assert set(np.unique(owner)) == {0, 1}
    ora = po.Oracle(po.Params.default(substeps=4), x0, edges, tets, kind="port")
    ora.permute_constraints(*single.schedule_order())
    seq = single.schedule_sequence()
    for fr in range(6):
        single.step_async(1 / 60, 1)
        for s in shards:
            s.step_async(1 / 60, 1)
        single.sync()
        for s in shards:
            s.sync()
        ora.step_sequence(1 / 60, seq)
        want = single.read_positions()
        assert np.array_equal(want, ora.positions()), f"single handle vs oracle, frame {fr}"
        got = np.where((owner == 0)[:, None], shards[0].read_positions(), shards[1].read_positions())
        assert np.array_equal(got, want), f"2 ranks on one GPU vs single handle, frame {fr}"
    for b in [single] + shards:
        b.close()
