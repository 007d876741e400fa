"""One body across TWO GPUs (BASELINE config 5 mechanics), single-process variant: needs a box with
>= 2 GPUs (`gpurun --gpus 2`), skipped elsewhere.  The sharded run must be BIT-IDENTICAL to a
single-GPU run of the same schedule (Options.plan_sms) and to the oracle replaying that schedule."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _need_two(capi):
    if capi.device_count() < 2:
        pytest.skip("needs two GPUs")


@pytest.mark.parametrize("flags", [0, 4], ids=["counters", "tagged"])
@pytest.mark.parametrize("mesh,order", [("kuhn12", "interleaved"), ("kuhn12", "strict"), ("icosphere001", "interleaved")])
def test_two_gpu_shard_bit_exact_vs_single_gpu_and_oracle(mesh, order, flags, capi, po, meshgen, golden):
    _need_two(capi)
    if mesh.startswith("kuhn"):
        x0, tets, edges = meshgen.kuhn_grid(int(mesh[4:]))
    else:
        m = golden(f"mesh_{mesh}.npz")
        x0, edges, tets = meshgen.place_body(m["vertices"], lowest_y=1.0), m["edges"], m["tets"]
    om = capi.ORDER_INTERLEAVED if order == "interleaved" else capi.ORDER_STRICT
    # small tiles so that both GPUs own several tiles of every phase and many tiles straddle the cut
    opt = capi.Options(backend=capi.BACKEND_TILE, order_mode=om, tile_vertices=0 if flags & 4 else 150, plan_sms=8, flags=flags)   # (the tagged hand-over needs the regular partitions: every phase covers every vertex)
    prm = capi.SolverParams.default(substeps=4)
    single = capi.Body(prm, x0, edges, tets, device=0, options=opt)
    shards = capi.sharded_bodies_one_process(prm, x0, edges, tets, devices=[0, 1], options=opt)
    if flags & 4 and mesh.startswith("kuhn"):
        assert "tagged" in shards[0].name() and "tagged" in single.name()
    owner = capi.shard_owner(shards[0])
    assert set(np.unique(owner)) == {0, 1}
    for a, b in zip(single.schedule_order(), shards[1].schedule_order()):
        assert np.array_equal(a, b)                                   # every rank built the same schedule
    ora = po.Oracle(po.Params.default(substeps=4), x0, edges, tets, kind="port")
    ora.permute_constraints(*single.schedule_order())
    seq = single.schedule_sequence()
    done = 0
    for fr in (1, 6, 20):
        n = fr - done
        single.step_async(1 / 60, n)
        for s in shards:                                              # launch every rank before any sync
            s.step_async(1 / 60, n)
        single.sync()
        for s in shards:
            s.sync()
        for _ in range(n):
            ora.step_sequence(1 / 60, seq)
        done = fr
        want = single.read_positions()
        assert np.array_equal(want, ora.positions()), f"single-GPU vs oracle, frame {fr}"
        got = np.where((owner == 0)[:, None], shards[0].read_positions(), shards[1].read_positions())
        assert np.array_equal(got, want), f"2-GPU shard vs single GPU, frame {fr}: max |d| = {np.abs(got - want).max():.3e}"
    for b in [single] + shards:
        b.close()
