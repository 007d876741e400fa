"""The outer (process / wire) seam on the GPU: a Python client replays what PBDRemoteWorld.cs sends
(SendInit :278-349, one STEP in flight :201-246) against

  * cs121-softbodysim_b200/pbd_server  -- this repository's PBD1 server (csrc/pbd_server.cpp), and
  * integration/_build/PBDServer_gpu   -- the REFERENCE's own comm_loop / sim_thread_fn / send_positions
    (compiled in place from /root/reference, unmodified) with integration/CudaStepper behind IStepper,

and checks that MSG_POSITIONS carries exactly 12 V bytes that are bit-equal to the in-process
pbd_step run, and that a second MSG_INIT with the same V/E/T restarts the body (re-INIT, Server.cpp:106-110)."""
import os
import time

import numpy as np
import pytest

from test_wire_cpu import Server

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SERVER = os.path.join(ROOT, "integration", "_build", "PBDServer_gpu")


def _mesh(golden, meshgen):
    m = golden("mesh_default.npz")
    return meshgen.place_body(m["vertices"], lowest_y=1.0), m["edges"], m["tets"]


def _in_process(capi, prm, x0, edges, tets, pinned, frames, **optkw):
    with capi.Body(prm, x0, edges, tets, pinned=pinned, device=0, options=capi.Options(**optkw)) as b:
        out = []
        for _ in range(frames):
            b.step(1 / 60)
            out.append(b.read_positions())
        return out


@pytest.mark.parametrize("order,extra", [("strict", []), ("riding", ["--order", "riding"])])
def test_pbd_server_positions_bit_equal_to_in_process_and_reinit(order, extra, pkg, capi, meshgen, golden):
    wire = pkg.wire
    x0, edges, tets = _mesh(golden, meshgen)
    pinned = np.argsort(-x0[:, 1])[:40].astype(np.uint32)            # top layer, PBDRemoteSoftBody.cs:163-183
    prm = capi.SolverParams.default(substeps=10)                       # BASELINE config 1 parameters
    pay = capi.pack_init_payload(prm, x0, edges, tets, pinned)
    want = _in_process(capi, prm, x0, edges, tets, pinned, 6, backend=capi.BACKEND_TILE,
                       order_mode=capi.ORDER_RIDING if order == "riding" else capi.ORDER_STRICT)
    with Server(pkg, *extra) as srv:
        with wire.PBD1Client(port=srv.port, timeout=120) as c:
            c.init(pay, len(x0))
            got, dts = [], []
            for _ in range(6):
                t0 = time.perf_counter()
                got.append(c.step(1 / 60))                             # raises unless type == POSITIONS and size == 12 V
                dts.append(time.perf_counter() - t0)
            for f, (a, b) in enumerate(zip(got, want)):
                assert a.shape == (len(x0), 3)
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"frame {f}: wire positions differ from pbd_step"
            # no Nagle / delayed-ACK stall: the reference's two-send reply costs ~40 ms per frame on loopback
            assert sorted(dts)[len(dts) // 2] < 0.030, dts
            # second MSG_INIT, same V/E/T: the body restarts from x0
            c.init(pay, len(x0))
            again = [c.step(1 / 60) for _ in range(3)]
            for f in range(3):
                assert np.array_equal(again[f].view(np.uint32), want[f].view(np.uint32)), f"re-INIT frame {f}"
            c.shutdown()
        assert srv.proc.wait(timeout=20) == 0
    assert any("Init received" in l for l in srv.lines + srv.proc.stdout.readlines())


def test_reference_server_with_cuda_stepper_adapter_reinit(pkg, capi, meshgen, golden):
    """The reference's own server loop with integration/CudaStepper.cpp behind IStepper.  INIT -> 5 steps
    -> second INIT with the same V/E/T (which the reference move-assigns into the SAME PBDState object)
    -> 5 steps: the second run must reproduce the first (the adapter's generation stamp detects it),
    and both equal the in-process run."""
    if not os.access(REF_SERVER, os.X_OK):
        pytest.skip("integration/_build/PBDServer_gpu was not built (needs /root/reference at build time)")
    wire = pkg.wire
    x0, edges, tets = _mesh(golden, meshgen)
    prm = capi.SolverParams.default(substeps=10)
    pay = capi.pack_init_payload(prm, x0, edges, tets)
    want = _in_process(capi, prm, x0, edges, tets, None, 5, backend=capi.BACKEND_AUTO, order_mode=capi.ORDER_STRICT)
    with Server(pkg, exe=REF_SERVER) as srv:
        with wire.PBD1Client(port=srv.port, timeout=120) as c:
            c.init(pay, len(x0))
            first = [c.step(1 / 60) for _ in range(5)]
            c.init(pay, len(x0))
            second = [c.step(1 / 60) for _ in range(5)]
            c.shutdown()
        assert srv.proc.wait(timeout=20) == 0
        tail = "".join(srv.lines + srv.proc.stdout.readlines())
    for f in range(5):
        assert np.array_equal(first[f].view(np.uint32), want[f].view(np.uint32)), f"frame {f}: adapter vs pbd_step"
        assert np.array_equal(second[f].view(np.uint32), first[f].view(np.uint32)), f"frame {f}: stale device state after re-INIT"
    assert "binds=2" in tail and "ok=1" in tail, tail
