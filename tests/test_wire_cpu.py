"""PBD1 wire server (cs121-softbodysim_b200/pbd_server, csrc/pbd_server.cpp): framing and error behaviour
that needs no GPU.  Reference behaviour being mirrored: CProgram/src/Server.cpp:20-149 (comm_loop),
Net.cpp:57-102; the client side is cs121-softbodysim_b200/wire.py (PBDRemoteWorld.cs:187-349)."""
import os
import re
import socket
import struct
import subprocess
import time

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class Server:
    """pbd_server (asks the OS for a port with --port 0 and prints it) or, with `exe`, the reference's
    server loop, which prints the port it was GIVEN (Net.cpp:88): that one gets a port chosen here."""

    def __init__(self, pkg, *extra, exe=None):
        pkg.build.build()
        port = "0" if exe is None else str(_free_port())
        exe = exe or pkg.build.SERVER
        self.proc = subprocess.Popen([exe, "--port", port, *extra], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        self.port = None
        self.lines = []
        t0 = time.time()
        while time.time() - t0 < 30:
            line = self.proc.stdout.readline()
            if not line:
                break
            self.lines.append(line)
            m = re.search(r"Listening on port (\d+)", line)
            if m:
                self.port = int(m.group(1))
                break
        assert self.port, "".join(self.lines)

    def close(self):
        if self.proc.poll() is None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        try:
            self.lines += self.proc.stdout.readlines()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def _closed(sock, wait=5.0):
    """True if the peer closed the connection within `wait` seconds (recv returns b'' or reset)."""
    sock.settimeout(wait)
    try:
        return sock.recv(1) == b""
    except socket.timeout:
        return False
    except OSError:
        return True


def test_server_binary_builds_and_prints_usage(pkg):
    pkg.build.build()
    assert os.access(pkg.build.SERVER, os.X_OK)
    out = subprocess.run([pkg.build.SERVER, "--help"], capture_output=True, text=True, timeout=10)
    assert out.returncode == 0 and "--port" in out.stdout
    bad = subprocess.run([pkg.build.SERVER, "--mode", "serial"], capture_output=True, text=True, timeout=10)
    assert bad.returncode == 1                                        # like main.cpp:44-52: unknown mode -> usage, exit 1


def test_step_before_init_is_ignored_and_bad_frames_disconnect(pkg):
    wire = pkg.wire
    with Server(pkg) as srv:
        # STEP before INIT: silently ignored, NO reply, connection stays open (Server.cpp:122)
        with wire.PBD1Client(port=srv.port, timeout=5) as c:
            c.send_raw(wire.pack_message(wire.MSG_STEP, struct.pack("<f", 1 / 60)))
            assert not _closed(c.sock, wait=0.5)
            # a STEP shorter than its float: disconnect (Server.cpp:116)
            c.send_raw(wire.pack_message(wire.MSG_STEP, b"\x00\x00"))
            assert _closed(c.sock)
        # bad magic: disconnect (Server.cpp:4-8, 25); the server keeps accepting clients
        with wire.PBD1Client(port=srv.port, timeout=5) as c:
            c.send_raw(struct.pack("<III", 0x12345678, wire.MSG_STEP, 4) + b"\0\0\0\0")
            assert _closed(c.sock)
        # unknown message type: disconnect (Server.cpp:141-143)
        with wire.PBD1Client(port=srv.port, timeout=5) as c:
            c.send_raw(wire.pack_message(77))
            assert _closed(c.sock)
        # MSG_SHUTDOWN: the server process exits (Server.cpp:138-140, main.cpp:92-97)
        with wire.PBD1Client(port=srv.port, timeout=5) as c:
            c.shutdown()
            assert _closed(c.sock)
        assert srv.proc.wait(timeout=10) == 0
    assert any("Shutdown" in l for l in srv.lines + srv.proc.stdout.readlines())


def test_malformed_init_is_refused_not_read_out_of_bounds(pkg, capi, meshgen):
    """The reference trusts V/E/T (Server.cpp:35-70 reads past the payload); here a MSG_INIT shorter than
    its own counts demand, or with an index >= V, closes the session -- and the server survives."""
    wire = pkg.wire
    x0, tets, edges = meshgen.kuhn_grid(3)
    good = capi.pack_init_payload(capi.SolverParams.default(), x0, edges, tets)
    with Server(pkg) as srv:
        with wire.PBD1Client(port=srv.port, timeout=5) as c:          # truncated payload
            c.init(good[:-8], len(x0))
            assert _closed(c.sock)
        with wire.PBD1Client(port=srv.port, timeout=5) as c:          # counts far larger than the payload
            lie = struct.pack("<III", 10**9, 10**9, 10**9) + good[12:]
            c.init(lie, 10**9)
            assert _closed(c.sock)
        with wire.PBD1Client(port=srv.port, timeout=5) as c:          # shorter than the 64-byte fixed part
            c.init(good[:40], len(x0))
            assert _closed(c.sock)
        bad_t = tets.copy(); bad_t[0, 0] = len(x0)
        with wire.PBD1Client(port=srv.port, timeout=5) as c:          # tet index == V
            c.init(capi.pack_init_payload(capi.SolverParams.default(), x0, edges, bad_t), len(x0))
            assert _closed(c.sock)
        nan_x = x0.copy(); nan_x[5, 1] = np.nan
        with wire.PBD1Client(port=srv.port, timeout=5) as c:          # NaN position (undefined sort order in the planner)
            c.init(capi.pack_init_payload(capi.SolverParams.default(), nan_x, edges, tets), len(x0))
            assert _closed(c.sock)
        assert srv.proc.poll() is None                                # still serving
        if capi.device_count() == 0:
            with wire.PBD1Client(port=srv.port, timeout=20) as c:     # a valid INIT without a GPU: refused (no CPU fallback)
                c.init(good, len(x0))
                assert _closed(c.sock, wait=15)
            assert srv.proc.poll() is None
    assert any("Init refused" in l for l in srv.lines)
