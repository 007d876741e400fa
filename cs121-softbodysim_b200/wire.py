"""PBD1 wire protocol, client side: what the Unity client speaks to PBDServer.

Mirror of ``PBDRemoteWorld.cs`` (``/root/reference/Assets/Scripts/Softbody/PBDRemoteWorld.cs``):
``SendInit`` :278-349 (payload layout), the STEP / POSITIONS exchange :201-246 (strictly one STEP in
flight; the reply must be ``type == MSG_POSITIONS`` and ``size == 12 V`` exactly, :228-231) and the
best-effort SHUTDOWN :253-275.  Used by the wire tests and by headless runs against
``cs121-softbodysim_b200/pbd_server`` (csrc/pbd_server.cpp) or the reference's own server.
"""
from __future__ import annotations

import socket
import struct

import numpy as np

MAGIC = 0x31444250            # 'PBD1' little-endian, PBDServer.h:47
MSG_INIT, MSG_STEP, MSG_POSITIONS, MSG_SHUTDOWN = 1, 2, 3, 4
HEADER = struct.Struct("<III")


class ProtocolError(RuntimeError):
    pass


def pack_message(msg_type: int, payload: bytes = b"") -> bytes:
    return HEADER.pack(MAGIC, msg_type, len(payload)) + payload


class PBD1Client:
    def __init__(self, host: str = "127.0.0.1", port: int = 7777, timeout: float | None = 60.0):
        self.sock = socket.create_connection((host, port), timeout=timeout)
        self.sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)      # NoDelay = true, PBDRemoteWorld.cs:193
        self.V = 0

    def close(self):
        if self.sock is not None:
            try:
                self.sock.close()
            finally:
                self.sock = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def send_raw(self, data: bytes):
        self.sock.sendall(data)

    def recv_exact(self, n: int) -> bytes:
        buf = bytearray()
        while len(buf) < n:
            chunk = self.sock.recv(n - len(buf))
            if not chunk:
                raise ConnectionError(f"server closed the connection after {len(buf)} of {n} bytes")
            buf += chunk
        return bytes(buf)

    def init(self, init_payload: bytes, V: int):
        """MSG_INIT (no reply).  ``init_payload`` as built by ``capi.pack_init_payload``."""
        self.V = V
        self.send_raw(pack_message(MSG_INIT, init_payload))

    def step(self, dt: float) -> np.ndarray:
        """MSG_STEP(dt) -> MSG_POSITIONS: float32 [V,3].  Raises ProtocolError on anything the Unity
        client would disconnect on."""
        self.send_raw(pack_message(MSG_STEP, struct.pack("<f", dt)))
        magic, typ, size = HEADER.unpack(self.recv_exact(HEADER.size))
        if magic != MAGIC or typ != MSG_POSITIONS or size != 12 * self.V:
            raise ProtocolError(f"bad reply: magic={magic:#x} type={typ} size={size}, expected MSG_POSITIONS of {12 * self.V} bytes")
        return np.frombuffer(self.recv_exact(size), dtype="<f4").reshape(-1, 3).copy()

    def shutdown(self):
        try:
            self.send_raw(pack_message(MSG_SHUTDOWN))
        except OSError:
            pass
