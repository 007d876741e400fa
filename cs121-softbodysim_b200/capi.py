"""ctypes binding of ``libpbd_b200.so`` (``include/pbd_b200.h``) and the Python-side mirror of the
reference's stepper seam.

Mirror of the reference (paths relative to /root/reference):

===========================  ==============================================================
here                         reference
===========================  ==============================================================
``SolverParams``             ``struct SolverParams``  CProgram/include/PBDServer.h:147-161
``StepStats``                ``perf::StepStats``      CProgram/include/PBDServer.h:75-81
``PBDState``                 ``struct PBDState`` as filled by the MSG_INIT decode,
                             CProgram/src/Server.cpp:72-104 (host arrays; device state is owned
                             by the stepper, keyed by ``state.generation`` so a second INIT is
                             detected)
``CudaStepper.name()``       ``IStepper::name``            PBDServer.h:263
``CudaStepper.step()``       ``IStepper::step``            PBDServer.h:264 / Sim.cpp:280-305
``CudaStepper.pack_positions()``  ``IStepper::pack_positions``  PBDServer.h:265 / Sim.cpp:307-316
===========================  ==============================================================

Everything computes on the GPU through the C ABI; if the library is missing or no CUDA device
is present the calls raise -- there is no CPU path in this package.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# PBD_B200_LIB: load another build of the same ABI (A/B timing of kernel variants on one box)
LIB_PATH = os.environ.get("PBD_B200_LIB") or os.path.join(HERE, "libpbd_b200.so")

PBD_OK, PBD_ERR_INVALID, PBD_ERR_INDEX, PBD_ERR_NO_DEVICE, PBD_ERR_CUDA, PBD_ERR_OOM, PBD_ERR_UNSUPPORTED = range(7)
BACKEND_AUTO, BACKEND_STREAM, BACKEND_TILE, BACKEND_JACOBI = 0, 1, 2, 3
ORDER_STRICT, ORDER_INTERLEAVED, ORDER_RIDING = 0, 1, 2
FLAG_STAGE_TIMING, FLAG_NO_GRAPH, FLAG_TAGGED_HANDOVER, FLAG_FAST_ARITH = 1, 2, 4, 8
ARRAY_INV_MASS, ARRAY_EDGE_REST, ARRAY_TET_REST, ARRAY_EDGE_LAMBDA, ARRAY_TET_LAMBDA, ARRAY_VELOCITY, ARRAY_XSTAR = range(7)

# every symbol include/pbd_b200.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "pbd_abi_version", "pbd_last_error", "pbd_device_count",
    "pbd_create", "pbd_create_from_init", "pbd_init_payload_size", "pbd_step", "pbd_step_async", "pbd_sync", "pbd_read_positions", "pbd_destroy",
    "pbd_backend_name", "pbd_get_info", "pbd_set_params", "pbd_get_schedule_order",
    "pbd_get_schedule_sequence", "pbd_get_array", "pbd_set_colliders", "pbd_set_surface", "pbd_read_normals",
    "pbd_shard_export", "pbd_shard_attach_ipc", "pbd_shard_attach_local", "pbd_shard_owner",
    "pbd_plan_create", "pbd_plan_get_info", "pbd_plan_get_order", "pbd_plan_get_sequence",
    "pbd_plan_get_edge_slots", "pbd_plan_get_tet_slots", "pbd_plan_destroy",
    "pbd_batch_create", "pbd_batch_step", "pbd_batch_step_async", "pbd_batch_sync",
    "pbd_batch_read_positions", "pbd_batch_get_info", "pbd_batch_get_schedule_order", "pbd_batch_destroy",
]


class PBDError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"pbd_b200 error {code}: {msg}")
        self.code = code


class SolverParams(C.Structure):
    """``pbd_params``: SolverParams in MSG_INIT wire order (48 bytes)."""
    _fields_ = [("substeps", C.c_uint32), ("iterations", C.c_uint32),
                ("dtHint", C.c_float), ("omega", C.c_float),
                ("edgeCompliance", C.c_float), ("volumeCompliance", C.c_float),
                ("gx", C.c_float), ("gy", C.c_float), ("gz", C.c_float),
                ("groundEnabled", C.c_uint32), ("groundY", C.c_float), ("friction", C.c_float)]

    @classmethod
    def default(cls, **kw):
        """Reference defaults, PBDServer.h:147-161 (== PBDRemoteWorld.cs:18-30)."""
        p = cls(2, 6, 1.0 / 60.0, 1.6, 5e-4, 0.0, 0.0, -9.81, 0.0, 1, 0.0, 0.2)
        for k, v in kw.items():
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, v)
        return p

    def copy(self):
        q = SolverParams()
        C.memmove(C.byref(q), C.byref(self), C.sizeof(self))
        return q


class StepStats(C.Structure):
    _fields_ = [("predictMs", C.c_double), ("solveMs", C.c_double), ("commitMs", C.c_double),
                ("packMs", C.c_double), ("totalMs", C.c_double)]


class Options(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("backend", C.c_uint32), ("order_mode", C.c_uint32),
                ("flags", C.c_uint32), ("tile_vertices", C.c_uint32), ("block_threads", C.c_uint32),
                ("max_phases", C.c_uint32), ("partitions", C.c_uint32), ("lanes_per_tet", C.c_uint32),
                ("tiles_per_sm", C.c_uint32), ("shard_world", C.c_uint32), ("shard_rank", C.c_uint32),
                ("plan_sms", C.c_uint32), ("jacobi_edge_stiffness", C.c_float), ("jacobi_volume_stiffness", C.c_float),
                ("reserved", C.c_uint32 * 1)]

    def __init__(self, **kw):
        super().__init__()
        self.struct_size = C.sizeof(Options)
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)


class Info(C.Structure):
    _fields_ = [("V", C.c_uint32), ("E", C.c_uint32), ("T", C.c_uint32), ("backend", C.c_uint32),
                ("edge_colors", C.c_uint32), ("tet_colors", C.c_uint32),
                ("edge_phases", C.c_uint32), ("tet_phases", C.c_uint32), ("tiles", C.c_uint32),
                ("launches_per_frame", C.c_uint32), ("grid_blocks", C.c_uint32), ("block_threads", C.c_uint32),
                ("partitions", C.c_uint32), ("lanes_per_tet", C.c_uint32), ("gather_wavefronts_permille", C.c_uint32 * 2),
                ("device_bytes", C.c_uint64), ("algorithmic_bytes_per_substep", C.c_uint64),
                ("plan_ms", C.c_double), ("upload_ms", C.c_double)]

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_ if not n.startswith("reserved")}
        d["gather_wavefronts_permille"] = list(d["gather_wavefronts_permille"])
        return d


_lib = None


def lib() -> C.CDLL:
    """Load the C-ABI library (built in-tree by ``build.py``).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u32, f32 = C.c_void_p, C.c_uint32, C.c_float
    L.pbd_abi_version.restype = C.c_int
    L.pbd_last_error.restype = C.c_char_p
    L.pbd_device_count.restype = C.c_int
    L.pbd_create.restype = vp
    L.pbd_create.argtypes = [C.POINTER(SolverParams), u32, u32, u32, vp, vp, vp, vp, u32, C.c_int,
                             C.POINTER(Options), C.POINTER(C.c_int)]
    L.pbd_create_from_init.restype = vp
    L.pbd_create_from_init.argtypes = [vp, C.c_uint64, C.c_int, C.POINTER(Options), C.POINTER(C.c_int)]
    L.pbd_init_payload_size.restype = C.c_uint64
    L.pbd_init_payload_size.argtypes = [u32, u32, u32, u32]
    L.pbd_step.argtypes = [vp, f32, C.POINTER(StepStats)]
    L.pbd_step_async.argtypes = [vp, f32, u32]
    L.pbd_sync.argtypes = [vp, C.POINTER(C.c_double)]
    L.pbd_read_positions.argtypes = [vp, vp, C.POINTER(C.c_double)]
    L.pbd_destroy.argtypes = [vp]
    L.pbd_destroy.restype = None
    L.pbd_backend_name.restype = C.c_char_p
    L.pbd_backend_name.argtypes = [vp]
    L.pbd_get_info.argtypes = [vp, C.POINTER(Info)]
    L.pbd_set_params.argtypes = [vp, C.POINTER(SolverParams)]
    L.pbd_get_schedule_order.argtypes = [vp, vp, vp]
    L.pbd_get_schedule_sequence.argtypes = [vp, vp]
    L.pbd_get_array.argtypes = [vp, C.c_int, vp]
    L.pbd_set_colliders.argtypes = [vp, vp, u32, f32]
    L.pbd_set_surface.argtypes = [vp, vp, u32]
    L.pbd_read_normals.argtypes = [vp, vp]
    L.pbd_shard_export.argtypes = [vp, vp]
    L.pbd_shard_attach_ipc.argtypes = [vp, vp]
    L.pbd_shard_attach_local.argtypes = [vp, u32]
    L.pbd_shard_owner.argtypes = [vp, vp]
    L.pbd_plan_create.restype = vp
    L.pbd_plan_create.argtypes = [u32, u32, u32, vp, vp, vp, C.POINTER(Options), C.POINTER(C.c_int)]
    L.pbd_plan_get_info.argtypes = [vp, C.POINTER(Info)]
    L.pbd_plan_get_order.argtypes = [vp, vp, vp]
    L.pbd_plan_get_sequence.argtypes = [vp, vp]
    L.pbd_plan_get_edge_slots.argtypes = [vp, vp, vp, vp]
    L.pbd_plan_get_tet_slots.argtypes = [vp, vp, vp, vp]
    L.pbd_plan_destroy.argtypes = [vp]
    L.pbd_plan_destroy.restype = None
    L.pbd_batch_create.restype = vp
    L.pbd_batch_create.argtypes = [C.POINTER(SolverParams), u32, vp, vp, vp, vp, vp, vp, C.c_int,
                                   C.POINTER(Options), C.POINTER(C.c_int)]
    L.pbd_batch_step.argtypes = [vp, f32, C.POINTER(StepStats)]
    L.pbd_batch_step_async.argtypes = [vp, f32, u32]
    L.pbd_batch_sync.argtypes = [vp, C.POINTER(C.c_double)]
    L.pbd_batch_read_positions.argtypes = [vp, vp, C.POINTER(C.c_double)]
    L.pbd_batch_get_info.argtypes = [vp, C.POINTER(Info)]
    L.pbd_batch_get_schedule_order.argtypes = [vp, vp, vp]
    L.pbd_batch_destroy.argtypes = [vp]
    L.pbd_batch_destroy.restype = None
    _lib = L
    return L


def _check(rc: int):
    if rc != PBD_OK:
        raise PBDError(rc, lib().pbd_last_error().decode())


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def device_count() -> int:
    return lib().pbd_device_count()


class Body:
    """One soft body resident on the GPU: thin object wrapper over pbd_create/step/read/destroy."""

    def __init__(self, params: SolverParams, x0, edges, tets, pinned=None, device: int = -1,
                 options: Options | None = None):
        L = lib()
        x0 = np.ascontiguousarray(x0, dtype=np.float32).reshape(-1, 3)
        edges = np.ascontiguousarray(edges, dtype=np.uint32).reshape(-1, 2)
        tets = np.ascontiguousarray(tets, dtype=np.uint32).reshape(-1, 4)
        pinned = np.ascontiguousarray(pinned if pinned is not None else [], dtype=np.uint32).ravel()
        self.V, self.E, self.T = x0.shape[0], edges.shape[0], tets.shape[0]
        self.params = params.copy()
        st = C.c_int(0)
        self.h = L.pbd_create(C.byref(self.params), self.V, self.E, self.T, _ptr(x0), _ptr(edges), _ptr(tets),
                              _ptr(pinned), pinned.size, device, C.byref(options) if options is not None else None,
                              C.byref(st))
        if not self.h:
            raise PBDError(st.value, L.pbd_last_error().decode())

    @classmethod
    def from_init_payload(cls, payload: bytes, device: int = -1, options: Options | None = None) -> "Body":
        """``pbd_create_from_init``: the body a MSG_INIT payload describes (Server.cpp:30-70)."""
        L = lib()
        buf = bytes(payload)
        self = cls.__new__(cls)
        st = C.c_int(0)
        self.h = L.pbd_create_from_init(buf, len(buf), device, C.byref(options) if options is not None else None,
                                        C.byref(st))
        if not self.h:
            raise PBDError(st.value, L.pbd_last_error().decode())
        i = self.info()
        self.V, self.E, self.T = i["V"], i["E"], i["T"]
        self.params = SolverParams.from_buffer_copy(buf[12:60])
        return self

    def close(self):
        if getattr(self, "h", None):
            try:
                lib().pbd_destroy(self.h)
            except TypeError:      # interpreter shutdown: module globals are already gone
                pass
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def step(self, dt: float, stats: StepStats | None = None):
        _check(lib().pbd_step(self.h, C.c_float(dt), C.byref(stats) if stats is not None else None))

    def step_async(self, dt: float, frames: int = 1):
        _check(lib().pbd_step_async(self.h, C.c_float(dt), frames))

    def sync(self) -> float:
        ms = C.c_double(0.0)
        _check(lib().pbd_sync(self.h, C.byref(ms)))
        return ms.value

    def read_positions(self, out: np.ndarray | None = None, out_ptr: int | None = None) -> np.ndarray | None:
        """Committed positions, caller's vertex order, float32 [V,3].  ``out_ptr`` may be the address
        of a (pinned) host buffer of 12 V bytes."""
        if out_ptr is not None:
            _check(lib().pbd_read_positions(self.h, C.c_void_p(out_ptr), None))
            return None
        if out is None:
            out = np.empty((self.V, 3), dtype=np.float32)
        assert out.dtype == np.float32 and out.size == self.V * 3 and out.flags.c_contiguous
        _check(lib().pbd_read_positions(self.h, out.ctypes.data_as(C.c_void_p), None))
        return out

    def set_params(self, params: SolverParams):
        self.params = params.copy()
        _check(lib().pbd_set_params(self.h, C.byref(self.params)))

    def set_surface(self, surface_tris):
        """``pbd_set_surface``: the boundary triangles (the asset's surfaceTriIds) normals are computed from."""
        t = np.ascontiguousarray(surface_tris, dtype=np.uint32).reshape(-1, 3)
        _check(lib().pbd_set_surface(self.h, _ptr(t), t.shape[0]))

    def read_normals(self) -> np.ndarray:
        """``pbd_read_normals``: float32 [V,3] area-weighted vertex normals of the committed positions."""
        out = np.empty((self.V, 3), dtype=np.float32)
        _check(lib().pbd_read_normals(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def set_colliders(self, colliders, particle_radius: float = 0.0):
        """Primitive colliders of the clamp stage (``pbd_set_colliders``); ``colliders`` as built by
        ``colliders_array`` (an empty list removes them)."""
        a = colliders_array(colliders) if not isinstance(colliders, np.ndarray) else np.ascontiguousarray(colliders)
        _check(lib().pbd_set_colliders(self.h, a.ctypes.data_as(C.c_void_p) if a.size else None, a.size, C.c_float(particle_radius)))

    def info(self) -> dict:
        i = Info()
        _check(lib().pbd_get_info(self.h, C.byref(i)))
        return i.as_dict()

    def name(self) -> str:
        return lib().pbd_backend_name(self.h).decode()

    def schedule_order(self):
        eo = np.empty(self.E, dtype=np.uint32)
        to = np.empty(self.T, dtype=np.uint32)
        _check(lib().pbd_get_schedule_order(self.h, _ptr(eo), _ptr(to)))
        return eo, to

    def schedule_sequence(self) -> np.ndarray:
        it = np.empty(self.E + self.T, dtype=np.uint32)
        _check(lib().pbd_get_schedule_sequence(self.h, it.ctypes.data_as(C.c_void_p)))
        return it

    def get_array(self, what: int) -> np.ndarray:
        shape = {ARRAY_INV_MASS: (self.V,), ARRAY_EDGE_REST: (self.E,), ARRAY_TET_REST: (self.T,),
                 ARRAY_EDGE_LAMBDA: (self.E,), ARRAY_TET_LAMBDA: (self.T,), ARRAY_VELOCITY: (self.V, 3),
                 ARRAY_XSTAR: (self.V, 3)}[what]
        out = np.zeros(shape, dtype=np.float32)
        _check(lib().pbd_get_array(self.h, what, out.ctypes.data_as(C.c_void_p)))
        return out


COLLIDER_SPHERE, COLLIDER_BOX, COLLIDER_CAPSULE = 0, 1, 2
COLLIDER_DTYPE = np.dtype([("type", "<u4"), ("p", "<f4", 3), ("q", "<f4", 4), ("d", "<f4", 3)])   # pbd_collider, 44 bytes


def colliders_array(items) -> np.ndarray:
    """``pbd_collider`` rows from dicts / tuples ``(type, position, quaternion xyzw, data)``:
    sphere data = (radius,), box = half extents, capsule = (radius, half height); the capsule axis is the
    collider's local Y (SoftBodyPrimitiveCollider.PrimitiveColliderData, SoftBodyPrimitiveCollider.cs:8-14)."""
    a = np.zeros(len(items), dtype=COLLIDER_DTYPE)
    for i, it in enumerate(items):
        if isinstance(it, dict):
            it = (it["type"], it["position"], it.get("rotation", (0, 0, 0, 1)), it["data"])
        t, p, q, d = it
        a[i]["type"] = t
        a[i]["p"] = p
        a[i]["q"] = q
        dd = list(d) + [0.0] * (3 - len(d))
        a[i]["d"] = dd[:3]
    return a


SHARD_EXPORT_BYTES = 128


class ShardedBody:
    """ONE body spread over several GPUs of a node (BASELINE config 5), this process holding rank
    ``rank`` of ``world`` (one process per GPU; ``dist`` = an initialised torch.distributed used only
    to exchange the CUDA IPC handles and to assemble read-back positions).  Every rank passes the
    same mesh.  Results are bit-identical to ``Body`` run with ``Options(plan_sms=...)`` equal to the
    SM count the sharded plan was made for."""

    def __init__(self, params: SolverParams, x0, edges, tets, rank: int, world: int, dist, device: int = -1,
                 options: Options | None = None, pinned=None):
        import torch
        opt = Options() if options is None else options
        opt.backend, opt.shard_world, opt.shard_rank = BACKEND_TILE, world, rank
        self.rank, self.world, self.dist = rank, world, dist
        self.body = Body(params, x0, edges, tets, pinned=pinned, device=device, options=opt)
        L = lib()
        mine = np.zeros(SHARD_EXPORT_BYTES, dtype=np.uint8)
        _check(L.pbd_shard_export(self.body.h, _ptr(mine)))
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.from_numpy(mine).to(dev)
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        allh = np.concatenate([o.cpu().numpy() for o in out]).astype(np.uint8)
        _check(L.pbd_shard_attach_ipc(self.body.h, _ptr(allh)))
        self.owner = np.zeros(self.body.V, dtype=np.uint8)
        _check(L.pbd_shard_owner(self.body.h, _ptr(self.owner)))
        dist.barrier()                       # every rank has mapped every peer before anyone steps
        self._dev = dev

    def step_async(self, dt: float, frames: int = 1):
        self.body.step_async(dt, frames)

    def sync(self) -> float:
        return self.body.sync()

    def step(self, dt: float):
        self.body.step_async(dt, 1)
        return self.body.sync()

    def read_positions_local(self) -> np.ndarray:
        """[V,3]; rows of vertices owned by another rank are stale."""
        return self.body.read_positions()

    def read_positions(self) -> np.ndarray:
        """All V committed positions on every rank (owned rows exchanged with an all-reduce)."""
        import torch
        pos = self.body.read_positions()
        pos[self.owner != self.rank] = 0.0
        t = torch.from_numpy(pos).to(self._dev)
        self.dist.all_reduce(t)              # every row has exactly one non-zero contributor
        return t.cpu().numpy()

    def info(self) -> dict:
        return self.body.info()

    def close(self):
        self.body.close()


def sharded_bodies_one_process(params: SolverParams, x0, edges, tets, devices, options: Options | None = None):
    """Single-process variant: one ``Body`` per device in ``devices`` (rank = position), attached
    with peer access.  Step with ``step_async`` on ALL of them before any ``sync``."""
    bodies = []
    for r, dev in enumerate(devices):
        opt = Options() if options is None else options
        o2 = Options()
        C.memmove(C.byref(o2), C.byref(opt), C.sizeof(opt))
        o2.backend, o2.shard_world, o2.shard_rank = BACKEND_TILE, len(devices), r
        bodies.append(Body(params, x0, edges, tets, device=dev, options=o2))
    arr = (C.c_void_p * len(bodies))(*[b.h for b in bodies])
    _check(lib().pbd_shard_attach_local(arr, len(bodies)))
    return bodies


def shard_owner(body: Body) -> np.ndarray:
    owner = np.zeros(body.V, dtype=np.uint8)
    _check(lib().pbd_shard_owner(body.h, _ptr(owner)))
    return owner


class Batch:
    """Many independent bodies stepped by one kernel per frame (pbd_batch_*; BASELINE config 4).
    ``bodies`` is a list of ``(x0 [V,3], edges [E,2], tets [T,4])`` with body-local indices."""

    def __init__(self, params: SolverParams, bodies, device: int = -1, options: Options | None = None):
        L = lib()
        xs = [np.ascontiguousarray(b[0], dtype=np.float32).reshape(-1, 3) for b in bodies]
        es = [np.ascontiguousarray(b[1], dtype=np.uint32).reshape(-1, 2) for b in bodies]
        ts = [np.ascontiguousarray(b[2], dtype=np.uint32).reshape(-1, 4) for b in bodies]
        self.n = len(bodies)
        self.v_off = np.concatenate([[0], np.cumsum([len(x) for x in xs])]).astype(np.uint64)
        self.e_off = np.concatenate([[0], np.cumsum([len(e) for e in es])]).astype(np.uint64)
        self.t_off = np.concatenate([[0], np.cumsum([len(t) for t in ts])]).astype(np.uint64)
        x0 = np.concatenate(xs) if xs else np.zeros((0, 3), np.float32)
        e = np.concatenate(es) if es else np.zeros((0, 2), np.uint32)
        t = np.concatenate(ts) if ts else np.zeros((0, 4), np.uint32)
        self.V, self.E, self.T = int(self.v_off[-1]), int(self.e_off[-1]), int(self.t_off[-1])
        self.params = params.copy()
        st = C.c_int(0)
        self.h = L.pbd_batch_create(C.byref(self.params), self.n, _ptr(self.v_off), _ptr(self.e_off), _ptr(self.t_off),
                                    _ptr(x0), _ptr(e), _ptr(t), device,
                                    C.byref(options) if options is not None else None, C.byref(st))
        if not self.h:
            raise PBDError(st.value, L.pbd_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            lib().pbd_batch_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def step(self, dt: float, stats: StepStats | None = None):
        _check(lib().pbd_batch_step(self.h, C.c_float(dt), C.byref(stats) if stats is not None else None))

    def step_async(self, dt: float, frames: int = 1):
        _check(lib().pbd_batch_step_async(self.h, C.c_float(dt), frames))

    def sync(self) -> float:
        ms = C.c_double(0.0)
        _check(lib().pbd_batch_sync(self.h, C.byref(ms)))
        return ms.value

    def read_positions(self, out: np.ndarray | None = None, out_ptr: int | None = None):
        """Committed positions of all bodies, concatenated in body order, float32 [sum V, 3]."""
        if out_ptr is not None:
            _check(lib().pbd_batch_read_positions(self.h, C.c_void_p(out_ptr), None))
            return None
        if out is None:
            out = np.empty((self.V, 3), dtype=np.float32)
        _check(lib().pbd_batch_read_positions(self.h, out.ctypes.data_as(C.c_void_p), None))
        return out

    def body_positions(self, b: int, pos: np.ndarray) -> np.ndarray:
        return pos[int(self.v_off[b]):int(self.v_off[b + 1])]

    def schedule_order(self, b: int | None = None):
        eo = np.empty(self.E, dtype=np.uint32)
        to = np.empty(self.T, dtype=np.uint32)
        _check(lib().pbd_batch_get_schedule_order(self.h, _ptr(eo), _ptr(to)))
        if b is None:
            return eo, to
        return (eo[int(self.e_off[b]):int(self.e_off[b + 1])], to[int(self.t_off[b]):int(self.t_off[b + 1])])

    def info(self) -> dict:
        i = Info()
        _check(lib().pbd_batch_get_info(self.h, C.byref(i)))
        return i.as_dict()


class Plan:
    """Host-only schedule (no CUDA needed): what ``pbd_create`` would build for these inputs."""

    def __init__(self, x0, edges, tets, options: Options | None = None):
        L = lib()
        x0 = np.ascontiguousarray(x0, dtype=np.float32).reshape(-1, 3)
        edges = np.ascontiguousarray(edges, dtype=np.uint32).reshape(-1, 2)
        tets = np.ascontiguousarray(tets, dtype=np.uint32).reshape(-1, 4)
        self.V, self.E, self.T = x0.shape[0], edges.shape[0], tets.shape[0]
        st = C.c_int(0)
        self.h = L.pbd_plan_create(self.V, self.E, self.T, _ptr(x0), _ptr(edges), _ptr(tets),
                                   C.byref(options) if options is not None else None, C.byref(st))
        if not self.h:
            raise PBDError(st.value, L.pbd_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            lib().pbd_plan_destroy(self.h)
            self.h = None

    __del__ = close

    def info(self) -> dict:
        i = Info()
        _check(lib().pbd_plan_get_info(self.h, C.byref(i)))
        return i.as_dict()

    def order(self):
        eo = np.empty(self.E, dtype=np.uint32)
        to = np.empty(self.T, dtype=np.uint32)
        _check(lib().pbd_plan_get_order(self.h, _ptr(eo), _ptr(to)))
        return eo, to

    def sequence(self) -> np.ndarray:
        it = np.empty(self.E + self.T, dtype=np.uint32)
        _check(lib().pbd_plan_get_sequence(self.h, it.ctypes.data_as(C.c_void_p)))
        return it

    def slots(self, tets: bool):
        n = self.T if tets else self.E
        ph, tl, co = (np.zeros(n, dtype=np.uint32) for _ in range(3))
        f = lib().pbd_plan_get_tet_slots if tets else lib().pbd_plan_get_edge_slots
        _check(f(self.h, _ptr(ph), _ptr(tl), _ptr(co)))
        return ph, tl, co


def pack_init_payload(params: SolverParams, x0, edge_ids, tet_ids, pinned=None) -> bytes:
    """The MSG_INIT payload the Unity client sends (PBDRemoteWorld.cs:294-349, read back by
    Server.cpp:30-70): u32 V,E,T | SolverParams (48 B) | u32 pinnedCount | pinned | x0 | edgeIds | tetIds,
    little-endian, packed."""
    x0 = np.ascontiguousarray(x0, dtype="<f4").reshape(-1, 3)
    edge_ids = np.ascontiguousarray(edge_ids, dtype="<u4").reshape(-1, 2)
    tet_ids = np.ascontiguousarray(tet_ids, dtype="<u4").reshape(-1, 4)
    pinned = np.ascontiguousarray(pinned if pinned is not None else [], dtype="<u4").ravel()
    head = np.array([x0.shape[0], edge_ids.shape[0], tet_ids.shape[0]], dtype="<u4").tobytes()
    return b"".join([head, bytes(params), np.array([pinned.size], dtype="<u4").tobytes(), pinned.tobytes(),
                     x0.tobytes(), edge_ids.tobytes(), tet_ids.tobytes()])


# ------------------------------------------------------------------ reference-shaped host mirror

class PBDState:
    """Host-side ``PBDState`` as ``comm_loop`` builds it from MSG_INIT (Server.cpp:72-104):
    V/E/T, params, world positions x (float32 [V,3]), edge and tet index arrays, pinned indices.
    ``x`` is refreshed by ``CudaStepper.pack_positions``; v/xStar/w/lambdas live in HBM."""
    _gen = 0

    def __init__(self, params: SolverParams, x0, edge_ids, tet_ids, pinned=None):
        self.params = params.copy()
        self.x = np.ascontiguousarray(x0, dtype=np.float32).reshape(-1, 3).copy()
        self.edge_ids = np.ascontiguousarray(edge_ids, dtype=np.uint32).reshape(-1, 2)
        self.tet_ids = np.ascontiguousarray(tet_ids, dtype=np.uint32).reshape(-1, 4)
        self.pinned = np.ascontiguousarray(pinned if pinned is not None else [], dtype=np.uint32).ravel()
        self.V, self.E, self.T = self.x.shape[0], self.edge_ids.shape[0], self.tet_ids.shape[0]
        PBDState._gen += 1
        self.generation = PBDState._gen  # a new INIT = a new generation -> device state is rebuilt


class CudaStepper:
    """``IStepper`` implemented on the GPU (what ``--mode gpu`` would select next to
    ``SerialStepper`` / ``ParallelStepper``, main.cpp:74-78)."""

    def __init__(self, device: int = -1, options: Options | None = None):
        self.device, self.options = device, options
        self._body: Body | None = None
        self._gen = None

    def name(self) -> str:
        return self._body.name() if self._body else "b200"

    def _bind(self, s: PBDState) -> Body:
        if self._body is None or self._gen != s.generation:
            if self._body:
                self._body.close()
            self._body = Body(s.params, s.x, s.edge_ids, s.tet_ids, s.pinned, self.device, self.options)
            self._gen = s.generation
        return self._body

    def step(self, s: PBDState, dt: float, out: StepStats):
        """Advance max(1,substeps) substeps of dt/substeps; ADDS into ``out`` (Sim.cpp:280-305)."""
        self._bind(s).step(dt, out)

    def pack_positions(self, s: PBDState, out_pos: np.ndarray | None = None, stats: StepStats | None = None):
        """All V committed positions in the caller's order (Sim.cpp:307-316); also refreshes s.x."""
        import time
        t0 = time.perf_counter()
        b = self._bind(s)
        if out_pos is None or out_pos.size != 3 * s.V:
            out_pos = np.empty(3 * s.V, dtype=np.float32)
        b.read_positions(out_pos)
        s.x[...] = out_pos.reshape(-1, 3)
        if stats is not None:
            stats.packMs += (time.perf_counter() - t0) * 1e3
        return out_pos

    def close(self):
        if self._body:
            self._body.close()
            self._body = None
