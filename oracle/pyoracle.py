"""ctypes face of the two CPU checkers -- ORACLE, test infrastructure, NOT product code.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs import this module.  Nothing under ``cs121-softbodysim_b200/`` does.

* ``kind="port"``      -> oracle/_build/libpbdoracle.so  (oracle/pbd_oracle.c, plain-C restatement)
* ``kind="reference"`` -> oracle/_ref/libpbdref.so       (the unmodified reference Sim.cpp compiled
  where it lies + oracle/ref_harness.cpp; built in the build container, travels prebuilt)

Both expose the same calls, so ``Oracle`` drives either.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "libpbdoracle.so")
REF_SO = os.path.join(HERE, "_ref", "libpbdref.so")

GET_W, GET_EDGE_REST, GET_TET_REST, GET_EDGE_LAMBDA, GET_TET_LAMBDA, GET_V, GET_XSTAR = range(7)


class Params(C.Structure):
    """SolverParams in MSG_INIT wire order (reference CProgram/src/Server.cpp:38-50);
    defaults = CProgram/include/PBDServer.h:147-161."""
    _fields_ = [("substeps", C.c_uint32), ("iterations", C.c_uint32),
                ("dtHint", C.c_float), ("omega", C.c_float),
                ("edgeCompliance", C.c_float), ("volumeCompliance", C.c_float),
                ("gx", C.c_float), ("gy", C.c_float), ("gz", C.c_float),
                ("groundEnabled", C.c_uint32), ("groundY", C.c_float), ("friction", C.c_float)]

    @classmethod
    def default(cls, **kw):
        p = cls(2, 6, 1.0 / 60.0, 1.6, 5e-4, 0.0, 0.0, -9.81, 0.0, 1, 0.0, 0.2)
        for k, v in kw.items():
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, v)
        return p

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def build(force: bool = False) -> None:
    """Compile the port (always possible) and, when /root/reference is present, the reference."""
    args = ["make", "-C", HERE, "-s"] + (["-B"] if force else []) + ["all"]
    subprocess.run(args, check=True, stdout=subprocess.DEVNULL)


def have(kind: str) -> bool:
    return os.path.exists(REF_SO if kind == "reference" else PORT_SO)


_libs: dict[str, tuple[C.CDLL, str]] = {}


def _load(kind: str):
    if kind in _libs:
        return _libs[kind]
    path, pre = (REF_SO, "pbdr_") if kind == "reference" else (PORT_SO, "pbdo_")
    if not os.path.exists(path) and kind == "port":
        build()
    lib = C.CDLL(path)
    f = lambda n: getattr(lib, pre + n)
    f("create").restype = C.c_void_p
    f("create").argtypes = [C.POINTER(Params), C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    f("destroy").argtypes = [C.c_void_p]
    f("permute_constraints").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    f("step").argtypes = [C.c_void_p, C.c_float]
    f("pack").argtypes = [C.c_void_p, C.c_void_p]
    f("get").argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    f("set_inv_mass").argtypes = [C.c_void_p, C.c_void_p]
    f("stats").argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    f("name").restype = C.c_char_p
    if kind == "reference":
        f("use_parallel").argtypes = [C.c_void_p, C.c_uint32]
    else:
        f("step_sequence").argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_uint64]
    _libs[kind] = (lib, pre)
    return _libs[kind]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    """One body stepped on the CPU by the reference (or its C restatement)."""

    def __init__(self, params: Params, x0, edges, tets, pinned=None, kind: str = "port", threads: int = 0):
        self.kind = kind
        self.lib, self.pre = _load(kind)
        x0 = np.ascontiguousarray(x0, dtype=np.float32).reshape(-1, 3)
        edges = np.ascontiguousarray(edges, dtype=np.uint32).reshape(-1, 2)
        tets = np.ascontiguousarray(tets, dtype=np.uint32).reshape(-1, 4)
        pinned = np.ascontiguousarray(pinned if pinned is not None else [], dtype=np.uint32)
        self.V, self.E, self.T = x0.shape[0], edges.shape[0], tets.shape[0]
        self.params = params
        self.h = self._f("create")(C.byref(params), self.V, self.E, self.T, _ptr(x0), _ptr(edges),
                                   _ptr(tets), _ptr(pinned), pinned.size)
        if threads and kind == "reference":
            self._f("use_parallel")(self.h, threads)

    def _f(self, n):
        return getattr(self.lib, self.pre + n)

    def close(self):
        if self.h:
            self._f("destroy")(self.h)
            self.h = None

    __del__ = close

    def permute_constraints(self, edge_order=None, tet_order=None):
        eo = np.ascontiguousarray(edge_order, dtype=np.uint32) if edge_order is not None else None
        to = np.ascontiguousarray(tet_order, dtype=np.uint32) if tet_order is not None else None
        if eo is not None:
            assert eo.size == self.E and np.array_equal(np.sort(eo), np.arange(self.E, dtype=np.uint32))
        if to is not None:
            assert to.size == self.T and np.array_equal(np.sort(to), np.arange(self.T, dtype=np.uint32))
        self._f("permute_constraints")(self.h, _ptr(eo), _ptr(to))

    def step(self, dt: float, frames: int = 1):
        for _ in range(frames):
            self._f("step")(self.h, C.c_float(dt))

    def set_params(self, params: "Params"):
        assert self.kind == "port"
        f = self._f("set_params")
        f.argtypes = [C.c_void_p, C.POINTER(Params)]
        f.restype = None
        f(self.h, C.byref(params))
        self.params = params

    def step_sequence(self, dt: float, items):
        assert self.kind == "port"
        items = np.ascontiguousarray(items, dtype=np.uint32)
        self._f("step_sequence")(self.h, C.c_float(dt), _ptr(items), items.size)

    def positions(self) -> np.ndarray:
        out = np.empty((self.V, 3), dtype=np.float32)
        self._f("pack")(self.h, _ptr(out))
        return out

    def get(self, what: int) -> np.ndarray:
        shape = {GET_W: (self.V,), GET_EDGE_REST: (self.E,), GET_TET_REST: (self.T,),
                 GET_EDGE_LAMBDA: (self.E,), GET_TET_LAMBDA: (self.T,), GET_V: (self.V, 3),
                 GET_XSTAR: (self.V, 3)}[what]
        out = np.empty(shape, dtype=np.float32)
        self._f("get")(self.h, what, _ptr(out))
        return out

    def set_inv_mass(self, w):
        w = np.ascontiguousarray(w, dtype=np.float32)
        assert w.size == self.V
        self._f("set_inv_mass")(self.h, _ptr(w))

    def set_colliders(self, colliders, particle_radius: float):
        """colliders: float32/uint32-compatible structured rows (type, p[3], q[4], d[3]) = 44 bytes each, as
        ``capi.colliders_array`` builds them (port only: PBDServer has no colliders)."""
        assert self.kind == "port"
        a = np.ascontiguousarray(colliders)
        assert a.dtype.itemsize == 44
        f = self._f("set_colliders")
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_float]
        f.restype = None
        f(self.h, a.ctypes.data_as(C.c_void_p) if a.size else None, a.size, C.c_float(particle_radius))

    def stats(self, reset: bool = True) -> dict:
        out = (C.c_double * 5)()
        self._f("stats")(self.h, out, 1 if reset else 0)
        return dict(zip(("predictMs", "solveMs", "commitMs", "packMs", "totalMs"), out))

    def name(self) -> str:
        return self._f("name")().decode()


def push_out(collider, particle_radius: float, point):
    """One point against one collider with the port's restatement of SoftBodyCollisionMath.ComputePushOut:
    returns (hit, push[3])."""
    lib, pre = _load("port")
    a = np.ascontiguousarray(collider)
    assert a.dtype.itemsize == 44 and a.size == 1
    p = np.ascontiguousarray(point, dtype=np.float32)
    out = np.zeros(3, dtype=np.float32)
    f = getattr(lib, pre + "push_out")
    f.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
    f.restype = C.c_int
    hit = f(a.ctypes.data_as(C.c_void_p), C.c_float(particle_radius), p.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return bool(hit), out


# ---------------------------------------------------------------- residual metrics (P3)

def residuals(pos, x0, edges, tets, ground_y=0.0):
    """Absolute edge residual RMS |len-rest|, relative total-volume error, min y, finite flag
    (absolute, not relative, edge residuals -- see SURVEY.md 7)."""
    pos = np.asarray(pos, dtype=np.float64).reshape(-1, 3)
    x0 = np.asarray(x0, dtype=np.float64).reshape(-1, 3)
    e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    t = np.asarray(tets, dtype=np.int64).reshape(-1, 4)

    def elen(p):
        return np.linalg.norm(p[e[:, 0]] - p[e[:, 1]], axis=1)

    def vol(p):
        a, b, c, d = p[t[:, 0]], p[t[:, 1]], p[t[:, 2]], p[t[:, 3]]
        return np.einsum("ij,ij->i", np.cross(b - a, c - a), d - a) / 6.0

    v0, v1 = vol(x0), vol(pos)
    used = np.zeros(len(pos), dtype=bool)
    used[t.ravel()] = True
    return {
        "edge_rms": float(np.sqrt(np.mean((elen(pos) - elen(x0)) ** 2))) if len(e) else 0.0,
        "vol_rel": float(abs(v1.sum() - v0.sum()) / max(abs(v0.sum()), 1e-30)) if len(t) else 0.0,
        "tet_vol_rms": float(np.sqrt(np.mean((v1 - v0) ** 2))) if len(t) else 0.0,
        "min_y_dynamic": float(pos[used, 1].min() - ground_y) if used.any() else 0.0,
        "finite": bool(np.isfinite(pos).all()),
    }


# ---------------------------------------------------------------- Jacobi + SOR comparison schedule (test infrastructure)

def jacobi_reference(params: "Params", x0, edges, tets, frames: int, dt: float, edge_k=0.9, tet_k=0.98, omega=None, pinned=None):
    """numpy float32 restatement of the reference's in-engine gather solver (SoftBodySolver.cs:379-527 ==
    SoftBodyCompute.compute:229-389) on top of the PBDServer integrator (predict / ground / commit, Sim.cpp:178-222)
    and PBDServer's inverse masses: what PBD_BACKEND_JACOBI computes.  Every operation is a float32 numpy op in the
    kernels' order (products rounded, sums left to right, per-vertex accumulation in ascending constraint order), so the
    GPU result can be compared bit for bit.  Returns positions [V,3] float32."""
    f = np.float32
    x = np.ascontiguousarray(x0, dtype=np.float32).reshape(-1, 3).copy()
    edges = np.ascontiguousarray(edges, dtype=np.int64).reshape(-1, 2)
    tets = np.ascontiguousarray(tets, dtype=np.int64).reshape(-1, 4)
    V = len(x)
    ora = Oracle(params, x, edges.astype(np.uint32), tets.astype(np.uint32), pinned=pinned, kind="port")
    w, e_rest, t_rest = ora.get(GET_W), ora.get(GET_EDGE_REST), ora.get(GET_TET_REST)
    ora.close()
    omega = f(params.omega if omega is None else omega)
    ss = max(1, int(params.substeps))
    sdt = f(f(dt) / f(ss))
    inv_dt = f(1.0) / sdt if sdt > 1e-12 else f(0.0)
    g = np.array([params.gx, params.gy, params.gz], np.float32) * sdt
    fr = f(1.0) - f(min(max(params.friction, 0.0), 1.0))
    v = np.zeros_like(x)
    dyn = w != 0

    def dot(a, b):
        return (a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1]) + a[:, 2] * b[:, 2]

    def cross(a, b):
        return np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1], a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2], a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], 1)

    def apply(p, vert, contrib, valid):
        # per-vertex accumulation in ascending constraint order, then p += (omega / cnt) * sum
        order = np.lexsort((np.arange(len(vert)), vert))
        order = order[valid[order]]
        s = np.zeros((V, 3), np.float32)
        for k in range(3):
            np.add.at(s[:, k], vert[order], contrib[order, k])
        cnt = np.bincount(vert[order], minlength=V)
        m = (cnt > 0) & dyn
        fac = (omega / cnt[m].astype(np.float32)).astype(np.float32)
        p[m] = p[m] + fac[:, None] * s[m]

    for _ in range(frames):
        for _ in range(ss):
            v[dyn] = v[dyn] + g
            p = x.copy()
            p[dyn] = x[dyn] + v[dyn] * sdt
            for _ in range(int(params.iterations)):
                if len(edges):
                    vert = np.concatenate([edges[:, 0], edges[:, 1]])           # constraint e seen from a, then from b
                    other = np.concatenate([edges[:, 1], edges[:, 0]])
                    eidx = np.concatenate([np.arange(len(edges))] * 2)
                    wi, wj = w[vert], w[other]
                    ws = wi + wj
                    d = p[vert] - p[other]
                    len2 = dot(d, d)
                    valid = (wi != 0) & (ws != 0) & ~(len2 < 1e-18)
                    with np.errstate(all="ignore"):
                        ln = np.sqrt(len2)
                        lam = f(-edge_k) * ((ln - e_rest[eidx]) / ws)
                        contrib = (d / ln[:, None]) * (lam * wi)[:, None]
                    order_key = eidx * 2                                           # adjacency order = ascending edge index
                    o = np.lexsort((order_key, vert))
                    apply_sorted = (vert[o], contrib[o], valid[o])
                    apply(p, *apply_sorted)
                if len(tets):
                    pa, pb, pc, pd = (p[tets[:, k]] for k in range(4))
                    wa, wb, wc, wd = (w[tets[:, k]] for k in range(4))
                    six = f(6.0)
                    ga = cross(pd - pb, pc - pb) / six
                    gb = cross(pc - pa, pd - pa) / six
                    gc = cross(pd - pa, pb - pa) / six
                    n = cross(pb - pa, pc - pa)
                    gd = n / six
                    wsum = ((wa * dot(ga, ga) + wb * dot(gb, gb)) + wc * dot(gc, gc)) + wd * dot(gd, gd)
                    massive = (((wa + wb) + wc) + wd) != 0
                    ok = massive & ~(wsum < 1e-20)
                    with np.errstate(all="ignore"):
                        vol = dot(n, pd - pa) / six
                        lam = f(-tet_k) * ((vol - t_rest) / wsum)
                    vert = np.concatenate([tets[:, k] for k in range(4)])
                    grads = np.concatenate([ga, gb, gc, gd])
                    wrole = np.concatenate([wa, wb, wc, wd])
                    lam4 = np.concatenate([lam] * 4)
                    valid = np.concatenate([ok] * 4) & (wrole != 0) & (w[vert] != 0)
                    contrib = grads * (lam4 * wrole)[:, None]
                    tidx = np.concatenate([np.arange(len(tets))] * 4)
                    o = np.lexsort((tidx, vert))
                    apply(p, vert[o], contrib[o], valid[o])
                if params.groundEnabled:
                    low = dyn & (p[:, 1] < f(params.groundY))
                    p[low, 1] = f(params.groundY)
            vel = (p - x) * inv_dt
            if params.groundEnabled:
                c = dyn & (p[:, 1] <= f(params.groundY) + f(1e-6))
                vel[c, 0] = vel[c, 0] * fr
                vel[c, 2] = vel[c, 2] * fr
                vel[c, 1] = np.maximum(vel[c, 1], f(0.0))
            v[dyn] = vel[dyn]
            v[~dyn] = 0
            x[dyn] = p[dyn]
    return x
