"""cs121-softbodysim_b200 -- B200-native XPBD soft-body substep behind the PBDServer stepper seam.

The directory name carries a hyphen (it mirrors the reference repository's name), so import it
with ``importlib.import_module("cs121-softbodysim_b200")``; ``__graft_entry__.py``, ``bench.py``
and ``tests/conftest.py`` do exactly that.

Contents (only what the hot path needs):
  csrc/      hand-written sm_100a CUDA kernels, the host-side schedule builder and the C ABI
             (``include/pbd_b200.h``) compiled into ``libpbd_b200.so``
  capi.py    ctypes binding of that C ABI + the host-side mirror of the reference's
             ``IStepper`` / ``PBDState`` / ``SolverParams`` / ``perf::StepStats``
  meshgen.py deterministic inputs (Kuhn tet grids, edge builder, placement)
  shard.py   which bodies a rank owns + the max-over-ranks timing reduction (multi-GPU batches)
  build.py   the nvcc command line (``-gencode arch=compute_100a,code=sm_100a -lineinfo``)
  wire.py    the PBD1 wire protocol, client side (what ``PBDRemoteWorld.cs`` speaks); the server side
             is ``csrc/pbd_server.cpp`` -> ``pbd_server``, a native binary on top of the C ABI

There is no CPU fallback: every solver call goes through ``libpbd_b200.so`` and raises if
the library or a CUDA device is missing.
"""
from . import meshgen  # noqa: F401

__all__ = ["meshgen"]


def __getattr__(name):
    # capi/build are imported lazily so that meshgen stays usable before the library is built
    if name in ("capi", "build", "shard", "wire"):
        import importlib

        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
