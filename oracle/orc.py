"""ctypes binding of the CPU oracle (oracle/kamr_oracle.c).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libkamr_oracle.so")


class OrcState(C.Structure):
    _fields_ = [(n, C.POINTER(C.c_double)) for n in ("df", "sdf", "flux", "w", "prim", "mflux", "qf", "sw")]


def build(force=False):
    src = os.path.join(HERE, "kamr_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-B", "libkamr_oracle.so"], stdout=subprocess.DEVNULL)
    return LIB


_lib = None
_parity_lib = None
FAST_DIR = os.path.join(HERE, "_fast")


def lib():
    global _lib, _parity_lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.orc_get_tau.restype = C.c_double
        _parity_lib = _lib
    return _lib


def use_fast():
    """Switch this process (and the workers it forks) to a build of the same source for SPEED: -O3 -march=native,
    compiled on the machine it runs on.  Only bench.py's timed CPU legs call this; every comparison
    uses the parity build (-O2 -ffp-contract=off).  Returns a note for the bench line."""
    global _lib
    lib()
    os.makedirs(FAST_DIR, exist_ok=True)
    out = os.path.join(FAST_DIR, "libkamr_oracle_fast.so")
    flags = ["-O3", "-march=native", "-std=c99", "-fPIC", "-fno-fast-math"]
    try:
        subprocess.check_call([os.environ.get("CC", "gcc")] + flags + ["-shared", "-o", out,
                                                                     os.path.join(HERE, "kamr_oracle.c"), "-lm"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        fast = C.CDLL(out)
        fast.orc_get_tau.restype = C.c_double
        _lib = fast
        return "built " + " ".join(flags[:2]) + " on this box for the timed leg"
    except Exception as e:  # pragma: no cover
        return "fast build failed (%r): timed with the parity build -O2 -ffp-contract=off" % (e,)


def use_parity():
    global _lib
    lib()
    _lib = _parity_lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def state_struct(st):
    s = OrcState()
    for n in ("df", "sdf", "flux", "w", "prim", "mflux", "qf", "sw"):
        setattr(s, n, _p(getattr(st, n)))
    return s


def slope(cfg, mesh, st):
    m = mesh.c_struct(); s = state_struct(st)
    rc = lib().orc_slope(C.byref(cfg), C.byref(m), C.byref(s))
    assert rc == 0, f"orc_slope rc={rc}"


def slope_level(cfg, mesh, st, level, transverse):
    m = mesh.c_struct(); s = state_struct(st)
    rc = lib().orc_slope_level(C.byref(cfg), C.byref(m), C.byref(s), int(level), int(transverse))
    assert rc == 0


def macro_slope(cfg, mesh, st):
    m = mesh.c_struct(); s = state_struct(st)
    assert lib().orc_macro_slope(C.byref(cfg), C.byref(m), C.byref(s)) == 0


def ib_solid_cells(cfg, mesh, st):
    m = mesh.c_struct(); s = state_struct(st)
    assert lib().orc_ib_solid_cells(C.byref(cfg), C.byref(m), C.byref(s)) == 0


def ib_solid_neighbors(cfg, mesh, st):
    m = mesh.c_struct(); s = state_struct(st)
    assert lib().orc_ib_solid_neighbors(C.byref(cfg), C.byref(m), C.byref(s)) == 0


def flux(cfg, mesh, st, dt):
    m = mesh.c_struct(); s = state_struct(st)
    rc = lib().orc_flux(C.byref(cfg), C.byref(m), C.byref(s), C.c_double(dt))
    assert rc == 0


def iterate(cfg, mesh, st, dt, want_residual=False):
    m = mesh.c_struct(); s = state_struct(st)
    res = np.zeros(2 * (cfg.dim + 2))
    rc = lib().orc_iterate(C.byref(cfg), C.byref(m), C.byref(s), C.c_double(dt), int(want_residual), _p(res))
    assert rc == 0, f"orc_iterate rc={rc}"
    return res


def step(cfg, mesh, st, dt, want_residual=False):
    m = mesh.c_struct(); s = state_struct(st)
    res = np.zeros(2 * (cfg.dim + 2))
    rc = lib().orc_step(C.byref(cfg), C.byref(m), C.byref(s), C.c_double(dt), int(want_residual), _p(res))
    assert rc == 0, f"orc_step rc={rc}"
    return res


def pair_map(dim, lev_a, lev_b):
    start = np.zeros(len(lev_a) + 1, dtype=np.int32)
    rc = lib().orc_pair_map(dim, len(lev_a), lev_a.ctypes.data_as(C.POINTER(C.c_int8)), len(lev_b),
                            lev_b.ctypes.data_as(C.POINTER(C.c_int8)), start.ctypes.data_as(C.POINTER(C.c_int32)))
    return rc, start


def solve_I_projection(dim, vmid, f, W, weight):
    """solve_I_projection, Theory/I-projection.jl:55-141.  vmid [dim, n] planes; f is shaved in place.
    Returns (lambda[dim+2], Newton systems solved)."""
    n = len(weight)
    lam = np.zeros(dim + 2)
    vm = np.ascontiguousarray(vmid, dtype=np.float64)
    lib().orc_solve_I_projection.restype = C.c_int
    rc = lib().orc_solve_I_projection(dim, n, _p(vm), _p(f), _p(np.ascontiguousarray(W, dtype=np.float64)),
                                      _p(weight), _p(lam))
    assert rc >= 0, "singular Newton system"
    return lam, rc
