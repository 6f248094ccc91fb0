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


def ps_criterion(cfg, mesh, st, threshold, ghost_flag=None):
    """update_criterion!(ka): returns (lohner [n_local, DIM, DIM+2], sensor [n_local], pre-buffer flag [n_local])."""
    m = mesh.c_struct(); s = state_struct(st)
    D, M = cfg.dim, cfg.dim + 2
    loh = np.zeros((mesh.n_local, D, M)); sen = np.zeros(mesh.n_local); flg = np.zeros(mesh.n_local, dtype=np.int32)
    gf = None
    if ghost_flag is not None:
        gfa = np.ascontiguousarray(ghost_flag, dtype=np.int32)
        gf = gfa.ctypes.data_as(C.POINTER(C.c_int32))
    f = lib().orc_ps_criterion
    f.restype = C.c_int
    rc = f(C.byref(cfg), C.byref(m), C.byref(s), C.c_double(threshold), gf, _p(loh), _p(sen),
           flg.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0, f"orc_ps_criterion rc={rc}"
    return loh, sen, flg


def vs_face_neighbors(dim, vmid, level, par):
    """[n, DIM, 2] same-or-coarser face neighbours (vs_face_neighbor, Velocity_space/Neighbor.jl:181), -1: none"""
    n = len(level)
    out = np.zeros((n, dim, 2), dtype=np.int32)
    vmid = np.ascontiguousarray(vmid, dtype=np.float64); level = np.ascontiguousarray(level, dtype=np.int8)
    f = lib().orc_vs_face_neighbors
    f.restype = C.c_int
    rc = f(int(dim), int(n), _p(vmid), level.ctypes.data_as(C.POINTER(C.c_int8)), C.byref(par),
           out.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0, f"orc_vs_face_neighbors rc={rc}"
    return out


def vs_resolution(cfg, mesh, st, par):
    m = mesh.c_struct(); s = state_struct(st)
    out = np.zeros(2)
    f = lib().orc_vs_resolution
    f.restype = C.c_int
    assert f(C.byref(cfg), C.byref(m), C.byref(s), C.byref(par), _p(out)) == 0
    return out


def vs_criterion(cfg, mesh, st, par):
    """(refine_flag, coarsen_ok) per local velocity point (uint8), vs_refine! / vs_coarsen! of Velocity_space/AMR.jl"""
    m = mesh.c_struct(); s = state_struct(st)
    npts = int(mesh.vs_off()[mesh.n_local])
    rf = np.zeros(npts, dtype=np.uint8); co = np.zeros(npts, dtype=np.uint8)
    f = lib().orc_vs_criterion
    f.restype = C.c_int
    u8 = C.POINTER(C.c_uint8)
    rc = f(C.byref(cfg), C.byref(m), C.byref(s), C.byref(par), rf.ctypes.data_as(u8), co.ctypes.data_as(u8))
    assert rc == 0, f"orc_vs_criterion rc={rc}"
    return rf, co


def project_cells(cfg, mesh, st, cells):
    """vs_conserved_correction! (Velocity_space/AMR.jl:120-133) on the listed local cells, in place"""
    m = mesh.c_struct(); s = state_struct(st)
    cells = np.ascontiguousarray(cells, dtype=np.int32)
    f = lib().orc_project_cells
    f.restype = C.c_int
    rc = f(C.byref(cfg), C.byref(m), C.byref(s), len(cells), cells.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0, f"orc_project_cells rc={rc}"


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
