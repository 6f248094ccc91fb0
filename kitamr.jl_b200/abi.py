"""ctypes mirror of include/kamr.h and loader of libkamr.so (the C-ABI drop-in boundary).

There is NO CPU fallback: if the CUDA library is missing, `load()` raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libkamr.so")

FLUX_CAIDVM, FLUX_DVM = 0, 1
MARCH_CAIDVM, MARCH_CIP, MARCH_EULER = 0, 1, 2
BC_MAXWELLIAN, BC_SUPERSONIC_INFLOW, BC_UNIFORM_OUTFLOW, BC_INTERPOLATED_OUTFLOW = 0, 1, 2, 3
OPT_KEEP_SDF = 1
DL_DF, DL_SDF, DL_FLUX, DL_W, DL_PRIM, DL_QF, DL_SW, DL_MFLUX = 1, 2, 4, 8, 16, 32, 64, 128

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_i8p = C.POINTER(C.c_int8)
c_f64p = C.POINTER(C.c_double)


class KamrConfig(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("ndf", C.c_int32), ("flux_type", C.c_int32), ("marching", C.c_int32),
        ("K", C.c_double), ("Pr", C.c_double), ("gamma", C.c_double), ("omega", C.c_double),
        ("mu_ref", C.c_double),
        ("device", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32),
        ("stream", C.c_void_p),
    ]


class KamrIB(C.Structure):
    _fields_ = [
        ("n_solid", C.c_int32), ("solid_cell", c_i32p), ("solid_nb_off", c_i32p), ("solid_nb_ids", c_i32p),
        ("n_sn", C.c_int32), ("sn_donor", c_i32p), ("sn_solid", c_i32p), ("sn_faceid", c_i32p),
        ("sn_aux", c_f64p), ("sn_normal", c_f64p), ("sn_bc", c_f64p),
        ("sn_nb_off", c_i32p), ("sn_nb_ids", c_i32p),
        ("cvc_off", c_i32p), ("cvc_index", c_i32p), ("cvc_gas_w", c_f64p), ("cvc_solid_w", c_f64p),
    ]


class KamrMesh(C.Structure):
    _fields_ = [
        ("n_local", C.c_int32), ("n_ghost", C.c_int32), ("n_solidnbr", C.c_int32),
        ("ds", c_f64p), ("mid", c_f64p), ("bound_enc", c_i32p), ("ps_level", c_i32p), ("cell_grid", c_i32p),
        ("n_grid", C.c_int32), ("grid_off", c_i64p), ("v_level", c_i8p), ("v_weight", c_f64p), ("v_mid", c_f64p),
        ("nb_state", c_i32p), ("nb_off", c_i32p), ("nb_ids", c_i32p),
        ("ps_maxlevel", C.c_int32), ("ps_minlevel", C.c_int32),
        ("n_face", C.c_int32), ("face_kind", c_i32p), ("face_here", c_i32p), ("face_there", c_i32p),
        ("face_dir", c_i32p), ("face_rot", c_f64p), ("face_mid", c_f64p), ("face_there_mid", c_f64p),
        ("n_bc", C.c_int32), ("bc_type", c_i32p), ("bc_prim", c_f64p),
        ("n_peer", C.c_int32), ("peer_rank", c_i32p), ("send_off", c_i32p), ("send_cells", c_i32p),
        ("recv_off", c_i32p),
        ("ib", C.POINTER(KamrIB)),
    ]


class KamrStats(C.Structure):
    _fields_ = [
        ("n_phase_local", C.c_int64), ("n_points_total", C.c_int64), ("n_relations", C.c_int64),
        ("n_slots", C.c_int64), ("kernel_launches", C.c_int64), ("device_bytes", C.c_int64),
        ("halo_bytes_per_step", C.c_int64), ("n_levels", C.c_int32), ("fused_cells", C.c_int32),
    ]


class KamrVsAdapt(C.Structure):
    """== kamr_vs_adapt (include/kamr.h): parameters of vs_refine! / vs_coarsen!, Velocity_space/AMR.jl:26-115"""
    _fields_ = [
        ("mode", C.c_int32), ("maxlevel", C.c_int32), ("trees", C.c_int32 * 3), ("pad_", C.c_int32),
        ("vmin", C.c_double * 3), ("vmax", C.c_double * 3),
        ("coeff_lohner", C.c_double), ("coeff_local", C.c_double), ("coeff_global", C.c_double),
        ("vr_density", C.c_double), ("vr_energy", C.c_double),
    ]


VS_MODE_LOHNER, VS_MODE_CONTRIBUTION = 0, 1


def vs_adapt(case, mode=VS_MODE_LOHNER, coeff_lohner=0.6, coeff_local=1e-2, coeff_global=0.125,
             vr_density=0.0, vr_energy=0.0) -> "KamrVsAdapt":
    """kamr_vs_adapt of a synthetic case with the reference's default coefficients (Solver/Types.jl:96-100)"""
    D = case.dim
    p = KamrVsAdapt()
    p.mode, p.maxlevel = int(mode), int(case.vs_maxlevel)
    for d in range(3):
        p.trees[d] = int(case.vs_trees_num[d]) if d < D else 1
        p.vmin[d] = float(case.quadrature[2 * d]) if d < D else 0.0
        p.vmax[d] = float(case.quadrature[2 * d + 1]) if d < D else 1.0
    p.coeff_lohner, p.coeff_local, p.coeff_global = coeff_lohner, coeff_local, coeff_global
    p.vr_density, p.vr_energy = vr_density, vr_energy
    return p


class KamrKernelTime(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("launches", C.c_int64), ("total_ms", C.c_double)]


def ptr(a: np.ndarray, ctype):
    """Pointer to a C-contiguous numpy array of the matching dtype (None -> NULL)."""
    if a is None:
        return C.cast(None, C.POINTER(ctype))
    assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return a.ctypes.data_as(C.POINTER(ctype))


# every symbol include/kamr.h declares (checked by the CPU test-suite)
EXPORTS = [
    "kamr_create", "kamr_destroy", "kamr_last_error", "kamr_version", "kamr_comm_unique_id", "kamr_comm_init",
    "kamr_upload_topology", "kamr_upload_state", "kamr_upload_aux", "kamr_download_state",
    "kamr_slope", "kamr_flux", "kamr_iterate", "kamr_step", "kamr_exchange_df", "kamr_sync",
    "kamr_get_stats", "kamr_get_pair_map", "kamr_get_cell_slots", "kamr_profile_enable", "kamr_profile_read", "kamr_set_option",
    "kamr_debug_exp_nonpos", "kamr_pack_cells", "kamr_unpack_cells", "kamr_ps_criterion",
    "kamr_vs_resolution", "kamr_vs_criterion", "kamr_project_cells", "kamr_migrate_begin", "kamr_migrate_finish",
]

_lib = None


def load(path: str | None = None):
    """dlopen libkamr.so and set the prototypes.  Raises if the library is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("KAMR_LIB", LIB_PATH)
    if not os.path.exists(p):
        raise RuntimeError(
            f"libkamr.so not found at {p}: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the hot path)")
    lib = C.CDLL(p)
    vp = C.c_void_p
    lib.kamr_create.argtypes = [C.POINTER(KamrConfig), C.POINTER(vp)]
    lib.kamr_destroy.argtypes = [vp]
    lib.kamr_last_error.argtypes = [vp]
    lib.kamr_last_error.restype = C.c_char_p
    lib.kamr_version.argtypes = []
    lib.kamr_comm_unique_id.argtypes = [vp]
    lib.kamr_comm_init.argtypes = [vp, vp]
    lib.kamr_upload_topology.argtypes = [vp, C.POINTER(KamrMesh)]
    lib.kamr_upload_state.argtypes = [vp, c_f64p, c_f64p, c_f64p]
    lib.kamr_upload_aux.argtypes = [vp, c_f64p, c_f64p, c_f64p]
    lib.kamr_download_state.argtypes = [vp, C.c_uint32] + [c_f64p] * 8
    lib.kamr_slope.argtypes = [vp]
    lib.kamr_flux.argtypes = [vp, C.c_double]
    lib.kamr_iterate.argtypes = [vp, C.c_double, C.c_int32, c_f64p]
    lib.kamr_step.argtypes = [vp, C.c_double, C.c_int32, c_f64p]
    lib.kamr_exchange_df.argtypes = [vp]
    lib.kamr_sync.argtypes = [vp]
    lib.kamr_get_stats.argtypes = [vp, C.POINTER(KamrStats)]
    lib.kamr_get_pair_map.argtypes = [vp, C.c_int32, C.c_int32, c_i32p, C.c_int32]
    lib.kamr_get_cell_slots.argtypes = [vp, C.c_int32, c_i32p, c_i32p, C.c_int32, c_i32p]
    lib.kamr_set_option.argtypes = [vp, C.c_int32, C.c_int32]
    lib.kamr_profile_enable.argtypes = [vp, C.c_int32]
    lib.kamr_profile_read.argtypes = [vp, C.POINTER(KamrKernelTime), C.c_int32, c_i32p]
    lib.kamr_debug_exp_nonpos.argtypes = [vp, c_f64p, c_f64p, C.c_int64]
    lib.kamr_pack_cells.argtypes = [vp, C.c_int32, c_i32p, c_f64p, c_f64p]
    lib.kamr_unpack_cells.argtypes = [vp, C.c_int32, c_i32p, c_f64p, c_f64p]
    lib.kamr_ps_criterion.argtypes = [vp, C.c_double, c_f64p, c_f64p]
    lib.kamr_migrate_begin.argtypes = [vp, C.c_int32, c_i32p, c_i32p, C.c_int32, c_i32p, c_i32p, C.POINTER(C.c_int64)]
    lib.kamr_migrate_finish.argtypes = [vp, C.c_int32, c_i32p]
    lib.kamr_project_cells.argtypes = [vp, C.c_int32, c_i32p]
    lib.kamr_vs_resolution.argtypes = [vp, C.POINTER(KamrVsAdapt), c_f64p]
    lib.kamr_vs_criterion.argtypes = [vp, C.POINTER(KamrVsAdapt), C.POINTER(C.c_uint8), C.POINTER(C.c_uint8)]
    for name in EXPORTS:
        if name != "kamr_last_error":
            getattr(lib, name).restype = C.c_int
    if path is None:
        _lib = lib
    return lib
