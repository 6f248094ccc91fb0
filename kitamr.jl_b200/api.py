"""Thin object wrapper over the C-ABI (include/kamr.h).  All compute happens in libkamr.so on the GPU;
this module only marshals numpy arrays across the boundary."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi
from .model import HostMesh, HostState


class KamrError(RuntimeError):
    pass


class Context:
    def __init__(self, cfg: abi.KamrConfig, lib=None):
        self.lib = lib or abi.load()
        self.cfg = cfg
        self.D, self.K, self.M = cfg.dim, cfg.ndf, cfg.dim + 2
        h = C.c_void_p()
        rc = self.lib.kamr_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise KamrError(self.lib.kamr_last_error(None).decode())
        self.h = h
        self.mesh = None

    def _ck(self, rc):
        if rc != 0:
            raise KamrError(self.lib.kamr_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.kamr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- multi-GPU -------------------------------------------------------------------------
    def unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        if self.lib.kamr_comm_unique_id(buf) != 0:
            raise KamrError(self.lib.kamr_last_error(None).decode())
        return buf.raw

    def comm_init(self, uid: bytes):
        buf = C.create_string_buffer(uid, 128)
        self._ck(self.lib.kamr_comm_init(self.h, buf))

    # ---- topology / state ------------------------------------------------------------------
    def upload_topology(self, mesh: HostMesh):
        self.mesh = mesh
        m = mesh.c_struct()
        self._ck(self.lib.kamr_upload_topology(self.h, C.byref(m)))

    def upload_state(self, st: HostState, aux=False):
        p = lambda a: a.ctypes.data_as(abi.c_f64p)
        self._ck(self.lib.kamr_upload_state(self.h, p(st.df), p(st.w), p(st.prim)))
        if aux:
            self._ck(self.lib.kamr_upload_aux(self.h, p(st.sdf), p(st.flux), p(st.mflux)))

    def download_state(self, st: HostState, mask=0xFF):
        p = lambda a: a.ctypes.data_as(abi.c_f64p)
        self._ck(self.lib.kamr_download_state(self.h, mask, p(st.df), p(st.sdf), p(st.flux), p(st.w), p(st.prim),
                                              p(st.qf), p(st.sw), p(st.mflux)))
        return st

    # ---- hot path --------------------------------------------------------------------------
    def slope(self):
        self._ck(self.lib.kamr_slope(self.h))

    def flux(self, dt):
        self._ck(self.lib.kamr_flux(self.h, dt))

    def iterate(self, dt, want_residual=False):
        res = np.zeros(2 * self.M)
        self._ck(self.lib.kamr_iterate(self.h, dt, int(want_residual), res.ctypes.data_as(abi.c_f64p)))
        return res

    def step(self, dt, want_residual=False):
        res = np.zeros(2 * self.M)
        self._ck(self.lib.kamr_step(self.h, dt, int(want_residual), res.ctypes.data_as(abi.c_f64p)))
        return res

    def exchange_df(self):
        self._ck(self.lib.kamr_exchange_df(self.h))

    def sync(self):
        self._ck(self.lib.kamr_sync(self.h))

    # ---- introspection ---------------------------------------------------------------------
    def stats(self) -> abi.KamrStats:
        s = abi.KamrStats()
        self._ck(self.lib.kamr_get_stats(self.h, C.byref(s)))
        return s

    def set_option(self, option, value):
        self._ck(self.lib.kamr_set_option(self.h, int(option), int(value)))

    def profile(self, on=True):
        self._ck(self.lib.kamr_profile_enable(self.h, int(on)))

    def profile_read(self):
        """{kernel name: (launches, total_ms)} since the last read."""
        buf = (abi.KamrKernelTime * 32)()
        n = C.c_int32(0)
        self._ck(self.lib.kamr_profile_read(self.h, buf, 32, C.byref(n)))
        return {buf[i].name.decode(): (int(buf[i].launches), float(buf[i].total_ms)) for i in range(n.value)}

    def pack_cells(self, cells):
        """(df, w) of the listed local cells, packed on the device (partition migration, Parallel/Partition.jl:339-388)"""
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        n_of = self.mesh.cell_n()[cells].astype(np.int64)
        df = np.empty(int(n_of.sum()) * self.K)
        w = np.empty(len(cells) * self.M)
        self._ck(self.lib.kamr_pack_cells(self.h, len(cells), cells.ctypes.data_as(abi.c_i32p),
                                          df.ctypes.data_as(abi.c_f64p), w.ctypes.data_as(abi.c_f64p)))
        return df, w

    def unpack_cells(self, cells, df, w):
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        df = np.ascontiguousarray(df, dtype=np.float64); w = np.ascontiguousarray(w, dtype=np.float64)
        self._ck(self.lib.kamr_unpack_cells(self.h, len(cells), cells.ctypes.data_as(abi.c_i32p),
                                            df.ctypes.data_as(abi.c_f64p), w.ctypes.data_as(abi.c_f64p)))

    def ps_criterion(self, threshold, want_lohner=True):
        """update_criterion!(ka) (Physical_space/AMR.jl:256-341) after slope(): (lohner [n_local, DIM, DIM+2] or None,
        ps_sensor [n_local])"""
        n = self.mesh.n_local
        loh = np.empty((n, self.D, self.M)) if want_lohner else None
        sen = np.empty(n)
        self._ck(self.lib.kamr_ps_criterion(self.h, float(threshold),
                                            loh.ctypes.data_as(abi.c_f64p) if want_lohner else None,
                                            sen.ctypes.data_as(abi.c_f64p)))
        return loh, sen

    def migrate_begin(self, cells, dest_rank, src_rank, src_cells, src_points):
        """phase 1 of the device-to-device partition migration (old topology); see include/kamr.h"""
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        cells, dest_rank, src_rank, src_cells = i32(cells), i32(dest_rank), i32(src_rank), i32(src_cells)
        src_points = np.ascontiguousarray(src_points, dtype=np.int64)
        p = lambda a: a.ctypes.data_as(abi.c_i32p)
        self._ck(self.lib.kamr_migrate_begin(self.h, len(cells), p(cells), p(dest_rank), len(src_rank), p(src_rank),
                                             p(src_cells), src_points.ctypes.data_as(C.POINTER(C.c_int64))))

    def migrate_finish(self, recv_cells):
        """phase 2 (new topology): arrival q becomes local cell recv_cells[q]"""
        recv_cells = np.ascontiguousarray(recv_cells, dtype=np.int32)
        self._ck(self.lib.kamr_migrate_finish(self.h, len(recv_cells), recv_cells.ctypes.data_as(abi.c_i32p)))

    def project_cells(self, cells):
        """conserved_I_porjection! of the listed local cells (vs_conserved_correction!, Velocity_space/AMR.jl:120-133)"""
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        self._ck(self.lib.kamr_project_cells(self.h, len(cells), cells.ctypes.data_as(abi.c_i32p)))

    def vs_resolution(self, par):
        """local maxima of vs_resolution(ps_data, kinfo) (Velocity_space/AMR.jl:139-166): [density, energy]"""
        out = np.zeros(2)
        self._ck(self.lib.kamr_vs_resolution(self.h, C.byref(par), out.ctypes.data_as(abi.c_f64p)))
        return out

    def vs_criterion(self, par):
        """(refine_flag, coarsen_ok) per local velocity point, host order (vs_refine! / vs_coarsen!, AMR.jl:26-115)"""
        npts = int(self.mesh.vs_off()[self.mesh.n_local])
        rf = np.zeros(npts, dtype=np.uint8); co = np.zeros(npts, dtype=np.uint8)
        u8 = C.POINTER(C.c_uint8)
        self._ck(self.lib.kamr_vs_criterion(self.h, C.byref(par), rf.ctypes.data_as(u8), co.ctypes.data_as(u8)))
        return rf, co

    def debug_exp_nonpos(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        self._ck(self.lib.kamr_debug_exp_nonpos(self.h, x.ctypes.data_as(abi.c_f64p), y.ctypes.data_as(abi.c_f64p),
                                                x.size))
        return y

    def pair_map(self, ga, gb):
        n = int(self.mesh.grid_off[ga + 1] - self.mesh.grid_off[ga])
        start = np.zeros(n + 1, dtype=np.int32)
        rc = self.lib.kamr_get_pair_map(self.h, ga, gb, start.ctypes.data_as(abi.c_i32p), n + 1)
        if rc == 1:
            return None
        if rc != 0:
            raise KamrError(self.lib.kamr_last_error(self.h).decode())
        return start

    def cell_slots(self, cell):
        face = np.zeros(64, dtype=np.int32); sign = np.zeros(64, dtype=np.int32)
        n = C.c_int32(0)
        self._ck(self.lib.kamr_get_cell_slots(self.h, cell, face.ctypes.data_as(abi.c_i32p),
                                              sign.ctypes.data_as(abi.c_i32p), 64, C.byref(n)))
        return face[: n.value].copy(), sign[: n.value].copy()
