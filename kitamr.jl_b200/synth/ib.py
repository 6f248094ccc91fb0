"""Synthetic immersed-boundary inputs (host side, set-up time): what the reference computes once per re-flatten and
the flattener ships as data to kernel (d).

  cell classification   src/Boundary/Circle.jl:33-35 (solid_flag), :124-143 (ghost_cell_flag)
  wall intersection     src/Boundary/Circle.jl:52-82 (calc_intersect for Circle / Sphere)
  cut velocity cells    src/Boundary/Immersed_boundary.jl:226-277, src/Velocity_space/Cut_cell.jl:13-77,126-200
                        (here: exact half-space clipping of the velocity cell by the plane v.n = 0; gas side v.n < 0)
"""
from __future__ import annotations

import ctypes as C
import itertools
from dataclasses import dataclass

import numpy as np

from .. import abi

EPS = 1e-12

FLUID, SOLID_GHOST, INSIDE_SOLID = 0, -1, -2


@dataclass
class Ball:
    """Circle (2-D) / Sphere (3-D) immersed boundary, solid inside; Maxwellian wall with prim `bc`.  `center` may
    hold several centres [m, DIM] (the weak-scaling workload places one body per GPU): every query uses the nearest."""
    centers: np.ndarray
    radius: float
    bc: np.ndarray

    def __post_init__(self):
        self.centers = np.atleast_2d(np.asarray(self.centers, dtype=np.float64))

    def nearest(self, x):
        d = np.linalg.norm(self.centers - np.asarray(x)[None, :], axis=1)
        return self.centers[int(np.argmin(d))]

    @property
    def center(self):
        return self.centers[0]

    def ghost_cell_flag(self, mid, ds):
        """any face-neighbour position on the other side of the boundary (NMT table, Abstract/Types.jl:12)"""
        D = len(mid)
        flag = 0
        for d in range(D):
            for s in (-1.0, 1.0):
                p = np.array(mid, dtype=np.float64); p[d] += s * ds[d]
                flag += 1 if np.linalg.norm(p - self.nearest(p)) > self.radius else -1
        return abs(flag) != 2 * D

    def calc_intersect(self, f_mid, s_mid):
        """aux point on the wall along the donor->solid axis and the outward (into the gas) unit normal."""
        r = self.radius
        f_mid = np.asarray(f_mid, dtype=np.float64); s_mid = np.asarray(s_mid, dtype=np.float64)
        c = self.nearest(s_mid)
        D = len(f_mid)
        if D == 2:
            # Circle.jl:52-69.  The reference returns the point relative to the centre (its examples centre the body at
            # the origin); here the centre is added back so that off-origin bodies work.  Identical for centre 0.
            if abs(f_mid[0] - s_mid[0]) < EPS:
                t = np.arccos((f_mid[0] - c[0]) / r)
                ap = np.array([r * np.cos(t), r * np.sin(t) if f_mid[1] > c[1] else -r * np.sin(t)])
            else:
                t = np.arcsin((f_mid[1] - c[1]) / r)
                ap = np.array([r * np.cos(t) if f_mid[0] > c[0] else -r * np.cos(t), r * np.sin(t)])
            return ap + c, ap / r
        # Circle.jl:70-82
        d = int(np.nonzero(np.abs(f_mid - s_mid) > EPS)[0][0])
        rm = f_mid - c
        x, y = rm[(d + 1) % 3], rm[(d + 2) % 3]
        rz = np.sqrt(r * r - x * x - y * y)
        dz = (rz - rm[d]) if abs((rz - rm[d]) / (s_mid[d] - f_mid[d])) < 1 else (-rz - rm[d])
        ap = f_mid.copy(); n = f_mid.copy()
        ap[d] = dz + f_mid[d]; n = rm.copy(); n[d] = dz + rm[d]
        return ap, n / r


def classify(forest, shape) -> np.ndarray:
    """per cell: FLUID / SOLID_GHOST / INSIDE_SOLID"""
    out = np.zeros(forest.n, dtype=np.int32)
    dist = np.min(np.linalg.norm(forest.mid[:, None, :] - shape.centers[None, :, :], axis=2), axis=1)
    for c in np.nonzero(dist <= shape.radius)[0]:
        out[c] = SOLID_GHOST if shape.ghost_cell_flag(forest.mid[c], forest.ds[c]) else INSIDE_SOLID
    return out


# ------------------------------------------------------------------------------------------------ cut cells
def _below_fraction(nv, alpha):
    """volume fraction of the unit cube {x in [0,1]^D : nv.x < alpha}, all nv > 0 (inclusion-exclusion).
    nv [m, D], alpha [m] -> [m]"""
    nv = np.atleast_2d(nv); alpha = np.atleast_1d(alpha)
    D = nv.shape[1]
    tot = np.zeros(len(alpha))
    for k in range(D + 1):
        for sub in itertools.combinations(range(D), k):
            a = alpha - (nv[:, list(sub)].sum(axis=1) if sub else 0.0)
            tot += (-1) ** k * np.where(a > 0, a, 0.0) ** D
    fact = 1.0
    for i in range(1, D + 1):
        fact *= i
    return np.clip(tot / (fact * np.prod(nv, axis=1)), 0.0, 1.0)


def cut_cells(normal, grid):
    """CuttedVelocityCells of one SolidNeighbor: indices of velocity cells crossed by v.n = 0 with their gas-side
    (v.n < 0) and solid-side (v.n > 0) measures.  Degenerate normals (2-D: any |n_i| < 1e-6; 3-D: more than one)
    give an empty list, as in the reference (Immersed_boundary.jl:227, :252-254)."""
    n = np.asarray(normal, dtype=np.float64)
    D = grid.dim
    small = np.abs(n) < 1e-6
    z = np.zeros(0)
    if (D == 2 and small.any()) or (D == 3 and small.sum() > 1):
        return np.zeros(0, np.int32), z, z
    ddu_all = grid.root_ds[None, :] / (2.0 ** grid.level.astype(np.float64))[:, None]
    cand = np.nonzero(np.abs(grid.mid @ n) <= 0.5 * np.linalg.norm(ddu_all, axis=1))[0]
    if len(cand) == 0:
        return np.zeros(0, np.int32), z, z
    act = [d for d in range(D) if not small[d]]
    ddu = ddu_all[cand]
    lo = grid.mid[cand] - 0.5 * ddu
    # map to the unit cube with positive normal components: x_d = (v_d - corner_d)/(+-ddu_d)
    nv = np.abs(n[act])[None, :] * ddu[:, act]
    corner = np.where(n[act][None, :] > 0, lo[:, act], lo[:, act] + ddu[:, act])
    alpha = -(corner @ n[act])                      # v.n < 0  <=>  nv.x < alpha
    frac = _below_fraction(nv, alpha)
    keep = (frac > 0.0) & (frac < 1.0)              # drop cells the plane only touches
    vol = np.prod(ddu, axis=1)
    return cand[keep].astype(np.int32), (frac * vol)[keep], ((1.0 - frac) * vol)[keep]


# ------------------------------------------------------------------------------------------------ host tables
@dataclass
class HostIB:
    """kamr_ib (include/kamr.h)"""
    solid_cell: np.ndarray
    solid_nb_off: np.ndarray
    solid_nb_ids: np.ndarray
    sn_donor: np.ndarray
    sn_solid: np.ndarray
    sn_faceid: np.ndarray
    sn_aux: np.ndarray
    sn_normal: np.ndarray
    sn_bc: np.ndarray
    sn_nb_off: np.ndarray
    sn_nb_ids: np.ndarray
    cvc_off: np.ndarray
    cvc_index: np.ndarray
    cvc_gas_w: np.ndarray
    cvc_solid_w: np.ndarray

    @property
    def n_solid(self):
        return len(self.solid_cell)

    @property
    def n_sn(self):
        return len(self.sn_donor)

    def c_struct_ptr(self):
        s = abi.KamrIB()
        p = abi.ptr
        s.n_solid = self.n_solid
        s.solid_cell = p(self.solid_cell, C.c_int32); s.solid_nb_off = p(self.solid_nb_off, C.c_int32)
        s.solid_nb_ids = p(self.solid_nb_ids, C.c_int32)
        s.n_sn = self.n_sn
        s.sn_donor = p(self.sn_donor, C.c_int32); s.sn_solid = p(self.sn_solid, C.c_int32)
        s.sn_faceid = p(self.sn_faceid, C.c_int32)
        s.sn_aux = p(self.sn_aux, C.c_double); s.sn_normal = p(self.sn_normal, C.c_double)
        s.sn_bc = p(self.sn_bc, C.c_double)
        s.sn_nb_off = p(self.sn_nb_off, C.c_int32); s.sn_nb_ids = p(self.sn_nb_ids, C.c_int32)
        s.cvc_off = p(self.cvc_off, C.c_int32); s.cvc_index = p(self.cvc_index, C.c_int32)
        s.cvc_gas_w = p(self.cvc_gas_w, C.c_double); s.cvc_solid_w = p(self.cvc_solid_w, C.c_double)
        self._keep = s
        return C.pointer(s)
