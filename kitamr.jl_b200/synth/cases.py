"""Synthetic stand-ins for the reference's example configurations.

Each `Case` carries what `Configure`/`Solver`/`Gas` carry in the reference
(src/Solver/Types.jl:121-149,381-409; src/Gas/Types.jl:25-38) plus the generated forest,
velocity grids and initial state.  Seeds: 20260101 + k, RNG = PCG64 keyed by global cell id so
every rank of a split sees the same field.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from .. import abi
from ..model import HostMesh, HostState, build_rank_view
from . import vgrid as vg
from .forest import Forest, partition

SEED_BASE = 20260101


def ref_vhs_vis(Kn, omega, alpha=1.0, T_ref=1.0):
    """src/Gas/Model.jl:5-9"""
    mu0 = 5.0 * (alpha + 1.0) * (alpha + 2.0) * math.sqrt(math.pi) / \
        (4.0 * alpha * (5.0 - 2.0 * omega) * (7.0 - 2.0 * omega)) * Kn
    return mu0 * T_ref ** (0.5 - omega)


@dataclass
class Gas:
    """src/Gas/Types.jl:5-38 (note: the reference calls ref_vhs_vis(Kn, αᵣ, ωᵣ), i.e. with αᵣ in the
    `omega` slot and ωᵣ in the `alpha` slot — reproduced here)."""
    Kn: float = 0.05
    Pr: float = 2 / 3
    K: float = 1.0
    gamma: float = 5 / 3
    omega: float = 0.5
    alpha_r: float = 1.0
    omega_r: float = 0.81
    mu_ref: float = None

    def __post_init__(self):
        if self.mu_ref is None:
            self.mu_ref = ref_vhs_vis(self.Kn, self.alpha_r, self.omega_r)


def get_conserved(prim, gamma):
    prim = np.asarray(prim, dtype=np.float64)
    D = len(prim) - 2
    w = np.zeros(D + 2)
    w[0] = prim[0]
    w[1:1 + D] = prim[0] * prim[1:1 + D]
    w[D + 1] = 0.5 * prim[0] / prim[-1] / (gamma - 1.0) + 0.5 * prim[0] * np.sum(prim[1:1 + D] ** 2)
    return w


def get_prim(w, gamma):
    w = np.asarray(w, dtype=np.float64)
    D = len(w) - 2
    p = np.zeros(D + 2)
    p[0] = w[0]
    p[1:1 + D] = w[1:1 + D] / w[0]
    p[D + 1] = 0.5 * w[0] / (gamma - 1.0) / (w[D + 1] - 0.5 * np.sum(w[1:1 + D] ** 2) / w[0])
    return p


def moments(mid, weight, df):
    """micro_to_macro (lib/KitCore/2D2F.jl:119, 3D1F.jl:109); df [n, ndf]."""
    D = mid.shape[1]
    w = np.zeros(D + 2)
    h = df[:, 0]
    w[0] = np.sum(weight * h)
    for d in range(D):
        w[1 + d] = np.sum(weight * mid[:, d] * h)
    e = np.sum(mid ** 2, axis=1) * h
    if df.shape[1] == 2:
        e = e + df[:, 1]
    w[D + 1] = 0.5 * np.sum(weight * e)
    return w


@dataclass
class Case:
    name: str
    dim: int
    ndf: int
    forest: Forest
    grids: list
    cell_grid: np.ndarray
    bc_type: np.ndarray
    bc_prim: np.ndarray
    gas: Gas
    quadrature: tuple
    vs_trees_num: tuple
    vs_maxlevel: int
    prim_fn: object                  # prim_fn(mid[dim]) -> prim[dim+2]
    seed: int
    cfl: float = 0.4
    marching: int = abi.MARCH_CAIDVM
    flux_type: int = abi.FLUX_CAIDVM
    noise: float = 0.01
    bound_enc: np.ndarray = None
    cell_class: np.ndarray = None    # synth.ib.classify(...) when an immersed boundary is present
    ib_shape: object = None

    # ------------------------------------------------------------------ config
    def config(self, device=0, rank=0, nranks=1, stream=None) -> abi.KamrConfig:
        g = self.gas
        return abi.KamrConfig(self.dim, self.ndf, self.flux_type, self.marching, g.K, g.Pr, g.gamma, g.omega,
                              g.mu_ref, device, rank, nranks, stream)

    def dt(self):
        """Δt_ξ of Status(config), src/Solver/Types.jl:509-520."""
        D = self.dim
        geo, q = self.forest.geometry, self.quadrature
        ds = [(geo[2 * i + 1] - geo[2 * i]) / self.forest.trees_num[i] / 2 ** self.forest.maxlevel for i in range(D)]
        U = [max(q[2 * i + 1], abs(q[2 * i])) - (q[2 * i + 1] - q[2 * i]) / self.vs_trees_num[i] /
             2 ** self.vs_maxlevel / 2 for i in range(D)]
        return self.cfl * min(a / b for a, b in zip(ds, U))

    # ------------------------------------------------------------------ partition / flatten
    partition_mode: str = "reference"   # "reference": partition_weight of the reference; "cost": device cost model

    def owner(self, nranks):
        """Weighted split of the Morton curve.  "reference": partition_weight (Parallel/Partition.jl:213-224): vs_num per
        cell, x2 for solid ghost cells; cells inside the body carry no velocity grid.  "cost": the weight function a
        GPU-aware shim hands to the reference's `partition!(p4est, weight)` hook (Partition.jl:228) instead: vs_num
        times the measured relative cost of the kernel class the cell falls into (regular 1, neighbour on
        another velocity grid 2.1, general 3, solid ghost cell 4, +5.5 per solid face of a donor)."""
        n_of = np.array([g.n for g in self.grids])[self.cell_grid].astype(np.float64)
        if self.cell_class is not None:
            n_of = np.where(self.cell_class == -2, 0.0, np.where(self.cell_class == -1, 2.0 * n_of, n_of))
        if self.partition_mode == "cost":
            n_of = n_of * self.cost_factors()
        return partition(n_of, nranks)

    def cost_factors(self):
        """relative device cost per velocity point of every forest cell (1 = regular cell), from the face list of the
        whole forest: which kernel class the cell will take (kamr_lib.cu build_topology)"""
        if getattr(self, "_cost", None) is not None:
            return self._cost
        full = build_rank_view(self.forest, self.grids, self.cell_grid, self.bc_type, self.bc_prim, self.ndf,
                               owner=None, rank=0, bound_enc_global=self.bound_enc, cell_class=self.cell_class,
                               ib_shape=self.ib_shape)
        nl = full.n_local
        kind, here, there = full.face_kind, full.face_here, full.face_there
        grid, be = full.cell_grid, full.bound_enc
        general = np.zeros(nl, bool); mapped = np.zeros(nl, bool); nsolid = np.zeros(nl)
        dom = kind == 0
        general[here[dom]] = True
        h, t, k = here[~dom], there[~dom], kind[~dom]
        for a, b in ((h, t), (t, h)):
            m = a < nl
            aa, bb, kk = a[m], b[m], k[m]
            sol = (bb >= nl) | (be[np.minimum(bb, len(be) - 1)] < 0)
            np.logical_or.at(general, aa, sol | (kk != 1))
            np.add.at(nsolid, aa, sol.astype(np.float64))
            bb_ok = np.minimum(bb, len(grid) - 1)
            np.logical_or.at(mapped, aa, (~sol) & (grid[aa] != grid[bb_ok]))
        f_local = np.where(general, 3.0, np.where(mapped, 2.1, 1.0)) + 5.5 * nsolid
        f_local = np.where(be[:nl] < 0, 2.0, f_local)        # solid ghost cells: 2 x vs_num above, x2 here = 4
        cost = np.ones(self.forest.n)
        cost[full.global_ids[:nl]] = f_local
        self._cost = cost
        return cost

    def rank_mesh(self, rank=0, nranks=1) -> HostMesh:
        owner = self.owner(nranks) if nranks > 1 else None
        return build_rank_view(self.forest, self.grids, self.cell_grid, self.bc_type, self.bc_prim, self.ndf,
                               owner=owner, rank=rank, bound_enc_global=self.bound_enc,
                               cell_class=self.cell_class, ib_shape=self.ib_shape)

    # ------------------------------------------------------------------ state
    def cell_df(self, gid):
        g = self.grids[int(self.cell_grid[gid])]
        prim = np.asarray(self.prim_fn(self.forest.mid[gid]), dtype=np.float64)
        df = vg.discrete_maxwell(g.mid, prim, self.ndf, self.gas.K)
        if self.noise:
            rng = np.random.Generator(np.random.PCG64([self.seed, int(gid)]))
            df = df * (1.0 + self.noise * rng.uniform(-1.0, 1.0, size=df.shape))
        return g, df

    def init_state(self, mesh: HostMesh) -> HostState:
        st = HostState.zeros(mesh)
        D, K, M = self.dim, self.ndf, self.dim + 2
        off = mesh.vs_off()
        for c in range(mesh.n_local + mesh.n_ghost):
            g, df = self.cell_df(int(mesh.global_ids[c]))
            st.df[off[c] * K: off[c + 1] * K] = df.T.ravel()
            w = moments(g.mid, g.weight, df)
            st.w[c * M:(c + 1) * M] = w
            st.prim[c * M:(c + 1) * M] = get_prim(w, self.gas.gamma)
        return st


# ---------------------------------------------------------------------- field helpers
def smooth_prim(dim, geometry, U0=None, amp=0.2):
    """ρ = 1 + 0.2 sin, U as given, λ = 1/(1 + 0.3 cos) at the cell centre."""
    lo = np.array(geometry[0::2]); hi = np.array(geometry[1::2])
    U0 = np.zeros(dim) if U0 is None else np.asarray(U0, dtype=np.float64)

    def fn(x):
        s = 2 * np.pi * (np.asarray(x) - lo) / (hi - lo)
        rho = 1.0 + amp * np.prod(np.sin(s))
        lam = 1.0 / (1.0 + 0.3 * np.prod(np.cos(s)))
        return np.concatenate([[rho], U0 + 0.1 * np.sin(s), [lam]])
    return fn


def _bcs(dim, kinds, prims):
    M = dim + 2
    bt = np.array(kinds, dtype=np.int32)
    bp = np.zeros((2 * dim, M))
    for i, p in enumerate(prims):
        if p is not None:
            bp[i] = p
        else:
            bp[i] = [1.0] + [0.0] * dim + [1.0]
    return bt, bp.ravel()


def _dedup_grids(per_cell_grids):
    keys = {}
    grids = []
    cell_grid = np.zeros(len(per_cell_grids), dtype=np.int32)
    for c, g in enumerate(per_cell_grids):
        k = g.key()
        if k not in keys:
            keys[k] = len(grids)
            grids.append(g)
        cell_grid[c] = keys[k]
    return grids, cell_grid


# ---------------------------------------------------------------------- cases
def smoke_s0(noise=0.01, trees=16, vtrees=16, tree_order="morton") -> Case:
    """test/runtests.jl:4-41: 2-D, NDF=2, 16x16 cells, 16x16 velocity points, Maxwellian walls in x
    (xmax wall moving with U=sqrt(5/6)), periodic in y, CAIDVM + CAIDVM_Marching."""
    geo = (-0.5, 0.5, -0.5, 0.5)
    forest = Forest.build(2, geo, (trees, trees), 0, periodic=(False, True), tree_order=tree_order)
    quad = (-5.0, 5.0, -5.0, 5.0)
    g = vg.root_grid(quad, (vtrees, vtrees))
    gas = Gas(K=0.0, Kn=0.075, omega=0.81, omega_r=0.81)
    bt, bp = _bcs(2, [abi.BC_MAXWELLIAN, abi.BC_MAXWELLIAN, abi.BC_UNIFORM_OUTFLOW, abi.BC_UNIFORM_OUTFLOW],
                  [[1., 0., 0., 1.], [1., math.sqrt(5 / 6), 0., 1.], None, None])
    return Case("S0-smoke", 2, 2, forest, [g], np.zeros(forest.n, np.int32), bt, bp, gas, quad, (vtrees, vtrees), 0,
                lambda x: np.array([1., 0., 0., 1.]), SEED_BASE, noise=noise)


def amr_case(dim=2, trees=4, maxlevel=2, vtrees=6, vs_maxlevel=2, ragged=True, periodic=None, bcs=None,
             seed=1, name=None, marching=abi.MARCH_CAIDVM, U0=None, quad_half=5.0, refine="ball",
             tree_order="morton") -> Case:
    """Small AMR case used by the parity tests: a refined ball/band in physical space (hanging faces,
    coarse/fine slope sweep) and, when `ragged`, Maxwellian-adapted velocity grids that differ from
    cell to cell (pair-list path)."""
    ndf = 2 if dim == 2 else 1
    geo = tuple([-0.5, 0.5] * dim)
    periodic = periodic if periodic is not None else (False,) * dim

    def refine_fn(l, mid, ds):
        if refine == "ball":
            r = np.sqrt(np.sum((mid - 0.1) ** 2, axis=1))
            return r < 0.42 / (l + 1)
        if refine == "band":
            return np.abs(mid[:, 0] + 0.5 * mid[:, 1]) < 0.25 / (l + 1)
        return np.zeros(len(mid), dtype=bool)

    forest = Forest.build(dim, geo, (trees,) * dim, maxlevel, refine_fn, periodic=periodic, tree_order=tree_order)
    quad = tuple([-quad_half, quad_half] * dim)
    gas = Gas(K=1.0 if dim == 2 else 0.0, Kn=0.05)
    prim_fn = smooth_prim(dim, geo, U0=U0)
    if ragged:
        per_cell = []
        cache = {}
        for c in range(forest.n):
            p = prim_fn(forest.mid[c])
            # quantise the prim that drives the grid so that neighbouring cells share grids in
            # patches but differ across patch borders (mimics the state jumps of rp_2D.jl:11-24)
            key = tuple(np.round(p * 4) / 4)
            if key not in cache:
                cache[key] = vg.maxwellian_grid(quad, (vtrees,) * dim, vs_maxlevel, np.array(key), ndf, gas.K)
            per_cell.append(cache[key])
        grids, cell_grid = _dedup_grids(per_cell)
    else:
        grids = [vg.maxwellian_grid(quad, (vtrees,) * dim, vs_maxlevel,
                                    np.array([1.0] + [0.0] * dim + [1.0]), ndf, gas.K)]
        cell_grid = np.zeros(forest.n, np.int32)
    if bcs is None:
        kinds = [abi.BC_SUPERSONIC_INFLOW, abi.BC_UNIFORM_OUTFLOW, abi.BC_MAXWELLIAN,
                 abi.BC_INTERPOLATED_OUTFLOW if dim == 2 else abi.BC_UNIFORM_OUTFLOW] + \
                ([abi.BC_UNIFORM_OUTFLOW, abi.BC_MAXWELLIAN] if dim == 3 else [])
        prims = [[1.0] + [0.3] + [0.0] * (dim - 1) + [1.0], None, [1.0] + [0.0] * dim + [0.9], None] + \
                ([None, [1.0] + [0.1] * dim + [1.1]] if dim == 3 else [])
        bcs = _bcs(dim, kinds, prims)
    return Case(name or f"amr{dim}d", dim, ndf, forest, grids, cell_grid, bcs[0], bcs[1], gas, quad,
                (vtrees,) * dim, vs_maxlevel, prim_fn, SEED_BASE + seed, marching=marching)


def uniform_case(dim=2, trees=64, maxlevel=0, vtrees=60, name=None, seed=3, refine_fn=None,
                 geometry=None, quadrature=None, U0=None, tree_order="morton") -> Case:
    """Bench-shaped case: one shared uniform velocity grid (example/airfoil: 60x60, VS level 0) on a
    (optionally geometry-refined) physical mesh, supersonic inflow at xmin and outflow elsewhere."""
    ndf = 2 if dim == 2 else 1
    geo = geometry or tuple([-0.5, 0.5] * dim)
    trees_t = trees if isinstance(trees, tuple) else (trees,) * dim
    forest = Forest.build(dim, geo, trees_t, maxlevel, refine_fn, tree_order=tree_order)
    quad = quadrature or tuple([-6.0, 6.0] * dim)
    vt = vtrees if isinstance(vtrees, tuple) else (vtrees,) * dim
    g = vg.root_grid(quad, vt)
    gas = Gas(K=1.0 if dim == 2 else 0.0, Kn=0.05)
    kinds = [abi.BC_SUPERSONIC_INFLOW] + [abi.BC_UNIFORM_OUTFLOW] * (2 * dim - 1)
    U = [0.5] + [0.0] * (dim - 1) if U0 is None else list(U0)
    prims = [[1.0] + U + [1.0]] + [None] * (2 * dim - 1)
    bt, bp = _bcs(dim, kinds, prims)
    return Case(name or f"uniform{dim}d", dim, ndf, forest, [g], np.zeros(forest.n, np.int32), bt, bp, gas, quad,
                vt, 0, smooth_prim(dim, geo, U0=U), SEED_BASE + seed)


# ---------------------------------------------------------------------- bench workloads
def cylinder_s2(copies=1, ps_maxlevel=7, box_level=4, vs_maxlevel=3, vtrees=16, trees=25, noise=0.01,
                ib=False, Ma=5.0, tree_order="morton") -> Case:
    """S2 cylinder2d (example/cylinder/cylinder.jl:5-47): 25x25 roots on [-16,16]^2, level `box_level`
    in max-norm(x)<5 (the converged dynamic-AMR region, cylinder_udf.jl:9-15), level `ps_maxlevel`
    within search_coeffi*ds_min = 4*ds_min of the r=1 circle; velocity grids 16x16 roots on
    [-10,10]^2 refined to level<=3 by maxwellian_refine_flag of the buffer IC (cylinder_udf.jl:16-27);
    Ma 5 inflow at xmin, UniformOutflow elsewhere; K=1, Kn=0.1, omega=0.81.

    `copies` > 1 places that many cylinders side by side in x (one connected forest of
    25*copies x 25 roots): the weak-scaling workload, one cylinder's worth of work per GPU.
    """
    W = 32.0
    geo = (-16.0, -16.0 + W * copies, -16.0, 16.0)
    centers = np.array([[W * k, 0.0] for k in range(copies)])
    ds_min = W / trees / 2 ** ps_maxlevel

    def rel(mid):
        d = mid[:, None, :] - centers[None, :, :]
        k = np.argmin(np.abs(d[:, :, 0]), axis=1)
        return d[np.arange(len(mid)), k]

    def refine_fn(l, mid, ds):
        x = rel(mid)
        r = np.sqrt(np.sum(x ** 2, axis=1))
        if l < box_level:
            return np.max(np.abs(x), axis=1) < 5.0
        half_diag = 0.5 * np.sqrt(np.sum(ds ** 2))
        return np.abs(r - 1.0) < 4.0 * ds_min + half_diag

    forest = Forest.build(2, geo, (trees * copies, trees), ps_maxlevel, refine_fn, tree_order=tree_order)
    quad = (-10.0, 10.0, -10.0, 10.0)
    gas = Gas(K=1.0, Kn=0.1, omega=0.81, omega_r=0.81)

    def prim_fn(x):
        d = np.asarray(x)[None, :] - centers
        xr = d[np.argmin(np.abs(d[:, 0]))]
        r = float(np.sqrt(np.sum(xr ** 2)))
        if r > 2.0:
            return np.array([1.0, Ma * math.sqrt(5 / 6), 0.0, 1.0])
        rr = max(r, 1.0)  # cells inside the body (no immersed boundary in this variant) take the wall state
        return np.array([1.0, (rr - 1.0) * Ma * math.sqrt(5 / 6), 0.0, 1.0])

    cache = {}
    per_cell = []
    for c in range(forest.n):
        p = prim_fn(forest.mid[c])
        key = (round(float(p[1]), 1),)
        if key not in cache:
            cache[key] = vg.maxwellian_grid(quad, (vtrees, vtrees), vs_maxlevel, np.array([1.0, key[0], 0.0, 1.0]),
                                            2, gas.K)
        per_cell.append(cache[key])
    grids, cell_grid = _dedup_grids(per_cell)
    bt, bp = _bcs(2, [abi.BC_SUPERSONIC_INFLOW] + [abi.BC_UNIFORM_OUTFLOW] * 3,
                  [[1.0, Ma * math.sqrt(5 / 6), 0.0, 1.0], None, None, None])
    case = Case(f"S2-cylinder2d{'-ib' if ib else ''}-x{copies}", 2, 2, forest, grids, cell_grid, bt, bp, gas, quad,
                (vtrees, vtrees), vs_maxlevel, prim_fn, SEED_BASE + 2, noise=noise)
    if ib:   # IB = [Circle(Maxwellian,[0.,0.],1.,true,4.0,[1.,0.,0.,1.])], example/cylinder/cylinder.jl:41
        from . import ib as ibm
        case.ib_shape = ibm.Ball(centers, 1.0, np.array([1.0, 0.0, 0.0, 1.0]))
        case.cell_class = ibm.classify(forest, case.ib_shape)
    return case


def sphere_s4(copies=1, ps_maxlevel=4, trees=16, vtrees=16, vs_maxlevel=2, noise=0.01, ib=True,
              shell=1.5, Ma=3.834) -> Case:
    """S4 sphere3d (example/sphere/sphere.jl:5-47, sphere_udf.jl): trees^3 roots on [-4,4]^3, static refinement
    to level L-2 in max-norm(x) < 1.2 and L-1 in < 0.8 outside the body (shock_wave_region), level L within
    search_coeffi*ds_min = 1.5*ds_min of the r = 0.5 sphere; velocity grids vtrees^3 roots on [-7.28,7.28]^3 refined to
    level <= 2 by maxwellian_refine_flag of the buffer IC (the stored static_vs_refine_flag is never called in the
    reference); NDF = 1, K = 0, Ma 3.834 inflow at xmin, UniformOutflow elsewhere, isothermal
    Maxwellian sphere.  `copies` places that many spheres side by side in x (weak scaling)."""
    W = 8.0
    geo = (-4.0, -4.0 + W * copies, -4.0, 4.0, -4.0, 4.0)
    centers = np.array([[W * k, 0.0, 0.0] for k in range(copies)])
    L = ps_maxlevel
    ds_min = W / trees / 2 ** L
    Tw = 1.0 + (5 / 3 - 1) * 0.5 * Ma ** 2

    def rel(mid):
        d = mid[:, None, :] - centers[None, :, :]
        k = np.argmin(np.abs(d[:, :, 0]), axis=1)
        return d[np.arange(len(mid)), k]

    def refine_fn(l, mid, ds):
        x = rel(mid)
        r = np.sqrt(np.sum(x ** 2, axis=1))
        mx = np.max(np.abs(x), axis=1)
        out = np.zeros(len(mid), dtype=bool)
        out |= (mx < 1.2) & (r > 0.5) & (l < L - 2)
        out |= (mx < 0.8) & (r > 0.5) & (l < L - 1)
        half_diag = 0.5 * np.sqrt(np.sum(ds ** 2))
        out |= np.abs(r - 0.5) < shell * ds_min + half_diag
        return out

    forest = Forest.build(3, geo, (trees * copies, trees, trees), L, refine_fn)
    quad = (-7.28, 7.28) * 3
    gas = Gas(K=0.0, Kn=0.03, omega=0.75, omega_r=0.75, mu_ref=5.0 * math.sqrt(math.pi) / 16.0 * 0.03)

    def prim_fn(x):
        d = np.asarray(x)[None, :] - centers
        xr = d[np.argmin(np.abs(d[:, 0]))]
        r = float(np.sqrt(np.sum(xr ** 2)))
        if r > 1.0:
            return np.array([1.0, Ma * math.sqrt(5 / 6), 0.0, 0.0, 1.0])
        rr = max(r, 0.5)
        return np.array([1.0, (rr - 0.5) / 0.5 * Ma * math.sqrt(5 / 6), 0.0, 0.0, 1.0 / (Tw - (rr - 0.5) * (Tw - 1.0))])

    cache = {}
    per_cell = []
    for c in range(forest.n):
        p = prim_fn(forest.mid[c])
        key = (round(float(p[1]), 0), round(float(p[4]), 1))
        if key not in cache:
            cache[key] = vg.maxwellian_grid(quad, (vtrees,) * 3, vs_maxlevel,
                                            np.array([1.0, key[0], 0.0, 0.0, max(key[1], 0.1)]), 1, gas.K)
        per_cell.append(cache[key])
    grids, cell_grid = _dedup_grids(per_cell)
    bt, bp = _bcs(3, [abi.BC_SUPERSONIC_INFLOW] + [abi.BC_UNIFORM_OUTFLOW] * 5,
                  [[1.0, Ma * math.sqrt(5 / 6), 0.0, 0.0, 1.0]] + [None] * 5)
    case = Case(f"S4-sphere3d-x{copies}", 3, 1, forest, grids, cell_grid, bt, bp, gas, quad, (vtrees,) * 3,
                vs_maxlevel, prim_fn, SEED_BASE + 4, noise=noise)
    if ib:
        from . import ib as ibm
        case.ib_shape = ibm.Ball(centers, 0.5, np.array([1.0, 0.0, 0.0, 0.0, 1.0 / Tw]))
        case.cell_class = ibm.classify(forest, case.ib_shape)
    return case


def x38like_s5(ps_maxlevel=4, trees=16, vtrees=10, vs_maxlevel=3, noise=0.01, radius=0.5, shell=0.3) -> Case:
    """S5 x38-like3d (example/X38/X38_julia.jl:12-58): trees^3 roots on [-4,5]x[-4,4]^2, level L-2 in max-norm(x) < 2,
    L-1 in < 1.2, level L within `shell` of the body surface; velocity grids vtrees^3 roots on the X38 quadrature box
    [-14.2,21.3]x[-17.75,17.75]^2 refined to level <= 3 by maxwellian_refine_flag of the buffer IC (1 000 points in the
    free stream, tens of thousands in the shock layer); NDF = 1, K = 0, Kn = 0.275, Ma 8 inflow at xmin,
    UniformOutflow elsewhere, cold Maxwellian wall (lambda_w = 56/300 as in the reference's IB record).
    The reference's body is a triangulated STL (`X38_normalized.stl`, missing from the repository); the surrogate body
    here is a sphere of radius `radius` at the origin — same cell classes (solid ghost cells, donors, cut velocity
    cells), different shape."""
    geo = (-4.0, 5.0, -4.0, 4.0, -4.0, 4.0)
    centers = np.array([[0.0, 0.0, 0.0]])
    L = ps_maxlevel
    Ma = 8.0
    lam_w = 56.0 / 300.0

    def refine_fn(l, mid, ds):
        r = np.sqrt(np.sum(mid ** 2, axis=1))
        mx = np.max(np.abs(mid), axis=1)
        out = np.zeros(len(mid), dtype=bool)
        out |= (mx < 2.0) & (r > radius) & (l < L - 2)
        out |= (mx < 1.2) & (r > radius) & (l < L - 1)
        half_diag = 0.5 * np.sqrt(np.sum(ds ** 2))
        out |= np.abs(r - radius) < shell + half_diag
        return out

    forest = Forest.build(3, geo, (trees,) * 3, L, refine_fn)
    quad = (-14.2, 21.3, -17.75, 17.75, -17.75, 17.75)
    gas = Gas(K=0.0, Kn=0.275, omega=0.81, omega_r=0.81)

    def prim_fn(x):
        r = float(np.sqrt(np.sum(np.asarray(x) ** 2)))
        if r > 2.0 * radius:
            return np.array([1.0, Ma * math.sqrt(5 / 6), 0.0, 0.0, 1.0])
        rr = max(r, radius)
        t = (rr - radius) / radius
        return np.array([1.0, t * Ma * math.sqrt(5 / 6), 0.0, 0.0, lam_w + t * (1.0 - lam_w)])

    cache = {}
    per_cell = []
    for c in range(forest.n):
        p = prim_fn(forest.mid[c])
        key = (round(float(p[1]), 0), round(float(p[4]), 1))
        if key not in cache:
            cache[key] = vg.maxwellian_grid(quad, (vtrees,) * 3, vs_maxlevel,
                                            np.array([1.0, key[0], 0.0, 0.0, max(key[1], 0.1)]), 1, gas.K)
        per_cell.append(cache[key])
    grids, cell_grid = _dedup_grids(per_cell)
    bt, bp = _bcs(3, [abi.BC_SUPERSONIC_INFLOW] + [abi.BC_UNIFORM_OUTFLOW] * 5,
                  [[1.0, Ma * math.sqrt(5 / 6), 0.0, 0.0, 1.0]] + [None] * 5)
    case = Case("S5-x38like3d", 3, 1, forest, grids, cell_grid, bt, bp, gas, quad, (vtrees,) * 3, vs_maxlevel,
                prim_fn, SEED_BASE + 5, noise=noise)
    from . import ib as ibm
    case.ib_shape = ibm.Ball(centers, radius, np.array([1.0, 0.0, 0.0, 0.0, lam_w]))
    case.cell_class = ibm.classify(forest, case.ib_shape)
    return case


def riemann_s1(ps_level=4, band_level=5, trees=16, vtrees=8, vs_maxlevel=3, noise=0.01,
               marching=abi.MARCH_CIP) -> Case:
    """S1 riemann2d (example/Riemann_problem/rp_2D.jl:26-72, BASELINE.json configs[0]): 16x16 roots on [-.5,.5]^2,
    uniform level `ps_level` with a diagonal band refined to `band_level` (hanging faces ~10 %); velocity grids 8x8
    roots on [-5,5]^2 refined to level <= 3 by maxwellian_refine_flag of each cell's own quadrant state
    (Riemann_2D_init :11-24), so grids differ across the state jumps; 4x UniformOutflow, K=1, Kn=1e-3, omega=0.81,
    CIP_Marching (:37)."""
    geo = (-0.5, 0.5, -0.5, 0.5)
    ds_band = 1.0 / trees / 2 ** ps_level

    def refine_fn(l, mid, ds):
        if l < ps_level:
            return np.ones(len(mid), dtype=bool)
        return np.abs(mid[:, 0] + mid[:, 1]) < 1.5 * ds_band

    forest = Forest.build(2, geo, (trees, trees), band_level, refine_fn)
    quad = (-5.0, 5.0, -5.0, 5.0)
    gas = Gas(K=1.0, Kn=1e-3, omega=0.81, omega_r=0.81)

    def state(rho, u, v, p):
        return np.array([rho, u, v, 0.5 * rho / p])

    def prim_fn(x):
        if x[0] >= 0.0 and x[1] >= 0.0:
            return state(0.5313, 0.0, 0.0, 0.4)
        if x[0] < 0.0 and x[1] >= 0.0:
            return state(1.0, 0.7276, 0.0, 1.0)
        if x[0] < 0.0 and x[1] < 0.0:
            return state(0.8, 0.0, 0.0, 1.0)
        return state(1.0, 0.0, 0.7276, 1.0)

    cache = {}
    per_cell = []
    for c in range(forest.n):
        key = tuple(prim_fn(forest.mid[c]))
        if key not in cache:
            cache[key] = vg.maxwellian_grid(quad, (vtrees, vtrees), vs_maxlevel, np.array(key), 2, gas.K)
        per_cell.append(cache[key])
    grids, cell_grid = _dedup_grids(per_cell)
    bt, bp = _bcs(2, [abi.BC_UNIFORM_OUTFLOW] * 4, [None] * 4)
    tag = {abi.MARCH_CIP: "cip", abi.MARCH_CAIDVM: "caidvm", abi.MARCH_EULER: "euler"}[marching]
    return Case(f"S1-riemann2d-{tag}", 2, 2, forest, grids, cell_grid, bt, bp, gas, quad, (vtrees, vtrees),
                vs_maxlevel, prim_fn, SEED_BASE + 1, noise=noise, marching=marching)


def _naca0012(x):
    """half thickness of a unit-chord NACA0012 at x in [0,1] (the reference's naca0012.csv is missing; analytic)"""
    x = np.clip(x, 0.0, 1.0)
    return 0.6 * (0.2969 * np.sqrt(x) - 0.1260 * x - 0.3516 * x ** 2 + 0.2843 * x ** 3 - 0.1015 * x ** 4)


def airfoil_s3(ps_maxlevel=7, box_level=4, trees=(24, 32), vtrees=60, noise=0.01) -> Case:
    """S3 airfoil2d (example/airfoil/airfoil.jl:5-46): 24x32 roots on [-3,7]x[-7,7], level `box_level` in the box
    (-1,3)x(-1,1), level `ps_maxlevel` within 4.5 ds_min of a unit-chord NACA0012 (analytic; no immersed boundary in
    this synthetic variant: the polygon IB needs the missing csv); ONE uniform 60x60 velocity grid on [-4,8]x[-6,6]
    (n = 3600 for every cell: the identical-grid fast path only); Ma 2 inflow at xmin, 3x InterpolatedOutflow."""
    geo = (-3.0, 7.0, -7.0, 7.0)
    ds_min = min(10.0 / trees[0], 14.0 / trees[1]) / 2 ** ps_maxlevel
    Ma = 2.0

    def refine_fn(l, mid, ds):
        x, y = mid[:, 0], mid[:, 1]
        if l < box_level:
            return (x > -1.0) & (x < 3.0) & (np.abs(y) < 1.0)
        half_diag = 0.5 * np.sqrt(np.sum(ds ** 2))
        dist = np.abs(np.abs(y) - _naca0012(x))
        dist = np.where((x < 0.0) | (x > 1.0), np.sqrt(np.minimum(x ** 2, (x - 1.0) ** 2) + y ** 2), dist)
        return dist < 4.5 * ds_min + half_diag

    forest = Forest.build(2, geo, trees, ps_maxlevel, refine_fn)
    quad = (-4.0, 8.0, -6.0, 6.0)
    gas = Gas(K=1.0, Kn=0.05, omega=0.81, omega_r=0.81)
    g = vg.root_grid(quad, (vtrees, vtrees))
    inflow = [1.0, Ma * math.sqrt(5 / 6), 0.0, 1.0]
    bt, bp = _bcs(2, [abi.BC_SUPERSONIC_INFLOW] + [abi.BC_INTERPOLATED_OUTFLOW] * 3, [inflow, None, None, None])
    return Case("S3-airfoil2d", 2, 2, forest, [g], np.zeros(forest.n, np.int32), bt, bp, gas, quad, (vtrees, vtrees),
                0, smooth_prim(2, geo, U0=inflow[1:3]), SEED_BASE + 3, noise=noise)


WORKLOADS = {
    "S2": cylinder_s2,
    "S2ib": lambda copies=1: cylinder_s2(copies=copies, ib=True),
    # the same cases with the dynamically refined box one level finer (54 436 / 45 936 cells): the cell counts SURVEY.md
    # §8d estimates for S2 / S3 (6e4 / 1e5).  The velocity spaces stay the configs' own (16x16 roots L<=3 -> mean vs_num
    # 928, where the survey guessed ~1500; 60x60), so the phase-space sizes are 5.0e7 and 1.65e8, not 1e8 and 3.6e8.
    "S2ib-big": lambda copies=1: cylinder_s2(copies=copies, ib=True, box_level=5),
    "S3-big": lambda copies=1: airfoil_s3(box_level=5),
    "S4": sphere_s4,
    "S1": lambda copies=1: riemann_s1(),
    "S1caidvm": lambda copies=1: riemann_s1(marching=abi.MARCH_CAIDVM),
    "S3": lambda copies=1: airfoil_s3(),
    "S5": lambda copies=1: x38like_s5(),
}
