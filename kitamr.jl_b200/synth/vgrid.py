"""Synthetic adaptive velocity grids in the reference's storage order.

Follows src/Velocity_space/Initialize.jl:9-37 (root grid, x fastest), Rebuild.jl:44-85 (a refined
cell is replaced in place by its 2^DIM children in RMT order, weight/2^DIM) and the analytic
initial refine flag of Criteria.jl:123-203 (`maxwellian_refine_flag`, contribution branch).
All merge-walks of the reference (Flux/Slope.jl:29, Flux/Flux.jl:151) rely on this depth-first
order on both grids.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass

import numpy as np
from scipy.special import erf

MAXWELLIAN_INIT_FLOOR = 1e-3  # Criteria.jl:121


def rmt(dim):
    """RMT[DIM]: child offsets, x fastest (src/Abstract/Types.jl:10)."""
    return np.array([o[::-1] for o in itertools.product(*[(-1, 1)] * dim)], dtype=np.float64)


@dataclass
class VGrid:
    dim: int
    level: np.ndarray    # [n] int8
    weight: np.ndarray   # [n]
    mid: np.ndarray      # [n, dim]
    root_ds: np.ndarray  # [dim]

    @property
    def n(self):
        return len(self.level)

    def key(self):
        return (self.level.tobytes(), self.mid.tobytes())


def root_grid(quadrature, trees_num) -> VGrid:
    dim = len(trees_num)
    ds = np.zeros(dim)
    axes = []
    for i in range(dim):
        half = (quadrature[2 * i + 1] - quadrature[2 * i]) / (2 * trees_num[i])
        ds[i] = 2 * half
        axes.append(np.linspace(quadrature[2 * i] + half, quadrature[2 * i + 1] - half, trees_num[i]))
    grids = np.meshgrid(*axes, indexing="ij")
    # Base.Iterators.product: first axis fastest
    mid = np.stack([g.ravel(order="F") for g in grids], axis=1)
    n = len(mid)
    return VGrid(dim, np.zeros(n, dtype=np.int8), np.full(n, np.prod(ds)), mid, ds)


def refine(grid: VGrid, flags: np.ndarray) -> VGrid:
    """refine_grid_stream!, Rebuild.jl:44-85."""
    dim = grid.dim
    nc = 2 ** dim
    if not flags.any():
        return grid
    counts = np.where(flags, nc, 1)
    starts = np.concatenate([[0], np.cumsum(counts)])
    nnew = int(starts[-1])
    src = np.repeat(np.arange(grid.n), counts)
    child = np.arange(nnew) - starts[src]
    level = grid.level[src].astype(np.int8)
    weight = grid.weight[src].copy()
    mid = grid.mid[src].copy()
    ref = flags[src]
    L = grid.level[src][ref].astype(np.int64)
    level[ref] = (L + 1).astype(np.int8)
    weight[ref] = weight[ref] / nc
    table = rmt(dim)
    mid[ref] = mid[ref] + 0.5 * (grid.root_ds[None, :] / (2.0 ** (L + 1))[:, None]) * table[child[ref]]
    return VGrid(dim, level, weight, mid, grid.root_ds)


def _maxwellian_1d(c, du, lam):
    """Criteria.jl:124-129"""
    sq = np.sqrt(lam); a = c - 0.5 * du; b = c + 0.5 * du
    I0 = 0.5 * np.sqrt(np.pi / lam) * (erf(sq * b) - erf(sq * a))
    I2 = (a * np.exp(-lam * a ** 2) - b * np.exp(-lam * b ** 2)) / (2 * lam) + I0 / (2 * lam)
    return I0, I2


def maxwellian_refine_flag(grid: VGrid, prim, ndf, K) -> np.ndarray:
    """Criteria.jl:165-203 without the :lohner branch."""
    dim = grid.dim
    rho, lam = prim[0], prim[-1]
    U = np.asarray(prim[1:1 + dim])
    du = grid.root_ds[None, :] / (2.0 ** grid.level.astype(np.float64))[:, None]
    I0 = np.zeros((grid.n, dim)); I2 = np.zeros((grid.n, dim))
    for d in range(dim):
        I0[:, d], I2[:, d] = _maxwellian_1d(grid.mid[:, d] - U[d], du[:, d], lam)
    pref = rho * (lam / np.pi) ** (dim / 2)
    drho = pref * np.prod(I0, axis=1)
    dc2 = np.zeros(grid.n)
    for d in range(dim):
        p = np.ones(grid.n)
        for e in range(dim):
            if e != d:
                p = p * I0[:, e]
        dc2 += I2[:, d] * p
    dc2 *= pref
    if ndf == 2:
        dE = 0.5 * (dc2 + (K / (2 * lam)) * drho)
        Eint = rho * (dim + K) / (4 * lam)
    else:
        dE = 0.5 * dc2
        Eint = rho * dim / (4 * lam)
    return np.maximum(drho / rho, dE / Eint) > MAXWELLIAN_INIT_FLOOR


def maxwellian_grid(quadrature, trees_num, maxlevel, prim, ndf, K) -> VGrid:
    """initialize_vs_data, Velocity_space/Initialize.jl:9-37."""
    g = root_grid(quadrature, trees_num)
    for _ in range(maxlevel):
        g = refine(g, maxwellian_refine_flag(g, prim, ndf, K))
    return g


def random_grid(quadrature, trees_num, maxlevel, rng, p=0.3) -> VGrid:
    """Randomly refined grid (stress input for the merge-walk / pair-map tests)."""
    g = root_grid(quadrature, trees_num)
    for _ in range(maxlevel):
        g = refine(g, rng.random(g.n) < p)
    return g


def discrete_maxwell(mid, prim, ndf, K):
    """lib/KitCore/2D2F.jl:1-13, 3D1F.jl:1-14 -> [n, ndf]."""
    dim = mid.shape[1]
    lam = prim[-1]
    c2 = np.sum((mid - np.asarray(prim[1:1 + dim])[None, :]) ** 2, axis=1)
    if dim == 2:
        h = prim[0] * (lam / np.pi) * np.exp(-lam * c2)
    else:
        h = prim[0] * (lam / np.pi) ** 1.5 * np.exp(-lam * c2)
    if ndf == 2:
        return np.stack([h, h * K / (2.0 * lam)], axis=1)
    return h[:, None]
