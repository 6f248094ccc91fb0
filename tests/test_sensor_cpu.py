"""CPU tests of the oracle's physical-space adaptation sensor (orc_ps_criterion): update_criterion!(ka) of
Physical_space/AMR.jl:256-341 with the Löhner estimator of Physical_space/Criteria.jl.  Pinned by an independently
written NumPy evaluation on uniform periodic meshes and by the estimator's invariants (zero on linear fields across
hanging faces, one-cell buffer)."""
import numpy as np
import pytest

from kitamr_jl_b200.synth import cases
from oracle import orc


def _prim(w, gamma, D):
    rho = w[..., 0]
    u = w[..., 1:D + 1] / rho[..., None]
    lam = 0.5 * rho / (gamma - 1.0) / (w[..., D + 1] - 0.5 * np.sum(w[..., 1:D + 1] ** 2, axis=-1) / rho)
    return np.concatenate([rho[..., None], u, lam[..., None]], axis=-1)


def _conserved(prim, gamma, D):
    rho = prim[..., 0]
    mom = prim[..., 1:D + 1] * rho[..., None]
    E = 0.5 * rho / prim[..., D + 1] / (gamma - 1.0) + 0.5 * rho * np.sum(prim[..., 1:D + 1] ** 2, axis=-1)
    return np.concatenate([rho[..., None], mom, E[..., None]], axis=-1)


def _periodic_uniform(dim, trees):
    from kitamr_jl_b200.synth.forest import Forest
    c = cases.uniform_case(dim=dim, trees=trees, vtrees=4, tree_order="lex")
    c.forest = Forest.build(dim, c.forest.geometry, c.forest.trees_num, 0, periodic=(True,) * dim, tree_order="lex")
    return c


def _set_fields(case, mesh, st, prim_fn, grad_fn=None):
    """w / prim from a primitive field evaluated at the cell midpoints; sw = the given gradient of w (or zero)"""
    D, M = case.dim, case.dim + 2
    gamma = case.config().gamma
    mid = mesh.mid.reshape(-1, D)
    prim = prim_fn(mid)
    w = _conserved(prim, gamma, D)
    st.w[:] = w.ravel()
    st.prim[:] = _prim(w, gamma, D).ravel()
    st.sw[:] = 0.0
    if grad_fn is not None:
        st.sw[:] = grad_fn(mid).ravel()   # [cell][dir][row]
    return gamma


def _lohner_twin(case, mesh, st, gamma):
    """uniform periodic mesh: every side is one same-size neighbour (dsL = dsR = ds), Criteria.jl:25-200"""
    D, M = case.dim, case.dim + 2
    n = mesh.n_local
    mid = mesh.mid.reshape(-1, D)[:n]
    ds = mesh.ds.reshape(-1, D)[:n]
    lo = mid.min(axis=0)
    idx = np.rint((mid - lo) / ds).astype(int)
    dims = idx.max(axis=0) + 1
    lin = -np.ones(tuple(dims), dtype=int)
    lin[tuple(idx.T)] = np.arange(n)
    w = st.w.reshape(-1, M)[:n]
    prim = st.prim.reshape(-1, M)[:n]
    sw = st.sw.reshape(-1, D, M)[:n]

    def vort(sw_, pr_):
        def vs(comp, d):
            return (sw_[:, d, comp] - pr_[:, comp] * sw_[:, d, 0]) / pr_[:, 0]
        if D == 2:
            return vs(1, 1) - vs(2, 0)
        c1 = vs(2, 2) - vs(3, 1); c2 = vs(3, 0) - vs(1, 2); c3 = vs(1, 1) - vs(2, 0)
        return np.sqrt(c1 * c1 + c2 * c2 + c3 * c3)

    def lval(l, c, r, h, eps):
        scale = h * np.abs(l) + 2 * h * np.abs(c) + h * np.abs(r)
        denom = h * np.abs(l - c) + h * np.abs(r - c) + eps * scale
        val = np.abs(h * l - 2 * h * c + h * r) / np.where(denom > 0, denom, 1.0)
        return np.where((scale < 1e-4 * h) | (denom <= 0), 0.0, val)

    out = np.zeros((n, D, M))
    om = vort(sw, prim)
    for d in range(D):
        sh = np.zeros(D, dtype=int); sh[d] = 1
        L = lin[tuple(((idx - sh) % dims).T)]
        R = lin[tuple(((idx + sh) % dims).T)]
        pL, pR = _prim(w[L], gamma, D), _prim(w[R], gamma, D)
        h = ds[:, d]
        eps = 0.2 * h
        for j in range(M):
            if j == 1:
                oL, oR = vort(sw[L], pL), vort(sw[R], pR)
                omega = np.maximum(np.abs(oL), np.maximum(np.abs(om), np.abs(oR)))
                vscale = np.maximum(np.sqrt(np.sum(prim[:, 1:D + 1] ** 2, axis=1)),
                                    1.0 / np.sqrt(np.maximum(np.abs(prim[:, M - 1]), np.finfo(float).eps)))
                ok = omega * h >= 2e-2 * vscale
                out[:, d, j] = np.where(ok, lval(oL, om, oR, h, eps), 0.0)
            else:
                jump = np.maximum(np.abs(pL[:, j] - prim[:, j]), np.abs(pR[:, j] - prim[:, j]))
                ok = jump >= 1e-3 * np.maximum(np.abs(prim[:, j]), 1e-4)
                out[:, d, j] = np.where(ok, lval(pL[:, j], prim[:, j], pR[:, j], h, eps), 0.0)
    return out


@pytest.mark.parametrize("dim", [2, 3])
def test_lohner_matches_numpy_twin_on_a_uniform_periodic_mesh(dim):
    case = _periodic_uniform(dim, 8 if dim == 2 else 5)
    mesh = case.rank_mesh()
    st = case.init_state(mesh)
    rng = np.random.default_rng(5)
    D, M = dim, dim + 2
    geo = np.asarray(case.forest.geometry, dtype=float).reshape(D, 2)
    span = geo[:, 1] - geo[:, 0]

    def prim_fn(x):
        s = 2 * np.pi * (x - geo[:, 0]) / span
        rho = 1.0 + 0.3 * np.sin(s[:, 0]) * np.cos(s[:, 1])
        u = [0.2 * np.sin(s[:, (k + 1) % D]) for k in range(D)]
        lam = 1.0 + 0.2 * np.cos(s[:, 0] + s[:, D - 1])
        return np.stack([rho] + u + [lam], axis=1)

    gamma = _set_fields(case, mesh, st, prim_fn)
    st.sw[:] = 0.05 * rng.standard_normal(st.sw.shape)   # the vorticity row is a function of sw only
    thr = 10.0   # nothing above: no buffer
    loh, sen, flg = orc.ps_criterion(case.config(), mesh, st, thr)
    tw = _lohner_twin(case, mesh, st, gamma)
    assert np.abs(tw).max() > 0.05                      # the field does exercise the estimator
    assert (tw[:, :, 1] != 0).any()                     # ... including the vorticity row
    assert np.allclose(loh, tw, rtol=1e-12, atol=1e-13)
    assert np.array_equal(sen, np.maximum(loh[:, :, 0].max(axis=1), loh[:, :, M - 1].max(axis=1)))
    assert not flg.any()


@pytest.mark.parametrize("name", ["amr2d", "amr3d"])
def test_lohner_vanishes_on_a_linear_field_across_hanging_faces(name):
    """dsL/dsR = ds, 0.75 ds (finer side: mean of the 2^(D-1) children) or 1.5 ds (coarser side, shifted to the cell's
    own transverse position with the neighbour's sw, Criteria.jl:147-158) make the second difference of a linear
    field vanish on every interior cell — a wrong factor or a missing shift leaves O(1) values."""
    if name == "amr2d":
        case = cases.amr_case(dim=2, trees=4, maxlevel=2, vtrees=4, vs_maxlevel=0, ragged=False, seed=3)
    else:
        case = cases.amr_case(dim=3, trees=3, maxlevel=1, vtrees=4, vs_maxlevel=0, ragged=False, seed=4)
    mesh = case.rank_mesh()
    st = case.init_state(mesh)
    D, M = case.dim, case.dim + 2
    g = np.array([0.11, -0.07, 0.05])[:D]
    lam0 = 1.3
    gamma = case.config().gamma

    def prim_fn(x):   # rho linear, u = 0, lambda constant: w is linear too
        rho = 2.0 + x @ g
        return np.stack([rho] + [np.zeros_like(rho)] * D + [np.full_like(rho, lam0)], axis=1)

    def grad_fn(x):
        sw = np.zeros((x.shape[0], D, M))
        for d in range(D):
            sw[:, d, 0] = g[d]
            sw[:, d, M - 1] = 0.5 * g[d] / lam0 / (gamma - 1.0)
        return sw

    _set_fields(case, mesh, st, prim_fn, grad_fn)
    loh, sen, flg = orc.ps_criterion(case.config(), mesh, st, 0.25)
    nbs = mesh.nb_state.reshape(-1, 2 * D)[:mesh.n_local]
    assert (nbs > 1).any() and (nbs == -1).any()        # the mesh does have hanging faces
    # the amplitude gate (jump >= 1e-3 |center|) lets these through: the values must be rounding noise
    assert np.abs(loh).max() < 1e-9
    assert not flg.any()
    # the same field with the transverse gradient withheld: cells next to a coarser neighbour now see a kink
    st.sw[:] = 0.0
    loh2, _, _ = orc.ps_criterion(case.config(), mesh, st, 0.25)
    coarse_side = (nbs == -1).any(axis=1)
    assert np.abs(loh2[coarse_side]).max() > 1e-3
    assert np.abs(loh2[~coarse_side & (nbs != 0).all(axis=1)]).max() < 1e-9


def test_domain_sides_zero_the_direction_and_solid_cells_are_skipped():
    case = cases.cylinder_s2(trees=5, ps_maxlevel=4, box_level=2, vtrees=8, vs_maxlevel=2, ib=True)
    mesh = case.rank_mesh()
    st = case.init_state(mesh)
    cfg = case.config()
    orc.step(cfg, mesh, st, case.dt())
    orc.slope(cfg, mesh, st)
    D, M = 2, 4
    loh, sen, flg = orc.ps_criterion(cfg, mesh, st, 0.25)
    assert np.isfinite(loh).all()
    nbs = mesh.nb_state.reshape(-1, 2 * D)[:mesh.n_local]
    be = mesh.bound_enc[:mesh.n_local]
    for d in range(D):
        edge = (nbs[:, 2 * d] == 0) | (nbs[:, 2 * d + 1] == 0)
        unbuffered = sen != 0.5
        assert (loh[edge & unbuffered][:, d, :] == 0).all()
    assert (loh[be < 0] == 0).all() and not flg[be < 0].any()
    # donor cells see the SolidNeighbor's zero state as a density jump (Immersed_boundary.jl:300-302: w = sw = 0);
    # the NaN primitives of that side switch the other rows off instead of poisoning them
    donors = be > 0
    assert donors.any() and flg[donors].mean() > 0.5


def test_one_cell_buffer():
    case = _periodic_uniform(2, 10)
    mesh = case.rank_mesh()
    st = case.init_state(mesh)
    D, M = 2, 4
    geo = np.asarray(case.forest.geometry, dtype=float).reshape(D, 2)
    xm = 0.5 * (geo[0, 0] + geo[0, 1])

    def prim_fn(x):   # a density step in x
        rho = np.where(x[:, 0] < xm, 1.0, 2.0)
        return np.stack([rho, 0 * rho, 0 * rho, np.ones_like(rho)], axis=1)

    _set_fields(case, mesh, st, prim_fn)
    thr = 0.25
    loh, sen, flg = orc.ps_criterion(case.config(), mesh, st, thr)
    n = mesh.n_local
    assert 0 < flg.sum() < n
    # expected buffer: unflagged cells with a flagged face neighbour
    nb_off = mesh.nb_off; nb_ids = mesh.nb_ids
    want = np.zeros(n, dtype=bool)
    for c in range(n):
        if flg[c]:
            continue
        ids = nb_ids[nb_off[c * 2 * D]:nb_off[(c + 1) * 2 * D]]
        want[c] = flg[ids[ids < n]].any()
    assert want.any()
    assert (loh[want] == 2 * thr).all() and (sen[want] == 2 * thr).all()
    rest = ~want & (flg == 0)
    assert (sen[rest] <= thr).all()
    assert (sen[flg == 1] > thr).all()
    # ghost flags: a single raised ghost flag would inflate exactly its local face neighbours — no ghosts here
    assert mesh.n_ghost == 0


@pytest.mark.parametrize("world", [2, 3])
def test_rank_partitions_with_ghost_flags_reproduce_the_single_rank_sensor(world):
    """the multi-rank contract of kamr_ps_criterion restated on the CPU: every rank evaluates its cells with the
    ghosts' w / sw filled from their owners and the owners' pre-buffer decisions as ghost flags
    (lohner_flag_exchange!, Parallel/Ghost.jl:939-978); the union equals the single-rank result bit for bit."""
    case = cases.amr_case(dim=2, trees=4, maxlevel=2, vtrees=4, vs_maxlevel=0, ragged=False, seed=31)
    full = case.rank_mesh()
    st = case.init_state(full)
    cfg = case.config()
    D, M = 2, 4
    rng = np.random.default_rng(9)
    nl = full.n_local
    prim = np.stack([1.0 + 0.5 * rng.random(nl)] + [0.2 * rng.standard_normal(nl) for _ in range(D)]
                    + [1.0 + 0.3 * rng.random(nl)], axis=1)
    w = _conserved(prim, cfg.gamma, D)
    st.w[: nl * M] = w.ravel(); st.prim[: nl * M] = _prim(w, cfg.gamma, D).ravel()
    st.sw[: nl * M * D] = 0.1 * rng.standard_normal(nl * M * D)
    _, sen_all, _ = orc.ps_criterion(cfg, full, st, 1e300)
    thr = float(np.quantile(sen_all[sen_all > 0], 0.9))   # few flagged cells: many buffers hinge on one neighbour
    loh1, sen1, flg1 = orc.ps_criterion(cfg, full, st, thr)
    assert 0 < flg1.sum() < nl and (sen1 == 2 * thr).any()
    index_of = {int(g): i for i, g in enumerate(full.global_ids[:nl])}
    seen = 0
    for r in range(world):
        mesh = case.rank_mesh(r, world)
        sr = case.init_state(mesh)
        nr = mesh.n_local + mesh.n_ghost
        gl = np.array([index_of[int(g)] for g in mesh.global_ids[:nr]], dtype=np.int64)
        sr.w[: nr * M] = st.w.reshape(-1, M)[gl].ravel()
        sr.prim[: nr * M] = st.prim.reshape(-1, M)[gl].ravel()
        sr.sw[: nr * M * D] = st.sw.reshape(-1, M * D)[gl].ravel()
        ghost_flag = flg1[gl[mesh.n_local:]]
        loh, sen, flg = orc.ps_criterion(case.config(rank=r, nranks=world), mesh, sr, thr, ghost_flag=ghost_flag)
        own = gl[: mesh.n_local]
        assert np.array_equal(flg, flg1[own])
        assert np.array_equal(loh, loh1[own]) and np.array_equal(sen, sen1[own])
        # without the owners' decisions the buffer misses the cells whose only flagged neighbour is a ghost
        if mesh.n_ghost and ghost_flag.any():
            loh_no, _, _ = orc.ps_criterion(case.config(rank=r, nranks=world), mesh, sr, thr)
            seen += int(not np.array_equal(loh_no, loh))
    assert seen > 0


@pytest.mark.parametrize("name", ["amr2d_ragged", "amr3d_ragged", "s2_ib_small"])
def test_criteria_reproduce_golden(name):
    """regression pin of the oracle's adaptation criteria (tests/golden/make_golden_criteria.py): flags bit for bit,
    Löhner values to rounding"""
    import os
    import sys
    gold_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, gold_dir)
    import make_golden_criteria as mgc
    gold = np.load(os.path.join(gold_dir, "criteria.npz"))
    got = mgc.evaluate(name)
    for k, v in got.items():
        ref = gold[f"{name}.{k}"]
        if v.dtype == np.uint8:
            assert np.array_equal(v, ref), k
        else:
            assert np.allclose(v, ref, rtol=1e-12, atol=1e-14), k
    assert got["above"].any() and np.unpackbits(got["refine_lohner"]).any() and np.unpackbits(got["coarsen_lohner"]).any()
