"""Properties of the CUDA path at BASELINE.json's full sizes, where the CPU oracle would take too long:
size-independent invariants of the step instead of an element-wise comparison.

  * the fused step (phase kernels that keep the face flux on chip) == slope! -> flux! -> iterate! one call at a time
    (different kernels, different summation trees, same algebra), on the bench workload S2 with its immersed boundary;
  * discrete conservation in a periodic box: sum_c vol_c w_c is unchanged by a step up to the quadrature error that
    the CAIDVM conservation correction removes exactly, with mismatched velocity grids and hanging faces present;
  * a uniform Maxwellian on identical grids is a fixed point of the step;
  * determinism: the same step twice from the same state gives the same bits (no atomics, fixed reduction trees).
"""
import numpy as np
import pytest

from util import local_pts, rel_l2

pytestmark = pytest.mark.gpu


def _run(case, mesh, st, steps, fused=True, residual=False):
    from kitamr_jl_b200 import abi, api
    ctx = api.Context(case.config(device=0))
    try:
        ctx.upload_topology(mesh)
        ctx.upload_state(st, aux=True)
        dt = case.dt()
        for _ in range(steps):
            if fused:
                ctx.step(dt, residual)
            else:
                ctx.slope(); ctx.flux(dt); ctx.iterate(dt, residual)
        return ctx.download_state(st.copy(), abi.DL_DF | abi.DL_W | abi.DL_PRIM)
    finally:
        ctx.close()


def test_s2_full_fused_equals_unfused(kamr_lib):
    from kitamr_jl_b200.synth import cases
    case = cases.cylinder_s2(ib=True)                       # the bench workload: 1.64e7 phase cells
    mesh = case.rank_mesh()
    st = case.init_state(mesh)
    a = _run(case, mesh, st, 2, fused=True)
    b = _run(case, mesh, st, 2, fused=False)
    K, M, nl = mesh.ndf, mesh.dim + 2, mesh.n_local
    assert np.isfinite(a.df).all()
    assert rel_l2(local_pts(mesh, a.df, K), local_pts(mesh, b.df, K)) <= 1e-12
    assert rel_l2(a.w[: nl * M], b.w[: nl * M]) <= 1e-12


def test_s2_full_deterministic(kamr_lib):
    from kitamr_jl_b200.synth import cases
    case = cases.cylinder_s2(ib=True)
    mesh = case.rank_mesh()
    st = case.init_state(mesh)
    a = _run(case, mesh, st, 3)
    b = _run(case, mesh, st, 3)
    assert np.array_equal(a.df, b.df) and np.array_equal(a.w, b.w)


@pytest.mark.parametrize("dim", [2, 3])
def test_periodic_box_conserves(kamr_lib, dim):
    """5 steps in a periodic box (ragged velocity grids, refined ball): total mass, momentum and energy stay put.
    w is advanced by the macro flux only (w += dt/vol * sum_faces fw), and every inner face adds +A fw to one cell
    and -A fw to the other, so the sums change by rounding only."""
    from kitamr_jl_b200.synth import cases
    if dim == 2:
        case = cases.amr_case(dim=2, trees=48, maxlevel=2, vtrees=16, vs_maxlevel=2, ragged=True,
                              periodic=(True, True), seed=41)
    else:
        case = cases.amr_case(dim=3, trees=10, maxlevel=1, vtrees=8, vs_maxlevel=1, ragged=True,
                              periodic=(True, True, True), seed=42)
    mesh = case.rank_mesh()
    st = case.init_state(mesh)
    out = _run(case, mesh, st, 5)
    M, nl = dim + 2, mesh.n_local
    vol = np.prod(mesh.ds.reshape(-1, dim)[:nl], axis=1)
    tot0 = (st.w.reshape(-1, M)[:nl] * vol[:, None]).sum(axis=0)
    tot1 = (out.w.reshape(-1, M)[:nl] * vol[:, None]).sum(axis=0)
    scale = (np.abs(st.w.reshape(-1, M)[:nl]) * vol[:, None]).sum(axis=0)
    assert np.all(np.abs(tot1 - tot0) / scale < 1e-13)


def test_uniform_maxwellian_is_a_fixed_point(kamr_lib):
    """identical velocity grids, uniform state, periodic box, no noise: every slope is exactly 0, every face flux
    cancels, the conservation correction is exactly 0, and relaxation towards M[prim_c] + S (q = 0 up to quadrature)
    leaves w untouched; f stays within the quadrature error of the discrete Maxwellian."""
    from kitamr_jl_b200.synth import cases
    case = cases.amr_case(dim=2, trees=16, maxlevel=1, vtrees=24, vs_maxlevel=0, ragged=False,
                          periodic=(True, True), seed=43)
    case.noise = 0.0
    case.prim_fn = lambda x: np.array([1.0, 0.3, -0.2, 1.1])
    mesh = case.rank_mesh()
    st = case.init_state(mesh)
    out = _run(case, mesh, st, 3)
    M, nl = 4, mesh.n_local
    assert rel_l2(out.w[: nl * M], st.w[: nl * M]) < 1e-13
    assert rel_l2(local_pts(mesh, out.df, 2), local_pts(mesh, st.df, 2)) < 1e-6


@pytest.mark.parametrize("workload,steps", [("S2ib", 2), ("S4", 1)])
def test_full_size_elementwise_vs_oracle(kamr_lib, workload, steps):
    """The bench workloads at BASELINE.json's full size, element by element against the CPU oracle (single process;
    S2ib: 1.64e7 phase cells, S4 sphere3d: 2.2e8): relative L2 <= 1e-12 per step on f, w and prim of the fluid cells."""
    from kitamr_jl_b200.synth import cases
    from oracle import orc
    case = cases.WORKLOADS[workload]()
    mesh = case.rank_mesh()
    st = case.init_state(mesh)
    out = _run(case, mesh, st, steps)
    ref = st.copy()
    cfg = case.config()
    for _ in range(steps):
        orc.step(cfg, mesh, ref, case.dt(), False)
    K, M, nl = mesh.ndf, mesh.dim + 2, mesh.n_local
    tol = 1e-12 * steps
    e_df = rel_l2(local_pts(mesh, out.df, K), local_pts(mesh, ref.df, K))
    e_w = rel_l2(out.w[: nl * M], ref.w[: nl * M])
    fluid = np.repeat(mesh.bound_enc[:nl] >= 0, M)
    e_p = rel_l2(out.prim[: nl * M][fluid], ref.prim[: nl * M][fluid])
    print(f"{case.name}: {mesh.n_phase_local()} phase cells, {steps} step(s): rel L2 df {e_df:.2e} w {e_w:.2e} prim {e_p:.2e}")
    assert e_df <= tol and e_w <= tol and e_p <= tol
