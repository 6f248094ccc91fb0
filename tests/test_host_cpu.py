"""Host-side logic: synthetic forest (p4est stand-in), velocity-grid ordering, and the flattener's face /
neighbour tables that cross the C-ABI (include/kamr.h).  Integer / index work is checked bit-exact."""
import itertools
import os

import numpy as np
import pytest

from kitamr_jl_b200 import abi
from kitamr_jl_b200.synth import cases
from kitamr_jl_b200.synth import vgrid as vg
from kitamr_jl_b200.synth.forest import FACE_BACKHANGING, FACE_DOMAIN, FACE_FULL, FACE_HANGING, Forest, partition


@pytest.mark.parametrize("dim", [2, 3])
def test_forest_balance_and_neighbor_symmetry(dim):
    case = cases.amr_case(dim=dim, trees=4 if dim == 2 else 3, maxlevel=3 if dim == 2 else 2, ragged=False,
                          vtrees=4, vs_maxlevel=0)
    f = case.forest
    vol = np.prod(f.ds, axis=1).sum()
    assert vol == pytest.approx(1.0, rel=1e-14)                    # leaves tile the domain
    for c in range(f.n):
        for face in range(2 * dim):
            state, nbs = f.face_neighbors(c, face)
            opp = face ^ 1
            if state == 0:
                assert len(nbs) == 0
            elif state == 1:
                s2, back = f.face_neighbors(nbs[0], opp)
                assert s2 == 1 and back == [c]
            elif state == -1:
                assert f.level[nbs[0]] == f.level[c] - 1          # 2:1 balance
                s2, back = f.face_neighbors(nbs[0], opp)
                assert s2 == 2 ** (dim - 1) and c in back
            else:
                assert state == 2 ** (dim - 1) and len(nbs) == state
                for j in nbs:
                    assert f.level[j] == f.level[c] + 1
                    s2, back = f.face_neighbors(j, opp)
                    assert s2 == -1 and back == [c]


@pytest.mark.parametrize("dim", [2, 3])
def test_face_list_covers_every_interface_once(dim):
    """initialize_faces! decision tree (Solver/Initialize.jl:10-32): every interior interface appears exactly once on
    a single rank, hanging faces as one record per fine cell with the coarse cell as `here`; face areas add up."""
    case = cases.amr_case(dim=dim, trees=4 if dim == 2 else 3, maxlevel=2 if dim == 2 else 1, ragged=False,
                          vtrees=4, vs_maxlevel=0, periodic=(False, True) + ((False,) if dim == 3 else ()))
    mesh = case.rank_mesh()
    f = case.forest
    D = dim
    seen = set()
    area_in = np.zeros(mesh.n_local)
    ds = mesh.ds.reshape(-1, D)
    for k in range(len(mesh.face_kind)):
        kind, here, there, d = mesh.face_kind[k], mesh.face_here[k], mesh.face_there[k], mesh.face_dir[k]
        a_here = np.prod([ds[here][t] for t in range(D) if t != d])
        if kind == FACE_DOMAIN:
            area_in[here] += a_here
            continue
        key = (min(here, there), max(here, there), int(d), float(mesh.face_mid[k * D + d]))
        assert key not in seen
        seen.add(key)
        assert kind != FACE_BACKHANGING                            # single rank: coarse local cell emits HangingFace
        a = a_here / 2 ** (D - 1) if kind == FACE_HANGING else a_here
        if kind == FACE_HANGING:
            assert f.level[there] == f.level[here] + 1
        area_in[here] += a
        area_in[there] += a
    # every cell's boundary is covered: sum of its face areas == its surface
    for c in range(mesh.n_local):
        surf = 2 * sum(np.prod([ds[c][t] for t in range(D) if t != d]) for d in range(D))
        assert area_in[c] == pytest.approx(surf, rel=1e-13)
    # face midpoints sit on the here cell's boundary, on the side rot says
    mid = mesh.mid.reshape(-1, D)
    for k in range(len(mesh.face_kind)):
        here, d, rot = mesh.face_here[k], mesh.face_dir[k], mesh.face_rot[k]
        assert mesh.face_mid[k * D + d] == mid[here][d] - 0.5 * rot * ds[here][d]


def test_velocity_grid_order_and_weights():
    """Root grid x fastest (Velocity_space/Initialize.jl:22-27); a refined cell is replaced IN PLACE by its children in
    RMT order with weight/2^D (Rebuild.jl:57-71, Abstract/Types.jl:10)."""
    g = vg.root_grid((-4.0, 4.0, -2.0, 2.0), (4, 2))
    assert np.allclose(g.mid[:4, 0], [-3, -1, 1, 3]) and np.all(g.mid[:4, 1] == -1.0)
    assert np.all(g.weight == 4.0)
    flags = np.zeros(g.n, dtype=bool); flags[1] = True
    r = vg.refine(g, flags)
    assert r.n == g.n + 3
    assert list(r.level[:6]) == [0, 1, 1, 1, 1, 0]
    assert np.allclose(r.mid[1:5], [[-1.5, -1.5], [-0.5, -1.5], [-1.5, -0.5], [-0.5, -0.5]])
    assert np.all(r.weight[1:5] == 1.0) and r.weight.sum() == g.weight.sum()
    # v = 0 is a cell corner (check_vs_setting, Solver/Types.jl:335-353): no point has a zero component
    assert np.all(r.mid != 0.0)


@pytest.mark.parametrize("dim", [2, 3])
def test_maxwellian_grid_is_nested_and_conservative(dim):
    ndf = 2 if dim == 2 else 1
    prim = np.array([1.0] + [0.5] * dim + [1.0])
    g = vg.maxwellian_grid(tuple([-5.0, 5.0] * dim), (6,) * dim, 2, prim, ndf, 1.0)
    assert g.weight.sum() == pytest.approx(10.0 ** dim, rel=1e-13)
    assert g.level.max() >= 1                                      # the bulk of the Maxwellian got refined
    far = np.linalg.norm(g.mid - 0.5, axis=1) > 4.5
    assert np.all(g.level[far] == 0)


def test_partition_is_contiguous_and_balanced():
    w = np.array([1, 1, 1, 10, 1, 1, 1, 10, 1, 1], dtype=float)
    owner = partition(w, 2)
    assert np.all(np.diff(owner) >= 0) and set(owner) == {0, 1}
    loads = [w[owner == r].sum() for r in range(2)]
    assert abs(loads[0] - loads[1]) <= 10


def test_mesh_struct_roundtrip_through_ctypes():
    case = cases.amr_case(dim=2, trees=3, maxlevel=1, vtrees=4, vs_maxlevel=1, ragged=True)
    mesh = case.rank_mesh()
    m = mesh.c_struct()
    assert m.n_local == mesh.n_local and m.n_face == len(mesh.face_kind) and m.n_grid == mesh.n_grid
    assert m.grid_off[mesh.n_grid] == len(mesh.v_level)
    assert m.nb_off[mesh.n_local * 4] == len(mesh.nb_ids)
    assert [m.face_here[i] for i in range(m.n_face)] == list(mesh.face_here)
    assert mesh.n_phase_local() == int(mesh.cell_n()[: mesh.n_local].sum())


def test_dt_follows_reference_formula():
    """Status(config): dt = CFL min_d(ds_min_d / U_d), U_d = max |quadrature| - half finest velocity cell
    (Solver/Types.jl:509-520)."""
    case = cases.smoke_s0()
    ds = 1.0 / 16
    U = 5.0 - (10.0 / 16) / 2
    assert case.dt() == pytest.approx(0.4 * ds / U, rel=1e-15)


def test_cost_weighted_partition_balances_the_device_cost_model():
    """bench.py --partition cost: the Morton split weighted by vs_num x the kernel-class cost factor (the weight function a
    GPU-aware shim hands to partition!(p4est, weight)) is contiguous along the curve, keeps every
    rank non-empty and balances the modelled cost better than the reference's vs_num weights."""
    from kitamr_jl_b200.synth import cases
    case = cases.cylinder_s2(copies=2, trees=6, ps_maxlevel=4, box_level=2, vtrees=8, vs_maxlevel=2, ib=True)
    n_of = np.array([g.n for g in case.grids])[case.cell_grid].astype(np.float64)
    n_of = np.where(case.cell_class == -2, 0.0, np.where(case.cell_class == -1, 2.0 * n_of, n_of))
    f = case.cost_factors()
    assert f.shape == (case.forest.n,) and f.min() >= 1.0 and f.max() > 3.0      # donors: 3 + 5.5 per solid face
    spread = {}
    for mode in ("reference", "cost"):
        case.partition_mode = mode
        ow = case.owner(4)
        assert np.all(np.diff(ow) >= 0) and set(ow.tolist()) == {0, 1, 2, 3}
        cost = np.array([(n_of * f)[ow == r].sum() for r in range(4)])
        spread[mode] = cost.max() / cost.mean()
    assert spread["cost"] < spread["reference"] and spread["cost"] < 1.15


def test_reference_arm_prints_the_bench_contract():
    """`bench.py --impl reference` (the CPU restatement on the host cores; no GPU involved) prints ONE JSON line with
    the contract's keys, the same workload / metric / unit as the GPU arm, and counts the fluid phase cells only."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup",
                        "0", "--workload", "S2ib"], capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in j, k
    assert j["impl"] == "reference" and j["unit"] == "cell-updates/s" and j["higher_is_better"] is True
    assert j["config"]["workload"] == "S2-cylinder2d-ib-x1" and j["config"]["phase_cells"] == 16407496
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"]["value"] == j["value"] and j["e2e"]["h2d_bytes_per_step"] == 0
    assert j["value"] > 0
