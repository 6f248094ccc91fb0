"""The reference's restart payload layout (IO/Restart.jl:192-199) round-trips through the flat host state, and a
reference dump (Kamr.dump_reference_step), when somebody has produced one, pins the oracle against the real KitAMR."""
import glob
import os

import numpy as np
import pytest

from kitamr_jl_b200 import restart
from kitamr_jl_b200.synth import cases


@pytest.mark.parametrize("fn", [
    lambda: cases.amr_case(dim=2, trees=4, maxlevel=2, vtrees=8, vs_maxlevel=2, ragged=True),
    lambda: cases.sphere_s4(trees=4, ps_maxlevel=2, vtrees=4, vs_maxlevel=1),
])
def test_payload_round_trip(fn, tmp_path):
    case = fn()
    mesh = case.rank_mesh()
    st = case.init_state(mesh)
    p = restart.payload_from_state(mesh, st)
    D, K, nl = mesh.dim, mesh.ndf, mesh.n_local
    # Julia's layouts: matrices are column-major over ALL points / quadrants of the rank
    assert p["ws"].shape == (nl, D + 2) and p["ws"].flags["F_CONTIGUOUS"]
    assert p["vs_df"].shape == (int(p["vs_nums"].sum()), K) and p["vs_df"].flags["F_CONTIGUOUS"]
    assert p["vs_midpoints"].shape[1] == D and p["vs_levels"].dtype == np.int8
    # a placeholder quadrant (InsideSolidData: vs_num = 0) in the middle of the rank must be skipped by the reader
    q = {k: v.copy() for k, v in p.items()}
    q["vs_nums"] = np.insert(q["vs_nums"], 3, 0)
    q["bound_encs"] = np.insert(q["bound_encs"], 3, -1)
    q["ws"] = np.asfortranarray(np.insert(q["ws"], 3, np.inf, axis=0))
    restart.save_npz(tmp_path / "restart_0.npz", q)
    back = restart.state_from_payload(restart.load_npz(tmp_path / "restart_0.npz"),
                                      root_weight=float(case.grids[0].weight[case.grids[0].level == 0][0]))
    assert list(back["keep"]) == [i for i in range(nl + 1) if i != 3]
    off = mesh.vs_off()
    assert np.array_equal(back["vs_off"], off[: nl + 1])
    assert np.array_equal(back["df"], st.df[: off[nl] * K])
    assert np.array_equal(back["w"], st.w[: nl * (D + 2)])
    assert np.array_equal(back["bound_enc"], mesh.bound_enc[:nl])
    # grids: same structure per cell (ids may be renumbered), weights rebuilt from the level
    for c in range(nl):
        g0, g1 = int(mesh.cell_grid[c]), int(back["cell_grid"][c])
        a = slice(mesh.grid_off[g0], mesh.grid_off[g0 + 1]); b = slice(back["grid_off"][g1], back["grid_off"][g1 + 1])
        assert np.array_equal(mesh.v_level[a], back["v_level"][b])
        assert np.array_equal(mesh.v_mid[a.start * D: a.stop * D], back["v_mid"][b.start * D: b.stop * D])
        assert np.array_equal(mesh.v_weight[a], back["v_weight"][b])


def test_reference_dump_pins_the_oracle():
    """tests/golden/reference_dumps/<case>/ — produced by Kamr.dump_reference_step under a real Julia + KitAMR.  None
    is committed yet (the build image has no Julia): until one is, parity stays unpinned."""
    from oracle import orc
    from kitamr_jl_b200 import abi
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_dumps")
    dumps = sorted(glob.glob(os.path.join(root, "*", "index.txt")))
    if not dumps:
        pytest.skip("no reference dump committed: the oracle is unpinned by the reference")
    for idx in dumps:
        mesh, st, after, meta = restart.read_reference_dump(os.path.dirname(idx))
        cfg = abi.KamrConfig(meta["dim"], meta["ndf"], int(meta.get("flux_type", 0)), int(meta.get("marching", 0)),
                             meta["K"], meta["Pr"], meta["gamma"], meta["omega"], meta["mu_ref"], 0, 0, 1, None)
        for _ in range(int(meta["steps"])):
            orc.step(cfg, mesh, st, meta["dt"], False)
        n = len(after["df"])
        assert np.linalg.norm(st.df[:n] - after["df"]) <= 1e-12 * np.linalg.norm(after["df"])
        assert np.linalg.norm(st.w[: len(after["w"])] - after["w"]) <= 1e-12 * np.linalg.norm(after["w"])
