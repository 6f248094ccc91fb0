"""Freezes outputs of the oracle's adaptation criteria (orc_ps_criterion, orc_vs_resolution, orc_vs_criterion) for three
cases: a regression pin of the restatement (see make_golden_cases.py for why these are not reference outputs).
    python tests/golden/make_golden_criteria.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import make_golden_cases as mg  # noqa: E402
from kitamr_jl_b200 import abi  # noqa: E402
from oracle import orc  # noqa: E402

NAMES = ["amr2d_ragged", "amr3d_ragged", "s2_ib_small"]
PS_THRESHOLD = 0.05


def evaluate(name):
    case = mg.CASES[name]()
    mesh = case.rank_mesh()
    st = case.init_state(mesh)
    cfg = case.config()
    orc.step(cfg, mesh, st, case.dt())
    orc.slope(cfg, mesh, st)
    loh, sen, flg = orc.ps_criterion(cfg, mesh, st, PS_THRESHOLD)
    out = {"lohner": loh, "sensor": sen, "above": flg.astype(np.uint8)}
    for mode, tag in ((abi.VS_MODE_LOHNER, "lohner"), (abi.VS_MODE_CONTRIBUTION, "contribution")):
        par = abi.vs_adapt(case, mode=mode)
        vr = orc.vs_resolution(cfg, mesh, st, par)
        par.vr_density, par.vr_energy = float(vr[0]), float(vr[1])
        rf, co = orc.vs_criterion(cfg, mesh, st, par)
        out[f"vr_{tag}"] = vr
        out[f"refine_{tag}"] = np.packbits(rf)
        out[f"coarsen_{tag}"] = np.packbits(co)
    return out


if __name__ == "__main__":
    blob = {}
    for name in NAMES:
        for k, v in evaluate(name).items():
            blob[f"{name}.{k}"] = v
    np.savez_compressed(os.path.join(HERE, "criteria.npz"), **blob)
    print("wrote criteria.npz:", {k: v.shape for k, v in blob.items() if k.startswith(NAMES[0])})
