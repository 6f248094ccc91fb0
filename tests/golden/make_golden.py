#!/usr/bin/env python
"""Regenerates tests/golden/*.npz: one oracle step (slope! + flux! + iterate!) of each fixture case, frozen as a
4096-point sample of df plus all of w, prim and qf.  Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import make_golden_cases as mg  # noqa: E402
from oracle import orc  # noqa: E402

for name, fn in mg.CASES.items():
    case = fn()
    mesh = case.rank_mesh()
    st = case.init_state(mesh)
    inp = mg.digest(st.df)
    dt = case.dt()
    res = orc.step(case.config(), mesh, st, dt, True)
    D = case.dim
    nl = mesh.n_local
    idx = mg.sample_idx(len(st.df))
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), input_digest=inp, dt=dt, sample_idx=idx,
                        df_sample=st.df[idx], w=st.w[: nl * (D + 2)], prim=st.prim[: nl * (D + 2)],
                        qf=st.qf[: nl * D], residual=res)
    print(name, "phase cells", mesh.n_phase_local(), "->", f"{name}.npz")
