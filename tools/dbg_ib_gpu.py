import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from kitamr_jl_b200 import api, abi
from kitamr_jl_b200.synth import cases
from oracle import orc
case = cases.cylinder_s2(trees=5, ps_maxlevel=4, box_level=2, vtrees=8, vs_maxlevel=2, ib=True)
mesh = case.rank_mesh(); st0 = case.init_state(mesh); cfg = case.config(); dt = case.dt()
ctx = api.Context(cfg); ctx.upload_topology(mesh); ctx.upload_state(st0, aux=True)
ref = st0.copy()
off = mesh.vs_off(); K = 2; M = 4
for it in range(10):
    orc.step(cfg, mesh, ref, dt, False); ctx.step(dt, False)
    out = ctx.download_state(st0.copy(), abi.DL_DF | abi.DL_W | abi.DL_PRIM)
    errs = []
    for c in range(mesh.n_local):
        a = out.df[off[c]*K:off[c+1]*K]; b = ref.df[off[c]*K:off[c+1]*K]
        errs.append(np.linalg.norm(a-b)/(np.linalg.norm(b)+1e-300))
    errs = np.array(errs)
    pe = np.abs(out.prim[:mesh.n_local*M]-ref.prim[:mesh.n_local*M]).reshape(-1, M).max(axis=1)
    top = np.argsort(-errs)[:5]
    print(f"step {it}: df rel L2 max {errs.max():.2e} median {np.median(errs):.2e}; prim abs max {pe.max():.2e}; worst cells",
          [(int(c), int(mesh.bound_enc[c]), f"{errs[c]:.1e}") for c in top])
