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
orc.slope(cfg, mesh, ref); ctx.slope()
orc.ib_solid_cells(cfg, mesh, ref); orc.ib_solid_neighbors(cfg, mesh, ref); orc.flux(cfg, mesh, ref, dt)
ctx.flux(dt)
out = ctx.download_state(st0.copy())
off = mesh.vs_off(); K = 2
for c in range(mesh.n_cell):
    a = out.df[off[c]*K:off[c+1]*K]; b = ref.df[off[c]*K:off[c+1]*K]
    e = np.linalg.norm(a-b)/(np.linalg.norm(b)+1e-300)
    if e > 1e-12:
        s = list(mesh.ib.solid_cell).index(c) if c in mesh.ib.solid_cell else -1
        nbs = mesh.ib.solid_nb_ids[mesh.ib.solid_nb_off[s]:mesh.ib.solid_nb_off[s+1]] if s >= 0 else []
        print("cell", c, "benc", mesh.bound_enc[c], "err", e, "grid", mesh.cell_grid[c], "nb grids", [int(mesh.cell_grid[j]) for j in nbs],
              "init-vs-ref", np.linalg.norm(st0.df[off[c]*K:off[c+1]*K]-b)/np.linalg.norm(b), "init-vs-out", np.linalg.norm(st0.df[off[c]*K:off[c+1]*K]-a)/np.linalg.norm(b))
print("w err", np.abs(out.w-ref.w).max())
