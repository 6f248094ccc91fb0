import sys, numpy as np
sys.path.insert(0, '.')
from kitamr_jl_b200 import abi, api
from kitamr_jl_b200.synth import cases
from oracle import orc
case = cases.x38like_s5(ps_maxlevel=2, trees=4, vtrees=5, vs_maxlevel=1)
mesh = case.rank_mesh(); st0 = case.init_state(mesh); cfg = case.config(device=0)
ctx = api.Context(cfg); ctx.upload_topology(mesh); ctx.upload_state(st0, aux=True)
ref = st0.copy(); dt = case.dt()
orc.slope(cfg, mesh, ref); ctx.slope()
orc.ib_solid_cells(cfg, mesh, ref); orc.ib_solid_neighbors(cfg, mesh, ref); orc.flux(cfg, mesh, ref, dt)
ctx.flux(dt)
out = ctx.download_state(st0.copy())
K = mesh.ndf; D = mesh.dim; off = mesh.vs_off(); nl = mesh.n_local
errs = []
for c in range(nl):
    a = out.flux[off[c]*K:off[c+1]*K]; b = ref.flux[off[c]*K:off[c+1]*K]
    nb = np.linalg.norm(b)
    errs.append(np.linalg.norm(a-b)/(nb if nb > 0 else 1.0))
errs = np.array(errs)
bad = np.nonzero(errs > 1e-10)[0]
print('bad cells', len(bad), 'of', nl)
kind = mesh.face_kind; here = mesh.face_here; there = mesh.face_there
for c in bad[:12]:
    fs = [(int(kind[f]), int(mesh.face_dir[f]), int(here[f]), int(there[f]), int(mesh.cell_grid[there[f]]) if kind[f] != 0 else -1, int(mesh.bound_enc[there[f]]) if kind[f]!=0 else 0) for f in range(len(kind)) if here[f] == c or (kind[f] != 0 and there[f] == c)]
    print(c, 'err %.2e' % errs[c], 'lvl', mesh.ps_level[c], 'be', mesh.bound_enc[c], 'grid', mesh.cell_grid[c], 'n', off[c+1]-off[c], 'ds', mesh.ds[c*D:(c+1)*D], fs)
    a = out.flux[off[c]*K:off[c+1]*K]; b = ref.flux[off[c]*K:off[c+1]*K]
    i = np.argmax(np.abs(a-b)); print('   worst point', i, a[i], b[i])
ctx.close()
