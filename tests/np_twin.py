"""NumPy twin of one KitAMR time step on a UNIFORM physical mesh with ONE shared velocity grid.

Test infrastructure: a second, independently written restatement of the reference step (vectorised over
the structured mesh instead of the face loop), used only to cross-check oracle/kamr_oracle.c.  Written
from the reference source, not from the C oracle:

  slopes      src/Flux/Slope.jl:20-24 (minmod), :68-116 (bound / inner), :458-488, :653-771
  inner flux  src/Flux/CAIDVM.jl:99-141, src/Flux/Flux.jl:84-136, :349-424 (upwind masks)
  domain flux src/Flux/CAIDVM.jl:4-97, src/Theory/Math.jl:251-282 (calc_ρw)
  update      src/Theory/Iterate.jl:96-162, lib/KitCore/2D2F.jl, 3D1F.jl, 2D.jl, 3D.jl, src/Gas/Model.jl:14

Array convention: f[cell axes (z, y, x)..., k, i] with velocity point i innermost.
"""
import numpy as np

from kitamr_jl_b200 import abi

EPS_KIT = 1e-12
EPS_MACH = 2.0 ** -52


def minmod(a, b):
    return 0.5 * (np.sign(a) + np.sign(b)) * np.minimum(np.abs(a), np.abs(b))


def get_prim(w, gamma):
    D = w.shape[-1] - 2
    p = np.empty_like(w)
    p[..., 0] = w[..., 0]
    for d in range(D):
        p[..., 1 + d] = w[..., 1 + d] / w[..., 0]
    p[..., D + 1] = 0.5 * w[..., 0] / (gamma - 1.0) / (w[..., D + 1] - 0.5 * np.sum(w[..., 1:1 + D] ** 2, axis=-1) / w[..., 0])
    return p


def maxwell(vm, prim, K, ndf):
    """vm [D, n]; prim [..., D+2] -> [..., ndf, n]"""
    D = vm.shape[0]
    c2 = 0.0
    for d in range(D):
        c2 = c2 + (vm[d] - prim[..., 1 + d, None]) ** 2
    lam = prim[..., D + 1, None]
    if D == 2:
        h = prim[..., 0, None] * (lam / np.pi) * np.exp(-lam * c2)
    else:
        h = prim[..., 0, None] * (lam / np.pi) ** 1.5 * np.exp(-lam * c2)
    if ndf == 2:
        return np.stack([h, h * K / (2.0 * lam)], axis=-2)
    return h[..., None, :]


def moments(vm, wt, f):
    """f [..., ndf, n] -> [..., D+2]"""
    D = vm.shape[0]
    h = f[..., 0, :]
    out = [np.sum(wt * h, axis=-1)]
    for d in range(D):
        out.append(np.sum(wt * vm[d] * h, axis=-1))
    e = np.sum(vm ** 2, axis=0) * h
    if f.shape[-2] == 2:
        e = e + f[..., 1, :]
    out.append(0.5 * np.sum(wt * e, axis=-1))
    return np.stack(out, axis=-1)


def shakhov(vm, F, prim, qf, Pr, K):
    D = vm.shape[0]
    c = [vm[d] - prim[..., 1 + d, None] for d in range(D)]
    c2 = sum(x * x for x in c)
    cq = sum(c[d] * qf[..., d, None] for d in range(D))
    lam = prim[..., D + 1, None]
    pre = 0.8 * (1 - Pr) * lam ** 2 / prim[..., 0, None] * cq
    if D == 2:
        out = [pre * (2 * lam * c2 + K - 5) * F[..., 0, :]]
        if F.shape[-2] == 2:
            out.append(pre * (2 * lam * c2 + K - 3) * F[..., 1, :])
    else:
        out = [pre * (2 * lam * c2 - 5) * F[..., 0, :]]
    return np.stack(out, axis=-2)


def i_projection(psi, wt, f, W):
    """solve_I_projection (Theory/I-projection.jl:55-141), vectorised; the Newton systems go through LAPACK
    (numpy.linalg.solve) like the reference's `Symmetric(J,:U) \\ G`.  Returns (lambda, shaved f)."""
    f = f.copy()
    lam = np.zeros(len(W))
    fm = 1.1 * f.min()
    if fm < 0:
        fp = f[f > 0].min()
        d = fp - fm
        neg = f < 0
        f[neg] = (f[neg] - fm) / d * fp
    tol = 1e-10 * max(1.0, np.linalg.norm(W))
    G_prev, stall = np.inf, 0
    for _ in range(10):
        c = wt * f * np.exp(lam @ psi)
        G = (psi * c).sum(axis=1) - W
        J = (psi * c) @ psi.T
        Gn = np.linalg.norm(G)
        if Gn < tol:
            break
        if Gn > 0.9 * G_prev:
            stall += 1
            if stall >= 2:
                break
        else:
            stall = 0
        G_prev = Gn
        dl = -np.linalg.solve(J, G)
        phi0 = (wt * f * np.exp(lam @ psi)).sum() - lam @ W
        slope, a = G @ dl, 1.0
        for _ in range(10):
            lt = lam + a * dl
            if (wt * f * np.exp(lt @ psi)).sum() - lt @ W <= phi0 + 1e-4 * a * slope:
                break
            a *= 0.5
        lam = lam + a * dl
    return lam, f


def heat_flux(vm, wt, f, prim):
    D = vm.shape[0]
    c = [vm[d] - prim[..., 1 + d, None] for d in range(D)]
    c2 = sum(x * x for x in c)
    q = []
    for d in range(D):
        s = np.sum(wt * c[d] * c2 * f[..., 0, :], axis=-1)
        if f.shape[-2] == 2:
            s = s + np.sum(wt * c[d] * f[..., 1, :], axis=-1)
        q.append(0.5 * s)
    return np.stack(q, axis=-1)


class Twin:
    def __init__(self, case, mesh):
        f = case.forest
        assert f.maxlevel == 0 and len(case.grids) == 1, "twin covers uniform meshes with one velocity grid"
        self.case, self.mesh = case, mesh
        self.D, self.K = case.dim, case.ndf
        self.shape = tuple(f.trees_num[::-1])          # (nz, ny, nx): x fastest in the cell order
        g = case.grids[0]
        self.vm = np.ascontiguousarray(g.mid.T)        # [D, n]
        self.wt = g.weight
        self.n = g.n
        self.ds = f.ds[0].copy()
        self.periodic = f.periodic
        D = self.D
        self.mid = f.mid.reshape(self.shape + (D,))
        self.bc_type = mesh.bc_type
        self.bc_prim = mesh.bc_prim.reshape(-1, D + 2)

    def axis(self, d):
        return self.D - 1 - d

    def unflat(self, st):
        D, K, n = self.D, self.K, self.n
        nc = int(np.prod(self.shape))
        f = st.df[: nc * K * n].reshape(self.shape + (K, n)).copy()
        w = st.w[: nc * (D + 2)].reshape(self.shape + (D + 2,)).copy()
        prim = st.prim[: nc * (D + 2)].reshape(self.shape + (D + 2,)).copy()
        return f, w, prim

    # -------------------------------------------------------------- slopes
    def slopes(self, f):
        D = self.D
        s = np.zeros((D,) + f.shape)
        for d in range(D):
            ax = self.axis(d)
            fl = np.roll(f, 1, axis=ax)    # left neighbour (lower coordinate)
            fr = np.roll(f, -1, axis=ax)
            sL = (f - fl) / self.ds[d]
            sR = (f - fr) / (-self.ds[d])
            sd = minmod(sL, sR)
            if not self.periodic[d]:
                lo = [slice(None)] * f.ndim; lo[ax] = 0
                hi = [slice(None)] * f.ndim; hi[ax] = -1
                # one-sided, unlimited: (f - f_n)/(x_c - x_n)  (Slope.jl:653-771)
                sd[tuple(lo)] = sR[tuple(lo)]
                sd[tuple(hi)] = sL[tuple(hi)]
            s[d] = sd
        return s

    # -------------------------------------------------------------- flux
    def _recon(self, f, s, dx, limited):
        """f [..., K, n]; s [D, ..., K, n]; dx [D, ..., n] (broadcast over K)"""
        D = self.D
        s_dx = sum(dx[t][..., None, :] * s[t] for t in range(D))
        if not limited:
            return f + s_dx
        s_abs = sum(self.ds[t] * np.abs(s[t]) for t in range(D))
        r = np.minimum(np.abs((f - EPS_MACH) / (0.5 * s_abs + EPS_KIT)), 1.0)
        return f + r * s_dx

    def flux(self, f, s, dt, gas):
        """returns (vs flux [..., K, n], macro flux [..., D+2]) accumulated over all faces of every cell"""
        D, K = self.D, self.K
        vm, wt = self.vm, self.wt
        flux = np.zeros_like(f)
        mflux = np.zeros(self.shape + (D + 2,))
        for d in range(D):
            ax = self.axis(d)
            A = np.prod([self.ds[t] for t in range(D) if t != d])
            vn = vm[d]
            # ---- interfaces between cell a (lower) and b = a+1 along d; a is `here` through its xmax face
            # (rot = -1): here-upwind v > 0 (v = 0 never occurs), there-upwind v < 0
            fa, sa, mida = f, s, self.mid
            fb, sb, midb = np.roll(f, -1, axis=ax), np.roll(s, -1, axis=ax + 1), np.roll(self.mid, -1, axis=ax)
            fmid = mida.copy(); fmid[..., d] = mida[..., d] - 0.5 * (-1.0) * self.ds[d]
            thmid = midb.copy()
            if self.periodic[d]:   # periodic alias of the neighbour sits across the face
                hi = [slice(None)] * self.mid.ndim; hi[ax] = -1
                alias = fmid[tuple(hi)].copy(); alias[..., d] = alias[..., d] - 0.5 * (-1.0) * self.ds[d]
                thmid[tuple(hi)] = alias
            dxa = [(fmid[..., t, None] - vm[t] * dt) - mida[..., t, None] for t in range(D)]
            dxb = [(fmid[..., t, None] - vm[t] * dt) - thmid[..., t, None] for t in range(D)]
            ma = self._recon(fa, sa, dxa, True) * vn
            mb = self._recon(fb, sb, dxb, True) * vn
            micro = np.where(vn > 0, ma, mb)              # [..., K, n]
            fw = moments(vm, wt, micro)
            area = -A                                      # rot * A with rot = -1
            valid = [slice(None)] * f.ndim
            if not self.periodic[d]:
                valid[ax] = slice(0, -1)
            v = tuple(valid)
            contrib = np.zeros_like(f); contrib[v] = micro[v] * area
            mcontrib = np.zeros_like(mflux); mcontrib[v[:-2]] = fw[v[:-2]] * area
            flux += contrib
            flux -= np.roll(contrib, 1, axis=ax)
            mflux += mcontrib
            mflux -= np.roll(mcontrib, 1, axis=ax)
            # ---- domain faces
            if not self.periodic[d]:
                for side, rot in ((0, 1.0), (1, -1.0)):
                    sl = [slice(None)] * f.ndim; sl[ax] = 0 if side == 0 else -1
                    sl = tuple(sl)
                    b = 2 * d + side
                    fc, sc, midc = f[sl], s[(slice(None),) + sl], self.mid[sl[:-2]]
                    fm = midc.copy(); fm[..., d] = midc[..., d] - 0.5 * rot * self.ds[d]
                    out = (rot * vn) <= 0                  # heavi: outgoing half
                    dx = [(fm[..., t, None] - vm[t] * dt) - midc[..., t, None] for t in range(D)]
                    bt = int(self.bc_type[b]); bc = self.bc_prim[b].copy()
                    if bt == abi.BC_UNIFORM_OUTFLOW:
                        m = fc * vn
                    elif bt == abi.BC_INTERPOLATED_OUTFLOW:
                        tmid = 2.0 * fm - midc
                        ndx = [(fm[..., t, None] - vm[t] * dt) - tmid[..., t, None] for t in range(D)]
                        tdf = fc + (tmid[..., d] - midc[..., d])[..., None, None] * sc[d]
                        m = np.where(out, self._recon(fc, sc, dx, False), self._recon(tdf, sc, ndx, False)) * vn
                    else:
                        rec = self._recon(fc, sc, dx, False)
                        bcs = np.broadcast_to(bc, midc.shape[:-1] + (D + 2,)).copy()
                        if bt == abi.BC_MAXWELLIAN:
                            SF = np.sum(np.where(out, wt * vn * rec[..., 0, :], 0.0), axis=-1)
                            c2 = sum((vm[t] - bc[1 + t]) ** 2 for t in range(D))
                            SG = np.sum(np.where(out, 0.0, wt * vn * np.exp(-bc[D + 1] * c2)))
                            SG = (bc[D + 1] / np.pi) ** (D / 2) * SG
                            bcs[..., 0] = -SF / SG
                        m = np.where(out, rec, maxwell(vm, bcs, gas.K, K)) * vn
                    flux[sl] += rot * A * m
                    mflux[sl[:-2]] += rot * A * moments(vm, wt, m)
        return flux, mflux

    # -------------------------------------------------------------- update
    def update(self, f, w, flux, mflux, dt, gas, marching):
        vm, wt, K = self.vm, self.wt, self.K
        vol = np.prod(self.ds)
        w = w + mflux * dt / vol
        prim_c = get_prim(w, gas.gamma)
        D = self.D
        tau = gas.mu_ref * 2.0 * prim_c[..., D + 1] ** (1 - gas.omega) / prim_c[..., 0]
        tau = tau[..., None, None]
        if marching == abi.MARCH_CAIDVM:
            f = f + dt / vol * flux
            prim = get_prim(moments(vm, wt, f), gas.gamma)
            Fc = maxwell(vm, prim_c, gas.K, K)
            f = f + (Fc - maxwell(vm, prim, gas.K, K))
            qf = heat_flux(vm, wt, f, prim_c)
            Fc = Fc + shakhov(vm, Fc, prim_c, qf, gas.Pr, gas.K)
            f = f * (tau / (tau + dt)) + dt / (tau + dt) * Fc
        elif marching == abi.MARCH_CIP:   # Theory/I-projection.jl:161-192
            f = f + dt / vol * flux
            qf = heat_flux(vm, wt, f, prim_c)
            Fc = maxwell(vm, prim_c, gas.K, K)
            Fc = Fc + shakhov(vm, Fc, prim_c, qf, gas.Pr, gas.K)
            psi = np.vstack([np.ones(self.n), vm, 0.5 * np.sum(vm ** 2, axis=0)])
            f = f.copy()
            for idx in np.ndindex(*f.shape[:-2]):           # conserved_I_porjection! cell by cell
                W = w[idx].copy()
                if K == 2:
                    W[-1] -= 0.5 * np.sum(wt * f[idx][1])
                lam, fh = i_projection(psi, wt, f[idx][0], W)
                f[idx][0] = fh * np.exp(lam @ psi)
            f = f * (tau / (tau + dt)) + dt / (tau + dt) * Fc
        else:  # Euler
            qf = heat_flux(vm, wt, f, prim_c)
            F = maxwell(vm, prim_c, gas.K, K)
            F = F + shakhov(vm, F, prim_c, qf, gas.Pr, gas.K)
            f = (f + dt / vol * flux) * tau / (tau + dt) + dt / (tau + dt) * F
        return f, w, prim_c, qf

    def step(self, st, dt):
        gas = self.case.gas
        f, w, _ = self.unflat(st)
        s = self.slopes(f)
        flux, mflux = self.flux(f, s, dt, gas)
        if self.case.flux_type == abi.FLUX_DVM:   # calc_flux(DVM, ...) returns no macro flux (Flux/DVM.jl:79-99)
            mflux = np.zeros_like(mflux)
        f2, w2, prim, qf = self.update(f, w, flux, mflux, dt, gas, self.case.marching)
        return dict(sdf=s, flux=flux, mflux=mflux, df=f2, w=w2, prim=prim, qf=qf)
