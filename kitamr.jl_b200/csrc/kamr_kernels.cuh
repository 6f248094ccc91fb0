// kamr_kernels.cuh — hand-written sm_100a kernels of the phase-space step.
//
//   K-a  slope_kernel        Flux/Slope.jl:29-116,278-452 (per-level sweep, limiter, transverse projection)
//        flux_kernel         Flux/Flux.jl:94-424 + Flux/CAIDVM.jl:4-141 as an atomic-free cell-centric gather
//   K-b  block_reduce / macro_slope_kernel   Theory/Math.jl:757-762, Slope.jl:1022-1036
//   K-c  update_kernel       Theory/Iterate.jl:96-162 (CAIDVM_Marching, Euler)
//        step_kernel         flux + update fused, convected f kept in shared memory
//   K-e  copy_segments       Parallel/Ghost.jl:757-808 (mirror pack / ghost unpack)
//
// All arithmetic is fp64; no tensor cores (stencil + segmented reduction, HBM/fp64-pipe bound).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include "kamr_types.h"

namespace kamr {

#define KAMR_PI 3.14159265358979323846

// ------------------------------------------------------------------------------------------------
// block-wide sum of NV doubles (warp shuffles + one shared-memory stage); result broadcast to all threads.
// `red` must hold NV*32 doubles.  Deterministic for a fixed block size.
template <int NV>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], off);
    }
    __syncthreads();  // protect `red` from a previous use
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) red[warp * NV + k] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double x = (lane < nwarp) ? red[lane * NV + k] : 0.0;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
            if (lane == 0) red[k] = x;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = red[k];
}

// ------------------------------------------------------------------------------------------------
// kinetics (lib/KitCore)
template <int D>
__device__ __forceinline__ void get_prim(const double* w, double gamma, double* prim) {
    prim[0] = w[0];
    double m2 = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) { prim[1 + d] = w[1 + d] / w[0]; m2 += w[1 + d] * w[1 + d]; }
    prim[D + 1] = 0.5 * w[0] / (gamma - 1.0) / (w[D + 1] - 0.5 * m2 / w[0]);
}

// Maxwellian coefficient rho*(lambda/pi)^(D/2)
template <int D>
__device__ __forceinline__ double maxwell_coef(const double* prim) {
    const double a = prim[D + 1] / KAMR_PI;
    return (D == 2) ? prim[0] * a : prim[0] * (a * sqrt(a));
}
template <int D>
__device__ __forceinline__ double c2_of(const double* v, const double* prim) {
    double c2 = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) { const double c = v[d] - prim[1 + d]; c2 += c * c; }
    return c2;
}
// discrete_maxwell (2D2F.jl:14-21, 3D1F.jl:15-23)
template <int D, int K>
__device__ __forceinline__ void maxwell(const double* v, const double* prim, double coef, double Kin, double* out) {
    const double h = coef * exp(-prim[D + 1] * c2_of<D>(v, prim));
    out[0] = h;
    if (K > 1) out[1] = h * Kin / (2.0 * prim[D + 1]);
}
// shakhov_part (2D2F.jl:22-67, 3D1F.jl:24-40)
template <int D, int K>
__device__ __forceinline__ void shakhov(const double* v, const double* F, const double* prim, const double* qf,
                                        double Pr, double Kin, double* out) {
    double cq = 0.0, c2 = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) { const double c = v[d] - prim[1 + d]; cq += c * qf[d]; c2 += c * c; }
    const double lam = prim[D + 1];
    const double c0 = 0.8 * (1 - Pr) * lam * lam / prim[0] * cq;
    if (D == 2) {
        out[0] = c0 * (2 * lam * c2 + Kin - 5) * F[0];
        if (K > 1) out[1] = c0 * (2 * lam * c2 + Kin - 3) * F[1];
    } else {
        out[0] = c0 * (2 * lam * c2 - 5) * F[0];
    }
}
// add wt * psi(v) * m to the D+2 moment accumulators (micro_to_macro, 2D2F.jl:119, 3D1F.jl:109);
// the energy slot accumulates the un-halved sum, callers multiply by 0.5 once at the end.
template <int D, int K>
__device__ __forceinline__ void add_moments(double* acc, double scale, const double* v, const double* m) {
    const double h = scale * m[0];
    acc[0] += h;
    double v2 = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) { acc[1 + d] += v[d] * h; v2 += v[d] * v[d]; }
    acc[D + 1] += v2 * h + ((K > 1) ? scale * m[1] : 0.0);
}

// ------------------------------------------------------------------------------------------------
// per-cell accessors
template <int D, int K>
struct CellPtr {
    const double* f;      // df planes
    const double* s;      // sdf planes
    const double* v;      // midpoint planes
    const double* wt;
    const int8_t* lev;
    int np;
    __device__ __forceinline__ CellPtr(const DevView& g, const CellInfo& c)
        : f(g.df + c.doff * K), s(g.sdf + c.doff * K * D), v(g.v_mid + c.goff * D), wt(g.v_weight + c.goff),
          lev(g.v_level + c.goff), np(c.np) {}
};

// limiter factor of positivity_preserving_reconstruct (CAIDVM.jl:134-139)
__device__ __forceinline__ double limiter(double f, double s_abs) {
    return fmin(fabs((f - EPS_MACH) / (0.5 * s_abs + EPS_KIT)), 1.);
}

// ------------------------------------------------------------------------------------------------
// Flux gathered by point i of cell `ci` from all of its face slots.
//   fl[k]   += sum_slots area * micro      (update_micro_flux!, Flux.jl:151-344, gather form)
//   mac[m]  += sum_slots area * <psi micro> restricted to what this point contributes
//              (calc_flux fw, CAIDVM.jl:119; update_macro_flux!, Flux.jl:116-136)
template <int D, int K>
__device__ __forceinline__ void point_flux(const DevView& g, const GasPar& gas, const CellInfo& ci, const Slot* slots,
                                           int ns, const double* rho_w, int i, double dt, double* fl, double* mac) {
    const CellPtr<D, K> own(g, ci);
    double v[D], f[K], s[K][D], r[K];
#pragma unroll
    for (int t = 0; t < D; ++t) v[t] = own.v[t * own.np + i];
    const double wt = own.wt[i];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        f[k] = own.f[k * own.np + i];
        double s_abs = 0.0;
#pragma unroll
        for (int t = 0; t < D; ++t) {
            s[k][t] = own.s[(t * K + k) * own.np + i];
            s_abs += ci.ds[t] * fabs(s[k][t]);
        }
        r[k] = limiter(f[k], s_abs);
    }
    for (int q = 0; q < ns; ++q) {
        const Slot& sl = slots[q];
        const int dir = sl.dir;
        const double vn = v[dir];
        const double x = sl.rot * vn;
        const bool own_up = sl.is_here ? (x <= 0.) : (x > 0.);
        const double A = sl.area;
        double m[K];
        if (sl.kind <= SLOT_NBR_SOLID) {
            if (own_up) {
                double dx[D];
#pragma unroll
                for (int t = 0; t < D; ++t) dx[t] = sl.fmid[t] - v[t] * dt - sl.own_mid[t];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    double s_dx = 0.0;
#pragma unroll
                    for (int t = 0; t < D; ++t) s_dx += dx[t] * s[k][t];
                    m[k] = (sl.kind == SLOT_INNER) ? (f[k] + r[k] * s_dx) * vn : (f[k] + s_dx) * vn;
                    fl[k] += A * m[k];
                }
                add_moments<D, K>(mac, A * wt, v, m);
            } else {
                const CellInfo& cn = g.cells[sl.nbr];
                const CellPtr<D, K> nb(g, cn);
                int j0 = i, cnt = 1;
                if (sl.rel >= 0) {
                    const int* st = g.pm_start + g.rel_off[sl.rel];
                    j0 = st[i];
                    cnt = max(1, st[i + 1] - j0);
                }
                const int li = own.lev[i];
                for (int j = j0; j < j0 + cnt; ++j) {
                    double vj[D];
                    if (sl.rel >= 0) {
#pragma unroll
                        for (int t = 0; t < D; ++t) vj[t] = nb.v[t * nb.np + j];
                    } else {
#pragma unroll
                        for (int t = 0; t < D; ++t) vj[t] = v[t];
                    }
                    const double vnj = vj[dir];
                    double dx[D];
#pragma unroll
                    for (int t = 0; t < D; ++t) dx[t] = sl.fmid[t] - vj[t] * dt - sl.nbr_mid[t];
                    double scale = 1.0, wq = wt;
                    if (cnt > 1) {
                        scale = 1.0 / (double)(1 << (D * (nb.lev[j] - li)));
                        wq = nb.wt[j];
                    }
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const double fj = nb.f[k * nb.np + j];
                        if (sl.kind == SLOT_INNER) {
                            double s_abs = 0.0, s_dx = 0.0;
#pragma unroll
                            for (int t = 0; t < D; ++t) {
                                const double sj = nb.s[(t * K + k) * nb.np + j];
                                s_abs += cn.ds[t] * fabs(sj);
                                s_dx += dx[t] * sj;
                            }
                            m[k] = (fj + limiter(fj, s_abs) * s_dx) * vnj;
                        } else {
                            m[k] = fj * vnj;
                        }
                        fl[k] += (A * m[k]) * scale;
                    }
                    add_moments<D, K>(mac, A * wq, vj, m);
                }
            }
        } else {
            // domain faces (calc_domain_flux, CAIDVM.jl:4-97); own_up == heavi (outgoing half)
            double dx[D];
#pragma unroll
            for (int t = 0; t < D; ++t) dx[t] = sl.fmid[t] - v[t] * dt - ci.mid[t];
            if (sl.kind == SLOT_BC_UNIFORM) {
#pragma unroll
                for (int k = 0; k < K; ++k) m[k] = f[k] * vn;
            } else if (own_up) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    double s_dx = 0.0;
#pragma unroll
                    for (int t = 0; t < D; ++t) s_dx += dx[t] * s[k][t];
                    m[k] = (f[k] + s_dx) * vn;
                }
            } else if (sl.kind == SLOT_BC_INTERP) {
                double tmid[D], ndx[D];
#pragma unroll
                for (int t = 0; t < D; ++t) {
                    tmid[t] = 2.0 * sl.fmid[t] - ci.mid[t];
                    ndx[t] = sl.fmid[t] - v[t] * dt - tmid[t];
                }
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const double tdf = f[k] + (tmid[dir] - ci.mid[dir]) * s[k][dir];
                    double s_dx = 0.0;
#pragma unroll
                    for (int t = 0; t < D; ++t) s_dx += ndx[t] * s[k][t];
                    m[k] = (tdf + s_dx) * vn;
                }
            } else {
                double bc[D + 2];
#pragma unroll
                for (int t = 0; t < D + 2; ++t) bc[t] = sl.bc[t];
                if (sl.kind == SLOT_BC_MAXWELL) bc[0] = rho_w[q];
                double F[K];
                maxwell<D, K>(v, bc, maxwell_coef<D>(bc), gas.K, F);
#pragma unroll
                for (int k = 0; k < K; ++k) m[k] = F[k] * vn;
            }
#pragma unroll
            for (int k = 0; k < K; ++k) fl[k] += A * m[k];
            add_moments<D, K>(mac, A * wt, v, m);
        }
    }
}

// Wall density of Maxwellian domain faces: rho_w = -SF/SG (calc_ρw, Theory/Math.jl:251-282).
// One block-wide reduction per Maxwellian slot of the cell (only boundary cells have any).
template <int D, int K>
__device__ __forceinline__ void wall_density(const DevView& g, const CellInfo& ci, const Slot* slots, int ns,
                                             double dt, double* rho_w, double* red) {
    const CellPtr<D, K> own(g, ci);
    for (int q = 0; q < ns; ++q) {
        if (slots[q].kind != SLOT_BC_MAXWELL) continue;  // uniform across the block
        const Slot& sl = slots[q];
        double acc[2] = {0.0, 0.0};
        for (int i = threadIdx.x; i < ci.n; i += blockDim.x) {
            double v[D];
#pragma unroll
            for (int t = 0; t < D; ++t) v[t] = own.v[t * own.np + i];
            const double vn = v[sl.dir], wt = own.wt[i];
            if (sl.rot * vn <= 0.) {
                double s_dx = 0.0;
#pragma unroll
                for (int t = 0; t < D; ++t)
                    s_dx += (sl.fmid[t] - v[t] * dt - ci.mid[t]) * own.s[(t * K + 0) * own.np + i];
                acc[0] += wt * vn * (own.f[i] + s_dx);
            } else {
                acc[1] += wt * vn * exp(-sl.bc[D + 1] * c2_of<D>(v, sl.bc));
            }
        }
        block_reduce<2>(acc, red);
        if (threadIdx.x == 0) {
            const double a = sl.bc[D + 1] / KAMR_PI;
            const double SG = ((D == 2) ? a : a * sqrt(a)) * acc[1];
            rho_w[q] = -acc[0] / SG;
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void load_slots(const DevView& g, const CellInfo& ci, Slot* sh_slots) {
    // cooperative copy of the cell's slot records into shared memory (ints)
    const int ns = ci.slot_end - ci.slot_begin;
    const int nwords = ns * (int)(sizeof(Slot) / sizeof(int));
    const int* src = reinterpret_cast<const int*>(g.slots + ci.slot_begin);
    int* dst = reinterpret_cast<int*>(sh_slots);
    for (int t = threadIdx.x; t < nwords; t += blockDim.x) dst[t] = src[t];
}

// ------------------------------------------------------------------------------------------------
// flux!(p4est, ka): vs_data.flux += sum over faces, ps_data.flux += macro flux (cell-centric gather)
template <int D, int K>
__global__ void __launch_bounds__(256) flux_kernel(DevView g, GasPar gas, const int* __restrict__ cell_list, double dt) {
    __shared__ Slot sh_slots[MAX_SLOTS];
    __shared__ double rho_w[MAX_SLOTS];
    __shared__ double red[(D + 2) * 32];
    __shared__ CellInfo ci;
    const int c = cell_list[blockIdx.x];
    if (threadIdx.x == 0) ci = g.cells[c];
    __syncthreads();
    load_slots(g, ci, sh_slots);
    __syncthreads();
    const int ns = ci.slot_end - ci.slot_begin;
    wall_density<D, K>(g, ci, sh_slots, ns, dt, rho_w, red);
    double mac[D + 2];
#pragma unroll
    for (int q = 0; q < D + 2; ++q) mac[q] = 0.0;
    double* flux = g.flux + ci.doff * K;
    for (int i = threadIdx.x; i < ci.n; i += blockDim.x) {
        double fl[K];
#pragma unroll
        for (int k = 0; k < K; ++k) fl[k] = 0.0;
        point_flux<D, K>(g, gas, ci, sh_slots, ns, rho_w, i, dt, fl, mac);
#pragma unroll
        for (int k = 0; k < K; ++k) flux[k * ci.np + i] += fl[k];
    }
    if (gas.flux_type == 0) {
        block_reduce<D + 2>(mac, red);
        if (threadIdx.x == 0) {
            mac[D + 1] *= 0.5;
#pragma unroll
            for (int q = 0; q < D + 2; ++q) g.mflux[(size_t)c * (D + 2) + q] += mac[q];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// one relaxation update of a cell (Theory/Iterate.jl:96-162), f staged in `fs` (shared or global),
// plane stride `fstride`.  On entry fs holds f^n (Euler) or the convected f (CAIDVM) ...
// The three phases are separated by two block reductions (moments, heat flux).
template <int D, int K>
struct UpdateShared {
    double prim_c[D + 2], prim[D + 2], qf[D], tau, coef_c, coef;
};

// CAIDVM_Marching phases 2+3 and the bookkeeping, given f_conv in fs and the two moment sets reduced.
template <int D, int K>
__device__ __forceinline__ void relax_phases(const DevView& g, const GasPar& gas, const CellInfo& ci, int c,
                                             double* fs, int fstride, double* fout, double dt, const double* w_new,
                                             const double* w0, double* red, UpdateShared<D, K>* us,
                                             int want_residual) {
    const CellPtr<D, K> own(g, ci);
    if (threadIdx.x == 0) {
        get_prim<D>(w_new, gas.gamma, us->prim_c);
        get_prim<D>(w0, gas.gamma, us->prim);
        us->tau = gas.mu_ref * 2.0 * pow(us->prim_c[D + 1], 1 - gas.omega) / us->prim_c[0];  // Gas/Model.jl:14
        us->coef_c = maxwell_coef<D>(us->prim_c);
        us->coef = maxwell_coef<D>(us->prim);
    }
    __syncthreads();
    // phase 2: conservation correction f += M[prim_c] - M[prim]; heat flux of the corrected f about prim_c
    double q[D];
#pragma unroll
    for (int d = 0; d < D; ++d) q[d] = 0.0;
    for (int i = threadIdx.x; i < ci.n; i += blockDim.x) {
        double v[D], Fc[K], F[K], f[K];
#pragma unroll
        for (int t = 0; t < D; ++t) v[t] = own.v[t * own.np + i];
        maxwell<D, K>(v, us->prim_c, us->coef_c, gas.K, Fc);
        maxwell<D, K>(v, us->prim, us->coef, gas.K, F);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            f[k] = fs[k * fstride + i] + (Fc[k] - F[k]);
            fs[k * fstride + i] = f[k];
        }
        const double wt = own.wt[i];
        const double c2 = c2_of<D>(v, us->prim_c);
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const double cd = v[d] - us->prim_c[1 + d];
            q[d] += wt * cd * c2 * f[0] + ((K > 1) ? wt * cd * f[1] : 0.0);
        }
    }
    block_reduce<D>(q, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int d = 0; d < D; ++d) { us->qf[d] = 0.5 * q[d]; g.qf[(size_t)c * D + d] = 0.5 * q[d]; }
    }
    __syncthreads();
    // phase 3: f = f*tau/(tau+dt) + dt/(tau+dt)*(M_c + S[M_c])
    const double tau = us->tau;
    const double a = tau / (tau + dt), b = dt / (tau + dt);
    for (int i = threadIdx.x; i < ci.n; i += blockDim.x) {
        double v[D], Fc[K], Fp[K];
#pragma unroll
        for (int t = 0; t < D; ++t) v[t] = own.v[t * own.np + i];
        maxwell<D, K>(v, us->prim_c, us->coef_c, gas.K, Fc);
        shakhov<D, K>(v, Fc, us->prim_c, us->qf, gas.Pr, gas.K, Fp);
#pragma unroll
        for (int k = 0; k < K; ++k) fout[k * ci.np + i] = fs[k * fstride + i] * a + b * (Fc[k] + Fp[k]);
    }
    if (threadIdx.x == 0) {
        double* prim_old = g.prim + (size_t)c * (D + 2);
        if (want_residual) {  // residual_check!, Solver/Finalize.jl:5-11
#pragma unroll
            for (int m = 0; m < D + 2; ++m) {
                const double dd = us->prim_c[m] - prim_old[m];
                g.res_cell[(size_t)c * 2 * (D + 2) + m] = dd * dd;
                g.res_cell[(size_t)c * 2 * (D + 2) + (D + 2) + m] = fabs(us->prim_c[m]);
            }
        }
#pragma unroll
        for (int m = 0; m < D + 2; ++m) {
            g.w[(size_t)c * (D + 2) + m] = w_new[m];
            prim_old[m] = us->prim_c[m];
            g.mflux[(size_t)c * (D + 2) + m] = 0.0;
        }
    }
}

// iterate!(CAIDVM_Marching | Euler).  STAGE = 1: convected f staged in dynamic shared memory;
// STAGE = 0: staged in the output array itself (cells too large for shared memory).
template <int D, int K, int STAGE>
__global__ void __launch_bounds__(256) update_kernel(DevView g, GasPar gas, const int* __restrict__ cell_list,
                                                     const double* __restrict__ fin_base, double* fout_base, double dt,
                                                     int want_residual) {
    extern __shared__ double dyn[];
    __shared__ double red[2 * (D + 2) * 32];
    __shared__ CellInfo ci;
    __shared__ UpdateShared<D, K> us;
    __shared__ double w_new[D + 2], w0s[D + 2];
    const int c = cell_list[blockIdx.x];
    if (threadIdx.x == 0) ci = g.cells[c];
    __syncthreads();
    const CellPtr<D, K> own(g, ci);
    const double* fin = fin_base + ci.doff * K;
    double* fout = fout_base + ci.doff * K;
    double* vflux = g.flux + ci.doff * K;
    double* fs = STAGE ? dyn : fout;
    const int fstride = STAGE ? ci.n : ci.np;
    const double dtv = dt / ci.vol;
    if (gas.marching == 0) {
        // phase 1: convection f += dt/vol*flux ; moments of the convected f
        double w0[D + 2];
#pragma unroll
        for (int m = 0; m < D + 2; ++m) w0[m] = 0.0;
        for (int i = threadIdx.x; i < ci.n; i += blockDim.x) {
            double v[D], f[K];
#pragma unroll
            for (int t = 0; t < D; ++t) v[t] = own.v[t * own.np + i];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                f[k] = fin[k * ci.np + i] + dtv * vflux[k * ci.np + i];
                fs[k * fstride + i] = f[k];
                vflux[k * ci.np + i] = 0.0;
            }
            add_moments<D, K>(w0, own.wt[i], v, f);
        }
        block_reduce<D + 2>(w0, red);
        if (threadIdx.x == 0) {
            w0[D + 1] *= 0.5;
#pragma unroll
            for (int m = 0; m < D + 2; ++m) {
                w0s[m] = w0[m];
                w_new[m] = g.w[(size_t)c * (D + 2) + m] + g.mflux[(size_t)c * (D + 2) + m] * dt / ci.vol;
            }
        }
        __syncthreads();
        relax_phases<D, K>(g, gas, ci, c, fs, fstride, fout, dt, w_new, w0s, red, &us, want_residual);
    } else {
        // Euler, Iterate.jl:131-162: qf from the pre-convection f, single relaxation with prim(w^{n+1})
        if (threadIdx.x == 0) {
#pragma unroll
            for (int m = 0; m < D + 2; ++m)
                w_new[m] = g.w[(size_t)c * (D + 2) + m] + g.mflux[(size_t)c * (D + 2) + m] * dt / ci.vol;
            get_prim<D>(w_new, gas.gamma, us.prim_c);
            us.tau = gas.mu_ref * 2.0 * pow(us.prim_c[D + 1], 1 - gas.omega) / us.prim_c[0];
            us.coef_c = maxwell_coef<D>(us.prim_c);
        }
        __syncthreads();
        double q[D];
#pragma unroll
        for (int d = 0; d < D; ++d) q[d] = 0.0;
        for (int i = threadIdx.x; i < ci.n; i += blockDim.x) {
            double v[D];
#pragma unroll
            for (int t = 0; t < D; ++t) v[t] = own.v[t * own.np + i];
            const double wt = own.wt[i];
            const double c2 = c2_of<D>(v, us.prim_c);
            const double f0 = fin[i];
            const double f1 = (K > 1) ? fin[ci.np + i] : 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const double cd = v[d] - us.prim_c[1 + d];
                q[d] += wt * cd * c2 * f0 + ((K > 1) ? wt * cd * f1 : 0.0);
            }
        }
        block_reduce<D>(q, red);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int d = 0; d < D; ++d) { us.qf[d] = 0.5 * q[d]; g.qf[(size_t)c * D + d] = 0.5 * q[d]; }
        }
        __syncthreads();
        const double tau = us.tau;
        const double a = tau / (tau + dt), b = dt / (tau + dt);
        for (int i = threadIdx.x; i < ci.n; i += blockDim.x) {
            double v[D], F[K], Fp[K];
#pragma unroll
            for (int t = 0; t < D; ++t) v[t] = own.v[t * own.np + i];
            maxwell<D, K>(v, us.prim_c, us.coef_c, gas.K, F);
            shakhov<D, K>(v, F, us.prim_c, us.qf, gas.Pr, gas.K, Fp);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                fout[k * ci.np + i] = (fin[k * ci.np + i] + dtv * vflux[k * ci.np + i]) * a + b * (F[k] + Fp[k]);
                vflux[k * ci.np + i] = 0.0;
            }
        }
        if (threadIdx.x == 0) {
            double* prim_old = g.prim + (size_t)c * (D + 2);
            if (want_residual) {
#pragma unroll
                for (int m = 0; m < D + 2; ++m) {
                    const double dd = us.prim_c[m] - prim_old[m];
                    g.res_cell[(size_t)c * 2 * (D + 2) + m] = dd * dd;
                    g.res_cell[(size_t)c * 2 * (D + 2) + (D + 2) + m] = fabs(us.prim_c[m]);
                }
            }
#pragma unroll
            for (int m = 0; m < D + 2; ++m) {
                g.w[(size_t)c * (D + 2) + m] = w_new[m];
                prim_old[m] = us.prim_c[m];
                g.mflux[(size_t)c * (D + 2) + m] = 0.0;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// fused flux! + iterate!(CAIDVM_Marching): the face flux of a cell never leaves the SM — the
// convected f is staged in shared memory, the result goes to the second df buffer.
template <int D, int K>
__global__ void __launch_bounds__(256) step_kernel(DevView g, GasPar gas, const int* __restrict__ cell_list, double dt,
                                                   int want_residual) {
    extern __shared__ double dyn[];  // n*K convected f
    __shared__ Slot sh_slots[MAX_SLOTS];
    __shared__ double rho_w[MAX_SLOTS];
    __shared__ double red[2 * (D + 2) * 32];
    __shared__ CellInfo ci;
    __shared__ UpdateShared<D, K> us;
    __shared__ double w_new[D + 2], w0s[D + 2];
    const int c = cell_list[blockIdx.x];
    if (threadIdx.x == 0) ci = g.cells[c];
    __syncthreads();
    load_slots(g, ci, sh_slots);
    __syncthreads();
    const int ns = ci.slot_end - ci.slot_begin;
    wall_density<D, K>(g, ci, sh_slots, ns, dt, rho_w, red);
    const CellPtr<D, K> own(g, ci);
    const double dtv = dt / ci.vol;
    double acc[2 * (D + 2)];  // [0,D+2): macro flux, [D+2, 2D+4): moments of the convected f
#pragma unroll
    for (int q = 0; q < 2 * (D + 2); ++q) acc[q] = 0.0;
    for (int i = threadIdx.x; i < ci.n; i += blockDim.x) {
        double fl[K], f[K], v[D];
#pragma unroll
        for (int k = 0; k < K; ++k) fl[k] = 0.0;
        point_flux<D, K>(g, gas, ci, sh_slots, ns, rho_w, i, dt, fl, acc);
#pragma unroll
        for (int t = 0; t < D; ++t) v[t] = own.v[t * own.np + i];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            f[k] = own.f[k * own.np + i] + dtv * fl[k];
            dyn[k * ci.n + i] = f[k];
        }
        add_moments<D, K>(acc + (D + 2), own.wt[i], v, f);
    }
    block_reduce<2 * (D + 2)>(acc, red);
    if (threadIdx.x == 0) {
        acc[D + 1] *= 0.5;
        acc[2 * D + 3] *= 0.5;
#pragma unroll
        for (int m = 0; m < D + 2; ++m) {
            const double mf = g.mflux[(size_t)c * (D + 2) + m] + ((gas.flux_type == 0) ? acc[m] : 0.0);
            w_new[m] = g.w[(size_t)c * (D + 2) + m] + mf * dt / ci.vol;
            w0s[m] = acc[D + 2 + m];
        }
    }
    __syncthreads();
    double* fout = g.df_new + ci.doff * K;
    relax_phases<D, K>(g, gas, ci, c, dyn, ci.n, fout, dt, w_new, w0s, red, &us, want_residual);
}

// ------------------------------------------------------------------------------------------------
// slopes: one block per (cell of the current level); all DIM directions in one pass.
//   sL = (1/nL) sum_nbr diff(f, P[f_nbr (+ dm . sdf_nbr)]) / dsL   (diff_vs!, Slope.jl:29-64, :278-333)
//   sdf = minmod(sL, sR) | sL | 0                                    (Slope.jl:90-116, 68-86, 471-472)
__device__ __forceinline__ double sgn_(double x) { return (double)((x > 0.) - (x < 0.)); }
__device__ __forceinline__ double minmod(double a, double b) {
    return 0.5 * (sgn_(a) + sgn_(b)) * fmin(fabs(a), fabs(b));
}

template <int D, int K>
__device__ __forceinline__ void side_diff(const DevView& g, const CellInfo& ci, const SlopeSide& sd, int i, int li,
                                          const double* f, double* out) {
#pragma unroll
    for (int k = 0; k < K; ++k) out[k] = 0.0;
    for (int a = 0; a < sd.n; ++a) {
        const CellInfo& cn = g.cells[sd.nbr[a]];
        const CellPtr<D, K> nb(g, cn);
        int j0 = i, cnt = 1;
        if (sd.rel[a] >= 0) {
            const int* st = g.pm_start + g.rel_off[sd.rel[a]];
            j0 = st[i];
            cnt = max(1, st[i + 1] - j0);
        }
        for (int j = j0; j < j0 + cnt; ++j) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double proj = nb.f[k * nb.np + j];
                if (sd.proj[a]) {
#pragma unroll
                    for (int t = 0; t < D; ++t) proj += sd.dm[a][t] * nb.s[(t * K + k) * nb.np + j];
                }
                if (cnt > 1)
                    out[k] += (f[k] - proj) / (double)(1 << (D * (nb.lev[j] - li))) / sd.ds;
                else
                    out[k] += (f[k] - proj) / sd.ds;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) out[k] /= (double)sd.n;
}

template <int D, int K>
__global__ void __launch_bounds__(256) slope_kernel(DevView g, const SlopeTask* __restrict__ tasks) {
    __shared__ SlopeTask tk;
    __shared__ CellInfo ci;
    {
        const int* src = reinterpret_cast<const int*>(tasks + blockIdx.x);
        int* dst = reinterpret_cast<int*>(&tk);
        for (int t = threadIdx.x; t < (int)(sizeof(SlopeTask) / sizeof(int)); t += blockDim.x) dst[t] = src[t];
    }
    __syncthreads();
    if (threadIdx.x == 0) ci = g.cells[tk.cell];
    __syncthreads();
    const CellPtr<D, K> own(g, ci);
    double* sdf = g.sdf + ci.doff * K * D;
    for (int i = threadIdx.x; i < ci.n; i += blockDim.x) {
        double f[K];
#pragma unroll
        for (int k = 0; k < K; ++k) f[k] = own.f[k * own.np + i];
        const int li = own.lev[i];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const SlopeDir& sd = tk.d[d];
            if (sd.mode == SLOPE_KEEP) continue;
            double sA[K], sB[K];
            if (sd.mode == SLOPE_ZERO) {
#pragma unroll
                for (int k = 0; k < K; ++k) sA[k] = 0.0;
            } else {
                side_diff<D, K>(g, ci, sd.A, i, li, f, sA);
                if (sd.mode == SLOPE_INNER) {
                    side_diff<D, K>(g, ci, sd.B, i, li, f, sB);
#pragma unroll
                    for (int k = 0; k < K; ++k) sA[k] = minmod(sA[k], sB[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) sdf[(d * K + k) * ci.np + i] = sA[k];
        }
    }
}

// update_macro_slope!, Slope.jl:1022-1036: sw[:,dir] = <psi sdf[:,:,dir]>
template <int D, int K>
__global__ void __launch_bounds__(256) macro_slope_kernel(DevView g, const int* __restrict__ cell_list) {
    __shared__ double red[D * (D + 2) * 32];
    __shared__ CellInfo ci;
    const int c = cell_list[blockIdx.x];
    if (threadIdx.x == 0) ci = g.cells[c];
    __syncthreads();
    const CellPtr<D, K> own(g, ci);
    double acc[D * (D + 2)];
#pragma unroll
    for (int q = 0; q < D * (D + 2); ++q) acc[q] = 0.0;
    for (int i = threadIdx.x; i < ci.n; i += blockDim.x) {
        double v[D];
#pragma unroll
        for (int t = 0; t < D; ++t) v[t] = own.v[t * own.np + i];
        const double wt = own.wt[i];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            double m[K];
#pragma unroll
            for (int k = 0; k < K; ++k) m[k] = own.s[(d * K + k) * own.np + i];
            add_moments<D, K>(acc + d * (D + 2), wt, v, m);
        }
    }
    block_reduce<D*(D + 2)>(acc, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
            acc[d * (D + 2) + D + 1] *= 0.5;
#pragma unroll
            for (int m = 0; m < D + 2; ++m) g.sw[(size_t)c * (D + 2) * D + d * (D + 2) + m] = acc[d * (D + 2) + m];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// residual sums over cells (residual_check! accumulators), single block, deterministic
__global__ void __launch_bounds__(256) residual_reduce_kernel(const double* __restrict__ res_cell,
                                                              const int* __restrict__ cell_list, int ncell, int nv,
                                                              double* out) {
    __shared__ double red[32];
    for (int q = 0; q < nv; ++q) {
        double a[1] = {0.0};
        for (int t = threadIdx.x; t < ncell; t += blockDim.x) a[0] += res_cell[(size_t)cell_list[t] * nv + q];
        block_reduce<1>(a, red);
        if (threadIdx.x == 0) out[q] = a[0];
    }
}

// halo pack / unpack: copy variable-length segments (one block per segment, grid-stride over segments)
__global__ void __launch_bounds__(256) copy_segments_kernel(const CopySeg* __restrict__ segs, int nseg,
                                                            const double* __restrict__ src, double* __restrict__ dst) {
    for (int s = blockIdx.x; s < nseg; s += gridDim.x) {
        const CopySeg sg = segs[s];
        const double* a = src + sg.src;
        double* b = dst + sg.dst;
        for (long long t = threadIdx.x; t < sg.len; t += blockDim.x) b[t] = a[t];
    }
}

__global__ void fill_kernel(double* p, long long n, double v) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        p[t] = v;
}

}  // namespace kamr
