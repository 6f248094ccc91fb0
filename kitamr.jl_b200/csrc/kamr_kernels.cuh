// kamr_kernels.cuh — hand-written sm_100a kernels of the phase-space step.
//
//   K-a  slope_kernel        Flux/Slope.jl:29-116,278-452 (dependency-wave sweep, minmod, transverse projection)
//        phase_kernel<FLUX>  Flux/Flux.jl:94-424 + Flux/CAIDVM.jl:4-141 as an atomic-free cell-centric gather
//   K-b  block_reduce / macro_slope_kernel   Theory/Math.jl:757-762, Slope.jl:1022-1036
//   K-c  phase_kernel<UPDATE>  Theory/Iterate.jl:96-130 (CAIDVM_Marching); euler_update_kernel :131-162
//        phase_kernel<FUSED>   flux + update fused: the face flux of a cell never leaves the SM
//   K-e  copy_segments       Parallel/Ghost.jl:757-808 (mirror pack / ghost unpack)
//
// All arithmetic is fp64; no tensor cores (stencil + segmented reduction, HBM / fp64-pipe bound).
// One CTA owns one physical cell; the velocity points of the cell are the contiguous inner dimension
// (planes of `np` doubles), so every global access of a warp is a run of 32 consecutive doubles.
#pragma once
#include <cuda_runtime.h>
#include <type_traits>
#include <math_constants.h>
#include "kamr_types.h"

namespace kamr {

#define KAMR_PI 3.14159265358979323846
#ifndef KAMR_UNROLL
#define KAMR_UNROLL 1
#endif
// big cells (n > ~1300 points in 2D2F, ~2000 in 3D1F) take PNT_BIG-thread CTAs, MINB_BIG of them per SM
#ifndef KAMR_PNT_BIG
#define KAMR_PNT_BIG 512
#endif
#ifndef KAMR_MINB_BIG
#define KAMR_MINB_BIG 2
#endif
constexpr int PNT_BIG = KAMR_PNT_BIG;
constexpr int MINB_BIG = KAMR_MINB_BIG;
constexpr int UNROLL = KAMR_UNROLL;
constexpr int KAMR_FLUX_DVM_ = 1;   // == KAMR_FLUX_DVM of include/kamr.h (this header does not include the C-ABI)  // point-loop unrolling of the phase kernels: independent points in flight per thread

// ------------------------------------------------------------------------------------------------
// block-wide sum of NV doubles (warp shuffles + one shared-memory stage); result broadcast to all threads.
// `red` must hold NV*(nwarp+1) doubles.  Deterministic for a fixed block size.
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
    // second stage: thread k adds the nwarp partials of value k in warp order (4 adds for a 128-thread CTA) and
    // leaves the total behind the partials
    if ((int)threadIdx.x < NV) {
        double x = red[threadIdx.x];
        for (int w = 1; w < nwarp; ++w) x += red[w * NV + threadIdx.x];
        red[nwarp * NV + threadIdx.x] = x;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = red[nwarp * NV + k];
}

// ------------------------------------------------------------------------------------------------
// A physical cell is owned by a thread-block CLUSTER of C CTAs (C = 1, 2, 4, 8): CTA r of the cluster takes the
// contiguous point range [r*P, (r+1)*P) of the cell, so a CTA never holds more than a few thousand points whatever
// the size of the velocity grid (the convected f of its range is staged in its shared memory), the number of cells in
// flight — and with it the L2 footprint of their neighbourhoods — stays bounded, and the three per-cell reductions of
// the step (moments, heat flux, wall density) go through distributed shared memory.
template <int C>
struct Cluster {
    static __device__ __forceinline__ unsigned rank() {
        if (C == 1) return 0u;
        unsigned r;
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
        return r;
    }
    static __device__ __forceinline__ void sync() {
        if (C > 1) {
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        }
    }
    // the double at the same shared-memory address in CTA r of the cluster
    static __device__ __forceinline__ double peer(const double* p, unsigned r) {
        const unsigned a = (unsigned)__cvta_generic_to_shared(p);
        unsigned ra;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(r));
        double v;
        asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
        return v;
    }
};

// Sum of NV doubles over all threads of all CTAs of the cluster, broadcast to every thread.  `xch` (NV doubles of
// shared memory, a buffer of its own per call site) publishes this CTA's totals to its peers; the ranks are added in
// rank order by every CTA, so all CTAs hold the same bits.  The caller ends the kernel with Cluster<C>::sync() so that
// no CTA exits while a peer still reads its buffer.
template <int NV, int C>
__device__ __forceinline__ void cluster_reduce(double (&v)[NV], double* red, double* xch) {
    block_reduce<NV>(v, red);
    if (C == 1) return;
    const int nwarp = (blockDim.x + 31) >> 5;
    if ((int)threadIdx.x < NV) xch[threadIdx.x] = red[nwarp * NV + threadIdx.x];
    Cluster<C>::sync();
    if ((int)threadIdx.x < NV) {
        double x = 0.0;
#pragma unroll
        for (int r = 0; r < C; ++r) x += Cluster<C>::peer(xch + threadIdx.x, (unsigned)r);
        red[threadIdx.x] = x;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = red[k];
    __syncthreads();
}

// packed velocity-grid statics (DevView::v_pack / v_tab): tables staged in shared memory, one word per point
template <int D>
__device__ __forceinline__ void load_vtab(const DevView& g, double* tab) {
    const int nw = D * g.n_vtab + VPK_LEVELS;
    for (int t = threadIdx.x; t < nw; t += blockDim.x) tab[t] = g.v_tab[t];
}
template <int D>
__device__ __forceinline__ void unpack_v(unsigned w, const double* __restrict__ tab, int ntab, double* v) {
    constexpr unsigned MSK = (1u << VPK_BITS) - 1u;
    v[0] = tab[w & MSK];
    v[1] = tab[ntab + ((w >> VPK_BITS) & MSK)];
    if (D == 3) v[2] = tab[2 * ntab + ((w >> (2 * VPK_BITS)) & MSK)];
}
template <int D>
__device__ __forceinline__ double unpack_wt(unsigned w, const double* __restrict__ tab, int ntab) {
    return tab[D * ntab + (w >> 27)];
}
// bit d = (v_d > 0): the side whose neighbour is upwind (see hot_flux)
template <int D>
__device__ __forceinline__ unsigned sign_bits(const double* v) {
    unsigned sg = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) sg |= (v[d] > 0.) ? (1u << d) : 0u;
    return sg;
}
__host__ __device__ inline size_t vtab_doubles(int D, int ntab) { return (size_t)D * ntab + VPK_LEVELS; }

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
// exp(x) for x <= 0 (every exponent on this path is -lambda*c^2): Cody-Waite reduction x = n ln2 + r,
// |r| <= ln2/2, degree-13 Taylor polynomial (truncation 4e-18), exponent-field scaling.  ~1 ulp like
// libdevice exp, without its overflow / NaN / large-argument paths.  Results below 2^-1022 go through a
// two-step scaling so they denormalise gradually.
__device__ __forceinline__ double exp_nonpos(double x) {
    const double MAGIC = 6755399441055744.0;  // 2^52 + 2^51: adding it rounds to nearest integer
    double t = fma(x, 1.4426950408889634074, MAGIC);
    const int n = __double2loint(t);
    t -= MAGIC;
    double r = fma(t, -6.93147180369123816490e-01, x);
    r = fma(t, -1.90821492927058770002e-10, r);
    // Estrin evaluation: the same 14 coefficients as a depth-4 tree instead of a 13-deep Horner chain (the chain's
    // fixed latency was a visible stall with ~6 warps per scheduler)
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double a0 = 1.0 + r;
    const double a1 = fma(1.6666666666666666e-01, r, 0.5);                        // 1/2! + r/3!
    const double a2 = fma(8.333333333333333e-03, r, 4.1666666666666664e-02);      // 1/4! + r/5!
    const double a3 = fma(1.984126984126984e-04, r, 1.388888888888889e-03);       // 1/6! + r/7!
    const double a4 = fma(2.7557319223985893e-06, r, 2.48015873015873e-05);       // 1/8! + r/9!
    const double a5 = fma(2.505210838544172e-08, r, 2.755731922398589e-07);       // 1/10! + r/11!
    const double a6 = fma(1.6059043836821613e-10, r, 2.08767569878681e-09);       // 1/12! + r/13!
    const double b0 = fma(a1, r2, a0), b1 = fma(a3, r2, a2), b2 = fma(a5, r2, a4);
    const double c0 = fma(b1, r4, b0), c1 = fma(a6, r4, b2);
    const double p = fma(c1, r8, c0);
    if (n >= -1021) return p * __hiloint2double((n + 1023) << 20, 0);
    if (n < -1100) return 0.0;
    const int h = n / 2;
    return (p * __hiloint2double((h + 1023) << 20, 0)) * __hiloint2double((n - h + 1023) << 20, 0);
}

// discrete_maxwell (2D2F.jl:14-21, 3D1F.jl:15-23)
template <int D, int K>
__device__ __forceinline__ void maxwell(const double* v, const double* prim, double coef, double Kin, double* out) {
    const double h = coef * exp_nonpos(-prim[D + 1] * c2_of<D>(v, prim));
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
// add scale * psi(v) * m to the D+2 moment accumulators (micro_to_macro, 2D2F.jl:119, 3D1F.jl:109);
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

// register-resident component select (a runtime index into a local array would go to local memory)
template <int D>
__device__ __forceinline__ double pick(const double* a, int d) {
    if (D == 2) return d == 0 ? a[0] : a[1];
    return d == 0 ? a[0] : (d == 1 ? a[1] : a[2]);
}

// DRAM -> L2 prefetch of a contiguous block (cp.async.bulk.prefetch.L2, one instruction, no destination): a CTA
// requests the state block of the cell that the CTA launched PF_DIST blocks later will work on, so that cell's first
// touch of its own planes is an L2 hit instead of a DRAM round trip.  p 16-byte aligned, bytes a multiple of 16.
#ifndef KAMR_PF_DIST
#define KAMR_PF_DIST 512
#endif
constexpr int PF_DIST = KAMR_PF_DIST;   // measured best on S1/S2 (gpurun_out/sweep6: 128 .. 2048 and "own cell at start")
__device__ __forceinline__ void l2_prefetch(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Loads of the state arrays (df, limited slopes).  Measured (gpurun_out/sweep4, S2): routing them around L1 with
// ld.global.cg leaves the regular kernel unchanged and slows the general one by 20 % (its second pass re-reads the
// cell's own planes from L1), so the default is a plain load; -DKAMR_USE_CG keeps the experiment reproducible.
#ifndef KAMR_USE_CG
__device__ __forceinline__ double ldg_stream(const double* p) { return *p; }
#else
__device__ __forceinline__ double ldg_stream(const double* p) { return __ldcg(p); }
#endif

// ------------------------------------------------------------------------------------------------
// per-cell accessors
template <int D, int K>
struct CellPtr {
    const double* f;      // df planes
    const double* s;      // sdf planes (raw slopes, reference semantics)
    const double* sl;     // limited slopes r*sdf (device-internal, see slope_kernel)
    const double* v;      // midpoint planes
    const double* wt;
    const int8_t* lev;
    int np;
    __device__ __forceinline__ CellPtr(const DevView& g, const CellInfo& c)
        : f(g.df + c.doff * K), s(g.sdf + c.doff * K * D), sl(g.sdl + c.doff * K * D), v(g.v_mid + c.goff * D),
          wt(g.v_weight + c.goff),
          lev(g.v_level + c.goff), np(c.np) {}
};

// dx = x_face - v*dt - x_cell evaluated in the reference's order without FMA contraction (CAIDVM.jl:108-109):
// face and cell midpoints are O(domain) while dx is O(cell size), so the rounding of the intermediate
// differences is what the result inherits; keeping the order keeps the bits.
__device__ __forceinline__ double face_dx(double fmid, double vdt, double mid) {
    return __dsub_rn(__dsub_rn(fmid, vdt), mid);
}

// limiter factor of positivity_preserving_reconstruct (CAIDVM.jl:134-139):
// min(|(f - eps())/(0.5 s_abs + EPS)|, 1).  The denominator is positive, so the quotient only has to be
// formed when the limiter is active (|f - eps()| < denominator); otherwise the min is exactly 1.
__device__ __forceinline__ double limiter(double f, double s_abs) {
    const double a = fabs(f - EPS_MACH), b = 0.5 * s_abs + EPS_KIT;
    return (a >= b) ? 1.0 : a / b;
}

// One point of a pair-mapped (mismatched velocity grids) neighbour-upwind contribution:
// update_micro_flux!, Flux.jl:151-344 in gather form.
//   fl[k]  += area * micro (mean over covering finer points / injection from the coarser point)
//   mac[m] += area * w_j psi(v_j) micro_j      (the neighbour's share of fw, CAIDVM.jl:119)
template <int D, int K>
__device__ __forceinline__ void pair_flux(const DevView& g, const Slot& sl, double wt, int li, int i, double dt,
                                          bool dvm, double* fl, double* mac) {
    const double* nf = g.df + sl.nbr_doff * K;
    // fluid/fluid: limited slopes; solid side under the DVM flux: the SolidNeighbor's own raw slopes (DVM.jl:91)
    const double* nsl = (sl.kind == SLOT_INNER ? g.sdl : g.sdf) + sl.nbr_doff * K * D;
    const double* nv = g.v_mid + sl.nbr_goff * D;
    const double* nwt = g.v_weight + sl.nbr_goff;
    const int8_t* nlev = g.v_level + sl.nbr_goff;
    const int np = sl.nbr_np;
    const int* st = g.pm_start + sl.rel_off;
    const int j0 = st[i];
    const int cnt = max(1, st[i + 1] - j0);
    const double A = sl.area;
    for (int j = j0; j < j0 + cnt; ++j) {
        double vj[D], dx[D], m[K];
#pragma unroll
        for (int t = 0; t < D; ++t) {
            vj[t] = nv[t * np + j];
            dx[t] = face_dx(sl.fmid[t], __dmul_rn(vj[t], dt), sl.nbr_mid[t]);
        }
        const double vnj = pick<D>(vj, sl.dir);
        double scale = 1.0, wq = wt;
        if (cnt > 1) {
            scale = 1.0 / (double)(1 << (D * (nlev[j] - li)));
            wq = nwt[j];
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double fj = nf[k * np + j];
            if (sl.kind == SLOT_INNER || dvm) {
                double s_dx = 0.0;
#pragma unroll
                for (int t = 0; t < D; ++t) s_dx += dx[t] * nsl[(t * K + k) * np + j];
                m[k] = (fj + s_dx) * vnj;
            } else {
                m[k] = fj * vnj;
            }
            fl[k] += (A * m[k]) * scale;
        }
        add_moments<D, K>(mac, A * wq, vj, m);
    }
}

// domain faces (calc_domain_flux, CAIDVM.jl:4-97); own_up == heavi (outgoing half)
template <int D, int K>
__device__ __forceinline__ void domain_flux(const Slot& sl, const CellInfo& ci, const GasPar& gas, const double* v,
                                            const double* f, const double* s /*[K][D]*/, double rho_w, bool own_up,
                                            double dt, double* fl) {
    const int dir = sl.dir;
    const double vn = pick<D>(v, dir);
    double m[K], dx[D];
#pragma unroll
    for (int t = 0; t < D; ++t) dx[t] = face_dx(sl.fmid[t], __dmul_rn(v[t], dt), ci.mid[t]);
    if (sl.kind == SLOT_BC_UNIFORM) {
#pragma unroll
        for (int k = 0; k < K; ++k) m[k] = f[k] * vn;
    } else if (own_up) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double s_dx = 0.0;
#pragma unroll
            for (int t = 0; t < D; ++t) s_dx += dx[t] * s[k * D + t];
            m[k] = (f[k] + s_dx) * vn;
        }
    } else if (sl.kind == SLOT_BC_INTERP) {
        double tmid[D], ndx[D];
#pragma unroll
        for (int t = 0; t < D; ++t) {
            tmid[t] = __dsub_rn(2.0 * sl.fmid[t], ci.mid[t]);
            ndx[t] = face_dx(sl.fmid[t], __dmul_rn(v[t], dt), tmid[t]);
        }
        const double dmid = pick<D>(tmid, dir) - pick<D>(ci.mid, dir);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double tdf = f[k] + dmid * pick<D>(s + k * D, dir);
            double s_dx = 0.0;
#pragma unroll
            for (int t = 0; t < D; ++t) s_dx += ndx[t] * s[k * D + t];
            m[k] = (tdf + s_dx) * vn;
        }
    } else {
        double bc[D + 2];
#pragma unroll
        for (int t = 0; t < D + 2; ++t) bc[t] = sl.bc[t];
        if (sl.kind == SLOT_BC_MAXWELL) bc[0] = rho_w;
        double F[K];
        maxwell<D, K>(v, bc, maxwell_coef<D>(bc), gas.K, F);
#pragma unroll
        for (int k = 0; k < K; ++k) m[k] = F[k] * vn;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) fl[k] += sl.area * m[k];
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
            const double vn = pick<D>(v, sl.dir), wt = own.wt[i];
            if (sl.rot * vn <= 0.) {
                double s_dx = 0.0;
#pragma unroll
                for (int t = 0; t < D; ++t)
                    s_dx += face_dx(sl.fmid[t], __dmul_rn(v[t], dt), ci.mid[t]) * own.s[(t * K + 0) * own.np + i];
                acc[0] += wt * vn * (own.f[i] + s_dx);
            } else {
                acc[1] += wt * vn * exp_nonpos(-sl.bc[D + 1] * c2_of<D>(v, sl.bc));
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

// cooperative copy of `nwords` 4-byte words into shared memory
__device__ __forceinline__ void copy_words(const void* src_, void* dst_, int nwords) {
    const int* src = reinterpret_cast<const int*>(src_);
    int* dst = reinterpret_cast<int*>(dst_);
    for (int t = threadIdx.x; t < nwords; t += blockDim.x) dst[t] = src[t];
}

// ------------------------------------------------------------------------------------------------
// The hot half of the face gather: fluid/fluid faces (FaceRec, sorted by direction and side at flatten time).
// For a point with v_d > 0 the neighbour across the LOW face of direction d is upwind and the cell itself is upwind
// at the HIGH face; for v_d < 0 the other way round (v_d = 0 never occurs, Solver/Types.jl:335-353, and would
// contribute 0).  So a point visits, per direction, the records of one side with its own data and the records of the
// other side with the neighbour's — no per-face upwind test.  Both halves use the LIMITED slopes r*sdf the slope
// kernel left in g.sdl, so there is no division: micro = (f + dx . (r s)) v_n
// (positivity_preserving_reconstruct, CAIDVM.jl:127-141).
// The neighbour values of the first record of every direction are loaded before any arithmetic so that the
// L2 round trips of the D directions overlap; further records of a side (hanging sub-faces) are rare.
template <int D, int K>
__device__ __forceinline__ void hot_flux(const DevView& g, const FaceRec* hot, const unsigned char* sb, int i,
                                         unsigned sg, double dt, const double* v, const double* f,
                                         const double* s /*[K][D] limited*/, double* fl) {
    double vdt[D];
    double nfv[D][K], nsv[D][K * D];
    int qn[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        vdt[d] = __dmul_rn(v[d], dt);
        const int sn = ((sg >> d) & 1u) ? 0 : 1;             // side whose neighbour is upwind (v_d > 0: the low face)
        const int q = sb[2 * d + sn];
        qn[d] = (q < sb[2 * d + sn + 1] && (hot[q].flags & 2)) ? q : -1;
        if (qn[d] >= 0) {
            const FaceRec& h = hot[q];
            const double* __restrict__ nf = g.df + h.nf_off + i;
            const double* __restrict__ nsl = g.sdl + h.nsl_off + i;
            const int np = h.np;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                nfv[d][k] = ldg_stream(nf + k * np);
#pragma unroll
                for (int t = 0; t < D; ++t) nsv[d][k * D + t] = ldg_stream(nsl + (t * K + k) * np);
            }
        }
    }
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const double vn = v[d];
        const int so = ((sg >> d) & 1u) ? 1 : 0;             // side where the cell itself is upwind
        for (int q = sb[2 * d + so]; q < sb[2 * d + so + 1]; ++q) {
            const FaceRec& h = hot[q];
            const double Avn = h.area * vn;
            double dx[D];
#pragma unroll
            for (int t = 0; t < D; ++t) dx[t] = face_dx(h.fmid[t], vdt[t], h.own_mid[t]);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double val = f[k];
#pragma unroll
                for (int t = 0; t < D; ++t) val += dx[t] * s[k * D + t];
                fl[k] += val * Avn;
            }
        }
        if (qn[d] >= 0) {
            const FaceRec& h = hot[qn[d]];
            const double Avn = h.area * vn;
            double dx[D];
#pragma unroll
            for (int t = 0; t < D; ++t) dx[t] = face_dx(h.fmid[t], vdt[t], h.nbr_mid[t]);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double val = nfv[d][k];
#pragma unroll
                for (int t = 0; t < D; ++t) val += dx[t] * nsv[d][k * D + t];
                fl[k] += val * Avn;
            }
        }
        const int sn = 1 - so;
#ifdef KAMR_HANG_BATCH   // (measured on S4: 11.0 ms against 9.65 ms for the plain loop below: the extra live registers cost more)
        // further sub-faces of a hanging side (2^(DIM-1) - 1 of them): all their neighbour values are requested before
        // any is used, so the side costs one more memory round trip, not one per sub-face
        constexpr int XN = (1 << (D - 1)) - 1;
        const int qb = sb[2 * d + sn] + 1, qe = sb[2 * d + sn + 1];
        if (qb < qe) {
            double xf[XN][K], xs[XN][K * D];
#pragma unroll
            for (int r = 0; r < XN; ++r) {
                const int q = qb + r;
                if (q < qe && (hot[q].flags & 2)) {
                    const FaceRec& h = hot[q];
                    const double* __restrict__ nf = g.df + h.nf_off + i;
                    const double* __restrict__ nsl = g.sdl + h.nsl_off + i;
                    const int np = h.np;
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        xf[r][k] = ldg_stream(nf + k * np);
#pragma unroll
                        for (int t = 0; t < D; ++t) xs[r][k * D + t] = ldg_stream(nsl + (t * K + k) * np);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < XN; ++r) {
                const int q = qb + r;
                if (q < qe && (hot[q].flags & 2)) {
                    const FaceRec& h = hot[q];
                    const double Avn = h.area * vn;
                    double dx[D];
#pragma unroll
                    for (int t = 0; t < D; ++t) dx[t] = face_dx(h.fmid[t], vdt[t], h.nbr_mid[t]);
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        double val = xf[r][k];
#pragma unroll
                        for (int t = 0; t < D; ++t) val += dx[t] * xs[r][k * D + t];
                        fl[k] += val * Avn;
                    }
                }
            }
        }
#else
        for (int q = sb[2 * d + sn] + 1; q < sb[2 * d + sn + 1]; ++q) {   // further sub-faces of a hanging side
            const FaceRec& h = hot[q];
            if (!(h.flags & 2)) continue;
            const double* __restrict__ nf = g.df + h.nf_off + i;
            const double* __restrict__ nsl = g.sdl + h.nsl_off + i;
            const int np = h.np;
            const double Avn = h.area * vn;
            double dx[D];
#pragma unroll
            for (int t = 0; t < D; ++t) dx[t] = face_dx(h.fmid[t], vdt[t], h.nbr_mid[t]);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double val = ldg_stream(nf + k * np);
#pragma unroll
                for (int t = 0; t < D; ++t) val += dx[t] * ldg_stream(nsl + (t * K + k) * np);
                fl[k] += val * Avn;
            }
        }
#endif
    }
}

// One neighbour point of a pair-mapped gather: what is loaded ...
template <int D, int K>
struct MapPt {
    unsigned w;          // packed statics of the neighbour's point
    double f[K], s[K * D];
};
template <int D, int K>
__device__ __forceinline__ void map_load(MapPt<D, K>& p, const double* __restrict__ nf, const double* __restrict__ nsl,
                                         const unsigned* __restrict__ npk, int np, int j) {
    p.w = npk[j];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        p.f[k] = ldg_stream(nf + k * np + j);
#pragma unroll
        for (int t = 0; t < D; ++t) p.s[k * D + t] = ldg_stream(nsl + (t * K + k) * np + j);
    }
}
// ... and what it contributes: fl[k] += area * micro (mean over the covering finer points / injection from the coarser
// point), mac[m] += area * w_j psi(v_j) micro_j (the neighbour's share of fw, CAIDVM.jl:119)
template <int D, int K>
__device__ __forceinline__ void map_apply(const MapPt<D, K>& p, const double* __restrict__ tab, int ntab, bool finer,
                                          int li, const double* fmid, const double* nmid, int d, double A, double wt,
                                          double dt, double* fl, double* mac) {
    double vj[D], dx[D], m[K];
    unpack_v<D>(p.w, tab, ntab, vj);
#pragma unroll
    for (int t = 0; t < D; ++t) dx[t] = face_dx(fmid[t], __dmul_rn(vj[t], dt), nmid[t]);
    double scale = 1.0, wq = wt;
    if (finer) {
        scale = 1.0 / (double)(1 << (D * ((int)(p.w >> 27) - li)));
        wq = unpack_wt<D>(p.w, tab, ntab);
    }
    const double vnj = pick<D>(vj, d);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double s_dx = 0.0;
#pragma unroll
        for (int t = 0; t < D; ++t) s_dx += dx[t] * p.s[k * D + t];
        m[k] = (p.f[k] + s_dx) * vnj;
        fl[k] += (A * m[k]) * scale;
    }
    add_moments<D, K>(mac, A * wq, vj, m);
}
// The pair-mapped neighbour-upwind half of one face for own point i (update_micro_flux!, Flux.jl:151-344 in gather
// form): the own point is covered by / covers points j0 .. j0+cnt-1 of the neighbour's grid (one entry when the
// neighbour is equal or coarser there).  Neighbour points are taken two at a time, loads before arithmetic, so a
// covering set of 2^DIM finer points costs half the memory round trips.
template <int D, int K>
__device__ __forceinline__ void mapped_gather(const DevView& g, const double* __restrict__ tab, int ntab,
                                              const double* __restrict__ nf, const double* __restrict__ nsl,
                                              long long ngoff, int np, int j0, int cnt, int li, const double* fmid,
                                              const double* nmid, int d, double A, double wt, double dt, double* fl,
                                              double* mac) {
    const unsigned* __restrict__ npk = g.v_pack + ngoff;
    const bool finer = cnt > 1;
    for (int j = j0; j < j0 + cnt; j += 2) {
        MapPt<D, K> a, b;
        const bool two = j + 1 < j0 + cnt;
        map_load<D, K>(a, nf, nsl, npk, np, j);
        if (two) map_load<D, K>(b, nf, nsl, npk, np, j + 1);
        map_apply<D, K>(a, tab, ntab, finer, li, fmid, nmid, d, A, wt, dt, fl, mac);
        if (two) map_apply<D, K>(b, tab, ntab, finer, li, fmid, nmid, d, A, wt, dt, fl, mac);
    }
}

// Work distribution of the pair-mapped pass: warps draw 32-point batches of the CTA's range from a shared counter, so
// the points with long gathers (covered by 2^DIM or more finer neighbour points) do not leave the other warps waiting
// at the next barrier.
__device__ __forceinline__ int warp_next_batch(int* ctr) {
    int b = 0;
    if ((threadIdx.x & 31) == 0) b = atomicAdd(ctr, 32);
    return __shfl_sync(0xffffffffu, b, 0);
}

// The macro flux of pair-mapped gathers is summed per 32-point batch with a fixed shuffle tree and kept in a
// [D+2][batches] table in shared memory; the batches are added in index order afterwards.  The result therefore does
// not depend on which warp drew which batch: the step stays bit-reproducible.
template <int NV>
__device__ __forceinline__ void warp_batch_add(double* mbat, int nb, int b, double (&mac)[NV]) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) mac[k] += __shfl_down_sync(0xffffffffu, mac[k], off);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) mbat[k * nb + b] += mac[k];
    }
}
__host__ __device__ inline int batch_count(int P) { return (P + 31) / 32 + 1; }

// The neighbour-upwind half of fluid/fluid faces whose neighbour lives on a DIFFERENT velocity grid (pair-mapped
// records, FaceRec::flags bit1 clear), gathered in a pass of its own (pass A2) after the identical-grid faces.
template <int D, int K>
__device__ __forceinline__ void mapped_flux(const DevView& g, const double* __restrict__ tab, int ntab,
                                            const FaceRec* hot, const unsigned char* sb, int i, unsigned sg, double dt,
                                            double wt, int li, double* fl, double* mac) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const int sn = ((sg >> d) & 1u) ? 0 : 1;
        for (int q = sb[2 * d + sn]; q < sb[2 * d + sn + 1]; ++q) {
            const FaceRec& h = hot[q];
            if (h.flags & 2) continue;   // block-uniform
            const int* __restrict__ st = g.pm_start + h.rel_off;
            const int j0 = st[i];
            const int cnt = max(1, st[i + 1] - j0);
            mapped_gather<D, K>(g, tab, ntab, g.df + h.nf_off, g.sdl + h.nsl_off, h.ngoff, h.np, j0, cnt, li, h.fmid,
                                h.nbr_mid, d, h.area, wt, dt, fl, mac);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// phase_kernel: the per-cell pipeline
//   MODE_FUSED   flux! + iterate!(CAIDVM_Marching): gather face fluxes, convect, moments, relax; reads
//                g.df (+ neighbours), writes g.df_new.  The vs flux never exists in memory.
//   MODE_FLUX    flux!(p4est, ka): vs_data.flux += gathered flux, ps_data.flux += macro flux
//   MODE_UPDATE  iterate!(CAIDVM_Marching) from stored vs_data.flux / ps_data.flux, in place
// STAGE_SMEM: the convected f (K planes) and the h-component of M[prim_c] (1 plane) are staged in dynamic
// shared memory; otherwise (cells larger than shared memory) in the output array, M[prim_c] recomputed.
//
// Pass A gathers the hot slots for every point; pass B (cells on the domain edge or with a neighbour on
// a different velocity grid only) adds the remaining contributions.  Splitting them keeps the hot loop
// free of the rare paths' registers.
enum PhaseMode : int { MODE_FUSED = 0, MODE_FLUX = 1, MODE_UPDATE = 2 };

template <int D, int K>
struct UpdateShared {
    double prim_c[D + 2], prim[D + 2], qf[D], tau, coef_c, coef, cb_c, cb;
};

// M[prim] at one point with the per-cell constants hoisted: h = coef e^{-lambda c^2}, b = h K/(2 lambda)
template <int D, int K>
__device__ __forceinline__ void maxwell_c(const double* v, const double* prim, double coef, double cb, double* out) {
    const double h = coef * exp_nonpos(-prim[D + 1] * c2_of<D>(v, prim));
    out[0] = h;
    if (K > 1) out[1] = h * cb;
}

// The update half shared by the phase kernels: given the block-partial sums acc = [macro flux | moments of the
// convected f] of this CTA's point range [p0, p1) and the convected f staged in fs, finish iterate!(CAIDVM_Marching)
// (Theory/Iterate.jl:108-126).  The reductions run over the whole cluster that owns the cell.  The point loops of
// phases 2 and 3 request the packed statics of TAIL_B2 / TAIL_B3 points before touching any of them, so a thread waits
// for one round trip per batch instead of one per point.
#ifndef KAMR_TAIL_B2
#define KAMR_TAIL_B2 2
#endif
#ifndef KAMR_TAIL_B3
#define KAMR_TAIL_B3 4
#endif
constexpr int TAIL_B2 = KAMR_TAIL_B2, TAIL_B3 = KAMR_TAIL_B3;
template <int D>
struct TailXch { double a[2 * (D + 2)], q[D]; };   // per-CTA exchange buffers of the two cluster reductions

template <int D, int K, int MODE, bool STAGE_SMEM, int NT, int C>
__device__ __forceinline__ void update_tail(const DevView& g, const GasPar& gas, int p0, int p1, int np, double vol, int c,
                                            double dt, int want_residual, double (&acc)[2 * (D + 2)],
                                            const unsigned* __restrict__ pk, const double* __restrict__ tab,
                                            double* fs, int fstride, int soff, double* __restrict__ fout, double* fch,
                                            double* red, TailXch<D>& xch, UpdateShared<D, K>& us, double* w_new,
                                            double* w0s) {
    const int ntab = g.n_vtab;
    cluster_reduce<2 * (D + 2), C>(acc, red, xch.a);
    // moments -> prim_c, prim, tau and the Maxwellian constants: two independent serial chains of fp64 divisions (and a
    // pow); lane 0 of warp 0 takes the conserved state, lane 0 of warp 1 the convected one.  Every CTA of the cluster
    // evaluates them from the same bits.
    if (threadIdx.x == 0) {
        acc[D + 1] *= 0.5;
#pragma unroll
        for (int m = 0; m < D + 2; ++m) {
            double mf = g.mflux[(size_t)c * (D + 2) + m];
            if (MODE == MODE_FUSED && gas.flux_type == 0) mf += acc[m];
            w_new[m] = g.w[(size_t)c * (D + 2) + m] + mf * dt / vol;
        }
        get_prim<D>(w_new, gas.gamma, us.prim_c);
        us.tau = gas.mu_ref * 2.0 * pow(us.prim_c[D + 1], 1 - gas.omega) / us.prim_c[0];  // Gas/Model.jl:14
        us.coef_c = maxwell_coef<D>(us.prim_c);
        us.cb_c = gas.K / (2.0 * us.prim_c[D + 1]);
    } else if (threadIdx.x == 32 % NT) {
        acc[2 * D + 3] *= 0.5;
#pragma unroll
        for (int m = 0; m < D + 2; ++m) w0s[m] = acc[D + 2 + m];
        get_prim<D>(w0s, gas.gamma, us.prim);
        us.coef = maxwell_coef<D>(us.prim);
        us.cb = gas.K / (2.0 * us.prim[D + 1]);
    }
    __syncthreads();

    // ---- phase 2: conservation correction f += M[prim_c] - M[prim]; heat flux of the corrected f about prim_c
    double prim_c[D + 2];
#pragma unroll
    for (int m = 0; m < D + 2; ++m) prim_c[m] = us.prim_c[m];
    const double coef_c = us.coef_c, cb_c = us.cb_c;
    constexpr bool STAGE_FC = STAGE_SMEM;   // h-component of M[prim_c] staged beside f, re-used by phase 3
    {
        double prim[D + 2];
#pragma unroll
        for (int m = 0; m < D + 2; ++m) prim[m] = us.prim[m];
        const double coef = us.coef, cb = us.cb;
        double q[D];
#pragma unroll
        for (int d = 0; d < D; ++d) q[d] = 0.0;
        for (int i0 = p0 + threadIdx.x; i0 < p1; i0 += TAIL_B2 * NT) {
            unsigned wb[TAIL_B2];
#pragma unroll
            for (int u = 0; u < TAIL_B2; ++u) {
                const int i = i0 + u * NT;
                wb[u] = (i < p1) ? pk[i] : 0u;
            }
#pragma unroll
            for (int u = 0; u < TAIL_B2; ++u) {
                const int i = i0 + u * NT;
                if (i < p1) {
                    double v[D], Fc[K], F[K], f[K];
                    unpack_v<D>(wb[u], tab, ntab, v);
                    const double wt = unpack_wt<D>(wb[u], tab, ntab);
                    maxwell_c<D, K>(v, prim_c, coef_c, cb_c, Fc);
                    maxwell_c<D, K>(v, prim, coef, cb, F);
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        f[k] = fs[k * fstride + (i - soff)] + (Fc[k] - F[k]);
                        fs[k * fstride + (i - soff)] = f[k];
                    }
                    if (STAGE_FC) fch[i - soff] = Fc[0];
                    // heat_flux (2D2F.jl:68-88, 3D1F.jl:41-72): q_d = 1/2 sum w c_d (c^2 h + b)
                    const double gq = wt * (c2_of<D>(v, prim_c) * f[0] + ((K > 1) ? f[1] : 0.0));
#pragma unroll
                    for (int d = 0; d < D; ++d) q[d] += (v[d] - prim_c[1 + d]) * gq;
                }
            }
        }
        cluster_reduce<D, C>(q, red, xch.q);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int d = 0; d < D; ++d) us.qf[d] = 0.5 * q[d];
        }
        __syncthreads();
    }

    // ---- phase 3: f = f*tau/(tau+dt) + dt/(tau+dt)*(M_c + S[M_c])
    {
        const double tau = us.tau;
        const double a = tau / (tau + dt), b = dt / (tau + dt);
        double qf[D];
#pragma unroll
        for (int d = 0; d < D; ++d) qf[d] = us.qf[d];
        for (int i0 = p0 + threadIdx.x; i0 < p1; i0 += TAIL_B3 * NT) {
            unsigned wb[TAIL_B3];
#pragma unroll
            for (int u = 0; u < TAIL_B3; ++u) {
                const int i = i0 + u * NT;
                wb[u] = (i < p1) ? pk[i] : 0u;
            }
#pragma unroll
            for (int u = 0; u < TAIL_B3; ++u) {
                const int i = i0 + u * NT;
                if (i < p1) {
                    double v[D], Fc[K], Fp[K];
                    unpack_v<D>(wb[u], tab, ntab, v);
                    if (STAGE_FC) {
                        Fc[0] = fch[i - soff];
                        if (K > 1) Fc[1] = Fc[0] * cb_c;
                    } else {
                        maxwell_c<D, K>(v, prim_c, coef_c, cb_c, Fc);
                    }
                    shakhov<D, K>(v, Fc, prim_c, qf, gas.Pr, gas.K, Fp);
#pragma unroll
                    for (int k = 0; k < K; ++k) fout[k * np + i] = fs[k * fstride + (i - soff)] * a + b * (Fc[k] + Fp[k]);
                }
            }
        }
    }
    if (threadIdx.x == 0 && Cluster<C>::rank() == 0) {
        double* prim_old = g.prim + (size_t)c * (D + 2);
#pragma unroll
        for (int d = 0; d < D; ++d) g.qf[(size_t)c * D + d] = us.qf[d];
        if (want_residual) {  // residual_check!, Solver/Finalize.jl:5-11
#pragma unroll
            for (int m = 0; m < D + 2; ++m) {
                const double dd = prim_c[m] - prim_old[m];
                g.res_cell[(size_t)c * 2 * (D + 2) + m] = dd * dd;
                g.res_cell[(size_t)c * 2 * (D + 2) + (D + 2) + m] = fabs(prim_c[m]);
            }
        }
#pragma unroll
        for (int m = 0; m < D + 2; ++m) {
            g.w[(size_t)c * (D + 2) + m] = w_new[m];
            prim_old[m] = prim_c[m];
            g.mflux[(size_t)c * (D + 2) + m] = 0.0;
        }
    }
    Cluster<C>::sync();   // no CTA leaves while a peer may still read its exchange buffers
}

// point range of CTA `rank` of a C-CTA cluster on a cell of n points: ranges are multiples of 4 points (32 B)
template <int C>
__device__ __forceinline__ void chunk_range(int n, unsigned rank, int& P, int& p0, int& p1) {
    P = (C == 1) ? ((n + 3) & ~3) : ((((n + C - 1) / C) + 3) & ~3);
    p0 = min((int)rank * P, n);
    p1 = min(p0 + P, n);
}
__host__ __device__ inline int chunk_points(int n, int C) { return (C == 1) ? ((n + 3) & ~3) : ((((n + C - 1) / C) + 3) & ~3); }

// dynamic shared memory of the phase kernels:
//   [velocity tables | f: K planes of P | face flux: K planes of P (later the h-plane of M[prim_c]) | moment columns]
// The fused kernels stage BOTH the cell's f and the gathered face flux of every point of the CTA's range, so the gather
// loop carries no moment accumulators (they were the registers it spilled); the moments are taken afterwards in a
// loop over the staged planes.  Kernels with pair-mapped gathers add a [D+2][batches] table for the neighbour-side
// macro flux (its quadrature runs over the NEIGHBOUR's points, CAIDVM.jl:119), see warp_batch_add.
template <int D, int K>
__host__ __device__ inline size_t phase_smem_bytes(int ntab, int P, bool stage, bool split, bool mapped) {
    return sizeof(double) * (((vtab_doubles(D, ntab) + 1) & ~(size_t)1) +
                             (stage ? (size_t)(split ? 2 * K : K + 1) * P : 0) +
                             (mapped && split ? (size_t)(D + 2) * batch_count(P) : 0));
}

template <int D, int K, int MODE, bool STAGE_SMEM, int NT, int MINB, int C>
__global__ void __launch_bounds__(NT, MINB)
    phase_kernel(DevView g, GasPar gas, const int* __restrict__ cell_list, double dt, int want_residual, int pf_dist) {
    extern __shared__ double dyn[];
    constexpr int NSLOT = (MODE == MODE_UPDATE) ? 1 : MaxSlots<D>::value;
    __shared__ Slot sh_slots[NSLOT];
    __shared__ FaceRec hot[NSLOT];
    __shared__ int rare[NSLOT];
    __shared__ double rho_w[NSLOT];
    __shared__ double red[2 * (D + 2) * 33];
    __shared__ CellInfo ci;
    __shared__ UpdateShared<D, K> us;
    __shared__ TailXch<D> xch;
    __shared__ double w_new[D + 2], w0s[D + 2];
    __shared__ int s_next;
    const int cidx = blockIdx.x / C;
    const int c = cell_list[cidx];
    copy_words(g.cells + c, &ci, (int)(sizeof(CellInfo) / sizeof(int)));
    double* tab = dyn;
    load_vtab<D>(g, tab);
    __syncthreads();
    const int n = ci.n, np = ci.np, ntab = g.n_vtab;
    int P, p0, p1;
    chunk_range<C>(n, Cluster<C>::rank(), P, p0, p1);
    const int ns = (MODE == MODE_UPDATE) ? 0 : ci.slot_end - ci.slot_begin;
    const int nrare = (MODE == MODE_UPDATE) ? 0 : ci.rare_count;
    if (MODE != MODE_UPDATE) {
        copy_words(g.hot + ci.hot_begin, hot, ci.side_begin[2 * D] * (int)(sizeof(FaceRec) / sizeof(int)));
        if (nrare > 0) {  // block-uniform: only cells on the domain edge / next to another velocity grid / the body
            copy_words(g.slots + ci.slot_begin, sh_slots, ns * (int)(sizeof(Slot) / sizeof(int)));
            copy_words(g.rare + ci.rare_begin, rare, nrare);
        }
        __syncthreads();
        // (a reduction over ALL points of the cell: every CTA of the cluster evaluates it, boundary cells only)
        if (ci.flags & CELL_HAS_MAXWELL_WALL) wall_density<D, K>(g, ci, sh_slots, ns, dt, rho_w, red);
    }
    const CellPtr<D, K> own(g, ci);
    const unsigned* __restrict__ pk = g.v_pack + ci.goff;
    const double dtv = dt / ci.vol;
    double* vflux = g.flux + ci.doff * K;
    double* fout = (MODE == MODE_FUSED ? g.df_new : g.df) + ci.doff * K;
    double* fsm = dyn + ((vtab_doubles(D, ntab) + 1) & ~(size_t)1);
    double* fs = STAGE_SMEM ? fsm : fout;
    const int fstride = STAGE_SMEM ? P : np;
    const int soff = STAGE_SMEM ? p0 : 0;
    double* fls = fsm + (size_t)K * P;    // staged face flux (SPLIT); afterwards the h-plane of M[prim_c]
    double* fch = fls;
    // SPLIT: the fused kernel with shared-memory staging keeps f and the gathered flux of every point in shared memory
    // and takes the moments in a loop of its own; the other instantiations accumulate them while they gather
    constexpr bool SPLIT = MODE == MODE_FUSED && STAGE_SMEM;
    double* mbat = fsm + (size_t)(2 * K) * P;   // [D+2][NB] macro flux of pair-mapped gathers (SPLIT only)
    const int NB = batch_count(P);

    // ---- phase 1: face fluxes, convection, moments
    double acc[2 * (D + 2)];  // [0,D+2): macro flux, [D+2, 2D+4): moments of the convected f
#pragma unroll
    for (int q = 0; q < 2 * (D + 2); ++q) acc[q] = 0.0;
    const bool has_mapped = (ci.flags & CELL_HAS_MAPPED) != 0;
    const bool any_pair = MODE != MODE_UPDATE && (has_mapped || nrare > 0);
    if (SPLIT && any_pair) {
        for (int t = threadIdx.x; t < (D + 2) * NB; t += NT) mbat[t] = 0.0;
    }
    unsigned wnext = (p0 + (int)threadIdx.x < p1) ? pk[p0 + threadIdx.x] : 0u;
    for (int i = p0 + threadIdx.x; i < p1; i += NT) {  // pass A
        double v[D], f[K], fl[K];
        const unsigned wcur = wnext;
        if (i + NT < p1) wnext = pk[i + NT];   // the next point's word travels while this point computes
        unpack_v<D>(wcur, tab, ntab, v);
        const double wt = unpack_wt<D>(wcur, tab, ntab);
        const unsigned sg = sign_bits<D>(v);
#pragma unroll
        for (int k = 0; k < K; ++k) { f[k] = ldg_stream(own.f + k * np + i); fl[k] = 0.0; }
        if (MODE != MODE_UPDATE) {
            double s[K * D];
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int t = 0; t < D; ++t) s[k * D + t] = ldg_stream(own.sl + (t * K + k) * np + i);
            hot_flux<D, K>(g, hot, ci.side_begin, i, sg, dt, v, f, s, fl);
            if (!SPLIT) add_moments<D, K>(acc, wt, v, fl);
        } else {
#pragma unroll
            for (int k = 0; k < K; ++k) { fl[k] = vflux[k * np + i]; vflux[k * np + i] = 0.0; }
        }
        if (MODE == MODE_FLUX) {
#pragma unroll
            for (int k = 0; k < K; ++k) vflux[k * np + i] += fl[k];
        } else if (SPLIT) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                fs[k * fstride + (i - soff)] = f[k];
                fls[k * fstride + (i - soff)] = fl[k];
            }
        } else {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                f[k] += dtv * fl[k];
                fs[k * fstride + (i - soff)] = f[k];
            }
            add_moments<D, K>(acc + (D + 2), wt, v, f);
        }
    }
    if (MODE != MODE_UPDATE && has_mapped) {   // pass A2 (block-uniform branch): pair-mapped neighbour-upwind halves
        if (threadIdx.x == 0) s_next = p0;
        __syncthreads();
        int ib_static = p0 + (int)(threadIdx.x & ~31u);
        for (;;) {
            // (without the batch table the sums live in per-thread accumulators: a fixed assignment keeps them reproducible)
            int ib;
            if (SPLIT) ib = warp_next_batch(&s_next); else { ib = ib_static; ib_static += NT; }
            if (ib >= p1) break;
            const int i = ib + (int)(threadIdx.x & 31);
            const bool on = i < p1;
            double v[D], flm[K], mac[D + 2];
            const unsigned wcur = on ? pk[i] : 0u;
            unpack_v<D>(wcur, tab, ntab, v);
            const double wt = unpack_wt<D>(wcur, tab, ntab);
#pragma unroll
            for (int k = 0; k < K; ++k) flm[k] = 0.0;
#pragma unroll
            for (int q = 0; q < D + 2; ++q) mac[q] = 0.0;
            // the neighbour's share of fw is weighted with the neighbour's points: kept apart from the own-weighted flux
            if (on) mapped_flux<D, K>(g, tab, ntab, hot, ci.side_begin, i, sign_bits<D>(v), dt, wt, (int)(wcur >> 27), flm, mac);
            if (SPLIT) warp_batch_add<D + 2>(mbat, NB, (ib - p0) >> 5, mac);
            if (!on) continue;
            if (MODE == MODE_FLUX) {
#pragma unroll
                for (int k = 0; k < K; ++k) vflux[k * np + i] += flm[k];
#pragma unroll
                for (int q = 0; q < D + 2; ++q) acc[q] += mac[q];
            } else if (SPLIT) {
#pragma unroll
                for (int k = 0; k < K; ++k) fs[k * fstride + (i - soff)] += dtv * flm[k];
            } else {
#pragma unroll
                for (int q = 0; q < D + 2; ++q) acc[q] += mac[q];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    flm[k] *= dtv;
                    fs[k * fstride + (i - soff)] += flm[k];
                }
                add_moments<D, K>(acc + (D + 2), wt, v, flm);
            }
        }
        __syncthreads();
    }
    if (MODE != MODE_UPDATE && nrare > 0) {  // pass B (block-uniform branch)
        for (int qq = 0; qq < nrare; ++qq) {
            const Slot& sl = sh_slots[rare[qq]];
            const int kind = sl.kind;
            const bool pair_slot = kind <= SLOT_NBR_SOLID && sl.rel_off >= 0;   // block-uniform
            for (int ib = p0 + (int)(threadIdx.x & ~31u); ib < p1; ib += NT) {
                const int i = ib + (int)(threadIdx.x & 31);
                const bool on = i < p1;
                double v[D], fl[K], mac[D + 2];
                const unsigned wcur = on ? pk[i] : 0u;
                unpack_v<D>(wcur, tab, ntab, v);
                const double wt = unpack_wt<D>(wcur, tab, ntab);
#pragma unroll
                for (int k = 0; k < K; ++k) fl[k] = 0.0;
#pragma unroll
                for (int q = 0; q < D + 2; ++q) mac[q] = 0.0;
                const double vn = pick<D>(v, sl.dir);
                const double x = sl.rot * vn;
                const bool own_up = sl.is_here ? (x <= 0.) : (x > 0.);
                bool own_weighted = true;   // fl's macro flux is its own moment (false: pair_flux added the neighbour's)
                bool paired = false;
                if (SPLIT && pair_slot) {   // (all lanes of the warp: the batch sum below is a warp collective)
                    paired = on && !own_up;
                    if (paired) pair_flux<D, K>(g, sl, wt, (int)(wcur >> 27), i, dt, gas.flux_type == KAMR_FLUX_DVM_, fl, mac);
                    warp_batch_add<D + 2>(mbat, NB, (ib - p0) >> 5, mac);
                    if (paired) own_weighted = false;
                }
                if (!on) continue;
                if (paired) {
                    // gathered above
                } else if (kind > SLOT_NBR_SOLID || (kind == SLOT_NBR_SOLID && own_up)) {
                    double f[K], s[K * D];  // raw (unlimited) slopes
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        f[k] = own.f[k * np + i];
#pragma unroll
                        for (int t = 0; t < D; ++t) s[k * D + t] = own.s[(t * K + k) * np + i];
                    }
                    if (kind > SLOT_NBR_SOLID) {
                        domain_flux<D, K>(sl, ci, gas, v, f, s, rho_w[rare[qq]], own_up, dt, fl);
                    } else {  // fluid side of a solid / SolidNeighbor face: unlimited, CAIDVM.jl:110
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            double s_dx = 0.0;
#pragma unroll
                            for (int t = 0; t < D; ++t)
                                s_dx += face_dx(sl.fmid[t], __dmul_rn(v[t], dt), sl.own_mid[t]) * s[k * D + t];
                            fl[k] += sl.area * ((f[k] + s_dx) * vn);
                        }
                    }
                    if (!SPLIT) add_moments<D, K>(acc, wt, v, fl);
                } else if (own_up) {
                    continue;  // own half of a pair-mapped fluid slot was gathered in pass A
                } else if (sl.rel_off < 0) {  // solid side, identical grids: f_wall v_n, CAIDVM.jl:111
                    const double* nf = g.df + sl.nbr_doff * K + i;
                    if (gas.flux_type == KAMR_FLUX_DVM_) {  // DVM.jl:91: (there_df + ndx . there_sdf) v_n, unlimited
                        const double* nsd = g.sdf + sl.nbr_doff * (K * D) + i;
                        double ndx[D];
#pragma unroll
                        for (int t = 0; t < D; ++t) ndx[t] = face_dx(sl.fmid[t], __dmul_rn(v[t], dt), sl.nbr_mid[t]);
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            double s_dx = 0.0;
#pragma unroll
                            for (int t = 0; t < D; ++t) s_dx += ndx[t] * nsd[(t * K + k) * sl.nbr_np];
                            fl[k] += sl.area * ((nf[k * sl.nbr_np] + s_dx) * vn);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < K; ++k) fl[k] += sl.area * (nf[k * sl.nbr_np] * vn);
                    }
                    if (!SPLIT) add_moments<D, K>(acc, wt, v, fl);
                } else {
                    pair_flux<D, K>(g, sl, wt, (int)(wcur >> 27), i, dt, gas.flux_type == KAMR_FLUX_DVM_, fl, acc);
                }
                if (MODE == MODE_FLUX) {
#pragma unroll
                    for (int k = 0; k < K; ++k) vflux[k * np + i] += fl[k];
                } else if (SPLIT) {
                    if (own_weighted) {
#pragma unroll
                        for (int k = 0; k < K; ++k) fls[k * fstride + (i - soff)] += fl[k];
                    } else {
#pragma unroll
                        for (int k = 0; k < K; ++k) fs[k * fstride + (i - soff)] += dtv * fl[k];
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        fl[k] *= dtv;
                        fs[k * fstride + (i - soff)] += fl[k];
                    }
                    add_moments<D, K>(acc + (D + 2), wt, v, fl);
                }
            }
        }
    }
    if (SPLIT) {   // moments of the staged planes: macro flux = <psi fl> (+ the pair-mapped batches), then f* = f + dt/V fl
        if (any_pair) {   // (block-uniform) batches in index order; the energy slots hold un-halved sums like acc's
            __syncthreads();
            if ((int)threadIdx.x < D + 2) {
                double x = 0.0;
                for (int b = 0; b < NB; ++b) x += mbat[threadIdx.x * NB + b];
                mbat[threadIdx.x * NB] = x;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
#pragma unroll
                for (int q = 0; q < D + 2; ++q) acc[q] = mbat[q * NB];
            }
        }
        for (int i = p0 + threadIdx.x; i < p1; i += NT) {
            double v[D], f[K], fl[K];
            const unsigned wcur = pk[i];
            unpack_v<D>(wcur, tab, ntab, v);
            const double wt = unpack_wt<D>(wcur, tab, ntab);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                fl[k] = fls[k * fstride + (i - soff)];
                f[k] = fs[k * fstride + (i - soff)] + dtv * fl[k];
                fs[k * fstride + (i - soff)] = f[k];
            }
            add_moments<D, K>(acc, wt, v, fl);
            add_moments<D, K>(acc + (D + 2), wt, v, f);
        }
    }
    if (MODE == MODE_FLUX) {
        if (gas.flux_type == 0) {
            block_reduce<D + 2>(*reinterpret_cast<double(*)[D + 2]>(acc), red);
            if (threadIdx.x == 0) {
                acc[D + 1] *= 0.5;
#pragma unroll
                for (int q = 0; q < D + 2; ++q) g.mflux[(size_t)c * (D + 2) + q] += acc[q];
            }
        }
        return;
    }
    if (MODE == MODE_FUSED && pf_dist > 0 && threadIdx.x == NT - 1 && (cidx + pf_dist) * C < (int)gridDim.x) {
        // DRAM -> L2 prefetch of this CTA's point range of the cell pf_dist cells ahead in the launch
        const CellInfo* __restrict__ nx = g.cells + cell_list[cidx + pf_dist];
        int P2, q0, q1;
        chunk_range<C>(nx->n, Cluster<C>::rank(), P2, q0, q1);
        if (q1 > q0) {
            const long long nd = nx->doff;
            const long long nnp = nx->np;
            const unsigned nb = (unsigned)((q1 - q0 + 1) & ~1) * 8u;
#pragma unroll
            for (int k = 0; k < K; ++k) l2_prefetch(g.df + nd * K + k * nnp + q0, nb);
#pragma unroll
            for (int k = 0; k < K * D; ++k) l2_prefetch(g.sdl + nd * (K * D) + k * nnp + q0, nb);
        }
    }
    update_tail<D, K, MODE, STAGE_SMEM, NT, C>(g, gas, p0, p1, np, ci.vol, c, dt, want_residual, acc, pk, tab, fs,
                                               fstride, soff, fout, fch, red, xch, us, w_new, w0s);
}

// The face flux of one point of a REGULAR cell from its loaded values (the arithmetic of phase_regular_kernel's gather
// loop as a function, used by the two-points-per-thread path): own-upwind face and neighbour-upwind face per direction,
// transverse dx = (x_t - v_t dt) - x_t formed once.
template <int D, int K>
__device__ __forceinline__ void regular_point_flux(const RegCell& rc, const double* v, unsigned sg, double dt,
                                                   const double* f, const double* s /*[K][D]*/,
                                                   const double (*nfv)[K], const double (*nsv)[K * D], double* fl) {
    double vdt[D], tdx[D];
#pragma unroll
    for (int t = 0; t < D; ++t) {
        vdt[t] = __dmul_rn(v[t], dt);
        tdx[t] = face_dx(rc.mid[t], vdt[t], rc.mid[t]);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) fl[k] = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const double vn = v[d];
        const int sn = ((sg >> d) & 1u) ? 0 : 1;
        {   // own-upwind face: the other side of this direction
            const RegSide& h = rc.side[2 * d + (sn ^ 1)];
            const double Avn = h.area * vn;
            const double dxd = face_dx(h.fmid, vdt[d], rc.mid[d]);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double val = f[k];
#pragma unroll
                for (int t = 0; t < D; ++t) val += (t == d ? dxd : tdx[t]) * s[k * D + t];
                fl[k] += val * Avn;
            }
        }
        {   // neighbour-upwind face
            const RegSide& h = rc.side[2 * d + sn];
            const double Avn = h.area * vn;
            const double dxd = face_dx(h.fmid, vdt[d], h.nmid);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double val = nfv[d][k];
#pragma unroll
                for (int t = 0; t < D; ++t) val += (t == d ? dxd : tdx[t]) * nsv[d][k * D + t];
                fl[k] += val * Avn;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// phase_regular_kernel: the fused flux + update for REGULAR cells — every one of the 2*DIM sides is a single
// fluid/fluid face to a same-size neighbour on the same velocity grid, and the face / neighbour midpoints equal the
// cell's own in the transverse coordinates bit for bit (checked at flatten time; the bulk of every mesh away from level
// jumps, domain edges, velocity-grid changes and the body).  Same arithmetic as phase_kernel<FUSED> pass A with the
// structure fixed at compile time: one RegCell record per cluster, no slot loops, no rare pass.  The side whose
// neighbour is upwind comes from the point's packed word (requested one iteration ahead), so all global loads of a
// point are requested together; the transverse dx = (x_t - v_t dt) - x_t is formed once per point instead of once per
// face.
template <int D, int K, bool STAGE_SMEM, int NT, int MINB, bool MAPPED, int C>
__global__ void __launch_bounds__(NT, MINB)
    phase_regular_kernel(DevView g, GasPar gas, const RegCell* __restrict__ recs, double dt, int want_residual,
                         int pf_dist) {
    extern __shared__ double dyn[];
    __shared__ RegCell rc;
    __shared__ double red[2 * (D + 2) * 33];
    __shared__ UpdateShared<D, K> us;
    __shared__ TailXch<D> xch;
    __shared__ double w_new[D + 2], w0s[D + 2];
    __shared__ int s_next;
    const int cidx = blockIdx.x / C;
    copy_words(recs + cidx, &rc, (int)(sizeof(RegCell) / sizeof(int)));
    double* tab = dyn;
    load_vtab<D>(g, tab);
    __syncthreads();
    const int n = rc.n, np = rc.np, c = rc.cell, ntab = g.n_vtab;
    int P, p0, p1;
    chunk_range<C>(n, Cluster<C>::rank(), P, p0, p1);
    const double dtv = dt / rc.vol;
    double* __restrict__ fout = g.df_new + rc.doff * K;
    double* fsm = dyn + ((vtab_doubles(D, ntab) + 1) & ~(size_t)1);
    double* fs = STAGE_SMEM ? fsm : fout;
    const int fstride = STAGE_SMEM ? P : np;
    const int soff = STAGE_SMEM ? p0 : 0;
    double* fls = fsm + (size_t)K * P;    // staged face flux (SPLIT); afterwards the h-plane of M[prim_c]
    double* fch = fls;
    // see phase_kernel; the regular small-cell kernel (128-thread CTAs) keeps the moments in its gather loop: there the
    // extra pass costs more than the few spilled accumulators (measured on S2: 0.42 -> 0.38 ms)
    constexpr bool SPLIT = STAGE_SMEM && NT > 128;
    double* mbat = fsm + (size_t)(2 * K) * P;   // [D+2][NB] macro flux of the pair-mapped gathers (MAPPED && SPLIT)
    const int NB = batch_count(P);
    const double* __restrict__ gdf = g.df;
    const double* __restrict__ gsl = g.sdl;
    const unsigned* __restrict__ pk = g.v_pack + rc.goff;
    const double* __restrict__ of = gdf + rc.doff * K;
    const double* __restrict__ os = gsl + rc.doff * K * D;
    double acc[2 * (D + 2)];
#pragma unroll
    for (int q = 0; q < 2 * (D + 2); ++q) acc[q] = 0.0;
    if (MAPPED && SPLIT) {
        for (int t = threadIdx.x; t < (D + 2) * NB; t += NT) mbat[t] = 0.0;
    }
#ifdef KAMR_PAIRS   // (measured slower, see DESIGN.md section 6 "tried and not adopted")
    constexpr bool PAIRS = !MAPPED;
#else
    constexpr bool PAIRS = false;
#endif
    if (PAIRS) {
        // Two points per thread (i even): 128-bit loads of the cell's own planes and, where both points take the same
        // upwind neighbour in a direction (sign runs of v_d are long: the exception is a pair that straddles a run
        // boundary), of the neighbour's; 64-bit otherwise.  Planes are 32-byte aligned and padded to a multiple of 4
        // points, so the partner of the last point of an odd cell is padding that exists in memory; it is computed
        // and staged like a point but never enters a moment or the output.
        for (int i = p0 + 2 * (int)threadIdx.x; i < p1; i += 2 * NT) {
            const bool two = i + 1 < p1;
            const uint2 w2 = *reinterpret_cast<const uint2*>(pk + i);
            double v0[D], v1[D];
            unpack_v<D>(w2.x, tab, ntab, v0);
            unpack_v<D>(w2.y, tab, ntab, v1);
            const unsigned sg0 = sign_bits<D>(v0), sg1 = sign_bits<D>(v1);
            double nf0[D][K], nf1[D][K], ns0[D][K * D], ns1[D][K * D];
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const RegSide& ha = rc.side[2 * d + (((sg0 >> d) & 1u) ? 0 : 1)];
                const RegSide& hb = rc.side[2 * d + (((sg1 >> d) & 1u) ? 0 : 1)];
                const double* __restrict__ fa = gdf + ha.ndoff * K + i;
                const double* __restrict__ sa = gsl + ha.ndoff * (K * D) + i;
                if ((((sg0 ^ sg1) >> d) & 1u) == 0u) {
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const double2 x = *reinterpret_cast<const double2*>(fa + k * np);
                        nf0[d][k] = x.x; nf1[d][k] = x.y;
#pragma unroll
                        for (int t = 0; t < D; ++t) {
                            const double2 y = *reinterpret_cast<const double2*>(sa + (t * K + k) * np);
                            ns0[d][k * D + t] = y.x; ns1[d][k * D + t] = y.y;
                        }
                    }
                } else {
                    const double* __restrict__ fb = gdf + hb.ndoff * K + i + 1;
                    const double* __restrict__ sb2 = gsl + hb.ndoff * (K * D) + i + 1;
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        nf0[d][k] = ldg_stream(fa + k * np); nf1[d][k] = ldg_stream(fb + k * np);
#pragma unroll
                        for (int t = 0; t < D; ++t) {
                            ns0[d][k * D + t] = ldg_stream(sa + (t * K + k) * np);
                            ns1[d][k * D + t] = ldg_stream(sb2 + (t * K + k) * np);
                        }
                    }
                }
            }
            double f0[K], f1[K], s0[K * D], s1[K * D], fl0[K], fl1[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const double2 x = *reinterpret_cast<const double2*>(of + k * np + i);
                f0[k] = x.x; f1[k] = x.y;
#pragma unroll
                for (int t = 0; t < D; ++t) {
                    const double2 y = *reinterpret_cast<const double2*>(os + (t * K + k) * np + i);
                    s0[k * D + t] = y.x; s1[k * D + t] = y.y;
                }
            }
            regular_point_flux<D, K>(rc, v0, sg0, dt, f0, s0, nf0, ns0, fl0);
            regular_point_flux<D, K>(rc, v1, sg1, dt, f1, s1, nf1, ns1, fl1);
            const int j = i - soff;
            if (SPLIT) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    *reinterpret_cast<double2*>(fs + k * fstride + j) = make_double2(f0[k], f1[k]);
                    *reinterpret_cast<double2*>(fls + k * fstride + j) = make_double2(fl0[k], fl1[k]);
                }
            } else {
                const double wt0 = unpack_wt<D>(w2.x, tab, ntab), wt1 = unpack_wt<D>(w2.y, tab, ntab);
                add_moments<D, K>(acc, wt0, v0, fl0);
#pragma unroll
                for (int k = 0; k < K; ++k) { f0[k] += dtv * fl0[k]; fs[k * fstride + j] = f0[k]; }
                add_moments<D, K>(acc + (D + 2), wt0, v0, f0);
                if (two) {
                    add_moments<D, K>(acc, wt1, v1, fl1);
#pragma unroll
                    for (int k = 0; k < K; ++k) { f1[k] += dtv * fl1[k]; fs[k * fstride + j + 1] = f1[k]; }
                    add_moments<D, K>(acc + (D + 2), wt1, v1, f1);
                }
            }
        }
    }
    unsigned wnext = (!PAIRS && p0 + (int)threadIdx.x < p1) ? pk[p0 + threadIdx.x] : 0u;
#pragma unroll UNROLL
    for (int i = p0 + threadIdx.x; i < (PAIRS ? p0 : p1); i += NT) {
        double v[D], vdt[D], tdx[D], f[K], s[K * D], fl[K];
        double nfv[D][K], nsv[D][K * D];
        bool mapped[D];
        const unsigned wcur = wnext;
        if (i + NT < p1) wnext = pk[i + NT];
        unpack_v<D>(wcur, tab, ntab, v);
        const double wt = unpack_wt<D>(wcur, tab, ntab);
        const unsigned sg = sign_bits<D>(v);
#pragma unroll
        for (int d = 0; d < D; ++d) {  // neighbour-upwind side: low face for v_d > 0, high face otherwise
            const RegSide& hs = rc.side[2 * d + (((sg >> d) & 1u) ? 0 : 1)];
            mapped[d] = MAPPED && hs.rel_off >= 0;
            if (mapped[d]) continue;   // pair-mapped neighbour: gathered in pass A2
            const long long nd = hs.ndoff;
            const double* __restrict__ nf = gdf + nd * K + i;
            const double* __restrict__ nsl = gsl + nd * (K * D) + i;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                nfv[d][k] = ldg_stream(nf + k * np);
#pragma unroll
                for (int t = 0; t < D; ++t) nsv[d][k * D + t] = ldg_stream(nsl + (t * K + k) * np);
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            f[k] = ldg_stream(of + k * np + i);
            fl[k] = 0.0;
#pragma unroll
            for (int t = 0; t < D; ++t) s[k * D + t] = ldg_stream(os + (t * K + k) * np + i);
        }
#pragma unroll
        for (int t = 0; t < D; ++t) {
            vdt[t] = __dmul_rn(v[t], dt);
            tdx[t] = face_dx(rc.mid[t], vdt[t], rc.mid[t]);
        }
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const double vn = v[d];
            const int sn = ((sg >> d) & 1u) ? 0 : 1;
            {   // own-upwind face: the other side of this direction
                const RegSide& h = rc.side[2 * d + (sn ^ 1)];
                const double Avn = h.area * vn;
                const double dxd = face_dx(h.fmid, vdt[d], rc.mid[d]);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    double val = f[k];
#pragma unroll
                    for (int t = 0; t < D; ++t) val += (t == d ? dxd : tdx[t]) * s[k * D + t];
                    fl[k] += val * Avn;
                }
            }
            if (!mapped[d]) {   // neighbour-upwind face
                const RegSide& h = rc.side[2 * d + sn];
                const double Avn = h.area * vn;
                const double dxd = face_dx(h.fmid, vdt[d], h.nmid);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    double val = nfv[d][k];
#pragma unroll
                    for (int t = 0; t < D; ++t) val += (t == d ? dxd : tdx[t]) * nsv[d][k * D + t];
                    fl[k] += val * Avn;
                }
            }
        }
        if (SPLIT) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                fs[k * fstride + (i - soff)] = f[k];
                fls[k * fstride + (i - soff)] = fl[k];
            }
        } else {
            add_moments<D, K>(acc, wt, v, fl);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                f[k] += dtv * fl[k];
                fs[k * fstride + (i - soff)] = f[k];
            }
            add_moments<D, K>(acc + (D + 2), wt, v, f);
        }
    }
    if (MAPPED) {
        // pass A2: neighbour-upwind halves across faces to another velocity grid (update_micro_flux!, Flux.jl:151-344 in
        // gather form): point i is covered by / covers points st[i] .. st[i+1]-1 over there.  Its macro flux is weighted
        // with the NEIGHBOUR's points, so it is kept apart from the own-weighted flux (per-thread columns / acc).
        if (threadIdx.x == 0) s_next = p0;
        __syncthreads();
        int ib_static = p0 + (int)(threadIdx.x & ~31u);
        for (;;) {
            int ib;
            if (SPLIT) ib = warp_next_batch(&s_next); else { ib = ib_static; ib_static += NT; }
            if (ib >= p1) break;
            const int i = ib + (int)(threadIdx.x & 31);
            const bool on = i < p1;
            double v[D], flm[K], mac[D + 2];
            const unsigned wcur = on ? pk[i] : 0u;
            unpack_v<D>(wcur, tab, ntab, v);
            const double wt = unpack_wt<D>(wcur, tab, ntab);
            const unsigned sg = sign_bits<D>(v);
#pragma unroll
            for (int k = 0; k < K; ++k) flm[k] = 0.0;
#pragma unroll
            for (int q = 0; q < D + 2; ++q) mac[q] = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const RegSide& h = rc.side[2 * d + (((sg >> d) & 1u) ? 0 : 1)];
                if (h.rel_off < 0 || !on) continue;
                const int* __restrict__ st = g.pm_start + h.rel_off;
                const int j0 = st[i];
                const int cnt = max(1, st[i + 1] - j0);
                double fm[D], nm[D];
#pragma unroll
                for (int t = 0; t < D; ++t) { fm[t] = (t == d) ? h.fmid : rc.mid[t]; nm[t] = (t == d) ? h.nmid : rc.mid[t]; }
                mapped_gather<D, K>(g, tab, ntab, gdf + h.ndoff * K, gsl + h.ndoff * (K * D), h.ngoff, h.np, j0, cnt,
                                    (int)(wcur >> 27), fm, nm, d, h.area, wt, dt, flm, mac);
            }
            if (SPLIT) warp_batch_add<D + 2>(mbat, NB, (ib - p0) >> 5, mac);
            if (!on) continue;
            if (SPLIT) {
#pragma unroll
                for (int k = 0; k < K; ++k) fs[k * fstride + (i - soff)] += dtv * flm[k];
            } else {
#pragma unroll
                for (int q = 0; q < D + 2; ++q) acc[q] += mac[q];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    flm[k] *= dtv;
                    fs[k * fstride + (i - soff)] += flm[k];
                }
                add_moments<D, K>(acc + (D + 2), wt, v, flm);
            }
        }
        __syncthreads();
    }
    if (SPLIT) {   // moments of the staged planes: macro flux = <psi fl> (+ the pair-mapped columns), then f* = f + dt/V fl
        if (MAPPED) {   // batches in index order (the barrier at the end of pass A2 ordered the table)
            if ((int)threadIdx.x < D + 2) {
                double x = 0.0;
                for (int b = 0; b < NB; ++b) x += mbat[threadIdx.x * NB + b];
                mbat[threadIdx.x * NB] = x;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
#pragma unroll
                for (int q = 0; q < D + 2; ++q) acc[q] = mbat[q * NB];
            }
        }
        for (int i = p0 + threadIdx.x; i < p1; i += NT) {
            double v[D], f[K], fl[K];
            const unsigned wcur = pk[i];
            unpack_v<D>(wcur, tab, ntab, v);
            const double wt = unpack_wt<D>(wcur, tab, ntab);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                fl[k] = fls[k * fstride + (i - soff)];
                f[k] = fs[k * fstride + (i - soff)] + dtv * fl[k];
                fs[k * fstride + (i - soff)] = f[k];
            }
            add_moments<D, K>(acc, wt, v, fl);
            add_moments<D, K>(acc + (D + 2), wt, v, f);
        }
    }
    if (pf_dist > 0 && threadIdx.x == NT - 1 && (cidx + pf_dist) * C < (int)gridDim.x) {
        const RegCell* __restrict__ nx = recs + cidx + pf_dist;
        int P2, q0, q1;
        chunk_range<C>(nx->n, Cluster<C>::rank(), P2, q0, q1);
        if (q1 > q0) {
            const long long nd = nx->doff;
            const long long nnp = nx->np;
            const unsigned nb = (unsigned)((q1 - q0 + 1) & ~1) * 8u;
#pragma unroll
            for (int k = 0; k < K; ++k) l2_prefetch(gdf + nd * K + k * nnp + q0, nb);
#pragma unroll
            for (int k = 0; k < K * D; ++k) l2_prefetch(gsl + nd * (K * D) + k * nnp + q0, nb);
        }
    }
    update_tail<D, K, MODE_FUSED, STAGE_SMEM, NT, C>(g, gas, p0, p1, np, rc.vol, c, dt, want_residual, acc, pk, tab, fs,
                                                     fstride, soff, fout, fch, red, xch, us, w_new, w0s);
}

// ------------------------------------------------------------------------------------------------
// iterate!(Euler), Theory/Iterate.jl:131-162: qf from the pre-convection f, single relaxation with prim(w^{n+1})
template <int D, int K>
__global__ void __launch_bounds__(256) euler_update_kernel(DevView g, GasPar gas, const int* __restrict__ cell_list,
                                                           double dt, int want_residual) {
    __shared__ double red[D * 32];
    __shared__ CellInfo ci;
    __shared__ UpdateShared<D, K> us;
    __shared__ double w_new[D + 2];
    const int c = cell_list[blockIdx.x];
    if (threadIdx.x == 0) ci = g.cells[c];
    __syncthreads();
    const CellPtr<D, K> own(g, ci);
    double* f = g.df + ci.doff * K;
    double* vflux = g.flux + ci.doff * K;
    const double dtv = dt / ci.vol;
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
        const double f0 = f[i];
        const double f1 = (K > 1) ? f[ci.np + i] : 0.0;
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
            f[k * ci.np + i] = (f[k * ci.np + i] + dtv * vflux[k * ci.np + i]) * a + b * (F[k] + Fp[k]);
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

// ------------------------------------------------------------------------------------------------
// iterate!(CIP_Marching), Theory/I-projection.jl:161-192: convection, heat flux, conserved I-projection of the
// h-component onto the updated conserved moments (solve_I_projection :55-141: Newton on the dual with Armijo
// backtracking; every Newton sum and every line-search objective is one block reduction), Shakhov relaxation.
// One CTA per cell, in place on g.df from the stored g.flux.  Every thread carries lambda and takes the (identical)
// scalar decisions from the broadcast reduction results, so the loop needs no extra broadcast.
// Donor cells of an immersed boundary first take positivity_preserving_ib! (Boundary/Positivity.jl).
template <int NV>
__device__ __forceinline__ void block_min(double (&v)[NV], double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] = fmin(v[k], __shfl_down_sync(0xffffffffu, v[k], off));
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) red[warp * NV + k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double x = red[k];
        for (int w = 1; w < nwarp; ++w) x = fmin(x, red[w * NV + k]);
        v[k] = x;
    }
    __syncthreads();
}

// Gaussian elimination with partial pivoting, M <= 5 (the oracle's orc_small_solve, same operation order)
template <int M>
__device__ __forceinline__ bool small_solve(double (&A)[M * M], double (&b)[M], double (&x)[M]) {
#pragma unroll
    for (int c = 0; c < M; ++c) {
        int piv = c;
        double best = fabs(A[c * M + c]);
#pragma unroll
        for (int r = c + 1; r < M; ++r)
            if (fabs(A[r * M + c]) > best) { best = fabs(A[r * M + c]); piv = r; }
        if (best == 0.0) return false;
#pragma unroll
        for (int r = c + 1; r < M; ++r) {
            if (r == piv) {
#pragma unroll
                for (int q = 0; q < M; ++q) { const double t = A[c * M + q]; A[c * M + q] = A[r * M + q]; A[r * M + q] = t; }
                const double t = b[c]; b[c] = b[r]; b[r] = t;
            }
        }
#pragma unroll
        for (int r = c + 1; r < M; ++r) {
            const double l = A[r * M + c] / A[c * M + c];
#pragma unroll
            for (int q = c; q < M; ++q) A[r * M + q] -= l * A[c * M + q];
            b[r] -= l * b[c];
        }
    }
#pragma unroll
    for (int r = M - 1; r >= 0; --r) {
        double t = b[r];
#pragma unroll
        for (int q = r + 1; q < M; ++q) t -= A[r * M + q] * x[q];
        x[r] = t / A[r * M + r];
    }
    return true;
}

template <int D>
__device__ __forceinline__ void psi_of(const double* v, double* psi) {
    double s2 = 0.0;
    psi[0] = 1.0;
#pragma unroll
    for (int d = 0; d < D; ++d) { psi[1 + d] = v[d]; s2 += v[d] * v[d]; }
    psi[D + 1] = s2 / 2;
}

// solve_I_projection, Theory/I-projection.jl:55-141, on the h-component plane `f` of one cell (one CTA): negative
// values shaved (:73-83) given the block minima mn = {min f, min positive f}, then the Newton iteration on the dual with
// Armijo backtracking (:88-136).  `lam` returns the multipliers; f_h *= exp(lam . psi) is left to the caller.  Each thread
// revisits only the points it owns (threadIdx.x + k blockDim.x), so no barrier is needed around the shave pass.
template <int D, int K>
__device__ __forceinline__ void solve_projection(const CellPtr<D, K>& own, double* __restrict__ f, int n, int np,
                                                 const double (&W)[D + 2], const double (&mn)[2], double* red,
                                                 double (&lam)[D + 2]) {
    constexpr int M = D + 2, NJ = M * (M + 1) / 2, NV = M + NJ;
    // ---- shave negative values (:73-83)
    const double f_min = 1.1 * mn[0];
    if (f_min < 0.) {
        const double fp = mn[1], dd = fp - f_min;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const double x = f[i];
            if (x < 0.) f[i] = (x - f_min) / dd * fp;
        }
    }
    // ---- Newton iteration on the dual (:88-136); each thread revisits only the points it wrote, so no barrier is needed
    // between the shave pass and the sums
#pragma unroll
    for (int m = 0; m < M; ++m) lam[m] = 0.0;
    {
        double nW = 0.0;
#pragma unroll
        for (int m = 0; m < M; ++m) nW += W[m] * W[m];
        nW = sqrt(nW);
        const double tol_eff = 1e-10 * fmax(1.0, nW);
        double G_prev = CUDART_INF;
        int stall = 0;
        for (int it = 0; it < 10; ++it) {
            double a[NV];
#pragma unroll
            for (int q = 0; q < NV; ++q) a[q] = 0.0;
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                double v[D], psi[M];
#pragma unroll
                for (int t = 0; t < D; ++t) v[t] = own.v[t * np + i];
                psi_of<D>(v, psi);
                double lp = 0.0;
#pragma unroll
                for (int m = 0; m < M; ++m) lp += lam[m] * psi[m];
                const double cc = own.wt[i] * f[i] * exp(lp);
                int q = M;
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const double cm = cc * psi[m];
                    a[m] += cm;
#pragma unroll
                    for (int j = m; j < M; ++j) a[q++] += cm * psi[j];
                }
            }
            block_reduce<NV>(a, red);
            const double Phi_sum = a[0];
            double G[M], Gn = 0.0;
#pragma unroll
            for (int m = 0; m < M; ++m) { G[m] = a[m] - W[m]; Gn += G[m] * G[m]; }
            Gn = sqrt(Gn);
            if (Gn < tol_eff) break;
            if (Gn > 0.9 * G_prev) {
                if (++stall >= 2) break;
            } else {
                stall = 0;
            }
            G_prev = Gn;
            double A[M * M], b[M], dl[M];
            {
                int q = M;
#pragma unroll
                for (int m = 0; m < M; ++m)
#pragma unroll
                    for (int j = m; j < M; ++j) { A[m * M + j] = a[q]; A[j * M + m] = a[q]; ++q; }
            }
#pragma unroll
            for (int m = 0; m < M; ++m) b[m] = G[m];
            if (!small_solve<M>(A, b, dl)) break;
            double lW = 0.0, slope = 0.0;
#pragma unroll
            for (int m = 0; m < M; ++m) { dl[m] = -dl[m]; lW += lam[m] * W[m]; slope += G[m] * dl[m]; }
            const double Phi0 = Phi_sum - lW;
            double alpha = 1.0;
            for (int ls = 0; ls < 10; ++ls) {
                double lt[M], ltW = 0.0;
#pragma unroll
                for (int m = 0; m < M; ++m) { lt[m] = lam[m] + alpha * dl[m]; ltW += lt[m] * W[m]; }
                double ph[1] = {0.0};
                for (int i = threadIdx.x; i < n; i += blockDim.x) {
                    double v[D], psi[M];
#pragma unroll
                    for (int t = 0; t < D; ++t) v[t] = own.v[t * np + i];
                    psi_of<D>(v, psi);
                    double lp = 0.0;
#pragma unroll
                    for (int m = 0; m < M; ++m) lp += lt[m] * psi[m];
                    ph[0] += own.wt[i] * f[i] * exp(lp);
                }
                block_reduce<1>(ph, red);
                if (ph[0] - ltW <= Phi0 + 1e-4 * alpha * slope) break;
                alpha *= 0.5;
            }
#pragma unroll
            for (int m = 0; m < M; ++m) lam[m] += alpha * dl[m];
        }
    }
}

template <int D, int K>
__global__ void __launch_bounds__(256) cip_update_kernel(DevView g, GasPar gas, const int* __restrict__ cell_list,
                                                         double dt, int want_residual) {
    constexpr int M = D + 2, NJ = M * (M + 1) / 2, NV = M + NJ;
    __shared__ double red[NV * 9];
    __shared__ CellInfo ci;
    __shared__ UpdateShared<D, K> us;
    __shared__ double w_new[M];
    const int c = cell_list[blockIdx.x];
    copy_words(g.cells + c, &ci, (int)(sizeof(CellInfo) / sizeof(int)));
    __syncthreads();
    const CellPtr<D, K> own(g, ci);
    const int n = ci.n, np = ci.np;
    double* f = g.df + ci.doff * K;
    double* vflux = g.flux + ci.doff * K;
    const double dtv = dt / ci.vol;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int m = 0; m < M; ++m)
            w_new[m] = g.w[(size_t)c * M + m] + g.mflux[(size_t)c * M + m] * dt / ci.vol;
    }
    if (ci.sn_count > 0) {
        // positivity_preserving_ib! (Boundary/Positivity.jl:1-43), donor cells only (block-uniform): the slope-
        // extrapolated correction of every SolidNeighbor face, micro = (sn.flux + ndx . sn.sdf) v_n A on the points the
        // wall side is upwind for (rot v_dir > 0), enters w and vs_data.flux limited by theta = min(theta_rho, theta_e)
        // so that density and internal energy of w stay positive
        __shared__ double theta_s;
        double we[M];
#pragma unroll
        for (int m = 0; m < M; ++m) we[m] = 0.0;
        for (int q = 0; q < ci.sn_count; ++q) {
            const DonorSn e = g.donor_sn[ci.sn_begin + q];
            const double* __restrict__ snflux = g.flux + e.doff * K;
            const double* __restrict__ snsdf = g.sdf + e.doff * (K * D);
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                double v[D], mic[K];
#pragma unroll
                for (int t = 0; t < D; ++t) v[t] = own.v[t * np + i];
                const double vn = pick<D>(v, e.dir);
                if (!(e.rot * vn > 0.)) continue;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    double dot = 0.0;
#pragma unroll
                    for (int t = 0; t < D; ++t) dot += (e.fmid[t] - v[t] * dt - e.snmid[t]) * snsdf[(t * K + k) * np + i];
                    mic[k] = (snflux[k * np + i] + dot) * vn * e.area;
                }
                add_moments<D, K>(we, own.wt[i], v, mic);
            }
        }
        block_reduce<M>(we, red);
        if (threadIdx.x == 0) {
            we[M - 1] *= 0.5;
#pragma unroll
            for (int m = 0; m < M; ++m) we[m] *= dt / ci.vol;
            const double delta = 1e-3;
            const double th_rho = we[0] > 0 ? 1.0 : fmin(1.0, (1 - delta) * w_new[0] / (fabs(we[0]) + EPS_MACH));
            double rub2 = 0.0, rube = 0.0, rue2 = 0.0;
#pragma unroll
            for (int d = 1; d <= D; ++d) { rub2 += w_new[d] * w_new[d]; rube += w_new[d] * we[d]; rue2 += we[d] * we[d]; }
            const double eb = w_new[M - 1] - rub2 / (2 * w_new[0]);
            const double ee = we[M - 1] - rube / w_new[0];
            const double gam = rue2 / (2 * w_new[0]);
            const double th_e = fmin(1.0, 2 * (1 - delta) * eb / (sqrt(ee * ee + 4 * gam * (1 - delta) * eb) - ee + EPS_MACH));
            const double th = fmin(th_rho, th_e);
#pragma unroll
            for (int m = 0; m < M; ++m) w_new[m] += th * we[m];
            theta_s = th;
        }
        __syncthreads();
        const double th = theta_s;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            double v[D], tot[K];
#pragma unroll
            for (int t = 0; t < D; ++t) v[t] = own.v[t * np + i];
#pragma unroll
            for (int k = 0; k < K; ++k) tot[k] = 0.0;
            for (int q = 0; q < ci.sn_count; ++q) {
                const DonorSn& e = g.donor_sn[ci.sn_begin + q];
                const double vn = pick<D>(v, e.dir);
                if (!(e.rot * vn > 0.)) continue;
                const double* __restrict__ snflux = g.flux + e.doff * K;
                const double* __restrict__ snsdf = g.sdf + e.doff * (K * D);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    double dot = 0.0;
#pragma unroll
                    for (int t = 0; t < D; ++t) dot += (e.fmid[t] - v[t] * dt - e.snmid[t]) * snsdf[(t * K + k) * np + i];
                    tot[k] += (snflux[k * np + i] + dot) * vn * e.area;
                }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) vflux[k * np + i] += th * tot[k];
        }
    }
    if (threadIdx.x == 0) {
        get_prim<D>(w_new, gas.gamma, us.prim_c);
        us.tau = gas.mu_ref * 2.0 * pow(us.prim_c[D + 1], 1 - gas.omega) / us.prim_c[0];
        us.coef_c = maxwell_coef<D>(us.prim_c);
        us.cb_c = gas.K / (2.0 * us.prim_c[D + 1]);
    }
    __syncthreads();
    double prim_c[M], W[M];
#pragma unroll
    for (int m = 0; m < M; ++m) { prim_c[m] = us.prim_c[m]; W[m] = w_new[m]; }
    // ---- convection (:177), heat flux of the convected f about prim_c (:179), internal energy of b (:145), minima
    double qe[D + 1], mn[2] = {CUDART_INF, CUDART_INF};
#pragma unroll
    for (int d = 0; d < D + 1; ++d) qe[d] = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double v[D], fk[K];
#pragma unroll
        for (int t = 0; t < D; ++t) v[t] = own.v[t * np + i];
        const double wt = own.wt[i];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            fk[k] = f[k * np + i] + dtv * vflux[k * np + i];
            f[k * np + i] = fk[k];
            vflux[k * np + i] = 0.0;
        }
        const double c2 = c2_of<D>(v, prim_c);
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const double cd = v[d] - prim_c[1 + d];
            qe[d] += wt * cd * c2 * fk[0] + ((K > 1) ? wt * cd * fk[1] : 0.0);
        }
        if (K > 1) qe[D] += wt * fk[1];
        mn[0] = fmin(mn[0], fk[0]);
        if (fk[0] > 0.) mn[1] = fmin(mn[1], fk[0]);
    }
    block_reduce<D + 1>(qe, red);
    block_min<2>(mn, red);
    double qf[D];
#pragma unroll
    for (int d = 0; d < D; ++d) qf[d] = 0.5 * qe[d];
    if (K > 1) W[M - 1] -= qe[D] / 2;
    double lam[M];
    solve_projection<D, K>(own, f, n, np, W, mn, red, lam);
    // ---- projection f_h *= exp(lambda.psi) (:147-149), relaxation towards M[prim_c] + S (:186-187)
    const double tau = us.tau;
    const double ra = tau / (tau + dt), rb = dt / (tau + dt);
    const double coef_c = us.coef_c;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double v[D], psi[M], F[K], Fp[K];
#pragma unroll
        for (int t = 0; t < D; ++t) v[t] = own.v[t * np + i];
        psi_of<D>(v, psi);
        double lp = 0.0;
#pragma unroll
        for (int m = 0; m < M; ++m) lp += lam[m] * psi[m];
        maxwell<D, K>(v, prim_c, coef_c, gas.K, F);
        shakhov<D, K>(v, F, prim_c, qf, gas.Pr, gas.K, Fp);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double x = f[k * np + i];
            if (k == 0) x *= exp(lp);
            f[k * np + i] = x * ra + rb * (F[k] + Fp[k]);
        }
    }
    if (threadIdx.x == 0) {
        double* prim_old = g.prim + (size_t)c * M;
#pragma unroll
        for (int d = 0; d < D; ++d) g.qf[(size_t)c * D + d] = qf[d];
        if (want_residual) {
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const double dd = prim_c[m] - prim_old[m];
                g.res_cell[(size_t)c * 2 * M + m] = dd * dd;
                g.res_cell[(size_t)c * 2 * M + M + m] = fabs(prim_c[m]);
            }
        }
#pragma unroll
        for (int m = 0; m < M; ++m) {
            g.w[(size_t)c * M + m] = w_new[m];
            prim_old[m] = prim_c[m];
            g.mflux[(size_t)c * M + m] = 0.0;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// slopes: one block per cell of the current dependency wave; all DIM directions in one pass.
//   sL = (1/nL) sum_nbr diff(f, P[f_nbr (+ dm . sdf_nbr)]) / dsL   (diff_vs!, Slope.jl:29-64, :278-333)
//   sdf = minmod(sL, sR) | sL | 0                                    (Slope.jl:90-116, 68-86, 471-472)
// minmod, Slope.jl:20-24: 0.5 (sign a + sign b) min(|a|,|b|)
__device__ __forceinline__ double minmod(double a, double b) {
    const double m = fmin(fabs(a), fabs(b));
    const bool pos = (a > 0.) && (b > 0.), neg = (a < 0.) && (b < 0.);
    return pos ? m : (neg ? -m : 0.0);
}

// N same-grid neighbours of one side, compile-time count: loads first, arithmetic after (a side of 2^(DIM-1) finer
// neighbours costs one memory round trip instead of one per neighbour).  The raw slopes of a projecting neighbour are
// requested with its value.
template <int D, int K, int N>
__device__ __forceinline__ void side_sum_fixed(const DevView& g, const SlopeNbr* nb, int i, const double* f, double* acc) {
    double nfv[N][K], nsv[N][K][D];
#pragma unroll
    for (int a = 0; a < N; ++a) {
        const double* __restrict__ nf = g.df + nb[a].doff * K + i;
        const double* __restrict__ nsd = g.sdf + nb[a].doff * (K * D) + i;
        const int np = nb[a].np;
        const bool pj = nb[a].proj != 0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            nfv[a][k] = nf[k * np];
#pragma unroll
            for (int t = 0; t < D; ++t) nsv[a][k][t] = pj ? nsd[(t * K + k) * np] : 0.0;
        }
    }
#pragma unroll
    for (int a = 0; a < N; ++a) {
        const bool pj = nb[a].proj != 0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double proj = nfv[a][k];
            if (pj) {
#pragma unroll
                for (int t = 0; t < D; ++t) proj += nb[a].dm[t] * nsv[a][k][t];
            }
            acc[k] += f[k] - proj;
        }
    }
}

// accumulated difference to the neighbours of one side: the general case (neighbours on other velocity grids, any count)
template <int K> struct KVec { double v[K]; };
template <int D, int K>
__device__ __forceinline__ KVec<K> side_sum_slow(const DevView& g, const SlopeNbr* nb, int cnt, int i, int li, KVec<K> f) {
    KVec<K> acc;
#pragma unroll
    for (int k = 0; k < K; ++k) acc.v[k] = 0.0;
    for (int a = 0; a < cnt; ++a) {
        const SlopeNbr& e = nb[a];
        const int np = e.np;
        if (e.rel_off < 0) {
            const double* __restrict__ nf = g.df + e.doff * K + i;
            const double* __restrict__ nsd = g.sdf + e.doff * (K * D) + i;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double proj = nf[k * np];
                if (e.proj) {
#pragma unroll
                    for (int t = 0; t < D; ++t) proj += e.dm[t] * nsd[(t * K + k) * np];
                }
                acc.v[k] += f.v[k] - proj;
            }
        } else {
            // pair-mapped neighbour (mismatched velocity grids): mean over the covering finer points or
            // injection from the coarser one (diff_vs!, Slope.jl:29-64)
            const double* nf = g.df + e.doff * K;
            const double* nsd = g.sdf + e.doff * K * D;
            const int8_t* nlev = g.v_level + e.goff;
            const int* st = g.pm_start + e.rel_off;
            const int j0 = st[i];
            const int cn = max(1, st[i + 1] - j0);
            for (int j = j0; j < j0 + cn; ++j) {
                const double scale = (cn > 1) ? 1.0 / (double)(1 << (D * (nlev[j] - li))) : 1.0;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    double proj = nf[k * np + j];
                    if (e.proj) {
#pragma unroll
                        for (int t = 0; t < D; ++t) proj += e.dm[t] * nsd[(t * K + k) * np + j];
                    }
                    acc.v[k] += (f.v[k] - proj) * scale;
                }
            }
        }
    }
    return acc;
}

template <int D, int K, bool GENERIC>
__device__ __forceinline__ void side_sum(const DevView& g, const SlopeNbr* nb, int cnt, int i, int li,
                                         const double* f, double* acc) {
    // Block-uniform fast path: every neighbour of the side lives on the cell's own velocity grid (point i <-> point i)
    constexpr int MAXN = 1 << (D - 1);
#if defined(KAMR_SLOPE_V2) || defined(KAMR_SLOPE_FIXED)
    bool same = cnt == 1 || cnt == MAXN;
#else
    bool same = false;
#endif
    if (GENERIC)
        for (int a = 0; a < cnt; ++a) same = same && nb[a].rel_off < 0;
    if (same) {
        if (cnt == 1) side_sum_fixed<D, K, 1>(g, nb, i, f, acc);
        else side_sum_fixed<D, K, MAXN>(g, nb, i, f, acc);
        return;
    }
    KVec<K> fv;
#pragma unroll
    for (int k = 0; k < K; ++k) fv.v[k] = f[k];
    const KVec<K> r = side_sum_slow<D, K>(g, nb, cnt, i, li, fv);
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] += r.v[k];
}

// Besides the reference's sdf (raw slopes) the kernel leaves the LIMITED slopes r*sdf in g.sdl, with
// r = min(|f - eps()| / (1/2 sum_t ds_t |sdf_t| + EPS), 1) of positivity_preserving_reconstruct
// (CAIDVM.jl:134-139): r depends on the cell's own f, sdf and ds only, so evaluating it here once per
// point replaces the 1 + 2*DIM evaluations (with their fp64 divisions) the flux gather would need.
// Raw sdf is written only where something reads it (tk.flags bit0: coarse neighbours of projecting cells,
// cells on domain / solid faces, halo mirrors) or when the host asked for it (raw_all: kamr_slope,
// KAMR_OPT_KEEP_SDF).
//
// Dependency sweep in ONE launch: tasks are ordered by wave; a task that projects the finished slopes of coarser
// cells computed by this same launch waits on their per-cell epoch flags (thread t spins on dependency t), and
// every task publishes its own flag when its stores are visible.  CUDA guarantees neither a dispatch order nor
// forward progress between CTAs, so a CTA does not take the task of its block index: it draws a ticket from a
// device-wide counter once it is running.  A task only depends on tasks earlier in the list, whose tickets were
// therefore drawn by CTAs that are already resident, so every wait is on a running (or finished) CTA and the sweep
// cannot deadlock under MPS, time-slicing or a debugger.  The spin is bounded all the same: on expiry the kernel
// raises g.err_flag (surfaced by the next kamr_sync / download / residual read as an error) instead of hanging.
constexpr unsigned SLOPE_SPIN_LIMIT = 1u << 25;   // x 64 ns sleep: seconds
#ifndef KAMR_SLOPE_THREADS
#ifdef KAMR_SLOPE_V2
#define KAMR_SLOPE_THREADS 768
#else
#define KAMR_SLOPE_THREADS 1024
#endif
#endif
template <int D, int K, bool GENERIC, int NT>
__global__ void __launch_bounds__(NT, KAMR_SLOPE_THREADS / NT) slope_kernel(DevView g, const SlopeTask* __restrict__ tasks, int raw_all,
                                                   int epoch, unsigned* __restrict__ ticket, unsigned ticket_base) {
    __shared__ SlopeTask tk;
    __shared__ CellInfo ci;
    __shared__ SlopeNbr nb[MAX_SLOPE_NB];
    __shared__ int s_task;
    int ti = blockIdx.x;
    if (epoch) {
        if (threadIdx.x == 0) s_task = (int)(atomicAdd(ticket, 1u) - ticket_base);
        __syncthreads();
        ti = s_task;
    }
    copy_words(tasks + ti, &tk, (int)(sizeof(SlopeTask) / sizeof(int)));
    __syncthreads();
    if (tk.dep_count > 0) {
        if ((int)threadIdx.x < tk.dep_count) {
            const volatile int* flag = g.slope_done + g.slope_deps[tk.dep_begin + threadIdx.x];
            unsigned spins = 0;
            while (*flag != epoch) {
                __nanosleep(64);
                if (++spins > SLOPE_SPIN_LIMIT) { atomicExch(g.err_flag, 1); break; }
            }
            __threadfence();
        }
        __syncthreads();
    }
    copy_words(g.cells + tk.cell, &ci, (int)(sizeof(CellInfo) / sizeof(int)));
    int base[D];
    {
        int b = 0;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const int cnt = tk.d[d].nA + tk.d[d].nB;
            copy_words(g.slope_nb + tk.d[d].nb_begin, nb + b, cnt * (int)(sizeof(SlopeNbr) / sizeof(int)));
            base[d] = b;
            b += cnt;
        }
    }
    __syncthreads();
    const CellPtr<D, K> own(g, ci);
    const int n = ci.n, np = ci.np;
    const bool raw = raw_all || (tk.flags & 1);
    double* sdf = g.sdf + ci.doff * K * D;
    double* sdl = g.sdl + ci.doff * K * D;
    // Block-uniform fast path: every direction is an inner stencil with ONE neighbour per side and no transverse
    // projection (same-level neighbours; their velocity grids may differ).  Indices into the neighbours' grids are
    // resolved for all 2*DIM sides first (identity or pair map), then all neighbour values are requested together.
    bool simple = true;
#pragma unroll
    for (int d = 0; d < D; ++d)
        simple = simple && tk.d[d].mode == SLOPE_INNER && tk.d[d].nA == 1 && tk.d[d].nB == 1 && !nb[base[d]].proj &&
                 !nb[base[d] + 1].proj;
    if (simple) {
        const double* __restrict__ df = g.df;
        for (int i = threadIdx.x; i < n; i += NT) {
            double f[K], s[D][K];
            int j0[2 * D], cn[2 * D];
#pragma unroll
            for (int q = 0; q < 2 * D; ++q) {
                const SlopeNbr& e = nb[base[q >> 1] + (q & 1)];
                j0[q] = i; cn[q] = 1;
                if (GENERIC && e.rel_off >= 0) {
                    const int* __restrict__ st = g.pm_start + e.rel_off;
                    j0[q] = st[i];
                    cn[q] = max(1, st[i + 1] - j0[q]);
                }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) f[k] = own.f[k * np + i];
            double nfv[2 * D][K];
#pragma unroll
            for (int q = 0; q < 2 * D; ++q) {
                const SlopeNbr& e = nb[base[q >> 1] + (q & 1)];
                const double* __restrict__ p = df + e.doff * K + j0[q];
#pragma unroll
                for (int k = 0; k < K; ++k) nfv[q][k] = p[k * e.np];
            }
#pragma unroll
            for (int d = 0; d < D; ++d) {
                double sAB[2][K];
#pragma unroll
                for (int sd2 = 0; sd2 < 2; ++sd2) {
                    const int q = 2 * d + sd2;
                    if (cn[q] == 1) {
#pragma unroll
                        for (int k = 0; k < K; ++k) sAB[sd2][k] = 0.0 + (f[k] - nfv[q][k]);
                    } else {  // covered by several finer points of the neighbour's grid: mean, diff_vs! Slope.jl:29-64
                        const SlopeNbr& e = nb[base[d] + sd2];
                        const int8_t* __restrict__ nlev = g.v_level + e.goff;
                        const int li = own.lev[i];
#pragma unroll
                        for (int k = 0; k < K; ++k) sAB[sd2][k] = 0.0;
                        for (int j = j0[q]; j < j0[q] + cn[q]; ++j) {
                            const double scale = 1.0 / (double)(1 << (D * (nlev[j] - li)));
#pragma unroll
                            for (int k = 0; k < K; ++k) sAB[sd2][k] += (f[k] - df[e.doff * K + k * e.np + j]) * scale;
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    s[d][k] = minmod(sAB[0][k] * tk.d[d].invA, sAB[1][k] * tk.d[d].invB);
                    if (raw) sdf[(d * K + k) * np + i] = s[d][k];
                }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double s_abs = 0.0;
#pragma unroll
                for (int d = 0; d < D; ++d) s_abs += ci.ds[d] * fabs(s[d][k]);
                const double r = limiter(f[k], s_abs);
#pragma unroll
                for (int d = 0; d < D; ++d) sdl[(d * K + k) * np + i] = r * s[d][k];
            }
        }
    } else
    for (int i = threadIdx.x; i < n; i += NT) {
        double f[K], s[D][K];
#pragma unroll
        for (int k = 0; k < K; ++k) f[k] = own.f[k * np + i];
        const int li = GENERIC ? (int)own.lev[i] : 0;
        // (the direction loop stays rolled: s[d] is indexed at run time and lives in local memory, a few L1-resident
        // bytes per point, in exchange for a kernel a third of the size — unrolled it overflowed the instruction cache)
#if defined(KAMR_SLOPE_V2) || defined(KAMR_SLOPE_ROLL)
#pragma unroll 1
#else
#pragma unroll
#endif
        for (int d = 0; d < D; ++d) {
            const SlopeDir& sd = tk.d[d];
            const int mode = sd.mode;
            if (mode == SLOPE_KEEP) {  // untouched by the reference's sweep: keep what is stored
#pragma unroll
                for (int k = 0; k < K; ++k) s[d][k] = sdf[(d * K + k) * np + i];
                continue;
            }
            double sB[K];
#pragma unroll
            for (int k = 0; k < K; ++k) { s[d][k] = 0.0; sB[k] = 0.0; }
            if (mode != SLOPE_ZERO) {
                side_sum<D, K, GENERIC>(g, nb + base[d], sd.nA, i, li, f, s[d]);
#pragma unroll
                for (int k = 0; k < K; ++k) s[d][k] *= sd.invA;
                if (mode == SLOPE_INNER) {
                    side_sum<D, K, GENERIC>(g, nb + base[d] + sd.nA, sd.nB, i, li, f, sB);
#pragma unroll
                    for (int k = 0; k < K; ++k) s[d][k] = minmod(s[d][k], sB[k] * sd.invB);
                }
            }
            if (raw) {
#pragma unroll
                for (int k = 0; k < K; ++k) sdf[(d * K + k) * np + i] = s[d][k];
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double s_abs = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d) s_abs += ci.ds[d] * fabs(s[d][k]);
            const double r = limiter(f[k], s_abs);
#pragma unroll
            for (int d = 0; d < D; ++d) sdl[(d * K + k) * np + i] = r * s[d][k];
        }
    }
    if (epoch) {  // publish: every thread's stores first, then the flag
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            *((volatile int*)(g.slope_done + tk.cell)) = epoch;
        }
    }
}

// slopes of REGULAR stencils (SlopeReg): same arithmetic as slope_kernel's SLOPE_INNER branch with one neighbour per
// side, all 2*DIM neighbour values requested up front.  MAPPED (SlopeRegMap): neighbours may live on other velocity
// grids; their point indices come from the pair maps (mean over covering finer points, diff_vs! Slope.jl:29-64).
template <int D, int K, int NT, bool MAPPED>
__global__ void __launch_bounds__(NT) slope_regular_kernel(DevView g, const void* __restrict__ tasks_, int raw_all,
                                                           int pf_dist) {
    using Task = typename std::conditional<MAPPED, SlopeRegMap, SlopeReg>::type;
    const Task* __restrict__ tasks = reinterpret_cast<const Task*>(tasks_);
    __shared__ Task tkm;
    copy_words(tasks + blockIdx.x, &tkm, (int)(sizeof(Task) / sizeof(int)));
    __syncthreads();
    const SlopeReg& tk = *reinterpret_cast<const SlopeReg*>(&tkm);
    const SlopeRegMap& tm = *reinterpret_cast<const SlopeRegMap*>(&tkm);   // only read when MAPPED
    const int n = tk.n, np = tk.np;
    const bool raw = raw_all || (tk.flags & 1);
    const double* __restrict__ df = g.df;
    if (pf_dist > 0 && threadIdx.x == NT - 1 && blockIdx.x + pf_dist < gridDim.x) {
        const SlopeReg* __restrict__ nx = reinterpret_cast<const SlopeReg*>(tasks + blockIdx.x + pf_dist);
        l2_prefetch(df + nx->doff * K, (unsigned)nx->np * 8u * K);
    }
    const double* __restrict__ own = df + tk.doff * K;
    double* sdf = g.sdf + tk.doff * K * D;
    double* sdl = g.sdl + tk.doff * K * D;
#ifdef KAMR_PAIRS   // (measured slower, see DESIGN.md section 6 "tried and not adopted")
    if (!MAPPED) {
        // Identical grids on all sides: point i of every neighbour is point i.  A thread takes the points i, i+1 with
        // 128-bit loads and stores (planes are 32-byte aligned and padded to a multiple of 4 points, so i+1 exists in
        // memory even when it is padding): half the load / store instructions and address arithmetic per point.
        for (int i = 2 * threadIdx.x; i < n; i += 2 * NT) {
            double2 f2[K], nf2[2 * D][K];
#pragma unroll
            for (int q = 0; q < 2 * D; ++q) {
                const double* __restrict__ p = df + tk.nb_doff[q] * K + i;
#pragma unroll
                for (int k = 0; k < K; ++k) nf2[q][k] = *reinterpret_cast<const double2*>(p + k * np);
            }
#pragma unroll
            for (int k = 0; k < K; ++k) f2[k] = *reinterpret_cast<const double2*>(own + k * np + i);
            double2 s2[D][K], l2[D][K];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                double f[K], s[D][K];
#pragma unroll
                for (int k = 0; k < K; ++k) f[k] = h ? f2[k].y : f2[k].x;
#pragma unroll
                for (int d = 0; d < D; ++d) {
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const double a = 0.0 + (f[k] - (h ? nf2[2 * d][k].y : nf2[2 * d][k].x));
                        const double b = 0.0 + (f[k] - (h ? nf2[2 * d + 1][k].y : nf2[2 * d + 1][k].x));
                        s[d][k] = minmod(a * tk.inv[2 * d], b * tk.inv[2 * d + 1]);
                        if (h) s2[d][k].y = s[d][k]; else s2[d][k].x = s[d][k];
                    }
                }
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    double s_abs = 0.0;
#pragma unroll
                    for (int d = 0; d < D; ++d) s_abs += tk.ds[d] * fabs(s[d][k]);
                    const double r = limiter(f[k], s_abs);
#pragma unroll
                    for (int d = 0; d < D; ++d) { if (h) l2[d][k].y = r * s[d][k]; else l2[d][k].x = r * s[d][k]; }
                }
            }
#pragma unroll
            for (int d = 0; d < D; ++d)
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    if (raw) *reinterpret_cast<double2*>(sdf + (d * K + k) * np + i) = s2[d][k];
                    *reinterpret_cast<double2*>(sdl + (d * K + k) * np + i) = l2[d][k];
                }
        }
        return;
    }
#endif
    for (int i = threadIdx.x; i < n; i += NT) {
        double f[K], nf[2 * D][K], s[D][K];
        int j0[2 * D], cn[2 * D];
#pragma unroll
        for (int q = 0; q < 2 * D; ++q) {
            j0[q] = i; cn[q] = 1;
            if (MAPPED && tm.nb_rel[q] >= 0) {
                const int* __restrict__ st = g.pm_start + tm.nb_rel[q];
                j0[q] = st[i];
                cn[q] = max(1, st[i + 1] - j0[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < 2 * D; ++q) {
            const double* __restrict__ p = df + tk.nb_doff[q] * K + j0[q];
            const int nnp = MAPPED ? tm.nb_np[q] : np;
#pragma unroll
            for (int k = 0; k < K; ++k) nf[q][k] = p[k * nnp];
        }
#pragma unroll
        for (int k = 0; k < K; ++k) f[k] = own[k * np + i];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            double ab[2][K];
#pragma unroll
            for (int sd2 = 0; sd2 < 2; ++sd2) {
                const int q = 2 * d + sd2;
                if (!MAPPED || cn[q] == 1) {
#pragma unroll
                    for (int k = 0; k < K; ++k) ab[sd2][k] = 0.0 + (f[k] - nf[q][k]);
                } else {
                    const int8_t* __restrict__ nlev = g.v_level + tm.nb_goff[q];
                    const int li = g.v_level[tm.goff + i];
                    const int nnp = tm.nb_np[q];
#pragma unroll
                    for (int k = 0; k < K; ++k) ab[sd2][k] = 0.0;
                    for (int j = j0[q]; j < j0[q] + cn[q]; ++j) {
                        const double scale = 1.0 / (double)(1 << (D * (nlev[j] - li)));
#pragma unroll
                        for (int k = 0; k < K; ++k) ab[sd2][k] += (f[k] - df[tk.nb_doff[q] * K + k * nnp + j]) * scale;
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                s[d][k] = minmod(ab[0][k] * tk.inv[2 * d], ab[1][k] * tk.inv[2 * d + 1]);
                if (raw) sdf[(d * K + k) * np + i] = s[d][k];
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double s_abs = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d) s_abs += tk.ds[d] * fabs(s[d][k]);
            const double r = limiter(f[k], s_abs);
#pragma unroll
            for (int d = 0; d < D; ++d) sdl[(d * K + k) * np + i] = r * s[d][k];
        }
    }
}

// sdl = r * sdf for cells whose raw slopes came from outside the slope kernel (ghost cells after the halo
// exchange, kamr_upload_aux)
template <int D, int K>
__global__ void __launch_bounds__(256) limit_kernel(DevView g, const int* __restrict__ cell_list) {
    __shared__ CellInfo ci;
    copy_words(g.cells + cell_list[blockIdx.x], &ci, (int)(sizeof(CellInfo) / sizeof(int)));
    __syncthreads();
    const CellPtr<D, K> own(g, ci);
    const int n = ci.n, np = ci.np;
    double* sdl = g.sdl + ci.doff * K * D;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double s[D], s_abs = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d) { s[d] = own.s[(d * K + k) * np + i]; s_abs += ci.ds[d] * fabs(s[d]); }
            const double r = limiter(own.f[k * np + i], s_abs);
#pragma unroll
            for (int d = 0; d < D; ++d) sdl[(d * K + k) * np + i] = r * s[d];
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
// kernel (d): immersed boundary.
//
// directional_extrapolate: value at point x on the target grid = sum_a w_a(v) (f_a + sdf_a . (x - x_a)) over the
// fluid cells a, w_a = max(0, v.l_a/|v|)^2 normalised over a, l_a the unit vector from x_a to x
// (update_solid_cell! :122-141, image_df :376-395, vs_extrapolate! :98-121 as a pair-map gather).
// direction weight max(0, v.l/|v|)^2 of one fluid cell, l the unit vector from its centre to x.  Evaluated with
// explicitly rounded operations (no FMA contraction) in the order of the reference expression
// `max(0., dot(u,l)/norm(u))^2` with `l /= norm(l)`: where v is perpendicular to l the weight is pure rounding noise
// (~1e-34) yet the reference normalises by the sum and only tests it against exactly 0, so the result at such points
// is decided by the last bit; matching the oracle there needs the same bits, not the same formula.
// Per-neighbour geometry of one extrapolation target x: l = x - x_a and l/|l|.  Point-independent, so each CTA
// evaluates it once (same operations, same bits as evaluating it per point).
template <int D>
struct IbGeom { double l[D], lh[D]; };
template <int D>
__device__ __forceinline__ void ib_geometry(const IbNbr* nb, int cnt, const double* x, IbGeom<D>* geo) {
    if ((int)threadIdx.x < cnt) {
        const int a = threadIdx.x;
        double l[D], nl = 0.0;
#pragma unroll
        for (int t = 0; t < D; ++t) { l[t] = __dsub_rn(x[t], nb[a].mid[t]); nl = __dadd_rn(nl, __dmul_rn(l[t], l[t])); }
        nl = __dsqrt_rn(nl);
#pragma unroll
        for (int t = 0; t < D; ++t) { geo[a].l[t] = l[t]; geo[a].lh[t] = __ddiv_rn(l[t], nl); }
    }
}
// v.l/|v| as a product with 1/|v| (one division per point instead of one per neighbour): the sign and the exact zeros
// of v.l, which decide the weights where v is perpendicular to l, are unchanged; magnitudes move by <= 1 ulp.
template <int D>
__device__ __forceinline__ double dir_weight(const IbGeom<D>& ge, const double* v, double inv_nu) {
    double dot = 0.0;
#pragma unroll
    for (int t = 0; t < D; ++t) dot = __dadd_rn(dot, __dmul_rn(v[t], ge.lh[t]));
    double q = __dmul_rn(dot, inv_nu);
    q = q > 0. ? q : 0.;
    return __dmul_rn(q, q);
}

#ifndef KAMR_IB_MINB
#define KAMR_IB_MINB 4   // 64 registers: some spill, twice the warps; measured best on S2 and S4 (gpurun_out/sweep11)
#endif
constexpr int IB_WREG = 4;  // direction weights of the first IB_WREG neighbours stay in registers between the passes

template <int D, int K>
__device__ __forceinline__ void directional_extrapolate(const DevView& g, const IbNbr* nb, const IbGeom<D>* geo,
                                                        int cnt, const double* v, int li, int i, double* out) {
    double nu = 0.0;
#pragma unroll
    for (int t = 0; t < D; ++t) nu = __dadd_rn(nu, __dmul_rn(v[t], v[t]));
    nu = __ddiv_rn(1.0, __dsqrt_rn(nu));
    double ws = 0.0, wreg[IB_WREG];
#pragma unroll
    for (int a = 0; a < IB_WREG; ++a) {
        wreg[a] = 0.0;
        if (a < cnt) { wreg[a] = dir_weight<D>(geo[a], v, nu); ws = __dadd_rn(ws, wreg[a]); }
    }
    for (int a = IB_WREG; a < cnt; ++a) ws = __dadd_rn(ws, dir_weight<D>(geo[a], v, nu));
#pragma unroll
    for (int k = 0; k < K; ++k) out[k] = 0.0;
    const double inv_ws = (ws == 0.) ? 0.0 : __ddiv_rn(1.0, ws);
    const double w_eq = 1.0 / cnt;
    auto gather = [&](int a, double wa) {
        const IbNbr& e = nb[a];
        const double wi = (ws == 0.) ? w_eq : __dmul_rn(wa, inv_ws);
        const double* sf = g.df + e.doff * K;
        const double* ss = g.sdf + e.doff * K * D;
        const int np = e.np;
        int j0 = i, cn = 1;
        if (e.rel_off >= 0) {
            const int* st = g.pm_start + e.rel_off;
            j0 = st[i];
            cn = max(1, st[i + 1] - j0);
        }
        const int8_t* slev = g.v_level + e.goff;
        for (int j = j0; j < j0 + cn; ++j) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double ddf = 0.0;
#pragma unroll
                for (int t = 0; t < D; ++t) ddf += ss[(t * K + k) * np + j] * geo[a].l[t];
                double val = sf[k * np + j] + ddf;
                if (cn > 1) val = val / (double)(1 << (D * (slev[j] - li)));
                out[k] += val * wi;
            }
        }
    };
#pragma unroll
    for (int a = 0; a < IB_WREG; ++a)
        if (a < cnt) gather(a, wreg[a]);
    for (int a = IB_WREG; a < cnt; ++a) gather(a, dir_weight<D>(geo[a], v, nu));
}

// update_solid_cell!: one CTA per solid ghost cell; also w = <psi f>, prim (Immersed_boundary.jl:139-140)
template <int D, int K>
__global__ void __launch_bounds__(256, KAMR_IB_MINB) solid_cell_kernel(DevView g, GasPar gas, const SolidTask* __restrict__ tasks,
                                                         double* __restrict__ df2) {
    __shared__ double red[(D + 2) * 32];
    __shared__ CellInfo ci;
    __shared__ SolidTask tk;
    __shared__ IbNbr nb[32];
    __shared__ IbGeom<D> geo[32];
    copy_words(tasks + blockIdx.x, &tk, (int)(sizeof(SolidTask) / sizeof(int)));
    __syncthreads();
    copy_words(g.cells + tk.cell, &ci, (int)(sizeof(CellInfo) / sizeof(int)));
    copy_words(g.ib_nb + tk.nb_begin, nb, tk.nb_count * (int)(sizeof(IbNbr) / sizeof(int)));
    __syncthreads();
    ib_geometry<D>(nb, tk.nb_count, ci.mid, geo);
    __syncthreads();
    const CellPtr<D, K> own(g, ci);
    const int n = ci.n, np = ci.np;
    double* out = g.df + ci.doff * K;
    double acc[D + 2];
#pragma unroll
    for (int q = 0; q < D + 2; ++q) acc[q] = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double v[D], f[K];
#pragma unroll
        for (int t = 0; t < D; ++t) v[t] = own.v[t * np + i];
        directional_extrapolate<D, K>(g, nb, geo, tk.nb_count, v, own.lev[i], i, f);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            out[k * np + i] = f[k];
            if (df2) df2[ci.doff * K + k * np + i] = f[k];
        }
        add_moments<D, K>(acc, own.wt[i], v, f);
    }
    block_reduce<D + 2>(acc, red);
    if (threadIdx.x == 0) {
        acc[D + 1] *= 0.5;
        double prim[D + 2];
        get_prim<D>(acc, gas.gamma, prim);
#pragma unroll
        for (int q = 0; q < D + 2; ++q) {
            g.w[(size_t)tk.cell * (D + 2) + q] = acc[q];
            g.prim[(size_t)tk.cell * (D + 2) + q] = prim[q];
        }
    }
}

// position of point i in the sorted cut-cell list [b, b+c), or -1
__device__ __forceinline__ int cvc_find(const int* __restrict__ idx, int b, int c, int i) {
    int lo = 0, hi = c;
    while (lo < hi) {
        const int m = (lo + hi) >> 1;
        if (idx[b + m] < i) lo = m + 1; else hi = m;
    }
    return (lo < c && idx[b + lo] == i) ? b + lo : -1;
}

// update_solid_neighbor!: one CTA per (donor cell, solid face).  Pass 1: image-point value, one-sided limited slope,
// extrapolation to the wall point, partial sums of the wall mass balance; pass 2: wall Maxwellian on the outgoing
// half, cut-cell blending.  Writes SolidNeighbor.vs_data.df / .sdf[:,:,dir] / .flux.
template <int D, int K>
__global__ void __launch_bounds__(256, KAMR_IB_MINB) solid_neighbor_kernel(DevView g, GasPar gas, const SnTask* __restrict__ tasks,
                                                             double* __restrict__ df2) {
    __shared__ double red[2 * 32];
    __shared__ CellInfo cp, cs, cn_;
    __shared__ SnTask tk;
    __shared__ IbNbr nb[8];
    __shared__ IbGeom<D> geo[8];
    __shared__ double rho_w_s;
    copy_words(tasks + blockIdx.x, &tk, (int)(sizeof(SnTask) / sizeof(int)));
    __syncthreads();
    copy_words(g.cells + tk.donor, &cp, (int)(sizeof(CellInfo) / sizeof(int)));
    copy_words(g.cells + tk.solid, &cs, (int)(sizeof(CellInfo) / sizeof(int)));
    copy_words(g.cells + tk.sn_cell, &cn_, (int)(sizeof(CellInfo) / sizeof(int)));
    copy_words(g.ib_nb + tk.nb_begin, nb, tk.nb_count * (int)(sizeof(IbNbr) / sizeof(int)));
    __syncthreads();
    const CellPtr<D, K> own(g, cp);
    const int n = cp.n, np = cp.np, dir = tk.dir;
    double ibp[D];
#pragma unroll
    for (int t = 0; t < D; ++t) ibp[t] = tk.aux[t] + cp.mid[t] - cs.mid[t];
    ib_geometry<D>(nb, tk.nb_count, ibp, geo);
    __syncthreads();
    const double dxf = pick<D>(ibp, dir) - pick<D>(cp.mid, dir);
    const double dxs = pick<D>(cp.mid, dir) - pick<D>(cn_.mid, dir);
    const double dxL = pick<D>(tk.aux, dir) - pick<D>(ibp, dir);
    const double dfl = pick<D>(cn_.mid, dir) - pick<D>(tk.aux, dir);
    const double inv_dxs = 1.0 / dxs, inv_dxf = 1.0 / dxf;   // block-uniform: products instead of per-point divisions
    const double* sdfS = g.df + cs.doff * K;   // solid cell's df
    const int8_t* slev = g.v_level + cs.goff;
    double* snf = g.df + cn_.doff * K;
    double* sns = g.sdf + cn_.doff * K * D + (size_t)dir * K * np;
    double* snflux = g.flux ? g.flux + cn_.doff * K : nullptr;   // read by positivity_preserving_ib! only (un-fused path)
    double bc[D + 2];
#pragma unroll
    for (int q = 0; q < D + 2; ++q) bc[q] = tk.bc[q];
    bc[0] = 1.0;
    const double coef = maxwell_coef<D>(bc);
    double acc[2] = {0.0, 0.0};  // SF, MuR
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double v[D], f[K], ibf[K];
        double vn = 0.0;  // v.n without FMA: its sign at points on the plane v.n = 0 is a last-bit matter (see dir_weight)
#pragma unroll
        for (int t = 0; t < D; ++t) { v[t] = own.v[t * np + i]; vn = __dadd_rn(vn, __dmul_rn(v[t], tk.normal[t])); }
        const int li = own.lev[i];
        directional_extrapolate<D, K>(g, nb, geo, tk.nb_count, v, li, i, ibf);
        int j0 = i, cc = 1;
        if (tk.rel_ps >= 0) {
            const int* st = g.pm_start + tk.rel_ps;
            j0 = st[i];
            cc = max(1, st[i + 1] - j0);
        }
        double aux0 = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            f[k] = own.f[k * np + i];
            double sL = 0.0;  // boundary_slope!, :396-435
            for (int j = j0; j < j0 + cc; ++j) {
                double d = f[k] - sdfS[k * cs.np + j];
                if (cc > 1) d = d / (double)(1 << (D * (slev[j] - li)));
                sL += d * inv_dxs;
            }
            double sv = minmod(sL, (ibf[k] - f[k]) * inv_dxf);
            sv = fmin(fabs((ibf[k] - EPS_MACH) / (sv * dxL + EPS_MACH)), 1.0) * sv;  // :452-457
            const double a = ibf[k] + sv * dxL;
            sns[k * np + i] = sv;
            if (snflux) snflux[k * np + i] = sv * dfl;                              // :474-475
            snf[k * np + i] = a;
            if (k == 0) aux0 = a;
        }
        const double M0 = coef * exp_nonpos(-bc[D + 1] * c2_of<D>(v, bc));
        const int q = cvc_find(g.cvc_index, tk.cvc_begin, tk.cvc_count, i);
        if (q >= 0) {  // cut velocity cell: gas part feeds SF, solid part MuR (cvc_density :338, cvc_Mu :347)
            acc[0] += g.cvc_gas_w[q] * vn * aux0;
            acc[1] += g.cvc_solid_w[q] * vn * M0;
        } else if (vn >= 0.) {
            acc[1] += own.wt[i] * vn * M0;
        } else {
            acc[0] += own.wt[i] * vn * aux0;
        }
    }
    block_reduce<2>(acc, red);
    if (threadIdx.x == 0) rho_w_s = -acc[0] / acc[1];
    __syncthreads();
    const double rho_w = rho_w_s;
    const double cb = gas.K / (2.0 * bc[D + 1]);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double v[D], vn = 0.0;
#pragma unroll
        for (int t = 0; t < D; ++t) { v[t] = own.v[t * np + i]; vn = __dadd_rn(vn, __dmul_rn(v[t], tk.normal[t])); }
        const int q = cvc_find(g.cvc_index, tk.cvc_begin, tk.cvc_count, i);
        double Mw[K];
        const double M1 = coef * exp_nonpos(-bc[D + 1] * c2_of<D>(v, bc));
        Mw[0] = M1 * rho_w;
        if (K > 1) Mw[1] = (M1 * cb) * rho_w;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double a = snf[k * np + i];
            if (q >= 0) {  // cvc_correction!, :360-365
                const double gw = g.cvc_gas_w[q], sw = g.cvc_solid_w[q];
                a = (gw * a + sw * Mw[k]) / (gw + sw);
            } else if (vn >= 0.) {
                a = Mw[k];
            }
            snf[k * np + i] = a;
            if (df2) df2[cn_.doff * K + k * np + i] = a;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// residual sums over cells (residual_check! accumulators): one block per accumulator, deterministic
__global__ void __launch_bounds__(1024) residual_reduce_kernel(const double* __restrict__ res_cell,
                                                               const int* __restrict__ cell_list, int ncell, int nv,
                                                               double* out) {
    __shared__ double red[33];
    const int q = blockIdx.x;
    double a[1] = {0.0};
    for (int t = threadIdx.x; t < ncell; t += blockDim.x) a[0] += res_cell[(size_t)cell_list[t] * nv + q];
    block_reduce<1>(a, red);
    if (threadIdx.x == 0) out[q] = a[0];
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

// One-sided halo: segments of a local array stored into the peers' arrays (peer pointers mapped with CUDA IPC; the
// stores travel over NVLink).  One block per segment, grid-stride over segments; 128-bit stores (blocks are 32-byte
// aligned and their lengths multiples of 4 doubles).
__global__ void __launch_bounds__(256) put_segments_kernel(const PutSeg* __restrict__ segs, int nseg,
                                                           const double* __restrict__ src, double* const* __restrict__ dst) {
    for (int s = blockIdx.x; s < nseg; s += gridDim.x) {
        const PutSeg sg = segs[s];
        const double2* a = reinterpret_cast<const double2*>(src + sg.src);
        double2* b = reinterpret_cast<double2*>(dst[sg.peer] + sg.dst);
        for (int t = threadIdx.x; t < sg.len / 2; t += blockDim.x) b[t] = a[t];
    }
}
// the same for short blocks of any length (macro slopes: (DIM+2)*DIM doubles per cell): 8 segments per block
__global__ void __launch_bounds__(256) put_small_kernel(const PutSeg* __restrict__ segs, int nseg,
                                                        const double* __restrict__ src, double* const* __restrict__ dst) {
    const int s = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (s >= nseg) return;
    const PutSeg sg = segs[s];
    for (int t = threadIdx.x & 31; t < sg.len; t += 32) dst[sg.peer][sg.dst + t] = src[sg.src + t];
}
// After the puts of one message: make them visible system-wide, then raise the message's flag in every receiving
// peer's memory (slot [my rank][kind] of its flag table) to `epoch`.  Stream order puts this after the put kernel.
__global__ void halo_signal_kernel(int* const* __restrict__ peer_flags, const int* __restrict__ peer_sel, int nsel, int slot,
                                   int epoch) {
    if ((int)threadIdx.x < nsel) {
        __threadfence_system();
        volatile int* f = peer_flags[peer_sel[threadIdx.x]] + slot;
        *f = epoch;
    }
}
// Before the first kernel that reads a message's ghost blocks: wait until every sending peer has raised its flag.
// Bounded: on expiry the kernel raises err_flag (surfaced at the next sync point) instead of hanging the device.
constexpr unsigned HALO_SPIN_LIMIT = 1u << 26;
__global__ void halo_wait_kernel(const int* __restrict__ flags, const int* __restrict__ slots, int nsel, int epoch,
                                 int* err_flag) {
    if ((int)threadIdx.x < nsel) {
        const volatile int* f = flags + slots[threadIdx.x];
        unsigned spins = 0;
        while (*f - epoch < 0) {   // (monotone counters; wrap-safe comparison)
            __nanosleep(100);
            if (++spins > HALO_SPIN_LIMIT) { atomicExch(err_flag, 2); break; }
        }
        __threadfence_system();
    }
}

// host layout (planes of n points, cells back to back) <-> device layout (planes padded to np): one block per cell
__global__ void __launch_bounds__(256) repack_kernel(const CellInfo* __restrict__ cells, const long long* __restrict__ host_off,
                                                     int ncell, int comps, double* __restrict__ padded,
                                                     double* __restrict__ packed, int to_padded) {
    for (int c = blockIdx.x; c < ncell; c += gridDim.x) {
        const long long doff = cells[c].doff * comps, hoff = host_off[c] * comps;
        const int n = cells[c].n, np = cells[c].np;
        for (int t = threadIdx.x; t < comps * n; t += blockDim.x) {
            const int p = t / n, i = t - p * n;
            if (to_padded) padded[doff + (long long)p * np + i] = packed[hoff + t];
            else packed[hoff + t] = padded[doff + (long long)p * np + i];
        }
    }
}

// the same for a list of cells: packed block q holds cell list[q] (partition migration, selective download)
__global__ void __launch_bounds__(256) repack_list_kernel(const CellInfo* __restrict__ cells, const int* __restrict__ list,
                                                          const long long* __restrict__ list_off, int nlist, int comps,
                                                          double* __restrict__ padded, double* __restrict__ packed,
                                                          int to_padded) {
    for (int q = blockIdx.x; q < nlist; q += gridDim.x) {
        const CellInfo& ci = cells[list[q]];
        const long long doff = ci.doff * comps, hoff = list_off[q] * comps;
        const int n = ci.n, np = ci.np;
        for (int t = threadIdx.x; t < comps * n; t += blockDim.x) {
            const int p = t / n, i = t - p * n;
            if (to_padded) padded[doff + (long long)p * np + i] = packed[hoff + t];
            else packed[hoff + t] = padded[doff + (long long)p * np + i];
        }
    }
}

__global__ void exp_nonpos_kernel(const double* __restrict__ x, double* __restrict__ y, long long n) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        y[t] = exp_nonpos(x[t]);
}

__global__ void fill_kernel(double* p, long long n, double v) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        p[t] = v;
}

// ------------------------------------------------------------------------------------------------
// Physical-space adaptation sensor: update_criterion!(ka), Physical_space/AMR.jl:256-341, with the
// Löhner estimator of Physical_space/Criteria.jl:14-200.  Event-driven (once per adaptation pass), macroscopic fields
// only: one thread per physical cell.  Every operation is an explicitly rounded one (no FMA contraction), so equal
// inputs give the oracle's bits and a flag never flips on a borderline sensor value.
namespace sensor {
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dvd(double a, double b) { return __ddiv_rn(a, b); }
// Julia's max: NaN if either argument is NaN
__device__ __forceinline__ double jmax(double a, double b) {
    if (a != a || b != b) return CUDART_NAN;
    return a > b ? a : b;
}
constexpr double LOHNER_ABS_FLOOR = 1e-4;          // Criteria.jl:2
constexpr double PRIMITIVE_REL_JUMP_FLOOR = 1e-3;  // Criteria.jl:3
constexpr double VORTICITY_JUMP_FLOOR = 2e-2;      // Criteria.jl:4

template <int D>
__device__ __forceinline__ void prim_of(const double* w, double gamma, double* prim) {   // lib/KitCore get_prim
    prim[0] = w[0];
    double m2 = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) { prim[1 + d] = dvd(w[1 + d], w[0]); m2 = add(m2, mul(w[1 + d], w[1 + d])); }
    prim[D + 1] = dvd(dvd(mul(0.5, w[0]), sub(gamma, 1.0)), sub(w[D + 1], dvd(mul(0.5, m2), w[0])));
}
// lohner_value, Criteria.jl:25-31
__device__ __forceinline__ double lohner_value(double l, double c, double r, double dsL, double dsR, double eps) {
    const double scale = add(add(mul(dsR, fabs(l)), mul(add(dsL, dsR), fabs(c))), mul(dsL, fabs(r)));
    if (scale < mul(LOHNER_ABS_FLOOR, dsL < dsR ? dsL : dsR)) return 0.0;
    const double denom = add(add(mul(dsR, fabs(sub(l, c))), mul(dsL, fabs(sub(r, c)))), mul(eps, scale));
    if (denom <= 0.0) return 0.0;
    return dvd(fabs(add(sub(mul(dsR, l), mul(add(dsL, dsR), c)), mul(dsL, r))), denom);
}
__device__ __forceinline__ bool amplitude_ok(double l, double c, double r) {   // Criteria.jl:33-37
    const double jump = jmax(fabs(sub(l, c)), fabs(sub(r, c)));
    const double scale = jmax(fabs(c), LOHNER_ABS_FLOOR);
    return jump >= mul(PRIMITIVE_REL_JUMP_FLOOR, scale);
}
// sw is [dir][row]; velocity_slope + vorticity, Criteria.jl:39-53
template <int D>
__device__ __forceinline__ double vslope(const double* sw, const double* prim, int comp, int dir) {
    return dvd(sub(sw[dir * (D + 2) + comp], mul(prim[comp], sw[dir * (D + 2)])), prim[0]);
}
template <int D>
__device__ __forceinline__ double vorticity(const double* sw, const double* prim) {
    if (D == 2) return sub(vslope<D>(sw, prim, 1, 1), vslope<D>(sw, prim, 2, 0));
    const double c1 = sub(vslope<D>(sw, prim, 2, 2), vslope<D>(sw, prim, 3, 1));
    const double c2 = sub(vslope<D>(sw, prim, 3, 0), vslope<D>(sw, prim, 1, 2));
    const double c3 = sub(vslope<D>(sw, prim, 1, 1), vslope<D>(sw, prim, 2, 0));
    return __dsqrt_rn(add(add(mul(c1, c1), mul(c2, c2)), mul(c3, c3)));
}
template <int D>
__device__ __forceinline__ bool vorticity_ok(double l, double c, double r, const double* prim, double h) {   // :55-67
    const double omega = jmax(fabs(l), jmax(fabs(c), fabs(r)));
    double speed2 = 0.0;
#pragma unroll
    for (int i = 1; i <= D; ++i) speed2 = add(speed2, mul(prim[i], prim[i]));
    const double lambda = jmax(fabs(prim[D + 1]), EPS_MACH);
    const double vscale = jmax(__dsqrt_rn(speed2), dvd(1.0, __dsqrt_rn(lambda)));
    return mul(omega, h) >= mul(VORTICITY_JUMP_FLOOR, vscale);
}
// one side of update_Lohner_inner_ps!, Criteria.jl:124-160.  Ids >= n_real are SolidNeighbor pseudo-cells, whose
// w and sw are zero in the reference (Boundary/Immersed_boundary.jl:300-302).
template <int D>
__device__ __forceinline__ void side(const CellInfo* __restrict__ cells, const double* __restrict__ w,
                                     const double* __restrict__ sw, const int* __restrict__ ids, int cnt, int n_real,
                                     const double* mid, int dir, bool coarser, double gamma, double* prim_out,
                                     double* sw_out) {
    constexpr int M = D + 2;
    double ws[M];
#pragma unroll
    for (int j = 0; j < M; ++j) ws[j] = 0.0;
#pragma unroll
    for (int q = 0; q < M * D; ++q) sw_out[q] = 0.0;
    for (int k = 0; k < cnt; ++k) {
        const int id = ids[k];
        const bool real = id < n_real;
#pragma unroll
        for (int j = 0; j < M; ++j) ws[j] = add(ws[j], real ? w[(size_t)id * M + j] : 0.0);
#pragma unroll
        for (int q = 0; q < M * D; ++q) sw_out[q] = add(sw_out[q], real ? sw[(size_t)id * M * D + q] : 0.0);
    }
    if (coarser) {
        const int id = ids[0];
        const bool real = id < n_real;
        const CellInfo& nb = cells[id];
#pragma unroll
        for (int j = 0; j < M; ++j) {
            double acc = 0.0;
            bool first = true;
#pragma unroll
            for (int t = 0; t < D; ++t) {
                if (t == dir) continue;
                const double term = mul(sub(mid[t], nb.mid[t]), real ? sw[(size_t)id * M * D + t * M + j] : 0.0);
                acc = first ? term : add(acc, term);
                first = false;
            }
            ws[j] = add(ws[j], acc);
        }
    }
    const double inv_n = (double)cnt;
#pragma unroll
    for (int j = 0; j < M; ++j) ws[j] = dvd(ws[j], inv_n);
#pragma unroll
    for (int q = 0; q < M * D; ++q) sw_out[q] = dvd(sw_out[q], inv_n);
    prim_of<D>(ws, gamma, prim_out);
}
template <int D>
__device__ __forceinline__ double sensor_of(const double* loh) {   // ps_sensor, Criteria.jl:14-23
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) { a = jmax(a, loh[d * (D + 2)]); b = jmax(b, loh[d * (D + 2) + D + 1]); }
    return jmax(a, b);
}
}  // namespace sensor

// lohner[c][dir][row] of every local cell and above[c] = (ps_sensor > threshold) as a double (1.0 / 0.0) so that the
// mirrors' decisions travel like any other per-cell field (lohner_flag_exchange!, Parallel/Ghost.jl:939-978)
template <int D>
__global__ void __launch_bounds__(128) ps_lohner_kernel(const CellInfo* __restrict__ cells,
                                                        const int* __restrict__ nb_state, const int* __restrict__ nb_off,
                                                        const int* __restrict__ nb_ids, const double* __restrict__ w,
                                                        const double* __restrict__ prim_all,
                                                        const double* __restrict__ sw_all, int n_local, int n_real,
                                                        double gamma, double threshold, double* __restrict__ lohner,
                                                        double* __restrict__ above) {
    constexpr int M = D + 2;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_local) return;
    double loh[M * D];
#pragma unroll
    for (int q = 0; q < M * D; ++q) loh[q] = 0.0;
    const CellInfo& ci = cells[c];
    bool fluid = ci.bound_enc >= 0;   // AMR.jl:264-265
    if (fluid) {
        double prim[M], sw[M * D], mid[D];
#pragma unroll
        for (int j = 0; j < M; ++j) prim[j] = prim_all[(size_t)c * M + j];
#pragma unroll
        for (int q = 0; q < M * D; ++q) sw[q] = sw_all[(size_t)c * M * D + q];
#pragma unroll
        for (int t = 0; t < D; ++t) mid[t] = ci.mid[t];
        const double omega = sensor::vorticity<D>(sw, prim);
#pragma unroll
        for (int dir = 0; dir < D; ++dir) {
            const int e = c * 2 * D + 2 * dir;
            const int sL = nb_state[e], sR = nb_state[e + 1];
            if (sL == 0 || sR == 0) continue;   // AMR.jl:160-255
            const double ds = ci.ds[dir];
            // AMR.jl:5-157: same level ds, finer 0.75 ds, coarser 1.5 ds
            const double dsL = sL == 1 ? ds : (sL > 1 ? sensor::mul(0.75, ds) : sensor::mul(1.5, ds));
            const double dsR = sR == 1 ? ds : (sR > 1 ? sensor::mul(0.75, ds) : sensor::mul(1.5, ds));
            double pL[M], pR[M], swL[M * D], swR[M * D];
            sensor::side<D>(cells, w, sw_all, nb_ids + nb_off[e], nb_off[e + 1] - nb_off[e], n_real, mid, dir, dsL > ds,
                            gamma, pL, swL);
            sensor::side<D>(cells, w, sw_all, nb_ids + nb_off[e + 1], nb_off[e + 2] - nb_off[e + 1], n_real, mid, dir,
                            dsR > ds, gamma, pR, swR);
            const double oL = sensor::vorticity<D>(swL, pL), oR = sensor::vorticity<D>(swR, pR);
            const bool use_vort = sensor::vorticity_ok<D>(oL, omega, oR, prim, dsL > dsR ? dsL : dsR);
            const double eps_l = sensor::mul(0.2, ds);
#pragma unroll
            for (int j = 0; j < M; ++j) {
                if (j == 1)
                    loh[dir * M + j] = use_vort ? sensor::lohner_value(oL, omega, oR, dsL, dsR, eps_l) : 0.0;
                else
                    loh[dir * M + j] = sensor::amplitude_ok(pL[j], prim[j], pR[j])
                                           ? sensor::lohner_value(pL[j], prim[j], pR[j], dsL, dsR, eps_l) : 0.0;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < M * D; ++q) lohner[(size_t)c * M * D + q] = loh[q];
    above[c] = (fluid && sensor::sensor_of<D>(loh) > threshold) ? 1.0 : 0.0;
}

// apply_amr_buffer!, AMR.jl:296-341: an unflagged fluid cell with a flagged fluid face neighbour (local or ghost) gets
// lohner .= 2 threshold.  Decisions are read from `above` (un-inflated), so the buffer stays one cell thick.
template <int D>
__global__ void __launch_bounds__(128) ps_buffer_kernel(const CellInfo* __restrict__ cells,
                                                        const int* __restrict__ nb_state, const int* __restrict__ nb_off,
                                                        const int* __restrict__ nb_ids, int n_local, int n_real,
                                                        double threshold, const double* __restrict__ above,
                                                        double* __restrict__ lohner, double* __restrict__ sensor_out) {
    constexpr int M = D + 2;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_local) return;
    bool flagged = false;
    if (cells[c].bound_enc >= 0 && above[c] == 0.0) {
        for (int f = 0; f < 2 * D && !flagged; ++f) {
            const int e = c * 2 * D + f;
            if (nb_state[e] == 0) continue;
            for (int k = nb_off[e]; k < nb_off[e + 1] && !flagged; ++k) {
                const int id = nb_ids[k];
                if (id >= n_real || cells[id].bound_enc < 0) continue;
                flagged = above[id] != 0.0;
            }
        }
    }
    double loh[M * D];
#pragma unroll
    for (int q = 0; q < M * D; ++q) {
        loh[q] = flagged ? sensor::mul(2.0, threshold) : lohner[(size_t)c * M * D + q];
        if (flagged) lohner[(size_t)c * M * D + q] = loh[q];
    }
    sensor_out[c] = sensor::sensor_of<D>(loh);
}

// rows of `width` doubles of a per-cell array: gather the rows of a list of cells into a contiguous buffer
__global__ void __launch_bounds__(256) gather_rows_kernel(const int* __restrict__ list, int n, int width,
                                                          const double* __restrict__ src, double* __restrict__ dst) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < (long long)n * width;
         t += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(t / width), j = (int)(t - (long long)q * width);
        dst[t] = src[(size_t)list[q] * width + j];
    }
}

// ------------------------------------------------------------------------------------------------
// Velocity-space adaptation inputs: vs_refine! / vs_coarsen! decisions per velocity point,
// Velocity_space/AMR.jl:26-115 with Velocity_space/Criteria.jl; the face-neighbour search of
// Velocity_space/Neighbor.jl as a painted finest-level lattice instead of sorted Morton keys.
struct VsPar {
    int mode, maxlevel;
    int gmax[MAXD];          // finest-level lattice extent: vs_trees_num << maxlevel
    double vmin[MAXD], h_fine[MAXD];
    double coeff_lohner, coeff_local, coeff_global, vr_density, vr_energy;
    double cell_weight;      // volume of a finest-level velocity cell (vs_resolution, AMR.jl:163)
};
struct VsGridTask {
    long long goff;          // velocity-grid statics of the grid
    long long nb_off;        // its face-neighbour table [n][DIM][2] in ints
    int n, np;
};

// One CTA per distinct velocity grid: every leaf paints its id over the finest-level lattice cells it covers
// (_corner_index, Neighbor.jl:63-66), then every leaf probes half a finest cell across each face on its centre line
// (vs_face_neighbor, Neighbor.jl:181-204) and reads the owner.  -1: velocity-domain boundary (vacuum).
template <int D>
__global__ void __launch_bounds__(256) vs_neighbors_kernel(const VsGridTask* __restrict__ tasks, VsPar par,
                                                           const double* __restrict__ v_mid,
                                                           const int8_t* __restrict__ v_level,
                                                           int* __restrict__ lattice_all, long long lattice_size,
                                                           int* __restrict__ nbt, int* __restrict__ err_flag) {
    const VsGridTask t = tasks[blockIdx.x];
    int* lat = lattice_all + (long long)blockIdx.x * lattice_size;
    for (long long q = threadIdx.x; q < lattice_size; q += blockDim.x) lat[q] = -1;
    __syncthreads();
    const double* v = v_mid + t.goff * D;
    const int8_t* lev = v_level + t.goff;
    for (int i = threadIdx.x; i < t.n; i += blockDim.x) {
        const int span = 1 << (par.maxlevel - lev[i]);
        int g[D];
        bool ok = true;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const double cell = __dmul_rn(par.h_fine[d], (double)span);
            const double x = __ddiv_rn(__dsub_rn(__dsub_rn(v[(size_t)d * t.np + i], __dmul_rn(0.5, cell)), par.vmin[d]),
                                       par.h_fine[d]);
            g[d] = (int)llrint(x);
            ok = ok && g[d] >= 0 && g[d] + span <= par.gmax[d];
        }
        if (!ok) { *err_flag = 3; continue; }
        const int cnt = D == 2 ? span * span : span * span * span;
        for (int o = 0; o < cnt; ++o) {
            long long lin = 0;
            int rem = o;
#pragma unroll
            for (int d = D - 1; d >= 0; --d) {
                const int od = d == 0 ? rem : rem / (d == 1 ? span : span * span);
                if (d > 0) rem -= od * (d == 1 ? span : span * span);
                lin = lin * par.gmax[d] + (g[d] + od);
            }
            lat[lin] = i;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < t.n; i += blockDim.x) {
        const int span = 1 << (par.maxlevel - lev[i]);
#pragma unroll
        for (int dim = 0; dim < D; ++dim) {
            const double cell = __dmul_rn(par.h_fine[dim], (double)span);
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const double dir = s ? 1.0 : -1.0;
                long long lin = 0;
                bool inside = true;
#pragma unroll
                for (int d = D - 1; d >= 0; --d) {
                    double coord = v[(size_t)d * t.np + i];
                    if (d == dim)
                        coord = __dadd_rn(coord, __dmul_rn(dir, __dadd_rn(__dmul_rn(0.5, cell), __dmul_rn(0.5, par.h_fine[d]))));
                    const double q = floor(__ddiv_rn(__dsub_rn(coord, par.vmin[d]), par.h_fine[d]));
                    inside = inside && q >= 0.0 && q < (double)par.gmax[d];
                    lin = lin * par.gmax[d] + (long long)(inside ? q : 0.0);
                }
                nbt[t.nb_off + ((long long)i * D + dim) * 2 + s] = inside ? lat[lin] : -1;
            }
        }
    }
}

namespace vsadapt {
using sensor::add; using sensor::sub; using sensor::mul; using sensor::dvd; using sensor::jmax;
constexpr double EPS_FLT = 1e-2, EPS_ABS = 1e-3, COARSEN_RATIO = 0.3;   // Criteria.jl:213-217
// _lohner_ratio, Criteria.jl:222-228
__device__ __forceinline__ double ratio(double L, double C, double R, double dsL, double dsR, double scale) {
    const double both = add(dsL, dsR);
    const double num = fabs(add(sub(mul(dsR, L), mul(both, C)), mul(dsL, R)));
    const double den = add(add(add(mul(dsR, fabs(sub(L, C))), mul(dsL, fabs(sub(R, C)))),
                               mul(EPS_FLT, add(add(mul(dsR, fabs(L)), mul(both, fabs(C))), mul(dsL, fabs(R))))),
                           mul(mul(EPS_ABS, scale), both));
    return den > 0 ? dvd(num, den) : 0.0;
}
}  // namespace vsadapt

// block-wide max of two values (result broadcast)
__device__ __forceinline__ void block_max2(double& a, double& b, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    for (int o = 16; o > 0; o >>= 1) {
        a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o));
        b = fmax(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    __syncthreads();
    if (lane == 0) { red[warp] = a; red[32 + warp] = b; }
    __syncthreads();
    a = red[0]; b = red[32];
    for (int q = 1; q < nwarp; ++q) { a = fmax(a, red[q]); b = fmax(b, red[32 + q]); }
}

// One CTA per local cell.  refine_flag / coarsen_ok in the host's point order (host_off[c] + i).
template <int D, int K>
__global__ void __launch_bounds__(256) vs_criterion_kernel(DevView g, VsPar par, const long long* __restrict__ host_off,
                                                           const long long* __restrict__ nb_off_of_grid,
                                                           const int* __restrict__ nbt,
                                                           unsigned char* __restrict__ refine_flag,
                                                           unsigned char* __restrict__ coarsen_ok) {
    using namespace vsadapt;
    constexpr int M = D + 2;
    __shared__ double red[64];
    const int c = blockIdx.x;
    const CellInfo& ci = g.cells[c];
    const CellPtr<D, K> own(g, ci);
    const int n = ci.n, np = ci.np;
    const bool lohner = par.mode == 0;
    double s1 = 0.0, s2 = 0.0;
    if (lohner) {   // vs_lohner_scales, Criteria.jl:238-248 (values are never NaN: fmax is max)
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            s1 = fmax(s1, fabs(own.f[i]));
            if (K == 2) s2 = fmax(s2, fabs(own.f[np + i]));
        }
        block_max2(s1, s2, red);
    }
    double w[M], U[D], ds[D];
#pragma unroll
    for (int j = 0; j < M; ++j) w[j] = g.w[(size_t)c * M + j];
    double U2 = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) { U[d] = g.prim[(size_t)c * M + 1 + d]; ds[d] = ci.ds[d]; U2 = add(U2, mul(U[d], U[d])); }
    const double eden = sub(w[M - 1], mul(mul(0.5, w[0]), U2));   // w[end] - 0.5 w[1] sum(U.^2)
    const double two_d = (double)(1 << D);
    const int* nb_tab = lohner ? nbt + nb_off_of_grid[ci.grid] : nullptr;
    const long long out0 = host_off[c];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double f[K], cdf[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {   // _criterion_cell!, AMR.jl:8-21
            f[k] = own.f[k * np + i];
            double mx = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const double a = fabs(mul(own.s[(d * K + k) * np + i], ds[d]));
                if (a > mx) mx = a;
            }
            cdf[k] = add(f[k], mx);
        }
        double S = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) { const double t = sub(U[d], own.v[d * np + i]); S = add(S, mul(t, t)); }
        const double wgt = own.wt[i];
        const int level = own.lev[i];
        // local_contribution_refine_flag / _coarsen_flag, Criteria.jl:19-49 (note the different association for NDF = 1)
        const double e2 = K == 2 ? mul(mul(0.5, add(mul(S, cdf[0]), cdf[K - 1])), wgt) : 0.0;
        const double e_ref = K == 2 ? fabs(e2) : fabs(mul(mul(mul(0.5, S), cdf[0]), wgt));
        const double e_co = K == 2 ? e2 : mul(mul(0.5, mul(S, cdf[0])), wgt);
        const double mass = dvd(mul(cdf[0], wgt), w[0]);
        const bool local_refine = jmax(dvd(e_ref, eden), mass) > par.coeff_local;
        const bool local_coarsen = jmax(dvd(e_co, eden), mass) < dvd(par.coeff_local, two_d);
        bool base_refine, ok;
        if (lohner) {   // vs_lohner_indicator, Criteria.jl:259-287
            double eta = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const double hi = mul(par.h_fine[d], (double)(1 << (par.maxlevel - level)));
                const int Ln = nb_tab[((long long)i * D + d) * 2], Rn = nb_tab[((long long)i * D + d) * 2 + 1];
                const double dsL = Ln < 0 ? hi : mul(0.5, add(hi, mul(par.h_fine[d], (double)(1 << (par.maxlevel - own.lev[Ln])))));
                const double dsR = Rn < 0 ? hi : mul(0.5, add(hi, mul(par.h_fine[d], (double)(1 << (par.maxlevel - own.lev[Rn])))));
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const double fL = Ln < 0 ? 0.0 : own.f[k * np + Ln], fR = Rn < 0 ? 0.0 : own.f[k * np + Rn];
                    eta = jmax(eta, ratio(fL, f[k], fR, dsL, dsR, k == 0 ? s1 : s2));
                }
            }
            base_refine = eta > par.coeff_lohner || local_refine;
            ok = eta < mul(COARSEN_RATIO, par.coeff_lohner) && local_coarsen;
        } else {   // global_contribution_{refine,coarsen}_flag, Criteria.jl:55-84
            const double e_gl = K == 2 ? e2 : e_co;
            const double m0 = mul(cdf[0], wgt);
            const bool global_refine = m0 > mul(par.coeff_global, par.vr_density) || e_gl > mul(par.vr_energy, par.coeff_global);
            const bool global_coarsen = m0 < dvd(mul(par.coeff_global, par.vr_density), two_d) &&
                                        e_gl < dvd(mul(par.vr_energy, par.coeff_global), two_d);
            base_refine = local_refine || global_refine;
            ok = local_coarsen && global_coarsen;
        }
        if (refine_flag) refine_flag[out0 + i] = (unsigned char)(level < par.maxlevel && base_refine);
        if (coarsen_ok) coarsen_ok[out0 + i] = (unsigned char)ok;
    }
}

// vs_resolution(ps_data, kinfo), AMR.jl:154-166: per local cell (density_max, energy_max) * weight; solid cells give 0
template <int D, int K>
__global__ void __launch_bounds__(256) vs_resolution_kernel(DevView g, VsPar par, double* __restrict__ out) {
    __shared__ double red[64];
    const int c = blockIdx.x;
    const CellInfo& ci = g.cells[c];
    const CellPtr<D, K> own(g, ci);
    const int n = ci.n, np = ci.np;
    double U[D];
#pragma unroll
    for (int d = 0; d < D; ++d) U[d] = g.prim[(size_t)c * (D + 2) + 1 + d];
    double dmax = -CUDART_INF, emax = -CUDART_INF;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double c2 = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) { const double t = sensor::sub(U[d], own.v[d * np + i]); c2 = sensor::add(c2, sensor::mul(t, t)); }
        const double f0 = own.f[i];
        dmax = fmax(dmax, f0);
        double e = sensor::mul(f0, c2);
        if (K == 2) { const double f1 = own.f[np + i]; dmax = fmax(dmax, f1); e = sensor::add(e, f1); }
        emax = fmax(emax, e);
    }
    block_max2(dmax, emax, red);
    if (threadIdx.x == 0) {
        const bool fluid = ci.bound_enc >= 0;
        out[2 * c] = fluid ? sensor::mul(dmax, par.cell_weight) : 0.0;
        out[2 * c + 1] = fluid ? sensor::mul(sensor::mul(0.5, emax), par.cell_weight) : 0.0;
    }
}

// conserved_I_porjection!(vs_data, ps_data.w), Theory/I-projection.jl:144-159, on a list of cells (the cells a
// velocity-space adaptation pass regridded, Velocity_space/AMR.jl:120-133): f_h is projected onto the cell's w
// (2D2F: with the internal energy of b removed).  One CTA per cell; the same Newton iteration as cip_update_kernel.
template <int D, int K>
__global__ void __launch_bounds__(256) project_cells_kernel(DevView g, const int* __restrict__ cell_list) {
    constexpr int M = D + 2, NJ = M * (M + 1) / 2, NV = M + NJ;
    __shared__ double red[NV * 9];
    __shared__ CellInfo ci;
    const int c = cell_list[blockIdx.x];
    copy_words(g.cells + c, &ci, (int)(sizeof(CellInfo) / sizeof(int)));
    __syncthreads();
    const CellPtr<D, K> own(g, ci);
    const int n = ci.n, np = ci.np;
    double* f = g.df + ci.doff * K;
    double W[M];
#pragma unroll
    for (int m = 0; m < M; ++m) W[m] = g.w[(size_t)c * M + m];
    double e[1] = {0.0}, mn[2] = {CUDART_INF, CUDART_INF};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double x = f[i];
        if (K > 1) e[0] += own.wt[i] * f[np + i];
        mn[0] = fmin(mn[0], x);
        if (x > 0.) mn[1] = fmin(mn[1], x);
    }
    if (K > 1) {
        block_reduce<1>(e, red);
        W[M - 1] -= e[0] / 2;
    }
    block_min<2>(mn, red);
    double lam[M];
    solve_projection<D, K>(own, f, n, np, W, mn, red, lam);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double v[D], psi[M];
#pragma unroll
        for (int t = 0; t < D; ++t) v[t] = own.v[t * np + i];
        psi_of<D>(v, psi);
        double lp = 0.0;
#pragma unroll
        for (int m = 0; m < M; ++m) lp += lam[m] * psi[m];
        f[i] *= exp(lp);
    }
}

// inverse of gather_rows_kernel: row q of the contiguous buffer goes to row list[q] of the per-cell array
__global__ void __launch_bounds__(256) scatter_rows_kernel(const int* __restrict__ list, int n, int width,
                                                           const double* __restrict__ src, double* __restrict__ dst) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < (long long)n * width;
         t += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(t / width), j = (int)(t - (long long)q * width);
        dst[(size_t)list[q] * width + j] = src[t];
    }
}

}  // namespace kamr
