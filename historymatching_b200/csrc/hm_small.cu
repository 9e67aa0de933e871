// Whole simulator of a SMALL grid in one kernel (sm_100a): one CTA per ensemble member, every
// time step (transmissibilities, multigrid-preconditioned CG pressure solve, fluxes, CFL count,
// all upwind sub-steps, observation gather) without leaving the SM.
//
// The notebooks run 20 x 20 grids (HistoryMatch.py:97, Optimise.py:64) with 30-200 members and
// re-run the ensemble dozens of times (IES, ES-MDA, EnOpt line searches).  On the streamed path
// that workload is pure launch latency (~390 launches per time step, 15 000 per forward run);
// here the state of a member - the multigrid hierarchy (operators + vectors of every level),
// pressure, CG vectors, saturation - lives in shared memory (<= 96 B per cell, grids of up to
// 2048 cells) and the forward run of the whole ensemble is ONE launch.
//
// Same algorithm as the streamed path (hm_sim.cu / hm_pressure.cu):
// identical transmissibility / flux / CFL expressions, the same V-cycle (hm_mg_onchip.cuh,
// here starting at level 0) inside the same PCG recurrence, the same convergence test
// ||r|| <= rtol ||q||.  Only the summation order of the dot products differs.
#include <type_traits>

#include "hm_mg_onchip.cuh"

using namespace hmsim;

namespace hmsim {
int sim_small_supported(const hm_sim_desc& d);
int sim_small(hm_ctx* ctx, const hm_sim_desc& d, int m0, int nm);
}  // namespace hmsim

namespace {

constexpr int kSmallThreads = 256;
constexpr int kSmallMaxPer = 8;  // cells per thread: the kernel is instantiated for 2, 4 and 8
constexpr int kSmallMaxCells = kSmallThreads * kSmallMaxPer;
constexpr int kSmallWarps = kSmallThreads / 32;
constexpr int kDensePitch = 33;  // row pitch of the dense coarse inverse (<= 32 x 32)

struct SmallArgs {
    // Hierarchy from the fine grid (level 0) down to the first level of <= 32 cells.  off[l] is where the
    // data of level l starts inside every all-level array; each level is surrounded by ny+1 zero entries, so
    // the 5-point stencil needs no boundary tests (boundary faces carry T = 0).  The pointer members are unused.
    OnchipMeta mt;
    Geo g;
    Fluid fl;
    Wells w;
    const double* K;
    int64_t K_ms, K_cs;
    const double* por;
    const double* S0;
    int64_t S0_ms;
    double dt;
    int n_steps, n_obs;
    const int32_t* obs_cell;
    double* S_last;
    double* S_hist;
    int hist_stride;
    double* obs;
    double* P_last;
    int32_t* status;
    int32_t* substeps;
    int32_t* cg_iters;
    int32_t* totals;  // [member][2]: CG iterations, sub-steps (summed over the time steps)
    double tol2;
    int max_iter;
    int unit_fluid;
    int lw;  // index of the coarsest level (solved exactly)
};

// Block-wide reductions with ONE barrier, result in every thread: each warp leaves its partial in slot
// `flip` of a double-buffered array, every thread sums the partials after the barrier.  A slot is
// rewritten two calls later, i.e. after at least one more barrier, so no trailing barrier is needed.
__device__ __forceinline__ double block_sum_all(double v, double (*red)[kSmallWarps], int& flip) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[flip][threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kSmallWarps; ++w) t += red[flip][w];
    flip ^= 1;
    return t;
}
__device__ __forceinline__ double block_min_all(double v, double (*red)[kSmallWarps], int& flip) {
    v = warp_min(v);
    if ((threadIdx.x & 31) == 0) red[flip][threadIdx.x >> 5] = v;
    __syncthreads();
    double t = INFINITY;
#pragma unroll
    for (int w = 0; w < kSmallWarps; ++w) t = fmin(t, red[flip][w]);
    flip ^= 1;
    return t;
}

// (A x) at the cell whose data sit at index o of the padded arrays; no boundary tests (zero pads, T = 0 on
// boundary faces).  `first` = this is cell 0 of the level (the pinned cell).
__device__ __forceinline__ double small_Ax(const double* __restrict__ x, const double* __restrict__ TX,
                                           const double* __restrict__ TY, int o, int ny, bool first, double pin) {
    const double xc = x[o];
    double y = TX[o] * (xc - x[o - ny]);
    y = fma(TX[o + ny], xc - x[o + ny], y);
    y = fma(TY[o], xc - x[o - 1], y);
    y = fma(TY[o + 1], xc - x[o + 1], y);
    if (first) y = fma(pin, xc, y);
    return y;
}

// Dense inverse of the coarsest-level operator (n <= 32 cells, SPD thanks to the pin) by in-place
// Gauss-Jordan elimination without pivoting; the CTA works on the n x n array, two barriers per pivot.
template <int NT>
__device__ __forceinline__ void small_coarse_inverse(const OnchipMeta& mt, const OnchipSmem<double>& s, double* Ainv,
                                                     int lw, double pinv) {
    const int n = mt.M[lw], ny = mt.ny[lw], o = mt.off[lw];
    int rr[4], cc[4];  // n*n <= 1024 entries: at most 4 per thread
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int t = threadIdx.x + q * NT;
        rr[q] = t < n * n ? t / n : -1;
        cc[q] = t - rr[q] * n;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int r = rr[q], c = cc[q];
        if (r < 0) continue;
        const double txl = s.TX[o + r], txh = s.TX[o + r + ny], tyl = s.TY[o + r], tyh = s.TY[o + r + 1];
        double v = 0.0;  // zero transmissibility on boundary faces: the tests below only pick the neighbour
        if (c == r) v = tyl + tyh + txl + txh + (r == 0 ? pinv : 0.0);
        if (c == r - ny) v = -txl;
        if (c == r + ny) v = -txh;
        if (c == r - 1 && tyl != 0.0) v = -tyl;
        if (c == r + 1 && tyh != 0.0) v = -tyh;
        Ainv[r * kDensePitch + c] = v;
    }
    __syncthreads();
    for (int k = 0; k < n; ++k) {
        const double pk = 1.0 / Ainv[k * kDensePitch + k];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int r = rr[q], c = cc[q];
            if (r >= 0 && r != k && c != k)
                Ainv[r * kDensePitch + c] =
                    fma(-Ainv[r * kDensePitch + k] * pk, Ainv[k * kDensePitch + c], Ainv[r * kDensePitch + c]);
        }
        __syncthreads();
        if ((int)threadIdx.x < n && (int)threadIdx.x != k) {
            Ainv[k * kDensePitch + threadIdx.x] *= pk;
            Ainv[threadIdx.x * kDensePitch + k] *= -pk;
        }
        if (threadIdx.x == 0) Ainv[k * kDensePitch + k] = pk;
        __syncthreads();
    }
}

__device__ __forceinline__ void small_coarse_apply(const OnchipMeta& mt, const OnchipSmem<double>& s,
                                                   const double* Ainv, int lw) {
    const int n = mt.M[lw], co = mt.off[lw];
    if ((int)threadIdx.x < n) {
        const double* row = Ainv + threadIdx.x * kDensePitch;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int j = 0;
        for (; j + 3 < n; j += 4) {
            a0 = fma(row[j], s.B[co + j], a0);
            a1 = fma(row[j + 1], s.B[co + j + 1], a1);
            a2 = fma(row[j + 2], s.B[co + j + 2], a2);
            a3 = fma(row[j + 3], s.B[co + j + 3], a3);
        }
        for (; j < n; ++j) a0 = fma(row[j], s.B[co + j], a0);
        s.X[co + threadIdx.x] = (a0 + a1) + (a2 + a3);
    }
}

// V-cycle of the small-grid kernel: Z = M^-1 B[level 0].  Same smoother and transfer as onchip_cycle
// (Chebyshev-weighted Jacobi, 2x2 aggregation), organised for few CTA barriers:
//   * the iterate of a level ping-pongs between two arrays (XA = s.X, XB): one barrier per sweep instead of
//     two, and the first pre-sweep (zero initial guess) needs no stencil;
//   * the coarsest level lw (<= 32 cells) is solved EXACTLY: its dense inverse Ainv is built once per time
//     step (small_coarse_inverse) and applied as one matrix-vector product.  (The streamed path recurses
//     down to 1 x 1 with smoothing only; an exact coarse solve is the better preconditioner and replaces
//     ~40 latency-bound one-warp sweeps per cycle.)
// Every level above lw does kNu pre- and kNu post-sweeps, so its result ends in XB; returns where level 0's is.
template <int NT, int PER>
__device__ __forceinline__ const double* small_vcycle(const OnchipMeta& mt, const OnchipSmem<double>& s, double* XB,
                                                      double pinv, int lw, const double* Ainv) {
    if (lw == 0) {  // the whole grid is the coarsest level: direct solve
        small_coarse_apply(mt, s, Ainv, 0);
        __syncthreads();
        return s.X;
    }
    double* buf[2] = {s.X, XB};
    constexpr int cPre = (kNu - 1) & 1;  // buffer that holds a level's iterate after pre-smoothing
    for (int l = 0; l < lw; ++l) {
        const int M = mt.M[l], ny = mt.ny[l], o = mt.off[l];
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int e = threadIdx.x + k * NT;
            if (e < M) buf[0][o + e] = cheb_w(0) * s.DV[o + e] * s.B[o + e];
        }
        __syncthreads();
        int c = 0;
        for (int sw = 1; sw < kNu; ++sw, c ^= 1) {
            const double wgt = cheb_w(sw);
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                const int e = threadIdx.x + k * NT;
                if (e < M)
                    buf[c ^ 1][o + e] = buf[c][o + e] + wgt * s.DV[o + e] *
                                        (s.B[o + e] - small_Ax(buf[c], s.TX, s.TY, o + e, ny, e == 0, pinv));
            }
            __syncthreads();
        }
        // residual restricted to level l+1 (the right-hand side there)
        const int cM = mt.M[l + 1], cny = mt.ny[l + 1], co = mt.off[l + 1], fnx = mt.nx[l];
        const float cinv = mt.inv_ny[l + 1];
        for (int e = threadIdx.x; e < cM; e += NT) {
            int ci, cj;
            cell_ij(e, cny, cinv, ci, cj);
            double r = 0.0;
#pragma unroll
            for (int di = 0; di < 2; ++di)
#pragma unroll
                for (int dj = 0; dj < 2; ++dj) {
                    const int i = 2 * ci + di, j = 2 * cj + dj;
                    if (i < fnx && j < ny) {
                        const int fe = i * ny + j;
                        r += s.B[o + fe] - small_Ax(buf[cPre], s.TX, s.TY, o + fe, ny, fe == 0, pinv);
                    }
                }
            s.B[co + e] = r;
        }
        __syncthreads();
    }
    small_coarse_apply(mt, s, Ainv, lw);
    __syncthreads();
    for (int l = lw - 1; l >= 0; --l) {
        const int M = mt.M[l], ny = mt.ny[l], o = mt.off[l], cny = mt.ny[l + 1], co = mt.off[l + 1];
        const float inv = mt.inv_ny[l];
        const double* xc = (l + 1 == lw ? s.X : XB) + co;  // result of the coarser level
        for (int e = threadIdx.x; e < M; e += NT) {
            int i, j;
            cell_ij(e, ny, inv, i, j);
            buf[cPre][o + e] += xc[(i >> 1) * cny + (j >> 1)];
        }
        __syncthreads();
        int c = cPre;
        for (int sw = 0; sw < kNu; ++sw, c ^= 1) {
            const double wgt = cheb_w(kNu - 1 - sw);
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                const int e = threadIdx.x + k * NT;
                if (e < M)
                    buf[c ^ 1][o + e] = buf[c][o + e] + wgt * s.DV[o + e] *
                                        (s.B[o + e] - small_Ax(buf[c], s.TX, s.TY, o + e, ny, e == 0, pinv));
            }
            __syncthreads();
        }
    }
    return XB;  // (kNu - 1) + kNu buffer flips: always odd
}

template <int kSmallPer>
__global__ void __launch_bounds__(kSmallThreads, kSmallPer <= 4 ? 2 : 1)
k_sim_small(const __grid_constant__ SmallArgs a) {
    extern __shared__ __align__(16) double smx[];
    __shared__ int wc[kMaxWells];
    __shared__ double wr[kMaxWells];
    __shared__ double red[2][kSmallWarps];
    __shared__ double Ainv[32 * kDensePitch];
    __shared__ double bc;
    int flip = 0;
    constexpr int NT = kSmallThreads;
    const OnchipMeta& mt = a.mt;
    const Geo& g = a.g;
    const int m = blockIdx.x, tid = threadIdx.x;
    const int M = g.M, nx = g.Nx, ny = g.Ny, tot = mt.total, lw = a.lw;
    const int o0 = mt.off[0];  // = ny + 1: level-0 data start inside a padded array
    const float inv_ny = mt.inv_ny[0];
    // all-level arrays (tot entries, pads included): XA, B, TX, TY, DV, XB; level-0 vectors padded like level 0
    // (M + 2 (ny+1) entries): P, Pd; plain M-entry vectors: AP, S
    OnchipSmem<double> s;
    s.X = smx;
    s.B = smx + tot;
    s.TX = smx + 2 * tot;
    s.TY = smx + 3 * tot;
    s.DV = smx + 4 * tot;
    double* XB = smx + 5 * tot;
    const int lenP = M + 2 * o0;
    double* P = smx + 6 * tot + o0;  // pressure (CG iterate); P[-o0 .. M+o0) is addressable
    double* Pd = P + lenP;           // CG search direction
    double* AP = Pd + M + o0;        // A * search direction
    double* S = AP + M;
    double* TX0 = s.TX + o0;         // level-0 views
    double* TY0 = s.TY + o0;
    double* DV0 = s.DV + o0;
    double* R = s.B + o0;            // CG residual = right-hand side of level 0
    for (int e = tid; e < 6 * tot + 2 * lenP + 2 * M; e += NT) smx[e] = 0.0;  // pads stay zero for good
    __syncthreads();

    const double* Kx = a.K + (int64_t)m * a.K_ms;
    const double* Ky = Kx + a.K_cs;
    const double pinv = perm_value(g, Kx[0]) + perm_value(g, Ky[0]);  // pin of the singular Neumann problem: A[0,0] += Kx[0] + Ky[0]
    for (int e = tid; e < M; e += NT) S[e] = a.S0[(int64_t)m * a.S0_ms + e];
    if (a.S_hist)
        for (int e = tid; e < M; e += NT)
            a.S_hist[(int64_t)m * hist_rows(a.n_steps, a.hist_stride) * M + e] = a.S0[(int64_t)m * a.S0_ms + e];
    __syncthreads();

    int cg_fail = 0, tot_iters = 0, tot_sub = 0;
    for (int step = 0; step < a.n_steps; ++step) {
        load_wells(a.w, m, step, wc, wr);
        // ---- transmissibilities (Appendix A.2); 1/mobility staged in Pd / AP ----------------------------
        for (int e = tid; e < M; e += NT) {
            const double mob = total_mobility(S[e], g);
            Pd[e] = 1.0 / (mob * perm_value(g, Kx[e]));
            AP[e] = 1.0 / (mob * perm_value(g, Ky[e]));
        }
        __syncthreads();
        double txl[kSmallPer], tyl[kSmallPer], dv[kSmallPer];
#pragma unroll
        for (int k = 0; k < kSmallPer; ++k) {
            const int e = tid + k * NT;
            if (e < M) {
                int i, j;
                cell_ij(e, ny, inv_ny, i, j);
                txl[k] = i > 0 ? g.cx / (Pd[e - ny] + Pd[e]) : 0.0;
                const double txh = i < nx - 1 ? g.cx / (Pd[e] + Pd[e + ny]) : 0.0;
                tyl[k] = j > 0 ? g.cy / (AP[e - 1] + AP[e]) : 0.0;
                const double tyh = j < ny - 1 ? g.cy / (AP[e] + AP[e + 1]) : 0.0;
                double d = tyl[k] + tyh + txl[k] + txh;
                if (e == 0) d += pinv;
                dv[k] = 1.0 / d;
            }
        }
        __syncthreads();  // Pd was the staging array: zero it again before it serves as a padded CG vector
#pragma unroll
        for (int k = 0; k < kSmallPer; ++k) {
            const int e = tid + k * NT;
            if (e < M) {
                TX0[e] = txl[k];
                TY0[e] = tyl[k];
                DV0[e] = dv[k];
                Pd[e] = 0.0;
            }
        }
        __syncthreads();
        for (int l = 0; l < lw; ++l) onchip_coarsen<double, NT>(mt, s, l, pinv);
        small_coarse_inverse<NT>(mt, s, Ainv, lw, pinv);

        // ---- PCG, warm start from the previous pressure -------------------------------------------------
        double q2 = 0.0;  // ||q||^2 with coincident wells merged
        if (tid == 0) {
            for (int i = 0; i < a.w.n; ++i) {
                bool first = true;
                for (int k = 0; k < i; ++k) first = first && (wc[k] != wc[i]);
                if (first) {
                    const double q = cell_source(wc[i], a.w.n, wc, wr);
                    q2 += q * q;
                }
            }
            bc = q2;
        }
        __syncthreads();
        q2 = bc;
        int iters = 0;
        if (q2 == 0.0) {  // no sources: the pinned system has the zero solution
            for (int e = tid; e < M; e += NT) P[e] = 0.0;
            __syncthreads();
        } else {
            double rr = 0.0;
            for (int e = tid; e < M; e += NT) {
                const double r = cell_source(e, a.w.n, wc, wr) - small_Ax(P, TX0, TY0, e, ny, e == 0, pinv);
                R[e] = r;
                rr = fma(r, r, rr);
            }
            rr = block_sum_all(rr, red, flip);
            double rz_prev = 1.0;
            bool conv = false;
            for (int k = 0; k < a.max_iter; ++k) {
                if (!(rr > a.tol2 * q2)) {  // converged (or NaN: stop, flagged by the status pass)
                    conv = true;
                    break;
                }
                iters = k + 1;
                const double* Z = small_vcycle<NT, kSmallPer>(mt, s, XB, pinv, lw, Ainv) + o0;  // Z = M^-1 R
                double rz = 0.0;
                for (int e = tid; e < M; e += NT) rz = fma(R[e], Z[e], rz);
                rz = block_sum_all(rz, red, flip);
                const double beta = k > 0 ? rz / rz_prev : 0.0;
                for (int e = tid; e < M; e += NT) Pd[e] = k > 0 ? fma(beta, Pd[e], Z[e]) : Z[e];
                __syncthreads();
                double pAp = 0.0;
                for (int e = tid; e < M; e += NT) {
                    const double ap = small_Ax(Pd, TX0, TY0, e, ny, e == 0, pinv);
                    AP[e] = ap;
                    pAp = fma(Pd[e], ap, pAp);
                }
                pAp = block_sum_all(pAp, red, flip);
                const double alpha = rz / pAp;
                rr = 0.0;
                for (int e = tid; e < M; e += NT) {
                    P[e] = fma(alpha, Pd[e], P[e]);
                    const double r = fma(-alpha, AP[e], R[e]);
                    R[e] = r;
                    rr = fma(r, r, rr);
                }
                rr = block_sum_all(rr, red, flip);
                rz_prev = rz;
            }
            if (!conv && rr > a.tol2 * q2) cg_fail = 1;
        }
        tot_iters += iters;
        if (a.cg_iters && tid == 0) a.cg_iters[(int64_t)m * a.n_steps + step] = iters;

        // ---- face fluxes, CFL bound, sub-step count (Appendix A.2 tail, A.3 head) ------------------------
        // (unguarded neighbour reads: the pads of P are zero and the boundary faces have T = 0)
        double vxl[kSmallPer], vyl[kSmallPer], vxh[kSmallPer], vyh[kSmallPer];
        double pm = INFINITY;
#pragma unroll
        for (int k = 0; k < kSmallPer; ++k) {
            const int e = tid + k * NT;
            if (e < M) {
                const double pc = P[e];
                vxl[k] = (P[e - ny] - pc) * TX0[e];
                vyl[k] = (P[e - 1] - pc) * TY0[e];
                vxh[k] = (pc - P[e + ny]) * TX0[e + ny];
                vyh[k] = (pc - P[e + 1]) * TY0[e + 1];
                const double vi = fmax(vxl[k], 0.0) + fmax(vyl[k], 0.0) - fmin(vxh[k], 0.0) - fmin(vyh[k], 0.0);
                const double fi = fmax(cell_source(e, a.w.n, wc, wr), 0.0);
                const double pv = g.h2 * (a.por ? a.por[e] : 1.0);
                pm = fmin(pm, pv / (vi + fi));
            }
        }
        pm = block_min_all(pm, red, flip);  // (its barrier also ends every read of TX / TY / P above)
        const double cfl = ((1.0 - (g.swc + g.sor)) / 3.0) * pm;
        const double xn = ceil(a.dt / cfl);
        const int n = (xn >= 0.0 && xn < 2.0e9) ? (int)xn : 0;  // inf cfl (no flow) -> 0; NaN -> 0
        tot_sub += n;
        if (a.substeps && tid == 0) a.substeps[(int64_t)m * a.n_steps + step] = n;

        // ---- upwind coefficients of the frozen flux field (as in k_sat_cluster), into the free arrays -----
        const double dts = n > 0 ? a.dt / (double)n : 0.0;
        double* cW = s.X + o0;
        double* cS = s.B + o0;
        double* cN = DV0;
        double* cE = XB + o0;
        double* cD = AP;
        double* cQ = TX0;
        double* fw = TY0;  // padded: fw[e +- ny], fw[e +- 1] always addressable, multiplied by a zero coefficient outside
#pragma unroll
        for (int k = 0; k < kSmallPer; ++k) {
            const int e = tid + k * NT;
            if (e < M) {
                const double dtx = dts / (g.h2 * (a.por ? a.por[e] : 1.0));
                const double hdt = 0.5 * dtx;
                const double q = cell_source(e, a.w.n, wc, wr) * dtx;
                // max(v,0) = (v+|v|)/2, min(v,0) = (v-|v|)/2 (exact)
                cW[e] = hdt * (vxl[k] + fabs(vxl[k]));
                cS[e] = hdt * (vyl[k] + fabs(vyl[k]));
                cN[e] = hdt * (fabs(vyh[k]) - vyh[k]);
                cE[e] = hdt * (fabs(vxh[k]) - vxh[k]);
                cD[e] = hdt * (((vyl[k] - vyh[k]) + (vxl[k] - vxh[k])) -
                               ((fabs(vyl[k]) + fabs(vyh[k])) + (fabs(vxl[k]) + fabs(vxh[k])))) + fmin(q, 0.0);
                cQ[e] = fmax(q, 0.0);
            }
        }
        // ---- Nts explicit upwind sub-steps (Appendix A.3) -------------------------------------------------
        for (int sub = 0; sub < n; ++sub) {
            double f[kSmallPer];
#pragma unroll
            for (int k = 0; k < kSmallPer; ++k) {
                const int e = tid + k * NT;
                if (e < M) {
                    f[k] = a.unit_fluid ? frac_flow_loop<true>(S[e], a.fl) : frac_flow_loop<false>(S[e], a.fl);
                    fw[e] = f[k];
                }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kSmallPer; ++k) {
                const int e = tid + k * NT;
                if (e < M) {
                    double acc = fma(cD[e], f[k], cQ[e]);
                    acc = fma(cW[e], fw[e - ny], acc);
                    acc = fma(cS[e], fw[e - 1], acc);
                    acc = fma(cN[e], fw[e + 1], acc);
                    acc = fma(cE[e], fw[e + ny], acc);
                    S[e] += acc;
                }
            }
            __syncthreads();
        }
        // ---- outputs of the step; the arrays borrowed by the transport loop get their zero pads back ---------
        if (a.obs)
            for (int j = tid; j < a.n_obs; j += NT)
                a.obs[((int64_t)m * a.n_steps + step) * a.n_obs + j] = S[a.obs_cell[j]];
        const int hrow = hist_row(step + 1, a.n_steps, a.hist_stride);
        if (a.S_hist && hrow >= 0)
            for (int e = tid; e < M; e += NT)
                a.S_hist[((int64_t)m * hist_rows(a.n_steps, a.hist_stride) + hrow) * M + e] = S[e];
        __syncthreads();
    }
    int bad = 0;
    for (int e = tid; e < M; e += NT) {
        const double v = S[e];
        a.S_last[(int64_t)m * M + e] = v;
        if (a.P_last) a.P_last[(int64_t)m * M + e] = P[e];
        bad |= !isfinite(v);
    }
    bad = __syncthreads_or(bad);
    if (tid == 0) {
        if (a.status) a.status[m] = (cg_fail ? HM_MEMBER_CG_NOT_CONVERGED : 0) | (bad ? HM_MEMBER_NON_FINITE : 0);
        a.totals[2 * m] = tot_iters;
        a.totals[2 * m + 1] = tot_sub;
    }
}

}  // namespace

namespace {

// hierarchy + shared-memory layout of the fused kernel (see SmallArgs::mt); returns the dynamic smem bytes
size_t small_layout(int Nx, int Ny, OnchipMeta& mt, int& lw) {
    int nx = Nx, ny = Ny, o = 0;
    mt = OnchipMeta{};
    while (true) {
        const int l = mt.n++;
        const int pad = ny + 1;
        mt.nx[l] = nx;
        mt.ny[l] = ny;
        mt.M[l] = nx * ny;
        mt.off[l] = o + pad;
        mt.inv_ny[l] = 1.0f / (float)ny;
        o += nx * ny + 2 * pad;
        if (nx * ny <= 32 || mt.n == kMaxLevels) break;
        nx = (nx + 1) / 2;
        ny = (ny + 1) / 2;
    }
    lw = mt.n - 1;
    mt.total = o;
    mt.wmin = 0x7fffffff;
    const size_t M = (size_t)Nx * Ny, lenP = M + 2 * (size_t)(Ny + 1);
    return ((size_t)6 * o + 2 * lenP + 2 * M) * sizeof(double);
}

}  // namespace

namespace hmsim {

// The fused kernel covers grids of 4..2048 cells (whose padded hierarchy fits shared memory) with the default
// preconditioner (multigrid V-cycle); sat_block != 0 forces the streamed path (the tests cross-check the two).
int sim_small_supported(const hm_sim_desc& d) {
    const int64_t M = (int64_t)d.Nx * d.Ny;
    if (!(M >= 4 && M <= kSmallMaxCells && d.precond == 0 && d.sat_block == 0)) return 0;
    OnchipMeta mt;
    int lw;
    return small_layout(d.Nx, d.Ny, mt, lw) <= 220 * 1024 && mt.M[lw] <= 32;
}

int sim_small(hm_ctx* ctx, const hm_sim_desc& d, int m0, int nm) {
    cudaStream_t st = ctx->stream;
    SmallArgs a{};
    Geo& g = a.g;
    g.Nx = d.Nx;
    g.Ny = d.Ny;
    g.M = d.Nx * d.Ny;
    g.R = d.Nx;
    g.nTiles = 1;
    const double hx = d.Lx / d.Nx, hy = d.Ly / d.Ny;
    g.cx = 2 * hy / hx;
    g.cy = 2 * hx / hy;
    g.h2 = hx * hy;
    g.vw = d.vw;
    g.vo = d.vo;
    g.swc = d.swc;
    g.sor = d.sor;
    g.k_transform = d.K_transform;
    g.k_a = d.K_a;
    g.k_b = d.K_b;
    a.fl.inv_range = 1.0 / (1.0 - d.swc - d.sor);
    a.fl.swc_ir = d.swc * a.fl.inv_range;
    a.fl.mr = d.vw / d.vo;
    a.unit_fluid = a.fl.inv_range == 1.0 && a.fl.swc_ir == 0.0 && a.fl.mr == 1.0;
    const int64_t M = g.M;

    const size_t smem = small_layout(d.Nx, d.Ny, a.mt, a.lw);

    a.w.n = d.n_wells;
    a.w.cell = d.well_cell + (int64_t)m0 * d.well_cell_member_stride;
    a.w.cell_ms = d.well_cell_member_stride;
    a.w.rate = d.well_rate + (int64_t)m0 * d.well_rate_member_stride;
    a.w.rate_ms = d.well_rate_member_stride;
    a.w.rate_ss = d.well_rate_step_stride;
    a.K = d.K + (int64_t)m0 * d.K_member_stride;
    a.K_ms = d.K_member_stride;
    a.K_cs = d.K_comp_stride;
    a.por = d.por;
    a.S0 = d.S0 + (int64_t)m0 * d.S0_member_stride;
    a.S0_ms = d.S0_member_stride;
    a.dt = d.dt;
    a.n_steps = d.n_steps;
    a.n_obs = d.obs ? d.n_obs : 0;
    a.obs_cell = d.obs_cell;
    a.S_last = d.S_last + (int64_t)m0 * M;
    a.S_hist = d.S_hist ? d.S_hist + (int64_t)m0 * hist_rows(d.n_steps, d.hist_stride) * M : nullptr;
    a.hist_stride = d.hist_stride;
    a.obs = d.obs ? d.obs + (int64_t)m0 * d.n_steps * d.n_obs : nullptr;
    a.P_last = d.P_last ? d.P_last + (int64_t)m0 * M : nullptr;
    a.status = d.status ? d.status + m0 : nullptr;
    a.substeps = d.substeps ? d.substeps + (int64_t)m0 * d.n_steps : nullptr;
    a.cg_iters = d.cg_iters ? d.cg_iters + (int64_t)m0 * d.n_steps : nullptr;
    const double rtol = d.cg_rtol > 0 ? d.cg_rtol : 1e-12;
    a.tol2 = rtol * rtol;
    a.max_iter = d.cg_max_iter > 0 ? d.cg_max_iter : 100 * (d.Nx + d.Ny) + 200;
    HM_CHECK(ctx->ws.get("small.totals", (size_t)2 * nm, &a.totals));

    auto kern = M <= 2 * kSmallThreads ? k_sim_small<2> : M <= 4 * kSmallThreads ? k_sim_small<4> : k_sim_small<8>;
    HM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    HM_CUDA(cudaEventCreate(&e0));
    HM_CUDA(cudaEventCreate(&e1));
    HM_CUDA(cudaEventRecord(e0, st));
    kern<<<nm, kSmallThreads, smem, st>>>(a);
    HM_CUDA(cudaEventRecord(e1, st));
    HM_CUDA(cudaGetLastError());
    std::vector<int32_t> totals((size_t)2 * nm);
    HM_CUDA(cudaMemcpyAsync(totals.data(), a.totals, totals.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    HM_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    HM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    int32_t it = 0, sub = 0;
    for (int i = 0; i < nm; ++i) {
        it = std::max(it, totals[2 * i]);
        sub = std::max(sub, totals[2 * i + 1]);
    }
    // one launch holds every phase; the solve is > 90 % of it and the time is booked there
    ctx->phase_ms[1] += ms;
    ctx->sim_stats.cg_iterations += it;
    ctx->sim_stats.sat_substeps += sub;
    ctx->sim_stats.kernel_launches += 1;
    ctx->sim_stats.cg_kernel_launches += 1;
    return HM_OK;
}

}  // namespace hmsim
