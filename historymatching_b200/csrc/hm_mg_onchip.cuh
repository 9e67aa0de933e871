// Multigrid hierarchy resident in shared memory: device functions shared by
//   * k_mg_onchip (hm_pressure.cu): the coarse levels (<= 4096 cells) of the streamed solver, and
//   * k_sim_small (hm_small.cu): the whole simulator of a small grid in one CTA.
// A CTA of NT threads owns one member; a thread handles up to PER cells of a level
// (cell e = tid + k*NT).  Every control variable is CTA-uniform.
#pragma once

#include "hm_sim_common.cuh"

namespace hmsim {

constexpr int kMaxLevels = 14;
constexpr int kWcycleMinCells = 64;  // recurse twice into a coarse level with at least this many cells
// NU Jacobi sweeps weighted by the reciprocal roots of the degree-NU Chebyshev polynomial on
// [2/8, 2] (the spectrum of D^-1 A lies in [0, 2]): same cost as damped Jacobi, markedly better
// smoothing.  Pre-smoothing applies the weights in order, post-smoothing in reverse order, which
// keeps the cycle symmetric (the smoother is a polynomial in D^-1 A either way).
#ifndef HM_NU
#define HM_NU 3
#endif
constexpr int kNu = HM_NU;
__device__ __forceinline__ double cheb_w(int i) {
    if (kNu == 2)  // interval [1/3, 2]
        return i == 0 ? 1.0 / 1.7559223176554566 : 1.0 / 0.57741101567787674;
    if (kNu == 3)  // interval [1/4, 2]: 1 / (1.125 + 0.875 cos(pi (2i+1) / 6))
        return i == 0 ? 1.0 / 1.8827722283113838 : i == 1 ? 1.0 / 1.125 : 1.0 / 0.36722777168861618;
    // kNu == 4, interval [1/5, 2]: 1 / (1.1 + 0.9 cos(pi (2i+1) / 8))
    return i == 0 ? 1.0 / 1.9314915792601581 : i == 1 ? 1.0 / 1.4444150891285809
         : i == 2 ? 1.0 / 0.75558491087141922 : 1.0 / 0.26850842073984186;
}

struct OnchipMeta {
    int n;                  // number of on-chip levels
    int nx[kMaxLevels], ny[kMaxLevels], M[kMaxLevels], off[kMaxLevels];
    float inv_ny[kMaxLevels];
    const void* TX[kMaxLevels];
    const void* TY[kMaxLevels];
    const void* dinv[kMaxLevels];
    int poff[kMaxLevels];   // vectorised cycle: offset of level l inside the PADDED arrays (X, TX, TY), a multiple of 4
    int lgpr[kMaxLevels];   // vectorised cycle: log2 of the 4-cell groups per row (ny / 4)
    int ptotal;             // vectorised cycle: length of a padded array
    int total;              // total cells over the on-chip levels
    int wmin;               // W-cycle: visit a coarse level twice if it has >= wmin cells (V-cycle: INT_MAX)
};

// X: iterate, B: right-hand side, TX / TY: low-face transmissibilities, DV: 1 / diagonal; all levels back to
// back, level l at offset mt.off[l]
template <typename T>
struct OnchipSmem {
    T *X, *B, *TX, *TY, *DV;
};

__device__ __forceinline__ void cell_ij(int e, int ny, float inv_ny, int& i, int& j) {
    i = __float2int_rd(((float)e + 0.5f) * inv_ny);  // exact for e < 2^22
    j = e - i * ny;
}

// (A x) at cell e = (i, j) of level l for the all-level vector array xv (level l at xv + mt.off[l])
template <typename T>
__device__ __forceinline__ T onchip_Axv(const OnchipMeta& mt, const OnchipSmem<T>& s, const T* xv, int l, int e, int i,
                                        int j, T pin) {
    const int ny = mt.ny[l], o = mt.off[l] + e;
    const T xc = xv[o];
    T y = 0;
    if (i > 0) y = s.TX[o] * (xc - xv[o - ny]);
    if (i < mt.nx[l] - 1) y = fma(s.TX[o + ny], xc - xv[o + ny], y);
    if (j > 0) y = fma(s.TY[o], xc - xv[o - 1], y);
    if (j < ny - 1) y = fma(s.TY[o + 1], xc - xv[o + 1], y);
    if (e == 0) y = fma(pin, xc, y);
    return y;
}
template <typename T>
__device__ __forceinline__ T onchip_Ax(const OnchipMeta& mt, const OnchipSmem<T>& s, int l, int e, int i, int j,
                                       T pin) {
    return onchip_Axv<T>(mt, s, s.X, l, e, i, j, pin);
}

// nsweep weighted-Jacobi sweeps on level l, in place (new values staged in registers).  zero_guess: the
// iterate is known to be zero, so the first sweep is x = w0 D^-1 b without a stencil (and one barrier).
template <typename T, int NT, int PER>
__device__ __forceinline__ void onchip_smooth(const OnchipMeta& mt, const OnchipSmem<T>& s, int l, T pin,
                                              int nsweep, bool reverse, bool zero_guess = false) {
    const int M = mt.M[l], ny = mt.ny[l], o = mt.off[l];
    const float inv = mt.inv_ny[l];
    int sw = 0;
    if (zero_guess) {
        const T wgt = (T)cheb_w(reverse ? kNu - 1 : 0);
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int e = threadIdx.x + k * NT;
            if (e < M) s.X[o + e] = (T)0 + wgt * s.DV[o + e] * (s.B[o + e] - (T)0);
        }
        __syncthreads();
        sw = 1;
    }
    for (; sw < nsweep; ++sw) {
        const T wgt = (T)cheb_w(reverse ? (kNu - 1 - sw % kNu) : sw % kNu);
        T xn[PER];
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int e = threadIdx.x + k * NT;
            if (e < M) {
                int i, j;
                cell_ij(e, ny, inv, i, j);
                xn[k] = s.X[o + e] + wgt * s.DV[o + e] * (s.B[o + e] - onchip_Ax<T>(mt, s, l, e, i, j, pin));
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int e = threadIdx.x + k * NT;
            if (e < M) s.X[o + e] = xn[k];
        }
        __syncthreads();
    }
}

// Residual of level l restricted to level l+1 (each coarse thread evaluates its own children);
// the coarse iterate is reset to zero.
template <typename T, int NT>
__device__ __forceinline__ void onchip_restrict(const OnchipMeta& mt, const OnchipSmem<T>& s, int l, T pin) {
    const int ny = mt.ny[l], nx = mt.nx[l], o = mt.off[l];
    const int cM = mt.M[l + 1], cny = mt.ny[l + 1], co = mt.off[l + 1];
    const float cinv = mt.inv_ny[l + 1];
    for (int e = threadIdx.x; e < cM; e += NT) {
        int ci, cj;
        cell_ij(e, cny, cinv, ci, cj);
        T r = 0;
#pragma unroll
        for (int di = 0; di < 2; ++di)
#pragma unroll
            for (int dj = 0; dj < 2; ++dj) {
                const int i = 2 * ci + di, j = 2 * cj + dj;
                if (i < nx && j < ny) {
                    const int fe = i * ny + j;
                    r += s.B[o + fe] - onchip_Ax<T>(mt, s, l, fe, i, j, pin);
                }
            }
        s.B[co + e] = r;
        s.X[co + e] = 0;
    }
    __syncthreads();
}

template <typename T, int NT>
__device__ __forceinline__ void onchip_prolong(const OnchipMeta& mt, const OnchipSmem<T>& s, int l) {
    const int M = mt.M[l], ny = mt.ny[l], o = mt.off[l], cny = mt.ny[l + 1], co = mt.off[l + 1];
    const float inv = mt.inv_ny[l];
    for (int e = threadIdx.x; e < M; e += NT) {
        int i, j;
        cell_ij(e, ny, inv, i, j);
        s.X[o + e] += s.X[co + (i >> 1) * cny + (j >> 1)];
    }
    __syncthreads();
}

// Coarse operators of level l+1 from level l (2x2 aggregation, piecewise-constant transfer, Galerkin / 2:
// coarse T = half the sum of the fine T crossing the coarse face); same arithmetic as k_mg_coarsen.
template <typename T, int NT>
__device__ __forceinline__ void onchip_coarsen(const OnchipMeta& mt, const OnchipSmem<T>& s, int l, T pin) {
    const int fnx = mt.nx[l], fny = mt.ny[l], fo = mt.off[l];
    const int cM = mt.M[l + 1], cny = mt.ny[l + 1], co = mt.off[l + 1];
    const float cinv = mt.inv_ny[l + 1];
    const T* TX = s.TX + fo;
    const T* TY = s.TY + fo;
    for (int e = threadIdx.x; e < cM; e += NT) {
        int I, J;
        cell_ij(e, cny, cinv, I, J);
        const int fi = 2 * I, fj = 2 * J;
        const bool j1 = fj + 1 < fny, i1 = fi + 1 < fnx;
        auto tx = [&](int i) {
            if (i >= fnx) return (T)0;
            return (T)0.5 * (TX[i * fny + fj] + (j1 ? TX[i * fny + fj + 1] : (T)0));
        };
        auto ty = [&](int j) {
            if (j >= fny) return (T)0;
            return (T)0.5 * (TY[fi * fny + j] + (i1 ? TY[(fi + 1) * fny + j] : (T)0));
        };
        const T txl = tx(fi), txh = tx(fi + 2), tyl = ty(fj), tyh = ty(fj + 2);
        T d = tyl + tyh + txl + txh;
        if (e == 0) d += pin;
        s.TX[co + e] = txl;
        s.TY[co + e] = tyl;
        s.DV[co + e] = (T)1 / d;
    }
    __syncthreads();
}

// Dense inverse of the operator of level l (n = mt.M[l] <= 32 cells; SPD thanks to the pin) -> A[n][n]
// (row pitch n), by Gauss-Jordan elimination without pivoting in ONE warp: lane r owns row r.  Always FP64,
// also under an FP32 hierarchy: the pinned operator is nearly singular (the constant mode is held by the pin
// alone) and its inverse has to be accurate for the cycle to stay symmetric positive definite.
constexpr int kDenseMax = 32;
template <typename T>
__device__ __forceinline__ void warp_dense_inverse(int n, int nx, int ny, const T* __restrict__ TX,
                                                   const T* __restrict__ TY, double pin, double* A /* [n][n] */) {
    const int r = threadIdx.x & 31;
    if (r < n) {
        const int i = r / ny, j = r - i * ny;
        const double txl = i > 0 ? (double)TX[r] : 0.0, txh = i < nx - 1 ? (double)TX[r + ny] : 0.0;
        const double tyl = j > 0 ? (double)TY[r] : 0.0, tyh = j < ny - 1 ? (double)TY[r + 1] : 0.0;
        for (int c = 0; c < n; ++c) A[r * n + c] = 0.0;
        A[r * n + r] = tyl + tyh + txl + txh + (r == 0 ? pin : 0.0);
        if (i > 0) A[r * n + r - ny] = -txl;
        if (i < nx - 1) A[r * n + r + ny] = -txh;
        if (j > 0) A[r * n + r - 1] = -tyl;
        if (j < ny - 1) A[r * n + r + 1] = -tyh;
    }
    __syncwarp();
    for (int k = 0; k < n; ++k) {
        const double pk = 1.0 / A[k * n + k];
        __syncwarp();
        if (r < n && r != k) {
            const double f = A[r * n + k] * pk;
            for (int c = 0; c < n; ++c)
                if (c != k) A[r * n + c] = fma(-f, A[k * n + c], A[r * n + c]);
            A[r * n + k] = -f;
        }
        __syncwarp();
        if (r < n && r != k) A[k * n + r] *= pk;
        if (r == k) A[k * n + k] = pk;
        __syncwarp();
    }
}

// One multigrid cycle on the shared-memory hierarchy, levels l0 .. mt.n-1: X[l0] = M^-1 B[l0] (X[l0] must be
// zero on entry).  The cycle (V, or W on the levels of at least `wmin` cells) is an explicit state machine:
// `left` packs, 4 bits per level, how many cycles are still to be run on that level.  The coarsest level
// mt.n-1 is solved exactly with its dense inverse `Ainv` (row pitch mt.M[mt.n-1]) when that is given, else
// by 3 kNu sweeps (a single cell: division).
template <typename T, int NT, int PER>
__device__ __forceinline__ void onchip_cycle(const OnchipMeta& mt, const OnchipSmem<T>& s, T pinv, int l0 = 0,
                                             const double* Ainv = nullptr) {
    unsigned long long left = ((mt.M[l0] >= mt.wmin) ? 2ull : 1ull) << (4 * l0);
    int l = l0;
    bool descend = true, fresh = true;  // fresh: the iterate of level l is still zero
    while (true) {
        if (descend) {  // start a cycle on level l
            if (l == mt.n - 1) {
                const int n = mt.M[l], o = mt.off[l];
                if (Ainv) {
                    if ((int)threadIdx.x < n) {
                        double acc = 0.0;  // exact solve (also on a W-cycle revisit, where it changes nothing)
                        for (int c = 0; c < n; ++c) acc = fma(Ainv[threadIdx.x * n + c], (double)s.B[o + c], acc);
                        s.X[o + threadIdx.x] = (T)acc;
                    }
                    __syncthreads();
                } else if (n == 1) {
                    if (threadIdx.x == 0) s.X[o] = s.B[o] * s.DV[o];
                    __syncthreads();
                } else {
                    onchip_smooth<T, NT, PER>(mt, s, l, pinv, 3 * kNu, false, fresh);
                }
                left -= 1ull << (4 * l);
                descend = false;
            } else {
                onchip_smooth<T, NT, PER>(mt, s, l, pinv, kNu, false, fresh);
                onchip_restrict<T, NT>(mt, s, l, pinv);
                ++l;
                fresh = true;
                left |= ((mt.M[l] >= mt.wmin) ? 2ull : 1ull) << (4 * l);
            }
        } else {  // a cycle on level l has just finished
            fresh = false;
            if ((left >> (4 * l)) & 15ull) {
                descend = true;
            } else if (l == l0) {
                break;
            } else {
                --l;
                onchip_prolong<T, NT>(mt, s, l);
                onchip_smooth<T, NT, PER>(mt, s, l, pinv, kNu, true);
                left -= 1ull << (4 * l);
            }
        }
    }
}

// ---- vectorised V-cycle for power-of-two hierarchies (k_mg_onchip_v) -----------------------------------------
// Every level above the coarsest has ny in {8, 16, 32, 64}, even nx, and halves exactly.  A thread handles groups of
// FOUR ADJACENT cells: iterate rows, operator rows, 1/diag and the right-hand side are 128-bit shared-memory
// accesses (11 load instructions per 4 cells instead of 44), and there are no boundary tests and no index
// divisions: X, TX, TY live in PADDED arrays (level l at mt.poff[l], at least ny + 4 zero entries on either side,
// boundary faces carry T = 0), B and DV in dense ones (mt.off[l]).  The residual is restricted with one warp
// shuffle (the two rows of a 2x2 aggregate sit in the same warp).  Same arithmetic per cell as onchip_cycle.
template <typename T>
__device__ __forceinline__ void v_stencil4(const OnchipMeta& mt, const OnchipSmem<T>& s, int l, int e, T pin, T (&xc)[4],
                                           T (&y)[4]) {
    const int ny = mt.ny[l], po = mt.poff[l] + e;
    T xu[4], xd[4], tx0[4], tx1[4], ty0[4];
    ldv<4>(s.X + po, xc);
    ldv<4>(s.X + po - ny, xu);
    ldv<4>(s.X + po + ny, xd);
    const T xl = s.X[po - 1], xh = s.X[po + 4];
    ldv<4>(s.TX + po, tx0);
    ldv<4>(s.TX + po + ny, tx1);
    ldv<4>(s.TY + po, ty0);
    const T tyr = s.TY[po + 4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const T lo = k == 0 ? xl : xc[k - 1], hi = k == 3 ? xh : xc[k + 1];
        const T tyh = k == 3 ? tyr : ty0[k + 1];
        T v = tx0[k] * (xc[k] - xu[k]);
        v = fma(tx1[k], xc[k] - xd[k], v);
        v = fma(ty0[k], xc[k] - lo, v);
        v = fma(tyh, xc[k] - hi, v);
        if (e + k == 0) v = fma(pin, xc[k], v);
        y[k] = v;
    }
}

template <typename T, int NT, int PER>
__device__ __forceinline__ void v_smooth(const OnchipMeta& mt, const OnchipSmem<T>& s, int l, T pin, bool reverse,
                                         bool zero_guess) {
    const int ng = mt.M[l] >> 2, o = mt.off[l], po = mt.poff[l];
    int sw = 0;
    if (zero_guess) {  // x = w0 D^-1 b
        const T wgt = (T)cheb_w(reverse ? kNu - 1 : 0);
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int g = threadIdx.x + k * NT;
            if (g < ng) {
                T dv[4], b[4], x[4];
                ldv<4>(s.DV + o + 4 * g, dv);
                ldv<4>(s.B + o + 4 * g, b);
#pragma unroll
                for (int i = 0; i < 4; ++i) x[i] = wgt * dv[i] * b[i];
                stv<4>(s.X + po + 4 * g, x);
            }
        }
        __syncthreads();
        sw = 1;
    }
    for (; sw < kNu; ++sw) {
        const T wgt = (T)cheb_w(reverse ? kNu - 1 - sw : sw);
        T xn[PER][4];
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int g = threadIdx.x + k * NT;
            if (g < ng) {
                T xc[4], y[4], dv[4], b[4];
                v_stencil4<T>(mt, s, l, 4 * g, pin, xc, y);
                ldv<4>(s.DV + o + 4 * g, dv);
                ldv<4>(s.B + o + 4 * g, b);
#pragma unroll
                for (int i = 0; i < 4; ++i) xn[k][i] = xc[i] + wgt * dv[i] * (b[i] - y[i]);
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int g = threadIdx.x + k * NT;
            if (g < ng) stv<4>(s.X + po + 4 * g, xn[k]);
        }
        __syncthreads();
    }
}

// residual of level l -> right-hand side of level l+1 (2x2 sums), coarse iterate reset to zero
template <typename T, int NT, int PER>
__device__ __forceinline__ void v_restrict(const OnchipMeta& mt, const OnchipSmem<T>& s, int l, T pin) {
    const int ng = mt.M[l] >> 2, o = mt.off[l], lg = mt.lgpr[l], gpr = 1 << lg;
    const int cny = mt.ny[l + 1], co = mt.off[l + 1], pco = mt.poff[l + 1];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int g = threadIdx.x + k * NT;
        if ((g & ~31) >= ng) break;  // warp-uniform
        T h0 = 0, h1 = 0;
        if (g < ng) {
            T xc[4], y[4], b[4];
            v_stencil4<T>(mt, s, l, 4 * g, pin, xc, y);
            ldv<4>(s.B + o + 4 * g, b);
            h0 = (b[0] - y[0]) + (b[1] - y[1]);
            h1 = (b[2] - y[2]) + (b[3] - y[3]);
        }
        const T v0 = h0 + __shfl_down_sync(0xffffffffu, h0, gpr);  // the row below: gpr lanes further
        const T v1 = h1 + __shfl_down_sync(0xffffffffu, h1, gpr);
        const int i = g >> lg, gi = g & (gpr - 1);
        if (g < ng && (i & 1) == 0) {
            const int ce = (i >> 1) * cny + 2 * gi;
            s.B[co + ce] = v0;
            s.B[co + ce + 1] = v1;
            s.X[pco + ce] = 0;
            s.X[pco + ce + 1] = 0;
        }
    }
    __syncthreads();
}

template <typename T, int NT, int PER>
__device__ __forceinline__ void v_prolong(const OnchipMeta& mt, const OnchipSmem<T>& s, int l) {
    const int ng = mt.M[l] >> 2, po = mt.poff[l], lg = mt.lgpr[l], gpr = 1 << lg;
    const int cny = mt.ny[l + 1], pco = mt.poff[l + 1];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int g = threadIdx.x + k * NT;
        if (g < ng) {
            const int i = g >> lg, gi = g & (gpr - 1);
            const T* xc = s.X + pco + (i >> 1) * cny + 2 * gi;
            const T c0 = xc[0], c1 = xc[1];
            T x[4];
            ldv<4>(s.X + po + 4 * g, x);
            x[0] += c0, x[1] += c0, x[2] += c1, x[3] += c1;
            stv<4>(s.X + po + 4 * g, x);
        }
    }
    __syncthreads();
}

// V-cycle on the padded hierarchy: X[level 0] = M^-1 B[level 0]; the coarsest level is solved with its dense inverse
template <typename T, int NT, int PER>
__device__ __forceinline__ void v_cycle(const OnchipMeta& mt, const OnchipSmem<T>& s, T pin, const double* Ainv) {
    const int lw = mt.n - 1;
    for (int l = 0; l < lw; ++l) {
        v_smooth<T, NT, PER>(mt, s, l, pin, false, true);
        v_restrict<T, NT, PER>(mt, s, l, pin);
    }
    {
        const int n = mt.M[lw], o = mt.off[lw], po = mt.poff[lw];
        if ((int)threadIdx.x < n) {
            double acc = 0.0;
            for (int c = 0; c < n; ++c) acc = fma(Ainv[threadIdx.x * n + c], (double)s.B[o + c], acc);
            s.X[po + threadIdx.x] = (T)acc;
        }
        __syncthreads();
    }
    for (int l = lw - 1; l >= 0; --l) {
        v_prolong<T, NT, PER>(mt, s, l);
        v_smooth<T, NT, PER>(mt, s, l, pin, true, false);
    }
}

}  // namespace hmsim
