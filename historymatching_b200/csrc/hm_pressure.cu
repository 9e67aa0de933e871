// Batched pressure solve: multigrid-preconditioned CG on the TPFA operator (sm_100a).
//
// Replaces scipy.sparse.linalg.spsolve inside TPFA_ResSim's pressure step
//  for a whole ensemble.  Matrix free: the operator of a
// level is given by its low-face transmissibilities TXl, TYl (+ the pin on cell 0),
//   (A x)_c = sum_faces T_f (x_c - x_nb) + pin * x_0 .
//
// Preconditioner: geometric multigrid, 2x2 cell aggregation, piecewise-constant
// transfer, Galerkin coarse operators scaled by 1/2 (coarse T = half the sum of the
// fine T crossing the coarse face, which is what re-discretisation would give),
// 2+2 damped-Jacobi sweeps, W-cycles on the coarse levels.  It is symmetric and
// fixed, so plain PCG applies.  Levels are processed in two ways:
//   * "streamed" levels (more than 4096 cells, and always level 0): row-tiled
//     kernels over HBM with halo rows recomputed in shared memory -
//     k_mg_down = 2 pre-sweeps + residual + restriction fused (50 B/cell),
//     k_mg_up   = prolongation + 2 post-sweeps (+ the (r,z) dot on level 0) (60 B/cell);
//   (precond 0 = V-cycle, the default; precond 2 = W-cycles on the shared-memory levels, more
//   robust for rough high-contrast fields at ~8x the coarse-level cost)
//   * all levels of at most 4096 cells: ONE kernel, one 1024-thread CTA per member,
//     the whole sub-hierarchy (operators + vectors, <= 218 KB) resident in shared
//     memory, no HBM traffic between the grid levels.
// CG itself is two more streamed kernels per iteration (k_cg_spmv 48 B/cell,
// k_cg_update 48 B/cell).  Members converge independently: a per-member `done`
// flag makes the CTAs of converged members exit at once.
#include "hm_sim_common.cuh"

namespace hmsim {

namespace {

constexpr int kMaxLevels = 14;
constexpr int kOnchipCells = 4096;
constexpr int kOnchipThreads = 1024;
constexpr int kWcycleMinCells = 64;  // recurse twice into a coarse level with at least this many cells
// Jacobi sweeps weighted by the roots of the degree-2 Chebyshev polynomial on [1/3, 2] (the
// spectrum of D^-1 A lies in [0, 2]): same cost as damped Jacobi, markedly better smoothing.
// Pre-smoothing applies (kW1, kW2), post-smoothing (kW2, kW1), which keeps the cycle symmetric.
constexpr double kW1 = 0.56950691011;  // 1 / 1.75592
constexpr double kW2 = 1.73205080757;  // 1 / 0.57735

struct Lvl {
    int nx, ny, M;
    int R, nTiles;
    const double* TX;
    const double* TY;
    const double* dinv;
    double* b;        // right-hand side (level 0: the CG residual)
    double* xa;       // iterate after pre-smoothing
    double* xb;       // iterate after post-smoothing = the level's result (level 0: z)
};

// (A x) at one cell.  xr points at the cell's row in a shared tile that has valid (finite) rows
// above and below; Tx / Ty point at the cell's own low-face transmissibilities.  No boundary
// tests: boundary faces carry T = 0 and every T array has a zero pad behind the last member, so
// the high faces Tx[ny] / Ty[1] are always readable and vanish where there is no neighbour.
__device__ __forceinline__ double stencil(const double* xr, int col, int ny, const double* __restrict__ Tx,
                                          const double* __restrict__ Ty, bool cell0, double pin) {
    const double xc = xr[col];
    double y = Tx[0] * (xc - xr[col - ny]);
    y = fma(Tx[ny], xc - xr[col + ny], y);
    y = fma(Ty[0], xc - xr[col - 1], y);
    y = fma(Ty[1], xc - xr[col + 1], y);
    if (cell0) y = fma(pin, xc, y);
    return y;
}

// ---- hierarchy ------------------------------------------------------------------------------
__global__ void k_mg_coarsen(int nm, Lvl f, int cnx, int cny, double* __restrict__ cTX,
                             double* __restrict__ cTY, double* __restrict__ cdinv,
                             const double* __restrict__ pin) {
    const int cM = cnx * cny;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)nm * cM) return;
    const int m = (int)(idx / cM), e = (int)(idx % cM);
    const int I = e / cny, J = e % cny;
    const double* TX = f.TX + (int64_t)m * f.M;
    const double* TY = f.TY + (int64_t)m * f.M;
    const int fi = 2 * I, fj = 2 * J;
    const bool j1 = fj + 1 < f.ny, i1 = fi + 1 < f.nx;
    auto tx = [&](int i) {  // half the sum of the fine x-faces on the low side of fine row i
        if (i >= f.nx) return 0.0;
        return 0.5 * (TX[i * f.ny + fj] + (j1 ? TX[i * f.ny + fj + 1] : 0.0));
    };
    auto ty = [&](int j) {
        if (j >= f.ny) return 0.0;
        return 0.5 * (TY[fi * f.ny + j] + (i1 ? TY[(fi + 1) * f.ny + j] : 0.0));
    };
    const double txl = tx(fi), txh = tx(fi + 2), tyl = ty(fj), tyh = ty(fj + 2);
    double d = tyl + tyh + txl + txh;
    if (e == 0) d += pin[m];
    cTX[idx] = txl;
    cTY[idx] = tyl;
    cdinv[idx] = 1.0 / d;
}

// ---- streamed level: pre-smoothing + residual + restriction --------------------------------------
// Rows [r0,r1) of the tile (r0 even).  x1 = w1 D^-1 b on rows [r0-2, r1+2), x2 = x1 + w2 D^-1 (b - A x1)
// on rows [r0-1, r1+1), residual on the tile rows, 2x2 sums of it to the coarse right-hand side.
// A warp walks whole grid rows (lanes along the contiguous index): no integer division.
__global__ void __launch_bounds__(kThreads)
k_mg_down(Lvl f, int cny, double* __restrict__ cb, const double* __restrict__ pin,
          const int* __restrict__ done) {
    extern __shared__ double sm[];
    const int m = blockIdx.x / f.nTiles, t = blockIdx.x % f.nTiles;
    if (done[m]) return;
    const int ny = f.ny, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nW = kThreads >> 5;
    const int r0 = t * f.R, r1 = min(r0 + f.R, f.nx), rows = r1 - r0;
    const int64_t off = (int64_t)m * f.M;
    const double* __restrict__ b = f.b + off;
    const double* __restrict__ dinv = f.dinv + off;
    const double* __restrict__ TX = f.TX + off;
    const double* __restrict__ TY = f.TY + off;
    const double pinv = pin[m];
    double* x1 = sm;                     // rows r0-2 .. r1+1   -> (rows+4) * ny
    double* x2 = sm + (f.R + 4) * ny;    // rows r0-1 .. r1     -> (rows+2) * ny
    for (int lr = warp; lr < rows + 4; lr += nW) {
        const int row = r0 - 2 + lr;
        const bool in = row >= 0 && row < f.nx;
        const int c0 = row * ny;
        for (int col = lane; col < ny; col += 32) x1[lr * ny + col] = in ? kW1 * dinv[c0 + col] * b[c0 + col] : 0.0;
    }
    __syncthreads();
    for (int lr = warp; lr < rows + 2; lr += nW) {
        const int row = r0 - 1 + lr;
        const bool in = row >= 0 && row < f.nx;
        const int c0 = row * ny;
        const double* xr = x1 + (lr + 1) * ny;
        for (int col = lane; col < ny; col += 32) {
            double v = 0.0;
            if (in) {
                const int c = c0 + col;
                v = xr[col] + kW2 * dinv[c] * (b[c] - stencil(xr, col, ny, TX + c, TY + c, c == 0, pinv));
            }
            x2[lr * ny + col] = v;
        }
    }
    __syncthreads();
    double* res = x1;  // x1 is dead: reuse for the residual of the tile rows, index (row-r0)*ny+col
    for (int lr = warp; lr < rows; lr += nW) {
        const int c0 = (r0 + lr) * ny;
        const double* xr = x2 + (lr + 1) * ny;
        for (int col = lane; col < ny; col += 32) {
            const int c = c0 + col;
            f.xa[off + c] = xr[col];
            res[lr * ny + col] = b[c] - stencil(xr, col, ny, TX + c, TY + c, c == 0, pinv);
        }
    }
    __syncthreads();
    const int crows = (rows + 1) / 2, cM = ((f.nx + 1) / 2) * cny;
    for (int I = warp; I < crows; I += nW) {
        const int lr = 2 * I;
        const bool two = lr + 1 < rows;
        double* out = cb + (int64_t)m * cM + (int64_t)(r0 / 2 + I) * cny;
        for (int J = lane; J < cny; J += 32) {
            const int col = 2 * J;
            const bool cc = col + 1 < ny;
            double sacc = res[lr * ny + col];
            if (cc) sacc += res[lr * ny + col + 1];
            if (two) {
                sacc += res[(lr + 1) * ny + col];
                if (cc) sacc += res[(lr + 1) * ny + col + 1];
            }
            out[J] = sacc;
        }
    }
}

// ---- streamed level: prolongation + post-smoothing (+ (r,z) on level 0) ----------------------------
template <bool DOT>
__global__ void __launch_bounds__(kThreads)
k_mg_up(Lvl f, int cny, const double* __restrict__ cx, const double* __restrict__ pin,
        const int* __restrict__ done, double* __restrict__ part_rz) {
    extern __shared__ double sm[];
    __shared__ double red[32];
    const int m = blockIdx.x / f.nTiles, t = blockIdx.x % f.nTiles;
    if (done[m]) return;
    const int ny = f.ny, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nW = kThreads >> 5;
    const int r0 = t * f.R, r1 = min(r0 + f.R, f.nx), rows = r1 - r0;
    const int64_t off = (int64_t)m * f.M;
    const int cM = ((f.nx + 1) / 2) * cny;
    const double* __restrict__ b = f.b + off;
    const double* __restrict__ dinv = f.dinv + off;
    const double* __restrict__ TX = f.TX + off;
    const double* __restrict__ TY = f.TY + off;
    const double* __restrict__ xa = f.xa + off;
    const double* __restrict__ xc = cx + (int64_t)m * cM;
    const double pinv = pin[m];
    double* x0 = sm;                     // rows r0-2 .. r1+1
    double* x3 = sm + (f.R + 4) * ny;    // rows r0-1 .. r1
    for (int lr = warp; lr < rows + 4; lr += nW) {
        const int row = r0 - 2 + lr;
        const bool in = row >= 0 && row < f.nx;
        const double* xar = xa + row * ny;
        const double* xcr = xc + (row >> 1) * cny;
        for (int col = lane; col < ny; col += 32) x0[lr * ny + col] = in ? xar[col] + xcr[col >> 1] : 0.0;
    }
    __syncthreads();
    for (int lr = warp; lr < rows + 2; lr += nW) {
        const int row = r0 - 1 + lr;
        const bool in = row >= 0 && row < f.nx;
        const int c0 = row * ny;
        const double* xr = x0 + (lr + 1) * ny;
        for (int col = lane; col < ny; col += 32) {
            double v = 0.0;
            if (in) {
                const int c = c0 + col;
                v = xr[col] + kW2 * dinv[c] * (b[c] - stencil(xr, col, ny, TX + c, TY + c, c == 0, pinv));
            }
            x3[lr * ny + col] = v;
        }
    }
    __syncthreads();
    double dot = 0.0;
    for (int lr = warp; lr < rows; lr += nW) {
        const int c0 = (r0 + lr) * ny;
        const double* xr = x3 + (lr + 1) * ny;
        for (int col = lane; col < ny; col += 32) {
            const int c = c0 + col;
            const double bc = b[c];
            const double v = xr[col] + kW1 * dinv[c] * (bc - stencil(xr, col, ny, TX + c, TY + c, c == 0, pinv));
            f.xb[off + c] = v;
            if (DOT) dot = fma(bc, v, dot);
        }
    }
    if (DOT) {
        dot = block_sum(dot, red);
        if (threadIdx.x == 0) part_rz[(int64_t)m * f.nTiles + t] = dot;
    }
}

// ---- all small levels in shared memory ----------------------------------------------------------
struct OnchipMeta {
    int n;                  // number of on-chip levels
    int nx[kMaxLevels], ny[kMaxLevels], M[kMaxLevels], off[kMaxLevels];
    float inv_ny[kMaxLevels];
    const double* TX[kMaxLevels];
    const double* TY[kMaxLevels];
    const double* dinv[kMaxLevels];
    int total;              // total cells over the on-chip levels
    int wmin;               // W-cycle: visit a coarse level twice if it has >= wmin cells (V-cycle: INT_MAX)
};

struct OnchipSmem {
    double *X, *B, *TX, *TY, *DV;
};

__device__ __forceinline__ void cell_ij(int e, int ny, float inv_ny, int& i, int& j) {
    i = __float2int_rd(((float)e + 0.5f) * inv_ny);  // exact for e < 2^22
    j = e - i * ny;
}

__device__ __forceinline__ double onchip_Ax(const OnchipMeta& mt, const OnchipSmem& s, int l, int e, int i, int j,
                                            double pin) {
    const int ny = mt.ny[l], o = mt.off[l] + e;
    const double xc = s.X[o];
    double y = 0.0;
    if (i > 0) y = s.TX[o] * (xc - s.X[o - ny]);
    if (i < mt.nx[l] - 1) y = fma(s.TX[o + ny], xc - s.X[o + ny], y);
    if (j > 0) y = fma(s.TY[o], xc - s.X[o - 1], y);
    if (j < ny - 1) y = fma(s.TY[o + 1], xc - s.X[o + 1], y);
    if (e == 0) y = fma(pin, xc, y);
    return y;
}

// nsweep damped-Jacobi sweeps on level l, in place (new values staged in registers)
__device__ __forceinline__ void onchip_smooth(const OnchipMeta& mt, const OnchipSmem& s, int l, double pin,
                                              int nsweep, double wa, double wb) {
    constexpr int PER = kOnchipCells / kOnchipThreads;
    const int M = mt.M[l], ny = mt.ny[l], o = mt.off[l];
    const float inv = mt.inv_ny[l];
    for (int sw = 0; sw < nsweep; ++sw) {
        const double wgt = (sw & 1) ? wb : wa;
        double xn[PER];
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int e = threadIdx.x + k * kOnchipThreads;
            if (e < M) {
                int i, j;
                cell_ij(e, ny, inv, i, j);
                xn[k] = s.X[o + e] + wgt * s.DV[o + e] * (s.B[o + e] - onchip_Ax(mt, s, l, e, i, j, pin));
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int e = threadIdx.x + k * kOnchipThreads;
            if (e < M) s.X[o + e] = xn[k];
        }
        __syncthreads();
    }
}

// Residual of level l restricted to level l+1 (each coarse thread evaluates its own children);
// the coarse iterate is reset to zero.
__device__ __forceinline__ void onchip_restrict(const OnchipMeta& mt, const OnchipSmem& s, int l, double pin) {
    const int ny = mt.ny[l], nx = mt.nx[l], o = mt.off[l];
    const int cM = mt.M[l + 1], cny = mt.ny[l + 1], co = mt.off[l + 1];
    const float cinv = mt.inv_ny[l + 1];
    for (int e = threadIdx.x; e < cM; e += kOnchipThreads) {
        int ci, cj;
        cell_ij(e, cny, cinv, ci, cj);
        double r = 0.0;
#pragma unroll
        for (int di = 0; di < 2; ++di)
#pragma unroll
            for (int dj = 0; dj < 2; ++dj) {
                const int i = 2 * ci + di, j = 2 * cj + dj;
                if (i < nx && j < ny) {
                    const int fe = i * ny + j;
                    r += s.B[o + fe] - onchip_Ax(mt, s, l, fe, i, j, pin);
                }
            }
        s.B[co + e] = r;
        s.X[co + e] = 0.0;
    }
    __syncthreads();
}

__device__ __forceinline__ void onchip_prolong(const OnchipMeta& mt, const OnchipSmem& s, int l) {
    const int M = mt.M[l], ny = mt.ny[l], o = mt.off[l], cny = mt.ny[l + 1], co = mt.off[l + 1];
    const float inv = mt.inv_ny[l];
    for (int e = threadIdx.x; e < M; e += kOnchipThreads) {
        int i, j;
        cell_ij(e, ny, inv, i, j);
        s.X[o + e] += s.X[co + (i >> 1) * cny + (j >> 1)];
    }
    __syncthreads();
}

// One CTA per member runs the cycle on the shared-memory hierarchy.  The cycle (V, or W on the
// levels of at least `wmin` cells) is an explicit state machine: `left` packs, 4 bits per level,
// how many cycles are still to be run on that level; every control variable is CTA-uniform.
__global__ void __launch_bounds__(kOnchipThreads, 1)
k_mg_onchip(const __grid_constant__ OnchipMeta mt, const double* __restrict__ b_in, double* __restrict__ x_out,
            const double* __restrict__ pin, const int* __restrict__ done) {
    extern __shared__ double sm[];
    const int m = blockIdx.x;
    if (done[m]) return;
    OnchipSmem s;
    s.X = sm;
    s.B = sm + mt.total;
    s.TX = sm + 2 * mt.total;
    s.TY = sm + 3 * mt.total;
    s.DV = sm + 4 * mt.total;
    for (int l = 0; l < mt.n; ++l) {
        const int64_t g = (int64_t)m * mt.M[l];
        for (int e = threadIdx.x; e < mt.M[l]; e += kOnchipThreads) {
            s.TX[mt.off[l] + e] = mt.TX[l][g + e];
            s.TY[mt.off[l] + e] = mt.TY[l][g + e];
            s.DV[mt.off[l] + e] = mt.dinv[l][g + e];
        }
    }
    const int M0 = mt.M[0];
    for (int e = threadIdx.x; e < M0; e += kOnchipThreads) {
        s.B[e] = b_in[(int64_t)m * M0 + e];
        s.X[e] = 0.0;
    }
    __syncthreads();
    const double pinv = pin[m];
    unsigned long long left = (M0 >= mt.wmin) ? 2ull : 1ull;
    int l = 0;
    bool descend = true;
    while (true) {
        if (descend) {  // start a cycle on level l
            if (l == mt.n - 1) {
                if (mt.M[l] == 1) {
                    if (threadIdx.x == 0) s.X[mt.off[l]] = s.B[mt.off[l]] * s.DV[mt.off[l]];
                    __syncthreads();
                } else {
                    onchip_smooth(mt, s, l, pinv, 8, kW1, kW2);
                }
                left -= 1ull << (4 * l);
                descend = false;
            } else {
                onchip_smooth(mt, s, l, pinv, 2, kW1, kW2);
                onchip_restrict(mt, s, l, pinv);
                ++l;
                left |= ((mt.M[l] >= mt.wmin) ? 2ull : 1ull) << (4 * l);
            }
        } else {  // a cycle on level l has just finished
            if ((left >> (4 * l)) & 15ull) {
                descend = true;
            } else if (l == 0) {
                break;
            } else {
                --l;
                onchip_prolong(mt, s, l);
                onchip_smooth(mt, s, l, pinv, 2, kW2, kW1);
                left -= 1ull << (4 * l);
            }
        }
    }
    for (int e = threadIdx.x; e < M0; e += kOnchipThreads) x_out[(int64_t)m * M0 + e] = s.X[e];
}

// ---- CG kernels -------------------------------------------------------------------------------
// r = q - A x0 (warm start), optional Jacobi z = r/diag; partial (r,z), (r,r); ||q||^2
template <bool JACOBI>
__global__ void __launch_bounds__(kThreads)
k_cg_init(Geo g, Wells w, int step, double* __restrict__ X, const double* __restrict__ TXl,
          const double* __restrict__ TYl, const double* __restrict__ dinv, const double* __restrict__ pin,
          double* __restrict__ Rv, double* __restrict__ Z, double* __restrict__ part_rz,
          double* __restrict__ part_rr, double* __restrict__ bb, int* __restrict__ done,
          int* __restrict__ iters, int* __restrict__ counters) {
    extern __shared__ double sm[];
    __shared__ int wc[kMaxWells];
    __shared__ double wr[kMaxWells];
    __shared__ double red[32];
    const int m = blockIdx.x / g.nTiles, t = blockIdx.x % g.nTiles;
    const int r0 = t * g.R, r1 = min(r0 + g.R, g.Nx), rows = r1 - r0;
    const int64_t off = (int64_t)m * g.M;
    load_wells(w, m, step, wc, wr);
    const int ny = g.Ny, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nW = kThreads >> 5;
    for (int lr = warp; lr < rows + 2; lr += nW) {
        const int row = r0 - 1 + lr;
        const bool in = row >= 0 && row < g.Nx;
        const double* xr = X + off + (int64_t)row * ny;
        for (int col = lane; col < ny; col += 32) sm[lr * ny + col] = in ? xr[col] : 0.0;
    }
    __syncthreads();
    double q2 = 0.0;  // ||q||^2 with coincident wells merged
    if (threadIdx.x == 0) {
        for (int i = 0; i < w.n; ++i) {
            bool first = true;
            for (int j = 0; j < i; ++j) first = first && (wc[j] != wc[i]);
            if (first) {
                const double q = cell_source(wc[i], w.n, wc, wr);
                q2 += q * q;
            }
        }
        red[0] = q2;
    }
    __syncthreads();
    q2 = red[0];
    __syncthreads();
    const double pinv = pin[m];
    double rz = 0.0, rr = 0.0;
    for (int lr = warp; lr < rows; lr += nW) {
        const int c0 = (r0 + lr) * ny;
        const double* xr = sm + (lr + 1) * ny;
        for (int col = lane; col < ny; col += 32) {
            const int c = c0 + col;
            double r = 0.0;
            if (q2 == 0.0) {  // no sources: the pinned system has the zero solution
                X[off + c] = 0.0;
            } else {
                r = cell_source(c, w.n, wc, wr) - stencil(xr, col, ny, TXl + off + c, TYl + off + c, c == 0, pinv);
            }
            Rv[off + c] = r;
            if (JACOBI) {
                const double z = r * dinv[off + c];
                Z[off + c] = z;
                rz = fma(r, z, rz);
            }
            rr = fma(r, r, rr);
        }
    }
    if (JACOBI) rz = block_sum(rz, red);
    rr = block_sum(rr, red);
    if (threadIdx.x == 0) {
        if (JACOBI) part_rz[(int64_t)m * g.nTiles + t] = rz;
        part_rr[(int64_t)m * g.nTiles + t] = rr;
        if (t == 0) {
            bb[m] = q2;
            iters[m] = 0;
            const int d = (q2 == 0.0);
            done[m] = d;
            if (d) atomicAdd(&counters[0], 1);
        }
    }
}

// convergence test of iteration k on the (r,r) partials; one thread per member
__global__ void k_cg_check(int nm, int nTiles, int k, double tol2, const double* __restrict__ part_rr,
                           const double* __restrict__ bb, int* __restrict__ done, int* __restrict__ iters,
                           int* __restrict__ counters) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nm || done[m]) return;
    double rr = 0.0;
    for (int t = 0; t < nTiles; ++t) rr += part_rr[(int64_t)m * nTiles + t];
    if (!(rr > tol2 * bb[m])) {  // converged (or NaN: stop, flagged by the status pass)
        done[m] = 1;
        atomicAdd(&counters[0], 1);
    } else {
        iters[m] = k + 1;
    }
}

// p' = z + beta p ; Ap' ; partial (p',Ap').  Parity buffers: iteration k reads (r,z)[k&1] and
// p[k&1], writes p[(k+1)&1].
__global__ void __launch_bounds__(kThreads)
k_cg_spmv(Geo g, int k, const double* __restrict__ Z, const double* __restrict__ Pin,
          double* __restrict__ Pout, double* __restrict__ AP, const double* __restrict__ TXl,
          const double* __restrict__ TYl, const double* __restrict__ pin,
          const double* __restrict__ part_rz_cur, const double* __restrict__ part_rz_prev,
          double* __restrict__ part_pAp, const int* __restrict__ done) {
    extern __shared__ double sm[];
    __shared__ double red[32];
    __shared__ double bc;
    const int m = blockIdx.x / g.nTiles, t = blockIdx.x % g.nTiles;
    if (done[m]) return;
    double beta = 0.0;
    if (k > 0) {
        const double rz = sum_partials(part_rz_cur + (int64_t)m * g.nTiles, g.nTiles, &bc);
        beta = rz / sum_partials(part_rz_prev + (int64_t)m * g.nTiles, g.nTiles, &bc);
    }
    const int r0 = t * g.R, r1 = min(r0 + g.R, g.Nx), rows = r1 - r0;
    const int64_t off = (int64_t)m * g.M;
    const int ny = g.Ny, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nW = kThreads >> 5;
    for (int lr = warp; lr < rows + 2; lr += nW) {
        const int row = r0 - 1 + lr;
        const bool in = row >= 0 && row < g.Nx, own = row >= r0 && row < r1;
        const int64_t c0 = off + (int64_t)row * ny;
        for (int col = lane; col < ny; col += 32) {
            double pn = 0.0;
            if (in) {
                pn = (k > 0) ? fma(beta, Pin[c0 + col], Z[c0 + col]) : Z[c0 + col];
                if (own) Pout[c0 + col] = pn;
            }
            sm[lr * ny + col] = pn;
        }
    }
    __syncthreads();
    const double pinv = pin[m];
    double pAp = 0.0;
    for (int lr = warp; lr < rows; lr += nW) {
        const int c0 = (r0 + lr) * ny;
        const double* xr = sm + (lr + 1) * ny;
        for (int col = lane; col < ny; col += 32) {
            const int c = c0 + col;
            const double ap = stencil(xr, col, ny, TXl + off + c, TYl + off + c, c == 0, pinv);
            AP[off + c] = ap;
            pAp = fma(xr[col], ap, pAp);
        }
    }
    pAp = block_sum(pAp, red);
    if (threadIdx.x == 0) part_pAp[(int64_t)m * g.nTiles + t] = pAp;
}

// x += a p ; r -= a Ap ; partial (r,r) [; Jacobi: z = r/diag, partial (r,z)]
template <bool JACOBI>
__global__ void __launch_bounds__(kThreads)
k_cg_update(Geo g, double* __restrict__ X, double* __restrict__ Rv, double* __restrict__ Z,
            const double* __restrict__ Pn, const double* __restrict__ AP, const double* __restrict__ dinv,
            const double* __restrict__ part_rz_cur, const double* __restrict__ part_pAp,
            double* __restrict__ part_rz_next, double* __restrict__ part_rr, const int* __restrict__ done) {
    __shared__ double red[32];
    __shared__ double bc;
    const int m = blockIdx.x / g.nTiles, t = blockIdx.x % g.nTiles;
    if (done[m]) return;
    const double rz = sum_partials(part_rz_cur + (int64_t)m * g.nTiles, g.nTiles, &bc);
    const double pAp = sum_partials(part_pAp + (int64_t)m * g.nTiles, g.nTiles, &bc);
    const double alpha = rz / pAp;
    const int r0 = t * g.R, r1 = min(r0 + g.R, g.Nx);
    const int64_t base = (int64_t)m * g.M + (int64_t)r0 * g.Ny;
    const int n = (r1 - r0) * g.Ny;
    double nrz = 0.0, nrr = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int64_t c = base + i;
        X[c] = fma(alpha, Pn[c], X[c]);
        const double r = fma(-alpha, AP[c], Rv[c]);
        Rv[c] = r;
        if (JACOBI) {
            const double z = r * dinv[c];
            Z[c] = z;
            nrz = fma(r, z, nrz);
        }
        nrr = fma(r, r, nrr);
    }
    if (JACOBI) nrz = block_sum(nrz, red);
    nrr = block_sum(nrr, red);
    if (threadIdx.x == 0) {
        if (JACOBI) part_rz_next[(int64_t)m * g.nTiles + t] = nrz;
        part_rr[(int64_t)m * g.nTiles + t] = nrr;
    }
}

}  // namespace

// ---- host side -----------------------------------------------------------------------------------
int pressure_solve(hm_ctx* ctx, const Geo& g, const Wells& w, int step, int nm, const double* TXl,
                   const double* TYl, const double* dinv, const double* pin, double* P, double rtol,
                   int max_iter, int precond, int* done, int* iters, int* counters, int* cg_batch,
                   int* iters_used, bool* all_done_out) {
    cudaStream_t st = ctx->stream;
    const int64_t M = g.M;
    const size_t vec = (size_t)nm * M;
    const size_t nPart = (size_t)nm * g.nTiles;
    const double tol2 = rtol * rtol;
    const bool jacobi = precond == 1 || M < 4;

    double *Rv, *Z, *AP, *Pa, *Pb, *part_rz, *part_rr, *part_pAp, *bb;
    HM_CHECK(ctx->ws.get("sim.r", vec, &Rv));
    HM_CHECK(ctx->ws.get("sim.z", vec, &Z));
    HM_CHECK(ctx->ws.get("sim.Ap", vec, &AP));
    HM_CHECK(ctx->ws.get("sim.pa", vec, &Pa));
    HM_CHECK(ctx->ws.get("sim.pb", vec, &Pb));
    HM_CHECK(ctx->ws.get("sim.part_rz", 2 * nPart, &part_rz));
    HM_CHECK(ctx->ws.get("sim.part_rr", nPart, &part_rr));
    HM_CHECK(ctx->ws.get("sim.part_pAp", nPart, &part_pAp));
    HM_CHECK(ctx->ws.get("sim.bb", (size_t)nm, &bb));

    const int grid = nm * g.nTiles;
    const size_t smem1 = (size_t)(g.R + 2) * g.Ny * sizeof(double);
    if (smem1 > 48 * 1024) {
        HM_CUDA(cudaFuncSetAttribute(k_cg_init<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
        HM_CUDA(cudaFuncSetAttribute(k_cg_init<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
        HM_CUDA(cudaFuncSetAttribute(k_cg_spmv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    }

    // ---- multigrid hierarchy ------------------------------------------------------------------
    Lvl lv[kMaxLevels];
    int nLev = 0, firstOn = 0;
    OnchipMeta mt{};
    size_t smemOn = 0;
    if (!jacobi) {
        int nx = g.Nx, ny = g.Ny;
        while (true) {
            Lvl& L = lv[nLev];
            L.nx = nx;
            L.ny = ny;
            L.M = nx * ny;
            if (nLev == 0) {  // level 0 shares the CG tiling (its (r,z) partials are summed per CG tile)
                L.R = g.R;
            } else {
                int R = std::max(2, std::min(nx, 4096 / ny));
                L.R = std::max(2, R & ~1);
            }
            L.nTiles = (nx + L.R - 1) / L.R;
            ++nLev;
            if ((nx == 1 && ny == 1) || nLev == kMaxLevels) break;
            nx = (nx + 1) / 2;
            ny = (ny + 1) / 2;
        }
        firstOn = 1;
        while (firstOn < nLev && lv[firstOn].M > kOnchipCells) ++firstOn;
        HM_REQUIRE(firstOn < nLev, "grid too large for the multigrid hierarchy");
        lv[0].TX = TXl;
        lv[0].TY = TYl;
        lv[0].dinv = dinv;
        lv[0].b = Rv;
        HM_CHECK(ctx->ws.get("mg.xa0", vec, &lv[0].xa));
        lv[0].xb = Z;
        for (int l = 1; l < nLev; ++l) {
            char name[32];
            const size_t n = (size_t)nm * lv[l].M;
            double *tx, *ty, *dv, *bq;
            snprintf(name, sizeof name, "mg.TX%d", l);
            HM_CHECK(ctx->ws.get(name, n + (size_t)lv[l].ny, &tx));  // + zero pads, see stencil()
            snprintf(name, sizeof name, "mg.TY%d", l);
            HM_CHECK(ctx->ws.get(name, n + 1, &ty));
            HM_CUDA(cudaMemsetAsync(tx + n, 0, (size_t)lv[l].ny * sizeof(double), st));
            HM_CUDA(cudaMemsetAsync(ty + n, 0, sizeof(double), st));
            snprintf(name, sizeof name, "mg.dv%d", l);
            HM_CHECK(ctx->ws.get(name, n, &dv));
            lv[l].TX = tx;
            lv[l].TY = ty;
            lv[l].dinv = dv;
            lv[l].b = nullptr;
            lv[l].xa = lv[l].xb = nullptr;
            if (l <= firstOn) {  // streamed levels and the first on-chip level exchange b / x through HBM
                snprintf(name, sizeof name, "mg.b%d", l);
                HM_CHECK(ctx->ws.get(name, n, &bq));
                lv[l].b = bq;
                snprintf(name, sizeof name, "mg.xb%d", l);
                HM_CHECK(ctx->ws.get(name, n, &lv[l].xb));
                if (l < firstOn) {
                    snprintf(name, sizeof name, "mg.xa%d", l);
                    HM_CHECK(ctx->ws.get(name, n, &lv[l].xa));
                }
            }
            const int64_t tot = (int64_t)nm * lv[l].M;
            k_mg_coarsen<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(nm, lv[l - 1], lv[l].nx, lv[l].ny, tx, ty, dv, pin);
            ctx->sim_stats.kernel_launches += 1;
        }
        mt.n = nLev - firstOn;
        int o = 0;
        for (int i = 0; i < mt.n; ++i) {
            const Lvl& L = lv[firstOn + i];
            mt.nx[i] = L.nx;
            mt.ny[i] = L.ny;
            mt.M[i] = L.M;
            mt.off[i] = o;
            mt.inv_ny[i] = 1.0f / (float)L.ny;
            mt.TX[i] = L.TX;
            mt.TY[i] = L.TY;
            mt.dinv[i] = L.dinv;
            o += L.M;
        }
        mt.total = o;
        mt.wmin = precond == 2 ? kWcycleMinCells : 0x7fffffff;
        smemOn = (size_t)5 * o * sizeof(double);
        HM_CUDA(cudaFuncSetAttribute(k_mg_onchip, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemOn));
        for (int l = 0; l < firstOn; ++l) {
            const size_t s2 = (size_t)(2 * lv[l].R + 6) * lv[l].ny * sizeof(double);
            if (s2 > 48 * 1024) {
                HM_CUDA(cudaFuncSetAttribute(k_mg_down, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2));
                HM_CUDA(cudaFuncSetAttribute(k_mg_up<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2));
                HM_CUDA(cudaFuncSetAttribute(k_mg_up<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2));
            }
        }
    }
    auto precondition = [&](int parity) {  // z = M^-1 r by one multigrid cycle; (r,z) partials -> parity slot
        for (int l = 0; l < firstOn; ++l) {
            const size_t s2 = (size_t)(2 * lv[l].R + 6) * lv[l].ny * sizeof(double);
            k_mg_down<<<nm * lv[l].nTiles, kThreads, s2, st>>>(lv[l], lv[l + 1].ny, lv[l + 1].b, pin, done);
        }
        k_mg_onchip<<<nm, kOnchipThreads, smemOn, st>>>(mt, lv[firstOn].b, lv[firstOn].xb, pin, done);
        for (int l = firstOn - 1; l >= 0; --l) {
            const size_t s2 = (size_t)(2 * lv[l].R + 6) * lv[l].ny * sizeof(double);
            if (l == 0)
                k_mg_up<true><<<nm * lv[l].nTiles, kThreads, s2, st>>>(lv[l], lv[l + 1].ny, lv[l + 1].xb, pin, done,
                                                                        part_rz + parity * nPart);
            else
                k_mg_up<false><<<nm * lv[l].nTiles, kThreads, s2, st>>>(lv[l], lv[l + 1].ny, lv[l + 1].xb, pin, done,
                                                                         nullptr);
        }
        ctx->sim_stats.kernel_launches += 2 * firstOn + 1;
        ctx->sim_stats.cg_kernel_launches += 2 * firstOn + 1;
    };
    // tiles start on even rows so that the 2x2 aggregates never straddle two tiles
    if (!jacobi) HM_REQUIRE(g.R % 2 == 0 || g.nTiles == 1, "level-0 tile height must be even");

    if (jacobi)
        k_cg_init<true><<<grid, kThreads, smem1, st>>>(g, w, step, P, TXl, TYl, dinv, pin, Rv, Z, part_rz, part_rr,
                                                        bb, done, iters, counters);
    else
        k_cg_init<false><<<grid, kThreads, smem1, st>>>(g, w, step, P, TXl, TYl, dinv, pin, Rv, Z, part_rz, part_rr,
                                                         bb, done, iters, counters);
    ctx->sim_stats.kernel_launches += 1;
    ctx->sim_stats.cg_kernel_launches += 1;

    int k = 0;
    bool all_done = false;
    const int chk_blocks = (nm + 127) / 128;
    while (k < max_iter && !all_done) {
        const int kend = std::min(max_iter, k + *cg_batch);
        for (; k < kend; ++k) {
            const int cur = k & 1, nxt = cur ^ 1;
            double* Pin = cur ? Pb : Pa;
            double* Pout = cur ? Pa : Pb;
            k_cg_check<<<chk_blocks, 128, 0, st>>>(nm, g.nTiles, k, tol2, part_rr, bb, done, iters, counters);
            if (!jacobi) precondition(cur);
            k_cg_spmv<<<grid, kThreads, smem1, st>>>(g, k, Z, Pin, Pout, AP, TXl, TYl, pin, part_rz + cur * nPart,
                                                      part_rz + nxt * nPart, part_pAp, done);
            if (jacobi)
                k_cg_update<true><<<grid, kThreads, 0, st>>>(g, P, Rv, Z, Pout, AP, dinv, part_rz + cur * nPart,
                                                              part_pAp, part_rz + nxt * nPart, part_rr, done);
            else
                k_cg_update<false><<<grid, kThreads, 0, st>>>(g, P, Rv, Z, Pout, AP, dinv, part_rz + cur * nPart,
                                                               part_pAp, part_rz + nxt * nPart, part_rr, done);
        }
        HM_CUDA(cudaMemcpyAsync(ctx->h_pinned, counters, sizeof(int), cudaMemcpyDeviceToHost, st));
        HM_CUDA(cudaStreamSynchronize(st));
        all_done = ctx->h_pinned[0] >= nm;
    }
    // the last update may have converged: one more test so that `done` is final
    k_cg_check<<<chk_blocks, 128, 0, st>>>(nm, g.nTiles, k, tol2, part_rr, bb, done, iters, counters);
    if (!all_done) {
        HM_CUDA(cudaMemcpyAsync(ctx->h_pinned, counters, sizeof(int), cudaMemcpyDeviceToHost, st));
        HM_CUDA(cudaStreamSynchronize(st));
        all_done = ctx->h_pinned[0] >= nm;
    }
    ctx->sim_stats.cg_iterations += k;
    ctx->sim_stats.kernel_launches += 3 * k + 1;
    ctx->sim_stats.cg_kernel_launches += 3 * k + 1;
    // adapt the convergence-check cadence to what this solve needed
    *cg_batch = jacobi ? std::max(8, std::min(64, k / 6 + 4)) : std::max(2, std::min(16, k / 4 + 1));
    *iters_used = k;
    *all_done_out = all_done;
    HM_CUDA(cudaGetLastError());
    return HM_OK;
}

}  // namespace hmsim
