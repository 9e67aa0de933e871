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
#include <type_traits>

#include "hm_mg_onchip.cuh"

namespace hmsim {

namespace {

constexpr int kOnchipCells = 4096;

// One grid level.  T is the arithmetic type of the preconditioner (double by default; float is an
// option - the V-cycle only has to be a good, fixed, symmetric approximation of A^-1 and CG itself
// stays FP64 - that costs ~10 % more iterations on smooth fields and up to 2x on rough ones).  On the top
// level (TOP) the right-hand side is the FP64 CG residual and the result z is written in FP64.
template <typename T>
struct Lvl {
    int nx, ny, M;
    int R, nTiles;
    const T* TX;
    const T* TY;
    const T* dinv;
    void* b;   // right-hand side: const double* on the top level, T* below
    T* xa;     // iterate after pre-smoothing
    void* xb;  // iterate after post-smoothing = the level's result: double* (z) on the top level, T* below
};

template <typename T>
__device__ __forceinline__ T* smem_as() {
    extern __shared__ __align__(16) unsigned char hm_smem_raw[];
    return reinterpret_cast<T*>(hm_smem_raw);
}

// (A x) at one cell.  xr points at the cell's row in a shared tile that has valid (finite) rows
// above and below; Tx / Ty point at the cell's own low-face transmissibilities.  No boundary
// tests: boundary faces carry T = 0 and every T array has a zero pad behind the last member, so
// the high faces Tx[ny] / Ty[1] are always readable and vanish where there is no neighbour.
template <typename T>
__device__ __forceinline__ T stencil(const T* xr, int col, int ny, const T* __restrict__ Tx,
                                     const T* __restrict__ Ty, bool cell0, T pin) {
    const T xc = xr[col];
    T y = Tx[0] * (xc - xr[col - ny]);
    y = fma(Tx[ny], xc - xr[col + ny], y);
    y = fma(Ty[0], xc - xr[col - 1], y);
    y = fma(Ty[1], xc - xr[col + 1], y);
    if (cell0) y = fma(pin, xc, y);
    return y;
}

// One stencil evaluation per cell of a whole grid row by a warp.  `emit(n_tag, col, c, xc, dv, bv, y)` receives
// n = n_tag.value consecutive cells starting at column col / cell c: their iterate, 1/diag, right-hand side and
// (A x).  Same arithmetic as stencil().
// NC != 0 (row length 32 NC known at compile time, a multiple of 128): a lane owns groups of FOUR ADJACENT cells
// (columns 128 g + 4 lane ..+3), so the operator rows, the three iterate rows, 1/diag and the right-hand side are
// 128-bit loads (7 shared-memory + 4 global load instructions per 4 cells instead of 28 + 16) and the results
// are stored as vectors: the streamed multigrid kernels were bound by the LSU instruction rate (ncu: 58-66 % LSU
// pipe, 34-37 % DRAM), not by HBM.  The global loads of a group pair are issued before its shared-memory phase.
// NC == 0: runtime row length, one cell per lane and step.
// (Measured and dropped: staging the operator rows in shared memory as well - 6 arrays per tile, 3 CTAs per SM
// instead of 5 - makes the pressure phase 9 % slower at 128^2 x 1024: occupancy matters more than the re-reads.)
template <typename T, typename TB, int NC, bool STAGE, typename F>
__device__ __forceinline__ void row_stencil(const T* xr, int ny, int lane, int c0, int li0,
                                            const T* __restrict__ TX, const T* __restrict__ TY,
                                            const T* __restrict__ dinv, const TB* __restrict__ b, const T* ds,
                                            const T* bs, T pinv, F&& emit) {
    if constexpr (NC != 0 && sizeof(T) == 4) {
        static_assert(NC % 4 == 0, "row length must be a multiple of 128");
        constexpr int NG = NC / 4;                                   // groups of 4 cells per lane
        constexpr int GB = (sizeof(T) == 4 && NG >= 2) ? 2 : 1;     // groups whose loads are in flight together
#pragma unroll
        for (int g0 = 0; g0 < NG; g0 += GB) {
            T tx0[GB][4], tx1[GB][4], ty0[GB][4], tyr[GB], dv[GB][4], bv[GB][4];
#pragma unroll
            for (int g = 0; g < GB; ++g) {
                const int col = 128 * (g0 + g) + 4 * lane, c = c0 + col;
                ldv<4>(TX + c, tx0[g]);
                ldv<4>(TX + c + ny, tx1[g]);
                ldv<4>(TY + c, ty0[g]);
                tyr[g] = TY[c + 4];
                if constexpr (STAGE) {
                    ldv<4>(ds + li0 + col, dv[g]);
                    ldv<4>(bs + li0 + col, bv[g]);
                } else {
                    ldv<4>(dinv + c, dv[g]);
                    TB bt[4];
                    ldv<4>(b + c, bt);
#pragma unroll
                    for (int k = 0; k < 4; ++k) bv[g][k] = (T)bt[k];
                }
            }
            asm volatile("" ::: "memory");  // keep every global load above the shared-memory phase
#pragma unroll
            for (int g = 0; g < GB; ++g) {
                const int col = 128 * (g0 + g) + 4 * lane, c = c0 + col;
                T xc[4], xu[4], xd[4], y[4];
                ldv<4>(xr + col, xc);
                ldv<4>(xr + col - ny, xu);
                ldv<4>(xr + col + ny, xd);
                const T xl = xr[col - 1], xh = xr[col + 4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const T lo = k == 0 ? xl : xc[k - 1], hi = k == 3 ? xh : xc[k + 1];
                    const T tyh = k == 3 ? tyr[g] : ty0[g][k + 1];
                    T v = tx0[g][k] * (xc[k] - xu[k]);
                    v = fma(tx1[g][k], xc[k] - xd[k], v);
                    v = fma(ty0[g][k], xc[k] - lo, v);
                    v = fma(tyh, xc[k] - hi, v);
                    if (c + k == 0) v = fma(pinv, xc[k], v);
                    y[k] = v;
                }
                emit(std::integral_constant<int, 4>{}, col, c, xc, dv[g], bv[g], y);
            }
        }
    } else if constexpr (NC != 0) {
        // FP64: one cell per lane and step (lanes along the row), the operator values of all the lane's cells
        // loaded first.  (Measured: the 4-adjacent-cells layout needs 60 registers in FP64 and is 14 % slower.)
        T tx0[NC], tx1[NC], ty0[NC], ty1[NC], dv[NC], bv[NC];
#pragma unroll
        for (int q = 0; q < NC; ++q) {
            const int col = lane + 32 * q, c = c0 + col;
            tx0[q] = TX[c];
            tx1[q] = TX[c + ny];
            ty0[q] = TY[c];
            ty1[q] = TY[c + 1];
            dv[q] = STAGE ? ds[li0 + col] : dinv[c];
            bv[q] = STAGE ? bs[li0 + col] : (T)b[c];
        }
        asm volatile("" ::: "memory");  // keep every global load above the shared-memory phase
#pragma unroll
        for (int q = 0; q < NC; ++q) {
            const int col = lane + 32 * q, c = c0 + col;
            const T xc[1] = {xr[col]};
            T y = tx0[q] * (xc[0] - xr[col - ny]);
            y = fma(tx1[q], xc[0] - xr[col + ny], y);
            y = fma(ty0[q], xc[0] - xr[col - 1], y);
            y = fma(ty1[q], xc[0] - xr[col + 1], y);
            if (c == 0) y = fma(pinv, xc[0], y);
            const T dvv[1] = {dv[q]}, bvv[1] = {bv[q]}, yv[1] = {y};
            emit(std::integral_constant<int, 1>{}, col, c, xc, dvv, bvv, yv);
        }
    } else {
        for (int col = lane; col < ny; col += 32) {
            const int c = c0 + col;
            const T xc[1] = {xr[col]};
            const T dvv[1] = {STAGE ? ds[li0 + col] : dinv[c]}, bvv[1] = {STAGE ? bs[li0 + col] : (T)b[c]};
            const T y[1] = {stencil<T>(xr, col, ny, TX + c, TY + c, c == 0, pinv)};
            emit(std::integral_constant<int, 1>{}, col, c, xc, dvv, bvv, y);
        }
    }
}

// ---- hierarchy ------------------------------------------------------------------------------
// FP64 level-0 operator -> arithmetic type of the preconditioner
template <typename T>
__global__ void k_mg_cast(int64_t n, const double* __restrict__ a, const double* __restrict__ b,
                          const double* __restrict__ c, T* __restrict__ oa, T* __restrict__ ob,
                          T* __restrict__ oc) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    oa[i] = (T)a[i];
    ob[i] = (T)b[i];
    oc[i] = (T)c[i];
}

template <typename T>
__global__ void k_mg_coarsen(int nm, Lvl<T> f, int cnx, int cny, T* __restrict__ cTX, T* __restrict__ cTY,
                             T* __restrict__ cdinv, const double* __restrict__ pin) {
    const int cM = cnx * cny;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)nm * cM) return;
    const int m = (int)(idx / cM), e = (int)(idx % cM);
    const int I = e / cny, J = e % cny;
    const T* TX = f.TX + (int64_t)m * f.M;
    const T* TY = f.TY + (int64_t)m * f.M;
    const int fi = 2 * I, fj = 2 * J;
    const bool j1 = fj + 1 < f.ny, i1 = fi + 1 < f.nx;
    auto tx = [&](int i) {  // half the sum of the fine x-faces on the low side of fine row i
        if (i >= f.nx) return (T)0;
        return (T)0.5 * (TX[i * f.ny + fj] + (j1 ? TX[i * f.ny + fj + 1] : (T)0));
    };
    auto ty = [&](int j) {
        if (j >= f.ny) return (T)0;
        return (T)0.5 * (TY[fi * f.ny + j] + (i1 ? TY[(fi + 1) * f.ny + j] : (T)0));
    };
    const T txl = tx(fi), txh = tx(fi + 2), tyl = ty(fj), tyh = ty(fj + 2);
    T d = tyl + tyh + txl + txh;
    if (e == 0) d += (T)pin[m];
    cTX[idx] = txl;
    cTY[idx] = tyl;
    cdinv[idx] = (T)1 / d;
}

// ---- streamed level: pre-smoothing + residual + restriction --------------------------------------
// Rows [r0,r1) of the tile (r0 even).  Sweep 1 (x = w0 D^-1 b, zero initial guess) covers rows
// [r0-NU, r1+NU); every further sweep shrinks the range by one row on either side, so after NU
// sweeps rows [r0-1, r1+1) are exact; then the residual on the tile rows and its 2x2 sums (the
// coarse right-hand side).  b, the operator and the sweeps ping-pong in shared memory; a warp
// walks whole grid rows (lanes along the contiguous index): no integer division.
template <typename T, bool TOP, int NY = 0>
__global__ void __launch_bounds__(kThreads)
k_mg_down(Lvl<T> f, int cny, T* __restrict__ cb, const double* __restrict__ pin, const int* __restrict__ done) {
    using TB = typename std::conditional<TOP, double, T>::type;
    T* sm = smem_as<T>();
    const int m = blockIdx.x / f.nTiles, t = blockIdx.x % f.nTiles;
    if (done[m]) return;
    constexpr int H = kNu;  // halo rows of the first sweep
    // NY != 0: row length known at compile time, the column loops unroll fully (independent loads in flight)
    const int ny = NY ? NY : f.ny, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nW = kThreads >> 5;
    const int r0 = t * f.R, r1 = min(r0 + f.R, f.nx), rows = r1 - r0;
    const int64_t off = (int64_t)m * f.M;
    const TB* __restrict__ b = static_cast<const TB*>(f.b) + off;
    const T* __restrict__ dinv = f.dinv + off;
    const T* __restrict__ TX = f.TX + off;
    const T* __restrict__ TY = f.TY + off;
    const T pinv = (T)pin[m];
    // FP32: right-hand side and D^-1 of the loaded rows are staged in shared memory as well; in FP64
    // that would halve the occupancy, they are re-read through L1 instead
    constexpr bool STAGE = sizeof(T) == 4;
    const int L = f.R + 2 * H;
    T* xa = sm;            // all arrays: local row lr <-> grid row r0 - H + lr
    T* xb = sm + L * ny;
    T* bs = sm + 2 * L * ny;
    T* ds = sm + 3 * L * ny;
    for (int lr = warp; lr < rows + 2 * H; lr += nW) {
        const int row = r0 - H + lr;
        const bool in = row >= 0 && row < f.nx;
        const int c0 = row * ny;
        if constexpr (NY != 0 && sizeof(T) == 4) {  // FP32: 4 adjacent cells per lane, 128-bit loads / stores
#pragma unroll
            for (int g = 0; g < NY / 128; ++g) {
                const int col = 128 * g + 4 * lane;
                TB bt[4] = {0, 0, 0, 0};
                T bv[4], dv[4] = {0, 0, 0, 0}, x0[4];
                const T zero[4] = {0, 0, 0, 0};
                if (in) {
                    ldv<4>(b + c0 + col, bt);
                    ldv<4>(dinv + c0 + col, dv);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    bv[k] = (T)bt[k];
                    x0[k] = (T)cheb_w(0) * dv[k] * bv[k];
                }
                if (STAGE) {
                    stv<4>(bs + lr * ny + col, bv);
                    stv<4>(ds + lr * ny + col, dv);
                }
                stv<4>(xa + lr * ny + col, x0);
                stv<4>(xb + lr * ny + col, zero);
            }
        } else {
#pragma unroll
            for (int col = lane; col < ny; col += 32) {
                const T bv = in ? (T)b[c0 + col] : (T)0, dv = in ? dinv[c0 + col] : (T)0;
                if (STAGE) {
                    bs[lr * ny + col] = bv;
                    ds[lr * ny + col] = dv;
                }
                xa[lr * ny + col] = (T)cheb_w(0) * dv * bv;
                xb[lr * ny + col] = (T)0;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int sw = 1; sw < kNu; ++sw) {  // sweep sw+1 is valid on local rows [sw, rows + 2H - sw)
        const T w = (T)cheb_w(sw);
        for (int lr = sw + warp; lr < rows + 2 * H - sw; lr += nW) {
            const int row = r0 - H + lr;
            if (row < 0 || row >= f.nx) continue;
            const int c0 = row * ny;
            const T* xr = xa + lr * ny;
            T* xo = xb + lr * ny;
            row_stencil<T, TB, NY / 32, STAGE>(xr, ny, lane, c0, lr * ny, TX, TY, dinv, b, ds, bs, pinv,
                                               [&](auto n, int col, int, const auto& xc, const auto& dv, const auto& bv,
                                                   const auto& y) {
                                                   constexpr int NV = decltype(n)::value;
                                                   T o[NV];
#pragma unroll
                                                   for (int k = 0; k < NV; ++k) o[k] = xc[k] + w * dv[k] * (bv[k] - y[k]);
                                                   stv<NV>(xo + col, o);
                                               });
        }
        __syncthreads();
        T* tsw = xa;
        xa = xb;
        xb = tsw;
    }
    // xa: pre-smoothed iterate, exact on local rows [H-1, H+rows+1); residual of the tile rows -> xb
    for (int lr = warp; lr < rows; lr += nW) {
        const int c0 = (r0 + lr) * ny;
        const T* xr = xa + (lr + H) * ny;
        T* xo = xb + lr * ny;
        T* xg = f.xa + off;
        row_stencil<T, TB, NY / 32, STAGE>(xr, ny, lane, c0, (lr + H) * ny, TX, TY, dinv, b, ds, bs, pinv,
                                           [&](auto n, int col, int c, const auto& xc, const auto&, const auto& bv,
                                               const auto& y) {
                                               constexpr int NV = decltype(n)::value;
                                               T o[NV], xv[NV];
#pragma unroll
                                               for (int k = 0; k < NV; ++k) o[k] = bv[k] - y[k], xv[k] = xc[k];
                                               stv<NV>(xg + c, xv);
                                               stv<NV>(xo + col, o);
                                           });
    }
    __syncthreads();
    const T* res = xb;
    const int crows = (rows + 1) / 2, cM = ((f.nx + 1) / 2) * cny;
    for (int I = warp; I < crows; I += nW) {
        const int lr = 2 * I;
        const bool two = lr + 1 < rows;
        T* out = cb + (int64_t)m * cM + (int64_t)(r0 / 2 + I) * cny;
        for (int J = lane; J < cny; J += 32) {
            const int col = 2 * J;
            const bool cc = col + 1 < ny;
            T sacc = res[lr * ny + col];
            if (cc) sacc += res[lr * ny + col + 1];
            if (two) {
                sacc += res[(lr + 1) * ny + col];
                if (cc) sacc += res[(lr + 1) * ny + col + 1];
            }
            out[J] = sacc;
        }
    }
}

// ---- streamed level: prolongation + post-smoothing (+ (r,z) on the top level) ----------------------
// x = xa + P xc on rows [r0-NU, r1+NU), then NU sweeps (weights in reverse order) on shrinking row
// ranges; the last sweep covers exactly the tile rows and is written to xb.
template <typename T, bool TOP, int NY = 0>
__global__ void __launch_bounds__(kThreads)
k_mg_up(Lvl<T> f, int cny, const T* __restrict__ cx, const double* __restrict__ pin,
        const int* __restrict__ done, double* __restrict__ part_rz) {
    using TB = typename std::conditional<TOP, double, T>::type;
    T* sm = smem_as<T>();
    __shared__ double red[32];
    const int m = blockIdx.x / f.nTiles, t = blockIdx.x % f.nTiles;
    if (done[m]) return;
    constexpr int H = kNu;
    // NY != 0: row length known at compile time, the column loops unroll fully (independent loads in flight)
    const int ny = NY ? NY : f.ny, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nW = kThreads >> 5;
    const int r0 = t * f.R, r1 = min(r0 + f.R, f.nx), rows = r1 - r0;
    const int64_t off = (int64_t)m * f.M;
    const int cM = ((f.nx + 1) / 2) * cny;
    const TB* __restrict__ b = static_cast<const TB*>(f.b) + off;
    TB* __restrict__ xout = static_cast<TB*>(f.xb) + off;
    const T* __restrict__ dinv = f.dinv + off;
    const T* __restrict__ TX = f.TX + off;
    const T* __restrict__ TY = f.TY + off;
    const T* __restrict__ xin = f.xa + off;
    const T* __restrict__ xc = cx + (int64_t)m * cM;
    const T pinv = (T)pin[m];
    constexpr bool STAGE = sizeof(T) == 4;
    const int L = f.R + 2 * H;
    T* xa = sm;
    T* xb = sm + L * ny;
    T* bs = sm + 2 * L * ny;
    T* ds = sm + 3 * L * ny;
    for (int lr = warp; lr < rows + 2 * H; lr += nW) {
        const int row = r0 - H + lr;
        const bool in = row >= 0 && row < f.nx;
        const int c0 = row * ny;
        const T* xcr = xc + (row >> 1) * cny;
        if constexpr (NY != 0 && sizeof(T) == 4) {  // FP32: 4 adjacent cells per lane, 128-bit loads / stores
#pragma unroll
            for (int g = 0; g < NY / 128; ++g) {
                const int col = 128 * g + 4 * lane;
                T x0[4] = {0, 0, 0, 0};
                const T zero[4] = {0, 0, 0, 0};
                if (in) {
                    ldv<4>(xin + c0 + col, x0);
                    const T ca = xcr[col >> 1], cb2 = xcr[(col >> 1) + 1];
                    x0[0] += ca, x0[1] += ca, x0[2] += cb2, x0[3] += cb2;
                }
                if (STAGE) {
                    TB bt[4] = {0, 0, 0, 0};
                    T bv[4], dv[4] = {0, 0, 0, 0};
                    if (in) {
                        ldv<4>(b + c0 + col, bt);
                        ldv<4>(dinv + c0 + col, dv);
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) bv[k] = (T)bt[k];
                    stv<4>(bs + lr * ny + col, bv);
                    stv<4>(ds + lr * ny + col, dv);
                }
                stv<4>(xa + lr * ny + col, x0);
                stv<4>(xb + lr * ny + col, zero);
            }
        } else {
#pragma unroll
            for (int col = lane; col < ny; col += 32) {
                if (STAGE) {
                    bs[lr * ny + col] = in ? (T)b[c0 + col] : (T)0;
                    ds[lr * ny + col] = in ? dinv[c0 + col] : (T)0;
                }
                xa[lr * ny + col] = in ? xin[c0 + col] + xcr[col >> 1] : (T)0;
                xb[lr * ny + col] = (T)0;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int sw = 0; sw < kNu - 1; ++sw) {  // sweep sw+1 is valid on local rows [sw+1, rows + 2H - sw - 1)
        const T w = (T)cheb_w(kNu - 1 - sw);
        for (int lr = sw + 1 + warp; lr < rows + 2 * H - sw - 1; lr += nW) {
            const int row = r0 - H + lr;
            if (row < 0 || row >= f.nx) continue;
            const int c0 = row * ny;
            const T* xr = xa + lr * ny;
            T* xo = xb + lr * ny;
            row_stencil<T, TB, NY / 32, STAGE>(xr, ny, lane, c0, lr * ny, TX, TY, dinv, b, ds, bs, pinv,
                                               [&](auto n, int col, int, const auto& xc, const auto& dv, const auto& bv,
                                                   const auto& y) {
                                                   constexpr int NV = decltype(n)::value;
                                                   T o[NV];
#pragma unroll
                                                   for (int k = 0; k < NV; ++k) o[k] = xc[k] + w * dv[k] * (bv[k] - y[k]);
                                                   stv<NV>(xo + col, o);
                                               });
        }
        __syncthreads();
        T* tsw = xa;
        xa = xb;
        xb = tsw;
    }
    double dot = 0.0;
    for (int lr = warp; lr < rows; lr += nW) {
        const int c0 = (r0 + lr) * ny;
        const T* xr = xa + (lr + H) * ny;
        row_stencil<T, TB, NY / 32, STAGE>(xr, ny, lane, c0, (lr + H) * ny, TX, TY, dinv, b, ds, bs, pinv,
                                           [&](auto n, int, int c, const auto& xc, const auto& dv, const auto& bv,
                                               const auto& y) {
                                               constexpr int NV = decltype(n)::value;
                                               TB o[NV];
#pragma unroll
                                               for (int k = 0; k < NV; ++k) {
                                                   const T v = xc[k] + (T)cheb_w(0) * dv[k] * (bv[k] - y[k]);
                                                   o[k] = (TB)v;
                                                   if (TOP) dot = fma((double)bv[k], (double)v, dot);
                                               }
                                               stv<NV>(xout + c, o);
                                           });
    }
    if (TOP) {
        dot = block_sum(dot, red);
        if (threadIdx.x == 0) part_rz[(int64_t)m * f.nTiles + t] = dot;
    }
}

// ---- all small levels in shared memory (device functions: hm_mg_onchip.cuh) -------------------------
// Dense inverse of the coarsest on-chip level (<= 32 cells) of every member, once per solve: one warp per
// member, Gauss-Jordan in shared memory (warp_dense_inverse), result [member][n][n] in global memory.
template <typename T>
__global__ void __launch_bounds__(256)
k_mg_dense_inverse(int nm, int n, int nx, int ny, const T* __restrict__ TX, const T* __restrict__ TY,
                   const double* __restrict__ pin, double* __restrict__ Ainv) {
    double* sm = smem_as<double>();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = blockIdx.x * 8 + warp;
    if (m >= nm) return;
    double* A = sm + warp * n * n;
    warp_dense_inverse<T>(n, nx, ny, TX + (int64_t)m * n, TY + (int64_t)m * n, pin[m], A);
    for (int e = lane; e < n * n; e += 32) Ainv[(int64_t)m * n * n + e] = A[e];
}

// One CTA per member runs the cycle (V, or W on the levels of at least `wmin` cells) on the shared-memory
// hierarchy; the coarsest level (<= 32 cells) is solved exactly with the precomputed dense inverse.
// FP64: 1024 threads, one CTA per SM (the hierarchy fills the shared memory).  FP32: the hierarchy is half the
// size, so two 512-thread CTAs (two members) share an SM and fill each other's barrier bubbles.
template <typename T>
struct OnchipCfg {
    static constexpr int NT = sizeof(T) == 4 ? 512 : 1024;
    static constexpr int CTAS = sizeof(T) == 4 ? 2 : 1;
};
template <typename T>
__global__ void __launch_bounds__(OnchipCfg<T>::NT, OnchipCfg<T>::CTAS)
k_mg_onchip(const __grid_constant__ OnchipMeta mt, const T* __restrict__ b_in, T* __restrict__ x_out,
            const double* __restrict__ pin, const int* __restrict__ done, const double* __restrict__ Ainv_g,
            int ainv_off /* bytes from the start of the dynamic shared memory, 8-aligned */) {
    constexpr int NT = OnchipCfg<T>::NT;
    T* sm = smem_as<T>();
    double* Ainv = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(sm) + ainv_off);
    const int m = blockIdx.x;
    if (done[m]) return;
    {
        const int nn = mt.M[mt.n - 1] * mt.M[mt.n - 1];
        for (int e = threadIdx.x; e < nn; e += NT) Ainv[e] = Ainv_g[(int64_t)m * nn + e];
    }
    OnchipSmem<T> s;
    s.X = sm;
    s.B = sm + mt.total;
    s.TX = sm + 2 * mt.total;
    s.TY = sm + 3 * mt.total;
    s.DV = sm + 4 * mt.total;
    for (int l = 0; l < mt.n; ++l) {
        const int64_t g = (int64_t)m * mt.M[l];
        const T* tx = static_cast<const T*>(mt.TX[l]) + g;
        const T* ty = static_cast<const T*>(mt.TY[l]) + g;
        const T* dv = static_cast<const T*>(mt.dinv[l]) + g;
        for (int e = threadIdx.x; e < mt.M[l]; e += NT) {
            s.TX[mt.off[l] + e] = tx[e];
            s.TY[mt.off[l] + e] = ty[e];
            s.DV[mt.off[l] + e] = dv[e];
        }
    }
    const int M0 = mt.M[0];
    for (int e = threadIdx.x; e < M0; e += NT) {
        s.B[e] = b_in[(int64_t)m * M0 + e];
        s.X[e] = 0;
    }
    __syncthreads();
    const T pinv = (T)pin[m];
    onchip_cycle<T, NT, kOnchipCells / NT>(mt, s, pinv, 0, Ainv);
    for (int e = threadIdx.x; e < M0; e += NT) x_out[(int64_t)m * M0 + e] = s.X[e];
}

// Vectorised variant for power-of-two hierarchies (v_cycle, hm_mg_onchip.cuh): padded X / TX / TY arrays, dense B /
// DV, dense inverse behind them.  Same launch shape as k_mg_onchip.
template <typename T>
__global__ void __launch_bounds__(OnchipCfg<T>::NT, OnchipCfg<T>::CTAS)
k_mg_onchip_v(const __grid_constant__ OnchipMeta mt, const T* __restrict__ b_in, T* __restrict__ x_out,
              const double* __restrict__ pin, const int* __restrict__ done, const double* __restrict__ Ainv_g,
              int ainv_off) {
    constexpr int NT = OnchipCfg<T>::NT;
    T* sm = smem_as<T>();
    double* Ainv = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(sm) + ainv_off);
    const int m = blockIdx.x;
    if (done[m]) return;
    OnchipSmem<T> s;
    s.X = sm;
    s.TX = sm + mt.ptotal;
    s.TY = sm + 2 * mt.ptotal;
    s.B = sm + 3 * mt.ptotal;
    s.DV = s.B + mt.total;
    const T zero[4] = {0, 0, 0, 0};
    for (int e = 4 * threadIdx.x; e < 3 * mt.ptotal; e += 4 * NT) stv<4>(sm + e, zero);  // iterate and the zero pads
    {
        const int nn = mt.M[mt.n - 1] * mt.M[mt.n - 1];
        for (int e = threadIdx.x; e < nn; e += NT) Ainv[e] = Ainv_g[(int64_t)m * nn + e];
    }
    __syncthreads();
    for (int l = 0; l < mt.n; ++l) {
        const int64_t g = (int64_t)m * mt.M[l];
        const T* tx = static_cast<const T*>(mt.TX[l]) + g;
        const T* ty = static_cast<const T*>(mt.TY[l]) + g;
        const T* dv = static_cast<const T*>(mt.dinv[l]) + g;
        for (int e = 4 * threadIdx.x; e < mt.M[l]; e += 4 * NT) {
            T v[4];
            ldv<4>(tx + e, v);
            stv<4>(s.TX + mt.poff[l] + e, v);
            ldv<4>(ty + e, v);
            stv<4>(s.TY + mt.poff[l] + e, v);
            ldv<4>(dv + e, v);
            stv<4>(s.DV + mt.off[l] + e, v);
        }
    }
    const int M0 = mt.M[0];
    for (int e = 4 * threadIdx.x; e < M0; e += 4 * NT) {
        T v[4];
        ldv<4>(b_in + (int64_t)m * M0 + e, v);
        stv<4>(s.B + e, v);
    }
    __syncthreads();
    v_cycle<T, NT, kOnchipCells / 4 / NT>(mt, s, (T)pin[m], Ainv);
    for (int e = 4 * threadIdx.x; e < M0; e += 4 * NT) {
        T v[4];
        ldv<4>(s.X + mt.poff[0] + e, v);
        stv<4>(x_out + (int64_t)m * M0 + e, v);
    }
}

// ---- CG kernels -------------------------------------------------------------------------------
// r = q - A x0 (warm start), optional Jacobi z = r/diag; partial (r,z), (r,r); ||q||^2
template <bool JACOBI>
__global__ void __launch_bounds__(kThreads)
k_cg_init(Geo g, Wells w, int step, double* __restrict__ X, const double* __restrict__ TXl,
          const double* __restrict__ TYl, const double* __restrict__ dinv, const double* __restrict__ pin,
          double* __restrict__ Rv, double* __restrict__ Z, double* __restrict__ part_rz,
          double* __restrict__ part_rr, double* __restrict__ bb, int* __restrict__ done,
          int* __restrict__ iters, int* __restrict__ counters, int init) {
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
        if (t == 0 && init) {  // init == 0: only the true residual of the current iterate is recomputed
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
        iters[m] += 1;  // iterations this member performs (member-local also across a reactivation, see k_cg_reactivate)
    }
}

// After a long solve: the recursively updated residual of CG drifts away from the true one (the gap grows with the
// iteration count and the condition number).  k_cg_init(init = 0) has recomputed r = q - A x; members whose TRUE
// residual is above the tolerance are reactivated.
__global__ void k_cg_reactivate(int nm, int nTiles, double tol2, const double* __restrict__ part_rr,
                                const double* __restrict__ bb, int* __restrict__ done, int* __restrict__ counters) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nm || !done[m]) return;
    double rr = 0.0;
    for (int t = 0; t < nTiles; ++t) rr += part_rr[(int64_t)m * nTiles + t];
    if (rr > tol2 * bb[m]) {
        done[m] = 0;
        atomicSub(&counters[0], 1);
    }
}

// p' = z + beta p ; Ap' ; partial (p',Ap').  Parity buffers: iteration k reads (r,z)[k&1] and
// p[k&1], writes p[(k+1)&1].
template <int NY = 0>
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
    const int ny = NY ? NY : g.Ny, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nW = kThreads >> 5;
    for (int lr = warp; lr < rows + 2; lr += nW) {
        const int row = r0 - 1 + lr;
        const bool in = row >= 0 && row < g.Nx, own = row >= r0 && row < r1;
        const int64_t c0 = off + (int64_t)row * ny;
#pragma unroll
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
#pragma unroll
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

// ---- multigrid hierarchy of one solve (host side) ---------------------------------------------------
template <typename T>
struct MgHierarchy {
    Lvl<T> lv[kMaxLevels];
    int nLev = 0, firstOn = 0, nm = 0;
    OnchipMeta mt{};
    size_t smemOn = 0;
    int ainvOff = 0;
    bool vecOn = false;  // the on-chip levels run the vectorised cycle (k_mg_onchip_v)
    const double* pin = nullptr;
    int* done = nullptr;
    double* part_rz = nullptr;
    size_t nPart = 0;
    double* Ainv = nullptr;  // [member][n][n]: dense inverse (FP64) of the coarsest level

    size_t smem_level(int l) const { return (size_t)(sizeof(T) == 4 ? 4 : 2) * (lv[l].R + 2 * kNu) * lv[l].ny * sizeof(T); }

    int build(hm_ctx* ctx, const Geo& g, int nm_, const double* TXl, const double* TYl, const double* dinv,
              const double* pin_, double* Rv, double* Z, bool wcycle, int* done_, double* part_rz_, size_t nPart_) {
        cudaStream_t st = ctx->stream;
        nm = nm_;
        pin = pin_;
        done = done_;
        part_rz = part_rz_;
        nPart = nPart_;
        const char* tag = sizeof(T) == 4 ? "f" : "d";
        int nx = g.Nx, ny = g.Ny;
        nLev = 0;
        while (true) {
            Lvl<T>& L = lv[nLev];
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
            // the hierarchy ends at the first level (below level 0) of <= 32 cells: it is solved exactly
            if ((nLev > 1 && nx * ny <= kDenseMax) || (nx == 1 && ny == 1) || nLev == kMaxLevels) break;
            nx = (nx + 1) / 2;
            ny = (ny + 1) / 2;
        }
        HM_REQUIRE(lv[nLev - 1].M <= kDenseMax, "multigrid hierarchy too deep");
        firstOn = 1;
        while (firstOn < nLev && lv[firstOn].M > kOnchipCells) ++firstOn;
        HM_REQUIRE(firstOn < nLev, "grid too large for the multigrid hierarchy");
        char name[40];
        for (int l = 0; l < nLev; ++l) {
            const size_t n = (size_t)nm * lv[l].M;
            T *tx, *ty, *dv;
            snprintf(name, sizeof name, "mg%s.TX%d", tag, l);
            HM_CHECK(ctx->ws.get(name, n + (size_t)lv[l].ny, &tx));  // + zero pads, see stencil()
            snprintf(name, sizeof name, "mg%s.TY%d", tag, l);
            HM_CHECK(ctx->ws.get(name, n + 1, &ty));
            snprintf(name, sizeof name, "mg%s.dv%d", tag, l);
            HM_CHECK(ctx->ws.get(name, n, &dv));
            HM_CUDA(cudaMemsetAsync(tx + n, 0, (size_t)lv[l].ny * sizeof(T), st));
            HM_CUDA(cudaMemsetAsync(ty + n, 0, sizeof(T), st));
            lv[l].TX = tx;
            lv[l].TY = ty;
            lv[l].dinv = dv;
            lv[l].b = nullptr;
            lv[l].xa = nullptr;
            lv[l].xb = nullptr;
            if (l == 0) {
                k_mg_cast<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>((int64_t)n, TXl, TYl, dinv, tx, ty, dv);
                lv[0].b = Rv;
                lv[0].xb = Z;
                snprintf(name, sizeof name, "mg%s.xa0", tag);
                HM_CHECK(ctx->ws.get(name, n, &lv[0].xa));
            } else {
                if (l <= firstOn) {  // streamed levels and the first on-chip level exchange b / x through HBM
                    T *bq, *xq;
                    snprintf(name, sizeof name, "mg%s.b%d", tag, l);
                    HM_CHECK(ctx->ws.get(name, n, &bq));
                    lv[l].b = bq;
                    snprintf(name, sizeof name, "mg%s.xb%d", tag, l);
                    HM_CHECK(ctx->ws.get(name, n, &xq));
                    lv[l].xb = xq;
                    if (l < firstOn) {
                        snprintf(name, sizeof name, "mg%s.xa%d", tag, l);
                        HM_CHECK(ctx->ws.get(name, n, &lv[l].xa));
                    }
                }
                k_mg_coarsen<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(nm, lv[l - 1], lv[l].nx, lv[l].ny, tx, ty,
                                                                              dv, pin);
            }
            ctx->sim_stats.kernel_launches += 1;
        }
        mt = OnchipMeta{};
        mt.n = nLev - firstOn;
        int o = 0;
        for (int i = 0; i < mt.n; ++i) {
            const Lvl<T>& L = lv[firstOn + i];
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
        mt.wmin = wcycle ? kWcycleMinCells : 0x7fffffff;
        ainvOff = (int)((((size_t)5 * o * sizeof(T)) + 7) & ~(size_t)7);
        smemOn = (size_t)ainvOff + (size_t)lv[nLev - 1].M * lv[nLev - 1].M * sizeof(double);
        // vectorised cycle (k_mg_onchip_v): V-cycle on a hierarchy whose levels above the coarsest have a power-of-two
        // row length in [8, 64], an even number of rows and halve exactly; padded layout for X, TX, TY
        vecOn = !wcycle && mt.n >= 2;
        int po = 0;
        for (int i = 0; i < mt.n; ++i) {
            const int ny_i = mt.ny[i], nx_i = mt.nx[i];
            if (i < mt.n - 1) {
                const bool pow2 = ny_i == 8 || ny_i == 16 || ny_i == 32 || ny_i == 64;
                vecOn = vecOn && pow2 && nx_i % 2 == 0 && mt.ny[i + 1] * 2 == ny_i && mt.nx[i + 1] * 2 == nx_i;
                int lg = 0;
                while ((4 << lg) < ny_i) ++lg;
                mt.lgpr[i] = lg;
            }
            // zero pad in front of the level: its own row length + 4 for the reads below index 0, and the row length of
            // the level above (stored in front of it), whose last row reads one row past its end
            const int pad = std::max(ny_i, i > 0 ? mt.ny[i - 1] : 0) + 4;
            po += (pad + 3) & ~3;
            mt.poff[i] = po;
            po += (mt.M[i] + 3) & ~3;
        }
        po += (mt.ny[mt.n - 1] + 4 + 3) & ~3;
        mt.ptotal = po;
        if (vecOn) {
            const size_t aoff = ((size_t)(3 * mt.ptotal + 2 * mt.total) * sizeof(T) + 7) & ~(size_t)7;
            const size_t sm_v = aoff + (size_t)lv[nLev - 1].M * lv[nLev - 1].M * sizeof(double);
            const size_t limit = (size_t)(232448 - 1024 * OnchipCfg<T>::CTAS) / OnchipCfg<T>::CTAS;
            if (sm_v <= limit) {
                ainvOff = (int)aoff;
                smemOn = sm_v;
                HM_CUDA(cudaFuncSetAttribute(k_mg_onchip_v<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemOn));
            } else {
                vecOn = false;
            }
        }
        HM_CUDA(cudaFuncSetAttribute(k_mg_onchip<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemOn));
        {   // dense inverse of the coarsest level, one warp per member
            const Lvl<T>& L = lv[nLev - 1];
            snprintf(name, sizeof name, "mg%s.Ainv", tag);
            HM_CHECK(ctx->ws.get(name, (size_t)nm * L.M * L.M, &Ainv));
            const size_t smd = (size_t)8 * L.M * L.M * sizeof(double);
            HM_CUDA(cudaFuncSetAttribute(k_mg_dense_inverse<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smd));
            k_mg_dense_inverse<T><<<(nm + 7) / 8, 256, smd, st>>>(nm, L.M, L.nx, L.ny, L.TX, L.TY, pin, Ainv);
            ctx->sim_stats.kernel_launches += 1;
        }
        size_t smax = 0;
        for (int l = 0; l < firstOn; ++l) smax = std::max(smax, smem_level(l));
        HM_REQUIRE(smax <= 200 * 1024, "row tile of a streamed multigrid level exceeds shared memory");
        if (smax > 40 * 1024) {
            HM_CUDA(cudaFuncSetAttribute(k_mg_down<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smax));
            HM_CUDA(cudaFuncSetAttribute(k_mg_down<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smax));
            HM_CUDA(cudaFuncSetAttribute(k_mg_up<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smax));
            HM_CUDA(cudaFuncSetAttribute(k_mg_up<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smax));
            HM_CUDA(cudaFuncSetAttribute(k_mg_down<T, true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smax));
            HM_CUDA(cudaFuncSetAttribute(k_mg_up<T, true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smax));
            HM_CUDA(cudaFuncSetAttribute(k_mg_down<T, true, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smax));
            HM_CUDA(cudaFuncSetAttribute(k_mg_up<T, true, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smax));
        }
        return HM_OK;
    }

    // z = M^-1 r by one multigrid cycle; the (r,z) partials go to the `parity` slot
    void apply(hm_ctx* ctx, int parity) {
        cudaStream_t st = ctx->stream;
        for (int l = 0; l < firstOn; ++l) {
            T* cb = static_cast<T*>(lv[l + 1].b);
            const int grid = nm * lv[l].nTiles;
            const size_t sm = smem_level(l);
            if (l == 0 && lv[0].ny == 128)
                k_mg_down<T, true, 128><<<grid, kThreads, sm, st>>>(lv[l], lv[l + 1].ny, cb, pin, done);
            else if (l == 0 && lv[0].ny == 512)
                k_mg_down<T, true, 512><<<grid, kThreads, sm, st>>>(lv[l], lv[l + 1].ny, cb, pin, done);
            else if (l == 0)
                k_mg_down<T, true><<<grid, kThreads, sm, st>>>(lv[l], lv[l + 1].ny, cb, pin, done);
            else
                k_mg_down<T, false><<<grid, kThreads, sm, st>>>(lv[l], lv[l + 1].ny, cb, pin, done);
        }
        if (vecOn)
            k_mg_onchip_v<T><<<nm, OnchipCfg<T>::NT, smemOn, st>>>(mt, static_cast<const T*>(lv[firstOn].b),
                                                                    static_cast<T*>(lv[firstOn].xb), pin, done, Ainv, ainvOff);
        else
            k_mg_onchip<T><<<nm, OnchipCfg<T>::NT, smemOn, st>>>(mt, static_cast<const T*>(lv[firstOn].b),
                                                                  static_cast<T*>(lv[firstOn].xb), pin, done, Ainv, ainvOff);
        for (int l = firstOn - 1; l >= 0; --l) {
            const T* cx = static_cast<const T*>(lv[l + 1].xb);
            const int grid = nm * lv[l].nTiles;
            const size_t sm = smem_level(l);
            double* prz = part_rz + parity * nPart;
            if (l == 0 && lv[0].ny == 128)
                k_mg_up<T, true, 128><<<grid, kThreads, sm, st>>>(lv[l], lv[l + 1].ny, cx, pin, done, prz);
            else if (l == 0 && lv[0].ny == 512)
                k_mg_up<T, true, 512><<<grid, kThreads, sm, st>>>(lv[l], lv[l + 1].ny, cx, pin, done, prz);
            else if (l == 0)
                k_mg_up<T, true><<<grid, kThreads, sm, st>>>(lv[l], lv[l + 1].ny, cx, pin, done, prz);
            else
                k_mg_up<T, false><<<grid, kThreads, sm, st>>>(lv[l], lv[l + 1].ny, cx, pin, done, nullptr);
        }
        ctx->sim_stats.kernel_launches += 2 * firstOn + 1;
        ctx->sim_stats.cg_kernel_launches += 2 * firstOn + 1;
    }
};

}  // namespace

// ---- host side -----------------------------------------------------------------------------------
int pressure_solve(hm_ctx* ctx, const Geo& g, const Wells& w, int step, int nm, const double* TXl,
                   const double* TYl, const double* dinv, const double* pin, double* P, double rtol,
                   int max_iter, int precond, int mg_switch_iters, int* done, int* iters, int* counters,
                   int* cg_batch, int* iters_used, bool* all_done_out) {
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
        HM_CUDA(cudaFuncSetAttribute(k_cg_spmv<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
        HM_CUDA(cudaFuncSetAttribute(k_cg_spmv<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
        HM_CUDA(cudaFuncSetAttribute(k_cg_spmv<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    }

    // ---- multigrid hierarchy --------------------------------------------------------------------------
    // precond 0 (default): V-cycle in FP32 arithmetic (operators, smoothing and transfers; CG itself, its
    // operator and the convergence test stay FP64, so the solution meets the same tolerance).  The cycle only
    // has to be a fixed SPD approximation of A^-1: on smooth log-normal fields FP32 costs ~3 % more iterations
    // at ~70 % of the time per iteration (measured at 128^2 x 1024: 104 vs 141 ms per 8 solves).  Fields that
    // need many iterations (rough, high contrast) lose more to the rounding noise of the cycle, so a solve that
    // has not converged after kSwitchIters iterations (twice that for the cold first solve of a run, which
    // starts from P = 0) restarts CG (beta = 0) with the FP64 cycle; when that happens on a warm-started solve
    // the rest of the forward run stays FP64 (ctx->mg_force64).  precond 4 = FP64 V-cycle, 2 = FP64 W-cycle,
    // 3 = FP32 V-cycle without fallback.
    const int kSwitchIters = mg_switch_iters > 0 ? mg_switch_iters : (step == 0 ? 80 : 40);
    MgHierarchy<float> mgf;
    MgHierarchy<double> mgd;
    const bool adaptive = precond == 0;
    bool mg32 = precond == 3 || (adaptive && !ctx->mg_force64);
    if (!jacobi) {
        if (mg32)
            HM_CHECK(mgf.build(ctx, g, nm, TXl, TYl, dinv, pin, Rv, Z, false, done, part_rz, nPart));
        else
            HM_CHECK(mgd.build(ctx, g, nm, TXl, TYl, dinv, pin, Rv, Z, precond == 2, done, part_rz, nPart));
    }
    auto precondition = [&](int parity) {
        if (mg32)
            mgf.apply(ctx, parity);
        else
            mgd.apply(ctx, parity);
    };
    // tiles start on even rows so that the 2x2 aggregates never straddle two tiles
    if (!jacobi) HM_REQUIRE(g.R % 2 == 0 || g.nTiles == 1, "level-0 tile height must be even");

    if (jacobi)
        k_cg_init<true><<<grid, kThreads, smem1, st>>>(g, w, step, P, TXl, TYl, dinv, pin, Rv, Z, part_rz, part_rr,
                                                        bb, done, iters, counters, 1);
    else
        k_cg_init<false><<<grid, kThreads, smem1, st>>>(g, w, step, P, TXl, TYl, dinv, pin, Rv, Z, part_rz, part_rr,
                                                         bb, done, iters, counters, 1);
    ctx->sim_stats.kernel_launches += 1;
    ctx->sim_stats.cg_kernel_launches += 1;

    int k = 0;
    bool all_done = false;
    bool restart = false;  // the next iteration starts a new Krylov space (preconditioner changed): p = z
    const int chk_blocks = (nm + 127) / 128;
    // Solves that took more than kVerifyIters iterations are verified against the TRUE residual and, if needed,
    // continued from a restart (at most twice): the residual gap of CG is negligible for the ~15-iteration solves of
    // the smooth priors and matters for ill-conditioned systems (anisotropic cells, extreme contrast: hundreds of
    // iterations).  Short solves pay nothing.
    constexpr int kVerifyIters = 50;
    for (int round = 0;; ++round) {
    while (k < max_iter && !all_done) {
        const int kend = std::min(max_iter, k + *cg_batch);
        for (; k < kend; ++k) {
            const int cur = k & 1, nxt = cur ^ 1;
            double* Pin = cur ? Pb : Pa;
            double* Pout = cur ? Pa : Pb;
            k_cg_check<<<chk_blocks, 128, 0, st>>>(nm, g.nTiles, k, tol2, part_rr, bb, done, iters, counters);
            if (!jacobi) precondition(cur);
            auto spmv = g.Ny == 128 ? k_cg_spmv<128> : g.Ny == 512 ? k_cg_spmv<512> : k_cg_spmv<0>;
            spmv<<<grid, kThreads, smem1, st>>>(g, restart ? 0 : k, Z, Pin, Pout, AP, TXl, TYl, pin,
                                                part_rz + cur * nPart, part_rz + nxt * nPart, part_pAp, done);
            restart = false;
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
        if (!all_done && adaptive && mg32 && !jacobi && k >= kSwitchIters) {
            HM_CHECK(mgd.build(ctx, g, nm, TXl, TYl, dinv, pin, Rv, Z, false, done, part_rz, nPart));
            mg32 = false;
            restart = true;
            if (step > 0 || mg_switch_iters > 0) ctx->mg_force64 = true;
            ctx->sim_stats.mg_fp64_fallbacks += 1;
        }
    }
    // the last update may have converged: one more test so that `done` is final
    k_cg_check<<<chk_blocks, 128, 0, st>>>(nm, g.nTiles, k, tol2, part_rr, bb, done, iters, counters);
    if (!all_done) {
        HM_CUDA(cudaMemcpyAsync(ctx->h_pinned, counters, sizeof(int), cudaMemcpyDeviceToHost, st));
        HM_CUDA(cudaStreamSynchronize(st));
        all_done = ctx->h_pinned[0] >= nm;
    }
    if (jacobi || !all_done || k < kVerifyIters || k >= max_iter || round >= 2) break;
    k_cg_init<false><<<grid, kThreads, smem1, st>>>(g, w, step, P, TXl, TYl, dinv, pin, Rv, Z, part_rz, part_rr, bb, done,
                                                     iters, counters, 0);
    k_cg_reactivate<<<chk_blocks, 128, 0, st>>>(nm, g.nTiles, tol2, part_rr, bb, done, counters);
    HM_CUDA(cudaMemcpyAsync(ctx->h_pinned, counters, sizeof(int), cudaMemcpyDeviceToHost, st));
    HM_CUDA(cudaStreamSynchronize(st));
    ctx->sim_stats.kernel_launches += 2;
    ctx->sim_stats.cg_kernel_launches += 2;
    if (ctx->h_pinned[0] >= nm) break;  // every member's true residual meets the tolerance
    all_done = false;
    restart = true;
    ctx->sim_stats.cg_restarts += 1;
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
