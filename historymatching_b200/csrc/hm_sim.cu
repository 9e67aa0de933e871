// Ensemble forward run of the two-phase TPFA simulator on sm_100a.
//
// Replaces utils.apply(comp1, ...) over TPFA_ResSim.ResSim.sim
// (tools/utils.py:155-242, HistoryMatch.py:358-387) for a whole ensemble.
// Algorithm: SURVEY.md Appendix A (Aarnes-Gimse-Lie TPFA.m / RelPerm.m / Upstream.m).
//
// HBM layout (all FP64, member-major, cell c = ix*Ny + iy fastest):
//   S, P, r, z, Ap, p[2], TXl, TYl, dinv, Vxl, Vyl : [member][M]
//   "l" = the LOW face of a cell: TXl[c] is the transmissibility of the face
//   between (ix-1,iy) and (ix,iy) (0 for ix=0); the high face of c is the low
//   face of c+Ny.  Same for TYl / c+1 and for the fluxes Vxl, Vyl.
// A CTA owns a tile of R whole grid rows of one member (rows are contiguous,
// so halo rows are coalesced loads); neighbours inside the tile come from a
// shared-memory copy of the tile + 2 halo rows.
//
// Kernels and their algorithmic HBM bytes per cell:
//   k_tpfa_setup   read S,K            write TXl,TYl,dinv        40 B / solve
//   k_cg_spmv      read z,p,TXl,TYl    write p',Ap               48 B / iteration
//   k_cg_update    read x,r,p,Ap,dinv  write x,r,z               64 B / iteration
//   k_flux_cfl     read P,TXl,TYl      write Vxl,Vyl             40 B / solve
//   k_sat_substep  read S,Vxl,Vyl      write S'                  32 B / sub-step
#include "hm_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxWells = 64;
constexpr int kTileCells = 2048;

struct Geo {
    int Nx, Ny, M;
    int R;       // grid rows per tile
    int nTiles;  // tiles per member
    double cx;   // 2*hy/hx
    double cy;   // 2*hx/hy
    double h2;   // hx*hy
    double vw, vo, swc, sor;
};

struct Wells {
    int n;
    const int32_t* cell;
    int64_t cell_ms;
    const double* rate;
    int64_t rate_ms;
    int64_t rate_ss;
};

// ---- small device helpers ---------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum, result valid in thread 0.  `red` holds >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        v = (lane < (blockDim.x >> 5)) ? red[lane] : 0.0;
        v = warp_sum(v);
    }
    return v;
}
__device__ __forceinline__ double block_min(double v, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_min(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        v = (lane < (blockDim.x >> 5)) ? red[lane] : INFINITY;
        v = warp_min(v);
    }
    return v;
}

// Deterministic sum of the per-tile partials of one member, computed
// identically by every CTA of that member (so all tiles take the same
// convergence decision and use the same alpha/beta).  Valid in all threads.
__device__ __forceinline__ double sum_partials(const double* part, int n, double* bcast) {
    if (threadIdx.x < 32) {
        double v = 0.0;
        for (int i = threadIdx.x; i < n; i += 32) v += part[i];
        v = warp_sum(v);
        if (threadIdx.x == 0) *bcast = v;
    }
    __syncthreads();
    double out = *bcast;
    __syncthreads();
    return out;
}

__device__ __forceinline__ void load_wells(const Wells& w, int m, int step, int* wc, double* wr) {
    for (int i = threadIdx.x; i < w.n; i += blockDim.x) {
        wc[i] = w.cell[(int64_t)m * w.cell_ms + i];
        wr[i] = w.rate[(int64_t)m * w.rate_ms + (int64_t)step * w.rate_ss + i];
    }
}
// net source of cell c (wells sharing a cell accumulate, like np.add.at)
__device__ __forceinline__ double cell_source(int c, int nw, const int* wc, const double* wr) {
    double q = 0.0;
    for (int i = 0; i < nw; ++i)
        if (wc[i] == c) q += wr[i];
    return q;
}

__device__ __forceinline__ double total_mobility(double s, const Geo& g) {
    const double se = (s - g.swc) / (1.0 - g.swc - g.sor);
    return se * se / g.vw + (1.0 - se) * (1.0 - se) / g.vo;
}
__device__ __forceinline__ double frac_flow(double s, const Geo& g) {
    const double se = (s - g.swc) / (1.0 - g.swc - g.sor);
    const double lw = se * se / g.vw;
    const double lo = (1.0 - se) * (1.0 - se) / g.vo;
    return lw / (lw + lo);
}

// ---- K1: mobility + harmonic transmissibilities (Appendix A.2) -------------------------------
__global__ void __launch_bounds__(kThreads)
k_tpfa_setup(Geo g, const double* __restrict__ S, const double* __restrict__ K, int64_t Kms,
             int64_t Kcs, double* __restrict__ TXl, double* __restrict__ TYl,
             double* __restrict__ dinv, double* __restrict__ pin) {
    extern __shared__ double sm[];
    const int m = blockIdx.x / g.nTiles, t = blockIdx.x % g.nTiles;
    const int r0 = t * g.R, r1 = min(r0 + g.R, g.Nx), rows = r1 - r0;
    const int64_t off = (int64_t)m * g.M;
    const double* Sm = S + off;
    const double* Kx = K + (int64_t)m * Kms;
    const double* Ky = Kx + Kcs;
    double* Lx = sm;
    double* Ly = sm + (g.R + 2) * g.Ny;

    for (int i = threadIdx.x; i < (rows + 2) * g.Ny; i += blockDim.x) {
        const int row = r0 - 1 + i / g.Ny, col = i % g.Ny;
        double lx = 1.0, ly = 1.0;
        if (row >= 0 && row < g.Nx) {
            const int c = row * g.Ny + col;
            const double mt = total_mobility(Sm[c], g);
            lx = 1.0 / (mt * Kx[c]);
            ly = 1.0 / (mt * Ky[c]);
        }
        Lx[i] = lx;
        Ly[i] = ly;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < rows * g.Ny; i += blockDim.x) {
        const int li = i + g.Ny;
        const int row = r0 + i / g.Ny, col = i % g.Ny, c = row * g.Ny + col;
        const double txl = row > 0 ? g.cx / (Lx[li - g.Ny] + Lx[li]) : 0.0;
        const double txh = row < g.Nx - 1 ? g.cx / (Lx[li] + Lx[li + g.Ny]) : 0.0;
        const double tyl = col > 0 ? g.cy / (Ly[li - 1] + Ly[li]) : 0.0;
        const double tyh = col < g.Ny - 1 ? g.cy / (Ly[li] + Ly[li + 1]) : 0.0;
        double d = tyl + tyh + txl + txh;
        if (c == 0) {  // pin of the singular Neumann problem: A[0,0] += Kx[0]+Ky[0]
            const double pv = Kx[0] + Ky[0];
            d += pv;
            pin[m] = pv;
        }
        TXl[off + c] = txl;
        TYl[off + c] = tyl;
        dinv[off + c] = 1.0 / d;
    }
}

// y = A x on one cell, x taken from the shared tile (li = local index incl. halo row)
__device__ __forceinline__ double apply_A(const Geo& g, const double* xs, int li, int row, int col,
                                          int c, const double* __restrict__ TXl,
                                          const double* __restrict__ TYl, double pin) {
    const double xc = xs[li];
    const double txl = TXl[c];
    const double tyl = TYl[c];
    const double txh = row < g.Nx - 1 ? TXl[c + g.Ny] : 0.0;
    const double tyh = col < g.Ny - 1 ? TYl[c + 1] : 0.0;
    double y = txl * (xc - xs[li - g.Ny]);
    y = fma(txh, xc - xs[li + g.Ny], y);
    if (col > 0) y = fma(tyl, xc - xs[li - 1], y);
    if (col < g.Ny - 1) y = fma(tyh, xc - xs[li + 1], y);
    if (c == 0) y = fma(pin, xc, y);
    return y;
}

// ---- K2a: r = q - A x0, z = r/diag; partial (r,z), (r,r) ------------------------------------
__global__ void __launch_bounds__(kThreads)
k_cg_init(Geo g, Wells w, int step, double* __restrict__ X, const double* __restrict__ TXl,
          const double* __restrict__ TYl, const double* __restrict__ dinv,
          const double* __restrict__ pin, double* __restrict__ Rv, double* __restrict__ Z,
          double* __restrict__ part_rz, double* __restrict__ part_rr, double* __restrict__ bb,
          int* __restrict__ done, int* __restrict__ iters, int* __restrict__ counters) {
    extern __shared__ double sm[];
    __shared__ int wc[kMaxWells];
    __shared__ double wr[kMaxWells];
    __shared__ double red[32];
    const int m = blockIdx.x / g.nTiles, t = blockIdx.x % g.nTiles;
    const int r0 = t * g.R, r1 = min(r0 + g.R, g.Nx), rows = r1 - r0;
    const int64_t off = (int64_t)m * g.M;
    load_wells(w, m, step, wc, wr);
    for (int i = threadIdx.x; i < (rows + 2) * g.Ny; i += blockDim.x) {
        const int row = r0 - 1 + i / g.Ny;
        sm[i] = (row >= 0 && row < g.Nx) ? X[off + (int64_t)row * g.Ny + i % g.Ny] : 0.0;
    }
    __syncthreads();
    // ||q||^2 with coincident wells merged
    double q2 = 0.0;
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
    for (int i = threadIdx.x; i < rows * g.Ny; i += blockDim.x) {
        const int li = i + g.Ny;
        const int row = r0 + i / g.Ny, col = i % g.Ny, c = row * g.Ny + col;
        double r, z;
        if (q2 == 0.0) {  // no sources: the pinned system has the zero solution
            X[off + c] = 0.0;
            r = 0.0;
            z = 0.0;
        } else {
            r = cell_source(c, w.n, wc, wr) - apply_A(g, sm, li, row, col, c, TXl + off, TYl + off, pinv);
            z = r * dinv[off + c];
        }
        Rv[off + c] = r;
        Z[off + c] = z;
        rz = fma(r, z, rz);
        rr = fma(r, r, rr);
    }
    rz = block_sum(rz, red);
    rr = block_sum(rr, red);
    if (threadIdx.x == 0) {
        part_rz[(int64_t)m * g.nTiles + t] = rz;
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

// ---- K2b: p' = z + beta p; Ap' ; partial (p',Ap') ------------------------------------------
// Parity buffers: iteration k reads partials[k&1] (written by update k-1 / init)
// and p[k&1], writes p[(k+1)&1].
__global__ void __launch_bounds__(kThreads)
k_cg_spmv(Geo g, int k, double tol2, const double* __restrict__ Z, const double* __restrict__ Pin,
          double* __restrict__ Pout, double* __restrict__ AP, const double* __restrict__ TXl,
          const double* __restrict__ TYl, const double* __restrict__ pin,
          const double* __restrict__ part_rz_cur, const double* __restrict__ part_rz_prev,
          const double* __restrict__ part_rr_cur, double* __restrict__ part_pAp,
          const double* __restrict__ bb, int* __restrict__ done, int* __restrict__ iters,
          int* __restrict__ counters) {
    extern __shared__ double sm[];
    __shared__ double red[32];
    __shared__ double bc;
    const int m = blockIdx.x / g.nTiles, t = blockIdx.x % g.nTiles;
    if (done[m]) return;
    const double rr = sum_partials(part_rr_cur + (int64_t)m * g.nTiles, g.nTiles, &bc);
    if (!(rr > tol2 * bb[m])) {  // converged (or NaN: stop, flagged later)
        if (t == 0 && threadIdx.x == 0) {
            done[m] = 1;
            atomicAdd(&counters[0], 1);
        }
        return;
    }
    const double rz = sum_partials(part_rz_cur + (int64_t)m * g.nTiles, g.nTiles, &bc);
    double beta = 0.0;
    if (k > 0) beta = rz / sum_partials(part_rz_prev + (int64_t)m * g.nTiles, g.nTiles, &bc);

    const int r0 = t * g.R, r1 = min(r0 + g.R, g.Nx), rows = r1 - r0;
    const int64_t off = (int64_t)m * g.M;
    for (int i = threadIdx.x; i < (rows + 2) * g.Ny; i += blockDim.x) {
        const int row = r0 - 1 + i / g.Ny;
        double pn = 0.0;
        if (row >= 0 && row < g.Nx) {
            const int64_t c = off + (int64_t)row * g.Ny + i % g.Ny;
            pn = (k > 0) ? fma(beta, Pin[c], Z[c]) : Z[c];
            if (row >= r0 && row < r1) Pout[c] = pn;
        }
        sm[i] = pn;
    }
    __syncthreads();
    const double pinv = pin[m];
    double pAp = 0.0;
    for (int i = threadIdx.x; i < rows * g.Ny; i += blockDim.x) {
        const int li = i + g.Ny;
        const int row = r0 + i / g.Ny, col = i % g.Ny, c = row * g.Ny + col;
        const double ap = apply_A(g, sm, li, row, col, c, TXl + off, TYl + off, pinv);
        AP[off + c] = ap;
        pAp = fma(sm[li], ap, pAp);
    }
    pAp = block_sum(pAp, red);
    if (threadIdx.x == 0) {
        part_pAp[(int64_t)m * g.nTiles + t] = pAp;
        if (t == 0) iters[m] = k + 1;
    }
}

// ---- K2c: x += a p; r -= a Ap; z = r/diag; partial (r,z), (r,r) ---------------------------------
__global__ void __launch_bounds__(kThreads)
k_cg_update(Geo g, double* __restrict__ X, double* __restrict__ Rv, double* __restrict__ Z,
            const double* __restrict__ Pn, const double* __restrict__ AP,
            const double* __restrict__ dinv, const double* __restrict__ part_rz_cur,
            const double* __restrict__ part_pAp, double* __restrict__ part_rz_next,
            double* __restrict__ part_rr_next, const int* __restrict__ done) {
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
        const double pc = Pn[c];
        X[c] = fma(alpha, pc, X[c]);
        const double r = fma(-alpha, AP[c], Rv[c]);
        const double z = r * dinv[c];
        Rv[c] = r;
        Z[c] = z;
        nrz = fma(r, z, nrz);
        nrr = fma(r, r, nrr);
    }
    nrz = block_sum(nrz, red);
    nrr = block_sum(nrr, red);
    if (threadIdx.x == 0) {
        part_rz_next[(int64_t)m * g.nTiles + t] = nrz;
        part_rr_next[(int64_t)m * g.nTiles + t] = nrr;
    }
}

// ---- K3: face fluxes + CFL bound (Appendix A.2 tail, A.3 head) ---------------------------------
__global__ void __launch_bounds__(kThreads)
k_flux_cfl(Geo g, Wells w, int step, const double* __restrict__ P, const double* __restrict__ TXl,
           const double* __restrict__ TYl, const double* __restrict__ por,
           double* __restrict__ Vxl, double* __restrict__ Vyl, double* __restrict__ part_pm) {
    extern __shared__ double sm[];
    __shared__ int wc[kMaxWells];
    __shared__ double wr[kMaxWells];
    __shared__ double red[32];
    const int m = blockIdx.x / g.nTiles, t = blockIdx.x % g.nTiles;
    const int r0 = t * g.R, r1 = min(r0 + g.R, g.Nx), rows = r1 - r0;
    const int64_t off = (int64_t)m * g.M;
    load_wells(w, m, step, wc, wr);
    for (int i = threadIdx.x; i < (rows + 2) * g.Ny; i += blockDim.x) {
        const int row = r0 - 1 + i / g.Ny;
        sm[i] = (row >= 0 && row < g.Nx) ? P[off + (int64_t)row * g.Ny + i % g.Ny] : 0.0;
    }
    __syncthreads();
    double pm = INFINITY;
    for (int i = threadIdx.x; i < rows * g.Ny; i += blockDim.x) {
        const int li = i + g.Ny;
        const int row = r0 + i / g.Ny, col = i % g.Ny, c = row * g.Ny + col;
        const double pc = sm[li];
        const double vxl = row > 0 ? (sm[li - g.Ny] - pc) * TXl[off + c] : 0.0;
        const double vyl = col > 0 ? (sm[li - 1] - pc) * TYl[off + c] : 0.0;
        const double vxh = row < g.Nx - 1 ? (pc - sm[li + g.Ny]) * TXl[off + c + g.Ny] : 0.0;
        const double vyh = col < g.Ny - 1 ? (pc - sm[li + 1]) * TYl[off + c + 1] : 0.0;
        Vxl[off + c] = vxl;
        Vyl[off + c] = vyl;
        // total influx of the cell, same association as the reference expression
        const double vi = fmax(vxl, 0.0) + fmax(vyl, 0.0) - fmin(vxh, 0.0) - fmin(vyh, 0.0);
        const double fi = fmax(cell_source(c, w.n, wc, wr), 0.0);
        const double pv = g.h2 * (por ? por[c] : 1.0);
        pm = fmin(pm, pv / (vi + fi));
    }
    pm = block_min(pm, red);
    if (threadIdx.x == 0) part_pm[(int64_t)m * g.nTiles + t] = pm;
}

// Nts = ceil(dt / cfl), cfl = ((1-swc-sor)/3) * min(pv/(Vi+fi))
__global__ void k_substep_count(Geo g, int n_members, double dt, const double* __restrict__ part_pm,
                                int* __restrict__ nts, int* __restrict__ counters,
                                int32_t* __restrict__ substeps_out, int n_steps, int step) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_members) return;
    double pm = INFINITY;
    for (int t = 0; t < g.nTiles; ++t) pm = fmin(pm, part_pm[(int64_t)m * g.nTiles + t]);
    const double cfl = ((1.0 - (g.swc + g.sor)) / 3.0) * pm;
    const double x = ceil(dt / cfl);
    int n = (x >= 0.0 && x < 2.0e9) ? (int)x : 0;  // inf cfl (no flow) -> 0; NaN -> 0
    nts[m] = n;
    atomicMax(&counters[1], n);
    if (substeps_out) substeps_out[(int64_t)m * n_steps + step] = n;
}

// ---- K4: one explicit upwind sub-step (Appendix A.3) ---------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_sat_substep(Geo g, Wells w, int step, int it, double dt, const int* __restrict__ nts,
              const double* __restrict__ Sin, double* __restrict__ Sout,
              const double* __restrict__ Vxl, const double* __restrict__ Vyl,
              const double* __restrict__ por) {
    extern __shared__ double sm[];  // fractional flow of tile + halo rows
    __shared__ int wc[kMaxWells];
    __shared__ double wr[kMaxWells];
    const int m = blockIdx.x / g.nTiles, t = blockIdx.x % g.nTiles;
    const int r0 = t * g.R, r1 = min(r0 + g.R, g.Nx), rows = r1 - r0;
    const int64_t off = (int64_t)m * g.M;
    const int n = nts[m];
    if (it >= n) {  // this member needs fewer sub-steps: carry its state over
        for (int i = threadIdx.x; i < rows * g.Ny; i += blockDim.x) {
            const int64_t c = off + (int64_t)r0 * g.Ny + i;
            Sout[c] = Sin[c];
        }
        return;
    }
    load_wells(w, m, step, wc, wr);
    for (int i = threadIdx.x; i < (rows + 2) * g.Ny; i += blockDim.x) {
        const int row = r0 - 1 + i / g.Ny;
        sm[i] = (row >= 0 && row < g.Nx) ? frac_flow(Sin[off + (int64_t)row * g.Ny + i % g.Ny], g) : 0.0;
    }
    __syncthreads();
    const double dts = dt / (double)n;
    for (int i = threadIdx.x; i < rows * g.Ny; i += blockDim.x) {
        const int li = i + g.Ny;
        const int row = r0 + i / g.Ny, col = i % g.Ny, c = row * g.Ny + col;
        const double vxl = Vxl[off + c];
        const double vyl = Vyl[off + c];
        const double vxh = row < g.Nx - 1 ? Vxl[off + c + g.Ny] : 0.0;
        const double vyh = col < g.Ny - 1 ? Vyl[off + c + 1] : 0.0;
        const double q = cell_source(c, w.n, wc, wr);
        const double fi = fmax(q, 0.0), fp = fmin(q, 0.0);
        const double dtx = dts / (g.h2 * (por ? por[c] : 1.0));
        // B row: [x2(c-Ny), y2(c-1), diag, -y1(c+1), -x1(c+Ny)] scaled by dtx
        double acc = (dtx * fmax(vxl, 0.0)) * sm[li - g.Ny];
        if (col > 0) acc = fma(dtx * fmax(vyl, 0.0), sm[li - 1], acc);
        const double diag = fp + fmin(vyl, 0.0) - fmax(vyh, 0.0) + fmin(vxl, 0.0) - fmax(vxh, 0.0);
        acc = fma(dtx * diag, sm[li], acc);
        if (col < g.Ny - 1) acc = fma(dtx * -fmin(vyh, 0.0), sm[li + 1], acc);
        acc = fma(dtx * -fmin(vxh, 0.0), sm[li + g.Ny], acc);
        Sout[off + c] = Sin[off + c] + (acc + fi * dtx);
    }
}

// ---- obs gather / history / status ------------------------------------------------------------
__global__ void k_gather_obs(int n_members, int M, int n_obs, const int32_t* __restrict__ obs_cell,
                             const double* __restrict__ S, double* __restrict__ obs, int n_steps,
                             int step) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_members * n_obs) return;
    const int m = i / n_obs, j = i % n_obs;
    obs[((int64_t)m * n_steps + step) * n_obs + j] = S[(int64_t)m * M + obs_cell[j]];
}

__global__ void k_copy_rows(int n_members, int M, const double* __restrict__ src, int64_t src_ms,
                            double* __restrict__ dst, int64_t dst_ms) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_members * M) return;
    const int64_t m = i / M, c = i % M;
    dst[m * dst_ms + c] = src[m * src_ms + c];
}

__global__ void k_member_status(int n_members, int M, const double* __restrict__ S,
                                const int* __restrict__ cg_fail, int32_t* __restrict__ status) {
    __shared__ int bad;
    const int m = blockIdx.x;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    int b = 0;
    for (int c = threadIdx.x; c < M; c += blockDim.x) b |= !isfinite(S[(int64_t)m * M + c]);
    if (b) bad = 1;
    __syncthreads();
    if (threadIdx.x == 0)
        status[m] = (cg_fail[m] ? HM_MEMBER_CG_NOT_CONVERGED : 0) | (bad ? HM_MEMBER_NON_FINITE : 0);
}

__global__ void k_mark_unconverged(int n_members, const int* __restrict__ done, int* __restrict__ cg_fail) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < n_members && !done[m]) cg_fail[m] = 1;
}

__global__ void k_record_iters(int n_members, const int* __restrict__ iters, int32_t* __restrict__ out,
                               int n_steps, int step) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < n_members) out[(int64_t)m * n_steps + step] = iters[m];
}

struct PhaseTimer {
    std::vector<cudaEvent_t> ev;
    std::vector<int> phase;
    cudaStream_t st;
    explicit PhaseTimer(cudaStream_t s) : st(s) {}
    void mark(int ph) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        ev.push_back(e);
        phase.push_back(ph);
    }
    void finish(double out[5]) {
        for (size_t i = 0; i + 1 < ev.size(); ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
            if (phase[i] >= 0 && phase[i] < 5) out[phase[i]] += ms;
        }
        for (auto e : ev) cudaEventDestroy(e);
        ev.clear();
    }
};

int sim_chunk(hm_ctx* ctx, const hm_sim_desc& d, int m0, int nm) {
    cudaStream_t st = ctx->stream;
    Geo g;
    g.Nx = d.Nx;
    g.Ny = d.Ny;
    g.M = d.Nx * d.Ny;
    g.R = std::max(1, std::min(d.Nx, kTileCells / d.Ny));
    g.nTiles = (d.Nx + g.R - 1) / g.R;
    const double hx = d.Lx / d.Nx, hy = d.Ly / d.Ny;
    g.cx = 2 * hy / hx;
    g.cy = 2 * hx / hy;
    g.h2 = hx * hy;
    g.vw = d.vw;
    g.vo = d.vo;
    g.swc = d.swc;
    g.sor = d.sor;
    const int64_t M = g.M;
    const size_t vec = (size_t)nm * M;
    const size_t nPart = (size_t)nm * g.nTiles;

    double *TXl, *TYl, *dinv, *P, *Rv, *Z, *AP, *Pa, *Pb, *Vxl, *Vyl, *Sa, *Sb;
    double *part_rz, *part_rr, *part_pAp, *part_pm, *bb, *pin;
    int *done, *iters, *nts, *cg_fail, *counters;
    HM_CHECK(ctx->ws.get("sim.TXl", vec, &TXl));
    HM_CHECK(ctx->ws.get("sim.TYl", vec, &TYl));
    HM_CHECK(ctx->ws.get("sim.dinv", vec, &dinv));
    HM_CHECK(ctx->ws.get("sim.P", vec, &P));
    HM_CHECK(ctx->ws.get("sim.r", vec, &Rv));
    HM_CHECK(ctx->ws.get("sim.z", vec, &Z));
    HM_CHECK(ctx->ws.get("sim.Ap", vec, &AP));
    HM_CHECK(ctx->ws.get("sim.pa", vec, &Pa));
    HM_CHECK(ctx->ws.get("sim.pb", vec, &Pb));
    HM_CHECK(ctx->ws.get("sim.Vxl", vec, &Vxl));
    HM_CHECK(ctx->ws.get("sim.Vyl", vec, &Vyl));
    HM_CHECK(ctx->ws.get("sim.Sa", vec, &Sa));
    HM_CHECK(ctx->ws.get("sim.Sb", vec, &Sb));
    HM_CHECK(ctx->ws.get("sim.part_rz", 2 * nPart, &part_rz));
    HM_CHECK(ctx->ws.get("sim.part_rr", 2 * nPart, &part_rr));
    HM_CHECK(ctx->ws.get("sim.part_pAp", nPart, &part_pAp));
    HM_CHECK(ctx->ws.get("sim.part_pm", nPart, &part_pm));
    HM_CHECK(ctx->ws.get("sim.bb", (size_t)nm, &bb));
    HM_CHECK(ctx->ws.get("sim.pin", (size_t)nm, &pin));
    HM_CHECK(ctx->ws.get("sim.done", (size_t)nm, &done));
    HM_CHECK(ctx->ws.get("sim.iters", (size_t)nm, &iters));
    HM_CHECK(ctx->ws.get("sim.nts", (size_t)nm, &nts));
    HM_CHECK(ctx->ws.get("sim.cg_fail", (size_t)nm, &cg_fail));
    HM_CHECK(ctx->ws.get("sim.counters", (size_t)4, &counters));

    // member-offset views of the caller's arrays
    const double* K = d.K + (int64_t)m0 * d.K_member_stride;
    Wells w;
    w.n = d.n_wells;
    w.cell = d.well_cell + (int64_t)m0 * d.well_cell_member_stride;
    w.cell_ms = d.well_cell_member_stride;
    w.rate = d.well_rate + (int64_t)m0 * d.well_rate_member_stride;
    w.rate_ms = d.well_rate_member_stride;
    w.rate_ss = d.well_rate_step_stride;
    double* S_hist = d.S_hist ? d.S_hist + (int64_t)m0 * (d.n_steps + 1) * M : nullptr;
    double* obs = d.obs ? d.obs + (int64_t)m0 * d.n_steps * d.n_obs : nullptr;
    int32_t* substeps = d.substeps ? d.substeps + (int64_t)m0 * d.n_steps : nullptr;
    int32_t* cg_iters_out = d.cg_iters ? d.cg_iters + (int64_t)m0 * d.n_steps : nullptr;

    const double rtol = d.cg_rtol > 0 ? d.cg_rtol : 1e-12;
    const double tol2 = rtol * rtol;
    const int max_iter = d.cg_max_iter > 0 ? d.cg_max_iter : 100 * (d.Nx + d.Ny) + 200;

    const int grid = nm * g.nTiles;
    const size_t smem1 = (size_t)(g.R + 2) * g.Ny * sizeof(double);
    const size_t smem2 = 2 * smem1;
    if (smem2 > 48 * 1024) {
        HM_CUDA(cudaFuncSetAttribute(k_tpfa_setup, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    }
    if (smem1 > 48 * 1024) {
        HM_CUDA(cudaFuncSetAttribute(k_cg_init, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
        HM_CUDA(cudaFuncSetAttribute(k_cg_spmv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
        HM_CUDA(cudaFuncSetAttribute(k_flux_cfl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
        HM_CUDA(cudaFuncSetAttribute(k_sat_substep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    }
    const int copy_blocks = (int)((vec + 255) / 256);

    // initial state: S <- S0, P <- 0 (cold start of the first solve), flags
    k_copy_rows<<<copy_blocks, 256, 0, st>>>(nm, (int)M, d.S0 + (int64_t)m0 * d.S0_member_stride,
                                              d.S0_member_stride, Sa, M);
    HM_CUDA(cudaMemsetAsync(P, 0, vec * sizeof(double), st));
    HM_CUDA(cudaMemsetAsync(cg_fail, 0, nm * sizeof(int), st));
    if (S_hist) k_copy_rows<<<copy_blocks, 256, 0, st>>>(nm, (int)M, Sa, M, S_hist, (int64_t)(d.n_steps + 1) * M);
    ctx->sim_stats.kernel_launches += 1 + (S_hist ? 1 : 0);

    PhaseTimer timer(st);
    double* Scur = Sa;
    double* Snxt = Sb;
    int cg_batch = 32;
    for (int step = 0; step < d.n_steps; ++step) {
        timer.mark(0);
        k_tpfa_setup<<<grid, kThreads, smem2, st>>>(g, Scur, K, d.K_member_stride, d.K_comp_stride, TXl,
                                                     TYl, dinv, pin);
        timer.mark(1);
        HM_CUDA(cudaMemsetAsync(counters, 0, 4 * sizeof(int), st));
        k_cg_init<<<grid, kThreads, smem1, st>>>(g, w, step, P, TXl, TYl, dinv, pin, Rv, Z, part_rz,
                                                  part_rr, bb, done, iters, counters);
        ctx->sim_stats.kernel_launches += 2;
        int k = 0;
        bool all_done = false;
        while (k < max_iter && !all_done) {
            const int kend = std::min(max_iter, k + cg_batch);
            for (; k < kend; ++k) {
                const int cur = k & 1, nxt = cur ^ 1;
                double* Pin = cur ? Pb : Pa;
                double* Pout = cur ? Pa : Pb;
                k_cg_spmv<<<grid, kThreads, smem1, st>>>(g, k, tol2, Z, Pin, Pout, AP, TXl, TYl, pin,
                                                          part_rz + cur * nPart, part_rz + nxt * nPart,
                                                          part_rr + cur * nPart, part_pAp, bb, done,
                                                          iters, counters);
                k_cg_update<<<grid, kThreads, 0, st>>>(g, P, Rv, Z, Pout, AP, dinv, part_rz + cur * nPart,
                                                        part_pAp, part_rz + nxt * nPart,
                                                        part_rr + nxt * nPart, done);
            }
            HM_CUDA(cudaMemcpyAsync(ctx->h_pinned, counters, sizeof(int), cudaMemcpyDeviceToHost, st));
            HM_CUDA(cudaStreamSynchronize(st));
            all_done = ctx->h_pinned[0] >= nm;
        }
        if (!all_done) {
            // one more convergence test (the last update may have converged), then flag the rest
            const int cur = k & 1, nxt = cur ^ 1;
            k_cg_spmv<<<grid, kThreads, smem1, st>>>(g, k, tol2, Z, cur ? Pb : Pa, cur ? Pa : Pb, AP, TXl,
                                                      TYl, pin, part_rz + cur * nPart,
                                                      part_rz + nxt * nPart, part_rr + cur * nPart,
                                                      part_pAp, bb, done, iters, counters);
        }
        if (cg_iters_out)
            k_record_iters<<<(nm + 127) / 128, 128, 0, st>>>(nm, iters, cg_iters_out, d.n_steps, step);
        ctx->sim_stats.cg_iterations += k;
        ctx->sim_stats.kernel_launches += 2 * k;
        ctx->sim_stats.cg_kernel_launches += 2 * k + 1;

        timer.mark(2);
        k_flux_cfl<<<grid, kThreads, smem1, st>>>(g, w, step, P, TXl, TYl, d.por, Vxl, Vyl, part_pm);
        k_substep_count<<<(nm + 127) / 128, 128, 0, st>>>(g, nm, d.dt, part_pm, nts, counters, substeps,
                                                            d.n_steps, step);
        HM_CUDA(cudaMemcpyAsync(ctx->h_pinned, counters, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
        HM_CUDA(cudaStreamSynchronize(st));
        const int max_nts = ctx->h_pinned[1];
        ctx->sim_stats.kernel_launches += 2;

        timer.mark(3);
        for (int it = 0; it < max_nts; ++it) {
            k_sat_substep<<<grid, kThreads, smem1, st>>>(g, w, step, it, d.dt, nts, Scur, Snxt, Vxl, Vyl,
                                                          d.por);
            std::swap(Scur, Snxt);
        }
        ctx->sim_stats.sat_substeps += max_nts;
        ctx->sim_stats.kernel_launches += max_nts;
        ctx->sim_stats.sat_kernel_launches += max_nts;

        timer.mark(4);
        if (obs && d.n_obs > 0) {
            k_gather_obs<<<(nm * d.n_obs + 255) / 256, 256, 0, st>>>(nm, (int)M, d.n_obs, d.obs_cell, Scur,
                                                                       obs, d.n_steps, step);
            ctx->sim_stats.kernel_launches += 1;
        }
        if (S_hist) {
            k_copy_rows<<<copy_blocks, 256, 0, st>>>(nm, (int)M, Scur, M, S_hist + (int64_t)(step + 1) * M,
                                                      (int64_t)(d.n_steps + 1) * M);
            ctx->sim_stats.kernel_launches += 1;
        }
        // adapt the convergence-check cadence to what this step needed
        cg_batch = std::max(8, std::min(64, k / 6 + 4));
        if (!all_done) {
            k_mark_unconverged<<<(nm + 127) / 128, 128, 0, st>>>(nm, done, cg_fail);
        }
    }
    timer.mark(-1);
    k_copy_rows<<<copy_blocks, 256, 0, st>>>(nm, (int)M, Scur, M, d.S_last + (int64_t)m0 * M, M);
    if (d.P_last) k_copy_rows<<<copy_blocks, 256, 0, st>>>(nm, (int)M, P, M, d.P_last + (int64_t)m0 * M, M);
    if (d.status) k_member_status<<<nm, 256, 0, st>>>(nm, (int)M, Scur, cg_fail, d.status + m0);
    ctx->sim_stats.kernel_launches += 1 + (d.P_last ? 1 : 0) + (d.status ? 1 : 0);
    HM_CUDA(cudaStreamSynchronize(st));
    timer.finish(ctx->phase_ms);
    HM_CUDA(cudaGetLastError());
    return HM_OK;
}

int validate(const hm_sim_desc& d) {
    HM_REQUIRE(d.n_members > 0, "n_members > 0");
    HM_REQUIRE(d.Nx > 0 && d.Ny > 0, "grid size");
    HM_REQUIRE(d.Lx > 0 && d.Ly > 0, "domain size");
    HM_REQUIRE(d.vw > 0 && d.vo > 0 && d.swc >= 0 && d.sor >= 0 && d.swc + d.sor < 1, "fluid");
    HM_REQUIRE(d.K && d.S0 && d.S_last, "K, S0, S_last are required");
    HM_REQUIRE(d.n_wells >= 0 && d.n_wells <= kMaxWells, "0 <= n_wells <= 64");
    HM_REQUIRE(d.n_wells == 0 || (d.well_cell && d.well_rate), "well arrays");
    HM_REQUIRE(d.n_steps >= 0 && d.dt > 0, "dt, n_steps");
    HM_REQUIRE(d.n_obs == 0 || d.obs_cell, "obs_cell");
    HM_REQUIRE((size_t)(d.Ny) * 3 * sizeof(double) * 2 <= 200 * 1024, "Ny too large for the row tile");
    return HM_OK;
}

}  // namespace

extern "C" int hm_sim_batch(hm_ctx* ctx, const hm_sim_desc* desc) {
    HM_REQUIRE(ctx && desc, "null ctx/desc");
    HM_CHECK(validate(*desc));
    HM_CUDA(cudaSetDevice(ctx->device));
    ctx->sim_stats = hm_sim_stats{};
    for (double& v : ctx->phase_ms) v = 0.0;
    const int chunk = desc->chunk_members > 0 ? std::min(desc->chunk_members, desc->n_members)
                                              : desc->n_members;
    for (int m0 = 0; m0 < desc->n_members; m0 += chunk)
        HM_CHECK(sim_chunk(ctx, *desc, m0, std::min(chunk, desc->n_members - m0)));
    ctx->launches += ctx->sim_stats.kernel_launches;
    return HM_OK;
}

extern "C" int hm_sim_get_stats(hm_ctx* ctx, hm_sim_stats* out) {
    HM_REQUIRE(ctx && out, "null");
    *out = ctx->sim_stats;
    return HM_OK;
}

extern "C" int hm_sim_get_phase_ms(hm_ctx* ctx, double out[5]) {
    HM_REQUIRE(ctx && out, "null");
    for (int i = 0; i < 5; ++i) out[i] = ctx->phase_ms[i];
    return HM_OK;
}

// Host-buffer entry point: the call a non-CUDA host (the reference's Python, via
// ctypes) makes.  Stages every array through ctx-owned device memory.
extern "C" int hm_sim_batch_host(hm_ctx* ctx, const hm_sim_desc* hd) {
    HM_REQUIRE(ctx && hd, "null ctx/desc");
    HM_CHECK(validate(*hd));
    HM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    hm_sim_desc d = *hd;
    const int64_t M = (int64_t)d.Nx * d.Ny;
    const int nm = d.n_members;
    auto up = [&](const char* name, const void* src, size_t bytes, void** dst) -> int {
        HM_CHECK(ctx->ws.get(name, bytes, dst));
        HM_CUDA(cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, st));
        return HM_OK;
    };
    void* p = nullptr;
    const size_t ncomp = d.K_comp_stride ? 2 : 1;
    const size_t nK = d.K_member_stride ? (size_t)nm : 1;
    // host K must be densely packed per member: [member][comp][M]
    HM_REQUIRE(d.K_member_stride == 0 || d.K_member_stride == (int64_t)(ncomp * M), "host K must be dense");
    HM_REQUIRE(d.K_comp_stride == 0 || d.K_comp_stride == M, "host K must be dense");
    HM_CHECK(up("h.K", hd->K, nK * ncomp * M * sizeof(double), &p));
    d.K = (const double*)p;
    if (hd->por) {
        HM_CHECK(up("h.por", hd->por, M * sizeof(double), &p));
        d.por = (const double*)p;
    }
    if (d.n_wells > 0) {
        const size_t nwc = d.well_cell_member_stride ? (size_t)nm : 1;
        HM_REQUIRE(d.well_cell_member_stride == 0 || d.well_cell_member_stride == d.n_wells, "dense well_cell");
        HM_CHECK(up("h.wc", hd->well_cell, nwc * d.n_wells * sizeof(int32_t), &p));
        d.well_cell = (const int32_t*)p;
        const size_t nsteps_r = d.well_rate_step_stride ? (size_t)d.n_steps : 1;
        const size_t nwr = d.well_rate_member_stride ? (size_t)nm : 1;
        HM_REQUIRE(d.well_rate_step_stride == 0 || d.well_rate_step_stride == d.n_wells, "dense well_rate");
        HM_REQUIRE(d.well_rate_member_stride == 0 ||
                       d.well_rate_member_stride == (int64_t)(nsteps_r * d.n_wells),
                   "dense well_rate");
        HM_CHECK(up("h.wr", hd->well_rate, nwr * nsteps_r * d.n_wells * sizeof(double), &p));
        d.well_rate = (const double*)p;
    }
    HM_REQUIRE(d.S0_member_stride == 0 || d.S0_member_stride == M, "dense S0");
    HM_CHECK(up("h.S0", hd->S0, (d.S0_member_stride ? (size_t)nm : 1) * M * sizeof(double), &p));
    d.S0 = (const double*)p;
    if (d.n_obs > 0) {
        HM_CHECK(up("h.obs_cell", hd->obs_cell, d.n_obs * sizeof(int32_t), &p));
        d.obs_cell = (const int32_t*)p;
    }
    HM_CHECK(ctx->ws.get("h.S_last", (size_t)nm * M * sizeof(double), &p));
    d.S_last = (double*)p;
    if (hd->S_hist) {
        HM_CHECK(ctx->ws.get("h.S_hist", (size_t)nm * (d.n_steps + 1) * M * sizeof(double), &p));
        d.S_hist = (double*)p;
    }
    if (hd->obs) {
        HM_CHECK(ctx->ws.get("h.obs", (size_t)nm * d.n_steps * d.n_obs * sizeof(double), &p));
        d.obs = (double*)p;
    }
    if (hd->P_last) {
        HM_CHECK(ctx->ws.get("h.P_last", (size_t)nm * M * sizeof(double), &p));
        d.P_last = (double*)p;
    }
    if (hd->status) {
        HM_CHECK(ctx->ws.get("h.status", (size_t)nm * sizeof(int32_t), &p));
        d.status = (int32_t*)p;
    }
    if (hd->substeps) {
        HM_CHECK(ctx->ws.get("h.substeps", (size_t)nm * d.n_steps * sizeof(int32_t), &p));
        d.substeps = (int32_t*)p;
    }
    if (hd->cg_iters) {
        HM_CHECK(ctx->ws.get("h.cg_iters", (size_t)nm * d.n_steps * sizeof(int32_t), &p));
        d.cg_iters = (int32_t*)p;
    }
    HM_CHECK(hm_sim_batch(ctx, &d));
    auto down = [&](void* dst, const void* src, size_t bytes) -> int {
        HM_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
        return HM_OK;
    };
    HM_CHECK(down(hd->S_last, d.S_last, (size_t)nm * M * sizeof(double)));
    if (hd->S_hist) HM_CHECK(down(hd->S_hist, d.S_hist, (size_t)nm * (d.n_steps + 1) * M * sizeof(double)));
    if (hd->obs) HM_CHECK(down(hd->obs, d.obs, (size_t)nm * d.n_steps * d.n_obs * sizeof(double)));
    if (hd->P_last) HM_CHECK(down(hd->P_last, d.P_last, (size_t)nm * M * sizeof(double)));
    if (hd->status) HM_CHECK(down(hd->status, d.status, (size_t)nm * sizeof(int32_t)));
    if (hd->substeps) HM_CHECK(down(hd->substeps, d.substeps, (size_t)nm * d.n_steps * sizeof(int32_t)));
    if (hd->cg_iters) HM_CHECK(down(hd->cg_iters, d.cg_iters, (size_t)nm * d.n_steps * sizeof(int32_t)));
    HM_CUDA(cudaStreamSynchronize(st));
    return HM_OK;
}
