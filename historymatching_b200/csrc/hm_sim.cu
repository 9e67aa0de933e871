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
#include <cooperative_groups.h>

#include <cstdlib>
#include <type_traits>

#include "hm_ptx.cuh"
#include "hm_sim_common.cuh"

using namespace hmsim;

namespace hmsim {
int pressure_solve(hm_ctx* ctx, const Geo& g, const Wells& w, int step, int nm, const double* TXl,
                   const double* TYl, const double* dinv, const double* pin, double* P, double rtol,
                   int max_iter, int precond, int mg_switch_iters, int* done, int* iters, int* counters,
                   int* cg_batch, int* iters_used, bool* all_done_out);
int sim_small_supported(const hm_sim_desc& d);
int sim_small(hm_ctx* ctx, const hm_sim_desc& d, int m0, int nm);
bool transport_tb_supported(const hm_sim_desc& d);
int transport_tb(hm_ctx* ctx, const hm_sim_desc& d, const Fluid& fl, const Wells& w, int step, int nm, int max_nts,
                 const int* nts, double* Scur, double* Snxt, const double* Vxl, const double* Vyl, double** Sresult,
                 int* launches);
}

namespace {

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
            lx = 1.0 / (mt * perm_value(g, Kx[c]));
            ly = 1.0 / (mt * perm_value(g, Ky[c]));
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
            const double pv = perm_value(g, Kx[0]) + perm_value(g, Ky[0]);
            d += pv;
            pin[m] = pv;
        }
        TXl[off + c] = txl;
        TYl[off + c] = tyl;
        dinv[off + c] = 1.0 / d;
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
// HBM-bound streaming kernel: 32 B per cell (S in, S out, the two low-face fluxes).
// A thread owns CPT cells of the tile (linear index e = tid + j*kThreads, coalesced), keeps
// their saturation in registers, publishes the fractional flow fw(S) of tile + 2 halo rows in
// shared memory and then applies the upwind stencil.  Wells touch a handful of cells and are
// applied as a fix-up pass after the sweep (the source terms are additive in the update).
constexpr int kSatCPT = kTileCells / kThreads;

// The flux arrays carry a zero pad (Ny resp. 1 elements) behind the last member, and the low
// faces of row 0 / column 0 are zero by construction, so the HIGH faces of the last row /
// last column can be read as Vxl[c+Ny] / Vyl[c+1] without any boundary test.
template <bool HAS_POR, bool FULL>
__device__ __forceinline__ void sat_tile_body(const Geo& g, const Fluid& fl, int nInt, double hdt0, double dts,
                                              const double* __restrict__ sp, double* __restrict__ op,
                                              const double* __restrict__ vx, const double* __restrict__ vy,
                                              const double* __restrict__ pp, double* fwp) {
    const int Ny = g.Ny;
    double s[kSatCPT];
#pragma unroll
    for (int j = 0; j < kSatCPT; ++j)
        s[j] = (FULL || threadIdx.x + j * kThreads < nInt) ? sp[j * kThreads] : 0.0;
#pragma unroll
    for (int j = 0; j < kSatCPT; ++j)
        if (FULL || threadIdx.x + j * kThreads < nInt) fwp[j * kThreads] = frac_flow_fast(s[j], fl);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kSatCPT; ++j) {
        if (FULL || threadIdx.x + j * kThreads < nInt) {
            const double vxl = vx[j * kThreads];
            const double vyl = vy[j * kThreads];
            const double vxh = vx[j * kThreads + Ny];
            const double vyh = vy[j * kThreads + 1];
            double hdt = hdt0;
            if (HAS_POR) hdt = 0.5 * (dts / (g.h2 * pp[j * kThreads]));
            const double* f = fwp + j * kThreads;
            // B row: [x2(c-Ny), y2(c-1), diag, -y1(c+1), -x1(c+Ny)] scaled by dtx, with
            // max(v,0) = (v+|v|)/2 and min(v,0) = (v-|v|)/2 (exact in floating point)
            double acc = (hdt * (vxl + fabs(vxl))) * f[-Ny];
            acc = fma(hdt * (vyl + fabs(vyl)), f[-1], acc);
            const double dg = ((vyl - vyh) + (vxl - vxh)) - ((fabs(vyl) + fabs(vyh)) + (fabs(vxl) + fabs(vxh)));
            acc = fma(hdt * dg, f[0], acc);
            acc = fma(hdt * (fabs(vyh) - vyh), f[1], acc);
            acc = fma(hdt * (fabs(vxh) - vxh), f[Ny], acc);
            op[j * kThreads] = s[j] + acc;
        }
    }
}

template <bool HAS_POR>
__global__ void __launch_bounds__(kThreads, 6)
k_sat_substep(Geo g, Fluid fl, Wells w, int step, int it, double dt, const int* __restrict__ nts,
              const double* __restrict__ Sin, double* __restrict__ Sout,
              const double* __restrict__ Vxl, const double* __restrict__ Vyl,
              const double* __restrict__ por) {
    extern __shared__ double fws[];  // fractional flow of tile + halo rows
    __shared__ int wc[kMaxWells];
    __shared__ double wr[kMaxWells];
    const int m = blockIdx.x / g.nTiles, t = blockIdx.x % g.nTiles;
    const int r0 = t * g.R, r1 = min(r0 + g.R, g.Nx), rows = r1 - r0;
    const int Ny = g.Ny, nInt = rows * Ny;
    const int64_t base = (int64_t)m * g.M + (int64_t)r0 * Ny;  // first interior cell
    const int n = nts[m];
    // per-thread base pointers: element j of this thread sits at a compile-time offset j*kThreads
    const double* __restrict__ sp = Sin + base + threadIdx.x;
    double* __restrict__ op = Sout + base + threadIdx.x;
    if (it >= n) {  // this member needs fewer sub-steps: carry its state over
#pragma unroll
        for (int j = 0; j < kSatCPT; ++j)
            if (threadIdx.x + j * kThreads < nInt) op[j * kThreads] = sp[j * kThreads];
        return;
    }
    if (w.n > 0) load_wells(w, m, step, wc, wr);
    // halo rows (row r0-1 and row r1); outside the domain the value is never used (zero flux)
    for (int e = threadIdx.x; e < 2 * Ny; e += kThreads) {
        const bool lowh = e < Ny;
        const int col = lowh ? e : e - Ny;
        const int row = lowh ? r0 - 1 : r1;
        double f = 0.0;
        if (row >= 0 && row < g.Nx) f = frac_flow_fast(Sin[(int64_t)m * g.M + (int64_t)row * Ny + col], fl);
        fws[lowh ? col : Ny + nInt + col] = f;
    }
    const double dts = dt / (double)n;
    const double hdt0 = 0.5 * (dts / g.h2);
    const double* __restrict__ vx = Vxl + base + threadIdx.x;
    const double* __restrict__ vy = Vyl + base + threadIdx.x;
    const double* __restrict__ pp = HAS_POR ? por + (int64_t)r0 * Ny + threadIdx.x : nullptr;
    double* fwp = fws + Ny + threadIdx.x;
    if (nInt == kTileCells)
        sat_tile_body<HAS_POR, true>(g, fl, nInt, hdt0, dts, sp, op, vx, vy, pp, fwp);
    else
        sat_tile_body<HAS_POR, false>(g, fl, nInt, hdt0, dts, sp, op, vx, vy, pp, fwp);
    if (w.n == 0) return;
    __syncthreads();
    // well fix-up: S += dtx * (min(q,0) * fw(S) + max(q,0)) on the cells that hold wells
    for (int i = threadIdx.x; i < w.n; i += kThreads) {
        const int c = wc[i];
        const int e = c - r0 * Ny;
        if (e < 0 || e >= nInt) continue;
        bool first = true;
        for (int k = 0; k < i; ++k) first = first && (wc[k] != c);
        if (!first) continue;
        const double q = cell_source(c, w.n, wc, wr);
        double dtx = dts / g.h2;
        if (HAS_POR) dtx = dts / (g.h2 * por[c]);
        Sout[base + e] += fma(dtx * fmin(q, 0.0), fws[Ny + e], fmax(q, 0.0) * dtx);
    }
}

// ---- K4b: the whole sub-step loop of one time step in ONE launch (thread-block clusters) ------------
// The face fluxes are frozen during the Nts sub-steps of a time step.  The tiles of one member form
// a thread-block cluster (nTiles <= 16 CTAs); every CTA keeps the saturation and the five upwind
// stencil coefficients of its cells in REGISTERS for all Nts sub-steps, publishes fw(S) of its rows
// in its own shared memory and pushes its first / last row into the neighbouring CTAs' halo rows
// through distributed shared memory.  One cluster barrier per sub-step (the fw tiles are double
// buffered).  HBM is touched once per time step (S in, fluxes in, S out) instead of once per
// sub-step, and the per-sub-step FP64 work drops from ~42 to ~20 operations per cell because the
// coefficients are not rebuilt.
// Halo exchange between the CTAs of a cluster without cluster-scope fences: the sender writes each
// value with st.async (a remote shared-memory store that performs complete_tx on an mbarrier of
// the RECEIVING CTA); the receiver posts the expected byte count on that mbarrier and waits on its
// phase.  (A cluster barrier per sub-step costs a GPU-scope MEMBAR plus an L1 invalidate, CCTL.IVALL,
// on this architecture and dominated the sub-step.)
// ---- K4c: streaming sub-step with bulk-copy (TMA) staging -------------------------------------------------
// Same arithmetic and traffic as k_sat_substep, different data movement.  k_sat_substep is bound by memory
// latency (ncu at 512^2: long-scoreboard stalls, 48 % of the DRAM peak): a thread first waits for its S values,
// and issues the flux loads only after the fw barrier.  Here ONE thread issues three bulk copies
// (cp.async.bulk, completion on an mbarrier) for everything the tile needs - the S rows incl. the two halo rows,
// the x-fluxes of R+1 rows, the y-fluxes (+ one element) - ~61 KB in flight per CTA, three CTAs per SM; the
// CTA then turns S into fw(S) in place (own cells keep S in registers), and applies the stencil from shared
// memory.  Needs an even row length (16-byte granularity of the bulk copies).
constexpr int kStreamThreads = 512;

// CPT cells per thread: 4 (tiles of 2048 cells, ~61 KB, three CTAs per SM) or 8 (tiles of 4096 cells, ~110 KB, two CTAs per
// SM: the per-CTA prologue - index arithmetic, barrier set-up, halo rows - is amortised over twice the cells and
// the halo rows are 2 in 8 + 2 instead of 2 in 4 + 2 rows of S).  The Geo passed in carries the matching R / nTiles.
template <bool HAS_POR, int CPT>
__global__ void __launch_bounds__(kStreamThreads, CPT == 4 ? 3 : 2)
k_sat_stream(Geo g, Fluid fl, Wells w, int step, int it, double dt, const int* __restrict__ nts,
             const double* __restrict__ Sin, double* __restrict__ Sout, const double* __restrict__ Vxl,
             const double* __restrict__ Vyl, const double* __restrict__ por) {
    extern __shared__ __align__(16) double sms[];  // fw/S rows [(R+2) Ny], Vx rows [(R+1) Ny], Vy [R Ny + 2]
    __shared__ __align__(8) unsigned long long bar_store;
    __shared__ int wc[kMaxWells];
    __shared__ double wr[kMaxWells];
    constexpr int NT = kStreamThreads;
    const int m = blockIdx.x / g.nTiles, t = blockIdx.x % g.nTiles;
    const int r0 = t * g.R, r1 = min(r0 + g.R, g.Nx), rows = r1 - r0;
    const int Ny = g.Ny, nInt = rows * Ny;
    const int64_t mbase = (int64_t)m * g.M;
    const int64_t base = mbase + (int64_t)r0 * Ny;  // first interior cell
    const int n = nts[m];
    if (it >= n) {  // this member needs fewer sub-steps: carry its state over
        for (int e = threadIdx.x; e < nInt; e += NT) Sout[base + e] = Sin[base + e];
        return;
    }
    double* Ss = sms;
    double* Vxs = sms + (g.R + 2) * Ny;
    double* Vys = Vxs + (g.R + 1) * Ny;
    const uint32_t bar = smem_u32(&bar_store);
    if (threadIdx.x == 0) {  // the copies start at once; the other threads meet the barrier after the CTA barrier below
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const int rowLo = max(r0 - 1, 0), rowHi = min(r1 + 1, g.Nx);
        const uint32_t bS = (uint32_t)(rowHi - rowLo) * Ny * 8u, bX = (uint32_t)(rows + 1) * Ny * 8u,
                       bY = (uint32_t)(nInt + 2) * 8u;
        mbar_expect_tx(bar, (int)(bS + bX + bY));
        bulk_g2s(smem_u32(Ss + (rowLo - (r0 - 1)) * Ny), Sin + mbase + (int64_t)rowLo * Ny, bS, bar);
        bulk_g2s(smem_u32(Vxs), Vxl + base, bX, bar);
        bulk_g2s(smem_u32(Vys), Vyl + base, bY, bar);
    }
    __syncwarp();  // lanes 1-31 of warp 0 poll the barrier below: not before lane 0 has initialised it
    if (w.n > 0) load_wells(w, m, step, wc, wr);
    // halo rows outside the domain: zero (the matching flux is zero, the value only has to be finite); the bulk
    // copies do not touch these rows
    if (t == 0)
        for (int e = threadIdx.x; e < Ny; e += NT) Ss[e] = 0.0;
    if (r1 == g.Nx)
        for (int e = threadIdx.x; e < Ny; e += NT) Ss[(rows + 1) * Ny + e] = 0.0;
    const double dts = dt / (double)n;
    const double hdt0 = 0.5 * (dts / g.h2);
    // one warp polls the mbarrier (512 spinning threads cost ~60 % of the issue slots, ncu), the others sleep in
    // the CTA barrier; the second wait returns at once and is each thread's own acquire of the copied data
    if (threadIdx.x < 32) mbar_wait(bar, 0);
    __syncthreads();
    mbar_wait(bar, 0);
    // S -> fw(S) in place; a thread keeps the saturation of its own cells
    double s[CPT];
    double* fwp = Ss + Ny + threadIdx.x;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        s[j] = 0.0;
        if (threadIdx.x + j * NT < nInt) {
            s[j] = fwp[j * NT];
            fwp[j * NT] = frac_flow_fast(s[j], fl);
        }
    }
    for (int e = threadIdx.x; e < 2 * Ny; e += NT) {
        const bool lowh = e < Ny;
        const int col = lowh ? e : e - Ny;
        if (lowh ? t > 0 : r1 < g.Nx) {
            double* q = Ss + (lowh ? col : (rows + 1) * Ny + col);
            *q = frac_flow_fast(*q, fl);
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const int e = threadIdx.x + j * NT;
        if (e < nInt) {
            const double vxl = Vxs[e], vyl = Vys[e], vxh = Vxs[e + Ny], vyh = Vys[e + 1];
            double hdt = hdt0;
            if (HAS_POR) hdt = 0.5 * (dts / (g.h2 * por[(int64_t)r0 * Ny + e]));
            const double* f = fwp + j * NT;
            // same expression tree as sat_tile_body
            double acc = (hdt * (vxl + fabs(vxl))) * f[-Ny];
            acc = fma(hdt * (vyl + fabs(vyl)), f[-1], acc);
            const double dg = ((vyl - vyh) + (vxl - vxh)) - ((fabs(vyl) + fabs(vyh)) + (fabs(vxl) + fabs(vxh)));
            acc = fma(hdt * dg, f[0], acc);
            acc = fma(hdt * (fabs(vyh) - vyh), f[1], acc);
            acc = fma(hdt * (fabs(vxh) - vxh), f[Ny], acc);
            Sout[base + e] = s[j] + acc;
        }
    }
    if (w.n == 0) return;
    __syncthreads();
    // well fix-up: S += dtx * (min(q,0) * fw(S) + max(q,0)) on the cells that hold wells
    for (int i = threadIdx.x; i < w.n; i += NT) {
        const int c = wc[i];
        const int e = c - r0 * Ny;
        if (e < 0 || e >= nInt) continue;
        bool first = true;
        for (int k = 0; k < i; ++k) first = first && (wc[k] != c);
        if (!first) continue;
        const double q = cell_source(c, w.n, wc, wr);
        double dtx = dts / g.h2;
        if (HAS_POR) dtx = dts / (g.h2 * por[c]);
        Sout[base + e] += fma(dtx * fmin(q, 0.0), Ss[Ny + e], fmax(q, 0.0) * dtx);
    }
}

template <bool HAS_POR, int NT, int CPT, int NY>
__global__ void __launch_bounds__(NT, NT <= 512 ? 2 : 1)
k_sat_cluster(Geo g, Fluid fl, Wells w, int step, double dt, const int* __restrict__ nts,
              const double* __restrict__ Sin, double* __restrict__ Sout, const double* __restrict__ Vxl,
              const double* __restrict__ Vyl, const double* __restrict__ por) {
    namespace cg = cooperative_groups;
    extern __shared__ double smc[];  // fw[2][(R+2)*Ny] (halo row, R tile rows, halo row), src[R*Ny]
    __shared__ int wc[kMaxWells];
    __shared__ double wr[kMaxWells];
    constexpr int kCells = NT * CPT;  // cells of a full tile (the Geo passed in carries the matching R / nTiles)
    cg::cluster_group cluster = cg::this_cluster();
    const int m = blockIdx.x / g.nTiles, t = blockIdx.x % g.nTiles;  // t == rank in the cluster
    const int r0 = t * g.R, r1 = min(r0 + g.R, g.Nx), rows = r1 - r0;
    const int Ny = NY ? NY : g.Ny;  // NY != 0: row length known at compile time (immediate shared-memory offsets)
    const int nInt = rows * Ny, fwLen = (g.R + 2) * Ny;
    const int64_t base = (int64_t)m * g.M + (int64_t)r0 * Ny;
    const int n = nts[m];
    double* const fwb0 = smc;
    double* const fwb1 = smc + fwLen;
    double* srcs = smc + 2 * fwLen;
    load_wells(w, m, step, wc, wr);
    for (int e = threadIdx.x; e < 2 * fwLen; e += NT) smc[e] = 0.0;  // halo rows of boundary tiles stay zero
    for (int e = threadIdx.x; e < nInt; e += NT) srcs[e] = 0.0;
    __syncthreads();
    const double dts = n > 0 ? dt / (double)n : 0.0;
    const double dtx0 = dts / g.h2;
    for (int i = threadIdx.x; i < w.n; i += NT) {  // wells of this tile: signed source * dtx
        const int c = wc[i], e = c - r0 * Ny;
        if (e < 0 || e >= nInt) continue;
        bool first = true;
        for (int k = 0; k < i; ++k) first = first && (wc[k] != c);
        if (!first) continue;
        double dtx = dtx0;
        if (HAS_POR) dtx = dts / (g.h2 * por[c]);
        srcs[e] = cell_source(c, w.n, wc, wr) * dtx;
    }
    __syncthreads();

    // Cell -> thread map.  Fast path (full tile, rows are whole warps, NT a multiple of Ny): thread (q, col)
    // owns the CPT consecutive grid rows q*CPT .. q*CPT+CPT-1 at column col, so that the x-neighbours inside the
    // row group are the thread's own registers (2 + 2 CPT shared-memory loads per sub-step instead of 4 CPT).
    // General path: cell e = tid + j*NT.
    const bool full = nInt == kCells;
    const bool fast = full && NT % Ny == 0 && Ny % 32 == 0;
    const int q = fast ? threadIdx.x / Ny : 0, col = fast ? threadIdx.x - q * Ny : 0;
    const int e0 = fast ? q * CPT * Ny + col : threadIdx.x, es = fast ? Ny : NT;
    double s[CPT], aW[CPT], aS[CPT], aN[CPT], aE[CPT], dg[CPT], sr[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const int e = e0 + j * es;
        s[j] = aW[j] = aS[j] = aN[j] = aE[j] = dg[j] = sr[j] = 0.0;
        if (e < nInt) {
            const int64_t gc = base + e;
            s[j] = Sin[gc];
            const double vxl = Vxl[gc], vyl = Vyl[gc], vxh = Vxl[gc + Ny], vyh = Vyl[gc + 1];
            double hdt = 0.5 * dtx0;
            if (HAS_POR) hdt = 0.5 * (dts / (g.h2 * por[(int64_t)r0 * Ny + e]));
            // max(v,0) = (v+|v|)/2, min(v,0) = (v-|v|)/2 (exact)
            aW[j] = hdt * (vxl + fabs(vxl));
            aS[j] = hdt * (vyl + fabs(vyl));
            aN[j] = hdt * (fabs(vyh) - vyh);
            aE[j] = hdt * (fabs(vxh) - vxh);
            const double qs = srcs[e];
            dg[j] = hdt * (((vyl - vyh) + (vxl - vxh)) - ((fabs(vyl) + fabs(vyh)) + (fabs(vxl) + fabs(vxh)))) +
                    fmin(qs, 0.0);
            sr[j] = fmax(qs, 0.0);
        }
    }
    // remote halo rows: my first row is the high halo of tile t-1, my last row the low halo of tile t+1
    __shared__ __align__(8) unsigned long long bars[2];
    const uint32_t bar0 = smem_u32(&bars[0]), bar1 = smem_u32(&bars[1]);
    const bool hasUp = t > 0, hasDn = t < g.nTiles - 1;
    const int nNbr = (hasUp ? 1 : 0) + (hasDn ? 1 : 0);
    if (threadIdx.x == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (fast && nNbr) {  // fast path: first expectation of either barrier, re-posted by the waiters in the loop
            mbar_expect_tx(bar0, nNbr * Ny * 8);
            mbar_expect_tx(bar1, nNbr * Ny * 8);
        }
    }
    // shared::cluster addresses of the neighbours' halo rows and barriers, per buffer parity
    uint32_t up0 = 0, up1 = 0, dn0 = 0, dn1 = 0, upb0 = 0, upb1 = 0, dnb0 = 0, dnb1 = 0;
    if (hasUp) {
        up0 = map_to_cta(smem_u32(fwb0 + (g.R + 1) * Ny), t - 1);
        up1 = map_to_cta(smem_u32(fwb1 + (g.R + 1) * Ny), t - 1);
        upb0 = map_to_cta(bar0, t - 1);
        upb1 = map_to_cta(bar1, t - 1);
    }
    if (hasDn) {
        dn0 = map_to_cta(smem_u32(fwb0), t + 1);
        dn1 = map_to_cta(smem_u32(fwb1), t + 1);
        dnb0 = map_to_cta(bar0, t + 1);
        dnb1 = map_to_cta(bar1, t + 1);
    }
    // Only the first and last row of a tile depend on the neighbours' data.  Per sub-step: publish fw
    // (own tile, and the two edge rows into the neighbours' halo rows), update the inner rows after a
    // CTA-local barrier, wait for the neighbours' rows, update the two edge rows.
    bool edge[CPT], sendUp[CPT], sendDn[CPT];
    bool anyEdge = false;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const int e = threadIdx.x + j * NT;
        const bool in = e < nInt;
        edge[j] = in && (e < Ny || e >= nInt - Ny);
        sendUp[j] = in && e < Ny && hasUp;
        sendDn[j] = in && e >= nInt - Ny && hasDn;
        anyEdge = anyEdge || edge[j];
    }
    cluster.sync();  // tiles zeroed and mbarriers initialised in every CTA before remote traffic starts
    if (fast) {
        // No per-cell predicates in the loop.  The first tile row is cell 0 of the threads of row group 0, the
        // last tile row cell CPT-1 of the last row group: exactly the threads that read the halo rows the
        // neighbours fill, so "readers of a halo row == senders of the matching edge row" and a neighbour can
        // never run more than one sub-step ahead of any reader (the fw tiles are double buffered).
        const bool sUp = hasUp && q == 0, sDn = hasDn && q == NT / Ny - 1;
        const bool needWait = sUp || sDn;
        const uint32_t colb = 8u * (uint32_t)col;
        const bool unit = fl.inv_range == 1.0 && fl.swc_ir == 0.0 && fl.mr == 1.0;
        // Transaction accounting of the halo mbarriers: every phase expects nNbr * Ny * 8 bytes.  The expectation of
        // a barrier's NEXT phase is posted by one of the threads that has just seen the current phase complete (the
        // edge row groups are the only waiters), so the 24 interior warps carry no barrier bookkeeping in the loop.
        // (A neighbour's bytes may arrive before the expectation is posted: the transaction count goes negative
        // for a moment, the phase cannot complete before the poster's arrival.)
        const bool poster = needWait && (int)threadIdx.x == (hasUp ? 0 : NT - Ny);
        const int haloBytes = nNbr * Ny * 8;
        auto substep = [&](auto unit_tag, double* __restrict__ fw, uint32_t mybar, uint32_t upA, uint32_t upB,
                           uint32_t dnA, uint32_t dnB, int parity) {
            constexpr bool U = decltype(unit_tag)::value;
            double f[CPT], ss[CPT];
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                f[j] = frac_flow_loop<U>(s[j], fl);
                fw[j * Ny] = f[j];
                ss[j] = s[j] + sr[j];  // injector source: beside the fw chain, not behind the update chain
            }
            st_async_f64_if(sUp, upA + colb, f[0], upB);
            st_async_f64_if(sDn, dnA + colb, f[CPT - 1], dnB);
            double acc[CPT];  // the terms that only need this thread's registers, before the barrier
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                acc[j] = fma(dg[j], f[j], ss[j]);
                if (j > 0) acc[j] = fma(aW[j], f[j - 1], acc[j]);
                if (j < CPT - 1) acc[j] = fma(aE[j], f[j + 1], acc[j]);
            }
            __syncthreads();
            if (needWait) {
                mbar_wait(mybar, parity);
                if (poster) mbar_expect_tx(mybar, haloBytes);
            }
            double fS[CPT], fN[CPT];
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                fS[j] = fw[j * Ny - 1];
                fN[j] = fw[j * Ny + 1];
            }
            const double fWest = fw[-Ny], fEast = fw[CPT * Ny];
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                double a2 = fma(aS[j], fS[j], acc[j]);
                a2 = fma(aN[j], fN[j], a2);
                if (j == 0) a2 = fma(aW[j], fWest, a2);
                if (j == CPT - 1) a2 = fma(aE[j], fEast, a2);
                s[j] = a2;
            }
        };
        double* const fwa = fwb0 + Ny + e0;
        double* const fwb = fwb1 + Ny + e0;
        auto run = [&](auto unit_tag) {
            int sub = 0;
            for (; sub + 1 < n; sub += 2) {
                const int par = (sub >> 1) & 1;
                substep(unit_tag, fwa, bar0, up0, upb0, dn0, dnb0, par);
                substep(unit_tag, fwb, bar1, up1, upb1, dn1, dnb1, par);
            }
            if (sub < n) substep(unit_tag, fwa, bar0, up0, upb0, dn0, dnb0, (sub >> 1) & 1);
        };
        if (unit)
            run(std::true_type{});
        else
            run(std::false_type{});
    } else
    for (int sub = 0; sub < n; ++sub) {
        const bool odd = sub & 1;
        double* fw = (odd ? fwb1 : fwb0) + Ny + threadIdx.x;
        const uint32_t up = (odd ? up1 : up0) + 8u * threadIdx.x, upb = odd ? upb1 : upb0;
        const uint32_t dn = (odd ? dn1 : dn0) + 8u * (threadIdx.x - (nInt - Ny)), dnb = odd ? dnb1 : dnb0;
        const uint32_t mybar = odd ? bar1 : bar0;
        if (threadIdx.x == 0 && nNbr) mbar_expect_tx(mybar, nNbr * Ny * 8);
        double f[CPT];
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            f[j] = frac_flow_fast(s[j], fl);
            if (full || threadIdx.x + j * NT < nInt) fw[j * NT] = f[j];
            if (sendUp[j]) st_async_f64(up + 8u * (j * NT), f[j], upb);
            if (sendDn[j]) st_async_f64(dn + 8u * (j * NT), f[j], dnb);
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            if ((full || threadIdx.x + j * NT < nInt) && !edge[j]) {
                const double* fp = fw + j * NT;
                double acc = aW[j] * fp[-Ny];
                acc = fma(aS[j], fp[-1], acc);
                acc = fma(dg[j], f[j], acc);
                acc = fma(aN[j], fp[1], acc);
                acc = fma(aE[j], fp[Ny], acc);
                s[j] += acc + sr[j];
            }
        }
        if (anyEdge) {
            if (nNbr) mbar_wait(mybar, (sub >> 1) & 1);
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                if (edge[j]) {
                    const double* fp = fw + j * NT;
                    double acc = aW[j] * fp[-Ny];
                    acc = fma(aS[j], fp[-1], acc);
                    acc = fma(dg[j], f[j], acc);
                    acc = fma(aN[j], fp[1], acc);
                    acc = fma(aE[j], fp[Ny], acc);
                    s[j] += acc + sr[j];
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const int e = e0 + j * es;
        if (e < nInt) Sout[base + e] = s[j];
    }
    cluster.sync();  // no CTA may exit while a neighbour can still write into its shared memory
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

// warm start of the next pressure solve: linear extrapolation in time, P <- 2 P - Pprev, Pprev <- P
__global__ void k_extrapolate(int64_t n, double* __restrict__ P, double* __restrict__ Pprev) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double p = P[i];
    P[i] = 2.0 * p - Pprev[i];
    Pprev[i] = p;
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
    if (sim_small_supported(d)) return sim_small(ctx, d, m0, nm);  // whole simulator in one kernel (hm_small.cu)
    cudaStream_t st = ctx->stream;
    Geo g;
    g.Nx = d.Nx;
    g.Ny = d.Ny;
    g.M = d.Nx * d.Ny;
    g.R = std::max(1, std::min(d.Nx, kTileCells / d.Ny));
    if (g.R < d.Nx && g.R > 1) g.R &= ~1;  // even tile height: 2x2 multigrid aggregates never straddle tiles
    g.nTiles = (d.Nx + g.R - 1) / g.R;
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
    Fluid fl;
    fl.inv_range = 1.0 / (1.0 - d.swc - d.sor);
    fl.swc_ir = d.swc * fl.inv_range;
    fl.mr = d.vw / d.vo;
    const int64_t M = g.M;
    const size_t vec = (size_t)nm * M;
    const size_t nPart = (size_t)nm * g.nTiles;

    double *TXl, *TYl, *dinv, *P, *Vxl, *Vyl, *Sa, *Sb;
    double *part_pm, *pin;
    int *done, *iters, *nts, *cg_fail, *counters;
    HM_CHECK(ctx->ws.get("sim.TXl", vec + (size_t)d.Ny, &TXl));  // + zero pads: high faces read unguarded
    HM_CHECK(ctx->ws.get("sim.TYl", vec + 1, &TYl));
    HM_CHECK(ctx->ws.get("sim.dinv", vec, &dinv));
    HM_CHECK(ctx->ws.get("sim.P", vec, &P));
    double* Pprev;
    HM_CHECK(ctx->ws.get("sim.Pprev", vec, &Pprev));
    HM_CHECK(ctx->ws.get("sim.Vxl", vec + (size_t)d.Ny, &Vxl));  // + zero pad, see sat_tile_body
    HM_CHECK(ctx->ws.get("sim.Vyl", vec + 2, &Vyl));  // + 2: k_sat_stream copies nInt + 2 elements
    HM_CHECK(ctx->ws.get("sim.Sa", vec, &Sa));
    HM_CHECK(ctx->ws.get("sim.Sb", vec, &Sb));
    HM_CHECK(ctx->ws.get("sim.part_pm", nPart, &part_pm));
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
    const int nHist = hist_rows(d.n_steps, d.hist_stride);
    double* S_hist = d.S_hist ? d.S_hist + (int64_t)m0 * nHist * M : nullptr;
    double* obs = d.obs ? d.obs + (int64_t)m0 * d.n_steps * d.n_obs : nullptr;
    int32_t* substeps = d.substeps ? d.substeps + (int64_t)m0 * d.n_steps : nullptr;
    int32_t* cg_iters_out = d.cg_iters ? d.cg_iters + (int64_t)m0 * d.n_steps : nullptr;

    const double rtol = d.cg_rtol > 0 ? d.cg_rtol : 1e-12;
    const int max_iter = d.cg_max_iter > 0 ? d.cg_max_iter : 100 * (d.Nx + d.Ny) + 200;

    const int grid = nm * g.nTiles;
    const size_t smem1 = (size_t)(g.R + 2) * g.Ny * sizeof(double);
    const size_t smem2 = 2 * smem1;
    if (smem2 > 48 * 1024) {
        HM_CUDA(cudaFuncSetAttribute(k_tpfa_setup, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    }
    if (smem1 > 48 * 1024) {
        HM_CUDA(cudaFuncSetAttribute(k_flux_cfl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
        HM_CUDA(cudaFuncSetAttribute(k_sat_substep<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
        HM_CUDA(cudaFuncSetAttribute(k_sat_substep<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    }
    const int copy_blocks = (int)((vec + 255) / 256);
    // transport: all sub-steps of a time step in one cluster launch when the member's tiles fit a cluster
    // (sat_block != 1).  Tile = 2048 cells = 1024 threads x 2 cells; measured alternatives at 128^2 x 1024 members:
    // 512 threads x 4 cells the same speed, 4096-cell tiles (512 x 8, clusters of 4) spill and are 25 % slower.
    // Row length 128 / 64 is a compile-time constant (immediate shared-memory offsets), other lengths are runtime.
    // sat_block 4: tiles of 1024 cells, 512 threads, two CTAs (of different members) per SM: the per-sub-step barriers
    // of the two CTAs are decoupled, so one fills the other's pipeline bubbles.
    Geo gc = g;
    const bool half_tiles = d.sat_block == 4 && d.Ny <= 512 && (1024 / d.Ny) >= 2 && d.Nx % (1024 / d.Ny) == 0;
    if (half_tiles) {
        gc.R = 1024 / d.Ny;
        gc.nTiles = d.Nx / gc.R;
    }
    // temporally blocked kernel (hm_transport.cu): the default where the grid qualifies
    const bool use_tb = (d.sat_block == 0 || d.sat_block == 7) && transport_tb_supported(d);
    bool use_cluster = !use_tb && d.sat_block != 1 && d.sat_block != 5 && d.sat_block != 6 && gc.nTiles <= 16;
    const int cluster_threads = half_tiles ? 512 : 1024;
    const size_t smem_cluster = ((size_t)2 * (gc.R + 2) * d.Ny + (size_t)gc.R * d.Ny) * sizeof(double);
    using cluster_fn = void (*)(Geo, Fluid, Wells, int, double, const int*, const double*, double*, const double*,
                                const double*, const double*);
    cluster_fn cluster_kernel = d.por ? k_sat_cluster<true, 1024, 2, 0> : k_sat_cluster<false, 1024, 2, 0>;
    if (d.Ny == 128) cluster_kernel = d.por ? k_sat_cluster<true, 1024, 2, 128> : k_sat_cluster<false, 1024, 2, 128>;
    if (d.Ny == 64) cluster_kernel = d.por ? k_sat_cluster<true, 1024, 2, 64> : k_sat_cluster<false, 1024, 2, 64>;
    if (half_tiles) {
        cluster_kernel = d.por ? k_sat_cluster<true, 512, 2, 0> : k_sat_cluster<false, 512, 2, 0>;
        if (d.Ny == 128) cluster_kernel = d.por ? k_sat_cluster<true, 512, 2, 128> : k_sat_cluster<false, 512, 2, 128>;
        if (d.Ny == 64) cluster_kernel = d.por ? k_sat_cluster<true, 512, 2, 64> : k_sat_cluster<false, 512, 2, 64>;
    }
    cudaLaunchConfig_t cluster_cfg{};
    cudaLaunchAttribute cluster_attr[1];
    if (use_cluster) {
        HM_CUDA(cudaFuncSetAttribute(cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cluster));
        if (gc.nTiles > 8) HM_CUDA(cudaFuncSetAttribute(cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cluster_cfg.gridDim = dim3((unsigned)(nm * gc.nTiles));
        cluster_cfg.blockDim = dim3(cluster_threads);
        cluster_cfg.dynamicSmemBytes = smem_cluster;
        cluster_cfg.stream = st;
        cluster_attr[0].id = cudaLaunchAttributeClusterDimension;
        cluster_attr[0].val.clusterDim.x = (unsigned)gc.nTiles;
        cluster_attr[0].val.clusterDim.y = 1;
        cluster_attr[0].val.clusterDim.z = 1;
        cluster_cfg.attrs = cluster_attr;
        cluster_cfg.numAttrs = 1;
        // how many CTAs of this kernel the GPU holds at once (clusters must fit a GPC).  No resident cluster (a MIG
        // slice, a part with fused-off SMs): the streaming kernel runs the same grid.
        int nc = 0;
        if (cudaOccupancyMaxActiveClusters(&nc, cluster_kernel, &cluster_cfg) != cudaSuccess) nc = 0;
        (void)cudaGetLastError();
        if (getenv("HM_DEBUG"))
            fprintf(stderr, "[hm] k_sat_cluster: cluster of %d CTAs x %d threads, %zu B smem: max active clusters %d "
                    "(%d CTAs on %d SMs)\n", gc.nTiles, cluster_threads, smem_cluster, nc, nc * gc.nTiles, ctx->sm_count);
        if (nc > 0)
            ctx->sim_stats.sat_resident_ctas = (int64_t)nc * gc.nTiles;
        else
            use_cluster = false;
    }

    // streaming transport (grids whose tiles do not fit a cluster, or sat_block 1): bulk-copy staged kernel when the
    // row length is even (16-byte copies), sat_block 5 forces the plain-load kernel
    auto stream_geo = [&](int tile_cells) {
        Geo gs = g;
        gs.R = std::max(1, std::min(d.Nx, tile_cells / d.Ny));
        if (gs.R < d.Nx && gs.R > 1) gs.R &= ~1;
        gs.nTiles = (d.Nx + gs.R - 1) / gs.R;
        return gs;
    };
    auto stream_smem = [&](const Geo& gs) { return ((size_t)(3 * gs.R + 3) * d.Ny + 2) * sizeof(double); };
    // 4096-cell tiles (8 cells per thread, 2 CTAs per SM) where they fit; sat_block 6 forces the 2048-cell tiles
    Geo gs = stream_geo(8 * kStreamThreads);
    int stream_cpt = 8;
    if (d.sat_block == 6 || stream_smem(gs) > 112 * 1024 || gs.nTiles == 1) {
        gs = stream_geo(4 * kStreamThreads);
        stream_cpt = 4;
    }
    const size_t smem_stream = stream_smem(gs);
    const bool use_stream_tma = d.sat_block != 5 && d.Ny % 2 == 0 && smem_stream <= 112 * 1024 &&
                                (int64_t)gs.R * d.Ny <= (int64_t)stream_cpt * kStreamThreads;
    const int grid_stream = nm * gs.nTiles;
    if (use_stream_tma) {
        HM_CUDA(cudaFuncSetAttribute(k_sat_stream<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_stream));
        HM_CUDA(cudaFuncSetAttribute(k_sat_stream<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_stream));
        HM_CUDA(cudaFuncSetAttribute(k_sat_stream<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_stream));
        HM_CUDA(cudaFuncSetAttribute(k_sat_stream<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_stream));
    }

    // initial state: S <- S0, P <- 0 (cold start of the first solve), flags
    k_copy_rows<<<copy_blocks, 256, 0, st>>>(nm, (int)M, d.S0 + (int64_t)m0 * d.S0_member_stride,
                                              d.S0_member_stride, Sa, M);
    HM_CUDA(cudaMemsetAsync(P, 0, vec * sizeof(double), st));
    HM_CUDA(cudaMemsetAsync(TXl + vec, 0, (size_t)d.Ny * sizeof(double), st));
    HM_CUDA(cudaMemsetAsync(TYl + vec, 0, sizeof(double), st));
    HM_CUDA(cudaMemsetAsync(Vxl + vec, 0, (size_t)d.Ny * sizeof(double), st));
    HM_CUDA(cudaMemsetAsync(Vyl + vec, 0, 2 * sizeof(double), st));
    HM_CUDA(cudaMemsetAsync(cg_fail, 0, nm * sizeof(int), st));
    if (S_hist) k_copy_rows<<<copy_blocks, 256, 0, st>>>(nm, (int)M, Sa, M, S_hist, (int64_t)nHist * M);
    ctx->sim_stats.kernel_launches += 1 + (S_hist ? 1 : 0);

    PhaseTimer timer(st);
    double* Scur = Sa;
    double* Snxt = Sb;
    int cg_batch = d.precond == 1 ? 32 : 8;
    for (int step = 0; step < d.n_steps; ++step) {
        timer.mark(0);
        k_tpfa_setup<<<grid, kThreads, smem2, st>>>(g, Scur, K, d.K_member_stride, d.K_comp_stride, TXl,
                                                     TYl, dinv, pin);
        timer.mark(1);
        HM_CUDA(cudaMemsetAsync(counters, 0, 4 * sizeof(int), st));
        if (step >= 2 && d.warm_start != 1) {  // P holds P_{k-1}, Pprev holds P_{k-2}
            k_extrapolate<<<copy_blocks, 256, 0, st>>>((int64_t)vec, P, Pprev);
            ctx->sim_stats.kernel_launches += 1;
        } else if (step == 1) {
            HM_CUDA(cudaMemcpyAsync(Pprev, P, vec * sizeof(double), cudaMemcpyDeviceToDevice, st));
        }
        int k = 0;
        bool all_done = false;
        HM_CHECK(pressure_solve(ctx, g, w, step, nm, TXl, TYl, dinv, pin, P, rtol, max_iter, d.precond, d.mg_switch_iters,
                                done, iters, counters, &cg_batch, &k, &all_done));
        ctx->sim_stats.kernel_launches += 1;
        if (cg_iters_out)
            k_record_iters<<<(nm + 127) / 128, 128, 0, st>>>(nm, iters, cg_iters_out, d.n_steps, step);

        timer.mark(2);
        k_flux_cfl<<<grid, kThreads, smem1, st>>>(g, w, step, P, TXl, TYl, d.por, Vxl, Vyl, part_pm);
        k_substep_count<<<(nm + 127) / 128, 128, 0, st>>>(g, nm, d.dt, part_pm, nts, counters, substeps,
                                                            d.n_steps, step);
        HM_CUDA(cudaMemcpyAsync(ctx->h_pinned, counters, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
        HM_CUDA(cudaStreamSynchronize(st));
        const int max_nts = ctx->h_pinned[1];
        ctx->sim_stats.kernel_launches += 2;

        timer.mark(3);
        int sat_launches = 0;
        if (use_tb) {
            double* Sres = nullptr;
            HM_CHECK(transport_tb(ctx, d, fl, w, step, nm, max_nts, nts, Scur, Snxt, Vxl, Vyl, &Sres, &sat_launches));
            if (Sres != Scur) std::swap(Scur, Snxt);
        } else if (use_cluster) {
            const double* porp = d.por;
            HM_CUDA(cudaLaunchKernelEx(&cluster_cfg, cluster_kernel, gc, fl, w, step, d.dt, (const int*)nts,
                                       (const double*)Scur, Snxt, (const double*)Vxl, (const double*)Vyl, porp));
            std::swap(Scur, Snxt);
            sat_launches = 1;
        } else {
            for (int it = 0; it < max_nts; ++it, ++sat_launches) {
                if (use_stream_tma) {
                    auto ks = d.por ? (stream_cpt == 8 ? k_sat_stream<true, 8> : k_sat_stream<true, 4>)
                                    : (stream_cpt == 8 ? k_sat_stream<false, 8> : k_sat_stream<false, 4>);
                    ks<<<grid_stream, kStreamThreads, smem_stream, st>>>(gs, fl, w, step, it, d.dt, nts, Scur, Snxt, Vxl, Vyl,
                                                                        d.por);
                } else if (d.por)
                    k_sat_substep<true><<<grid, kThreads, smem1, st>>>(g, fl, w, step, it, d.dt, nts, Scur, Snxt,
                                                                        Vxl, Vyl, d.por);
                else
                    k_sat_substep<false><<<grid, kThreads, smem1, st>>>(g, fl, w, step, it, d.dt, nts, Scur, Snxt,
                                                                         Vxl, Vyl, nullptr);
                std::swap(Scur, Snxt);
            }
        }
        ctx->sim_stats.sat_substeps += max_nts;
        ctx->sim_stats.kernel_launches += sat_launches;
        ctx->sim_stats.sat_kernel_launches += sat_launches;

        timer.mark(4);
        if (obs && d.n_obs > 0) {
            k_gather_obs<<<(nm * d.n_obs + 255) / 256, 256, 0, st>>>(nm, (int)M, d.n_obs, d.obs_cell, Scur,
                                                                       obs, d.n_steps, step);
            ctx->sim_stats.kernel_launches += 1;
        }
        const int hrow = hist_row(step + 1, d.n_steps, d.hist_stride);
        if (S_hist && hrow >= 0) {
            k_copy_rows<<<copy_blocks, 256, 0, st>>>(nm, (int)M, Scur, M, S_hist + (int64_t)hrow * M, (int64_t)nHist * M);
            ctx->sim_stats.kernel_launches += 1;
        }
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
    HM_REQUIRE(d.K_transform == 0 || d.K_transform == 1, "K_transform: 0 = permeability, 1 = K_a + exp(K_b x)");
    HM_REQUIRE(d.Ny <= 1024, "Ny <= 1024 (row tiles of at least two grid rows must fit 2048 cells)");
    HM_REQUIRE(d.precond >= 0 && d.precond <= 4,
               "precond: 0 = multigrid V-cycle (FP32 cycle, FP64 fallback), 1 = Jacobi, 2 = FP64 W-cycle, 3 = FP32 V-cycle, "
               "4 = FP64 V-cycle");
    return HM_OK;
}

}  // namespace

extern "C" int hm_sim_batch(hm_ctx* ctx, const hm_sim_desc* desc) {
    HM_REQUIRE(ctx && desc, "null ctx/desc");
    HM_CHECK(validate(*desc));
    HM_CUDA(cudaSetDevice(ctx->device));
    ctx->sim_stats = hm_sim_stats{};
    ctx->mg_force64 = false;
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
        HM_CHECK(ctx->ws.get("h.S_hist", (size_t)nm * hist_rows(d.n_steps, d.hist_stride) * M * sizeof(double), &p));
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
    if (hd->S_hist)
        HM_CHECK(down(hd->S_hist, d.S_hist, (size_t)nm * hist_rows(d.n_steps, d.hist_stride) * M * sizeof(double)));
    if (hd->obs) HM_CHECK(down(hd->obs, d.obs, (size_t)nm * d.n_steps * d.n_obs * sizeof(double)));
    if (hd->P_last) HM_CHECK(down(hd->P_last, d.P_last, (size_t)nm * M * sizeof(double)));
    if (hd->status) HM_CHECK(down(hd->status, d.status, (size_t)nm * sizeof(int32_t)));
    if (hd->substeps) HM_CHECK(down(hd->substeps, d.substeps, (size_t)nm * d.n_steps * sizeof(int32_t)));
    if (hd->cg_iters) HM_CHECK(down(hd->cg_iters, d.cg_iters, (size_t)nm * d.n_steps * sizeof(int32_t)));
    HM_CUDA(cudaStreamSynchronize(st));
    return HM_OK;
}
