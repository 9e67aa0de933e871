// Temporally blocked upwind transport: k_sat_tb.
//
// Replaces, for a whole ensemble, the CFL sub-step loop of the reference simulator's saturation step
// (TPFA_ResSim saturation_step_upwind, called through ResSim.sim at HistoryMatch.py:224,362; SURVEY.md Appendix
// A.3):   repeat Nts times:  S <- S + dtx * (B fw(S)) + fi * dtx.
//
// The face fluxes are frozen during the Nts sub-steps of a time step, so a block of cells can be advanced many
// sub-steps on chip.  Layout:
//   * a CTA (512 threads) owns a tile of R rows x W columns = 4096 cells; a thread owns a patch of 4 rows x 2
//     columns and keeps, in REGISTERS for the whole launch, the 8 saturations and the 22 signed face
//     coefficients dtx*v of the patch (2.75 coefficients per cell instead of the 5 one-sided ones of
//     k_sat_cluster, so an SM holds 4096 cells instead of 2048);
//   * per sub-step a thread evaluates fw(S) of its 8 cells, publishes them in shared memory, applies the 10 faces
//     inside its patch from registers, and after ONE CTA barrier the 12 faces on the patch boundary from shared
//     memory.  The published row is split into its even and its odd columns, so that every access of the loop is
//     a conflict-free 64-bit access: 20 shared-memory accesses per 8 cells (k_sat_cluster: 32);
//   * the upwind value of a face is chosen by the sign of the coefficient (integer compare on the high word +
//     two selects on the ALU pipe): S_lo -= w f_up, S_hi += w f_up is 4 FP64 FMAs per cell, with the 7 of fw(S)
//     11 FP64 operations per cell and sub-step (k_sat_cluster: 13) - the FP64 pipe is what binds this kernel;
//   * the cx x cy CTAs of a thread-block cluster tile a strip of cx*R rows x cy*W = Ny columns; halo rows and
//     halo columns travel through distributed shared memory with st.async + mbarrier (no cluster barrier in the
//     loop), double buffered like the tiles.
// Temporal blocking across HBM: where a member does not fit one cluster (512^2: 64 tiles), it is cut into
// overlapping row strips.  One launch ("round") advances every strip k sub-steps: a strip is loaded with k extra
// rows at each inner edge, the error of the unknown outside travels one row per sub-step, and after k sub-steps
// the rows at distance >= k from the inner edges are exact and written back.  HBM is touched once per k sub-steps
// (40 B per valid cell per round instead of 32 B per cell per sub-step), at the price of (cx R) / (cx R - 2k)
// redundant cell updates.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "hm_ptx.cuh"
#include "hm_sim_common.cuh"

// the sub-step lambda captures the patch registers (S, wx, wy) by reference before the first work item fills them
#pragma nv_diag_suppress 549

namespace hmsim {

constexpr int kTbThreads = 512;
constexpr int kTbCells = 4096;

struct TbArgs {
    int Nx, Ny;
    int64_t M;
    int cx, cy;     // CTAs of a cluster along the rows / along the columns
    int nStrips;    // row strips per member
    int stride;     // rows between the starts of consecutive strips
    int halo;       // rows that are not written back at an inner strip edge = sub-steps per round
    double h2;      // hx * hy
    int step;       // time step (well rate schedule)
    int it0, kmax;  // this launch advances sub-steps it0 .. it0 + kmax - 1 (clipped to the member's count)
    int nWork;      // work items (member, strip) of the launch, dealt round-robin to the persistent clusters
    double dt;
    const int* nts;
    const double* Sin;
    double* Sout;
    const double* Vxl;
    const double* Vyl;
};

// upwind value of a face with signed coefficient w (flow from the low to the high cell when w > 0); the sign test
// is an integer compare on the high word: no FP64 pipe slot (a positive denormal counts as "not positive", its
// contribution is below 1e-300 either way)
__device__ __forceinline__ double upwind(double w, double f_lo, double f_hi) {
    return __double2hiint(w) > 0 ? f_lo : f_hi;
}

template <int W>
struct TbLayout {
    static constexpr int H = W / 2;            // threads per row group (each owns 2 columns)
    static constexpr int R = kTbCells / W;     // rows of the CTA tile
    static constexpr int RSD = W + 2;          // doubles per published row: E[0..H] (E[H] = right halo), O[-1..H-1] (O[-1] = left halo)
    static constexpr int NQ = R / 4;           // row groups
    static constexpr int BUF = (R + 2) * RSD;  // one fw buffer: halo row, R tile rows, halo row
    static constexpr int OO = H + 2;           // offset of O[c] behind E[c] in a row
    // staging of the NEXT work item (bulk copies, one mbarrier): S rows, x-flux rows (one more), y-flux rows (two more columns)
    static constexpr int ST_S = 2 * BUF, ST_X = ST_S + R * W, ST_Y = ST_X + (R + 1) * W, TOTAL = ST_Y + R * (W + 2);
};

template <int W, bool UNIT>
__global__ void __launch_bounds__(kTbThreads, 1) k_sat_tb(TbArgs a, Fluid fl, Wells w) {
    namespace cg = cooperative_groups;
    using L = TbLayout<W>;
    constexpr int H = L::H, R = L::R, RSD = L::RSD, NQ = L::NQ, BUF = L::BUF, OO = L::OO;
    extern __shared__ __align__(16) double smt[];  // fw[2][BUF], staging
    __shared__ int wcs[2][kMaxWells];  // wells and sub-step count of this and of the next work item (cp.async)
    __shared__ double wrs[2][kMaxWells];
    __shared__ int ntss[2];
    __shared__ int wl_tid[kMaxWells], wl_slot[kMaxWells];
    __shared__ double wl_neg[kMaxWells], wl_pos[kMaxWells];
    __shared__ int wl_n[2];
    __shared__ __align__(8) unsigned long long bars[4];  // halo exchange (per fw buffer), staging, CTA split barrier

    cg::cluster_group cluster = cg::this_cluster();
    const int csize = a.cx * a.cy;
    const int rank = (int)cluster.block_rank();
    const int cxi = rank / a.cy, cyi = rank - cxi * a.cy;
    const int sMax = max(a.Nx - a.cx * R, 0);
    const int tid = threadIdx.x, q = tid / H, c = tid - q * H;
    const int col0 = cyi * W;
    // Persistent clusters: cluster k advances the work items (member, strip) k, k + nClusters, ... of this round.
    const int nClusters = gridDim.x / csize, nWork = a.nWork;
    int work = blockIdx.x / csize;

    for (int e = tid; e < 2 * BUF; e += kTbThreads) smt[e] = 0.0;  // halo slots without a neighbour stay zero
    double* const stS = smt + L::ST_S;
    double* const stX = smt + L::ST_X;
    double* const stY = smt + L::ST_Y;

    // halo exchange: this thread's one row neighbour (up or down) and one column neighbour (left or right)
    const uint32_t bar0 = smem_u32(&bars[0]), bar1 = smem_u32(&bars[1]), ldbar = smem_u32(&bars[2]), ctabar = smem_u32(&bars[3]);
    int cpar = 0;
    const bool hasUp = cxi > 0, hasDn = cxi < a.cx - 1, hasLf = cyi > 0, hasRt = cyi < a.cy - 1;
    const int haloBytes = 8 * (W * ((int)hasUp + (int)hasDn) + R * ((int)hasLf + (int)hasRt));
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar1, 1);
        mbar_init(ldbar, 1);
        mbar_init(ctabar, kTbThreads);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (haloBytes) {  // first expectation of either barrier, re-posted in the loop by a waiter
            mbar_expect_tx(bar0, haloBytes);
            mbar_expect_tx(bar1, haloBytes);
        }
    }
    const bool sUp = hasUp && q == 0, sDn = hasDn && q == NQ - 1, sLf = hasLf && c == 0, sRt = hasRt && c == H - 1;
    const bool sX = sUp || sDn, needWait = sX || sLf || sRt;
    const int posterTid = (hasUp || hasLf) ? 0 : hasRt ? H - 1 : (NQ - 1) * H;
    const bool poster = needWait && tid == posterTid;
    double* const fb0 = smt;
    double* const fb1 = smt + BUF;
    const int tb = (4 * q + 1) * RSD + c;  // E slot of the patch's first row
    // my first row is the lower halo row (R) of the tile above, my last row the upper halo row (-1) of the tile below
    const int xoff = sUp ? (R + 1) * RSD + c : c;
    const int xrank = sUp ? rank - a.cy : sDn ? rank + a.cy : rank;
    // my first column is the right halo column E[H] of the left tile, my last column the left halo O[-1] of the right tile
    const int yoff = (4 * q + 1) * RSD + (sLf ? H : H + 1);
    const int yrank = sLf ? rank - 1 : sRt ? rank + 1 : rank;
    const uint32_t xa0 = map_to_cta(smem_u32(fb0 + xoff), xrank), xa1 = map_to_cta(smem_u32(fb1 + xoff), xrank);
    const uint32_t xb0 = map_to_cta(bar0, xrank), xb1 = map_to_cta(bar1, xrank);
    const uint32_t ya0 = map_to_cta(smem_u32(fb0 + yoff), yrank), ya1 = map_to_cta(smem_u32(fb1 + yoff), yrank);
    const uint32_t yb0 = map_to_cta(bar0, yrank), yb1 = map_to_cta(bar1, yrank);

    // A work item's tile into the staging buffers: bulk copies issued by warp 0, completion on ldbar.  A tile of whole grid
    // rows (cy == 1) is contiguous in memory: three copies; otherwise one copy per tile row and array.  Rows beyond the
    // grid (ragged last tile of a single-strip member) are not copied and read as zero below.  The item's wells and its
    // sub-step count follow with cp.async (completion: cp.async.wait_all of the issuing threads + the next CTA barrier).
    const int ysd = a.cy == 1 ? W : W + 2;  // row pitch of the staged y-fluxes
    auto prefetch = [&](int wk, int slot) {
        const int pm = wk / a.nStrips, ps = wk - pm * a.nStrips;
        if (tid < 32) {
            const int tr0 = min(ps * a.stride, sMax) + cxi * R;
            const int rowsS = max(0, min(R, a.Nx - tr0)), rowsX = max(0, min(R + 1, a.Nx - tr0));
            const int64_t g0 = (int64_t)pm * a.M + (int64_t)tr0 * a.Ny + col0;
            if (a.cy == 1) {
                if (tid == 0) {
                    const uint32_t bS = 8u * rowsS * W, bX = 8u * rowsX * W, bY = rowsS ? bS + 16u : 0u;
                    mbar_expect_tx(ldbar, (int)(bS + bX + bY));
                    if (bS) bulk_g2s(smem_u32(stS), a.Sin + g0, bS, ldbar);
                    if (bX) bulk_g2s(smem_u32(stX), a.Vxl + g0, bX, ldbar);
                    if (bY) bulk_g2s(smem_u32(stY), a.Vyl + g0, bY, ldbar);
                }
            } else {
                if (tid == 0) mbar_expect_tx(ldbar, 8 * (rowsS * W + rowsX * W + rowsS * (W + 2)));
                __syncwarp();
                for (int r = tid; r < rowsS; r += 32) bulk_g2s(smem_u32(stS + r * W), a.Sin + g0 + (int64_t)r * a.Ny, 8u * W, ldbar);
                for (int r = tid; r < rowsX; r += 32) bulk_g2s(smem_u32(stX + r * W), a.Vxl + g0 + (int64_t)r * a.Ny, 8u * W, ldbar);
                for (int r = tid; r < rowsS; r += 32)
                    bulk_g2s(smem_u32(stY + r * (W + 2)), a.Vyl + g0 + (int64_t)r * a.Ny, 8u * (W + 2), ldbar);
            }
        } else if (tid < 32 + w.n) {
            const int i = tid - 32;
            cp_async_4(smem_u32(&wcs[slot][i]), w.cell + (int64_t)pm * w.cell_ms + i);
            cp_async_8(smem_u32(&wrs[slot][i]), w.rate + (int64_t)pm * w.rate_ms + (int64_t)a.step * w.rate_ss + i);
        } else if (tid == 32 + kMaxWells) {
            cp_async_4(smem_u32(&ntss[slot]), a.nts + pm);
        }
    };
    __syncthreads();  // ldbar initialised before warp 0 uses it
    if (work < nWork) prefetch(work, 0);
    cluster.sync();  // tiles zeroed and mbarriers initialised in every CTA before remote traffic starts

    // state of the patch: saturations and signed face coefficients dtx * v.  wx[r][j]: face between rows r-1 and r
    // of the patch (r = 0..4), wy[r][j]: face between columns j-1 and j (j = 0..2).
    double S[4][2], wx[5][2], wy[4][3];
    int wcode = -1, nwl = 0;
    if (tid < 2) wl_n[tid] = 0;  // visible behind the barrier / cluster barrier below

    // Hazards (as in k_sat_cluster): the threads that read a halo row / column filled by a neighbour are exactly the
    // threads that send the matching edge row / column to that neighbour, and a sender has waited for the complete
    // previous phase, so no neighbour runs more than one sub-step ahead of a reader; tiles and halos are double
    // buffered.  The sub-step counter (buffer and mbarrier phase) runs on across the work items of a cluster, whose CTAs
    // all execute the same sequence of sub-steps, so no cluster barrier is needed between items or at the end: a CTA
    // has received everything its neighbours send before it leaves its last sub-step.
    auto substep = [&](double* __restrict__ fw, uint32_t mybar, uint32_t xa, uint32_t xb, uint32_t ya, uint32_t yb,
                       int parity) {
        double f[4][2];
        if constexpr (NQ <= 2) {
            // every warp owns an edge row: the two candidate edge rows first, sent at once, so that the neighbouring tile
            // has them as early as possible; the other rows depend (through `one`) on a volatile asm behind the sends,
            // which keeps the compiler from hoisting them above
            f[0][0] = frac_flow_loop<UNIT>(S[0][0], fl);
            f[0][1] = frac_flow_loop<UNIT>(S[0][1], fl);
            f[3][0] = frac_flow_loop<UNIT>(S[3][0], fl);
            f[3][1] = frac_flow_loop<UNIT>(S[3][1], fl);
            if (sX) {
                st_async_f64(xa, sUp ? f[0][0] : f[3][0], xb);
                st_async_f64(xa + 8u * OO, sUp ? f[0][1] : f[3][1], xb);
            }
            double one;
            asm volatile("mov.f64 %0, 0d3FF0000000000000;" : "=d"(one));
#pragma unroll
            for (int r = 1; r < 3; ++r) {
                f[r][0] = frac_flow_one<UNIT>(S[r][0], fl, one);
                f[r][1] = frac_flow_one<UNIT>(S[r][1], fl, one);
            }
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                f[r][0] = frac_flow_loop<UNIT>(S[r][0], fl);
                f[r][1] = frac_flow_loop<UNIT>(S[r][1], fl);
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            fw[r * RSD] = f[r][0];
            fw[r * RSD + OO] = f[r][1];
        }
        // Split CTA barrier (an mbarrier counting all threads): arrive as soon as this thread's fw values are published, wait
        // only where the neighbours' values are needed - the halo sends, the faces inside the patch and the well terms of
        // a warp overlap with the other warps' publishing (measured at 128^2: 18.2 -> 16.2 ms per launch against bar.sync)
        // (whole-row tiles: the barrier operations sit behind a branch on a run-time constant - a scheduling fence for ptxas,
        // which otherwise sinks the arrive below the sends and the interior faces: measured at 512^2, 488 vs 459 ms per step)
        if (NQ > 2 || a.kmax > 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ctabar) : "memory");
        // edge rows / columns into the neighbours' halo slots.  (STAS cannot be predicated: one branch per direction,
        // the row direction is warp-uniform, the column direction is taken by one lane per warp.)
        if (NQ > 2 && sX) {
            st_async_f64(xa, sUp ? f[0][0] : f[3][0], xb);
            st_async_f64(xa + 8u * OO, sUp ? f[0][1] : f[3][1], xb);
        }
        if (sLf || sRt) {
#pragma unroll
            for (int r = 0; r < 4; ++r) st_async_f64(ya + 8u * (r * RSD), sLf ? f[r][0] : f[r][1], yb);
        }
        // the 10 faces inside the patch
#pragma unroll
        for (int r = 1; r < 4; ++r)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const double fu = upwind(wx[r][j], f[r - 1][j], f[r][j]);
                S[r - 1][j] = fma(-wx[r][j], fu, S[r - 1][j]);
                S[r][j] = fma(wx[r][j], fu, S[r][j]);
            }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const double fu = upwind(wy[r][1], f[r][0], f[r][1]);
            S[r][0] = fma(-wy[r][1], fu, S[r][0]);
            S[r][1] = fma(wy[r][1], fu, S[r][1]);
        }
        if (wcode >= 0) {  // S += dtx * (min(q,0) fw(S) + max(q,0)) on the cells that hold wells: one lane of one warp
            const int we0 = wcode & 255;
            const double qn = wl_neg[we0], qp = wl_pos[we0];
            switch ((wcode >> 8) & 255) {  // the slot is fixed for the work item: a jump table, then two FP64 operations
                case 0: S[0][0] += fma(qn, f[0][0], qp); break;
                case 1: S[0][1] += fma(qn, f[0][1], qp); break;
                case 2: S[1][0] += fma(qn, f[1][0], qp); break;
                case 3: S[1][1] += fma(qn, f[1][1], qp); break;
                case 4: S[2][0] += fma(qn, f[2][0], qp); break;
                case 5: S[2][1] += fma(qn, f[2][1], qp); break;
                case 6: S[3][0] += fma(qn, f[3][0], qp); break;
                default: S[3][1] += fma(qn, f[3][1], qp); break;
            }
            if (wcode >> 16) {  // further wells in the same 4 x 2 patch (rare)
                for (int e = we0 + 1; e < nwl; ++e) {
                    if (wl_tid[e] != tid) continue;
                    const int sl = wl_slot[e];
                    const double qn2 = wl_neg[e], qp2 = wl_pos[e];
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int j = 0; j < 2; ++j)
                            if (sl == 2 * r + j) S[r][j] += fma(qn2, f[r][j], qp2);
                }
            }
        }
        if (NQ > 2 || a.kmax > 0) {
            mbar_wait(ctabar, cpar);
            cpar ^= 1;
        }
        // the 12 faces on the boundary of the patch
        auto halo_wait = [&]() {
            mbar_wait(mybar, parity);
            if (poster) mbar_expect_tx(mybar, haloBytes);
        };
        auto top = [&]() {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const double fn = fw[-RSD + j * OO];
                S[0][j] = fma(wx[0][j], upwind(wx[0][j], fn, f[0][j]), S[0][j]);
            }
        };
        auto bottom = [&]() {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const double fs = fw[4 * RSD + j * OO];
                S[3][j] = fma(-wx[4][j], upwind(wx[4][j], f[3][j], fs), S[3][j]);
            }
        };
        auto sides = [&]() {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const double fl_ = fw[r * RSD + OO - 1];  // O[c-1]: column 2c-1
                S[r][0] = fma(wy[r][0], upwind(wy[r][0], fl_, f[r][0]), S[r][0]);
                const double fr_ = fw[r * RSD + 1];  // E[c+1]: column 2c+2
                S[r][1] = fma(-wy[r][2], upwind(wy[r][2], f[r][1], fr_), S[r][1]);
            }
        };
        if constexpr (NQ <= 2) {
            // Tiles of two row groups (W = 512): EVERY warp owns an edge row and waits for a halo row each sub-step.  First the
            // faces whose neighbour lives in this CTA, then - behind the wait - the two that face another tile of the cluster:
            // the DSMEM latency overlaps with 10 of the 12 faces (measured at 512^2: 488 -> 462 ms per step; on taller tiles,
            // where only 4 of 16 warps wait, the extra branches cost more than they hide: 15.6 -> 17.3 ms at 128^2).
            bool waited = false;
            if (sLf || sRt) {  // a halo column is needed by sides()
                halo_wait();
                waited = true;
            }
            sides();
            if (!sUp) top();
            if (!sDn) bottom();
            if (sX) {
                if (!waited) halo_wait();
                if (sUp) top();
                if (sDn) bottom();
            }
        } else {
            if (needWait) halo_wait();
            top();
            bottom();
            sides();
        }
    };
    double* const fwa = fb0 + tb;
    double* const fwb = fb1 + tb;
    int gsub = 0;  // sub-steps executed by this cluster so far: buffer = gsub & 1, mbarrier phase = (gsub >> 1) & 1

    for (int item = 0; work < nWork; work += nClusters, ++item) {
        const int m = work / a.nStrips, sidx = work - m * a.nStrips;
        const int s0 = min(sidx * a.stride, sMax);      // first grid row of the strip
        const int vlo = sidx == 0 ? 0 : s0 + a.halo;    // rows [vlo, vhi) are exact after the round and written back
        const int vhi = sidx == a.nStrips - 1 ? a.Nx : min((sidx + 1) * a.stride, sMax) + a.halo;
        const int row0 = s0 + cxi * R + 4 * q;          // first grid row of the thread's patch
        const int slot = item & 1;

        // staged tile -> registers.  Rows beyond the grid hold S = 0 and zero coefficients: inert.
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();  // this item's wells and sub-step count (cp.async of other threads) are visible
        const int n = ntss[slot];
        const int nr = max(0, min(a.kmax, n - a.it0));  // sub-steps of this launch
        const double dts = n > 0 ? a.dt / (double)n : 0.0;
        const double dtx = dts / a.h2;
        const int* wc = wcs[slot];
        const double* wr = wrs[slot];
        // wells of this tile: (owning thread, cell slot in its patch, dtx * min(q,0), dtx * max(q,0)); built by warp 1 while
        // the other warps turn the staged tile into registers
        for (int i = tid - 32; i >= 0 && i < w.n; i += kTbThreads) {
            const int cc = wc[i];
            bool first = true;
            for (int k = 0; k < i; ++k) first = first && (wc[k] != cc);
            if (!first) continue;
            const int gr = cc / a.Ny, gc = cc - gr * a.Ny;
            const int lr = gr - (s0 + cxi * R), lc = gc - col0;
            if (lr < 0 || lr >= R || lc < 0 || lc >= W) continue;
            const double qs = cell_source(cc, w.n, wc, wr) * dtx;
            const int e = atomicAdd(&wl_n[slot], 1);
            wl_tid[e] = (lr >> 2) * H + (lc >> 1);
            wl_slot[e] = (lr & 3) * 2 + (lc & 1);
            wl_neg[e] = fmin(qs, 0.0);
            wl_pos[e] = fmax(qs, 0.0);
        }
        mbar_wait(ldbar, item & 1);
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            wx[r][0] = wx[r][1] = 0.0;
            if (row0 + r < a.Nx) {
                const double2 v = *reinterpret_cast<const double2*>(stX + (4 * q + r) * W + 2 * c);
                wx[r][0] = dtx * v.x;
                wx[r][1] = dtx * v.y;
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            S[r][0] = S[r][1] = 0.0;
            wy[r][0] = wy[r][1] = wy[r][2] = 0.0;
            if (row0 + r < a.Nx) {
                const double2 sv = *reinterpret_cast<const double2*>(stS + (4 * q + r) * W + 2 * c);
                const double* yp = stY + (4 * q + r) * ysd + 2 * c;
                const double2 v = *reinterpret_cast<const double2*>(yp);
                S[r][0] = sv.x;
                S[r][1] = sv.y;
                wy[r][0] = dtx * v.x;
                wy[r][1] = dtx * v.y;
                wy[r][2] = dtx * yp[2];  // column Ny is column 0 of the next row: a zero face
            }
        }
        __syncthreads();  // the staging buffers are free; the well list (built by warp 1 meanwhile) is complete
        if (work + nClusters < nWork) prefetch(work + nClusters, slot ^ 1);  // overlaps this item's sub-steps
        if (tid == 0) wl_n[slot ^ 1] = 0;  // the next item's counter: last read before this item's first barrier
        nwl = wl_n[slot];
        // this thread's first entry of the list, its cell slot and "the patch holds further wells", packed into one
        // register (-1: the patch holds no well)
        wcode = -1;
        for (int e = nwl - 1; e >= 0; --e)
            if (wl_tid[e] == tid) wcode = e | (wl_slot[e] << 8) | (wcode >= 0 ? 1 << 16 : 0);

        int sub = 0;
        if ((gsub & 1) && nr > 0) {  // an odd number of sub-steps so far: the next one uses the second buffer
            substep(fwb, bar1, xa1, xb1, ya1, yb1, (gsub >> 1) & 1);
            ++sub, ++gsub;
        }
        for (; sub + 1 < nr; sub += 2, gsub += 2) {
            const int par = (gsub >> 1) & 1;
            substep(fwa, bar0, xa0, xb0, ya0, yb0, par);
            substep(fwb, bar1, xa1, xb1, ya1, yb1, par);
        }
        if (sub < nr) {
            substep(fwa, bar0, xa0, xb0, ya0, yb0, (gsub >> 1) & 1);
            ++gsub;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int gr = row0 + r;
            if (gr >= vlo && gr < vhi)
                *reinterpret_cast<double2*>(a.Sout + (int64_t)m * a.M + (int64_t)gr * a.Ny + col0 + 2 * c) =
                    make_double2(S[r][0], S[r][1]);
        }
    }
}

using tb_fn = void (*)(TbArgs, Fluid, Wells);

// Tile width: the largest of 512 / 256 / 128 / 64 that divides the row length.  A tile of whole grid rows (W == Ny) has no
// column neighbours: its halo exchange is whole warps sending whole rows, and its staging copies are contiguous.
static int tb_tile_cols(int Ny) {
    if (getenv("HM_TB_W") && Ny % atoi(getenv("HM_TB_W")) == 0) return atoi(getenv("HM_TB_W"));  // development
    for (int W : {512, 256, 128, 64})
        if (Ny % W == 0) return W;
    return 0;
}

static tb_fn tb_kernel(int W, bool unit) {
    switch (W) {
        case 512: return unit ? k_sat_tb<512, true> : k_sat_tb<512, false>;
        case 256: return unit ? k_sat_tb<256, true> : k_sat_tb<256, false>;
        case 128: return unit ? k_sat_tb<128, true> : k_sat_tb<128, false>;
        default: return unit ? k_sat_tb<64, true> : k_sat_tb<64, false>;
    }
}

static size_t tb_smem(int W) {
    const int n = W == 512 ? TbLayout<512>::TOTAL : W == 256 ? TbLayout<256>::TOTAL : W == 128 ? TbLayout<128>::TOTAL
                                                                                                  : TbLayout<64>::TOTAL;
    return (size_t)n * sizeof(double);
}

bool transport_tb_supported(const hm_sim_desc& d) {
    if (d.por) return false;  // the face coefficients carry dtx of BOTH cells of a face: uniform pore volume only
    const int W = tb_tile_cols(d.Ny);
    return W > 0 && d.Ny / W <= 16;
}

static int tb_launch_cfg(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, int W, int csize, unsigned grid,
                         cudaStream_t st) {
    *cfg = cudaLaunchConfig_t{};
    cfg->gridDim = dim3(grid);
    cfg->blockDim = dim3(kTbThreads);
    cfg->dynamicSmemBytes = tb_smem(W);
    cfg->stream = st;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg->attrs = attr;
    cfg->numAttrs = 1;
    return HM_OK;
}

// Strip geometry for a cluster of cx x cy tiles and k sub-steps per round (see the header comment).
static void tb_strips(int Nx, int R, int cx, int k, int* nStrips, int* stride) {
    const int RS = cx * R;
    if (RS >= Nx) {
        *nStrips = 1;
        *stride = RS;
        return;
    }
    *stride = RS - 2 * k;
    *nStrips = (Nx - RS + *stride - 1) / *stride + 1;
}

// Chooses the cluster shape and the round length for this grid (desc.tb_cluster_rows / tb_halo override), and
// advances every member by its nts[m] sub-steps: Scur -> (ping-pong) -> *Sresult.
int transport_tb(hm_ctx* ctx, const hm_sim_desc& d, const Fluid& fl, const Wells& w, int step, int nm, int max_nts,
                 const int* nts, double* Scur, double* Snxt, const double* Vxl, const double* Vyl, double** Sresult,
                 int* launches) {
    cudaStream_t st = ctx->stream;
    *Sresult = Scur;
    *launches = 0;
    if (max_nts <= 0) return HM_OK;  // no flow in any member
    const int W = tb_tile_cols(d.Ny);
    const int R = kTbCells / W, cy = d.Ny / W;
    const bool unit = fl.inv_range == 1.0 && fl.swc_ir == 0.0 && fl.mr == 1.0;
    tb_fn kern = tb_kernel(W, unit);
    HM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tb_smem(W)));
    HM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));

    // candidates: cx tiles along the rows (cluster of cx*cy <= 16 CTAs), k sub-steps per round.  Cost model:
    // waves of resident clusters x (k sub-steps + launch prologue / epilogue), measured constants in us.
    const int cx_need = (d.Nx + R - 1) / R;
    int best_cx = 0, best_k = 0, best_active = 0;
    double best_cost = 1e300;
    for (int cx = 1; cx <= std::min(cx_need, 16 / cy); ++cx) {
        if (d.tb_cluster_rows > 0 && cx != std::min(d.tb_cluster_rows, std::min(cx_need, 16 / cy))) continue;
        const int csize = cx * cy;
        int& active = ctx->tb_active[W == 512 ? 3 : W == 256 ? 2 : W == 128][unit][csize];
        if (active == 0) {
            cudaLaunchConfig_t cfg;
            cudaLaunchAttribute attr[1];
            tb_launch_cfg(&cfg, attr, W, csize, (unsigned)csize, st);
            int nc = 0;
            if (cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) != cudaSuccess) nc = 0;
            (void)cudaGetLastError();
            active = nc > 0 ? nc : -1;
        }
        if (active <= 0) continue;
        const bool single = cx * R >= d.Nx;
        for (int k : {4, 8, 12, 16, 24, 32}) {
            if (!single && d.tb_halo > 0 && k != d.tb_halo) continue;
            if (!single && 4 * k > cx * R) break;  // at least half of a strip's rows must be written back
            int nStrips, stride;
            tb_strips(d.Nx, R, cx, k, &nStrips, &stride);
            const double waves = std::ceil((double)nm * nStrips / active);
            const double rounds = single ? 1.0 : std::ceil((double)max_nts / k);
            const double per_round = (single ? max_nts : k) * 0.8 + 1.5;  // us per work item: sub-steps + staging / set-up
            const double cost = waves * rounds * per_round;
            if (cost < best_cost) best_cost = cost, best_cx = cx, best_k = single ? 0 : k, best_active = active;
            if (single) break;
        }
    }
    if (best_cx == 0) {
        hm::set_error("k_sat_tb: no cluster shape of this grid (%d x %d) can be resident on the device", d.Nx, d.Ny);
        return HM_ERR_ARG;
    }
    TbArgs a;
    a.Nx = d.Nx;
    a.Ny = d.Ny;
    a.M = (int64_t)d.Nx * d.Ny;
    a.cx = best_cx;
    a.cy = cy;
    a.halo = best_k;
    tb_strips(d.Nx, R, best_cx, best_k, &a.nStrips, &a.stride);
    a.h2 = (d.Lx / d.Nx) * (d.Ly / d.Ny);
    a.step = step;
    a.dt = d.dt;
    a.nts = nts;
    a.Vxl = Vxl;
    a.Vyl = Vyl;
    const int csize = a.cx * a.cy;
    ctx->sim_stats.sat_resident_ctas = (int64_t)best_active * csize;
    ctx->sim_stats.sat_tb_cluster = csize;
    ctx->sim_stats.sat_tb_strips = a.nStrips;
    ctx->sim_stats.sat_tb_halo = a.halo;
    if (getenv("HM_DEBUG") && step == 0)
        fprintf(stderr, "[hm] k_sat_tb<%d>: cluster %d x %d, %d strip(s) of %d rows, stride %d, %d sub-steps per round, "
                "%d resident clusters (%d of %d SMs)\n", W, a.cx, a.cy, a.nStrips, a.cx * R, a.stride, a.halo, best_active,
                best_active * csize, ctx->sm_count);
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    if (getenv("HM_TB_MAX_CLUSTERS")) best_active = std::max(1, std::min(best_active, atoi(getenv("HM_TB_MAX_CLUSTERS"))));  // development
    a.nWork = nm * a.nStrips;
    tb_launch_cfg(&cfg, attr, W, csize, (unsigned)(std::min(a.nWork, best_active) * csize), st);
    const int kround = a.nStrips == 1 ? max_nts : a.halo;
    int n_launch = 0;
    for (int it0 = 0; it0 < max_nts; it0 += kround, ++n_launch) {
        a.it0 = it0;
        a.kmax = kround;
        a.Sin = Scur;
        a.Sout = Snxt;
        HM_CUDA(cudaLaunchKernelEx(&cfg, kern, a, fl, w));
        std::swap(Scur, Snxt);
    }
    *Sresult = Scur;
    *launches = n_launch;
    ctx->sim_stats.sat_cell_updates += (int64_t)nm * a.nStrips * csize * kTbCells * max_nts;
    return HM_OK;
}

}  // namespace hmsim
