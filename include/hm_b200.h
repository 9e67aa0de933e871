/* hm_b200.h - C ABI of the B200-native history-matching hot path.
 *
 * The reference (patnr/HistoryMatching) is pure Python and has no FFI; its
 * boundary for this path is three Python seams.  Each entry
 * point below states the reference interface it replaces.  The Python host
 * side (historymatching_b200/) binds these with ctypes; INTEGRATION.md shows
 * the stub a maintainer of the reference would add.
 *
 * Conventions
 *  - all floating point data is FP64 (the reference is numpy float64 throughout);
 *  - "device" pointers are caller-owned CUDA device memory on the ctx's device;
 *    the library owns only the workspace inside the ctx;
 *  - work is enqueued on the ctx stream (hm_set_stream); calls that return
 *    per-member results synchronise that stream before returning where stated;
 *  - ensemble axis first: a matrix "(N,M)" is row-major with members as rows
 *    ("transposed" convention, HistoryMatch.py:574-575);
 *  - flat cell index c = ix*Ny + iy (C-order ravel of an (Nx,Ny) field,
 *    HistoryMatch.py:163);
 *  - every function returns 0 on success, a negative hm_status otherwise;
 *    hm_last_error() gives the message.  One ctx per host thread / GPU.
 */
#ifndef HM_B200_H
#define HM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hm_ctx hm_ctx;

enum hm_status {
    HM_OK = 0,
    HM_ERR_CUDA = -1,     /* a CUDA / cuSOLVER call failed */
    HM_ERR_ARG = -2,      /* invalid argument */
    HM_ERR_NUMERIC = -3,  /* factorisation failed / non-finite result */
    HM_ERR_NO_DEVICE = -4
};

/* per-member status bits written by hm_sim_batch */
#define HM_MEMBER_CG_NOT_CONVERGED 1
#define HM_MEMBER_NON_FINITE 2

int hm_version(void);
const char* hm_last_error(void);

int hm_ctx_create(int device, hm_ctx** out);
int hm_ctx_destroy(hm_ctx* ctx);
/* cudaStream_t passed as void*; NULL = the legacy default stream */
int hm_set_stream(hm_ctx* ctx, void* cuda_stream);
int hm_synchronize(hm_ctx* ctx);
/* number of this library's kernels launched on the ctx since creation (its own kernels only,
 * cuSOLVER calls are not counted) */
int hm_launch_count(hm_ctx* ctx, int64_t* out);

/* ---------------------------------------------------------------------------
 * Ensemble forward run.  Replaces, for a whole ensemble at once, the
 * multiprocessing map  utils.apply(comp1, ...)  (tools/utils.py:155-242,
 * HistoryMatch.py:358-387) over  ResSim.sim(dt, nSteps, S0)  of the external
 * TPFA_ResSim package (HistoryMatch.py:224,362; Optimise.py:116): per time
 * step a TPFA pressure solve, CFL sub-stepped explicit upwind saturation
 * transport and the gather of saturations at the observation cells
 * (obs_model, HistoryMatch.py:212-213).
 * ------------------------------------------------------------------------- */
typedef struct hm_sim_desc {
    int32_t n_members;
    int32_t Nx, Ny;
    double Lx, Ly;
    double vw, vo, swc, sor; /* fluid: viscosities, irreducible saturations */

    const double* K;         /* permeability, [member][comp][M] */
    int64_t K_member_stride; /* elements; 0 = one field shared by all members */
    int64_t K_comp_stride;   /* elements from Kx to Ky; 0 = isotropic */
    const double* por;       /* [M] porosity or NULL (= 1) */

    int32_t n_wells;
    const int32_t* well_cell; /* [member][well] flat cell index */
    int64_t well_cell_member_stride; /* 0 = shared */
    const double* well_rate;  /* signed: + injection, - production; [member][step][well] */
    int64_t well_rate_member_stride; /* 0 = shared */
    int64_t well_rate_step_stride;   /* 0 = constant in time */

    const double* S0;         /* initial water saturation [member][M] */
    int64_t S0_member_stride; /* 0 = shared */
    double dt;
    int32_t n_steps;

    int32_t n_obs;
    const int32_t* obs_cell;  /* [n_obs] cells whose saturation is observed */

    double* S_last;   /* out [member][M] saturation after the last step */
    double* S_hist;   /* out NULL or [member][n_hist][M], row 0 = S0; n_hist = n_steps+1, or with hist_stride = k > 1
                       * 1 + ceil(n_steps / k): the states after steps k, 2k, ... and after the last step */
    double* obs;      /* out NULL or [member][n_steps][n_obs] */
    double* P_last;   /* out NULL or [member][M] pressure of the last step */
    int32_t* status;  /* out NULL or [member] HM_MEMBER_* bits */
    int32_t* substeps; /* out NULL or [member][n_steps] CFL sub-step counts */
    int32_t* cg_iters; /* out NULL or [member][n_steps] pressure-solver iterations */

    double cg_rtol;      /* <=0: default 1e-12 (||r|| <= rtol ||q||) */
    int32_t cg_max_iter; /* <=0: default 100*(Nx+Ny)+200 */
    int32_t chunk_members; /* <=0: all members in one launch wave */
    int32_t precond;       /* pressure preconditioner (CG itself, its operator and the convergence test are always FP64):
                            * 0 = multigrid V-cycle, default: the cycle runs in FP32 arithmetic; a solve that needs more than
                            *     40 iterations restarts with the FP64 cycle and the rest of the call stays FP64
                            *     (hm_sim_stats.mg_fp64_fallbacks); grids of <= 2048 cells: fused kernel, FP64 cycle;
                            * 1 = Jacobi, 2 = FP64 multigrid W-cycle, 3 = FP32 V-cycle without fallback, 4 = FP64 V-cycle */
    int32_t mg_switch_iters; /* precond 0: iterations after which a solve switches to the FP64 cycle; <= 0: default (40, and
                            * 80 for the cold first solve of a call).  A positive value also makes the switch sticky at once. */
    int32_t sat_block;     /* kernel selection.  0 = automatic: grids of <= 2048 cells run the whole simulator in ONE kernel,
                            * one CTA per member (hm_small.cu); larger grids take the streamed path with the cluster
                            * transport kernel (all sub-steps of a time step in one launch) where a member's tiles fit a
                            * thread-block cluster, else the streaming transport kernel.  1 = streamed path, streaming
                            * transport kernel (one sub-step per launch).  2 = streamed path, cluster transport kernel.
                            * The streaming kernel stages its tile with bulk copies (cp.async.bulk) when Ny is even.
                            * 4 = as 2 with tiles of 1024 cells, 512 threads, two CTAs per SM (measured: same speed).
                            * 5 = as 1 with the plain-load streaming kernel.  6 = as 1 with 2048-cell tiles (default: 4096
                            * cells, 8 per thread, where the tile fits 110 KB of shared memory).
                            * Automatic choice on the streamed path: the temporally blocked kernel k_sat_tb (7) where the grid
                            * qualifies (Ny a multiple of 64, no porosity field), else the cluster kernel (2) where a member's
                            * tiles fit a cluster, else the streaming kernel (1).  7 = streamed path, k_sat_tb forced. */
    int32_t hist_stride;   /* <= 1: S_hist holds every step (the reference's ResSim.sim output); k > 1: every k-th step and
                            * the last one - the saturation history of a large ensemble for plotting / animation cells
                            * (HistoryMatch.py:233, 1212-1214) without n_steps+1 fields per member */
    int32_t warm_start;    /* initial guess of a pressure solve: 0 = linear extrapolation of the two previous pressures
                            * (default), 1 = the previous pressure (measured: same iteration counts) */
    int32_t tb_cluster_rows; /* temporally blocked transport kernel (sat_block 0 / 7): tiles of a cluster along the grid rows;
                            * <= 0: chosen from the grid and the device's cluster occupancy */
    int32_t tb_halo;       /* the same kernel: sub-steps per round (= overlap rows of neighbouring row strips) when a member
                            * does not fit one cluster; <= 0: automatic */
    int32_t K_transform;   /* 0: K holds permeabilities.  1: K holds the notebook's log-permeability parameter x and the
                            * permeability is K_a + exp(K_b x) - perm_transf, HistoryMatch.py:137-138 (0.1 + exp(5 x)) -
                            * evaluated where the transmissibilities are built, so that an ensemble update can hand its
                            * parameter matrix to the forward run as it is */
    double K_a, K_b;
} hm_sim_desc;

/* statistics of the last hm_sim_batch on this ctx (host side) */
typedef struct hm_sim_stats {
    int64_t cg_iterations;   /* sum over steps of the iterations launched */
    int64_t sat_substeps;    /* sum over steps of max-over-members sub-steps */
    int64_t kernel_launches; /* kernels launched by the call */
    int64_t cg_kernel_launches;
    int64_t sat_kernel_launches;
    int64_t mg_fp64_fallbacks; /* pressure solves that switched from the FP32 to the FP64 multigrid cycle */
    int64_t cg_restarts;       /* long pressure solves continued after the check of the true residual */
    int64_t sat_resident_ctas; /* cluster transport kernel: CTAs the GPU holds at once (0: other transport path) */
    int64_t sat_tb_cluster;    /* temporally blocked transport kernel: CTAs per cluster (0: other transport path), */
    int64_t sat_tb_strips;     /* row strips per member (1: a cluster holds the whole member), */
    int64_t sat_tb_halo;       /* sub-steps per round = overlap rows of neighbouring strips (0 with one strip) */
    int64_t sat_cell_updates;  /* cell updates executed by the transport kernels, redundant halo-row updates included */
} hm_sim_stats;

/* All pointers in the descriptor are DEVICE pointers.  Synchronises the ctx
 * stream before returning (the sub-step count is data dependent). */
int hm_sim_batch(hm_ctx* ctx, const hm_sim_desc* desc);
/* Same with HOST pointers: stages through device memory owned by the ctx
 * (host->device and device->host copies happen inside the call). */
int hm_sim_batch_host(hm_ctx* ctx, const hm_sim_desc* desc);
int hm_sim_get_stats(hm_ctx* ctx, hm_sim_stats* out);
/* ms of device time per phase of the last hm_sim_batch: [setup, cg, flux, saturation, obs] */
int hm_sim_get_phase_ms(hm_ctx* ctx, double out[5]);

/* ---------------------------------------------------------------------------
 * Dense FP64 building block (tensor-core DMMA).  C = alpha*op(A)*op(B) + beta*C,
 * row-major, op = transpose when the flag is non-zero.  Replaces the numpy
 * "@" products of HistoryMatch.py:583-586, 920, 928-941.
 * ------------------------------------------------------------------------- */
int hm_dgemm(hm_ctx* ctx, int transA, int transB, int64_t m, int64_t n, int64_t k,
             double alpha, const double* A, int64_t lda, const double* B, int64_t ldb,
             double beta, double* C, int64_t ldc);

/* ---------------------------------------------------------------------------
 * ES analysis.  Replaces ens_update0 (HistoryMatch.py:578-586):
 *   E + D pinv(S^T S + (N-1) I) S^T X,  S = center(Eo) decorr,
 *   D = (obs - Eo - perturbs) decorr,  X = center(E).
 * E (N,M) is updated in place; the M columns (parameters) are independent, so a
 * parameter-sharded multi-GPU caller passes its own column block as E with M =
 * the number of columns it holds (Eo, perturbs are always the full (N,p)
 * blocks).  ldE = row stride of E.
 * ------------------------------------------------------------------------- */
int hm_es_update(hm_ctx* ctx, int64_t N, int64_t M, int64_t p, double* E, int64_t ldE,
                 const double* Eo, const double* obs, const double* perturbs,
                 const double* decorr);
int hm_es_update_host(hm_ctx* ctx, int64_t N, int64_t M, int64_t p, double* E,
                      const double* Eo, const double* obs, const double* perturbs,
                      const double* decorr);

/* Localised ES.  Replaces ens_update0_loc (HistoryMatch.py:774-797):
 * one tapered local analysis per parameter (column of E).  taper is (M,p)
 * row-major for the M columns held; entries with sqrt(taper) <= 1e-2 are
 * excluded exactly as in the reference. */
int hm_les_update(hm_ctx* ctx, int64_t N, int64_t M, int64_t p, double* E, int64_t ldE,
                  const double* Eo, const double* obs, const double* perturbs,
                  const double* decorr, const double* taper);

/* Distance taper computed on the device.  Replaces
 * loc.bump(loc.pairwise_distances(xy_prm, xy_obs) / radius, sharpness)
 * (tools/localization.py:9-92; HistoryMatch.py:717,863): out is (M,p). */
int hm_taper_bump(hm_ctx* ctx, int64_t M, int64_t p, const double* xy_prm /* (M,2) */,
                  const double* xy_obs /* (p,2) */, double radius, double sharpness,
                  double* out);

/* Anomalies and mean.  Replaces utils.center (tools/utils.py:10-28) along axis 0:
 * X = E - mean (may alias E), mean (M,) may be NULL; scale = sqrt(N/(N-1)) if rescale. */
int hm_center(hm_ctx* ctx, int64_t N, int64_t M, const double* E, int64_t ldE, double* X,
              int64_t ldX, double* mean, int rescale);

/* One Gauss-Newton step of the iterative smoother in the ensemble subspace.
 * Replaces the loop body of IES (HistoryMatch.py:927-942): given W (N,N), the
 * predicted data Eo (N,p) of E = x0 + W X0, y = obs decorr, Dp = perturbs decorr
 * it overwrites W with W + xStep * (grad_y + grad_b) covw. */
int hm_ies_step(hm_ctx* ctx, int64_t N, int64_t p, double* W, const double* Eo,
                const double* obs, const double* perturbs, const double* decorr, double xStep);

/* Localised iterative smoother.  Replaces the loop body of ILES (HistoryMatch.py:1031-1062):
 * Ws holds one (N,N) weight matrix per parameter, (M,N,N) row-major, updated in place with the
 * tapered Gauss-Newton step of every parameter; taper is (M,p). */
int hm_iles_step(hm_ctx* ctx, int64_t N, int64_t M, int64_t p, double* Ws, const double* Eo,
                 const double* obs, const double* perturbs, const double* decorr,
                 const double* taper, double xStep);
/* E[:, i] = x0[i] + Ws[i] X0[:, i]  (recompose, HistoryMatch.py:1020-1021); E, X0 are (N,M). */
int hm_iles_recompose(hm_ctx* ctx, int64_t N, int64_t M, const double* Ws, const double* X0,
                      const double* x0, double* E);

/* Ensemble covariance / correlation fields.  Replaces utils.cov / utils.corr (tools/utils.py:31-55) as used
 * by the correlation dashboards and the max-correlation paths (HistoryMatch.py:478-482, 738-748, 829-833)
 * without pulling the ensemble to the host: a (N,M), b (N,q), out (M,q) row-major,
 *   out = center(a)^T center(b) / (N-1)                                   (corr == 0)
 *   out = clip(cov / std(a, ddof=1)[:,None] / std(b, ddof=1)[None,:], -999, 999)   (corr != 0). */
int hm_corr(hm_ctx* ctx, int64_t N, int64_t M, int64_t q, const double* a, int64_t lda,
            const double* b, int64_t ldb, double* out, int corr);

/* Strided block copy dst[r*ldd + c] = src[r*lds + c], rows x cols (device pointers): the pack / unpack step of the
 * member-row <-> parameter-column re-sharding around the all-to-all of the multi-GPU analysis (SURVEY.md section 8(e);
 * the reference has no counterpart - its ensemble lives in one process, tools/utils.py:155-242). */
int hm_copy2d(hm_ctx* ctx, int64_t rows, int64_t cols, const double* src, int64_t lds, double* dst, int64_t ldd);

/* dst[j][i][:] = src[i][j][:] for a dense (d0, d1, d2) array (device pointers, out of place).  Used by the scalable
 * separable prior sampler (the stand-in for geostat.gaussian_fields, tools/geostat.py:86-99, at grid sizes where its dense
 * M x M covariance is infeasible): the left factor of Fx Z Fy^T is applied to all members in ONE GEMM on an
 * (Nx, N, Ny) layout, this kernel returns the member-major layout. */
int hm_swap01(hm_ctx* ctx, int64_t d0, int64_t d1, int64_t d2, const double* src, double* dst);

#ifdef __cplusplus
}
#endif
#endif /* HM_B200_H */
