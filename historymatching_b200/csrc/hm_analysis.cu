// Ensemble-smoother analysis on sm_100a: ES, localised ES, IES step, taper.
//
// Replaces ens_update0 / ens_update0_loc / the IES loop body
// (HistoryMatch.py:578-586, 774-797, 927-942) and
// loc.bump(loc.pairwise_distances(...)) (tools/localization.py:9-92).
// Big products go through the DMMA GEMM (hm_gemm.cu); the p x p / N x N
// factorisations use cuSOLVER (a library call); the
// per-parameter local analyses are a hand-written batched Cholesky in shared
// memory, one CTA per parameter.
//
// Algebra used (all exact rewrites of the reference expressions):
//  * S = center(Eo) decorr has zero column sums, so S^T X = S^T E: the
//    parameter ensemble never needs centring.
//  * C = S^T S + (N-1) I is SPD with eigenvalues >= N-1, so pinv(C) = C^-1
//    and a Cholesky solve replaces the SVD pseudo-inverse.
//  * ES is evaluated as E + (D C^-1)(S^T E): 4 N p M flops instead of the
//    reference's left-to-right 2 N^2 M.
#include <algorithm>

#include "hm_common.cuh"

namespace hm {
int dgemm(hm_ctx* ctx, bool tA, bool tB, int64_t m, int64_t n, int64_t k, double alpha,
          const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C,
          int64_t ldc);
}

namespace {

constexpr size_t kMaxWorkSmem = 227 * 1024;  // dynamic shared memory a CTA can opt into on sm_100

int ensure_solver(hm_ctx* ctx) {
    if (!ctx->solver) {
        HM_CUSOLVER(cusolverDnCreate(&ctx->solver));
        HM_CUSOLVER(cusolverDnSetStream(ctx->solver, ctx->stream));
    }
    return HM_OK;
}

// ---- column means / anomalies (utils.center, tools/utils.py:10-28) -----------------------------
// One thread per column, coalesced across columns; rows are split over
// blockIdx.y slabs only for the mean pass when N is large.
__global__ void k_col_mean(int64_t N, int64_t M, const double* __restrict__ E, int64_t ldE,
                           double* __restrict__ mean) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    double s = 0.0;
    for (int64_t i = 0; i < N; ++i) s += E[i * ldE + j];
    mean[j] = s / (double)N;
}
__global__ void k_sub_mean(int64_t N, int64_t M, const double* __restrict__ E, int64_t ldE,
                           double* __restrict__ X, int64_t ldX, const double* __restrict__ mean,
                           double scale) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = blockIdx.y;
    if (j >= M) return;
    X[i * ldX + j] = (E[i * ldE + j] - mean[j]) * scale;
}

// D0 = obs - Eo - perturbs (HistoryMatch.py:584)
__global__ void k_innovation(int64_t N, int64_t p, const double* __restrict__ obs,
                             const double* __restrict__ Eo, const double* __restrict__ pert,
                             double* __restrict__ D0) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * p) return;
    D0[i] = obs[i % p] - Eo[i] - pert[i];
}

__global__ void k_add_diag(int64_t n, double* __restrict__ A, int64_t lda, double v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) A[i * lda + i] += v;
}

// out = a*X + b*I - c*W  style helpers for the IES step
__global__ void k_ies_grad_b(int64_t N, double* __restrict__ G, const double* __restrict__ W,
                             double nm1) {
    // G += (N-1) (I - W)
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * N) return;
    const int64_t r = i / N, c = i % N;
    G[i] += nm1 * ((r == c ? 1.0 : 0.0) - W[i]);
}
__global__ void k_axpy(int64_t n, double a, const double* __restrict__ x, double* __restrict__ y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = fma(a, x[i], y[i]);
}
__global__ void k_ies_resid(int64_t N, int64_t p, const double* __restrict__ y,
                            const double* __restrict__ Dp, const double* __restrict__ Eow,
                            double* __restrict__ out) {
    // out = y - Dp - Eo_w
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * p) return;
    out[i] = y[i % p] - Dp[i] - Eow[i];
}
__global__ void k_set_identity(int64_t n, double* __restrict__ A) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * n) return;
    A[i] = (i / n == i % n) ? 1.0 : 0.0;
}

// ---- taper (tools/localization.py:9-92) ----------------------------------------------------------
__global__ void k_taper_bump(int64_t M, int64_t p, const double* __restrict__ xy_prm,
                             const double* __restrict__ xy_obs, double radius, double sharp,
                             double* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * p) return;
    const int64_t i = idx / p, j = idx % p;
    const double dx = xy_prm[2 * i] - xy_obs[2 * j], dy = xy_prm[2 * i + 1] - xy_obs[2 * j + 1];
    const double x = sqrt(dx * dx + dy * dy) / radius;
    double v = 0.0;
    if (fabs(x) < 1.0) {
        v = exp(1.0 - 1.0 / (1.0 - x * x));
        if (sharp != 1.0) v = pow(v, sharp);
    }
    out[idx] = v;
}

// ---- batched local analysis (HistoryMatch.py:783-793) ----------------------------------------
// One CTA per parameter i.  Inputs: A = S^T S (p,p), B = S^T E (p,M) whose
// column i is this parameter's right-hand side.  Builds the tapered system
// Ci = (c c^T) o A[jj,jj] + (N-1) I on the active observations jj
// (sqrt(taper) > 1e-2), Cholesky-factorises it in shared memory, solves and
// writes w~ = c o Ci^-1 (c o B[jj,i]) back into column i of B (zeros on
// inactive rows), so that the caller finishes with E += D B.
__global__ void __launch_bounds__(256)
k_local_analysis(int64_t M, int p, double nm1, const double* __restrict__ A,
                 const double* __restrict__ taper, double* __restrict__ B, int64_t ldB,
                 int* __restrict__ fail, double* gws, size_t gws_stride) {
    // Workspace: dynamic shared memory where the packed triangle of the tapered p x p system fits (p <= 236), otherwise this CTA's slab of
    // a global-memory workspace (L1 / L2 resident; the CTA barrier orders its accesses) - same algorithm, no size limit.
    extern __shared__ double sm_dyn[];
    double* sm = gws ? gws + blockIdx.x * gws_stride : sm_dyn;
    for (int64_t i = blockIdx.x; i < M; i += gridDim.x) {
    double* c = sm;           // [p] sqrt(taper) of active obs
    double* rhs = sm + p;     // [p]
    int* idx = (int*)(sm + 2 * p);  // [p] active obs indices
    double* colk = sm + 2 * p + (p + 1) / 2;  // [p] the current column of the factor, contiguous (conflict-free reads)
    double* L = colk + p;  // lower triangle, packed by rows: (a, b <= a) at a (a + 1) / 2 + b
    __shared__ int s_n;
    __shared__ int s_fail;
    const int tid = threadIdx.x, nt = blockDim.x;
    auto tri = [](int a, int b) { return a * (a + 1) / 2 + b; };

    // ordered compaction of the active set (ascending j, like boolean indexing)
    if (tid == 0) {
        int n = 0;
        for (int j = 0; j < p; ++j) {
            const double cj = sqrt(taper[i * p + j]);
            if (cj > 1e-2) {
                idx[n] = j;
                c[n] = cj;
                ++n;
            }
        }
        s_n = n;
        s_fail = 0;
    }
    __syncthreads();
    const int n = s_n;
    for (int a = tid; a < n; a += nt) rhs[a] = c[a] * B[(int64_t)idx[a] * ldB + i];
    // rows of the triangle go to warps, the columns b <= a of a row to lanes: no integer division in the loops
    const int lane = tid & 31, wrp = tid >> 5, nwarp = nt >> 5;
    for (int a = wrp; a < n; a += nwarp) {
        const double ca = c[a];
        const double* Arow = A + (int64_t)idx[a] * p;
        double* La = L + tri(a, 0);
        for (int b = lane; b <= a; b += 32) La[b] = ca * c[b] * Arow[idx[b]] + (a == b ? nm1 : 0.0);
    }
    __syncthreads();
    // right-looking Cholesky
    for (int k = 0; k < n; ++k) {
        const double dkk = L[tri(k, k)];
        if (tid == 0 && !(dkk > 0.0)) s_fail = 1;
        const double d = sqrt(dkk);
        __syncthreads();
        for (int a = k + tid; a < n; a += nt) {
            const double v = (a == k) ? d : L[tri(a, k)] / d;
            L[tri(a, k)] = v;
            colk[a] = v;
        }
        __syncthreads();
        for (int a = k + 1 + wrp; a < n; a += nwarp) {
            const double lak = colk[a];
            double* La = L + tri(a, 0);
            for (int b = k + 1 + lane; b <= a; b += 32) La[b] -= lak * colk[b];
        }
        __syncthreads();
    }
    // forward / backward substitution (warp 0; columns processed in order)
    if (tid < 32) {
        for (int k = 0; k < n; ++k) {
            const double yk = rhs[k] / L[tri(k, k)];
            __syncwarp();
            if (tid == 0) rhs[k] = yk;
            for (int a = k + 1 + tid; a < n; a += 32) rhs[a] -= L[tri(a, k)] * yk;
            __syncwarp();
        }
        for (int k = n - 1; k >= 0; --k) {
            const double wk = rhs[k] / L[tri(k, k)];
            __syncwarp();
            if (tid == 0) rhs[k] = wk;
            for (int a = tid; a < k; a += 32) rhs[a] -= L[tri(k, a)] * wk;
            __syncwarp();
        }
    }
    __syncthreads();
    for (int j = tid; j < p; j += nt) B[(int64_t)j * ldB + i] = 0.0;
    __syncthreads();
    for (int a = tid; a < n; a += nt) B[(int64_t)idx[a] * ldB + i] = c[a] * rhs[a];
    if (tid == 0 && s_fail) atomicExch(fail, 1);
    __syncthreads();  // the workspace is reused by this CTA's next parameter
    }
}

// ---- batched localised IES step (HistoryMatch.py:1034-1059) -------------------------------------
// One CTA per parameter i with its own N x N weight matrix Wi.  Everything is in shared memory:
//   Ai = Wi^-1 (in-place Gauss-Jordan with partial pivoting; the reference uses pinv of the square,
//   non-singular Wi), Y0 = center(Ai) Si, G = Di Y0^T + (N-1)(I - Wi), C = Y0 Y0^T + (N-1) I
//   (= the inverse of the reference's SVD expression for covw), Cholesky of C, dW = G C^-1,
//   Wi += xStep dW.  Si / Di are the tapered, active columns of S / D as in k_local_analysis.
__global__ void __launch_bounds__(1024, 1)
k_iles_step(int N, int64_t M, int p, double xStep, const double* __restrict__ S, const double* __restrict__ D,
            const double* __restrict__ taper, double* __restrict__ Ws, int* __restrict__ fail, double* gws,
            size_t gws_stride) {
    // Workspace (3 N^2 + N p doubles): dynamic shared memory where it fits (N <= ~72 at p = 160), otherwise this CTA's slab
    // of a global-memory workspace - same algorithm, no size limit (the notebook's N = 200 runs this way).
    extern __shared__ double sm_dyn[];
    double* sm = gws ? gws + blockIdx.x * gws_stride : sm_dyn;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wrp = tid >> 5, nwarp = nt >> 5;
    for (int64_t ip = blockIdx.x; ip < M; ip += gridDim.x) {
    __syncthreads();  // the workspace and the flags are reused by this CTA's next parameter
    double* A = sm;                 // N*N  : Wi -> Wi^-1 -> centred
    double* G = A + N * N;          // N*N
    double* Cc = G + N * N;         // N*N
    double* Y0 = Cc + N * N;        // N*p: Y0 transposed, n rows (active observations) of pitch N
    double* c = Y0 + (size_t)N * p; // p
    double* colmean = c + p;        // N
    int* idx = (int*)(colmean + N); // p
    int* ipiv = idx + p;            // N
    __shared__ int s_n, s_piv, s_fail;
    double* W = Ws + ip * N * N;
    const double nm1 = (double)(N - 1);

    if (tid == 0) {
        int n = 0;
        for (int j = 0; j < p; ++j) {
            const double cj = sqrt(taper[ip * p + j]);
            if (cj > 1e-2) {
                idx[n] = j;
                c[n] = cj;
                ++n;
            }
        }
        s_n = n;
        s_fail = 0;
    }
    __syncthreads();
    const int n = s_n;
    if (n == 0) continue;  // no active observation: dW = 0 (HistoryMatch.py:1039-1040)
    for (int e = tid; e < N * N; e += nt) A[e] = W[e];
    __syncthreads();
    // in-place Gauss-Jordan inversion with partial pivoting
    for (int k = 0; k < N; ++k) {
        if (tid < 32) {
            double best = -1.0;
            int bi = k;
            for (int i = k + tid; i < N; i += 32) {
                const double v = fabs(A[i * N + k]);
                if (v > best) {
                    best = v;
                    bi = i;
                }
            }
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) {
                    best = ob;
                    bi = oi;
                }
            }
            if (tid == 0) {
                s_piv = bi;
                ipiv[k] = bi;
                if (!(best > 0.0)) s_fail = 1;
            }
        }
        __syncthreads();
        const int pv = s_piv;
        if (pv != k)
            for (int j = tid; j < N; j += nt) {
                const double t0 = A[k * N + j];
                A[k * N + j] = A[pv * N + j];
                A[pv * N + j] = t0;
            }
        __syncthreads();
        const double dinv = 1.0 / A[k * N + k];
        __syncthreads();
        for (int j = tid; j < N; j += nt) A[k * N + j] = (j == k) ? dinv : A[k * N + j] * dinv;
        __syncthreads();
        // column k of the other rows becomes -f * dinv, the rest is eliminated; the multipliers f are stashed first (they
        // are overwritten).  Rows go to warps, columns to lanes: no integer division in the loops.
        for (int i = tid; i < N; i += nt) Cc[i] = A[i * N + k];
        __syncthreads();
        for (int i = wrp; i < N; i += nwarp) {
            if (i == k) continue;
            const double f = Cc[i];
            double* Ai = A + i * N;
            const double* Ak = A + k * N;
            for (int j = lane; j < N; j += 32) Ai[j] = (j == k) ? -f * dinv : Ai[j] - f * Ak[j];
        }
        __syncthreads();
    }
    for (int k = N - 1; k >= 0; --k) {  // undo the row swaps as column swaps
        const int pv = ipiv[k];
        if (pv != k)
            for (int i = tid; i < N; i += nt) {
                const double t0 = A[i * N + k];
                A[i * N + k] = A[i * N + pv];
                A[i * N + pv] = t0;
            }
        __syncthreads();
    }
    // centre the columns of Wi^-1 (axis 0)
    for (int j = tid; j < N; j += nt) {
        double sacc = 0.0;
        for (int i = 0; i < N; ++i) sacc += A[i * N + j];
        colmean[j] = sacc / (double)N;
    }
    __syncthreads();
    for (int i = wrp; i < N; i += nwarp)
        for (int j = lane; j < N; j += 32) A[i * N + j] -= colmean[j];
    __syncthreads();
    // Y0 = A Si,  Si[k][a] = S[k][idx[a]] c[a]; kept TRANSPOSED (Y0t[a][r], row pitch N) so that the N^2 n loop below reads it
    // with consecutive lanes on consecutive addresses (the [r][a] layout made every lane walk its own row: one 32-byte
    // sector per lane and load - with the workspace in global memory that was most of the N = 200 step)
    for (int r = wrp; r < N; r += nwarp)
        for (int a = lane; a < n; a += 32) {
            double acc = 0.0;
            for (int k = 0; k < N; ++k) acc = fma(A[r * N + k], S[(int64_t)k * p + idx[a]], acc);
            Y0[a * N + r] = acc * c[a];
        }
    __syncthreads();
    // G = Di Y0^T + (N-1)(I - Wi), kept transposed (Gt[l][r]) for the solves ;  C = Y0 Y0^T + (N-1) I
    for (int r = wrp; r < N; r += nwarp)
        for (int l = lane; l < N; l += 32) {
            double g = 0.0, cc = 0.0;
            for (int a = 0; a < n; ++a) {
                const double y = Y0[a * N + l];
                g = fma(D[(int64_t)r * p + idx[a]] * c[a], y, g);
                cc = fma(Y0[a * N + r], y, cc);
            }
            G[l * N + r] = g + nm1 * ((r == l ? 1.0 : 0.0) - W[r * N + l]);
            Cc[r * N + l] = cc + (r == l ? nm1 : 0.0);
        }
    __syncthreads();
    // Cholesky of C (lower, in place); the scaled column k is also kept contiguous (colk) for the trailing update
    double* colk = colmean;  // the column means are no longer needed
    for (int k = 0; k < N; ++k) {
        const double dkk = Cc[k * N + k];
        if (tid == 0 && !(dkk > 0.0)) s_fail = 1;
        const double d = sqrt(dkk);
        __syncthreads();
        for (int a = k + tid; a < N; a += nt) {
            const double v = (a == k) ? d : Cc[a * N + k] / d;
            Cc[a * N + k] = v;
            colk[a] = v;
        }
        __syncthreads();
        for (int a = k + 1 + wrp; a < N; a += nwarp) {
            const double cak = colk[a];
            for (int b = k + 1 + lane; b <= a; b += 32) Cc[a * N + b] -= cak * colk[b];
        }
        __syncthreads();
    }
    // dW rows: solve C x = G[r,:]^T, one row r per thread = column r of Gt (consecutive threads on consecutive addresses)
    for (int r = tid; r < N; r += nt) {
        double* g = G + r;
        for (int k = 0; k < N; ++k) {
            double v = g[k * N];
            for (int a = 0; a < k; ++a) v -= Cc[k * N + a] * g[a * N];
            g[k * N] = v / Cc[k * N + k];
        }
        for (int k = N - 1; k >= 0; --k) {
            double v = g[k * N];
            for (int a = k + 1; a < N; ++a) v -= Cc[a * N + k] * g[a * N];
            g[k * N] = v / Cc[k * N + k];
        }
    }
    __syncthreads();
    for (int r = wrp; r < N; r += nwarp)
        for (int l = lane; l < N; l += 32) W[r * N + l] = fma(xStep, G[l * N + r], W[r * N + l]);
    if (tid == 0 && s_fail) atomicExch(fail, 1);
    }
}

// E[:, i] = x0[i] + Ws[i] X0[:, i]   (recompose, HistoryMatch.py:1020-1021)
__global__ void k_iles_recompose(int N, int64_t M, const double* __restrict__ Ws, const double* __restrict__ X0,
                                 const double* __restrict__ x0, double* __restrict__ E) {
    const int64_t i = blockIdx.x;
    const double* W = Ws + i * N * N;
    for (int r = threadIdx.x; r < N; r += blockDim.x) {
        double acc = 0.0;
        for (int k = 0; k < N; ++k) acc = fma(W[r * N + k], X0[(int64_t)k * M + i], acc);
        E[(int64_t)r * M + i] = x0[i] + acc;
    }
}

// S = center(Eo) decorr, D = (obs - Eo - perturbs) decorr (HistoryMatch.py:580-584)
int whiten(hm_ctx* ctx, int64_t N, int64_t p, const double* Eo, const double* obs,
           const double* perturbs, const double* decorr, double** S_out, double** D_out) {
    cudaStream_t st = ctx->stream;
    double *tmp, *mean, *S, *D;
    HM_CHECK(ctx->ws.get("an.tmp", (size_t)N * p, &tmp));
    HM_CHECK(ctx->ws.get("an.mean", (size_t)p, &mean));
    HM_CHECK(ctx->ws.get("an.S", (size_t)N * p, &S));
    HM_CHECK(ctx->ws.get("an.D", (size_t)N * p, &D));
    const int tb = 128;
    k_col_mean<<<(unsigned)((p + tb - 1) / tb), tb, 0, st>>>(N, p, Eo, p, mean);
    k_sub_mean<<<dim3((unsigned)((p + tb - 1) / tb), (unsigned)N), tb, 0, st>>>(N, p, Eo, p, tmp, p, mean, 1.0);
    HM_CHECK(hm::dgemm(ctx, false, false, N, p, p, 1.0, tmp, p, decorr, p, 0.0, S, p));
    k_innovation<<<(unsigned)((N * p + 255) / 256), 256, 0, st>>>(N, p, obs, Eo, perturbs, tmp);
    ctx->launches += 3;
    HM_CHECK(hm::dgemm(ctx, false, false, N, p, p, 1.0, tmp, p, decorr, p, 0.0, D, p));
    *S_out = S;
    *D_out = D;
    return HM_OK;
}

int check_info(hm_ctx* ctx, int* d_info, const char* what) {
    HM_CUDA(cudaMemcpyAsync(ctx->h_pinned, d_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    HM_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_pinned[0] != 0) {
        hm::set_error("%s failed: info = %d", what, ctx->h_pinned[0]);
        return HM_ERR_NUMERIC;
    }
    return HM_OK;
}

// check the first `count` info words of "an.info" with ONE stream synchronisation (deferred checks of a call sequence)
int check_infos(hm_ctx* ctx, int count, const char* what) {
    int* info;
    HM_CHECK(ctx->ws.get("an.info", (size_t)4, &info));
    HM_CUDA(cudaMemcpyAsync(ctx->h_pinned, info, count * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    HM_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < count; ++i)
        if (ctx->h_pinned[i] != 0) {
            hm::set_error("%s failed: info[%d] = %d", what, i, ctx->h_pinned[i]);
            return HM_ERR_NUMERIC;
        }
    return HM_OK;
}

// Cholesky factor of the SPD n x n matrix A (row-major == column-major by symmetry), in place.  slot >= 0: the status
// goes to word `slot` of "an.info" and is NOT checked here (no host synchronisation): the caller ends with check_infos.
int chol_factor(hm_ctx* ctx, int n, double* A, int slot = -1) {
    HM_CHECK(ensure_solver(ctx));
    int lwork = 0;
    HM_CUSOLVER(cusolverDnDpotrf_bufferSize(ctx->solver, CUBLAS_FILL_MODE_LOWER, n, A, n, &lwork));
    double* work;
    int* info;
    HM_CHECK(ctx->ws.get("an.potrf_work", (size_t)std::max(lwork, 1), &work));
    HM_CHECK(ctx->ws.get("an.info", (size_t)4, &info));
    HM_CUSOLVER(cusolverDnDpotrf(ctx->solver, CUBLAS_FILL_MODE_LOWER, n, A, n, work, lwork, info + std::max(slot, 0)));
    return slot >= 0 ? HM_OK : check_info(ctx, info, "Cholesky factorisation (potrf)");
}
// In place: Brow (nrhs x n, row-major) <- Brow A^-1, using the factor from chol_factor.
// (column-major view of Brow is n x nrhs = Brow^T, and A^-1 Brow^T = (Brow A^-1)^T.)
int chol_solve_right(hm_ctx* ctx, int n, const double* Afac, int nrhs, double* Brow, int slot = -1) {
    int* info;
    HM_CHECK(ctx->ws.get("an.info", (size_t)4, &info));
    HM_CUSOLVER(cusolverDnDpotrs(ctx->solver, CUBLAS_FILL_MODE_LOWER, n, nrhs, Afac, n, Brow, n, info + std::max(slot, 0)));
    return slot >= 0 ? HM_OK : check_info(ctx, info, "Cholesky solve (potrs)");
}

}  // namespace

extern "C" int hm_center(hm_ctx* ctx, int64_t N, int64_t M, const double* E, int64_t ldE, double* X,
                         int64_t ldX, double* mean, int rescale) {
    HM_REQUIRE(ctx && E && X, "null pointer");
    HM_REQUIRE(N > 0 && M > 0, "empty ensemble");
    HM_CUDA(cudaSetDevice(ctx->device));
    double* mu = mean;
    if (!mu) HM_CHECK(ctx->ws.get("an.center_mean", (size_t)M, &mu));
    const int tb = 128;
    const unsigned gx = (unsigned)((M + tb - 1) / tb);
    k_col_mean<<<gx, tb, 0, ctx->stream>>>(N, M, E, ldE, mu);
    const double scale = (rescale && N > 1) ? sqrt((double)N / (double)(N - 1)) : 1.0;
    k_sub_mean<<<dim3(gx, (unsigned)N), tb, 0, ctx->stream>>>(N, M, E, ldE, X, ldX, mu, scale);
    ctx->launches += 2;
    HM_CUDA(cudaGetLastError());
    return HM_OK;
}

extern "C" int hm_taper_bump(hm_ctx* ctx, int64_t M, int64_t p, const double* xy_prm,
                             const double* xy_obs, double radius, double sharpness, double* out) {
    HM_REQUIRE(ctx && xy_prm && xy_obs && out, "null pointer");
    HM_REQUIRE(radius > 0, "radius > 0");
    HM_CUDA(cudaSetDevice(ctx->device));
    k_taper_bump<<<(unsigned)((M * p + 255) / 256), 256, 0, ctx->stream>>>(M, p, xy_prm, xy_obs, radius,
                                                                            sharpness, out);
    ctx->launches += 1;
    HM_CUDA(cudaGetLastError());
    return HM_OK;
}

extern "C" int hm_es_update(hm_ctx* ctx, int64_t N, int64_t M, int64_t p, double* E, int64_t ldE,
                            const double* Eo, const double* obs, const double* perturbs,
                            const double* decorr) {
    HM_REQUIRE(ctx && E && Eo && obs && perturbs && decorr, "null pointer");
    HM_REQUIRE(N > 1 && M > 0 && p > 0 && ldE >= M, "shape");
    HM_CUDA(cudaSetDevice(ctx->device));
    double *S, *D, *C, *G;
    HM_CHECK(whiten(ctx, N, p, Eo, obs, perturbs, decorr, &S, &D));
    HM_CHECK(ctx->ws.get("an.C", (size_t)p * p, &C));
    HM_CHECK(ctx->ws.get("an.G", (size_t)p * M, &G));
    // C = S^T S + (N-1) I
    HM_CHECK(hm::dgemm(ctx, true, false, p, p, N, 1.0, S, p, S, p, 0.0, C, p));
    k_add_diag<<<(unsigned)((p + 127) / 128), 128, 0, ctx->stream>>>(p, C, p, (double)(N - 1));
    ctx->launches += 1;
    // one host synchronisation for the whole update (the factorisation status is checked at the end): the update is a
    // dozen short launches, and every host round trip between them is exposed to the host's scheduling noise
    HM_CHECK(chol_factor(ctx, (int)p, C, 0));
    // D <- D C^-1
    HM_CHECK(chol_solve_right(ctx, (int)p, C, (int)N, D, 1));
    // G = S^T E ; E += (D C^-1) G
    HM_CHECK(hm::dgemm(ctx, true, false, p, M, N, 1.0, S, p, E, ldE, 0.0, G, M));
    HM_CHECK(hm::dgemm(ctx, false, false, N, M, p, 1.0, D, p, G, M, 1.0, E, ldE));
    return check_infos(ctx, 2, "ES update: Cholesky factorisation / solve of S^T S + (N-1) I");
}

extern "C" int hm_es_update_host(hm_ctx* ctx, int64_t N, int64_t M, int64_t p, double* E,
                                 const double* Eo, const double* obs, const double* perturbs,
                                 const double* decorr) {
    HM_REQUIRE(ctx && E && Eo && obs && perturbs && decorr, "null pointer");
    HM_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    double *dE, *dEo, *dobs, *dpert, *ddec;
    HM_CHECK(ctx->ws.get("h.E", (size_t)N * M, &dE));
    HM_CHECK(ctx->ws.get("h.Eo", (size_t)N * p, &dEo));
    HM_CHECK(ctx->ws.get("h.obsv", (size_t)p, &dobs));
    HM_CHECK(ctx->ws.get("h.pert", (size_t)N * p, &dpert));
    HM_CHECK(ctx->ws.get("h.decorr", (size_t)p * p, &ddec));
    HM_CUDA(cudaMemcpyAsync(dE, E, (size_t)N * M * 8, cudaMemcpyHostToDevice, st));
    HM_CUDA(cudaMemcpyAsync(dEo, Eo, (size_t)N * p * 8, cudaMemcpyHostToDevice, st));
    HM_CUDA(cudaMemcpyAsync(dobs, obs, (size_t)p * 8, cudaMemcpyHostToDevice, st));
    HM_CUDA(cudaMemcpyAsync(dpert, perturbs, (size_t)N * p * 8, cudaMemcpyHostToDevice, st));
    HM_CUDA(cudaMemcpyAsync(ddec, decorr, (size_t)p * p * 8, cudaMemcpyHostToDevice, st));
    HM_CHECK(hm_es_update(ctx, N, M, p, dE, M, dEo, dobs, dpert, ddec));
    HM_CUDA(cudaMemcpyAsync(E, dE, (size_t)N * M * 8, cudaMemcpyDeviceToHost, st));
    HM_CUDA(cudaStreamSynchronize(st));
    return HM_OK;
}

extern "C" int hm_les_update(hm_ctx* ctx, int64_t N, int64_t M, int64_t p, double* E, int64_t ldE,
                             const double* Eo, const double* obs, const double* perturbs,
                             const double* decorr, const double* taper) {
    HM_REQUIRE(ctx && E && Eo && obs && perturbs && decorr && taper, "null pointer");
    HM_REQUIRE(N > 1 && M > 0 && p > 0 && ldE >= M, "shape");
    HM_CUDA(cudaSetDevice(ctx->device));
    size_t smem = ((size_t)3 * p + (p + 1) / 2 + (size_t)p * (p + 1) / 2) * sizeof(double);  // packed lower triangle: two CTAs per SM at p = 160
    double *S, *D, *A, *B;
    int* fail;
    HM_CHECK(whiten(ctx, N, p, Eo, obs, perturbs, decorr, &S, &D));
    HM_CHECK(ctx->ws.get("an.C", (size_t)p * p, &A));
    HM_CHECK(ctx->ws.get("an.G", (size_t)p * M, &B));
    HM_CHECK(ctx->ws.get("an.info", (size_t)4, &fail));
    HM_CUDA(cudaMemsetAsync(fail, 0, sizeof(int), ctx->stream));
    HM_CHECK(hm::dgemm(ctx, true, false, p, p, N, 1.0, S, p, S, p, 0.0, A, p));
    HM_CHECK(hm::dgemm(ctx, true, false, p, M, N, 1.0, S, p, E, ldE, 0.0, B, M));
    double* gws = nullptr;
    size_t gstride = 0;
    unsigned grid = (unsigned)M;
    if (smem > kMaxWorkSmem) {  // the tapered system does not fit shared memory: global-memory workspace, persistent CTAs
        gstride = (smem / sizeof(double) + 15) & ~(size_t)15;
        grid = (unsigned)std::min<int64_t>(M, (int64_t)ctx->sm_count * 4);
        HM_CHECK(ctx->ws.get("an.loc_ws", gstride * grid, &gws));
        smem = 0;
    }
    HM_CUDA(cudaFuncSetAttribute(k_local_analysis, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_local_analysis<<<grid, 256, smem, ctx->stream>>>(M, (int)p, (double)(N - 1), A, taper, B, M, fail, gws, gstride);
    ctx->launches += 1;
    HM_CUDA(cudaGetLastError());
    HM_CHECK(hm::dgemm(ctx, false, false, N, M, p, 1.0, D, p, B, M, 1.0, E, ldE));
    return check_info(ctx, fail, "local analysis Cholesky");
}

extern "C" int hm_ies_step(hm_ctx* ctx, int64_t N, int64_t p, double* W, const double* Eo,
                           const double* obs, const double* perturbs, const double* decorr,
                           double xStep) {
    HM_REQUIRE(ctx && W && Eo && obs && perturbs && decorr, "null pointer");
    HM_REQUIRE(N > 1 && p > 0, "shape");
    HM_CUDA(cudaSetDevice(ctx->device));
    HM_CHECK(ensure_solver(ctx));
    cudaStream_t st = ctx->stream;
    double *y, *Dp, *Eow, *LU, *Winv, *mean, *Y0, *res, *G, *Cw, *work;
    int *ipiv, *info;
    HM_CHECK(ctx->ws.get("ies.y", (size_t)p, &y));
    HM_CHECK(ctx->ws.get("ies.Dp", (size_t)N * p, &Dp));
    HM_CHECK(ctx->ws.get("ies.Eow", (size_t)N * p, &Eow));
    HM_CHECK(ctx->ws.get("ies.LU", (size_t)N * N, &LU));
    HM_CHECK(ctx->ws.get("ies.Winv", (size_t)N * N, &Winv));
    HM_CHECK(ctx->ws.get("ies.mean", (size_t)N, &mean));
    HM_CHECK(ctx->ws.get("ies.Y0", (size_t)N * p, &Y0));
    HM_CHECK(ctx->ws.get("ies.res", (size_t)N * p, &res));
    HM_CHECK(ctx->ws.get("ies.G", (size_t)N * N, &G));
    HM_CHECK(ctx->ws.get("ies.Cw", (size_t)N * N, &Cw));
    HM_CHECK(ctx->ws.get("ies.ipiv", (size_t)N, &ipiv));
    HM_CHECK(ctx->ws.get("an.info", (size_t)4, &info));
    // y = obs decorr ; Dp = perturbs decorr ; Eo_w = Eo decorr   (HistoryMatch.py:911-912, 927)
    HM_CHECK(hm::dgemm(ctx, false, false, 1, p, p, 1.0, obs, p, decorr, p, 0.0, y, p));
    HM_CHECK(hm::dgemm(ctx, false, false, N, p, p, 1.0, perturbs, p, decorr, p, 0.0, Dp, p));
    HM_CHECK(hm::dgemm(ctx, false, false, N, p, p, 1.0, Eo, p, decorr, p, 0.0, Eow, p));
    // W^-1 by LU (pinv(W) of the reference; W is square and non-singular).  Row-major W is
    // column-major W^T and inv(W^T) read back row-major is inv(W).
    HM_CUDA(cudaMemcpyAsync(LU, W, (size_t)N * N * 8, cudaMemcpyDeviceToDevice, st));
    int lwork = 0;
    HM_CUSOLVER(cusolverDnDgetrf_bufferSize(ctx->solver, (int)N, (int)N, LU, (int)N, &lwork));
    HM_CHECK(ctx->ws.get("ies.getrf_work", (size_t)std::max(lwork, 1), &work));
    HM_CUSOLVER(cusolverDnDgetrf(ctx->solver, (int)N, (int)N, LU, (int)N, work, ipiv, info));
    HM_CHECK(check_info(ctx, info, "LU factorisation of W (getrf)"));
    k_set_identity<<<(unsigned)((N * N + 255) / 256), 256, 0, st>>>(N, Winv);
    HM_CUSOLVER(cusolverDnDgetrs(ctx->solver, CUBLAS_OP_N, (int)N, (int)N, LU, (int)N, ipiv, Winv, (int)N, info));
    HM_CHECK(check_info(ctx, info, "LU solve (getrs)"));
    // Y0 = center(W^-1) Eo_w   (HistoryMatch.py:928)
    HM_CHECK(hm_center(ctx, N, N, Winv, N, Winv, N, mean, 0));
    HM_CHECK(hm::dgemm(ctx, false, false, N, p, N, 1.0, Winv, N, Eow, p, 0.0, Y0, p));
    // grad = (y - Dp - Eo_w) Y0^T + (N-1)(I - W)   (HistoryMatch.py:931-932)
    k_ies_resid<<<(unsigned)((N * p + 255) / 256), 256, 0, st>>>(N, p, y, Dp, Eow, res);
    HM_CHECK(hm::dgemm(ctx, false, true, N, N, p, 1.0, res, p, Y0, p, 0.0, G, N));
    k_ies_grad_b<<<(unsigned)((N * N + 255) / 256), 256, 0, st>>>(N, G, W, (double)(N - 1));
    // covw = (Y0 Y0^T + (N-1) I)^-1   (HistoryMatch.py:935-938, via the SVD there)
    HM_CHECK(hm::dgemm(ctx, false, true, N, N, p, 1.0, Y0, p, Y0, p, 0.0, Cw, N));
    k_add_diag<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(N, Cw, N, (double)(N - 1));
    HM_CHECK(chol_factor(ctx, (int)N, Cw));
    // dW = grad covw ; W += xStep dW   (HistoryMatch.py:941-942)
    HM_CHECK(chol_solve_right(ctx, (int)N, Cw, (int)N, G));
    k_axpy<<<(unsigned)((N * N + 255) / 256), 256, 0, st>>>(N * N, xStep, G, W);
    ctx->launches += 5;  // set_identity, ies_resid, ies_grad_b, add_diag, axpy
    HM_CUDA(cudaGetLastError());
    return HM_OK;
}

extern "C" int hm_iles_step(hm_ctx* ctx, int64_t N, int64_t M, int64_t p, double* Ws, const double* Eo,
                            const double* obs, const double* perturbs, const double* decorr,
                            const double* taper, double xStep) {
    HM_REQUIRE(ctx && Ws && Eo && obs && perturbs && decorr && taper, "null pointer");
    HM_REQUIRE(N > 1 && M > 0 && p > 0, "shape");
    HM_CUDA(cudaSetDevice(ctx->device));
    size_t smem = ((size_t)3 * N * N + (size_t)N * p + p + N) * sizeof(double) + (size_t)(p + N) * sizeof(int);
    double *S, *D;
    int* fail;
    HM_CHECK(whiten(ctx, N, p, Eo, obs, perturbs, decorr, &S, &D));
    HM_CHECK(ctx->ws.get("an.info", (size_t)4, &fail));
    HM_CUDA(cudaMemsetAsync(fail, 0, sizeof(int), ctx->stream));
    double* gws = nullptr;
    size_t gstride = 0;
    unsigned grid = (unsigned)M, block = 256;
    if (smem > kMaxWorkSmem) {  // 3 N^2 + N p doubles do not fit shared memory: global-memory workspace, persistent CTAs
        gstride = (smem / sizeof(double) + 15) & ~(size_t)15;
        // 1024 threads, one resident CTA per SM: the N x N sweeps of a column step are spread over four times the threads
        // and the resident matrices (148 x 3 N^2 doubles) stay in L2.  Measured at N = 200, M = 400, p = 160 (ms per step):
        // 256 threads x 4 CTAs per SM 32.9, 512 x 2 26.1, 1024 x 1 (grid = SMs / 2 x SMs) 21.1 / 19.6
        block = 1024;
        grid = (unsigned)std::min<int64_t>(M, (int64_t)ctx->sm_count * 2);
        HM_CHECK(ctx->ws.get("an.loc_ws", gstride * grid, &gws));
        smem = 0;
    }
    HM_CUDA(cudaFuncSetAttribute(k_iles_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_iles_step<<<grid, block, smem, ctx->stream>>>((int)N, M, (int)p, xStep, S, D, taper, Ws, fail, gws, gstride);
    ctx->launches += 1;
    HM_CUDA(cudaGetLastError());
    return check_info(ctx, fail, "localised IES step (singular Wi or non-SPD Gauss-Newton matrix)");
}

extern "C" int hm_iles_recompose(hm_ctx* ctx, int64_t N, int64_t M, const double* Ws, const double* X0,
                                 const double* x0, double* E) {
    HM_REQUIRE(ctx && Ws && X0 && x0 && E, "null pointer");
    HM_CUDA(cudaSetDevice(ctx->device));
    k_iles_recompose<<<(unsigned)M, 64, 0, ctx->stream>>>((int)N, M, Ws, X0, x0, E);
    ctx->launches += 1;
    HM_CUDA(cudaGetLastError());
    return HM_OK;
}

// ---- ensemble covariance / correlation fields (utils.cov / utils.corr, tools/utils.py:31-55) ----------
// out (M,q) = center(a)^T center(b) / (N-1) [ / std(a) / std(b), clipped to +-999 ].  center(b) has zero
// column sums, so a itself need not be centred for the product.  Streaming and HBM bound: `a` (N,M) is read
// twice (two-pass variance, as numpy) by one thread per column, coalesced across columns; the correlated
// series b is small (q columns, usually 1: the dashboards correlate a field with one well observation,
// HistoryMatch.py:478-482, 738-748, 829-833) and is read through the read-only path as a warp broadcast.
namespace {

constexpr int kCorrQ = 8;  // columns of b handled per pass over a

// Bc = center(b) (N,q) dense; sb[t] = sample standard deviation of column t (ddof = 1)
__global__ void k_corr_prep(int64_t N, int64_t q, const double* __restrict__ b, int64_t ldb,
                            double* __restrict__ Bc, double* __restrict__ sb) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= q) return;
    double s = 0.0;
    for (int64_t i = 0; i < N; ++i) s += b[i * ldb + t];
    const double mu = s / (double)N;
    double v = 0.0;
    for (int64_t i = 0; i < N; ++i) {
        const double d = b[i * ldb + t] - mu;
        Bc[i * q + t] = d;
        v = fma(d, d, v);
    }
    sb[t] = sqrt(v / (double)(N - 1));
}

// A block owns 32 adjacent columns; its 8 warps split the ensemble rows (row i goes to warp i mod 8), so 8 x as many
// loads are in flight as with one thread per column, and the second pass over the block's 32-column slab comes from L2.
// Partial sums are combined through shared memory in a fixed order (deterministic).
constexpr int kCorrSlices = 8;

__global__ void __launch_bounds__(32 * kCorrSlices)
k_corr_fields(int64_t N, int64_t M, int64_t q, int64_t t0, const double* __restrict__ a, int64_t lda,
              const double* __restrict__ Bc, const double* __restrict__ sb, double* __restrict__ out,
              int corr) {
    __shared__ double red[kCorrSlices][kCorrQ + 1][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int64_t j = (int64_t)blockIdx.x * 32 + cx;
    const bool live = j < M;
    const int nq = (int)min((int64_t)kCorrQ, q - t0);
    const double* col = a + (live ? j : 0);
    double s0 = 0.0, s1 = 0.0;
    int64_t i = ry;
    for (; i + kCorrSlices < N; i += 2 * kCorrSlices) {
        s0 += col[i * lda];
        s1 += col[(i + kCorrSlices) * lda];
    }
    if (i < N) s0 += col[i * lda];
    red[ry][0][cx] = s0 + s1;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < kCorrSlices; ++r) s += red[r][0][cx];
    const double mu = s / (double)N;
    __syncthreads();
    double v = 0.0, acc[kCorrQ];
#pragma unroll
    for (int t = 0; t < kCorrQ; ++t) acc[t] = 0.0;
    for (i = ry; i < N; i += kCorrSlices) {
        const double d = col[i * lda] - mu;
        v = fma(d, d, v);
#pragma unroll
        for (int t = 0; t < kCorrQ; ++t)
            if (t < nq) acc[t] = fma(d, __ldg(Bc + i * q + t0 + t), acc[t]);
    }
    red[ry][0][cx] = v;
#pragma unroll
    for (int t = 0; t < kCorrQ; ++t) red[ry][t + 1][cx] = acc[t];
    __syncthreads();
    if (ry != 0 || !live) return;
    v = 0.0;
#pragma unroll
    for (int r = 0; r < kCorrSlices; ++r) v += red[r][0][cx];
    const double sa = sqrt(v / (double)(N - 1));
#pragma unroll
    for (int t = 0; t < kCorrQ; ++t) {
        if (t < nq) {
            double c = 0.0;
#pragma unroll
            for (int r = 0; r < kCorrSlices; ++r) c += red[r][t + 1][cx];
            c /= (double)(N - 1);
            if (corr) {
                c = c / sa / sb[t0 + t];
                if (!isnan(c)) c = fmin(fmax(c, -999.0), 999.0);  // 0/0 stays NaN, as with np.clip
            }
            out[j * q + t0 + t] = c;
        }
    }
}

}  // namespace

extern "C" int hm_corr(hm_ctx* ctx, int64_t N, int64_t M, int64_t q, const double* a, int64_t lda,
                       const double* b, int64_t ldb, double* out, int corr) {
    HM_REQUIRE(ctx && a && b && out, "null pointer");
    HM_REQUIRE(N > 1 && M > 0 && q > 0, "need N > 1 members and non-empty a, b");
    HM_REQUIRE(lda >= M && ldb >= q, "row strides");
    HM_CUDA(cudaSetDevice(ctx->device));
    double *Bc, *sb;
    HM_CHECK(ctx->ws.get("an.corr_Bc", (size_t)(N * q), &Bc));
    HM_CHECK(ctx->ws.get("an.corr_sb", (size_t)q, &sb));
    k_corr_prep<<<(unsigned)((q + 63) / 64), 64, 0, ctx->stream>>>(N, q, b, ldb, Bc, sb);
    for (int64_t t0 = 0; t0 < q; t0 += kCorrQ)
        k_corr_fields<<<(unsigned)((M + 31) / 32), 32 * kCorrSlices, 0, ctx->stream>>>(N, M, q, t0, a, lda, Bc, sb, out, corr);
    ctx->launches += 1 + (q + kCorrQ - 1) / kCorrQ;
    HM_CUDA(cudaGetLastError());
    return HM_OK;
}

// ---- strided block copy: the pack / unpack step of the member <-> parameter-column re-sharding --------------------
namespace {
__global__ void __launch_bounds__(256) k_copy2d(int64_t rows, int64_t cols, const double* __restrict__ src, int64_t lds,
                                                double* __restrict__ dst, int64_t ldd) {
    const int64_t n = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols, c = i - r * cols;
        dst[r * ldd + c] = src[r * lds + c];
    }
}
}  // namespace

extern "C" int hm_copy2d(hm_ctx* ctx, int64_t rows, int64_t cols, const double* src, int64_t lds, double* dst,
                         int64_t ldd) {
    HM_REQUIRE(ctx, "null ctx");
    if (rows <= 0 || cols <= 0) return HM_OK;
    HM_REQUIRE(src && dst && lds >= cols && ldd >= cols, "pointers / row strides");
    HM_CUDA(cudaSetDevice(ctx->device));
    const int64_t n = rows * cols;
    const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16);
    k_copy2d<<<grid, 256, 0, ctx->stream>>>(rows, cols, src, lds, dst, ldd);
    ctx->launches += 1;
    HM_CUDA(cudaGetLastError());
    return HM_OK;
}

// ---- swap of the two leading axes of a (d0, d1, d2) array: dst[j][i][:] = src[i][j][:] ------------------------------
namespace {
__global__ void __launch_bounds__(256) k_swap01(int64_t d0, int64_t d1, int64_t d2, const double* __restrict__ src,
                                                double* __restrict__ dst) {
    const int64_t n = d0 * d1 * d2;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = e % d2, ij = e / d2, j = ij % d0, i = ij / d0;  // e indexes dst[i][j][k], i < d1, j < d0
        dst[e] = src[(j * d1 + i) * d2 + k];
    }
}
}  // namespace

extern "C" int hm_swap01(hm_ctx* ctx, int64_t d0, int64_t d1, int64_t d2, const double* src, double* dst) {
    HM_REQUIRE(ctx && src && dst && src != dst, "pointers");
    if (d0 <= 0 || d1 <= 0 || d2 <= 0) return HM_OK;
    HM_CUDA(cudaSetDevice(ctx->device));
    const int64_t n = d0 * d1 * d2;
    const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16);
    k_swap01<<<grid, 256, 0, ctx->stream>>>(d0, d1, d2, src, dst);
    ctx->launches += 1;
    HM_CUDA(cudaGetLastError());
    return HM_OK;
}
