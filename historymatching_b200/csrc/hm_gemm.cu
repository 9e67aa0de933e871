// FP64 tensor-core GEMM for the analysis step (sm_100a).
//
// Replaces the numpy "@" products of the ensemble update
// (HistoryMatch.py:583-586, 920, 928-941).  tcgen05 has no FP64 kind, so the
// FP64 tensor path on Blackwell is the warp-level DMMA
// (mma.sync.aligned.m8n8k4.f64, SASS DMMA.8x8x4).
//
// C[m,n] = alpha * op(A)[m,k] * op(B)[k,n] + beta * C, everything row-major.
// CTA tile (32*WM) x (32*WN) x 16, one warp per 32x32 sub-tile (4x4 DMMA
// tiles, 64 accumulator registers), operands staged in shared memory as
// As[m][k], Bs[n][k] with a row pitch of 20 doubles (conflict-free 64-bit
// fragment loads), register-prefetched double buffering.
#include "hm_common.cuh"

namespace {

constexpr int BK = 16;
constexpr int LDS = BK + 4;

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Element (row i, reduction index kk) of a "rows x k" operand lives at base[i*s_row + kk*s_k].
struct Operand {
    const double* base;
    int64_t s_row, s_k;
    int64_t rows;
};

template <int ROWS, int NT>
__device__ __forceinline__ void load_tile(const Operand& op, int64_t row0, int64_t k0, int64_t K,
                                          double (&reg)[(ROWS * BK + NT - 1) / NT]) {
    constexpr int PER = (ROWS * BK + NT - 1) / NT;
    // thread -> element mapping runs fastest along the unit-stride direction of the operand
    if (op.s_k == 1) {
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int e = threadIdx.x + j * NT;
            const int kk = e % BK, r = e / BK;
            const int64_t gr = row0 + r, gk = k0 + kk;
            reg[j] = (e < ROWS * BK && gr < op.rows && gk < K) ? op.base[gr * op.s_row + gk] : 0.0;
        }
    } else {
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int e = threadIdx.x + j * NT;
            const int r = e % ROWS, kk = e / ROWS;
            const int64_t gr = row0 + r, gk = k0 + kk;
            reg[j] = (e < ROWS * BK && gr < op.rows && gk < K) ? op.base[gr * op.s_row + gk * op.s_k] : 0.0;
        }
    }
}

template <int ROWS, int NT>
__device__ __forceinline__ void store_tile(const Operand& op, double* sm,
                                           const double (&reg)[(ROWS * BK + NT - 1) / NT]) {
    constexpr int PER = (ROWS * BK + NT - 1) / NT;
    if (op.s_k == 1) {
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int e = threadIdx.x + j * NT;
            if (e < ROWS * BK) sm[(e / BK) * LDS + (e % BK)] = reg[j];
        }
    } else {
        // operand whose rows are contiguous in memory (k strided): kept k-major in shared memory, sm[k][row] with a pitch of
        // ROWS + 4 doubles - consecutive lanes store consecutive rows (the [row][k] layout gave 4-way bank conflicts here)
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int e = threadIdx.x + j * NT;
            if (e < ROWS * BK) sm[(e / ROWS) * (ROWS + 4) + (e % ROWS)] = reg[j];
        }
    }
}

// AKM / BKM: the operand is k-major in shared memory (its rows are contiguous in global memory, see store_tile)
template <int WM, int WN, bool AKM, bool BKM>
__global__ void __launch_bounds__(WM * WN * 32)
k_dgemm(Operand A, Operand B, int64_t K, double alpha, double beta, double* __restrict__ C,
        int64_t ldc) {
    constexpr int BM = 32 * WM, BN = 32 * WN, NT = WM * WN * 32;
    extern __shared__ double smem[];
    double* As = smem;                  // [2][BM][LDS]
    double* Bs = smem + 2 * BM * LDS;   // [2][BN][LDS]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp / WN, wn = warp % WN;
    const int64_t row0 = (int64_t)blockIdx.y * BM, col0 = (int64_t)blockIdx.x * BN;
    const int g = lane >> 2, q = lane & 3;

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    double ra[(BM * BK + NT - 1) / NT], rb[(BN * BK + NT - 1) / NT];
    const int64_t nk = (K + BK - 1) / BK;
    load_tile<BM, NT>(A, row0, 0, K, ra);
    load_tile<BN, NT>(B, col0, 0, K, rb);
    store_tile<BM, NT>(A, As, ra);
    store_tile<BN, NT>(B, Bs, rb);
    __syncthreads();
    for (int64_t kt = 0; kt < nk; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nk) {
            load_tile<BM, NT>(A, row0, (kt + 1) * BK, K, ra);
            load_tile<BN, NT>(B, col0, (kt + 1) * BK, K, rb);
        }
        // fragment element (row 8 i + g, k = ks + q): [row][k] layout (pitch LDS) or [k][row] layout (pitch ROWS + 4); both
        // are conflict-free 64-bit loads (per half-warp: 4 values of g x 4 of q on 16 distinct banks)
        const double* as = As + cur * BM * LDS + (AKM ? q * (BM + 4) + wm * 32 + g : (wm * 32 + g) * LDS + q);
        const double* bs = Bs + cur * BN * LDS + (BKM ? q * (BN + 4) + wn * 32 + g : (wn * 32 + g) * LDS + q);
#pragma unroll
        for (int ks = 0; ks < BK; ks += 4) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = AKM ? as[ks * (BM + 4) + i * 8] : as[i * 8 * LDS + ks];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = BKM ? bs[ks * (BN + 4) + j * 8] : bs[j * 8 * LDS + ks];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        if (kt + 1 < nk) {
            store_tile<BM, NT>(A, As + (cur ^ 1) * BM * LDS, ra);
            store_tile<BN, NT>(B, Bs + (cur ^ 1) * BN * LDS, rb);
        }
        __syncthreads();
    }
    // epilogue: lane holds C[g][2q], C[g][2q+1] of every 8x8 tile
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = row0 + wm * 32 + i * 8 + g;
        if (r >= A.rows) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t c = col0 + wn * 32 + j * 8 + 2 * q;
            double* dst = C + r * ldc + c;
            if (c + 1 < B.rows) {
                double v0 = alpha * acc[i][j][0], v1 = alpha * acc[i][j][1];
                if (beta != 0.0) {
                    v0 = fma(beta, dst[0], v0);
                    v1 = fma(beta, dst[1], v1);
                }
                dst[0] = v0;
                dst[1] = v1;
            } else if (c < B.rows) {
                double v0 = alpha * acc[i][j][0];
                if (beta != 0.0) v0 = fma(beta, dst[0], v0);
                dst[0] = v0;
            }
        }
    }
}

template <int WM, int WN, bool AKM, bool BKM>
int launch2(hm_ctx* ctx, const Operand& A, const Operand& B, int64_t K, double alpha, double beta, double* C, int64_t ldc) {
    constexpr int BM = 32 * WM, BN = 32 * WN;
    const size_t smem = (size_t)2 * (BM + BN) * LDS * sizeof(double);
    // per launch: the attribute belongs to the current device's context (a process may drive several GPUs)
    HM_CUDA(cudaFuncSetAttribute(k_dgemm<WM, WN, AKM, BKM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((B.rows + BN - 1) / BN), (unsigned)((A.rows + BM - 1) / BM));
    k_dgemm<WM, WN, AKM, BKM><<<grid, WM * WN * 32, smem, ctx->stream>>>(A, B, K, alpha, beta, C, ldc);
    HM_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return HM_OK;
}

template <int WM, int WN>
int launch(hm_ctx* ctx, const Operand& A, const Operand& B, int64_t K, double alpha, double beta,
           double* C, int64_t ldc) {
    const bool akm = A.s_k != 1, bkm = B.s_k != 1;  // the shared-memory layout follows the operand's unit-stride direction
    if (akm) return bkm ? launch2<WM, WN, true, true>(ctx, A, B, K, alpha, beta, C, ldc)
                        : launch2<WM, WN, true, false>(ctx, A, B, K, alpha, beta, C, ldc);
    return bkm ? launch2<WM, WN, false, true>(ctx, A, B, K, alpha, beta, C, ldc)
               : launch2<WM, WN, false, false>(ctx, A, B, K, alpha, beta, C, ldc);
}

}  // namespace

namespace hm {

// internal entry used by the analysis code
int dgemm(hm_ctx* ctx, bool tA, bool tB, int64_t m, int64_t n, int64_t k, double alpha,
          const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C,
          int64_t ldc) {
    if (m <= 0 || n <= 0) return HM_OK;
    Operand a{A, tA ? 1 : lda, tA ? lda : 1, m};
    Operand b{B, tB ? ldb : 1, tB ? 1 : ldb, n};
    auto padded = [&](int64_t bm, int64_t bn) {
        return ((m + bm - 1) / bm) * bm * (((n + bn - 1) / bn) * bn);
    };
    // 160-row tiles fit p = 160 observations exactly; otherwise the square tile
    const int64_t v44 = padded(128, 128), v52 = padded(160, 64), v25 = padded(64, 160);
    if (v52 < v44 && v52 <= v25) return launch<5, 2>(ctx, a, b, k, alpha, beta, C, ldc);
    if (v25 < v44) return launch<2, 5>(ctx, a, b, k, alpha, beta, C, ldc);
    return launch<4, 4>(ctx, a, b, k, alpha, beta, C, ldc);
}

}  // namespace hm

extern "C" int hm_dgemm(hm_ctx* ctx, int transA, int transB, int64_t m, int64_t n, int64_t k,
                        double alpha, const double* A, int64_t lda, const double* B, int64_t ldb,
                        double beta, double* C, int64_t ldc) {
    HM_REQUIRE(ctx && A && B && C, "null pointer");
    HM_REQUIRE(m >= 0 && n >= 0 && k >= 0, "negative dimension");
    HM_CUDA(cudaSetDevice(ctx->device));
    return hm::dgemm(ctx, transA != 0, transB != 0, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}
