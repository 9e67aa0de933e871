// FP64 tensor-core GEMM for the analysis step (sm_100a).
//
// Replaces the numpy "@" products of the ensemble update
// (HistoryMatch.py:583-586, 920, 928-941).  tcgen05 has no FP64 kind, so the
// FP64 tensor path on Blackwell is the warp-level DMMA
// (mma.sync.aligned.m8n8k4.f64, SASS DMMA.8x8x4).
//
// C[m,n] = alpha * op(A)[m,k] * op(B)[k,n] + beta * C, everything row-major.
// CTA tile (32*WM) x (32*WN) x 16, one warp per 32x32 sub-tile (4x4 DMMA
// tiles, 64 accumulator registers).  Operands are staged in shared memory by
// cp.async (LDGSTS, zero-filled outside the matrices) through a 3-stage ring:
// a "full" mbarrier per stage completes when every thread's copies of that
// k-block have landed (cp.async.mbarrier.arrive), an "empty" mbarrier when all
// warps have consumed it.  There is no CTA-wide barrier in the main loop, so a
// warp can run up to one k-block ahead of the slowest one and the FP64 pipe is
// not drained at every block (the register-prefetch version lost 10 % of its
// stall samples at its __syncthreads and ran at 0.75 of cuBLAS; this one
// runs at 0.92 - 0.93).  Copies are 16 bytes wide where the operand's alignment
// allows, 8 bytes otherwise.
// Layout per stage: As[m][k], Bs[n][k] with a row pitch of 20 doubles, or
// k-major [k][row] with a pitch of rows + 4 for an operand whose rows are
// contiguous in memory - conflict-free 64-bit fragment loads in both.
#include "hm_common.cuh"
#include "hm_ptx.cuh"

namespace {

using hmsim::mbar_init;
using hmsim::mbar_wait;
using hmsim::smem_u32;

// Measured on B200, W (1024 x 1024) @ X0 (1024 x 16384), TFLOP/s (cuBLAS 34.9): BK 16 x 3 stages 31.9; 4 stages 30.4;
// BK 8 x 4 stages 29.9; BK 32 x 2 stages 26.7; 8-byte copies only 30.9; copies requested before instead of after the
// block's multiplication 31.5 (profiles/dgemm_r2_shapes.txt).
constexpr int BK = 16;
constexpr int LDS = BK + 4;
constexpr int STAGES = 3;

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
// 8-byte asynchronous copy; src_bytes = 0 writes zeros and does not touch the source
__device__ __forceinline__ void cp_async_8z(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// 16-byte copy past L1; src_bytes in {0, 8, 16}, the rest of the 16 bytes is zero-filled
__device__ __forceinline__ void cp_async_16z(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// arrive on `bar` once all cp.async issued so far by this thread have completed (the barrier's count includes it)
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Element (row i, reduction index kk) of a "rows x k" operand lives at base[i*s_row + kk*s_k].
struct Operand {
    const double* base;
    int64_t s_row, s_k;
    int64_t rows;
    int vec;  // 16-byte copies are possible: base 16-byte aligned, the non-unit stride even
};

// Issue the copies of one ROWS x BK operand block into a stage.  KM: the operand's rows are contiguous in memory
// (s_row == 1), consecutive threads take consecutive rows and the stage is k-major; otherwise s_k == 1 and
// consecutive threads run along k.
template <int ROWS, int NT, bool KM>
__device__ __forceinline__ void issue_tile(const Operand& op, int64_t row0, int64_t k0, int64_t K, double* sm) {
    if (op.vec) {  // 16-byte copies of two elements adjacent along the unit-stride direction (alignment checked by the host)
        constexpr int PER = (ROWS * BK / 2 + NT - 1) / NT;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int e = threadIdx.x + j * NT;
            if (e < ROWS * BK / 2) {
                const int r = KM ? 2 * (e % (ROWS / 2)) : e / (BK / 2), kk = KM ? e / (ROWS / 2) : 2 * (e % (BK / 2));
                const int64_t gr = row0 + r, gk = k0 + kk;
                const bool ok = gr < op.rows && gk < K;
                const bool ok2 = KM ? gr + 1 < op.rows : gk + 1 < K;
                const double* src = ok ? op.base + gr * op.s_row + gk * op.s_k : op.base;
                cp_async_16z(smem_u32(sm + (KM ? kk * (ROWS + 4) + r : r * LDS + kk)), src, ok ? (ok2 ? 16 : 8) : 0);
            }
        }
        return;
    }
    constexpr int PER = (ROWS * BK + NT - 1) / NT;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const int e = threadIdx.x + j * NT;
        if (e < ROWS * BK) {
            const int r = KM ? e % ROWS : e / BK, kk = KM ? e / ROWS : e % BK;
            const int64_t gr = row0 + r, gk = k0 + kk;
            const bool ok = gr < op.rows && gk < K;
            const double* src = ok ? op.base + gr * op.s_row + gk * op.s_k : op.base;
            cp_async_8z(smem_u32(sm + (KM ? kk * (ROWS + 4) + r : r * LDS + kk)), src, ok ? 8 : 0);
        }
    }
}

// AKM / BKM: the operand is k-major in shared memory (its rows are contiguous in global memory, see issue_tile)
template <int WM, int WN, bool AKM, bool BKM>
__global__ void __launch_bounds__(WM * WN * 32)
k_dgemm(Operand A, Operand B, int64_t K, double alpha, double beta, double* __restrict__ C,
        int64_t ldc) {
    constexpr int BM = 32 * WM, BN = 32 * WN, NT = WM * WN * 32, NW = WM * WN;
    constexpr int STAGE = (BM + BN) * LDS;  // doubles per stage: A block, then B block
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t bars[2 * STAGES];  // full[s], empty[s]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp / WN, wn = warp % WN;
    const int64_t row0 = (int64_t)blockIdx.y * BM, col0 = (int64_t)blockIdx.x * BN;
    const int g = lane >> 2, q = lane & 3;
    const uint32_t bar0 = smem_u32(bars);
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (STAGES + s); };

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full(s), NT);
            mbar_init(empty(s), NW);
        }
    }
    __syncthreads();

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int64_t nk = (K + BK - 1) / BK;
    auto issue = [&](int64_t kb, int s) {  // k-block kb -> stage s
        double* as = smem + s * STAGE;
        issue_tile<BM, NT, AKM>(A, row0, kb * BK, K, as);
        issue_tile<BN, NT, BKM>(B, col0, kb * BK, K, as + BM * LDS);
        cp_async_arrive(full(s));
    };
    // Block kt + PF is requested after block kt has been multiplied; it goes into the stage block kt - 1 used.
    constexpr int PF = STAGES - 1;
    const int nkb = (int)nk;
    auto refill = [&](int kt) {
        const int nb = kt + PF, lb = nb - STAGES;
        if (nb < nkb) {
            if (lb >= 0) mbar_wait(empty(lb % STAGES), (lb / STAGES) & 1);  // every warp is done with the stage's last block
            issue(nb, nb % STAGES);
        }
    };
#pragma unroll
    for (int b = 0; b < PF; ++b)
        if (b < nkb) issue(b, b);

    int s = 0, ph = 0;  // stage / parity of k-block kt
    for (int kt = 0; kt < nkb; ++kt) {
        mbar_wait(full(s), ph);
        // fragment element (row 8 i + g, k = ks + q): [row][k] layout (pitch LDS) or [k][row] layout (pitch ROWS + 4); both
        // are conflict-free 64-bit loads (per half-warp: 4 values of g x 4 of q on 16 distinct banks)
        const double* as = smem + s * STAGE + (AKM ? q * (BM + 4) + wm * 32 + g : (wm * 32 + g) * LDS + q);
        const double* bs = smem + s * STAGE + BM * LDS + (BKM ? q * (BN + 4) + wn * 32 + g : (wn * 32 + g) * LDS + q);
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
        __syncwarp();
        if (lane == 0) mbar_arrive(empty(s));  // this warp is done with block kt
        refill(kt);
        if (++s == STAGES) {
            s = 0;
            ph ^= 1;
        }
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
    const size_t smem = (size_t)STAGES * (BM + BN) * LDS * sizeof(double);
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
    auto vec_ok = [](const double* p, int64_t ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ld % 2 == 0; };
    Operand a{A, tA ? 1 : lda, tA ? lda : 1, m, vec_ok(A, lda)};
    Operand b{B, tB ? ldb : 1, tB ? 1 : ldb, n, vec_ok(B, ldb)};
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
