// Device helpers shared by the simulator kernels (hm_pressure.cu, hm_sim.cu).
#pragma once

#include "hm_common.cuh"

namespace hmsim {

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
    int k_transform;  // hm_sim_desc.K_transform: 1 = the K array holds x, permeability = k_a + exp(k_b x)
    double k_a, k_b;
};

// permeability of a cell from the caller's K array (perm_transf of the notebooks, HistoryMatch.py:137-138, when asked for)
__device__ __forceinline__ double perm_value(const Geo& g, double k) { return g.k_transform ? g.k_a + exp(g.k_b * k) : k; }

// Saturation history layout (hm_sim_desc.hist_stride): row 0 = S0, then the states after steps k, 2k, ... and after
// the last step.  hist_row(step_done) = the row of the state after `step_done` steps, or -1 if it is not stored.
__host__ __device__ __forceinline__ int hist_rows(int n_steps, int stride) {
    return stride <= 1 ? n_steps + 1 : 1 + (n_steps + stride - 1) / stride;
}
__host__ __device__ __forceinline__ int hist_row(int step_done, int n_steps, int stride) {
    if (stride <= 1) return step_done;
    if (step_done % stride == 0) return step_done / stride;
    return step_done == n_steps ? hist_rows(n_steps, stride) - 1 : -1;
}

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

// ---- fractional flow for the transport kernels ---------------------------------------------------
struct Fluid {
    double inv_range;  // 1 / (1 - swc - sor)
    double swc_ir;     // swc / (1 - swc - sor)
    double mr;         // mobility ratio vw / vo:  fw = se^2 / (se^2 + mr (1-se)^2)
};
// a / b from the FP64 reciprocal seed MUFU.RCP64H (one special-function operation on the high word, no
// FP32 round trip; measured seed error 9.9e-7) and one cubically convergent correction applied to the
// QUOTIENT: q0 = a r, e = 1 - b r, q = q0 + q0 (e + e^2).  4 FP64 operations + 1 XU operation; q0 is computed
// beside e, so the dependent chain behind the seed is 3 operations deep (e, e + e^2, q) - the transport kernels
// are bound by the latency of their FP64 dependency chain, not by FP64 throughput (profiles/README.md).
// Measured on B200 over b in [0.25, 2]: max relative error of the quotient 2e-16.
__device__ __forceinline__ double fast_div_h(double a, double b) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    const double q0 = a * r;
    const double e = fma(-b, r, 1.0);
    return fma(q0, fma(e, e, e), q0);
}
__device__ __forceinline__ double frac_flow_fast(double s, const Fluid& f) {
    const double se = fma(s, f.inv_range, -f.swc_ir);
    const double t = 1.0 - se;
    const double a = se * se;
    return fast_div_h(a, fma(f.mr * t, t, a));  // denominator >= min(1, mr)/2 > 0 for every saturation
}
// UNIT: swc = sor = 0 and vw = vo (the reference's default fluid): fw = s^2 / (s^2 + (1-s)^2)
template <bool UNIT>
__device__ __forceinline__ double frac_flow_loop(double s, const Fluid& f) {
    const double se = UNIT ? s : fma(s, f.inv_range, -f.swc_ir);
    const double t = 1.0 - se;
    const double a = se * se;
    return fast_div_h(a, UNIT ? fma(t, t, a) : fma(f.mr * t, t, a));
}
// the same with the constant 1 passed in (a register the caller obtained from a volatile asm: orders the evaluation behind it)
template <bool UNIT>
__device__ __forceinline__ double frac_flow_one(double s, const Fluid& f, double one) {
    const double se = UNIT ? s : fma(s, f.inv_range, -f.swc_ir);
    const double t = one - se;
    const double a = se * se;
    return fast_div_h(a, UNIT ? fma(t, t, a) : fma(f.mr * t, t, a));
}

// Vector access helpers: N consecutive elements at a 16-byte aligned address as 128-bit transactions.
template <int N, typename U>
__device__ __forceinline__ void ldv(const U* p, U (&v)[N]) {
    if constexpr (N == 4 && sizeof(U) == 4) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
    } else if constexpr (N == 4 && sizeof(U) == 8) {
        const double2 t0 = *reinterpret_cast<const double2*>(p), t1 = *reinterpret_cast<const double2*>(p + 2);
        v[0] = t0.x, v[1] = t0.y, v[2] = t1.x, v[3] = t1.y;
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) v[k] = p[k];
    }
}
template <int N, typename U>
__device__ __forceinline__ void stv(U* p, const U (&v)[N]) {
    if constexpr (N == 4 && sizeof(U) == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else if constexpr (N == 4 && sizeof(U) == 8) {
        *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
        *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) p[k] = v[k];
    }
}

}  // namespace hmsim
