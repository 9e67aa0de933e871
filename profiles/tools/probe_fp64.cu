// Micro-benchmark: FP64 pipe throughput of one B200 SM (DFMA / DADD / DMUL warp instructions per clock per SM
// sub-partition) and the cost of MUFU.RCP64H and of predicated-off FP64 instructions.  The transport kernel
// k_sat_cluster issues 26 FP64 + 2 MUFU warp instructions per sub-step and warp (SASS, hm_sim.cu); this probe
// gives the roofline those are measured against.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_fp64 probe_fp64.cu && ./probe_fp64
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_probe(double* out, int iters, double x, int never) {
    __shared__ double sh[1024];
    sh[threadIdx.x] = x;
    __syncthreads();
    int ia[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) ia[j] = threadIdx.x + j;
    double a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = x + threadIdx.x * 1e-9 + j;
    const double b = 1.0000001, c = 1e-9;
    const bool p = never != 0;  // always false at run time, unknown to the compiler
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) a[j] = fma(a[j], b, c);          // DFMA
            if (MODE == 1) a[j] = a[j] + c;                 // DADD
            if (MODE == 2) a[j] = a[j] * b;                 // DMUL
            if (MODE == 3) {                                // DFMA + 1 MUFU.RCP64H per 13 DFMA-equivalents
                a[j] = fma(a[j], b, c);
                if (j == 0) {
                    double r;
                    asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a[0]));
                    a[1] = fma(r, c, a[1]);
                }
            }
            if (MODE == 5) {                                // DFMA + an independent integer multiply-add (issue-port test)
                a[j] = fma(a[j], b, c);
                ia[j] = ia[j] * 3 + never;
            }
            if (MODE == 6) {                                // DFMA + an independent shared-memory load
                a[j] = fma(a[j], b, c);
                ia[j] += (int)__double_as_longlong(*(volatile double*)&sh[(threadIdx.x + j * 32 + i) & 1023]);
            }
            if (MODE == 7) {                                // integer multiply-add only
                ia[j] = ia[j] * 3 + never;
            }
            if (MODE == 4) {                                // DFMA + a predicated-off DADD
                a[j] = fma(a[j], b, c);
                asm volatile("{ .reg .pred q; setp.ne.s32 q, %1, 0; @q add.f64 %0, %0, 1.0; }" : "+d"(a[j]) : "r"((int)p));
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j] + ia[j];
    if (s == 123.456) out[0] = s;
}

template <int MODE>
void run(const char* name, int inst_per_iter, int sms, double mhz) {
    double* out;
    cudaMalloc(&out, 8);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_probe<MODE><<<sms, 1024>>>(out, 100, 1.0, 0);
    cudaEventRecord(e0);
    k_probe<MODE><<<sms, 1024>>>(out, iters, 1.0, 0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cycles = ms * 1e-3 * mhz * 1e6;
    const double warp_inst_per_smsp = (double)iters * inst_per_iter * 8 /* warps per SMSP */;
    printf("%-34s %8.3f ms  %6.3f cycles per FP64 warp instruction per SM sub-partition  (%.1f FP64 lanes/clk/SM)\n", name, ms,
           cycles / warp_inst_per_smsp, 32.0 * 4 * warp_inst_per_smsp / cycles);
    cudaFree(out);
}

// shared-memory crossbar against warp shuffles: 32 warps per SM, 8 independent operations per thread and iteration
//   MODE 0: 8 LDS.64   MODE 1: 16 SHFL.32 (= 8 double shuffles)   MODE 2: 8 LDS.64 + 16 SHFL.32   MODE 3: 8 LDS.64 + 8 DFMA
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_xbar(double* out, int iters, int never) {
    __shared__ double sh[2048];
    sh[threadIdx.x] = threadIdx.x;
    sh[threadIdx.x + 1024] = 1.0;
    __syncthreads();
    double a[8], v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = j, v[j] = threadIdx.x + j;
    int idx = threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0 || MODE == 2 || MODE == 3) a[j] += *(volatile double*)&sh[(idx + 32 * j) & 2047];
            if (MODE == 1 || MODE == 2) v[j] = __shfl_down_sync(0xffffffffu, v[j], 1);
            if (MODE == 3) v[j] = fma(v[j], 1.0000001, 1e-9);
        }
        idx += never;
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j] + v[j];
    if (s == 123.456) out[0] = s;
}
template <int MODE>
void run_xbar(const char* name, int sms, double mhz) {
    double* out;
    cudaMalloc(&out, 8);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_xbar<MODE><<<sms, 1024>>>(out, 100, 0);
    cudaEventRecord(e0);
    k_xbar<MODE><<<sms, 1024>>>(out, iters, 0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %8.3f ms  %6.2f cycles per SM for one warp's 8 operations\n", name, ms,
           ms * 1e-3 * mhz * 1e6 / (iters * 32.0));
    cudaFree(out);
}

// dependent-issue latency: ONE warp per SM sub-partition, one dependent chain per thread
template <int MODE>
__global__ void __launch_bounds__(128, 1) k_latency(double* out, int iters, double x) {
    double a = x + threadIdx.x * 1e-9;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) a = fma(a, b, c);
            if (MODE == 1) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(a));
        }
    }
    if (a == 123.456) out[0] = a;
}
template <int MODE>
void run_latency(const char* name, int sms, double mhz) {
    double* out;
    cudaMalloc(&out, 8);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_latency<MODE><<<sms, 128>>>(out, 100, 1.0);
    cudaEventRecord(e0);
    k_latency<MODE><<<sms, 128>>>(out, iters, 1.0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-34s %8.3f ms  %6.2f cycles of dependent-issue latency\n", name, ms, ms * 1e-3 * mhz * 1e6 / (iters * 8.0));
    cudaFree(out);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1e3;
    printf("%s, %d SMs, %.0f MHz (max SM clock, used to convert time into cycles)\n", p.name, p.multiProcessorCount, mhz);
    run<0>("DFMA x8 chains", 8, p.multiProcessorCount, mhz);
    run<1>("DADD x8 chains", 8, p.multiProcessorCount, mhz);
    run<2>("DMUL x8 chains", 8, p.multiProcessorCount, mhz);
    run<3>("DFMA + MUFU.RCP64H (1 per 8)", 9, p.multiProcessorCount, mhz);
    run<4>("DFMA + predicated-off DADD", 8, p.multiProcessorCount, mhz);
    run<5>("DFMA + independent IMAD (1:1)", 8, p.multiProcessorCount, mhz);
    run<6>("DFMA + independent LDS.64 (1:1)", 8, p.multiProcessorCount, mhz);
    run<7>("IMAD only (per-IMAD cycles)", 8, p.multiProcessorCount, mhz);
    run_xbar<0>("8 x LDS.64 per thread", p.multiProcessorCount, mhz);
    run_xbar<1>("8 x 64-bit shuffle (16 SHFL.32) per thread", p.multiProcessorCount, mhz);
    run_xbar<2>("8 x LDS.64 + 8 x 64-bit shuffle per thread", p.multiProcessorCount, mhz);
    run_xbar<3>("8 x LDS.64 + 8 x DFMA per thread", p.multiProcessorCount, mhz);
    run_latency<0>("DFMA dependent chain (1 warp/SMSP)", p.multiProcessorCount, mhz);
    run_latency<1>("MUFU.RCP64H dependent chain", p.multiProcessorCount, mhz);
    return 0;
}
