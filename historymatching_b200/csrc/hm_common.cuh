// Shared internals of libhm_b200: context, workspace arena, error plumbing.
#pragma once

#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/hm_b200.h"

namespace hm {

void set_error(const char* fmt, ...);

#define HM_CUDA(call)                                                                      \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            hm::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return HM_ERR_CUDA;                                                            \
        }                                                                                  \
    } while (0)

#define HM_CUSOLVER(call)                                                                  \
    do {                                                                                   \
        cusolverStatus_t s_ = (call);                                                      \
        if (s_ != CUSOLVER_STATUS_SUCCESS) {                                               \
            hm::set_error("%s:%d: %s -> cusolver status %d", __FILE__, __LINE__, #call, (int)s_); \
            return HM_ERR_CUDA;                                                            \
        }                                                                                  \
    } while (0)

#define HM_CHECK(call)              \
    do {                            \
        int rc_ = (call);           \
        if (rc_ != HM_OK) return rc_; \
    } while (0)

#define HM_REQUIRE(cond, msg)                                            \
    do {                                                                 \
        if (!(cond)) {                                                   \
            hm::set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, msg); \
            return HM_ERR_ARG;                                           \
        }                                                                \
    } while (0)

// Named, grow-only device buffers: the ctx workspace.  A buffer keeps its
// address until a larger size is requested, so steady-state calls allocate nothing.
struct Arena {
    struct Buf {
        void* ptr = nullptr;
        size_t bytes = 0;
    };
    std::map<std::string, Buf> bufs;

    int get(const char* name, size_t bytes, void** out) {
        Buf& b = bufs[name];
        if (b.bytes < bytes) {
            if (b.ptr) cudaFree(b.ptr);
            b.ptr = nullptr;
            b.bytes = 0;
            cudaError_t e = cudaMalloc(&b.ptr, bytes);
            if (e != cudaSuccess) {
                set_error("cudaMalloc(%zu bytes) for workspace '%s' failed: %s", bytes, name,
                          cudaGetErrorString(e));
                return HM_ERR_CUDA;
            }
            b.bytes = bytes;
        }
        *out = b.ptr;
        return HM_OK;
    }
    template <typename T>
    int get(const char* name, size_t count, T** out) {
        void* p = nullptr;
        int rc = get(name, count * sizeof(T), &p);
        *out = static_cast<T*>(p);
        return rc;
    }
    void release() {
        for (auto& kv : bufs)
            if (kv.second.ptr) cudaFree(kv.second.ptr);
        bufs.clear();
    }
};

}  // namespace hm

struct hm_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    size_t l2_bytes = 0;
    hm::Arena ws;
    cusolverDnHandle_t solver = nullptr;
    int32_t* h_pinned = nullptr;  // small pinned scratch for device->host scalars
    hm_sim_stats sim_stats{};
    int64_t launches = 0;  // kernels of this library launched on the ctx
    double phase_ms[5] = {0, 0, 0, 0, 0};
    int tb_active[4][2][17] = {};  // k_sat_tb<W, UNIT>: resident clusters per cluster size (0 = not queried, -1 = none)
    bool mg_force64 = false;  // pressure_solve: the FP32 multigrid cycle converged too slowly in this forward run
};
