// Context management and error plumbing of libhm_b200 (C ABI in include/hm_b200.h).
#include "hm_common.cuh"

namespace hm {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

}  // namespace hm

extern "C" int hm_version(void) { return 100; }

extern "C" const char* hm_last_error(void) { return hm::g_err; }

extern "C" int hm_ctx_create(int device, hm_ctx** out) {
    HM_REQUIRE(out, "out is null");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        hm::set_error("no CUDA device visible: libhm_b200 has no CPU fallback");
        return HM_ERR_NO_DEVICE;
    }
    HM_REQUIRE(device >= 0 && device < n, "device index out of range");
    HM_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    HM_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        hm::set_error("device %d is sm_%d%d; libhm_b200 is built for sm_100a only", device, prop.major,
                      prop.minor);
        return HM_ERR_NO_DEVICE;
    }
    hm_ctx* c = new hm_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->l2_bytes = (size_t)prop.l2CacheSize;
    cudaError_t e = cudaMallocHost((void**)&c->h_pinned, 64 * sizeof(int32_t));
    if (e != cudaSuccess) {
        delete c;
        hm::set_error("cudaMallocHost failed: %s", cudaGetErrorString(e));
        return HM_ERR_CUDA;
    }
    *out = c;
    return HM_OK;
}

extern "C" int hm_ctx_destroy(hm_ctx* ctx) {
    if (!ctx) return HM_OK;
    cudaSetDevice(ctx->device);
    ctx->ws.release();
    if (ctx->solver) cusolverDnDestroy(ctx->solver);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    delete ctx;
    return HM_OK;
}

extern "C" int hm_set_stream(hm_ctx* ctx, void* cuda_stream) {
    HM_REQUIRE(ctx, "null ctx");
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    if (ctx->solver) HM_CUSOLVER(cusolverDnSetStream(ctx->solver, ctx->stream));
    return HM_OK;
}

extern "C" int hm_synchronize(hm_ctx* ctx) {
    HM_REQUIRE(ctx, "null ctx");
    HM_CUDA(cudaSetDevice(ctx->device));
    HM_CUDA(cudaStreamSynchronize(ctx->stream));
    return HM_OK;
}

extern "C" int hm_launch_count(hm_ctx* ctx, int64_t* out) {
    HM_REQUIRE(ctx && out, "null");
    *out = ctx->launches;
    return HM_OK;
}
