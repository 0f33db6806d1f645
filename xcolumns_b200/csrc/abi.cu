// Context management and error reporting of the C ABI (include/xcolumns_b200.h).
#include <new>

#include "xc_common.cuh"

extern "C" int xc_abi_version(void) { return XC_ABI_VERSION; }

extern "C" const char *xc_strerror(int code)
{
    switch (code) {
    case XC_OK: return "ok";
    case XC_ERR_INVALID: return "invalid argument";
    case XC_ERR_UNSUPPORTED: return "unsupported dtype/metric/shape combination";
    case XC_ERR_CUDA: return "CUDA runtime error (see xc_last_cuda_error)";
    case XC_ERR_NOMEM: return "out of memory";
    default: return "unknown error";
    }
}

extern "C" int xc_ctx_create(int device, xc_ctx **out)
{
    if (!out) return XC_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device < 0 || device >= count) return XC_ERR_CUDA;
    xc_ctx *ctx = new (std::nothrow) xc_ctx();
    if (!ctx) return XC_ERR_NOMEM;
    ctx->device = device;
    ctx->launches = 0;
    ctx->last_err = cudaSuccess;
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
    for (int i = 0; i < 8; ++i) ctx->coop_blocks_cache[i] = 0;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        delete ctx;
        return XC_ERR_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->red_partials = nullptr;
    ctx->red_counter = nullptr;
    cudaSetDevice(device);
    if (cudaMalloc(&ctx->red_partials, sizeof(double) * XC_RED_MAX_BLOCKS) != cudaSuccess ||
        cudaMalloc(&ctx->red_counter, 64) != cudaSuccess || cudaMemset(ctx->red_counter, 0, 64) != cudaSuccess) {
        delete ctx;
        return XC_ERR_CUDA;
    }
    *out = ctx;
    return XC_OK;
}

extern "C" void xc_ctx_destroy(xc_ctx *ctx)
{
    if (!ctx) return;
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->red_partials) cudaFree(ctx->red_partials);
    if (ctx->red_counter) cudaFree(ctx->red_counter);
    delete ctx;
}

extern "C" const char *xc_last_cuda_error(xc_ctx *ctx)
{
    return ctx ? cudaGetErrorString(ctx->last_err) : "no context";
}

extern "C" int64_t xc_launch_count(xc_ctx *ctx) { return ctx ? ctx->launches : -1; }
extern "C" int xc_sm_count(xc_ctx *ctx) { return ctx ? ctx->sm_count : -1; }

int xc_ctx_scratch(xc_ctx *ctx, size_t bytes, void **out)
{
    if (ctx->scratch_bytes < bytes) {
        if (ctx->scratch) cudaFree(ctx->scratch);
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
        size_t want = bytes < (1u << 20) ? (1u << 20) : bytes;
        XC_CUDA_TRY(ctx, cudaMalloc(&ctx->scratch, want));
        ctx->scratch_bytes = want;
    }
    *out = ctx->scratch;
    return XC_OK;
}
