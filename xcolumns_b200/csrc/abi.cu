// Context management and error reporting of the C ABI (include/xcolumns_b200.h).
#include <cstdlib>
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
    ctx->aux_ready = false;
    ctx->pipe_active = false;
    ctx->pipe_forked = false;
    ctx->timing_on = false;
    ctx->timing_commits = false;
    ctx->timing_count = ctx->timing_cap = 0;
    ctx->timing_ev = nullptr;
    ctx->timing_rows = nullptr;
    ctx->stage[0] = ctx->stage[1] = nullptr;
    XcDeviceGuard guard(ctx);   // allocate on `device`, leave the caller's current device untouched
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
    XcDeviceGuard guard(ctx);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->red_partials) cudaFree(ctx->red_partials);
    if (ctx->red_counter) cudaFree(ctx->red_counter);
    if (ctx->aux_ready) {
        for (int i = 0; i <= XC_PIPE_MAX_LAG; ++i) {
            cudaStreamDestroy(ctx->aux[i]);
            cudaEventDestroy(ctx->ev_k[i]);
            cudaEventDestroy(ctx->ev_c[i]);
            cudaStreamDestroy(ctx->pstream[i]);
            cudaEventDestroy(ctx->ev_p[i]);
        }
        for (int i = 0; i <= XC_PIPE_MAX_LAG + 1; ++i) cudaEventDestroy(ctx->ev_join[i]);
        cudaStreamDestroy(ctx->cstream);
        cudaEventDestroy(ctx->ev_fork);
        cudaEventDestroy(ctx->ev_pro);
        cudaEventDestroy(ctx->ev_util);
    }
    for (int i = 0; i < 2 * ctx->timing_cap; ++i) cudaEventDestroy(ctx->timing_ev[i]);
    free(ctx->timing_ev);
    for (int i = 0; i < 2; ++i)
        if (ctx->stage[i]) {
            cudaFreeHost(ctx->stage[i]);
            cudaEventDestroy(ctx->stage_ev[i]);
        }
    free(ctx->timing_rows);
    delete ctx;
}

// ---- per-launch timing of the streaming batch kernels (bench.py's roofline leg) --------------------------
// While enabled, the batched-sweep entry points bracket every batch kernel with CUDA events on the stream it
// is launched on.  xc_timing_read synchronises the device and returns, per launch, start / end in ms relative
// to the first recorded event plus the rows it processed.
extern "C" int xc_timing_enable(xc_ctx *ctx, int on)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx) return XC_ERR_INVALID;
    ctx->timing_on = on != 0;
    ctx->timing_commits = on == 2;
    ctx->timing_count = 0;
    return XC_OK;
}

int xc_timing_slot(xc_ctx *ctx, int64_t rows, cudaEvent_t *start, cudaEvent_t *end)
{
    if (ctx->timing_count == ctx->timing_cap) {
        const int cap = ctx->timing_cap ? 2 * ctx->timing_cap : 256;
        cudaEvent_t *ev = static_cast<cudaEvent_t *>(realloc(ctx->timing_ev, sizeof(cudaEvent_t) * 2 * cap));
        if (!ev) return XC_ERR_NOMEM;
        ctx->timing_ev = ev;
        int64_t *rw = static_cast<int64_t *>(realloc(ctx->timing_rows, sizeof(int64_t) * cap));
        if (!rw) return XC_ERR_NOMEM;
        ctx->timing_rows = rw;
        for (int i = 2 * ctx->timing_cap; i < 2 * cap; ++i) XC_CUDA_TRY(ctx, cudaEventCreate(&ctx->timing_ev[i]));
        ctx->timing_cap = cap;
    }
    const int i = ctx->timing_count++;
    ctx->timing_rows[i] = rows;
    *start = ctx->timing_ev[2 * i];
    *end = ctx->timing_ev[2 * i + 1];
    return XC_OK;
}

extern "C" int xc_timing_read(xc_ctx *ctx, int cap, double *start_ms_host, double *end_ms_host, int64_t *rows_host,
                              int *count_host)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !count_host || cap < 0) return XC_ERR_INVALID;
    XC_CUDA_TRY(ctx, cudaDeviceSynchronize());
    const int n = ctx->timing_count < cap ? ctx->timing_count : cap;
    for (int i = 0; i < n; ++i) {
        float a = 0.f, b = 0.f;
        XC_CUDA_TRY(ctx, cudaEventElapsedTime(&a, ctx->timing_ev[0], ctx->timing_ev[2 * i]));
        XC_CUDA_TRY(ctx, cudaEventElapsedTime(&b, ctx->timing_ev[0], ctx->timing_ev[2 * i + 1]));
        if (start_ms_host) start_ms_host[i] = a;
        if (end_ms_host) end_ms_host[i] = b;
        if (rows_host) rows_host[i] = ctx->timing_rows[i];
    }
    *count_host = ctx->timing_count;
    ctx->timing_count = 0;
    return XC_OK;
}

int xc_ctx_aux_streams(xc_ctx *ctx)
{
    if (ctx->aux_ready) return XC_OK;
    int lo = 0, hi = 0;
    XC_CUDA_TRY(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    // (cudaDeviceGetStreamPriorityRange: `lo` is the numerically greatest = lowest priority, `hi` the highest)
    for (int i = 0; i <= XC_PIPE_MAX_LAG; ++i) {
        XC_CUDA_TRY(ctx, cudaStreamCreateWithPriority(&ctx->aux[i], cudaStreamNonBlocking, lo));
        XC_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_k[i], cudaEventDisableTiming));
        XC_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_c[i], cudaEventDisableTiming));
        XC_CUDA_TRY(ctx, cudaStreamCreateWithPriority(&ctx->pstream[i], cudaStreamNonBlocking, hi));
        XC_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_p[i], cudaEventDisableTiming));
    }
    for (int i = 0; i <= XC_PIPE_MAX_LAG + 1; ++i)
        XC_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming));
    XC_CUDA_TRY(ctx, cudaStreamCreateWithPriority(&ctx->cstream, cudaStreamNonBlocking, hi));
    XC_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    XC_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_pro, cudaEventDisableTiming));
    XC_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_util, cudaEventDisableTiming));
    ctx->aux_ready = true;
    return XC_OK;
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

// ---- pseudo-random permutation of 0..n-1 (visiting order of a batched sweep) ----------------------------
// A 4-round balanced Feistel network on the smallest even-width power-of-two domain >= n, with cycle
// walking (re-encrypt until the value falls below n): a bijection of [0, n) computed independently per
// element, one 6 us kernel instead of a key-sort (torch.randperm: ~95 us at n = 307 k, measured).
namespace {
__device__ __forceinline__ uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

__global__ void __launch_bounds__(256) permutation_kernel(int64_t n, uint64_t seed, int half_bits, int32_t *out)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const uint32_t mask = (1u << half_bits) - 1u;
    uint32_t k0 = mix32((uint32_t)seed), k1 = mix32((uint32_t)(seed >> 32) ^ 0x9e3779b9U);
    uint64_t x = (uint64_t)i;
    do {
        uint32_t l = (uint32_t)(x >> half_bits) & mask, r = (uint32_t)x & mask;
#pragma unroll
        for (int round = 0; round < 4; ++round) {
            uint32_t f = mix32(r ^ (round & 1 ? k1 : k0) ^ (0x85ebca6bU * (uint32_t)(round + 1))) & mask;
            uint32_t t = l ^ f;
            l = r;
            r = t;
        }
        x = ((uint64_t)l << half_bits) | r;
    } while (x >= (uint64_t)n);
    out[i] = (int32_t)x;
}
}  // namespace

extern "C" int xc_permutation(xc_ctx *ctx, int64_t n, uint64_t seed, int32_t *out, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !out || n < 0 || n > 0x7fffffffLL) return XC_ERR_INVALID;
    if (n == 0) return XC_OK;
    int bits = 2;
    while ((1ULL << bits) < (uint64_t)n) bits += 2;   // even width, domain < 4n -> < 4 walks expected
    permutation_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, seed, bits / 2, out);
    XC_LAUNCHED(ctx);
    return XC_OK;
}
