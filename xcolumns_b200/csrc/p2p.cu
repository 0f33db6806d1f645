// Peer-memory windows for the sharded (one process per GPU) batched sweep: every rank allocates one
// window, exports it with CUDA IPC, and maps the windows of all other ranks of the box.  The commit
// kernel (bca_batched.cu: bca_commit_p2p_kernel) then signals and reads through NVLink / NVSwitch
// directly, without a collective library call on the critical path (the
// reference is single-process, this is the B200-native replacement for its per-instance state update).
#include <cstring>
#include <new>

#include "xc_common.cuh"

extern "C" int xc_p2p_create(xc_ctx *ctx, int world, int rank, int64_t payload_bytes, xc_p2p **out,
                             void *ipc_handle_out)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !out || !ipc_handle_out || world < 1 || world > XC_P2P_MAX_WORLD || rank < 0 || rank >= world ||
        payload_bytes <= 0)
        return XC_ERR_INVALID;
    *out = nullptr;
    xc_p2p *w = new (std::nothrow) xc_p2p();
    if (!w) return XC_ERR_NOMEM;
    w->world = world;
    w->rank = rank;
    w->bytes = (size_t)XC_P2P_HEADER + (size_t)payload_bytes;
    w->epoch = 0;
    w->opened = false;
    w->windows_dev = nullptr;
    for (int r = 0; r < XC_P2P_MAX_WORLD; ++r) w->windows[r] = nullptr;
    void *base = nullptr;
    cudaError_t e = cudaMalloc(&base, w->bytes);
    if (e == cudaSuccess) e = cudaMemset(base, 0, w->bytes);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&w->windows_dev), sizeof(uint8_t *) * XC_P2P_MAX_WORLD);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, base);
    if (e != cudaSuccess) {
        ctx->last_err = e;
        if (base) cudaFree(base);
        if (w->windows_dev) cudaFree(w->windows_dev);
        delete w;
        return XC_ERR_CUDA;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    std::memcpy(ipc_handle_out, &h, sizeof(h));
    w->windows[rank] = static_cast<uint8_t *>(base);
    *out = w;
    return XC_OK;
}

// handles: world * 64 bytes, the handle of rank r at offset 64 r (as gathered by the host shim)
extern "C" int xc_p2p_open(xc_ctx *ctx, xc_p2p *w, const void *handles)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !w || !handles || w->opened) return XC_ERR_INVALID;
    for (int r = 0; r < w->world; ++r) {
        if (r == w->rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const uint8_t *>(handles) + 64 * r, sizeof(h));
        void *p = nullptr;
        XC_CUDA_TRY(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        w->windows[r] = static_cast<uint8_t *>(p);
    }
    XC_CUDA_TRY(ctx, cudaMemcpy(w->windows_dev, w->windows, sizeof(uint8_t *) * XC_P2P_MAX_WORLD, cudaMemcpyHostToDevice));
    w->opened = true;
    return XC_OK;
}

extern "C" void *xc_p2p_payload(xc_p2p *w) { return w ? w->windows[w->rank] + XC_P2P_HEADER : nullptr; }

// error word of the local window (0 = fine, otherwise the commit number that timed out); synchronises
extern "C" int xc_p2p_error(xc_ctx *ctx, xc_p2p *w, unsigned *out)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !w || !out) return XC_ERR_INVALID;
    XC_CUDA_TRY(ctx, cudaMemcpy(out, w->windows[w->rank] + 4 * XC_P2P_ERR_WORD, 4, cudaMemcpyDeviceToHost));
    if (*out) XC_CUDA_TRY(ctx, cudaMemset(w->windows[w->rank] + 4 * XC_P2P_ERR_WORD, 0, 4));   // reported once (windows are reused)
    return XC_OK;
}

extern "C" void xc_p2p_destroy(xc_ctx *ctx, xc_p2p *w)
{
    if (!w) return;
    XcDeviceGuard xc_guard__(ctx);
    for (int r = 0; r < w->world; ++r) {
        if (!w->windows[r]) continue;
        if (r == w->rank) cudaFree(w->windows[r]);
        else cudaIpcCloseMemHandle(w->windows[r]);
    }
    if (w->windows_dev) cudaFree(w->windows_dev);
    delete w;
}
