// Host-side helpers of the boundary (no device code): materialising the dense 0/1 prediction
// matrix the reference API returns for dense inputs.  For AmazonCat-13K-shape outputs this is a
// 16 GB host write (first-touch page faults + zeroing); it is spread over host threads.
#include <sys/mman.h>

#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

#include "xc_common.cuh"

namespace {

template <typename TO, typename TV>
void fill_rows(TO *out, int64_t ld, int64_t m, const int32_t *idx, const TV *val, int k, int64_t r0, int64_t r1,
               bool zero)
{
    for (int64_t i = r0; i < r1; ++i) {
        TO *row = out + i * ld;
        if (zero) std::memset(row, 0, sizeof(TO) * (size_t)m);
        for (int t = 0; t < k; ++t) {
            int j = idx[i * k + t];
            if (j >= 0 && j < m) row[j] = val ? (TO)val[i * k + t] : (TO)1;
        }
    }
}

template <typename TO, typename TV>
int fill_threads(void *out, int64_t n, int64_t m, int64_t ld, const int32_t *idx, const void *val, int k, int nthreads,
                 bool zero)
{
    if (nthreads < 1) nthreads = 1;
    int64_t chunk = (n + nthreads - 1) / nthreads;
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
        int64_t r0 = t * chunk, r1 = std::min(n, r0 + chunk);
        if (r0 >= r1) break;
        th.emplace_back(fill_rows<TO, TV>, (TO *)out, ld, m, idx, (const TV *)val, k, r0, r1, zero);
    }
    for (auto &x : th) x.join();
    return XC_OK;
}

int default_threads()
{
    unsigned hc = std::thread::hardware_concurrency();
    return (int)std::min<unsigned>(hc ? hc : 1, 32);
}

int fill_dispatch(void *out_host, int out_dtype, int64_t n, int64_t m, int64_t ld, const int32_t *idx_host,
                  const void *val_host, int val_dtype, int k, int nthreads, bool zero)
{
    if (!out_host || !idx_host || n < 0 || m <= 0 || ld < m || k < 1) return XC_ERR_INVALID;
    if (nthreads <= 0) nthreads = default_threads();
    if ((zero ? n * m : n * (int64_t)k) < (1 << 22)) nthreads = 1;
    if (out_dtype == XC_F32 && (!val_host || val_dtype == XC_F32)) return fill_threads<float, float>(out_host, n, m, ld, idx_host, val_host, k, nthreads, zero);
    if (out_dtype == XC_F32 && val_dtype == XC_F64) return fill_threads<float, double>(out_host, n, m, ld, idx_host, val_host, k, nthreads, zero);
    if (out_dtype == XC_F64 && (!val_host || val_dtype == XC_F64)) return fill_threads<double, double>(out_host, n, m, ld, idx_host, val_host, k, nthreads, zero);
    if (out_dtype == XC_F64 && val_dtype == XC_F32) return fill_threads<double, float>(out_host, n, m, ld, idx_host, val_host, k, nthreads, zero);
    return XC_ERR_UNSUPPORTED;
}

}  // namespace

extern "C" int xc_fill_pred_dense_host(void *out_host, int out_dtype, int64_t n, int64_t m, int64_t ld,
                                       const int32_t *idx_host, const void *val_host, int val_dtype, int k,
                                       int nthreads)
{
    return fill_dispatch(out_host, out_dtype, n, m, ld, idx_host, val_host, val_dtype, k, nthreads, true);
}

// the scatter alone, into a matrix that xc_zero_host already cleared (the zero-fill of a 16 GB result
// runs on host threads WHILE the scores are uploaded and swept; only this tiny step is left at the end)
extern "C" int xc_scatter_pred_dense_host(void *out_host, int out_dtype, int64_t n, int64_t m, int64_t ld,
                                          const int32_t *idx_host, const void *val_host, int val_dtype, int k,
                                          int nthreads)
{
    return fill_dispatch(out_host, out_dtype, n, m, ld, idx_host, val_host, val_dtype, k, nthreads, false);
}

extern "C" int xc_zero_host(void *p, int64_t bytes, int nthreads)
{
    if (!p || bytes < 0) return XC_ERR_INVALID;
    if (nthreads <= 0) nthreads = default_threads();
    if (bytes < (1 << 24)) nthreads = 1;
#ifdef MADV_HUGEPAGE
    // fresh anonymous memory: ask for transparent huge pages so that first touch costs one fault per
    // 2 MB instead of one per 4 KB (the page faults, not the stores, dominate a 16 GB clear)
    {
        const uintptr_t lo = (reinterpret_cast<uintptr_t>(p) + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1);
        const uintptr_t hi = (reinterpret_cast<uintptr_t>(p) + (uintptr_t)bytes) & ~(uintptr_t)((2u << 20) - 1);
        if (hi > lo) madvise(reinterpret_cast<void *>(lo), hi - lo, MADV_HUGEPAGE);
    }
#endif
    const int64_t chunk = ((bytes + nthreads - 1) / nthreads + 4095) & ~(int64_t)4095;
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
        const int64_t b0 = t * chunk, b1 = std::min(bytes, b0 + chunk);
        if (b0 >= b1) break;
        th.emplace_back([=] { std::memset(static_cast<uint8_t *>(p) + b0, 0, (size_t)(b1 - b0)); });
    }
    for (auto &x : th) x.join();
    return XC_OK;
}

// ---- pageable host memory -> device ---------------------------------------------------------------------------
// What a caller of the reference passes is an ordinary numpy array: pageable memory, which the driver can only
// DMA through its own small staging buffers (~10 GB/s).  Here host threads copy chunk c + 1 into one of two pinned
// staging buffers while the DMA of chunk c runs, so the upload proceeds at min(threaded memcpy, PCIe) instead.
// Rows of width_bytes are copied from src (pitch src_pitch) to dst (pitch dst_pitch).  The call returns when the
// last chunk has been queued on `stream`; src may be reused once the stream has passed that copy.
namespace {
constexpr size_t kStageBytes = (size_t)128 << 20;

void copy_rows_threads(uint8_t *dst, const uint8_t *src, int64_t src_pitch, int64_t width, int64_t rows, int nthreads)
{
    const int64_t per = (rows + nthreads - 1) / nthreads;
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
        const int64_t r0 = t * per, r1 = std::min(rows, r0 + per);
        if (r0 >= r1) break;
        th.emplace_back([=] {
            if (src_pitch == width) std::memcpy(dst + r0 * width, src + r0 * src_pitch, (size_t)((r1 - r0) * width));
            else for (int64_t r = r0; r < r1; ++r) std::memcpy(dst + r * width, src + r * src_pitch, (size_t)width);
        });
    }
    for (auto &x : th) x.join();
}
}  // namespace

extern "C" int xc_h2d_staged(xc_ctx *ctx, void *dst_dev, int64_t dst_pitch, const void *src_host, int64_t src_pitch,
                             int64_t width_bytes, int64_t rows, int nthreads, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !dst_dev || !src_host || width_bytes <= 0 || rows < 0 || dst_pitch < width_bytes || src_pitch < width_bytes)
        return XC_ERR_INVALID;
    if ((size_t)width_bytes > kStageBytes) return XC_ERR_UNSUPPORTED;
    if (rows == 0) return XC_OK;
    if (nthreads <= 0) nthreads = std::min(default_threads(), 16);
    if (!ctx->stage[0]) {
        for (int i = 0; i < 2; ++i) {
            XC_CUDA_TRY(ctx, cudaHostAlloc(&ctx->stage[i], kStageBytes, cudaHostAllocDefault));
            XC_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->stage_ev[i], cudaEventDisableTiming));
        }
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t rows_per_chunk = std::max<int64_t>(1, (int64_t)(kStageBytes / (size_t)width_bytes));
    int c = 0;
    for (int64_t r0 = 0; r0 < rows; r0 += rows_per_chunk, ++c) {
        const int64_t nr = std::min(rows_per_chunk, rows - r0);
        const int b = c & 1;
        if (c >= 2) XC_CUDA_TRY(ctx, cudaEventSynchronize(ctx->stage_ev[b]));   // the DMA out of this buffer is done
        copy_rows_threads(static_cast<uint8_t *>(ctx->stage[b]), static_cast<const uint8_t *>(src_host) + r0 * src_pitch,
                          src_pitch, width_bytes, nr, nr * width_bytes < (1 << 22) ? 1 : nthreads);
        XC_CUDA_TRY(ctx, cudaMemcpy2DAsync(static_cast<uint8_t *>(dst_dev) + r0 * dst_pitch, (size_t)dst_pitch, ctx->stage[b],
                                           (size_t)width_bytes, (size_t)width_bytes, (size_t)nr, cudaMemcpyHostToDevice, st));
        XC_CUDA_TRY(ctx, cudaEventRecord(ctx->stage_ev[b], st));
    }
    // the staging buffers belong to the context: the next call may overwrite them, so wait for the last DMAs here
    for (int b = 0; b < 2 && b < c; ++b) XC_CUDA_TRY(ctx, cudaEventSynchronize(ctx->stage_ev[b]));
    return XC_OK;
}
