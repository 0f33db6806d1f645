// Host-side helpers of the boundary (no device code): materialising the dense 0/1 prediction
// matrix the reference API returns for dense inputs.  For AmazonCat-13K-shape outputs this is a
// 16 GB host write (first-touch page faults + zeroing); it is spread over host threads.
#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

#include "xc_common.cuh"

namespace {

template <typename TO, typename TV>
void fill_rows(TO *out, int64_t ld, int64_t m, const int32_t *idx, const TV *val, int k, int64_t r0, int64_t r1)
{
    for (int64_t i = r0; i < r1; ++i) {
        TO *row = out + i * ld;
        std::memset(row, 0, sizeof(TO) * (size_t)m);
        for (int t = 0; t < k; ++t) {
            int j = idx[i * k + t];
            if (j >= 0 && j < m) row[j] = val ? (TO)val[i * k + t] : (TO)1;
        }
    }
}

template <typename TO, typename TV>
int fill_threads(void *out, int64_t n, int64_t m, int64_t ld, const int32_t *idx, const void *val, int k, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    int64_t chunk = (n + nthreads - 1) / nthreads;
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
        int64_t r0 = t * chunk, r1 = std::min(n, r0 + chunk);
        if (r0 >= r1) break;
        th.emplace_back(fill_rows<TO, TV>, (TO *)out, ld, m, idx, (const TV *)val, k, r0, r1);
    }
    for (auto &x : th) x.join();
    return XC_OK;
}

}  // namespace

extern "C" int xc_fill_pred_dense_host(void *out_host, int out_dtype, int64_t n, int64_t m, int64_t ld,
                                       const int32_t *idx_host, const void *val_host, int val_dtype, int k,
                                       int nthreads)
{
    if (!out_host || !idx_host || n < 0 || m <= 0 || ld < m || k < 1) return XC_ERR_INVALID;
    if (nthreads <= 0) {
        unsigned hc = std::thread::hardware_concurrency();
        nthreads = (int)std::min<unsigned>(hc ? hc : 1, 32);
    }
    if (n * m < (1 << 22)) nthreads = 1;
    if (out_dtype == XC_F32 && (!val_host || val_dtype == XC_F32)) return fill_threads<float, float>(out_host, n, m, ld, idx_host, val_host, k, nthreads);
    if (out_dtype == XC_F32 && val_dtype == XC_F64) return fill_threads<float, double>(out_host, n, m, ld, idx_host, val_host, k, nthreads);
    if (out_dtype == XC_F64 && (!val_host || val_dtype == XC_F64)) return fill_threads<double, double>(out_host, n, m, ld, idx_host, val_host, k, nthreads);
    if (out_dtype == XC_F64 && val_dtype == XC_F32) return fill_threads<double, float>(out_host, n, m, ld, idx_host, val_host, k, nthreads);
    return XC_ERR_UNSUPPORTED;
}
