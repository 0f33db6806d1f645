// Weighted per-instance top-k (dense + CSR), thresholding, dense scatter.
// Replaces xcolumns/weighted_prediction.py:25-88 and numba_csr_functions.py:456-484, 586-629.
#include "xc_scan.cuh"

namespace {

constexpr int kThreads = 256;

template <typename TE, typename G, int R, class Xf = XfMulAdd<G>>
__global__ void __launch_bounds__(kThreads, (R == 1 && sizeof(G) == 4) ? 6 : 1)
topk_dense_kernel(const TE *__restrict__ eta, int64_t n_rows, int64_t m, int64_t ld,
                  const int32_t *__restrict__ rows, Xf xf, int k, int32_t *__restrict__ out_idx,
                  G *__restrict__ out_val, bool vec_ok)
{
    const int lane = lane_id();
    const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t grp = warp; grp * R < n_rows; grp += nwarps) {
        const TE *rp[R];
        int64_t oi[R];
        int dummy[R];
        WarpTopK<G> tk[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int64_t i = grp * R + r;
            oi[r] = i < n_rows ? i : -1;
            if (i >= n_rows) i = grp * R;  // re-scan a valid row, result discarded
            int64_t row = rows ? (int64_t)rows[i] : i;
            rp[r] = eta + row * ld;
            tk[r].init();
            dummy[r] = -1;
        }
        xc_scan_rows<TE, G, R, false, Xf>(rp, m, vec_ok, xf, tk, dummy, k);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int src = warp_rank_src(tk[r].idx, k);
            int j = __shfl_sync(XC_FULL, tk[r].idx, src);
            G v = __shfl_sync(XC_FULL, tk[r].val, src);
            if (oi[r] >= 0 && lane < k) {
                out_idx[oi[r] * k + lane] = j;
                if (out_val) out_val[oi[r] * k + lane] = v;
            }
        }
    }
}

// one warp per CSR row; stored entries only
template <typename T>
__global__ void __launch_bounds__(kThreads)
topk_csr_kernel(const T *__restrict__ data, const int32_t *__restrict__ indices,
                const int64_t *__restrict__ indptr, int64_t n_rows, const T *__restrict__ a,
                const T *__restrict__ b, int k, int32_t *__restrict__ out_idx, T *__restrict__ out_val)
{
    const int lane = lane_id();
    const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    XfMulAdd<T> xf{a, b};
    for (int64_t i = warp; i < n_rows; i += nwarps) {
        const int64_t s = indptr[i], e = indptr[i + 1];
        WarpTopK<T> tk;
        tk.init();
        // list entries carry the POSITION inside the row (ascending position == ascending label)
        for (int64_t q0 = s; q0 < e; q0 += 32) {
            int64_t q = q0 + lane;
            T g[1];
            g[0] = (T)NAN;
            if (q < e) g[0] = xf.template apply_one<T>(indices[q], data[q]);
            if (__any_sync(XC_FULL, tk.passes(g[0]))) xc_scan_insert<T, 1, false>(tk, g, q0 - s, 1, k, -1);
        }
        int src = warp_rank_src(tk.idx, k);
        int pos = __shfl_sync(XC_FULL, tk.idx, src);
        T v = __shfl_sync(XC_FULL, tk.val, src);
        if (lane < k) {
            bool ok = pos != 0x7fffffff;
            out_idx[i * k + lane] = ok ? indices[s + pos] : -1;
            if (out_val) out_val[i * k + lane] = ok ? v : (T)1;
        }
    }
}

template <typename TE, typename G>
__global__ void __launch_bounds__(kThreads)
threshold_dense_kernel(const TE *__restrict__ eta, int64_t n_rows, int64_t m, int64_t ld, XfMulAdd<G> xf, G th,
                       TE *__restrict__ out, int64_t ld_out)
{
    const int64_t total = n_rows * m;
    for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (int64_t)gridDim.x * kThreads) {
        int64_t i = t / m, j = t - i * m;
        G g = xf.template apply_one<TE>(j, eta[i * ld + j]);
        out[i * ld_out + j] = (g >= th) ? (TE)1 : (TE)0;
    }
}

template <typename TO, typename TV>
__global__ void __launch_bounds__(kThreads)
scatter_pred_kernel(const int32_t *__restrict__ pred_idx, const TV *__restrict__ val, int k, int64_t n_rows,
                    TO *__restrict__ out, int64_t ld_out)
{
    const int64_t total = n_rows * k;
    for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (int64_t)gridDim.x * kThreads) {
        int j = pred_idx[t];
        if (j >= 0) out[(t / k) * ld_out + j] = val ? (TO)val[t] : (TO)1;
    }
}

// ---- k > 32: one CTA per row, radix select -------------------------------------------------------------------
// The warp-list kernels hold the running top-k in 32 lanes.  Larger k (the sparsifier of the on-disk formats keeps
// the top 100 - 1000 scores of a row, experiments/utils.py:199-211) selects by value instead: the k-th largest
// gain of the row is found digit by digit (11-bit histograms in shared memory over an order-preserving integer
// image of the gain), then one ordered pass writes every label above it plus, for ties at the threshold, the
// lowest label ids -- in ascending label order, the compact-prediction layout.  The row is read once from HBM and
// then from L2 (a row is at most a few hundred KB).
constexpr int kSelThreads = 512;
constexpr int kSelDigit = 11;
constexpr int kSelBins = 1 << kSelDigit;

__device__ __forceinline__ uint32_t sel_key(float g)
{
    const uint32_t u = __float_as_uint(g + 0.0f);
    if (g != g) return 0u;   // NaN sorts last
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ uint64_t sel_key(double g)
{
    const uint64_t u = (uint64_t)__double_as_longlong(g + 0.0);
    if (g != g) return 0ull;
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
template <typename G> struct SelKey;
template <> struct SelKey<float> { using type = uint32_t; static constexpr int bits = 32; };
template <> struct SelKey<double> { using type = uint64_t; static constexpr int bits = 64; };

// inclusive block scan of one int per thread (kSelThreads threads); returns the block total in *total
__device__ __forceinline__ int sel_block_scan(int v, int *warp_sums, int *total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(XC_FULL, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        int w = lane < kSelThreads / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(XC_FULL, w, o);
            if (lane >= o) w += y;
        }
        if (lane < kSelThreads / 32) warp_sums[lane] = w;
    }
    __syncthreads();
    const int base = wid ? warp_sums[wid - 1] : 0;
    *total = warp_sums[kSelThreads / 32 - 1];
    __syncthreads();   // warp_sums is reused by the next scan
    return base + x;
}

template <typename TE, typename G>
__global__ void __launch_bounds__(kSelThreads)
topk_select_kernel(const TE *__restrict__ eta, int64_t n_rows, int64_t m, int64_t ld, const int32_t *__restrict__ rows,
                   XfMulAdd<G> xf, int k, int32_t *__restrict__ out_idx, G *__restrict__ out_val)
{
    using Key = typename SelKey<G>::type;
    __shared__ int hist[kSelBins];
    __shared__ int warp_sums[kSelThreads / 32];
    __shared__ Key s_prefix;
    __shared__ int s_need;
    for (int64_t i = blockIdx.x; i < n_rows; i += gridDim.x) {
        const int64_t row = rows ? (int64_t)rows[i] : i;
        const TE *rp = eta + row * ld;
        // ---- the k-th largest key, one digit per pass
        Key prefix = 0, known = 0;   // `known`: mask of the digits fixed so far
        int need = k;                // how many of the elements matching the prefix are still wanted
        for (int shift = SelKey<G>::bits - kSelDigit; ; shift -= kSelDigit) {
            const int sh = shift < 0 ? 0 : shift;
            const int width = shift < 0 ? kSelDigit + shift : kSelDigit;   // the last digit may be narrower
            for (int b = threadIdx.x; b < kSelBins; b += kSelThreads) hist[b] = 0;
            __syncthreads();
            for (int64_t c = threadIdx.x; c < m; c += kSelThreads) {
                const Key key = sel_key(xf.template apply_one<TE>(c, rp[c]));
                if ((key & known) == prefix) atomicAdd(&hist[(int)((key >> sh) & (Key)((1 << width) - 1))], 1);
            }
            __syncthreads();
            if (threadIdx.x == 0) {   // walk the bins from the top: 2048 adds, negligible next to the pass over the row
                int acc = 0, d = (1 << width) - 1;
                for (; d > 0; --d) {
                    if (acc + hist[d] >= need) break;
                    acc += hist[d];
                }
                s_prefix = prefix | ((Key)d << sh);
                s_need = need - acc;
            }
            __syncthreads();
            prefix = s_prefix;
            need = s_need;
            known |= (Key)((1 << width) - 1) << sh;
            __syncthreads();
            if (sh == 0) break;
        }
        // prefix is now the exact key of the k-th largest gain; `need` of the elements equal to it are taken
        // ---- ordered output: labels ascending
        int written = 0, eq_seen = 0;
        for (int64_t c0 = 0; c0 < m; c0 += kSelThreads) {
            const int64_t c = c0 + threadIdx.x;
            G g = (G)0;
            Key key = 0;
            if (c < m) {
                g = xf.template apply_one<TE>(c, rp[c]);
                key = sel_key(g);
            }
            const int is_eq = (c < m && key == prefix) ? 1 : 0;
            int tot_eq;
            const int eq_rank = eq_seen + sel_block_scan(is_eq, warp_sums, &tot_eq) - is_eq;
            const int take = (c < m && (key > prefix || (is_eq && eq_rank < need))) ? 1 : 0;
            int tot_take;
            const int pos = written + sel_block_scan(take, warp_sums, &tot_take) - take;
            if (take && pos < k) {
                out_idx[i * k + pos] = (int32_t)c;
                if (out_val) out_val[i * k + pos] = g;
            }
            written += tot_take;
            eq_seen += tot_eq;
        }
        __syncthreads();
    }
}

template <typename K>
int grid_for(xc_ctx *ctx, K kernel, int64_t work_warps)
{
    // Balanced persistent grid: every warp gets the same number of row tasks (+-1).  With W warps
    // resident per full wave and T tasks, waves = ceil(T / W) and only ceil(T / waves) warps are
    // launched, spread evenly over the SMs -- instead of a full first wave and a mostly empty last
    // one (T = 1.48 W used to cost two full task times).
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0);
    if (per_sm < 1) per_sm = 1;
    const int wpc = kThreads / 32;
    const int64_t full_warps = (int64_t)ctx->sm_count * per_sm * wpc;
    if (work_warps < 1) work_warps = 1;
    const int64_t waves = (work_warps + full_warps - 1) / full_warps;
    const int64_t warps = (work_warps + waves - 1) / waves;
    int64_t grid = (warps + wpc - 1) / wpc;
    const int64_t cap = (int64_t)ctx->sm_count * per_sm;
    return (int)(grid < cap ? grid : cap);
}

template <typename TE, typename G>
int launch_topk_dense(xc_ctx *ctx, const void *eta, int64_t n_rows, int64_t m, int64_t ld, const int32_t *rows,
                      const void *a, const void *b, int k, int32_t *out_idx, void *out_val, cudaStream_t st)
{
    constexpr int V = 16 / sizeof(TE);
    bool vec_ok = xc_aligned16(eta) && (ld % V == 0);
    XfMulAdd<G> xf{(const G *)a, (const G *)b};
    // R rows per warp amortise the coefficient loads; small problems use R = 1 to fill the GPU
    // one row per warp unless wide coefficient vectors (a / b beyond L1) make register reuse pay;
        // see dense_rows_per_warp in bca_batched.cu for the measurement
    const int64_t coef_bytes = ((a ? 1 : 0) + (b ? 1 : 0)) * m * (int64_t)sizeof(G);
    const int rr = coef_bytes <= 160 * 1024 ? 1 : (coef_bytes <= 512 * 1024 ? 2 : 4);
    if (vec_ok && a && b && xc_aligned16(a) && xc_aligned16(b) && coef_bytes <= 1024 * 1024) {
        // both vectors present and aligned: 128-bit coefficient loads, one row per warp (see XfMulAddVec)
        XfMulAddVec<G> xfv{(const G *)a, (const G *)b};
        auto kern = topk_dense_kernel<TE, G, 1, XfMulAddVec<G>>;
        int grid = grid_for(ctx, kern, n_rows);
        kern<<<grid, kThreads, 0, st>>>((const TE *)eta, n_rows, m, ld, rows, xfv, k, out_idx, (G *)out_val, vec_ok);
    } else if (rr == 4) {
        auto kern = topk_dense_kernel<TE, G, 4>;
        int grid = grid_for(ctx, kern, (n_rows + 3) / 4);
        kern<<<grid, kThreads, 0, st>>>((const TE *)eta, n_rows, m, ld, rows, xf, k, out_idx, (G *)out_val, vec_ok);
    } else if (rr == 2) {
        auto kern = topk_dense_kernel<TE, G, 2>;
        int grid = grid_for(ctx, kern, (n_rows + 1) / 2);
        kern<<<grid, kThreads, 0, st>>>((const TE *)eta, n_rows, m, ld, rows, xf, k, out_idx, (G *)out_val, vec_ok);
    } else {
        auto kern = topk_dense_kernel<TE, G, 1>;
        int grid = grid_for(ctx, kern, n_rows);
        kern<<<grid, kThreads, 0, st>>>((const TE *)eta, n_rows, m, ld, rows, xf, k, out_idx, (G *)out_val, vec_ok);
    }
    XC_LAUNCHED(ctx);
    return XC_OK;
}

}  // namespace

extern "C" int xc_topk_dense(xc_ctx *ctx, const void *eta, int eta_dtype, int64_t n_rows, int64_t m, int64_t ld,
                             const int32_t *rows, const void *a, const void *b, int g_dtype, int k,
                             int32_t *out_idx, void *out_val, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !eta || !out_idx || n_rows < 0 || m <= 0 || ld < m) return XC_ERR_INVALID;
    if (k < 1 || k > m) return XC_ERR_INVALID;
    if (n_rows == 0) return XC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (k > 32) {   // one CTA per row, radix select (any k <= m)
        if (m > 0x7fffffffLL) return XC_ERR_UNSUPPORTED;
        const int64_t cap = (int64_t)ctx->sm_count * 4;
        const int grid = (int)(n_rows < cap ? n_rows : cap);
#define XC_SEL(TE, G)                                                                                              \
    topk_select_kernel<TE, G><<<grid, kSelThreads, 0, st>>>((const TE *)eta, n_rows, m, ld, rows,                     \
                                                            XfMulAdd<G>{(const G *)a, (const G *)b}, k, out_idx,    \
                                                            (G *)out_val)
        if (eta_dtype == XC_F32 && g_dtype == XC_F32) XC_SEL(float, float);
        else if (eta_dtype == XC_F32 && g_dtype == XC_F64) XC_SEL(float, double);
        else if (eta_dtype == XC_F64 && g_dtype == XC_F64) XC_SEL(double, double);
        else return XC_ERR_UNSUPPORTED;
#undef XC_SEL
        XC_LAUNCHED(ctx);
        return XC_OK;
    }
    if (eta_dtype == XC_F32 && g_dtype == XC_F32)
        return launch_topk_dense<float, float>(ctx, eta, n_rows, m, ld, rows, a, b, k, out_idx, out_val, st);
    if (eta_dtype == XC_F32 && g_dtype == XC_F64)
        return launch_topk_dense<float, double>(ctx, eta, n_rows, m, ld, rows, a, b, k, out_idx, out_val, st);
    if (eta_dtype == XC_F64 && g_dtype == XC_F64)
        return launch_topk_dense<double, double>(ctx, eta, n_rows, m, ld, rows, a, b, k, out_idx, out_val, st);
    return XC_ERR_UNSUPPORTED;
}

extern "C" int xc_topk_csr(xc_ctx *ctx, const void *data, int dtype, const int32_t *indices, const int64_t *indptr,
                           int64_t n_rows, const void *a, const void *b, int k, int32_t *out_idx, void *out_val,
                           void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !indptr || !out_idx || n_rows < 0) return XC_ERR_INVALID;
    if (k < 1 || k > 32) return XC_ERR_INVALID;
    if (n_rows == 0) return XC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == XC_F32) {
        auto kern = topk_csr_kernel<float>;
        int grid = grid_for(ctx, kern, n_rows);
        kern<<<grid, kThreads, 0, st>>>((const float *)data, indices, indptr, n_rows, (const float *)a,
                                        (const float *)b, k, out_idx, (float *)out_val);
    } else if (dtype == XC_F64) {
        auto kern = topk_csr_kernel<double>;
        int grid = grid_for(ctx, kern, n_rows);
        kern<<<grid, kThreads, 0, st>>>((const double *)data, indices, indptr, n_rows, (const double *)a,
                                        (const double *)b, k, out_idx, (double *)out_val);
    } else {
        return XC_ERR_UNSUPPORTED;
    }
    XC_LAUNCHED(ctx);
    return XC_OK;
}

extern "C" int xc_threshold_dense(xc_ctx *ctx, const void *eta, int eta_dtype, int64_t n_rows, int64_t m,
                                  int64_t ld, const void *a, const void *b, int g_dtype, double th, void *out,
                                  int64_t ld_out, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !eta || !out || n_rows < 0 || m <= 0 || ld < m || ld_out < m) return XC_ERR_INVALID;
    if (n_rows == 0) return XC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t blocks64 = (n_rows * m + kThreads - 1) / kThreads;
    int grid = (int)(blocks64 < (int64_t)ctx->sm_count * 16 ? blocks64 : (int64_t)ctx->sm_count * 16);
    if (eta_dtype == XC_F32 && g_dtype == XC_F32) {
        XfMulAdd<float> xf{(const float *)a, (const float *)b};
        threshold_dense_kernel<float, float><<<grid, kThreads, 0, st>>>((const float *)eta, n_rows, m, ld, xf,
                                                                         (float)th, (float *)out, ld_out);
    } else if (eta_dtype == XC_F32 && g_dtype == XC_F64) {
        XfMulAdd<double> xf{(const double *)a, (const double *)b};
        threshold_dense_kernel<float, double><<<grid, kThreads, 0, st>>>((const float *)eta, n_rows, m, ld, xf, th,
                                                                          (float *)out, ld_out);
    } else if (eta_dtype == XC_F64 && g_dtype == XC_F64) {
        XfMulAdd<double> xf{(const double *)a, (const double *)b};
        threshold_dense_kernel<double, double><<<grid, kThreads, 0, st>>>((const double *)eta, n_rows, m, ld, xf,
                                                                           th, (double *)out, ld_out);
    } else {
        return XC_ERR_UNSUPPORTED;
    }
    XC_LAUNCHED(ctx);
    return XC_OK;
}

extern "C" int xc_scatter_pred_dense(xc_ctx *ctx, const int32_t *pred_idx, const void *val, int val_dtype, int k,
                                     int64_t n_rows, void *out, int out_dtype, int64_t ld_out, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !pred_idx || !out || k < 1 || n_rows < 0) return XC_ERR_INVALID;
    if (n_rows == 0) return XC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t blocks64 = (n_rows * k + kThreads - 1) / kThreads;
    int grid = (int)(blocks64 < (int64_t)ctx->sm_count * 16 ? blocks64 : (int64_t)ctx->sm_count * 16);
#define XC_SC(TO, TV) \
    scatter_pred_kernel<TO, TV><<<grid, kThreads, 0, st>>>(pred_idx, (const TV *)val, k, n_rows, (TO *)out, ld_out)
    if (out_dtype == XC_F32 && (val == nullptr || val_dtype == XC_F32)) XC_SC(float, float);
    else if (out_dtype == XC_F32 && val_dtype == XC_F64) XC_SC(float, double);
    else if (out_dtype == XC_F64 && (val == nullptr || val_dtype == XC_F64)) XC_SC(double, double);
    else if (out_dtype == XC_F64 && val_dtype == XC_F32) XC_SC(double, float);
    else return XC_ERR_UNSUPPORTED;
#undef XC_SC
    XC_LAUNCHED(ctx);
    return XC_OK;
}
