// Frank-Wolfe iterate: fused weighted top-k + confusion accumulation, closed-form metric
// gradient, uniform line search over the step size.  Replaces the inner loop of
// xcolumns/frank_wolfe.py:589-670 (predict_weighted_per_instance :601, calculate_confusion_matrix
// :604, autograd gradient :591-596, _find_best_alpha/uniform_search :615 + utils.py:174-184).
#include "xc_scan.cuh"

namespace {

constexpr int kThreads = 256;

// ---- fused iterate, dense ------------------------------------------------------------------------
// gains = eta * a + b with separate IEEE multiply/add in the dtype numpy would use
// (weighted_prediction.py:37-41), top-k per row, then for the k selected labels
//   tp[j] += y_true[i][j],  cnt[j] += 1        (float64 atomics)
template <typename TE, int R>
__global__ void __launch_bounds__(kThreads)
fw_iterate_dense_kernel(const TE *__restrict__ eta, int64_t n, int64_t m, int64_t ld, const TE *__restrict__ y_true,
                        int64_t ld_true, XfMulAdd<TE> xf, int k, double *tp, double *cnt,
                        int32_t *__restrict__ pred_idx, bool vec_ok)
{
    const int lane = lane_id();
    const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t grp = warp; grp * R < n; grp += nwarps) {
        const TE *rp[R];
        int64_t row_id[R];
        int dummy[R];
        WarpTopK<TE> tk[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int64_t i = grp * R + r;
            row_id[r] = i < n ? i : -1;
            if (i >= n) i = grp * R;
            rp[r] = eta + i * ld;
            tk[r].init();
            dummy[r] = -1;
        }
        xc_scan_rows<TE, TE, R, false>(rp, m, vec_ok, xf, tk, dummy, k);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (row_id[r] < 0) continue;
            int j = tk[r].idx;
            if (lane < k && j != 0x7fffffff) {
                atomicAdd(tp + j, (double)__ldg(y_true + row_id[r] * ld_true + j));
                atomicAdd(cnt + j, 1.0);
            }
            if (pred_idx) {
                int src = warp_rank_src(j, k);
                int v = __shfl_sync(XC_FULL, j, src);
                if (lane < k) pred_idx[row_id[r] * k + lane] = v == 0x7fffffff ? -1 : v;
            }
        }
    }
}

__device__ __forceinline__ int64_t csr_find(const int32_t *idx, int64_t s, int64_t e, int j)
{
    while (s < e) {
        int64_t mid = (s + e) >> 1;
        int v = idx[mid];
        if (v == j) return mid;
        if (v < j) s = mid + 1; else e = mid;
    }
    return -1;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
fw_iterate_csr_kernel(const T *__restrict__ data, const int32_t *__restrict__ indices,
                      const int64_t *__restrict__ indptr, int64_t n, const T *__restrict__ t_data,
                      const int32_t *__restrict__ t_idx, const int64_t *__restrict__ t_ptr, XfMulAdd<T> xf, int k,
                      double *tp, double *cnt, int32_t *__restrict__ pred_idx)
{
    const int lane = lane_id();
    const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t i = warp; i < n; i += nwarps) {
        const int64_t s = indptr[i], e = indptr[i + 1];
        WarpTopK<T> tk;
        tk.init();
        for (int64_t q0 = s; q0 < e; q0 += 32) {
            int64_t q = q0 + lane;
            T g[1];
            g[0] = (T)NAN;
            if (q < e) g[0] = xf.template apply_one<T>(indices[q], data[q]);
            if (__any_sync(XC_FULL, tk.passes(g[0]))) xc_scan_insert<T, 1, false>(tk, g, q0 - s, 1, k, -1);
        }
        int src = warp_rank_src(tk.idx, k);
        int pos = __shfl_sync(XC_FULL, tk.idx, src);
        int j = pos == 0x7fffffff ? -1 : indices[s + pos];
        if (lane < k && j >= 0) {
            int64_t y = csr_find(t_idx, t_ptr[i], t_ptr[i + 1], j);
            if (y >= 0) atomicAdd(tp + j, (double)t_data[y]);
            atomicAdd(cnt + j, 1.0);
        }
        if (pred_idx && lane < k) pred_idx[i * k + lane] = j;
    }
}

// ---- confusion vectors of the iterate: C_i = [tp, fp, fn, tn] (4 stacked m-vectors) ----------------
// fp = cnt - tp, fn = colsum(y_true) - tp, optional /n, tn = -tp - fp - fn + (1 | n)  or -1
// (confusion_matrix.py:386-399)
__global__ void __launch_bounds__(kThreads)
fw_make_conf_kernel(const double *tp_raw, const double *cnt, const double *colsum, int64_t m, double n, int normalize,
                    int skip_tn, double *Ci)
{
    int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (j >= m) return;
    double t = tp_raw[j], f = cnt[j] - t, g = colsum[j] - t;
    if (normalize) { t = t / n; f = f / n; g = g / n; }
    Ci[j] = t;
    Ci[m + j] = f;
    Ci[2 * m + j] = g;
    Ci[3 * m + j] = skip_tn ? -1.0 : ((-t - f) - g) + (normalize ? 1.0 : n);
}

// ---- metric value + next classifier -------------------------------------------------------------------
struct Grad4 { double v, gtp, gfp, gfn, gtn; };

__device__ __forceinline__ Grad4 metric_grad(int metric, double tp, double fp, double fn, double tn, double c,
                                             double b2, double e)
{
    Grad4 r;
    r.gtn = 0.0;
    if (metric == XC_METRIC_FBETA) {
        double D = b2 * (tp + fp) + tp + fn + e;
        r.v = c * tp / D;
        r.gtp = c * (D - tp * c) / (D * D);
        r.gfp = -c * tp * b2 / (D * D);
        r.gfn = -c * tp / (D * D);
    } else if (metric == XC_METRIC_PRECISION) {
        double D = tp + fp + e;
        r.v = tp / D;
        r.gtp = (fp + e) / (D * D);
        r.gfp = -tp / (D * D);
        r.gfn = 0.0;
    } else if (metric == XC_METRIC_RECALL) {
        double D = tp + fn + e;
        r.v = tp / D;
        r.gtp = (fn + e) / (D * D);
        r.gfp = 0.0;
        r.gfn = -tp / (D * D);
    } else if (metric == XC_METRIC_JACCARD) {
        double D = tp + fp + fn + e;
        r.v = tp / D;
        r.gtp = (fp + fn + e) / (D * D);
        r.gfp = -tp / (D * D);
        r.gfn = r.gfp;
    } else {
        double Dp = tp + fn + e, Dn = tn + fp + e;
        double tpr = tp / Dp, tnr = tn / Dn;
        double tpr_tp = (fn + e) / (Dp * Dp), tpr_fn = -tp / (Dp * Dp);
        double tnr_fp = -tn / (Dn * Dn), tnr_tn = (fp + e) / (Dn * Dn);
        double fpw, fnw;  // d metric / d tpr, d metric / d tnr
        if (metric == XC_METRIC_BALANCED_ACC) {
            r.v = (tpr + tnr) / 2.0;
            fpw = 0.5; fnw = 0.5;
        } else if (metric == XC_METRIC_GMEAN) {
            r.v = sqrt(tpr * tnr);
            fpw = 0.5 * tnr / r.v; fnw = 0.5 * tpr / r.v;
        } else {
            double s = tpr + tnr;
            r.v = 2.0 * tpr * tnr / s;
            fpw = 2.0 * tnr * tnr / (s * s); fnw = 2.0 * tpr * tpr / (s * s);
        }
        r.gtp = fpw * tpr_tp;
        r.gfn = fpw * tpr_fn;
        r.gfp = fnw * tnr_fp;
        r.gtn = fnw * tnr_tn;
    }
    return r;
}

__global__ void __launch_bounds__(1024)
fw_metric_grad_kernel(xc_metric_params p, const double *tp, const double *fp, const double *fn, const double *tn,
                      int64_t m, float *a_out, float *b_out, double *value)
{
    __shared__ double sm[32];
    double s = 0.0;
    const double sgn = p.maximize ? 1.0 : -1.0;
    const double inv_m = 1.0 / (double)m;
    for (int64_t j = threadIdx.x; j < m; j += 1024) {
        Grad4 g = metric_grad(p.metric, tp[j], fp[j], fn[j], tn ? tn[j] : -1.0, p.c1, p.beta2, p.eps);
        s += g.v;
        if (a_out) {
            double gtp = g.gtp * inv_m, gfp = g.gfp * inv_m, gfn = g.gfn * inv_m, gtn = g.gtn * inv_m;
            a_out[j] = (float)(sgn * (((gtp - gfp) - gfn) + gtn));  // frank_wolfe.py:595
            b_out[j] = (float)(sgn * (gfp - gtn));                  // :596
        }
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = warp_sum(sm[threadIdx.x]);
        if (threadIdx.x == 0 && value) *value = v * inv_m;
    }
}

// ---- uniform line search -------------------------------------------------------------------------------
// block = AT consecutive grid points; threads stride over the labels, keeping AT running sums.
constexpr int AT = 16;

__global__ void __launch_bounds__(kThreads)
fw_alpha_eval_kernel(xc_metric_params p, const double *__restrict__ C, const double *__restrict__ Ci, int64_t m,
                     const double *__restrict__ alphas, int64_t n_alphas, double *__restrict__ vals)
{
    __shared__ double sm[AT][kThreads / 32];
    const int64_t q0 = (int64_t)blockIdx.x * AT;  // q = 0 is alpha = 0, q >= 1 is alphas[q-1]
    double al[AT], acc[AT];
#pragma unroll
    for (int t = 0; t < AT; ++t) {
        int64_t q = q0 + t;
        al[t] = (q == 0 || q > n_alphas) ? 0.0 : alphas[q - 1];
        acc[t] = 0.0;
    }
    const bool use_tn = p.metric >= XC_METRIC_BALANCED_ACC;
    for (int64_t j = threadIdx.x; j < m; j += kThreads) {
        const double tp = C[j], fp = C[m + j], fn = C[2 * m + j], tn = use_tn ? C[3 * m + j] : 0.0;
        const double tpi = Ci[j], fpi = Ci[m + j], fni = Ci[2 * m + j], tni = use_tn ? Ci[3 * m + j] : 0.0;
#pragma unroll
        for (int t = 0; t < AT; ++t) {
            const double a1 = al[t], a0 = 1.0 - a1;  // frank_wolfe.py:393-398
            acc[t] += xc_binary_metric(p.metric, a0 * tp + a1 * tpi, a0 * fp + a1 * fpi, a0 * fn + a1 * fni,
                                       a0 * tn + a1 * tni, p.c1, p.beta2, p.eps);
        }
    }
#pragma unroll
    for (int t = 0; t < AT; ++t) {
        double v = warp_sum(acc[t]);
        if ((threadIdx.x & 31) == 0) sm[t][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < AT) {
        double v = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) v += sm[threadIdx.x][w];
        int64_t q = q0 + threadIdx.x;
        if (q <= n_alphas) vals[q] = v / (double)m;
    }
}

// first strict maximum over q = 0..n_alphas (utils.py:177-184)
__global__ void __launch_bounds__(1024)
fw_alpha_pick_kernel(const double *__restrict__ vals, const double *__restrict__ alphas, int64_t n_alphas,
                     double *result)
{
    __shared__ double sv[32];
    __shared__ long long sq[32];
    double bv = -INFINITY;
    long long bq = 0x7fffffffffffffffLL;
    for (int64_t q = threadIdx.x; q <= n_alphas; q += 1024) {
        double v = vals[q];
        if (v > bv || (v == bv && q < bq)) { bv = v; bq = q; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(XC_FULL, bv, o);
        long long oq = __shfl_xor_sync(XC_FULL, bq, o);
        if (ov > bv || (ov == bv && oq < bq)) { bv = ov; bq = oq; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; sq[threadIdx.x >> 5] = bq; }
    __syncthreads();
    if (threadIdx.x < 32) {
        bv = sv[threadIdx.x]; bq = sq[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_xor_sync(XC_FULL, bv, o);
            long long oq = __shfl_xor_sync(XC_FULL, bq, o);
            if (ov > bv || (ov == bv && oq < bq)) { bv = ov; bq = oq; }
        }
        if (threadIdx.x == 0) {
            // NaN everywhere -> keep alpha = 0 like the reference (no score > best_val)
            if (bq == 0x7fffffffffffffffLL) { bq = 0; bv = vals[0]; }
            result[0] = bq == 0 ? 0.0 : alphas[bq - 1];
            result[1] = bv;
        }
    }
}

__global__ void __launch_bounds__(kThreads)
fw_combine_kernel(double *C, const double *Ci, int64_t m4, const double *alpha_dev)
{
    const double a1 = *alpha_dev;
    int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (j < m4) C[j] = (1.0 - a1) * C[j] + a1 * Ci[j];  // frank_wolfe.py:633-636
}

template <typename K>
int grid_for(xc_ctx *ctx, K kernel, int64_t work_warps)
{
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0);
    if (per_sm < 1) per_sm = 1;
    int64_t full = (int64_t)ctx->sm_count * per_sm;
    int64_t need = (work_warps + (kThreads / 32) - 1) / (kThreads / 32);
    if (need < 1) need = 1;
    return (int)(need < full ? need : full);
}

template <typename TE>
int launch_fw_dense(xc_ctx *ctx, const void *eta, int64_t n, int64_t m, int64_t ld, const void *y_true,
                    int64_t ld_true, const void *a, const void *b, int k, double *tp, double *cnt, int32_t *pred_idx,
                    cudaStream_t st)
{
    constexpr int V = 16 / sizeof(TE);
    bool vec_ok = xc_aligned16(eta) && (ld % V == 0);
    XfMulAdd<TE> xf{(const TE *)a, (const TE *)b};
    const int64_t coef_bytes = 2 * m * (int64_t)sizeof(TE);
    const int rr = coef_bytes <= 160 * 1024 ? 1 : (coef_bytes <= 512 * 1024 ? 2 : 4);
#define XC_GO(R)                                                                                             \
    {                                                                                                        \
        auto kern = fw_iterate_dense_kernel<TE, R>;                                                          \
        int grid = grid_for(ctx, kern, (n + R - 1) / R);                                                     \
        kern<<<grid, kThreads, 0, st>>>((const TE *)eta, n, m, ld, (const TE *)y_true, ld_true, xf, k, tp,   \
                                        cnt, pred_idx, vec_ok);                                              \
    }
    if (rr == 4) XC_GO(4)
    else if (rr == 2) XC_GO(2)
    else XC_GO(1)
#undef XC_GO
    XC_LAUNCHED(ctx);
    return XC_OK;
}

}  // namespace

extern "C" int xc_fw_iterate_dense(xc_ctx *ctx, const void *eta, int dtype, int64_t n, int64_t m, int64_t ld,
                                   const void *y_true, int64_t ld_true, const void *a, const void *b, int k,
                                   double *tp, double *cnt, int32_t *pred_idx, void *stream)
{
    if (!ctx || !eta || !y_true || !tp || !cnt || n <= 0 || m <= 0 || ld < m || ld_true < m) return XC_ERR_INVALID;
    if (k < 1 || k > 32 || k > m) return XC_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    XC_CUDA_TRY(ctx, cudaMemsetAsync(tp, 0, sizeof(double) * m, st));
    XC_CUDA_TRY(ctx, cudaMemsetAsync(cnt, 0, sizeof(double) * m, st));
    if (dtype == XC_F32) return launch_fw_dense<float>(ctx, eta, n, m, ld, y_true, ld_true, a, b, k, tp, cnt, pred_idx, st);
    if (dtype == XC_F64) return launch_fw_dense<double>(ctx, eta, n, m, ld, y_true, ld_true, a, b, k, tp, cnt, pred_idx, st);
    return XC_ERR_UNSUPPORTED;
}

extern "C" int xc_fw_iterate_csr(xc_ctx *ctx, const void *data, int dtype, const int32_t *indices,
                                 const int64_t *indptr, int64_t n, int64_t m, const void *t_data,
                                 const int32_t *t_idx, const int64_t *t_ptr, const void *a, const void *b, int k,
                                 double *tp, double *cnt, int32_t *pred_idx, void *stream)
{
    if (!ctx || !indptr || !t_ptr || !tp || !cnt || n <= 0 || m <= 0) return XC_ERR_INVALID;
    if (k < 1 || k > 32) return XC_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    XC_CUDA_TRY(ctx, cudaMemsetAsync(tp, 0, sizeof(double) * m, st));
    XC_CUDA_TRY(ctx, cudaMemsetAsync(cnt, 0, sizeof(double) * m, st));
    if (dtype == XC_F32) {
        auto kern = fw_iterate_csr_kernel<float>;
        int grid = grid_for(ctx, kern, n);
        XfMulAdd<float> xf{(const float *)a, (const float *)b};
        kern<<<grid, kThreads, 0, st>>>((const float *)data, indices, indptr, n, (const float *)t_data, t_idx, t_ptr,
                                        xf, k, tp, cnt, pred_idx);
    } else if (dtype == XC_F64) {
        auto kern = fw_iterate_csr_kernel<double>;
        int grid = grid_for(ctx, kern, n);
        XfMulAdd<double> xf{(const double *)a, (const double *)b};
        kern<<<grid, kThreads, 0, st>>>((const double *)data, indices, indptr, n, (const double *)t_data, t_idx, t_ptr,
                                        xf, k, tp, cnt, pred_idx);
    } else {
        return XC_ERR_UNSUPPORTED;
    }
    XC_LAUNCHED(ctx);
    return XC_OK;
}

extern "C" int xc_fw_make_conf(xc_ctx *ctx, const double *tp_raw, const double *cnt, const double *colsum, int64_t m,
                               double n, int normalize, int skip_tn, double *Ci, void *stream)
{
    if (!ctx || !tp_raw || !cnt || !colsum || !Ci || m <= 0) return XC_ERR_INVALID;
    fw_make_conf_kernel<<<(unsigned)((m + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        tp_raw, cnt, colsum, m, n, normalize, skip_tn, Ci);
    XC_LAUNCHED(ctx);
    return XC_OK;
}

extern "C" int xc_fw_metric_grad(xc_ctx *ctx, const xc_metric_params *p, const double *C, int64_t m, float *a_out,
                                 float *b_out, double *value_dev, void *stream)
{
    if (!ctx || !p || !C || m <= 0) return XC_ERR_INVALID;
    if ((a_out == nullptr) != (b_out == nullptr)) return XC_ERR_INVALID;
    if (p->metric < 0 || p->metric > XC_METRIC_HMEAN) return XC_ERR_INVALID;
    fw_metric_grad_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(*p, C, C + m, C + 2 * m, C + 3 * m, m, a_out, b_out,
                                                                value_dev);
    XC_LAUNCHED(ctx);
    return XC_OK;
}

extern "C" int xc_fw_alpha_search(xc_ctx *ctx, const xc_metric_params *p, const double *C, const double *Ci,
                                  int64_t m, const double *alphas_dev, int64_t n_alphas, double *vals_dev,
                                  double *result_dev, void *stream)
{
    if (!ctx || !p || !C || !Ci || !vals_dev || !result_dev || m <= 0 || n_alphas < 0) return XC_ERR_INVALID;
    if (n_alphas > 0 && !alphas_dev) return XC_ERR_INVALID;
    if (p->metric < 0 || p->metric > XC_METRIC_HMEAN) return XC_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned grid = (unsigned)((n_alphas + 1 + AT - 1) / AT);
    fw_alpha_eval_kernel<<<grid, kThreads, 0, st>>>(*p, C, Ci, m, alphas_dev, n_alphas, vals_dev);
    XC_LAUNCHED(ctx);
    fw_alpha_pick_kernel<<<1, 1024, 0, st>>>(vals_dev, alphas_dev, n_alphas, result_dev);
    XC_LAUNCHED(ctx);
    return XC_OK;
}

extern "C" int xc_fw_combine(xc_ctx *ctx, double *C, const double *Ci, int64_t m4, const double *alpha_dev,
                             void *stream)
{
    if (!ctx || !C || !Ci || !alpha_dev || m4 <= 0) return XC_ERR_INVALID;
    fw_combine_kernel<<<(unsigned)((m4 + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(C, Ci, m4,
                                                                                                          alpha_dev);
    XC_LAUNCHED(ctx);
    return XC_OK;
}
