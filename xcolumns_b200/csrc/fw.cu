// Frank-Wolfe iterate: fused weighted top-k + confusion accumulation, closed-form metric
// gradient, uniform line search over the step size.  Replaces the inner loop of
// xcolumns/frank_wolfe.py:589-670 (predict_weighted_per_instance :601, calculate_confusion_matrix
// :604, autograd gradient :591-596, _find_best_alpha/uniform_search :615 + utils.py:174-184).
#include <cstdlib>

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

__global__ void __launch_bounds__(256)
fw_metric_grad_kernel(xc_metric_params p, const double *tp, const double *fp, const double *fn, const double *tn,
                      int64_t m, float *a_out, float *b_out, double *value, double *partials, unsigned *counter)
{
    __shared__ double sm[8];
    double s = 0.0;
    const double sgn = p.maximize ? 1.0 : -1.0;
    const double inv_m = 1.0 / (double)m;
    for (int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x; j < m; j += (int64_t)gridDim.x * 256) {
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
    double bsum = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < 8; ++w) bsum += sm[w];
    double total;
    if (xc_grid_sum_last(bsum, partials, counter, &total) && value) *value = total * inv_m;
}

// ---- uniform line search -------------------------------------------------------------------------------
// The reference evaluates the metric of (1-alpha) C + alpha C_i at ~10^4 grid points and keeps the
// first strict maximum (utils.py:174-184): 10^4 x m IEEE float64 divisions per iteration (0.6-1.8 ms
// at m = 31 k, several times the streaming pass).  For the metrics of the form c*tp/D with D linear
// in the confusion entries (precision, recall, F-beta, Jaccard) the search runs in two stages:
//   1. every grid point in float32 from a per-label linearisation (T0 + a dT) / (D0 + a dD)
//      (2 FFMA + 1 MUFU.RCP + 1 FMUL + 1 FADD per term; measured 110 us.  A float64 variant with a
//      Newton-refined reciprocal was measured at 394 us: vector FP64 is the scarce resource here);
//   2. the grid points within 2e-5 (relative) of the float32 maximum -- a superset of every point
//      that can be the exact maximum, the float32 pass is accurate to ~2e-6 -- are re-evaluated
//      with the reference's exact float64 expression; the first strict maximum among them wins.
// If more than ALPHA_MAX_CAND points qualify (objective flat in alpha to 2e-5) the whole grid is
// evaluated exactly.  Metrics that use tn always take the exact path.
constexpr int AT_FULL = 16;         // grid points per block, float64 kernel over the whole grid
constexpr int AT_CAND = 4;          // ... over the candidate list (few points: spread them over the SMs)
constexpr int AT32 = 32;            // grid points per block, stage-1 kernel
constexpr int ALPHA_MAX_CAND = 2048;

struct AlphaCtl {
    int count;   // number of candidate slots (or n_alphas + 1 when full)
    int full;    // 1: evaluate the whole grid in float64
};

__global__ void __launch_bounds__(kThreads)
fw_alpha_prep_kernel(xc_metric_params p, const double *__restrict__ C, const double *__restrict__ Ci, int64_t m,
                     float4 *__restrict__ lin)
{
    int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (j >= m) return;
    const double tp = C[j], fp = C[m + j], fn = C[2 * m + j];
    const double tpi = Ci[j], fpi = Ci[m + j], fni = Ci[2 * m + j];
    double D0, D1;
    if (p.metric == XC_METRIC_PRECISION) { D0 = tp + fp + p.eps; D1 = tpi + fpi + p.eps; }
    else if (p.metric == XC_METRIC_RECALL) { D0 = tp + fn + p.eps; D1 = tpi + fni + p.eps; }
    else if (p.metric == XC_METRIC_JACCARD) { D0 = tp + fp + fn + p.eps; D1 = tpi + fpi + fni + p.eps; }
    else { D0 = p.beta2 * (tp + fp) + tp + fn + p.eps; D1 = p.beta2 * (tpi + fpi) + tpi + fni + p.eps; }
    lin[j] = make_float4((float)tp, (float)(tpi - tp), (float)D0, (float)(D1 - D0));
}

constexpr int ALPHA_LSPLIT = 4;  // label slices per grid point tile (blockIdx.y)

__global__ void __launch_bounds__(kThreads)
fw_alpha_evalfast_kernel(const float4 *__restrict__ lin, int64_t m, const double *__restrict__ alphas,
                         int64_t n_alphas, float scale, float *__restrict__ vals_fast)
{
    __shared__ float sm[AT32][kThreads / 32];
    const int64_t q0 = (int64_t)blockIdx.x * AT32;
    float al[AT32], acc[AT32];
#pragma unroll
    for (int t = 0; t < AT32; ++t) {
        int64_t q = q0 + t;
        al[t] = (q == 0 || q > n_alphas) ? 0.f : (float)alphas[q - 1];
        acc[t] = 0.f;
    }
    // <= m / (256 * ALPHA_LSPLIT) terms per float32 accumulator (30 at m = 31 k): worst-case
    // accumulation error ~2e-6 relative, well inside the 2e-5 candidate window of stage 2
    for (int64_t j = (int64_t)blockIdx.y * kThreads + threadIdx.x; j < m; j += (int64_t)kThreads * ALPHA_LSPLIT) {
        const float4 l = __ldg(lin + j);
#pragma unroll
        for (int t = 0; t < AT32; ++t)
            acc[t] += __fdividef(fmaf(al[t], l.y, l.x), fmaf(al[t], l.w, l.z));
    }
#pragma unroll
    for (int t = 0; t < AT32; ++t) {
        float v = acc[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(XC_FULL, v, o);
        if ((threadIdx.x & 31) == 0) sm[t][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < AT32) {
        float v = 0.f;
        for (int w = 0; w < kThreads / 32; ++w) v += sm[threadIdx.x][w];
        int64_t q = q0 + threadIdx.x;
        if (q <= n_alphas) atomicAdd(vals_fast + q, v * scale);
    }
}

// candidates = grid points within 2e-5 (relative) of the stage-1 maximum, in grid order
__global__ void __launch_bounds__(1024)
fw_alpha_cand_kernel(const float *__restrict__ vf, int64_t n_alphas, int *__restrict__ cand_q, AlphaCtl *ctl)
{
    __shared__ float s_max[32];
    __shared__ int s_cnt[1024];
    __shared__ int s_base;
    const int64_t total = n_alphas + 1;
    float mx = -INFINITY;
    for (int64_t q = threadIdx.x; q < total; q += 1024) mx = fmaxf(mx, vf[q]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(XC_FULL, mx, o));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = s_max[0];
    for (int w = 1; w < 32; ++w) mx = fmaxf(mx, s_max[w]);
    const float thr = mx - 2e-5f * fabsf(mx) - 1e-37f;
    // contiguous slice per thread so that the compacted list stays in grid order
    const int64_t per = (total + 1023) / 1024;
    const int64_t b = threadIdx.x * per, e = min(total, b + per);
    int cnt = 0;
    for (int64_t q = b; q < e; ++q) cnt += (vf[q] >= thr) || !(vf[q] == vf[q]);  // NaN: keep
    s_cnt[threadIdx.x] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int t = 0; t < 1024; ++t) { int c = s_cnt[t]; s_cnt[t] = run; run += c; }
        s_base = run;
        const bool full = run > ALPHA_MAX_CAND || !(mx == mx);
        ctl->full = full ? 1 : 0;
        ctl->count = full ? (int)total : run;
    }
    __syncthreads();
    if (s_base <= ALPHA_MAX_CAND) {
        int o = s_cnt[threadIdx.x];
        for (int64_t q = b; q < e; ++q)
            if ((vf[q] >= thr) || !(vf[q] == vf[q])) cand_q[o++] = (int)q;
    }
}

// float64 evaluation with the reference's expression (frank_wolfe.py:393-398).  Slot t evaluates grid
// point q = cand_q[t] (candidate mode) or q = t (full mode / ctl == nullptr); q = 0 is alpha = 0.
template <int AT>
__global__ void __launch_bounds__(kThreads)
fw_alpha_eval_kernel(xc_metric_params p, const double *__restrict__ C, const double *__restrict__ Ci, int64_t m,
                     const double *__restrict__ alphas, int64_t n_alphas, const int *__restrict__ cand_q,
                     const AlphaCtl *__restrict__ ctl, double *__restrict__ vals, int want_full)
{
    __shared__ double sm[AT][kThreads / 32];
    const int64_t s0 = (int64_t)blockIdx.x * AT;
    const bool full = ctl == nullptr || ctl->full;
    const int64_t count = ctl == nullptr ? n_alphas + 1 : ctl->count;
    if ((int)full != want_full || s0 >= count) return;  // the other launch handles this mode
    double al[AT], acc[AT];
#pragma unroll
    for (int t = 0; t < AT; ++t) {
        int64_t slot = s0 + t;
        int64_t q = slot < count ? (full ? slot : (int64_t)cand_q[slot]) : 0;
        al[t] = (q == 0) ? 0.0 : alphas[q - 1];
        acc[t] = 0.0;
    }
    const bool use_tn = p.metric >= XC_METRIC_BALANCED_ACC;
    // gridDim.y label slices (candidate mode): partial sums are combined with float64 atomics
    for (int64_t j = (int64_t)blockIdx.y * kThreads + threadIdx.x; j < m; j += (int64_t)kThreads * gridDim.y) {
        const double tp = C[j], fp = C[m + j], fn = C[2 * m + j], tn = use_tn ? C[3 * m + j] : 0.0;
        const double tpi = Ci[j], fpi = Ci[m + j], fni = Ci[2 * m + j], tni = use_tn ? Ci[3 * m + j] : 0.0;
#pragma unroll
        for (int t = 0; t < AT; ++t) {
            const double a1 = al[t], a0 = 1.0 - a1;
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
        int64_t slot = s0 + threadIdx.x;
        if (slot < count) {
            if (gridDim.y > 1) atomicAdd(vals + slot, v / (double)m);
            else vals[slot] = v / (double)m;
        }
    }
}

// first strict maximum over the evaluated slots (slots are in grid order; utils.py:177-184)
__global__ void __launch_bounds__(1024)
fw_alpha_pick_kernel(const double *__restrict__ vals, const double *__restrict__ alphas, int64_t n_alphas,
                     const int *__restrict__ cand_q, const AlphaCtl *__restrict__ ctl, double *result)
{
    __shared__ double sv[32];
    __shared__ long long sq[32];
    const bool full = ctl == nullptr || ctl->full;
    const int64_t count = ctl == nullptr ? n_alphas + 1 : ctl->count;
    double bv = -INFINITY;
    long long bq = 0x7fffffffffffffffLL;
    for (int64_t q = threadIdx.x; q < count; q += 1024) {
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
            const long long q = full ? bq : (long long)cand_q[bq];
            result[0] = q == 0 ? 0.0 : alphas[q - 1];
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

template <typename TE>
int launch_fw_dense(xc_ctx *ctx, const void *eta, int64_t n, int64_t m, int64_t ld, const void *y_true,
                    int64_t ld_true, const void *a, const void *b, int k, double *tp, double *cnt, int32_t *pred_idx,
                    cudaStream_t st)
{
    constexpr int V = 16 / sizeof(TE);
    bool vec_ok = xc_aligned16(eta) && (ld % V == 0);
    XfMulAdd<TE> xf{(const TE *)a, (const TE *)b};
    const int64_t coef_bytes = 2 * m * (int64_t)sizeof(TE);
    int rr = coef_bytes <= 160 * 1024 ? 1 : (coef_bytes <= 512 * 1024 ? 2 : 4);
    if (const char *e = getenv("XCOLUMNS_B200_DENSE_R")) {
        int v = atoi(e);
        if (v == 1 || v == 2 || v == 4) rr = v;
    }
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
    int64_t blocks = (m + 255) / 256;
    int grid = (int)(blocks < XC_RED_MAX_BLOCKS ? blocks : XC_RED_MAX_BLOCKS);
    if (grid > ctx->sm_count * 4) grid = ctx->sm_count * 4;
    fw_metric_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*p, C, C + m, C + 2 * m, C + 3 * m, m, a_out, b_out,
                                                                  value_dev, ctx->red_partials, ctx->red_counter);
    XC_LAUNCHED(ctx);
    return XC_OK;
}

extern "C" int64_t xc_fw_alpha_scratch_bytes(int64_t m, int64_t n_alphas)
{
    // exact values | stage-1 values | candidate list | control | per-label linearisation
    return (n_alphas + 1) * 8 + (n_alphas + 1) * 8 + ALPHA_MAX_CAND * 4 + 64 + m * 16 + 256;
}

extern "C" int xc_fw_alpha_search(xc_ctx *ctx, const xc_metric_params *p, const double *C, const double *Ci,
                                  int64_t m, const double *alphas_dev, int64_t n_alphas, double *vals_dev,
                                  double *result_dev, void *stream)
{
    if (!ctx || !p || !C || !Ci || !vals_dev || !result_dev || m <= 0 || n_alphas < 0) return XC_ERR_INVALID;
    if (n_alphas > 0 && !alphas_dev) return XC_ERR_INVALID;
    if (p->metric < 0 || p->metric > XC_METRIC_HMEAN) return XC_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    // carve the caller's scratch (xc_fw_alpha_scratch_bytes)
    uint8_t *base = reinterpret_cast<uint8_t *>(vals_dev);
    double *vals = vals_dev;
    size_t off = (size_t)(n_alphas + 1) * 8;
    float *vals_fast = reinterpret_cast<float *>(base + off);
    off += (size_t)(n_alphas + 1) * 8;
    int *cand_q = reinterpret_cast<int *>(base + off);
    off += ALPHA_MAX_CAND * 4;
    AlphaCtl *ctl = reinterpret_cast<AlphaCtl *>(base + off);
    off += 64;
    off = (off + 31) & ~(size_t)31;
    float4 *lin = reinterpret_cast<float4 *>(base + off);
    const unsigned grid_full = (unsigned)((n_alphas + 1 + AT_FULL - 1) / AT_FULL);
    const bool two_stage = p->metric <= XC_METRIC_JACCARD && n_alphas >= 256;
    if (two_stage) {
        fw_alpha_prep_kernel<<<(unsigned)((m + kThreads - 1) / kThreads), kThreads, 0, st>>>(*p, C, Ci, m, lin);
        XC_LAUNCHED(ctx);
        XC_CUDA_TRY(ctx, cudaMemsetAsync(vals_fast, 0, sizeof(float) * (size_t)(n_alphas + 1), st));
        XC_CUDA_TRY(ctx, cudaMemsetAsync(vals, 0, sizeof(double) * (size_t)ALPHA_MAX_CAND, st));
        const float scale = (float)((p->metric == XC_METRIC_FBETA ? p->c1 : 1.0) / (double)m);
        dim3 g32((unsigned)((n_alphas + 1 + AT32 - 1) / AT32), ALPHA_LSPLIT);
        fw_alpha_evalfast_kernel<<<g32, kThreads, 0, st>>>(lin, m, alphas_dev, n_alphas, scale, vals_fast);
        XC_LAUNCHED(ctx);
        fw_alpha_cand_kernel<<<1, 1024, 0, st>>>(vals_fast, n_alphas, cand_q, ctl);
        XC_LAUNCHED(ctx);
        // candidate mode: <= ALPHA_MAX_CAND slots, 2 per block; blocks past the count exit at once
        fw_alpha_eval_kernel<AT_CAND><<<dim3(ALPHA_MAX_CAND / AT_CAND, ALPHA_LSPLIT), kThreads, 0, st>>>(
            *p, C, Ci, m, alphas_dev, n_alphas, cand_q, ctl, vals, 0);
        XC_LAUNCHED(ctx);
        // fallback over the whole grid: every block exits unless the candidate kernel asked for it
        fw_alpha_eval_kernel<AT_FULL><<<grid_full, kThreads, 0, st>>>(*p, C, Ci, m, alphas_dev, n_alphas, cand_q, ctl,
                                                                      vals, 1);
        XC_LAUNCHED(ctx);
        fw_alpha_pick_kernel<<<1, 1024, 0, st>>>(vals, alphas_dev, n_alphas, cand_q, ctl, result_dev);
        XC_LAUNCHED(ctx);
    } else {
        fw_alpha_eval_kernel<AT_FULL><<<grid_full, kThreads, 0, st>>>(*p, C, Ci, m, alphas_dev, n_alphas, nullptr,
                                                                      nullptr, vals, 1);
        XC_LAUNCHED(ctx);
        fw_alpha_pick_kernel<<<1, 1024, 0, st>>>(vals, alphas_dev, n_alphas, nullptr, nullptr, result_dev);
        XC_LAUNCHED(ctx);
    }
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

// ---- one Frank-Wolfe iteration as two host calls (the all-reduce of the iterate's raw sums, if
// any, happens between them) -----------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(kThreads) f32_to_f64_kernel(const float *a, double *o, int64_t m)
{
    int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (j < m) o[j] = (double)a[j];
}
}  // namespace

extern "C" int xc_fw_step_begin(xc_ctx *ctx, const xc_metric_params *p, int have_grad, const void *eta, int dtype,
                                int64_t n, int64_t m, int64_t ld, const void *y_true, int64_t ld_true,
                                const double *Cm, float *a_row, float *b_row, double *ab64, int k, double *raw,
                                double *scal, void *stream)
{
    if (!ctx || !p || !eta || !y_true || !a_row || !b_row || !raw || !scal) return XC_ERR_INVALID;
    int rc;
    if (have_grad) {  // value of the running confusion vectors + next classifier -> (a_row, b_row)
        if (!Cm) return XC_ERR_INVALID;
        rc = xc_fw_metric_grad(ctx, p, Cm, m, a_row, b_row, scal + 0, stream);
        if (rc) return rc;
    }
    const void *a = a_row, *b = b_row;
    if (dtype == XC_F64) {  // numpy promotes the float32 classifier rows to float64 gains
        if (!ab64) return XC_ERR_INVALID;
        unsigned g = (unsigned)((m + kThreads - 1) / kThreads);
        f32_to_f64_kernel<<<g, kThreads, 0, (cudaStream_t)stream>>>(a_row, ab64, m);
        XC_LAUNCHED(ctx);
        f32_to_f64_kernel<<<g, kThreads, 0, (cudaStream_t)stream>>>(b_row, ab64 + m, m);
        XC_LAUNCHED(ctx);
        a = ab64;
        b = ab64 + m;
    }
    return xc_fw_iterate_dense(ctx, eta, dtype, n, m, ld, y_true, ld_true, a, b, k, raw, raw + m, nullptr, stream);
}

extern "C" int xc_fw_step_finish(xc_ctx *ctx, const xc_metric_params *p, int first, const double *raw,
                                 const double *colsum, int64_t m, double n_global, int normalize, int skip_tn,
                                 double *Cm, double *Ci, const double *alphas_dev, int64_t n_alphas,
                                 double fixed_alpha, double *scratch_dev, double *scal, void *stream)
{
    if (!ctx || !p || !raw || !colsum || !Cm || !Ci || !scal) return XC_ERR_INVALID;
    int rc;
    if (first) {  // classifier 0: its confusion vectors ARE the running ones (frank_wolfe.py:564-572)
        rc = xc_fw_make_conf(ctx, raw, raw + m, colsum, m, n_global, normalize, skip_tn, Cm, stream);
        if (rc) return rc;
        return xc_fw_metric_grad(ctx, p, Cm, m, nullptr, nullptr, scal + 0, stream);
    }
    rc = xc_fw_make_conf(ctx, raw, raw + m, colsum, m, n_global, normalize, skip_tn, Ci, stream);
    if (rc) return rc;
    rc = xc_fw_metric_grad(ctx, p, Ci, m, nullptr, nullptr, scal + 1, stream);  // utility of classifier i
    if (rc) return rc;
    if (alphas_dev) {
        rc = xc_fw_alpha_search(ctx, p, Cm, Ci, m, alphas_dev, n_alphas, scratch_dev, scal + 2, stream);
        if (rc) return rc;
    } else {
        XC_CUDA_TRY(ctx, cudaMemcpyAsync(scal + 2, &fixed_alpha, sizeof(double), cudaMemcpyHostToDevice,
                                         (cudaStream_t)stream));
    }
    rc = xc_fw_combine(ctx, Cm, Ci, 4 * m, scal + 2, stream);
    if (rc) return rc;
    return xc_fw_metric_grad(ctx, p, Cm, m, nullptr, nullptr, scal + 4, stream);  // new utility
}
