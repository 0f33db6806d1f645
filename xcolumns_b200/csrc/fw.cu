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
template <typename TE, int R, class Xf>
__global__ void __launch_bounds__(kThreads, R == 1 ? 6 : 1)
fw_iterate_dense_kernel(const TE *__restrict__ eta, int64_t n, int64_t m, int64_t ld, const TE *__restrict__ y_true,
                        int64_t ld_true, Xf xf, int k, double *tp, double *cnt, int32_t *__restrict__ pred_idx,
                        bool vec_ok)
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
        xc_scan_rows<TE, TE, R, false, Xf>(rp, m, vec_ok, xf, tk, dummy, k);
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

// No budget (k = 0): every label with eta * a + b >= 0 is predicted (weighted_prediction.py:55-56 with the
// th = 0 the Frank-Wolfe driver passes, frank_wolfe.py:601).  Column-oriented: 128 consecutive labels per
// CTA, two row lanes, a strip of rows; eta AND y_true are streamed (2 n m sizeof bytes per iterate).
constexpr int K0_COLS = 128, K0_STRIP = 512;

template <typename TE>
__global__ void __launch_bounds__(kThreads)
fw_iterate_dense_k0_kernel(const TE *__restrict__ eta, int64_t n, int64_t m, int64_t ld,
                           const TE *__restrict__ y_true, int64_t ld_true, const TE *__restrict__ a,
                           const TE *__restrict__ b, double *tp, double *cnt)
{
    const int64_t j = (int64_t)blockIdx.x * K0_COLS + (threadIdx.x % K0_COLS);
    if (j >= m) return;
    const int lanes = kThreads / K0_COLS;
    const int64_t r0 = (int64_t)blockIdx.y * K0_STRIP, r1 = min(n, r0 + K0_STRIP);
    const TE aj = a ? a[j] : (TE)1, bj = b ? b[j] : (TE)0;
    double t = 0.0, c = 0.0;
    for (int64_t i = r0 + threadIdx.x / K0_COLS; i < r1; i += lanes) {
        TE g = ld_stream(eta + i * ld + j);
        if (a) g = XfMulAdd<TE>::mul_rn(g, aj);
        if (b) g = XfMulAdd<TE>::add_rn(g, bj);
        if (g >= (TE)0) {
            t += (double)ld_stream(y_true + i * ld_true + j);
            c += 1.0;
        }
    }
    if (c != 0.0) {
        atomicAdd(tp + j, t);
        atomicAdd(cnt + j, c);
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

// No budget on CSR rows (k = 0): the STORED labels of a row whose gain data * a[j] + b[j] is >= 0 are predicted
// (numba_csr_functions.py:516-517 through :631-653 with the th = 0 of frank_wolfe.py:601); labels a row does not
// store are never predicted, whatever b says.  One warp per row, lanes over the stored entries.
template <typename T>
__global__ void __launch_bounds__(kThreads)
fw_iterate_csr_k0_kernel(const T *__restrict__ data, const int32_t *__restrict__ indices,
                         const int64_t *__restrict__ indptr, int64_t n, const T *__restrict__ t_data,
                         const int32_t *__restrict__ t_idx, const int64_t *__restrict__ t_ptr, XfMulAdd<T> xf,
                         double *tp, double *cnt)
{
    const int lane = lane_id();
    const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t i = warp; i < n; i += nwarps) {
        const int64_t s = indptr[i], e = indptr[i + 1];
        const int64_t ts = t_ptr[i], te = t_ptr[i + 1];
        for (int64_t q = s + lane; q < e; q += 32) {
            const int j = indices[q];
            const T g = xf.template apply_one<T>(j, data[q]);
            if (g >= (T)0) {
                int64_t y = csr_find(t_idx, ts, te, j);
                if (y >= 0) atomicAdd(tp + j, (double)t_data[y]);
                atomicAdd(cnt + j, 1.0);
            }
        }
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

// metric_grad of the call's objective, scaled so that value = (1/m) sum_j v_j and grad = (1/m) g_j also
// hold for the mixed objectives sum_j [(1 - alpha) tp_j / k + alpha metric_j / m] (frank_wolfe.py:838-915)
__device__ __forceinline__ Grad4 metric_grad_p(const xc_metric_params &p, double tp, double fp, double fn, double tn)
{
    Grad4 g = metric_grad(p.metric, tp, fp, fn, tn, p.c1, p.beta2, p.eps);
    if (p.mix == 1) {
        const double w = (1.0 - p.mix_alpha) * (p.mix_m / p.mix_k);
        g.v = p.mix_alpha * g.v + w * tp;
        g.gtp = p.mix_alpha * g.gtp + w;
        g.gfp = p.mix_alpha * g.gfp;
        g.gfn = p.mix_alpha * g.gfn;
        g.gtn = p.mix_alpha * g.gtn;
    } else if (p.mix == 3) {
        // sum_j [(1 - alpha) recall_j + alpha precision_j] (frank_wolfe.py:917-938); scaled by m because the
        // callers report (1/m) sum_j v_j and (1/m) g_j
        const Grad4 r = metric_grad(XC_METRIC_RECALL, tp, fp, fn, tn, p.c1, p.beta2, p.eps);
        const double wa = p.mix_alpha * p.mix_m, wr = (1.0 - p.mix_alpha) * p.mix_m;
        g.v = wa * g.v + wr * r.v;
        g.gtp = wa * g.gtp + wr * r.gtp;
        g.gfp = wa * g.gfp + wr * r.gfp;
        g.gfn = wa * g.gfn + wr * r.gfn;
        g.gtn = wa * g.gtn + wr * r.gtn;
    }
    return g;
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
        Grad4 g = metric_grad_p(p, tp[j], fp[j], fn[j], tn ? tn[j] : -1.0);
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
// in the confusion entries (precision, recall, F-beta, Jaccard) the search runs in three stages
// (screen, refinement -- described at fw_alpha_refine_kernel --, float64):
//   1. every grid point in float32 from a per-label linearisation T(a)/D(a) = (T0 + a dT) / (D0 + a dD).
//      The reciprocal unit (MUFU, 16 lanes/clk/SM) is the scarce pipe, so two labels share one
//      reciprocal: T1/D1 + T2/D2 = (T1 D2 + T2 D1) / (D1 D2), factors kept in product form so that no
//      precision is lost to cancelling polynomial coefficients (9 FP32 + 1 MUFU per label pair;
//      measured: 110 us for the one-label-per-reciprocal version at m = 31 k, 10^4 points);
//   2. the grid points within 1e-5 (relative) of the float32 maximum -- the float32 pass is accurate to
//      ~3e-6 in the worst case, so this is a superset of every point that can be the float64 maximum --
//      are pruned by the float32 difference refinement and the survivors re-evaluated with the
//      reference's float64 expression; the first strict maximum among them wins.  The candidates are spread over all SMs (label slices chosen on the device from the
//      candidate count); partial sums are combined in a fixed order, so the result is reproducible.
// Up to ALPHA_MAX_CAND (>= the default 10^4-point grid) screen candidates go through the refinement: on
// many rows the objective gets so flat in alpha that ALL grid points are within 1e-5 of the maximum (seen
// at 112 k rows from the 8th iteration on) -- the refinement then costs about as much as the screen
// (~80 us) where the float64 evaluation of the whole grid costs 830 us.  Only finer grids than that fall
// back to the float64 grid.  Metrics that use tn always take that path.
constexpr int AT64 = 4;             // grid points per block tile, float64 kernel
constexpr int AT32 = 32;            // grid points per block, stage-1 kernel
constexpr int ALPHA_LS = 8;         // label slices of the stage-1 kernel (blockIdx.y)
constexpr int ALPHA_LS2_MAX = 32;   // max label slices of the float64 kernel
constexpr int ALPHA_MAX_CAND = 16384;  // >= the default grid (10^4 points): a flat objective keeps them all
constexpr float ALPHA_WINDOW = 1e-5f;
constexpr int ATR = 16;             // candidate slots per block, refinement kernel
constexpr int ALPHA_LSR = 8;        // label slices of the refinement kernel (blockIdx.y)
constexpr float ALPHA_REFINE_ULP = 2.5e-7f;  // error unit of the refinement's bound (~2 ulp of float32)

struct AlphaCtl {
    int count;   // number of slots to evaluate in float64 (candidates, or n_alphas + 1 when full)
    int full;    // 1: slots are the grid points themselves
    int slices;  // label slices per slot tile
    int tiles;   // ceil(count / AT64)
    int qstar;   // stage-1 arg-max (grid point index)
    int pad[3];
};

__device__ __forceinline__ float4 alpha_lin(const xc_metric_params &p, double tp, double fp, double fn, double tpi,
                                            double fpi, double fni, float *E)
{
    double D0, D1;
    if (p.metric == XC_METRIC_PRECISION) { D0 = tp + fp + p.eps; D1 = tpi + fpi + p.eps; }
    else if (p.metric == XC_METRIC_RECALL) { D0 = tp + fn + p.eps; D1 = tpi + fni + p.eps; }
    else if (p.metric == XC_METRIC_JACCARD) { D0 = tp + fp + fn + p.eps; D1 = tpi + fpi + fni + p.eps; }
    else { D0 = p.beta2 * (tp + fp) + tp + fn + p.eps; D1 = p.beta2 * (tpi + fpi) + tpi + fni + p.eps; }
    if (E) *E = (float)((tpi - tp) * D0 - tp * (D1 - D0));  // dT D0 - T0 dD, formed in float64
    return make_float4((float)tp, (float)(tpi - tp), (float)D0, (float)(D1 - D0));
}

__global__ void __launch_bounds__(kThreads)
fw_alpha_prep_kernel(xc_metric_params p, const double *__restrict__ C, const double *__restrict__ Ci, int64_t m,
                     float4 *__restrict__ lin, float *__restrict__ linE)
{
    int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (j == m && (m & 1)) {  // pad to a whole label pair: term 0 / 1
        lin[m] = make_float4(0.f, 0.f, 1.f, 0.f);
        linE[m] = 0.f;
    }
    if (j >= m) return;
    lin[j] = alpha_lin(p, C[j], C[m + j], C[2 * m + j], Ci[j], Ci[m + j], Ci[2 * m + j], linE + j);
}

__device__ __forceinline__ float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// part[slice][q] = sum over the slice's label pairs of T/D at grid point q (q = 0 is alpha = 0)
__global__ void __launch_bounds__(kThreads)
fw_alpha_evalfast_kernel(const float4 *__restrict__ lin, int64_t npairs, const double *__restrict__ alphas,
                         int64_t n_alphas, float *__restrict__ part)
{
    __shared__ float sm[AT32][kThreads];
    const int64_t q0 = (int64_t)blockIdx.x * AT32;
    float al[AT32], acc[AT32];
#pragma unroll
    for (int t = 0; t < AT32; ++t) {
        int64_t q = q0 + t;
        al[t] = (q == 0 || q > n_alphas) ? 0.f : (float)alphas[q - 1];
        acc[t] = 0.f;
    }
    // the linearisation comes from L2 (~1 us away under load) and one pair only feeds ~300 instructions:
    // the next pair is fetched while the current one is evaluated
    int64_t pr = (int64_t)blockIdx.y * kThreads + threadIdx.x;
    const int64_t stride = (int64_t)kThreads * ALPHA_LS;
    float4 u = make_float4(0.f, 0.f, 1.f, 0.f), w = u;
    if (pr < npairs) { u = __ldg(lin + 2 * pr); w = __ldg(lin + 2 * pr + 1); }
    for (; pr < npairs; pr += stride) {
        float4 un = u, wn = w;
        if (pr + stride < npairs) { un = __ldg(lin + 2 * (pr + stride)); wn = __ldg(lin + 2 * (pr + stride) + 1); }
#pragma unroll
        for (int t = 0; t < AT32; ++t) {
            const float Ta = fmaf(al[t], u.y, u.x), Da = fmaf(al[t], u.w, u.z);
            const float Tb = fmaf(al[t], w.y, w.x), Db = fmaf(al[t], w.w, w.z);
            const float N = fmaf(Ta, Db, Tb * Da);
            acc[t] = fmaf(N, rcp_approx(Da * Db), acc[t]);
        }
        u = un;
        w = wn;
    }
    // block reduction through shared memory in a fixed order: thread (t, s) adds the 32 values
    // sm[t][32 s .. 32 s + 32) (rotated by the lane so that the 32 lanes hit 32 different banks)
#pragma unroll
    for (int t = 0; t < AT32; ++t) sm[t][threadIdx.x] = acc[t];
    __syncthreads();
    const int t = threadIdx.x >> 3, s = threadIdx.x & 7, lane = threadIdx.x & 31;
    float v = 0.f;
    for (int i = 0; i < 32; ++i) v += sm[t][s * 32 + ((i + lane) & 31)];
    v += __shfl_xor_sync(XC_FULL, v, 1);
    v += __shfl_xor_sync(XC_FULL, v, 2);
    v += __shfl_xor_sync(XC_FULL, v, 4);
    const int64_t q = q0 + t;
    if (s == 0 && q <= n_alphas) part[(int64_t)blockIdx.y * (n_alphas + 1) + q] = v;
}

// candidates = grid points within ALPHA_WINDOW (relative) of the stage-1 maximum, in grid order
__global__ void __launch_bounds__(1024)
fw_alpha_cand_kernel(const float *__restrict__ part, int64_t n_alphas, float scale, float *__restrict__ vfast,
                     int *__restrict__ cand_q, AlphaCtl *ctl, int eval_ctas)
{
    __shared__ float s_max[32];
    __shared__ int s_arg[32];
    __shared__ int s_warp[32];
    const int64_t total = n_alphas + 1;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // contiguous slice per thread so that the compacted list stays in grid order
    const int64_t per = (total + 1023) / 1024;
    const int64_t b = threadIdx.x * per, e = min(total, b + per);
    float mx = -INFINITY;
    int arg = 0x7fffffff;
    bool nan = false;
    // pass 1, coalesced: combine the label-slice partials in a fixed order, first maximum
#pragma unroll 4
    for (int64_t q = threadIdx.x; q < total; q += 1024) {
        float v = 0.f;
#pragma unroll
        for (int sl = 0; sl < ALPHA_LS; ++sl) v += __ldg(part + (int64_t)sl * total + q);
        v *= scale;
        vfast[q] = v;
        nan |= !(v == v);
        if (v > mx) { mx = v; arg = (int)q; }  // q ascending per thread: first maximum
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(XC_FULL, mx, o);
        const int oa = __shfl_xor_sync(XC_FULL, arg, o);
        if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
    }
    if (lane == 0) { s_max[wid] = mx; s_arg[wid] = arg; }
    const int any_nan = __syncthreads_or(nan);
    mx = s_max[0];
    arg = s_arg[0];
    for (int w = 1; w < 32; ++w)
        if (s_max[w] > mx || (s_max[w] == mx && s_arg[w] < arg)) { mx = s_max[w]; arg = s_arg[w]; }
    const float thr = mx - ALPHA_WINDOW * fabsf(mx) - 1e-37f;
    int cnt = 0;   // (the barrier above also published vfast to the whole block)
    for (int64_t q = b; q < e; ++q) cnt += vfast[q] >= thr;
    // exclusive scan over the 1024 per-thread counts
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(XC_FULL, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = s_warp[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(XC_FULL, wi, o);
            if (lane >= o) wi += y;
        }
        s_warp[lane] = wi - w;  // exclusive prefix of the warp totals
        if (lane == 31) s_arg[0] = wi;
    }
    __syncthreads();
    const int run = s_arg[0];
    const bool full = run > ALPHA_MAX_CAND || any_nan || arg == 0x7fffffff;
    if (!full) {
        int o = s_warp[wid] + incl - cnt;
        for (int64_t q = b; q < e; ++q)
            if (vfast[q] >= thr) cand_q[o++] = (int)q;
    }
    if (threadIdx.x == 0) {
        const int count = full ? (int)total : run;
        const int tiles = (count + AT64 - 1) / AT64;
        int slices = full ? 1 : eval_ctas / (tiles > 0 ? tiles : 1);
        slices = slices < 1 ? 1 : (slices > ALPHA_LS2_MAX ? ALPHA_LS2_MAX : slices);
        ctl->count = count;
        ctl->full = full ? 1 : 0;
        ctl->slices = slices;
        ctl->tiles = tiles;
        ctl->qstar = full ? 0 : arg;
    }
}

// Stage 1.5: with alpha* the stage-1 arg-max, the difference of one label's term between a candidate
// and alpha* has the closed form
//     T(a)/D(a) - T(a*)/D(a*) = (a - a*) E / (D(a) D(a*)),   E = dT D0 - T0 dD  (formed in float64)
// which float32 evaluates to a relative accuracy of a few ulp OF THE DIFFERENCE (the direct float32
// values only resolve ~1e-6 of the objective itself, which on a flat objective leaves hundreds of
// candidates).  Per candidate the kernel accumulates d = sum_j t_j and a bound on its rounding error
// err = ulp * sum_j |t_j| (amp_j(a) + amp_j(a*) + 10), amp = (|D0| + a |dD|) / D(a) being the
// cancellation amplification of the float32 denominator (the constant covers the reciprocal, the
// products and the worst-case rounding of the ~36-deep float32 summation).  A candidate survives iff
// d + err >= max_s (d_s - err_s): only those can be the float64 maximum.
__global__ void __launch_bounds__(kThreads)
fw_alpha_refine_kernel(const float4 *__restrict__ lin, const float *__restrict__ linE, int64_t m,
                       const double *__restrict__ alphas, const int *__restrict__ cand_q,
                       const AlphaCtl *__restrict__ ctl, float *__restrict__ part_d, float *__restrict__ part_e)
{
    __shared__ float sm[2][ATR][kThreads / 32];
    if (ctl->full) return;
    const int count = ctl->count;
    const int qs = ctl->qstar;
    const float astar = qs == 0 ? 0.f : (float)alphas[qs - 1];
    // the candidate count is only known on the device: a fixed grid walks the slot tiles
    for (int s0 = blockIdx.x * ATR; s0 < count; s0 += gridDim.x * ATR) {
    float al[ATR], accd[ATR], acce[ATR];
#pragma unroll
    for (int t = 0; t < ATR; ++t) {
        const int slot = s0 + t;
        const int q = slot < count ? cand_q[slot] : qs;
        al[t] = q == 0 ? 0.f : (float)alphas[q - 1];
        accd[t] = 0.f;
        acce[t] = 0.f;
    }
    int64_t j = (int64_t)blockIdx.y * kThreads + threadIdx.x;
    const int64_t stride = (int64_t)kThreads * ALPHA_LSR;
    float4 l = make_float4(0.f, 0.f, 1.f, 0.f);
    float le = 0.f;
    if (j < m) { l = __ldg(lin + j); le = __ldg(linE + j); }
    for (; j < m; j += stride) {
        float4 ln = l;
        float len = le;
        if (j + stride < m) { ln = __ldg(lin + j + stride); len = __ldg(linE + j + stride); }
        const float aD0 = fabsf(l.z), adD = fabsf(l.w);
        const float rstar = rcp_approx(fmaf(astar, l.w, l.z));
        const float ampstar = fmaf(astar, adD, aD0) * rstar + 10.f;
        const float er = le * rstar;
#pragma unroll
        for (int t = 0; t < ATR; ++t) {
            const float rs = rcp_approx(fmaf(al[t], l.w, l.z));
            const float tt = er * rs;
            accd[t] += tt;
            acce[t] = fmaf(fabsf(tt), fmaf(fmaf(al[t], adD, aD0), rs, ampstar), acce[t]);
        }
        l = ln;
        le = len;
    }
#pragma unroll
    for (int t = 0; t < ATR; ++t) {
        float v = accd[t], w = acce[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v += __shfl_xor_sync(XC_FULL, v, o);
            w += __shfl_xor_sync(XC_FULL, w, o);
        }
        if ((threadIdx.x & 31) == 0) { sm[0][t][threadIdx.x >> 5] = v; sm[1][t][threadIdx.x >> 5] = w; }
    }
    __syncthreads();
    if (threadIdx.x < ATR) {
        float v = 0.f, w = 0.f;
        for (int x = 0; x < kThreads / 32; ++x) { v += sm[0][threadIdx.x][x]; w += sm[1][threadIdx.x][x]; }
        const int slot = s0 + threadIdx.x;
        if (slot < count) {
            part_d[blockIdx.y * ALPHA_MAX_CAND + slot] = v;
            part_e[blockIdx.y * ALPHA_MAX_CAND + slot] = w;
        }
    }
    __syncthreads();   // sm is reused by the next tile
    }
}

// prune the candidate list with the refined differences; writes the final list + control block.
// hi_buf: scratch of >= count floats (the stage-1 value array, no longer needed at this point).
__global__ void __launch_bounds__(1024)
fw_alpha_cand2_kernel(const float *__restrict__ part_d, const float *__restrict__ part_e,
                      const double *__restrict__ alphas, const int *__restrict__ cand_q,
                      const AlphaCtl *__restrict__ ctl, float *__restrict__ hi_buf, int *__restrict__ cand_q2,
                      AlphaCtl *ctl2, int eval_ctas)
{
    __shared__ float s_lo[32];
    __shared__ int s_warp[32];
    __shared__ int s_run;
    if (ctl->full) {
        if (threadIdx.x == 0) *ctl2 = *ctl;
        return;
    }
    const int count = ctl->count, qs = ctl->qstar;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const double astar = qs == 0 ? 0.0 : alphas[qs - 1];
    // pass 1 (coalesced): refined difference and its error bound per slot, best lower bound
    float lo = -INFINITY;
    for (int slot = threadIdx.x; slot < count; slot += 1024) {
        const int q = cand_q[slot];
        float d = 0.f, er = 0.f;
#pragma unroll
        for (int sl = 0; sl < ALPHA_LSR; ++sl) {
            d += part_d[sl * ALPHA_MAX_CAND + slot];
            er += part_e[sl * ALPHA_MAX_CAND + slot];
        }
        const float delta = (float)((q == 0 ? 0.0 : alphas[q - 1]) - astar);
        d *= delta;
        er = er * fabsf(delta) * ALPHA_REFINE_ULP;
        if (q == qs) { d = 0.f; er = 0.f; }
        if (!(d == d) || !(er == er)) { d = 0.f; er = INFINITY; }  // not a number: keep, bound nothing
        hi_buf[slot] = d + er;
        lo = fmaxf(lo, d - er);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lo = fmaxf(lo, __shfl_xor_sync(XC_FULL, lo, o));
    if (lane == 0) s_lo[wid] = lo;
    __syncthreads();   // also publishes hi_buf to the block
    lo = s_lo[0];
    for (int w = 1; w < 32; ++w) lo = fmaxf(lo, s_lo[w]);
    // pass 2: contiguous slices so that the surviving list stays in grid order
    const int per = (count + 1023) / 1024;
    const int b0 = threadIdx.x * per, b1 = min(count, b0 + per);
    int cnt = 0;
    for (int slot = b0; slot < b1; ++slot) cnt += hi_buf[slot] >= lo;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(XC_FULL, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = s_warp[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(XC_FULL, wi, o);
            if (lane >= o) wi += y;
        }
        s_warp[lane] = wi - w;
        if (lane == 31) s_run = wi;
    }
    __syncthreads();
    int o = s_warp[wid] + incl - cnt;
    for (int slot = b0; slot < b1; ++slot)
        if (hi_buf[slot] >= lo) cand_q2[o++] = cand_q[slot];
    if (threadIdx.x == 0) {
        const int run = s_run;
        const int tiles = (run + AT64 - 1) / AT64;
        int slices = eval_ctas / (tiles > 0 ? tiles : 1);
        slices = slices < 1 ? 1 : (slices > ALPHA_LS2_MAX ? ALPHA_LS2_MAX : slices);
        ctl2->count = run;
        ctl2->full = 0;
        ctl2->slices = slices;
        ctl2->tiles = tiles;
        ctl2->qstar = qs;
    }
}

// float64 evaluation with the reference's expression (frank_wolfe.py:393-398).  Slot t evaluates grid
// point q = cand_q[t] (candidate mode) or q = t (full mode / ctl == nullptr); q = 0 is alpha = 0.
// Work item = (slot tile of AT64 slots, label slice); partial sums go to vals[slice * vstride + slot].
__global__ void __launch_bounds__(kThreads)
fw_alpha_eval_kernel(xc_metric_params p, const double *__restrict__ C, const double *__restrict__ Ci, int64_t m,
                     const double *__restrict__ alphas, int64_t n_alphas, const int *__restrict__ cand_q,
                     const AlphaCtl *__restrict__ ctl, double *__restrict__ vals, int64_t vstride)
{
    __shared__ double sm[AT64][kThreads / 32];
    const bool full = ctl == nullptr || ctl->full;
    const int64_t count = ctl == nullptr ? n_alphas + 1 : ctl->count;
    const int slices = ctl == nullptr ? 1 : ctl->slices;
    const int64_t tiles = (count + AT64 - 1) / AT64;
    const bool use_tn = p.metric >= XC_METRIC_BALANCED_ACC && p.metric <= XC_METRIC_HMEAN;
    const int64_t per_slice = (m + slices - 1) / slices;
    for (int64_t item = blockIdx.x; item < tiles * slices; item += gridDim.x) {
        const int64_t tile = item / slices;
        const int slice = (int)(item - tile * slices);
        const int64_t s0 = tile * AT64;
        double al[AT64], acc[AT64];
#pragma unroll
        for (int t = 0; t < AT64; ++t) {
            int64_t slot = s0 + t;
            int64_t q = slot < count ? (full ? slot : (int64_t)cand_q[slot]) : 0;
            al[t] = (q == 0) ? 0.0 : alphas[q - 1];
            acc[t] = 0.0;
        }
        const int64_t jb = slice * per_slice, je = min(m, jb + per_slice);
        for (int64_t j = jb + threadIdx.x; j < je; j += kThreads) {
            const double tp = C[j], fp = C[m + j], fn = C[2 * m + j], tn = use_tn ? C[3 * m + j] : 0.0;
            const double tpi = Ci[j], fpi = Ci[m + j], fni = Ci[2 * m + j], tni = use_tn ? Ci[3 * m + j] : 0.0;
#pragma unroll
            for (int t = 0; t < AT64; ++t) {
                const double a1 = al[t], a0 = 1.0 - a1;
                acc[t] += xc_metric_eval(p, a0 * tp + a1 * tpi, a0 * fp + a1 * fpi, a0 * fn + a1 * fni,
                                           a0 * tn + a1 * tni);
            }
        }
#pragma unroll
        for (int t = 0; t < AT64; ++t) {
            double v = warp_sum(acc[t]);
            if ((threadIdx.x & 31) == 0) sm[t][threadIdx.x >> 5] = v;
        }
        __syncthreads();
        if (threadIdx.x < AT64) {
            double v = 0.0;
            for (int w = 0; w < kThreads / 32; ++w) v += sm[threadIdx.x][w];
            int64_t slot = s0 + threadIdx.x;
            if (slot < count) vals[(int64_t)slice * vstride + slot] = v;
        }
        __syncthreads();
    }
}

// first strict maximum over the evaluated slots (slots are in grid order; utils.py:177-184); the
// label-slice partials of a slot are added in slice order, then divided by m
__global__ void __launch_bounds__(1024)
fw_alpha_pick_kernel(const double *__restrict__ vals, int64_t vstride, int64_t m, const double *__restrict__ alphas,
                     int64_t n_alphas, const int *__restrict__ cand_q, const AlphaCtl *__restrict__ ctl,
                     double *result)
{
    __shared__ double sv[32];
    __shared__ long long sq[32];
    const bool full = ctl == nullptr || ctl->full;
    const int64_t count = ctl == nullptr ? n_alphas + 1 : ctl->count;
    const int slices = ctl == nullptr ? 1 : ctl->slices;
    double bv = -INFINITY;
    long long bq = 0x7fffffffffffffffLL;
    // one warp per slot: lane l fetches slice l's partial, fixed shuffle tree (reproducible)
    for (int64_t q = threadIdx.x >> 5; q < count; q += 32) {
        const int l = threadIdx.x & 31;
        double v = l < slices ? vals[(int64_t)l * vstride + q] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(XC_FULL, v, o);
        v = v / (double)m;
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
            if (bq == 0x7fffffffffffffffLL) {
                bq = 0;
                bv = NAN;
            }
            const long long q = full ? bq : (long long)cand_q[bq];
            result[0] = q == 0 ? 0.0 : alphas[q - 1];
            result[1] = bv;
        }
    }
}

// ---- ternary search (utils.py:187-201 as called at frank_wolfe.py:399-400) ------------------------------
// One CTA walks the reference's loop: both probes of a round are reduced over the labels in a fixed
// order (thread-strided partial sums, shuffle tree, warp order), so the sequence of decisions -- and
// with it the returned step -- is reproducible.  ~18 rounds for eps = 1e-3.
__global__ void __launch_bounds__(1024)
fw_alpha_ternary_kernel(xc_metric_params p, const double *__restrict__ C, const double *__restrict__ Ci, int64_t m,
                        double eps, double *result)
{
    __shared__ double sm[2][32];
    const bool use_tn = p.metric >= XC_METRIC_BALANCED_ACC && p.metric <= XC_METRIC_HMEAN;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    auto eval2 = [&](double a1, double a2, double &v1, double &v2) {
        double s1 = 0.0, s2 = 0.0;
        for (int64_t j = threadIdx.x; j < m; j += 1024) {
            const double tp = C[j], fp = C[m + j], fn = C[2 * m + j], tn = use_tn ? C[3 * m + j] : 0.0;
            const double tpi = Ci[j], fpi = Ci[m + j], fni = Ci[2 * m + j], tni = use_tn ? Ci[3 * m + j] : 0.0;
            const double b1 = 1.0 - a1, b2 = 1.0 - a2;
            s1 += xc_metric_eval(p, b1 * tp + a1 * tpi, b1 * fp + a1 * fpi, b1 * fn + a1 * fni, b1 * tn + a1 * tni);
            s2 += xc_metric_eval(p, b2 * tp + a2 * tpi, b2 * fp + a2 * fpi, b2 * fn + a2 * fni, b2 * tn + a2 * tni);
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        __syncthreads();  // previous round's readers are done with sm
        if (lane == 0) { sm[0][wid] = s1; sm[1][wid] = s2; }
        __syncthreads();
        v1 = 0.0;
        v2 = 0.0;
        for (int w = 0; w < 32; ++w) { v1 += sm[0][w]; v2 += sm[1][w]; }
        const double div = (p.mix == 1 || p.mix == 3) ? 1.0 : (double)m;
        v1 = v1 / div;
        v2 = v2 / div;
    };
    double low = 0.0, high = 1.0;
    while (high - low > eps) {
        const double mid1 = low + (high - low) / 3.0;
        const double mid2 = high - (high - low) / 3.0;
        double f1, f2;
        eval2(mid1, mid2, f1, f2);
        if (f1 < f2) high = mid2;  // (sic) utils.py:194-197
        else low = mid1;
    }
    const double best = (low + high) / 2.0;
    double fb, dummy;
    eval2(best, best, fb, dummy);
    if (threadIdx.x == 0) {
        result[0] = best;
        result[1] = fb;
    }
}

__global__ void __launch_bounds__(kThreads)
fw_combine_kernel(double *C, const double *Ci, int64_t m4, const double *alpha_dev)
{
    const double a1 = *alpha_dev;
    int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (j < m4) C[j] = (1.0 - a1) * C[j] + a1 * Ci[j];  // frank_wolfe.py:633-636
}

// ---- fused per-label stages of one dense iteration ---------------------------------------------------
// (1) confusion vectors of the newest classifier from the raw sums (fw_make_conf_kernel), its
//     utility (deterministic grid sum) and, for the two-stage search, the per-label linearisation.
__global__ void __launch_bounds__(256)
fw_conf_prep_kernel(xc_metric_params p, const double *__restrict__ tp_raw, const double *__restrict__ cnt,
                    const double *__restrict__ colsum, int64_t m, double n, int normalize, int skip_tn,
                    const double *__restrict__ Cm, double *__restrict__ out, float4 *__restrict__ lin,
                    float *__restrict__ linE, double *value, double *partials, unsigned *counter)
{
    __shared__ double sm[8];
    double s = 0.0;
    for (int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x; j < m; j += (int64_t)gridDim.x * 256) {
        double t = tp_raw[j], f = cnt[j] - t, g = colsum[j] - t;
        if (normalize) { t = t / n; f = f / n; g = g / n; }
        const double tn = skip_tn ? -1.0 : ((-t - f) - g) + (normalize ? 1.0 : n);
        out[j] = t;
        out[m + j] = f;
        out[2 * m + j] = g;
        out[3 * m + j] = tn;
        s += metric_grad_p(p, t, f, g, tn).v;
        if (lin) lin[j] = alpha_lin(p, Cm[j], Cm[m + j], Cm[2 * m + j], t, f, g, linE + j);
    }
    if (lin && (m & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        lin[m] = make_float4(0.f, 0.f, 1.f, 0.f);
        linE[m] = 0.f;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    double bsum = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < 8; ++w) bsum += sm[w];
    double total;
    if (xc_grid_sum_last(bsum, partials, counter, &total) && value) *value = total * (1.0 / (double)m);
}

// (2) C = (1 - alpha) C + alpha Ci (skipped when alpha_dev == nullptr), utility of the new C, the next
//     classifier from its gradient (frank_wolfe.py:591-596) and re-zeroing of the raw sums
__global__ void __launch_bounds__(256)
fw_finish_kernel(xc_metric_params p, double *__restrict__ C, const double *__restrict__ Ci, int64_t m,
                 const double *__restrict__ alpha_dev, float *__restrict__ a_next, float *__restrict__ b_next,
                 double *__restrict__ raw_zero, double *value, double *value_next, double *partials,
                 unsigned *counter)
{
    __shared__ double sm[8];
    double s = 0.0;
    const double sgn = p.maximize ? 1.0 : -1.0;
    const double inv_m = 1.0 / (double)m;
    const double a1 = alpha_dev ? *alpha_dev : 0.0;
    for (int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x; j < m; j += (int64_t)gridDim.x * 256) {
        double c[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            c[t] = C[t * m + j];
            if (alpha_dev) {
                c[t] = (1.0 - a1) * c[t] + a1 * Ci[t * m + j];
                C[t * m + j] = c[t];
            }
        }
        Grad4 g = metric_grad_p(p, c[0], c[1], c[2], c[3]);
        s += g.v;
        if (a_next) {
            double gtp = g.gtp * inv_m, gfp = g.gfp * inv_m, gfn = g.gfn * inv_m, gtn = g.gtn * inv_m;
            a_next[j] = (float)(sgn * (((gtp - gfp) - gfn) + gtn));
            b_next[j] = (float)(sgn * (gfp - gtn));
        }
        if (raw_zero) { raw_zero[j] = 0.0; raw_zero[m + j] = 0.0; }
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    double bsum = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < 8; ++w) bsum += sm[w];
    double total;
    if (xc_grid_sum_last(bsum, partials, counter, &total)) {
        if (value) *value = total * inv_m;
        if (value_next) *value_next = total * inv_m;
    }
}

// ---- micro-averaged objectives (metrics.py:68-100: binary_metric(tp.sum(), fp.sum(), fn.sum(), tn.sum())) ----
// The objective only sees the four sums, its gradient is the same for every label and the line search is a
// scalar problem: one CTA does all of it.  S = sums of the running matrix, Si = sums of the newest
// classifier's matrix (label-strided partial sums, shuffle tree, warp order: reproducible).
//   scal[1] = metric(Si); scal[2..3] = step (grid / ternary / fixed) and its value; scal[4] = metric of the
//   combined sums; scal[5..6] = the classifier (a, b) every label gets next (frank_wolfe.py:591-596).
__global__ void __launch_bounds__(1024)
fw_micro_kernel(xc_metric_params p, const double *__restrict__ C, const double *__restrict__ Ci, int64_t m, int first,
                const double *__restrict__ alphas, int64_t n_alphas, double fixed_alpha, double ternary_eps,
                double *scal, double *scal_next)
{
    __shared__ double sm[8][32];
    __shared__ double s_sum[8];
    __shared__ double s_bv[32];
    __shared__ long long s_bq[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double part[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) part[t] = 0.0;
    for (int64_t j = threadIdx.x; j < m; j += 1024) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            part[t] += C[t * m + j];
            if (!first) part[4 + t] += Ci[t * m + j];
        }
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const double v = warp_sum(part[t]);
        if (lane == 0) sm[t][wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        double v = 0.0;
        for (int w = 0; w < 32; ++w) v += sm[threadIdx.x][w];
        s_sum[threadIdx.x] = v;
    }
    __syncthreads();
    const double S[4] = {s_sum[0], s_sum[1], s_sum[2], s_sum[3]};
    const double Si[4] = {s_sum[4], s_sum[5], s_sum[6], s_sum[7]};
    auto f_at = [&](double a1) {
        const double a0 = 1.0 - a1;
        return xc_binary_metric(p.metric, a0 * S[0] + a1 * Si[0], a0 * S[1] + a1 * Si[1], a0 * S[2] + a1 * Si[2],
                                a0 * S[3] + a1 * Si[3], p.c1, p.beta2, p.eps);
    };
    double alpha = 0.0, best = 0.0;
    if (!first) {
        if (ternary_eps > 0.0) {  // utils.py:187-201 verbatim
            double low = 0.0, high = 1.0;
            while (high - low > ternary_eps) {
                const double mid1 = low + (high - low) / 3.0, mid2 = high - (high - low) / 3.0;
                if (f_at(mid1) < f_at(mid2)) high = mid2;
                else low = mid1;
            }
            alpha = (low + high) / 2.0;
            best = f_at(alpha);
        } else if (alphas) {      // utils.py:174-184: first strict maximum over {0} + grid
            double bv = -INFINITY;
            long long bq = 0x7fffffffffffffffLL;
            for (int64_t q = threadIdx.x; q <= n_alphas; q += 1024) {
                const double v = f_at(q == 0 ? 0.0 : alphas[q - 1]);
                if (v > bv || (v == bv && q < bq)) { bv = v; bq = q; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(XC_FULL, bv, o);
                const long long oq = __shfl_xor_sync(XC_FULL, bq, o);
                if (ov > bv || (ov == bv && oq < bq)) { bv = ov; bq = oq; }
            }
            if (lane == 0) { s_bv[wid] = bv; s_bq[wid] = bq; }
            __syncthreads();
            bv = s_bv[0];
            bq = s_bq[0];
            for (int w = 1; w < 32; ++w)
                if (s_bv[w] > bv || (s_bv[w] == bv && s_bq[w] < bq)) { bv = s_bv[w]; bq = s_bq[w]; }
            if (bq == 0x7fffffffffffffffLL) bq = 0;  // all NaN: keep alpha = 0 like the reference
            alpha = bq == 0 ? 0.0 : alphas[bq - 1];
            best = bv;
        } else {
            alpha = fixed_alpha;
            best = f_at(alpha);
        }
    }
    if (threadIdx.x == 0) {
        double T[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) T[t] = first ? S[t] : (1.0 - alpha) * S[t] + alpha * Si[t];
        const Grad4 g = metric_grad(p.metric, T[0], T[1], T[2], T[3], p.c1, p.beta2, p.eps);
        const double sgn = p.maximize ? 1.0 : -1.0;
        if (first) {
            scal[0] = g.v;
        } else {
            scal[1] = xc_binary_metric(p.metric, Si[0], Si[1], Si[2], Si[3], p.c1, p.beta2, p.eps);
            scal[2] = alpha;
            scal[3] = best;
            scal[4] = g.v;
        }
        scal[5] = sgn * (((g.gtp - g.gfp) - g.gfn) + g.gtn);
        scal[6] = sgn * (g.gfp - g.gtn);
        if (scal_next) scal_next[0] = g.v;
    }
}

__global__ void __launch_bounds__(kThreads)
fw_micro_fill_kernel(const double *__restrict__ ab, int64_t m, float *__restrict__ a_next, float *__restrict__ b_next,
                     double *__restrict__ raw_zero)
{
    const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (j >= m) return;
    if (a_next) {
        a_next[j] = (float)ab[0];
        b_next[j] = (float)ab[1];
    }
    if (raw_zero) { raw_zero[j] = 0.0; raw_zero[m + j] = 0.0; }
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
    if (k == 0) {
        if (pred_idx) return XC_ERR_INVALID;  // no compact prediction without a budget
        dim3 grid((unsigned)((m + K0_COLS - 1) / K0_COLS), (unsigned)((n + K0_STRIP - 1) / K0_STRIP));
        fw_iterate_dense_k0_kernel<TE><<<grid, kThreads, 0, st>>>((const TE *)eta, n, m, ld, (const TE *)y_true, ld_true,
                                                                  (const TE *)a, (const TE *)b, tp, cnt);
        XC_LAUNCHED(ctx);
        return XC_OK;
    }
    constexpr int V = 16 / sizeof(TE);
    bool vec_ok = xc_aligned16(eta) && (ld % V == 0);
    XfMulAdd<TE> xf{(const TE *)a, (const TE *)b};
    const int64_t coef_bytes = 2 * m * (int64_t)sizeof(TE);
    int rr = coef_bytes <= 160 * 1024 ? 1 : (coef_bytes <= 512 * 1024 ? 2 : 4);
    // with 128-bit coefficient loads one row per warp wins even when the classifier spills out of L1
    // (measured at 14 k x 31 k, 248 KB of coefficients: R = 1 286 us, R = 2 311 us)
    if (vec_ok && a && b && xc_aligned16(a) && xc_aligned16(b) && coef_bytes <= 1024 * 1024) rr = 1;
    if (const char *e = getenv("XCOLUMNS_B200_DENSE_R")) {
        int v = atoi(e);
        if (v == 1 || v == 2 || v == 4) rr = v;
    }
    // classifier rows present and 16-byte aligned (the host shim pads the classifier matrices): vector loads
    const bool coef_vec = vec_ok && a && b && xc_aligned16(a) && xc_aligned16(b);
    XfMulAddVec<TE> xfv{(const TE *)a, (const TE *)b};
#define XC_GO(R)                                                                                             \
    if (coef_vec) {                                                                                          \
        auto kern = fw_iterate_dense_kernel<TE, R, XfMulAddVec<TE>>;                                      \
        int grid = grid_for(ctx, kern, (n + R - 1) / R);                                                     \
        kern<<<grid, kThreads, 0, st>>>((const TE *)eta, n, m, ld, (const TE *)y_true, ld_true, xfv, k, tp,  \
                                        cnt, pred_idx, vec_ok);                                              \
    } else {                                                                                                 \
        auto kern = fw_iterate_dense_kernel<TE, R, XfMulAdd<TE>>;                                         \
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
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !eta || !y_true || !tp || !cnt || n <= 0 || m <= 0 || ld < m || ld_true < m) return XC_ERR_INVALID;
    if (k < 0 || k > 32 || k > m) return XC_ERR_INVALID;   // k == 0: no budget, threshold at 0
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
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !indptr || !t_ptr || !tp || !cnt || n <= 0 || m <= 0) return XC_ERR_INVALID;
    if (k < 0 || k > 32 || (k == 0 && pred_idx)) return XC_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    XC_CUDA_TRY(ctx, cudaMemsetAsync(tp, 0, sizeof(double) * m, st));
    XC_CUDA_TRY(ctx, cudaMemsetAsync(cnt, 0, sizeof(double) * m, st));
    if (k == 0 && dtype == XC_F32) {
        auto kern = fw_iterate_csr_k0_kernel<float>;
        int grid = grid_for(ctx, kern, n);
        XfMulAdd<float> xf{(const float *)a, (const float *)b};
        kern<<<grid, kThreads, 0, st>>>((const float *)data, indices, indptr, n, (const float *)t_data, t_idx, t_ptr,
                                        xf, tp, cnt);
    } else if (k == 0 && dtype == XC_F64) {
        auto kern = fw_iterate_csr_k0_kernel<double>;
        int grid = grid_for(ctx, kern, n);
        XfMulAdd<double> xf{(const double *)a, (const double *)b};
        kern<<<grid, kThreads, 0, st>>>((const double *)data, indices, indptr, n, (const double *)t_data, t_idx, t_ptr,
                                        xf, tp, cnt);
    } else if (dtype == XC_F32) {
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
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !tp_raw || !cnt || !colsum || !Ci || m <= 0) return XC_ERR_INVALID;
    fw_make_conf_kernel<<<(unsigned)((m + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        tp_raw, cnt, colsum, m, n, normalize, skip_tn, Ci);
    XC_LAUNCHED(ctx);
    return XC_OK;
}

extern "C" int xc_fw_metric_grad(xc_ctx *ctx, const xc_metric_params *p, const double *C, int64_t m, float *a_out,
                                 float *b_out, double *value_dev, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
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

namespace {
// scratch layout of the line search (all offsets 32-byte aligned)
struct AlphaScratch {
    double *vals;      // float64 partial sums: [slice][vstride]
    int64_t vstride;
    float *part;       // stage-1 partial sums: [ALPHA_LS][n_alphas + 1]
    float *vfast;      // stage-1 values [n_alphas + 1]
    int *cand_q;       // [ALPHA_MAX_CAND] stage-1 candidates
    int *cand_q2;      // [ALPHA_MAX_CAND] candidates that survive the refinement
    float *part_d;     // refinement partial sums [ALPHA_LSR][ALPHA_MAX_CAND]
    float *part_e;
    AlphaCtl *ctl;     // final control block (diagnostics: xc_fw_alpha_ctl_offset)
    AlphaCtl *ctl1;    // stage-1 control block
    float4 *lin;       // [m + 1]
    float *linE;       // [m + 1]
    size_t bytes;
};

AlphaScratch alpha_scratch(void *base_, int64_t m, int64_t n_alphas)
{
    auto up = [](size_t v) { return (v + 31) & ~(size_t)31; };
    uint8_t *base = reinterpret_cast<uint8_t *>(base_);
    AlphaScratch s;
    const int64_t total = n_alphas + 1;
    s.vstride = ALPHA_MAX_CAND;
    const int64_t nvals = total > (int64_t)ALPHA_MAX_CAND * ALPHA_LS2_MAX ? total : (int64_t)ALPHA_MAX_CAND * ALPHA_LS2_MAX;
    size_t off = 0;
    s.vals = reinterpret_cast<double *>(base + off);
    off = up(off + (size_t)nvals * 8);
    s.part = reinterpret_cast<float *>(base + off);
    off = up(off + (size_t)total * 4 * ALPHA_LS);
    s.vfast = reinterpret_cast<float *>(base + off);
    off = up(off + (size_t)total * 4);
    s.cand_q = reinterpret_cast<int *>(base + off);
    off = up(off + (size_t)ALPHA_MAX_CAND * 4);
    s.cand_q2 = reinterpret_cast<int *>(base + off);
    off = up(off + (size_t)ALPHA_MAX_CAND * 4);
    s.part_d = reinterpret_cast<float *>(base + off);
    off = up(off + (size_t)ALPHA_MAX_CAND * 4 * ALPHA_LSR);
    s.part_e = reinterpret_cast<float *>(base + off);
    off = up(off + (size_t)ALPHA_MAX_CAND * 4 * ALPHA_LSR);
    s.ctl = reinterpret_cast<AlphaCtl *>(base + off);
    off = up(off + 64);
    s.ctl1 = reinterpret_cast<AlphaCtl *>(base + off);
    off = up(off + 64);
    s.lin = reinterpret_cast<float4 *>(base + off);
    off = up(off + (size_t)(m + 1) * 16);
    s.linE = reinterpret_cast<float *>(base + off);
    off = up(off + (size_t)(m + 1) * 4);
    s.bytes = off + 256;
    return s;
}

int eval_grid(xc_ctx *ctx)
{
    static int per_sm = 0;
    if (!per_sm) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fw_alpha_eval_kernel, kThreads, 0);
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 4) per_sm = 4;
    }
    return ctx->sm_count * per_sm;
}

bool alpha_two_stage(const xc_metric_params *p, int64_t n_alphas)
{
    static int force_full = -1;  // $XCOLUMNS_B200_FW_SEARCH=full: float64 over the whole grid
    if (force_full < 0) {
        const char *e = getenv("XCOLUMNS_B200_FW_SEARCH");
        force_full = (e && e[0] == 'f') ? 1 : 0;
    }
    return !force_full && !p->mix && p->metric <= XC_METRIC_JACCARD && n_alphas >= 256;
}

bool alpha_refine()
{
    static int off = -1;  // $XCOLUMNS_B200_FW_REFINE=0: skip the stage-1.5 pruning
    if (off < 0) {
        const char *e = getenv("XCOLUMNS_B200_FW_REFINE");
        off = (e && e[0] == '0') ? 1 : 0;
    }
    return !off;
}

// line search proper; `lin` must already hold the linearisation when the two-stage path applies
int alpha_search_launch(xc_ctx *ctx, const xc_metric_params *p, const double *C, const double *Ci, int64_t m,
                        const double *alphas_dev, int64_t n_alphas, const AlphaScratch &s, double *result_dev,
                        cudaStream_t st)
{
    const int grid64 = eval_grid(ctx);
    if (alpha_two_stage(p, n_alphas)) {
        const float scale = (float)((p->metric == XC_METRIC_FBETA ? p->c1 : 1.0) / (double)m);
        dim3 g32((unsigned)((n_alphas + 1 + AT32 - 1) / AT32), ALPHA_LS);
        fw_alpha_evalfast_kernel<<<g32, kThreads, 0, st>>>(s.lin, (m + 1) / 2, alphas_dev, n_alphas, s.part);
        XC_LAUNCHED(ctx);
        fw_alpha_cand_kernel<<<1, 1024, 0, st>>>(s.part, n_alphas, scale, s.vfast, s.cand_q, s.ctl1, grid64);
        XC_LAUNCHED(ctx);
        const int *cq = s.cand_q;
        const AlphaCtl *cc = s.ctl1;
        if (alpha_refine()) {
            fw_alpha_refine_kernel<<<dim3(2 * ctx->sm_count, ALPHA_LSR), kThreads, 0, st>>>(
                s.lin, s.linE, m, alphas_dev, s.cand_q, s.ctl1, s.part_d, s.part_e);
            XC_LAUNCHED(ctx);
            fw_alpha_cand2_kernel<<<1, 1024, 0, st>>>(s.part_d, s.part_e, alphas_dev, s.cand_q, s.ctl1, s.vfast, s.cand_q2,
                                                      s.ctl, grid64);
            XC_LAUNCHED(ctx);
            cq = s.cand_q2;
            cc = s.ctl;
        }
        fw_alpha_eval_kernel<<<grid64, kThreads, 0, st>>>(*p, C, Ci, m, alphas_dev, n_alphas, cq, cc, s.vals, s.vstride);
        XC_LAUNCHED(ctx);
        fw_alpha_pick_kernel<<<1, 1024, 0, st>>>(s.vals, s.vstride, m, alphas_dev, n_alphas, cq, cc, result_dev);
        XC_LAUNCHED(ctx);
    } else {
        fw_alpha_eval_kernel<<<grid64, kThreads, 0, st>>>(*p, C, Ci, m, alphas_dev, n_alphas, nullptr, nullptr, s.vals,
                                                          s.vstride);
        XC_LAUNCHED(ctx);
        fw_alpha_pick_kernel<<<1, 1024, 0, st>>>(s.vals, s.vstride, m, alphas_dev, n_alphas, nullptr, nullptr, result_dev);
        XC_LAUNCHED(ctx);
    }
    return XC_OK;
}

int reduce_grid(xc_ctx *ctx, int64_t m)
{
    int64_t blocks = (m + 255) / 256;
    int grid = (int)(blocks < XC_RED_MAX_BLOCKS ? blocks : XC_RED_MAX_BLOCKS);
    if (grid > ctx->sm_count * 4) grid = ctx->sm_count * 4;
    return grid;
}
}  // namespace

extern "C" int64_t xc_fw_alpha_scratch_bytes(int64_t m, int64_t n_alphas)
{
    return (int64_t)alpha_scratch(nullptr, m, n_alphas).bytes;
}

extern "C" int64_t xc_fw_alpha_ctl_offset(int64_t m, int64_t n_alphas)
{
    AlphaScratch s = alpha_scratch(nullptr, m, n_alphas);
    return (int64_t)(reinterpret_cast<uint8_t *>(s.ctl) - reinterpret_cast<uint8_t *>(s.vals));
}

extern "C" int xc_fw_alpha_search(xc_ctx *ctx, const xc_metric_params *p, const double *C, const double *Ci,
                                  int64_t m, const double *alphas_dev, int64_t n_alphas, double *scratch_dev,
                                  double *result_dev, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !p || !C || !Ci || !scratch_dev || !result_dev || m <= 0 || n_alphas < 0) return XC_ERR_INVALID;
    if (n_alphas > 0 && !alphas_dev) return XC_ERR_INVALID;
    if (p->metric < 0 || p->metric > XC_METRIC_HMEAN) return XC_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    AlphaScratch s = alpha_scratch(scratch_dev, m, n_alphas);
    if (alpha_two_stage(p, n_alphas)) {
        fw_alpha_prep_kernel<<<(unsigned)((m + 1 + kThreads - 1) / kThreads), kThreads, 0, st>>>(*p, C, Ci, m, s.lin, s.linE);
        XC_LAUNCHED(ctx);
    }
    return alpha_search_launch(ctx, p, C, Ci, m, alphas_dev, n_alphas, s, result_dev, st);
}

extern "C" int xc_fw_alpha_ternary(xc_ctx *ctx, const xc_metric_params *p, const double *C, const double *Ci,
                                   int64_t m, double eps, double *result_dev, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !p || !C || !Ci || !result_dev || m <= 0 || !(eps > 0.0)) return XC_ERR_INVALID;
    if (p->metric < 0 || p->metric > XC_METRIC_HMEAN) return XC_ERR_INVALID;
    fw_alpha_ternary_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(*p, C, Ci, m, eps, result_dev);
    XC_LAUNCHED(ctx);
    return XC_OK;
}

extern "C" int xc_fw_combine(xc_ctx *ctx, double *C, const double *Ci, int64_t m4, const double *alpha_dev,
                             void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !C || !Ci || !alpha_dev || m4 <= 0) return XC_ERR_INVALID;
    fw_combine_kernel<<<(unsigned)((m4 + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(C, Ci, m4,
                                                                                                          alpha_dev);
    XC_LAUNCHED(ctx);
    return XC_OK;
}

// ---- one dense Frank-Wolfe iteration as two host calls (the all-reduce of the iterate's raw sums, if
// any, happens between them) -----------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(kThreads)
f32_to_f64_kernel(const float *a, const float *b, double *oa, double *ob, int64_t m)
{
    int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (j < m) {
        oa[j] = (double)a[j];
        ob[j] = (double)b[j];
    }
}
}  // namespace

extern "C" int xc_fw_step_begin(xc_ctx *ctx, const void *eta, int dtype, int64_t n, int64_t m, int64_t ld,
                                const void *y_true, int64_t ld_true, const float *a_row, const float *b_row,
                                double *ab64, int k, double *raw, int raw_is_zero, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !eta || !y_true || !a_row || !b_row || !raw) return XC_ERR_INVALID;
    if (n <= 0 || m <= 0 || ld < m || ld_true < m || k < 0 || k > 32 || k > m) return XC_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const void *a = a_row, *b = b_row;
    if (dtype == XC_F64) {  // numpy promotes the float32 classifier rows to float64 gains
        if (!ab64) return XC_ERR_INVALID;
        double *b64 = ab64 + ((m + 1) & ~(int64_t)1);  // keeps the second vector 16-byte aligned
        f32_to_f64_kernel<<<(unsigned)((m + kThreads - 1) / kThreads), kThreads, 0, st>>>(a_row, b_row, ab64, b64, m);
        XC_LAUNCHED(ctx);
        a = ab64;
        b = b64;
    } else if (dtype != XC_F32) {
        return XC_ERR_UNSUPPORTED;
    }
    if (!raw_is_zero) XC_CUDA_TRY(ctx, cudaMemsetAsync(raw, 0, sizeof(double) * 2 * m, st));
    if (dtype == XC_F32)
        return launch_fw_dense<float>(ctx, eta, n, m, ld, y_true, ld_true, a, b, k, raw, raw + m, nullptr, st);
    return launch_fw_dense<double>(ctx, eta, n, m, ld, y_true, ld_true, a, b, k, raw, raw + m, nullptr, st);
}

extern "C" int xc_fw_step_finish(xc_ctx *ctx, const xc_metric_params *p, int first, double *raw,
                                 const double *colsum, int64_t m, double n_global, int normalize, int skip_tn,
                                 double *Cm, double *Ci, const double *alphas_dev, int64_t n_alphas,
                                 double fixed_alpha, double *scratch_dev, double *scal, float *a_next, float *b_next,
                                 double *scal_next, int zero_raw, double ternary_eps, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !p || !raw || !colsum || !Cm || !Ci || !scal || m <= 0) return XC_ERR_INVALID;
    if ((a_next == nullptr) != (b_next == nullptr)) return XC_ERR_INVALID;
    if (p->metric < 0 || p->metric > XC_METRIC_HMEAN) return XC_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = reduce_grid(ctx, m);
    double *zr = zero_raw ? raw : nullptr;
    if (p->mix == 2) {  // micro-averaged objective: scalar value / gradient / line search in one CTA
        const unsigned gm = (unsigned)((m + kThreads - 1) / kThreads);
        fw_make_conf_kernel<<<gm, kThreads, 0, st>>>(raw, raw + m, colsum, m, n_global, normalize, skip_tn,
                                                     first ? Cm : Ci);
        XC_LAUNCHED(ctx);
        fw_micro_kernel<<<1, 1024, 0, st>>>(*p, Cm, Ci, m, first, alphas_dev, n_alphas, fixed_alpha, ternary_eps, scal,
                                            scal_next);
        XC_LAUNCHED(ctx);
        if (!first) {
            fw_combine_kernel<<<(unsigned)((4 * m + kThreads - 1) / kThreads), kThreads, 0, st>>>(Cm, Ci, 4 * m, scal + 2);
            XC_LAUNCHED(ctx);
        }
        if (a_next || zr) {
            fw_micro_fill_kernel<<<gm, kThreads, 0, st>>>(scal + 5, m, a_next, b_next, zr);
            XC_LAUNCHED(ctx);
        }
        return XC_OK;
    }
    if (first) {  // classifier 0: its confusion vectors ARE the running ones (frank_wolfe.py:564-572)
        fw_conf_prep_kernel<<<grid, 256, 0, st>>>(*p, raw, raw + m, colsum, m, n_global, normalize, skip_tn, nullptr, Cm,
                                                  nullptr, nullptr, scal + 0, ctx->red_partials, ctx->red_counter);
        XC_LAUNCHED(ctx);
        if (a_next || zr) {
            fw_finish_kernel<<<grid, 256, 0, st>>>(*p, Cm, nullptr, m, nullptr, a_next, b_next, zr, nullptr,
                                                   scal_next, ctx->red_partials, ctx->red_counter);
            XC_LAUNCHED(ctx);
        }
        return XC_OK;
    }
    const bool ternary = ternary_eps > 0.0;
    const bool search = alphas_dev != nullptr && !ternary;
    if (search && !scratch_dev) return XC_ERR_INVALID;
    AlphaScratch s = alpha_scratch(scratch_dev, m, n_alphas);
    const bool two_stage = search && alpha_two_stage(p, n_alphas);
    fw_conf_prep_kernel<<<grid, 256, 0, st>>>(*p, raw, raw + m, colsum, m, n_global, normalize, skip_tn, Cm, Ci,
                                              two_stage ? s.lin : nullptr, s.linE, scal + 1, ctx->red_partials,
                                              ctx->red_counter);
    XC_LAUNCHED(ctx);
    if (ternary) {
        int rc = xc_fw_alpha_ternary(ctx, p, Cm, Ci, m, ternary_eps, scal + 2, stream);
        if (rc) return rc;
    } else if (search) {
        int rc = alpha_search_launch(ctx, p, Cm, Ci, m, alphas_dev, n_alphas, s, scal + 2, st);
        if (rc) return rc;
    } else {
        XC_CUDA_TRY(ctx, cudaMemcpyAsync(scal + 2, &fixed_alpha, sizeof(double), cudaMemcpyHostToDevice, st));
    }
    fw_finish_kernel<<<grid, 256, 0, st>>>(*p, Cm, Ci, m, scal + 2, a_next, b_next, zr, scal + 4, scal_next,
                                           ctx->red_partials, ctx->red_counter);
    XC_LAUNCHED(ctx);
    return XC_OK;
}
