// Batched block-Jacobi BCA sweep: per-label coefficient kernel + streaming batch kernels.
//
// Replaces the per-instance step of xcolumns/block_coordinate.py:132-293 (dense + CSR) and
// :539-580 (coverage) evaluated against a state frozen for one batch of rows.  For metrics whose
// marginal gain is affine in eta (precision / recall / F-beta)
//     unselected label:  gain = A_j  + B_j  * eta_ij
//     selected label:    gain = A'_j + B'_j * eta_ij      (own contribution removed)
// so the dense batch is ONE streaming pass: 16-byte loads of eta, one FMA per element, a
// warp-distributed top-k list seeded with the row's current selection (which gives a tight
// threshold from the first element on), then float64 atomic deltas for the rows that changed.
#include <cstdlib>

#include "xc_scan.cuh"
#include "xc_tma.cuh"

namespace {

constexpr int kThreads = 256;

// ---- coefficient kernel -------------------------------------------------------------------------
// D = u*(tp+fp) + v*(tp+fn) + E with E = eps*n_div (eps is added AFTER dividing by n in the
// reference, block_coordinate.py:176-183 + metrics.py), c = numerator factor:
//   precision (c,u,v) = (1,1,0); recall (1,0,1); F-beta (1+b^2, b^2, 1).
__device__ __forceinline__ void bca_coef_of(const xc_metric_params &p, double t, double f, double g, float2 *cn,
                                            float2 *cs)
{
    const double sgn = p.maximize ? 1.0 : -1.0;
    const double E = p.eps * p.n_div;
    double Bn, An, Bs, As;  // gain = A + B * eta for an unselected (n) / currently selected (s) label
    if (p.metric == XC_METRIC_BALANCED_ACC) {
        // tp + fn = sum_i eta_ij and tn + fp = sum_i (1 - eta_ij) do not depend on the prediction, so
        // gain = [eta / (S + E) - (1 - eta) / (NS + E)] / 2 for selected and unselected labels alike
        const double S = t + g;
        const double NS = p.n_rows - S;
        Bn = Bs = 0.5 * (1.0 / (S + E) + 1.0 / (NS + E));
        An = As = -0.5 / (NS + E);
    } else if (p.metric == XC_METRIC_PREC_AT_K) {
        // tp / k on the normalised counts (metrics.py:513): every label gains eta / (k n)
        Bn = Bs = 1.0 / (p.c1 * p.n_div);
        An = As = 0.0;
    } else {
        double c, u, v;
        if (p.metric == XC_METRIC_PRECISION) { c = 1.0; u = 1.0; v = 0.0; }
        else if (p.metric == XC_METRIC_RECALL) { c = 1.0; u = 0.0; v = 1.0; }
        else { c = p.c1; u = p.beta2; v = 1.0; }
        const double D = u * (t + f) + v * (t + g) + E;
        const double ct = c * t;
        // unselected: predict j for this row -> D grows by u
        Bn = c / (D + u);
        An = ct / (D + u) - ct / D;
        // selected: un-predicting j -> D shrinks by u
        Bs = c / (D - u);
        As = ct / D - ct / (D - u);
    }
    if (p.mix == 1) {
        // (1 - alpha) * tp / k + alpha * metric / m, summed over the labels (block_coordinate.py:848-1045):
        // the instance-precision part adds (1 - alpha) eta / (k n) to every gain
        const double w1 = (1.0 - p.mix_alpha) / (p.mix_k * p.n_div), w2 = p.mix_alpha / p.mix_m;
        Bn = w1 + w2 * Bn; An = w2 * An;
        Bs = w1 + w2 * Bs; As = w2 * As;
    }
    *cn = make_float2((float)(sgn * Bn), (float)(sgn * An));
    *cs = make_float2((float)(sgn * Bs), (float)(sgn * As));
}


__global__ void __launch_bounds__(kThreads)
bca_coef_kernel(xc_metric_params p, double *tp, double *fp, double *fn, double *dtp, double *dfp, double *dfn,
                int64_t m, float2 *coef_n, float2 *coef_s)
{
    const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (j >= m) return;
    double t = tp[j], f = fp[j], g = fn[j];
    if (dtp) {
        t += dtp[j]; f += dfp[j]; g += dfn[j];
        tp[j] = t; fp[j] = f; fn[j] = g;
        dtp[j] = 0.0; dfp[j] = 0.0; dfn[j] = 0.0;
    }
    bca_coef_of(p, t, f, g, coef_n + j, coef_s + j);
}

__device__ __forceinline__ float4 bca_rec_of(const xc_metric_params &p, double t, double f, double g);

// ---- commit over peer memory (rows sharded over the GPUs of one box): system-scope flag helpers ---------------
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Exchange of one batch's deltas between the ranks of a box, PUSH model.  After its batch kernel every rank copies
// its delta buffer into its slot of every peer's "inbox" (remote STORES over NVLink are posted: no round trip),
// fences, and the last CTA to finish raises the rank's flag in every peer's window.  The commit kernel then waits
// for all flags and reads nothing but LOCAL memory.  (First version: the commit kernel PULLED the peers' buffers
// with remote loads -- three dependent rounds of NVLink round trips per label; measured on 8 x B200: 62 us per
// commit, which made the commit chain, not the streaming, the critical path of a strong-scaled sweep.)
// Window payload: [NB own delta buffers | world x NB inbox buffers], each xc_bca_delta_stride(m) bytes.
// A slot is overwritten NB = 2 (lag + 1) commits later; by then its reader has left the commit that read it (see the
// flag argument at xc_bca_pipe_sweep).
constexpr int kCommitThreads = 64;

__global__ void __launch_bounds__(kCommitThreads)
bca_push_kernel(uint8_t *const *windows, int world, int rank, unsigned epoch, int64_t own_off, int64_t inbox_off,
                int64_t stride, int nb, int cur, int64_t m)
{
    __shared__ int s_last;
    uint8_t *mine = windows[rank];
    const double *src = reinterpret_cast<const double *>(mine + own_off + (int64_t)cur * stride);
    const int64_t slot = inbox_off + ((int64_t)rank * nb + cur) * stride;
    for (int64_t i = (int64_t)blockIdx.x * kCommitThreads + threadIdx.x; i < 3 * m; i += (int64_t)gridDim.x * kCommitThreads) {
        const double v = src[i];
        for (int r = 0; r < world; ++r)
            if (r != rank) reinterpret_cast<double *>(windows[r] + slot)[i] = v;
    }
    __threadfence_system();   // this thread's stores are visible system-wide before the ticket
    __syncthreads();
    // (one ticket and one flag per buffer: the pushes of consecutive batches run on different streams and may
    //  overlap or finish out of order; pushes into the same buffer are NB batches apart on one stream)
    unsigned *ticket = reinterpret_cast<unsigned *>(mine) + XC_P2P_TICKET_WORD + cur;
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last) {   // every CTA of this rank has pushed buffer `cur`: tell the peers
        if (threadIdx.x == 0) *ticket = 0;
        if (threadIdx.x < world) {
            __threadfence_system();
            st_release_sys(reinterpret_cast<unsigned *>(windows[threadIdx.x]) + cur * XC_P2P_MAX_WORLD + rank, epoch);
        }
    }
}

// One commit = fold the deltas of one batch into the float64 state, refresh the gain coefficients (or the
// Jaccard / G-mean / H-mean records) and clear the delta buffer a later batch will accumulate into.
//   windows == nullptr : single process, the deltas are this GPU's own buffer `cur`
//   windows != nullptr : wait for every rank's flag (bca_push_kernel), then add the W buffers -- the peers' from the
//                        local inbox, in rank order, so every rank adds the same numbers in the same order and the
//                        replicated state stays bit-equal.  A rank that waits longer than ~4 s raises the window's
//                        error word and leaves.
//   cur < 0            : no fold -- coefficients of the current state only (start of a sweep)
// 64 threads and <= 64 registers per CTA: 4096 registers, which is what six resident CTAs of the streaming
// batch kernel (6 x 256 x 40) leave free on an SM, so a commit never has to wait for a streaming CTA to
// retire when the next batch is already running (pipelined sweep, see xc_bca_pipe_sweep).
__global__ void __launch_bounds__(kCommitThreads, 16)
bca_commit_kernel(xc_metric_params p, double *tp, double *fp, double *fn, uint8_t *const *windows, int world, int rank,
                  unsigned epoch, double *local, int64_t inbox_off, int64_t stride, int nb, int cur, int clr, int64_t m,
                  float2 *coef_n, float2 *coef_s, float2 *coef_n2, float2 *coef_s2, int rec)
{
    __shared__ int s_fail;
    const uint8_t *mine = windows ? windows[rank] : nullptr;
    if (windows && cur >= 0) {
        if (threadIdx.x == 0) s_fail = 0;
        __syncthreads();
        if (threadIdx.x < world) {
            const unsigned *flag = reinterpret_cast<const unsigned *>(mine) + cur * XC_P2P_MAX_WORLD + threadIdx.x;
            const unsigned long long t0 = global_timer_ns();
            // the flag of (buffer, sender) only grows; a peer may already be a rotation ahead (wrap-safe signed distance)
            while ((int)(ld_acquire_sys(flag) - epoch) < 0) {
                if (global_timer_ns() - t0 > 4000000000ULL) {
                    s_fail = 1;
                    reinterpret_cast<unsigned *>(windows[rank])[XC_P2P_ERR_WORD] = epoch;
                    break;
                }
            }
        }
        __syncthreads();
        if (s_fail) return;
    }
    for (int64_t j = (int64_t)blockIdx.x * kCommitThreads + threadIdx.x; j < m; j += (int64_t)gridDim.x * kCommitThreads) {
        double t = tp[j], f = fp[j], g = fn[j];
        if (cur >= 0) {
            const double *dl = reinterpret_cast<const double *>(reinterpret_cast<const uint8_t *>(local) +
                                                                (int64_t)cur * stride);
            double d[3] = {0.0, 0.0, 0.0};
            if (windows) {   // every rank's buffer is local memory by now; added in rank order
#pragma unroll 4
                for (int r = 0; r < world; ++r) {
                    const double *src = r == rank ? dl
                                                  : reinterpret_cast<const double *>(mine + inbox_off +
                                                                                     ((int64_t)r * nb + cur) * stride);
                    d[0] += __ldcv(src + j);
                    d[1] += __ldcv(src + m + j);
                    d[2] += __ldcv(src + 2 * m + j);
                }
            } else {
                d[0] = dl[j]; d[1] = dl[m + j]; d[2] = dl[2 * m + j];
            }
            t += d[0]; f += d[1]; g += d[2];
            tp[j] = t; fp[j] = f; fn[j] = g;
        }
        if (clr >= 0) {
            double *nxt = reinterpret_cast<double *>(reinterpret_cast<uint8_t *>(local) + (int64_t)clr * stride);
            nxt[j] = 0.0; nxt[m + j] = 0.0; nxt[2 * m + j] = 0.0;
        }
        if (rec) {
            reinterpret_cast<float4 *>(coef_n)[j] = bca_rec_of(p, t, f, g);   // coef_n is the 16-byte record array here
        } else {
            float2 cn, cs;
            bca_coef_of(p, t, f, g, &cn, &cs);
            coef_n[j] = cn; coef_s[j] = cs;
            if (coef_n2) { coef_n2[j] = cn; coef_s2[j] = cs; }
        }
    }
}

// ---- shared epilogue: compare new selection with the old one, emit deltas, store the row ----------
// lanes < k hold: old label / old eta (sorted as stored) and the new list entry (tk.idx).
template <typename TE>
__device__ __forceinline__ void bca_commit_row(const TE *__restrict__ row_ptr, int k, int old_j, TE old_e,
                                               int new_j_unsorted, int32_t *__restrict__ pred_row, double *dtp,
                                               double *dfp, double *dfn)
{
    const int lane = lane_id();
    // sort new labels ascending
    int src = warp_rank_src(new_j_unsorted, k);
    int new_j = __shfl_sync(XC_FULL, new_j_unsorted, src);
    bool stays_old = false, stays_new = false;
    for (int t = 0; t < k; ++t) {
        int nj = __shfl_sync(XC_FULL, new_j, t);
        int oj = __shfl_sync(XC_FULL, old_j, t);
        stays_old |= (nj == old_j);
        stays_new |= (oj == new_j);
    }
    if (lane < k) {
        const TE one = (TE)1;
        if (!stays_old && old_j >= 0) {  // label leaves the prediction
            atomicAdd(dtp + old_j, -(double)old_e);
            atomicAdd(dfp + old_j, -(double)(TE)(one - old_e));
            atomicAdd(dfn + old_j, (double)old_e);
        }
        if (!stays_new && new_j != 0x7fffffff) {  // label enters
            TE e = row_ptr[new_j];
            atomicAdd(dtp + new_j, (double)e);
            atomicAdd(dfp + new_j, (double)(TE)(one - e));
            atomicAdd(dfn + new_j, -(double)e);
        }
        pred_row[lane] = new_j == 0x7fffffff ? -1 : new_j;
    }
}

// seed a list with the row's current selection evaluated under the "selected" coefficients
template <typename G>
__device__ __forceinline__ void seed_list(WarpTopK<G> &tk, G g, int j, int k)
{
    // k sequential warp-uniform inserts (k <= 32, once per row)
    tk.init();
    for (int t = 0; t < k; ++t) {
        G gt = __shfl_sync(XC_FULL, g, t);
        int jt = __shfl_sync(XC_FULL, j, t);
        if (jt >= 0) tk.insert(gt, jt, k);
    }
}

template <typename TE, int R, bool DEEP = false>
__global__ void __launch_bounds__(kThreads, DEEP ? 4 : ((R == 1 && sizeof(TE) == 4) ? 6 : 1))
bca_batch_dense_kernel(const TE *__restrict__ eta, int64_t m, int64_t ld, const int32_t *__restrict__ rows,
                       int64_t n_rows, int k, const float2 *__restrict__ coef_n, const float2 *__restrict__ coef_s,
                       int32_t *__restrict__ pred_idx, double *dtp, double *dfp, double *dfn, bool vec_ok,
                       int32_t *__restrict__ snap)
{
    const int lane = lane_id();
    const int wpc = (int)(blockDim.x >> 5);   // 8 warps per CTA, or 2 for launches of one row per warp (see launcher)
    const int64_t warp = (int64_t)blockIdx.x * wpc + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * wpc;
    XfAffine xf{coef_n};
    for (int64_t grp = warp; grp * R < n_rows; grp += nwarps) {
        const TE *rp[R];
        int64_t row_id[R];
        int old_j[R];
        TE old_e[R];
        WarpTopK<float> tk[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int64_t i = grp * R + r;
            bool valid = i < n_rows;
            if (!valid) i = grp * R;
            int64_t row = rows ? (int64_t)rows[i] : i;
            row_id[r] = valid ? row : -1;
            rp[r] = eta + row * ld;
            old_j[r] = -1;
            old_e[r] = (TE)0;
            float g = 0.f;
            if (lane < k) {
                old_j[r] = pred_idx[row * k + lane];
                // the row's selection BEFORE this sweep, kept for roll-backs (every row is visited once per sweep,
                // so the snapshot is complete when the sweep is)
                if (snap && valid) snap[row * k + lane] = old_j[r];
                if (old_j[r] >= 0) {
                    old_e[r] = rp[r][old_j[r]];
                    float2 cs = __ldg(coef_s + old_j[r]);
                    g = fmaf(cs.x, (float)old_e[r], cs.y);
                }
            }
            seed_list(tk[r], g, old_j[r], k);
        }
        xc_scan_rows<TE, float, R, true, XfAffine, DEEP>(rp, m, vec_ok, xf, tk, old_j, k);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (row_id[r] >= 0)  // warp-uniform
                bca_commit_row<TE>(rp[r], k, old_j[r], old_e[r], tk[r].idx, pred_idx + row_id[r] * k, dtp, dfp, dfn);
        }
    }
}

// ---- metrics whose gain is NOT affine in eta: Jaccard, G-mean, H-mean ------------------------------------
// Against a frozen state the gain of an unselected label is still a closed form of eta and four per-label
// numbers.  Evaluating psi(C + row) - psi(C - row) directly in float32 would lose the gain (~1/n) in the
// rounding of psi (~1); the DIFFERENCE is therefore expanded algebraically so that float32 keeps ~1e-6
// relative accuracy of the gain itself:
//   Jaccard  (D = tp+fp+fn+E):        g = (eta (D + tp) - tp) / (D (D + 1 - eta))
//   with x = tp/P, y = tn/N (P = tp+fn+E and N = tn+fp+E do not depend on the prediction),
//        dx = eta/P, dy = -(1-eta)/N, x' = x+dx, y' = y+dy:
//   G-mean:  g = (x dy + y dx + dx dy) / (sqrt(x' y') + sqrt(x y))
//   H-mean:  g = 2 (x x' dy + y y' dx) / ((x' + y') (x + y))
// The k currently selected labels of a row (own contribution to remove) are evaluated in float64 with
// the reference's expression straight from the state vectors.
__device__ __forceinline__ float4 bca_rec_of(const xc_metric_params &p, double t, double f, double g)
{
    const double E = p.eps * p.n_div;
    if (p.metric == XC_METRIC_JACCARD) {
        const double D = t + f + g + E;
        return make_float4((float)(D + t), (float)t, (float)D, (float)(D + 1.0));
    }
    const double tn = p.n_rows - t - f - g;
    const double P = t + g + E, N = tn + f + E;
    return make_float4((float)(t / P), (float)(tn / N), (float)(1.0 / P), (float)(1.0 / N));
}

template <int METRIC>
struct XfRecord {
    const float4 *rec;
    float sgn;
    // mixed utility (1 - alpha) tp / k + alpha metric / m (block_coordinate.py:960-1045): the instance-precision part
    // adds w1 * eta to every gain, the metric part is scaled by w2; (0, 1) for the plain metric
    float w1, w2;
    __device__ static XfRecord make(const float4 *rec, const xc_metric_params &p)
    {
        XfRecord x{rec, p.maximize ? 1.f : -1.f, 0.f, 1.f};
        if (p.mix == 1) {
            x.w1 = (float)((1.0 - p.mix_alpha) / (p.mix_k * p.n_div));
            x.w2 = (float)(p.mix_alpha / p.mix_m);
        }
        return x;
    }
    __device__ __forceinline__ float gain(const float4 r, float e) const
    {
        float g;
        if (METRIC == XC_METRIC_JACCARD) {
            g = fmaf(e, r.x, -r.y) * __frcp_rn(r.z * (r.w - e));
        } else {
            const float dx = e * r.z, dy = (e - 1.f) * r.w;
            const float xn = r.x + dx, yn = r.y + dy;
            if (METRIC == XC_METRIC_GMEAN) {
                const float num = fmaf(r.x, dy, fmaf(r.y, dx, dx * dy));
                const float den = sqrtf(fmaxf(xn * yn, 0.f)) + sqrtf(r.x * r.y);
                g = num * __frcp_rn(fmaxf(den, 1e-30f));
            } else {
                const float num = 2.f * fmaf(r.x * xn, dy, r.y * yn * dx);
                g = num * __frcp_rn(fmaxf((xn + yn) * (r.x + r.y), 1e-30f));
            }
        }
        return sgn * fmaf(w2, g, w1 * e);
    }
    template <typename TE, int V>
    __device__ __forceinline__ void apply_vec(int64_t c, const TE (&e)[V], float (&g)[V], float (&ca)[V],
                                              float (&cb)[V], bool first) const
    {
#pragma unroll
        for (int v = 0; v < V; ++v) g[v] = gain(__ldg(rec + c + v), (float)e[v]);
    }
    template <typename TE>
    __device__ __forceinline__ float apply_one(int64_t c, TE e) const { return gain(__ldg(rec + c), (float)e); }
};

__global__ void __launch_bounds__(kThreads)
bca_rec_kernel(xc_metric_params p, double *tp, double *fp, double *fn, double *dtp, double *dfp, double *dfn, int64_t m,
               float4 *rec)
{
    const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (j >= m) return;
    double t = tp[j], f = fp[j], g = fn[j];
    if (dtp) {
        t += dtp[j]; f += dfp[j]; g += dfn[j];
        tp[j] = t; fp[j] = f; fn[j] = g;
        dtp[j] = 0.0; dfp[j] = 0.0; dfn[j] = 0.0;
    }
    rec[j] = bca_rec_of(p, t, f, g);
}

// gain of keeping a currently selected label (float64, reference expression, own contribution removed)
__device__ __forceinline__ double bca_selected_gain(const xc_metric_params &p, double t, double f, double g, double e,
                                                    double om)
{
    const double nd = p.n_div;
    const double tn = p.n_rows - t - f - g;
    const double keep = xc_metric_eval(p, t / nd, f / nd, g / nd, tn / nd);
    const double drop = xc_metric_eval(p, (t - e) / nd, (f - om) / nd, (g + e) / nd, (tn + om) / nd);
    const double d = keep - drop;
    return p.maximize ? d : -d;
}

template <typename TE, int METRIC>
__global__ void __launch_bounds__(kThreads)
bca_batch_dense_rec_kernel(xc_metric_params p, const TE *__restrict__ eta, int64_t m, int64_t ld,
                           const int32_t *__restrict__ rows, int64_t n_rows, int k, const float4 *__restrict__ rec,
                           const double *__restrict__ tp, const double *__restrict__ fp,
                           const double *__restrict__ fn, int32_t *__restrict__ pred_idx, double *dtp, double *dfp,
                           double *dfn, bool vec_ok, int32_t *__restrict__ snap)
{
    const int lane = lane_id();
    const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    const XfRecord<METRIC> xf = XfRecord<METRIC>::make(rec, p);
    const TE one = (TE)1;
    for (int64_t i = warp; i < n_rows; i += nwarps) {
        const int64_t row = rows ? (int64_t)rows[i] : i;
        const TE *rp[1] = {eta + row * ld};
        int old_j[1] = {-1};
        TE old_e = (TE)0;
        float g = 0.f;
        if (lane < k) {
            old_j[0] = pred_idx[row * k + lane];
            if (snap) snap[row * k + lane] = old_j[0];
            if (old_j[0] >= 0) {
                const int j = old_j[0];
                old_e = rp[0][j];
                g = (float)bca_selected_gain(p, tp[j], fp[j], fn[j], (double)old_e, (double)(TE)(one - old_e));
            }
        }
        WarpTopK<float> tk[1];
        seed_list(tk[0], g, old_j[0], k);
        xc_scan_rows<TE, float, 1, true>(rp, m, vec_ok, xf, tk, old_j, k);
        bca_commit_row<TE>(rp[0], k, old_j[0], old_e, tk[0].idx, pred_idx + row * k, dtp, dfp, dfn);
    }
}

// ---- coverage ------------------------------------------------------------------------------------
// gain = Ef_j * eta (unselected) | Ef_j / (1 - eta) * eta (selected), optionally mixed with
// precision@k: alpha * gain + (1 - alpha) * eta / k   (block_coordinate.py:562-569)
__device__ __forceinline__ void atomic_mul(double *addr, double f)
{
    unsigned long long *a = reinterpret_cast<unsigned long long *>(addr);
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        old = atomicCAS(a, assumed, __double_as_longlong(__longlong_as_double(assumed) * f));
    } while (assumed != old);
}

template <typename T>
__device__ __forceinline__ double cov_gain(double Ef, T eta, bool sel, double alpha, int k)
{
    double ef = sel ? Ef / (double)(T)((T)1 - eta) : Ef;
    double g = ef * (double)eta;
    if (alpha < 1.0) g = alpha * g + (1.0 - alpha) * (double)eta / (double)k;
    return g;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
cov_batch_csr_kernel(const T *__restrict__ data, const int32_t *__restrict__ indices,
                     const int64_t *__restrict__ indptr, const int32_t *__restrict__ rows, int64_t n_rows, int k,
                     double alpha, const double *__restrict__ Ef, int32_t *__restrict__ pred_idx, double *dEf)
{
    const int lane = lane_id();
    const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    const T one = (T)1;
    for (int64_t w = warp; w < n_rows; w += nwarps) {
        const int64_t row = rows ? (int64_t)rows[w] : w;
        const int64_t s = indptr[row], e = indptr[row + 1];
        int32_t *pred_row = pred_idx + row * k;
        int old_j = -1;
        if (lane < k) old_j = pred_row[lane];
        WarpTopK<double> tk;
        tk.init();
        for (int64_t q0 = s; q0 < e; q0 += 32) {
            int64_t q = q0 + lane;
            double g[1];
            g[0] = NAN;
            const int j = q < e ? indices[q] : -2;
            bool sel = false;
            for (int t = 0; t < k; ++t) sel |= (__shfl_sync(XC_FULL, old_j, t) == j);
            if (q < e) g[0] = cov_gain<T>(Ef[j], data[q], sel, alpha, k);
            if (__any_sync(XC_FULL, tk.passes(g[0]))) xc_scan_insert<double, 1, false>(tk, g, q0 - s, 1, k, -1);
        }
        int src = warp_rank_src(tk.idx, k);
        int pos = __shfl_sync(XC_FULL, tk.idx, src);
        int new_j = pos == 0x7fffffff ? 0x7fffffff : indices[s + pos];
        T new_e = pos == 0x7fffffff ? (T)0 : data[s + pos];
        bool stays_old = false, stays_new = false;
        for (int t = 0; t < k; ++t) {
            int nj = __shfl_sync(XC_FULL, new_j, t);
            int oj = __shfl_sync(XC_FULL, old_j, t);
            stays_old |= (nj == old_j);
            stays_new |= (oj == new_j);
        }
        if (lane < k) {
            if (!stays_old && old_j >= 0) {
                int64_t y = s, z = e;
                while (y < z) {
                    int64_t mid = (y + z) >> 1;
                    int v = indices[mid];
                    if (v == old_j) { atomic_mul(dEf + old_j, 1.0 / (double)(T)(one - data[mid])); break; }
                    if (v < old_j) y = mid + 1; else z = mid;
                }
            }
            if (!stays_new && new_j != 0x7fffffff) atomic_mul(dEf + new_j, (double)(T)(one - new_e));
            pred_row[lane] = new_j == 0x7fffffff ? -1 : new_j;
        }
    }
}

// dense coverage batch: gains need Ef_j in float64 per element -> a = Ef as float32 coefficient would
// lose the 1e-4 contract only marginally, but coverage products span many decades; keep float64.
template <typename T>
struct XfCov {
    const double *Ef;
    double alpha;
    double inv_k;
    template <typename TE, int V>
    __device__ __forceinline__ void apply_vec(int64_t c, const TE (&e)[V], double (&g)[V], double (&ca)[V],
                                              double (&cb)[V], bool first) const
    {
        if (first) {
#pragma unroll
            for (int v = 0; v < V; ++v) ca[v] = __ldg(Ef + c + v);
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
            double x = ca[v] * (double)e[v];
            if (alpha < 1.0) x = alpha * x + (1.0 - alpha) * (double)e[v] * inv_k;
            g[v] = x;
        }
    }
    template <typename TE>
    __device__ __forceinline__ double apply_one(int64_t c, TE e) const
    {
        double x = __ldg(Ef + c) * (double)e;
        if (alpha < 1.0) x = alpha * x + (1.0 - alpha) * (double)e * inv_k;
        return x;
    }
};

template <typename TE, int R>
__global__ void __launch_bounds__(kThreads)
cov_batch_dense_kernel(const TE *__restrict__ eta, int64_t m, int64_t ld, const int32_t *__restrict__ rows,
                       int64_t n_rows, int k, double alpha, const double *__restrict__ Ef,
                       int32_t *__restrict__ pred_idx, double *dEf, bool vec_ok)
{
    const int lane = lane_id();
    const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    XfCov<TE> xf{Ef, alpha, 1.0 / (double)k};
    const TE one = (TE)1;
    for (int64_t grp = warp; grp * R < n_rows; grp += nwarps) {
        const TE *rp[R];
        int64_t row_id[R];
        int old_j[R];
        TE old_e[R];
        WarpTopK<double> tk[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int64_t i = grp * R + r;
            bool valid = i < n_rows;
            if (!valid) i = grp * R;
            int64_t row = rows ? (int64_t)rows[i] : i;
            row_id[r] = valid ? row : -1;
            rp[r] = eta + row * ld;
            old_j[r] = -1;
            old_e[r] = (TE)0;
            double g = 0.0;
            if (lane < k) {
                old_j[r] = pred_idx[row * k + lane];
                if (old_j[r] >= 0) {
                    old_e[r] = rp[r][old_j[r]];
                    g = cov_gain<TE>(Ef[old_j[r]], old_e[r], true, alpha, k);
                }
            }
            seed_list(tk[r], g, old_j[r], k);
        }
        xc_scan_rows<TE, double, R, true>(rp, m, vec_ok, xf, tk, old_j, k);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (row_id[r] < 0) continue;
            int src = warp_rank_src(tk[r].idx, k);
            int new_j = __shfl_sync(XC_FULL, tk[r].idx, src);
            bool stays_old = false, stays_new = false;
            for (int t = 0; t < k; ++t) {
                int nj = __shfl_sync(XC_FULL, new_j, t);
                int oj = __shfl_sync(XC_FULL, old_j[r], t);
                stays_old |= (nj == old_j[r]);
                stays_new |= (oj == new_j);
            }
            if (lane < k) {
                if (!stays_old && old_j[r] >= 0) atomic_mul(dEf + old_j[r], 1.0 / (double)(TE)(one - old_e[r]));
                if (!stays_new && new_j != 0x7fffffff) atomic_mul(dEf + new_j, (double)(TE)(one - rp[r][new_j]));
                pred_idx[row_id[r] * k + lane] = new_j == 0x7fffffff ? -1 : new_j;
            }
        }
    }
}

// =================================================================================================
// TMA-pipelined dense batch kernel (float32, 16-byte aligned rows).
//
// One CTA per SM, persistent over groups of 32 rows.  A producer warp streams the group's rows tile
// by tile (512 columns = 2 KB per row, plus the 4 KB coefficient tile from L2) into a 3-stage
// shared-memory ring with cp.async.bulk (TMA, 1-D) and mbarrier transaction counts; lane l of the
// producer owns row l of the group.  Eight consumer warps own 4 rows each: they read the tile with
// conflict-free LDS.128, reuse the coefficient chunk for their 4 rows, run the same
// FMA / max / compare common path as the LDG kernel and touch the top-k lists only on a hit.
// Up to 3 x 68 KB per SM are in flight regardless of what the consumer warps are doing, so a warp
// that sits in the list-update slow path no longer stops the HBM stream.
// =================================================================================================
constexpr int TMA_NCW = 8;                      // consumer warps
constexpr int TMA_RW = 4;                       // rows per consumer warp
constexpr int TMA_ROWS = TMA_NCW * TMA_RW;      // rows per CTA pass (= producer lanes)
constexpr int TMA_TC = 512;                     // columns per tile
constexpr int TMA_STAGES = 3;
constexpr int TMA_ETA_BYTES = TMA_ROWS * TMA_TC * 4;
constexpr int TMA_COEF_BYTES = TMA_TC * 8;
constexpr int TMA_STAGE_BYTES = TMA_ETA_BYTES + TMA_COEF_BYTES;
constexpr int TMA_SMEM_BYTES = TMA_STAGES * TMA_STAGE_BYTES + 2 * TMA_STAGES * 8;
constexpr int TMA_NPW = 4;                      // producer warps (bulk-copy issue is serialised per warp)
constexpr int TMA_PROWS = TMA_ROWS / TMA_NPW;   // rows each producer warp feeds
constexpr int TMA_THREADS = (TMA_NCW + TMA_NPW) * 32;
static_assert(TMA_PROWS <= 32, "one producer lane per row");

// list update for one row: only elements that pass the threshold are broadcast
template <bool SKIP>
__device__ __forceinline__ void tma_insert4(WarpTopK<float> &tk, float g0, float g1, float g2, float g3, int cbase,
                                            int k, int old_idx)
{
    const int lane = lane_id();
    unsigned pm = (tk.passes(g0) ? 1u : 0u) | (tk.passes(g1) ? 2u : 0u) | (tk.passes(g2) ? 4u : 0u) |
                  (tk.passes(g3) ? 8u : 0u);
    unsigned bal = __ballot_sync(XC_FULL, pm != 0);
    while (bal) {
        const int src = __ffs(bal) - 1;
        bal &= bal - 1;
        unsigned em = __shfl_sync(XC_FULL, pm, src);
        while (em) {
            const int i = __ffs(em) - 1;
            em &= em - 1;
            const float mine = i == 0 ? g0 : (i == 1 ? g1 : (i == 2 ? g2 : g3));
            const float gv = __shfl_sync(XC_FULL, mine, src);
            const int j = cbase + src * 4 + i;
            if (xc_better(gv, j, tk.thr, tk.thr_j)) {
                if (SKIP) {
                    if (__any_sync(XC_FULL, lane < k && old_idx == j)) continue;
                }
                tk.insert(gv, j, k);
            }
        }
    }
}

__global__ void __launch_bounds__(TMA_THREADS, 1)
bca_batch_dense_tma_kernel(const float *__restrict__ eta, int64_t m, int64_t ld, const int32_t *__restrict__ rows,
                           int64_t n_rows, int k, const float2 *__restrict__ coef_n,
                           const float2 *__restrict__ coef_s, int32_t *__restrict__ pred_idx, double *dtp,
                           double *dfp, double *dfn)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + TMA_STAGES * TMA_STAGE_BYTES);
    uint64_t *empty = full + TMA_STAGES;
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < TMA_STAGES; ++s) {
            mbar_init(full + s, TMA_NPW);
            mbar_init(empty + s, TMA_NCW);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int64_t n_groups = (n_rows + TMA_ROWS - 1) / TMA_ROWS;
    const int64_t mcopy = ((m + 3) / 4) * 4;  // <= ld (ld % 4 == 0): whole 16-byte units
    const int n_tiles = (int)((mcopy + TMA_TC - 1) / TMA_TC);
    uint32_t it = 0;  // running tile counter -> stage / phase

    if (warp >= TMA_NCW) {
        // ------------------------------ producers -----------------------------
        const int pw = warp - TMA_NCW;
        for (int64_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
            const int64_t i = g * TMA_ROWS + pw * TMA_PROWS + lane;
            const bool valid = lane < TMA_PROWS && i < n_rows;
            const int64_t row = valid ? (rows ? (int64_t)rows[i] : i) : 0;
            const float *src_row = eta + row * ld;
            const int nvalid = __popc(__ballot_sync(XC_FULL, valid));
            for (int t = 0; t < n_tiles; ++t, ++it) {
                const int s = it % TMA_STAGES;
                const uint32_t ph = (it / TMA_STAGES) & 1u;
                mbar_wait(empty + s, ph ^ 1u);  // slot free (first TMA_STAGES waits pass immediately)
                const int64_t c0 = (int64_t)t * TMA_TC;
                const int cols = (int)(mcopy - c0 < TMA_TC ? mcopy - c0 : TMA_TC);
                uint8_t *st = smem + s * TMA_STAGE_BYTES;
                if (lane == 0)
                    mbar_arrive_expect_tx(full + s, (uint32_t)(nvalid * cols * 4 + (pw == 0 ? TMA_COEF_BYTES : 0)));
                __syncwarp();
                if (valid)
                    bulk_g2s(st + (pw * TMA_PROWS + lane) * (TMA_TC * 4), src_row + c0, (uint32_t)(cols * 4), full + s);
                if (pw == 0 && lane == 0) bulk_g2s(st + TMA_ETA_BYTES, coef_n + c0, TMA_COEF_BYTES, full + s);
            }
        }
    } else {
        // ------------------------------ consumers -----------------------------
        const int w = warp;
        for (int64_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
            int64_t row_id[TMA_RW];
            const float *rp[TMA_RW];
            int old_j[TMA_RW];
            float old_e[TMA_RW];
            WarpTopK<float> tk[TMA_RW];
#pragma unroll
            for (int r = 0; r < TMA_RW; ++r) {
                const int64_t i = g * TMA_ROWS + w * TMA_RW + r;
                const bool valid = i < n_rows;
                const int64_t row = valid ? (rows ? (int64_t)rows[i] : i) : 0;
                row_id[r] = valid ? row : -1;
                rp[r] = eta + row * ld;
                old_j[r] = -1;
                old_e[r] = 0.f;
                float gs = 0.f;
                if (valid && lane < k) {
                    old_j[r] = pred_idx[row * k + lane];
                    if (old_j[r] >= 0) {
                        old_e[r] = rp[r][old_j[r]];
                        float2 cs = __ldg(coef_s + old_j[r]);
                        gs = fmaf(cs.x, old_e[r], cs.y);
                    }
                }
                seed_list(tk[r], gs, old_j[r], k);
                if (!valid) tk[r].thr = INFINITY;  // nothing passes: the smem row is never written
            }
            for (int t = 0; t < n_tiles; ++t, ++it) {
                const int s = it % TMA_STAGES;
                const uint32_t ph = (it / TMA_STAGES) & 1u;
                mbar_wait(full + s, ph);
                const uint8_t *st = smem + s * TMA_STAGE_BYTES;
                const float *se = reinterpret_cast<const float *>(st) + (w * TMA_RW) * TMA_TC;
                const float2 *sc = reinterpret_cast<const float2 *>(st + TMA_ETA_BYTES);
                const int64_t c0 = (int64_t)t * TMA_TC;
                const bool edge = c0 + TMA_TC > m;  // tile reaches past the last label
#pragma unroll
                for (int c = 0; c < TMA_TC / 128; ++c) {
                    const int colw = c * 128 + lane * 4;
                    const int col = (int)c0 + colw;
                    const float4 u = *reinterpret_cast<const float4 *>(sc + colw);
                    const float4 v = *reinterpret_cast<const float4 *>(sc + colw + 2);
                    float gq[TMA_RW][4];
                    bool hit = false;
#pragma unroll
                    for (int r = 0; r < TMA_RW; ++r) {
                        const float4 e = *reinterpret_cast<const float4 *>(se + r * TMA_TC + colw);
                        gq[r][0] = fmaf(u.x, e.x, u.y);
                        gq[r][1] = fmaf(u.z, e.y, u.w);
                        gq[r][2] = fmaf(v.x, e.z, v.y);
                        gq[r][3] = fmaf(v.z, e.w, v.w);
                        if (edge) {
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (col + q >= m) gq[r][q] = NAN;
                        }
                        hit |= tk[r].passes(fmaxf(fmaxf(gq[r][0], gq[r][1]), fmaxf(gq[r][2], gq[r][3])));
                    }
                    if (__any_sync(XC_FULL, hit)) {
#pragma unroll
                        for (int r = 0; r < TMA_RW; ++r)
                            tma_insert4<true>(tk[r], gq[r][0], gq[r][1], gq[r][2], gq[r][3], (int)c0 + c * 128, k,
                                              old_j[r]);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + s);  // this warp is done with the slot
            }
#pragma unroll
            for (int r = 0; r < TMA_RW; ++r) {
                if (row_id[r] >= 0)
                    bca_commit_row<float>(rp[r], k, old_j[r], old_e[r], tk[r].idx, pred_idx + row_id[r] * k, dtp, dfp,
                                          dfn);
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads) cov_fold_kernel(double *Ef, double *dEf, int64_t m)
{
    int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (j < m) {
        Ef[j] *= dEf[j];
        dEf[j] = 1.0;
    }
}

// ---- CSR batch: one warp per row, candidates = stored labels ----------------------------------------
// Rows with up to 32 * BC_RMAX stored labels are staged in registers: all loads of the row are issued up
// front, the coefficient gathers follow in one wave, and everything after that (which stored entries
// are currently selected, the probability of a label that leaves, the label / probability of a new
// selection) is resolved with shuffles instead of dependent loads and binary searches.  Per row that is
// 3 dependent memory round trips (row bounds -> entries -> coefficients) instead of ~12.
constexpr int BC_RMAX = 4;

// gain of one stored entry: affine coefficient pairs (METRIC < 0) or the record form of Jaccard / G-mean /
// H-mean (unselected: float32 difference formula, selected: float64 reference expression on the state)
template <typename T, int METRIC>
struct CsrGain {
    const float2 *coef_n, *coef_s;
    const float4 *rec;
    const double *tp, *fp, *fn;
    xc_metric_params p;
    __device__ __forceinline__ float operator()(bool sel, int j, T ev) const
    {
        if (METRIC < 0) {
            const float2 cf = __ldg((sel ? coef_s : coef_n) + j);
            return fmaf(cf.x, (float)ev, cf.y);
        }
        if (sel) return (float)bca_selected_gain(p, tp[j], fp[j], fn[j], (double)ev, (double)(T)((T)1 - ev));
        const auto xf = XfRecord<(METRIC < 0 ? XC_METRIC_JACCARD : METRIC)>::make(rec, p);
        return xf.gain(__ldg(rec + j), (float)ev);
    }
};

template <typename T, int METRIC>
__global__ void __launch_bounds__(kThreads, METRIC < 0 ? 6 : 1)
bca_batch_csr_kernel(const T *__restrict__ data, const int32_t *__restrict__ indices,
                     const int64_t *__restrict__ indptr, const int32_t *__restrict__ rows, int64_t n_rows, int k,
                     CsrGain<T, METRIC> gain, int32_t *__restrict__ pred_idx, double *dtp, double *dfp, double *dfn)
{
    const int lane = lane_id();
    const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    const T one = (T)1;
    for (int64_t w = warp; w < n_rows; w += nwarps) {
        const int64_t row = rows ? (int64_t)rows[w] : w;
        const int64_t s = indptr[row], e = indptr[row + 1];
        int32_t *pred_row = pred_idx + row * k;
        int old_j = -1;
        if (lane < k) old_j = pred_row[lane];
        int new_j;
        T new_e, old_e = (T)0;
        bool old_found = false;
        if (e - s <= 32 * BC_RMAX) {
            // ---- register-staged row
            const int nz = (int)(e - s);
            int idx[BC_RMAX];
            T val[BC_RMAX];
#pragma unroll
            for (int t = 0; t < BC_RMAX; ++t) {
                const int q = lane + 32 * t;
                idx[t] = q < nz ? indices[s + q] : -2;
                val[t] = q < nz ? data[s + q] : (T)0;
            }
            bool sel[BC_RMAX];
#pragma unroll
            for (int t = 0; t < BC_RMAX; ++t) sel[t] = false;
            for (int x = 0; x < k; ++x) {
                const int px = __shfl_sync(XC_FULL, old_j, x);
#pragma unroll
                for (int t = 0; t < BC_RMAX; ++t) {
                    const bool hit = px >= 0 && idx[t] == px;
                    sel[t] |= hit;
                    const unsigned bal = __ballot_sync(XC_FULL, hit);
                    if (bal) {   // warp-uniform: the owner hands the probability of old label x to lane x
                        const T v = __shfl_sync(XC_FULL, val[t], __ffs(bal) - 1);
                        if (lane == x) { old_e = v; old_found = true; }
                    }
                }
            }
            WarpTopK<float> tk;
            tk.init();
            float g[BC_RMAX][1];
#pragma unroll
            for (int t = 0; t < BC_RMAX; ++t) {
                g[t][0] = NAN;
                if (idx[t] >= 0) g[t][0] = gain(sel[t], idx[t], val[t]);
            }
            // the row's current selection goes in first: it sets a tight threshold, so that in a converged
            // sweep hardly any other entry reaches the list (an empty list would take all 32 lanes of the
            // first chunk through the serial insertion path: ~1500 warp instructions per row, measured as
            // the bound of this kernel)
#pragma unroll
            for (int t = 0; t < BC_RMAX; ++t) {
                float gs[1];
                gs[0] = sel[t] ? g[t][0] : NAN;
                if (32 * t < nz && __any_sync(XC_FULL, tk.passes(gs[0])))
                    xc_scan_insert<float, 1, false>(tk, gs, 32 * t, 1, k, -1);
                if (sel[t]) g[t][0] = NAN;
            }
#pragma unroll
            for (int t = 0; t < BC_RMAX; ++t)
                if (32 * t < nz && __any_sync(XC_FULL, tk.passes(g[t][0])))
                    xc_scan_insert<float, 1, false>(tk, g[t], 32 * t, 1, k, -1);
            // positions inside the row, ascending (== ascending label); fetch label / probability by shuffle
            const int src = warp_rank_src(tk.idx, k);
            const int pos = __shfl_sync(XC_FULL, tk.idx, src);
            const bool none = pos == 0x7fffffff;
            const int want_lane = none ? 0 : (pos & 31), want_t = none ? 0 : (pos >> 5);
            int gj = 0;
            T ge = (T)0;
#pragma unroll
            for (int t = 0; t < BC_RMAX; ++t) {
                const int vj = __shfl_sync(XC_FULL, idx[t], want_lane);
                const T ve = __shfl_sync(XC_FULL, val[t], want_lane);
                if (t == want_t) { gj = vj; ge = ve; }
            }
            new_j = none ? 0x7fffffff : gj;
            new_e = none ? (T)0 : ge;
        } else {
            // ---- long row: stream it, look the leaving labels up afterwards
            WarpTopK<float> tk;
            tk.init();
            for (int64_t q0 = s; q0 < e; q0 += 32) {
                int64_t q = q0 + lane;
                float g[1];
                g[0] = NAN;
                const int j = q < e ? indices[q] : -2;
                bool sel = false;
                for (int t = 0; t < k; ++t) sel |= (__shfl_sync(XC_FULL, old_j, t) == j);
                if (q < e) g[0] = gain(sel, j, data[q]);
                if (__any_sync(XC_FULL, tk.passes(g[0]))) xc_scan_insert<float, 1, false>(tk, g, q0 - s, 1, k, -1);
            }
            int src = warp_rank_src(tk.idx, k);
            int pos = __shfl_sync(XC_FULL, tk.idx, src);
            new_j = pos == 0x7fffffff ? 0x7fffffff : indices[s + pos];
            new_e = pos == 0x7fffffff ? (T)0 : data[s + pos];
            if (lane < k && old_j >= 0) {
                int64_t y = s, z = e;
                while (y < z) {
                    int64_t mid = (y + z) >> 1;
                    int v = indices[mid];
                    if (v == old_j) { old_e = data[mid]; old_found = true; break; }
                    if (v < old_j) y = mid + 1; else z = mid;
                }
            }
        }
        bool stays_old = false, stays_new = false;
        for (int t = 0; t < k; ++t) {
            int nj = __shfl_sync(XC_FULL, new_j, t);
            int oj = __shfl_sync(XC_FULL, old_j, t);
            stays_old |= (nj == old_j);
            stays_new |= (oj == new_j);
        }
        if (lane < k) {
            if (!stays_old && old_j >= 0) {
                if (old_found) {   // eta of the leaving label (a label the row does not store only had fp = 1)
                    atomicAdd(dtp + old_j, -(double)old_e);
                    atomicAdd(dfp + old_j, -(double)(T)(one - old_e));
                    atomicAdd(dfn + old_j, (double)old_e);
                } else {
                    atomicAdd(dfp + old_j, -1.0);
                }
            }
            if (!stays_new && new_j != 0x7fffffff) {
                atomicAdd(dtp + new_j, (double)new_e);
                atomicAdd(dfp + new_j, (double)(T)(one - new_e));
                atomicAdd(dfn + new_j, -(double)new_e);
            }
            pred_row[lane] = new_j == 0x7fffffff ? -1 : new_j;
        }
    }
}

// ---- CSR batch, second generation: rows of up to 128 stored labels, selection by warp reductions ------------------
// The first kernel above builds the row's top-k through the serial list insertion shared with the dense scan
// (~1500-2000 warp instructions per row, measured: the kernel was issue bound at 4.5 % of the HBM roofline).  A
// CSR row of this path is tiny -- the whole row sits in 4 registers per lane -- so the k best are taken by k
// rounds of two hardware warp reductions (REDUX.MAX on an order-preserving integer image of the gain, REDUX.MIN
// on the position for ties: lower position = lower label id, the same rule as xc_better), ~12 instructions each.
// Which stored entries are currently selected is resolved by k broadcasts; the probability of a label that
// leaves is looked up only when a label really leaves (rare once a sweep has converged).
// Labels whose deltas change are appended once to a "touched" list, so that the fold after the batch only
// recomputes the coefficients of those labels instead of all m (C4: m = 670 k, but <= 2 k B labels can change).
__device__ __forceinline__ unsigned gain_key(float g)
{
    const unsigned u = __float_as_uint(g + 0.0f);   // -0 -> +0: equal gains must have equal images
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct TouchList {
    int32_t *flag;    // [m] 0 / 1
    int32_t *list;    // [m]
    int32_t *count;   // [0] entries, [1] ticket of the fold kernel
    __device__ __forceinline__ void add(int j) const
    {
        if (flag && atomicExch(flag + j, 1) == 0) list[atomicAdd(count, 1)] = j;
    }
};

template <typename T>
__global__ void __launch_bounds__(kThreads, 6)
bca_batch_csr_redux_kernel(const T *__restrict__ data, const int32_t *__restrict__ indices,
                           const int64_t *__restrict__ indptr, const int32_t *__restrict__ rows, int64_t n_rows, int k,
                           const float2 *__restrict__ coef_n, const float2 *__restrict__ coef_s,
                           int32_t *__restrict__ pred_idx, double *dtp, double *dfp, double *dfn, TouchList touch)
{
    const int lane = lane_id();
    const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    const T one = (T)1;
    for (int64_t w = warp; w < n_rows; w += nwarps) {
        const int64_t row = rows ? (int64_t)rows[w] : w;
        const int64_t s = indptr[row], e = indptr[row + 1];
        if (e - s > 32 * BC_RMAX) continue;   // cannot happen: the launcher only takes this kernel for max_row_nnz <= 128
        const int nz = (int)(e - s);
        int32_t *pred_row = pred_idx + row * k;
        int old_j = -1;
        if (lane < k) old_j = pred_row[lane];
        int idx[BC_RMAX];
        T val[BC_RMAX];
#pragma unroll
        for (int t = 0; t < BC_RMAX; ++t) {
            const int q = lane + 32 * t;
            idx[t] = q < nz ? indices[s + q] : -2;
            val[t] = q < nz ? data[s + q] : (T)0;
        }
        unsigned selm = 0;   // bit t: my entry t is one of the row's current labels
        for (int x = 0; x < k; ++x) {
            const int px = __shfl_sync(XC_FULL, old_j, x);
#pragma unroll
            for (int t = 0; t < BC_RMAX; ++t) selm |= (px >= 0 && idx[t] == px) ? (1u << t) : 0u;
        }
        unsigned key[BC_RMAX];
#pragma unroll
        for (int t = 0; t < BC_RMAX; ++t) {
            key[t] = 0u;   // 0 = no entry / already taken (the image of any float is >= 0x007fffff)
            if (idx[t] >= 0) {
                const float2 cf = __ldg(((selm >> t) & 1u ? coef_s : coef_n) + idx[t]);
                key[t] = gain_key(fmaf(cf.x, (float)val[t], cf.y));
            }
        }
        // k rounds: best remaining entry of the row
        unsigned my_pos = 0xffffffffu;   // lane t < k: position of the t-th best
        for (int t = 0; t < k; ++t) {
            unsigned bk = key[0];
            int bs = 0;
#pragma unroll
            for (int u = 1; u < BC_RMAX; ++u)
                if (key[u] > bk) { bk = key[u]; bs = u; }
            const unsigned wmax = __reduce_max_sync(XC_FULL, bk);
            const unsigned pos = (wmax != 0u && bk == wmax) ? (unsigned)(bs * 32 + lane) : 0xffffffffu;
            const unsigned wpos = __reduce_min_sync(XC_FULL, pos);
            if (lane == t) my_pos = wpos;
            if (wpos != 0xffffffffu && (int)(wpos & 31u) == lane) {
#pragma unroll
                for (int u = 0; u < BC_RMAX; ++u)
                    if ((int)(wpos >> 5) == u) key[u] = 0u;
            }
        }
        // ascending position == ascending label: rank the k positions, then fetch label / probability by shuffle
        int rank = 0;
        for (int t = 0; t < k; ++t) {
            const unsigned o = __shfl_sync(XC_FULL, my_pos, t);
            rank += (o < my_pos) ? 1 : 0;   // positions are distinct; the "none" marker sorts last (rank by lane)
            rank += (o == my_pos && t < lane) ? 1 : 0;
        }
        unsigned pos_sorted = 0xffffffffu;
        for (int t = 0; t < k; ++t) {
            const unsigned bal = __ballot_sync(XC_FULL, lane < k && rank == t);
            const unsigned v = __shfl_sync(XC_FULL, my_pos, bal ? __ffs(bal) - 1 : 0);
            if (lane == t) pos_sorted = v;
        }
        const bool none = pos_sorted == 0xffffffffu;
        const int want_lane = none ? 0 : (int)(pos_sorted & 31u), want_t = none ? 0 : (int)(pos_sorted >> 5);
        int gj = 0;
        T ge = (T)0;
#pragma unroll
        for (int t = 0; t < BC_RMAX; ++t) {
            const int vj = __shfl_sync(XC_FULL, idx[t], want_lane);
            const T ve = __shfl_sync(XC_FULL, val[t], want_lane);
            if (t == want_t) { gj = vj; ge = ve; }
        }
        const int new_j = (none || lane >= k) ? 0x7fffffff : gj;
        const T new_e = none ? (T)0 : ge;
        bool stays_old = false, stays_new = false;
        for (int t = 0; t < k; ++t) {
            const int nj = __shfl_sync(XC_FULL, new_j, t);
            const int oj = __shfl_sync(XC_FULL, old_j, t);
            stays_old |= (nj == old_j);
            stays_new |= (oj == new_j);
        }
        const bool leaves = lane < k && !stays_old && old_j >= 0;
        const unsigned leaving = __ballot_sync(XC_FULL, leaves);
        if (leaving == 0u && __all_sync(XC_FULL, lane >= k || stays_new || new_j == 0x7fffffff)) continue;  // unchanged row
        // probability of every label that leaves (its entry is somewhere in the row's registers)
        T old_e = (T)0;
        bool old_found = false;
        for (unsigned lm = leaving; lm; lm &= lm - 1) {
            const int x = __ffs(lm) - 1;
            const int px = __shfl_sync(XC_FULL, old_j, x);
#pragma unroll
            for (int t = 0; t < BC_RMAX; ++t) {
                const unsigned bal = __ballot_sync(XC_FULL, idx[t] == px);
                if (bal) {
                    const T v = __shfl_sync(XC_FULL, val[t], __ffs(bal) - 1);
                    if (lane == x) { old_e = v; old_found = true; }
                }
            }
        }
        if (lane < k) {
            if (leaves) {
                if (old_found) {   // eta of the leaving label (a label the row does not store only had fp = 1)
                    atomicAdd(dtp + old_j, -(double)old_e);
                    atomicAdd(dfp + old_j, -(double)(T)(one - old_e));
                    atomicAdd(dfn + old_j, (double)old_e);
                } else {
                    atomicAdd(dfp + old_j, -1.0);
                }
                touch.add(old_j);
            }
            if (!stays_new && new_j != 0x7fffffff) {
                atomicAdd(dtp + new_j, (double)new_e);
                atomicAdd(dfp + new_j, (double)(T)(one - new_e));
                atomicAdd(dfn + new_j, -(double)new_e);
                touch.add(new_j);
            }
            pred_row[lane] = new_j == 0x7fffffff ? -1 : new_j;
        }
    }
}

// fold + coefficients for the touched labels only; the last block re-arms the list
__global__ void __launch_bounds__(kThreads)
bca_fold_touched_kernel(xc_metric_params p, double *tp, double *fp, double *fn, double *dtp, double *dfp, double *dfn,
                        TouchList touch, float2 *coef_n, float2 *coef_s)
{
    const int cnt = touch.count[0];
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < cnt; i += gridDim.x * kThreads) {
        const int j = touch.list[i];
        const double t = tp[j] + dtp[j], f = fp[j] + dfp[j], g = fn[j] + dfn[j];
        tp[j] = t; fp[j] = f; fn[j] = g;
        dtp[j] = 0.0; dfp[j] = 0.0; dfn[j] = 0.0;
        touch.flag[j] = 0;
        bca_coef_of(p, t, f, g, coef_n + j, coef_s + j);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(touch.count + 1, 1) == (int)gridDim.x - 1) {   // every block has read the count
            touch.count[0] = 0;
            touch.count[1] = 0;
        }
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

// Rows one warp streams at a time in the LDG kernel.  Measured on B200 at m = 13 000 (profiles/
// r01_notes.md): R = 1 -> 2.67 ms / sweep (0.91 of the measured HBM peak, 40 warps / SM), R = 2 ->
// 2.91 ms, R = 4 -> 2.85 ms: with the 8 m-byte coefficient vector resident in L1, more independent
// warps hide the list-update slow path and the per-row seed/commit latency better than register
// reuse of the coefficients does.  Wider label spaces, whose coefficients no longer fit L1, reuse
// them for 2 / 4 rows.  $XCOLUMNS_B200_DENSE_R overrides.
int dense_rows_per_warp(int64_t m)
{
    static int force_r = -1;
    if (force_r < 0) {
        const char *e = getenv("XCOLUMNS_B200_DENSE_R");
        force_r = e ? atoi(e) : 0;
    }
    if (force_r == 1 || force_r == 2 || force_r == 4) return force_r;
    if (m * 8 <= 160 * 1024) return 1;
    if (m * 8 <= 512 * 1024) return 2;
    return 4;
}

// $XCOLUMNS_B200_DENSE_DEEP=1 selects the deep-prefetch variant for sub-wave batches.  Measured on B200 at the
// 8-GPU shard shape (38 375 rows x 13 000, 8 commits, pipelined): 0.452 ms per sweep against 0.400 ms for the
// standard kernel -- four of its CTAs fill an SM's register file, so the NEXT batch's CTAs cannot move in
// while this batch drains, which costs more than the deeper loads gain.  It stays opt-in.
bool dense_deep_ok()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("XCOLUMNS_B200_DENSE_DEEP");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

// $XCOLUMNS_B200_DENSE_SMALL_CTA=0: always 256-thread CTAs
bool dense_small_cta_ok()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("XCOLUMNS_B200_DENSE_SMALL_CTA");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

// 0 = auto (LDG kernel), 1 = force the LDG kernel, 2 = force the TMA-ring kernel
int dense_path_override()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("XCOLUMNS_B200_DENSE_PATH");
        v = !e ? 0 : (e[0] == 'l' ? 1 : (e[0] == 't' ? 2 : 0));
    }
    return v;
}

int launch_batch_dense_tma(xc_ctx *ctx, const float *eta, int64_t m, int64_t ld, const int32_t *rows, int64_t n_rows,
                           int k, const float *coef_n, const float *coef_s, int32_t *pred_idx, double *dtp,
                           double *dfp, double *dfn, cudaStream_t st)
{
    static bool attr_set = false;
    if (!attr_set) {
        XC_CUDA_TRY(ctx, cudaFuncSetAttribute(bca_batch_dense_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              TMA_SMEM_BYTES));
        attr_set = true;
    }
    int64_t groups = (n_rows + TMA_ROWS - 1) / TMA_ROWS;
    int grid = (int)(groups < ctx->sm_count ? groups : ctx->sm_count);
    bca_batch_dense_tma_kernel<<<grid, TMA_THREADS, TMA_SMEM_BYTES, st>>>(
        eta, m, ld, rows, n_rows, k, (const float2 *)coef_n, (const float2 *)coef_s, pred_idx, dtp, dfp, dfn);
    XC_LAUNCHED(ctx);
    return XC_OK;
}

template <typename TE>
int launch_batch_dense(xc_ctx *ctx, const void *eta, int64_t m, int64_t ld, const int32_t *rows, int64_t n_rows, int k,
                       const float *coef_n, const float *coef_s, int32_t *pred_idx, double *dtp, double *dfp,
                       double *dfn, cudaStream_t st, int32_t *snap = nullptr)
{
    constexpr int V = 16 / sizeof(TE);
    bool vec_ok = xc_aligned16(eta) && (ld % V == 0) && xc_aligned16(coef_n);
    if (sizeof(TE) == 4 && vec_ok && dense_path_override() != 1 && !snap) {
        // the coefficient tile is copied in whole TMA_TC units: needs the padded coefficient
        // arrays the host shim allocates (xc_bca_coef_len)
        // Measured (profiles/r01_notes.md): the TMA ring keeps 200 KB / SM in flight but its 8
        // lock-stepped consumer warps expose the per-group seed/commit latency: 3.05 ms / sweep vs
        // 2.67 ms for the 40-warp LDG kernel.  It therefore stays opt-in.
        if (dense_path_override() == 2)
            return launch_batch_dense_tma(ctx, (const float *)eta, m, ld, rows, n_rows, k, coef_n, coef_s, pred_idx,
                                          dtp, dfp, dfn, st);
    }
#define XC_GO(R)                                                                                              \
    {                                                                                                         \
        auto kern = bca_batch_dense_kernel<TE, R>;                                                            \
        int grid = grid_for(ctx, kern, (n_rows + R - 1) / R);                                                 \
        int threads = kThreads;                                                                               \
        if (R == 1 && dense_small_cta_ok() && (int64_t)grid * (kThreads / 32) >= n_rows) {                    \
            /* one row per warp (a batch of at most one wave): a CTA holds its registers until its slowest */ \
            /* warp is done, so 2-warp CTAs hand the SM's slots to the next batch's rows sooner than       */ \
            /* 8-warp CTAs (rows differ in how often they touch their top-k list)                          */ \
            threads = 64;                                                                                     \
            grid = (int)((n_rows + 1) / 2);                                                                   \
        }                                                                                                     \
        kern<<<grid, threads, 0, st>>>((const TE *)eta, m, ld, rows, n_rows, k, (const float2 *)coef_n,       \
                                       (const float2 *)coef_s, pred_idx, dtp, dfp, dfn, vec_ok, snap);        \
    }
    const int rr = dense_rows_per_warp(m);
    if (rr == 4) XC_GO(4)
    else if (rr == 2) XC_GO(2)
    else if (sizeof(TE) == 4 && vec_ok && dense_deep_ok() &&
             n_rows * 4 <= (int64_t)ctx->sm_count * 6 * (kThreads / 32) * 3) {
        // fewer rows than 3/4 of a wave of the 6-CTA kernel: the launch cannot fill the GPU, so each warp keeps
        // twice the bytes in flight instead (4 CTAs per SM of the deep variant hold a 2/3-wave batch at once)
        auto kern = bca_batch_dense_kernel<TE, 1, true>;
        int grid = grid_for(ctx, kern, n_rows);
        kern<<<grid, kThreads, 0, st>>>((const TE *)eta, m, ld, rows, n_rows, k, (const float2 *)coef_n,
                                        (const float2 *)coef_s, pred_idx, dtp, dfp, dfn, vec_ok, snap);
    }
    else XC_GO(1)
#undef XC_GO
    XC_LAUNCHED(ctx);
    return XC_OK;
}

}  // namespace

extern "C" int64_t xc_bca_coef_len(int64_t m) { return ((m + TMA_TC - 1) / TMA_TC) * TMA_TC; }

extern "C" int xc_bca_wave_rows(xc_ctx *ctx, int dtype, int64_t m)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx) return XC_ERR_INVALID;
    if (dtype != XC_F32 && dtype != XC_F64) return XC_ERR_UNSUPPORTED;
    if (dtype == XC_F32 && dense_path_override() == 2) return ctx->sm_count * TMA_ROWS;
    const int rr = dense_rows_per_warp(m);
    int per_sm = 0;
#define XC_OCC(TE, R) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bca_batch_dense_kernel<TE, R>, kThreads, 0)
    if (dtype == XC_F32) { if (rr == 4) XC_OCC(float, 4); else if (rr == 2) XC_OCC(float, 2); else XC_OCC(float, 1); }
    else { if (rr == 4) XC_OCC(double, 4); else if (rr == 2) XC_OCC(double, 2); else XC_OCC(double, 1); }
#undef XC_OCC
    if (per_sm < 1) per_sm = 1;
    return ctx->sm_count * per_sm * (kThreads / 32) * rr;
}

extern "C" int xc_bca_coef(xc_ctx *ctx, const xc_metric_params *p, double *tp, double *fp, double *fn, double *dtp,
                           double *dfp, double *dfn, int64_t m, float *coef_n, float *coef_s, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !p || !tp || !fp || !fn || !coef_n || !coef_s || m <= 0) return XC_ERR_INVALID;
    if ((dtp || dfp || dfn) && !(dtp && dfp && dfn)) return XC_ERR_INVALID;
    if (p->metric != XC_METRIC_PRECISION && p->metric != XC_METRIC_RECALL && p->metric != XC_METRIC_FBETA &&
        p->metric != XC_METRIC_BALANCED_ACC && p->metric != XC_METRIC_PREC_AT_K)
        return XC_ERR_UNSUPPORTED;  // gain not affine in eta
    bca_coef_kernel<<<(unsigned)((m + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        *p, tp, fp, fn, dtp, dfp, dfn, m, (float2 *)coef_n, (float2 *)coef_s);
    XC_LAUNCHED(ctx);
    return XC_OK;
}

extern "C" int xc_bca_rec(xc_ctx *ctx, const xc_metric_params *p, double *tp, double *fp, double *fn, double *dtp,
                          double *dfp, double *dfn, int64_t m, float *rec, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !p || !tp || !fp || !fn || !rec || m <= 0) return XC_ERR_INVALID;
    if ((dtp || dfp || dfn) && !(dtp && dfp && dfn)) return XC_ERR_INVALID;
    if (p->metric != XC_METRIC_JACCARD && p->metric != XC_METRIC_GMEAN && p->metric != XC_METRIC_HMEAN)
        return XC_ERR_UNSUPPORTED;
    if (p->metric != XC_METRIC_JACCARD && p->skip_tn) return XC_ERR_INVALID;  // G-mean / H-mean need the real tn
    bca_rec_kernel<<<(unsigned)((m + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        *p, tp, fp, fn, dtp, dfp, dfn, m, (float4 *)rec);
    XC_LAUNCHED(ctx);
    return XC_OK;
}

static int batch_dense_rec_impl(xc_ctx *ctx, const xc_metric_params *p, const void *eta, int dtype, int64_t m,
                                int64_t ld, const int32_t *rows, int64_t n_rows, int k, const float *rec,
                                const double *tp, const double *fp, const double *fn, int32_t *pred_idx, double *dtp,
                                double *dfp, double *dfn, void *stream, int32_t *snap)
{
    if (!ctx || !p || !eta || !rec || !tp || !fp || !fn || !pred_idx || !dtp || !dfp || !dfn || m <= 0 || ld < m ||
        n_rows < 0)
        return XC_ERR_INVALID;
    if (k < 1 || k > 32 || k > m) return XC_ERR_INVALID;
    if (p->metric != XC_METRIC_JACCARD && p->metric != XC_METRIC_GMEAN && p->metric != XC_METRIC_HMEAN)
        return XC_ERR_UNSUPPORTED;
    if (n_rows == 0) return XC_OK;
    cudaStream_t st = (cudaStream_t)stream;
#define XC_GO(TE, METRIC)                                                                                        \
    {                                                                                                            \
        constexpr int V = 16 / sizeof(TE);                                                                       \
        const bool vec_ok = xc_aligned16(eta) && (ld % V == 0);                                                  \
        auto kern = bca_batch_dense_rec_kernel<TE, METRIC>;                                                      \
        int grid = grid_for(ctx, kern, n_rows);                                                                  \
        kern<<<grid, kThreads, 0, st>>>(*p, (const TE *)eta, m, ld, rows, n_rows, k, (const float4 *)rec, tp, fp, \
                                        fn, pred_idx, dtp, dfp, dfn, vec_ok, snap);                              \
    }
    if (dtype == XC_F32) {
        if (p->metric == XC_METRIC_JACCARD) XC_GO(float, XC_METRIC_JACCARD)
        else if (p->metric == XC_METRIC_GMEAN) XC_GO(float, XC_METRIC_GMEAN)
        else XC_GO(float, XC_METRIC_HMEAN)
    } else if (dtype == XC_F64) {
        if (p->metric == XC_METRIC_JACCARD) XC_GO(double, XC_METRIC_JACCARD)
        else if (p->metric == XC_METRIC_GMEAN) XC_GO(double, XC_METRIC_GMEAN)
        else XC_GO(double, XC_METRIC_HMEAN)
    } else {
        return XC_ERR_UNSUPPORTED;
    }
#undef XC_GO
    XC_LAUNCHED(ctx);
    return XC_OK;
}

extern "C" int xc_bca_batch_dense_rec(xc_ctx *ctx, const xc_metric_params *p, const void *eta, int dtype, int64_t m,
                                      int64_t ld, const int32_t *rows, int64_t n_rows, int k, const float *rec,
                                      const double *tp, const double *fp, const double *fn, int32_t *pred_idx,
                                      double *dtp, double *dfp, double *dfn, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    return batch_dense_rec_impl(ctx, p, eta, dtype, m, ld, rows, n_rows, k, rec, tp, fp, fn, pred_idx, dtp, dfp, dfn, stream,
                                nullptr);
}

extern "C" int64_t xc_bca_delta_stride(int64_t m) { return ((3 * m * 8 + 255) / 256) * 256; }

static int batch_dense_impl(xc_ctx *ctx, const void *eta, int dtype, int64_t m, int64_t ld, const int32_t *rows,
                            int64_t n_rows, int k, const float *coef_n, const float *coef_s, int32_t *pred_idx,
                            double *dtp, double *dfp, double *dfn, void *stream, int32_t *snap)
{
    if (!ctx || !eta || !coef_n || !coef_s || !pred_idx || !dtp || !dfp || !dfn || m <= 0 || ld < m || n_rows < 0)
        return XC_ERR_INVALID;
    if (k < 1 || k > 32 || k > m) return XC_ERR_INVALID;
    if (n_rows == 0) return XC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == XC_F32)
        return launch_batch_dense<float>(ctx, eta, m, ld, rows, n_rows, k, coef_n, coef_s, pred_idx, dtp, dfp, dfn, st, snap);
    if (dtype == XC_F64)
        return launch_batch_dense<double>(ctx, eta, m, ld, rows, n_rows, k, coef_n, coef_s, pred_idx, dtp, dfp, dfn, st, snap);
    return XC_ERR_UNSUPPORTED;
}

extern "C" int xc_bca_batch_dense(xc_ctx *ctx, const void *eta, int dtype, int64_t m, int64_t ld, const int32_t *rows,
                                  int64_t n_rows, int k, const float *coef_n, const float *coef_s, int32_t *pred_idx,
                                  double *dtp, double *dfp, double *dfn, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    return batch_dense_impl(ctx, eta, dtype, m, ld, rows, n_rows, k, coef_n, coef_s, pred_idx, dtp, dfp, dfn, stream, nullptr);
}

namespace {
template <typename T, int METRIC>
int launch_batch_csr(xc_ctx *ctx, const xc_metric_params *p, const void *data, const int32_t *indices,
                     const int64_t *indptr, const int32_t *rows, int64_t n_rows, int k, const float *coef_n,
                     const float *coef_s, const float *rec, const double *tp, const double *fp, const double *fn,
                     int32_t *pred_idx, double *dtp, double *dfp, double *dfn, cudaStream_t st)
{
    auto kern = bca_batch_csr_kernel<T, METRIC>;
    int grid = grid_for(ctx, kern, n_rows);
    CsrGain<T, METRIC> gain{(const float2 *)coef_n, (const float2 *)coef_s, (const float4 *)rec, tp, fp, fn,
                            p ? *p : xc_metric_params{}};
    kern<<<grid, kThreads, 0, st>>>((const T *)data, indices, indptr, rows, n_rows, k, gain, pred_idx, dtp, dfp, dfn);
    XC_LAUNCHED(ctx);
    return XC_OK;
}
}  // namespace

extern "C" int xc_bca_batch_csr(xc_ctx *ctx, const void *data, int dtype, const int32_t *indices,
                                const int64_t *indptr, const int32_t *rows, int64_t n_rows, int k,
                                const float *coef_n, const float *coef_s, int32_t *pred_idx, double *dtp, double *dfp,
                                double *dfn, int max_row_nnz, int32_t *touch_flag, int32_t *touch_list,
                                int32_t *touch_ctl, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !indptr || !coef_n || !coef_s || !pred_idx || !dtp || !dfp || !dfn || n_rows < 0) return XC_ERR_INVALID;
    if (k < 1 || k > 32) return XC_ERR_INVALID;
    if ((touch_flag || touch_list || touch_ctl) && !(touch_flag && touch_list && touch_ctl)) return XC_ERR_INVALID;
    if (n_rows == 0) return XC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (max_row_nnz > 0 && max_row_nnz <= 32 * BC_RMAX && (dtype == XC_F32 || dtype == XC_F64)) {
        // every row fits the registers of one warp: selection by warp reductions, touched-label list
        TouchList touch{touch_flag, touch_list, touch_ctl};
        if (dtype == XC_F32) {
            auto kern = bca_batch_csr_redux_kernel<float>;
            kern<<<grid_for(ctx, kern, n_rows), kThreads, 0, st>>>((const float *)data, indices, indptr, rows, n_rows, k,
                                                                   (const float2 *)coef_n, (const float2 *)coef_s,
                                                                   pred_idx, dtp, dfp, dfn, touch);
        } else {
            auto kern = bca_batch_csr_redux_kernel<double>;
            kern<<<grid_for(ctx, kern, n_rows), kThreads, 0, st>>>((const double *)data, indices, indptr, rows, n_rows, k,
                                                                    (const float2 *)coef_n, (const float2 *)coef_s,
                                                                    pred_idx, dtp, dfp, dfn, touch);
        }
        XC_LAUNCHED(ctx);
        return XC_OK;
    }
    if (touch_flag) return XC_ERR_UNSUPPORTED;   // the streaming kernel does not maintain the touched list
    if (dtype == XC_F32)
        return launch_batch_csr<float, -1>(ctx, nullptr, data, indices, indptr, rows, n_rows, k, coef_n, coef_s, nullptr,
                                           nullptr, nullptr, nullptr, pred_idx, dtp, dfp, dfn, st);
    if (dtype == XC_F64)
        return launch_batch_csr<double, -1>(ctx, nullptr, data, indices, indptr, rows, n_rows, k, coef_n, coef_s, nullptr,
                                            nullptr, nullptr, nullptr, pred_idx, dtp, dfp, dfn, st);
    return XC_ERR_UNSUPPORTED;
}

// fold of the pending deltas + coefficients of the labels on the touched list only (the list is emptied)
extern "C" int xc_bca_fold_touched(xc_ctx *ctx, const xc_metric_params *p, double *tp, double *fp, double *fn,
                                   double *dtp, double *dfp, double *dfn, int32_t *touch_flag, int32_t *touch_list,
                                   int32_t *touch_ctl, float *coef_n, float *coef_s, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !p || !tp || !fp || !fn || !dtp || !dfp || !dfn || !touch_flag || !touch_list || !touch_ctl || !coef_n ||
        !coef_s)
        return XC_ERR_INVALID;
    if (p->metric != XC_METRIC_PRECISION && p->metric != XC_METRIC_RECALL && p->metric != XC_METRIC_FBETA &&
        p->metric != XC_METRIC_BALANCED_ACC && p->metric != XC_METRIC_PREC_AT_K)
        return XC_ERR_UNSUPPORTED;
    TouchList touch{touch_flag, touch_list, touch_ctl};
    bca_fold_touched_kernel<<<ctx->sm_count, kThreads, 0, (cudaStream_t)stream>>>(*p, tp, fp, fn, dtp, dfp, dfn, touch,
                                                                                  (float2 *)coef_n, (float2 *)coef_s);
    XC_LAUNCHED(ctx);
    return XC_OK;
}

extern "C" int xc_bca_batch_csr_rec(xc_ctx *ctx, const xc_metric_params *p, const void *data, int dtype,
                                    const int32_t *indices, const int64_t *indptr, const int32_t *rows, int64_t n_rows,
                                    int k, const float *rec, const double *tp, const double *fp, const double *fn,
                                    int32_t *pred_idx, double *dtp, double *dfp, double *dfn, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !p || !indptr || !rec || !tp || !fp || !fn || !pred_idx || !dtp || !dfp || !dfn || n_rows < 0)
        return XC_ERR_INVALID;
    if (k < 1 || k > 32) return XC_ERR_INVALID;
    if (p->metric != XC_METRIC_JACCARD && p->metric != XC_METRIC_GMEAN && p->metric != XC_METRIC_HMEAN)
        return XC_ERR_UNSUPPORTED;
    if (n_rows == 0) return XC_OK;
    cudaStream_t st = (cudaStream_t)stream;
#define XC_GO(T, METRIC)                                                                                          \
    return launch_batch_csr<T, METRIC>(ctx, p, data, indices, indptr, rows, n_rows, k, nullptr, nullptr, rec, tp, fp, \
                                       fn, pred_idx, dtp, dfp, dfn, st)
    if (dtype == XC_F32) {
        if (p->metric == XC_METRIC_JACCARD) XC_GO(float, XC_METRIC_JACCARD);
        if (p->metric == XC_METRIC_GMEAN) XC_GO(float, XC_METRIC_GMEAN);
        XC_GO(float, XC_METRIC_HMEAN);
    }
    if (dtype == XC_F64) {
        if (p->metric == XC_METRIC_JACCARD) XC_GO(double, XC_METRIC_JACCARD);
        if (p->metric == XC_METRIC_GMEAN) XC_GO(double, XC_METRIC_GMEAN);
        XC_GO(double, XC_METRIC_HMEAN);
    }
#undef XC_GO
    return XC_ERR_UNSUPPORTED;
}

extern "C" int xc_cov_batch_csr(xc_ctx *ctx, const void *data, int dtype, const int32_t *indices,
                                const int64_t *indptr, const int32_t *rows, int64_t n_rows, int k, double alpha,
                                const double *Ef, int32_t *pred_idx, double *dEf, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !indptr || !Ef || !pred_idx || !dEf || n_rows < 0) return XC_ERR_INVALID;
    if (k < 1 || k > 32) return XC_ERR_INVALID;
    if (n_rows == 0) return XC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == XC_F32) {
        auto kern = cov_batch_csr_kernel<float>;
        int grid = grid_for(ctx, kern, n_rows);
        kern<<<grid, kThreads, 0, st>>>((const float *)data, indices, indptr, rows, n_rows, k, alpha, Ef, pred_idx, dEf);
    } else if (dtype == XC_F64) {
        auto kern = cov_batch_csr_kernel<double>;
        int grid = grid_for(ctx, kern, n_rows);
        kern<<<grid, kThreads, 0, st>>>((const double *)data, indices, indptr, rows, n_rows, k, alpha, Ef, pred_idx, dEf);
    } else {
        return XC_ERR_UNSUPPORTED;
    }
    XC_LAUNCHED(ctx);
    return XC_OK;
}

extern "C" int xc_cov_batch_dense(xc_ctx *ctx, const void *eta, int dtype, int64_t m, int64_t ld, const int32_t *rows,
                                  int64_t n_rows, int k, double alpha, const double *Ef, int32_t *pred_idx,
                                  double *dEf, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !eta || !Ef || !pred_idx || !dEf || m <= 0 || ld < m || n_rows < 0) return XC_ERR_INVALID;
    if (k < 1 || k > 32 || k > m) return XC_ERR_INVALID;
    if (n_rows == 0) return XC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == XC_F32) {
        bool vec_ok = xc_aligned16(eta) && (ld % 4 == 0);
        auto kern = cov_batch_dense_kernel<float, 1>;
        int grid = grid_for(ctx, kern, n_rows);
        kern<<<grid, kThreads, 0, st>>>((const float *)eta, m, ld, rows, n_rows, k, alpha, Ef, pred_idx, dEf, vec_ok);
    } else if (dtype == XC_F64) {
        bool vec_ok = xc_aligned16(eta) && (ld % 2 == 0);
        auto kern = cov_batch_dense_kernel<double, 1>;
        int grid = grid_for(ctx, kern, n_rows);
        kern<<<grid, kThreads, 0, st>>>((const double *)eta, m, ld, rows, n_rows, k, alpha, Ef, pred_idx, dEf, vec_ok);
    } else {
        return XC_ERR_UNSUPPORTED;
    }
    XC_LAUNCHED(ctx);
    return XC_OK;
}

extern "C" int xc_cov_fold(xc_ctx *ctx, double *Ef, double *dEf, int64_t m, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !Ef || !dEf || m <= 0) return XC_ERR_INVALID;
    cov_fold_kernel<<<(unsigned)((m + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(Ef, dEf, m);
    XC_LAUNCHED(ctx);
    return XC_OK;
}

// ---- one whole batched sweep as ONE host call (single process: no exchange between the batches) ----------
// For small problems the sweep is bound by the host, not the GPU: at 10 000 x 1 000 the ~20 calls of a sweep
// cost ~0.5 ms through the Python shim while the kernels need ~0.06 ms.  This entry point issues the same
// launches (coefficients / records with the fold of the pending deltas, batch kernel, ... , final fold) from
// C; `rec` selects the Jaccard / G-mean / H-mean record kernels.  order: the sweep's visiting order (device).
extern "C" int xc_bca_sweep_dense(xc_ctx *ctx, const xc_metric_params *p, const void *eta, int dtype, int64_t m,
                                  int64_t ld, const int32_t *order, int64_t n_order, int64_t batch, int k, float *coef_a,
                                  float *coef_s, int32_t *pred_idx, double *tp, double *fp, double *fn, double *dtp,
                                  double *dfp, double *dfn, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !p || !order || n_order < 0 || batch < 1) return XC_ERR_INVALID;
    const bool rec = p->metric == XC_METRIC_JACCARD || p->metric == XC_METRIC_GMEAN || p->metric == XC_METRIC_HMEAN;
    for (int64_t lo = 0; lo <= n_order; lo += batch) {   // the last trip (lo >= n_order or empty) only folds
        int rc = rec ? xc_bca_rec(ctx, p, tp, fp, fn, dtp, dfp, dfn, m, coef_a, stream)
                     : xc_bca_coef(ctx, p, tp, fp, fn, dtp, dfp, dfn, m, coef_a, coef_s, stream);
        if (rc) return rc;
        const int64_t hi = lo + batch < n_order ? lo + batch : n_order;
        if (hi <= lo) break;
        rc = rec ? xc_bca_batch_dense_rec(ctx, p, eta, dtype, m, ld, order + lo, hi - lo, k, coef_a, tp, fp, fn, pred_idx,
                                          dtp, dfp, dfn, stream)
                 : xc_bca_batch_dense(ctx, eta, dtype, m, ld, order + lo, hi - lo, k, coef_a, coef_s, pred_idx, dtp,
                                      dfp, dfn, stream);
        if (rc) return rc;
        if (hi == n_order) {   // fold the last batch
            rc = rec ? xc_bca_rec(ctx, p, tp, fp, fn, dtp, dfp, dfn, m, coef_a, stream)
                     : xc_bca_coef(ctx, p, tp, fp, fn, dtp, dfp, dfn, m, coef_a, coef_s, stream);
            return rc;
        }
    }
    return XC_OK;
}

extern "C" int xc_bca_sweep_csr(xc_ctx *ctx, const xc_metric_params *p, const void *data, int dtype,
                                const int32_t *indices, const int64_t *indptr, int64_t m, const int32_t *order,
                                int64_t n_order, int64_t batch, int k, float *coef_n, float *coef_s, int32_t *pred_idx,
                                double *tp, double *fp, double *fn, double *dtp, double *dfp, double *dfn,
                                int max_row_nnz, int32_t *touch_flag, int32_t *touch_list, int32_t *touch_ctl,
                                int refresh, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !p || !order || n_order < 0 || batch < 1) return XC_ERR_INVALID;
    const bool rec = p->metric == XC_METRIC_JACCARD || p->metric == XC_METRIC_GMEAN || p->metric == XC_METRIC_HMEAN;
    // touched-label lists: rows that fit one warp's registers, affine-gain metrics; the fold after a batch then
    // only visits the labels the batch changed instead of all m
    const bool lists = !rec && touch_flag && touch_list && touch_ctl && max_row_nnz > 0 && max_row_nnz <= 32 * BC_RMAX;
    auto fold_all = [&]() {   // coef_n is the record array for the record metrics
        return rec ? xc_bca_rec(ctx, p, tp, fp, fn, dtp, dfp, dfn, m, coef_n, stream)
                   : xc_bca_coef(ctx, p, tp, fp, fn, dtp, dfp, dfn, m, coef_n, coef_s, stream);
    };
    auto fold = [&]() {
        return lists ? xc_bca_fold_touched(ctx, p, tp, fp, fn, dtp, dfp, dfn, touch_flag, touch_list, touch_ctl, coef_n,
                                           coef_s, stream)
                     : fold_all();
    };
    if (lists && refresh) {   // the host changed the state: every coefficient from scratch (pending deltas are zero)
        int rc = fold_all();
        if (rc) return rc;
    }
    for (int64_t lo = 0; lo <= n_order; lo += batch) {
        int rc = XC_OK;
        if (!lists) rc = fold_all();
        if (rc) return rc;
        const int64_t hi = lo + batch < n_order ? lo + batch : n_order;
        if (hi <= lo) break;
        rc = rec ? xc_bca_batch_csr_rec(ctx, p, data, dtype, indices, indptr, order + lo, hi - lo, k, coef_n, tp, fp, fn,
                                        pred_idx, dtp, dfp, dfn, stream)
                 : xc_bca_batch_csr(ctx, data, dtype, indices, indptr, order + lo, hi - lo, k, coef_n, coef_s, pred_idx, dtp,
                                    dfp, dfn, lists ? max_row_nnz : 0, lists ? touch_flag : nullptr,
                                    lists ? touch_list : nullptr, lists ? touch_ctl : nullptr, stream);
        if (rc) return rc;
        if (lists) {
            rc = fold();
            if (rc) return rc;
        }
        if (hi == n_order) return lists ? XC_OK : fold_all();
    }
    return XC_OK;
}

// ---- pipelined sweeps: batches overlap, commits are applied `lag` batches late ---------------------------------
// The strict block-Jacobi order  K_0, commit_0, K_1, commit_1, ...  leaves the GPU under-filled at every batch
// boundary: the tail of the batch kernel, the commit (a cross-GPU barrier when the rows are sharded), the ramp of
// the next kernel -- and a batch smaller than one wave of resident warps (strong scaling: 38 k rows per GPU cut
// into 8 commits = 0.68 wave) cannot saturate HBM by itself.  With lag = 1 batch g (global index, counted over
// all sweeps) is evaluated against the state after commit g-2: K_g and K_{g+1} have no dependency on each other,
// so they run on two streams, and the chain
//     K_g -> commit_g -> K_{g+2}                      (stream g & 1, coefficient set g & 1)
//     commit_{g-1} -> commit_g                        (event: the state is folded in batch order)
// keeps two batch kernels in flight; the commit of one overlaps the streaming of the other (the commit kernel is
// sized to fit next to six resident streaming CTAs).  The pipeline does NOT drain between sweeps: a sweep's
// prologue (its visiting order, a snapshot of the prediction for roll-backs) only waits for the previous sweep's
// last two batch KERNELS (a row must not be visited by two sweeps at once), not for their commits, and the
// utility of the finished sweep is evaluated behind its last commit while the next sweep already streams.  The
// caller's stream waits for that utility only.  Staleness grows from "up to one batch" to "one to two batches";
// measured against the sequential reference this moves the final macro-F1 by < 1e-6 (scripts/probe_lag.py,
// tests/test_gpu_baseline_shapes.py).  lag = 0 is the strict order on the caller's stream.
//
// Delta buffers: batch g accumulates into buffer g % NB, NB = 2 (lag + 1); commit_g folds buffer g % NB and
// clears buffer (g + lag + 1) % NB, the one batch g + lag + 1 will use -- with sharded rows that buffer was last
// read by the peers during commit g - lag - 1, which every peer has left before it raised its flag for commit g.
// No host-side clearing, no race with a slower peer.
namespace {

// ---- no drain between sweeps ------------------------------------------------------------------------------------
// When a new sweep starts, the previous sweep's last lag batches may still be streaming.  A row must not be visited
// by two sweeps at once, so the new sweep's first batches must not contain rows of those batches.  Instead of
// waiting for them (a drain: the tail of a batch kernel is a whole row time, ~50 us, every sweep), the new visiting
// order is repaired: within its first `window` positions the rows that are still "busy" are moved behind the
// others (stable partition), which puts them beyond the first lag batches -- behind at least one commit of the new
// sweep, by which time the old sweep's kernels have retired.  Only which rows share a batch matters to a
// block-Jacobi sweep, not the order inside a batch, and the order stays a permutation.
// The raw permutation is generated into a scratch buffer; two small kernels write the sweep's order from it:
// order_count_kernel counts the free rows of every 256-entry block of the window, order_fix_kernel copies the
// positions beyond the window, partitions the window stably (free rows first, then the busy ones: every block
// adds up the counts of the blocks before it -- deterministic, the same order on every run) and stamps the rows
// of THIS sweep's last `lag` batches with the sweep's mark for the next sweep's test (every row is handled by
// exactly one thread, the only one to read and write its stamp).
__global__ void __launch_bounds__(256)
order_count_kernel(const int32_t *__restrict__ raw, int64_t window, const int32_t *__restrict__ stamp,
                   int32_t prev_mark, int32_t *__restrict__ counts)
{
    __shared__ int s_cnt[8];
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const bool free_row = i < window && stamp[raw[i]] != prev_mark;
    const unsigned mf = __ballot_sync(XC_FULL, free_row);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = __popc(mf);
    __syncthreads();
    if (threadIdx.x == 0) {
        int c = 0;
        for (int w = 0; w < 8; ++w) c += s_cnt[w];
        counts[blockIdx.x] = c;
    }
}

__global__ void __launch_bounds__(256)
order_fix_kernel(const int32_t *__restrict__ raw, int32_t *__restrict__ order, int64_t n, int64_t window,
                 int32_t *__restrict__ stamp, int32_t prev_mark, int64_t tail_from, int32_t mark,
                 const int32_t *__restrict__ counts)
{
    __shared__ int s_red[8], s_red2[8];
    __shared__ int s_before, s_total, s_wfree[8];
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool valid = i < n;
    const int32_t r = valid ? raw[i] : 0;
    int64_t pos = i;
    if ((int64_t)blockIdx.x * 256 < window) {   // block-uniform: this block holds window positions
        const int nblk = (int)((window + 255) / 256);
        int before = 0, total = 0;
        for (int b = threadIdx.x; b < nblk; b += 256) {
            const int c = counts[b];
            total += c;
            if (b < (int)blockIdx.x) before += c;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            before += __shfl_xor_sync(XC_FULL, before, o);
            total += __shfl_xor_sync(XC_FULL, total, o);
        }
        const bool inw = valid && i < window;
        const bool busy = inw && stamp[r] == prev_mark;
        const unsigned mf = __ballot_sync(XC_FULL, inw && !busy);
        if (lane == 0) { s_red[wid] = before; s_red2[wid] = total; s_wfree[wid] = __popc(mf); }
        __syncthreads();
        if (threadIdx.x == 0) {
            int bsum = 0, tsum = 0;
            for (int w = 0; w < 8; ++w) { bsum += s_red[w]; tsum += s_red2[w]; }
            s_before = bsum;
            s_total = tsum;
        }
        __syncthreads();
        int wfree_before = 0;
        for (int w = 0; w < wid; ++w) wfree_before += s_wfree[w];
        const int free_rank = s_before + wfree_before + __popc(mf & ((1u << lane) - 1u));   // free rows before me
        if (inw) pos = busy ? (int64_t)s_total + (i - free_rank) : free_rank;
    }
    if (valid) {
        order[pos] = r;
        if (pos >= tail_from) stamp[r] = mark;
    }
}

struct PipeCommit {
    xc_ctx *ctx;
    xc_p2p *w;
    const xc_metric_params *p;
    double *tp, *fp, *fn;
    double *local;                 // this rank's NB delta buffers
    int64_t win_off, inbox_off;    // byte offsets inside a window: own buffers, inbox (world x NB buffers)
    int64_t stride, m;
    int nb, rec;
};

// my deltas of buffer `cur` into every peer's inbox, then my flag (sharded rows only); see the call site for the
// stream it is issued on
int launch_push(const PipeCommit &c, int cur, cudaStream_t st)
{
    if (!c.w) return XC_OK;
    const int64_t want = (3 * c.m + kCommitThreads - 1) / kCommitThreads;
    const int grid = (int)(want < c.ctx->sm_count ? want : c.ctx->sm_count);
    const unsigned epoch = ++c.w->epoch;
    bca_push_kernel<<<grid, kCommitThreads, 0, st>>>(c.w->windows_dev, c.w->world, c.w->rank, epoch, c.win_off, c.inbox_off,
                                                     c.stride, c.nb, cur, c.m);
    XC_LAUNCHED(c.ctx);
    return XC_OK;
}

int launch_commit(const PipeCommit &c, int cur, int clr, float *set_a, float *set_b, int64_t clen, cudaStream_t st)
{
    const int64_t want = (c.m + kCommitThreads - 1) / kCommitThreads;
    const int grid = (int)(want < c.ctx->sm_count ? want : c.ctx->sm_count);
    float2 *an = (float2 *)set_a, *as = c.rec ? nullptr : (float2 *)(set_a + 2 * clen);
    float2 *bn = set_b ? (float2 *)set_b : nullptr, *bs = set_b ? (float2 *)(set_b + 2 * clen) : nullptr;
    uint8_t *const *windows = c.w ? c.w->windows_dev : nullptr;
    const int world = c.w ? c.w->world : 1, rank = c.w ? c.w->rank : 0;
    const unsigned epoch = (c.w && cur >= 0) ? c.w->epoch : 0u;   // the epoch launch_push raised the flags with
    bca_commit_kernel<<<grid, kCommitThreads, 0, st>>>(*c.p, c.tp, c.fp, c.fn, windows, world, rank, epoch, c.local,
                                                       c.inbox_off, c.stride, c.nb, cur, clr, c.m, an, as, bn, bs, c.rec);
    XC_LAUNCHED(c.ctx);
    return XC_OK;
}

}  // namespace

static_assert(sizeof(xc_bca_pipe_args) == 200, "xc_bca_pipe_args layout (mirrored by ctypes in _lib.py)");

// payload bytes of a peer window for the pipelined sweep: NB own delta buffers + world x NB inbox buffers
extern "C" int64_t xc_bca_window_bytes(int64_t m, int lag, int world)
{
    const int64_t nb = 2 * ((lag < 0 ? 0 : (lag > XC_PIPE_MAX_LAG ? XC_PIPE_MAX_LAG : lag)) + 1);
    return (nb + (int64_t)(world < 1 ? 1 : world) * nb) * xc_bca_delta_stride(m);
}

extern "C" int xc_bca_pipe_buffers(int lag) { return 2 * ((lag < 0 ? 0 : (lag > XC_PIPE_MAX_LAG ? XC_PIPE_MAX_LAG : lag)) + 1); }

// the caller's stream waits for everything the pipeline still has in flight (before the host reads or rewrites
// the prediction / the state: recompute, roll-back, end of the call)
extern "C" int xc_bca_pipe_join(xc_ctx *ctx, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx) return XC_ERR_INVALID;
    if (!ctx->pipe_active) return XC_OK;
    if (ctx->pipe_forked) {
        for (int i = 0; i <= XC_PIPE_MAX_LAG + 1; ++i) {
            XC_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_join[i], i <= XC_PIPE_MAX_LAG ? ctx->aux[i] : ctx->cstream));
            XC_CUDA_TRY(ctx, cudaStreamWaitEvent((cudaStream_t)stream, ctx->ev_join[i], 0));
        }
    }
    ctx->pipe_active = ctx->pipe_forked = false;
    return XC_OK;
}

extern "C" int xc_bca_pipe_sweep(xc_ctx *ctx, xc_p2p *w, const xc_bca_pipe_args *a, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !a) return XC_ERR_INVALID;
    const xc_metric_params *p = a->params;
    const int64_t m = a->m, n_order = a->n_rows;
    if (!p || !a->eta || !a->coef || !a->pred_idx || !a->tp || !a->fp || !a->fn || n_order < 0 || a->batch < 1 ||
        a->n_batches < 0 || a->batch0 < 0 || m <= 0 || a->ld < m)
        return XC_ERR_INVALID;
    if (n_order > 0 && !a->order) return XC_ERR_INVALID;
    if (!w && !a->delta) return XC_ERR_INVALID;
    if (w && !w->opened) return XC_ERR_INVALID;
    if (a->k < 1 || a->k > 32 || a->k > m) return XC_ERR_INVALID;
    if (p->metric < 0 || p->metric > XC_METRIC_PREC_AT_K) return XC_ERR_INVALID;
    static_assert(2 * (XC_PIPE_MAX_LAG + 1) <= XC_P2P_MAX_BUF, "one flag row / ticket per delta buffer");
    if (a->n_batches * a->batch < n_order) return XC_ERR_INVALID;
    if (a->util_out && !a->util_params) return XC_ERR_INVALID;
    const bool rec = p->metric == XC_METRIC_JACCARD || p->metric == XC_METRIC_GMEAN || p->metric == XC_METRIC_HMEAN;
    if (rec && p->metric != XC_METRIC_JACCARD && p->skip_tn) return XC_ERR_INVALID;
    int lag = a->lag < 0 ? 0 : (a->lag > XC_PIPE_MAX_LAG ? XC_PIPE_MAX_LAG : a->lag);
    if (rec) lag = 0;   // the record kernels read the float64 state of the selected labels: no overlap with a commit
    const int S = lag + 1, NB = 2 * S;   // S batches in flight on S streams, S coefficient sets, 2 S delta buffers
    const int64_t stride = xc_bca_delta_stride(m), clen = xc_bca_coef_len(m);
    if (w && (size_t)(XC_P2P_HEADER + xc_bca_window_bytes(m, lag, w->world)) > w->bytes) return XC_ERR_INVALID;
    cudaStream_t caller = (cudaStream_t)stream;
    double *local = w ? reinterpret_cast<double *>(w->windows[w->rank] + XC_P2P_HEADER) : a->delta;
    PipeCommit c{ctx, w, p, a->tp, a->fp, a->fn, local, (int64_t)XC_P2P_HEADER, (int64_t)XC_P2P_HEADER + NB * stride,
                 stride, m, NB, rec ? 1 : 0};
    float *set[XC_PIPE_MAX_LAG + 1];
    for (int i = 0; i <= XC_PIPE_MAX_LAG; ++i) set[i] = a->coef + (int64_t)(i < S ? i : 0) * 4 * clen;
    // $XCOLUMNS_B200_PIPE_SERIAL=1 (tests): the same dependency order on ONE stream, nothing overlaps
    const bool serial = getenv("XCOLUMNS_B200_PIPE_SERIAL") && atoi(getenv("XCOLUMNS_B200_PIPE_SERIAL")) == 1;
    const bool forked = S > 1 && !serial;
    int rc;
    // fresh: the caller hands over a state the commits of earlier calls did not produce (first call, after a join,
    // XC_PIPE_FORK); otherwise the sweep continues on the coefficient sets the previous sweep's commits left behind
    // (same dependency structure whether the batches overlap on two streams or are serialised on one)
    const bool fresh = !ctx->pipe_active || (a->flags & XC_PIPE_FORK) || (forked != ctx->pipe_forked);
    if (fresh) {
        if (ctx->pipe_active) {
            rc = xc_bca_pipe_join(ctx, stream);
            if (rc) return rc;
        }
        // coefficients of the state the caller hands over, for every set
        for (int i = 0; i < S; i += 2) {
            rc = launch_commit(c, -1, -1, set[i], i + 1 < S ? set[i + 1] : nullptr, clen, caller);
            if (rc) return rc;
        }
    }
    // batch kernels: one low-priority stream per batch in flight; commits (and the sweep's utility): ONE
    // high-priority stream -- they are serial anyway (the state is folded in batch order), and the priority lets a
    // commit's few small CTAs be dispatched ahead of the pending CTAs of the batch kernels queued behind it (the
    // hardware otherwise hands out CTAs grid by grid in launch order: measured, a commit waited a whole row time)
    cudaStream_t st[XC_PIPE_MAX_LAG + 1], cst = caller;
    for (int i = 0; i <= XC_PIPE_MAX_LAG; ++i) st[i] = caller;
    if (forked) {
        rc = xc_ctx_aux_streams(ctx);
        if (rc) return rc;
        for (int i = 0; i < S; ++i) st[i] = ctx->aux[i];
        cst = ctx->cstream;
        if (fresh) {
            XC_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, caller));
            for (int i = 0; i < S; ++i) XC_CUDA_TRY(ctx, cudaStreamWaitEvent(st[i], ctx->ev_fork, 0));
            XC_CUDA_TRY(ctx, cudaStreamWaitEvent(cst, ctx->ev_fork, 0));
        }
    }
    ctx->pipe_active = true;
    ctx->pipe_forked = forked;
    // ---- prologue on the stream of the sweep's first batch: the visiting order.  The previous sweep's last `lag`
    // batch kernels may still run: either the new order is repaired so that its first batches avoid their rows
    // (no drain, see order_fix_kernel), or this sweep's kernels wait for them.
    const int s0 = (int)(a->batch0 % S);
    const bool shuffle = (a->flags & XC_PIPE_SHUFFLE) != 0;
    int32_t *order = a->order + (shuffle ? (a->sweep & 1) * n_order : 0);
    // where this sweep's last `lag` batches begin in its order
    const int64_t my_tail = a->n_batches > lag ? (a->n_batches - lag) * a->batch : 0;
    bool repair = false;
    if (!fresh && lag > 0) {   // (also in the serialised test schedule: the order is part of the algorithm)
        const int64_t head = (int64_t)lag * a->batch, tail = a->prev_tail_from >= 0 ? n_order - a->prev_tail_from : -1;
        repair = shuffle && tail >= 0 && head + tail <= my_tail && my_tail <= n_order && a->n_batches >= S;
        if (!repair && forked)
            for (int i = 0; i < S; ++i)
                if (i != s0) XC_CUDA_TRY(ctx, cudaStreamWaitEvent(st[s0], ctx->ev_k[i], 0));
    }
    if (shuffle && n_order > 0) {
        // layout of a->order: [order A | order B | raw permutation | stamps | per-block counts]
        int32_t *raw = a->order + 2 * n_order, *stamp = a->order + 3 * n_order, *counts = a->order + 4 * n_order;
        rc = xc_permutation(ctx, n_order, a->seed, raw, st[s0]);
        if (rc) return rc;
        const int64_t window = repair ? (int64_t)lag * a->batch + (n_order - a->prev_tail_from) : 0;
        const int32_t prev_mark = (int32_t)(a->sweep & 0x3fffffff), mark = (int32_t)((a->sweep + 1) & 0x3fffffff);
        if (window > 0) {
            order_count_kernel<<<(unsigned)((window + 255) / 256), 256, 0, st[s0]>>>(raw, window, stamp, prev_mark, counts);
            XC_LAUNCHED(ctx);
        }
        order_fix_kernel<<<(unsigned)((n_order + 255) / 256), 256, 0, st[s0]>>>(
            raw, order, n_order, window, stamp, prev_mark, my_tail < n_order ? my_tail : n_order, mark, counts);
        XC_LAUNCHED(ctx);
    }
    if (forked) {
        XC_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_pro, st[s0]));
        for (int i = 0; i < S; ++i)
            if (i != s0) XC_CUDA_TRY(ctx, cudaStreamWaitEvent(st[i], ctx->ev_pro, 0));
    }
    // ---- batches
    for (int64_t b = 0; b < a->n_batches; ++b) {
        const int64_t g = a->batch0 + b;
        const int si = (int)(g % S);
        const int cur = (int)(g % NB), clr = (int)((g + S) % NB);
        const int64_t lo = b * a->batch < n_order ? b * a->batch : n_order;
        const int64_t hi = lo + a->batch < n_order ? lo + a->batch : n_order;
        double *d = reinterpret_cast<double *>(reinterpret_cast<uint8_t *>(local) + (int64_t)cur * stride);
        if (hi > lo) {
            cudaEvent_t e0 = nullptr, e1 = nullptr;
            if (ctx->timing_on) {
                rc = xc_timing_slot(ctx, hi - lo, &e0, &e1);
                if (rc) return rc;
                XC_CUDA_TRY(ctx, cudaEventRecord(e0, st[si]));
            }
            rc = rec ? batch_dense_rec_impl(ctx, p, a->eta, a->dtype, m, a->ld, order + lo, hi - lo, a->k, set[si], a->tp,
                                            a->fp, a->fn, a->pred_idx, d, d + m, d + 2 * m, st[si], a->pred_snapshot)
                     : batch_dense_impl(ctx, a->eta, a->dtype, m, a->ld, order + lo, hi - lo, a->k, set[si],
                                        set[si] + 2 * clen, a->pred_idx, d, d + m, d + 2 * m, st[si], a->pred_snapshot);
            if (rc) return rc;
            if (e1) XC_CUDA_TRY(ctx, cudaEventRecord(e1, st[si]));
        }
        if (forked) {
            // commit_g follows K_g (and, on its own stream, commit_{g-1}).  Sharded rows: the batch's deltas are
            // published as soon as the batch is done, by a push on a HIGH-priority stream of its own -- a push depends on
            // nothing but its batch, so it neither queues behind the (serial) commits of earlier batches nor, as a
            // low-priority launch on the batch stream would, behind the pending CTAs of the batch kernels already in
            // the hardware queue (measured at 8 GPUs: 36 us per commit in the first case, pushes delayed by a whole
            // row time in the second; profiles/r02_notes.md section 2)
            XC_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_k[si], st[si]));
            if (w) {
                XC_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->pstream[si], ctx->ev_k[si], 0));
                rc = launch_push(c, cur, ctx->pstream[si]);
                if (rc) return rc;
                XC_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_p[si], ctx->pstream[si]));
                XC_CUDA_TRY(ctx, cudaStreamWaitEvent(cst, ctx->ev_p[si], 0));
            } else {
                XC_CUDA_TRY(ctx, cudaStreamWaitEvent(cst, ctx->ev_k[si], 0));
            }
        } else {
            rc = launch_push(c, cur, st[si]);
            if (rc) return rc;
        }
        cudaEvent_t c0 = nullptr, c1 = nullptr;
        if (ctx->timing_on && ctx->timing_commits) {   // (diagnostics: rows = 0 marks a commit in the timing log)
            rc = xc_timing_slot(ctx, 0, &c0, &c1);
            if (rc) return rc;
            XC_CUDA_TRY(ctx, cudaEventRecord(c0, cst));
        }
        rc = launch_commit(c, cur, clr, set[si], nullptr, clen, cst);
        if (rc) return rc;
        if (c1) XC_CUDA_TRY(ctx, cudaEventRecord(c1, cst));
        if (forked) {   // K_{g+S}, the next kernel on this batch stream, follows commit_g
            XC_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_c[si], cst));
            XC_CUDA_TRY(ctx, cudaStreamWaitEvent(st[si], ctx->ev_c[si], 0));
        }
        if (b + 1 == a->n_batches && a->util_out) {
            // the finished sweep's utility, behind its last commit; the next sweep's first commit queues behind it on
            // the commit stream, its batch kernels do not wait for it
            rc = xc_utility_launch(ctx, a->util_params, a->agg, a->tp, a->fp, a->fn, nullptr, a->util_tn_rows, m,
                                   a->util_out, cst);
            if (rc) return rc;
        }
    }
    if (a->n_batches == 0 && a->util_out) {
        rc = xc_utility_launch(ctx, a->util_params, a->agg, a->tp, a->fp, a->fn, nullptr, a->util_tn_rows, m, a->util_out,
                               cst);
        if (rc) return rc;
    }
    if (forked) {   // the caller's stream sees the utility (and with it the state after the last commit), nothing else
        XC_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_util, cst));
        XC_CUDA_TRY(ctx, cudaStreamWaitEvent(caller, ctx->ev_util, 0));
    }
    return XC_OK;
}

// ---- coverage: whole sweeps, dense state, utility on the device ------------------------------------------------
// ref: block_coordinate.py:600-701.  Ef_j = prod_i (1 - yhat_ij eta_ij); a batch accumulates multiplicative
// factors into dEf (1 = unchanged) which the fold multiplies into Ef.  With the rows sharded over several ranks
// the host shim all-reduces dEf (product) between the batch kernel and the fold, so these one-call sweeps are the
// single-process form.
namespace {

template <typename T>
__global__ void __launch_bounds__(256)
cov_state_dense_kernel(const T *__restrict__ eta, int64_t ld, int64_t n, const int32_t *__restrict__ pred_idx, int k,
                       double *Ef)
{
    const T one = (T)1;
    const int64_t total = n * k;
    for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
        const int pj = pred_idx[t];
        if (pj < 0) continue;
        const T e = eta[(t / k) * ld + pj];
        if (e != (T)0) atomic_mul(Ef + pj, (double)(T)((double)one * (1.0 - (double)e)));   // zeros are not stored entries
    }
}

__global__ void __launch_bounds__(256) cov_fill_kernel(double *x, double v, int64_t m)
{
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j < m) x[j] = v;
}

// out[0] = 1 - mean(Ef)   (block_coordinate.py:592), deterministic block-ordered sum
__global__ void __launch_bounds__(256)
cov_utility_kernel(const double *__restrict__ Ef, int64_t m, double *out, double *partials, unsigned *counter)
{
    __shared__ double sm[8];
    double s = 0.0;
    for (int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x; j < m; j += (int64_t)gridDim.x * 256) s += Ef[j];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    double bsum = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < 8; ++w) bsum += sm[w];
    double total;
    if (xc_grid_sum_last(bsum, partials, counter, &total)) *out = 1.0 - total / (double)m;
}

}  // namespace

extern "C" int xc_cov_state_dense(xc_ctx *ctx, const void *eta, int dtype, int64_t n, int64_t m, int64_t ld,
                                  const int32_t *pred_idx, int k, double *Ef, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !eta || !pred_idx || !Ef || n <= 0 || m <= 0 || ld < m || k < 1 || k > 32) return XC_ERR_INVALID;
    if (dtype != XC_F32 && dtype != XC_F64) return XC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    cov_fill_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(Ef, 1.0, m);
    XC_LAUNCHED(ctx);
    const int64_t blocks = (n * k + 255) / 256, cap = (int64_t)ctx->sm_count * 16;
    const int grid = (int)(blocks < cap ? blocks : cap);
    if (dtype == XC_F32) cov_state_dense_kernel<float><<<grid, 256, 0, st>>>((const float *)eta, ld, n, pred_idx, k, Ef);
    else cov_state_dense_kernel<double><<<grid, 256, 0, st>>>((const double *)eta, ld, n, pred_idx, k, Ef);
    XC_LAUNCHED(ctx);
    return XC_OK;
}

extern "C" int xc_cov_utility(xc_ctx *ctx, const double *Ef, int64_t m, double *out_dev, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !Ef || !out_dev || m <= 0) return XC_ERR_INVALID;
    int64_t blocks = (m + 255) / 256;
    const int grid = (int)(blocks < XC_RED_MAX_BLOCKS ? blocks : XC_RED_MAX_BLOCKS);
    cov_utility_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(Ef, m, out_dev, ctx->red_partials, ctx->red_counter);
    XC_LAUNCHED(ctx);
    return XC_OK;
}

extern "C" int xc_cov_sweep_csr(xc_ctx *ctx, const void *data, int dtype, const int32_t *indices, const int64_t *indptr,
                                int64_t m, const int32_t *order, int64_t n_order, int64_t batch, int k, double alpha,
                                double *Ef, int32_t *pred_idx, double *dEf, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !order || n_order < 0 || batch < 1 || m <= 0) return XC_ERR_INVALID;
    for (int64_t lo = 0; lo < n_order; lo += batch) {
        const int64_t hi = lo + batch < n_order ? lo + batch : n_order;
        int rc = xc_cov_batch_csr(ctx, data, dtype, indices, indptr, order + lo, hi - lo, k, alpha, Ef, pred_idx, dEf, stream);
        if (rc) return rc;
        rc = xc_cov_fold(ctx, Ef, dEf, m, stream);
        if (rc) return rc;
    }
    return XC_OK;
}

extern "C" int xc_cov_sweep_dense(xc_ctx *ctx, const void *eta, int dtype, int64_t m, int64_t ld, const int32_t *order,
                                  int64_t n_order, int64_t batch, int k, double alpha, double *Ef, int32_t *pred_idx,
                                  double *dEf, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !order || n_order < 0 || batch < 1 || m <= 0) return XC_ERR_INVALID;
    for (int64_t lo = 0; lo < n_order; lo += batch) {
        const int64_t hi = lo + batch < n_order ? lo + batch : n_order;
        int rc = xc_cov_batch_dense(ctx, eta, dtype, m, ld, order + lo, hi - lo, k, alpha, Ef, pred_idx, dEf, stream);
        if (rc) return rc;
        rc = xc_cov_fold(ctx, Ef, dEf, m, stream);
        if (rc) return rc;
    }
    return XC_OK;
}
