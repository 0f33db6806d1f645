// Label-wise confusion sums (tp / fp / fn), column sums and the utility reduction.
// Replaces xcolumns/confusion_matrix.py:160-234, :364-399 and
// numba_csr_functions.py:116-258 (sorted-index merges + scatter-add).
//
// Two summation orders (include/xcolumns_b200.h): XC_SUM_FAST splits the rows over the grid and
// combines float64 partial sums with atomics; XC_SUM_ORDERED keeps one running sum per label and
// visits the rows in order, which reproduces numpy's axis-0 reduction / numba's row loop bit for
// bit (used by the sequential-exact BCA at sweep boundaries).
#include "xc_scan.cuh"

namespace {

constexpr int kThreads = 256;

template <typename T>
__device__ __forceinline__ T ldx(const T *p) { return __ldg(p); }

// ---- dense x dense, axis 0 -------------------------------------------------------------------
// thread = one column; blockIdx.y = row chunk.  ACC = double (atomic combine when chunks > 1) or
// float (ordered only, mimics dtype=None on float32 input).
template <typename T, typename ACC>
__global__ void __launch_bounds__(kThreads)
confmat_dense_ax0_kernel(const T *__restrict__ yt, int64_t ldt, const T *__restrict__ yp, int64_t ldp, int64_t n,
                         int64_t m, int64_t rows_per_chunk, double *tp, double *fp, double *fn, bool atomic)
{
    const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (j >= m) return;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
    const int64_t r1 = min(n, r0 + rows_per_chunk);
    ACC stp = 0, sfp = 0, sfn = 0;
    const T one = (T)1;
#pragma unroll 4
    for (int64_t i = r0; i < r1; ++i) {
        T y = ldx(yt + i * ldt + j), p = ldx(yp + i * ldp + j);
        stp = stp + (ACC)(T)(y * p);
        sfp = sfp + (ACC)(T)((one - y) * p);
        sfn = sfn + (ACC)(T)(y * (one - p));
    }
    if (atomic) {
        atomicAdd(tp + j, (double)stp);
        atomicAdd(fp + j, (double)sfp);
        atomicAdd(fn + j, (double)sfn);
    } else {
        tp[j] = (double)stp;
        fp[j] = (double)sfp;
        fn[j] = (double)sfn;
    }
}

// ---- dense x dense, axis 1: one warp per row ---------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
confmat_dense_ax1_kernel(const T *__restrict__ yt, int64_t ldt, const T *__restrict__ yp, int64_t ldp, int64_t n,
                         int64_t m, double *tp, double *fp, double *fn)
{
    const int lane = lane_id();
    const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    const T one = (T)1;
    for (int64_t i = warp; i < n; i += nwarps) {
        double stp = 0, sfp = 0, sfn = 0;
        for (int64_t j = lane; j < m; j += 32) {
            T y = ldx(yt + i * ldt + j), p = ldx(yp + i * ldp + j);
            stp += (double)(T)(y * p);
            sfp += (double)(T)((one - y) * p);
            sfn += (double)(T)(y * (one - p));
        }
        stp = warp_sum(stp);
        sfp = warp_sum(sfp);
        sfn = warp_sum(sfn);
        if (lane == 0) {
            tp[i] = stp;
            fp[i] = sfp;
            fn[i] = sfn;
        }
    }
}

// ---- column sums of a dense matrix (fast order) ------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
colsum_dense_kernel(const T *__restrict__ x, int64_t ld, int64_t n, int64_t m, int64_t rows_per_chunk,
                    double *out, bool vec_ok)
{
    constexpr int V = XcVec<T>::V;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
    const int64_t r1 = min(n, r0 + rows_per_chunk);
    if (vec_ok) {
        const int64_t j = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * V;
        if (j >= m) return;
        if (j + V <= m) {
            double s[V];
#pragma unroll
            for (int v = 0; v < V; ++v) s[v] = 0.0;
            int64_t i = r0;
            for (; i + 4 <= r1; i += 4) {
                T e[4][V];
#pragma unroll
                for (int u = 0; u < 4; ++u) XcVec<T>::load(x + (i + u) * ld + j, e[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int v = 0; v < V; ++v) s[v] += (double)e[u][v];
            }
            for (; i < r1; ++i) {
                T e[V];
                XcVec<T>::load(x + i * ld + j, e);
#pragma unroll
                for (int v = 0; v < V; ++v) s[v] += (double)e[v];
            }
#pragma unroll
            for (int v = 0; v < V; ++v) atomicAdd(out + j + v, s[v]);
        } else {
            for (int64_t jj = j; jj < m; ++jj) {
                double s = 0.0;
                for (int64_t i = r0; i < r1; ++i) s += (double)ld_stream(x + i * ld + jj);
                atomicAdd(out + jj, s);
            }
        }
    } else {
        const int64_t j0 = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * V;
        for (int64_t jj = j0; jj < min(m, j0 + V); ++jj) {
            double s = 0.0;
            for (int64_t i = r0; i < r1; ++i) s += (double)ld_stream(x + i * ld + jj);
            atomicAdd(out + jj, s);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
colsum_csr_kernel(const T *__restrict__ data, const int32_t *__restrict__ indices, int64_t nnz, double *out)
{
    for (int64_t q = (int64_t)blockIdx.x * kThreads + threadIdx.x; q < nnz; q += (int64_t)gridDim.x * kThreads)
        atomicAdd(out + indices[q], (double)data[q]);
}

// ---- dense truth, compact prediction ------------------------------------------------------------
// Warp w handles 32 consecutive rows of ONE prediction slot.  Frequent labels repeat inside such a
// warp (head labels are predicted for a large share of the rows), so equal labels are combined
// with __match_any_sync and only the group leader issues the float64 atomics: the hot addresses
// see up to 32x fewer atomics (measured 244 us -> see profiles/ for n*k = 1.5M entries).
template <typename T>
__global__ void __launch_bounds__(kThreads)
confmat_compact_fast_kernel(const T *__restrict__ yt, int64_t ld, const int32_t *__restrict__ pred, int k,
                            int64_t n, double *tp, double *fp)
{
    const int lane = lane_id();
    const T one = (T)1;
    const int64_t n32 = (n + 31) / 32;           // warps per slot
    const int64_t total_warps = n32 * k;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    for (int64_t w = warp0; w < total_warps; w += (int64_t)gridDim.x * (kThreads / 32)) {
        const int slot = (int)(w / n32);
        const int64_t i = (w - (int64_t)slot * n32) * 32 + lane;
        int j = -1;
        double vt = 0.0, vf = 0.0;
        if (i < n) {
            j = pred[i * k + slot];
            if (j >= 0) {
                T y = yt[i * ld + j];
                vt = (double)y;
                vf = (double)(T)(one - y);
            }
        }
        unsigned grp = __match_any_sync(XC_FULL, j);
        const bool leader = (__ffs(grp) - 1) == lane;
        double st = 0.0, sf = 0.0;
        unsigned rest = grp;
        while (__any_sync(XC_FULL, rest != 0)) {
            int src = rest ? __ffs(rest) - 1 : lane;
            double ot = __shfl_sync(XC_FULL, vt, src);
            double of = __shfl_sync(XC_FULL, vf, src);
            if (rest) {
                st += ot;
                sf += of;
                rest &= rest - 1;
            }
        }
        if (leader && j >= 0) {
            atomicAdd(tp + j, st);
            atomicAdd(fp + j, sf);
        }
    }
}

__global__ void __launch_bounds__(kThreads) sub_kernel(const double *a, const double *b, double *out, int64_t m)
{
    int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (j < m) out[j] = a[j] - b[j];
}

// ordered: thread = one label, rows in order; pred rows (k ids) are warp-broadcast loads
template <typename T>
__global__ void __launch_bounds__(kThreads)
confmat_compact_ordered_kernel(const T *__restrict__ yt, int64_t ld, const int32_t *__restrict__ pred, int k,
                               int64_t n, int64_t m, double *tp, double *fp, double *fn)
{
    const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (j >= m) return;
    double stp = 0.0, sfp = 0.0, sfn = 0.0;
    const T one = (T)1;
    const int jj = (int)j;
    // The three running sums are chains of dependent float64 adds in row order (that order IS the contract), but the
    // loads are not: U rows are fetched before their adds are issued, so the L2 latency of a row is paid once per U
    // rows instead of once per two (a label's thread walks all n rows; with m = 1 000 only four CTAs exist and
    // nothing else hides the latency -- this pass follows every sweep of the sequential mode).
    constexpr int U = 8;
    int64_t i = 0;
    for (; i + U <= n; i += U) {
        T y[U];
        bool sel[U];
#pragma unroll
        for (int u = 0; u < U; ++u) y[u] = ldx(yt + (i + u) * ld + j);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            sel[u] = false;
            for (int t = 0; t < k; ++t) sel[u] |= (__ldg(pred + (i + u) * k + t) == jj);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (sel[u]) {
                stp += (double)y[u];
                sfp += (double)(T)(one - y[u]);
            } else {
                sfn += (double)y[u];
            }
        }
    }
    for (; i < n; ++i) {
        T y = ldx(yt + i * ld + j);
        bool sel = false;
        for (int t = 0; t < k; ++t) sel |= (__ldg(pred + i * k + t) == jj);
        if (sel) {
            stp += (double)y;
            sfp += (double)(T)(one - y);
        } else {
            sfn += (double)y;
        }
    }
    tp[j] = stp;
    fp[j] = sfp;
    fn[j] = sfn;
}

// Same sums, same order, with the prediction as a bitmap: word (i, j / 32) holds the "row i predicts label j" bits of
// the 32 labels one warp owns, so a row costs a warp ONE broadcast load instead of k loads + k compares per lane.
// Measured motive (ncu launch list of a C1 call, profiles/r02_launches_exact_c1.csv): the pass above took 3.3 ms for
// 10 000 x 1 000 -- 650 cycles per row and thread, the k dependent load -> compare pairs of a run-time k serialise on
// an in-order warp -- which is 8 % of a sequential sweep.
__global__ void __launch_bounds__(kThreads)
pred_bitmap_kernel(const int32_t *__restrict__ pred, int64_t total, int k, int64_t words_per_row, uint32_t *bitmap)
{
    const int64_t q = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (q >= total) return;
    const int j = pred[q];
    if (j >= 0) atomicOr(bitmap + (q / k) * words_per_row + (j >> 5), 1u << (j & 31));
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
confmat_bitmap_ordered_kernel(const T *__restrict__ yt, int64_t ld, const uint32_t *__restrict__ bitmap,
                              int64_t words_per_row, int64_t n, int64_t m, double *tp, double *fp, double *fn)
{
    const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;   // kThreads is a multiple of 32: a warp's labels
    if (j >= m) return;                                                // share the word j >> 5
    double stp = 0.0, sfp = 0.0, sfn = 0.0;
    const T one = (T)1;
    const uint32_t *bw = bitmap + (j >> 5);
    const int bit = (int)(j & 31);
    constexpr int U = 8;
    int64_t i = 0;
    for (; i + U <= n; i += U) {
        T y[U];
        uint32_t w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            y[u] = ldx(yt + (i + u) * ld + j);
            w[u] = __ldg(bw + (i + u) * words_per_row);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if ((w[u] >> bit) & 1u) {
                stp += (double)y[u];
                sfp += (double)(T)(one - y[u]);
            } else {
                sfn += (double)y[u];
            }
        }
    }
    for (; i < n; ++i) {
        const T y = ldx(yt + i * ld + j);
        if ((__ldg(bw + i * words_per_row) >> bit) & 1u) {
            stp += (double)y;
            sfp += (double)(T)(one - y);
        } else {
            sfn += (double)y;
        }
    }
    tp[j] = stp;
    fp[j] = sfp;
    fn[j] = sfn;
}

// ---- CSR x CSR ------------------------------------------------------------------------------------
// products exactly as numba forms them: a*b in T; a*(1.0-b) in float64 rounded once to T
template <typename T> __device__ __forceinline__ T mul_round(T a, T b) { return (T)(a * b); }
template <typename T> __device__ __forceinline__ T mul_om_round(T a, T b)
{
    return (T)__dmul_rn((double)a, __dsub_rn(1.0, (double)b));
}

// position of label j in the sorted slice idx[s..e), or -1
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

// One row's contributions; the `add` functor is either an atomic or a plain RMW.
template <typename T, class Add>
__device__ __forceinline__ void confmat_csr_row(const T *t_data, const int32_t *t_idx, int64_t ts, int64_t te,
                                                const T *p_data, const int32_t *p_idx, int64_t ps, int64_t pe,
                                                int lane, int nlanes, double *tp, double *fp, double *fn, Add add)
{
    for (int64_t x = ps + lane; x < pe; x += nlanes) {  // over prediction entries: tp and fp
        int j = p_idx[x];
        T pv = p_data[x];
        int64_t y = csr_find(t_idx, ts, te, j);
        if (y >= 0) {
            add(tp + j, (double)mul_round(pv, t_data[y]));
            add(fp + j, (double)mul_om_round(pv, t_data[y]));
        } else {
            add(fp + j, (double)pv);
        }
    }
    for (int64_t y = ts + lane; y < te; y += nlanes) {  // over truth entries: fn
        int j = t_idx[y];
        T tv = t_data[y];
        int64_t x = csr_find(p_idx, ps, pe, j);
        add(fn + j, x >= 0 ? (double)mul_om_round(tv, p_data[x]) : (double)tv);
    }
}

struct AddAtomic { __device__ __forceinline__ void operator()(double *p, double v) const { atomicAdd(p, v); } };
struct AddPlain { __device__ __forceinline__ void operator()(double *p, double v) const { *p = *p + v; } };
// float32 running sums stored in the double output array's low precision domain (dtype=None)
struct AddPlainF32 {
    __device__ __forceinline__ void operator()(double *p, double v) const { *p = (double)((float)*p + (float)v); }
};

template <typename T>
__global__ void __launch_bounds__(kThreads)
confmat_csr_fast_kernel(const T *t_data, const int32_t *t_idx, const int64_t *t_ptr, const T *p_data,
                        const int32_t *p_idx, const int64_t *p_ptr, int64_t n, double *tp, double *fp, double *fn)
{
    const int lane = lane_id();
    const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t i = warp; i < n; i += nwarps)
        confmat_csr_row<T>(t_data, t_idx, t_ptr[i], t_ptr[i + 1], p_data, p_idx, p_ptr[i], p_ptr[i + 1], lane, 32, tp,
                           fp, fn, AddAtomic());
}

// single CTA, rows strictly in order (labels inside a row are distinct, so lanes never collide)
template <typename T, class Add>
__global__ void __launch_bounds__(128)
confmat_csr_ordered_kernel(const T *t_data, const int32_t *t_idx, const int64_t *t_ptr, const T *p_data,
                           const int32_t *p_idx, const int64_t *p_ptr, int64_t n, double *tp, double *fp,
                           double *fn)
{
    for (int64_t i = 0; i < n; ++i) {
        confmat_csr_row<T>(t_data, t_idx, t_ptr[i], t_ptr[i + 1], p_data, p_idx, p_ptr[i], p_ptr[i + 1], threadIdx.x,
                           blockDim.x, tp, fp, fn, Add());
        __syncthreads();
    }
}

// CSR truth + compact prediction of ones
template <typename T, class Add>
__device__ __forceinline__ void confmat_csrc_row(const T *t_data, const int32_t *t_idx, int64_t ts, int64_t te,
                                                 const int32_t *pred, int k, int lane, int nlanes, double *tp,
                                                 double *fp, double *fn, Add add)
{
    const T one = (T)1;
    for (int x = lane; x < k; x += nlanes) {
        int j = pred[x];
        if (j < 0) continue;
        int64_t y = csr_find(t_idx, ts, te, j);
        if (y >= 0) {
            add(tp + j, (double)mul_round(one, t_data[y]));
            add(fp + j, (double)mul_om_round(one, t_data[y]));
        } else {
            add(fp + j, 1.0);
        }
    }
    if (!fn) return;  // caller derives fn = colsum - tp (the row's unselected entries are never touched)
    for (int64_t y = ts + lane; y < te; y += nlanes) {
        int j = t_idx[y];
        bool sel = false;
        for (int x = 0; x < k; ++x) sel |= (pred[x] == j);
        add(fn + j, sel ? (double)mul_om_round(t_data[y], one) : (double)t_data[y]);
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
confmat_csrc_fast_kernel(const T *t_data, const int32_t *t_idx, const int64_t *t_ptr, const int32_t *pred, int k,
                         int64_t n, double *tp, double *fp, double *fn)
{
    const int lane = lane_id();
    const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t i = warp; i < n; i += nwarps)
        confmat_csrc_row<T>(t_data, t_idx, t_ptr[i], t_ptr[i + 1], pred + i * k, k, lane, 32, tp, fp, fn, AddAtomic());
}

template <typename T>
__global__ void __launch_bounds__(128)
confmat_csrc_ordered_kernel(const T *t_data, const int32_t *t_idx, const int64_t *t_ptr, const int32_t *pred, int k,
                            int64_t n, double *tp, double *fp, double *fn)
{
    for (int64_t i = 0; i < n; ++i) {
        confmat_csrc_row<T>(t_data, t_idx, t_ptr[i], t_ptr[i + 1], pred + i * k, k, threadIdx.x, blockDim.x, tp, fp,
                            fn, AddPlain());
        __syncthreads();
    }
}

// ---- CSR truth, compact prediction, ORDERED sums through a column-major copy -------------------------
// The row-walking ordered kernel above is one CTA doing binary searches row after row (15 us per row).
// The per-label sums only need each label's contributions in row order, which is exactly a CSC column
// of y_proba (built once per call by the host shim with a stable sort; the probabilities never change).
// One warp per label: 32 entries at a time are loaded / matched against their rows' predictions in
// parallel, then added in lane order (a 32-step shuffle chain), i.e. strictly in row order.
// Only valid when every predicted label is stored in its row (always true for predictions the sweeps
// produce); confmat_lone_check_kernel verifies it and the host falls back to the row-walking kernel.
template <typename T>
__global__ void __launch_bounds__(kThreads)
confmat_csc_ordered_kernel(const T *__restrict__ c_data, const int32_t *__restrict__ c_rows,
                           const int64_t *__restrict__ c_ptr, const int32_t *__restrict__ pred, int k, int64_t m,
                           double *tp, double *fp, double *fn)
{
    const int lane = lane_id();
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const T one = (T)1;
    for (int64_t j = warp0; j < m; j += (int64_t)gridDim.x * (kThreads / 32)) {
        const int64_t s = c_ptr[j], e = c_ptr[j + 1];
        double stp = 0.0, sfp = 0.0, sfn = 0.0;   // identical in every lane
        for (int64_t q0 = s; q0 < e; q0 += 32) {
            const int64_t q = q0 + lane;
            double ctp = 0.0, cfp = 0.0, cfn = 0.0;
            if (q < e) {
                const int64_t i = c_rows[q];
                const T v = c_data[q];
                bool sel = false;
                for (int t = 0; t < k; ++t) sel |= (__ldg(pred + i * k + t) == (int)j);
                if (sel) {
                    ctp = (double)mul_round(one, v);
                    cfp = (double)mul_om_round(one, v);
                    cfn = (double)mul_om_round(v, one);
                } else {
                    cfn = (double)v;
                }
            }
            const int cnt = (int)min((int64_t)32, e - q0);
            for (int l = 0; l < cnt; ++l) {   // strictly in row order; adding an exact 0.0 is a no-op
                stp = stp + __shfl_sync(XC_FULL, ctp, l);
                sfp = sfp + __shfl_sync(XC_FULL, cfp, l);
                sfn = sfn + __shfl_sync(XC_FULL, cfn, l);
            }
        }
        if (lane == 0) {
            tp[j] = stp;
            fp[j] = sfp;
            fn[j] = sfn;
        }
    }
}

// flag = 1 if some predicted label is not stored in its row
__global__ void __launch_bounds__(kThreads)
confmat_lone_check_kernel(const int32_t *__restrict__ t_idx, const int64_t *__restrict__ t_ptr,
                          const int32_t *__restrict__ pred, int k, int64_t n, int *flag)
{
    const int64_t total = n * k;
    for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (int64_t)gridDim.x * kThreads) {
        const int j = pred[t];
        if (j < 0) continue;
        const int64_t i = t / k;
        if (csr_find(t_idx, t_ptr[i], t_ptr[i + 1], j) < 0) *flag = 1;
    }
}

// ---- utility: mean / sum over labels of the binary metric, fixed reduction order -----------------
__global__ void __launch_bounds__(256)
utility_kernel(xc_metric_params p, int agg, const double *tp, const double *fp, const double *fn, const double *tn,
               double tn_rows, int64_t m, double *out, double *partials, unsigned *counter)
{
    __shared__ double sm[8];
    double s = 0.0;
    for (int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x; j < m; j += (int64_t)gridDim.x * 256) {
        // tn: the stored vector, or (tn_rows >= 0) -tp - fp - fn + rows formed on the fly (confusion_matrix.py:397)
        double t4 = tn_rows >= 0.0 ? ((-tp[j] - fp[j]) - fn[j]) + tn_rows : (tn ? tn[j] : -1.0);
        s += xc_metric_eval(p, tp[j] / p.n_div, fp[j] / p.n_div, fn[j] / p.n_div, t4 / p.n_div);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    double bsum = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < 8; ++w) bsum += sm[w];
    double total;
    if (xc_grid_sum_last(bsum, partials, counter, &total)) *out = agg == 0 ? total / (double)m : total;
}

inline int cap_grid(xc_ctx *ctx, int64_t blocks, int per_sm = 16)
{
    int64_t cap = (int64_t)ctx->sm_count * per_sm;
    if (blocks < 1) blocks = 1;
    return (int)(blocks < cap ? blocks : cap);
}

}  // namespace

extern "C" int xc_confmat_dense(xc_ctx *ctx, const void *y_true, int64_t ldt, const void *y_pred, int64_t ldp,
                                int dtype, int64_t n, int64_t m, int axis, int order, int acc_f32, double *tp,
                                double *fp, double *fn, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !y_true || !y_pred || !tp || !fp || !fn || n <= 0 || m <= 0 || ldt < m || ldp < m) return XC_ERR_INVALID;
    if (axis != 0 && axis != 1) return XC_ERR_INVALID;
    if (dtype != XC_F32 && dtype != XC_F64) return XC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    if (axis == 1) {
        int grid = cap_grid(ctx, (n + 7) / 8);
        if (dtype == XC_F32)
            confmat_dense_ax1_kernel<float><<<grid, kThreads, 0, st>>>((const float *)y_true, ldt, (const float *)y_pred, ldp, n, m, tp, fp, fn);
        else
            confmat_dense_ax1_kernel<double><<<grid, kThreads, 0, st>>>((const double *)y_true, ldt, (const double *)y_pred, ldp, n, m, tp, fp, fn);
        XC_LAUNCHED(ctx);
        return XC_OK;
    }
    int64_t col_blocks = (m + kThreads - 1) / kThreads;
    int64_t chunks = 1;
    if (order == XC_SUM_FAST) {
        int64_t want = ((int64_t)ctx->sm_count * 8 + col_blocks - 1) / col_blocks;
        chunks = want < 1 ? 1 : want;
        if (chunks > (n + 63) / 64) chunks = (n + 63) / 64;
        if (chunks > 65535) chunks = 65535;
    }
    int64_t rpc = (n + chunks - 1) / chunks;
    chunks = (n + rpc - 1) / rpc;
    bool atomic = chunks > 1;
    if (atomic) {
        XC_CUDA_TRY(ctx, cudaMemsetAsync(tp, 0, sizeof(double) * m, st));
        XC_CUDA_TRY(ctx, cudaMemsetAsync(fp, 0, sizeof(double) * m, st));
        XC_CUDA_TRY(ctx, cudaMemsetAsync(fn, 0, sizeof(double) * m, st));
    }
    dim3 grid((unsigned)col_blocks, (unsigned)chunks);
    if (dtype == XC_F32) {
        if (acc_f32 && !atomic)
            confmat_dense_ax0_kernel<float, float><<<grid, kThreads, 0, st>>>((const float *)y_true, ldt, (const float *)y_pred, ldp, n, m, rpc, tp, fp, fn, false);
        else
            confmat_dense_ax0_kernel<float, double><<<grid, kThreads, 0, st>>>((const float *)y_true, ldt, (const float *)y_pred, ldp, n, m, rpc, tp, fp, fn, atomic);
    } else {
        confmat_dense_ax0_kernel<double, double><<<grid, kThreads, 0, st>>>((const double *)y_true, ldt, (const double *)y_pred, ldp, n, m, rpc, tp, fp, fn, atomic);
    }
    XC_LAUNCHED(ctx);
    return XC_OK;
}

extern "C" int xc_colsum_dense(xc_ctx *ctx, const void *x, int dtype, int64_t n, int64_t m, int64_t ld, double *out,
                               void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !x || !out || n <= 0 || m <= 0 || ld < m) return XC_ERR_INVALID;
    if (dtype != XC_F32 && dtype != XC_F64) return XC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int V = dtype == XC_F32 ? 4 : 2;
    bool vec_ok = xc_aligned16(x) && (ld % V == 0);
    int64_t col_blocks = (m + (int64_t)kThreads * V - 1) / ((int64_t)kThreads * V);
    int64_t chunks = ((int64_t)ctx->sm_count * 8 + col_blocks - 1) / col_blocks;
    if (chunks > (n + 31) / 32) chunks = (n + 31) / 32;
    if (chunks < 1) chunks = 1;
    if (chunks > 65535) chunks = 65535;
    int64_t rpc = (n + chunks - 1) / chunks;
    chunks = (n + rpc - 1) / rpc;
    XC_CUDA_TRY(ctx, cudaMemsetAsync(out, 0, sizeof(double) * m, st));
    dim3 grid((unsigned)col_blocks, (unsigned)chunks);
    if (dtype == XC_F32)
        colsum_dense_kernel<float><<<grid, kThreads, 0, st>>>((const float *)x, ld, n, m, rpc, out, vec_ok);
    else
        colsum_dense_kernel<double><<<grid, kThreads, 0, st>>>((const double *)x, ld, n, m, rpc, out, vec_ok);
    XC_LAUNCHED(ctx);
    return XC_OK;
}

extern "C" int xc_colsum_csr(xc_ctx *ctx, const void *data, int dtype, const int32_t *indices, int64_t nnz, int64_t m,
                             double *out, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !out || nnz < 0 || m <= 0) return XC_ERR_INVALID;
    if (dtype != XC_F32 && dtype != XC_F64) return XC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    XC_CUDA_TRY(ctx, cudaMemsetAsync(out, 0, sizeof(double) * m, st));
    if (nnz == 0) return XC_OK;
    int grid = cap_grid(ctx, (nnz + kThreads - 1) / kThreads);
    if (dtype == XC_F32) colsum_csr_kernel<float><<<grid, kThreads, 0, st>>>((const float *)data, indices, nnz, out);
    else colsum_csr_kernel<double><<<grid, kThreads, 0, st>>>((const double *)data, indices, nnz, out);
    XC_LAUNCHED(ctx);
    return XC_OK;
}

extern "C" int xc_confmat_dense_compact(xc_ctx *ctx, const void *y_true, int dtype, int64_t ld,
                                        const int32_t *pred_idx, int k, int64_t n, int64_t m, int order,
                                        const double *colsum, double *tp, double *fp, double *fn, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !y_true || !pred_idx || !tp || !fp || !fn || n <= 0 || m <= 0 || ld < m || k < 1) return XC_ERR_INVALID;
    if (dtype != XC_F32 && dtype != XC_F64) return XC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    if (order == XC_SUM_ORDERED) {
        int grid = (int)((m + kThreads - 1) / kThreads);
        const int64_t words_per_row = (m + 31) / 32;
        const size_t bitmap_bytes = (size_t)n * (size_t)words_per_row * 4;
        if (bitmap_bytes <= ((size_t)256 << 20)) {
            void *scratch = nullptr;
            int rc = xc_ctx_scratch(ctx, bitmap_bytes, &scratch);
            if (rc) return rc;
            uint32_t *bitmap = (uint32_t *)scratch;
            XC_CUDA_TRY(ctx, cudaMemsetAsync(bitmap, 0, bitmap_bytes, st));
            const int64_t total = n * (int64_t)k;
            pred_bitmap_kernel<<<(unsigned)((total + kThreads - 1) / kThreads), kThreads, 0, st>>>(pred_idx, total, k,
                                                                                               words_per_row, bitmap);
            XC_LAUNCHED(ctx);
            if (dtype == XC_F32)
                confmat_bitmap_ordered_kernel<float><<<grid, kThreads, 0, st>>>((const float *)y_true, ld, bitmap, words_per_row, n, m, tp, fp, fn);
            else
                confmat_bitmap_ordered_kernel<double><<<grid, kThreads, 0, st>>>((const double *)y_true, ld, bitmap, words_per_row, n, m, tp, fp, fn);
            XC_LAUNCHED(ctx);
            return XC_OK;
        }
        if (dtype == XC_F32)
            confmat_compact_ordered_kernel<float><<<grid, kThreads, 0, st>>>((const float *)y_true, ld, pred_idx, k, n, m, tp, fp, fn);
        else
            confmat_compact_ordered_kernel<double><<<grid, kThreads, 0, st>>>((const double *)y_true, ld, pred_idx, k, n, m, tp, fp, fn);
        XC_LAUNCHED(ctx);
        return XC_OK;
    }
    if (!colsum) {
        int rc = xc_colsum_dense(ctx, y_true, dtype, n, m, ld, fn, stream);
        if (rc) return rc;
        colsum = fn;
    }
    XC_CUDA_TRY(ctx, cudaMemsetAsync(tp, 0, sizeof(double) * m, st));
    XC_CUDA_TRY(ctx, cudaMemsetAsync(fp, 0, sizeof(double) * m, st));
    int grid = cap_grid(ctx, (((n + 31) / 32) * k + (kThreads / 32) - 1) / (kThreads / 32));
    if (dtype == XC_F32)
        confmat_compact_fast_kernel<float><<<grid, kThreads, 0, st>>>((const float *)y_true, ld, pred_idx, k, n, tp, fp);
    else
        confmat_compact_fast_kernel<double><<<grid, kThreads, 0, st>>>((const double *)y_true, ld, pred_idx, k, n, tp, fp);
    XC_LAUNCHED(ctx);
    sub_kernel<<<(unsigned)((m + kThreads - 1) / kThreads), kThreads, 0, st>>>(colsum, tp, fn, m);
    XC_LAUNCHED(ctx);
    return XC_OK;
}

extern "C" int xc_confmat_csr(xc_ctx *ctx, const void *t_data, const int32_t *t_idx, const int64_t *t_ptr,
                              const void *p_data, const int32_t *p_idx, const int64_t *p_ptr, int dtype, int64_t n,
                              int64_t m, int order, int acc_f32, double *tp, double *fp, double *fn, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !t_ptr || !p_ptr || !tp || !fp || !fn || n <= 0 || m <= 0) return XC_ERR_INVALID;
    if (dtype != XC_F32 && dtype != XC_F64) return XC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    XC_CUDA_TRY(ctx, cudaMemsetAsync(tp, 0, sizeof(double) * m, st));
    XC_CUDA_TRY(ctx, cudaMemsetAsync(fp, 0, sizeof(double) * m, st));
    XC_CUDA_TRY(ctx, cudaMemsetAsync(fn, 0, sizeof(double) * m, st));
#define XC_ARGS(T) (const T *)t_data, t_idx, t_ptr, (const T *)p_data, p_idx, p_ptr, n, tp, fp, fn
    if (order == XC_SUM_ORDERED) {
        if (dtype == XC_F32 && acc_f32) confmat_csr_ordered_kernel<float, AddPlainF32><<<1, 128, 0, st>>>(XC_ARGS(float));
        else if (dtype == XC_F32) confmat_csr_ordered_kernel<float, AddPlain><<<1, 128, 0, st>>>(XC_ARGS(float));
        else confmat_csr_ordered_kernel<double, AddPlain><<<1, 128, 0, st>>>(XC_ARGS(double));
    } else {
        int grid = cap_grid(ctx, (n + 7) / 8);
        if (dtype == XC_F32) confmat_csr_fast_kernel<float><<<grid, kThreads, 0, st>>>(XC_ARGS(float));
        else confmat_csr_fast_kernel<double><<<grid, kThreads, 0, st>>>(XC_ARGS(double));
    }
#undef XC_ARGS
    XC_LAUNCHED(ctx);
    return XC_OK;
}

extern "C" int xc_confmat_csr_compact(xc_ctx *ctx, const void *t_data, const int32_t *t_idx, const int64_t *t_ptr,
                                      int dtype, const int32_t *pred_idx, int k, int64_t n, int64_t m, int order,
                                      double *tp, double *fp, double *fn, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !t_ptr || !pred_idx || !tp || !fp || n <= 0 || m <= 0 || k < 1) return XC_ERR_INVALID;
    if (!fn && order == XC_SUM_ORDERED) return XC_ERR_INVALID;  // fn may only be skipped in the fast order
    if (dtype != XC_F32 && dtype != XC_F64) return XC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    XC_CUDA_TRY(ctx, cudaMemsetAsync(tp, 0, sizeof(double) * m, st));
    XC_CUDA_TRY(ctx, cudaMemsetAsync(fp, 0, sizeof(double) * m, st));
    if (fn) XC_CUDA_TRY(ctx, cudaMemsetAsync(fn, 0, sizeof(double) * m, st));
    if (order == XC_SUM_ORDERED) {
        if (dtype == XC_F32) confmat_csrc_ordered_kernel<float><<<1, 128, 0, st>>>((const float *)t_data, t_idx, t_ptr, pred_idx, k, n, tp, fp, fn);
        else confmat_csrc_ordered_kernel<double><<<1, 128, 0, st>>>((const double *)t_data, t_idx, t_ptr, pred_idx, k, n, tp, fp, fn);
    } else {
        int grid = cap_grid(ctx, (n + 7) / 8);
        if (dtype == XC_F32) confmat_csrc_fast_kernel<float><<<grid, kThreads, 0, st>>>((const float *)t_data, t_idx, t_ptr, pred_idx, k, n, tp, fp, fn);
        else confmat_csrc_fast_kernel<double><<<grid, kThreads, 0, st>>>((const double *)t_data, t_idx, t_ptr, pred_idx, k, n, tp, fp, fn);
    }
    XC_LAUNCHED(ctx);
    return XC_OK;
}

int xc_utility_launch(xc_ctx *ctx, const xc_metric_params *p, int agg, const double *tp, const double *fp,
                      const double *fn, const double *tn, double tn_rows, int64_t m, double *out_dev, cudaStream_t st)
{
    if (!ctx || !p || !tp || !fp || !fn || !out_dev || m <= 0) return XC_ERR_INVALID;
    if (p->metric < 0 || p->metric > XC_METRIC_PREC_AT_K) return XC_ERR_INVALID;
    int64_t blocks = (m + 255) / 256;
    int grid = (int)(blocks < XC_RED_MAX_BLOCKS ? blocks : XC_RED_MAX_BLOCKS);
    if (grid > ctx->sm_count * 4) grid = ctx->sm_count * 4;
    utility_kernel<<<grid, 256, 0, st>>>(*p, agg, tp, fp, fn, tn, tn_rows, m, out_dev, ctx->red_partials,
                                         ctx->red_counter);
    XC_LAUNCHED(ctx);
    return XC_OK;
}

extern "C" int xc_utility(xc_ctx *ctx, const xc_metric_params *p, int agg, const double *tp, const double *fp,
                          const double *fn, const double *tn, int64_t m, double *out_dev, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    return xc_utility_launch(ctx, p, agg, tp, fp, fn, tn, -1.0, m, out_dev, (cudaStream_t)stream);
}

extern "C" int xc_confmat_csc_ordered(xc_ctx *ctx, const void *c_data, int dtype, const int32_t *c_rows,
                                      const int64_t *c_ptr, const int32_t *t_idx, const int64_t *t_ptr,
                                      const int32_t *pred_idx, int k, int64_t n, int64_t m, double *tp, double *fp,
                                      double *fn, int *lone_flag_dev, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !c_rows || !c_ptr || !t_ptr || !pred_idx || !tp || !fp || !fn || !lone_flag_dev) return XC_ERR_INVALID;
    if (n <= 0 || m <= 0 || k < 1) return XC_ERR_INVALID;
    if (dtype != XC_F32 && dtype != XC_F64) return XC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    XC_CUDA_TRY(ctx, cudaMemsetAsync(lone_flag_dev, 0, sizeof(int), st));
    confmat_lone_check_kernel<<<cap_grid(ctx, (n * k + kThreads - 1) / kThreads), kThreads, 0, st>>>(t_idx, t_ptr, pred_idx,
                                                                                                  k, n, lone_flag_dev);
    XC_LAUNCHED(ctx);
    int grid = cap_grid(ctx, (m + 7) / 8);
    if (dtype == XC_F32)
        confmat_csc_ordered_kernel<float><<<grid, kThreads, 0, st>>>((const float *)c_data, c_rows, c_ptr, pred_idx, k, m, tp, fp, fn);
    else
        confmat_csc_ordered_kernel<double><<<grid, kThreads, 0, st>>>((const double *)c_data, c_rows, c_ptr, pred_idx, k, m, tp, fp, fn);
    XC_LAUNCHED(ctx);
    return XC_OK;
}
