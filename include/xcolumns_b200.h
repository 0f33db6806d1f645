/*
 * xcolumns_b200 -- C ABI of the B200-native prediction-optimisation path.
 *
 * Drop-in boundary for the hot path of mwydmuch/xCOLUMNs 0.0.3.
 * The reference has no FFI of its own (pure Python + numba); each entry point below names
 * the reference function(s) it replaces ("ref:", paths relative to the reference repo).
 * INTEGRATION.md shows the ctypes stub a maintainer would add on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - all calls are asynchronous on `stream` unless documented otherwise;
 *   - return value: 0 on success, negative XC_ERR_* otherwise (never throws);
 *   - matrices are row-major with leading dimension `ld` (elements); the 128-bit fast path
 *     needs ld % 4 == 0 (f32) / ld % 2 == 0 (f64) and a 16-byte aligned base, otherwise a
 *     scalar path is used;
 *   - label ids are int32, CSR indptr is int64, accumulators ("state") are float64;
 *   - a compact prediction is an [n, k] int32 matrix of label ids, ascending inside a row,
 *     -1 marks an unused slot (CSR rows with fewer than k stored labels).
 */
#ifndef XCOLUMNS_B200_H
#define XCOLUMNS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XC_API __attribute__((visibility("default")))

#define XC_ABI_VERSION 3

/* element types of probability matrices / weight vectors */
enum { XC_F32 = 0, XC_F64 = 1 };

/* binary metrics, ref: xcolumns/metrics.py:585-944 */
enum {
    XC_METRIC_PRECISION = 0,    /* tp / (tp + fp + eps)                               :603 */
    XC_METRIC_RECALL = 1,       /* tp / (tp + fn + eps)                               :652 */
    XC_METRIC_FBETA = 2,        /* (1+b^2) tp / (b^2 (tp+fp) + tp + fn + eps)         :703 */
    XC_METRIC_JACCARD = 3,      /* tp / (tp + fp + fn + eps)                          :797 */
    XC_METRIC_BALANCED_ACC = 4, /* (tpr + tnr) / 2                                    :843 */
    XC_METRIC_GMEAN = 5,        /* sqrt(tpr * tnr)                                    :890 */
    XC_METRIC_HMEAN = 6,        /* 2 tpr tnr / (tpr + tnr)                            :939 */
    XC_METRIC_PREC_AT_K = 7     /* tp / k  (c1 carries k)                             :513 */
};

/* summation order of label-wise reductions */
enum {
    XC_SUM_FAST = 0,   /* float64 atomics, any order (value parity ~1e-12 relative)           */
    XC_SUM_ORDERED = 1 /* one running sum per label, rows in order: bit-identical to numpy's  */
                       /* axis-0 reduction / numba's row loop (sequential-exact BCA)          */
};

enum {
    XC_OK = 0,
    XC_ERR_INVALID = -1,     /* bad argument (shape, k, dtype, metric id, NULL pointer)   */
    XC_ERR_UNSUPPORTED = -2, /* valid request this build has no kernel for                */
    XC_ERR_CUDA = -3,        /* a CUDA runtime call failed; see xc_last_cuda_error        */
    XC_ERR_NOMEM = -4
};

typedef struct xc_ctx xc_ctx;

typedef struct {
    int32_t metric;    /* XC_METRIC_*                                                          */
    int32_t maximize;  /* 1: ascend, 0: descend (ref: block_coordinate.py:187-188)             */
    int32_t skip_tn;   /* 1: tn is a constant -1 vector (ref: confusion_matrix.py:391-393)     */
    int32_t mix;       /* 1: mixed utility (ref: block_coordinate.py:848-1045, frank_wolfe.py:838-915): */
                       /*    ((1 - mix_alpha) * (tp / mix_k)) + ((mix_alpha * metric) / mix_m)          */
                       /* 2: micro average, metric(tp.sum(), fp.sum(), fn.sum(), tn.sum())              */
                       /*    (ref: metrics.py:68-100; Frank-Wolfe objective only)                       */
                       /* 3: (1 - mix_alpha) * recall + mix_alpha * precision, summed over the labels   */
                       /*    (ref: frank_wolfe.py:917-938; metric = XC_METRIC_PRECISION, FW only)       */
    double c1;         /* 1 + beta**2, computed by the host exactly like python does           */
    double beta2;      /* beta**2                                                              */
    double eps;        /* epsilon of the metric (metric_kwargs["epsilon"], default 1e-9)       */
    double n_div;      /* n if normalize_conf_matrix else 1 (ref: block_coordinate.py:149-151) */
    double n_rows;     /* number of instances (all ranks), whatever the normalisation           */
    double mix_alpha;  /* weight of the macro metric in the mixed utility                      */
    double mix_k;      /* k of precision@k                                                     */
    double mix_m;      /* number of labels                                                     */
} xc_metric_params;

/* ---- context -------------------------------------------------------------------------- */
XC_API int xc_abi_version(void);
XC_API const char *xc_strerror(int code);
/* one context per device and host thread: scratch buffers, SM count, launch counter */
XC_API int xc_ctx_create(int device, xc_ctx **out);
XC_API void xc_ctx_destroy(xc_ctx *ctx);
XC_API const char *xc_last_cuda_error(xc_ctx *ctx);
/* number of kernels this context has launched so far (bench.py: gpu_launches) */
XC_API int64_t xc_launch_count(xc_ctx *ctx);
XC_API int xc_sm_count(xc_ctx *ctx);

/* Pseudo-random permutation of 0..n-1 (visiting order of one batched sweep), written to out[n].
 * The sequential mode does not use it: there the order comes from numpy's Generator on the host,
 * bit-identical to the reference (block_coordinate.py:413-419).                                */
XC_API int xc_permutation(xc_ctx *ctx, int64_t n, uint64_t seed, int32_t *out, void *stream);

/* ---- weighted per-instance top-k ------------------------------------------------------ */
/* ref: weighted_prediction.py:25-60 (_predict_weighted_per_instance_dense), :91-220.
 * gains = eta [* a] [+ b] evaluated in `g_dtype` (separate multiply and add, no FMA, so the
 * selection is bit-comparable with numpy); per row the k largest gains, ties -> lowest label.
 * rows: optional list of n_rows row ids to process (NULL = rows 0..n_rows-1); output row r
 * belongs to rows[r].  out_val (optional) receives the gains in g_dtype.  1 <= k <= 32.    */
XC_API int xc_topk_dense(xc_ctx *ctx, const void *eta, int eta_dtype, int64_t n_rows, int64_t m,
                         int64_t ld, const int32_t *rows, const void *a, const void *b,
                         int g_dtype, int k, int32_t *out_idx, void *out_val, void *stream);
/* ref: numba_csr_functions.py:586-629 (numba_predict_weighted_per_instance_csr), :456-484.
 * Rows with nnz <= k keep all their labels; unused slots get id -1 in out_idx and (when
 * out_val != NULL) value 1 -- the host shim turns them into the reference's (0, 1) filler.  */
XC_API int xc_topk_csr(xc_ctx *ctx, const void *data, int dtype, const int32_t *indices,
                       const int64_t *indptr, int64_t n_rows, const void *a, const void *b, int k,
                       int32_t *out_idx, void *out_val, void *stream);
/* k == 0 branch of weighted_prediction.py:57-58: out[i][j] = (gain >= th) in out dtype=eta dtype */
XC_API int xc_threshold_dense(xc_ctx *ctx, const void *eta, int eta_dtype, int64_t n_rows,
                              int64_t m, int64_t ld, const void *a, const void *b, int g_dtype,
                              double th, void *out, int64_t ld_out, void *stream);
/* scatter a compact prediction into a dense 0/1 (or gain-valued) matrix that was zeroed */
XC_API int xc_scatter_pred_dense(xc_ctx *ctx, const int32_t *pred_idx, const void *val,
                                 int val_dtype, int k, int64_t n_rows, void *out, int out_dtype,
                                 int64_t ld_out, void *stream);

/* HOST helper (all pointers are host pointers): writes the dense n x m prediction matrix the
 * reference returns for dense inputs (weighted_prediction.py:35,47; block_coordinate.py:191-198)
 * from a compact prediction, zero-filling and scattering with `nthreads` host threads
 * (0 = hardware concurrency, at most 32).  val_host may be NULL (ones).                       */
XC_API int xc_fill_pred_dense_host(void *out_host, int out_dtype, int64_t n, int64_t m, int64_t ld,
                                   const int32_t *idx_host, const void *val_host, int val_dtype,
                                   int k, int nthreads);
/* the two halves of the above: clear `bytes` of host memory on host threads (started by the shim at
 * call entry, overlapping the upload and the sweeps), then only scatter into the cleared matrix     */
XC_API int xc_zero_host(void *p, int64_t bytes, int nthreads);
XC_API int xc_scatter_pred_dense_host(void *out_host, int out_dtype, int64_t n, int64_t m, int64_t ld,
                                   const int32_t *idx_host, const void *val_host, int val_dtype,
                                   int k, int nthreads);

/* ---- label-wise confusion sums -------------------------------------------------------- */
/* ref: confusion_matrix.py:160-202, :364-399 (calculate_confusion_matrix), dense x dense.
 * axis 0: tp/fp/fn have m entries; axis 1: n entries.  Products are formed in the input
 * dtype, sums in float64 (acc_f32 = 1 with XC_SUM_ORDERED keeps float32 running sums like
 * dtype=None does).  normalize / tn are m-vector epilogues left to the host shim.           */
XC_API int xc_confmat_dense(xc_ctx *ctx, const void *y_true, int64_t ldt, const void *y_pred,
                            int64_t ldp, int dtype, int64_t n, int64_t m, int axis, int order,
                            int acc_f32, double *tp, double *fp, double *fn, void *stream);
/* same sums with a compact prediction (k ids per row); colsum (optional, fast order only) is
 * sum_i y_true[i][j] so that fn = colsum - tp without a second pass over y_true.
 * XC_SUM_ORDERED: one thread per label adds its column strictly in row order; the prediction is
 * first expanded into a bitmap (n x ceil(m / 32) words in the context's scratch buffer, when that
 * is <= 256 MB) so that a row costs a warp one broadcast load instead of k loads per lane.   */
XC_API int xc_confmat_dense_compact(xc_ctx *ctx, const void *y_true, int dtype, int64_t ld,
                                    const int32_t *pred_idx, int k, int64_t n, int64_t m,
                                    int order, const double *colsum, double *tp, double *fp,
                                    double *fn, void *stream);
/* ref: numba_csr_functions.py:144-182, 217-258 via confusion_matrix.py:174-228 (both CSR) */
XC_API int xc_confmat_csr(xc_ctx *ctx, const void *t_data, const int32_t *t_idx,
                          const int64_t *t_ptr, const void *p_data, const int32_t *p_idx,
                          const int64_t *p_ptr, int dtype, int64_t n, int64_t m, int order,
                          int acc_f32, double *tp, double *fp, double *fn, void *stream);
/* CSR truth, compact prediction of ones */
/* fn may be NULL with XC_SUM_FAST: only tp / fp are accumulated (k look-ups per row instead of a pass
 * over the whole row); the caller derives fn = colsum(y_true) - tp.                                  */
XC_API int xc_confmat_csr_compact(xc_ctx *ctx, const void *t_data, const int32_t *t_idx,
                                  const int64_t *t_ptr, int dtype, const int32_t *pred_idx, int k,
                                  int64_t n, int64_t m, int order, double *tp, double *fp,
                                  double *fn, void *stream);
/* XC_SUM_ORDERED sums for a CSR truth matrix through its column-major copy (c_data, c_rows, c_ptr:
 * for every label its stored entries in row order -- a stable sort of the CSR by label, built once
 * per call by the host shim).  One warp per label adds that label's contributions strictly in row
 * order, bit-identical to xc_confmat_csr_compact(XC_SUM_ORDERED).  Only valid when every predicted
 * label is stored in its row: *lone_flag_dev is set to 1 otherwise and the caller must use the
 * row-walking entry point instead.                                                              */
XC_API int xc_confmat_csc_ordered(xc_ctx *ctx, const void *c_data, int dtype, const int32_t *c_rows,
                                  const int64_t *c_ptr, const int32_t *t_idx, const int64_t *t_ptr,
                                  const int32_t *pred_idx, int k, int64_t n, int64_t m, double *tp,
                                  double *fp, double *fn, int *lone_flag_dev, void *stream);
/* column sums of a dense / CSR matrix in float64 (fast order) */
XC_API int xc_colsum_dense(xc_ctx *ctx, const void *x, int dtype, int64_t n, int64_t m, int64_t ld,
                           double *out, void *stream);
XC_API int xc_colsum_csr(xc_ctx *ctx, const void *data, int dtype, const int32_t *indices,
                         int64_t nnz, int64_t m, double *out, void *stream);
/* ref: block_coordinate.py:54-90 (_calculate_utility): mean (agg=0) or sum (agg=1) over labels
 * of the binary metric on (tp,fp,fn,tn)/n_div; result written to *out_dev (device double).   */
XC_API int xc_utility(xc_ctx *ctx, const xc_metric_params *p, int agg, const double *tp,
                      const double *fp, const double *fn, const double *tn, int64_t m,
                      double *out_dev, void *stream);

/* ---- BCA, sequential-exact mode ("Gauss-Seidel", reference instance order) -------------- */
/* ref: block_coordinate.py:132-209 (_bc_with_0approx_step_dense) for i in order (:448-463).
 * One cooperative launch per sweep; float64 arithmetic in the reference's operation order
 * (compiled without FMA contraction).  tp/fp/fn/tn are the un-normalised running sums,
 * updated in place; pred_idx [n, k] is updated in place.  greedy: skip the removal step.    */
XC_API int xc_bca_exact_sweep_dense(xc_ctx *ctx, const void *eta, int dtype, int64_t n, int64_t m,
                                    int64_t ld, const int32_t *order, int64_t n_order, int k,
                                    const xc_metric_params *p, int greedy, int32_t *pred_idx,
                                    double *tp, double *fp, double *fn, double *tn, void *stream);
/* Online / greedy micro-batch (ref: block_coordinate.py:132-209 called with greedy=True, only_pred=True,
 * followed by confusion_matrix.py:402-435 _update_unnormalized_confusion_matrix, as driven by
 * experiments/omma_wrappers_online_methods.py:192-270): rows 0..n_rows-1 in sequence, each predicted from
 * the current state (its own contribution is not removed), then the state is advanced with the row of
 * y_true (NULL: with the probabilities themselves, the "ETU" variant).  One cluster launch per micro-batch;
 * label spaces beyond 32 k return XC_ERR_UNSUPPORTED.  The state stays on the device between calls.  */
XC_API int xc_bca_online_dense(xc_ctx *ctx, const void *eta, int dtype, int64_t n_rows, int64_t m,
                               int64_t ld, const void *y_true, int64_t ld_true, int k,
                               const xc_metric_params *p, int32_t *pred_idx, double *tp, double *fp,
                               double *fn, double *tn, void *stream);
/* The same for CSR rows (ref: block_coordinate.py:212-293 with greedy=True, only_pred=True, then
 * confusion_matrix.py:421-432 -> numba_csr_functions.py:421-452; experiments/omma_wrappers_online_methods.py:223-266):
 * candidates are the row's stored labels; t_*: the CSR rows of y_true (pass the probability rows again for the
 * "ETU" variant).  The reference adds 1 to tn of every label per instance; here that is replayed lazily and
 * bit-exactly (tn_last [m] int32, zero before the first call; step0 = instances of earlier calls); on return tn
 * is up to date for all labels.  pred_idx [n_rows, k]: ascending labels, -1 for rows with fewer than k entries. */
XC_API int xc_bca_online_csr(xc_ctx *ctx, const void *data, int dtype, const int32_t *indices,
                             const int64_t *indptr, const void *t_data, const int32_t *t_indices,
                             const int64_t *t_indptr, int64_t n_rows, int64_t m, int k,
                             const xc_metric_params *p, int32_t *pred_idx, double *tp, double *fp, double *fn,
                             double *tn, int32_t *tn_last, int64_t step0, void *stream);
/* k == 0 (no budget, ref: block_coordinate.py:199-200): every label with gain >= 0 is predicted.
 * pred is a dense [n, ld_pred] 0/1 matrix of eta's dtype, updated in place.                     */
XC_API int xc_bca_exact_sweep_dense_k0(xc_ctx *ctx, const void *eta, int dtype, int64_t n, int64_t m,
                                       int64_t ld, const int32_t *order, int64_t n_order,
                                       const xc_metric_params *p, int greedy, void *pred,
                                       int64_t ld_pred, double *tp, double *fp, double *fn,
                                       double *tn, void *stream);
/* ref: block_coordinate.py:212-293 (_bc_with_0approx_step_csr) + numba_csr_functions.py
 * :386-452, :456-466, :500-546.  tn may be NULL when p->skip_tn.                              */
XC_API int xc_bca_exact_sweep_csr(xc_ctx *ctx, const void *data, int dtype, const int32_t *indices,
                                  const int64_t *indptr, int64_t n, int64_t m,
                                  const int32_t *order, int64_t n_order, int k,
                                  const xc_metric_params *p, int greedy, int32_t *pred_idx,
                                  double *tp, double *fp, double *fn, double *tn, void *stream);
/* ref: block_coordinate.py:539-580 (_bc_for_coverage_step_csr); Ef updated in place */
XC_API int xc_cov_exact_sweep_csr(xc_ctx *ctx, const void *data, int dtype, const int32_t *indices,
                                  const int64_t *indptr, int64_t n, int64_t m,
                                  const int32_t *order, int64_t n_order, int k, double alpha,
                                  int greedy, int32_t *pred_idx, double *Ef, void *stream);
/* ref: numba_csr_functions.py:325-382 as called at block_coordinate.py:665/676 */
XC_API int xc_cov_state_csr(xc_ctx *ctx, const void *data, int dtype, const int32_t *indices,
                            const int64_t *indptr, int64_t n, int64_t m, const int32_t *pred_idx,
                            int k, int order, double *Ef, void *stream);

/* ---- BCA, batched block-Jacobi mode ----------------------------------------------------- */
/* Per-label gain coefficients from the frozen state: for the metrics
 * whose gain is affine in eta (precision, recall, F-beta)
 *     gain_ij = A_j + B_j * eta_ij          label j not selected in row i   -> coef_n[j] = (B, A)
 *     gain_ij = A'_j + B'_j * eta_ij        label j currently selected      -> coef_s[j] = (B', A')
 * Before computing them the pending deltas are folded into the state and cleared:
 *     tp += dtp; fp += dfp; fn += dfn; d* = 0     (d* may be NULL).
 * ref: block_coordinate.py:158-185 evaluated against a frozen state.                        */
XC_API int xc_bca_coef(xc_ctx *ctx, const xc_metric_params *p, double *tp, double *fp, double *fn,
                       double *dtp, double *dfp, double *dfn, int64_t m, float *coef_n,
                       float *coef_s, void *stream);   /* coef_*: xc_bca_coef_len(m) float2 each */
/* affine-gain metrics: precision, recall, F-beta and balanced accuracy (for the latter tp + fn and
 * tn + fp are prediction-independent column sums, so the gain is state-free like recall's)      */
/* Number of float2 entries coef_n / coef_s must provide for m labels (m rounded up to whole
 * coefficient tiles of the TMA-pipelined kernel; entries past m are never used for a gain).    */
XC_API int64_t xc_bca_coef_len(int64_t m);
/* Rows one full wave of the dense batch kernel covers (SMs x resident CTAs x warps x rows per
 * warp): batches that are a multiple of it leave no partially filled last wave.               */
XC_API int xc_bca_wave_rows(xc_ctx *ctx, int dtype, int64_t m);
/* One batch: rows[0..n_rows) stream past the frozen coefficients, every row re-selects its k
 * best labels (own contribution removed via coef_s), pred_idx rows are rewritten and the
 * confusion deltas of all changed rows are accumulated into dtp/dfp/dfn (float64 atomics). */
XC_API int xc_bca_batch_dense(xc_ctx *ctx, const void *eta, int dtype, int64_t m, int64_t ld,
                              const int32_t *rows, int64_t n_rows, int k, const float *coef_n,
                              const float *coef_s, int32_t *pred_idx, double *dtp, double *dfp,
                              double *dfn, void *stream);
/* CSR rows.  max_row_nnz: an upper bound of the stored labels per row if the caller knows one, else 0; rows of at
 * most 128 labels are held in registers and selected by warp reductions.  touch_*: optional "touched label" list
 * (flag [m] zero-initialised, list [m], ctl [4] zero-initialised; only with 0 < max_row_nnz <= 128): every label
 * whose deltas the batch changes is appended once, so that xc_bca_fold_touched can fold / refresh just those
 * labels instead of all m.  ref: block_coordinate.py:212-293 (_bc_with_0approx_step_csr), frozen state.        */
XC_API int xc_bca_batch_csr(xc_ctx *ctx, const void *data, int dtype, const int32_t *indices,
                            const int64_t *indptr, const int32_t *rows, int64_t n_rows, int k,
                            const float *coef_n, const float *coef_s, int32_t *pred_idx,
                            double *dtp, double *dfp, double *dfn, int max_row_nnz, int32_t *touch_flag,
                            int32_t *touch_list, int32_t *touch_ctl, void *stream);
XC_API int xc_bca_fold_touched(xc_ctx *ctx, const xc_metric_params *p, double *tp, double *fp, double *fn,
                               double *dtp, double *dfp, double *dfn, int32_t *touch_flag,
                               int32_t *touch_list, int32_t *touch_ctl, float *coef_n, float *coef_s,
                               void *stream);
/* coverage: gain = Ef_j * eta (not selected) or Ef_j / (1 - eta) * eta (selected); the batch
 * accumulates multiplicative factors into dEf (init 1), folded by xc_cov_fold.               */
XC_API int xc_cov_batch_csr(xc_ctx *ctx, const void *data, int dtype, const int32_t *indices,
                            const int64_t *indptr, const int32_t *rows, int64_t n_rows, int k,
                            double alpha, const double *Ef, int32_t *pred_idx, double *dEf,
                            void *stream);
XC_API int xc_cov_batch_dense(xc_ctx *ctx, const void *eta, int dtype, int64_t m, int64_t ld,
                              const int32_t *rows, int64_t n_rows, int k, double alpha,
                              const double *Ef, int32_t *pred_idx, double *dEf, void *stream);
XC_API int xc_cov_fold(xc_ctx *ctx, double *Ef, double *dEf, int64_t m, void *stream);

/* One whole batched coverage sweep of a single process as one host call (batch kernel + fold per `batch`
 * entries of `order`); dEf must hold ones on entry and holds ones on return.                          */
XC_API int xc_cov_sweep_csr(xc_ctx *ctx, const void *data, int dtype, const int32_t *indices,
                            const int64_t *indptr, int64_t m, const int32_t *order, int64_t n_order,
                            int64_t batch, int k, double alpha, double *Ef, int32_t *pred_idx, double *dEf,
                            void *stream);
XC_API int xc_cov_sweep_dense(xc_ctx *ctx, const void *eta, int dtype, int64_t m, int64_t ld,
                              const int32_t *order, int64_t n_order, int64_t batch, int k, double alpha,
                              double *Ef, int32_t *pred_idx, double *dEf, void *stream);
/* Ef of a compact prediction over DENSE rows (any order, float64 compare-and-swap products); zero
 * probabilities count as "not stored", i.e. the CSR semantics of block_coordinate.py:539-580 that the
 * coverage path follows for both layouts.  ref: numba_csr_functions.py:325-382.                       */
XC_API int xc_cov_state_dense(xc_ctx *ctx, const void *eta, int dtype, int64_t n, int64_t m, int64_t ld,
                              const int32_t *pred_idx, int k, double *Ef, void *stream);
/* out_dev[0] = 1 - mean(Ef), ref: block_coordinate.py:583-597 (the alpha = 1 part)                    */
XC_API int xc_cov_utility(xc_ctx *ctx, const double *Ef, int64_t m, double *out_dev, void *stream);

/* Jaccard / G-mean / H-mean: the gain is not affine in eta but still a closed form of eta and four
 * per-label numbers (rec: 4 floats per label, 16-byte aligned; see csrc/bca_batched.cu for the algebra).
 * xc_bca_rec is the counterpart of xc_bca_coef (folds the pending deltas, writes the records);
 * xc_bca_batch_dense_rec the counterpart of xc_bca_batch_dense (tp/fp/fn: the frozen state, read for the
 * k currently selected labels of every row).  xc_bca_sweep_dense_pipe takes the record array in place of the
 * coefficient sets for these metrics.                              */
XC_API int xc_bca_rec(xc_ctx *ctx, const xc_metric_params *p, double *tp, double *fp, double *fn,
                      double *dtp, double *dfp, double *dfn, int64_t m, float *rec, void *stream);
XC_API int xc_bca_batch_csr_rec(xc_ctx *ctx, const xc_metric_params *p, const void *data, int dtype,
                                const int32_t *indices, const int64_t *indptr, const int32_t *rows,
                                int64_t n_rows, int k, const float *rec, const double *tp,
                                const double *fp, const double *fn, int32_t *pred_idx, double *dtp,
                                double *dfp, double *dfn, void *stream);
XC_API int xc_bca_batch_dense_rec(xc_ctx *ctx, const xc_metric_params *p, const void *eta, int dtype,
                                  int64_t m, int64_t ld, const int32_t *rows, int64_t n_rows, int k,
                                  const float *rec, const double *tp, const double *fp,
                                  const double *fn, int32_t *pred_idx, double *dtp, double *dfp,
                                  double *dfn, void *stream);

/* One whole batched sweep of a single process as one host call: per batch of `batch` entries of `order`
 * the coefficient (or record) kernel with the fold of the pending deltas and the batch kernel, then a
 * final fold, i.e. tp/fp/fn end up as the state after the sweep.  coef_a: the coefficient pairs of
 * xc_bca_coef, or the records of xc_bca_rec for Jaccard / G-mean / H-mean.  Small problems are bound by
 * the host-side call overhead of the per-batch entry points, not by the GPU.                       */
XC_API int xc_bca_sweep_dense(xc_ctx *ctx, const xc_metric_params *p, const void *eta, int dtype, int64_t m,
                              int64_t ld, const int32_t *order, int64_t n_order, int64_t batch, int k,
                              float *coef_a, float *coef_s, int32_t *pred_idx, double *tp, double *fp,
                              double *fn, double *dtp, double *dfp, double *dfn, void *stream);
/* CSR form; max_row_nnz / touch_* as in xc_bca_batch_csr (NULL: every fold visits all m labels); refresh != 0:
 * the host changed tp/fp/fn since the last call, all coefficients are recomputed first.                */
XC_API int xc_bca_sweep_csr(xc_ctx *ctx, const xc_metric_params *p, const void *data, int dtype,
                            const int32_t *indices, const int64_t *indptr, int64_t m,
                            const int32_t *order, int64_t n_order, int64_t batch, int k, float *coef_n,
                            float *coef_s, int32_t *pred_idx, double *tp, double *fp, double *fn,
                            double *dtp, double *dfp, double *dfn, int max_row_nnz, int32_t *touch_flag,
                            int32_t *touch_list, int32_t *touch_ctl, int refresh, void *stream);

/* ---- commits over peer memory (rows sharded over the GPUs of one box) -------------------- */
/* One window per rank: cudaMalloc'ed, exported with CUDA IPC (ipc_handle_out: 64 bytes), mapped by
 * every other rank with xc_p2p_open(handles = the world * 64 gathered bytes).  xc_p2p_payload
 * returns the local payload (device pointer, payload_bytes long, zero-initialised); for the batched
 * sweep it holds xc_bca_pipe_buffers(lag) delta buffers of xc_bca_delta_stride(m) bytes each,
 * [dtp | dfp | dfn].        */
typedef struct xc_p2p xc_p2p;
XC_API int xc_p2p_create(xc_ctx *ctx, int world, int rank, int64_t payload_bytes, xc_p2p **out,
                         void *ipc_handle_out);
XC_API int xc_p2p_open(xc_ctx *ctx, xc_p2p *w, const void *handles);
XC_API void *xc_p2p_payload(xc_p2p *w);
XC_API int xc_p2p_error(xc_ctx *ctx, xc_p2p *w, unsigned *out);
XC_API void xc_p2p_destroy(xc_ctx *ctx, xc_p2p *w);
XC_API int64_t xc_bca_delta_stride(int64_t m);
/* Pipelined block-Jacobi sweeps over this rank's DENSE rows, one host call per sweep, single process (w = NULL)
 * or with the rows sharded over the GPUs of one box (w = the peer window; every rank calls with the same
 * n_batches / lag / flags).  ref: block_coordinate.py:415-493 (the sweep loop: shuffle :419, per-instance steps
 * :448-463, utility :465-476) evaluated batch-wise.
 * Per batch g = batch0 + b (rows order[b * batch ...]): the streaming batch kernel accumulates the deltas of its
 * rows into delta buffer g % NB; the commit then folds that buffer -- with a window: a push kernel copies it into
 * this rank's slot of every peer's inbox over NVLink and raises the rank's flag there, the commit kernel waits for
 * every rank's flag and adds the W buffers (local memory) in rank order, so that the replicated state stays
 * bit-identical -- into tp/fp/fn, refreshes gain-coefficient set g % (lag + 1) (or the records of
 * Jaccard / G-mean / H-mean) and clears buffer (g + lag + 1) % NB.  NB = xc_bca_pipe_buffers(lag) buffers of
 * xc_bca_delta_stride(m) bytes each: at the start of the window's payload (xc_bca_window_bytes), or `delta` when w
 * is NULL; all zero before the first call, never touched by the host afterwards.  batch0: number of batches of all earlier calls on these
 * buffers (the rotation continues across sweeps).
 * lag = 0: strict order K_0, commit_0, K_1, ... on `stream`.  lag = L (1..3): batch g sees the state after commit
 * g - L - 1; L + 1 consecutive batch kernels run concurrently on as many internal streams, every commit overlaps
 * the other batches' streaming, and the pipeline keeps running ACROSS calls: a sweep only waits for the previous sweep's
 * batch kernels, the previous sweep's utility is computed behind its last commit, and `stream` is made to wait
 * for that utility only.  XC_PIPE_FORK (or the first call, or a call after xc_bca_pipe_join) re-synchronises with
 * `stream` first and recomputes every coefficient set from tp/fp/fn (needed after the host changed the state or
 * the prediction).  xc_bca_pipe_join makes `stream` wait for everything in flight; it must precede any access
 * to pred_idx / tp / fp / fn from `stream`.  Record metrics always run with lag 0.                          */
enum { XC_PIPE_FORK = 1, XC_PIPE_SHUFFLE = 2 };
typedef struct {
    const xc_metric_params *params;
    const void *eta;          /* [n_rows, ld] row-major                                                        */
    int32_t dtype;            /* XC_F32 / XC_F64                                                               */
    int32_t k;
    int64_t m, ld, n_rows;
    int64_t batch, n_batches; /* rows per commit; commits of this sweep (>= ceil(n_rows / batch): ragged shards) */
    int64_t batch0;
    int32_t lag;
    int32_t flags;            /* XC_PIPE_*                                                                     */
    uint64_t seed;            /* XC_PIPE_SHUFFLE: visiting order = xc_permutation(n_rows, seed), ref :419      */
    int64_t sweep;            /* sweep counter, +1 per call: with XC_PIPE_SHUFFLE the order lives in order + (sweep & 1) * n_rows */
    int32_t *order;           /* XC_PIPE_SHUFFLE: 4 * n_rows + n_rows / 256 + 8 int32 of scratch [order A | order B |  */
                              /* raw order | stamps | counts], zero before the first call; else the visiting order */
    float *coef;              /* (lag + 1) sets of 4 * xc_bca_coef_len(m) floats, each [coef_n | coef_s] (or records) */
    int32_t *pred_idx;        /* [n_rows, k], rewritten                                                        */
    int32_t *pred_snapshot;   /* optional [n_rows, k]: every visited row's selection as it was BEFORE this sweep */
                              /* (written by the batch kernels; complete when the sweep is; roll-back)           */
    double *tp, *fp, *fn;     /* replicated float64 state                                                     */
    double *delta;            /* w == NULL: the NB delta buffers                                               */
    const xc_metric_params *util_params;  /* optional: utility of the state after this sweep (ref :468-476) -> */
    double *util_out;         /*           util_out[0] (device)                                                */
    int32_t agg;              /* 0 mean, 1 sum                                                                 */
    int32_t reserved;
    double util_tn_rows;      /* >= 0: tn = -tp - fp - fn + util_tn_rows inside the utility; < 0: tn = -1     */
    int64_t prev_tail_from;   /* position in the PREVIOUS sweep's order where its last `lag` batches begin, -1:   */
                              /* unknown.  With XC_PIPE_SHUFFLE and a known tail the new order is repaired so    */
                              /* that its first `lag` batches avoid those rows and the sweeps overlap without a  */
                              /* drain; otherwise the new sweep's kernels wait for the previous sweep's.          */
} xc_bca_pipe_args;
XC_API int xc_bca_pipe_buffers(int lag);
/* payload bytes a peer window needs for xc_bca_pipe_sweep: NB own delta buffers + world x NB inbox buffers (every
 * rank PUSHES its deltas into its slot of every peer's inbox after a batch, so a commit only reads local memory) */
XC_API int64_t xc_bca_window_bytes(int64_t m, int lag, int world);
XC_API int xc_bca_pipe_sweep(xc_ctx *ctx, xc_p2p *w, const xc_bca_pipe_args *a, void *stream);
XC_API int xc_bca_pipe_join(xc_ctx *ctx, void *stream);

/* ---- host -> device upload of PAGEABLE host memory (what a numpy caller of the reference passes) ---- */
/* rows x width_bytes from src_host (pitch src_pitch, ordinary pageable memory) to dst_dev (pitch dst_pitch):
 * host threads copy chunk c + 1 into one of two pinned staging buffers while the DMA of chunk c runs.
 * Returns when the last DMA has completed.  Replaces the implicit staged copy of the reference's
 * torch / numpy conversion at its entry points (block_coordinate.py:388-401).                         */
XC_API int xc_h2d_staged(xc_ctx *ctx, void *dst_dev, int64_t dst_pitch, const void *src_host,
                         int64_t src_pitch, int64_t width_bytes, int64_t rows, int nthreads, void *stream);

/* ---- per-launch timing of the streaming batch kernels (measurement only) ------------------ */
/* While enabled, xc_bca_pipe_sweep brackets every batch kernel with CUDA events on the stream it
 * is launched on.  xc_timing_read synchronises the device and returns for up to `cap` launches the start
 * and end time (ms since the first recorded event) and the rows processed (HOST arrays), the number of
 * recorded launches in *count_host, and clears the log.                                              */
XC_API int xc_timing_enable(xc_ctx *ctx, int on);   /* 0 off, 1 batch kernels, 2 also commits (logged with rows = 0) */
XC_API int xc_timing_read(xc_ctx *ctx, int cap, double *start_ms_host, double *end_ms_host,
                          int64_t *rows_host, int *count_host);

/* ---- Frank-Wolfe iterate ----------------------------------------------------------------- */
/* ref: frank_wolfe.py:601-606: weighted top-k of every row with the linear classifier (a, b)
 * fused with the accumulation of tp_j = sum_i y_true[i][j] * yhat[i][j] and
 * cnt_j = sum_i yhat[i][j] (float64).  a, b and y_true have eta's dtype (numpy promotes the
 * float32 classifier rows to it); gains = eta * a + b with separate IEEE multiply and add.
 * y_true may alias eta.  tp/cnt are zeroed by the call.  pred_idx [n, k] is optional.
 * k = 0 (no budget): every label with a gain >= 0 is predicted; on CSR rows only the STORED
 * labels take part (numba_csr_functions.py:516-517, :631-653); pred_idx must be NULL then.    */
XC_API int xc_fw_iterate_dense(xc_ctx *ctx, const void *eta, int dtype, int64_t n, int64_t m,
                               int64_t ld, const void *y_true, int64_t ld_true, const void *a,
                               const void *b, int k, double *tp, double *cnt, int32_t *pred_idx,
                               void *stream);
XC_API int xc_fw_iterate_csr(xc_ctx *ctx, const void *data, int dtype, const int32_t *indices,
                             const int64_t *indptr, int64_t n, int64_t m, const void *t_data,
                             const int32_t *t_idx, const int64_t *t_ptr, const void *a,
                             const void *b, int k, double *tp, double *cnt, int32_t *pred_idx,
                             void *stream);
/* ref: confusion_matrix.py:386-399 on the iterate: Ci = [tp, fp, fn, tn] (4 stacked m-vectors)
 * with fp = cnt - tp, fn = colsum(y_true) - tp, optional / n, tn = -tp - fp - fn + (1 | n) or
 * the constant -1 when skip_tn.                                                               */
XC_API int xc_fw_make_conf(xc_ctx *ctx, const double *tp_raw, const double *cnt,
                           const double *colsum, int64_t m, double n, int normalize, int skip_tn,
                           double *Ci, void *stream);
/* ref: frank_wolfe.py:591-596 (+ :368-376): value of the macro-averaged metric on the stacked
 * confusion vectors C and the next linear classifier a = dtp - dfp - dfn + dtn, b = dfp - dtn
 * (negated when minimising), closed-form gradients, stored as float32 like :503-504.
 * a_out/b_out may both be NULL (value only).                                                  */
XC_API int xc_fw_metric_grad(xc_ctx *ctx, const xc_metric_params *p, const double *C, int64_t m,
                             float *a_out, float *b_out, double *value_dev, void *stream);
/* ref: frank_wolfe.py:379-404 + utils.py:174-184 (uniform_search): evaluates the metric of
 * (1-alpha) C + alpha Ci at alpha = 0 and at alphas_dev[0..n_alphas) (the host builds the grid
 * with numpy.arange so the grid points are bit-identical), keeps the FIRST strict maximum.
 * For the c*tp/D metrics a float32 pass over the whole grid pre-selects the points that can be
 * the float64 maximum, a float32 evaluation of their DIFFERENCES to the pre-selected maximum (with a
 * rigorous rounding bound) prunes them to a handful, and those are re-evaluated with the reference's
 * float64 expression.  $XCOLUMNS_B200_FW_SEARCH=full evaluates the whole grid in float64 instead.
 * scratch_dev: xc_fw_alpha_scratch_bytes(m, n_alphas) bytes; result_dev[0] = alpha, [1] = value. */
XC_API int64_t xc_fw_alpha_scratch_bytes(int64_t m, int64_t n_alphas);
XC_API int xc_fw_alpha_search(xc_ctx *ctx, const xc_metric_params *p, const double *C,
                              const double *Ci, int64_t m, const double *alphas_dev,
                              int64_t n_alphas, double *scratch_dev, double *result_dev,
                              void *stream);
/* ref: utils.py:187-201 (ternary_search) on the same objective, eps = alpha_tolerance
 * (frank_wolfe.py:399-400, :627); result_dev[0] = alpha, [1] = value.  Follows the reference's
 * comparison verbatim (it moves `high` when f(mid1) < f(mid2)).                                  */
XC_API int xc_fw_alpha_ternary(xc_ctx *ctx, const xc_metric_params *p, const double *C,
                               const double *Ci, int64_t m, double eps, double *result_dev,
                               void *stream);
/* C = (1 - alpha) C + alpha Ci on the 4 stacked m-vectors, alpha read from device memory */
XC_API int xc_fw_combine(xc_ctx *ctx, double *C, const double *Ci, int64_t m4,
                         const double *alpha_dev, void *stream);

/* One dense Frank-Wolfe iteration as two calls (ref: frank_wolfe.py:589-637); between them the
 * caller all-reduces raw[0..2m) when rows are sharded over ranks.
 *   begin : fused weighted top-k + accumulation of raw = [tp, cnt] with the classifier held in the
 *           float32 rows (a_row, b_row) (e.g. row i of the classifier matrices; written by the
 *           previous finish call).  ab64: 2*(m+1) doubles, used when dtype == XC_F64 (the float32
 *           rows are widened like numpy promotes them).  Rows that are 16-byte aligned are read
 *           with 128-bit loads.  raw_is_zero != 0: raw was already cleared by
 *           the previous finish call (zero_raw), no memset is issued.
 *   finish: first != 0: Cm = confusion vectors of classifier 0, scal[0] = metric(Cm).  Otherwise
 *           Ci = confusion vectors of classifier i, scal[1] = metric(Ci), scal[2..3] = line search
 *           (alphas_dev != NULL) or the fixed step (alpha passed in fixed_alpha), Cm = (1-a) Cm + a Ci,
 *           scal[4] = metric(Cm); ternary_eps > 0 selects the ternary search instead of the grid.
 *           In both cases, when a_next/b_next are given, the NEXT classifier
 *           (gradient of the metric at the new Cm, :591-596) is written to them and scal_next[0]
 *           receives metric(Cm) (the next iteration's "old utility"); zero_raw != 0 clears raw.
 *           The per-label work is fused into two kernels (confusion vectors + utility +
 *           line-search linearisation; combine + utility + gradient).                           */
XC_API int xc_fw_step_begin(xc_ctx *ctx, const void *eta, int dtype, int64_t n, int64_t m, int64_t ld,
                            const void *y_true, int64_t ld_true, const float *a_row,
                            const float *b_row, double *ab64, int k, double *raw, int raw_is_zero,
                            void *stream);
XC_API int xc_fw_step_finish(xc_ctx *ctx, const xc_metric_params *p, int first, double *raw,
                             const double *colsum, int64_t m, double n_global, int normalize,
                             int skip_tn, double *Cm, double *Ci, const double *alphas_dev,
                             int64_t n_alphas, double fixed_alpha, double *scratch_dev, double *scal,
                             float *a_next, float *b_next, double *scal_next, int zero_raw,
                             double ternary_eps, void *stream);
/* byte offset, inside the line-search scratch, of the control block {int count, full, slices, tiles}
 * the two-stage search leaves behind (diagnostics: number of float64 candidates)               */
XC_API int64_t xc_fw_alpha_ctl_offset(int64_t m, int64_t n_alphas);

#ifdef __cplusplus
}
#endif
#endif /* XCOLUMNS_B200_H */
