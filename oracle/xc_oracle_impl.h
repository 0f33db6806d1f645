/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Type-generic bodies of the CPU restatement; included twice by xc_oracle.c with
 *   T   = float / double   (dtype of the probability matrix, "eta")
 *   SFX = f32 / f64
 * Every function cites the reference lines (relative to /root/reference) whose
 * arithmetic it restates. Floating-point operations are written in exactly the
 * order and precision numpy/numba evaluate them so results are bit-comparable
 * with the live reference on tie-free inputs. Compile with -ffp-contract=off.
 */

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)

/* ---------------------------------------------------------------------------------
 * Weighted per-instance top-k, dense rows.
 * Restates xcolumns/weighted_prediction.py:25-60 for the case where eta, a and b
 * share dtype T (numpy keeps T): gains = eta [* a] [+ b], separate mul and add.
 * Selection: k largest gains, ties -> lowest label index (reference: np.argpartition,
 * tie order unspecified). Output indices ascending by label.
 * --------------------------------------------------------------------------------- */
int FN(orc_topk_dense)(const T *eta, int64_t n, int64_t m, int64_t ld, const T *a,
                       const T *b, int k, int32_t *out_idx, T *out_val)
{
    if (k <= 0 || k > m) return -1;
    double *g = (double *)malloc(sizeof(double) * (size_t)m);
    int32_t *sel = (int32_t *)malloc(sizeof(int32_t) * (size_t)k);
    if (!g || !sel) return -2;
    for (int64_t i = 0; i < n; ++i) {
        const T *row = eta + i * ld;
        for (int64_t j = 0; j < m; ++j) {
            T v = row[j];
            if (a) v = v * a[j];
            if (b) v = v + b[j];
            g[j] = (double)v; /* exact widening; ordering unchanged */
        }
        orc_select_topk(g, m, k, sel);
        for (int t = 0; t < k; ++t) {
            out_idx[i * k + t] = sel[t];
            if (out_val) out_val[i * k + t] = (T)g[sel[t]];
        }
    }
    free(g);
    free(sel);
    return 0;
}

/* Mixed precision variant: eta in T, a/b in double -> numpy promotes gains to float64
 * (weighted_prediction.py:37-41 with a float64 weight vector). */
int FN(orc_topk_dense_wd)(const T *eta, int64_t n, int64_t m, int64_t ld, const double *a,
                          const double *b, int k, int32_t *out_idx, double *out_val)
{
    if (k <= 0 || k > m) return -1;
    double *g = (double *)malloc(sizeof(double) * (size_t)m);
    int32_t *sel = (int32_t *)malloc(sizeof(int32_t) * (size_t)k);
    if (!g || !sel) return -2;
    for (int64_t i = 0; i < n; ++i) {
        const T *row = eta + i * ld;
        for (int64_t j = 0; j < m; ++j) {
            double v = (double)row[j];
            if (a) v = v * a[j];
            if (b) v = v + b[j];
            g[j] = v;
        }
        orc_select_topk(g, m, k, sel);
        for (int t = 0; t < k; ++t) {
            out_idx[i * k + t] = sel[t];
            if (out_val) out_val[i * k + t] = g[sel[t]];
        }
    }
    free(g);
    free(sel);
    return 0;
}

/* ---------------------------------------------------------------------------------
 * Weighted per-instance top-k, CSR rows.
 * Restates xcolumns/numba_csr_functions.py:586-629 (+ :456-484): per row, gains over the
 * stored entries only; rows with nnz > k keep the k best (label-ascending), rows with
 * nnz <= k keep every stored label and the remaining slots stay at their
 * pre-allocated (index 0, value 1) filling (:599-601).
 * out_idx / out_val are n*k, row stride k. a, b already cast to T (:72-75 of
 * weighted_prediction.py).
 * --------------------------------------------------------------------------------- */
int FN(orc_topk_csr)(const T *data, const int32_t *indices, const int64_t *indptr, int64_t n,
                     const T *a, const T *b, int k, int keep_scores, int32_t *out_idx,
                     T *out_val)
{
    if (k <= 0) return -1;
    int64_t cap = 0;
    for (int64_t i = 0; i < n; ++i)
        if (indptr[i + 1] - indptr[i] > cap) cap = indptr[i + 1] - indptr[i];
    double *g = (double *)malloc(sizeof(double) * (size_t)(cap + 1));
    int32_t *sel = (int32_t *)malloc(sizeof(int32_t) * (size_t)k);
    if (!g || !sel) return -2;
    for (int64_t i = 0; i < n; ++i) {
        int64_t s = indptr[i], e = indptr[i + 1], nz = e - s;
        for (int t = 0; t < k; ++t) {
            out_idx[i * k + t] = 0;
            out_val[i * k + t] = (T)1;
        }
        for (int64_t q = 0; q < nz; ++q) {
            T v = data[s + q];
            if (a) v = v * a[indices[s + q]];
            if (b) v = v + b[indices[s + q]];
            g[q] = (double)v;
        }
        if (nz > k) {
            orc_select_topk(g, nz, k, sel); /* positions, ascending == label ascending */
            for (int t = 0; t < k; ++t) {
                out_idx[i * k + t] = indices[s + sel[t]];
                if (keep_scores) out_val[i * k + t] = (T)g[sel[t]];
            }
        } else {
            for (int64_t q = 0; q < nz; ++q) {
                out_idx[i * k + q] = indices[s + q];
                if (keep_scores) out_val[i * k + q] = (T)g[q];
            }
        }
    }
    free(g);
    free(sel);
    return 0;
}

/* ---------------------------------------------------------------------------------
 * Label-wise confusion sums, dense.  Restates confusion_matrix.py:160-202 + numpy's
 * axis-0 reduction order (row after row, one running sum per column -- verified
 * bit-for-bit against numpy 2.3.5 by tests/golden/make_golden.py):
 *   tp = sum_i acc( T(y_true*y_pred) ), fp = sum_i acc( T((1-y_true)*y_pred) ),
 *   fn = sum_i acc( T(y_true*(1-y_pred)) )
 * acc_is_T != 0 mimics dtype=None (running sums kept in T); otherwise float64 sums.
 * Outputs are always returned as double (exact widening of T sums).
 * --------------------------------------------------------------------------------- */
int FN(orc_confmat_dense)(const T *y_true, int64_t ldt, const T *y_pred, int64_t ldp, int64_t n,
                          int64_t m, int acc_is_T, double *tp, double *fp, double *fn)
{
    if (acc_is_T) {
        T *stp = (T *)calloc((size_t)m, sizeof(T)), *sfp = (T *)calloc((size_t)m, sizeof(T)),
          *sfn = (T *)calloc((size_t)m, sizeof(T));
        if (!stp || !sfp || !sfn) return -2;
        for (int64_t i = 0; i < n; ++i)
            for (int64_t j = 0; j < m; ++j) {
                T y = y_true[i * ldt + j], p = y_pred[i * ldp + j];
                stp[j] = stp[j] + (T)(y * p);
                sfp[j] = sfp[j] + (T)(((T)1 - y) * p);
                sfn[j] = sfn[j] + (T)(y * ((T)1 - p));
            }
        for (int64_t j = 0; j < m; ++j) {
            tp[j] = (double)stp[j];
            fp[j] = (double)sfp[j];
            fn[j] = (double)sfn[j];
        }
        free(stp);
        free(sfp);
        free(sfn);
    } else {
        for (int64_t j = 0; j < m; ++j) tp[j] = fp[j] = fn[j] = 0.0;
        for (int64_t i = 0; i < n; ++i)
            for (int64_t j = 0; j < m; ++j) {
                T y = y_true[i * ldt + j], p = y_pred[i * ldp + j];
                tp[j] += (double)(T)(y * p);
                fp[j] += (double)(T)(((T)1 - y) * p);
                fn[j] += (double)(T)(y * ((T)1 - p));
            }
    }
    return 0;
}

/* Same sums with the prediction given compactly (n x k label ids, -1 = empty slot):
 * identical values because y_pred is 0/1 and adding an exact 0 is a no-op. */
int FN(orc_confmat_dense_compact)(const T *y_true, int64_t ldt, const int32_t *pred_idx, int k,
                                  int64_t n, int64_t m, double *tp, double *fp, double *fn)
{
    uint8_t *mark = (uint8_t *)calloc((size_t)m, 1);
    if (!mark) return -2;
    for (int64_t j = 0; j < m; ++j) tp[j] = fp[j] = fn[j] = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        for (int t = 0; t < k; ++t)
            if (pred_idx[i * k + t] >= 0) mark[pred_idx[i * k + t]] = 1;
        for (int64_t j = 0; j < m; ++j) {
            T y = y_true[i * ldt + j];
            if (mark[j]) {
                tp[j] += (double)y;
                fp[j] += (double)(T)((T)1 - y);
            } else {
                fn[j] += (double)y;
            }
        }
        for (int t = 0; t < k; ++t)
            if (pred_idx[i * k + t] >= 0) mark[pred_idx[i * k + t]] = 0;
    }
    free(mark);
    return 0;
}

/* ---------------------------------------------------------------------------------
 * CSR helpers: the two sorted-index merges of numba_csr_functions.py:116-140 (a*b on the
 * intersection, rounded to a's dtype) and :186-213 (a*(1.0-b) over all of a; numba
 * types "1.0 - b" as float64, the store into a's float32 buffer rounds once).
 * --------------------------------------------------------------------------------- */
static inline T FN(orc_mul_round)(T a, T b) { return (T)(a * b); }
static inline T FN(orc_mul_om_round)(T a, T b) { return (T)((double)a * (1.0 - (double)b)); }

/* Confusion sums for CSR (confusion_matrix.py:174-228, numba_csr_functions.py:144-182,
 * 217-258), axis 0, rows visited in order, scatter-add into acc (double or T).      */
int FN(orc_confmat_csr)(const T *t_data, const int32_t *t_idx, const int64_t *t_ptr,
                        const T *p_data, const int32_t *p_idx, const int64_t *p_ptr, int64_t n,
                        int64_t m, int acc_is_T, double *tp, double *fp, double *fn)
{
    T *stp = NULL, *sfp = NULL, *sfn = NULL;
    if (acc_is_T) {
        stp = (T *)calloc((size_t)m, sizeof(T));
        sfp = (T *)calloc((size_t)m, sizeof(T));
        sfn = (T *)calloc((size_t)m, sizeof(T));
        if (!stp || !sfp || !sfn) return -2;
    }
    for (int64_t j = 0; j < m; ++j) tp[j] = fp[j] = fn[j] = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        int64_t ts = t_ptr[i], te = t_ptr[i + 1], ps = p_ptr[i], pe = p_ptr[i + 1];
        /* tp: pred (a) * true (b) on the intersection */
        int64_t x = ps, y = ts;
        while (x < pe && y < te) {
            if (p_idx[x] < t_idx[y]) ++x;
            else if (p_idx[x] == t_idx[y]) {
                T v = FN(orc_mul_round)(p_data[x], t_data[y]);
                if (acc_is_T) stp[p_idx[x]] = stp[p_idx[x]] + v; else tp[p_idx[x]] += (double)v;
                ++x; ++y;
            } else ++y;
        }
        /* fp: pred (a) * (1 - true (b)) over all pred entries */
        x = ps; y = ts;
        while (x < pe) {
            T v;
            if (y >= te || p_idx[x] < t_idx[y]) { v = p_data[x]; }
            else if (p_idx[x] == t_idx[y]) { v = FN(orc_mul_om_round)(p_data[x], t_data[y]); ++y; }
            else { ++y; continue; }
            if (acc_is_T) sfp[p_idx[x]] = sfp[p_idx[x]] + v; else fp[p_idx[x]] += (double)v;
            ++x;
        }
        /* fn: true (a) * (1 - pred (b)) over all true entries */
        x = ts; y = ps;
        while (x < te) {
            T v;
            if (y >= pe || t_idx[x] < p_idx[y]) { v = t_data[x]; }
            else if (t_idx[x] == p_idx[y]) { v = FN(orc_mul_om_round)(t_data[x], p_data[y]); ++y; }
            else { ++y; continue; }
            if (acc_is_T) sfn[t_idx[x]] = sfn[t_idx[x]] + v; else fn[t_idx[x]] += (double)v;
            ++x;
        }
    }
    if (acc_is_T) {
        for (int64_t j = 0; j < m; ++j) { tp[j] = (double)stp[j]; fp[j] = (double)sfp[j]; fn[j] = (double)sfn[j]; }
        free(stp); free(sfp); free(sfn);
    }
    return 0;
}

/* ---------------------------------------------------------------------------------
 * One BCA sweep, dense rows, sequential (Gauss-Seidel) -- restates
 * block_coordinate.py:132-209 applied for i in order (:448-463).
 * pred is a dense n x m 0/1 byte matrix (the reference keeps a dense y_pred of dtype T).
 * State tp/fp/fn/tn: float64 un-normalised sums, updated in place.
 *   n_div  = n if normalize_conf_matrix else 1 (python int -> float64 divisor)
 *   greedy : skip the "remove own contribution" block (:157)
 * --------------------------------------------------------------------------------- */
int FN(orc_bca_dense_sweep)(const T *eta, int64_t n, int64_t m, int64_t ld, uint8_t *pred,
                            const int64_t *order, int64_t n_order, int k, int metric,
                            double c1, double beta2, double eps, double n_div, int skip_tn,
                            int maximize, int greedy, double *tp, double *fp, double *fn,
                            double *tn)
{
    (void)n;
    double *g = (double *)malloc(sizeof(double) * (size_t)m);
    int32_t *sel = (int32_t *)malloc(sizeof(int32_t) * (size_t)(k > 0 ? k : 1));
    if (!g || !sel) return -2;
    for (int64_t s = 0; s < n_order; ++s) {
        int64_t i = order[s];
        const T *row = eta + i * ld;
        uint8_t *yp = pred + i * m;
        for (int64_t j = 0; j < m; ++j) {
            T p = row[j];
            T om = (T)1 - p;                 /* (1 - y_proba_i) in T, :159 */
            T y = yp[j] ? (T)1 : (T)0;
            if (!greedy) {                   /* :157-163 */
                tp[j] -= (double)(T)(y * p);
                fp[j] -= (double)(T)(y * om);
                fn[j] -= (double)(T)(((T)1 - y) * p);
                if (!skip_tn) tn[j] -= (double)(T)(((T)1 - y) * om);
            }
            double pos_tp = tp[j] + (double)p;      /* :166-172 */
            double pos_fp = fp[j] + (double)om;
            double neg_fn = fn[j] + (double)p;
            double neg_tn = skip_tn ? tn[j] : tn[j] + (double)om;
            double up = orc_binary_metric(metric, pos_tp / n_div, pos_fp / n_div, fn[j] / n_div,
                                          tn[j] / n_div, c1, beta2, eps);
            double un = orc_binary_metric(metric, tp[j] / n_div, fp[j] / n_div, neg_fn / n_div,
                                          neg_tn / n_div, c1, beta2, eps);
            double gain = up - un;           /* :129 */
            g[j] = maximize ? gain : -gain;  /* we pick the LARGEST of g (:187-195) */
        }
        if (k > 0) {
            orc_select_topk(g, m, k, sel);
            memset(yp, 0, (size_t)m);
            for (int t = 0; t < k; ++t) yp[sel[t]] = 1;
        } else {
            for (int64_t j = 0; j < m; ++j) yp[j] = (g[j] >= 0.0) ? 1 : 0; /* :200 */
        }
        for (int64_t j = 0; j < m; ++j) {    /* :203-209 */
            T p = row[j];
            T om = (T)1 - p;
            T y = yp[j] ? (T)1 : (T)0;
            tp[j] += (double)(T)(y * p);
            fp[j] += (double)(T)(y * om);
            fn[j] += (double)(T)(((T)1 - y) * p);
            if (!skip_tn) tn[j] += (double)(T)(((T)1 - y) * om);
        }
    }
    free(g);
    free(sel);
    return 0;
}

/* ---------------------------------------------------------------------------------
 * One BCA sweep, CSR rows, sequential -- restates block_coordinate.py:212-293 with
 * numba_csr_functions.py:386-452 (sub/add), :456-466 (top-k of the stored entries),
 * :500-546 (row replace).  pred is n x k label ids per row, ascending, with
 * pred_len[i] <= k valid entries (rows with nnz_i < k keep all nnz_i labels, :465-466).
 * Prediction values are all 1 (T).  tn == NULL means skip_tn (tn_const is passed to the metric);
 * otherwise tn is updated literally like numba does: +-1 on ALL m labels, then the three products
 * on the touched ones (numba_csr_functions.py:413-417, :448-452).
 * --------------------------------------------------------------------------------- */
int FN(orc_bca_csr_sweep)(const T *data, const int32_t *indices, const int64_t *indptr,
                          int64_t n, int64_t m, int32_t *pred_idx, int32_t *pred_len,
                          const int64_t *order, int64_t n_order, int k, int metric, double c1,
                          double beta2, double eps, double n_div, int maximize, int greedy,
                          double *tp, double *fp, double *fn, double *tn, double tn_const)
{
    (void)n; (void)m;
    int64_t cap = 0;
    for (int64_t s = 0; s < n_order; ++s) {
        int64_t i = order[s];
        if (indptr[i + 1] - indptr[i] > cap) cap = indptr[i + 1] - indptr[i];
    }
    double *g = (double *)malloc(sizeof(double) * (size_t)(cap + 1));
    int32_t *sel = (int32_t *)malloc(sizeof(int32_t) * (size_t)k);
    if (!g || !sel) return -2;
    const T one = (T)1;
    for (int64_t s = 0; s < n_order; ++s) {
        int64_t i = order[s];
        int64_t ts = indptr[i], te = indptr[i + 1], nz = te - ts;
        int32_t *pi = pred_idx + i * k;
        for (int pass = 0; pass < 2; ++pass) {
            /* pass 0: remove the row's contribution (:243-246), pass 1: add it back with
             * the new prediction (:290-293).                                            */
            if (pass == 0 && greedy) goto gains;
            {
                double sgn = pass == 0 ? -1.0 : 1.0;
                int np_ = pred_len[i];
                int64_t x = 0, y = ts;
                if (tn) for (int64_t j = 0; j < m; ++j) tn[j] += sgn;   /* tn -= 1 / tn += 1 on every label */
                while (x < np_ && y < te) { /* tp on the intersection */
                    if (pi[x] < indices[y]) ++x;
                    else if (pi[x] == indices[y]) {
                        double v = (double)FN(orc_mul_round)(one, data[y]);
                        tp[pi[x]] += sgn * v;
                        if (tn) tn[pi[x]] -= sgn * v;
                        ++x; ++y;
                    }
                    else ++y;
                }
                x = 0; y = ts;
                while (x < np_) { /* fp over pred entries */
                    if (y >= te || pi[x] < indices[y]) { fp[pi[x]] += sgn * (double)one; if (tn) tn[pi[x]] -= sgn * (double)one; ++x; }
                    else if (pi[x] == indices[y]) {
                        double v = (double)FN(orc_mul_om_round)(one, data[y]);
                        fp[pi[x]] += sgn * v;
                        if (tn) tn[pi[x]] -= sgn * v;
                        ++x; ++y;
                    }
                    else ++y;
                }
                y = 0;
                for (int64_t q = ts; q < te; ++q) { /* fn over true entries */
                    while (y < np_ && pi[y] < indices[q]) ++y;
                    double v = (y < np_ && pi[y] == indices[q]) ? (double)FN(orc_mul_om_round)(data[q], one)
                                                                : (double)data[q];
                    fn[indices[q]] += sgn * v;
                    if (tn) tn[indices[q]] -= sgn * v;
                }
            }
            if (pass == 1) break;
        gains:
            for (int64_t q = 0; q < nz; ++q) { /* :248-282 */
                int32_t j = indices[ts + q];
                T t = data[ts + q];
                T om = one - t;
                double neg_tp = tp[j], neg_fp = fp[j], pos_fn = fn[j];
                double pos_tpp = (neg_tp + (double)t) / n_div;
                double pos_fpp = (neg_fp + (double)om) / n_div;
                double neg_fnn = (pos_fn + (double)t) / n_div;
                neg_tp /= n_div; neg_fp /= n_div; pos_fn /= n_div;
                double pos_tn = tn_const, neg_tnn = tn_const;   /* block_coordinate.py:260-264 */
                if (tn) {
                    pos_tn = tn[j];
                    neg_tnn = (pos_tn + (double)om) / n_div;
                    pos_tn /= n_div;
                }
                double up = orc_binary_metric(metric, pos_tpp, pos_fpp, pos_fn, pos_tn, c1, beta2, eps);
                double un = orc_binary_metric(metric, neg_tp, neg_fp, neg_fnn, neg_tnn, c1, beta2, eps);
                double gain = up - un;
                g[q] = maximize ? gain : -gain;
            }
            if (nz > k) {
                orc_select_topk(g, nz, k, sel);
                for (int t = 0; t < k; ++t) pi[t] = indices[ts + sel[t]];
                pred_len[i] = k;
            } else {
                for (int64_t q = 0; q < nz; ++q) pi[q] = indices[ts + q];
                for (int64_t q = nz; q < k; ++q) pi[q] = -1;
                pred_len[i] = (int32_t)nz;
            }
        }
    }
    free(g);
    free(sel);
    return 0;
}

/* ---------------------------------------------------------------------------------
 * Coverage: failure-probability state Ef[j] = prod_i (1 - yhat_ij * eta_ij).
 * Recompute restates numba_csr_functions.py:325-382 as called at
 * block_coordinate.py:665/676 (a = y_pred, b = y_proba), rows in order.
 * --------------------------------------------------------------------------------- */
int FN(orc_cov_state_csr)(const T *data, const int32_t *indices, const int64_t *indptr, int64_t n,
                          int64_t m, const int32_t *pred_idx, const int32_t *pred_len, int k,
                          double *Ef)
{
    const T one = (T)1;
    for (int64_t j = 0; j < m; ++j) Ef[j] = 1.0;
    for (int64_t i = 0; i < n; ++i) {
        int64_t y = indptr[i], te = indptr[i + 1];
        const int32_t *pi = pred_idx + i * k;
        for (int x = 0; x < pred_len[i]; ++x) {
            while (y < te && indices[y] < pi[x]) ++y;
            if (y < te && indices[y] == pi[x]) Ef[pi[x]] *= (double)FN(orc_mul_om_round)(one, data[y]);
            else Ef[pi[x]] *= (double)one;
        }
    }
    return 0;
}

/* One coverage sweep over CSR rows -- restates block_coordinate.py:539-580.
 * Rows with nnz <= k are padded with label 0 then sorted (:573-576); those padded ids
 * are kept in pred_idx exactly as the reference stores them.                          */
int FN(orc_cov_csr_sweep)(const T *data, const int32_t *indices, const int64_t *indptr, int64_t n,
                          int64_t m, int32_t *pred_idx, const int64_t *order, int64_t n_order,
                          int k, double alpha, int greedy, double *Ef)
{
    (void)n; (void)m;
    int64_t cap = 0;
    for (int64_t s = 0; s < n_order; ++s) {
        int64_t i = order[s];
        if (indptr[i + 1] - indptr[i] > cap) cap = indptr[i + 1] - indptr[i];
    }
    double *g = (double *)malloc(sizeof(double) * (size_t)(cap + 1));
    int32_t *sel = (int32_t *)malloc(sizeof(int32_t) * (size_t)k);
    if (!g || !sel) return -2;
    const T one = (T)1;
    for (int64_t s = 0; s < n_order; ++s) {
        int64_t i = order[s];
        int64_t ts = indptr[i], te = indptr[i + 1], nz = te - ts;
        int32_t *pi = pred_idx + i * k;
        if (!greedy) { /* :562-564  Ef[idx] /= 1 - T(p*t) */
            int64_t x = 0, y = ts;
            while (x < k && y < te) {
                if (pi[x] < indices[y]) ++x;
                else if (pi[x] == indices[y]) { Ef[pi[x]] /= (double)(T)(one - FN(orc_mul_round)(one, data[y])); ++x; ++y; }
                else ++y;
            }
        }
        for (int64_t q = 0; q < nz; ++q) { /* :567-569 */
            T t = data[ts + q];
            double gain = Ef[indices[ts + q]] * (double)t;
            if (alpha < 1.0) {
                T w = (T)((T)(1.0 - alpha) * t);  /* python float is a weak scalar -> T */
                w = (T)(w / (T)k);
                gain = alpha * gain + (double)w;
            }
            g[q] = gain;
        }
        if (nz > k) { /* :570-572 */
            orc_select_topk(g, nz, k, sel);
            for (int t = 0; t < k; ++t) pi[t] = indices[ts + sel[t]];
        } else {      /* :573-576: resize to k, fill with 0, sort */
            int64_t pad = k - nz;
            for (int64_t q = 0; q < pad; ++q) pi[q] = 0;
            for (int64_t q = 0; q < nz; ++q) pi[pad + q] = indices[ts + q];
        }
        { /* :579-580  Ef[idx] *= 1 - T(p*t) */
            int64_t x = 0, y = ts;
            while (x < k && y < te) {
                if (pi[x] < indices[y]) ++x;
                else if (pi[x] == indices[y]) { Ef[pi[x]] *= (double)(T)(one - FN(orc_mul_round)(one, data[y])); ++x; ++y; }
                else ++y;
            }
        }
    }
    free(g);
    free(sel);
    return 0;
}

#undef FN
#undef CAT
#undef CAT_
