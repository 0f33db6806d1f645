/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the prediction-optimisation hot path of
 * mwydmuch/xCOLUMNs 0.0.3 (the reference; line numbers below are relative to
 * /root/reference).  It exists so that tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs can CHECK the CUDA path; nothing
 * under xcolumns_b200/ may import, link or call it.
 *
 * Parity status: PINNED.  tests/golden/make_golden.py imported the live reference in
 * the build container (numpy 2.3.5, numba 0.65, scipy 1.18) and stored its outputs
 * under tests/golden/ (npz files); tests/test_oracle_golden.py replays them against this
 * file bit-for-bit (label indices, float64 state, per-sweep utilities).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no FMA contraction, IEEE
 * float/double arithmetic exactly as numpy/numba perform it on x86-64).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Metric ids shared with include/xcolumns_b200.h (XC_METRIC_*). */
enum {
    ORC_PRECISION = 0,
    ORC_RECALL = 1,
    ORC_FBETA = 2,
    ORC_JACCARD = 3,
    ORC_BALANCED_ACC = 4,
    ORC_GMEAN = 5,
    ORC_HMEAN = 6,
    ORC_PREC_AT_K = 7 /* tp / k (metrics.py:497-513) */
};

/* Binary metrics on one label's (tp, fp, fn, tn), float64, operation order of the
 * python expressions in xcolumns/metrics.py:
 *   precision :603   tp / (tp + fp + epsilon)
 *   recall    :652   tp / (tp + fn + epsilon)
 *   fbeta     :703   (1 + beta**2) * tp / ((beta**2 * (tp + fp)) + tp + fn + epsilon)
 *   jaccard   :797   tp / (tp + fp + fn + epsilon)
 *   bal. acc  :843-845, gmean :890-892 ((tpr*tnr)**0.5 == sqrt), hmean :939-941
 * c1 = 1 + beta**2 and beta2 = beta**2 are computed by the python caller.            */
/* Mixed utilities of block_coordinate.py:848-1045 / frank_wolfe.py:838-915:
 *     (1 - alpha) * binary_precision_at_k(tp, k) + alpha * binary_metric(...) / m
 * = ((1 - alpha) * (tp / k)) + ((alpha * metric) / m) in python's evaluation order.  The mix is
 * process-wide test state (orc_set_mix) so that every restated step picks it up unchanged.   */
static int g_mix_on = 0;
static double g_mix_alpha = 1.0, g_mix_k = 1.0, g_mix_m = 1.0;
void orc_set_mix(int on, double alpha, double k, double m)
{
    g_mix_on = on;
    g_mix_alpha = alpha;
    g_mix_k = k;
    g_mix_m = m;
}

static inline double orc_base_metric(int metric, double tp, double fp, double fn, double tn,
                                     double c1, double beta2, double eps);

static inline double orc_binary_metric(int metric, double tp, double fp, double fn, double tn,
                                       double c1, double beta2, double eps)
{
    double v = orc_base_metric(metric, tp, fp, fn, tn, c1, beta2, eps);
    if (g_mix_on) v = ((1.0 - g_mix_alpha) * (tp / g_mix_k)) + ((g_mix_alpha * v) / g_mix_m);
    return v;
}

static inline double orc_base_metric(int metric, double tp, double fp, double fn, double tn,
                                     double c1, double beta2, double eps)
{
    switch (metric) {
    case ORC_PREC_AT_K: return tp / c1; /* metrics.py:513, c1 carries k */
    case ORC_PRECISION: return tp / ((tp + fp) + eps);
    case ORC_RECALL: return tp / ((tp + fn) + eps);
    case ORC_FBETA: return (c1 * tp) / ((((beta2 * (tp + fp)) + tp) + fn) + eps);
    case ORC_JACCARD: return tp / (((tp + fp) + fn) + eps);
    default: {
        double tpr = tp / ((tp + fn) + eps);
        double tnr = tn / ((tn + fp) + eps);
        if (metric == ORC_BALANCED_ACC) return (tpr + tnr) / 2.0;
        if (metric == ORC_GMEAN) return sqrt(tpr * tnr);
        return ((2.0 * tpr) * tnr) / (tpr + tnr); /* ORC_HMEAN */
    }
    }
}

/* positions of the k largest values of g[0..len), ties -> lowest position; result
 * ascending by position.  (Reference: np.argpartition, tie order unspecified --
 * parity is only claimed on inputs without ties at the k-th boundary.)              */
static void orc_select_topk(const double *g, int64_t len, int k, int32_t *sel)
{
    /* sel kept sorted by (value desc, position asc): sel[k-1] is the current worst */
    int cnt = 0;
    for (int64_t j = 0; j < len; ++j) {
        double v = g[j];
        if (cnt == k) {
            double w = g[sel[k - 1]];
            if (!(v > w)) continue; /* equal value, higher position loses */
        }
        int pos = cnt < k ? cnt : k - 1;
        while (pos > 0 && g[sel[pos - 1]] < v) { sel[pos] = sel[pos - 1]; --pos; }
        sel[pos] = (int32_t)j;
        if (cnt < k) ++cnt;
    }
    /* sort by position */
    for (int a = 1; a < cnt; ++a) {
        int32_t x = sel[a];
        int b = a - 1;
        while (b >= 0 && sel[b] > x) { sel[b + 1] = sel[b]; --b; }
        sel[b + 1] = x;
    }
}

#define T float
#define SFX f32
#include "xc_oracle_impl.h"
#undef T
#undef SFX

#define T double
#define SFX f64
#include "xc_oracle_impl.h"
#undef T
#undef SFX

/* per-label metric vector (python side takes .mean()/.sum() with numpy so the
 * pairwise summation order of the reference is reproduced, block_coordinate.py:54-90) */
void orc_binary_metric_vec(int metric, const double *tp, const double *fp, const double *fn,
                           const double *tn, int64_t m, double div, double c1, double beta2,
                           double eps, double *out)
{
    for (int64_t j = 0; j < m; ++j)
        out[j] = orc_binary_metric(metric, tp[j] / div, fp[j] / div, fn[j] / div, tn[j] / div, c1,
                                   beta2, eps);
}

/* Frank-Wolfe line search -- restates utils.py:174-184 (uniform grid, first strict
 * maximum wins) over frank_wolfe.py:393-398 for a macro-averaged built-in metric, all in
 * float64 with a plain left-to-right label sum (the reference sums pairwise and, for
 * float32 confusion vectors, evaluates alpha=0 in float32: agreement is to ~1e-12, the
 * FW contract is 1e-4).  alphas[] is the grid the caller built with numpy
 * (np.arange(low+step, high, step)), evaluated after alpha = 0.                       */
void orc_fw_alpha_search(int metric, const double *tp, const double *fp, const double *fn,
                         const double *tn, const double *tp_i, const double *fp_i,
                         const double *fn_i, const double *tn_i, int64_t m, const double *alphas,
                         int64_t n_alphas, double c1, double beta2, double eps,
                         double *best_alpha, double *best_val, int skip_zero)
{
    /* skip_zero: evaluate only the listed points (single-point evaluation for the ternary search) */
    double best = 0.0, bv = 0.0;
    for (int64_t q = skip_zero ? 0 : -1; q < n_alphas; ++q) {
        double al = q < 0 ? 0.0 : alphas[q];
        double s = 0.0;
        for (int64_t j = 0; j < m; ++j) {
            double a = (1.0 - al) * tp[j] + al * tp_i[j];
            double b = (1.0 - al) * fp[j] + al * fp_i[j];
            double c = (1.0 - al) * fn[j] + al * fn_i[j];
            double d = (1.0 - al) * tn[j] + al * tn_i[j];
            s += orc_binary_metric(metric, a, b, c, d, c1, beta2, eps);
        }
        s /= (double)m;
        if (q < 0 || (skip_zero && q == 0) || s > bv) { bv = s; best = al; } /* always a maximum, also when the FW
                                                        driver minimises (frank_wolfe.py:616) */
    }
    *best_alpha = best;
    *best_val = bv;
}
