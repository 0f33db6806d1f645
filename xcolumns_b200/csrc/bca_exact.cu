// Sequential-exact BCA sweeps: the reference's instance order and float64 operation order.
//
// Dense rows (xcolumns/block_coordinate.py:132-209): one cooperative launch per sweep.  The m
// labels are spread over the whole grid, each thread keeps the float64 state (tp, fp, fn, tn)
// of its labels in registers for the whole sweep, and every instance costs ONE grid barrier:
//   1. remove the row's contribution, evaluate psi(+row) - psi(-row) per label   (registers)
//   2. block-level top-k  -> k candidates per block in a double-buffered exchange array
//   3. grid barrier
//   4. every block merges all candidates (same deterministic result everywhere), re-adds the
//      row with the new selection to its own labels; block 0 stores the new prediction row.
// eta of the next instance is prefetched before the barrier.  The chain is latency bound
// (~2-3 us per instance) by construction: instance i+1 needs instance i's commit.
//
// CSR rows (:212-293) and coverage (:539-580): the candidates of a step are only the row's
// stored labels, so a single warp walks the order; the state lives in L2.
//
// This file must be compiled with -fmad=false: every + - * / below is a separate IEEE
// operation, exactly like numpy/numba evaluate the reference expressions.
#include <cooperative_groups.h>
#include <cstdlib>

#include "xc_scan.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int EX_THREADS = 128;
constexpr int EX_WARPS = EX_THREADS / 32;

// Candidate = (gain, label).  The gain is carried as an order-preserving int64 key of the double
// (FP64 compares and selects run on the scarce FP64 pipe of this chip -- ~16 lanes/clk/SM measured --
// and every arg-max round is a chain of them; integer compares are 4x cheaper and 2-cycle issue).
struct alignas(16) Cand {
    long long g;   // sortable key of the float64 gain
    int j;
    int pad;
};

__device__ __forceinline__ long long gain_key(double x)
{
    if (!(x == x)) return (long long)0x8000000000000000ULL;          // NaN sorts below everything
    long long b = __double_as_longlong(x + 0.0);                      // -0.0 -> +0.0
    return b >= 0 ? b : (b ^ 0x7fffffffffffffffLL);                   // negative: reverse magnitude order
}
constexpr long long KEY_NEG_INF = (long long)0x8000000000000001ULL;   // below every real gain, above NaN key

// candidates written by other blocks: read through L2 (L1 is not coherent across SMs)
__device__ __forceinline__ Cand cand_load_cg(const Cand *p)
{
    longlong2 v = __ldcg(reinterpret_cast<const longlong2 *>(p));
    Cand c;
    c.g = v.x;
    c.j = (int)(v.y & 0xffffffffLL);
    c.pad = 0;
    return c;
}

__device__ __forceinline__ Cand cand_best(Cand a, Cand b) { return xc_better(b.g, b.j, a.g, a.j) ? b : a; }  // int64 keys

__device__ __forceinline__ Cand warp_argmax(Cand c)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Cand t;
        t.g = __shfl_xor_sync(XC_FULL, c.g, o);   // long long shuffle (2 x 32 bit)
        t.j = __shfl_xor_sync(XC_FULL, c.j, o);
        t.pad = 0;
        c = cand_best(c, t);
    }
    return c;
}

// best candidate of the block, returned to every thread; sm has EX_WARPS entries
__device__ __forceinline__ Cand block_argmax(Cand c, Cand *sm)
{
    c = warp_argmax(c);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
    __syncthreads();
    Cand r;
    const int lane = threadIdx.x & 31;
    if (lane < EX_WARPS) r = sm[lane];
    else { r.g = KEY_NEG_INF; r.j = 0x7fffffff; }
    r.pad = 0;
    r = warp_argmax(r);
    __syncthreads();
    return r;
}

template <typename TE, int L>
__global__ void __launch_bounds__(EX_THREADS)
bca_exact_dense_kernel(const TE *__restrict__ eta, int64_t m, int64_t ld, const int32_t *__restrict__ order,
                       int64_t n_order, int k, xc_metric_params p, int greedy, int32_t *pred_idx, double *tp,
                       double *fp, double *fn, double *tn, Cand *cand)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ Cand sm[EX_WARPS];
    __shared__ int s_sel[32];
    const int64_t stride = (int64_t)gridDim.x * EX_THREADS;
    const int64_t j0 = (int64_t)blockIdx.x * EX_THREADS + threadIdx.x;
    const int nblk = gridDim.x;
    const bool use_tn = !p.skip_tn;
    const double nd = p.n_div;
    const TE one = (TE)1;

    double stp[L], sfp[L], sfn[L], stn[L];
    TE pv[L], pnext[L];
#pragma unroll
    for (int l = 0; l < L; ++l) {
        int64_t j = j0 + l * stride;
        bool ok = j < m;
        stp[l] = ok ? tp[j] : 0.0;
        sfp[l] = ok ? fp[j] : 0.0;
        sfn[l] = ok ? fn[j] : 0.0;
        stn[l] = ok ? tn[j] : 0.0;
        pnext[l] = (TE)0;
    }
    if (n_order > 0) {
        const TE *rp = eta + (int64_t)order[0] * ld;
#pragma unroll
        for (int l = 0; l < L; ++l) {
            int64_t j = j0 + l * stride;
            if (j < m) pnext[l] = __ldg(rp + j);
        }
    }

    for (int64_t s = 0; s < n_order; ++s) {
        const int64_t row = order[s];
        int32_t *prow = pred_idx + row * k;
#pragma unroll
        for (int l = 0; l < L; ++l) pv[l] = pnext[l];
        if (s + 1 < n_order) {  // prefetch the next instance's probabilities
            const TE *rn = eta + (int64_t)order[s + 1] * ld;
#pragma unroll
            for (int l = 0; l < L; ++l) {
                int64_t j = j0 + l * stride;
                if (j < m) pnext[l] = __ldg(rn + j);
            }
        }
        // ---- 1. remove own contribution, gains -----------------------------------------------
        double gain[L];
#pragma unroll
        for (int l = 0; l < L; ++l) {
            const int64_t j = j0 + l * stride;
            gain[l] = -INFINITY;
            if (j < m) {
                bool sel = false;
                for (int t = 0; t < k; ++t) sel |= (prow[t] == (int)j);
                const TE pe = pv[l];
                const TE om = one - pe;           // (1 - y_proba_i) in TE          :159
                const TE y = sel ? one : (TE)0;
                if (!greedy) {                    //                                :157-163
                    stp[l] = stp[l] - (double)(TE)(y * pe);
                    sfp[l] = sfp[l] - (double)(TE)(y * om);
                    sfn[l] = sfn[l] - (double)(TE)((one - y) * pe);
                    if (use_tn) stn[l] = stn[l] - (double)(TE)((one - y) * om);
                }
                const double pos_tp = stp[l] + (double)pe;   //                     :166-172
                const double pos_fp = sfp[l] + (double)om;
                const double neg_fn = sfn[l] + (double)pe;
                const double neg_tn = use_tn ? stn[l] + (double)om : stn[l];
                const double up = xc_metric_eval(p, pos_tp / nd, pos_fp / nd, sfn[l] / nd, stn[l] / nd);
                const double un = xc_metric_eval(p, stp[l] / nd, sfp[l] / nd, neg_fn / nd, neg_tn / nd);
                const double g = up - un;         //                                :129
                gain[l] = p.maximize ? g : -g;
            }
        }
        // ---- 2. block-level top-k ---------------------------------------------------------------
        Cand *slot = cand + ((s & 1) * (int64_t)nblk + blockIdx.x) * k;
        unsigned taken = 0;
        for (int t = 0; t < k; ++t) {
            Cand c;
            c.g = KEY_NEG_INF;
            c.j = 0x7fffffff;
            c.pad = 0;
#pragma unroll
            for (int l = 0; l < L; ++l) {
                int64_t j = j0 + l * stride;
                if (j < m && !((taken >> l) & 1u)) {
                    Cand d;
                    d.g = gain_key(gain[l]);
                    d.j = (int)j;
                    d.pad = 0;
                    c = cand_best(c, d);
                }
            }
            Cand w = block_argmax(c, sm);
#pragma unroll
            for (int l = 0; l < L; ++l)
                if ((int64_t)w.j == j0 + l * stride) taken |= 1u << l;
            if (threadIdx.x == 0) slot[t] = w;
        }
        // ---- 3. exchange --------------------------------------------------------------------------
        __threadfence();
        grid.sync();
        // ---- 4. merge all candidates (identical in every block) ----------------------------------
        const Cand *all = cand + (s & 1) * (int64_t)nblk * k;
        const int total = nblk * k;
        // every thread scans a strided slice; "taken" by label id comparison
        int picked[32];
        for (int t = 0; t < k; ++t) {
            Cand c;
            c.g = KEY_NEG_INF;
            c.j = 0x7fffffff;
            c.pad = 0;
            for (int q = threadIdx.x; q < total; q += EX_THREADS) {
                Cand d = cand_load_cg(all + q);
                bool used = false;
                for (int u = 0; u < t; ++u) used |= (picked[u] == d.j);
                if (!used) c = cand_best(c, d);
            }
            Cand w = block_argmax(c, sm);
            picked[t] = w.j;
        }
        if (threadIdx.x < 32) s_sel[threadIdx.x] = threadIdx.x < k ? picked[threadIdx.x] : 0x7fffffff;
        // ---- 5. re-add with the new selection -------------------------------------------------------
#pragma unroll
        for (int l = 0; l < L; ++l) {
            const int64_t j = j0 + l * stride;
            if (j < m) {
                bool sel = false;
                for (int t = 0; t < k; ++t) sel |= (picked[t] == (int)j);
                const TE pe = pv[l];
                const TE om = one - pe;
                const TE y = sel ? one : (TE)0;
                stp[l] = stp[l] + (double)(TE)(y * pe);      //                     :203-209
                sfp[l] = sfp[l] + (double)(TE)(y * om);
                sfn[l] = sfn[l] + (double)(TE)((one - y) * pe);
                if (use_tn) stn[l] = stn[l] + (double)(TE)((one - y) * om);
            }
        }
        if (blockIdx.x == 0) {
            __syncthreads();
            if (threadIdx.x < 32) {  // store ascending by label
                int mine = s_sel[threadIdx.x];
                int src = warp_rank_src(mine, k);
                int v = __shfl_sync(XC_FULL, mine, src);
                if (threadIdx.x < k) prow[threadIdx.x] = v;
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int l = 0; l < L; ++l) {
        int64_t j = j0 + l * stride;
        if (j < m) {
            tp[j] = stp[l];
            fp[j] = sfp[l];
            fn[j] = sfn[l];
            if (use_tn) tn[j] = stn[l];
        }
    }
}

// ---- dense, one thread-block cluster ---------------------------------------------------------------
// Same arithmetic as bca_exact_dense_kernel, but the labels live in ONE cluster of up to 16 CTAs, so the
// per-instance exchange goes through distributed shared memory and a cluster barrier (~0.2 us) instead
// of global memory and a grid barrier (~3-5 us):
//   warp top-k (k rounds of shuffle arg-max) -> block top-k in smem -> every CTA pushes its k candidates
//   into all peers' smem (DSMEM stores) -> barrier.cluster -> every CTA merges the <= 16*k candidates.
constexpr int CL_MAX = 16;

// ---- candidate exchange through asynchronous DSMEM stores that complete on the RECEIVER's mbarrier -----------------
// (st.async ... mbarrier::complete_tx::bytes): the sender does not wait for its remote stores, the receiver waits for
// the bytes it expects.  The first version used plain DSMEM stores + barrier.cluster: its release fence (ERRBAR in the
// SASS) made every CTA wait for the round trip of its own remote stores before it could even arrive -- 8 % of all
// warp samples of the sweep plus 5 % in the barrier itself (ncu, profiles/r02_notes.md section 7).
__device__ __forceinline__ uint32_t smem_addr_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_addr_of(uint32_t local_smem_addr, uint32_t cta_rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void st_async_16(uint32_t remote_addr, long long a, long long b, uint32_t remote_mbar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];"
                 :: "r"(remote_addr), "l"(a), "l"(b), "r"(remote_mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t mbar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint32_t mbar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
    } while (!done);
}

// top-k of `cnt` candidates in smem by ranking: thread t < cnt counts the candidates that beat
// candidate t (cnt broadcast-friendly smem reads, no shuffle chains, no serial rounds) and, if fewer
// than k do, stores it at dst[rank].  All threads of the CTA call it; the caller synchronises.
// Measured motivation (profiles/r01_notes.md): with k rounds of warp arg-max run by warp 0 alone, 47 %
// of all warp samples of the sweep were other warps parked at the barriers behind those merges.
// The caller has filled dst[0 .. k) with empty candidates (slots no candidate claims -- fewer than k real
// candidates -- must read as empty) and has put a block barrier between that fill and this call: the fill used to
// live here behind a barrier of its own, two of the six block barriers of an instance.
// Four threads share one candidate: each counts the better ones among a quarter of the list, two shuffles add the
// partial ranks.  With cnt = 20 .. 40 candidates the old "thread t ranks candidate t" left one or two warps
// walking the whole list while the rest of the CTA sat at the barrier behind them (ncu source page,
// profiles/r02_notes.md section 7: 16 % of all warp samples of the sweep were that wait).
__device__ __forceinline__ void rank_topk_smem(const Cand *src, int cnt, int k, Cand *dst)
{
    const int lane = threadIdx.x & 31;
    const int total = cnt * 4;
    for (int base = (int)threadIdx.x - lane; base < total; base += (int)blockDim.x) {   // warp-uniform trip count
        const int u = base + lane;
        const bool active = u < total;
        const int c = active ? (u >> 2) : 0, h = u & 3;
        const Cand me = src[c];
        int rank = 0;
        if (active) {
            for (int q = h; q < cnt; q += 4) {
                const Cand o = src[q];
                rank += xc_better(o.g, o.j, me.g, me.j) ? 1 : 0;   // strict order: labels are distinct
            }
        }
        rank += __shfl_xor_sync(XC_FULL, rank, 1);
        rank += __shfl_xor_sync(XC_FULL, rank, 2);
        if (active && h == 0 && rank < k) dst[rank] = me;
    }
}

template <typename TE, int L, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
bca_exact_dense_cluster_kernel(const TE *__restrict__ eta, int64_t m, int64_t ld, const int32_t *__restrict__ order,
                               int64_t n_order, int k, xc_metric_params p, int greedy, int32_t *pred_idx, double *tp,
                               double *fp, double *fn, double *tn, const TE *__restrict__ addback, int64_t ld_add)
{
    // order == nullptr: rows 0 .. n_order-1 in sequence.  addback != nullptr (online mode, greedy): the state
    // is updated with the row of `addback` (the true labels) instead of the probabilities themselves
    // (confusion_matrix.py:402-435 after a step with only_pred=True).
    cg::cluster_group cluster = cg::this_cluster();
    const int nc = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    constexpr int NW = THREADS / 32;
    __shared__ Cand s_blk[32];                // block top-k
    __shared__ Cand s_in[2][CL_MAX * 32];     // candidates pushed by every CTA of the cluster (double buffered)
    __shared__ Cand s_fin[32];                // final top-k of the step
    __shared__ Cand s_tmp[NW * 32];           // per-warp top-k, densely packed: warp w owns [w * k, (w + 1) * k)
    __shared__ Cand s_keys[L == 1 ? NW * 32 : 1];   // per-warp key exchange (L == 1)
    __shared__ alignas(8) unsigned long long s_mbar[2];   // one mbarrier per input buffer
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t stride = (int64_t)nc * THREADS;
    const int64_t j0 = (int64_t)rank * THREADS + threadIdx.x;
    const bool use_tn = !p.skip_tn;
    const double nd = p.n_div;
    const TE one = (TE)1;

    double stp[L], sfp[L], sfn[L], stn[L];
    TE pv[L], pnext[L], av[L], anext[L];
#pragma unroll
    for (int l = 0; l < L; ++l) {
        int64_t j = j0 + l * stride;
        bool ok = j < m;
        stp[l] = ok ? tp[j] : 0.0;
        sfp[l] = ok ? fp[j] : 0.0;
        sfn[l] = ok ? fn[j] : 0.0;
        stn[l] = ok ? tn[j] : 0.0;
        pnext[l] = (TE)0;
        anext[l] = (TE)0;
    }
    if (n_order > 0) {
        const int64_t r0 = order ? (int64_t)order[0] : 0;
#pragma unroll
        for (int l = 0; l < L; ++l) {
            int64_t j = j0 + l * stride;
            if (j < m) {
                pnext[l] = __ldg(eta + r0 * ld + j);
                if (addback) anext[l] = __ldg(addback + r0 * ld_add + j);
            }
        }
    }
    // the row's current selection is prefetched one step ahead like the row itself (a row is visited once per
    // sweep, so nobody rewrites it before its turn)
    int myp_next = -1;
    if (n_order > 0 && lane < k) myp_next = __ldg(pred_idx + (order ? (int64_t)order[0] : 0) * k + lane);
    // The visiting order is read TWO steps ahead: an in-order warp stalls at the first consumer of a load, and with
    // one warp per scheduler nothing else can issue meanwhile (ncu, profiles/r02_exact_cluster_source.csv: 6 % of
    // all samples sat on `row = order[s]`, 2 % on the address arithmetic behind `order[s + 1]`).
    // (kept as the raw 32-bit value: widening it right behind the load would make THAT the first consumer)
    int64_t row_cur = n_order > 0 ? (order ? (int64_t)order[0] : 0) : 0;
    int row_n1 = n_order > 1 ? (order ? order[1] : 1) : 0;
    // the index arithmetic of the DSMEM push (q / k, q % k: ~25 instructions each with a run-time k) is loop
    // invariant for a thread's first -- and, unless nc * k > THREADS, only -- candidate
    const int push_dst0 = (int)threadIdx.x / k, push_r0 = (int)threadIdx.x % k;
    Cand c_empty;
    c_empty.g = KEY_NEG_INF; c_empty.j = 0x7fffffff; c_empty.pad = 0;
    double stn_n[L];   // used with skip_tn only: tn is never updated then, neither is its quotient
#pragma unroll
    for (int l = 0; l < L; ++l) stn_n[l] = stn[l] / nd;
    const uint32_t mbar0 = smem_addr_u32(&s_mbar[0]), mbar1 = smem_addr_u32(&s_mbar[1]);
    const uint32_t in0 = smem_addr_u32(&s_in[0][0]), in1 = smem_addr_u32(&s_in[1][0]);
    if (threadIdx.x == 0) {
        mbar_init(mbar0, 1);
        mbar_init(mbar1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const uint32_t push_bytes = (uint32_t)(nc * k) * (uint32_t)sizeof(Cand);
    cluster.sync();   // every CTA's mbarriers are initialised before anybody pushes

    for (int64_t s = 0; s < n_order; ++s) {
        const int64_t row = row_cur;
        int32_t *prow = pred_idx + row * k;
        const int myp = myp_next;   // the row's current selection, one label per lane
        // s_blk was last read by the push of the previous instance (two block barriers ago)
        if (threadIdx.x < k) s_blk[threadIdx.x] = c_empty;
        // Arm this instance's input buffer: nc * k candidates are expected.  Pushes of faster peers may already have
        // landed -- the phase cannot complete before this arrival.  The buffer's previous use (instance s - 2) is over
        // everywhere: a peer pushes instance s only after it has received every CTA's push of s - 1, which a CTA sends
        // after its final ranking of s - 2.
        const uint32_t mbar = (s & 1) ? mbar1 : mbar0;
        if (threadIdx.x == 0) mbar_arrive_expect_tx(mbar, push_bytes);
#pragma unroll
        for (int l = 0; l < L; ++l) { pv[l] = pnext[l]; av[l] = anext[l]; }
        if (s + 1 < n_order) {
            const int64_t rnext = (int64_t)row_n1;
            row_cur = rnext;
            if (s + 2 < n_order) row_n1 = order ? order[s + 2] : (int)(s + 2);
            if (lane < k) myp_next = __ldg(pred_idx + rnext * k + lane);
#pragma unroll
            for (int l = 0; l < L; ++l) {
                int64_t j = j0 + l * stride;
                if (j < m) {
                    pnext[l] = __ldg(eta + rnext * ld + j);
                    if (addback) anext[l] = __ldg(addback + rnext * ld_add + j);
                }
            }
        }
        // ---- 1. remove own contribution, gains (block_coordinate.py:157-185) --------------------------
        double gain[L];
#pragma unroll
        for (int l = 0; l < L; ++l) {
            const int64_t j = j0 + l * stride;
            gain[l] = -INFINITY;
            bool sel = false;
            for (int t = 0; t < k; ++t) sel |= (__shfl_sync(XC_FULL, myp, t) == (int)j);
            if (j < m) {
                const TE pe = pv[l];
                const TE om = one - pe;
                const TE y = sel ? one : (TE)0;
                if (!greedy) {
                    stp[l] = stp[l] - (double)(TE)(y * pe);
                    sfp[l] = sfp[l] - (double)(TE)(y * om);
                    sfn[l] = sfn[l] - (double)(TE)((one - y) * pe);
                    if (use_tn) stn[l] = stn[l] - (double)(TE)((one - y) * om);
                }
                const double pos_tp = stp[l] + (double)pe;
                const double pos_fp = sfp[l] + (double)om;
                const double neg_fn = sfn[l] + (double)pe;
                const double neg_tn = use_tn ? stn[l] + (double)om : stn[l];
                // skip_tn: tn is never updated (the constant -1 vector of confusion_matrix.py:391-393), so its
                // quotient is loop invariant: 2 of the 10 float64 divisions of a label and instance
                const double pos_tn_n = use_tn ? stn[l] / nd : stn_n[l];
                const double neg_tn_n = use_tn ? neg_tn / nd : stn_n[l];
                const double up = xc_metric_eval(p, pos_tp / nd, pos_fp / nd, sfn[l] / nd, pos_tn_n);
                const double un = xc_metric_eval(p, stp[l] / nd, sfp[l] / nd, neg_fn / nd, neg_tn_n);
                const double g = up - un;
                gain[l] = p.maximize ? g : -g;
            }
        }
        // ---- 2a. warp top-k ---------------------------------------------------------------------------
        if (L == 1) {
            // one gain per lane: every lane ranks itself against the other 31 through smem
            // (32 broadcast reads instead of k dependent 5-level shuffle arg-max rounds)
            Cand me;
            me.g = j0 < m ? gain_key(gain[0]) : KEY_NEG_INF;
            me.j = j0 < m ? (int)j0 : 0x7fffffff;
            me.pad = 0;
            if (k <= 8) {
                // small budgets: k rounds of hardware warp reductions on the two halves of the key (REDUX.MAX on
                // the signed high word, then on the unsigned low word among the lanes that hold that high word);
                // the lowest such lane wins, and lanes are in label order, so ties go to the lowest label.  ~11
                // instructions per round against 32 x (LDS.128 + a 64-bit compare chain) for the ranking below.
                const int khi = (int)(me.g >> 32);
                const unsigned klo = (unsigned)(me.g & 0xffffffffLL);
                bool alive = true;
                for (int r = 0; r < k; ++r) {
                    const int mh = __reduce_max_sync(XC_FULL, alive ? khi : (int)0x80000000);
                    const bool c1 = alive && khi == mh;
                    const unsigned ml = __reduce_max_sync(XC_FULL, c1 ? klo : 0u);
                    const unsigned b = __ballot_sync(XC_FULL, c1 && klo == ml);
                    if (lane == __ffs(b) - 1) {
                        s_tmp[warp * k + r] = me;   // padding lanes carry the empty candidate
                        alive = false;
                    }
                }
            } else {
                Cand *wk = s_keys + warp * 32;
                wk[lane] = me;
                if (lane < k) s_tmp[warp * k + lane] = c_empty;   // fewer than k labels in this warp: unused slots stay empty
                __syncwarp();
                int rank = 0;
#pragma unroll 8
                for (int q = 0; q < 32; ++q) {
                    const Cand o = wk[q];
                    rank += xc_better(o.g, o.j, me.g, me.j) ? 1 : 0;
                }
                if (rank < k) s_tmp[warp * k + rank] = me;
            }
        } else {
            long long gkey[L];
#pragma unroll
            for (int l = 0; l < L; ++l) gkey[l] = gain_key(gain[l]);
            unsigned taken = 0;
            for (int r = 0; r < k; ++r) {
                Cand c;
                c.g = KEY_NEG_INF; c.j = 0x7fffffff; c.pad = -1;
#pragma unroll
                for (int l = 0; l < L; ++l) {
                    int64_t j = j0 + l * stride;
                    if (j < m && !((taken >> l) & 1u) && xc_better(gkey[l], (int)j, c.g, c.j)) {
                        c.g = gkey[l]; c.j = (int)j; c.pad = l;
                    }
                }
                Cand w = c;
                int wl = lane;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    long long og = __shfl_xor_sync(XC_FULL, w.g, o);
                    int oj = __shfl_xor_sync(XC_FULL, w.j, o);
                    int ol = __shfl_xor_sync(XC_FULL, wl, o);
                    if (xc_better(og, oj, w.g, w.j)) { w.g = og; w.j = oj; wl = ol; }
                }
                if (lane == wl && c.pad >= 0) taken |= 1u << c.pad;
                if (lane == 0) { Cand o; o.g = w.g; o.j = w.j; o.pad = 0; s_tmp[warp * k + r] = o; }
            }
        }
        __syncthreads();   // (1) warp candidates packed in s_tmp, s_blk emptied; every warp is past the previous re-add
        // ---- 2b. block top-k by ranking the NW * k warp candidates, then push to every CTA ----------------
        Cand *inbuf = s_in[s & 1];
        {
            // s_fin was last read by the re-add of the previous instance, which every warp left before barrier (1)
            if (threadIdx.x < k) s_fin[threadIdx.x] = c_empty;
            rank_topk_smem(s_tmp, NW * k, k, s_blk);
            __syncthreads();   // (2)
            // my block's k candidates go to slot `rank` of every CTA's input buffer (asynchronous DSMEM stores that
            // complete on the receiver's mbarrier; my own CTA is one of the receivers)
            const uint32_t in_base = (s & 1) ? in1 : in0;
            int dst_rank = push_dst0, r = push_r0;
            for (int q = threadIdx.x; q < nc * k;) {
                const Cand c = s_blk[r];
                const long long w1 = (long long)(((unsigned long long)(unsigned)c.pad << 32) | (unsigned)c.j);
                st_async_16(cluster_addr_of(in_base + (uint32_t)(rank * k + r) * (uint32_t)sizeof(Cand), (uint32_t)dst_rank),
                            c.g, w1, cluster_addr_of(mbar, (uint32_t)dst_rank));
                q += THREADS;
                if (q < nc * k) { dst_rank = q / k; r = q % k; }
            }
        }
        mbar_wait_parity(mbar, (uint32_t)((s >> 1) & 1));   // all nc * k candidates of this instance have landed
        // ---- 3. every CTA ranks the nc * k candidates (identical result everywhere) ---------------------
        rank_topk_smem(inbuf, nc * k, k, s_fin);
        __syncthreads();   // (3)
        // ---- 4. re-add with the new selection (block_coordinate.py:203-209) --------------------------------
        const int fin = lane < k ? s_fin[lane].j : -1;
#pragma unroll
        for (int l = 0; l < L; ++l) {
            const int64_t j = j0 + l * stride;
            bool sel = false;
            for (int t = 0; t < k; ++t) sel |= (__shfl_sync(XC_FULL, fin, t) == (int)j);
            if (j < m) {
                const TE pe = addback ? av[l] : pv[l];
                const TE om = one - pe;
                const TE y = sel ? one : (TE)0;
                stp[l] = stp[l] + (double)(TE)(y * pe);
                sfp[l] = sfp[l] + (double)(TE)(y * om);
                sfn[l] = sfn[l] + (double)(TE)((one - y) * pe);
                if (use_tn) stn[l] = stn[l] + (double)(TE)((one - y) * om);
            }
        }
        if (rank == 0 && warp == 0) {   // store the new row ascending by label
            int src = warp_rank_src(fin < 0 ? 0x7fffffff : fin, k);
            int v = __shfl_sync(XC_FULL, fin, src);
            if (lane < k) prow[lane] = v;
        }
    }
#pragma unroll
    for (int l = 0; l < L; ++l) {
        int64_t j = j0 + l * stride;
        if (j < m) {
            tp[j] = stp[l];
            fp[j] = sfp[l];
            fn[j] = sfn[l];
            if (use_tn) tn[j] = stn[l];
        }
    }
    cluster.sync();   // nobody exits while peers may still address its shared memory
}

// ---- k = 0 (no budget), dense: every label decides on its own (gain >= 0, block_coordinate.py:200),
// so the sweep decomposes per label: thread j walks the whole order for label j -- no exchange at all,
// still the reference's exact arithmetic.  pred is a dense [n, ldp] 0/1 matrix of the input dtype.
template <typename TE>
__global__ void __launch_bounds__(128)
bca_exact_dense_k0_kernel(const TE *__restrict__ eta, int64_t m, int64_t ld, const int32_t *__restrict__ order,
                          int64_t n_order, xc_metric_params p, int greedy, TE *pred, int64_t ldp, double *tp,
                          double *fp, double *fn, double *tn)
{
    const int64_t j = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (j >= m) return;
    const bool use_tn = !p.skip_tn;
    const double nd = p.n_div;
    const TE one = (TE)1;
    double stp = tp[j], sfp = fp[j], sfn = fn[j], stn = tn[j];
    for (int64_t s = 0; s < n_order; ++s) {
        const int64_t row = order[s];
        const TE pe = __ldg(eta + row * ld + j);
        const TE om = one - pe;
        TE y = pred[row * ldp + j];
        if (!greedy) {
            stp = stp - (double)(TE)(y * pe);
            sfp = sfp - (double)(TE)(y * om);
            sfn = sfn - (double)(TE)((one - y) * pe);
            if (use_tn) stn = stn - (double)(TE)((one - y) * om);
        }
        const double pos_tp = stp + (double)pe, pos_fp = sfp + (double)om, neg_fn = sfn + (double)pe;
        const double neg_tn = use_tn ? stn + (double)om : stn;
        const double up = xc_metric_eval(p, pos_tp / nd, pos_fp / nd, sfn / nd, stn / nd);
        const double un = xc_metric_eval(p, stp / nd, sfp / nd, neg_fn / nd, neg_tn / nd);
        double g = up - un;
        if (p.maximize) g = -g;           // :187-188
        y = (g <= 0.0) ? one : (TE)0;     // :191, :200
        pred[row * ldp + j] = y;
        stp = stp + (double)(TE)(y * pe);
        sfp = sfp + (double)(TE)(y * om);
        sfn = sfn + (double)(TE)((one - y) * pe);
        if (use_tn) stn = stn + (double)(TE)((one - y) * om);
    }
    tp[j] = stp;
    fp[j] = sfp;
    fn[j] = sfn;
    if (use_tn) tn[j] = stn;
}

// ---- CSR: one warp walks the order ------------------------------------------------------------------
// products exactly as numba forms them (numba_csr_functions.py:133, :206)
template <typename T> __device__ __forceinline__ T mul_round(T a, T b) { return (T)(a * b); }
template <typename T> __device__ __forceinline__ T mul_om_round(T a, T b) { return (T)((double)a * (1.0 - (double)b)); }

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

// +-(row contribution) for one CSR row given its compact prediction (lane x < k holds label pj).
// tn (optional): numba subtracts 1 from EVERY label and adds the three products back on the touched
// ones (numba_csr_functions.py:413-417 / :448-452).  For an untouched label (x - 1) + 1 == x exactly
// in float64 (both steps are exact), so only the touched labels are rewritten, with numba's
// operation order; each touched label is handled by exactly one lane.
template <typename T>
__device__ __forceinline__ void csr_apply_row(const T *data, const int32_t *indices, int64_t ts, int64_t te, int pj,
                                              int k, double sgn, double *tp, double *fp, double *fn, double *tn)
{
    const int lane = lane_id();
    const T one = (T)1;
    if (lane < k && pj >= 0) {  // numba_csr_functions.py:400-407 / 435-442
        int64_t y = csr_find(indices, ts, te, pj);
        if (y >= 0) {
            const double vtp = (double)mul_round(one, data[y]);
            const double vft = (double)mul_om_round(one, data[y]);
            tp[pj] = tp[pj] + sgn * vtp;
            fp[pj] = fp[pj] + sgn * vft;
            if (tn) {  // label in pred and in the row: fn_data = t * (1 - 1) = 0
                const double vfn = (double)mul_om_round(data[y], one);
                tn[pj] = (((tn[pj] + sgn) - sgn * vtp) - sgn * vft) - sgn * vfn;
            }
        } else {
            fp[pj] = fp[pj] + sgn * (double)one;
            if (tn) tn[pj] = (tn[pj] + sgn) - sgn * (double)one;
        }
    }
    for (int64_t q0 = ts; q0 < te; q0 += 32) {  // :408-411 / 443-446
        int64_t q = q0 + lane;
        int j = q < te ? indices[q] : -2;
        bool sel = false;
        for (int t = 0; t < k; ++t) sel |= (__shfl_sync(XC_FULL, pj, t) == j);
        if (q < te) {
            const double vfn = sel ? (double)mul_om_round(data[q], one) : (double)data[q];
            fn[j] = fn[j] + sgn * vfn;
            if (tn && !sel) tn[j] = (tn[j] + sgn) - sgn * vfn;
        }
    }
    __syncwarp();
}

template <typename T>
__global__ void __launch_bounds__(32)
bca_exact_csr_kernel(const T *__restrict__ data, const int32_t *__restrict__ indices,
                     const int64_t *__restrict__ indptr, const int32_t *__restrict__ order, int64_t n_order, int k,
                     xc_metric_params p, int greedy, int32_t *pred_idx, double *tp, double *fp, double *fn, double *tn)
{
    const int lane = lane_id();
    const double nd = p.n_div;
    const T one = (T)1;
    for (int64_t s = 0; s < n_order; ++s) {
        const int64_t row = order[s];
        const int64_t ts = indptr[row], te = indptr[row + 1], nz = te - ts;
        int32_t *prow = pred_idx + row * k;
        int pj = lane < k ? prow[lane] : -1;
        if (!greedy) csr_apply_row<T>(data, indices, ts, te, pj, k, -1.0, tp, fp, fn, tn);
        WarpTopK<double> tk;
        tk.init();
        for (int64_t q0 = ts; q0 < te; q0 += 32) {  // block_coordinate.py:248-282
            int64_t q = q0 + lane;
            double g[1];
            g[0] = NAN;
            if (q < te) {
                const int j = indices[q];
                const T t = data[q];
                const T om = one - t;
                double neg_tp = tp[j], neg_fp = fp[j], pos_fn = fn[j];
                const double pos_tpp = (neg_tp + (double)t) / nd;
                const double pos_fpp = (neg_fp + (double)om) / nd;
                const double neg_fnn = (pos_fn + (double)t) / nd;
                neg_tp = neg_tp / nd;
                neg_fp = neg_fp / nd;
                pos_fn = pos_fn / nd;
                double pos_tn = -1.0, neg_tnn = -1.0;       // block_coordinate.py:260-264
                if (tn) {
                    pos_tn = tn[j];
                    neg_tnn = (pos_tn + (double)om) / nd;
                    pos_tn = pos_tn / nd;
                }
                const double up = xc_metric_eval(p, pos_tpp, pos_fpp, pos_fn, pos_tn);
                const double un = xc_metric_eval(p, neg_tp, neg_fp, neg_fnn, neg_tnn);
                const double gg = up - un;
                g[0] = p.maximize ? gg : -gg;
            }
            if (__any_sync(XC_FULL, tk.passes(g[0]))) xc_scan_insert<double, 1, false>(tk, g, q0 - ts, 1, k, -1);
        }
        // nz > k: the k best (ascending label); else all stored labels (numba_csr_functions.py:456-466)
        (void)nz;
        int src = warp_rank_src(tk.idx, k);
        int pos = __shfl_sync(XC_FULL, tk.idx, src);
        pj = (lane < k && pos != 0x7fffffff) ? indices[ts + pos] : -1;
        if (lane < k) prow[lane] = pj;
        csr_apply_row<T>(data, indices, ts, te, pj, k, 1.0, tp, fp, fn, tn);
    }
}

// ---- CSR, online / greedy micro-batch -----------------------------------------------------------------
// ref: block_coordinate.py:212-293 called with greedy=True, only_pred=True (the row is predicted from the current
// confusion matrix, nothing is removed or re-added), followed by confusion_matrix.py:402-435 ->
// numba_add_to_unnormalized_confusion_matrix_csr with the row of y_true, as driven by
// experiments/omma_wrappers_online_methods.py:223-266.
//
// The reference's update adds 1 to tn of EVERY label per instance (numba_csr_functions.py:448) and corrects the
// touched ones -- O(m) per instance.  Here tn is kept lazily: last[j] counts the instances whose "+1" label j
// has already received; a label is brought up to date when a row touches it (and all labels at the end of the
// launch).  Bit-exact: a float64 "+1" is only rounded when the sum crosses into the next binade, so c pending
// additions are replayed as exact block additions up to the binade boundary plus one rounded addition across it.
__device__ __forceinline__ double tn_fast_forward(double x, long long c)
{
    while (c > 0) {
        if (!(x >= 1.0) || !(x < 4503599627370496.0)) {   // below 1 (or absurdly large): plain rounded additions
            x = x + 1.0;
            --c;
            continue;
        }
        const double limit = ldexp(1.0, ilogb(x) + 1);        // x in [limit / 2, limit)
        long long s = (long long)ceil(limit - x) - 1;         // additions that stay below the boundary: exact
        if (s > c) s = c;
        if (s > 0) {
            x = x + (double)s;
            c -= s;
        } else {
            x = x + 1.0;                                      // the addition that crosses the boundary (rounded)
            --c;
        }
    }
    return x;
}

__global__ void __launch_bounds__(256) tn_flush_kernel(double *tn, int32_t *last, int64_t m, int32_t now)
{
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j >= m) return;
    const int c = now - last[j];
    if (c > 0) tn[j] = tn_fast_forward(tn[j], c);
    last[j] = now;
}

// bring tn of the labels of one CSR row up to date (lane-parallel, labels of a row are distinct)
__device__ __forceinline__ void tn_catch_up_row(const int32_t *indices, int64_t s, int64_t e, double *tn, int32_t *last,
                                                int32_t now)
{
    for (int64_t q = s + lane_id(); q < e; q += 32) {
        const int j = indices[q];
        const int c = now - last[j];
        if (c > 0) {
            tn[j] = tn_fast_forward(tn[j], c);
            last[j] = now;
        }
    }
    __syncwarp();
}

template <typename T>
__global__ void __launch_bounds__(32)
bca_online_csr_kernel(const T *__restrict__ data, const int32_t *__restrict__ indices,
                      const int64_t *__restrict__ indptr, const T *__restrict__ t_data,
                      const int32_t *__restrict__ t_indices, const int64_t *__restrict__ t_indptr, int64_t n_rows, int k,
                      xc_metric_params p, int32_t *pred_idx, double *tp, double *fp, double *fn, double *tn,
                      int32_t *tn_last, int32_t step0)
{
    const int lane = lane_id();
    const double nd = p.n_div;
    const T one = (T)1;
    const bool use_tn = !p.skip_tn;
    for (int64_t i = 0; i < n_rows; ++i) {
        const int32_t now = step0 + (int32_t)i;   // instances whose "+1" every label should have received so far
        const int64_t ts = indptr[i], te = indptr[i + 1];
        if (use_tn) tn_catch_up_row(indices, ts, te, tn, tn_last, now);
        WarpTopK<double> tk;
        tk.init();
        for (int64_t q0 = ts; q0 < te; q0 += 32) {  // block_coordinate.py:243-282 (greedy: no removal)
            const int64_t q = q0 + lane;
            double g[1];
            g[0] = NAN;
            if (q < te) {
                const int j = indices[q];
                const T t = data[q];
                const T om = one - t;
                double neg_tp = tp[j], neg_fp = fp[j], pos_fn = fn[j];
                const double pos_tpp = (neg_tp + (double)t) / nd;
                const double pos_fpp = (neg_fp + (double)om) / nd;
                const double neg_fnn = (pos_fn + (double)t) / nd;
                neg_tp = neg_tp / nd;
                neg_fp = neg_fp / nd;
                pos_fn = pos_fn / nd;
                double pos_tn = tn[j], neg_tnn = pos_tn;     // :258-264: raw tn on both sides when skip_tn
                if (use_tn) {
                    neg_tnn = (pos_tn + (double)om) / nd;
                    pos_tn = pos_tn / nd;
                }
                const double up = xc_metric_eval(p, pos_tpp, pos_fpp, pos_fn, pos_tn);
                const double un = xc_metric_eval(p, neg_tp, neg_fp, neg_fnn, neg_tnn);
                const double gg = up - un;
                g[0] = p.maximize ? gg : -gg;
            }
            if (__any_sync(XC_FULL, tk.passes(g[0]))) xc_scan_insert<double, 1, false>(tk, g, q0 - ts, 1, k, -1);
        }
        const int src = warp_rank_src(tk.idx, k);
        const int pos = __shfl_sync(XC_FULL, tk.idx, src);
        const int pj = (lane < k && pos != 0x7fffffff) ? indices[ts + pos] : -1;
        if (lane < k) pred_idx[i * k + lane] = pj;
        // confusion_matrix.py:402-435 with the row of y_true (prediction entries are ones)
        const int64_t us = t_indptr[i], ue = t_indptr[i + 1];
        if (use_tn) tn_catch_up_row(t_indices, us, ue, tn, tn_last, now);   // (predicted labels: caught up above)
        csr_apply_row<T>(t_data, t_indices, us, ue, pj, k, 1.0, tp, fp, fn, use_tn ? tn : nullptr);
        if (use_tn) {   // the touched labels have now received this instance's "+1" as well
            if (lane < k && pj >= 0) tn_last[pj] = now + 1;
            for (int64_t q = us + lane; q < ue; q += 32) tn_last[t_indices[q]] = now + 1;
            __syncwarp();
        }
    }
}

// ---- CSR, register-staged fast path ------------------------------------------------------------------
// The generic kernel above walks every step through ~30 dependent L2 round trips (order -> indptr ->
// row -> binary searches -> state ...): 26 us per instance.  Here a parallel prologue resolves all
// order-dependent indirections for the whole sweep (row start, row length, the row's current
// prediction -- a row is visited once per sweep, so its prediction cannot change before its turn), the
// walking warp prefetches the next row's entries into registers one step ahead, matches prediction and
// row by shuffles instead of searches, keeps the touched state in registers from "remove" to "re-add"
// and stores it once.  What remains per step is one L2 gather of the state, the float64 gains and the
// warp top-k.  Rows with more than 32 * CSR_RMAX stored labels take the generic per-row routine.
constexpr int CSR_RMAX = 4;

__global__ void __launch_bounds__(256)
csr_step_meta_kernel(const int32_t *__restrict__ order, const int64_t *__restrict__ indptr,
                     const int32_t *__restrict__ pred_idx, int k, int64_t n_order, int64_t *__restrict__ step_ts,
                     int32_t *__restrict__ step_nz, int32_t *__restrict__ step_pred)
{
    const int64_t s = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (s >= n_order) return;
    const int64_t row = order[s];
    const int64_t ts = indptr[row];
    step_ts[s] = ts;
    step_nz[s] = (int32_t)(indptr[row + 1] - ts);
    for (int t = 0; t < k; ++t) step_pred[s * k + t] = pred_idx[row * k + t];
}

template <typename T>
__global__ void __launch_bounds__(32)
bca_exact_csr_fast_kernel(const T *__restrict__ data, const int32_t *__restrict__ indices,
                          const int32_t *__restrict__ order, const int64_t *__restrict__ step_ts,
                          const int32_t *__restrict__ step_nz, const int32_t *__restrict__ step_pred, int64_t n_order,
                          int k, xc_metric_params p, int greedy, int32_t *pred_idx, double *tp, double *fp, double *fn,
                          double *tn)
{
    const int lane = lane_id();
    const double nd = p.n_div;
    const T one = (T)1;
    // software pipeline: entries of step s+1 and metadata of step s+2 are in flight during step s
    int64_t ts_c = 0, ts_n = 0;
    int nz_c = 0, nz_n = 0, pj_c = -1, pj_n = -1;
    int idx_c[CSR_RMAX], idx_n[CSR_RMAX];
    T val_c[CSR_RMAX], val_n[CSR_RMAX];
#pragma unroll
    for (int t = 0; t < CSR_RMAX; ++t) { idx_c[t] = idx_n[t] = -2; val_c[t] = val_n[t] = (T)0; }
    if (n_order > 0) {
        ts_c = step_ts[0];
        nz_c = step_nz[0];
        pj_c = lane < k ? step_pred[lane] : -1;
#pragma unroll
        for (int t = 0; t < CSR_RMAX; ++t) {
            int q = lane + 32 * t;
            if (q < nz_c && nz_c <= 32 * CSR_RMAX) { idx_c[t] = indices[ts_c + q]; val_c[t] = data[ts_c + q]; }
        }
    }
    if (n_order > 1) {
        ts_n = step_ts[1];
        nz_n = step_nz[1];
        pj_n = lane < k ? step_pred[k + lane] : -1;
    }
    for (int64_t s = 0; s < n_order; ++s) {
        // ---- prefetch: entries of step s+1 (its row start is already known), metadata of step s+2
#pragma unroll
        for (int t = 0; t < CSR_RMAX; ++t) {
            int q = lane + 32 * t;
            idx_n[t] = -2;
            val_n[t] = (T)0;
            if (s + 1 < n_order && q < nz_n && nz_n <= 32 * CSR_RMAX) {
                idx_n[t] = indices[ts_n + q];
                val_n[t] = data[ts_n + q];
            }
        }
        int64_t ts_nn = 0;
        int nz_nn = 0, pj_nn = -1;
        if (s + 2 < n_order) {
            ts_nn = step_ts[s + 2];
            nz_nn = step_nz[s + 2];
            pj_nn = lane < k ? step_pred[(s + 2) * k + lane] : -1;
        }
        const int64_t row = order[s];
        int32_t *prow = pred_idx + row * k;

        if (nz_c > 32 * CSR_RMAX) {
            // ---- long row: generic per-row routine (global-memory searches) ------------------------------
            const int64_t ts = ts_c, te = ts_c + nz_c;
            int pj = pj_c;
            if (!greedy) csr_apply_row<T>(data, indices, ts, te, pj, k, -1.0, tp, fp, fn, tn);
            WarpTopK<double> tk;
            tk.init();
            for (int64_t q0 = ts; q0 < te; q0 += 32) {
                int64_t q = q0 + lane;
                double g[1];
                g[0] = NAN;
                if (q < te) {
                    const int j = indices[q];
                    const T tv = data[q];
                    const T om = one - tv;
                    double neg_tp = tp[j], neg_fp = fp[j], pos_fn = fn[j];
                    const double pos_tpp = (neg_tp + (double)tv) / nd;
                    const double pos_fpp = (neg_fp + (double)om) / nd;
                    const double neg_fnn = (pos_fn + (double)tv) / nd;
                    neg_tp = neg_tp / nd; neg_fp = neg_fp / nd; pos_fn = pos_fn / nd;
                    double pos_tn = -1.0, neg_tnn = -1.0;
                    if (tn) { pos_tn = tn[j]; neg_tnn = (pos_tn + (double)om) / nd; pos_tn = pos_tn / nd; }
                    const double up = xc_metric_eval(p, pos_tpp, pos_fpp, pos_fn, pos_tn);
                    const double un = xc_metric_eval(p, neg_tp, neg_fp, neg_fnn, neg_tnn);
                    const double gg = up - un;
                    g[0] = p.maximize ? gg : -gg;
                }
                if (__any_sync(XC_FULL, tk.passes(g[0]))) xc_scan_insert<double, 1, false>(tk, g, q0 - ts, 1, k, -1);
            }
            int src = warp_rank_src(tk.idx, k);
            int pos = __shfl_sync(XC_FULL, tk.idx, src);
            pj = (lane < k && pos != 0x7fffffff) ? indices[ts + pos] : -1;
            if (lane < k) prow[lane] = pj;
            csr_apply_row<T>(data, indices, ts, te, pj, k, 1.0, tp, fp, fn, tn);
        } else {
            // ---- register-staged row ---------------------------------------------------------------------
            // which stored entries are currently predicted / which predicted labels are not stored
            bool sel[CSR_RMAX];
            bool pj_found = false;
#pragma unroll
            for (int t = 0; t < CSR_RMAX; ++t) sel[t] = false;
            for (int x = 0; x < k; ++x) {
                const int px = __shfl_sync(XC_FULL, pj_c, x);
                bool any = false;
#pragma unroll
                for (int t = 0; t < CSR_RMAX; ++t) {
                    const bool hit = px >= 0 && idx_c[t] == px;
                    sel[t] |= hit;
                    any |= hit;
                }
                const bool f = __any_sync(XC_FULL, any);
                if (lane == x) pj_found = f;
            }
            // state of the touched labels -> registers
            double stp[CSR_RMAX], sfp[CSR_RMAX], sfn[CSR_RMAX], stn[CSR_RMAX];
#pragma unroll
            for (int t = 0; t < CSR_RMAX; ++t) {
                stp[t] = sfp[t] = sfn[t] = 0.0;
                stn[t] = -1.0;
                if (idx_c[t] >= 0) {
                    stp[t] = tp[idx_c[t]];
                    sfp[t] = fp[idx_c[t]];
                    sfn[t] = fn[idx_c[t]];
                    if (tn) stn[t] = tn[idx_c[t]];
                }
            }
            // predicted label that the row does not store (lane x < k): fp / tn only
            const bool lone = lane < k && pj_c >= 0 && !pj_found;
            double lfp = 0.0, ltn = 0.0;
            if (lone) {
                lfp = fp[pj_c];
                if (tn) ltn = tn[pj_c];
            }
            // ---- remove (numba_csr_functions.py:386-417)
            if (!greedy) {
#pragma unroll
                for (int t = 0; t < CSR_RMAX; ++t) {
                    if (idx_c[t] < 0) continue;
                    if (sel[t]) {
                        const double vtp = (double)mul_round(one, val_c[t]);
                        const double vft = (double)mul_om_round(one, val_c[t]);
                        const double vfn = (double)mul_om_round(val_c[t], one);
                        stp[t] = stp[t] - vtp;
                        sfp[t] = sfp[t] - vft;
                        sfn[t] = sfn[t] - vfn;
                        if (tn) stn[t] = (((stn[t] - 1.0) + vtp) + vft) + vfn;
                    } else {
                        const double vfn = (double)val_c[t];
                        sfn[t] = sfn[t] - vfn;
                        if (tn) stn[t] = (stn[t] - 1.0) + vfn;
                    }
                }
                if (lone) {
                    lfp = lfp - (double)one;
                    if (tn) ltn = (ltn - 1.0) + (double)one;
                }
            }
            // ---- gains (block_coordinate.py:248-282) and top-k of the stored entries
            WarpTopK<double> tk;
            tk.init();
#pragma unroll
            for (int t = 0; t < CSR_RMAX; ++t) {
                double g[1];
                g[0] = NAN;
                if (idx_c[t] >= 0) {
                    const T tv = val_c[t];
                    const T om = one - tv;
                    double neg_tp = stp[t], neg_fp = sfp[t], pos_fn = sfn[t];
                    const double pos_tpp = (neg_tp + (double)tv) / nd;
                    const double pos_fpp = (neg_fp + (double)om) / nd;
                    const double neg_fnn = (pos_fn + (double)tv) / nd;
                    neg_tp = neg_tp / nd; neg_fp = neg_fp / nd; pos_fn = pos_fn / nd;
                    double pos_tn = -1.0, neg_tnn = -1.0;
                    if (tn) { pos_tn = stn[t]; neg_tnn = (pos_tn + (double)om) / nd; pos_tn = pos_tn / nd; }
                    const double up = xc_metric_eval(p, pos_tpp, pos_fpp, pos_fn, pos_tn);
                    const double un = xc_metric_eval(p, neg_tp, neg_fp, neg_fnn, neg_tnn);
                    const double gg = up - un;
                    g[0] = p.maximize ? gg : -gg;
                }
                if (32 * t < nz_c && __any_sync(XC_FULL, tk.passes(g[0])))
                    xc_scan_insert<double, 1, false>(tk, g, 32 * t, 1, k, -1);
            }
            // new selection: positions inside the row, ascending (== ascending label)
            const int src = warp_rank_src(tk.idx, k);
            const int pos = __shfl_sync(XC_FULL, tk.idx, src);
            // label of position `pos` lives in lane pos % 32, slot pos / 32
            int newp = -1;
            {
                const int want_lane = pos == 0x7fffffff ? 0 : (pos & 31), want_t = pos == 0x7fffffff ? 0 : (pos >> 5);
                int got = -1;
#pragma unroll
                for (int t = 0; t < CSR_RMAX; ++t) {
                    const int v = __shfl_sync(XC_FULL, idx_c[t], want_lane);
                    if (t == want_t) got = v;
                }
                if (lane < k && pos != 0x7fffffff) newp = got;
            }
            if (lane < k) prow[lane] = newp;
            // ---- re-add with the new selection (numba_csr_functions.py:421-452), store once
            bool nsel[CSR_RMAX];
#pragma unroll
            for (int t = 0; t < CSR_RMAX; ++t) nsel[t] = false;
            for (int x = 0; x < k; ++x) {
                const int px = __shfl_sync(XC_FULL, newp, x);
#pragma unroll
                for (int t = 0; t < CSR_RMAX; ++t) nsel[t] |= (px >= 0 && idx_c[t] == px);
            }
#pragma unroll
            for (int t = 0; t < CSR_RMAX; ++t) {
                if (idx_c[t] < 0) continue;
                if (nsel[t]) {
                    const double vtp = (double)mul_round(one, val_c[t]);
                    const double vft = (double)mul_om_round(one, val_c[t]);
                    const double vfn = (double)mul_om_round(val_c[t], one);
                    stp[t] = stp[t] + vtp;
                    sfp[t] = sfp[t] + vft;
                    sfn[t] = sfn[t] + vfn;
                    if (tn) stn[t] = (((stn[t] + 1.0) - vtp) - vft) - vfn;
                } else {
                    const double vfn = (double)val_c[t];
                    sfn[t] = sfn[t] + vfn;
                    if (tn) stn[t] = (stn[t] + 1.0) - vfn;
                }
                tp[idx_c[t]] = stp[t];
                fp[idx_c[t]] = sfp[t];
                fn[idx_c[t]] = sfn[t];
                if (tn) tn[idx_c[t]] = stn[t];
            }
            if (lone && !greedy) {   // the new selection is a subset of the row: this label only left
                fp[pj_c] = lfp;
                if (tn) tn[pj_c] = ltn;
            }
            __syncwarp();
        }
        // ---- rotate the pipeline
        ts_c = ts_n; nz_c = nz_n; pj_c = pj_n;
#pragma unroll
        for (int t = 0; t < CSR_RMAX; ++t) { idx_c[t] = idx_n[t]; val_c[t] = val_n[t]; }
        ts_n = ts_nn; nz_n = nz_nn; pj_n = pj_nn;
    }
}

// ---- coverage (block_coordinate.py:539-580) ---------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(32)
cov_exact_csr_kernel(const T *__restrict__ data, const int32_t *__restrict__ indices,
                     const int64_t *__restrict__ indptr, const int32_t *__restrict__ order, int64_t n_order, int k,
                     double alpha, int greedy, int32_t *pred_idx, double *Ef)
{
    const int lane = lane_id();
    const T one = (T)1;
    for (int64_t s = 0; s < n_order; ++s) {
        const int64_t row = order[s];
        const int64_t ts = indptr[row], te = indptr[row + 1];
        int32_t *prow = pred_idx + row * k;
        int pj = lane < k ? prow[lane] : -1;
        if (!greedy && lane < k && pj >= 0) {  // :562-564
            int64_t y = csr_find(indices, ts, te, pj);
            if (y >= 0) Ef[pj] = Ef[pj] / (double)(T)(one - mul_round(one, data[y]));
        }
        __syncwarp();
        WarpTopK<double> tk;
        tk.init();
        for (int64_t q0 = ts; q0 < te; q0 += 32) {  // :567-569
            int64_t q = q0 + lane;
            double g[1];
            g[0] = NAN;
            if (q < te) {
                const T t = data[q];
                double gg = Ef[indices[q]] * (double)t;
                if (alpha < 1.0) {
                    T w = (T)((T)(1.0 - alpha) * t);
                    w = (T)(w / (T)k);
                    gg = alpha * gg + (double)w;
                }
                g[0] = gg;
            }
            if (__any_sync(XC_FULL, tk.passes(g[0]))) xc_scan_insert<double, 1, false>(tk, g, q0 - ts, 1, k, -1);
        }
        int src = warp_rank_src(tk.idx, k);
        int pos = __shfl_sync(XC_FULL, tk.idx, src);
        pj = (lane < k && pos != 0x7fffffff) ? indices[ts + pos] : -1;
        if (lane < k) prow[lane] = pj;
        if (lane < k && pj >= 0) {  // :579-580
            T t = data[ts + pos];
            Ef[pj] = Ef[pj] * (double)(T)(one - mul_round(one, t));
        }
        __syncwarp();
    }
}

// Ef = prod_i (1 - yhat_ij eta_ij), rows in order (numba_csr_functions.py:325-382)
template <typename T>
__global__ void __launch_bounds__(32)
cov_state_ordered_kernel(const T *__restrict__ data, const int32_t *__restrict__ indices,
                         const int64_t *__restrict__ indptr, int64_t n, const int32_t *__restrict__ pred_idx, int k,
                         double *Ef)
{
    const int lane = lane_id();
    const T one = (T)1;
    for (int64_t i = 0; i < n; ++i) {
        int pj = lane < k ? pred_idx[i * k + lane] : -1;
        if (pj >= 0) {
            int64_t y = csr_find(indices, indptr[i], indptr[i + 1], pj);
            Ef[pj] = Ef[pj] * (y >= 0 ? (double)mul_om_round(one, data[y]) : (double)one);
        }
        __syncwarp();
    }
}

__device__ __forceinline__ void atomic_mul_d(double *addr, double f)
{
    unsigned long long *a = reinterpret_cast<unsigned long long *>(addr);
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        old = atomicCAS(a, assumed, __double_as_longlong(__longlong_as_double(assumed) * f));
    } while (assumed != old);
}

template <typename T>
__global__ void __launch_bounds__(256)
cov_state_fast_kernel(const T *__restrict__ data, const int32_t *__restrict__ indices,
                      const int64_t *__restrict__ indptr, int64_t n, const int32_t *__restrict__ pred_idx, int k,
                      double *Ef)
{
    const T one = (T)1;
    const int64_t total = n * k;
    for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
        int pj = pred_idx[t];
        if (pj < 0) continue;
        int64_t i = t / k;
        int64_t y = csr_find(indices, indptr[i], indptr[i + 1], pj);
        if (y >= 0) atomic_mul_d(Ef + pj, (double)mul_om_round(one, data[y]));
    }
}

__global__ void fill_kernel(double *x, double v, int64_t m)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) x[j] = v;
}

template <typename TE, int L>
int launch_exact_dense(xc_ctx *ctx, int grid, const void *eta, int64_t m, int64_t ld, const int32_t *order,
                       int64_t n_order, int k, const xc_metric_params *p, int greedy, int32_t *pred_idx, double *tp,
                       double *fp, double *fn, double *tn, Cand *cand, cudaStream_t st)
{
    const TE *eta_t = (const TE *)eta;
    xc_metric_params pp = *p;
    void *args[] = {&eta_t, &m, &ld, &order, &n_order, &k, &pp, &greedy, &pred_idx, &tp, &fp, &fn, &tn, &cand};
    XC_CUDA_TRY(ctx, cudaLaunchCooperativeKernel((void *)bca_exact_dense_kernel<TE, L>, dim3(grid), dim3(EX_THREADS),
                                                 args, 0, st));
    XC_LAUNCHED(ctx);
    return XC_OK;
}

template <typename TE, int L, int THREADS>
int launch_exact_cluster(xc_ctx *ctx, int nc, const void *eta, int64_t m, int64_t ld, const int32_t *order,
                         int64_t n_order, int k, const xc_metric_params *p, int greedy, int32_t *pred_idx, double *tp,
                         double *fp, double *fn, double *tn, cudaStream_t st, const void *addback = nullptr,
                         int64_t ld_add = 0)
{
    auto kern = bca_exact_dense_cluster_kernel<TE, L, THREADS>;
    if (nc > 8) XC_CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nc);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = nc;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess || max_clusters < 1) {
        cudaGetLastError();
        return XC_ERR_UNSUPPORTED;  // caller falls back to the grid-barrier kernel
    }
    XC_CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, kern, (const TE *)eta, m, ld, order, n_order, k, *p, greedy, pred_idx, tp,
                                        fp, fn, tn, (const TE *)addback, ld_add));
    XC_LAUNCHED(ctx);
    return XC_OK;
}

// cluster shape for m labels: THREADS * L labels per CTA, at most 16 CTAs; prefers a portable (<= 8) cluster
template <typename TE>
int dispatch_exact_cluster(xc_ctx *ctx, const void *eta, int64_t m, int64_t ld, const int32_t *order, int64_t n_order,
                           int k, const xc_metric_params *p, int greedy, int32_t *pred_idx, double *tp, double *fp,
                           double *fn, double *tn, cudaStream_t st, const void *addback = nullptr, int64_t ld_add = 0)
{
    auto ncta = [&](int64_t per_cta) { return (int)((m + per_cta - 1) / per_cta); };
#define XC_TRY(LL, TH)                                                                                               \
    {                                                                                                                \
        int nc = ncta((int64_t)(TH) * (LL));                                                                         \
        if (nc <= CL_MAX)                                                                                            \
            return launch_exact_cluster<TE, LL, TH>(ctx, nc, eta, m, ld, order, n_order, k, p, greedy, pred_idx, tp, \
                                                    fp, fn, tn, st, addback, ld_add);                                \
    }
    // Few labels per CTA and many CTAs: the per-instance cost is the float64 divisions of the gains (10 per label;
    // vector FP64 is ~16 lanes / clk / SM on this part), which a single SM with 512 labels needs ~5 us for --
    // the cluster barrier and the DSMEM exchange cost ~1.5 us whatever the cluster size.  $XCOLUMNS_B200_EXACT_CTA=512
    // restores the wide CTAs.
    static int wide = -1;
    if (wide < 0) {
        const char *e = getenv("XCOLUMNS_B200_EXACT_CTA");
        wide = (e && atoi(e) == 512) ? 1 : 0;
    }
    if (!wide) {
        if (ncta(128) <= CL_MAX) XC_TRY(1, 128)
        if (ncta(256) <= CL_MAX) XC_TRY(1, 256)
    }
    if (ncta(512) <= 8) XC_TRY(1, 512)
    if (ncta(1024) <= 8) XC_TRY(2, 512)
    XC_TRY(1, 512)
    XC_TRY(2, 512)
    XC_TRY(4, 512)
    XC_TRY(8, 256)
#undef XC_TRY
    return XC_ERR_UNSUPPORTED;
}

template <typename TE>
int dispatch_exact_dense(xc_ctx *ctx, const void *eta, int64_t m, int64_t ld, const int32_t *order, int64_t n_order,
                         int k, const xc_metric_params *p, int greedy, int32_t *pred_idx, double *tp, double *fp,
                         double *fn, double *tn, cudaStream_t st)
{
    // co-resident capacity of the largest-register variant bounds the grid for all variants
    int per_sm = 0;
    XC_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bca_exact_dense_kernel<TE, 8>, EX_THREADS, 0));
    if (per_sm < 1) return XC_ERR_UNSUPPORTED;
    if (per_sm > 4) per_sm = 4;  // barrier cost grows with the grid; 4 x 148 x 128 x 8 labels = 606k labels
    int64_t cap = (int64_t)ctx->sm_count * per_sm;
    int64_t need = (m + EX_THREADS - 1) / EX_THREADS;
    // prefer one wave of <= sm_count blocks when L stays small
    int64_t grid = need < (int64_t)ctx->sm_count ? need : (int64_t)ctx->sm_count;
    int64_t L = (m + grid * EX_THREADS - 1) / (grid * EX_THREADS);
    if (L > 8) {
        grid = need < cap ? need : cap;
        L = (m + grid * EX_THREADS - 1) / (grid * EX_THREADS);
    }
    if (L > 8) return XC_ERR_UNSUPPORTED;
    void *scratch = nullptr;
    int rc = xc_ctx_scratch(ctx, sizeof(Cand) * 2 * (size_t)grid * (size_t)k, &scratch);
    if (rc) return rc;
    Cand *cand = (Cand *)scratch;
#define XC_GO(LL) return launch_exact_dense<TE, LL>(ctx, (int)grid, eta, m, ld, order, n_order, k, p, greedy, pred_idx, tp, fp, fn, tn, cand, st)
    if (L <= 1) XC_GO(1);
    if (L <= 2) XC_GO(2);
    if (L <= 4) XC_GO(4);
    XC_GO(8);
#undef XC_GO
}

}  // namespace

extern "C" int xc_bca_exact_sweep_dense(xc_ctx *ctx, const void *eta, int dtype, int64_t n, int64_t m, int64_t ld,
                                        const int32_t *order, int64_t n_order, int k, const xc_metric_params *p,
                                        int greedy, int32_t *pred_idx, double *tp, double *fp, double *fn, double *tn,
                                        void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !eta || !order || !p || !pred_idx || !tp || !fp || !fn || !tn) return XC_ERR_INVALID;
    if (n <= 0 || m <= 0 || ld < m || n_order < 0 || k < 1 || k > 32 || k > m) return XC_ERR_INVALID;
    if (p->metric < 0 || p->metric > XC_METRIC_PREC_AT_K) return XC_ERR_INVALID;
    if (n_order == 0) return XC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    static int grid_only = -1;   // $XCOLUMNS_B200_EXACT_PATH=grid forces the cooperative-grid kernel
    if (grid_only < 0) {
        const char *e = getenv("XCOLUMNS_B200_EXACT_PATH");
        grid_only = (e && e[0] == 'g') ? 1 : 0;
    }
    if (!grid_only) {
        int rc = XC_ERR_UNSUPPORTED;
        if (dtype == XC_F32) rc = dispatch_exact_cluster<float>(ctx, eta, m, ld, order, n_order, k, p, greedy, pred_idx, tp, fp, fn, tn, st);
        else if (dtype == XC_F64) rc = dispatch_exact_cluster<double>(ctx, eta, m, ld, order, n_order, k, p, greedy, pred_idx, tp, fp, fn, tn, st);
        if (rc != XC_ERR_UNSUPPORTED) return rc;
    }
    if (dtype == XC_F32)
        return dispatch_exact_dense<float>(ctx, eta, m, ld, order, n_order, k, p, greedy, pred_idx, tp, fp, fn, tn, st);
    if (dtype == XC_F64)
        return dispatch_exact_dense<double>(ctx, eta, m, ld, order, n_order, k, p, greedy, pred_idx, tp, fp, fn, tn, st);
    return XC_ERR_UNSUPPORTED;
}

extern "C" int xc_bca_online_dense(xc_ctx *ctx, const void *eta, int dtype, int64_t n_rows, int64_t m, int64_t ld,
                                   const void *y_true, int64_t ld_true, int k, const xc_metric_params *p,
                                   int32_t *pred_idx, double *tp, double *fp, double *fn, double *tn, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !eta || !p || !pred_idx || !tp || !fp || !fn || !tn) return XC_ERR_INVALID;
    if (n_rows < 0 || m <= 0 || ld < m || k < 1 || k > 32 || k > m || (y_true && ld_true < m)) return XC_ERR_INVALID;
    if (p->metric < 0 || p->metric > XC_METRIC_PREC_AT_K) return XC_ERR_INVALID;
    if (n_rows == 0) return XC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == XC_F32)
        return dispatch_exact_cluster<float>(ctx, eta, m, ld, nullptr, n_rows, k, p, 1, pred_idx, tp, fp, fn, tn, st, y_true, ld_true);
    if (dtype == XC_F64)
        return dispatch_exact_cluster<double>(ctx, eta, m, ld, nullptr, n_rows, k, p, 1, pred_idx, tp, fp, fn, tn, st, y_true, ld_true);
    return XC_ERR_UNSUPPORTED;
}

/* tn_last [m] int32 (zero before the first call): see bca_online_csr_kernel; step0: instances processed by earlier
 * calls.  On return tn is up to date for every label (tn_last = step0 + n_rows). */
extern "C" int xc_bca_online_csr(xc_ctx *ctx, const void *data, int dtype, const int32_t *indices, const int64_t *indptr,
                                 const void *t_data, const int32_t *t_indices, const int64_t *t_indptr, int64_t n_rows,
                                 int64_t m, int k, const xc_metric_params *p, int32_t *pred_idx, double *tp, double *fp,
                                 double *fn, double *tn, int32_t *tn_last, int64_t step0, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !indptr || !t_indptr || !p || !pred_idx || !tp || !fp || !fn || !tn || !tn_last) return XC_ERR_INVALID;
    if (n_rows < 0 || m <= 0 || k < 1 || k > 32 || step0 < 0 || step0 + n_rows > 0x7fffffffLL) return XC_ERR_INVALID;
    if (p->metric < 0 || p->metric > XC_METRIC_PREC_AT_K) return XC_ERR_INVALID;
    if (n_rows == 0) return XC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == XC_F32)
        bca_online_csr_kernel<float><<<1, 32, 0, st>>>((const float *)data, indices, indptr, (const float *)t_data, t_indices,
                                                       t_indptr, n_rows, k, *p, pred_idx, tp, fp, fn, tn, tn_last,
                                                       (int32_t)step0);
    else if (dtype == XC_F64)
        bca_online_csr_kernel<double><<<1, 32, 0, st>>>((const double *)data, indices, indptr, (const double *)t_data,
                                                        t_indices, t_indptr, n_rows, k, *p, pred_idx, tp, fp, fn, tn,
                                                        tn_last, (int32_t)step0);
    else
        return XC_ERR_UNSUPPORTED;
    XC_LAUNCHED(ctx);
    if (!p->skip_tn) {
        tn_flush_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(tn, tn_last, m, (int32_t)(step0 + n_rows));
        XC_LAUNCHED(ctx);
    }
    return XC_OK;
}

extern "C" int xc_bca_exact_sweep_dense_k0(xc_ctx *ctx, const void *eta, int dtype, int64_t n, int64_t m, int64_t ld,
                                           const int32_t *order, int64_t n_order, const xc_metric_params *p,
                                           int greedy, void *pred, int64_t ld_pred, double *tp, double *fp, double *fn,
                                           double *tn, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !eta || !order || !p || !pred || !tp || !fp || !fn || !tn) return XC_ERR_INVALID;
    if (n <= 0 || m <= 0 || ld < m || ld_pred < m || n_order < 0) return XC_ERR_INVALID;
    if (p->metric < 0 || p->metric > XC_METRIC_PREC_AT_K) return XC_ERR_INVALID;
    if (n_order == 0) return XC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned grid = (unsigned)((m + 127) / 128);
    if (dtype == XC_F32)
        bca_exact_dense_k0_kernel<float><<<grid, 128, 0, st>>>((const float *)eta, m, ld, order, n_order, *p, greedy, (float *)pred, ld_pred, tp, fp, fn, tn);
    else if (dtype == XC_F64)
        bca_exact_dense_k0_kernel<double><<<grid, 128, 0, st>>>((const double *)eta, m, ld, order, n_order, *p, greedy, (double *)pred, ld_pred, tp, fp, fn, tn);
    else
        return XC_ERR_UNSUPPORTED;
    XC_LAUNCHED(ctx);
    return XC_OK;
}

extern "C" int xc_bca_exact_sweep_csr(xc_ctx *ctx, const void *data, int dtype, const int32_t *indices,
                                      const int64_t *indptr, int64_t n, int64_t m, const int32_t *order,
                                      int64_t n_order, int k, const xc_metric_params *p, int greedy, int32_t *pred_idx,
                                      double *tp, double *fp, double *fn, double *tn, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !indptr || !order || !p || !pred_idx || !tp || !fp || !fn) return XC_ERR_INVALID;
    if (n <= 0 || m <= 0 || n_order < 0 || k < 1 || k > 32) return XC_ERR_INVALID;
    if (p->metric < 0 || p->metric > XC_METRIC_PREC_AT_K) return XC_ERR_INVALID;
    if (!p->skip_tn && !tn) return XC_ERR_INVALID;
    if (n_order == 0) return XC_OK;
    double *tn_arg = p->skip_tn ? nullptr : tn;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype != XC_F32 && dtype != XC_F64) return XC_ERR_UNSUPPORTED;
    static int generic_only = -1;   // $XCOLUMNS_B200_EXACT_CSR=generic forces the unstaged kernel
    if (generic_only < 0) {
        const char *e = getenv("XCOLUMNS_B200_EXACT_CSR");
        generic_only = (e && e[0] == 'g') ? 1 : 0;
    }
    if (generic_only) {
        if (dtype == XC_F32)
            bca_exact_csr_kernel<float><<<1, 32, 0, st>>>((const float *)data, indices, indptr, order, n_order, k, *p, greedy, pred_idx, tp, fp, fn, tn_arg);
        else
            bca_exact_csr_kernel<double><<<1, 32, 0, st>>>((const double *)data, indices, indptr, order, n_order, k, *p, greedy, pred_idx, tp, fp, fn, tn_arg);
        XC_LAUNCHED(ctx);
        return XC_OK;
    }
    // per-step metadata resolved in parallel: row start (int64), row length (int32), current prediction
    void *scratch = nullptr;
    const size_t off_nz = (size_t)n_order * 8, off_pred = off_nz + (((size_t)n_order * 4 + 7) & ~(size_t)7);
    int rc = xc_ctx_scratch(ctx, off_pred + (size_t)n_order * k * 4, &scratch);
    if (rc) return rc;
    int64_t *step_ts = (int64_t *)scratch;
    int32_t *step_nz = (int32_t *)((uint8_t *)scratch + off_nz);
    int32_t *step_pred = (int32_t *)((uint8_t *)scratch + off_pred);
    csr_step_meta_kernel<<<(unsigned)((n_order + 255) / 256), 256, 0, st>>>(order, indptr, pred_idx, k, n_order, step_ts,
                                                                            step_nz, step_pred);
    XC_LAUNCHED(ctx);
    if (dtype == XC_F32)
        bca_exact_csr_fast_kernel<float><<<1, 32, 0, st>>>((const float *)data, indices, order, step_ts, step_nz, step_pred, n_order, k, *p, greedy, pred_idx, tp, fp, fn, tn_arg);
    else
        bca_exact_csr_fast_kernel<double><<<1, 32, 0, st>>>((const double *)data, indices, order, step_ts, step_nz, step_pred, n_order, k, *p, greedy, pred_idx, tp, fp, fn, tn_arg);
    XC_LAUNCHED(ctx);
    return XC_OK;
}

extern "C" int xc_cov_exact_sweep_csr(xc_ctx *ctx, const void *data, int dtype, const int32_t *indices,
                                      const int64_t *indptr, int64_t n, int64_t m, const int32_t *order,
                                      int64_t n_order, int k, double alpha, int greedy, int32_t *pred_idx, double *Ef,
                                      void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !indptr || !order || !pred_idx || !Ef) return XC_ERR_INVALID;
    if (n <= 0 || m <= 0 || n_order < 0 || k < 1 || k > 32) return XC_ERR_INVALID;
    if (n_order == 0) return XC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == XC_F32)
        cov_exact_csr_kernel<float><<<1, 32, 0, st>>>((const float *)data, indices, indptr, order, n_order, k, alpha, greedy, pred_idx, Ef);
    else if (dtype == XC_F64)
        cov_exact_csr_kernel<double><<<1, 32, 0, st>>>((const double *)data, indices, indptr, order, n_order, k, alpha, greedy, pred_idx, Ef);
    else
        return XC_ERR_UNSUPPORTED;
    XC_LAUNCHED(ctx);
    return XC_OK;
}

extern "C" int xc_cov_state_csr(xc_ctx *ctx, const void *data, int dtype, const int32_t *indices,
                                const int64_t *indptr, int64_t n, int64_t m, const int32_t *pred_idx, int k, int order,
                                double *Ef, void *stream)
{
    XcDeviceGuard xc_guard__(ctx);
    if (!ctx || !indptr || !pred_idx || !Ef || n <= 0 || m <= 0 || k < 1 || k > 32) return XC_ERR_INVALID;
    if (dtype != XC_F32 && dtype != XC_F64) return XC_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    fill_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(Ef, 1.0, m);
    XC_LAUNCHED(ctx);
    if (order == XC_SUM_ORDERED) {
        if (dtype == XC_F32) cov_state_ordered_kernel<float><<<1, 32, 0, st>>>((const float *)data, indices, indptr, n, pred_idx, k, Ef);
        else cov_state_ordered_kernel<double><<<1, 32, 0, st>>>((const double *)data, indices, indptr, n, pred_idx, k, Ef);
    } else {
        int64_t blocks = (n * k + 255) / 256;
        int64_t cap = (int64_t)ctx->sm_count * 16;
        int grid = (int)(blocks < cap ? blocks : cap);
        if (dtype == XC_F32) cov_state_fast_kernel<float><<<grid, 256, 0, st>>>((const float *)data, indices, indptr, n, pred_idx, k, Ef);
        else cov_state_fast_kernel<double><<<grid, 256, 0, st>>>((const double *)data, indices, indptr, n, pred_idx, k, Ef);
    }
    XC_LAUNCHED(ctx);
    return XC_OK;
}
