// Streaming row scan shared by the weighted top-k, batched-BCA and Frank-Wolfe kernels.
//
// One warp streams R rows of the probability matrix at a time with 128-bit loads
// (ld.global.nc.L1::no_allocate), turns every element into a gain with a per-label transform
// and keeps, per row, a warp-distributed top-k list.  Coefficient vectors are loaded once per
// column chunk and reused for the R rows in registers, so the L1/L2 traffic for them is 1/R of
// the HBM stream.  The common case per 16-byte chunk is: 4 gains, 3 max, 1 compare; the list
// is only touched when some lane beats the current k-th best (rare once the list is warm).
#pragma once
#include "xc_common.cuh"

template <typename TE> struct XcVec;
template <> struct XcVec<float> {
    static constexpr int V = 4;
    __device__ static __forceinline__ void load(const float *p, float (&e)[4])
    {
        float4 v = ld_stream_f4(p);
        e[0] = v.x; e[1] = v.y; e[2] = v.z; e[3] = v.w;
    }
};
template <> struct XcVec<double> {
    static constexpr int V = 2;
    __device__ static __forceinline__ void load(const double *p, double (&e)[2])
    {
        double2 v = ld_stream_d2(p);
        e[0] = v.x; e[1] = v.y;
    }
};

// ---- per-label transforms -----------------------------------------------------------------
// gains = eta [* a] [+ b], separate IEEE multiply and add in G (numpy semantics, no FMA).
template <typename G>
struct XfMulAdd {
    const G *a;
    const G *b;
    template <typename TE, int V>
    __device__ __forceinline__ void apply_vec(int64_t c, const TE (&e)[V], G (&g)[V], G (&ca)[V], G (&cb)[V],
                                              bool first) const
    {
        if (first) {  // coefficients for this column chunk, shared by the R rows
#pragma unroll
            for (int v = 0; v < V; ++v) {
                ca[v] = a ? __ldg(a + c + v) : G(1);
                cb[v] = b ? __ldg(b + c + v) : G(0);
            }
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
            G x = (G)e[v];
            if (a) x = mul_rn(x, ca[v]);
            if (b) x = add_rn(x, cb[v]);
            g[v] = x;
        }
    }
    template <typename TE>
    __device__ __forceinline__ G apply_one(int64_t c, TE e) const
    {
        G x = (G)e;
        if (a) x = mul_rn(x, __ldg(a + c));
        if (b) x = add_rn(x, __ldg(b + c));
        return x;
    }
    __device__ static __forceinline__ float mul_rn(float x, float y) { return __fmul_rn(x, y); }
    __device__ static __forceinline__ float add_rn(float x, float y) { return __fadd_rn(x, y); }
    __device__ static __forceinline__ double mul_rn(double x, double y) { return __dmul_rn(x, y); }
    __device__ static __forceinline__ double add_rn(double x, double y) { return __dadd_rn(x, y); }
};

// Same arithmetic, both vectors present and 16-byte aligned: one 128-bit load per vector and chunk
// instead of 2 V scalar loads and no null checks in the streaming loop (measured on the Frank-Wolfe
// iterate at 14 k x 31 k: the scalar version kept the LSU pipe 68 % busy and capped the kernel at
// 0.78 of the copy bandwidth).
template <typename G>
struct XfMulAddVec {
    const G *a;
    const G *b;
    template <typename TE, int V>
    __device__ __forceinline__ void apply_vec(int64_t c, const TE (&e)[V], G (&g)[V], G (&ca)[V], G (&cb)[V],
                                              bool first) const
    {
        static_assert(V * sizeof(G) == 16 || V * sizeof(G) == 32, "one or two 16-byte loads per vector");
        if (first) {
            constexpr int Q = (int)(V * sizeof(G) / 16);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const float4 u = __ldg(reinterpret_cast<const float4 *>(a + c) + q);
                const float4 w = __ldg(reinterpret_cast<const float4 *>(b + c) + q);
                memcpy(reinterpret_cast<char *>(ca) + 16 * q, &u, 16);
                memcpy(reinterpret_cast<char *>(cb) + 16 * q, &w, 16);
            }
        }
#pragma unroll
        for (int v = 0; v < V; ++v) g[v] = XfMulAdd<G>::add_rn(XfMulAdd<G>::mul_rn((G)e[v], ca[v]), cb[v]);
    }
    template <typename TE>
    __device__ __forceinline__ G apply_one(int64_t c, TE e) const
    {
        return XfMulAdd<G>::add_rn(XfMulAdd<G>::mul_rn((G)e, __ldg(a + c)), __ldg(b + c));
    }
};

// gains = fma(B_j, eta, A_j) with interleaved float2 coefficients (B_j, A_j) (batched BCA)
struct XfAffine {
    const float2 *coef;
    template <typename TE, int V>
    __device__ __forceinline__ void apply_vec(int64_t c, const TE (&e)[V], float (&g)[V], float (&ca)[V],
                                              float (&cb)[V], bool first) const
    {
        if (first) {
            if (V == 4) {
                const float4 *p = reinterpret_cast<const float4 *>(coef + c);
                float4 u = __ldg(p), w = __ldg(p + 1);
                ca[0] = u.x; cb[0] = u.y; ca[1] = u.z; cb[1] = u.w;
                ca[2] = w.x; cb[2] = w.y; ca[3] = w.z; cb[3] = w.w;
            } else {
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    float2 t = __ldg(coef + c + v);
                    ca[v] = t.x; cb[v] = t.y;
                }
            }
        }
#pragma unroll
        for (int v = 0; v < V; ++v) g[v] = fmaf(ca[v], (float)e[v], cb[v]);
    }
    template <typename TE>
    __device__ __forceinline__ float apply_one(int64_t c, TE e) const
    {
        float2 t = __ldg(coef + c);
        return fmaf(t.x, (float)e, t.y);
    }
};

// ---- slow path: some lane holds a candidate for row-list tk ---------------------------------
// Inlined on purpose: taking references to the register-resident list / gains from a real call
// would force them into local memory in the streaming loop (measured: 8 STL.128 per trip).
template <typename G, int V, bool SKIP>
__device__ __forceinline__ void xc_scan_insert(WarpTopK<G> &tk, const G (&g)[V], int64_t cbase, int stride, int k,
                                               int old_idx)
{
    const int lane = lane_id();
    bool hit = false;
#pragma unroll
    for (int v = 0; v < V; ++v) hit |= tk.passes(g[v]);
    unsigned bal = __ballot_sync(XC_FULL, hit);
    while (bal) {
        int src = __ffs(bal) - 1;
        bal &= bal - 1;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            G gv = __shfl_sync(XC_FULL, g[v], src);
            int j = (int)(cbase + (int64_t)src * stride + v);
            if (xc_better(gv, j, tk.thr, tk.thr_j)) {  // warp-uniform
                if (SKIP) {
                    if (__any_sync(XC_FULL, lane < k && old_idx == j)) continue;
                }
                tk.insert(gv, j, k);
            }
        }
    }
}

template <typename G, int V>
__device__ __forceinline__ G xc_vmax(const G (&g)[V])
{
    G mx = g[0];
#pragma unroll
    for (int v = 1; v < V; ++v) mx = fmax(mx, g[v]);  // FMNMX drops NaN operands; all-NaN stays NaN
    return mx;
}

// ---- the scan ---------------------------------------------------------------------------------
// rp[r]: start of row r; vec_ok: rows are 16-byte aligned (base aligned and ld % V == 0).
// old_idx[r]: (SKIP) lanes < k hold the labels already seeded into tk[r]; candidates equal to
// one of them are ignored (their gain under the "selected" formula is already in the list).
// DEEP: four 16-byte chunks per row in flight instead of two (R = 1 only).  At full occupancy the extra registers
// cost a resident CTA and the kernel gets slower (measured in round 1); a launch that cannot fill the GPU anyway
// (a batch of less than one wave of rows: strong scaling, tiny commits) is bound by the bytes each warp keeps in
// flight, and there the deeper loop wins.
template <typename TE, typename G, int R, bool SKIP, class Xf, bool DEEP = false>
__device__ __forceinline__ void xc_scan_rows(const TE *const (&rp)[R], int64_t m, bool vec_ok, const Xf &xf,
                                             WarpTopK<G> (&tk)[R], const int (&old_idx)[R], int k)
{
    constexpr int V = XcVec<TE>::V;
    constexpr int STEP = 32 * V;  // columns one warp covers per 16-byte load
    const int lane = lane_id();
    const int64_t mv = vec_ok ? (m / V) * V : 0;
    const int64_t m2 = (mv / (2 * STEP)) * (2 * STEP);  // part covered by full, unguarded double steps
    const int64_t m4 = (DEEP && R == 1) ? (mv / (4 * STEP)) * (4 * STEP) : 0;  // ... by quadruple steps
    const G qnan = (G)NAN;

    // ---- lists built from scratch (no seed): bound the threshold from the first double step ----
    if (!SKIP && m2 > 0) {
        const int64_t cA = (int64_t)lane * V;
        G ca[V], cb[V];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            TE eA[V], eB[V];
            XcVec<TE>::load(rp[r] + cA, eA);
            XcVec<TE>::load(rp[r] + cA + STEP, eB);
            G gA[V], gB[V];
            xf.template apply_vec<TE, V>(cA, eA, gA, ca, cb, true);
            xf.template apply_vec<TE, V>(cA + STEP, eB, gB, ca, cb, true);
            G mx = fmax(xc_vmax<G, V>(gA), xc_vmax<G, V>(gB));
            tk[r].prime(mx == mx ? mx : (G)-INFINITY, k);
        }
    }

    // ---- main loop: two 16-byte chunks per row in flight, no bounds checks, no local memory ----
    // (a generalised loop with 4 chunks in flight was measured slower on the Frank-Wolfe iterate -- 350 vs
    // 334 us at 14 k x 31 k -- and its code shape cost the batched-BCA kernel 4 registers / one CTA per SM)
    if (DEEP && R == 1) {
        for (int64_t c0 = 0; c0 < m4; c0 += 4 * STEP) {
            const int64_t cA = c0 + (int64_t)lane * V;
            TE e[4][V];
#pragma unroll
            for (int q = 0; q < 4; ++q) XcVec<TE>::load(rp[0] + cA + q * STEP, e[q]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {   // two pairs, each handled like one double step of the loop below
                G gA[V], gB[V], ca[V], cb[V];
                xf.template apply_vec<TE, V>(cA + (2 * h) * STEP, e[2 * h], gA, ca, cb, true);
                bool hit = tk[0].passes(xc_vmax<G, V>(gA));
                xf.template apply_vec<TE, V>(cA + (2 * h + 1) * STEP, e[2 * h + 1], gB, ca, cb, true);
                hit |= tk[0].passes(xc_vmax<G, V>(gB));
                if (__any_sync(XC_FULL, hit)) {
                    xc_scan_insert<G, V, SKIP>(tk[0], gA, c0 + (2 * h) * STEP, V, k, old_idx[0]);
                    xc_scan_insert<G, V, SKIP>(tk[0], gB, c0 + (2 * h + 1) * STEP, V, k, old_idx[0]);
                }
            }
        }
    }
    for (int64_t c0 = m4; c0 < m2; c0 += 2 * STEP) {
        const int64_t cA = c0 + (int64_t)lane * V;
        const int64_t cB = cA + STEP;
        TE eA[R][V], eB[R][V];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            XcVec<TE>::load(rp[r] + cA, eA[r]);
            XcVec<TE>::load(rp[r] + cB, eB[r]);
        }
        G gA[R][V], gB[R][V], ca[V], cb[V];
        bool hit = false;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            xf.template apply_vec<TE, V>(cA, eA[r], gA[r], ca, cb, r == 0);
            hit |= tk[r].passes(xc_vmax<G, V>(gA[r]));
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            xf.template apply_vec<TE, V>(cB, eB[r], gB[r], ca, cb, r == 0);
            hit |= tk[r].passes(xc_vmax<G, V>(gB[r]));
        }
        if (__any_sync(XC_FULL, hit)) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                xc_scan_insert<G, V, SKIP>(tk[r], gA[r], c0, V, k, old_idx[r]);
                xc_scan_insert<G, V, SKIP>(tk[r], gB[r], c0 + STEP, V, k, old_idx[r]);
            }
        }
    }
    // ---- guarded single steps for the rest of the vectorisable part ------------------------------
    for (int64_t c0 = m2; c0 < mv; c0 += STEP) {
        const int64_t c = c0 + (int64_t)lane * V;
        const bool in = c < mv;
        G g[R][V], ca[V], cb[V];
        bool hit = false;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (in) {
                TE e[V];
                XcVec<TE>::load(rp[r] + c, e);
                xf.template apply_vec<TE, V>(c, e, g[r], ca, cb, r == 0);
            } else {
#pragma unroll
                for (int v = 0; v < V; ++v) g[r][v] = qnan;
            }
            hit |= tk[r].passes(xc_vmax<G, V>(g[r]));
        }
        if (__any_sync(XC_FULL, hit)) {
#pragma unroll
            for (int r = 0; r < R; ++r) xc_scan_insert<G, V, SKIP>(tk[r], g[r], c0, V, k, old_idx[r]);
        }
    }
    // ---- scalar tail (and the whole row when it is not 16-byte aligned): one column per lane ----
    for (int64_t c0 = mv; c0 < m; c0 += 32) {
        const int64_t c = c0 + lane;
        G g1[R][1];
        bool hit = false;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            g1[r][0] = (c < m) ? xf.template apply_one<TE>(c, ld_stream(rp[r] + c)) : qnan;
            hit |= tk[r].passes(g1[r][0]);
        }
        if (__any_sync(XC_FULL, hit)) {
#pragma unroll
            for (int r = 0; r < R; ++r) xc_scan_insert<G, 1, SKIP>(tk[r], g1[r], c0, 1, k, old_idx[r]);
        }
    }
}

// rank-sort helper: lane t (< k) learns which lane holds the label of rank t (ascending label id)
__device__ __forceinline__ int warp_rank_src(int idx, int k)
{
    const int lane = lane_id();
    int rank = 0;
    for (int t = 0; t < k; ++t) {
        int o = __shfl_sync(XC_FULL, idx, t);
        rank += (o < idx) || (o == idx && t < lane);
    }
    int src = 0;
    for (int t = 0; t < k; ++t) {
        unsigned bal = __ballot_sync(XC_FULL, lane < k && rank == t);
        if (lane == t) src = __ffs(bal) - 1;
    }
    return src;
}
