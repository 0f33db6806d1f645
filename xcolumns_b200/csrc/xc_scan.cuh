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
template <typename TE, typename G, int R, bool SKIP, class Xf>
__device__ __forceinline__ void xc_scan_rows(const TE *const (&rp)[R], int64_t m, bool vec_ok, const Xf &xf,
                                             WarpTopK<G> (&tk)[R], const int (&old_idx)[R], int k)
{
    constexpr int V = XcVec<TE>::V;
    constexpr int STEP = 32 * V;  // columns one warp covers per 16-byte load
    const int lane = lane_id();
    const int64_t mv = vec_ok ? (m / V) * V : 0;
    const int64_t m2 = (mv / (2 * STEP)) * (2 * STEP);  // part covered by full, unguarded double steps
    const G qnan = (G)NAN;

    // ---- main loop: two 16-byte chunks per row in flight, no bounds checks, no local memory ----
    for (int64_t c0 = 0; c0 < m2; c0 += 2 * STEP) {
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

// ---- CTA-cooperative scan: coefficients staged through shared memory ------------------------------
// When the per-label coefficient vectors outgrow L1 (2 m sizeof(G) > ~160 KB, e.g. the Frank-Wolfe
// classifier at m = 31 k) every warp re-reads them from L2 for every row: coefficient traffic equals
// the HBM stream and the kernel sits at 0.70 of the copy bandwidth.  Here all warps of a CTA walk the
// column chunks in lock-step, the chunk's coefficients are copied once per CTA into a double-buffered
// shared-memory tile with cp.async (chunk c+1 lands while chunk c is consumed; one barrier per chunk)
// and are read back with conflict-free LDS.128.  One row per warp, so many independent CTAs per SM
// keep the HBM queue full while a warp is in the list-update slow path.
constexpr int XC_STAGE_BYTES = 16384;  // per buffer

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// gains = eta * a + b (XfMulAdd semantics, both vectors present), tile layout: G ca[TC] | G cb[TC]
template <typename G>
struct StageMulAdd {
    static constexpr int TC = XC_STAGE_BYTES / (2 * (int)sizeof(G));
    static constexpr int V = 16 / (int)sizeof(G);
    const G *a;
    const G *b;
    bool al16;  // a and b are 16-byte aligned (chunk starts are multiples of TC)

    // every thread of the CTA calls; copies columns [cbeg, min(m, cbeg + TC)) of a and b
    __device__ __forceinline__ void issue(char *buf, int64_t cbeg, int64_t m, int tid, int nthreads) const
    {
        G *sa = reinterpret_cast<G *>(buf), *sb = sa + TC;
        const int ncol = (int)min((int64_t)TC, m - cbeg);
        const G *ga = a + cbeg, *gb = b + cbeg;
        if (al16) {
            const int nv = ncol / V;
            for (int t = tid; t < nv; t += nthreads) {
                cp_async16(sa + t * V, ga + t * V);
                cp_async16(sb + t * V, gb + t * V);
            }
            for (int t = nv * V + tid; t < ncol; t += nthreads) {
                sa[t] = __ldg(ga + t);
                sb[t] = __ldg(gb + t);
            }
        } else {
            for (int t = tid; t < ncol; t += nthreads) {
                if (sizeof(G) == 4) {
                    cp_async4(sa + t, ga + t);
                    cp_async4(sb + t, gb + t);
                } else {
                    cp_async8(sa + t, ga + t);
                    cp_async8(sb + t, gb + t);
                }
            }
        }
    }
    // col: column inside the tile (multiple of V)
    template <typename TE>
    __device__ __forceinline__ void apply_vec(const char *buf, int col, const TE (&e)[V], G (&g)[V]) const
    {
        const G *sa = reinterpret_cast<const G *>(buf) + col;
        G ca[V], cb[V];
        const float4 u = *reinterpret_cast<const float4 *>(sa);
        const float4 w = *reinterpret_cast<const float4 *>(sa + TC);
        memcpy(ca, &u, 16);
        memcpy(cb, &w, 16);
#pragma unroll
        for (int v = 0; v < V; ++v) g[v] = XfMulAdd<G>::add_rn(XfMulAdd<G>::mul_rn((G)e[v], ca[v]), cb[v]);
    }
    template <typename TE>
    __device__ __forceinline__ G apply_one(const char *buf, int col, TE e) const
    {
        const G *sa = reinterpret_cast<const G *>(buf) + col;
        return XfMulAdd<G>::add_rn(XfMulAdd<G>::mul_rn((G)e, sa[0]), sa[TC]);
    }
};

// gains = fma(B_j, eta, A_j), interleaved float2 coefficients (batched BCA); tile: float2 coef[TC]
struct StageAffine {
    static constexpr int TC = XC_STAGE_BYTES / 8;
    const float2 *coef;  // 16-byte aligned (allocated by the host shim)

    __device__ __forceinline__ void issue(char *buf, int64_t cbeg, int64_t m, int tid, int nthreads) const
    {
        float2 *sc = reinterpret_cast<float2 *>(buf);
        const int ncol = (int)min((int64_t)TC, m - cbeg);
        const int nv = ncol / 2;
        for (int t = tid; t < nv; t += nthreads) cp_async16(sc + 2 * t, coef + cbeg + 2 * t);
        if ((ncol & 1) && tid == 0) sc[ncol - 1] = __ldg(coef + cbeg + ncol - 1);
    }
    template <typename TE, int V>
    __device__ __forceinline__ void apply_vec(const char *buf, int col, const TE (&e)[V], float (&g)[V]) const
    {
        const float2 *sc = reinterpret_cast<const float2 *>(buf) + col;
#pragma unroll
        for (int v = 0; v < V; v += 2) {
            const float4 u = *reinterpret_cast<const float4 *>(sc + v);
            g[v] = fmaf(u.x, (float)e[v], u.y);
            g[v + 1] = fmaf(u.z, (float)e[v + 1], u.w);
        }
    }
    template <typename TE>
    __device__ __forceinline__ float apply_one(const char *buf, int col, TE e) const
    {
        const float2 t = reinterpret_cast<const float2 *>(buf)[col];
        return fmaf(t.x, (float)e, t.y);
    }
};

// One chunk of one row: columns [cbeg, cend) inside one tile, coefficients in `buf`, tile_off = column of
// cbeg inside the tile.  Rows must be 16-byte aligned (vec_ok); unaligned inputs use the unstaged kernels.
template <typename TE, typename G, bool SKIP, class Stage>
__device__ __forceinline__ void xc_scan_chunk_staged(const TE *rp, int64_t cbeg, int64_t cend, const Stage &st,
                                                     const char *buf, WarpTopK<G> &tk, int old_idx, int k,
                                                     int tile_off = 0)
{
    constexpr int V = XcVec<TE>::V;
    constexpr int STEP = 32 * V;
    const int lane = lane_id();
    const int len = (int)(cend - cbeg);
    const int len2 = (len / (2 * STEP)) * (2 * STEP);
    const int lenv = (len / V) * V;
    const G qnan = (G)NAN;
    const TE *p = rp + cbeg;
    for (int c0 = 0; c0 < len2; c0 += 2 * STEP) {
        const int cA = c0 + lane * V, cB = cA + STEP;
        TE eA[V], eB[V];
        XcVec<TE>::load(p + cA, eA);
        XcVec<TE>::load(p + cB, eB);
        G gA[V], gB[V];
        st.template apply_vec<TE>(buf, tile_off + cA, eA, gA);
        st.template apply_vec<TE>(buf, tile_off + cB, eB, gB);
        bool hit = tk.passes(xc_vmax<G, V>(gA)) | tk.passes(xc_vmax<G, V>(gB));
        if (__any_sync(XC_FULL, hit)) {
            xc_scan_insert<G, V, SKIP>(tk, gA, cbeg + c0, V, k, old_idx);
            xc_scan_insert<G, V, SKIP>(tk, gB, cbeg + c0 + STEP, V, k, old_idx);
        }
    }
    for (int c0 = len2; c0 < lenv; c0 += STEP) {
        const int c = c0 + lane * V;
        G g[V];
        if (c < lenv) {
            TE e[V];
            XcVec<TE>::load(p + c, e);
            st.template apply_vec<TE>(buf, tile_off + c, e, g);
        } else {
#pragma unroll
            for (int v = 0; v < V; ++v) g[v] = qnan;
        }
        if (__any_sync(XC_FULL, tk.passes(xc_vmax<G, V>(g)))) xc_scan_insert<G, V, SKIP>(tk, g, cbeg + c0, V, k, old_idx);
    }
    for (int c0 = lenv; c0 < len; c0 += 32) {
        const int c = c0 + lane;
        G g1[1];
        g1[0] = (c < len) ? st.template apply_one<TE>(buf, tile_off + c, ld_stream(p + c)) : qnan;
        if (__any_sync(XC_FULL, tk.passes(g1[0]))) xc_scan_insert<G, 1, SKIP>(tk, g1, cbeg + c0, 1, k, old_idx);
    }
}

// Flat (row group, chunk) pipeline state shared by the staged kernels.  Usage per CTA:
//   XcStagePipe<Stage> pipe(st, m, smem);  pipe.prime();
//   for each row group:  for (c = 0; c < pipe.nch; ++c) { const char *buf = pipe.acquire(more_work_follows);
//                                                         xc_scan_chunk_staged(..., buf, ...); }
template <class Stage>
struct XcStagePipe {
    const Stage &st;
    const int64_t m;
    char *smem;  // 2 * XC_STAGE_BYTES
    int nch;
    int next_c;   // chunk id the next issue() will copy
    unsigned it;  // flat step counter (buffer = it & 1)

    __device__ __forceinline__ XcStagePipe(const Stage &s, int64_t m_, char *smem_) : st(s), m(m_), smem(smem_)
    {
        nch = (int)((m + Stage::TC - 1) / Stage::TC);
        next_c = 0;
        it = 0;
    }
    __device__ __forceinline__ void prime()
    {
        st.issue(smem, 0, m, threadIdx.x, blockDim.x);
        cp_async_commit();
        next_c = nch > 1 ? 1 : 0;
    }
    // wait for the current chunk's tile, start the copy of the following one; returns the tile to read
    __device__ __forceinline__ const char *acquire(bool more)
    {
        cp_async_wait_all();
        __syncthreads();  // tile `it` is visible to everyone AND everyone is done with tile `it - 1`
        char *cur = smem + (it & 1) * XC_STAGE_BYTES;
        if (more) {
            st.issue(smem + ((it + 1) & 1) * XC_STAGE_BYTES, (int64_t)next_c * Stage::TC, m, threadIdx.x, blockDim.x);
            cp_async_commit();
            next_c = next_c + 1 == nch ? 0 : next_c + 1;
        }
        ++it;
        return cur;
    }
};

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
