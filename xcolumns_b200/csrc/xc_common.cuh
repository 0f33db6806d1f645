// Shared device/host helpers for the xcolumns_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/xcolumns_b200.h"

constexpr int XC_PIPE_MAX_LAG = 3;   // pipelined sweeps: at most 4 batches in flight

struct xc_ctx {
    int device;
    int sm_count;
    int64_t launches;
    cudaError_t last_err;
    // scratch for the cooperative sequential-exact sweep (candidate exchange + barrier)
    void *scratch;
    size_t scratch_bytes;
    int coop_blocks_cache[8];
    // scratch of the deterministic multi-block reductions (block partials + ticket counter)
    double *red_partials;
    unsigned *red_counter;
    // pipelined batched sweep (bca_batched.cu): two internal streams + ordering events, created on first use
    cudaStream_t aux[XC_PIPE_MAX_LAG + 1];   // batch kernels, one stream per batch in flight (low priority)
    cudaStream_t cstream;                    // commits + sweep utilities (high priority: dispatched ahead of pending batch CTAs)
    cudaStream_t pstream[XC_PIPE_MAX_LAG + 1];   // pushes of the batch deltas to the peers (high priority, one per batch stream)
    cudaEvent_t ev_p[XC_PIPE_MAX_LAG + 1];
    cudaEvent_t ev_fork, ev_c[XC_PIPE_MAX_LAG + 1], ev_join[XC_PIPE_MAX_LAG + 2], ev_k[XC_PIPE_MAX_LAG + 1], ev_pro, ev_util;
    bool aux_ready;
    bool pipe_active;            // sweeps issued since the last join: the coefficient sets follow the commits
    bool pipe_forked;            // ... on the internal streams (work the caller's stream has not joined yet)
    // optional per-launch timing of the batch kernels (xc_timing_*): events in launch order
    bool timing_on;
    bool timing_commits;         // xc_timing_enable(2): commits are logged too (rows = 0)
    int timing_count, timing_cap;
    cudaEvent_t *timing_ev;      // 2 per launch (start, end)
    int64_t *timing_rows;        // rows of the launch
    // pinned staging buffers of xc_h2d_staged (host.cu), allocated on first use
    void *stage[2];
    cudaEvent_t stage_ev[2];
};

constexpr int XC_RED_MAX_BLOCKS = 1024;

// Peer-memory window of one rank (csrc/p2p.cu): cudaMalloc'ed, exported with CUDA IPC and mapped by
// every other rank of the box.  Header layout (32-bit words): flag[buffer][sender] at buffer * XC_P2P_MAX_WORLD + sender
// for the XC_P2P_MAX_BUF rotating delta buffers (a flag holds the epoch of the last push into that buffer slot; pushes
// of different buffers may complete out of order, pushes into one buffer are stream-ordered), the error word, one
// ticket counter per buffer (the pushes of different buffers run concurrently); [XC_P2P_HEADER, ...) payload.
constexpr int XC_P2P_MAX_WORLD = 16;
constexpr int XC_P2P_MAX_BUF = 8;                                   // 2 * (XC_PIPE_MAX_LAG + 1)
constexpr int XC_P2P_ERR_WORD = XC_P2P_MAX_WORLD * XC_P2P_MAX_BUF;  // 128
constexpr int XC_P2P_TICKET_WORD = XC_P2P_ERR_WORD + 1;             // 129 .. 136
constexpr int XC_P2P_HEADER = 1024;
struct xc_p2p {
    int world;
    int rank;
    size_t bytes;
    uint8_t *windows[XC_P2P_MAX_WORLD];  // windows[rank] is the local allocation, the others are IPC mappings
    uint8_t **windows_dev;               // the same table in device memory
    unsigned epoch;                      // commits performed so far (identical on every rank)
    bool opened;
};

#define XC_FULL 0xffffffffu

#define XC_CUDA_TRY(ctx, expr)                    \
    do {                                          \
        cudaError_t e__ = (expr);                 \
        if (e__ != cudaSuccess) {                 \
            (ctx)->last_err = e__;                \
            return XC_ERR_CUDA;                   \
        }                                         \
    } while (0)

// call after every kernel launch: counts it and catches launch-configuration errors
#define XC_LAUNCHED(ctx)                          \
    do {                                          \
        (ctx)->launches++;                        \
        cudaError_t e__ = cudaGetLastError();     \
        if (e__ != cudaSuccess) {                 \
            (ctx)->last_err = e__;                \
            return XC_ERR_CUDA;                   \
        }                                         \
    } while (0)

int xc_ctx_scratch(xc_ctx *ctx, size_t bytes, void **out);
int xc_ctx_aux_streams(xc_ctx *ctx);
int xc_timing_slot(xc_ctx *ctx, int64_t rows, cudaEvent_t *start, cudaEvent_t *end);
int xc_utility_launch(xc_ctx *ctx, const xc_metric_params *p, int agg, const double *tp, const double *fp,
                      const double *fn, const double *tn, double tn_rows, int64_t m, double *out_dev, cudaStream_t st);

// Every entry point runs on the context's device whatever the caller's current device is, and leaves the
// caller's current device as it found it (a single process may drive several GPUs, one context each).
struct XcDeviceGuard {
    int prev;
    bool switched;
    explicit XcDeviceGuard(const xc_ctx *ctx) : prev(-1), switched(false)
    {
        if (!ctx) return;
        if (cudaGetDevice(&prev) == cudaSuccess && prev != ctx->device) switched = cudaSetDevice(ctx->device) == cudaSuccess;
    }
    ~XcDeviceGuard()
    {
        if (switched) cudaSetDevice(prev);
    }
    XcDeviceGuard(const XcDeviceGuard &) = delete;
    XcDeviceGuard &operator=(const XcDeviceGuard &) = delete;
};

static inline bool xc_aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------------------------------------
// streaming loads: read-only path, do not allocate in L1 (each byte of eta is used once)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_stream_f4(const float *p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ double2 ld_stream_d2(const double *p)
{
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_stream(const float *p)
{
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ double ld_stream(const double *p)
{
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// (gain, label) ordering used everywhere: larger gain first, ties -> lower label id
template <typename G>
__device__ __forceinline__ bool xc_better(G g1, int j1, G g2, int j2)
{
    return (g1 > g2) || (g1 == g2 && j1 < j2);
}

// ---------------------------------------------------------------------------------------------
// Warp-distributed top-k list (k <= 32): lane l holds the l-th best (gain, label).
// Empty slots hold (-inf, INT_MAX).  thr / thr_j mirror slot k-1 in every lane.
// ---------------------------------------------------------------------------------------------
template <typename G>
struct WarpTopK {
    G val;
    int idx;
    G thr;
    int thr_j;

    __device__ __forceinline__ void init()
    {
        val = -INFINITY;
        idx = 0x7fffffff;
        thr = -INFINITY;
        thr_j = 0x7fffffff;
    }
    // warp-uniform (g, j); all 32 lanes must call
    __device__ __forceinline__ void insert(G g, int j, int k)
    {
        const int lane = lane_id();
        bool ahead = (lane < k) && xc_better(val, idx, g, j);
        int pos = __popc(__ballot_sync(XC_FULL, ahead));
        if (pos < k) {  // warp-uniform
            G pv = __shfl_up_sync(XC_FULL, val, 1);
            int pi = __shfl_up_sync(XC_FULL, idx, 1);
            if (lane == pos) {
                val = g;
                idx = j;
            } else if (lane > pos && lane < k) {
                val = pv;
                idx = pi;
            }
            // the threshold only ever tightens: while the list is still filling, a bound set by prime()
            // (k elements >= thr are known to exist) stays in force
            const G nv = __shfl_sync(XC_FULL, val, k - 1);
            const int nj = __shfl_sync(XC_FULL, idx, k - 1);
            if (nv > thr || (nv == thr && nj < thr_j)) {
                thr = nv;
                thr_j = nj;
            }
        }
    }
    // Lower bound for the final k-th best from per-lane maxima of values that WILL be offered to the list:
    // the k-th largest of the 32 lane maxima has at least k offered values at or above it.  Without it an
    // empty list sends every lane of the first chunks through the serial insertion path.
    // lane_max: NaN-free maximum of this lane's values (-inf if it has none); all 32 lanes must call.
    __device__ __forceinline__ void prime(G lane_max, int k)
    {
        G m = lane_max;
        G bound = -INFINITY;
        for (int r = 0; r < k; ++r) {
            G w = m;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) w = fmax(w, __shfl_xor_sync(XC_FULL, w, o));
            bound = w;  // r-th largest lane maximum (with multiplicity)
            // retire ONE lane holding it
            const unsigned bal = __ballot_sync(XC_FULL, m == w);
            if (lane_id() == __ffs(bal) - 1) m = -INFINITY;
        }
        if (bound > thr) {
            thr = bound;
            thr_j = 0x7fffffff;  // ties with the bound still enter
        }
    }
    // does (g, j) have a chance to enter?  (cheap per-lane filter)
    __device__ __forceinline__ bool passes(G g) const { return g >= thr; }
};

// sort the first k lanes' labels ascending; returns in lane r (< k) the label with rank r.
// Invalid labels (INT_MAX) sort last.
__device__ __forceinline__ int warp_sort_labels(int idx, int k)
{
    const int lane = lane_id();
    int rank = 0;
    for (int t = 0; t < k; ++t) {
        int o = __shfl_sync(XC_FULL, idx, t);
        rank += (o < idx) || (o == idx && t < lane);
    }
    int out = 0x7fffffff;
    for (int t = 0; t < k; ++t) {
        unsigned bal = __ballot_sync(XC_FULL, lane < k && rank == t);
        int src = __ffs(bal) - 1;
        int v = __shfl_sync(XC_FULL, idx, src < 0 ? 0 : src);
        if (lane == t) out = v;
    }
    return out;
}

// Deterministic grid-wide sum: every block stores its partial, the block that draws the last ticket
// adds the partials in block order (fixed order -> bit-reproducible) and re-arms the counter.
// Returns true in thread 0 of that last block with the total in *total.
__device__ __forceinline__ bool xc_grid_sum_last(double block_partial, double *partials, unsigned *counter,
                                                 double *total)
{
    __shared__ bool s_last;
    __shared__ double s_total;
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = block_partial;
        __threadfence();
        unsigned t = atomicAdd(counter, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return false;
    if (threadIdx.x < 32) {
        double v = 0.0;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += 32) v += __ldcg(partials + b);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(XC_FULL, v, o);
        if (threadIdx.x == 0) {
            s_total = v;
            *counter = 0;
        }
    }
    __syncthreads();
    *total = s_total;
    return threadIdx.x == 0;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(XC_FULL, v, o);
    return v;
}

// binary metrics in float64, operation order of xcolumns/metrics.py (see include/xcolumns_b200.h)
__device__ __forceinline__ double xc_binary_metric(int metric, double tp, double fp, double fn, double tn,
                                                   double c1, double beta2, double eps)
{
    switch (metric) {
    case XC_METRIC_PRECISION: return tp / ((tp + fp) + eps);
    case XC_METRIC_RECALL: return tp / ((tp + fn) + eps);
    case XC_METRIC_FBETA: return (c1 * tp) / ((((beta2 * (tp + fp)) + tp) + fn) + eps);
    case XC_METRIC_JACCARD: return tp / (((tp + fp) + fn) + eps);
    case XC_METRIC_PREC_AT_K: return tp / c1;
    default: {
        double tpr = tp / ((tp + fn) + eps);
        double tnr = tn / ((tn + fp) + eps);
        if (metric == XC_METRIC_BALANCED_ACC) return (tpr + tnr) / 2.0;
        if (metric == XC_METRIC_GMEAN) return sqrt(tpr * tnr);
        return ((2.0 * tpr) * tnr) / (tpr + tnr);
    }
    }
}

// the binary metric of the call, including the mixed utilities
//   (1 - alpha) * binary_precision_at_k(tp, k) + alpha * metric / m     (python evaluation order)
__device__ __forceinline__ double xc_metric_eval(const xc_metric_params &p, double tp, double fp, double fn, double tn)
{
    double v = xc_binary_metric(p.metric, tp, fp, fn, tn, p.c1, p.beta2, p.eps);
    if (p.mix == 1) v = ((1.0 - p.mix_alpha) * (tp / p.mix_k)) + ((p.mix_alpha * v) / p.mix_m);
    // mix == 3 (frank_wolfe.py:917-938): (1 - alpha) * recall + alpha * precision, metric id = precision
    if (p.mix == 3) v = ((1.0 - p.mix_alpha) * (tp / ((tp + fn) + p.eps))) + (p.mix_alpha * v);
    return v;
}
