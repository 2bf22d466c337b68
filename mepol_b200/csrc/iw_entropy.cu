// K2 / K3 / K4: importance-weighted k-NN entropy + KL (forward), its hand-written backward, and
// the reverse-neighbour CSR that makes the backward a deterministic gather.
//
// Reference arithmetic being replaced (paths relative to /root/reference/):
//   src/algorithms/mepol.py:121-138  u_j = exp(cumsum_t(lt - lb)) per trajectory, w = u / sum(u)
//   src/algorithms/mepol.py:148      W_i = sum_{m<k} w[idx[i, m]]            (indices[:, :-1])
//   src/algorithms/mepol.py:150      V_i = r_k^ns * pi^(ns/2) / Gamma(ns/2+1)
//   src/algorithms/mepol.py:152      H   = -sum_i (W_i/k) log(W_i/(V_i+eps) + eps) + B
//   src/algorithms/mepol.py:165      KL  = (1/N) sum_i log(k/(N W_i) + eps)
//   loss.backward() (mepol.py:278)   closed form, SURVEY.md section 8a.
//
// Data flow (all arrays length N unless noted; T = float or double storage):
//   iw_scan    (1 CTA / trajectory): c = segmented inclusive cumsum(lt - lb)            -> c[N]
//              per-trajectory max m_t and sum exp(c - m_t); the LAST CTA to finish folds them into
//              scal[3] = M = max c, scal[4] = S = sum exp(c - M)     (fixed order: deterministic)
//   iw_gather  (1 thread / row): W_i = sum_m exp(c[idx[i,m]] - M)/S; row terms of H and KL;
//              gW_i = dH/dW_i; block partials of {H, KL, sum gW_i W_i}; the last CTA reduces the
//              partials in fixed order into scal[0..2].
//   iw_bwd     (1 CTA / trajectory): g_w[j] = sum_{e in rev[j]} gW[col[e]]  (reverse CSR gather),
//              g_c[j] = w_j (g_w[j] - s), grad_lt = grad_H * reverse segmented cumsum(g_c).
//              s = sum_j g_w[j] w_j = sum_i gW_i W_i is already known from the forward, so the
//              backward needs no grid-wide dependency.
// Algorithmic HBM bytes per state at k=4, fp32 storage: scan 12, gather 8+8k+... see
// DESIGN.md section "IW kernels".
#include <math.h>

#include "common.cuh"

namespace mepol {

constexpr int kScanThreads = 256;
constexpr int kGatherThreads = 256;
constexpr int kGatherRows = 4;      // rows per thread of the gather kernel (memory-level parallelism)
constexpr double kLogPi = 1.1447298858494001741434273513531;

struct IwWorkspace {
    double *traj_max;   // [n_traj]
    double *traj_sum;   // [n_traj]
    double *blk_part;   // [n_gather_blocks * 4]
    unsigned *counters; // [2] last-block-done tickets (zeroed by the forward launcher)
    size_t bytes;
};

static int gather_blocks_for(int64_t n) {
    const int64_t per = (int64_t)kGatherThreads * kGatherRows;
    return (int)((n + per - 1) / per);
}

static void carve_iw(IwWorkspace &w, int64_t n, int64_t n_traj, void *ws) {
    Carver c(ws);
    w.traj_max = c.take<double>((size_t)n_traj);
    w.traj_sum = c.take<double>((size_t)n_traj);
    w.blk_part = c.take<double>((size_t)gather_blocks_for(n) * 4);
    w.counters = c.take<unsigned>(4);
    w.bytes = c.bytes();
}

// Block-wide inclusive scan of one value per thread (sum), returns the thread's inclusive prefix;
// total over the block is left in *block_total (valid for every thread after the call).
__device__ __forceinline__ double block_inclusive_scan(double v, double *s_warp, double *block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double t = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double u = __shfl_up_sync(0xffffffffu, t, o);
            if (lane >= o) t += u;
        }
        s_warp[lane] = t;  // inclusive warp totals
    }
    __syncthreads();
    const double prev = warp > 0 ? s_warp[warp - 1] : 0.0;
    *block_total = s_warp[(blockDim.x >> 5) - 1];
    __syncthreads();
    return v + prev;
}

__device__ __forceinline__ double block_reduce_sum(double v, double *s_warp) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_warp[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (warp == 0) {
        t = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0.0;
        t = warp_sum(t);
    }
    return t;  // valid in warp 0
}
__device__ __forceinline__ double block_reduce_max(double v, double *s_warp) {
    v = warp_max(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_warp[warp] = v;
    __syncthreads();
    double t = -__longlong_as_double(0x7ff0000000000000LL);
    if (warp == 0) {
        t = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : t;
        t = warp_max(t);
    }
    return t;
}


// exp / log at the STORAGE precision of the log-probs (fp32 kernels: MUFU-based expf / logf; sums stay in float64)
template <typename T>
struct RowMath;
template <>
struct RowMath<float> {
    static __device__ __forceinline__ float exp_(float x) { return expf(x); }
    static __device__ __forceinline__ float log_(float x) { return logf(x); }
};
template <>
struct RowMath<double> {
    static __device__ __forceinline__ double exp_(double x) { return exp(x); }
    static __device__ __forceinline__ double log_(double x) { return log(x); }
};

// Last-block fold of the per-trajectory (max, sum exp(c - max)) pairs into scal[3] (M), scal[4] (S), scal[6],
// in a fixed order (deterministic).  L2 loads (ld.global.cg: the pairs were written by other CTAs), unrolled so
// that a thread's loads are in flight together -- the volatile loops of the first version were chains of
// ~700-cycle round trips (9 + 18 of them for 2098 trajectories): microseconds of single-CTA tail.
__device__ __forceinline__ void fold_traj_pairs(const double *__restrict__ traj_max, const double *__restrict__ traj_sum,
                                                int n_traj, double *s_warp, double *s_bcast, double *__restrict__ scal) {
    const double NEG_INF = -__longlong_as_double(0x7ff0000000000000LL);
    constexpr int kU = 8;
    const int nthr = blockDim.x;
    double gm = NEG_INF;
    for (int i0 = threadIdx.x; i0 < n_traj; i0 += kU * nthr) {
        double tm[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) tm[u] = (i0 + u * nthr < n_traj) ? __ldcg(traj_max + i0 + u * nthr) : NEG_INF;
#pragma unroll
        for (int u = 0; u < kU; ++u) gm = fmax(gm, tm[u]);
    }
    gm = block_reduce_max(gm, s_warp);
    if (threadIdx.x == 0) *s_bcast = gm;
    __syncthreads();
    gm = *s_bcast;
    double gs = 0.0;
    for (int i0 = threadIdx.x; i0 < n_traj; i0 += kU * nthr) {
        double tm[kU], ts[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const bool ok = i0 + u * nthr < n_traj;
            tm[u] = ok ? __ldcg(traj_max + i0 + u * nthr) : NEG_INF;
            ts[u] = ok ? __ldcg(traj_sum + i0 + u * nthr) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < kU; ++u)
            if (ts[u] != 0.0) gs += ts[u] * exp(tm[u] - gm);   // empty trajectories contribute 0 (NaN != 0 propagates)
    }
    gs = block_reduce_sum(gs, s_warp);
    if (threadIdx.x == 0) {
        scal[3] = gm;
        scal[4] = gs;
        // The reference forms exp(cumsum) directly in float64 (mepol.py:130): it overflows to inf
        // (-> inf/inf = nan weights) above log(DBL_MAX) and the whole batch underflows to 0/0 below
        // log(min subnormal).  The stabilised form here would silently survive; keep the reference's
        // numeric-error behaviour by flagging those regimes.
        scal[6] = (gm > 709.782712893384 || gm < -745.1332191019412) ? 1.0 : 0.0;
    }
}

// ------------------------------------------------------------------------------------ K2a scan
template <typename T>
__global__ void __launch_bounds__(kScanThreads)
    iw_scan_kernel(const T *__restrict__ lt, const T *__restrict__ lb, const int32_t *__restrict__ traj_off,
                   int n_traj, T *__restrict__ c_out, double *__restrict__ traj_max,
                   double *__restrict__ traj_sum, unsigned *__restrict__ counters, double *__restrict__ scal) {
    __shared__ double s_warp[32];
    __shared__ double s_bcast;
    __shared__ unsigned s_ticket;
    const int t = blockIdx.x;
    const int lo = traj_off[t], hi = traj_off[t + 1];
    const double NEG_INF = -__longlong_as_double(0x7ff0000000000000LL);
    double carry = 0.0, local_max = NEG_INF;
    for (int base = lo; base < hi; base += kScanThreads) {
        const int j = base + threadIdx.x;
        double dv = 0.0;
        if (j < hi) dv = (double)lt[j] - (double)lb[j];
        double total;
        double incl = block_inclusive_scan(dv, s_warp, &total) + carry;
        if (j < hi) {
            const T stored = (T)incl;
            c_out[j] = stored;
            // NaN propagates: fmax would drop it, so track it explicitly via the sum below
            local_max = fmax(local_max, (double)stored);
        }
        carry += total;
    }
    double m = block_reduce_max(local_max, s_warp);
    if (threadIdx.x == 0) s_bcast = m;
    __syncthreads();
    m = s_bcast;
    // second pass over the (L1/L2-resident) trajectory: sum exp(c - m_t)
    double part = 0.0;
    for (int j = lo + threadIdx.x; j < hi; j += kScanThreads) part += exp((double)c_out[j] - m);
    part = block_reduce_sum(part, s_warp);
    if (threadIdx.x == 0) {
        traj_max[t] = m;
        traj_sum[t] = (hi > lo) ? part : 0.0;
        __threadfence();
        s_ticket = atomicAdd(counters + 0, 1u);
    }
    __syncthreads();
    if (s_ticket != (unsigned)(n_traj - 1)) return;
    // last CTA: fold per-trajectory (max, sum) pairs in fixed order
    __threadfence();
    fold_traj_pairs(traj_max, traj_sum, n_traj, s_warp, &s_bcast, scal);
}

// ---------------------------------------------------------------------------------- K2b gather
// KT = k when it is known at compile time (4: every shipped configuration but GridWorld), 0 = run-time loop.  With a
// run-time trip count the index and weight loads of a row form a chain; unrolled, a thread has kGatherRows * k
// index loads and then as many gathers in flight.
template <typename T, int KT>
__global__ void __launch_bounds__(kGatherThreads)
    iw_gather_kernel(const T *__restrict__ c, const int32_t *__restrict__ idx, int k1, int n,
                     const T *__restrict__ rk, int ns, double log_G, double B, double eps,
                     T *__restrict__ gW, double *__restrict__ blk_part, unsigned *__restrict__ counters,
                     double *__restrict__ scal, double log_k_over_n, double logV_const) {
    __shared__ double s_warp[32];
    __shared__ unsigned s_ticket;
    const int k = k1 - 1;
    const double M = scal[3], S = scal[4];
    // per-thread constants, amortised over kGatherRows rows
    const double inv_S = 1.0 / S, inv_k = 1.0 / (double)k;
    // log_k_over_n = log k - log n and logV_const = (ns/2) log pi - lgamma(ns/2+1) come from the host: two
    // float64 logs per thread were a quarter of this kernel's instructions
    const T Mt = (T)M;
    double h_term = 0.0, kl_term = 0.0, s_term = 0.0, bad = 0.0;
    // gather phase first for every row of the thread: kGatherRows * k independent loads in flight
    double Wr[kGatherRows];
#pragma unroll
    for (int r = 0; r < kGatherRows; ++r) {
        const int i = (blockIdx.x * kGatherRows + r) * kGatherThreads + threadIdx.x;
        Wr[r] = 0.0;
        if (i < n) {
            const int32_t *ir = idx + (size_t)i * k1;
            double W = 0.0;
            if (KT > 0) {
                int nb[KT > 0 ? KT : 1];
                T cv[KT > 0 ? KT : 1];
#pragma unroll
                for (int m = 0; m < KT; ++m) nb[m] = ir[m];
#pragma unroll
                for (int m = 0; m < KT; ++m) cv[m] = c[nb[m]];
#pragma unroll
                for (int m = 0; m < KT; ++m) W += (double)RowMath<T>::exp_(cv[m] - Mt);   // storage-precision exp, fp64 sum
            } else {
                for (int m = 0; m < k; ++m) W += (double)RowMath<T>::exp_(c[ir[m]] - Mt);
            }
            Wr[r] = W * inv_S;
        }
    }
#pragma unroll
    for (int r = 0; r < kGatherRows; ++r) {
        const int i = (blockIdx.x * kGatherRows + r) * kGatherThreads + threadIdx.x;
        if (i >= n) continue;
        const double W = Wr[r];
        double g, ht, kt;
        if (eps == 0.0) {
            // log-space: log q = log W - log V, log V = ns log r + (ns/2) log pi - lgamma(ns/2+1);
            // r^ns overflows float32 already for ns=29, r>~20 -- never formed explicitly.
            const T logW = RowMath<T>::log_((T)W);
            const T logr = RowMath<T>::log_(rk[i]);
            const double logq = (double)logW - ((double)ns * (double)logr + logV_const);
            ht = (W * inv_k) * logq;
            g = -inv_k * (logq + 1.0);
            kt = log_k_over_n - (double)logW;
        } else {
            // literal float64 evaluation of mepol.py:150-152,165 (MountainCar's eps=1e-15)
            const double rr = (double)rk[i];
            const double V = pow(rr, (double)ns) * exp(0.5 * (double)ns * kLogPi) / exp(log_G);
            const double qv = W / (V + eps);
            ht = (W * inv_k) * log(qv + eps);
            g = -inv_k * (log(qv + eps) + qv / (qv + eps));
            kt = log((double)k / ((double)n * W) + eps);
        }
        if (gW) gW[i] = (T)g;
        h_term += ht;
        kl_term += kt;
        s_term += g * W;
        if (!(ht == ht) || !(kt == kt) || fabs(ht) == INFINITY || fabs(kt) == INFINITY) bad += 1.0;
    }
    const double bh = block_reduce_sum(h_term, s_warp);
    const double bk = block_reduce_sum(kl_term, s_warp);
    const double bs = block_reduce_sum(s_term, s_warp);
    const double bb = block_reduce_sum(bad, s_warp);
    if (threadIdx.x == 0) {
        blk_part[(size_t)blockIdx.x * 4 + 0] = bh;
        blk_part[(size_t)blockIdx.x * 4 + 1] = bk;
        blk_part[(size_t)blockIdx.x * 4 + 2] = bs;
        blk_part[(size_t)blockIdx.x * 4 + 3] = bb;
        __threadfence();
        s_ticket = atomicAdd(counters + 1, 1u);
    }
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    __threadfence();
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    {   // L2 loads, four blocks' partials in flight per thread (see fold_traj_pairs)
        constexpr int kU = 4;
        for (unsigned b0 = threadIdx.x; b0 < gridDim.x; b0 += kU * kGatherThreads) {
            double2 lo2[kU], hi2[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const unsigned b = b0 + u * kGatherThreads;
                const double2 *pp = reinterpret_cast<const double2 *>(blk_part + (size_t)b * 4);
                lo2[u] = b < gridDim.x ? __ldcg(pp) : make_double2(0.0, 0.0);
                hi2[u] = b < gridDim.x ? __ldcg(pp + 1) : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                a0 += lo2[u].x;
                a1 += lo2[u].y;
                a2 += hi2[u].x;
                a3 += hi2[u].y;
            }
        }
    }
    a0 = block_reduce_sum(a0, s_warp);
    a1 = block_reduce_sum(a1, s_warp);
    a2 = block_reduce_sum(a2, s_warp);
    a3 = block_reduce_sum(a3, s_warp);
    if (threadIdx.x == 0) {
        const double poison = scal[6] != 0.0 ? __longlong_as_double(0x7ff8000000000000LL) : 0.0;
        scal[0] = -a0 + B + poison;          // H
        scal[1] = a1 / (double)n + poison;   // KL before the max(0, .) clamp
        scal[2] = a2;                 // s = sum_i gW_i W_i
        scal[5] = a3;                 // rows whose terms were nan/inf
        scal[7] = 1.0;                // unit upstream gradient for the fused forward+backward entry point
    }
}

// ------------------------------------------------------------------------------------- K3 bwd
template <typename T>
__global__ void __launch_bounds__(kScanThreads)
    iw_bwd_kernel(const T *__restrict__ c, const T *__restrict__ gW, const double *__restrict__ scal,
                  const int32_t *__restrict__ traj_off, const int32_t *__restrict__ rowptr,
                  const int32_t *__restrict__ col, const double *__restrict__ grad_H,
                  T *__restrict__ grad_lt) {
    __shared__ double s_warp[32];
    const int t = blockIdx.x;
    const int lo = traj_off[t], hi = traj_off[t + 1];
    const double M = scal[3], S = scal[4], s = scal[2];
    const double gH = grad_H[0];
    const T Mt = (T)M;
    double carry = 0.0;
    // walk the trajectory backwards: position p = 0 is the LAST step
    for (int base = 0; base < hi - lo; base += kScanThreads) {
        const int p = base + threadIdx.x;
        const int j = hi - 1 - p;
        double gc = 0.0;
        if (j >= lo) {
            double gw = 0.0;
            const int e0 = rowptr[j], e1 = rowptr[j + 1];
            for (int e = e0; e < e1; ++e) gw += (double)gW[col[e]];
            const double w = (double)RowMath<T>::exp_(c[j] - Mt) / S;
            gc = w * (gw - s);
        }
        double total;
        const double incl = block_inclusive_scan(gc, s_warp, &total) + carry;
        if (j >= lo) grad_lt[j] = (T)(gH * incl);
        carry += total;
    }
}

// ---------------------------------------------------------------- warp-per-trajectory variants
// With thousands of short trajectories (the N=1M benchmark has 2098 of length 500) a CTA per
// trajectory spends its time in __syncthreads; here one warp scans one trajectory with shuffles only.
// Each lane owns kLaneChunk CONSECUTIVE steps (16-byte vector loads / stores): the prefix inside a
// chunk is a register-local chain, ONE warp scan of the 32 chunk totals per 512 steps links them
// (the first version ran a 5-stage shuffle scan per 32 steps: sixteen dependent scans per trajectory
// with 14 warps per SM to hide them, 22 us for 8 MB), and the chunk's (max, sum exp) is formed from
// registers with independent exps.  The trajectory is read exactly once.
constexpr int kWarpScanWarps = 8;
constexpr int kLaneChunk = 16;

__device__ __forceinline__ double warp_inclusive_scan(double v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// out[u] = p[j0 + u] for j0 + u in [lo_valid, hi_valid), 0 elsewhere; 16-byte vectors when the whole chunk is
// inside the range and aligned (trajectory offsets are multiples of 4 in every shipped configuration)
template <typename T>
__device__ __forceinline__ void load_chunk(const T *__restrict__ p, int j0, int lo_valid, int hi_valid, T (&out)[kLaneChunk]) {
    constexpr int V = 16 / (int)sizeof(T);
    if (j0 >= lo_valid && j0 + kLaneChunk <= hi_valid && (reinterpret_cast<uintptr_t>(p + j0) & 15) == 0) {
        const uint4 *q = reinterpret_cast<const uint4 *>(p + j0);
#pragma unroll
        for (int v = 0; v < kLaneChunk / V; ++v) {
            const uint4 r = q[v];
            memcpy(&out[v * V], &r, 16);
        }
    } else {
#pragma unroll
        for (int u = 0; u < kLaneChunk; ++u) {
            const int jj = j0 + u;
            out[u] = (jj >= lo_valid && jj < hi_valid) ? p[jj] : (T)0;
        }
    }
}
template <typename T>
__device__ __forceinline__ void store_chunk(T *__restrict__ p, int j0, int lo_valid, int hi_valid, const T (&in)[kLaneChunk]) {
    constexpr int V = 16 / (int)sizeof(T);
    if (j0 >= lo_valid && j0 + kLaneChunk <= hi_valid && (reinterpret_cast<uintptr_t>(p + j0) & 15) == 0) {
        uint4 *q = reinterpret_cast<uint4 *>(p + j0);
#pragma unroll
        for (int v = 0; v < kLaneChunk / V; ++v) {
            uint4 r;
            memcpy(&r, &in[v * V], 16);
            q[v] = r;
        }
    } else {
#pragma unroll
        for (int u = 0; u < kLaneChunk; ++u) {
            const int jj = j0 + u;
            if (jj >= lo_valid && jj < hi_valid) p[jj] = in[u];
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kWarpScanWarps * 32)
    iw_scan_warp_kernel(const T *__restrict__ lt, const T *__restrict__ lb, const int32_t *__restrict__ traj_off,
                        int n_traj, T *__restrict__ c_out, double *__restrict__ traj_max,
                        double *__restrict__ traj_sum, unsigned *__restrict__ counters, double *__restrict__ scal) {
    __shared__ double s_warp[32];
    __shared__ double s_bcast;
    __shared__ unsigned s_ticket;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = blockIdx.x * kWarpScanWarps + warp;
    const double NEG_INF = -__longlong_as_double(0x7ff0000000000000LL);
    if (t < n_traj) {
        const int lo = traj_off[t], hi = traj_off[t + 1];
        double carry = 0.0, m = NEG_INF, ssum = 0.0;
        for (int base = lo; base < hi; base += 32 * kLaneChunk) {
            const int j0 = base + lane * kLaneChunk;
            T a[kLaneChunk], b[kLaneChunk];
            load_chunk(lt, j0, lo, hi, a);
            load_chunk(lb, j0, lo, hi, b);
            double pre[kLaneChunk];
            pre[0] = (double)a[0] - (double)b[0];
#pragma unroll
            for (int u = 1; u < kLaneChunk; ++u) pre[u] = pre[u - 1] + ((double)a[u] - (double)b[u]);
            const double incl = warp_inclusive_scan(pre[kLaneChunk - 1], lane);
            const double off = incl - pre[kLaneChunk - 1] + carry;
            carry += __shfl_sync(0xffffffffu, incl, 31);
            T stored[kLaneChunk];
            double lm = NEG_INF;
#pragma unroll
            for (int u = 0; u < kLaneChunk; ++u) {
                stored[u] = (T)(pre[u] + off);
                if (j0 + u < hi) lm = fmax(lm, (double)stored[u]);    // NaN is ignored here and poisons the sum below
            }
            store_chunk(c_out, j0, lo, hi, stored);
            if (lm > m) {               // rescale the running (max, sum exp(c - max)) pair
                ssum *= exp(m - lm);
                m = lm;
            }
            // storage-precision exp like the gather kernel's weights (a float64 exp per step was half of this kernel in
            // the fp32 build: 13.4 -> see profiles/); the difference c - max is formed in float64 first
#pragma unroll
            for (int u = 0; u < kLaneChunk; ++u)
                if (j0 + u < hi) ssum += (double)RowMath<T>::exp_((T)((double)stored[u] - m));
        }
        // fold the 32 (max, sum) pairs
        const double wm = warp_max(m);
        // a lane without steps contributes 0; a lane that only saw NaN (m still -inf, sum NaN) keeps its NaN
        const double part = warp_sum(m == NEG_INF ? (ssum != ssum ? ssum : 0.0) : ssum * exp(m - wm));
        if (lane == 0) {
            traj_max[t] = wm;
            traj_sum[t] = (hi > lo) ? part : 0.0;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_ticket = atomicAdd(counters + 0, 1u);
    }
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    // last CTA: fold per-trajectory (max, sum) pairs in fixed order
    __threadfence();
    fold_traj_pairs(traj_max, traj_sum, n_traj, s_warp, &s_bcast, scal);
}

// Backward in the many-trajectories regime, two launches:
//   iw_bwd_elem    one thread per particle (full occupancy for the latency-bound CSR gathers):
//                  g_c[j] = w_j (sum_{e in rev[j]} gW[col[e]] - s), written into the output buffer
//   iw_bwd_rscan   one warp per trajectory: reverse inclusive scan of g_c in place, times grad_H
template <typename T>
__global__ void __launch_bounds__(256)
    iw_bwd_elem_kernel(const T *__restrict__ c, const T *__restrict__ gW, const double *__restrict__ scal, int n,
                       const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col, T *__restrict__ gc_out) {
    // (Two particles per thread with the rowptr / col / gW loads of a level issued together measured SLOWER, 20.1 vs
    // 18.6 us at N = 1M: the kernel is paced by the 4 M random 4-byte gathers, a 32-byte L2 sector each, not by the
    // dependent round trips.)
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n) return;
    const double M = scal[3], S = scal[4], s = scal[2];
    double gw = 0.0;
    const int e0 = rowptr[j], e1 = rowptr[j + 1];
    for (int e = e0; e < e1; ++e) gw += (double)gW[col[e]];
    const double w = (double)RowMath<T>::exp_(c[j] - (T)M) / S;
    gc_out[j] = (T)(w * (gw - s));
}

template <typename T>
__global__ void __launch_bounds__(kWarpScanWarps * 32)
    iw_bwd_rscan_kernel(const int32_t *__restrict__ traj_off, int n_traj, const double *__restrict__ grad_H,
                        T *__restrict__ g) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = blockIdx.x * kWarpScanWarps + warp;
    if (t >= n_traj) return;
    const int lo = traj_off[t], hi = traj_off[t + 1];
    const double gH = grad_H[0];
    double carry = 0.0;
    // walk backwards in blocks of 512 steps; lane 0 owns the LAST kLaneChunk steps of a block
    for (int top = hi; top > lo; top -= 32 * kLaneChunk) {
        const int j0 = top - (lane + 1) * kLaneChunk;       // the lane's chunk is [j0, j0 + kLaneChunk), clipped at lo
        T v[kLaneChunk];
        load_chunk(g, j0, lo, hi, v);
        double suf[kLaneChunk];                              // suffix sums inside the chunk
        suf[kLaneChunk - 1] = (double)v[kLaneChunk - 1];
#pragma unroll
        for (int u = kLaneChunk - 2; u >= 0; --u) suf[u] = suf[u + 1] + (double)v[u];
        const double incl = warp_inclusive_scan(suf[0], lane);   // lanes in walking order: lane 0 first
        const double off = incl - suf[0] + carry;
        carry += __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
        for (int u = 0; u < kLaneChunk; ++u) v[u] = (T)(gH * (suf[u] + off));
        store_chunk(g, j0, lo, hi, v);
    }
}

// -------------------------------------------------------------------------------- K4 reverse CSR
constexpr int kCsrThreads = 256;

__global__ void csr_count_kernel(const int32_t *__restrict__ idx, int n, int k1, int32_t *__restrict__ deg) {
    const int k = k1 - 1;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)n * k) return;
    const int i = (int)(e / k), m = (int)(e % k);
    atomicAdd(deg + idx[(size_t)i * k1 + m], 1);
}

// exclusive scan of deg[0..n) -> rowptr[0..n], three small kernels (block sums, scan of sums, add)
__global__ void __launch_bounds__(kCsrThreads)
    csr_block_sum_kernel(const int32_t *__restrict__ deg, int n, int items, int32_t *__restrict__ bsum) {
    __shared__ int s[kCsrThreads / 32];
    const int base = blockIdx.x * kCsrThreads * items;
    int v = 0;
    for (int it = 0; it < items; ++it) {
        const int j = base + it * kCsrThreads + threadIdx.x;
        if (j < n) v += deg[j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < kCsrThreads / 32; ++w) t += s[w];
        bsum[blockIdx.x] = t;
    }
}
__global__ void csr_scan_sums_kernel(int32_t *__restrict__ bsum, int nb) {
    // single thread block, serial over a few thousand entries at most
    if (threadIdx.x == 0) {
        int run = 0;
        for (int b = 0; b < nb; ++b) {
            const int v = bsum[b];
            bsum[b] = run;
            run += v;
        }
    }
}
__global__ void __launch_bounds__(kCsrThreads)
    csr_finish_kernel(const int32_t *__restrict__ deg, int n, int items, const int32_t *__restrict__ bsum,
                      int32_t *__restrict__ rowptr, int32_t *__restrict__ cursor) {
    __shared__ int s_w[kCsrThreads / 32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = bsum[blockIdx.x];
    __syncthreads();
    const int base = blockIdx.x * kCsrThreads * items;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int it = 0; it < items; ++it) {
        const int j = base + it * kCsrThreads + threadIdx.x;
        const int v = j < n ? deg[j] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        int prev = 0;
        for (int w = 0; w < warp; ++w) prev += s_w[w];
        int tot = 0;
        for (int w = 0; w < kCsrThreads / 32; ++w) tot += s_w[w];
        const int excl = s_carry + prev + incl - v;
        if (j < n) {
            rowptr[j] = excl;
            cursor[j] = excl;
        }
        if (j == n - 1) rowptr[n] = excl + v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += tot;
        __syncthreads();
    }
}
__global__ void csr_fill_kernel(const int32_t *__restrict__ idx, int n, int k1, int32_t *__restrict__ cursor,
                                int32_t *__restrict__ col) {
    const int k = k1 - 1;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)n * k) return;
    const int i = (int)(e / k), m = (int)(e % k);
    const int pos = atomicAdd(cursor + idx[(size_t)i * k1 + m], 1);
    col[pos] = i;
}
// Per-row sort: atomics filled each row in arbitrary order; sorting makes the backward's summation
// order (and therefore its bits) reproducible run to run.  The in-degree of a particle is NOT bounded
// by k: with exact duplicates (MountainCar's wall states) every copy lists the same k lowest-index
// copies, so a few hub rows collect thousands of entries.  Rows of up to kCsrShortRow entries are
// insertion-sorted by one thread (<= ~16k steps); longer ones are queued for csr_sort_long_rows_kernel.
constexpr int kCsrShortRow = 256;   // k = 50 (GridWorld) makes ~50-100-entry rows the norm: still one thread's job
__global__ void csr_sort_rows_kernel(const int32_t *__restrict__ rowptr, int n, int32_t *__restrict__ col,
                                     int32_t *__restrict__ long_list, int32_t *__restrict__ long_count) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int e0 = rowptr[j], e1 = rowptr[j + 1];
    if (e1 - e0 > kCsrShortRow) {
        long_list[atomicAdd(long_count, 1)] = j;   // list order is arbitrary; every row is sorted on its own
        return;
    }
    for (int a = e0 + 1; a < e1; ++a) {
        const int v = col[a];
        int b = a - 1;
        while (b >= e0 && col[b] > v) {
            col[b + 1] = col[b];
            --b;
        }
        col[b + 1] = v;
    }
}

// Long rows: one CTA per row, LSD radix sort (4-bit digits) between the row and a scratch copy.  Every
// thread owns a contiguous slice of the row and a private column of the 16 x 256 histogram, so the
// scatter is stable without atomics; O(deg * passes / 256) per row whatever the degree.
constexpr int kCsrLongThreads = 256;
__global__ void __launch_bounds__(kCsrLongThreads)
    csr_sort_long_rows_kernel(const int32_t *__restrict__ rowptr, int32_t *__restrict__ col, int32_t *__restrict__ tmp,
                              const int32_t *__restrict__ long_list, const int32_t *__restrict__ long_count, int passes) {
    __shared__ int hist[16 * kCsrLongThreads];
    __shared__ int warp_tot[kCsrLongThreads / 32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int n_long = *long_count;
    for (int r = blockIdx.x; r < n_long; r += gridDim.x) {
        const int j = long_list[r];
        const int e0 = rowptr[j], deg = rowptr[j + 1] - e0;
        const int per = (deg + kCsrLongThreads - 1) / kCsrLongThreads;
        const int lo = min(deg, t * per), hi = min(deg, lo + per);
        int32_t *src = col + e0, *dst = tmp + e0;
        for (int p = 0; p < passes; ++p) {
            const int shift = 4 * p;
#pragma unroll
            for (int dg = 0; dg < 16; ++dg) hist[dg * kCsrLongThreads + t] = 0;
            for (int e = lo; e < hi; ++e) ++hist[((src[e] >> shift) & 15) * kCsrLongThreads + t];
            __syncthreads();
            // exclusive scan of the 4096 counters in (digit, thread) order: thread t owns entries 16t .. 16t+15
            int loc[16], sum = 0;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                loc[q] = sum;
                sum += hist[16 * t + q];
            }
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            if (lane == 31) warp_tot[warp] = incl;
            __syncthreads();
            int base = incl - sum;
            for (int w = 0; w < warp; ++w) base += warp_tot[w];
#pragma unroll
            for (int q = 0; q < 16; ++q) hist[16 * t + q] = base + loc[q];
            __syncthreads();
            for (int e = lo; e < hi; ++e) {
                const int v = src[e];
                dst[hist[((v >> shift) & 15) * kCsrLongThreads + t]++] = v;
            }
            __syncthreads();
            int32_t *sw = src;
            src = dst;
            dst = sw;
        }
    }
}

constexpr int kWarpPathMinTraj = 512;   // below this a CTA per trajectory has more parallelism
constexpr int kCsrItems = 8;
static int csr_blocks_for(int64_t n) { return (int)((n + (int64_t)kCsrThreads * kCsrItems - 1) / ((int64_t)kCsrThreads * kCsrItems)); }

}  // namespace mepol

using namespace mepol;

extern "C" int mepol_rev_csr_workspace_bytes(int64_t n, int k, size_t *bytes) {
    MEPOL_REQUIRE(bytes && n > 0 && k > 0, MEPOL_E_INVALID, "rev_csr_workspace_bytes: bad arguments");
    Carver c(nullptr);
    c.take<int32_t>((size_t)n + 1);               // deg
    c.take<int32_t>((size_t)n + 1);               // cursor
    c.take<int32_t>((size_t)csr_blocks_for(n) + 1);
    c.take<int32_t>((size_t)n * k / kCsrShortRow + 2);   // rows queued for the long-row sort (+ their count)
    c.take<int32_t>((size_t)n * k);                      // scratch copy of col for the long-row sort
    *bytes = c.bytes();
    return MEPOL_OK;
}

extern "C" int mepol_rev_csr(const int32_t *idx, int64_t n, int k1, int32_t *rowptr, int32_t *col,
                             void *workspace, size_t workspace_bytes, void *stream) {
    MEPOL_REQUIRE(idx && rowptr && col && workspace, MEPOL_E_INVALID, "rev_csr: null pointer");
    MEPOL_REQUIRE(n > 0 && k1 >= 2 && n * (int64_t)(k1 - 1) < (1LL << 31), MEPOL_E_INVALID,
                  "rev_csr: need n > 0, k1 >= 2 and n*k < 2^31 (n=%lld k1=%d)", (long long)n, k1);
    size_t need;
    mepol_rev_csr_workspace_bytes(n, k1 - 1, &need);
    MEPOL_REQUIRE(workspace_bytes >= need, MEPOL_E_WORKSPACE, "rev_csr: workspace %zu < %zu", workspace_bytes, need);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Carver c(workspace);
    int32_t *deg = c.take<int32_t>((size_t)n + 1);
    int32_t *cursor = c.take<int32_t>((size_t)n + 1);
    const int nb = csr_blocks_for(n);
    int32_t *bsum = c.take<int32_t>((size_t)nb + 1);
    const int k = k1 - 1;
    int32_t *long_list = c.take<int32_t>((size_t)n * k / kCsrShortRow + 2);
    int32_t *long_count = long_list + (size_t)n * k / kCsrShortRow + 1;
    int32_t *tmp = c.take<int32_t>((size_t)n * k);
    MEPOL_CUDA_CHECK(cudaMemsetAsync(long_count, 0, sizeof(int32_t), st));
    MEPOL_CUDA_CHECK(cudaMemsetAsync(deg, 0, sizeof(int32_t) * ((size_t)n + 1), st));
    const int64_t edges = n * (int64_t)k;
    const int eb = (int)((edges + 255) / 256);
    csr_count_kernel<<<eb, 256, 0, st>>>(idx, (int)n, k1, deg);
    csr_block_sum_kernel<<<nb, kCsrThreads, 0, st>>>(deg, (int)n, kCsrItems, bsum);
    csr_scan_sums_kernel<<<1, 32, 0, st>>>(bsum, nb);
    csr_finish_kernel<<<nb, kCsrThreads, 0, st>>>(deg, (int)n, kCsrItems, bsum, rowptr, cursor);
    csr_fill_kernel<<<eb, 256, 0, st>>>(idx, (int)n, k1, cursor, col);
    csr_sort_rows_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(rowptr, (int)n, col, long_list, long_count);
    {
        // key = source row index < n; an even number of 4-bit passes leaves the sorted row in col
        int bits = 1;
        while ((1LL << bits) < n) ++bits;
        int passes = (bits + 3) / 4;
        passes += passes & 1;
        csr_sort_long_rows_kernel<<<device_sm_count(), kCsrLongThreads, 0, st>>>(rowptr, col, tmp, long_list, long_count,
                                                                                passes);
    }
    MEPOL_CUDA_CHECK(cudaGetLastError());
    return MEPOL_OK;
}

extern "C" int mepol_iw_workspace_bytes(int64_t n, int64_t n_traj, size_t *bytes) {
    MEPOL_REQUIRE(bytes && n > 0 && n_traj > 0, MEPOL_E_INVALID, "iw_workspace_bytes: bad arguments");
    IwWorkspace w;
    carve_iw(w, n, n_traj, nullptr);
    *bytes = w.bytes;
    return MEPOL_OK;
}

template <typename T>
static int iw_fwd_impl(const T *lt, const T *lb, const int32_t *traj_off, int64_t n_traj, int64_t n,
                       const int32_t *idx, int k1, const T *rk, int ns, double log_G, double B, double eps,
                       T *c, T *gW, double *scal, void *workspace, size_t workspace_bytes, void *stream) {
    MEPOL_REQUIRE(lt && lb && traj_off && idx && rk && c && scal && workspace, MEPOL_E_INVALID,
                  "iw_entropy_fwd: null pointer");
    MEPOL_REQUIRE(n > 0 && n < (1LL << 31) && n_traj > 0 && n_traj <= n && k1 >= 2 && ns > 0, MEPOL_E_INVALID,
                  "iw_entropy_fwd: bad sizes n=%lld n_traj=%lld k1=%d ns=%d", (long long)n, (long long)n_traj, k1, ns);
    IwWorkspace w;
    carve_iw(w, n, n_traj, workspace);
    MEPOL_REQUIRE(workspace_bytes >= w.bytes, MEPOL_E_WORKSPACE, "iw_entropy_fwd: workspace %zu < %zu",
                  workspace_bytes, w.bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MEPOL_CUDA_CHECK(cudaMemsetAsync(w.counters, 0, 4 * sizeof(unsigned), st));
    if (n_traj >= kWarpPathMinTraj)
        iw_scan_warp_kernel<T><<<(int)((n_traj + kWarpScanWarps - 1) / kWarpScanWarps), kWarpScanWarps * 32, 0, st>>>(
            lt, lb, traj_off, (int)n_traj, c, w.traj_max, w.traj_sum, w.counters, scal);
    else
        iw_scan_kernel<T><<<(int)n_traj, kScanThreads, 0, st>>>(lt, lb, traj_off, (int)n_traj, c, w.traj_max,
                                                                w.traj_sum, w.counters, scal);
    auto gather = k1 == 5 ? iw_gather_kernel<T, 4> : iw_gather_kernel<T, 0>;
    gather<<<gather_blocks_for(n), kGatherThreads, 0, st>>>(
        c, idx, k1, (int)n, rk, ns, log_G, B, eps, gW, w.blk_part, w.counters, scal,
        log((double)(k1 - 1)) - log((double)n), 0.5 * (double)ns * kLogPi - log_G);
    MEPOL_CUDA_CHECK(cudaGetLastError());
    return MEPOL_OK;
}

template <typename T>
static int iw_bwd_impl(const T *c, const T *gW, const double *scal, const int32_t *traj_off, int64_t n_traj,
                       int64_t n, const int32_t *rowptr, const int32_t *col, const double *grad_H, T *grad_lt,
                       void *stream) {
    MEPOL_REQUIRE(c && gW && scal && traj_off && rowptr && col && grad_H && grad_lt, MEPOL_E_INVALID,
                  "iw_entropy_bwd: null pointer");
    MEPOL_REQUIRE(n > 0 && n_traj > 0, MEPOL_E_INVALID, "iw_entropy_bwd: bad sizes");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_traj >= kWarpPathMinTraj)
    {
        iw_bwd_elem_kernel<T><<<(int)((n + 255) / 256), 256, 0, st>>>(c, gW, scal, (int)n, rowptr, col, grad_lt);
        iw_bwd_rscan_kernel<T><<<(int)((n_traj + kWarpScanWarps - 1) / kWarpScanWarps), kWarpScanWarps * 32, 0, st>>>(
            traj_off, (int)n_traj, grad_H, grad_lt);
    }
    else
        iw_bwd_kernel<T><<<(int)n_traj, kScanThreads, 0, st>>>(c, gW, scal, traj_off, rowptr, col, grad_H, grad_lt);
    MEPOL_CUDA_CHECK(cudaGetLastError());
    return MEPOL_OK;
}

extern "C" int mepol_iw_entropy_fwd_f32(const float *logp_t, const float *logp_b, const int32_t *traj_off,
                                        int64_t n_traj, int64_t n, const int32_t *idx, int k1, const float *rk,
                                        int ns, double log_G, double B, double eps, float *c, float *gW,
                                        double *scal, void *workspace, size_t workspace_bytes, void *stream) {
    return iw_fwd_impl<float>(logp_t, logp_b, traj_off, n_traj, n, idx, k1, rk, ns, log_G, B, eps, c, gW, scal,
                              workspace, workspace_bytes, stream);
}
extern "C" int mepol_iw_entropy_fwd_f64(const double *logp_t, const double *logp_b, const int32_t *traj_off,
                                        int64_t n_traj, int64_t n, const int32_t *idx, int k1, const double *rk,
                                        int ns, double log_G, double B, double eps, double *c, double *gW,
                                        double *scal, void *workspace, size_t workspace_bytes, void *stream) {
    return iw_fwd_impl<double>(logp_t, logp_b, traj_off, n_traj, n, idx, k1, rk, ns, log_G, B, eps, c, gW, scal,
                               workspace, workspace_bytes, stream);
}
// Forward and backward in one call, the backward for dH = 1 (scal[7] holds the 1.0 it reads as grad_H):
// what an autograd bridge wants when it knows a gradient will be asked for -- the caller scales
// grad_unit by the incoming dH afterwards.  Same kernels, same results as the two separate calls.
template <typename T>
static int iw_fwd_bwd_impl(const T *lt, const T *lb, const int32_t *traj_off, int64_t n_traj, int64_t n,
                           const int32_t *idx, int k1, const T *rk, int ns, double log_G, double B, double eps,
                           const int32_t *rowptr, const int32_t *col, T *c, T *gW, double *scal, T *grad_unit,
                           void *workspace, size_t workspace_bytes, void *stream) {
    MEPOL_REQUIRE(gW && grad_unit && rowptr && col, MEPOL_E_INVALID, "iw_entropy_fwd_bwd: null pointer");
    int rc = iw_fwd_impl<T>(lt, lb, traj_off, n_traj, n, idx, k1, rk, ns, log_G, B, eps, c, gW, scal, workspace,
                            workspace_bytes, stream);
    if (rc != MEPOL_OK) return rc;
    return iw_bwd_impl<T>(c, gW, scal, traj_off, n_traj, n, rowptr, col, scal + 7, grad_unit, stream);
}
extern "C" int mepol_iw_entropy_fwd_bwd_f32(const float *logp_t, const float *logp_b, const int32_t *traj_off,
                                            int64_t n_traj, int64_t n, const int32_t *idx, int k1, const float *rk,
                                            int ns, double log_G, double B, double eps, const int32_t *rowptr,
                                            const int32_t *col, float *c, float *gW, double *scal, float *grad_unit,
                                            void *workspace, size_t workspace_bytes, void *stream) {
    return iw_fwd_bwd_impl<float>(logp_t, logp_b, traj_off, n_traj, n, idx, k1, rk, ns, log_G, B, eps, rowptr, col, c,
                                  gW, scal, grad_unit, workspace, workspace_bytes, stream);
}
extern "C" int mepol_iw_entropy_fwd_bwd_f64(const double *logp_t, const double *logp_b, const int32_t *traj_off,
                                            int64_t n_traj, int64_t n, const int32_t *idx, int k1, const double *rk,
                                            int ns, double log_G, double B, double eps, const int32_t *rowptr,
                                            const int32_t *col, double *c, double *gW, double *scal,
                                            double *grad_unit, void *workspace, size_t workspace_bytes, void *stream) {
    return iw_fwd_bwd_impl<double>(logp_t, logp_b, traj_off, n_traj, n, idx, k1, rk, ns, log_G, B, eps, rowptr, col, c,
                                   gW, scal, grad_unit, workspace, workspace_bytes, stream);
}
extern "C" int mepol_iw_entropy_bwd_f32(const float *c, const float *gW, const double *scal,
                                        const int32_t *traj_off, int64_t n_traj, int64_t n,
                                        const int32_t *rowptr, const int32_t *col, const double *grad_H,
                                        float *grad_logp_t, void *stream) {
    return iw_bwd_impl<float>(c, gW, scal, traj_off, n_traj, n, rowptr, col, grad_H, grad_logp_t, stream);
}
extern "C" int mepol_iw_entropy_bwd_f64(const double *c, const double *gW, const double *scal,
                                        const int32_t *traj_off, int64_t n_traj, int64_t n,
                                        const int32_t *rowptr, const int32_t *col, const double *grad_H,
                                        double *grad_logp_t, void *stream) {
    return iw_bwd_impl<double>(c, gW, scal, traj_off, n_traj, n, rowptr, col, grad_H, grad_logp_t, stream);
}
