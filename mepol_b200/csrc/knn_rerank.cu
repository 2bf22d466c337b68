// Exact float64 re-rank of the pruned candidates, and the exact brute-force fallback for rows whose
// candidate buffer overflowed.
//
// This is where the reference's ordering semantics are reproduced: sklearn ranks on float64
// squared distances of the float32 inputs (_argkmin.pyx.tp:492-498) and keeps the first-scanned
// (lowest-index) element among exact ties (sklearn/utils/_heap.pyx:46), then sorts each row
// ascending (_argkmin.pyx.tp:193-198).  Here: d2 = sum_c ((double)x_c - (double)y_c)^2 (every
// difference exact, the sum correct to ~d*1e-16), order = ascending (d2, index).
#include "knn_internal.cuh"

namespace mepol {

constexpr int kRerankWarps = 4;

__device__ __forceinline__ bool key_less(double da, int ia, double db, int ib) {
    return da < db || (da == db && ia < ib);
}

__device__ __forceinline__ double exact_d2(const float *__restrict__ a, const float *__restrict__ b, int d) {
    double s = 0.0;
    for (int c = 0; c < d; ++c) {
        double t = (double)a[c] - (double)b[c];
        s = fma(t, t, s);
    }
    return s;
}

// Eight lanes per query row (four rows per warp): a row carries ~6 candidates after pruning, so a whole
// warp per row left 26 lanes idle behind six serial 29-term fma chains (786 us at N = 1M in round 1).
// Candidates of all splits are gathered into shared memory, their exact distances computed one candidate
// per lane (the same sequential fma chain as every other path: distances are bit-identical across
// kernels), then each candidate's rank is counted; ranks < k1 are output.
constexpr int kRerankSub = 8;                               // lanes per row
constexpr int kRerankRowsPerWarp = 32 / kRerankSub;
__global__ void __launch_bounds__(kRerankWarps * 32)
    knn_rerank_kernel(const float *__restrict__ ref, const float *__restrict__ q, int n_q, int d, int k1,
                      int cap, int n_split, const int32_t *__restrict__ cand,
                      const int32_t *__restrict__ cand_cnt, unsigned char *__restrict__ flag,
                      int32_t *__restrict__ idx_out, double *__restrict__ dist_out,
                      double *__restrict__ rk_out, int max_cand) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane / kRerankSub, sl = lane % kRerankSub;
    const int slot_in_block = warp * kRerankRowsPerWarp + sub;
    constexpr int kRowsPerBlock = kRerankWarps * kRerankRowsPerWarp;
    double *sd = reinterpret_cast<double *>(smem_raw) + (size_t)slot_in_block * max_cand;
    int *si = reinterpret_cast<int *>(reinterpret_cast<double *>(smem_raw) + (size_t)kRowsPerBlock * max_cand) +
              (size_t)slot_in_block * max_cand;
    const int row = blockIdx.x * kRowsPerBlock + slot_in_block;
    const bool active = row < n_q && !flag[row];            // flagged rows belong to the exact fallback
    // gather candidate indices (splits are disjoint reference ranges -> no duplicates)
    int total = 0;
    if (active) {
        for (int s = 0; s < n_split; ++s) {
            const size_t slot = (size_t)row * n_split + s;
            const int c = cand_cnt[slot];
            for (int e = sl; e < c; e += kRerankSub) si[total + e] = cand[slot * cap + e];
            total += c;
        }
        if (total < k1) {  // cannot happen for finite inputs; stay safe
            if (sl == 0) flag[row] = 1;
            total = 0;
        }
    }
    __syncwarp();
    if (total > 0) {
        const float *qr = q + (size_t)row * d;
        for (int e = sl; e < total; e += kRerankSub) sd[e] = exact_d2(qr, ref + (size_t)si[e] * d, d);
    }
    __syncwarp();
    for (int e = sl; e < total; e += kRerankSub) {
        const double de = sd[e];
        const int ie = si[e];
        int rank = 0;
        for (int o = 0; o < total; ++o) rank += key_less(sd[o], si[o], de, ie) ? 1 : 0;
        if (rank < k1) {
            idx_out[(size_t)row * k1 + rank] = ie;
            const double r = sqrt(de);
            dist_out[(size_t)row * k1 + rank] = r;
            if (rk_out && rank == k1 - 1) rk_out[row] = r;
        }
    }
}

int launch_rerank(const float *ref, const float *q, const KnnPlan &p, int32_t *idx_out, double *dist_out,
                  double *rk_out, cudaStream_t st) {
    const int max_cand = p.cap * p.n_split;
    const int rows_per_block = kRerankWarps * kRerankRowsPerWarp;
    const size_t smem = (size_t)rows_per_block * max_cand * (sizeof(double) + sizeof(int));
    MEPOL_CUDA_CHECK(cudaFuncSetAttribute(knn_rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
    const int blocks = (int)((p.n_q + rows_per_block - 1) / rows_per_block);
    knn_rerank_kernel<<<blocks, kRerankWarps * 32, smem, st>>>(ref, q, (int)p.n_q, p.d, p.k1, p.cap,
                                                               p.n_split, p.cand, p.cand_cnt, p.flag,
                                                               idx_out, dist_out, rk_out, max_cand);
    MEPOL_CUDA_CHECK(cudaGetLastError());
    return MEPOL_OK;
}

// ------------------------------------------------------------------------------------------
// Exact fallback: for every flagged row, k1 rounds of "smallest (d2, index) key strictly greater
// than the previous one" over ALL references in float64.  O(k1 * n_ref * d) per row -- only rows
// whose band held more than `cap` candidates (massive duplication) come here.
// ------------------------------------------------------------------------------------------
constexpr int kFallbackThreads = 256;

__global__ void __launch_bounds__(kFallbackThreads)
    knn_exact_fallback_kernel(const float *__restrict__ ref, int n_ref, const float *__restrict__ q, int n_q,
                              int d, int k1, const unsigned char *__restrict__ flag,
                              int32_t *__restrict__ idx_out, double *__restrict__ dist_out,
                              double *__restrict__ rk_out, unsigned long long *__restrict__ stats) {
    __shared__ double s_d[kFallbackThreads / 32];
    __shared__ int s_i[kFallbackThreads / 32];
    __shared__ double s_last_d;
    __shared__ int s_last_i;
    __shared__ float s_q[MEPOL_MAX_DIM];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    __shared__ int s_rows[kFallbackThreads];
    __shared__ int s_nrows;
    const int n_chunks = (n_q + kFallbackThreads - 1) / kFallbackThreads;
    for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const int my_row = chunk * kFallbackThreads + tid;
        const bool mine = my_row < n_q && flag[my_row] == 1;   // 2 = taken by the parallel path
        if (tid == 0) s_nrows = 0;
        if (!__syncthreads_or(mine ? 1 : 0)) continue;  // the common case: nothing flagged here
        if (mine) s_rows[atomicAdd(&s_nrows, 1)] = my_row;
        __syncthreads();
        const int n_rows = s_nrows;
        for (int t = 0; t < n_rows; ++t) {
            const int row = s_rows[t];
            __syncthreads();
            if (tid < d) s_q[tid] = q[(size_t)row * d + tid];
            if (tid == 0) {
                s_last_d = -1.0;
                s_last_i = -1;
                if (stats) atomicAdd(stats + 1, 1ULL);
            }
            __syncthreads();
            for (int m = 0; m < k1; ++m) {
                const double ld = s_last_d;
                const int li = s_last_i;
                double bd = INF;
                int bi = 0x7fffffff;
                for (int j = tid; j < n_ref; j += kFallbackThreads) {
                    const double dj = exact_d2(s_q, ref + (size_t)j * d, d);
                    if (key_less(ld, li, dj, j) && key_less(dj, j, bd, bi)) {
                        bd = dj;
                        bi = j;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double od = __shfl_xor_sync(0xffffffffu, bd, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (key_less(od, oi, bd, bi)) {
                        bd = od;
                        bi = oi;
                    }
                }
                if (lane == 0) {
                    s_d[warp] = bd;
                    s_i[warp] = bi;
                }
                __syncthreads();
                if (tid == 0) {
                    for (int w = 1; w < kFallbackThreads / 32; ++w)
                        if (key_less(s_d[w], s_i[w], bd, bi)) {
                            bd = s_d[w];
                            bi = s_i[w];
                        }
                    const bool ok = bi != 0x7fffffff;
                    idx_out[(size_t)row * k1 + m] = ok ? bi : -1;
                    const double r = ok ? sqrt(bd) : INF;
                    dist_out[(size_t)row * k1 + m] = r;
                    if (rk_out && m == k1 - 1) rk_out[row] = r;
                    s_last_d = bd;
                    s_last_i = bi;
                }
                __syncthreads();
            }
        }
        __syncthreads();
    }
}


// ------------------------------------------------------------------------------------------
// Parallel exact fallback for the first kFbMaxRows flagged rows: the reference range is cut into
// chunks, one CTA computes the exact float64 top-k1 of (row, chunk), a warp merges the chunks.
// A handful of overflowing rows at N=1M would otherwise serialise 1M x d fp64 distances on one SM.
// ------------------------------------------------------------------------------------------
__global__ void fallback_list_kernel(unsigned char *__restrict__ flag, int n_q, int32_t *__restrict__ list,
                                     int *__restrict__ count, int max_rows) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_q || flag[row] != 1) return;
    const int slot = atomicAdd(count, 1);
    if (slot < max_rows) {
        list[slot] = row;
        flag[row] = 2;
    }
}

__global__ void __launch_bounds__(kFallbackThreads)
    fallback_partial_kernel(const float *__restrict__ ref, int n_ref, const float *__restrict__ q, int d, int k1,
                            const int32_t *__restrict__ list, const int *__restrict__ count, int max_rows,
                            int n_chunks, double *__restrict__ part_d, int32_t *__restrict__ part_i) {
    __shared__ double s_d[kFallbackThreads / 32];
    __shared__ int s_i[kFallbackThreads / 32];
    __shared__ double s_last_d;
    __shared__ int s_last_i;
    __shared__ float s_q[MEPOL_MAX_DIM];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    int n_rows = *count;
    if (n_rows > max_rows) n_rows = max_rows;
    const int chunk = blockIdx.x;
    const int per = (n_ref + n_chunks - 1) / n_chunks;
    const int lo = chunk * per;
    const int hi = lo + per < n_ref ? lo + per : n_ref;
    for (int slot = blockIdx.y; slot < n_rows; slot += gridDim.y) {
        const int row = list[slot];
        __syncthreads();
        if (tid < d) s_q[tid] = q[(size_t)row * d + tid];
        if (tid == 0) {
            s_last_d = -1.0;
            s_last_i = -1;
        }
        __syncthreads();
        for (int m = 0; m < k1; ++m) {
            const double ld = s_last_d;
            const int li = s_last_i;
            double bd = INF;
            int bi = 0x7fffffff;
            for (int j = lo + tid; j < hi; j += kFallbackThreads) {
                const double dj = exact_d2(s_q, ref + (size_t)j * d, d);
                if (key_less(ld, li, dj, j) && key_less(dj, j, bd, bi)) {
                    bd = dj;
                    bi = j;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, bd, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (key_less(od, oi, bd, bi)) {
                    bd = od;
                    bi = oi;
                }
            }
            if (lane == 0) {
                s_d[warp] = bd;
                s_i[warp] = bi;
            }
            __syncthreads();
            if (tid == 0) {
                for (int w = 1; w < kFallbackThreads / 32; ++w)
                    if (key_less(s_d[w], s_i[w], bd, bi)) {
                        bd = s_d[w];
                        bi = s_i[w];
                    }
                const size_t o = ((size_t)slot * n_chunks + chunk) * k1 + m;
                part_d[o] = bd;       // INF / 0x7fffffff when the chunk has fewer than m+1 references
                part_i[o] = bi;
                s_last_d = bd;
                s_last_i = bi;
            }
            __syncthreads();
        }
    }
}

// one warp per flagged row: rank-count the n_chunks*k1 partial winners
__global__ void __launch_bounds__(128)
    fallback_merge_kernel(const int32_t *__restrict__ list, const int *__restrict__ count, int max_rows, int n_chunks,
                          int k1, const double *__restrict__ part_d, const int32_t *__restrict__ part_i,
                          int32_t *__restrict__ idx_out, double *__restrict__ dist_out, double *__restrict__ rk_out,
                          unsigned long long *__restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * 4 + (threadIdx.x >> 5);
    int n_rows = *count;
    if (n_rows > max_rows) n_rows = max_rows;
    if (slot >= n_rows) return;
    const int row = list[slot];
    const int total = n_chunks * k1;
    const double *pd = part_d + (size_t)slot * total;
    const int32_t *pi = part_i + (size_t)slot * total;
    for (int e = lane; e < total; e += 32) {
        const double de = pd[e];
        const int ie = pi[e];
        if (ie == 0x7fffffff) continue;
        int rank = 0;
        for (int o = 0; o < total; ++o) rank += key_less(pd[o], pi[o], de, ie) ? 1 : 0;
        if (rank < k1) {
            idx_out[(size_t)row * k1 + rank] = ie;
            const double r = sqrt(de);
            dist_out[(size_t)row * k1 + rank] = r;
            if (rk_out && rank == k1 - 1) rk_out[row] = r;
        }
    }
    if (lane == 0 && stats) atomicAdd(stats + 1, 1ULL);
}

int fallback_chunks_for(int k1) { return k1 <= 16 ? 32 : 8; }

int launch_exact_fallback(const float *ref, const float *q, const KnnPlan &p, int32_t *idx_out,
                          double *dist_out, double *rk_out, cudaStream_t st) {
    // parallel path for the first kFbMaxRows flagged rows ...
    const int n_chunks = fallback_chunks_for(p.k1);
    MEPOL_CUDA_CHECK(cudaMemsetAsync(p.fb_count, 0, sizeof(int), st));
    fallback_list_kernel<<<(unsigned)((p.n_q + 255) / 256), 256, 0, st>>>(p.flag, (int)p.n_q, p.fb_list, p.fb_count,
                                                                      p.fb_max_rows);
    {
        dim3 grid((unsigned)n_chunks, (unsigned)(p.fb_max_rows < 64 ? p.fb_max_rows : 64));
        fallback_partial_kernel<<<grid, kFallbackThreads, 0, st>>>(ref, (int)p.n_ref, q, p.d, p.k1, p.fb_list,
                                                                   p.fb_count, p.fb_max_rows, n_chunks, p.fb_part_d,
                                                                   p.fb_part_i);
        fallback_merge_kernel<<<(unsigned)((p.fb_max_rows + 3) / 4), 128, 0, st>>>(
            p.fb_list, p.fb_count, p.fb_max_rows, n_chunks, p.k1, p.fb_part_d, p.fb_part_i, idx_out, dist_out, rk_out,
            p.stats);
    }
    // ... and the serial-per-row kernel for whatever is left (massive duplication)
    int blocks = device_sm_count() * 4;
    const int n_row_chunks = (int)((p.n_q + kFallbackThreads - 1) / kFallbackThreads);
    if (blocks > n_row_chunks) blocks = n_row_chunks;
    if (blocks < 1) blocks = 1;
    knn_exact_fallback_kernel<<<blocks, kFallbackThreads, 0, st>>>(ref, (int)p.n_ref, q, (int)p.n_q, p.d,
                                                                   p.k1, p.flag, idx_out, dist_out, rk_out,
                                                                   p.stats);
    MEPOL_CUDA_CHECK(cudaGetLastError());
    return MEPOL_OK;
}

}  // namespace mepol
