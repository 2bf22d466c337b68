// K1 (SIMT form): fp32 pruning pass of the exact k-NN search.
//
// Replaces the hot loop of sklearn's brute ArgKmin
// (sklearn/metrics/_pairwise_distances_reduction/_argkmin.pyx.tp:471-510: GEMM middle term + norms
//  + heap_push for every (query, reference) pair).  Three uses:
//   * the search itself for small particle dimensions (GridWorld / MountainCar / Ant-7d: a 2..7-wide
//     contraction cannot feed tensor cores and fp16 is far too coarse for their tiny neighbour
//     radii), for k+1 > 12 (GridWorld k=50) and for d > 29 (Humanoid-47);
//   * the RESCUE pass behind the tensor-core pruner: rows whose fp16 error band overflowed the
//     candidate buffer (clustered / trajectory-like data, where neighbour distances are small
//     compared with ||x||) are re-pruned here with the ~1000x tighter fp32 band before anything
//     falls back to the exact float64 brute force.
//
// Work decomposition: one thread owns ROWS query rows (their centred coordinates live in
// registers) and streams over reference tiles staged in shared memory with cp.async double
// buffering; every lane of a warp reads the same reference element, so the tile reads are
// conflict-free broadcasts.  The score  s = ||y'||^2 - 2 x'.y'  (the row-constant ||x'||^2 is
// dropped) is compared against the row's pruning threshold; survivors are appended to the row's
// buffer in shared memory (row-major, padded: conflict-free per lane and per row).  When a buffer
// fills, the WARP compacts it cooperatively (rank counting, ballot re-pack) -- a serial per-thread
// selection cost 90 ms at k=50 (profiles/).  Exactness comes from the error band + fp64 re-rank.
#include "knn_internal.cuh"

namespace mepol {

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// fp32 error bound of the score for a row with centred norm a (A = max reference norm), u = 2^-24:
//   |fl(x^.y^) - x'.y'| <= (2u + u^2) a A + gamma_d (1+u)^2 a A,   |n^ - ||y'||^2| <= u A^2,
//   final fma rounding <= u (A^2 + 2 a A).   e <= u((2d + 6) a A + 2 A^2); 1.25x slack on top.
__device__ __forceinline__ float simt_band(float a, float A, int d) {
    const float u = 5.9604645e-8f;
    float e = 1.25f * u * ((2.0f * d + 6.0f) * a * A + 2.0f * A * A);
    return 2.0f * e + 1e-37f;
}

constexpr int kSimtGroup = 4;   // references per inner step = most appends a row can make between checks

template <int DPAD, int ROWS>
__global__ void __launch_bounds__(kSimtThreads)
    knn_prune_simt_kernel(const float *__restrict__ qs, const float *__restrict__ q_anorm, int n_q,
                          const float *__restrict__ rs, const float *__restrict__ r_nrm, int n_ref,
                          const float *__restrict__ glob, int d_true, int k1, int cap, int n_split,
                          int32_t *__restrict__ cand, int32_t *__restrict__ cand_cnt, unsigned char *__restrict__ flag,
                          int rescue, unsigned long long rescue_min_rows, int addr_nsplit,
                          unsigned long long *__restrict__ stats) {
    // Rescue mode only pays when MANY rows overflowed the fp16 band (its CTAs sweep every reference
    // for whatever flagged rows they own); a sparse set is cheaper in the chunk-parallel exact fallback.
    if (rescue && stats[3] <= rescue_min_rows) return;
    constexpr int R = kSimtThreads * ROWS;
    constexpr int BN = kSimtTileRefs;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *tile = reinterpret_cast<float *>(smem_raw);                 // [2][BN][DPAD]
    float *tnrm = tile + 2 * BN * DPAD;                                // [2][BN]
    const int stride = cap + 1;
    float *bufv = tnrm + 2 * BN;                                       // [R][cap+1]
    int *bufi = reinterpret_cast<int *>(bufv + (size_t)stride * R);    // [R][cap+1]
    const int soft = cap - kSimtGroup;                                 // a row may hold this many between checks

    const int tid = threadIdx.x, lane = tid & 31, warp0 = tid - lane;
    const int split = blockIdx.y;
    const float INF = __int_as_float(0x7f800000);
    const int tiles_total = (n_ref + BN - 1) / BN;
    const int tiles_per = (tiles_total + n_split - 1) / n_split;
    const int tile_lo = split * tiles_per;
    int tile_hi = tile_lo + tiles_per;
    if (tile_hi > tiles_total) tile_hi = tiles_total;

    float q[ROWS][DPAD];
    float tau[ROWS], band[ROWS];
    int cnt[ROWS], row[ROWS];
    bool ovf[ROWS], active[ROWS];
    const float A = glob[G_AMAX];
    bool any_active = false;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        row[r] = blockIdx.x * R + r * kSimtThreads + tid;
        // rescue mode: only rows the tensor-core pruner flagged (and that exist) are searched
        active[r] = row[r] < n_q && (!rescue || flag[row[r]] == 1);
        any_active |= active[r];
        const int rr = row[r] < n_q ? row[r] : n_q - 1;
        const float4 *src = reinterpret_cast<const float4 *>(qs + (size_t)rr * DPAD);
#pragma unroll
        for (int c = 0; c < DPAD / 4; ++c) {
            float4 v = src[c];
            q[r][4 * c] = v.x; q[r][4 * c + 1] = v.y; q[r][4 * c + 2] = v.z; q[r][4 * c + 3] = v.w;
        }
        tau[r] = active[r] ? INF : -INF;              // inactive rows never collect anything
        band[r] = simt_band(q_anorm[rr], A, d_true);
        cnt[r] = 0;
        ovf[r] = false;
    }
    if (!__syncthreads_or(any_active ? 1 : 0)) return;   // whole CTA has nothing to do (common in rescue mode)
    unsigned n_compactions = 0;

    // Cooperative compaction of the buffer of (row slot r, lane L): up to `cap` entries, up to four
    // per lane.  rank counting finds the smallest value with >= k1 entries at or below it, tau
    // tightens to it + band, survivors are re-packed in order with ballot prefixes.
    auto coop_compact = [&](int r, int L) {
        const int base = (r * kSimtThreads + warp0 + L) * stride;
        const int n = __shfl_sync(0xffffffffu, cnt[r], L);
        const float tau_L = __shfl_sync(0xffffffffu, tau[r], L);
        const float band_L = __shfl_sync(0xffffffffu, band[r], L);
        float v[4];
        int id[4], c[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = lane + 32 * i;
            v[i] = e < n ? bufv[base + e] : INF;
            id[i] = e < n ? bufi[base + e] : -1;
            c[i] = 0;
        }
        for (int e2 = 0; e2 < n; ++e2) {
            const float o = bufv[base + e2];           // broadcast read
#pragma unroll
            for (int i = 0; i < 4; ++i) c[i] += (o <= v[i]) ? 1 : 0;
        }
        float tk = INF;
#pragma unroll
        for (int i = 0; i < 4; ++i) tk = fminf(tk, (c[i] >= k1) ? v[i] : INF);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tk = fminf(tk, __shfl_xor_sync(0xffffffffu, tk, o));
        const float nt = fminf(tau_L, tk + band_L);
        __syncwarp();
        int offset = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool keep = (lane + 32 * i) < n && v[i] <= nt;
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int pos = offset + __popc(m & ((1u << lane) - 1u));
                bufv[base + pos] = v[i];               // every entry was read into registers above
                bufi[base + pos] = id[i];
            }
            offset += __popc(m);
        }
        __syncwarp();
        if (lane == L) {
            tau[r] = nt;
            cnt[r] = offset;
            ++n_compactions;
            if (offset > soft) {                       // the band alone overfills the buffer
                ovf[r] = true;
                cnt[r] = 0;
                tau[r] = -INF;
            }
        }
    };

    auto load_tile = [&](int t, int stage) {
        const int ref0 = t * BN;
        float *dst = tile + stage * BN * DPAD;
        constexpr int VEC = BN * DPAD / 4;
        for (int v = tid; v < VEC; v += kSimtThreads) {
            int rj = (v * 4) / DPAD;
            int src_row = ref0 + rj < n_ref ? ref0 + rj : n_ref - 1;
            cp_async16(dst + v * 4, rs + (size_t)src_row * DPAD + (v * 4) % DPAD);
        }
        if (tid < BN) {
            int src_row = ref0 + tid;
            // out-of-range references get an infinite norm -> score +inf -> never collected
            tnrm[stage * BN + tid] = src_row < n_ref ? r_nrm[src_row] : INF;
        }
    };

    if (tile_lo < tile_hi) load_tile(tile_lo, 0);
    cp_async_commit();
    for (int t = tile_lo; t < tile_hi; ++t) {
        const int stage = (t - tile_lo) & 1;
        if (t + 1 < tile_hi) load_tile(t + 1, stage ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const float *ts = tile + stage * BN * DPAD;
        const float *tn = tnrm + stage * BN;
        const int ref0 = t * BN;
#pragma unroll 1
        for (int j0 = 0; j0 < BN; j0 += kSimtGroup) {
            float acc[ROWS][4];
#pragma unroll
            for (int r = 0; r < ROWS; ++r)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) acc[r][jj] = 0.f;
#pragma unroll
            for (int c = 0; c < DPAD; c += 4) {
                float4 y[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                    y[jj] = *reinterpret_cast<const float4 *>(ts + (j0 + jj) * DPAD + c);
#pragma unroll
                for (int r = 0; r < ROWS; ++r)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        acc[r][jj] = fmaf(q[r][c], y[jj].x, acc[r][jj]);
                        acc[r][jj] = fmaf(q[r][c + 1], y[jj].y, acc[r][jj]);
                        acc[r][jj] = fmaf(q[r][c + 2], y[jj].z, acc[r][jj]);
                        acc[r][jj] = fmaf(q[r][c + 3], y[jj].w, acc[r][jj]);
                    }
            }
            const float4 nn = *reinterpret_cast<const float4 *>(tn + j0);
            const float nv[4] = {nn.x, nn.y, nn.z, nn.w};
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const int base = (r * kSimtThreads + tid) * stride;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const float s = fmaf(-2.0f, acc[r][jj], nv[jj]);
                    if (s < tau[r]) {   // strict: +inf scores of padding references never enter
                        bufv[base + cnt[r]] = s;
                        bufi[base + cnt[r]] = ref0 + j0 + jj;
                        ++cnt[r];
                    }
                }
                // warp-uniform check: any row of this slot past its soft capacity?
                unsigned full = __ballot_sync(0xffffffffu, cnt[r] > soft);
                while (full) {
                    const int L = __ffs(full) - 1;
                    full &= full - 1;
                    coop_compact(r, L);
                }
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();

    unsigned long long n_cand = 0;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        unsigned need = __ballot_sync(0xffffffffu, active[r] && !ovf[r] && cnt[r] > k1);
        while (need) {
            const int L = __ffs(need) - 1;
            need &= need - 1;
            coop_compact(r, L);
        }
        if (active[r]) {
            // addr_nsplit = candidate lists per row in the output layout (differs from n_split only in
            // rescue mode, where this single sweep replaces all lists of the row)
            const size_t slot = (size_t)row[r] * addr_nsplit + split;
            if (ovf[r]) {
                flag[row[r]] = 1;
                cand_cnt[slot] = 0;
            } else {
                cand_cnt[slot] = cnt[r];
                n_cand += cnt[r];
                if (rescue) {
                    flag[row[r]] = 0;                  // re-pruned successfully: back to the normal re-rank
                    for (int s2 = 1; s2 < addr_nsplit; ++s2) cand_cnt[slot + s2] = 0;
                }
            }
        }
        __syncwarp();
        // coalesced write-out of the 32 rows of this warp / slot
        for (int L = 0; L < 32; ++L) {
            const int nL = __shfl_sync(0xffffffffu, (active[r] && !ovf[r]) ? cnt[r] : 0, L);
            const int rowL = __shfl_sync(0xffffffffu, row[r], L);
            const int baseL = (r * kSimtThreads + warp0 + L) * stride;
            for (int e = lane; e < nL; e += 32) cand[((size_t)rowL * addr_nsplit + split) * cap + e] = bufi[baseL + e];
        }
        __syncwarp();
    }
    if (stats) {
        n_cand = (unsigned long long)warp_sum((double)n_cand);
        unsigned long long nc = (unsigned long long)warp_sum((double)n_compactions);
        if (lane == 0) {
            atomicAdd(stats + 0, n_cand);
            atomicAdd(stats + 2, nc);
        }
    }
}

template <int DPAD, int ROWS>
static int launch_simt_inst(const KnnPlan &p, int n_split, int cap, int rescue, cudaStream_t st) {
    const int R = kSimtThreads * ROWS;
    const size_t smem = (size_t)(2 * kSimtTileRefs * DPAD + 2 * kSimtTileRefs) * 4 + (size_t)(cap + 1) * R * 8;
    auto kern = knn_prune_simt_kernel<DPAD, ROWS>;
    MEPOL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((p.n_q + R - 1) / R), (unsigned)n_split);
    kern<<<grid, kSimtThreads, smem, st>>>(p.qs, p.q_anorm, (int)p.n_q, p.rs, p.r_nrm, (int)p.n_ref, p.glob, p.d,
                                           p.k1, cap, n_split, p.cand, p.cand_cnt, p.flag, rescue,
                                           (unsigned long long)rescue_min_rows(p), rescue ? p.n_split : n_split, p.stats);
    MEPOL_CUDA_CHECK(cudaGetLastError());
    return MEPOL_OK;
}

static int launch_simt_any(const KnnPlan &p, int rows_per_thread, int n_split, int cap, int rescue, cudaStream_t st) {
#define MEPOL_SIMT_CASE(DP)                                                                                  \
    case DP:                                                                                                 \
        return rows_per_thread == 2 ? launch_simt_inst<DP, 2>(p, n_split, cap, rescue, st)                   \
                                    : launch_simt_inst<DP, 1>(p, n_split, cap, rescue, st);
    switch (p.dpad) {
        MEPOL_SIMT_CASE(4)
        MEPOL_SIMT_CASE(8)
        MEPOL_SIMT_CASE(16)
        MEPOL_SIMT_CASE(24)
        MEPOL_SIMT_CASE(32)
        case 48:
            return launch_simt_inst<48, 1>(p, n_split, cap, rescue, st);
        case 64:
            return launch_simt_inst<64, 1>(p, n_split, cap, rescue, st);
        default:
            break;
    }
#undef MEPOL_SIMT_CASE
    set_error("knn simt: unsupported padded dimension %d", p.dpad);
    return MEPOL_E_UNSUPPORTED;
}

int launch_prune_simt(const KnnPlan &p, cudaStream_t st) {
    return launch_simt_any(p, p.rows_per_thread, p.n_split, p.cap, 0, st);
}

// Rescue pass behind the tensor-core pruner: same candidate layout as knn_tc.cu (one list of 32 per row).
int launch_rescue_simt(const KnnPlan &p, cudaStream_t st) {
    return launch_simt_any(p, p.dpad <= 32 ? 2 : 1, 1, 32, 1, st);
}

}  // namespace mepol
