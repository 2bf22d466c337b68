// K1 (tensor-core form): tcgen05 bf16 pruning pass of the exact k-NN search for sm_100a.
//
// Replaces the same reference hot loop as knn_simt.cu (sklearn ArgKmin,
// _argkmin.pyx.tp:471-510: dgemm middle term + norms + heap_push per pair), but the
// ||y||^2 - 2 x.y score tile is produced by the 5th-generation tensor cores:
//
//   A (queries)    [x~_0 .. x~_{d-1}, 1, 1, 1, 0..]            bf16, K = 32
//   B (references) [-2y~_0 .. -2y~_{d-1}, n_hi, n_mid, n_lo, 0..] bf16, K = 32
//   D = A . B^T  = ||y'||^2 - 2 x~.y~   (fp32 accumulator in TMEM; the row-constant ||x'||^2 is
//                                         dropped, it does not change a row's ranking)
//
// so the accumulator IS the ranking key and the epilogue does no arithmetic on it: one
// tcgen05.ld (32 columns of the thread's own row), a 3-input-min tree (FMNMX3), one compare with
// the row's pruning threshold, and a rare slow path that appends survivors to the row's candidate
// buffer (knn_internal.cuh).  Exactness comes from the proven error band (see tc_band) and the
// fp64 re-rank, never from bf16.
//
// CTA = 256 query rows (two M=128 accumulators) x a stream of 128-reference B tiles.
//   warp 8      bulk-copy producer (cp.async.bulk -> smem, mbarrier complete_tx), 6-stage ring
//   warp 9      TMEM allocator + single-thread tcgen05.mma issuer (cta_group::1, M128 N128 K16 x2)
//   warps 0-7   epilogue: warp w owns TMEM lanes 32*(w%4).. of accumulator w/4, one row per thread
// TMEM: 2 buffers x 2 accumulators x 128 columns = 512 columns (all of it) -> MMA of tile j+1
// overlaps the epilogue of tile j.  Persistent grid (one CTA per SM) over row blocks.
//
// Operand tiles are pre-arranged in global memory by knn_prep.cu in the UMMA canonical K-major
// no-swizzle order (8x16-byte core matrices, 128 B each): a 128-row x 32-col tile is 8 KB
// contiguous, so one 1-D bulk copy per tile lands it in shared memory ready for the descriptor:
//   byte offset(row r, 16-byte k-chunk c) = ((r/8)*4 + c)*128 + (r%8)*16
//   descriptor: LBO (next core matrix along K) = 128 B, SBO (next 8-row group) = 512 B.
#include <cuda_bf16.h>

#include "knn_internal.cuh"

namespace mepol {

constexpr int kTcRows = 256;
constexpr int kTcTileRefs = 128;
constexpr int kTcStages = 6;
constexpr int kTcTileBytes = 128 * kTcKPad * 2;  // 8192
constexpr int kTcSoftCap = 32;                   // candidates a row may carry between column groups / emit
constexpr int kTcCap = 64;                       // physical buffer entries per row (soft cap + one 32-column group)
constexpr int kTcBufStride = kTcCap + 1;         // row-major, padded: conflict-free both per-lane and per-row
constexpr int kTcEpiWarps = 8;
constexpr int kTcThreads = (kTcEpiWarps + 2) * 32;
constexpr int kTcTmemCols = 512;

constexpr int kSmemA = 0;
constexpr int kSmemB = kSmemA + 2 * kTcTileBytes;
constexpr int kSmemBufV = kSmemB + kTcStages * kTcTileBytes;
constexpr int kSmemBufI = kSmemBufV + kTcBufStride * kTcRows * 4;
constexpr int kSmemBar = (kSmemBufI + kTcBufStride * kTcRows * 4 + 127) / 128 * 128;
constexpr int kSmemTotal = kSmemBar + 512;

bool tc_supported(int d, int k1) { return d + 3 <= kTcKPad && 2 * k1 + 8 <= kTcSoftCap; }

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (clean launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 25)) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t *r = reinterpret_cast<uint32_t *>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// UMMA shared-memory descriptor, K-major, no swizzle (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1, [61,64) layout type = 0.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// Instruction descriptor (InstrDescriptor): c_format F32 (bit 4), a/b format BF16 (bits 7, 10),
// K-major A and B, N>>3 at bit 17, M>>4 at bit 24.
constexpr uint32_t kTcIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTcTileRefs >> 3) << 17) |
                              ((uint32_t)(128 >> 4) << 24);

// Error band of the tensor-core score for one query row (DESIGN.md "pruning band"):
//   s  = ||y'||^2 - 2 x'.y'  (exact),   s~ = TC( [x~,1,1,1] . [-2y~, n_hi, n_mid, n_lo] )
//   |x~.y~ - x'.y'| <= ||x~|| ||y~-y'|| + ||x~-x'|| ||y'|| <= (a+dx) D + dx A
//   |n_hi+n_mid+n_lo - ||y'||^2| <= 2^-24 A^2,  accumulation <= 2^-20 * sum|terms|
// with a = ||x'||, dx = ||x~-x'|| (this row), A = max ||y'||, D = max ||y~-y'|| (fitted set).
__device__ __forceinline__ float tc_band(float a, float dx, float A, float D) {
    const float dot_err = (a + dx) * D + dx * A;
    const float acc = 9.5367432e-7f * (2.0f * (a + dx) * (A + D) + 1.01f * A * A);
    const float e = 1.05f * (2.0f * dot_err + 5.9604645e-8f * A * A + acc);
    return 2.0f * e + 1e-37f;
}

template <bool kDump>
__global__ void __launch_bounds__(kTcThreads, 1)
    knn_prune_tc_kernel(const uint16_t *__restrict__ qb, const float *__restrict__ q_anorm,
                        const float *__restrict__ q_delta, int n_q, const uint16_t *__restrict__ rb, int n_ref,
                        int n_ref_tiles, int n_row_blocks, const float *__restrict__ glob, int k1,
                        uint32_t lbo_bytes, uint32_t sbo_bytes, int32_t *__restrict__ cand,
                        int32_t *__restrict__ cand_cnt, unsigned char *__restrict__ flag,
                        unsigned long long *__restrict__ stats, float *__restrict__ dump, int epi_mode) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // barrier slots
    const uint32_t bar_full = sbase + kSmemBar;                    // [stages]
    const uint32_t bar_empty = bar_full + 8 * kTcStages;           // [stages]
    const uint32_t bar_acc_full = bar_empty + 8 * kTcStages;       // [2]
    const uint32_t bar_acc_empty = bar_acc_full + 16;              // [2]
    const uint32_t bar_a_full = bar_acc_empty + 16;
    const uint32_t bar_a_empty = bar_a_full + 8;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + kSmemBar + 8 * (2 * kTcStages + 6));

    if (threadIdx.x == 0) {
        for (int s = 0; s < kTcStages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_acc_full + 8 * b, 1);
            mbar_init(bar_acc_empty + 8 * b, kTcEpiWarps);
        }
        mbar_init(bar_a_full, 1);
        mbar_init(bar_a_empty, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kTcEpiWarps + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)kTcTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == kTcEpiWarps) {
        // ===================== bulk-copy producer =====================
        if (lane == 0) {
            uint32_t it = 0, rbi = 0;
            for (int rblk = blockIdx.x; rblk < n_row_blocks; rblk += gridDim.x, ++rbi) {
                mbar_wait(bar_a_empty, (rbi & 1) ^ 1);
                mbar_arrive_expect_tx(bar_a_full, 2 * kTcTileBytes);
                bulk_g2s(sbase + kSmemA, qb + (size_t)rblk * kTcRows * kTcKPad, 2 * kTcTileBytes, bar_a_full);
                const int tiles = kDump ? 1 : n_ref_tiles;
                // every CTA starts its sweep over the reference tiles at a different offset so that
                // the 148 SMs do not hammer the same L2 lines at the same moment
                const int rot = kDump ? 0 : (int)(((long long)blockIdx.x * n_ref_tiles) / gridDim.x);
                for (int j = 0; j < tiles; ++j, ++it) {
                    int jt = j + rot;
                    if (jt >= n_ref_tiles) jt -= n_ref_tiles;
                    const uint32_t st = it % kTcStages;
                    mbar_wait(bar_empty + 8 * st, ((it / kTcStages) & 1) ^ 1);
                    mbar_arrive_expect_tx(bar_full + 8 * st, kTcTileBytes);
                    bulk_g2s(sbase + kSmemB + st * kTcTileBytes, rb + (size_t)jt * kTcTileRefs * kTcKPad, kTcTileBytes,
                             bar_full + 8 * st);
                }
            }
        }
    } else if (warp == kTcEpiWarps + 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t it = 0, rbi = 0, acc_it = 0;
            for (int rblk = blockIdx.x; rblk < n_row_blocks; rblk += gridDim.x, ++rbi) {
                mbar_wait(bar_a_full, rbi & 1);
                const int tiles = kDump ? 1 : n_ref_tiles;
                for (int j = 0; j < tiles; ++j, ++it, ++acc_it) {
                    const uint32_t st = it % kTcStages;
                    const uint32_t buf = acc_it & 1;
                    mbar_wait(bar_acc_empty + 8 * buf, ((acc_it >> 1) & 1) ^ 1);
                    mbar_wait(bar_full + 8 * st, (it / kTcStages) & 1);
                    tc_fence_after();
                    const uint32_t b_addr = sbase + kSmemB + st * kTcTileBytes;
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        const uint32_t a_addr = sbase + kSmemA + t * kTcTileBytes;
                        const uint32_t d_addr = tmem_base + buf * 256 + t * 128;
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
                            // one K=16 step = two 16-byte k-chunks = 256 bytes further along K
                            const uint64_t ad = umma_desc(a_addr + ks * 256, lbo_bytes, sbo_bytes);
                            const uint64_t bd = umma_desc(b_addr + ks * 256, lbo_bytes, sbo_bytes);
                            tc_mma_bf16(d_addr, ad, bd, kTcIdesc, ks);
                        }
                    }
                    tc_commit(bar_empty + 8 * st);        // smem stage reusable once these MMAs retire
                    tc_commit(bar_acc_full + 8 * buf);    // accumulators ready for the epilogue
                }
                tc_commit(bar_a_empty);                   // query tiles reusable
            }
        }
    } else {
        // ===================== epilogue: one query row per thread =====================
        const int t = warp >> 2;                       // accumulator (row tile) of this warp
        const int row_in_tile = (warp & 3) * 32 + lane;
        const int rloc = t * 128 + row_in_tile;        // row inside the CTA's 256-row block
        const int warp_row0 = rloc - lane;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        float *bufv = reinterpret_cast<float *>(smem + kSmemBufV);
        int *bufi = reinterpret_cast<int *>(smem + kSmemBufI);
        const int my_base = rloc * kTcBufStride;
        const float A = glob[G_AMAX], D = glob[G_DELTA_MAX];
        const float INF = __int_as_float(0x7f800000);
        uint32_t acc_it = 0;
        unsigned n_compactions = 0;
        unsigned long long n_cand = 0;

        // Cooperative compaction of lane L's buffer (up to 64 entries, two per lane): the k1-th
        // smallest is found by rank counting over 64 shuffles, tau tightens to t + band, survivors
        // are re-packed with ballot prefixes.  A few hundred warp instructions, no serial scan.
        auto coop_compact = [&](int L, int &cnt, float &tau, float band, bool &ovf) {
            const int base = (warp_row0 + L) * kTcBufStride;
            const int n = __shfl_sync(0xffffffffu, cnt, L);
            const float tau_L = __shfl_sync(0xffffffffu, tau, L);
            const float band_L = __shfl_sync(0xffffffffu, band, L);
            const bool ok0 = lane < n, ok1 = lane + 32 < n;
            const float v0 = ok0 ? bufv[base + lane] : INF;
            const float v1 = ok1 ? bufv[base + 32 + lane] : INF;
            const int i0 = ok0 ? bufi[base + lane] : -1;
            const int i1 = ok1 ? bufi[base + 32 + lane] : -1;
            int c0 = 0, c1 = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float o0 = __shfl_sync(0xffffffffu, v0, j);
                const float o1 = __shfl_sync(0xffffffffu, v1, j);
                c0 += (o0 <= v0 ? 1 : 0) + (o1 <= v0 ? 1 : 0);
                c1 += (o0 <= v1 ? 1 : 0) + (o1 <= v1 ? 1 : 0);
            }
            float tk = fminf((c0 >= k1) ? v0 : INF, (c1 >= k1) ? v1 : INF);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tk = fminf(tk, __shfl_xor_sync(0xffffffffu, tk, o));
            const float nt = fminf(tau_L, tk + band_L);
            const bool k0 = ok0 && v0 <= nt, k1b = ok1 && v1 <= nt;
            const unsigned m0 = __ballot_sync(0xffffffffu, k0);
            const unsigned m1 = __ballot_sync(0xffffffffu, k1b);
            const unsigned lt = (1u << lane) - 1u;
            const int p0 = __popc(m0 & lt), p1 = __popc(m0) + __popc(m1 & lt);
            __syncwarp();
            if (k0) {
                bufv[base + p0] = v0;
                bufi[base + p0] = i0;
            }
            if (k1b) {
                bufv[base + p1] = v1;
                bufi[base + p1] = i1;
            }
            __syncwarp();
            if (lane == L) {
                tau = nt;
                cnt = __popc(m0) + __popc(m1);
                ++n_compactions;
                if (cnt > kTcSoftCap) {                // the band alone overfills the buffer: exact fallback row
                    ovf = true;
                    cnt = 0;
                    tau = -INF;
                }
            }
        };

        for (int rblk = blockIdx.x; rblk < n_row_blocks; rblk += gridDim.x) {
            const int row = rblk * kTcRows + rloc;
            const bool live = row < n_q;
            float tau = live ? INF : -INF;             // padding rows never collect anything
            const float band = live ? tc_band(q_anorm[row], q_delta[row], A, D) : 0.f;
            int cnt = 0;
            bool ovf = false;
            const int tiles = kDump ? 1 : n_ref_tiles;
            const int rot = kDump ? 0 : (int)(((long long)blockIdx.x * n_ref_tiles) / gridDim.x);
            for (int j = 0; j < tiles; ++j, ++acc_it) {
                int jt = j + rot;
                if (jt >= n_ref_tiles) jt -= n_ref_tiles;
                const uint32_t buf = acc_it & 1;
                mbar_wait(bar_acc_full + 8 * buf, (acc_it >> 1) & 1);
                tc_fence_after();
                const uint32_t tbase = tmem_base + lane_base + buf * 256 + t * 128;
                // pull the thread's whole 128-column row into registers and hand the TMEM buffer
                // back to the MMA issuer at once: the filter below overlaps the next tile's MMA
                float v[4][32];
                __syncwarp();
                if (epi_mode != 2) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) tc_ld32(tbase + c * 32, v[c]);
                    tc_wait_ld();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_acc_empty + 8 * buf);
                if (epi_mode != 0) continue;          // bring-up timing modes: no filter
                if (kDump) {
                    if (rblk == 0)
#pragma unroll
                        for (int c = 0; c < 4; ++c)
#pragma unroll
                            for (int jj = 0; jj < 32; ++jj) dump[(size_t)rloc * kTcTileRefs + c * 32 + jj] = v[c][jj];
                    continue;
                }
                float gm[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float m[11];
#pragma unroll
                    for (int g = 0; g < 10; ++g) m[g] = fmin3(v[c][3 * g], v[c][3 * g + 1], v[c][3 * g + 2]);
                    m[10] = fminf(v[c][30], v[c][31]);
                    const float m0 = fmin3(m[0], m[1], m[2]), m1 = fmin3(m[3], m[4], m[5]);
                    const float m2 = fmin3(m[6], m[7], m[8]), m3 = fminf(m[9], m[10]);
                    gm[c] = fminf(fmin3(m0, m1, m2), m3);
                }
                const float mm = fminf(fminf(gm[0], gm[1]), fminf(gm[2], gm[3]));
                if (__any_sync(0xffffffffu, mm < tau)) {
                    // warp-uniform slow path.  Invariant: cnt <= kTcSoftCap (32) on entry to a
                    // 32-column group, a group appends at most 32, the buffer holds 64.
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (!__any_sync(0xffffffffu, gm[c] < tau)) continue;
                        const int ref0 = jt * kTcTileRefs + c * 32;
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) {
                            if (v[c][jj] < tau && ref0 + jj < n_ref) {
                                bufv[my_base + cnt] = v[c][jj];
                                bufi[my_base + cnt] = ref0 + jj;
                                ++cnt;
                            }
                        }
                        unsigned full = __ballot_sync(0xffffffffu, cnt > kTcSoftCap);
                        while (full) {
                            const int L = __ffs(full) - 1;
                            full &= full - 1;
                            coop_compact(L, cnt, tau, band, ovf);
                        }
                    }
                }
            }
            if (!kDump) {
                // final tightening + coalesced write-out of every row of the warp
                unsigned need = __ballot_sync(0xffffffffu, live && !ovf && cnt > k1);
                while (need) {
                    const int L = __ffs(need) - 1;
                    need &= need - 1;
                    coop_compact(L, cnt, tau, band, ovf);
                }
                if (live) {
                    if (ovf) {
                        flag[row] = 1;
                        cand_cnt[row] = 0;
                    } else {
                        cand_cnt[row] = cnt;
                        n_cand += cnt;
                    }
                }
                __syncwarp();
                for (int L = 0; L < 32; ++L) {
                    const int nL = __shfl_sync(0xffffffffu, (live && !ovf) ? cnt : 0, L);
                    const int rowL = rblk * kTcRows + warp_row0 + L;
                    if (lane < nL) cand[(size_t)rowL * kTcSoftCap + lane] = bufi[(warp_row0 + L) * kTcBufStride + lane];
                }
                __syncwarp();
            }
        }
        if (!kDump && stats) {
            n_cand = (unsigned long long)warp_sum((double)n_cand);
            unsigned long long nc = (unsigned long long)warp_sum((double)n_compactions);
            if (lane == 0) {
                atomicAdd(stats + 0, n_cand);
                atomicAdd(stats + 2, nc);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kTcEpiWarps + 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTcTmemCols)
                     : "memory");
    }
}

static void tc_layout_params(uint32_t &lbo, uint32_t &sbo) {
    lbo = 128;
    sbo = 512;
    const char *sw = getenv("MEPOL_TC_SWAP_LBO_SBO");   // bring-up aid only
    if (sw && sw[0] == '1') {
        lbo = 512;
        sbo = 128;
    }
}

int launch_prune_tc(const KnnPlan &p, cudaStream_t st) {
    uint32_t lbo, sbo;
    tc_layout_params(lbo, sbo);
    auto kern = knn_prune_tc_kernel<false>;
    MEPOL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    const int n_row_blocks = (int)(p.n_q_pad / kTcRows);
    const int n_ref_tiles = (int)(p.n_ref_pad / kTcTileRefs);
    int grid = device_sm_count();
    if (grid > n_row_blocks) grid = n_row_blocks;
    const char *em = getenv("MEPOL_TC_EPI_MODE");        // bring-up aid only (results invalid when set)
    const int epi_mode = em ? atoi(em) : 0;
    kern<<<grid, kTcThreads, kSmemTotal, st>>>(p.qb, p.q_anorm, p.q_delta, (int)p.n_q, p.rb, (int)p.n_ref, n_ref_tiles,
                                               n_row_blocks, p.glob, p.k1, lbo, sbo, p.cand, p.cand_cnt, p.flag,
                                               p.stats, nullptr, epi_mode);
    MEPOL_CUDA_CHECK(cudaGetLastError());
    return MEPOL_OK;
}

// Bring-up / test hook: raw accumulator tile (first 256 query rows x first 128 references).
int launch_tc_dump(const KnnPlan &p, float *dump, cudaStream_t st) {
    uint32_t lbo, sbo;
    tc_layout_params(lbo, sbo);
    auto kern = knn_prune_tc_kernel<true>;
    MEPOL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    kern<<<1, kTcThreads, kSmemTotal, st>>>(p.qb, p.q_anorm, p.q_delta, (int)p.n_q, p.rb, (int)p.n_ref, 1, 1, p.glob,
                                            p.k1, lbo, sbo, p.cand, p.cand_cnt, p.flag, nullptr, dump, 0);
    MEPOL_CUDA_CHECK(cudaGetLastError());
    return MEPOL_OK;
}

}  // namespace mepol
