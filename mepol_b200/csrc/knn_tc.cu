// K1 (tensor-core form): tcgen05 fp16 pruning pass of the exact k-NN search for sm_100a.
//
// Replaces the same reference hot loop as knn_simt.cu (sklearn ArgKmin,
// _argkmin.pyx.tp:471-510: dgemm middle term + norms + heap_push per pair), but the
// ||y||^2 - 2 x.y score tile is produced by the 5th-generation tensor cores:
//
//   A (queries)    [x~_0 .. x~_{d-1}, 1, 1, 1, 0..]            fp16, K = 32
//   B (references) [-2y~_0 .. -2y~_{d-1}, n_hi, n_mid, n_lo, 0..] fp16, K = 32
//   D = A . B^T  = ||y'||^2 - 2 x~.y~   (fp32 accumulator in TMEM; the row-constant ||x'||^2 is
//                                         dropped, it does not change a row's ranking)
//
// so the accumulator IS the ranking key and the epilogue does no arithmetic on it: tcgen05.ld of
// the thread's own 128-column row, a 3-input-min tree (FMNMX3), one compare with the row's pruning
// threshold, and a rare slow path that appends survivors to the row's candidate buffer.
// Exactness comes from the proven error band (tc_band) and the fp64 re-rank, never from fp16.
//
// The operands are fp16 (not bf16): knn_prep.cu scales the centred data by a power of two so that
// every fitted row norm is <= 128, which keeps -2y~ and ||y'||^2 inside fp16's range and makes the
// band ~8x tighter than bf16's (5.3 instead of 8.5 candidates per row at N = 1M, d = 29).
//
// CTA = 256 query rows (two M=128 accumulators) x a stream of 128-reference B tiles.
//   warp 8      bulk-copy producer (cp.async.bulk -> smem, mbarrier complete_tx), 4-stage ring
//   warp 9      TMEM allocator + tcgen05.mma issuer of accumulator 0 (cta_group::1, M128 N128 K16 x2)
//   warp 10     tcgen05.mma issuer of accumulator 1
//   warps 0-7   epilogue: warp w owns TMEM lanes 32*(w%4).. of accumulator w/4, one row per thread
// TMEM: 2 buffers x 2 accumulators x 128 columns = 512 columns (all of it).  The epilogue pulls a
// tile into registers and releases the TMEM buffer before filtering, so the MMA of tile j+1/j+2
// overlaps the filter of tile j.  Persistent grid (one CTA per SM) over row blocks.
//
// Measured on B200 (profiles/r01_tc_pipeline_notes.md): with K = 32 the tensor pipe needs only 256
// cycles per tile; the kernel is bound by two serial loops of fixed latencies that are both ~870
// cycles per tile -- per buffer: release -> issuer sees it (~390) -> two MMAs + commit issued (~320)
// -> MMA retires -> epilogue's barrier query (~200) -> tcgen05.ld (~160) -> release; per epilogue
// warp: query + load + min tree (~270) + survivor path (~70 average) + loop (~180).  An mbarrier
// query costs ~200 cycles even when the phase has completed; TMEM (4 fp32 accumulators) caps the
// depth at two buffers.  The structure below keeps every role's serial chain per tile short.
//
// Operand tiles are pre-arranged in global memory by knn_prep.cu in the UMMA canonical K-major
// no-swizzle order (8x16-byte core matrices, 128 B each): a 128-row x 32-col tile is 8 KB
// contiguous, so one 1-D bulk copy per tile lands it in shared memory ready for the descriptor:
//   byte offset(row r, 16-byte k-chunk c) = ((r/8)*4 + c)*128 + (r%8)*16
//   descriptor: LBO (next core matrix along K) = 128 B, SBO (next 8-row group) = 512 B.
#include <type_traits>

#include "tc_ptx.cuh"

namespace mepol {

constexpr int kTcRows = 256;
constexpr int kTcStages = 4;
constexpr int kTcSoftCap = 32;                   // candidates a row may carry between column groups / emit
constexpr int kTcCap = 64;                       // physical buffer entries per row (soft cap + one 32-column group)
constexpr int kTcBufStride = kTcCap + 1;         // row-major, padded: conflict-free both per-lane and per-row
constexpr int kTcEpiWarps = 8;
constexpr int kTcThreads = (kTcEpiWarps + 4) * 32;   // 8 epilogue warps + producer + issuer + 2 idle (whole warpgroup for setmaxnreg)
constexpr int kTcEpiRegs = 216;                        // setmaxnreg budget of the epilogue warpgroups
constexpr int kTcWarmTiles = 256;                      // tau-only tiles at the start of a unit (see warm_of)
constexpr int kTcAuxRegs = 40;                         //   ... and of the producer/issuer warpgroup

constexpr int kSmemA = 0;
constexpr int kSmemB = kSmemA + 2 * kTcTileBytes;
constexpr int kSmemBufV = kSmemB + kTcStages * kTcTileBytes;
constexpr int kSmemBufI = kSmemBufV + kTcBufStride * kTcRows * 4;
constexpr int kSmemStage = kSmemBufI + kTcBufStride * kTcRows * 4;            // per-warp 32x33 transpose tiles
constexpr int kSmemBar = (kSmemStage + kTcEpiWarps * 32 * 33 * 4 + 127) / 128 * 128;
constexpr int kSmemTotal = kSmemBar + 512;

bool tc_supported(int d, int k1) { return d + 3 <= kTcKPad && 2 * k1 + 8 <= kTcSoftCap; }

#ifdef MEPOL_BRINGUP
#define TC_PROF_NOW() clock64()
#else
#define TC_PROF_NOW() 0LL
#endif

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }
__device__ __forceinline__ float ld_volatile_f32(const float *p) { return *reinterpret_cast<const volatile float *>(p); }

// kVar (experiment switch, MEPOL_TC_VARIANT): 0 = TMEM buffers are handed back to the issuers through
// mbarriers; 1 = through named hardware barriers (bar.arrive by the 128 epilogue threads of an
// accumulator, bar.sync by its issuer warp).
template <bool kDump, int K1, int kVar>
__global__ void __launch_bounds__(kTcThreads, 1)
    knn_prune_tc_kernel(const uint16_t *__restrict__ qb, const float *__restrict__ q_anorm,
                        const float *__restrict__ q_delta, int n_q, const uint16_t *__restrict__ rb, int n_ref,
                        int n_ref_tiles, int n_row_blocks, const float *__restrict__ glob, int k1,
                        int32_t *__restrict__ cand, int32_t *__restrict__ cand_cnt, unsigned char *__restrict__ flag,
                        unsigned long long *__restrict__ stats, float *__restrict__ dump, int epi_mode_arg, int n_full_blocks,
                        int tail_split, int n_split, int warm_tiles, int k_steps) {
    extern __shared__ __align__(1024) unsigned char smem[];
#if defined(MEPOL_BRINGUP) && defined(MEPOL_BRINGUP_MODES)
    const int epi_mode = epi_mode_arg;      // bring-up timing modes (results invalid when non-zero)
#else
    constexpr int epi_mode = 0;             // shipped build: the per-tile mode checks fold away
    (void)epi_mode_arg;
#endif
    const uint32_t sbase = smem_u32(smem);
    // warp index through a shuffle: tells the compiler it is warp-uniform, so that per-role addresses
    // (TMEM, barriers, descriptors) live in uniform registers and tcgen05 operands need no R2UR waterfall
    const int warp = kVar >= 2 ? __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0) : (int)(threadIdx.x >> 5), lane = threadIdx.x & 31;
    // barrier slots.  done0/done1[s]: the MMAs of accumulator 0 / 1 that read B stage s have retired
    // (tcgen05.commit): their accumulator is ready, and once both fired the stage may be refilled.
    const uint32_t bar_full = sbase + kSmemBar;                    // [stages] producer -> issuers
    const uint32_t bar_done = bar_full + 8 * kTcStages;            // [stages] issuer 1 -> producer + epilogue of accumulator 1
    const uint32_t bar_done0 = bar_done + 8 * kTcStages;           // [stages] issuer 0 -> producer + epilogue of accumulator 0
    const uint32_t bar_acc_empty = bar_done0 + 8 * kTcStages;      // [2 buffers][2 accumulators] epilogue -> issuer
    const uint32_t bar_a_full = bar_acc_empty + 32;
    const uint32_t bar_a_empty = bar_a_full + 8;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + kSmemBar + 8 * (3 * kTcStages + 6));
    volatile uint32_t *ts_s = reinterpret_cast<volatile uint32_t *>(smem + kSmemBar + 320);   // bring-up time stamps [16]

    if (threadIdx.x == 0) {
        for (int s = 0; s < kTcStages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_done + 8 * s, 1);
            mbar_init(bar_done0 + 8 * s, 1);
        }
        for (int b = 0; b < 4; ++b) mbar_init(bar_acc_empty + 8 * b, kTcEpiWarps / 2);
        mbar_init(bar_a_full, 1);
        mbar_init(bar_a_empty, 2);
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
    // Work units.  Row blocks [0, n_full_blocks) are whole units (all reference tiles).  The remaining
    // "tail" blocks -- the last, partially filled round of the persistent grid, which matters when a
    // multi-GPU shard leaves only a few rounds -- are cut into tail_split reference ranges each, so the
    // last round keeps all SMs busy for 1/tail_split of a round.  Splits of a row write separate
    // candidate lists (slot sp of n_split) that the re-rank merges.
    const int n_units = n_full_blocks + (n_row_blocks - n_full_blocks) * tail_split;
    const int tiles_per_split = (n_ref_tiles + tail_split - 1) / tail_split;
    auto unit_of = [&](int u, int &rblk, int &sp, int &tile0) {
        if (u < n_full_blocks) {
            rblk = u;
            sp = 0;
            tile0 = 0;
            return kDump ? 1 : n_ref_tiles;
        }
        const int v = u - n_full_blocks;
        rblk = n_full_blocks + v / tail_split;
        sp = v - (v / tail_split) * tail_split;
        tile0 = sp * tiles_per_split;
        int hi = tile0 + tiles_per_split;
        if (hi > n_ref_tiles) hi = n_ref_tiles;
        return hi - tile0;
    };
    // Warm-up tiles.  A streaming top-k1 appends ~k1*ln(n) survivors per row, most of them while tau
    // is still loose.  The first `warm` tiles of a unit are therefore only used to bound the k1-th
    // smallest score: the k1 smallest TILE MINIMA (k1 distinct references) are tracked branch-free in
    // registers and nothing is appended; the same tiles are swept again at the end of the unit, now
    // against a tight tau.  ~3 % more MMA work at N = 1M removes two thirds of the compactions; for
    // short sweeps up to half of the tiles are warmed (measured 2-4 % faster at N = 32k-131k than an
    // eighth, profiles/r01_tc_pipeline_notes.md).
    auto warm_of = [&](int tiles) { return (K1 > 0 && !kDump) ? min(warm_tiles, tiles >> 1) : 0; };
    // Register re-distribution at warpgroup granularity (setmaxnreg): the epilogue keeps a whole
    // 128-column row in registers, the producer / issuer / idle warps need almost none.
    // (each role branch issues its own setmaxnreg so that ptxas budgets the branch it sits in)
#define TC_AUX_SETMAXNREG() asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kTcAuxRegs))

    if (warp == kTcEpiWarps) {
        // ===================== bulk-copy producer =====================
        TC_AUX_SETMAXNREG();
        if (lane == 0) {
            uint32_t it = 0, rbi = 0;
            for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++rbi) {
                int rblk, sp, tile0;
                const int tiles = unit_of(u, rblk, sp, tile0);
                mbar_wait(bar_a_empty, (rbi & 1) ^ 1);
                mbar_arrive_expect_tx(bar_a_full, 2 * kTcTileBytes);
                bulk_g2s(sbase + kSmemA, qb + (size_t)rblk * kTcRows * kTcKPad, 2 * kTcTileBytes, bar_a_full);
                // The sweep over the reference tiles starts at the row block's own position: CTAs do
                // not hit the same L2 lines at the same moment, and in a self-search the rows' own
                // neighbourhood (same trajectory) is seen first, which tightens tau early.
                const int rot = kDump ? 0 : (rblk * 2) % tiles;
                const int sweep = tiles + warm_of(tiles);
                for (int j = 0; j < sweep; ++j, ++it) {
                    int jt = j + rot;
                    if (jt >= tiles) jt -= tiles;
                    if (jt >= tiles) jt -= tiles;
                    jt += tile0;
                    const uint32_t st = it % kTcStages;
                    mbar_wait(bar_done0 + 8 * st, ((it / kTcStages) & 1) ^ 1);   // both accumulators' MMAs
                    mbar_wait(bar_done + 8 * st, ((it / kTcStages) & 1) ^ 1);    // have read the stage
                    mbar_arrive_expect_tx(bar_full + 8 * st, kTcTileBytes);
                    bulk_g2s(sbase + kSmemB + st * kTcTileBytes, rb + (size_t)jt * kTcTileRefs * kTcKPad, kTcTileBytes,
                             bar_full + 8 * st);
                }
            }
        }
    } else if (warp == kTcEpiWarps + 1 || warp == kTcEpiWarps + 2) {
        // ===================== MMA issuers (one per accumulator) =====================
        // Warp 9 drives accumulator 0 (query rows 0-127), warp 10 accumulator 1 (rows 128-255): each
        // waits only for its own four epilogue warps and issues its own two K=16 steps, so the two
        // row tiles are independent pipelines that merely share the B stages.  A tcgen05 instruction
        // costs the issuing thread ~70 cycles, and one thread issuing all four MMAs + two commits of
        // a tile (~470 cycles) sat on the critical release -> MMA -> ready loop of accumulator 1.
        // The whole warp runs the loop so that the two barrier waits of a tile are ONE warp-wide
        // try_wait (lane 1 polls full[stage], the other lanes poll acc_empty[buffer]); lane 0 issues.
        TC_AUX_SETMAXNREG();
        const int t = warp - (kTcEpiWarps + 1);
        uint64_t a_desc[2];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) a_desc[ks] = umma_desc(sbase + kSmemA + t * kTcTileBytes + ks * 256, 128, 512);
        const uint64_t b_desc0 = umma_desc(sbase + kSmemB, 128, 512);
        const uint32_t my_commit = t == 0 ? bar_done0 : bar_done;
        uint32_t it = 0, rbi = 0;
        unsigned long long chain[2] = {0, 0};   // bring-up only: release->detect, detect->issued
        for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++rbi) {
            int rblk, sp, tile0;
            const int tiles = unit_of(u, rblk, sp, tile0);
            mbar_wait(bar_a_full, rbi & 1);
            const int sweep = tiles + warm_of(tiles);
            for (int j = 0; j < sweep; ++j, ++it) {
                const uint32_t st = it % kTcStages;
                const uint32_t buf = it & 1;
                uint32_t spins = 0;
                if (kVar == 1) {
                    // stage first (long complete by the time the buffer comes back), then the hardware
                    // barrier the accumulator's 128 epilogue threads arrive on after their TMEM load
                    while (!__all_sync(0xffffffffu, mbar_try_wait(bar_full + 8 * st, (it / kTcStages) & 1))) {
                        if (++spins > (1u << 25)) __trap();
                    }
                    if (it >= 2) named_bar_sync(1 + buf * 2 + t, 160);
                } else {
                    const uint32_t my_bar = (lane == 1) ? bar_full + 8 * st : bar_acc_empty + 8 * (buf * 2 + t);
                    const uint32_t my_par = (lane == 1) ? ((it / kTcStages) & 1) : (((it >> 1) & 1) ^ 1);
                    while (!__all_sync(0xffffffffu, mbar_try_wait(my_bar, my_par))) {
                        if (++spins > (1u << 25)) __trap();
                    }
                }
                tc_fence_after();
#ifdef MEPOL_BRINGUP
                const uint32_t ck_detect = (uint32_t)clock64();
                if (t == 1 && lane == 0 && it >= 2) chain[0] += (uint32_t)(ck_detect - ts_s[(it - 2) & 7]);
#endif
                if (kVar >= 2 ? elect_one() : lane == 0) {
                    // one K=16 step = two 16-byte k-chunks = 256 bytes further along K
                    const uint64_t bd0 = b_desc0 + (uint64_t)((st * kTcTileBytes) >> 4);
                    const uint64_t bd1 = bd0 + (256 >> 4);
                    const uint32_t d0 = tmem_base + buf * 256 + t * 128;
                    tc_mma_f16(d0, a_desc[0], bd0, kTcIdesc, 0);
                    if (k_steps > 1) tc_mma_f16(d0, a_desc[1], bd1, kTcIdesc, 1);   // d + 3 <= 16 fits one K=16 step
                    tc_commit(my_commit + 8 * st);
#ifdef MEPOL_BRINGUP
                    const uint32_t ck_issued = (uint32_t)clock64();
                    if (t == 1) {
                        ts_s[8 + (it & 7)] = ck_issued;
                        chain[1] += (uint32_t)(ck_issued - ck_detect);
                    }
#endif
                }
                __syncwarp();
            }
            if (lane == 0) tc_commit(bar_a_empty);   // query tiles reusable once every MMA of the block retired
            __syncwarp();
        }
        if (kVar == 1) {
            // consume the arrivals of the last two tiles so that every barrier ends the kernel idle
            for (uint32_t i = it >= 2 ? it - 2 : 0; i < it; ++i) named_bar_sync(1 + (i & 1) * 2 + t, 160);
        }
#ifdef MEPOL_BRINGUP
        if (dump && lane == 0 && !kDump && t == 1) {
            atomicAdd(reinterpret_cast<unsigned long long *>(dump) + 6, chain[0]);
            atomicAdd(reinterpret_cast<unsigned long long *>(dump) + 7, chain[1]);
            atomicAdd(reinterpret_cast<unsigned long long *>(dump) + 10, (unsigned long long)it);
        }
#endif
    } else if (warp < kTcEpiWarps) {
        // ===================== epilogue: one query row per thread =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kTcEpiRegs));
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
        // Loop-invariant addresses are made opaque to the compiler: it otherwise re-derives them from
        // S2R (tid, cluster CTA id) reads in every tile, a ~60-cycle dependent chain in front of the
        // TMEM load.
        uint32_t my_done = t == 0 ? bar_done0 : bar_done;
        uint32_t tmem_row = tmem_base + lane_base + t * 128;
        uint32_t acc_empty_addr = bar_acc_empty + 8 * t;
        asm volatile("mov.u32 %0, %0;\n\tmov.u32 %1, %1;\n\tmov.u32 %2, %2;" : "+r"(my_done), "+r"(tmem_row), "+r"(acc_empty_addr));
        uint32_t it = 0;
        unsigned n_compactions = 0;
        unsigned long long n_cand = 0;
        long long prof[6] = {0, 0, 0, 0, 0, 0};   // bring-up only: cycles in wait / ld / tree / slow, slow entries, tiles
        unsigned long long chainE[2] = {0, 0};
        bool next_ready = false;

        float *stg = reinterpret_cast<float *>(smem + kSmemStage) + warp * (32 * 33);   // per-warp transpose tile
        // Running k1 smallest scores of the thread's row, ascending (K1 > 0 only): tau follows the
        // CURRENT k1-th smallest + band instead of waiting for the next compaction, which halves the
        // number of survivors that have to be appended.  Inserting +inf is a no-op.
        float topk[K1 > 0 ? K1 : 1];
        auto topk_insert = [&](float x) {
#pragma unroll
            for (int i = 0; i < (K1 > 0 ? K1 : 0); ++i) {
                const float lo = fminf(topk[i], x);
                x = fmaxf(topk[i], x);
                topk[i] = lo;
            }
        };

        // Cooperative compaction of lane L's buffer (up to 64 entries, two per lane).  The k1-th
        // smallest comes from the running list (K1 > 0) or from rank counting over 64 shuffles
        // (generic k1); tau tightens to t + band, survivors are re-packed with ballot prefixes.
        auto coop_compact = [&](int L, int &cnt, float &tau, float band, bool &ovf) {
            const int base = (warp_row0 + L) * kTcBufStride;
            const int n = __shfl_sync(0xffffffffu, cnt, L);
            const float tau_L = __shfl_sync(0xffffffffu, tau, L);
            const bool ok0 = lane < n, ok1 = lane + 32 < n;
            const float v0 = ok0 ? bufv[base + lane] : INF;
            const float v1 = ok1 ? bufv[base + 32 + lane] : INF;
            const int i0 = ok0 ? bufi[base + lane] : -1;
            const int i1 = ok1 ? bufi[base + 32 + lane] : -1;
            float nt = tau_L;
            if (K1 == 0) {
                const float band_L = __shfl_sync(0xffffffffu, band, L);
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
                nt = fminf(tau_L, tk + band_L);
            }
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

        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            int rblk, sp, tile0;
            const int tiles = unit_of(u, rblk, sp, tile0);
            const int row = rblk * kTcRows + rloc;
            const bool live = row < n_q;
            float tau = live ? INF : -INF;             // padding rows never collect anything
            const float band = live ? tc_band(q_anorm[row], q_delta[row], A, D) : 0.f;
            int cnt = 0;
            bool ovf = false;
#pragma unroll
            for (int i = 0; i < (K1 > 0 ? K1 : 1); ++i) topk[i] = INF;
            const int rot = kDump ? 0 : (rblk * 2) % tiles;
            const int warm = warm_of(tiles), sweep = tiles + warm;
            // The tile body is instantiated twice (warm-up tiles / sweep proper) so that neither loop
            // carries the other's per-tile branches on its critical path.
            auto tile = [&](auto warm_c, int j) {
                constexpr bool kWarmTile = decltype(warm_c)::value;
                int jt = j + rot;
                if (jt >= tiles) jt -= tiles;
                if (jt >= tiles) jt -= tiles;
                jt += tile0;
                const uint32_t buf = it & 1;
                const uint32_t st = it % kTcStages;
                const long long pt0 = TC_PROF_NOW();
                if (!next_ready) mbar_wait(my_done + 8 * st, (it / kTcStages) & 1);
                tc_fence_after();
                const long long pt1 = TC_PROF_NOW();
                const uint32_t tbase = tmem_row + buf * 256;
                // pull the thread's whole 128-column row into registers and hand the TMEM buffer
                // back to the MMA issuer at once: the filter below overlaps the next tiles' MMAs
                float v[4][32];
                __syncwarp();                          // .sync.aligned TMEM ops need a converged warp
                if (epi_mode != 2) tc_ld128(tbase, v);
                {   // probe the next tile's barrier while the TMEM loads are in flight
                    const uint32_t st1 = (it + 1) % kTcStages;
                    next_ready = mbar_test_wait(my_done + 8 * st1, ((it + 1) / kTcStages) & 1);
                }
                if (epi_mode != 2) tc_wait_ld();
                tc_fence_before();
                if (kVar == 1) {
                    named_bar_arrive(1 + buf * 2 + t, 160);
                } else {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty_addr + 16 * buf);
                }
                const long long pt2 = TC_PROF_NOW();
#ifdef MEPOL_BRINGUP
                if (warp == 7 && lane == 0) {
                    ts_s[it & 7] = (uint32_t)pt2;                                  // release of this tile's buffer
                    chainE[0] += (uint32_t)((uint32_t)pt1 - ts_s[8 + (it & 7)]);   // issued -> seen ready
                    chainE[1] += (uint32_t)(pt2 - pt1);                            // seen -> loaded + released
                }
#endif
                prof[0] += pt1 - pt0;
                prof[1] += pt2 - pt1;
                ++prof[5];
                if (epi_mode == 1 || epi_mode == 2) return;   // bring-up timing modes: no filter
                if (kDump) {
                    if (rblk == 0)
#pragma unroll
                        for (int c = 0; c < 4; ++c)
#pragma unroll
                            for (int jj = 0; jj < 32; ++jj) dump[(size_t)rloc * kTcTileRefs + c * 32 + jj] = v[c][jj];
                    return;
                }
                float gm[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    // 32 -> 1 in 16 instructions (the minimum for 3-input mins): 10 triples + the last two
                    // raw values make 12, four triples make 4, then one 3-input and one 2-input min
                    float m[10];
#pragma unroll
                    for (int g = 0; g < 10; ++g) m[g] = fmin3(v[c][3 * g], v[c][3 * g + 1], v[c][3 * g + 2]);
                    const float m0 = fmin3(m[0], m[1], m[2]), m1 = fmin3(m[3], m[4], m[5]);
                    const float m2 = fmin3(m[6], m[7], m[8]), m3 = fmin3(m[9], v[c][30], v[c][31]);
                    gm[c] = fminf(fmin3(m0, m1, m2), m3);
                    if (c == 1 && !next_ready) {
                        // second probe of the next tile's barrier, anchored inside the min tree: the
                        // first one (under the TMEM load) is usually a little too early, and the
                        // blocking wait at the top of the loop costs ~200 cycles even when complete
                        const uint32_t st1 = (it + 1) % kTcStages;
                        next_ready = mbar_test_wait_after(my_done + 8 * st1, ((it + 1) / kTcStages) & 1, gm[c]);
                    }
                }
                const float mm = fminf(fmin3(gm[0], gm[1], gm[2]), gm[3]);
                if (kWarmTile) {
                    // warm-up tile: the tile minimum is the score of one reference of this tile, so the
                    // k1 smallest minima bound the k1-th smallest score from above (padding references
                    // carry a score above every real one and never are a minimum)
                    topk_insert(mm);
                    if (live) tau = topk[K1 > 0 ? K1 - 1 : 0] + band;
                    return;
                }
                const long long pt3 = TC_PROF_NOW();
                prof[2] += pt3 - pt2;
                if (__any_sync(0xffffffffu, mm < tau) || epi_mode == 3) {
                    ++prof[4];
                    // Slow path, per lane and shuffle-free.  Each lane builds the bit mask of its own
                    // row's survivors in the 32-column group straight from registers, parks the group
                    // in its private row of the staging tile (registers cannot be indexed by a run-time
                    // column) and then walks the mask: lanes with survivors run the short loop
                    // together, the others idle.  No ballot / shuffle chain and no cross-lane
                    // dependency, so a typical late entry (one survivor in one row) is ~150 cycles.
                    // Invariant: cnt <= kTcSoftCap (32) on entry, a group appends <= 32, buffer = 64.
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (!(__any_sync(0xffffffffu, gm[c] < tau) || (epi_mode == 3 && c == 0))) continue;
                        const int ref0 = jt * kTcTileRefs + c * 32;
                        unsigned mask = 0;
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) mask |= (v[c][jj] < tau ? 1u : 0u) << jj;
                        if (ref0 + 32 > n_ref) mask &= ref0 < n_ref ? (1u << (n_ref - ref0)) - 1u : 0u;   // padding columns
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) stg[lane * 33 + jj] = v[c][jj];
                        while (mask) {
                            const int jj = __ffs(mask) - 1;
                            mask &= mask - 1;
                            const float x = stg[lane * 33 + jj];
                            if (x < tau) {                 // tau may have tightened inside this loop
                                bufv[my_base + cnt] = x;
                                bufi[my_base + cnt] = ref0 + jj;
                                ++cnt;
                                if (K1 > 0) {
                                    topk_insert(x);
                                    tau = fminf(tau, topk[K1 > 0 ? K1 - 1 : 0] + band);
                                }
                            }
                        }
                        unsigned full = __ballot_sync(0xffffffffu, cnt > kTcSoftCap);
                        while (full) {
                            const int L = __ffs(full) - 1;
                            full &= full - 1;
                            coop_compact(L, cnt, tau, band, ovf);
                        }
                    }
                    prof[3] += TC_PROF_NOW() - pt3;
                }
            };
            int j = 0;
            for (; j < warm; ++j, ++it) tile(std::true_type{}, j);
            if (warm > 0) {
                // the sweep proper sees every tile once more: restart the list, keep tau
#pragma unroll
                for (int i = 0; i < (K1 > 0 ? K1 : 1); ++i) topk[i] = INF;
            }
            for (; j < sweep; ++j, ++it) tile(std::false_type{}, j);
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
                        cand_cnt[(size_t)row * n_split + sp] = 0;
                        if (stats) atomicAdd(stats + 3, 1ULL);   // rows whose fp16 band overflowed
                    } else {
                        cand_cnt[(size_t)row * n_split + sp] = cnt;
                        n_cand += cnt;
                    }
                }
                __syncwarp();
                for (int L = 0; L < 32; ++L) {
                    const int nL = __shfl_sync(0xffffffffu, (live && !ovf) ? cnt : 0, L);
                    const int rowL = rblk * kTcRows + warp_row0 + L;
                    if (lane < nL)
                        cand[((size_t)rowL * n_split + sp) * kTcSoftCap + lane] = bufi[(warp_row0 + L) * kTcBufStride + lane];
                }
                __syncwarp();
            }
        }
#ifdef MEPOL_BRINGUP
        if (!kDump && dump && lane == 0)
            for (int i = 0; i < 6; ++i) atomicAdd(reinterpret_cast<unsigned long long *>(dump) + i, (unsigned long long)prof[i]);
        if (!kDump && dump && warp == 7 && lane == 0) {
            atomicAdd(reinterpret_cast<unsigned long long *>(dump) + 8, chainE[0]);
            atomicAdd(reinterpret_cast<unsigned long long *>(dump) + 9, chainE[1]);
        }
#endif
        if (!kDump && stats) {
            n_cand = (unsigned long long)warp_sum((double)n_cand);
            unsigned long long nc = (unsigned long long)warp_sum((double)n_compactions);
            if (lane == 0) {
                atomicAdd(stats + 0, n_cand);
                atomicAdd(stats + 2, nc);
            }
        }
    } else {
        TC_AUX_SETMAXNREG();   // (setmaxnreg is warpgroup-wide)
    }
#undef TC_AUX_SETMAXNREG

    tc_fence_before();
    __syncthreads();
    if (warp == kTcEpiWarps + 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTcTmemCols)
                     : "memory");
    }
}

int launch_prune_tc(const KnnPlan &p, cudaStream_t st) {
    // k+1 = 5 (every shipped MEPOL config but GridWorld) gets the register-resident running top-k1
    int variant = 2;                                    // MEPOL_TC_VARIANT = 0 / 1 select the round-1 forms (capi.cu::tc_variant)
    if (const char *tv = getenv("MEPOL_TC_VARIANT")) variant = atoi(tv);
    auto kern = p.k1 == 5 ? (variant == 1 ? knn_prune_tc_kernel<false, 5, 1> : variant == 2 ? knn_prune_tc_kernel<false, 5, 2> : knn_prune_tc_kernel<false, 5, 0>)
                          : (variant == 1 ? knn_prune_tc_kernel<false, 0, 1> : variant == 2 ? knn_prune_tc_kernel<false, 0, 2> : knn_prune_tc_kernel<false, 0, 0>);
    const int smem_bytes = kSmemTotal;
    MEPOL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    const int n_row_blocks = (int)(p.n_q_pad / kTcRows);
    const int n_ref_tiles = (int)(p.n_ref_pad / kTcTileRefs);
    int grid = device_sm_count();
    const int n_full_blocks = p.tc_tail_split > 1 ? (n_row_blocks / grid) * grid : n_row_blocks;
    const int n_units = n_full_blocks + (n_row_blocks - n_full_blocks) * p.tc_tail_split;
    if (grid > n_units) grid = n_units;
#if defined(MEPOL_BRINGUP) && defined(MEPOL_BRINGUP_MODES)
    const char *em = getenv("MEPOL_TC_EPI_MODE");        // bring-up timing aid (results invalid when set):
    const int epi_mode = em ? atoi(em) : 0;              // 1 = TMEM loads only, 2 = barriers + MMA only
#else
    const int epi_mode = 0;
#endif
    int warm_tiles = kTcWarmTiles;
    if (const char *wt = getenv("MEPOL_TC_WARM_TILES")) warm_tiles = atoi(wt);   // tuning aid; any value >= 0 is exact
    if (warm_tiles < 0) warm_tiles = 0;
    float *prof_dev = nullptr;
#ifdef MEPOL_BRINGUP
    static unsigned long long *prof_buf = nullptr;
    if (getenv("MEPOL_TC_PROF")) {
        if (!prof_buf) cudaMalloc(&prof_buf, 128);
        cudaMemsetAsync(prof_buf, 0, 128, st);
        prof_dev = reinterpret_cast<float *>(prof_buf);
    }
#endif
    kern<<<grid, kTcThreads, smem_bytes, st>>>(p.qb, p.q_anorm, p.q_delta, (int)p.n_q, p.rb, (int)p.n_ref, n_ref_tiles,
                                               n_row_blocks, p.glob, p.k1, p.cand, p.cand_cnt, p.flag, p.stats, prof_dev,
                                               epi_mode, n_full_blocks, p.tc_tail_split, p.n_split, warm_tiles,
                                               p.d + 3 <= 16 ? 1 : 2);
    MEPOL_CUDA_CHECK(cudaGetLastError());
#ifdef MEPOL_BRINGUP
    if (prof_dev) {
        unsigned long long h[16];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, prof_dev, sizeof(h), cudaMemcpyDeviceToHost);
        const double w = (double)grid * kTcEpiWarps, tl = (double)h[5] / w;
        fprintf(stderr, "[tc prof] per warp: tiles %.0f  cycles/tile: wait %.1f ld %.1f tree %.1f slow %.1f  slow entries/tile %.4f  cycles/entry %.0f\n",
                tl, h[0] / w / tl, h[1] / w / tl, h[2] / w / tl, h[3] / w / tl, (double)h[4] / h[5], h[4] ? (double)h[3] / h[4] : 0.0);
        if (h[10])
            fprintf(stderr, "[tc chain] cycles/tile: release->issuer detect %.1f  detect->issued %.1f  issued->epilogue sees ready %.1f  seen->loaded+released %.1f\n",
                    (double)h[6] / h[10], (double)h[7] / h[10], (double)h[8] / h[10], (double)h[9] / h[10]);
    }
#endif
    return MEPOL_OK;
}

// Test hook: raw accumulator tile (first 256 query rows x first 128 references).
int launch_tc_dump(const KnnPlan &p, float *dump, cudaStream_t st) {
    auto kern = knn_prune_tc_kernel<true, 0, 0>;
    MEPOL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    kern<<<1, kTcThreads, kSmemTotal, st>>>(p.qb, p.q_anorm, p.q_delta, (int)p.n_q, p.rb, (int)p.n_ref, 1, 1, p.glob,
                                            p.k1, p.cand, p.cand_cnt, p.flag, nullptr, dump, 0, 1, 1, 1, 0, p.d + 3 <= 16 ? 1 : 2);
    MEPOL_CUDA_CHECK(cudaGetLastError());
    return MEPOL_OK;
}

}  // namespace mepol
