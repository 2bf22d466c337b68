// C ABI entry points of the k-NN search (include/mepol_b200.h) and the search plan.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "knn_internal.cuh"

namespace mepol {

constexpr int kTcAutoMinDim = 4;        // AUTO picks the tensor-core pruner for d >= this ...
constexpr int64_t kTcAutoMinRefs = 512;   // ... and at least this many fitted rows


static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int device_sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
            cached = sms;
        else
            return 148;
    }
    return cached;
}

int mean_blocks_for(int64_t n_ref);

// Which tensor-core pruner serves MEPOL_KNN_TC (read at every launch; an A/B switch for the measurements
// in profiles/, every variant is exact):
//   unset / 3 -> knn_tc2.cu, N = 256 MMAs, fp16 accumulators, packed epilogue (the shipped kernel)
//   4 / 5     -> knn_tc2.cu with two 128-column buffers per accumulator (5: both unloaded per epilogue iteration)
//   0, 1, 2   -> knn_tc.cu, the fp32-accumulator kernel of round 1 (2 = with warp-uniform tcgen05 operands)
int tc_variant() {
    const char *tv = getenv("MEPOL_TC_VARIANT");
    return tv && *tv ? atoi(tv) : 3;
}

static int simt_dpad(int d) {
    const int opts[] = {4, 8, 16, 24, 32, 48, 64};
    for (int o : opts)
        if (d <= o) return o;
    return -1;
}

// Fills the plan; with ws == nullptr only sizes are computed (pointers are offsets from null).
int knn_make_plan(KnnPlan &p, const float *ref, int64_t n_ref, const float *q, int64_t n_q, int d, int k1,
                  int algo, void *ws) {
    MEPOL_REQUIRE(n_ref > 0 && n_q > 0 && n_ref < (1LL << 31) && n_q < (1LL << 31), MEPOL_E_INVALID,
                  "knn: n_ref=%lld n_q=%lld out of range", (long long)n_ref, (long long)n_q);
    MEPOL_REQUIRE(d >= 1 && d <= MEPOL_MAX_DIM, MEPOL_E_INVALID, "knn: d=%d outside [1,%d]", d, MEPOL_MAX_DIM);
    MEPOL_REQUIRE(k1 >= 1 && k1 <= MEPOL_MAX_K1, MEPOL_E_INVALID, "knn: k1=%d outside [1,%d]", k1, MEPOL_MAX_K1);
    // mirrors sklearn's ValueError "Expected n_neighbors <= n_samples_fit" (neighbors/_base.py:834)
    MEPOL_REQUIRE(k1 <= n_ref, MEPOL_E_INVALID, "knn: Expected n_neighbors <= n_samples_fit, got k1=%d n_ref=%lld",
                  k1, (long long)n_ref);
    MEPOL_REQUIRE(algo == MEPOL_KNN_AUTO || algo == MEPOL_KNN_SIMT || algo == MEPOL_KNN_TC || algo == MEPOL_KNN_GRID,
                  MEPOL_E_INVALID, "knn: unknown algo %d", algo);
    // AUTO: the tensor-core pruner wins from d = 4 up (2.3-5x over the SIMT pruner at N = 10k-50k on
    // gaussian, uniform and trajectory-like data, profiles/r01_tc_lowdim.md).  For d <= 3 the
    // neighbour radii are so small next to the row norms that the fp16 band passes too many
    // candidates (rows overflow into the exact fallback); there the fp32 SIMT pruner is both
    // band-tight and cheap.  Either way the result is exact.
    // d <= 3 is the reference's kd-tree regime: the uniform-grid search is O(N k) instead of O(N^2).
    if (algo == MEPOL_KNN_AUTO) {
        if (grid_supported(d, k1))
            algo = MEPOL_KNN_GRID;
        else
            algo = (tc_supported(d, k1) && n_ref >= kTcAutoMinRefs && d >= kTcAutoMinDim) ? MEPOL_KNN_TC : MEPOL_KNN_SIMT;
    }
    if (algo == MEPOL_KNN_GRID && !grid_supported(d, k1)) {
        set_error("knn: grid search does not support d=%d k1=%d", d, k1);
        return MEPOL_E_UNSUPPORTED;
    }
    if (algo == MEPOL_KNN_TC && !tc_supported(d, k1)) {
        set_error("knn: tensor-core pruner does not support d=%d k1=%d", d, k1);
        return MEPOL_E_UNSUPPORTED;
    }
    memset(&p, 0, sizeof(p));
    p.n_q = n_q;
    p.n_ref = n_ref;
    p.d = d;
    p.k1 = k1;
    p.algo = algo;
    p.dpad = simt_dpad(d);
    p.cap = k1 <= 8 ? 32 : (k1 <= 24 ? 64 : 128);
    p.mean_blocks = mean_blocks_for(n_ref);
    const int sms = device_sm_count();
    if (algo == MEPOL_KNN_SIMT) {
        p.rows_per_thread = (p.cap == 32 && p.dpad <= 32) ? 2 : 1;
        p.rows_per_cta = kSimtThreads * p.rows_per_thread;
        const int64_t row_tiles = (n_q + p.rows_per_cta - 1) / p.rows_per_cta;
        int64_t want = (2LL * sms + row_tiles - 1) / row_tiles;
        int64_t max_split = 1024 / p.cap;
        if (max_split > 16) max_split = 16;
        const int64_t ref_tiles = (n_ref + kSimtTileRefs - 1) / kSimtTileRefs;
        if (max_split > ref_tiles / 4) max_split = ref_tiles / 4;
        if (max_split < 1) max_split = 1;
        p.n_split = (int)(want < 1 ? 1 : (want > max_split ? max_split : want));
    } else if (algo == MEPOL_KNN_GRID) {
        p.rows_per_thread = 1;
        p.rows_per_cta = 256;
        p.n_split = 1;
        p.cap = 1;    // no candidate lists: the grid kernel writes the final rows itself
        p.grid_g = grid_cells_per_dim(n_ref, d, k1);
        p.grid_cells = p.grid_g;
        for (int c = 1; c < d; ++c) p.grid_cells *= p.grid_g;
    } else {
        p.rows_per_thread = 1;
        p.rows_per_cta = 256;
        p.n_split = 1;
        p.cap = 32;   // kTcSoftCap of knn_tc.cu (tc_supported guarantees 2*k1 + 8 <= 32)
        p.n_ref_pad = (n_ref + 255) / 256 * 256;   // whole 256-reference B tiles (knn_tc2.cu); knn_tc.cu takes them as 2 x 128
        p.n_q_pad = (n_q + 255) / 256 * 256;
        {
            // Tail balancing of the persistent grid: with few rounds (multi-GPU shards), the blocks of the
            // last partial round are cut into reference-range splits that together fill the SMs.
            const int64_t nrb = p.n_q_pad / 256, tiles = p.n_ref_pad / 128;
            const int64_t rounds = nrb / sms, tail = nrb - rounds * sms;
            p.tc_tail_split = 1;
            if (rounds <= 12 && tail > 0 && tiles >= 256) {
                // The tail takes ceil(tail * s / sms) rounds of 1/s length: pick the split (<= 4 ranges of >= 256
                // tiles each) that makes it shortest, the smallest such split on a draw.  Shorter ranges lose
                // more to their own warm-up and append ramp than the balance wins (N = 64k: 108 tail blocks cut
                // in 4 would finish in 0.75 of a round but measured 0.72 vs 0.62 ms, profiles/r02_tc2_notes.md).
                int64_t max_s = tiles / 256 < 4 ? tiles / 256 : 4;
                if (const char *ms = getenv("MEPOL_TC_MAX_TAIL_SPLIT")) max_s = atoi(ms);   // tuning aid; every value is exact
                double best = 1e30;
                for (int64_t s_ = 1; s_ <= max_s; ++s_) {
                    const double len = (double)((tail * s_ + sms - 1) / sms) / (double)s_;
                    if (len < best - 1e-9) {
                        best = len;
                        p.tc_tail_split = (int)s_;
                    }
                }
            }
            p.n_split = p.tc_tail_split;
        }
    }
    // a query slice of the fitted matrix can reuse the packed fp32 reference rows
    if (ref && q && q >= ref && q + n_q * d <= ref + n_ref * d && ((q - ref) % d) == 0) {
        p.q_is_ref = true;
        p.q_row0 = (q - ref) / d;
    }
    Carver c(ws);
    p.mean_partial = c.take<double>((size_t)p.mean_blocks * 3 * MEPOL_MAX_DIM);   // sum, min, max
    p.mean = c.take<double>(MEPOL_MAX_DIM + 2);   // column means + the operand scale
    p.glob = c.take<float>(G_COUNT);
    p.stats = c.take<unsigned long long>(MEPOL_KNN_NSTATS);
    p.r_anorm = c.take<float>((size_t)n_ref);
    p.q_anorm = c.take<float>((size_t)n_q);
    // fp32 operands: the SIMT pruner's input, and the rescue pass's input behind the TC pruner
    p.rs = c.take<float>((size_t)n_ref * p.dpad);
    p.r_nrm = c.take<float>((size_t)n_ref);
    if (p.q_is_ref) {
        p.qs = ws ? p.rs + p.q_row0 * p.dpad : nullptr;
        if (algo == MEPOL_KNN_SIMT) p.q_anorm = ws ? p.r_anorm + p.q_row0 : nullptr;
    } else {
        p.qs = c.take<float>((size_t)n_q * p.dpad);
    }
    if (algo == MEPOL_KNN_TC) {
        p.rb = c.take<uint16_t>((size_t)p.n_ref_pad * kTcKPad);
        p.qb = c.take<uint16_t>((size_t)p.n_q_pad * kTcKPad);
        p.q_delta = c.take<float>((size_t)n_q);
        p.q_ahalf = c.take<float>((size_t)n_q);
    }
    if (algo == MEPOL_KNN_GRID) {
        p.grid_pts = c.take<float>((size_t)n_ref * 4);
        p.grid_par = c.take<unsigned char>(grid_par_bytes());
        p.grid_start = c.take<int32_t>((size_t)p.grid_cells + 1);
        p.grid_cursor = c.take<int32_t>((size_t)p.grid_cells);
        p.grid_cell_of = c.take<int32_t>((size_t)n_ref);
    }
    p.cand = c.take<int32_t>((size_t)n_q * p.n_split * p.cap);
    p.cand_cnt = c.take<int32_t>((size_t)n_q * p.n_split);
    p.flag = c.take<unsigned char>((size_t)n_q);
    p.fb_max_rows = (int)(n_q < 16384 ? n_q : 16384);   // rows the chunk-parallel exact fallback can take
    p.fb_list = c.take<int32_t>((size_t)p.fb_max_rows);
    p.fb_count = c.take<int>(4);
    p.fb_part_d = c.take<double>((size_t)p.fb_max_rows * fallback_chunks_for(k1) * k1);
    p.fb_part_i = c.take<int32_t>((size_t)p.fb_max_rows * fallback_chunks_for(k1) * k1);
    p.bytes = c.bytes();
    return MEPOL_OK;
}

// ---- optional per-launch timing of the dominant (prune) kernel, for bench.py's roofline -------
constexpr int kProfSlots = 64;
static bool g_prof_on = false;
static cudaEvent_t g_prof_ev[kProfSlots][2];
static bool g_prof_made = false;
static int g_prof_count = 0;

static void prof_record(int which, cudaStream_t st) {
    if (!g_prof_on || g_prof_count >= kProfSlots) return;
    if (!g_prof_made) {
        for (int i = 0; i < kProfSlots; ++i) {
            cudaEventCreate(&g_prof_ev[i][0]);
            cudaEventCreate(&g_prof_ev[i][1]);
        }
        g_prof_made = true;
    }
    cudaEventRecord(g_prof_ev[g_prof_count][which], st);
    if (which == 1) ++g_prof_count;
}

}  // namespace mepol

using namespace mepol;

extern "C" int mepol_profile_enable(int on) {
    g_prof_on = on != 0;
    g_prof_count = 0;
    return MEPOL_OK;
}

extern "C" int mepol_profile_read(float *ms_out, int max_n, int *n_out) {
    MEPOL_REQUIRE(ms_out && n_out && max_n >= 0, MEPOL_E_INVALID, "profile_read: bad arguments");
    int n = g_prof_count < max_n ? g_prof_count : max_n;
    for (int i = 0; i < n; ++i) {
        MEPOL_CUDA_CHECK(cudaEventSynchronize(g_prof_ev[i][1]));
        MEPOL_CUDA_CHECK(cudaEventElapsedTime(ms_out + i, g_prof_ev[i][0], g_prof_ev[i][1]));
    }
    *n_out = n;
    g_prof_count = 0;
    return MEPOL_OK;
}

extern "C" const char *mepol_last_error(void) { return g_err; }
extern "C" int mepol_version(void) { return 100; }

extern "C" int mepol_device_info(int *sm_count, int *cc_major, int *cc_minor) {
    int dev = 0;
    MEPOL_CUDA_CHECK(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    MEPOL_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return MEPOL_OK;
}

extern "C" int mepol_knn_workspace_bytes(int64_t n_q, int64_t n_ref, int d, int k1, int algo, size_t *bytes) {
    MEPOL_REQUIRE(bytes != nullptr, MEPOL_E_INVALID, "knn_workspace_bytes: bytes is null");
    KnnPlan p;
    // worst case over aliasing: size as if the queries were packed separately
    int rc = knn_make_plan(p, nullptr, n_ref, nullptr, n_q, d, k1, algo, nullptr);
    if (rc != MEPOL_OK) return rc;
    *bytes = p.bytes;
    return MEPOL_OK;
}

extern "C" int mepol_knn_search_f32(const float *ref, int64_t n_ref, const float *q, int64_t n_q, int d, int k1,
                                    int algo, int32_t *idx_out, double *dist_out, double *rk_out, void *workspace,
                                    size_t workspace_bytes, int64_t *stats, void *stream) {
    MEPOL_REQUIRE(ref && q && idx_out && dist_out && workspace, MEPOL_E_INVALID, "knn_search: null pointer");
    KnnPlan p;
    int rc = knn_make_plan(p, ref, n_ref, q, n_q, d, k1, algo, workspace);
    if (rc != MEPOL_OK) return rc;
    MEPOL_REQUIRE(workspace_bytes >= p.bytes, MEPOL_E_WORKSPACE, "knn_search: workspace %zu < %zu bytes",
                  workspace_bytes, p.bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MEPOL_CUDA_CHECK(cudaMemsetAsync(p.flag, 0, (size_t)n_q, st));
    if (p.algo == MEPOL_KNN_TC && p.n_split > 1)   // rows of whole units only fill slot 0 of their candidate lists
        MEPOL_CUDA_CHECK(cudaMemsetAsync(p.cand_cnt, 0, sizeof(int32_t) * (size_t)n_q * p.n_split, st));
    MEPOL_CUDA_CHECK(cudaMemsetAsync(p.stats, 0, sizeof(unsigned long long) * MEPOL_KNN_NSTATS, st));
    if ((rc = launch_col_mean(ref, n_ref, d, p, st)) != MEPOL_OK) return rc;
    if (p.algo == MEPOL_KNN_GRID) {
        // no pruning pass: the grid kernel produces the final rows (rows it cannot fill go to the fallback)
        prof_record(0, st);
        rc = launch_grid_search(ref, q, p, idx_out, dist_out, rk_out, st);
        prof_record(1, st);
        if (rc != MEPOL_OK) return rc;
        if ((rc = launch_exact_fallback(ref, q, p, idx_out, dist_out, rk_out, st)) != MEPOL_OK) return rc;
        if (stats)
            MEPOL_CUDA_CHECK(cudaMemcpyAsync(stats, p.stats, sizeof(int64_t) * MEPOL_KNN_NSTATS,
                                             cudaMemcpyDeviceToDevice, st));
        return MEPOL_OK;
    }
    if ((rc = launch_pack(ref, n_ref, d, p, true, false, st)) != MEPOL_OK) return rc;
    if ((p.algo == MEPOL_KNN_TC || !p.q_is_ref) && (rc = launch_pack(q, n_q, d, p, false, false, st)) != MEPOL_OK) return rc;
    prof_record(0, st);
    if (p.algo == MEPOL_KNN_TC)
        {
            const int tv = tc_variant();
            rc = tv == 4 ? launch_prune_tc2(p, 128, 1, st) : tv == 5 ? launch_prune_tc2(p, 128, 2, st)
                 : (tv >= 0 && tv <= 2) ? launch_prune_tc(p, st) : launch_prune_tc2(p, 256, 1, st);
        }
    else
        rc = launch_prune_simt(p, st);
    prof_record(1, st);
    if (rc != MEPOL_OK) return rc;
    // rows whose fp16 band overflowed are re-pruned with the fp32 band before the exact fallback (the fp32 operands
    // of that pass are only produced when it is going to run: both launches exit at once otherwise)
    if (p.algo == MEPOL_KNN_TC) {
        if ((rc = launch_pack(ref, n_ref, d, p, true, true, st)) != MEPOL_OK) return rc;
        if (!p.q_is_ref && (rc = launch_pack(q, n_q, d, p, false, true, st)) != MEPOL_OK) return rc;
        if ((rc = launch_rescue_simt(p, st)) != MEPOL_OK) return rc;
    }
#ifdef MEPOL_TC2_DBGMODES
    if (getenv("MEPOL_TC2_DBG") || getenv("MEPOL_TC2_KSTEPS")) return MEPOL_OK;   // timing experiments: scores are invalid, stop here
#endif
#if defined(MEPOL_BRINGUP) && defined(MEPOL_BRINGUP_MODES)
    {
        const char *em = getenv("MEPOL_TC_EPI_MODE");   // bring-up timing aid: results are invalid, stop here
        if (em && atoi(em) != 0) return MEPOL_OK;
    }
#endif
    if ((rc = launch_rerank(ref, q, p, idx_out, dist_out, rk_out, st)) != MEPOL_OK) return rc;
    if ((rc = launch_exact_fallback(ref, q, p, idx_out, dist_out, rk_out, st)) != MEPOL_OK) return rc;
    if (stats)
        MEPOL_CUDA_CHECK(cudaMemcpyAsync(stats, p.stats, sizeof(int64_t) * MEPOL_KNN_NSTATS,
                                         cudaMemcpyDeviceToDevice, st));
    return MEPOL_OK;
}

// Bring-up / test hook of the tensor-core pruner: raw fp32 accumulator tile of the first 256 query
// rows against the first 128 references (scores ||y'||^2 - 2 x~.y~), out [256, 128] float (device).
extern "C" int mepol_debug_tc_scores(const float *ref, int64_t n_ref, const float *q, int64_t n_q, int d, int k1,
                                     float *out, void *workspace, size_t workspace_bytes, void *stream) {
    MEPOL_REQUIRE(ref && q && out && workspace, MEPOL_E_INVALID, "debug_tc_scores: null pointer");
    KnnPlan p;
    int rc = knn_make_plan(p, ref, n_ref, q, n_q, d, k1, MEPOL_KNN_TC, workspace);
    if (rc != MEPOL_OK) return rc;
    MEPOL_REQUIRE(workspace_bytes >= p.bytes, MEPOL_E_WORKSPACE, "debug_tc_scores: workspace %zu < %zu bytes",
                  workspace_bytes, p.bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if ((rc = launch_col_mean(ref, n_ref, d, p, st)) != MEPOL_OK) return rc;
    if ((rc = launch_pack(ref, n_ref, d, p, true, false, st)) != MEPOL_OK) return rc;
    if ((rc = launch_pack(q, n_q, d, p, false, false, st)) != MEPOL_OK) return rc;
    return launch_tc_dump(p, out, st);
}

// Same hook for knn_tc2.cu: the fp16 accumulator tile of the first 256 query rows against the first
// 256 references, widened to float, out [256, 256] float (device).
extern "C" int mepol_debug_tc2_scores(const float *ref, int64_t n_ref, const float *q, int64_t n_q, int d, int k1,
                                      float *out, void *workspace, size_t workspace_bytes, void *stream) {
    MEPOL_REQUIRE(ref && q && out && workspace, MEPOL_E_INVALID, "debug_tc2_scores: null pointer");
    KnnPlan p;
    int rc = knn_make_plan(p, ref, n_ref, q, n_q, d, k1, MEPOL_KNN_TC, workspace);
    if (rc != MEPOL_OK) return rc;
    MEPOL_REQUIRE(workspace_bytes >= p.bytes, MEPOL_E_WORKSPACE, "debug_tc2_scores: workspace %zu < %zu bytes",
                  workspace_bytes, p.bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if ((rc = launch_col_mean(ref, n_ref, d, p, st)) != MEPOL_OK) return rc;
    if ((rc = launch_pack(ref, n_ref, d, p, true, false, st)) != MEPOL_OK) return rc;
    if ((rc = launch_pack(q, n_q, d, p, false, false, st)) != MEPOL_OK) return rc;
    return launch_tc2_dump(p, out, st);
}

__global__ void widen_idx_kernel(const int32_t *__restrict__ in, int64_t *__restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

extern "C" int mepol_knn_search_host_f32(const float *x, int64_t n, int d, int k1, int algo, int64_t *idx_out,
                                         double *dist_out) {
    MEPOL_REQUIRE(x && idx_out && dist_out, MEPOL_E_INVALID, "knn_search_host: null pointer");
    size_t ws_bytes = 0;
    int rc = mepol_knn_workspace_bytes(n, n, d, k1, algo, &ws_bytes);
    if (rc != MEPOL_OK) return rc;
    float *dx = nullptr;
    int32_t *didx = nullptr;
    int64_t *didx64 = nullptr;
    double *ddist = nullptr;
    void *ws = nullptr;
    cudaStream_t st = nullptr;
    auto cleanup = [&]() {
        cudaFree(dx);
        cudaFree(didx);
        cudaFree(didx64);
        cudaFree(ddist);
        cudaFree(ws);
    };
#define MEPOL_HOST_CHECK(expr)                                                                         \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            set_error("%s failed: %s", #expr, cudaGetErrorString(_e));                                 \
            cleanup();                                                                                 \
            return MEPOL_E_CUDA;                                                                       \
        }                                                                                              \
    } while (0)
    MEPOL_HOST_CHECK(cudaMalloc(&dx, sizeof(float) * (size_t)n * d));
    MEPOL_HOST_CHECK(cudaMalloc(&didx, sizeof(int32_t) * (size_t)n * k1));
    MEPOL_HOST_CHECK(cudaMalloc(&didx64, sizeof(int64_t) * (size_t)n * k1));
    MEPOL_HOST_CHECK(cudaMalloc(&ddist, sizeof(double) * (size_t)n * k1));
    MEPOL_HOST_CHECK(cudaMalloc(&ws, ws_bytes));
    MEPOL_HOST_CHECK(cudaMemcpyAsync(dx, x, sizeof(float) * (size_t)n * d, cudaMemcpyHostToDevice, st));
    rc = mepol_knn_search_f32(dx, n, dx, n, d, k1, algo, didx, ddist, nullptr, ws, ws_bytes, nullptr, st);
    if (rc != MEPOL_OK) {
        cleanup();
        return rc;
    }
    const int64_t tot = n * (int64_t)k1;
    widen_idx_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(didx, didx64, tot);
    MEPOL_HOST_CHECK(cudaMemcpyAsync(idx_out, didx64, sizeof(int64_t) * (size_t)tot, cudaMemcpyDeviceToHost, st));
    MEPOL_HOST_CHECK(cudaMemcpyAsync(dist_out, ddist, sizeof(double) * (size_t)tot, cudaMemcpyDeviceToHost, st));
    MEPOL_HOST_CHECK(cudaStreamSynchronize(st));
#undef MEPOL_HOST_CHECK
    cleanup();
    return MEPOL_OK;
}
