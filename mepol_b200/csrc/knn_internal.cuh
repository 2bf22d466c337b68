// Internal plan / launcher declarations of the exact k-NN search (prep -> prune -> re-rank -> fallback).
#pragma once
#include "common.cuh"

namespace mepol {

constexpr int kSimtThreads = 128;  // one query row (or two) per thread
constexpr int kSimtTileRefs = 64;  // reference rows staged per shared-memory tile
constexpr double kTcScaledNormMax = 128.0;   // centred rows are scaled to norms <= this (fp16 range)
constexpr double kTcQueryAbsMax = 30000.0;   // scaled query entries beyond this do not fit fp16
constexpr int kTcKPad = 32;        // fp16 operand width of the tensor-core pruner (d + 3 norm columns <= 32)

// Element index (in fp16 units) of (row, column c) inside the tensor-core operand arrays: 128-row
// tiles of 8 KB, each tile in UMMA canonical K-major no-swizzle order (8-row x 16-byte core
// matrices of 128 B; next core matrix along K at +128 B, next 8-row group at +512 B).
__host__ __device__ inline size_t tc_elem_index(int64_t row, int c) {
    const size_t tile = (size_t)(row >> 7);
    const int r = (int)(row & 127);
    return tile * 4096 + (size_t)(((r >> 3) * 4 + (c >> 3)) * 64 + (r & 7) * 8 + (c & 7));
}

// indices into the float "glob" array produced by the pack kernels (stored as uint bits, atomicMax)
// G_AHALF_MAX: max over the fitted rows of the fp16 operand's norm over its first 16 columns (knn_tc2.cu)
enum { G_AMAX = 0, G_DELTA_MAX = 1, G_ATILDE_MAX = 2, G_AHALF_MAX = 3, G_COUNT = 4 };

struct KnnPlan {
    int64_t n_q, n_ref;
    int d, k1, algo;
    int dpad;          // SIMT operand width (multiple of 4)
    int rows_per_thread;
    int rows_per_cta;
    int cap;           // candidate slots per (row, split)
    int n_split;       // reference-range splits (more CTAs for small batches)
    int mean_blocks;
    int tc_tail_split; // reference-range splits of the row blocks in the last, partial round (knn_tc.cu)
    bool q_is_ref;     // query rows alias the packed reference rows
    int64_t q_row0;    // first reference row of the query slice when q_is_ref
    // workspace slices
    double *mean_partial, *mean;
    float *rs, *r_nrm, *r_anorm, *qs, *q_anorm;
    float *glob;
    int32_t *cand, *cand_cnt;
    unsigned char *flag;
    unsigned long long *stats;
    // tensor-core pruner operands (fp16 stored as uint16), padded to tile multiples
    uint16_t *rb, *qb;
    float *q_delta;
    float *q_ahalf;    // per query row: norm of the fp16 operand over its first 16 columns (knn_tc2.cu's band)
    int64_t n_ref_pad, n_q_pad;
    // uniform-grid search (d <= 3, knn_grid.cu)
    int grid_g, grid_cells;
    void *grid_pts, *grid_par;
    int32_t *grid_start, *grid_cursor, *grid_cell_of;
    // parallel exact fallback scratch
    int fb_max_rows;
    int32_t *fb_list, *fb_part_i;
    int *fb_count;
    double *fb_part_d;
    size_t bytes;
};

int knn_make_plan(KnnPlan &p, const float *ref, int64_t n_ref, const float *q, int64_t n_q, int d,
                  int k1, int algo, void *ws);

// knn_prep.cu
int launch_col_mean(const float *ref, int64_t n_ref, int d, const KnnPlan &p, cudaStream_t st);
int launch_pack(const float *x, int64_t n, int d, const KnnPlan &p, bool is_ref, bool lazy, cudaStream_t st);
// Rescue or fallback?  The chunk-parallel exact fallback streams every reference k1 times per flagged
// row (~F * k1 * n_ref * d); one rescue sweep costs about a SIMT pass over the row blocks that hold a
// flagged row (<= n_q * n_ref * d).  The sweep wins once F * k1 exceeds ~n_q / 8; below that, and always
// within what the fallback's scratch holds, the flagged rows stay with the fallback.
inline unsigned long long rescue_min_rows(const KnnPlan &p) {
    long long thr = p.n_q / (8LL * p.k1);
    if (thr < 64) thr = 64;
    if (thr > p.fb_max_rows) thr = p.fb_max_rows;
    return (unsigned long long)thr;
}
// knn_simt.cu
int launch_prune_simt(const KnnPlan &p, cudaStream_t st);
int launch_rescue_simt(const KnnPlan &p, cudaStream_t st);
// knn_tc.cu
bool tc_supported(int d, int k1);
int launch_prune_tc(const KnnPlan &p, cudaStream_t st);
int launch_tc_dump(const KnnPlan &p, float *dump, cudaStream_t st);
// knn_tc2.cu
int launch_prune_tc2(const KnnPlan &p, int nc, int fuse, cudaStream_t st);
int launch_tc2_dump(const KnnPlan &p, float *dump, cudaStream_t st);
// knn_grid.cu
bool grid_supported(int d, int k1);
int grid_cells_per_dim(int64_t n_ref, int d, int k1);
size_t grid_par_bytes();
int launch_grid_search(const float *ref, const float *q, const KnnPlan &p, int32_t *idx_out, double *dist_out,
                       double *rk_out, cudaStream_t st);
// knn_rerank.cu
int launch_rerank(const float *ref, const float *q, const KnnPlan &p, int32_t *idx_out,
                  double *dist_out, double *rk_out, cudaStream_t st);
int fallback_chunks_for(int k1);
int launch_exact_fallback(const float *ref, const float *q, const KnnPlan &p, int32_t *idx_out,
                          double *dist_out, double *rk_out, cudaStream_t st);

// ------------------------------------------------------------------------------------------
// Per-row candidate buffer shared by both pruners.  Each thread owns one query row; its buffer is
// a column of a [cap][rows] shared-memory array (stride = rows, so a warp touching slot e of 32
// different rows hits 32 different banks).
//
// Safety argument (DESIGN.md "pruning band"): let e bound |approx score - true score| for this
// row.  With t = k1-th smallest approximate score seen so far, any reference whose approximate
// score exceeds t + 2e has k1 references strictly closer in exact arithmetic, so it can never be
// one of the k1 nearest -- dropping it is exact.  compact() re-derives t from the buffer, tightens
// tau = t + band (band = 2e) and drops entries above it.  If the buffer stays full the row is
// flagged for the exact fallback kernel.
// ------------------------------------------------------------------------------------------
struct RowBuf {
    float *v;   // scores, element e at v[e * stride]
    int *ix;    // reference indices
    int stride;
};

__device__ __noinline__ static void rowbuf_compact(RowBuf b, int &cnt, float &tau, int k1, float band,
                                                   int cap, bool &overflow) {
    float t = __int_as_float(0x7f800000);  // +inf when fewer than k1 entries
    if (cnt >= k1) {
        float last_v = -__int_as_float(0x7f800000);
        int last_p = -1;
        for (int m = 0; m < k1; ++m) {
            float best = __int_as_float(0x7f800000);
            int bp = cnt;
            for (int p = 0; p < cnt; ++p) {
                float x = b.v[p * b.stride];
                bool after = (x > last_v) || (x == last_v && p > last_p);
                if (after && x < best) {
                    best = x;
                    bp = p;
                }
            }
            last_v = best;
            last_p = bp;
        }
        t = last_v;
    }
    float new_tau = t + band;
    if (new_tau < tau) tau = new_tau;
    int w = 0;
    for (int p = 0; p < cnt; ++p) {
        float x = b.v[p * b.stride];
        int id = b.ix[p * b.stride];
        if (x <= tau) {
            b.v[w * b.stride] = x;
            b.ix[w * b.stride] = id;
            ++w;
        }
    }
    cnt = w;
    if (cnt >= cap) {
        overflow = true;
        cnt = 0;
        tau = -__int_as_float(0x7f800000);  // nothing more is collected for a flagged row
    }
}

__device__ __forceinline__ void rowbuf_push(RowBuf b, int &cnt, float &tau, float score, int ref_idx,
                                            int k1, float band, int cap, bool &overflow,
                                            unsigned &n_compactions) {
    b.v[cnt * b.stride] = score;
    b.ix[cnt * b.stride] = ref_idx;
    ++cnt;
    if (cnt == cap) {
        rowbuf_compact(b, cnt, tau, k1, band, cap, overflow);
        ++n_compactions;
    }
}

}  // namespace mepol
