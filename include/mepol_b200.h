/*
 * mepol_b200.h -- C ABI of the B200-native MEPOL k-NN state-entropy hot path.
 *
 * The reference (RiccZamboni/mepol) has no FFI/plugin layer: its hot path is a set of module-level
 * Python functions in src/algorithms/mepol.py.  Each entry point below names the reference lines
 * whose arithmetic it replaces; INTEGRATION.md shows the ctypes stub a maintainer adds on the
 * reference side.  Conventions:
 *   - every function returns 0 on success, <0 on error (MEPOL_E_*); the message is available from
 *     mepol_last_error() (thread-local).  Nothing throws or aborts across the ABI.
 *   - device functions take caller-owned DEVICE pointers, a caller-owned workspace and a CUDA
 *     stream (cudaStream_t passed as void*); they allocate nothing and never synchronise.
 *   - there is no CPU fallback: without a CUDA device every compute entry point returns
 *     MEPOL_E_CUDA.
 */
#ifndef MEPOL_B200_H
#define MEPOL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MEPOL_OK 0
#define MEPOL_E_INVALID (-1)   /* bad argument (null pointer, k1 > n_ref, d out of range ...) */
#define MEPOL_E_WORKSPACE (-2) /* workspace too small */
#define MEPOL_E_CUDA (-3)      /* CUDA runtime error / no device */
#define MEPOL_E_UNSUPPORTED (-4)

#define MEPOL_KNN_AUTO 0 /* grid search for d <= 3, tensor-core pruner for 4 <= d <= 29 and k1 <= 12, else SIMT */
#define MEPOL_KNN_SIMT 1 /* fp32 SIMT pruner + fp64 re-rank (any d <= 64, k1 <= 60) */
#define MEPOL_KNN_TC 2   /* tcgen05 fp16 pruner (N=256 MMAs, fp16 accumulators) + fp64 re-rank (d <= 29, k1 <= 12) */
#define MEPOL_KNN_GRID 3 /* uniform-grid exact search in fp64 (d <= 3, k1 <= 60): the kd-tree regime */

#define MEPOL_MAX_DIM 64
#define MEPOL_MAX_K1 60
#define MEPOL_KNN_NSTATS 8

const char *mepol_last_error(void);
int mepol_version(void);
/* sm count / compute capability of the current device; MEPOL_E_CUDA when there is none. */
int mepol_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* Optional measurement hook used by bench.py: when enabled, every mepol_knn_search_f32 call records
 * a CUDA-event pair around its pruning kernel (the dominant kernel of the path) on the caller's
 * stream, up to 64 launches; mepol_profile_read waits for them and returns the per-launch
 * milliseconds.  No effect on results.  BENCH-ONLY: the event table is process-global and not thread-safe --
 * call it from one thread, around searches issued from that thread. */
int mepol_profile_enable(int on);
int mepol_profile_read(float *ms_out, int max_n, int *n_out);

/* ---------------------------------------------------------------------------------------------
 * Exact k-NN search.  Replaces src/algorithms/mepol.py:190-192
 *     nbrs = NearestNeighbors(n_neighbors=k+1, metric='euclidean', algorithm='auto', n_jobs=..)
 *     nbrs.fit(next_states); distances, indices = nbrs.kneighbors(next_states)
 * (arithmetic owned by scikit-learn: _argkmin.pyx.tp:471-510, _heap.pyx:46).
 *
 * ref   [n_ref, d] float32 row-major: the fitted particle set (device).
 * q     [n_q,  d] float32 row-major: query rows (device); may alias rows of ref (self is NOT
 *       removed, exactly like kneighbors(X) with X passed explicitly).
 * k1    = k + 1 neighbours per row.
 * idx_out  [n_q, k1] int32  -- neighbour indices into ref, ascending by (distance, index).
 * dist_out [n_q, k1] float64 -- Euclidean distances, sqrt of the exact float64 squared distance.
 * rk_out   [n_q] float64, optional (may be NULL) -- dist_out[:, k1-1] as a contiguous column, the
 *       only distance the objective consumes (mepol.py:150).
 * stats    optional DEVICE int64[MEPOL_KNN_NSTATS]: {candidates re-ranked, rows sent to the exact
 *       fallback, buffer compactions, rows whose fp16 band overflowed in the tensor-core pruner, 0...}.
 * Results are exact: candidates are pruned with a proven error band and re-ranked in float64;
 * rows whose candidate buffer overflows are recomputed by an exact float64 brute-force kernel.
 */
int mepol_knn_workspace_bytes(int64_t n_q, int64_t n_ref, int d, int k1, int algo, size_t *bytes);
int mepol_knn_search_f32(const float *ref, int64_t n_ref, const float *q, int64_t n_q, int d,
                         int k1, int algo, int32_t *idx_out, double *dist_out, double *rk_out,
                         void *workspace, size_t workspace_bytes, int64_t *stats, void *stream);

/* Test hook of the tensor-core pruner: writes the raw fp32 accumulator tile (scores
 * ||y'||^2 - 2 x~.y~ of the first 256 query rows against the first 128 references) to out
 * [256, 128] float32 (device).  Lets the tests check the tcgen05 contraction in isolation. */
int mepol_debug_tc_scores(const float *ref, int64_t n_ref, const float *q, int64_t n_q, int d,
                          int k1, float *out, void *workspace, size_t workspace_bytes,
                          void *stream);

/* The same hook for the fp16-accumulator pruner (knn_tc2.cu): first 256 query rows against the first
 * 256 references, out [256, 256] float32 (device). */
int mepol_debug_tc2_scores(const float *ref, int64_t n_ref, const float *q, int64_t n_q, int d,
                           int k1, float *out, void *workspace, size_t workspace_bytes,
                           void *stream);

/* Host-buffer convenience form of the same call (what a ctypes/cgo-style binding of the reference
 * would use): x [n, d] float32 HOST, self-search, outputs HOST int64 / float64 [n, k1] exactly as
 * sklearn returns them.  Allocates, copies H2D, searches, copies D2H, frees, synchronises. */
int mepol_knn_search_host_f32(const float *x, int64_t n, int d, int k1, int algo, int64_t *idx_out,
                              double *dist_out);

/* ---------------------------------------------------------------------------------------------
 * Reverse-neighbour CSR, built once per search and reused by every backward pass of the epoch.
 * Replaces the scatter-add (index_put_ accumulate) that autograd runs for
 * `importance_weights[indices[:, :-1]]` (mepol.py:148) with a deterministic gather.
 * idx [n, k1] int32 (device).  Only the first k = k1-1 columns are used (mepol.py:148 "[:, :-1]").
 * rowptr [n+1] int32, col [n*k] int32 (device, caller-owned): col[rowptr[j]..rowptr[j+1]) lists,
 * ascending, the rows i whose neighbour set contains j.
 */
int mepol_rev_csr_workspace_bytes(int64_t n, int k, size_t *bytes);
int mepol_rev_csr(const int32_t *idx, int64_t n, int k1, int32_t *rowptr, int32_t *col,
                  void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Importance-weighted k-NN entropy + KL, forward.  Replaces
 *   compute_importance_weights  mepol.py:114-139   (per-trajectory exp(cumsum(lt-lb)), normalise)
 *   compute_entropy             mepol.py:142-154
 *   compute_kl                  mepol.py:157-174   (value before the clamp; the caller clamps)
 * logp_t, logp_b [n]: log pi_target / log pi_behavioural of every particle, trajectory-major.
 * traj_off [n_traj+1] int32: CSR offsets of the trajectories inside [0, n).
 * idx [n, k1] int32, rk [n]: k-th neighbour distance (column k of the search result).
 * ns: dimension of the particle space, log_G = lgamma(ns/2+1), B = log k - digamma(k), eps: the
 *     reference's eps argument (0 for every shipped config but MountainCar's 1e-15).
 * Outputs (device): c [n] stabilised log-weights (cumsum), gW [n] dH/dW_i (may be NULL when no
 * backward is needed), scal double[8] = {H, KL_raw, s=sum_i gW_i*W_i, M=max c, S=sum exp(c-M),
 * n_nan_or_inf_terms, float64-exp range flag, 1.0 (the unit upstream gradient of the fused call below)}.
 * The _f32 form stores c/gW in float32 and evaluates per-row terms in fp32 with fp64 accumulation;
 * the _f64 form is float64 end to end (what the drop-in uses to reproduce the float64 reference).
 */
int mepol_iw_workspace_bytes(int64_t n, int64_t n_traj, size_t *bytes);
int mepol_iw_entropy_fwd_f32(const float *logp_t, const float *logp_b, const int32_t *traj_off,
                             int64_t n_traj, int64_t n, const int32_t *idx, int k1, const float *rk,
                             int ns, double log_G, double B, double eps, float *c, float *gW,
                             double *scal, void *workspace, size_t workspace_bytes, void *stream);
int mepol_iw_entropy_fwd_f64(const double *logp_t, const double *logp_b, const int32_t *traj_off,
                             int64_t n_traj, int64_t n, const int32_t *idx, int k1,
                             const double *rk, int ns, double log_G, double B, double eps,
                             double *c, double *gW, double *scal, void *workspace,
                             size_t workspace_bytes, void *stream);

/* Backward: grad_logp_t[j] = grad_H * dH/d logp_t[j]   (the hand-written replacement of
 * loss.backward() through mepol.py:121-152; d/d logp_b is the negative).  grad_H is a DEVICE
 * scalar (float64).  c, gW, scal come from the forward; rowptr/col from mepol_rev_csr. */
int mepol_iw_entropy_bwd_f32(const float *c, const float *gW, const double *scal,
                             const int32_t *traj_off, int64_t n_traj, int64_t n,
                             const int32_t *rowptr, const int32_t *col, const double *grad_H,
                             float *grad_logp_t, void *stream);
int mepol_iw_entropy_bwd_f64(const double *c, const double *gW, const double *scal,
                             const int32_t *traj_off, int64_t n_traj, int64_t n,
                             const int32_t *rowptr, const int32_t *col, const double *grad_H,
                             double *grad_logp_t, void *stream);

/* Forward + backward in one call (the backward for dH = 1; scale grad_unit by the upstream gradient):
 * what the autograd bridge issues when a gradient will be asked for (policy_update, mepol.py:268-281).
 * Same kernels and bit-identical results as mepol_iw_entropy_fwd_* followed by mepol_iw_entropy_bwd_*. */
int mepol_iw_entropy_fwd_bwd_f32(const float *logp_t, const float *logp_b, const int32_t *traj_off,
                                 int64_t n_traj, int64_t n, const int32_t *idx, int k1, const float *rk,
                                 int ns, double log_G, double B, double eps, const int32_t *rowptr,
                                 const int32_t *col, float *c, float *gW, double *scal, float *grad_unit,
                                 void *workspace, size_t workspace_bytes, void *stream);
int mepol_iw_entropy_fwd_bwd_f64(const double *logp_t, const double *logp_b, const int32_t *traj_off,
                                 int64_t n_traj, int64_t n, const int32_t *idx, int k1,
                                 const double *rk, int ns, double log_G, double B, double eps,
                                 const int32_t *rowptr, const int32_t *col, double *c, double *gW,
                                 double *scal, double *grad_unit, void *workspace,
                                 size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Diagonal-Gaussian log-probability of the policy head.  Replaces the elementwise tail of
 * GaussianPolicy.get_log_p (src/policy.py:43-51):
 *     log p = sum_j -0.5 * (log(2 pi) + 2 log_std_j + (a_j - mean_j)^2 / (exp(log_std_j) + 1e-7)^2)
 * mean [n, A] is the MLP output (policy.py:54, computed by the caller), actions [n, A], log_std [A] (all device,
 * row-major), logp [n].  The backward takes d loss / d logp [n] and writes d / d mean [n, A] and d / d log_std [A]
 * (reduced over the n states in a fixed order).  A <= 64.
 */
int mepol_gauss_logp_workspace_bytes(int64_t n, int A, size_t *bytes);
int mepol_gauss_logp_fwd_f32(const float *mean, const float *actions, const float *log_std, int64_t n, int A,
                             float *logp, void *stream);
int mepol_gauss_logp_fwd_f64(const double *mean, const double *actions, const double *log_std, int64_t n,
                             int A, double *logp, void *stream);
int mepol_gauss_logp_bwd_f32(const float *mean, const float *actions, const float *log_std,
                             const float *g_logp, int64_t n, int A, float *g_mean, float *g_log_std,
                             void *workspace, size_t workspace_bytes, void *stream);
int mepol_gauss_logp_bwd_f64(const double *mean, const double *actions, const double *log_std,
                             const double *g_logp, int64_t n, int A, double *g_mean, double *g_log_std,
                             void *workspace, size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MEPOL_B200_H */
