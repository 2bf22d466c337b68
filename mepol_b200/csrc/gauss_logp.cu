// Diagonal-Gaussian log-probability of the policy head and its backward, one kernel each.
//
// Replaces the elementwise tail of GaussianPolicy.get_log_p (reference src/policy.py:43-51):
//     log p(a | s) = sum_j -0.5 * ( log(2 pi) + 2 ls_j + (a_j - mu_j)^2 / (exp(ls_j) + eps)^2 ),  eps = 1e-7
// where mu = mean(net(s)) comes from the MLP (cuBLAS, outside this file) and ls = log_std is the
// state-independent parameter.  In torch that tail is ~8 elementwise launches forward and ~12 through
// autograd backward on [n, A] temporaries; here the forward reads mu and a once and writes log p [n],
// the backward reads them once more and writes d mu [n, A] plus d ls [A]:
//     d log p / d mu_j = (a_j - mu_j) / (exp(ls_j) + eps)^2
//     d log p / d ls_j = -1 + (a_j - mu_j)^2 exp(ls_j) / (exp(ls_j) + eps)^3
// d ls is a reduction over the n states: block partials in shared memory, written per block, and the last
// block to finish folds them in fixed order (deterministic, like the IW kernels).
// The output feeds the importance-weight scan (iw_entropy.cu) as logp_target / logp_behav.
#include <math.h>

#include "common.cuh"

namespace mepol {

constexpr int kLogpThreads = 256;
constexpr int kLogpMaxA = 64;
constexpr double kLog2Pi = 1.8378770664093454835606594728112;
constexpr double kPolicyEps = 1e-7;   // src/utils/dtypes.py:7

template <typename T>
__global__ void __launch_bounds__(kLogpThreads)
    gauss_logp_fwd_kernel(const T *__restrict__ mean, const T *__restrict__ act, const T *__restrict__ log_std, int64_t n,
                          int A, T *__restrict__ logp) {
    __shared__ double s_ls[kLogpMaxA], s_inv_var[kLogpMaxA];
    if (threadIdx.x < A) {
        const double ls = (double)log_std[threadIdx.x];
        const double sd = exp(ls) + kPolicyEps;
        s_ls[threadIdx.x] = ls;
        s_inv_var[threadIdx.x] = 1.0 / (sd * sd);
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * kLogpThreads + threadIdx.x;
    if (i >= n) return;
    const T *m = mean + i * A, *a = act + i * A;
    double acc = 0.0;
    for (int j = 0; j < A; ++j) {
        const double df = (double)a[j] - (double)m[j];
        acc += -0.5 * (kLog2Pi + 2.0 * s_ls[j] + df * df * s_inv_var[j]);
    }
    logp[i] = (T)acc;
}

template <typename T>
__global__ void __launch_bounds__(kLogpThreads)
    gauss_logp_bwd_kernel(const T *__restrict__ mean, const T *__restrict__ act, const T *__restrict__ log_std,
                          const T *__restrict__ g_logp, int64_t n, int A, T *__restrict__ g_mean,
                          double *__restrict__ part, unsigned *__restrict__ counter, T *__restrict__ g_log_std) {
    __shared__ double s_inv_var[kLogpMaxA], s_k3[kLogpMaxA];
    __shared__ double s_acc[kLogpThreads / 32][kLogpMaxA];
    __shared__ unsigned s_ticket;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < A) {
        const double e = exp((double)log_std[threadIdx.x]);
        const double sd = e + kPolicyEps;
        s_inv_var[threadIdx.x] = 1.0 / (sd * sd);
        s_k3[threadIdx.x] = e / (sd * sd * sd);
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * kLogpThreads + threadIdx.x;
    const double g = i < n ? (double)g_logp[i] : 0.0;
    for (int j = 0; j < A; ++j) {
        double gls = 0.0;
        if (i < n) {
            const double df = (double)act[i * A + j] - (double)mean[i * A + j];
            g_mean[i * A + j] = (T)(g * df * s_inv_var[j]);
            gls = g * (-1.0 + df * df * s_k3[j]);
        }
        gls = warp_sum(gls);
        if (lane == 0) s_acc[warp][j] = gls;
    }
    __syncthreads();
    if (threadIdx.x < A) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kLogpThreads / 32; ++w) t += s_acc[w][threadIdx.x];
        part[(size_t)blockIdx.x * A + threadIdx.x] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(counter, 1u);
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    __threadfence();
    // last block: fold the per-block partials in block order (fixed order -> reproducible bits)
    if (threadIdx.x < A) {
        double t = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) t += ((volatile double *)part)[(size_t)b * A + threadIdx.x];
        g_log_std[threadIdx.x] = (T)t;
    }
    if (threadIdx.x == 0) *counter = 0;   // ready for the next launch on this workspace
}

static int logp_blocks(int64_t n) { return (int)((n + kLogpThreads - 1) / kLogpThreads); }

template <typename T>
static int logp_fwd_impl(const T *mean, const T *act, const T *log_std, int64_t n, int A, T *logp, void *stream) {
    MEPOL_REQUIRE(mean && act && log_std && logp, MEPOL_E_INVALID, "gauss_logp_fwd: null pointer");
    MEPOL_REQUIRE(n > 0 && A >= 1 && A <= kLogpMaxA, MEPOL_E_INVALID, "gauss_logp_fwd: n=%lld A=%d (A <= %d)",
                  (long long)n, A, kLogpMaxA);
    gauss_logp_fwd_kernel<T><<<logp_blocks(n), kLogpThreads, 0, static_cast<cudaStream_t>(stream)>>>(mean, act, log_std, n,
                                                                                                    A, logp);
    MEPOL_CUDA_CHECK(cudaGetLastError());
    return MEPOL_OK;
}

template <typename T>
static int logp_bwd_impl(const T *mean, const T *act, const T *log_std, const T *g_logp, int64_t n, int A, T *g_mean,
                         T *g_log_std, void *workspace, size_t workspace_bytes, void *stream) {
    MEPOL_REQUIRE(mean && act && log_std && g_logp && g_mean && g_log_std && workspace, MEPOL_E_INVALID,
                  "gauss_logp_bwd: null pointer");
    MEPOL_REQUIRE(n > 0 && A >= 1 && A <= kLogpMaxA, MEPOL_E_INVALID, "gauss_logp_bwd: n=%lld A=%d (A <= %d)",
                  (long long)n, A, kLogpMaxA);
    const int nb = logp_blocks(n);
    Carver c(workspace);
    unsigned *counter = c.take<unsigned>(4);
    double *part = c.take<double>((size_t)nb * A);
    MEPOL_REQUIRE(workspace_bytes >= c.bytes(), MEPOL_E_WORKSPACE, "gauss_logp_bwd: workspace %zu < %zu", workspace_bytes,
                  c.bytes());
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MEPOL_CUDA_CHECK(cudaMemsetAsync(counter, 0, 4 * sizeof(unsigned), st));
    gauss_logp_bwd_kernel<T><<<nb, kLogpThreads, 0, st>>>(mean, act, log_std, g_logp, n, A, g_mean, part, counter, g_log_std);
    MEPOL_CUDA_CHECK(cudaGetLastError());
    return MEPOL_OK;
}

}  // namespace mepol

using namespace mepol;

extern "C" int mepol_gauss_logp_workspace_bytes(int64_t n, int A, size_t *bytes) {
    MEPOL_REQUIRE(bytes && n > 0 && A >= 1 && A <= kLogpMaxA, MEPOL_E_INVALID, "gauss_logp_workspace_bytes: bad arguments");
    Carver c(nullptr);
    c.take<unsigned>(4);
    c.take<double>((size_t)logp_blocks(n) * A);
    *bytes = c.bytes();
    return MEPOL_OK;
}
extern "C" int mepol_gauss_logp_fwd_f32(const float *mean, const float *actions, const float *log_std, int64_t n, int A,
                                        float *logp, void *stream) {
    return logp_fwd_impl<float>(mean, actions, log_std, n, A, logp, stream);
}
extern "C" int mepol_gauss_logp_fwd_f64(const double *mean, const double *actions, const double *log_std, int64_t n, int A,
                                        double *logp, void *stream) {
    return logp_fwd_impl<double>(mean, actions, log_std, n, A, logp, stream);
}
extern "C" int mepol_gauss_logp_bwd_f32(const float *mean, const float *actions, const float *log_std, const float *g_logp,
                                        int64_t n, int A, float *g_mean, float *g_log_std, void *workspace,
                                        size_t workspace_bytes, void *stream) {
    return logp_bwd_impl<float>(mean, actions, log_std, g_logp, n, A, g_mean, g_log_std, workspace, workspace_bytes, stream);
}
extern "C" int mepol_gauss_logp_bwd_f64(const double *mean, const double *actions, const double *log_std,
                                        const double *g_logp, int64_t n, int A, double *g_mean, double *g_log_std,
                                        void *workspace, size_t workspace_bytes, void *stream) {
    return logp_bwd_impl<double>(mean, actions, log_std, g_logp, n, A, g_mean, g_log_std, workspace, workspace_bytes, stream);
}
