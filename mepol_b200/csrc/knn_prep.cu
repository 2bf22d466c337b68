// K0 "prep": column means of the fitted set, then per-row centring / packing of the operands both
// pruners consume.  Replaces the per-chunk upcast + row-norm work sklearn does on the fly
// (sklearn/metrics/_pairwise_distances_reduction/_base.pyx.tp:46-84 row norms,
//  _middle_term_computer.pyx.tp:329-396 float32->float64 upcast copies).
//
// All per-row arithmetic is float64 (B200 has full-rate-enough fp64 and the pass is a single
// streaming read of X): x' = x - mean exactly, ||x'||^2 exactly, then
//   SIMT operand   xs[row][0..dpad)  = fl32(x')               (zero padded)
//   r_nrm[row]     = fl32(||x'||^2), r_anorm[row] >= ||x'||   (rounded up)
//   TC operand     xb[row][0..32)    = bf16 row (see knn_tc.cu for the column meaning)
//   q_delta[row]   >= ||bf16(x') - x'||  (rounded up), and global maxima via atomicMax on the
//   (non-negative) float bit patterns.
#include <cuda_bf16.h>

#include "knn_internal.cuh"

namespace mepol {

constexpr int kMeanThreadsX = 64;
constexpr int kMeanThreadsY = 4;
constexpr int kMeanRowsPerBlock = 1024;

__global__ void __launch_bounds__(kMeanThreadsX *kMeanThreadsY)
    col_sum_partial_kernel(const float *__restrict__ x, int64_t n, int d, double *__restrict__ partial) {
    __shared__ double sm[kMeanThreadsY][kMeanThreadsX];
    const int cx = threadIdx.x, ry = threadIdx.y;
    const int64_t row0 = (int64_t)blockIdx.x * kMeanRowsPerBlock;
    int64_t row1 = row0 + kMeanRowsPerBlock;
    if (row1 > n) row1 = n;
    double s = 0.0;
    if (cx < d)
        for (int64_t r = row0 + ry; r < row1; r += kMeanThreadsY) s += (double)x[r * d + cx];
    sm[ry][cx] = s;
    __syncthreads();
    if (ry == 0 && cx < d) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < kMeanThreadsY; ++i) t += sm[i][cx];
        partial[(int64_t)blockIdx.x * MEPOL_MAX_DIM + cx] = t;
    }
}

__global__ void col_mean_final_kernel(const double *__restrict__ partial, int n_blocks, int64_t n, int d,
                                      double *__restrict__ mean, float *__restrict__ glob) {
    const int c = threadIdx.x;
    if (c < d) {
        double t = 0.0;
        for (int b = 0; b < n_blocks; ++b) t += partial[(int64_t)b * MEPOL_MAX_DIM + c];  // fixed order
        // The mean only has to be the SAME value for queries and references; rounding it to
        // float keeps x - mean exactly representable in float64 for every float32 input.
        mean[c] = (double)(float)(t / (double)n);
    }
    if (c < G_COUNT) glob[c] = 0.0f;
}

int launch_col_mean(const float *ref, int64_t n_ref, int d, const KnnPlan &p, cudaStream_t st) {
    dim3 blk(kMeanThreadsX, kMeanThreadsY);
    col_sum_partial_kernel<<<p.mean_blocks, blk, 0, st>>>(ref, n_ref, d, p.mean_partial);
    col_mean_final_kernel<<<1, 64, 0, st>>>(p.mean_partial, p.mean_blocks, n_ref, d, p.mean, p.glob);
    MEPOL_CUDA_CHECK(cudaGetLastError());
    return MEPOL_OK;
}

int mean_blocks_for(int64_t n_ref) { return (int)((n_ref + kMeanRowsPerBlock - 1) / kMeanRowsPerBlock); }

// One thread per row.  dpad is the SIMT operand width; when xb != nullptr the bf16 tensor-core
// operand is produced as well (refs: [-2*y~ (d cols), n_hi, n_mid, n_lo, 0...]; queries:
// [x~ (d cols), 1, 1, 1, 0...]) -- see knn_tc.cu.
__global__ void __launch_bounds__(128)
    pack_rows_kernel(const float *__restrict__ x, int64_t n, int64_t n_pad, int d, int dpad,
                     const double *__restrict__ mean, float *__restrict__ xs, float *__restrict__ nrm,
                     float *__restrict__ anorm, uint16_t *__restrict__ xb, float *__restrict__ delta,
                     float *__restrict__ glob, int is_ref) {
    __shared__ double smean[MEPOL_MAX_DIM];
    if (threadIdx.x < d) smean[threadIdx.x] = mean[threadIdx.x];
    __syncthreads();
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float a_up = 0.f, delta_up = 0.f, atilde_up = 0.f;
    if (row < n) {
        const float *xr = x + row * d;
        float *o = xs ? xs + row * dpad : nullptr;
        double n2 = 0.0, dl2 = 0.0, nt2 = 0.0;
        for (int c = 0; c < d; ++c) {
            double v = (double)xr[c] - smean[c];
            n2 += v * v;
            if (o) o[c] = (float)v;
            if (xb) {
                __nv_bfloat16 hb = __double2bfloat16(v);
                double hv = (double)__bfloat162float(hb);
                dl2 += (hv - v) * (hv - v);
                nt2 += hv * hv;
                __nv_bfloat16 st = is_ref ? __float2bfloat16(-2.0f * __bfloat162float(hb)) : hb;  // x2 exact
                xb[tc_elem_index(row, c)] = __bfloat16_as_ushort(st);
            }
        }
        if (o)
            for (int c = d; c < dpad; ++c) o[c] = 0.f;
        if (nrm) nrm[row] = (float)n2;
        a_up = __double2float_ru(sqrt(n2) * (1.0 + 1e-12));
        anorm[row] = a_up;
        if (xb) {
            delta_up = __double2float_ru(sqrt(dl2) * (1.0 + 1e-12));
            atilde_up = __double2float_ru(sqrt(nt2) * (1.0 + 1e-12));
            if (delta) delta[row] = delta_up;
            if (is_ref) {
                // ||y'||^2 split over three bf16 columns: hi + mid + lo reproduces it to ~2^-24
                __nv_bfloat16 h = __double2bfloat16(n2);
                double r1 = n2 - (double)__bfloat162float(h);
                __nv_bfloat16 m = __double2bfloat16(r1);
                double r2 = r1 - (double)__bfloat162float(m);
                __nv_bfloat16 l = __double2bfloat16(r2);
                xb[tc_elem_index(row, d)] = __bfloat16_as_ushort(h);
                xb[tc_elem_index(row, d + 1)] = __bfloat16_as_ushort(m);
                xb[tc_elem_index(row, d + 2)] = __bfloat16_as_ushort(l);
            } else {
                const uint16_t one = __bfloat16_as_ushort(__float2bfloat16(1.0f));
                xb[tc_elem_index(row, d)] = one;
                xb[tc_elem_index(row, d + 1)] = one;
                xb[tc_elem_index(row, d + 2)] = one;
            }
            for (int c = d + 3; c < kTcKPad; ++c) xb[tc_elem_index(row, c)] = 0;
        }
    } else if (row < n_pad && xb) {
        // padding rows of the tensor-core operand: references get a huge norm so they never win
        for (int c = 0; c < kTcKPad; ++c) xb[tc_elem_index(row, c)] = 0;
        if (is_ref) xb[tc_elem_index(row, d)] = __bfloat16_as_ushort(__float2bfloat16(3.0e38f));
    }
    if (is_ref) {
        // max over the fitted set (bit patterns of non-negative floats order like unsigned ints)
        float a = warp_max(a_up), dm = warp_max(delta_up), at = warp_max(atilde_up);
        if ((threadIdx.x & 31) == 0) {
            atomicMax(reinterpret_cast<unsigned *>(glob + G_AMAX), __float_as_uint(a));
            if (xb) {
                atomicMax(reinterpret_cast<unsigned *>(glob + G_DELTA_MAX), __float_as_uint(dm));
                atomicMax(reinterpret_cast<unsigned *>(glob + G_ATILDE_MAX), __float_as_uint(at));
            }
        }
    }
}

int launch_pack(const float *x, int64_t n, int d, const KnnPlan &p, bool is_ref, cudaStream_t st) {
    const bool tc = (p.algo == MEPOL_KNN_TC);
    int64_t n_pad = is_ref ? p.n_ref_pad : p.n_q_pad;
    if (!tc) n_pad = n;
    const int64_t rows = n_pad > n ? n_pad : n;
    const int blocks = (int)((rows + 127) / 128);
    // fp32 operands are produced in both modes (SIMT search / rescue pass); query rows that alias
    // the fitted matrix reuse the reference rows
    float *xs = is_ref ? p.rs : (p.q_is_ref ? nullptr : p.qs);
    pack_rows_kernel<<<blocks, 128, 0, st>>>(
        x, n, n_pad, d, p.dpad, p.mean, xs, is_ref ? p.r_nrm : nullptr,
        is_ref ? p.r_anorm : p.q_anorm, tc ? (is_ref ? p.rb : p.qb) : nullptr,
        tc ? (is_ref ? nullptr : p.q_delta) : nullptr, p.glob, is_ref ? 1 : 0);
    MEPOL_CUDA_CHECK(cudaGetLastError());
    return MEPOL_OK;
}

}  // namespace mepol
