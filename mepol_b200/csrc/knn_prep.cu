// K0 "prep": column means of the fitted set, then per-row centring / packing of the operands both
// pruners consume.  Replaces the per-chunk upcast + row-norm work sklearn does on the fly
// (sklearn/metrics/_pairwise_distances_reduction/_base.pyx.tp:46-84 row norms,
//  _middle_term_computer.pyx.tp:329-396 float32->float64 upcast copies).
//
// All per-row arithmetic is float64 (B200 has full-rate-enough fp64 and the pass is a single
// streaming read of X): x' = sigma * (x - mean) exactly (sigma = 2^e keeps fp16 in range), ||x'||^2 exactly, then
//   SIMT operand   xs[row][0..dpad)  = fl32(x')               (zero padded)
//   r_nrm[row]     = fl32(||x'||^2), r_anorm[row] >= ||x'||   (rounded up)
//   TC operand     xb[row][0..32)    = fp16 row (see knn_tc.cu for the column meaning)
//   q_delta[row]   >= ||fp16(x') - x'||  (rounded up), and global maxima via atomicMax on the
//   (non-negative) float bit patterns.
#include <cuda_fp16.h>

#include "knn_internal.cuh"

namespace mepol {

constexpr int kMeanThreadsX = 64;
constexpr int kMeanThreadsY = 4;
constexpr int kMeanRowsPerBlock = 1024;

__global__ void __launch_bounds__(kMeanThreadsX *kMeanThreadsY)
    col_sum_partial_kernel(const float *__restrict__ x, int64_t n, int d, double *__restrict__ partial) {
    __shared__ double sm[kMeanThreadsY][kMeanThreadsX];
    __shared__ float smin[kMeanThreadsY][kMeanThreadsX], smax[kMeanThreadsY][kMeanThreadsX];
    const int cx = threadIdx.x, ry = threadIdx.y;
    const int64_t row0 = (int64_t)blockIdx.x * kMeanRowsPerBlock;
    int64_t row1 = row0 + kMeanRowsPerBlock;
    if (row1 > n) row1 = n;
    double s = 0.0;
    float lo = __int_as_float(0x7f800000), hi = -lo;
    if (cx < d)
        for (int64_t r = row0 + ry; r < row1; r += kMeanThreadsY) {
            const float v = x[r * d + cx];
            s += (double)v;
            lo = fminf(lo, v);
            hi = fmaxf(hi, v);
        }
    sm[ry][cx] = s;
    smin[ry][cx] = lo;
    smax[ry][cx] = hi;
    __syncthreads();
    if (ry == 0 && cx < d) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < kMeanThreadsY; ++i) {
            t += sm[i][cx];
            lo = fminf(lo, smin[i][cx]);
            hi = fmaxf(hi, smax[i][cx]);
        }
        double *o = partial + (int64_t)blockIdx.x * (3 * MEPOL_MAX_DIM);
        o[cx] = t;
        o[MEPOL_MAX_DIM + cx] = (double)lo;
        o[2 * MEPOL_MAX_DIM + cx] = (double)hi;
    }
}

constexpr int kMeanFinalY = 16;

__global__ void __launch_bounds__(64 * kMeanFinalY)
    col_mean_final_kernel(const double *__restrict__ partial, int n_blocks, int64_t n, int d,
                          double *__restrict__ mean, float *__restrict__ glob) {
    __shared__ double r2[64];
    __shared__ double ssum[kMeanFinalY][64], slo[kMeanFinalY][64], shi[kMeanFinalY][64];
    const int c = threadIdx.x, y = threadIdx.y;
    {   // kMeanFinalY interleaved chains per column, combined in a fixed order (deterministic)
        double t = 0.0, lo = 1e300, hi = -1e300;
        if (c < d)
            for (int b = y; b < n_blocks; b += kMeanFinalY) {
                const double *pb = partial + (int64_t)b * (3 * MEPOL_MAX_DIM);
                t += pb[c];
                lo = fmin(lo, pb[MEPOL_MAX_DIM + c]);
                hi = fmax(hi, pb[2 * MEPOL_MAX_DIM + c]);
            }
        ssum[y][c] = t;
        slo[y][c] = lo;
        shi[y][c] = hi;
    }
    __syncthreads();
    if (y != 0) return;
    double ext = 0.0;
    if (c < d) {
        double t = 0.0, lo = 1e300, hi = -1e300;
        for (int i = 0; i < kMeanFinalY; ++i) {
            t += ssum[i][c];
            lo = fmin(lo, slo[i][c]);
            hi = fmax(hi, shi[i][c]);
        }
        // The mean only has to be the SAME value for queries and references; rounding it to
        // float keeps x - mean exactly representable in float64 for every float32 input.
        const double m = (double)(float)(t / (double)n);
        mean[c] = m;
        ext = fmax(hi - m, m - lo);
    }
    r2[c] = ext * ext;
    __syncwarp();
    asm volatile("bar.sync 1, 64;" ::: "memory");      // the 64 threads of row y == 0 (two warps)
    if (c == 0) {
        // Power-of-two scale sigma with sigma * max_row ||x - mean|| <= kTcScaledNormMax, from the
        // per-column extents (an upper bound of the max norm).  Scaling by 2^e is exact, so the
        // fp32 operands lose nothing; the fp16 tensor-core operands stay inside fp16's range
        // (entries -2*y~ <= 2*128, ||y'||^2 <= 16384).
        double a2 = 0.0;
        for (int i = 0; i < d; ++i) a2 += r2[i];
        int e = 0;
        const double ab = sqrt(a2);
        if (ab > 0.0 && ab < 1e300) {
            e = (int)floor(log2(kTcScaledNormMax / ab));
            while (ldexp(ab, e) > kTcScaledNormMax) --e;
            e = max(-100, min(100, e));
        }
        mean[MEPOL_MAX_DIM] = ldexp(1.0, e);
    }
    if (c < G_COUNT) glob[c] = 0.0f;
}

int launch_col_mean(const float *ref, int64_t n_ref, int d, const KnnPlan &p, cudaStream_t st) {
    dim3 blk(kMeanThreadsX, kMeanThreadsY);
    col_sum_partial_kernel<<<p.mean_blocks, blk, 0, st>>>(ref, n_ref, d, p.mean_partial);
    col_mean_final_kernel<<<1, dim3(64, kMeanFinalY), 0, st>>>(p.mean_partial, p.mean_blocks, n_ref, d, p.mean, p.glob);
    MEPOL_CUDA_CHECK(cudaGetLastError());
    return MEPOL_OK;
}

int mean_blocks_for(int64_t n_ref) { return (int)((n_ref + kMeanRowsPerBlock - 1) / kMeanRowsPerBlock); }

// One thread per row, one CTA per 128 rows (= one tensor-core operand tile).  dpad is the SIMT operand
// width; when xb != nullptr the fp16 tensor-core operand is produced as well (refs: [-2*y~ (d cols),
// n_hi, n_mid, n_lo, 0...]; queries: [x~ (d cols), 1, 1, 1, 0...]) -- see knn_tc2.cu.
// Global traffic is staged through shared memory: the 128 x d input block is read as one contiguous
// range, each thread packs its row out of / into shared memory, and the SIMT rows (128 x dpad floats,
// contiguous) and the operand tile (8 KB, contiguous in the UMMA canonical order) leave as coalesced
// stores.  The thread-per-row global accesses of round 1 (116-byte strided reads, 2-byte scattered
// stores) made this kernel 8x slower than its bytes: 199 -> see profiles/r02 launch list.
constexpr int kPackRows = 128;
__host__ __device__ inline int tc_tile_index(int r, int c) { return ((r >> 3) * 4 + (c >> 3)) * 64 + (r & 7) * 8 + (c & 7); }

__global__ void __launch_bounds__(kPackRows)
    pack_rows_kernel(const float *__restrict__ x, int64_t n, int64_t n_pad, int d, int dpad,
                     const double *__restrict__ mean, float *__restrict__ xs, float *__restrict__ nrm,
                     float *__restrict__ anorm, uint16_t *__restrict__ xb, float *__restrict__ delta,
                     float *__restrict__ ahalf, float *__restrict__ glob, int is_ref,
                     const unsigned long long *__restrict__ gate, unsigned long long gate_min) {
    // lazy launch (fp32 operands of the rescue pass behind the tensor-core pruner): nothing to do unless the pruner
    // flagged more rows than the exact fallback takes (the same test knn_prune_simt_kernel makes in rescue mode)
    if (gate && *gate <= gate_min) return;
    extern __shared__ __align__(16) unsigned char pack_smem[];
    __shared__ double smean[MEPOL_MAX_DIM];
    float *s_in = reinterpret_cast<float *>(pack_smem);                         // [128][d]
    float *s_xs = s_in + kPackRows * d;                                         // [128][dpad + 1] (padded: no bank conflicts)
    uint16_t *s_xb = reinterpret_cast<uint16_t *>(s_xs + kPackRows * (dpad + 1));   // [4096] one operand tile
    const int tid = threadIdx.x;
    if (tid < d) smean[tid] = mean[tid];
    const double sigma = mean[MEPOL_MAX_DIM];
    const int64_t row0 = (int64_t)blockIdx.x * kPackRows;
    const int nrows = row0 < n ? (int)((n - row0) < kPackRows ? (n - row0) : kPackRows) : 0;
    {   // contiguous block read (16-byte vectors when the block starts on a 16-byte boundary)
        const int total = nrows * d;
        const float *src = x + row0 * d;
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {    // a query slice of the fitted matrix may start anywhere
            const float4 *src4 = reinterpret_cast<const float4 *>(src);
            float4 *dst4 = reinterpret_cast<float4 *>(s_in);
            for (int i = tid; i < (total >> 2); i += kPackRows) dst4[i] = src4[i];
            for (int i = (total & ~3) + tid; i < total; i += kPackRows) s_in[i] = src[i];
        } else {
            for (int i = tid; i < total; i += kPackRows) s_in[i] = src[i];
        }
    }
    __syncthreads();
    const int64_t row = row0 + tid;
    const int xs_stride = dpad + 1;
    float a_up = 0.f, delta_up = 0.f, atilde_up = 0.f, ahalf_up = 0.f;
    if (row < n) {
        const float *xr = s_in + tid * d;
        float *o = s_xs + tid * xs_stride;
        double n2 = 0.0, dl2 = 0.0, nt2 = 0.0, nh2 = 0.0;
        bool in_range = true;
        for (int c = 0; c < d; ++c) {
            const double v = ((double)xr[c] - smean[c]) * sigma;   // exact: sigma is a power of two
            n2 += v * v;
            o[c] = (float)v;
            in_range = in_range && fabs(v) <= kTcQueryAbsMax;
        }
        for (int c = d; c < dpad; ++c) o[c] = 0.f;
        if (nrm) nrm[row] = (float)n2;
        a_up = __double2float_ru(sqrt(n2) * (1.0 + 1e-12));
        anorm[row] = a_up;
        if (xb) {
            // fp16 operand.  A query far outside the fitted set's extent (fp16 would overflow) gets a
            // zero operand and an infinite band: the kernel flags it and the exact fallback serves it.
            for (int c = 0; c < d; ++c) {
                const double v = in_range ? ((double)xr[c] - smean[c]) * sigma : 0.0;
                const __half hb = __double2half(v);
                const double hv = (double)__half2float(hb);
                dl2 += (hv - v) * (hv - v);
                nt2 += hv * hv;
                if (c < 16) nh2 += hv * hv;
                const __half st = is_ref ? __float2half(-2.0f * __half2float(hb)) : hb;   // x2 exact
                s_xb[tc_tile_index(tid, c)] = __half_as_ushort(st);
            }
            delta_up = in_range ? __double2float_ru(sqrt(dl2) * (1.0 + 1e-12)) : __int_as_float(0x7f800000);
            atilde_up = __double2float_ru(sqrt(nt2) * (1.0 + 1e-12));
            ahalf_up = __double2float_ru(sqrt(nh2) * (1.0 + 1e-12));
            if (delta) delta[row] = delta_up;
            if (ahalf) ahalf[row] = ahalf_up;
            if (is_ref) {
                // ||y'||^2 split over three fp16 columns: hi + mid + lo reproduces it to ~2^-24
                const __half h = __double2half(n2);
                const double r1 = n2 - (double)__half2float(h);
                const __half m = __double2half(r1);
                const double r2 = r1 - (double)__half2float(m);
                const __half l = __double2half(r2);
                s_xb[tc_tile_index(tid, d)] = __half_as_ushort(h);
                s_xb[tc_tile_index(tid, d + 1)] = __half_as_ushort(m);
                s_xb[tc_tile_index(tid, d + 2)] = __half_as_ushort(l);
            } else {
                const uint16_t one = __half_as_ushort(__float2half(1.0f));
                s_xb[tc_tile_index(tid, d)] = one;
                s_xb[tc_tile_index(tid, d + 1)] = one;
                s_xb[tc_tile_index(tid, d + 2)] = one;
            }
            for (int c = d + 3; c < kTcKPad; ++c) s_xb[tc_tile_index(tid, c)] = 0;
        }
    } else if (row < n_pad && xb) {
        // padding rows of the tensor-core operand: references get a norm above every real score
        // (real scores stay below 3 * kTcScaledNormMax^2 = 49152) so they never win
        for (int c = 0; c < kTcKPad; ++c) s_xb[tc_tile_index(tid, c)] = 0;
        if (is_ref) s_xb[tc_tile_index(tid, d)] = __half_as_ushort(__float2half(60000.0f));
    }
    __syncthreads();
    if (xs) {   // rows row0 .. row0 + nrows of the SIMT operand are one contiguous range
        const int total = nrows * dpad;
        float *dst = xs + row0 * dpad;
        for (int i = tid; i < total; i += kPackRows) {
            const int r = i / dpad, c = i - r * dpad;
            dst[i] = s_xs[r * xs_stride + c];
        }
    }
    if (xb && row0 < n_pad) {   // the whole 8 KB tile (n_pad is a multiple of 128)
        const uint4 *src4 = reinterpret_cast<const uint4 *>(s_xb);
        uint4 *dst4 = reinterpret_cast<uint4 *>(xb + (size_t)blockIdx.x * 4096);
        for (int i = tid; i < 512; i += kPackRows) dst4[i] = src4[i];
    }
    if (is_ref) {
        // max over the fitted set (bit patterns of non-negative floats order like unsigned ints)
        float a = warp_max(a_up), dm = warp_max(delta_up), at = warp_max(atilde_up), ah = warp_max(ahalf_up);
        if ((threadIdx.x & 31) == 0) {
            atomicMax(reinterpret_cast<unsigned *>(glob + G_AMAX), __float_as_uint(a));
            if (xb) {
                atomicMax(reinterpret_cast<unsigned *>(glob + G_DELTA_MAX), __float_as_uint(dm));
                atomicMax(reinterpret_cast<unsigned *>(glob + G_ATILDE_MAX), __float_as_uint(at));
                atomicMax(reinterpret_cast<unsigned *>(glob + G_AHALF_MAX), __float_as_uint(ah));
            }
        }
    }
}

// lazy = false: the operands of the search itself (SIMT: fp32 rows + norms; tensor-core: fp16 operand tiles + norms).
// lazy = true (tensor-core searches only): the fp32 rows + norms the RESCUE pass reads, produced behind the pruning
//   kernel and only when it flagged enough rows (device-side gate) -- on the usual search that is 128 B per row of
//   writes that nobody reads (134 MB at N = 1M, on every rank of a sharded search).
int launch_pack(const float *x, int64_t n, int d, const KnnPlan &p, bool is_ref, bool lazy, cudaStream_t st) {
    const bool tc = (p.algo == MEPOL_KNN_TC) && !lazy;
    int64_t n_pad = is_ref ? p.n_ref_pad : p.n_q_pad;
    if (!tc) n_pad = n;
    const int64_t rows = n_pad > n ? n_pad : n;
    const int blocks = (int)((rows + kPackRows - 1) / kPackRows);
    const size_t smem = (size_t)kPackRows * d * 4 + (size_t)kPackRows * (p.dpad + 1) * 4 + 8192;
    if (smem > 48 * 1024)
        MEPOL_CUDA_CHECK(cudaFuncSetAttribute(pack_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // fp32 operands are produced in both modes (SIMT search / rescue pass); query rows that alias
    // the fitted matrix reuse the reference rows
    float *xs = is_ref ? p.rs : (p.q_is_ref ? nullptr : p.qs);
    float *nrm = is_ref ? p.r_nrm : nullptr;
    if (tc) xs = nrm = nullptr;      // produced lazily, see above
    pack_rows_kernel<<<blocks, kPackRows, smem, st>>>(
        x, n, n_pad, d, p.dpad, p.mean, xs, nrm,
        is_ref ? p.r_anorm : p.q_anorm, tc ? (is_ref ? p.rb : p.qb) : nullptr,
        tc ? (is_ref ? nullptr : p.q_delta) : nullptr, tc ? (is_ref ? nullptr : p.q_ahalf) : nullptr, p.glob,
        is_ref ? 1 : 0, lazy ? p.stats + 3 : nullptr, lazy ? rescue_min_rows(p) : 0ULL);
    MEPOL_CUDA_CHECK(cudaGetLastError());
    return MEPOL_OK;
}

}  // namespace mepol
