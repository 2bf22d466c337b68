// Exact k-NN for tiny feature dimensions (d <= 3) on a uniform grid -- the replacement of the
// reference's kd-tree regime (sklearn picks KDTree for d <= 15, neighbors/_base.py:615-626; the
// GridWorld / MountainCar configurations of the reference have ns = 2, scripts/tae/grid_world.sh,
// mountain_car.sh) where a dense O(N^2) scan -- tensor cores or not -- is the wrong algorithm.
//
//   G0  bounding box of the fitted set (from the column min/max of knn_prep's first pass) -> cell size
//   G1  cell id per fitted point + histogram            (atomicAdd on a per-cell counter)
//   G2  exclusive scan of the histogram                 (cell -> first slot)
//   G3  scatter: points sorted by cell, 16 B each (x, y, z, original index)
//   G4  one warp per query: visit the query's cell, then Chebyshev rings of cells around it, 32
//       candidates at a time; exact float64 d2 (the same fma chain as knn_rerank.cu, so distances are
//       bit-identical to the other paths); the k1 best are kept sorted across the warp's registers
//       (2 entries per lane) under the key (d2, index) -- the reference's tie rule
//       (sklearn/utils/_heap.pyx:46); stop once the k1-th best d2 is strictly below the squared
//       distance from the query to the nearest face of the visited block (no unvisited point can enter).
//
// The order of points inside a cell comes from atomics and is not reproducible; the output is,
// because the kept set and its order only depend on (d2, index).  Rows that cannot fill their list
// (non-finite input) are flagged for the exact brute-force fallback of knn_rerank.cu.
#include "knn_internal.cuh"

namespace mepol {

constexpr int kGridWarps = 8;
constexpr int kGridScanThreads = 1024;

struct GridPar {
    double lo[3], h[3];     // box origin and cell edge per dimension
    float lo_f[3], inv_h[3];
};

__device__ __forceinline__ int grid_coord(float x, float lo, float inv_h, int g) {
    const float t = floorf((x - lo) * inv_h);
    int c = (t == t) ? (int)fminf(fmaxf(t, 0.f), (float)(g - 1)) : 0;   // NaN -> cell 0 (the row is flagged later)
    return c;
}

__global__ void grid_setup_kernel(const double *__restrict__ partial, int n_blocks, int d, int g, GridPar *par) {
    const int c = threadIdx.x;
    if (c >= 3) return;
    double lo = 0.0, hi = 0.0;
    if (c < d) {
        lo = 1e300;
        hi = -1e300;
        for (int b = 0; b < n_blocks; ++b) {
            const double *pb = partial + (int64_t)b * (3 * MEPOL_MAX_DIM);
            lo = fmin(lo, pb[MEPOL_MAX_DIM + c]);
            hi = fmax(hi, pb[2 * MEPOL_MAX_DIM + c]);
        }
        if (!(lo <= hi) || !(hi - lo < 1e300)) {   // empty / non-finite column: one degenerate cell row
            lo = 0.0;
            hi = 0.0;
        }
    }
    double h = (hi - lo) / (double)g;
    if (!(h > 0.0)) h = 1.0;
    par->lo[c] = lo;
    par->h[c] = h;
    par->lo_f[c] = (float)lo;       // lo is a float32 value (a column minimum): exact
    par->inv_h[c] = (float)(1.0 / h);
}

template <int D>
__device__ __forceinline__ int grid_cell_of(const float *__restrict__ x, const GridPar &par, int g) {
    int cell = 0;
#pragma unroll
    for (int c = D - 1; c >= 0; --c) cell = cell * g + grid_coord(x[c], par.lo_f[c], par.inv_h[c], g);
    return cell;
}

template <int D>
__global__ void grid_count_kernel(const float *__restrict__ ref, int n_ref, const GridPar *__restrict__ par, int g,
                                  int32_t *__restrict__ cell_of, int32_t *__restrict__ cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_ref) return;
    const GridPar p = *par;
    const int cell = grid_cell_of<D>(ref + (size_t)i * D, p, g);
    cell_of[i] = cell;
    atomicAdd(cnt + cell, 1);
}

// exclusive scan of cnt[0..n) into start[0..n]; cnt becomes the per-cell write cursor (zeroed)
__global__ void __launch_bounds__(kGridScanThreads)
    grid_scan_kernel(int32_t *__restrict__ cnt, int32_t *__restrict__ start, int n) {
    __shared__ int s_sum[kGridScanThreads];
    const int t = threadIdx.x;
    const int per = (n + kGridScanThreads - 1) / kGridScanThreads;
    const int a = min(n, t * per), b = min(n, a + per);
    int s = 0;
    for (int i = a; i < b; ++i) s += cnt[i];
    s_sum[t] = s;
    __syncthreads();
    for (int o = 1; o < kGridScanThreads; o <<= 1) {
        const int v = t >= o ? s_sum[t - o] : 0;
        __syncthreads();
        s_sum[t] += v;
        __syncthreads();
    }
    int run = s_sum[t] - s;
    for (int i = a; i < b; ++i) {
        const int c = cnt[i];
        start[i] = run;
        cnt[i] = 0;
        run += c;
    }
    if (t == kGridScanThreads - 1) start[n] = s_sum[t];
}

template <int D>
__global__ void grid_scatter_kernel(const float *__restrict__ ref, int n_ref, const int32_t *__restrict__ cell_of,
                                    const int32_t *__restrict__ start, int32_t *__restrict__ cursor,
                                    float4 *__restrict__ pts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_ref) return;
    const int cell = cell_of[i];
    const int pos = start[cell] + atomicAdd(cursor + cell, 1);
    const float *x = ref + (size_t)i * D;
    pts[pos] = make_float4(x[0], D > 1 ? x[1] : 0.f, D > 2 ? x[2] : 0.f, __int_as_float(i));
}

__device__ __forceinline__ bool grid_key_less(double da, int ia, double db, int ib) {
    return da < db || (da == db && ia < ib);
}
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double shfl_up_d(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }

// kSlots = 1: k1 <= 32, kSlots = 2: k1 <= 64.  sorted_q: the queries ARE the fitted points and are
// walked in cell order (neighbouring warps touch the same cells).
template <int D, int kSlots>
__global__ void __launch_bounds__(kGridWarps * 32)
    grid_query_kernel(const float *__restrict__ q, int n_q, int sorted_q, const float4 *__restrict__ pts,
                      const int32_t *__restrict__ start, const GridPar *__restrict__ par, int g, int k1,
                      int32_t *__restrict__ idx_out, double *__restrict__ dist_out, double *__restrict__ rk_out,
                      unsigned char *__restrict__ flag, unsigned long long *__restrict__ stats) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * kGridWarps + warp;
    if (slot >= n_q) return;
    const GridPar p = *par;
    int row = slot;
    float qx[3] = {0.f, 0.f, 0.f};
    if (sorted_q) {
        const float4 me = pts[slot];
        row = __float_as_int(me.w);
        qx[0] = me.x;
        qx[1] = me.y;
        qx[2] = me.z;
    } else {
#pragma unroll
        for (int c = 0; c < D; ++c) qx[c] = q[(size_t)row * D + c];
    }
    int cq[3] = {0, 0, 0};
#pragma unroll
    for (int c = 0; c < D; ++c) cq[c] = grid_coord(qx[c], p.lo_f[c], p.inv_h[c], g);

    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    // sorted list: position l in (d0, i0) of lane l, position 32 + l in (d1, i1)
    double d0 = INF, d1 = INF;
    int i0 = 0x7fffffff, i1 = 0x7fffffff;
    double kth_d = INF;
    int kth_i = 0x7fffffff;
    const int kth_lane = (k1 - 1) & 31, kth_slot = (k1 - 1) >> 5;
    unsigned long long evaluated = 0;

    auto insert = [&](double dn, int in) {
        const unsigned l0 = __ballot_sync(0xffffffffu, grid_key_less(d0, i0, dn, in));
        int pos = __popc(l0);
        if (kSlots > 1) pos += __popc(__ballot_sync(0xffffffffu, grid_key_less(d1, i1, dn, in)));
        if (pos >= k1) return;
        if (kSlots > 1) {
            const double carry_d = shfl_d(d0, 31);
            const int carry_i = __shfl_sync(0xffffffffu, i0, 31);
            const double u_d = shfl_up_d(d1);
            const int u_i = __shfl_up_sync(0xffffffffu, i1, 1);
            const int my = 32 + lane;
            if (my > pos) {
                d1 = lane == 0 ? carry_d : u_d;
                i1 = lane == 0 ? carry_i : u_i;
            } else if (my == pos) {
                d1 = dn;
                i1 = in;
            }
        }
        {
            const double u_d = shfl_up_d(d0);
            const int u_i = __shfl_up_sync(0xffffffffu, i0, 1);
            if (lane > pos) {
                d0 = u_d;
                i0 = u_i;
            } else if (lane == pos) {
                d0 = dn;
                i0 = in;
            }
        }
        kth_d = shfl_d(kth_slot ? d1 : d0, kth_lane);
        kth_i = __shfl_sync(0xffffffffu, kth_slot ? i1 : i0, kth_lane);
    };

    // all points of the cells [cell_a, cell_b] (contiguous in the sorted array)
    auto visit = [&](int cell_a, int cell_b) {
        const int s = start[cell_a], e = start[cell_b + 1];
        for (int base = s; base < e; base += 32) {
            const int i = base + lane;
            double dn = INF;
            int in = 0x7fffffff;
            if (i < e) {
                const float4 y = pts[i];
                in = __float_as_int(y.w);
                double t = (double)qx[0] - (double)y.x;
                dn = fma(t, t, 0.0);
                if (D > 1) {
                    t = (double)qx[1] - (double)y.y;
                    dn = fma(t, t, dn);
                }
                if (D > 2) {
                    t = (double)qx[2] - (double)y.z;
                    dn = fma(t, t, dn);
                }
            }
            evaluated += (i < e) ? 1 : 0;
            unsigned hits = __ballot_sync(0xffffffffu, grid_key_less(dn, in, kth_d, kth_i));
            while (hits) {
                const int l = __ffs(hits) - 1;
                hits &= hits - 1;
                const double cd = shfl_d(dn, l);
                const int ci = __shfl_sync(0xffffffffu, in, l);
                if (grid_key_less(cd, ci, kth_d, kth_i)) insert(cd, ci);   // kth may have tightened (warp-uniform)
            }
        }
    };

    for (int r = 0;; ++r) {
        // ---- cells at Chebyshev distance exactly r from the query's cell, as x-contiguous spans
        if (D == 1) {
            if (r == 0) visit(cq[0], cq[0]);
            else {
                if (cq[0] - r >= 0) visit(cq[0] - r, cq[0] - r);
                if (cq[0] + r < g) visit(cq[0] + r, cq[0] + r);
            }
        } else {
            const int z_lo = D > 2 ? max(0, cq[2] - r) : 0, z_hi = D > 2 ? min(g - 1, cq[2] + r) : 0;
            const int y_lo = max(0, cq[1] - r), y_hi = min(g - 1, cq[1] + r);
            const int x_lo = max(0, cq[0] - r), x_hi = min(g - 1, cq[0] + r);
            for (int z = z_lo; z <= z_hi; ++z)
                for (int y = y_lo; y <= y_hi; ++y) {
                    const int rowbase = (D > 2 ? z * g + y : y) * g;
                    const bool shell = abs(y - cq[1]) == r || (D > 2 && abs(z - cq[2]) == r);
                    if (shell) {
                        visit(rowbase + x_lo, rowbase + x_hi);
                    } else {
                        if (cq[0] - r >= 0) visit(rowbase + cq[0] - r, rowbase + cq[0] - r);
                        if (cq[0] + r < g && r > 0) visit(rowbase + cq[0] + r, rowbase + cq[0] + r);
                    }
                }
        }
        // ---- can an unvisited point still enter?  Distance from the query to the nearest face of
        // the visited block that has cells behind it; the margin covers the float32 rounding of the cell
        // assignment, which grows with the grid: (x - lo) * inv_h carries ~3 ulp of a value up to g, i.e.
        // <= g * 3 * 2^-24 cells (4e-4 for g <= 2048, but 0.19 for the 2^20 cells a 1-D grid may have).
        double bound = INF;
        bool covered = true;
#pragma unroll
        for (int c = 0; c < D; ++c) {
            const double eps = fmax(2e-3, 4.0 * (double)g * 5.9604644775390625e-8) * p.h[c];
            if (cq[c] + r + 1 < g) {
                covered = false;
                bound = fmin(bound, p.lo[c] + (double)(cq[c] + r + 1) * p.h[c] - (double)qx[c] - eps);
            }
            if (cq[c] - r > 0) {
                covered = false;
                bound = fmin(bound, (double)qx[c] - (p.lo[c] + (double)(cq[c] - r) * p.h[c]) - eps);
            }
        }
        if (covered) break;
        if (bound > 0.0 && kth_d < bound * bound) break;
    }

    if (!(kth_d < INF)) {   // fewer than k1 finite distances: the exact fallback serves the row
        if (lane == 0) flag[row] = 1;
    } else {
        if (lane < k1) {
            idx_out[(size_t)row * k1 + lane] = i0;
            dist_out[(size_t)row * k1 + lane] = sqrt(d0);
        }
        if (kSlots > 1 && 32 + lane < k1) {
            idx_out[(size_t)row * k1 + 32 + lane] = i1;
            dist_out[(size_t)row * k1 + 32 + lane] = sqrt(d1);
        }
        if (rk_out && lane == 0) rk_out[row] = sqrt(kth_d);
    }
    if (stats) {
        evaluated = (unsigned long long)warp_sum((double)evaluated);
        if (lane == 0) atomicAdd(stats + 0, evaluated);
    }
}

// cells per dimension: about max(4, k1/3) points per cell on average
int grid_cells_per_dim(int64_t n_ref, int d, int k1) {
    const double occ = k1 / 3.0 > 4.0 ? k1 / 3.0 : 4.0;
    const double cells = (double)n_ref / occ;
    int g = (int)floor(pow(cells > 1.0 ? cells : 1.0, 1.0 / d));
    const int gmax = d == 1 ? (1 << 20) : (d == 2 ? 2048 : 160);
    if (g < 1) g = 1;
    if (g > gmax) g = gmax;
    return g;
}

bool grid_supported(int d, int k1) { return d >= 1 && d <= 3 && k1 <= 64; }

template <int D>
static int launch_grid_d(const float *ref, const float *q, const KnnPlan &p, int32_t *idx_out, double *dist_out,
                         double *rk_out, cudaStream_t st) {
    const int g = p.grid_g;
    const int n_cells = p.grid_cells;
    GridPar *par = reinterpret_cast<GridPar *>(p.grid_par);
    MEPOL_CUDA_CHECK(cudaMemsetAsync(p.grid_cursor, 0, sizeof(int32_t) * (size_t)n_cells, st));
    grid_setup_kernel<<<1, 32, 0, st>>>(p.mean_partial, p.mean_blocks, D, g, par);
    const int nb = (int)((p.n_ref + 255) / 256);
    grid_count_kernel<D><<<nb, 256, 0, st>>>(ref, (int)p.n_ref, par, g, p.grid_cell_of, p.grid_cursor);
    grid_scan_kernel<<<1, kGridScanThreads, 0, st>>>(p.grid_cursor, p.grid_start, n_cells);
    grid_scatter_kernel<D><<<nb, 256, 0, st>>>(ref, (int)p.n_ref, p.grid_cell_of, p.grid_start, p.grid_cursor,
                                                reinterpret_cast<float4 *>(p.grid_pts));
    const int sorted_q = (p.q_is_ref && p.q_row0 == 0 && p.n_q == p.n_ref) ? 1 : 0;
    const int qb = (int)((p.n_q + kGridWarps - 1) / kGridWarps);
    if (p.k1 <= 32)
        grid_query_kernel<D, 1><<<qb, kGridWarps * 32, 0, st>>>(q, (int)p.n_q, sorted_q,
                                                               reinterpret_cast<const float4 *>(p.grid_pts), p.grid_start,
                                                               par, g, p.k1, idx_out, dist_out, rk_out, p.flag, p.stats);
    else
        grid_query_kernel<D, 2><<<qb, kGridWarps * 32, 0, st>>>(q, (int)p.n_q, sorted_q,
                                                               reinterpret_cast<const float4 *>(p.grid_pts), p.grid_start,
                                                               par, g, p.k1, idx_out, dist_out, rk_out, p.flag, p.stats);
    MEPOL_CUDA_CHECK(cudaGetLastError());
    return MEPOL_OK;
}

int launch_grid_search(const float *ref, const float *q, const KnnPlan &p, int32_t *idx_out, double *dist_out,
                       double *rk_out, cudaStream_t st) {
    switch (p.d) {
        case 1: return launch_grid_d<1>(ref, q, p, idx_out, dist_out, rk_out, st);
        case 2: return launch_grid_d<2>(ref, q, p, idx_out, dist_out, rk_out, st);
        case 3: return launch_grid_d<3>(ref, q, p, idx_out, dist_out, rk_out, st);
    }
    set_error("knn: grid search supports d <= 3, got %d", p.d);
    return MEPOL_E_UNSUPPORTED;
}

size_t grid_par_bytes() { return sizeof(GridPar); }

}  // namespace mepol
