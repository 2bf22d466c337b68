/*
 * oracle/knn_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the k-NN search the reference performs at
 *   /root/reference/src/algorithms/mepol.py:190-192
 *     NearestNeighbors(n_neighbors=k+1, metric='euclidean', algorithm='auto').fit(X).kneighbors(X)
 * The arithmetic lives in the third-party dependency scikit-learn (pinned 0.22.2.post1 in
 * /root/reference/requirements.txt:3; the container ships 1.9.0, whose sources were read):
 *   - brute path (d > 15):  sklearn/metrics/_pairwise_distances_reduction/_argkmin.pyx.tp:471-510
 *       d2 = ||x||^2 + (-2 x.y) + ||y||^2 in float64 on float32 data upcast to float64,
 *       clamped at 0, pushed on a fixed-size max-heap; result narrowed to float32 before sqrt
 *       (sklearn/metrics/_dist_metrics.pyx.tp:1018-1019).
 *   - kd_tree path (d <= 15): exact float64 reduced distance sum (x_i-y_i)^2, same heap.
 *   - heap rule: sklearn/utils/_heap.pyx:46 -- a value is rejected when val >= heap top, so among
 *     equal distances the first scanned (lowest index) is kept.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * call this file.  The product path (mepol_b200/) never does.
 *
 * Pinning: checked against sklearn 1.9.0 outputs generated in the build container by
 * oracle/gen_golden.py (tests/golden/knn_*.npz).  The reference itself ships no golden vectors.
 *
 * Deviation, documented: the order of *exactly equal* distances inside a row is an artefact of
 * sklearn's heap layout + introsort; this oracle orders ties by ascending index.  Neighbour SETS
 * are identical; on tie-free data the output is index-exact.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* fixed-size max-heap keyed on value; mirrors the acceptance rule of _heap.pyx:46 */
static void heap_offer(double *val, int64_t *ind, int size, double v, int64_t vi) {
    if (v >= val[0]) return; /* strict-less acceptance: ties lose to earlier entries */
    int cur = 0;
    for (;;) {
        int l = 2 * cur + 1, r = l + 1, s;
        if (l >= size) break;
        if (r >= size) {
            if (val[l] > v) s = l; else break;
        } else if (val[l] >= val[r]) {
            if (v < val[l]) s = l; else break;
        } else {
            if (v < val[r]) s = r; else break;
        }
        val[cur] = val[s];
        ind[cur] = ind[s];
        cur = s;
    }
    val[cur] = v;
    ind[cur] = vi;
}

/* ascending by (value, index): deterministic stand-in for simultaneous_sort (_sorting.pyx) */
static void sort_pairs(double *val, int64_t *ind, int n) {
    for (int i = 1; i < n; ++i) {
        double v = val[i];
        int64_t vi = ind[i];
        int j = i - 1;
        while (j >= 0 && (val[j] > v || (val[j] == v && ind[j] > vi))) {
            val[j + 1] = val[j];
            ind[j + 1] = ind[j];
            --j;
        }
        val[j + 1] = v;
        ind[j + 1] = vi;
    }
}

/*
 * mode 0 = "brute"  (expanded float64 form, float32 narrowing of the returned distance)
 * mode 1 = "kd_tree" (direct float64 differences, no narrowing)
 * q: [n_q, d] float32 row-major, ref: [n_ref, d] float32 row-major.
 * out_dist: [n_q, k1] float64, out_idx: [n_q, k1] int64, rows sorted ascending.
 * returns 0 on success, -1 on bad arguments (k1 > n_ref mirrors sklearn's ValueError).
 */
int mepol_oracle_knn(const float *q, int64_t n_q, const float *ref, int64_t n_ref, int d, int k1,
                     int mode, double *out_dist, int64_t *out_idx, int n_threads) {
    if (!q || !ref || !out_dist || !out_idx || d <= 0 || k1 <= 0 || k1 > n_ref) return -1;
    double *rnorm = (double *)malloc(sizeof(double) * (size_t)n_ref);
    if (!rnorm) return -1;
    for (int64_t j = 0; j < n_ref; ++j) {
        double s = 0.0;
        for (int c = 0; c < d; ++c) {
            double v = (double)ref[j * d + c];
            s += v * v;
        }
        rnorm[j] = s;
    }
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#else
    (void)n_threads;
#endif
#pragma omp parallel
    {
        double *hv = (double *)malloc(sizeof(double) * (size_t)k1);
        int64_t *hi = (int64_t *)malloc(sizeof(int64_t) * (size_t)k1);
        double *xr = (double *)malloc(sizeof(double) * (size_t)d);
#pragma omp for schedule(dynamic, 64)
        for (int64_t i = 0; i < n_q; ++i) {
            double xn = 0.0;
            for (int c = 0; c < d; ++c) {
                xr[c] = (double)q[i * d + c];
                xn += xr[c] * xr[c];
            }
            for (int m = 0; m < k1; ++m) {
                hv[m] = INFINITY;
                hi[m] = 0;
            }
            for (int64_t j = 0; j < n_ref; ++j) {
                const float *y = ref + j * d;
                double d2;
                if (mode == 0) {
                    double dot = 0.0;
                    for (int c = 0; c < d; ++c) dot += xr[c] * (double)y[c];
                    d2 = xn + (-2.0 * dot) + rnorm[j];
                    if (d2 < 0.0) d2 = 0.0;
                } else {
                    d2 = 0.0;
                    for (int c = 0; c < d; ++c) {
                        double t = xr[c] - (double)y[c];
                        d2 += t * t;
                    }
                }
                heap_offer(hv, hi, k1, d2, j);
            }
            sort_pairs(hv, hi, k1);
            for (int m = 0; m < k1; ++m) {
                double r;
                if (mode == 0) {
                    float narrowed = (float)hv[m];
                    r = (double)sqrtf(narrowed);
                } else {
                    r = sqrt(hv[m]);
                }
                out_dist[i * k1 + m] = r;
                out_idx[i * k1 + m] = hi[m];
            }
        }
        free(hv);
        free(hi);
        free(xr);
    }
    free(rnorm);
    return 0;
}

int mepol_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
