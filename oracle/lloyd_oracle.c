/* C restatement of one scikit-learn Lloyd iteration for d = 3 -- TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product path never links or dlopens it.
 *
 * Follows sklearn/cluster/_k_means_lloyd.pyx (scikit-learn 1.9.0; the library behind the
 * reference's KMeans call at members/jasraj/land_use_classification/core.py:227-228):
 *   - distances as  ||c||^2 - 2 x.c  in float64          (pyx:191-203)
 *   - argmin with strict '<', lowest index wins ties      (pyx:205-213)
 *   - per-thread sums / counts, merged afterwards         (pyx:215-218, 143-152)
 * The merge here is in thread-index order (sklearn's is lock-arrival order), so this
 * oracle is run-to-run deterministic.  Points may be given as float32 SoA (what the GPU
 * path holds) and are widened to float64 and mean-centred on the fly, mirroring
 * sklearn/cluster/_kmeans.py:1487-1493 without a 24 B/point host copy.
 *
 * Pinned by tests/test_oracle.py against the numpy restatement, the golden fixtures and
 * live scikit-learn.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_MAX_THREADS 256

static int resolve_threads(int n_threads) {
#ifdef _OPENMP
  if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
  n_threads = 1;
#endif
  if (n_threads > ORACLE_MAX_THREADS) n_threads = ORACLE_MAX_THREADS;
  return n_threads;
}

int oracle_num_threads(void) { return resolve_threads(0); }

/* One E-step (+ M-step sums when sums != NULL) over n points.
 * x,y,z: float32 SoA.  mean[3]: subtracted in float64 before anything else.
 * centers: k x 3 float64, in the SAME (mean-centred) frame.
 * labels: int32[n] out.  sums: k x 3 float64 out (centred frame).  counts: k float64 out.
 * inertia_out (optional): sum of squared distances to the assigned centre (direct form,
 * sklearn/cluster/_k_means_common.pyx:94-124). */
int oracle_lloyd_step_f32soa(const float* x, const float* y, const float* z, int64_t n,
                             const double* mean, const double* centers, int k,
                             int32_t* labels, double* sums, double* counts,
                             double* inertia_out, int n_threads) {
  if (k <= 0 || n < 0) return -1;
  n_threads = resolve_threads(n_threads);
  double* cn = (double*)malloc(sizeof(double) * (size_t)k);
  double* tsums = (double*)calloc((size_t)n_threads * (size_t)k * 4, sizeof(double));
  double* tin = (double*)calloc((size_t)n_threads, sizeof(double));
  if (!cn || !tsums || !tin) { free(cn); free(tsums); free(tin); return -2; }
  for (int j = 0; j < k; ++j) {
    const double* c = centers + 3 * (size_t)j;
    cn[j] = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
  }
  const double mx = mean ? mean[0] : 0.0, my = mean ? mean[1] : 0.0, mz = mean ? mean[2] : 0.0;
#pragma omp parallel num_threads(n_threads)
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num();
#else
    const int t = 0;
#endif
    double* ls = tsums + (size_t)t * (size_t)k * 4;
    double lin = 0.0;
#pragma omp for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      const double px = (double)x[i] - mx, py = (double)y[i] - my, pz = (double)z[i] - mz;
      double best = cn[0] - 2.0 * (px * centers[0] + py * centers[1] + pz * centers[2]);
      int lab = 0;
      for (int j = 1; j < k; ++j) {
        const double* c = centers + 3 * (size_t)j;
        const double d = cn[j] - 2.0 * (px * c[0] + py * c[1] + pz * c[2]);
        if (d < best) { best = d; lab = j; }
      }
      labels[i] = lab;
      if (sums) {
        double* a = ls + 4 * (size_t)lab;
        a[0] += px; a[1] += py; a[2] += pz; a[3] += 1.0;
      }
      if (inertia_out) {
        const double* c = centers + 3 * (size_t)lab;
        const double dx = px - c[0], dy = py - c[1], dz = pz - c[2];
        lin += dx * dx + dy * dy + dz * dz;
      }
    }
    tin[t] = lin;
  }
  if (sums) {
    memset(sums, 0, sizeof(double) * (size_t)k * 3);
    memset(counts, 0, sizeof(double) * (size_t)k);
    for (int t = 0; t < n_threads; ++t) {
      const double* ls = tsums + (size_t)t * (size_t)k * 4;
      for (int j = 0; j < k; ++j) {
        sums[3 * j + 0] += ls[4 * j + 0];
        sums[3 * j + 1] += ls[4 * j + 1];
        sums[3 * j + 2] += ls[4 * j + 2];
        counts[j] += ls[4 * j + 3];
      }
    }
  }
  if (inertia_out) {
    double s = 0.0;
    for (int t = 0; t < n_threads; ++t) s += tin[t];
    *inertia_out = s;
  }
  free(cn); free(tsums); free(tin);
  return 0;
}

/* Same, for float64 AoS points X[n][3] already in the centred frame (what sklearn holds). */
int oracle_lloyd_step_f64(const double* X, int64_t n, const double* centers, int k,
                          int32_t* labels, double* sums, double* counts, int n_threads) {
  if (k <= 0 || n < 0) return -1;
  n_threads = resolve_threads(n_threads);
  double* cn = (double*)malloc(sizeof(double) * (size_t)k);
  double* tsums = (double*)calloc((size_t)n_threads * (size_t)k * 4, sizeof(double));
  if (!cn || !tsums) { free(cn); free(tsums); return -2; }
  for (int j = 0; j < k; ++j) {
    const double* c = centers + 3 * (size_t)j;
    cn[j] = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
  }
#pragma omp parallel num_threads(n_threads)
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num();
#else
    const int t = 0;
#endif
    double* ls = tsums + (size_t)t * (size_t)k * 4;
#pragma omp for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      const double px = X[3 * i], py = X[3 * i + 1], pz = X[3 * i + 2];
      double best = cn[0] - 2.0 * (px * centers[0] + py * centers[1] + pz * centers[2]);
      int lab = 0;
      for (int j = 1; j < k; ++j) {
        const double* c = centers + 3 * (size_t)j;
        const double d = cn[j] - 2.0 * (px * c[0] + py * c[1] + pz * c[2]);
        if (d < best) { best = d; lab = j; }
      }
      labels[i] = lab;
      if (sums) {
        double* a = ls + 4 * (size_t)lab;
        a[0] += px; a[1] += py; a[2] += pz; a[3] += 1.0;
      }
    }
  }
  if (sums) {
    memset(sums, 0, sizeof(double) * (size_t)k * 3);
    memset(counts, 0, sizeof(double) * (size_t)k);
    for (int t = 0; t < n_threads; ++t) {
      const double* ls = tsums + (size_t)t * (size_t)k * 4;
      for (int j = 0; j < k; ++j) {
        sums[3 * j + 0] += ls[4 * j + 0];
        sums[3 * j + 1] += ls[4 * j + 1];
        sums[3 * j + 2] += ls[4 * j + 2];
        counts[j] += ls[4 * j + 3];
      }
    }
  }
  free(cn); free(tsums);
  return 0;
}
