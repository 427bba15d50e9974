/*
 * tip_oracle.c — plain-C restatement of the reference's hot-path arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY: built into oracle/libtiporacle.so and loaded by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.  The
 * product package never links or loads this file.
 *
 * Build: gcc -O3 -mavx2 -fopenmp -ffp-contract=off -fno-fast-math -shared -fPIC (oracle/c_oracle.py:build).
 * Vectorising the eight independent stride-8 accumulators does not change any rounding.
 * -ffp-contract=off is REQUIRED: the reference's NumPy arithmetic never fuses a*b+c.
 *
 * What each function follows (paths relative to /root/reference):
 *   oracle_pairwise_sumsq_*   NumPy `add.reduce` pairwise order applied to x*x, which is
 *                             what np.linalg.norm(axis=2) does in src/core/surprise.py:640
 *   oracle_dsa_*              src/core/surprise.py:558-651 (DSA.__call__, _dsa_distances,
 *                             _get_closest_ats): min / first-occurrence argmin in the input dtype
 *   oracle_kde_eval           scipy==1.4.1 _stats.gaussian_kernel_estimate loop nest, called
 *                             from src/core/stable_kde.py:101 (third-party, restated)
 *   oracle_deepgini_*         src/core/deepgini.py:33-34
 *   oracle_kmnc_*             src/core/neuron_coverage.py:82-94 (+ sum_score :8-22)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PW_BLOCK 128

/* ---- NumPy pairwise sum of squared differences (or squares when y == NULL) ---------- */
#define DEFINE_PAIRWISE(NAME, T)                                                          \
  static T NAME(const T* x, const T* y, int64_t n) {                                      \
    if (n < 8) {                                                                          \
      T res = (T)0;                                                                       \
      for (int64_t i = 0; i < n; i++) {                                                   \
        T dlt = y ? (T)(x[i] - y[i]) : x[i];                                              \
        T sq = (T)(dlt * dlt);                                                            \
        res = (T)(res + sq);                                                              \
      }                                                                                   \
      return res;                                                                         \
    } else if (n <= PW_BLOCK) {                                                           \
      T r[8];                                                                             \
      for (int j = 0; j < 8; j++) {                                                       \
        T dlt = y ? (T)(x[j] - y[j]) : x[j];                                              \
        r[j] = (T)(dlt * dlt);                                                            \
      }                                                                                   \
      int64_t i;                                                                          \
      for (i = 8; i < n - (n % 8); i += 8) {                                              \
        for (int j = 0; j < 8; j++) {                                                     \
          T dlt = y ? (T)(x[i + j] - y[i + j]) : x[i + j];                                \
          T sq = (T)(dlt * dlt);                                                          \
          r[j] = (T)(r[j] + sq);                                                          \
        }                                                                                 \
      }                                                                                   \
      T res = (T)((T)((T)(r[0] + r[1]) + (T)(r[2] + r[3])) +                              \
                  (T)((T)(r[4] + r[5]) + (T)(r[6] + r[7])));                              \
      for (; i < n; i++) {                                                                \
        T dlt = y ? (T)(x[i] - y[i]) : x[i];                                              \
        T sq = (T)(dlt * dlt);                                                            \
        res = (T)(res + sq);                                                              \
      }                                                                                   \
      return res;                                                                         \
    } else {                                                                              \
      int64_t n2 = n / 2;                                                                 \
      n2 -= n2 % 8;                                                                       \
      T a = NAME(x, y, n2);                                                               \
      T b = NAME(x + n2, y ? y + n2 : NULL, n - n2);                                      \
      return (T)(a + b);                                                                  \
    }                                                                                     \
  }

DEFINE_PAIRWISE(pw_f32, float)
DEFINE_PAIRWISE(pw_f64, double)

float oracle_pairwise_sumsq_f32(const float* x, const float* y, int64_t n) { return pw_f32(x, y, n); }
double oracle_pairwise_sumsq_f64(const double* x, const double* y, int64_t n) { return pw_f64(x, y, n); }

/* ---- DSA -------------------------------------------------------------------------- */
/* For every test row: stage 1 = nearest train row of the predicted class (distance and
 * ORIGINAL train index, first occurrence on ties); stage 2 = distance from THAT TRAIN ROW
 * to the nearest train row of any other class (surprise.py:627-629).  Rows whose
 * predicted class has no train rows, or with no other-class rows, get idx -1 / NaN. */
#define DEFINE_DSA(NAME, T, PW, SQRT)                                                     \
  void NAME(const T* train, const int64_t* train_pred, int64_t n, int64_t d,              \
            const T* test, const int64_t* test_pred, int64_t m,                           \
            T* dist_a, T* dist_b, int64_t* idx_a, int threads) {                          \
    (void)threads;                                                                        \
    _Pragma("omp parallel for schedule(dynamic, 4) num_threads(threads)")                 \
    for (int64_t t = 0; t < m; t++) {                                                     \
      const T* x = test + t * d;                                                          \
      int64_t c = test_pred[t];                                                           \
      T best = (T)INFINITY;                                                               \
      int64_t bi = -1;                                                                    \
      for (int64_t i = 0; i < n; i++) {                                                   \
        if (train_pred[i] != c) continue;                                                 \
        T dist = SQRT(PW(x, train + i * d, d));                                           \
        if (bi < 0 || dist < best) { best = dist; bi = i; }                               \
      }                                                                                   \
      idx_a[t] = bi;                                                                      \
      dist_a[t] = bi < 0 ? (T)NAN : best;                                                 \
      if (bi < 0) { dist_b[t] = (T)NAN; continue; }                                       \
      const T* w = train + bi * d;                                                        \
      T bestb = (T)INFINITY;                                                              \
      int found = 0;                                                                      \
      for (int64_t i = 0; i < n; i++) {                                                   \
        if (train_pred[i] == c) continue;                                                 \
        T dist = SQRT(PW(w, train + i * d, d));                                           \
        if (!found || dist < bestb) { bestb = dist; found = 1; }                          \
      }                                                                                   \
      dist_b[t] = found ? bestb : (T)NAN;                                                 \
    }                                                                                     \
  }

DEFINE_DSA(oracle_dsa_f32, float, pw_f32, sqrtf)
DEFINE_DSA(oracle_dsa_f64, double, pw_f64, sqrt)

/* Same arithmetic, parallel over TRAIN rows instead of test rows (few test rows against a very large
 * training set, e.g. the in-bench parity check of the C5 slice): every thread scans a contiguous,
 * ascending chunk (first occurrence inside the chunk), chunks merge on (distance, index). */
#define DEFINE_DSA_ROWS(NAME, T, PW, SQRT)                                                \
  void NAME(const T* train, const int64_t* train_pred, int64_t n, int64_t d,              \
            const T* test, const int64_t* test_pred, int64_t m,                           \
            T* dist_a, T* dist_b, int64_t* idx_a, int threads) {                          \
    (void)threads;                                                                        \
    for (int64_t t = 0; t < m; t++) {                                                     \
      const T* x = test + t * d;                                                          \
      const int64_t c = test_pred[t];                                                     \
      T best = (T)INFINITY;                                                               \
      int64_t bi = -1;                                                                    \
      _Pragma("omp parallel num_threads(threads)")                                        \
      {                                                                                   \
        T lb = (T)INFINITY;                                                               \
        int64_t li = -1;                                                                  \
        _Pragma("omp for schedule(static) nowait")                                        \
        for (int64_t i = 0; i < n; i++) {                                                 \
          if (train_pred[i] != c) continue;                                               \
          T dist = SQRT(PW(x, train + i * d, d));                                         \
          if (li < 0 || dist < lb) { lb = dist; li = i; }                                 \
        }                                                                                 \
        _Pragma("omp critical")                                                           \
        {                                                                                 \
          if (li >= 0 && (bi < 0 || lb < best || (lb == best && li < bi))) { best = lb; bi = li; } \
        }                                                                                 \
      }                                                                                   \
      idx_a[t] = bi;                                                                      \
      dist_a[t] = bi < 0 ? (T)NAN : best;                                                 \
      if (bi < 0) { dist_b[t] = (T)NAN; continue; }                                       \
      const T* w = train + bi * d;                                                        \
      T bestb = (T)INFINITY;                                                              \
      int found = 0;                                                                      \
      _Pragma("omp parallel num_threads(threads)")                                        \
      {                                                                                   \
        T lb = (T)INFINITY;                                                               \
        int lf = 0;                                                                       \
        _Pragma("omp for schedule(static) nowait")                                        \
        for (int64_t i = 0; i < n; i++) {                                                 \
          if (train_pred[i] == c) continue;                                               \
          T dist = SQRT(PW(w, train + i * d, d));                                         \
          if (!lf || dist < lb) { lb = dist; lf = 1; }                                    \
        }                                                                                 \
        _Pragma("omp critical")                                                           \
        {                                                                                 \
          if (lf && (!found || lb < bestb)) { bestb = lb; found = 1; }                    \
        }                                                                                 \
      }                                                                                   \
      dist_b[t] = found ? bestb : (T)NAN;                                                 \
    }                                                                                     \
  }

DEFINE_DSA_ROWS(oracle_dsa_rows_f32, float, pw_f32, sqrtf)
DEFINE_DSA_ROWS(oracle_dsa_rows_f64, double, pw_f64, sqrt)

/* ---- Gaussian KDE evaluate (scipy 1.4.1 gaussian_kernel_estimate) ------------------- */
/* points_w: n x d whitened train, xi_w: m x d whitened test, weight = 1/n, norm as in
 * the Cython source.  estimate[j] += weight * (exp(-arg/2) * norm) accumulated over i in
 * ascending order, arg accumulated over k in ascending order — the literal loop nest
 * (i outer, j inner there; per-j accumulation order over i is what matters and is kept). */
void oracle_kde_eval(const double* points_w, int64_t n, const double* xi_w, int64_t m, int64_t d,
                     double weight, double norm, double* estimate, int threads) {
  (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads)
  for (int64_t j = 0; j < m; j++) {
    const double* q = xi_w + j * d;
    double est = 0.0;
    for (int64_t i = 0; i < n; i++) {
      const double* p = points_w + i * d;
      double arg = 0.0;
      for (int64_t k = 0; k < d; k++) {
        double r = p[k] - q[k];
        arg += r * r;
      }
      arg = exp(-arg / 2.0) * norm;
      est += weight * arg;
    }
    estimate[j] = est;
  }
}

/* ---- DeepGini ---------------------------------------------------------------------- */
#define DEFINE_GINI(NAME, T, PW)                                                          \
  void NAME(const T* p, int64_t n, int64_t c, int64_t* pred, T* gini, int threads) {      \
    (void)threads;                                                                        \
    _Pragma("omp parallel for schedule(static) num_threads(threads)")                     \
    for (int64_t r = 0; r < n; r++) {                                                     \
      const T* row = p + r * c;                                                           \
      int64_t best = 0;                                                                   \
      for (int64_t k = 1; k < c; k++)                                                     \
        if (row[k] > row[best]) best = k;                                                 \
      pred[r] = best;                                                                     \
      gini[r] = (T)((T)1 - PW(row, NULL, c));                                             \
    }                                                                                     \
  }

DEFINE_GINI(oracle_deepgini_f32, float, pw_f32)
DEFINE_GINI(oracle_deepgini_f64, double, pw_f64)

/* ---- KMNC -------------------------------------------------------------------------- */
/* thresholds t[i][j] are passed in exactly as NumPy built them ((k+1) x d, same dtype as
 * the comparison dtype); literal loop over sections: profile = t[i] <= a < t[i+1]. */
#define DEFINE_KMNC(NAME, T)                                                              \
  void NAME(const T* act, int64_t n, int64_t d, const T* thresh, int64_t sections,        \
            int32_t* bucket, int64_t* score, int threads) {                               \
    (void)threads;                                                                        \
    _Pragma("omp parallel for schedule(static) num_threads(threads)")                     \
    for (int64_t r = 0; r < n; r++) {                                                     \
      int64_t cnt = 0;                                                                    \
      for (int64_t j = 0; j < d; j++) {                                                   \
        T a = act[r * d + j];                                                             \
        int32_t b = -1;                                                                   \
        for (int64_t i = 0; i < sections; i++) {                                          \
          if (thresh[i * d + j] <= a && a < thresh[(i + 1) * d + j]) { b = (int32_t)i; cnt++; } \
        }                                                                                 \
        bucket[r * d + j] = b;                                                            \
      }                                                                                   \
      score[r] = cnt;                                                                     \
    }                                                                                     \
  }

DEFINE_KMNC(oracle_kmnc_f32, float)
DEFINE_KMNC(oracle_kmnc_f64, double)

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
