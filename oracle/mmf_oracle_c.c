/*
 * mmf_oracle_c.c -- C restatement of the per-series arithmetic of oracle/mmf_oracle.py.  TEST INFRASTRUCTURE ONLY
 * (tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; never the product path).
 *
 * It exists so that the CPU baseline timed next to the GPU is a fair one: float64 accumulation like the NumPy
 * oracle, but one tight pass over each series and all host cores (pthreads), instead of NumPy temporaries.
 * Spec: DESIGN.md section 2 items 5-6 == mmf_oracle.solve_series / fit_forecast_packed, which restate the fit + predict
 * of the reference UDF (group_apply/02_Fine_Grained_Demand_Forecasting.py:435-494) in the whitened calendar basis.
 * PARITY STATUS: parity unpinned against the reference's SARIMAX arithmetic (see the header of mmf_oracle.py);
 * tests/test_oracle.py pins this file against the NumPy oracle to 1e-9.
 *
 *   y        [n, ld]         float32, NaN / Inf = missing
 *   A        [n_rows, 16]    float64 whitened design (rows [0,t_fit) fit, the rest predict), from mmf_oracle.whiten
 *   kept     [16]            0/1: column retained on the calendar
 *   out      [n, n_pred]     float64 predictions for rows [pred_start, pred_start + n_pred)
 *   status   [n]             0 ok, 1 no observed fit row, 2 a column was dropped for this series' mask
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define P 16
#define PIVOT_TOL 1e-3

static int solve_masked(const float* y, const double* A, const int* kept, int t_fit, double* gamma) {
  double G[P][P], L[P][P], b[P];
  int keep[P];
  memset(G, 0, sizeof(G));
  memset(L, 0, sizeof(L));
  memset(b, 0, sizeof(b));
  for (int t = 0; t < t_fit; ++t) {
    if (!isfinite(y[t])) continue;
    const double* a = A + (size_t)t * P;
    const double v = (double)y[t];
    for (int i = 0; i < P; ++i) {
      b[i] += a[i] * v;
      for (int j = 0; j <= i; ++j) G[i][j] += a[i] * a[j];
    }
  }
  int dropped = 0;
  for (int j = 0; j < P; ++j) {
    keep[j] = 0;
    const double gjj = G[j][j];
    if (!kept[j] || gjj <= 0.0) continue;
    double d = gjj;
    for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
    if (d <= PIVOT_TOL * gjj) { dropped = 1; continue; }
    keep[j] = 1;
    L[j][j] = sqrt(d);
    for (int i = j + 1; i < P; ++i) {
      double s = G[i][j];
      for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
      L[i][j] = s / L[j][j];
    }
  }
  double z[P];
  for (int j = 0; j < P; ++j) {
    if (!keep[j]) { z[j] = 0.0; continue; }
    double s = b[j];
    for (int k = 0; k < j; ++k) if (keep[k]) s -= L[j][k] * z[k];
    z[j] = s / L[j][j];
  }
  for (int j = P - 1; j >= 0; --j) {
    if (!keep[j]) { gamma[j] = 0.0; continue; }
    double s = z[j];
    for (int i = j + 1; i < P; ++i) if (keep[i]) s -= L[i][j] * gamma[i];
    gamma[j] = s / L[j][j];
  }
  return dropped ? 2 : 0;
}

typedef struct {
  const float* y; int64_t lo, hi, ld; int32_t t_fit; const double* A; const int* kp;
  int32_t pred_start, n_pred; double* out; int32_t* status;
  int cpu;                    /* >= 0: pin the worker to this core (reproducible placement for the timed baseline) */
  float* touch; size_t touch_bytes;   /* first-touch job: zero this slice from the pinned worker */
} job_t;

static void pin_self(int cpu) {
  if (cpu < 0) return;
  cpu_set_t one;
  CPU_ZERO(&one);
  CPU_SET(cpu, &one);
  pthread_setaffinity_np(pthread_self(), sizeof(one), &one);
}

/* k-th core of the process affinity mask, or -1 */
static int kth_cpu(const cpu_set_t* set, int k) {
  for (int c = 0; c < CPU_SETSIZE; ++c)
    if (CPU_ISSET(c, set) && k-- == 0) return c;
  return -1;
}

static void* toucher(void* arg) {
  const job_t* jb = (const job_t*)arg;
  pin_self(jb->cpu);
  memset(jb->touch, 0, jb->touch_bytes);
  return NULL;
}

static void* worker(void* arg) {
  const job_t* jb = (const job_t*)arg;
  pin_self(jb->cpu);
  const int32_t t_fit = jb->t_fit, n_pred = jb->n_pred;
  const double* A = jb->A;
  for (int64_t i = jb->lo; i < jb->hi; ++i) {
    const float* yr = jb->y + i * jb->ld;
    double g[P];
    int nobs = 0, nmiss = 0;
    for (int j = 0; j < P; ++j) g[j] = 0.0;
    for (int t = 0; t < t_fit; ++t) {
      const float v = yr[t];
      if (!isfinite(v)) { ++nmiss; continue; }
      ++nobs;
      const double* a = A + (size_t)t * P;
      const double dv = (double)v;
      for (int j = 0; j < P; ++j) g[j] += a[j] * dv;          /* fully observed: G_i = I, gamma = b */
    }
    int st = 0;
    if (nobs == 0) {
      st = 1;
      for (int k = 0; k < n_pred; ++k) jb->out[i * n_pred + k] = NAN;
    } else {
      if (nmiss > 0) st = solve_masked(yr, A, jb->kp, t_fit, g);
      else for (int j = 0; j < P; ++j) if (!jb->kp[j]) g[j] = 0.0;
      for (int k = 0; k < n_pred; ++k) {
        const double* a = A + (size_t)(jb->pred_start + k) * P;
        double s = 0.0;
        for (int j = 0; j < P; ++j) s += a[j] * g[j];
        jb->out[i * n_pred + k] = s;
      }
    }
    jb->status[i] = st;
  }
  return NULL;
}

/* NUMA-aware sample buffer for the timed CPU baseline: n rows of ld floats whose pages are first touched by the same
 * pinned thread (k-th core of the affinity mask, rows [k*per, (k+1)*per)) that mmf_oracle_fit_forecast later assigns to
 * those rows, so every worker streams from its own socket's memory whatever box the bench lands on.  free() it with
 * mmf_oracle_free_local.  The caller copies the sample in afterwards (the pages are already placed). */
float* mmf_oracle_alloc_local(int64_t n, int64_t ld, int32_t n_threads) {
  cpu_set_t set;
  const int have = (sched_getaffinity(0, sizeof(set), &set) == 0) ? CPU_COUNT(&set) : 1;
  if (n_threads <= 0) n_threads = have;
  if ((int64_t)n_threads > n) n_threads = n > 0 ? (int)n : 1;
  float* buf = NULL;
  if (posix_memalign((void**)&buf, 4096, (size_t)n * (size_t)ld * sizeof(float) + 4096) != 0) return NULL;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
  job_t* jobs = (job_t*)calloc((size_t)n_threads, sizeof(job_t));
  const int64_t per = (n + n_threads - 1) / n_threads;
  for (int k = 0; k < n_threads; ++k) {
    const int64_t lo = k * per < n ? k * per : n, hi = (lo + per < n) ? lo + per : n;
    jobs[k].cpu = n_threads <= have ? kth_cpu(&set, k) : -1;
    jobs[k].touch = buf + lo * ld;
    jobs[k].touch_bytes = (size_t)(hi - lo) * (size_t)ld * sizeof(float);
    pthread_create(&th[k], NULL, toucher, &jobs[k]);
  }
  for (int k = 0; k < n_threads; ++k) pthread_join(th[k], NULL);
  free(th);
  free(jobs);
  return buf;
}
void mmf_oracle_free_local(float* p) { free(p); }

/* n_threads <= 0: one thread per core of the affinity mask, each pinned to its core.  Returns the threads used. */
int mmf_oracle_fit_forecast(const float* y, int64_t n, int64_t ld, int32_t t_fit, const double* A, const int32_t* kept,
                            int32_t pred_start, int32_t n_pred, double* out, int32_t* status, int32_t n_threads) {
  int kp[P];
  for (int j = 0; j < P; ++j) kp[j] = kept[j];
  cpu_set_t set;
  const int have = (sched_getaffinity(0, sizeof(set), &set) == 0) ? CPU_COUNT(&set) : 1;
  if (n_threads <= 0) n_threads = have;
  if ((int64_t)n_threads > n) n_threads = n > 0 ? (int)n : 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
  job_t* jobs = (job_t*)malloc(sizeof(job_t) * (size_t)n_threads);
  const int64_t per = (n + n_threads - 1) / n_threads;
  for (int k = 0; k < n_threads; ++k) {
    const int64_t lo = k * per, hi = (lo + per < n) ? lo + per : n;
    jobs[k] = (job_t){y, lo < n ? lo : n, hi, ld, t_fit, A, kp, pred_start, n_pred, out, status,
                      n_threads <= have ? kth_cpu(&set, k) : -1, NULL, 0};
    pthread_create(&th[k], NULL, worker, &jobs[k]);
  }
  for (int k = 0; k < n_threads; ++k) pthread_join(th[k], NULL);
  free(th);
  free(jobs);
  return n_threads;
}
