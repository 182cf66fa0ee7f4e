/*
 * maxsim_oracle.c — CPU ORACLE, TEST INFRASTRUCTURE ONLY (plain C restatement).
 *
 * Restates the reference's exhaustive late-interaction scoring for one shard:
 *   - the dense contraction `D_packed @ Q.T`           (CB/modeling/colbert.py:304)
 *   - the per-document running max over its token rows  (CB/modeling/segmented_maxsim.cpp:22-47),
 *     zero-initialised when relu != 0 (segmented_maxsim.cpp:58-59 torch::zeros) or the true max
 *     (padded/GPU path: -9999 fill then max, CB/modeling/colbert.py:240-241) when relu == 0
 *   - the sum over query tokens                         (segmented_maxsim.cpp:92 / colbert.py:263)
 *   - ranking by descending score then `[:k]`           (CB/search/index_storage.py:95,
 *                                                        CB/searcher.py:132)
 * Threading mirrors the reference's scheme: contiguous document ranges of ceil(ndocs/nthreads)
 * per pthread (segmented_maxsim.cpp:25-28).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may call this.  Parity is pinned
 * through the npz fixtures under tests/golden (outputs of the reference itself), see oracle/maxsim_oracle.py.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  const float* Q;      /* [nq, dim] one query */
  const float* D;      /* [sum(doclens), dim] */
  const int64_t* off;  /* [n + 1] */
  int nq, dim, relu;
  int64_t p0, p1;
  float* out;          /* [n] */
} job_t;

static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  float* m = (float*)malloc(sizeof(float) * (size_t)j->nq);
  for (int64_t p = j->p0; p < j->p1; ++p) {
    const float init = j->relu ? 0.0f : -9999.0f; /* |dot| <= 1 for normalised rows */
    for (int i = 0; i < j->nq; ++i) m[i] = init;
    for (int64_t t = j->off[p]; t < j->off[p + 1]; ++t) {
      const float* d = j->D + t * j->dim;
      for (int i = 0; i < j->nq; ++i) {
        const float* q = j->Q + (int64_t)i * j->dim;
        float acc = 0.0f;
        for (int c = 0; c < j->dim; ++c) acc += q[c] * d[c];
        if (acc > m[i]) m[i] = acc;
      }
    }
    float s = 0.0f;
    for (int i = 0; i < j->nq; ++i) s += m[i];
    j->out[p] = s;
  }
  free(m);
  return NULL;
}

/* out[b][p] = sum_i max_j <Q[b][i], D_p[j]>;  returns 0 on success. */
int flmr_oracle_maxsim_scores(const float* Q, int n_queries, int nq, const float* D,
                              const int32_t* doclens, int64_t n_passages, int dim, int relu,
                              int nthreads, float* out) {
  if (!Q || !D || !doclens || !out || nq <= 0 || dim <= 0 || n_passages < 0) return 1;
  if (nthreads < 1) nthreads = 1;
  int64_t* off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_passages + 1));
  if (!off) return 2;
  off[0] = 0;
  for (int64_t p = 0; p < n_passages; ++p) off[p + 1] = off[p] + doclens[p];
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
  job_t* jobs = (job_t*)malloc(sizeof(job_t) * (size_t)nthreads);
  const int64_t per = (n_passages + nthreads - 1) / nthreads;
  for (int b = 0; b < n_queries; ++b) {
    int started = 0;
    for (int t = 0; t < nthreads; ++t) {
      job_t* j = &jobs[t];
      j->Q = Q + (int64_t)b * nq * dim;
      j->D = D;
      j->off = off;
      j->nq = nq;
      j->dim = dim;
      j->relu = relu;
      j->p0 = t * per < n_passages ? t * per : n_passages;
      j->p1 = (t + 1) * per < n_passages ? (t + 1) * per : n_passages;
      j->out = out + (int64_t)b * n_passages;
      if (j->p0 < j->p1 && pthread_create(&th[t], NULL, worker, j) == 0) {
        started |= 1 << (t & 30);
        j->dim = dim; /* mark started below */
      } else {
        j->p1 = j->p0 - 1; /* not started */
        if (j->p0 < (t + 1) * per && j->p0 < n_passages) { /* thread creation failed: run inline */
          j->p1 = (t + 1) * per < n_passages ? (t + 1) * per : n_passages;
          worker(j);
          j->p1 = j->p0 - 1;
        }
      }
    }
    for (int t = 0; t < nthreads; ++t)
      if (jobs[t].p1 >= jobs[t].p0 && jobs[t].p0 < jobs[t].p1) pthread_join(th[t], NULL);
    (void)started;
  }
  free(jobs);
  free(th);
  free(off);
  return 0;
}

/* Descending score, ties by ascending id; pads with (-inf, -1) when n < k. */
int flmr_oracle_topk(const float* scores, int64_t n, int k, int64_t pid_base, float* out_s,
                     int64_t* out_p) {
  if (!scores || !out_s || !out_p || k < 1) return 1;
  char* taken = (char*)calloc((size_t)(n > 0 ? n : 1), 1);
  if (!taken) return 2;
  for (int r = 0; r < k; ++r) {
    int64_t best = -1;
    for (int64_t p = 0; p < n; ++p)
      if (!taken[p] && (best < 0 || scores[p] > scores[best])) best = p;
    if (best < 0) {
      out_s[r] = -INFINITY;
      out_p[r] = -1;
    } else {
      taken[best] = 1;
      out_s[r] = scores[best];
      out_p[r] = best + pid_base;
    }
  }
  free(taken);
  return 0;
}
