/*
 * oracle.c — plain C restatement of the reference's aggregation rules.  TEST INFRASTRUCTURE ONLY
 * (same status as oracle/ref_numpy.py: only tests/, __graft_entry__.smoke() and bench.py's CPU legs
 * may load it; the product package never does).
 *
 * It follows the reference's algorithm (file:line cited per function) with float64 accumulation
 * ("arbiter" arithmetic) and pthreads over independent units, so it can check the CUDA path at sizes the
 * Python-loop oracle cannot reach and serve as an all-cores CPU baseline.  It is itself pinned: the
 * CPU tests compare it against oracle/ref_numpy.py (which is pinned bit-for-bit to the reference's
 * golden vectors) on the committed fixtures.
 *
 * Build: make -C oracle   ->  oracle/liboracle_c.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* Minimal fork-join runtime on pthreads (this image has no libgomp): items [0, total) are handed
 * out in chunks from a shared counter; fn(ctx, begin, end, thread_id). */
typedef void (*orc_body)(void* ctx, int64_t begin, int64_t end, int tid);
typedef struct { orc_body fn; void* ctx; int64_t total, chunk; int64_t next; pthread_mutex_t mu; int tid_seq; } orc_job;

int orc_num_threads(void) {
  const char* e = getenv("ORC_THREADS");
  if (e && atoi(e) > 0) return atoi(e);
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}
static void* orc_worker(void* arg) {
  orc_job* j = (orc_job*)arg;
  pthread_mutex_lock(&j->mu);
  const int tid = j->tid_seq++;
  pthread_mutex_unlock(&j->mu);
  for (;;) {
    pthread_mutex_lock(&j->mu);
    const int64_t b = j->next;
    j->next += j->chunk;
    pthread_mutex_unlock(&j->mu);
    if (b >= j->total) break;
    const int64_t e = b + j->chunk < j->total ? b + j->chunk : j->total;
    j->fn(j->ctx, b, e, tid);
  }
  return NULL;
}
static void orc_parallel_for(int64_t total, int64_t chunk, orc_body fn, void* ctx) {
  int nt = orc_num_threads();
  if (nt > 256) nt = 256;
  if ((int64_t)nt > (total + chunk - 1) / chunk) nt = (int)((total + chunk - 1) / chunk);
  if (nt < 1) nt = 1;
  orc_job j = {fn, ctx, total, chunk, 0, PTHREAD_MUTEX_INITIALIZER, 0};
  pthread_t th[256];
  for (int t = 1; t < nt; ++t) pthread_create(&th[t], NULL, orc_worker, &j);
  orc_worker(&j);
  for (int t = 1; t < nt; ++t) pthread_join(th[t], NULL);
}

/* defences.py:16-21 — d2[i][j] = sum_k fl32(g_i[k] - g_j[k])^2, accumulated in float64.
 * Column-blocked so the n rows of a block stay in cache while all pairs are visited. */
typedef struct { const float* G; int n; int64_t d, ld; double* loc; } pw_ctx;
static void pw_body(void* vc, int64_t b0, int64_t b1, int tid) {
  pw_ctx* c = (pw_ctx*)vc;
  const int64_t B = 2048;
  const int n = c->n;
  double* loc = c->loc + (size_t)tid * n * n;
  for (int64_t b = b0; b < b1; ++b) {
    const int64_t c0 = b * B, c1 = (c0 + B < c->d) ? c0 + B : c->d;
    for (int i = 1; i < n; ++i) {
      const float* gi = c->G + (int64_t)i * c->ld;
      for (int j = 0; j < i; ++j) {
        const float* gj = c->G + (int64_t)j * c->ld;
        double acc = 0.0;
        for (int64_t k = c0; k < c1; ++k) {
          const float df = gi[k] - gj[k];              /* fp32 subtract, as the reference */
          acc += (double)df * (double)df;
        }
        loc[(size_t)i * n + j] += acc;
      }
    }
  }
}
void orc_pairwise_sqdist(const float* G, int n, int64_t d, int64_t ld, double* d2) {
  const int64_t B = 2048;
  const int64_t nblk = (d + B - 1) / B;
  int nt = orc_num_threads(); if (nt > 256) nt = 256;
  pw_ctx c = {G, n, d, ld, (double*)calloc((size_t)nt * n * n, sizeof(double))};
  orc_parallel_for(nblk, 4, pw_body, &c);
  memset(d2, 0, sizeof(double) * (size_t)n * n);
  for (int t = 0; t < nt; ++t)
    for (size_t e = 0; e < (size_t)n * n; ++e) d2[e] += c.loc[(size_t)t * n * n + e];
  free(c.loc);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j) d2[(size_t)j * n + i] = d2[(size_t)i * n + j];
}

static int cmp_double(const void* a, const void* b) {
  const double x = *(const double*)a, y = *(const double*)b;
  return (x > y) - (x < y);
}

static int visit_pos(int u) { return u == 1 ? 0 : (u == 0 ? 1 : u); }

typedef struct { const double* dist; int n; const unsigned char* alive; int take; double* score; } ks_ctx;
static int cmp_double(const void* a, const void* b);
static void ks_body(void* vc, int64_t u0, int64_t u1, int tid) {
  ks_ctx* c = (ks_ctx*)vc; (void)tid;
  const int n = c->n;
  double* tmp = (double*)malloc(sizeof(double) * (size_t)n);
  for (int u = (int)u0; u < (int)u1; ++u) {
    c->score[u] = NAN;
    if (c->alive && !c->alive[u]) continue;
    int k = 0;
    for (int v = 0; v < n; ++v)
      if (v != u && (!c->alive || c->alive[v])) tmp[k++] = c->dist[(size_t)u * n + v];
    qsort(tmp, (size_t)k, sizeof(double), cmp_double);
    double s = 0.0;
    for (int t = 0; t < c->take; ++t) s += tmp[t];
    c->score[u] = s;
  }
  free(tmp);
}

/* defences.py:23-42 on a dense table `dist` (n x n, float64 values of the fp32 distances), restricted to
 * alive users; keep = users_count - corrupted_count with Python slice semantics; strict-< scan in the
 * reference's visit order [1,0,2,...]; scores summed ascending in float64. */
int orc_krum_select(const double* dist, int n, const unsigned char* alive, int users_count, int corrupted_count,
                    double* scores_out) {
  int n_alive = 0;
  for (int u = 0; u < n; ++u) n_alive += alive ? alive[u] : 1;
  const int len = n_alive - 1;
  const int m = users_count - corrupted_count;
  const int take = m >= 0 ? (m < len ? m : len) : (len + m > 0 ? len + m : 0);
  double best = 1e20; int best_pos = -1, best_idx = -1;
  double* score = (double*)malloc(sizeof(double) * (size_t)n);
  ks_ctx kc = {dist, n, alive, take, score};
  orc_parallel_for(n, 8, ks_body, &kc);
  if (n >= 2) {   /* a 1-row table has no dict keys at all; a last survivor of a larger table still is a key */
    for (int u = 0; u < n; ++u) {
      if (alive && !alive[u]) continue;
      const double s = score[u];
      const int pos = visit_pos(u);
      if (!(s < 1e20)) continue;
      if (best_pos < 0 || s < best || (s == best && pos < best_pos)) { best = s; best_pos = pos; best_idx = u; }
    }
  }
  if (scores_out) memcpy(scores_out, score, sizeof(double) * (size_t)n);
  free(score);
  return best_idx;
}

/* Relative gap between the winning score and the runner-up of one scan (0 = exact tie, resolved by
 * visit order; inf when there is no runner-up).  Used by the tests to tell a selection the CUDA path
 * must reproduce from one inside fp32 noise. */
static double scan_margin(const double* score, int n, const unsigned char* alive, int best_idx) {
  if (best_idx < 0) return INFINITY;
  double second = INFINITY;
  for (int u = 0; u < n; ++u) {
    if (u == best_idx || (alive && !alive[u])) continue;
    if (score[u] < second) second = score[u];
  }
  const double best = score[best_idx];
  if (!(second < INFINITY) || best == 0.0) return INFINITY;
  return (second - best) / fabs(best);
}

/* orc_krum_select + margin */
int orc_krum_select_m(const double* dist, int n, const unsigned char* alive, int users_count, int corrupted_count,
                      double* margin_out) {
  double* score = (double*)malloc(sizeof(double) * (size_t)n);
  const int idx = orc_krum_select(dist, n, alive, users_count, corrupted_count, score);
  if (margin_out) *margin_out = scan_margin(score, n, alive, idx);
  free(score);
  return idx;
}

/* defences.py:57-68 — theta = n - 2f rounds of Krum with removal on one table.
 * margins_out (may be NULL): per-round relative gap to the runner-up. */
int orc_bulyan_select_m(const double* dist, int n, int f, int* sel_out, double* margins_out) {
  const int theta = n - 2 * f;
  unsigned char* alive = (unsigned char*)malloc((size_t)n);
  double* score = (double*)malloc(sizeof(double) * (size_t)n);
  memset(alive, 1, (size_t)n);
  for (int r = 0; r < theta; ++r) {
    const int idx = orc_krum_select(dist, n, alive, n - r, f, score);
    sel_out[r] = idx;
    if (margins_out) margins_out[r] = scan_margin(score, n, alive, idx);
    if (idx < 0) { free(alive); free(score); return r; }
    alive[idx] = 0;
  }
  free(alive); free(score);
  return theta;
}
int orc_bulyan_select(const double* dist, int n, int f, int* sel_out) {
  return orc_bulyan_select_m(dist, n, f, sel_out, NULL);
}

typedef struct { float key; float dev; int row; } kd_t;
static int cmp_float(const void* a, const void* b) {
  const float x = *(const float*)a, y = *(const float*)b;
  return (x > y) - (x < y);
}
static int cmp_kd(const void* a, const void* b) {
  const kd_t* x = (const kd_t*)a; const kd_t* y = (const kd_t*)b;
  if (x->key != y->key) return (x->key > y->key) - (x->key < y->key);
  return (x->row > y->row) - (x->row < y->row);          /* stable: earlier client first */
}

/* defences.py:44-52 — per column: fp32 median (even count: fl32((a+b))/2), fp32 deviations, the k of
 * smallest magnitude with ties in client order; mean of the kept deviations in float64, + med.
 * rows: optional row gather (Bulyan stage 2 passes its selection sequence).  out is float64. */
typedef struct { const float* G; int64_t ld; const int* rows; int n, keep; double* out; } tm_ctx;
static void tm_body(void* vc, int64_t c0, int64_t c1, int tid) {
  tm_ctx* t = (tm_ctx*)vc; (void)tid;
  const int n = t->n;
  float* col = (float*)malloc(sizeof(float) * (size_t)n);
  kd_t* kd = (kd_t*)malloc(sizeof(kd_t) * (size_t)n);
  for (int64_t c = c0; c < c1; ++c) {
    for (int r = 0; r < n; ++r) col[r] = t->G[(int64_t)(t->rows ? t->rows[r] : r) * t->ld + c];
    for (int r = 0; r < n; ++r) { kd[r].dev = col[r]; kd[r].row = r; }
    qsort(col, (size_t)n, sizeof(float), cmp_float);
    float med;
    if (n & 1) med = col[n / 2];
    else { const float s = col[n / 2 - 1] + col[n / 2]; med = s / 2.0f; }
    for (int r = 0; r < n; ++r) { const float dv = kd[r].dev - med; kd[r].dev = dv; kd[r].key = fabsf(dv); }
    qsort(kd, (size_t)n, sizeof(kd_t), cmp_kd);
    double s = 0.0;
    for (int r = 0; r < t->keep; ++r) s += (double)kd[r].dev;
    t->out[c] = t->keep > 0 ? s / t->keep + (double)med : NAN;
  }
  free(col); free(kd);
}
void orc_trimmed_mean(const float* G, int n_total, int64_t d, int64_t ld, const int* rows, int n, int corrupted_count,
                      double* out) {
  const int k0 = n - corrupted_count - 1;
  const int keep = k0 >= 0 ? (k0 < n ? k0 : n) : (n + k0 > 0 ? n + k0 : 0);
  (void)n_total;
  tm_ctx t = {G, ld, rows, n, keep, out};
  orc_parallel_for(d, 64, tm_body, &t);
}

/* defences.py:13-14 */
void orc_mean(const float* G, int n, int64_t d, int64_t ld, double* out) {
  for (int64_t c = 0; c < d; ++c) {
    double s = 0.0;
    for (int r = 0; r < n; ++r) s += (double)G[(int64_t)r * ld + c];
    out[c] = s / n;
  }
}

/* malicious.py:18-19,35 — mu, population sigma over the f malicious rows, crafted = mu - z*sigma (float64). */
void orc_alie(const float* G, int f, int64_t d, int64_t ld, double z, double* mu, double* sigma, double* crafted) {
  for (int64_t c = 0; c < d; ++c) {
    double s = 0.0;
    for (int r = 0; r < f; ++r) s += (double)G[(int64_t)r * ld + c];
    const double m = s / f;
    double v = 0.0;
    for (int r = 0; r < f; ++r) { const double t = (double)G[(int64_t)r * ld + c] - m; v += t * t; }
    const double sd = sqrt(v / f);
    if (mu) mu[c] = m;
    if (sigma) sigma[c] = sd;
    if (crafted) crafted[c] = m - z * sd;
  }
}
