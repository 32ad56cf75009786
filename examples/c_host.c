/* c_host.c -- a plain C host of libvecsearch_b200.so: no Python, no torch.
 *
 * Shows that the drop-in boundary (include/vecsearch_b200.h) is a self-contained C ABI: the calls
 * below are what a cgo / JNI / N-API / ctypes binding of the reference's collection.add and
 * collection.query (backend/app/main.py:735-740, 761-765) would make.  It builds a small corpus,
 * asks for the top-5 of a few queries through the HOST-buffer entry point and checks the answer
 * against a brute-force cosine loop written here in C (exit code 0 = identical rankings).
 *
 *   gcc -std=c99 -O2 -Iinclude examples/c_host.c -o /tmp/c_host \
 *       -Lmultimodal-image-similarity-search_b200 -lvecsearch_b200 -lm \
 *       -Wl,-rpath,$PWD/multimodal-image-similarity-search_b200
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "vecsearch_b200.h"

#define N 20000
#define D 512
#define B 4
#define K 5

static float frand(unsigned long long* s) { /* xorshift, uniform in (-1, 1) */
  *s ^= *s << 13;
  *s ^= *s >> 7;
  *s ^= *s << 17;
  return (float)((double)(*s >> 11) / 9007199254740992.0 * 2.0 - 1.0);
}

int main(void) {
  unsigned long long seed = 88172645463325252ULL;
  float* X = (float*)malloc(sizeof(float) * N * D);
  float* Q = (float*)malloc(sizeof(float) * B * D);
  if (!X || !Q) return 2;
  for (long i = 0; i < (long)N * D; ++i) X[i] = frand(&seed);
  for (long i = 0; i < (long)B * D; ++i) Q[i] = frand(&seed);

  vs_index_t* ix = NULL;
  if (vs_create(0, D, VS_F32, 0, &ix) != VS_OK) {
    fprintf(stderr, "vs_create: %s\n", vs_last_error());
    return 3;   /* e.g. no sm_100 GPU: the library has no CPU fallback */
  }
  int64_t first = -1;
  if (vs_add_host(ix, X, N, &first) != VS_OK || first != 0 || vs_count(ix) != N) {
    fprintf(stderr, "vs_add_host: %s\n", vs_last_error());
    return 4;
  }
  float scores[B * K];
  int64_t rows[B * K];
  if (vs_query_topk_host(ix, Q, B, K, NULL, VS_Q_AUTO, scores, rows) != VS_OK) {
    fprintf(stderr, "vs_query_topk_host: %s\n", vs_last_error());
    return 5;
  }
  int bad = 0;
  for (int b = 0; b < B; ++b) {
    double qn = 0;
    for (int e = 0; e < D; ++e) qn += (double)Q[b * D + e] * Q[b * D + e];
    qn = sqrt(qn);
    /* brute force: K passes of "best not yet taken" (ties -> lower row) */
    int taken[K];
    for (int j = 0; j < K; ++j) {
      double best = -2;
      int arg = -1;
      for (int i = 0; i < N; ++i) {
        int skip = 0;
        for (int t = 0; t < j; ++t) skip |= (taken[t] == i);
        if (skip) continue;
        double dot = 0, xn = 0;
        for (int e = 0; e < D; ++e) {
          dot += (double)Q[b * D + e] * X[(long)i * D + e];
          xn += (double)X[(long)i * D + e] * X[(long)i * D + e];
        }
        const double c = dot / (qn * sqrt(xn));
        if (c > best) {
          best = c;
          arg = i;
        }
      }
      taken[j] = arg;
      if (rows[b * K + j] != arg || fabs(scores[b * K + j] - best) > 1e-5) {
        fprintf(stderr, "query %d rank %d: got row %lld score %.7f, want row %d score %.7f\n", b, j,
                (long long)rows[b * K + j], scores[b * K + j], arg, best);
        ++bad;
      }
    }
  }
  printf("c_host: %d queries x top-%d over %d x %d f32 rows, %llu kernel launches, %s\n", B, K, N, D,
         (unsigned long long)vs_launch_count(), bad ? "MISMATCH" : "rankings identical to brute force");
  vs_destroy(ix);
  free(X);
  free(Q);
  return bad ? 1 : 0;
}
