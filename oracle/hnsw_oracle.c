/* hnsw_oracle.c -- TEST INFRASTRUCTURE, not product code (only tests/, tools/ and bench.py's
 * reference arm may build or call it).
 *
 * CPU restatement of the approximate index the reference actually queries: chromadb's collection
 * with metadata {"hnsw:space": "cosine"} (backend/app/utils.py:127-130) is an hnswlib index
 * (chroma-hnswlib fork; chromadb>=0.4.13 per requirements.txt:10 -- neither is vendored or
 * installable here, so this follows the PUBLISHED algorithm: Malkov & Yashunin, "Efficient and
 * robust approximate nearest neighbor search using Hierarchical Navigable Small World graphs",
 * Alg. 1-5, with hnswlib's concrete choices):
 *   - cosine space: vectors are L2-normalised at insert, distance = 1 - <a, b>;
 *   - M = 16 links per node on the upper layers, 2M = 32 on layer 0 (chroma defaults hnsw:M = 16,
 *     hnsw:construction_ef = 100, hnsw:search_ef = 10);
 *   - level of a new node = floor(-ln(U(0,1)) / ln(M));
 *   - insertion: greedy descent (ef = 1) to the node's level + 1, then on every layer at or below it
 *     an ef_construction beam search, neighbour SELECTION BY THE HEURISTIC (Alg. 4: keep a candidate
 *     only if it is closer to the new point than to every neighbour kept so far), bidirectional
 *     links, and the same heuristic to shrink a neighbour's list that overflows;
 *   - query: greedy descent to layer 1, beam search on layer 0 with ef = max(search_ef, k).
 * Its purpose here is ONE number the spec asks for: recall@k of that index against the exact result
 * (which is what the CUDA engine returns), plus its CPU query rate, on synthetic CLIP-shaped data.
 * The random level draws use our own generator, so graphs differ from hnswlib's link for link;
 * recall statistics do not depend on that.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  float d;
  int id;
} cand_t;

typedef struct {
  int dim, M, M0, efc, cap, n;
  int max_level, entry;
  double mult;
  uint64_t rng;
  float* vec;        /* [cap][dim] normalised */
  int* level;        /* [cap] */
  int* link0;        /* [cap][M0 + 1]: count, ids */
  int** linkup;      /* [cap] -> [level][M + 1] */
  uint32_t* visited; /* [cap] */
  uint32_t epoch;
  cand_t *heap_a, *heap_b, *tmp; /* scratch */
  int heap_cap;
} hnsw_t;

static float dist(const hnsw_t* h, const float* a, const float* b) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int i = 0;
  for (; i + 4 <= h->dim; i += 4) {
    s0 += a[i] * b[i];
    s1 += a[i + 1] * b[i + 1];
    s2 += a[i + 2] * b[i + 2];
    s3 += a[i + 3] * b[i + 3];
  }
  for (; i < h->dim; ++i) s0 += a[i] * b[i];
  return 1.0f - ((s0 + s1) + (s2 + s3));
}

static double urand(hnsw_t* h) { /* xorshift64* */
  h->rng ^= h->rng >> 12;
  h->rng ^= h->rng << 25;
  h->rng ^= h->rng >> 27;
  return (double)((h->rng * 2685821657736338717ULL) >> 11) / 9007199254740992.0;
}

/* binary heaps on cand_t; `far` = max-heap on d (result set W), `near` = min-heap on d (candidates C) */
static void push(cand_t* hp, int* n, cand_t c, int maxheap) {
  int i = (*n)++;
  hp[i] = c;
  while (i > 0) {
    int p = (i - 1) / 2;
    int up = maxheap ? hp[i].d > hp[p].d : hp[i].d < hp[p].d;
    if (!up) break;
    cand_t t = hp[i];
    hp[i] = hp[p];
    hp[p] = t;
    i = p;
  }
}
static cand_t pop(cand_t* hp, int* n, int maxheap) {
  cand_t top = hp[0];
  hp[0] = hp[--(*n)];
  int i = 0;
  for (;;) {
    int l = 2 * i + 1, r = l + 1, b = i;
    if (l < *n && (maxheap ? hp[l].d > hp[b].d : hp[l].d < hp[b].d)) b = l;
    if (r < *n && (maxheap ? hp[r].d > hp[b].d : hp[r].d < hp[b].d)) b = r;
    if (b == i) break;
    cand_t t = hp[i];
    hp[i] = hp[b];
    hp[b] = t;
    i = b;
  }
  return top;
}

static int* links(const hnsw_t* h, int id, int lvl) {
  return lvl == 0 ? h->link0 + (size_t)id * (h->M0 + 1) : h->linkup[id] + (size_t)(lvl - 1) * (h->M + 1);
}

/* Alg. 2: beam search on one layer; result (unsorted max-heap) left in heap_a, size returned */
static int search_layer(hnsw_t* h, const float* q, int ep, float ep_d, int ef, int lvl) {
  cand_t* W = h->heap_a;
  cand_t* C = h->heap_b;
  int nw = 0, nc = 0;
  if (++h->epoch == 0) {
    memset(h->visited, 0, (size_t)h->cap * sizeof(uint32_t));
    h->epoch = 1;
  }
  cand_t e = {ep_d, ep};
  push(W, &nw, e, 1);
  push(C, &nc, e, 0);
  h->visited[ep] = h->epoch;
  while (nc > 0) {
    cand_t c = pop(C, &nc, 0);
    if (c.d > W[0].d && nw >= ef) break;
    const int* l = links(h, c.id, lvl);
    for (int j = 1; j <= l[0]; ++j) {
      const int nb = l[j];
      if (h->visited[nb] == h->epoch) continue;
      h->visited[nb] = h->epoch;
      const float d = dist(h, q, h->vec + (size_t)nb * h->dim);
      if (nw < ef || d < W[0].d) {
        cand_t x = {d, nb};
        push(C, &nc, x, 0);
        push(W, &nw, x, 1);
        if (nw > ef) pop(W, &nw, 1);
      }
    }
  }
  return nw;
}

static int cmp_cand(const void* a, const void* b) {
  const cand_t *x = (const cand_t*)a, *y = (const cand_t*)b;
  return x->d < y->d ? -1 : x->d > y->d ? 1 : (x->id > y->id) - (x->id < y->id);
}

/* Alg. 4 (hnswlib getNeighborsByHeuristic2): cands sorted ascending by distance to the base point;
 * keep c iff it is closer to the base than to every already kept neighbour.  Returns kept count. */
static int select_heuristic(const hnsw_t* h, cand_t* cands, int n, int M, int* out) {
  int kept = 0;
  for (int i = 0; i < n && kept < M; ++i) {
    int ok = 1;
    for (int j = 0; j < kept; ++j) {
      const float d = dist(h, h->vec + (size_t)cands[i].id * h->dim, h->vec + (size_t)out[j] * h->dim);
      if (d < cands[i].d) {
        ok = 0;
        break;
      }
    }
    if (ok) out[kept++] = cands[i].id;
  }
  return kept;
}

hnsw_t* hnsw_create(int dim, int capacity, int M, int ef_construction, uint64_t seed) {
  hnsw_t* h = (hnsw_t*)calloc(1, sizeof(hnsw_t));
  if (!h) return NULL;
  h->dim = dim;
  h->M = M;
  h->M0 = 2 * M;
  h->efc = ef_construction;
  h->cap = capacity;
  h->max_level = -1;
  h->entry = -1;
  h->mult = 1.0 / log((double)M);
  h->rng = seed ? seed : 88172645463325252ULL;
  h->vec = (float*)malloc((size_t)capacity * dim * sizeof(float));
  h->level = (int*)calloc(capacity, sizeof(int));
  h->link0 = (int*)calloc((size_t)capacity * (h->M0 + 1), sizeof(int));
  h->linkup = (int**)calloc(capacity, sizeof(int*));
  h->visited = (uint32_t*)calloc(capacity, sizeof(uint32_t));
  h->heap_cap = capacity + 8;
  h->heap_a = (cand_t*)malloc((size_t)h->heap_cap * sizeof(cand_t));
  h->heap_b = (cand_t*)malloc((size_t)h->heap_cap * sizeof(cand_t));
  h->tmp = (cand_t*)malloc((size_t)h->heap_cap * sizeof(cand_t));
  if (!h->vec || !h->level || !h->link0 || !h->linkup || !h->visited || !h->heap_a || !h->heap_b || !h->tmp) return NULL;
  return h;
}

void hnsw_destroy(hnsw_t* h) {
  if (!h) return;
  for (int i = 0; i < h->n; ++i) free(h->linkup[i]);
  free(h->vec);
  free(h->level);
  free(h->link0);
  free(h->linkup);
  free(h->visited);
  free(h->heap_a);
  free(h->heap_b);
  free(h->tmp);
  free(h);
}

int hnsw_count(const hnsw_t* h) { return h->n; }

/* Alg. 1.  Returns the id (= insertion order) or -1 when full. */
int hnsw_add(hnsw_t* h, const float* x) {
  if (h->n >= h->cap) return -1;
  const int id = h->n;
  float* v = h->vec + (size_t)id * h->dim;
  double ss = 0.0;
  for (int i = 0; i < h->dim; ++i) ss += (double)x[i] * x[i];
  const float inv = ss > 0.0 ? (float)(1.0 / sqrt(ss)) : 0.f;   /* cosine space: normalise at insert */
  for (int i = 0; i < h->dim; ++i) v[i] = x[i] * inv;
  const int lvl = (int)(-log(1.0 - urand(h)) * h->mult);
  h->level[id] = lvl;
  if (lvl > 0) h->linkup[id] = (int*)calloc((size_t)lvl * (h->M + 1), sizeof(int));
  h->n++;
  if (h->entry < 0) {
    h->entry = id;
    h->max_level = lvl;
    return id;
  }
  int ep = h->entry;
  float ep_d = dist(h, v, h->vec + (size_t)ep * h->dim);
  for (int l = h->max_level; l > lvl; --l) {   /* greedy descent, ef = 1 */
    int changed = 1;
    while (changed) {
      changed = 0;
      const int* ln = links(h, ep, l);
      for (int j = 1; j <= ln[0]; ++j) {
        const float d = dist(h, v, h->vec + (size_t)ln[j] * h->dim);
        if (d < ep_d) {
          ep_d = d;
          ep = ln[j];
          changed = 1;
        }
      }
    }
  }
  int sel[128];
  for (int l = lvl < h->max_level ? lvl : h->max_level; l >= 0; --l) {
    const int Mmax = l == 0 ? h->M0 : h->M;
    int nw = search_layer(h, v, ep, ep_d, h->efc, l);
    memcpy(h->tmp, h->heap_a, (size_t)nw * sizeof(cand_t));
    qsort(h->tmp, nw, sizeof(cand_t), cmp_cand);
    const int ns = select_heuristic(h, h->tmp, nw, h->M, sel);
    int* mine = links(h, id, l);
    mine[0] = ns;
    for (int j = 0; j < ns; ++j) mine[1 + j] = sel[j];
    ep = h->tmp[0].id;   /* closest found becomes the next layer's entry */
    ep_d = h->tmp[0].d;
    for (int j = 0; j < ns; ++j) {   /* back links, shrinking with the same heuristic on overflow */
      const int nb = sel[j];
      int* ln = links(h, nb, l);
      if (ln[0] < Mmax) {
        ln[++ln[0]] = id;
      } else {
        const float* bv = h->vec + (size_t)nb * h->dim;
        cand_t pool[130];
        int np = 0;
        pool[np].d = dist(h, bv, v);
        pool[np++].id = id;
        for (int t = 1; t <= ln[0]; ++t) {
          pool[np].d = dist(h, bv, h->vec + (size_t)ln[t] * h->dim);
          pool[np++].id = ln[t];
        }
        qsort(pool, np, sizeof(cand_t), cmp_cand);
        int keep[128];
        const int nk = select_heuristic(h, pool, np, Mmax, keep);
        ln[0] = nk;
        for (int t = 0; t < nk; ++t) ln[1 + t] = keep[t];
      }
    }
  }
  if (lvl > h->max_level) {
    h->max_level = lvl;
    h->entry = id;
  }
  return id;
}

/* Alg. 5: k nearest (ascending distance) with beam width ef = max(ef_search, k); returns count. */
int hnsw_search(hnsw_t* h, const float* q, int k, int ef_search, int* out_ids, float* out_dist) {
  if (h->entry < 0) return 0;
  float* qn = (float*)malloc((size_t)h->dim * sizeof(float));
  double ss = 0.0;
  for (int i = 0; i < h->dim; ++i) ss += (double)q[i] * q[i];
  const float inv = ss > 0.0 ? (float)(1.0 / sqrt(ss)) : 0.f;
  for (int i = 0; i < h->dim; ++i) qn[i] = q[i] * inv;
  int ep = h->entry;
  float ep_d = dist(h, qn, h->vec + (size_t)ep * h->dim);
  for (int l = h->max_level; l > 0; --l) {
    int changed = 1;
    while (changed) {
      changed = 0;
      const int* ln = links(h, ep, l);
      for (int j = 1; j <= ln[0]; ++j) {
        const float d = dist(h, qn, h->vec + (size_t)ln[j] * h->dim);
        if (d < ep_d) {
          ep_d = d;
          ep = ln[j];
          changed = 1;
        }
      }
    }
  }
  const int ef = ef_search > k ? ef_search : k;
  int nw = search_layer(h, qn, ep, ep_d, ef, 0);
  memcpy(h->tmp, h->heap_a, (size_t)nw * sizeof(cand_t));
  qsort(h->tmp, nw, sizeof(cand_t), cmp_cand);
  const int m = nw < k ? nw : k;
  for (int i = 0; i < m; ++i) {
    out_ids[i] = h->tmp[i].id;
    out_dist[i] = h->tmp[i].d;
  }
  free(qn);
  return m;
}

/* bulk helpers for the ctypes driver */
int hnsw_add_many(hnsw_t* h, const float* x, int n) {
  for (int i = 0; i < n; ++i)
    if (hnsw_add(h, x + (size_t)i * h->dim) < 0) return i;
  return n;
}
void hnsw_search_many(hnsw_t* h, const float* q, int nq, int k, int ef_search, int* out_ids, float* out_dist) {
  for (int i = 0; i < nq; ++i) {
    const int m = hnsw_search(h, q + (size_t)i * h->dim, k, ef_search, out_ids + (size_t)i * k, out_dist + (size_t)i * k);
    for (int j = m; j < k; ++j) {
      out_ids[(size_t)i * k + j] = -1;
      out_dist[(size_t)i * k + j] = INFINITY;
    }
  }
}
