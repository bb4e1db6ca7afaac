/* oracle/knn_oracle.c
 *
 * TEST INFRASTRUCTURE ONLY.  A plain-C, CPU restatement of the reference's
 * knnQuery / knnQueryBatch path (B-R-P/NMSLIB-ZIG), used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline leg as the CHECKER.
 * The product (nmslib_zig_b200/csrc) never includes, links or loads this file.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks every function
 * here against the UNMODIFIED reference compiled from /root/reference
 * (oracle/_ref/libnmslib_ref.so + libref_harness.so) and against the committed
 * golden vectors in tests/golden/ that were produced by that same reference
 * (tests/golden/make_golden.py), plus the three assertions the reference's own
 * tests hold for this path (lib.zig:1292-1299, :1419-1424).
 *
 * Every function cites the reference file:line it restates.  Written from the
 * published algorithm; no reference source text is reproduced.  Compiled with
 * -ffp-contract=off so that the summation order below is what actually runs.
 */
#include "knn_oracle.h"

#include <math.h>
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* Pairwise distances                                                         */
/* ------------------------------------------------------------------------- */

/* L2SqrSIMD, src/distcomp_lp.cc:304-365 (SSE2 branch): four fp32 lane
 * accumulators over the first 4*floor(d/4) elements (the 16-wide unrolled loop
 * and the 4-wide loop feed the same four lanes), horizontal sum t0+t1+t2+t3,
 * then a scalar tail. */
float orc_l2sqr(const float* a, const float* b, size_t d) {
  float lane[4] = {0.f, 0.f, 0.f, 0.f};
  size_t d4 = (d / 4) * 4, i;
  for (i = 0; i < d4; i += 4) {
    for (int j = 0; j < 4; ++j) {
      float df = a[i + j] - b[i + j];
      lane[j] = lane[j] + df * df;
    }
  }
  float res = lane[0] + lane[1] + lane[2] + lane[3];
  for (; i < d; ++i) {
    float df = a[i] - b[i];
    res += df * df;
  }
  return res;
}

/* L2NormSIMD, src/distcomp_lp.cc:368-371 (what space "l2" returns through
 * SpaceLp::HiddenDistance, space_lp.h:57-58). */
float orc_l2(const float* a, const float* b, size_t d) { return sqrtf(orc_l2sqr(a, b, d)); }

/* ScalarProductSIMD, src/distcomp_scalar.cc:194-245: same lane structure. */
static float orc_dot(const float* a, const float* b, size_t d) {
  float lane[4] = {0.f, 0.f, 0.f, 0.f};
  size_t d4 = (d / 4) * 4, i;
  for (i = 0; i < d4; i += 4)
    for (int j = 0; j < 4; ++j) lane[j] = lane[j] + a[i + j] * b[i + j];
  float res = lane[0] + lane[1] + lane[2] + lane[3];
  for (; i < d; ++i) res += a[i] * b[i];
  return res;
}

/* NormScalarProductSIMD, src/distcomp_scalar.cc:84-168.  a = data point (left),
 * b = query (right) -- query.cc:60.  Returns 0 when either squared norm is below
 * 2*FLT_MIN, otherwise clamp(sum / sqrt(n1) / sqrt(n2), -1, 1). */
float orc_norm_scalar_product(const float* a, const float* b, size_t d) {
  float ls[4] = {0, 0, 0, 0}, l1[4] = {0, 0, 0, 0}, l2[4] = {0, 0, 0, 0};
  size_t d4 = (d / 4) * 4, i;
  for (i = 0; i < d4; i += 4)
    for (int j = 0; j < 4; ++j) {
      ls[j] = ls[j] + a[i + j] * b[i + j];
      l1[j] = l1[j] + a[i + j] * a[i + j];
      l2[j] = l2[j] + b[i + j] * b[i + j];
    }
  float sum = ls[0] + ls[1] + ls[2] + ls[3];
  float n1 = l1[0] + l1[1] + l1[2] + l1[3];
  float n2 = l2[0] + l2[1] + l2[2] + l2[3];
  for (; i < d; ++i) {
    sum += a[i] * b[i];
    n1 += a[i] * a[i];
    n2 += b[i] * b[i];
  }
  const float eps = FLT_MIN * 2;
  if (n1 < eps || n2 < eps) return 0.f;
  float v = sum / sqrtf(n1) / sqrtf(n2);
  if (v > 1.f) v = 1.f;
  if (v < -1.f) v = -1.f;
  return v;
}

/* CosineSimilarity, src/distcomp_scalar.cc:268-271. */
float orc_cosine(const float* a, const float* b, size_t d) {
  float v = 1.f - orc_norm_scalar_product(a, b, d);
  return v > 0.f ? v : 0.f;
}

/* AngularDistance, src/distcomp_scalar.cc:254-258: acos of the clamped normalised scalar product. */
float orc_angular(const float* a, const float* b, size_t d) { return acosf(orc_norm_scalar_product(a, b, d)); }

/* L1NormSIMD<float>, src/distcomp_lp.cc:190-251 (the SSE2 branch, which is what x86 builds run): four
 * lanes of |a - b| (16-float unrolled, then 4-float groups), lanes summed t0+t1+t2+t3 into a double, the
 * scalar tail accumulated in that double. */
float orc_l1(const float* a, const float* b, size_t d) {
  float lane[4] = {0.f, 0.f, 0.f, 0.f};
  size_t d4 = (d / 4) * 4, i = 0;
  for (; i < d4; i += 4)
    for (int j = 0; j < 4; ++j) lane[j] = lane[j] + fabsf(a[i + j] - b[i + j]);
  double res = lane[0] + lane[1] + lane[2] + lane[3];
  for (; i < d; ++i) res += fabs(a[i] - b[i]);
  return (float)res;
}

/* LInfNormSIMD, src/distcomp_lp.cc:77-139: max |a_i - b_i| (exact whatever the order). */
float orc_linf(const float* a, const float* b, size_t d) {
  float res = 0.f;
  for (size_t i = 0; i < d; ++i) {
    const float v = fabsf(a[i] - b[i]);
    if (v > res) res = v;
  }
  return res;
}

/* SpaceNegativeScalarProduct::HiddenDistance, src/space/space_scalar.cc:60-68. */
float orc_negdot(const float* a, const float* b, size_t d) { return -orc_dot(a, b, d); }

/* l2SqrSIFTPrecomp*, src/distcomp_l2sqr_sift.cc:41-151 -- every variant is the same
 * exact integer: sum(a^2) + sum(b^2) - 2 sum(a b).  The reference stores the int32
 * norms behind the 128 bytes (space_l2sqr_sift.cc:141-148); we recompute them, which
 * is the same number. */
int32_t orc_l2sqr_sift(const uint8_t* a, const uint8_t* b) {
  int32_t na = 0, nb = 0, dot = 0;
  for (int i = 0; i < 128; ++i) {
    na += (int32_t)a[i] * a[i];
    nb += (int32_t)b[i] * b[i];
    dot += (int32_t)a[i] * b[i];
  }
  return na + nb - 2 * dot;
}

/* L2Sqr16Ext / L2SqrExt (AVX branch), include/method/hnsw_distfunc_opt_impl_inline.h:42-122.
 * Eight fp32 lanes over the first 16*floor(d/16) elements.  d % 16 == 0 -> the
 * eight lanes are summed left to right (L2Sqr16Ext, chosen at hnsw.cc:379-385);
 * otherwise lanes j and j+4 are folded, a 4-wide loop and a scalar tail follow. */
float orc_hnsw_l2sqr(const float* a, const float* b, size_t d) {
  float l8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  size_t d16 = (d / 16) * 16, d4 = (d / 4) * 4, i;
  for (i = 0; i < d16; i += 8)
    for (int j = 0; j < 8; ++j) {
      float df = a[i + j] - b[i + j];
      l8[j] = l8[j] + df * df;
    }
  if (d % 16 == 0) return l8[0] + l8[1] + l8[2] + l8[3] + l8[4] + l8[5] + l8[6] + l8[7];
  float l4[4];
  for (int j = 0; j < 4; ++j) l4[j] = l8[j] + l8[j + 4];
  for (; i < d4; i += 4)
    for (int j = 0; j < 4; ++j) {
      float df = a[i + j] - b[i + j];
      l4[j] = l4[j] + df * df;
    }
  float res = l4[0] + l4[1] + l4[2] + l4[3];
  for (; i < d; ++i) {
    float df = a[i] - b[i];
    res += df * df;
  }
  return res;
}

/* ScalarProduct (AVX branch), hnsw_distfunc_opt_impl_inline.h:124-173. */
float orc_hnsw_dot(const float* a, const float* b, size_t d) {
  float l8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  size_t d16 = (d / 16) * 16, d4 = (d / 4) * 4, i;
  for (i = 0; i < d16; i += 8)
    for (int j = 0; j < 8; ++j) l8[j] = l8[j] + a[i + j] * b[i + j];
  float l4[4];
  for (int j = 0; j < 4; ++j) l4[j] = l8[j] + l8[j + 4];
  for (; i < d4; i += 4)
    for (int j = 0; j < 4; ++j) l4[j] = l4[j] + a[i + j] * b[i + j];
  float res = l4[0] + l4[1] + l4[2] + l4[3];
  for (; i < d; ++i) res += a[i] * b[i];
  return res;
}

/* ------------------------------------------------------------------------- */
/* KNNQueue (include/knnqueue.h:28-81) + KNNQuery::CheckAndAddToResult        */
/* (src/knnquery.cc:66-80): bounded max-heap of (distance, object); a new     */
/* candidate enters when the queue is not full or when top.distance >         */
/* candidate (strict).  std::pair ordering makes the evicted element the      */
/* largest (distance, Object*); we stand in the insertion position for the    */
/* pointer (that is what the reference does in practice).      */
/* Distances are held as double so one queue serves float and int spaces.     */
/* ------------------------------------------------------------------------- */
typedef struct {
  double d;
  int64_t pos;
} qitem_t;

typedef struct {
  qitem_t* h;
  size_t n, k;
} knnq_t;

static int q_less(const qitem_t* x, const qitem_t* y) {
  return x->d < y->d || (x->d == y->d && x->pos < y->pos);
}
static void q_sift_up(knnq_t* q, size_t i) {
  while (i > 0) {
    size_t p = (i - 1) / 2;
    if (!q_less(&q->h[p], &q->h[i])) break;
    qitem_t t = q->h[p];
    q->h[p] = q->h[i];
    q->h[i] = t;
    i = p;
  }
}
static void q_sift_down(knnq_t* q, size_t i) {
  for (;;) {
    size_t l = 2 * i + 1, r = l + 1, m = i;
    if (l < q->n && q_less(&q->h[m], &q->h[l])) m = l;
    if (r < q->n && q_less(&q->h[m], &q->h[r])) m = r;
    if (m == i) break;
    qitem_t t = q->h[m];
    q->h[m] = q->h[i];
    q->h[i] = t;
    i = m;
  }
}
static void q_push(knnq_t* q, double d, int64_t pos) {
  if (q->n < q->k) { /* knnqueue.h:56-57 */
    q->h[q->n].d = d;
    q->h[q->n].pos = pos;
    q_sift_up(q, q->n++);
  } else if (q->k > 0 && q->h[0].d > d) { /* knnqueue.h:59 (strict) */
    q->h[0].d = d;
    q->h[0].pos = pos;
    q_sift_down(q, 0);
  }
}
/* extract_knn_results, nmslib_c.cpp:313-327: pop in descending order, reverse. */
static size_t q_drain_ascending(knnq_t* q, qitem_t* out) {
  size_t found = q->n;
  for (size_t j = found; j-- > 0;) {
    out[j] = q->h[0];
    q->h[0] = q->h[--q->n];
    q_sift_down(q, 0);
  }
  return found;
}

/* ------------------------------------------------------------------------- */
/* SeqSearch<dist_t>::Search(KNNQuery*), src/method/seqsearch.cc:144-150      */
/* ------------------------------------------------------------------------- */
int orc_seq_knn(int space, const void* data, size_t n, size_t dim, const int32_t* ext_ids,
                const void* queries, size_t nq, size_t k, int32_t* out_ids, float* out_dists,
                int32_t* out_counts, int threads) {
  if (space < ORC_SPACE_L2 || space > ORC_SPACE_ANGULAR || k == 0) return -1;
  if (space == ORC_SPACE_L2SQR_SIFT && dim != 128) return -2; /* space_l2sqr_sift.cc:137 CHECK */
  if (threads < 1) threads = 1;
  int failed = 0;
#pragma omp parallel num_threads(threads)
  {
    knnq_t q;
    q.h = (qitem_t*)malloc(sizeof(qitem_t) * k);
    qitem_t* sorted = (qitem_t*)malloc(sizeof(qitem_t) * k);
    q.k = k;
    if (!q.h || !sorted) {
#pragma omp atomic write
      failed = 1;
    } else {
#pragma omp for schedule(dynamic, 4)
      for (long long qi = 0; qi < (long long)nq; ++qi) {
        q.n = 0;
        if (space == ORC_SPACE_L2SQR_SIFT) {
          const uint8_t* qv = (const uint8_t*)queries + (size_t)qi * 128;
          for (size_t i = 0; i < n; ++i)
            q_push(&q, (double)orc_l2sqr_sift((const uint8_t*)data + i * 128, qv), (int64_t)i);
        } else {
          const float* qv = (const float*)queries + (size_t)qi * dim;
          const float* base = (const float*)data;
          for (size_t i = 0; i < n; ++i) {
            const float* x = base + i * dim; /* data point LEFT, query RIGHT (query.cc:60) */
            float d;
            switch (space) {
              case ORC_SPACE_L2: d = orc_l2(x, qv, dim); break;
              case ORC_SPACE_L2SQR: d = orc_l2sqr(x, qv, dim); break;
              case ORC_SPACE_COSINE: d = orc_cosine(x, qv, dim); break;
              case ORC_SPACE_L1: d = orc_l1(x, qv, dim); break;
              case ORC_SPACE_LINF: d = orc_linf(x, qv, dim); break;
              case ORC_SPACE_ANGULAR: d = orc_angular(x, qv, dim); break;
              default: d = orc_negdot(x, qv, dim); break;
            }
            q_push(&q, (double)d, (int64_t)i);
          }
        }
        size_t found = q_drain_ascending(&q, sorted);
        out_counts[qi] = (int32_t)found;
        for (size_t j = 0; j < k; ++j) {
          if (j < found) {
            out_ids[(size_t)qi * k + j] =
                ext_ids ? ext_ids[sorted[j].pos] : (int32_t)sorted[j].pos;
            out_dists[(size_t)qi * k + j] = (float)sorted[j].d; /* nmslib_c.cpp:317 */
          } else {
            out_ids[(size_t)qi * k + j] = -1;
            out_dists[(size_t)qi * k + j] = INFINITY;
          }
        }
      }
    }
    free(q.h);
    free(sorted);
  }
  return failed ? -3 : 0;
}

/* ------------------------------------------------------------------------- */
/* HNSW optimized index: file format written by Hnsw::SaveIndex /             */
/* SaveOptimizedIndex (src/method/hnsw.cc:748-806), read as in                */
/* LoadOptimizedIndex (:1025-1074).  Little-endian, unpadded.                 */
/* ------------------------------------------------------------------------- */
struct orc_hnsw {
  uint32_t total;
  uint64_t mem_per_obj, off_level0, off_data;
  int32_t maxlevel;
  uint32_t enterpoint;
  uint64_t maxM, maxM0;
  int32_t dist_func; /* hnsw.h:50-58: 1 L2Sqr16Ext 2 L2SqrExt 3 NormCosine 4 NegDot */
  uint64_t search_method;
  char* level0;      /* total * mem_per_obj bytes */
  char** links;      /* per node: level*(maxM+1)*4 bytes or NULL */
  uint64_t dim;
};

static int rd(FILE* f, void* p, size_t n) { return fread(p, 1, n, f) == n ? 0 : -1; }

orc_hnsw_t* orc_hnsw_load(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) return NULL;
  orc_hnsw_t* h = (orc_hnsw_t*)calloc(1, sizeof(*h));
  uint32_t optim = 0;
  int bad = rd(f, &optim, 4);
  if (!bad && optim != 1) bad = 1; /* hnsw.cc:756: only the optimized flat index */
  bad |= rd(f, &h->total, 4) | rd(f, &h->mem_per_obj, 8) | rd(f, &h->off_level0, 8) |
         rd(f, &h->off_data, 8) | rd(f, &h->maxlevel, 4) | rd(f, &h->enterpoint, 4) |
         rd(f, &h->maxM, 8) | rd(f, &h->maxM0, 8) | rd(f, &h->dist_func, 4) |
         rd(f, &h->search_method, 8);
  if (bad) goto fail;
  h->dim = (h->off_level0 - 16) / 4; /* hnsw.cc:377 vectorlength_ */
  h->level0 = (char*)malloc((size_t)h->total * h->mem_per_obj + 64);
  h->links = (char**)calloc(h->total, sizeof(char*));
  if (!h->level0 || !h->links) goto fail;
  if (rd(f, h->level0, (size_t)h->total * h->mem_per_obj)) goto fail;
  for (uint32_t i = 0; i < h->total; ++i) {
    uint32_t sz; /* SIZEMASS_TYPE = unsigned int */
    if (rd(f, &sz, 4)) goto fail;
    if (sz) {
      h->links[i] = (char*)malloc(sz);
      if (!h->links[i] || rd(f, h->links[i], sz)) goto fail;
    }
  }
  fclose(f);
  return h;
fail:
  fclose(f);
  orc_hnsw_free(h);
  return NULL;
}

void orc_hnsw_free(orc_hnsw_t* h) {
  if (!h) return;
  if (h->links)
    for (uint32_t i = 0; i < h->total; ++i) free(h->links[i]);
  free(h->links);
  free(h->level0);
  free(h);
}

void orc_hnsw_info(const orc_hnsw_t* h, uint64_t* total, uint64_t* dim, uint64_t* maxM,
                   uint64_t* maxM0, int32_t* maxlevel, uint32_t* enterpoint, int32_t* dist_func) {
  *total = h->total;
  *dim = h->dim;
  *maxM = h->maxM;
  *maxM0 = h->maxM0;
  *maxlevel = h->maxlevel;
  *enterpoint = h->enterpoint;
  *dist_func = h->dist_func;
}

/* fstdistfunc_ selection, hnsw.cc:70-81 + getDistFunc */
static float hnsw_dist(const orc_hnsw_t* h, const float* q, const float* x) {
  switch (h->dist_func) {
    case 1:
    case 2: return orc_hnsw_l2sqr(q, x, h->dim);
    case 3: { /* NormCosine, hnsw.cc:78-81 */
      float s = orc_hnsw_dot(q, x, h->dim);
      if (s > 1.f) s = 1.f;
      if (s < -1.f) s = -1.f;
      float v = 1.f - s;
      return v > 0.f ? v : 0.f;
    }
    default: return -orc_hnsw_dot(q, x, h->dim); /* NegativeDotProduct, hnsw.cc:70-73 */
  }
}
static const float* hnsw_vec(const orc_hnsw_t* h, uint32_t node) {
  return (const float*)(h->level0 + (size_t)node * h->mem_per_obj + h->off_data + 16);
}
static const int32_t* hnsw_links0(const orc_hnsw_t* h, uint32_t node) {
  return (const int32_t*)(h->level0 + (size_t)node * h->mem_per_obj + h->off_level0);
}
static int32_t hnsw_ext_id(const orc_hnsw_t* h, uint32_t node) {
  int32_t id;
  memcpy(&id, h->level0 + (size_t)node * h->mem_per_obj + h->off_data, 4);
  return id;
}

/* SortArrBI<float,int>, include/sort_arr_bi.h:30-216 */
typedef struct {
  float key;
  int used;
  int32_t data;
} sitem_t;
typedef struct {
  sitem_t* v;
  size_t cap, n;
} sarr_t;

/* push_or_replace_non_empty_exp, sort_arr_bi.h:159-199: returns the insertion index
 * (n when the item was dropped because the array is full and key >= last key). */
static size_t sarr_push_exp(sarr_t* s, float key, int32_t data) {
  size_t curr = s->n - 1;
  if (s->v[curr].key <= key) {
    if (s->n < s->cap) {
      s->v[s->n].used = 0;
      s->v[s->n].key = key;
      s->v[s->n].data = data;
      return s->n++;
    }
    return s->n;
  }
  size_t prev = curr, d = 1;
  while (curr > 0 && s->v[curr].key > key) {
    prev = curr;
    curr -= d;
    d *= 2;
    if (d > curr) d = curr;
  }
  if (curr < prev) { /* std::lower_bound over [curr, prev) on key */
    size_t lo = curr, hi = prev;
    while (lo < hi) {
      size_t mid = lo + (hi - lo) / 2;
      if (s->v[mid].key < key) lo = mid + 1; else hi = mid;
    }
    curr = lo;
  }
  if (s->n < s->cap) s->n++;
  if (s->n - (1 + curr) > 0)
    memmove(&s->v[curr + 1], &s->v[curr], (s->n - (1 + curr)) * sizeof(sitem_t));
  s->v[curr].used = 0;
  s->v[curr].key = key;
  s->v[curr].data = data;
  return curr;
}

/* merge_with_sorted_items, sort_arr_bi.h:116-155 (only reached when more than 100
 * candidates come out of one expansion, i.e. maxM0 > 100).  Stable merge: existing
 * items precede equal new ones, like std::inplace_merge. */
static size_t sarr_merge(sarr_t* s, const sitem_t* items, size_t qty, sitem_t* tmp) {
  if (!qty) return s->n;
  if (qty > s->cap) qty = s->cap;
  size_t left = s->cap - s->n, keep_old, take_new;
  if (left >= qty) {
    keep_old = s->n;
    take_new = qty;
  } else {
    size_t rm = 0;
    while (qty > left + rm && s->n > rm && items[left + rm].key < s->v[s->n - rm - 1].key) rm++;
    keep_old = s->n - rm;
    take_new = left + rm;
  }
  size_t i = 0, j = 0, o = 0;
  while (i < keep_old || j < take_new) {
    if (j >= take_new || (i < keep_old && !(items[j].key < s->v[i].key))) tmp[o++] = s->v[i++];
    else tmp[o++] = items[j++];
  }
  memcpy(s->v, tmp, o * sizeof(sitem_t));
  s->n = o;
  size_t ret = 0;
  while (ret < s->n && s->v[ret].used) ++ret;
  return ret;
}

static int sitem_cmp(const void* x, const void* y) {
  float a = ((const sitem_t*)x)->key, b = ((const sitem_t*)y)->key;
  return (a > b) - (a < b);
}

/* NormalizeVect, include/method/hnsw.h:486-497 */
static void normalize_vect(float* v, size_t d) {
  float sum = 0;
  for (size_t i = 0; i < d; ++i) sum += v[i] * v[i];
  if (sum != 0.0f) {
    sum = 1 / sqrtf(sum);
    for (size_t i = 0; i < d; ++i) v[i] *= sum;
  }
}

/* greedy descent through the upper layers, hnsw_distfunc_opt.cc:168-198 */
static uint32_t hnsw_descend(const orc_hnsw_t* h, const float* q, float* curdist_out,
                             int64_t* evals) {
  uint32_t cur = h->enterpoint;
  float curdist = hnsw_dist(h, q, hnsw_vec(h, cur));
  ++*evals;
  for (int lvl = h->maxlevel; lvl > 0; --lvl) {
    int changed = 1;
    while (changed) {
      changed = 0;
      const int32_t* lk = (const int32_t*)(h->links[cur] + (h->maxM + 1) * (size_t)(lvl - 1) * 4);
      int size = lk[0];
      for (int j = 1; j <= size; ++j) {
        uint32_t t = (uint32_t)lk[j];
        float d = hnsw_dist(h, q, hnsw_vec(h, t));
        ++*evals;
        if (d < curdist) {
          curdist = d;
          cur = t;
          changed = 1;
        }
      }
    }
  }
  *curdist_out = curdist;
  return cur;
}

/* Hnsw::SearchV1Merge, src/method/hnsw_distfunc_opt.cc:152-283 */
static void hnsw_search_v1merge(const orc_hnsw_t* h, const float* q, size_t k, size_t ef,
                                uint8_t* visited, knnq_t* res, int64_t* evals) {
  float curdist;
  uint32_t cur = hnsw_descend(h, q, &curdist, evals);
  sarr_t s;
  s.cap = ef > k ? ef : k;
  s.n = 0;
  s.v = (sitem_t*)malloc(sizeof(sitem_t) * (s.cap + 1));
  size_t buf_cap = 1 + (h->maxM > h->maxM0 ? h->maxM : h->maxM0);
  sitem_t* buf = (sitem_t*)malloc(sizeof(sitem_t) * buf_cap);
  sitem_t* tmp = (sitem_t*)malloc(sizeof(sitem_t) * (s.cap + 1));
  s.v[0].used = 0; /* push_unsorted_grow, sort_arr_bi.h:61-67 */
  s.v[0].key = curdist;
  s.v[0].data = (int32_t)cur;
  s.n = 1;
  size_t curr_elem = 0;
  visited[cur] = 1;
  while (curr_elem < (s.n < ef ? s.n : ef)) {
    s.v[curr_elem].used = 1;
    uint32_t node = (uint32_t)s.v[curr_elem].data;
    ++curr_elem;
    size_t qty = 0;
    float top_key = s.v[s.n - 1].key;
    const int32_t* lk = hnsw_links0(h, node);
    int size = lk[0];
    for (int j = 1; j <= size; ++j) {
      uint32_t t = (uint32_t)lk[j];
      if (!visited[t]) {
        visited[t] = 1;
        float d = hnsw_dist(h, q, hnsw_vec(h, t));
        ++*evals;
        if (d < top_key || s.n < ef) {
          buf[qty].key = d;
          buf[qty].used = 0;
          buf[qty].data = (int32_t)t;
          ++qty;
        }
      }
    }
    if (qty) {
      qsort(buf, qty, sizeof(sitem_t), sitem_cmp);
      if (qty > 100) { /* MERGE_BUFFER_ALGO_SWITCH_THRESHOLD, hnsw_distfunc_opt.cc:35 */
        size_t ins = sarr_merge(&s, buf, qty, tmp);
        if (ins < curr_elem) curr_elem = ins;
      } else {
        for (size_t ii = 0; ii < qty; ++ii) {
          size_t ins = sarr_push_exp(&s, buf[ii].key, buf[ii].data);
          if (ins < curr_elem) curr_elem = ins;
        }
      }
    }
    while (curr_elem < s.n && s.v[curr_elem].used) ++curr_elem;
  }
  for (size_t i = 0; i < k && i < s.n; ++i) q_push(res, (double)s.v[i].key, s.v[i].data);
  free(s.v);
  free(buf);
  free(tmp);
}

/* tiny binary heaps for SearchOld (std::priority_queue<EvaluatedMSWNodeInt>) */
typedef struct {
  float d;
  int32_t e;
} hitem_t;
typedef struct {
  hitem_t* h;
  size_t n, cap;
} heap_t;
static void heap_push(heap_t* p, float d, int32_t e) {
  if (p->n == p->cap) {
    p->cap = p->cap ? p->cap * 2 : 64;
    p->h = (hitem_t*)realloc(p->h, p->cap * sizeof(hitem_t));
  }
  size_t i = p->n++;
  p->h[i].d = d;
  p->h[i].e = e;
  while (i > 0) {
    size_t up = (i - 1) / 2;
    if (!(p->h[up].d < p->h[i].d)) break;
    hitem_t t = p->h[up];
    p->h[up] = p->h[i];
    p->h[i] = t;
    i = up;
  }
}
static void heap_pop(heap_t* p) {
  p->h[0] = p->h[--p->n];
  size_t i = 0;
  for (;;) {
    size_t l = 2 * i + 1, r = l + 1, m = i;
    if (l < p->n && p->h[m].d < p->h[l].d) m = l;
    if (r < p->n && p->h[m].d < p->h[r].d) m = r;
    if (m == i) break;
    hitem_t t = p->h[m];
    p->h[m] = p->h[i];
    p->h[i] = t;
    i = m;
  }
}

/* Hnsw::SearchOld, src/method/hnsw_distfunc_opt.cc:46-150 */
static void hnsw_search_old(const orc_hnsw_t* h, const float* q, size_t ef, uint8_t* visited,
                            knnq_t* res, int64_t* evals) {
  float curdist;
  uint32_t cur = hnsw_descend(h, q, &curdist, evals);
  heap_t cand = {0, 0, 0}, closest = {0, 0, 0};
  heap_push(&cand, -curdist, (int32_t)cur);
  heap_push(&closest, curdist, (int32_t)cur);
  q_push(res, (double)curdist, cur);
  visited[cur] = 1;
  while (cand.n) {
    hitem_t ev = cand.h[0];
    float lower = closest.h[0].d;
    if (-ev.d > lower) break;
    heap_pop(&cand);
    const int32_t* lk = hnsw_links0(h, (uint32_t)ev.e);
    int size = lk[0];
    for (int j = 1; j <= size; ++j) {
      uint32_t t = (uint32_t)lk[j];
      if (!visited[t]) {
        visited[t] = 1;
        float d = hnsw_dist(h, q, hnsw_vec(h, t));
        ++*evals;
        if (closest.h[0].d > d || closest.n < ef) {
          heap_push(&cand, -d, (int32_t)t);
          q_push(res, (double)d, t);
          heap_push(&closest, d, (int32_t)t);
          if (closest.n > ef) heap_pop(&closest);
        }
      }
    }
  }
  free(cand.h);
  free(closest.h);
}

/* Hnsw<float>::Search(KNNQuery*) dispatch, src/method/hnsw.cc:717-746, followed by
 * the result extraction of nmslib_c.cpp:313-327 (ids = external ids stored in the
 * 16-byte Object header of each level-0 record). */
int orc_hnsw_knn(const orc_hnsw_t* h, const float* queries, size_t nq, size_t dim, size_t k,
                 size_t ef, int algo, int32_t* out_ids, float* out_dists, int32_t* out_counts,
                 int64_t* n_eval, int threads) {
  if (!h || dim != h->dim || k == 0 || ef == 0) return -1;
  if (threads < 1) threads = 1;
  int use_old = (algo == 2) || (algo == 0 && ef >= 1000);
#pragma omp parallel num_threads(threads)
  {
    uint8_t* visited = (uint8_t*)malloc(h->total);
    float* qcopy = (float*)malloc(sizeof(float) * dim);
    knnq_t res;
    res.h = (qitem_t*)malloc(sizeof(qitem_t) * k);
    res.k = k;
    qitem_t* sorted = (qitem_t*)malloc(sizeof(qitem_t) * k);
#pragma omp for schedule(dynamic, 4)
    for (long long qi = 0; qi < (long long)nq; ++qi) {
      memset(visited, 0, h->total); /* VisitedList epoch == a fresh array, hnsw.h:569-591 */
      memcpy(qcopy, queries + (size_t)qi * dim, sizeof(float) * dim);
      if (h->dist_func == 3) normalize_vect(qcopy, dim); /* hnsw_distfunc_opt.cc:160-162 */
      res.n = 0;
      int64_t evals = 0;
      if (use_old) hnsw_search_old(h, qcopy, ef, visited, &res, &evals);
      else hnsw_search_v1merge(h, qcopy, k, ef, visited, &res, &evals);
      size_t found = q_drain_ascending(&res, sorted);
      out_counts[qi] = (int32_t)found;
      if (n_eval) n_eval[qi] = evals;
      for (size_t j = 0; j < k; ++j) {
        if (j < found) {
          out_ids[(size_t)qi * k + j] = hnsw_ext_id(h, (uint32_t)sorted[j].pos);
          out_dists[(size_t)qi * k + j] = (float)sorted[j].d;
        } else {
          out_ids[(size_t)qi * k + j] = -1;
          out_dists[(size_t)qi * k + j] = INFINITY;
        }
      }
    }
    free(visited);
    free(qcopy);
    free(res.h);
    free(sorted);
  }
  return 0;
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
