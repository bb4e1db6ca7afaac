/* oracle/knn_oracle.h -- TEST INFRASTRUCTURE ONLY (see knn_oracle.c). */
#ifndef KNN_ORACLE_H
#define KNN_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* space codes shared with tests (NOT with the product; the product has its own enum) */
enum {
  ORC_SPACE_L2 = 0,          /* sqrtf(sum (x-y)^2)                 space_lp.h:57-58 */
  ORC_SPACE_L2SQR = 1,       /* sum (x-y)^2 (L2SqrSIMD directly)   distcomp_lp.cc:304-365 */
  ORC_SPACE_COSINE = 2,      /* max(0, 1 - nsp)                    distcomp_scalar.cc:268-271 */
  ORC_SPACE_NEGDOT = 3,      /* -dot                               space_scalar.cc:60-68 */
  ORC_SPACE_L2SQR_SIFT = 4,  /* n1 + n2 - 2 dot, int32             distcomp_l2sqr_sift.cc:41-50 */
  ORC_SPACE_L1 = 5,          /* L1NormSIMD                         distcomp_lp.cc:190-251 */
  ORC_SPACE_LINF = 6,        /* LInfNormSIMD                       distcomp_lp.cc:77-139 */
  ORC_SPACE_ANGULAR = 7      /* acos(nsp)                          distcomp_scalar.cc:254-258 */
};

/* pairwise distances */
float orc_l2sqr(const float* a, const float* b, size_t d);
float orc_l2(const float* a, const float* b, size_t d);
float orc_norm_scalar_product(const float* a, const float* b, size_t d);
float orc_cosine(const float* a, const float* b, size_t d);
float orc_negdot(const float* a, const float* b, size_t d);
float orc_l1(const float* a, const float* b, size_t d);
float orc_linf(const float* a, const float* b, size_t d);
float orc_angular(const float* a, const float* b, size_t d);
int32_t orc_l2sqr_sift(const uint8_t* a, const uint8_t* b); /* 128-D */
/* HNSW optimized-index kernels (8-lane AVX summation order) */
float orc_hnsw_l2sqr(const float* a, const float* b, size_t d);
float orc_hnsw_dot(const float* a, const float* b, size_t d);

/* Sequential search (seqsearch.cc:144-150 + knnqueue.h:55-64 + nmslib_c.cpp:313-327).
 * data: [n][dim] float (or [n][128] uint8 for ORC_SPACE_L2SQR_SIFT), queries alike.
 * ext_ids may be NULL (position is the id).  Outputs are [nq][k]; counts[q] = min(k, n).
 * threads > 1 parallelises over queries with OpenMP (the result does not depend on it). */
int orc_seq_knn(int space, const void* data, size_t n, size_t dim, const int32_t* ext_ids,
                const void* queries, size_t nq, size_t k, int32_t* out_ids, float* out_dists,
                int32_t* out_counts, int threads);

/* HNSW optimized index (hnsw.cc:774-806 writer, :1025-1074 reader) */
typedef struct orc_hnsw orc_hnsw_t;
orc_hnsw_t* orc_hnsw_load(const char* path);
void orc_hnsw_free(orc_hnsw_t* h);
/* header fields, for tests */
void orc_hnsw_info(const orc_hnsw_t* h, uint64_t* total, uint64_t* dim, uint64_t* maxM,
                   uint64_t* maxM0, int32_t* maxlevel, uint32_t* enterpoint, int32_t* dist_func);
/* algo: 0 = hybrid (old iff ef >= 1000, hnsw.cc:724), 1 = v1merge, 2 = old.
 * n_eval (may be NULL): per-query number of distance evaluations, for roofline bytes. */
int orc_hnsw_knn(const orc_hnsw_t* h, const float* queries, size_t nq, size_t dim, size_t k,
                 size_t ef, int algo, int32_t* out_ids, float* out_dists, int32_t* out_counts,
                 int64_t* n_eval, int threads);

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
