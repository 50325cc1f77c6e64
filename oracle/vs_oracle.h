/* oracle/vs_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's CPU algorithm for the flat top-K / range / batch-iterator
 * hot path and its distance kernels. It is the checker for the CUDA path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it. The product libraries
 * (libvsgpu.so / libvecsim_b200.so) never link, load or call anything in this directory.
 *
 * Parity status: PINNED. tests/test_oracle_vs_reference.py checks every function below bit-for-bit
 * against oracle/_ref/libvecsim_ref.so (the unmodified reference compiled by oracle/Makefile) and
 * against the known-answer values of the reference's own unit tests (tests/golden/).
 *
 * Numbering of `type` / `metric` follows VecSimType / VecSimMetric
 * (reference src/VecSim/vec_sim_common.h:60-69,84-88).
 */
#ifndef VS_ORACLE_H
#define VS_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { VSO_FLOAT32 = 0, VSO_FLOAT64 = 1, VSO_BFLOAT16 = 2, VSO_FLOAT16 = 3, VSO_INT8 = 4, VSO_UINT8 = 5 };
enum { VSO_L2 = 0, VSO_IP = 1, VSO_COSINE = 2 };
enum { VSO_BY_SCORE = 0, VSO_BY_ID = 1 };

/* Which x86 tier of the reference the floating-point kernels restate.
 *   VSO_TIER_AVX512      what an AVX512F/BW/VL/VNNI/VBMI2/BF16 host dispatches, with the
 *                        half-precision-accumulating AVX512_FP16 kernels masked off (App. A4).
 *   VSO_TIER_AVX512_NOBF16  same, but bf16 IP through the AVX512BW+VBMI2 kernel (no vdpbf16ps).
 *   VSO_TIER_NAIVE       the scalar functions of spaces/IP/IP.cpp, spaces/L2/L2.cpp.
 */
enum { VSO_TIER_AVX512 = 0, VSO_TIER_AVX512_NOBF16 = 1, VSO_TIER_NAIVE = 2 };
void vso_set_tier(int tier);
int vso_get_tier(void);

/* bytes of one stored row / processed query (utils/vec_utils.cpp:296-302). */
size_t vso_stored_size(int type, int metric, size_t dim);

/* dist(stored, query): both blobs in *processed* form (normalised / norm appended). */
double vso_distance(int type, int metric, size_t dim, const void *a, const void *b);

/* In-place normalisation (spaces/normalize/normalize_naive.h:23-88). int8/uint8 need dim+4 bytes. */
void vso_normalize(int type, size_t dim, void *blob);

/* Type conversions (types/bfloat16.h:23-39, types/float16.h:33-117). */
uint16_t vso_f32_to_bf16(float f);
float vso_bf16_to_f32(uint16_t h);
uint16_t vso_f32_to_fp16(float f);
float vso_fp16_to_f32(uint16_t h);

/* ---- flat index (reference: algorithms/brute_force) ---- */
typedef struct vso_flat vso_flat;
vso_flat *vso_flat_new(int type, size_t dim, int metric, int multi, size_t block_size);
void vso_flat_free(vso_flat *f);
int vso_flat_add(vso_flat *f, const void *blob, size_t label);    /* returns #new labels... see .c */
int vso_flat_delete(vso_flat *f, size_t label);                   /* returns #deleted vectors */
size_t vso_flat_size(const vso_flat *f);
size_t vso_flat_label_count(const vso_flat *f);
/* raw access for tests that want to feed the same rows to the GPU store */
const void *vso_flat_row(const vso_flat *f, size_t id);
size_t vso_flat_label_of(const vso_flat *f, size_t id);

/* returns number of results written (<= k). code: 0 OK, 1 timed out. timeout!=0 emulates a
 * timeout callback returning 1. */
size_t vso_flat_topk(const vso_flat *f, const void *query, size_t k, int order, int timeout,
                     size_t *labels, double *scores, int *code);
/* returns total number of results (only `cap` written); -1 on the reference's exception cases. */
long vso_flat_range(const vso_flat *f, const void *query, double radius, int order, int timeout,
                    size_t cap, size_t *labels, double *scores, int *code);
double vso_flat_distance_from(const vso_flat *f, size_t label, const void *query);

typedef struct vso_bi vso_bi;
vso_bi *vso_bi_new(const vso_flat *f, const void *query);
size_t vso_bi_next(vso_bi *it, size_t n, int order, size_t *labels, double *scores, int *code);
int vso_bi_has_next(const vso_bi *it);
void vso_bi_reset(vso_bi *it);
void vso_bi_free(vso_bi *it);

/* ---- HNSW (reference: algorithms/hnsw), vs_oracle_hnsw.c: single-threaded build + top-k + range ---- */
typedef struct vso_hnsw vso_hnsw;
vso_hnsw *vso_hnsw_new(int type, size_t dim, int metric, size_t M, size_t ef_construction, size_t ef_runtime,
                       double epsilon);
void vso_hnsw_free(vso_hnsw *g);
void vso_hnsw_add(vso_hnsw *g, const void *blob, size_t label); /* raw caller blob; label must be new */
size_t vso_hnsw_size(const vso_hnsw *g);
/* HNSWIndex_Multi (algorithms/hnsw/hnsw_multi.h): labels may repeat; top-k / range return each label once with its best
 * score. Call before the first query; the graph itself is built exactly as for a single-value index. */
void vso_hnsw_set_multi(vso_hnsw *g, int multi);
void vso_hnsw_mark_deleted(vso_hnsw *g, size_t id, int deleted);
void vso_hnsw_info(const vso_hnsw *g, long *entry, long *max_level); /* -1/-1 when empty */
uint32_t vso_hnsw_level(const vso_hnsw *g, size_t id);
size_t vso_hnsw_links(const vso_hnsw *g, size_t id, size_t level, uint32_t *out);
size_t vso_hnsw_dist_count(const vso_hnsw *g);
/* ascending (score, label); ef_runtime 0 = index default; returns the number of results (<= k) */
size_t vso_hnsw_topk(vso_hnsw *g, const void *query, size_t k, size_t ef_runtime, size_t *labels, double *scores);
/* sorted by (score, label); epsilon 0 = index default; returns the total found, writes <= cap */
size_t vso_hnsw_range(vso_hnsw *g, const void *query, double radius, double epsilon, size_t cap, size_t *labels,
                      double *scores);

typedef struct vso_hnsw_bi vso_hnsw_bi;
vso_hnsw_bi *vso_hnsw_bi_new(vso_hnsw *g, const void *query, size_t ef_runtime);
void vso_hnsw_bi_free(vso_hnsw_bi *it);
void vso_hnsw_bi_reset(vso_hnsw_bi *it);
int vso_hnsw_bi_has_next(const vso_hnsw_bi *it);
size_t vso_hnsw_bi_next(vso_hnsw_bi *it, size_t n_res, size_t label_count, size_t *labels, double *scores);

#ifdef __cplusplus
}
#endif
#endif
