/* oracle/vs_oracle.c — TEST INFRASTRUCTURE ONLY (see vs_oracle.h).
 *
 * CPU restatement, in plain C, of the reference's flat-index hot path. Every function cites the
 * reference file:line (relative to /root/reference/src/VecSim) whose behaviour it restates.
 * Nothing here is shared with, linked into or called from the product path.
 *
 * Parity: PINNED against oracle/_ref (the unmodified reference) and the reference's own
 * known-answer tests by tests/test_oracle_vs_reference.py and tests/test_golden.py.
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off -mfma  (FMA only where fmaf()/fma() is written).
 */
#include "vs_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static int g_tier = VSO_TIER_AVX512;
void vso_set_tier(int tier) { g_tier = tier; }
int vso_get_tier(void) { return g_tier; }

/* ------------------------------------------------------------------------------------------ */
/* type conversions                                                                           */

/* types/bfloat16.h:23-30 — round-to-nearest-even on the upper 16 bits. */
uint16_t vso_f32_to_bf16(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    u += ((u >> 16) & 1u) + 0x7FFFu;
    return (uint16_t)(u >> 16);
}
/* types/bfloat16.h:32-39 — the 16 bits become the high half of an fp32. */
float vso_bf16_to_f32(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static float bits_f(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static uint32_t f_bits(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}
/* types/float16.h:33-50 — exact widening (all fp16 values are representable in fp32). */
float vso_fp16_to_f32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t em = h & 0x7fffu;
    uint32_t exp = em & 0x7c00u;
    uint32_t out;
    if (exp == 0x7c00u) { /* inf / nan */
        out = (em << 13) + ((uint32_t)(127 - 15) << 23) + ((uint32_t)(128 - 16) << 23);
    } else if (exp == 0) { /* zero / subnormal: renormalise through an fp32 subtract */
        uint32_t o = (em << 13) + ((uint32_t)(127 - 15) << 23);
        out = f_bits(bits_f(o + (1u << 23)) - bits_f(113u << 23));
    } else {
        out = (em << 13) + ((uint32_t)(127 - 15) << 23);
    }
    return bits_f(out | sign);
}
/* types/float16.h:60-117 — note: drops the low 12 mantissa bits, then rounds half up at bit 12. */
uint16_t vso_f32_to_fp16(float input) {
    uint32_t fint = f_bits(input);
    uint32_t sign = fint & 0x80000000u;
    fint ^= sign;
    const uint32_t f32infty = 255u << 23;
    const uint32_t round_mask = ~0xfffu;
    uint32_t o = (fint > f32infty) ? 0x7e00u : 0x7c00u;
    float fscale = bits_f(fint & round_mask) * bits_f(15u << 23);
    float cap = bits_f((31u << 23) - 0x1000u);
    if (cap < fscale) fscale = cap; /* std::min(fscale, cap) */
    int32_t fint2 = (int32_t)(f_bits(fscale) - round_mask);
    if (fint < f32infty) o = (uint32_t)(fint2 >> 13);
    return (uint16_t)(o | (sign >> 16));
}

static size_t type_size(int type) {
    switch (type) {
    case VSO_FLOAT32: return 4;
    case VSO_FLOAT64: return 8;
    case VSO_BFLOAT16:
    case VSO_FLOAT16: return 2;
    default: return 1;
    }
}
/* utils/vec_utils.cpp:296-302 */
size_t vso_stored_size(int type, int metric, size_t dim) {
    size_t s = type_size(type) * dim;
    if (metric == VSO_COSINE && (type == VSO_INT8 || type == VSO_UINT8)) s += sizeof(float);
    return s;
}

/* ------------------------------------------------------------------------------------------ */
/* floating-point kernels                                                                     */
/*
 * The AVX512 kernels all have the shape "L lanes x A accumulators, residual first":
 *   r = dim % (L*A);  head = r % L;  nfull = r / L
 *   acc0[l]   = x[l]*y[l]           for l < head      (a multiply, not an FMA)
 *   acc(a)[l] = fma(x[..], y[..])   for a = 1..nfull  (one L-block each)
 *   then whole chunks of L*A elements, block a of each chunk into acc(a) with FMA
 *   s = combine(acc0..); result = lane tree-reduction(s)
 * spaces/IP/IP_AVX512F_FP32.h:19-56, L2/L2_AVX512F_FP32.h:21-59 (L=16,A=2)
 * spaces/IP/IP_AVX512F_FP64.h:19-56, L2/L2_AVX512F_FP64.h:21-60 (L=8, A=2)
 * spaces/IP/IP_AVX512F_FP16.h:27-68, L2/L2_AVX512F_FP16.h:28-70 (L=16,A=2 after cvtph_ps)
 * spaces/IP/IP_F16C_FP16.h:28-80,   L2/L2_F16C_FP16.h:28-82    (L=8, A=4, sequential lane sum)
 * Lane reduction of the 512-bit kernels = GCC's _mm512_reduce_add_ps/pd: halves, quarters,
 * then pairs (avx512fintrin.h __MM512_REDUCE_OP).
 */
#define MAXL 16
#define MAXA 4

static float lanes_f32(const float *x, const float *y, size_t dim, int L, int A, int is_l2,
                       int seq_reduce) {
    float acc[MAXA][MAXL];
    memset(acc, 0, sizeof(acc));
    const size_t chunk = (size_t)L * A;
    const size_t r = dim % chunk, head = r % L, nfull = r / L;
    size_t p = 0;
    for (size_t l = 0; l < head; l++) {
        if (is_l2) {
            float d = x[l] - y[l];
            acc[0][l] = d * d;
        } else {
            acc[0][l] = x[l] * y[l];
        }
    }
    p += head;
    for (size_t a = 1; a <= nfull; a++) {
        for (int l = 0; l < L; l++) {
            if (is_l2) {
                float d = x[p + l] - y[p + l];
                acc[a][l] = fmaf(d, d, acc[a][l]);
            } else {
                acc[a][l] = fmaf(x[p + l], y[p + l], acc[a][l]);
            }
        }
        p += L;
    }
    while (p < dim) {
        for (int a = 0; a < A; a++) {
            for (int l = 0; l < L; l++) {
                if (is_l2) {
                    float d = x[p + l] - y[p + l];
                    acc[a][l] = fmaf(d, d, acc[a][l]);
                } else {
                    acc[a][l] = fmaf(x[p + l], y[p + l], acc[a][l]);
                }
            }
            p += L;
        }
    }
    float s[MAXL];
    for (int l = 0; l < L; l++) {
        if (A == 1) s[l] = acc[0][l];
        else if (A == 2) s[l] = acc[0][l] + acc[1][l];
        else s[l] = (acc[0][l] + acc[1][l]) + (acc[2][l] + acc[3][l]);
    }
    if (seq_reduce) { /* AVX_utils.h:33-38 */
        float t = s[0];
        for (int l = 1; l < L; l++) t = t + s[l];
        return t;
    }
    for (int w = L / 2; w >= 1; w /= 2)
        for (int l = 0; l < w; l++) s[l] = s[l + w] + s[l];
    return s[0];
}

static double lanes_f64(const double *x, const double *y, size_t dim, int is_l2) {
    enum { L = 8, A = 2 };
    double acc[A][L];
    memset(acc, 0, sizeof(acc));
    const size_t r = dim % (L * A), head = r % L, nfull = r / L;
    size_t p = 0;
    for (size_t l = 0; l < head; l++) {
        if (is_l2) {
            double d = x[l] - y[l];
            acc[0][l] = d * d;
        } else {
            acc[0][l] = x[l] * y[l];
        }
    }
    p += head;
    for (size_t a = 1; a <= nfull; a++) {
        for (int l = 0; l < L; l++) {
            double u = x[p + l], v = y[p + l];
            if (is_l2) {
                double d = u - v;
                acc[a][l] = fma(d, d, acc[a][l]);
            } else {
                acc[a][l] = fma(u, v, acc[a][l]);
            }
        }
        p += L;
    }
    while (p < dim) {
        for (int a = 0; a < A; a++) {
            for (int l = 0; l < L; l++) {
                double u = x[p + l], v = y[p + l];
                if (is_l2) {
                    double d = u - v;
                    acc[a][l] = fma(d, d, acc[a][l]);
                } else {
                    acc[a][l] = fma(u, v, acc[a][l]);
                }
            }
            p += L;
        }
    }
    double s[L];
    for (int l = 0; l < L; l++) s[l] = acc[0][l] + acc[1][l];
    for (int w = L / 2; w >= 1; w /= 2)
        for (int l = 0; l < w; l++) s[l] = s[l + w] + s[l];
    return s[0];
}

static float tree16(float *s) {
    for (int w = 8; w >= 1; w /= 2)
        for (int l = 0; l < w; l++) s[l] = s[l + w] + s[l];
    return s[0];
}

/* fma with x86 DAZ+FTZ semantics, as VDPBF16PS applies them irrespective of MXCSR. */
static float flush(float v) { return fpclassify(v) == FP_SUBNORMAL ? copysignf(0.0f, v) : v; }
static float fma_ftz(float a, float b, float c) { return flush(fmaf(flush(a), flush(b), flush(c))); }

/* spaces/IP/IP_AVX512_BF16_VL_BF16.h:23-47 — one 16-lane accumulator fed by vdpbf16ps: lane p
 * takes the bf16 pair (2p, 2p+1) of each 32-element block, the odd element first, each step a
 * separately rounded FMA (Intel SDM pseudo-code; verified on an AVX512_BF16 host, 32M lanes).
 * The residual block (first dim%32 elements, zero padded) goes first. */
static float bf16_ip_dpbf16(const uint16_t *x, const uint16_t *y, size_t dim) {
    float s[16];
    memset(s, 0, sizeof(s));
    size_t r = dim % 32, p = 0;
    if (r) {
        for (int l = 0; l < 16; l++) {
            size_t e0 = 2 * (size_t)l, e1 = e0 + 1;
            float a1 = e1 < r ? vso_bf16_to_f32(x[e1]) : 0.0f, b1 = e1 < r ? vso_bf16_to_f32(y[e1]) : 0.0f;
            float a0 = e0 < r ? vso_bf16_to_f32(x[e0]) : 0.0f, b0 = e0 < r ? vso_bf16_to_f32(y[e0]) : 0.0f;
            s[l] = fma_ftz(a1, b1, s[l]);
            s[l] = fma_ftz(a0, b0, s[l]);
        }
        p += r;
    }
    do {
        for (int l = 0; l < 16; l++) {
            s[l] = fma_ftz(vso_bf16_to_f32(x[p + 2 * l + 1]), vso_bf16_to_f32(y[p + 2 * l + 1]), s[l]);
            s[l] = fma_ftz(vso_bf16_to_f32(x[p + 2 * l]), vso_bf16_to_f32(y[p + 2 * l]), s[l]);
        }
        p += 32;
    } while (p < dim);
    return tree16(s);
}

/* spaces/IP/IP_AVX512BW_VBMI2_BF16.h:39-76, L2/L2_AVX512BW_VBMI2_BF16.h:42-78 — one 16-lane
 * accumulator. Residual: lane j <- element j of the first 16 (if r>=16), then lane j <- element j
 * of the next r%16 (zero padded, still an FMA). Main: per 32 elements, lane 4q+i takes element
 * 8q+i (unpacklo) and then element 8q+4+i (unpackhi). */
static float bf16_vbmi2(const uint16_t *x, const uint16_t *y, size_t dim, int is_l2) {
    float s[16];
    memset(s, 0, sizeof(s));
    size_t r = dim % 32, p = 0;
#define BF_STEP(lane, ea, eb, valid)                                        \
    do {                                                                    \
        float a_ = (valid) ? vso_bf16_to_f32(x[ea]) : 0.0f;                 \
        float b_ = (valid) ? vso_bf16_to_f32(y[eb]) : 0.0f;                 \
        if (is_l2) {                                                        \
            float d_ = a_ - b_;                                             \
            s[lane] = fmaf(d_, d_, s[lane]);                                \
        } else {                                                            \
            s[lane] = fmaf(a_, b_, s[lane]);                                \
        }                                                                   \
    } while (0)
    if (r) {
        if (r >= 16) {
            for (int l = 0; l < 16; l++) BF_STEP(l, p + l, p + l, 1);
            p += 16;
        }
        if (r != 16) {
            size_t h = r % 16;
            for (int l = 0; l < 16; l++) BF_STEP(l, p + l, p + l, (size_t)l < h);
            p += h;
        }
    }
    do {
        for (int q = 0; q < 4; q++)
            for (int i = 0; i < 4; i++) BF_STEP(4 * q + i, p + 8 * q + i, p + 8 * q + i, 1);
        for (int q = 0; q < 4; q++)
            for (int i = 0; i < 4; i++) BF_STEP(4 * q + i, p + 8 * q + 4 + i, p + 8 * q + 4 + i, 1);
        p += 32;
    } while (p < dim);
#undef BF_STEP
    return tree16(s);
}

/* Scalar kernels: spaces/IP/IP.cpp:185-238, spaces/L2/L2.cpp:76-133. The reference compiles
 * these translation units without -mfma, so multiply and add round separately. */
static float naive_f32(const float *x, const float *y, size_t dim, int is_l2) {
    float res = 0;
    for (size_t i = 0; i < dim; i++) {
        if (is_l2) {
            float t = x[i] - y[i];
            res += t * t;
        } else {
            res += x[i] * y[i];
        }
    }
    return res;
}
static double naive_f64(const double *x, const double *y, size_t dim, int is_l2) {
    double res = 0;
    for (size_t i = 0; i < dim; i++) {
        if (is_l2) {
            double t = x[i] - y[i];
            res += t * t;
        } else {
            res += x[i] * y[i];
        }
    }
    return res;
}

/* Integer kernels are exact in any order: IP.cpp:247-286, L2.cpp:139-174 and the VNNI tiers
 * IP_AVX512F_BW_VL_VNNI_INT8.h:27-77, ..._UINT8.h:33-106 differ only in the width of the
 * accumulator, which never overflows inside the dimension cap of spaces.h:57-66. */
static long long int_dot(const void *a, const void *b, size_t dim, int is_unsigned) {
    long long res = 0;
    if (is_unsigned) {
        const uint8_t *x = a, *y = b;
        for (size_t i = 0; i < dim; i++) res += (int)x[i] * (int)y[i];
    } else {
        const int8_t *x = a, *y = b;
        for (size_t i = 0; i < dim; i++) res += (int)x[i] * (int)y[i];
    }
    return res;
}
static long long int_l2(const void *a, const void *b, size_t dim, int is_unsigned) {
    long long res = 0;
    if (is_unsigned) {
        const uint8_t *x = a, *y = b;
        for (size_t i = 0; i < dim; i++) {
            int d = (int)x[i] - (int)y[i];
            res += d * d;
        }
    } else {
        const int8_t *x = a, *y = b;
        for (size_t i = 0; i < dim; i++) {
            int d = (int)x[i] - (int)y[i];
            res += d * d;
        }
    }
    return res;
}

static void widen16(int type, const void *src, float *dst, size_t dim) {
    const uint16_t *h = src;
    for (size_t i = 0; i < dim; i++)
        dst[i] = type == VSO_BFLOAT16 ? vso_bf16_to_f32(h[i]) : vso_fp16_to_f32(h[i]);
}

/* Dispatch thresholds restate spaces/IP_space.cpp:435-887 and spaces/L2_space.cpp:185-517 for an
 * x86 host with the AVX512 feature set. */
double vso_distance(int type, int metric, size_t dim, const void *a, const void *b) {
    const int is_l2 = metric == VSO_L2;
    const int naive = g_tier == VSO_TIER_NAIVE;
    switch (type) {
    case VSO_FLOAT32: {
        float s = (naive || dim < 8) ? naive_f32(a, b, dim, is_l2) : lanes_f32(a, b, dim, 16, 2, is_l2, 0);
        return is_l2 ? s : 1.0f - s;
    }
    case VSO_FLOAT64: {
        double s = (naive || dim < 4) ? naive_f64(a, b, dim, is_l2) : lanes_f64(a, b, dim, is_l2);
        return is_l2 ? s : 1.0 - s;
    }
    case VSO_BFLOAT16: {
        float s;
        if (naive || dim < 32) {
            float *x = malloc(sizeof(float) * (2 * dim + 1)), *y = x + dim;
            widen16(type, a, x, dim);
            widen16(type, b, y, dim);
            s = naive_f32(x, y, dim, is_l2);
            free(x);
        } else if (is_l2 || g_tier == VSO_TIER_AVX512_NOBF16) {
            s = bf16_vbmi2(a, b, dim, is_l2);
        } else {
            s = bf16_ip_dpbf16(a, b, dim);
        }
        return is_l2 ? s : 1.0f - s;
    }
    case VSO_FLOAT16: {
        float *x = malloc(sizeof(float) * (2 * dim + 1)), *y = x + dim;
        widen16(type, a, x, dim);
        widen16(type, b, y, dim);
        float s;
        if (naive || dim < 8) s = naive_f32(x, y, dim, is_l2);
        else if (dim < 16) s = lanes_f32(x, y, dim, 8, 4, is_l2, 1);
        else s = lanes_f32(x, y, dim, 16, 2, is_l2, 0);
        free(x);
        return is_l2 ? s : 1.0f - s;
    }
    case VSO_INT8:
    case VSO_UINT8: {
        const int u = type == VSO_UINT8;
        if (is_l2) return (float)int_l2(a, b, dim, u);
        long long ip = int_dot(a, b, dim, u);
        if (metric == VSO_IP) return (float)(1 - ip);
        float na, nb;
        memcpy(&na, (const char *)a + dim, 4);
        memcpy(&nb, (const char *)b + dim, 4);
        return 1.0f - (float)ip / (na * nb);
    }
    }
    return NAN;
}

/* ------------------------------------------------------------------------------------------ */
/* normalisation: spaces/normalize/normalize_naive.h:23-88, compute_norm.h:17-31              */
void vso_normalize(int type, size_t dim, void *blob) {
    switch (type) {
    case VSO_FLOAT32: {
        float *v = blob;
        double sum = 0;
        for (size_t i = 0; i < dim; i++) sum += (double)v[i] * (double)v[i];
        float norm = (float)sqrt(sum);
        for (size_t i = 0; i < dim; i++) v[i] = v[i] / norm;
        return;
    }
    case VSO_FLOAT64: {
        double *v = blob;
        double sum = 0;
        for (size_t i = 0; i < dim; i++) sum += v[i] * v[i];
        double norm = sqrt(sum);
        for (size_t i = 0; i < dim; i++) v[i] = v[i] / norm;
        return;
    }
    case VSO_BFLOAT16:
    case VSO_FLOAT16: {
        uint16_t *v = blob;
        float *tmp = malloc(sizeof(float) * (dim + 1));
        float sum = 0;
        for (size_t i = 0; i < dim; i++) {
            float val = type == VSO_BFLOAT16 ? vso_bf16_to_f32(v[i]) : vso_fp16_to_f32(v[i]);
            tmp[i] = val;
            sum += val * val;
        }
        float norm = sqrtf(sum);
        for (size_t i = 0; i < dim; i++)
            v[i] = type == VSO_BFLOAT16 ? vso_f32_to_bf16(tmp[i] / norm) : vso_f32_to_fp16(tmp[i] / norm);
        free(tmp);
        return;
    }
    case VSO_INT8:
    case VSO_UINT8: {
        uint64_t sum = 0;
        if (type == VSO_INT8) {
            const int8_t *v = blob;
            for (size_t i = 0; i < dim; i++) sum += (uint64_t)((int)v[i] * (int)v[i]);
        } else {
            const uint8_t *v = blob;
            for (size_t i = 0; i < dim; i++) sum += (uint64_t)((int)v[i] * (int)v[i]);
        }
        float norm = (float)sqrt((double)sum);
        memcpy((char *)blob + dim, &norm, 4);
        return;
    }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* flat index: algorithms/brute_force/brute_force.h, brute_force_single.h, brute_force_multi.h */
struct vso_flat {
    int type, metric, multi;
    size_t dim, stored, block_size;
    size_t count, cap;
    char *rows;     /* count x stored, dense by internal id (blocks are a host-allocation detail) */
    size_t *labels; /* idToLabelMapping */
};

vso_flat *vso_flat_new(int type, size_t dim, int metric, int multi, size_t block_size) {
    vso_flat *f = calloc(1, sizeof(*f));
    f->type = type;
    f->metric = metric;
    f->multi = multi;
    f->dim = dim;
    f->stored = vso_stored_size(type, metric, dim);
    f->block_size = block_size ? block_size : 1024;
    return f;
}
void vso_flat_free(vso_flat *f) {
    if (!f) return;
    free(f->rows);
    free(f->labels);
    free(f);
}
size_t vso_flat_size(const vso_flat *f) { return f->count; }
size_t vso_flat_label_count(const vso_flat *f) {
    if (!f->multi) return f->count;
    size_t n = 0; /* distinct labels */
    for (size_t i = 0; i < f->count; i++) {
        size_t j = 0;
        for (; j < i; j++)
            if (f->labels[j] == f->labels[i]) break;
        n += j == i;
    }
    return n;
}
const void *vso_flat_row(const vso_flat *f, size_t id) { return f->rows + id * f->stored; }
size_t vso_flat_label_of(const vso_flat *f, size_t id) { return f->labels[id]; }

/* preprocessForStorage / preprocessQuery: spaces/computer/preprocessors.h:49-141 (cosine only). */
static void *preprocess(const vso_flat *f, const void *blob) {
    char *out = malloc(f->stored + 8);
    memcpy(out, blob, type_size(f->type) * f->dim);
    if (f->metric == VSO_COSINE) vso_normalize(f->type, f->dim, out);
    return out;
}

static long find_label(const vso_flat *f, size_t label) {
    for (size_t i = 0; i < f->count; i++)
        if (f->labels[i] == label) return (long)i;
    return -1;
}

/* brute_force_single.h:134-148 (update copies the caller's raw bytes, no preprocessing),
 * brute_force_multi.h addVector (always appends), brute_force.h:174-193 appendVector. */
int vso_flat_add(vso_flat *f, const void *blob, size_t label) {
    if (!f->multi) {
        long id = find_label(f, label);
        if (id >= 0) {
            /* updateElement(id, vector_data): for int8/uint8 cosine the reference reads 4 bytes
             * past the caller's blob; the restatement copies only what the caller owns. */
            memcpy(f->rows + (size_t)id * f->stored, blob, type_size(f->type) * f->dim);
            return 0;
        }
    }
    if (f->count == f->cap) {
        f->cap = f->cap ? f->cap * 2 : 64;
        f->rows = realloc(f->rows, f->cap * f->stored);
        f->labels = realloc(f->labels, f->cap * sizeof(size_t));
    }
    void *p = preprocess(f, blob);
    memcpy(f->rows + f->count * f->stored, p, f->stored);
    free(p);
    f->labels[f->count++] = label;
    return 1;
}

/* brute_force.h:195-224 removeVector: the last row moves into the hole. Multi: every vector of
 * the label goes (brute_force_multi.h deleteVector, highest ids first is not needed because ids
 * are re-resolved after each swap). */
int vso_flat_delete(vso_flat *f, size_t label) {
    int deleted = 0;
    for (;;) {
        long id = find_label(f, label);
        if (id < 0) break;
        size_t last = f->count - 1;
        if ((size_t)id != last) {
            memcpy(f->rows + (size_t)id * f->stored, f->rows + last * f->stored, f->stored);
            f->labels[id] = f->labels[last];
        }
        f->count--;
        deleted++;
        if (!f->multi) break;
    }
    return deleted;
}

/* --- max-heap of (score, label) pairs under std::pair's lexicographic order ---------------- */
/* utils/vecsim_stl.h:65-84 (std::priority_queue<pair<dist,label>, ..., std::less>). */
typedef struct {
    double score;
    size_t label;
} pairsl;
static int pair_less(pairsl a, pairsl b) {
    if (a.score < b.score) return 1;
    if (b.score < a.score) return 0;
    return a.label < b.label;
}
typedef struct {
    pairsl *v;
    size_t n, cap;
} heap_t;
static void heap_push(heap_t *h, pairsl x) {
    if (h->n == h->cap) {
        h->cap = h->cap ? h->cap * 2 : 16;
        h->v = realloc(h->v, h->cap * sizeof(pairsl));
    }
    size_t i = h->n++;
    h->v[i] = x;
    while (i > 0) {
        size_t p = (i - 1) / 2;
        if (!pair_less(h->v[p], h->v[i])) break;
        pairsl t = h->v[p];
        h->v[p] = h->v[i];
        h->v[i] = t;
        i = p;
    }
}
static void heap_pop(heap_t *h) {
    h->v[0] = h->v[--h->n];
    size_t i = 0;
    for (;;) {
        size_t l = 2 * i + 1, r = l + 1, m = i;
        if (l < h->n && pair_less(h->v[m], h->v[l])) m = l;
        if (r < h->n && pair_less(h->v[m], h->v[r])) m = r;
        if (m == i) break;
        pairsl t = h->v[m];
        h->v[m] = h->v[i];
        h->v[i] = t;
        i = m;
    }
}
/* Multi-value heap: utils/updatable_heap.h:66-111 — one entry per label, emplace on an existing
 * label only lowers its score, pop removes the max score (ties: max label). */
static long heap_find_label(const heap_t *h, size_t label) {
    for (size_t i = 0; i < h->n; i++)
        if (h->v[i].label == label) return (long)i;
    return -1;
}
static void heap_rebuild(heap_t *h) {
    heap_t t = {0};
    for (size_t i = 0; i < h->n; i++) heap_push(&t, h->v[i]);
    free(h->v);
    *h = t;
}

static int cmp_label(const void *a, const void *b) {
    const pairsl *x = a, *y = b;
    return x->label < y->label ? -1 : x->label > y->label;
}
static int cmp_score_label(const void *a, const void *b) {
    const pairsl *x = a, *y = b;
    if (x->score < y->score) return -1;
    if (y->score < x->score) return 1;
    return cmp_label(a, b);
}

static double flat_score(const vso_flat *f, size_t id, const void *pq) {
    return vso_distance(f->type, f->metric, f->dim, f->rows + id * f->stored, pq);
}
/* DistType is float for everything but fp64 (brute_force_factory.cpp:49-80). */
static double dist_lowest(const vso_flat *f) { return f->type == VSO_FLOAT64 ? -1.7976931348623157e308 : -3.402823466e38; }

/* brute_force.h:242-291 + vec_sim.cpp:345-357 */
size_t vso_flat_topk(const vso_flat *f, const void *query, size_t k, int order, int timeout,
                     size_t *labels, double *scores, int *code) {
    if (code) *code = 0;
    if (k == 0) return 0;
    void *pq = preprocess(f, query);
    heap_t h = {0};
    double upper = dist_lowest(f);
    for (size_t id = 0; id < f->count; id++) {
        if (timeout) {
            if (code) *code = 1;
            free(pq);
            free(h.v);
            return 0;
        }
        double score = flat_score(f, id, pq);
        if (score < upper || h.n < k) {
            pairsl x = {score, f->labels[id]};
            if (f->multi) {
                long pos = heap_find_label(&h, x.label);
                if (pos >= 0) {
                    if (score < h.v[pos].score) {
                        h.v[pos].score = score;
                        heap_rebuild(&h);
                    }
                } else {
                    heap_push(&h, x);
                }
            } else {
                heap_push(&h, x);
            }
            if (h.n > k) heap_pop(&h);
            upper = h.v[0].score;
        }
    }
    size_t n = h.n;
    for (size_t i = n; i-- > 0;) {
        labels[i] = h.v[0].label;
        scores[i] = h.v[0].score;
        heap_pop(&h);
    }
    free(h.v);
    free(pq);
    if (order == VSO_BY_ID) {
        pairsl *t = malloc(sizeof(pairsl) * (n + 1));
        for (size_t i = 0; i < n; i++) t[i] = (pairsl){scores[i], labels[i]};
        qsort(t, n, sizeof(pairsl), cmp_label);
        for (size_t i = 0; i < n; i++) labels[i] = t[i].label, scores[i] = t[i].score;
        free(t);
    }
    return n;
}

/* brute_force.h:293-326, vec_sim_index.h:246-252, vec_sim.cpp:359-369. Results with equal score
 * come back ordered by label (std::sort is unstable there; any order of ties is reference-valid). */
long vso_flat_range(const vso_flat *f, const void *query, double radius, int order, int timeout,
                    size_t cap, size_t *labels, double *scores, int *code) {
    if (code) *code = 0;
    if (order != VSO_BY_ID && order != VSO_BY_SCORE) return -1;
    if (radius < 0) return -1;
    void *pq = preprocess(f, query);
    double r = f->type == VSO_FLOAT64 ? radius : (double)(float)radius;
    pairsl *res = malloc(sizeof(pairsl) * (f->count + 1));
    size_t n = 0;
    for (size_t id = 0; id < f->count; id++) {
        if (timeout) {
            if (code) *code = 1;
            break;
        }
        double score = flat_score(f, id, pq);
        if (score <= r) {
            if (f->multi) { /* unique_results_container: min score per label */
                size_t j = 0;
                for (; j < n; j++)
                    if (res[j].label == f->labels[id]) break;
                if (j < n) {
                    if (score < res[j].score) res[j].score = score;
                    continue;
                }
            }
            res[n++] = (pairsl){score, f->labels[id]};
        }
    }
    qsort(res, n, sizeof(pairsl), order == VSO_BY_ID ? cmp_label : cmp_score_label);
    for (size_t i = 0; i < n && i < cap; i++) labels[i] = res[i].label, scores[i] = res[i].score;
    free(res);
    free(pq);
    return (long)n;
}

/* brute_force_single.h:200-212 — the caller's blob is used as is (no preprocessing). */
double vso_flat_distance_from(const vso_flat *f, size_t label, const void *query) {
    double best = NAN;
    for (size_t i = 0; i < f->count; i++) {
        if (f->labels[i] != label) continue;
        double s = vso_distance(f->type, f->metric, f->dim, f->rows + i * f->stored, query);
        if (isnan(best) || s < best) best = s;
        if (!f->multi) break;
    }
    return best;
}

/* ------------------------------------------------------------------------------------------ */
/* batch iterator: bf_batch_iterator.h:59-214, bfs_batch_iterator.h:24-41,                     */
/* bfm_batch_iterator.h:24-53. Each Next returns the n best (score, then label) not returned   */
/* yet; within a group of equal scores the reference's own order is unspecified                */
/* (std::nth_element / unstable sort), so (score,label) is one of its valid outputs.           */
struct vso_bi {
    const vso_flat *f;
    void *pq;
    pairsl *scores;
    size_t n, pos, returned, label_count;
    int computed;
};
vso_bi *vso_bi_new(const vso_flat *f, const void *query) {
    vso_bi *it = calloc(1, sizeof(*it));
    it->f = f;
    it->pq = preprocess(f, query);
    it->label_count = vso_flat_label_count(f);
    return it;
}
static void bi_compute(vso_bi *it) {
    const vso_flat *f = it->f;
    it->scores = malloc(sizeof(pairsl) * (f->count + 1));
    it->n = 0;
    for (size_t id = 0; id < f->count; id++) {
        double s = flat_score(f, id, it->pq);
        if (f->multi) {
            size_t j = 0;
            for (; j < it->n; j++)
                if (it->scores[j].label == f->labels[id]) break;
            if (j < it->n) {
                if (s < it->scores[j].score) it->scores[j].score = s;
                continue;
            }
        }
        it->scores[it->n++] = (pairsl){s, f->labels[id]};
    }
    qsort(it->scores, it->n, sizeof(pairsl), cmp_score_label);
    it->label_count = it->n;
    it->computed = 1;
}
size_t vso_bi_next(vso_bi *it, size_t n, int order, size_t *labels, double *scores, int *code) {
    if (code) *code = 0;
    if (!it->computed) bi_compute(it);
    size_t m = it->n - it->pos;
    if (n < m) m = n;
    pairsl *out = malloc(sizeof(pairsl) * (m + 1));
    memcpy(out, it->scores + it->pos, sizeof(pairsl) * m);
    it->pos += m;
    it->returned += m;
    if (order == VSO_BY_ID) qsort(out, m, sizeof(pairsl), cmp_label);
    for (size_t i = 0; i < m; i++) labels[i] = out[i].label, scores[i] = out[i].score;
    free(out);
    return m;
}
int vso_bi_has_next(const vso_bi *it) { return it->returned != it->label_count; }
void vso_bi_reset(vso_bi *it) {
    free(it->scores);
    it->scores = NULL;
    it->n = it->pos = it->returned = 0;
    it->computed = 0;
}
void vso_bi_free(vso_bi *it) {
    if (!it) return;
    free(it->scores);
    free(it->pq);
    free(it);
}
