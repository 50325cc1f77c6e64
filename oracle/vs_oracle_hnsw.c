/* oracle/vs_oracle_hnsw.c — TEST INFRASTRUCTURE ONLY (see vs_oracle.h).
 *
 * Plain-C restatement of the reference's single-threaded HNSW (paths relative to
 * /root/reference/src/VecSim/algorithms/hnsw): build, top-k, range. Distances come from
 * vso_distance (vs_oracle.c), i.e. the same bit-exact kernels the flat oracle uses.
 *
 *   level draw              hnsw.h:418-422 (std::default_random_engine = minstd_rand0,
 *                           uniform_real_distribution<double> = libstdc++ generate_canonical)
 *   insert                  hnsw.h:1567-1602 insertElementToGraph, :1860-1960 store/index
 *   searchLayer             hnsw.h:682-721, processCandidate :530-613
 *   neighbour selection     hnsw.h:725-799 getNeighborsByHeuristic2
 *   connect / revisit       hnsw.h:870-941, :801-868
 *   greedy descent          hnsw.h:1210-1258, :1967-1981
 *   top-k                   hnsw.h:1983-2084
 *   range                   hnsw.h:2086-2186, :615-680
 *
 * Heaps restate libstdc++'s push_heap / pop_heap (bits/stl_heap.h) because the order of the result
 * heap's underlying array is observable: mutuallyConnectNewElement iterates it (hnsw.h:878-880) and
 * links are appended in that order when fewer than M candidates exist.
 * Known freedom: std::sort's order among equal distances (hnsw.h:760) is unspecified; equal
 * distances are ordered by id here (the GPU builder does the same). Random real-valued data has no
 * such ties.
 *
 * Parity status: PINNED against oracle/_ref (tests/test_oracle_vs_reference.py::test_hnsw_*): same
 * levels, entry point, link lists in order, and identical query results, for all six types. */
#include "vs_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    double d;
    size_t key; /* id, or label for the query result heap */
} pr_t;

typedef struct {
    pr_t *a;
    size_t n, cap;
} heap_t;

/* std::pair<DistType, key> operator< */
static int pr_less(pr_t x, pr_t y) { return x.d < y.d || (!(y.d < x.d) && x.key < y.key); }

static void h_reserve(heap_t *h, size_t n) {
    if (n > h->cap) {
        h->cap = n * 2 + 16;
        h->a = (pr_t *)realloc(h->a, h->cap * sizeof(pr_t));
    }
}
static void sift_up(pr_t *a, size_t hole, size_t top, pr_t v) {
    while (hole > top) {
        size_t parent = (hole - 1) / 2;
        if (!pr_less(a[parent], v)) break;
        a[hole] = a[parent];
        hole = parent;
    }
    a[hole] = v;
}
static void h_push(heap_t *h, double d, size_t key) {
    h_reserve(h, h->n + 1);
    pr_t v = {d, key};
    h->n++;
    sift_up(h->a, h->n - 1, 0, v);
}
/* pop_heap + pop_back: the hole sinks to a leaf along the larger children, then the former last
 * element is pushed up from there */
static void h_pop(heap_t *h) {
    size_t len = h->n - 1;
    pr_t v = h->a[len];
    h->n = len;
    if (len == 0) return;
    size_t hole = 0, child = 0;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (pr_less(h->a[child], h->a[child - 1])) child--;
        h->a[hole] = h->a[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        h->a[hole] = h->a[child - 1];
        hole = child - 1;
    }
    sift_up(h->a, hole, 0, v);
}

struct vso_hnsw {
    int type, metric;
    size_t dim, stored, M, M0, efc, ef;
    double epsilon, mult;
    uint32_t rng; /* minstd_rand0 state */
    size_t n, cap;
    unsigned char *rows;
    size_t *labels;
    uint8_t *deleted;
    uint32_t *levels;
    uint32_t **links; /* links[id]: level 0 record (1 + M0) followed by `level` records of (1 + M) */
    long entry, max_level;
    uint32_t *tags;
    uint32_t tag;
    size_t n_dist;
    int multi; /* HNSWIndex_Multi: several vectors per label; queries return each label once with its best score */
};

static uint32_t *rec(const vso_hnsw *g, size_t id, size_t level) {
    return level == 0 ? g->links[id] : g->links[id] + (1 + g->M0) + (level - 1) * (1 + g->M);
}
static const void *row(const vso_hnsw *g, size_t id) { return g->rows + id * g->stored; }
static double dist(vso_hnsw *g, size_t id, const void *q) {
    g->n_dist++;
    return vso_distance(g->type, g->metric, g->dim, row(g, id), q);
}
static double dist_max(const vso_hnsw *g) { return g->type == VSO_FLOAT64 ? 1.7976931348623157e308 : 3.402823466e38; }

vso_hnsw *vso_hnsw_new(int type, size_t dim, int metric, size_t M, size_t ef_construction, size_t ef_runtime,
                       double epsilon) {
    vso_hnsw *g = (vso_hnsw *)calloc(1, sizeof(*g));
    g->type = type;
    g->metric = metric;
    g->dim = dim;
    g->stored = vso_stored_size(type, metric, dim);
    g->M = M ? M : 16;
    g->M0 = 2 * g->M;
    g->efc = ef_construction ? ef_construction : 200;
    if (g->efc < g->M) g->efc = g->M;
    g->ef = ef_runtime ? ef_runtime : 10;
    g->epsilon = epsilon > 0 ? epsilon : 0.01;
    g->mult = 1 / log(1.0 * (double)g->M);
    g->rng = 100; /* hnsw.h:230 */
    g->entry = g->max_level = -1;
    return g;
}

void vso_hnsw_free(vso_hnsw *g) {
    if (!g) return;
    for (size_t i = 0; i < g->n; i++) free(g->links[i]);
    free(g->links);
    free(g->rows);
    free(g->labels);
    free(g->deleted);
    free(g->levels);
    free(g->tags);
    free(g);
}

/* std::minstd_rand0: x <- 16807 x mod (2^31 - 1); min 1, max 2^31 - 2 */
static uint32_t minstd(vso_hnsw *g) {
    g->rng = (uint32_t)(((uint64_t)g->rng * 16807u) % 2147483647u);
    return g->rng;
}
/* libstdc++ generate_canonical<double, 53>: k = 2 draws, range R = 2^31 - 2 */
static double canonical(vso_hnsw *g) {
    const double R = 2147483646.0;
    double sum = (double)(minstd(g) - 1u);
    sum += (double)(minstd(g) - 1u) * R;
    double ret = sum / (R * R);
    if (ret >= 1.0) ret = nextafter(1.0, 0.0);
    return ret;
}
static size_t draw_level(vso_hnsw *g) {
    double r = -log(canonical(g)) * g->mult;
    return (size_t)r;
}

static uint32_t fresh_tag(vso_hnsw *g) { return ++g->tag; }

/* greedySearchLevel (hnsw.h:1210-1258) */
static void greedy(vso_hnsw *g, const void *q, size_t level, size_t *best, double *cur, int running_query) {
    size_t best_alive = *best;
    int changed;
    do {
        changed = 0;
        const uint32_t *r = rec(g, *best, level);
        for (uint32_t i = 0; i < r[0]; i++) {
            size_t c = r[1 + i];
            double d = dist(g, c, q);
            if (d < *cur) {
                *cur = d;
                *best = c;
                changed = 1;
                if (!running_query && !g->deleted[c]) best_alive = c;
            }
        }
    } while (changed);
    if (!running_query) *best = best_alive;
}

/* searchLayer / searchBottomLayer_WithTimeout. by_label: result heap keyed (dist, label). */
static void search_layer(vso_hnsw *g, size_t ep, const void *q, size_t level, size_t ef, int by_label, heap_t *top) {
    heap_t cand = {0};
    uint32_t tag = fresh_tag(g);
    double lower;
    top->n = 0;
    if (!g->deleted[ep]) {
        double d = dist(g, ep, q);
        lower = d;
        h_push(top, d, by_label ? g->labels[ep] : ep);
        h_push(&cand, -d, ep);
    } else {
        lower = dist_max(g);
        h_push(&cand, -lower, ep);
    }
    g->tags[ep] = tag;
    while (cand.n) {
        pr_t c = cand.a[0];
        if (-c.d > lower && top->n >= ef) break;
        h_pop(&cand);
        const uint32_t *r = rec(g, c.key, level);
        for (uint32_t j = 0; j < r[0]; j++) {
            size_t id = r[1 + j];
            if (g->tags[id] == tag) continue;
            g->tags[id] = tag;
            double d = dist(g, id, q);
            if (lower > d || top->n < ef) {
                h_push(&cand, -d, id);
                if (!g->deleted[id]) h_push(top, d, by_label ? g->labels[id] : id);
                if (top->n > ef) h_pop(top);
                if (top->n) lower = top->a[0].d;
            }
        }
    }
    free(cand.a);
}

/* The result set of a multi-value search: vecsim_stl::updatable_max_heap<DistType, labelType>
 * (utils/updatable_heap.h:24-130; HNSWIndex_Multi::getNewMaxPriorityQueue / emplaceToHeap, hnsw_multi.h:62-71, :103-106).
 * One entry per label; emplace inserts a new label or lowers the score of a known one, never raises it; top / pop act on
 * the maximum under (score, label) — top_ptr() picks the largest label among the largest scores. Kept as a flat array:
 * this is a checker, not a data structure. */
static long ms_find(const heap_t *h, size_t label) {
    for (size_t i = 0; i < h->n; i++)
        if (h->a[i].key == label) return (long)i;
    return -1;
}
static void ms_emplace(heap_t *h, double d, size_t label) {
    long i = ms_find(h, label);
    if (i < 0) {
        h_reserve(h, h->n + 1);
        h->a[h->n].d = d;
        h->a[h->n].key = label;
        h->n++;
    } else if (h->a[i].d > d) {
        h->a[i].d = d;
    }
}
static size_t ms_top(const heap_t *h) {
    size_t m = 0;
    for (size_t i = 1; i < h->n; i++)
        if (pr_less(h->a[m], h->a[i])) m = i;
    return m;
}
static void ms_pop(heap_t *h) {
    size_t m = ms_top(h);
    h->a[m] = h->a[h->n - 1];
    h->n--;
}

/* searchBottomLayer_WithTimeout on a multi-value index (hnsw.h:530-613, :1983-2035 with the heap above): the candidate
 * set is keyed by internal id exactly as in the single-value search, the result set by label. */
static void search_layer_multi(vso_hnsw *g, size_t ep, const void *q, size_t ef, heap_t *top) {
    heap_t cand = {0};
    uint32_t tag = fresh_tag(g);
    double lower;
    top->n = 0;
    if (!g->deleted[ep]) {
        double d = dist(g, ep, q);
        lower = d;
        ms_emplace(top, d, g->labels[ep]);
        h_push(&cand, -d, ep);
    } else {
        lower = dist_max(g);
        h_push(&cand, -lower, ep);
    }
    g->tags[ep] = tag;
    while (cand.n) {
        pr_t c = cand.a[0];
        if (-c.d > lower && top->n >= ef) break;
        h_pop(&cand);
        const uint32_t *r = rec(g, c.key, 0);
        for (uint32_t j = 0; j < r[0]; j++) {
            size_t id = r[1 + j];
            if (g->tags[id] == tag) continue;
            g->tags[id] = tag;
            double d = dist(g, id, q);
            if (lower > d || top->n < ef) {
                h_push(&cand, -d, id);
                if (!g->deleted[id]) ms_emplace(top, d, g->labels[id]);
                if (top->n > ef) ms_pop(top);
                if (top->n) lower = top->a[ms_top(top)].d;
            }
        }
    }
    free(cand.a);
}

static int cmp_dist_id(const void *a, const void *b) {
    const pr_t *x = (const pr_t *)a, *y = (const pr_t *)b;
    if (x->d < y->d) return -1;
    if (y->d < x->d) return 1;
    return x->key < y->key ? -1 : (x->key > y->key ? 1 : 0);
}
static int cmp_key_dist(const void *a, const void *b) {
    const pr_t *x = (const pr_t *)a, *y = (const pr_t *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->d < y->d ? -1 : (y->d < x->d ? 1 : 0);
}

/* getNeighborsByHeuristic2_internal (hnsw.h:745-799): list[n] -> kept in place, returns kept
 * count; removed ids go to `removed` (may be NULL). */
static size_t heuristic(vso_hnsw *g, pr_t *list, size_t n, size_t M, size_t *removed, size_t *n_removed) {
    if (n_removed) *n_removed = 0;
    if (n < M) return n;
    qsort(list, n, sizeof(pr_t), cmp_dist_id);
    pr_t *keep = (pr_t *)malloc((M + 1) * sizeof(pr_t));
    size_t nk = 0, i = 0;
    for (; i < n && nk < M; i++) {
        int good = 1;
        for (size_t j = 0; j < nk; j++) {
            double d = vso_distance(g->type, g->metric, g->dim, row(g, keep[j].key), row(g, list[i].key));
            g->n_dist++;
            if (d < list[i].d) {
                good = 0;
                break;
            }
        }
        if (good) keep[nk++] = list[i];
        else if (removed) removed[(*n_removed)++] = list[i].key;
    }
    for (; i < n; i++)
        if (removed) removed[(*n_removed)++] = list[i].key;
    memcpy(list, keep, nk * sizeof(pr_t));
    free(keep);
    return nk;
}

static int contains(const size_t *a, size_t n, size_t v) {
    for (size_t i = 0; i < n; i++)
        if (a[i] == v) return 1;
    return 0;
}

/* revisitNeighborConnections (hnsw.h:801-868), single-threaded */
static void revisit(vso_hnsw *g, size_t level, size_t new_id, pr_t nb, uint32_t *new_rec, uint32_t *nb_rec) {
    size_t maxM = level ? g->M : g->M0;
    size_t nc = nb_rec[0] + 1;
    pr_t *cand = (pr_t *)malloc(nc * sizeof(pr_t));
    size_t *removed = (size_t *)malloc(nc * sizeof(size_t));
    cand[0].d = nb.d;
    cand[0].key = new_id;
    for (uint32_t j = 0; j < nb_rec[0]; j++) {
        cand[1 + j].d = vso_distance(g->type, g->metric, g->dim, row(g, nb_rec[1 + j]), row(g, nb.key));
        cand[1 + j].key = nb_rec[1 + j];
        g->n_dist++;
    }
    size_t n_removed = 0;
    heuristic(g, cand, nc, maxM, removed, &n_removed);
    int new_chosen = !contains(removed, n_removed, new_id);
    uint32_t kept = 0;
    for (uint32_t i = 0; i < nb_rec[0]; i++)
        if (!contains(removed, n_removed, nb_rec[1 + i])) nb_rec[1 + kept++] = nb_rec[1 + i];
    if (new_rec[0] < maxM && !g->deleted[new_id] && !g->deleted[nb.key]) {
        new_rec[1 + new_rec[0]++] = (uint32_t)nb.key;
        if (new_chosen && kept < maxM) nb_rec[1 + kept++] = (uint32_t)new_id;
    }
    nb_rec[0] = kept;
    free(cand);
    free(removed);
}

/* mutuallyConnectNewElement (hnsw.h:870-941); returns the next closest entry point */
static size_t connect(vso_hnsw *g, size_t new_id, heap_t *top, size_t level) {
    size_t maxM = level ? g->M : g->M0;
    size_t n = top->n;
    pr_t *list = (pr_t *)malloc((n + 1) * sizeof(pr_t));
    memcpy(list, top->a, n * sizeof(pr_t)); /* the heap's underlying array, in order */
    size_t next_ep;
    if (n < g->M) {
        size_t best = 0;
        for (size_t i = 1; i < n; i++)
            if (list[i].d < list[best].d) best = i;
        next_ep = list[best].key;
    } else {
        n = heuristic(g, list, n, g->M, NULL, NULL);
        next_ep = list[0].key;
    }
    uint32_t *new_rec = rec(g, new_id, level);
    for (size_t i = 0; i < n; i++) {
        size_t nb = list[i].key;
        if (new_rec[0] == maxM) break;
        if (g->deleted[new_id] || g->deleted[nb]) continue;
        uint32_t *nb_rec = rec(g, nb, level);
        if (nb_rec[0] < maxM) {
            new_rec[1 + new_rec[0]++] = (uint32_t)nb;
            nb_rec[1 + nb_rec[0]++] = (uint32_t)new_id;
            continue;
        }
        revisit(g, level, new_id, list[i], new_rec, nb_rec);
    }
    free(list);
    return next_ep;
}

static void grow(vso_hnsw *g) {
    if (g->n < g->cap) return;
    size_t cap = g->cap ? g->cap * 2 : 1024;
    g->rows = (unsigned char *)realloc(g->rows, cap * g->stored);
    g->labels = (size_t *)realloc(g->labels, cap * sizeof(size_t));
    g->deleted = (uint8_t *)realloc(g->deleted, cap);
    g->levels = (uint32_t *)realloc(g->levels, cap * sizeof(uint32_t));
    g->links = (uint32_t **)realloc(g->links, cap * sizeof(uint32_t *));
    g->tags = (uint32_t *)realloc(g->tags, cap * sizeof(uint32_t));
    memset(g->tags + g->cap, 0, (cap - g->cap) * sizeof(uint32_t));
    g->cap = cap;
}

static void *processed(const vso_hnsw *g, const void *blob) {
    void *p = calloc(1, g->stored + 8);
    size_t raw = g->stored - ((g->metric == VSO_COSINE && (g->type == VSO_INT8 || g->type == VSO_UINT8)) ? 4 : 0);
    memcpy(p, blob, raw);
    if (g->metric == VSO_COSINE) vso_normalize(g->type, g->dim, p);
    return p;
}

/* appendVector (hnsw.h:1948-1960); labels are assumed new (the single-value overwrite path deletes first) */
void vso_hnsw_add(vso_hnsw *g, const void *blob, size_t label) {
    void *p = processed(g, blob);
    grow(g);
    size_t level = draw_level(g);
    size_t id = g->n++;
    memcpy(g->rows + id * g->stored, p, g->stored);
    g->labels[id] = label;
    g->deleted[id] = 0;
    g->levels[id] = (uint32_t)level;
    g->links[id] = (uint32_t *)calloc((1 + g->M0) + level * (1 + g->M), sizeof(uint32_t));
    long prev_ep = g->entry, prev_max = g->max_level;
    if ((long)level > prev_max) {
        g->entry = (long)id;
        g->max_level = (long)level;
    }
    if (prev_ep >= 0) {
        size_t cur = (size_t)prev_ep;
        long max_common;
        if ((long)level < prev_max) {
            max_common = (long)level;
            double cd = dist(g, cur, p);
            for (long l = prev_max; l > (long)level; l--) greedy(g, p, (size_t)l, &cur, &cd, 0);
        } else {
            max_common = prev_max;
        }
        heap_t top = {0};
        for (long l = max_common; l >= 0; l--) {
            search_layer(g, cur, p, (size_t)l, g->efc, 0, &top);
            if (top.n) cur = connect(g, id, &top, (size_t)l);
        }
        free(top.a);
    }
    free(p);
}

size_t vso_hnsw_size(const vso_hnsw *g) { return g->n; }
void vso_hnsw_set_multi(vso_hnsw *g, int multi) { g->multi = multi != 0; }
void vso_hnsw_mark_deleted(vso_hnsw *g, size_t id, int deleted) { g->deleted[id] = deleted ? 1 : 0; }
void vso_hnsw_info(const vso_hnsw *g, long *entry, long *max_level) {
    *entry = g->entry;
    *max_level = g->max_level;
}
uint32_t vso_hnsw_level(const vso_hnsw *g, size_t id) { return g->levels[id]; }
/* links of `id` at `level` into out (capacity M0); returns the count */
size_t vso_hnsw_links(const vso_hnsw *g, size_t id, size_t level, uint32_t *out) {
    const uint32_t *r = rec(g, id, level);
    memcpy(out, r + 1, r[0] * sizeof(uint32_t));
    return r[0];
}
size_t vso_hnsw_dist_count(const vso_hnsw *g) { return g->n_dist; }

/* searchBottomLayerEP (hnsw.h:1967-1981) */
static long bottom_ep(vso_hnsw *g, const void *q, double *cd) {
    if (g->entry < 0) return -1;
    size_t cur = (size_t)g->entry;
    *cd = dist(g, cur, q);
    for (long l = g->max_level; l > 0; l--) greedy(g, q, (size_t)l, &cur, cd, 1);
    return (long)cur;
}

/* topKQuery (hnsw.h:2037-2084): ascending (score, label). ef_runtime 0 = index default. */
size_t vso_hnsw_topk(vso_hnsw *g, const void *query, size_t k, size_t ef_runtime, size_t *labels, double *scores) {
    if (g->n == 0 || k == 0) return 0;
    void *q = processed(g, query);
    size_t ef = ef_runtime ? ef_runtime : g->ef;
    if (ef < k) ef = k;
    double cd;
    long ep = bottom_ep(g, q, &cd);
    size_t n = 0;
    if (ep >= 0) {
        heap_t top = {0};
        if (g->multi) {
            search_layer_multi(g, (size_t)ep, q, ef, &top);
            while (top.n > k) ms_pop(&top);
            n = top.n;
            for (size_t i = n; i-- > 0;) {
                size_t m = ms_top(&top);
                scores[i] = top.a[m].d;
                labels[i] = top.a[m].key;
                ms_pop(&top);
            }
        } else {
            search_layer(g, (size_t)ep, q, 0, ef, 1, &top);
            while (top.n > k) h_pop(&top);
            n = top.n;
            for (size_t i = n; i-- > 0;) {
                scores[i] = top.a[0].d;
                labels[i] = top.a[0].key;
                h_pop(&top);
            }
        }
        free(top.a);
    }
    free(q);
    return n;
}

/* rangeQuery (hnsw.h:2086-2186, :615-680): results sorted by (score, label); returns the total, writes <= cap */
size_t vso_hnsw_range(vso_hnsw *g, const void *query, double radius_in, double epsilon, size_t cap, size_t *labels,
                      double *scores) {
    if (g->n == 0) return 0;
    void *q = processed(g, query);
    if (epsilon == 0.0) epsilon = g->epsilon;
    const int f64 = g->type == VSO_FLOAT64;
    const double radius = f64 ? radius_in : (double)(float)radius_in;
    double cd;
    long epl = bottom_ep(g, q, &cd);
    size_t found = 0;
    pr_t *res = NULL;
    size_t res_cap = 0;
    if (epl >= 0) {
        size_t ep = (size_t)epl;
        heap_t cand = {0};
        uint32_t tag = fresh_tag(g);
        double ep_dist, dyn, bound;
#define ROUND_DT(x) (f64 ? (x) : (double)(float)(x))
#define EMIT(id_, d_)                                                      \
    do {                                                                   \
        if (found == res_cap) {                                            \
            res_cap = res_cap * 2 + 16;                                    \
            res = (pr_t *)realloc(res, res_cap * sizeof(pr_t));            \
        }                                                                  \
        res[found].d = (d_);                                               \
        res[found].key = g->labels[(id_)];                                 \
        found++;                                                           \
    } while (0)
        if (g->deleted[ep]) {
            ep_dist = dist_max(g);
            bound = dyn = ep_dist;
        } else {
            ep_dist = dist(g, ep, q);
            dyn = ep_dist;
            if (ep_dist <= radius) {
                EMIT(ep, ep_dist);
                dyn = radius;
            }
            bound = ROUND_DT(dyn * (1.0 + epsilon));
        }
        h_push(&cand, -ep_dist, ep);
        g->tags[ep] = tag;
        while (cand.n) {
            pr_t c = cand.a[0];
            if (-c.d > bound) break;
            h_pop(&cand);
            if (-c.d < dyn && -c.d >= radius) {
                dyn = -c.d;
                bound = ROUND_DT(dyn * (1.0 + epsilon));
            }
            const uint32_t *r = rec(g, c.key, 0);
            for (uint32_t j = 0; j < r[0]; j++) {
                size_t id = r[1 + j];
                if (g->tags[id] == tag) continue;
                g->tags[id] = tag;
                double d = dist(g, id, q);
                if (d < bound) {
                    h_push(&cand, -d, id);
                    if (d <= radius && !g->deleted[id]) EMIT(id, d);
                }
            }
        }
        free(cand.a);
    }
    if (g->multi && found > 1) {
        /* unique_results_container (containers/vecsim_results_container.h:32-60): one result per label, its lowest score */
        qsort(res, found, sizeof(pr_t), cmp_key_dist);
        size_t o = 0;
        for (size_t i = 0; i < found; i++)
            if (i == 0 || res[i].key != res[o - 1].key) res[o++] = res[i];
        found = o;
    }
    qsort(res, found, sizeof(pr_t), cmp_dist_id);
    for (size_t i = 0; i < found && i < cap; i++) {
        labels[i] = res[i].key;
        scores[i] = res[i].d;
    }
    free(res);
    free(q);
    return found;
}

/* ---- batch iterator (hnsw_batch_iterator.h:59-267, hnsw_single_batch_iterator.h:36-78) ---- */
struct vso_hnsw_bi {
    vso_hnsw *g;
    void *q;
    size_t ef, returned;
    long entry; /* -1 = not computed */
    int depleted;
    double lower;
    heap_t cand, extras; /* min-heaps: stored negated through pr_greater */
    uint8_t *visited;
    size_t *ret_labels; /* multi-value: labels handed out so far (HNSWMulti_BatchIterator::returned) */
    size_t n_ret, cap_ret;
};
static int bi_returned(const vso_hnsw_bi *it, size_t label) {
    for (size_t i = 0; i < it->n_ret; i++)
        if (it->ret_labels[i] == label) return 1;
    return 0;
}
static void bi_mark_returned(vso_hnsw_bi *it, size_t label) {
    if (it->n_ret == it->cap_ret) {
        it->cap_ret = it->cap_ret * 2 + 64;
        it->ret_labels = (size_t *)realloc(it->ret_labels, it->cap_ret * sizeof(size_t));
    }
    it->ret_labels[it->n_ret++] = label;
}

/* std::greater<pair>: min-heap ordering for the same push/pop machinery */
static void mh_sift_up(pr_t *a, size_t hole, pr_t v) {
    while (hole > 0) {
        size_t parent = (hole - 1) / 2;
        if (!pr_less(v, a[parent])) break; /* comp(parent, v) = parent > v */
        a[hole] = a[parent];
        hole = parent;
    }
    a[hole] = v;
}
static void mh_push(heap_t *h, double d, size_t key) {
    h_reserve(h, h->n + 1);
    pr_t v = {d, key};
    h->n++;
    mh_sift_up(h->a, h->n - 1, v);
}
static void mh_pop(heap_t *h) {
    size_t len = h->n - 1;
    pr_t v = h->a[len];
    h->n = len;
    if (len == 0) return;
    size_t hole = 0, child = 0;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (pr_less(h->a[child - 1], h->a[child])) child--; /* comp(child, child-1) = child > child-1 */
        h->a[hole] = h->a[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        h->a[hole] = h->a[child - 1];
        hole = child - 1;
    }
    mh_sift_up(h->a, hole, v);
}

vso_hnsw_bi *vso_hnsw_bi_new(vso_hnsw *g, const void *query, size_t ef_runtime) {
    vso_hnsw_bi *it = (vso_hnsw_bi *)calloc(1, sizeof(*it));
    it->g = g;
    it->q = processed(g, query);
    it->ef = ef_runtime ? ef_runtime : g->ef;
    it->entry = -1;
    it->lower = INFINITY;
    it->visited = (uint8_t *)calloc(g->n + 1, 1);
    return it;
}
void vso_hnsw_bi_free(vso_hnsw_bi *it) {
    if (!it) return;
    free(it->q);
    free(it->cand.a);
    free(it->extras.a);
    free(it->visited);
    free(it->ret_labels);
    free(it);
}
void vso_hnsw_bi_reset(vso_hnsw_bi *it) {
    it->returned = 0;
    it->depleted = 0;
    it->lower = INFINITY;
    it->cand.n = it->extras.n = 0;
    it->n_ret = 0;
    it->entry = -1; /* the reference keeps entry_point; it is recomputed identically on the next call */
    memset(it->visited, 0, it->g->n + 1);
}
int vso_hnsw_bi_has_next(const vso_hnsw_bi *it) { return !(it->depleted && it->extras.n == 0); }

/* getNextResults (:206-249): ascending (score, label); returns the number written */
size_t vso_hnsw_bi_next(vso_hnsw_bi *it, size_t n_res, size_t label_count, size_t *labels, double *scores) {
    vso_hnsw *g = it->g;
    size_t orig_ef = it->ef;
    if (orig_ef < n_res) it->ef = n_res;
    if (it->returned == 0) {
        double cd;
        it->entry = bottom_ep(g, it->q, &cd);
    }
    heap_t top = {0};
    size_t n = 0;
    if (it->entry < 0) {
        it->depleted = 1;
    } else {
        if (it->returned == 0 && it->extras.n == 0 && it->cand.n == 0) {
            size_t ep = (size_t)it->entry;
            it->lower = g->deleted[ep] ? dist_max(g) : dist(g, ep, it->q);
            it->visited[ep] = 1;
            mh_push(&it->cand, it->lower, ep);
        }
        const int multi = g->multi; /* HNSWMulti_BatchIterator (hnsw_multi_batch_iterator.h:39-99): label-keyed result set,
                                       labels already handed out are skipped */
        while (top.n < it->ef && it->extras.n) { /* fillFromExtras */
            if (multi) {
                if (!bi_returned(it, it->extras.a[0].key)) ms_emplace(&top, it->extras.a[0].d, it->extras.a[0].key);
            } else {
                h_push(&top, it->extras.a[0].d, it->extras.a[0].key);
            }
            mh_pop(&it->extras);
        }
        if (top.n != it->ef) {
            while (it->cand.n) { /* scanGraphInternal */
                double cd = it->cand.a[0].d;
                size_t cur = it->cand.a[0].key;
                if (cd > it->lower && top.n >= it->ef) break;
                if (!g->deleted[cur]) { /* updateHeaps */
                    if (multi) {
                        if (it->lower > cd || top.n < it->ef) {
                            const size_t label = g->labels[cur];
                            if (!bi_returned(it, label)) {
                                ms_emplace(&top, cd, label);
                                if (top.n > it->ef) {
                                    const size_t m = ms_top(&top);
                                    mh_push(&it->extras, top.a[m].d, top.a[m].key);
                                    ms_pop(&top);
                                }
                                it->lower = top.a[ms_top(&top)].d;
                            }
                        }
                    } else if (top.n < it->ef) {
                        h_push(&top, cd, g->labels[cur]);
                        it->lower = top.a[0].d;
                    } else if (it->lower > cd) {
                        h_push(&top, cd, g->labels[cur]);
                        mh_push(&it->extras, top.a[0].d, top.a[0].key);
                        h_pop(&top);
                        it->lower = top.a[0].d;
                    }
                }
                mh_pop(&it->cand);
                const uint32_t *r = rec(g, cur, 0);
                for (uint32_t j = 0; j < r[0]; j++) {
                    size_t id = r[1 + j];
                    if (it->visited[id]) continue;
                    it->visited[id] = 1;
                    mh_push(&it->cand, dist(g, id, it->q), id);
                }
            }
            if (top.n < it->ef) it->depleted = 1;
        }
        if (multi) {
            while (top.n > n_res) { /* prepareResults */
                const size_t m = ms_top(&top);
                mh_push(&it->extras, top.a[m].d, top.a[m].key);
                ms_pop(&top);
            }
            n = top.n;
            for (size_t i = n; i-- > 0;) {
                const size_t m = ms_top(&top);
                scores[i] = top.a[m].d;
                labels[i] = top.a[m].key;
                bi_mark_returned(it, top.a[m].key);
                ms_pop(&top);
            }
        } else {
            while (top.n > n_res) { /* prepareResults */
                mh_push(&it->extras, top.a[0].d, top.a[0].key);
                h_pop(&top);
            }
            n = top.n;
            for (size_t i = n; i-- > 0;) {
                scores[i] = top.a[0].d;
                labels[i] = top.a[0].key;
                h_pop(&top);
            }
        }
    }
    free(top.a);
    it->returned += n;
    if (it->returned == label_count) it->depleted = 1;
    it->ef = orig_ef;
    return n;
}
