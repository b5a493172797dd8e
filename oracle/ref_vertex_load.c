/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * CPU restatement of the reference's `GenericGraph::vertex_load`
 *   (reference: src/generic_graph/generic_graph.rs:1310-1385)
 * in plain C.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The CUDA
 * product path never links, imports or calls it.
 *
 * Parity status: PINNED.  The restatement reproduces bit-for-bit the only
 * numeric golden vector the reference holds for this path
 *   (tests/integration_test_graph.rs:238-275: K4 -> [3;4]/[0;4], edgeless
 *    -> [0;4], the 8-vertex graph -> exact f64 vector),
 * see tests/test_oracle.py.  The Rust crate itself cannot be compiled in
 * this environment (no rustc/cargo, dependencies not vendored).
 *
 * The restatement is literal on purpose: same containers (two FIFO queues
 * swapped per level, an `ordering` stack, per-vertex predecessor vectors in
 * discovery order), same visiting order (neighbours in adjacency storage
 * order), same floating-point sequence (`b += b_k` then `b -= 1.0`;
 * `fraction = b_k / len as f64` pushed onto each predecessor in order).
 * That fixes the rounding, so results are bit-identical to the crate.
 *
 * Graph input: order-preserving CSR (row_ptr u64[n+1], col_idx u32[nnz]) —
 * the flattening of `container(i).neighbors()` (graph_traits.rs:72,
 * graph.rs:62-64, sw_graph.rs:195-199) for i in 0..n.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define NONE_DIST UINT64_MAX /* Option<usize>::None (generic_graph.rs:1317) */

typedef struct {
    uint32_t *data;
    uint32_t len, cap;
} vec_u32;

static void vec_push(vec_u32 *v, uint32_t x)
{
    if (v->len == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 4;
        v->data = (uint32_t *)realloc(v->data, (size_t)v->cap * sizeof(uint32_t));
    }
    v->data[v->len++] = x;
}

/* Per-thread scratch; mirrors the allocations at generic_graph.rs:1312-1318. */
typedef struct {
    uint32_t n;
    uint32_t *queue0, *queue1; /* VecDeque<usize> used strictly FIFO */
    uint32_t *ordering;        /* Vec<usize> used as a stack */
    double *b_k;
    uint64_t *distance;
    vec_u32 *predecessor;
} scratch_t;

static int scratch_init(scratch_t *s, uint32_t n)
{
    size_t m = n ? n : 1;
    s->n = n;
    s->queue0 = (uint32_t *)malloc(m * sizeof(uint32_t));
    s->queue1 = (uint32_t *)malloc(m * sizeof(uint32_t));
    s->ordering = (uint32_t *)malloc(m * sizeof(uint32_t));
    s->b_k = (double *)malloc(m * sizeof(double));
    s->distance = (uint64_t *)malloc(m * sizeof(uint64_t));
    s->predecessor = (vec_u32 *)calloc(m, sizeof(vec_u32));
    return s->queue0 && s->queue1 && s->ordering && s->b_k && s->distance && s->predecessor ? 0 : -1;
}

static void scratch_free(scratch_t *s)
{
    if (s->predecessor)
        for (uint32_t j = 0; j < s->n; ++j) free(s->predecessor[j].data);
    free(s->predecessor);
    free(s->queue0);
    free(s->queue1);
    free(s->ordering);
    free(s->b_k);
    free(s->distance);
    memset(s, 0, sizeof(*s));
}

/*
 * One iteration of the `for i in 0..vertex_count()` loop, generic_graph.rs:1321-1383.
 * Adds source i's contribution into b[].  Returns the number of directed adjacency
 * entries inspected (E_s) through *edges_seen and reached vertices through *reached.
 */
static void one_source(scratch_t *s, const uint64_t *row_ptr, const uint32_t *col_idx,
                       uint32_t i, int include_endpoints, double *b,
                       uint64_t *edges_seen, uint64_t *reached)
{
    const uint32_t n = s->n;
    /* :1324-1332 — the reference skips this for i == 0 because the vectors are
     * freshly constructed in exactly this state; doing it always is identical. */
    for (uint32_t j = 0; j < n; ++j) {
        s->b_k[j] = 1.0;
        s->distance[j] = NONE_DIST;
        s->predecessor[j].len = 0;
    }

    uint64_t depth = 0;                       /* :1335 */
    uint32_t *q0 = s->queue0, *q1 = s->queue1;
    uint32_t q0_head = 0, q0_tail = 0, q1_tail = 0;
    uint32_t ord_len = 0;
    uint64_t e_seen = 0;

    q0[q0_tail++] = i;                        /* :1336 */
    s->distance[i] = depth;                   /* :1337 */

    while (q0_head < q0_tail) {               /* :1341 pop_front */
        uint32_t index = q0[q0_head++];
        s->ordering[ord_len++] = index;       /* :1342 */
        for (uint64_t e = row_ptr[index]; e < row_ptr[index + 1]; ++e) { /* :1344 */
            uint32_t neighbor = col_idx[e];
            ++e_seen;
            uint64_t d = s->distance[neighbor];
            if (d != NONE_DIST) {             /* :1345 */
                if (d == depth + 1)           /* :1346 */
                    vec_push(&s->predecessor[neighbor], index); /* :1347 */
            } else {                          /* :1351 */
                s->distance[neighbor] = depth + 1;              /* :1352 */
                q1[q1_tail++] = neighbor;                       /* :1353 */
                vec_push(&s->predecessor[neighbor], index);     /* :1354 */
            }
        }
        if (q0_head == q0_tail) {             /* :1357 queue0.is_empty() */
            uint32_t *t = q0; q0 = q1; q1 = t; /* :1358 swap */
            q0_head = 0; q0_tail = q1_tail; q1_tail = 0;
            depth += 1;                       /* :1359 */
        }
    }
    if (reached) *reached += ord_len;
    if (edges_seen) *edges_seen += e_seen;

    while (ord_len > 0) {                     /* :1364 ordering.pop() */
        uint32_t index = s->ordering[--ord_len];
        if (ord_len == 0) break;              /* :1366-1368 skip last (= source) */
        b[index] += s->b_k[index];            /* :1371 */
        if (!include_endpoints) b[index] -= 1.0; /* :1372-1374 */
        const vec_u32 *p = &s->predecessor[index];
        double fraction = s->b_k[index] / (double)p->len; /* :1377 */
        for (uint32_t k = 0; k < p->len; ++k) /* :1378-1380 */
            s->b_k[p->data[k]] += fraction;
    }
}

/* vertex_load over an explicit list of sources, processed in the given order
 * (ascending 0..n-1 reproduces generic_graph.rs:1321 exactly).  out[n] is overwritten.
 * stats (optional, 2 x u64): [0] += sum_s E_s (directed entries inspected), [1] += sum_s |R_s|. */
int ref_vertex_load_sources(uint32_t n, const uint64_t *row_ptr, const uint32_t *col_idx,
                            const uint32_t *sources, uint32_t n_sources,
                            int include_endpoints, double *out, uint64_t *stats)
{
    scratch_t s;
    if (scratch_init(&s, n) != 0) { scratch_free(&s); return -1; }
    for (uint32_t j = 0; j < n; ++j) out[j] = 0.0; /* :1315 */
    uint64_t e = 0, r = 0;
    for (uint32_t k = 0; k < n_sources; ++k)
        one_source(&s, row_ptr, col_idx, sources[k], include_endpoints, out, &e, &r);
    if (stats) { stats[0] += e; stats[1] += r; }
    scratch_free(&s);
    return 0;
}

/* The reference entry point: all sources ascending. */
int ref_vertex_load(uint32_t n, const uint64_t *row_ptr, const uint32_t *col_idx,
                    int include_endpoints, double *out, uint64_t *stats)
{
    uint32_t *src = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    if (!src) return -1;
    for (uint32_t j = 0; j < n; ++j) src[j] = j;
    int rc = ref_vertex_load_sources(n, row_ptr, col_idx, src, n, include_endpoints, out, stats);
    free(src);
    return rc;
}

/*
 * Timing variant for the CPU baseline: the same per-source routine, sources split over
 * host threads (OpenMP, static chunks), per-thread private b[] reduced at the end in
 * thread order.  The crate itself is single-threaded; this is "all host cores" in
 * BASELINE.md section 3.  Summation order across threads differs from the serial code,
 * so this entry point is used for timing and for <=1e-9 comparisons, never for
 * bit-exact checks.  n_threads <= 0 means omp default.  Returns threads used (>0) or -1.
 */
int ref_vertex_load_sources_mt(uint32_t n, const uint64_t *row_ptr, const uint32_t *col_idx,
                               const uint32_t *sources, uint32_t n_sources,
                               int include_endpoints, int n_threads, double *out, uint64_t *stats)
{
    int used = 1;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
    used = n_threads;
#else
    n_threads = 1;
#endif
    double *partial = (double *)calloc((size_t)used * (n ? n : 1), sizeof(double));
    uint64_t *pst = (uint64_t *)calloc((size_t)used * 2, sizeof(uint64_t));
    int failed = 0;
    if (!partial || !pst) { free(partial); free(pst); return -1; }
#ifdef _OPENMP
#pragma omp parallel num_threads(used)
#endif
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        scratch_t s;
        if (scratch_init(&s, n) != 0) {
            failed = 1;
        } else {
            double *b = partial + (size_t)t * n;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
            for (int64_t k = 0; k < (int64_t)n_sources; ++k)
                one_source(&s, row_ptr, col_idx, sources[k], include_endpoints, b,
                           &pst[2 * t], &pst[2 * t + 1]);
        }
        scratch_free(&s);
    }
    for (uint32_t j = 0; j < n; ++j) {
        double acc = 0.0;
        for (int t = 0; t < used; ++t) acc += partial[(size_t)t * n + j];
        out[j] = acc;
    }
    if (stats)
        for (int t = 0; t < used; ++t) { stats[0] += pst[2 * t]; stats[1] += pst[2 * t + 1]; }
    free(partial);
    free(pst);
    return failed ? -1 : used;
}

/*
 * Per-source intermediate state, for exactness checks of the GPU path:
 *   dist[v]  = BFS distance from `source` (UINT32_MAX if unreached)   (:1337,:1352)
 *   npred[v] = predecessor[v].len()                                   (:1347,:1354)
 *   sigma[v] = number of shortest paths source->v as f64.  NOT used by the load value
 *              (the crate splits equally, :1377); provided because BASELINE.json's
 *              north_star asks for a path-count exactness check.
 *   bk[v]    = b_k[v] after the backward sweep (1.0 for unreached), optional.
 */
int ref_bfs_debug(uint32_t n, const uint64_t *row_ptr, const uint32_t *col_idx, uint32_t source,
                  uint32_t *dist, uint32_t *npred, double *sigma, double *bk)
{
    scratch_t s;
    if (source >= n) return -2;
    if (scratch_init(&s, n) != 0) { scratch_free(&s); return -1; }
    double *b = (double *)calloc(n, sizeof(double));
    if (!b) { scratch_free(&s); return -1; }

    /* sigma needs BFS visiting order: recompute it with the same traversal. */
    one_source(&s, row_ptr, col_idx, source, 1, b, NULL, NULL);
    for (uint32_t v = 0; v < n; ++v) {
        if (dist) dist[v] = s.distance[v] == NONE_DIST ? UINT32_MAX : (uint32_t)s.distance[v];
        if (npred) npred[v] = s.predecessor[v].len;
        if (bk) bk[v] = s.b_k[v];
    }
    if (sigma) {
        /* order vertices by distance (counting sort), then sigma[v] = sum over preds. */
        uint32_t maxd = 0;
        for (uint32_t v = 0; v < n; ++v)
            if (s.distance[v] != NONE_DIST && s.distance[v] > maxd) maxd = (uint32_t)s.distance[v];
        uint32_t *cnt = (uint32_t *)calloc((size_t)maxd + 2, sizeof(uint32_t));
        uint32_t *ord = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
        for (uint32_t v = 0; v < n; ++v)
            if (s.distance[v] != NONE_DIST) cnt[s.distance[v] + 1]++;
        for (uint32_t d = 0; d <= maxd; ++d) cnt[d + 1] += cnt[d];
        for (uint32_t v = 0; v < n; ++v)
            if (s.distance[v] != NONE_DIST) ord[cnt[s.distance[v]]++] = v;
        uint32_t reached = cnt[maxd];
        for (uint32_t v = 0; v < n; ++v) sigma[v] = 0.0;
        sigma[source] = 1.0;
        for (uint32_t k = 0; k < reached; ++k) {
            uint32_t v = ord[k];
            if (v == source) continue;
            double acc = 0.0;
            const vec_u32 *p = &s.predecessor[v];
            for (uint32_t t = 0; t < p->len; ++t) acc += sigma[p->data[t]];
            sigma[v] = acc;
        }
        free(cnt);
        free(ord);
    }
    free(b);
    scratch_free(&s);
    return 0;
}

int ref_omp_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
