/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see ref_vertex_load.c).
 *
 * CPU restatement of the reference's all-source BFS measures that SURVEY section 8(f) ranks first
 * after vertex_load:
 *   Bfs iterator (two queues, `handled` flags, depth)  src/generic_graph/iterators.rs:208-290
 *   GenericGraph::diameter                              src/generic_graph/generic_graph.rs:1116-1136
 *   GenericGraph::longest_shortest_path_from_index      src/generic_graph/generic_graph.rs:1140-1144
 *   GenericGraph::closeness_centrality                  src/generic_graph/generic_graph.rs:1387-1403
 *   GenericGraph::is_connected                          src/generic_graph/generic_graph.rs:735-741
 * Pinned by the reference's own known answers for diameter
 * (tests/integration_test_graph.rs:166-196: path5 -> 4, ring4 -> 2, ring40 -> 20, K40 -> 1),
 * see tests/test_measures.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint32_t n;
    const uint64_t *row_ptr;
    const uint32_t *col_idx;
    uint8_t *handled;
    uint32_t *queue0, *queue1;
    uint32_t q0_head, q0_tail, q1_tail;
    uint64_t depth;
} bfs_t;

static int bfs_alloc(bfs_t *b, uint32_t n, const uint64_t *row_ptr, const uint32_t *col_idx)
{
    size_t m = n ? n : 1;
    memset(b, 0, sizeof(*b));
    b->n = n; b->row_ptr = row_ptr; b->col_idx = col_idx;
    b->handled = (uint8_t *)malloc(m);
    b->queue0 = (uint32_t *)malloc(m * sizeof(uint32_t));
    b->queue1 = (uint32_t *)malloc(m * sizeof(uint32_t));
    return b->handled && b->queue0 && b->queue1 ? 0 : -1;
}
static void bfs_free(bfs_t *b) { free(b->handled); free(b->queue0); free(b->queue1); }

/* Bfs::reuse (iterators.rs:247-259) */
static void bfs_reuse(bfs_t *b, uint32_t index)
{
    memset(b->handled, 0, b->n ? b->n : 1);
    b->q0_head = b->q0_tail = b->q1_tail = 0;
    b->depth = 0;
    if (index < b->n) { b->queue0[b->q0_tail++] = index; b->handled[index] = 1; }
}

/* Iterator::next (iterators.rs:268-285): returns 1 and (index, depth), or 0 when exhausted */
static int bfs_next(bfs_t *b, uint32_t *index, uint64_t *depth)
{
    for (;;) {
        if (b->q0_head < b->q0_tail) {
            uint32_t v = b->queue0[b->q0_head++];
            for (uint64_t e = b->row_ptr[v]; e < b->row_ptr[v + 1]; ++e) {
                uint32_t i = b->col_idx[e];
                if (!b->handled[i]) { b->handled[i] = 1; b->queue1[b->q1_tail++] = i; }
            }
            *index = v; *depth = b->depth;
            return 1;
        }
        if (b->q1_tail == 0) return 0;
        uint32_t *t = b->queue0; b->queue0 = b->queue1; b->queue1 = t;
        b->q0_head = 0; b->q0_tail = b->q1_tail; b->q1_tail = 0;
        b->depth += 1;
    }
}

/* per source: eccentricity, sum of distances, reached count (what the three measures are made of) */
int ref_distance_measures(uint32_t n, const uint64_t *row_ptr, const uint32_t *col_idx, const uint32_t *sources,
                          uint32_t n_sources, uint32_t *ecc, uint64_t *sum_dist, uint32_t *reached)
{
    bfs_t b;
    if (bfs_alloc(&b, n, row_ptr, col_idx) != 0) { bfs_free(&b); return -1; }
    for (uint32_t k = 0; k < n_sources; ++k) {
        uint32_t idx, cnt = 0; uint64_t d, last = 0, sum = 0;
        bfs_reuse(&b, sources[k]);
        while (bfs_next(&b, &idx, &d)) { last = d; sum += d; ++cnt; }
        if (ecc) ecc[k] = (uint32_t)last;
        if (sum_dist) sum_dist[k] = sum;
        if (reached) reached[k] = cnt;
    }
    bfs_free(&b);
    return 0;
}

/* diameter (generic_graph.rs:1116-1136): -1 = None (empty or not connected) */
int64_t ref_diameter(uint32_t n, const uint64_t *row_ptr, const uint32_t *col_idx)
{
    bfs_t b;
    if (n == 0) return -1;                                    /* is_connected()? -> None */
    if (bfs_alloc(&b, n, row_ptr, col_idx) != 0) { bfs_free(&b); return -2; }
    uint32_t idx, cnt = 0; uint64_t d;
    bfs_reuse(&b, 0);                                         /* is_connected: everything reachable from 0 */
    while (bfs_next(&b, &idx, &d)) ++cnt;
    int64_t result = -1;
    if (cnt == n) {
        uint64_t max = 0;
        for (uint32_t index = 1; index < n; ++index) {        /* "from every node (except 1 node)" */
            uint64_t depth = 0;
            bfs_reuse(&b, index);
            while (bfs_next(&b, &idx, &d)) depth = d;
            if (depth > max) max = depth;
        }
        result = (int64_t)max;
    }
    bfs_free(&b);
    return result;
}

/* closeness_centrality (generic_graph.rs:1387-1403) */
int ref_closeness_centrality(uint32_t n, const uint64_t *row_ptr, const uint32_t *col_idx, double *out)
{
    bfs_t b;
    if (bfs_alloc(&b, n, row_ptr, col_idx) != 0) { bfs_free(&b); return -1; }
    uint64_t *count = (uint64_t *)calloc(n ? n : 1, sizeof(uint64_t));
    if (!count) { bfs_free(&b); return -1; }
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t idx; uint64_t d;
        bfs_reuse(&b, i);
        while (bfs_next(&b, &idx, &d)) count[idx] += d;
    }
    double val = (double)(n - 1);
    for (uint32_t v = 0; v < n; ++v) out[v] = val / (double)count[v];
    free(count);
    bfs_free(&b);
    return 0;
}
