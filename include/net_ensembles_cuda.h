/*
 * net_ensembles_cuda — C ABI of the B200-native `vertex_load` hot path.
 *
 * This header is the drop-in boundary: the entry points below are exactly what a
 * Rust `extern "C"` block (rust/net_ensembles_cuda.rs), a ctypes stub
 * (net_ensembles_b200/_lib.py) or the C++ mirror (include/net_ensembles_cuda.hpp)
 * bind.  Plain pointers and sizes only; no torch / C++ types cross this line.
 *
 * Reference interface replaced (all paths relative to the reference repo):
 *   GenericGraph::vertex_load(&self, include_endpoints: bool) -> Vec<f64>
 *       src/generic_graph/generic_graph.rs:1310-1385
 *   MeasurableGraphQuantities::vertex_load + blanket impl for every ensemble
 *       src/traits/graph_traits.rs:267, :272-331 (method at :328-330)
 * Input side (what the exporter flattens, once per sampled graph):
 *   container_iter()/neighbors()/degree()   generic_graph.rs:222, graph_traits.rs:72,75
 *   NodeContainer.adj: Vec<usize>           src/graph.rs:33-37, neighbors :62-64
 *   SwContainer.adj: Vec<SwEdge>            src/sw_graph.rs:158-162, neighbors :195-199
 *
 * Semantics (identical to the reference): *load* centrality — a vertex's accumulated
 * value is split EQUALLY among its BFS predecessors (generic_graph.rs:1377-1380); it
 * is not sigma-weighted Brandes betweenness.  Path counts (sigma) are available from
 * nec_bfs_debug only and never enter the load value.
 *
 * Conventions
 *   - every function returning int returns NEC_OK (0) on success, a negative NEC_E*
 *     code otherwise; nec_last_error(ctx) then describes the failure.
 *   - there is NO CPU fallback: without a usable CUDA device nec_init fails.
 *   - host pointers passed in are never retained past the call.
 *   - a context is not thread-safe; use one context per host thread.
 *   - graphs are simple undirected graphs given as an order-preserving CSR:
 *     row_ptr[n+1] (u64, row_ptr[0]==0, row_ptr[n]==nnz), col_idx[nnz] (u32 < n),
 *     every edge stored in both endpoint rows (as the crate's adjacency lists do).
 */
#ifndef NET_ENSEMBLES_CUDA_H
#define NET_ENSEMBLES_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NEC_OK 0
#define NEC_EINVAL (-1)   /* bad argument (null pointer, col_idx out of range, ...) */
#define NEC_ECUDA (-2)    /* CUDA runtime error (message in nec_last_error) */
#define NEC_ENOMEM (-3)   /* device or host allocation failed */
#define NEC_ENODEVICE (-4)/* no CUDA device / wrong architecture: no CPU fallback exists */
#define NEC_EINTERNAL (-5)/* kernel reported an inconsistency (queue overflow, ...) */

typedef struct nec_ctx nec_ctx;     /* opaque: device, stream, workspace arenas */
typedef struct nec_graph nec_graph; /* opaque: device-resident CSR of one sampled graph */

/* Library / build identification ("net_ensembles_cuda <version> sm_100a"). */
const char *nec_version(void);

/* Create a context on the listed CUDA devices.  n_devices == 1: one GPU (devices == NULL selects the current
 * device).  n_devices > 1 (devices == NULL: 0..n_devices-1): ONE process drives all of them — a worker thread,
 * stream and workspace per device, the graph replicated at upload, the sources of every nec_vertex_load* call
 * dealt to the devices in 64-source blocks, and the devices' partial load vectors combined with a single
 * ncclAllReduce(sum, f64, n) over NVLink (ncclCommInitAll; libnccl.so.2 is resolved at run time, only here).
 * nec_vertex_load on such a context is the reference's one call (graph_traits.rs:328-330) on all GPUs of the box;
 * nec_vertex_load_batch shards by graph and nec_distance_measures by source (nothing to combine).  The
 * one-process-per-GPU form (one single-device context per rank + the caller's own all-reduce on the buffer of
 * nec_vertex_load_sources_device) stays available, see INTEGRATION.md. */
int nec_init(nec_ctx **out, const int *devices, int n_devices);
void nec_destroy(nec_ctx *ctx);

/* Message for the last failing call on this context ("" if none).  ctx may be NULL:
 * then the message of the last failing nec_init on this thread is returned. */
const char *nec_last_error(const nec_ctx *ctx);

/* Run all work of this context on the given cudaStream_t (passed as void*); NULL
 * restores the context's own stream.  Lets a host framework (e.g. torch) time and
 * order the kernels on its current stream. */
int nec_set_stream(nec_ctx *ctx, void *cuda_stream);

/* Cap the scratch arena (bytes of HBM the engine may use for per-batch state);
 * 0 restores the default (80 % of free device memory at first use). */
int nec_set_workspace_limit(nec_ctx *ctx, uint64_t bytes);

/* Engine selection (tests and profiling; the default 0 chooses by graph size and structure):
 *   NEC_ENGINE_AUTO    0  n in [64, 180000]: one source per CTA, per-source state in shared
 *                         memory - unless n >= 8192, the call has >= 512 sources and a one-off
 *                         probe of the graph (kept with the nec_graph) finds that a vertex sits on
 *                         few distinct BFS levels within a batch (random, scale-free graphs): then,
 *                         and for n > 180000, the compact batched engine (64, 128 or 256 sources per
 *                         batch, a team of CTAs per batch); n < 64: the source-minor batched engine
 *                         (32 or 64 sources per CTA)
 *   NEC_ENGINE_BATCHED 1  always the source-minor batched engine
 *   NEC_ENGINE_MID     2  the one-source-per-CTA engine whenever n <= 180000
 *   NEC_ENGINE_WIDE    3  the compact batched engine whatever n */
#define NEC_ENGINE_AUTO 0
#define NEC_ENGINE_BATCHED 1
#define NEC_ENGINE_MID 2
#define NEC_ENGINE_WIDE 3
int nec_set_engine(nec_ctx *ctx, int engine);

/* Upload one graph (H2D copy of the CSR; validates row_ptr monotonicity and
 * col_idx < n).  Replaces the per-call traversal of the crate's adjacency lists. */
int nec_graph_upload(nec_ctx *ctx, uint32_t n, uint64_t nnz, const uint64_t *row_ptr,
                     const uint32_t *col_idx, nec_graph **out);
/* Same graph, handed over as the crate stores it: vertex v's neighbours are the `degrees[v]` elements of the
 * contiguous slice rows[v] (`AdjList::edges()`: &[usize] src/graph.rs:145-150, &[SwEdge] src/sw_graph.rs:164-170),
 * each element `elem_stride` bytes with the neighbour id (`to_width` = 4 or 8 bytes, little endian) at byte
 * `to_offset`.  The library flattens, narrows to u32, validates and stages the rows itself (all host cores, pinned
 * memory), preserving adjacency order: the export fast path that spares the caller a CSR in pageable memory. */
int nec_graph_upload_adjacency(nec_ctx *ctx, uint32_t n, const void *const *rows, const uint64_t *degrees,
                               uint32_t elem_stride, uint32_t to_offset, uint32_t to_width, nec_graph **out);
void nec_graph_free(nec_ctx *ctx, nec_graph *g);

/* Vertex and adjacency-entry counts of a device graph; its CSR back on the host (row_ptr[n+1], col_idx[nnz]). */
int nec_graph_dims(const nec_graph *g, uint32_t *n, uint64_t *nnz);
int nec_graph_download(nec_ctx *ctx, const nec_graph *g, uint64_t *row_ptr, uint32_t *col_idx);

/* ErEnsembleC::randomize on the device (src/er_c.rs:128-143): one `gen::<f64>() <= prob` per unordered pair in
 * lexicographic order, drawn from rand_pcg's Pcg64 by jump-ahead, so that the sampled graph - adjacency ORDER
 * included - and the generator state afterwards are bit-identical to what the crate produces from the same state
 * (pinned by TestData/unchaning_erc_{1,2}.json).  rng_in / rng_out: state low, state high, increment low,
 * increment high (u128 as two u64).  The graph is created device-resident (no host adjacency lists, no upload);
 * nec_graph_download fetches it if the host wants the lists.  Removes the host's O(n^2) loop from simple
 * sampling (benches/bench_vertex_load.rs:20-26: randomize, then vertex_load). */
int nec_erc_randomize(nec_ctx *ctx, uint32_t n, double prob, const uint64_t rng_in[4], uint64_t rng_out[4], nec_graph **out);

/* vertex_load(include_endpoints) over all sources: out[n] (host) is overwritten.
 * Synchronous: returns after the result is in `out`.
 * Replaces generic_graph.rs:1310 / graph_traits.rs:328-330. */
int nec_vertex_load(nec_ctx *ctx, const nec_graph *g, int include_endpoints, double *out);

/* Partial load: only the contributions of the listed sources (the `for i in ...` loop
 * body of generic_graph.rs:1321-1383 restricted to i in sources).  This is the
 * multi-GPU sharding primitive and the sampled-parity primitive.  sources are vertex
 * ids < n; duplicates count twice. */
int nec_vertex_load_sources(nec_ctx *ctx, const nec_graph *g, const uint32_t *sources,
                            uint32_t n_sources, int include_endpoints, double *out);

/* Same, but the result stays on the device: d_out is a device pointer to n doubles
 * (overwritten); no D2H of the result — the caller may hand d_out straight to
 * ncclAllReduce on the same stream.  The call still waits for the kernels in order to
 * read back a few bytes of per-batch status (errors are reported, never swallowed).
 * sources is a HOST pointer (copied during the call). */
int nec_vertex_load_sources_device(nec_ctx *ctx, const nec_graph *g, const uint32_t *sources,
                                   uint32_t n_sources, int include_endpoints, double *d_out);

/* Many small graphs in one launch (one CTA per graph, one warp per source, all state in
 * shared memory):
 * the shape of a Markov chain that measures after every step
 * (benches/common/mod.rs:79-90).  Graph i has n_per_graph[i] vertices
 * (1..NEC_BATCH_MAX_N); its row pointers are
 * row_ptr_concat[row_ptr_offsets[i] .. row_ptr_offsets[i]+n_i] (inclusive end, i.e.
 * n_i+1 entries, LOCAL: starting at 0) and its neighbour ids (local, < n_i) are
 * col_idx_concat[col_offsets[i] + row_ptr_local[...]].  out_concat receives the load
 * vectors back to back (sum n_i doubles).  Results are reproducible run to run and agree
 * with nec_vertex_load on each graph to rounding (sources are added in a different order). */
#define NEC_BATCH_MAX_N 512u
int nec_vertex_load_batch(nec_ctx *ctx, uint32_t n_graphs, const uint32_t *n_per_graph,
                          const uint64_t *row_ptr_offsets, const uint32_t *row_ptr_concat,
                          const uint64_t *col_offsets, const uint32_t *col_idx_concat,
                          int include_endpoints, double *out_concat);

/* Device-resident Markov chains (one CTA per graph, as above, but the graphs STAY on the device between
 * measurements).  A chain's step changes one or two edges (er_c.rs:154-174, er_m.rs:223-236, sw.rs:299-309), so
 * instead of re-exporting every CSR the host ships the LIST OPERATIONS it performed on its adjacency lists and the
 * device replays them in order — adjacency order (which defines the rounding of the load) stays identical:
 *     op = kind << 60 | v << 40 | x << 20 | y      (local vertex ids < 2^20)
 *     kind 1  push x onto v's list                         NodeContainer::push graph.rs:101-110, SwContainer::push sw_graph.rs:259-279
 *     kind 2  swap_remove the entry x from v's list        graph.rs:134-142 (Vec::swap_remove), sw_graph.rs:350-442
 *     kind 3  retarget the entry x of v's list to y        rewire / reset of a small-world root edge, sw_graph.rs:350-442
 * add_edge(a,b) = push(a,b), push(b,a); remove_edge(a,b) = swap_remove(a,b), swap_remove(b,a).
 * nec_chains_create uploads the initial graphs (same batch format as nec_vertex_load_batch; `slack` = spare
 * entries per adjacency row beyond the largest degree).  nec_chains_step_and_load applies graph g's operations
 * ops[op_offsets[g] .. op_offsets[g+1]) (op_offsets == NULL: measure without stepping), then runs
 * vertex_load(include_endpoints) on every graph; the load vectors arrive back to back in out_concat, or - with
 * out_concat == NULL - stay in the set's pinned host buffer, readable through nec_chains_result until the next
 * call.  A list that outgrows its row is reported (NEC_EINTERNAL): re-create the set with more slack.
 * nec_chains_download returns the device adjacency as a compact CSR batch (graph g's n_g + 1 local row pointers
 * start at sum_{h<g} (n_h + 1)) for checking it against the host's lists. */
typedef struct nec_chains nec_chains;
int nec_chains_create(nec_ctx *ctx, uint32_t n_graphs, const uint32_t *n_per_graph, const uint64_t *row_ptr_offsets,
                      const uint32_t *row_ptr_concat, const uint64_t *col_offsets, const uint32_t *col_idx_concat,
                      uint32_t slack, nec_chains **out);
int nec_chains_step_and_load(nec_ctx *ctx, nec_chains *chains, const uint32_t *op_offsets, const uint64_t *ops,
                             int include_endpoints, double *out_concat);
const double *nec_chains_result(const nec_chains *chains);
int nec_chains_download(nec_ctx *ctx, nec_chains *chains, uint32_t *row_ptr_concat, uint32_t *col_idx_concat,
                        uint64_t col_capacity, uint64_t *nnz);
void nec_chains_free(nec_ctx *ctx, nec_chains *chains);

/* Distance measures over the listed sources — the all-source BFS measures of the crate that sit
 * next to vertex_load in its benches (benches/common/mod.rs:53-64): for source sources[i]
 *   ecc[i]      = depth of the BFS from it   (longest_shortest_path_from_index, generic_graph.rs:1140-1144)
 *   sum_dist[i] = sum of the distances to every vertex it reaches (closeness_centrality, :1387-1403:
 *                 count[v] = sum_s d(s,v), which equals this sum for source v in an undirected graph)
 *   reached[i]  = number of vertices it reaches, itself included (connectivity test of diameter, :1116-1136)
 * Integer results, exact.  Forward (bit-parallel, 32 or 64 sources per pass) only; any output may be NULL. */
int nec_distance_measures(nec_ctx *ctx, const nec_graph *g, const uint32_t *sources, uint32_t n_sources,
                          uint32_t *ecc, uint64_t *sum_dist, uint32_t *reached);

/* How many sources the engine works on concurrently ("one resident wave": batch slots x sources per batch,
 * as planned for this graph, this device's free memory and a job of n_sources_total sources).  Callers that
 * sample sources or cut a job into calls should use multiples of it: a call with fewer sources leaves CTAs
 * idle, a call with slightly more runs a second, nearly empty wave. */
int nec_wave_size(nec_ctx *ctx, const nec_graph *g, uint64_t n_sources_total, uint32_t *sources_per_wave);

/* Intermediate state of one source, for exactness checks: dist[n] (UINT32_MAX =
 * unreached), npred[n] (number of BFS predecessors), sigma[n] (f64 shortest-path
 * counts; NOT used by the load), bk[n] (accumulated value b_k after the backward
 * sweep; 1.0 for unreached).  Any output pointer may be NULL. */
int nec_bfs_debug(nec_ctx *ctx, const nec_graph *g, uint32_t source, uint32_t *dist,
                  uint32_t *npred, double *sigma, double *bk);

/* Counters of the last nec_vertex_load* call on this context. */
typedef struct nec_stats {
    uint64_t n_sources;       /* sources processed */
    uint64_t n_batches;       /* batches (of batch_width sources) */
    uint64_t reached_pairs;   /* sum_s |R_s|  (vertices reached, source included) */
    uint64_t edge_pairs;      /* sum_s E_s    (directed adjacency entries in s's component) */
    uint64_t vertex_visits;   /* (vertex, level, batch) queue entries expanded, forward pass */
    uint32_t max_depth;       /* largest BFS depth seen */
    uint32_t kernel_launches; /* kernels of this library launched by the call */
    uint32_t wide_batches;    /* batches re-run with 32-bit distances (depth >= 255) */
    uint32_t n_slots;         /* resident batch slots (CTAs) used */
    float kernel_ms;          /* CUDA-event time of all kernels of the call (traversal + reduction) */
    float engine_ms;          /* CUDA-event time of the traversal kernel(s) alone (the dominant kernel) */
    float frac_reset;         /* share of CTA time spent resetting per-batch state ... */
    float frac_forward;       /* ... in the forward (BFS) pass ... */
    float frac_backward;      /* ... in the backward (accumulation) pass (clock64, summed over CTAs) */
    uint64_t workspace_bytes; /* scratch arena in use */
    uint32_t engine;          /* traversal kernel that ran: 1 batched (vertex_load_engine), 2 mid (mid_engine),
                                 3 small-graph batch (small_graph_kernel), 4 compact batched (wide_engine);
                                 first | second << 4 if one engine handed sources to another (1 | 2 << 4, 4 | 1 << 4) */
    uint32_t batch_width;     /* sources per batch of the batched engines in this call: 32, 64 or 128 (0: not used) */
    uint32_t team_size;       /* batched engine: CTAs per batch slot (thread-block cluster size; 1 = one CTA per slot) */
    uint32_t cta_threads;     /* threads per CTA of the traversal kernel that ran */
    uint32_t pull_levels;     /* batched engine: (batch, level) expansions done in pull direction ... */
    uint32_t push_levels;     /* ... and in push direction */
    uint32_t requeued_batches;/* batches re-run with the full level queue after an overflow of the memory-trimmed one */
    uint32_t reserved0;
} nec_stats;
int nec_get_stats(const nec_ctx *ctx, nec_stats *out);

#ifdef __cplusplus
}
#endif
#endif /* NET_ENSEMBLES_CUDA_H */
