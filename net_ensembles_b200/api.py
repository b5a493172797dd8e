"""Python host side of the vertex_load path: mirrors the reference's operator interface
(`vertex_load(include_endpoints)` on graphs and ensembles — /root/reference
src/generic_graph/generic_graph.rs:1310, src/traits/graph_traits.rs:267,272-331) over the
C ABI.  Everything numeric happens in the CUDA library; this file only moves buffers.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import NecStats, c_int, c_u32, c_u64, c_vp  # noqa: F401


class NecError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("net_ensembles_cuda error %d: %s" % (code, message))
        self.code = code


def _ptr(a):
    return a.ctypes.data_as(c_vp) if a is not None else None


class Context:
    """One GPU, one stream, one workspace (nec_ctx) — or, with `devices=[...]`, ONE context over several GPUs of
    the box: every vertex_load call then uses all of them (sources dealt to the devices, one NCCL all-reduce
    behind the C ABI).  Not thread-safe."""

    def __init__(self, device=None, devices=None):
        self._lib = _lib.load()
        h = c_vp()
        if devices is not None:
            devices = [int(d) for d in devices]
            dev, count = (c_int * len(devices))(*devices), len(devices)
        else:
            dev, count = ((c_int * 1)(device) if device is not None else None), 1
        self.n_devices = count
        rc = self._lib.nec_init(ctypes.byref(h), dev, count)
        if rc != 0:
            raise NecError(rc, self._lib.nec_last_error(None).decode())
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.nec_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc):
        if rc != 0:
            raise NecError(rc, self._lib.nec_last_error(self._h).decode())

    def set_stream(self, cuda_stream):
        """cuda_stream: integer cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream) or None."""
        self._check(self._lib.nec_set_stream(self._h, c_vp(cuda_stream) if cuda_stream else None))

    def set_workspace_limit(self, nbytes):
        self._check(self._lib.nec_set_workspace_limit(self._h, int(nbytes)))

    def set_engine(self, engine):
        """0 auto, 1 source-minor batched (32 or 64 sources per CTA), 2 mid (one source per CTA, default for 64 <= n <= 180000
        unless a one-off structural probe of the graph prefers 3), 3 compact batched (64 / 128 / 256 sources per team of CTAs,
        default for n > 180000)."""
        self._check(self._lib.nec_set_engine(self._h, int(engine)))

    def stats(self):
        s = NecStats()
        self._check(self._lib.nec_get_stats(self._h, ctypes.byref(s)))
        return s.as_dict()

    def upload(self, row_ptr, col_idx):
        return DeviceGraph(self, row_ptr, col_idx)


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


class DeviceGraph:
    """Device-resident CSR of one sampled graph (nec_graph): export once, measure many times."""

    def __init__(self, ctx, row_ptr=None, col_idx=None, _host=None, _adopt=None):
        self.ctx = ctx
        if _adopt is not None:             # a graph the library created on the device (nec_erc_randomize)
            n, nnz = c_u32(0), c_u64(0)
            ctx._lib.nec_graph_dims(_adopt, ctypes.byref(n), ctypes.byref(nnz))
            self.n, self.nnz, self._h = int(n.value), int(nnz.value), _adopt
            return
        if _host is not None:
            # straight from a host mirror's adjacency lists (export fast path: no CSR is built on the Python side)
            self.n, self.nnz = _host.vertex_count(), 2 * _host.edge_count()
            h = c_vp()
            ctx._check(ctx._lib.nech_upload_cuda(_host._h, ctx._h, ctypes.byref(h)))
            self._h = h
            return
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        col_idx = np.ascontiguousarray(col_idx, dtype=np.uint32)
        self.n = int(row_ptr.shape[0]) - 1
        self.nnz = int(col_idx.shape[0])
        h = c_vp()
        ctx._check(ctx._lib.nec_graph_upload(ctx._h, self.n, self.nnz, _ptr(row_ptr), _ptr(col_idx), ctypes.byref(h)))
        self._h = h

    def free(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            self.ctx._lib.nec_graph_free(self.ctx._h, self._h)
        self._h = None

    __del__ = free

    def vertex_load(self, include_endpoints):
        out = np.empty(self.n, dtype=np.float64)
        self.ctx._check(self.ctx._lib.nec_vertex_load(self.ctx._h, self._h, int(bool(include_endpoints)), _ptr(out)))
        return out

    def vertex_load_sources(self, sources, include_endpoints):
        src = np.ascontiguousarray(sources, dtype=np.uint32)
        out = np.empty(self.n, dtype=np.float64)
        self.ctx._check(self.ctx._lib.nec_vertex_load_sources(self.ctx._h, self._h, _ptr(src), src.shape[0],
                                                              int(bool(include_endpoints)), _ptr(out)))
        return out

    def vertex_load_sources_device(self, sources, include_endpoints, d_out_ptr):
        """Partial load into a caller-owned device buffer of n doubles (integer device pointer)."""
        src = np.ascontiguousarray(sources, dtype=np.uint32)
        self.ctx._check(self.ctx._lib.nec_vertex_load_sources_device(self.ctx._h, self._h, _ptr(src), src.shape[0],
                                                                     int(bool(include_endpoints)), c_vp(d_out_ptr)))

    def wave_size(self, n_sources_total=None):
        """Sources the engine works on concurrently for a job of `n_sources_total` (default: all vertices)."""
        w = ctypes.c_uint32(0)
        total = self.n if n_sources_total is None else int(n_sources_total)
        self.ctx._check(self.ctx._lib.nec_wave_size(self.ctx._h, self._h, total, ctypes.byref(w)))
        return int(w.value)

    def distance_measures(self, sources=None):
        """(ecc, sum_dist, reached) per listed source (all vertices by default); exact integers."""
        src = np.arange(self.n, dtype=np.uint32) if sources is None else np.ascontiguousarray(sources, dtype=np.uint32)
        k = src.shape[0]
        ecc, sumd, reached = np.zeros(k, np.uint32), np.zeros(k, np.uint64), np.zeros(k, np.uint32)
        self.ctx._check(self.ctx._lib.nec_distance_measures(self.ctx._h, self._h, _ptr(src), k, _ptr(ecc), _ptr(sumd), _ptr(reached)))
        return ecc, sumd, reached

    def diameter(self):
        """`GenericGraph::diameter` (generic_graph.rs:1116-1136): None if empty or not connected."""
        if self.n == 0:
            return None
        ecc, _, reached = self.distance_measures()
        if int(reached[0]) != self.n:
            return None
        return int(ecc[1:].max()) if self.n > 1 else 0          # the crate skips vertex 0; same maximum

    def longest_shortest_path_from_index(self, index):
        """generic_graph.rs:1140-1144."""
        if not 0 <= index < self.n:
            return None
        return int(self.distance_measures([index])[0][0])

    def closeness_centrality(self):
        """`GenericGraph::closeness_centrality` (generic_graph.rs:1387-1403): (N-1) / sum_s d(s, v);
        the sum over sources of the distance TO v equals the sum of distances FROM v (undirected)."""
        _, sumd, _ = self.distance_measures()
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.float64(self.n - 1) / sumd.astype(np.float64)

    def download(self):
        """(row_ptr u64[n+1], col_idx u32[nnz]) of the device graph."""
        rp, col = np.empty(self.n + 1, np.uint64), np.empty(self.nnz, np.uint32)
        self.ctx._check(self.ctx._lib.nec_graph_download(self.ctx._h, self._h, _ptr(rp), _ptr(col)))
        return rp, col

    def bfs_debug(self, source):
        n = self.n
        dist, npred = np.empty(n, np.uint32), np.empty(n, np.uint32)
        sigma, bk = np.empty(n, np.float64), np.empty(n, np.float64)
        self.ctx._check(self.ctx._lib.nec_bfs_debug(self.ctx._h, self._h, int(source), _ptr(dist), _ptr(npred), _ptr(sigma), _ptr(bk)))
        return dist, npred, sigma, bk


def vertex_load_batch(ctx, graphs, include_endpoints):
    """graphs: list of (row_ptr, col_idx) with LOCAL ids, each n <= NEC_BATCH_MAX_N.
    One launch, one CTA per graph.  Returns a list of load vectors."""
    ns = np.array([len(rp) - 1 for rp, _ in graphs], dtype=np.uint32)
    rp_off = np.zeros(len(graphs), dtype=np.uint64)
    col_off = np.zeros(len(graphs), dtype=np.uint64)
    if len(graphs):
        rp_off[1:] = np.cumsum(ns[:-1].astype(np.uint64) + 1)
        col_off[1:] = np.cumsum([len(c) for _, c in graphs][:-1])
    rp_cat = np.concatenate([np.asarray(rp, dtype=np.uint32) for rp, _ in graphs]) if len(graphs) else np.zeros(0, np.uint32)
    col_cat = np.concatenate([np.asarray(c, dtype=np.uint32) for _, c in graphs] + [np.zeros(0, np.uint32)])
    out = np.empty(int(ns.sum()), dtype=np.float64)
    ctx._check(ctx._lib.nec_vertex_load_batch(ctx._h, len(graphs), _ptr(ns), _ptr(rp_off), _ptr(np.ascontiguousarray(rp_cat)),
                                              _ptr(col_off), _ptr(np.ascontiguousarray(col_cat)),
                                              int(bool(include_endpoints)), _ptr(out)))
    return np.split(out, np.cumsum(ns)[:-1]) if len(graphs) else []


class ChainBatch:
    """A fixed set of Markov chains measured together (benches/common/mod.rs:79-90 run on many chains).

    device_resident=True (default): the chains' adjacency lists are uploaded ONCE into slack-padded device rows;
    a measurement step runs `m_steps` Markov steps on every host chain (all host cores), ships only the list
    operations they performed (a few bytes per chain) and replays them on the device copy, so adjacency order -
    hence every rounding - stays identical to the host's; then vertex_load of all graphs, one CTA per graph.
    device_resident=False: every step re-exports all CSRs and uploads them again (nec_vertex_load_batch)."""

    def __init__(self, chains, device_resident=True):
        self.chains = list(chains)
        self._handles = (c_vp * len(self.chains))(*[c._h for c in self.chains])
        self._ns = np.array([c.vertex_count() for c in self.chains], dtype=np.int64)
        self._out = np.empty(int(self._ns.sum()), dtype=np.float64)
        self._uniform = len(self.chains) > 0 and bool((self._ns == self._ns[0]).all())
        self._cuts = np.cumsum(self._ns)[:-1]
        self.device_resident = bool(device_resident) and len(self.chains) > 0 and int(self._ns.max(initial=0)) <= _lib.NEC_BATCH_MAX_N
        self.mode = ("device-resident adjacency rows; a step ships the chains' list operations" if self.device_resident
                     else "CSR re-exported and re-uploaded every step")
        self._dev, self._dev_ctx = None, None
        self.h2d_bytes_last_step = None
        self.uploads = 0
        self.stats = None

    def _shape(self, flat):
        if self._uniform:
            return flat.reshape(len(self.chains), int(self._ns[0]))
        return np.split(flat, self._cuts)

    def close(self):
        if self._dev is not None and getattr(self._dev_ctx, "_h", None):
            self._dev_ctx._lib.nech_chain_batch_free(self._dev)
        self._dev = None

    __del__ = close

    def step_and_load(self, ctx, m_steps, include_endpoints):
        """Returns the load vectors: a (chains, n) array when all graphs have the same size, else a list of
        views.  The buffer is reused by the next call."""
        lib = ctx._lib
        if not self.chains:
            return []
        if not self.device_resident:
            ctx._check(lib.nech_chains_step_and_load_cuda(self._handles, len(self.chains), int(m_steps), ctx._h,
                                                          int(bool(include_endpoints)), _ptr(self._out)))
            st = _lib.NecStats()
            lib.nech_last_chain_stats(ctypes.byref(st))
            self.stats = st.as_dict()          # counters of this step
            return self._shape(self._out)
        if self._dev is None or self._dev_ctx is not ctx:
            self.close()
            h = lib.nech_chain_batch_new(self._handles, len(self.chains), ctx._h)
            if not h:
                raise NecError(-1, lib.nec_last_error(ctx._h).decode())
            self._dev, self._dev_ctx = c_vp(h), ctx
        ctx._check(lib.nech_chain_batch_step(self._dev, int(m_steps), int(bool(include_endpoints)), None))
        st, h2d, ups = _lib.NecStats(), c_u64(0), c_u64(0)
        lib.nech_chain_batch_stats(self._dev, ctypes.byref(st), ctypes.byref(h2d), ctypes.byref(ups))
        self.stats, self.h2d_bytes_last_step, self.uploads = st.as_dict(), int(h2d.value), int(ups.value)
        ptr = lib.nech_chain_batch_result(self._dev)         # the library's pinned result buffer: no extra copy
        flat = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_double)), shape=(int(self._ns.sum()),))
        return self._shape(flat)

    def device_adjacency(self):
        """The device copy's adjacency lists, graph by graph (order included) - for checking them against the hosts'."""
        if self._dev is None:
            raise ValueError("no device copy yet: call step_and_load first")
        total_rp = int((self._ns + 1).sum())
        cap = int(sum(2 * c.edge_count() for c in self.chains)) + 16
        rp, col, nnz = np.zeros(total_rp, np.uint32), np.zeros(cap, np.uint32), c_u64(0)
        self._dev_ctx._check(self._dev_ctx._lib.nech_chain_batch_download(self._dev, _ptr(rp), _ptr(col), cap, ctypes.byref(nnz)))
        out, at_rp, at_col = [], 0, 0
        for n in self._ns:
            r = rp[at_rp:at_rp + int(n) + 1]
            out.append([col[at_col + int(r[v]):at_col + int(r[v + 1])].tolist() for v in range(int(n))])
            at_rp += int(n) + 1
            at_col += int(r[int(n)])
        return out


def chains_step_and_load(ctx, chains, m_steps, include_endpoints):
    """Markov-chain measurement (benches/common/mod.rs:79-90): `m_steps` Markov steps on every
    chain, then vertex_load of all their graphs (export + call happen in C++).  Convenience form of
    ChainBatch for a one-off call; returns a list of load vectors (copies)."""
    res = ChainBatch(chains).step_and_load(ctx, m_steps, include_endpoints)
    return [np.array(r) for r in res]


class _HostGraph:
    """Base of the host-side mirrors: owns a nech handle (adjacency lists live in C++)."""

    def __init__(self, handle):
        self._lib = _lib.load_host()       # input side only: no CUDA library is needed to build or step graphs
        if not handle:
            raise ValueError(self._lib.nech_last_error().decode())
        self._h = c_vp(handle)

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.nech_free(self._h)
            self._h = None

    def vertex_count(self):
        return int(self._lib.nech_vertex_count(self._h))

    def edge_count(self):
        return int(self._lib.nech_edge_count(self._h))

    def add_edge(self, a, b):
        """True if added, False if the edge exists (GraphErrors::EdgeExists)."""
        rc = self._lib.nech_add_edge(self._h, a, b)
        if rc < 0:
            raise ValueError(self._lib.nech_last_error().decode())
        return bool(rc)

    def remove_edge(self, a, b):
        rc = self._lib.nech_remove_edge(self._h, a, b)
        if rc < 0:
            raise ValueError(self._lib.nech_last_error().decode())
        return bool(rc)

    def export_csr(self):
        """Order-preserving CSR (row_ptr u64[n+1], col_idx u32[nnz]) of the adjacency lists."""
        n, nnz = self.vertex_count(), int(self._lib.nech_nnz(self._h))
        row_ptr, col = np.empty(n + 1, np.uint64), np.empty(nnz, np.uint32)
        self._lib.nech_export_csr(self._h, _ptr(row_ptr), _ptr(col))
        return row_ptr, col

    def adjacency(self):
        rp, col = self.export_csr()
        return [col[int(rp[i]):int(rp[i + 1])].tolist() for i in range(len(rp) - 1)]

    def vertex_load(self, include_endpoints, ctx=None):
        """Drop-in for `vertex_load(include_endpoints)`: export to CSR, run on the GPU, return
        a fresh f64 vector.  Raises NecError on any device failure (no CPU fallback)."""
        ctx = ctx or default_context()
        out = np.empty(self.vertex_count(), dtype=np.float64)
        ctx._check(ctx._lib.nech_vertex_load_cuda(self._h, ctx._h, int(bool(include_endpoints)), _ptr(out)))
        return out

    def to_device(self, ctx=None):
        """Device-resident CSR of the current adjacency lists ("export once per sampled graph")."""
        ctx = ctx or default_context()
        return DeviceGraph(ctx, _host=self)

    # the all-source BFS measures that sit beside vertex_load in the crate's benches
    def diameter(self, ctx=None):
        dg = self.to_device(ctx)
        try:
            return dg.diameter()
        finally:
            dg.free()

    def closeness_centrality(self, ctx=None):
        dg = self.to_device(ctx)
        try:
            return dg.closeness_centrality()
        finally:
            dg.free()

    def longest_shortest_path_from_index(self, index, ctx=None):
        dg = self.to_device(ctx)
        try:
            return dg.longest_shortest_path_from_index(index)
        finally:
            dg.free()


class Graph(_HostGraph):
    """`Graph<T>` = GenericGraph with plain adjacency lists (src/graph.rs:309)."""

    def __init__(self, n, _handle=None):
        lib = _lib.load_host()
        super().__init__(_handle if _handle is not None else lib.nech_graph_new(n))


class _Ensemble(_HostGraph):
    def graph(self):
        return self

    def randomize(self):
        if self._lib.nech_randomize(self._h) != 0:
            raise ValueError(self._lib.nech_last_error().decode())

    def m_steps(self, count):
        if self._lib.nech_m_steps(self._h, count) != 0:
            raise ValueError(self._lib.nech_last_error().decode())

    def m_step(self):
        self.m_steps(1)

    def rng_state(self):
        """(state, increment) of the Pcg64 as Python ints."""
        s = np.zeros(4, dtype=np.uint64)
        self._lib.nech_rng_state(self._h, _ptr(s))
        return int(s[0]) | (int(s[1]) << 64), int(s[2]) | (int(s[3]) << 64)


class ErEnsembleC(_Ensemble):
    """ErEnsembleC::new(n, c_target, Pcg64::seed_from_u64(seed)) (src/er_c.rs:239)."""

    def __init__(self, n, c_target, seed):
        super().__init__(_lib.load_host().nech_erc_new(n, c_target, seed))
        self._params = {"prob": float(c_target) / (n - 1) if n > 1 else 0.0, "c_target": float(c_target)}   # er_c.rs: prob = c_target / (n - 1)

    def randomize_cuda(self, ctx=None):
        """`randomize()` (src/er_c.rs:128-143) drawn ON THE DEVICE with Pcg64 jump-ahead: same graph, same adjacency
        order and same generator state afterwards as the host loop, but the sample is born device-resident.  Returns
        its DeviceGraph (measure it, then free it); the host lists are brought in step as well."""
        ctx = ctx or default_context()
        h = c_vp()
        ctx._check(ctx._lib.nech_erc_randomize_cuda(self._h, ctx._h, ctypes.byref(h)))
        return DeviceGraph(ctx, _adopt=h)


class ErEnsembleM(_Ensemble):
    """ErEnsembleM::new(n, m, rng) (src/er_m.rs:245); O(n^2) memory like the crate."""

    def __init__(self, n, m, seed):
        super().__init__(_lib.load_host().nech_erm_new(n, m, seed))
        self._params = {"m": int(m)}


class SwEnsemble(_Ensemble):
    """SwEnsemble::new(n, r_prob, rng) / new_with_distance (src/sw.rs:214,234)."""

    def __init__(self, n, r_prob, seed, ring_distance=2):
        super().__init__(_lib.load_host().nech_sw_new(n, r_prob, seed, ring_distance))
        self._params = {"r_prob": float(r_prob)}

    def root_targets(self):
        out = np.empty(int(self._lib.nech_nnz(self._h)), np.uint32)
        self._lib.nech_sw_root_targets(self._h, _ptr(out))
        return out


class BAensemble(_Ensemble):
    """BAensemble::new(n, rng, m, source_n) (src/barabasi_albert.rs:89); O(n^2) like the crate."""

    def __init__(self, n, seed, m, source_n):
        super().__init__(_lib.load_host().nech_ba_new(n, m, source_n, seed))


def gnm_scalable(n, m, seed):
    """G(n,m) for sizes the crate's ErEnsembleM cannot reach; not seed-identical to the crate."""
    return Graph(n, _handle=_lib.load_host().nech_gnm_scalable(n, m, seed))


def ba_scalable(n, m, source_n, seed):
    """Barabasi-Albert for sizes the crate's BAensemble cannot reach; not seed-identical."""
    return Graph(n, _handle=_lib.load_host().nech_ba_scalable(n, m, source_n, seed))


# ------------------------------------------------------------------------------------------------
# The crate's serde save-states (`serde_support`, e.g. generic_graph.rs:31-32, er_c.rs:71-80; lib.rs:147-180 shows the
# save / resume idiom; TestData/unchaning_*.json are such dumps).  The format on the INPUT side of the path: a topology
# stored by the crate goes to the GPU with its adjacency order intact.
# ------------------------------------------------------------------------------------------------
class SavedState:
    """What a serde_json dump of a Graph / ErEnsembleC / ErEnsembleM / SwEnsemble holds for this path.

    kind         "graph" | "erc" | "erm" | "sw"
    graph        host `Graph` with exactly the stored adjacency lists (order preserved); `.vertex_load(...)`, `.to_device(ctx)`
    params       the ensemble's parameters as stored (prob / c_target, m, r_prob)
    rng          (state, increment) of the Pcg64 as Python ints, or None
    root_targets small-world only: per adjacency entry the `originally_to` of a root edge, -1 for a loose entry (CSR order)
    """

    def __init__(self, kind, graph, params, rng, root_targets, erm_lists=None):
        self.kind, self.graph, self.params, self.rng, self.root_targets = kind, graph, params, rng, root_targets
        self.erm_lists = erm_lists          # ErEnsembleM: (all_edges, possible_edges, current_edges) as (k, 2) uint32 arrays, or None

    def vertex_load(self, include_endpoints, ctx=None):
        return self.graph.vertex_load(include_endpoints, ctx)

    def resume(self):
        """A live ensemble of this package in exactly the saved state - lists, parameters, generator, ErM's edge lists, Sw's root
        edges - so that `m_steps` / `randomize` continue as the crate's own would (lib.rs:147-180, the resume idiom).  Nothing
        is drawn on the way."""
        if self.kind not in ("erc", "erm", "sw") or self.rng is None:
            raise ValueError("resume: the save-state holds no ensemble with a generator (kind %r)" % self.kind)
        rp, col = self.graph.export_csr()
        n = len(rp) - 1
        mask = (1 << 64) - 1
        rng4 = np.array([self.rng[0] & mask, self.rng[0] >> 64, self.rng[1] & mask, self.rng[1] >> 64], dtype=np.uint64)
        lib = _lib.load_host()
        none = None
        if self.kind == "erc":
            cls, params = ErEnsembleC, {"prob": float(self.params["prob"]), "c_target": float(self.params["c_target"])}
            h = lib.nech_restore(1, n, _ptr(rp), _ptr(col), none, params["c_target"], params["prob"], 0, _ptr(rng4), none, 0, none, 0, none, 0)
        elif self.kind == "erm":
            if self.erm_lists is None:
                raise ValueError("resume: an ErEnsembleM needs its all_edges / possible_edges / current_edges lists")
            cls, params = ErEnsembleM, {"m": int(self.params["m"])}
            al, po, cu = (np.ascontiguousarray(x, dtype=np.uint32).reshape(-1, 2) for x in self.erm_lists)
            h = lib.nech_restore(2, n, _ptr(rp), _ptr(col), none, 0.0, 0.0, params["m"], _ptr(rng4),
                                 _ptr(al), len(al), _ptr(po), len(po), _ptr(cu), len(cu))
        else:
            cls, params = SwEnsemble, {"r_prob": float(self.params["r_prob"])}
            roots = np.where(self.root_targets < 0, 0xFFFFFFFF, self.root_targets).astype(np.uint32)
            h = lib.nech_restore(3, n, _ptr(rp), _ptr(col), _ptr(roots), params["r_prob"], 0.0, 0, _ptr(rng4), none, 0, none, 0, none, 0)
        if not h:
            raise ValueError(lib.nech_last_error().decode())
        e = cls.__new__(cls)
        _HostGraph.__init__(e, h)
        e._params = params
        return e


def graph_from_adjacency(adjacency):
    """Host `Graph` whose adjacency lists are exactly `adjacency` (list of lists of neighbour ids), in that order.
    Raises ValueError on ids out of range, self-loops, repeated entries or entries without a counterpart."""
    n = len(adjacency)
    row_ptr = np.zeros(n + 1, np.uint64)
    row_ptr[1:] = np.cumsum([len(a) for a in adjacency], dtype=np.uint64)
    flat = [x for a in adjacency for x in a]
    if any((not isinstance(x, (int, np.integer))) or x < 0 or x >= max(n, 1) for x in flat):
        raise ValueError("graph_from_adjacency: neighbour id out of range")
    col = np.asarray(flat, dtype=np.uint32) if flat else np.zeros(0, np.uint32)
    lib = _lib.load_host()
    handle = lib.nech_graph_from_adjacency(n, _ptr(row_ptr), _ptr(col))
    if not handle:
        raise ValueError(lib.nech_last_error().decode())
    return Graph(n, _handle=handle)


def load_serde_json(src):
    """Parses a serde_json dump of the crate (`src`: path, open file, JSON text or the parsed dict) into a SavedState.
    Layout: {"graph": {"next_id", "edge_count", "vertices": [{"id", "adj": [...], "node": ...}], "phantom"},
    <parameters>, "rng": {"state", "increment"}}; a bare Graph is the inner object alone; small-world adjacency entries are
    {"to": k, "originally_to": null | k0}.  Unknown keys (ErM's edge lists, node payloads) are ignored."""
    import json
    if isinstance(src, dict):
        doc = src
    elif hasattr(src, "read"):
        doc = json.load(src)
    elif isinstance(src, str) and src.lstrip().startswith("{"):
        doc = json.loads(src)
    else:
        with open(src) as fh:
            doc = json.load(fh)
    inner = doc["graph"] if "graph" in doc else doc
    if "vertices" not in inner:
        raise ValueError("load_serde_json: no `vertices` array (not a net_ensembles graph dump)")
    verts = inner["vertices"]
    n = len(verts)
    if "next_id" in inner and int(inner["next_id"]) != n:
        raise ValueError("load_serde_json: next_id %s does not match %d vertices" % (inner["next_id"], n))
    adjacency, roots, is_sw = [], [], False
    for i, v in enumerate(verts):
        if int(v.get("id", i)) != i:
            raise ValueError("load_serde_json: vertex %d carries id %s" % (i, v.get("id")))
        row = []
        for e in v["adj"]:
            if isinstance(e, dict):
                is_sw = True
                row.append(int(e["to"]))
                roots.append(-1 if e.get("originally_to") is None else int(e["originally_to"]))
            else:
                row.append(int(e))
                roots.append(-1)
        adjacency.append(row)
    g = graph_from_adjacency(adjacency)
    if "edge_count" in inner and int(inner["edge_count"]) != g.edge_count():
        raise ValueError("load_serde_json: edge_count %s does not match the %d edges of the lists" % (inner["edge_count"], g.edge_count()))
    kind = "graph"
    if "graph" in doc:
        kind = "erc" if "c_target" in doc else "erm" if "m" in doc else "sw" if ("r_prob" in doc or is_sw) else "graph"
    params = {k: doc[k] for k in ("prob", "c_target", "m", "r_prob") if k in doc}
    rng = None
    if isinstance(doc.get("rng"), dict) and "state" in doc["rng"]:
        rng = (int(doc["rng"]["state"]), int(doc["rng"]["increment"]))
    erm_lists = None
    if kind == "erm" and all(k in doc for k in ("all_edges", "possible_edges", "current_edges")):
        erm_lists = tuple(np.asarray(doc[k], dtype=np.uint32).reshape(-1, 2) for k in ("all_edges", "possible_edges", "current_edges"))
    return SavedState(kind, g, params, rng, np.asarray(roots, dtype=np.int64) if is_sw else None, erm_lists)


def dump_serde_json(ensemble, path=None):
    """The inverse for this package's ensembles (ErEnsembleC / ErEnsembleM / SwEnsemble / Graph): a dict in the crate's
    layout - topology in storage order, parameters, generator state, ErEnsembleM's three edge lists - i.e. everything the
    crate's Deserialize needs to resume the ensemble."""
    import json
    adj = ensemble.adjacency()
    roots = ensemble.root_targets() if isinstance(ensemble, SwEnsemble) else None
    verts, at = [], 0
    for i, row in enumerate(adj):
        if roots is None:
            entries = [int(x) for x in row]
        else:
            entries = [{"to": int(x), "originally_to": (None if int(roots[at + k]) < 0 or int(roots[at + k]) == 0xFFFFFFFF else int(roots[at + k]))} for k, x in enumerate(row)]
        at += len(row)
        verts.append({"id": i, "adj": entries, "node": {}})
    inner = {"next_id": len(adj), "edge_count": ensemble.edge_count(), "vertices": verts, "phantom": None}
    if isinstance(ensemble, _Ensemble):
        doc = {"graph": inner}
        doc.update(getattr(ensemble, "_params", {}))
        state, inc = ensemble.rng_state()
        doc["rng"] = {"state": state, "increment": inc}
        if isinstance(ensemble, ErEnsembleM):
            for which, key in enumerate(("all_edges", "possible_edges", "current_edges")):
                pairs = np.empty((int(ensemble._lib.nech_erm_list_size(ensemble._h, which)), 2), np.uint32)
                if ensemble._lib.nech_erm_list(ensemble._h, which, _ptr(pairs)) != 0:
                    raise ValueError(ensemble._lib.nech_last_error().decode())
                doc[key] = pairs.tolist()
    else:
        doc = inner
    if path is not None:
        with open(path, "w") as fh:
            json.dump(doc, fh)
    return doc
