"""Test-side helpers: the oracle wrapper (oracle/ is test infrastructure — only tests/,
smoke() and bench.py's CPU legs touch it) and small graph builders."""
import ctypes

import numpy as np

c_u32, c_int, c_vp = ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p


def _p(a):
    return a.ctypes.data_as(c_vp) if a is not None else None


class Oracle:
    def __init__(self, path):
        self.lib = ctypes.CDLL(path)
        self.lib.ref_omp_max_threads.restype = c_int

    def vertex_load(self, row_ptr, col, include_endpoints, sources=None, threads=None):
        row_ptr = np.ascontiguousarray(row_ptr, np.uint64)
        col = np.ascontiguousarray(col, np.uint32)
        n = len(row_ptr) - 1
        out = np.zeros(n, np.float64)
        stats = np.zeros(2, np.uint64)
        if sources is None:
            sources = np.arange(n, dtype=np.uint32)
        sources = np.ascontiguousarray(sources, np.uint32)
        if threads is None:
            rc = self.lib.ref_vertex_load_sources(c_u32(n), _p(row_ptr), _p(col), _p(sources), c_u32(len(sources)),
                                                  c_int(int(include_endpoints)), _p(out), _p(stats))
            assert rc == 0
        else:
            rc = self.lib.ref_vertex_load_sources_mt(c_u32(n), _p(row_ptr), _p(col), _p(sources), c_u32(len(sources)),
                                                     c_int(int(include_endpoints)), c_int(threads), _p(out), _p(stats))
            assert rc > 0
        self.last_stats = {"edge_pairs": int(stats[0]), "reached_pairs": int(stats[1])}
        return out

    def bfs_debug(self, row_ptr, col, source):
        row_ptr = np.ascontiguousarray(row_ptr, np.uint64)
        col = np.ascontiguousarray(col, np.uint32)
        n = len(row_ptr) - 1
        dist, npred = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        sigma, bk = np.zeros(n, np.float64), np.zeros(n, np.float64)
        rc = self.lib.ref_bfs_debug(c_u32(n), _p(row_ptr), _p(col), c_u32(source), _p(dist), _p(npred), _p(sigma), _p(bk))
        assert rc == 0
        return dist, npred, sigma, bk

    def distance_measures(self, row_ptr, col, sources=None):
        row_ptr = np.ascontiguousarray(row_ptr, np.uint64)
        col = np.ascontiguousarray(col, np.uint32)
        n = len(row_ptr) - 1
        src = np.arange(n, dtype=np.uint32) if sources is None else np.ascontiguousarray(sources, np.uint32)
        ecc, sumd, reached = np.zeros(len(src), np.uint32), np.zeros(len(src), np.uint64), np.zeros(len(src), np.uint32)
        rc = self.lib.ref_distance_measures(c_u32(n), _p(row_ptr), _p(col), _p(src), c_u32(len(src)), _p(ecc), _p(sumd), _p(reached))
        assert rc == 0
        return ecc, sumd, reached

    def diameter(self, row_ptr, col):
        row_ptr = np.ascontiguousarray(row_ptr, np.uint64)
        col = np.ascontiguousarray(col, np.uint32)
        self.lib.ref_diameter.restype = ctypes.c_int64
        d = int(self.lib.ref_diameter(c_u32(len(row_ptr) - 1), _p(row_ptr), _p(col)))
        assert d >= -1
        return None if d < 0 else d

    def closeness_centrality(self, row_ptr, col):
        row_ptr = np.ascontiguousarray(row_ptr, np.uint64)
        col = np.ascontiguousarray(col, np.uint32)
        out = np.zeros(len(row_ptr) - 1, np.float64)
        assert self.lib.ref_closeness_centrality(c_u32(len(row_ptr) - 1), _p(row_ptr), _p(col), _p(out)) == 0
        return out

    def max_threads(self):
        return int(self.lib.ref_omp_max_threads())


def csr_from_adj(adj):
    n = len(adj)
    rp = np.zeros(n + 1, np.uint64)
    if n:
        rp[1:] = np.cumsum([len(a) for a in adj])
    col = np.array([x for a in adj for x in a], dtype=np.uint32)
    return rp, col


def adj_from_edges(n, edges):
    """Adjacency lists as Graph::add_edge builds them (push b onto a's list, a onto b's)."""
    adj = [[] for _ in range(n)]
    for a, b in edges:
        adj[a].append(b)
        adj[b].append(a)
    return adj


def close(a, b, rel=1e-9, abs_floor=1e-9):
    """The parity bar of BASELINE.json's north_star: relative 1e-9 (absolute floor for values < 1:
    leaves have load_false == 0 exactly)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return bool(np.all(np.abs(a - b) <= np.maximum(rel * np.maximum(np.abs(a), np.abs(b)), abs_floor)))


def edge_case_graphs():
    """(name, adjacency) — the shapes the reference tests plus the ones it never covers."""
    out = {}
    out["single"] = [[]]
    out["two_isolated"] = [[], []]
    out["one_edge"] = adj_from_edges(2, [(0, 1)])
    out["edgeless5"] = [[] for _ in range(5)]
    out["K5"] = adj_from_edges(5, [(i, j) for i in range(5) for j in range(i + 1, 5)])
    out["path7"] = adj_from_edges(7, [(i, i + 1) for i in range(6)])
    out["ring9"] = adj_from_edges(9, [(i, (i + 1) % 9) for i in range(9)])
    out["star8"] = adj_from_edges(8, [(0, i) for i in range(1, 8)])
    out["two_components"] = adj_from_edges(9, [(0, 1), (1, 2), (2, 0), (2, 3), (4, 5), (5, 6), (6, 7), (7, 4), (4, 6)])
    # 7-vertex graph on which load != Brandes betweenness
    out["load_vs_brandes7"] = adj_from_edges(7, [(0, 1), (0, 2), (0, 3), (1, 4), (2, 4), (3, 5), (4, 6), (5, 6)])
    out["grid4x5"] = adj_from_edges(20, [(r * 5 + c, r * 5 + c + 1) for r in range(4) for c in range(4)] +
                                    [(r * 5 + c, (r + 1) * 5 + c) for r in range(3) for c in range(5)])
    # hub with degree > 32 and > 64 (exercises the 32-neighbour chunking)
    out["star70"] = adj_from_edges(71, [(0, i) for i in range(1, 71)])
    out["wheel40"] = adj_from_edges(41, [(0, i) for i in range(1, 41)] + [(i, i % 40 + 1) for i in range(1, 41)])
    return out
