#!/usr/bin/env python3
"""Regenerates tests/golden/fixtures.json.  Run in the BUILD container only: it reads the
reference checkout at /root/reference (absent on the GPU box; tests only read the JSON).

Contents
  known_answers : the reference's own known-answer vectors for vertex_load, transcribed from
                  /root/reference/tests/integration_test_graph.rs:238-275 (K4, 4 isolated
                  vertices, the 8-vertex graph with exact f64 expectations).
  graphs        : the 8 seeded ensemble fixtures /root/reference/TestData/unchaning_*.json,
                  reduced to what the hot path needs: constructor parameters + seed
                  (tests/integration_test_{er_c,er_m,sw}.rs), adjacency lists in storage
                  order, small-world root targets, final Pcg64 state/increment.
                  These are INPUTS (the reference stores no load values for them).
  derived_load  : vertex_load(true/false) of those 8 graphs computed by the pinned oracle
                  (oracle/ref_vertex_load.c), as f64 hex strings, plus integer checksums.
                  DERIVED here, not taken from the reference; cross-checked against networkx
                  load_centrality in tests/test_oracle.py.
"""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/TestData"
SEEDS = {  # tests/integration_test_er_c.rs:138-159, _er_m.rs:75-97, _sw.rs:142-189
    "erc_1": dict(kind="erc", n=123, c=2.0, seed=123929),
    "erc_2": dict(kind="erc", n=233, c=10.0, seed=1929),
    "erm_1": dict(kind="erm", n=123, m=921, seed=123929),
    "erm_2": dict(kind="erm", n=133, m=3000, seed=1929),
    "sw_1": dict(kind="sw", n=123, r_prob=0.1, seed=123929),
    "sw_2": dict(kind="sw", n=133, r_prob=0.3, seed=1929),
    "sw_3": dict(kind="sw", n=123, r_prob=0.0, seed=1229),
    "sw_4": dict(kind="sw", n=133, r_prob=1.0, seed=1922239),
}


def oracle_load(adj, incl):
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "build", "liboracle.so"))
    n = len(adj)
    rp = np.zeros(n + 1, dtype=np.uint64)
    rp[1:] = np.cumsum([len(a) for a in adj])
    col = np.array([x for a in adj for x in a], dtype=np.uint32)
    out = np.zeros(n)
    rc = lib.ref_vertex_load(ctypes.c_uint32(n), rp.ctypes.data_as(ctypes.c_void_p),
                             col.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(incl),
                             out.ctypes.data_as(ctypes.c_void_p), None)
    assert rc == 0
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit("reference checkout not present; fixtures.json is committed, nothing to do")
    edges8 = [(0, 1), (0, 2), (2, 1), (5, 1), (2, 3), (2, 4), (4, 3), (5, 3), (4, 5), (4, 7),
              (4, 6), (5, 7), (5, 6), (6, 7)]
    doc = {
        "known_answers": [
            {"name": "K4", "n": 4, "edges_in_insertion_order": [[i, j] for i in range(4) for j in range(i + 1, 4)],
             "include_endpoints_true": [3.0] * 4, "include_endpoints_false": [0.0] * 4,
             "source": "tests/integration_test_graph.rs:240-250"},
            {"name": "isolated4", "n": 4, "edges_in_insertion_order": [],
             "include_endpoints_true": [0.0] * 4,
             "source": "tests/integration_test_graph.rs:252-254"},
            {"name": "paper8", "n": 8, "edges_in_insertion_order": [list(e) for e in edges8],
             "include_endpoints_false": [0.0, 4.666666666666666, 8.0, 0.6666666666666665,
                                         8.666666666666668, 10.0, 0.0, 0.0],
             "source": "tests/integration_test_graph.rs:257-274"},
        ],
        "graphs": {},
        "derived_load": {},
    }
    for name, meta in SEEDS.items():
        d = json.load(open(os.path.join(REF, "unchaning_%s.json" % name)))
        verts = d["graph"]["vertices"]
        g = dict(meta)
        if meta["kind"] == "sw":
            g["adj"] = [[e["to"] for e in v["adj"]] for v in verts]
            g["root_target"] = [[-1 if e["originally_to"] is None else e["originally_to"] for e in v["adj"]] for v in verts]
        else:
            g["adj"] = [v["adj"] for v in verts]
        g["edge_count"] = d["graph"]["edge_count"]
        g["rng_state"] = str(d["rng"]["state"])
        g["rng_increment"] = str(d["rng"]["increment"])
        doc["graphs"][name] = g
        t, f = oracle_load(g["adj"], 1), oracle_load(g["adj"], 0)
        doc["derived_load"][name] = {
            "true_hex": [float(x).hex() for x in t], "false_hex": [float(x).hex() for x in f],
            "sum_true": float(t.sum()), "sum_false": float(f.sum()),
            "argmax_true": int(t.argmax()), "max_true": float(t.max()),
        }
    out = os.path.join(ROOT, "tests", "golden", "fixtures.json")
    with open(out, "w") as fh:
        json.dump(doc, fh, separators=(",", ":"))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
