"""net_ensembles_b200 — B200-native `vertex_load` for net_ensembles graphs and ensembles.

The product is the CUDA library behind the C ABI in include/net_ensembles_cuda.h; this
package is the Python host side: a ctypes binding (`_lib`) and a mirror of the reference's
operator interface for this path (`Graph`, `SwGraph`-backed `SwEnsemble`, `ErEnsembleC`,
`ErEnsembleM`, `BAensemble`, each with `vertex_load(include_endpoints)`).
"""
from .api import (BAensemble, ChainBatch, Context, DeviceGraph, ErEnsembleC, ErEnsembleM, Graph, NecError,  # noqa: F401
                  SavedState, SwEnsemble, chains_step_and_load, default_context, dump_serde_json, gnm_scalable, ba_scalable,
                  graph_from_adjacency, load_serde_json, vertex_load_batch)

__all__ = ["Context", "DeviceGraph", "Graph", "ErEnsembleC", "ErEnsembleM", "SwEnsemble", "BAensemble",
           "NecError", "ChainBatch", "chains_step_and_load", "default_context", "gnm_scalable", "ba_scalable", "vertex_load_batch",
           "SavedState", "load_serde_json", "dump_serde_json", "graph_from_adjacency"]
