"""ctypes binding of the C ABI declared in include/net_ensembles_cuda.h (nec_*) and of the
host-side graph/ensemble mirror (nech_*).  This is the stub a Python host would use; the Rust
stub is rust/net_ensembles_cuda.rs.  Loading fails loudly if the CUDA library is missing —
there is no CPU fallback and no PyTorch/eager implementation of the path."""
import ctypes
import os

from . import _build

c_u32, c_u64, c_int, c_dbl, c_vp = ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int, ctypes.c_double, ctypes.c_void_p

NEC_OK, NEC_EINVAL, NEC_ECUDA, NEC_ENOMEM, NEC_ENODEVICE, NEC_EINTERNAL = 0, -1, -2, -3, -4, -5
NEC_BATCH_MAX_N = 512


class NecStats(ctypes.Structure):
    _fields_ = [("n_sources", c_u64), ("n_batches", c_u64), ("reached_pairs", c_u64), ("edge_pairs", c_u64),
                ("vertex_visits", c_u64), ("max_depth", c_u32), ("kernel_launches", c_u32),
                ("wide_batches", c_u32), ("n_slots", c_u32), ("kernel_ms", ctypes.c_float), ("engine_ms", ctypes.c_float), ("frac_reset", ctypes.c_float), ("frac_forward", ctypes.c_float), ("frac_backward", ctypes.c_float),
                ("workspace_bytes", c_u64), ("engine", c_u32), ("batch_width", c_u32),
                ("team_size", c_u32), ("cta_threads", c_u32), ("pull_levels", c_u32), ("push_levels", c_u32),
                ("requeued_batches", c_u32), ("reserved0", c_u32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/net_ensembles_cuda.h declares: name -> (restype, argtypes)
NEC_SYMBOLS = {
    "nec_version": (ctypes.c_char_p, []),
    "nec_init": (c_int, [ctypes.POINTER(c_vp), ctypes.POINTER(c_int), c_int]),
    "nec_destroy": (None, [c_vp]),
    "nec_last_error": (ctypes.c_char_p, [c_vp]),
    "nec_set_stream": (c_int, [c_vp, c_vp]),
    "nec_set_workspace_limit": (c_int, [c_vp, c_u64]),
    "nec_set_engine": (c_int, [c_vp, c_int]),
    "nec_graph_upload": (c_int, [c_vp, c_u32, c_u64, c_vp, c_vp, ctypes.POINTER(c_vp)]),
    "nec_graph_upload_adjacency": (c_int, [c_vp, c_u32, c_vp, c_vp, c_u32, c_u32, c_u32, ctypes.POINTER(c_vp)]),
    "nec_graph_free": (None, [c_vp, c_vp]),
    "nec_graph_dims": (c_int, [c_vp, ctypes.POINTER(c_u32), ctypes.POINTER(c_u64)]),
    "nec_graph_download": (c_int, [c_vp, c_vp, c_vp, c_vp]),
    "nec_erc_randomize": (c_int, [c_vp, c_u32, c_dbl, c_vp, c_vp, ctypes.POINTER(c_vp)]),
    "nec_vertex_load": (c_int, [c_vp, c_vp, c_int, c_vp]),
    "nec_vertex_load_sources": (c_int, [c_vp, c_vp, c_vp, c_u32, c_int, c_vp]),
    "nec_vertex_load_sources_device": (c_int, [c_vp, c_vp, c_vp, c_u32, c_int, c_vp]),
    "nec_vertex_load_batch": (c_int, [c_vp, c_u32, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp]),
    "nec_chains_create": (c_int, [c_vp, c_u32, c_vp, c_vp, c_vp, c_vp, c_vp, c_u32, ctypes.POINTER(c_vp)]),
    "nec_chains_step_and_load": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_vp]),
    "nec_chains_result": (c_vp, [c_vp]),
    "nec_chains_download": (c_int, [c_vp, c_vp, c_vp, c_vp, c_u64, c_vp]),
    "nec_chains_free": (None, [c_vp, c_vp]),
    "nec_distance_measures": (c_int, [c_vp, c_vp, c_vp, c_u32, c_vp, c_vp, c_vp]),
    "nec_bfs_debug": (c_int, [c_vp, c_vp, c_u32, c_vp, c_vp, c_vp, c_vp]),
    "nec_wave_size": (c_int, [c_vp, c_vp, c_u64, c_vp]),
    "nec_get_stats": (c_int, [c_vp, ctypes.POINTER(NecStats)]),
}
NECH_SYMBOLS = {
    "nech_last_error": (ctypes.c_char_p, []),
    "nech_graph_new": (c_vp, [c_u64]),
    "nech_graph_from_adjacency": (c_vp, [c_u64, c_vp, c_vp]),
    "nech_restore": (c_vp, [c_int, c_u64, c_vp, c_vp, c_vp, c_dbl, c_dbl, c_u64, c_vp, c_vp, c_u64, c_vp, c_u64, c_vp, c_u64]),
    "nech_erm_list_size": (c_u64, [c_vp, c_int]),
    "nech_erm_list": (c_int, [c_vp, c_int, c_vp]),
    "nech_erc_new": (c_vp, [c_u64, c_dbl, c_u64]),
    "nech_erm_new": (c_vp, [c_u64, c_u64, c_u64]),
    "nech_sw_new": (c_vp, [c_u64, c_dbl, c_u64, c_u64]),
    "nech_ba_new": (c_vp, [c_u64, c_u64, c_u64, c_u64]),
    "nech_gnm_scalable": (c_vp, [c_u64, c_u64, c_u64]),
    "nech_ba_scalable": (c_vp, [c_u64, c_u64, c_u64, c_u64]),
    "nech_free": (None, [c_vp]),
    "nech_vertex_count": (c_u64, [c_vp]),
    "nech_edge_count": (c_u64, [c_vp]),
    "nech_nnz": (c_u64, [c_vp]),
    "nech_add_edge": (c_int, [c_vp, c_u32, c_u32]),
    "nech_remove_edge": (c_int, [c_vp, c_u32, c_u32]),
    "nech_randomize": (c_int, [c_vp]),
    "nech_m_steps": (c_int, [c_vp, c_u64]),
    "nech_export_csr": (c_int, [c_vp, c_vp, c_vp]),
    "nech_sw_root_targets": (c_int, [c_vp, c_vp]),
    "nech_rng_state": (c_int, [c_vp, c_vp]),
    "nech_pcg64_draws": (None, [c_u64, c_u64, c_vp]),
}
# the drop-in calls over host-mirror handles: compiled into the CUDA library (csrc/host/host_bridge.cpp)
BRIDGE_SYMBOLS = {
    "nech_erc_randomize_cuda": (c_int, [c_vp, c_vp, ctypes.POINTER(c_vp)]),
    "nech_upload_cuda": (c_int, [c_vp, c_vp, ctypes.POINTER(c_vp)]),
    "nech_vertex_load_cuda": (c_int, [c_vp, c_vp, c_int, c_vp]),
    "nech_chains_step_and_load_cuda": (c_int, [c_vp, c_u32, c_u64, c_vp, c_int, c_vp]),
    "nech_last_chain_stats": (None, [ctypes.POINTER(NecStats)]),
    "nech_chain_batch_new": (c_vp, [c_vp, c_u32, c_vp]),
    "nech_chain_batch_free": (None, [c_vp]),
    "nech_chain_batch_step": (c_int, [c_vp, c_u64, c_int, c_vp]),
    "nech_chain_batch_result": (c_vp, [c_vp]),
    "nech_chain_batch_stats": (None, [c_vp, ctypes.POINTER(NecStats), ctypes.POINTER(c_u64), ctypes.POINTER(c_u64)]),
    "nech_chain_batch_download": (c_int, [c_vp, c_vp, c_vp, c_u64, ctypes.POINTER(c_u64)]),
}

_lib = None
_host = None


def load_host():
    """The graph / ensemble mirror (libnec_host.so): input producers only, no CUDA behind it."""
    global _host
    if _host is not None:
        return _host
    path = os.environ.get("NEC_HOST_LIB_PATH") or _build.HOST_LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`" % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in NECH_SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _host = lib
    return lib


def lib_path():
    # NEC_LIB_PATH: A/B experiments with two builds on the same GPU box (scripts/)
    return os.environ.get("NEC_LIB_PATH") or _build.LIB_PATH


def load():
    """Loads (never builds) the in-tree shared library and declares every symbol."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(
            "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(net_ensembles_b200 has no CPU fallback)" % path)
    lib = ctypes.CDLL(path)
    for table in (NEC_SYMBOLS, BRIDGE_SYMBOLS):
        for name, (res, args) in table.items():
            fn = getattr(lib, name)          # AttributeError if the export is missing
            fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib
