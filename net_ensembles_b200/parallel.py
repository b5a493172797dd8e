"""Multi-GPU vertex_load: one process per GPU, every rank holds the full CSR, the sources are
partitioned, the per-vertex f64 partial loads are combined with ONE all-reduce (NCCL over
NVLink on GPUs; gloo in the CPU tests).  Each source's contribution to b[] is independent
(/root/reference src/generic_graph/generic_graph.rs:1321 — the loop body only accumulates
into b), so no other exchange exists on this path.  torch is plumbing here (device buffer,
stream, process group); all load arithmetic is in the CUDA library."""
import numpy as np

BLOCK = 32  # sources travel in blocks of one engine batch so that a batch keeps neighbouring ids


def shard_sources(n_sources, rank, world, block=BLOCK):
    """Block-cyclic partition of 0..n_sources-1: block b (32 consecutive ids) goes to rank b % world.
    Consecutive ids stay together (ring-like graphs have neighbouring BFS trees for neighbouring
    ids, which keeps a batch's frontier coherent) and per-rank work differs by at most one block."""
    ids = np.arange(n_sources, dtype=np.uint32)
    return ids[(ids // block) % world == rank]


def vertex_load_sharded(device_graph, include_endpoints, group=None, out_tensor=None):
    """All ranks call this with their replica of the graph; returns a torch f64 CUDA tensor with
    the full load vector on every rank."""
    import torch
    import torch.distributed as dist
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)
    n = device_graph.n
    if out_tensor is None:
        out_tensor = torch.empty(n, dtype=torch.float64, device="cuda")
    device_graph.ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    device_graph.vertex_load_sources_device(shard_sources(n, rank, world), include_endpoints, out_tensor.data_ptr())
    if world > 1:
        dist.all_reduce(out_tensor, op=dist.ReduceOp.SUM, group=group)
    return out_tensor


def all_reduce_partials(partial, group=None):
    """Host-side combine used by the CPU (gloo) tests of the sharding logic: partial is a numpy
    f64 vector holding this rank's partial load."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(partial, dtype=np.float64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.numpy()
