"""N>1 host logic on CPU: world_size-2 gloo.  The per-rank partial loads come from the oracle
here (no GPU in this container); what is under test is the partition + single all-reduce."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from net_ensembles_b200 import _build
    from net_ensembles_b200.parallel import all_reduce_partials, shard_sources
    from tests.helpers import Oracle
    import net_ensembles_b200 as ne
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    e = ne.SwEnsemble(300, 0.1, 77)
    rp, col = e.export_csr()
    orc = Oracle(_build.ORACLE_PATH)
    mine = shard_sources(300, rank, world)
    full = all_reduce_partials(orc.vertex_load(rp, col, True, sources=mine))
    want = orc.vertex_load(rp, col, True)
    ok = bool(np.all(np.abs(full - want) <= 1e-9 * np.maximum(np.abs(want), 1.0)))
    ret[rank] = (ok, len(mine))
    dist.destroy_process_group()


def test_shard_sources_partition():
    from net_ensembles_b200.parallel import shard_sources
    for n, world in ((0, 2), (1, 2), (31, 4), (1000, 3), (65536, 8)):
        parts = [shard_sources(n, r, world) for r in range(world)]
        allv = np.sort(np.concatenate(parts))
        assert allv.tolist() == list(range(n))
        sizes = [len(p) for p in parts]
        assert max(sizes) - min(sizes) <= 32
        for p in parts:                      # whole 32-blocks stay on one rank
            if len(p):
                assert np.all((p // 32) % world == (p[0] // 32) % world)


def test_two_rank_gloo_allreduce(built):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0][0] and ret[1][0]
    assert ret[0][1] + ret[1][1] == 300
