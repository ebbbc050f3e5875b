"""world_size-2 gloo test of the multi-GPU plumbing (channel sharding, scatter/gather, max-reduce)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_ch, frames, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_package
    pkg = load_package()
    par = __import__("r8brain_free_src_b200.parallel", fromlist=["x"])
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(n_ch * frames, dtype=torch.float64).reshape(n_ch, frames) if rank == 0 else None
    mine = par.scatter_channels(full, n_ch, frames, dist, device=torch.device("cpu"), dtype=torch.float64)
    start, count = par.shard_channels(n_ch, world, rank)
    ok = mine.shape == (count, frames) and (count == 0 or float(mine[0, 0]) == start * frames)
    # stand-in for the per-shard GPU work: every rank's output count comes from the SAME host scheduler
    plan = pkg.Plan(44100.0, 96000.0, frames)
    n_out = plan.simulate([frames])[0]
    y = mine[:, :1].repeat(1, max(n_out, 1)) + 1.0
    back = par.gather_channels(y, n_ch, dist)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # bench.py's max-over-ranks timing reduction
    if rank == 0:
        exp = torch.arange(n_ch, dtype=torch.float64) * frames + 1.0
        ok = ok and back.shape == (n_ch, max(n_out, 1)) and bool(torch.all(back[:, 0] == exp))
    q.put((rank, bool(ok), float(t.item()), n_out))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_ch", [7, 2, 1])
def test_gloo_world2_scatter_gather(n_ch):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_ch, 4096, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    assert all(r[2] == 2.0 for r in res)
    assert len({r[3] for r in res}) == 1


def test_shard_partition_properties():
    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_package
    load_package()
    par = __import__("r8brain_free_src_b200.parallel", fromlist=["x"])
    for n in (0, 1, 5, 8, 1023, 1024, 8192):
        for w in (1, 2, 3, 4, 8):
            parts = [par.shard_channels(n, w, r) for r in range(w)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
