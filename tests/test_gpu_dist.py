"""NCCL on hardware: one [channels, frames] buffer on rank 0 is scattered over the ranks (ncclSend/ncclRecv of row slabs),
every rank resamples its channels on its own GPU, the outputs are gathered back and checked against the oracle.
Needs >= 2 GPUs (skipped on the one-GPU box; run with `gpurun --gpus 2`)."""
import os
import socket
import sys

import numpy as np
import pytest

import oracle_util as ou

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_ch, lens, src, dst, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from __graft_entry__ import load_package
    pkg = load_package()
    par = __import__("r8brain_free_src_b200.parallel", fromlist=["x"])
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    start, count = par.shard_channels(n_ch, world, rank)
    plan = pkg.Plan(src, dst, max(lens), 2.0, pkg.ATTEN_24)
    batch = pkg.Batch(plan, max(count, 1), rank)
    x = ou.white_noise(n_ch, int(sum(lens)), 5) if rank == 0 else None
    outs, pos = [], 0
    for l in lens:
        full = torch.from_numpy(np.ascontiguousarray(x[:, pos:pos + l])).to(dev) if rank == 0 else None
        mine = par.scatter_channels(full, n_ch, l, dist, device=dev, dtype=torch.float64)
        y = batch.process(mine if count > 0 else torch.zeros((1, l), dtype=torch.float64, device=dev))
        back = par.gather_channels(y[:count].contiguous(), n_ch, dist)
        if rank == 0:
            outs.append(back.cpu().numpy())
        pos += l
    ok, worst = True, 0.0
    if rank == 0:
        ref = ou.best_oracle()
        for c in range(n_ch):
            r = ref.Resampler(src, dst, max(lens), 2.0, pkg.ATTEN_24)
            pos = 0
            for i, l in enumerate(lens):
                yr = r.process(x[c, pos:pos + l])
                pos += l
                ok = ok and len(yr) == outs[i].shape[1]
                if ok and len(yr):
                    m, rr = ou.parity_metrics(outs[i][c], yr)
                    worst = max(worst, m)
                    ok = ok and m <= 32 * ou.EPS and rr <= 4 * ou.EPS
    q.put((rank, bool(ok), worst / ou.EPS))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_ch", [5, 2])
def test_nccl_scatter_process_gather(pkg, n_ch):
    if pkg.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_ch, [8192, 8192, 1000], 44100.0, 96000.0, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
