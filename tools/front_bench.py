#!/usr/bin/env python3
"""One process, every GPU of the box behind ONE C-ABI batch (r8bgpu_batch_create(plan, n, R8BGPU_DEVICE_ALL)):
end-to-end rate of r8bgpu_batch_process_host() on NUMA-placed pinned buffers, cfg 2 shape (1024 channels per GPU)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
import numpy as np  # noqa: E402

n_dev = pkg.device_count()
per, block = 1024, 65536
plan = pkg.Plan(44100.0, 96000.0, block, 2.0, pkg.ATTEN_24)
res = {}
for numa in ((True,) if os.environ.get("R8BGPU_NO_HUGEPAGES") else (True, False)):
    batch = pkg.Batch(plan, per * n_dev, pkg.DEVICE_ALL)
    cap = (plan.max_out_len + 7) // 8 * 8
    if numa:
        hx, hy = batch.host_alloc(block), batch.host_alloc(cap)
    else:
        import torch
        hx = torch.empty((per * n_dev, block), dtype=torch.float64).pin_memory().numpy()
        hy = torch.empty((per * n_dev, cap), dtype=torch.float64).pin_memory().numpy()
    rng = np.random.default_rng(1)
    for c0 in range(0, per * n_dev, 256):
        hx[c0:c0 + 256] = rng.uniform(-1.0, 1.0, size=(256, block))
    for _ in range(2):
        batch.process_host_ptr(hx.ctypes.data, block, block, hy.ctypes.data, cap, cap)
    k = 6
    t0 = time.perf_counter()
    for _ in range(k):
        n = batch.process_host_ptr(hx.ctypes.data, block, block, hy.ctypes.data, cap, cap)
    dt = time.perf_counter() - t0
    res["numa_buffers" if numa else "plain_pinned"] = {"Msamples_per_s": 1e-6 * per * n_dev * block * k / dt, "ms_per_call": dt / k * 1e3,
                                                       "d2h_gbs_per_gpu": per * n * 8 * k / dt / 1e9}
    shards = batch.shards()
    del batch
print(json.dumps({"n_gpus": n_dev, "channels": per * n_dev, "shards": shards, **res}))
