"""Multi-device batches behind the C-ABI (r8bgpu_batch_create(plan, n, R8BGPU_DEVICE_ALL)) against the oracle.

On a one-GPU box R8BGPU_FORCE_SHARDS deals several shards to the same device, which exercises the whole front (channel
ranges, per-shard worker threads, NUMA-placed host buffers); with >= 2 GPUs the shards land on different devices.
"""
import os

import numpy as np
import pytest

import oracle_util as ou

pytestmark = pytest.mark.gpu


def _run(pkg, ref, n_ch, src, dst, lens, numa_buffers):
    plan = pkg.Plan(src, dst, max(lens), 2.0, pkg.ATTEN_24)
    batch = pkg.Batch(plan, n_ch, pkg.DEVICE_ALL)
    shards = batch.shards()
    assert sum(s[2] for s in shards) == n_ch and [s[1] for s in shards] == list(np.cumsum([0] + [s[2] for s in shards[:-1]]))
    x = ou.white_noise(n_ch, int(sum(lens)), 77)
    rs = [ref.Resampler(src, dst, max(lens), 2.0, pkg.ATTEN_24) for _ in range(n_ch)]
    cap = plan.max_out_len
    if numa_buffers:
        hin, hout = batch.host_alloc(max(lens)), batch.host_alloc(cap)
    else:
        hin, hout = np.empty((n_ch, max(lens))), np.empty((n_ch, cap))
    pos = 0
    for l in lens:
        hin[:, :l] = x[:, pos:pos + l]
        n = batch.process_host_ptr(hin.ctypes.data, hin.shape[1], l, hout.ctypes.data, hout.shape[1], cap)
        for c in range(n_ch):
            yr = rs[c].process(x[c, pos:pos + l])
            assert len(yr) == n
            if n:
                m, r = ou.parity_metrics(hout[c, :n], yr)
                assert m <= 32 * ou.EPS and r <= 4 * ou.EPS, (c, m / ou.EPS, r / ou.EPS)
        pos += l
    # device-pointer calls are refused on a front, and say where to go instead
    if len(shards) > 1:
        with pytest.raises(pkg.R8bGpuError, match="shard"):
            batch.process_ptr(1, 1, 1, 1, 1, 1)
    if numa_buffers:
        pkg.host_free(hin)
        pkg.host_free(hout)
    return shards


@pytest.mark.parametrize("numa_buffers", [False, True])
def test_forced_shards_on_one_device(pkg, ref, monkeypatch, numa_buffers):
    monkeypatch.setenv("R8BGPU_FORCE_SHARDS", "3")
    shards = _run(pkg, ref, 7, 44100.0, 96000.0, [4096, 1000, 0, 4096, 17], numa_buffers)
    assert len(shards) == 3 and [s[2] for s in shards] == [3, 3, 1]


def test_all_visible_devices(pkg, ref):
    if pkg.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    shards = _run(pkg, ref, 2 * pkg.device_count() + 1, 48000.0, 44100.0, [8192, 8192, 3000], True)
    assert len({s[0] for s in shards}) == pkg.device_count()


def test_one_channel_is_an_ordinary_batch(pkg):
    plan = pkg.Plan(44100.0, 96000.0, 1024)
    b = pkg.Batch(plan, 1, pkg.DEVICE_ALL)
    assert len(b.shards()) == 1
