"""CPU: the v2 fused kernel's own arithmetic, run thread by thread on the host, against the oracle.

csrc/r8b_fused2_core.cuh holds the per-thread phase functions of k_up2_frac2 (real-input FFT split, slot orders,
spectrum multiply, inverse passes, interpolation bookkeeping); they compile for the host.  tests/cpp/fused2_emul.cpp
drives them with the engine's own host code (plan, schedule, tables, tile geometry).  What this cannot see is the
device-only part: barriers, bulk copies and the transposed store staging -- those are covered by the GPU parity tests.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_util

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "r8brain-free-src_b200", "csrc")
EPS = 2.0 ** -52


def _cuda_include():
    for d in (os.environ.get("CUDA_HOME"), "/usr/local/cuda"):
        if d and os.path.exists(os.path.join(d, "include", "cuda_runtime.h")):
            return os.path.join(d, "include")
    return None


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    inc = _cuda_include()
    if inc is None:
        pytest.skip("CUDA headers not found")
    so = str(tmp_path_factory.mktemp("f2emul") / "libf2emul.so")
    srcs = [os.path.join(HERE, "cpp", "fused2_emul.cpp")] + [os.path.join(CSRC, f) for f in
                                                             ("r8b_plan.cpp", "r8b_design.cpp", "r8b_hosttab.cpp")]
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-I" + inc, "-o", so] + srcs, check=True)
    L = C.CDLL(so)
    L.f2emul_create.restype = C.c_void_p
    L.f2emul_create.argtypes = [C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int]
    L.f2emul_process.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.f2emul_destroy.argtypes = [C.c_void_p]
    return L


def _run(L, src, dst, max_len, lens, glog=-1, atten=180.15, tb=2.0, seed=7):
    ref = oracle_util.best_oracle()
    h = L.f2emul_create(src, dst, max_len, tb, atten, glog)
    assert h, "rate pair does not plan to BlockConv(2x) -> whole-stepping interpolator"
    rs = ref.Resampler(src, dst, max_len, tb, atten)
    rng = np.random.default_rng(seed)
    worst = se = sy = 0.0
    total = 0
    for n in lens:
        x = rng.uniform(-1.0, 1.0, n)
        out = np.zeros(int(n * dst / src * 1.25) + 4096)
        k = L.f2emul_process(h, x.ctypes.data, n, out.ctypes.data, len(out))
        yr = rs.process(x)
        assert k == len(yr), (k, len(yr))
        if k:
            d = out[:k] - yr
            worst = max(worst, float(np.max(np.abs(d))) / float(np.max(np.abs(yr))))
            se += float(np.sum(d * d))
            sy += float(np.sum(yr * yr))
            total += k
    L.f2emul_destroy(h)
    assert total > 0
    assert worst <= 32 * EPS, worst / EPS
    assert (se / sy) ** 0.5 <= 4 * EPS, (se / sy) ** 0.5 / EPS


@pytest.mark.parametrize("glog", [0, 1, 2])
def test_cfg2_chain_all_lane_splits(emul, glog):
    _run(emul, 44100.0, 96000.0, 8192, [8192, 8192, 8192], glog=glog)


@pytest.mark.parametrize("src,dst,lens", [(44100.0, 96000.0, [8192, 8192, 3, 0, 4097, 8191]),
                                          (48000.0, 44100.0, [8191, 8191, 4000]),       # padded y layout, 19 groups
                                          (44100.0, 48000.0, [4096] * 3)])
def test_tensor_path_interpolation(emul, src, dst, lens):
    """glog = 8 selects the m8n8k4 formulation (fragment index functions of r8b_fused2_core.cuh)."""
    _run(emul, src, dst, max(lens), lens, glog=8)


@pytest.mark.parametrize("src,dst,lens,atten", [(96000.0, 44100.0, [8192, 8192, 5, 0, 4099, 8191], 180.15),
                                                (48000.0, 22050.0, [4096] * 4, 206.91),
                                                (100000.0, 44100.0, [8192, 8192], 136.45)])
def test_up1_pair_blockconv_to_interpolator(emul, src, dst, lens, atten):
    """BlockConvolver 1/1 -> whole-stepping interpolator (the tail of the decimating chains): 2048-point inverse
    transform mirroring the real-input forward one, tensor-path interpolation over a 4096-sample tile."""
    _run(emul, src, dst, max(lens), lens, glog=8, atten=atten)


def test_ragged_blocks_history_ring_and_misaligned_rows(emul):
    # odd lengths shift the block base parity (plain-load path), tiny and empty blocks reach into the history ring
    _run(emul, 44100.0, 96000.0, 8192, [1, 0, 4097, 777, 8192, 3, 8191, 5000])


def test_downsampling_chain_padded_y_layout(emul):
    _run(emul, 48000.0, 44100.0, 8191, [8191, 8191, 4000])   # in_step 320: padded y layout, 10-phase groups


def test_full_block_and_presets(emul):
    _run(emul, 44100.0, 96000.0, 65536, [65536, 65536])
    _run(emul, 44100.0, 48000.0, 4096, [4096] * 4, atten=136.45)


def test_seeded_sweep_of_rate_pairs_and_ragged_calls(emul):
    """Every rate pair here plans to BlockConvolver (2x or 1x) -> whole-stepping interpolator; presets, transition
    bands, call lengths and the interpolation formulation (register-tiled FMA / tensor-path fragments) are drawn."""
    rng = np.random.default_rng(4242)
    pairs = [(44100.0, 96000.0), (48000.0, 44100.0), (44100.0, 48000.0), (32000.0, 44100.0), (48000.0, 88200.0),
             (96000.0, 44100.0), (88200.0, 48000.0), (44100.0, 64000.0), (100000.0, 44100.0), (22050.0, 32000.0)]
    done = 0
    for src, dst in pairs:
        for _ in range(2):
            att = float(rng.choice([136.45, 180.15, 206.91]))
            tb = float(rng.choice([2.0, 3.0, 6.0]))
            max_len = int(rng.choice([2048, 4096, 8192]))
            lens = [max_len, max_len] + [int(v) for v in rng.integers(0, max_len + 1, 4)] + [max_len]
            glog = int(rng.choice([-1, 8]))
            h = emul.f2emul_create(src, dst, max_len, tb, att, glog)
            if not h:   # this preset does not take the fused chain for the pair (kernel too long for the tile)
                continue
            emul.f2emul_destroy(h)
            _run(emul, src, dst, max_len, lens, glog=glog, atten=att, tb=tb, seed=int(rng.integers(1 << 30)))
            done += 1
    assert done >= 12, done
