"""GPU: caller-side sample formats through r8bgpu_batch_process_host_fmt /
r8bgpu_batch_process_fmt.

The conversions restate what the reference's callers do on the CPU -- oneshot<Tin,Tout>() casts
"(double) ip[i]" / "(Tout) op[i]" (CDSPResampler.h:592-651).  Two checks per case:
  1. bit-exact: the typed path equals "our fp64 planar path + the same C conversion done in numpy"
     (integer work -> exact);
  2. against the compiled reference fed the widened input: at most one unit of the output format
     (a 5-eps fp64 difference can only flip a value that sits on a rounding/truncation boundary).
"""
import numpy as np
import pytest

import oracle_util as ou

pytestmark = pytest.mark.gpu


def pcm16(n_ch, n, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(-20000, 20000, size=(n_ch, n), dtype=np.int16)


def c_cast(y, dtype):
    """C++ narrowing conversion of a double array (float: nearest; ints: toward zero, saturating)."""
    if dtype == np.float32:
        return y.astype(np.float32)
    info = np.iinfo(dtype)
    return np.clip(np.trunc(y), info.min, info.max).astype(dtype)


def ref_run(ref, src, dst, max_in, x64, lens):
    out = []
    for c in range(x64.shape[0]):
        rs = ref.Resampler(src, dst, max_in, 2.0, 180.15)
        pos, acc = 0, []
        for l in lens:
            acc.append(rs.process(x64[c, pos:pos + l]))
            pos += l
        out.append(np.concatenate(acc))
    return np.stack(out)


@pytest.mark.parametrize("interleaved", [False, True])
@pytest.mark.parametrize("out_dtype", [np.int16, np.float32, np.int32])
def test_int16_in(pkg, ref, interleaved, out_dtype):
    src, dst, n_ch, lens = 44100.0, 96000.0, 37, [4096, 1000, 4096, 33]
    x = pcm16(n_ch, sum(lens), 5)
    a = pkg.ResamplerBatch(n_ch, src, dst, 4096, device=0)
    b = pkg.ResamplerBatch(n_ch, src, dst, 4096, device=0)
    pos, got, want64 = 0, [], []
    for l in lens:
        blk = x[:, pos:pos + l]
        y = a.batch.process_host_fmt(blk.T if interleaved else blk, out_dtype=out_dtype, interleaved=interleaved)
        got.append(y.T if interleaved else y)
        want64.append(b.process(blk.astype(np.float64)))
        pos += l
    got, want64 = np.concatenate(got, axis=1), np.concatenate(want64, axis=1)
    assert got.dtype == out_dtype and got.shape == want64.shape
    assert np.array_equal(got, c_cast(want64, out_dtype))          # 1. exact vs our own fp64 path
    yr = ref_run(ref, src, dst, 4096, x.astype(np.float64), lens)
    assert yr.shape == got.shape
    want = c_cast(yr, out_dtype)
    if out_dtype == np.float32:
        assert np.max(np.abs(got.astype(np.float64) - want.astype(np.float64)) / np.maximum(np.abs(yr), 1.0)) <= 2.0 ** -23
    else:
        assert np.max(np.abs(got.astype(np.int64) - want.astype(np.int64))) <= 1
        assert np.mean(got != want) < 1e-6                         # 2. boundary flips are vanishingly rare


def pack24(v):
    v = v.astype(np.int32)
    return np.stack([v & 0xff, (v >> 8) & 0xff, (v >> 16) & 0xff], axis=-1).astype(np.uint8)


def unpack24(b):
    b = b.astype(np.int32)
    v = b[..., 0] | (b[..., 1] << 8) | (b[..., 2] << 16)
    return np.where(v >= 1 << 23, v - (1 << 24), v).astype(np.int32)


@pytest.mark.parametrize("interleaved", [False, True])
def test_packed24_scaled(pkg, interleaved):
    # 24-bit PCM normalised to +-1 on the way in and back to 24-bit on the way out (power-of-two scales are exact)
    src, dst, n_ch, l = 48000.0, 44100.0, 5, 3000
    rng = np.random.default_rng(9)
    v = rng.integers(-(1 << 22), 1 << 22, size=(n_ch, l), dtype=np.int32)
    a = pkg.ResamplerBatch(n_ch, src, dst, l, device=0)
    b = pkg.ResamplerBatch(n_ch, src, dst, l, device=0)
    raw = pack24(v.T if interleaved else v)
    y = a.batch.process_host_fmt(raw, interleaved=interleaved, fmt=pkg.S24, out_fmt=pkg.S24,
                                 in_scale=2.0 ** -23, out_scale=2.0 ** 23)
    got = unpack24(y)
    got = got.T if interleaved else got
    want = c_cast(b.process(v.astype(np.float64) * 2.0 ** -23) * 2.0 ** 23, np.int32)
    assert got.shape == want.shape and got.shape[1] > 0
    assert np.array_equal(got, np.clip(want, -(1 << 23), (1 << 23) - 1))


def test_saturation_and_float_in(pkg):
    src, dst, n_ch, l = 44100.0, 96000.0, 3, 2048
    x = ou.white_noise(n_ch, l, 3).astype(np.float32)
    a = pkg.ResamplerBatch(n_ch, src, dst, l, device=0)
    b = pkg.ResamplerBatch(n_ch, src, dst, l, device=0)
    y = a.batch.process_host_fmt(x, out_dtype=np.int16, out_scale=1.0e6)   # far beyond int16
    w = b.process(x.astype(np.float64)) * 1.0e6
    assert np.array_equal(y, c_cast(w, np.int16))
    assert y.max() == 32767 and y.min() == -32768


def test_device_buffers(pkg):
    import torch
    src, dst, n_ch, l = 44100.0, 96000.0, 40, 4096
    x = ou.white_noise(n_ch, l, 11).astype(np.float32)
    a = pkg.ResamplerBatch(n_ch, src, dst, l, device=0)
    b = pkg.ResamplerBatch(n_ch, src, dst, l, device=0)
    cap = a.plan.max_out_len
    d_in = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()              # interleaved float32 [l][n_ch]
    d_out = torch.zeros((n_ch, cap), dtype=torch.float32, device="cuda")   # planar float32
    a.batch.set_stream(torch.cuda.current_stream().cuda_stream)
    n = a.batch.process_fmt(pkg.Buffer.make(d_in.data_ptr(), pkg.F32, True, n_ch),
                            l, pkg.Buffer.make(d_out.data_ptr(), pkg.F32, False, cap), cap, host=False)
    torch.cuda.synchronize()
    want = b.process(x.astype(np.float64)).astype(np.float32)
    assert n == want.shape[1]
    assert np.array_equal(d_out[:, :n].cpu().numpy(), want)


def test_passthrough_and_errors(pkg):
    a = pkg.ResamplerBatch(2, 48000.0, 48000.0, 256, device=0)
    x = pcm16(2, 256, 1)
    y = a.batch.process_host_fmt(x, out_dtype=np.float32)
    assert np.array_equal(y, x.astype(np.float32))
    with pytest.raises(pkg.R8bGpuError):
        a.batch.process_fmt(pkg.Buffer.make(x.ctypes.data, 99, False, 256), 256,
                            pkg.Buffer.make(x.ctypes.data, pkg.S16, False, 256), 256, host=True)
    with pytest.raises(pkg.R8bGpuError):   # interleaved stride below the channel count
        a.batch.process_fmt(pkg.Buffer.make(x.ctypes.data, pkg.S16, True, 1), 256,
                            pkg.Buffer.make(x.ctypes.data, pkg.S16, False, 256), 256, host=True)


def test_planar_formats_are_converted_inside_the_resampling_kernels(pkg):
    """Planar typed buffers need no conversion kernels on the chains whose first / last kernel is the fused one: the
    gather widens, the tensor-path stores narrow.  Launches per call: fused kernel + history copy (2), where the
    interleaved layout adds the two transposing conversion kernels (4).  Results equal the separate-kernel path
    bit for bit (R8BGPU_NO_FORMAT_FUSION) -- ragged calls included (history ring filled from typed blocks)."""
    import os
    src, dst, n_ch, lens = 44100.0, 96000.0, 5, [4096, 777, 4096, 1, 4096]
    rng = np.random.default_rng(3)
    x = rng.uniform(-0.9, 0.9, size=(n_ch, sum(lens))).astype(np.float32)
    outs = {}
    for mode in ("fused", "separate"):
        if mode == "separate":
            os.environ["R8BGPU_NO_FORMAT_FUSION"] = "1"
        try:
            rb = pkg.ResamplerBatch(n_ch, src, dst, 4096, device=0)
            pos, ys, per_call = 0, [], []
            for l in lens:
                l0 = rb.batch.kernel_launches
                ys.append(rb.batch.process_host_fmt(x[:, pos:pos + l], out_dtype=np.int32, out_scale=2.0 ** 30))
                per_call.append(rb.batch.kernel_launches - l0)
                pos += l
            outs[mode] = (np.concatenate(ys, axis=1), per_call)
        finally:
            os.environ.pop("R8BGPU_NO_FORMAT_FUSION", None)
    assert np.array_equal(outs["fused"][0], outs["separate"][0])
    assert max(outs["fused"][1]) == 2 and max(outs["separate"][1]) == 4, (outs["fused"][1], outs["separate"][1])


@pytest.mark.parametrize("src,dst", [(192000.0, 44100.0), (96000.0, 44100.0), (44100.0, 88200.0)])
def test_typed_io_on_other_fused_chains(pkg, ref, src, dst):
    # decimating chain (typed output from the 1x fused pair), 1x pair first (typed gather), lone 2x block convolver
    n_ch, lens = 3, [8192, 8192, 100, 8192]
    x = pcm16(n_ch, sum(lens), 8)
    a = pkg.ResamplerBatch(n_ch, src, dst, 8192, device=0)
    b = pkg.ResamplerBatch(n_ch, src, dst, 8192, device=0)
    pos = 0
    for l in lens:
        blk = x[:, pos:pos + l]
        got = a.batch.process_host_fmt(blk, out_dtype=np.float32)
        want = b.process(blk.astype(np.float64))
        assert np.array_equal(got, c_cast(want, np.float32))
        pos += l
