"""GPU parity: the CUDA path (through the C-ABI) against the compiled reference oracle.

Tolerance: per channel, max|d| <= 32*eps*max|y| and rms(d) <= 4*eps*rms(y),
eps = 2^-52, with per-call output counts EXACTLY equal.  The reference disagrees with itself at
the 10-eps level between FFT back-ends, so bit equality is not defined for this path.
"""
import numpy as np
import pytest

import oracle_util as ou

pytestmark = pytest.mark.gpu

MAX_TOL = 32 * ou.EPS
RMS_TOL = 4 * ou.EPS


def run_both(pkg, oracle, src, dst, lens, n_ch=3, tb=2.0, atten=180.15, extfft=0, seed=1, max_in=None):
    max_in = max_in or max(lens)
    x = ou.white_noise(n_ch, int(sum(lens)), seed)
    rb = pkg.ResamplerBatch(n_ch, src, dst, max_in, tb, atten, device=0, extfft=extfft)
    rs = [oracle.Resampler(src, dst, max_in, tb, atten) for _ in range(n_ch)]
    pos = 0
    ys, yr = [[] for _ in range(n_ch)], [[] for _ in range(n_ch)]
    for l in lens:
        y = rb.process(x[:, pos:pos + l])
        for c in range(n_ch):
            r = rs[c].process(x[c, pos:pos + l])
            assert len(r) == y.shape[1], "per-call count differs: ref %d gpu %d (l=%d)" % (len(r), y.shape[1], l)
            ys[c].append(y[c])
            yr[c].append(r)
        pos += l
    return [np.concatenate(a) for a in ys], [np.concatenate(a) for a in yr]


def check(ys, yr, max_tol=MAX_TOL, rms_tol=RMS_TOL):
    worst = (0.0, 0.0)
    for a, b in zip(ys, yr):
        assert len(a) == len(b) and len(a) > 0
        m, r = ou.parity_metrics(a, b)
        worst = (max(worst[0], m), max(worst[1], r))
        assert m <= max_tol, "max err %.3g eps" % (m / ou.EPS)
        assert r <= rms_tol, "rms err %.3g eps" % (r / ou.EPS)
    return worst


CHAINS = [
    (44100.0, 96000.0),    # BASELINE cfg 1/2: BlockConv 2x -> whole-step interp 147/160
    (48000.0, 44100.0),    # cfg 3: BlockConv 2x (NormFreq .459) -> whole-step 320/147
    (48000.0, 47999.0),    # cfg 5: BlockConv 2x -> order-2 interpolated bank
    (44100.0, 88200.0),    # BlockConv 2x only
    (44100.0, 176400.0),   # BlockConv 2x -> HBUp
    (192000.0, 44100.0),   # HBDown -> BlockConv 1/1 -> whole-step
    (96000.0, 48000.0),    # BlockConv 1/2
    (48000.0, 16000.0),    # BlockConv 1/3
    (44100.0, 192000.0),   # BlockConv -> interp -> BlockConv -> HBUp (intermediate interpolation)
    (44100.0, 22050.5),    # fractional downsampling just above 2x
    (44100.0, 132300.0),   # BlockConv 3/1 (time-domain zero-stuffing)
    (32000.0, 48000.0),    # BlockConv 3/2 (reference-exact power-of-two decimation on a 3x stream)
    (48000.0, 36000.0),    # BlockConv 3/4
    (48000.0, 32000.0),    # BlockConv 2/3
    (8000.0, 48000.0),     # BlockConv 3/1 -> HBUp (third-band taps)
]


@pytest.mark.parametrize("src,dst", CHAINS)
def test_chain_parity(pkg, ref, src, dst):
    ys, yr = run_both(pkg, ref, src, dst, [8192] * 4 + [1000, 1, 0, 17, 8192], max_in=8192)
    w = check(ys, yr)
    print("%g->%g: max %.2f eps, rms %.2f eps" % (src, dst, w[0] / ou.EPS, w[1] / ou.EPS))


def test_dsd_cascade_extfft(pkg, ref_e1):
    # cfg 4: 44100 -> 2822400 with R8B_EXTFFT=1: BlockConv 2x -> HBUp x5
    ys, yr = run_both(pkg, ref_e1, 44100.0, 2822400.0, [2048] * 5 + [100], n_ch=2, extfft=1, max_in=2048)
    check(ys, yr)


def test_hbdown_cascade(pkg, ref):
    # (EXTFFT variant below)
    ys, yr = run_both(pkg, ref, 2822400.0, 44100.0, [65536] * 8, n_ch=2, max_in=65536)
    check(ys, yr)


def test_decimating_chains_are_fused(pkg, ref):
    """The five half-band decimators of 2822400 -> 44100 run as ONE kernel (k_hbdown_cascade) followed by the block
    convolver: two launches per call (+ the history copy); 192000 -> 44100 runs its 1x BlockConvolver and interpolator
    in the fused kernel.  Ragged calls reach across call boundaries into the cascade's recomputed history."""
    plan = pkg.Plan(2822400.0, 44100.0, 65536, 2.0, pkg.ATTEN_24)
    batch = pkg.Batch(plan, 2, 0)
    names = batch.stage_kernels()
    assert names[0] == ("k_hbdown_cascade", 5) and all(n == ("(fused)", 0) for n in names[1:5]), names
    x = ou.white_noise(2, 65536, 3)
    batch.process_host(x)
    l0 = batch.kernel_launches
    batch.process_host(x)
    assert batch.kernel_launches - l0 == 3, batch.kernel_launches - l0  # cascade, block convolver, history copy
    b2 = pkg.Batch(pkg.Plan(192000.0, 44100.0, 8192, 2.0, pkg.ATTEN_24), 2, 0)
    assert [n[0] for n in b2.stage_kernels()] == ["k_hbdown", "k_up2_frac2", "(fused)"]
    ys, yr = run_both(pkg, ref, 2822400.0, 44100.0, [65536, 1000, 65536, 7, 0, 33333, 65536, 65536, 65536], n_ch=3, max_in=65536)
    check(ys, yr)
    ys, yr = run_both(pkg, ref, 705600.0, 44100.0, [30000] * 6, n_ch=2, max_in=30000)
    check(ys, yr)


def test_full_block_size(pkg, ref):
    # BASELINE block size: 65536-sample calls, counts 138963, 142664, 142663 ...
    ys, yr = run_both(pkg, ref, 44100.0, 96000.0, [65536] * 3, n_ch=2)
    assert len(ys[0]) == 138963 + 142664 + 142663
    check(ys, yr)


def test_presets_and_transition_bands(pkg, ref):
    for atten in (pkg.ATTEN_16, pkg.ATTEN_16IR, 206.91):
        for tb in (2.0, 5.0):
            ys, yr = run_both(pkg, ref, 44100.0, 48000.0, [4096] * 6, n_ch=1, tb=tb, atten=atten)
            check(ys, yr)


def test_chunking_invariance(pkg):
    # Whole-stepping chains: the same stream fed in different block sizes gives the same output to
    # rounding (the reference is bit-invariant; our FFT tiles are anchored per call).
    x = ou.white_noise(2, 40000, 7)
    outs = []
    for lens in ([40000], [4096] * 9 + [3136], [1000] * 40, [1] * 50 + [39950]):
        rb = pkg.ResamplerBatch(2, 44100.0, 96000.0, 40000, 2.0, pkg.ATTEN_24, device=0)
        pos, acc = 0, []
        for l in lens:
            acc.append(rb.process(x[:, pos:pos + l]))
            pos += l
        outs.append(np.concatenate(acc, axis=1))
    for o in outs[1:]:
        assert o.shape == outs[0].shape
        m, r = ou.parity_metrics(o[0], outs[0][0])
        assert m <= MAX_TOL and r <= RMS_TOL


def test_clear_restarts_stream(pkg, ref):
    x = ou.white_noise(1, 20000, 3)
    rb = pkg.ResamplerBatch(1, 44100.0, 96000.0, 20000, device=0)
    a = rb.process(x)
    rb.clear()
    b = rb.process(x)
    assert np.array_equal(a, b)


def test_zero_input_gives_zero_output(pkg):
    rb = pkg.ResamplerBatch(2, 48000.0, 47999.0, 8192, device=0)
    y = rb.process(np.zeros((2, 8192)))
    assert y.shape[1] > 0 and not np.any(y)


def test_impulse_and_dc_gain(pkg):
    # unity DC gain end to end (filter gain = upsampling factor, bank rows sum to 1)
    rb = pkg.ResamplerBatch(1, 44100.0, 96000.0, 16384, device=0)
    y = np.concatenate([rb.process(np.ones((1, 16384)))[0] for _ in range(2)])
    assert abs(y[-1000:].mean() - 1.0) < 1e-13


def test_device_pointer_api_and_strides(pkg, ref):
    import torch
    n_ch, l = 5, 6000
    x = ou.white_noise(n_ch, l * 2, 11)
    plan = pkg.Plan(44100.0, 96000.0, l, 2.0, pkg.ATTEN_24)
    b = pkg.Batch(plan, n_ch, 0)
    xin = torch.zeros((n_ch, 2 * l + 13), dtype=torch.float64, device="cuda:0")  # padded stride
    xin[:, :2 * l] = torch.from_numpy(x).cuda()
    out = torch.empty((n_ch, plan.max_out_len + 5), dtype=torch.float64, device="cuda:0")
    got = []
    for c in range(2):
        y = b.process(xin[:, c * l:(c + 1) * l], out)
        got.append(y.cpu().numpy().copy())
    got = np.concatenate(got, axis=1)
    for c in range(n_ch):
        r = ref.Resampler(44100.0, 96000.0, l, 2.0, pkg.ATTEN_24)
        yr = np.concatenate([r.process(x[c, :l]), r.process(x[c, l:])])
        assert len(yr) == got.shape[1]
        m, rr = ou.parity_metrics(got[c], yr)
        assert m <= MAX_TOL and rr <= RMS_TOL
    assert b.kernel_launches > 0


def test_errors_are_loud(pkg):
    plan = pkg.Plan(44100.0, 96000.0, 1024)
    b = pkg.Batch(plan, 1, 0)
    with pytest.raises(pkg.R8bGpuError):
        b.process_host(np.zeros((1, 2048)))  # l > MaxInLen
    with pytest.raises(pkg.R8bGpuError):
        pkg.Plan(44100.0, 96000.0, 1024, phase=1)  # minimum phase not implemented


# ---- committed golden vectors (generated from the reference by tests/golden/make_golden.py) ----
import os  # noqa: E402

_VEC = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors.npz"))
_NAMES = [str(n) for n in _VEC["names"]]


@pytest.mark.parametrize("name", _NAMES)
def test_gpu_matches_golden_fixture(pkg, name):
    p = _VEC[name + "/params"]
    lens = [int(v) for v in _VEC[name + "/lens"]]
    x = _VEC[name + "/x"]
    stride = int(p[5])
    rb = pkg.ResamplerBatch(2, p[0], p[1], max(lens), p[2], p[3], device=0, extfft=int(p[4]))
    pos, ys, counts = 0, [], []
    for l in lens:
        y = rb.process(np.stack([x[pos:pos + l], -0.5 * x[pos:pos + l]]))
        pos += l
        ys.append(y)
        counts.append(y.shape[1])
    assert counts == [int(v) for v in _VEC[name + "/counts"]]
    y = np.concatenate(ys, axis=1)
    m, r = ou.parity_metrics(y[0, ::stride], _VEC[name + "/y_sub"])
    assert m <= MAX_TOL and r <= RMS_TOL, (m / ou.EPS, r / ou.EPS)
    # linearity across channels: channel 1 was fed -0.5*x
    m2, r2 = ou.parity_metrics(y[1], -0.5 * y[0])
    assert m2 <= MAX_TOL and r2 <= RMS_TOL


def test_gpu_drums_kat(pkg):
    """The reference's own golden pair (24-bit WAV), 0.5 s excerpt, both channels in one batch."""
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "drums_excerpt.npz"))
    src, dst = d["src"].astype(np.float64) / 2 ** 23, d["dst"]
    rb = pkg.ResamplerBatch(2, 44100.0, 96000.0, 4096, 2.0, pkg.ATTEN_24, device=0)
    xs = np.concatenate([src.T, np.zeros((2, 8192))], axis=1)
    y = np.concatenate([rb.process(xs[:, i:i + 4096]) for i in range(0, xs.shape[1], 4096)], axis=1)[:, :len(dst)]
    q = np.clip(np.round(y * 2 ** 23), -2 ** 23, 2 ** 23 - 1)
    diff = (q.T - dst)[4800:]  # rmscompare.cpp skips 50 ms at the edges
    assert np.max(np.abs(diff)) <= 1
    assert 20 * np.log10(np.sqrt(np.mean((diff / 2 ** 23) ** 2))) <= -141.0


def test_mirror_class_api(pkg, ref):
    """r8b::CDSPResampler24-shaped object: process / getters / oneshot behave like the reference's."""
    rs = pkg.CDSPResampler24(44100.0, 96000.0, 4096)
    r = ref.Resampler(44100.0, 96000.0, 4096, 2.0, pkg.ATTEN_24)
    assert rs.getMaxOutLen(0) == r.max_out_len
    assert rs.getInLenBeforeOutPos(0) == r.in_len_before_out_pos(0)
    assert rs.getInputRequiredForOutput(1000) == r.input_required_for_output(1000)
    assert rs.getInLenBeforeOutStart(0) == r.in_len_before_out_start(0)
    assert rs.getLatency() == 0 and rs.getLatencyFrac() == r.latency_frac()
    x = ou.white_noise(1, 10000, 21)[0]
    a = rs.oneshot(x, 21000)
    b = r.oneshot(x, 21000)
    m, rr = ou.parity_metrics(a, b)
    assert m <= MAX_TOL and rr <= RMS_TOL
    same = pkg.CDSPResampler(48000.0, 48000.0, 64)
    assert np.array_equal(same.process(x[:64]), x[:64])


def test_hbdown_cascade_extfft(pkg, ref_e1):
    # R8B_EXTFFT=1 doubles the reference's block length: the 1/2 BlockConvolver needs 8192-point tiles
    ys, yr = run_both(pkg, ref_e1, 2822400.0, 44100.0, [65536] * 8, n_ch=2, extfft=1, max_in=65536)
    check(ys, yr)


def test_long_filters(pkg, ref):
    # narrow transition band => long kernels (K = 6817 at 0.5 %): exercises the largest FFT tiles
    for src, dst, tb in [(96000.0, 48000.0, 1.0), (44100.0, 48000.0, 1.0), (48000.0, 44100.0, 3.0)]:
        ys, yr = run_both(pkg, ref, src, dst, [8192] * 4, n_ch=1, tb=tb, max_in=8192)
        check(ys, yr)


def test_fasttiming_parity(pkg):
    """R8B_FASTTIMING=1 plans against the reference compiled with the same macro."""
    if not ou.have_ref("e0_ft"):
        pytest.skip("oracle/_ref fast-timing build missing")
    ref_ft = ou.RefOracle("e0_ft")
    for src, dst in [(48000.0, 47999.0), (44100.0, 22050.5), (192000.0, 44101.0)]:
        lens = [4096] * 10 + [1000, 1, 0, 17, 4096]
        x = ou.white_noise(2, sum(lens), 5)
        plan = pkg.Plan(src, dst, 4096, 2.0, pkg.ATTEN_24, fasttiming=1)
        b = pkg.Batch(plan, 2, 0)
        rs = [ref_ft.Resampler(src, dst, 4096, 2.0, pkg.ATTEN_24) for _ in range(2)]
        pos, ya, yb = 0, [[], []], [[], []]
        for l in lens:
            y = b.process_host(x[:, pos:pos + l])
            for c in range(2):
                r = rs[c].process(x[c, pos:pos + l])
                assert len(r) == y.shape[1]
                ya[c].append(y[c])
                yb[c].append(r)
            pos += l
        check([np.concatenate(a) for a in ya], [np.concatenate(a) for a in yb])


def test_tiny_blocks_and_many_channels(pkg, ref):
    # MaxInLen of a few samples (thousands of calls before the first output) and a wide batch
    for max_in, n_calls in [(1, 2600), (7, 500)]:
        x = ou.white_noise(1, max_in * n_calls, 9)
        rb = pkg.ResamplerBatch(1, 44100.0, 96000.0, max_in, device=0)
        r = ref.Resampler(44100.0, 96000.0, max_in, 2.0, pkg.ATTEN_24)
        ya, yb = [], []
        for c in range(n_calls):
            a = rb.process(x[:, c * max_in:(c + 1) * max_in])
            b = r.process(x[0, c * max_in:(c + 1) * max_in])
            assert a.shape[1] == len(b)
            ya.append(a[0])
            yb.append(b)
        check([np.concatenate(ya)], [np.concatenate(yb)])
    n_ch = 3000
    x = ou.white_noise(n_ch, 3 * 1024, 13)
    rb = pkg.ResamplerBatch(n_ch, 48000.0, 44100.0, 1024, device=0)
    y = np.concatenate([rb.process(x[:, i:i + 1024]) for i in range(0, 3072, 1024)], axis=1)
    for c in (0, 1, 1499, 2999):
        r = ref.Resampler(48000.0, 44100.0, 1024, 2.0, pkg.ATTEN_24)
        yr = np.concatenate([r.process(x[c, i:i + 1024]) for i in range(0, 3072, 1024)])
        assert len(yr) == y.shape[1]
        m, rr = ou.parity_metrics(y[c], yr)
        assert m <= MAX_TOL and rr <= RMS_TOL


@pytest.mark.parametrize("src,dst", [
    (48000.0, 48001.0),     # order-2 bank, rows drift DOWNWARD slowly (staged run, poly_dir = -1)
    (48000.0, 47990.0),     # upward drift fast enough to need several staged chunks per tile pair
    (44100.0, 48001.3),     # rows jump all over the bank: nothing staged, rows read from global memory
    (50000.0, 49999.5),     # very slow drift: one row serves many outputs, wrap-around of the row index
])
def test_poly_bank_row_staging(pkg, ref, src, dst):
    ys, yr = run_both(pkg, ref, src, dst, [8192] * 5 + [333, 8192, 5], n_ch=2, max_in=8192)
    check(ys, yr)


# ---- BASELINE block size (65536 frames per call) on every BASELINE chain: SURVEY.md appendix A count vectors -------------
BASELINE_COUNTS = [
    (48000.0, 44100.0, 0, [58679, 60211, 60211, 60211, 60212, 60211]),          # cfg 3
    (48000.0, 47999.0, 0, [63835, 65535, 65534, 65535, 65535, 65534]),          # cfg 5
    (192000.0, 44100.0, 0, [13516, 15053, 15052, 15053, 15053, 15053]),         # cfg 3b (first chain with CDSPHBDownsampler)
    (44100.0, 2822400.0, 1, [3954306, 4194304, 4194304]),                       # cfg 4 (R8B_EXTFFT = 1)
    (2822400.0, 44100.0, 1, [0, 0, 0, 347, 1024, 1024, 1024, 1024]),            # five half-band decimators (R8B_EXTFFT = 1)
]


@pytest.mark.parametrize("src,dst,extfft,counts", BASELINE_COUNTS)
def test_baseline_block_size_counts_and_parity(pkg, ref, ref_e1, src, dst, extfft, counts):
    oracle = ref_e1 if extfft else ref
    n_ch, l = 2, 65536
    x = ou.white_noise(n_ch, l * len(counts), 11)
    rb = pkg.ResamplerBatch(n_ch, src, dst, l, 2.0, pkg.ATTEN_24, device=0, extfft=extfft)
    rs = [oracle.Resampler(src, dst, l, 2.0, pkg.ATTEN_24) for _ in range(n_ch)]
    for i, want in enumerate(counts):
        y = rb.process(x[:, i * l:(i + 1) * l])
        assert y.shape[1] == want, (i, y.shape[1], want)
        for c in range(n_ch):
            yr = rs[c].process(x[c, i * l:(i + 1) * l])
            assert len(yr) == want
            if want:
                m, r = ou.parity_metrics(y[c], yr)
                assert m <= MAX_TOL and r <= RMS_TOL, (i, c, m / ou.EPS, r / ou.EPS)


def test_bench_shape_device_pointers_on_a_user_stream(pkg, ref):
    """The exact shape bench.py times: 1024 channels x 65536 frames through r8bgpu_batch_process() on a non-default
    stream, padded output rows; sampled channels against the oracle, every channel for finiteness and count."""
    import torch
    n_ch, l, calls = 1024, 65536, 3
    plan = pkg.Plan(44100.0, 96000.0, l, 2.0, pkg.ATTEN_24)
    batch = pkg.Batch(plan, n_ch, 0)
    dev = torch.device("cuda", 0)
    cap = (plan.max_out_len + 7) // 8 * 8
    rng = np.random.default_rng(2024)
    xs = [rng.uniform(-1.0, 1.0, size=(n_ch, l)) for _ in range(2)]
    dx = [torch.from_numpy(a).to(dev) for a in xs]
    out = torch.empty((n_ch, cap), dtype=torch.float64, device=dev)
    st = torch.cuda.Stream(dev)
    batch.set_stream(st.cuda_stream)
    check_ch = [0, 1, 341, 682, 1023]
    rs = {c: ref.Resampler(44100.0, 96000.0, l, 2.0, pkg.ATTEN_24) for c in check_ch}
    want = [138963, 142664, 142663]
    for i in range(calls):
        with torch.cuda.stream(st):
            n = batch.process_ptr(dx[i & 1].data_ptr(), l, l, out.data_ptr(), cap, cap)
        st.synchronize()
        assert n == want[i]
        y = out[:, :n].cpu().numpy()
        assert np.all(np.isfinite(y))
        for c in check_ch:
            yr = rs[c].process(xs[i & 1][c])
            assert len(yr) == n
            m, r = ou.parity_metrics(y[c], yr)
            assert m <= MAX_TOL and r <= RMS_TOL, (i, c, m / ou.EPS, r / ou.EPS)
    batch.set_stream(None)


@pytest.mark.parametrize("src,dst", [(96000.0, 48000.0), (192000.0, 48000.0), (352800.0, 44100.0), (48000.0, 36000.0)])
@pytest.mark.parametrize("tb", [10.0, 20.0, 45.0])
def test_short_kernels_reference_exact_decimation(pkg, ref, src, dst, tb):
    """Wide transition bands make the low-pass kernel short; the reference then runs its power-of-two decimation on
    blocks of 64..512 points, and reference-exact results need tiles of exactly that size (radix-2 transforms)."""
    for atten in (pkg.ATTEN_16IR, pkg.ATTEN_24):
        ys, yr = run_both(pkg, ref, src, dst, [3000, 1, 4096, 777, 4096], n_ch=2, tb=tb, atten=atten, max_in=4096)
        check(ys, yr)
