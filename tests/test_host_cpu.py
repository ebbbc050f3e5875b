"""CPU tests of the product's host side: C-ABI surface, planner, filter design, scheduler.

No compute call is made (there is no CPU fallback); everything here is the plan-level API.
"""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle_util as ou
from port_oracle import PortOracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
VEC = np.load(os.path.join(G, "ref_vectors.npz"))
NAMES = [str(n) for n in VEC["names"]]


def test_library_exports_every_declared_symbol(pkg):
    hdr = open(os.path.join(ROOT, "include", "r8bgpu.h")).read()
    declared = set(re.findall(r"R8BGPU_API[^;(]*?\b(r8bgpu_\w+)\s*\(", hdr))
    assert len(declared) >= 30
    L = ctypes.CDLL(pkg.lib_path())
    for name in sorted(declared):
        assert hasattr(L, name), "libr8bgpu.so does not export " + name
    assert declared == set(pkg._SYMBOLS), declared ^ set(pkg._SYMBOLS)


def test_cubin_is_sm100a(pkg):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", pkg.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


@pytest.mark.parametrize("name", NAMES)
def test_scheduler_matches_golden_counts(pkg, name):
    p = VEC[name + "/params"]
    lens = [int(v) for v in VEC[name + "/lens"]]
    plan = pkg.Plan(p[0], p[1], max(lens), p[2], p[3], extfft=int(p[4]))
    assert plan.simulate(lens) == [int(v) for v in VEC[name + "/counts"]]
    assert plan.max_out_len == int(p[6])
    assert plan.in_len_before_out_pos(0) == int(p[7])
    assert plan.in_len_before_out_pos(1000) == int(p[8])


RATES = [(44100.0, 96000.0), (48000.0, 44100.0), (48000.0, 47999.0), (192000.0, 44100.0), (44100.0, 88200.0),
         (44100.0, 176400.0), (96000.0, 48000.0), (44100.0, 192000.0), (44100.0, 132300.0), (48000.0, 32000.0),
         (32000.0, 48000.0), (48000.0, 36000.0), (2822400.0, 44100.0), (48000.0, 16000.0), (8000.0, 48000.0),
         (44100.0, 22050.5), (1.0, 2.0), (44100.0, 44100.0), (11025.0, 96000.0), (96000.0, 11025.0)]


@pytest.mark.parametrize("src,dst", RATES)
def test_plan_equals_port_design(pkg, src, dst):
    """Product planner/designer (C++) and the oracle port (C) are separate implementations of the
    same reference formulas on the same libm: chains and filter data must agree bit for bit."""
    plan = pkg.Plan(src, dst, 4096, 2.0, pkg.ATTEN_24)
    r = PortOracle().Resampler(src, dst, 4096, 2.0, pkg.ATTEN_24)
    st = plan.stages()
    assert [s["kind"] for s in st] == r.stage_kinds()
    assert plan.max_out_len == r.max_out_len
    for i in range(len(st)):
        assert np.array_equal(plan.stage_data(i), r.stage_data(i)), "stage %d data differs" % i
    for pos in (0, 1, 999, 54321):
        assert plan.in_len_before_out_pos(pos) == r.in_len_before_out_pos(pos)


@pytest.mark.parametrize("src,dst", RATES)
def test_scheduler_equals_reference(pkg, ref, src, dst):
    rng = np.random.default_rng(int(src + dst))
    lens = [4096, 4096] + [int(v) for v in rng.integers(0, 4097, 10)] + [0, 1, 4096]
    plan = pkg.Plan(src, dst, 4096, 2.0, pkg.ATTEN_24)
    r = ref.Resampler(src, dst, 4096, 2.0, pkg.ATTEN_24)
    x = np.zeros(4096)
    assert plan.simulate(lens) == [len(r.process(x[:l])) for l in lens]
    assert plan.max_out_len == r.max_out_len
    for n in (1, 10, 1000):
        assert plan.input_required_for_output(n) == r.input_required_for_output(n)


def test_in_len_before_out_start_matches_reference(pkg, ref):
    for src, dst in [(44100.0, 96000.0), (48000.0, 44100.0), (192000.0, 44100.0)]:
        plan = pkg.Plan(src, dst, 1024, 2.0, pkg.ATTEN_24)
        r = ref.Resampler(src, dst, 1024, 2.0, pkg.ATTEN_24)
        cs = np.cumsum(plan.simulate([1] * 8192))
        mine = int(np.nonzero(cs > 0)[0][0])
        assert mine == r.in_len_before_out_start(0)
        # the two "instant" functions agree with the iterative one, as bench/zerotest.cpp:115-128 checks
        assert plan.in_len_before_out_pos(0) == mine


def test_lowpass_spectrum_matches_reference(pkg, ref):
    for nf, tb, att, gain in [(0.5, 2.0, 180.15, 2.0), (0.459375, 2.0, 180.15, 2.0), (0.5, 0.5, 218.0, 2.0),
                              (1 / 3, 30.0, 109.56, 3.0), (0.3, 5.0, 55.0, 1.0), (0.5, 45.0, 49.0, 1.0)]:
        p = pkg.Plan.single_stage(0, [nf, tb, att, gain, 1, 1], 1024)
        st = p.stages()[0]
        h = p.stage_data(0)
        r = ref.lpfilter(nf, tb, att, gain)
        assert st["kernel_len"] == r["kernel_len"] and st["block_len_bits"] == r["block_len_bits"]
        K = st["kernel_len"]
        L = (K - 1) // 2
        b2 = 2 << r["block_len_bits"]
        z = np.zeros(b2)
        z[:L + 1] = h[L:]
        z[b2 - L:] = h[:L]
        H = np.fft.rfft(z).real
        assert np.max(np.abs(H - r["spectrum"])) <= 8 * ou.EPS * gain  # both sides carry FFT rounding


def test_frac_banks_and_halfbands_bit_exact(pkg, ref):
    for src, dst, att, third in [(88200.0, 96000.0, 180.15, 0), (96000.0, 44100.0, 180.15, 0),
                                 (48000.0, 47999.0, 180.15, 0), (48000.0, 47999.0, 136.45, 0),
                                 (48000.0, 47999.0, 109.56, 1), (88200.0, 96000.0, 206.91, 1)]:
        ok, a, b = ref.whole_stepping(src, dst)
        r = ref.fracbank(b if ok else -1, 1 if ok else 3, 2 if ok else 8, att, bool(third))
        p = pkg.Plan.single_stage(1, [src, dst, att, third], 1024)
        assert np.array_equal(p.stage_data(0).reshape(r["table"].shape), r["table"])
    for third in (0, 1):
        for steep in range(8):
            for att in (50.0, 100.0, 136.45, 180.15, 206.91, 300.0):
                t, a = ref.hbfilter(att, steep, bool(third))
                p = pkg.Plan.single_stage(3, [att, steep, third], 1024)
                assert np.array_equal(p.stage_data(0), t) and p.stages()[0]["atten"] == a


def test_passthrough_and_errors(pkg):
    assert pkg.Plan(48000.0, 48000.0, 256).passthrough
    for bad in [dict(src_rate=0.0, dst_rate=1.0, max_in_len=16), dict(src_rate=1.0, dst_rate=2.0, max_in_len=0),
                dict(src_rate=1.0, dst_rate=2.0, max_in_len=16, phase=1),
                dict(src_rate=1.0, dst_rate=2.0, max_in_len=16, trans_band=0.1),
                dict(src_rate=1.0, dst_rate=2.0, max_in_len=16, atten=300.0)]:
        with pytest.raises(pkg.R8bGpuError):
            pkg.Plan(**bad)
    with pytest.raises(pkg.R8bGpuError):
        pkg.Plan(44100.0, 96000.0, 64).simulate([65])  # l > MaxInLen


def test_no_cpu_fallback(pkg):
    """Without a CUDA device the batch API must fail loudly, never compute on the host."""
    if pkg.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(pkg.R8bGpuError):
        pkg.Batch(pkg.Plan(44100.0, 96000.0, 64), 2)
    with pytest.raises(pkg.R8bGpuError):
        pkg.CDSPResampler24(44100.0, 96000.0, 64)


def test_product_never_touches_oracle():
    """The shipped sources must not reference oracle/ (only tests, smoke() and bench.py may)."""
    pdir = os.path.join(ROOT, "r8brain-free-src_b200")
    for base, _, files in os.walk(pdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".inc")):
                txt = open(os.path.join(base, f), errors="replace").read()
                assert "oracle" not in txt.lower() or f == "__init__.py" and "oracle" not in txt.lower(), f
    inc = open(os.path.join(ROOT, "include", "r8bgpu.h")).read()
    assert "oracle" not in inc.lower()


def test_fasttiming_scheduler_matches_reference(pkg):
    """R8B_FASTTIMING=1 (r8bconf.h:123-132): the drifting-accumulator timing of the non-whole interpolator."""
    if not ou.have_ref("e0_ft"):
        pytest.skip("oracle/_ref fast-timing build missing")
    ref_ft = ou.RefOracle("e0_ft")
    assert ref_ft.L.r8bref_fasttiming() == 1
    for src, dst in [(48000.0, 47999.0), (44100.0, 22050.5), (44100.0, 96001.0), (44100.0, 96000.0)]:
        lens = [4096] * 6 + [1000, 0, 1, 4096]
        plan = pkg.Plan(src, dst, 4096, 2.0, pkg.ATTEN_24, fasttiming=1)
        r = ref_ft.Resampler(src, dst, 4096, 2.0, pkg.ATTEN_24)
        x = np.zeros(4096)
        assert plan.simulate(lens) == [len(r.process(x[:l])) for l in lens]


@pytest.mark.skipif(not ou.have_ref('e0'), reason='compiled reference not present')
def test_whole_stepping_decision_matches_reference(pkg):
    """Whole-stepping vs order-2 bank is decided by a bounded subtractive GCD on the doubles
    (CDSPFracInterpolator.h:609-673): the iteration budget and the OutStep limit are part of the plan."""
    ref = ou.RefOracle("e0")
    pairs = [(float(n), 1.0) for n in range(1, 160)] + [(1.0, float(n)) for n in range(1, 160)]
    pairs += [(float(a), float(b)) for a in (147, 148, 149, 150, 151, 233, 377, 1499, 1500, 1501, 1502, 3001)
              for b in (1, 2, 3, 89, 144, 1500, 1501, 2999)]
    pairs += [(44100.0, 96000.0), (48000.0, 47999.0), (44100.5, 96000.0), (0.1, 0.3), (96000.0, 88200.0)]
    for src, dst in pairs:
        if src == dst:
            continue
        ok, a, b = ref.whole_stepping(src, dst)
        st = pkg.Plan.single_stage(1, [src, dst, 180.15, 0], 1024).stages()[0]
        assert (st["order"] == 0) == ok, (src, dst, ok, st)
        if ok:
            assert (st["in_step"], st["out_step"]) == (a, b), (src, dst, a, b, st)


def test_multi_device_host_plumbing(tmp_path):
    """csrc/r8b_multi.cpp on the CPU: shard worker pool (results by shard, error relay from the worker's own thread,
    20000 back-to-back rounds), NUMA thread placement inside the process mask, pinned allocator refusing without a device."""
    import shutil
    import subprocess
    cuda = next((d for d in (os.environ.get("CUDA_HOME"), "/usr/local/cuda")
                 if d and os.path.exists(os.path.join(d, "include", "cuda_runtime.h"))), None)
    if cuda is None or shutil.which("g++") is None:
        pytest.skip("CUDA headers / g++ not found")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "multi_host")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I" + os.path.join(cuda, "include"), "-o", exe,
                    os.path.join(root, "tests", "cpp", "multi_host.cpp"),
                    os.path.join(root, "r8brain-free-src_b200", "csrc", "r8b_multi.cpp"),
                    "-L" + os.path.join(cuda, "lib64"), "-lcudart_static", "-ldl", "-lrt"], check=True)
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0 and "FAIL" not in r.stdout and "OK" in r.stdout, r.stdout


@pytest.mark.skipif(not (ou.have_ref('e0') and ou.have_ref('e1')), reason='compiled reference not present')
def test_scheduler_random_sweep_matches_reference(pkg):
    """Seeded sweep over rate pairs x presets x transition bands x R8B_EXTFFT with ragged call lengths: per-call output
    counts, getMaxOutLen, getInLenBeforeOutPos and getInputRequiredForOutput are the reference's integers
    (CDSPResampler.h:117-214, 559-575) on every planner branch the sweep reaches."""
    rng = np.random.default_rng(20260923)
    rates = [8000.0, 11025.0, 16000.0, 22050.0, 32000.0, 44100.0, 48000.0, 88200.0, 96000.0, 176400.0, 192000.0,
             352800.0, 384000.0, 2822400.0, 47999.0, 44100.5, 12345.0, 96001.0]
    attens = [pkg.ATTEN_16, pkg.ATTEN_16IR, pkg.ATTEN_24, 206.91, 60.0]
    tbs = [0.5, 2.0, 3.0, 7.0, 20.0, 45.0]
    refs = {0: ou.RefOracle("e0"), 1: ou.RefOracle("e1")}
    kinds = set()
    n_done = 0
    for _ in range(400):
        src, dst = (float(v) for v in rng.choice(rates, 2, replace=False))
        att, tb, ext = float(rng.choice(attens)), float(rng.choice(tbs)), int(rng.integers(0, 2))
        max_len = int(rng.choice([64, 1000, 4096]))
        lens = [max_len] * 3 + [int(v) for v in rng.integers(0, max_len + 1, 8)] + [0, 1, max_len]
        plan = pkg.Plan(src, dst, max_len, tb, att, extfft=ext)
        r = refs[ext].Resampler(src, dst, max_len, tb, att)
        x = np.zeros(max_len)
        what = (src, dst, att, tb, ext, max_len)
        assert plan.simulate(lens) == [len(r.process(x[:l])) for l in lens], what
        assert plan.max_out_len == r.max_out_len, what
        for pos in (0, 17, 1000):
            assert plan.in_len_before_out_pos(pos) == r.in_len_before_out_pos(pos), what
        for n in (1, 999):
            assert plan.input_required_for_output(n) == r.input_required_for_output(n), what
        kinds.update(s["kind"] for s in plan.stages())
        n_done += 1
    assert n_done == 400 and len(kinds) >= 5, kinds  # BlockConv, both interpolators, both half-band stages were reached


def test_filter_design_random_sweep_matches_reference(pkg, ref):
    """Seeded sweep of the host filter design against the reference's (CDSPFIRFilter.h:220-537,
    CDSPFracInterpolator.h:61-189): low-pass kernel length, block size and spectrum over random cut-offs, transition
    bands, attenuations and gains; fractional-delay banks bit for bit over random whole-stepping / order-2 ratios."""
    rng = np.random.default_rng(77)
    for _ in range(60):
        nf = float(rng.choice([0.5, 1 / 3, 0.25, 0.125, float(rng.uniform(0.05, 0.5))]))
        tb = float(rng.choice([0.5, 1.0, 2.0, 3.0, 7.5, 20.0, 45.0, float(rng.uniform(0.5, 45.0))]))
        att = float(rng.choice([49.0, 60.0, 109.56, 136.45, 180.15, 206.91, 218.0, float(rng.uniform(49.0, 218.0))]))
        gain = float(rng.choice([1.0, 2.0, 3.0, 0.5, 0.03125]))
        p = pkg.Plan.single_stage(0, [nf, tb, att, gain, 1, 1], 1024)
        st = p.stages()[0]
        r = ref.lpfilter(nf, tb, att, gain)
        what = (nf, tb, att, gain)
        assert st["kernel_len"] == r["kernel_len"] and st["block_len_bits"] == r["block_len_bits"], what
        h = p.stage_data(0)
        L = (st["kernel_len"] - 1) // 2
        b2 = 2 << r["block_len_bits"]
        z = np.zeros(b2)
        z[:L + 1] = h[L:]
        z[b2 - L:] = h[:L]
        assert np.max(np.abs(np.fft.rfft(z).real - r["spectrum"])) <= 8 * ou.EPS * gain, what
        # DC gain (what the chain's level rests on): the exact tap sum is the reference's DC bin
        import math
        assert abs(math.fsum(h) - r["spectrum"][0]) <= 8 * ou.EPS * gain, what
    rates = [44100.0, 48000.0, 88200.0, 96000.0, 47999.0, 32000.0, 12345.0, 96001.0, 176400.0, 22050.5]
    for _ in range(30):
        src, dst = (float(v) for v in rng.choice(rates, 2, replace=False))
        att, third = float(rng.choice([109.56, 136.45, 180.15, 206.91])), int(rng.integers(0, 2))
        ok, a, b = ref.whole_stepping(src, dst)
        r = ref.fracbank(b if ok else -1, 1 if ok else 3, 2 if ok else 8, att, bool(third))
        p = pkg.Plan.single_stage(1, [src, dst, att, third], 1024)
        assert np.array_equal(p.stage_data(0).reshape(r["table"].shape), r["table"]), (src, dst, att, third)
