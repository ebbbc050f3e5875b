"""CPU tests of the CHECKERS: the C restatement (oracle/r8b_oracle.c) pinned against the reference.

Runs without a GPU.  `oracle/_ref` (the compiled reference) is used when present; the committed
fixtures in tests/golden/ (generated from that same reference by make_golden.py) carry the pin to
machines where /root/reference does not exist.
"""
import os

import numpy as np
import pytest

import oracle_util as ou
from port_oracle import PortOracle

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VEC = np.load(os.path.join(G, "ref_vectors.npz"))
NAMES = [str(n) for n in VEC["names"]]
# chains whose BlockConvolver decimates by a power of two: the reference keeps only the low part of
# each block spectrum (CDSPBlockConvolver.h:329-344), the ideal operator restated by the port does not
BLOCK_SPECTRUM_CASES = {"half_96000_48000", "up3_32000_48000"}


def case(name):
    p = VEC[name + "/params"]
    return dict(src=p[0], dst=p[1], tb=p[2], att=p[3], ext=int(p[4]), stride=int(p[5]), max_out=int(p[6]),
                ilb0=int(p[7]), ilb1000=int(p[8]), lens=[int(v) for v in VEC[name + "/lens"]],
                counts=[int(v) for v in VEC[name + "/counts"]], x=VEC[name + "/x"], y_sub=VEC[name + "/y_sub"],
                stats=VEC[name + "/y_stats"])


def run(res, c):
    pos, ys, counts = 0, [], []
    for l in c["lens"]:
        y = res.process(c["x"][pos:pos + l])
        pos += l
        ys.append(y)
        counts.append(len(y))
    return np.concatenate(ys), counts


@pytest.mark.parametrize("name", NAMES)
def test_port_matches_golden(name):
    c = case(name)
    r = PortOracle(extfft=c["ext"]).Resampler(c["src"], c["dst"], max(c["lens"]), c["tb"], c["att"])
    y, counts = run(r, c)
    assert counts == c["counts"]
    assert r.max_out_len == c["max_out"]
    assert r.in_len_before_out_pos(0) == c["ilb0"] and r.in_len_before_out_pos(1000) == c["ilb1000"]
    m, rms = ou.parity_metrics(y[::c["stride"]], c["y_sub"])
    if name in BLOCK_SPECTRUM_CASES:
        assert m < 1e-9 and rms < 1e-9
    else:
        # the port accumulates in long double; the reference's FFT path sits ~2 eps rms from it
        assert m <= 16 * ou.EPS and rms <= 4 * ou.EPS, (m / ou.EPS, rms / ou.EPS)


@pytest.mark.parametrize("name", NAMES)
def test_reference_reproduces_golden(name):
    c = case(name)
    flavor = "e1" if c["ext"] else "e0"
    if not ou.have_ref(flavor):
        pytest.skip("oracle/_ref not built")
    r = ou.RefOracle(flavor).Resampler(c["src"], c["dst"], max(c["lens"]), c["tb"], c["att"])
    y, counts = run(r, c)
    assert counts == c["counts"]
    assert np.array_equal(y[::c["stride"]], c["y_sub"])
    assert len(y) == int(c["stats"][0]) and abs(y.sum() - c["stats"][1]) <= 1e-9 * max(1.0, abs(c["stats"][1]))


def test_port_vs_reference_random_chunking():
    if not ou.have_ref("e0"):
        pytest.skip("oracle/_ref not built")
    ref = ou.RefOracle("e0")
    rng = np.random.default_rng(42)
    for src, dst in [(44100.0, 96000.0), (48000.0, 47999.0), (192000.0, 44100.0), (44100.0, 88200.0)]:
        lens = [int(v) for v in rng.integers(0, 3000, 8)]
        x = rng.uniform(-1, 1, sum(lens))
        a = ref.Resampler(src, dst, 3000)
        b = PortOracle().Resampler(src, dst, 3000)
        pos = 0
        ya, yb = [], []
        for l in lens:
            p, q = a.process(x[pos:pos + l]), b.process(x[pos:pos + l])
            assert len(p) == len(q)
            ya.append(p)
            yb.append(q)
            pos += l
        m, rms = ou.parity_metrics(np.concatenate(yb), np.concatenate(ya))
        assert m <= 16 * ou.EPS and rms <= 4 * ou.EPS


EDGE = 4800  # 50 ms at 96 kHz


def _drums_check(process_fn, max_in):
    d = np.load(os.path.join(G, "drums_excerpt.npz"))
    src, dst = d["src"].astype(np.float64) / 2 ** 23, d["dst"]
    for ch in range(2):
        y = process_fn(ch, src[:, ch], len(dst))
        q = np.clip(np.round(y * 2 ** 23), -2 ** 23, 2 ** 23 - 1)
        # the reference's own comparison tool skips 50 ms at the file edges (bench/rmscompare.cpp:80-91);
        # the author's file differs by up to 3 LSB in its first 6 samples (dither/start-up of his tool)
        diff = (q - dst[:, ch])[EDGE:]
        assert np.max(np.abs(diff)) <= 1, "more than 1 LSB from the author's 24-bit conversion"
        rms_db = 20 * np.log10(np.sqrt(np.mean((diff / 2 ** 23) ** 2)) + 1e-30)
        assert rms_db <= -141.0, rms_db  # bench/README.md:9-11 "files are equal" threshold


def _feed_until(res, x, oplen, max_in):
    out, pos = [], 0
    got = 0
    while got < oplen:
        if pos < len(x):
            chunk = x[pos:pos + max_in]
            pos += len(chunk)
        else:
            chunk = np.zeros(max_in)
        y = res.process(chunk)
        out.append(y)
        got += len(y)
    return np.concatenate(out)[:oplen]


def test_port_drums_kat():
    """The reference's only golden vector (bench/DrumsSrc.wav -> DrumsDst96.wav), 0.5 s excerpt."""
    def f(ch, x, oplen):
        return _feed_until(PortOracle().Resampler(44100.0, 96000.0, 4096, 2.0, 180.15), x, oplen, 4096)
    _drums_check(f, 4096)


def test_reference_drums_kat_full_file():
    wav = "/root/reference/bench/DrumsSrc.wav"
    if not (ou.have_ref("e0") and os.path.exists(wav)):
        pytest.skip("needs /root/reference and oracle/_ref")
    import sys
    sys.path.insert(0, G)
    from make_golden import read_wav24
    _, s = read_wav24(wav)
    _, d = read_wav24("/root/reference/bench/DrumsDst96.wav")
    ref = ou.RefOracle("e0")
    for ch in range(2):
        r = ref.Resampler(44100.0, 96000.0, 65536, 2.0, 180.15)
        y = _feed_until(r, s[:, ch].astype(np.float64) / 2 ** 23, len(d), 65536)
        q = np.clip(np.round(y * 2 ** 23), -2 ** 23, 2 ** 23 - 1)
        diff = (q - d[:, ch])[EDGE:-EDGE]
        assert np.max(np.abs(diff)) <= 1
        assert 20 * np.log10(np.sqrt(np.mean((diff / 2 ** 23) ** 2))) <= -141.0


# ---- sample-format semantics: what "(double) ip[i]" / "(Tout) op[i]" do in the reference ----------
@pytest.mark.parametrize("in_dtype,out_dtype", [
    (np.int16, np.float32), (np.int16, np.int16), (np.int16, np.int32), (np.float32, np.int16), (np.int32, np.float32),
])
def test_reference_oneshot_casts_are_the_c_cast_model(ref, in_dtype, out_dtype):
    """The GPU sample-format tests compare against 'fp64 path + C conversion in numpy' (tests/test_gpu_formats.py::c_cast).
    This pins that model to the reference itself: oneshot<Tin,Tout>() == c_cast(oneshot<double,double>(widened input))."""
    rng = np.random.default_rng(3)
    if np.issubdtype(in_dtype, np.integer):
        x = rng.integers(-12000, 12000, size=5000).astype(in_dtype)
    else:
        x = (rng.uniform(-12000, 12000, size=5000)).astype(in_dtype)
    oplen = 10000
    a = ref.Resampler(44100.0, 96000.0, 1024, 2.0, 180.15).oneshot_typed(x, oplen, out_dtype)
    y = ref.Resampler(44100.0, 96000.0, 1024, 2.0, 180.15).oneshot(x.astype(np.float64), oplen)
    if out_dtype == np.float32:
        want = y.astype(np.float32)                       # round to nearest
    else:
        info = np.iinfo(out_dtype)
        assert y.min() > info.min and y.max() < info.max  # in range: the reference's cast is defined
        want = np.trunc(y).astype(out_dtype)              # toward zero
    assert np.array_equal(a, want)
