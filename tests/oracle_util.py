"""ctypes access to the CHECKERS under oracle/ (test infrastructure; never used by the product).

RefOracle   -- oracle/_ref/libr8bref_*.so: the unmodified reference headers compiled by oracle/Makefile.
PortOracle  -- oracle/libr8boracle.so: our own C restatement (oracle/r8b_oracle.c).
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
PORT_LIB = os.path.join(ROOT, "oracle", "libr8boracle.so")

_dp = C.c_void_p


def ref_lib_path(flavor="e0"):
    return os.path.join(REF_DIR, "libr8bref_%s.so" % flavor)


def have_ref(flavor="e0"):
    return os.path.exists(ref_lib_path(flavor))


def have_port():
    return os.path.exists(PORT_LIB)


def cpu_supports_fast():
    try:
        flags = open("/proc/cpuinfo").read()
        return " avx2" in flags and " fma" in flags
    except OSError:
        return False


class RefOracle:
    """flavor: e0 (R8B_EXTFFT=0), e1 (=1), e0_ooura, e0_fast / e1_fast (timing builds)."""

    _cache = {}

    def __new__(cls, flavor="e0"):
        if flavor in cls._cache:
            return cls._cache[flavor]
        self = super().__new__(cls)
        cls._cache[flavor] = self
        self._init(flavor)
        return self

    def _init(self, flavor):
        self.flavor = flavor
        self.name = "reference(%s)" % flavor
        L = C.CDLL(ref_lib_path(flavor))
        self.L = L
        L.r8bref_create.restype = C.c_void_p
        L.r8bref_create.argtypes = [C.c_double, C.c_double, C.c_int, C.c_double, C.c_double]
        L.r8bref_delete.argtypes = [C.c_void_p]
        L.r8bref_clear.argtypes = [C.c_void_p]
        for f in ("r8bref_max_out_len",):
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("r8bref_in_len_before_out_pos", "r8bref_input_required_for_output",
                  "r8bref_in_len_before_out_start"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_int]
        L.r8bref_latency_frac.restype = C.c_double
        L.r8bref_latency_frac.argtypes = [C.c_void_p]
        L.r8bref_process.argtypes = [C.c_void_p, _dp, C.c_int, _dp, C.c_int]
        L.r8bref_oneshot.argtypes = [C.c_void_p, _dp, C.c_int, _dp, C.c_int]
        if hasattr(L, "r8bref_oneshot_typed"):
            L.r8bref_oneshot_typed.restype = C.c_int
            L.r8bref_oneshot_typed.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.r8bref_stage_blockconv.restype = C.c_void_p
        L.r8bref_stage_blockconv.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
        L.r8bref_stage_frac.restype = C.c_void_p
        L.r8bref_stage_frac.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int]
        L.r8bref_stage_hbup.restype = C.c_void_p
        L.r8bref_stage_hbup.argtypes = [C.c_double, C.c_int, C.c_int]
        L.r8bref_stage_hbdown.restype = C.c_void_p
        L.r8bref_stage_hbdown.argtypes = [C.c_double, C.c_int, C.c_int]
        L.r8bref_stage_delete.argtypes = [C.c_void_p]
        L.r8bref_stage_clear.argtypes = [C.c_void_p]
        L.r8bref_stage_max_out_len.argtypes = [C.c_void_p, C.c_int]
        L.r8bref_stage_in_len_before_out_pos.argtypes = [C.c_void_p, C.c_int]
        L.r8bref_stage_process.argtypes = [C.c_void_p, _dp, C.c_int, _dp, C.c_int]
        L.r8bref_lpfilter.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double,
                                      C.POINTER(C.c_int), C.POINTER(C.c_int), _dp, C.c_int]
        L.r8bref_fracbank.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                      C.POINTER(C.c_int), _dp, C.c_long]
        L.r8bref_hbfilter.argtypes = [C.c_double, C.c_int, C.c_int, _dp, C.POINTER(C.c_double)]
        L.r8bref_whole_stepping.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.r8bref_bench.restype = C.c_double
        L.r8bref_bench.argtypes = [C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                                   C.c_int, C.c_int, _dp, C.c_long, C.POINTER(C.c_double), C.POINTER(C.c_long)]
        self.extfft = L.r8bref_extfft()
        oracle = self

        class Resampler:
            def __init__(self, src, dst, max_in_len, tb=2.0, atten=180.15):
                self.h = L.r8bref_create(src, dst, max_in_len, tb, atten)
                self.max_out_len = L.r8bref_max_out_len(self.h)
                self._out = np.empty(self.max_out_len + 16)

            def __del__(self):
                if getattr(self, "h", None):
                    L.r8bref_delete(self.h)
                    self.h = None

            def process(self, x):
                x = np.ascontiguousarray(x, dtype=np.float64)
                n = L.r8bref_process(self.h, x.ctypes.data, len(x), self._out.ctypes.data, len(self._out))
                return self._out[:n].copy()

            def clear(self):
                L.r8bref_clear(self.h)

            def in_len_before_out_pos(self, p):
                return L.r8bref_in_len_before_out_pos(self.h, p)

            def input_required_for_output(self, n):
                return L.r8bref_input_required_for_output(self.h, n)

            def in_len_before_out_start(self, p=0):
                return L.r8bref_in_len_before_out_start(self.h, p)

            def latency_frac(self):
                return L.r8bref_latency_frac(self.h)

            def oneshot(self, x, oplen):
                x = np.ascontiguousarray(x, dtype=np.float64)
                out = np.empty(oplen)
                L.r8bref_oneshot(self.h, x.ctypes.data, len(x), out.ctypes.data, oplen)
                return out

            def oneshot_typed(self, x, oplen, out_dtype):
                """The reference's oneshot<Tin,Tout>() with ITS OWN sample conversions."""
                codes = {"float64": 0, "float32": 1, "int16": 2, "int32": 4}
                x = np.ascontiguousarray(x)
                out = np.empty(oplen, dtype=out_dtype)
                rc = L.r8bref_oneshot_typed(self.h, codes[x.dtype.name], x.ctypes.data, len(x),
                                            codes[np.dtype(out_dtype).name], out.ctypes.data, oplen)
                assert rc == 0
                return out

        class Stage:
            def __init__(self, h):
                self.h = h

            def __del__(self):
                if getattr(self, "h", None):
                    L.r8bref_stage_delete(self.h)
                    self.h = None

            def max_out_len(self, l):
                return L.r8bref_stage_max_out_len(self.h, l)

            def in_len_before_out_pos(self, p):
                return L.r8bref_stage_in_len_before_out_pos(self.h, p)

            def process(self, x):
                x = np.ascontiguousarray(x, dtype=np.float64)
                out = np.empty(self.max_out_len(len(x)) + 16)
                n = L.r8bref_stage_process(self.h, x.ctypes.data, len(x), out.ctypes.data, len(out))
                return out[:n].copy()

            def clear(self):
                L.r8bref_stage_clear(self.h)

        self.Resampler = Resampler
        self._Stage = Stage

    def stage_blockconv(self, norm_freq, tb, atten, gain, up, down):
        return self._Stage(self.L.r8bref_stage_blockconv(norm_freq, tb, atten, gain, up, down))

    def stage_frac(self, src, dst, atten, third=False):
        return self._Stage(self.L.r8bref_stage_frac(src, dst, atten, int(third)))

    def stage_hbup(self, atten, steep, third=False):
        return self._Stage(self.L.r8bref_stage_hbup(atten, steep, int(third)))

    def stage_hbdown(self, atten, steep, third=False):
        return self._Stage(self.L.r8bref_stage_hbdown(atten, steep, int(third)))

    def lpfilter(self, norm_freq, tb, atten, gain):
        """-> dict(kernel_len, block_len_bits, latency, spectrum[0..B2/2])"""
        bits, lat = C.c_int(0), C.c_int(0)
        klen = self.L.r8bref_lpfilter(norm_freq, tb, atten, gain, C.byref(bits), C.byref(lat), None, 0)
        n = (2 << bits.value) // 2 + 1
        sp = np.empty(n)
        self.L.r8bref_lpfilter(norm_freq, tb, atten, gain, C.byref(bits), C.byref(lat), sp.ctypes.data, n)
        return dict(kernel_len=klen, block_len_bits=bits.value, latency=lat.value, spectrum=sp)

    def fracbank(self, init_fracs, elsize, interp_points, atten, third=False):
        fr = C.c_int(0)
        flen = self.L.r8bref_fracbank(init_fracs, elsize, interp_points, atten, int(third), C.byref(fr), None, 0)
        n = (fr.value + 1) * flen * elsize
        t = np.empty(n)
        self.L.r8bref_fracbank(init_fracs, elsize, interp_points, atten, int(third), C.byref(fr), t.ctypes.data, n)
        return dict(filter_len=flen, fracs=fr.value, table=t.reshape(fr.value + 1, flen, elsize))

    def hbfilter(self, atten, steep, third=False):
        taps = np.zeros(16)
        att = C.c_double(0)
        n = self.L.r8bref_hbfilter(atten, steep, int(third), taps.ctypes.data, C.byref(att))
        return taps[:n].copy(), att.value

    def whole_stepping(self, s, d):
        a, b = C.c_int(0), C.c_int(0)
        ok = self.L.r8bref_whole_stepping(s, d, C.byref(a), C.byref(b))
        return bool(ok), a.value, b.value

    def bench(self, src, dst, block_len, tb, atten, x, n_warm, n_calls, n_threads):
        """x: [n_ch, block_len] float64.  Returns (seconds, out_samples_per_channel, checksum)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        cs, no = C.c_double(0), C.c_long(0)
        secs = self.L.r8bref_bench(src, dst, block_len, tb, atten, x.shape[0], n_warm, n_calls, n_threads,
                                   x.ctypes.data, x.shape[1], C.byref(cs), C.byref(no))
        return secs, no.value, cs.value


def best_oracle(extfft=0):
    """The strongest checker available: the compiled reference if present, else the C port."""
    flavor = "e1" if extfft else "e0"
    if have_ref(flavor):
        return RefOracle(flavor)
    from port_oracle import PortOracle  # noqa: deferred; defined next to this file
    return PortOracle(extfft=extfft)


def white_noise(n_ch, n, seed=0):
    """Uniform [-1,1) noise, one independent stream per channel."""
    out = np.empty((n_ch, n))
    for c in range(n_ch):
        rng = np.random.default_rng([seed, c, 0x9E3779B9])
        out[c] = rng.uniform(-1.0, 1.0, n)
    return out


EPS = 2.0 ** -52


def parity_metrics(y, yref):
    """(max abs diff / max|yref|, rms diff / rms yref) -- the tolerance of SURVEY.md section 8c."""
    y = np.asarray(y, dtype=np.float64)
    yref = np.asarray(yref, dtype=np.float64)
    d = y - yref
    mx = float(np.max(np.abs(yref))) if yref.size else 0.0
    rm = float(np.sqrt(np.mean(yref * yref))) if yref.size else 0.0
    if mx == 0.0:
        return (float(np.max(np.abs(d))) if d.size else 0.0, 0.0)
    return float(np.max(np.abs(d))) / mx, float(np.sqrt(np.mean(d * d))) / max(rm, 1e-300)
