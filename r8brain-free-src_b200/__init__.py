"""r8brain-free-src_b200 -- Python mirror of the reference's resampler interface over the B200 C-ABI.

The product is libr8bgpu.so (hand-written sm_100a kernels + host planner, see csrc/ and
include/r8bgpu.h).  This module is a thin ctypes binding that mirrors the reference's
front-end names and argument meaning for the process() path:

    r8b::CDSPResampler(Src, Dst, MaxInLen, ReqTransBand=2, ReqAtten=206.91)   CDSPResampler.h:117-120
    r8b::CDSPResampler16 / 16IR / 24                                           CDSPResampler.h:729-810
    process / clear / oneshot / getMaxOutLen / getInLenBeforeOutPos /
    getInputRequiredForOutput / getInLenBeforeOutStart / getLatency[Frac]      CDSPResampler.h:406-651

plus `ResamplerBatch`, the channel-batched form of example.cpp:30-67 (one resampler per channel,
same block length for every channel).  There is no CPU fallback: constructing a batch without a
CUDA device raises.  The directory name contains '-', so import it through
`__graft_entry__.load_package()` (registers it as module `r8brain_free_src_b200`).
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

STAGE_NAMES = {0: "blockconv", 1: "frac_whole", 2: "frac_poly", 3: "hbup", 4: "hbdown"}
ATTEN_16 = 136.45   # CDSPResampler16   (CDSPResampler.h:745-746)
ATTEN_16IR = 109.56  # CDSPResampler16IR (:776-777)
ATTEN_24 = 180.15   # CDSPResampler24   (:806-807)


class StageInfo(C.Structure):
    _fields_ = [("kind", C.c_int), ("up", C.c_int), ("down", C.c_int), ("kernel_len", C.c_int),
                ("latency", C.c_int), ("ref_input_len", C.c_int), ("block_len_bits", C.c_int),
                ("fracs", C.c_int), ("in_step", C.c_int), ("out_step", C.c_int), ("order", C.c_int),
                ("max_out_len", C.c_int), ("atten", C.c_double), ("data_len", C.c_int)]


# Every symbol include/r8bgpu.h declares: name -> (restype, argtypes)
_SYMBOLS = {
    "r8bgpu_last_error": (C.c_char_p, []),
    "r8bgpu_version": (C.c_char_p, []),
    "r8bgpu_plan_create": (C.c_void_p, [C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int]),
    "r8bgpu_plan_create_stage": (C.c_void_p, [C.c_int, C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int]),
    "r8bgpu_plan_destroy": (None, [C.c_void_p]),
    "r8bgpu_plan_max_out_len": (C.c_int, [C.c_void_p]),
    "r8bgpu_plan_in_len_before_out_pos": (C.c_int, [C.c_void_p, C.c_int]),
    "r8bgpu_plan_input_required_for_output": (C.c_int, [C.c_void_p, C.c_int]),
    "r8bgpu_plan_latency_frac": (C.c_double, [C.c_void_p]),
    "r8bgpu_plan_is_passthrough": (C.c_int, [C.c_void_p]),
    "r8bgpu_plan_stage_count": (C.c_int, [C.c_void_p]),
    "r8bgpu_plan_stage_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(StageInfo)]),
    "r8bgpu_plan_stage_data": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "r8bgpu_plan_describe": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "r8bgpu_plan_simulate": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]),
    "r8bgpu_device_count": (C.c_int, []),
    "r8bgpu_batch_create": (C.c_void_p, [C.c_void_p, C.c_int, C.c_int]),
    "r8bgpu_batch_destroy": (None, [C.c_void_p]),
    "r8bgpu_batch_shard_count": (C.c_int, [C.c_void_p]),
    "r8bgpu_batch_shard_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "r8bgpu_batch_shard": (C.c_void_p, [C.c_void_p, C.c_int]),
    "r8bgpu_batch_host_alloc": (C.c_void_p, [C.c_void_p, C.c_size_t, C.c_int]),
    "r8bgpu_batch_clear": (C.c_int, [C.c_void_p]),
    "r8bgpu_batch_channels": (C.c_int, [C.c_void_p]),
    "r8bgpu_batch_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "r8bgpu_batch_process": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.c_int]),
    "r8bgpu_batch_process_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.c_int]),
    "r8bgpu_batch_process_fmt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "r8bgpu_batch_process_host_fmt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "r8bgpu_batch_sync": (C.c_int, [C.c_void_p]),
    "r8bgpu_batch_kernel_launches": (C.c_ulonglong, [C.c_void_p]),
    "r8bgpu_batch_device_bytes": (C.c_ulonglong, [C.c_void_p]),
    "r8bgpu_batch_stage_kernel": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_int]),
    "r8bgpu_batch_set_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "r8bgpu_batch_stage_time_ms": (C.c_double, [C.c_void_p, C.c_int, C.POINTER(C.c_ulonglong)]),
    "r8bgpu_measure_fp64_tflops": (C.c_double, [C.c_int]),
    "r8bgpu_host_alloc": (C.c_void_p, [C.c_size_t]),
    "r8bgpu_host_free": (None, [C.c_void_p]),
}


class R8bGpuError(RuntimeError):
    pass


# r8bgpu_sample_format / r8bgpu_buffer (include/r8bgpu.h)
F64, F32, S16, S24, S32 = 0, 1, 2, 3, 4
FORMAT_BYTES = {F64: 8, F32: 4, S16: 2, S24: 3, S32: 4}
_NP_FORMATS = {"float64": F64, "float32": F32, "int16": S16, "int32": S32}


class Buffer(C.Structure):
    _fields_ = [("data", C.c_void_p), ("format", C.c_int), ("interleaved", C.c_int),
                ("stride", C.c_size_t), ("scale", C.c_double)]

    @classmethod
    def make(cls, ptr, fmt, interleaved, stride, scale=1.0):
        return cls(C.c_void_p(int(ptr) if ptr else None), int(fmt), int(bool(interleaved)), int(stride), float(scale))


def lib_path():
    return _build.LIB


def lib():
    """Load (building first if stale and nvcc is present) the C-ABI library."""
    global _lib
    if _lib is None:
        path = _build.build()
        L = C.CDLL(path)
        for name, (res, args) in _SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError here == header/library mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _err():
    return lib().r8bgpu_last_error().decode("utf-8", "replace")


def device_count():
    return lib().r8bgpu_device_count()


def measure_fp64_tflops(device=-1):
    """Measured DFMA ceiling of the device (bench.py's secondary roofline)."""
    v = lib().r8bgpu_measure_fp64_tflops(int(device))
    if v < 0:
        raise R8bGpuError(_err())
    return v


class Plan:
    """Immutable stage chain + filters for (src, dst): what CDSPResampler's constructor decides."""

    def __init__(self, src_rate, dst_rate, max_in_len, trans_band=2.0, atten=206.91, phase=0,
                 extfft=0, fasttiming=0, _handle=None):
        self._h = _handle if _handle is not None else lib().r8bgpu_plan_create(
            float(src_rate), float(dst_rate), int(max_in_len), float(trans_band), float(atten),
            int(phase), int(extfft), int(fasttiming))
        if not self._h:
            raise R8bGpuError(_err())
        self.src_rate, self.dst_rate, self.max_in_len = float(src_rate), float(dst_rate), int(max_in_len)

    @classmethod
    def single_stage(cls, kind, params, max_in_len, extfft=0):
        arr = (C.c_double * len(params))(*[float(p) for p in params])
        h = lib().r8bgpu_plan_create_stage(int(kind), arr, len(params), int(max_in_len), int(extfft))
        if not h:
            raise R8bGpuError(_err())
        return cls(0.0, 0.0, max_in_len, _handle=h)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.r8bgpu_plan_destroy(self._h)
            self._h = None

    @property
    def max_out_len(self):
        return lib().r8bgpu_plan_max_out_len(self._h)

    @property
    def passthrough(self):
        return bool(lib().r8bgpu_plan_is_passthrough(self._h))

    def in_len_before_out_pos(self, pos):
        return lib().r8bgpu_plan_in_len_before_out_pos(self._h, int(pos))

    def input_required_for_output(self, n):
        return lib().r8bgpu_plan_input_required_for_output(self._h, int(n))

    def latency_frac(self):
        return lib().r8bgpu_plan_latency_frac(self._h)

    def stages(self):
        out = []
        for i in range(lib().r8bgpu_plan_stage_count(self._h)):
            info = StageInfo()
            if lib().r8bgpu_plan_stage_info(self._h, i, C.byref(info)) != 0:
                raise R8bGpuError(_err())
            d = {f: getattr(info, f) for f, _ in StageInfo._fields_}
            d["name"] = STAGE_NAMES[info.kind]
            out.append(d)
        return out

    def stage_data(self, i):
        n = lib().r8bgpu_plan_stage_data(self._h, int(i), None, 0)
        if n < 0:
            raise R8bGpuError(_err())
        a = np.empty(n, dtype=np.float64)
        lib().r8bgpu_plan_stage_data(self._h, int(i), a.ctypes.data, n)
        return a

    def describe(self):
        n = lib().r8bgpu_plan_describe(self._h, None, 0)
        buf = C.create_string_buffer(n + 1)
        lib().r8bgpu_plan_describe(self._h, buf, n + 1)
        return buf.value.decode()

    def simulate(self, lens):
        """Per-call output counts the scheduler would return (CPU only, no GPU needed)."""
        lens = [int(v) for v in lens]
        a = (C.c_int * len(lens))(*lens)
        o = (C.c_int * len(lens))()
        if lib().r8bgpu_plan_simulate(self._h, a, len(lens), o) != 0:
            raise R8bGpuError(_err())
        return list(o)


DEVICE_ALL, DEVICE_CURRENT = -1, -2
_host_allocs = {}


def host_free(arr):
    """Release an array from Batch.host_alloc()."""
    p = _host_allocs.pop(arr.ctypes.data, None)
    if p is not None:
        lib().r8bgpu_host_free(C.c_void_p(p))


class Batch:
    """n_channels independent streams resampled in lock-step: on one GPU (device >= 0, or DEVICE_CURRENT), or sharded
    over every visible GPU behind the C-ABI (DEVICE_ALL: host-buffer calls only)."""

    def __init__(self, plan, n_channels, device=-2):
        self.plan = plan
        self.n_channels = int(n_channels)
        self._h = lib().r8bgpu_batch_create(plan._h, self.n_channels, int(device))
        if not self._h:
            raise R8bGpuError(_err())

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.r8bgpu_batch_destroy(self._h)
            self._h = None

    def shards(self):
        """[(device, first_channel, n_channels, numa_node)] -- one entry for a single-device batch."""
        out = []
        for i in range(lib().r8bgpu_batch_shard_count(self._h)):
            d, c0, n, node = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
            if lib().r8bgpu_batch_shard_info(self._h, i, C.byref(d), C.byref(c0), C.byref(n), C.byref(node)) != 0:
                raise R8bGpuError(_err())
            out.append((d.value, c0.value, n.value, node.value))
        return out

    def host_alloc(self, samples_per_channel, dtype="float64"):
        """Pinned planar [n_channels, samples_per_channel] numpy array, rows on the NUMA node of the owning GPU.
        Keep the returned array alive while in use; release with host_free(arr)."""
        import numpy as np
        dt = np.dtype(dtype)
        p = lib().r8bgpu_batch_host_alloc(self._h, int(samples_per_channel), dt.itemsize)
        if not p:
            raise R8bGpuError(_err())
        n = self.n_channels * int(samples_per_channel)
        buf = (C.c_char * (n * dt.itemsize)).from_address(p)
        arr = np.frombuffer(buf, dtype=dt, count=n).reshape(self.n_channels, int(samples_per_channel))
        arr.flags.writeable = True
        _host_allocs[arr.ctypes.data] = p
        return arr

    def clear(self):
        if lib().r8bgpu_batch_clear(self._h) != 0:
            raise R8bGpuError(_err())

    def set_stream(self, cuda_stream_ptr):
        lib().r8bgpu_batch_set_stream(self._h, C.c_void_p(int(cuda_stream_ptr) if cuda_stream_ptr else None))

    def sync(self):
        if lib().r8bgpu_batch_sync(self._h) != 0:
            raise R8bGpuError(_err())

    @property
    def kernel_launches(self):
        return int(lib().r8bgpu_batch_kernel_launches(self._h))

    @property
    def device_bytes(self):
        return int(lib().r8bgpu_batch_device_bytes(self._h))

    def stage_kernels(self):
        """[(kernel_name, n_plan_stages_covered)] per plan stage."""
        out = []
        for i in range(len(self.plan.stages())):
            buf = C.create_string_buffer(64)
            n = lib().r8bgpu_batch_stage_kernel(self._h, i, buf, 64)
            out.append((buf.value.decode(), n))
        return out

    def set_timing(self, enable=True):
        lib().r8bgpu_batch_set_timing(self._h, int(bool(enable)))

    def stage_times(self):
        """[(stage_name, accumulated_ms, launches)] since set_timing(True); synchronises."""
        out = []
        for i, st in enumerate(self.plan.stages()):
            n = C.c_ulonglong(0)
            ms = lib().r8bgpu_batch_stage_time_ms(self._h, i, C.byref(n))
            if ms < 0:
                raise R8bGpuError(_err())
            out.append((st["name"], ms, int(n.value)))
        return out

    def process_ptr(self, d_in, in_stride, l, d_out, out_stride, out_cap):
        """Raw device-pointer call (asynchronous).  Returns samples produced per channel."""
        n = lib().r8bgpu_batch_process(self._h, C.c_void_p(int(d_in) if d_in else None), int(in_stride), int(l),
                                       C.c_void_p(int(d_out) if d_out else None), int(out_stride), int(out_cap))
        if n < 0:
            raise R8bGpuError(_err())
        return n

    def process_host_ptr(self, h_in, in_stride, l, h_out, out_stride, out_cap):
        n = lib().r8bgpu_batch_process_host(self._h, C.c_void_p(int(h_in) if h_in else None), int(in_stride), int(l),
                                            C.c_void_p(int(h_out) if h_out else None), int(out_stride), int(out_cap))
        if n < 0:
            raise R8bGpuError(_err())
        return n

    def process_fmt(self, buf_in, l, buf_out, out_cap, host):
        """Typed buffers (Buffer.make(...)): host=True -> r8bgpu_batch_process_host_fmt, else device."""
        fn = lib().r8bgpu_batch_process_host_fmt if host else lib().r8bgpu_batch_process_fmt
        n = fn(self._h, C.byref(buf_in), int(l), C.byref(buf_out), int(out_cap))
        if n < 0:
            raise R8bGpuError(_err())
        return n

    def process_host_fmt(self, x, out_dtype=None, interleaved=False, in_scale=1.0, out_scale=1.0, fmt=None,
                         out_fmt=None):
        """x: numpy array of int16/int32/float32/float64 samples, planar [n_channels, l] or (interleaved=True)
        [l, n_channels]; packed 24-bit is uint8 [..., 3] with fmt=S24.  Returns the same layout in out_dtype
        (default: the input's) -- the conversions of oneshot<Tin,Tout>() (CDSPResampler.h:592-651)."""
        x = np.ascontiguousarray(x)
        fi = _NP_FORMATS[x.dtype.name] if fmt is None else fmt
        shape = x.shape[:2]
        l, nch = (shape[0], shape[1]) if interleaved else (shape[1], shape[0])
        if nch != self.n_channels:
            raise ValueError("channel count mismatch")
        if out_fmt is None:
            out_fmt = fi if out_dtype is None else _NP_FORMATS[np.dtype(out_dtype).name]
        cap = max(self.plan.max_out_len, 1)
        np_out = {F64: np.float64, F32: np.float32, S16: np.int16, S32: np.int32, S24: np.uint8}[out_fmt]
        tail = (3,) if out_fmt == S24 else ()
        y = np.empty(((cap, nch) if interleaved else (nch, cap)) + tail, dtype=np_out)
        bi = Buffer.make(x.ctypes.data, fi, interleaved, nch if interleaved else l, in_scale)
        bo = Buffer.make(y.ctypes.data, out_fmt, interleaved, nch if interleaved else cap, out_scale)
        n = self.process_fmt(bi, l, bo, cap, host=True)
        return (y[:n] if interleaved else y[:, :n]).copy()

    def process_host(self, x):
        """x: float64 numpy [n_channels, l] (C-contiguous rows).  Returns [n_channels, n_out]."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.ndim != 2 or x.shape[0] != self.n_channels:
            raise ValueError("expected [n_channels, l]")
        cap = self.plan.max_out_len
        y = np.empty((self.n_channels, max(cap, 1)), dtype=np.float64)
        n = self.process_host_ptr(x.ctypes.data, x.shape[1], x.shape[1], y.ctypes.data, y.shape[1], cap)
        return y[:, :n].copy()

    def process(self, x, out=None):
        """x: CUDA float64 torch tensor [n_channels, l]; returns a view [n_channels, n_out] of `out`
        (allocated when None).  Runs on torch's current stream."""
        import torch
        assert x.is_cuda and x.dtype == torch.float64 and x.dim() == 2 and x.shape[0] == self.n_channels
        assert x.stride(1) == 1
        cap = max(self.plan.max_out_len, 1)
        if out is None:
            out = torch.empty((self.n_channels, cap), dtype=torch.float64, device=x.device)
        assert out.stride(1) == 1 and out.shape[1] >= cap
        self.set_stream(torch.cuda.current_stream(x.device).cuda_stream)
        n = self.process_ptr(x.data_ptr(), x.stride(0), x.shape[1], out.data_ptr(), out.stride(0), out.shape[1])
        return out[:, :n]


class ResamplerBatch:
    """Channel-batched counterpart of the loop in example.cpp:30-67 (host numpy in/out)."""

    def __init__(self, n_channels, src_rate, dst_rate, max_in_len, trans_band=2.0, atten=ATTEN_24,
                 device=-1, extfft=0):
        self.plan = Plan(src_rate, dst_rate, max_in_len, trans_band, atten, extfft=extfft)
        self.batch = Batch(self.plan, n_channels, device)

    def process(self, x):
        return self.batch.process_host(x)

    def clear(self):
        self.batch.clear()

    def getMaxOutLen(self, _max_in_len=0):
        return self.plan.max_out_len


class CDSPResampler:
    """Single-stream object with the reference's method names (host buffers, n_channels == 1)."""

    def __init__(self, SrcSampleRate, DstSampleRate, aMaxInLen, ReqTransBand=2.0, ReqAtten=206.91,
                 device=-1, extfft=0):
        self.MaxInLen = int(aMaxInLen)
        self.plan = Plan(SrcSampleRate, DstSampleRate, aMaxInLen, ReqTransBand, ReqAtten, extfft=extfft)
        self.batch = Batch(self.plan, 1, device)

    def process(self, ip):
        """ip: 1-D float64 array of l <= MaxInLen samples; returns the produced samples."""
        ip = np.ascontiguousarray(ip, dtype=np.float64).reshape(1, -1)
        if self.plan.passthrough:
            return ip[0]  # the reference returns the input buffer itself (CDSPResampler.h:563-574)
        return self.batch.process_host(ip)[0]

    def clear(self):
        self.batch.clear()

    def getMaxOutLen(self, _max_in_len=0):
        return self.plan.max_out_len

    def getInLenBeforeOutPos(self, ReqOutPos):
        return self.plan.in_len_before_out_pos(ReqOutPos)

    def getInputRequiredForOutput(self, ReqOutSamples):
        return self.plan.input_required_for_output(ReqOutSamples)

    def getLatency(self):
        return 0

    def getLatencyFrac(self):
        return self.plan.latency_frac()

    def getInLenBeforeOutStart(self, ReqOutPos=0):
        """Feeds single zero samples until the output passes ReqOutPos, then clears
        (CDSPResampler.h:443-464); evaluated on the host scheduler, no kernels run."""
        n = 4096
        while True:
            cs = np.cumsum(self.plan.simulate([1] * n))
            hit = np.nonzero(cs > ReqOutPos)[0]
            if len(hit):
                return int(hit[0])
            n *= 2

    def oneshot(self, ip, oplen):
        """Resample a whole signal: feeds MaxInLen chunks, then zeros, until oplen samples exist
        (CDSPResampler.h:592-651).  Returns a float64 array of oplen samples."""
        ip = np.ascontiguousarray(ip, dtype=np.float64)
        out = np.empty(int(oplen), dtype=np.float64)
        got, pos = 0, 0
        zeros = None
        while got < oplen:
            if pos < len(ip):
                chunk = ip[pos:pos + self.MaxInLen]
                pos += len(chunk)
            else:
                if zeros is None:
                    zeros = np.zeros(self.MaxInLen)
                chunk = zeros
            y = self.process(chunk)
            w = min(len(y), oplen - got)
            out[got:got + w] = y[:w]
            got += w
        self.clear()
        return out


class CDSPResampler16(CDSPResampler):
    def __init__(self, SrcSampleRate, DstSampleRate, aMaxInLen, ReqTransBand=2.0, **kw):
        super().__init__(SrcSampleRate, DstSampleRate, aMaxInLen, ReqTransBand, ATTEN_16, **kw)


class CDSPResampler16IR(CDSPResampler):
    def __init__(self, SrcSampleRate, DstSampleRate, aMaxInLen, ReqTransBand=2.0, **kw):
        super().__init__(SrcSampleRate, DstSampleRate, aMaxInLen, ReqTransBand, ATTEN_16IR, **kw)


class CDSPResampler24(CDSPResampler):
    def __init__(self, SrcSampleRate, DstSampleRate, aMaxInLen, ReqTransBand=2.0, **kw):
        super().__init__(SrcSampleRate, DstSampleRate, aMaxInLen, ReqTransBand, ATTEN_24, **kw)
