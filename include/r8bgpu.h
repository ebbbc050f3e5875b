/* r8bgpu.h -- C-ABI of the B200-native sample-rate-conversion engine (libr8bgpu.so).
 *
 * Drop-in boundary for the r8b::CDSPResampler::process() path.  Plain C: opaque handles,
 * doubles, ints and raw pointers only -- no C++ or torch types cross this boundary.
 *
 * What each entry point replaces in the reference (file:line into avaneev/r8brain-free-src):
 *
 *   r8bgpu_plan_create            CDSPResampler::CDSPResampler()                CDSPResampler.h:117-394
 *                                 (and r8b_create()                              DLL/r8bsrc.cpp:64-88)
 *   r8bgpu_plan_max_out_len       CDSPResampler::getMaxOutLen()                 CDSPResampler.h:502-505
 *   r8bgpu_plan_in_len_before_out_pos   ::getInLenBeforeOutPos()                CDSPResampler.h:406-419
 *                                 (and r8b_inlen()                               DLL/r8bsrc.cpp:95-98)
 *   r8bgpu_plan_input_required_for_output  ::getInputRequiredForOutput()        CDSPResampler.h:476-484
 *   r8bgpu_plan_latency_frac      ::getLatencyFrac()                            CDSPResampler.h:491-494
 *   r8bgpu_batch_create           N x "new CDSPResampler24(...)", one object per channel
 *                                                                                example.cpp:30-41
 *   r8bgpu_batch_clear            CDSPResampler::clear() on every channel       CDSPResampler.h:521-529
 *                                 (and r8b_clear()                               DLL/r8bsrc.cpp:99-100)
 *   r8bgpu_batch_process          the per-channel loop "Resamps[i]->process(in[i], l, op)"
 *                                                                                example.cpp:61-67,
 *                                                                                CDSPResampler.h:559-575
 *                                 (and r8b_process()                             DLL/r8bsrc.cpp:101-105)
 *   r8bgpu_batch_process_fmt / _host_fmt   same, plus the sample conversion loops of
 *                                 oneshot<Tin,Tout>()                            CDSPResampler.h:592-651
 *   r8bgpu_batch_process_host     same, with host buffers (H2D + kernels + D2H); this is what the
 *                                 single-object r8b::CDSPResampler::process() shim in
 *                                 include/r8b/CDSPResampler.h calls.
 *
 * Conventions
 *   - Audio is planar: channel c's samples start at base + c*stride (stride in doubles).
 *   - Every channel of a batch receives the same number of input samples `l` per call and
 *     therefore produces the same number of output samples, which is the return value.
 *   - The reference has no error channel (R8BASSERT compiles out, r8bconf.h:20-29).  Here a
 *     negative return value / NULL handle signals failure; r8bgpu_last_error() (thread-local)
 *     says why.  There is NO CPU fallback: without a usable CUDA device every batch call fails.
 *   - A batch is bound to one CUDA device and one stream; calls on one batch must be serialised
 *     by the caller (same rule as one reference object = one thread at a time, README.md:52-55).
 *     Plans are immutable and may be shared.
 */
#ifndef R8BGPU_H_INCLUDED
#define R8BGPU_H_INCLUDED

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define R8BGPU_API __declspec(dllexport)
#else
#define R8BGPU_API __attribute__((visibility("default")))
#endif

typedef struct r8bgpu_plan r8bgpu_plan;
typedef struct r8bgpu_batch r8bgpu_batch;

/* Stage kinds reported by r8bgpu_plan_stage_info(). */
enum {
    R8BGPU_STAGE_BLOCKCONV = 0,  /* CDSPBlockConvolver  */
    R8BGPU_STAGE_FRAC_WHOLE = 1, /* CDSPFracInterpolator, whole-number stepping */
    R8BGPU_STAGE_FRAC_POLY = 2,  /* CDSPFracInterpolator, 2nd-order interpolated bank */
    R8BGPU_STAGE_HBUP = 3,       /* CDSPHBUpsampler */
    R8BGPU_STAGE_HBDOWN = 4      /* CDSPHBDownsampler */
};

typedef struct r8bgpu_stage_info {
    int kind;
    int up, down;          /* BLOCKCONV */
    int kernel_len;        /* BLOCKCONV: taps; FRAC: taps per filter; HB: one-sided taps */
    int latency;           /* BLOCKCONV: reference Latency (InputLen + L) */
    int ref_input_len;     /* BLOCKCONV: reference InputLen */
    int block_len_bits;    /* BLOCKCONV: reference BlockLenBits */
    int fracs;             /* FRAC: filter-bank fractions */
    int in_step, out_step; /* FRAC_WHOLE */
    int order;             /* FRAC: 0 or 2 */
    int max_out_len;       /* reference getMaxOutLen() chain value after this stage */
    double atten;          /* FRAC/HB: attenuation of the selected table row */
    int data_len;          /* doubles r8bgpu_plan_stage_data() provides */
} r8bgpu_stage_info;

R8BGPU_API const char* r8bgpu_last_error(void);
R8BGPU_API const char* r8bgpu_version(void);

/* ---- plan (host only; needs no GPU) ----------------------------------------------------- */

/* phase: 0 = fprLinearPhase (the only one implemented).  extfft / fasttiming carry the
 * reference's compile-time R8B_EXTFFT / R8B_FASTTIMING (r8bconf.h:132,146), which change the
 * emission timing / interpolation timing of the chain. */
R8BGPU_API r8bgpu_plan* r8bgpu_plan_create(double src_rate, double dst_rate, int max_in_len,
                                           double trans_band, double atten, int phase, int extfft,
                                           int fasttiming);
/* Test hook: a one-stage chain (see r8b_plan.h Plan::build_single). */
R8BGPU_API r8bgpu_plan* r8bgpu_plan_create_stage(int kind, const double* params, int n_params,
                                                 int max_in_len, int extfft);
R8BGPU_API void r8bgpu_plan_destroy(r8bgpu_plan* plan);
R8BGPU_API int r8bgpu_plan_max_out_len(const r8bgpu_plan* plan);
R8BGPU_API int r8bgpu_plan_in_len_before_out_pos(const r8bgpu_plan* plan, int req_out_pos);
R8BGPU_API int r8bgpu_plan_input_required_for_output(const r8bgpu_plan* plan, int req_out_samples);
R8BGPU_API double r8bgpu_plan_latency_frac(const r8bgpu_plan* plan);
R8BGPU_API int r8bgpu_plan_is_passthrough(const r8bgpu_plan* plan);
R8BGPU_API int r8bgpu_plan_stage_count(const r8bgpu_plan* plan);
R8BGPU_API int r8bgpu_plan_stage_info(const r8bgpu_plan* plan, int stage, r8bgpu_stage_info* info);
/* BLOCKCONV: time-domain taps h[-L..L]; FRAC: bank [(fracs+1)][taps][order+1]; HB: taps. */
R8BGPU_API int r8bgpu_plan_stage_data(const r8bgpu_plan* plan, int stage, double* out, int cap);
R8BGPU_API int r8bgpu_plan_describe(const r8bgpu_plan* plan, char* buf, int cap);
/* Dry-run of the integer scheduler: counts[i] = what process() would return for lens[i]. */
R8BGPU_API int r8bgpu_plan_simulate(const r8bgpu_plan* plan, const int* lens, int n_calls, int* counts);

/* ---- batch (GPU) ------------------------------------------------------------------------- */

R8BGPU_API int r8bgpu_device_count(void);
/* device >= 0: that CUDA device.  R8BGPU_DEVICE_ALL (-1): every visible device -- the channels are sharded
 * contiguously, ceil(n/G) per GPU, and each shard is an ordinary single-device batch driven by its own worker thread
 * (bound to the GPU's NUMA node), stream set and PCIe link; there is no device-to-device traffic.  Such a batch takes
 * HOST buffers (r8bgpu_batch_process_host / _host_fmt); device buffers go to the shards (r8bgpu_batch_shard()).  With
 * one visible device, or one channel, this is an ordinary batch.  R8BGPU_DEVICE_CURRENT (-2): the current device.
 * Replaces the caller-side loop over per-channel objects spread over threads (example.cpp:30-67). */
#define R8BGPU_DEVICE_ALL (-1)
#define R8BGPU_DEVICE_CURRENT (-2)
R8BGPU_API r8bgpu_batch* r8bgpu_batch_create(const r8bgpu_plan* plan, int n_channels, int device);
/* Shards of a batch (1 for a single-device batch): device, channel range and NUMA node (-1: unknown / one node). */
R8BGPU_API int r8bgpu_batch_shard_count(const r8bgpu_batch* batch);
R8BGPU_API int r8bgpu_batch_shard_info(const r8bgpu_batch* batch, int shard, int* device, int* first_channel,
                                       int* n_channels, int* numa_node);
/* The single-device batch behind shard `shard` (owned by `batch`; for device-pointer calls on its GPU). */
R8BGPU_API r8bgpu_batch* r8bgpu_batch_shard(r8bgpu_batch* batch, int shard);
/* Page-locked planar host buffer [channels][samples_per_channel] of `sample_bytes`-wide samples whose rows sit on the
 * NUMA node of the GPU that owns the channel (mmap + mbind + cudaHostRegister); free with r8bgpu_host_free(). */
R8BGPU_API void* r8bgpu_batch_host_alloc(const r8bgpu_batch* batch, size_t samples_per_channel, int sample_bytes);
R8BGPU_API void r8bgpu_batch_destroy(r8bgpu_batch* batch);
R8BGPU_API int r8bgpu_batch_clear(r8bgpu_batch* batch);
R8BGPU_API int r8bgpu_batch_channels(const r8bgpu_batch* batch);
/* stream: a cudaStream_t (NULL = the legacy default stream, which is also the default). */
R8BGPU_API int r8bgpu_batch_set_stream(r8bgpu_batch* batch, void* stream);

/* Device-pointer call; asynchronous on the batch's stream.  l <= MaxInLen.  Output sample i of
 * channel c lands in d_out[c*out_ch_stride + i]; out_cap is the room per channel (use
 * r8bgpu_plan_max_out_len()).  Returns samples produced per channel, or < 0. */
R8BGPU_API int r8bgpu_batch_process(r8bgpu_batch* batch, const double* d_in, size_t in_ch_stride, int l,
                                    double* d_out, size_t out_ch_stride, int out_cap);
/* Host-pointer call: copies in, runs, copies the produced samples out, synchronises. */
R8BGPU_API int r8bgpu_batch_process_host(r8bgpu_batch* batch, const double* h_in, size_t in_ch_stride,
                                         int l, double* h_out, size_t out_ch_stride, int out_cap);
R8BGPU_API int r8bgpu_batch_sync(r8bgpu_batch* batch);

/* ---- caller-side sample formats ---------------------------------------------------------
 * What the reference's callers do on the CPU around process(): CDSPResampler::oneshot<Tin,Tout>()
 * converts "(double) ip[i]" on the way in and "(Tout) op[i]" on the way out (CDSPResampler.h:592-651),
 * and WAV front-ends de-interleave frames (bench/r8bfreesrc.cpp:106-137).  Here the narrow samples
 * cross PCIe / HBM as they are and are widened / narrowed on the device.
 *   in : x = (double) v * scale        out: v = (T) (y * scale)
 * With scale = 1 these are exactly the C++ conversions of oneshot(): widening is exact, float output
 * rounds to nearest, integer output truncates toward zero (out-of-range values, undefined in the
 * reference, saturate; NaN -> 0).  R8BGPU_S24 is packed 3-byte little-endian. */
typedef enum {
    R8BGPU_F64 = 0,
    R8BGPU_F32 = 1,
    R8BGPU_S16 = 2,
    R8BGPU_S24 = 3,
    R8BGPU_S32 = 4
} r8bgpu_sample_format;

typedef struct {
    void* data;      /* host (…_host_fmt) or device (…_fmt) memory; never written when used as input */
    int format;      /* r8bgpu_sample_format */
    int interleaved; /* 0: planar, channel c starts at c*stride; 1: frame f starts at f*stride, channel c at +c */
    size_t stride;   /* in samples of `format` */
    double scale;    /* see above; 1.0 for the reference's plain casts */
} r8bgpu_buffer;

/* As r8bgpu_batch_process() / r8bgpu_batch_process_host() with typed buffers.  out_cap = room per
 * channel in samples. */
R8BGPU_API int r8bgpu_batch_process_fmt(r8bgpu_batch* batch, const r8bgpu_buffer* d_in, int l,
                                        const r8bgpu_buffer* d_out, int out_cap);
R8BGPU_API int r8bgpu_batch_process_host_fmt(r8bgpu_batch* batch, const r8bgpu_buffer* h_in, int l,
                                             const r8bgpu_buffer* h_out, int out_cap);

/* Number of kernels this batch has launched since creation. */
R8BGPU_API unsigned long long r8bgpu_batch_kernel_launches(const r8bgpu_batch* batch);
/* Per-stage device timing for profiling/bench: when enabled every stage launch is bracketed
 * by CUDA events on the batch's stream.  r8bgpu_batch_stage_time_ms() synchronises and returns the
 * accumulated milliseconds (and launch count) of one stage since timing was (re-)enabled. */
R8BGPU_API int r8bgpu_batch_set_timing(r8bgpu_batch* batch, int enable);
R8BGPU_API double r8bgpu_batch_stage_time_ms(r8bgpu_batch* batch, int stage, unsigned long long* launches);
/* Name of the kernel that executes plan stage `stage`; returns the number of consecutive plan stages
 * that kernel covers (0: the stage is folded into an earlier stage's kernel), < 0 on error. */
R8BGPU_API int r8bgpu_batch_stage_kernel(const r8bgpu_batch* batch, int stage, char* name, int cap);
/* Bytes of device memory held by the batch (state rings + tables + staging). */
R8BGPU_API unsigned long long r8bgpu_batch_device_bytes(const r8bgpu_batch* batch);

/* Calibration for the bench's secondary (fp64) roofline: measured DFMA throughput of `device` (< 0: current) in
 * TFLOP/s, register-resident, ~10 ms; < 0 on error.  Replaces nothing in the reference. */
R8BGPU_API double r8bgpu_measure_fp64_tflops(int device);

/* Page-locked host memory for the *_host entry points (optional but faster). */
R8BGPU_API void* r8bgpu_host_alloc(size_t bytes);
R8BGPU_API void r8bgpu_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* R8BGPU_H_INCLUDED */
