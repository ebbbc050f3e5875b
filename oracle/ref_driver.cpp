// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin extern "C" driver over the UNMODIFIED reference headers that live under
// /root/reference (included with -I/root/reference; no reference source is
// copied into this repository).  It is compiled by oracle/Makefile into
// oracle/_ref/libr8bref_e{0,1}[_fast].so and is used
//   * by tests/ as the ground-truth oracle for the CUDA path and for pinning
//     the C restatement in oracle/r8b_oracle.c,
//   * by bench.py (cpu_baseline / --impl reference) as the reference's own CPU
//     implementation of the process() path (R8B_PFFFT_DOUBLE build).
//
// Reference entry points exercised:
//   r8b::CDSPResampler                CDSPResampler.h:117-575
//   r8b::CDSPBlockConvolver           CDSPBlockConvolver.h:62-354
//   r8b::CDSPFracInterpolator         CDSPFracInterpolator.h:707-922
//   r8b::CDSPHBUpsampler/Downsampler  CDSPHBUpsampler.h:573-732, CDSPHBDownsampler.h:47-239
//   r8b::CDSPFIRFilterCache           CDSPFIRFilter.h:598-694
//   r8b::CDSPFracDelayFilterBank      CDSPFracInterpolator.h:61-240

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <chrono>
#include <thread>
#include <vector>
#include <memory>

#include "CDSPResampler.h"

using namespace r8b;

#define R8BREF_API extern "C" __attribute__((visibility("default")))

R8BREF_API int r8bref_extfft() { return R8B_EXTFFT; }
R8BREF_API int r8bref_pffft_double() { return R8B_PFFFT_DOUBLE; }
R8BREF_API int r8bref_fasttiming() { return R8B_FASTTIMING; }

// ---------------------------------------------------------------- full resampler

R8BREF_API void* r8bref_create(double src, double dst, int max_in_len, double tb, double atten)
{
    return new CDSPResampler(src, dst, max_in_len, tb, atten, fprLinearPhase);
}
R8BREF_API void r8bref_delete(void* h) { delete (CDSPResampler*) h; }
R8BREF_API void r8bref_clear(void* h) { ((CDSPResampler*) h)->clear(); }
R8BREF_API int r8bref_max_out_len(void* h) { return ((CDSPResampler*) h)->getMaxOutLen(0); }
R8BREF_API int r8bref_in_len_before_out_pos(void* h, int p)
{
    return ((CDSPResampler*) h)->getInLenBeforeOutPos(p);
}
R8BREF_API int r8bref_input_required_for_output(void* h, int n)
{
    return ((CDSPResampler*) h)->getInputRequiredForOutput(n);
}
R8BREF_API int r8bref_in_len_before_out_start(void* h, int p)
{
    return ((CDSPResampler*) h)->getInLenBeforeOutStart(p);
}
R8BREF_API double r8bref_latency_frac(void* h) { return ((CDSPResampler*) h)->getLatencyFrac(); }

// process(): copies at most out_cap samples of the returned block into out.
R8BREF_API int r8bref_process(void* h, const double* in, int l, double* out, int out_cap)
{
    double* op = nullptr;
    const int n = ((CDSPResampler*) h)->process(const_cast<double*>(in), l, op);
    const int c = n < out_cap ? n : out_cap;
    if (c > 0 && out != nullptr) memcpy(out, op, (size_t) c * sizeof(double));
    return n;
}

// oneshot<double,double>
R8BREF_API void r8bref_oneshot(void* h, const double* in, int inlen, double* out, int outlen)
{
    ((CDSPResampler*) h)->oneshot(in, inlen, out, outlen);
}

// oneshot<Tin,Tout> with the reference's own sample conversions (CDSPResampler.h:592-651): pins the cast
// semantics that the device-side sample formats restate.  Type codes follow r8bgpu_sample_format:
// 0 double, 1 float, 2 int16, 4 int32.  Returns 0, or -1 for an unknown type pair.
template <typename Tin>
static int oneshot_out(CDSPResampler* rs, const Tin* in, int inlen, int tout, void* out, int outlen)
{
    switch (tout) {
    case 0: rs->oneshot(in, inlen, (double*) out, outlen); return 0;
    case 1: rs->oneshot(in, inlen, (float*) out, outlen); return 0;
    case 2: rs->oneshot(in, inlen, (short*) out, outlen); return 0;
    case 4: rs->oneshot(in, inlen, (int*) out, outlen); return 0;
    default: return -1;
    }
}

R8BREF_API int r8bref_oneshot_typed(void* h, int tin, const void* in, int inlen, int tout, void* out, int outlen)
{
    CDSPResampler* rs = (CDSPResampler*) h;
    switch (tin) {
    case 0: return oneshot_out(rs, (const double*) in, inlen, tout, out, outlen);
    case 1: return oneshot_out(rs, (const float*) in, inlen, tout, out, outlen);
    case 2: return oneshot_out(rs, (const short*) in, inlen, tout, out, outlen);
    case 4: return oneshot_out(rs, (const int*) in, inlen, tout, out, outlen);
    default: return -1;
    }
}

// ---------------------------------------------------------------- single stages

struct RefStage {
    CDSPProcessor* proc = nullptr;
    std::vector<double> buf;  // private copy of the input (stages may process in place)
    std::vector<double> outbuf;
    ~RefStage() { delete proc; }
};

R8BREF_API void* r8bref_stage_blockconv(double norm_freq, double tb, double atten, double gain,
                                        int up, int down)
{
    RefStage* s = new RefStage;
    s->proc = new CDSPBlockConvolver(
        CDSPFIRFilterCache::getLPFilter(norm_freq, tb, atten, fprLinearPhase, gain), up, down, 0.0);
    return s;
}
R8BREF_API void* r8bref_stage_frac(double src, double dst, double atten, int is_third)
{
    RefStage* s = new RefStage;
    s->proc = new CDSPFracInterpolator(src, dst, atten, is_third != 0, 0.0);
    return s;
}
R8BREF_API void* r8bref_stage_hbup(double atten, int steep, int is_third)
{
    RefStage* s = new RefStage;
    s->proc = new CDSPHBUpsampler(atten, steep, is_third != 0, 0.0);
    return s;
}
R8BREF_API void* r8bref_stage_hbdown(double atten, int steep, int is_third)
{
    RefStage* s = new RefStage;
    s->proc = new CDSPHBDownsampler(atten, steep, is_third != 0, 0.0);
    return s;
}
R8BREF_API void r8bref_stage_delete(void* h) { delete (RefStage*) h; }
R8BREF_API void r8bref_stage_clear(void* h) { ((RefStage*) h)->proc->clear(); }
R8BREF_API int r8bref_stage_max_out_len(void* h, int l) { return ((RefStage*) h)->proc->getMaxOutLen(l); }
R8BREF_API int r8bref_stage_in_len_before_out_pos(void* h, int p)
{
    return ((RefStage*) h)->proc->getInLenBeforeOutPos(p);
}
R8BREF_API int r8bref_stage_process(void* h, const double* in, int l, double* out, int out_cap)
{
    RefStage* s = (RefStage*) h;
    s->buf.assign(in, in + l);
    const int mo = s->proc->getMaxOutLen(l);
    if ((int) s->outbuf.size() < mo + 16) s->outbuf.resize((size_t) mo + 16);
    double* op = s->outbuf.data();
    const int n = s->proc->process(s->buf.data(), l, op);
    const int c = n < out_cap ? n : out_cap;
    if (c > 0 && out != nullptr) memcpy(out, op, (size_t) c * sizeof(double));
    return n;
}

// ---------------------------------------------------------------- filter-design dumps

// Low-pass kernel: returns kernel length; fills geometry and (optionally) the
// zero-phase spectrum H[k], k=0..B2/2 (B2 = 2<<BlockLenBits), with the FFT's
// inverse-scaling constant divided out so that H[0] == gain.
R8BREF_API int r8bref_lpfilter(double norm_freq, double tb, double atten, double gain,
                               int* block_len_bits, int* latency, double* spectrum, int spectrum_cap)
{
    CDSPFIRFilter& f = CDSPFIRFilterCache::getLPFilter(norm_freq, tb, atten, fprLinearPhase, gain);
    const int bits = f.getBlockLenBits();
    if (block_len_bits) *block_len_bits = bits;
    if (latency) *latency = f.getLatency();
    const int klen = f.getKernelLen();
    if (spectrum != nullptr) {
        const int b2 = 2 << bits;
        CDSPRealFFTKeeper ffto(bits + 1);
        const double inv = 1.0 / ffto->getInvMulConst();
        const double* kb = f.getKernelBlock();
        // "ZP" layout (CDSPRealFFT.h:395-414): kb[0]=DC, kb[1]=Nyquist, kb[2k]=kb[2k+1]=Re H[k].
        for (int k = 0; k <= b2 / 2 && k < spectrum_cap; k++) {
            double v;
            if (k == 0) v = kb[0];
            else if (k == b2 / 2) v = kb[1];
            else v = kb[2 * k];
            spectrum[k] = v * inv;
        }
    }
    f.unref();
    return klen;
}

// Fractional-delay bank: returns FilterLen; table_out receives
// (fracs+1) rows x FilterLen taps x elsize coefficients in NATURAL order
// [row][tap][coef] (SIMD shuffling of the reference table undone).
R8BREF_API int r8bref_fracbank(int init_fracs, int elsize, int interp_points, double atten,
                               int is_third, int* fracs_out, double* table_out, long cap)
{
    CDSPFracDelayFilterBank& fb = CDSPFracDelayFilterBankCache::getFilterBank(
        init_fracs, elsize, interp_points, atten, is_third != 0, false);
    const int flen = fb.getFilterLen();
    const int fracs = fb.getFilterFracs();
    if (fracs_out) *fracs_out = fracs;
    if (table_out != nullptr) {
        long o = 0;
        for (int r = 0; r <= fracs; r++) {
            const double* p = &fb[r];
            for (int t = 0; t < flen; t++) {
                for (int c = 0; c < elsize; c++) {
                    double v;
#if defined(R8B_SIMD_ISH)
                    if (elsize == 1) v = p[t];
                    else {
                        // pairs of taps stored as [a0,b0,a1,b1,...] (CDSPFracInterpolator.h:350-414)
                        const int pair = t >> 1, lane = t & 1;
                        v = p[pair * 2 * elsize + c * 2 + lane];
                    }
#else
                    v = p[t * elsize + c];
#endif
                    if (o < cap) table_out[o] = v;
                    o++;
                }
            }
        }
    }
    fb.unref();
    return flen;
}

R8BREF_API int r8bref_hbfilter(double atten, int steep, int is_third, double* taps, double* att_out)
{
    const double* flt = nullptr;
    int fltt = 0;
    double att = 0.0;
    if (is_third) CDSPHBUpsampler::getHBFilterThird(atten, steep, flt, fltt, att);
    else CDSPHBUpsampler::getHBFilter(atten, steep, flt, fltt, att);
    if (taps) memcpy(taps, flt, (size_t) fltt * sizeof(double));
    if (att_out) *att_out = att;
    return fltt;
}

R8BREF_API int r8bref_whole_stepping(double s, double d, int* in_step, int* out_step)
{
    int a = 0, b = 0;
    const bool ok = getWholeStepping(s, d, a, b);
    if (in_step) *in_step = a;
    if (out_step) *out_step = b;
    return ok ? 1 : 0;
}

// ---------------------------------------------------------------- CPU baseline runner
//
// The reference's usage pattern (example.cpp:30-67): one CDSPResampler per
// channel, every channel fed the same block length per call.  Channels are
// partitioned statically over n_threads std::threads (objects are independent;
// the global caches are mutex-guarded, README.md:52-55).  Input is planar
// [n_ch][block_len]; the same block is re-fed n_calls times (the arithmetic
// does not depend on the data).  Returns seconds spent in process() for the
// timed calls (max over threads); construction is excluded, as in
// bench/r8bfreesrc.cpp:117-127.  *checksum receives a sum of outputs so the
// work cannot be optimised away; *out_samples the outputs per channel.
R8BREF_API double r8bref_bench(double src, double dst, int block_len, double tb, double atten,
                               int n_ch, int n_warm, int n_calls, int n_threads,
                               const double* in, long in_ch_stride,
                               double* checksum, long* out_samples)
{
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_ch) n_threads = n_ch;
    std::vector<std::unique_ptr<CDSPResampler>> rs((size_t) n_ch);
    for (int c = 0; c < n_ch; c++)
        rs[(size_t) c].reset(new CDSPResampler(src, dst, block_len, tb, atten, fprLinearPhase));
    std::vector<double> secs((size_t) n_threads, 0.0), sums((size_t) n_threads, 0.0);
    std::vector<long> outs((size_t) n_threads, 0);
    auto work = [&](int t) {
        const int c0 = (int) ((long) n_ch * t / n_threads);
        const int c1 = (int) ((long) n_ch * (t + 1) / n_threads);
        std::vector<double> tmp((size_t) block_len);
        double s = 0.0;
        long no = 0;
        for (int call = 0; call < n_warm; call++)
            for (int c = c0; c < c1; c++) {
                double* op;
                memcpy(tmp.data(), in + (long) c * in_ch_stride, (size_t) block_len * sizeof(double));
                rs[(size_t) c]->process(tmp.data(), block_len, op);
            }
        const auto t0 = std::chrono::steady_clock::now();
        for (int call = 0; call < n_calls; call++)
            for (int c = c0; c < c1; c++) {
                double* op;
                // process() never writes its input (CDSPResampler.h:541-543) but takes
                // a non-const pointer.
                const int n = rs[(size_t) c]->process(
                    const_cast<double*>(in + (long) c * in_ch_stride), block_len, op);
                if (n > 0) s += op[0] + op[n - 1];
                if (c == c0) no += n;
            }
        const auto t1 = std::chrono::steady_clock::now();
        secs[(size_t) t] = std::chrono::duration<double>(t1 - t0).count();
        sums[(size_t) t] = s;
        outs[(size_t) t] = no;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; t++) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    double mx = 0.0, cs = 0.0;
    for (int t = 0; t < n_threads; t++) {
        if (secs[(size_t) t] > mx) mx = secs[(size_t) t];
        cs += sums[(size_t) t];
    }
    if (checksum) *checksum = cs;
    if (out_samples) *out_samples = outs[0];
    return mx;
}
