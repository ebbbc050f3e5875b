// include/r8b/CDSPResampler.h -- header-style C++ front-end in namespace r8b over the r8bgpu C-ABI.
//
// Drop-in for the reference's CDSPResampler.h on the process() path: same class names,
// constructor arguments, method names and ownership rules (CDSPResampler.h:117-651,729-810 of
// avaneev/r8brain-free-src); the per-channel CPU pipeline behind them is replaced by sm_100a
// kernels reached through include/r8bgpu.h.  Link with -lr8bgpu.
//
//   r8b::CDSPResampler / CDSPResampler16 / CDSPResampler16IR / CDSPResampler24
//        one stream per object, HOST buffers; process() = H2D + kernels + D2H per call.
//        Meant for drop-in correctness; it cannot be fast (one PCIe round trip per call).
//   r8b::CDSPResamplerBatch (new)
//        N independent channels processed in lock-step -- the shape of example.cpp:30-67 --
//        with host OR device planar buffers.  This is the intended production entry.
//
// Configuration macros (r8bconf.h surface).  The FFT back-end selectors R8B_IPP, R8B_PFFFT,
// R8B_PFFFT_DOUBLE, R8B_FLOATFFT are accepted and ignored (there is no CPU FFT and no CPU
// fallback).  R8B_EXTFFT and R8B_FASTTIMING change the plan exactly as they change the reference
// (block length -> emission latency; interpolator timing) and are forwarded to the planner.
// R8BASSERT / R8BCONSOLE keep their meaning (no-ops unless defined by the user).
#ifndef R8B_CDSPRESAMPLER_B200_INCLUDED
#define R8B_CDSPRESAMPLER_B200_INCLUDED

#include <cstddef>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <vector>

#include "../r8bgpu.h"

#ifndef R8B_EXTFFT
#define R8B_EXTFFT 0
#endif
#ifndef R8B_FASTTIMING
#define R8B_FASTTIMING 0
#endif
#ifndef R8BASSERT
#define R8BASSERT(e)
#endif
#ifndef R8BCONSOLE
#define R8BCONSOLE(...)
#endif

namespace r8b {

/// Filter phase response (CDSPFIRFilter.h:34-46).  Only fprLinearPhase is implemented.
enum EDSPFilterPhaseResponse { fprLinearPhase = 0, fprMinPhase = 1 };

/// N channels resampled in lock-step on one GPU.
class CDSPResamplerBatch {
public:
    CDSPResamplerBatch(const int NumChannels, const double SrcSampleRate, const double DstSampleRate,
                       const int aMaxInLen, const double ReqTransBand = 2.0, const double ReqAtten = 206.91,
                       const EDSPFilterPhaseResponse ReqPhase = fprLinearPhase, const int Device = -1)
        : Plan(r8bgpu_plan_create(SrcSampleRate, DstSampleRate, aMaxInLen, ReqTransBand, ReqAtten,
                                  (int) ReqPhase, R8B_EXTFFT, R8B_FASTTIMING))
        , Batch(NULL)
        , Channels(NumChannels)
        , Dev(Device)
        , MaxInLen(aMaxInLen)
    {
        R8BASSERT(Plan != NULL);
        if (Plan != NULL) {
            char buf[1024];
            r8bgpu_plan_describe(Plan, buf, (int) sizeof(buf));
            R8BCONSOLE("%s", buf);
            (void) buf;
        }
    }

    ~CDSPResamplerBatch()
    {
        if (Batch != NULL) r8bgpu_batch_destroy(Batch);
        if (Plan != NULL) r8bgpu_plan_destroy(Plan);
    }

    /// false: the constructor arguments were refused (out-of-range parameters, fprMinPhase ...) or, after the first
    /// call, no CUDA device / no memory for the batch.  The reference has no such state (R8BASSERT compiles out and
    /// bad parameters are undefined behaviour); here every accessor of an invalid object returns 0 and every
    /// process() returns -1, and getLastError() says why.
    bool isValid() const { return Plan != NULL && !Failed; }
    const char* getLastError() const { return r8bgpu_last_error(); }
    int getNumChannels() const { return Channels; }
    int getMaxOutLen(const int /* MaxInLen */ = 0) const { return Plan ? r8bgpu_plan_max_out_len(Plan) : 0; }
    int getInLenBeforeOutPos(const int ReqOutPos) const { return Plan ? r8bgpu_plan_in_len_before_out_pos(Plan, ReqOutPos) : 0; }
    int getInputRequiredForOutput(const int ReqOutSamples) const { return Plan ? r8bgpu_plan_input_required_for_output(Plan, ReqOutSamples) : 0; }
    int getLatency() const { return 0; }
    double getLatencyFrac() const { return Plan ? r8bgpu_plan_latency_frac(Plan) : 0.0; }

    void clear()
    {
        if (Batch != NULL) r8bgpu_batch_clear(Batch);
    }

    /// Host planar buffers: channel c at ip + c*InStride / op + c*OutStride (strides in doubles).
    /// Returns samples written per channel (same for all channels), or -1.
    int process(const double* ip, const size_t InStride, const int l, double* op, const size_t OutStride,
                const int OutCap)
    {
        if (!ensure()) return -1;
        return r8bgpu_batch_process_host(Batch, ip, InStride, l, op, OutStride, OutCap);
    }

    /// Device planar buffers; asynchronous on the batch stream (see r8bgpu_batch_set_stream()).
    int processDevice(const double* d_ip, const size_t InStride, const int l, double* d_op, const size_t OutStride,
                      const int OutCap)
    {
        if (!ensure()) return -1;
        return r8bgpu_batch_process(Batch, d_ip, InStride, l, d_op, OutStride, OutCap);
    }

    /// Typed buffers (int16 / packed int24 / int32 / float32 / float64, planar or interleaved): the
    /// sample conversions of oneshot<Tin,Tout>() (CDSPResampler.h:592-651) run on the device.
    int process(const r8bgpu_buffer& ip, const int l, const r8bgpu_buffer& op, const int OutCap)
    {
        if (!ensure()) return -1;
        return r8bgpu_batch_process_host_fmt(Batch, &ip, l, &op, OutCap);
    }

    int processDevice(const r8bgpu_buffer& d_ip, const int l, const r8bgpu_buffer& d_op, const int OutCap)
    {
        if (!ensure()) return -1;
        return r8bgpu_batch_process_fmt(Batch, &d_ip, l, &d_op, OutCap);
    }

    void setStream(void* CudaStream)
    {
        if (ensure()) r8bgpu_batch_set_stream(Batch, CudaStream);
    }

    void sync()
    {
        if (Batch != NULL) r8bgpu_batch_sync(Batch);
    }

    r8bgpu_batch* handle()
    {
        ensure();
        return Batch;
    }

private:
    r8bgpu_plan* Plan;
    r8bgpu_batch* Batch;
    int Channels;
    int Dev;
    int MaxInLen;
    bool Failed = false; // batch creation was tried and refused: do not retry on every call

    bool ensure()
    {
        if (Batch == NULL && Plan != NULL && !Failed) {
            Batch = r8bgpu_batch_create(Plan, Channels, Dev);
            Failed = (Batch == NULL);
        }
        R8BASSERT(Batch != NULL);
        return Batch != NULL;
    }

    CDSPResamplerBatch(const CDSPResamplerBatch&);
    CDSPResamplerBatch& operator=(const CDSPResamplerBatch&);
};

/// "Pull" use of a batch for real-time callers (README.md:132-146 of the reference: "a pull method ... calls the
/// resampling process until the output buffer is filled", keeping the excess output for the next request).  The
/// reference leaves that loop to the caller; this helper is the same loop around CDSPResamplerBatch with host buffers:
/// pull() asks `fill` for input blocks of at most MaxInLen frames per channel until `n` output frames per channel are
/// available, hands them out, and keeps what is left over.  Output is exactly the push-mode stream, in the same order.
class CDSPResamplerPull {
public:
    CDSPResamplerPull(const int NumChannels, const double SrcSampleRate, const double DstSampleRate, const int aMaxInLen,
                      const double ReqTransBand = 2.0, const double ReqAtten = 206.91, const int Device = -1)
        : Rs(NumChannels, SrcSampleRate, DstSampleRate, aMaxInLen, ReqTransBand, ReqAtten, fprLinearPhase, Device)
        , Channels(NumChannels)
        , MaxInLen(aMaxInLen)
        , Cap(Rs.getMaxOutLen() > 0 ? Rs.getMaxOutLen() : 1)
        , Have(0)
        , In((size_t) NumChannels * (size_t) aMaxInLen)
        , Out((size_t) NumChannels * (size_t) Cap)
        , Fifo((size_t) NumChannels)
    {
    }

    bool isValid() const { return Rs.isValid(); }
    CDSPResamplerBatch& batch() { return Rs; }

    /// fill(double* ip, size_t stride, int maxFrames) -> frames written per channel (planar, channel c at ip + c*stride);
    /// returning 0 ends the stream (pull() then returns fewer than n frames).  op: planar, channel c at op + c*OutStride.
    template <typename Fill>
    int pull(Fill fill, double* op, const size_t OutStride, const int n)
    {
        while (Have < n) {
            const int l = fill(&In[0], (size_t) MaxInLen, MaxInLen);
            if (l <= 0) break;
            const int got = Rs.process(&In[0], (size_t) MaxInLen, l, &Out[0], (size_t) Cap, Cap);
            if (got < 0) return -1;
            for (int c = 0; c < Channels; c++)
                Fifo[(size_t) c].insert(Fifo[(size_t) c].end(), Out.begin() + (size_t) c * Cap, Out.begin() + (size_t) c * Cap + got);
            Have += got;
        }
        const int give = Have < n ? Have : n;
        for (int c = 0; c < Channels; c++) {
            std::vector<double>& f = Fifo[(size_t) c];
            std::copy(f.begin(), f.begin() + give, op + (size_t) c * OutStride);
            f.erase(f.begin(), f.begin() + give);
        }
        Have -= give;
        return give;
    }

    void clear()
    {
        Rs.clear();
        for (size_t c = 0; c < Fifo.size(); c++) Fifo[c].clear();
        Have = 0;
    }

private:
    CDSPResamplerBatch Rs;
    int Channels, MaxInLen, Cap, Have;
    std::vector<double> In, Out;
    std::vector<std::vector<double> > Fifo;
};

/// Single-stream object with the reference's exact call shape.
class CDSPResampler {
public:
    CDSPResampler(const double SrcSampleRate, const double DstSampleRate, const int aMaxInLen,
                  const double ReqTransBand = 2.0, const double ReqAtten = 206.91,
                  const EDSPFilterPhaseResponse ReqPhase = fprLinearPhase)
        : Impl(1, SrcSampleRate, DstSampleRate, aMaxInLen, ReqTransBand, ReqAtten, ReqPhase)
        , MaxInLen(aMaxInLen)
        , IsSame(SrcSampleRate == DstSampleRate)
    {
        R8BASSERT(SrcSampleRate > 0.0);
        R8BASSERT(DstSampleRate > 0.0);
        R8BASSERT(aMaxInLen > 0);
        OutBuf.resize((size_t) (Impl.getMaxOutLen() > 0 ? Impl.getMaxOutLen() : 1));
    }

    virtual ~CDSPResampler() {}

    /// Not in the reference: see CDSPResamplerBatch::isValid().  An invalid object produces no output: process()
    /// returns 0 samples, oneshot() zero-fills, getInLenBeforeOutStart() returns 0 -- none of them loops.
    bool isValid() const { return IsSame || (Impl.isValid() && !Broken); }
    const char* getLastError() const { return Impl.getLastError(); }

    virtual int getInLenBeforeOutPos(const int ReqOutPos) const { return Impl.getInLenBeforeOutPos(ReqOutPos); }
    int getInputRequiredForOutput(const int ReqOutSamples) const { return Impl.getInputRequiredForOutput(ReqOutSamples); }
    virtual int getLatency() const { return 0; }
    virtual double getLatencyFrac() const { return Impl.getLatencyFrac(); }
    virtual int getMaxOutLen(const int /* MaxInLen */) const { return Impl.getMaxOutLen(); }
    virtual void clear() { Impl.clear(); }

    /// As CDSPResampler.h:443-464: feeds single samples until the output passes ReqOutPos.
    int getInLenBeforeOutStart(const int ReqOutPos = 0)
    {
        int inc = 0, outc = 0;
        while (true) {
            double ins = 0.0;
            double* op;
            outc += process(&ins, 1, op);
            if (!isValid()) return 0; // a failed object never produces output: do not spin
            if (outc > ReqOutPos) {
                clear();
                return inc;
            }
            inc++;
        }
    }

    /// As CDSPResampler.h:559-575: `op0` receives a pointer to an internal buffer that stays valid
    /// until the next call; the input is never written; equal rates hand the input back.
    virtual int process(double* ip0, int l, double*& op0)
    {
        R8BASSERT(l >= 0);
        if (IsSame) {
            op0 = ip0;
            return l;
        }
        op0 = &OutBuf[0];
        const int n = Impl.process(ip0, (size_t) l, l, op0, OutBuf.size(), (int) OutBuf.size());
        R8BASSERT(n >= 0);
        if (n < 0) Broken = true; // sticky: isValid() turns false, the loops below stop
        return n < 0 ? 0 : n;
    }

    /// One-shot conversion of a whole signal (semantics of CDSPResampler.h:592-651): the input is fed in
    /// MaxInLen-sized pieces, followed by silence, until `oplen` output samples have been collected; the
    /// stream state is cleared afterwards.  Tin/Tout may be any arithmetic sample type.
    template <typename Tin, typename Tout>
    void oneshot(const Tin* ip, int iplen, Tout* op, int oplen)
    {
        std::vector<double> chunk((size_t) MaxInLen, 0.0);
        int fed = 0, got = 0;
        while (got < oplen) {
            const int n = (fed < iplen) ? ((iplen - fed < MaxInLen) ? iplen - fed : MaxInLen) : MaxInLen;
            if (fed < iplen) {
                for (int i = 0; i < n; i++) chunk[(size_t) i] = (double) ip[fed + i];
                fed += n;
                if (fed >= iplen && n < MaxInLen) std::fill(chunk.begin() + n, chunk.end(), 0.0);
            } else if (fed == iplen) {
                std::fill(chunk.begin(), chunk.end(), 0.0); // from here on: silence
                fed++;
            }
            double* res = NULL;
            int produced = process(&chunk[0], n, res);
            if (!isValid()) { // no device, refused parameters, launch failure: silence instead of an endless loop
                for (int i = got; i < oplen; i++) op[i] = (Tout) 0;
                return;
            }
            if (produced > oplen - got) produced = oplen - got;
            for (int i = 0; i < produced; i++) op[got + i] = (Tout) res[i];
            got += produced;
        }
        clear();
    }

private:
    CDSPResamplerBatch Impl;
    std::vector<double> OutBuf;
    int MaxInLen;
    bool IsSame;
    bool Broken = false;
};

class CDSPResampler16 : public CDSPResampler {
public:
    CDSPResampler16(const double SrcSampleRate, const double DstSampleRate, const int aMaxInLen,
                    const double ReqTransBand = 2.0)
        : CDSPResampler(SrcSampleRate, DstSampleRate, aMaxInLen, ReqTransBand, 136.45, fprLinearPhase) {}
};

class CDSPResampler16IR : public CDSPResampler {
public:
    CDSPResampler16IR(const double SrcSampleRate, const double DstSampleRate, const int aMaxInLen,
                      const double ReqTransBand = 2.0)
        : CDSPResampler(SrcSampleRate, DstSampleRate, aMaxInLen, ReqTransBand, 109.56, fprLinearPhase) {}
};

class CDSPResampler24 : public CDSPResampler {
public:
    CDSPResampler24(const double SrcSampleRate, const double DstSampleRate, const int aMaxInLen,
                    const double ReqTransBand = 2.0)
        : CDSPResampler(SrcSampleRate, DstSampleRate, aMaxInLen, ReqTransBand, 180.15, fprLinearPhase) {}
};

} // namespace r8b

#endif // R8B_CDSPRESAMPLER_B200_INCLUDED
