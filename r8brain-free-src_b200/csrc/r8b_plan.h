// r8b_plan.h -- host-side resampling plan and integer "shadow scheduler".
//
// A Plan is the B200 engine's restatement of what r8b::CDSPResampler's constructor decides
// (CDSPResampler.h:117-394): which stages form the chain for (SrcSampleRate, DstSampleRate),
// their filters, and the latency bookkeeping.  Stages are expressed as operators on
// ABSOLUTELY INDEXED streams (sample n of the stage input since clear()); how many samples
// each stage has emitted after N inputs is a closed-form integer function, identical for every
// channel of a batch, so one Schedule instance drives all channels.
//
//   BlockConv (U,D)   z[q] = sum_k h[k] * xu[D*q - k],  xu[t] = x[t/U] if U|t else 0
//                     emitted(N) = max(0, ceil((U*N - Latency)/D)),  Latency = InputLen + L
//                     (CDSPBlockConvolver.h:62-185, 252-354, 512-593)
//   FracWhole         out[j] = sum_i bank[(j*InStep) % OutStep][i] * x[(j*InStep)/OutStep - fll + i]
//                     produced while  p_j + fl2 <= N-1            (CDSPFracInterpolator.h:991-1060)
//   FracPoly          order-2 interpolated bank with the resettable-counter timing
//                     (CDSPFracInterpolator.h:1069-1179, 907-919)
//   HBUp (T taps)     out[2n] = x[n]; out[2n+1] = sum_k f[k](x[n-k] + x[n+1+k]); emitted = 2*max(0,N-T)
//                     (CDSPHBUpsampler.h:674-732)
//   HBDown (T taps)   out[m] = x[2m] + sum_k f[k](x[2m+1+2k] + x[2m-1-2k]); emitted = max(0,N/2-(T-1))
//                     (CDSPHBDownsampler.h:137-239)
//
// Only the linear-phase presets are planned (fprMinPhase is out of scope).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "r8b_design.h"

namespace r8bgpu {

enum StageKind { ST_BLOCKCONV = 0, ST_FRAC_WHOLE = 1, ST_FRAC_POLY = 2, ST_HBUP = 3, ST_HBDOWN = 4 };

struct StageDesc {
    StageKind kind;
    // --- BlockConv
    int up = 1, down = 1;
    int ref_input_len = 0;   // the reference's InputLen (emission timing only)
    int latency = 0;         // the reference's Latency = InputLen + L
    int ref_prev_len = 0;    // the reference's PrevInputLen (overlap carried between its blocks)
    bool block_exact = false; // power-of-two D: the reference inverse-transforms only the lower
                             // 1/D of each block spectrum (CDSPBlockConvolver.h:329-344), which is
                             // NOT plain decimation; such stages reproduce the reference's block
                             // segmentation (FFT size 2<<BlockLenBits, blocks every InputLen).
    double norm_freq = 0, trans_band = 0, gain = 0;
    LowpassDesign lp;
    // --- Frac
    double src_rate = 0, dst_rate = 0; // as seen by this stage
    bool is_third = false;
    int in_step = 0, out_step = 0;
    bool fasttiming = false; // R8B_FASTTIMING: drifting accumulator instead of the resettable counter
    FracBank bank;
    // --- Halfband
    int hb_taps = 0;
    int steep_index = 0;
    double hb_atten = 0;
    std::vector<double> hb;
    // --- derived
    int max_out_len = 0;     // reference getMaxOutLen chain value after this stage
    int src_history = 0;     // how many source samples before "inputs so far" may be re-read
};

struct Plan {
    double src_rate = 0, dst_rate = 0;
    int max_in_len = 0;
    double trans_band = 2.0, atten = 0;
    int extfft = 0;
    bool passthrough = false; // SrcSampleRate == DstSampleRate (CDSPResampler.h:135-138)
    std::vector<StageDesc> stages;
    int max_out_len = 0;      // CurMaxOutLen (CDSPResampler.h:502-505)
    std::string error;

    // Returns false (and sets error) for configurations this engine does not plan.
    bool build(double src, double dst, int max_in_len, double tb, double atten, int phase, int extfft,
               int fasttiming);

    // Test hook: a chain consisting of ONE stage, so that each kernel can be checked against the
    // corresponding reference stage class in isolation.  kind: StageKind; a[]: BLOCKCONV
    // {norm_freq, trans_band, atten, gain, up, down}; FRAC_* {src, dst, atten, is_third};
    // HBUP/HBDOWN {atten, steep_index, is_third}.
    bool build_single(int kind, const double* a, int max_in_len, int extfft);

    int in_len_before_out_pos(int req_out_pos) const; // CDSPResampler.h:406-419
    int input_required_for_output(int n) const;       // :476-484
    std::string describe() const;                      // R8BCONSOLE-style plan dump
};

// Range of absolute output indices [e0,e1) a stage emits during one process() call and the
// number of source samples [n0,n1) that became available to it.
struct StageCall {
    long long n0 = 0, n1 = 0;
    long long e0 = 0, e1 = 0;
    // FracPoly only: timing state at the first output of this call.
    int in_counter0 = 0, in_pos_int0 = 0;
    double in_pos_shift = 0.0, fpos0 = 0.0;
    long long p0 = 0;
    long long p_last = 0; // read position of the last output of this call (FracPoly)
    // R8B_FASTTIMING: the position sequence is inherently sequential (fpos += step with rounding), so the
    // host walks it and hands the kernel one (position - p0, fraction) pair per output of the call
    std::vector<int> ft_dp;
    std::vector<double> ft_fpos;
};

struct Schedule {
    const Plan* plan = nullptr;
    std::vector<long long> n_in, n_out; // per stage totals since clear()
    // FracPoly timing state (at most one such stage per chain, but keep per stage)
    struct PolyState {
        int in_counter = 0, in_pos_int = 0;
        double in_pos_shift = 0.0, fpos = 0.0;
        long long p = 0;
    };
    std::vector<PolyState> poly;

    void init(const Plan* p);
    void clear();
    // Advance by l input samples; fills one StageCall per stage; returns samples emitted by the chain.
    int advance(int l, std::vector<StageCall>& calls);
};

// emitted-sample count helpers (exposed for tests)
long long blockconv_emitted(const StageDesc& s, long long n_in);
long long frac_whole_emitted(const StageDesc& s, long long n_in);
long long hbup_emitted(const StageDesc& s, long long n_in);
long long hbdown_emitted(const StageDesc& s, long long n_in);

} // namespace r8bgpu
