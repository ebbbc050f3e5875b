// r8b_design.h -- host-side (fp64, strict IEEE) filter design for the B200 resampler plan.
//
// Fresh implementation of the *design-time* mathematics that the reference performs once
// per resampler object; none of this is on the per-sample hot path.  Every function cites
// the reference location whose numerical result it has to reproduce:
//
//   design_lowpass()    CDSPFIRFilter.h:220-537 (buildLPFilter), CDSPSincFilterGen.h:312-337,
//                       r8bbase.h:666-755 (sine recurrence), :1154-1157 (pow_a), :1192-1212 (I0)
//   design_frac_bank()  CDSPFracInterpolator.h:61-189, :279-341; CDSPSincFilterGen.h:452-552;
//                       r8bbase.h:934-961, :1014-1024
//   select_halfband()   CDSPHBUpsampler.h:47-552 (first tap set whose attenuation >= requested)
//   whole_stepping()    CDSPFracInterpolator.h:609-673
//
// Must be compiled without FMA contraction (-ffp-contract=off): the reference's formulas
// contain cancellations (e.g. its own asinh()) whose value depends on the exact operation order.
#pragma once
#include <vector>

namespace r8bgpu {

struct LowpassDesign {
    int kernel_len = 0;       // K = 2L+1
    int half_len = 0;         // L (the reference's filter "latency")
    int block_len_bits = 0;   // reference BlockLenBits (incl. R8B_EXTFFT), used only for emission timing
    std::vector<double> taps; // h[-L..L] stored at [0..K), scaled so that sum == gain
};

// norm_freq in (0,1], trans_band in percent, atten in dB (positive), gain > 0.
// Returns false when parameters are outside the reference's accepted range.
bool design_lowpass(double norm_freq, double trans_band, double atten, double gain, int extfft,
                    LowpassDesign& out);

struct FracBank {
    int filter_len = 0;   // taps per filter (even)
    int fracs = 0;        // number of fractional positions; rows 0..fracs are addressable
    int order = 0;        // 0: one coefficient per tap (whole stepping), 2: c0,c1,c2 per tap
    double atten = 0.0;   // attenuation after rounding to the table row
    std::vector<double> table; // [(fracs+1)][filter_len][order+1]
};

// init_fracs = -1 selects the automatic count (non-whole stepping, order 2, 8 interpolation points);
// otherwise init_fracs = OutStep (whole stepping, order 0).
void design_frac_bank(int init_fracs, double req_atten, bool is_third, FracBank& out);

struct HalfbandTaps {
    int ntaps = 0;
    double atten = 0.0;
    const double* taps = nullptr; // points into the static table
};
HalfbandTaps select_halfband(double req_atten, int steep_index, bool is_third);

bool whole_stepping(double src_rate, double dst_rate, int& in_step, int& out_step);

int bit_occupancy(int v); // r8bbase.h:766-803: number of significant bits of v (v>=0), 1 for v==0

} // namespace r8bgpu
