// r8b_fused_common.cuh -- pieces shared by the two fused "2x BlockConvolver -> FracInterpolator" kernels
// (r8b_fused.cu: tile pairs, one CTA per pair; r8b_fused2.cu: persistent CTA, two half-CTA pipelines).
// Everything here is plain arithmetic on pointers, so it also compiles for the host: the CPU emulation of
// the v2 kernel (tests/cpp/fused2_emul.cu) runs these very functions thread by thread.
#pragma once
#include "r8b_fft.cuh"
#include "r8b_kernels.h"

namespace r8bgpu {

R8B_HD double src_read_f(const SrcView& v, int ch, long long n)
{
    if (n >= v.avail) return 0.0;
    if (n >= v.cur_base) return R8B_LDG(v.cur + (long long) ch * v.cur_stride + (n - v.cur_base));
    return R8B_LDG(v.ring + (long long) ch * v.ring_stride + (n & v.ring_mask));
}

R8B_HD void dst_write_f(const DstView& v, int ch, long long idx, double x)
{
    v.ptr[(long long) ch * v.stride + ((idx - v.base) & v.mask)] = x;
}

constexpr int FM = 4096;            // FFT length of the fused kernels
constexpr int FPL = fft_padded_len(FM);

R8B_HD int ylay(int i, int ysh) { return i + (i >> ysh); }

// Twiddles from shared memory, laid out [q][r] so that the 16 consecutive lanes of a half-warp read 16
// consecutive entries (the natural [r*q] indexing is an up-to-16-way bank conflict for even q):
//   tw2t[q*16 + r] = W_256^(r q)            (r, q < 16)  -- passes with NCUR = 256
//   tw1t[q*16 + r] = W_M^(r q)              (r, q < 16)
//   W_M^(R q), R = 16 r_hi + r_lo < 256  =  tw2t[q*16 + r_hi] * tw1t[q*16 + r_lo]   -- passes with NCUR = M
// (one extra complex multiply, <= ~1.5 ulp, instead of walking a 64 KB table through L1/L2).
R8B_HD double2 tw_pair(const double2* __restrict__ tw2t, const double2* __restrict__ tw1t, int r, int q)
{
    const double2 c = tw2t[q * 16 + (r >> 4)], f = tw1t[q * 16 + (r & 15)];
    return make_double2(fma(c.x, f.x, -c.y * f.y), fma(c.x, f.y, c.y * f.x));
}

// One radix-16 DIF pass over blocks of NCUR points in padded shared memory; butterfly g of M/16.
template <int NCUR>
R8B_HD void fwd_pass(double2* __restrict__ s, const double2* __restrict__ tw2_s, int g)
{
    constexpr int D = NCUR / 16;
    const int blk = g / D, r = g % D;
    const int base = blk * NCUR + r;
    double2 v[16];
#pragma unroll
    for (int j = 0; j < 16; j++) v[j] = s[fft_pad(base + j * D)];
    Network<16, +1>::run(v);
#pragma unroll
    for (int q = 0; q < 16; q++) {
        double2 x = v[bitrev<16>(q)];
        if (D > 1 && q > 0) x = cmul<+1>(x, tw2_s[q * 16 + r]); // NCUR == 256: W_256^(r q), [q][r] layout
        s[fft_pad(base + q * D)] = x;
    }
}

template <int NCUR>
R8B_HD void inv_pass(double2* __restrict__ s, const double2* __restrict__ tw2_s, int g)
{
    constexpr int D = NCUR / 16;
    const int blk = g / D, r = g % D;
    const int base = blk * NCUR + r;
    double2 v[16];
#pragma unroll
    for (int q = 0; q < 16; q++) {
        double2 x = s[fft_pad(base + q * D)];
        if (D > 1 && q > 0) x = cmul<-1>(x, tw2_s[q * 16 + r]);
        v[q] = x;
    }
    Network<16, -1>::run(v);
#pragma unroll
    for (int j = 0; j < 16; j++) s[fft_pad(base + j * D)] = v[bitrev<16>(j)];
}

} // namespace r8bgpu
