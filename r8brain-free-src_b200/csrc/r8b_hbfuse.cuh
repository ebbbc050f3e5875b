// r8b_hbfuse.cuh -- the LAST TWO stages of a half-band 2x upsampler cascade as one operator
// (CDSPHBUpsampler::process twice in a row, CDSPHBUpsampler.h:674-732):
//
//   u = s_{c-2}  --f (T1 taps)-->  v = s_{c-1}  --g (T2 taps)-->  y = s_c
//   v[2n] = u[n]      v[2n+1] = sum_j f[j] (u[n-j] + u[n+1+j])
//   y[2m] = v[m]      y[2m+1] = sum_j g[j] (v[m-j] + v[m+1+j])        stream values at negative indices are 0
//
// One work item = four consecutive v positions m .. m+3 (m even) -> eight consecutive y values.  The v
// window it needs (2*T2 + 3 values) is produced in registers from a window of T2 + 1 + 2*T1 samples of u;
// v never exists in shared memory.  Compared with running the two stages separately this removes the
// largest intermediate buffer of the cascade (half of its shared memory), its store + reload, and one
// block-wide barrier; the price is that odd v samples near item boundaries are computed twice.
// Summation order of every FIR output is the one of the single-stage kernel (tap 0 first, fma ascending).
//
// The function is plain C++ over an accessor so that the index algebra can be unit-tested on the host
// (tests/cpp/hbfuse_check.cpp).
#pragma once

#if defined(__CUDACC__)
#define R8B_HD __host__ __device__ __forceinline__
#define R8B_HDC __host__ __device__ constexpr
#else
#define R8B_HD inline
#define R8B_HDC constexpr
#endif

namespace r8bgpu {

R8B_HDC int hbf_floor_half(int e) { return e >= 0 ? e / 2 : -((-e + 1) / 2); }

// Number of u samples one item reads, and the offset of its first one relative to h = m / 2.
template <int T1, int T2>
struct HbFuseGeom {
    static constexpr int NA = hbf_floor_half(1 - T2);     // first n (relative to h) whose v pair is touched
    static constexpr int UB = NA - T1 + 1;                // u window starts at h + UB
    static constexpr int UW = T2 + 1 + 2 * T1;            // u window length
    static constexpr int VW = 2 * T2 + 3;                 // v window length, starts at m - T2 + 1
};

// U: callable (int s) -> double returning u[h + UB + s], 0 <= s < UW.
// m: absolute index of the first v position (even); v indices < 0 are zeros when may_be_negative.
// y8[0..7] receives y[2m .. 2m+7].
template <int T1, int T2, typename U>
R8B_HD void hb_fused_item(const double (&f)[T1], const double (&g)[T2], U u, long long m, bool may_be_negative,
                          double (&y8)[8])
{
    using G = HbFuseGeom<T1, T2>;
    double uw[G::UW];
#pragma unroll
    for (int s = 0; s < G::UW; s++) uw[s] = u(s);
    double vw[G::VW];
#pragma unroll
    for (int t = 0; t < G::VW; t++) {
        const int e = t + 1 - T2; // v index relative to m (m even => parity of e is the parity of the index)
        double val;
        if ((e & 1) == 0) {
            val = uw[hbf_floor_half(e) - G::UB];
        } else {
            const int n = (e - 1) / 2; // exact: e - 1 is even
            const int c = n - G::UB;   // index of u[n] in uw
            val = f[0] * (uw[c + 1] + uw[c]);
#pragma unroll
            for (int j = 1; j < T1; j++) val = fma(f[j], uw[c + 1 + j] + uw[c - j], val);
        }
        if (may_be_negative && m + e < 0) val = 0.0;
        vw[t] = val;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int c = T2 - 1 + i; // vw index of v[m + i]
        double od = g[0] * (vw[c + 1] + vw[c]);
#pragma unroll
        for (int j = 1; j < T2; j++) od = fma(g[j], vw[c + 1 + j] + vw[c - j], od);
        y8[2 * i] = vw[c];
        y8[2 * i + 1] = od;
    }
}

} // namespace r8bgpu
