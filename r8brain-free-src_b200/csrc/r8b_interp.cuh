// r8b_interp.cuh -- inner products of order-2 filter-bank rows with windows of samples
// (CDSPFracInterpolator::convolve2, CDSPFracInterpolator.h:1069-1179).
//
// Bank row layout: [tap][c0, c1, c2]; tap coefficient c = c0 + c1*x + c2*x^2 evaluated as
// fma(c2, x2, fma(c1, x, c0)) and accumulated in ascending tap order (one fma per tap) -- the order is
// part of the parity contract, the grouping below only changes how the operands are FETCHED:
// taps are taken six at a time, each pair of taps = 6 doubles = three 128-bit loads (rows start 16-byte
// aligned because filter lengths are even), so the loads of a group are in flight before its first fma.
//
// Measured on B200 (tools/microbench.cu): a warp-wide LDS.128 whose lanes all read the same address still
// costs 2.1 clk, an LDS.64 with lane stride 1 costs 2 clk, stride 2 costs 4 clk.  A loop that fetches the
// 72 row coefficients and 24 samples per output is therefore shared-memory bound (~170 clk per 32 outputs
// against 36 clk of DFMA issue).  poly_block4 amortises the coefficient loads over four consecutive
// outputs of one thread, which share the bank row whenever the row index drifts slowly.
#pragma once
#include <cuda_runtime.h>

namespace r8bgpu {

// two taps from three double2: (c0,c1) (c2,c0') (c1',c2')
__device__ __forceinline__ double poly_two_taps(double acc, double2 a, double2 b, double2 c, double x, double x2,
                                                double y0, double y1)
{
    acc = fma(fma(b.x, x2, fma(a.y, x, a.x)), y0, acc);
    acc = fma(fma(c.y, x2, fma(c.x, x, b.y)), y1, acc);
    return acc;
}

// ROW_SMEM: the row was staged in shared memory (plain loads), otherwise it is read from global memory
// through the read-only path.
template <bool ROW_SMEM>
__device__ __forceinline__ double2 poly_ld2(const double2* p)
{
    return ROW_SMEM ? *p : __ldg(p);
}
template <bool ROW_SMEM>
__device__ __forceinline__ double poly_ld1(const double* p)
{
    return ROW_SMEM ? *p : __ldg(p);
}

// One output.  row: first coefficient of the bank row (16-byte aligned, flen even -- else the scalar path
// runs); y(i): the sample that multiplies tap i.
template <bool ROW_SMEM, typename YF>
__device__ __forceinline__ double poly_row_dot(const double* __restrict__ row, int flen, double x, double x2, YF y)
{
    double acc = 0.0;
    if ((flen & 1) != 0 || (reinterpret_cast<size_t>(row) & 15) != 0) {
        for (int i = 0; i < flen; i++) {
            const double c = fma(poly_ld1<ROW_SMEM>(row + 3 * i + 2), x2,
                                 fma(poly_ld1<ROW_SMEM>(row + 3 * i + 1), x, poly_ld1<ROW_SMEM>(row + 3 * i)));
            acc = fma(c, y(i), acc);
        }
        return acc;
    }
    const double2* __restrict__ r2 = reinterpret_cast<const double2*>(row);
    int i = 0;
#pragma unroll 1
    for (; i + 6 <= flen; i += 6) {
        double2 v[9];
        double w[6];
#pragma unroll
        for (int u = 0; u < 9; u++) v[u] = poly_ld2<ROW_SMEM>(r2 + (3 * i) / 2 + u);
#pragma unroll
        for (int u = 0; u < 6; u++) w[u] = y(i + u);
#pragma unroll
        for (int t = 0; t < 3; t++) acc = poly_two_taps(acc, v[3 * t], v[3 * t + 1], v[3 * t + 2], x, x2, w[2 * t], w[2 * t + 1]);
    }
#pragma unroll 1
    for (; i < flen; i += 2) {
        const double2 a = poly_ld2<ROW_SMEM>(r2 + (3 * i) / 2), b = poly_ld2<ROW_SMEM>(r2 + (3 * i) / 2 + 1),
                      c = poly_ld2<ROW_SMEM>(r2 + (3 * i) / 2 + 2);
        acc = poly_two_taps(acc, a, b, c, x, x2, y(i), y(i + 1));
    }
    return acc;
}

// Four consecutive outputs that share one staged bank row (flen even, row in shared memory) and whose
// windows start N samples apart: y(j) is the sample at offset j from the FIRST output's window start, so
// output r, tap i reads y(N*r + i).  Per group of six taps: 9 coefficient loads and 6 + 3N sample loads feed
// 72 fma (instead of 36 + 24 loads).  Same per-output arithmetic as poly_row_dot.
template <int N, typename YF>
__device__ __forceinline__ void poly_block4(const double* __restrict__ row, int flen, const double (&x)[4],
                                            const double (&x2)[4], YF y, double (&acc)[4])
{
    const double2* __restrict__ r2 = reinterpret_cast<const double2*>(row);
#pragma unroll
    for (int r = 0; r < 4; r++) acc[r] = 0.0;
    int i = 0;
#pragma unroll 1
    for (; i + 6 <= flen; i += 6) {
        double2 v[9];
        double w[6 + 3 * N];
#pragma unroll
        for (int u = 0; u < 9; u++) v[u] = r2[(3 * i) / 2 + u];
#pragma unroll
        for (int u = 0; u < 6 + 3 * N; u++) w[u] = y(i + u);
#pragma unroll
        for (int t = 0; t < 3; t++)
#pragma unroll
            for (int r = 0; r < 4; r++)
                acc[r] = poly_two_taps(acc[r], v[3 * t], v[3 * t + 1], v[3 * t + 2], x[r], x2[r], w[N * r + 2 * t],
                                       w[N * r + 2 * t + 1]);
    }
#pragma unroll 1
    for (; i < flen; i += 2) {
        const double2 a = r2[(3 * i) / 2], b = r2[(3 * i) / 2 + 1], c = r2[(3 * i) / 2 + 2];
        double w[2 + 3 * N];
#pragma unroll
        for (int u = 0; u < 2 + 3 * N; u++) w[u] = y(i + u);
#pragma unroll
        for (int r = 0; r < 4; r++) acc[r] = poly_two_taps(acc[r], a, b, c, x[r], x2[r], w[N * r], w[N * r + 1]);
    }
}

} // namespace r8bgpu
