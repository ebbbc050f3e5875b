// Host check of the index algebra in csrc/r8b_hbfuse.cuh: the fused two-stage item must equal, bit for bit,
// two plain half-band upsampler stages applied one after the other (same summation order), including the
// "negative stream indices are zeros" rule at the start of the stream.
#include <cmath>
#include <cstdio>
#include <vector>

using std::fma;
#include "../../r8brain-free-src_b200/csrc/r8b_hbfuse.cuh"

static double urand(unsigned long long& s)
{
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    return (double) (s >> 11) / 9007199254740992.0 * 2.0 - 1.0;
}

template <int T>
static void stage(const std::vector<double>& x, long long x0, const double (&f)[T], std::vector<double>& y, long long y0)
{
    // y covers absolute indices [y0, y0 + y.size()); x covers [x0, x0 + x.size()), zeros elsewhere are NOT assumed:
    // the caller sizes x generously.  Negative output indices are zeros.
    auto X = [&](long long n) { return n < 0 ? 0.0 : x[(size_t) (n - x0)]; };
    for (size_t i = 0; i < y.size(); i++) {
        const long long a = y0 + (long long) i;
        if (a < 0) { y[i] = 0.0; continue; }
        const long long n = a >> 1;
        if ((a & 1) == 0) { y[i] = X(n); continue; }
        double od = f[0] * (X(n + 1) + X(n));
        for (int j = 1; j < T; j++) od = fma(f[j], X(n + 1 + j) + X(n - j), od);
        y[i] = od;
    }
}

template <int T1, int T2>
static int check()
{
    using G = r8bgpu::HbFuseGeom<T1, T2>;
    unsigned long long seed = 0x9E3779B97F4A7C15ULL + T1 * 131 + T2;
    double f[T1], g[T2];
    for (int j = 0; j < T1; j++) f[j] = urand(seed) * 0.6;
    for (int j = 0; j < T2; j++) g[j] = urand(seed) * 0.6;
    const long long x0 = -64;
    std::vector<double> u(600);
    for (auto& v : u) v = urand(seed);
    for (long long n = x0; n < 0; n++) u[(size_t) (n - x0)] = 0.0; // stream values at negative indices are zeros
    std::vector<double> v(1000), y(1900);
    const long long v0 = -40, y0 = -20;
    stage<T1>(u, x0, f, v, v0);
    stage<T2>(v, v0, g, y, y0);
    int bad = 0;
    for (long long m = 0; m <= 400; m += 2) { // even first positions, starting AT the stream start
        const long long h = m / 2;
        auto U = [&](int s) { return u[(size_t) (h + G::UB + s - x0)]; };
        double y8[8];
        r8bgpu::hb_fused_item<T1, T2>(f, g, U, m, true, y8);
        for (int i = 0; i < 8; i++)
            if (y8[i] != y[(size_t) (2 * m + i - y0)]) bad++;
    }
    if (bad) std::printf("T1=%d T2=%d: %d mismatches\n", T1, T2, bad);
    return bad;
}

int main()
{
    int bad = 0;
    bad += check<4, 3>();
    bad += check<5, 4>();
    bad += check<6, 5>();
    bad += check<2, 1>();
    bad += check<1, 1>();
    bad += check<3, 2>();
    bad += check<11, 6>();
    bad += check<2, 2>();
    std::printf(bad ? "FAIL\n" : "hbfuse ok\n");
    return bad ? 1 : 0;
}
