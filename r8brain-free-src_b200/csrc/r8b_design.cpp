// r8b_design.cpp -- see r8b_design.h.  Strict-IEEE host code (build with -ffp-contract=off).
#include "r8b_design.h"

#include <climits>
#include <cmath>
#include <cstring>

#include "r8b_tables.inc"

namespace r8bgpu {

namespace {

const double kPi = 3.14159265358979324; // r8bbase.h:179 (literal as in the reference)

inline double sq(double x) { return x * x; }

// The reference evaluates its formulas with its own log-based inverse hyperbolic sine
// (r8bbase.h:1175-1178); for large negative arguments that expression cancels heavily, so
// the exact form -- not the mathematically exact asinh -- is part of the filter definition.
inline double asinh_logform(double v) { return std::log(v + std::sqrt(v * v + 1.0)); }

// |v|^p through exp/log with the 1e-300 guard (r8bbase.h:1154-1157).
inline double pow_abs(double v, double p) { return std::exp(p * std::log(std::fabs(v) + 1e-300)); }

// Abramowitz-Stegun polynomial I0 (r8bbase.h:1192-1212).
double bessel_i0(double x)
{
    const double ax = std::fabs(x);
    if (ax < 3.75) {
        double y = x / 3.75;
        y *= y;
        return 1.0 + y * (3.5156229 + y * (3.0899424 + y * (1.2067492 + y * (0.2659732 +
               y * (0.360768e-1 + y * 0.45813e-2)))));
    }
    const double y = 3.75 / ax;
    return std::exp(ax) / std::sqrt(ax) * (0.39894228 + y * (0.1328592e-1 + y * (0.225319e-2 +
           y * (-0.157565e-2 + y * (0.916281e-2 + y * (-0.2057706e-1 + y * (0.2635537e-1 +
           y * (-0.1647633e-1 + y * 0.392377e-2))))))));
}

// Kaiser window sampled at integer positions wn, wn+1, ... on a support of half-length len2,
// optionally shifted by a fractional delay (CDSPSincFilterGen.h:230-241, 586-605).
struct KaiserWindow {
    double beta, mul, len2_inv, frac_shift;
    int wn;
    void init(double beta_, double len2, int first, double frac_delay)
    {
        beta = beta_ < 1.0 ? 1.0 : (beta_ > 350.0 ? 350.0 : beta_);
        mul = 1.0 / bessel_i0(beta);
        len2_inv = 1.0 / len2;
        frac_shift = frac_delay * len2_inv;
        wn = first;
    }
    double next()
    {
        const double n = 1.0 - sq(wn * len2_inv + frac_shift);
        wn++;
        if (n <= 0.0) return 0.0;
        return bessel_i0(beta * std::sqrt(n)) * mul;
    }
};

// sin(phase + k*step)*gain by the two-term recurrence (r8bbase.h:697-749).
struct SineRecurrence {
    double cur, prev, incr;
    SineRecurrence(double step, double phase, double gain)
        : cur(std::sin(phase) * gain), prev(std::sin(phase - step) * gain), incr(2.0 * std::cos(step)) {}
    double next()
    {
        const double r = cur;
        cur = incr * r - prev;
        prev = r;
        return r;
    }
};

} // namespace

int bit_occupancy(int v)
{
    unsigned u = (unsigned) v;
    int n = 1;
    while (u > 1u) {
        u >>= 1;
        n++;
    }
    return n;
}

bool design_lowpass(double norm_freq, double trans_band, double req_atten, double gain, int extfft,
                    LowpassDesign& out)
{
    // accepted ranges: CDSPFIRFilter.h:67-110
    if (!(norm_freq > 0.0 && norm_freq <= 1.0)) return false;
    if (!(trans_band >= 0.5 && trans_band <= 45.0)) return false;
    if (!(req_atten >= 49.0 && req_atten <= 218.0)) return false;
    if (!(gain > 0.0)) return false;

    const double tb = trans_band * 0.01;
    const int cls = tb >= 0.25 ? 0 : (tb >= 0.10 ? 1 : 2);
    const int lvl = req_atten >= 117.0 ? 0 : (req_atten >= 60.0 ? 1 : 2);
    static const double kClassShift[3][3] = {
        {1.60, 1.91, 2.25}, {0.69, 0.73, 1.13}, {0.21, 0.25, 0.36}}; // CDSPFIRFilter.h:228-276
    double atten = -req_atten;
    atten -= kClassShift[cls][lvl];

    int ci = (int) std::floor((-atten - 49.0) * 264 / 176.25 + 0.5); // :278-284
    if (ci < 0) ci = 0;
    if (ci > 264) ci = 264;
    atten -= R8B_ATTCORR[cls][ci] / R8B_ATTCORR_SCALE[cls];

    // Window power and half-length/cut-off formulas: CDSPFIRFilter.h:373-448.
    const double pwr = 7.43932822146293e-8 * sq(atten) +
        0.000102747434588003 * std::cos(0.00785021930010397 * atten) *
            std::cos(0.633854318781239 + 0.103208573657699 * atten) -
        0.00798132247867036 - 0.000903555213543865 * atten -
        0.0969365532127236 * std::exp(0.0779275237937911 * atten) -
        1.37304948662012e-5 * atten * std::cos(0.00785021930010397 * atten);

    double hl, fo1;
    if (pwr <= 0.067665322581) {
        if (cls == 0) {
            hl = 2.6778150875894 / tb +
                300.547590563091 * ::atan(::atan(2.68959772209918 * pwr)) /
                    (5.5099277187035 * tb - tb * ::tanh(std::cos(asinh_logform(atten))));
            fo1 = 0.987205355829873 * tb +
                1.00011788929851 *
                    std::atan2(-0.321432067051302 - 6.19131357321578 * std::sqrt(pwr),
                               hl + -1.14861472207245 / (hl - 14.1821147585957) +
                                   ::pow(0.9521145021664,
                                         ::pow(std::atan2(1.12018764830637, tb),
                                               2.10988901686912 * hl - 20.9691278378345)));
        } else if (cls == 1) {
            hl = (1.56688617018066 + 142.064321294568 * pwr +
                  0.00419441117131136 * std::cos(243.633511747297 * pwr) -
                  0.022953443903576 * atten -
                  0.026629568860284 * std::cos(127.715550622571 * pwr)) / tb;
            fo1 = 0.982299356642411 * tb +
                0.999441744774215 *
                    asinh_logform((-0.361783054039583 - 5.80540593623676 * std::sqrt(pwr)) / hl);
        } else {
            hl = (2.45739657014937 +
                  269.183679500541 * pwr *
                      std::cos(5.73225668178813 +
                               std::atan2(::cosh(0.988861169868941 - 17.2201556280744 * pwr),
                                          1.08340138240431 * pwr))) / tb;
            fo1 = 2.291956939 * tb + 0.01942450693 * sq(tb) * hl - 4.67538973161837 * pwr * tb -
                1.668433124 * tb * ::pow(pwr, pwr);
        }
    } else {
        if (cls == 0) {
            hl = (1.50258368698213 +
                  158.556968859477 * asinh_logform(pwr) * ::tanh(57.9466246871383 * ::tanh(pwr)) -
                  0.0105440479814834 * atten) / tb;
            fo1 = 0.994024401639321 * tb +
                (-0.236282717577215 - 6.8724924545387 * std::sqrt(std::sin(pwr))) / hl;
        } else if (cls == 1) {
            hl = (1.50277377248945 +
                  158.222625721046 * asinh_logform(pwr) *
                      ::tanh(1.02875299001715 + 42.072277322604 * pwr) -
                  0.0108380943845632 * atten) / tb;
            fo1 = 0.992539376734551 * tb +
                (-0.251747813037178 -
                 6.74159892452584 * std::sqrt(::tanh(::tanh(::tan(pwr))))) / hl;
        } else {
            hl = (1.15990238966306 * pwr - 5.02124037125213 * sq(pwr) -
                  0.158676856669827 * atten *
                      std::cos(1.1609073390614 * pwr - 6.33932586197475 * pwr * sq(pwr))) / tb;
            fo1 = 0.867344453126885 * tb + 0.052693817907757 * tb * std::log(pwr) +
                0.0895511178735932 * tb * ::atan(59.7538527741309 * pwr) -
                0.0745653568081453 * pwr * tb;
        }
    }

    const double len2 = 0.25 * hl / norm_freq;             // :455
    const double freq2 = kPi * (1.0 - fo1) * norm_freq;     // :457
    if (!(len2 >= 2.0)) return false;
    const int L = (int) std::floor(len2);
    const int K = L + L + 1;

    out.kernel_len = K;
    out.half_len = L;
    out.block_len_bits = bit_occupancy(K - 1) + (extfft ? 1 : 0); // :461
    out.taps.assign((size_t) K, 0.0);

    // Windowed sinc, centre outwards (CDSPSincFilterGen.h:312-337): sine recurrence scaled
    // by 1/pi, Kaiser(beta=125) raised to `pwr`.
    KaiserWindow win;
    win.init(125.0, len2, 0, 0.0);
    const double wpow = std::fabs(pwr);
    SineRecurrence osc(freq2, 0.0, 1.0 / kPi);
    osc.next(); // sin(0)
    double* c = out.taps.data() + L;
    c[0] = freq2 * pow_abs(win.next(), wpow) / kPi;
    for (int t = 1; t <= L; t++) {
        const double v = osc.next() * pow_abs(win.next(), wpow) / t;
        c[t] = v;
        c[-t] = v;
    }

    // DC normalisation: sequential sum from the first tap (CDSPFIRFilter.h:492-500).  The
    // reference folds a power-of-two FFT scale into the same multiplier, which is exact.
    double s = 0.0;
    for (int i = 0; i < K; i++) s += out.taps[(size_t) i];
    s = gain / s;
    for (int i = 0; i < K; i++) out.taps[(size_t) i] *= s;
    return true;
}

// ----------------------------------------------------------------------------------------------

namespace {

// One fractional-delay windowed-sinc row, evaluated tap by tap in closed form.
//
// Operator (what CDSPSincFilterGen.h:452-552 computes with running state): for tap position
// t = -fl2 .. fl2-1 and u = t + fd,
//     row[t] = 0                                   if u lies outside [-len2, len2]   (only the edge taps can)
//            = w(t)^wpow                           where u is a zero of the sinc argument
//                                                  (fd ~ 0 at t = 0, fd ~ 1 at t = -1: the 0/0 limit)
//            = (-1)^t * sin(pi fd)/pi * w(t)^wpow / u   everywhere else,
// w(t) = Kaiser window at (t + fd)/len2.  sin(pi (t+fd)) = (-1)^t sin(pi fd) is what makes one sine enough.
// The row is then DC-normalised by its plain left-to-right sum (r8bbase.h:934-961).  Each value is
// produced by the same operations in the same order as the reference, which keeps the bank bit-exact
// (tests/test_host_cpu.py::test_frac_banks_and_halfbands_bit_exact).
struct FracRow {
    double len2, fd, beta, mul, wpow, sinc_gain;
    int fl2, unit_tap; // unit_tap: the tap that carries the pure window value, or INT_MIN

    FracRow(double len2_, double frac_delay, double beta_, double wpow_) : len2(len2_), fd(frac_delay), wpow(wpow_)
    {
        fl2 = (int) std::ceil(len2); // CDSPSincFilterGen.h:168-176
        beta = beta_ < 1.0 ? 1.0 : (beta_ > 350.0 ? 350.0 : beta_);
        mul = 1.0 / bessel_i0(beta);
        sinc_gain = std::sin(fd * kPi) / kPi;
        const double kTiny = 2.3e-13; // the reference's threshold for "fd is an integer"
        unit_tap = std::fabs(fd - 1.0) < kTiny ? -1 : (std::fabs(fd) < kTiny ? 0 : INT_MIN);
    }
    double window(int t) const
    {
        const double n = 1.0 - sq(t * (1.0 / len2) + fd * (1.0 / len2));
        return n <= 0.0 ? 0.0 : bessel_i0(beta * std::sqrt(n)) * mul;
    }
    double tap(int t) const
    {
        const double u = t + fd;
        if ((t == -fl2 && u < -len2) || (t == fl2 - 1 && u > len2)) return 0.0;
        const double w = pow_abs(window(t), wpow);
        if (t == unit_tap) return w;
        const double s = (t & 1) ? -sinc_gain : sinc_gain;
        return s * w / u;
    }
};

void frac_delay_row(double* op, int stride, int filter_len, double len2, double frac_delay, double beta, double wpow)
{
    const FracRow row(len2, frac_delay, beta, wpow);
    double sum = 0.0;
    for (int i = 0; i < filter_len; i++) {
        const double v = row.tap(i - row.fl2);
        op[(size_t) i * stride] = v;
        sum += v;
    }
    const double norm = 1.0 / sum;
    for (int i = 0; i < filter_len; i++) op[(size_t) i * stride] *= norm;
}

} // namespace

void design_frac_bank(int init_fracs, double req_atten, bool is_third, FracBank& out)
{
    // Window row selection: CDSPFracInterpolator.h:279-341.
    const double(*rows)[3] = is_third ? R8B_FRACWIN3 : R8B_FRACWIN2;
    const int nrows = is_third ? 10 : 12;
    const int base = is_third ? 6 : 8;
    int r = 0;
    while (r != nrows - 1 && rows[r][2] < req_atten) r++;
    const double beta = rows[r][0];
    const double wpow = std::fabs(rows[r][1]);
    const double att = rows[r][2];
    const int flen = base + r * 2;

    const bool whole = (init_fracs != -1);
    const int elsize = whole ? 1 : 3;
    const int interp_points = whole ? 2 : 8;
    const int fracs = whole ? init_fracs : (int) std::ceil(::pow(6.4, att / 50.0)); // :80-96
    const int pc2 = interp_points / 2;
    const int nrows_tab = fracs + interp_points;
    const size_t fsize = (size_t) flen * elsize;

    std::vector<double> tab(fsize * nrows_tab, 0.0);
    const double len2 = flen / 2; // integer division, as in the reference (:101)
    double* p = tab.data();
    for (int i = -pc2 + 1; i <= fracs + pc2; i++) { // :107-116
        const double fd = (double) (fracs - i) / fracs;
        frac_delay_row(p, elsize, flen, len2, fd, beta, wpow);
        p += fsize;
    }

    if (!whole) {
        // 2nd-order 8-point spline coefficients computed in place (row r reads rows r..r+7 and
        // receives the polynomial of row r+3): CDSPFracInterpolator.h:128-147, r8bbase.h:1014-1024.
        const double k = 1.31578947368421052e-2;
        double* q = tab.data();
        double* const qend = tab.data() + (size_t) (fracs + 1) * fsize;
        while (q < qend) {
            const double xm3 = q[0], xm2 = q[fsize], xm1 = q[2 * fsize], x0 = q[3 * fsize],
                         x1 = q[4 * fsize], x2 = q[5 * fsize], x3 = q[6 * fsize], x4 = q[7 * fsize];
            q[0] = x0;
            q[1] = (61.0 * (x1 - xm1) + 16.0 * (xm2 - x2) + 3.0 * (x3 - xm3)) * k;
            q[2] = (106.0 * (xm1 + x1) + 10.0 * x3 + 6.0 * xm3 - 3.0 * x4 - 29.0 * (xm2 + x2) -
                    167.0 * x0) * k;
            q += elsize;
        }
    }

    out.filter_len = flen;
    out.fracs = fracs;
    out.order = whole ? 0 : 2;
    out.atten = att;
    out.table.assign(tab.begin(), tab.begin() + (size_t) (fracs + 1) * fsize);
}

HalfbandTaps select_halfband(double req_atten, int steep_index, bool is_third)
{
    // Families beyond the last one reuse it (CDSPHBUpsampler.h:303-315 "else" branch).
    int steep = steep_index < 0 ? 0 : steep_index;
    int max_steep = 0;
    for (int i = 0; i < R8B_HB_INDEX_COUNT; i++)
        if (R8B_HB_INDEX[i].third == (is_third ? 1 : 0) && R8B_HB_INDEX[i].steep > max_steep)
            max_steep = R8B_HB_INDEX[i].steep;
    if (steep > max_steep) steep = max_steep;
    HalfbandTaps res;
    for (int i = 0; i < R8B_HB_INDEX_COUNT; i++) {
        if (R8B_HB_INDEX[i].third != (is_third ? 1 : 0) || R8B_HB_INDEX[i].steep != steep) continue;
        res.ntaps = R8B_HB_INDEX[i].ntaps;
        res.atten = R8B_HB_INDEX[i].atten;
        res.taps = &R8B_HB_TAPS[R8B_HB_INDEX[i].offs];
        if (!(R8B_HB_INDEX[i].atten < req_atten)) break; // first set reaching the request
    }
    return res;
}

bool whole_stepping(double src_rate, double dst_rate, int& in_step, int& out_step)
{
    // The interpolator steps by whole bank rows when Src/Dst reduces to InStep/OutStep with a small OutStep.
    // The reference looks for the common divisor on the doubles themselves and gives up after a fixed budget
    // of 149 equality checks, i.e. 148 reduction steps (CDSPFracInterpolator.h:609-628), then rejects OutStep > 1500 (:664-670); both
    // limits decide which rates take the order-2 bank instead, so they are part of the plan and kept as
    // explicit guards.  Differences of exactly representable rates are exact, so this is Euclid's
    // subtractive algorithm on the pair (larger, smaller) = (a, b) -> (b, |a - b|).
    constexpr int kMaxReductions = 148, kMaxOutStep = 1500;
    double a = src_rate, b = dst_rate;
    int reductions = 0;
    while (a != b) {
        if (++reductions > kMaxReductions) return false;
        const double d = std::fabs(a - b);
        a = b;
        b = d;
    }
    if (!(b > 0.0)) return false;
    const double qi = src_rate / b, qo = dst_rate / b;
    if (qi != std::floor(qi) || qo != std::floor(qo) || qi > 2147483647.0 || qo > (double) kMaxOutStep) return false;
    in_step = (int) qi;
    out_step = (int) qo;
    return true;
}

} // namespace r8bgpu
