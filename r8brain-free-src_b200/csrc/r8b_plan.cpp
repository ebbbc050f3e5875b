// r8b_plan.cpp -- see r8b_plan.h.  Strict-IEEE host code (build with -ffp-contract=off).
#include "r8b_plan.h"

#include <cmath>
#include <cstdio>

namespace r8bgpu {

namespace {

long long ceil_div(long long a, long long b) { return (a + b - 1) / b; } // a >= 0, b > 0

bool make_blockconv(StageDesc& s, double norm_freq, double tb, double atten, double gain, int up,
                    int down, int extfft, std::string& err)
{
    s.kind = ST_BLOCKCONV;
    s.up = up;
    s.down = down;
    s.norm_freq = norm_freq;
    s.trans_band = tb;
    s.gain = gain;
    if (!design_lowpass(norm_freq, tb, atten, gain, extfft, s.lp)) {
        err = "low-pass design parameters out of range";
        return false;
    }
    // Block geometry of the reference convolver -- needed only to reproduce WHEN samples are
    // emitted (CDSPBlockConvolver.h:75-146); the CUDA tiles use their own FFT size.
    const int K = s.lp.kernel_len, L = s.lp.half_len;
    const int b2 = 2 << s.lp.block_len_bits;
    int ushift = bit_occupancy(up) - 1;
    int prev_len, in_len;
    if ((1 << ushift) == up) {
        prev_len = (K - 1 + up - 1) / up;
        in_len = b2 - prev_len * up;
    } else {
        ushift = -1;
        prev_len = K - 1;
        in_len = b2 - prev_len;
    }
    int latency = in_len + L;
    const int dshift = bit_occupancy(down) - 1;
    if ((1 << dshift) == down && down > 1) {
        if (ushift > 0) {
            err = "power-of-two up- and down-factors together are not planned by the reference";
            return false;
        }
        const int ilc = in_len & (down - 1);
        prev_len += ilc;
        in_len -= ilc;
        latency -= ilc;
        s.block_exact = true;
    }
    s.ref_input_len = in_len;
    s.ref_prev_len = prev_len;
    s.latency = latency;
    const int lg = (L + up - 1) / up + 1;
    s.src_history = (latency + L + up - 1) / up + lg + 40;
    if (s.block_exact) s.src_history = 2 * b2 + 16;
    return true;
}

bool make_frac(StageDesc& s, double src, double dst, double atten, bool is_third, int fasttiming,
               std::string& err)
{
    s.src_rate = src;
    s.dst_rate = dst;
    s.is_third = is_third;
    int a = 0, b = 0;
    if (whole_stepping(src, dst, a, b)) {
        s.kind = ST_FRAC_WHOLE;
        s.in_step = a;
        s.out_step = b;
        design_frac_bank(b, atten, is_third, s.bank);
    } else {
        s.fasttiming = fasttiming != 0;
        s.kind = ST_FRAC_POLY;
        design_frac_bank(-1, atten, is_third, s.bank);
    }
    s.src_history = s.bank.filter_len + 8;
    return true;
}

void make_hb(StageDesc& s, StageKind kind, double atten, int steep, bool is_third)
{
    s.kind = kind;
    s.steep_index = steep;
    s.is_third = is_third;
    const HalfbandTaps t = select_halfband(atten, steep, is_third);
    s.hb_taps = t.ntaps;
    s.hb_atten = t.atten;
    s.hb.assign(t.taps, t.taps + t.ntaps);
    s.src_history = (kind == ST_HBUP ? 2 : 4) * t.ntaps + 8;
}

int stage_max_out_len(const StageDesc& s, int max_in)
{
    switch (s.kind) {
    case ST_BLOCKCONV:
        return (max_in * s.up + s.down - 1) / s.down;              // CDSPBlockConvolver.h:208-213
    case ST_FRAC_WHOLE:
    case ST_FRAC_POLY:
        return (int) std::ceil(max_in * s.dst_rate / s.src_rate) + 1; // CDSPFracInterpolator.h:827-832
    case ST_HBUP:
        return max_in * 2;                                           // CDSPHBUpsampler.h:648-653
    case ST_HBDOWN:
        return (max_in + 1) >> 1;                                    // CDSPHBDownsampler.h:113-118
    }
    return 0;
}

int stage_in_len_before_out_pos(const StageDesc& s, int pos)
{
    switch (s.kind) {
    case ST_BLOCKCONV: // CDSPBlockConvolver.h:192-196 (LatencyFrac == 0 for linear phase)
        return (int) ((s.latency + (double) pos * s.down) / s.up + 0.0 * s.down / s.up);
    case ST_FRAC_WHOLE: // CDSPFracInterpolator.h:802-811
        return s.bank.filter_len / 2 +
            (int) ((0 + (double) pos * s.in_step) / s.out_step + 0.0 * s.in_step / s.out_step);
    case ST_FRAC_POLY: // :813-814
        return s.bank.filter_len / 2 + (int) (0.0 + pos * s.src_rate / s.dst_rate);
    case ST_HBUP: // CDSPHBUpsampler.h:633-636
        return s.hb_taps + (int) ((0 + 0.0 + pos) * 0.5);
    case ST_HBDOWN: // CDSPHBDownsampler.h:98-101
        return (2 * s.hb_taps - 1) + (int) ((0 + 0.0 + pos) * 2.0);
    }
    return 0;
}

} // namespace

long long blockconv_emitted(const StageDesc& s, long long n)
{
    const long long avail = (long long) s.up * n - s.latency;
    return avail <= 0 ? 0 : ceil_div(avail, s.down);
}

long long frac_whole_emitted(const StageDesc& s, long long n)
{
    // outputs j >= 0 with floor(j*InStep/OutStep) + fl2 <= n-1
    const long long fl2 = s.bank.filter_len / 2;
    const long long pmax = n - 1 - fl2;
    if (pmax < 0) return 0;
    // floor(j*a/b) <= pmax  <=>  j*a <= pmax*b + b - 1  <=>  j <= (pmax*b + b - 1)/a
    const long long a = s.in_step, b = s.out_step;
    return (pmax * b + b - 1) / a + 1;
}

long long hbup_emitted(const StageDesc& s, long long n)
{
    const long long c = n - s.hb_taps;
    return c <= 0 ? 0 : 2 * c;
}

long long hbdown_emitted(const StageDesc& s, long long n)
{
    const long long c = n / 2 - (s.hb_taps - 1);
    return c <= 0 ? 0 : c;
}

bool Plan::build(double src, double dst, int max_in, double tb, double att, int phase, int ext,
                 int fasttiming)
{
    src_rate = src;
    dst_rate = dst;
    max_in_len = max_in;
    trans_band = tb;
    atten = att;
    extfft = ext ? 1 : 0;
    stages.clear();
    passthrough = false;
    error.clear();
    max_out_len = max_in;

    if (!(src > 0.0) || !(dst > 0.0) || max_in <= 0) {
        error = "invalid sample rates or MaxInLen";
        return false;
    }
    if (phase != 0) {
        error = "only fprLinearPhase is implemented (minimum-phase is out of scope)";
        return false;
    }
    if (src == dst) { // CDSPResampler.h:135-138
        passthrough = true;
        return true;
    }

    auto push_bc = [&](double nf, double tbw, double gain, int up, int down) -> bool {
        StageDesc s;
        if (!make_blockconv(s, nf, tbw, att, gain, up, down, extfft, error)) return false;
        stages.push_back(std::move(s));
        return true;
    };
    auto push_frac = [&](double s_rate, double d_rate, bool third) -> bool {
        StageDesc s;
        if (!make_frac(s, s_rate, d_rate, att, third, fasttiming, error)) return false;
        stages.push_back(std::move(s));
        return true;
    };
    auto push_hb = [&](StageKind k, int steep, bool third) {
        StageDesc s;
        make_hb(s, k, att, steep, third);
        stages.push_back(std::move(s));
    };

    bool done = false;

    // (1) single-step common ratios, CDSPResampler.h:146-172
    static const int kCommon[5][2] = {{1, 2}, {1, 3}, {2, 3}, {3, 2}, {3, 4}};
    for (int i = 0; i < 5 && !done; i++) {
        const int num = kCommon[i][0], den = kCommon[i][1];
        if (src * num == dst * den) {
            if (!push_bc(1.0 / (num > den ? num : den), tb, num, num, den)) return false;
            done = true;
        }
    }

    // (2) whole 2^c or 3*2^c upsampling, :176-216
    for (int i = 2; i <= 3 && !done; i++) {
        bool found = false;
        int c = 0;
        while (true) {
            const double nsr = src * (i << c);
            if (nsr == dst) {
                found = true;
                break;
            }
            if (nsr > dst) break;
            c++;
        }
        if (found) {
            if (!push_bc(1.0 / i, tb, i, i, 1)) return false;
            for (int k = 0; k < c; k++) push_hb(ST_HBUP, k, i == 3);
            done = true;
        }
    }

    if (!done && dst * 2.0 > src) {
        // (3) upsampling or fractional downsampling down to 2X, :218-333
        const double nf = (dst > src ? 0.5 : 0.5 * dst / src);
        if (!push_bc(nf, tb, 2.0, 2, 1)) return false;

        const double tbw = 0.0175;
        const double thresh = src / (1.0 - tbw * tb);
        int c = 0, div = 1;
        while (true) {
            const int ndiv = div * 2;
            if (dst < thresh * ndiv) break;
            div = ndiv;
            c++;
        }
        int c2 = 0, div2 = 1;
        while (true) {
            const int ndiv = div * (c2 == 0 ? 3 : 2);
            if (dst < thresh * ndiv) break;
            div2 = ndiv;
            c2++;
        }
        const double src2 = src * 2.0;
        int t1, t2;
        if (c == 1 && whole_stepping(src2, dst, t1, t2)) c = 0;

        if (c > 0) {
            int num;
            if (c2 > 0 && div2 > div) {
                div = div2;
                c = c2;
                num = 3;
            } else {
                num = 2;
            }
            if (!push_frac(src2 * div, dst, false)) return false;
            double tb2 = (1.0 - src * div / dst) / tbw;
            if (tb2 > 45.0) tb2 = 45.0; // CDSPFIRFilter::getLPMaxTransBand()
            if (!push_bc(1.0 / num, tb2, num, num, 1)) return false;
            for (int k = 1; k < c; k++) push_hb(ST_HBUP, k - 1, num == 3);
        } else {
            if (!push_frac(src2, dst, false)) return false;
        }
        done = true;
    }

    if (!done) {
        // (4) downsampling with half-band decimators, :337-393
        double check = dst * 4.0;
        int c = 0;
        double fin_gain = 1.0;
        while (check <= src) {
            c++;
            check *= 2.0;
            fin_gain *= 0.5;
        }
        const int srdiv = (1 << c);
        int downf;
        double nf = 0.5;
        bool use_interp = true, third = false;
        for (downf = 2; downf <= 3; downf++) {
            if (dst * srdiv * downf == src) {
                nf = 1.0 / downf;
                use_interp = false;
                third = (downf == 3);
                break;
            }
        }
        if (use_interp) {
            downf = 1;
            nf = dst * srdiv / src;
            third = (nf * 3.0 <= 1.0);
        }
        for (int k = 0; k < c; k++) push_hb(ST_HBDOWN, c - 1 - k, third);
        if (!push_bc(nf, tb, fin_gain, 1, downf)) return false;
        if (use_interp && !push_frac(src, dst * srdiv, third)) return false;
    }

    // Buffer-length chain (addProcessor, CDSPResampler.h:677-700).
    int cur = max_in;
    for (auto& s : stages) {
        cur = stage_max_out_len(s, cur);
        s.max_out_len = cur;
    }
    max_out_len = cur;
    return true;
}

bool Plan::build_single(int kind, const double* a, int max_in, int ext)
{
    src_rate = 0;
    dst_rate = 0;
    max_in_len = max_in;
    extfft = ext ? 1 : 0;
    stages.clear();
    passthrough = false;
    error.clear();
    StageDesc s;
    switch (kind) {
    case ST_BLOCKCONV:
        atten = a[2];
        trans_band = a[1];
        if (!make_blockconv(s, a[0], a[1], a[2], a[3], (int) a[4], (int) a[5], extfft, error)) return false;
        break;
    case ST_FRAC_WHOLE:
    case ST_FRAC_POLY:
        atten = a[2];
        if (!make_frac(s, a[0], a[1], a[2], a[3] != 0.0, 0, error)) return false;
        break;
    case ST_HBUP:
    case ST_HBDOWN:
        atten = a[0];
        make_hb(s, (StageKind) kind, a[0], (int) a[1], a[2] != 0.0);
        break;
    default:
        error = "unknown stage kind";
        return false;
    }
    s.max_out_len = stage_max_out_len(s, max_in);
    max_out_len = s.max_out_len;
    stages.push_back(std::move(s));
    return true;
}

int Plan::in_len_before_out_pos(int req_out_pos) const
{
    int req = req_out_pos;
    for (int c = (int) stages.size() - 1; c >= 0; c--)
        req = stage_in_len_before_out_pos(stages[(size_t) c], req);
    return req;
}

int Plan::input_required_for_output(int n) const
{
    if (n < 1) return 0;
    return in_len_before_out_pos(n - 1) + 1;
}

std::string Plan::describe() const
{
    char b[512];
    std::string r;
    snprintf(b, sizeof b, "* plan: src=%.1f dst=%.1f len=%i tb=%.1f att=%.2f extfft=%i\n", src_rate,
             dst_rate, max_in_len, trans_band, atten, extfft);
    r += b;
    for (const auto& s : stages) {
        switch (s.kind) {
        case ST_BLOCKCONV:
            snprintf(b, sizeof b, "BlockConv: flt_len=%i in_len=%i io=%i/%i latency=%i nfreq=%.4f gain=%.3f\n",
                     s.lp.kernel_len, s.ref_input_len, s.up, s.down, s.latency, s.norm_freq, s.gain);
            break;
        case ST_FRAC_WHOLE:
            snprintf(b, sizeof b, "FracInterp: src=%.2f dst=%.2f taps=%i whole step=%i/%i third=%i\n",
                     s.src_rate, s.dst_rate, s.bank.filter_len, s.in_step, s.out_step, (int) s.is_third);
            break;
        case ST_FRAC_POLY:
            snprintf(b, sizeof b, "FracInterp: src=%.2f dst=%.2f taps=%i fracs=%i order=2 third=%i\n",
                     s.src_rate, s.dst_rate, s.bank.filter_len, s.bank.fracs, (int) s.is_third);
            break;
        case ST_HBUP:
            snprintf(b, sizeof b, "HBUp: sti=%i third=%i taps=%i att=%.1f\n", s.steep_index,
                     (int) s.is_third, s.hb_taps, s.hb_atten);
            break;
        case ST_HBDOWN:
            snprintf(b, sizeof b, "HBDown: sti=%i third=%i taps=%i att=%.1f\n", s.steep_index,
                     (int) s.is_third, s.hb_taps, s.hb_atten);
            break;
        }
        r += b;
    }
    return r;
}

// ----------------------------------------------------------------------------------------------

void Schedule::init(const Plan* p)
{
    plan = p;
    clear();
}

void Schedule::clear()
{
    const size_t n = plan->stages.size();
    n_in.assign(n, 0);
    n_out.assign(n, 0);
    poly.assign(n, PolyState()); // InitFracPos == 0 for linear-phase chains (CDSPFracInterpolator.h:834-858)
}

namespace {

// Position of the k-th output after the call-start state (k >= 1); mirrors the reference's
// expression order exactly: ((InCounter + InPosShift) * ssr) / dsr   (CDSPFracInterpolator.h:1161-1166)
inline void poly_pos(const Schedule::PolyState& st, double ssr, double dsr, long long k,
                     long long& p, int& ni, double& fpos)
{
    const int ic = st.in_counter + (int) k;
    const double next_pos = (ic + st.in_pos_shift) * ssr / dsr;
    ni = (int) next_pos;
    p = st.p + (ni - st.in_pos_int);
    fpos = next_pos - ni;
}

} // namespace

int Schedule::advance(int l, std::vector<StageCall>& calls)
{
    const auto& st = plan->stages;
    calls.assign(st.size(), StageCall());
    long long feed0 = 0, feed1 = 0;
    for (size_t i = 0; i < st.size(); i++) {
        StageCall& c = calls[i];
        if (i == 0) {
            c.n0 = n_in[0];
            c.n1 = n_in[0] + l;
        } else {
            c.n0 = feed0;
            c.n1 = feed1;
        }
        c.e0 = n_out[i];
        const StageDesc& s = st[i];
        long long e1 = c.e0;
        switch (s.kind) {
        case ST_BLOCKCONV: e1 = blockconv_emitted(s, c.n1); break;
        case ST_FRAC_WHOLE: e1 = frac_whole_emitted(s, c.n1); break;
        case ST_HBUP: e1 = hbup_emitted(s, c.n1); break;
        case ST_HBDOWN: e1 = hbdown_emitted(s, c.n1); break;
        case ST_FRAC_POLY: {
            PolyState& ps = poly[i];
            c.in_counter0 = ps.in_counter;
            c.in_pos_int0 = ps.in_pos_int;
            c.in_pos_shift = ps.in_pos_shift;
            c.fpos0 = ps.fpos;
            c.p0 = ps.p;
            const long long fl2 = s.bank.filter_len / 2;
            const long long pmax = c.n1 - 1 - fl2; // produce while p <= pmax
            long long cnt = 0;
            if (s.fasttiming) {
                // CDSPFracInterpolator.h:1153-1158: fpos += FracStep; PosIncr = (int) fpos; fpos -= PosIncr
                const double frac_step = s.src_rate / s.dst_rate; // :713
                c.p_last = ps.p;
                while (ps.p <= pmax) {
                    c.ft_dp.push_back((int) (ps.p - c.p0));
                    c.ft_fpos.push_back(ps.fpos);
                    c.p_last = ps.p;
                    cnt++;
                    ps.fpos += frac_step;
                    const int inc = (int) ps.fpos;
                    ps.fpos -= inc;
                    ps.p += inc;
                }
                e1 = c.e0 + cnt;
                break;
            }
            if (ps.p <= pmax) {
                // largest k with p_k <= pmax (p_k is non-decreasing in k); k = 0 qualifies.
                long long lo = 0;
                long long hi = (long long) ((double) (pmax - ps.p + 2) * s.dst_rate / s.src_rate) + 4;
                auto pk = [&](long long k) {
                    long long p;
                    int ni;
                    double f;
                    poly_pos(ps, s.src_rate, s.dst_rate, k, p, ni, f);
                    return p;
                };
                while (pk(hi) <= pmax) hi *= 2;
                while (hi - lo > 1) {
                    const long long mid = lo + (hi - lo) / 2;
                    if (pk(mid) <= pmax) lo = mid;
                    else hi = mid;
                }
                cnt = lo + 1;
            }
            e1 = c.e0 + cnt;
            c.p_last = ps.p;
            if (cnt > 1) {
                long long p;
                int ni;
                double f;
                poly_pos(ps, s.src_rate, s.dst_rate, cnt - 1, p, ni, f);
                c.p_last = p;
            }
            if (cnt > 0) {
                long long p;
                int ni;
                double f;
                poly_pos(ps, s.src_rate, s.dst_rate, cnt, p, ni, f);
                ps.in_counter += (int) cnt;
                ps.p = p;
                ps.in_pos_int = ni;
                ps.fpos = f;
            }
            if (ps.in_counter > 1000) { // once per process() call, CDSPFracInterpolator.h:907-919
                ps.in_counter = 0;
                ps.in_pos_int = 0;
                ps.in_pos_shift = ps.fpos * s.dst_rate / s.src_rate;
            }
            break;
        }
        }
        c.e1 = e1;
        n_in[i] = c.n1;
        n_out[i] = e1;
        feed0 = c.e0;
        feed1 = c.e1;
    }
    if (st.empty()) return l;
    return (int) (calls.back().e1 - calls.back().e0);
}

} // namespace r8bgpu
