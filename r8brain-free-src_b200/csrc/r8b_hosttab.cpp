// r8b_hosttab.cpp -- see r8b_hosttab.h.
#include "r8b_hosttab.h"

#include "r8b_fused2_core.cuh"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>

#include "r8b_fft.cuh"

namespace r8bgpu {

namespace {
// Spectrum of the polyphase-packed filter in FFT slot order (see k_blockconv).
template <int M>
void fill_slot_order(const std::vector<double2>& nat, std::vector<double2>& out)
{
    out.resize((size_t) M);
    for (int k = 0; k < M; k++) out[(size_t) slot_of<M>(k)] = nat[(size_t) k];
}

} // namespace

void build_spectrum(const StageDesc& s, int fft_log2, std::vector<double2>& spec_slots,
                    std::vector<double2>& tw, double* nyq_gain)
{
    const int M = 1 << fft_log2;
    const int L = s.lp.half_len, U = (s.up > 2) ? 1 : s.up; // up = 3 runs on the zero-stuffed stream
    const long double two_pi = 6.283185307179586476925286766559005768L;
    std::vector<long double> cs((size_t) M), sn((size_t) M);
    for (int k = 0; k < M; k++) {
        // exact octant symmetries are not needed at long-double accuracy
        const long double a = two_pi * (long double) k / (long double) M;
        cs[(size_t) k] = cosl(a);
        sn[(size_t) k] = sinl(a);
    }
    tw.resize((size_t) M);
    for (int k = 0; k < M; k++) tw[(size_t) k] = make_double2((double) cs[(size_t) k], (double) -sn[(size_t) k]);

    // g[j] = h[U*j] + i*h[U*j+1] (U==2) or h[j] (U==1), j in [-lg, lg]
    const int lg = (L + U - 1) / U;
    const double* h = s.lp.taps.data() + L; // h[-L..L]
    auto tap = [&](long long idx) -> long double {
        return (idx < -L || idx > L) ? 0.0L : (long double) h[idx];
    };
    const long double scale = 1.0L / ((long double) M * (U == 2 ? 2.0L : 1.0L));
    std::vector<double2> nat((size_t) M);
    for (int k = 0; k < M; k++) {
        long double re = 0.0L, im = 0.0L;
        for (int j = -lg; j <= lg; j++) {
            const long double gr = tap((long long) U * j);
            const long double gi = (U == 2) ? tap((long long) U * j + 1) : 0.0L;
            if (gr == 0.0L && gi == 0.0L) continue;
            const int idx = (int) ((((long long) j * k) % M + M) % M);
            const long double c = cs[(size_t) idx], sv = -sn[(size_t) idx]; // exp(-i*2pi*j*k/M)
            re += gr * c - gi * sv;
            im += gr * sv + gi * c;
        }
        nat[(size_t) k] = make_double2((double) (re * scale), (double) (im * scale));
    }
    if (s.block_exact) {
        // Power-of-two decimation in the reference = inverse transform of only the lowest 1/D of
        // the block spectrum (CDSPBlockConvolver.h:329-344).  Same thing here: the bins that the
        // shorter inverse FFT never sees are zeroed and the full-length inverse is sampled every
        // D-th point.  (The folded Nyquist term kb[z]*p[z]-kb[z+1]*p[z+1] is the product of two
        // stop-band values, far below one ulp of the output, and is dropped.)
        const int keep = M / (2 * s.down);
        if (nyq_gain) *nyq_gain = nat[(size_t) keep].x;
        for (int k = keep; k <= M - keep; k++) nat[(size_t) k] = make_double2(0.0, 0.0);
    }
    switch (fft_log2) {
    case 6: fill_slot_order<64>(nat, spec_slots); break;
    case 7: fill_slot_order<128>(nat, spec_slots); break;
    case 8: fill_slot_order<256>(nat, spec_slots); break;
    case 9: fill_slot_order<512>(nat, spec_slots); break;
    case 10: fill_slot_order<1024>(nat, spec_slots); break;
    case 11: fill_slot_order<2048>(nat, spec_slots); break;
    case 13: fill_slot_order<8192>(nat, spec_slots); break;
    default: fill_slot_order<4096>(nat, spec_slots); break;
    }
}


std::vector<double2> build_tw_tab(const std::vector<double2>& tw)
{
    std::vector<double2> tt(512);
    for (int q = 0; q < 16; q++)
        for (int r = 0; r < 16; r++) {
            tt[(size_t) (q * 16 + r)] = tw[(size_t) ((r * q) * 16)];   // W_256^(r q) = W_4096^(16 r q)
            tt[(size_t) (256 + q * 16 + r)] = tw[(size_t) (r * q)];    // W_4096^(r q)
        }
    return tt;
}

std::vector<double2> build_c_tab(const std::vector<double2>& spec, const std::vector<double2>& tw, int up)
{
    using namespace f2;
    if (up == 1) {
        // H[k]/2 for k = 0..N from the slot-ordered table of FFT(h)/M (the halving is exact)
        auto hk = [&](int k) {
            const double2 v = spec[(size_t) slot_of<FM>(k)];
            return make_double2(0.5 * v.x, 0.5 * v.y);
        };
        std::vector<double2> ct((size_t) 4 * 3 * HT + 3);
        for (int u = 0; u < 4; u++)
            for (int ht = 0; ht < HT; ht++) {
                const int k = c_freq(ht, u);
                double2* e = &ct[(size_t) (u * 3) * HT + ht];
                e[0] = tw[(size_t) k];
                e[HT] = hk(k);
                e[2 * HT] = hk(FN - k); // k = 0: the Nyquist bin
            }
        ct[(size_t) 12 * HT] = tw[(size_t) (FN / 2)];
        ct[(size_t) 12 * HT + 1] = hk(FN / 2);
        ct[(size_t) 12 * HT + 2] = hk(FN / 2);
        return ct;
    }
    return std::vector<double2>(); // up 2: phase C runs inside the first inverse pass (build_cd_tab)
}

std::vector<double2> build_cd_tab(const std::vector<double2>& spec, const std::vector<double2>& tw)
{
    using namespace f2;
    std::vector<double2> ct((size_t) 17 * HT);
    for (int g = 0; g < HT; g++) {
        for (int q3 = 0; q3 < 16; q3++) ct[(size_t) q3 * HT + g] = spec[(size_t) (16 * g + q3)];
        ct[(size_t) 16 * HT + g] = tw[(size_t) ((g >> 4) + 16 * (g & 15))];
    }
    return ct;
}

std::vector<double2> build_c_tab_v1(const std::vector<double2>& spec)
{
    constexpr int NT = 512, NC = FM / (2 * NT);
    std::vector<double2> ct((size_t) NC * 2 * NT);
    for (int u = 0; u < NC; u++)
        for (int tid = 0; tid < NT; tid++) {
            const int s1 = 16 * ((tid >> 3) + 64 * u) + (tid & 7);
            const int k = freq_of<FM>(s1);
            const int s2 = slot_of<FM>((FM - k) & (FM - 1));
            ct[(size_t) (2 * u) * NT + tid] = spec[(size_t) s1];
            ct[(size_t) (2 * u + 1) * NT + tid] = spec[(size_t) s2];
        }
    return ct;
}

FusedGeom fused_geometry(const StageDesc& s, const StageDesc& f)
{
    FusedGeom g;
    if (!(s.kind == ST_BLOCKCONV && (s.up == 2 || s.up == 1) && s.down == 1 && !s.block_exact &&
          (f.kind == ST_FRAC_WHOLE || (f.kind == ST_FRAC_POLY && s.up == 2))))
        return g;
    g.up = s.up;
    // half support of the filter as seen from one tile sample: polyphase branches for up 2 (input-rate samples)
    const int lg = s.up == 2 ? (s.lp.half_len + 1) / 2 : s.lp.half_len;
    const int flen = f.bank.filter_len, fll = flen / 2 - 1;
    int dmax = 0;
    if (f.kind == ST_FRAC_WHOLE)
        dmax = (int) (((long long) 9 * f.in_step + f.out_step - 1) / f.out_step) + 1; // up to 10 phases per group
    const int yl = (fll + 2) & ~1;
    const int yr = (dmax + flen - yl + 2 + 1) & ~1;
    // stream positions a tile can own: the valid part of its (up * 4096)-sample window minus the interpolation margins
    const int smax = (s.up * (4096 - 2 * lg) - yl - yr) & ~1;
    if (!(smax >= 1024 && (f.kind == ST_FRAC_POLY || f.in_step < smax / 2))) return g;
    g.ok = true;
    g.lg = lg;
    g.yl = yl;
    g.yr = yr;
    g.span_max = smax;
    g.ysh = 31;
    if (f.kind == ST_FRAC_WHOLE && (f.in_step & 1) == 0) {
        // lanes step by in_step doubles through the tile: make the padded stride odd
        int sh = 0;
        while (((f.in_step >> sh) & 1) == 0) sh++;
        g.ysh = sh < 4 ? 4 : sh;
        if (((f.in_step + (f.in_step >> g.ysh)) & 1) == 0) g.ysh = 31; // cannot fix; accept conflicts
    }
    return g;
}

// IR is 8 or 10, whichever spreads the phase groups more evenly over 16 warps (v1) / pairs of groups over 8 (v2).
int choose_group_ir(const StageDesc& s)
{
    int ir = 8;
    const int g8 = (s.out_step + 7) / 8, g10 = (s.out_step + 9) / 10;
    const int c8 = ((g8 + 15) / 16) * 8, c10 = ((g10 + 15) / 16) * 10;
    // measured (v1): the 10-phase variant spills registers; where 8-phase groups can start every call on a 64-byte
    // output boundary (out_step % 8 == 0: cfg 2) it loses by ~4 % despite the even task split, elsewhere the split
    // wins (cfg 3, out_step 147: 1.88 vs 1.98 ms)
    if (s.out_step % 8 != 0 && c10 < c8) ir = 10;
    if (const char* e = getenv("R8BGPU_IR")) ir = atoi(e) == 10 ? 10 : 8;
    return ir;
}

GroupBank build_group_bank(const StageDesc& s, int ir, bool frag_order)
{
    GroupBank B;
    const int os = s.out_step, flen = s.bank.filter_len;
    B.ir = ir;
    B.off.resize((size_t) os);
    B.row.resize((size_t) os);
    for (int r = 0; r < os; r++) {
        const long long pos = (long long) r * s.in_step;
        B.off[(size_t) r] = (int) (pos / os);
        B.row[(size_t) r] = (int) (pos % os);
    }
    B.n_groups = (os + ir - 1) / ir;
    // window offset of "phase" pr >= 0 counted from cycle 0 (pr >= os continues in later cycles)
    auto offx = [&](int pr) { return B.off[(size_t) (pr % os)] + (pr / os) * s.in_step; };
    int dmax = 0;
    for (int r0 = 0; r0 < os; r0++) dmax = std::max(dmax, offx(r0 + ir - 1) - offx(r0));
    B.smaxp = (flen + dmax + 3) & ~3;
    // one entry per possible first phase r0: lets a call start its groups at e0 mod 8
    B.gb.assign((size_t) os * B.smaxp * ir, 0.0);
    B.go.resize((size_t) os);
    for (int r0 = 0; r0 < os; r0++) {
        B.go[(size_t) r0] = B.off[(size_t) r0];
        for (int r = 0; r < ir; r++) {
            const int pr = r0 + r;
            const int dr = offx(pr) - offx(r0);
            const double* rowp = s.bank.table.data() + (size_t) B.row[(size_t) (pr % os)] * flen;
            for (int i = 0; i < flen; i++) {
                const int tap = dr + i;
                const size_t at = frag_order ? (size_t) (tap & ~3) * 8 + (size_t) r * 4 + (tap & 3) : (size_t) tap * ir + r;
                B.gb[(size_t) r0 * B.smaxp * ir + at] = rowp[i];
            }
        }
    }
    return B;
}

void fused_whole_fields(FusedParams& p, const StageDesc& f, long long e0, long long e1)
{
    p.mode = 0;
    p.flen = f.bank.filter_len;
    p.fll = p.flen / 2 - 1;
    p.e0 = e0;
    p.e1 = e1;
    p.in_step = f.in_step;
    p.out_step = f.out_step;
    p.p_lo = ((e0 * f.in_step) / f.out_step) & ~1LL; // even (positions are >= 0)
    p.p_hi = ((e1 - 1) * f.in_step) / f.out_step + 1;
    // groups of 8 phases start at e0 mod 8: every 64-byte output row is then aligned in the caller's buffer
    p.wrap = (f.out_step % 8 == 0 && !getenv("R8BGPU_NO_ALIGN")) ? 1 : 0;
    p.delta = p.wrap ? (int) (e0 & 7) : 0;
}

void fused2_tiles(FusedParams& p, const FusedGeom& g, int cur_parity)
{
    // One tile per half-CTA, no pairing.  Spans are multiples of 4 so that every tile's FFT window starts on the
    // same parity of the input index, and p_lo gives way by one sample pair where that makes the windows start
    // 16-byte aligned in the caller's block.
    const long long w0 = (g.up == 1 ? p.p_lo - g.yl : (p.p_lo - g.yl) / 2) - g.lg;
    const int back = g.up == 1 ? 1 : 2; // stream positions per input sample
    if (cur_parity >= 0 && (((w0 & 1) != 0) != (cur_parity != 0)) && p.p_lo >= back) p.p_lo -= back;
    const long long range = p.p_hi - p.p_lo, smax = g.span_max & ~3;
    const long long nt = (range + smax - 1) / smax;
    p.n_tiles = (int) nt;
    p.span = nt > 0 ? (int) (((range + nt - 1) / nt + 3) & ~3LL) : 4;
}

int fused2_choose_glog(int span, int in_step, int out_step, int ir)
{
    // lanes = (32 >> glog) stepping cycles x (1 << glog) phase groups, 3 cycles per lane: fewest rounds of tasks
    // over a half-CTA's 8 warps, ties to the wider cycle dimension
    const int cyc = span / in_step + 2, ng = (out_step + ir - 1) / ir;
    int best = 0, best_rounds = INT_MAX;
    for (int gl = 0; gl <= 2; gl++) {
        const int tasks = ((ng + (1 << gl) - 1) >> gl) * ((cyc + (96 >> gl) - 1) / (96 >> gl));
        const int rounds = (tasks + 7) / 8;
        if (rounds < best_rounds) {
            best_rounds = rounds;
            best = gl;
        }
    }
    if (const char* e = getenv("R8BGPU_F2_GLOG")) best = atoi(e) & 3;
    return best > 2 ? 2 : best;
}

int fused2_choose_mbu(int span, int in_step, int out_step)
{
    // a tile owns ~span / in_step + 1 stepping cycles, handled in pairs of 8-cycle blocks (16 consecutive cycles)
    const int cycles = span / in_step + 2, n_mb = 2 * ((cycles - 1) / 16 + 1), n_groups = (out_step + 7) / 8;
    int best = 3, best_cost = INT_MAX;
    for (int mbu = 2; mbu <= 4; mbu++) {
        const int units = n_groups * ((n_mb + mbu - 1) / mbu);
        const int cost = ((units + 7) / 8) * mbu;
        if (cost < best_cost || (cost == best_cost && mbu == 3)) {
            best_cost = cost;
            best = mbu;
        }
    }
    if (const char* e = getenv("R8BGPU_F2_MBU")) best = atoi(e);
    return best < 2 ? 2 : (best > 4 ? 4 : best);
}

} // namespace r8bgpu
