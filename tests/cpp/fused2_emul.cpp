// CPU emulation of the v2 fused kernel (csrc/r8b_fused2.cu), for the tests that run without a GPU.
//
// The kernel's arithmetic lives in per-thread phase functions (csrc/r8b_fused2_core.cuh) that compile for the
// host.  This harness runs those very functions "thread" after "thread", one loop per barrier interval, on
// tiles laid out by the engine's own host code (plan, schedule, tables, tile geometry), so index algebra,
// slot orders, the real-input FFT split and the interpolation bookkeeping are checked against the oracle
// before a GPU is involved.  TEST INFRASTRUCTURE: not part of the product, never linked into libr8bgpu.so.
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../r8brain-free-src_b200/csrc/r8b_fused2_core.cuh"
#include "../../r8brain-free-src_b200/csrc/r8b_hosttab.h"
#include "../../r8brain-free-src_b200/csrc/r8b_plan.h"

using namespace r8bgpu;
using namespace r8bgpu::f2;

namespace {

struct Emul {
    Plan plan;
    Schedule sched;
    std::vector<StageCall> calls;
    FusedGeom fg;
    GroupBank B;
    std::vector<double2> spec, tw, tw_tab, c_tab, cd_tab;
    std::vector<double> ring; // the whole past of the input stream (power-of-two ring, zero before the start)
    long long ring_mask = 0;
    int glog_force = -1;
    bool tc = false; // interpolation through the m8n8k4 formulation (glog_force == 8)
};

// the tensor-path interpolation: one m8n8k4 product per (block of 8 cycles, K-step), emulated on whole "warps" with the
// fragment layouts of mma.sync (A: lane = 4*row + k, B: lane = 4*n + k, C: lane = 4*row + col/2)
template <bool PADV>
void interp_tc(const FusedParams& p, const DstView& dst, const Tile& t, const double* yb, const double* sbank, const int* s_goff,
               const int* s_i, double* s_o)
{
    MmaTile mt;
    mt.load(s_i);
    const int n_groups = (p.out_step + 7) / 8;
    const int n_mu = mma_units(p, mt.c_cnt), ksteps = p.smaxp >> 2;
    for (int w = 0; w < HT / 32; w++) { // the kernel deals units to its 8 warps round-robin
        MmaUnit mu;
        mu.set(w, n_groups);
        for (int unit = w; unit < n_mu; unit += HT / 32, mu.advance(HT / 32, n_groups)) {
            const int MBU = mma_mbu(p);
            double acc[MBU_MAX][32][2] = {};
            for (int ks = 0; ks < ksteps; ks++) {
                double b[32];
                for (int lane = 0; lane < 32; lane++) b[lane] = sbank[mma_b_index(p, mu, lane) + ks * 32];
                for (int i = 0; i < MBU; i++) {
                    double a[32];
                    for (int lane = 0; lane < 32; lane++) {
                        const int yi = mma_a_index(p, mt, mu, s_goff[mu.g], i, lane) + 4 * ks;
                        a[lane] = PADV ? yb[ylay(yi, p.ysh)] : yb[yi];
                    }
                    for (int lane = 0; lane < 32; lane++) {
                        const int row = lane >> 2, col = 2 * (lane & 3);
                        for (int k = 0; k < 4; k++) {
                            acc[i][lane][0] = fma(a[4 * row + k], b[4 * col + k], acc[i][lane][0]);
                            acc[i][lane][1] = fma(a[4 * row + k], b[4 * (col + 1) + k], acc[i][lane][1]);
                        }
                    }
                }
            }
            for (int i = 0; i < MBU; i++)
                for (int lane = 0; lane < 32; lane++) mma_store(p, dst, t.ch, mt, s_o, mu, i, lane, acc[i][lane][0], acc[i][lane][1]);
        }
    }
}

template <int IR, bool PADV, int GLOG>
void run_units(const FusedParams& p, const SrcView& src, const DstView& dst, const Emul& E)
{
    std::vector<double2> buf((size_t) FPL2);
    const double2* tw2 = E.tw_tab.data();
    const double2* twf = tw2 + 256;
    const int n_groups = (p.out_step + IR - 1) / IR, esz = p.smaxp * IR;
    // this call's bank selection, as the kernel's bulk copies lay it out
    std::vector<double> sbank((size_t) n_groups * esz);
    std::vector<int> s_goff((size_t) n_groups);
    for (int g = 0; g < n_groups; g++) {
        memcpy(&sbank[(size_t) g * esz], p.gbank + (long long) (p.delta + g * IR) * esz, (size_t) esz * sizeof(double));
        s_goff[(size_t) g] = p.goff[p.delta + g * IR];
    }
    const int n_units = p.n_tiles * p.n_ch;
    for (int u = 0; u < n_units; u++) {
        const Tile t = tile_of(p, u);
        const int path = tile_input_path(src, t);
        if (path == 2) // the bulk copy
            memcpy(buf.data() + fft_pad(FN), tile_run(src, t), FM * sizeof(double));
        for (int ht = 0; ht < HT; ht++) {
            double2 v[8];
            if (path == 2) {
                for (int j = 0; j < 8; j++) v[j] = buf[(size_t) (fft_pad(FN) + ht + 256 * j)];
            } else {
                gather_tile(v, src, t, path, ht);
            }
            fwd_pass1_r8(v, buf.data(), tw2, twf, ht);
        }
        int s_i[8];
        double* s_o = nullptr;
        interp_prepare(p, dst, t, s_i, &s_o);
        for (int ht = 0; ht < FN / 16; ht++) fwd_pass<256>(buf.data(), tw2, ht);
        if (p.up == 1) for (int ht = 0; ht < FN / 16; ht++) fwd_pass<16>(buf.data(), tw2, ht);
        else for (int ht = 0; ht < FN / 16; ht++) fwd_pass16_skew(buf.data(), ht);
        if (p.up == 1) {
            std::vector<double2> z1((size_t) HT * 4), z2((size_t) HT * 4);
            for (int ht = 0; ht < HT; ht++) {
                double2 a[4], b[4];
                c_load(buf.data(), ht, a, b);
                for (int i = 0; i < 4; i++) {
                    z1[(size_t) ht * 4 + i] = a[i];
                    z2[(size_t) ht * 4 + i] = b[i];
                }
            }
            const double2 ze = buf[(size_t) fft_pad(slot_of<FN>(FN / 2))];
            for (int ht = 0; ht < HT; ht++)
                for (int i = 0; i < 4; i++) c1_pair_tab(p, buf.data(), ht, i, z1[(size_t) ht * 4 + i], z2[(size_t) ht * 4 + i]);
            c1_pair_mid(p, buf.data(), ze);
        } else { // phase C inside the first inverse pass: every "thread" fetches, then (after the barrier) computes
            std::vector<double2> z1((size_t) HT * 8), z2((size_t) HT * 8);
            for (int ht = 0; ht < HT; ht++) {
                double2 a[8], b[8];
                cd1_load(buf.data(), ht, a, b);
                for (int i = 0; i < 8; i++) {
                    z1[(size_t) ht * 8 + i] = a[i];
                    z2[(size_t) ht * 8 + i] = b[i];
                }
            }
            for (int ht = 0; ht < HT; ht++) {
                double2 a[8], b[8];
                for (int i = 0; i < 8; i++) {
                    a[i] = z1[(size_t) ht * 8 + i];
                    b[i] = z2[(size_t) ht * 8 + i];
                }
                cd1_compute(p, buf.data(), ht, a, b);
            }
        }
        if (p.up == 1) {
            for (int ht = 0; ht < FN / 16; ht++) inv_pass<16>(buf.data(), tw2, ht);
            for (int ht = 0; ht < FN / 16; ht++) inv_pass<256>(buf.data(), tw2, ht);
            std::vector<double2> v((size_t) HT * 8);
            for (int ht = 0; ht < HT; ht++) {
                double2 a[8];
                inv1_last_load(buf.data(), tw2, twf, ht, a);
                for (int i = 0; i < 8; i++) v[(size_t) ht * 8 + i] = a[i];
            }
            for (int ht = 0; ht < HT; ht++) {
                double2 a[8];
                for (int i = 0; i < 8; i++) a[i] = v[(size_t) ht * 8 + i];
                y_store1<PADV>(buf.data(), a, ht, t.w, p.ysh);
            }
        } else {
        for (int ht = 0; ht < HT; ht++) inv_pass<256>(buf.data(), tw2, ht);
        {
            std::vector<double2> v((size_t) HT * 16);
            for (int ht = 0; ht < HT; ht++) {
                double2 a[16];
                inv3_load(buf.data(), tw2, twf, ht, a);
                for (int i = 0; i < 16; i++) v[(size_t) ht * 16 + i] = a[i];
            }
            for (int ht = 0; ht < HT; ht++) {
                double2 a[16];
                for (int i = 0; i < 16; i++) a[i] = v[(size_t) ht * 16 + i];
                y_store<PADV>(buf.data(), a, ht, t.w, p.ysh);
            }
        }
        }
        if (s_i[0] > 0 && E.tc) {
            if constexpr (IR == 8) interp_tc<PADV>(p, dst, t, reinterpret_cast<const double*>(buf.data()), sbank.data(), s_goff.data(), s_i, s_o);
        } else if (s_i[0] > 0) {
            const double* yb = reinterpret_cast<const double*>(buf.data());
            const int n_tasks = TaskGeom<IR, GLOG>::n_tasks(p, s_i[1]);
            for (int task = 0; task < n_tasks; task++)
                for (int lane = 0; lane < 32; lane++) {
                    TaskGeom<IR, GLOG> g;
                    g.set(p, s_goff.data(), task, lane);
                    int yo[IQ2];
                    interp_windows<IR, GLOG>(p, g, s_i, yo);
                    double acc[IR][IQ2];
                    interp_acc<IR, PADV>(yb, sbank.data() + (size_t) g.grp * esz, yo, p.smaxp, p.ysh, acc);
                    interp_store_direct<IR, GLOG>(p, dst, t.ch, g, s_i, s_o, acc);
                }
        }
    }
}

template <int IR, bool PADV>
void run_glog(const FusedParams& p, const SrcView& src, const DstView& dst, const Emul& E)
{
    if (p.glog == 2) run_units<IR, PADV, 2>(p, src, dst, E);
    else if (p.glog == 1) run_units<IR, PADV, 1>(p, src, dst, E);
    else run_units<IR, PADV, 0>(p, src, dst, E);
}

} // namespace

extern "C" {

// A "2x BlockConvolver -> whole-stepping interpolator" resampler for ONE channel; returns NULL when the rate
// pair does not plan to that chain.
void* f2emul_create(double src, double dst, int max_in_len, double tb, double atten, int glog_force)
{
    Emul* E = new Emul;
    if (!E->plan.build(src, dst, max_in_len, tb, atten, 0, 0, 0) || E->plan.stages.size() != 2 ||
        E->plan.stages[1].kind != ST_FRAC_WHOLE) {
        delete E;
        return nullptr;
    }
    E->fg = fused_geometry(E->plan.stages[0], E->plan.stages[1]);
    if (!E->fg.ok) {
        delete E;
        return nullptr;
    }
    E->sched.init(&E->plan);
    E->tc = glog_force == 8 || E->fg.up == 1;
    E->B = build_group_bank(E->plan.stages[1], E->tc ? 8 : choose_group_ir(E->plan.stages[1]), E->tc);
    build_spectrum(E->plan.stages[0], 12, E->spec, E->tw, nullptr);
    E->tw_tab = build_tw_tab(E->tw);
    E->c_tab = build_c_tab(E->spec, E->tw, E->fg.up);
    E->cd_tab = build_cd_tab(E->spec, E->tw);
    E->ring.assign((size_t) 1 << 22, 0.0);
    E->ring_mask = ((long long) 1 << 22) - 1;
    E->glog_force = E->tc ? -1 : glog_force;
    return E;
}

void f2emul_destroy(void* h) { delete (Emul*) h; }

// One process() call: l input samples at x (any alignment), up to out_cap outputs; returns the count.
int f2emul_process(void* h, const double* x, int l, double* out, int out_cap)
{
    Emul& E = *(Emul*) h;
    const int n_out = E.sched.advance(l, E.calls);
    if (n_out > out_cap) return -1;
    const StageCall& c = E.calls[0];
    const StageCall& fc = E.calls[1];
    const StageDesc& f = E.plan.stages[1];
    if (n_out > 0) {
        FusedParams p;
        memset(&p, 0, sizeof p);
        fused_whole_fields(p, f, fc.e0, fc.e1);
        fused2_tiles(p, E.fg, (int) (c.n0 & 1));
        p.yl = E.fg.yl;
        p.lg = E.fg.lg;
        p.ysh = E.fg.ysh;
        p.spec = E.spec.data();
        p.tw = E.tw.data();
        p.c_tab = E.c_tab.data();
        p.cd_tab = E.cd_tab.data();
        p.up = E.fg.up;
        p.ylen = E.fg.up * 4096;
        p.gbank = E.B.gb.data();
        p.goff = E.B.go.data();
        p.smaxp = E.B.smaxp;
        p.ir = E.B.ir;
        p.gbank_smem_len = E.B.n_groups * E.B.smaxp * E.B.ir;
        p.n_ch = 1;
        p.mbu = fused2_choose_mbu(p.span, f.in_step, f.out_step);
        p.glog = E.tc ? 0 : E.glog_force >= 0 ? E.glog_force : fused2_choose_glog(p.span, f.in_step, f.out_step, p.ir);
        SrcView src;
        src.ring = E.ring.data();
        src.ring_stride = (long long) E.ring.size();
        src.ring_mask = E.ring_mask;
        src.cur = x;
        src.cur_stride = l;
        src.cur_base = c.n0;
        src.avail = c.n1;
        DstView dst;
        dst.ptr = out;
        dst.stride = out_cap;
        dst.mask = -1;
        dst.base = fc.e0;
        const bool pad = p.ysh != 31;
        if (p.ir == 10) {
            if (pad) run_glog<10, true>(p, src, dst, E);
            else run_glog<10, false>(p, src, dst, E);
        } else {
            if (pad) run_glog<8, true>(p, src, dst, E);
            else run_glog<8, false>(p, src, dst, E);
        }
    }
    for (int i = 0; i < l; i++) E.ring[(size_t) ((c.n0 + i) & E.ring_mask)] = x[i];
    return n_out;
}

} // extern "C"
