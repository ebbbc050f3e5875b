// r8b_capi.cu -- the engine behind include/r8bgpu.h: device-resident per-channel state, the
// per-call launch sequence, and the extern "C" boundary.
//
// Per-channel state in HBM, all planar [channel][...]:
//   * input history ring      : the most recent source samples the first stage may re-read
//   * one ring per stage link : the stream between stage i and i+1, absolute-indexed,
//                               capacity = pow2 >= (max samples per call + look-back of stage i+1)
// Shared read-only per plan: filter spectrum (slot order, pre-scaled), FFT twiddles, fractional
// delay bank.  All integer scheduling state lives on the host (Schedule), once per batch.
#include "../../include/r8bgpu.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "r8b_fft.cuh"
#include "r8b_hosttab.h"
#include "r8b_kernels.h"
#include "r8b_multi.h"
#include "r8b_plan.h"

using namespace r8bgpu;

namespace {

thread_local std::string g_err;

void set_err(const std::string& s) { g_err = s; }

bool cuda_ok(cudaError_t e, const char* what)
{
    if (e == cudaSuccess) return true;
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
}

long long next_pow2(long long v)
{
    long long p = 1;
    while (p < v) p <<= 1;
    return p;
}

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) {
            ok = false;
            return;
        }
        if (prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard()
    {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

int choose_fft_log2(int lg, int max_log2)
{
    if (const char* e = getenv("R8BGPU_FFT_LOG2")) {
        const int v = atoi(e);
        if (v >= 10 && v <= max_log2 && (1 << v) - 2 * lg >= 64) return v;
    }
    int best = -1;
    double best_cost = 0.0;
    for (int b = 10; b <= max_log2; b++) {
        const int m = 1 << b;
        const int valid = m - 2 * lg;
        if (valid < 64) continue;
        const double cost = (double) b * m / valid;
        if (best < 0 || cost < best_cost) {
            best = b;
            best_cost = cost;
        }
    }
    return best;
}

struct StageDev {
    // BLOCKCONV
    int fft_log2 = 0, lg = 0, virt_up = 1;
    double nyq_gain = 0.0;
    double2* spec = nullptr;
    double2* tw = nullptr;
    // FRAC
    double* bank = nullptr;
    // source ring of this stage (for stage 0: the input history ring)
    double* ring = nullptr;
    long long ring_cap = 0;
    // fusion: a 2x BLOCKCONV immediately followed by a FRAC stage runs as ONE kernel
    bool fused_with_next = false; // on the BLOCKCONV stage
    bool fused_into_prev = false; // on the FRAC stage (its source ring is never materialised)
    int* phase_off = nullptr;
    int* phase_row = nullptr;
    double* gbank = nullptr; // whole stepping: grouped, pre-shifted, zero-padded bank (see FusedParams)
    int* goff = nullptr;
    int gbank_len = 0, gbank_smem_len = 0, smaxp = 0, ir = 8;
    int yl = 0, yr = 0, ysh = 31, span_max = 0, bank_in_smem = 0;
    // R8B_FASTTIMING position tables (FRAC_POLY stages of fast-timing plans)
    int* ft_dp = nullptr;
    double* ft_fpos = nullptr;
    int* h_dp[2] = {nullptr, nullptr};
    double* h_fpos[2] = {nullptr, nullptr};
    cudaEvent_t ft_ev[2] = {nullptr, nullptr};
    int ft_cur = 0;
    int casc_len = 0; // >= 2 on the first stage of a run of HBUP stages executed by k_hbup_cascade
    int down_casc_len = 0; // >= 2 on the first stage of a run of HBDOWN stages executed by k_hbdown_cascade
    HbDownCascParams down_casc; // its tile plan (taps, halos, shared-memory layout)
    int down_casc_smem = 0;
    // v2 fused kernel (r8b_fused2.cu): [q][r] twiddle tables for the bulk copy; on the BLOCKCONV stage
    double2* tw_tab = nullptr;
    double2* c_tab = nullptr;   // v2 fused kernel: phase C operands in thread order
    double2* cd_tab = nullptr;   // v2 fused kernel, up 2: operands of phase C fused into the first inverse pass
    double2* c_tab_v1 = nullptr; // round-1 fused kernel: its two spectrum values per frequency pair in thread order
    bool bank_frag_order = false; // grouped bank stored in mma fragment order (only the tensor-path interpolation reads it)
    bool f2_ok = false;
    bool f2_poly = false; // order-2 interpolator on the v2 kernel's tensor path (decided per call: near-integer ratios)
    bool f2_copy = false; // BlockConvolver 2/1 alone on the v2 kernel (phase E copies the 2x stream out)
    FusedGeom fgeom;
};

} // namespace

struct r8bgpu_plan {
    Plan p;
};

// A multi-device batch (r8bgpu_batch_create(plan, n, -1) on a box with several GPUs) is a FRONT: it owns no device
// state, only one ordinary single-device batch per shard (contiguous channel ranges) and the worker threads that
// drive them side by side (r8b_multi.h).
struct ShardFront {
    std::vector<r8bgpu_batch*> shards;
    std::vector<int> ch0, device, numa;
    std::unique_ptr<ShardPool> pool;
};

struct r8bgpu_batch {
    std::unique_ptr<ShardFront> front; // non-null: multi-device front (everything below except plan/n_ch is unused)
    const Plan* plan = nullptr;
    Plan plan_copy; // batches own a copy so the plan handle may be destroyed first
    int n_ch = 0;
    int device = 0;
    cudaStream_t stream = nullptr;
    Schedule sched;
    std::vector<StageDev> dev;
    std::vector<StageCall> calls;
    unsigned long long launches = 0;
    unsigned long long dev_bytes = 0;
    // optional per-stage device timing (CUDA events on the launch stream)
    bool timing = false;
    struct EvPair {
        int stage;
        cudaEvent_t a, b;
    };
    std::vector<EvPair> events;
    std::vector<double> stage_ms;
    std::vector<unsigned long long> stage_launches;
    // staging + pipeline resources for the host-pointer entry point
    double* st_in = nullptr;
    double* st_out = nullptr;
    unsigned char* raw_in = nullptr;   // narrow-format staging (r8b_format.cu)
    unsigned char* raw_out = nullptr;
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr, s_comp = nullptr;
    int host_groups = 1;
    std::vector<cudaEvent_t> ev_h2d, ev_k;
    unsigned long long* prof = nullptr; // R8BGPU_PROFILE: phase cycle counters of the fused kernel
    int n_sm = 0;     // SMs of the device (grid of the persistent v2 fused kernel)
    int f2_flags = 6; // v2 fused kernel: bit 0 ping-pong token, bit 1 bulk-copied input tiles, bit 2 interpolation on the fp64 tensor path
    unsigned long long prof_ctas = 0;

    ~r8bgpu_batch()
    {
        if (front) {
            front->pool.reset();
            for (r8bgpu_batch* sb : front->shards) delete sb;
            return;
        }
        DeviceGuard g(device);
        if (prof != nullptr) {
            unsigned long long h[10] = {};
            cudaDeviceSynchronize();
            cudaMemcpy(h, prof, sizeof h, cudaMemcpyDeviceToHost);
            static const char* nm[8] = {"gather+fwd1", "fwd2", "fwd3", "C(split*G)", "inv1", "inv2", "inv3+ystore", "interp"};
            unsigned long long tot = 0;
            for (int i = 0; i < 8; i++) tot += h[i];
            fprintf(stderr, "[r8bgpu profile] k_up2_frac phases, mean clk per CTA over %llu CTAs:\n", prof_ctas);
            for (int i = 0; i < 8; i++)
                fprintf(stderr, "  %-12s %9.0f  (%4.1f %%)\n", nm[i], prof_ctas ? (double) h[i] / prof_ctas : 0.0,
                        tot ? 100.0 * h[i] / tot : 0.0);
            if (h[8] + h[9] > 0)
                fprintf(stderr, "  order-2 bank, clk per CTA: 4-output groups %.0f, queued single outputs %.0f\n",
                        (double) h[8] / prof_ctas, (double) h[9] / prof_ctas);
            cudaFree(prof);
        }
        for (auto& d : dev) {
            cudaFree(d.spec);
            cudaFree(d.tw);
            cudaFree(d.tw_tab);
            cudaFree(d.c_tab);
            cudaFree(d.c_tab_v1);
            cudaFree(d.cd_tab);
            cudaFree(d.bank);
            cudaFree(d.ring);
            cudaFree(d.phase_off);
            cudaFree(d.phase_row);
            cudaFree(d.gbank);
            cudaFree(d.goff);
            cudaFree(d.ft_dp);
            cudaFree(d.ft_fpos);
            for (int k = 0; k < 2; k++) {
                if (d.h_dp[k]) cudaFreeHost(d.h_dp[k]);
                if (d.h_fpos[k]) cudaFreeHost(d.h_fpos[k]);
                if (d.ft_ev[k]) cudaEventDestroy(d.ft_ev[k]);
            }
        }
        cudaFree(st_in);
        cudaFree(st_out);
        cudaFree(raw_in);
        cudaFree(raw_out);
        for (auto e : ev_h2d) cudaEventDestroy(e);
        for (auto e : ev_k) cudaEventDestroy(e);
        if (s_h2d) cudaStreamDestroy(s_h2d);
        if (s_d2h) cudaStreamDestroy(s_d2h);
        if (s_comp) cudaStreamDestroy(s_comp);
    }
};

extern "C" {

const char* r8bgpu_last_error(void) { return g_err.c_str(); }
const char* r8bgpu_version(void) { return "r8bgpu 0.1 (sm_100a)"; }

r8bgpu_plan* r8bgpu_plan_create(double src, double dst, int max_in_len, double tb, double atten, int phase,
                                int extfft, int fasttiming)
{
    std::unique_ptr<r8bgpu_plan> h(new r8bgpu_plan);
    if (!h->p.build(src, dst, max_in_len, tb, atten, phase, extfft, fasttiming)) {
        set_err("plan_create: " + h->p.error);
        return nullptr;
    }
    return h.release();
}

r8bgpu_plan* r8bgpu_plan_create_stage(int kind, const double* params, int n_params, int max_in_len, int extfft)
{
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n_params && i < 8; i++) a[i] = params[i];
    std::unique_ptr<r8bgpu_plan> h(new r8bgpu_plan);
    if (max_in_len <= 0 || !h->p.build_single(kind, a, max_in_len, extfft)) {
        set_err("plan_create_stage: " + h->p.error);
        return nullptr;
    }
    return h.release();
}

void r8bgpu_plan_destroy(r8bgpu_plan* plan) { delete plan; }
int r8bgpu_plan_max_out_len(const r8bgpu_plan* plan) { return plan->p.max_out_len; }
int r8bgpu_plan_in_len_before_out_pos(const r8bgpu_plan* plan, int pos) { return plan->p.in_len_before_out_pos(pos); }
int r8bgpu_plan_input_required_for_output(const r8bgpu_plan* plan, int n) { return plan->p.input_required_for_output(n); }
double r8bgpu_plan_latency_frac(const r8bgpu_plan*) { return 0.0; } // linear-phase chains leave no fractional latency
int r8bgpu_plan_is_passthrough(const r8bgpu_plan* plan) { return plan->p.passthrough ? 1 : 0; }
int r8bgpu_plan_stage_count(const r8bgpu_plan* plan) { return (int) plan->p.stages.size(); }

static int stage_data_len(const StageDesc& s)
{
    switch (s.kind) {
    case ST_BLOCKCONV: return s.lp.kernel_len;
    case ST_FRAC_WHOLE:
    case ST_FRAC_POLY: return (int) s.bank.table.size();
    default: return s.hb_taps;
    }
}

int r8bgpu_plan_stage_info(const r8bgpu_plan* plan, int stage, r8bgpu_stage_info* info)
{
    if (stage < 0 || stage >= (int) plan->p.stages.size() || info == nullptr) {
        set_err("stage_info: bad stage index");
        return -1;
    }
    const StageDesc& s = plan->p.stages[(size_t) stage];
    memset(info, 0, sizeof *info);
    info->kind = (int) s.kind;
    info->up = s.up;
    info->down = s.down;
    info->max_out_len = s.max_out_len;
    info->data_len = stage_data_len(s);
    switch (s.kind) {
    case ST_BLOCKCONV:
        info->kernel_len = s.lp.kernel_len;
        info->latency = s.latency;
        info->ref_input_len = s.ref_input_len;
        info->block_len_bits = s.lp.block_len_bits;
        break;
    case ST_FRAC_WHOLE:
    case ST_FRAC_POLY:
        info->kernel_len = s.bank.filter_len;
        info->fracs = s.bank.fracs;
        info->in_step = s.in_step;
        info->out_step = s.out_step;
        info->order = s.bank.order;
        info->atten = s.bank.atten;
        break;
    default:
        info->kernel_len = s.hb_taps;
        info->atten = s.hb_atten;
        break;
    }
    return 0;
}

int r8bgpu_plan_stage_data(const r8bgpu_plan* plan, int stage, double* out, int cap)
{
    if (stage < 0 || stage >= (int) plan->p.stages.size()) {
        set_err("stage_data: bad stage index");
        return -1;
    }
    const StageDesc& s = plan->p.stages[(size_t) stage];
    const double* src = s.kind == ST_BLOCKCONV ? s.lp.taps.data()
        : (s.kind == ST_FRAC_WHOLE || s.kind == ST_FRAC_POLY) ? s.bank.table.data() : s.hb.data();
    const int n = stage_data_len(s);
    const int c = n < cap ? n : cap;
    if (out != nullptr && c > 0) memcpy(out, src, (size_t) c * sizeof(double));
    return n;
}

int r8bgpu_plan_describe(const r8bgpu_plan* plan, char* buf, int cap)
{
    const std::string d = plan->p.describe();
    if (buf != nullptr && cap > 0) {
        const size_t n = d.size() < (size_t) cap - 1 ? d.size() : (size_t) cap - 1;
        memcpy(buf, d.data(), n);
        buf[n] = 0;
    }
    return (int) d.size();
}

int r8bgpu_plan_simulate(const r8bgpu_plan* plan, const int* lens, int n_calls, int* counts)
{
    Schedule sc;
    sc.init(&plan->p);
    std::vector<StageCall> calls;
    for (int i = 0; i < n_calls; i++) {
        if (lens[i] < 0 || lens[i] > plan->p.max_in_len) {
            set_err("simulate: block length outside [0, MaxInLen]");
            return -1;
        }
        counts[i] = sc.advance(lens[i], calls);
    }
    return 0;
}

// ------------------------------------------------------------------------------------------

int r8bgpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

static r8bgpu_batch* create_front(const r8bgpu_plan* plan, int n_channels, int n_sh, int ndev)
{
    std::unique_ptr<r8bgpu_batch> b(new r8bgpu_batch);
    b->plan_copy = plan->p;
    b->plan = &b->plan_copy;
    b->n_ch = n_channels;
    b->device = -1;
    b->front.reset(new ShardFront);
    ShardFront& F = *b->front;
    const int per = (n_channels + n_sh - 1) / n_sh; // ceil(C / G) channels per shard, the last one takes the rest
    for (int s = 0; s * per < n_channels; s++) {
        const int c0 = s * per, n = std::min(per, n_channels - c0), dev = s % ndev;
        r8bgpu_batch* sb = r8bgpu_batch_create(plan, n, dev);
        if (sb == nullptr) return nullptr; // (b's destructor releases the shards made so far)
        F.shards.push_back(sb);
        F.ch0.push_back(c0);
        F.device.push_back(dev);
        F.numa.push_back(gpu_numa_node(dev));
    }
    F.pool.reset(new ShardPool(F.numa));
    return b.release();
}

// run fn(shard batch, shard index) on every shard's worker thread; all shards see the same block lengths, so they
// return the same count.  Any failure fails the call (the shards' schedules then disagree: the caller must clear()).
static int front_run(r8bgpu_batch* b, const std::function<int(r8bgpu_batch*, int)>& fn)
{
    ShardFront& F = *b->front;
    std::vector<std::string> errs;
    const std::function<int(int)> job = [&](int s) { return fn(F.shards[(size_t) s], s); };
    const std::function<std::string()> err = [] { return g_err; };
    const std::vector<int> r = F.pool->run_all(job, &errs, err);
    for (size_t s = 0; s < r.size(); s++)
        if (r[s] < 0) {
            set_err("shard " + std::to_string(s) + " (device " + std::to_string(F.device[s]) + "): " + errs[s]);
            return -1;
        }
    for (size_t s = 1; s < r.size(); s++)
        if (r[s] != r[0]) {
            set_err("multi-device batch: shards disagree on the output count");
            return -1;
        }
    return r.empty() ? 0 : r[0];
}

r8bgpu_batch* r8bgpu_batch_create(const r8bgpu_plan* plan, int n_channels, int device)
{
    if (plan == nullptr || n_channels <= 0 || n_channels > 65535) {
        set_err("batch_create: need a plan and 1..65535 channels");
        return nullptr;
    }
    int ndev = 0;
    if (!cuda_ok(cudaGetDeviceCount(&ndev), "batch_create: cudaGetDeviceCount") || ndev == 0) {
        if (g_err.empty()) set_err("batch_create: no CUDA device (this engine has no CPU fallback)");
        return nullptr;
    }
    if (device == R8BGPU_DEVICE_ALL) {
        // shard over every visible device (R8BGPU_FORCE_SHARDS=n: n shards dealt round-robin to the devices -- lets a
        // one-GPU box exercise the multi-device path)
        int n_sh = ndev;
        if (const char* e = getenv("R8BGPU_FORCE_SHARDS")) n_sh = atoi(e) > 0 ? atoi(e) : ndev;
        if (n_sh > n_channels) n_sh = n_channels;
        if (n_sh > 1) return create_front(plan, n_channels, n_sh, ndev);
        device = 0;
        if (ndev > 1 && !cuda_ok(cudaGetDevice(&device), "batch_create: cudaGetDevice")) return nullptr;
    }
    if (device < 0 && !cuda_ok(cudaGetDevice(&device), "batch_create: cudaGetDevice")) return nullptr;
    if (device >= ndev) {
        set_err("batch_create: device index out of range");
        return nullptr;
    }
    DeviceGuard g(device);
    if (!g.ok) {
        set_err("batch_create: cudaSetDevice failed");
        return nullptr;
    }
    std::unique_ptr<r8bgpu_batch> b(new r8bgpu_batch);
    b->plan_copy = plan->p;
    b->plan = &b->plan_copy;
    b->n_ch = n_channels;
    b->device = device;
    b->sched.init(b->plan);
    if (!cuda_ok(cudaDeviceGetAttribute(&b->n_sm, cudaDevAttrMultiProcessorCount, device), "batch_create: SM count")) return nullptr;
    if (const char* e = getenv("R8BGPU_F2_FLAGS")) b->f2_flags = atoi(e);
    const auto& st = b->plan->stages;
    b->dev.resize(st.size());
    for (size_t i = 0; i < st.size(); i++) {
        const StageDesc& s = st[i];
        StageDev& d = b->dev[i];
        const long long emit_in = (i == 0) ? 0 : st[i - 1].max_out_len;
        // Fusable pair: [BlockConv 2/1 with a kernel that fits M=4096 tiles] -> [FracInterp].
        if (i + 1 < st.size() && !getenv("R8BGPU_NO_FUSION")) {
            FusedGeom fg = fused_geometry(s, st[i + 1]);
            if (fg.ok && fg.up == 1) {
                // the 1x pair exists only in the v2 kernel with the tensor-path interpolation: its bank must fit
                const GroupBank tb = build_group_bank(st[i + 1], 8, true);
                if (getenv("R8BGPU_FUSED_V1") || !(b->f2_flags & 4) || tb.n_groups > 192 ||
                    fused2_smem_bytes(tb.n_groups * tb.smaxp * tb.ir, false) > kFused2SmemMax)
                    fg.ok = false;
            }
            if (fg.ok) {
                // (opt-in: measured slower than the round-1 kernel's staged-row path)
                if (st[i + 1].kind == ST_FRAC_POLY && fg.up == 2 && (b->f2_flags & 4) && !getenv("R8BGPU_FUSED_V1") && getenv("R8BGPU_POLY_V2") &&
                    (st[i + 1].bank.filter_len & 1) == 0 && st[i + 1].bank.filter_len <= 32)
                    d.f2_poly = true;
                d.fused_with_next = true;
                b->dev[i + 1].fused_into_prev = true;
                d.fgeom = fg;
                d.yl = fg.yl;
                d.yr = fg.yr;
                d.span_max = fg.span_max;
                d.ysh = fg.ysh;
            }
        }
        if (s.kind == ST_HBUP && !d.fused_into_prev && !getenv("R8BGPU_NO_FUSION")) {
            size_t c = 1;
            while (i + c < st.size() && st[i + c].kind == ST_HBUP && c < 6) c++;
            if (c >= 2) {
                d.casc_len = (int) c;
                for (size_t k = 1; k < c; k++) b->dev[i + k].fused_into_prev = true;
            }
        }
        long long extra_history = 0;
        if (s.kind == ST_HBDOWN && !d.fused_into_prev && !getenv("R8BGPU_NO_FUSION")) {
            size_t c = 1;
            while (i + c < st.size() && st[i + c].kind == ST_HBDOWN && c < 6) c++;
            if (c >= 2) {
                HbDownCascParams& cp = d.down_casc;
                memset(&cp, 0, sizeof cp);
                cp.n_stages = (int) c;
                for (size_t k = 0; k < c; k++) {
                    cp.ntaps[k] = st[i + k].hb_taps;
                    for (int j = 0; j < st[i + k].hb_taps; j++) cp.taps[k][j] = st[i + k].hb[(size_t) j];
                }
                int hbd_budget = 6400; // doubles per CTA (4 CTAs per SM; measured on 2822400->44100: 3200 0.97, 6400 0.46, 12800 0.53, 25000 0.79 ms)
                if (const char* e = getenv("R8BGPU_HBD_SMEM_DOUBLES")) hbd_budget = atoi(e);
                d.down_casc_smem = hbdown_cascade_plan(cp, hbd_budget);
                if (d.down_casc_smem > 0) {
                    d.down_casc_len = (int) c;
                    for (size_t k = 1; k < c; k++) b->dev[i + k].fused_into_prev = true;
                    // the cascade recomputes intermediate samples of earlier calls from the source: keep its whole reach
                    extra_history = 2LL * cp.back[0] + (2LL << c) + 64;
                }
            }
        }
        if (d.fused_into_prev) {
            d.ring_cap = 0; // the link stream lives only in shared memory
        } else {
            d.ring_cap = next_pow2(std::max<long long>(s.src_history, extra_history) + emit_in + 64);
            const size_t ring_bytes = (size_t) d.ring_cap * (size_t) n_channels * sizeof(double);
            if (!cuda_ok(cudaMalloc(&d.ring, ring_bytes), "batch_create: cudaMalloc(ring)")) return nullptr;
            b->dev_bytes += ring_bytes;
        }
        if (s.kind == ST_BLOCKCONV) {
            // up-factors other than 1 and 2 (the planner only makes 3) run as a 1x convolution over the
            // zero-stuffed stream, exactly as the reference does
            d.virt_up = (s.up > 2) ? s.up : 1;
            const int up_eff = (s.up > 2) ? 1 : s.up;
            d.lg = (s.lp.half_len + up_eff - 1) / up_eff;
            d.fft_log2 = choose_fft_log2(d.lg, up_eff == 1 ? 13 : 12);
            if (s.block_exact) { // tiles == the reference's own blocks
                d.lg = s.ref_prev_len - s.lp.half_len;
                d.fft_log2 = s.lp.block_len_bits + 1;
                // the tile IS the reference's block (2 << BlockLenBits): 64 .. 8192 points for 1x stages (short kernels run on
                // plain radix-2 transforms); 2x stages have tiles of 1024 .. 4096
                const int lo = up_eff == 1 ? 6 : 10, hi = up_eff == 1 ? 13 : 12;
                if (d.fft_log2 < lo || d.fft_log2 > hi) {
                    char msg[200];
                    snprintf(msg, sizeof msg, "batch_create: reference-exact decimation needs a %d-point block transform; "
                             "this build has %d..%d points for such stages", 1 << d.fft_log2, 1 << lo, 1 << hi);
                    set_err(msg);
                    return nullptr;
                }
            }
            // a 2x BlockConvolver that no interpolator follows runs on the v2 fused kernel too (its phase E copies the stream
            // out) when the polyphase branches fit 4096-point tiles
            if (!d.fused_with_next && s.up == 2 && s.down == 1 && !s.block_exact && !getenv("R8BGPU_NO_FUSION") &&
                !getenv("R8BGPU_FUSED_V1") && 2 * (4096 - 2 * d.lg) >= 2048) {
                d.f2_copy = true;
                d.fgeom = FusedGeom();
                d.fgeom.ok = true;
                d.fgeom.up = 2;
                d.fgeom.lg = d.lg;
                d.fgeom.span_max = (2 * (4096 - 2 * d.lg)) & ~3;
                d.fft_log2 = 12;
            }
            if (d.fused_with_next) d.fft_log2 = 12; // the fused kernel is built for M = 4096
            if (d.fft_log2 < 0) {
                set_err("batch_create: low-pass kernel too long for the in-shared-memory FFT tiles");
                return nullptr;
            }
            std::vector<double2> spec, tw;
            build_spectrum(s, d.fft_log2, spec, tw, &d.nyq_gain);
            const size_t nb = spec.size() * sizeof(double2);
            if (!cuda_ok(cudaMalloc(&d.spec, nb), "cudaMalloc(spec)")) return nullptr;
            if (!cuda_ok(cudaMalloc(&d.tw, nb), "cudaMalloc(tw)")) return nullptr;
            if (!cuda_ok(cudaMemcpy(d.spec, spec.data(), nb, cudaMemcpyHostToDevice), "copy spec")) return nullptr;
            if (!cuda_ok(cudaMemcpy(d.tw, tw.data(), nb, cudaMemcpyHostToDevice), "copy tw")) return nullptr;
            b->dev_bytes += 2 * nb;
            if (d.fused_with_next || d.f2_copy) { // conflict-free [q][r] twiddle tables, one 8 KB bulk copy per CTA in the v2 kernel
                const std::vector<double2> tt = build_tw_tab(tw);
                if (!cuda_ok(cudaMalloc(&d.tw_tab, tt.size() * sizeof(double2)), "cudaMalloc(tw_tab)")) return nullptr;
                if (!cuda_ok(cudaMemcpy(d.tw_tab, tt.data(), tt.size() * sizeof(double2), cudaMemcpyHostToDevice), "copy tw_tab")) return nullptr;
                if (d.fgeom.up == 1) {
                    const std::vector<double2> ctb = build_c_tab(spec, tw, 1);
                    if (!cuda_ok(cudaMalloc(&d.c_tab, ctb.size() * sizeof(double2)), "cudaMalloc(c_tab)")) return nullptr;
                    if (!cuda_ok(cudaMemcpy(d.c_tab, ctb.data(), ctb.size() * sizeof(double2), cudaMemcpyHostToDevice), "copy c_tab")) return nullptr;
                }
                if (d.fgeom.up == 2) {
                    const std::vector<double2> cd = build_cd_tab(spec, tw);
                    if (!cuda_ok(cudaMalloc(&d.cd_tab, cd.size() * sizeof(double2)), "cudaMalloc(cd_tab)")) return nullptr;
                    if (!cuda_ok(cudaMemcpy(d.cd_tab, cd.data(), cd.size() * sizeof(double2), cudaMemcpyHostToDevice), "copy cd_tab")) return nullptr;
                }
                if (d.fused_with_next && d.fgeom.up == 2) {
                    const std::vector<double2> c1 = build_c_tab_v1(spec);
                    if (!cuda_ok(cudaMalloc(&d.c_tab_v1, c1.size() * sizeof(double2)), "cudaMalloc(c_tab_v1)")) return nullptr;
                    if (!cuda_ok(cudaMemcpy(d.c_tab_v1, c1.data(), c1.size() * sizeof(double2), cudaMemcpyHostToDevice), "copy c_tab_v1")) return nullptr;
                }
            }
        } else if (s.kind == ST_FRAC_WHOLE || s.kind == ST_FRAC_POLY) {
            const size_t nb = s.bank.table.size() * sizeof(double);
            if (!cuda_ok(cudaMalloc(&d.bank, nb), "cudaMalloc(bank)")) return nullptr;
            if (!cuda_ok(cudaMemcpy(d.bank, s.bank.table.data(), nb, cudaMemcpyHostToDevice), "copy bank")) return nullptr;
            b->dev_bytes += nb;
            if (s.kind == ST_FRAC_POLY && s.fasttiming) {
                const size_t cap = (size_t) s.max_out_len + 16;
                if (!cuda_ok(cudaMalloc(&d.ft_dp, cap * sizeof(int)), "cudaMalloc(ft)")) return nullptr;
                if (!cuda_ok(cudaMalloc(&d.ft_fpos, cap * sizeof(double)), "cudaMalloc(ft)")) return nullptr;
                for (int k = 0; k < 2; k++) {
                    if (!cuda_ok(cudaMallocHost(&d.h_dp[k], cap * sizeof(int)), "cudaMallocHost(ft)")) return nullptr;
                    if (!cuda_ok(cudaMallocHost(&d.h_fpos[k], cap * sizeof(double)), "cudaMallocHost(ft)")) return nullptr;
                    cudaEventCreateWithFlags(&d.ft_ev[k], cudaEventDisableTiming);
                }
            }
            if (s.kind == ST_FRAC_WHOLE) {
                // per output phase r: floor(r*InStep/OutStep) and the bank row (r*InStep) % OutStep; grouped bank for
                // the fused kernels: IR consecutive phases share one y window
                // (the tensor-path interpolation of the v2 kernel works on groups of exactly 8 phases)
                const bool want_f2 = d.fused_into_prev && i > 0 && !getenv("R8BGPU_FUSED_V1");
                bool tc_bank = want_f2 && (b->f2_flags & 4);
                GroupBank B = build_group_bank(s, tc_bank ? 8 : choose_group_ir(s), tc_bank);
                auto f2_fits = [&](const GroupBank& gb) {
                    return fused2_smem_bytes(gb.n_groups * gb.smaxp * gb.ir, false) <= kFused2SmemMax && gb.n_groups <= 192;
                };
                if (tc_bank && !f2_fits(B)) { // the v2 kernel will not run this pair: the v1 kernel reads the plain layout
                    tc_bank = false;
                    B = build_group_bank(s, choose_group_ir(s), false);
                }
                d.bank_frag_order = tc_bank;
                const size_t tb = B.off.size() * sizeof(int);
                if (!cuda_ok(cudaMalloc(&d.phase_off, tb), "cudaMalloc(phase)")) return nullptr;
                if (!cuda_ok(cudaMalloc(&d.phase_row, tb), "cudaMalloc(phase)")) return nullptr;
                cudaMemcpy(d.phase_off, B.off.data(), tb, cudaMemcpyHostToDevice);
                cudaMemcpy(d.phase_row, B.row.data(), tb, cudaMemcpyHostToDevice);
                d.gbank_smem_len = B.n_groups * B.smaxp * B.ir;
                d.ir = B.ir;
                d.gbank_len = (int) B.gb.size();
                d.smaxp = B.smaxp;
                if (!cuda_ok(cudaMalloc(&d.gbank, B.gb.size() * sizeof(double)), "cudaMalloc(gbank)")) return nullptr;
                if (!cuda_ok(cudaMalloc(&d.goff, B.go.size() * sizeof(int)), "cudaMalloc(goff)")) return nullptr;
                cudaMemcpy(d.gbank, B.gb.data(), B.gb.size() * sizeof(double), cudaMemcpyHostToDevice);
                cudaMemcpy(d.goff, B.go.data(), B.go.size() * sizeof(int), cudaMemcpyHostToDevice);
                b->dev_bytes += B.gb.size() * sizeof(double);
                d.bank_in_smem = (fused_smem_bytes(d.gbank_smem_len) <= 220 * 1024) ? 1 : 0;
                // the persistent two-pipeline kernel needs the call's whole bank in shared memory
                if (want_f2 && f2_fits(B))
                    b->dev[i - 1].f2_ok = true;
            }
        }
    }
    r8bgpu_batch* raw = b.release();
    if (r8bgpu_batch_clear(raw) != 0) {
        delete raw;
        return nullptr;
    }
    return raw;
}

void r8bgpu_batch_destroy(r8bgpu_batch* batch) { delete batch; }
int r8bgpu_batch_channels(const r8bgpu_batch* b) { return b->n_ch; }
unsigned long long r8bgpu_batch_kernel_launches(const r8bgpu_batch* b)
{
    if (!b->front) return b->launches;
    unsigned long long n = 0;
    for (const r8bgpu_batch* sb : b->front->shards) n += sb->launches;
    return n;
}
unsigned long long r8bgpu_batch_device_bytes(const r8bgpu_batch* b)
{
    if (!b->front) return b->dev_bytes;
    unsigned long long n = 0;
    for (const r8bgpu_batch* sb : b->front->shards) n += sb->dev_bytes;
    return n;
}
int r8bgpu_batch_shard_count(const r8bgpu_batch* b) { return b->front ? (int) b->front->shards.size() : 1; }
int r8bgpu_batch_shard_info(const r8bgpu_batch* b, int shard, int* device, int* first_channel, int* n_channels, int* numa_node)
{
    if (shard < 0 || shard >= r8bgpu_batch_shard_count(b)) {
        set_err("batch_shard_info: shard index out of range");
        return -1;
    }
    if (!b->front) {
        if (device) *device = b->device;
        if (first_channel) *first_channel = 0;
        if (n_channels) *n_channels = b->n_ch;
        if (numa_node) *numa_node = gpu_numa_node(b->device);
        return 0;
    }
    const ShardFront& F = *b->front;
    if (device) *device = F.device[(size_t) shard];
    if (first_channel) *first_channel = F.ch0[(size_t) shard];
    if (n_channels) *n_channels = F.shards[(size_t) shard]->n_ch;
    if (numa_node) *numa_node = F.numa[(size_t) shard];
    return 0;
}
r8bgpu_batch* r8bgpu_batch_shard(r8bgpu_batch* b, int shard)
{
    if (shard < 0 || shard >= r8bgpu_batch_shard_count(b)) {
        set_err("batch_shard: shard index out of range");
        return nullptr;
    }
    return b->front ? b->front->shards[(size_t) shard] : b;
}

void* r8bgpu_batch_host_alloc(const r8bgpu_batch* b, size_t samples_per_channel, int sample_bytes)
{
    if (b == nullptr || sample_bytes <= 0 || samples_per_channel == 0) {
        set_err("batch_host_alloc: bad arguments");
        return nullptr;
    }
    const size_t row = samples_per_channel * (size_t) sample_bytes;
    std::vector<NumaRange> ranges;
    const int n_sh = r8bgpu_batch_shard_count(b);
    for (int s = 0; s < n_sh; s++) {
        int c0 = 0, n = 0, node = -1;
        r8bgpu_batch_shard_info(b, s, nullptr, &c0, &n, &node);
        ranges.push_back({(size_t) c0 * row, (size_t) n * row, node});
    }
    void* p = numa_host_alloc(row * (size_t) b->n_ch, ranges);
    if (p == nullptr) set_err("batch_host_alloc: mmap / cudaHostRegister failed");
    return p;
}

int r8bgpu_batch_set_timing(r8bgpu_batch* b, int enable)
{
    if (b->front) {
        int rc = 0;
        for (r8bgpu_batch* sb : b->front->shards) rc |= r8bgpu_batch_set_timing(sb, enable);
        return rc;
    }
    DeviceGuard g(b->device);
    for (auto& e : b->events) {
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    b->events.clear();
    b->stage_ms.assign(b->plan->stages.size(), 0.0);
    b->stage_launches.assign(b->plan->stages.size(), 0);
    b->timing = enable != 0;
    return 0;
}

// Synchronises the stream, folds the pending event pairs into per-stage totals and returns the
// accumulated device time (ms) of `stage` since timing was enabled; *launches = kernel launches.
double r8bgpu_batch_stage_time_ms(r8bgpu_batch* b, int stage, unsigned long long* launches)
{
    if (b->front) { // the shards run side by side: report the slowest one
        double worst = 0.0;
        for (r8bgpu_batch* sb : b->front->shards) {
            const double t = r8bgpu_batch_stage_time_ms(sb, stage, launches);
            if (t < 0.0) return t;
            if (t > worst) worst = t;
        }
        return worst;
    }
    if (stage < 0 || stage >= (int) b->stage_ms.size()) {
        set_err("stage_time_ms: timing not enabled or bad stage");
        return -1.0;
    }
    DeviceGuard g(b->device);
    if (!cuda_ok(cudaStreamSynchronize(b->stream), "stage_time_ms: sync")) return -1.0;
    for (auto& e : b->events) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) {
            b->stage_ms[(size_t) e.stage] += ms;
            b->stage_launches[(size_t) e.stage]++;
        }
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    b->events.clear();
    if (launches) *launches = b->stage_launches[(size_t) stage];
    return b->stage_ms[(size_t) stage];
}

// Which kernel executes plan stage `stage`, and how many consecutive plan stages it covers
// (0 = this stage is folded into the kernel of an earlier stage).
int r8bgpu_batch_stage_kernel(const r8bgpu_batch* b, int stage, char* name, int cap)
{
    if (b->front) return r8bgpu_batch_stage_kernel(b->front->shards[0], stage, name, cap);
    if (stage < 0 || stage >= (int) b->plan->stages.size()) {
        set_err("stage_kernel: bad stage index");
        return -1;
    }
    const StageDev& d = b->dev[(size_t) stage];
    const StageDesc& s = b->plan->stages[(size_t) stage];
    const char* nm = "";
    int span = 1;
    if (d.fused_into_prev) {
        nm = "(fused)";
        span = 0;
    } else if (d.fused_with_next) {
        nm = d.f2_ok ? "k_up2_frac2" : d.f2_poly ? "k_up2_frac2<poly>|k_up2_frac" : "k_up2_frac";
        span = 2;
    } else if (d.down_casc_len >= 2) {
        nm = "k_hbdown_cascade";
        span = d.down_casc_len;
    } else if (d.casc_len >= 2) {
        nm = "k_hbup_cascade";
        span = d.casc_len;
    } else {
        switch (s.kind) {
        case ST_BLOCKCONV: nm = d.f2_copy ? "k_up2_frac2<copy>" : "k_blockconv"; break;
        case ST_FRAC_WHOLE: nm = "k_frac<false>"; break;
        case ST_FRAC_POLY: nm = "k_frac<true>"; break;
        case ST_HBUP: nm = "k_hbup"; break;
        default: nm = "k_hbdown"; break;
        }
    }
    if (name != nullptr && cap > 0) {
        strncpy(name, nm, (size_t) cap - 1);
        name[cap - 1] = 0;
    }
    return span;
}

int r8bgpu_batch_set_stream(r8bgpu_batch* b, void* stream)
{
    if (b->front) {
        if (stream == nullptr) return 0;
        set_err("batch_set_stream: a multi-device batch has one stream per shard (use r8bgpu_batch_shard())");
        return -1;
    }
    if ((cudaStream_t) stream != b->stream) {
        // work queued on the old stream (previous calls share the rings with the next ones) must be ordered before anything
        // the new stream runs: an event recorded there, waited for here
        DeviceGuard g(b->device);
        cudaEvent_t ev;
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess) {
            cudaEventRecord(ev, b->stream);
            cudaStreamWaitEvent((cudaStream_t) stream, ev, 0);
            cudaEventDestroy(ev);
        }
    }
    b->stream = (cudaStream_t) stream;
    return 0;
}

int r8bgpu_batch_clear(r8bgpu_batch* b)
{
    if (b->front) {
        int rc = 0;
        for (r8bgpu_batch* sb : b->front->shards) rc |= r8bgpu_batch_clear(sb);
        return rc == 0 ? 0 : -1;
    }
    DeviceGuard g(b->device);
    b->sched.clear();
    for (auto& d : b->dev) {
        if (d.ring == nullptr) continue;
        if (!cuda_ok(cudaMemsetAsync(d.ring, 0, (size_t) d.ring_cap * (size_t) b->n_ch * sizeof(double), b->stream),
                     "batch_clear: cudaMemsetAsync"))
            return -1;
    }
    // clear() is rare; finishing it here keeps the device path (batch stream) and the host path
    // (internal pipeline streams) ordered without cross-stream events
    if (!cuda_ok(cudaStreamSynchronize(b->stream), "batch_clear: sync")) return -1;
    return 0;
}

int r8bgpu_batch_sync(r8bgpu_batch* b)
{
    if (b->front) {
        int rc = 0;
        for (r8bgpu_batch* sb : b->front->shards) rc |= r8bgpu_batch_sync(sb);
        return rc == 0 ? 0 : -1;
    }
    DeviceGuard g(b->device);
    return cuda_ok(cudaStreamSynchronize(b->stream), "batch_sync") ? 0 : -1;
}

// R8B_FASTTIMING: ship this call's host-walked (position, fraction) sequence to the device, in stream order
// ahead of the kernels that read it.  Two pinned staging buffers alternate; an event guards their reuse.
static bool upload_fasttiming(r8bgpu_batch* b, cudaStream_t st)
{
    const Plan& P = *b->plan;
    for (size_t i = 0; i < P.stages.size(); i++) {
        const StageDesc& s = P.stages[i];
        if (s.kind != ST_FRAC_POLY || !s.fasttiming) continue;
        StageDev& d = b->dev[i];
        const StageCall& c = b->calls[i];
        const size_t n = c.ft_dp.size();
        if (n == 0) continue;
        const int k = (d.ft_cur ^= 1);
        cudaEventSynchronize(d.ft_ev[k]);
        memcpy(d.h_dp[k], c.ft_dp.data(), n * sizeof(int));
        memcpy(d.h_fpos[k], c.ft_fpos.data(), n * sizeof(double));
        if (!cuda_ok(cudaMemcpyAsync(d.ft_dp, d.h_dp[k], n * sizeof(int), cudaMemcpyHostToDevice, st), "fasttiming upload")) return false;
        if (!cuda_ok(cudaMemcpyAsync(d.ft_fpos, d.h_fpos[k], n * sizeof(double), cudaMemcpyHostToDevice, st), "fasttiming upload")) return false;
        cudaEventRecord(d.ft_ev[k], st);
    }
    return true;
}

// Launches every kernel of one process() call (already scheduled in b->calls) for the channel range
// [ch0, ch0+nch) on stream st.  d_in / d_out point at the FIRST channel of that range.
// TypedIO: when the first / last kernel of the chain converts caller-side sample formats itself (fuses_input_format(),
// fuses_output_format()), d_in / d_out address planar samples of that format and the strides count samples.
struct TypedIO {
    int in_fmt = FMT_F64, out_fmt = FMT_F64;
    double in_scale = 1.0, out_scale = 1.0;
};

// The v2 fused kernel widens a planar typed block while it gathers its tiles / narrows in its tensor-path stores.
static bool fuses_input_format(const r8bgpu_batch* b)
{
    return !b->plan->passthrough && !b->dev.empty() && ((b->dev[0].fused_with_next && b->dev[0].f2_ok &&
           b->plan->stages[1].kind == ST_FRAC_WHOLE) || b->dev[0].f2_copy) && !getenv("R8BGPU_NO_FORMAT_FUSION");
}
static bool fuses_output_format(const r8bgpu_batch* b)
{
    const size_t ns = b->dev.size();
    return !b->plan->passthrough && ns >= 2 && b->plan->stages[ns - 1].kind == ST_FRAC_WHOLE && b->dev[ns - 1].fused_into_prev &&
           b->dev[ns - 2].f2_ok && b->dev[ns - 1].bank_frag_order && !getenv("R8BGPU_NO_FORMAT_FUSION");
}

static void launch_call(r8bgpu_batch* b, const double* d_in, size_t in_stride, int l, double* d_out,
                        size_t out_stride, int ch0, int nch, cudaStream_t st, const TypedIO& tio = TypedIO())
{
    const Plan& P = *b->plan;
    const size_t ns = P.stages.size();
    for (size_t i = 0; i < ns; i++) {
        const StageDesc& s = P.stages[i];
        const StageCall& c = b->calls[i];
        const StageDev& d = b->dev[i];
        if (d.fused_into_prev) continue; // handled together with the previous stage
        const bool fused = d.fused_with_next;
        const size_t last = fused ? i + 1 : (d.casc_len >= 2 ? i + (size_t) d.casc_len - 1 : d.down_casc_len >= 2 ? i + (size_t) d.down_casc_len - 1 : i); // stage whose output this launch produces
        if (b->calls[last].e1 <= b->calls[last].e0) continue;
        SrcView src;
        src.ring = d.ring + (long long) ch0 * d.ring_cap;
        src.ring_stride = d.ring_cap;
        src.ring_mask = d.ring_cap - 1;
        if (i == 0) {
            src.cur = d_in;
            src.cur_stride = (long long) in_stride;
            src.cur_base = c.n0;
            src.cur_fmt = tio.in_fmt;
            src.cur_scale = tio.in_scale;
        } else {
            src.cur = nullptr;
            src.cur_stride = 0;
            src.cur_base = LLONG_MAX;
        }
        src.avail = c.n1;
        DstView dst;
        if (last + 1 == ns) {
            dst.ptr = d_out;
            dst.stride = (long long) out_stride;
            dst.mask = -1;
            dst.base = b->calls[last].e0;
            dst.fmt = tio.out_fmt;
            dst.scale = tio.out_scale;
        } else {
            dst.ptr = b->dev[last + 1].ring + (long long) ch0 * b->dev[last + 1].ring_cap;
            dst.stride = b->dev[last + 1].ring_cap;
            dst.mask = b->dev[last + 1].ring_cap - 1;
            dst.base = 0;
        }
        r8bgpu_batch::EvPair ev{(int) i, nullptr, nullptr};
        if (b->timing) {
            cudaEventCreate(&ev.a);
            cudaEventCreate(&ev.b);
            cudaEventRecord(ev.a, st);
        }
        if (d.down_casc_len >= 2) {
            HbDownCascParams p = d.down_casc;
            p.e0 = b->calls[last].e0;
            p.e1 = b->calls[last].e1;
            p.n_tiles = (int) ((p.e1 - p.e0 + p.w - 1) / p.w);
            launch_hbdown_cascade(p, d.down_casc_smem, src, dst, nch, st);
            b->launches++;
        } else if (d.casc_len >= 2) {
            const int cl = d.casc_len;
            HbCascadeParams p;
            memset(&p, 0, sizeof p);
            p.n_stages = cl;
            for (int k = 0; k < cl; k++) {
                const StageDesc& h = P.stages[i + (size_t) k];
                p.ntaps[k] = h.hb_taps;
                for (int j = 0; j < h.hb_taps; j++) p.taps[k][j] = h.hb[(size_t) j];
            }
            p.e0 = b->calls[last].e0;
            p.e1 = b->calls[last].e1;
            // halos, from the last stage backwards (see k_hbup_cascade)
            p.lo_off[cl] = 0;
            p.hi_off[cl] = 0;
            for (int k = cl - 1; k >= 0; k--) {
                const int T = p.ntaps[k];
                p.lo_off[k] = (p.lo_off[k + 1] + 1) / 2 + T - 1;
                p.hi_off[k] = (p.hi_off[k + 1] >= 1 ? (p.hi_off[k + 1] - 1) / 2 : -1) + T + 1;
            }
            p.fuse_last2 = (cl >= 2 && hb_last2_supported(p.ntaps[cl - 2], p.ntaps[cl - 1]) && !getenv("R8BGPU_HB_NO_LAST2")) ? 1 : 0;
            const int nbuf = p.fuse_last2 ? cl - 1 : cl; // streams 0 .. nbuf-1 live in shared memory
            int halo = 0;
            for (int k = 0; k < nbuf; k++) halo += p.lo_off[k] + p.hi_off[k] + 8;
            int budget = 7000; // doubles of shared memory per CTA (4 CTAs of 128 threads per SM; measured best); buffers carry a 5/4 skew
            if (const char* e = getenv("R8BGPU_HB_SMEM_DOUBLES")) budget = atoi(e);
            int w = (((budget * 4) / 5 - halo) / ((1 << nbuf) - 1)) & ~31;
            if (w > 1024) w = 1024;
            if (w < 32) w = 32;
            p.w = w;
            int off = 0;
            for (int k = 0; k < nbuf; k++) {
                p.boff[k] = off; // buffer k starts at its own lo bound
                off += (((w << k) + p.lo_off[k] + p.hi_off[k] + 8) * 5 + 3) / 4 + 2; // + slack: threads work in quads; 5/4 skew
                off = (off + 1) & ~1;
            }
            const int smem_bytes = off * (int) sizeof(double);
            p.a0 = p.e0 >> cl;
            const long long span = p.e1 - (p.a0 << cl);
            p.n_tiles = (int) ((span + ((long long) w << cl) - 1) / ((long long) w << cl));
            launch_hbup_cascade(p, smem_bytes, src, dst, nch, st);
            b->launches++;
        } else if (fused) {
            const StageDesc& f = P.stages[i + 1];
            const StageCall& fc = b->calls[i + 1];
            const StageDev& fd = b->dev[i + 1];
            FusedParams p;
            memset(&p, 0, sizeof p);
            if (f.kind == ST_FRAC_WHOLE) {
                fused_whole_fields(p, f, fc.e0, fc.e1);
            } else {
                p.mode = 1;
                p.flen = f.bank.filter_len;
                p.fll = p.flen / 2 - 1;
                p.e0 = fc.e0;
                p.e1 = fc.e1;
                p.p_lo = fc.p0 & ~1LL; // even (positions are >= 0)
                p.p_hi = fc.p_last + 1;
                p.in_step = f.in_step;
                p.out_step = f.out_step;
            }
            // order-2 bank on the v2 kernel: ratios within 1e-3 of an integer 1..3 (windows of consecutive outputs N apart)
            bool v2_poly = false;
            if (p.mode == 1 && d.f2_poly) {
                const double ratio = f.src_rate / f.dst_rate;
                const long long nn = llround(ratio);
                v2_poly = nn >= 1 && nn <= 3 && fabs(ratio - (double) nn) < 1e-3 * (double) nn;
                if (v2_poly) p.poly_n = (int) nn;
            }
            const bool v2 = (d.f2_ok && p.mode == 0) || v2_poly;
            if (v2) {
                fused2_tiles(p, d.fgeom, i == 0 ? (int) (c.n0 & 1) : -1);
            } else {
                const long long range = p.p_hi - p.p_lo;
                long long nt = (range + d.span_max - 1) / d.span_max;
                if (nt > 1 && (nt & 1)) nt++;
                p.n_tiles = (int) nt;
                p.span = (int) (((range + nt - 1) / nt + 1) & ~1LL);
            }
            p.yl = d.yl;
            p.lg = d.lg;
            p.ysh = d.ysh;
            p.spec = d.spec;
            p.tw = d.tw;
            p.bank = fd.bank;
            p.bank_len = (int) f.bank.table.size();
            p.bank_in_smem = fd.bank_in_smem;
            p.gbank = fd.gbank;
            p.gbank_len = fd.gbank_len;
            p.smaxp = fd.smaxp;
            p.goff = fd.goff;
            p.ir = fd.ir;
            p.gbank_smem_len = fd.gbank_smem_len;
            {
                // store staging area behind the bank, if shared memory allows (whole stepping, 8-phase groups)
                const int used = fused_fixed_doubles() + (p.bank_in_smem ? ((p.gbank_smem_len + 1) & ~1) : 0);
                p.stage_off = (p.mode == 0 && p.ir == 8 && !getenv("R8BGPU_NO_STAGE") &&
                               (used + fused_stage_doubles()) * 8 <= 224 * 1024) ? used : 0;
            }
            p.phase_off = fd.phase_off;
            p.phase_row = fd.phase_row;
            p.fracs = f.bank.fracs;
            p.ssr = f.src_rate;
            p.dsr = f.dst_rate;
            p.in_counter0 = fc.in_counter0;
            p.in_pos_int0 = fc.in_pos_int0;
            p.in_pos_shift = fc.in_pos_shift;
            p.fpos0 = fc.fpos0;
            p.p0 = fc.p0;
            p.pos_dp = fd.ft_dp;
            p.pos_fpos = fd.ft_fpos;
            p.bank = fd.bank;
            if (p.mode == 1 && !v2_poly && (p.flen & 1) == 0 && !getenv("R8BGPU_BANK_GLOBAL")) {
                // bank-row drift per output, in rows: frac(ssr/dsr) * fracs upward, or (1 - frac) * fracs downward
                const double ratio = p.ssr / p.dsr, fr = ratio - floor(ratio);
                const double outs = 2.0 * p.span / ratio + 4.0; // outputs one tile pair can own
                const int row_words = 6 * p.flen; // 32-bit words per bank row
                p.poly_row_stride = 3 * p.flen + ((row_words % 8) == 4 ? 0 : 2);
                const int cap = (224 * 1024 - fused_smem_bytes(0) - fused_poly_queue_bytes()) /
                                (p.poly_row_stride * (int) sizeof(double));
                const double up = fr * p.fracs * outs + 4.0, dn = (1.0 - fr) * p.fracs * outs + 4.0;
                const double need = up < dn ? up : dn;
                const int chunks = (int) ceil(need / cap);
                if (cap >= 8 && chunks <= 4) { // more pieces than that: the rows are not a short run, read them from L2
                    p.poly_dir = up < dn ? 1 : -1;
                    p.poly_rows_cap = cap;
                    p.poly_chunks = chunks < 1 ? 1 : chunks;
                    const long long nn = llround(ratio);
                    if (nn >= 1 && nn <= 3 && !getenv("R8BGPU_POLY_SINGLE")) {
                        // four consecutive outputs per thread: lanes step by 4*nn samples through the tile -> padded
                        // y layout (i + (i >> 4), the only padding the tile buffers have room for)
                        p.poly_n = (int) nn;
                        p.ysh = 4;
                    }
                }
            }
            if (p.poly_chunks < 1) p.poly_chunks = 1;
            if (b->prof == nullptr && getenv("R8BGPU_PROFILE")) {
                cudaMalloc(&b->prof, 10 * sizeof(unsigned long long));
                cudaMemset(b->prof, 0, 10 * sizeof(unsigned long long));
            }
            p.prof = b->prof;
#ifdef R8BGPU_EXPERIMENTS
            if (const char* e = getenv("R8BGPU_DEBUG")) p.debug = atoi(e);
#endif
            if (b->prof) b->prof_ctas += (unsigned long long) ((p.n_tiles + 1) / 2) * nch;
            if (v2) {
                p.n_ch = nch;
                p.flags = b->f2_flags;
                p.tw_tab = d.tw_tab;
                p.c_tab = d.c_tab;
                p.cd_tab = d.cd_tab;
                p.up = d.fgeom.up;
                p.ylen = d.fgeom.up * 4096;
                if (!fd.bank_frag_order) p.flags &= ~4; // (the bank layout decides: see batch_create)
                p.stage_off = (p.ir == 8 && !(p.flags & 4) && fused2_smem_bytes(p.gbank_smem_len, true) <= kFused2SmemMax &&
                               !getenv("R8BGPU_NO_STAGE")) ? fused2_stage_off(p.gbank_smem_len) : 0;
                if (v2_poly) { // plain y layout; no grouped bank, no staging area in shared memory
                    p.ysh = 31;
                    p.gbank_smem_len = 0;
                    p.stage_off = 0;
                }
                p.glog = v2_poly ? 0 : fused2_choose_glog(p.span, f.in_step, f.out_step, p.ir);
                p.mbu = v2_poly ? 3 : fused2_choose_mbu(p.span, f.in_step, f.out_step);
                launch_up2_frac2(p, src, dst, b->n_sm, st);
            } else {
                p.c_tab = d.c_tab_v1;
                launch_up2_frac(p, src, dst, nch, st);
            }
            b->launches++;
        } else
        switch (s.kind) {
        case ST_BLOCKCONV: {
            if (d.f2_copy) {
                FusedParams fp;
                memset(&fp, 0, sizeof fp);
                fp.mode = 2;
                fp.e0 = c.e0;
                fp.e1 = c.e1;
                fp.p_lo = c.e0 & ~1LL;
                fp.p_hi = c.e1;
                fused2_tiles(fp, d.fgeom, i == 0 ? (int) (c.n0 & 1) : -1);
                fp.lg = d.lg;
                fp.ysh = 31;
                fp.spec = d.spec;
                fp.tw = d.tw;
                fp.tw_tab = d.tw_tab;
                fp.c_tab = d.c_tab;
                fp.cd_tab = d.cd_tab;
                fp.up = 2;
                fp.ylen = 8192;
                fp.ir = 8;
                fp.out_step = 8; // one (unused) phase group
                fp.smaxp = 4;
                fp.n_ch = nch;
                fp.flags = b->f2_flags & 2;
                launch_up2_frac2(fp, src, dst, b->n_sm, st);
                b->launches++;
                break;
            }
            BlockConvParams p;
            const int up_eff = d.virt_up > 1 ? 1 : s.up;
            p.up = up_eff;
            p.src_up = d.virt_up;
            p.down = s.down;
            p.lg = d.lg;
            p.fft_log2 = d.fft_log2;
            p.e0 = c.e0;
            p.e1 = c.e1;
            p.m0 = (c.e0 * s.down) / up_eff;             // floor; indices are >= 0
            p.m1 = ((c.e1 - 1) * s.down) / up_eff + 1;
            if (s.block_exact) {
                // tile b = reference block b: owns positions [b*InputLen - L, (b+1)*InputLen - L)
                const long long il = s.ref_input_len, L = s.lp.half_len;
                const long long b0 = (p.m0 + L) / il, b1 = (p.m1 - 1 + L) / il;
                p.m0 = b0 * il - L;
                p.adv = (int) il;
                p.n_tiles = (int) (b1 - b0 + 1);
            } else {
                const int adv_max = (1 << d.fft_log2) - 2 * d.lg;
                const long long span = p.m1 - p.m0;
                long long nt = (span + adv_max - 1) / adv_max;
                if (nt > 1 && (nt & 1)) nt++; // tiles are transformed in pairs
                p.n_tiles = (int) nt;
                p.adv = (int) ((span + nt - 1) / nt);
            }
            p.trunc = s.block_exact ? s.down : 0;
            p.nyq_gain = d.nyq_gain;
            p.spec = d.spec;
            p.tw = d.tw;
            launch_blockconv(p, src, dst, nch, st);
            b->launches++;
            break;
        }
        case ST_FRAC_WHOLE:
        case ST_FRAC_POLY: {
            FracParams p;
            memset(&p, 0, sizeof p);
            p.flen = s.bank.filter_len;
            p.fll = s.bank.filter_len / 2 - 1;
            p.e0 = c.e0;
            p.e1 = c.e1;
            p.bank = d.bank;
            p.in_step = s.in_step;
            p.out_step = s.out_step;
            p.fracs = s.bank.fracs;
            p.ssr = s.src_rate;
            p.dsr = s.dst_rate;
            p.in_counter0 = c.in_counter0;
            p.in_pos_int0 = c.in_pos_int0;
            p.in_pos_shift = c.in_pos_shift;
            p.fpos0 = c.fpos0;
            p.p0 = c.p0;
            p.pos_dp = d.ft_dp;
            p.pos_fpos = d.ft_fpos;
            if (s.kind == ST_FRAC_WHOLE) launch_frac_whole(p, src, dst, nch, st);
            else launch_frac_poly(p, src, dst, nch, st);
            b->launches++;
            break;
        }
        case ST_HBUP:
        case ST_HBDOWN: {
            HbParams p;
            memset(&p, 0, sizeof p);
            p.ntaps = s.hb_taps;
            p.e0 = c.e0;
            p.e1 = c.e1;
            for (int k = 0; k < s.hb_taps; k++) p.taps[k] = s.hb[(size_t) k];
            if (s.kind == ST_HBUP) launch_hbup(p, src, dst, nch, st);
            else launch_hbdown(p, src, dst, nch, st);
            b->launches++;
            break;
        }
        }
        if (b->timing) {
            cudaEventRecord(ev.b, st);
            b->events.push_back(ev);
        }
    }
    // Keep the most recent input samples for the next calls.
    if (l > 0) {
        const StageCall& c0 = b->calls[0];
        const StageDev& d0 = b->dev[0];
        long long from = c0.n1 - d0.ring_cap;
        if (from < c0.n0) from = c0.n0;
        launch_save_tail(d_in, (long long) in_stride, c0.n0, from, c0.n1, d0.ring + (long long) ch0 * d0.ring_cap, d0.ring_cap,
                         d0.ring_cap - 1, nch, st, tio.in_fmt, tio.in_scale);
        b->launches++;
    }
}

int r8bgpu_batch_process(r8bgpu_batch* b, const double* d_in, size_t in_stride, int l, double* d_out,
                         size_t out_stride, int out_cap)
{
    if (b == nullptr || l < 0 || l > b->plan->max_in_len) {
        set_err("batch_process: l must be in [0, MaxInLen]");
        return -1;
    }
    if (l > 0 && d_in == nullptr) {
        set_err("batch_process: null input");
        return -1;
    }
    if (b->front) {
        set_err("batch_process: device buffers live on one GPU; call the shards of a multi-device batch (r8bgpu_batch_shard())");
        return -1;
    }
    DeviceGuard g(b->device);
    const Plan& P = *b->plan;
    const cudaStream_t st = b->stream;
    if (P.passthrough) { // SrcSampleRate == DstSampleRate: the reference hands the input back
        if (l > out_cap) {
            set_err("batch_process: output capacity too small");
            return -1;
        }
        if (l > 0 && !cuda_ok(cudaMemcpy2DAsync(d_out, out_stride * sizeof(double), d_in,
                                                in_stride * sizeof(double), (size_t) l * sizeof(double),
                                                (size_t) b->n_ch, cudaMemcpyDeviceToDevice, st),
                              "batch_process: passthrough copy"))
            return -1;
        return l;
    }

    Schedule saved = b->sched;
    const int n_out = b->sched.advance(l, b->calls);
    if (n_out > out_cap || (n_out > 0 && d_out == nullptr)) {
        b->sched = saved;
        set_err("batch_process: output capacity too small for this call");
        return -1;
    }

    if (!upload_fasttiming(b, st)) {
        b->sched = saved;
        return -1;
    }
    launch_call(b, d_in, in_stride, l, d_out, out_stride, 0, b->n_ch, st);
    if (!cuda_ok(cudaGetLastError(), "batch_process: kernel launch")) {
        b->sched = saved; // a refused launch did nothing: the call did not happen
        return -1;
    }
    return n_out;
}

// ---- staging shared by the host path and the sample-format paths ----------------------------
} // extern "C" (helpers below have C++ linkage)

static bool ensure_staging(r8bgpu_batch* b)
{
    if (b->st_in != nullptr) return true;
    const Plan& P = *b->plan;
    const size_t in_cap = (size_t) P.max_in_len;
    const size_t o_cap = ((size_t) P.max_out_len + 3) & ~(size_t) 3; // rows 32-byte aligned
    if (!cuda_ok(cudaMalloc(&b->st_in, in_cap * b->n_ch * sizeof(double)), "staging: cudaMalloc(in)")) return false;
    if (!cuda_ok(cudaMalloc(&b->st_out, o_cap * b->n_ch * sizeof(double)), "staging: cudaMalloc(out)")) return false;
    b->dev_bytes += (in_cap + o_cap) * b->n_ch * sizeof(double);
    if (!cuda_ok(cudaStreamCreateWithFlags(&b->s_h2d, cudaStreamNonBlocking), "staging: stream")) return false;
    if (!cuda_ok(cudaStreamCreateWithFlags(&b->s_d2h, cudaStreamNonBlocking), "staging: stream")) return false;
    if (!cuda_ok(cudaStreamCreateWithFlags(&b->s_comp, cudaStreamNonBlocking), "staging: stream")) return false;
    int groups = 8;
    if (const char* e = getenv("R8BGPU_HOST_GROUPS")) groups = atoi(e);
    if (groups < 1) groups = 1;
    while (groups > 1 && b->n_ch / groups < 32) groups /= 2; // keep every group a full-GPU launch
    b->host_groups = groups;
    b->ev_h2d.resize((size_t) groups);
    b->ev_k.resize((size_t) groups);
    for (int i = 0; i < groups; i++) {
        cudaEventCreateWithFlags(&b->ev_h2d[(size_t) i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&b->ev_k[(size_t) i], cudaEventDisableTiming);
    }
    return true;
}

static bool buffer_is_plain(const r8bgpu_buffer& d)
{
    return d.format == R8BGPU_F64 && !d.interleaved && d.scale == 1.0;
}

static bool check_buffer(const r8bgpu_batch* b, const r8bgpu_buffer* d, const char* what)
{
    if (d == nullptr || format_bytes(d->format) == 0) {
        set_err(std::string(what) + ": unknown sample format");
        return false;
    }
    if (d->interleaved && d->stride < (size_t) b->n_ch) {
        set_err(std::string(what) + ": interleaved stride smaller than the channel count");
        return false;
    }
    if (!(d->scale == d->scale) || d->scale == 0.0) {
        set_err(std::string(what) + ": scale must be a non-zero number");
        return false;
    }
    return true;
}

// Host-pointer path.  The batch is cut into channel groups that flow through a three-stage pipeline
//   copy stream A: H2D(group g+1)  |  compute stream: kernels(group g)  |  copy stream B: D2H(group g-1)
// so the two PCIe directions and the SMs work at the same time (channels are independent, so a group
// is a self-contained sub-batch).  Staging buffers are per channel, so groups never alias.
// Narrow / interleaved sample formats cross PCIe as they are and are widened (narrowed) on the
// device by r8b_format.cu, in the compute stage of the same pipeline.
static int process_host_impl(r8bgpu_batch* b, const r8bgpu_buffer& in, int l, const r8bgpu_buffer& out, int out_cap);

// every shard processes its own channel range of the caller's buffers on its own thread, stream set and PCIe link
static int process_host_front(r8bgpu_batch* b, const r8bgpu_buffer& in, int l, const r8bgpu_buffer& out, int out_cap)
{
    const ShardFront& F = *b->front;
    auto view = [](const r8bgpu_buffer& d, int c0) {
        r8bgpu_buffer v = d;
        if (d.data != nullptr) {
            const size_t e = (size_t) format_bytes(d.format);
            v.data = (unsigned char*) d.data + (d.interleaved ? (size_t) c0 * e : (size_t) c0 * d.stride * e);
        }
        return v;
    };
    return front_run(b, [&](r8bgpu_batch* sb, int s) {
        return process_host_impl(sb, view(in, F.ch0[(size_t) s]), l, view(out, F.ch0[(size_t) s]), out_cap);
    });
}

static int process_host_impl(r8bgpu_batch* b, const r8bgpu_buffer& in, int l, const r8bgpu_buffer& out, int out_cap)
{
    if (b->front) return process_host_front(b, in, l, out, out_cap);
    if (l < 0 || l > b->plan->max_in_len) {
        set_err("batch_process_host: l must be in [0, MaxInLen]");
        return -1;
    }
    if (l > 0 && in.data == nullptr) {
        set_err("batch_process_host: null input");
        return -1;
    }
    DeviceGuard g(b->device);
    const Plan& P = *b->plan;
    const size_t in_cap = (size_t) P.max_in_len;
    const size_t o_cap = ((size_t) P.max_out_len + 3) & ~(size_t) 3;
    if (!ensure_staging(b)) return -1;
    const bool in_plain = buffer_is_plain(in), out_plain = buffer_is_plain(out);
    const size_t ein = (size_t) format_bytes(in.format), eout = (size_t) format_bytes(out.format);
    if (!in_plain && b->raw_in == nullptr) {
        if (!cuda_ok(cudaMalloc(&b->raw_in, in_cap * b->n_ch * 8), "process_host: cudaMalloc(raw in)")) return -1;
        b->dev_bytes += in_cap * b->n_ch * 8;
    }
    if (!out_plain && b->raw_out == nullptr) {
        if (!cuda_ok(cudaMalloc(&b->raw_out, o_cap * b->n_ch * 8), "process_host: cudaMalloc(raw out)")) return -1;
        b->dev_bytes += o_cap * b->n_ch * 8;
    }
    int n = l;
    const Schedule saved = b->sched;
    // any failure after the schedule has advanced: put it back and drain the pipeline streams, so that the rings and the
    // schedule still agree on the next call (the failed call then simply did not happen)
    auto fail = [&]() {
        b->sched = saved;
        cudaStreamSynchronize(b->s_h2d);
        cudaStreamSynchronize(b->s_comp);
        cudaStreamSynchronize(b->s_d2h);
        return -1;
    };
    if (!P.passthrough) {
        n = b->sched.advance(l, b->calls);
        if (n > out_cap || (n > 0 && out.data == nullptr)) {
            b->sched = saved;
            set_err("process_host: output capacity too small for this call");
            return -1;
        }
    } else if (l > out_cap) {
        set_err("process_host: output capacity too small");
        return -1;
    }
    // order after any device-path work queued on the batch stream (the two paths share the rings)
    if (!cuda_ok(cudaStreamSynchronize(b->stream), "process_host: sync(batch stream)")) return fail();
    if (!P.passthrough && !upload_fasttiming(b, b->s_comp)) return fail();
    const int G = b->host_groups;
    const unsigned char* hin = (const unsigned char*) in.data;
    unsigned char* hout = (unsigned char*) out.data;
    for (int gi = 0; gi < G; gi++) {
        const int ch0 = (int) ((long long) b->n_ch * gi / G);
        const int ch1 = (int) ((long long) b->n_ch * (gi + 1) / G);
        const int nch = ch1 - ch0;
        if (nch <= 0) continue;
        double* din = b->st_in + (size_t) ch0 * in_cap;
        double* dout = b->st_out + (size_t) ch0 * o_cap;
        unsigned char* rin = in_plain ? nullptr : b->raw_in + (size_t) ch0 * in_cap * 8;
        unsigned char* rout = out_plain ? nullptr : b->raw_out + (size_t) ch0 * o_cap * 8;
        if (l > 0) {
            cudaError_t e;
            if (in_plain)
                e = cudaMemcpy2DAsync(din, in_cap * 8, hin + (size_t) ch0 * in.stride * 8, in.stride * 8,
                                      (size_t) l * 8, (size_t) nch, cudaMemcpyHostToDevice, b->s_h2d);
            else if (in.interleaved) // device copy: compact [l][nch]
                e = cudaMemcpy2DAsync(rin, (size_t) nch * ein, hin + (size_t) ch0 * ein, in.stride * ein,
                                      (size_t) nch * ein, (size_t) l, cudaMemcpyHostToDevice, b->s_h2d);
            else // device copy: [nch][in_cap]
                e = cudaMemcpy2DAsync(rin, in_cap * ein, hin + (size_t) ch0 * in.stride * ein, in.stride * ein,
                                      (size_t) l * ein, (size_t) nch, cudaMemcpyHostToDevice, b->s_h2d);
            if (!cuda_ok(e, "process_host: H2D")) return fail();
        }
        cudaEventRecord(b->ev_h2d[(size_t) gi], b->s_h2d);
        cudaStreamWaitEvent(b->s_comp, b->ev_h2d[(size_t) gi], 0);
        TypedIO tio;
        const bool in_fused = !in_plain && !in.interleaved && fuses_input_format(b);
        const bool out_fused = !out_plain && !out.interleaved && fuses_output_format(b);
        if (in_fused) {
            tio.in_fmt = in.format;
            tio.in_scale = in.scale;
        } else if (!in_plain) {
            launch_to_f64(in.format, rin, in.interleaved != 0, in.interleaved ? (size_t) nch : in_cap, din, in_cap, l,
                          nch, in.scale, b->s_comp);
            if (l > 0) b->launches++;
        }
        if (out_fused) {
            tio.out_fmt = out.format;
            tio.out_scale = out.scale;
        }
        if (P.passthrough) {
            if (l > 0) cudaMemcpy2DAsync(dout, o_cap * sizeof(double), din, in_cap * sizeof(double),
                                         (size_t) l * sizeof(double), (size_t) nch, cudaMemcpyDeviceToDevice, b->s_comp);
        } else {
            launch_call(b, in_fused ? (const double*) rin : din, in_cap, l, out_fused ? (double*) rout : dout, o_cap, ch0, nch,
                        b->s_comp, tio);
        }
        if (!out_plain && !out_fused) {
            launch_from_f64(out.format, rout, out.interleaved != 0, out.interleaved ? (size_t) nch : o_cap, dout, o_cap,
                            n, nch, out.scale, b->s_comp);
            if (n > 0) b->launches++;
        }
        cudaEventRecord(b->ev_k[(size_t) gi], b->s_comp);
        cudaStreamWaitEvent(b->s_d2h, b->ev_k[(size_t) gi], 0);
        if (n > 0) {
            cudaError_t e;
            if (out_plain)
                e = cudaMemcpy2DAsync(hout + (size_t) ch0 * out.stride * 8, out.stride * 8, dout, o_cap * 8,
                                      (size_t) n * 8, (size_t) nch, cudaMemcpyDeviceToHost, b->s_d2h);
            else if (out.interleaved)
                e = cudaMemcpy2DAsync(hout + (size_t) ch0 * eout, out.stride * eout, rout, (size_t) nch * eout,
                                      (size_t) nch * eout, (size_t) n, cudaMemcpyDeviceToHost, b->s_d2h);
            else
                e = cudaMemcpy2DAsync(hout + (size_t) ch0 * out.stride * eout, out.stride * eout, rout, o_cap * eout,
                                      (size_t) n * eout, (size_t) nch, cudaMemcpyDeviceToHost, b->s_d2h);
            if (!cuda_ok(e, "process_host: D2H")) return fail();
        }
    }
    if (!cuda_ok(cudaStreamSynchronize(b->s_d2h), "process_host: sync")) return fail();
    if (!cuda_ok(cudaStreamSynchronize(b->s_comp), "process_host: sync")) return fail();
    if (!cuda_ok(cudaGetLastError(), "process_host: kernel launch")) return fail();
    return n;
}

extern "C" {

int r8bgpu_batch_process_host(r8bgpu_batch* b, const double* h_in, size_t in_stride, int l, double* h_out,
                              size_t out_stride, int out_cap)
{
    if (b == nullptr) {
        set_err("batch_process_host: null batch");
        return -1;
    }
    const r8bgpu_buffer in = { const_cast<double*>(h_in), R8BGPU_F64, 0, in_stride, 1.0 };
    const r8bgpu_buffer out = { h_out, R8BGPU_F64, 0, out_stride, 1.0 };
    return process_host_impl(b, in, l, out, out_cap);
}

int r8bgpu_batch_process_host_fmt(r8bgpu_batch* b, const r8bgpu_buffer* h_in, int l, const r8bgpu_buffer* h_out,
                                  int out_cap)
{
    if (b == nullptr) {
        set_err("batch_process_host_fmt: null batch");
        return -1;
    }
    if (!check_buffer(b, h_in, "batch_process_host_fmt(in)") || !check_buffer(b, h_out, "batch_process_host_fmt(out)"))
        return -1;
    return process_host_impl(b, *h_in, l, *h_out, out_cap);
}

// Device buffers in any format: widen into the staging block, run, narrow into the caller's buffer --
// all on the batch stream, asynchronous like r8bgpu_batch_process().
int r8bgpu_batch_process_fmt(r8bgpu_batch* b, const r8bgpu_buffer* d_in, int l, const r8bgpu_buffer* d_out,
                             int out_cap)
{
    if (b == nullptr) {
        set_err("batch_process_fmt: null batch");
        return -1;
    }
    if (b->front) {
        set_err("batch_process_fmt: device buffers live on one GPU; call the shards of a multi-device batch (r8bgpu_batch_shard())");
        return -1;
    }
    if (!check_buffer(b, d_in, "batch_process_fmt(in)") || !check_buffer(b, d_out, "batch_process_fmt(out)")) return -1;
    const bool in_plain = buffer_is_plain(*d_in), out_plain = buffer_is_plain(*d_out);
    if (in_plain && out_plain)
        return r8bgpu_batch_process(b, (const double*) d_in->data, d_in->stride, l, (double*) d_out->data,
                                    d_out->stride, out_cap);
    if (l < 0 || l > b->plan->max_in_len) {
        set_err("batch_process_fmt: l must be in [0, MaxInLen]");
        return -1;
    }
    if (l > 0 && d_in->data == nullptr) {
        set_err("batch_process_fmt: null input");
        return -1;
    }
    DeviceGuard g(b->device);
    const Plan& P = *b->plan;
    const size_t in_cap = (size_t) P.max_in_len;
    const size_t o_cap = ((size_t) P.max_out_len + 3) & ~(size_t) 3;
    if (!ensure_staging(b)) return -1;
    const cudaStream_t st = b->stream;
    int n = l;
    if (!P.passthrough) {
        Schedule saved = b->sched;
        n = b->sched.advance(l, b->calls);
        if (n > out_cap || (n > 0 && d_out->data == nullptr)) {
            b->sched = saved;
            set_err("batch_process_fmt: output capacity too small for this call");
            return -1;
        }
        if (!upload_fasttiming(b, st)) return -1;
    } else if (l > out_cap) {
        set_err("batch_process_fmt: output capacity too small");
        return -1;
    }
    const double* src = (const double*) d_in->data;
    size_t src_stride = d_in->stride;
    TypedIO tio;
    const bool in_fused = !in_plain && !d_in->interleaved && fuses_input_format(b);
    const bool out_fused = !out_plain && !d_out->interleaved && fuses_output_format(b);
    if (in_fused) { // the first kernel reads the caller's samples as they are
        tio.in_fmt = d_in->format;
        tio.in_scale = d_in->scale;
    } else if (!in_plain) {
        launch_to_f64(d_in->format, d_in->data, d_in->interleaved != 0, d_in->stride, b->st_in, in_cap, l, b->n_ch,
                      d_in->scale, st);
        if (l > 0) b->launches++;
        src = b->st_in;
        src_stride = in_cap;
    }
    double* dst = (out_plain || out_fused) ? (double*) d_out->data : b->st_out;
    const size_t dst_stride = (out_plain || out_fused) ? d_out->stride : o_cap;
    if (out_fused) { // the last kernel narrows in its stores
        tio.out_fmt = d_out->format;
        tio.out_scale = d_out->scale;
    }
    if (P.passthrough) {
        if (l > 0) cudaMemcpy2DAsync(dst, dst_stride * sizeof(double), src, src_stride * sizeof(double),
                                     (size_t) l * sizeof(double), (size_t) b->n_ch, cudaMemcpyDeviceToDevice, st);
    } else {
        launch_call(b, src, src_stride, l, dst, dst_stride, 0, b->n_ch, st, tio);
    }
    if (!out_plain && !out_fused) {
        launch_from_f64(d_out->format, d_out->data, d_out->interleaved != 0, d_out->stride, b->st_out, o_cap, n, b->n_ch,
                        d_out->scale, st);
        if (n > 0) b->launches++;
    }
    if (!cuda_ok(cudaGetLastError(), "batch_process_fmt: kernel launch")) return -1;
    return n;
}

double r8bgpu_measure_fp64_tflops(int device)
{
    int dev = device;
    if (dev < 0 && !cuda_ok(cudaGetDevice(&dev), "measure_fp64")) return -1.0;
    DeviceGuard guard(dev);
    if (!guard.ok) {
        set_err("measure_fp64: cannot select device");
        return -1.0;
    }
    const double tf = measure_dfma_tflops();
    if (tf < 0.0) cuda_ok(cudaGetLastError(), "measure_fp64");
    return tf;
}

void* r8bgpu_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (!cuda_ok(cudaMallocHost(&p, bytes), "host_alloc")) return nullptr;
    return p;
}

void r8bgpu_host_free(void* p)
{
    if (p == nullptr) return;
    if (!numa_host_free(p)) cudaFreeHost(p);
}

} // extern "C"
