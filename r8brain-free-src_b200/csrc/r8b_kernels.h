// r8b_kernels.h -- launch interface between the host engine (r8b_engine.cu) and the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>

#include <atomic>

namespace r8bgpu {

// Opt a kernel into more than 48 KB of dynamic shared memory, once per (kernel, device): thread-safe, and
// devices beyond the bitmap simply set the attribute on every launch.
template <auto Kernel>
inline void ensure_dyn_smem(int bytes)
{
    static std::atomic<unsigned long long> done[4];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 256 && ((done[dev >> 6].load(std::memory_order_acquire) >> (dev & 63)) & 1ull)) return;
    // the opt-in limit (227 KB per CTA on sm_100) covers static + dynamic shared memory together
    cudaFuncAttributes fa{};
    if (cudaFuncGetAttributes(&fa, Kernel) == cudaSuccess && bytes > 227 * 1024 - (int) fa.sharedSizeBytes)
        bytes = 227 * 1024 - (int) fa.sharedSizeBytes;
    if (cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) {
        cudaGetLastError();
        return; // not marked done: the launch reports the error
    }
    if (dev >= 0 && dev < 256) done[dev >> 6].fetch_or(1ull << (dev & 63), std::memory_order_release);
}

// caller-side sample formats (values of r8bgpu_sample_format, include/r8bgpu.h)
enum { FMT_F64 = 0, FMT_F32 = 1, FMT_S16 = 2, FMT_S24 = 3, FMT_S32 = 4 };

// A per-channel sample stream addressed by ABSOLUTE sample index n (n = 0 is the first sample
// after clear()).  Samples with n >= cur_base are read from the caller's block of this
// process() call; older samples come from a power-of-two ring that holds the recent past.
// Indices < 0 land in ring slots that are still zero (rings are cleared at clear()).
struct SrcView {
    const double* ring;    // [n_ch][ring_stride]
    long long ring_stride;
    long long ring_mask;   // capacity-1
    const double* cur;     // [n_ch][cur_stride] or nullptr
    long long cur_stride;
    long long cur_base;    // absolute index of cur[0]; LLONG_MAX when there is no cur block
    long long avail;       // samples with n >= avail do not exist yet (read as 0)
    // Caller-side sample format of the cur block (FMT_*; FMT_F64 = plain doubles).  Only the v2 fused kernel and the
    // history copy read typed blocks: cur then points at planar samples of that format, cur_stride counts samples, and
    // a value is (double) sample * cur_scale -- the conversion of CDSPResampler::oneshot<Tin,Tout>()
    // (CDSPResampler.h:592-651) done in the gather instead of in a kernel of its own.
    int cur_fmt = 0;
    double cur_scale = 1.0;
};

// Destination stream: either a ring (mask = capacity-1, base = 0) or a linear block whose
// element 0 is absolute index `base` (mask = -1).
struct DstView {
    double* ptr;           // [n_ch][stride]
    long long stride;
    long long mask;
    long long base;
    // Sample format of a LINEAR destination (the v2 fused kernel's tensor-path stores only): ptr then addresses planar
    // samples of that format, stride counts samples, and a stored value is (T) (y * scale).
    int fmt = 0;
    double scale = 1.0;
};

struct BlockConvParams {
    int up, down;          // up is 1 or 2 here; other up-factors run as up = 1 on a zero-stuffed view
    int src_up;            // > 1: tile positions index the zero-stuffed stream x[t/src_up] (t % src_up == 0)
    int lg;                // half support of the polyphase filters, in input samples
    int fft_log2;          // log2(M)
    int adv;               // valid input-rate positions per tile (<= M - 2*lg)
    long long m0, m1;      // input-rate positions [m0,m1) whose outputs may be needed
    long long e0, e1;      // output indices to write
    int n_tiles;
    int trunc;             // 0, or D for reference-exact power-of-two decimation (see k_blockconv)
    double nyq_gain;       // scaled filter response at bin M/(2D) (trunc only)
    const double2* spec;   // filter spectrum in slot order, pre-scaled (device)
    const double2* tw;     // twiddles exp(-2*pi*i*k/M) (device)
};

struct FracParams {
    int flen, fll;
    long long e0, e1;
    const double* bank;    // device; [(fracs+1)][flen][order+1]
    // whole stepping
    int in_step, out_step;
    // polynomial (order 2)
    int fracs;
    double ssr, dsr;
    int in_counter0, in_pos_int0;
    double in_pos_shift, fpos0;
    long long p0;
    const int* pos_dp;       // R8B_FASTTIMING: per-output position - p0 and fraction (else nullptr)
    const double* pos_fpos;
};

struct HbParams {
    int ntaps;
    long long e0, e1;
    double taps[14];
};

// Fused chain of up to 6 half-band 2x DOWNsamplers (k_hbdown_cascade): every intermediate rate lives in shared memory.
// Stream s is the input of stage s (stream 0 = the cascade's source, stream n_stages = its output).
struct HbDownCascParams {
    int n_stages;
    int ntaps[6];
    double taps[6][14];
    long long e0, e1;      // output indices of the LAST stage to write
    int w;                 // final outputs per tile
    int n_tiles;
    int back[7];           // stream-s samples needed below 2^(n-s) * m for final output m (back[n] = 0)
    int boff[7];           // offsets (doubles) of the per-stream buffers in dynamic shared memory: even half, then odd half
    int cap[7];            // doubles per half of stream s's buffer
};
int hbdown_cascade_plan(HbDownCascParams& p, int smem_budget_doubles); // fills w, back, boff, cap; returns smem bytes (0: does not fit)
void launch_hbdown_cascade(const HbDownCascParams& p, int smem_bytes, const SrcView& src, const DstView& dst, int n_ch, cudaStream_t st);

// Fused chain of up to 6 half-band 2x upsamplers (k_hbup_cascade).
struct HbCascadeParams {
    int n_stages;
    int ntaps[6];
    double taps[6][14];
    long long e0, e1;      // output indices of the LAST stage to write
    long long a0;          // first tile starts at this position of the cascade's input stream
    int w;                 // tile width in input-stream samples
    int n_tiles;
    int lo_off[7], hi_off[7]; // stage-k stream range a tile needs: [2^k*A - lo_off[k], 2^k*(A+w) + hi_off[k])
    int boff[7];           // offsets (doubles) of the per-stage buffers in dynamic shared memory
    int fuse_last2;        // stages c-2 and c-1 run as one pass (no buffer for stream c-1)
};
bool hb_last2_supported(int t1, int t2);
void launch_hbup_cascade(const HbCascadeParams& p, int smem_bytes, const SrcView& src, const DstView& dst,
                         int n_ch, cudaStream_t st);

// Fused 2x BlockConvolver + fractional interpolator (r8b_fused.cu).  Positions are indices of the
// 2x-rate stream between the two stages.
struct FusedParams {
    int mode;              // 0 whole stepping, 1 order-2 bank, 2 (v2 kernel only) no interpolator: the 2x stream itself is the output
    int n_tiles;           // tiles of `span` owned positions each, processed in pairs
    int stage_off;         // offset (doubles) of the store staging area in dynamic smem, 0 = none
    int debug;             // profiling experiments only (R8BGPU_DEBUG): bit0 skip interp stores, bit1 skip tap loop
    unsigned long long* prof; // optional: 8 phase-cycle accumulators (R8BGPU_PROFILE), else nullptr
    int span;              // even
    long long p_lo, p_hi;  // owned position range of this call [p_lo, p_hi), p_lo even
    int yl;                // left margin of a tile's valid range (>= fll, even)
    int lg;                // half support of the polyphase low-pass, input samples
    int ysh;               // y layout: index i lives at i + (i >> ysh) (31 = plain)
    const double2* spec;   // low-pass spectrum, slot order, pre-scaled by 1/(2M)
    const double2* tw;     // exp(-2*pi*i*k/M)
    // interpolator
    const double* bank;
    int bank_len, bank_in_smem;
    int flen, fll;
    long long e0, e1;      // outputs of this call
    int in_step, out_step;
    const double* gbank;   // whole stepping: for EVERY first phase r0 in [0,out_step): [smaxp][ir] shifted, zero-padded taps
                           // of phases r0..r0+ir-1 (phases past out_step-1 continue in the next stepping cycle)
    int gbank_len, smaxp;  // doubles in gbank; padded window length (multiple of 4)
    int ir;                // phases per group (8)
    int gbank_smem_len;    // doubles of the per-call bank selection kept in shared memory (n_groups * smaxp * ir)
    int delta, wrap;       // first phase of group 0 (= e0 mod 8 when out_step % 8 == 0, else 0); wrap = rows may span cycles
    const int* goff;       // [out_step] floor(r0*in_step/out_step)
    const int* phase_off;  // [out_step] floor(r*in_step/out_step)
    const int* phase_row;  // [out_step] (r*in_step) % out_step
    int fracs;
    double ssr, dsr;
    int in_counter0, in_pos_int0;
    double in_pos_shift, fpos0;
    long long p0;
    const int* pos_dp;       // R8B_FASTTIMING tables (else nullptr)
    const double* pos_fpos;
    // order-2 bank: when the bank row index drifts slowly and monotonically (mod fracs) with the output index,
    // the rows a tile pair needs are a short circular run that is staged in shared memory
    int poly_dir;            // +1 rows ascend with k, -1 descend, 0 no staging
    int poly_rows_cap;       // rows of shared memory available for the run
    int poly_row_stride;     // doubles between staged rows: 3*flen padded so that consecutive rows start 4 (mod 8)
                             // banks apart -- lanes that sit on different rows then load without bank conflicts
    int poly_n;              // > 0: input positions advance by ~poly_n per output; threads take 4 consecutive outputs
    int poly_chunks;         // a pair's outputs are processed in this many pieces, each with its own run (>= 1)
    // v2 kernel (r8b_fused2.cu): persistent CTAs, one tile per half-CTA
    int n_ch;                // channels of this launch (work units = n_ch * n_tiles)
    int glog;                // log2 of the phase groups a warp covers per instruction (lanes = 32>>glog cycles x 1<<glog groups)
    int flags;               // bit 0: ping-pong token around the interpolation, bit 1: bulk-copy input tiles
    const double2* tw_tab;   // 512 entries: tw2t[q*16+r] = W_256^(r q), then tw1t[q*16+r] = W_M^(r q)
    int mbu;                 // tensor-path interpolation: blocks of 8 stepping cycles per work unit (2..4; 0 = 3)
    int up;                  // BlockConvolver up-factor of the fused pair: 2 (default, also when 0) or 1
    int ylen;                // doubles of the tile's stream between the two stages held in shared memory (2*FM for up 2, FM for up 1)
    const double2* cd_tab;   // up == 2, phase C fused into the first inverse pass: [q3 < 16][g < 256] spectrum at slot 16 g + q3,
                             // then [g < 256] W_M^((g >> 4) + 16 (g & 15))
    const double2* c_tab;    // up == 1: phase C operands in the order the threads consume them: [u < 4][item < 3][ht < 256] --
                             // W_M^k, H[k]/2, H[N-k]/2 for k = c_freq(ht, u) -- then 3 entries for k = N/2 (v1 kernel: its own table)
};
int fused_smem_bytes(int bank_doubles_in_smem);
int fused_max_span(int lg, int yl, int yr);
int fused_stage_doubles();
int fused_fixed_doubles();
int fused_poly_queue_bytes();
void launch_up2_frac(const FusedParams& p, const SrcView& src, const DstView& dst, int n_ch, cudaStream_t st);
// Calibration: best-of-3 TFLOP/s of a register-resident DFMA stream (16 independent chains per thread, 32 warps per
// SM) on the current device -- the fp64 ceiling the bench reports next to the HBM roofline.  < 0 on error.
double measure_dfma_tflops();

constexpr int kFused2SmemMax = 227 * 1024 - 1024; // dynamic part; the kernel's static shared memory is < 1 KB
int fused2_smem_bytes(int bank_doubles, bool staged);
int fused2_stage_off(int bank_doubles);
void launch_up2_frac2(const FusedParams& p, const SrcView& src, const DstView& dst, int n_sm, cudaStream_t st);

int blockconv_smem_bytes(int fft_log2, int up);
cudaError_t blockconv_configure(); // opt-in shared memory attributes; call once per device

void launch_blockconv(const BlockConvParams& p, const SrcView& src, const DstView& dst, int n_ch,
                      cudaStream_t st);
void launch_frac_whole(const FracParams& p, const SrcView& src, const DstView& dst, int n_ch,
                       cudaStream_t st);
void launch_frac_poly(const FracParams& p, const SrcView& src, const DstView& dst, int n_ch,
                      cudaStream_t st);
void launch_hbup(const HbParams& p, const SrcView& src, const DstView& dst, int n_ch, cudaStream_t st);
void launch_hbdown(const HbParams& p, const SrcView& src, const DstView& dst, int n_ch, cudaStream_t st);
// copy cur[n0..n1) into the ring (history for later calls)
void launch_save_tail(const double* cur, long long cur_stride, long long cur_base, long long n0,
                      long long n1, double* ring, long long ring_stride, long long ring_mask, int n_ch,
                      cudaStream_t st, int fmt = 0, double scale = 1.0);

// Caller-side sample formats (r8b_format.cu); values match r8bgpu_sample_format in include/r8bgpu.h.
__host__ __device__ int format_bytes(int fmt); // 0: unknown format
// raw (any format; planar: channel c at c*raw_stride, interleaved: frame f at f*raw_stride) <-> planar fp64
bool launch_to_f64(int fmt, const void* raw, bool interleaved, size_t raw_stride, double* f64, size_t f64_stride,
                   int n, int n_ch, double scale, cudaStream_t st);
bool launch_from_f64(int fmt, void* raw, bool interleaved, size_t raw_stride, const double* f64, size_t f64_stride,
                     int n, int n_ch, double scale, cudaStream_t st);

} // namespace r8bgpu
