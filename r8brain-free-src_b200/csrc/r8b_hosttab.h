// r8b_hosttab.h -- host-side tables and tile geometry of the device kernels, shared by the engine
// (r8b_capi.cu) and by the CPU emulation of the fused kernel that the tests run without a GPU
// (tests/cpp/fused2_emul.cu).  Nothing here touches the device.
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "r8b_kernels.h"
#include "r8b_plan.h"

namespace r8bgpu {

// Low-pass spectrum of a BlockConvolver stage (CDSPFIRFilter.h:492-520 as arithmetic: direct long-double DFT of
// the designed taps, polyphase-packed g_0 + i g_1 for 2x stages, pre-scaled by 1/M resp. 1/(2M)) in FFT slot order,
// and the twiddle table exp(-2 pi i k / M).
void build_spectrum(const StageDesc& s, int fft_log2, std::vector<double2>& spec_slots, std::vector<double2>& tw,
                    double* nyq_gain);

// [q][r] twiddle tables of the fused kernels: 256 entries W_256^(r q), then 256 entries W_4096^(r q)
std::vector<double2> build_tw_tab(const std::vector<double2>& tw4096);
// phase C operands of the v2 fused kernel's 1x pair in thread order (FusedParams::c_tab, the layout c1_pair_tab() reads);
// empty for up = 2, whose phase C runs inside the first inverse pass (build_cd_tab)
std::vector<double2> build_c_tab(const std::vector<double2>& spec_slots4096, const std::vector<double2>& tw4096, int up);

// "2x BlockConvolver -> FracInterpolator" pair: margins and span of the M = 4096 tiles
struct FusedGeom {
    bool ok = false;
    int lg = 0, yl = 0, yr = 0, span_max = 0, ysh = 31;
    int up = 2; // up-factor of the BlockConvolver: 2, or 1 (v2 kernel with the tensor-path interpolation only)
};
FusedGeom fused_geometry(const StageDesc& bc, const StageDesc& frac);

// Whole-stepping bank re-laid for the fused kernels: for EVERY first phase r0 a [smaxp][ir] block holding the
// filters of phases r0..r0+ir-1 pre-shifted by their window offsets and zero-padded (CDSPFracInterpolator.h:991-1060
// reads bank[(j*InStep) % OutStep] at window floor(j*InStep/OutStep)).
struct GroupBank {
    int ir = 8, smaxp = 0, n_groups = 0;
    std::vector<double> gb;   // [out_step][smaxp][ir]
    std::vector<int> go;      // [out_step] floor(r0*in_step/out_step)
    std::vector<int> off, row; // per phase: window offset, bank row
};
int choose_group_ir(const StageDesc& frac);
// operands of the v2 kernel's fused phase C + first inverse pass (FusedParams::cd_tab)
std::vector<double2> build_cd_tab(const std::vector<double2>& spec_slots4096, const std::vector<double2>& tw4096);
// the same for the round-1 fused kernel (tile pairs, 512 threads): [u < 4][item < 2][tid < 512] = spectrum at the slot
// s1 = 16*((tid>>3) + 64u) + (tid&7) and at the slot of the mirrored frequency
std::vector<double2> build_c_tab_v1(const std::vector<double2>& spec_slots4096);

// frag_order (ir == 8 only): within every block of 4 taps the 32 values are stored as [phase][tap % 4] -- the B-fragment
// order of mma.sync m8n8k4, so a warp's load of one K-step is 256 contiguous bytes
GroupBank build_group_bank(const StageDesc& frac, int ir, bool frag_order = false);

// Whole-stepping call: the fields of FusedParams that follow from the interpolator stage and this call's output
// range [e0, e1) alone (positions are indices of the 2x-rate stream; p_lo even).
void fused_whole_fields(FusedParams& p, const StageDesc& frac, long long e0, long long e1);

// Per-call tile geometry of the v2 fused kernel (one tile per half-CTA): fills p.p_lo (input: first needed position,
// even), p.n_tiles, p.span.  cur_parity >= 0: parity of the caller's block base index, so that FFT windows start on
// 16-byte boundaries of the block (bulk-copied input tiles); -1: no constraint.
void fused2_tiles(FusedParams& p, const FusedGeom& g, int cur_parity);
int fused2_choose_glog(int span, int in_step, int out_step, int ir);
// blocks of 8 stepping cycles per work unit of the tensor-path interpolation: fewest (rounds over a half's 8 warps) x blocks
int fused2_choose_mbu(int span, int in_step, int out_step);

} // namespace r8bgpu
