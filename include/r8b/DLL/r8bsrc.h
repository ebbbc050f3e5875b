/* include/r8b/DLL/r8bsrc.h -- the reference's plain-C style DLL interface (DLL/r8bsrc.h:31-134 of
 * avaneev/r8brain-free-src: r8b_create / r8b_delete / r8b_inlen / r8b_clear / r8b_process) over this engine.
 *
 * Same names, argument order, enum values and ownership rules: the handle is opaque; r8b_process() hands back a
 * pointer into a buffer owned by the resampler (valid until the next call), or the input pointer itself when the rates
 * are equal.  The reference declares the output parameter as a C++ reference inside extern "C" (`double*& op0`,
 * DLL/r8bsrc.h:131-132); callers compiled against that header pass the address of their pointer, which is the ABI of
 * `double** op0` -- the C-clean spelling used here (C++ callers that wrote `op` keep compiling through the inline
 * overload at the end).  Build: csrc/r8bsrc_shim.cpp -> libr8bsrc.so (links libr8bgpu.so).  No CPU fallback: without
 * a CUDA device r8b_process() returns 0 samples and r8b_last_error() (an addition) says why.
 */
#ifndef R8BSRC_B200_INCLUDED
#define R8BSRC_B200_INCLUDED

#if defined(_WIN32)
#define R8BSRC_DECL __declspec(dllexport)
#else
#define R8BSRC_DECL __attribute__((visibility("default")))
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef void* CR8BResampler;

enum ER8BResamplerRes {
    r8brr16 = 0,   /* 16-bit precision resampler */
    r8brr16IR = 1, /* 16-bit precision resampler for impulse responses */
    r8brr24 = 2    /* 24-bit precision resampler (including 32-bit floating point) */
};

R8BSRC_DECL CR8BResampler r8b_create(double SrcSampleRate, double DstSampleRate, int MaxInLen, double ReqTransBand,
                                     enum ER8BResamplerRes Res);
R8BSRC_DECL void r8b_delete(CR8BResampler rs);
R8BSRC_DECL int r8b_inlen(CR8BResampler rs, int ReqOutSamples);
R8BSRC_DECL void r8b_clear(CR8BResampler rs);
R8BSRC_DECL int r8b_process(CR8BResampler rs, double* ip0, int l, double** op0);
/* Not in the reference: the reason for the last failure on this thread ("" if none). */
R8BSRC_DECL const char* r8b_last_error(void);

#ifdef __cplusplus
} /* extern "C" */
/* source compatibility with the reference's `double*&` spelling */
inline int r8b_process(CR8BResampler rs, double* ip0, int l, double*& op0) { return r8b_process(rs, ip0, l, &op0); }
#endif

#endif /* R8BSRC_B200_INCLUDED */
