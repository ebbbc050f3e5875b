// r8bsrc_shim.cpp -- libr8bsrc.so: the reference's DLL entry points (DLL/r8bsrc.cpp:64-107) on top of the header
// front-end (include/r8b/CDSPResampler.h), which in turn calls the C-ABI of libr8bgpu.so.  Built by
// r8brain-free-src_b200/build.py next to libr8bgpu.so.
#include "../../include/r8b/DLL/r8bsrc.h"

#include "../../include/r8b/CDSPResampler.h"

using namespace r8b;

extern "C" {

CR8BResampler r8b_create(double SrcSampleRate, double DstSampleRate, int MaxInLen, double ReqTransBand, enum ER8BResamplerRes Res)
{
    if (Res == r8brr16) return new CDSPResampler16(SrcSampleRate, DstSampleRate, MaxInLen, ReqTransBand);
    if (Res == r8brr16IR) return new CDSPResampler16IR(SrcSampleRate, DstSampleRate, MaxInLen, ReqTransBand);
    return new CDSPResampler24(SrcSampleRate, DstSampleRate, MaxInLen, ReqTransBand);
}

void r8b_delete(CR8BResampler rs) { delete (CDSPResampler*) rs; }

int r8b_inlen(CR8BResampler rs, int ReqOutSamples) { return ((CDSPResampler*) rs)->getInputRequiredForOutput(ReqOutSamples); }

void r8b_clear(CR8BResampler rs) { ((CDSPResampler*) rs)->clear(); }

int r8b_process(CR8BResampler rs, double* ip0, int l, double** op0)
{
    double* op = 0;
    const int n = ((CDSPResampler*) rs)->process(ip0, l, op);
    *op0 = op;
    return n;
}

const char* r8b_last_error(void) { return r8bgpu_last_error(); }

} // extern "C"
