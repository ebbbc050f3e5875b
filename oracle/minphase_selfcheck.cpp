// oracle/minphase_selfcheck.cpp -- evidence program (test infrastructure, not product).
//
// Runs the UNMODIFIED reference in minimum-phase mode (fprMinPhase) on a fixed noise signal and dumps the
// fp64 output.  Built twice -- PFFFT-double and the default Ooura FFT -- the two dumps differ by
// max 9.7e-5 / rms 2.3e-5 (signal rms 0.57): the cepstral transform (CDSPRealFFT.h:681-785) takes
// log|H| of stop-band bins that are pure FFT rounding noise, so the reference's min-phase kernel is only
// defined to about -88 dB.  See DESIGN.md "Minimum phase".
//
//   g++ -O2 -ffp-contract=off -std=c++17 -I$REF -DR8B_PFFFT_DOUBLE=1 minphase_selfcheck.cpp \
//       -x c $REF/fft/pffft_double.c -o _ref/mp_pffft && ./_ref/mp_pffft > /tmp/a.bin
//   g++ -O2 -ffp-contract=off -std=c++17 -I$REF minphase_selfcheck.cpp -o _ref/mp_ooura && ./_ref/mp_ooura > /tmp/b.bin
#include <cstdio>
#include <cmath>
#include <vector>
#include "CDSPResampler.h"
using namespace r8b;
int main(){
    CDSPResampler rs(44100.0, 96000.0, 4096, 2.0, 180.15, fprMinPhase);
    std::vector<double> x(4096*8);
    unsigned long long s=88172645463325252ULL;
    for(auto&v:x){ s^=s<<13; s^=s>>7; s^=s<<17; v=(double)(s>>11)/9007199254740992.0*2-1; }
    std::vector<double> out;
    for(int b=0;b<8;b++){ double*op; int n=rs.process(&x[b*4096],4096,op); out.insert(out.end(),op,op+n);}
    fwrite(out.data(),8,out.size(),stdout);
    fprintf(stderr,"n=%zu\n",out.size());
}
