// tests/cpp/shim_demo.cpp -- uses include/r8b/CDSPResampler.h exactly the way example.cpp:30-67 uses the
// reference header: one CDSPResampler24 per channel, the same block length for every channel.
//   shim_demo <in.f64> <out.f64> <n_ch> <frames> <src> <dst> <block>      (planar raw doubles)
// Without a GPU, "shim_demo --plan <src> <dst> <block>" prints only plan-level getters.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "r8b/CDSPResampler.h"

int main(int argc, char** argv)
{
    if (argc == 5 && strcmp(argv[1], "--plan") == 0) {
        r8b::CDSPResampler24 rs(atof(argv[2]), atof(argv[3]), atoi(argv[4]));
        printf("%d %d %d %g\n", rs.getMaxOutLen(0), rs.getInLenBeforeOutPos(0), rs.getInputRequiredForOutput(1000),
               rs.getLatencyFrac());
        return 0;
    }
    if (argc == 2 && strcmp(argv[1], "--failures") == 0) {
        // objects that cannot work must say so and must not hang: (a) a refused plan (minimum phase), (b) on a box
        // without a CUDA device, a valid plan whose batch cannot be created
        r8b::CDSPResampler bad(44100.0, 96000.0, 1024, 2.0, 180.15, r8b::fprMinPhase);
        float out[64];
        short in[16] = {1, 2, 3};
        bad.oneshot(in, 16, out, 64);
        int zeros = 0;
        for (int i = 0; i < 64; i++) zeros += out[i] == 0.0f;
        printf("%d %d %d %d %d\n", (int) bad.isValid(), bad.getInLenBeforeOutStart(0), zeros, bad.getInLenBeforeOutPos(0),
               bad.getInputRequiredForOutput(10));
        r8b::CDSPResampler24 ok(44100.0, 96000.0, 1024);
        const int before = (int) ok.isValid();
        const int start = ok.getInLenBeforeOutStart(0); // terminates with or without a device
        printf("%d %d %d\n", before, (int) ok.isValid(), start);
        return 0;
    }
    if (argc == 9 && strcmp(argv[8], "--pull") == 0) {
        // pull mode: the same stream requested in odd-sized pieces must equal the push-mode output
        const int n_ch = atoi(argv[3]), frames = atoi(argv[4]), block = atoi(argv[7]);
        std::vector<double> in((size_t) n_ch * frames);
        FILE* f = fopen(argv[1], "rb");
        if (!f || fread(in.data(), sizeof(double), in.size(), f) != in.size()) return 3;
        fclose(f);
        r8b::CDSPResamplerPull rs(n_ch, atof(argv[5]), atof(argv[6]), block, 2.0, 180.15);
        int pos = 0;
        auto fill = [&](double* ip, size_t stride, int maxFrames) {
            const int l = frames - pos < maxFrames ? frames - pos : maxFrames;
            for (int c = 0; c < n_ch; c++) memcpy(ip + c * stride, &in[(size_t) c * frames + pos], sizeof(double) * (size_t) l);
            pos += l;
            return l;
        };
        std::vector<std::vector<double> > out((size_t) n_ch);
        std::vector<double> piece((size_t) n_ch * 1000);
        for (int req = 1;; req = req % 997 + 101) {
            const int got = rs.pull(fill, piece.data(), 1000, req > 1000 ? 1000 : req);
            if (got <= 0) break;
            for (int c = 0; c < n_ch; c++) out[(size_t) c].insert(out[(size_t) c].end(), piece.begin() + c * 1000, piece.begin() + c * 1000 + got);
        }
        f = fopen(argv[2], "wb");
        for (int c = 0; c < n_ch; c++) {
            const long long n = (long long) out[(size_t) c].size();
            fwrite(&n, sizeof n, 1, f);
            fwrite(out[(size_t) c].data(), sizeof(double), (size_t) n, f);
        }
        fclose(f);
        return 0;
    }
    if (argc == 9 && strcmp(argv[8], "--batch") == 0) {
        // r8b::CDSPResamplerBatch with Device = -1: every visible GPU behind one object (R8BGPU_FORCE_SHARDS on one GPU),
        // buffers from r8bgpu_batch_host_alloc (rows on the owning GPU's NUMA node)
        const int n_ch = atoi(argv[3]), frames = atoi(argv[4]), block = atoi(argv[7]);
        std::vector<double> in((size_t) n_ch * frames);
        FILE* f = fopen(argv[1], "rb");
        if (!f || fread(in.data(), sizeof(double), in.size(), f) != in.size()) return 3;
        fclose(f);
        r8b::CDSPResamplerBatch rs(n_ch, atof(argv[5]), atof(argv[6]), block, 2.0, 180.15);
        const int cap = rs.getMaxOutLen();
        double* hin = (double*) r8bgpu_batch_host_alloc(rs.handle(), (size_t) block, 8);
        double* hout = (double*) r8bgpu_batch_host_alloc(rs.handle(), (size_t) cap, 8);
        if (!hin || !hout) return 4;
        std::vector<std::vector<double> > out((size_t) n_ch);
        for (int pos = 0; pos < frames; pos += block) {
            const int l = frames - pos < block ? frames - pos : block;
            for (int c = 0; c < n_ch; c++) memcpy(hin + (size_t) c * block, &in[(size_t) c * frames + pos], sizeof(double) * (size_t) l);
            const int n = rs.process(hin, (size_t) block, l, hout, (size_t) cap, cap);
            if (n < 0) return 5;
            for (int c = 0; c < n_ch; c++) out[(size_t) c].insert(out[(size_t) c].end(), hout + (size_t) c * cap, hout + (size_t) c * cap + n);
        }
        f = fopen(argv[2], "wb");
        for (int c = 0; c < n_ch; c++) {
            const long long n = (long long) out[(size_t) c].size();
            fwrite(&n, sizeof n, 1, f);
            fwrite(out[(size_t) c].data(), sizeof(double), (size_t) n, f);
        }
        fclose(f);
        printf("%d\n", r8bgpu_batch_shard_count(rs.handle()));
        r8bgpu_host_free(hin);
        r8bgpu_host_free(hout);
        return 0;
    }
    if (argc != 8) return 2;
    const int n_ch = atoi(argv[3]), frames = atoi(argv[4]), block = atoi(argv[7]);
    const double src = atof(argv[5]), dst = atof(argv[6]);
    std::vector<double> in((size_t) n_ch * frames);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(in.data(), sizeof(double), in.size(), f) != in.size()) return 3;
    fclose(f);
    std::vector<r8b::CDSPResampler24*> rs;
    for (int c = 0; c < n_ch; c++) rs.push_back(new r8b::CDSPResampler24(src, dst, block));
    std::vector<std::vector<double> > out((size_t) n_ch);
    for (int pos = 0; pos < frames; pos += block) {
        const int l = frames - pos < block ? frames - pos : block;
        for (int c = 0; c < n_ch; c++) {
            double* op;
            const int n = rs[(size_t) c]->process(&in[(size_t) c * frames + pos], l, op);
            out[(size_t) c].insert(out[(size_t) c].end(), op, op + n);
        }
    }
    f = fopen(argv[2], "wb");
    for (int c = 0; c < n_ch; c++) {
        const long long n = (long long) out[(size_t) c].size();
        fwrite(&n, sizeof n, 1, f);
        fwrite(out[(size_t) c].data(), sizeof(double), (size_t) n, f);
        delete rs[(size_t) c];
    }
    fclose(f);
    return 0;
}
