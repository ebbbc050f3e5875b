/* tests/cpp/dll_demo.c -- a plain C caller of the reference's DLL interface (DLL/r8bsrc.h), linked against
 * libr8bsrc.so:  dll_demo <in.f64> <out.f64> <frames> <src> <dst> <block> <res>   (one channel of raw doubles)
 * "dll_demo --inlen <src> <dst> <block> <res> <n>" prints r8b_inlen() (no GPU needed). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "r8b/DLL/r8bsrc.h"

int main(int argc, char** argv)
{
    if (argc == 7 && strcmp(argv[1], "--inlen") == 0) {
        CR8BResampler rs = r8b_create(atof(argv[2]), atof(argv[3]), atoi(argv[4]), 2.0, (enum ER8BResamplerRes) atoi(argv[5]));
        printf("%d\n", r8b_inlen(rs, atoi(argv[6])));
        r8b_delete(rs);
        return 0;
    }
    if (argc != 8) return 2;
    {
        const int frames = atoi(argv[3]), block = atoi(argv[6]);
        double* in = (double*) malloc(sizeof(double) * (size_t) frames);
        FILE* f = fopen(argv[1], "rb");
        CR8BResampler rs;
        long long total = 0;
        int pos;
        if (!f || fread(in, sizeof(double), (size_t) frames, f) != (size_t) frames) return 3;
        fclose(f);
        rs = r8b_create(atof(argv[4]), atof(argv[5]), block, 2.0, (enum ER8BResamplerRes) atoi(argv[7]));
        f = fopen(argv[2], "wb");
        for (pos = 0; pos < frames; pos += block) {
            const int l = frames - pos < block ? frames - pos : block;
            double* op = NULL;
            const int n = r8b_process(rs, in + pos, l, &op);
            if (n > 0) fwrite(op, sizeof(double), (size_t) n, f);
            total += n;
            if (pos == 0) r8b_clear(rs), r8b_process(rs, in + pos, l, &op); /* clear() + the same block again: same result */
        }
        fclose(f);
        if (total == 0 && r8b_last_error()[0]) fprintf(stderr, "%s\n", r8b_last_error());
        r8b_delete(rs);
        free(in);
        printf("%lld\n", total);
    }
    return 0;
}
