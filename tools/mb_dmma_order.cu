// tools/mb_dmma_order.cu -- in which order does DMMA.8x8x4 accumulate its four products?  Random fragments, device
// result compared bit for bit with candidate orders evaluated in fp64 on the host.
#include <cstdio>
#include <cstring>
#include <cmath>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void k(const double* a, const double* b, const double* c, double* d)
{
    const int lane = threadIdx.x;
    double c0 = c[2 * lane], c1 = c[2 * lane + 1];
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a[lane]), "d"(b[lane]));
    d[2 * lane] = c0;
    d[2 * lane + 1] = c1;
}

int main()
{
    double ha[32], hb[32], hc[64], hd[64], *da, *db, *dc, *dd;
    cudaMalloc(&da, 256); cudaMalloc(&db, 256); cudaMalloc(&dc, 512); cudaMalloc(&dd, 512);
    int hits[4] = {0, 0, 0, 0}, total = 0;
    srand(7);
    for (int trial = 0; trial < 200; trial++) {
        for (int i = 0; i < 32; i++) { ha[i] = (rand() / (double) RAND_MAX - 0.5) * exp2((double) (rand() % 20)); hb[i] = rand() / (double) RAND_MAX - 0.5; }
        for (int i = 0; i < 64; i++) hc[i] = (rand() / (double) RAND_MAX - 0.5) * 100;
        cudaMemcpy(da, ha, 256, cudaMemcpyHostToDevice); cudaMemcpy(db, hb, 256, cudaMemcpyHostToDevice); cudaMemcpy(dc, hc, 512, cudaMemcpyHostToDevice);
        k<<<1, 32>>>(da, db, dc, dd);
        cudaMemcpy(hd, dd, 512, cudaMemcpyDeviceToHost);
        for (int lane = 0; lane < 32; lane++)
            for (int e = 0; e < 2; e++) {
                const int row = lane >> 2, col = 2 * (lane & 3) + e;
                double A[4], B[4];
                for (int q = 0; q < 4; q++) { A[q] = ha[4 * row + q]; B[q] = hb[4 * col + q]; }
                const double c = hc[2 * lane + e], got = hd[2 * lane + e];
                double s0 = c; for (int q = 0; q < 4; q++) s0 = fma(A[q], B[q], s0);                 // ascending chain from c
                double s1 = c; for (int q = 3; q >= 0; q--) s1 = fma(A[q], B[q], s1);                // descending chain
                double s2 = fma(A[0], B[0], 0.0); for (int q = 1; q < 4; q++) s2 = fma(A[q], B[q], s2); s2 += c; // products first, c last
                double s3 = fma(A[1], B[1], A[0] * B[0]) + fma(A[3], B[3], A[2] * B[2]) + c;         // pairwise
                hits[0] += memcmp(&s0, &got, 8) == 0; hits[1] += memcmp(&s1, &got, 8) == 0;
                hits[2] += memcmp(&s2, &got, 8) == 0; hits[3] += memcmp(&s3, &got, 8) == 0;
                total++;
            }
    }
    printf("of %d results: ascending FMA chain from C %d | descending %d | products then +C %d | pairwise %d\n", total, hits[0], hits[1], hits[2], hits[3]);
    return 0;
}
