// tools/mb_dmma.cu -- is the fp64 tensor path (mma.sync m8n8k4 f64 = SASS DMMA) worth using for the whole-step
// interpolator?  (1) raw DMMA throughput, (2) the interpolation as a strided-Hankel GEMM out of shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mb_dmma tools/mb_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void k_dmma_raw(double* out, int iters)
{
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; i++) c[i][0] = c[i][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// interpolation of one phase group: out[c][r] = sum_s y[c*in_step + o + s] * Bp[s][r], r < 8, s < 32, cycles c in MB blocks of 8
constexpr int YLEN = 8704, NG = 20, SM_TAPS = 32;
template <int MB>
__global__ void __launch_bounds__(512, 1) k_interp_dmma(double* out, const double* gbank, int iters, int in_step)
{
    extern __shared__ double sm[];
    double* y = sm;
    double* sbank = sm + YLEN;
    for (int i = threadIdx.x; i < YLEN; i += blockDim.x) y[i] = 1e-3 * i;
    for (int i = threadIdx.x; i < NG * SM_TAPS * 8; i += blockDim.x) sbank[i] = gbank[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int arow = lane >> 2, acol = lane & 3;   // A fragment: row = lane/4, col = lane%4
    const int bk = lane & 3, bn = lane >> 2;       // B fragment: k = lane%4, n = lane/4
    double tot = 0.0;
    for (int it = 0; it < iters; it++) {
        const int g = (warp + it) % NG;
        int yo[MB];
#pragma unroll
        for (int m = 0; m < MB; m++) yo[m] = (((8 * m + arow) * in_step + it * 7) % (YLEN - 64)) + acol;
        double c[MB][2];
#pragma unroll
        for (int m = 0; m < MB; m++) c[m][0] = c[m][1] = 0.0;
        const double* gb = sbank + g * SM_TAPS * 8 + bk * 8 + bn;
#pragma unroll
        for (int ks = 0; ks < SM_TAPS / 4; ks++) {
            const double b = gb[ks * 32];
#pragma unroll
            for (int m = 0; m < MB; m++) dmma(c[m][0], c[m][1], y[yo[m] + 4 * ks], b);
        }
#pragma unroll
        for (int m = 0; m < MB; m++) tot += c[m][0] + c[m][1];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = tot;
}

template <typename F>
float timeit(F f)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main()
{
    double *out, *gbank;
    cudaMalloc(&out, 148 * 1024 * 8);
    cudaMalloc(&gbank, NG * SM_TAPS * 8 * 8);
    cudaMemset(gbank, 0, NG * SM_TAPS * 8 * 8);
    const int iters = 8192;
    for (int nt : {128, 256, 512, 1024}) {
        float m1 = timeit([&] { k_dmma_raw<1><<<148, nt>>>(out, iters); });
        float m4 = timeit([&] { k_dmma_raw<4><<<148, nt>>>(out, iters); });
        float m8 = timeit([&] { k_dmma_raw<8><<<148, nt>>>(out, iters); });
        auto tf = [&](float ms, int ilp) { return 2.0 * 256 * (double) ilp * iters * (nt / 32) * 148 / (ms * 1e-3) / 1e12; };
        printf("DMMA m8n8k4 raw, %2d warps/SM: ILP1 %.2f TF  ILP4 %.2f TF  ILP8 %.2f TF\n", nt / 32, tf(m1, 1), tf(m4, 4), tf(m8, 8));
    }
    const int smem = (YLEN + NG * SM_TAPS * 8) * 8, it2 = 2000;
    cudaFuncSetAttribute(k_interp_dmma<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_interp_dmma<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int nt : {256, 512}) {
        float a = timeit([&] { k_interp_dmma<6><<<148, nt, smem>>>(out, gbank, it2, 147); });
        float b = timeit([&] { k_interp_dmma<3><<<148, nt, smem>>>(out, gbank, it2, 147); });
        // one warp-iteration = MB*8 cycles x 8 phases x 32 taps
        auto clk_per_out = [&](float ms, int mb) { return ms * 1e-3 * 1.965e9 / ((double) it2 * (nt / 32) * mb * 64); };
        printf("interp via DMMA, nt %d: MB=6 %.3f clk per output (24-tap equiv %.1f G out/s/GPU), MB=3 %.3f clk per output\n", nt,
               clk_per_out(a, 6), 148 * 1.965 / clk_per_out(a, 6), clk_per_out(b, 3));
    }
    printf("%s (reference: register-tiled DFMA loop = 18 clk per warp-tap per 768 outputs/32 taps -> %.3f clk per output)\n",
           cudaGetErrorString(cudaGetLastError()), 18.0 * 32 / 768);
    return 0;
}
