// tools/mb_interp.cu -- the whole-step interpolation tap loop in isolation: where should the bank live?
//   A: bank in shared memory (broadcast LDS.128), B: bank in __constant__ memory (indexed LDC), C: bank via __ldg (L1)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mb_interp tools/mb_interp.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int SMAXP = 32, NG = 20, YLEN = 8704;
__constant__ double cbank[NG * SMAXP * 8];

template <int MODE, int IR, int IQ>
__global__ void __launch_bounds__(512, 1) k_interp(double* out, const double* gbank, int iters, int in_step)
{
    extern __shared__ double sm[];
    double* y = sm;
    double* sbank = sm + YLEN;
    for (int i = threadIdx.x; i < YLEN; i += blockDim.x) y[i] = 1e-3 * i;
    for (int i = threadIdx.x; i < NG * SMAXP * 8; i += blockDim.x) sbank[i] = gbank[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double tot = 0.0;
    for (int it = 0; it < iters; it++) {
        const int g = (warp + it) % NG;
        int yo[IQ];
#pragma unroll
        for (int q = 0; q < IQ; q++) yo[q] = ((lane + 32 * q + it) * in_step) % (YLEN - SMAXP - 64);
        double acc[IR][IQ];
#pragma unroll
        for (int r = 0; r < IR; r++)
#pragma unroll
            for (int q = 0; q < IQ; q++) acc[r][q] = 0.0;
        const double* gb = (MODE == 0 ? sbank : MODE == 1 ? cbank : gbank) + g * SMAXP * 8;
#pragma unroll 4
        for (int s = 0; s < SMAXP; s++) {
            double yv[IQ];
#pragma unroll
            for (int q = 0; q < IQ; q++) yv[q] = y[yo[q] + s];
#pragma unroll
            for (int r = 0; r < IR; r += 2) {
                double2 b;
                if (MODE == 2) b = __ldg(reinterpret_cast<const double2*>(gb + s * 8 + r));
                else b = *reinterpret_cast<const double2*>(gb + s * 8 + r);
#pragma unroll
                for (int q = 0; q < IQ; q++) {
                    acc[r][q] = fma(b.x, yv[q], acc[r][q]);
                    acc[r + 1][q] = fma(b.y, yv[q], acc[r + 1][q]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < IR; r++)
#pragma unroll
            for (int q = 0; q < IQ; q++) tot += acc[r][q];
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

template <int MODE, int IR, int IQ>
void run(const char* name, double* out, double* gbank, int nt)
{
    const int smem = (YLEN + NG * SMAXP * 8) * 8, iters = 2000;
    cudaFuncSetAttribute(k_interp<MODE, IR, IQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float ms = timeit([&] { k_interp<MODE, IR, IQ><<<148, nt, smem>>>(out, gbank, iters, 147); });
    const double clk = ms * 1e-3 * 1.965e9 / iters;          // per task-iteration of all warps of the SM
    const double per_warp_iter = clk / (nt / 32) / SMAXP;    // SM clocks per (warp, tap)
    printf("%-28s nt %3d: %.1f clk per task round, %.2f clk per warp-tap (DFMA floor %.2f), fp64 pipe %.0f %%\n", name, nt, clk,
           per_warp_iter, IR * IQ / 2.0, 100.0 * (IR * IQ / 2.0) / per_warp_iter);
}

int main()
{
    double *out, *gbank;
    cudaMalloc(&out, 148 * 512 * 8);
    cudaMalloc(&gbank, NG * SMAXP * 8 * 8);
    static double h[NG * SMAXP * 8];
    for (int i = 0; i < NG * SMAXP * 8; i++) h[i] = 1e-4 * (i % 97);
    cudaMemcpy(gbank, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(cbank, h, sizeof(h));
    for (int nt : {256, 512}) {
        run<0, 8, 3>("smem bank  8x3", out, gbank, nt);
        run<1, 8, 3>("const bank 8x3", out, gbank, nt);
        run<2, 8, 3>("ldg bank   8x3", out, gbank, nt);
        run<0, 8, 4>("smem bank  8x4", out, gbank, nt);
        run<1, 8, 4>("const bank 8x4", out, gbank, nt);
        run<0, 4, 6>("smem bank  4x6", out, gbank, nt);
        run<1, 4, 6>("const bank 4x6", out, gbank, nt);
        run<1, 6, 4>("const bank 6x4", out, gbank, nt);
        run<0, 6, 4>("smem bank  6x4", out, gbank, nt);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
