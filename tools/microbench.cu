// tools/microbench.cu -- B200 fp64 pipe / shared-memory calibration used to size the kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b)
{
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed DADD/DMUL (non-fused) stream
template <int ILP>
__global__ void k_daddmul(double* out, int iters, double a, double b)
{
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) acc[i] = __dadd_rn(__dmul_rn(acc[i], a), b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int W>  // W = 1: LDS.64 ; W = 2: LDS.128
__global__ void k_lds(double* out, int iters, int stride)
{
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
    __syncthreads();
    double s = 0;
    int idx = (threadIdx.x * stride * W) & 8191;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (W == 1) s += sm[(idx + u * 32 * W) & 8191];
            else {
                double2 v = *reinterpret_cast<double2*>(&sm[(idx + u * 64) & 8190]);
                s += v.x + v.y;
            }
        }
        idx = (idx + 1024) & 8191;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
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
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    double* out;
    cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double));
    const int iters = 4096;
    for (int warps = 1; warps <= 32; warps *= 2) {
        const int nt = warps * 32, nb = p.multiProcessorCount;
        float m1 = timeit([&] { k_dfma<1><<<nb, nt>>>(out, iters, 1.0000001, 1e-9); });
        float m4 = timeit([&] { k_dfma<4><<<nb, nt>>>(out, iters, 1.0000001, 1e-9); });
        float m16 = timeit([&] { k_dfma<16><<<nb, nt>>>(out, iters, 1.0000001, 1e-9); });
        float a16 = timeit([&] { k_daddmul<16><<<nb, nt>>>(out, iters, 1.0000001, 1e-9); });
        auto tf = [&](float ms, int ilp, int flop) { return (double) nb * nt * ilp * iters * flop / (ms * 1e-3) / 1e12; };
        printf("warps/SM %2d: DFMA ILP1 %.2f TF  ILP4 %.2f TF  ILP16 %.2f TF | DMUL+DADD ILP16 %.2f TF(2 instr)\n", warps,
               tf(m1, 1, 2), tf(m4, 4, 2), tf(m16, 16, 2), tf(a16, 16, 2));
    }
    // dependent-chain latency: 1 warp, ILP1
    {
        float ms = timeit([&] { k_dfma<1><<<1, 32>>>(out, 1 << 20, 1.0000001, 1e-9); });
        printf("DFMA dependent latency ~ %.2f ns per op (x clock GHz = cycles)\n", ms * 1e6 / (1 << 20));
    }
    cudaFuncSetAttribute(k_lds<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_lds<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int w = 1; w <= 2; w++)
        for (int stride = 1; stride <= 17; stride += (stride == 1 ? 15 : 1)) {
            const int nt = 256, nb = p.multiProcessorCount;
            float ms = (w == 1) ? timeit([&] { k_lds<1><<<nb, nt, 65536>>>(out, 4096, stride); })
                                : timeit([&] { k_lds<2><<<nb, nt, 65536>>>(out, 4096, stride); });
            double bytes = (double) nb * nt * 4096 * 8 * 8 * w;
            printf("LDS.%d stride %2d: %.1f GB/s per SM-clk -> %.1f B/clk/SM (at 1.9 GHz)\n", 64 * w, stride, bytes / (ms * 1e-3) / 1e9,
                   bytes / (ms * 1e-3) / nb / 1.9e9);
        }
    return 0;
}
