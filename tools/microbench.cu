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

// fp64 <-> int32 conversions and fp64 division (the timing arithmetic of the order-2 interpolator)
template <int WHAT, int ILP>  // 0: D2I+I2D round trip, 1: __ddiv_rn, 2: (double)int only, 3: __double2int_rz only
__global__ void k_cvt(double* out, int iters, double a)
{
    double acc[ILP];
    int ia[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { acc[i] = threadIdx.x * 1e-3 + i + 1.5; ia[i] = threadIdx.x + i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (WHAT == 0) acc[i] = (double) __double2int_rz(acc[i]) + a;
            else if (WHAT == 1) acc[i] = __ddiv_rn(acc[i], a);
            else if (WHAT == 2) { acc[i] += (double) ia[i]; ia[i] ^= it; }
            else { ia[i] += __double2int_rz(acc[i]); acc[i] += a; }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i] + ia[i];
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
    {
        const int nt = 512, nb = p.multiProcessorCount, it = 2048;
        auto rate = [&](float ms, int ops) { return (double) nt / 32 * 8 * it * ops / (ms * 1e-3) / 1.9e9; };
        float m0 = timeit([&] { k_cvt<0, 8><<<nb, nt>>>(out, it, 0.37); });
        float m1 = timeit([&] { k_cvt<1, 8><<<nb, nt>>>(out, it, 1.0000001); });
        float m2 = timeit([&] { k_cvt<2, 8><<<nb, nt>>>(out, it, 0.37); });
        float m3 = timeit([&] { k_cvt<3, 8><<<nb, nt>>>(out, it, 0.37); });
        printf("warp-instr/clk/SM: D2I+I2D+DADD triple %.3f | DDIV %.3f | I2D+DADD %.3f | D2I+DADD %.3f\n", rate(m0, 1), rate(m1, 1),
               rate(m2, 1), rate(m3, 1));
    }
    cudaFuncSetAttribute(k_lds<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_lds<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    // stride 0 = every lane reads the same address (broadcast); the last column is warp-wide load
    // instructions retired per clock per SM (1.0 = one wavefront per instruction)
    const int strides[] = {0, 1, 2, 4, 5, 8, 9, 16, 17};
    for (int w = 1; w <= 2; w++)
        for (int stride : strides) {
            const int nt = 512, nb = p.multiProcessorCount;
            float ms = (w == 1) ? timeit([&] { k_lds<1><<<nb, nt, 65536>>>(out, 4096, stride); })
                                : timeit([&] { k_lds<2><<<nb, nt, 65536>>>(out, 4096, stride); });
            double bytes = (double) nb * nt * 4096 * 8 * 8 * w;
            printf("LDS.%d stride %2d: %.1f B/clk/SM (at 1.9 GHz), %.3f warp-loads/clk/SM\n", 64 * w, stride,
                   bytes / (ms * 1e-3) / nb / 1.9e9, (double) (nt / 32) * 4096 * 8 / (ms * 1e-3) / 1.9e9);
        }
    return 0;
}
