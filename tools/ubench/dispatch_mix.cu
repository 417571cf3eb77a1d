// Micro-benchmark: which instruction classes overlap with FP64 issue on B200 (sm_100a)?
// Every warp runs, per loop iteration, 8 independent DFMAs (16 cycles of its scheduler's FP64 pipe) plus NX instructions of
// another class, all independent of each other (inline PTX, volatile).  12 warps per SM = 3 per scheduler, so without any
// interference an iteration costs 3 x 16 = 48 cycles per scheduler (+ loop overhead); anything above that is interference.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/dispatch_mix tools/ubench/dispatch_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int NX, int NF> __global__ void __launch_bounds__(384) mix(double *io, int iters) {
    extern __shared__ double2 sm[];
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = io[threadIdx.x + i];
    const double b = io[0] * 1e-9 + 1.0, c = io[1];
    int x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
    const int y = (int)io[2] + 3, z = (int)io[3] + 1;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_double2(i, 1);
    __syncthreads();
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16 + (threadIdx.x >> 5) * 512;
    double l0, l1;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NF; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i & 7]) : "d"(b), "d"(c));
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            if (MODE == 1) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[i & 7]) : "r"(y), "r"(z));
            if (MODE == 2) asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[i & 7]) : "r"(y));
            if (MODE == 3) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(l0), "=d"(l1) : "r"(saddr + 8192 * (i & 3)));
            if (MODE == 4) asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(saddr + 8192 * (i & 3)), "d"(b), "d"(c));
            if (MODE == 5) asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(x[i & 7]));
            if (MODE == 6) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a[(i + NF) & 7]) : "d"(c));
            if (MODE == 7) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float *)&x[i & 7]) : "f"(1.0001f), "f"(0.5f));
            if (MODE == 8) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(l0) : "r"(saddr / 2 + 4096 * (i & 3)));
        }
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += a[i] + x[i];
    io[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int NX, int NF = 8> double run(double *d, int sms, int iters, int clk) {
    cudaFuncSetAttribute(mix<MODE, NX, NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 16);
    mix<MODE, NX, NF><<<sms, 384, 4096 * 16>>>(d, iters);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    mix<MODE, NX, NF><<<sms, 384, 4096 * 16>>>(d, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1e-3 * clk * 1e3 / iters;      // cycles per iteration (= per scheduler: its 3 warps each run one iteration)
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double *d; cudaMalloc(&d, (size_t)p.multiProcessorCount * 384 * 64); cudaMemset(d, 0, (size_t)p.multiProcessorCount * 384 * 64);
    const int iters = 20000, sms = p.multiProcessorCount;
    printf("cycles per iteration per scheduler (3 warps x [8 DFMA + NX other]); 8 DFMA alone = 48 + loop overhead\n");
    printf("8 DFMA only            : %6.1f      0 DFMA, loop only: %6.1f\n", run<0, 0>(d, sms, iters, clk), run<0, 0, 0>(d, sms, iters, clk));
    printf("+ IMAD  x4 / x8 / x16  : %6.1f %6.1f %6.1f     IMAD x16 alone: %6.1f\n", run<1, 4>(d, sms, iters, clk), run<1, 8>(d, sms, iters, clk), run<1, 16>(d, sms, iters, clk), run<1, 16, 0>(d, sms, iters, clk));
    printf("+ LOP   x4 / x8 / x16  : %6.1f %6.1f %6.1f     LOP  x16 alone: %6.1f\n", run<2, 4>(d, sms, iters, clk), run<2, 8>(d, sms, iters, clk), run<2, 16>(d, sms, iters, clk), run<2, 16, 0>(d, sms, iters, clk));
    printf("+ FFMA  x4 / x8 / x16  : %6.1f %6.1f %6.1f     FFMA x16 alone: %6.1f\n", run<7, 4>(d, sms, iters, clk), run<7, 8>(d, sms, iters, clk), run<7, 16>(d, sms, iters, clk), run<7, 16, 0>(d, sms, iters, clk));
    printf("+ DADD  x4 / x8        : %6.1f %6.1f\n", run<6, 4>(d, sms, iters, clk), run<6, 8>(d, sms, iters, clk));
    printf("+ LDS.128 x2 / x4 / x8 : %6.1f %6.1f %6.1f     LDS.128 x8 alone: %6.1f\n", run<3, 2>(d, sms, iters, clk), run<3, 4>(d, sms, iters, clk), run<3, 8>(d, sms, iters, clk), run<3, 8, 0>(d, sms, iters, clk));
    printf("+ LDS.64  x2 / x4 / x8 : %6.1f %6.1f %6.1f     LDS.64  x8 alone: %6.1f\n", run<8, 2>(d, sms, iters, clk), run<8, 4>(d, sms, iters, clk), run<8, 8>(d, sms, iters, clk), run<8, 8, 0>(d, sms, iters, clk));
    printf("+ STS.128 x2 / x4 / x8 : %6.1f %6.1f %6.1f     STS.128 x8 alone: %6.1f\n", run<4, 2>(d, sms, iters, clk), run<4, 4>(d, sms, iters, clk), run<4, 8>(d, sms, iters, clk), run<4, 8, 0>(d, sms, iters, clk));
    printf("+ SHFL.32 x2 / x4 / x8 : %6.1f %6.1f %6.1f     SHFL x8 alone:    %6.1f\n", run<5, 2>(d, sms, iters, clk), run<5, 4>(d, sms, iters, clk), run<5, 8>(d, sms, iters, clk), run<5, 8, 0>(d, sms, iters, clk));
    return 0;
}
