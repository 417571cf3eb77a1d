// Micro-benchmark: how close to the FP64 pipe peak does the in-register radix-32 DFT of the forward kernel get on its own
// (no memory traffic), at the occupancies the kernels run at?  Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a
//   -I aus-rust_b200/csrc -o tools/ubench/fp64_butterfly tools/ubench/fp64_butterfly.cu
#include "aus_device.cuh"
#include <cstdio>
#include <cstdlib>
using namespace aus;

template <int MAXREG>
__global__ void __maxnreg__(MAXREG) bfly(double2 *io, int iters) {
    Cx<double> v[32];
    const size_t base = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 32;
#pragma unroll
    for (int r = 0; r < 32; ++r) { const double2 t = io[base + r]; v[r] = Cx<double>{t.x, t.y}; }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        dft_dif<32>(v);
#pragma unroll
        for (int r = 0; r < 32; ++r) { v[r].x *= 0.17677669529663687; v[r].y *= 0.17677669529663687; }   // keep it bounded (64 DMUL)
    }
#pragma unroll
    for (int r = 0; r < 32; ++r) io[base + r] = make_double2(v[r].x, v[r].y);
}

// dependent-chain probe: ILP independent DFMA chains per thread
template <int ILP> __global__ void chain(double *io, int iters) {
    double a[ILP];
    for (int i = 0; i < ILP; ++i) a[i] = io[threadIdx.x + i];
    const double b = io[0], c = io[1];
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += a[i];
    io[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double peak_wi = sms * 4.0 * clk * 1e3 / 2.0;      // FP64 warp-instructions / s
    double2 *d; cudaMalloc(&d, (size_t)sms * 512 * 32 * sizeof(double2)); cudaMemset(d, 0, (size_t)sms * 512 * 32 * sizeof(double2));
    const int iters = 2000;
    const double wi_per_iter = 388 + 64;                     // per thread = per warp-instruction count
    for (int warps : {4, 8, 12}) {
        float ms = warps == 12 ? timeit([&] { bfly<168><<<sms, 384>>>(d, iters); })
                 : warps == 8  ? timeit([&] { bfly<255><<<sms, 256>>>(d, iters); })
                               : timeit([&] { bfly<255><<<sms, 128>>>(d, iters); });
        const double wi = (double)sms * warps * iters * wi_per_iter;
        printf("radix-32 butterfly, %2d warps/SM: %.3f ms, %.1f %% of FP64 issue peak (%.0f cycles per butterfly per SM-scheduler warp)\n",
               warps, ms, 100.0 * wi / (ms * 1e-3) / peak_wi, ms * 1e-3 * clk * 1e3 / iters);
    }
    double *dd = (double *)d;
    for (int warps : {4, 12}) {
        float m1 = timeit([&] { chain<1><<<sms, warps * 32>>>(dd, 20000); });
        float m2 = timeit([&] { chain<2><<<sms, warps * 32>>>(dd, 20000); });
        float m4 = timeit([&] { chain<4><<<sms, warps * 32>>>(dd, 20000); });
        float m8 = timeit([&] { chain<8><<<sms, warps * 32>>>(dd, 20000); });
        printf("DFMA chains, %2d warps/SM: cycles per DFMA per warp: ILP1 %.2f  ILP2 %.2f  ILP4 %.2f  ILP8 %.2f\n", warps,
               m1 * 1e-3 * clk * 1e3 / 20000, m2 * 1e-3 * clk * 1e3 / 40000, m4 * 1e-3 * clk * 1e3 / 80000, m8 * 1e-3 * clk * 1e3 / 160000);
    }
    printf("clock %d kHz, %d SMs\n", clk, sms);
    return 0;
}
