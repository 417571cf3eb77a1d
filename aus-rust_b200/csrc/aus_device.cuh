// aus_device.cuh -- device-side building blocks: complex helpers, PTX wrappers (mbarrier, TMA bulk copy,
// named barriers), in-register radix-R DFTs and thread-group reductions/scans.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace aus {

// ------------------------------------------------------------------------------------------------ complex
template <typename T> struct alignas(2 * sizeof(T)) Cx { T x, y; };

template <typename T> __device__ __forceinline__ Cx<T> operator+(Cx<T> a, Cx<T> b) { return {a.x + b.x, a.y + b.y}; }
template <typename T> __device__ __forceinline__ Cx<T> operator-(Cx<T> a, Cx<T> b) { return {a.x - b.x, a.y - b.y}; }
template <typename T> __device__ __forceinline__ Cx<T> cmul(Cx<T> a, Cx<T> b) {
    return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}

// read-only (non-coherent) table loads
__device__ __forceinline__ Cx<double> ldg_cx(const Cx<double> *p) {
    const double2 v = __ldg(reinterpret_cast<const double2 *>(p));
    return Cx<double>{v.x, v.y};
}
__device__ __forceinline__ Cx<float> ldg_cx(const Cx<float> *p) {
    const float2 v = __ldg(reinterpret_cast<const float2 *>(p));
    return Cx<float>{v.x, v.y};
}

// ------------------------------------------------------------------------------------------------ compile-time loop
template <int I, int N, typename F> __device__ __forceinline__ void static_for(F &&f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

__host__ __device__ constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v >> 1); }
__host__ __device__ constexpr int bitrev(int v, int bits) {
    int r = 0;
    for (int b = 0; b < bits; ++b) r |= ((v >> b) & 1) << (bits - 1 - b);
    return r;
}

// cos/sin(2*pi*i/32), i = 0..15, correctly rounded doubles
__host__ __device__ constexpr double w32_cos(int i) {
    constexpr double t[16] = {1.0,
                              0.98078528040323044912618223613424, 0.92387953251128675612818318939679,
                              0.83146961230254523707878837761791, 0.70710678118654752440084436210485,
                              0.55557023301960222474283081394853, 0.38268343236508977172845998403040,
                              0.19509032201612826784828486847702, 0.0,
                              -0.19509032201612826784828486847702, -0.38268343236508977172845998403040,
                              -0.55557023301960222474283081394853, -0.70710678118654752440084436210485,
                              -0.83146961230254523707878837761791, -0.92387953251128675612818318939679,
                              -0.98078528040323044912618223613424};
    return t[i];
}
__host__ __device__ constexpr double w32_sin(int i) {
    constexpr double t[16] = {0.0,
                              0.19509032201612826784828486847702, 0.38268343236508977172845998403040,
                              0.55557023301960222474283081394853, 0.70710678118654752440084436210485,
                              0.83146961230254523707878837761791, 0.92387953251128675612818318939679,
                              0.98078528040323044912618223613424, 1.0,
                              0.98078528040323044912618223613424, 0.92387953251128675612818318939679,
                              0.83146961230254523707878837761791, 0.70710678118654752440084436210485,
                              0.55557023301960222474283081394853, 0.38268343236508977172845998403040,
                              0.19509032201612826784828486847702};
    return t[i];
}

// ------------------------------------------------------------------------------------------------ radix-R DFT
// Forward DFT (sign -1) of v[0..R) held in registers: decimation-in-frequency radix-2 network, fully unrolled.
// The result is left in bit-reversed order: natural output q is v[bitrev(q, log2 R)].  All indices are
// compile-time, so the permutation costs nothing.
template <int R, typename T> __device__ __forceinline__ void dft_dif(Cx<T> (&v)[R]) {
    static_assert(R >= 1 && R <= 32 && (R & (R - 1)) == 0, "radix must be a power of two <= 32");
    constexpr int LG = ilog2(R);
    static_for<0, LG>([&](auto s_) {
        constexpr int s = decltype(s_)::value;
        constexpr int half = R >> (s + 1);
        static_for<0, R / (2 * half)>([&](auto b_) {
            constexpr int blk = decltype(b_)::value * 2 * half;
            static_for<0, half>([&](auto k_) {
                constexpr int k = decltype(k_)::value;
                constexpr int idx = k * (16 / half);          // twiddle W_32^idx = W_{2*half}^k
                const Cx<T> a = v[blk + k], b = v[blk + k + half];
                v[blk + k] = a + b;
                const Cx<T> d = a - b;
                if constexpr (idx == 0) {
                    v[blk + k + half] = d;
                } else if constexpr (idx == 8) {              // multiply by -i
                    v[blk + k + half] = Cx<T>{d.y, -d.x};
                } else {
                    constexpr T c = (T)w32_cos(idx), sn = (T)w32_sin(idx);
                    v[blk + k + half] = Cx<T>{d.x * c + d.y * sn, d.y * c - d.x * sn};   // d * (c - i s)
                }
            });
        });
    });
}

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// order prior generic-proxy accesses to shared memory before subsequent async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
// 1-D TMA bulk copy global -> shared (SASS: UBLKCP), completion bytes signalled on `bar`.
// dst, src 16-byte aligned; bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// 1-D TMA bulk copy shared -> global (bulk async-group completion)
__device__ __forceinline__ void tma_store_1d(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// Synchronise the TPF threads that cooperate on one frame ("frame group").  grp = group index inside the CTA.
template <int TPF, int BLOCK> __device__ __forceinline__ void group_sync(int grp) {
    if constexpr (TPF <= 32) {
        __syncwarp();
    } else if constexpr (TPF == BLOCK) {
        __syncthreads();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(TPF) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ group collectives
// All-reduce (sum) of NV doubles over the TPF threads of a frame group.  Fixed combination order -> deterministic.
// scratch: shared, NV * (TPF/32) doubles per group (only touched when TPF > 32).
template <int TPF, int BLOCK, int NV>
__device__ __forceinline__ void group_allreduce(double (&v)[NV], double *scratch, int t, int grp) {
    constexpr int W = TPF < 32 ? TPF : 32;
#pragma unroll
    for (int off = W / 2; off >= 1; off >>= 1) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
    }
    if constexpr (TPF > 32) {
        constexpr int NW = TPF / 32;
        const int w = t >> 5;
        if ((t & 31) == 0) {
#pragma unroll
            for (int i = 0; i < NV; ++i) scratch[w * NV + i] = v[i];
        }
        group_sync<TPF, BLOCK>(grp);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double s = scratch[i];
            for (int ww = 1; ww < NW; ++ww) s += scratch[ww * NV + i];
            v[i] = s;
        }
        group_sync<TPF, BLOCK>(grp);
    }
}

// Inclusive prefix sum over the group's threads (ascending t) of two doubles at once.
template <int TPF, int BLOCK>
__device__ __forceinline__ void group_inclusive_scan2(double &a, double &b, double *scratch, int t, int grp) {
    constexpr int W = TPF < 32 ? TPF : 32;
    const int l = t & (W - 1);
#pragma unroll
    for (int off = 1; off < W; off <<= 1) {
        const double ua = __shfl_up_sync(0xffffffffu, a, off, W);
        const double ub = __shfl_up_sync(0xffffffffu, b, off, W);
        if (l >= off) { a += ua; b += ub; }
    }
    if constexpr (TPF > 32) {
        constexpr int NW = TPF / 32;
        const int w = t >> 5;
        if ((t & 31) == 31) { scratch[2 * w] = a; scratch[2 * w + 1] = b; }
        group_sync<TPF, BLOCK>(grp);
        double pa = 0.0, pb = 0.0;
        for (int ww = 0; ww < NW; ++ww)
            if (ww < w) { pa += scratch[2 * ww]; pb += scratch[2 * ww + 1]; }
        a += pa; b += pb;
        group_sync<TPF, BLOCK>(grp);
    }
}

// ------------------------------------------------------------------------------------------------ math traits
template <typename T> struct Mth;
template <> struct Mth<double> {
    static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
    static __device__ __forceinline__ double log2_(double x) { return log2(x); }
};
template <> struct Mth<float> {
    static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float log2_(float x) { return log2f(x); }
};

}  // namespace aus
