// aus_device.cuh -- device-side building blocks: complex helpers, PTX wrappers (mbarrier, TMA bulk copy,
// named barriers), in-register radix-R DFTs and thread-group reductions/scans.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace aus {

// ------------------------------------------------------------------------------------------------ complex
template <typename T> struct alignas(2 * sizeof(T)) Cx { T x, y; };

__device__ __forceinline__ Cx<double> operator+(Cx<double> a, Cx<double> b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ Cx<double> operator-(Cx<double> a, Cx<double> b) { return {a.x - b.x, a.y - b.y}; }
// f32: Blackwell's packed f32x2 ALU ops (SASS FADD2 / FMUL2) do a whole complex add / sub / scale in one issue slot;
// the f32 kernel is issue-bound, and (re, im) already sits in an aligned register pair.
__device__ __forceinline__ unsigned long long cx_bits(Cx<float> a) { return *reinterpret_cast<unsigned long long *>(&a); }
__device__ __forceinline__ Cx<float> cx_from(unsigned long long u) { return *reinterpret_cast<Cx<float> *>(&u); }
__device__ __forceinline__ Cx<float> operator+(Cx<float> a, Cx<float> b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(cx_bits(a)), "l"(cx_bits(b)));
    return cx_from(r);
}
__device__ __forceinline__ Cx<float> operator-(Cx<float> a, Cx<float> b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(cx_bits(a)), "l"(cx_bits(b)));
    return cx_from(r);
}
// element-wise product (x*x', y*y'): the window multiply of the packed real input
__device__ __forceinline__ Cx<double> emul(Cx<double> a, Cx<double> b) { return {a.x * b.x, a.y * b.y}; }
__device__ __forceinline__ Cx<float> emul(Cx<float> a, Cx<float> b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(cx_bits(a)), "l"(cx_bits(b)));
    return cx_from(r);
}
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

// FP64 twiddles live in the constant bank so DFMA/DMUL read them as c[bank][offset] operands; a 64-bit literal
// would cost two MOVs at every use.  FP32 twiddles fit a 32-bit immediate.
static __constant__ double c_w32_cos[16] = {1.0,
                                     0.98078528040323044912618223613424, 0.92387953251128675612818318939679,
                                     0.83146961230254523707878837761791, 0.70710678118654752440084436210485,
                                     0.55557023301960222474283081394853, 0.38268343236508977172845998403040,
                                     0.19509032201612826784828486847702, 0.0,
                                     -0.19509032201612826784828486847702, -0.38268343236508977172845998403040,
                                     -0.55557023301960222474283081394853, -0.70710678118654752440084436210485,
                                     -0.83146961230254523707878837761791, -0.92387953251128675612818318939679,
                                     -0.98078528040323044912618223613424};
static __constant__ double c_w32_sin[16] = {0.0,
                                     0.19509032201612826784828486847702, 0.38268343236508977172845998403040,
                                     0.55557023301960222474283081394853, 0.70710678118654752440084436210485,
                                     0.83146961230254523707878837761791, 0.92387953251128675612818318939679,
                                     0.98078528040323044912618223613424, 1.0,
                                     0.98078528040323044912618223613424, 0.92387953251128675612818318939679,
                                     0.83146961230254523707878837761791, 0.70710678118654752440084436210485,
                                     0.55557023301960222474283081394853, 0.38268343236508977172845998403040,
                                     0.19509032201612826784828486847702};
// tan(2*pi*i/32) for the 6-operation butterflies (entry 8, the pole, is never used: W^8 = -i is handled as a swap)
#define AUS_W32_TAN \
    0.0, 0.19891236737965800691159762264467623, 0.41421356237309504880168872420969808, \
    0.66817863791929891999775768652308076, 1.0, 1.4966057626654890176011351349424769, \
    2.4142135623730950488016887242096981, 5.0273394921258481045149750710640724, 0.0, \
    -5.0273394921258481045149750710640724, -2.4142135623730950488016887242096981, -1.4966057626654890176011351349424769, \
    -1.0, -0.66817863791929891999775768652308076, -0.41421356237309504880168872420969808, \
    -0.19891236737965800691159762264467623
__host__ __device__ constexpr double w32_tan(int i) {
    constexpr double t[16] = {AUS_W32_TAN};
    return t[i];
}
static __constant__ double c_w32_tan[16] = {AUS_W32_TAN};
template <typename T, int IDX> __device__ __forceinline__ T tw32_tan() {
    if constexpr (sizeof(T) == 8) return c_w32_tan[IDX];
    else { constexpr T c = (T)w32_tan(IDX); return c; }
}
template <typename T, int IDX> __device__ __forceinline__ T tw32_cos() {
    if constexpr (sizeof(T) == 8) return c_w32_cos[IDX];
    else { constexpr T c = (T)w32_cos(IDX); return c; }
}
template <typename T, int IDX> __device__ __forceinline__ T tw32_sin() {
    if constexpr (sizeof(T) == 8) return c_w32_sin[IDX];
    else { constexpr T c = (T)w32_sin(IDX); return c; }
}

// Pass-2 input twiddles are the powers w^r, r = 1..31, of one per-thread root w = tab[stride].  The 31 table loads per
// frame (124 L1 wavefronts in f64, and a 16 KB table that does not fit next to the window in what the shared-memory
// carve-out leaves of L1) are replaced by four anchors w^1, w^8, w^16, w^24 and a running product in between: at most
// seven multiplications deep (~1e-15 relative in f64).  f32 keeps the table (it is issue-bound, not L1-bound).
#ifndef AUS_TW2_RECUR
#define AUS_TW2_RECUR(T) (sizeof(T) == 8)
#endif
template <typename T, typename F> __device__ __forceinline__ void for_pass2_twiddles(const Cx<T> *tab, int stride, F &&f) {
    if constexpr (AUS_TW2_RECUR(T)) {
        const Cx<T> w1 = ldg_cx(tab + stride);
        Cx<T> a8 = ldg_cx(tab + 8 * stride), a16 = ldg_cx(tab + 16 * stride), a24 = ldg_cx(tab + 24 * stride);
        Cx<T> a0 = w1;
        f(std::integral_constant<int, 1>{}, a0);
        f(std::integral_constant<int, 8>{}, a8);
        f(std::integral_constant<int, 16>{}, a16);
        f(std::integral_constant<int, 24>{}, a24);
        static_for<1, 8>([&](auto i_) {
            constexpr int i = decltype(i_)::value;
            if constexpr (i < 7) { a0 = cmul(a0, w1); f(std::integral_constant<int, 1 + i>{}, a0); }
            a8 = cmul(a8, w1);
            f(std::integral_constant<int, 8 + i>{}, a8);
            a16 = cmul(a16, w1);
            f(std::integral_constant<int, 16 + i>{}, a16);
            a24 = cmul(a24, w1);
            f(std::integral_constant<int, 24 + i>{}, a24);
        });
    } else {
        static_for<1, 32>([&](auto r_) { f(r_, ldg_cx(tab + decltype(r_)::value * stride)); });
    }
}

// ------------------------------------------------------------------------------------------------ radix-R DFT
// Forward DFT (sign -1) of v[0..R) held in registers, fully unrolled.  Natural-order input; the result is left in
// bit-reversed order: natural output q is v[bitrev(q, log2 R)].  All indices are compile-time, so permutations cost
// nothing.
//
// The network is radix-2 decimation in TIME run on the bit-reversed view w[i] = v[bitrev(i)] (which makes the input
// natural in v and the output bit-reversed in v -- the same interface as a decimation-in-frequency network).  DIT puts
// the twiddle on ONE input of a butterfly, which allows the 6-operation form
//     g = b (1 - i tan)   (2 FMA),      a +- cos * g   (4 FMA)
// instead of the 8 operations of  (a - b) * (cos - i sin)  after the subtraction: 34 of the 80 butterflies of a radix-32
// pass carry a non-trivial twiddle, so a pass costs 388 instead of 456 floating-point operations.
// f32: the four FMAs of  a +- cos * g  are two packed fma.rn.f32x2.
__device__ __forceinline__ Cx<double> cx_fma(double s, Cx<double> g, Cx<double> a) { return {fma(s, g.x, a.x), fma(s, g.y, a.y)}; }
__device__ __forceinline__ Cx<float> cx_fma(float s, Cx<float> g, Cx<float> a) {
    const Cx<float> ss{s, s};
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(cx_bits(ss)), "l"(cx_bits(g)), "l"(cx_bits(a)));
    return cx_from(r);
}
template <int R, typename T> __device__ __forceinline__ void dft_dif(Cx<T> (&v)[R]) {
    static_assert(R >= 1 && R <= 32 && (R & (R - 1)) == 0, "radix must be a power of two <= 32");
    constexpr int LG = ilog2(R);
    static_for<0, LG>([&](auto s_) {
        constexpr int half = 1 << decltype(s_)::value;            // butterflies span `half` positions of the view
        static_for<0, R / (2 * half)>([&](auto b_) {
            constexpr int blk = decltype(b_)::value * 2 * half;
            static_for<0, half>([&](auto k_) {
                constexpr int k = decltype(k_)::value;
                constexpr int idx = k * (16 / half);              // twiddle W_32^idx = W_{2*half}^k
                constexpr int ia = bitrev(blk + k, LG), ib = bitrev(blk + k + half, LG);
                const Cx<T> a = v[ia], b = v[ib];
                if constexpr (idx == 0) {
                    v[ia] = a + b;
                    v[ib] = a - b;
                } else if constexpr (idx == 8) {                  // W = -i:  W b = (b.y, -b.x)
                    const Cx<T> g{b.y, -b.x};
                    v[ia] = a + g;
                    v[ib] = a - g;
                } else {
                    const T tn = tw32_tan<T, idx>(), c = tw32_cos<T, idx>();
                    const Cx<T> g{fma(tn, b.y, b.x), fma(-tn, b.x, b.y)};       // b (1 - i tan) = W b / cos
                    v[ia] = cx_fma(c, g, a);
                    v[ib] = cx_fma(-c, g, a);
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
// L2 prefetch of a byte range (16-byte aligned start, size multiple of 16)
__device__ __forceinline__ void prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
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

// The same collectives in the arithmetic type of the kernel (R = float in f32 mode: half the shuffles of the double
// versions and no conversions; f32 results are compared at 1e-4) -- used by the transposed-layout feature pass.
template <int TPF, int BLOCK, int NV, typename R>
__device__ __forceinline__ void group_allreduce_t(R (&v)[NV], double *scratch_d, int t, int grp) {
    constexpr int W = TPF < 32 ? TPF : 32;
#pragma unroll
    for (int off = W / 2; off >= 1; off >>= 1) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
    }
    if constexpr (TPF > 32) {
        R *scratch = reinterpret_cast<R *>(scratch_d);
        constexpr int NW = TPF / 32;
        const int w = t >> 5;
        if ((t & 31) == 0) {
#pragma unroll
            for (int i = 0; i < NV; ++i) scratch[w * NV + i] = v[i];
        }
        group_sync<TPF, BLOCK>(grp);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            R s = scratch[i];
            for (int ww = 1; ww < NW; ++ww) s += scratch[ww * NV + i];
            v[i] = s;
        }
        group_sync<TPF, BLOCK>(grp);
    }
}
template <int TPF, int BLOCK, typename R>
__device__ __forceinline__ void group_inclusive_scan1(R &a, double *scratch_d, int t, int grp) {
    constexpr int W = TPF < 32 ? TPF : 32;
    const int l = t & (W - 1);
#pragma unroll
    for (int off = 1; off < W; off <<= 1) {
        const R ua = __shfl_up_sync(0xffffffffu, a, off, W);
        if (l >= off) a += ua;
    }
    if constexpr (TPF > 32) {
        R *scratch = reinterpret_cast<R *>(scratch_d);
        constexpr int NW = TPF / 32;
        const int w = t >> 5;
        if ((t & 31) == 31) scratch[w] = a;
        group_sync<TPF, BLOCK>(grp);
        R pa = (R)0;
        for (int ww = 0; ww < NW; ++ww)
            if (ww < w) pa += scratch[ww];
        a += pa;
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
// The f64 feature pass needs one sqrt and one log2 per bin at ~1e-12 relative accuracy (tolerance 1e-9), not the
// ~45 FP64 instructions each of the correctly rounded library versions.
//   sqrt:  y0 = rsqrt.approx (MUFU.RSQ64H, ~2^-20), one Newton step -> ~1e-12; 4 FP64 ops.
//   log2:  x = 2^e z, z in [1,2); 128-entry table of (1/c_i, -log2(1/c_i)) on the top 7 mantissa bits,
//          r = z/c_i - 1 (|r| < 2^-8), degree-5 Taylor of log2(1+r) (|err| < 1e-15); 7 FP64 ops.
// Zero, subnormal, negative, inf and NaN arguments take the library function (out of line, rare).
constexpr int LOG2_TAB = 128;

// true for +0, subnormals, negatives, inf and NaN: the arguments the fast paths below do not handle
__device__ __forceinline__ bool abnormal(double x) { return (unsigned)__double2hiint(x) - 0x00100000u >= 0x7fe00000u; }
__device__ __forceinline__ bool abnormal(float) { return false; }

template <typename T> struct Mth;
// log2 polynomial coefficients and the exponent-conversion constant as constant-bank operands: as 64-bit literals each
// costs two MOVs per use, which the compiler re-materialises inside the rolled per-bin loops (5 of 35 instructions per bin)
static __constant__ double c_log2_k[6] = {0.28853900817779268147,      // 1/(5 ln2)
                                   -0.36067376022224085184,     // -1/(4 ln2)
                                   0.48089834696298780245,      // 1/(3 ln2)
                                   -0.72134752044448170368,     // -1/(2 ln2)
                                   1.44269504088896340736,      // 1/ln2
                                   4503601774854144.0};         // 2^52 + 2^31
template <> struct Mth<double> {
    static __device__ __forceinline__ double sqrt_exact(double x) { return sqrt(x); }
    static __device__ __forceinline__ double log2_exact(double x) { return log2(x); }
    // positive normal x only (see abnormal())
    static __device__ __forceinline__ double sqrt_(double x) {
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
        const double h = x * y;
        const double r = fma(-h, y, 1.0);
        return fma(h * 0.5, r, h);
    }
    static __device__ __forceinline__ double log2_(double x, const double2 *tab) {
        const unsigned hi = (unsigned)__double2hiint(x);
        const double2 tc = tab[((hi >> 13) & (LOG2_TAB - 1)) * 8];     // 8 replicas per entry: see Cfg::LOGREP
        const double z = __hiloint2double((int)((hi & 0x000fffffu) | 0x3ff00000u), __double2loint(x));
        // exponent as a double without an I2F: 2^52 + 2^31 + (e ^ 0x80000000) trick
        const double ed = __hiloint2double(0x43300000, (int)(((hi >> 20) - 1023u) ^ 0x80000000u)) - c_log2_k[5];
        const double r = fma(z, tc.x, -1.0);
        double pl = fma(r, c_log2_k[0], c_log2_k[1]);
        pl = fma(pl, r, c_log2_k[2]);
        pl = fma(pl, r, c_log2_k[3]);
        pl = fma(pl, r, c_log2_k[4]);
        return fma(pl, r, ed + tc.y);
    }
};
template <> struct Mth<float> {
    static __device__ __forceinline__ float sqrt_exact(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float log2_exact(float x) { return log2f(x); }
    // opt-in f32 mode (tolerance 1e-4): single-MUFU approximations.  sqrt.approx.ftz(0) = 0 and lg2.approx.ftz(0) = -inf,
    // so zero bins keep the reference's NaN / -inf behaviour; max errors ~2^-22 (relative / absolute).
    static __device__ __forceinline__ float sqrt_(float x) {
        float r;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
    static __device__ __forceinline__ float log2_(float x, const double2 *) {
        float r;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
};

// any-of over the TPF threads of a frame group (warp vote; cross-warp through shared scratch)
template <int TPF, int BLOCK>
__device__ __forceinline__ bool group_any(bool pred, double *scratch, int t, int grp) {
    if constexpr (TPF >= 32) {
        bool w = __any_sync(0xffffffffu, pred);
        if constexpr (TPF > 32) {
            int *flag = reinterpret_cast<int *>(scratch);
            if (t == 0) *flag = 0;
            group_sync<TPF, BLOCK>(grp);
            if (w && (t & 31) == 0) *flag = 1;
            group_sync<TPF, BLOCK>(grp);
            w = *flag != 0;
            group_sync<TPF, BLOCK>(grp);
        }
        return w;
    } else {
        const unsigned b = __ballot_sync(0xffffffffu, pred);
        const unsigned lane = threadIdx.x & 31u;
        const unsigned m = ((TPF == 32 ? 0xffffffffu : ((1u << TPF) - 1u))) << (lane & ~(unsigned)(TPF - 1));
        return (b & m) != 0;
    }
}

// Order-preserving integer key of a non-negative power value (NaN sorts above everything): f64 -> high word,
// f32 -> bit pattern >> 1 (so it fits the same unsigned compare); key >= *_MIN_NORMAL <=> positive normal or larger.
__device__ __forceinline__ unsigned power_hi(double x) { return (unsigned)__double2hiint(x); }
__device__ __forceinline__ unsigned power_hi(float x) { return __float_as_uint(x) >> 1; }
template <typename T> __device__ __forceinline__ bool power_is_normal(unsigned key) {
    if constexpr (sizeof(T) == 8) return key - 0x00100000u < 0x7fe00000u;          // [2^-1022, inf)
    else return key - (0x00800000u >> 1) < ((0x7f800000u - 0x00800000u) >> 1);     // [2^-126, inf)
}
// a value with that key (low bits zero): a lower bound of the smallest power, good enough to test q = p / P
template <typename T> __device__ __forceinline__ T power_from_hi(unsigned key) {
    if constexpr (sizeof(T) == 8) return __hiloint2double((int)key, 0);
    else return __uint_as_float(key << 1);
}
// min over the TPF threads of a frame group
template <int TPF, int BLOCK>
__device__ __forceinline__ unsigned group_min_u32(unsigned v, double *scratch, int t, int grp) {
    constexpr int W = TPF < 32 ? TPF : 32;
#pragma unroll
    for (int off = W / 2; off >= 1; off >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, off));
    if constexpr (TPF > 32) {
        unsigned *s = reinterpret_cast<unsigned *>(scratch);
        if ((t & 31) == 0) s[t >> 5] = v;
        group_sync<TPF, BLOCK>(grp);
        for (int w = 0; w < TPF / 32; ++w) v = min(v, s[w]);
        group_sync<TPF, BLOCK>(grp);
    }
    return v;
}

}  // namespace aus
