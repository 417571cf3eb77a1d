// aus_stft_kernel.cuh -- fused frame-gather + window + packed real FFT + |X|^2 + 14 spectral features (K1),
// optional spectra output (K2) and features from caller-supplied magnitudes (K3).  sm_100a.
//
// Replaces, per frame: spectrum/fft.rs:128-176 (framing, window, FFT), spectrum/fft_tools.rs:41-55 (magnitude),
// analysis/analyzer.rs:41-81 + analysis/spectral_analysis_tools.rs:79-290 (features).
//
// Work decomposition ("frame group"): TPF = N/64 threads cooperate on one frame; every thread always holds
// P = 32 complex points in registers.  The N-point real transform is an M = N/2 point complex Stockham FFT of
// z[n] = x[2n] + i x[2n+1] with radices {TPF, 32} (M <= 1024) or {32, 32, M/1024} (M > 1024), followed by the
// real-to-complex post-twiddle fused into the feature pass.  One shared-memory buffer of M(+pad) complex per
// group is the TMA landing zone for the frame's samples, then the exchange space between passes, then the home
// of the power spectrum and its running sum during the feature passes; spectra never touch HBM unless the
// caller asks for them.
//
// Code-size discipline matters as much as FLOPs here: the kernel is instruction-issue bound, so the two radix-32
// passes share ONE butterfly body (a rolled 2-trip loop with linear address generators) and the per-bin feature
// loops are rolled, keeping the hot loop inside the instruction cache.
#pragma once
#include "aus_device.cuh"

#ifndef AUS_F64_BLOCK_R
#define AUS_F64_BLOCK_R 256      // f64 CTA size of the register-resident path; registers are allocated per 4 warps
#endif
#ifndef AUS_UNROLL_PASSES
// register-resident kernels: two copies of the radix-32 pass body with compile-time strides instead of one rolled loop
// (saves the per-load address arithmetic).  Measured at fft 2048: f32 +7 % (features) / +9 % (spectra), f64 spectra-only
// +10 %, f64 both outputs +8 %; the 12-warp shared-memory-tail kernel keeps the rolled body (-1.4 % unrolled).
#define AUS_UNROLL_PASSES 1
#endif
#ifndef AUS_PASSB_W
#define AUS_PASSB_W 4            // bins per block of the second feature pass (shared-memory tail)
#endif
#ifndef AUS_F32_MINB
#define AUS_F32_MINB 2              // f32 CTAs (256 threads) per SM
#endif
#ifndef AUS_WINDOW_PRELOAD
#define AUS_WINDOW_PRELOAD 1     // fetch the pass-1 window values before waiting for the frame's bulk copy
#endif
#ifndef AUS_L2_PREFETCH
#define AUS_L2_PREFETCH 0        // experiment: cp.async.bulk.prefetch.L2 of the next frame / row while this one is processed.
                                 // Measured on 2 h inputs: forward features -1..-6 %, spectra -4..-6 %, analyzer +1 %, irstft +-4 %
                                 // (the neighbouring frames other warps are loading already keep L2 warm) -> off.
#endif

namespace aus {

// RR: register-resident tail (aus_stft_r.cuh).  It wants ~230 registers in f64, so those CTAs are
// 8 warps x 255 registers instead of 12 x 168; that also shrinks the shared-memory carve-out enough for the window table
// to stay in L1 (hit rate 68 % -> 99 %).
template <typename T, int LOG2N, bool RR = false> struct Cfg {
    static constexpr int N = 1 << LOG2N;
    static constexpr int M = N / 2;
    static constexpr int K = M + 1;
    static constexpr int TPF = M / 32;                       // threads per frame, 1..256
    static constexpr bool THREE = (M > 1024);
    static constexpr int R1 = THREE ? 32 : TPF;              // first radix (TPF == 1: single radix-32 pass)
    static constexpr int R3 = THREE ? M / 1024 : 1;
    // f64: 12 warps x 168 registers per SM (3 per scheduler) when the group buffers fit; f32: 2 CTAs x 8 warps x 128 regs
    static constexpr int BLOCK = (sizeof(T) == 8 && RR) ? AUS_F64_BLOCK_R
                                 : (sizeof(T) == 8 && TPF >= 16 && TPF <= 128) ? 384 : 256;
    static constexpr int MINB = (sizeof(T) == 4) ? AUS_F32_MINB : 1;
    static constexpr int FPB = BLOCK / TPF;                  // frame groups per CTA
    // register budget of the forward kernel: all 64 Ki registers of the SM split over the resident warps (allocation
    // granularity: 8 registers per thread).  __maxnreg__ instead of __launch_bounds__, which rounds the CTA up to 128 threads.
    static constexpr int MAXREG_RAW = (65536 / (BLOCK * MINB)) & ~7;
    static constexpr int MAXREG = MAXREG_RAW > 255 ? 255 : MAXREG_RAW;
    static constexpr int XS = (R1 >= 16) ? ilog2(R1) : 5;    // exchange-layout pad shift
    static constexpr int BUF = M + M / 16 + 2;               // complex elements per group buffer: slots 0 .. zslot(M), even
    static constexpr int NW = TPF > 32 ? TPF / 32 : 1;
    static constexpr int SCR = TPF > 32 ? 16 * NW : 2;       // doubles of cross-warp reduction scratch per group
    __host__ __device__ static constexpr size_t off_scratch() { return (size_t)FPB * BUF * sizeof(Cx<T>); }
    __host__ __device__ static constexpr size_t off_bar() { return off_scratch() + (size_t)FPB * SCR * sizeof(double); }
    __host__ __device__ static constexpr size_t off_log() { return (off_bar() + (size_t)FPB * sizeof(uint64_t) + 15) & ~(size_t)15; }
    // f64: the log2 table is replicated over the 8 sixteen-byte bank columns (lane l reads column l & 7), so a lookup
    // with 32 random indices is conflict-free by construction; f32 uses log2f and needs no table
    static constexpr int LOGREP = sizeof(T) == 8 ? 8 : 1;
    __host__ __device__ static constexpr size_t smem_bytes() { return off_log() + (size_t)LOG2_TAB * LOGREP * sizeof(double2); }
    static_assert(LOG2N >= 6 && LOG2N <= 14, "fft_size must be 64..16384");
};

// Everything the kernels need besides the template parameters.  Filled on the host (aus_b200.cu).
template <typename T> struct StftParams {
    const T *audio;                    // all files, contiguous
    const T *mags;                     // K3: [frames][K] magnitudes (audio unused)
    const unsigned long long *file_off;   // [n_files] first sample of each file in `audio`
    const unsigned long long *file_len;   // [n_files]
    const unsigned long long *file_first; // [n_files + 1] first global frame index of each file
    int n_files;
    unsigned long long total_frames;
    int hop;
    int wtype;
    int tma_ok;                        // audio base is 16-byte aligned
    int npass;                         // always 2; a runtime value so the shared radix-32 body stays one rolled loop
    int direct;                        // 1: full aligned frames are read global -> registers (no shared staging); 0: TMA staging
    const Cx<T> *win2;                 // [M] (0.5 w[2n], 0.5 w[2n+1]) -- the 1/2 of the real-FFT split is folded in
    // twiddles are stored in CONSUMPTION order, so the 32 lanes of a warp always read 32 consecutive entries:
    const Cx<T> *tw2;                  // [32][C2] pass-2 twiddles  W_M^(c R3 r) at [r][c], C2 = min(TPF, 32)
    const Cx<T> *tw3;                  // [R3][1024] pass-3 twiddles W_M^(j r) at [r][j]           (M > 1024 only)
    const Cx<T> *twNt;                 // [16][TPF] post-twiddles (cos, sin)(2 pi (16 t + i) / N) at [i][t]
    const Cx<T> *twN;                  // [M/2 + 1] (cos, sin)(2 pi k / N), natural order (inverse pre-twiddle)
    const double2 *log2tab;            // [LOG2_TAB] (1/c_i, -log2(1/c_i))
    double *raw;                       // [frames][RAW_N] per-frame sums for finalize_kernel, or nullptr (no features)
    T *feat;                           // [frames][14] (written by finalize_kernel)
    Cx<T> *spec;                       // [frames][K] or nullptr
    double f0;                         // sample_rate / fft_size   (fft_tools.rs:201)
    double slope_sx, slope_den;        // n(n-1)/2 and n*Sxx - Sx^2 with n = K (spectral_analysis_tools.rs:177-180)
    int pick[6];                       // bins a1-, a1+, b1- (= a2-), b1+ (= a2+), b2-, b2+ (s_a_t.rs:198-201)
    int r_lo[2], r_hi[2];              // half-open bin ranges [ceil(m_fl), floor(m_fu)) of bands 1 and 2
    double band_fl[3], band_fu[3];     // Hz
    double band_tl[3], band_tu[3];     // fractional parts m_fl - floor(m_fl), m_fu - floor(m_fu)
    double band_sx[3], band_den[3];    // sum_x and n*sum_x2 - sum_x^2 (host, reference summation order)
};

// layout of one raw record (doubles): everything finalize_kernel needs to produce the 14 features of a frame
enum RawIdx { RAW_SM = 0, RAW_SMK, RAW_SPK, RAW_B1P, RAW_B1PK, RAW_B2P, RAW_B2PK, RAW_P, RAW_D2, RAW_D3, RAW_D4,
              RAW_ENT, RAW_SLQ, RAW_ROLL0, RAW_ROLL1, RAW_ROLL2, RAW_ROLL3, RAW_PICK0, RAW_N = RAW_PICK0 + 6 };

// window.rs:26-98 evaluated for one index (used only for the short last frame of a file, fft.rs:148-159)
static __device__ __noinline__ double window_value(int wtype, double i, double L) {
    const double PI = 3.14159265358979323846264338327950288;
    switch (wtype) {
    case 0: return (2.0 / (L - 1.0)) * ((L - 1.0) / 2.0 - fabs(i - (L - 1.0) / 2.0));
    case 1: return 0.42 - 0.5 * cos((2.0 * PI * i) / L) + 0.08 * cos((4.0 * PI * i) / L);
    case 2: return 0.5 - 0.5 * cos((2.0 * PI * i) / (L - 1.0));
    case 3: return 0.54 - 0.46 * cos((2.0 * PI * i) / (L - 1.0));
    default: return 1.0;
    }
}

// The short last frame of a file (fft.rs:148-159): L < fft_size samples times the window of length L (half-scaled like
// win2), zero-padded.  Kept out of line so that the hot loop's register allocation does not see it; the loop is inside, so
// the call overhead (a spill of the caller's live registers) is paid once per frame, and eight samples per trip keep eight
// independent cos() chains in flight (one warp works alone here: a dependent FP64 operation is ~25 cycles away).
// window.rs:26-98.  cos(2 pi i / D) is evaluated as cospi(2 i / D): the same value to the last ulp or two (neither device
// function is the reference's libm anyway), but straight-line code -- cos() carries a branch to its large-argument path, which
// keeps the compiler from interleaving the eight evaluations of a trip (measured: 460 cycles per sample instead of ~80).
template <int W> __device__ __forceinline__ double short_window(double id, double Ld) {
    if constexpr (W == 0) return (2.0 / (Ld - 1.0)) * ((Ld - 1.0) / 2.0 - fabs(id - (Ld - 1.0) / 2.0));
    else if constexpr (W == 1) return 0.42 - 0.5 * cospi((2.0 * id) / Ld) + 0.08 * cospi((4.0 * id) / Ld);
    else if constexpr (W == 2) return 0.5 - 0.5 * cospi((2.0 * id) / (Ld - 1.0));
    else if constexpr (W == 3) return 0.54 - 0.46 * cospi((2.0 * id) / (Ld - 1.0));
    else return 1.0;
}
constexpr int SSF_W = 8;       // samples per trip = independent cos() chains in flight
template <typename T, int W> __device__ __forceinline__ void stage_short_frame_w(T *bufr, const T *src, int L, int N, int t, int tpf) {
    const double Ld = (double)L;
    for (int i0 = t; i0 < N; i0 += SSF_W * tpf) {
        if (i0 >= L) {                                                 // the zero padding (i0 is the trip's smallest index)
#pragma unroll
            for (int c = 0; c < SSF_W; ++c)
                if (i0 + c * tpf < N) bufr[i0 + c * tpf] = (T)0;
            continue;
        }
        double w[SSF_W];
        T x[SSF_W];
#pragma unroll
        for (int c = 0; c < SSF_W; ++c) {
            const int i = i0 + c * tpf;
            x[c] = i < L ? src[i] : (T)0;
            w[c] = short_window<W>((double)(i < L ? i : 0), Ld);
        }
#pragma unroll
        for (int c = 0; c < SSF_W; ++c) {
            const int i = i0 + c * tpf;
            if (i < N) bufr[i] = i < L ? (T)((double)x[c] * (0.5 * w[c])) : (T)0;
        }
    }
}
template <typename T>
static __device__ __noinline__ void stage_short_frame(T *bufr, const T *src, int L, int N, int wtype, int t, int tpf) {
    switch (wtype) {
    case 0: stage_short_frame_w<T, 0>(bufr, src, L, N, t, tpf); break;
    case 1: stage_short_frame_w<T, 1>(bufr, src, L, N, t, tpf); break;
    case 2: stage_short_frame_w<T, 2>(bufr, src, L, N, t, tpf); break;
    case 3: stage_short_frame_w<T, 3>(bufr, src, L, N, t, tpf); break;
    default: stage_short_frame_w<T, 4>(bufr, src, L, N, t, tpf); break;
    }
}

template <int SHIFT> __device__ __forceinline__ int padi(int i) { return i + (i >> SHIFT); }

// ---------------------------------------------------------------------------------------------------------
// Forward packed FFT of the frame staged in buf (as N reals == M complex, unpadded).  On return buf holds
// Z'[k] = FFT_M(z')[k] in the "Z layout" (index k + (k >> 4)), where z' = 0.5 * window * z.
template <typename T, int LOG2N>
__device__ __forceinline__ void frame_fft(Cx<T> *buf, const StftParams<T> &p, int t, int grp, bool prewindowed,
                                          const Cx<T> *gsrc, Cx<T> (&v)[32], bool wpre, bool vin = false) {
    using C = Cfg<T, LOG2N>;
    constexpr int M = C::M, TPF = C::TPF, R1 = C::R1, R3 = C::R3, XS = C::XS;

    if constexpr (M == 32) {                                  // one thread per frame: a single radix-32 pass
        static_for<0, 32>([&](auto r_) {
            constexpr int r = decltype(r_)::value;
            Cx<T> z = buf[r];
            if (!prewindowed) { const Cx<T> w = ldg_cx(&p.win2[r]); z.x *= w.x; z.y *= w.y; }
            v[r] = z;
        });
        dft_dif<32>(v);
        static_for<0, 32>([&](auto r_) {
            constexpr int r = decltype(r_)::value;
            buf[padi<4>(r)] = v[bitrev(r, 5)];
        });
        group_sync<TPF, C::BLOCK>(grp);
        return;
    } else if constexpr (R1 == 32) {
        // ---- passes 1 and 2 are both radix 32 with one butterfly per thread (j = t): one rolled loop whose
        //      loads / table reads / stores are linear in r:  address = base + r * stride.
        //      pass 1 (Ns = 1):  in  = buf[j + r TPF] (raw samples)   x win2[j + r TPF]      out = X[33 j + r]
        //      pass 2 (Ns = 32): in  = X[pad5(j + r TPF)]             x tw2[r][j mod 32]     out = Z or X layout
        const int j = t;
        int lbase = j, lstride = TPF;
        const Cx<T> *tab = p.win2 + j;
        int tstride = TPF;
        int sbase = 33 * j, sstride = 1;
#pragma unroll 1
        for (int pass = 0; pass < p.npass; ++pass) {
            if (pass == 0) {
                if (vin) {                                    // v[r] already holds the pass-1 input buf[j + r TPF] (K4, R1 == 32)
                } else if (gsrc != nullptr) {                 // samples straight from global memory (16-byte, coalesced)
#pragma unroll
                    for (int r = 0; r < 32; ++r) v[r] = emul(ldg_cx(gsrc + lbase + r * lstride), ldg_cx(tab + r * tstride));
                } else if (prewindowed) {
#pragma unroll
                    for (int r = 0; r < 32; ++r) v[r] = buf[lbase + r * lstride];
                } else if (wpre) {                            // v already holds the window (loaded before the wait for the frame)
#pragma unroll
                    for (int r = 0; r < 32; ++r) v[r] = emul(buf[lbase + r * lstride], v[r]);
                } else {
#pragma unroll
                    for (int r = 0; r < 32; ++r) v[r] = emul(buf[lbase + r * lstride], ldg_cx(tab + r * tstride));
                }
            } else {
                v[0] = buf[lbase];
                for_pass2_twiddles<T>(tab, tstride, [&](auto r_, const Cx<T> &w) {
                    constexpr int r = decltype(r_)::value;
                    v[r] = cmul(buf[lbase + r * lstride], w);
                });
            }
            group_sync<TPF, C::BLOCK>(grp);                   // every load done before buf is overwritten
            dft_dif<32>(v);
            static_for<0, 32>([&](auto r_) {
                constexpr int r = decltype(r_)::value;
                buf[sbase + r * sstride] = v[bitrev(r, 5)];
            });
            group_sync<TPF, C::BLOCK>(grp);
            // address generators of pass 2
            lbase = j + (j >> 5);
            lstride = TPF + TPF / 32;
            tab = p.tw2 + (j & 31);                            // W_M^((j mod 32) R3 r) at tw2[r][j mod 32]
            tstride = 32;
            if constexpr (!C::THREE) {
                sbase = j + (j >> 4);                          // Z layout: pad4(j + 32 r) = j + (j >> 4) + 34 r
                sstride = 34;
            } else {
                const int b = (j >> 5) * 1024 + (j & 31);      // exchange layout: pad5(b + 32 r) = b + (b >> 5) + 33 r
                sbase = b + (b >> 5);
                sstride = 33;
            }
        }
    } else {
        // ---- pass 1: radix R1 < 32, Ns = 1, B1 = 32 / R1 butterflies per thread, input = windowed samples
        constexpr int B1 = 32 / R1;
        static_for<0, B1>([&](auto b_) {
            constexpr int b = decltype(b_)::value;
            const int j = t + TPF * b;
            static_for<0, R1>([&](auto r_) {
                constexpr int r = decltype(r_)::value;
                const int n = j + r * (M / R1);
                Cx<T> z = buf[n];
                if (!prewindowed) { const Cx<T> w = ldg_cx(&p.win2[n]); z.x *= w.x; z.y *= w.y; }
                v[b * R1 + r] = z;
            });
        });
        group_sync<TPF, C::BLOCK>(grp);
        static_for<0, B1>([&](auto b_) {
            constexpr int b = decltype(b_)::value;
            const int j = t + TPF * b;
            Cx<T> u[R1];
            static_for<0, R1>([&](auto r_) { u[decltype(r_)::value] = v[b * R1 + decltype(r_)::value]; });
            dft_dif<R1>(u);
            static_for<0, R1>([&](auto r_) {
                constexpr int r = decltype(r_)::value;
                buf[padi<XS>(j * R1 + r)] = u[bitrev(r, ilog2(R1))];
            });
        });
        group_sync<TPF, C::BLOCK>(grp);
        // ---- pass 2: radix 32, Ns = R1 = TPF, one butterfly per thread (j = t), natural-order output
        {
            const int j = t;
            v[0] = buf[padi<XS>(j)];
            for_pass2_twiddles<T>(p.tw2 + j, TPF, [&](auto r_, const Cx<T> &w) {
                constexpr int r = decltype(r_)::value;
                v[r] = cmul(buf[padi<XS>(j + r * TPF)], w);
            });
            group_sync<TPF, C::BLOCK>(grp);
            dft_dif<32>(v);
            static_for<0, 32>([&](auto r_) {
                constexpr int r = decltype(r_)::value;
                buf[padi<4>(j + r * TPF)] = v[bitrev(r, 5)];
            });
            group_sync<TPF, C::BLOCK>(grp);
        }
    }

    // ---- pass 3 (M > 1024): radix R3, Ns = 1024, B3 = 32 / R3 butterflies per thread
    if constexpr (C::THREE) {
        constexpr int B3 = 32 / R3;
        static_for<0, B3>([&](auto b_) {
            constexpr int b = decltype(b_)::value;
            const int j = t + TPF * b;
            static_for<0, R3>([&](auto r_) {
                constexpr int r = decltype(r_)::value;
                Cx<T> z = buf[padi<XS>(j + r * 1024)];
                if constexpr (r > 0) z = cmul(z, ldg_cx(&p.tw3[r * 1024 + j]));
                v[b * R3 + r] = z;
            });
        });
        group_sync<TPF, C::BLOCK>(grp);
        static_for<0, B3>([&](auto b_) {
            constexpr int b = decltype(b_)::value;
            const int j = t + TPF * b;
            Cx<T> u[R3];
            static_for<0, R3>([&](auto r_) { u[decltype(r_)::value] = v[b * R3 + decltype(r_)::value]; });
            dft_dif<R3>(u);
            static_for<0, R3>([&](auto r_) {
                constexpr int r = decltype(r_)::value;
                buf[padi<4>(j + r * 1024)] = u[bitrev(r, ilog2(R3))];
            });
        });
        group_sync<TPF, C::BLOCK>(grp);
    }
}

// X[k] (k <= M/2) and X[M-k] from the packed transform: A = Z'[k], B = Z'[M-k], W = (cos, sin)(2 pi k / N).
template <typename T>
__device__ __forceinline__ void post_twiddle(Cx<T> A, Cx<T> B, Cx<T> W, Cx<T> &Xlo, Cx<T> &Xhi) {
    const T ex = A.x + B.x, ey = A.y - B.y;          // E = A + conj(B)
    const T ox = A.y + B.y, oy = B.x - A.x;          // O = -i (A - conj(B))
    const T tx = W.x * ox + W.y * oy;                // T = (c - i s) * O
    const T ty = W.x * oy - W.y * ox;
    Xlo = Cx<T>{ex + tx, ey + ty};
    Xhi = Cx<T>{ex - tx, ty - ey};                   // conj(E - T)
}

// Per-thread running sums of the first feature pass (arithmetic type T; cross-thread combination is always double).
template <typename T> struct FeatAcc {
    T s_m = 0, s_mk = 0, s_pk = 0, lo_sum = 0, hi_sum = 0, mid = 0;
    T s_m2 = 0, s_mk2 = 0, s_pk2 = 0;                 // second accumulator set (the high bins): halves the add chains
    unsigned pmin = 0x7ff00000u;                      // smallest power bin, as the high word of its f64 pattern (f32: bits >> 1)
    __device__ __forceinline__ void add(T kf, T m, T pw) {
        s_m += m;
        s_mk += m * kf;
        s_pk += pw * kf;
    }
    __device__ __forceinline__ void add2(T kf, T m, T pw) {
        s_m2 += m;
        s_mk2 += m * kf;
        s_pk2 += pw * kf;
    }
    __device__ __forceinline__ void merge() { s_m += s_m2; s_mk += s_mk2; s_pk += s_pk2; }
};
#ifndef AUS_SPLIT_TAIL_ACC
#define AUS_SPLIT_TAIL_ACC 0     // measured neutral at fft 2048 (1.277e8 vs 1.281e8 frames/s)
#endif
#ifndef AUS_TAIL_PAIRS
#define AUS_TAIL_PAIRS 4         // pairs (k, M - k) per trip of the first feature pass
#endif
#ifndef AUS_EARLY_TMA
#define AUS_EARLY_TMA 0          // frame_features32: issue the next frame's bulk copy as soon as the group buffer is free.
                                 // Measured SLOWER (1.23e8 vs 1.27e8 frames/s, twice, also with a load-free locate): off
#endif
#ifndef AUS_SWEEP_W
#define AUS_SWEEP_W 4            // bins per trip of frame_features32's sweep
#endif

// Slot of bin b in the group buffer ("Z layout").
__device__ __forceinline__ int zslot(int b) { return b + (b >> 4); }

// During the feature passes a slot holds two scalars: the bin's power and the running sum up to it.  Lanes 8
// (f64) / 16 (f32) apart map to the same banks (slot stride 17), so which half holds the power alternates with
// bit 3 (4) of the owning thread: half = hpar(owner) for the power, 1 - hpar(owner) for the running sum.
template <typename T> __device__ __forceinline__ int hpar(int owner) { return (owner >> (sizeof(T) == 8 ? 3 : 4)) & 1; }
template <int M, int TPF> __device__ __forceinline__ int bin_owner(int b) {
    return b < M / 2 ? (b >> 4) : (b == M / 2 ? TPF - 1 : ((M - b) >> 4));
}
template <typename T, int M, int TPF> __device__ __forceinline__ T &pow_of(Cx<T> *buf, int b) {
    return reinterpret_cast<T *>(buf)[2 * zslot(b) + hpar<T>(bin_owner<M, TPF>(b))];
}
template <typename T, int M, int TPF> __device__ __forceinline__ T &cum_of(Cx<T> *buf, int b) {
    return reinterpret_cast<T *>(buf)[2 * zslot(b) + 1 - hpar<T>(bin_owner<M, TPF>(b))];
}

// ---------------------------------------------------------------------------------------------------------
// Feature passes over one frame.  On entry slot(b).x = |X[b]|^2 for every bin b = 0..M (written by the owning
// threads, not yet synchronised).  Thread t owns the bins [16 t, 16 t + 16) ("low") and (M - 16 t - 16, M - 16 t]
// ("high"); the last thread also owns bin M/2.  Writes the frame's raw record; finalize_kernel turns it into the 14
// features (one thread per frame, so the scalar divisions / exp / log run SIMD-efficiently).
template <typename T, int LOG2N>
__device__ __forceinline__ void frame_features(Cx<T> *buf, FeatAcc<T> &acc, double *scratch, const double2 *logtab,
                                               const StftParams<T> &p, unsigned long long g, bool active, int t,
                                               int grp, bool mags_from_sqrt) {
    using C = Cfg<T, LOG2N>;
    constexpr int M = C::M, K = C::K, TPF = C::TPF;
    const int hb = M - 16 * t - 15;                   // first (lowest) bin of this thread's high chunk

    // The fast sqrt / log2 need positive normal arguments.  Powers are >= 0, so one running minimum (an integer min
    // per bin) plus the finiteness of the totals decides it for the whole frame: a zero / subnormal / non-finite power
    // bin sends the frame through the exact library functions, which also reproduce the reference's NaN / -inf rules.
    // (f32 uses sqrtf / log2f throughout, which handle every argument; K3 gets its magnitudes from the caller.)
    const bool bad_p = sizeof(T) == 8 && mags_from_sqrt && active &&   // (a group past the end of the batch skips the cold paths)
                       (!power_is_normal<T>(acc.pmin) || !(acc.lo_sum + acc.hi_sum + acc.mid < (T)INFINITY));
    if (group_any<TPF, C::BLOCK>(bad_p, scratch, t, grp)) {
        acc.s_m = 0;
        acc.s_mk = 0;
#pragma unroll 1
        for (int u = 0; u < 33; ++u) {
            const int b = u < 16 ? 16 * t + u : (u < 32 ? hb + (u - 16) : M / 2);
            if (u == 32 && t != TPF - 1) break;
            const T m = Mth<T>::sqrt_exact(pow_of<T, M, TPF>(buf, b));
            acc.s_m += m;
            acc.s_mk += m * (T)b;
        }
    }
    group_sync<TPF, C::BLOCK>(grp);                   // every power bin visible to the whole group

    // sub-band sums for the three slope regions (s_a_t.rs:190-230): bins [r_lo, r_hi) of bands 1 and 2, all threads
    T b1p = 0, b1pk = 0, b2p = 0, b2pk = 0;
    const bool bands_low = p.r_hi[1] <= M / 2;        // every band bin in the lower half: owner = b >> 4, no selects
#pragma unroll 1
    for (int b = p.r_lo[0] + t; b < p.r_hi[1]; b += TPF) {
        const T pw = bands_low ? reinterpret_cast<const T *>(buf)[2 * zslot(b) + hpar<T>(b >> 4)] : pow_of<T, M, TPF>(buf, b);
        const T pk = pw * (T)b;
        if (b < p.r_hi[0]) { b1p += pw; b1pk += pk; }
        if (b >= p.r_lo[1]) { b2p += pw; b2pk += pk; }
    }

    double r[10] = {(double)acc.s_m, (double)acc.s_mk, (double)acc.s_pk, (double)b1p, (double)b1pk, (double)b2p,
                    (double)b2pk, (double)acc.lo_sum, (double)acc.hi_sum, (double)acc.mid};
    group_allreduce<TPF, C::BLOCK, 10>(r, scratch, t, grp);
    const double S_m = r[0], S_mk = r[1];
    const double tot_lo = r[7], tot_hi = r[8], tot_mid = r[9];
    const double P = tot_lo + tot_mid + tot_hi;

    // prefix of the power spectrum in bin order: low chunks ascending t, bin M/2, high chunks descending t
    double inc_lo = (double)acc.lo_sum, inc_hi = (double)acc.hi_sum;
    group_inclusive_scan2<TPF, C::BLOCK>(inc_lo, inc_hi, scratch, t, grp);
    const T base_lo = (T)(inc_lo - (double)acc.lo_sum);
    const T base_hi = (T)(tot_lo + tot_mid + (tot_hi - inc_hi));

    // central moments about the (magnitude-weighted) centroid, in units of bins; entropy and log-mean from
    // q = p / P (a nearly pure tone keeps full relative accuracy, like the reference's pmf); running sum -> slot.y
    const T kappa = (T)(S_mk / S_m);
    const T inv_p = (T)(1.0 / P);
    T d2s = 0, d3s = 0, d4s = 0, ent = 0, slq = 0;
    // q = p / P of the smallest bin must itself be a positive normal number (P finite and > 0 implied)
    const unsigned gpmin = group_min_u32<TPF, C::BLOCK>(acc.pmin, scratch, t, grp);
    const bool badq = sizeof(T) == 8 && active && (!power_is_normal<T>(gpmin) || !(P > 0.0 && P < (double)INFINITY) ||
                                                   abnormal((T)(power_from_hi<T>(gpmin) * inv_p)));
    // four bins at a time: all shared-memory loads first, then four independent sqrt-free math chains, so the
    // load-to-use latency is paid once per block.  s0 = slot of the first bin, the 4th bin sits at s0 + 3 + skip.
    T *const bufT = reinterpret_cast<T *>(buf);
    const int hp = hpar<T>(t);                        // this thread's power half; running sum in the other one
    const double2 *const ltab = logtab + (C::LOGREP > 1 ? (threadIdx.x & 7) : 0);   // this lane's bank column
    constexpr int PW = AUS_PASSB_W;                   // bins per block: that many independent log2 chains in flight
    auto blockw = [&](int s0, int skip, T d0, T &cum) {
        T pw[PW], q[PW], l[PW];
#pragma unroll
        for (int c = 0; c < PW; ++c) pw[c] = bufT[2 * (s0 + c + (c == PW - 1 ? skip : 0)) + hp];
#pragma unroll
        for (int c = 0; c < PW; ++c) {
            const T d = d0 + (T)c;
            const T dd = d * d;
            const T w2 = dd * pw[c];
            d2s += w2;
            d3s += w2 * d;
            d4s += w2 * dd;
            q[c] = pw[c] * inv_p;                     // make_spectrum_pmf, s_a_t.rs:284-290
        }
#pragma unroll
        for (int c = 0; c < PW; ++c) l[c] = Mth<T>::log2_(q[c], ltab);
#pragma unroll
        for (int c = 0; c < PW; ++c) {
            ent += q[c] * l[c];
            slq += l[c];
            cum += pw[c];
            bufT[2 * (s0 + c + (c == PW - 1 ? skip : 0)) + 1 - hp] = cum;
        }
    };
    {
        T cum = base_lo;
        T dlo = (T)(16 * t) - kappa;                   // running (bin - centroid): no integer -> double conversion per block
#pragma unroll 1
        for (int i0 = 0; i0 < 16; i0 += PW, dlo += (T)PW) blockw(17 * t + i0, 0, dlo, cum);   // zslot(16 t + i) = 17 t + i
        // high chunk, ascending bins hb .. hb + 15: slots hs .. hs + 14 and hs + 16 (hs + 15 is the pad slot)
        cum = base_hi;
        const int hs = zslot(hb);
        T dhi = (T)hb - kappa;
#pragma unroll 1
        for (int i0 = 0; i0 < 16; i0 += PW, dhi += (T)PW) blockw(hs + i0, i0 == 16 - PW ? 1 : 0, dhi, cum);
        if (t == TPF - 1) {                            // bin M/2 sits between the low and the high chunks
            const T pw = bufT[2 * zslot(M / 2) + hp];
            const T d = (T)(M / 2) - kappa;
            const T dd = d * d;
            const T w2 = dd * pw;
            d2s += w2;
            d3s += w2 * d;
            d4s += w2 * dd;
            const T q = pw * inv_p;
            const T l = Mth<T>::log2_(q, ltab);
            ent += q * l;
            slq += l;
            bufT[2 * zslot(M / 2) + 1 - hp] = (T)tot_lo + pw;
        }
    }
    // zero / subnormal / non-finite q somewhere: redo the two log sums with the library log2 (0 * -inf = NaN and
    // sum = -inf exactly like pmf * log2(pmf) and ln(m) in the reference, s_a_t.rs:93-114)
    if (badq) {                                        // group-uniform
        ent = 0;
        slq = 0;
#pragma unroll 1
        for (int u = 0; u < 33; ++u) {
            const int b = u < 16 ? 16 * t + u : (u < 32 ? hb + (u - 16) : M / 2);
            if (u == 32 && t != TPF - 1) break;
            const T q = pow_of<T, M, TPF>(buf, b) * inv_p;
            const T l = Mth<T>::log2_exact(q);
            ent += q * l;
            slq += l;
        }
    }
    double q[5] = {(double)d2s, (double)d3s, (double)d4s, (double)ent, (double)slq};
    group_allreduce<TPF, C::BLOCK, 5>(q, scratch, t, grp);      // (its barriers also publish every slot.y)
    if constexpr (TPF <= 32) __syncwarp();

    if (active && p.raw != nullptr) {
        double *rec = p.raw + g * (unsigned long long)RAW_N;
        // roll-off (s_a_t.rs:137-145): index = number of bins whose inclusive cumulative power is below theta * P.
        // The running sums in slot.y are monotone over b = 0..M: one binary search per threshold, one thread each.
#pragma unroll 1
        for (int qq = t; qq < 4; qq += TPF) {
            const T th = (T)((qq == 0 ? 0.5 : qq == 1 ? 0.75 : qq == 2 ? 0.9 : 0.95) * P);
            int lo = 0, hi = K;                        // first index in [0, K] whose running sum is not below th
#pragma unroll 1
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (cum_of<T, M, TPF>(buf, mid) < th) lo = mid + 1; else hi = mid;
            }
            if constexpr (sizeof(T) == 8) {
                // Knife edge: the threshold sits within rounding distance of a running sum, where the reference's own
                // arithmetic -- a sequential sum of p[i] / P compared with theta (s_a_t.rs:137-145) -- decides.  Redo it
                // exactly that way (one thread, rare): identical result whenever the power values are identical, e.g.
                // analyzer() on caller-supplied magnitudes with exact ties.
                const double tol = 1e-12 * P;
                const bool amb = (lo <= M && fabs((double)cum_of<T, M, TPF>(buf, lo) - (double)th) <= tol) ||
                                 (lo >= 1 && fabs((double)th - (double)cum_of<T, M, TPF>(buf, lo - 1)) <= tol);
                if (amb && !(P > 0.0)) {                   // silent (or NaN) frame: the reference's loop stops at bin 0
                    lo = 0;
                } else if (amb) {
                    const double theta = qq == 0 ? 0.5 : qq == 1 ? 0.75 : qq == 2 ? 0.9 : 0.95;
                    double pseq = 0.0;                     // power_spectrum.iter().sum(): sequential (analyzer.rs:44)
                    for (int i = 0; i < K; ++i) pseq += (double)pow_of<T, M, TPF>(buf, i);
                    double csum = 0.0;
                    int i = -1;
                    while (csum < theta && i < K - 1) {
                        ++i;
                        csum += (double)pow_of<T, M, TPF>(buf, i) / pseq;
                    }
                    lo = i;
                }
            }
            rec[RAW_ROLL0 + qq] = (double)lo;
        }
#pragma unroll 1
        for (int qq = t; qq < 6; qq += TPF) rec[RAW_PICK0 + qq] = (double)pow_of<T, M, TPF>(buf, p.pick[qq]);
        if (t == 0) {
            rec[RAW_SM] = r[0]; rec[RAW_SMK] = r[1]; rec[RAW_SPK] = r[2];
            rec[RAW_B1P] = r[3]; rec[RAW_B1PK] = r[4]; rec[RAW_B2P] = r[5]; rec[RAW_B2PK] = r[6];
            rec[RAW_P] = P;
            rec[RAW_D2] = q[0]; rec[RAW_D3] = q[1]; rec[RAW_D4] = q[2]; rec[RAW_ENT] = q[3]; rec[RAW_SLQ] = q[4];
        }
    }
}

#ifndef AUS_COLD_OUTLINE
#define AUS_COLD_OUTLINE 0       // experiment: frame_features32() with its silent-frame and knife-edge fallbacks out of line
#endif
#ifndef AUS_STAGGER_NS
#define AUS_STAGGER_NS 0         // experiment: start frame group g of a CTA g * AUS_STAGGER_NS later (short launches run in lockstep)
#endif
#ifndef AUS_PSUM_ENTROPY
#define AUS_PSUM_ENTROPY 0       // experiment: entropy / flatness sums over p instead of q = p / P (one multiply per bin less)
#endif
#ifndef AUS_FEAT32
#define AUS_FEAT32 1             // f64, one warp per frame: latency-trimmed second half of the feature computation (below)
#endif

// Sum of NV <= 16 doubles over a warp with 16 instead of 5 NV 64-bit shuffles (recursive halving): every step a lane keeps
// the half of the values its lane bit selects and adds the partner's copy; afterwards lane t holds the total of value
// t >> 1 (values >= NV are zero).  Fixed order -> deterministic.
template <int NV> __device__ __forceinline__ double warp_reduce_scatter(const double (&in)[NV], int t) {
    static_assert(NV <= 16, "at most 16 values");
    double v8[8];
    {
        const bool h = (t & 16) != 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const double lo = in[i], hi = i + 8 < NV ? in[i + 8] : 0.0;
            const double keep = h ? hi : lo, send = h ? lo : hi;
            v8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    double v4[4];
    {
        const bool h = (t & 8) != 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double keep = h ? v8[i + 4] : v8[i], send = h ? v8[i] : v8[i + 4];
            v4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    double v2[2];
    {
        const bool h = (t & 4) != 0;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const double keep = h ? v4[i + 2] : v4[i], send = h ? v4[i] : v4[i + 2];
            v2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    const bool h = (t & 2) != 0;
    const double keep = h ? v2[1] : v2[0], send = h ? v2[0] : v2[1];
    double v1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
    return v1;
}

// Cold paths of frame_features32(), out of line so that the hot function's code and register allocation do not carry them:
// silent / subnormal / non-finite frames (exact sqrt and log2, the reference's NaN / -inf rules) and the knife-edge roll-off
// threshold (the reference's own sequential arithmetic, s_a_t.rs:137-145).
template <int M, int TPF>
static __device__ __noinline__ double2 cold_exact_magnitudes(Cx<double> *buf, int t) {
    double a = 0.0, b2 = 0.0;
#pragma unroll 1
    for (int u = 0; u < 33; ++u) {
        if (u == 32 && t != TPF - 1) break;
        const int b = u < 16 ? 16 * t + u : (u < 32 ? M - 16 * t - 15 + (u - 16) : M / 2);
        const double m = sqrt(pow_of<double, M, TPF>(buf, b));
        a += m;
        b2 += m * (double)b;
    }
    return double2{a, b2};
}
template <int M, int TPF>
static __device__ __noinline__ double2 cold_exact_log_sums(Cx<double> *buf, int t, double inv_p) {
    double e = 0.0, l2 = 0.0;                            // library log2: 0 * -inf = NaN, sum = -inf, like the reference
#pragma unroll 1
    for (int u = 0; u < 33; ++u) {
        if (u == 32 && t != TPF - 1) break;
        const int b = u < 16 ? 16 * t + u : (u < 32 ? M - 16 * t - 15 + (u - 16) : M / 2);
        const double q = pow_of<double, M, TPF>(buf, b) * inv_p;
        const double l = log2(q);
        e += q * l;
        l2 += l;
    }
    return double2{e, l2};
}
template <int M, int TPF>
static __device__ __noinline__ int cold_rolloff_sequential(Cx<double> *buf, double theta) {
    constexpr int K = M + 1;
    double pseq = 0.0;                                   // power_spectrum.iter().sum(), analyzer.rs:44
    for (int i = 0; i < K; ++i) pseq += pow_of<double, M, TPF>(buf, i);
    double csum = 0.0;
    int i = -1;
    while (csum < theta && i < K - 1) {
        ++i;
        csum += pow_of<double, M, TPF>(buf, i) / pseq;
    }
    return i;
}

// ---------------------------------------------------------------------------------------------------------
// frame_features() for f64 frame groups of exactly one warp (fft_size 2048), rewritten around what the clock64() phase
// timings of profiles/r2_k1_pc_experiment.md showed: a warp spends most of this half waiting on dependency chains (a
// dependent FP64 operation is ~25 cycles away, a dependent shared-memory load ~60), not on any pipe.  Same arithmetic per
// bin (q = p / P, table log2, central moments about the magnitude centroid), but
//   * no per-bin running sum is stored: the chunk sums of the first pass, scanned over the lanes, give the running sum at
//     the 64 chunk ends in registers; two ballots per threshold find the chunk, a sum tree over its 16 bins finds the bin
//     (was: 32 dependent adds + stores per thread, then a 10-step binary search of dependent loads by four lanes);
//   * the sub-band sums load eight bins at once and add them as a tree (was: a rolled loop of dependent load -> multiply -> add);
//   * two accumulator sets in the per-bin sweep (halves the add chains), sums combined by recursive halving and a scan
//     (39 instead of 85 64-bit shuffles), the record written by the lanes that hold the values.
// Knife-edge thresholds fall back to the reference's own sequential arithmetic exactly like frame_features().
template <int LOG2N>
__device__ __forceinline__ void frame_features32(Cx<double> *buf, FeatAcc<double> &acc, const double2 *logtab,
                                                 const StftParams<double> &p, unsigned long long g, bool active, int t,
                                                 const double *next_src, uint64_t *bar) {
    using T = double;
    using C = Cfg<T, LOG2N>;
    constexpr int M = C::M, K = C::K, TPF = C::TPF;
    static_assert(TPF == 32, "one warp per frame");
    const unsigned FULL = 0xffffffffu;
    T *const bufT = reinterpret_cast<T *>(buf);
    const int hp = hpar<T>(t);
    // power of bin 16 c + i (low chunk of lane c) / bin M - 16 c - i (its high chunk), i = 0..15
    auto lo_ptr = [&](int c) { return bufT + 2 * (17 * c) + hpar<T>(c); };                  // element i at [2 i]
    auto hi_at = [&](int c, int i) -> T {
        const int h0c = zslot(M - 16 * c);
        return bufT[2 * (i == 0 ? h0c : h0c - 1 - i) + hpar<T>(c)];
    };
    // ---- zero / subnormal / non-finite power somewhere: magnitudes again with the exact square root (cold)
    const bool bad_p = !power_is_normal<T>(acc.pmin) || !(acc.lo_sum + acc.hi_sum + acc.mid < (T)INFINITY);
#if AUS_COLD_OUTLINE
    if (__any_sync(FULL, bad_p)) {
        const double2 mm = cold_exact_magnitudes<M, TPF>(buf, t);
        acc.s_m = mm.x;
        acc.s_mk = mm.y;
    }
#else
    if (__any_sync(FULL, bad_p)) {
        acc.s_m = 0;
        acc.s_mk = 0;
#pragma unroll 1
        for (int u = 0; u < 33; ++u) {
            if (u == 32 && t != TPF - 1) break;
            const int b = u < 16 ? 16 * t + u : (u < 32 ? M - 16 * t - 15 + (u - 16) : M / 2);
            const T m = sqrt(pow_of<T, M, TPF>(buf, b));
            acc.s_m += m;
            acc.s_mk += m * (T)b;
        }
    }
#endif
    __syncwarp();                                     // every power bin visible to the whole warp

    // ---- sub-band sums (s_a_t.rs:190-230), eight bins per trip: loads first, independent products, tree sums
    T b1p = 0, b1pk = 0, b2p = 0, b2pk = 0;
#pragma unroll 1
    for (int b0 = p.r_lo[0] + t; b0 < p.r_hi[1]; b0 += 8 * TPF) {
        T pv[8], a1[8], k1[8], a2[8], k2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int b = b0 + j * TPF;
            pv[j] = b < p.r_hi[1] ? pow_of<T, M, TPF>(buf, b) : (T)0;
        }
        const T bf0 = (T)b0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int b = b0 + j * TPF;
            const T pk = pv[j] * (bf0 + (T)(j * TPF));
            const bool in1 = b < p.r_hi[0], in2 = b >= p.r_lo[1];
            a1[j] = in1 ? pv[j] : (T)0; k1[j] = in1 ? pk : (T)0;
            a2[j] = in2 ? pv[j] : (T)0; k2[j] = in2 ? pk : (T)0;
        }
        auto tree8 = [](const T (&x)[8]) { return ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7])); };
        b1p += tree8(a1); b1pk += tree8(k1); b2p += tree8(a2); b2pk += tree8(k2);
    }
    // ---- totals every lane needs: magnitude sums (butterfly), chunk sums (inclusive scan; its last lane has the totals)
    T sm2[2] = {acc.s_m, acc.s_mk};
    T inc_lo = acc.lo_sum, inc_hi = acc.hi_sum;
    unsigned gpmin = acc.pmin;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        sm2[0] += __shfl_xor_sync(FULL, sm2[0], off);
        sm2[1] += __shfl_xor_sync(FULL, sm2[1], off);
        const T ua = __shfl_up_sync(FULL, inc_lo, off), ub = __shfl_up_sync(FULL, inc_hi, off);
        if (t >= off) { inc_lo += ua; inc_hi += ub; }
        gpmin = min(gpmin, __shfl_xor_sync(FULL, gpmin, off));
    }
    const T S_m = sm2[0], S_mk = sm2[1];
    const T tot_lo = __shfl_sync(FULL, inc_lo, 31), tot_hi = __shfl_sync(FULL, inc_hi, 31), tot_mid = __shfl_sync(FULL, acc.mid, 31);
    const T P = tot_lo + tot_mid + tot_hi;
    const T kappa = S_mk / S_m;
    const T inv_p = 1.0 / P;
    const bool badq = !power_is_normal<T>(gpmin) || !(P > 0.0 && P < (double)INFINITY) ||
                      abnormal((T)(power_from_hi<T>(gpmin) * inv_p));

    // ---- roll-off (s_a_t.rs:137-145): index = number of bins whose inclusive running sum is below theta * P.  Chunk ends in
    //      ascending bin order: low chunks of lanes 0..31, bin M/2, high chunks of lanes 31..0.  Lane 2 q + 1 settles
    //      threshold q: two ballots find the chunk, a sum tree over the chunk's 16 bins finds the bin.
    const T base_lo = inc_lo - acc.lo_sum;                               // running sum before this lane's low chunk
    const T base_hi = tot_lo + tot_mid + (tot_hi - inc_hi);              // ... before its high chunk
    const T end_hi = base_hi + acc.hi_sum;
    const int vi = t >> 1;
    const int wq = (t & 1) ? vi : 4;                                     // threshold this lane settles (4: none)
    int n_lo = 0, n_hi = 0;
    T my_th = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const T th = (q == 0 ? 0.5 : q == 1 ? 0.75 : q == 2 ? 0.9 : 0.95) * P;
        const int nl = __popc(__ballot_sync(FULL, inc_lo < th));        // NaN compares false: 0, walk chunk 0
        const int nh = __popc(__ballot_sync(FULL, end_hi < th));
        if (wq == q) { n_lo = nl; n_hi = nh; my_th = th; }
    }
    const int c_lo = n_lo < 32 ? n_lo : 31, c_hi = n_hi < 32 ? 31 - n_hi : 0;
    const T bl = __shfl_sync(FULL, base_lo, c_lo), bh = __shfl_sync(FULL, base_hi, c_hi);
    int roll_res = 0;
    if (wq < 4) {
        int res = 0, first = 0;
        bool walk = false, high = false, amb = false;
        T base = 0.0, c_at = 0.0, c_prev = 0.0;
        if (n_lo < 32) {                                                 // crossing inside low chunk n_lo
            first = 16 * c_lo; base = bl; walk = true;
        } else if (!(tot_lo + tot_mid < my_th)) {                        // bin M/2
            res = M / 2; c_at = tot_lo + tot_mid; c_prev = tot_lo;
        } else if (n_hi < 32) {                                          // high chunk of lane c_hi, ascending bins
            first = M - 16 * c_hi - 15; base = bh; walk = true; high = true;
        } else {                                                         // even the last bin stays below (rounding): K
            res = K; c_prev = P;
        }
        if (walk) {
            T x[16];                                                     // the chunk's powers in ascending bin order
            if (!high) {
                const T *const lp = lo_ptr(c_lo);
#pragma unroll
                for (int e = 0; e < 16; ++e) x[e] = lp[2 * e];
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) x[e] = hi_at(c_hi, 15 - e);
            }
            T s2[8], s4[4];
#pragma unroll
            for (int e = 0; e < 8; ++e) s2[e] = x[2 * e] + x[2 * e + 1];
#pragma unroll
            for (int e = 0; e < 4; ++e) s4[e] = s2[2 * e] + s2[2 * e + 1];
            const T s8 = s4[0] + s4[1];
            T run = base;
            int j = 0;
            T c = run + s8;                                              // bins 0..7 of the chunk
            if (c < my_th) { run = c; j = 8; }
            c = run + (j == 8 ? s4[2] : s4[0]);
            if (c < my_th) { run = c; j += 4; }
            c = run + (j == 12 ? s2[6] : j == 8 ? s2[4] : j == 4 ? s2[2] : s2[0]);
            if (c < my_th) { run = c; j += 2; }
            T xa = x[0], xb = x[1];
#pragma unroll
            for (int e = 2; e < 16; e += 2)
                if (j == e) { xa = x[e]; xb = x[e + 1]; }
            c = run + xa;
            if (c < my_th) { run = c; j += 1; c = run + xb; }
            if (c < my_th) amb = true;                                   // the chunk ends and this tree disagree by rounding
            res = first + j;
            c_at = c;
            c_prev = run;
        }
        // Knife edge: the threshold sits within rounding distance of a running sum, where the reference's own arithmetic
        // -- a sequential sum of p[i] / P compared with theta -- decides.  Redo it exactly that way (rare).
        const double tol = 1e-12 * P;
        if (!amb && res < K) amb = fabs(c_at - my_th) <= tol || (res >= 1 && fabs(my_th - c_prev) <= tol);
        if (!amb && res == K) amb = fabs(my_th - c_prev) <= tol;
        if (amb && !(P > 0.0)) {                                         // silent (or NaN) frame: the reference's loop stops at bin 0
            res = 0;
        } else if (amb) {
            const double theta = wq == 0 ? 0.5 : wq == 1 ? 0.75 : wq == 2 ? 0.9 : 0.95;
#if AUS_COLD_OUTLINE
            res = cold_rolloff_sequential<M, TPF>(buf, theta);
#else
            double pseq = 0.0;                                           // power_spectrum.iter().sum(), analyzer.rs:44
            for (int i = 0; i < K; ++i) pseq += pow_of<T, M, TPF>(buf, i);
            double csum = 0.0;
            int i = -1;
            while (csum < theta && i < K - 1) {
                ++i;
                csum += pow_of<T, M, TPF>(buf, i) / pseq;
            }
            res = i;
#endif
        }
        roll_res = res;
    }
    // the six bins the sub-band interpolation picks (s_a_t.rs:198-201), fetched now: nothing may read the buffer after the sweep
    T pick_val = 0.0;
    if ((t & 1) && t >= 9 && t < 21) pick_val = pow_of<T, M, TPF>(buf, p.pick[(t - 9) / 2]);
    // ---- the per-bin sweep: central moments about the centroid (in bins), q log2 q, log2 q; two accumulator sets
    const double2 *const ltab = logtab + (threadIdx.x & 7);
    T d2s[2] = {0, 0}, d3s[2] = {0, 0}, d4s[2] = {0, 0}, ent[2] = {0, 0}, slq[2] = {0, 0};
    constexpr int SW = AUS_SWEEP_W;
    auto sweep4 = [&](const T (&pv)[SW], T d0) {
        T q[SW], l[SW];
#if AUS_PSUM_ENTROPY
        // sum p log2 p and sum log2 p; -sum q log2 q = log2 P - S / P and sum log2 q = sum log2 p - K log2 P follow per frame
#pragma unroll
        for (int c = 0; c < SW; ++c) q[c] = pv[c];
#else
#pragma unroll
        for (int c = 0; c < SW; ++c) q[c] = pv[c] * inv_p;                // make_spectrum_pmf, s_a_t.rs:284-290
#endif
#pragma unroll
        for (int c = 0; c < SW; ++c) l[c] = Mth<T>::log2_(q[c], ltab);
#pragma unroll
        for (int c = 0; c < SW; ++c) {
            const T d = d0 + (T)c;
            const T dd = d * d;
            const T w2 = dd * pv[c];
            d2s[c & 1] += w2;
            d3s[c & 1] = fma(w2, d, d3s[c & 1]);
            d4s[c & 1] = fma(w2, dd, d4s[c & 1]);
            ent[c & 1] = fma(q[c], l[c], ent[c & 1]);
            slq[c & 1] += l[c];
        }
    };
    {
        const T *const lp = lo_ptr(t);
        T dlo = (T)(16 * t) - kappa;
#pragma unroll 1
        for (int i0 = 0; i0 < 16; i0 += SW, dlo += (T)SW) {
            T pv[SW];
#pragma unroll
            for (int c = 0; c < SW; ++c) pv[c] = lp[2 * (i0 + c)];
            sweep4(pv, dlo);
        }
        // high chunk, ascending bins M - 16 t - 15 + u  <->  element i = 15 - u (i = 0 sits one slot apart)
        const int h0 = zslot(M - 16 * t);
        T dhi = (T)(M - 16 * t - 15) - kappa;
#pragma unroll 1
        for (int u0 = 0; u0 < 16; u0 += SW, dhi += (T)SW) {
            T pv[SW];
#pragma unroll
            for (int c = 0; c < SW; ++c) {
                const int i = 15 - u0 - c;
                pv[c] = bufT[2 * (i == 0 ? h0 : h0 - 1 - i) + hp];
            }
            sweep4(pv, dhi);
        }
        if (t == TPF - 1) {                             // bin M/2 sits between the low and the high chunks
            const T pw = acc.mid;
            const T d = (T)(M / 2) - kappa;
            const T dd = d * d, w2 = dd * pw;
            const T q = AUS_PSUM_ENTROPY ? pw : pw * inv_p;
            const T l = Mth<T>::log2_(q, ltab);
            d2s[0] += w2;
            d3s[0] = fma(w2, d, d3s[0]);
            d4s[0] = fma(w2, dd, d4s[0]);
            ent[0] = fma(q, l, ent[0]);
            slq[0] += l;
        }
    }
    T e_sum = ent[0] + ent[1], l_sum = slq[0] + slq[1];
#if AUS_COLD_OUTLINE
    if (badq) {
        const double2 el = cold_exact_log_sums<M, TPF>(buf, t, inv_p);
        e_sum = el.x;
        l_sum = el.y;
    }
#else
    if (badq) {                                          // warp-uniform, cold: library log2 (0 * -inf = NaN, sum = -inf)
        e_sum = 0;
        l_sum = 0;
#pragma unroll 1
        for (int u = 0; u < 33; ++u) {
            if (u == 32 && t != TPF - 1) break;
            const int b = u < 16 ? 16 * t + u : (u < 32 ? M - 16 * t - 15 + (u - 16) : M / 2);
            const T q = pow_of<T, M, TPF>(buf, b) * inv_p;
            const T l = log2(q);
            e_sum += q * l;
            l_sum += l;
        }
    }
#endif
    // ---- nothing reads the group buffer any more: the next frame's bulk copy starts now and overlaps the reduction, the
    //      record stores and the set-up of the next iteration (its latency, ~1 000 cycles, used to be waited for in full)
    fence_proxy_async();
    __syncwarp();
    if (next_src != nullptr && t == 0) {
        mbar_arrive_expect_tx(bar, (uint32_t)(C::N * sizeof(T)));
        tma_load_1d(buf, next_src, (uint32_t)(C::N * sizeof(T)), bar);
    }
    // ---- ten sums by recursive halving: lane t ends up with the total of value t >> 1
    const T r10[10] = {acc.s_pk, d2s[0] + d2s[1], d3s[0] + d3s[1], d4s[0] + d4s[1], e_sum, l_sum, b1p, b1pk, b2p, b2pk};
    T tot = warp_reduce_scatter<10>(r10, t);
#if AUS_PSUM_ENTROPY
    if (!badq) {                                         // lanes 8, 9 hold sum p log2 p, lanes 10, 11 sum log2 p
        const T l2P = log2(P);
        if (vi == 4) tot = fma(tot, inv_p, -l2P);        // sum q log2 q = S / P - log2 P
        if (vi == 5) tot = fma(-(T)K, l2P, tot);         // sum log2 q = sum log2 p - K log2 P
    }
#endif
    double *rec = p.raw + g * (unsigned long long)RAW_N;
    const bool wr = active && p.raw != nullptr;
    if (wr) {
        // even lanes 0..18 hold one reduced value each; odd lanes 9..19 fetch the six picked bins; 21, 23, 25 the rest
        if (!(t & 1) && vi < 10) rec[vi == 0 ? (int)RAW_SPK : vi <= 5 ? (int)RAW_D2 + vi - 1 : (int)RAW_B1P + vi - 6] = tot;
        if (wq < 4) rec[RAW_ROLL0 + wq] = (double)roll_res;
        if ((t & 1) && t >= 9 && t < 21) rec[RAW_PICK0 + (t - 9) / 2] = pick_val;
        if (t == 21) rec[RAW_SM] = S_m;
        if (t == 23) rec[RAW_SMK] = S_mk;
        if (t == 25) rec[RAW_P] = P;
    }
}

// One thread per frame: raw record -> the 14 features in the declaration order of analyzer.rs:10-23.
template <typename T>
__global__ void __launch_bounds__(256) finalize_kernel(const StftParams<T> p, int K) {
    asm volatile("griddepcontrol.wait;" ::: "memory");       // launched as a programmatic dependent of the feature kernel
    const unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.total_frames) return;
    const double *rec = p.raw + g * (unsigned long long)RAW_N;
    const double f0 = p.f0, n = (double)K;
    const double S_m = rec[RAW_SM], S_mk = rec[RAW_SMK], S_pk = rec[RAW_SPK], P = rec[RAW_P];
    const double var = f0 * f0 * (rec[RAW_D2] / P);
    double out[14];
    out[0] = f0 * S_mk / S_m;                                                    // centroid, s_a_t.rs:79-85
    out[1] = var;                                                                // :238-244
    out[2] = (f0 * f0 * f0 * (rec[RAW_D3] / P)) / (var * sqrt(var));             // :153-160  (var^1.5)
    out[3] = (f0 * f0 * f0 * f0 * (rec[RAW_D4] / P)) / (var * var);              // :122-129
    out[4] = -rec[RAW_ENT];                                                      // :93-99   -sum q log2 q
    out[5] = exp(0.5 * 0.69314718055994530942 * (rec[RAW_SLQ] / n + log2(P))) * n / S_m;   // :107-114
#pragma unroll
    for (int i = 0; i < 4; ++i) out[6 + i] = f0 * fmin(rec[RAW_ROLL0 + i], n - 1.0);       // :137-145
    out[10] = (n * S_pk - p.slope_sx * P) / p.slope_den;                         // :167-182
    const double pk0 = rec[RAW_PICK0], pk1 = rec[RAW_PICK0 + 1], pk2 = rec[RAW_PICK0 + 2];
    const double pk3 = rec[RAW_PICK0 + 3], pk4 = rec[RAW_PICK0 + 4], pk5 = rec[RAW_PICK0 + 5];
    const bool gap = p.pick[2] < p.pick[3];                // bin floor(m) sits between bands 1 and 2
    const double b1p = rec[RAW_B1P], b1pk = rec[RAW_B1PK], b2p = rec[RAW_B2P], b2pk = rec[RAW_B2PK];
    const double bp[3] = {b1p, b2p, b1p + b2p + (gap ? pk2 : 0.0)};
    const double bpk[3] = {b1pk, b2pk, b1pk + b2pk + (gap ? pk2 * (double)p.pick[2] : 0.0)};
    const double pal[3] = {pk0, pk2, pk0}, pah[3] = {pk1, pk3, pk1};
    const double pbl[3] = {pk2, pk4, pk4}, pbh[3] = {pk3, pk5, pk5};
#pragma unroll
    for (int b = 0; b < 3; ++b) {                                                // :190-230
        const double yl = pal[b] + p.band_tl[b] * (pah[b] - pal[b]);
        const double yu = pbl[b] + p.band_tu[b] * (pbh[b] - pbl[b]);
        const double sum_y = yl + bp[b] + yu;
        const double sum_xy = p.band_fl[b] * yl + f0 * bpk[b] + p.band_fu[b] * yu;
        out[11 + b] = (n * sum_xy - p.band_sx[b] * sum_y) / p.band_den[b];
    }
    T *dst = p.feat + g * 14;
#pragma unroll
    for (int i = 0; i < 14; ++i) dst[i] = (T)out[i];
}

// ---------------------------------------------------------------------------------------------------------
// Shared-memory tail of the forward kernel (every fft_size): on entry buf holds Z' in the Z layout (frame_fft()).
template <typename T, int LOG2N>
__device__ __forceinline__ bool frame_tail_s(Cx<T> *buf, double *scratch, const double2 *logtab, const StftParams<T> &p,
                                             unsigned long long g, bool active, int t, int grp, bool want_feat,
                                             const T *next_src = nullptr, uint64_t *bar = nullptr) {
    using C = Cfg<T, LOG2N>;
    constexpr int M = C::M, K = C::K, TPF = C::TPF;
    // ---- post-twiddle + power + first feature pass; |X|^2 replaces Z in place (each thread touches only its slots)
    FeatAcc<T> acc;
    Cx<T> *spec = (p.spec != nullptr && active) ? p.spec + g * (unsigned long long)K : nullptr;
    // four pairs (k, M - k) at a time: loads first, then eight independent post-twiddle / sqrt chains.
    // slots: zslot(16 t + i) = 17 t + i;  zslot(M - 16 t - i) = h0 - 1 - i for i >= 1 and h0 for i = 0.
    const int h0 = zslot(M - 16 * t);
    T *const bufT = reinterpret_cast<T *>(buf);
    const int hp = hpar<T>(t);
    // bin numbers as running floating-point values: converted from the thread index once, not per bin (I2F.F64 is slow)
    T klo_run = (T)(16 * t), khi_run = (T)(M - 16 * t);
    constexpr int PT = (sizeof(T) == 8 && TPF == 32) ? AUS_TAIL_PAIRS : 4;
#pragma unroll 1
    for (int i0 = 0; i0 < 16; i0 += PT, klo_run += (T)PT, khi_run -= (T)PT) {
        Cx<T> A[PT], B[PT], W[PT];
#pragma unroll
        for (int c = 0; c < PT; ++c) {
            const int i = i0 + c;
            A[c] = buf[17 * t + i];
            B[c] = buf[(c == 0 && i0 == 0) ? (t == 0 ? 0 : h0) : h0 - 1 - i];      // k = 0 pairs with itself
            W[c] = ldg_cx(&p.twNt[i * TPF + t]);
        }
        T plo[PT], phi[PT];
        if (!want_feat) {
            // K2 (spectra only): X replaces Z in place, then the whole row leaves in coalesced stores below
#pragma unroll
            for (int c = 0; c < PT; ++c) {
                const int i = i0 + c;
                Cx<T> Xlo, Xhi;
                post_twiddle(A[c], B[c], W[c], Xlo, Xhi);
                buf[17 * t + i] = Xlo;
                buf[(c == 0 && i0 == 0) ? h0 : h0 - 1 - i] = Xhi;
            }
            continue;
        }
#pragma unroll
        for (int c = 0; c < PT; ++c) {
            Cx<T> Xlo, Xhi;
            post_twiddle(A[c], B[c], W[c], Xlo, Xhi);
            plo[c] = Xlo.x * Xlo.x + Xlo.y * Xlo.y;
            phi[c] = Xhi.x * Xhi.x + Xhi.y * Xhi.y;
        }
#pragma unroll
        for (int c = 0; c < PT; ++c) {
            const int i = i0 + c;
            bufT[2 * (17 * t + i) + hp] = plo[c];
            bufT[2 * ((c == 0 && i0 == 0) ? h0 : h0 - 1 - i) + hp] = phi[c];
        }
        {
            T mlo[PT], mhi[PT];
#pragma unroll
            for (int c = 0; c < PT; ++c) {
                acc.pmin = min(acc.pmin, min(power_hi(plo[c]), power_hi(phi[c])));
                mlo[c] = Mth<T>::sqrt_(plo[c]);
                mhi[c] = Mth<T>::sqrt_(phi[c]);
            }
            const T klo = klo_run, khi = khi_run;
#pragma unroll
            for (int c = 0; c < PT; ++c) {
                acc.lo_sum += plo[c];
                acc.hi_sum += phi[c];
                acc.add(klo + (T)c, mlo[c], plo[c]);
                if constexpr (AUS_SPLIT_TAIL_ACC) acc.add2(khi - (T)c, mhi[c], phi[c]);
                else acc.add(khi - (T)c, mhi[c], phi[c]);
            }
        }
    }
    if (!want_feat) {
        if (t == TPF - 1) {
            const Cx<T> A = buf[zslot(M / 2)];
            buf[zslot(M / 2)] = Cx<T>{A.x + A.x, -(A.y + A.y)};     // X[M/2] = conj(Z[M/2])
        }
        group_sync<TPF, C::BLOCK>(grp);
        if (spec) {
#pragma unroll 4
            for (int k = t; k <= M; k += TPF) spec[k] = buf[zslot(k)];   // consecutive lanes -> consecutive bins
        }
    } else if (t == TPF - 1) {
        const Cx<T> A = buf[zslot(M / 2)];
        const Cx<T> X = Cx<T>{A.x + A.x, -(A.y + A.y)};             // X[M/2] = conj(Z[M/2])
        const T pm = X.x * X.x + X.y * X.y;
        bufT[2 * zslot(M / 2) + hp] = pm;
        acc.mid = pm;
        acc.pmin = min(acc.pmin, power_hi(pm));
        if (want_feat) acc.add((T)(M / 2), Mth<T>::sqrt_(pm), pm);
    }
    if (want_feat) acc.merge();
    if (want_feat) {
        // (a version generalised to half-warp groups -- fft 1024, two frames per warp -- measured 4 % slower than
        //  frame_features() there AND cost this one-warp instantiation its whole gain (1.22e8 vs 1.27e8 frames/s): the
        //  kernel is sensitive to such code-generation details, so only one-warp groups take this pass, in this exact form)
        if constexpr (AUS_FEAT32 && sizeof(T) == 8 && TPF == 32) {
            frame_features32<LOG2N>(buf, acc, logtab, p, g, active, t, next_src, bar);
            return next_src != nullptr;                 // the next frame's bulk copy is already in flight
        } else {
            frame_features<T, LOG2N>(buf, acc, scratch, logtab, p, g, active, t, grp, true);
        }
    }
    return false;
}

// register-resident variants, defined in aus_stft_r.cuh
template <typename T, int LOG2N, typename C, bool UNR>
__device__ __forceinline__ void frame_fft_keep(Cx<T> *buf, const StftParams<T> &p, int t, int grp, bool prewindowed,
                                               const Cx<T> *gsrc, Cx<T> (&v)[32], bool wpre);
template <typename T, int LOG2N, typename C, bool SHFL, bool FEAT>
__device__ __forceinline__ void frame_tail_r(Cx<T> (&v)[32], Cx<T> *buf, double *scratch, const double2 *logtab,
                                             const StftParams<T> &p, unsigned long long g, bool active, int t, int grp);

// ---------------------------------------------------------------------------------------------------------
// FEAT = false (register-resident variant only): spectra-only build -- no feature code, no log2 table; the smaller
// register footprint is worth +11 % on the spectra path (f64, fft 2048: 59 % -> 66 % of the HBM roofline).
template <typename T, int LOG2N, bool RR, bool FEAT>
__global__ void __maxnreg__((Cfg<T, LOG2N, RR>::MAXREG)) stft_kernel(const StftParams<T> p) {
    using C = Cfg<T, LOG2N, RR>;
    static_assert(RR || FEAT, "the spectra-only build exists for the register-resident variant");
    constexpr int N = C::N, TPF = C::TPF, FPB = C::FPB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int grp = threadIdx.x / TPF, t = threadIdx.x % TPF;
    Cx<T> *buf = reinterpret_cast<Cx<T> *>(smem_raw) + (size_t)grp * C::BUF;
    double *scratch = reinterpret_cast<double *>(smem_raw + C::off_scratch()) + (size_t)grp * C::SCR;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + C::off_bar()) + grp;
    double2 *logtab = reinterpret_cast<double2 *>(smem_raw + C::off_log());
    if constexpr (FEAT) {
        for (int i = threadIdx.x; i < LOG2_TAB * C::LOGREP; i += C::BLOCK) logtab[i] = p.log2tab[i / C::LOGREP];
    }
    if (t == 0) mbar_init(bar, 1);
    fence_mbar_init();
    __syncthreads();
    uint32_t parity = 0;
#if AUS_STAGGER_NS
    // All frame groups of an SM start in the same phase (bulk-copy wait, then loads, then butterflies ...) and, on a short
    // job, stay in lockstep: every pipe is hit by all of them at once and idles in between (a 3-round launch runs at half
    // the steady-state rate per round).  Spreading the starts over about one frame time breaks the lockstep.
    if (FEAT && sizeof(T) == 8 && TPF == 32) __nanosleep((unsigned)grp * AUS_STAGGER_NS);
#endif

    const unsigned long long stride = (unsigned long long)gridDim.x * FPB;
    const unsigned long long iters = (p.total_frames + stride - 1) / stride;
    const bool want_feat = FEAT && p.raw != nullptr;
    int f = 0;
    // frame g of the batch -> its samples: file (frames only move forward, so the search resumes at fq), start, length
    // (first frame, first frame of the next file, length, base pointer) of the file the last located frame lies in: the next
    // frame of this group is located from them without touching the descriptor arrays again when it is in the same file
    unsigned long long cur_first = 0, cur_next = 0, cur_len = 0;
    const T *cur_base = p.audio;
    auto locate = [&](unsigned long long gq, int &fq, const T *&srcq, int &Lq) -> bool {
        srcq = p.audio;
        Lq = 0;
        if (gq >= p.total_frames) return false;
        while (gq >= p.file_first[fq + 1]) ++fq;
        cur_first = p.file_first[fq];
        cur_next = p.file_first[fq + 1];
        cur_len = p.file_len[fq];
        cur_base = p.audio + p.file_off[fq];
        const unsigned long long s = (gq - cur_first) * (unsigned long long)p.hop;
        const unsigned long long rem = cur_len > s ? cur_len - s : 0ull;
        Lq = rem < (unsigned long long)N ? (int)rem : N;                // fft.rs:130-144
        srcq = cur_base + s;
        return true;
    };
    bool staged = false;                                                // this iteration's bulk copy was issued early (below)
#pragma unroll 1
    for (unsigned long long it = 0; it < iters; ++it) {
        const unsigned long long g = it * stride + (unsigned long long)blockIdx.x * FPB + grp;
        const T *src;
        int L;
        const bool active = locate(g, f, src, L);
        // A group made of whole warps that has run out of frames is done (frames only move forward, and nobody waits for it:
        // its barriers are its own).  Without this the last, partly filled round pushes all-zero frames through the cold
        // silent-frame paths -- on a one-file call that doubled the launch time.  Groups that share a warp keep going (their
        // shuffles and __syncwarp are warp-wide) and skip only the cold paths.
        if constexpr (TPF >= 32) {
            if (!active) break;
        }
#if AUS_L2_PREFETCH
        if (active && t == 0 && g + stride < p.file_first[f + 1]) {     // same file: its next frame of this group, if full
            const unsigned long long s2 = (g + stride - p.file_first[f]) * (unsigned long long)p.hop;
            const T *src2 = p.audio + p.file_off[f] + s2;
            if (s2 + N <= p.file_len[f] && (reinterpret_cast<uintptr_t>(src2) & 15) == 0)
                prefetch_l2(src2, (uint32_t)(N * sizeof(T)));
        }
#endif
        const bool prewindowed = L < N;                                 // short last frame, fft.rs:148-159
        const bool aligned = active && !prewindowed && p.tma_ok && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
        const bool direct = aligned && p.direct && C::R1 == 32;        // group-uniform
        const bool use_tma = aligned && !direct;
        // ---- stage the frame's samples into the group buffer (TMA bulk copy when 16-byte aligned and full length;
        //      otherwise a cooperative copy that also applies the short last-frame window).  Groups sharing a warp
        //      may take different paths, so no barrier sits inside a branch.
        if (use_tma) {
            if (t == 0 && !staged) {
                mbar_arrive_expect_tx(bar, (uint32_t)(N * sizeof(T)));
                tma_load_1d(buf, src, (uint32_t)(N * sizeof(T)), bar);
            }
        } else if (!direct) {
            T *bufr = reinterpret_cast<T *>(buf);
            if (prewindowed) {                                          // one call per frame, not one per sample (cold)
                stage_short_frame<T>(bufr, src, L, N, p.wtype, t, TPF);
            } else {
#pragma unroll 1
                for (int i = t; i < N; i += TPF) bufr[i] = i < L ? src[i] : (T)0;
            }
        }
        // the window values this thread multiplies in pass 1 do not depend on the frame: fetch them (L1 / L2) while the bulk
        // copy is still in flight, into the registers that receive the frame afterwards
        Cx<T> v[32];
        bool wpre = false;
        // (only where registers allow: the f64 spectra-only build, +3 %; with features the 128 live registers spill)
        if constexpr (C::R1 == 32 && AUS_WINDOW_PRELOAD && RR && !FEAT && sizeof(T) == 8) {
            if (!prewindowed && !direct) {                             // group-uniform
#pragma unroll
                for (int r = 0; r < 32; ++r) v[r] = ldg_cx(p.win2 + t + r * TPF);
                wpre = true;
            }
        }
        if (use_tma) {
            mbar_wait(bar, parity);
            parity ^= 1u;
        }
        group_sync<TPF, C::BLOCK>(grp);

        const T *half2 = nullptr;                                       // second half of the next frame, still to be copied
        if constexpr (RR) {
            // the last pass, the real-FFT split and the first feature pass stay in registers (aus_stft_r.cuh)
            frame_fft_keep<T, LOG2N, C, (AUS_UNROLL_PASSES != 0)>(buf, p, t, grp, prewindowed,
                                                                    direct ? reinterpret_cast<const Cx<T> *>(src) : nullptr, v, wpre);
            staged = false;
            if (!want_feat) {
                // spectra only: every thread has read its last-pass inputs, so the next frame's bulk copy starts now and
                // overlaps the butterflies, the real-FFT split and the stores.  One warp per group: the tail works from
                // registers and the whole frame goes.  Larger groups exchange partners through the upper half of the buffer,
                // so only the first half of the frame goes now and the second half follows after the tail (same mbarrier).
                const T *srcn;
                int L2, f2 = f;
                const bool act2 = it + 1 < iters && locate(g + stride, f2, srcn, L2);
                const bool tma2 = act2 && L2 == N && p.tma_ok && !(p.direct && C::R1 == 32) &&
                                  ((reinterpret_cast<uintptr_t>(srcn) & 15) == 0);      // = the next iteration's use_tma
                fence_proxy_async();                                    // pass-1 stores (generic proxy) before the async write
                group_sync<TPF, C::BLOCK>(grp);                         // (outside the branch: groups may share a warp)
                if (tma2) {                                             // group-uniform
                    constexpr uint32_t FIRST = (uint32_t)((TPF <= 32 ? N : N / 2) * sizeof(T));
                    if (t == 0) {
                        mbar_arrive_expect_tx(bar, (uint32_t)(N * sizeof(T)));
                        tma_load_1d(buf, srcn, FIRST, bar);
                    }
                    staged = true;
                    if constexpr (TPF > 32) half2 = srcn + N / 2;
                }
            }
            // partner exchange: by shuffles (f32; spectra-only builds, where the buffer now belongs to the next frame) or
            // through the group buffer (f64 with features: frees registers for the feature pass)
            if constexpr (!FEAT) {
                frame_tail_r<T, LOG2N, C, (TPF <= 32), false>(v, buf, scratch, logtab, p, g, active, t, grp);
            } else {
                frame_tail_r<T, LOG2N, C, (sizeof(T) == 4 && TPF <= 32), true>(v, buf, scratch, logtab, p, g, active, t, grp);
            }
        } else {
            frame_fft<T, LOG2N>(buf, p, t, grp, prewindowed, direct ? reinterpret_cast<const Cx<T> *>(src) : nullptr, v, wpre);
            const T *next_src = nullptr;
            if constexpr (AUS_FEAT32 && AUS_EARLY_TMA && sizeof(T) == 8 && TPF == 32) {
                // the next frame of this warp, located now: the feature pass issues its bulk copy the moment the buffer is free
                const unsigned long long g2 = g + stride;
                if (want_feat && active && g2 < cur_next && p.tma_ok && !p.direct) {       // same file: no descriptor loads
                    const unsigned long long s2 = (g2 - cur_first) * (unsigned long long)p.hop;
                    const T *srcn = cur_base + s2;
                    if (s2 + N <= cur_len && (reinterpret_cast<uintptr_t>(srcn) & 15) == 0) next_src = srcn;
                }
            }
            staged = frame_tail_s<T, LOG2N>(buf, scratch, logtab, p, g, active, t, grp, want_feat, next_src, bar);
        }
#ifndef AUS_EXP_NOFENCE
#define AUS_EXP_NOFENCE 0        // timing experiment only (violates the proxy ordering rules): no fence at the loop end
#endif
        if (!(AUS_EXP_NOFENCE || (!RR && staged))) {                    // (already fenced where the early copy was issued)
            fence_proxy_async();                                        // generic accesses to buf before the next TMA write
            group_sync<TPF, C::BLOCK>(grp);
        }
        if constexpr (RR && TPF > 32) {
            if (half2 != nullptr && t == 0)                             // second half of the early copy (see above)
                tma_load_1d(reinterpret_cast<T *>(buf) + N / 2, half2, (uint32_t)((N / 2) * sizeof(T)), bar);
        }
    }
}

}  // namespace aus
