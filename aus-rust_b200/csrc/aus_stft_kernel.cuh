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
// group is the TMA landing zone for the frame's samples and then the exchange space between passes; spectra
// never touch HBM unless the caller asks for them.
#pragma once
#include "aus_device.cuh"

namespace aus {

template <typename T, int LOG2N> struct Cfg {
    static constexpr int N = 1 << LOG2N;
    static constexpr int M = N / 2;
    static constexpr int K = M + 1;
    static constexpr int TPF = M / 32;                       // threads per frame, 1..256
    static constexpr bool THREE = (M > 1024);
    static constexpr int R1 = THREE ? 32 : TPF;              // first radix (TPF == 1: single radix-32 pass)
    static constexpr int R3 = THREE ? M / 1024 : 1;
    static constexpr int BLOCK = 256;
    static constexpr int FPB = BLOCK / TPF;                  // frame groups per CTA
    static constexpr int XS = (R1 >= 16) ? ilog2(R1) : 5;    // exchange-layout pad shift
    static constexpr int BUF = M + M / 16 + 8;               // complex elements per group buffer (max pad M/16)
    static constexpr int NW = TPF > 32 ? TPF / 32 : 1;
    static constexpr int SCR = 16 * NW + 8;                  // doubles of reduction scratch per group (+6 picks)
    static constexpr size_t smem_bytes() {
        return (size_t)FPB * BUF * sizeof(Cx<T>) + (size_t)FPB * SCR * sizeof(double) + (size_t)FPB * sizeof(uint64_t);
    }
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
    const Cx<T> *win2;                 // [M] (0.5 w[2n], 0.5 w[2n+1]) -- the 1/2 of the real-FFT split is folded in
    const Cx<T> *twM;                  // [M] exp(-2 pi i k / M)
    const Cx<T> *twN;                  // [M/2 + 1] (cos, sin)(2 pi k / N)
    T *feat;                           // [frames][14] or nullptr
    Cx<T> *spec;                       // [frames][K] or nullptr
    double f0;                         // sample_rate / fft_size   (fft_tools.rs:201)
    double slope_sx, slope_den;        // n(n-1)/2 and n*Sxx - Sx^2 with n = K (spectral_analysis_tools.rs:177-180)
    int pick[6];                       // bins a1-, a1+, b1- (= a2-), b1+ (= a2+), b2-, b2+ (…:198-201)
    int r_lo[2], r_hi[2];              // half-open bin ranges [ceil(m_fl), floor(m_fu)) of bands 1 and 2
    double band_fl[3], band_fu[3];     // Hz
    double band_tl[3], band_tu[3];     // fractional parts m_fl - floor(m_fl), m_fu - floor(m_fu)
    double band_sx[3], band_den[3];    // sum_x and n*sum_x2 - sum_x^2 (host, reference summation order)
    int bands_low;                     // all band bins <= M/2 (skip band tests on the mirrored half)
};

// window.rs:26-98 evaluated for one index (used only for the short last frame of a file, fft.rs:148-159)
__device__ __forceinline__ double window_value(int wtype, double i, double L) {
    const double PI = 3.14159265358979323846264338327950288;
    switch (wtype) {
    case 0: return (2.0 / (L - 1.0)) * ((L - 1.0) / 2.0 - fabs(i - (L - 1.0) / 2.0));
    case 1: return 0.42 - 0.5 * cos((2.0 * PI * i) / L) + 0.08 * cos((4.0 * PI * i) / L);
    case 2: return 0.5 - 0.5 * cos((2.0 * PI * i) / (L - 1.0));
    case 3: return 0.54 - 0.46 * cos((2.0 * PI * i) / (L - 1.0));
    default: return 1.0;
    }
}

template <int SHIFT> __device__ __forceinline__ int padi(int i) { return i + (i >> SHIFT); }

// ---------------------------------------------------------------------------------------------------------
// Forward packed FFT of the frame staged in buf (as N reals == M complex, unpadded).  On return buf holds
// Z'[k] = FFT_M(z')[k] in the "Z layout" (index k + (k >> 4)), where z' = 0.5 * window * z.
template <typename T, int LOG2N>
__device__ __forceinline__ void frame_fft(Cx<T> *buf, const StftParams<T> &p, int t, int grp, bool prewindowed) {
    using C = Cfg<T, LOG2N>;
    constexpr int M = C::M, TPF = C::TPF, R1 = C::R1, R3 = C::R3, XS = C::XS;
    Cx<T> v[32];

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
    } else {
        // ---- pass 1: radix R1, Ns = 1, B1 = 32 / R1 butterflies per thread, input = windowed samples
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
        group_sync<TPF, C::BLOCK>(grp);                       // every load done before buf is overwritten
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

        // ---- pass 2: radix 32, Ns = R1, one butterfly per thread (j = t)
        {
            const int j = t;
            const int jm = C::THREE ? (j & 31) * R3 : j;      // twiddle exponent step: W_M^(jm * r)
            static_for<0, 32>([&](auto r_) {
                constexpr int r = decltype(r_)::value;
                Cx<T> z = buf[padi<XS>(j + r * TPF)];
                if constexpr (r > 0) z = cmul(z, ldg_cx(&p.twM[jm * r]));
                v[r] = z;
            });
            group_sync<TPF, C::BLOCK>(grp);
            dft_dif<32>(v);
            if constexpr (!C::THREE) {
                static_for<0, 32>([&](auto r_) {
                    constexpr int r = decltype(r_)::value;
                    buf[padi<4>(j + r * TPF)] = v[bitrev(r, 5)];      // natural order k = j + r * R1
                });
            } else {
                const int base = (j >> 5) * 1024 + (j & 31);
                static_for<0, 32>([&](auto r_) {
                    constexpr int r = decltype(r_)::value;
                    buf[padi<XS>(base + r * 32)] = v[bitrev(r, 5)];
                });
            }
            group_sync<TPF, C::BLOCK>(grp);
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
                    if constexpr (r > 0) z = cmul(z, ldg_cx(&p.twM[j * r]));
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

// Per-thread running sums of the feature pass (arithmetic type T; cross-thread combination is always double).
template <typename T> struct FeatAcc {
    T s_m = 0, s_mk = 0, s_pk = 0, s_l = 0, s_pl = 0;
    T b1p = 0, b1pk = 0, b2p = 0, b2pk = 0;
    template <bool BANDS>
    __device__ __forceinline__ void add(int k, T m, T pw, const int (&rlo)[2], const int (&rhi)[2]) {
        const T kf = (T)k;
        s_m += m;
        s_mk += m * kf;
        const T pk = pw * kf;
        s_pk += pk;
        const T l = Mth<T>::log2_(pw);
        s_l += l;
        s_pl += pw * l;                               // 0 * -inf = NaN, like pmf * log2(pmf) in the reference
        if constexpr (BANDS) {
            if (k >= rlo[0] && k < rhi[0]) { b1p += pw; b1pk += pk; }
            if (k >= rlo[1] && k < rhi[1]) { b2p += pw; b2pk += pk; }
        }
    }
};

// ---------------------------------------------------------------------------------------------------------
// Feature pass over one frame.  Thread t owns the pairs k in [16 t, 16 t + 16): bins k ("low") and M - k ("high"),
// thread 0 additionally DC / Nyquist (pair 0), the last thread bin M/2.  plo/phi/pmid hold the power of those bins.
template <typename T, int LOG2N>
__device__ __forceinline__ void frame_features(const T (&plo)[16], const T (&phi)[16], T pmid, const FeatAcc<T> &acc,
                                               const T *picks_smem, double *scratch, const StftParams<T> &p,
                                               unsigned long long g, bool active, int t, int grp) {
    using C = Cfg<T, LOG2N>;
    constexpr int M = C::M, K = C::K, TPF = C::TPF;

    T lo_sum = 0, hi_sum = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) { lo_sum += plo[i]; hi_sum += phi[i]; }

    double r[12] = {(double)acc.s_m, (double)acc.s_mk, (double)acc.s_pk, (double)acc.s_l, (double)acc.s_pl,
                    (double)acc.b1p, (double)acc.b1pk, (double)acc.b2p, (double)acc.b2pk,
                    (double)lo_sum, (double)hi_sum, (double)pmid};
    group_allreduce<TPF, C::BLOCK, 12>(r, scratch, t, grp);
    const double S_m = r[0], S_mk = r[1], S_pk = r[2], S_l = r[3], S_pl = r[4];
    const double tot_lo = r[9], tot_hi = r[10], tot_mid = r[11];
    const double P = tot_lo + tot_mid + tot_hi;

    // prefix of the power spectrum in bin order: low chunks ascending t, bin M/2, high chunks descending t
    double inc_lo = (double)lo_sum, inc_hi = (double)hi_sum;
    group_inclusive_scan2<TPF, C::BLOCK>(inc_lo, inc_hi, scratch, t, grp);
    const T base_lo = (T)(inc_lo - (double)lo_sum);
    const T base_hi = (T)(tot_lo + tot_mid + (tot_hi - inc_hi));

    // second pass: central moments about the (magnitude-weighted) centroid, in units of bins, and the
    // roll-off counts: index = number of bins whose inclusive cumulative power is still below theta * P
    const T kappa = (T)(S_mk / S_m);
    const T th0 = (T)(0.5 * P), th1 = (T)(0.75 * P), th2 = (T)(0.9 * P), th3 = (T)(0.95 * P);
    T d2s = 0, d3s = 0, d4s = 0;
    int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    auto moment = [&](int k, T pw) {
        const T d = (T)k - kappa;
        const T dd = d * d;
        const T w2 = dd * pw;
        d2s += w2;
        d3s += w2 * d;
        d4s += w2 * dd;
    };
    auto count = [&](T cum) {
        c0 += (cum < th0) ? 1 : 0;
        c1 += (cum < th1) ? 1 : 0;
        c2 += (cum < th2) ? 1 : 0;
        c3 += (cum < th3) ? 1 : 0;
    };
    {
        T cum = base_lo;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            moment(16 * t + i, plo[i]);
            cum += plo[i];
            count(cum);
        }
        if (t == TPF - 1) {
            moment(M / 2, pmid);
            count((T)(tot_lo + (double)pmid));
        }
        cum = base_hi;
#pragma unroll
        for (int i = 15; i >= 0; --i) {
            moment(M - 16 * t - i, phi[i]);
            cum += phi[i];
            count(cum);
        }
    }
    double q[5] = {(double)d2s, (double)d3s, (double)d4s,
                   (double)(c0 + (c1 << 16)), (double)(c2 + (c3 << 16))};      // counts <= 8193 each: exact in f64
    group_allreduce<TPF, C::BLOCK, 5>(q, scratch, t, grp);

    if (t == 0 && active && p.feat != nullptr) {
        const double f0 = p.f0, n = (double)K;
        const double var = f0 * f0 * (q[0] / P);
        const unsigned int cnta = (unsigned int)q[3], cntb = (unsigned int)q[4];
        const int i50 = (int)(cnta & 0xffffu), i75 = (int)(cnta >> 16);
        const int i90 = (int)(cntb & 0xffffu), i95 = (int)(cntb >> 16);
        double out[14];
        out[0] = f0 * S_mk / S_m;                                                // centroid, s_a_t.rs:79-85
        out[1] = var;                                                            // :238-244
        out[2] = (f0 * f0 * f0 * (q[1] / P)) / (var * sqrt(var));                // :153-160  (var^1.5)
        out[3] = (f0 * f0 * f0 * f0 * (q[2] / P)) / (var * var);                 // :122-129
        out[4] = -(S_pl / P - log2(P));                                          // :93-99   -sum q log2 q
        out[5] = exp(0.5 * 0.69314718055994530942 * S_l / n) * n / S_m;          // :107-114 (ln m = ln2/2 * log2 p)
        out[6] = f0 * (double)min(i50, K - 1);                                   // :137-145
        out[7] = f0 * (double)min(i75, K - 1);
        out[8] = f0 * (double)min(i90, K - 1);
        out[9] = f0 * (double)min(i95, K - 1);
        out[10] = (n * S_pk - p.slope_sx * P) / p.slope_den;                     // :167-182
        const double pk0 = (double)picks_smem[0], pk1 = (double)picks_smem[1], pk2 = (double)picks_smem[2];
        const double pk3 = (double)picks_smem[3], pk4 = (double)picks_smem[4], pk5 = (double)picks_smem[5];
        const bool gap = p.pick[2] < p.pick[3];            // bin floor(m) sits between bands 1 and 2
        const double bp[3] = {r[5], r[7], r[5] + r[7] + (gap ? pk2 : 0.0)};
        const double bpk[3] = {r[6], r[8], r[6] + r[8] + (gap ? pk2 * (double)p.pick[2] : 0.0)};
        const double pal[3] = {pk0, pk2, pk0}, pah[3] = {pk1, pk3, pk1};
        const double pbl[3] = {pk2, pk4, pk4}, pbh[3] = {pk3, pk5, pk5};
#pragma unroll
        for (int b = 0; b < 3; ++b) {                                            // :190-230
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
}

// ---------------------------------------------------------------------------------------------------------
template <typename T, int LOG2N>
__global__ void __launch_bounds__(256, (sizeof(T) == 4 ? 2 : 1)) stft_kernel(const StftParams<T> p) {
    using C = Cfg<T, LOG2N>;
    constexpr int N = C::N, M = C::M, K = C::K, TPF = C::TPF, FPB = C::FPB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int grp = threadIdx.x / TPF, t = threadIdx.x % TPF;
    Cx<T> *buf = reinterpret_cast<Cx<T> *>(smem_raw) + (size_t)grp * C::BUF;
    double *scratch = reinterpret_cast<double *>(smem_raw + (size_t)FPB * C::BUF * sizeof(Cx<T>)) + (size_t)grp * C::SCR;
    T *picks = reinterpret_cast<T *>(scratch + 16 * C::NW);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + (size_t)FPB * C::BUF * sizeof(Cx<T>) +
                                                 (size_t)FPB * C::SCR * sizeof(double)) + grp;
    if (t == 0) mbar_init(bar, 1);
    fence_mbar_init();
    __syncthreads();
    uint32_t parity = 0;

    const unsigned long long stride = (unsigned long long)gridDim.x * FPB;
    const unsigned long long iters = (p.total_frames + stride - 1) / stride;
    int f = 0;
    for (unsigned long long it = 0; it < iters; ++it) {
        const unsigned long long g = it * stride + (unsigned long long)blockIdx.x * FPB + grp;
        const bool active = g < p.total_frames;
        const T *src = p.audio;
        int L = 0;
        if (active) {
            while (g >= p.file_first[f + 1]) ++f;                       // frames only move forward
            const unsigned long long s = (g - p.file_first[f]) * (unsigned long long)p.hop;
            const unsigned long long len = p.file_len[f];
            const unsigned long long rem = len > s ? len - s : 0ull;
            L = rem < (unsigned long long)N ? (int)rem : N;             // fft.rs:130-144
            src = p.audio + p.file_off[f] + s;
        }
        const bool prewindowed = L < N;                                 // short last frame, fft.rs:148-159
        const bool use_tma = active && !prewindowed && p.tma_ok && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
        // ---- stage the frame's samples into the group buffer (TMA bulk copy when 16-byte aligned and full length;
        //      otherwise a cooperative copy that also applies the short last-frame window).  Groups sharing a warp
        //      may take different paths, so no barrier sits inside a branch.
        if (use_tma) {
            if (t == 0) {
                mbar_arrive_expect_tx(bar, (uint32_t)(N * sizeof(T)));
                tma_load_1d(buf, src, (uint32_t)(N * sizeof(T)), bar);
            }
        } else {
            T *bufr = reinterpret_cast<T *>(buf);
            for (int i = t; i < N; i += TPF) {
                T x = (T)0;
                if (i < L) {
                    x = src[i];
                    if (prewindowed) x = (T)((double)x * (0.5 * window_value(p.wtype, (double)i, (double)L)));
                }
                bufr[i] = x;
            }
        }
        if (use_tma) {
            mbar_wait(bar, parity);
            parity ^= 1u;
        }
        group_sync<TPF, C::BLOCK>(grp);

        frame_fft<T, LOG2N>(buf, p, t, grp, prewindowed);

        // ---- post-twiddle + power + first feature pass
        T plo[16], phi[16], pmid = (T)0;
        FeatAcc<T> acc;
        const bool want_feat = p.feat != nullptr;
        Cx<T> *spec = (p.spec != nullptr && active) ? p.spec + g * (unsigned long long)K : nullptr;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int k = 16 * t + i;
            const Cx<T> A = buf[padi<4>(k)];
            const Cx<T> B = buf[padi<4>(k == 0 ? 0 : M - k)];
            const Cx<T> W = ldg_cx(&p.twN[k]);
            Cx<T> Xlo, Xhi;
            post_twiddle(A, B, W, Xlo, Xhi);
            if (spec) { spec[k] = Xlo; spec[M - k] = Xhi; }
            plo[i] = Xlo.x * Xlo.x + Xlo.y * Xlo.y;
            phi[i] = Xhi.x * Xhi.x + Xhi.y * Xhi.y;
            if (want_feat) {
                acc.template add<true>(k, Mth<T>::sqrt_(plo[i]), plo[i], p.r_lo, p.r_hi);
                if (p.bands_low) acc.template add<false>(M - k, Mth<T>::sqrt_(phi[i]), phi[i], p.r_lo, p.r_hi);
                else             acc.template add<true>(M - k, Mth<T>::sqrt_(phi[i]), phi[i], p.r_lo, p.r_hi);
            }
        }
        if (t == TPF - 1) {
            const Cx<T> A = buf[padi<4>(M / 2)];
            const Cx<T> X = Cx<T>{A.x + A.x, -(A.y + A.y)};             // X[M/2] = conj(Z[M/2])
            if (spec) spec[M / 2] = X;
            pmid = X.x * X.x + X.y * X.y;
            if (want_feat) {
                if (p.bands_low) acc.template add<false>(M / 2, Mth<T>::sqrt_(pmid), pmid, p.r_lo, p.r_hi);
                else             acc.template add<true>(M / 2, Mth<T>::sqrt_(pmid), pmid, p.r_lo, p.r_hi);
            }
        }
        if (want_feat) {
            for (int q = t; q < 6; q += TPF) {                          // the six interpolation end-point bins
                const int k = p.pick[q];
                const int kk = (k <= M / 2) ? k : M - k;
                Cx<T> Xlo, Xhi;
                post_twiddle(buf[padi<4>(kk)], buf[padi<4>(kk == 0 ? 0 : M - kk)], ldg_cx(&p.twN[kk]), Xlo, Xhi);
                const Cx<T> X = (k <= M / 2) ? Xlo : Xhi;
                picks[q] = X.x * X.x + X.y * X.y;
            }
        }
        fence_proxy_async();                                            // generic accesses to buf before the next TMA write
        group_sync<TPF, C::BLOCK>(grp);                                 // buf free for the next frame; picks visible
        if (want_feat) frame_features<T, LOG2N>(plo, phi, pmid, acc, picks, scratch, p, g, active, t, grp);
        group_sync<TPF, C::BLOCK>(grp);
    }
}

// K3: analyzer over caller-supplied magnitude rows (analysis/analyzer.rs:41-81, the loop at :99-104 / mp.rs:86-88)
template <typename T, int LOG2N>
__global__ void __launch_bounds__(256, 2) analyzer_kernel(const StftParams<T> p) {
    using C = Cfg<T, LOG2N>;
    constexpr int M = C::M, K = C::K, TPF = C::TPF, FPB = C::FPB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int grp = threadIdx.x / TPF, t = threadIdx.x % TPF;
    double *scratch = reinterpret_cast<double *>(smem_raw) + (size_t)grp * C::SCR;
    T *picks = reinterpret_cast<T *>(scratch + 16 * C::NW);
    const unsigned long long stride = (unsigned long long)gridDim.x * FPB;
    const unsigned long long iters = (p.total_frames + stride - 1) / stride;
    for (unsigned long long it = 0; it < iters; ++it) {
        const unsigned long long g = it * stride + (unsigned long long)blockIdx.x * FPB + grp;
        const bool active = g < p.total_frames;
        const T *row = p.mags + (active ? g : 0ull) * (unsigned long long)K;
        T plo[16], phi[16], pmid = (T)0;
        FeatAcc<T> acc;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int k = 16 * t + i;
            const T ml = active ? row[k] : (T)0, mh = active ? row[M - k] : (T)0;
            plo[i] = ml * ml;                                           // make_power_spectrum, s_a_t.rs:273-279
            phi[i] = mh * mh;
            acc.template add<true>(k, ml, plo[i], p.r_lo, p.r_hi);
            acc.template add<true>(M - k, mh, phi[i], p.r_lo, p.r_hi);
        }
        if (t == TPF - 1) {
            const T mm = active ? row[M / 2] : (T)0;
            pmid = mm * mm;
            acc.template add<true>(M / 2, mm, pmid, p.r_lo, p.r_hi);
        }
        for (int q = t; q < 6; q += TPF) {
            const T mm = active ? row[p.pick[q]] : (T)0;
            picks[q] = mm * mm;
        }
        group_sync<TPF, C::BLOCK>(grp);
        frame_features<T, LOG2N>(plo, phi, pmid, acc, picks, scratch, p, g, active, t, grp);
        group_sync<TPF, C::BLOCK>(grp);
    }
}

}  // namespace aus
