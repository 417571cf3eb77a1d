// aus_stft_r.cuh -- register-resident tail of the forward kernel (the last FFT pass stays in registers), the
// second feature pass on the transposed power layout (any fft_size) and the analyzer kernel K3 built on it.  sm_100a.
//
// The L1/shared data pipe is the busiest unit of the fused kernel (ncu: 63-70 % of its wavefront peak at 18 % occupancy),
// so this path removes shared-memory round trips instead of arithmetic:
//   * the last radix-32 pass leaves Z[j + TPF s], s = 0..31, in the registers of thread j (no store);
//   * the partner Z[M - k] of the real-FFT split (fft.rs:171-172 via the packed transform) lives in thread
//     (TPF - j) mod TPF at s' = 31 - s, so half of the registers are swapped with it -- by warp shuffles, in place, or
//     through a staging area in the free half of the group buffer (f64 with features, where registers are short);
//   * X[k], X[M-k] are formed in registers and, when the caller wants spectra (K2), stored straight to HBM
//     (consecutive lanes -> consecutive bins, fully coalesced) -- no staging buffer, no second launch;
//   * |X|^2 goes to shared memory once, transposed (bin b at slot b + b/32), so that each thread then walks 32
//     consecutive bins for the order-dependent work (running sum -> roll-off) together with the second feature pass.
// Which launches use it is a measured table: use_rr() in aus_b200.cu.
#pragma once
#include "aus_stft_kernel.cuh"

namespace aus {

// Transposed power layout: bin b at T-slot b + b/32, so a thread's 32 consecutive bins [32 t, 32 t + 32) are 32
// consecutive slots from 33 t (conflict-free both ways); bin M takes the pad slot that closes the last chunk.
__device__ __forceinline__ int pslot_lt(int b) { return b + (b >> 5); }                          // b < M
template <int M> __device__ __forceinline__ int pslot(int b) { return b + (b >> 5) - (b == M ? 1 : 0); }
template <int M> struct RLay {
    static constexpr int CO = (M + M / 32 + 2 + 1) & ~1;      // T-element offset of the running-sum region
};

// (cos, sin)(2 pi s / 64), s = 0..15: the post-twiddle of bin k = t + TPF s is W_N^t x W_64^s for every fft_size here
// (N = 64 TPF), so a thread needs one table entry per frame and the rest are constant-bank operands (f64) / immediates.
#define AUS_W64_COS \
    1.0, 0.99518472667219688624483695310947992, 0.98078528040323044912618223613423904, \
    0.95694033573220886493579788698026997, 0.92387953251128675612818318939678829, 0.88192126434835502971275686366038835, \
    0.83146961230254523707878837761790576, 0.7730104533627369608109066097584698, 0.70710678118654752440084436210484904, \
    0.63439328416364549821517161322549337, 0.55557023301960222474283081394853287, 0.47139673682599764855638762590525438, \
    0.38268343236508977172845998403039887, 0.29028467725446236763619237581739527, 0.19509032201612826784828486847702224, \
    0.098017140329560601994195563888641846
#define AUS_W64_SIN \
    0.0, 0.098017140329560601994195563888641846, 0.19509032201612826784828486847702224, \
    0.29028467725446236763619237581739527, 0.38268343236508977172845998403039887, 0.47139673682599764855638762590525438, \
    0.55557023301960222474283081394853287, 0.63439328416364549821517161322549337, 0.70710678118654752440084436210484904, \
    0.7730104533627369608109066097584698, 0.83146961230254523707878837761790576, 0.88192126434835502971275686366038835, \
    0.92387953251128675612818318939678829, 0.95694033573220886493579788698026997, 0.98078528040323044912618223613423904, \
    0.99518472667219688624483695310947992
static __constant__ double c_w64_cos[16] = {AUS_W64_COS};
static __constant__ double c_w64_sin[16] = {AUS_W64_SIN};
__host__ __device__ constexpr double w64_cos(int i) {
    constexpr double t[16] = {AUS_W64_COS};
    return t[i];
}
__host__ __device__ constexpr double w64_sin(int i) {
    constexpr double t[16] = {AUS_W64_SIN};
    return t[i];
}
template <typename T, int S> __device__ __forceinline__ Cx<T> tw64() {
    if constexpr (sizeof(T) == 8) return Cx<T>{c_w64_cos[S], c_w64_sin[S]};
    else { constexpr T c = (T)w64_cos(S), sn = (T)w64_sin(S); return Cx<T>{c, sn}; }
}

template <int W> __device__ __forceinline__ Cx<double> shfl_cx(Cx<double> a, int src) {
    return Cx<double>{__shfl_sync(0xffffffffu, a.x, src, W), __shfl_sync(0xffffffffu, a.y, src, W)};
}
template <int W> __device__ __forceinline__ Cx<float> shfl_cx(Cx<float> a, int src) {
    return cx_from(__shfl_sync(0xffffffffu, cx_bits(a), src, W));
}
// ---------------------------------------------------------------------------------------------------------
// Forward packed FFT like frame_fft(), but the last pass stays in registers: on return natural output s of
// thread j's final radix-32 butterfly, Z'[j + TPF s], is v[bitrev(s, 5)], and every thread of the group has
// finished reading buf (which may be overwritten).
template <typename T, int LOG2N, typename C, bool UNR>
__device__ __forceinline__ void frame_fft_keep(Cx<T> *buf, const StftParams<T> &p, int t, int grp, bool prewindowed,
                                               const Cx<T> *gsrc, Cx<T> (&v)[32], bool wpre) {
    constexpr int M = C::M, TPF = C::TPF, R1 = C::R1, R3 = C::R3, XS = C::XS;

    if constexpr (M == 32) {                                  // one thread per frame: a single radix-32 pass
        static_for<0, 32>([&](auto r_) {
            constexpr int r = decltype(r_)::value;
            Cx<T> z = buf[r];
            if (!prewindowed) { const Cx<T> w = ldg_cx(&p.win2[r]); z.x *= w.x; z.y *= w.y; }
            v[r] = z;
        });
        dft_dif<32>(v);
    } else if constexpr (R1 == 32) {
        // same rolled two-trip loop as frame_fft(): one radix-32 butterfly body shared by both passes
        const int j = t;
        int lbase = j, lstride = TPF;
        const Cx<T> *tab = p.win2 + j;
        int tstride = TPF;
        int sbase = 33 * j, sstride = 1;
        auto one_pass = [&](int pass) {
            if (pass == 0) {
                if (gsrc != nullptr) {
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
            if (pass == 0 || C::THREE) {                      // two-pass sizes keep the second pass in registers
                static_for<0, 32>([&](auto r_) {
                    constexpr int r = decltype(r_)::value;
                    buf[sbase + r * sstride] = v[bitrev(r, 5)];
                });
                group_sync<TPF, C::BLOCK>(grp);
                lbase = j + (j >> 5);
                lstride = TPF + TPF / 32;
                tab = p.tw2 + (j & 31);
                tstride = 32;
                const int b = (j >> 5) * 1024 + (j & 31);     // (three-pass sizes) exchange layout: pad5(b + 32 r)
                sbase = b + (b >> 5);
                sstride = 33;
            }
        };
        if constexpr (UNR) {
            // two copies of the pass body with compile-time strides save the per-load address arithmetic of the shared
            // (rolled) body (AUS_UNROLL_PASSES)
            one_pass(0);
            one_pass(1);
        } else {
#pragma unroll 1
            for (int pass = 0; pass < p.npass; ++pass) one_pass(pass);
        }
        if constexpr (C::THREE) {
            // ---- pass 3 (M > 1024): radix R3, Ns = 1024, B3 = 32 / R3 butterflies per thread, kept in registers: butterfly
            //      b of thread t yields bins t + TPF b + 1024 r = t + TPF (b + B3 r), i.e. natural index s = b + B3 r
            constexpr int B3 = 32 / R3;
            Cx<T> w[32];
            static_for<0, B3>([&](auto b_) {
                constexpr int b = decltype(b_)::value;
                const int jj = t + TPF * b;
                static_for<0, R3>([&](auto r_) {
                    constexpr int r = decltype(r_)::value;
                    Cx<T> z = buf[padi<XS>(jj + r * 1024)];
                    if constexpr (r > 0) z = cmul(z, ldg_cx(&p.tw3[r * 1024 + jj]));
                    w[b * R3 + r] = z;
                });
            });
            group_sync<TPF, C::BLOCK>(grp);
            static_for<0, B3>([&](auto b_) {
                constexpr int b = decltype(b_)::value;
                Cx<T> u[R3];
                static_for<0, R3>([&](auto r_) { u[decltype(r_)::value] = w[b * R3 + decltype(r_)::value]; });
                dft_dif<R3>(u);
                static_for<0, R3>([&](auto r_) {
                    constexpr int r = decltype(r_)::value;
                    v[bitrev(b + B3 * r, 5)] = u[bitrev(r, ilog2(R3))];
                });
            });
        }
    } else {
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
        {
            const int j = t;
            v[0] = buf[padi<XS>(j)];
            for_pass2_twiddles<T>(p.tw2 + j, TPF, [&](auto r_, const Cx<T> &w) {
                constexpr int r = decltype(r_)::value;
                v[r] = cmul(buf[padi<XS>(j + r * TPF)], w);
            });
            group_sync<TPF, C::BLOCK>(grp);
            dft_dif<32>(v);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Second half of the feature computation on the transposed layout.  On entry the power of every bin b = 0..M sits at
// pw[pslot<M>(b)] (written by whichever thread produced it, not yet synchronised) and (s_m, s_mk, s_pk, p_part, pmin) are
// this thread's partial sums over the bins it produced.  Thread t then owns the 32 consecutive bins [32 t, 32 t + 32)
// (the last thread also bin M).  Writes the frame's raw record (see finalize_kernel).
template <typename T, int LOG2N, typename C, bool STORE_CUM>
__device__ __forceinline__ void frame_features_r(T *pw, T *cw, T s_m, T s_mk, T s_pk, T p_part, unsigned pmin, double *scratch,
                                                 const double2 *logtab, const StftParams<T> &p, unsigned long long g,
                                                 bool active, int t, int grp, bool mags_from_sqrt) {
    constexpr int M = C::M, K = C::K, TPF = C::TPF;
    const int c0 = 33 * t;                              // slot of bin 32 t
    const int cnt = (t == TPF - 1) ? 33 : 32;           // bins in this thread's chunk (bin M: slot c0 + 32)
    group_sync<TPF, C::BLOCK>(grp);                     // every power bin visible to the whole group

    // zero / subnormal / non-finite power somewhere: magnitudes again with the exact square root (cold)
    const bool bad_p = sizeof(T) == 8 && mags_from_sqrt && active && (!power_is_normal<T>(pmin) || !(p_part < (T)INFINITY));
    if (group_any<TPF, C::BLOCK>(bad_p, scratch, t, grp)) {
        s_m = 0;
        s_mk = 0;
#pragma unroll 1
        for (int j = 0; j < cnt; ++j) {
            const T m = Mth<T>::sqrt_exact(pw[c0 + j]);
            s_m += m;
            s_mk += m * (T)(32 * t + j);
        }
    }

    // sub-band sums for the three slope regions (s_a_t.rs:190-230): bins [r_lo, r_hi) of bands 1 and 2, all threads
    T b1p = 0, b1pk = 0, b2p = 0, b2pk = 0;
#pragma unroll 1
    for (int b = p.r_lo[0] + t; b < p.r_hi[1]; b += TPF) {
        const T pv = pw[pslot<M>(b)];
        const T pk = pv * (T)b;
        if (b < p.r_hi[0]) { b1p += pv; b1pk += pk; }
        if (b >= p.r_lo[1]) { b2p += pv; b2pk += pk; }
    }
    T r[8] = {s_m, s_mk, s_pk, b1p, b1pk, b2p, b2pk, p_part};       // combined in the kernel's arithmetic type
    group_allreduce_t<TPF, C::BLOCK, 8, T>(r, scratch, t, grp);
    const double S_m = (double)r[0], S_mk = (double)r[1], P = (double)r[7];

    // central moments about the (magnitude-weighted) centroid in bins; entropy and log-mean from q = p / P; the
    // chunk-local running sum goes to the second region (global running sum = chunk base + local, see below)
    const T kappa = (T)(S_mk / S_m);
    const T inv_p = (T)(1.0 / P);
    T d2s = 0, d3s = 0, d4s = 0, ent = 0, slq = 0;
    bool badq = false;                                   // (f32 uses lg2.approx, which handles every argument)
    if constexpr (sizeof(T) == 8) {
        const unsigned gpmin = group_min_u32<TPF, C::BLOCK>(pmin, scratch, t, grp);
        badq = active && (!power_is_normal<T>(gpmin) || !(P > 0.0 && P < (double)INFINITY) || abnormal((T)(power_from_hi<T>(gpmin) * inv_p)));
    }
    const double2 *const ltab = logtab + (C::LOGREP > 1 ? (threadIdx.x & 7) : 0);
    T cum = 0;
    {
        T dj = (T)(32 * t) - kappa;                     // bin - centroid of the block's first bin (kept as a running value:
#pragma unroll 1                                         //  no integer -> double conversion per bin)
        for (int j0 = 0; j0 < 32; j0 += 4, dj += (T)4) {
            T pv[4], q[4], l[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) pv[c] = pw[c0 + j0 + c];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const T d = dj + (T)c;
                const T dd = d * d;
                const T w2 = dd * pv[c];
                d2s += w2;
                d3s += w2 * d;
                d4s += w2 * dd;
                q[c] = pv[c] * inv_p;                    // make_spectrum_pmf, s_a_t.rs:284-290
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) l[c] = Mth<T>::log2_(q[c], ltab);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                ent += q[c] * l[c];
                slq += l[c];
                cum += pv[c];
                if constexpr (STORE_CUM) cw[c0 + j0 + c] = cum;
            }
        }
        if (t == TPF - 1) {                              // bin M closes the last chunk
            const T pv = pw[c0 + 32];
            const T d = (T)M - kappa;
            const T dd = d * d;
            const T w2 = dd * pv;
            d2s += w2;
            d3s += w2 * d;
            d4s += w2 * dd;
            const T q = pv * inv_p;
            const T l = Mth<T>::log2_(q, ltab);
            ent += q * l;
            slq += l;
            cum += pv;
            if constexpr (STORE_CUM) cw[c0 + 32] = cum;
        }
    }
    if (badq) {                                          // group-uniform, cold: library log2 (0 * -inf = NaN, sum = -inf)
        ent = 0;
        slq = 0;
#pragma unroll 1
        for (int j = 0; j < cnt; ++j) {
            const T q = pw[c0 + j] * inv_p;
            const T l = Mth<T>::log2_exact(q);
            ent += q * l;
            slq += l;
        }
    }
    T q[5] = {d2s, d3s, d4s, ent, slq};
    group_allreduce_t<TPF, C::BLOCK, 5, T>(q, scratch, t, grp);

    // running sum at the end of each chunk (bin order), published in the chunk's pad slot of the power region (the last
    // chunk's pad slot holds bin M; its end value is the total and never needed)
    T inc = cum;
    group_inclusive_scan1<TPF, C::BLOCK, T>(inc, scratch, t, grp);
    if (t < TPF - 1) pw[c0 + 32] = inc;
    group_sync<TPF, C::BLOCK>(grp);                      // chunk ends and local running sums visible to the searching threads

    // roll-off (s_a_t.rs:137-145): index = number of bins whose inclusive running sum is below theta * P; one thread per
    // threshold.  Level 1: binary search over the chunk ends for the chunk where the running sum first reaches the
    // threshold; level 2: inside that chunk (global running sum = end of the previous chunk + local).
    auto theta_of = [](int qq) { return qq == 0 ? 0.5 : qq == 1 ? 0.75 : qq == 2 ? 0.9 : 0.95; };
    auto level1 = [&](T th, int &c, T &bs) {
        int lo = 0, hi = TPF - 1;                          // first chunk in [0, TPF - 1] whose end is not below th
#pragma unroll 1
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (pw[33 * mid + 32] < th) lo = mid + 1; else hi = mid;
        }
        c = lo;
        bs = c > 0 ? pw[33 * c - 1] : (T)0;                // end of chunk c - 1 = pad slot 33 (c - 1) + 32
    };
    // Knife edge: the threshold sits within rounding distance of a running sum, where the reference's own arithmetic -- a
    // sequential sum of p[i] / P compared with theta (s_a_t.rs:137-145) -- decides.  Redo it exactly that way (one thread,
    // rare): identical result whenever the power values are identical.
    auto settle = [&](int qq, T th, int res, bool amb, T c_at, T c_prev) -> int {
        if constexpr (sizeof(T) == 8) {
            const double tol = 1e-12 * P;
            if (!amb && res < K)
                amb = fabs((double)c_at - (double)th) <= tol || (res >= 1 && fabs((double)th - (double)c_prev) <= tol);
            if (!amb && res == K) amb = fabs((double)th - (double)c_prev) <= tol;
            if (amb && !(P > 0.0)) {                       // silent (or NaN) frame: the reference's loop stops at bin 0
                res = 0;
            } else if (amb) {
                const double theta = theta_of(qq);
                double pseq = 0.0;                         // power_spectrum.iter().sum(): sequential (analyzer.rs:44)
                for (int i = 0; i < K; ++i) pseq += (double)pw[pslot<M>(i)];
                double csum = 0.0;
                int i = -1;
                while (csum < theta && i < K - 1) {
                    ++i;
                    csum += (double)pw[pslot<M>(i)] / pseq;
                }
                res = i;
            }
        }
        return res;
    };
    double *rec = p.raw + g * (unsigned long long)RAW_N;
    const bool wr = active && p.raw != nullptr;
    {
#pragma unroll 1
        for (int qq = t; qq < 4; qq += TPF) {
            const T th = (T)(theta_of(qq) * P);
            int c;
            T bs;
            level1(th, c, bs);
            const int n = (c == TPF - 1) ? 33 : 32;
            int res;
            bool amb = false;
            T c_at = (T)0, c_prev = bs;
            if constexpr (STORE_CUM) {
                const T *cl = cw + 33 * c;
                int lo = 0, hi = n;                        // first j in [0, n] with !(base + local[j] < th)
#pragma unroll 1
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (bs + cl[mid] < th) lo = mid + 1; else hi = mid;
                }
                res = 32 * c + lo;                         // = K when even bin M stays below (the reference then answers K - 1)
                if (lo < n) c_at = bs + cl[lo];
                else amb = c < TPF - 1;                    // inside the spectrum this happens only by rounding of the scan
                if (lo > 0) c_prev = bs + cl[lo - 1];      // running sum through bin res - 1
            } else {
                const T *pl = pw + 33 * c;                 // no stored running sums: walk the chunk's powers (a warp-
                                                           // cooperative scan of the chunk measured no faster)
                T run = bs;
                int j = n;                                 // first bin of the chunk whose running sum is not below th
                // eight bins per trip: the loads go out together, so the shared-memory latency is paid four times per
                // chunk instead of thirty-two (the sum itself stays sequential, as in the reference)
#pragma unroll 1
                for (int j0 = 0; j0 < 32 && j == n; j0 += 8) {
                    T x[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) x[e] = pl[j0 + e];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        if (j == n) {
                            const T nx = run + x[e];
                            if (!(nx < th)) { c_at = nx; j = j0 + e; }
                            else run = nx;
                        }
                    }
                }
                if (j == n && n == 33) {                   // bin M closes the last chunk
                    const T nx = run + pl[32];
                    if (!(nx < th)) { c_at = nx; j = 32; }
                    else run = nx;
                }
                res = 32 * c + j;
                if (j == n) amb = c < TPF - 1;
                c_prev = run;
            }
            res = settle(qq, th, res, amb, c_at, c_prev);
            if (wr) rec[RAW_ROLL0 + qq] = (double)res;
        }
    }

    if (wr) {
#pragma unroll 1
        for (int qq = t; qq < 6; qq += TPF) rec[RAW_PICK0 + qq] = (double)pw[pslot<M>(p.pick[qq])];
        if (t == 0) {
            rec[RAW_SM] = (double)r[0]; rec[RAW_SMK] = (double)r[1]; rec[RAW_SPK] = (double)r[2];
            rec[RAW_B1P] = (double)r[3]; rec[RAW_B1PK] = (double)r[4]; rec[RAW_B2P] = (double)r[5]; rec[RAW_B2PK] = (double)r[6];
            rec[RAW_P] = P;
            rec[RAW_D2] = (double)q[0]; rec[RAW_D3] = (double)q[1]; rec[RAW_D4] = (double)q[2]; rec[RAW_ENT] = (double)q[3];
            rec[RAW_SLQ] = (double)q[4];
        }
    }
}

#ifndef AUS_FEATR32
#define AUS_FEATR32 1            // one-warp frame groups: latency-trimmed frame_features_r (see frame_features_r32)
#endif

// warp_reduce_scatter for the kernel's arithmetic type (float shuffles are one instruction, double ones two)
template <int NV, typename R> __device__ __forceinline__ R warp_reduce_scatter_t(const R (&in)[NV], int t) {
    static_assert(NV <= 16, "at most 16 values");
    R v8[8];
    {
        const bool h = (t & 16) != 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const R lo = in[i], hi = i + 8 < NV ? in[i + 8] : (R)0;
            const R keep = h ? hi : lo, send = h ? lo : hi;
            v8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    R v4[4];
    {
        const bool h = (t & 8) != 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const R keep = h ? v8[i + 4] : v8[i], send = h ? v8[i] : v8[i + 4];
            v4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    R v2[2];
    {
        const bool h = (t & 4) != 0;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const R keep = h ? v4[i + 2] : v4[i], send = h ? v4[i] : v4[i + 2];
            v2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    const bool h = (t & 2) != 0;
    const R keep = h ? v2[1] : v2[0], send = h ? v2[0] : v2[1];
    R v1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
    return v1;
}

// ---------------------------------------------------------------------------------------------------------
// frame_features_r() for frame groups of exactly one warp (fft_size 2048; f64 and f32), with the dependency chains of
// the general version cut the way frame_features32() does it for the shared-memory tail (profiles/r2_k1_pc_experiment.md:
// a warp spends this half of the work waiting on dependent operations, not on a pipe):
//   * the running sum at the 32 chunk ends comes from a warp scan of the chunk sums and stays in registers; one ballot per
//     roll-off threshold finds the chunk and a sum tree over its 32 bins finds the bin (was: chunk ends published to
//     shared memory, a 5-step binary search of dependent loads, then 32 dependent adds or a 5-step search of stored sums);
//   * no per-bin running sum is stored even in the forward kernels (STORE_CUM builds);
//   * sub-band sums load eight bins at once and add them as trees; ten sums are combined by recursive halving.
// Same arithmetic per bin; knife-edge thresholds fall back to the reference's sequential arithmetic (f64) as before.
template <typename T, int LOG2N, typename C>
__device__ __forceinline__ void frame_features_r32(T *pw, T s_m, T s_mk, T s_pk, T p_part, unsigned pmin, const double2 *logtab,
                                                   const StftParams<T> &p, unsigned long long g, bool active, int t,
                                                   bool mags_from_sqrt) {
    constexpr int M = C::M, K = C::K, TPF = C::TPF;
    static_assert(TPF == 32, "one warp per frame");
    const unsigned FULL = 0xffffffffu;
    const int c0 = 33 * t;                              // slot of bin 32 t
    const int cnt = (t == TPF - 1) ? 33 : 32;           // bins in this thread's chunk (bin M: slot c0 + 32)
    __syncwarp();                                       // every power bin visible to the whole warp

    // zero / subnormal / non-finite power somewhere: magnitudes again with the exact square root (cold)
    const bool bad_p = sizeof(T) == 8 && mags_from_sqrt && (!power_is_normal<T>(pmin) || !(p_part < (T)INFINITY));
    if (__any_sync(FULL, bad_p)) {
        s_m = 0;
        s_mk = 0;
#pragma unroll 1
        for (int j = 0; j < cnt; ++j) {
            const T m = Mth<T>::sqrt_exact(pw[c0 + j]);
            s_m += m;
            s_mk += m * (T)(32 * t + j);
        }
    }
    // ---- sub-band sums (s_a_t.rs:190-230), eight bins per trip: loads first, independent products, tree sums
    T b1p = 0, b1pk = 0, b2p = 0, b2pk = 0;
#pragma unroll 1
    for (int b0 = p.r_lo[0] + t; b0 < p.r_hi[1]; b0 += 8 * TPF) {
        T pv[8], a1[8], k1[8], a2[8], k2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int b = b0 + j * TPF;
            pv[j] = b < p.r_hi[1] ? pw[pslot<M>(b)] : (T)0;
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
    // ---- totals every lane needs: sum m, sum m k, P (butterfly), smallest power (f64)
    T r3[3] = {s_m, s_mk, p_part};
    unsigned gpmin = pmin;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
        for (int i = 0; i < 3; ++i) r3[i] += __shfl_xor_sync(FULL, r3[i], off);
        if constexpr (sizeof(T) == 8) gpmin = min(gpmin, __shfl_xor_sync(FULL, gpmin, off));
    }
    const double S_m = (double)r3[0], S_mk = (double)r3[1], P = (double)r3[2];
    const T kappa = (T)(S_mk / S_m);
    const T inv_p = (T)(1.0 / P);
    bool badq = false;                                   // (f32 uses lg2.approx, which handles every argument)
    if constexpr (sizeof(T) == 8)
        badq = !power_is_normal<T>(gpmin) || !(P > 0.0 && P < (double)INFINITY) || abnormal((T)(power_from_hi<T>(gpmin) * inv_p));

    // ---- the per-bin sweep over this thread's 32 consecutive bins: central moments, q log2 q, log2 q, chunk sum
    const double2 *const ltab = logtab + (C::LOGREP > 1 ? (threadIdx.x & 7) : 0);
    T d2s[2] = {0, 0}, d3s[2] = {0, 0}, d4s[2] = {0, 0}, ent[2] = {0, 0}, slq[2] = {0, 0}, cs[2] = {0, 0};
    {
        T dj = (T)(32 * t) - kappa;
#pragma unroll 1
        for (int j0 = 0; j0 < 32; j0 += 4, dj += (T)4) {
            T pv[4], q[4], l[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) pv[c] = pw[c0 + j0 + c];
#pragma unroll
            for (int c = 0; c < 4; ++c) q[c] = pv[c] * inv_p;               // make_spectrum_pmf, s_a_t.rs:284-290
#pragma unroll
            for (int c = 0; c < 4; ++c) l[c] = Mth<T>::log2_(q[c], ltab);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const T d = dj + (T)c;
                const T dd = d * d;
                const T w2 = dd * pv[c];
                d2s[c & 1] += w2;
                d3s[c & 1] += w2 * d;
                d4s[c & 1] += w2 * dd;
                ent[c & 1] += q[c] * l[c];
                slq[c & 1] += l[c];
                cs[c & 1] += pv[c];
            }
        }
    }
    T p_last = (T)0;                                     // bin M closes the last chunk
    if (t == TPF - 1) {
        p_last = pw[c0 + 32];
        const T d = (T)M - kappa;
        const T dd = d * d;
        const T w2 = dd * p_last;
        const T q = p_last * inv_p;
        const T l = Mth<T>::log2_(q, ltab);
        d2s[0] += w2;
        d3s[0] += w2 * d;
        d4s[0] += w2 * dd;
        ent[0] += q * l;
        slq[0] += l;
    }
    T e_sum = ent[0] + ent[1], l_sum = slq[0] + slq[1];
    if (badq) {                                          // warp-uniform, cold: library log2 (0 * -inf = NaN, sum = -inf)
        e_sum = 0;
        l_sum = 0;
#pragma unroll 1
        for (int j = 0; j < cnt; ++j) {
            const T q = pw[c0 + j] * inv_p;
            const T l = Mth<T>::log2_exact(q);
            e_sum += q * l;
            l_sum += l;
        }
    }
    // ---- running sum at the chunk ends: inclusive scan of the chunk sums of bins [32 t, 32 t + 32)
    const T csum = cs[0] + cs[1];
    T inc = csum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const T u = __shfl_up_sync(FULL, inc, off);
        if (t >= off) inc += u;
    }
    const T base = inc - csum;                           // running sum before this lane's chunk
    // ---- roll-off (s_a_t.rs:137-145): lane 2 q + 1 settles threshold q
    const int vi = t >> 1;
    const int wq = (t & 1) ? vi : 4;
    int n_below = 0;
    T my_th = (T)0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const T th = (T)((q == 0 ? 0.5 : q == 1 ? 0.75 : q == 2 ? 0.9 : 0.95) * P);
        const int nb = __popc(__ballot_sync(FULL, inc < th));            // chunks that end below the threshold
        if (wq == q) { n_below = nb; my_th = th; }
    }
    const int cidx = n_below < 32 ? n_below : 31;
    const T cbase = __shfl_sync(FULL, base, cidx);
    const T tot32 = __shfl_sync(FULL, inc, 31);                          // running sum through bin M - 1
    const T pM = __shfl_sync(FULL, p_last, 31);
    int roll_res = 0;
    if (wq < 4) {
        int res;
        T c_at = (T)0, c_prev = (T)0;
        bool amb = false;
        if (n_below < 32) {
            const T *const pl = pw + 33 * cidx;
            T x[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) x[e] = pl[e];
            T s2[16], s4[8], s8[4], s16[2];
#pragma unroll
            for (int e = 0; e < 16; ++e) s2[e] = x[2 * e] + x[2 * e + 1];
#pragma unroll
            for (int e = 0; e < 8; ++e) s4[e] = s2[2 * e] + s2[2 * e + 1];
#pragma unroll
            for (int e = 0; e < 4; ++e) s8[e] = s4[2 * e] + s4[2 * e + 1];
            s16[0] = s8[0] + s8[1];
            s16[1] = s8[2] + s8[3];
            T run = cbase;
            int j = 0;
            T c = run + s16[0];
            if (c < my_th) { run = c; j = 16; }
            T sel = s8[0];
#pragma unroll
            for (int e = 1; e < 4; ++e) if (j == 8 * e) sel = s8[e];
            c = run + sel;
            if (c < my_th) { run = c; j += 8; }
            sel = s4[0];
#pragma unroll
            for (int e = 1; e < 8; ++e) if (j == 4 * e) sel = s4[e];
            c = run + sel;
            if (c < my_th) { run = c; j += 4; }
            sel = s2[0];
#pragma unroll
            for (int e = 1; e < 16; ++e) if (j == 2 * e) sel = s2[e];
            c = run + sel;
            if (c < my_th) { run = c; j += 2; }
            T xa = x[0], xb = x[1];
#pragma unroll
            for (int e = 2; e < 32; e += 2) if (j == e) { xa = x[e]; xb = x[e + 1]; }
            c = run + xa;
            if (c < my_th) { run = c; j += 1; c = run + xb; }
            res = 32 * cidx + j;
            if (c < my_th) {                                             // nothing in this chunk reaches the threshold
                if (cidx == TPF - 1) {                                   // ... the last chunk also holds bin M
                    run = c;
                    c = run + pM;
                    res = M;
                    if (c < my_th) { res = K; run = c; }                 // even bin M stays below (the reference answers K - 1)
                } else {
                    amb = true;                                          // scan and tree disagree by rounding
                }
            }
            c_at = c;
            c_prev = run;
        } else {                                                         // all 32 chunk ends below: bin M or beyond
            const T c = tot32 + pM;
            res = M;
            c_at = c;
            c_prev = tot32;
            if (c < my_th) { res = K; c_prev = c; }
        }
        if constexpr (sizeof(T) == 8) {
            // Knife edge: the threshold sits within rounding distance of a running sum, where the reference's own arithmetic
            // -- a sequential sum of p[i] / P compared with theta (s_a_t.rs:137-145) -- decides (rare)
            const double tol = 1e-12 * P;
            if (!amb && res < K) amb = fabs((double)c_at - (double)my_th) <= tol || (res >= 1 && fabs((double)my_th - (double)c_prev) <= tol);
            if (!amb && res == K) amb = fabs((double)my_th - (double)c_prev) <= tol;
            if (amb && !(P > 0.0)) {                                     // silent (or NaN) frame: the reference's loop stops at bin 0
                res = 0;
            } else if (amb) {
                const double theta = wq == 0 ? 0.5 : wq == 1 ? 0.75 : wq == 2 ? 0.9 : 0.95;
                double pseq = 0.0;                                       // power_spectrum.iter().sum(), analyzer.rs:44
                for (int i = 0; i < K; ++i) pseq += (double)pw[pslot<M>(i)];
                double csq = 0.0;
                int i = -1;
                while (csq < theta && i < K - 1) {
                    ++i;
                    csq += (double)pw[pslot<M>(i)] / pseq;
                }
                res = i;
            }
        }
        roll_res = res;
    }
    T pick_val = (T)0;
    if ((t & 1) && t >= 9 && t < 21) pick_val = pw[pslot<M>(p.pick[(t - 9) / 2])];
    // ---- ten sums by recursive halving: lane t ends up with the total of value t >> 1
    const T r10[10] = {s_pk, d2s[0] + d2s[1], d3s[0] + d3s[1], d4s[0] + d4s[1], e_sum, l_sum, b1p, b1pk, b2p, b2pk};
    const T tot = warp_reduce_scatter_t<10, T>(r10, t);
    if (active && p.raw != nullptr) {
        double *rec = p.raw + g * (unsigned long long)RAW_N;
        if (!(t & 1) && vi < 10) rec[vi == 0 ? (int)RAW_SPK : vi <= 5 ? (int)RAW_D2 + vi - 1 : (int)RAW_B1P + vi - 6] = (double)tot;
        if (wq < 4) rec[RAW_ROLL0 + wq] = (double)roll_res;
        if ((t & 1) && t >= 9 && t < 21) rec[RAW_PICK0 + (t - 9) / 2] = (double)pick_val;
        if (t == 21) rec[RAW_SM] = S_m;
        if (t == 23) rec[RAW_SMK] = S_mk;
        if (t == 25) rec[RAW_P] = P;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Real-FFT split + power + first feature pass straight from the registers left by frame_fft_keep().
template <typename T, int LOG2N, typename C, bool SHFL, bool FEAT>
__device__ __forceinline__ void frame_tail_r(Cx<T> (&v)[32], Cx<T> *buf, double *scratch, const double2 *logtab,
                                             const StftParams<T> &p, unsigned long long g, bool active, int t, int grp) {
    constexpr int M = C::M, K = C::K, TPF = C::TPF;
    const bool want_feat = FEAT && p.raw != nullptr;
    Cx<T> *spec = (p.spec != nullptr && active) ? p.spec + g * (unsigned long long)K : nullptr;
    T *const pw = reinterpret_cast<T *>(buf);

    // sums over the bins this thread produces: low bins k = t + TPF s and high bins M - k, s = 0..15.  With c_s = TPF s,
    //   sum m k = t S_lo + C_lo + (M - t) S_hi - C_hi,   S = sum m, C = sum m c_s   (c_s: compile-time constants)
    T s_lo = 0, s_hi = 0, c_lo = 0, c_hi = 0, p_lo = 0, p_hi = 0, pc_lo = 0, pc_hi = 0;
    unsigned pmin = 0x7ff00000u;
    // thread 0 also holds the self-paired bin M/2 = t + TPF 16: X[M/2] = conj(Z[M/2]) (x2: the halved window); it enters
    // the sums as one more "low" bin, before the register that holds it is reused by the exchange below
    if (t == 0) {
        const Cx<T> zmid = v[bitrev(16, 5)];
        const Cx<T> X = Cx<T>{zmid.x + zmid.x, -(zmid.y + zmid.y)};
        if (spec != nullptr) spec[M / 2] = X;
        if (want_feat) {
            const T pm = X.x * X.x + X.y * X.y;
            pw[pslot_lt(M / 2)] = pm;
            pmin = power_hi(pm);
            const T mm = Mth<T>::sqrt_(pm);
            s_lo = mm;
            c_lo = mm * (T)(M / 2);
            p_lo = pm;
            pc_lo = pm * (T)(M / 2);
        }
    }
    // ---- the partner Z'[M - k] of k = t + TPF s is natural output 31 - s of thread (TPF - t) mod TPF (thread 0 pairs with
    //      itself one step further: M - TPF s = TPF (32 - s), and its k = 0 with Z'[0] itself).
    //      SHFL: the upper half of the registers is swapped with the partner by warp shuffles, in place; nothing touches
    //      the group buffer, so a spectra-only launch can already stream the next frame into it.
    //      !SHFL: a [16][TPF] staging area in the free upper part of the group buffer -- row s' - 16, column j, plus Z'[0]
    //      in slot 16 TPF -- where the partner of pair s is slot (16 - s) TPF - t for every thread; twice the shared-memory
    //      traffic, but 64 registers are free right after the store, which the f64 feature pass needs (measured +4..12 %).
    const Cx<T> wt = ldg_cx(&p.twN[t]);                  // (cos, sin)(2 pi t / N)
    constexpr int PB = 4;                                // pairs per batch
    static_assert(!SHFL || TPF <= 32, "the shuffle exchange needs the whole frame group in one warp");
    Cx<T> *const xb = buf + (C::BUF - 16 * TPF - 2);
    const int partner = (TPF - t) & (TPF - 1);
    auto exchange = [&](auto s_) {                       // shuffle variant: in place, v[bitrev(31 - s)] <- Z'[M - k]
        constexpr int s = decltype(s_)::value;
        constexpr int dst = bitrev(31 - s, 5), alt = bitrev((32 - s) & 31, 5);
        Cx<T> snd = v[dst];
        if (t == 0) snd = v[alt];
        if constexpr (TPF > 1) v[dst] = shfl_cx<TPF>(snd, partner);
        else v[dst] = snd;
    };
    if constexpr (SHFL) {
        // walking s downwards keeps thread 0's sources intact until they are sent; pairs are consumed in the same order
        static_for<0, PB>([&](auto c_) { exchange(std::integral_constant<int, 15 - decltype(c_)::value>{}); });
    } else {
        static_for<16, 32>([&](auto s_) {
            constexpr int s = decltype(s_)::value;
            xb[(s - 16) * TPF + t] = v[bitrev(s, 5)];
        });
        if (t == 0) xb[16 * TPF] = v[0];
        group_sync<TPF, C::BLOCK>(grp);
    }
    static_for<0, 16 / PB>([&](auto b_) {
        constexpr int i0 = 16 - PB * (decltype(b_)::value + 1);           // batch = pairs i0 .. i0 + PB - 1, descending batches
        Cx<T> B[PB];
        if constexpr (SHFL) {
            if constexpr (i0 > 0)
                static_for<0, PB>([&](auto c_) { exchange(std::integral_constant<int, i0 - 1 - decltype(c_)::value>{}); });
            static_for<0, PB>([&](auto c_) { B[decltype(c_)::value] = v[bitrev(31 - i0 - decltype(c_)::value, 5)]; });
        } else {
#pragma unroll
            for (int c = 0; c < PB; ++c) B[c] = xb[(16 - i0 - c) * TPF - t];
        }
        Cx<T> Xlo[PB], Xhi[PB];
        static_for<0, PB>([&](auto c_) {
            constexpr int c = decltype(c_)::value, s = i0 + c;
            const Cx<T> W = s == 0 ? wt : cmul(wt, tw64<T, s>());       // (cos, sin)(2 pi (t + TPF s) / N)
            post_twiddle(v[bitrev(s, 5)], B[c], W, Xlo[c], Xhi[c]);
        });
        if (spec != nullptr) {
#pragma unroll
            for (int c = 0; c < PB; ++c) {
                const int k = t + TPF * (i0 + c);
                spec[k] = Xlo[c];                          // consecutive lanes -> consecutive bins
                spec[M - k] = Xhi[c];
            }
        }
        if (want_feat) {
            T plo[PB], phi[PB], mlo[PB], mhi[PB];
#pragma unroll
            for (int c = 0; c < PB; ++c) {
                const int k = t + TPF * (i0 + c);
                plo[c] = Xlo[c].x * Xlo[c].x + Xlo[c].y * Xlo[c].y;
                phi[c] = Xhi[c].x * Xhi[c].x + Xhi[c].y * Xhi[c].y;
                pw[pslot_lt(k)] = plo[c];
                pw[(i0 + c == 0) ? pslot<M>(M - k) : pslot_lt(M - k)] = phi[c];   // only k = 0 pairs with bin M
                if constexpr (sizeof(T) == 8) pmin = min(pmin, min(power_hi(plo[c]), power_hi(phi[c])));
                mlo[c] = Mth<T>::sqrt_(plo[c]);
                mhi[c] = Mth<T>::sqrt_(phi[c]);
            }
#pragma unroll
            for (int c = 0; c < PB; ++c) {
                const T cs = (T)(TPF * (i0 + c));
                s_lo += mlo[c];
                s_hi += mhi[c];
                c_lo += mlo[c] * cs;
                c_hi += mhi[c] * cs;
                p_lo += plo[c];
                p_hi += phi[c];
                pc_lo += plo[c] * cs;
                pc_hi += phi[c] * cs;
            }
        }
    });
    const T tf = (T)t, mt = (T)(M - t);
    const T s_m = s_lo + s_hi;
    const T s_mk = (tf * s_lo + c_lo) + (mt * s_hi - c_hi);
    const T s_pk = (tf * p_lo + pc_lo) + (mt * p_hi - pc_hi);
    const T p_part = p_lo + p_hi;
    if (want_feat) {
        if constexpr (AUS_FEATR32 && TPF == 32) frame_features_r32<T, LOG2N, C>(pw, s_m, s_mk, s_pk, p_part, pmin, logtab, p, g, active, t, true);
        else frame_features_r<T, LOG2N, C, true>(pw, pw + RLay<M>::CO, s_m, s_mk, s_pk, p_part, pmin, scratch, logtab, p, g, active, t, grp, true);
    }
}

// ---------------------------------------------------------------------------------------------------------
// K3: analyzer over caller-supplied magnitude rows (analysis/analyzer.rs:41-81, the loop at :99-104 / mp.rs:86-88).
// HBM-bound by design (K * sizeof(T) bytes in, 14 values out, ~22 flops per bin), so the layout is chosen for memory-level
// parallelism: no FFT buffer, only the transposed power row in shared memory (33 TPF + 1 slots per frame group), two
// (f64) or three (f32) 256-thread CTAs per SM, and every thread issues its 32 coalesced row loads back to back.
template <typename T, int LOG2N> struct CfgA {
    static constexpr int N = 1 << LOG2N;
    static constexpr int M = N / 2;
    static constexpr int K = M + 1;
    static constexpr int TPF = M / 32;
    static constexpr int BLOCK = 256;
    static constexpr int MINB = (sizeof(T) == 4) ? 3 : 2;
    static constexpr int FPB = BLOCK / TPF;
    static constexpr int PW = (33 * TPF + 1 + 1) & ~1;        // T slots per group: bins 0..M transposed (pslot), even
    static constexpr int NW = TPF > 32 ? TPF / 32 : 1;
    static constexpr int SCR = TPF > 32 ? 16 * NW : 2;
    static constexpr int LOGREP = sizeof(T) == 8 ? 8 : 1;
    static constexpr int MAXREG = (65536 / (BLOCK * MINB)) & ~7;
    __host__ __device__ static constexpr size_t off_scratch() { return ((size_t)FPB * PW * sizeof(T) + 15) & ~(size_t)15; }
    __host__ __device__ static constexpr size_t off_log() { return (off_scratch() + (size_t)FPB * SCR * sizeof(double) + 15) & ~(size_t)15; }
    __host__ __device__ static constexpr size_t smem_bytes() {
        return off_log() + (sizeof(T) == 8 ? (size_t)LOG2_TAB * LOGREP * sizeof(double2) : 0);
    }
};

template <typename T, int LOG2N>
__global__ void __maxnreg__((CfgA<T, LOG2N>::MAXREG)) analyzer_kernel(const StftParams<T> p) {
    using C = CfgA<T, LOG2N>;
    constexpr int M = C::M, K = C::K, TPF = C::TPF, FPB = C::FPB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int grp = threadIdx.x / TPF, t = threadIdx.x % TPF;
    T *const pw = reinterpret_cast<T *>(smem_raw) + (size_t)grp * C::PW;
    double *scratch = reinterpret_cast<double *>(smem_raw + C::off_scratch()) + (size_t)grp * C::SCR;
    double2 *logtab = reinterpret_cast<double2 *>(smem_raw + C::off_log());
    if constexpr (sizeof(T) == 8) {
        for (int i = threadIdx.x; i < LOG2_TAB * C::LOGREP; i += C::BLOCK) logtab[i] = p.log2tab[i / C::LOGREP];
    }
    __syncthreads();
    const unsigned long long stride = (unsigned long long)gridDim.x * FPB;
    const unsigned long long iters = (p.total_frames + stride - 1) / stride;
    const T tf = (T)t;
#pragma unroll 1
    for (unsigned long long it = 0; it < iters; ++it) {
        const unsigned long long g = it * stride + (unsigned long long)blockIdx.x * FPB + grp;
        const bool active = g < p.total_frames;
        if constexpr (TPF >= 32) {
            if (!active) break;                            // whole-warp groups past the last row are done (see stft_kernel)
        }
        const T *row = p.mags + (active ? g : 0ull) * (unsigned long long)K;
#if AUS_L2_PREFETCH
        if (t == 0 && g + stride + 1 < p.total_frames) {   // this group's next row -> L2 (16-byte aligned superset; the row after
            const uintptr_t a = reinterpret_cast<uintptr_t>(p.mags + (g + stride) * (unsigned long long)K);   // it exists, so in bounds)
            prefetch_l2(reinterpret_cast<const void *>(a & ~(uintptr_t)15), (uint32_t)((K * sizeof(T) + 16 + 15) & ~(size_t)15));
        }
#endif
        // thread t takes bins t + TPF s (consecutive lanes -> consecutive bins), thread 0 also bin M
        T mv[32];
#pragma unroll
        for (int s = 0; s < 32; ++s) mv[s] = active ? __ldg(row + t + TPF * s) : (T)0;
        const T m_last = (active && t == 0) ? __ldg(row + M) : (T)0;
        T s_m = 0, c_m = 0, s_p = 0, c_p = 0;              // sum m, sum m (TPF s), sum p, sum p (TPF s)
        unsigned pmin = 0x7ff00000u;
#pragma unroll
        for (int s = 0; s < 32; ++s) {
            const T m = mv[s];
            const T pv = m * m;                            // make_power_spectrum, s_a_t.rs:273-279
            pw[pslot_lt(t + TPF * s)] = pv;
            if constexpr (sizeof(T) == 8) pmin = min(pmin, power_hi(pv));
            const T cs = (T)(TPF * s);
            s_m += m;
            c_m += m * cs;
            s_p += pv;
            c_p += pv * cs;
        }
        T s_mk = tf * s_m + c_m, s_pk = tf * s_p + c_p;
        if (t == 0) {
            const T pv = m_last * m_last;
            pw[33 * TPF - 1] = pv;                         // pslot<M>(M)
            pmin = min(pmin, power_hi(pv));
            s_m += m_last;
            s_mk += m_last * (T)M;
            s_p += pv;
            s_pk += pv * (T)M;
        }
        if constexpr (AUS_FEATR32 && TPF == 32) frame_features_r32<T, LOG2N, C>(pw, s_m, s_mk, s_pk, s_p, pmin, logtab, p, g, active, t, false);
        else frame_features_r<T, LOG2N, C, false>(pw, nullptr, s_m, s_mk, s_pk, s_p, pmin, scratch, logtab, p, g, active, t, grp, false);
        group_sync<TPF, C::BLOCK>(grp);                    // the searches are done before the next row overwrites pw
    }
}

}  // namespace aus
