// aus_stft_pc.cuh -- producer/consumer build of the fused f64 STFT + 14-feature kernel (K1) for frame groups of one
// warp (fft_size 2048: M = 1024 = 32 threads x 32 points).  sm_100a.
//
// Why: ncu's per-instruction stall samples of the single-role kernel (profiles/r2_k1_pc_ncu_summary.md) show that a warp
// alone needs ~20 600 cycles per frame, 16 100 of them pure latency (a dependent DFMA costs ~25 cycles on B200; smem / L2
// round trips; the bulk copy of the next frame; one-lane searches), while three resident warps per scheduler are all the
// register file allows when every warp carries the 32 complex points (128 registers) of its frame.  The feature passes
// need few registers but long dependency chains, the FFT needs many registers and little latency hiding.  So the two are
// split over warps of different register budgets (setmaxnreg, as warp-specialised GEMMs do):
//   * 8 producer warps (2 warpgroups, 160 registers): stage a frame (TMA bulk copy), both radix-32 passes, real-FFT
//     split, |X|^2, |X| and the two magnitude sums; the power row goes to a compact 8.7 KB shared buffer;
//   * 8 consumer warps (2 warpgroups, 96 registers): everything analyzer.rs:41-81 does with the power row -- moments,
//     entropy, log-mean, sub-band sums, running sums and the four roll-off searches -- and the raw record.
// Producer w and consumer w hand the power buffer back and forth through two mbarriers (full / empty), so the FFT of
// frame i+1 overlaps the feature pass of frame i: 16 warps per SM instead of 12, and the latency-bound half of the work
// no longer holds 128 registers hostage.
//
// Arithmetic differences to the single-role kernel (all inside the 1e-9 contract, tests/test_gpu_parity.py):
//   * entropy and log-mean come from  S = sum p log2 p  and  L = sum log2 p  (no division by P per bin):
//     sum q log2 q = S / P - log2 P,  sum log2 q = L - K log2 P.  The difference carries an ABSOLUTE error of ~4e-15
//     bits, so the 1e-9 relative contract holds for entropies above ~4e-6 bits (a frame whose power sits in one bin to
//     better than -120 dB falls below that; the single-role kernels form q log2 q per bin and have no such floor) --
//     one more reason this build is opt-in;
//     (a degree-3 log2 polynomial, two FP64 operations per bin cheaper, was tried and dropped: its 1.5e-11 error per
//     bin breaks the contract for entropies below ~0.01 bits; tools/fit_log2_poly.py has the fit);
//   * sum p k comes from the first central moment: sum p k = sum p (k - kappa) + kappa P;
//   * roll-off: the running sum at chunk ends lives in registers (warp scan); a threshold's chunk is found with two
//     ballots, then one lane walks that chunk's 16 bins.  Knife edges fall back to the reference's own sequential
//     arithmetic exactly like the single-role kernel (s_a_t.rs:137-145).
#pragma once
#include "aus_stft_kernel.cuh"

#ifndef AUS_PC_REG_P
#define AUS_PC_REG_P 160
#endif
#ifndef AUS_PC_REG_C
#define AUS_PC_REG_C 96
#endif
#ifndef AUS_PC_PASSB_W
#define AUS_PC_PASSB_W 4
#endif

#ifndef AUS_PC_ABLATE
#define AUS_PC_ABLATE 0          // timing experiments only (results are garbage): 1 no consumer work, 2 no butterflies,
#endif                           // 4 no producer tail, 8 no shared-memory exchange in the FFT passes

#ifndef AUS_PC_TIMING
#define AUS_PC_TIMING 0          // development: per-phase clock64() totals in g_pc_timing (read with aus_debug_pc_timing)
#endif

namespace aus {

#if AUS_PC_TIMING
__device__ unsigned long long g_pc_timing[40];
#define PC_T(i) do { const long long now_ = clock64(); tacc[i] += now_ - tlast; tlast = now_; } while (0)
#else
#define PC_T(i) do { } while (0)
#endif

template <int LOG2N> struct CfgPC {
    static constexpr int N = 1 << LOG2N;
    static constexpr int M = N / 2;
    static constexpr int K = M + 1;
    static constexpr int TPF = M / 32;
    static_assert(TPF == 32, "the producer/consumer kernel is built for one warp per frame (fft_size 2048)");
    static constexpr int NPAIR = 8;                          // producer/consumer warp pairs per CTA
    static constexpr int BLOCK = 64 * NPAIR;                 // 512 threads: warpgroups 0-1 produce, 2-3 consume
    static constexpr int BUF = M + M / 16 + 2;               // complex slots of one FFT buffer (same layout as Cfg)
    // compact power buffer (doubles): low bins b < M/2 at b + b/16; high bins b > M/2 at HI0 + j + j/16 with j = M - b;
    // bin M/2 at MID; the pad slot that closes each 16-bin chunk (17 c + 16) carries the owning lane's magnitude sums
    static constexpr int HI0 = M / 2 + M / 32;               // 544
    static constexpr int MID = 2 * HI0;                      // 1088
    static constexpr int HDR = MID + 1;                      // one header word: smallest power key of the frame
    static constexpr int PWB = (HDR + 1 + 1) & ~1;           // 1090 doubles
    static constexpr int LOGREP = 8;
    __host__ __device__ static constexpr size_t off_pw() { return (size_t)NPAIR * BUF * sizeof(Cx<double>); }
    __host__ __device__ static constexpr size_t off_bar() { return off_pw() + (size_t)NPAIR * PWB * sizeof(double); }
    __host__ __device__ static constexpr size_t off_log() { return (off_bar() + 3 * NPAIR * sizeof(uint64_t) + 15) & ~(size_t)15; }
    __host__ __device__ static constexpr size_t smem_bytes() { return off_log() + (size_t)LOG2_TAB * LOGREP * sizeof(double2); }
};

template <int LOG2N> __device__ __forceinline__ int pc_slot(int b) {
    using C = CfgPC<LOG2N>;
    if (b < C::M / 2) return b + (b >> 4);
    if (b == C::M / 2) return C::MID;
    const int j = C::M - b;
    return C::HI0 + j + (j >> 4);
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int NREG> __device__ __forceinline__ void reg_alloc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(NREG));
}
template <int NREG> __device__ __forceinline__ void reg_dealloc() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(NREG));
}

// ---------------------------------------------------------------------------------------------------------
// Forward packed FFT of the frame staged in buf, M = 1024 = 32 x 32: the two radix-32 passes of frame_fft() (one shared
// butterfly body, rolled) with the pass-1 multipliers already in v[] -- the half-scaled window pairs, fetched by the
// caller while the frame's bulk copy was in flight, or (1, 1) for a frame that was windowed while it was staged.
template <int LOG2N>
__device__ __forceinline__ void pc_fft(Cx<double> *buf, const StftParams<double> &p, int t, Cx<double> (&v)[32],
                                       long long (&tacc)[24], long long &tlast) {
    constexpr int TPF = CfgPC<LOG2N>::TPF;
    const int j = t;
    int lbase = j, lstride = TPF;
    int sbase = 33 * j, sstride = 1;
#pragma unroll 1
    for (int pass = 0; pass < p.npass; ++pass) {
        if (pass == 0) {
#pragma unroll
            for (int r = 0; r < 32; ++r) v[r] = emul((AUS_PC_ABLATE & 8) ? v[(r + 1) & 31] : buf[lbase + r * lstride], v[r]);
        } else {
            if (!(AUS_PC_ABLATE & 8)) v[0] = buf[lbase];
            for_pass2_twiddles<double>(p.tw2 + (j & 31), 32, [&](auto r_, const Cx<double> &w) {
                constexpr int r = decltype(r_)::value;
                v[r] = cmul((AUS_PC_ABLATE & 8) ? v[r] : buf[lbase + r * lstride], w);
            });
        }
        __syncwarp();                                     // every load done before buf is overwritten
        PC_T(14 + 3 * pass);
        if (!(AUS_PC_ABLATE & 2)) dft_dif<32>(v);
        PC_T(15 + 3 * pass);
        if (!(AUS_PC_ABLATE & 8) || pass == 1) {
        static_for<0, 32>([&](auto r_) {
            constexpr int r = decltype(r_)::value;
            buf[sbase + r * sstride] = v[bitrev(r, 5)];
        });
        }
        __syncwarp();
        PC_T(16 + 3 * pass);
        lbase = j + (j >> 5);                             // pass 2: exchange layout in, Z layout out
        lstride = TPF + TPF / 32;
        sbase = j + (j >> 4);
        sstride = 34;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Producer tail: real-FFT split of the Z' left in buf by frame_fft() (Z layout), |X|^2 into the compact power buffer,
// |X| sums.  Thread t owns the pairs (k, M - k), k = 16 t + i.  Returns this lane's (sum m, sum m k) in the pad slots.
template <int LOG2N>
__device__ __forceinline__ void pc_producer_tail(Cx<double> *buf, double *pwb, const StftParams<double> &p, int t,
                                                 const double *next_src, uint64_t *bar_in) {
    using C = CfgPC<LOG2N>;
    constexpr int M = C::M, TPF = C::TPF;
    double s_m = 0, s_mk = 0;
    unsigned pmin = 0x7ff00000u;
    const int h0 = zslot(M - 16 * t);
    double klo_run = (double)(16 * t), khi_run = (double)(M - 16 * t);
    double *const plo_dst = pwb + 17 * t, *const phi_dst = pwb + C::HI0 + 17 * t;
#pragma unroll 1
    for (int i0 = 0; i0 < 16; i0 += 4, klo_run += 4.0, khi_run -= 4.0) {
        Cx<double> A[4], B[4], W[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int i = i0 + c;
            A[c] = buf[17 * t + i];
            B[c] = buf[(c == 0 && i0 == 0) ? (t == 0 ? 0 : h0) : h0 - 1 - i];      // k = 0 pairs with itself
            W[c] = ldg_cx(&p.twNt[i * TPF + t]);
        }
        double plo[4], phi[4], mlo[4], mhi[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            Cx<double> Xlo, Xhi;
            post_twiddle(A[c], B[c], W[c], Xlo, Xhi);
            plo[c] = Xlo.x * Xlo.x + Xlo.y * Xlo.y;
            phi[c] = Xhi.x * Xhi.x + Xhi.y * Xhi.y;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            plo_dst[i0 + c] = plo[c];
            phi_dst[i0 + c] = phi[c];
            pmin = min(pmin, min(power_hi(plo[c]), power_hi(phi[c])));
            mlo[c] = Mth<double>::sqrt_(plo[c]);
            mhi[c] = Mth<double>::sqrt_(phi[c]);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            s_m += mlo[c];
            s_m += mhi[c];
            s_mk = fma(mlo[c], klo_run + (double)c, s_mk);
            s_mk = fma(mhi[c], khi_run - (double)c, s_mk);
        }
    }
    Cx<double> Amid = Cx<double>{0.0, 0.0};
    if (t == TPF - 1) Amid = buf[zslot(M / 2)];
    // every Z' value has been read: the FFT buffer is free, so the next frame's bulk copy starts now and overlaps the
    // rest of this tail, the hand-over to the consumer and the window preload of the next frame
    fence_proxy_async();
    __syncwarp();
    if (next_src != nullptr && t == 0) {
        mbar_arrive_expect_tx(bar_in, (uint32_t)(C::N * sizeof(double)));
        tma_load_1d(buf, next_src, (uint32_t)(C::N * sizeof(double)), bar_in);
    }
    if (t == TPF - 1) {
        const Cx<double> X = Cx<double>{Amid.x + Amid.x, -(Amid.y + Amid.y)};   // X[M/2] = conj(Z[M/2])
        const double pm = X.x * X.x + X.y * X.y;
        pwb[C::MID] = pm;
        pmin = min(pmin, power_hi(pm));
        const double mm = Mth<double>::sqrt_(pm);
        s_m += mm;
        s_mk = fma(mm, (double)(M / 2), s_mk);
    }
    // zero / subnormal / non-finite power somewhere in the frame: the fast square root does not apply.  Redo this lane's
    // magnitude sums with the library sqrt (cold; reproduces the reference's NaN / zero rules).
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) pmin = min(pmin, __shfl_xor_sync(0xffffffffu, pmin, off));
    const bool bad = !power_is_normal<double>(pmin) || __any_sync(0xffffffffu, !(s_m < (double)INFINITY));
    if (bad) {
        __syncwarp();
        s_m = 0;
        s_mk = 0;
#pragma unroll 1
        for (int i = 0; i < 16; ++i) {
            const double ml = sqrt(plo_dst[i]), mh = sqrt(phi_dst[i]);
            s_m += ml;
            s_m += mh;
            s_mk += ml * (double)(16 * t + i);
            s_mk += mh * (double)(M - 16 * t - i);
        }
        if (t == TPF - 1) {
            const double mm = sqrt(pwb[C::MID]);
            s_m += mm;
            s_mk += mm * (double)(M / 2);
        }
    }
    plo_dst[16] = s_m;
    phi_dst[16] = s_mk;
    if (t == 0) reinterpret_cast<unsigned long long *>(pwb)[C::HDR] = bad ? 0ull : (unsigned long long)pmin;
}

// ---------------------------------------------------------------------------------------------------------
// Consumer: the 14 features' raw record from one power row (analyzer.rs:41-81).  Thread t owns the low chunk
// [16 t, 16 t + 16) and the high chunk (M - 16 t - 16, M - 16 t]; lane 31 also bin M/2.
template <int LOG2N>
__device__ __forceinline__ void pc_consumer(const double *pwb, const double2 *logtab, const StftParams<double> &p,
                                            unsigned long long g, int t, long long (&tacc)[24], long long &tlast) {
    using C = CfgPC<LOG2N>;
    constexpr int M = C::M, K = C::K, TPF = C::TPF;
    const double *const plo = pwb + 17 * t, *const phi = pwb + C::HI0 + 17 * t;
    const unsigned FULL = 0xffffffffu;

    // ---- magnitude sums of the producer -> centroid in bins
    double sm2[2] = {plo[16], phi[16]};
    const unsigned long long hdr = reinterpret_cast<const unsigned long long *>(pwb)[C::HDR];
    // sub-band sums for the three slope regions (s_a_t.rs:190-230): bins [r_lo, r_hi) of bands 1 and 2.  Eight bins per
    // trip (one trip at 44.1 kHz): loads first, independent products, tree sums -- no loop-carried chain inside a trip.
    double b1p = 0, b1pk = 0, b2p = 0, b2pk = 0;
#pragma unroll 1
    for (int b0 = p.r_lo[0] + t; b0 < p.r_hi[1]; b0 += 8 * TPF) {
        double pv[8], a1[8], k1[8], a2[8], k2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int b = b0 + j * TPF;
            pv[j] = b < p.r_hi[1] ? pwb[pc_slot<LOG2N>(b)] : 0.0;
        }
        const double bf0 = (double)b0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int b = b0 + j * TPF;
            const double pk = pv[j] * (bf0 + (double)(j * TPF));
            const bool in1 = b < p.r_hi[0], in2 = b >= p.r_lo[1];
            a1[j] = in1 ? pv[j] : 0.0; k1[j] = in1 ? pk : 0.0;
            a2[j] = in2 ? pv[j] : 0.0; k2[j] = in2 ? pk : 0.0;
        }
        auto tree8 = [](const double (&x)[8]) { return ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7])); };
        b1p += tree8(a1); b1pk += tree8(k1); b2p += tree8(a2); b2pk += tree8(k2);
    }
    PC_T(1);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        sm2[0] += __shfl_xor_sync(FULL, sm2[0], off);
        sm2[1] += __shfl_xor_sync(FULL, sm2[1], off);
    }
    const double S_m = sm2[0], S_mk = sm2[1];
    const double kappa = S_mk / S_m;

    // ---- one sweep over the thread's 33 bins: central moments about the centroid (bins), sum p log2 p, sum log2 p,
    //      chunk sums.  PW bins at a time: loads first, then PW independent chains.
    PC_T(2);
    const double2 *const ltab = logtab + (threadIdx.x & 7);
    double d1s = 0, d2s = 0, d3s = 0, d4s = 0, plp = 0, sl = 0, lo_sum = 0, hi_sum = 0;
    constexpr int PW = AUS_PC_PASSB_W;
    auto block = [&](const double *src, int step, double d0, double &csum) {
        double pv[PW], l[PW];
#pragma unroll
        for (int c = 0; c < PW; ++c) pv[c] = src[c * step];
#pragma unroll
        for (int c = 0; c < PW; ++c) l[c] = Mth<double>::log2_(pv[c], ltab);
#pragma unroll
        for (int c = 0; c < PW; ++c) {
            const double d = d0 + (double)c;
            const double dd = d * d;
            const double w2 = dd * pv[c];
            d1s = fma(pv[c], d, d1s);
            d2s += w2;
            d3s = fma(w2, d, d3s);
            d4s = fma(w2, dd, d4s);
            plp = fma(pv[c], l[c], plp);
            sl += l[c];
            csum += pv[c];
        }
    };
    {
        double dlo = (double)(16 * t) - kappa;
#pragma unroll 1
        for (int i0 = 0; i0 < 16; i0 += PW, dlo += (double)PW) block(plo + i0, 1, dlo, lo_sum);
        // high chunk in ascending bin order: bin M - 16 t - 15 + u sits at phi[15 - u]
        double dhi = (double)(M - 16 * t - 15) - kappa;
#pragma unroll 1
        for (int u0 = 0; u0 < 16; u0 += PW, dhi += (double)PW) block(phi + 15 - u0, -1, dhi, hi_sum);
    }
    double mid = 0.0;
    if (t == TPF - 1) {
        mid = pwb[C::MID];
        const double d = (double)(M / 2) - kappa;
        const double dd = d * d, w2 = dd * mid;
        const double l = Mth<double>::log2_(mid, ltab);
        d1s = fma(mid, d, d1s);
        d2s += w2;
        d3s = fma(w2, d, d3s);
        d4s = fma(w2, dd, d4s);
        plp = fma(mid, l, plp);
        sl += l;
    }
    PC_T(3);
    // ---- combine: the ten sums by recursive halving (lane t ends up with the total of value t >> 1); the two chunk sums by
    //      an inclusive scan, whose last lane also has the totals
    const double r10[10] = {d1s, d2s, d3s, d4s, plp, sl, b1p, b1pk, b2p, b2pk};
    const double tot = warp_reduce_scatter<10>(r10, t);
    double inc_lo = lo_sum, inc_hi = hi_sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double ua = __shfl_up_sync(FULL, inc_lo, off), ub = __shfl_up_sync(FULL, inc_hi, off);
        if (t >= off) { inc_lo += ua; inc_hi += ub; }
    }
    PC_T(4);
    const double tot_lo = __shfl_sync(FULL, inc_lo, 31), tot_hi = __shfl_sync(FULL, inc_hi, 31);
    const double tot_mid = __shfl_sync(FULL, mid, 31);
    const double P = tot_lo + tot_mid + tot_hi;
    const int vi = t >> 1;                               // the value whose total this lane holds
    double val = tot;
    // sum q log2 q and sum log2 q from the p-based sums; P must be a positive normal number for the table log2
    const bool slow = hdr == 0ull || !(P > 2.3e-308 && P < 1e300);
    if (!slow) {
        const double l2p = Mth<double>::log2_(P, ltab);
        if (vi == 0) val = tot + kappa * P;              // sum p k  =  sum p (k - kappa) + kappa P
        if (vi == 4) val = tot / P - l2p;
        if (vi == 5) val = tot - (double)K * l2p;
    } else {
        // zero / subnormal / non-finite power somewhere: the reference's own per-bin arithmetic with the library log2
        // (0 * -inf = NaN, sum = -inf: s_a_t.rs:93-114), and sum p k without the centroid (kappa is NaN for silence)
        const double inv_p = 1.0 / P;
        double e = 0, sq = 0, pk = 0;
#pragma unroll 1
        for (int u = 0; u < 33; ++u) {
            if (u == 32 && t != TPF - 1) break;
            const int b = u < 16 ? 16 * t + u : (u < 32 ? M - 16 * t - 15 + (u - 16) : M / 2);
            const double pv = pwb[pc_slot<LOG2N>(b)];
            const double q = pv * inv_p;
            const double l = log2(q);
            e += q * l;
            sq += l;
            pk += pv * (double)b;
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            e += __shfl_xor_sync(FULL, e, off);
            sq += __shfl_xor_sync(FULL, sq, off);
            pk += __shfl_xor_sync(FULL, pk, off);
        }
        if (vi == 0) val = pk;
        if (vi == 4) val = e;
        if (vi == 5) val = sq;
    }

    PC_T(5);
    // ---- roll-off (s_a_t.rs:137-145): index = number of bins whose inclusive running sum is below theta * P.
    //      Chunk ends in ascending bin order: low chunks t = 0..31, bin M/2, high chunks t = 31..0.  Lane 2 q + 1 settles
    //      threshold q: two ballots find the chunk, then a sum tree over the chunk's 16 bins finds the bin.
    const double base_lo = inc_lo - lo_sum;                              // running sum before this lane's low chunk
    const double base_hi = tot_lo + tot_mid + (tot_hi - inc_hi);         // ... before its high chunk
    const double end_hi = base_hi + hi_sum;
    double *rec = p.raw + g * (unsigned long long)RAW_N;
    const int wq = (t & 1) ? vi : 4;                                     // threshold this lane settles (4: none)
    int n_lo = 0, n_hi = 0;
    double my_th = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const double theta = q == 0 ? 0.5 : q == 1 ? 0.75 : q == 2 ? 0.9 : 0.95;
        const double th = theta * P;
        const int nl = __popc(__ballot_sync(FULL, inc_lo < th));        // NaN compares false: 0, walk chunk 0
        const int nh = __popc(__ballot_sync(FULL, end_hi < th));
        if (wq == q) { n_lo = nl; n_hi = nh; my_th = th; }
    }
    const int c_lo = n_lo < 32 ? n_lo : 31, c_hi = n_hi < 32 ? 31 - n_hi : 0;
    const double bl = __shfl_sync(FULL, base_lo, c_lo), bh = __shfl_sync(FULL, base_hi, c_hi);
    PC_T(6);
    if (wq < 4) {
        int res, first = 0, dir = 1;
        const double *src = nullptr;
        double base = 0.0, c_at = 0.0, c_prev = 0.0;
        bool amb = false;
        if (n_lo < 32) {                                                 // crossing inside low chunk n_lo
            first = 16 * c_lo; src = pwb + 17 * c_lo; base = bl; res = 0;
        } else if (!(tot_lo + tot_mid < my_th)) {                        // bin M/2
            res = M / 2; c_at = tot_lo + tot_mid; c_prev = tot_lo;
        } else if (n_hi < 32) {                                          // high chunk of lane c_hi, ascending bins
            first = M - 16 * c_hi - 15; dir = -1; src = pwb + C::HI0 + 17 * c_hi + 15; base = bh; res = 0;
        } else {                                                         // even the last bin stays below (rounding): K
            res = K; c_prev = P;
        }
        if (src != nullptr) {
            double x[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) x[e] = src[e * dir];
            double s2[8], s4[4], s8;
#pragma unroll
            for (int e = 0; e < 8; ++e) s2[e] = x[2 * e] + x[2 * e + 1];
#pragma unroll
            for (int e = 0; e < 4; ++e) s4[e] = s2[2 * e] + s2[2 * e + 1];
            s8 = s4[0] + s4[1];
            double run = base;
            int j = 0;
            double c = run + s8;                                         // bins 0..7 of the chunk
            if (c < my_th) { run = c; j = 8; }
            c = run + (j == 8 ? s4[2] : s4[0]);
            if (c < my_th) { run = c; j += 4; }
            c = run + (j == 12 ? s2[6] : j == 8 ? s2[4] : j == 4 ? s2[2] : s2[0]);
            if (c < my_th) { run = c; j += 2; }
            const double xa = src[j * dir], xb = src[(j + 1) * dir];
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
        if (amb) {
            const double theta = wq == 0 ? 0.5 : wq == 1 ? 0.75 : wq == 2 ? 0.9 : 0.95;
            double pseq = 0.0;                                           // power_spectrum.iter().sum(), analyzer.rs:44
            for (int i = 0; i < K; ++i) pseq += pwb[pc_slot<LOG2N>(i)];
            double csum = 0.0;
            int i = -1;
            while (csum < theta && i < K - 1) {
                ++i;
                csum += pwb[pc_slot<LOG2N>(i)] / pseq;
            }
            res = i;
        }
        rec[RAW_ROLL0 + wq] = (double)res;
    }
    // the record: even lanes 0..18 hold one reduced value each; odd lanes 9..19 fetch the six picked bins; 21, 23, 25 the rest
    if (!(t & 1) && vi < 10) rec[vi == 0 ? (int)RAW_SPK : vi <= 5 ? (int)RAW_D2 + vi - 1 : (int)RAW_B1P + vi - 6] = val;
    if ((t & 1) && t >= 9 && t < 21) rec[RAW_PICK0 + (t - 9) / 2] = pwb[pc_slot<LOG2N>(p.pick[(t - 9) / 2])];
    if (t == 21) rec[RAW_SM] = S_m;
    if (t == 23) rec[RAW_SMK] = S_mk;
    if (t == 25) rec[RAW_P] = P;
    __syncwarp();
    PC_T(7);
}

// ---------------------------------------------------------------------------------------------------------
template <int LOG2N>
__global__ void __launch_bounds__(CfgPC<LOG2N>::BLOCK, 1) stft_pc_kernel(const StftParams<double> p) {
    using C = CfgPC<LOG2N>;
    using T = double;
    constexpr int N = C::N, TPF = C::TPF, NP = C::NPAIR;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, t = threadIdx.x & 31;
    const bool producer = warp < NP;
    const int w = producer ? warp : warp - NP;
    Cx<T> *buf = reinterpret_cast<Cx<T> *>(smem_raw) + (size_t)w * C::BUF;
    double *pwb = reinterpret_cast<double *>(smem_raw + C::off_pw()) + (size_t)w * C::PWB;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + C::off_bar());
    uint64_t *bar_in = bars + w, *bar_full = bars + NP + w, *bar_empty = bars + 2 * NP + w;
    double2 *logtab = reinterpret_cast<double2 *>(smem_raw + C::off_log());
    for (int i = threadIdx.x; i < LOG2_TAB * C::LOGREP; i += C::BLOCK) logtab[i] = p.log2tab[i / C::LOGREP];
    if (threadIdx.x < NP) {
        mbar_init(bars + threadIdx.x, 1);
        mbar_init(bars + NP + threadIdx.x, 32);
        mbar_init(bars + 2 * NP + threadIdx.x, 32);
    }
    fence_mbar_init();
    __syncthreads();

    const unsigned long long stride = (unsigned long long)gridDim.x * NP;
    const unsigned long long iters = (p.total_frames + stride - 1) / stride;
    long long tacc[24] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = clock64();
    (void)tacc; (void)tlast;

    if (producer) {
        reg_alloc<AUS_PC_REG_P>();
        uint32_t par_in = 0, par_empty = 1;
        // frame g of the batch -> its file (frames only move forward), start, length.  The file's descriptor stays in
        // registers; the global arrays are read again only when a frame crosses into the next file.
        int f = -1;
        unsigned long long cur_first = 0, cur_next = 0, cur_len = 0;
        const T *cur_base = p.audio;
        auto locate = [&](unsigned long long gq, const T *&srcq, int &Lq) -> bool {
            srcq = p.audio;
            Lq = 0;
            if (gq >= p.total_frames) return false;
            if (gq >= cur_next) {
                do {
                    ++f;
                    cur_next = p.file_first[f + 1];
                } while (gq >= cur_next);
                cur_first = p.file_first[f];
                cur_len = p.file_len[f];
                cur_base = p.audio + p.file_off[f];
            }
            const unsigned long long s = (gq - cur_first) * (unsigned long long)p.hop;
            const unsigned long long rem = cur_len > s ? cur_len - s : 0ull;
            Lq = rem < (unsigned long long)N ? (int)rem : N;                    // fft.rs:130-144
            srcq = cur_base + s;
            return true;
        };
        auto tma_able = [&](bool act, const T *srcq, int Lq) {
            return act && Lq == N && p.tma_ok && ((reinterpret_cast<uintptr_t>(srcq) & 15) == 0);
        };
        const unsigned long long g0 = (unsigned long long)blockIdx.x * NP + w;
        const T *src;
        int L;
        bool active = locate(g0, src, L);
        bool use_tma = tma_able(active, src, L);
        if (use_tma && t == 0) {                                               // first frame: nothing to overlap with
            mbar_arrive_expect_tx(bar_in, (uint32_t)(N * sizeof(T)));
            tma_load_1d(buf, src, (uint32_t)(N * sizeof(T)), bar_in);
        }
#pragma unroll 1
        for (unsigned long long it = 0; it < iters; ++it) {
            const bool prewindowed = L < N;                                     // short last frame, fft.rs:148-159
            Cx<T> v[32];
            // pass-1 multipliers: the window values do not depend on the frame, so they are fetched (L1 / L2) before the
            // wait for the bulk copy, into the registers that receive the frame afterwards
            if (use_tma) {
#pragma unroll
                for (int r = 0; r < 32; ++r) v[r] = ldg_cx(p.win2 + t + r * TPF);
                PC_T(8);
                mbar_wait(bar_in, par_in);
                par_in ^= 1u;
                PC_T(9);
            } else {
                // unaligned start or the short last frame of a file (windowed here at its own length, fft.rs:148-159)
                T *bufr = reinterpret_cast<T *>(buf);
#pragma unroll 1
                for (int i = t; i < N; i += TPF) {
                    T x = (T)0;
                    if (i < L) {
                        x = src[i];
                        if (prewindowed) x = (T)((double)x * (0.5 * window_value(p.wtype, (double)i, (double)L)));
                    }
                    bufr[i] = x;
                }
                if (!prewindowed) {
#pragma unroll
                    for (int r = 0; r < 32; ++r) v[r] = ldg_cx(p.win2 + t + r * TPF);
                } else {
#pragma unroll
                    for (int r = 0; r < 32; ++r) v[r] = Cx<T>{1.0, 1.0};
                }
            }
            __syncwarp();
            pc_fft<LOG2N>(buf, p, t, v, tacc, tlast);
            PC_T(10);
            // the next frame of this warp: located now so that its descriptor loads overlap the tail
            const T *src_n;
            int L_n;
            const bool active_n = it + 1 < iters && locate(g0 + (it + 1) * stride, src_n, L_n);
            const bool tma_n = tma_able(active_n, src_n, L_n);
            PC_T(11);
            mbar_wait(bar_empty, par_empty);                                    // consumer is done with the power buffer
            par_empty ^= 1u;
            PC_T(12);
            if (!(AUS_PC_ABLATE & 4)) pc_producer_tail<LOG2N>(buf, pwb, p, t, tma_n ? src_n : nullptr, bar_in);
            else {
                fence_proxy_async();
                __syncwarp();
                if (tma_n && t == 0) {
                    mbar_arrive_expect_tx(bar_in, (uint32_t)(N * sizeof(T)));
                    tma_load_1d(buf, src_n, (uint32_t)(N * sizeof(T)), bar_in);
                }
            }
            mbar_arrive(bar_full);                                              // release: this lane's power / pad stores
            PC_T(13);
            src = src_n; L = L_n; active = active_n; use_tma = tma_n;
        }
    } else {
        reg_dealloc<AUS_PC_REG_C>();
        uint32_t par_full = 0;
#pragma unroll 1
        for (unsigned long long it = 0; it < iters; ++it) {
            const unsigned long long g = it * stride + (unsigned long long)blockIdx.x * NP + w;
            mbar_wait(bar_full, par_full);
            par_full ^= 1u;
            PC_T(0);
            if (g < p.total_frames && !(AUS_PC_ABLATE & 1)) pc_consumer<LOG2N>(pwb, logtab, p, g, t, tacc, tlast);
            mbar_arrive(bar_empty);
        }
    }
#if AUS_PC_TIMING
    if (t == 0 && blockIdx.x == 0 && (warp == 0 || warp == NP)) {
        for (int i = 0; i < 24; ++i) atomicAdd(&g_pc_timing[i], (unsigned long long)tacc[i]);
        atomicAdd(&g_pc_timing[32 + (producer ? 0 : 1)], iters);
    }
#endif
}

}  // namespace aus
