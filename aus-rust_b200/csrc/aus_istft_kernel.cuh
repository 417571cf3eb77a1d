// aus_istft_kernel.cuh -- inverse real STFT: Hermitian pre-twiddle + packed inverse FFT + window + deterministic
// overlap-add + window-norm divide + peak tracking (K4), and the -6 dBFS rescale pass (K5).  sm_100a.
//
// Replaces spectrum/fft.rs:204-268 (`irstft`), spectrum/fft_tools.rs:69-112 (`overlap_add`) and, as a batch of
// one frame without window/normalisation, spectrum/fft.rs:66-93 (`irfft`).
//
// A frame group (same TPF / 32-points-per-thread layout as the forward kernel) owns one *segment*: a contiguous
// frame range [a, b) of one file and the output samples [a*hop, b*hop) (to the end of the signal for the last
// segment).  It walks its frames in ascending order and accumulates each windowed chunk straight into the output
// array (L2-resident read-modify-write, no atomics), so every sample sums its covering frames in ascending frame
// order exactly like fft_tools.rs:102-105.  The ceil(N/hop)-1 frames before `a` that still overlap the segment are
// recomputed instead of communicated.  A sample is divided by its window norm the moment its last frame lands.
#pragma once
#include "aus_stft_kernel.cuh"
#include "aus_stft_r.cuh"

#ifndef AUS_K4_PRE_REG
#define AUS_K4_PRE_REG 0            // experiment: correct, -2 %
#endif

namespace aus {

template <typename T> struct IstftParams {
    const Cx<T> *spec;                       // [total_frames][K]
    T *out;                                  // all files' outputs, contiguous
    const unsigned long long *seg_file;      // [n_seg] file index
    const unsigned long long *seg_a, *seg_b; // [n_seg] frame range inside the file
    // A "file" here is one spectrogram or one frame-range shard of it (multi-GPU irstft, aus_irstft_multi): rows
    // [row0, ...) of the true spectrogram are present in `spec` starting at row file_first[file], and out[file_out_off]
    // is output sample samp0 of the true signal.  Frame indices (seg_a, seg_b, h) are always those of the true file.
    const unsigned long long *file_first;    // [n_files] row of `spec` that holds this file's frame row0
    const unsigned long long *file_frames;   // [n_files] frames of the true spectrogram (fft.rs:208)
    const unsigned long long *file_row0;     // [n_files] first frame present in `spec`
    const unsigned long long *file_samp0;    // [n_files] output sample held by out[file_out_off[file]]
    const unsigned long long *file_out_off;  // [n_files] first output sample of each file in `out`
    unsigned long long *file_max_bits;       // [n_files] bit pattern of max(out, 0.0) as f64 (fft.rs:255-258)
    unsigned long long n_seg;
    unsigned long long total_rows;           // rows of `spec` (f32: the array's last row is not over-read)
    int hop;
    unsigned hop_magic;                      // floor(2^32 / hop) + 1: x / hop == umulhi(x, hop_magic) for x < 2^15
    const T *rnorm;                          // [hop] reciprocal steady-state window norm (fast path)
    const Cx<T> *winf;                       // [M] (w[2n], w[2n+1]) full-scale window pairs (fast path)
    int fast_ok;                             // windowed, even hop < N: frames with every covering frame present take the fast path
    int halo;                                // ceil(N / hop) - 1
    int apply_window;                        // 0: irfft (plain inverse, no window / norm)
    int tma_ok;
    StftParams<T> fwd;                       // tables (win2, tw2, tw3, twN) for frame_fft
};

template <typename T, int LOG2N>
__global__ void __launch_bounds__(Cfg<T, LOG2N>::BLOCK, Cfg<T, LOG2N>::MINB) istft_kernel(const IstftParams<T> p) {
    using C = Cfg<T, LOG2N>;
    constexpr int N = C::N, M = C::M, K = C::K, TPF = C::TPF, FPB = C::FPB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int grp = threadIdx.x / TPF, t = threadIdx.x % TPF;
    Cx<T> *buf = reinterpret_cast<Cx<T> *>(smem_raw) + (size_t)grp * C::BUF;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + C::off_bar()) + grp;
    if (t == 0) mbar_init(bar, 1);
    fence_mbar_init();
    __syncthreads();
    uint32_t parity = 0;

    const unsigned long long stride = (unsigned long long)gridDim.x * FPB;
    const unsigned long long iters = (p.n_seg + stride - 1) / stride;
    const long long hop = p.hop;
    for (unsigned long long it = 0; it < iters; ++it) {
        const unsigned long long sidx = it * stride + (unsigned long long)blockIdx.x * FPB + grp;
        const bool seg_ok = sidx < p.n_seg;
        long long a = 0, b = 0, F = 0, h_begin = 0;
        unsigned long long gfirst = 0, ooff = 0;
        long long row0 = 0, samp0 = 0;
        int file = 0;
        if (seg_ok) {
            file = (int)p.seg_file[sidx];
            a = (long long)p.seg_a[sidx];
            b = (long long)p.seg_b[sidx];
            gfirst = p.file_first[file];
            F = (long long)p.file_frames[file];
            row0 = (long long)p.file_row0[file];
            samp0 = (long long)p.file_samp0[file];
            ooff = p.file_out_off[file];
            h_begin = a - p.halo > 0 ? a - p.halo : 0;
        }
        // groups that share a warp must run the same number of frame steps (barriers are warp-wide for TPF <= 32)
        long long steps = seg_ok ? b - h_begin : 0;
        if constexpr (TPF < 32) {
#pragma unroll
            for (int off = 16; off >= TPF; off >>= 1) {
                const long long o = __shfl_xor_sync(0xffffffffu, steps, off);
                steps = o > steps ? o : steps;
            }
        }
        const long long own_lo = a * hop;
        const long long out_len = (F > 0) ? (long long)N + hop * (F - 1) : 0;
        const long long own_hi = (b == F) ? out_len : b * hop;
        T vmax = (T)0;

        for (long long st = 0; st < steps; ++st) {
            const long long h = h_begin + st;
            const bool active = seg_ok && h < b;
            const Cx<T> *row = p.spec + (gfirst + (unsigned long long)(active ? h - row0 : 0)) * (unsigned long long)K;
#if AUS_L2_PREFETCH
            if (active && t == 0 && h + 1 < b && ((K * sizeof(Cx<T>)) % 16 == 0) && p.tma_ok)
                prefetch_l2(row + K, (uint32_t)(K * sizeof(Cx<T>)));        // next frame of this segment -> L2
#endif
            // ---- stage the K bins of this frame with a TMA bulk copy.  f64 rows (K x 16 bytes) are 16-byte sized and aligned.
            //      f32 rows are K x 8 = 8 (mod 16) bytes, so every second row starts 8 bytes into a 16-byte unit: the copy
            //      takes the enclosing 16-byte aligned range (one extra element before or after the row; never past the end of
            //      the array: its last row goes the cooperative way) and the frame then lives one element further (`sh`).
            constexpr bool ROW16 = (K * sizeof(Cx<T>)) % 16 == 0;
            const uintptr_t ra = reinterpret_cast<uintptr_t>(row);
            int sh = 0;
            bool use_tma = active && p.tma_ok;
            if constexpr (ROW16) {
                use_tma = use_tma && (ra & 15) == 0;
            } else {
                use_tma = use_tma && gfirst + (unsigned long long)(h - row0) + 1 < p.total_rows;
                if (use_tma) sh = (int)((ra & 15) >> 3);
            }
            constexpr uint32_t ROWB = (uint32_t)((K * sizeof(Cx<T>) + 15) & ~(size_t)15);     // f32: 8208 for either alignment
            if (use_tma) {
                if (t == 0) {
                    mbar_arrive_expect_tx(bar, ROWB);
                    tma_load_1d(buf, reinterpret_cast<const void *>(ra & ~(uintptr_t)15), ROWB, bar);
                }
            } else {
                for (int i = t; i < K; i += TPF) buf[i] = active ? row[i] : Cx<T>{(T)0, (T)0};
            }
            Cx<T> *const fb = buf + sh;                    // the frame (and everything derived from it below)
            // pre-twiddle factors of this thread's 16 pairs: exact table values (the first samples of a Hann-windowed
            // signal divide by a window norm of ~1e-15, so every rounding of the pre-twiddle is amplified there), fetched
            // while the row copy is in flight instead of inside the pre-twiddle loop
            Cx<T> Wp[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) Wp[i] = ldg_cx(&p.fwd.twN[t + TPF * i]);
            if (use_tma) {
                mbar_wait(bar, parity);
                parity ^= 1u;
            }
            group_sync<TPF, C::BLOCK>(grp);

            // ---- Hermitian pre-twiddle, in place: buf[k] <- conj(Zi[k]),
            //      Zi[k] = (Y[k] + conj Y[M-k]) + i W_N^{-k} (Y[k] - conj Y[M-k])   (fft.rs:221-229 folded in)
            //      The twiddles (cos, sin)(2 pi k / N) of this thread's pairs were loaded before the wait for the row (Wp).
            // Experiment, off: f32, fft 2048 (one warp per frame, two radix-32 passes), pre-twiddled values kept out of shared memory.
            // Thread t forms conj Zi[k] and conj Zi[M-k] for its k = t + 32 i; the first is its own pass-1 input r = i,
            // the second is input 31 - i of thread 32 - t (thread 0: its own input 32 - i), handed over by a shuffle --
            // 64 shared-memory accesses per thread less, 32 shuffles more: measured -2 % (the shuffles use the same data pipe).
            constexpr bool PRE_REG = AUS_K4_PRE_REG && sizeof(T) == 4 && TPF == 32 && C::R1 == 32 && !C::THREE;
            if constexpr (PRE_REG) {
                Cx<T> v[32];
                auto pretw = [&](Cx<T> Yk, Cx<T> Ym, const Cx<T> W, Cx<T> &A, Cx<T> &B) {
                    const T sx = Yk.x + Ym.x, sy = Yk.y - Ym.y;
                    const T dx = Yk.x - Ym.x, dy = Yk.y + Ym.y;
                    const T ux = W.x * dx - W.y * dy;
                    const T uy = W.x * dy + W.y * dx;
                    A = Cx<T>{sx - uy, -(sy + ux)};
                    B = Cx<T>{sx + uy, sy - ux};
                };
                static_for<0, 16>([&](auto i_) {
                    constexpr int i = decltype(i_)::value;
                    const int k = t + TPF * i;
                    Cx<T> Yk = fb[k], Ym = fb[M - k];
                    if (i == 0 && t == 0) { Yk.y = (T)0; Ym.y = (T)0; }        // DC / Nyquist: fft.rs:235
                    pretw(Yk, Ym, Wp[i], v[i], v[31 - i]);
                });
                Cx<T> mid = Cx<T>{(T)0, (T)0}, dummy;
                if (t == 0) pretw(fb[M / 2], fb[M / 2], Cx<T>{(T)0, (T)1}, mid, dummy);    // k = M/2 pairs with itself
                const int partner = (TPF - t) & (TPF - 1);
                static_for<0, 16>([&](auto c_) {                               // r = 31 .. 16: thread 0's sources stay intact
                    constexpr int r = 31 - decltype(c_)::value;
                    Cx<T> snd = v[r];
                    if (t == 0) snd = (r == 16) ? mid : v[r == 16 ? 16 : r - 1];
                    v[r] = shfl_cx<TPF>(snd, partner);
                });
                frame_fft<T, LOG2N>(fb, p.fwd, t, grp, /*prewindowed=*/true, nullptr, v, false, /*vin=*/true);
            } else {
            static_for<0, 17>([&](auto i_) {
                constexpr int i = decltype(i_)::value;
                const int k = t + TPF * i;                 // pairs (k, M-k), k = 0 .. M/2
                if (i < 16 || t == 0) {
                    Cx<T> Yk = fb[k], Ym = fb[M - k];
                    if (k == 0) { Yk.y = (T)0; Ym.y = (T)0; }   // imaginary parts of DC / Nyquist vanish under `.re` (fft.rs:235)
                    Cx<T> W;
                    if constexpr (i == 16) W = Cx<T>{(T)0, (T)1};               // only t == 0: k = M/2, angle pi/2
                    else W = Wp[i];
                    const T sx = Yk.x + Ym.x, sy = Yk.y - Ym.y;
                    const T dx = Yk.x - Ym.x, dy = Yk.y + Ym.y;
                    const T ux = W.x * dx - W.y * dy;      // U = W^{-k} D, W^{-k} = c + i s
                    const T uy = W.x * dy + W.y * dx;
                    fb[k] = Cx<T>{sx - uy, -(sy + ux)};
                    if (k != 0 && k != M / 2) fb[M - k] = Cx<T>{sx + uy, sy - ux};
                }
            });
            group_sync<TPF, C::BLOCK>(grp);

            {
                Cx<T> v[32];
                frame_fft<T, LOG2N>(fb, p.fwd, t, grp, /*prewindowed=*/true, nullptr, v, false);     // buf: conj(z) in the Z layout
            }
            }

            // ---- window + overlap-add into the output (ascending frame order), finalise completed samples.
            //      u = offset inside the frame.  Frames h - b (b >= 1) also cover the sample iff b <= (N-1-u)/hop, frames
            //      h + a iff a <= u/hop; both quotients come from a 32-bit multiply-high (p.hop_magic), not a division.
            // Fast path (almost every frame): windowed, even hop, every frame that can cover this frame's samples exists
            // (h >= halo) and a later frame follows.  Then both samples of a butterfly output share their class -- has a
            // partial sum from earlier frames iff u < N - hop, complete after this frame iff u < hop -- so each pair is one
            // 16-byte read-modify-write, the window norm is a table multiply, and nothing is computed per sample.
            const bool fastf = active && p.fast_ok && h >= p.halo && h < F - 1 &&
                               ((reinterpret_cast<uintptr_t>(p.out + ooff + (h * hop - samp0)) & (2 * sizeof(T) - 1)) == 0);
            if (fastf) {
                using V2 = typename std::conditional<sizeof(T) == 8, double2, float2>::type;
                T *outh = p.out + ooff + (h * hop - samp0);
                const long long rlo = own_lo - h * hop, rhi = own_hi - h * hop;     // owned range, relative (even)
                const int ulo = rlo < 0 ? 0 : (rlo > N ? N : (int)rlo), uhi = rhi < 0 ? 0 : (rhi > N ? N : (int)rhi);
                const int nprev = N - p.hop;
#pragma unroll 1
                for (int i0 = 0; i0 < 32; i0 += 8) {
                    V2 prev[8], rn[8];
                    Cx<T> z[8], w[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const int n = t + TPF * (i0 + c), u0 = 2 * n;
                        prev[c] = V2{(T)0, (T)0};
                        rn[c] = V2{(T)1, (T)1};
                        if (u0 >= ulo && u0 < uhi && u0 < nprev) prev[c] = __ldcg(reinterpret_cast<const V2 *>(outh + u0));
                        if (u0 < p.hop) rn[c] = __ldg(reinterpret_cast<const V2 *>(p.rnorm + u0));   // with the other loads
                        z[c] = fb[padi<4>(n)];
                        w[c] = ldg_cx(&p.winf[n]);
                    }
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const int u0 = 2 * (t + TPF * (i0 + c));
                        if (u0 < ulo || u0 >= uhi) continue;
                        T v0 = z[c].x * w[c].x + prev[c].x;                 // fft.rs:234-236 + overlap-add (ascending frames)
                        T v1 = -z[c].y * w[c].y + prev[c].y;
                        if (u0 < p.hop) {                                   // no later frame reaches it: fft.rs:250-258
                            v0 *= rn[c].x;
                            v1 *= rn[c].y;
                            if (v0 > vmax) vmax = v0;                       // f64::max skips NaN
                            if (v1 > vmax) vmax = v1;
                        }
                        *reinterpret_cast<V2 *>(outh + u0) = V2{v0, v1};
                    }
                }
            } else if (active) {
                T *outh = p.out + ooff + (h * hop - samp0);
                const long long rlo = own_lo - h * hop, rhi = own_hi - h * hop;
                const int ulo = rlo < 0 ? 0 : (rlo > N ? N : (int)rlo), uhi = rhi < 0 ? 0 : (rhi > N ? N : (int)rhi);
                const bool first_frame = (h == 0), last_frame = (h == F - 1);
#pragma unroll 1
                for (int i0 = 0; i0 < 32; i0 += 8) {
                    // eight butterfly outputs (16 samples) at a time: all partial sums are fetched from L2 first, so the
                    // read-modify-write latency is paid once per batch instead of once per sample
                    T prev[8][2];
                    Cx<T> z[8], w[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const int n = t + TPF * (i0 + c);
                        z[c] = fb[padi<4>(n)];
                        w[c] = ldg_cx(&p.fwd.win2[n]);
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int u = 2 * n + e;
                            const int back = p.hop == 1 ? N - 1 - u : (int)__umulhi((unsigned)(N - 1 - u), p.hop_magic);
                            const bool need = p.apply_window && u >= ulo && u < uhi && back > 0 && !first_frame;
                            prev[c][e] = need ? __ldcg(outh + u) : (T)0;                      // an earlier frame started it
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const int n = t + TPF * (i0 + c);
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int u = 2 * n + e;
                            if (u < ulo || u >= uhi) continue;
                            T val = (e == 0) ? z[c].x : -z[c].y;
                            if (!p.apply_window) { outh[u] = val; continue; }
                            val = val * ((T)2 * (e == 0 ? w[c].x : w[c].y)) + prev[c][e];      // fft.rs:234-236 + overlap-add
                            const int fwd = p.hop == 1 ? u : (int)__umulhi((unsigned)u, p.hop_magic);   // u / hop
                            if (fwd == 0 || last_frame) {                                      // no later frame reaches it
                                const int back = p.hop == 1 ? N - 1 - u : (int)__umulhi((unsigned)(N - 1 - u), p.hop_magic);
                                const long long nb = (long long)back < h ? back : h;          // frames h-nb .. h cover it
                                T norm = (T)0;                                                 // fft.rs:239-243, ascending frames
                                for (int b = (int)nb; b >= 0; --b) {
                                    const int idx = u + b * p.hop;
                                    const Cx<T> ww = ldg_cx(&p.fwd.win2[idx >> 1]);
                                    const T wv = (T)2 * ((idx & 1) ? ww.y : ww.x);
                                    norm += wv * wv;
                                }
                                val /= norm;                                                   // fft.rs:250-252
                                if (val > vmax) vmax = val;                                    // f64::max skips NaN
                            }
                            outh[u] = val;
                        }
                    }
                }
            }
            fence_proxy_async();
            group_sync<TPF, C::BLOCK>(grp);                 // frame h's writes visible before frame h+1 reads them
        }
        // ---- per-file peak (signed max, starts at 0.0): positive IEEE doubles order like unsigned integers
        if (p.apply_window) {
            constexpr int W = TPF < 32 ? TPF : 32;
#pragma unroll
            for (int off = W / 2; off >= 1; off >>= 1) {
                const T o = __shfl_xor_sync(0xffffffffu, vmax, off);
                vmax = o > vmax ? o : vmax;
            }
            if ((t & (W - 1)) == 0 && seg_ok && vmax > (T)0)
                atomicMax(p.file_max_bits + file, (unsigned long long)__double_as_longlong((double)vmax));
        }
    }
}

// K5: audio *= 10^(-6/20) / maxval per file (fft.rs:260-265).  grid.y = file.
template <typename T>
__global__ void peak_scale_kernel(T *out, const unsigned long long *file_out_off, const unsigned long long *file_out_len,
                                  const unsigned long long *file_max_bits, double level) {
    const int file = blockIdx.y;
    const double maxval = __longlong_as_double((long long)file_max_bits[file]);
    const T scaler = (T)(level / maxval);
    T *o = out + file_out_off[file];
    const unsigned long long n = file_out_len[file];
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        o[i] *= scaler;
}

}  // namespace aus
