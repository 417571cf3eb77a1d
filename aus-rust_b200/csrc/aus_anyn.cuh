// aus_anyn.cuh -- fft sizes the radix-32 kernels do not cover: anything that is not a power of two in [64, 16384].
//
// rustfft plans any length (spectrum/fft.rs:120-121, :205-206), so `rstft(audio, 1000, 250, ..)` is legal upstream.  Here
// such a size is reformulated as a chirp-z (Bluestein) transform on top of the power-of-two kernels:
//
//     X[k] = c[k] * sum_n (x[n] c[n]) conj(c)[k - n],      c[n] = exp(-i pi n^2 / N)   (n^2 reduced mod 2N, exact)
//
// i.e. a circular convolution of length Lp = 2^ceil(log2(2N - 1)), done with two forward transforms of length Lp:
//     A = a * c (zero-padded)  ->  FFT_Lp  ->  * FFT_Lp(conj c, wrapped)  ->  conj  ->  FFT_Lp  ->  conj / Lp  ->  * c.
// A complex FFT_Lp is two launches' worth of K2 rows (real parts and imaginary parts as two real signals; FFT is linear),
// so the whole path is: pre kernel -> K2 (2F rows) -> multiply kernel -> K2 (2F rows) -> post kernel, in f64 whatever the
// caller's precision, chunked so that the workspace stays bounded.  Features of such frames come from the generic
// warp-per-row analyzer (spectral_generic_kernel, analyzer order); the inverse transform feeds conj(full spectrum) through
// the same machinery and overlap-adds with a gather kernel (one thread per output sample, covering frames in ascending
// order like fft_tools.rs:102-105).  This is a compatibility path (about 20x the work of a native size), not a tuned one.
#pragma once

namespace anyn {

static inline bool native_size(size_t N) {
    const int lg = log2_exact(N);
    return lg >= 6 && lg <= 14;
}
static inline size_t padded(size_t N) {
    size_t L = 64;
    while (L < 2 * N - 1) L <<= 1;
    return L;
}

struct Tab {
    size_t N = 0, Lp = 0;
    DevBuf chirp, bhat, win[5];
    bool has_win[5] = {false, false, false, false, false};
};

// in-place radix-2 FFT on the host (long double): only used once per fft_size for the Lp-point chirp filter
static void host_fft(std::vector<std::complex<long double>> &a) {
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    const long double PI = 3.14159265358979323846264338327950288L;
    for (size_t len = 2; len <= n; len <<= 1) {
        std::vector<std::complex<long double>> w(len / 2);
        for (size_t k = 0; k < len / 2; ++k) w[k] = std::polar(1.0L, -2.0L * PI * (long double)k / (long double)len);
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; ++k) {
                const std::complex<long double> u = a[i + k], v = a[i + k + len / 2] * w[k];
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
    }
}

static int get_tab(aus_ctx *c, size_t N, Tab **out) {
    auto it = c->anytabs.find(N);
    if (it != c->anytabs.end()) {
        *out = it->second;
        return AUS_OK;
    }
    Tab *t = new (std::nothrow) Tab();
    if (!t) return fail(AUS_ERR_ALLOC, "out of host memory");
    t->N = N;
    t->Lp = padded(N);
    const long double PI = 3.14159265358979323846264338327950288L;
    std::vector<double2> ch(N);
    std::vector<std::complex<long double>> b(t->Lp, std::complex<long double>(0.0L, 0.0L));
    for (size_t n = 0; n < N; ++n) {
        const unsigned long long r = ((unsigned long long)n * n) % (2ull * N);          // exact argument reduction
        const long double ang = PI * (long double)r / (long double)N;
        const long double cs = cosl(ang), sn = sinl(ang);
        ch[n] = double2{(double)cs, (double)(-sn)};                                      // c[n] = exp(-i pi n^2 / N)
        const std::complex<long double> bc(cs, sn);                                      // conj c[n]
        b[n] = bc;
        if (n) b[t->Lp - n] = bc;
    }
    host_fft(b);
    std::vector<double2> bh(t->Lp);
    for (size_t k = 0; k < t->Lp; ++k) bh[k] = double2{(double)b[k].real(), (double)b[k].imag()};
    int rc = t->chirp.ensure(N * sizeof(double2));
    if (rc == AUS_OK) rc = t->bhat.ensure(t->Lp * sizeof(double2));
    if (rc != AUS_OK) { delete t; return rc; }
    cudaError_t e = cudaMemcpyAsync(t->chirp.p, ch.data(), N * sizeof(double2), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(t->bhat.p, bh.data(), t->Lp * sizeof(double2), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);                          // the host vectors are temporaries
    if (e != cudaSuccess) {
        t->chirp.release();
        t->bhat.release();
        delete t;
        return fail(AUS_ERR_CUDA, std::string("chirp tables: ") + cudaGetErrorString(e));
    }
    c->anytabs[N] = t;
    *out = t;
    return AUS_OK;
}
static int need_window(aus_ctx *c, Tab *t, int wtype) {
    if (t->has_win[wtype]) return AUS_OK;
    std::vector<double> w(t->N);
    AUS_TRY(aus_generate_window(wtype, t->N, w.data()));
    AUS_TRY(t->win[wtype].ensure(t->N * sizeof(double)));
    CUDA_TRY(cudaMemcpyAsync(t->win[wtype].p, w.data(), t->N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    t->has_win[wtype] = true;
    return AUS_OK;
}
static void release_all(aus_ctx *c) {
    for (auto &kv : c->anytabs) {
        kv.second->chirp.release();
        kv.second->bhat.release();
        for (DevBuf &w : kv.second->win) w.release();
        delete kv.second;
    }
    c->anytabs.clear();
}

// ---- kernels.  One CTA per frame (row); A holds two real rows of Lp doubles per frame: real parts, then imaginary parts.
// frame g of the batch: windowed like fft.rs:128-168 (a short last frame gets the window of its own length), times the chirp
template <typename T>
__global__ void __launch_bounds__(256) pre_audio_kernel(const T *audio, const unsigned long long *file_off, const unsigned long long *file_len,
                                                       const unsigned long long *file_first, int n_files, unsigned long long g0, int N,
                                                       int Lp, int hop, int wtype, const double *win, const double2 *chirp, double *A) {
    __shared__ unsigned long long s_start;
    __shared__ int s_L;
    const unsigned long long g = g0 + blockIdx.x;
    if (threadIdx.x == 0) {
        int lo = 0, hi = n_files - 1;                          // last file whose first frame is <= g
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (file_first[mid] <= g) lo = mid; else hi = mid - 1;
        }
        const unsigned long long s = (g - file_first[lo]) * (unsigned long long)hop, len = file_len[lo];
        const unsigned long long rem = len > s ? len - s : 0ull;
        s_L = rem < (unsigned long long)N ? (int)rem : N;       // fft.rs:130-144
        s_start = file_off[lo] + s;
    }
    __syncthreads();
    const int L = s_L;
    const T *src = audio + s_start;
    double *ar = A + (unsigned long long)blockIdx.x * 2ull * (unsigned long long)Lp, *ai = ar + Lp;
    for (int n = threadIdx.x; n < Lp; n += blockDim.x) {
        double re = 0.0, im = 0.0;
        if (n < L) {
            const double w = (L == N) ? win[n] : aus::window_value(wtype, (double)n, (double)L);
            const double a = (double)src[n] * w;
            const double2 ch = chirp[n];
            re = a * ch.x;
            im = a * ch.y;
        }
        ar[n] = re;
        ai[n] = im;
    }
}
// row r of a spectrogram: conj of the full spectrum rebuilt as fft.rs:221-229 does (N even), times the chirp
template <typename T>
__global__ void __launch_bounds__(256) pre_spec_kernel(const aus::Cx<T> *spec, unsigned long long r0, int N, int Lp, const double2 *chirp,
                                                      double *A) {
    const int K = N / 2 + 1;
    const aus::Cx<T> *row = spec + (r0 + blockIdx.x) * (unsigned long long)K;
    double *ar = A + (unsigned long long)blockIdx.x * 2ull * (unsigned long long)Lp, *ai = ar + Lp;
    for (int k = threadIdx.x; k < Lp; k += blockDim.x) {
        double re = 0.0, im = 0.0;
        if (k < N) {
            double fx, fy;                                       // full[k]
            if (k < K) { fx = (double)row[k].x; fy = (double)row[k].y; }
            else { fx = (double)row[N - k].x; fy = -(double)row[N - k].y; }
            const double2 ch = chirp[k];
            re = fx * ch.x + fy * ch.y;                          // conj(full) * c = (fx - i fy)(cx + i cy)
            im = fx * ch.y - fy * ch.x;
        }
        ar[k] = re;
        ai[k] = im;
    }
}
// S: half spectra (Lp/2 + 1 bins) of the two real rows of every frame.  A <- conj(FFT(a) * bhat), again as two real rows.
__global__ void __launch_bounds__(256) mul_kernel(const double2 *S, const double2 *bhat, int Lp, double *A) {
    const int Kp = Lp / 2 + 1;
    const double2 *sr = S + (unsigned long long)blockIdx.x * 2ull * (unsigned long long)Kp, *si = sr + Kp;
    double *ar = A + (unsigned long long)blockIdx.x * 2ull * (unsigned long long)Lp, *ai = ar + Lp;
    for (int k = threadIdx.x; k < Lp; k += blockDim.x) {
        double2 R, I;
        if (k < Kp) { R = sr[k]; I = si[k]; }
        else { R = sr[Lp - k]; R.y = -R.y; I = si[Lp - k]; I.y = -I.y; }     // Hermitian halves of the two real signals
        const double ax = R.x - I.y, ay = R.y + I.x;             // FFT(re + i im) = R + i I
        const double2 b = bhat[k];
        ar[k] = ax * b.x - ay * b.y;                             // conj(product): real part kept, imaginary part negated
        ai[k] = -(ax * b.y + ay * b.x);
    }
}
// element m of the convolution, times the chirp: X[m] = c[m] * conj(FFT(conj C)[m]) / Lp
__device__ __forceinline__ double2 post_value(const double2 *sr, const double2 *si, const double2 *chirp, int m, double inv_lp) {
    const double2 R = sr[m], I = si[m];
    const double ex = (R.x - I.y) * inv_lp, ey = -(R.y + I.x) * inv_lp;
    const double2 ch = chirp[m];
    return double2{ex * ch.x - ey * ch.y, ex * ch.y + ey * ch.x};
}
template <typename T>
__global__ void __launch_bounds__(256) post_fwd_kernel(const double2 *S, const double2 *chirp, int N, int Lp, aus::Cx<T> *spec, double *mags) {
    const int Kp = Lp / 2 + 1, K = N / 2 + 1;
    const double2 *sr = S + (unsigned long long)blockIdx.x * 2ull * (unsigned long long)Kp, *si = sr + Kp;
    const double inv_lp = 1.0 / (double)Lp;
    for (int m = threadIdx.x; m < K; m += blockDim.x) {
        const double2 x = post_value(sr, si, chirp, m, inv_lp);
        if (spec) spec[(unsigned long long)blockIdx.x * K + m] = aus::Cx<T>{(T)x.x, (T)x.y};
        if (mags) mags[(unsigned long long)blockIdx.x * K + m] = sqrt(x.x * x.x + x.y * x.y);      // fft_tools.rs:41-55
    }
}
__global__ void __launch_bounds__(256) post_inv_kernel(const double2 *S, const double2 *chirp, int N, int Lp, double *y) {
    const int Kp = Lp / 2 + 1;
    const double2 *sr = S + (unsigned long long)blockIdx.x * 2ull * (unsigned long long)Kp, *si = sr + Kp;
    const double inv_lp = 1.0 / (double)Lp;
    for (int n = threadIdx.x; n < N; n += blockDim.x) y[(unsigned long long)blockIdx.x * N + n] = post_value(sr, si, chirp, n, inv_lp).x;   // `.re`, fft.rs:235
}
// window + overlap-add + window norm + peak (fft.rs:234-258, fft_tools.rs:69-112) as a gather: one thread per output sample
// of a shard, covering frames added in ascending order.  grid.y = shard.
template <typename T>
__global__ void __launch_bounds__(256) ola_kernel(const double *y, T *out, const unsigned long long *first, const unsigned long long *frames,
                                                 const unsigned long long *row0, const unsigned long long *samp0,
                                                 const unsigned long long *ooff, const unsigned long long *olen, int N, int hop,
                                                 const double *win, int apply_window, unsigned long long *max_bits) {
    const int sh = blockIdx.y;
    const unsigned long long F = frames[sh], base = first[sh], r0 = row0[sh], s0 = samp0[sh], n = olen[sh];
    T *o = out + ooff[sh];
    double vmax = 0.0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long s = s0 + i;
        const unsigned long long h_lo = s >= (unsigned long long)N ? (s - N) / hop + 1 : 0ull;
        unsigned long long h_hi = s / hop;
        if (h_hi > F - 1) h_hi = F - 1;
        double acc = 0.0, norm = 0.0;
        for (unsigned long long h = h_lo; h <= h_hi; ++h) {
            const unsigned long long u = s - h * hop;
            const double v = y[(base + (h - r0)) * (unsigned long long)N + u];
            if (apply_window) {
                const double w = win[u];
                acc += v * w;
                norm += w * w;
            } else {
                acc += v;
            }
        }
        double val = acc;
        if (apply_window) {
            val = acc / norm;                                    // fft.rs:250-252
            const double vt = (double)(T)val;
            if (vt > vmax) vmax = vt;                            // f64::max skips NaN
        }
        o[i] = (T)val;
    }
    if (apply_window) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, vmax, off);
            vmax = ov > vmax ? ov : vmax;
        }
        if ((threadIdx.x & 31) == 0 && vmax > 0.0) atomicMax(max_bits + sh, (unsigned long long)__double_as_longlong(vmax));
    }
}

}  // namespace anyn
