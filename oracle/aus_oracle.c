/* CPU oracle / baseline port of the aus STFT + spectral-feature path (plain C, f64).
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into the product library.  Only tests/, smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  PARITY STATUS: parity unpinned
 * (see oracle/aus_oracle.py header): the reference is Rust + rustfft ^6.2.0 and cannot be built
 * in this image; this file restates its arithmetic and its *structure* (so that the timed CPU
 * baseline does the same work per frame as the reference does):
 *   - rstft: complex N-point transform of real data, fresh buffers per frame   (fft.rs:119-178)
 *   - complex_to_polar_rstft: sqrt AND atan2 per bin, phases discarded          (fft_tools.rs:41-55, mp.rs:43)
 *   - analyzer: ~14 passes, per-frame allocations, pow() calls                  (analyzer.rs:41-81)
 *   - stft_analysis: serial STFT + polar, row copies, analyzer on a thread pool (mp.rs:26-119)
 * rustfft itself is replaced by a table-driven iterative radix-4 FFT (forward sign e^{-2 pi i jk/N},
 * un-normalised in both directions, as rustfft documents; radix 4 so that the timed CPU baseline is not a
 * strawman next to rustfft's radix-4 / radix-8 butterflies -- no SIMD intrinsics here, gcc -O3 only).  Paths are relative to /root/reference/src.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define AUS_PI 3.14159265358979323846264338327950288

typedef struct { double re, im; } cplx;

/* ------------------------------------------------------------------ window.rs:26-118 */
void aus_oracle_generate_window(int type, size_t L, double *w) {
    double Lf = (double)L;
    for (size_t i = 0; i < L; ++i) {
        double x = (double)i;
        switch (type) {
        case 0: w[i] = (2.0 / (Lf - 1.0)) * ((Lf - 1.0) / 2.0 - fabs(x - (Lf - 1.0) / 2.0)); break; /* :29 */
        case 1: w[i] = 0.42 - 0.5 * cos((2.0 * AUS_PI * x) / Lf) + 0.08 * cos((4.0 * AUS_PI * x) / Lf); break; /* :46 */
        case 2: w[i] = 0.5 - 0.5 * cos((2.0 * AUS_PI * x) / (Lf - 1.0)); break;   /* :64 */
        case 3: w[i] = 0.54 - 0.46 * cos((2.0 * AUS_PI * x) / (Lf - 1.0)); break; /* :81 */
        default: w[i] = 1.0; break;                                              /* :96 */
        }
    }
}

/* ------------------------------------------------------------------ FFT plan (stands in for rustfft) */
typedef struct { size_t n; int pow2; cplx *tw; uint32_t *rev; } fft_plan;

static int plan_init(fft_plan *p, size_t n) {
    p->n = n; p->pow2 = n && !(n & (n - 1));
    p->tw = (cplx *)malloc(sizeof(cplx) * (n ? n : 1));
    p->rev = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    if (!p->tw || !p->rev) return -1;
    for (size_t k = 0; k < n; ++k) {
        double a = -2.0 * AUS_PI * (double)k / (double)n;
        p->tw[k].re = cos(a); p->tw[k].im = sin(a);
    }
    if (p->pow2) {
        int lg = 0; while (((size_t)1 << lg) < n) ++lg;
        for (size_t i = 0; i < n; ++i) {
            uint32_t r = 0;
            for (int b = 0; b < lg; ++b) if (i & ((size_t)1 << b)) r |= 1u << (lg - 1 - b);
            p->rev[i] = r;
        }
    }
    return 0;
}
static void plan_free(fft_plan *p) { free(p->tw); free(p->rev); }

/* in-place transform; inverse != 0 conjugates the twiddles (no 1/N) */
static void fft_process(const fft_plan *p, cplx *x, int inverse) {
    size_t n = p->n;
    if (n <= 1) return;
    if (!p->pow2) {   /* O(n^2) DFT: only used by small tests */
        cplx *y = (cplx *)malloc(sizeof(cplx) * n);
        for (size_t k = 0; k < n; ++k) {
            double sr = 0.0, si = 0.0;
            for (size_t j = 0; j < n; ++j) {
                cplx w = p->tw[(j * k) % n];
                double wi = inverse ? -w.im : w.im;
                sr += x[j].re * w.re - x[j].im * wi;
                si += x[j].re * wi + x[j].im * w.re;
            }
            y[k].re = sr; y[k].im = si;
        }
        memcpy(x, y, sizeof(cplx) * n); free(y);
        return;
    }
    for (size_t i = 0; i < n; ++i) { size_t r = p->rev[i]; if (i < r) { cplx t = x[i]; x[i] = x[r]; x[r] = t; } }
    /* radix-4 decimation in time on the bit-reversed data (two radix-2 stages merged: three twiddle products per four
     * points, half the sweeps over the array), after one plain radix-2 stage when log2 n is odd.  The transform is the
     * one rustfft documents -- forward sign, no scaling; only the summation grouping differs (~1e-16). */
    size_t m = 1;                                    /* size of the finished sub-transforms */
    int lg = 0; while (((size_t)1 << lg) < n) ++lg;
    if (lg & 1) {
        for (size_t i = 0; i < n; i += 2) {
            cplx a = x[i], b = x[i + 1];
            x[i].re = a.re + b.re; x[i].im = a.im + b.im;
            x[i + 1].re = a.re - b.re; x[i + 1].im = a.im - b.im;
        }
        m = 2;
    }
    const double sgn = inverse ? -1.0 : 1.0;         /* twiddle conjugation and the sign of the +-i rotation */
    for (; m < n; m <<= 2) {
        size_t step = n / (4 * m);
        for (size_t i = 0; i < n; i += 4 * m)
            for (size_t j = 0; j < m; ++j) {
                cplx w1 = p->tw[j * step], w2 = p->tw[2 * j * step], w3 = p->tw[3 * j * step];
                w1.im *= sgn; w2.im *= sgn; w3.im *= sgn;
                cplx a = x[i + j], b = x[i + j + m], c = x[i + j + 2 * m], d = x[i + j + 3 * m];
                double Br = b.re * w2.re - b.im * w2.im, Bi = b.re * w2.im + b.im * w2.re;
                double Cr = c.re * w1.re - c.im * w1.im, Ci = c.re * w1.im + c.im * w1.re;
                double Dr = d.re * w3.re - d.im * w3.im, Di = d.re * w3.im + d.im * w3.re;
                double a1r = a.re + Br, a1i = a.im + Bi, b1r = a.re - Br, b1i = a.im - Bi;
                double sr = Cr + Dr, si = Ci + Di, dr = Cr - Dr, di = Ci - Di;
                /* (-i)(dr + i di) = di - i dr  (forward);  (+i)(dr + i di) = -di + i dr  (inverse) */
                double vr = sgn * di, vi = -sgn * dr;
                x[i + j].re = a1r + sr;         x[i + j].im = a1i + si;
                x[i + j + 2 * m].re = a1r - sr; x[i + j + 2 * m].im = a1i - si;
                x[i + j + m].re = b1r + vr;     x[i + j + m].im = b1i + vi;
                x[i + j + 3 * m].re = b1r - vr; x[i + j + 3 * m].im = b1i - vi;
            }
    }
}

/* ------------------------------------------------------------------ fft.rs */
size_t aus_oracle_num_frames(size_t n, size_t N, size_t hop) {           /* loop shape at fft.rs:126-176 */
    if (n <= N) return 1;
    return (n - N + hop - 1) / hop + 1;
}

/* one frame of fft.rs:128-176: window (short window for a short last frame), zero pad, complex FFT */
static void stft_one_frame(const fft_plan *pl, const double *audio, size_t n, size_t N, size_t hop,
                           int wtype, const double *window, size_t h, cplx *out_bins /* K */) {
    size_t start = h * hop;
    size_t end = (start + N < n) ? start + N : n;                         /* :130-135 (strict <) */
    size_t L = end - start;
    cplx *buf = (cplx *)malloc(sizeof(cplx) * N);                         /* fresh Vec per frame, :152/:162 */
    if (L < N) {                                                          /* num_zeros > 0, :148-159 */
        double *sw = (double *)malloc(sizeof(double) * (L ? L : 1));
        aus_oracle_generate_window(wtype, L, sw);
        for (size_t i = 0; i < L; ++i) { buf[i].re = audio[start + i] * sw[i]; buf[i].im = 0.0; }
        for (size_t i = L; i < N; ++i) { buf[i].re = 0.0; buf[i].im = 0.0; }
        free(sw);
    } else {
        for (size_t i = 0; i < N; ++i) { buf[i].re = audio[start + i] * window[i]; buf[i].im = 0.0; }
    }
    fft_process(pl, buf, 0);                                              /* :171 */
    memcpy(out_bins, buf, sizeof(cplx) * (N / 2 + 1));                    /* :172 */
    free(buf);
}

int aus_oracle_rstft(const double *audio, size_t n, size_t N, size_t hop, int wtype, double *out /* F*K*2 */) {
    if (hop == 0 || hop > N || N < 2) return -1;
    fft_plan pl; if (plan_init(&pl, N)) return -2;
    double *window = (double *)malloc(sizeof(double) * N);
    aus_oracle_generate_window(wtype, N, window);
    size_t F = aus_oracle_num_frames(n, N, hop), K = N / 2 + 1;
    for (size_t h = 0; h < F; ++h) stft_one_frame(&pl, audio, n, N, hop, wtype, window, h, (cplx *)out + h * K);
    free(window); plan_free(&pl);
    return 0;
}

int aus_oracle_rfft(const double *audio, size_t N, double *out /* K*2 */) {           /* fft.rs:37-47 */
    fft_plan pl; if (plan_init(&pl, N)) return -2;
    cplx *buf = (cplx *)malloc(sizeof(cplx) * N);
    for (size_t i = 0; i < N; ++i) { buf[i].re = audio[i]; buf[i].im = 0.0; }
    fft_process(&pl, buf, 0);
    memcpy(out, buf, sizeof(cplx) * (N / 2 + 1));
    free(buf); plan_free(&pl);
    return 0;
}

static void hermitian_extend(const cplx *s, size_t N, cplx *full) {                   /* fft.rs:76-83 / :221-229 */
    size_t K = N / 2 + 1;
    for (size_t i = 0; i < K; ++i) full[i] = s[i];
    for (size_t i = 0; i + 1 < N / 2; ++i) { full[K + i].re = s[N / 2 - 1 - i].re; full[K + i].im = -s[N / 2 - 1 - i].im; }
}

int aus_oracle_irfft(const double *spec /* K*2 */, size_t N, double *audio /* N */) { /* fft.rs:66-93 */
    fft_plan pl; if (plan_init(&pl, N)) return -2;
    cplx *full = (cplx *)malloc(sizeof(cplx) * N);
    hermitian_extend((const cplx *)spec, N, full);
    fft_process(&pl, full, 1);
    for (size_t i = 0; i < N; ++i) audio[i] = full[i].re;
    free(full); plan_free(&pl);
    return 0;
}

int aus_oracle_irstft(const double *spec /* F*K*2 */, size_t F, size_t N, size_t hop, int wtype,
                      double *audio /* N + hop*(F-1) */) {                            /* fft.rs:204-268 */
    if (F == 0 || hop == 0) return -1;
    fft_plan pl; if (plan_init(&pl, N)) return -2;
    size_t K = N / 2 + 1, out_len = N + hop * (F - 1);
    double *w = (double *)malloc(sizeof(double) * N);
    double *norm = (double *)calloc(out_len, sizeof(double));
    double *chunks = (double *)malloc(sizeof(double) * F * N);
    cplx *full = (cplx *)malloc(sizeof(cplx) * N);
    aus_oracle_generate_window(wtype, N, w);
    for (size_t h = 0; h < F; ++h) {
        hermitian_extend((const cplx *)spec + h * K, N, full);
        fft_process(&pl, full, 1);                                                    /* :231 */
        for (size_t j = 0; j < N; ++j) chunks[h * N + j] = full[j].re * w[j];         /* :234-236 */
        for (size_t j = 0; j < N; ++j) norm[h * hop + j] += pow(w[j], 2.0);           /* :239-243 */
    }
    /* overlap_add, fft_tools.rs:69-112: each sample sums its covering frames in ascending frame order */
    for (size_t i = 0; i < out_len; ++i) {
        size_t lo = (i >= N) ? (i - N) / hop + 1 : 0;
        size_t hi = i / hop; if (hi > F - 1) hi = F - 1;
        double s = 0.0;
        for (size_t h = lo; h <= hi; ++h) s += chunks[h * N + (i - h * hop)];
        audio[i] = s;
    }
    for (size_t i = 0; i < out_len; ++i) audio[i] /= norm[i];                         /* :250-252 */
    double maxval = 0.0;
    for (size_t i = 0; i < out_len; ++i) maxval = fmax(audio[i], maxval);             /* :255-258 */
    double scaler = pow(10.0, -6.0 / 20.0) / maxval;                                  /* :261-262 */
    for (size_t i = 0; i < out_len; ++i) audio[i] *= scaler;
    free(w); free(norm); free(chunks); free(full); plan_free(&pl);
    return 0;
}

/* ------------------------------------------------------------------ spectral_analysis_tools.rs:79-290 */
static double dot_product(const double *a, const double *b, size_t n) {               /* :253-259 */
    double s = 0.0; for (size_t i = 0; i < n; ++i) s += a[i] * b[i]; return s;
}
static double seq_sum(const double *a, size_t n) { double s = 0.0; for (size_t i = 0; i < n; ++i) s += a[i]; return s; }

static double slope_region(const double *p, const double *f, size_t K, double fl, double fu, uint32_t sr) { /* :190-230 */
    double fund = (double)sr / ((double)(K - 1) * 2.0);
    double m_fl = fl / fund, m_fu = fu / fund, n = (double)K;
    size_t ac = (size_t)ceil(m_fl), af = (size_t)floor(m_fl), bc = (size_t)ceil(m_fu), bf = (size_t)floor(m_fu);
    if (bc >= K || ac > bf) return NAN;   /* the reference panics here */
    double yl = p[af] + (m_fl - (double)af) * (p[ac] - p[af]);
    double yu = p[bf] + (m_fu - (double)bf) * (p[bc] - p[bf]);
    double sum_x = fl + seq_sum(f + ac, bf - ac) + fu;
    double sum_y = yl + seq_sum(p + ac, bf - ac) + yu;
    double sum_x2 = pow(fl, 2.0) + dot_product(f + ac, f + ac, bf - ac) + pow(fu, 2.0);
    double sum_xy = fl * yl + dot_product(p + ac, f + ac, bf - ac) + fu * yu;
    return (n * sum_xy - sum_x * sum_y) / (n * sum_x2 - pow(sum_x, 2.0));
}

static double roll_off(const double *p, const double *f, size_t K, double psum, double n) {  /* :137-145 */
    int64_t i = -1; double cum = 0.0;
    while (cum < n && i < (int64_t)K - 1) { i += 1; cum += p[i] / psum; }
    return f[i < 0 ? 0 : i];
}

/* analyzer.rs:41-81; out[14] in the declaration order of analyzer.rs:10-23 */
void aus_oracle_analyzer(const double *m, size_t K, size_t N, uint32_t sr, double *out) {
    double *p = (double *)malloc(sizeof(double) * K);       /* make_power_spectrum, :273-279 */
    for (size_t i = 0; i < K; ++i) p[i] = pow(m[i], 2.0);
    double msum = seq_sum(m, K), psum = seq_sum(p, K);
    double *pmf = (double *)malloc(sizeof(double) * K);     /* make_spectrum_pmf, :284-290 */
    for (size_t i = 0; i < K; ++i) pmf[i] = p[i] / psum;
    double *f = (double *)malloc(sizeof(double) * K);       /* rfftfreq per frame, fft_tools.rs:199-206 */
    double f0 = (double)sr / (double)N; f[0] = 0.0;
    for (size_t i = 1; i < K; ++i) f[i] = f0 * (double)i;
    double c = dot_product(m, f, K) / msum;                 /* :79-85 */
    double var = 0.0, sk = 0.0, ku = 0.0, ent = 0.0, lg = 0.0;
    for (size_t i = 0; i < K; ++i) var += pow(f[i] - c, 2.0) * pmf[i];   /* :238-244 */
    for (size_t i = 0; i < K; ++i) sk += pow(f[i] - c, 3.0) * pmf[i];    /* :153-160 */
    sk /= pow(var, 1.5);
    for (size_t i = 0; i < K; ++i) ku += pow(f[i] - c, 4.0) * pmf[i];    /* :122-129 */
    ku /= pow(var, 2.0);
    for (size_t i = 0; i < K; ++i) ent += pmf[i] * log2(pmf[i]);         /* :93-99 */
    for (size_t i = 0; i < K; ++i) lg += log(m[i]);                      /* :107-114 */
    lg /= (double)K;
    double n = (double)K;                                                /* slope, :167-182 */
    double *x = (double *)malloc(sizeof(double) * K);
    for (size_t i = 0; i < K; ++i) x[i] = (double)i;
    double sum_x = n * (n - 1.0) / 2.0, sum_x2 = n * (n - 1.0) * (2.0 * n - 1.0) / 6.0;
    out[0] = c; out[1] = var; out[2] = sk; out[3] = ku; out[4] = -ent;
    out[5] = exp(lg) * (double)K / msum;
    out[6] = roll_off(p, f, K, psum, 0.5); out[7] = roll_off(p, f, K, psum, 0.75);
    out[8] = roll_off(p, f, K, psum, 0.9); out[9] = roll_off(p, f, K, psum, 0.95);
    out[10] = (n * dot_product(p, x, K) - sum_x * psum) / (n * sum_x2 - pow(sum_x, 2.0));
    out[11] = slope_region(p, f, K, 0.0, 1000.0, sr);
    out[12] = slope_region(p, f, K, 1000.0, 5000.0, sr);
    out[13] = slope_region(p, f, K, 0.0, 5000.0, sr);
    free(p); free(pmf); free(f); free(x);
}

/* ------------------------------------------------------------------ mp.rs:26-119 */
typedef struct {
    /* reference-structure job: analyzer over a private copy of contiguous magnitude rows */
    double *rows; size_t nrows, K, N; uint32_t sr; double *out;
    /* all-parallel job: STFT + magnitude + analyzer over a frame range */
    const double *audio; size_t n, hop, h0, h1; int wtype; const double *window; int all_parallel;
} job_t;

static void *job_run(void *arg) {
    job_t *j = (job_t *)arg;
    if (!j->all_parallel) {
        for (size_t r = 0; r < j->nrows; ++r) aus_oracle_analyzer(j->rows + r * j->K, j->K, j->N, j->sr, j->out + r * 14);
        return NULL;
    }
    fft_plan pl; if (plan_init(&pl, j->N)) return NULL;
    cplx *bins = (cplx *)malloc(sizeof(cplx) * j->K);
    double *mag = (double *)malloc(sizeof(double) * j->K);
    for (size_t h = j->h0; h < j->h1; ++h) {
        stft_one_frame(&pl, j->audio, j->n, j->N, j->hop, j->wtype, j->window, h, bins);
        for (size_t k = 0; k < j->K; ++k) mag[k] = sqrt(pow(bins[k].re, 2.0) + pow(bins[k].im, 2.0));
        aus_oracle_analyzer(mag, j->K, j->N, j->sr, j->out + (h - j->h0) * 14);
    }
    free(bins); free(mag); plan_free(&pl);
    return NULL;
}

/* mode 0: the reference's structure (serial rstft + sqrt/atan2 polar, row copies, analyzer on the pool);
 * mode 1: every stage parallel over frame ranges, no atan2 (the "not flattered" variant of BASELINE.md 3).
 * threads == 0 -> 1.  Output: F x 14 in frame order (mp.rs:100-116). */
int aus_oracle_stft_features_mt(const double *audio, size_t n, size_t N, size_t hop, int wtype, uint32_t sr,
                                size_t threads, int mode, double *out) {
    if (hop == 0 || hop > N || N < 2) return -1;
    if (threads == 0) threads = 1;
    size_t F = aus_oracle_num_frames(n, N, hop), K = N / 2 + 1;
    size_t per = (F + threads - 1) / threads;                                      /* mp.rs:54 */
    job_t *jobs = (job_t *)calloc(threads, sizeof(job_t));
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    double *window = (double *)malloc(sizeof(double) * N);
    aus_oracle_generate_window(wtype, N, window);
    double *mags = NULL;
    if (mode == 0) {
        double *spec = (double *)malloc(sizeof(double) * 2 * F * K);
        if (aus_oracle_rstft(audio, n, N, hop, wtype, spec)) { free(spec); return -2; }   /* mp.rs:42, serial */
        mags = (double *)malloc(sizeof(double) * F * K);
        double *phases = (double *)malloc(sizeof(double) * F * K);
        for (size_t i = 0; i < F * K; ++i) {                                       /* mp.rs:43, serial */
            mags[i] = sqrt(pow(spec[2 * i], 2.0) + pow(spec[2 * i + 1], 2.0));
            phases[i] = atan2(spec[2 * i + 1], spec[2 * i]);
        }
        free(phases); free(spec);                                                  /* phases discarded */
    }
    size_t started = 0;
    for (size_t t = 0; t < threads; ++t) {
        size_t s = t * per, e = s + per < F ? s + per : F;
        if (s >= F) break;
        job_t *j = &jobs[t];
        j->K = K; j->N = N; j->sr = sr; j->out = out + s * 14; j->all_parallel = mode;
        if (mode == 0) {
            j->nrows = e - s;
            j->rows = (double *)malloc(sizeof(double) * (e - s) * K);              /* deep copy, mp.rs:66-75 */
            memcpy(j->rows, mags + s * K, sizeof(double) * (e - s) * K);
        } else {
            j->audio = audio; j->n = n; j->hop = hop; j->h0 = s; j->h1 = e; j->wtype = wtype; j->window = window;
        }
        pthread_create(&tid[t], NULL, job_run, j);
        ++started;
    }
    for (size_t t = 0; t < started; ++t) { pthread_join(tid[t], NULL); free(jobs[t].rows); }
    free(mags); free(window); free(jobs); free(tid);
    return 0;
}

/* mp::stft_analysis(audio, fft_size, sample_rate, max_num_threads): hop = N/2, Hamming (mp.rs:42) */
int aus_oracle_stft_analysis(const double *audio, size_t n, size_t N, uint32_t sr, size_t threads, double *out) {
    return aus_oracle_stft_features_mt(audio, n, N, N / 2, 3, sr, threads, 0, out);
}

/* ------------------------------------------------------------------ synthetic input */
static inline uint64_t mix64(uint64_t seed, uint64_t file_id, uint64_t n) {
    uint64_t z = seed * 0x9E3779B97F4A7C15ull + file_id * 0xD1B54A32D192ED03ull + n * 0x94D049BB133111EBull
               + 0x2545F4914F6CDD1Dull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
void aus_oracle_synth(double *x, size_t count, size_t n0, uint32_t sr, double freq, uint64_t seed, uint64_t file_id) {
    double coef = 2.0 * AUS_PI * freq / (double)sr;                                /* synthesis.rs:52 */
    for (size_t i = 0; i < count; ++i) {
        double u = (double)(mix64(seed, file_id, n0 + i) >> 11) * 0x1p-52 - 1.0;
        x[i] = 0.5 * sin(coef * (double)(n0 + i)) + 0.1 * u;
    }
}
