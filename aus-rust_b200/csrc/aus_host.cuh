// aus_host.cuh -- host-side pieces shared by the translation units of libaus_b200.so: error state, the per-device
// context, and the kernel launch templates.  The kernels of one fft_size live in one translation unit (aus_kernels.cu
// compiled with -DAUS_LG=6..14) so that `make -j` builds the sizes in parallel; aus_b200.cu holds everything else.
#pragma once
#include "../../include/aus_b200.h"
#include "aus_istft_kernel.cuh"
#include "aus_stft_pc.cuh"

#include <algorithm>
#include <atomic>
#include <map>
#include <string>
#include <vector>

using namespace aus;

// ------------------------------------------------------------------------------------------------ errors
inline thread_local std::string g_err = "";
inline int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define CUDA_TRY(...)                                                                               \
    do {                                                                                            \
        cudaError_t e_ = (__VA_ARGS__);                                                             \
        if (e_ != cudaSuccess) {                                                                    \
            (void)cudaGetLastError(); /* un-latch, or the next launch check would report this one */ \
            return fail(AUS_ERR_CUDA, std::string(#__VA_ARGS__) + ": " + cudaGetErrorString(e_));   \
        }                                                                                           \
    } while (0)
#define AUS_TRY(...)                   \
    do {                               \
        int r_ = (__VA_ARGS__);        \
        if (r_ != AUS_OK) return r_;   \
    } while (0)

// ------------------------------------------------------------------------------------------------ context
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return AUS_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            p = nullptr;
            return fail(AUS_ERR_ALLOC, std::string("cudaMalloc: ") + cudaGetErrorString(e));
        }
        cap = want;
        return AUS_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct Tables {
    void *tw2 = nullptr, *tw3 = nullptr, *twNt = nullptr, *twN = nullptr;
    void *win2[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    void *winf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};     // full-scale window pairs (inverse transform)
};

namespace anyn { struct Tab; }        // chirp-z tables of one non-native fft_size (aus_anyn.cuh)

struct aus_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    int sm_count = 148;
    uint64_t launches = 0;
    int direct_load = 0;                      // experiment knob, see aus_ctx_set_direct_load
    uint64_t chunk_samples = 32ull << 20;     // host pipeline granularity: 32 Mi samples = 256 MB of f64
    int rr_mode = -1;                         // forward-kernel variant: -1 auto, 0 shared-memory tail, 1 register-resident tail,
                                              // 2 producer/consumer warps (f64 features at fft_size 2048; aus_stft_pc.cuh)
    std::map<int, Tables> tables;          // key = log2n * 2 + precision
    void *log2tab = nullptr;               // LOG2_TAB x (1/c_i, -log2(1/c_i)) for the table-driven f64 log2
    DevBuf in64, in32, out64, out32, aux64, aux32, desc, misc, raw, normtab;
    long long normtab_key = -1;            // (fft_size, hop, window, precision) the reciprocal-norm table was built for
    const unsigned long long *k5_out_off = nullptr, *k5_out_len = nullptr;   // K5 descriptors of the last irstft_k4 (in `desc`)
    DevBuf acorr;                          // padded chunks of aus_autocorrelation_batch
    std::map<size_t, anyn::Tab *> anytabs; // fft sizes outside the radix-32 kernels' range (aus_anyn.cuh), by fft_size
    DevBuf any_a, any_s, any_y, any_mag, any_desc, any_par, any_feat;   // their workspace
    std::vector<unsigned long long> desc_stage, desc_last;   // packed descriptor image being built / last uploaded to `desc`
    bool desc_valid = false;
    void *hostreg_ptr = nullptr;           // caller buffer currently page-locked by the host pipeline (cudaHostRegister)
    int hostreg_mode = -1;                 // pageable inputs: -1 auto (2 when >= hostreg_min_bytes), 0 driver staging, 1 cudaHostRegister,
                                           // 2 own pinned staging ring filled by several host threads
    void *hstage[2] = {nullptr, nullptr};  // the staging ring (pinned host memory)
    size_t hstage_cap[2] = {0, 0};
    static constexpr size_t XSTAGE_BYTES = (size_t)64 << 20;
    void *xstage[2] = {nullptr, nullptr};  // pinned ring of the generic big host <-> device moves (xfer_h2d / xfer_d2h)
    cudaEvent_t ev_x[2] = {nullptr, nullptr};
    void *hout[2] = {nullptr, nullptr};    // pinned ring for feature rows on their way to a pageable output array
    size_t hout_cap[2] = {0, 0};
    size_t hostreg_min_bytes = (size_t)64 << 20;
    DevBuf pin64[2], pin32[2], pout64[2], pout32[2], ppcm[2];   // double buffers of the chunked host pipeline
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
};

// ------------------------------------------------------------------------------------------------ launches
template <typename T, int LG, bool RR, bool FEAT> int launch_stft_v(aus_ctx *c, const StftParams<T> &p) {
    using C = Cfg<T, LG, RR>;
    static std::atomic<bool> attr_done[64];
    auto kern = stft_kernel<T, LG, RR, FEAT>;
    const size_t smem = FEAT ? C::smem_bytes() : C::off_log();          // the spectra-only build carries no log2 table
    if (!attr_done[c->device & 63]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done[c->device & 63] = true;
    }
    static std::atomic<int> occ_cache[64];
    int occ = occ_cache[c->device & 63].load();
    if (occ < 1) {
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::BLOCK, smem));
        if (occ < 1) return fail(AUS_ERR_CUDA, "stft kernel does not fit on an SM");
        occ_cache[c->device & 63].store(occ);
    }
    const unsigned long long need = (p.total_frames + C::FPB - 1) / C::FPB;
    const unsigned long long cap = (unsigned long long)c->sm_count * occ;
    const unsigned int grid = (unsigned int)std::max<unsigned long long>(1, std::min(need, cap));
    kern<<<grid, C::BLOCK, smem, c->stream>>>(p);
    CUDA_TRY(cudaGetLastError());
    c->launches++;
    return AUS_OK;
}
// Which forward kernel: shared-memory tail or register-resident tail (aus_stft_r.cuh).  Measured on B200 over
// fft_size 64..16384 x {f64, f32} x {features, spectra, both} on 0.5-2 h inputs (tools/kbench.py, AUS_B200_RR=0|1):
//   f32: the register-resident tail wins everywhere (features +3..32 %, spectra +18..90 %, both outputs in one launch
//        +17..110 %);
//   f64: both outputs +15..80 %; spectra alone +15..70 % (fft_size 8192: +-0); features alone +4..35 % up to fft_size 512
//        and +8 % at 16384, but -3..-5 % at 1024..8192 (its CTAs are 8 warps x 255 registers against 12 x 168 there,
//        and the FP64 feature pass needs the resident warps).
// c->rr_mode (aus_ctx_set_forward_variant / AUS_B200_RR) overrides the table for ablations.
inline bool use_rr(const aus_ctx *c, int lg, bool f32, bool feat, bool spec) {
    if (c->rr_mode >= 0 && c->rr_mode <= 1) return c->rr_mode == 1;
    if (f32 || !feat || spec) return true;
    return lg <= 9 || lg >= 14;                          // f64 features alone
}
// Producer/consumer build (aus_stft_pc.cuh): f64, features only, one warp per frame.  Measured on B200 against the
// single-role kernel on 1 h inputs: 1.15-1.19e8 frames/s against 1.22e8 (profiles/r2_k1_pc_experiment.md), so it is an
// opt-in variant (rr_mode 2 / AUS_B200_RR=2), not the default.
template <int LG> constexpr bool pc_built() { return LG == 11; }
template <int LG> int launch_stft_pc(aus_ctx *c, const StftParams<double> &p) {
    if constexpr (pc_built<LG>()) {
        using C = CfgPC<LG>;
        static std::atomic<bool> attr_done[64];
        auto kern = stft_pc_kernel<LG>;
        const size_t smem = C::smem_bytes();
        if (!attr_done[c->device & 63]) {
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_done[c->device & 63] = true;
        }
        const unsigned long long need = (p.total_frames + C::NPAIR - 1) / C::NPAIR;
        const unsigned int grid = (unsigned int)std::max<unsigned long long>(1, std::min<unsigned long long>(need, (unsigned long long)c->sm_count));
        kern<<<grid, C::BLOCK, smem, c->stream>>>(p);
        CUDA_TRY(cudaGetLastError());
        c->launches++;
        return AUS_OK;
    } else {
        return fail(AUS_ERR_UNSUPPORTED, "producer/consumer kernel is not built for this fft_size");
    }
}
template <typename T, int LG> int launch_stft_t(aus_ctx *c, const StftParams<T> &p) {
    if constexpr (sizeof(T) == 8 && pc_built<LG>()) {
        if (p.raw != nullptr && p.spec == nullptr && !p.direct && c->rr_mode == 2)
            return launch_stft_pc<LG>(c, p);
    }
    if (use_rr(c, LG, sizeof(T) == 4, p.raw != nullptr, p.spec != nullptr))
        return p.raw != nullptr ? launch_stft_v<T, LG, true, true>(c, p) : launch_stft_v<T, LG, true, false>(c, p);
    return launch_stft_v<T, LG, false, true>(c, p);
}
template <typename T, int LG> int launch_analyzer_t(aus_ctx *c, const StftParams<T> &p) {
    using C = CfgA<T, LG>;
    static std::atomic<bool> attr_done[64];
    auto kern = analyzer_kernel<T, LG>;
    const size_t smem = C::smem_bytes();
    if (!attr_done[c->device & 63]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done[c->device & 63] = true;
    }
    static std::atomic<int> occ_cache[64];
    int occ = occ_cache[c->device & 63].load();
    if (occ < 1) {
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::BLOCK, smem));
        if (occ < 1) return fail(AUS_ERR_CUDA, "analyzer kernel does not fit on an SM");
        occ_cache[c->device & 63].store(occ);
    }
    const unsigned long long need = (p.total_frames + C::FPB - 1) / C::FPB;
    const unsigned long long cap = (unsigned long long)c->sm_count * occ;
    const unsigned int grid = (unsigned int)std::max<unsigned long long>(1, std::min(need, cap));
    kern<<<grid, C::BLOCK, smem, c->stream>>>(p);
    CUDA_TRY(cudaGetLastError());
    c->launches++;
    return AUS_OK;
}
template <typename T, int LG> int launch_istft_t(aus_ctx *c, const IstftParams<T> &p) {
    using C = Cfg<T, LG>;
    static std::atomic<bool> attr_done[64];
    auto kern = istft_kernel<T, LG>;
    const size_t smem = C::off_log();                                     // no log2 table in the inverse kernel
    if (!attr_done[c->device & 63]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done[c->device & 63] = true;
    }
    static std::atomic<int> occ_cache[64];
    int occ = occ_cache[c->device & 63].load();
    if (occ < 1) {
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::BLOCK, smem));
        if (occ < 1) return fail(AUS_ERR_CUDA, "istft kernel does not fit on an SM");
        occ_cache[c->device & 63].store(occ);
    }
    const unsigned long long need = (p.n_seg + C::FPB - 1) / C::FPB;
    const unsigned long long cap = (unsigned long long)c->sm_count * occ;
    const unsigned int grid = (unsigned int)std::max<unsigned long long>(1, std::min(need, cap));
    kern<<<grid, C::BLOCK, smem, c->stream>>>(p);
    CUDA_TRY(cudaGetLastError());
    c->launches++;
    return AUS_OK;
}

