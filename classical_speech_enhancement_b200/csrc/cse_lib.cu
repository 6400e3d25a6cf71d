// libcse_sm100a: C ABI (include/cse.h) over the hand-written kernels.
// Built for sm_100a by nvcc (build.py) and, for CPU-side kernel tests only, by g++ with
// -DCSE_EMU (tests/emu).  Host code here only validates arguments, sizes launches and
// enqueues kernels on the caller's stream; it never allocates or synchronises.
#define CSE_EMU_IMPL
#include "cse_common.cuh"
#include "k_stft.cuh"
#include "k_enhance.cuh"
#include "k_noise.cuh"
#include "k_score.cuh"
#include "k_stoi_stream.cuh"
#include "k_select.cuh"

#include <mutex>
#include <cstdarg>

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
static int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(CSE_ECUDA, "%s: %s", what, cudaGetErrorString(e));
    return CSE_OK;
}
// Opt-in to > 48 KB of dynamic shared memory once per (kernel instantiation, device) instead of on every launch.
#define CSE_SMEM_OPT_IN(kfn, bytes)                                                                        \
    do {                                                                                                    \
        static unsigned long long cse_done_[2] = {0, 0};                                                    \
        int cse_dev_ = 0;                                                                                   \
        cudaGetDevice(&cse_dev_);                                                                           \
        if ((size_t)(bytes) > 48 * 1024 && !((cse_done_[(cse_dev_ >> 6) & 1] >> (cse_dev_ & 63)) & 1ull)) { \
            cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes));           \
            cse_done_[(cse_dev_ >> 6) & 1] |= 1ull << (cse_dev_ & 63);                                       \
        }                                                                                                   \
    } while (0)
static bool valid_nfft(int n_fft) { return n_fft == 256 || n_fft == 512 || n_fft == 1024 || n_fft == 2048; }
#define CSE_REQUIRE(cond, ...) do { if (!(cond)) return fail(CSE_EINVAL, __VA_ARGS__); } while (0)

extern "C" {

int cse_abi_version(void) { return CSE_ABI_VERSION; }
int cse_dtype(void) { return CSE_REAL_BITS; }
const char* cse_last_error(void) { return g_err; }
int cse_bins_padded(int n_fft) { return cse_nbp(n_fft); }
int cse_num_frames(int length, int hop) { return hop > 0 ? 1 + length / hop : 0; }

// ------------------------------------------------------------------ tables
size_t cse_tables_bytes(void) { return sizeof(CseTables); }

static CseTables* host_tables() {
    static CseTables* t = nullptr;
    static std::once_flag once;
    std::call_once(once, []() {
        t = new CseTables;
        memset(t, 0, sizeof(CseTables));
        const double PI = 3.14159265358979323846;
        for (int k = 0; k < CSE_TW_N / 2; ++k) {
            const double a = -2.0 * PI * k / CSE_TW_N;
            t->tw[k] = mk2((real)cos(a), (real)sin(a));
        }
        auto hann = [&](real* w, int n) { for (int i = 0; i < n; ++i) w[i] = (real)(0.5 - 0.5 * cos(2.0 * PI * i / n)); };
        hann(t->hann256, 256); hann(t->hann512, 512); hann(t->hann1024, 1024); hann(t->hann2048, 2048);
        // np.hanning(258)[1:-1]: 0.5 - 0.5 cos(2 pi (i+1) / 257)
        for (int i = 0; i < 256; ++i) t->stoi_win[i] = (real)(0.5 - 0.5 * cos(2.0 * PI * (i + 1) / 257.0));
        cse_fill_resampler(t);
    });
    return t;
}

int cse_tables_init(void* tables, void* stream) {
    CSE_REQUIRE(tables != nullptr, "tables is NULL");
    cudaError_t e = cudaMemcpyAsync(tables, host_tables(), sizeof(CseTables), cudaMemcpyHostToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(CSE_ECUDA, "tables copy: %s", cudaGetErrorString(e));
    return CSE_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ K1 STFT + PSD
template <int LOG2N, int F>
static int launch_stft(const CseTables* T, const real* wav, const real* minus, int U, int L, int hop, int nf,
                       real floor_, real2* Y, real* P, void* stream) {
    constexpr int M = (1 << LOG2N) / 2;
    const size_t smem = (size_t)F * CSE_FFT_STRIDE(M) * sizeof(real2);
    auto kfn = stft_psd_kernel<LOG2N, F>;
    CSE_SMEM_OPT_IN(kfn, smem);
    dim3 grid((nf + F - 1) / F, U);
    CSE_LAUNCH(kfn, grid, 256, smem, stream, T, wav, minus, L, hop, nf, floor_, Y, P);
    return check_launch("stft_psd_kernel");
}

extern "C" int cse_stft_psd(const void* tables, const void* wav, const void* minus, int n_utts, int length, int n_fft,
                 int hop, double psd_floor, void* Y, void* P, void* stream) {
    CSE_REQUIRE(tables && wav, "tables/wav is NULL");
    CSE_REQUIRE(valid_nfft(n_fft), "n_fft %d not in {256,512,1024,2048}", n_fft);
    CSE_REQUIRE(hop > 0 && hop <= n_fft / 2 && hop % 2 == 0, "hop %d must be even and <= n_fft/2", hop);
    CSE_REQUIRE(n_utts > 0 && length > n_fft / 2, "need n_utts > 0 and length > n_fft/2 (reflect padding)");
    const int nf = cse_num_frames(length, hop);
    const CseTables* T = (const CseTables*)tables;
    const real* w = (const real*)wav;
    const real* m = (const real*)minus;
    switch (n_fft) {
        case 256: return launch_stft<8, 8>(T, w, m, n_utts, length, hop, nf, (real)psd_floor, (real2*)Y, (real*)P, stream);
        case 512: return launch_stft<9, 8>(T, w, m, n_utts, length, hop, nf, (real)psd_floor, (real2*)Y, (real*)P, stream);
        case 1024: return launch_stft<10, 8>(T, w, m, n_utts, length, hop, nf, (real)psd_floor, (real2*)Y, (real*)P, stream);
        default: return launch_stft<11, 4>(T, w, m, n_utts, length, hop, nf, (real)psd_floor, (real2*)Y, (real*)P, stream);
    }
}

// ------------------------------------------------------------------ K3+K4 gain + ISTFT
// Dynamic shared memory of enhance_kernel<ALG, LOG2N, STAGED> (must mirror the carve-up at the top of the kernel):
// FFT buffer + zero cell, overlap-add ring, steady-state window sum-of-squares, parameter slots, window pairs,
// per-pass twiddles, 1/(N wss) pairs, then (16-byte aligned)
// the TMA tile of F frames of Y (+ of a time-varying noise PSD) and its mbarrier.
#ifndef CSE_ENH_STAGED_MAX_LOG2N
#define CSE_ENH_STAGED_MAX_LOG2N 10
#endif
#ifndef CSE_ENH_STAGED_MIN_LOG2N
#define CSE_ENH_STAGED_MIN_LOG2N 8
#endif
template <int LOG2N>
static size_t enhance_smem_bytes(int hop, int noise_tv, bool staged) {
    typedef EnhanceCfg<LOG2N> C;
    const int W = C::NFFT + (C::F - 1) * hop;
    size_t s = (size_t)(C::F * C::XST + 1) * sizeof(real2) + (size_t)(C::M + FftTwLayout<LOG2N - 1, true>::SIZE + C::M / 2) * sizeof(real2) +
               (size_t)(16 + W + hop) * sizeof(real);
    s += 128;                                                        // alignment slack of the tile
    if (staged) s += (size_t)C::F * cse_nbp(C::NFFT) * (sizeof(real2) + (noise_tv ? sizeof(real) : 0));
    return s + 16;                                                   // mbarrier
}
template <int ALG, int LOG2N>
static int launch_enhance(const EnhanceArgs& a, int n_items, void* stream) {
    typedef EnhanceCfg<LOG2N> C;
    // n_fft 2048: the tile would cost a resident CTA (see k_enhance.cuh)
    constexpr bool STAGED = LOG2N <= CSE_ENH_STAGED_MAX_LOG2N && LOG2N >= CSE_ENH_STAGED_MIN_LOG2N;
    const size_t smem = enhance_smem_bytes<LOG2N>(a.hop, a.noise_tv, STAGED);
    if (a.noise_tv == 2) {
        if constexpr (ALG == 0) return fail(CSE_EINVAL, "spectral subtraction takes the noise PSD, not the a-posteriori SNR (noise_tv 2)");
        else {
            auto kfn = enhance_kernel<ALG, LOG2N, STAGED, true, true>;
            CSE_SMEM_OPT_IN(kfn, enhance_smem_bytes<LOG2N>(C::NFFT / 2, 1, STAGED));
            CSE_LAUNCH(kfn, n_items, C::NT, smem, stream, a);
        }
    } else if (a.noise_tv) {
        auto kfn = enhance_kernel<ALG, LOG2N, STAGED, true>;
        CSE_SMEM_OPT_IN(kfn, enhance_smem_bytes<LOG2N>(C::NFFT / 2, 1, STAGED));   // covers every hop of the instantiation
        CSE_LAUNCH(kfn, n_items, C::NT, smem, stream, a);
    } else {
        auto kfn = enhance_kernel<ALG, LOG2N, STAGED, false>;
        CSE_SMEM_OPT_IN(kfn, enhance_smem_bytes<LOG2N>(C::NFFT / 2, 0, STAGED));
        CSE_LAUNCH(kfn, n_items, C::NT, smem, stream, a);
    }
    return check_launch("enhance_kernel");
}
template <int ALG>
static int dispatch_enhance(const EnhanceArgs& a, int n_fft, int n_items, void* stream) {
    switch (n_fft) {
        case 256: return launch_enhance<ALG, 8>(a, n_items, stream);
        case 512: return launch_enhance<ALG, 9>(a, n_items, stream);
        case 1024: return launch_enhance<ALG, 10>(a, n_items, stream);
        default: return launch_enhance<ALG, 11>(a, n_items, stream);
    }
}
static real alg_eps(int algorithm) { return algorithm == CSE_ALG_MMSE ? R(1e-12) : R(1e-10); }

// items [item0, item0 + n_items) of the utterance-major (utt, param) product
static int enhance_items(const void* tables, int algorithm, const void* Y, const void* N, int noise_tv, int length,
                         int n_fft, int hop, const cse_params* params, int n_params, int item0, int n_items,
                         void* out, void* stream, const int* item_list = nullptr) {
    EnhanceArgs a;
    a.item_list = item_list;
    a.T = (const CseTables*)tables; a.Y = (const real2*)Y; a.N = (const real*)N; a.params = params;
    a.out = (real*)out; a.noise_tv = noise_tv; a.L = length; a.hop = hop;
    a.n_frames = cse_num_frames(length, hop); a.n_params = n_params; a.item0 = item0; a.eps = alg_eps(algorithm);
    a.hop_shift = (hop & (hop - 1)) == 0 ? cse_ilog2(hop) : -1;
    switch (algorithm) {
        case CSE_ALG_SS: return dispatch_enhance<0>(a, n_fft, n_items, stream);
        case CSE_ALG_WIENER: return dispatch_enhance<1>(a, n_fft, n_items, stream);
        case CSE_ALG_MMSE: return dispatch_enhance<2>(a, n_fft, n_items, stream);
        case CSE_ALG_OMLSA: return dispatch_enhance<3>(a, n_fft, n_items, stream);
    }
    return fail(CSE_EINVAL, "unknown algorithm %d", algorithm);
}

extern "C" int cse_enhance(const void* tables, int algorithm, const void* Y, const void* N, int noise_tv, int n_utts,
                int length, int n_fft, int hop, const cse_params* params, int n_params, void* out, void* stream) {
    CSE_REQUIRE(tables && Y && N && params && out, "NULL argument");
    CSE_REQUIRE(valid_nfft(n_fft), "n_fft %d not in {256,512,1024,2048}", n_fft);
    CSE_REQUIRE(hop > 0 && hop <= n_fft / 2 && hop % 2 == 0, "hop %d must be even and <= n_fft/2", hop);
    CSE_REQUIRE(n_utts > 0 && n_params > 0 && length > n_fft / 2, "bad sizes");
    CSE_REQUIRE(algorithm >= 0 && algorithm <= 3, "unknown algorithm %d", algorithm);
    return enhance_items(tables, algorithm, Y, N, noise_tv, length, n_fft, hop, params, n_params, 0,
                         n_utts * n_params, out, stream);
}

extern "C" int cse_gamma(const void* Y, const void* N, int noise_tv, int n_utts, int length, int n_fft, int hop,
                         double noise_mu, double eps, void* G, void* stream) {
    CSE_REQUIRE(Y && N && G, "NULL argument");
    CSE_REQUIRE(valid_nfft(n_fft) && hop > 0 && n_utts > 0 && length > n_fft / 2, "bad sizes");
    CSE_REQUIRE(noise_tv == 0 || noise_tv == 1, "noise_tv must be 0 (static PSD) or 1 (time-varying PSD)");
    const int nf = cse_num_frames(length, hop), nb = n_fft / 2 + 1, nbp = cse_nbp(n_fft);
    real mu = (real)noise_mu;
    if (noise_mu >= 0.0) mu = r_clip(mu, R(0), R(0.9999));               // mmse.py:51, advanced_mmse.py:61
    CSE_LAUNCH(gamma_kernel, dim3((nbp + 127) / 128, n_utts), 128, 0, stream, (const real2*)Y, (const real*)N, noise_tv, nf, nb,
               nbp, mu, (real)eps, (real*)G);
    return check_launch("gamma_kernel");
}

extern "C" int cse_gamma_groups(int n_utts, int length, int n_fft, const cse_gamma_group* groups, int n_groups, void* stream) {
    CSE_REQUIRE(groups, "NULL argument");
    CSE_REQUIRE(valid_nfft(n_fft) && n_utts > 0 && n_groups > 0 && length > n_fft / 2, "bad sizes");
    const int nb = n_fft / 2 + 1, nbp = cse_nbp(n_fft);
    for (int k = 0; k < n_groups; ++k) {
        CSE_REQUIRE(groups[k].Y && groups[k].N && groups[k].G, "NULL argument in group %d", k);
        CSE_REQUIRE(groups[k].hop > 0, "bad hop in group %d", k);
        CSE_REQUIRE(groups[k].noise_tv == 0 || groups[k].noise_tv == 1, "noise_tv must be 0 (static PSD) or 1 (time-varying PSD)");
    }
    for (int g0 = 0; g0 < n_groups; g0 += CSE_GAMMA_MAX_GROUPS) {
        GammaGroupsArgs ga;
        ga.nb = nb; ga.nbp = nbp;
        const int n = n_groups - g0 < CSE_GAMMA_MAX_GROUPS ? n_groups - g0 : CSE_GAMMA_MAX_GROUPS;
        for (int k = 0; k < CSE_GAMMA_MAX_GROUPS; ++k) {
            const cse_gamma_group& src = groups[g0 + (k < n ? k : 0)];
            GammaGroup& g = ga.g[k];
            g.Y = (const real2*)src.Y; g.N = (const real*)src.N; g.out = (real*)src.G; g.noise_tv = src.noise_tv;
            g.nf = cse_num_frames(length, src.hop);
            real mu = (real)src.noise_mu;
            if (src.noise_mu >= 0.0) mu = r_clip(mu, R(0), R(0.9999));           // mmse.py:51, advanced_mmse.py:61
            g.mu = mu; g.eps = (real)src.eps;
        }
        CSE_LAUNCH(gamma_groups_kernel, dim3((nbp + 127) / 128, n_utts, n), 128, 0, stream, ga);
        const int rc = check_launch("gamma_groups_kernel");
        if (rc != CSE_OK) return rc;
    }
    return CSE_OK;
}

extern "C" int cse_enhance_list(const void* tables, int algorithm, const void* Y, const void* N, int noise_tv, int length,
                                int n_fft, int hop, const cse_params* params, int n_params, const int* items, int n_items,
                                void* out, void* stream) {
    CSE_REQUIRE(tables && Y && N && params && out && items, "NULL argument");
    CSE_REQUIRE(valid_nfft(n_fft), "n_fft %d not in {256,512,1024,2048}", n_fft);
    CSE_REQUIRE(hop > 0 && hop <= n_fft / 2 && hop % 2 == 0, "hop %d must be even and <= n_fft/2", hop);
    CSE_REQUIRE(n_params > 0 && n_items > 0 && length > n_fft / 2, "bad sizes");
    CSE_REQUIRE(algorithm >= 0 && algorithm <= 3, "unknown algorithm %d", algorithm);
    return enhance_items(tables, algorithm, Y, N, noise_tv, length, n_fft, hop, params, n_params, 0, n_items, out, stream, items);
}

// ---- grouped launches (several noise-PSD groups of one instantiation)
template <int ALG, int LOG2N>
static int launch_enhance_groups(const EnhanceGroupsArgs& ga, int n_blocks, void* stream) {
    typedef EnhanceCfg<LOG2N> C;
    constexpr bool STAGED = LOG2N <= CSE_ENH_STAGED_MAX_LOG2N && LOG2N >= CSE_ENH_STAGED_MIN_LOG2N;
    size_t smem = 0;
    for (int k = 0; k < ga.n_groups; ++k) {
        const size_t sk = enhance_smem_bytes<LOG2N>(ga.g[k].hop, ga.noise_tv, STAGED);
        smem = sk > smem ? sk : smem;
    }
    if (ga.noise_tv == 2) {
        if constexpr (ALG == 0) return fail(CSE_EINVAL, "spectral subtraction takes the noise PSD, not the a-posteriori SNR (noise_tv 2)");
        else {
            auto kfn = enhance_groups_kernel<ALG, LOG2N, STAGED, true, true>;
            CSE_SMEM_OPT_IN(kfn, enhance_smem_bytes<LOG2N>(C::NFFT / 2, 1, STAGED));
            CSE_LAUNCH(kfn, n_blocks, C::NT, smem, stream, ga);
        }
    } else if (ga.noise_tv) {
        auto kfn = enhance_groups_kernel<ALG, LOG2N, STAGED, true>;
        CSE_SMEM_OPT_IN(kfn, enhance_smem_bytes<LOG2N>(C::NFFT / 2, 1, STAGED));
        CSE_LAUNCH(kfn, n_blocks, C::NT, smem, stream, ga);
    } else {
        auto kfn = enhance_groups_kernel<ALG, LOG2N, STAGED, false>;
        CSE_SMEM_OPT_IN(kfn, enhance_smem_bytes<LOG2N>(C::NFFT / 2, 0, STAGED));
        CSE_LAUNCH(kfn, n_blocks, C::NT, smem, stream, ga);
    }
    return check_launch("enhance_groups_kernel");
}
template <int ALG>
static int dispatch_enhance_groups(const EnhanceGroupsArgs& ga, int n_fft, int n_blocks, void* stream) {
    switch (n_fft) {
        case 256: return launch_enhance_groups<ALG, 8>(ga, n_blocks, stream);
        case 512: return launch_enhance_groups<ALG, 9>(ga, n_blocks, stream);
        case 1024: return launch_enhance_groups<ALG, 10>(ga, n_blocks, stream);
        default: return launch_enhance_groups<ALG, 11>(ga, n_blocks, stream);
    }
}

extern "C" int cse_enhance_groups(const void* tables, int algorithm, int noise_tv, int n_utts, int length, int n_fft,
                                  const cse_enhance_group* groups, int n_groups, void* stream) {
    CSE_REQUIRE(tables && groups, "NULL argument");
    CSE_REQUIRE(valid_nfft(n_fft), "n_fft %d not in {256,512,1024,2048}", n_fft);
    CSE_REQUIRE(n_utts > 0 && n_groups > 0 && length > n_fft / 2, "bad sizes");
    CSE_REQUIRE(algorithm >= 0 && algorithm <= 3, "unknown algorithm %d", algorithm);
    CSE_REQUIRE(noise_tv >= 0 && noise_tv <= 2, "noise_tv %d not in {0,1,2}", noise_tv);
    for (int k = 0; k < n_groups; ++k) {
        CSE_REQUIRE(groups[k].Y && groups[k].N && groups[k].params && groups[k].out, "NULL argument in group %d", k);
        CSE_REQUIRE(groups[k].hop > 0 && groups[k].hop <= n_fft / 2 && groups[k].hop % 2 == 0, "hop %d must be even and <= n_fft/2", groups[k].hop);
        CSE_REQUIRE(groups[k].n_params > 0, "group %d has no parameter rows", k);
    }
    for (int g0 = 0; g0 < n_groups; g0 += CSE_ENH_MAX_GROUPS) {       // as many launches as the descriptor table needs
        EnhanceGroupsArgs ga;
        ga.T = (const CseTables*)tables; ga.noise_tv = noise_tv; ga.L = length; ga.eps = alg_eps(algorithm);
        ga.n_groups = n_groups - g0 < CSE_ENH_MAX_GROUPS ? n_groups - g0 : CSE_ENH_MAX_GROUPS;
        int blocks = 0;
        for (int k = 0; k < ga.n_groups; ++k) {
            const cse_enhance_group& src = groups[g0 + k];
            EnhanceGroup& g = ga.g[k];
            g.Y = (const real2*)src.Y; g.N = (const real*)src.N; g.params = src.params; g.out = (real*)src.out;
            g.hop = src.hop; g.n_frames = cse_num_frames(length, src.hop); g.n_params = src.n_params;
            g.hop_shift = (src.hop & (src.hop - 1)) == 0 ? cse_ilog2(src.hop) : -1;
            g.first_block = blocks; g.reserved = 0;
            blocks += n_utts * src.n_params;
        }
        for (int k = ga.n_groups; k < CSE_ENH_MAX_GROUPS; ++k) ga.g[k] = ga.g[0];
        int rc = CSE_OK;
        switch (algorithm) {
            case CSE_ALG_SS: rc = dispatch_enhance_groups<0>(ga, n_fft, blocks, stream); break;
            case CSE_ALG_WIENER: rc = dispatch_enhance_groups<1>(ga, n_fft, blocks, stream); break;
            case CSE_ALG_MMSE: rc = dispatch_enhance_groups<2>(ga, n_fft, blocks, stream); break;
            default: rc = dispatch_enhance_groups<3>(ga, n_fft, blocks, stream); break;
        }
        if (rc != CSE_OK) return rc;
    }
    return CSE_OK;
}

#include "cse_lib_noise_score.inl"
