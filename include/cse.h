/* libcse_sm100a - C ABI of the B200-native enhancement-and-scoring sweep.
 *
 * Drop-in boundary for the hot path of Katja39/Classical_Speech_Enhancement.  The reference
 * has no FFI layer: its seam is the Python call `algorithm_function(noisy, sr, **param_dict)`
 * plus the metric calls inside `optimize_parameters`
 * (Code/speech_enhancement_comparison.py:165,177-184).  The Python package
 * `classical_speech_enhancement_b200` re-exposes those entry points and binds the functions
 * below with ctypes (INTEGRATION.md shows the binding a reference maintainer would add).
 *
 * Conventions
 *  - plain C types only; every pointer marked [dev] is CALLER-OWNED device memory, contiguous,
 *    16-byte aligned (PyTorch tensors' data_ptr()); the library allocates nothing;
 *  - `real` is float in the default build and double when built with -DCSE_FP64
 *    (cse_dtype() reports 32 or 64; same symbols either way);
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing
 *    synchronises; results may be read after the caller synchronises the stream;
 *  - the device is the caller's current device; no mutable global state (the FFT twiddle /
 *    window / resampler tables are arguments the caller keeps in device memory, see
 *    cse_tables_bytes / cse_tables_init); re-entrant across streams and devices;
 *  - every function returns 0 (CSE_OK) or a negative CSE_E* code and never throws/exits;
 *    cse_last_error() returns a thread-local message for the last failure.
 *
 * Spectrogram layout: frame-major, bin-fastest, `[utt][frame][cse_bins_padded(n_fft)]`
 * (the transpose of librosa's (bins, frames)) so that a warp reading one frame is coalesced
 * and the frame march of the decision-directed recursion walks memory forwards.
 */
#ifndef CSE_H_
#define CSE_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSE_ABI_VERSION 1

enum cse_status {
    CSE_OK = 0,
    CSE_EINVAL = -1,        /* bad argument (message in cse_last_error) */
    CSE_ECUDA = -2,         /* a CUDA runtime call or launch failed */
    CSE_EUNSUPPORTED = -3,  /* valid in the reference but outside this build's range */
    CSE_EWORKSPACE = -4     /* caller-provided workspace too small */
};

/* Algorithms: reference table Code/speech_enhancement_comparison.py:395-401. */
enum cse_algorithm {
    CSE_ALG_SS = 0,      /* spectral_subtraction  Code/spectral_subtractor.py:6  */
    CSE_ALG_WIENER = 1,  /* wiener_filter         Code/wiener_filter.py:7        */
    CSE_ALG_MMSE = 2,    /* mmse                  Code/mmse.py:6                 */
    CSE_ALG_OMLSA = 3    /* advanced_mmse         Code/advanced_mmse.py:7        */
};

/* One grid point.  v[] by algorithm (names are the reference's keyword arguments):
 *   SS     : v0 alpha, v1 beta
 *   WIENER : v0 alpha, v1 gain_floor
 *   MMSE   : v0 alpha, v1 ksi_min, v2 gain_min, v3 gain_max, v4 noise_mu
 *   OMLSA  : v0 alpha, v1 ksi_min, v2 gain_floor, v3 noise_mu, v4 q, v5 v_max
 * noise_mu < 0 disables the recursive smoothing of a time-varying noise PSD (the reference
 * applies it only when the PSD has more than one frame and noise_method != "true_noise":
 * Code/mmse.py:48, Code/advanced_mmse.py:60). */
typedef struct cse_params { double v[8]; } cse_params;

/* One scored candidate (cse_dtype()==32 layout; the FP64 build uses double for the first two). */
#ifdef CSE_FP64
typedef struct cse_score_t { double stoi; double snr; int32_t lag; int32_t flags; } cse_score_t;
#else
typedef struct cse_score_t { float stoi; float snr; int32_t lag; int32_t flags; } cse_score_t;
#endif
#define CSE_FLAG_VALID 1      /* waveform finite -> candidate takes part in the selection     */
#define CSE_FLAG_ALIGNED 2    /* a lag was estimated (reference skips it when N < 256 samples) */
#define CSE_FLAG_SNR_INF 4    /* residual energy exactly 0: reference returns float('inf')     */
#define CSE_FLAG_STOI_SHORT 8 /* fewer than 30 STOI frames: pystoi returns 1e-5                */

int cse_abi_version(void);
int cse_dtype(void);                    /* 32 or 64 */
const char* cse_last_error(void);
/* Operating range of this build (checked by every entry point, CSE_EUNSUPPORTED / CSE_EINVAL otherwise;
 * the reference itself has no such limits):
 *   n_fft in {256, 512, 1024, 2048}; win_length = n_fft, periodic Hann, center=True, reflect padding;
 *   hop even, 0 < hop <= n_fft/2 (the overlap-add kernel owns sample pairs); length > n_fft/2;
 *   cse_noise_percentile: n_frames <= 8192;  cse_noise_mintrack: 5 <= n_frames <= 4096
 *     (about 65 s / 32 s of audio at hop 128; the per-bin series is sorted / scanned in one CTA);
 *   scoring (cse_prepare_clean, cse_score*, cse_sweep): sr = 16000 and length <= cse_max_score_length(sr)
 *     (about 38 s: the clean-side STOI kernel keeps the band envelopes of an utterance in shared memory). */
int cse_max_score_length(int sr);
int cse_bins_padded(int n_fft);         /* bin stride of the spectrogram layout */
int cse_num_frames(int length, int hop); /* 1 + length / hop (librosa center=True) */

/* Constant tables (twiddles, Hann windows, STOI resampler taps, band edges).  The caller
 * allocates cse_tables_bytes() bytes on each device once and passes the pointer to every call. */
size_t cse_tables_bytes(void);
int cse_tables_init(void* tables /*[dev]*/, void* stream);

/* STFT + PSD.  Replaces librosa.stft(...)+abs()**2 at Code/spectral_subtractor.py:25-26,
 * wiener_filter.py:35-37, mmse.py:29-32, advanced_mmse.py:39-40, noise_estimation.py:184-188.
 * wav [U][L]; if `minus` is non-NULL the transform is taken of (wav - minus) and the PSD is
 * floored at `psd_floor` (the oracle noise PSD of noise_estimation.py:128-147).
 * Y (interleaved re,im) [U][nf][nbp] may be NULL; P [U][nf][nbp] may be NULL. */
int cse_stft_psd(const void* tables, const void* wav, const void* minus, int n_utts, int length,
                 int n_fft, int hop, double psd_floor, void* Y, void* P, void* stream);

/* PercentileNoiseEstimator.estimate, Code/noise_estimation.py:20-56.  P [U][nf][nbp] ->
 * N [U][nbp].  Workspace: cse_noise_workspace_bytes(). */
size_t cse_noise_workspace_bytes(int n_utts, int n_frames, int n_fft);
int cse_noise_percentile(const void* P, int n_utts, int n_frames, int n_fft, double percentile,
                         double eps, void* N, void* workspace, size_t workspace_bytes, void* stream);

/* MinTrackingNoiseEstimator.estimate, Code/noise_estimation.py:64-99.  P -> N [U][nf][nbp]. */
int cse_noise_mintrack(const void* P, int n_utts, int n_frames, int n_fft, double eps, void* N,
                       void* workspace, size_t workspace_bytes, void* stream);

/* The candidate-invariant front of the Wiener / MMSE / Log-MMSE gain rules, once per (utterance, shape, noise PSD,
 * noise_mu): gamma[u][t][k] = max(|Y|^2 / N', eps), N' = max(N, eps) smoothed recursively over the frames with noise_mu
 * when noise_mu >= 0 and the PSD is time-varying (Code/wiener_filter.py:45,61, mmse.py:45-57,67,
 * advanced_mmse.py:57-66,76).  N as for cse_enhance (noise_tv 0 / 1); G [U][nf][nbp].  cse_enhance* take G in place
 * of N with noise_tv = 2 (and ignore the candidates' noise_mu): hundreds of candidates then share this work. */
int cse_gamma(const void* Y, const void* N, int noise_tv, int n_utts, int length, int n_fft, int hop,
              double noise_mu, double eps, void* G, void* stream);
/* Grouped form: several (noise PSD, noise_mu) combinations of ONE n_fft in one launch (a single pair needs about 36;
 * launched one by one their dependent frame chains cost more than the gain kernels).  `groups` is a HOST array. */
typedef struct cse_gamma_group {
    const void* Y;      /* [dev] spectrogram of (n_fft, hop) */
    const void* N;      /* [dev] noise PSD as noise_tv says (0 / 1) */
    void* G;            /* [dev] [U][nf][nbp] */
    int noise_tv;
    int hop;
    double noise_mu;    /* < 0: no smoothing */
    double eps;
} cse_gamma_group;
int cse_gamma_groups(int n_utts, int length, int n_fft, const cse_gamma_group* groups, int n_groups, void* stream);

/* Gain + ISTFT for n_utts x n_params candidates (utterance-major): out[(u*n_params+c)][L].
 * noise_tv: 0 -> N is [U][nbp] (static), 1 -> N is [U][nf][nbp] (time-varying), 2 -> N is the a-posteriori SNR
 * [U][nf][nbp] of cse_gamma (algorithms WIENER, MMSE, OMLSA only).
 * eps: the algorithm's epsilon (1e-10; 1e-12 for MMSE, Code/mmse.py:17).
 * Replaces the body of the four reference entry points after their STFT / noise_estimation
 * calls, including librosa.istft(..., length=L). */
int cse_enhance(const void* tables, int algorithm, const void* Y, const void* N, int noise_tv,
                int n_utts, int length, int n_fft, int hop, const cse_params* params /*[dev]*/,
                int n_params, void* out, void* stream);

/* Clean-side scoring caches (alignment spectra, STOI VAD mask / band envelopes / segment
 * statistics, signal energy), one record of cse_clean_cache_bytes() per utterance. */
size_t cse_clean_cache_bytes(int length, int sr);
size_t cse_clean_workspace_bytes(int n_utts, int length, int sr);
int cse_prepare_clean(const void* tables, const void* clean, int n_utts, int length, int sr,
                      void* cache, void* workspace, size_t workspace_bytes, void* stream);

/* finalize_enhanced + calculate_stoi + calculate_snr for n_utts x per_utt waveforms
 * (Code/speech_enhancement_comparison.py:171-184, evaluation_metrics.py:30-58).
 * finalize=0 scores the waveform as is (the baseline row, :116-118). */
size_t cse_score_workspace_bytes(int n_items, int length, int sr);
int cse_score(const void* tables, const void* wav, int n_utts, int per_utt, int length, int sr,
              const void* clean, const void* cache, int finalize, cse_score_t* scores,
              void* workspace, size_t workspace_bytes, void* stream);

/* The batched sweep: cse_enhance + cse_score in chunks of `chunk_items` candidates whose
 * waveforms live in the workspace (sized by cse_sweep_workspace_bytes).  scores [U*n_params]. */
size_t cse_sweep_workspace_bytes(int chunk_items, int length, int sr);
int cse_sweep(const void* tables, int algorithm, const void* Y, const void* N, int noise_tv,
              int n_utts, int length, int n_fft, int hop, const cse_params* params, int n_params,
              int sr, const void* clean, const void* cache, cse_score_t* scores,
              int chunk_items, void* workspace, size_t workspace_bytes, void* stream);

/* Chunk-level forms of the two halves of cse_sweep, for callers that drive the chunk loop
 * themselves (the Python engine does, to bracket each kernel family with CUDA events).
 * Items are indices into the utterance-major (utterance, param) product; `out` / `wav` hold only
 * the n_items waveforms of the chunk, `scores` is the base of the full [n_utts*n_params] table. */
int cse_enhance_items(const void* tables, int algorithm, const void* Y, const void* N, int noise_tv,
                      int length, int n_fft, int hop, const cse_params* params, int n_params,
                      int item0, int n_items, void* out, void* stream);
/* Sparse form: `items` [dev, int32] lists the n_items (utterance * n_params + param) candidates to compute, out[i][L]
 * receives candidate items[i].  Used to re-materialise the winners of a sweep (at most three per utterance and
 * algorithm, Code/speech_enhancement_comparison.py:188-216 keeps their waveforms) without recomputing the grid. */
int cse_enhance_list(const void* tables, int algorithm, const void* Y, const void* N, int noise_tv,
                     int length, int n_fft, int hop, const cse_params* params, int n_params,
                     const int* items, int n_items, void* out, void* stream);
/* Grouped form: several noise-PSD groups of ONE (algorithm, n_fft, noise_tv) in one launch - for small batches (one
 * pair is the reference's own call pattern, Code/speech_enhancement_comparison.py:108-252), whose groups are far
 * smaller than the GPU.  Group k computes all n_utts * n_params candidates of its (Y, N, params, hop) into its own
 * `out` [n_utts * n_params][length].  `groups` is a HOST array (copied into the launch parameters). */
typedef struct cse_enhance_group {
    const void* Y;            /* [dev] as for cse_enhance, shape of (n_fft, hop) */
    const void* N;            /* [dev] noise PSD or a-posteriori SNR, as noise_tv says */
    const cse_params* params; /* [dev] n_params rows */
    void* out;                /* [dev] [n_utts * n_params][length] */
    int hop;
    int n_params;
} cse_enhance_group;
int cse_enhance_groups(const void* tables, int algorithm, int noise_tv, int n_utts, int length, int n_fft,
                       const cse_enhance_group* groups, int n_groups, void* stream);
int cse_score_items(const void* tables, const void* wav, int item0, int n_items, int per_utt,
                    int length, int sr, const void* clean, const void* cache, int finalize,
                    cse_score_t* scores, void* workspace, size_t workspace_bytes, void* stream);

/* cse_score_items as its two kernels (alignment first, then SNR + STOI, which reads the lag and
 * flags the alignment left in the workspace): same arguments. */
int cse_align_items(const void* tables, const void* wav, int item0, int n_items, int per_utt,
                    int length, int sr, const void* clean, const void* cache, int finalize,
                    cse_score_t* scores, void* workspace, size_t workspace_bytes, void* stream);
int cse_stoi_items(const void* tables, const void* wav, int item0, int n_items, int per_utt,
                   int length, int sr, const void* clean, const void* cache, int finalize,
                   cse_score_t* scores, void* workspace, size_t workspace_bytes, void* stream);

/* Nominal score table from the unique candidates' records: out[u][p] = unique[base[p] + u*stride[p]]
 * (all [dev]; base / stride are int32 [n_points]).  Grid points that differ only in parameters the
 * reference ignores (noise_percentile under min_tracking, noise_mu under percentile) are computed
 * once and receive the identical record here, before the host-side selection scan. */
int cse_expand_scores(const cse_score_t* unique_scores, const int* base, const int* stride,
                      int n_utts, int n_points, cse_score_t* out, void* stream);

/* The three winners of optimize_parameters' sequential scan with hysteresis
 * (Code/speech_enhancement_comparison.py:186-216): a candidate replaces the running best of a criterion
 * only if it beats it by more than 1e-6 (stoi), 1e-3 (pesq), 1e-5 (balance), in grid order - an
 * order-dependent scan, not an argmax.  table [n_utts][n_points] is a nominal score table (grid order);
 * pesq [n_utts][n_points] doubles holds the host-side PESQ of every candidate (NaN = calculate_pesq
 * returned None, candidate skipped, :180-181) or is NULL (PESQ = 0.0 for every candidate: only the
 * `stoi` winner is then meaningful).  winners [n_utts][3] in the order stoi, pesq, balance; index -1 =
 * no candidate qualified.  Comparisons are done in double on the table's values, exactly as the host
 * scan (grid.select_best) does them. */
typedef struct cse_winner_t {
    int32_t index;      /* grid index of the winner, -1 if none */
    int32_t lag;        /* its alignment lag (finalize_enhanced) */
    int32_t flags;      /* its CSE_FLAG_* */
    int32_t reserved;
    double score;       /* the criterion's value */
    double stoi, pesq, snr; /* the winner's other metrics (snr = +inf under CSE_FLAG_SNR_INF) */
} cse_winner_t;
int cse_select_best(const cse_score_t* table, const double* pesq, int n_utts, int n_points,
                    cse_winner_t* winners, void* stream);

/* Host-side probe used by the tests: evaluates the special-function fits the gain kernels
 * inline (which = 0: exp(-v/2)[(1+v)I0(v/2)+vI1(v/2)] of Code/mmse.py:92-96; 1: E1(v) of
 * Code/advanced_mmse.py:103) at x[0..n) in the library's precision. */
int cse_debug_special(int which, const double* x, double* y, int n);

#ifdef __cplusplus
}
#endif
#endif /* CSE_H_ */
