// ------------------------------------------------------------------ K2 noise estimators
static int next_pow2(int n) { int p = 32; while (p < n) p <<= 1; return p; }   // >= one element per lane of the sorting warp

// (k, effective percentile) of Code/noise_estimation.py:29-41; short-signal rule :194-195,226-232
static void quiet_rule(int nf, double percentile, int* k, double* pct, double* floor_rel) {
    if (nf < 5) { *k = nf; *pct = nf < 2 ? 50.0 : 25.0; *floor_rel = 0.0; return; }
    int min_frames = 10;
    if (nf < 30) {
        min_frames = std::max(2, nf / 4);
        const int target = std::max(3, (int)(nf * 0.15));
        percentile = std::min(50.0, 100.0 * target / nf);
    }
    int kk = std::max(min_frames, (int)ceil(nf * (percentile / 100.0)));
    kk = std::min(kk, std::max(1, (int)ceil(nf * 0.30)));
    *k = std::min(kk, nf);
    *pct = percentile;
    *floor_rel = 0.02;
}

extern "C" size_t cse_noise_workspace_bytes(int n_utts, int n_frames, int n_fft) {
    (void)n_fft;
    return (size_t)n_utts * n_frames * (sizeof(double) + sizeof(int)) + 64;
}

extern "C" int cse_noise_percentile(const void* P, int n_utts, int n_frames, int n_fft, double percentile, double eps,
                                    void* N, void* workspace, size_t workspace_bytes, void* stream) {
    CSE_REQUIRE(P && N && workspace, "NULL argument");
    CSE_REQUIRE(valid_nfft(n_fft) && n_utts > 0 && n_frames > 0, "bad sizes");
    CSE_REQUIRE(n_frames <= 8192, "n_frames %d > 8192 unsupported", n_frames);
    if (workspace_bytes < cse_noise_workspace_bytes(n_utts, n_frames, n_fft)) return fail(CSE_EWORKSPACE, "noise workspace too small");
    const int nb = n_fft / 2 + 1, nbp = cse_nbp(n_fft);
    int k; double pct, floor_rel;
    quiet_rule(n_frames, percentile, &k, &pct, &floor_rel);
    // numpy's virtual index for method="linear": n*q + (alpha + q*(1 - alpha - beta)) - 1, alpha = beta = 1
    const double q = pct / 100.0;
    const double vi = k * q + (1.0 + q * (1.0 - 1.0 - 1.0)) - 1.0;
    int lo = (int)floor(vi);
    double frac = vi - lo;
    if (lo < 0) { lo = 0; frac = 0; }
    if (lo >= k - 1) { lo = k - 1; frac = 0; }
    double* energy = (double*)workspace;
    int* quiet = (int*)(energy + (size_t)n_utts * n_frames);
    CSE_LAUNCH(frame_logenergy_kernel, dim3(n_frames, n_utts), 128, 40 * sizeof(double), stream,
               (const real*)P, n_frames, nb, nbp, (real)eps, energy);
    CSE_LAUNCH(quiet_select_kernel, n_utts, 256, 0, stream, (const double*)energy, n_frames, k, quiet);
    const int n_pad = next_pow2(n_frames), k_pad = next_pow2(k);
    const int bpc = std::max(1, std::min(8, 8192 / n_pad));
    const size_t smem = (size_t)bpc * n_pad * sizeof(real);
    if (smem > 48 * 1024) cudaFuncSetAttribute(percentile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    CSE_LAUNCH(percentile_kernel, dim3((nb + bpc - 1) / bpc, n_utts), 32 * bpc, smem, stream, (const real*)P, n_frames,
               nb, nbp, n_pad, (const int*)quiet, k, k_pad, lo, (real)frac, (real)floor_rel, (real)eps, (real*)N);
    return check_launch("noise_percentile");
}

extern "C" int cse_noise_mintrack(const void* P, int n_utts, int n_frames, int n_fft, double eps, void* N,
                                  void* workspace, size_t workspace_bytes, void* stream) {
    (void)workspace; (void)workspace_bytes;
    CSE_REQUIRE(P && N, "NULL argument");
    CSE_REQUIRE(valid_nfft(n_fft) && n_utts > 0 && n_frames > 0, "bad sizes");
    CSE_REQUIRE(n_frames <= 4096, "n_frames %d > 4096 unsupported", n_frames);
    if (n_frames < 5)
        return fail(CSE_EUNSUPPORTED, "n_frames < 5: the reference's short-signal rule applies, use cse_noise_percentile");
    const int nb = n_fft / 2 + 1, nbp = cse_nbp(n_fft);
    const double a = std::max(0.8, std::min(0.95, 1.0 - 5.0 / n_frames));     // noise_estimation.py:73-75
    int w = std::min(std::max(3, 50), n_frames);                               // :97-99
    if (w % 2 == 0) w += 1;
    const int n_pad = next_pow2(n_frames);
    const int bpc = std::max(1, std::min(8, 4096 / n_pad));
    const size_t smem = (size_t)2 * bpc * n_pad * sizeof(real);
    if (smem > 48 * 1024) cudaFuncSetAttribute(mintrack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    CSE_LAUNCH(mintrack_kernel, dim3((nb + bpc - 1) / bpc, n_utts), 32 * bpc, smem, stream, (const real*)P, n_frames, nb,
               nbp, n_pad, (real)a, w / 2, R(0.01), (real)eps, (real*)N);
    return check_launch("noise_mintrack");
}

// ------------------------------------------------------------------ K5 scoring
static size_t align_smem() { return (size_t)(CSE_FFT_STRIDE(CSE_CORR_P) + FftTwLayout<CSE_CORR_LOG2P, false>::SIZE + 2) * sizeof(real2) + 48 * sizeof(double); }
static size_t stoi_smem(const ScoreGeom& g) {      // clean_stoi_kernel
    return sizeof(real) * ((size_t)CSE_STOI_T * CSE_FFT_STRIDE(256) * 2 + (size_t)CSE_STOI_T * (CSE_STOI_K1 - CSE_STOI_K0) +
                           (size_t)CSE_NBANDS * (g.nfrm + 1) + 4 + 256 + 320) + sizeof(int) * (size_t)(g.nfr + 2);
}
static size_t stoi_stream_smem() {
    return 40 * sizeof(double) + sizeof(real) * ((size_t)CSE_STOI_T * CSE_FFT_STRIDE(256) * 2 + 5 * CSE_RS_A2 +
                                                 (size_t)CSE_STOI_RB * 128 + 256 + 320 +
                                                 (size_t)3 * (CSE_STOI_K1 - CSE_STOI_K0) + 32 + 128);   // + split twiddles (2 reals) and bin offsets per band bin, run descriptors, hop-block rows of the tile
}
static size_t up64(size_t x) { return (x + 63) & ~(size_t)63; }
static int check_sr(int sr) {
    if (sr != CSE_SR) return fail(CSE_EUNSUPPORTED, "scoring is built for sr=16000 (got %d): the reference resamples to 16 kHz before this path", sr);
    return CSE_OK;
}
#define CSE_MAX_SMEM (227 * 1024)

extern "C" int cse_max_score_length(int sr) {
    if (sr != CSE_SR) return 0;
    int lo = 1, hi = 1 << 24;                      // stoi_smem grows monotonically with the length
    while (lo < hi) {
        const int mid = lo + (hi - lo + 1) / 2;
        if (stoi_smem(score_geom(mid)) <= CSE_MAX_SMEM) lo = mid; else hi = mid - 1;
    }
    return lo;
}
extern "C" size_t cse_clean_cache_bytes(int length, int sr) { (void)sr; return length > 0 ? score_geom(length).bytes : 0; }
extern "C" size_t cse_clean_workspace_bytes(int n_utts, int length, int sr) {
    (void)sr;
    if (length <= 0 || n_utts <= 0) return 0;
    const ScoreGeom g = score_geom(length);
    return up64((size_t)n_utts * g.n10 * sizeof(double)) + up64((size_t)n_utts * (g.nfr + 1) * sizeof(double));
}

extern "C" int cse_prepare_clean(const void* tables, const void* clean, int n_utts, int length, int sr, void* cache,
                                 void* workspace, size_t workspace_bytes, void* stream) {
    CSE_REQUIRE(tables && clean && cache && workspace, "NULL argument");
    CSE_REQUIRE(n_utts > 0 && length > 0, "bad sizes");
    if (int rc = check_sr(sr)) return rc;
    if (workspace_bytes < cse_clean_workspace_bytes(n_utts, length, sr)) return fail(CSE_EWORKSPACE, "clean workspace too small");
    ScoreArgs a;
    memset(&a, 0, sizeof(a));
    a.T = (const CseTables*)tables; a.wav = (const real*)clean; a.clean = (const real*)clean;
    a.cache = (unsigned char*)cache; a.per_utt = 1; a.finalize = 0; a.item0 = 0; a.g = score_geom(length);
    if (stoi_smem(a.g) > CSE_MAX_SMEM) return fail(CSE_EUNSUPPORTED, "utterance too long for the STOI kernel's shared memory (%d samples)", length);
    double* y10d = (double*)workspace;
    double* energies = (double*)((unsigned char*)workspace + up64((size_t)n_utts * a.g.n10 * sizeof(double)));
    auto ka = align_kernel<true>;
    CSE_SMEM_OPT_IN(ka, align_smem());
    CSE_LAUNCH(ka, n_utts, 512, align_smem(), stream, a);
    const size_t vsm = (size_t)(8 * (CSE_RS_A + 17) + 40) * sizeof(double);
    CSE_LAUNCH(clean_vad_kernel, n_utts, 256, vsm, stream, a, y10d, energies);
    auto ks = clean_stoi_kernel;
    cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stoi_smem(a.g));
    CSE_LAUNCH(ks, n_utts, 256, stoi_smem(a.g), stream, a, (const double*)y10d);
    return check_launch("prepare_clean");
}

extern "C" size_t cse_score_workspace_bytes(int n_items, int length, int sr) {
    (void)sr;
    if (length <= 0 || n_items <= 0) return 0;
    const ScoreGeom g = score_geom(length);
    return up64((size_t)n_items * score_row_reals(g.nfrm) * sizeof(real)) + up64((size_t)n_items * 2 * sizeof(int));
}

// scores items [item0, item0 + n_items); wav holds only those items.  which: 1 = alignment kernel,
// 2 = SNR/STOI kernel, 3 = both (the stoi kernel consumes the lag / flags the alignment kernel left
// in the workspace).
static int score_items(const void* tables, const void* wav, int item0, int n_items, int per_utt, int length,
                       const void* clean, const void* cache, int finalize, cse_score_t* scores, void* workspace,
                       void* stream, int which = 3) {
    ScoreArgs a;
    memset(&a, 0, sizeof(a));
    a.T = (const CseTables*)tables; a.wav = (const real*)wav; a.clean = (const real*)clean;
    a.cache = (unsigned char*)cache; a.scores = scores; a.per_utt = per_utt; a.finalize = finalize; a.item0 = item0;
    a.g = score_geom(length);
    if (stoi_smem(a.g) > CSE_MAX_SMEM) return fail(CSE_EUNSUPPORTED, "utterance too long for the STOI kernel's shared memory (%d samples)", length);
    a.y10 = (real*)workspace;
    a.lagflags = (int*)((unsigned char*)workspace + up64((size_t)n_items * score_row_reals(a.g.nfrm) * sizeof(real)));
    if (which & 1) {
        auto ka = align_kernel<false>;
        CSE_SMEM_OPT_IN(ka, align_smem());
        CSE_LAUNCH(ka, n_items, 512, align_smem(), stream, a);
    }
    if (which & 2) {
        auto ks = stoi_stream_kernel;
        CSE_SMEM_OPT_IN(ks, stoi_stream_smem());
        CSE_LAUNCH(ks, n_items, 256, stoi_stream_smem(), stream, a);
    }
    return check_launch("score");
}

extern "C" int cse_score(const void* tables, const void* wav, int n_utts, int per_utt, int length, int sr,
                         const void* clean, const void* cache, int finalize, cse_score_t* scores, void* workspace,
                         size_t workspace_bytes, void* stream) {
    CSE_REQUIRE(tables && wav && clean && cache && scores && workspace, "NULL argument");
    CSE_REQUIRE(n_utts > 0 && per_utt > 0 && length > 0, "bad sizes");
    if (int rc = check_sr(sr)) return rc;
    const int n_items = n_utts * per_utt;
    if (workspace_bytes < cse_score_workspace_bytes(n_items, length, sr)) return fail(CSE_EWORKSPACE, "score workspace too small");
    return score_items(tables, wav, 0, n_items, per_utt, length, clean, cache, finalize, scores, workspace, stream);
}

extern "C" size_t cse_sweep_workspace_bytes(int chunk_items, int length, int sr) {
    if (chunk_items <= 0 || length <= 0) return 0;
    return up64((size_t)chunk_items * length * sizeof(real)) + cse_score_workspace_bytes(chunk_items, length, sr);
}

extern "C" int cse_sweep(const void* tables, int algorithm, const void* Y, const void* N, int noise_tv, int n_utts,
                         int length, int n_fft, int hop, const cse_params* params, int n_params, int sr,
                         const void* clean, const void* cache, cse_score_t* scores, int chunk_items, void* workspace,
                         size_t workspace_bytes, void* stream) {
    CSE_REQUIRE(tables && Y && N && params && clean && cache && scores && workspace, "NULL argument");
    CSE_REQUIRE(valid_nfft(n_fft), "n_fft %d not in {256,512,1024,2048}", n_fft);
    CSE_REQUIRE(hop > 0 && hop <= n_fft / 2 && hop % 2 == 0, "hop %d must be even and <= n_fft/2", hop);
    CSE_REQUIRE(n_utts > 0 && n_params > 0 && length > n_fft / 2 && chunk_items > 0, "bad sizes");
    CSE_REQUIRE(algorithm >= 0 && algorithm <= 3, "unknown algorithm %d", algorithm);
    if (int rc = check_sr(sr)) return rc;
    if (workspace_bytes < cse_sweep_workspace_bytes(chunk_items, length, sr)) return fail(CSE_EWORKSPACE, "sweep workspace too small");
    void* wavs = workspace;
    void* score_ws = (unsigned char*)workspace + up64((size_t)chunk_items * length * sizeof(real));
    const int total = n_utts * n_params;
    for (int i0 = 0; i0 < total; i0 += chunk_items) {
        const int n = std::min(chunk_items, total - i0);
        if (int rc = enhance_items(tables, algorithm, Y, N, noise_tv, length, n_fft, hop, params, n_params, i0, n, wavs, stream)) return rc;
        if (int rc = score_items(tables, wavs, i0, n, n_params, length, clean, cache, 1, scores, score_ws, stream)) return rc;
    }
    return CSE_OK;
}

extern "C" int cse_enhance_items(const void* tables, int algorithm, const void* Y, const void* N, int noise_tv,
                                 int length, int n_fft, int hop, const cse_params* params, int n_params, int item0,
                                 int n_items, void* out, void* stream) {
    CSE_REQUIRE(tables && Y && N && params && out, "NULL argument");
    CSE_REQUIRE(valid_nfft(n_fft), "n_fft %d not in {256,512,1024,2048}", n_fft);
    CSE_REQUIRE(hop > 0 && hop <= n_fft / 2 && hop % 2 == 0, "hop %d must be even and <= n_fft/2", hop);
    CSE_REQUIRE(n_params > 0 && n_items > 0 && item0 >= 0 && length > n_fft / 2, "bad sizes");
    CSE_REQUIRE(algorithm >= 0 && algorithm <= 3, "unknown algorithm %d", algorithm);
    return enhance_items(tables, algorithm, Y, N, noise_tv, length, n_fft, hop, params, n_params, item0, n_items, out, stream);
}

extern "C" int cse_score_items(const void* tables, const void* wav, int item0, int n_items, int per_utt, int length,
                               int sr, const void* clean, const void* cache, int finalize, cse_score_t* scores,
                               void* workspace, size_t workspace_bytes, void* stream) {
    CSE_REQUIRE(tables && wav && clean && cache && scores && workspace, "NULL argument");
    CSE_REQUIRE(n_items > 0 && item0 >= 0 && per_utt > 0 && length > 0, "bad sizes");
    if (int rc = check_sr(sr)) return rc;
    if (workspace_bytes < cse_score_workspace_bytes(n_items, length, sr)) return fail(CSE_EWORKSPACE, "score workspace too small");
    return score_items(tables, wav, item0, n_items, per_utt, length, clean, cache, finalize, scores, workspace, stream);
}

// The two halves of cse_score_items as separate launches (same arguments), so that a caller can
// bracket each kernel with its own events: cse_align_items first, then cse_stoi_items.
extern "C" int cse_align_items(const void* tables, const void* wav, int item0, int n_items, int per_utt, int length,
                               int sr, const void* clean, const void* cache, int finalize, cse_score_t* scores,
                               void* workspace, size_t workspace_bytes, void* stream) {
    CSE_REQUIRE(tables && wav && clean && cache && scores && workspace, "NULL argument");
    CSE_REQUIRE(n_items > 0 && item0 >= 0 && per_utt > 0 && length > 0, "bad sizes");
    if (int rc = check_sr(sr)) return rc;
    if (workspace_bytes < cse_score_workspace_bytes(n_items, length, sr)) return fail(CSE_EWORKSPACE, "score workspace too small");
    return score_items(tables, wav, item0, n_items, per_utt, length, clean, cache, finalize, scores, workspace, stream, 1);
}
extern "C" int cse_stoi_items(const void* tables, const void* wav, int item0, int n_items, int per_utt, int length,
                              int sr, const void* clean, const void* cache, int finalize, cse_score_t* scores,
                              void* workspace, size_t workspace_bytes, void* stream) {
    CSE_REQUIRE(tables && wav && clean && cache && scores && workspace, "NULL argument");
    CSE_REQUIRE(n_items > 0 && item0 >= 0 && per_utt > 0 && length > 0, "bad sizes");
    if (int rc = check_sr(sr)) return rc;
    if (workspace_bytes < cse_score_workspace_bytes(n_items, length, sr)) return fail(CSE_EWORKSPACE, "score workspace too small");
    return score_items(tables, wav, item0, n_items, per_utt, length, clean, cache, finalize, scores, workspace, stream, 2);
}

// ------------------------------------------------------------------ nominal score table
// Broadcast of the unique candidates' scores to every nominal grid point (duplicates through dead
// parameters get the identical record, as the reference would compute them), on the device:
// out[u][p] = unique[base[p] + u * stride[p]]  (16-byte records; one thread per record).
__global__ void __launch_bounds__(256) expand_scores_kernel(const cse_score_t* __restrict__ uniq, const int* __restrict__ base,
                                                            const int* __restrict__ stride, int n_utts, int n_points,
                                                            cse_score_t* __restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x, u = blockIdx.y;
    if (p < n_points && u < n_utts) out[(size_t)u * n_points + p] = uniq[(size_t)base[p] + (size_t)u * stride[p]];
}

extern "C" int cse_expand_scores(const cse_score_t* unique_scores, const int* base, const int* stride, int n_utts,
                                 int n_points, cse_score_t* out, void* stream) {
    CSE_REQUIRE(unique_scores && base && stride && out, "NULL argument");
    CSE_REQUIRE(n_utts > 0 && n_points > 0, "bad sizes");
    CSE_LAUNCH(expand_scores_kernel, dim3((n_points + 255) / 256, n_utts), 256, 0, stream, unique_scores, base, stride,
               n_utts, n_points, out);
    return check_launch("expand_scores");
}

// ------------------------------------------------------------------ K6 selection
extern "C" int cse_select_best(const cse_score_t* table, const double* pesq, int n_utts, int n_points,
                               cse_winner_t* winners, void* stream) {
    CSE_REQUIRE(table && winners, "NULL argument");
    CSE_REQUIRE(n_utts > 0 && n_points > 0, "bad sizes");
    CSE_LAUNCH(select_best_kernel, n_utts, 32 * CSE_SEL_CRITERIA, 0, stream, table, pesq, n_points, winners);
    return check_launch("select_best");
}

// ------------------------------------------------------------------ host-side probes for tests
// Evaluates the gain rules' special-function fits on the host (same code the kernels inline).
extern "C" int cse_debug_special(int which, const double* x, double* y, int n) {
    CSE_REQUIRE(x && y && n >= 0 && (which == 0 || which == 1), "bad argument");
    for (int i = 0; i < n; ++i) y[i] = which == 0 ? (double)cse_mmse_bessel_term((real)x[i]) : (double)cse_expint_e1((real)x[i]);
    return CSE_OK;
}
