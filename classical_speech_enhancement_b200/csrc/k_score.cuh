// K5: per-candidate post-processing and scoring on the device.
//
// Restates, for sr = 16 kHz:
//   align_to_reference   Code/speech_enhancement_comparison.py:38-69   (align_kernel)
//   finalize_enhanced    Code/speech_enhancement_comparison.py:92-106,175 (lag shift, length
//                        match, finite check, clip - applied on the fly by xhat())
//   calculate_stoi       Code/evaluation_metrics.py:30-36 -> pystoi.stoi(extended=False)
//   calculate_snr        Code/evaluation_metrics.py:39-58
//
// Alignment: the reference takes the argmax over |lag| <= 1600 of a 32 000-sample full
// cross-correlation.  Only 3201 lags are needed, so the signal is cut into blocks of
// B = 8192 - 2*1600 samples; two blocks are packed as real/imaginary parts of one 8192-point
// complex FFT, multiplied with the cached spectrum of the matching (2*1600 longer) clean blocks
// packed the same way, accumulated over block pairs in registers and inverse-transformed once:
//   Re IFFT( sum_pairs Q_pair * conj(Z_pair) )[k + 1600] = sum_n clean0[n + k] * sig[n].
// The forward transforms are DIF (bit-reversed output), the inverse is DIT (bit-reversed
// input), so no reordering pass exists; the cached clean spectra are stored in the order in which
// the candidate kernel's last forward pass consumes them (see align_kernel).  Mean removal of the
// candidate is applied afterwards (correlation is linear).
//
// STOI: polyphase 16k->10k resampling through a de-interleaved shared-memory tile, frames
// gathered through the clean signal's VAD list (cached per utterance), 512-point real FFTs as
// batched 256-point complex DIF FFTs, third-octave band energies, and the 30-frame segment
// correlation against cached clean-side statistics.
#pragma once
#include "cse_fft.cuh"
#include "cse_resampler_taps.h"

// Resampler taps of the candidate kernel as compile-time-indexed constant-bank values: with the row
// loop fully unrolled every tap arrives through a uniform register of its packed FFMA2, and the
// structurally zero taps (each output phase only reaches 116-117 of the 136 tile rows) are
// skipped at compile time.
#ifdef CSE_EMU
static const real c_rs5[CSE_RS_ROWS * 5] = {CSE_RS_TAP_VALUES5};
#else
// five taps per row, unpadded: the packed FFMA2 form takes its taps from uniform registers, and 20 taps
// of four rows arrive in five 16-byte uniform loads
__device__ __constant__ __align__(16) real c_rs5[CSE_RS_ROWS * 5] = {CSE_RS_TAP_VALUES5};
#endif

#define CSE_SR 16000
#define CSE_CORR_LOG2P 13
#define CSE_CORR_P (1 << CSE_CORR_LOG2P)
#define CSE_MAXLAG 1600
#define CSE_CORR_B (CSE_CORR_P - 2 * CSE_MAXLAG)
#define CSE_CORR_SECONDS 2
#define CSE_NBANDS 15
#define CSE_NSEG 30
#define CSE_STOI_T 8                   // STOI frames transformed per batch
#define CSE_STOI_K0 7                  // first / one-past-last FFT bin used by the bands
#define CSE_STOI_K1 219
#define CSE_RS_A 256                   // resampler tile: input groups of 8 samples per pass
#define CSE_RS_GROWS CSE_RS_ROWS

struct CleanHeader {
    int K;              // frames kept by the VAD
    int J;              // number of 30-frame segments (0 -> pystoi returns 1e-5)
    int aligned;        // correlation window >= 256 samples
    int reserved;
    double energy;      // sum(clean^2)
    double ref_mean;    // mean(clean[:Nc])
    double pad[4];
};

struct ScoreGeom {
    int L, Nc, maxlag, nblocks, npairs, n10, nfr, nfrm, jmax;
    size_t off_kept, off_need, off_hbmap, off_xtob, off_seg, off_rsum, off_q, bytes;
};

// reals per candidate of the score workspace: the 15 x (K-1) band envelopes of the STOI kernel
CSE_HD size_t score_row_reals(int nfrm) { return ((size_t)CSE_NBANDS * (nfrm + 1) + 15) & ~(size_t)15; }

static inline ScoreGeom score_geom(int L) {
    ScoreGeom g;
    g.L = L;
    g.Nc = L < CSE_CORR_SECONDS * CSE_SR ? L : CSE_CORR_SECONDS * CSE_SR;
    g.maxlag = g.Nc - 1 < CSE_MAXLAG ? g.Nc - 1 : CSE_MAXLAG;
    g.nblocks = (g.Nc + CSE_CORR_B - 1) / CSE_CORR_B;
    g.npairs = (g.nblocks + 1) / 2;
    g.n10 = (5 * L + 7) / 8;
    g.nfr = g.n10 > 256 ? (g.n10 - 256 + 127) / 128 : 0;
    g.nfrm = g.nfr > 1 ? g.nfr - 1 : 0;
    g.jmax = g.nfrm >= CSE_NSEG ? g.nfrm - CSE_NSEG + 1 : 0;
    auto up = [](size_t x) { return (x + 63) & ~(size_t)63; };
    size_t o = up(sizeof(CleanHeader));
    g.off_kept = o; o = up(o + sizeof(int) * (size_t)(g.nfr + 1));
    g.off_need = o; o = up(o + (size_t)((g.n10 + 4) / 5 + 8));       // one byte per resampler group of 5 outputs
    g.off_hbmap = o; o = up(o + sizeof(int) * (size_t)3 * (g.nfr + 2));   // per 128-sample hop-block: (jA, jB, #kept below)
    g.off_xtob = o; o = up(o + sizeof(real) * (size_t)CSE_NBANDS * (g.nfrm + 1));
    g.off_seg = o; o = up(o + sizeof(real) * (size_t)3 * CSE_NBANDS * (g.jmax + 1));
    g.off_rsum = o; o = up(o + sizeof(real) * (size_t)(2 * CSE_MAXLAG + 1));
    g.off_q = o; o = up(o + sizeof(real2) * (size_t)g.npairs * CSE_CORR_P);
    g.bytes = o;
    return g;
}

// ---------------------------------------------------------------- resampler taps (host)
static double cse_i0_series(double x) {          // modified Bessel I0 for the Kaiser window
    double s = 1.0, term = 1.0;
    const double q = 0.25 * x * x;
    for (int k = 1; k < 64; ++k) { term *= q / ((double)k * k); s += term; if (term < 1e-20 * s) break; }
    return s;
}
// pystoi 0.4.1 utils.resample_oct(x, 10000, 16000): h = kaiser(581, beta) * 2*5*fc*sinc(2 fc t),
// fc = 1/16, beta = 0.1102 (60 - 8.7); scipy.resample_poly(x, 5, 8, window=h/sum(h)) multiplies by
// up=5 and centres it: y[m] = sum_j x[j] h5[8m + 290 - 5j].  Tile row jj = i + 64 (i = j - 8a).
static void cse_fill_resampler(CseTables* t) {
    const double PI = 3.14159265358979323846;
    const int half = 290, n = 2 * half + 1;
    const double beta = 0.1102 * (60.0 - 8.7), fc = 1.0 / 16.0;
    static double h[581];
    double sum = 0.0;
    for (int i = 0; i < n; ++i) {
        const double tt = i - half;
        const double xs = 2.0 * fc * tt;
        const double sinc = tt == 0 ? 1.0 : sin(PI * xs) / (PI * xs);
        const double r = 2.0 * i / (n - 1) - 1.0;
        const double kaiser = cse_i0_series(beta * sqrt(1.0 - r * r)) / cse_i0_series(beta);
        h[i] = kaiser * 2.0 * 5.0 * fc * sinc;
        sum += h[i];
    }
    const int edges[16] = {7, 9, 11, 14, 17, 22, 27, 34, 43, 55, 69, 87, 109, 138, 174, 219};
    for (int i = 0; i < 16; ++i) t->stoi_edges[i] = edges[i];
    for (int jj = 0; jj < CSE_RS_GROWS; ++jj)
        for (int p = 0; p < 8; ++p) {
            const int idx = 8 * p + 610 - 5 * jj;
            const double v = (p < 5 && idx >= 0 && idx <= 580) ? 5.0 * h[idx] / sum : 0.0;
            t->rs[jj][p] = (real)v;
            t->rs_d[jj][p] = v;
        }
}

// ---------------------------------------------------------------- finalize on the fly
// enhanced -> align shift -> match_length -> clip(-1, 1)  (speech_enhancement_comparison.py:62-67,29-36,105)
CSE_D real xraw(const real* __restrict__ sig, int i, int lag, int L) {
    const int j = i - lag;
    return (j < 0 || j >= L || i < 0 || i >= L) ? R(0) : sig[j];
}
CSE_D real xhat(const real* __restrict__ sig, int i, int lag, int L, bool finalize) {
    const real v = xraw(sig, i, lag, L);
    return finalize ? r_clip(v, R(-1), R(1)) : v;
}

struct ScoreArgs {
    const CseTables* T;
    const real* wav;        // [n_items][L] candidates (or the clean signals in the prepare pass)
    const real* clean;      // [U][L]
    unsigned char* cache;   // [U] records of geom.bytes
    cse_score_t* scores;    // [n_items]
    real* y10;              // workspace [n_items][score_row_reals(g)]: band-envelope scratch rows of the STOI kernel
    int* lagflags;          // workspace [n_items][2]
    int per_utt, finalize, item0;   // block b scores global item item0 + b; wav / workspace are chunk-local
    ScoreGeom g;
};

// ---------------------------------------------------------------- alignment
template <bool CLEAN>
__global__ void __launch_bounds__(512, 2) align_kernel(ScoreArgs a) {
    constexpr int P = CSE_CORR_P, NT = 512, PER = P / NT, M = CSE_MAXLAG, B = CSE_CORR_B;
    CSE_DYN_SMEM(smem_raw);
    real2* z = reinterpret_cast<real2*>(smem_raw);                              // CSE_FFT_STRIDE(P)
    real2* tws = z + CSE_FFT_STRIDE(P);                                         // radix-8 pass twiddles (the radix-2 pass reads the global table, unit stride)
    double* scratch = reinterpret_cast<double*>(tws + FftTwLayout<CSE_CORR_LOG2P, false>::SIZE + 1);                   // 40 doubles
    const int tid = threadIdx.x, li = blockIdx.x, item = a.item0 + li;
    const int u = CLEAN ? item : item / a.per_utt;
    const ScoreGeom& g = a.g;
    const int L = g.L, Nc = g.Nc;
    const real* __restrict__ sig = a.wav + (size_t)li * L;
    unsigned char* rec = a.cache + (size_t)u * g.bytes;
    CleanHeader* hdr = reinterpret_cast<CleanHeader*>(rec);
    real2* Q = reinterpret_cast<real2*>(rec + g.off_q);
    real* rsum = reinterpret_cast<real*>(rec + g.off_rsum);

    // mean over the correlation window; a non-finite sample there makes the reference's whole
    // correlation NaN, and np.argmax of an all-NaN array is index 0, i.e. lag = -max_lag
    load_pass_twiddles<CSE_CORR_LOG2P, false>(tws, a.T->tw, tid, NT);
    double s = 0.0, e2 = 0.0;
    int bad = 0;
    int i_first = tid;
#if !defined(CSE_FP64)
    if (!CLEAN && (reinterpret_cast<size_t>(sig) & 15) == 0) {
        // candidate side: 16-byte loads, eight samples summed in fp32 before they join the double
        // accumulator (one conversion and one double add per eight samples; a non-finite sample makes
        // its group sum non-finite, which is then examined sample by sample)
        const float4* __restrict__ v4 = reinterpret_cast<const float4*>(sig);
        const int n4 = Nc >> 2;
        for (int g4 = tid; g4 < n4; g4 += 2 * NT) {
            const float4 q1 = v4[g4];
            const float4 q2 = g4 + NT < n4 ? v4[g4 + NT] : make_float4(0.f, 0.f, 0.f, 0.f);
            const float gs = ((q1.x + q1.y) + (q1.z + q1.w)) + ((q2.x + q2.y) + (q2.z + q2.w));
            s += (double)gs;
            if (!r_finite(gs)) {
                if (!r_finite(q1.x) || !r_finite(q1.y) || !r_finite(q1.z) || !r_finite(q1.w) ||
                    !r_finite(q2.x) || !r_finite(q2.y) || !r_finite(q2.z) || !r_finite(q2.w)) bad = 1;
            }
        }
        i_first = 4 * n4 + tid;
    }
#endif
    for (int i = i_first; i < (CLEAN ? L : Nc); i += NT) {
        const real v = sig[i];
        if (i < Nc) { s += (double)v; if (!r_finite(v)) bad = 1; }
        if (CLEAN) e2 += (double)v * (double)v;
    }
    const double total = block_sum<double>(s, scratch);
    const double nbad = block_sum<double>((double)bad, scratch);
    const double mean = total / (double)Nc;
    const bool aligned = Nc >= 256 && (a.finalize || CLEAN);
    if (CLEAN) {
        const double en = block_sum<double>(e2, scratch);
        if (tid == 0) { hdr->energy = en; hdr->ref_mean = mean; hdr->aligned = Nc >= 256; }
    } else if (!aligned || nbad > 0.0) {
        if (tid == 0) {
            a.lagflags[2 * li] = aligned ? -g.maxlag : 0;
            a.lagflags[2 * li + 1] = aligned ? (CSE_FLAG_VALID | CSE_FLAG_ALIGNED) : CSE_FLAG_VALID;
        }
        return;
    }
    if (CLEAN && Nc < 256) return;

    const real meanr = (real)mean;
    if (CLEAN) {
        for (int pair = 0; pair < g.npairs; ++pair) {
            const int b1 = 2 * pair, b2 = b1 + 1;
            for (int idx = tid; idx < P; idx += NT) {
                // clean blocks, 2*M samples longer, zero outside [0, Nc)
                const int n1 = b1 * B - M + idx, n2 = b2 * B - M + idx;
                const real re = (n1 >= 0 && n1 < Nc) ? sig[n1] - meanr : R(0);
                const real im = (b2 < g.nblocks && n2 >= 0 && n2 < Nc) ? sig[n2] - meanr : R(0);
                z[SIDX(idx)] = mk2(re, im);
            }
            __syncthreads();
            fft_dif<CSE_CORR_LOG2P, false, -1>(z, 1, 0, tws, tid, NT, a.T->tw);
            // stored in the order the candidate side consumes it: thread t of group half gi multiplies
            // outputs (t + gi NT) 8 + m, m < 8, so slot ((gi 8 + m) NT + t) keeps a warp's reads contiguous
            for (int idx = tid; idx < P; idx += NT) {
                const int grp = idx >> 3, m = idx & 7, gi = grp / NT, t = grp - gi * NT;
                Q[(size_t)pair * P + (gi * 8 + m) * NT + t] = z[SIDX(idx)];
            }
            __syncthreads();
        }
        // rsum[k + M] = sum of clean0[m] over the m that lag k overlaps with (for the mean correction)
        double t0 = 0.0;
        for (int i = tid; i < Nc; i += NT) t0 += (double)(sig[i] - meanr);
        const double tot0 = block_sum<double>(t0, scratch);
        for (int kk = tid; kk <= 2 * M; kk += NT) {
            const int k = kk - M;
            double part = 0.0;
            if (k >= 0) { for (int m = 0; m < k && m < Nc; ++m) part += (double)(sig[m] - meanr); }
            else { for (int m = Nc + k < 0 ? 0 : Nc + k; m < Nc; ++m) part += (double)(sig[m] - meanr); }
            rsum[kk] = (real)(tot0 - part);
        }
        return;
    }

    // Candidate side.  The 8192-point transform (13 stages = radix 2 x four radix-8 passes) never runs
    // a pass that only moves data: the first radix-2 stage is applied while the block pair is loaded
    // (its upper inputs are zero beyond B - P/2), the last forward pass multiplies its register outputs
    // by the cached clean spectrum and accumulates over block pairs in registers, the first inverse
    // pass starts from those registers, and the last inverse stage is evaluated only for the
    // 2 maxlag + 1 lags that are searched.
    static_assert(CSE_CORR_LOG2P == 13 && PER == 16, "pass structure below: 2 x 8 x 8 x 8 x 8, two radix-8 groups per thread");
    constexpr int HP = P / 2, ZH = SIDX(HP);
    const real2* __restrict__ twg = a.T->tw;                  // flat W_8192^k, k < 4096
    real2 acc[2][8];
#pragma unroll
    for (int gi = 0; gi < 2; ++gi)
#pragma unroll
        for (int m = 0; m < 8; ++m) acc[gi][m] = mk2(R(0), R(0));
    for (int pair = 0; pair < g.npairs; ++pair) {
        const int b1 = 2 * pair, b2 = b1 + 1;
        const real* __restrict__ s1 = sig + b1 * B;
        const real* __restrict__ s2 = sig + b2 * B;
        const int r1 = Nc - b1 * B, r2 = b2 < g.nblocks ? Nc - b2 * B : 0;      // samples left in each block
        {   // the next pair's samples start their way from DRAM now (one 128-byte line per thread)
            const int pf = (b1 + 2) * B + tid * (128 / (int)sizeof(real));
            if (pf < Nc && pf < (b1 + 4) * B) cse_prefetch_l2(sig + pf);
        }
        if (r1 >= B && r2 >= B) {                 // both blocks lie inside the window (uniform): no per-sample bounds
#pragma unroll
            for (int k = 0; k < HP / NT; ++k) {
                const int i = tid + k * NT, ih = i + HP;
                const real2 lo = mk2(s1[i], s2[i]);
                real2 hi = mk2(R(0), R(0));
                if (k * NT < B - HP && ih < B) hi = mk2(s1[ih], s2[ih]);
                real2* pz = z + SIDX(i);
                pz[0] = cadd(lo, hi);
                pz[ZH] = cmul(csub(lo, hi), twg[i]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < HP / NT; ++k) {
                const int i = tid + k * NT, ih = i + HP;
                const real2 lo = mk2(i < r1 ? s1[i] : R(0), i < r2 ? s2[i] : R(0));            // i < HP < B
                real2 hi = mk2(R(0), R(0));
                if (k * NT < B - HP && ih < B) hi = mk2(ih < r1 ? s1[ih] : R(0), ih < r2 ? s2[ih] : R(0));
                real2* pz = z + SIDX(i);
                pz[0] = cadd(lo, hi);
                pz[ZH] = cmul(csub(lo, hi), twg[i]);
            }
        }
        __syncthreads();
        dif_pass<CSE_CORR_LOG2P, 3, false, (HP >> 1), 0>(z, 1, 0, tws, tid, NT); __syncthreads();
        dif_pass<CSE_CORR_LOG2P, 3, false, (HP >> 4), 0>(z, 1, 0, tws, tid, NT); __syncthreads();
        dif_pass<CSE_CORR_LOG2P, 3, false, (HP >> 7), 0>(z, 1, 0, tws, tid, NT); __syncthreads();
#pragma unroll
        for (int gi = 0; gi < 2; ++gi) {
            const int grp = tid + gi * NT;
            const real2* pz = z + SIDX(grp * 8);
            real2 v[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) v[m] = pz[m];
            bfly8_dif<false, true>(v, mk2(R(1), R(0)), mk2(R(1), R(0)), mk2(R(1), R(0)));
            const real2* __restrict__ q = Q + (size_t)pair * P + gi * 8 * NT + tid;
#pragma unroll
            for (int m = 0; m < 8; ++m) acc[gi][m] = cadd(acc[gi][m], cmulc(q[m * NT], v[m]));
        }
        __syncthreads();
    }
#pragma unroll
    for (int gi = 0; gi < 2; ++gi) {
        const int grp = tid + gi * NT;
        real2* pz = z + SIDX(grp * 8);
        bfly8_dit<true, true>(acc[gi], mk2(R(1), R(0)), mk2(R(1), R(0)), mk2(R(1), R(0)));
#pragma unroll
        for (int m = 0; m < 8; ++m) pz[m] = acc[gi][m];
    }
    __syncthreads();
    dit_pass<CSE_CORR_LOG2P, 3, true, 8, 0>(z, 1, 0, tws, tid, NT); __syncthreads();
    dit_pass<CSE_CORR_LOG2P, 3, true, 64, 0>(z, 1, 0, tws, tid, NT); __syncthreads();
    dit_pass<CSE_CORR_LOG2P, 3, true, 512, 0>(z, 1, 0, tws, tid, NT); __syncthreads();
    // first maximum over k = -maxlag .. maxlag; output k + M of the last radix-2 stage is
    // z[k + M] + z[k + M + P/2] conj(W_P^(k + M)), of which only the real part is needed
    const real invP = R(1) / (real)P;
    real best = -cse_inf();
    int bestk = 0x7fffffff;
    for (int k = -g.maxlag + tid; k <= g.maxlag; k += NT) {
        const int i = k + M;
        const real2 z0 = z[SIDX(i)], z1 = z[SIDX(i) + ZH], w = twg[i];
        const real c = (z0.x + (z1.x * w.x + z1.y * w.y)) * invP - meanr * rsum[i];
        if (c > best) { best = c; bestk = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const real ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int ok = __shfl_xor_sync(0xffffffffu, bestk, o);
        if (ob > best || (ob == best && ok < bestk)) { best = ob; bestk = ok; }
    }
    real* sb = reinterpret_cast<real*>(scratch);
    int* sk = reinterpret_cast<int*>(scratch + 20);
    __syncthreads();
    if ((tid & 31) == 0) { sb[tid >> 5] = best; sk[tid >> 5] = bestk; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < NT / 32; ++w)
            if (sb[w] > best || (sb[w] == best && sk[w] < bestk)) { best = sb[w]; bestk = sk[w]; }
        if (bestk == 0x7fffffff) bestk = -g.maxlag;      // all-NaN: np.argmax returns index 0
        a.lagflags[2 * li] = bestk;
        a.lagflags[2 * li + 1] = CSE_FLAG_VALID | CSE_FLAG_ALIGNED;
    }
}

// ---------------------------------------------------------------- resampler
// One pass of the clean-side (double precision) resampler: outputs y10[5a + p], a in [a0, a0 + A), from
// the de-interleaved tile xs[c][ap] (c = sample index mod 8, AP = A + 17 columns), taps from the table G.
// The candidate side has its own packed form in k_stoi_stream.cuh.
template <class TA, class TG, class TO>
CSE_D void resample_pass(const real* __restrict__ sig, int L, int lag, bool finalize, const TG* __restrict__ G /*[136][8]*/,
                         TA* xs, int a0, TO* __restrict__ y10, int n10, int tid, int nth,
                         const unsigned char* __restrict__ need = nullptr) {
    constexpr int A = CSE_RS_A, AP = A + 17;
    const int j0 = 8 * a0 - 64;
    for (int jj = tid; jj < 8 * AP; jj += nth) xs[(jj & 7) * AP + (jj >> 3)] = (TA)xhat(sig, j0 + jj, lag, L, finalize);
    __syncthreads();
    for (int al = tid; al < A; al += nth) {
        const int a = a0 + al;
        if (5 * a < n10 && (need == nullptr || need[a])) {
            TA acc[5] = {0, 0, 0, 0, 0};
            {
                for (int o = 0; o < 17; ++o) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const TA x = xs[c * AP + al + o];
                        const TG* g = G + (size_t)(8 * o + c) * 8;
#pragma unroll
                        for (int p = 0; p < 5; ++p) acc[p] += x * (TA)g[p];
                    }
                }
            }
#pragma unroll
            for (int p = 0; p < 5; ++p) if (5 * a + p < n10) y10[5 * a + p] = (TO)acc[p];
        }
    }
    __syncthreads();
}

#define CSE_RS_A2 512                  // candidate-side resampler tile (k_stoi_stream.cuh): groups of 5 outputs per pass

// clean-side resampling + VAD (pystoi remove_silent_frames, mask from the clean signal only)
// grid U, block 256; y10 (real) goes to the workspace for clean_stoi.
__global__ void __launch_bounds__(256) clean_vad_kernel(ScoreArgs a, double* __restrict__ y10d, double* __restrict__ energies) {
    CSE_DYN_SMEM(smem_raw);
    double* xs = reinterpret_cast<double*>(smem_raw);             // 8 * (A + 17)
    double* scratch = xs + 8 * (CSE_RS_A + 17);                   // 40
    const int tid = threadIdx.x, u = blockIdx.x;
    const ScoreGeom& g = a.g;
    const real* sig = a.wav + (size_t)u * g.L;
    double* yd = y10d + (size_t)u * g.n10;
    unsigned char* rec = a.cache + (size_t)u * g.bytes;
    CleanHeader* hdr = reinterpret_cast<CleanHeader*>(rec);
    int* kept = reinterpret_cast<int*>(rec + g.off_kept);
    const int na = (g.n10 + 4) / 5;
    for (int a0 = 0; a0 < na; a0 += CSE_RS_A)
        resample_pass<double, double, double>(sig, g.L, 0, false, &a.T->rs_d[0][0], xs, a0, yd, g.n10, tid, 256);
    __threadfence_block();
    __syncthreads();
    // frame energies in dB: 20 log10(||w * frame|| + EPS)
    const double EPS = 2.220446049250313e-16;
    double* en = energies + (size_t)u * (g.nfr + 1);
    double emax = -1e300;
    for (int f = tid; f < g.nfr; f += 256) {
        double e = 0.0;
        for (int n = 0; n < 256; ++n) { const double v = (double)a.T->stoi_win[n] * yd[128 * f + n]; e += v * v; }
        e = 20.0 * log10(sqrt(e) + EPS);
        en[f] = e;
        emax = e > emax ? e : emax;
    }
    // block max
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, emax, o); emax = t > emax ? t : emax; }
    __syncthreads();
    if ((tid & 31) == 0) scratch[tid >> 5] = emax;
    __syncthreads();
    emax = scratch[0];
    for (int w = 1; w < 8; ++w) emax = scratch[w] > emax ? scratch[w] : emax;
    __syncthreads();
    // mask = (max - 40 dB - e) < 0, then ordered compaction into the kept-frame list
    if (tid == 0) {
        int K = 0;
        for (int f = 0; f < g.nfr; ++f) if ((emax - 40.0 - en[f]) < 0.0) kept[K++] = f;
        hdr->K = K;
        hdr->J = (K - 1 >= CSE_NSEG) ? (K - 1) - CSE_NSEG + 1 : 0;
    }
    __threadfence_block();
    __syncthreads();
    // resampler groups a (outputs 5a..5a+4) that land in a kept frame's samples [128k, 128k + 256)
    {
        unsigned char* need = rec + g.off_need;
        const int K = hdr->K;
        for (int a5 = tid; a5 < na; a5 += 256) {
            const int lo = 5 * a5, hi = lo + 4;           // output sample range of this group
            // frames f with 128f <= hi and 128f + 255 >= lo
            int f0 = (lo - 255 + 127) / 128;
            if (f0 < 0) f0 = 0;
            const int f1 = hi / 128;
            unsigned char nd = 0;
            for (int j = 0; j < K && !nd; ++j) { const int f = kept[j]; if (f >= f0 && f <= f1) nd = 1; else if (f > f1) break; }
            need[a5] = nd;
        }
        // hop-block map for the streaming STOI kernel.  Hop-block hb = y10[128 hb, 128 hb + 128) is the first
        // half of kept frame hb (index jA in the kept list) and the second half of kept frame hb-1 (which
        // starts the overlap-added block jB = index(hb-1) + 1); count = kept frames with frame index < hb.
        int* hbmap = reinterpret_cast<int*>(rec + g.off_hbmap);
        for (int hb = tid; hb < g.nfr + 2; hb += 256) {
            int ja = -1, jb = -1, cnt = 0;
            for (int j = 0; j < K; ++j) {
                const int f = kept[j];
                if (f < hb) ++cnt;
                if (f == hb) ja = j;
                if (f == hb - 1) jb = j + 1;
            }
            hbmap[3 * hb] = ja; hbmap[3 * hb + 1] = jb; hbmap[3 * hb + 2] = cnt;
        }
    }
}

// ---------------------------------------------------------------- clean-side STOI caches
// Band envelopes of the clean signal and, per (segment, band), its norm, mean and 1/(centred norm +
// EPS) into the per-utterance cache (the candidate side is k_stoi_stream.cuh).  y10d: the clean
// signal at 10 kHz (double, from clean_vad_kernel).
__global__ void __launch_bounds__(256, 4) clean_stoi_kernel(ScoreArgs a, const double* __restrict__ y10d) {
    constexpr int T = CSE_STOI_T, BST = CSE_FFT_STRIDE(256), NK = CSE_STOI_K1 - CSE_STOI_K0, NT = 256;
    CSE_DYN_SMEM(smem_raw);
    const ScoreGeom& g = a.g;
    const int tid = threadIdx.x, u = blockIdx.x;
    unsigned char* rec = a.cache + (size_t)u * g.bytes;
    const CleanHeader* hdr = reinterpret_cast<const CleanHeader*>(rec);
    const int* __restrict__ kept = reinterpret_cast<const int*>(rec + g.off_kept);
    real* xtob_c = reinterpret_cast<real*>(rec + g.off_xtob);
    real* seg_c = reinterpret_cast<real*>(rec + g.off_seg);
    const int K = hdr->K, Kf = K > 0 ? K - 1 : 0, J = hdr->J;

    real2* fbuf = reinterpret_cast<real2*>(smem_raw);                         // T * BST
    real* pw = reinterpret_cast<real*>(fbuf + T * BST);                       // T * NK
    real* ytob = pw + T * NK;                                                 // 15 * Kf
    real* w_s = ytob + ((CSE_NBANDS * (g.nfrm + 1) + 3) & ~3);                // 256: np.hanning(258)[1:-1] (16-byte aligned)
    real2* tws = reinterpret_cast<real2*>(w_s + 256);                         // per-pass twiddles of the 256-point FFT
    int* kept_s = reinterpret_cast<int*>(tws + 160);                          // nfr + 2 kept-frame indices

    for (int i = tid; i < 256; i += NT) w_s[i] = a.T->stoi_win[i];
    load_pass_twiddles<8, true>(tws, a.T->tw, tid, NT);
    for (int i = tid; i < K; i += NT) kept_s[i] = kept[i];
    __syncthreads();
    if (Kf < CSE_NSEG) return;
    const int* __restrict__ edges = a.T->stoi_edges;
    const double* ydu = y10d + (size_t)u * g.n10;

    for (int m0 = 0; m0 < Kf; m0 += T) {
        // frame m = w .* (B_m ++ B_{m+1}), B_j[n] = w[n] y10[128 kept[j] + n] + w[128+n] y10[128 kept[j-1] + 128 + n]
        // (the overlap-added blocks of the silence-removed signal)
        for (int idx = tid; idx < T * 128; idx += NT) {
            const int f = idx >> 7, mm = idx & 127, m = m0 + f;
            const int j = mm < 64 ? m : m + 1, nn = (2 * mm) & 127, n = 2 * mm;
            real2 v = mk2(R(0), R(0));
            if (m < Kf) {
                const int k1 = 128 * kept_s[j] + nn;
                real b0 = w_s[nn] * (real)ydu[k1], b1 = w_s[nn + 1] * (real)ydu[k1 + 1];
                if (j > 0) {
                    const int k0 = 128 * kept_s[j - 1] + 128 + nn;
                    b0 = r_fma(w_s[nn + 128], (real)ydu[k0], b0);
                    b1 = r_fma(w_s[nn + 129], (real)ydu[k0 + 1], b1);
                }
                v = mk2(w_s[n] * b0, w_s[n + 1] * b1);
            }
            fbuf[f * BST + SIDX(mm)] = v;
            fbuf[f * BST + SIDX(mm + 128)] = mk2(R(0), R(0));
        }
        __syncthreads();
        fft_dif<8, false, 0>(fbuf, T, BST, tws, tid, NT);
        for (int idx = tid; idx < T * NK; idx += NT) {
            const int f = idx / NK, k = CSE_STOI_K0 + (idx - f * NK);
            const real2* zf = fbuf + f * BST;
            const real2 z0 = zf[SIDX(brev_n(k, 8))], z1 = zf[SIDX(brev_n(256 - k, 8))];
            const real2 E = mk2(R(0.5) * (z0.x + z1.x), R(0.5) * (z0.y - z1.y));
            const real2 O = mk2(R(0.5) * (z0.y + z1.y), R(-0.5) * (z0.x - z1.x));
            const real2 X = cadd(E, cmul(O, tw_load(a.T->tw, k * (CSE_TW_N / 512))));
            pw[idx] = X.x * X.x + X.y * X.y;
        }
        __syncthreads();
        for (int idx = tid; idx < T * CSE_NBANDS; idx += NT) {
            const int f = idx / CSE_NBANDS, b = idx - f * CSE_NBANDS, m = m0 + f;
            if (m < Kf) {
                real sacc = R(0);
                for (int k = edges[b]; k < edges[b + 1]; ++k) sacc += pw[f * NK + k - CSE_STOI_K0];
                ytob[b * Kf + m] = r_sqrt(sacc);
            }
        }
        __syncthreads();
    }
    const real EPS = R(2.220446049250313e-16);
    for (int i = tid; i < CSE_NBANDS * Kf; i += NT) xtob_c[i] = ytob[i];
    for (int idx = tid; idx < J * CSE_NBANDS; idx += NT) {
        const int b = idx / J, j = idx - b * J;        // (band, segment), segment fastest
        const real* x = ytob + b * Kf + j;
        real s1 = R(0), s2 = R(0);
        for (int n = 0; n < CSE_NSEG; ++n) { s1 += x[n]; s2 = r_fma(x[n], x[n], s2); }
        const real mean = s1 / R(CSE_NSEG);
        real c2 = R(0);
        for (int n = 0; n < CSE_NSEG; ++n) { const real d = x[n] - mean; c2 = r_fma(d, d, c2); }
        seg_c[(size_t)0 * J * CSE_NBANDS + idx] = r_sqrt(s2);
        seg_c[(size_t)1 * J * CSE_NBANDS + idx] = mean;
        seg_c[(size_t)2 * J * CSE_NBANDS + idx] = R(1) / (r_sqrt(c2) + EPS);
    }
}
