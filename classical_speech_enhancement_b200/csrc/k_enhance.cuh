// K3+K4: per-candidate gain (frame march) fused with the inverse real FFT and the weighted
// overlap-add of librosa.istft.  One CTA per (utterance, grid point).
//
// Restates, after their STFT / noise_estimation calls:
//   ALG 0  spectral_subtraction  Code/spectral_subtractor.py:37-62
//   ALG 1  wiener_filter         Code/wiener_filter.py:47-94
//   ALG 2  mmse                  Code/mmse.py:48-118
//   ALG 3  advanced_mmse         Code/advanced_mmse.py:51-135
// and librosa.istft(S, hop_length, win_length=n_fft, window="hann", center=True, length=L).
//
// The decision-directed recursion is sequential over frames but independent per bin, so each
// thread carries (previous gain, previous a-posteriori SNR, smoothed noise PSD) for its bin pairs
// in registers (both bins of a pair as the two halves of packed FP32 instructions) and marches F
// frames per iteration.  The F gained spectra are turned into F packed
// half-size inverse FFTs in shared memory (F is chosen so that every thread has a radix-8
// butterfly per pass), windowed, overlap-added into a ring buffer, normalised by the window
// sum-of-squares and streamed out, F*hop finished samples per iteration.
//
// Thread layout: NTB = min(n_fft/4, 256) "pair" threads own the bin pairs (s, M-s), M = n_fft/2
// (slot 0 is DC + Nyquist); one extra warp carries the self-paired bin M/2 in its lane 0 so that
// no thread does extra work in front of the barrier; all NTB+32 threads share the FFT /
// overlap-add loops.  Y / N reads (coalesced, issued for the NEXT iteration right after the FFT
// so that they fly during the overlap-add) and waveform writes are coalesced.  Round-1 profiles and
// the changes they drove: profiles/r01*_ncu.md, DESIGN.md section 8.
#pragma once
#include "cse_fft.cuh"
#include "cse_special.cuh"

// ---------------------------------------------------------------- TMA bulk copy + mbarrier (sm_100a)
// The spectrogram layout is frame-major, so the F frames of one iteration are ONE contiguous block of
// F * nbp bins: a 1-D bulk copy (cp.async.bulk, the TMA engine; SASS UBLKCP) moves it into shared memory
// without a tensor map, and completion is signalled on an mbarrier (complete_tx::bytes).  One thread
// issues the copy; nobody spends issue slots, registers or address arithmetic on the tile.
#if defined(CSE_EMU)
// CPU emulation (tests): the issuing thread copies synchronously; the CTA barrier at the end of every
// iteration orders the copy before the next iteration's reads, so waiting is a no-op.
typedef unsigned long long cse_mbar_t;
CSE_D void cse_mbar_init(cse_mbar_t*, int) {}
CSE_D void cse_mbar_expect_tx(cse_mbar_t*, unsigned) {}
CSE_D void cse_bulk_g2s(void* dst, const void* src, unsigned bytes, cse_mbar_t*) { memcpy(dst, src, bytes); }
CSE_D void cse_mbar_wait(cse_mbar_t*, unsigned) {}
#else
typedef unsigned long long cse_mbar_t;
CSE_D unsigned cse_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
CSE_D void cse_mbar_init(cse_mbar_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cse_smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
CSE_D void cse_mbar_expect_tx(cse_mbar_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cse_smem_u32(bar)), "r"(bytes) : "memory");
}
CSE_D void cse_bulk_g2s(void* dst, const void* src, unsigned bytes, cse_mbar_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(cse_smem_u32(dst)), "l"(src), "r"(bytes), "r"(cse_smem_u32(bar)) : "memory");
}
CSE_D void cse_mbar_wait(cse_mbar_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "CSE_MBAR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra CSE_MBAR_DONE_%=;\n"
        "bra CSE_MBAR_WAIT_%=;\n"
        "CSE_MBAR_DONE_%=:\n"
        "}\n" ::"r"(cse_smem_u32(bar)), "r"(parity) : "memory");
}
#endif

// The gain rule for two bins at once (the thread's pair (s, M-s)): lane .x = bin a, lane .y = bin b,
// every FP add / mul / fma issued as one packed instruction for both bins; the decision-directed
// state (previous gain, previous a-posteriori SNR, smoothed noise PSD) is packed the same way.
struct GainState2 { real2 g_prev, gam_prev, nsm; };

// TV = false (static noise PSD: the percentile estimator, or any method on < 5 frames): the caller passes the
// PSD already floored at eps and its reciprocal, both loop-invariant - the per-frame floor and the MUFU
// reciprocal disappear from the frame march (same values, bit-identical results).
#ifndef CSE_GAIN_TRIM
#define CSE_GAIN_TRIM 1      // drop clamps that can only bind through rounding (see the comments at each site)
#endif
// GAM = true: `Nraw` IS the a-posteriori SNR gamma = max(|Y|^2 / N', eps), N' the floored (and, for the recursive
// algorithms, smoothed) noise PSD: everything in front of the decision-directed recursion is the same for all
// candidates of one (utterance, noise PSD, noise_mu) and is computed once per utterance by gamma_kernel.
template <int ALG, bool TV, bool GAM = false>
CSE_D void gain_pair(real2 Ya, real2 Yb, real2 Nraw, real2 rNstat, bool first, GainState2& st, const real* __restrict__ pv,
                     real eps, bool smooth, real2& Sa, real2& Sb) {
    static_assert(!(GAM && ALG == 0), "spectral subtraction needs the PSD itself");
    // |Y|^2 of the two bins with scalar FMAs: gathering (re, re') and (im, im') pairs for a packed form costs eight moves
    const real2 Pw = GAM ? mk2(R(0), R(0)) : mk2(r_fma(Ya.x, Ya.x, Ya.y * Ya.y), r_fma(Yb.x, Yb.x, Yb.y * Yb.y));
    real2 Nt = TV ? p_max(Nraw, p_set(eps)) : Nraw;
    if (ALG == 0) {
        const real2 Pc = p_max(p_fma(p_set(-pv[0]), Nt, Pw), p_mul(p_set(pv[1]), Nt));
        real g[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const real pw = e ? Pw.y : Pw.x, pc = e ? Pc.y : Pc.x;
            // (comparisons written so that a NaN power takes the first branch and propagates, as np.maximum does)
            g[e] = !(pw <= R(1e-30)) ? r_fsqrt(pc * r_rcp(pw)) : (pw > R(0) ? r_sqrt(pc) / r_sqrt(pw) : R(-1));
        }
        Sa = g[0] < R(0) ? mk2(r_sqrt(Pc.x), R(0)) : cscale(Ya, g[0]);
        Sb = g[1] < R(0) ? mk2(r_sqrt(Pc.y), R(0)) : cscale(Yb, g[1]);
        return;
    }
    if (ALG >= 2 && TV && !GAM && smooth) {
        const real mu = (ALG == 2) ? pv[4] : pv[3];
        if (!first) Nt = p_fma(p_set(mu), st.nsm, p_mul(p_set(pv[9]), Nt));
        st.nsm = Nt;
        // a convex combination of two values >= eps: the reference's second floor (mmse.py:57, advanced_mmse.py:66)
        // can only bind by one rounding of eps itself
        if (!CSE_GAIN_TRIM) Nt = p_max(Nt, p_set(eps));
    }
    const real2 gam = GAM ? Nraw : p_max(p_mul(Pw, TV ? p_rcp(Nt) : rNstat), p_set(eps));
    const real2 gm1 = p_add(gam, p_set(R(-1)));
    const real2 direct = p_max(gm1, p_set(R(0)));
    const real alpha = pv[0];
    const real2 rec = p_fma(p_set(alpha), p_mul(p_mul(st.g_prev, st.g_prev), st.gam_prev), p_mul(p_set(pv[8]), direct));
    real2 G;
    if (ALG == 1) {
        const real2 xi = p_max(first ? direct : rec, p_set(R(1e-10)));
        const real2 wg = p_mul(xi, p_rcp(p_add(xi, p_set(R(1)))));                 // xi / (1 + xi) < 1: the upper clip is idle
        G = CSE_GAIN_TRIM ? p_max(wg, p_set(pv[1])) : p_clip(wg, pv[1], R(1));
    } else {
        const real2 xi = p_max(first ? gm1 : rec, p_set(pv[1]));
        const real2 r = p_rcp(p_add(xi, p_set(R(1))));
        const real2 xr = p_mul(xi, r);
        if (ALG == 2) {
            const real2 v = p_clip(p_mul(xr, gam), eps, R(80));
            const real2 A = p_mul(p_mul(p_set(R(0.88622692545275801365)), p_sqrt(v)), p_rcp(p_add(gam, p_set(eps))));
            G = p_clip(p_mul(A, cse_mmse_bessel_term2(v)), pv[2], pv[3]);
        } else {
            const real gf = pv[2], q = pv[4], vmax = pv[5], lg2gf = pv[6];
            const real2 v = p_clip(p_mul(xr, gam), R(1e-12), vmax);
            const real2 enegv = p_exp2(p_mul(v, p_set(-CSE_LOG2E)));                  // exp(-v)
            const real2 lg2 = p_add(p_log2(xr), cse_half_e1_log2_2(v, enegv));
            // p = 1 / (1 + (1-q) / (q Lambda + eps)), Lambda = exp(v)/(1+xi); numerator and denominator
            // multiplied by exp(-v) so that one exponential serves both E1 and the presence probability
            const real2 A = p_fma(p_set(eps), enegv, p_mul(p_set(q), r));
            // A > 0 and the denominator exceeds it: p is in (0, 1] up to one rounding, the reference's clip is idle
            const real2 pr = p_mul(A, p_rcp(p_fma(p_set(pv[10]), enegv, A)));
            const real2 p = CSE_GAIN_TRIM ? pr : p_clip(pr, R(0), R(1));
            G = p_clip(p_exp2(p_fma(p, p_add(lg2, p_set(-lg2gf)), p_set(lg2gf))), gf, R(1));
        }
    }
    st.g_prev = G;
    st.gam_prev = gam;
    Sa = cscale(Ya, G.x);
    Sb = cscale(Yb, G.y);
}

struct EnhanceArgs {
    const CseTables* T;
    const real2* Y;      // [U][nf][nbp]
    const real* N;       // [U][nbp] or [U][nf][nbp]
    const cse_params* params;
    real* out;           // [(item - item0)][L]  (or [i][L] for item_list[i])
    const int* item_list; // optional: explicit items (u * n_params + c) of a sparse launch (winners' re-materialisation)
    int noise_tv, L, hop, n_frames, n_params, item0;
    int hop_shift;       // log2(hop) when hop is a power of two, else -1
    real eps;
};

#ifndef CSE_ENH_MB_SMALL
#define CSE_ENH_MB_SMALL 5    // resident CTAs per SM asked of ptxas for the <= 200-thread variants (n_fft <= 512)
#endif
#ifndef CSE_ENH_MB_LARGE
#define CSE_ENH_MB_LARGE 3    // same for the 288-thread variants (n_fft >= 1024): 72 registers; 2 (104 registers, no spills) measured slower
#endif
#ifndef CSE_ENH_CAP
#define CSE_ENH_CAP 256
#endif
template <int LOG2N> struct EnhanceCfg {
    static constexpr int NFFT = 1 << LOG2N, M = NFFT / 2;
    static constexpr int NPAIR = M / 2;                      // pair slots s: bins (s, M-s); slot 0 = (DC, Nyquist)
    static constexpr int NTB = NPAIR < CSE_ENH_CAP ? NPAIR : CSE_ENH_CAP;    // pair threads
    static constexpr int PPT = NPAIR / NTB;                  // pair slots per pair thread
    static constexpr int NT = NTB + 32;                      // + one warp whose lane 0 owns the self-paired bin M/2
    static constexpr int F = (8 * NTB) / M;                  // frames per iteration: F * M/8 butterflies == NTB
    static constexpr int XST = CSE_FFT_STRIDE(M);            // per-frame stride of the packed half-size spectra
    static constexpr int KMAX = ((NFFT + (F - 1) * (NFFT / 2)) / 2 + NT - 1) / NT;   // overlap-add pair-positions per thread (hop <= n_fft/2)
};

// Emission of a sample pair near the signal edges (first / last frames, odd lengths): the window
// sum-of-squares of the frames that really cover each sample (librosa window_sumsquare).  Rare; a
// __noinline__ version was measured 3 % slower (caller-saved spills on the hot path), so it stays inline.
template <int NFFT>
CSE_D void emit_edge_pair(real2 acc, int p, int i, int L, int nf, int hop, int hop_shift, real scale, bool vec2,
                                            const real* __restrict__ w, const real* wsteady, real* __restrict__ out) {
    real ws[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int pe = p + e;
        const int tmax = hop_shift >= 0 ? (pe >> hop_shift) : pe / hop, r = pe - tmax * hop;
        const int kmax = hop_shift >= 0 ? ((NFFT - 1 - r) >> hop_shift) : (NFFT - 1 - r) / hop;   // frames tmax-k, k <= kmax, cover pe
        if (tmax - kmax >= 0 && tmax < nf) ws[e] = wsteady[r];
        else {
            real s = R(0);
            for (int k = 0; k <= kmax; ++k) {
                const int t = tmax - k;
                if (t >= 0 && t < nf) s = r_fma(w[r + k * hop], w[r + k * hop], s);
            }
            ws[e] = s;
        }
    }
#ifdef CSE_FP64
    const real tiny = 2.2250738585072014e-308;
#else
    const real tiny = 1.17549435e-38f;
#endif
    const real o0 = ws[0] > tiny ? acc.x * scale / ws[0] : acc.x * scale;
    const real o1 = ws[1] > tiny ? acc.y * scale / ws[1] : acc.y * scale;
    if (vec2 && i >= 0 && i + 1 < L) *reinterpret_cast<real2*>(out + i) = mk2(o0, o1);
    else {
        if (i >= 0) out[i] = o0;
        if (i + 1 < L) out[i + 1] = o1;
    }
}

// One CTA per (utterance, grid point).  Each pair thread owns the bins (s, M-s) of its PPT pair
// slots: the packed half-size inverse-FFT input Z[s] = E + iO, Z[M-s] = conj(E) + i conj(O)
// (E = X[s] + conj X[M-s], O = (X[s] - conj X[M-s]) W_N^-s) needs exactly those two gained bins,
// so it is formed in registers and written straight into the FFT buffer - no separate split
// pass, no exchange through shared memory.
//
// STAGED (the product path for n_fft <= 1024): the F frames of Y (and of a time-varying noise PSD) an
// iteration needs arrive in shared memory by ONE bulk copy each (TMA engine, mbarrier completion), issued
// by thread 0 right after the gain phase has consumed the previous tile - the copy flies during the FFT and
// the overlap-add, no thread holds prefetched spectra in registers across those phases, and the gain phase
// reads its bins with LDS instead of LDG + 64-bit address arithmetic.  !STAGED is the round-1 register
// prefetch, kept for n_fft = 2048, whose tile would cost a resident CTA.
//
// The body is a device function of (arguments, block index) so that two entry points share it: enhance_kernel
// (one noise-PSD group per launch, the block index is the grid's) and enhance_groups_kernel (several groups of
// one instantiation per launch, each block finds its group first).
template <int ALG, int LOG2N, bool STAGED, bool TV, bool GAM>
__device__ __forceinline__ void enhance_body(const EnhanceArgs& a, const int bx) {
    typedef EnhanceCfg<LOG2N> C;
    constexpr int NFFT = C::NFFT, M = C::M, LOG2M = LOG2N - 1, NTB = C::NTB, PPT = C::PPT, NT = C::NT, F = C::F;
    constexpr int XST = C::XST;
    CSE_DYN_SMEM(smem_raw);
    // Shared memory: the arrays whose size is known at compile time come first, so that their addresses are the
    // base plus an immediate; the hop-dependent ones (ring, window sums, TMA tile) follow.
    constexpr int TWSZ = FftTwLayout<LOG2M, true>::SIZE;
    real2* xs = reinterpret_cast<real2*>(smem_raw);                    // F * XST (+ 1 spare)
    real2* w2s = xs + F * XST + 1;                                     // M window pairs (w[2m], w[2m+1])
    real2* tws = w2s + M;                                              // per-pass twiddles of the half-size FFT (FftTwLayout)
    real2* inv_ws_s = tws + TWSZ;                                      // M/2 (hop/2 used): 1/(N wss) of the emitted pairs
    real* pv_s = reinterpret_cast<real*>(inv_ws_s + M / 2);            // 8 params + derived constants (16 slots)
    real* ring = pv_s + 16;                                            // W
    constexpr unsigned OFF_W2S = (unsigned)((F * XST + 1) * sizeof(real2)), OFF_INV = OFF_W2S + (unsigned)((M + TWSZ) * sizeof(real2));
    constexpr unsigned OFF_RING = OFF_INV + (unsigned)((M / 2) * sizeof(real2) + 16 * sizeof(real));
    const int hop = a.hop;
    const int W = NFFT + (F - 1) * hop;
    real* wsteady = ring + W;                                          // hop: steady-state window sum-of-squares
    const int tid = threadIdx.x;
    const int item = a.item_list ? a.item_list[bx] : a.item0 + bx;
    const int u = item / a.n_params, c = item - u * a.n_params;
    const int nbp = cse_nbp(NFFT);
    const int nf = a.n_frames, L = a.L;
    const real* __restrict__ w = cse_hann(a.T, NFFT);

    if (tid < 8) {
        real v = (real)a.params[c].v[tid];
        if (ALG == 2 && tid == 4) v = r_clip(v, R(0), R(0.9999));           // mmse.py:51
        if (ALG == 3 && tid == 3) v = r_clip(v, R(0), R(0.9999));           // advanced_mmse.py:61
        if (ALG == 3 && tid == 4) v = r_clip(v, R(1e-3), R(1) - R(1e-3));   // advanced_mmse.py:72
        pv_s[tid] = v;
    }
    for (int i = tid; i < W; i += NT) ring[i] = R(0);
    // shared-memory tail (after the tables whose size depends on the hop): see enhance_smem_bytes()
    const int W_ = NFFT + (F - 1) * hop;
    // (a byte offset from the shared-memory base, not pointer <-> integer casts: the compiler must keep seeing a
    // shared-memory pointer, or the tile reads become generic loads with 64-bit address arithmetic)
    const unsigned tail_off = (OFF_RING + (unsigned)((W_ + hop) * sizeof(real)) + 15u) & ~15u;
    unsigned char* tail = smem_raw + tail_off;
    real2* ytile = reinterpret_cast<real2*>(tail);                                    // [F][nbp]  (STAGED)
    real* ntile = reinterpret_cast<real*>(ytile + (STAGED ? F * cse_nbp(NFFT) : 0));  // [F][nbp]  (STAGED, time-varying noise)
    cse_mbar_t* bar = reinterpret_cast<cse_mbar_t*>(ntile + ((STAGED && TV) ? F * cse_nbp(NFFT) : 0));
    if (STAGED && tid == 0) cse_mbar_init(bar, 1);
    for (int m = tid; m < M; m += NT) w2s[m] = mk2(w[2 * m], w[2 * m + 1]);
    load_pass_twiddles<LOG2M, true>(tws, a.T->tw, tid, NT);
    for (int r = tid; r < hop; r += NT) {
        real s = R(0);
        for (int n = r; n < NFFT; n += hop) s = r_fma(w[n], w[n], s);
        wsteady[r] = s;
    }
    __syncthreads();
    if (tid == 0) {
        if (ALG == 3) pv_s[6] = r_log(pv_s[2]) * CSE_LOG2E;             // log2(gain_floor)
        pv_s[8] = R(1) - pv_s[0];                                       // 1 - alpha
        pv_s[9] = R(1) - ((ALG == 2) ? pv_s[4] : pv_s[3]);              // 1 - noise_mu
        pv_s[10] = R(1) - pv_s[4];                                      // 1 - q (Log-MMSE)
    }
    __syncthreads();
    const real mu_raw = (ALG == 2) ? (real)a.params[c].v[4] : (ALG == 3) ? (real)a.params[c].v[3] : R(-1);
    const bool smooth = (ALG >= 2) && TV && !GAM && (mu_raw >= R(0));

    const real2* __restrict__ Yu = a.Y + (size_t)u * nf * nbp;
    const real* __restrict__ Nu = a.N + (size_t)u * (TV ? (size_t)nf * nbp : (size_t)nbp);
    real* __restrict__ out = a.out + (size_t)bx * L;

    const bool is_pair = tid < NTB, is_mid = tid == NTB;       // lane 0 of the extra warp: bin M/2
    const int n_slots = is_pair ? PPT : (is_mid ? 1 : 0);
    GainState2 st[PPT];
    real nstat[PPT][2];
    real2 rstat[PPT];
    real2 twc[PPT];                                           // -i W_N^s of each pair slot
    real2 yv[PPT][2][F];
    real nv[PPT][2][F];
    // bins (ka, kb) of slot i: pair thread -> (s, M - s) with s = tid + i*NTB (slot 0: (0, M)); mid lane -> (M/2, -)
    auto bin_a = [&](int i) { return is_pair ? tid + i * NTB : M / 2; };
    auto bin_b = [&](int i) { return M - (tid + i * NTB); };
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        st[i].g_prev = mk2(R(1), R(1)); st[i].gam_prev = mk2(R(1), R(1)); st[i].nsm = mk2(R(0), R(0));
        // static PSD: floored once here (the algorithms' np.maximum(noise_psd, eps)), reciprocal taken once
        nstat[i][0] = (!TV && i < n_slots) ? r_max(Nu[bin_a(i)], a.eps) : R(1);
        nstat[i][1] = (!TV && is_pair) ? r_max(Nu[bin_b(i)], a.eps) : ((!TV && STAGED && i < n_slots) ? r_max(Nu[M / 2], a.eps) : R(1));
        rstat[i] = mk2(r_rcp(nstat[i][0]), r_rcp(nstat[i][1]));
        const real2 tws_i = tw_load(a.T->tw, (is_pair ? tid + i * NTB : (STAGED ? M / 2 : 0)) * (CSE_TW_N / NFFT));
        twc[i] = mk2(tws_i.y, -tws_i.x);                      // -i W_N^s: conj-multiplying by it gives i D W_N^-s in one step
    }
    auto fetch = [&](int t0) {
        if (t0 + F <= nf) {
            // fast path (all F frames exist): running row pointer, compile-time frame offsets
            const int row0 = t0 * nbp;            // < 2^31: nf <= 8192 frames of <= 1032 bins
#pragma unroll
            for (int i = 0; i < PPT; ++i) {
                if (i < n_slots) {
                    const real2* __restrict__ pa = Yu + row0 + bin_a(i);
                    const real* __restrict__ na = Nu + (TV ? row0 : 0) + bin_a(i);
#pragma unroll
                    for (int f = 0; f < F; ++f) {
                        yv[i][0][f] = pa[f * nbp];
                        nv[i][0][f] = TV ? na[f * nbp] : nstat[i][0];
                    }
                    if (is_pair) {
                        const real2* __restrict__ pb = Yu + row0 + bin_b(i);
                        const real* __restrict__ nb_ = Nu + (TV ? row0 : 0) + bin_b(i);
#pragma unroll
                        for (int f = 0; f < F; ++f) {
                            yv[i][1][f] = pb[f * nbp];
                            nv[i][1][f] = TV ? nb_[f * nbp] : nstat[i][1];
                        }
                    }
                }
            }
            return;
        }
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            const int ka = bin_a(i), kb = bin_b(i);
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const int t = t0 + f;
                const int row = t * nbp;
                const bool on = i < n_slots && t < nf;
                yv[i][0][f] = on ? Yu[row + ka] : mk2(R(0), R(0));
                nv[i][0][f] = on ? (TV ? Nu[row + ka] : nstat[i][0]) : R(1);
                const bool onb = on && is_pair;
                yv[i][1][f] = onb ? Yu[row + kb] : mk2(R(0), R(0));
                nv[i][1][f] = onb ? (TV ? Nu[row + kb] : nstat[i][1]) : R(1);
            }
        }
    };
    // lanes without a second bin (the mid lane) keep a harmless dummy in their b half
#pragma unroll
    for (int i = 0; i < PPT; ++i)
#pragma unroll
        for (int f = 0; f < F; ++f) { yv[i][0][f] = mk2(R(0), R(0)); yv[i][1][f] = mk2(R(0), R(0)); nv[i][0][f] = R(1); nv[i][1][f] = R(1); }
    // STAGED: thread 0 arms the mbarrier with the tile's byte count and issues the bulk copies of the frames
    // [t0, min(t0 + F, nf)) - contiguous in the frame-major layout - into the shared-memory tile.
    auto issue_tile = [&](int t0) {
        const int rows = nf - t0 < F ? nf - t0 : F;
        const unsigned yb = (unsigned)(rows * nbp * sizeof(real2)), nbytes = TV ? (unsigned)(rows * nbp * sizeof(real)) : 0u;
        cse_mbar_expect_tx(bar, yb + nbytes);
        cse_bulk_g2s(ytile, Yu + (size_t)t0 * nbp, yb, bar);
        if (TV) cse_bulk_g2s(ntile, Nu + (size_t)t0 * nbp, nbytes, bar);
    };
    unsigned tile_parity = 0;
    if (STAGED) { if (tid == 0) issue_tile(0); }        // (the mbarrier was initialised before the barriers above)
    else fetch(0);

    // Overlap-add.  The last inverse-FFT pass leaves every frame in NATURAL order (cse_fft.cuh, NAT), so the
    // sample pair m of frame f sits at xs[f * XST + m] and its window pair at w2s[m]: a thread that owns the
    // window pair-positions jj = tid + k*NT adds, for every frame f of the iteration with 0 <= m = jj - f*hop/2 < M,
    // x_f[m] * w[m] - two shared loads at (per-thread base) + (per-frame uniform offset) and one packed FMA per
    // term, no index table.  Which (k, f) terms exist is loop-invariant: one bit each in `fmask`.
    constexpr int KMAX = C::KMAX;
    const int hh = hop >> 1;
    static_assert(KMAX * F <= 32, "term mask");
    unsigned fmask = 0;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
#pragma unroll
        for (int f = 0; f < F; ++f) {
            const int jj = tid + k * NT, m = jj - f * hh;
            if (jj < W / 2 && m >= 0 && m < M) fmask |= 1u << (k * F + f);
        }

    // Steady-state normalisation of the pairs this thread emits (window positions j < F*hop): the
    // window sum-of-squares only depends on j mod hop there, so 1/(N * wss) is loop-invariant.
    for (int r2 = tid; r2 < (hop >> 1); r2 += NT)
        inv_ws_s[r2] = mk2(R(1) / ((real)NFFT * wsteady[2 * r2]), R(1) / ((real)NFFT * wsteady[2 * r2 + 1]));
    // Steady-state overlap-add (interior iterations, power-of-two hop): everything below is loop-invariant per
    // thread - how many window pair-positions it owns, which of them finish (are emitted) in an iteration - and
    // the loop runs on 32-bit shared-window addresses held in registers.
    // Four values carry the whole steady-state loop; they are made opaque to the compiler (empty asm), which at
    // this register budget otherwise re-derives them - thread index, masks, window base - inside the loop:
    //   tb    shared-window address of xs[tid]: every x / window / ring operand is tb + immediate (+ uniform)
    //   pm    term mask (KMAX*F bits) | emit mask (KMAX bits) << 20 | number of owned positions << 26
    //   raddr address of this thread's ring slot for k = 0, advanced by F*hop samples per iteration
    //   rend  end of the ring
    int kcnt = 0;
    unsigned emask = 0;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        if (tid + k * NT < W / 2) ++kcnt;
        if (2 * (tid + k * NT) < F * hop) emask |= 1u << k;
    }
    static_assert(KMAX * F <= 20 && KMAX <= 6, "packed loop masks");
    unsigned pm = fmask | (emask << 20) | ((unsigned)kcnt << 26);
    unsigned tb = cse_saddr(smem_raw) + (unsigned)(tid * sizeof(real2));
    unsigned raddr = tb + OFF_RING, rend = cse_saddr(ring) + (unsigned)(W * sizeof(real));
    CSE_OPAQUE(pm); CSE_OPAQUE(tb); CSE_OPAQUE(rend);
    const bool fast_ok = a.hop_shift >= 0 && ((L & 1) == 0);
    __syncthreads();                         // inv_ws_s is read by other threads than its writers

    const int total_pos = L + M;             // padded positions [0, L + M) must be emitted
    int ring_base = 0;                       // ring slot of padded position t0 * hop
    // iterations [t_lo, t_hi) are steady: every frame covering the emitted positions exists (t0*hop >= n_fft,
    // t0 + F <= n_frames) and the positions map inside [0, L) ((t0 + F)*hop <= L + M)
    int t_lo = (NFFT + hop - 1) / hop, t_hi = (nf - F < (L + M) / hop - F ? nf - F : (L + M) / hop - F) + 1;
    if (!fast_ok) t_hi = t_lo;
    real* outp = out - M + 2 * tid;          // + t0 * hop: where this thread's k = 0 pair of the iteration goes
    CSE_OPAQUE(t_lo); CSE_OPAQUE(t_hi);
    const real scale = R(1) / (real)NFFT;    // irfft normalisation (1/2 of the split, 1/M of the FFT)
    const bool vec2 = (L & 1) == 0;          // waveform rows 8-byte aligned -> paired stores

    for (int t0 = 0; t0 * hop < total_pos; t0 += F) {
        const bool any = t0 < nf;
        if (any) {
            if (STAGED) { cse_mbar_wait(bar, tile_parity); tile_parity ^= 1u; }
            // the candidate's parameters, read once per iteration: left in shared memory they would be re-read for
            // every frame (the compiler cannot prove that the spectrum stores of the frame loop do not alias them)
            real pvr[12];
#pragma unroll
            for (int q = 0; q < 12; ++q) pvr[q] = pv_s[q];
            const real* pv = pvr;
#pragma unroll
            for (int i = 0; i < PPT; ++i) {
                if (i < n_slots) {
                    const int s = tid + i * NTB;
#pragma unroll
                    for (int f = 0; f < F; ++f) {
                        const int t = t0 + f;
                        real2* xf = xs + f * XST;
                        if (t >= nf) {
                            if (is_pair) { xf[SIDX(s)] = mk2(R(0), R(0)); if (s > 0) xf[SIDX(M - s)] = mk2(R(0), R(0)); }
                            else xf[SIDX(M / 2)] = mk2(R(0), R(0));
                            continue;
                        }
                        real2 xa, xb;
                        if (STAGED) {
                            // every thread runs the generic pair formula: the mid lane with s = M/2 (both "bins" are its
                            // own bin, the two stores hit the same cell with the same value), slot 0 with (DC, Nyquist)
                            // whose imaginary parts an inverse real FFT ignores
                            const int sa = is_pair ? s : M / 2, ka = sa, kb = M - sa;
                            const real2 Ya = ytile[f * nbp + ka], Yb = ytile[f * nbp + kb];
                            const real2 Nn = TV ? mk2(ntile[f * nbp + ka], ntile[f * nbp + kb]) : mk2(nstat[i][0], nstat[i][1]);
                            gain_pair<ALG, TV, GAM>(Ya, Yb, Nn, rstat[i], t == 0, st[i], pv, a.eps, smooth, xa, xb);
                            if (sa == 0) { xa.y = R(0); xb.y = R(0); }
                            const real2 cb = mk2(xb.x, -xb.y);                     // conj X[M-s]
                            const real2 E = cadd(xa, cb), D = csub(xa, cb);
                            const real2 iO = cmulc(D, twc[i]);                     // i D W_N^-s (twc holds -i W_N^s)
                            const real2 T = csub(E, iO);
                            xf[SIDX(sa)] = cadd(E, iO);                            // E + iO
                            if (sa > 0) xf[SIDX(M - sa)] = mk2(T.x, -T.y);         // conj(E) + i conj(O) = conj(E - iO)
                            continue;
                        }
                        gain_pair<ALG, TV, GAM>(yv[i][0][f], yv[i][1][f], mk2(nv[i][0][f], nv[i][1][f]), rstat[i], t == 0, st[i], pv, a.eps, smooth, xa, xb);
                        if (!is_pair) { xf[SIDX(M / 2)] = mk2(R(2) * xa.x, R(-2) * xa.y); continue; }   // 2 conj X[M/2]
                        if (s == 0) { xf[0] = mk2(xa.x + xb.x, xa.x - xb.x); continue; }                // DC, Nyquist (real)
                        const real2 cb = mk2(xb.x, -xb.y);                     // conj X[M-s]
                        const real2 E = cadd(xa, cb), D = csub(xa, cb);
                        const real2 iO = cmulc(D, twc[i]);                     // i D W_N^-s (twc holds -i W_N^s)
                        const real2 T = csub(E, iO);
                        xf[SIDX(s)] = cadd(E, iO);                             // E + iO
                        xf[SIDX(M - s)] = mk2(T.x, -T.y);                      // conj(E) + i conj(O) = conj(E - iO)
                    }
                }
            }
            __syncthreads();
            // the tile has been consumed by every thread: refill it for the next iteration while the FFT and
            // the overlap-add run (reads through the generic proxy are ordered before the copy by the barrier)
            if (STAGED && tid == 0 && t0 + F < nf) issue_tile(t0 + F);
            static_assert(F * (M / 8) <= NT, "one butterfly per thread and pass");
            fft_dif<LOG2M, true, 0, false, true, true>(xs, F, XST, tws, tid, NT);   // group barriers, natural-order output
            if (!STAGED) fetch(t0 + F);      // next iteration's spectra fly during the overlap-add (issued after
                                             // the FFT so that they are not live across its register-hungry passes)
        }
        // overlap-add the F windowed frames two samples at a time (sample pair 2m,2m+1 of a frame
        // is one complex FFT output), emit the F*hop positions no later frame touches
        const int p_begin = t0 * hop;
        const int emit_end = p_begin + F * hop;
        // steady state: all frames covering the emitted positions exist and the positions map inside [0, L)
        if (t0 >= t_lo && t0 < t_hi) {
            const unsigned Wb = (unsigned)(W * sizeof(real)), hh8 = (unsigned)(hh * sizeof(real2));
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                if (k < (int)(pm >> 26)) {
                    unsigned ra = raddr + (unsigned)(k * NT * 2 * sizeof(real));
                    if (ra >= rend) ra -= Wb;
                    real2 acc = cse_lds_r2(ra);
#pragma unroll
                    for (int f = 0; f < F; ++f) {
                        if ((pm >> (k * F + f)) & 1u) {
                            const unsigned at = tb + (unsigned)((k * NT + f * XST) * sizeof(real2)) - (unsigned)f * hh8;
                            acc = cfma2(cse_lds_r2(at), cse_lds_r2(at + OFF_W2S - (unsigned)(f * XST * sizeof(real2))), acc);
                        }
                    }
                    const bool emit = (pm >> (20 + k)) & 1u;
                    cse_sts_r2(ra, emit ? mk2(R(0), R(0)) : acc);
                    if (emit) {
                        const real2 iw = inv_ws_s[(tid + k * NT) & (hh - 1)];
                        *reinterpret_cast<real2*>(outp + 2 * k * NT) = mk2(acc.x * iw.x, acc.y * iw.y);
                    }
                }
            }
        } else {
        const bool steady = p_begin >= NFFT && t0 + F <= nf && p_begin >= M && emit_end <= L + M;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            const int jj = tid + k * NT;
            if (jj >= W / 2) break;
            const int j = 2 * jj, p = p_begin + j;
            int slot = ring_base + j;
            if (slot >= W) slot -= W;
            real2 acc = *reinterpret_cast<real2*>(ring + slot);
            if (any) {                       // frames beyond n_frames were written as zero spectra
#pragma unroll
                for (int f = 0; f < F; ++f) {
                    const int m = jj - f * hh;
                    if (m >= 0 && m < M) acc = cfma2(xs[f * XST + m], w2s[m], acc);
                }
            }
            // the ring slot is settled first (same address register as the load above), then finished pairs go out
            const bool emit = p < emit_end;
            *reinterpret_cast<real2*>(ring + slot) = emit ? mk2(R(0), R(0)) : acc;
            if (emit) {
                const int i = p - M;
                if (steady && vec2) {
                    // interior of the signal: every covering frame exists, all samples are in range
                    const real2 iw = inv_ws_s[(j % hop) >> 1];
                    *reinterpret_cast<real2*>(out + i) = mk2(acc.x * iw.x, acc.y * iw.y);
                } else if (i >= -1 && i < L) {
                    emit_edge_pair<NFFT>(acc, p, i, L, nf, hop, a.hop_shift, scale, vec2, w, wsteady, out);
                }
            }
        }
        }
        outp += F * hop;
        CSE_OPAQUE_PTR(outp);
        ring_base += F * hop;
        while (ring_base >= W) ring_base -= W;
        raddr += (unsigned)(F * hop * sizeof(real));
        while (raddr >= rend) raddr -= (unsigned)(W * sizeof(real));
        __syncthreads();
    }
}



template <int ALG, int LOG2N, bool STAGED, bool TV, bool GAM = false>
__global__ void __launch_bounds__(EnhanceCfg<LOG2N>::NT, (EnhanceCfg<LOG2N>::NT > 200 ? CSE_ENH_MB_LARGE : CSE_ENH_MB_SMALL)) enhance_kernel(EnhanceArgs a) {
    enhance_body<ALG, LOG2N, STAGED, TV, GAM>(a, (int)blockIdx.x);
}

// Several noise-PSD groups of ONE instantiation (algorithm, n_fft, kind of noise input) in one launch: the groups of a
// small batch - one pair is the reference's own call pattern - are a dozen to a few hundred candidates each, far below
// a wave; launched together they fill the GPU.  The descriptors travel as kernel parameters (no device copy, no
// lifetime to manage); a block finds its group by a scan over at most CSE_ENH_MAX_GROUPS first-block indices (uniform).
#define CSE_ENH_MAX_GROUPS 24
struct EnhanceGroup {
    const real2* Y;
    const real* N;
    const cse_params* params;
    real* out;                       // [n_utts * n_params][L] of this group
    int hop, n_frames, hop_shift, n_params, first_block, reserved;
};
struct EnhanceGroupsArgs {
    const CseTables* T;
    int noise_tv, L, n_groups;
    real eps;
    EnhanceGroup g[CSE_ENH_MAX_GROUPS];
};
template <int ALG, int LOG2N, bool STAGED, bool TV, bool GAM = false>
__global__ void __launch_bounds__(EnhanceCfg<LOG2N>::NT, (EnhanceCfg<LOG2N>::NT > 200 ? CSE_ENH_MB_LARGE : CSE_ENH_MB_SMALL)) enhance_groups_kernel(EnhanceGroupsArgs ga) {
    int gi = 0;
    for (int k = 1; k < ga.n_groups; ++k)
        if ((int)blockIdx.x >= ga.g[k].first_block) gi = k;
    const EnhanceGroup& g = ga.g[gi];
    EnhanceArgs a;
    a.T = ga.T; a.Y = g.Y; a.N = g.N; a.params = g.params; a.out = g.out; a.item_list = nullptr;
    a.noise_tv = ga.noise_tv; a.L = ga.L; a.hop = g.hop; a.n_frames = g.n_frames; a.n_params = g.n_params; a.item0 = 0;
    a.hop_shift = g.hop_shift; a.eps = ga.eps;
    enhance_body<ALG, LOG2N, STAGED, TV, GAM>(a, (int)blockIdx.x - g.first_block);
}


// The candidate-invariant front of the gain rules, once per (utterance, STFT shape, noise PSD, noise_mu):
//   N'(t) = max(N(t), eps)                          (np.maximum(noise_psd, eps): wiener_filter.py:45, mmse.py:45, advanced_mmse.py:57)
//   N'(t) = mu N'(t-1) + (1 - mu) N'(t),  t > 0       (recursive smoothing: mmse.py:48-57, advanced_mmse.py:60-66; mu < 0: none)
//   gamma(t) = max(|Y(t)|^2 / N'(t), eps)            (a-posteriori SNR: wiener_filter.py:61, mmse.py:67, advanced_mmse.py:76)
// One thread per (utterance, bin) walks the frames; a warp reads 32 neighbouring bins of a frame (coalesced).
// Same operations, in the same order, as gain_pair<..., GAM = false>.
CSE_D void gamma_body(const real2* __restrict__ Y, const real* __restrict__ N, int noise_tv, int nf, int nb, int nbp, real mu,
                      real eps, real* __restrict__ out, const int k, const int u) {
    if (k >= nbp) return;
    const real2* __restrict__ Yu = Y + (size_t)u * nf * nbp + k;
    const real* __restrict__ Nu = N + (size_t)u * (noise_tv ? (size_t)nf * nbp : (size_t)nbp) + k;
    real* __restrict__ o = out + (size_t)u * nf * nbp + k;
    if (k >= nb) { for (int t = 0; t < nf; ++t) o[(size_t)t * nbp] = eps; return; }      // padding bins
    const bool smooth = noise_tv && mu >= R(0);
    const real one_minus_mu = R(1) - mu;
    const real nstat = noise_tv ? R(0) : r_max(Nu[0], eps);
    const real rstat = noise_tv ? R(0) : r_rcp(nstat);
    real nsm = R(0);
    // eight frames' loads are issued before their (sequential) use: the chain is bound by memory latency, and after
    // inlining into the grouped entry point the compiler no longer batches the loads by itself
    constexpr int UF = 8;
    auto step = [&](int t, real2 y, real nraw) {
        const real pw = r_fma(y.x, y.x, y.y * y.y);
        real g;
        if (noise_tv) {
            real nt = r_max(nraw, eps);
            if (smooth) {
                if (t > 0) nt = r_fma(mu, nsm, one_minus_mu * nt);
                nsm = nt;
                if (!CSE_GAIN_TRIM) nt = r_max(nt, eps);
            }
            g = pw * r_rcp(nt);
        } else g = pw * rstat;
        o[(size_t)t * nbp] = r_max(g, eps);
    };
    int t = 0;
    for (; t + UF <= nf; t += UF) {
        real2 y[UF];
        real nn[UF];
#pragma unroll
        for (int j = 0; j < UF; ++j) {
            y[j] = __ldg(Yu + (size_t)(t + j) * nbp);
            nn[j] = noise_tv ? __ldg(Nu + (size_t)(t + j) * nbp) : R(0);
        }
#pragma unroll
        for (int j = 0; j < UF; ++j) step(t + j, y[j], nn[j]);
    }
    for (; t < nf; ++t) step(t, __ldg(Yu + (size_t)t * nbp), noise_tv ? __ldg(Nu + (size_t)t * nbp) : R(0));
}

__global__ void __launch_bounds__(128) gamma_kernel(const real2* __restrict__ Y, const real* __restrict__ N, int noise_tv, int nf,
                                                    int nb, int nbp, real mu, real eps, real* __restrict__ out) {
    gamma_body(Y, N, noise_tv, nf, nb, nbp, mu, eps, out, (int)(blockIdx.x * blockDim.x + threadIdx.x), (int)blockIdx.y);
}

// Several (noise PSD, noise_mu) groups of one n_fft in one launch (blockIdx.z = group): a single pair needs ~36 of them,
// each a chain of n_frames dependent steps on 513 threads - one after the other they cost more than the gain kernels.
#define CSE_GAMMA_MAX_GROUPS 32
struct GammaGroup {
    const real2* Y;
    const real* N;
    real* out;
    int noise_tv, nf;
    real mu, eps;
};
struct GammaGroupsArgs {
    int nb, nbp;
    GammaGroup g[CSE_GAMMA_MAX_GROUPS];
};
__global__ void __launch_bounds__(128) gamma_groups_kernel(GammaGroupsArgs ga) {
    const GammaGroup& g = ga.g[blockIdx.z];
    gamma_body(g.Y, g.N, g.noise_tv, g.nf, ga.nb, ga.nbp, g.mu, g.eps, g.out, (int)(blockIdx.x * blockDim.x + threadIdx.x), (int)blockIdx.y);
}
