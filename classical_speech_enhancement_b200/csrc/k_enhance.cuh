// K3+K4: per-candidate gain (frame march) fused with the inverse real FFT and the weighted
// overlap-add of librosa.istft.  One CTA per (utterance, grid point); one thread per bin.
//
// Restates, after their STFT / noise_estimation calls:
//   ALG 0  spectral_subtraction  Code/spectral_subtractor.py:37-62
//   ALG 1  wiener_filter         Code/wiener_filter.py:47-94
//   ALG 2  mmse                  Code/mmse.py:48-118
//   ALG 3  advanced_mmse         Code/advanced_mmse.py:51-135
// and librosa.istft(S, hop_length, win_length=n_fft, window="hann", center=True, length=L).
//
// The decision-directed recursion is sequential over frames but independent per bin, so each
// thread carries (previous gain, previous a-posteriori SNR, smoothed noise PSD) for its bin in
// registers and marches F frames per iteration; the F gained spectra are turned into F packed
// half-size inverse FFTs in shared memory (all threads busy), windowed, overlap-added into a
// ring buffer, normalised by the running window sum-of-squares and streamed out, F*hop
// finished samples per iteration.  Y / N reads and waveform writes are coalesced; the
// spectra are prefetched into registers one iteration ahead.
#pragma once
#include "cse_fft.cuh"
#include "cse_special.cuh"

struct GainState { real g_prev, gam_prev, nsm; };

template <int ALG>
CSE_D real2 gain_apply(real2 Yv, real Nraw, bool first, GainState& st, const real* __restrict__ pv,
                       real eps, bool smooth) {
    const real Pw = Yv.x * Yv.x + Yv.y * Yv.y;
    real Nt = r_max(Nraw, eps);
    if (ALG == 0) {
        // Pc = max(P - alpha N, beta N); S = sqrt(Pc) * exp(j angle(Y))
        const real Pc = r_max(Pw - pv[0] * Nt, pv[1] * Nt);
        if (Pw > R(0)) { const real g = r_sqrt(Pc / Pw); return mk2(Yv.x * g, Yv.y * g); }
        return mk2(r_sqrt(Pc), R(0));
    }
    if (ALG >= 2 && smooth) {           // recursive smoothing of a time-varying noise PSD
        const real mu = (ALG == 2) ? pv[4] : pv[3];
        Nt = first ? Nt : r_fma(mu, st.nsm, (R(1) - mu) * Nt);
        st.nsm = Nt;
        Nt = r_max(Nt, eps);
    }
    const real gam = r_max(Pw / Nt, eps);
    const real direct = r_max(gam - R(1), R(0));
    const real alpha = pv[0];
    real G;
    if (ALG == 1) {
        real xi = first ? direct : r_fma(alpha, st.g_prev * st.g_prev * st.gam_prev, (R(1) - alpha) * direct);
        xi = r_max(xi, R(1e-10));
        G = r_clip(xi / (R(1) + xi), pv[1], R(1));
    } else {
        const real ksi_min = pv[1];
        real xi = first ? (gam - R(1)) : r_fma(alpha, st.g_prev * st.g_prev * st.gam_prev, (R(1) - alpha) * direct);
        xi = r_max(xi, ksi_min);
        if (ALG == 2) {
            const real v = r_clip(xi * gam / (R(1) + xi), eps, R(80));
            const real A = R(0.88622692545275801365) * r_sqrt(v) / (gam + eps);
            G = r_clip(A * cse_mmse_bessel_term(v), pv[2], pv[3]);
        } else {
            const real gf = pv[2], q = pv[4], vmax = pv[5], lngf = pv[6];
            const real v = r_clip(xi * gam / (R(1) + xi), R(1e-12), vmax);
            const real opx = R(1) + xi;
            const real lg = r_log(xi / opx) + R(0.5) * cse_expint_e1(v);          // ln(G_lsa)
            const real lam = r_exp(v) / opx;
            const real p = r_clip(R(1) / (R(1) + (R(1) - q) / r_fma(q, lam, eps)), R(0), R(1));
            G = r_clip(r_exp(r_fma(p, lg - lngf, lngf)), gf, R(1));               // G_lsa^p gf^(1-p)
        }
    }
    st.g_prev = G;
    st.gam_prev = gam;
    return mk2(Yv.x * G, Yv.y * G);
}

struct EnhanceArgs {
    const CseTables* T;
    const real2* Y;      // [U][nf][nbp]
    const real* N;       // [U][nbp] or [U][nf][nbp]
    const cse_params* params;
    real* out;           // [(item - item0)][L]
    int noise_tv, L, hop, n_frames, n_params, item0;
    real eps;
};

template <int ALG, int LOG2N, int F>
__global__ void __launch_bounds__(1 << (LOG2N - 1)) enhance_kernel(EnhanceArgs a) {
    constexpr int NFFT = 1 << LOG2N, M = NFFT / 2, LOG2M = LOG2N - 1;
    constexpr int XST = CSE_FFT_STRIDE(M) + 2;        // per-frame stride: M padded values + Nyquist slot
    CSE_DYN_SMEM(smem_raw);
    real2* xs = reinterpret_cast<real2*>(smem_raw);                    // F * XST
    real* ring = reinterpret_cast<real*>(xs + F * XST);                // W
    const int hop = a.hop;
    const int W = NFFT + (F - 1) * hop;
    real* ring2 = ring + W;                                            // window sum-of-squares ring
    real* pv_s = ring2 + W;                                            // 8 params
    const int tid = threadIdx.x;                                       // == bin index, blockDim.x == M
    const int item = a.item0 + blockIdx.x;
    const int u = item / a.n_params, c = item - u * a.n_params;
    const int nbp = cse_nbp(NFFT);
    const int nf = a.n_frames, L = a.L;
    const real* __restrict__ w = cse_hann(a.T, NFFT);

    if (tid < 8) {
        real v = (real)a.params[c].v[tid];
        if (ALG == 2 && tid == 4) v = r_clip(v, R(0), R(0.9999));       // mmse.py:51
        if (ALG == 3 && tid == 3) v = r_clip(v, R(0), R(0.9999));       // advanced_mmse.py:61
        if (ALG == 3 && tid == 4) v = r_clip(v, R(1e-3), R(1) - R(1e-3));   // advanced_mmse.py:72
        pv_s[tid] = v;
    }
    for (int i = tid; i < 2 * W; i += M) ring[i] = R(0);
    __syncthreads();
    if (ALG == 3 && tid == 0) pv_s[6] = r_log(pv_s[2]);
    __syncthreads();
    real pv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pv[i] = pv_s[i];
    const real mu_raw = (ALG == 2) ? (real)a.params[c].v[4] : (ALG == 3) ? (real)a.params[c].v[3] : R(-1);
    const bool smooth = (ALG >= 2) && a.noise_tv && (mu_raw >= R(0));

    const real2* __restrict__ Yu = a.Y + (size_t)u * nf * nbp;
    const real* __restrict__ Nu = a.N + (size_t)u * (a.noise_tv ? (size_t)nf * nbp : (size_t)nbp);
    real* __restrict__ out = a.out + (size_t)blockIdx.x * L;

    GainState st{R(1), R(1), R(0)}, stN{R(1), R(1), R(0)};     // stN: Nyquist bin, thread 0 only
    const real n_static = a.noise_tv ? R(0) : Nu[tid];
    const real n_static_ny = (!a.noise_tv && tid == 0) ? Nu[M] : R(0);

    real2 ycur[F], ynxt[F];
    real ncur[F], nnxt[F];
    auto prefetch = [&](int t0, real2* yy, real* nn) {
#pragma unroll
        for (int f = 0; f < F; ++f) {
            const int t = t0 + f;
            if (t < nf) {
                yy[f] = Yu[(size_t)t * nbp + tid];
                nn[f] = a.noise_tv ? Nu[(size_t)t * nbp + tid] : n_static;
            } else { yy[f] = mk2(R(0), R(0)); nn[f] = R(1); }
        }
    };
    prefetch(0, ycur, ncur);

    const int total_pos = L + M;             // padded positions [0, L + M) must be emitted
    int ring_base = 0;                       // ring slot of padded position t0 * hop
    const real scale = R(1) / (real)NFFT;    // irfft normalisation (1/2 of the split and 1/M of the FFT)

    for (int t0 = 0; t0 * hop < total_pos; t0 += F) {
        const bool any = t0 < nf;
        if (any) {
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const int t = t0 + f;
                if (t < nf) {
                    xs[f * XST + SIDX(tid)] = gain_apply<ALG>(ycur[f], ncur[f], t == 0, st, pv, a.eps, smooth);
                    if (tid == 0) {
                        const real2 yn = Yu[(size_t)t * nbp + M];
                        const real nn = a.noise_tv ? Nu[(size_t)t * nbp + M] : n_static_ny;
                        xs[f * XST + CSE_FFT_STRIDE(M)] = gain_apply<ALG>(yn, nn, t == 0, stN, pv, a.eps, smooth);
                    }
                } else {
                    xs[f * XST + SIDX(tid)] = mk2(R(0), R(0));
                    if (tid == 0) xs[f * XST + CSE_FFT_STRIDE(M)] = mk2(R(0), R(0));
                }
            }
            prefetch(t0 + F, ynxt, nnxt);
            __syncthreads();
            // pre-split (in place): Z[k] = E + iO, Z[M-k] = conj(E) + i conj(O);
            // E = X[k] + conj X[M-k], O = (X[k] - conj X[M-k]) W_N^-k   (scale 1/2 folded into `scale`)
            for (int idx = tid; idx < F * (M / 2 + 1); idx += M) {
                const int f = idx / (M / 2 + 1), k = idx - f * (M / 2 + 1);
                real2* xf = xs + f * XST;
                if (k == 0) {
                    const real x0 = xf[0].x, xm = xf[CSE_FFT_STRIDE(M)].x;
                    xf[0] = mk2(x0 + xm, x0 - xm);
                } else if (k == M / 2) {
                    const real2 x = xf[SIDX(k)];
                    xf[SIDX(k)] = mk2(R(2) * x.x, R(-2) * x.y);
                } else {
                    const real2 xa = xf[SIDX(k)], xb = xf[SIDX(M - k)];
                    const real2 E = mk2(xa.x + xb.x, xa.y - xb.y);
                    const real2 D = mk2(xa.x - xb.x, xa.y + xb.y);
                    const real2 O = cmulc(D, tw_load(a.T->tw, k * (CSE_TW_N / NFFT)));   // D * W_N^-k
                    xf[SIDX(k)] = mk2(E.x - O.y, E.y + O.x);            // E + iO
                    xf[SIDX(M - k)] = mk2(E.x + O.y, O.x - E.y);        // conj(E) + i conj(O)
                }
            }
            __syncthreads();
            fft_dif<LOG2M, true>(xs, F, XST, a.T->tw, tid, M);
        }
        // overlap-add the F windowed frames, emit the F*hop positions no later frame touches
        const int p_begin = t0 * hop;
        const int emit_end = p_begin + F * hop;
        for (int j = tid; j < W; j += M) {
            const int p = p_begin + j;
            int slot = ring_base + j;
            if (slot >= W) slot -= W;
            real acc = ring[slot], acc2 = ring2[slot];
            if (any) {
#pragma unroll
                for (int f = 0; f < F; ++f) {
                    const int n = j - f * hop;               // sample index inside frame t0+f
                    if (n >= 0 && n < NFFT && t0 + f < nf) {
                        const real2 zz = xs[f * XST + SIDX(brev_n(n >> 1, LOG2M))];
                        const real wv = w[n];
                        acc = r_fma((n & 1) ? zz.y : zz.x, wv * scale, acc);
                        acc2 = r_fma(wv, wv, acc2);
                    }
                }
            }
            if (p < emit_end) {
                const int i = p - M;
                if (i >= 0 && i < L) {
#ifdef CSE_FP64
                    const real tiny = 2.2250738585072014e-308;
#else
                    const real tiny = 1.17549435e-38f;
#endif
                    out[i] = acc2 > tiny ? acc / acc2 : acc;
                }
                ring[slot] = R(0);
                ring2[slot] = R(0);
            } else {
                ring[slot] = acc;
                ring2[slot] = acc2;
            }
        }
        ring_base += F * hop;
        while (ring_base >= W) ring_base -= W;
#pragma unroll
        for (int f = 0; f < F; ++f) { ycur[f] = ynxt[f]; ncur[f] = nnxt[f]; }
        __syncthreads();
    }
}
