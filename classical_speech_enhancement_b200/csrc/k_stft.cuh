// K1: fused reflect-pad + Hann window + real FFT + |X|^2.
//
// Restates librosa.stft(y, n_fft, hop, win_length=n_fft, window="hann", center=True,
// pad_mode="reflect") followed by np.abs(.)**2 as the reference calls them
// (Code/spectral_subtractor.py:25-26, wiener_filter.py:35-37, mmse.py:29-32,
// advanced_mmse.py:39-40, noise_estimation.py:184-188; oracle noise :128-147).
//
// One CTA transforms F consecutive frames of one utterance: the windowed frames are packed
// two real samples per complex value, transformed by F batched half-size DIF FFTs in shared
// memory, and unpacked to the n_fft/2+1 bins on the way out (split step reads the
// bit-reversed FFT output in place).  Output rows are written bin-fastest, coalesced.
#pragma once
#include "cse_fft.cuh"

template <int LOG2N /* log2(n_fft) */, int F>
__global__ void __launch_bounds__(256) stft_psd_kernel(const CseTables* __restrict__ T,
                                                       const real* __restrict__ wav,
                                                       const real* __restrict__ minus, int L, int hop,
                                                       int n_frames, real psd_floor, real2* __restrict__ Y,
                                                       real* __restrict__ P) {
    constexpr int NFFT = 1 << LOG2N, M = NFFT / 2, LOG2M = LOG2N - 1;
    constexpr int BST = CSE_FFT_STRIDE(M);
    CSE_DYN_SMEM(smem_raw);
    real2* z = reinterpret_cast<real2*>(smem_raw);
    const int tid = threadIdx.x, nth = blockDim.x;
    const int u = blockIdx.y, t0 = blockIdx.x * F;
    const int nbp = cse_nbp(NFFT);
    const real* __restrict__ w = cse_hann(T, NFFT);
    const real* x = wav + (size_t)u * L;
    const real* xm = minus ? minus + (size_t)u * L : nullptr;

    for (int idx = tid; idx < F * M; idx += nth) {
        const int f = idx / M, m = idx - f * M;
        const int t = t0 + f;
        real2 v = mk2(R(0), R(0));
        if (t < n_frames) {
            int p0 = t * hop + 2 * m - M, p1 = p0 + 1;           // position in the unpadded signal
            if (p0 < 0) p0 = -p0;
            if (p0 >= L) p0 = 2 * (L - 1) - p0;
            if (p1 < 0) p1 = -p1;
            if (p1 >= L) p1 = 2 * (L - 1) - p1;
            real a = x[p0], b = x[p1];
            if (xm) { a -= xm[p0]; b -= xm[p1]; }
            v = mk2(a * w[2 * m], b * w[2 * m + 1]);
        }
        z[f * BST + SIDX(m)] = v;
    }
    __syncthreads();
    fft_dif<LOG2M, false>(z, F, BST, T->tw, tid, nth);

    // split: X[k] = E + W_N^k O,  E = (Z[k] + conj Z[M-k])/2,  O = -i (Z[k] - conj Z[M-k])/2
    for (int idx = tid; idx < F * nbp; idx += nth) {
        const int f = idx / nbp, k = idx - f * nbp;
        const int t = t0 + f;
        if (t >= n_frames) continue;
        real2 X = mk2(R(0), R(0));
        real pw = R(0);
        if (k <= M) {
            const real2* zf = z + f * BST;
            const real2 z0 = zf[SIDX(brev_n(k == M ? 0 : k, LOG2M))];
            if (k == 0) X = mk2(z0.x + z0.y, R(0));
            else if (k == M) X = mk2(z0.x - z0.y, R(0));
            else {
                const real2 z1 = zf[SIDX(brev_n(M - k, LOG2M))];
                const real2 E = mk2(R(0.5) * (z0.x + z1.x), R(0.5) * (z0.y - z1.y));
                const real2 O = mk2(R(0.5) * (z0.y + z1.y), R(-0.5) * (z0.x - z1.x));
                X = cadd(E, cmul(O, tw_load(T->tw, k * (CSE_TW_N / NFFT))));
            }
            pw = X.x * X.x + X.y * X.y;
            if (xm) pw = r_max(pw, psd_floor);
        }
        const size_t o = ((size_t)u * n_frames + t) * nbp + k;
        if (Y) Y[o] = X;
        if (P) P[o] = pw;
    }
}
