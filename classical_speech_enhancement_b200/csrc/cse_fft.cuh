// Block-cooperative in-place FFTs in shared memory.
//
// Two mirrored radix-2 flow graphs, executed three stages at a time in registers (radix-8
// passes, one __syncthreads per pass):
//   fft_dif<LOG2N, INV>  natural-order input  -> bit-reversed output
//   fft_dit<LOG2N, INV>  bit-reversed input   -> natural-order output
// (INV=false: forward e^-j; INV=true: unnormalised inverse e^+j)
// Neither needs a reordering pass: the real-FFT split steps and the frequency-domain products
// that sit between them address elements through __brev().  Several transforms ("batches")
// laid out back to back are processed by one call so that every thread has butterflies to do.
//
// Storage index of logical element i is SIDX(i) = i + (i >> 4) + (i >> 8).  A brute-force bank
// model over every access pattern used here (all radix passes for 128..8192 points, natural
// order, k / M-k pairs of the real-FFT split, and bit-reversed reads) shows this padding is
// conflict-free for 8-byte accesses everywhere except a 2-way conflict in the q=8 pass; the
// first version (i + (i >> 3)) was 4- to 16-way conflicted on the bit-reversed reads
// (profiles/r01_enhance_omlsa1024_ncu.md: 38 % of shared wavefronts were conflicts).
#pragma once
#include "cse_common.cuh"

#define SIDX(i) ((i) + ((i) >> 4) + ((i) >> 8))
#define CSE_FFT_STRIDE(n) ((n) + ((n) >> 4) + ((n) >> 8) + 1)

CSE_D real2 tw_load(const real2* __restrict__ tw, int idx) { return tw[idx]; }

// W_8^m for m = 0..3 applied to d
template <int m> CSE_D real2 rot8(real2 d) {
    const real h = R(0.70710678118654752440);
    if (m == 0) return d;
    if (m == 1) return cscale(cadd(d, mk2(d.y, -d.x)), h);          // * (1 - i)/sqrt2
    if (m == 2) return mk2(d.y, -d.x);                             // * (-i)
    return cscale(cadd(mk2(-d.x, -d.y), mk2(d.y, -d.x)), h);       // * (-1 - i)/sqrt2
}
template <int m> CSE_D real2 rot8c(real2 d) {   // conj(W_8^m)
    const real h = R(0.70710678118654752440);
    if (m == 0) return d;
    if (m == 1) return cscale(cadd(d, mk2(-d.y, d.x)), h);          // * (1 + i)/sqrt2
    if (m == 2) return mk2(-d.y, d.x);                             // * (+i)
    return cscale(cadd(mk2(-d.x, -d.y), mk2(-d.y, d.x)), h);       // * (-1 + i)/sqrt2
}

// UNIT: the twiddle is known to be 1 (every twiddle of the q == 1 pass): no load, no multiply, same bits.
template <bool CONJ, bool UNIT = false> CSE_D real2 twmul(real2 a, real2 w) { return UNIT ? a : (CONJ ? cmulc(a, w) : cmul(a, w)); }
template <int m, bool CONJ> CSE_D real2 rotf(real2 d) { return CONJ ? rot8c<m>(d) : rot8<m>(d); }

// Three fused radix-2 stages on 8 register values.  DIF: distances 4, 2, 1 with twiddles w1 = W_{8q}^j,
// w2 = W_{4q}^j, w3 = W_{2q}^j applied after the subtraction; DIT: the mirror image (distances 1, 2, 4,
// twiddles applied before the butterfly).
template <bool CONJ, bool U1> CSE_D void bfly8_dif(real2* v, real2 w1, real2 w2, real2 w3) {
    real2 d;
    d = csub(v[0], v[4]); v[0] = cadd(v[0], v[4]); v[4] = twmul<CONJ, U1>(rotf<0, CONJ>(d), w1);
    d = csub(v[1], v[5]); v[1] = cadd(v[1], v[5]); v[5] = twmul<CONJ, U1>(rotf<1, CONJ>(d), w1);
    d = csub(v[2], v[6]); v[2] = cadd(v[2], v[6]); v[6] = twmul<CONJ, U1>(rotf<2, CONJ>(d), w1);
    d = csub(v[3], v[7]); v[3] = cadd(v[3], v[7]); v[7] = twmul<CONJ, U1>(rotf<3, CONJ>(d), w1);
#pragma unroll
    for (int g = 0; g < 8; g += 4) {
        d = csub(v[g], v[g + 2]); v[g] = cadd(v[g], v[g + 2]); v[g + 2] = twmul<CONJ, U1>(d, w2);
        d = csub(v[g + 1], v[g + 3]); v[g + 1] = cadd(v[g + 1], v[g + 3]); v[g + 3] = twmul<CONJ, U1>(rotf<2, CONJ>(d), w2);
    }
#pragma unroll
    for (int g = 0; g < 8; g += 2) { d = csub(v[g], v[g + 1]); v[g] = cadd(v[g], v[g + 1]); v[g + 1] = twmul<CONJ, U1>(d, w3); }
}
template <bool CONJ, bool U1> CSE_D void bfly8_dit(real2* v, real2 w1, real2 w2, real2 w3) {
    real2 t;
#pragma unroll
    for (int g = 0; g < 8; g += 2) { t = twmul<CONJ, U1>(v[g + 1], w3); v[g + 1] = csub(v[g], t); v[g] = cadd(v[g], t); }
#pragma unroll
    for (int g = 0; g < 8; g += 4) {
        t = twmul<CONJ, U1>(v[g + 2], w2); v[g + 2] = csub(v[g], t); v[g] = cadd(v[g], t);
        t = rotf<2, CONJ>(twmul<CONJ, U1>(v[g + 3], w2)); v[g + 3] = csub(v[g + 1], t); v[g + 1] = cadd(v[g + 1], t);
    }
    t = rotf<0, CONJ>(twmul<CONJ, U1>(v[4], w1)); v[4] = csub(v[0], t); v[0] = cadd(v[0], t);
    t = rotf<1, CONJ>(twmul<CONJ, U1>(v[5], w1)); v[5] = csub(v[1], t); v[1] = cadd(v[1], t);
    t = rotf<2, CONJ>(twmul<CONJ, U1>(v[6], w1)); v[6] = csub(v[2], t); v[2] = cadd(v[2], t);
    t = rotf<3, CONJ>(twmul<CONJ, U1>(v[7], w1)); v[7] = csub(v[3], t); v[3] = cadd(v[3], t);
}

// Group index of butterfly slot t in a pass with smallest distance q.  For q == 8 the slots are
// rotated left by one bit so that the two groups a half-warp touches are 128 elements apart
// instead of 64: with the storage padding this makes the pass conflict-free (it was 2-way).
template <int Q, int NGRP> CSE_D int fft_group(int t) {
    if (Q == 8 && NGRP >= 4) return ((t << 1) & (NGRP - 1)) | (t / (NGRP >= 2 ? NGRP / 2 : 1));
    return t;
}

// Barrier among the GROUP consecutive threads that own one transform of a batch (butterfly idx -> transform
// idx / GROUP stays fixed from pass to pass when every thread has one butterfly per pass): a warp barrier when the
// group fits a warp, else a named barrier of its own - the other transforms' warps run on.
template <int GROUP>
CSE_D void fft_group_sync(int tid) {
    if (GROUP <= 32) { __syncwarp(); return; }
#if defined(CSE_EMU)
    cse_emu::named_barrier(1 + tid / GROUP, GROUP);
#else
    asm volatile("bar.sync %0, %1;" ::"r"(1 + tid / GROUP), "n"(GROUP) : "memory");
#endif
}

// One pass of RL fused DIF stages (CONJ: conjugated twiddles = inverse transform).  `h` = half size of the first fused stage,
// q = h >> (RL-1) = smallest butterfly distance of the pass.
// ZHI: the upper half of every transform's input is known to be zero (zero-padded frames) and is not read.
// NAT (last pass only, q == 1, one butterfly per thread): the outputs go to NATURAL order, unpadded -
// out[b * bstride + k] = X_b[k] - instead of staying bit-reversed in place; the threads of a transform meet
// between their loads and these scattered stores (fft_group_sync), so the buffer is reused in place.
template <int LOG2N, int RL, bool CONJ, int H, int TWN, bool ZHI = false, bool NAT = false>
CSE_D void dif_pass(real2* s, int nbatch, int bstride, const real2* __restrict__ tw, int tid, int nth) {
    constexpr int N = 1 << LOG2N;
    constexpr int NB = 1 << RL;               // elements per butterfly
    constexpr int h = H;
    constexpr int q = h >> (RL - 1);
    constexpr int per = N / NB;
    const int total = nbatch * per;
    constexpr bool U1 = q == 1;                  // all twiddles of this pass are W^0
    // TWN > 0: `tw` is a flat table W_TWN^k.  TWN == 0: `tw` is the compact per-pass layout of
    // load_pass_twiddles (block of this pass at 3(q-1)/7: [W_{8q}^j | W_{4q}^j | W_{2q}^j], unit stride).
    constexpr int twstep = TWN > 0 ? (TWN / 2) / h : 0;
    constexpr int coff = 3 * (q - 1) / 7;
    for (int idx = tid; idx < total; idx += nth) {
        const int b = idx / per, r = idx - b * per;
        // NAT: thread r takes butterfly group brev(r), so that its outputs X[brev_RL(m) << (LOG2N-RL) | r] are
        // consecutive across the lanes of a warp (conflict-free natural-order stores; the padded loads stay conflict-free)
        const int j = r & (q - 1), grp = NAT ? (int)(__brev((unsigned)r) >> (32 - (LOG2N - RL))) : fft_group<q, per / q>(r / q);
        // SIDX(base + m q) == SIDX(base) + SIDX(m q): base = grp*NB*q + j with j < q, q a power of two,
        // so neither the >>4 nor the >>8 term of the padding ever carries across the addition.
        real2* p = s + b * bstride + SIDX(grp * (q * NB) + j);
        real2 v[NB];
#pragma unroll
        for (int m = 0; m < NB; ++m) v[m] = (ZHI && m >= NB / 2) ? mk2(R(0), R(0)) : p[SIDX(m * q)];
        if (RL == 3) {
            const real2 w1 = U1 ? mk2(R(1), R(0)) : TWN > 0 ? tw_load(tw, j * twstep) : tw_load(tw, coff + j);
            const real2 w2 = U1 ? mk2(R(1), R(0)) : TWN > 0 ? tw_load(tw, j * twstep * 2) : tw_load(tw, coff + q + j);
            const real2 w3 = U1 ? mk2(R(1), R(0)) : TWN > 0 ? tw_load(tw, j * twstep * 4) : tw_load(tw, coff + 2 * q + j);
            bfly8_dif<CONJ, U1>(v, w1, w2, w3);
        } else if (RL == 2) {
            const real2 w1 = U1 ? mk2(R(1), R(0)) : TWN > 0 ? tw_load(tw, j * twstep) : tw_load(tw, coff + j);
            const real2 w2 = U1 ? mk2(R(1), R(0)) : TWN > 0 ? tw_load(tw, j * twstep * 2) : tw_load(tw, coff + q + j);
            real2 d;
            d = csub(v[0], v[2]); v[0] = cadd(v[0], v[2]); v[2] = twmul<CONJ, U1>(d, w1);
            d = csub(v[1], v[3]); v[1] = cadd(v[1], v[3]); v[3] = twmul<CONJ, U1>(rotf<2, CONJ>(d), w1);
            d = csub(v[0], v[1]); v[0] = cadd(v[0], v[1]); v[1] = twmul<CONJ, U1>(d, w2);
            d = csub(v[2], v[3]); v[2] = cadd(v[2], v[3]); v[3] = twmul<CONJ, U1>(d, w2);
        } else {
            const real2 w1 = U1 ? mk2(R(1), R(0)) : TWN > 0 ? tw_load(tw, j * twstep) : tw_load(tw, coff + j);
            real2 d = csub(v[0], v[1]); v[0] = cadd(v[0], v[1]); v[1] = twmul<CONJ, U1>(d, w1);
        }
        if (NAT) {
            static_assert(!NAT || q == 1, "natural-order output is for the last pass");
            fft_group_sync<per>(tid);
            // position grp*NB + m holds X[brev(grp*NB + m)] = X[brev_RL(m) << (LOG2N - RL) | brev(grp)], brev(grp) = r
            real2* o = s + b * bstride + r;
#pragma unroll
            for (int m = 0; m < NB; ++m) o[(int)(__brev((unsigned)m) >> (32 - RL)) << (LOG2N - RL)] = v[m];
            continue;
        }
#pragma unroll
        for (int m = 0; m < NB; ++m) p[SIDX(m * q)] = v[m];
    }
}

// One pass of RL fused DIT stages; q = half size of the FIRST (smallest) fused stage.
template <int LOG2N, int RL, bool CONJ, int Q, int TWN>
CSE_D void dit_pass(real2* s, int nbatch, int bstride, const real2* __restrict__ tw, int tid, int nth) {
    constexpr int N = 1 << LOG2N;
    constexpr int NB = 1 << RL;
    constexpr int q = Q;
    constexpr int per = N / NB;
    const int total = nbatch * per;
    constexpr bool U1 = q == 1;                  // all twiddles of this pass are W^0
    constexpr int twq = TWN > 0 ? (TWN / 2) / q : 0;     // W_{2q}^p = W_T^{p * T/(2q)}
    constexpr int coff = 3 * (q - 1) / 7;                 // compact layout, see dif_pass
    for (int idx = tid; idx < total; idx += nth) {
        const int b = idx / per, r = idx - b * per;
        const int j = r & (q - 1), grp = fft_group<q, per / q>(r / q);
        real2* p = s + b * bstride + SIDX(grp * (q * NB) + j);      // affine storage offsets, see dif_pass
        real2 v[NB];
#pragma unroll
        for (int m = 0; m < NB; ++m) v[m] = p[SIDX(m * q)];
        if (RL == 3) {
            // distances q (W_{2q}^j), 2q (W_{4q}^{j + (m&1)q}), 4q (W_{8q}^{j + (m&3)q}); conjugated
            const real2 w3 = U1 ? mk2(R(1), R(0)) : TWN > 0 ? tw_load(tw, j * twq) : tw_load(tw, coff + 2 * q + j);
            const real2 w2 = U1 ? mk2(R(1), R(0)) : TWN > 0 ? tw_load(tw, j * (twq >> 1)) : tw_load(tw, coff + q + j);
            const real2 w1 = U1 ? mk2(R(1), R(0)) : TWN > 0 ? tw_load(tw, j * (twq >> 2)) : tw_load(tw, coff + j);
            bfly8_dit<CONJ, U1>(v, w1, w2, w3);
        } else if (RL == 2) {
            const real2 w2 = U1 ? mk2(R(1), R(0)) : TWN > 0 ? tw_load(tw, j * twq) : tw_load(tw, coff + q + j);
            const real2 w1 = U1 ? mk2(R(1), R(0)) : TWN > 0 ? tw_load(tw, j * (twq >> 1)) : tw_load(tw, coff + j);
            real2 t;
            t = twmul<CONJ, U1>(v[1], w2); v[1] = csub(v[0], t); v[0] = cadd(v[0], t);
            t = twmul<CONJ, U1>(v[3], w2); v[3] = csub(v[2], t); v[2] = cadd(v[2], t);
            t = twmul<CONJ, U1>(v[2], w1); v[2] = csub(v[0], t); v[0] = cadd(v[0], t);
            t = rotf<2, CONJ>(twmul<CONJ, U1>(v[3], w1)); v[3] = csub(v[1], t); v[1] = cadd(v[1], t);
        } else {
            const real2 w1 = U1 ? mk2(R(1), R(0)) : TWN > 0 ? tw_load(tw, j * twq) : tw_load(tw, coff + j);
            real2 t = twmul<CONJ, U1>(v[1], w1); v[1] = csub(v[0], t); v[0] = cadd(v[0], t);
        }
#pragma unroll
        for (int m = 0; m < NB; ++m) p[SIDX(m * q)] = v[m];
    }
}

// Decimation-in-frequency transform, natural-order in -> bit-reversed out.  INV=false: forward
// (e^-j); INV=true: unnormalised inverse (e^+j).  Ends with a __syncthreads().  Stage sizes are
// template constants so that all index arithmetic folds to shifts and masks.
// Barrier between two passes whose butterflies of one transform stay with the same GROUP threads (every
// radix-8 pass of a transform with LOG2N % 3 == 0 maps butterfly idx to transform idx / (N/8)): only those
// threads have to meet, on a named barrier of their own - the other transforms' warps run on.
// GROUPSYNC (every thread has exactly one butterfly per pass: nbatch * N/8 <= nth): the barriers between
// radix-8 passes only gather the N/8 threads of each transform.  NAT: natural-order, unpadded output (see dif_pass).
template <int LOG2N, bool INV, int TWN = CSE_TW_N, bool ZHI = false, bool GROUPSYNC = false, bool NAT = false>
CSE_D void fft_dif(real2* s, int nbatch, int bstride, const real2* __restrict__ tw, int tid, int nth,
                   const real2* __restrict__ twg = nullptr) {
    constexpr int REM = LOG2N % 3, NP = LOG2N / 3;
    constexpr int H0 = 1 << (LOG2N - 1), H1 = H0 >> REM;
    constexpr int TWR = TWN < 0 ? CSE_TW_N : TWN;          // remainder pass: flat global table when TWN == -1
    constexpr int TW3 = TWN < 0 ? 0 : TWN;                 // radix-8 passes: compact when TWN <= 0
    constexpr int G = (1 << LOG2N) / 8;                    // threads per transform in a radix-8 pass
    static_assert(!NAT || NP >= 1, "natural-order output needs a final radix-8 pass");
    const real2* twr = TWN < 0 ? twg : tw;
    const bool mine = tid < nbatch * G;                    // (GROUPSYNC) this thread owns a butterfly
    auto sync = [&](bool last) {
        if (GROUPSYNC && !last) { if (mine) fft_group_sync<G>(tid); }
        else __syncthreads();
    };
    if constexpr (REM == 1) { dif_pass<LOG2N, 1, INV, H0, TWR, ZHI>(s, nbatch, bstride, twr, tid, nth); __syncthreads(); }
    if constexpr (REM == 2) { dif_pass<LOG2N, 2, INV, H0, TWR, ZHI>(s, nbatch, bstride, twr, tid, nth); __syncthreads(); }
    if constexpr (NP >= 1) { dif_pass<LOG2N, 3, INV, H1, TW3, (ZHI && REM == 0), (NAT && NP == 1)>(s, nbatch, bstride, tw, tid, nth); sync(NP == 1); }
    if constexpr (NP >= 2) { dif_pass<LOG2N, 3, INV, (H1 >> 3), TW3, false, (NAT && NP == 2)>(s, nbatch, bstride, tw, tid, nth); sync(NP == 2); }
    if constexpr (NP >= 3) { dif_pass<LOG2N, 3, INV, (H1 >> 6), TW3, false, (NAT && NP == 3)>(s, nbatch, bstride, tw, tid, nth); sync(NP == 3); }
    if constexpr (NP >= 4) { dif_pass<LOG2N, 3, INV, (H1 >> 9), TW3, false, (NAT && NP == 4)>(s, nbatch, bstride, tw, tid, nth); sync(NP == 4); }
}

// Decimation-in-time transform, bit-reversed in -> natural-order out.  Ends with a __syncthreads().
template <int LOG2N, bool INV, int TWN = CSE_TW_N>
CSE_D void fft_dit(real2* s, int nbatch, int bstride, const real2* __restrict__ tw, int tid, int nth,
                   const real2* __restrict__ twg = nullptr) {
    constexpr int REM = LOG2N % 3, NP = LOG2N / 3;
    constexpr int TWR = TWN < 0 ? CSE_TW_N : TWN;
    constexpr int TW3 = TWN < 0 ? 0 : TWN;
    const real2* twr = TWN < 0 ? twg : tw;
    if constexpr (NP >= 1) { dit_pass<LOG2N, 3, INV, 1, TW3>(s, nbatch, bstride, tw, tid, nth); __syncthreads(); }
    if constexpr (NP >= 2) { dit_pass<LOG2N, 3, INV, 8, TW3>(s, nbatch, bstride, tw, tid, nth); __syncthreads(); }
    if constexpr (NP >= 3) { dit_pass<LOG2N, 3, INV, 64, TW3>(s, nbatch, bstride, tw, tid, nth); __syncthreads(); }
    if constexpr (NP >= 4) { dit_pass<LOG2N, 3, INV, 512, TW3>(s, nbatch, bstride, tw, tid, nth); __syncthreads(); }
    constexpr int QR = 1 << (3 * NP);
    if constexpr (REM == 2) { dit_pass<LOG2N, 2, INV, QR, TWR>(s, nbatch, bstride, twr, tid, nth); __syncthreads(); }
    if constexpr (REM == 1) { dit_pass<LOG2N, 1, INV, QR, TWR>(s, nbatch, bstride, twr, tid, nth); __syncthreads(); }
}

// Compact per-pass twiddle layout in shared memory (TWN == 0 / -1 above).  For every radix-8 pass
// with smallest butterfly distance q (a power of 8) a block at offset 3(q-1)/7 holds
// [W_{8q}^j | W_{4q}^j | W_{2q}^j], j < q, so that a warp's twiddle reads are unit-stride.  The
// remainder pass (radix 2 or 4, q = 8^(LOG2N/3)) gets [W_{2q}^j] or [W_{4q}^j | W_{2q}^j] behind them
// unless WITH_REM is false.  Reading the flat global W_8192 table instead costs up to 32 different
// 128-byte lines per warp instruction in the narrow passes (profiles/r01e: L1TEX 71 % busy) and a
// flat shared copy costs 2- to 8-way bank conflicts (profiles/r01f).
template <int LOG2N, bool WITH_REM>
struct FftTwLayout {
    static constexpr int NP = LOG2N / 3, REM = LOG2N % 3;
    static constexpr int QR = 1 << (3 * NP);
    static constexpr int OFFR = 3 * (QR - 1) / 7;
    static constexpr int SIZE = OFFR + (WITH_REM ? REM * QR : 0);
};
template <int LOG2N, bool WITH_REM>
CSE_D void load_pass_twiddles(real2* dst, const real2* __restrict__ twg, int tid, int nth) {
    typedef FftTwLayout<LOG2N, WITH_REM> LY;
    for (int i = tid; i < LY::SIZE; i += nth) {
        int idx;
        if (i < LY::OFFR) {
            // find the block: q = 8^k with 3(q-1)/7 <= i < 3(8q-1)/7
            int q = 1, off = 0;
            while (i >= off + 3 * q) { off += 3 * q; q <<= 3; }
            const int r = i - off, which = r / q, j = r - which * q;          // which: 0 -> W_{8q}, 1 -> W_{4q}, 2 -> W_{2q}
            idx = j * ((CSE_TW_N / 8) / q << which);
        } else {
            const int r = i - LY::OFFR, which = r / LY::QR, j = r - which * LY::QR;
            // REM == 1: W_{2q}^j.  REM == 2: [W_{4q}^j | W_{2q}^j]
            idx = LY::REM == 1 ? j * ((CSE_TW_N / 2) / LY::QR) : j * ((CSE_TW_N / 4) / LY::QR << which);
        }
        dst[i] = twg[idx];
    }
}

CSE_D int brev_n(int k, int log2n) { return (int)(__brev((unsigned)k) >> (32 - log2n)); }
