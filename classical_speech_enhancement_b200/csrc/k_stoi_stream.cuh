// K5b (candidate side): SNR + STOI in ONE streaming pass over the candidate waveform.
//
// Restates calculate_snr (Code/evaluation_metrics.py:39-58) and pystoi.stoi(extended=False)
// (Code/evaluation_metrics.py:30-36) for one finalized candidate, as k_score.cuh describes, but
// without the global 10 kHz scratch signal of the first version (profiles/r01f: 491 KB of DRAM
// traffic per candidate against 192 KB algorithmic).  The waveform is read exactly once:
//
//   for each tile of 4096 input samples (2560 output samples = 20 hop-blocks of 128):
//     1. load the tile (+ filter margins) de-interleaved into shared memory; the same loads feed the
//        SNR sums and the finite check of finalize_enhanced;
//     2. polyphase-resample it into a shared-memory tile (packed two-group FFMA2 form);
//     3. fold the needed hop-blocks into the overlap-added blocks of the silence-removed signal,
//          B_j[n] = w[n] y10[128 kept[j] + n] + w[128+n] y10[128 kept[j-1] + 128 + n],
//        kept in a ring of 32 blocks (hop-block -> (j, j') map and running kept count cached per
//        utterance by the clean-side pass);
//     4. as soon as 9 consecutive blocks are complete, transform 8 frames  w .* (B_m ++ B_{m+1})
//        (batched 256-point complex DIF FFTs), take third-octave band envelopes.
//   then the 30-frame segment correlation against the cached clean statistics.
// Only the 15 x K band envelopes (14 KB) go through a global scratch row (they are re-read into
// shared memory for the correlation when they fit).
#pragma once
#include "k_score.cuh"

#define CSE_STOI_RB 32                 // ring of overlap-added blocks
#define CSE_STOI_HBT (5 * CSE_RS_A2 / 128)   // hop-blocks produced per resampler tile (20)

__global__ void __launch_bounds__(256, 4) stoi_stream_kernel(ScoreArgs a) {
    constexpr int T = CSE_STOI_T, BST = CSE_FFT_STRIDE(256), NK = CSE_STOI_K1 - CSE_STOI_K0, NT = 256;
    constexpr int H = CSE_RS_A2 / 2, AP2 = H + 18 /* 17 needed; 18 makes the de-interleaving stores 2-way instead of 4-way conflicted */, RB = CSE_STOI_RB, HBT = CSE_STOI_HBT, YT = 5 * CSE_RS_A2;
    CSE_DYN_SMEM(smem_raw);
    const ScoreGeom& g = a.g;
    const int tid = threadIdx.x, li = blockIdx.x, item = a.item0 + li;
    const int u = item / a.per_utt;
    const unsigned char* rec = a.cache + (size_t)u * g.bytes;
    const CleanHeader* hdr = reinterpret_cast<const CleanHeader*>(rec);
    const int* __restrict__ hbmap = reinterpret_cast<const int*>(rec + g.off_hbmap);   // [nfr + 2][3]: jA, jB, count
    const real* __restrict__ xtob = reinterpret_cast<const real*>(rec + g.off_xtob);
    const real* __restrict__ seg_c = reinterpret_cast<const real*>(rec + g.off_seg);
    const unsigned char* __restrict__ need = rec + g.off_need;
    const int K = hdr->K, Kf = K > 0 ? K - 1 : 0, J = hdr->J;
    const int nhb = g.nfr + 1;

    double* scratch = reinterpret_cast<double*>(smem_raw);                    // 40 doubles
    real2* fbuf = reinterpret_cast<real2*>(scratch + 40);                     // T * BST; input tile during resampling
    real* yt = reinterpret_cast<real*>(fbuf + T * BST);                       // YT resampled samples; band powers during FFT batches
    real* ring = yt + YT;                                                     // RB * 128
    real* w_s = ring + RB * 128;                                              // 256
    real2* tws = reinterpret_cast<real2*>(w_s + 256);                         // 160
    real2* ktw_s = tws + 160;                                                 // NK: -i W_512^k of the real-FFT split
    unsigned* koff_s = reinterpret_cast<unsigned*>(ktw_s + NK);               // NK: storage offsets of bins k | 256-k
    real2* xs2 = fbuf;
    real* pw = yt;

    for (int i = tid; i < 256; i += NT) w_s[i] = a.T->stoi_win[i];
    load_pass_twiddles<8, true>(tws, a.T->tw, tid, NT);
    for (int i = tid; i < NK; i += NT) {
        const int k = CSE_STOI_K0 + i;
        const real2 wk = tw_load(a.T->tw, k * (CSE_TW_N / 512));
        ktw_s[i] = mk2(wk.y, -wk.x);                                          // -i W_512^k
        koff_s[i] = (unsigned)SIDX(brev_n(k, 8)) | ((unsigned)SIDX(brev_n(256 - k, 8)) << 16);
    }
    // Band sums are done by one warp per frame: lane = a run of <= 9 consecutive bins inside one
    // band (29 runs cover the 15 bands); the first lane of a band adds its neighbours' partials.
    // The run descriptor (first bin | length << 8 | band << 12 | runs of the band if first << 16) sits in shared memory.
    static_assert(T * 32 == NT, "one warp per frame of a batch");
    unsigned* run_s = koff_s + NK;                                            // 32
    int* hb_s2 = reinterpret_cast<int*>(run_s + 32);                          // 2 x 64: this tile's (HBT + 1) rows of hbmap, double-buffered by tile parity
    if (tid < 32) {
        unsigned desc = 0;
        int r0 = 0;
        for (int b = 0; b < CSE_NBANDS; ++b) {
            const int e0 = a.T->stoi_edges[b], wd = a.T->stoi_edges[b + 1] - e0, n = (wd + 8) / 9;
            if (tid >= r0 && tid < r0 + n) {
                const int i = tid - r0, lo = (i * wd) / n, hi = ((i + 1) * wd) / n;
                desc = (unsigned)(e0 + lo - CSE_STOI_K0) | ((unsigned)(hi - lo) << 8) | ((unsigned)b << 12) | ((unsigned)(i == 0 ? n : 0) << 16);
            }
            r0 += n;
        }
        run_s[tid] = desc;
    }
    const real* __restrict__ sig = a.wav + (size_t)li * g.L;
    const real* __restrict__ cl = a.clean + (size_t)u * g.L;
    real* ytob_g = a.y10 + (size_t)li * score_row_reals(g.nfrm);   // scratch row: 15 * Kf band envelopes
    const int lag = a.lagflags[2 * li];
    int flags = a.lagflags[2 * li + 1];
    const bool fin = a.finalize != 0;
    const int L = g.L, n10 = g.n10;
    const bool do_stoi = Kf >= CSE_NSEG;
    __syncthreads();

    real pn = R(0);
    int bad = 0, m_next = 0;
    const int na = (n10 + 4) / 5;
    for (int a0 = 0, t = 0; a0 < na; a0 += CSE_RS_A2, ++t) {
        int* hb_s = hb_s2 + 64 * (t & 1);
        int gap = 0;
        // ---- 1. input tile, SNR sums and finite check on the samples this tile owns
        {
            constexpr int TOT = 8 * (CSE_RS_A2 + 17), PERT = (TOT + NT - 1) / NT;
            static_assert(NT == 256 && H == 256 && PERT == 17, "tile-load index algebra below assumes 256 threads, 512 groups");
            const int j0 = 8 * a0 - 64;
            if (tid < 3 * (HBT + 1)) {
                const int row = HBT * t + tid / 3;
                hb_s[tid] = row <= nhb ? hbmap[3 * HBT * t + tid] : -1;
            }
            if (tid < HBT && HBT * t + tid < nhb) {            // does a kept frame start after a removed one in this tile?
                const int ja = hbmap[3 * (HBT * t + tid)], jb = hbmap[3 * (HBT * t + tid) + 1];
                gap = ja >= 0 && ja != jb;
            }
            if (tid >= 128 && j0 + TOT + (tid - 128) * (128 / (int)sizeof(real)) < L)      // next tile's lines start moving from DRAM to L2
                cse_prefetch_l2(sig + j0 + TOT + (tid - 128) * (128 / (int)sizeof(real)));
            const bool interior = lag == 0 && j0 >= 0 && j0 + TOT <= L;      // uniform: no bounds, no shift
            real raw[PERT], cv[PERT];
            if (interior) {
                const real* __restrict__ sp = sig + j0 + tid;
                const real* __restrict__ cp = cl + j0 + tid;
#pragma unroll
                for (int k = 0; k < PERT; ++k) {
                    raw[k] = (k < PERT - 1 || tid + k * NT < TOT) ? sp[k * NT] : R(0);
                    cv[k] = (k >= 1 && k < PERT - 1) || (k == 0 && tid >= 64) || (k == PERT - 1 && tid < 64) ? cp[k * NT] : R(0);
                }
            } else {
#pragma unroll
                for (int k = 0; k < PERT; ++k) {
                    const int jj = tid + k * NT, j = j0 + jj;
                    const bool own = jj >= 64 && jj < 64 + 8 * CSE_RS_A2 && j < L;        // each sample owned by one tile
                    raw[k] = jj < TOT ? xraw(sig, j, lag, L) : R(0);
                    cv[k] = own ? cl[j] : R(0);
                }
            }
            // jj = tid + 256 k  ->  c = tid & 7, column ap = (tid >> 3) + 32 k: columns < 256 feed the .x half of
            // the pair tile, columns >= 256 the .y half (column - 256); the 18 columns of overlap feed both.
            const int c = tid & 7, ap0 = tid >> 3;
            real2* row = xs2 + c * AP2 + ap0;
#pragma unroll
            for (int k = 0; k < PERT; ++k) {
                const int jj = tid + k * NT, j = j0 + jj;
                const real v = fin ? r_clip(raw[k], R(-1), R(1)) : raw[k];
                const bool own = interior ? ((k >= 1 && k < PERT - 1) || (k == 0 && tid >= 64) || (k == PERT - 1 && tid < 64))
                                          : (jj >= 64 && jj < 64 + 8 * CSE_RS_A2 && j < L);
                if (own) {
                    if (!r_finite(raw[k])) bad = 1;
                    const real d = cv[k] - v;
                    pn = r_fma(d, d, pn);
                }
                if (k < 8) row[32 * k].x = v;
                else if (k == 8) { if (ap0 + 256 < AP2) row[256].x = v; row[0].y = v; }
                else if (k < PERT - 1 || jj < TOT) row[32 * (k - 8)].y = v;
            }
        }
        const int anygap = __syncthreads_or(gap);
        if (!do_stoi) continue;                                                   // uniform: SNR only (the next tile load
                                                                                  // only overwrites what nobody reads)
        // ---- 2. resample into the shared tile: yt[5 (a - a0) + p]
        for (int al = tid; al < H; al += NT) {
            const int alo = a0 + al, ahi = alo + H;
            const bool want_lo = 5 * alo < n10 && need[alo], want_hi = 5 * ahi < n10 && need[ahi];
            if (want_lo || want_hi) {
                real2 acc[5];
#pragma unroll
                for (int p = 0; p < 5; ++p) acc[p] = mk2(R(0), R(0));
#pragma unroll
                for (int jj = 6; jj <= 128; ++jj) {
                    const real2 x = xs2[(jj & 7) * AP2 + al + (jj >> 3)];
#pragma unroll
                    for (int p = 0; p < 5; ++p) {
                        const int idx = 8 * p + 610 - 5 * jj;
                        if (idx >= 0 && idx <= 580) acc[p] = cfma2(x, mk2(c_rs5[jj * 5 + p], c_rs5[jj * 5 + p]), acc[p]);
                    }
                }
#pragma unroll
                for (int p = 0; p < 5; ++p) { yt[5 * al + p] = acc[p].x; yt[5 * (al + H) + p] = acc[p].y; }
            }
        }
        __syncthreads();
        // ---- 3. fold the tile's hop-blocks into the overlap-added blocks B_j.  A block's first
        // contribution (the second half of kept frame j-1) ASSIGNS, the other one ADDS; when both come
        // from one hop-block they are written together, otherwise the add runs after a barrier.
        {
            static_assert(NT == 256 && HBT % 2 == 0, "each half of the CTA takes every other hop-block; a thread keeps one sample index n");
            const int n = tid & 127, par = tid >> 7;
            const real w0 = w_s[n], w1 = w_s[128 + n];
#pragma unroll
            for (int i = 0; i < HBT / 2; ++i) {
                const int hl = 2 * i + par;
                const int ja = hb_s[3 * hl], jb = hb_s[3 * hl + 1];
                if (HBT * t + hl < nhb && jb >= 0 && jb < K) {
                    const real y = yt[128 * hl + n];
                    real v = w1 * y;
                    if (ja == jb) v = r_fma(w0, y, v);        // consecutive kept frames share the hop-block
                    ring[(jb & (RB - 1)) * 128 + n] = v;
                }
            }
            __syncthreads();
            if (anygap) {                                     // uniform; rare: only where speech resumes after removed frames
#pragma unroll
                for (int i = 0; i < HBT / 2; ++i) {
                    const int hl = 2 * i + par;
                    const int ja = hb_s[3 * hl], jb = hb_s[3 * hl + 1];
                    if (HBT * t + hl < nhb && ja >= 0 && ja != jb) {      // first half of kept frame ja completes block ja
                        const int slot = (ja & (RB - 1)) * 128 + n;
                        ring[slot] = r_fma(w0, yt[128 * hl + n], ja > 0 ? ring[slot] : R(0));
                    }
                }
                __syncthreads();
            }
        }
        // ---- 4. transform every batch of T frames whose blocks are complete
        const int hb_end = HBT * (t + 1) < nhb ? HBT * (t + 1) : nhb;
        const int jdone = hb_s[3 * (hb_end - HBT * t) + 2];              // kept frames with index < hb_end -> blocks B_0..B_{jdone-1} complete
        const bool last = a0 + CSE_RS_A2 >= na;
        while (m_next < Kf && (m_next + T <= jdone - 1 || last)) {
            const int m0 = m_next;
            {
                // thread = one packed sample pair mm of every other frame: window pair and block half are fixed
                static_assert(NT == 256 && T == 8, "frame load: 128 pairs x 2 frame parities");
                const int mm = tid & 127, par = tid >> 7;
                const real2 wv = *reinterpret_cast<const real2*>(w_s + 2 * mm);
                const int nn = (2 * mm) & 127, joff = mm < 64 ? 0 : 1;
                real2* dst = fbuf + SIDX(mm);
#pragma unroll
                for (int i = 0; i < T / 2; ++i) {
                    const int f = 2 * i + par, m = m0 + f;
                    real2 v = mk2(R(0), R(0));
                    if (m < Kf) v = cfma2(wv, *reinterpret_cast<const real2*>(ring + ((m + joff) & (RB - 1)) * 128 + nn), v);
                    dst[f * BST] = v;                 // upper half of the zero-padded frame: never stored, the first pass knows
                }
            }
            __syncthreads();
            fft_dif<8, false, 0, true>(fbuf, T, BST, tws, tid, NT);
            {
                const int lane = tid & 31, wf = tid >> 5;
                const unsigned desc = run_s[lane];
                const int run_lo = desc & 0xff, run_len = (desc >> 8) & 0xf, run_band = (desc >> 12) & 0xf, run_lead = desc >> 16;
                const real2* zf = fbuf + wf * BST;
                real* pwf = pw + wf * NK;
                // chunks of 32 ALIGNED bins: the bit-reversed reads of a warp then differ only in their top
                // five address bits, which the storage padding spreads over all banks
#pragma unroll
                for (int c = 0; c < (CSE_STOI_K1 + 31) / 32; ++c) {
                    const int i = lane + 32 * c - CSE_STOI_K0;
                    if (i >= 0 && i < NK) {
                        const unsigned off = koff_s[i];
                        // X[k] = (S + D (-i W^k)) / 2 with S = z0 + conj z1, D = z0 - conj z1; the 1/2 joins the square root
                        const real2 z0 = zf[off & 0xffffu], z1 = zf[off >> 16];
                        const real2 cz1 = mk2(z1.x, -z1.y);
                        const real2 X = cadd(cadd(z0, cz1), cmul(csub(z0, cz1), ktw_s[i]));
                        pwf[i] = r_fma(X.x, X.x, X.y * X.y);
                    }
                }
                __syncwarp();
                real sacc = R(0);
#pragma unroll
                for (int i = 0; i < 9; ++i) if (i < run_len) sacc += pwf[run_lo + i];
#pragma unroll
                for (int d = 1; d < 5; ++d) {
                    const real o = __shfl_down_sync(0xffffffffu, sacc, d);
                    if (d < run_lead) sacc += o;
                }
                const int m = m0 + wf;
                if (run_lead > 0 && m < Kf) ytob_g[run_band * Kf + m] = R(0.5) * r_sqrt(sacc);
            }
            __syncthreads();
            m_next += T;
        }
    }

    // ---- SNR, validity
    const double nbad = block_sum<double>((double)bad, scratch);
    const double pnoise = block_sum<double>((double)pn, scratch);
    if (nbad > 0.0) {
        if (tid == 0) { cse_score_t sc; sc.stoi = R(0); sc.snr = R(0); sc.lag = lag; sc.flags = 0; a.scores[item] = sc; }
        return;
    }
    if (tid == 0) {
        real snr;
        if (pnoise == 0.0) { snr = cse_inf(); flags |= CSE_FLAG_SNR_INF; }
        else snr = (real)(10.0 * log10(hdr->energy / (pnoise + 1e-10)));
        a.scores[item].snr = snr;
        a.scores[item].lag = lag;
    }
    if (!do_stoi) {
        if (tid == 0) { a.scores[item].stoi = R(1e-5); a.scores[item].flags = flags | CSE_FLAG_STOI_SHORT; }
        return;
    }
    // ---- segment correlation; envelopes come back into shared memory when they fit
    __threadfence_block();
    __syncthreads();
    real* ysm = reinterpret_cast<real*>(fbuf);
    const int nenv = CSE_NBANDS * Kf;
    const bool fits = 2 * nenv <= 2 * T * BST + YT + RB * 128;              // fbuf + yt + ring are free now
    if (fits) {
        for (int i = tid; i < nenv; i += NT) { ysm[i] = ytob_g[i]; ysm[nenv + i] = xtob[i]; }
        __syncthreads();
    }
    const real* ytob = fits ? ysm : ytob_g;
    const real* xenv = fits ? ysm + nenv : xtob;
    const real EPS = R(2.220446049250313e-16);
    const real clipc = R(1) + R(5.623413251903491);
    real dsum = R(0);
    for (int idx = tid; idx < J * CSE_NBANDS; idx += NT) {
        const int b = idx / J, j = idx - b * J;
        const real* x = xenv + b * Kf + j;
        const real* y = ytob + b * Kf + j;
        const real xn = seg_c[idx], xmean = seg_c[(size_t)J * CSE_NBANDS + idx], xinv = seg_c[(size_t)2 * J * CSE_NBANDS + idx];
        real yr[CSE_NSEG];                                                  // the segment stays in registers
        real y2 = R(0);
#pragma unroll
        for (int n = 0; n < CSE_NSEG; ++n) { yr[n] = y[n]; y2 = r_fma(yr[n], yr[n], y2); }
        const real alpha = xn / (r_sqrt(y2) + EPS);
        real s1 = R(0);
#pragma unroll
        for (int n = 0; n < CSE_NSEG; ++n) { yr[n] = r_min(alpha * yr[n], clipc * x[n]); s1 += yr[n]; }
        const real ymean = s1 / R(CSE_NSEG);
        real c2 = R(0), cx = R(0);
#pragma unroll
        for (int n = 0; n < CSE_NSEG; ++n) {
            const real d = yr[n] - ymean;
            c2 = r_fma(d, d, c2);
            cx = r_fma(d, x[n] - xmean, cx);
        }
        dsum += cx * xinv / (r_sqrt(c2) + EPS);
    }
    const double dtot = block_sum<double>((double)dsum, scratch);
    if (tid == 0) {
        a.scores[item].stoi = (real)(dtot / ((double)J * CSE_NBANDS));
        a.scores[item].flags = flags;
    }
}
