// K2: noise-PSD estimators of Code/noise_estimation.py.
//
//   percentile   (:20-56)  frame log-energy -> k quietest frames -> per-bin percentile (numpy
//                          "linear" interpolation) over those frames, floored at 0.02*median_t(P)
//   min_tracking (:64-99)  per-bin IIR smoothing over frames, centred running minimum
//                          (scipy.ndimage.minimum_filter1d, mode="nearest"), floored at
//                          0.01*median_t(P)
//
// A CTA owns a tile of BPC bins of one utterance and keeps their whole frame series in shared
// memory; one warp per bin does the order statistics with an in-place bitonic sort (series are
// padded with +inf to a power of two).  These run once per (utterance, STFT shape, method,
// percentile, eps) - amortised over hundreds of candidates - so the layout favours simplicity:
// reads are 32-byte sectors (8 adjacent bins per frame row).
#pragma once
#include "cse_common.cuh"

// ascending in-place bitonic sort of a[0..n) (n a power of two) by one warp
CSE_D void warp_bitonic_sort(real* a, int n, int lane) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < n; i += 32) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const real x = a[i], y = a[ixj];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { a[i] = y; a[ixj] = x; }
                }
            }
            __syncwarp();
        }
    }
}

// The same network with the series in registers: lane l holds elements l*EPL .. l*EPL+EPL-1 of a
// sequence of 32*EPL values.  Compare-exchange distances below EPL stay inside a lane, the others are
// one shuffle per element - about a quarter of the instructions of the shared-memory version, which
// matters because the medians make these kernels 2 % of a full sweep (profiles/r01_bench_launches).
template <int EPL> CSE_D void warp_sort_regs(real (&v)[EPL], int lane) {
#pragma unroll
    for (int k = 2; k <= 32 * EPL; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= EPL) {
                const int lj = j / EPL;
                const bool lower = (lane & lj) == 0;
#pragma unroll
                for (int e = 0; e < EPL; ++e) {
                    const real o = __shfl_xor_sync(0xffffffffu, v[e], lj);
                    const bool up = ((lane * EPL + e) & k) == 0;
                    v[e] = (lower == up) ? r_min(v[e], o) : r_max(v[e], o);
                }
            } else {
#pragma unroll
                for (int e = 0; e < EPL; ++e) {
                    if ((e & j) == 0) {
                        const bool up = ((lane * EPL + e) & k) == 0;
                        const real lo = r_min(v[e], v[e | j]), hi = r_max(v[e], v[e | j]);
                        v[e] = up ? lo : hi;
                        v[e | j] = up ? hi : lo;
                    }
                }
            }
        }
    }
}
template <int EPL> CSE_D real warp_pick(const real (&v)[EPL], int i, int lane) {     // element i (warp-uniform) of the sequence
    real x = v[0];
#pragma unroll
    for (int e = 1; e < EPL; ++e) x = (i % EPL) == e ? v[e] : x;
    return __shfl_sync(0xffffffffu, x, i / EPL);
}
template <int EPL> CSE_D void sorted_pair_regs(const real* a, int n, int lane, int i0, int i1, real& x0, real& x1) {
    real v[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) { const int i = lane * EPL + e; v[e] = i < n ? a[i] : cse_inf(); }
    warp_sort_regs<EPL>(v, lane);
    x0 = warp_pick<EPL>(v, i0, lane);
    x1 = warp_pick<EPL>(v, i1, lane);
}
// Order statistics i0, i1 of a[0..n) (n <= n_pad = a power of two >= 32; a[n..n_pad) is +inf).  Series of up
// to 1024 values are sorted in registers and leave `a` untouched; longer ones are sorted in place.
CSE_D void sorted_pair(real* a, int n, int n_pad, int lane, int i0, int i1, real& x0, real& x1) {
    switch (n_pad) {
        case 32: sorted_pair_regs<1>(a, n, lane, i0, i1, x0, x1); return;
        case 64: sorted_pair_regs<2>(a, n, lane, i0, i1, x0, x1); return;
        case 128: sorted_pair_regs<4>(a, n, lane, i0, i1, x0, x1); return;
        case 256: sorted_pair_regs<8>(a, n, lane, i0, i1, x0, x1); return;
        case 512: sorted_pair_regs<16>(a, n, lane, i0, i1, x0, x1); return;
        case 1024: sorted_pair_regs<32>(a, n, lane, i0, i1, x0, x1); return;
        default: break;
    }
    warp_bitonic_sort(a, n_pad, lane);
    x0 = a[i0];
    x1 = a[i1];
}
// np.median of n values from its two middle order statistics (equal when n is odd)
CSE_D real median_of(real lo, real hi) { return (lo + hi) * R(0.5); }

// mean over bins of log(max(P, eps)) per frame, accumulated in double (the ordering of these
// energies is a discrete decision: Code/noise_estimation.py:44-47)
__global__ void __launch_bounds__(128) frame_logenergy_kernel(const real* __restrict__ P, int n_frames, int nb,
                                                              int nbp, real eps, double* __restrict__ energy) {
    CSE_DYN_SMEM(smem_raw);
    double* scratch = reinterpret_cast<double*>(smem_raw);
    const int t = blockIdx.x, u = blockIdx.y;
    const real* row = P + ((size_t)u * n_frames + t) * nbp;
    double acc = 0.0;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) acc += log((double)r_max(row[b], eps));
    acc = block_sum<double>(acc, scratch);
    if (threadIdx.x == 0) energy[(size_t)u * n_frames + t] = acc / (double)nb;
}

// quiet[u][r] = index of the frame with the r-th lowest energy, r < k (ties: lower index first)
__global__ void __launch_bounds__(256) quiet_select_kernel(const double* __restrict__ energy, int n_frames, int k,
                                                           int* __restrict__ quiet) {
    const int u = blockIdx.x;
    const double* e = energy + (size_t)u * n_frames;
    for (int t = threadIdx.x; t < n_frames; t += blockDim.x) {
        const double et = e[t];
        int rank = 0;
        for (int s = 0; s < n_frames; ++s) {
            const double es = e[s];
            rank += (es < et) || (es == et && s < t);
        }
        if (rank < k) quiet[(size_t)u * k + rank] = t;
    }
}

// grid (ceil(nb / BPC), U); block = 32 * BPC; smem = BPC * n_pad reals
__global__ void __launch_bounds__(256) percentile_kernel(const real* __restrict__ P, int n_frames, int nb, int nbp,
                                                         int n_pad, const int* __restrict__ quiet, int k, int k_pad,
                                                         int lo_idx, real frac, real floor_rel, real eps,
                                                         real* __restrict__ N) {
    CSE_DYN_SMEM(smem_raw);
    real* s = reinterpret_cast<real*>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int u = blockIdx.y, b = blockIdx.x * (blockDim.x >> 5) + warp;
    if (b >= nb) return;                       // whole warp exits together; only __syncwarp below
    real* a = s + (size_t)warp * n_pad;
    const real* Pu = P + (size_t)u * n_frames * nbp + b;
    // percentile over the quiet frames
    for (int i = lane; i < k_pad; i += 32) a[i] = i < k ? Pu[(size_t)quiet[(size_t)u * k + i] * nbp] : cse_inf();
    __syncwarp();
    real pct;
    {
        real x0, x1;
        sorted_pair(a, k, k_pad, lane, lo_idx, min(lo_idx + 1, k - 1), x0, x1);
        const real d = x1 - x0;
        pct = frac >= R(0.5) ? x1 - d * (R(1) - frac) : x0 + d * frac;   // numpy _lerp
    }
    __syncwarp();
    // median over all frames
    for (int i = lane; i < n_pad; i += 32) a[i] = i < n_frames ? Pu[(size_t)i * nbp] : cse_inf();
    __syncwarp();
    real m0, m1;
    sorted_pair(a, n_frames, n_pad, lane, (n_frames - 1) >> 1, n_frames >> 1, m0, m1);
    const real med = median_of(m0, m1);
    if (lane == 0) N[(size_t)u * nbp + b] = r_max(r_max(pct, floor_rel * med), eps);
}

// grid (ceil(nb / BPC), U); block = 32 * BPC; smem = 2 * BPC * n_pad reals
__global__ void __launch_bounds__(256) mintrack_kernel(const real* __restrict__ P, int n_frames, int nb, int nbp,
                                                       int n_pad, real a_smooth, int half, real floor_rel, real eps,
                                                       real* __restrict__ N) {
    CSE_DYN_SMEM(smem_raw);
    real* s = reinterpret_cast<real*>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int u = blockIdx.y, b = blockIdx.x * (blockDim.x >> 5) + warp;
    if (b >= nb) return;
    real* a = s + (size_t)warp * 2 * n_pad;    // power series, later sorted
    real* sm = a + n_pad;                      // smoothed series
    const real* Pu = P + (size_t)u * n_frames * nbp + b;
    for (int i = lane; i < n_pad; i += 32) a[i] = i < n_frames ? Pu[(size_t)i * nbp] : cse_inf();
    __syncwarp();
    if (lane == 0) {                           // S_0 = P_0; S_t = a S_{t-1} + (1-a) P_t
        real acc = a[0];
        sm[0] = acc;
        const real oma = R(1) - a_smooth;
        for (int t = 1; t < n_frames; ++t) { acc = a_smooth * acc + oma * a[t]; sm[t] = acc; }
    }
    __syncwarp();
    real m0, m1;
    sorted_pair(a, n_frames, n_pad, lane, (n_frames - 1) >> 1, n_frames >> 1, m0, m1);
    const real fl = r_max(floor_rel * median_of(m0, m1), eps);
    real* Nu = N + (size_t)u * n_frames * nbp + b;
    for (int t = lane; t < n_frames; t += 32) {
        const int t0 = max(t - half, 0), t1 = min(t + half, n_frames - 1);
        real m = sm[t0];
        for (int i = t0 + 1; i <= t1; ++i) m = r_min(m, sm[i]);
        Nu[(size_t)t * nbp] = r_max(m, fl);
    }
}
