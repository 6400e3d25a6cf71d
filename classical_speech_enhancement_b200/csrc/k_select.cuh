// K6: the three-way best-candidate selection of optimize_parameters
// (Code/speech_enhancement_comparison.py:186-216) on the device.
//
// The reference walks the grid in itertools.product order and replaces a running best only when a
// candidate beats it by more than a tolerance (1e-6 STOI, 1e-3 PESQ, 1e-5 balance): the result
// depends on the order, so it is NOT an argmax.  One warp per (utterance, criterion) performs that
// very scan, 32 grid points per step: every lane tests its point against the running best; the
// FIRST lane that passes becomes the new best, then only the lanes after it are re-tested against
// the raised threshold, and so on until no lane passes (the loop runs once per accepted update,
// and running bests rise rarely).  All comparisons are in double on the values the host scan
// would see (the table's STOI widened to double; the balance score formed in double as
// 0.5*stoi + 0.5*(max(0, pesq)/4.5), Code/evaluation_metrics.py:104-114), so the winners equal
// those of grid.select_best_batch bit for bit.
#pragma once
#include "cse_common.cuh"

#define CSE_SEL_CRITERIA 3    // 0 stoi, 1 pesq, 2 balance (the order of grid.TOL)

__global__ void __launch_bounds__(32 * CSE_SEL_CRITERIA) select_best_kernel(
    const cse_score_t* __restrict__ table, const double* __restrict__ pesq, int n_points, cse_winner_t* __restrict__ out) {
    const int u = blockIdx.x, crit = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const cse_score_t* __restrict__ row = table + (size_t)u * n_points;
    const double* __restrict__ prow = pesq ? pesq + (size_t)u * n_points : nullptr;
    const double tol = crit == 0 ? 1e-6 : (crit == 1 ? 1e-3 : 1e-5);
    double best = -1.0;
    int bidx = -1;
    for (int t0 = 0; t0 < n_points; t0 += 32) {
        const int i = t0 + lane;
        bool ok = false;
        double v = 0.0;
        if (i < n_points) {
            const cse_score_t r = row[i];
            const double pq = prow ? prow[i] : 0.0;
            ok = (r.flags & CSE_FLAG_VALID) && !(pq != pq);      // NaN PESQ: calculate_pesq returned None -> skipped (:180-181)
            const double st = (double)r.stoi;
            v = crit == 0 ? st : (crit == 1 ? pq : 0.5 * st + 0.5 * ((pq > 0.0 ? pq : 0.0) / 4.5));
        }
        unsigned pending = __ballot_sync(0xffffffffu, ok && v > best + tol);
        while (pending) {
            const int l = __ffs((int)pending) - 1;
            best = __shfl_sync(0xffffffffu, v, l);
            bidx = t0 + l;
            pending = __ballot_sync(0xffffffffu, ok && lane > l && v > best + tol);
        }
    }
    if (lane == 0) {
        cse_winner_t w;
        w.index = bidx; w.lag = 0; w.flags = 0; w.reserved = 0;
        w.score = best; w.stoi = 0.0; w.pesq = 0.0; w.snr = 0.0;
        if (bidx >= 0) {
            const cse_score_t r = row[bidx];
            w.lag = r.lag; w.flags = r.flags;
            w.stoi = (double)r.stoi;
            w.pesq = prow ? prow[bidx] : 0.0;
            w.snr = (r.flags & CSE_FLAG_SNR_INF) ? (double)INFINITY : (double)r.snr;
        }
        out[(size_t)u * CSE_SEL_CRITERIA + crit] = w;
    }
}
