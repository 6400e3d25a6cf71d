// Special-function combinations of the MMSE-STSA and Log-MMSE gain rules.
// Chebyshev coefficients come from tools/fit_special.py (scipy.special as ground truth);
// host+device so the CPU tests can sweep them against scipy without a GPU.
#pragma once
#include "cse_common.cuh"
#include "cse_special_coeffs.h"

#ifdef CSE_FP64
#define CSE_COEF(name) CSE_##name##_F64_COEFFS
#define CSE_COEF_N(name) CSE_##name##_F64_N
#else
#define CSE_COEF(name) CSE_##name##_F32_COEFFS
#define CSE_COEF_N(name) CSE_##name##_F32_N
#endif

template <int N> CSE_HD real clenshaw(const real (&c)[N], real t) {
    real b1 = R(0), b2 = R(0);
    const real t2 = t + t;
#pragma unroll
    for (int k = N - 1; k >= 1; --k) {
        const real b0 = r_fma(t2, b1, c[k] - b2);
        b2 = b1;
        b1 = b0;
    }
    return r_fma(t, b1, c[0] - b2);
}

// M(v) = exp(-v/2) * [(1+v) I0(v/2) + v I1(v/2)],  0 <= v <= 80   (Code/mmse.py:92-96)
CSE_HD real cse_mmse_bessel_term(real v) {
    if (v <= R(16)) {
        const real c[CSE_COEF_N(M_LO)] = CSE_COEF(M_LO);
        return clenshaw(c, v * R(0.125) - R(1));
    }
    const real c[CSE_COEF_N(M_HI)] = CSE_COEF(M_HI);
    return clenshaw(c, R(40) / v - R(1.5)) * r_sqrt(v);
}

// E1(v) = scipy.special.expn(1, v),  1e-12 <= v <= 80   (Code/advanced_mmse.py:103)
CSE_HD real cse_expint_e1(real v) {
    if (v <= R(1)) {
        const real c[CSE_COEF_N(E_LO)] = CSE_COEF(E_LO);
        return clenshaw(c, v + v - R(1)) - r_log(v);
    }
    const real c[CSE_COEF_N(E_HI)] = CSE_COEF(E_HI);
    const real t = (R(2) / v - R(1.0125)) * R(1.0 / 0.9875);
    return clenshaw(c, t) * r_exp(-v) / v;
}

// Device-side fast variants used by the gain kernels.  fp32: power-basis Horner (one FFMA per
// term, coefficients become FFMA immediates) + single-MUFU rcp / lg2 / ex2; fp64: the exact forms.
template <int N> CSE_HD real horner(const real (&a)[N], real t) {
    real r = a[N - 1];
#pragma unroll
    for (int k = N - 2; k >= 0; --k) r = r_fma(r, t, a[k]);
    return r;
}
#ifdef CSE_FP64
#define CSE_POLY_EVAL(name, t) clenshaw(c_##name, t)
#define CSE_POLY_DECL(name) const real c_##name[CSE_COEF_N(name)] = CSE_COEF(name)
#else
#define CSE_POLY_EVAL(name, t) horner(c_##name, t)
#define CSE_POLY_DECL(name) const real c_##name[CSE_##name##_F32_N] = CSE_##name##_F32_POLY
#endif

// ---- two-lane forms (both range branches evaluated for both lanes, selected per lane) ----
template <int N> CSE_D real2 horner2(const real (&a)[N], real2 t) {
    real2 r = p_set(a[N - 1]);
#pragma unroll
    for (int k = N - 2; k >= 0; --k) r = p_fma(r, t, p_set(a[k]));
    return r;
}
#ifdef CSE_FP64
template <int N> CSE_D real2 poly2(const real (&c)[N], real2 t) { return mk2(clenshaw(c, t.x), clenshaw(c, t.y)); }
#else
template <int N> CSE_D real2 poly2(const real (&a)[N], real2 t) { return horner2(a, t); }
#endif
CSE_D real2 cse_mmse_bessel_term2(real2 v) {
    CSE_POLY_DECL(M_LO);
    CSE_POLY_DECL(M_HI);
    const real2 lo = poly2(c_M_LO, p_fma(v, p_set(R(0.125)), p_set(R(-1))));
    const real2 hi = p_mul(poly2(c_M_HI, p_fma(p_rcp(v), p_set(R(40)), p_set(R(-1.5)))), p_sqrt(v));   // discarded below 16 (v >= eps > 0)
    return mk2(v.x <= R(16) ? lo.x : hi.x, v.y <= R(16) ? lo.y : hi.y);
}
// enegv = exp(-v) (shared with the speech-presence probability of the caller)
CSE_D real2 cse_half_e1_log2_2(real2 v, real2 enegv) {
    CSE_POLY_DECL(E_LO);
    CSE_POLY_DECL(E_HI);
    // both branches are evaluated on v itself: outside its range a branch only produces a value (possibly inf / NaN)
    // that the per-lane select below discards, so no clamping is needed (v is already clipped to [1e-12, v_max])
    const real2 lo = p_mul(p_set(R(0.5)), p_sub(p_mul(p_set(CSE_LOG2E), poly2(c_E_LO, p_fma(v, p_set(R(2)), p_set(R(-1))))), p_log2(v)));
    const real2 rv = p_rcp(v);
    const real2 t = p_mul(p_fma(rv, p_set(R(2)), p_set(R(-1.0125))), p_set(R(1.0 / 0.9875)));
    const real2 hi = p_mul(p_mul(p_set(R(0.5) * CSE_LOG2E), poly2(c_E_HI, t)), p_mul(enegv, rv));
    return mk2(v.x <= R(1) ? lo.x : hi.x, v.y <= R(1) ? lo.y : hi.y);
}
