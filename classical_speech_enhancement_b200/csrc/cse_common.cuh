// Common definitions for the cse kernels (sm_100a; also built by g++ under CSE_EMU for tests).
#pragma once
#ifdef CSE_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#define CSE_LAUNCH(kern, grid, block, smem, stream, ...) \
    kern<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
#define CSE_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

#include "../../include/cse.h"

#ifdef CSE_FP64
typedef double real;
typedef double2 real2;
#define CSE_REAL_BITS 64
#else
typedef float real;
typedef float2 real2;
#define CSE_REAL_BITS 32
#endif
#define R(x) ((real)(x))
#define CSE_HD __host__ __device__ __forceinline__
#define CSE_D __device__ __forceinline__

CSE_HD real2 mk2(real a, real b) { real2 r; r.x = a; r.y = b; return r; }
#if !defined(CSE_FP64) && !defined(CSE_EMU) && defined(__CUDA_ARCH__)
// Blackwell packed FP32: one FADD2 / FMUL2 / FFMA2 instruction works on both halves of a 64-bit
// register pair (operand broadcast, swap and per-half negation are free modifiers), so a complex
// add is ONE instruction and a complex multiply TWO.  tools/micro/ffma2_bench.cu: 65.8 TFLOP/s
// with FFMA2 against 42 TFLOP/s with scalar FFMA on this B200; in these issue-bound kernels the
// halved instruction count is what matters.
#define CSE_PACKED_F32 1
CSE_D real2 cadd(real2 a, real2 b) { return __fadd2_rn(a, b); }
CSE_D real2 csub(real2 a, real2 b) { return __fadd2_rn(a, mk2(-b.x, -b.y)); }
// (the scalar halves of `a` are written as plain broadcasts: ptxas then uses the .F32 operand form and needs no register moves;
// the swap and the per-half sign go on `b`, where they are operand modifiers)
CSE_D real2 cmul(real2 a, real2 b) { return __ffma2_rn(mk2(a.y, a.y), mk2(-b.y, b.x), __fmul2_rn(mk2(a.x, a.x), b)); }
CSE_D real2 cmulc(real2 a, real2 b) { /* a * conj(b) */ return __ffma2_rn(mk2(a.y, a.y), mk2(b.y, b.x), __fmul2_rn(mk2(a.x, a.x), mk2(b.x, -b.y))); }
CSE_D real2 cscale(real2 a, real s) { return __fmul2_rn(a, mk2(s, s)); }
CSE_D real2 cfma2(real2 a, real2 b, real2 c) { return __ffma2_rn(a, b, c); }      // elementwise a*b + c
#else
CSE_HD real2 cadd(real2 a, real2 b) { return mk2(a.x + b.x, a.y + b.y); }
CSE_HD real2 csub(real2 a, real2 b) { return mk2(a.x - b.x, a.y - b.y); }
CSE_HD real2 cmul(real2 a, real2 b) { return mk2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
CSE_HD real2 cmulc(real2 a, real2 b) { /* a * conj(b) */ return mk2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
CSE_HD real2 cscale(real2 a, real s) { return mk2(a.x * s, a.y * s); }
CSE_HD real2 cfma2(real2 a, real2 b, real2 c) { return mk2(a.x * b.x + c.x, a.y * b.y + c.y); }
#endif
CSE_HD real2 cconj(real2 a) { return mk2(a.x, -a.y); }

// precision-generic math (float intrinsics only where their error is far below the 1e-4 budget)
#ifdef CSE_FP64
CSE_HD real r_sqrt(real x) { return sqrt(x); }
CSE_HD real r_exp(real x) { return exp(x); }
CSE_HD real r_log(real x) { return log(x); }
CSE_HD real r_log10(real x) { return log10(x); }
CSE_HD real r_pow(real x, real y) { return pow(x, y); }
CSE_HD real r_abs(real x) { return fabs(x); }
CSE_HD real r_max(real a, real b) { return fmax(a, b); }
CSE_HD real r_min(real a, real b) { return fmin(a, b); }
CSE_HD real r_floor(real x) { return floor(x); }
CSE_HD real r_fma(real a, real b, real c) { return fma(a, b, c); }
#else
CSE_HD real r_sqrt(real x) { return sqrtf(x); }
CSE_HD real r_exp(real x) { return expf(x); }
CSE_HD real r_log(real x) { return logf(x); }
CSE_HD real r_log10(real x) { return log10f(x); }
CSE_HD real r_pow(real x, real y) { return powf(x, y); }
CSE_HD real r_abs(real x) { return fabsf(x); }
CSE_HD real r_max(real a, real b) { return fmaxf(a, b); }
CSE_HD real r_min(real a, real b) { return fminf(a, b); }
CSE_HD real r_floor(real x) { return floorf(x); }
CSE_HD real r_fma(real a, real b, real c) { return fmaf(a, b, c); }
#endif
// Fast forms for the fp32 gain rules: ONE MUFU instruction each (rcp / lg2 / ex2 / rsq .approx.ftz,
// relative error <= ~2e-7 .. 1e-6 over the ranges the gain rules use - far inside the 1e-4
// waveform budget; no argument here is ever denormal).  The FP64 build and the CPU emulation use
// the exact forms.
// Hint: pull the 128-byte line at p into L2 (no register, no dependency); nothing under the emulator.
#if defined(CSE_EMU)
CSE_D void cse_prefetch_l2(const void*) {}
#else
CSE_D void cse_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif
#if defined(CSE_FP64) || defined(CSE_EMU)
CSE_HD real r_rcp(real x) { return R(1) / x; }
CSE_HD real r_fexp2(real x) { return (real)exp2((double)x); }
CSE_HD real r_flog2(real x) { return (real)log2((double)x); }
CSE_HD real r_fsqrt(real x) { return r_sqrt(x); }
#else
CSE_D real r_rcp(real x) { real y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
CSE_D real r_fexp2(real x) { real y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
CSE_D real r_flog2(real x) { real y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
CSE_D real r_fsqrt(real x) { real y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return x * y; }
#endif
// Shared-memory accesses through 32-bit shared-window addresses (ld.shared / st.shared with a register
// address): hot loops that hold a few base addresses in registers and index them with packed byte offsets
// get LDS / STS with no generic-pointer arithmetic around them.  Under the CPU emulation an "address" is the
// byte offset from the CTA's dynamic shared memory.
#if defined(CSE_EMU)
CSE_D unsigned cse_saddr(const void* p) { return (unsigned)((const unsigned char*)p - cse_emu::dyn_smem()); }
CSE_D real2 cse_lds_r2(unsigned a) { return *reinterpret_cast<const real2*>(cse_emu::dyn_smem() + a); }
CSE_D uint2 cse_lds_u2(unsigned a) { return *reinterpret_cast<const uint2*>(cse_emu::dyn_smem() + a); }
CSE_D void cse_sts_r2(unsigned a, real2 v) { *reinterpret_cast<real2*>(cse_emu::dyn_smem() + a) = v; }
#else
CSE_D unsigned cse_saddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
CSE_D uint2 cse_lds_u2(unsigned a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
#ifdef CSE_FP64
CSE_D real2 cse_lds_r2(unsigned a) { real2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a)); return v; }
CSE_D void cse_sts_r2(unsigned a, real2 v) { asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory"); }
#else
CSE_D real2 cse_lds_r2(unsigned a) { real2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v; }
CSE_D void cse_sts_r2(unsigned a, real2 v) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory"); }
#endif
#endif
// Makes a register value opaque to the optimiser (it can then neither be re-derived nor folded): used where the
// compiler would rather recompute loop-invariant index data inside a hot loop than keep it in a register.
#if defined(CSE_EMU)
#define CSE_OPAQUE(x) ((void)0)
#define CSE_OPAQUE_PTR(x) ((void)0)
#else
#define CSE_OPAQUE(x) asm volatile("" : "+r"(x))
#define CSE_OPAQUE_PTR(x) asm volatile("" : "+l"(x))
#endif
// Two-lane ("packed pair") helpers: the same scalar computation for two independent values held
// in the halves of a real2.  Arithmetic maps to FADD2 / FMUL2 / FFMA2 on sm_100a; min/max and the
// MUFU functions have no packed form and are applied per half.
CSE_D real2 p_set(real s) { return mk2(s, s); }
CSE_D real2 p_add(real2 a, real2 b) { return cadd(a, b); }
CSE_D real2 p_sub(real2 a, real2 b) { return csub(a, b); }
CSE_D real2 p_fma(real2 a, real2 b, real2 c) { return cfma2(a, b, c); }
#ifdef CSE_PACKED_F32
CSE_D real2 p_mul(real2 a, real2 b) { return __fmul2_rn(a, b); }
#else
CSE_D real2 p_mul(real2 a, real2 b) { return mk2(a.x * b.x, a.y * b.y); }
#endif
CSE_D real2 p_max(real2 a, real2 b) { return mk2(r_max(a.x, b.x), r_max(a.y, b.y)); }
CSE_D real2 p_min(real2 a, real2 b) { return mk2(r_min(a.x, b.x), r_min(a.y, b.y)); }
CSE_D real2 p_clip(real2 a, real lo, real hi) { return mk2(r_min(r_max(a.x, lo), hi), r_min(r_max(a.y, lo), hi)); }
CSE_D real2 p_rcp(real2 a) { return mk2(r_rcp(a.x), r_rcp(a.y)); }
CSE_D real2 p_exp2(real2 a) { return mk2(r_fexp2(a.x), r_fexp2(a.y)); }
CSE_D real2 p_log2(real2 a) { return mk2(r_flog2(a.x), r_flog2(a.y)); }
CSE_D real2 p_sqrt(real2 a) { return mk2(r_fsqrt(a.x), r_fsqrt(a.y)); }
#define CSE_LOG2E R(1.44269504088896340736)
#define CSE_LN2 R(0.69314718055994530942)
// numpy's maximum/minimum/clip propagate NaN, fmax/fmin drop it.  The reference relies on
// that only through np.nan_to_num, which the gain kernels restate explicitly.
CSE_HD real r_clip(real x, real lo, real hi) { return r_min(r_max(x, lo), hi); }
CSE_HD bool r_finite(real x) { return (x - x) == R(0); }
CSE_HD real cse_inf() { return (real)INFINITY; }

// ---------------------------------------------------------------- geometry helpers
static inline int cse_ilog2(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }
CSE_HD int cse_nbp(int n_fft) { return ((n_fft / 2 + 1) + 7) & ~7; }   // bins padded to 8

// ---------------------------------------------------------------- constant tables
// One blob in device memory, filled by cse_tables_init.  Offsets in units of `real`.
#define CSE_TW_N 8192                 // twiddle table: W_8192^k = exp(-2*pi*i*k/8192), k < 4096
#define CSE_MAX_NFFT 2048
#define CSE_RS_TAPS 581               // STOI 16k->10k resampler (pystoi resample_oct), 5 phases
#define CSE_RS_ROWS 136               // tile rows jj = (j - 8a) + 64 that can reach the 5 outputs 5a+p
struct CseTables {
    real2 tw[CSE_TW_N / 2];
    real hann256[256], hann512[512], hann1024[1024], hann2048[2048];   // periodic Hann (librosa)
    real stoi_win[256];                                                 // np.hanning(258)[1:-1]
    real rs[CSE_RS_ROWS][8];          // rs[jj][p] = h5[8p + 610 - 5jj] (0 outside the filter), p < 5
    double rs_d[CSE_RS_ROWS][8];      // same in double for the clean-side VAD decision
    int stoi_edges[16];               // third-octave band edges as FFT-bin indices [lo, hi)
};
CSE_D const real* cse_hann(const CseTables* T, int n_fft) {
    return n_fft == 256 ? T->hann256 : n_fft == 512 ? T->hann512 : n_fft == 1024 ? T->hann1024 : T->hann2048;
}

// ---------------------------------------------------------------- block reductions
CSE_D real warp_sum(real v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
CSE_D double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Sum over the block; result valid in every thread.  `scratch` >= 33 elements of shared memory.
// blockDim.x must be a multiple of 32.
template <class T> CSE_D T block_sum(T v, T* scratch) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        T s = lane < nw ? scratch[lane] : (T)0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) scratch[32] = s;
    }
    __syncthreads();
    return scratch[32];
}
