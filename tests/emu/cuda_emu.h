// CPU thread-emulation of the small CUDA subset the cse kernels use.
//
// TEST INFRASTRUCTURE ONLY.  This header lets `g++ -DCSE_EMU -x c++ cse_lib.cu` build the
// very same kernel sources into tests/emu/libcse_emu.so so that indexing, synchronisation
// and fp32 error budgets can be checked in a container that has no GPU.  One OS thread per
// CUDA thread, pthread barriers for __syncthreads / warp collectives, blocks run a few at a
// time.  The product package never loads this library (it loads libcse_sm100a.so and fails
// loudly when that is missing); only tests/ does, and only under `-m "not gpu"`.
#pragma once
#ifndef CSE_EMU
#error "cuda_emu.h is only for the CSE_EMU host build"
#endif
#include <pthread.h>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __restrict__ __restrict
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };
struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(16))) double2 { double x, y; };
struct int2 { int x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
struct uint2 { unsigned x, y; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }

namespace cse_emu {
struct BlockCtx {
    pthread_barrier_t block_bar;
    std::vector<pthread_barrier_t> warp_bar;
    std::vector<uint64_t> xch;      // 32 slots per warp
    unsigned char* smem;
    unsigned nthreads;
    std::atomic<int> orflag[2];     // __syncthreads_or, alternating by call parity
    std::mutex named_mu;            // named barriers (bar.sync id, count), created on first use
    std::map<int, struct NamedBarrier*> named;
};
struct Tls { BlockCtx* blk; unsigned lane, warp, orphase; };
extern thread_local Tls tls;
}  // namespace cse_emu
extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;

#ifdef CSE_EMU_IMPL
namespace cse_emu { thread_local Tls tls; }
thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;
#endif

static inline void __syncthreads() { pthread_barrier_wait(&cse_emu::tls.blk->block_bar); }
static inline int __syncthreads_or(int pred) {
    cse_emu::Tls& t = cse_emu::tls;
    std::atomic<int>& f = t.blk->orflag[t.orphase++ & 1];
    if (pred) f.store(1);
    pthread_barrier_wait(&t.blk->block_bar);
    const int r = f.load();
    pthread_barrier_wait(&t.blk->block_bar);
    if (t.lane == 0 && t.warp == 0) f.store(0);     // next use of this flag is two calls (>= two barriers) away
    return r;
}
namespace cse_emu {
// Named barriers (bar.sync id, count / bar.arrive id, count): `count` arrivals complete a phase; bar.sync arrives and
// waits for the phase to complete, bar.arrive only arrives.  A thread that runs ahead into the next phase of the same
// barrier waits until the previous phase has completed (the hardware's behaviour for a barrier still in use).
struct NamedBarrier {
    std::mutex mu;
    std::condition_variable cv;
    unsigned arrived = 0;
    unsigned long generation = 0;
};
static inline NamedBarrier* named_barrier_get(int id) {
    BlockCtx* b = tls.blk;
    std::lock_guard<std::mutex> g(b->named_mu);
    auto it = b->named.find(id);
    if (it == b->named.end()) it = b->named.emplace(id, new NamedBarrier).first;
    return it->second;
}
static inline void named_barrier_op(int id, unsigned count, bool wait) {
    NamedBarrier* nb = named_barrier_get(id);
    std::unique_lock<std::mutex> lk(nb->mu);
    const unsigned long gen = nb->generation;
    if (++nb->arrived == count) {
        nb->arrived = 0;
        ++nb->generation;
        nb->cv.notify_all();
    } else if (wait) {
        nb->cv.wait(lk, [&] { return nb->generation != gen; });
    }
}
static inline void named_barrier(int id, unsigned count) { named_barrier_op(id, count, true); }
static inline void named_barrier_arrive(int id, unsigned count) { named_barrier_op(id, count, false); }
}  // namespace cse_emu
static inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&cse_emu::tls.blk->warp_bar[cse_emu::tls.warp]); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }

namespace cse_emu {
template <class T> static inline T exchange(T v, unsigned src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    Tls& t = tls;
    uint64_t* slots = &t.blk->xch[t.warp * 32];
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    slots[t.lane] = raw;
    pthread_barrier_wait(&t.blk->warp_bar[t.warp]);
    unsigned nl = std::min(32u, t.blk->nthreads - t.warp * 32);
    T out = v;
    if (src_lane < nl) memcpy(&out, &slots[src_lane], sizeof(T));
    pthread_barrier_wait(&t.blk->warp_bar[t.warp]);
    return out;
}
}  // namespace cse_emu
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    unsigned lane = cse_emu::tls.lane;
    return cse_emu::exchange(v, (lane / width) * width + (unsigned)(src % width));
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return cse_emu::exchange(v, cse_emu::tls.lane ^ (unsigned)m); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
    unsigned s = cse_emu::tls.lane + d;
    return cse_emu::exchange(v, s < 32 ? s : cse_emu::tls.lane);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
    unsigned l = cse_emu::tls.lane;
    return cse_emu::exchange(v, l >= d ? l - d : l);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned r = 0;
    for (unsigned l = 0; l < 32; ++l) { int p = __shfl_sync(0xffffffffu, pred, (int)l); if (p && l < std::min(32u, cse_emu::tls.blk->nthreads - cse_emu::tls.warp * 32)) r |= 1u << l; }
    return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float atomicAdd(float* p, float v) {
    uint32_t* ip = (uint32_t*)p; uint32_t old = __atomic_load_n(ip, __ATOMIC_RELAXED), nw; float f;
    do { memcpy(&f, &old, 4); f += v; memcpy(&nw, &f, 4); } while (!__atomic_compare_exchange_n(ip, &old, nw, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED));
    memcpy(&f, &old, 4); return f;
}
static inline double atomicAdd(double* p, double v) {
    uint64_t* ip = (uint64_t*)p; uint64_t old = __atomic_load_n(ip, __ATOMIC_RELAXED), nw; double f;
    do { memcpy(&f, &old, 8); f += v; memcpy(&nw, &f, 8); } while (!__atomic_compare_exchange_n(ip, &old, nw, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED));
    memcpy(&f, &old, 8); return f;
}
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicMax(int* p, int v) { int o = __atomic_load_n(p, __ATOMIC_RELAXED); while (o < v && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED)) {} return o; }
static inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }

static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline void sincospif(float x, float* s, float* c) { *s = (float)sin(M_PI * (double)x); *c = (float)cos(M_PI * (double)x); }
static inline void sincospi(double x, double* s, double* c) { *s = sin(M_PI * x); *c = cos(M_PI * x); }
static inline unsigned __brev(unsigned x) { x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1); x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2); x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4); x = ((x >> 8) & 0x00ff00ffu) | ((x & 0x00ff00ffu) << 8); return (x >> 16) | (x << 16); }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
using std::isfinite; using std::isnan; using std::isinf; using std::min; using std::max;

namespace cse_emu {
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
#ifdef CSE_EMU_IMPL
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    const unsigned nthreads = block.x * block.y * block.z;
    const unsigned nwarps = (nthreads + 31) / 32;
    const unsigned long nblocks = (unsigned long)grid.x * grid.y * grid.z;
    unsigned conc = 4;
    if (const char* e = getenv("CSE_EMU_BLOCKS")) conc = (unsigned)std::max(1, atoi(e));
    for (unsigned long b0 = 0; b0 < nblocks; b0 += conc) {
        unsigned long b1 = std::min(nblocks, b0 + conc);
        std::vector<BlockCtx*> ctxs;
        std::vector<std::thread> threads;
        for (unsigned long b = b0; b < b1; ++b) {
            BlockCtx* c = new BlockCtx;
            c->orflag[0].store(0); c->orflag[1].store(0);
            c->nthreads = nthreads;
            pthread_barrier_init(&c->block_bar, nullptr, nthreads);
            c->warp_bar.resize(nwarps);
            for (unsigned w = 0; w < nwarps; ++w) pthread_barrier_init(&c->warp_bar[w], nullptr, std::min(32u, nthreads - w * 32));
            c->xch.assign(nwarps * 32, 0);
            c->smem = (unsigned char*)aligned_alloc(128, ((smem + 127) / 128 + 1) * 128);
            memset(c->smem, 0xCD, smem);      // poison: uninitialised shared memory reads show up
            ctxs.push_back(c);
            uint3 bi{(unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((unsigned long)grid.x * grid.y))};
            for (unsigned t = 0; t < nthreads; ++t) {
                threads.emplace_back([=, &body]() {
                    tls.blk = c; tls.lane = t % 32; tls.warp = t / 32; tls.orphase = 0;
                    threadIdx = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
                    blockIdx = bi; blockDim = block; gridDim = grid;
                    body();
                });
            }
        }
        for (auto& th : threads) th.join();
        for (BlockCtx* c : ctxs) {
            pthread_barrier_destroy(&c->block_bar);
            for (auto& wb : c->warp_bar) pthread_barrier_destroy(&wb);
            for (auto& nb : c->named) delete nb.second;
            free(c->smem);
            delete c;
        }
    }
}
#endif
static inline unsigned char* dyn_smem() { return tls.blk->smem; }
}  // namespace cse_emu

#define CSE_LAUNCH(kern, grid, block, smem, stream, ...) \
    cse_emu::launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { kern(__VA_ARGS__); })
#define CSE_DYN_SMEM(name) unsigned char* name = cse_emu::dyn_smem()
