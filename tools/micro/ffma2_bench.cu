// Microbenchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) issue/throughput on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE>
__global__ void k(float2* out, float a, float b, int iters) {
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    const float2 x = make_float2(a, a * 1.0001f), y = make_float2(b, b * 0.9999f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { acc[i].x = fmaf(acc[i].x, x.x, y.x); acc[i].y = fmaf(acc[i].y, x.y, y.y); }
            else acc[i] = __ffma2_rn(acc[i], x, y);
        }
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s.x += acc[i].x; s.y += acc[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float2* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float2));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148 * 8, 256>>>(out, 0.999f, 0.001f, iters); else k<1><<<148 * 8, 256>>>(out, 0.999f, 0.001f, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 16 * iters * 148.0 * 8 * 256;
        printf("%s: %.3f ms  %.1f TFLOP/s\n", mode ? "FFMA2" : "FFMA ", ms, flops / ms / 1e9);
    }
    return 0;
}
