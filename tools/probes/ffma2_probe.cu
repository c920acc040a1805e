// Issue-rate probe: FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float ffma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
template <int MODE>
__global__ void k(float* out, int iters, float s) {
    float a[8]; u64 p[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; p[i] = ((u64)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 0.5f); }
    u64 ps = ((u64)__float_as_uint(s) << 32) | __float_as_uint(s);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) a[i] = ffma1(a[i], s, s);
                else if (MODE == 1) p[i] = ffma2(p[i], ps, ps);
                else { if (i & 1) a[i] = ffma1(a[i], s, s); else p[i] = ffma2(p[i], ps, ps); }
            }
        }
    }
    float r = 0; for (int i = 0; i < 8; ++i) r += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> float run(float* d, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(d, 10, 0.999f); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(d, iters, 0.999f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    const int iters = 20000;
    const double inst = 148.0 * 8 * 256 * iters * 64;
    float m0 = run<0>(d, iters), m1 = run<1>(d, iters), m2 = run<2>(d, iters);
    printf("FFMA : %.2f ms  %.2f Tinst/s (thread)  %.1f TFLOP/s\n", m0, inst / m0 / 1e9, 2 * inst / m0 / 1e9);
    printf("FFMA2: %.2f ms  %.2f Tinst/s (thread)  %.1f TFLOP/s\n", m1, inst / m1 / 1e9, 4 * inst / m1 / 1e9);
    printf("mixed: %.2f ms  %.2f Tinst/s (thread)\n", m2, inst / m2 / 1e9);
    return 0;
}
