// Micro-benchmark: scattered global reductions, scalar vs vector (red.global.add.v4.f32 / v2.f32).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_bench red_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void red_v4(float *p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_v2(float *p, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ uint32_t hash(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// mode 0: 12 scalar reds (3 planes x 4 taps), 1: 4 x v4 (interleaved texel), 2: 3 planes x 2 rows x v2, 3: 4 scalar
template <int MODE>
__global__ void k(float *buf, int texels, int iters, int spread) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        // neighbouring lanes hit neighbouring texels (like pixels of one face), warps are spread
        const uint32_t base = hash((tid >> 5) * 977u + it * 131u) % (uint32_t)(texels - 4096);
        const int t = base + (threadIdx.x & 31) * spread;
        const int W = 200;
        if (MODE == 0) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float *p = buf + (size_t)c * texels + t;
                atomicAdd(p, 1.f); atomicAdd(p + 1, 1.f); atomicAdd(p + W, 1.f); atomicAdd(p + W + 1, 1.f);
            }
        } else if (MODE == 1) {
            float *p = buf + (size_t)t * 4;
            red_v4(p, 1.f, 1.f, 1.f, 0.f); red_v4(p + 4, 1.f, 1.f, 1.f, 0.f);
            red_v4(p + 4 * W, 1.f, 1.f, 1.f, 0.f); red_v4(p + 4 * W + 4, 1.f, 1.f, 1.f, 0.f);
        } else if (MODE == 2) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float *p = buf + (size_t)c * texels + (t & ~1);
                red_v2(p, 1.f, 1.f); red_v2(p + W, 1.f, 1.f);
            }
        } else {
            float *p = buf + t;
            atomicAdd(p, 1.f); atomicAdd(p + 1, 1.f); atomicAdd(p + W, 1.f); atomicAdd(p + W + 1, 1.f);
        }
    }
}

template <int MODE>
float run(float *buf, int texels, int iters, int spread) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(buf, texels, 4, spread);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(buf, texels, iters, spread);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    const int texels = 64 * 40000;      // cfg2: 64 views x 200 x 200
    float *buf; cudaMalloc(&buf, (size_t)texels * 4 * sizeof(float) + (1 << 20));
    cudaMemset(buf, 0, (size_t)texels * 4 * sizeof(float));
    const int iters = 64;
    const double lanes = 148.0 * 8 * 256 * iters;
    for (int spread = 1; spread <= 2; ++spread) {
        float m0 = run<0>(buf, texels, iters, spread), m1 = run<1>(buf, texels, iters, spread);
        float m2 = run<2>(buf, texels, iters, spread), m3 = run<3>(buf, texels, iters, spread);
        printf("spread %d: 12 scalar %.3f ms (%.2f cyc/lane-op/SM) | 4 x v4 %.3f ms (%.2f cyc/lane-op/SM) | 6 x v2 %.3f ms (%.2f) | 4 scalar %.3f ms (%.2f)\n",
               spread, m0, m0 * 1e-3 * 1.965e9 * 148 / (lanes * 12), m1, m1 * 1e-3 * 1.965e9 * 148 / (lanes * 4), m2,
               m2 * 1e-3 * 1.965e9 * 148 / (lanes * 6), m3, m3 * 1e-3 * 1.965e9 * 148 / (lanes * 4));
    }
    cudaError_t e = cudaGetLastError();
    printf("status %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
