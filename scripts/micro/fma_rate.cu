// microbenchmark: fp32 FMA issue rate per SM (FFMA vs FFMA2, register vs shared-memory operands) -- decides whether
// the diffusion loops of rnn_fwd/rnn_bwd are FMA-throughput or latency bound.   nvcc -arch=sm_100a -O3 fma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_ffma(float* out, int iters, float a, float b) {
    float acc[20];
    for (int i = 0; i < 20; ++i) acc[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 20; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0; for (int i = 0; i < 20; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, int iters, float a, float b) {
    unsigned long long acc[10];
    for (int i = 0; i < 10; ++i) { float x = threadIdx.x * 0.001f + i; asm("mov.b64 %0, {%1, %1};" : "=l"(acc[i]) : "f"(x)); }
    unsigned long long aa, bb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 10; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(aa), "l"(bb));
    }
    float s = 0; for (int i = 0; i < 10; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(acc[i])); s += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// the diffusion inner loop: per iteration 1 LDS.32 + 5 broadcast LDS.128 + 10 FFMA2
__global__ void k_diff(float* out, int iters) {
    __shared__ __align__(16) float PT[20 * 20];
    __shared__ float Z[32 * 68];
    for (int i = threadIdx.x; i < 400; i += blockDim.x) PT[i] = 0.01f * i;
    for (int i = threadIdx.x; i < 32 * 68; i += blockDim.x) Z[i] = 0.001f * i;
    __syncthreads();
    unsigned long long a64[10];
    for (int i = 0; i < 10; ++i) a64[i] = 0ull;
    const int lane = threadIdx.x & 31;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 2
        for (int j = 0; j < 19; ++j) {
            const float zv = Z[j * 68 + lane];
            unsigned long long zx;
            asm("mov.b64 %0, {%1, %1};" : "=l"(zx) : "f"(zv));
            const ulonglong2* pr = reinterpret_cast<const ulonglong2*>(PT + j * 20);
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const ulonglong2 pv = pr[q];
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a64[2 * q]) : "l"(pv.x), "l"(zx));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a64[2 * q + 1]) : "l"(pv.y), "l"(zx));
            }
        }
    }
    float s = 0; for (int i = 0; i < 10; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a64[i])); s += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    for (int warps : {4, 8, 16, 32}) {
        const int iters = 20000;
        for (int which = 0; which < 3; ++which) {
            float ms = 0;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (which == 0) k_ffma<<<148, warps * 32>>>(out, iters, 1.0001f, 0.5f);
                else if (which == 1) k_ffma2<<<148, warps * 32>>>(out, iters, 1.0001f, 0.5f);
                else k_diff<<<148, warps * 32>>>(out, iters / 19);
                cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            }
            const double fma = (which == 2 ? (double)(iters / 19) * 19 * 20 : (double)iters * 20) * warps * 32;   // per SM
            printf("%s warps/SM=%2d: %.3f ms -> %.1f FMA/ns/SM (at %.2f GHz max: %.1f FMA/clk/SM)\n", which == 0 ? "FFMA " : which == 1 ? "FFMA2" : "diff ",
                   warps, ms, fma / (ms * 1e6), clk * 1e-6, fma / (ms * 1e6) / (clk * 1e-6));
        }
    }
    return 0;
}
