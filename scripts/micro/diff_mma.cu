// microbenchmark + check of the warp-level tensor-path diffusion (f16_common.cuh::diffuse_mma) against the FMA version
// (diffuse2 + store_cols2) and a double-precision host evaluation.  One CTA per SM, 8 worker warps, each warp repeats
// one (sample, term) task: 64 columns x 19 nodes, as in one diffusion phase of rnn_fwd / rnn_bwd at M = 3.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I eeg-gnn-ssl_b200/csrc -I include scripts/micro/diff_mma.cu -o scripts/micro/diff_mma
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "f16_common.cuh"

using namespace dcgru::f16;

constexpr int ZLD = 68;

__global__ void __launch_bounds__(256, 1) k_diff(const float* Z, const float* P, int N, int iters, int which, uint8_t* out, long long* cyc) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* slots = smem;                                          // 2 chunk slots (one per term)
    float* ZH = reinterpret_cast<float*>(smem + 2 * SLOT);          // [128][ZLD]
    float* PTs = ZH + 128 * ZLD;                                    // [s][term][j][NPAD]
    uint8_t* src16 = smem + 2 * SLOT + 128 * ZLD * 4 + ((SB * 2 * PT_STRIDE * 4 + 1023) / 1024) * 1024;   // hi / lo chunk of Z
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 2 * SLOT / 4; i += 256) reinterpret_cast<uint32_t*>(slots)[i] = 0u;
    for (int i = tid; i < 128 * ZLD; i += 256) { const int r = i / ZLD, c = i % ZLD; ZH[i] = (c < 64 && (r & 31) < N) ? Z[r * 64 + c] : 0.f; }
    for (int i = tid; i < SB * 2 * PT_STRIDE; i += 256) PTs[i] = 0.f;
    for (int i = tid; i < 128 * 8; i += 256) {                      // the state as the epilogue leaves it: hi / lo chunk
        const int r = i >> 3, u = i & 7;
        float v[8];
        for (int j = 0; j < 8; ++j) v[j] = ((r & 31) < N) ? Z[r * 64 + 8 * u + j] : 0.f;
        uint4 hi, lo;
        split8(v, hi, lo);
        *reinterpret_cast<uint4*>(src16 + k128_off(r, 8 * u)) = hi;
        *reinterpret_cast<uint4*>(src16 + PLANE + k128_off(r, 8 * u)) = lo;
    }
    __syncthreads();
    load_pt(PTs, P, 0, SB, N, 2, 0, tid, 256);
    __syncthreads();
    const int s = warp & 3, m = warp >> 2;                          // task of this warp
    const float* pt = PTs + (s * 2 + m) * PT_STRIDE;
    uint8_t* sl = slots + m * SLOT;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (which == 0) {
            float acc[NPAD][2];
            diffuse2(ZH + (s * RP) * ZLD + 2 * lane, ZLD, N, pt, acc);
            store_cols2(sl, s * RP, lane, N, acc, 1.f, RG * 8);
        } else if (which == 1) {
            PFrag pf;
            load_pfrag(pt, lane, pf);
            diffuse_mma(ZH + (s * RP) * ZLD, ZLD, N, pf, sl, s * RP, lane, 1.f);
        } else {
            PFrag pf;
            load_pfrag(pt, lane, pf);
            diffuse_mma16(src16, s * RP, pf, sl, s * RP, lane, 1.f);
        }
        __syncwarp();
    }
    long long t1 = clock64();
    __syncthreads();
    if (blockIdx.x == 0) {
        for (int i = tid; i < 2 * SLOT / 4; i += 256) reinterpret_cast<uint32_t*>(out)[i] = reinterpret_cast<uint32_t*>(slots)[i];
        if (tid == 0) cyc[0] = (t1 - t0) / iters;
    }
}

// raw issue rate of the legacy tensor path: 8 independent accumulators per warp, `warps` warps per SM
__global__ void k_hmma(float* out, int iters) {
    float d[8][4];
    uint32_t a[4] = {0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u}, b[2] = {0x3c003c00u, 0x3c003c00u};
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) d[i][j] = (float)threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) mma_f16_16816(d[i], a, b);
    }
    float s = 0.f;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float h2f(unsigned short h) {                                  // fp16 -> fp32 (host)
    const unsigned s = (h >> 15) & 1, e = (h >> 10) & 31, f = h & 1023;
    float v;
    if (e == 0) v = ldexpf((float)f, -24);
    else if (e == 31) v = f ? NAN : INFINITY;
    else v = ldexpf((float)(f | 1024), (int)e - 25);
    return s ? -v : v;
}

int main() {
    const int N = 19;
    std::vector<float> Z(128 * 64), P(SB * 2 * N * N);
    srand(7);
    for (auto& v : Z) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    for (auto& v : P) v = (rand() / (float)RAND_MAX - 0.5f);
    float *dZ, *dP; uint8_t* dout; long long* dcyc;
    cudaMalloc(&dZ, Z.size() * 4); cudaMalloc(&dP, P.size() * 4); cudaMalloc(&dout, 2 * SLOT); cudaMalloc(&dcyc, 8);
    cudaMemcpy(dZ, Z.data(), Z.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dP, P.data(), P.size() * 4, cudaMemcpyHostToDevice);
    const int smem = 2 * SLOT + 128 * ZLD * 4 + ((SB * 2 * PT_STRIDE * 4 + 1023) / 1024) * 1024 + SLOT + 1024;
    cudaFuncSetAttribute(k_diff, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<uint8_t> out(2 * SLOT);
    for (int which = 0; which < 3; ++which) {
        long long cyc = 0;
        k_diff<<<148, 256, smem>>>(dZ, dP, N, 200, which, dout, dcyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(out.data(), dout, 2 * SLOT, cudaMemcpyDeviceToHost);
        cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0, maxpad = 0;
        for (int m = 0; m < 2; ++m)
            for (int s = 0; s < SB; ++s)
                for (int n = 0; n < 32; ++n)
                    for (int c = 0; c < 64; ++c) {
                        const int row = s * 32 + n;
                        const unsigned off = (row >> 3) * 1024 + (row & 7) * 128 + ((((c >> 3) ^ (row & 7)) << 4) | ((c & 7) << 1));
                        const unsigned short hh = *reinterpret_cast<unsigned short*>(&out[m * SLOT + off]);
                        const unsigned short ll = *reinterpret_cast<unsigned short*>(&out[m * SLOT + PLANE + off]);
                        const double got = (double)h2f(hh) + (double)h2f(ll);
                        if (n >= N) { maxpad = fmax(maxpad, fabs(got)); continue; }
                        double ref = 0;
                        for (int j = 0; j < N; ++j) ref += (double)P[((s * 2 + m) * N + n) * N + j] * Z[(s * 32 + j) * 64 + c];
                        maxerr = fmax(maxerr, fabs(got - ref)); maxref = fmax(maxref, fabs(ref));
                    }
        printf("%s: %lld cycles per (sample, term) task with 8 warps/SM; max|err| / max|ref| = %.3e (pad rows max %.1e)\n",
               which == 0 ? "FMA (diffuse2 + store_cols2)  " : which == 1 ? "mma.sync, fp32 source        " : "mma.sync, hi/lo chunk source ", cyc, maxerr / maxref, maxpad);
    }
    float* dump; cudaMalloc(&dump, 148 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    for (int warps : {4, 8, 16}) {
        float ms = 0;
        const int iters = 20000;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            k_hmma<<<148, warps * 32>>>(dump, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        }
        const double n = (double)iters * 8 * warps;                 // HMMAs per SM
        const double cyc = ms * 1e-3 * clk * 1e3;
        printf("HMMA.16816.F32 warps/SM=%2d: %.2f cycles per HMMA per SM sub-partition, %.0f fp16 MAC/clk/SM\n", warps, cyc / (n / 4), n * 2048 / cyc);
    }
    return 0;
}
