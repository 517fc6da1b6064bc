// Graph kernels: supports -> diffusion polynomials, and raw clip -> correlation-graph supports.
#include "common.cuh"
#include "dw.cuh"

namespace dcgru {

// P[b][m-1] = the matrix that maps Z to the m-th diffusion term of model/cell.py:76-93
// (recurrence run on the identity in float64; x0/x1 carried across supports as the reference does)
__global__ void graph_poly_kernel(int B, int N, int K, int S, const float* s0, long long bs0,
                                  const float* s1, long long bs1, float* P) {
    __shared__ double Sm[NP * NP], X0[NP * NP], X1[NP * NP], X2[NP * NP];
    const int b = blockIdx.x, tid = threadIdx.x, NN = N * N;
    const int M1 = S * K;
    for (int i = tid; i < NN; i += blockDim.x) X0[i] = ((i / N) == (i % N)) ? 1.0 : 0.0;
    int m = 0;
    double *x0 = X0, *x1 = X1, *x2 = X2;
    for (int s = 0; s < S; ++s) {
        const float* sp = (s == 0) ? s0 + (size_t)b * bs0 : s1 + (size_t)b * bs1;
        __syncthreads();
        for (int i = tid; i < NN; i += blockDim.x) Sm[i] = (double)sp[i];
        __syncthreads();
        for (int i = tid; i < NN; i += blockDim.x) {           // x1 = S x0
            int r = i / N, c = i % N;
            double a = 0.0;
            for (int k = 0; k < N; ++k) a += Sm[r * N + k] * x0[k * N + c];
            x1[i] = a;
            P[((size_t)b * M1 + m) * NN + i] = (float)a;
        }
        ++m;
        for (int k2 = 2; k2 <= K; ++k2) {
            __syncthreads();
            for (int i = tid; i < NN; i += blockDim.x) {       // x2 = 2 S x1 - x0
                int r = i / N, c = i % N;
                double a = 0.0;
                for (int k = 0; k < N; ++k) a += Sm[r * N + k] * x1[k * N + c];
                a = 2.0 * a - x0[i];
                x2[i] = a;
                P[((size_t)b * M1 + m) * NN + i] = (float)a;
            }
            ++m;
            double* t = x0; x0 = x1; x1 = x2; x2 = t;          // x1, x0 = x2, x1
        }
    }
}

cudaError_t launch_graph_poly(int B, int N, int K, int S, const float* const* sup, const long long* bstride,
                              float* P, cudaStream_t st) {
    if (K == 0 || S == 0) return cudaSuccess;
    graph_poly_kernel<<<B, 128, 0, st>>>(B, N, K, S, sup[0], bstride[0], S > 1 ? sup[1] : nullptr,
                                         S > 1 ? bstride[1] : 0, P);
    return cudaGetLastError();
}

// One CTA per clip: |normalised zero-lag cross-correlation| -> directed top-k -> random-walk supports
// (data/dataloader_detection.py:258-307,343-347; data/data_utils.py:174-222; utils.py:220-230)
__global__ void __launch_bounds__(NT) corr_supports_kernel(int T, int N, int F, const float* clip, long long sb,
                                                           long long st_, float scale, float shift, int top_k,
                                                           float* adj_out, float* sup0, float* sup1) {
    extern __shared__ __align__(16) double sm[];
    const int FL = F | 1;                        // odd row stride: conflict-free across rows
    double* X = sm;                              // [N][FL] as double: one conversion per element, not one per pair and element
    __shared__ double G[NP * NP];
    __shared__ float A[NP * NP];
    __shared__ float dr[NP], dc[NP];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int npair = N * (N + 1) / 2;
    int pi = 0, pj = 0;
    if (tid < npair) {                           // decode pair index -> (i <= j)
        int rem = tid;
        for (int i = 0; i < N; ++i) {
            int cnt = N - i;
            if (rem < cnt) { pi = i; pj = i + rem; break; }
            rem -= cnt;
        }
    }
    double acc = 0.0;
    const float* base = clip + (size_t)b * sb;
    // the next step's slice is fetched into registers while this one is reduced (the loop was a chain of exposed global and
    // shared-memory latencies: 0.47 ms for 512 clips of 60 steps); four independent fp64 accumulators per pair
    constexpr int PF = 8;                        // N*F <= PF * NT elements per step (19 x 100 = 1900 <= 2048)
    const int nel = N * F;
    const bool pre = nel <= PF * NT;
    float nxt[PF];
    if (pre) {
#pragma unroll
        for (int k = 0; k < PF; ++k) { const int idx = tid + k * NT; nxt[k] = idx < nel ? __ldcs(base + idx) : 0.f; }
    }
    for (int t = 0; t < T; ++t) {
        __syncthreads();
        if (pre) {
#pragma unroll
            for (int k = 0; k < PF; ++k) {
                const int idx = tid + k * NT;
                if (idx < nel) { const int n = idx / F, f = idx - n * F; X[n * FL + f] = (double)(nxt[k] * scale + shift); }
            }
            if (t + 1 < T) {
#pragma unroll
                for (int k = 0; k < PF; ++k) { const int idx = tid + k * NT; nxt[k] = idx < nel ? __ldcs(base + (size_t)(t + 1) * st_ + idx) : 0.f; }
            }
        } else {
            for (int idx = tid; idx < nel; idx += NT) {
                int n = idx / F, f = idx - n * F;
                X[n * FL + f] = (double)(base[(size_t)t * st_ + idx] * scale + shift);
            }
        }
        __syncthreads();
        if (tid < npair) {
            const double* xi = X + pi * FL;
            const double* xj = X + pj * FL;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int f = 0;
            for (; f + 3 < F; f += 4) {
                a0 = fma(xi[f], xj[f], a0);
                a1 = fma(xi[f + 1], xj[f + 1], a1);
                a2 = fma(xi[f + 2], xj[f + 2], a2);
                a3 = fma(xi[f + 3], xj[f + 3], a3);
            }
            for (; f < F; ++f) a0 = fma(xi[f], xj[f], a0);
            acc += (a0 + a1) + (a2 + a3);
        }
    }
    if (tid < npair) { G[pi * N + pj] = acc; G[pj * N + pi] = acc; }
    __syncthreads();
    for (int idx = tid; idx < N * N; idx += NT) {
        int i = idx / N, j = idx - i * N;
        float v = 1.0f;
        if (i != j) {
            double g = G[idx], cxx = G[i * N + i], cyy = G[j * N + j];
            if (cxx != 0.0 && cyy != 0.0) g = g / sqrt(cxx * cyy);
            v = fabsf((float)g);
        }
        A[idx] = v;
    }
    __syncthreads();
    if (tid < N) {                               // keep the top_k largest off-diagonal entries of row tid
        const int i = tid;
        unsigned keep = 1u << i;
        for (int k = 0; k < top_k && k < N - 1; ++k) {
            int best = -1; float bv = -1.f;
            for (int j = 0; j < N; ++j) {
                if ((keep >> j) & 1u) continue;
                float v = A[i * N + j];
                if (v > bv) { bv = v; best = j; }
            }
            if (best >= 0) keep |= 1u << best;
        }
        for (int j = 0; j < N; ++j)
            if (!((keep >> j) & 1u)) A[i * N + j] = 0.f;
    }
    __syncthreads();
    if (tid < N) {
        float s = 0.f, c = 0.f;
        for (int j = 0; j < N; ++j) { s += A[tid * N + j]; c += A[j * N + tid]; }
        dr[tid] = (s != 0.f) ? 1.0f / s : 0.f;
        dc[tid] = (c != 0.f) ? 1.0f / c : 0.f;
    }
    __syncthreads();
    for (int idx = tid; idx < N * N; idx += NT) {
        int i = idx / N, j = idx - i * N;
        size_t o = (size_t)b * N * N + idx;
        if (adj_out != nullptr) adj_out[o] = A[idx];
        sup0[o] = dr[j] * A[j * N + i];          // (D^-1 A)^T
        sup1[o] = dc[j] * A[idx];                // (D_c^-1 A^T)^T
    }
}

cudaError_t launch_corr_supports(int B, int T, int N, int F, const float* clip, long long sb, long long st_,
                                 float scale, float shift, int top_k, float* adj, float* s0, float* s1,
                                 cudaStream_t st) {
    int smem = N * (F | 1) * 8;
    cudaError_t e = cudaFuncSetAttribute(corr_supports_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    corr_supports_kernel<<<B, NT, smem, st>>>(T, N, F, clip, sb, st_, scale, shift, top_k, adj, s0, s1);
    return cudaGetLastError();
}

}  // namespace dcgru
