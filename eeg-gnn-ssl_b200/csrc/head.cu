// Fused classification head of DCRNNModel_classification (model/model.py:257-270):
//     last = h_seq[seq_len[b]-1, b]            (utils.last_relevant_pytorch, utils.py:346-357)
//     a    = relu(dropout(last))               (B, N, H)
//     z    = a @ fc_w^T + fc_b                 (B, N, C)
//     logits[b, c] = max_n z[b, n, c]          (torch.max over nodes)
// and its backward, which leaves the upstream gradient of the encoder's top layer in the SPARSE form the BPTT kernels
// take: one (B, N*H) slab d_hsel that belongs to step sel_t[b] of sample b, instead of the dense (T, B, N*H) tensor the
// gather's autograd would materialise (149 MB of zeros per step at BASELINE config 2).
// HBM-bound and tiny (B*N*H floats in, B*C out): one CTA per sample, 128 threads, fixed-order reductions (deterministic).
#include <cstring>

#include "common.cuh"
#include "dw.cuh"

namespace dcgru {

constexpr int HD_THREADS = 128;
constexpr int HD_MAXC = 16;

// a[n][j] of one sample into shared memory; returns nothing.  drop (B,N,H) may be nullptr.
__device__ __forceinline__ void head_load_a(const float* hseq, const int* sel_t, const float* drop, int B, int T, int N, int H,
                                            int b, float* a) {
    int t = sel_t ? sel_t[b] : T - 1;
    t = t < 0 ? 0 : (t >= T ? T - 1 : t);
    const float* src = hseq + ((size_t)t * B + b) * N * H;
    const float* dm = drop ? drop + (size_t)b * N * H : nullptr;
    for (int i = threadIdx.x; i < N * H; i += HD_THREADS) {
        float v = src[i];
        if (dm) v *= dm[i];
        a[i] = v > 0.f ? v : 0.f;
    }
}

__global__ void __launch_bounds__(HD_THREADS) cls_head_fwd_kernel(int B, int T, int N, int H, int C, const float* hseq, const int* sel_t,
                                                                  const float* drop, const float* W, const float* bias, float* logits,
                                                                  int* arg) {
    extern __shared__ float sm[];
    float* a = sm;                       // N*H
    float* z = sm + N * H;               // N*C
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    head_load_a(hseq, sel_t, drop, B, T, N, H, b, a);
    __syncthreads();
    for (int pair = warp; pair < N * C; pair += HD_THREADS / 32) {       // one warp per (node, class) dot product
        const int n = pair / C, c = pair - n * C;
        float s = 0.f;
        for (int j = lane; j < H; j += 32) s = fmaf(a[n * H + j], W[c * H + j], s);
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) z[pair] = s + bias[c];
    }
    __syncthreads();
    if (threadIdx.x < C) {
        const int c = threadIdx.x;
        float best = z[c];
        int bi = 0;
        for (int n = 1; n < N; ++n) {
            const float v = z[n * C + c];
            if (v > best) { best = v; bi = n; }                           // first maximum wins
        }
        logits[(size_t)b * C + c] = best;
        arg[(size_t)b * C + c] = bi;
    }
}

// d_hsel[b] (N*H) = sum_c [n == arg[b,c]] d_logits[b,c] * fc_w[c] * d relu(drop*h)/dh;  dwpart[b][c][j] = d_logits[b,c] * a[arg][j]
__global__ void __launch_bounds__(HD_THREADS) cls_head_bwd_kernel(int B, int T, int N, int H, int C, const float* hseq, const int* sel_t,
                                                                  const float* drop, const float* W, const int* arg,
                                                                  const float* dlogits, float* d_hsel, float* dwpart) {
    extern __shared__ float sm[];
    float* a = sm;
    __shared__ int s_arg[HD_MAXC];
    __shared__ float s_dl[HD_MAXC];
    const int b = blockIdx.x;
    head_load_a(hseq, sel_t, drop, B, T, N, H, b, a);
    if (threadIdx.x < C) {
        s_arg[threadIdx.x] = arg[(size_t)b * C + threadIdx.x];
        s_dl[threadIdx.x] = dlogits[(size_t)b * C + threadIdx.x];
    }
    __syncthreads();
    const float* dm = drop ? drop + (size_t)b * N * H : nullptr;
    for (int i = threadIdx.x; i < N * H; i += HD_THREADS) {
        const int n = i / H, j = i - n * H;
        float g = 0.f;
        for (int c = 0; c < C; ++c)
            if (s_arg[c] == n) g = fmaf(s_dl[c], W[c * H + j], g);
        g = a[i] > 0.f ? g : 0.f;
        if (dm) g *= dm[i];
        d_hsel[(size_t)b * N * H + i] = g;
    }
    for (int i = threadIdx.x; i < C * H; i += HD_THREADS) {
        const int c = i / H, j = i - c * H;
        dwpart[(size_t)b * C * H + i] = s_dl[c] * a[s_arg[c] * H + j];
    }
}

// d_fc_w[c][j] = sum_b dwpart[b][c][j] (fixed order), d_fc_b[c] = sum_b d_logits[b][c]
__global__ void cls_head_reduce_kernel(int B, int C, int H, const float* dwpart, const float* dlogits, float* dW, float* db) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < C * H) {
        float s = 0.f;
        for (int b = 0; b < B; ++b) s += dwpart[(size_t)b * C * H + i];
        dW[i] = s;
    }
    if (i < C) {
        float s = 0.f;
        for (int b = 0; b < B; ++b) s += dlogits[(size_t)b * C + i];
        db[i] = s;
    }
}

// dense[t][b][:] = (t == clamp(sel_t[b])) ? d_hsel[b][:] : 0   -- bridge for the BPTT kernels that take a dense gradient
__global__ void scatter_sel_kernel(int B, int T, int NH, const float* d_hsel, const int* sel_t, float* dense) {
    const size_t total = (size_t)T * B * (NH / 4);
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int q = (int)(idx % (NH / 4));
        const size_t r = idx / (NH / 4);
        const int b = (int)(r % B), t = (int)(r / B);
        int s = sel_t ? sel_t[b] : T - 1;
        s = s < 0 ? 0 : (s >= T ? T - 1 : s);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s == t) v = reinterpret_cast<const float4*>(d_hsel + (size_t)b * NH)[q];
        reinterpret_cast<float4*>(dense)[idx] = v;
    }
}

bool cls_head_supported(int N, int H, int C) { return N >= 1 && N <= 32 && H >= 1 && H <= 1024 && C >= 1 && C <= HD_MAXC; }

cudaError_t launch_cls_head_fwd(int B, int T, int N, int H, int C, const float* hseq, const int* sel_t, const float* drop,
                                const float* W, const float* bias, float* logits, int* arg, cudaStream_t st) {
    const size_t smem = (size_t)(N * H + N * C) * 4;
    cudaError_t e = cudaFuncSetAttribute(cls_head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cls_head_fwd_kernel<<<B, HD_THREADS, smem, st>>>(B, T, N, H, C, hseq, sel_t, drop, W, bias, logits, arg);
    return cudaGetLastError();
}

cudaError_t launch_cls_head_bwd(int B, int T, int N, int H, int C, const float* hseq, const int* sel_t, const float* drop,
                                const float* W, const int* arg, const float* dlogits, float* d_hsel, float* dW, float* db,
                                float* dwpart, cudaStream_t st) {
    const size_t smem = (size_t)N * H * 4;
    cudaError_t e = cudaFuncSetAttribute(cls_head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cls_head_bwd_kernel<<<B, HD_THREADS, smem, st>>>(B, T, N, H, C, hseq, sel_t, drop, W, arg, dlogits, d_hsel, dwpart);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int n = C * H;
    cls_head_reduce_kernel<<<(n + 127) / 128, 128, 0, st>>>(B, C, H, dwpart, dlogits, dW, db);
    return cudaGetLastError();
}

// out = a (+ b) (* m): the two element-wise steps of the decoder's BPTT orchestration (capi.cu): total gradient of out_t =
// upstream + autoregressive feedback, and the dropout mask on the projection's input gradient.  n % 4 == 0, 16-byte aligned.
__global__ void ew_add_mul_kernel(const float* a, const float* b, const float* m, float* out, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<const float4*>(a)[i];
        if (b) { const float4 w = reinterpret_cast<const float4*>(b)[i]; v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w; }
        if (m) { const float4 w = reinterpret_cast<const float4*>(m)[i]; v.x *= w.x; v.y *= w.y; v.z *= w.z; v.w *= w.w; }
        reinterpret_cast<float4*>(out)[i] = v;
    }
}
cudaError_t launch_ew_add_mul(const float* a, const float* b, const float* m, float* out, size_t n, cudaStream_t st) {
    const size_t n4 = n / 4;
    int grid = (int)((n4 + 255) / 256);
    if (grid > 1184) grid = 1184;
    if (grid < 1) grid = 1;
    ew_add_mul_kernel<<<grid, 256, 0, st>>>(a, b, m, out, n4);
    return cudaGetLastError();
}

cudaError_t launch_scatter_sel(int B, int T, int NH, const float* d_hsel, const int* sel_t, float* dense, cudaStream_t st) {
    scatter_sel_kernel<<<1184, 256, 0, st>>>(B, T, NH, d_hsel, sel_t, dense);
    return cudaGetLastError();
}

}  // namespace dcgru
