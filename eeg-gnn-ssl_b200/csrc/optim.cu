// Fused optimiser step over the flat gradient / parameter buffers: global-norm clip + Adam with
// L2 weight decay, the tail of the reference's training step (train.py:273-275: clip_grad_norm_, optimizer.step with
// torch.optim.Adam(lr, weight_decay)).  Two launches instead of ~20 element-wise kernels:
//   1. sumsq_partial_kernel : fixed-order partial sums of g^2 (deterministic); also advances the device step counter
//   2. clip_adam_kernel     : every block re-reduces the partials (<= 512 values) to the total norm, then
//                             g <- gs * g (gs = 1/world folds the data-parallel average into this pass)
//                             g <- g * min(1, max_norm / (norm + 1e-6));  g' = g + wd * p
//                             m <- b1 m + (1-b1) g';  v <- b2 v + (1-b2) g'^2
//                             p <- p - lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// Learning rate and step count live in device memory, so the step can sit inside a CUDA graph while a scheduler
// (the reference's CosineAnnealingLR, train.py:224) changes the rate between replays.
#include "common.cuh"
#include "dw.cuh"

namespace dcgru {

constexpr int OPT_MAXPART = 512;

// partial sums in double: torch's clip_grad_norm_ reduces per-tensor norms, this reduces one flat buffer; double
// keeps the two within rounding of each other for any parameter count
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, size_t n, double* partial, int* step) {
    __shared__ double red[256];
    const size_t per = (n + gridDim.x - 1) / gridDim.x;
    const size_t i0 = (size_t)blockIdx.x * per, i1 = (i0 + per < n) ? i0 + per : n;
    double a = 0.0;
    for (size_t i = i0 + threadIdx.x; i < i1; i += 256) { const double x = g[i]; a = fma(x, x, a); }
    red[threadIdx.x] = a;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = red[0];
        if (blockIdx.x == 0) step[0] += 1;
    }
}

__global__ void __launch_bounds__(256) clip_adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, size_t n, const double* partial, int npart,
                                                        const float* lr_dev, const int* step, float beta1, float beta2,
                                                        float eps, float wd, float max_norm, float gscale, float* norm_out) {
    __shared__ double red[256];
    double a = 0.0;
    for (int i = threadIdx.x; i < npart; i += 256) a += partial[i];       // same order in every block
    red[threadIdx.x] = a;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    const float norm = (float)(sqrt(red[0]) * (double)gscale);             // norm of the scaled (averaged) gradient
    if (blockIdx.x == 0 && threadIdx.x == 0 && norm_out) norm_out[0] = norm;
    float coef = 1.f;
    if (max_norm > 0.f) { coef = max_norm / (norm + 1e-6f); if (coef > 1.f) coef = 1.f; }
    coef *= gscale;
    // bias corrections in double, once per thread (torch computes them in double on the host; 1 - beta2^t in fp32
    // loses ~6e-5 relative at t = 1), integer step count
    const double t = (double)step[0];
    const double bc1 = 1.0 - pow((double)beta1, t), bc2 = 1.0 - pow((double)beta2, t);
    const float step_size = (float)((double)lr_dev[0] / bc1), inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const float gc = g[i] * coef;
        const float pi = p[i];
        const float gd = fmaf(wd, pi, gc);
        const float mi = fmaf(beta1, m[i], (1.f - beta1) * gd);
        const float vi = fmaf(beta2, v[i], (1.f - beta2) * gd * gd);
        g[i] = gc;
        m[i] = mi;
        v[i] = vi;
        p[i] = pi - step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
    }
}

int clip_adam_npart(size_t n) {
    size_t b = (n + 4095) / 4096;
    if (b < 1) b = 1;
    if (b > OPT_MAXPART) b = OPT_MAXPART;
    return (int)b;
}

cudaError_t launch_clip_adam(float* p, float* g, float* m, float* v, size_t n, const float* lr_dev, int* step,
                             float beta1, float beta2, float eps, float wd, float max_norm, float gscale,
                             double* partial, float* norm_out, cudaStream_t st) {
    const int npart = clip_adam_npart(n);
    sumsq_partial_kernel<<<npart, 256, 0, st>>>(g, n, partial, step);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    size_t blocks = (n + 1023) / 1024;
    if (blocks < 1) blocks = 1;
    if (blocks > 592) blocks = 592;
    clip_adam_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, g, m, v, n, partial, npart, lr_dev, step, beta1, beta2, eps, wd,
                                                       max_norm, gscale, norm_out);
    return cudaGetLastError();
}

}  // namespace dcgru
