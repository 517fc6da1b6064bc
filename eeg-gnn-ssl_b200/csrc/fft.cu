// Input features on the device: raw resampled EEG -> per-second log-amplitude spectrum -> random
// reflect / random scale augmentation -> standardisation, i.e. what the DataLoader workers do per clip in
//   data/dataloader_detection.py:58-72   (computeSliceMatrix: 1-second windows of FREQUENCY = 200 samples, is_fft)
//   data/data_utils.py:13-34             (computeFFT: fft(n = 200), first floor(n/2) = 100 bins, |.|, 0 -> 1e-8, log)
//   data/dataloader_detection.py:233-256 (_random_reflect: swap channel pairs; _random_scale: += log(scale) with FFT)
//   utils.py:402-403                     (StandardScaler.transform: (x - mean) / std)
// One window = 200 real samples -> 100 complex bins.  HBM-bound byte work (800 B in, 400 B out per window; 0.70 GB per
// 512-clip batch of 60-second clips), so the DFT must cost less than the stream: the real-input symmetries fold the
// 200-point DFT into four 50-term real sums per bin,
//     a_j = x_j + x_{200-j}, d_j = x_j - x_{200-j}                      (cos / sin are even / odd around j = 100)
//     even k:  Re = sum_{j<50} (a_j + a_{100-j}) cos(pi j k / 100) + a_50 cos(pi k / 2)
//              Im = -sum_{0<j<50} (d_j - d_{100-j}) sin(pi j k / 100)
//     odd  k:  Re = sum_{j<50} (a_j - a_{100-j}) cos(pi j k / 100)
//              Im = -[ sum_{0<j<50} (d_j + d_{100-j}) sin(pi j k / 100) + d_50 sin(pi k / 2) ]
// = 51 packed (re, im) FMAs (fma.rn.f32x2) instead of 400 scalar ones -- and they serve two bins: the twiddles of bin
// 100 - k are those of bin k times (-1)^j, so sums split by the parity of j give both.  Thread = bin k (and its mirror), its
// 51 (cos, sin) twiddle pairs live in registers for the whole kernel;
// a warp holds bins of one parity, so every folded (C_j, S_j) pair it needs comes from a broadcast LDS.128 for all 32 lanes.  A CTA
// (128 threads: warp = (bin parity, half of the windows)) is persistent over items = (channel of a clip, 16 consecutive windows):
// one cp.async.bulk fetches the next item's samples (12.8 KB, mbarrier) while this one is transformed ->
// fold into shared memory -> 51 packed FMAs per (bin, window) -> log amplitude -> shared -> coalesced stores with the
// augmentation and the scaler applied on the way out.  fp32 arithmetic on fp32 samples (the sums have 51 terms; measured
// error vs the float64 reference <= 2e-6 of the largest feature, tests/test_gpu_fft.py).
#include "common.cuh"
#include "dw.cuh"
#include "tc_common.cuh"

namespace dcgru {
using namespace tc;

constexpr int FFT_W = 200;           // samples per window (FREQUENCY * time_step_size, constants.py / args.py)
constexpr int FFT_K = FFT_W / 2;     // bins kept
constexpr int FFT_THREADS = 128;
constexpr int FFT_FW = 16;           // windows per batch
constexpr int FFT_FLD = 52;          // folded (C, S) pairs per parity and window (51 used)

struct FftParams {
    int B, N, T;
    const float* signal;             // (B, N, T*200)
    long long sig_sb, sig_sn;        // element strides of b and n
    const int* perm;                 // (B, N) destination channel of source channel n, or nullptr
    const float* log_scale;          // (B) added to the log amplitude, or nullptr
    const float* mean; const float* stdv; int stat_len;   // 0 (none), 1 or N
    float* raw;                      // (B, T, N, 100) log amplitude before augmentation / scaling, or nullptr
    float* x;                        // (B, T, N, 100) or nullptr
    const float* tw;                 // 400 floats: cos(pi m / 100), sin(pi m / 100), m = 0..199
    long long nwin;                  // B * N * T
};

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}

// Work item = (channel of a clip, block of FFT_FW consecutive windows): the samples of an item are one contiguous range
// whatever the strides of the signal, and the augmentation / scaler constants are per item.
__global__ void __launch_bounds__(FFT_THREADS, 3) fft_features_kernel(const FftParams p) {
    // raw samples of two items (the next one is in flight while this one is transformed): one cp.async.bulk per item
    __shared__ __align__(128) float rawb[2][FFT_FW][FFT_W];
    // folded values, interleaved so that one LDS.128 yields two ready (C_j, S_j) operand pairs of the packed FMA
    __shared__ __align__(16) float2 fold[FFT_FW][2][FFT_FLD];     // [window][even | odd bins][j] = (C_j, S_j)
    __shared__ __align__(16) float outb[FFT_FW][FFT_K];
    __shared__ uint64_t bar_raw[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // warp = (bin parity, window half).  A lane owns bin k AND its mirror 100 - k: the mirror's twiddles are the same up to
    // (-1)^j, so splitting the sums by the parity of j yields both bins from one set of FMAs (and one set of loads)
    const int parity = warp & 1;                                  // even / odd bins
    const int team = warp >> 1;                                   // windows [8 team, 8 team + 8)
    const int k = 2 * lane + parity;                              // even: 0..50 (lanes 0..25), odd: 1..49 (lanes 0..24)
    const bool kvalid = k <= 50;
    const int kmir = 100 - k;
    const bool mvalid = kvalid && kmir > 50 && kmir < 100;        // k = 0 and k = 50 have no mirror among the 100 bins kept
    unsigned long long tw2[51];                                   // (cos, sin)(pi j k / 100) as packed pairs
#pragma unroll
    for (int j = 0; j <= 50; ++j) {
        const int m = (j * (kvalid ? k : 0)) % 200;
        tw2[j] = pack2(p.tw[m], p.tw[200 + m]);
    }
    if (tid == 0) {
        mbar_init(&bar_raw[0], 1);
        mbar_init(&bar_raw[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int nbt = (p.T + FFT_FW - 1) / FFT_FW;                  // window blocks per channel
    const int nitem = p.B * p.N * nbt;
    auto prefetch = [&](int item, int buf) {                      // one thread
        const int pair = item / nbt, tb = item - pair * nbt;
        const int b = pair / p.N, n = pair - b * p.N;
        const int t0 = tb * FFT_FW;
        const int nw = (p.T - t0 < FFT_FW) ? (p.T - t0) : FFT_FW;
        const uint32_t bytes = (uint32_t)(nw * FFT_W * 4);
        mbar_expect_tx(&bar_raw[buf], bytes);
        bulk_g2s(&rawb[buf][0][0], p.signal + (long long)b * p.sig_sb + (long long)n * p.sig_sn + (long long)t0 * FFT_W, bytes,
                 &bar_raw[buf]);
    };
    int item = blockIdx.x;
    if (tid == 0 && item < nitem) prefetch(item, 0);
    for (int it = 0; item < nitem; item += gridDim.x, ++it) {
        const int buf = it & 1;
        const int pair = item / nbt, tb = item - pair * nbt;
        const int b = pair / p.N, n = pair - b * p.N;
        const int t0 = tb * FFT_FW;
        const int nw = (p.T - t0 < FFT_FW) ? (p.T - t0) : FFT_FW;
        // constants of the item (loaded now, used by the stores at the end: the latency hides behind the transform)
        const int no = p.perm ? p.perm[(long long)b * p.N + n] : n;           // destination channel of source channel n
        const float ls = p.log_scale ? p.log_scale[b] : 0.f;
        const float mean = p.stat_len ? p.mean[p.stat_len == 1 ? 0 : no] : 0.f;
        const float sd = p.stat_len ? p.stdv[p.stat_len == 1 ? 0 : no] : 1.f;
        __syncthreads();                                          // every thread is done with item it-1 (fold, outb, rawb[buf^1])
        if (tid == 0 && item + (int)gridDim.x < nitem) prefetch(item + gridDim.x, buf ^ 1);
        mbar_wait(&bar_raw[buf], (it >> 1) & 1);
        // ---- fold: work unit = (window, j), j = 0..50: x_j, x_{200-j}, x_{100-j}, x_{100+j} -------------------------------
        for (int i = tid; i < nw * 51; i += FFT_THREADS) {
            const int w = i / 51, j = i - w * 51;
            const float* s = rawb[buf][w];
            float ce, se, co, so;
            if (j == 0) {
                const float x0 = s[0], x100 = s[100];
                ce = x0 + x100; co = x0 - x100; se = 0.f; so = 0.f;
            } else if (j == 50) {
                const float x50 = s[50], x150 = s[150];
                ce = x50 + x150;            // a_50 (even bins: times cos(pi k / 2))
                co = 0.f;
                se = 0.f;
                so = x50 - x150;            // d_50 (odd bins: times sin(pi k / 2))
            } else {
                const float xa = s[j], xb = s[200 - j], xc = s[100 - j], xd = s[100 + j];
                const float aj = xa + xb, dj = xa - xb;           // a_j, d_j
                const float ar = xc + xd, dr = xc - xd;           // a_{100-j}, d_{100-j}
                ce = aj + ar; co = aj - ar; se = dj - dr; so = dj + dr;
            }
            fold[w][0][j] = make_float2(ce, se);
            fold[w][1][j] = make_float2(co, so);
            if (j == 50) { fold[w][0][51] = make_float2(0.f, 0.f); fold[w][1][51] = make_float2(0.f, 0.f); }
        }
        __syncthreads();
        // ---- 51 packed FMAs per (bin, window): (re, im) += (C_j, S_j) * (cos, sin); folded values are warp-wide broadcasts.
        //      Unrolled over the windows so that the loads of window w+1 and the logarithm of window w-1 overlap the FMAs of w.
        if (kvalid) {
            float pw[FFT_FW / 2], pm[FFT_FW / 2];
#pragma unroll
            for (int wi = 0; wi < FFT_FW / 2; ++wi) {
                const int w = team * (FFT_FW / 2) + wi;
                pw[wi] = 1.f; pm[wi] = 1.f;
                if (w < nw) {
                    const ulonglong2* F = reinterpret_cast<const ulonglong2*>(fold[w][parity]);
                    unsigned long long e0 = 0ull, e1 = 0ull, o0 = 0ull, o1 = 0ull;      // (re, im) sums over even / odd j, two chains each
#pragma unroll
                    for (int q = 0; q < 24; q += 2) {
                        const ulonglong2 f = F[q], g = F[q + 1];                         // j = 2q, 2q+1 | 2q+2, 2q+3
                        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(e0) : "l"(f.x), "l"(tw2[2 * q]));
                        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(o0) : "l"(f.y), "l"(tw2[2 * q + 1]));
                        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(e1) : "l"(g.x), "l"(tw2[2 * q + 2]));
                        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(o1) : "l"(g.y), "l"(tw2[2 * q + 3]));
                    }
                    {
                        const ulonglong2 f = F[24], g = F[25];        // j = 48, 49, 50 (51 is padding)
                        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(e0) : "l"(f.x), "l"(tw2[48]));
                        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(o0) : "l"(f.y), "l"(tw2[49]));
                        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(e1) : "l"(g.x), "l"(tw2[50]));
                    }
                    const float ere = __uint_as_float((unsigned)e0) + __uint_as_float((unsigned)e1);
                    const float eim = __uint_as_float((unsigned)(e0 >> 32)) + __uint_as_float((unsigned)(e1 >> 32));
                    const float ore = __uint_as_float((unsigned)o0) + __uint_as_float((unsigned)o1);
                    const float oim = __uint_as_float((unsigned)(o0 >> 32)) + __uint_as_float((unsigned)(o1 >> 32));
                    const float re = ere + ore, im = eim + oim, rm = ere - ore, imm = eim - oim;
                    pw[wi] = fmaf(re, re, im * im);               // |X_k|^2;  log|X| = log(|X|^2) / 2
                    pm[wi] = fmaf(rm, rm, imm * imm);             // |X_{100-k}|^2
                }
            }
#pragma unroll
            for (int wi = 0; wi < FFT_FW / 2; ++wi) {
                const int w = team * (FFT_FW / 2) + wi;
                if (w < nw) {
                    outb[w][k < FFT_K ? k : 0] = pw[wi] == 0.0f ? -18.420680743952367f : 0.5f * logf(pw[wi]);      // data_utils.py:30: amp == 0 -> 1e-8
                    if (mvalid) outb[w][kmir] = pm[wi] == 0.0f ? -18.420680743952367f : 0.5f * logf(pm[wi]);
                }
            }
        }
        __syncthreads();
        // ---- coalesced stores: raw features, and augmented + standardised x ------------------------------------------
        const long long dst0 = (((long long)b * p.T + t0) * p.N + n) * FFT_K, dstx0 = (((long long)b * p.T + t0) * p.N + no) * FFT_K;
        const long long wstride = (long long)p.N * FFT_K;         // next window of the same channel
        for (int i = tid; i < nw * (FFT_K / 4); i += FFT_THREADS) {
            const int w = i / (FFT_K / 4), q = i - w * (FFT_K / 4);
            float4 v = *reinterpret_cast<const float4*>(&outb[w][4 * q]);
            if (p.raw) __stcs(reinterpret_cast<float4*>(p.raw + dst0 + w * wstride) + q, v);
            if (p.x) {
                if (p.log_scale) { v.x += ls; v.y += ls; v.z += ls; v.w += ls; }
                if (p.stat_len > 0) { v.x = (v.x - mean) / sd; v.y = (v.y - mean) / sd; v.z = (v.z - mean) / sd; v.w = (v.w - mean) / sd; }
                __stcs(reinterpret_cast<float4*>(p.x + dstx0 + w * wstride) + q, v);
            }
        }
    }
}

// twiddle table cos / sin(pi m / 100), m = 0..199: written by a one-block kernel in front of every launch (float64 sincospi,
// exact zeros and ones at the multiples of pi / 2) -- stream-ordered, capturable, nothing allocated by the library
__device__ float g_fft_tw[400];
__global__ void fft_twiddle_kernel() {
    const int m = threadIdx.x;
    if (m < 200) {
        double s, c;
        sincospi((double)m / 100.0, &s, &c);
        g_fft_tw[m] = (float)c;
        g_fft_tw[200 + m] = (float)s;
    }
}

cudaError_t launch_fft_features(int B, int N, int T, const float* signal, long long sig_sb, long long sig_sn, const int* perm,
                                const float* log_scale, const float* mean, const float* stdv, int stat_len, float* raw, float* x,
                                int nsms, cudaStream_t st) {
    FftParams p;
    p.B = B; p.N = N; p.T = T; p.signal = signal; p.sig_sb = sig_sb; p.sig_sn = sig_sn; p.perm = perm; p.log_scale = log_scale;
    p.mean = mean; p.stdv = stdv; p.stat_len = stat_len; p.raw = raw; p.x = x;
    p.nwin = (long long)B * N * T;
    float* tw = nullptr;
    cudaError_t e = cudaGetSymbolAddress(reinterpret_cast<void**>(&tw), g_fft_tw);
    if (e != cudaSuccess) return e;
    p.tw = tw;
    fft_twiddle_kernel<<<1, 256, 0, st>>>();
    long long nbatch = (long long)B * N * ((T + FFT_FW - 1) / FFT_FW);
    long long grid = (long long)nsms * 3;                        // persistent: 3 CTAs of 128 threads per SM (register-limited)
    if (grid > nbatch) grid = nbatch;
    fft_features_kernel<<<(unsigned)grid, FFT_THREADS, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace dcgru
