// Minimal tcgen05 GEMM used to validate the UMMA plumbing of tc_common.cuh on hardware:
//   C[128 x N] = A[128 x K] * B[N x K]^T   (A, B row-major with K contiguous, K % 32 == 0, N in {64,128,192,256})
// Same structure as the production kernels: all threads stage K-chunks of both operands (hi/lo split) into
// double-buffered K-group-major tiles, one thread issues the 3xTF32 MMAs asynchronously, completion is
// tracked with mbarriers, the accumulator is read back from TMEM with tcgen05.ld.
#include "tc_common.cuh"
#include "common.cuh"

namespace dcgru {
using namespace tc;

constexpr int ST_KC = 32;                         // k per chunk (4 MMA k-steps)

__global__ void __launch_bounds__(NT, 1) tc_selftest_kernel(const float* A, const float* B, float* C, int N, int K) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t mbar[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int a_bytes = 128 * ST_KC * 4, b_bytes = N * ST_KC * 4;
    const int stage_bytes = 2 * a_bytes + 2 * b_bytes;
    if (warp == 0) tmem_alloc<256>(&tmem_slot);
    if (tid == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); mbar_fence_init(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = tmem_slot;
    const uint32_t idesc = make_idesc_tf32(128, N);
    const int nchunks = K / ST_KC;
    for (int i = 0; i < nchunks; ++i) {
        const int s = i & 1;
        if (i >= 2) mbar_wait(&mbar[s], ((i >> 1) - 1) & 1);
        uint8_t* st = smem_raw + s * stage_bytes;
        float4* a_hi = reinterpret_cast<float4*>(st);
        float4* a_lo = reinterpret_cast<float4*>(st + a_bytes);
        float4* b_hi = reinterpret_cast<float4*>(st + 2 * a_bytes);
        float4* b_lo = reinterpret_cast<float4*>(st + 2 * a_bytes + b_bytes);
        for (int idx = tid; idx < 128 * (ST_KC / 4); idx += NT) {
            int kg = idx / 128, r = idx - kg * 128;
            float4 v = *reinterpret_cast<const float4*>(A + (size_t)r * K + i * ST_KC + kg * 4);
            float4 h, l;
            split4(v, h, l);
            a_hi[kg * 128 + r] = h;
            a_lo[kg * 128 + r] = l;
        }
        for (int idx = tid; idx < N * (ST_KC / 4); idx += NT) {
            int kg = idx / N, r = idx - kg * N;
            float4 v = *reinterpret_cast<const float4*>(B + (size_t)r * K + i * ST_KC + kg * 4);
            float4 h, l;
            split4(v, h, l);
            b_hi[kg * N + r] = h;
            b_lo[kg * N + r] = l;
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            issue_3xtf32(taddr, smem_u32(a_hi), smem_u32(a_lo), 128, smem_u32(b_hi), smem_u32(b_lo), N,
                         ST_KC / 8, idesc, i > 0);
            umma_commit(&mbar[s]);
        }
    }
    mbar_wait(&mbar[(nchunks - 1) & 1], ((nchunks - 1) >> 1) & 1);
    tc_fence_after();
    {
        const int row = 32 * (warp & 3) + lane;
        const int half = warp >> 2, ncol = N / 2;
        for (int cb = half * ncol; cb < (half + 1) * ncol; cb += 32) {
            float v[32];
            tmem_ld32(taddr + ((uint32_t)(32 * (warp & 3)) << 16) + cb, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) C[(size_t)row * N + cb + j] = v[j];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(taddr);
}

// Layout probe: ONE kind::tf32 MMA (128 x N x 8) over caller-provided raw shared-memory images of A and B with
// caller-provided descriptor fields.  With one-hot / index-valued images the result shows which shared-memory
// word the tensor core reads for a logical (row, k): that is how the MN-major operand layout of dw_mm.cu was
// pinned down on hardware.
__global__ void __launch_bounds__(128, 1) tc_probe_kernel(const float* Aimg, int a_bytes, const float* Bimg, int b_bytes,
                                                          uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo,
                                                          uint32_t idesc, uint32_t a_type, uint32_t b_type, float* D, int N) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* sa = reinterpret_cast<float*>(smem_raw);
    float* sb = reinterpret_cast<float*>(smem_raw + 16384);
    for (int i = tid; i < a_bytes / 4; i += 128) sa[i] = Aimg[i];
    for (int i = tid; i < b_bytes / 4; i += 128) sb[i] = Bimg[i];
    if (warp == 0) tmem_alloc<256>(&tmem_slot);
    if (tid == 0) { mbar_init(&mbar, 1); mbar_fence_init(); }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = tmem_slot;
    if (tid == 0) {
        const uint64_t da = make_smem_desc(smem_u32(sa), a_lbo, a_sbo) | ((uint64_t)a_type << 61);
        const uint64_t db = make_smem_desc(smem_u32(sb), b_lbo, b_sbo) | ((uint64_t)b_type << 61);
        umma_tf32(taddr, da, db, idesc, 0u);
        umma_commit(&mbar);
    }
    mbar_wait(&mbar, 0);
    tc_fence_after();
    for (int cb = 0; cb < N; cb += 16) {
        float v[16];
        tmem_ld16(taddr + ((uint32_t)(32 * warp) << 16) + cb, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) D[(size_t)tid * N + cb + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(taddr);
}

cudaError_t launch_tc_probe(const float* Aimg, int a_bytes, const float* Bimg, int b_bytes, uint32_t a_lbo, uint32_t a_sbo,
                            uint32_t b_lbo, uint32_t b_sbo, uint32_t idesc, uint32_t a_type, uint32_t b_type, float* D, int N,
                            cudaStream_t st) {
    const int smem = 32768;
    cudaError_t e = cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    tc_probe_kernel<<<1, 128, smem, st>>>(Aimg, a_bytes, Bimg, b_bytes, a_lbo, a_sbo, b_lbo, b_sbo, idesc, a_type, b_type, D, N);
    return cudaGetLastError();
}

cudaError_t launch_tc_selftest(const float* A, const float* B, float* C, int N, int K, cudaStream_t st) {
    int smem = 2 * (2 * 128 * ST_KC * 4 + 2 * N * ST_KC * 4);
    cudaError_t e = cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    tc_selftest_kernel<<<1, NT, smem, st>>>(A, B, C, N, K);
    return cudaGetLastError();
}

}  // namespace dcgru
