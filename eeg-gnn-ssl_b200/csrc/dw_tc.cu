// Tensor-core (tcgen05, 3xTF32) weight-gradient kernel:  dW[kk][o] = sum_rows G[row][kk] * dA[row][o].
//
// GEMM view: M = kk (a 128-row slab of W rows per CTA), N = o (<= 192 output columns), K = rows =
// every (step, sample, node) -- 583 680 of them at BASELINE config 2 -- split over CTAs.
//   A[m = kk][k = row]  = diffuse(Z)^T : produced on the fly by the CTA's threads from the saved
//                         activations (x, h_prev, r*h_prev) and the per-sample polynomials P_m
//   B[n = o][k = row]   = dA^T         : loaded row-major from HBM, transposed in registers
// Both are written hi/lo-split straight into K-group-major UMMA tiles (tc_common.cuh), 40 K-rows
// (= 2 sample steps x 20 padded nodes = 5 MMA k-steps) per chunk, double buffered: while the tensor
// core consumes chunk i the threads produce chunk i+1.  The accumulator (128 x N fp32) lives in
// TMEM for the CTA's whole row range and is written once as a split-K partial; dw.cu's
// reduce_cell_kernel sums the partials in fixed order.  The bias gradient rides along as an extra
// all-ones A row in a slab that has a spare row (db[o] = sum_rows 1 * dA[row][o]).
#include "common.cuh"
#include "dw.cuh"
#include "tc_common.cuh"

namespace dcgru {
using namespace tc;

constexpr int TKR = 40;                 // K rows per chunk
constexpr int TKG = TKR / 4;            // K groups per chunk
constexpr int TZC = 48;                 // max source columns a 128-row slab can touch (+ slack)

struct DwTcSmem {
    int a_bytes, b_bytes, stage_bytes, zc_off, pt_off, total;
};
__host__ __device__ inline DwTcSmem dwtc_smem(int M, int nco) {
    DwTcSmem s;
    s.a_bytes = TKG * 128 * 16;                     // one of hi / lo
    s.b_bytes = TKG * nco * 16;
    s.stage_bytes = 2 * s.a_bytes + 2 * s.b_bytes;
    s.zc_off = 2 * s.stage_bytes;
    s.pt_off = s.zc_off + 2 * NP * TZC * 4;
    s.total = s.pt_off + 2 * (M - 1) * NP * NP * 4;
    return s;
}

__global__ void __launch_bounds__(NT, 1) dw_tc_kernel(const DwParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t mbar[2];
    __shared__ uint32_t tmem_slot;
    const DwJob job = p.jobs[blockIdx.x];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = p.N, H = p.H, M = p.M, M1 = M - 1, B = p.B, H3 = 3 * H;
    const int nco = job.nco, kk0 = job.kk0, nkk = job.nz;          // nz field = rows of this slab
    const DwTcSmem L = dwtc_smem(M, nco);
    float* Zc = reinterpret_cast<float*>(smem + L.zc_off);          // [2][NP][TZC]
    float* PT = reinterpret_cast<float*>(smem + L.pt_off);          // [2][M1][NP(j)][NP(n)]
    // source columns this slab needs: absolute z columns c_lo..c_hi of [x | h]
    const int c_lo = kk0 / M, c_hi = (kk0 + nkk - 1) / M;
    const int ncz = (nkk > 0) ? (c_hi - c_lo + 1) : 0;
    const int zoff = (job.type == 0) ? 0 : p.fin;                    // h columns start after the x columns
    const bool ones_row = (job.z0 != 0);                             // z0 field = "carry the db row" flag

    if (warp == 0) tmem_alloc<256>(&tmem_slot);
    if (tid == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); mbar_fence_init(); }
    for (int idx = tid; idx < 2 * L.stage_bytes / 16; idx += NT)
        reinterpret_cast<float4*>(smem)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = tmem_slot;
    const uint32_t idesc = make_idesc_tf32(128, nco);

    const size_t NH = (size_t)N * H;
    const long total_q = (long)p.T * B;                              // sample steps
    const long nchunk = (total_q + 1) / 2;
    int it = 0;
    for (long ch = blockIdx.y; ch < nchunk; ch += gridDim.y, ++it) {
        const int s_ = it & 1;
        if (it >= 2) mbar_wait(&mbar[s_], ((it >> 1) - 1) & 1);
        uint8_t* st = smem + s_ * L.stage_bytes;
        float4* a_hi = reinterpret_cast<float4*>(st);
        float4* a_lo = reinterpret_cast<float4*>(st + L.a_bytes);
        float4* b_hi = reinterpret_cast<float4*>(st + 2 * L.a_bytes);
        float4* b_lo = reinterpret_cast<float4*>(st + 2 * L.a_bytes + L.b_bytes);
        // ---- the two sample steps of this chunk ------------------------------------------------------
        long q[2] = {2 * ch, 2 * ch + 1};
        int tq[2], bq[2];
        bool vq[2];
        for (int s = 0; s < 2; ++s) {
            vq[s] = q[s] < total_q;
            tq[s] = vq[s] ? (int)(q[s] / B) : 0;
            bq[s] = vq[s] ? (int)(q[s] % B) : 0;
        }
        // ---- PT[s][m1][j][n] and Zc[s][j][ccl] ---------------------------------------------------------
        for (int idx = tid; idx < 2 * M1 * NP * NP; idx += NT) {
            int n = idx % NP, j = (idx / NP) % NP, sm = idx / (NP * NP);
            int s = sm / max(M1, 1), m1 = sm - s * max(M1, 1);
            float v = 0.f;
            if (n < N && j < N && vq[s]) v = p.P[(((size_t)bq[s] * M1 + m1) * N + n) * N + j];
            PT[idx] = v;
        }
        for (int idx = tid; idx < 2 * NP * ncz; idx += NT) {
            int ccl = idx % ncz, j = (idx / ncz) % NP, s = idx / (ncz * NP);
            float v = 0.f;
            if (j < N && vq[s]) {
                const int t = tq[s], b = bq[s];
                const int c = c_lo + ccl - zoff;                    // column within x or within h
                const size_t ro = (size_t)b * N + j;
                if (job.type == 0) {
                    const float* xs; long long xsb;
                    if (p.mode == 0) { xs = p.x + (size_t)t * p.xs_t; xsb = p.xs_b; }
                    else {
                        if (p.layer == 0) {
                            xsb = (long long)N * p.Fo;
                            if (t == 0) xs = nullptr;
                            else if ((p.teacher_mask >> (t - 1)) & 1ull) xs = p.targets + (size_t)(t - 1) * B * N * p.Fo;
                            else xs = p.out + (size_t)(t - 1) * B * N * p.Fo;
                        } else { xsb = (long long)NH; xs = p.hseq + ((size_t)t * p.ncell + (p.layer - 1)) * B * NH; }
                    }
                    if (xs != nullptr) v = xs[(size_t)b * xsb + j * p.fin + c];
                } else {
                    const float* hp = (p.mode == 0)
                        ? ((t == 0) ? p.h0 : p.hseq + (size_t)(t - 1) * B * NH)
                        : ((t == 0) ? p.h0 + (size_t)p.layer * B * NH
                                    : p.hseq + ((size_t)(t - 1) * p.ncell + p.layer) * B * NH);
                    v = hp[ro * H + c];
                    if (job.type == 2) {
                        const float* rc = (p.mode == 0) ? p.ruc + (size_t)t * B * NH * 3
                                                        : p.ruc + ((size_t)t * p.ncell + p.layer) * B * NH * 3;
                        v *= rc[ro * H3 + c];
                    }
                }
            }
            Zc[(s * NP + j) * TZC + ccl] = v;
        }
        __syncthreads();
        // ---- A tile: G^T, one task = (sample, source column, node quad) ---------------------------------
        for (int id = tid; id < 2 * ncz * 5; id += NT) {
            int ccl = id % ncz, t1 = id / ncz;
            int q5 = t1 % 5, s = t1 / 5;
            const float* zp = Zc + (s * NP) * TZC + ccl;
            const int kg = s * 5 + q5;
            const int kbase = (c_lo + ccl) * M - kk0;                // local row of the m = 0 term
            {   // m = 0: identity
                if (kbase >= 0 && kbase < nkk) {
                    float4 v = make_float4(zp[(4 * q5) * TZC], zp[(4 * q5 + 1) * TZC], zp[(4 * q5 + 2) * TZC],
                                           zp[(4 * q5 + 3) * TZC]);
                    float4 h, l;
                    split4(v, h, l);
                    a_hi[kg * 128 + kbase] = h;
                    a_lo[kg * 128 + kbase] = l;
                }
            }
            for (int m1 = 0; m1 < M1; ++m1) {
                const int kl = kbase + m1 + 1;
                if (kl < 0 || kl >= nkk) continue;
                const float* pp = PT + ((size_t)(s * M1 + m1) * NP) * NP + 4 * q5;
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int j = 0; j < N; ++j) {
                    float z = zp[j * TZC];
                    float4 pv = *reinterpret_cast<const float4*>(pp + j * NP);
                    a.x = fmaf(pv.x, z, a.x); a.y = fmaf(pv.y, z, a.y);
                    a.z = fmaf(pv.z, z, a.z); a.w = fmaf(pv.w, z, a.w);
                }
                float4 h, l;
                split4(a, h, l);
                a_hi[kg * 128 + kl] = h;
                a_lo[kg * 128 + kl] = l;
            }
        }
        if (ones_row && tid < TKG) {                                 // db row: 1 for every real (sample, node)
            int s = tid / 5, q5 = tid % 5;
            float4 v;
            v.x = (vq[s] && 4 * q5 + 0 < N) ? 1.f : 0.f;
            v.y = (vq[s] && 4 * q5 + 1 < N) ? 1.f : 0.f;
            v.z = (vq[s] && 4 * q5 + 2 < N) ? 1.f : 0.f;
            v.w = (vq[s] && 4 * q5 + 3 < N) ? 1.f : 0.f;
            a_hi[tid * 128 + 127] = v;
        }
        // ---- B tile: dA^T, one task = (K group, 4 output columns) ------------------------------------------
        {
            const int noq = nco >> 2;
            for (int id = tid; id < TKG * noq; id += NT) {
                int oq = id % noq, kg = id / noq;
                int s = kg / 5, q5 = kg - s * 5;
                float4 r[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    int n = 4 * q5 + i;
                    r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (vq[s] && n < N) {
                        const float* da = (p.mode == 0) ? p.dA + (size_t)tq[s] * B * NH * 3
                                                        : p.dA + ((size_t)tq[s] * p.ncell + p.layer) * B * NH * 3;
                        r[i] = *reinterpret_cast<const float4*>(da + ((size_t)bq[s] * N + n) * H3 + job.o0 + 4 * oq);
                    }
                }
                float4 c0 = make_float4(r[0].x, r[1].x, r[2].x, r[3].x);
                float4 c1 = make_float4(r[0].y, r[1].y, r[2].y, r[3].y);
                float4 c2 = make_float4(r[0].z, r[1].z, r[2].z, r[3].z);
                float4 c3 = make_float4(r[0].w, r[1].w, r[2].w, r[3].w);
                float4 h, l;
                float4* bh = b_hi + kg * nco + 4 * oq;
                float4* bl = b_lo + kg * nco + 4 * oq;
                split4(c0, h, l); bh[0] = h; bl[0] = l;
                split4(c1, h, l); bh[1] = h; bl[1] = l;
                split4(c2, h, l); bh[2] = h; bl[2] = l;
                split4(c3, h, l); bh[3] = h; bl[3] = l;
            }
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            issue_3xtf32(taddr, smem_u32(a_hi), smem_u32(a_lo), 128, smem_u32(b_hi), smem_u32(b_lo), nco,
                         TKR / 8, idesc, it > 0);
            umma_commit(&mbar[s_]);
        }
    }
    // ---- epilogue: TMEM -> split-K partial -----------------------------------------------------------------
    const bool any = it > 0;
    if (any) mbar_wait(&mbar[(it - 1) & 1], ((it - 1) >> 1) & 1);
    tc_fence_after();
    {
        const int row = 32 * (warp & 3) + lane;
        const int half = warp >> 2, ncol = nco / 2;
        const size_t psz = (size_t)(p.fin + H) * M * H3;
        float* part = p.part + (size_t)blockIdx.y * psz;
        for (int cb = half * ncol; cb < (half + 1) * ncol; cb += 32) {
            float v[32];
            if (any) tmem_ld32(taddr + ((uint32_t)(32 * (warp & 3)) << 16) + cb, v);
            else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
            if (row < nkk) {
                float4* dst = reinterpret_cast<float4*>(part + (size_t)(kk0 + row) * H3 + job.o0 + cb);
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            if (ones_row && row == 127) {
                float* pb = p.partb + (size_t)blockIdx.y * H3 + job.o0 + cb;
#pragma unroll
                for (int j = 0; j < 32; ++j) pb[j] = v[j];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(taddr);
}

int dw_tc_smem_bytes(int M, int nco_max) { return dwtc_smem(M, nco_max).total; }

cudaError_t launch_dw_tc(const DwParams& p, int njobs, int nco_max, cudaStream_t st) {
    int smem = dw_tc_smem_bytes(p.M, nco_max);
    cudaError_t e = cudaFuncSetAttribute(dw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    dim3 grid(njobs, p.nsplit);
    dw_tc_kernel<<<grid, NT, smem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace dcgru
