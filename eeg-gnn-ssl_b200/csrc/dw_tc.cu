// Tensor-core (tcgen05, 3xTF32) weight-gradient kernel:  dW[kk][o] = sum_rows G[row][kk] * dA[row][o].
//
// GEMM view: M = kk (a 128-row slab of W rows per CTA), N = o (<= 192 output columns), K = rows =
// every (step, sample, node) -- 583 680 of them at BASELINE config 2 -- split over CTAs.
//   A[m = kk][k = row]  = diffuse(Z)^T : produced on the fly by the CTA's threads from the saved
//                         activations (x, h_prev, r*h_prev) and the per-sample polynomials P_m
//   B[n = o][k = row]   = dA^T         : loaded row-major from HBM, transposed in registers
// Both are written hi/lo-split straight into K-group-major UMMA tiles (tc_common.cuh), 40 K-rows
// (= 2 sample steps x 20 padded nodes = 5 MMA k-steps) per chunk, double buffered: while the tensor
// core consumes chunk i the 512 threads produce chunk i+1; the global loads of chunk i+1 are issued
// into registers before the shared-memory work of chunk i so their latency is hidden.
// The accumulator (128 x N fp32) lives in TMEM; because the tensor core truncates when it adds into
// the fp32 accumulator, it is flushed into the CTA's split-K partial every FLUSH chunks so that the
// accumulation bias stays far below the 1e-4 parity budget.  dw.cu's reduce_cell_kernel sums the
// partials in fixed order.  The bias gradient rides along as an extra all-ones A row in a slab that
// has a spare row (db[o] = sum_rows 1 * dA[row][o]).
#include "common.cuh"
#include "dw.cuh"
#include "tc_common.cuh"

namespace dcgru {
using namespace tc;

constexpr int TNT = 512;                // threads
constexpr int TKR = 40;                 // K rows per chunk
constexpr int TKG = TKR / 4;            // K groups per chunk
constexpr int TFLUSH = 16;              // chunks between TMEM -> partial flushes (640 K rows)
constexpr int PTS = NP * NP;            // one padded transposed polynomial

struct DwTcSmem {
    int a_bytes, b_bytes, stage_bytes, zc_off, pt_off, total, zld;
};
__host__ __device__ inline DwTcSmem dwtc_smem(int M, int nco) {
    DwTcSmem s;
    s.a_bytes = TKG * 128 * 16;                     // one of hi / lo
    s.b_bytes = TKG * nco * 16;
    s.stage_bytes = 2 * s.a_bytes + 2 * s.b_bytes;
    s.zld = ((128 / M + 2) + 3) & ~3;               // source columns a 128-row slab can touch
    s.zc_off = 2 * s.stage_bytes;
    s.pt_off = s.zc_off + 2 * NP * s.zld * 4;
    s.total = s.pt_off + 2 * (M - 1) * PTS * 4;
    return s;
}

// PT[b][m1][j][n] = P[b][m1][n][j], zero padded to NP x NP (contiguous per sample -> float4 copies)
__global__ void make_pt_kernel(const float* P, int B, int M1, int N, float* PT) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)B * M1 * PTS;
    if (i >= total) return;
    int n = (int)(i % NP), j = (int)((i / NP) % NP);
    size_t bm = i / PTS;
    PT[i] = (n < N && j < N) ? P[(bm * N + n) * N + j] : 0.f;
}

struct SrcOne {             // where one sample step of a chunk lives (all scalars: stays in registers)
    const float* zsrc;      // base pointer of the Z source rows (nullptr = zeros)
    const float* rsrc;      // r gate rows (type 2)
    const float* dA;        // dA rows
    const float* pt;        // padded P^T of the sample
    bool valid;
};

// sample step q = t*B + b (32-bit arithmetic: T*B < 2^31)
__device__ __forceinline__ SrcOne sample_source(const DwParams& p, const DwJob& job, unsigned q) {
    const int N = p.N, H = p.H, B = p.B;
    const size_t NH = (size_t)N * H;
    SrcOne o;
    o.valid = q < (unsigned)p.T * (unsigned)B;
    const unsigned tq = o.valid ? q / (unsigned)B : 0u;
    const int t = (int)tq, b = o.valid ? (int)(q - tq * (unsigned)B) : 0;
    o.pt = p.dY + (size_t)b * (p.M - 1) * PTS;               // dY field carries the PT buffer for this kernel
    const size_t cs = (p.mode == 0) ? (size_t)t * B * NH * 3 : ((size_t)t * p.ncell + p.layer) * B * NH * 3;
    o.dA = p.dA + cs + (size_t)b * N * 3 * H;
    o.rsrc = p.ruc + cs + (size_t)b * N * 3 * H;
    const float* z = nullptr;
    if (job.type == 0) {
        if (p.mode == 0) z = p.x + (size_t)t * p.xs_t + (size_t)b * p.xs_b;
        else if (p.layer == 0) {
            const size_t nfo = (size_t)N * p.Fo;
            if (t > 0) z = (((p.teacher_mask >> (t - 1)) & 1ull) ? p.targets : p.out) + ((size_t)(t - 1) * B + b) * nfo;
        } else z = p.hseq + (((size_t)t * p.ncell + (p.layer - 1)) * B + b) * NH;
    } else {
        if (p.mode == 0) z = (t == 0) ? p.h0 + (size_t)b * NH : p.hseq + ((size_t)(t - 1) * B + b) * NH;
        else z = (t == 0) ? p.h0 + ((size_t)p.layer * B + b) * NH
                          : p.hseq + (((size_t)(t - 1) * p.ncell + p.layer) * B + b) * NH;
    }
    o.zsrc = o.valid ? z : nullptr;
    return o;
}

__global__ void __launch_bounds__(TNT, 1) dw_tc_kernel(const DwParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t mbar[2];
    __shared__ uint32_t tmem_slot;
    const DwJob job = p.jobs[blockIdx.x];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = p.N, H = p.H, M = p.M, M1 = M - 1, H3 = 3 * H;
    const int nco = job.nco, kk0 = job.kk0, nkk = job.nz;          // nz field = rows of this slab
    const DwTcSmem L = dwtc_smem(M, nco);
    const int ZLD = L.zld;
    float* Zc = reinterpret_cast<float*>(smem + L.zc_off);          // [2][NP][ZLD]
    float* PT = reinterpret_cast<float*>(smem + L.pt_off);          // [2][M1][NP(j)][NP(n)]
    const int c_lo = kk0 / M, c_hi = (kk0 + nkk - 1) / M;           // absolute columns of [x | h]
    const int ncz = (nkk > 0) ? (c_hi - c_lo + 1) : 0;
    const int zoff = (job.type == 0) ? 0 : p.fin;
    const bool ones_row = (job.z0 != 0);                             // z0 field = "carry the db row" flag

    if (warp == 0) tmem_alloc<256>(&tmem_slot);
    if (tid == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); mbar_fence_init(); }
    for (int idx = tid; idx < 2 * L.stage_bytes / 16; idx += TNT)
        reinterpret_cast<float4*>(smem)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = tmem_slot;
    const uint32_t idesc = make_idesc_tf32(128, nco);

    // ---- per-thread task table (identical for every chunk) ------------------------------------------------
    // A: (sample, source column, node quad)
    const int a_cnt = 2 * ncz * 5;
    int a_s = 0, a_q5 = 0, a_ccl = 0;
    if (tid < a_cnt) { a_ccl = tid % ncz; int t1 = tid / ncz; a_q5 = t1 % 5; a_s = t1 / 5; }
    // B: (K group, output column): 4 rows of one dA column -> one float4 of the tile; consecutive lanes own
    // consecutive columns, so both the global loads and the shared stores are contiguous
    constexpr int BS = 4;                                             // TKG*192/512 rounded up
    int b_off[BS], b_src[BS];                                         // tile slot kg*nco + o (or -1) / dA offset, s in bit 30
    unsigned b_nmask = 0;                                             // 4 node-valid bits per slot
#pragma unroll
    for (int k = 0; k < BS; ++k) {
        int idx = tid + k * TNT;
        b_off[k] = -1; b_src[k] = 0;
        if (idx < TKG * nco) {
            const int kg = idx / nco, o = idx - kg * nco;
            const int s = kg >= 5, q5 = kg - 5 * s;
            b_off[k] = idx;
            b_src[k] = (s << 30) | ((4 * q5) * H3 + job.o0 + o);
            for (int i = 0; i < 4; ++i)
                if (4 * q5 + i < N) b_nmask |= 1u << (4 * k + i);
        }
    }
    // Z: up to 4 source elements per thread
    constexpr int ZS = 4;
    int z_sm[ZS], z_src[ZS], z_r[ZS];                                 // smem offset / (node*stride + col), s in bit 30 / r-gate offset
    const long long zstride = (job.type == 0) ? p.fin : H;
#pragma unroll
    for (int k = 0; k < ZS; ++k) {
        int idx = tid + k * TNT;
        z_sm[k] = -1; z_src[k] = 0; z_r[k] = 0;
        if (idx < 2 * NP * ncz) {
            int ccl = idx % ncz, j = (idx / ncz) % NP, s = idx / (ncz * NP);
            z_sm[k] = (s * NP + j) * ZLD + ccl;
            z_src[k] = (j < N) ? ((s << 30) | (int)(j * zstride + (c_lo + ccl - zoff))) : -1;
            z_r[k] = j * 3 * H + (c_lo + ccl - zoff);
        }
    }
    // P^T: float4 slots
    constexpr int PS = 3;
    const int pt_f4 = 2 * M1 * PTS / 4;

    float4 breg[BS];
    float zreg[ZS], rreg[ZS];
    float4 preg[PS];
    bool cur_v0 = false, cur_v1 = false;       // validity of the two sample steps held in the registers
    const int pt_half = M1 * PTS / 4;          // float4s of one sample's P^T

    auto prefetch = [&](long ch) {
        const SrcOne s0 = sample_source(p, job, (unsigned)(2 * ch));
        const SrcOne s1 = sample_source(p, job, (unsigned)(2 * ch + 1));
        cur_v0 = s0.valid; cur_v1 = s1.valid;
        // dA rows of this thread's B task
#pragma unroll
        for (int k = 0; k < BS; ++k) {
            breg[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (b_off[k] >= 0) {
                const int s = (b_src[k] >> 30) & 1;
                if (s ? s1.valid : s0.valid) {
                    const float* da = (s ? s1.dA : s0.dA) + (b_src[k] & 0x3FFFFFFF);
                    const unsigned nm = b_nmask >> (4 * k);
                    if (nm & 1u) breg[k].x = da[0];
                    if (nm & 2u) breg[k].y = da[H3];
                    if (nm & 4u) breg[k].z = da[2 * H3];
                    if (nm & 8u) breg[k].w = da[3 * H3];
                }
            }
        }
#pragma unroll
        for (int k = 0; k < ZS; ++k) {
            float v = 0.f, rr = 1.f;
            if (z_sm[k] >= 0 && z_src[k] >= 0) {
                const int s = (z_src[k] >> 30) & 1, off = z_src[k] & 0x3FFFFFFF;
                const float* zs = s ? s1.zsrc : s0.zsrc;
                if (zs != nullptr) {
                    v = zs[off];
                    if (job.type == 2) rr = (s ? s1.rsrc : s0.rsrc)[z_r[k]];
                }
            }
            zreg[k] = v;
            rreg[k] = rr;                                             // multiplied when stored: keeps the loads in flight
        }
#pragma unroll
        for (int k = 0; k < PS; ++k) {
            int idx = tid + k * TNT;
            preg[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < pt_f4) {
                const int s = idx >= pt_half;
                const int r = idx - s * pt_half;
                if (s ? s1.valid : s0.valid) preg[k] = reinterpret_cast<const float4*>(s ? s1.pt : s0.pt)[r];
            }
        }
    };

    const size_t psz = (size_t)(p.fin + H) * M * H3;
    float* part = p.part + (size_t)blockIdx.y * psz;
    bool flushed = false;
    // TMEM -> partial (+=).  Called by all threads after the MMAs issued so far have completed.
    auto flush = [&]() {
        tc_fence_after();
        const int row = 32 * (warp & 3) + lane;
        const int cg = warp >> 2, ncol = nco / 4;
        for (int cb = cg * ncol; cb < (cg + 1) * ncol; cb += 16) {
            float v[16];
            tmem_ld16(taddr + ((uint32_t)(32 * (warp & 3)) << 16) + cb, v);
            if (row < nkk) {
                float4* dst = reinterpret_cast<float4*>(part + (size_t)(kk0 + row) * H3 + job.o0 + cb);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    if (flushed) { float4 q = dst[j]; o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w; }
                    dst[j] = o;
                }
            }
            if (ones_row && row == 127) {
                float* pb = p.partb + (size_t)blockIdx.y * H3 + job.o0 + cb;
#pragma unroll
                for (int j = 0; j < 16; ++j) pb[j] = (flushed ? pb[j] : 0.f) + v[j];
            }
        }
        flushed = true;
        tc_fence_before();
        __syncthreads();
    };

    const long nchunk = ((long)p.T * p.B + 1) / 2;
    int it = 0;              // chunks done by this CTA
    int since_flush = 0;     // chunks accumulated in TMEM since the last flush
    if ((long)blockIdx.y < nchunk) prefetch(blockIdx.y);
    for (long ch = blockIdx.y; ch < nchunk; ch += gridDim.y, ++it) {
        const int s_ = it & 1;
        if (it >= 2) mbar_wait(&mbar[s_], ((it >> 1) - 1) & 1);
        uint8_t* st = smem + s_ * L.stage_bytes;
        float4* a_hi = reinterpret_cast<float4*>(st);
        float4* a_lo = reinterpret_cast<float4*>(st + L.a_bytes);
        float4* b_hi = reinterpret_cast<float4*>(st + 2 * L.a_bytes);
        float4* b_lo = reinterpret_cast<float4*>(st + 2 * L.a_bytes + L.b_bytes);
        const bool v0 = cur_v0, v1 = cur_v1;
        // ---- registers of this chunk -> shared memory --------------------------------------------------------
#pragma unroll
        for (int k = 0; k < ZS; ++k)
            if (z_sm[k] >= 0) Zc[z_sm[k]] = zreg[k] * rreg[k];
#pragma unroll
        for (int k = 0; k < PS; ++k) {
            int idx = tid + k * TNT;
            if (idx < pt_f4) reinterpret_cast<float4*>(PT)[idx] = preg[k];
        }
#pragma unroll
        for (int k = 0; k < BS; ++k) {                                // B tile: dA^T
            if (b_off[k] >= 0) {
                float4 h, l;
                split4(breg[k], h, l);
                b_hi[b_off[k]] = h;
                b_lo[b_off[k]] = l;
            }
        }
        if (ones_row && tid < TKG) {                                 // db row: 1 for every real (sample, node)
            int s = tid / 5, q5 = tid % 5;
            bool v = s ? v1 : v0;
            float4 o;
            o.x = (v && 4 * q5 + 0 < N) ? 1.f : 0.f;
            o.y = (v && 4 * q5 + 1 < N) ? 1.f : 0.f;
            o.z = (v && 4 * q5 + 2 < N) ? 1.f : 0.f;
            o.w = (v && 4 * q5 + 3 < N) ? 1.f : 0.f;
            a_hi[tid * 128 + 127] = o;
        }
        __syncthreads();
        // ---- global loads of the next chunk (latency hidden behind the A production below) ---------------------
        if (ch + gridDim.y < nchunk) prefetch(ch + gridDim.y);
        // ---- A tile: G^T ------------------------------------------------------------------------------------------
        if (tid < a_cnt) {
            const float* zp = Zc + (a_s * NP) * ZLD + a_ccl;
            const int kg = a_s * 5 + a_q5;
            const int kbase = (c_lo + a_ccl) * M - kk0;              // local row of the m = 0 term
            if (kbase >= 0 && kbase < nkk) {
                float4 v = make_float4(zp[(4 * a_q5) * ZLD], zp[(4 * a_q5 + 1) * ZLD], zp[(4 * a_q5 + 2) * ZLD],
                                       zp[(4 * a_q5 + 3) * ZLD]);     // (dynamic quad index: read from smem, not zc[])
                float4 h, l;
                split4(v, h, l);
                a_hi[kg * 128 + kbase] = h;
                a_lo[kg * 128 + kbase] = l;
            }
            float zc[NP];                                             // the source column (rows >= N are zero)
#pragma unroll
            for (int j = 0; j < NP; ++j) zc[j] = zp[j * ZLD];
            for (int m1 = 0; m1 < M1; ++m1) {
                const int kl = kbase + m1 + 1;
                if (kl < 0 || kl >= nkk) continue;
                const float* pp = PT + (size_t)(a_s * M1 + m1) * PTS + 4 * a_q5;
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int j = 0; j < NP; ++j) {                        // PT rows j >= N are zero
                    float4 pv = *reinterpret_cast<const float4*>(pp + j * NP);
                    a.x = fmaf(pv.x, zc[j], a.x); a.y = fmaf(pv.y, zc[j], a.y);
                    a.z = fmaf(pv.z, zc[j], a.z); a.w = fmaf(pv.w, zc[j], a.w);
                }
                float4 h, l;
                split4(a, h, l);
                a_hi[kg * 128 + kl] = h;
                a_lo[kg * 128 + kl] = l;
            }
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            issue_3xtf32(taddr, smem_u32(a_hi), smem_u32(a_lo), 128, smem_u32(b_hi), smem_u32(b_lo), nco,
                         TKR / 8, idesc, since_flush > 0);
            umma_commit(&mbar[s_]);
        }
        ++since_flush;
        if (since_flush == TFLUSH) {
            mbar_wait(&mbar[s_], (it >> 1) & 1);                      // this chunk's MMAs (hence all earlier) done
            flush();
            since_flush = 0;
        }
    }
    if (since_flush > 0) {
        mbar_wait(&mbar[(it - 1) & 1], ((it - 1) >> 1) & 1);
        flush();
    } else if (!flushed) {                                            // CTA had no chunk at all: zero partial
        const int row = tid & 127, q = tid >> 7;
        if (row < nkk)
            for (int c = q * (nco / 4); c < (q + 1) * (nco / 4); ++c) part[(size_t)(kk0 + row) * H3 + job.o0 + c] = 0.f;
        if (ones_row && tid < nco) p.partb[(size_t)blockIdx.y * H3 + job.o0 + tid] = 0.f;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(taddr);
}

int dw_tc_smem_bytes(int M, int nco_max) { return dwtc_smem(M, nco_max).total; }
size_t dw_tc_pt_floats(int B, int M) { return (size_t)B * (M - 1) * PTS; }

cudaError_t launch_make_pt(const float* P, int B, int M, int N, float* PT, cudaStream_t st) {
    size_t total = (size_t)B * (M - 1) * PTS;
    if (total == 0) return cudaSuccess;
    make_pt_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(P, B, M - 1, N, PT);
    return cudaGetLastError();
}

cudaError_t launch_dw_tc(const DwParams& p, int njobs, int nco_max, cudaStream_t st) {
    int smem = dw_tc_smem_bytes(p.M, nco_max);
    cudaError_t e = cudaFuncSetAttribute(dw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    dim3 grid(njobs, p.nsplit);
    dw_tc_kernel<<<grid, TNT, smem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace dcgru
