// Persistent encoder-layer backward (BPTT) on the tensor cores (tcgen05, 3xTF32) -- SURVEY A.4.
//
// One CTA = 4 samples (M = 128 rows = 4 x 32, one warp per sample: see tc_common.cuh) for all T steps, time running backwards.
// Per step two UMMA GEMMs whose A operand is the elementwise gate gradient (no diffusion on the input
// side) and whose 192 output columns are the M = 3 diffusion terms of the 64 hidden columns; the
// transposed diffusion sum_m P_m^T is applied on the output side by the producer warps:
//     E1   du, dc, dH*u, dA_c, dA_u                                  (thread = (row, column half))
//     B1   D1[:,0:192] = dA_c (K=64)          @ Wc_h^T   -> diffT -> d(rH)
//     E2   dA_r = d(rH) * h * r(1-r),  dH += d(rH) * r
//     B2   D2[:,0:192] = [dA_u | dA_r] (K=128) @ Wg_h^T  -> diffT -> dH +=
// (the u half of B2 is issued right after B1, so it overlaps the diffT of B1).  dA = [dA_r|dA_u|dA_c] is
// stored for the bulk weight-gradient kernel and for the input-gradient kernel (dX is not part of the
// recurrence, so it is computed in bulk afterwards).
//
// Roles: warps 0-7 elementwise + A tiles + diffT, warp 8 MMA issue, warp 9 TMA weight loads (12 chunk blocks
// per step, pre-split / pre-tiled by pack_w_bwd_kernel, 3-slot ring).  Synchronisation is mbarrier-only
// apart from two 128-thread named barriers around the diffT staging planes.
// A slots are laid out [row group of 8][K-group pair][8 rows x 32 B] (K-major, 32-byte swizzle, SBO = 512 B).  Optional operand
// image (daimg): warp 10 copies every finished A slot (hi and lo) to HBM with one tensor-map TMA store each into
// the row-major image DA[cta*T + t][hi|lo][96 rows = sample*24 + node][192 columns r|u|c] -- the B operand of the weight-gradient
// GEMM (dw_mm.cu).
#include <cstring>

#include "common.cuh"
#include "dw.cuh"
#include "tc_common.cuh"
#include "tmap.cuh"

namespace dcgru {
using namespace tc;

constexpr int BT_SB = TC_SB;
constexpr int BT_RP = TC_RP;
constexpr int BT_ROWS = 128;
constexpr int BT_H = 64;
constexpr int BT_M = 3;
constexpr int BT_NCOL = BT_H * BT_M;                       // 192 output columns (kk = c*M + m)
constexpr int BT_CK = 16;                                  // K (= o) per chunk: 2 MMA k-steps
constexpr int BT_KG = BT_CK / 4;
constexpr int BT_A_BYTES = BT_KG * BT_ROWS * 16;           // one of hi / lo
constexpr int BT_A_SLOT = 2 * BT_A_BYTES;                  // 16 KB
constexpr int BT_B_SLOT = 2 * BT_KG * BT_NCOL * 16;        // 24 KB
constexpr int BT_CHUNKS = 12;                              // per step: 4 (B1) + 4 (B2 u) + 4 (B2 r)
constexpr int BT_PLD = 36;                                 // staging plane / IO tile row stride (floats)
constexpr int BT_DLD = BT_H + 4;                           // dH row stride
constexpr int BT_OFF_A = 0;                                // 4 slots
constexpr int BT_OFF_B = BT_OFF_A + 4 * BT_A_SLOT;         // 3 slots
constexpr int BT_OFF_S = BT_OFF_B + 3 * BT_B_SLOT;         // 2 planes [128][36] (also the IO tiles 0)
constexpr int BT_OFF_DH = BT_OFF_S + 2 * BT_ROWS * BT_PLD * 4;
constexpr int BT_SMEM = BT_OFF_DH + BT_ROWS * BT_DLD * 4 + 1024;   // + slack: the swizzled slots need an aligned base
constexpr int BT_NPROD = 256;
constexpr int BT_THREADS = 352;                            // + warp 10: operand-image dump
constexpr int BT_RG_F4 = BT_KG * 8;                        // float4s per 8-row group of an A slot (512 B)
constexpr uint32_t BT_D1 = 0, BT_D2 = 256;                 // TMEM column bases

// weight image: 12 blocks in consumption order, each [hi: kg][n = kk] [lo: kg][n] float4 over 4 consecutive o
//   0-3 : Wc rows of the h part, o = 16i ..      (B1)
//   4-7 : Wg rows of the h part, o = 64 + 16i .. (B2, u columns)
//   8-11: Wg rows of the h part, o = 16i ..      (B2, r columns)
__global__ void pack_w_bwd_kernel(const float* Wg, const float* Wc, int fin, float* img) {
    const int blk = blockIdx.x;
    float4* hi = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(img) + (size_t)blk * BT_B_SLOT);
    float4* lo = hi + BT_KG * BT_NCOL;
    for (int idx = threadIdx.x; idx < BT_KG * BT_NCOL; idx += blockDim.x) {
        const int kg = idx / BT_NCOL, kk = idx - kg * BT_NCOL;
        const size_t row = (size_t)fin * BT_M + kk;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (blk < 4) v[e] = Wc[row * BT_H + 16 * blk + 4 * kg + e];
            else if (blk < 8) v[e] = Wg[row * 2 * BT_H + BT_H + 16 * (blk - 4) + 4 * kg + e];
            else v[e] = Wg[row * 2 * BT_H + 16 * (blk - 8) + 4 * kg + e];
        }
        float4 h, l;
        split4(make_float4(v[0], v[1], v[2], v[3]), h, l);
        hi[idx] = h;
        lo[idx] = l;
    }
}

// weight image of the input-gradient GEMM (fin == 64): 12 blocks, o = 16*blk .. over [Wg_x | Wc_x] columns
__global__ void pack_w_dx_kernel(const float* Wg, const float* Wc, float* img) {
    const int blk = blockIdx.x;
    float4* hi = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(img) + (size_t)blk * BT_B_SLOT);
    float4* lo = hi + BT_KG * BT_NCOL;
    for (int idx = threadIdx.x; idx < BT_KG * BT_NCOL; idx += blockDim.x) {
        const int kg = idx / BT_NCOL, kk = idx - kg * BT_NCOL;               // kk: x-part rows 0..191 (fin = 64)
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int o = 16 * blk + 4 * kg + e;
            v[e] = (o < 2 * BT_H) ? Wg[(size_t)kk * 2 * BT_H + o] : Wc[(size_t)kk * BT_H + o - 2 * BT_H];
        }
        float4 h, l;
        split4(make_float4(v[0], v[1], v[2], v[3]), h, l);
        hi[idx] = h;
        lo[idx] = l;
    }
}

struct BwdTcParams {
    int B, T, N, act;
    int mode;                   // 0: BPTT (B1, B2), 1: bulk input gradient dX[t] = diffT(dA[t] @ [Wg_x|Wc_x]^T), fin == 64
    float* dx;                  // mode 1: (T, B, N*64)
    const float* h0;
    const float* hseq;
    const float* ruc;
    const float* P;
    const float* d_hseq;
    const float* d_hlast;
    const float* wimg;
    float* dh0;
    float* dA;
    uint8_t* daimg;             // mode 0: operand image of dA for dw_mm (nullptr: not saved)
};

__device__ __forceinline__ void bt_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bt_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bt_bulk(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void prod_barrier() {             // the 8 producer warps (the IO tiles alias the planes and
    asm volatile("bar.sync 1, 256;\n" ::: "memory");         // the A ring across column halves, so halves cannot sync alone)
}
// 32 lanes x 16 columns registers -> TMEM is not needed here; TMEM is only read (tmem_ld16)

__global__ void __launch_bounds__(BT_THREADS, 1) seq_bwd_tc_kernel(const BwdTcParams p,
                                                                   const __grid_constant__ CUtensorMap tm_d) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar_bfull[3], bar_afull[4], bar_cdone[4], bar_stored[4], bar_d1free, bar_d2free, bar_b1done;
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t dA_desc[4][2][2];                    // [slot][k-step][hi, lo]
    __shared__ uint64_t dB_desc[3][2][2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = p.N;
    const int b0 = blockIdx.x * BT_SB;
    float* DH = reinterpret_cast<float*>(smem + BT_OFF_DH);

    if (warp == 0) tmem_alloc<512>(&tmem_slot);
    if (tid == 0) {
        for (int i = 0; i < 3; ++i) mbar_init(&bar_bfull[i], 1);
        for (int i = 0; i < 4; ++i) { mbar_init(&bar_afull[i], BT_NPROD / 32); mbar_init(&bar_cdone[i], 1); mbar_init(&bar_stored[i], 1); }
        mbar_init(&bar_d1free, BT_NPROD / 32);
        mbar_init(&bar_d2free, BT_NPROD / 32);
        mbar_init(&bar_b1done, 1);
        mbar_fence_init();
        for (int sl = 0; sl < 4; ++sl)
            for (int k = 0; k < 2; ++k) {
                const uint32_t hi = smem_u32(smem + BT_OFF_A + sl * BT_A_SLOT) + k * 256;
                dA_desc[sl][k][0] = make_smem_desc_k32(hi, BT_RG_F4 * 16);
                dA_desc[sl][k][1] = make_smem_desc_k32(hi + BT_A_BYTES, BT_RG_F4 * 16);
            }
        for (int sl = 0; sl < 3; ++sl)
            for (int k = 0; k < 2; ++k) {
                const uint32_t hi = smem_u32(smem + BT_OFF_B + sl * BT_B_SLOT) + 2 * k * BT_NCOL * 16;
                dB_desc[sl][k][0] = make_smem_desc(hi, BT_NCOL * 16, 128);
                dB_desc[sl][k][1] = make_smem_desc(hi + BT_KG * BT_NCOL * 16, BT_NCOL * 16, 128);
            }
    }
    for (int idx = tid; idx < BT_ROWS * BT_DLD; idx += BT_THREADS) DH[idx] = 0.f;
    for (int idx = tid; idx < 4 * BT_A_SLOT / 16; idx += BT_THREADS)
        reinterpret_cast<float4*>(smem + BT_OFF_A)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = tmem_slot;
    const unsigned total_chunks = (unsigned)p.T * BT_CHUNKS;
    const uint8_t* wimg = reinterpret_cast<const uint8_t*>(p.wimg);
    const size_t NH = (size_t)N * BT_H;
    const bool dump = (p.daimg != nullptr) && p.mode == 0;

    if (warp == 8) {
        // =================================== MMA issuer =========================================================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(128, BT_NCOL);
            int q = 0, step = 0, sb = 0, kb = 0;
            for (unsigned g = 0; g < total_chunks; ++g) {
                const int sa = g & 3;
                // one combined poll (weights landed, A slot built): see seq_fwd_tc.cu
                mbar_wait2(&bar_bfull[sb], kb, &bar_afull[sa], (g >> 2) & 1);
                if (step > 0) {                                           // the previous step's reads of D are done
                    if (q == 0) mbar_wait(&bar_d1free, (step - 1) & 1);
                    if (q == 4 && p.mode == 0) mbar_wait(&bar_d2free, (step - 1) & 1);
                }
                tc_fence_after();
                const uint32_t d = taddr + ((q < 4 || p.mode == 1) ? BT_D1 : BT_D2);
                uint32_t acc = (q == 0 || (q == 4 && p.mode == 0)) ? 0u : 1u;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const uint64_t ah = dA_desc[sa][k][0], al = dA_desc[sa][k][1];
                    const uint64_t bh = dB_desc[sb][k][0], bl = dB_desc[sb][k][1];
                    umma_tf32(d, al, bh, idesc, acc);
                    umma_tf32(d, ah, bl, idesc, 1u);
                    umma_tf32(d, ah, bh, idesc, 1u);
                    acc = 1u;
                }
                umma_commit(&bar_cdone[sa]);
                // B1 complete: its own barrier, one phase per step.  (bar_cdone of B1's last slot cannot be used for
                // this: the producers refill that slot with a B2 chunk before they wait, and if that chunk's MMAs also
                // finish before a late warp polls, the slot barrier has advanced two phases and the poll never succeeds.)
                if (q == 3 && p.mode == 0) umma_commit(&bar_b1done);
                if (++q == BT_CHUNKS) { q = 0; ++step; }
                if (++sb == 3) { sb = 0; kb ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == 9) {
        // =================================== TMA weight loader ==================================================
        if (lane == 0) {
            auto load_w = [&](int q, int slot) {
                bt_expect_tx(&bar_bfull[slot], BT_B_SLOT);
                bt_bulk(smem + BT_OFF_B + slot * BT_B_SLOT, wimg + (size_t)q * BT_B_SLOT, BT_B_SLOT, &bar_bfull[slot]);
            };
            load_w(0, 0);
            load_w(1, 1);
            int q2 = 2, sb2 = 2;
            for (unsigned g = 0; g + 2 < total_chunks; ++g) {
                if (g >= 1) mbar_wait(&bar_cdone[(g - 1) & 3], ((g - 1) >> 2) & 1);   // frees weight slot sb2
                load_w(q2, sb2);
                if (++q2 == BT_CHUNKS) q2 = 0;
                if (++sb2 == 3) sb2 = 0;
            }
        }
        __syncwarp();
    } else if (warp == 10) {
        // =================================== operand-image dump ==================================================
        // lane = (hi|lo, sample): one TMA tensor store per finished A slot part and sample: box (8 o, 8 rows, 2 column
        // octets, 3 row groups) = the 24 rows of the sample that can be non-zero
        if (dump) {
            const int part = lane >> 2, s = lane & 3;
            if (lane == 0) tma_prefetch_desc(&tm_d);
            int q = 0, t = p.T - 1;
            for (unsigned g = 0; g < total_chunks; ++g) {
                const int sa = g & 3;
                // column quad of the chunk in [r | u | c] order: B1 = c, then u, then r
                const int og0 = (q < 4) ? 32 + 4 * q : (q < 8 ? 16 + 4 * (q - 4) : 4 * (q - 8));
                const int rg0 = (((blockIdx.x * p.T + t) * 2 + part) * BT_SB + s) * TC_RG;
                mbar_wait(&bar_afull[sa], (g >> 2) & 1);
                if (lane < 2 * BT_SB) {
                    const uint8_t* src = smem + BT_OFF_A + sa * BT_A_SLOT + part * BT_A_BYTES + s * (4 * BT_RG_F4 * 16);
                    tma_store_4d(&tm_d, 0, 0, og0 / 2, rg0, src);
                    bulk_commit();
                }
                bulk_wait_read();
                __syncwarp();
                if (lane == 0) bt_arrive(&bar_stored[sa]);
                if (++q == BT_CHUNKS) { q = 0; --t; }
            }
            bulk_wait_all();
        }
        __syncwarp();
    } else {
        // =================================== producers ===========================================================
        const int row = tid & 127, hf = tid >> 7;
        const int s_ = row / BT_RP, n_ = row - s_ * BT_RP;
        const int b_ = b0 + s_;
        const bool rvalid = (n_ < N) && (b_ < p.B);
        // column j = n_ of the polynomials (= row n_ of P^T), kept in registers for the whole sequence
        float PT1[NP], PT2[NP];
#pragma unroll
        for (int n = 0; n < NP; ++n) {
            PT1[n] = 0.f; PT2[n] = 0.f;
            if (rvalid && n < N) {
                PT1[n] = p.P[(((size_t)b_ * 2 + 0) * N + n) * N + n_];
                PT2[n] = p.P[(((size_t)b_ * 2 + 1) * N + n) * N + n_];
            }
        }
        // ---- warp-private IO tiles: tile 0 in the staging planes, tile 1 in the A ring (both idle when used) ----
        float* tile0 = reinterpret_cast<float*>(smem + BT_OFF_S) + warp * (32 * BT_PLD);
        float* tile1 = reinterpret_cast<float*>(smem + BT_OFF_A) + warp * (32 * BT_PLD);
        const int rq = lane >> 3, f4 = lane & 7;
        size_t grow[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = 32 * (warp & 3) + rq + 4 * i;
            const int s = r / BT_RP, n = r - s * BT_RP, b = b0 + s;
            grow[i] = (n < N && b < p.B) ? ((size_t)b * N + n) : ~(size_t)0;
        }
        auto tile_load = [&](float* tile, const float* base, int ld, int col0) {          // async: cp.async
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float* d = tile + (rq + 4 * i) * BT_PLD + 4 * f4;
                if (base != nullptr && grow[i] != ~(size_t)0) cp_async16(d, base + grow[i] * ld + col0 + 4 * f4);
                else *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto tile_get = [&](const float* tile, float (&v)[32]) {
            const float4* d = reinterpret_cast<const float4*>(tile + lane * BT_PLD);
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float4 q = d[j]; v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w; }
        };
        auto tile_store = [&](float* tile, const float (&v)[32], float* base, int ld, int col0) {
            __syncwarp();
            float4* d = reinterpret_cast<float4*>(tile + lane * BT_PLD);
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (grow[i] != ~(size_t)0)
                    *reinterpret_cast<float4*>(base + grow[i] * ld + col0 + 4 * f4) =
                        *reinterpret_cast<const float4*>(tile + (rq + 4 * i) * BT_PLD + 4 * f4);
            __syncwarp();
        };
        // one A chunk (16 o) of the running chunk counter g from 16 of this thread's 32 values (owner half only)
        unsigned g = 0;
        auto put_chunk = [&](const float (&v)[32], int owner_hf, int sub) {
            const int sa = g & 3;
            if (g >= 4)                                                   // the MMAs that read this slot are done, and so is its
                mbar_wait2(&bar_cdone[sa], ((g >> 2) - 1) & 1,            // copy to the operand image (one combined poll)
                           dump ? &bar_stored[sa] : &bar_cdone[sa], ((g >> 2) - 1) & 1);
            if (hf == owner_hf) {
                float4* a_hi = reinterpret_cast<float4*>(smem + BT_OFF_A + sa * BT_A_SLOT);
                float4* a_lo = reinterpret_cast<float4*>(smem + BT_OFF_A + sa * BT_A_SLOT + BT_A_BYTES);
#pragma unroll
                for (int kg = 0; kg < BT_KG; ++kg) {
                    float4 h, l;
                    split4(make_float4(v[16 * sub + 4 * kg], v[16 * sub + 4 * kg + 1], v[16 * sub + 4 * kg + 2],
                                       v[16 * sub + 4 * kg + 3]), h, l);
                    a_hi[k32_idx<BT_KG / 2>(kg, row)] = h;
                    a_lo[k32_idx<BT_KG / 2>(kg, row)] = l;
                }
                fence_async_smem();
            }
            __syncwarp();
            if (lane == 0) bt_arrive(&bar_afull[sa]);
            ++g;
        };
        auto wait_chunk = [&](unsigned gc) {                              // MMAs of chunk gc (and all earlier) done
            mbar_wait(&bar_cdone[gc & 3], (gc >> 2) & 1);
            tc_fence_after();
        };
        // transposed diffusion of D (TMEM, 192 columns) -> 32 own columns, added into out[]
        float* S1 = reinterpret_cast<float*>(smem + BT_OFF_S);
        float* S2 = S1 + BT_ROWS * BT_PLD;
        const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
        auto diff_t = [&](uint32_t dcol, float (&out)[32]) {
#pragma unroll
            for (int ps = 0; ps < 2; ++ps) {
                prod_barrier();                                            // planes free (previous readers done)
#pragma unroll
                for (int i3 = 0; i3 < 3; ++i3) {
                    float v[16];
                    tmem_ld16(taddr + lane_base + dcol + 96 * hf + 48 * ps + 16 * i3, v);
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int kl = 16 * i3 + e, cl = kl / 3, m = kl - 3 * cl;          // compile-time after unrolling
                        if (m == 0) out[16 * ps + cl] += v[e];                               // identity term: own row
                        else if (m == 1) S1[row * BT_PLD + 16 * hf + cl] = v[e];
                        else S2[row * BT_PLD + 16 * hf + cl] = v[e];
                    }
                }
                prod_barrier();                                            // planes written by every row
                if (n_ < NP) {                                            // (all lanes of a warp read the same plane rows: broadcast)
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                        const float* z1 = S1 + (s_ * BT_RP) * BT_PLD + 16 * hf + 4 * q4;
                        const float* z2 = S2 + (s_ * BT_RP) * BT_PLD + 16 * hf + 4 * q4;
#pragma unroll
                        for (int nb = 0; nb < NP; nb += 5) {
                            float4 u1[5], u2[5];
#pragma unroll
                            for (int n = 0; n < 5; ++n) {
                                u1[n] = *reinterpret_cast<const float4*>(z1 + (nb + n) * BT_PLD);
                                u2[n] = *reinterpret_cast<const float4*>(z2 + (nb + n) * BT_PLD);
                            }
#pragma unroll
                            for (int n = 0; n < 5; ++n) {
                                a0 = fmaf(PT1[nb + n], u1[n].x, a0); a1 = fmaf(PT1[nb + n], u1[n].y, a1);
                                a2 = fmaf(PT1[nb + n], u1[n].z, a2); a3 = fmaf(PT1[nb + n], u1[n].w, a3);
                                a0 = fmaf(PT2[nb + n], u2[n].x, a0); a1 = fmaf(PT2[nb + n], u2[n].y, a1);
                                a2 = fmaf(PT2[nb + n], u2[n].z, a2); a3 = fmaf(PT2[nb + n], u2[n].w, a3);
                            }
                        }
                        out[16 * ps + 4 * q4 + 0] += a0; out[16 * ps + 4 * q4 + 1] += a1;
                        out[16 * ps + 4 * q4 + 2] += a2; out[16 * ps + 4 * q4 + 3] += a3;
                    }
                }
            }
        };

        // initial carry: d_hlast
        {
            float v[32];
            tile_load(tile0, p.d_hlast, BT_H, 32 * hf);
            cp_async_commit();
            cp_async_wait<0>();
            __syncwarp();
            tile_get(tile0, v);
            if (rvalid) {
                float4* z4 = reinterpret_cast<float4*>(DH + row * BT_DLD + 32 * hf);
#pragma unroll
                for (int j = 0; j < 8; ++j) z4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            __syncwarp();
        }
        if (p.mode == 1) {
            // ---- bulk input gradient: one GEMM (K = 192: dA_r | dA_u | dA_c) + diffT per step, steps independent -----
            for (int t = 0; t < p.T; ++t) {
                const float* dA = p.dA + (size_t)t * p.B * NH * 3;
                const unsigned g0 = g;
                float w0[32];
#pragma unroll
                for (int part = 0; part < 3; ++part) {
                    tile_load(tile0, dA, 3 * BT_H, part * BT_H + 32 * hf);
                    cp_async_commit();
                    cp_async_wait<0>();
                    __syncwarp();
                    tile_get(tile0, w0);
                    __syncwarp();
                    put_chunk(w0, 0, 0); put_chunk(w0, 0, 1); put_chunk(w0, 1, 0); put_chunk(w0, 1, 1);
                }
                wait_chunk(g0 + 11);
#pragma unroll
                for (int j = 0; j < 32; ++j) w0[j] = 0.f;
                diff_t(BT_D1, w0);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) bt_arrive(&bar_d1free);
                prod_barrier();                                           // planes idle: tile 0 may be reused
                tile_store(tile0, w0, p.dx + (size_t)t * p.B * NH, BT_H, 32 * hf);
                prod_barrier();
            }
        } else
        for (int t = p.T - 1; t >= 0; --t) {
            const float* hprev = (t == 0) ? p.h0 : p.hseq + (size_t)(t - 1) * p.B * NH;
            const float* ruc = p.ruc + (size_t)t * p.B * NH * 3;
            float* dA = p.dA + (size_t)t * p.B * NH * 3;
            const unsigned g0 = g;                                        // first chunk of this step
            float hp[32], dAu[32], w0[32], w1[32];
            if (dump && g >= 4) {                                         // tile 1 aliases the A ring: its copies must have been read
#pragma unroll
                for (unsigned i = 1; i <= 4; ++i) mbar_wait(&bar_stored[(g - i) & 3], ((g - i) >> 2) & 1);
            }
            // ---- E1 ------------------------------------------------------------------------------------------
            // (all MMAs of the previous step are complete: the A ring and the planes are idle)
            tile_load(tile0, ruc, 3 * BT_H, BT_H + 32 * hf);              // u
            tile_load(tile1, ruc, 3 * BT_H, 2 * BT_H + 32 * hf);          // c
            cp_async_commit();
            cp_async_wait<0>();
            __syncwarp();
            tile_get(tile0, w0);                                          // u
            tile_get(tile1, w1);                                          // c
            __syncwarp();
            tile_load(tile0, hprev, BT_H, 32 * hf);
            tile_load(tile1, p.d_hseq ? p.d_hseq + (size_t)t * p.B * NH : nullptr, BT_H, 32 * hf);
            cp_async_commit();
            cp_async_wait<0>();
            __syncwarp();
            tile_get(tile0, hp);
            {
                float dup[32];
                tile_get(tile1, dup);
                prod_barrier();                                          // tile 1 lives in the A ring: all reads before any A write
                float4* z4 = reinterpret_cast<float4*>(DH + row * BT_DLD + 32 * hf);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 zz = z4[j];
                    const float dd[4] = {zz.x + dup[4 * j], zz.y + dup[4 * j + 1], zz.z + dup[4 * j + 2], zz.w + dup[4 * j + 3]};
                    float nd[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float u = w0[4 * j + e], cv = w1[4 * j + e], d = dd[e];
                        const float du = d * (hp[4 * j + e] - cv);
                        const float dc = d * (1.f - u);
                        nd[e] = d * u;
                        w1[4 * j + e] = dc * ((p.act == 0) ? (1.f - cv * cv) : (cv > 0.f ? 1.f : 0.f));   // dA_c
                        dAu[4 * j + e] = du * u * (1.f - u);
                    }
                    z4[j] = make_float4(nd[0], nd[1], nd[2], nd[3]);
                }
            }
            // ---- B1 operand: dA_c (4 chunks), then its global copy ---------------------------------------------
            put_chunk(w1, 0, 0); put_chunk(w1, 0, 1); put_chunk(w1, 1, 0); put_chunk(w1, 1, 1);
            tile_store(tile0, w1, dA, 3 * BT_H, 2 * BT_H + 32 * hf);
            tile_store(tile0, dAu, dA, 3 * BT_H, BT_H + 32 * hf);
            // ---- B2, u half: reuses the A slots of B1 as soon as B1's MMAs have consumed them -----------------------
            put_chunk(dAu, 0, 0); put_chunk(dAu, 0, 1); put_chunk(dAu, 1, 0); put_chunk(dAu, 1, 1);
            // ---- d(rH) = diffT(D1) -----------------------------------------------------------------------------------
            mbar_wait(&bar_b1done, (p.T - 1 - t) & 1);                  // all B1 MMAs of this step are complete
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < 32; ++j) w0[j] = 0.f;
            diff_t(BT_D1, w0);                                            // w0 = d(rH), own 32 columns
            tc_fence_before();
            __syncwarp();
            if (lane == 0) bt_arrive(&bar_d1free);
            // ---- E2 ------------------------------------------------------------------------------------------------
            prod_barrier();                                             // nobody still reads the planes (tile 0)
            tile_load(tile0, ruc, 3 * BT_H, 32 * hf);                     // r
            cp_async_commit();
            cp_async_wait<0>();
            __syncwarp();
            tile_get(tile0, w1);                                          // r
            __syncwarp();
            {
                float4* z4 = reinterpret_cast<float4*>(DH + row * BT_DLD + 32 * hf);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 zz = z4[j];
                    zz.x += w0[4 * j] * w1[4 * j]; zz.y += w0[4 * j + 1] * w1[4 * j + 1];
                    zz.z += w0[4 * j + 2] * w1[4 * j + 2]; zz.w += w0[4 * j + 3] * w1[4 * j + 3];
                    z4[j] = zz;
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) w0[j] = w0[j] * hp[j] * w1[j] * (1.f - w1[j]);               // dA_r
            }
            put_chunk(w0, 0, 0); put_chunk(w0, 0, 1); put_chunk(w0, 1, 0); put_chunk(w0, 1, 1);
            tile_store(tile0, w0, dA, 3 * BT_H, 32 * hf);
            // ---- dH += diffT(D2) ---------------------------------------------------------------------------------------
            wait_chunk(g0 + 11);
#pragma unroll
            for (int j = 0; j < 32; ++j) w1[j] = 0.f;
            diff_t(BT_D2, w1);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) bt_arrive(&bar_d2free);
            if (rvalid) {
                float4* z4 = reinterpret_cast<float4*>(DH + row * BT_DLD + 32 * hf);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 zz = z4[j];
                    zz.x += w1[4 * j]; zz.y += w1[4 * j + 1]; zz.z += w1[4 * j + 2]; zz.w += w1[4 * j + 3];
                    z4[j] = zz;
                }
            }
            prod_barrier();                                             // planes idle before the next step's IO tiles
        }
        // ---- dh0 ----------------------------------------------------------------------------------------------------
        if (p.mode == 0) {
            float v[32];
            const float4* z4 = reinterpret_cast<const float4*>(DH + row * BT_DLD + 32 * hf);
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float4 q = z4[j]; v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w; }
            tile_store(tile0, v, p.dh0, BT_H, 32 * hf);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(taddr);
}

size_t seq_bwd_tc_wimg_bytes() { return (size_t)BT_CHUNKS * BT_B_SLOT; }
size_t seq_bwd_tc_daimg_bytes(int B, int T) {
    return (size_t)((B + BT_SB - 1) / BT_SB) * T * 2 * TC_IMG_ROWS * 3 * BT_H * 4;
}
bool seq_bwd_tc_supported(int N, int H, int M, int smem_limit) {
    return H == BT_H && M == BT_M && N <= NP && BT_SMEM + 2304 <= smem_limit;
}

cudaError_t launch_seq_bwd_tc(int B, int T, int N, int fin, int act, const float* h0, const float* hseq,
                              const float* ruc, const float* P, const float* Wg, const float* Wc,
                              const float* d_hseq, const float* d_hlast, float* wimg, float* dh0, float* dA,
                              void* daimg, cudaStream_t st) {
    pack_w_bwd_kernel<<<BT_CHUNKS, 256, 0, st>>>(Wg, Wc, fin, wimg);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    BwdTcParams p;
    p.mode = 0; p.dx = nullptr;
    p.B = B; p.T = T; p.N = N; p.act = act; p.h0 = h0; p.hseq = hseq; p.ruc = ruc; p.P = P;
    p.d_hseq = d_hseq; p.d_hlast = d_hlast; p.wimg = wimg; p.dh0 = dh0; p.dA = dA;
    p.daimg = reinterpret_cast<uint8_t*>(daimg);
    e = cudaFuncSetAttribute(seq_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BT_SMEM);
    if (e != cudaSuccess) return e;
    CUtensorMap tm;
    memset(&tm, 0, sizeof tm);
    if (daimg) {
        const unsigned long long rowb = 3 * BT_H * 4;
        const unsigned long long dims[4] = {8, 8, 3 * BT_H / 8,
                                            (unsigned long long)((B + BT_SB - 1) / BT_SB) * T * 2 * BT_SB * TC_RG};
        const unsigned long long str[4] = {4, rowb, 32, 8 * rowb};
        const unsigned box[4] = {8, 8, BT_KG / 2, TC_RG};
        e = make_tmap_f32(&tm, daimg, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_32B);
        if (e != cudaSuccess) return e;
    }
    seq_bwd_tc_kernel<<<(B + BT_SB - 1) / BT_SB, BT_THREADS, BT_SMEM, st>>>(p, tm);
    return cudaGetLastError();
}

// dX for an encoder layer whose input is the hidden sequence of the layer below (fin == H == 64)
cudaError_t launch_dx_tc(int B, int T, int N, const float* P, const float* Wg, const float* Wc, const float* dA,
                         float* wimg, float* dx, cudaStream_t st) {
    pack_w_dx_kernel<<<BT_CHUNKS, 256, 0, st>>>(Wg, Wc, wimg);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    BwdTcParams p;
    memset(&p, 0, sizeof p);
    p.mode = 1; p.B = B; p.T = T; p.N = N; p.P = P; p.wimg = wimg; p.dA = const_cast<float*>(dA); p.dx = dx;
    e = cudaFuncSetAttribute(seq_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BT_SMEM);
    if (e != cudaSuccess) return e;
    CUtensorMap tm;
    memset(&tm, 0, sizeof tm);
    seq_bwd_tc_kernel<<<(B + BT_SB - 1) / BT_SB, BT_THREADS, BT_SMEM, st>>>(p, tm);
    return cudaGetLastError();
}

}  // namespace dcgru
