// BPTT of the recurrent part of one encoder layer on the tensor cores (tcgen05, 2xFP16): all T steps in reverse for a
// tile of 4 samples per CTA (autograd of model/cell.py:182-210 driven by model/model.py:93-96).
// Per step (dh = gradient of h_t, carried in registers; thread = (row, column half)):
//   E1 : dh += d_hseq[t];  du = dh*(h_{t-1}-c), dc = dh*(1-u), dA_c = dc*act'(c), dA_u = du*u(1-u), dh' = dh*u
//   B1 : d(r*h) = sum_m (P_m^T dA_c) @ Wc_h,m^T                                   (K = M*H, N = H)
//   E2 : dr = d(rh)*h_{t-1}, dA_r = dr*r(1-r), dh' += d(rh)*r
//   B2 : dh_{t-1} = dh' + sum_m (P_m^T [dA_r | dA_u]) @ Wg_h,m^T                   (K = 2*M*H, N = H)
// The transposed diffusion is applied on the INPUT side (P^T commutes with the column contraction), which makes the
// backward structurally the forward: the epilogues write the term-0 chunks [dA_c], [dA_u], [dA_r] (scaled, hi/lo fp16)
// into two dedicated slots, one warp per (sample, term) applies P_m^T to them on the warp-level tensor path (mma.sync
// 2xFP16, f16_common.cuh::diffuse_mma16) into a 3-slot ring, kind::f16 MMAs accumulate in TMEM.  The u part of B2 does
// not depend on B1, so its chunks are diffused while B1's MMAs complete.
// The term-0 chunks [dA_c], [dA_u], [dA_r] (scaled, hi/lo) are exactly the rows of the dA operand image that the
// weight-gradient GEMM (dw_mm16.cu), the input-gradient GEMM (bulk_dp.cu) and the bias gradient read: the dump warp
// stores them with tensor-map TMA; nothing else is written per step.
// Saved forward state (r, u, c, h_{t-1}) and the upstream gradient are prefetched one step ahead into TMEM by four
// loader warps (thread = row), so the epilogues never wait for global memory.
// Gradient scaling: every fp16 operand derived from the gradient is multiplied by s = *scale_ptr, a power of two chosen
// from the upstream gradient's magnitude (grad_scale_kernel), and the accumulators are multiplied by 1/s when read.
//   warps 0-7 workers (epilogues + diffusion), warp 8 MMA issuer, warp 9 weight loader, warp 10 image dump,
//   warps 11-14 state loaders
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "dw.cuh"
#include "f16_common.cuh"
#include "tmap.cuh"

namespace dcgru {
using namespace f16;

constexpr int RB_H = 64;
constexpr int RB_THREADS = 480;
constexpr int RB_NWORK = 256;
constexpr int RB_NW = 4;                             // weight ring slots
constexpr int RB_WPIECE = RB_H * 128;                // one plane of a chunk's weights: 64 rows x 128 B
constexpr int RB_NS = 5;                             // chunk slots: 0 = [dA_c] then [dA_r], 1 = [dA_u], 2..4 = ring of the diffused chunks
constexpr int RB_OFF_W = RB_NS * SLOT;
constexpr int RB_OFF_PT = RB_OFF_W + RB_NW * RB_WPIECE;
// TMEM columns
constexpr int RB_ACC1 = 0, RB_ACC2 = 64, RB_U = 128, RB_C = 192, RB_HP = 256, RB_DUP = 320, RB_R = 384, RB_HP1 = 448;   // h_{t-1}: 256 / 448 alternate

struct RnnBwdParams {
    int B, T, N, M, act, dump, dbg;
    int img_T, img_t0;            // slab of step t in the dA image: tile * img_T + img_t0 + t
    const float* h0; const float* hseq; const float* ruc;
    const float* P;
    const float* d_hseq; const float* d_hlast;
    const float* d_hsel; const int* sel_t;   // sparse upstream gradient: slab b belongs to step sel_t[b] (head.cu), or nullptr
    const uint8_t* wimg;          // B1 planes [m][hi|lo], then B2 planes [2m+g][hi|lo] (g = 0: r, 1: u), 8 KB each
    const float* scale_ptr;
    float* dh0;
};

// chunk i of a step (0 <= i < 3M): kind 0 = c (-> acc1), 1 = u, 2 = r (-> acc2); term m
__device__ __forceinline__ void rb_chunk(int i, int M, int& kind, int& m) {
    if (i == 0) { kind = 0; m = 0; }
    else if (i == 1) { kind = 1; m = 0; }
    else if (i <= M) { kind = 0; m = i - 1; }
    else if (i < 2 * M) { kind = 1; m = i - M; }
    else { kind = 2; m = i - 2 * M; }
}

// slot and fill index (how many times the slot was filled before) of chunk i of processed step k: the term-0 chunks live in
// slots 0 / 1 for a whole step (they are the sources of the diffusion), the diffused chunks rotate through slots 2..4
__device__ __forceinline__ void rb_slot(int i, int k, int M, int& slot, unsigned& fill) {
    if (i == 0) { slot = 0; fill = 2u * k; }
    else if (i == 1) { slot = 1; fill = (unsigned)k; }
    else if (i == 2 * M) { slot = 0; fill = 2u * k + 1u; }
    else {
        const unsigned d = (unsigned)k * (3 * (M - 1)) + (unsigned)(i < 2 * M ? i - 2 : i - 3);
        slot = 2 + (int)(d % 3u); fill = d / 3u;
    }
}

// timing experiment (DCGRU_DBG & 32): clock64 stamps of CTA 0, worker thread 0: [step][0..9]
__device__ long long rb_dbg[64 * 16];

__device__ __forceinline__ void rb_worker_bar() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }

__global__ void __launch_bounds__(RB_THREADS, 1) rnn_bwd_kernel(const RnnBwdParams p, const __grid_constant__ CUtensorMap tm_img) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar_afull[RB_NS], bar_aempty[RB_NS], bar_stored[RB_NS], bar_wfull[RB_NW], bar_wempty[RB_NW];
    __shared__ uint64_t bar_gafull, bar_gbfull, bar_gafree, bar_gbfree, bar_b1, bar_b2;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = p.N, M = p.M, T = p.T;
    const int tile = blockIdx.x, b0 = tile * SB;
    const int nch = 3 * M;
    uint8_t* Aslots = smem;
    uint8_t* Wring = smem + RB_OFF_W;
    float* PTs = reinterpret_cast<float*>(smem + RB_OFF_PT);
    const bool dump = p.dump != 0;
    const size_t NH = (size_t)N * RB_H;

    if (warp == 0) tmem_alloc<512>(&tmem_slot);
    if (tid == 0) {
        for (int i = 0; i < RB_NS; ++i) { mbar_init(&bar_afull[i], RB_NWORK / 32); mbar_init(&bar_aempty[i], 1); mbar_init(&bar_stored[i], 1); }
        for (int i = 0; i < RB_NW; ++i) { mbar_init(&bar_wfull[i], 1); mbar_init(&bar_wempty[i], 1); }
        mbar_init(&bar_gafull, 4); mbar_init(&bar_gbfull, 4);
        mbar_init(&bar_gafree, RB_NWORK / 32); mbar_init(&bar_gbfree, RB_NWORK / 32);
        mbar_init(&bar_b1, 1); mbar_init(&bar_b2, 1);
        mbar_fence_init();
    }
    for (int i = tid; i < RB_NS * SLOT / 16; i += RB_THREADS) reinterpret_cast<uint4*>(Aslots)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < SB * (M - 1) * PT_STRIDE; i += RB_THREADS) PTs[i] = 0.f;
    __syncthreads();
    load_pt(PTs, p.P, b0, p.B, N, M - 1, 1, tid, RB_THREADS);                     // rows of P: (P^T z)[n] = sum_j P[j][n] z[j]
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = tmem_slot;

    if (warp == 8) {
        // =================================== MMA issuer =================================================================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_f16(128, RB_H);
            const uint32_t a_base = smem_u32(Aslots), w_base = smem_u32(Wring);
            const int last_c = (M == 1) ? 0 : M;
            unsigned pc = 0;
            for (int k = 0; k < T; ++k) {
                for (int i = 0; i < nch; ++i) {
                    int kind, m, slot;
                    unsigned fill;
                    rb_chunk(i, M, kind, m);
                    rb_slot(i, k, M, slot, fill);
                    DBG_PROG(k * 100 + i);
                    const uint32_t ah = a_base + slot * SLOT, al = ah + PLANE;
                    const uint32_t d = taddr + (kind == 0 ? RB_ACC1 : RB_ACC2);
                    const uint32_t first = (i <= 1) ? 0u : 1u;
                    const int ws0 = pc % RB_NW, ws1 = (pc + 1) % RB_NW;
                    mbar_wait2(&bar_afull[slot], fill & 1, &bar_wfull[ws0], (pc / RB_NW) & 1);
                    tc_fence_after();
                    const uint32_t bh = w_base + ws0 * RB_WPIECE, bl = w_base + ws1 * RB_WPIECE;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        umma_f16(d, make_desc_k128(al + 32 * ks), make_desc_k128(bh + 32 * ks), idesc, (ks == 0) ? first : 1u);
                        umma_f16(d, make_desc_k128(ah + 32 * ks), make_desc_k128(bh + 32 * ks), idesc, 1u);
                    }
                    umma_commit(&bar_wempty[ws0]);
                    mbar_wait(&bar_wfull[ws1], ((pc + 1) / RB_NW) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_f16(d, make_desc_k128(ah + 32 * ks), make_desc_k128(bl + 32 * ks), idesc, 1u);
                    umma_commit(&bar_wempty[ws1]);
                    umma_commit(&bar_aempty[slot]);
                    pc += 2;
                    if (i == last_c) umma_commit(&bar_b1);
                }
                umma_commit(&bar_b2);
            }
        }
        __syncwarp();
    } else if (warp == 9) {
        // =================================== weight loader ==============================================================
        if (lane == 0) {
            unsigned pc = 0;
            for (int k = 0; k < T; ++k)
                for (int i = 0; i < nch; ++i) {
                    int kind, m;
                    rb_chunk(i, M, kind, m);
                    const int piece0 = (kind == 0) ? 2 * m : 2 * M + 2 * (2 * m + (kind == 1 ? 1 : 0));
                    for (int pl = 0; pl < 2; ++pl, ++pc) {
                        const int ws = pc % RB_NW;
                        DBG_PROG(k * 100 + i);
                        if (pc >= RB_NW) mbar_wait(&bar_wempty[ws], ((pc / RB_NW) - 1) & 1);
                        mbar_expect_tx(&bar_wfull[ws], RB_WPIECE);
                        bulk_g2s(Wring + ws * RB_WPIECE, p.wimg + (size_t)(piece0 + pl) * RB_WPIECE, RB_WPIECE, &bar_wfull[ws]);
                    }
                }
        }
        __syncwarp();
    } else if (warp == 10) {
        // =================================== dA-image dump ================================================================
        if (dump) {
            const int plane = lane >> 2, s = lane & 3;
            if (lane == 0) tma_prefetch_desc(&tm_img);
            for (int k = 0; k < T; ++k) {
                const int t = T - 1 - k;
                const long slab = (long)tile * p.img_T + p.img_t0 + t;
                for (int i = 0; i < nch; ++i) {
                    int slot;
                    unsigned fill;
                    rb_slot(i, k, M, slot, fill);
                    DBG_PROG(k * 100 + i);
                    mbar_wait(&bar_afull[slot], fill & 1);
                    const int col = (i == 0) ? 2 * RB_H : (i == 1 ? RB_H : (i == 2 * M ? 0 : -1));   // image columns r | u | c
                    if (col >= 0) {
                        if (lane < 2 * SB) {
                            tma_store_2d(&tm_img, col, (int)((slab * 2 + plane) * IMG_ROWS + s * (RG * 8)),
                                         Aslots + slot * SLOT + plane * PLANE + s * (RP * 128));
                            bulk_commit();
                        }
                        bulk_wait_read();
                    }
                    // every lane has seen this chunk's phase before lane 0 lets the slot be refilled: found by the stress build --
                    // without it, lanes that lag behind lane 0 (only its arrival gates the producers) can be overtaken by two
                    // phases on a ring slot and then wait for ever
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_stored[slot]);
                }
            }
            bulk_wait_all();
        }
        __syncwarp();
    } else if (warp >= 11) {
        // =================================== state loaders: thread = row ==================================================
        const int quad = warp & 3, row = 32 * quad + lane;
        const int s = row >> 5, n = row & 31, b = b0 + s;
        const bool valid = n < N && b < p.B;
        const uint32_t tb = taddr + ((uint32_t)(32 * quad) << 16);
        const size_t rbase = valid ? ((size_t)b * N + n) : 0;
        int sel = -1;
        if (p.d_hsel && valid) { sel = p.sel_t ? p.sel_t[b] : T - 1; sel = sel < 0 ? 0 : (sel >= T ? T - 1 : sel); }
        auto load_cols = [&](const float* src, uint32_t tcol) {          // 64 floats of this row -> TMEM columns
#pragma unroll 1
            for (int cb = 0; cb < RB_H; cb += 32) {
                float v[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (valid && src) q = __ldcs(reinterpret_cast<const float4*>(src + cb) + j);
                    v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
                }
                tmem_st32(tb + tcol + cb, v);
            }
        };
        for (int k = 0; k < T; ++k) {
            const int t = T - 1 - k;
            const float* ruc = p.ruc + ((size_t)t * p.B * N + rbase) * 3 * RB_H;
            const float* hp = (t > 0) ? p.hseq + (size_t)(t - 1) * p.B * NH + rbase * RB_H : p.h0 + rbase * RB_H;
            DBG_PROG(k * 10 + 1);
            if (k >= 1) mbar_wait(&bar_gafree, (k - 1) & 1);
            tc_fence_after();
            DBG_PROG(k * 10 + 2);
            // group A (needed by E1): u, c, upstream gradient, h_{t-1}.  h_{t-1} is also read by E2, so its TMEM columns alternate
            // between two buffers: the copy of step k is still being read when the one of step k+1 arrives (every worker warp
            // passes E2 of step k-1 before it frees group A of step k)
            load_cols(ruc + RB_H, RB_U);
            load_cols(ruc + 2 * RB_H, RB_C);
            load_cols(hp, (k & 1) ? RB_HP1 : RB_HP);
            load_cols(p.d_hsel ? (sel == t ? p.d_hsel + rbase * RB_H : nullptr)
                               : (p.d_hseq ? p.d_hseq + (size_t)t * p.B * NH + rbase * RB_H : nullptr), RB_DUP);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_gafull);
            DBG_PROG(k * 10 + 3);
            if (k >= 1) mbar_wait(&bar_gbfree, (k - 1) & 1);
            tc_fence_after();
            DBG_PROG(k * 10 + 4);
            load_cols(ruc, RB_R);                                         // group B (needed by E2 only): r
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_gbfull);
        }
    } else {
        // =================================== workers =====================================================================
        const int quad = warp & 3, half = warp >> 2;
        const int row = 32 * quad + lane;
        const int b_ = b0 + quad;
        const bool rvalid = lane < N && b_ < p.B;
        const uint32_t tb = taddr + ((uint32_t)(32 * quad) << 16) + half * 32;
        const float gs = __ldg(p.scale_ptr), inv_gs = 1.f / gs;
        // slot of chunk i of step k, once its previous content was consumed by the MMAs (and dumped)
        // (need_mma = false: the MMAs that read the previous content are known to be complete through bar_b1 / bar_b2)
        auto acquire = [&](int i, int k, bool need_mma = true) -> int {
            int slot;
            unsigned fill;
            rb_slot(i, k, M, slot, fill);
            if (fill >= 1) {
                if (need_mma) mbar_wait(&bar_aempty[slot], (fill - 1) & 1);
                if (dump) mbar_wait(&bar_stored[slot], (fill - 1) & 1);
            }
            return slot;
        };
        auto publish = [&](int slot) {
            tc_fence_before();
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_afull[slot]);
        };
        auto put8 = [&](uint8_t* sl, int col, const float (&v)[8]) {   // 8 columns of this row -> chunk (scaled, hi / lo)
            float w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = v[j] * gs;
            uint4 hi, lo;
            split8(w, hi, lo);
            const uint32_t off = k128_off(row, col);
            *reinterpret_cast<uint4*>(sl + off) = hi;
            *reinterpret_cast<uint4*>(sl + PLANE + off) = lo;
        };
        // diffused chunks of one gate: one warp per (sample, term) applies P_m^T to the term-0 chunk in slot `src` (scaled
        // hi / lo fp16, written by the epilogue) on the warp-level tensor path; 4 tasks per chunk, each arriving with count 2.
        // i0 = chunk index of term 1 of this gate
        auto diffuse_group = [&](int src, int i0, int k) {
            for (int m = 1; m < M; ++m) {
                const int s = (warp - ((m - 1) * SB)) & 7;
                if (s < SB) {
                    const int slot = acquire(i0 + m - 1, k);
                    PFrag pf;
                    load_pfrag(PTs + (s * (M - 1) + (m - 1)) * PT_STRIDE, lane, pf);
                    diffuse_mma16(Aslots + src * SLOT, s * RP, pf, Aslots + slot * SLOT, s * RP, lane, 1.f);
                    tc_fence_before();
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_n(&bar_afull[slot], 2);
                }
            }
        };
        float dhp[32];                                                  // dh*u (+ d(rh)*r): the elementwise part of dh_{t-1}
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rvalid && p.d_hlast) q = *reinterpret_cast<const float4*>(p.d_hlast + ((size_t)b_ * N + lane) * RB_H + half * 32 + 4 * j);
            dhp[4 * j] = q.x; dhp[4 * j + 1] = q.y; dhp[4 * j + 2] = q.z; dhp[4 * j + 3] = q.w;
        }
        for (int k = 0; k < T; ++k) {
            // ---- E1 ----------------------------------------------------------------------------------------------------
            const bool rec = p.dbg && blockIdx.x == 0 && tid == 0 && k < 64;
            long long* es = rb_dbg + k * 16;
            if (rec) es[0] = clock64();
            DBG_PROG(k * 10 + 0);
            mbar_wait(&bar_gafull, k & 1);
            DBG_PROG(k * 10 + 1);
            const uint32_t thp = tb + ((k & 1) ? RB_HP1 : RB_HP);
            if (rec) es[1] = clock64();
            if (k >= 1) mbar_wait(&bar_b2, (k - 1) & 1);                // dh_t's GEMM part (B2 of step t+1) is in acc2
            tc_fence_after();
            if (rec) es[2] = clock64();
            DBG_PROG(k * 10 + 2);
            uint8_t* slc = Aslots + acquire(0, k, false) * SLOT;        // chunk [dA_c] (slot 0; [dA_r] of step k-1 was read before bar_b2)
            uint8_t* slu = Aslots + acquire(1, k, false) * SLOT;        // chunk [dA_u] (slot 1; likewise)
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {                            // 16 columns at a time (TMEM loads are paid per instruction)
                float u[16], c[16], hp[16], g[16], a2[16];
                tmem_ld16_nw(tb + RB_U + 16 * cc, u);
                tmem_ld16_nw(tb + RB_C + 16 * cc, c);
                tmem_ld16_nw(thp + 16 * cc, hp);
                tmem_ld16_nw(tb + RB_DUP + 16 * cc, g);
                if (k >= 1) tmem_ld16_nw(tb + RB_ACC2 + 16 * cc, a2);
                tmem_wait_ld();
                if (k >= 1) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) g[j] += a2[j] * inv_gs;
                }
#pragma unroll
                for (int h8 = 0; h8 < 2; ++h8) {
                    float dac[8], dau[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int e = 8 * h8 + j;
                        const float dh = rvalid ? dhp[16 * cc + e] + g[e] : 0.f;
                        const float du = dh * (hp[e] - c[e]);
                        const float dc = dh * (1.f - u[e]);
                        dac[j] = (p.act == 0) ? dc * (1.f - c[e] * c[e]) : (c[e] > 0.f ? dc : 0.f);
                        dau[j] = du * u[e] * (1.f - u[e]);
                        dhp[16 * cc + e] = dh * u[e];
                    }
                    const int col = half * 32 + 16 * cc + 8 * h8;
                    put8(slc, col, dac);
                    put8(slu, col, dau);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_gafree);                    // u, c, d_hseq[t] consumed
            publish(0);                                                 // [dA_c]
            publish(1);                                                 // [dA_u]
            if (rec) es[3] = clock64();
            rb_worker_bar();                                            // both column halves of every row are written
            if (rec) es[4] = clock64();
            // ---- diffusion for B1, then the u half of B2 (independent of B1: fills B1's MMA latency) ---------------------
            DBG_PROG(k * 10 + 3);
            diffuse_group(0, 2, k);
            if (rec) es[5] = clock64();
            DBG_PROG(k * 10 + 4);
            diffuse_group(1, M + 1, k);
            if (rec) es[6] = clock64();
            DBG_PROG(k * 10 + 5);
            mbar_wait2(&bar_b1, k & 1, &bar_gbfull, k & 1);            // d(rh) is in acc1; r of this step is in TMEM
            tc_fence_after();
            if (rec) es[7] = clock64();
            // ---- E2 ----------------------------------------------------------------------------------------------------
            DBG_PROG(k * 10 + 6);
            uint8_t* slr = Aslots + acquire(2 * M, k, false) * SLOT;    // chunk [dA_r] (slot 0: every reader of [dA_c] is done, bar_b1)
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                float a1[16], hp[16], r[16];
                tmem_ld16_nw(tb + RB_ACC1 + 16 * cc, a1);
                tmem_ld16_nw(thp + 16 * cc, hp);
                tmem_ld16_nw(tb + RB_R + 16 * cc, r);
                tmem_wait_ld();
#pragma unroll
                for (int h8 = 0; h8 < 2; ++h8) {
                    float dar[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int e = 8 * h8 + j;
                        const float drh = rvalid ? a1[e] * inv_gs : 0.f;
                        dar[j] = drh * hp[e] * r[e] * (1.f - r[e]);
                        dhp[16 * cc + e] += drh * r[e];
                    }
                    const int col = half * 32 + 16 * cc + 8 * h8;
                    put8(slr, col, dar);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_gbfree);                    // r, h_{t-1} consumed
            publish(0);                                                 // [dA_r]
            if (rec) es[8] = clock64();
            DBG_PROG(k * 10 + 7);
            rb_worker_bar();
            DBG_PROG(k * 10 + 8);
            diffuse_group(0, 2 * M + 1, k);
            DBG_PROG(k * 10 + 9);
            if (rec) es[9] = clock64();
        }
        // ---- dh0 = dh' + B2 of the last processed step (t = 0) ------------------------------------------------------------
        mbar_wait(&bar_b2, (T - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            float a2[8];
            tmem_ld8(tb + RB_ACC2 + 8 * cc, a2);
            if (rvalid) {
                float4* d = reinterpret_cast<float4*>(p.dh0 + ((size_t)b_ * N + lane) * RB_H + half * 32 + 8 * cc);
                d[0] = make_float4(dhp[8 * cc] + a2[0] * inv_gs, dhp[8 * cc + 1] + a2[1] * inv_gs, dhp[8 * cc + 2] + a2[2] * inv_gs, dhp[8 * cc + 3] + a2[3] * inv_gs);
                d[1] = make_float4(dhp[8 * cc + 4] + a2[4] * inv_gs, dhp[8 * cc + 5] + a2[5] * inv_gs, dhp[8 * cc + 6] + a2[6] * inv_gs, dhp[8 * cc + 7] + a2[7] * inv_gs);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(taddr);
}

// ---- gradient scale: s = 2^e with max|g| * s in [8, 16) (1 when the gradient is all zero / not finite) ---------------------
__global__ void __launch_bounds__(256) grad_absmax_kernel(const float* a, size_t na, const float* b, size_t nb, unsigned* out) {
    float m = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x, i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    auto scan = [&](const float* q, size_t n) {                        // 16-byte loads, 4 in flight per thread (n % 4 == 0, 16-byte aligned)
        const float4* q4 = reinterpret_cast<const float4*>(q);
        const size_t n4 = n / 4;
        size_t i = i0;
        for (; i + 3 * stride < n4; i += 4 * stride) {
            const float4 v0 = __ldcs(q4 + i), v1 = __ldcs(q4 + i + stride), v2 = __ldcs(q4 + i + 2 * stride), v3 = __ldcs(q4 + i + 3 * stride);
            m = fmaxf(m, fmaxf(fmaxf(fmaxf(fabsf(v0.x), fabsf(v0.y)), fmaxf(fabsf(v0.z), fabsf(v0.w))),
                               fmaxf(fmaxf(fabsf(v1.x), fabsf(v1.y)), fmaxf(fabsf(v1.z), fabsf(v1.w)))));
            m = fmaxf(m, fmaxf(fmaxf(fmaxf(fabsf(v2.x), fabsf(v2.y)), fmaxf(fabsf(v2.z), fabsf(v2.w))),
                               fmaxf(fmaxf(fabsf(v3.x), fabsf(v3.y)), fmaxf(fabsf(v3.z), fabsf(v3.w)))));
        }
        for (; i < n4; i += stride) {
            const float4 v = __ldcs(q4 + i);
            m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
    };
    if (a) scan(a, na);
    if (b) scan(b, nb);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));          // non-negative floats order like their bit patterns
}
__global__ void grad_scale_kernel(unsigned* maxbits, float* scale) {
    const float m = __uint_as_float(*maxbits);
    float s = 1.f;
    if (m > 0.f && m < 3.0e38f) {
        int e;
        frexpf(m, &e);                                                  // m = f * 2^e, f in [0.5, 1)
        int sh = 4 - e;                                                 // max * s in [8, 16)
        sh = sh > 100 ? 100 : (sh < -100 ? -100 : sh);
        s = ldexpf(1.f, sh);
    }
    *scale = s;
    *maxbits = 0u;
}
// scale[0] <- s, computed on the stream from the two upstream gradients (either may be nullptr); scratch: one unsigned, zeroed here
cudaError_t launch_grad_scale(const float* a, size_t na, const float* b, size_t nb, const float* c, size_t nc, unsigned* scratch,
                              float* scale, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(unsigned), st);
    if (e != cudaSuccess) return e;
    if (a || b || !c) {
        grad_absmax_kernel<<<1184, 256, 0, st>>>(a, na, b, nb, scratch);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    if (c) {
        const int grid = (int)((nc / 4 + 255) / 256 < 1184 ? (nc / 4 + 255) / 256 : 1184);
        grad_absmax_kernel<<<grid < 1 ? 1 : grid, 256, 0, st>>>(c, nc, nullptr, 0, scratch);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    grad_scale_kernel<<<1, 1, 0, st>>>(scratch, scale);
    return cudaGetLastError();
}

cudaError_t rnn_bwd_read_dbg(long long* out, int n) {
    return cudaMemcpyFromSymbol(out, rb_dbg, sizeof(long long) * (n < 1024 ? n : 1024));
}
size_t rnn_bwd_wimg_bytes(int M) { return (size_t)(2 * M + 4 * M) * RB_WPIECE; }
int rnn_bwd_smem_bytes(int M) { return RB_OFF_PT + SB * (M - 1) * PT_STRIDE * 4 + 1024; }
bool rnn_bwd_supported(int N, int H, int M, int smem_limit) {
    return H == RB_H && N <= NPAD && M >= 1 && M <= 7 && rnn_bwd_smem_bytes(M) + 1024 <= smem_limit;
}

cudaError_t rnn_bwd_pack_weights(const float* Wg, const float* Wc, int fin, int M, void* wimg, cudaStream_t st) {
    cudaError_t e = launch_pack_w16(Wg, Wc, fin, RB_H, M, 4, RB_H, M, wimg, st);
    if (e != cudaSuccess) return e;
    return launch_pack_w16(Wg, Wc, fin, RB_H, M, 5, RB_H, 2 * M, reinterpret_cast<uint8_t*>(wimg) + (size_t)2 * M * RB_WPIECE, st);
}

// dA image: [tile*T + t][hi|lo][96][3H] fp16, columns r | u | c, values scaled by *scale_ptr
cudaError_t launch_rnn_bwd(int B, int T, int N, int fin, int M, int act, const float* h0, const float* hseq, const float* ruc,
                           const float* P, const float* Wg, const float* Wc, const float* d_hseq, const float* d_hlast,
                           const float* d_hsel, const int* sel_t, void* wimg, const float* scale_ptr, float* dh0, void* daimg,
                           cudaStream_t st, int img_T, int img_t0) {
    cudaError_t e = cudaSuccess;
    if (Wg) {
        e = rnn_bwd_pack_weights(Wg, Wc, fin, M, wimg, st);
        if (e != cudaSuccess) return e;
    }
    RnnBwdParams p;
    memset(&p, 0, sizeof p);
    p.B = B; p.T = T; p.N = N; p.M = M; p.act = act; p.dump = daimg != nullptr;
    p.img_T = img_T > 0 ? img_T : T; p.img_t0 = img_T > 0 ? img_t0 : 0;
    { const char* e = getenv("DCGRU_DBG"); p.dbg = e ? (atoi(e) & 32) : 0; }
    p.h0 = h0; p.hseq = hseq; p.ruc = ruc; p.P = P; p.d_hseq = d_hseq; p.d_hlast = d_hlast;
    p.d_hsel = d_hsel; p.sel_t = sel_t;
    p.wimg = reinterpret_cast<const uint8_t*>(wimg); p.scale_ptr = scale_ptr; p.dh0 = dh0;
    CUtensorMap tm;
    memset(&tm, 0, sizeof tm);
    const int ntile = g16_ntile(B);
    if (daimg) {
        const unsigned long long dims[2] = {(unsigned long long)3 * RB_H, (unsigned long long)ntile * p.img_T * 2 * IMG_ROWS};
        const unsigned long long str[2] = {2, (unsigned long long)3 * RB_H * 2};
        const unsigned box[2] = {64, RG * 8};
        e = make_tmap(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, daimg, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (e != cudaSuccess) return e;
    }
    const int smem = rnn_bwd_smem_bytes(M);
    e = cudaFuncSetAttribute(rnn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    rnn_bwd_kernel<<<ntile, RB_THREADS, smem, st>>>(p, tm);
    return cudaGetLastError();
}

}  // namespace dcgru

namespace dcgru {
// operand image [tile*T+t][hi|lo][96][cols] (scaled by *scale_ptr) -> row-major fp32 (T,B,N,cols): diagnostics, and the
// bridge to the first-generation weight-gradient kernels
// image slab t' lands in output slab (t' / group) * stride + off + t' % group (decoder: layers interleaved per step)
__global__ void img_to_rows_kernel(const __half* img, int B, int T, int N, int cols, const float* scale_ptr, float* out, int group,
                                   int stride, int off) {
    const float inv = 1.f / scale_ptr[0];
    const size_t total = (size_t)T * B * N * cols;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx % cols);
        size_t r = idx / cols;
        const int n = (int)(r % N); r /= N;
        const int b = (int)(r % B);
        const int t = (int)(r / B);
        const size_t slab = (size_t)(b / f16::SB) * T + t;
        const size_t o = ((slab * 2) * f16::IMG_ROWS + (b % f16::SB) * (f16::RG * 8) + n) * cols + c;
        const int ot = (t / group) * stride + off + t % group;
        out[(((size_t)ot * B + b) * N + n) * cols + c] = (__half2float(img[o]) + __half2float(img[o + (size_t)f16::IMG_ROWS * cols])) * inv;
    }
}
cudaError_t launch_img_to_rows(const void* img, int B, int T, int N, int cols, const float* scale_ptr, float* out, int group,
                               int stride, int off, cudaStream_t st) {
    img_to_rows_kernel<<<1184, 256, 0, st>>>(reinterpret_cast<const __half*>(img), B, T, N, cols, scale_ptr, out, group, stride, off);
    return cudaGetLastError();
}
}  // namespace dcgru
