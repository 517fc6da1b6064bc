// Recurrent part of one encoder layer on the tensor cores (tcgen05, 2xFP16): all T steps for a tile of 4 samples per
// CTA, no relaunch (model/model.py:93-96 x model/cell.py:182-210).  The non-recurrent x-part of both projections,
// incl. the biases, was hoisted into one bulk GEMM over all steps (bulk_dp.cu) and arrives as the pre-activation
// XP (T,B,N,3H); per step this kernel only does what depends on h_{t-1}:
//     gate : D[:, 0:2H]  = XP_t[:, 0:2H]  + sum_m (P_m h_{t-1})      @ Wg_h,m     -> r = sigmoid(D[:, 0:H])
//     cand : D[:, 2H:3H] = XP_t[:, 2H:3H] + sum_m (P_m (r*h_{t-1}))  @ Wc_h,m     -> c = act(.), u = sigmoid(D[:, H:2H])
//     h_t = u*h_{t-1} + (1-u)*c
// K order is term-major (kk = m*H + c), so with H = 64 one chunk (f16_common.cuh) = one diffusion term:
//   * term 0 (the state itself) is written as hi/lo fp16 by the epilogue that produced it (slot 0),
//   * terms m >= 1 by one worker warp per (sample, term): fp32 diffusion from the state tile ZH (lane = two columns,
//     registers = the 19 output rows), slots 1 and 2 alternate.
// XP_t is loaded straight into the accumulator (TMEM) one or two steps ahead by four loader warps and the MMAs
// accumulate on top of it; the accumulator is double buffered, the GRU state also lives in a TMEM stash.
//   warps 0-7   workers: diffusion tasks + the two epilogues (thread = (row, column half))
//   warp 8      MMA issuer;  warp 9 weight loader (cp.async.bulk ring of hi / lo planes, L2 resident)
//   warp 10     operand-image dump (tensor-map TMA store of every A chunk: columns [KXP, KXP + 2*M*H) of the image)
//   warps 11-14 XP loaders (thread = row: global -> registers -> tcgen05.st)
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "dw.cuh"
#include "f16_common.cuh"
#include "tmap.cuh"

namespace dcgru {
using namespace f16;

constexpr int RF_H = 64;
constexpr int RF_THREADS = 480;
constexpr int RF_NWORK = 256;
constexpr int RF_ZLD = RF_H + 4;                    // state tile row stride (floats): conflict-free float4 row writes
constexpr int RF_NW = 4;                            // weight ring slots
constexpr int RF_WSLOT = 16 * 1024;                 // one plane of a gate chunk (128 rows x 128 B); candidate planes are 8 KB
constexpr int RF_OFF_A = 0;                         // 3 chunk slots
constexpr int RF_OFF_W = 3 * SLOT;
constexpr int RF_OFF_ZH = RF_OFF_W + RF_NW * RF_WSLOT;
constexpr int RF_OFF_PT = RF_OFF_ZH + 128 * RF_ZLD * 4;
constexpr int RF_ACC1 = 192, RF_STASH = 384, RF_RSTASH = 448;   // TMEM columns: accumulators at 0 / 192, h stash, r stash

struct RnnFwdParams {
    int B, T, N, M, act, dump, img_col0, dbg;
    int img_T, img_t0;          // slab of step t in the operand image: tile * img_T + img_t0 + t
    int mma_diff;               // diffusion on the warp-level tensor path (DCGRU_MMA_DIFF=0: fp32 FMA loop)
    const float* xp;            // (T,B,N,3H)
    const float* h0;            // (B,N*H)
    const float* P;             // (B,M-1,N,N)
    const uint8_t* wimg;        // gate planes [m][hi|lo] (16 KB each), then candidate planes [m][hi|lo] (8 KB each)
    float* hseq;                // (T,B,N*H)
    float* ruc;                 // (T,B,N,3H) or nullptr
};

// timing experiment (DCGRU_DBG & 16): clock64 stamps of CTA 0 -- worker thread 0: [t][0..9], issuer: [t][10..15]
__device__ long long rf_dbg[64 * 16 + 64 * 8];   // + per-step stamps inside the first diffusion term (offset 1024)

__device__ __forceinline__ void rf_worker_bar() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }

__global__ void __launch_bounds__(RF_THREADS, 1) rnn_fwd_kernel(const RnnFwdParams p, const __grid_constant__ CUtensorMap tm_img) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar_afull[3], bar_aempty[3], bar_stored[3], bar_wfull[RF_NW], bar_wempty[RF_NW];
    __shared__ uint64_t bar_xpfull[2], bar_accfree[2], bar_gate, bar_cand;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = p.N, M = p.M, T = p.T;
    const int tile = blockIdx.x, b0 = tile * SB;
    uint8_t* Aslots = smem + RF_OFF_A;
    uint8_t* Wring = smem + RF_OFF_W;
    float* ZH = reinterpret_cast<float*>(smem + RF_OFF_ZH);
    float* PTs = reinterpret_cast<float*>(smem + RF_OFF_PT);
    const bool dump = p.dump != 0;
    const size_t NH = (size_t)N * RF_H;

    if (warp == 0) tmem_alloc<512>(&tmem_slot);
    if (tid == 0) {
        mbar_init(&bar_afull[0], RF_NWORK / 32);
        mbar_init(&bar_afull[1], RF_NWORK / 32);
        mbar_init(&bar_afull[2], RF_NWORK / 32);
        for (int i = 0; i < 3; ++i) { mbar_init(&bar_aempty[i], 1); mbar_init(&bar_stored[i], 1); }
        for (int i = 0; i < RF_NW; ++i) { mbar_init(&bar_wfull[i], 1); mbar_init(&bar_wempty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&bar_xpfull[i], 4); mbar_init(&bar_accfree[i], RF_NWORK / 32); }
        mbar_init(&bar_gate, 1);
        mbar_init(&bar_cand, 1);
        mbar_fence_init();
    }
    for (int i = tid; i < 3 * SLOT / 16; i += RF_THREADS) reinterpret_cast<uint4*>(Aslots)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < 128 * RF_ZLD; i += RF_THREADS) ZH[i] = 0.f;
    for (int i = tid; i < SB * (M - 1) * PT_STRIDE; i += RF_THREADS) PTs[i] = 0.f;
    __syncthreads();
    load_pt(PTs, p.P, b0, p.B, N, M - 1, 0, tid, RF_THREADS);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = tmem_slot;

    if (warp == 8) {
        // =================================== MMA issuer =================================================================
        if (lane == 0) {
            const uint32_t idg = make_idesc_f16(128, 2 * RF_H), idc = make_idesc_f16(128, RF_H);
            const uint32_t a_base = smem_u32(Aslots), w_base = smem_u32(Wring);
            unsigned pc = 0, fills[3] = {0, 0, 0};
            for (int t = 0; t < T; ++t) {
                const int acc_i = t & 1;
                const uint32_t dacc = taddr + acc_i * RF_ACC1;
                const bool rec = (p.dbg & 16) && blockIdx.x == 0 && t < 64;
                if (rec) rf_dbg[t * 16 + 10] = clock64();
                mbar_wait(&bar_xpfull[acc_i], (t >> 1) & 1);            // XP_t sits in the accumulator
                if (rec) rf_dbg[t * 16 + 11] = clock64();
                for (int ph = 0; ph < 2; ++ph) {
                    const uint32_t d = dacc + (ph ? 2 * RF_H : 0), idesc = ph ? idc : idg;
                    for (int m = 0; m < M; ++m) {
                        const int slot = m == 0 ? 0 : 1 + ((m - 1) & 1);
                        const uint32_t ah = a_base + slot * SLOT, al = ah + PLANE;
                        const int ws0 = pc % RF_NW, ws1 = (pc + 1) % RF_NW;
                        mbar_wait2(&bar_afull[slot], fills[slot] & 1, &bar_wfull[ws0], (pc / RF_NW) & 1);
                        ++fills[slot];
                        tc_fence_after();
                        const uint32_t bh = w_base + ws0 * RF_WSLOT, bl = w_base + ws1 * RF_WSLOT;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            umma_f16(d, make_desc_k128(al + 32 * ks), make_desc_k128(bh + 32 * ks), idesc, 1u);
                            umma_f16(d, make_desc_k128(ah + 32 * ks), make_desc_k128(bh + 32 * ks), idesc, 1u);
                        }
                        umma_commit(&bar_wempty[ws0]);
                        mbar_wait(&bar_wfull[ws1], ((pc + 1) / RF_NW) & 1);
                        tc_fence_after();
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            umma_f16(d, make_desc_k128(ah + 32 * ks), make_desc_k128(bl + 32 * ks), idesc, 1u);
                        umma_commit(&bar_wempty[ws1]);
                        umma_commit(&bar_aempty[slot]);
                        pc += 2;
                    }
                    umma_commit(ph ? &bar_cand : &bar_gate);
                    if (rec) rf_dbg[t * 16 + 12 + ph] = clock64();
                }
            }
        }
        __syncwarp();
    } else if (warp == 9) {
        // =================================== weight loader ==============================================================
        if (lane == 0) {
            const int per_step = 4 * M;                                  // planes per step: 2M gate (16 KB) + 2M candidate (8 KB)
            const uint8_t* wc = p.wimg + (size_t)2 * M * RF_WSLOT;
            unsigned pc = 0;
            for (int t = 0; t < T; ++t)
                for (int q = 0; q < per_step; ++q, ++pc) {
                    const int ws = pc % RF_NW;
                    if (pc >= RF_NW) mbar_wait_relaxed(&bar_wempty[ws], ((pc / RF_NW) - 1) & 1);
                    const bool gate = q < 2 * M;
                    const uint32_t bytes = gate ? RF_WSLOT : RF_WSLOT / 2;
                    const uint8_t* src = gate ? p.wimg + (size_t)q * RF_WSLOT : wc + (size_t)(q - 2 * M) * (RF_WSLOT / 2);
                    mbar_expect_tx(&bar_wfull[ws], bytes);
                    bulk_g2s(Wring + ws * RF_WSLOT, src, bytes, &bar_wfull[ws]);
                }
        }
        __syncwarp();
    } else if (warp == 10) {
        // =================================== operand-image dump ==========================================================
        if (dump) {
            const int plane = lane >> 2, s = lane & 3;
            if (lane == 0) tma_prefetch_desc(&tm_img);
            unsigned fills[3] = {0, 0, 0};
            for (int t = 0; t < T; ++t) {
                const long slab = (long)tile * p.img_T + p.img_t0 + t;
                for (int ci = 0; ci < 2 * M; ++ci) {
                    const int m = ci < M ? ci : ci - M;
                    const int slot = m == 0 ? 0 : 1 + ((m - 1) & 1);
                    mbar_wait_relaxed(&bar_afull[slot], fills[slot] & 1);
                    ++fills[slot];
                    if (lane < 2 * SB) {
                        tma_store_2d(&tm_img, p.img_col0 + 64 * ci, (int)((slab * 2 + plane) * IMG_ROWS + s * (RG * 8)),
                                     Aslots + slot * SLOT + plane * PLANE + s * (RP * 128));
                        bulk_commit();
                    }
                    bulk_wait_read();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_stored[slot]);
                }
            }
            bulk_wait_all();
        }
        __syncwarp();
    } else if (warp >= 11) {
        // =================================== XP loaders: thread = row ====================================================
        const int quad = warp & 3, row = 32 * quad + lane;
        const int s = row >> 5, n = row & 31, b = b0 + s;
        const bool valid = n < N && b < p.B;
        const uint32_t lane_base = (uint32_t)(32 * quad) << 16;
        for (int t = 0; t < T; ++t) {
            const int acc_i = t & 1;
            if (t >= 2) mbar_wait_relaxed(&bar_accfree[acc_i], ((t >> 1) - 1) & 1, 512);
            tc_fence_after();
            const float4* src = reinterpret_cast<const float4*>(p.xp + (((size_t)t * p.B + (valid ? b : 0)) * N + (valid ? n : 0)) * (3 * RF_H));
#pragma unroll 1
            for (int cb = 0; cb < 3 * RF_H; cb += 32) {
                float v[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (valid) q = __ldcs(src + cb / 4 + j);
                    v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
                }
                tmem_st32(taddr + lane_base + acc_i * RF_ACC1 + cb, v);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_xpfull[acc_i]);
        }
    } else {
        // =================================== workers =====================================================================
        const int quad = warp & 3, half = warp >> 2;
        const int row = 32 * quad + lane;
        const int b_ = b0 + quad;
        const bool rvalid = lane < N && b_ < p.B;
        const uint32_t lane_base = (uint32_t)(32 * quad) << 16;
        unsigned fills[3] = {0, 0, 0};                                  // fills STARTED per slot (identical in every thread)
        // before writing fill #f of a slot: the MMAs (and the image dump) of fill #f-1 have read it
        auto acquire = [&](int slot, bool need_mma) {
            const unsigned f = fills[slot];
            if (f >= 1) {
                if (need_mma) mbar_wait(&bar_aempty[slot], (f - 1) & 1);
                if (dump) mbar_wait(&bar_stored[slot], (f - 1) & 1);
            }
        };
        // ---- warp-private staging tile for coalesced global stores; lives in slots 1-2, which are idle in epilogue 2 ----
        float* stg = reinterpret_cast<float*>(Aslots + SLOT) + warp * (32 * 36);
        const int rq = lane >> 3, f4 = lane & 7;
        int grow[8];                                                    // (b*N + n) of the rows this lane moves, or -1
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int n = rq + 4 * i;
            grow[i] = (n < N && b_ < p.B) ? (b_ * N + n) : -1;
        }
        auto stage_store = [&](const float (&v)[32], float* base, int ld, int col0) {
            __syncwarp();
            float4* d = reinterpret_cast<float4*>(stg + lane * 36);
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (grow[i] >= 0)
                    *reinterpret_cast<float4*>(base + (size_t)grow[i] * ld + col0 + 4 * f4) =
                        *reinterpret_cast<const float4*>(stg + (rq + 4 * i) * 36 + 4 * f4);
        };
        // this thread's 32 state columns -> state tile row and slot 0 (hi / lo)
        auto put_state = [&](const float (&v)[32]) {
            float4* z4 = reinterpret_cast<float4*>(ZH + row * RF_ZLD + half * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) z4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 hi, lo;
                split8(&v[8 * j], hi, lo);
                const uint32_t off = k128_off(row, half * 32 + 8 * j);
                *reinterpret_cast<uint4*>(Aslots + off) = hi;
                *reinterpret_cast<uint4*>(Aslots + PLANE + off) = lo;
            }
        };
        auto put_state8 = [&](int col, const float (&v)[8]) {          // 8 columns of this row
            float4* z4 = reinterpret_cast<float4*>(ZH + row * RF_ZLD + col);
            z4[0] = make_float4(v[0], v[1], v[2], v[3]);
            z4[1] = make_float4(v[4], v[5], v[6], v[7]);
            uint4 hi, lo;
            split8(v, hi, lo);
            const uint32_t off = k128_off(row, col);
            *reinterpret_cast<uint4*>(Aslots + off) = hi;
            *reinterpret_cast<uint4*>(Aslots + PLANE + off) = lo;
        };
        auto publish_slot0 = [&]() {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_afull[0]);
            ++fills[0];
        };
        // diffusion of one phase: one warp per (sample, term): 64 columns, two per lane (each P^T row read from shared memory
        // feeds 40 FMAs: with one column per lane the loop was bound by the LDS issue rate at ~55 FMA/clk/SM, measured
        // with scripts/micro/fma_rate.cu).  M = 3: terms 1 and 2 run side by side on warps 0-3 / 4-7; M = 5: two rounds,
        // the MMAs of round one overlap the diffusion of round two.  A task arrives with count 2 (4 tasks = 8 arrivals).
        auto diffuse_phase = [&]() {
            for (int m = 1; m < M; ++m) {
                const int slot = 1 + ((m - 1) & 1);
                const int s = (warp - ((m - 1) * SB)) & 7;
                if (s < SB) {
                    acquire(slot, true);
                    const float* pt = PTs + (s * (M - 1) + (m - 1)) * PT_STRIDE;
                    if (p.mma_diff) {
                        // warp-level tensor path (f16_common.cuh::diffuse_mma16): the source is the term-0 chunk itself (slot 0,
                        // hi / lo fp16, just published by the epilogue); rows 0..23 of the sample's block are written
                        PFrag pf;
                        load_pfrag(pt, lane, pf);
                        diffuse_mma16(Aslots, s * RP, pf, Aslots + slot * SLOT, s * RP, lane, 1.f);
                    } else {
                        float acc[NPAD][2];
                        diffuse2(ZH + (s * RP) * RF_ZLD + 2 * lane, RF_ZLD, N, pt, acc);
                        // (rows N..23 are dumped to the operand image and aliased by the staging tiles: rewrite them as zeros)
                        store_cols2(Aslots + slot * SLOT, s * RP, lane, N, acc, 1.f, RG * 8);
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_n(&bar_afull[slot], 2);
                }
                ++fills[slot];
            }
        };
        {   // state <- h0
            float hp[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                if (rvalid) q = *reinterpret_cast<const float4*>(p.h0 + ((size_t)b_ * N + lane) * RF_H + half * 32 + 4 * j);
                hp[4 * j] = q.x; hp[4 * j + 1] = q.y; hp[4 * j + 2] = q.z; hp[4 * j + 3] = q.w;
            }
            put_state(hp);
            tmem_st32(taddr + lane_base + RF_STASH + half * 32, hp);
            publish_slot0();
            rf_worker_bar();
        }
        for (int t = 0; t < T; ++t) {
            const uint32_t dacc = taddr + lane_base + (t & 1) * RF_ACC1;
            float* hout = p.hseq + (size_t)t * p.B * NH;
            const bool rec = (p.dbg & 16) && blockIdx.x == 0 && tid == 0 && t < 64;
            long long* es = rf_dbg + t * 16;
            if (rec) es[0] = clock64();
            // ---- gate ------------------------------------------------------------------------------------------------
            diffuse_phase();
            if (rec) es[1] = clock64();
            mbar_wait(&bar_gate, t & 1);
            tc_fence_after();
            if (rec) es[2] = clock64();
            {   // epilogue 1: r = sigmoid(gate[:, 0:H]) -> r*h  (one 32-column TMEM load per thread: loads are paid per instruction)
                float rk[32], v[32];
                tmem_ld32(dacc + half * 32, rk);
                const float4* z4 = reinterpret_cast<const float4*>(ZH + row * RF_ZLD + half * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 z = z4[j];
                    rk[4 * j] = fast_sigmoid(rk[4 * j]); rk[4 * j + 1] = fast_sigmoid(rk[4 * j + 1]);
                    rk[4 * j + 2] = fast_sigmoid(rk[4 * j + 2]); rk[4 * j + 3] = fast_sigmoid(rk[4 * j + 3]);
                    v[4 * j] = rvalid ? rk[4 * j] * z.x : 0.f; v[4 * j + 1] = rvalid ? rk[4 * j + 1] * z.y : 0.f;
                    v[4 * j + 2] = rvalid ? rk[4 * j + 2] * z.z : 0.f; v[4 * j + 3] = rvalid ? rk[4 * j + 3] * z.w : 0.f;
                }
                acquire(0, false);                                      // (bar_gate covers the MMAs that read slot 0)
                put_state(v);                                           // ZH <- r*h, slot 0 <- hi/lo(r*h)
                publish_slot0();
                if (p.ruc) tmem_st32(taddr + lane_base + RF_RSTASH + half * 32, rk);   // r waits in TMEM for the stores of epilogue 2
            }
            if (rec) es[3] = clock64();
            rf_worker_bar();                                            // r*h of every row is visible
            if (rec) es[4] = clock64();
            // ---- candidate ---------------------------------------------------------------------------------------------
            diffuse_phase();
            if (rec) es[5] = clock64();
            mbar_wait(&bar_cand, t & 1);
            tc_fence_after();
            if (rec) es[6] = clock64();
            {   // epilogue 2: 64 columns, each warp half takes 32.  Everything it needs is on chip: the candidate and update-gate
                // pre-activations in the accumulator, h_{t-1} and r in the TMEM stashes.
                float cv[32], uv[32], hp[32];
                tmem_ld32_nw(dacc + 2 * RF_H + half * 32, cv);
                tmem_ld32_nw(dacc + RF_H + half * 32, uv);
                tmem_ld32_nw(taddr + lane_base + RF_STASH + half * 32, hp);
                tmem_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_accfree[t & 1]);        // the accumulator may take XP_{t+2}
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float c, u;
                    if (p.act == 0) tanh_sigmoid(cv[j], uv[j], c, u);
                    else { c = fmaxf(cv[j], 0.f); u = fast_sigmoid(uv[j]); }
                    cv[j] = c; uv[j] = u;
                    hp[j] = rvalid ? u * hp[j] + (1.f - u) * c : 0.f;
                }
                acquire(0, false);
                put_state(hp);                                          // ZH <- h_t, slot 0 <- hi/lo(h_t): term 0 of the next gate
                publish_slot0();
                tmem_st32(taddr + lane_base + RF_STASH + half * 32, hp);
                if (rec) es[7] = clock64();
                // the staging tiles alias slots 1-2: their last chunks must have been read by the MMAs (bar_cand) and dumped
                if (dump && M > 1) {
                    if (fills[1] >= 1) mbar_wait(&bar_stored[1], (fills[1] - 1) & 1);
                    if (fills[2] >= 1) mbar_wait(&bar_stored[2], (fills[2] - 1) & 1);
                }
                stage_store(hp, hout, RF_H, half * 32);
                if (p.ruc) {
                    float* ruc = p.ruc + (size_t)t * p.B * NH * 3;
                    stage_store(uv, ruc, 3 * RF_H, RF_H + half * 32);
                    stage_store(cv, ruc, 3 * RF_H, 2 * RF_H + half * 32);
                    tmem_ld32(taddr + lane_base + RF_RSTASH + half * 32, uv);
                    stage_store(uv, ruc, 3 * RF_H, half * 32);
                }
            }
            if (rec) es[8] = clock64();
            rf_worker_bar();                                            // h_t of every row is visible; staging is free again
            if (rec) es[9] = clock64();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(taddr);
}

cudaError_t rnn_fwd_read_dbg(long long* out, int n) {
    return cudaMemcpyFromSymbol(out, rf_dbg, sizeof(long long) * (n < 1536 ? n : 1536));
}
size_t rnn_fwd_wimg_bytes(int M) { return (size_t)2 * M * (RF_WSLOT + RF_WSLOT / 2); }
int rnn_fwd_smem_bytes(int M) { return RF_OFF_PT + SB * (M - 1) * PT_STRIDE * 4 + 1024; }
bool rnn_fwd_supported(int N, int H, int M, int smem_limit) {
    return H == RF_H && N <= NPAD && M >= 1 && M <= 7 && rnn_fwd_smem_bytes(M) + 1024 <= smem_limit;
}

cudaError_t rnn_fwd_pack_weights(const float* Wg, const float* Wc, int fin, int M, void* wimg, cudaStream_t st) {
    cudaError_t e = launch_pack_w16(Wg, Wc, fin, RF_H, M, 2, 2 * RF_H, M, wimg, st);
    if (e != cudaSuccess) return e;
    return launch_pack_w16(Wg, Wc, fin, RF_H, M, 3, RF_H, M, reinterpret_cast<uint8_t*>(wimg) + (size_t)2 * M * RF_WSLOT, st);
}

// wimg: rnn_fwd_wimg_bytes(M), filled here; img: operand image base or nullptr (img_cols fp16 values per row)
cudaError_t launch_rnn_fwd(int B, int T, int N, int fin, int M, int act, const float* xp, const float* h0, const float* P,
                           const float* Wg, const float* Wc, void* wimg, float* hseq, float* ruc, void* img, int img_cols,
                           int img_col0, cudaStream_t st, int img_T, int img_t0) {
    cudaError_t e = cudaSuccess;
    if (Wg) {
        e = rnn_fwd_pack_weights(Wg, Wc, fin, M, wimg, st);
        if (e != cudaSuccess) return e;
    }
    RnnFwdParams p;
    memset(&p, 0, sizeof p);
    p.B = B; p.T = T; p.N = N; p.M = M; p.act = act; p.dump = img != nullptr; p.img_col0 = img_col0;
    p.img_T = img_T > 0 ? img_T : T; p.img_t0 = img_T > 0 ? img_t0 : 0;
    { const char* e = getenv("DCGRU_DBG"); p.dbg = e ? (atoi(e) & (16 | 64 | 128 | 256 | 512)) : 0; }
    { const char* e = getenv("DCGRU_MMA_DIFF"); p.mma_diff = !(e && e[0] == '0'); }
    p.xp = xp; p.h0 = h0; p.P = P; p.wimg = reinterpret_cast<const uint8_t*>(wimg); p.hseq = hseq; p.ruc = ruc;
    CUtensorMap tm;
    memset(&tm, 0, sizeof tm);
    const int ntile = g16_ntile(B);
    if (img) {
        const unsigned long long dims[2] = {(unsigned long long)img_cols, (unsigned long long)ntile * p.img_T * 2 * IMG_ROWS};
        const unsigned long long str[2] = {2, (unsigned long long)img_cols * 2};
        const unsigned box[2] = {64, RG * 8};
        e = make_tmap(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, img, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (e != cudaSuccess) return e;
    }
    const int smem = rnn_fwd_smem_bytes(M);
    e = cudaFuncSetAttribute(rnn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    rnn_fwd_kernel<<<ntile, RF_THREADS, smem, st>>>(p, tm);
    return cudaGetLastError();
}

}  // namespace dcgru
