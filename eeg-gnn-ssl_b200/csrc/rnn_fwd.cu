// Recurrent part of one encoder layer on the tensor cores (tcgen05, 2xFP16): all T steps for a tile of 4 samples per
// CTA, no relaunch (model/model.py:93-96 x model/cell.py:182-210).  The non-recurrent x-part of both projections,
// incl. the biases, was hoisted into one bulk GEMM over all steps (bulk_dp.cu) and arrives as the pre-activation
// XP (T,B,N,3H); per step this kernel only does what depends on h_{t-1}:
//     gate : D[:, 0:2H]  = XP_t[:, 0:2H]  + sum_m (P_m h_{t-1})      @ Wg_h,m     -> r = sigmoid(D[:, 0:H])
//     cand : D[:, 2H:3H] = XP_t[:, 2H:3H] + sum_m (P_m (r*h_{t-1}))  @ Wc_h,m     -> c = act(.), u = sigmoid(D[:, H:2H])
//     h_t = u*h_{t-1} + (1-u)*c
// K order is term-major (kk = m*H + c), so with H = 64 one chunk (f16_common.cuh) = one diffusion term:
//   * term 0 (the state itself) is written as hi/lo fp16 by the epilogue that produced it (slot 0),
//   * terms m >= 1 by the worker warps, one term at a time (warp = (sample, column half)): P_m applied to the term-0
//     chunk on the warp-level tensor path (mma.sync 2xFP16, f16_common.cuh::diffuse_mma16), slots 1 and 2 alternate; the
//     tcgen05 MMAs of term m run while term m+1 is being diffused.
// XP_t is loaded straight into the accumulator (TMEM) two steps ahead by four loader warps and the MMAs accumulate on
// top of it; the accumulator is double buffered, the GRU state lives in a TMEM stash.  The workers never touch global
// memory inside the loop: the epilogues write r, u, c back over their pre-activations in the accumulator and h_t into
// the stash, and the loader warps move them to h_seq / ruc (TMEM -> warp-private staging tile -> 128-byte rows) while
// the workers are already in the next step.
//   warps 0-7   workers: diffusion tasks + the two epilogues (thread = (row, column half))
//   warp 8      MMA issuer;  warp 9 weight loader (cp.async.bulk ring of hi / lo planes, L2 resident)
//   warp 10     operand-image dump (tensor-map TMA store of every A chunk: columns [KXP, KXP + 2*M*H) of the image)
//   warps 11-14 output stores of step t, then XP loads of step t+2 (thread = row)
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
constexpr int RF_STG_LD = 36;                       // staging tile row stride (floats)
constexpr int RF_STG = 32 * RF_STG_LD * 4;          // one warp-private staging tile: 32 rows x 32 columns
constexpr int RF_NW = 4;                            // weight ring slots
constexpr int RF_WSLOT = 16 * 1024;                 // one plane of a gate chunk (128 rows x 128 B); candidate planes are 8 KB
constexpr int RF_OFF_A = 0;                         // 3 chunk slots
constexpr int RF_OFF_W = 3 * SLOT;
constexpr int RF_OFF_STG = RF_OFF_W + RF_NW * RF_WSLOT;          // 4 staging tiles (loader warps)
constexpr int RF_OFF_PT = RF_OFF_STG + 4 * RF_STG;
constexpr int RF_ACC1 = 192, RF_STASH = 384;                     // TMEM columns: accumulators at 0 / 192, h stash

struct RnnFwdParams {
    int B, T, N, M, act, dump, img_col0, dbg;
    int img_T, img_t0;          // slab of step t in the operand image: tile * img_T + img_t0 + t
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
    __shared__ uint64_t bar_xpfull[2], bar_final[2], bar_hfree, bar_gate, bar_cand;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = p.N, M = p.M, T = p.T;
    const int tile = blockIdx.x, b0 = tile * SB;
    uint8_t* Aslots = smem + RF_OFF_A;
    uint8_t* Wring = smem + RF_OFF_W;
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
        for (int i = 0; i < 2; ++i) { mbar_init(&bar_xpfull[i], 4); mbar_init(&bar_final[i], RF_NWORK / 32); }
        mbar_init(&bar_hfree, 4);
        mbar_init(&bar_gate, 1);
        mbar_init(&bar_cand, 1);
        mbar_fence_init();
    }
    for (int i = tid; i < 3 * SLOT / 16; i += RF_THREADS) reinterpret_cast<uint4*>(Aslots)[i] = make_uint4(0, 0, 0, 0);
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
        // =================================== output stores + XP loads: thread = row ======================================
        const int quad = warp & 3, row = 32 * quad + lane;
        const int s = row >> 5, n = row & 31, b = b0 + s;
        const bool valid = n < N && b < p.B;
        const uint32_t lane_base = (uint32_t)(32 * quad) << 16;
        float* stg = reinterpret_cast<float*>(smem + RF_OFF_STG + (warp - 11) * RF_STG);
        const int rq = lane >> 3, f4 = lane & 7;
        int grow[8];                                                    // (b*N + n) of the rows this lane moves, or -1
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int nn = rq + 4 * i;
            grow[i] = (nn < N && b < p.B) ? (b * N + nn) : -1;
        }
        // 32 columns of this warp's 32 rows: TMEM -> staging tile -> whole 128-byte row pieces in global memory
        auto move32 = [&](uint32_t tcol, float* base, int ld, int col0) {
            float v[32];
            tmem_ld32(taddr + lane_base + tcol, v);
            __syncwarp();
            float4* d = reinterpret_cast<float4*>(stg + lane * RF_STG_LD);
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (grow[i] >= 0)
                    __stcs(reinterpret_cast<float4*>(base + (size_t)grow[i] * ld + col0 + 4 * f4),
                           *reinterpret_cast<const float4*>(stg + (rq + 4 * i) * RF_STG_LD + 4 * f4));
        };
        auto load_xp = [&](int t) {
            const int acc_i = t & 1;
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
        };
        load_xp(0);
        if (T > 1) load_xp(1);
        for (int t = 0; t < T; ++t) {
            const int acc_i = t & 1;
            mbar_wait_relaxed(&bar_final[acc_i], (t >> 1) & 1, 512);    // r | u | c sit in the accumulator, h_t in the stash
            tc_fence_after();
            move32(RF_STASH, p.hseq + (size_t)t * p.B * NH, RF_H, 0);
            move32(RF_STASH + 32, p.hseq + (size_t)t * p.B * NH, RF_H, 32);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_hfree);                     // the stash may take h_{t+1}
            if (p.ruc) {
                float* ruc = p.ruc + (size_t)t * p.B * NH * 3;
#pragma unroll 1
                for (int cb = 0; cb < 3 * RF_H; cb += 32) move32(acc_i * RF_ACC1 + cb, ruc, 3 * RF_H, cb);
            }
            if (t + 2 < T) load_xp(t + 2);
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
        // this thread's 32 state columns -> slot 0 (hi / lo): term 0 for the tensor core and source of the diffusion
        auto put_state = [&](const float (&v)[32]) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 hi, lo;
                split8(&v[8 * j], hi, lo);
                const uint32_t off = k128_off(row, half * 32 + 8 * j);
                *reinterpret_cast<uint4*>(Aslots + off) = hi;
                *reinterpret_cast<uint4*>(Aslots + PLANE + off) = lo;
            }
        };
        auto publish_slot0 = [&]() {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_afull[0]);
            ++fills[0];
        };
        // diffusion of one phase, one term at a time: warp = (sample quad, column half), two 16-column groups per task on
        // the warp-level tensor path (f16_common.cuh::diffuse_mma16); the source is the term-0 chunk itself (slot 0, hi / lo
        // fp16, just published by the epilogue).  The tcgen05 MMAs of term m overlap the diffusion of term m+1.
        auto diffuse_phase = [&](long long* st) {
            for (int m = 1; m < M; ++m) {
                const int slot = 1 + ((m - 1) & 1);
                if (st && m == 1) st[0] = clock64();
                // terms 1 and 2 reuse the slots of the previous phase's last two terms: their MMAs completed before bar_gate /
                // bar_cand, which this thread has already waited for -> only the image dump has to be awaited (a try_wait costs
                // ~200 cycles even when the phase completed long ago)
                acquire(slot, m > 2);
                if (st && m == 1) st[1] = clock64();
                PFrag pf;
                load_pfrag(PTs + (quad * (M - 1) + (m - 1)) * PT_STRIDE, lane, pf);
                if (st && m == 1) st[2] = clock64();
                diffuse_mma16(Aslots, quad * RP, pf, Aslots + slot * SLOT, quad * RP, lane, 1.f, 2 * half, 2 * half + 2);
                if (st && m == 1) st[3] = clock64();
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_afull[slot]);
                ++fills[slot];
                if (st && m == 1) st[4] = clock64();
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
            const bool rec = (p.dbg & 16) && blockIdx.x == 0 && tid == 0 && t < 64;
            long long* es = rf_dbg + t * 16;
            if (rec) es[0] = clock64();
            // ---- gate ------------------------------------------------------------------------------------------------
            diffuse_phase(rec ? rf_dbg + 1024 + t * 8 : nullptr);
            if (rec) es[1] = clock64();
            mbar_wait(&bar_gate, t & 1);
            tc_fence_after();
            if (rec) es[2] = clock64();
            {   // epilogue 1: r = sigmoid(gate[:, 0:H]) -> r*h  (h_{t-1} from the TMEM stash)
                float rk[32], v[32];
                tmem_ld32_nw(dacc + half * 32, rk);
                tmem_ld32_nw(taddr + lane_base + RF_STASH + half * 32, v);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    rk[j] = fast_sigmoid(rk[j]);
                    v[j] = rvalid ? rk[j] * v[j] : 0.f;
                }
                acquire(0, false);                                      // (bar_gate covers the MMAs that read slot 0)
                put_state(v);                                           // slot 0 <- hi/lo(r*h)
                publish_slot0();
                if (p.ruc) tmem_st32(dacc + half * 32, rk);             // r replaces its pre-activation: the loader warps store it
            }
            if (rec) es[3] = clock64();
            // (no CTA-wide barrier: a warp diffuses exactly the rows and columns of slot 0 that it wrote itself)
            if (rec) es[4] = clock64();
            // ---- candidate ---------------------------------------------------------------------------------------------
            diffuse_phase(nullptr);
            if (rec) es[5] = clock64();
            mbar_wait(&bar_cand, t & 1);
            tc_fence_after();
            if (rec) es[6] = clock64();
            {   // epilogue 2: 64 columns, each warp half takes 32.  Everything it needs is on chip: the candidate and update-gate
                // pre-activations in the accumulator, h_{t-1} in the TMEM stash.
                float cv[32], uv[32], hp[32];
                tmem_ld32_nw(dacc + 2 * RF_H + half * 32, cv);
                tmem_ld32_nw(dacc + RF_H + half * 32, uv);
                tmem_ld32_nw(taddr + lane_base + RF_STASH + half * 32, hp);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float c, u;
                    if (p.act == 0) tanh_sigmoid(cv[j], uv[j], c, u);
                    else { c = fmaxf(cv[j], 0.f); u = fast_sigmoid(uv[j]); }
                    cv[j] = c; uv[j] = u;
                    hp[j] = rvalid ? u * hp[j] + (1.f - u) * c : 0.f;
                }
                acquire(0, false);
                put_state(hp);                                          // slot 0 <- hi/lo(h_t): term 0 of the next gate
                publish_slot0();
                if (rec) es[7] = clock64();
                if (t >= 1) mbar_wait(&bar_hfree, (t - 1) & 1);         // the loader warps have read h_{t-1} (long ago)
                tc_fence_after();
                tmem_st32(taddr + lane_base + RF_STASH + half * 32, hp);
                if (p.ruc) {                                            // u, c replace their pre-activations
                    tmem_st32(dacc + RF_H + half * 32, uv);
                    tmem_st32(dacc + 2 * RF_H + half * 32, cv);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_final[t & 1]);          // outputs of step t are complete: loader warps take over
            }
            if (rec) es[8] = clock64();
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
