// Persistent encoder-layer forward on the tensor cores (tcgen05, 3xTF32): all T steps of one layer for a
// group of 4 samples per CTA, no relaunch (model/model.py:93-96 x model/cell.py:182-210).
//
// Per step the three projections of the cell run as UMMA GEMMs with M = 128 rows (4 samples x 32 rows, one
// warp per sample: see tc_common.cuh), accumulating in one 128 x 192 fp32 TMEM tile:
//     X : D[:, 0:192] (=)  diffuse(x_t)      @ [Wg_x | Wc_x]     (K = Fin*M)
//     Hg: D[:, 0:128] (+)= diffuse(h_{t-1})  @  Wg_h             (K = H*M)      -> r = sigmoid(D[:,0:64] + bg)  (u: later)
//     Hc: D[:,128:192](+)= diffuse(r*h_{t-1})@  Wc_h             (K = H*M)      -> c = act(. + bc), u = sigmoid(D[:,64:128] + bg), GRU update
// A operand: each thread owns one (sample, node) row and keeps that row of the diffusion polynomials P_m
// in registers for the whole sequence; for every 8-column chunk of [x | h] it forms the M diffusion terms
// of 4 columns (20 broadcast float4 reads + 40 FMAs per column quad), splits them hi/lo and writes them
// in kk = c*M + m order into a UMMA tile (2 stages) laid out [row group of 8][K-group pair][8 rows x 32 B]
// (K-major, 32-byte swizzle: one MMA k-step = one 256-byte atom per row group, SBO = 768 B between row groups).
// Optional operand image (gsave): a dump warp copies every finished A stage (hi and lo) to HBM with one
// tensor-map TMA store each (a 4-D box that scatters the 32-byte row pieces of the tile) into the row-major image
// G[cta*T + t][hi|lo][96 rows = sample*24 + node][KKP floats]: the diffused operands [x | h | r*h] in kk order.
// The weight-gradient GEMM (dw_mm.cu) reads that image back with swizzling TMA loads as an MN-major UMMA
// operand, so dW needs no recomputation of the diffusion at all.
// B operand: the weights, pre-split and pre-tiled once per launch by pack_w_fwd_kernel, streamed chunk by
// chunk from L2 with cp.async.bulk (TMA) into a 3-slot ring, each load issued one chunk ahead at the top
// of the iteration so its latency hides behind two A-tile productions; mbarriers track "weights landed"
// and "MMAs done".  x_t is streamed 8 columns at a time with cp.async into a 4-slot ring (two chunks ahead);
// the hidden state stays resident in shared memory (r*h overwrites it for the candidate) and in a TMEM stash.
// The epilogues read the accumulator with tcgen05.ld (thread = row) and fuse bias, sigmoid/tanh, r*h and
// the GRU update.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "dw.cuh"
#include "tc_common.cuh"
#include "tmap.cuh"

namespace dcgru {
using namespace tc;

constexpr int FT_SB = TC_SB;             // samples per CTA
constexpr int FT_RP = TC_RP;             // rows per sample
constexpr int FT_ROWS = 128;
constexpr int FT_CC = 8;                 // source columns per chunk
constexpr int FT_H = 64;
constexpr int FT_M = 3;
constexpr int FT_KK = FT_CC * FT_M;      // 24 K values per chunk = 3 MMA k-steps
constexpr int FT_KG = FT_KK / 4;         // 6 K groups
constexpr int FT_A_BYTES = FT_KG * FT_ROWS * 16;          // one of hi / lo
constexpr int FT_A_STAGE = 2 * FT_A_BYTES;
constexpr int FT_BX_BYTES = 2 * FT_KG * 192 * 16;         // hi + lo, N = 192
constexpr int FT_BG_BYTES = 2 * FT_KG * 128 * 16;
constexpr int FT_BC_BYTES = 2 * FT_KG * 64 * 16;
constexpr int FT_XLD = 12;                                  // x chunk row stride (floats): 8 columns + pad, conflict-free float4 rows
constexpr int FT_ZLD = FT_H + 4;                            // hidden-state row stride (floats), same reason
constexpr int FT_XSLOT = FT_ROWS * FT_XLD * 4;             // one x chunk: [128 rows][8 cols (+4 pad)]
// shared memory map
constexpr int FT_OFF_A = 0;                                // 2 stages
constexpr int FT_OFF_B = FT_OFF_A + 2 * FT_A_STAGE;        // 3 slots
constexpr int FT_OFF_X = FT_OFF_B + 3 * FT_BX_BYTES;       // 3 slots
constexpr int FT_XRING = 4;                                // x chunk ring slots (prefetch distance 2, race-free)
constexpr int FT_OFF_ZH = FT_OFF_X + FT_XRING * FT_XSLOT;  // [128][64] hidden state (or r*h)
constexpr int FT_SMEM = FT_OFF_ZH + FT_ROWS * FT_ZLD * 4 + 1024;   // + slack: the swizzled tiles need an aligned base

__host__ __device__ inline int ft_nxc(int fin) { return (fin + FT_CC - 1) / FT_CC; }
__host__ __device__ inline size_t ft_wimg_bytes(int fin) {
    return (size_t)ft_nxc(fin) * FT_BX_BYTES + (size_t)(FT_H / FT_CC) * (FT_BG_BYTES + FT_BC_BYTES);
}

// ---- weight image: chunk blocks in the order the kernel consumes them, each [hi: kg][n] [lo: kg][n] float4 ----
// block kinds: X chunk i (n < 128 from Wg, n >= 128 from Wc, rows (8i+cc)*M+m), then Hg chunks, then Hc chunks
__global__ void pack_w_fwd_kernel(const float* Wg, const float* Wc, int fin, float* img) {
    const int nxc = ft_nxc(fin), nhc = FT_H / FT_CC;
    const int blk = blockIdx.x;
    int kind, ci, N;
    size_t off;
    if (blk < nxc) { kind = 0; ci = blk; N = 192; off = (size_t)blk * FT_BX_BYTES; }
    else if (blk < nxc + nhc) { kind = 1; ci = blk - nxc; N = 128; off = (size_t)nxc * FT_BX_BYTES + (size_t)ci * FT_BG_BYTES; }
    else { kind = 2; ci = blk - nxc - nhc; N = 64;
           off = (size_t)nxc * FT_BX_BYTES + (size_t)nhc * FT_BG_BYTES + (size_t)ci * FT_BC_BYTES; }
    float4* hi = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(img) + off);
    float4* lo = hi + FT_KG * N;
    const int cbase = (kind == 0) ? ci * FT_CC : fin + ci * FT_CC;       // first source column of the chunk
    const int cend = (kind == 0) ? fin : fin + FT_H;
    for (int idx = threadIdx.x; idx < FT_KG * N; idx += blockDim.x) {
        const int kg = idx / N, n = idx - kg * N;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int kl = kg * 4 + e;                                   // local kk: (cc, m)
            const int cc = kl / FT_M, m = kl - cc * FT_M;
            const int c = cbase + cc;
            float w = 0.f;
            if (c < cend) {
                const size_t row = (size_t)c * FT_M + m;
                if (kind == 0) w = (n < 128) ? Wg[row * 128 + n] : Wc[row * 64 + (n - 128)];
                else if (kind == 1) w = Wg[row * 128 + n];
                else w = Wc[row * 64 + n];
            }
            v[e] = w;
        }
        float4 h, l;
        split4(make_float4(v[0], v[1], v[2], v[3]), h, l);
        hi[idx] = h;
        lo[idx] = l;
    }
}

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct FwdTcParams {
    int B, T, N, fin, act, dbg;
    const float* x; long long xs_t, xs_b;
    const float* h0;
    const float* P;
    const float* bg; const float* bc;
    const float* wimg;
    float* hseq;
    float* ruc;
    uint8_t* gsave;             // operand image for dw_mm (nullptr: not saved)
    long long* dbgbuf;          // timing experiment (DCGRU_DBG & 4): clock64 stamps of CTA 0
};

__device__ __forceinline__ void bulk_copy(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void producer_barrier() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }
// fast gates for the tensor-core path: ex2.approx based, ~1e-6 absolute error (the fp32 FMA path keeps expf/tanhf)
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

constexpr int FT_NPROD = 256;            // producer / epilogue threads (warps 0-7)
constexpr int FT_STASH = 192;            // TMEM columns 192..255: h_{t-1}, thread = row (the accumulator uses 0..191)
constexpr int FT_THREADS = 352;          // + warp 8: MMA issue, warp 9: TMA weight loads, warp 10: operand-image dump
constexpr int FT_RG_F4 = FT_KG * 8;      // float4s per 8-row group of an A tile (768 B)
__device__ __forceinline__ int ft_a_idx(int kg, int row) { return k32_idx<FT_KG / 2>(kg, row); }

// Warp-specialised: warps 0-7 build the A tiles and run the epilogues, warp 8 streams weights and x with TMA
// and issues the MMAs; the only synchronisation inside a step is through mbarriers (no CTA-wide barrier).
__global__ void __launch_bounds__(FT_THREADS, 1) seq_fwd_tc_kernel(const FwdTcParams p,
                                                                   const __grid_constant__ CUtensorMap tm_g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar_bfull[3], bar_xfull[FT_XRING], bar_afull[2], bar_done[2], bar_stored[2], bar_epi;
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float sbias[3 * FT_H];                      // bg (r | u) | bc
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = p.N, fin = p.fin;
    const int b0 = blockIdx.x * FT_SB;
    if (tid < 3 * FT_H) sbias[tid] = (tid < 2 * FT_H) ? p.bg[tid] : p.bc[tid - 2 * FT_H];
    float* ZH = reinterpret_cast<float*>(smem + FT_OFF_ZH);              // [128][FT_ZLD]

    if (warp == 0) tmem_alloc<256>(&tmem_slot);
    if (tid == 0) {
        for (int i = 0; i < 3; ++i) mbar_init(&bar_bfull[i], 1);
        for (int i = 0; i < FT_XRING; ++i) mbar_init(&bar_xfull[i], FT_NPROD);
        for (int i = 0; i < 2; ++i) { mbar_init(&bar_afull[i], FT_NPROD / 32); mbar_init(&bar_done[i], 1); mbar_init(&bar_stored[i], 1); }
        mbar_init(&bar_epi, FT_NPROD / 32);
        mbar_fence_init();
    }
    // hidden state <- h0, x ring <- 0 (rows of pad nodes / missing samples stay zero for ever)
    for (int idx = tid; idx < FT_ROWS * (FT_H / 4); idx += FT_THREADS) {
        const int r = idx / (FT_H / 4), c4 = (idx - r * (FT_H / 4)) * 4;
        const int s = r / FT_RP, n = r - s * FT_RP, b = b0 + s;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < N && b < p.B) v = *reinterpret_cast<const float4*>(p.h0 + ((size_t)b * N + n) * FT_H + c4);
        *reinterpret_cast<float4*>(ZH + r * FT_ZLD + c4) = v;
    }
    for (int idx = tid; idx < FT_XRING * FT_XSLOT / 16; idx += FT_THREADS)
        reinterpret_cast<float4*>(smem + FT_OFF_X)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_async_smem();                      // the zeros must be ordered before the TMA writes into the ring
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = tmem_slot;

    const int nxc = ft_nxc(fin), nhc = FT_H / FT_CC;
    const int per_step = nxc + 2 * nhc;
    const unsigned total_chunks = (unsigned)p.T * per_step;
    const unsigned total_x = (unsigned)p.T * nxc;
    const uint8_t* wimg = reinterpret_cast<const uint8_t*>(p.wimg);
    const size_t NH = (size_t)N * FT_H;
    const bool dump = p.gsave != nullptr;

    if (warp == 8) {
        // =================================== TMA + MMA issuer ===================================================
        // One thread issues every MMA, so its instruction stream is the pipeline's clock: all shared-memory
        // descriptors are precomputed (a first version rebuilt them per MMA and, with integer divisions for the
        // chunk bookkeeping, spent ~2500 cycles per chunk here -- measured with clock64 stamps).
        __shared__ uint64_t dA[2][3][2];                  // [stage][k-step][hi, lo]
        __shared__ uint64_t dB[3][3][3][2];               // [slot][kind: X, Hg, Hc][k-step][hi, lo]
        if (lane == 0) {
            for (int st = 0; st < 2; ++st)
                for (int k = 0; k < 3; ++k) {
                    const uint32_t hi = smem_u32(smem + FT_OFF_A + st * FT_A_STAGE) + k * 256;
                    dA[st][k][0] = make_smem_desc_k32(hi, FT_RG_F4 * 16);
                    dA[st][k][1] = make_smem_desc_k32(hi + FT_A_BYTES, FT_RG_F4 * 16);
                }
            for (int sl = 0; sl < 3; ++sl)
                for (int kind = 0; kind < 3; ++kind) {
                    const int ncols = kind == 0 ? 192 : (kind == 1 ? 128 : 64);
                    for (int k = 0; k < 3; ++k) {
                        const uint32_t hi = smem_u32(smem + FT_OFF_B + sl * FT_BX_BYTES) + 2 * k * ncols * 16;
                        dB[sl][kind][k][0] = make_smem_desc(hi, ncols * 16, 128);
                        dB[sl][kind][k][1] = make_smem_desc(hi + FT_KG * ncols * 16, ncols * 16, 128);
                    }
                }
        }
        __syncwarp();
        const uint32_t idesc[3] = {make_idesc_tf32(128, 192), make_idesc_tf32(128, 128), make_idesc_tf32(128, 64)};
        if (lane == 0) {
            int q = 0, t = 0;                 // chunk within step / step
            int sb = 0, kb = 0;               // weight slot of chunk g and its use count parity
            for (unsigned g = 0; g < total_chunks; ++g) {
                const int sa = g & 1;
                const bool rec = (p.dbg & 4) && blockIdx.x == 0 && g < 128;
                if (rec) p.dbgbuf[g * 8 + 0] = clock64();
                if (rec) p.dbgbuf[g * 8 + 1] = clock64();
                // one combined poll (weights landed, A tile built): this thread issues MMAs almost synchronously (the
                // tensor core queues only ~2 of them), so every cycle it spends elsewhere is a cycle the pipe drains
                // (a non-blocking look-ahead at the next chunk's barriers from inside the issue loop was tried: slower)
                mbar_wait2(&bar_bfull[sb], kb, &bar_afull[sa], (g >> 1) & 1);
                if (rec) p.dbgbuf[g * 8 + 2] = clock64();
                if (q == 0 && t > 0) mbar_wait(&bar_epi, (t - 1) & 1);   // the previous step's TMEM reads are done
                tc_fence_after();
                const int kind = (q < nxc) ? 0 : (q < nxc + nhc ? 1 : 2);
                const uint32_t d = taddr + (kind == 2 ? 128u : 0u);
                const uint32_t id = idesc[kind];
                uint32_t acc = (q != 0) ? 1u : 0u;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const uint64_t ah = dA[sa][k][0], al = dA[sa][k][1];
                    const uint64_t bh = dB[sb][kind][k][0], bl = dB[sb][kind][k][1];
                    umma_tf32(d, al, bh, id, acc);      // small terms first
                    umma_tf32(d, ah, bl, id, 1u);
                    umma_tf32(d, ah, bh, id, 1u);
                    acc = 1u;
                }
                umma_commit(&bar_done[sa]);
                if (rec) p.dbgbuf[g * 8 + 3] = clock64();
                // advance the bookkeeping without divisions
                if (++q == per_step) { q = 0; ++t; }
                if (++sb == 3) { sb = 0; kb ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == 9) {
        // =================================== TMA weight loader ==================================================
        // chunk g+2's weights go into the ring slot chunk g-1 used, as soon as that chunk's MMAs have completed
        const uint8_t* wg_base = wimg + (size_t)nxc * FT_BX_BYTES;
        const uint8_t* wc_base = wg_base + (size_t)nhc * FT_BG_BYTES;
        auto load_w = [&](int q, int slot) {                              // q = chunk index within a step
            const uint8_t* src; uint32_t bytes;
            if (q < nxc) { src = wimg + (size_t)q * FT_BX_BYTES; bytes = FT_BX_BYTES; }
            else if (q < nxc + nhc) { src = wg_base + (size_t)(q - nxc) * FT_BG_BYTES; bytes = FT_BG_BYTES; }
            else { src = wc_base + (size_t)(q - nxc - nhc) * FT_BC_BYTES; bytes = FT_BC_BYTES; }
            mbar_expect_tx(&bar_bfull[slot], bytes);
            bulk_copy(smem + FT_OFF_B + slot * FT_BX_BYTES, src, bytes, &bar_bfull[slot]);
        };
        if (lane == 0) {
            load_w(0, 0);
            if (total_chunks > 1) load_w(1 % per_step, 1);
            int q2 = 2 % per_step, sb2 = 2;   // chunk g+2: index within step, slot
            for (unsigned g = 0; g + 2 < total_chunks; ++g) {
                if (g >= 1) mbar_wait(&bar_done[(g - 1) & 1], ((g - 1) >> 1) & 1);       // frees weight slot sb2
                load_w(q2, sb2);
                if (++q2 == per_step) q2 = 0;
                if (++sb2 == 3) sb2 = 0;
            }
        }
        __syncwarp();
    } else if (warp == 10) {
        // =================================== operand-image dump ==================================================
        // lane = (hi|lo, sample): one TMA tensor store per finished A stage part and sample: box (8 kk, 8 rows,
        // 3 K-group pairs, 3 row groups) = the 24 rows of the sample that can be non-zero, in shared-memory order,
        // scattered into the row-major image.  The stage is released to the producers (bar_stored) once the
        // stores have read their shared-memory source.
        if (dump) {
            const int part = lane >> 2, s = lane & 3;
            if (lane == 0) tma_prefetch_desc(&tm_g);
            int q = 0;
            int rg0 = ((blockIdx.x * p.T * 2 + part) * FT_SB + s) * TC_RG;  // image row group of (cta, t = 0, part, sample)
            for (unsigned g = 0; g < total_chunks; ++g) {
                const int sa = g & 1;
                const bool rec = (p.dbg & 4) && blockIdx.x == 0 && lane == 0 && g < 128;
                mbar_wait(&bar_afull[sa], (g >> 1) & 1);
                if (rec) p.dbgbuf[1024 + g * 4 + 0] = clock64();
                if (lane < 2 * FT_SB) {
                    const uint8_t* src = smem + FT_OFF_A + sa * FT_A_STAGE + part * FT_A_BYTES + s * (4 * FT_RG_F4 * 16);
                    tma_store_4d(&tm_g, 0, 0, q * (FT_KG / 2), rg0, src);
                    bulk_commit();
                }
                if (rec) p.dbgbuf[1024 + g * 4 + 1] = clock64();
                bulk_wait_read();
                __syncwarp();
                if (rec) p.dbgbuf[1024 + g * 4 + 2] = clock64();
                if (lane == 0) mbar_arrive(&bar_stored[sa]);
                if (++q == per_step) { q = 0; rg0 += 2 * FT_SB * TC_RG; }
            }
            bulk_wait_all();
        }
        __syncwarp();
    } else {
        // =================================== producers / epilogue ================================================
        const int row = tid & 127, half = tid >> 7;
        const int s_ = row / FT_RP, n_ = row - s_ * FT_RP;
        const int b_ = b0 + s_;
        const bool rvalid = (n_ < N) && (b_ < p.B);
        // this row of the diffusion polynomials, kept in registers for the whole sequence
        float P1[NP], P2[NP];
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            P1[j] = 0.f; P2[j] = 0.f;
            if (rvalid && j < N) {
                P1[j] = p.P[(((size_t)b_ * 2 + 0) * N + n_) * N + j];
                P2[j] = p.P[(((size_t)b_ * 2 + 1) * N + n_) * N + j];
            }
        }
        unsigned g = 0, xq = 0;
        // x chunk xq (= t*nxc + i) -> ring slot xq % 3: one 16-byte cp.async per thread; the slot's mbarrier
        // completes when every producer thread's copy has landed (cp.async.mbarrier.arrive.noinc)
        auto issue_x = [&](unsigned q_) {
            if (q_ >= total_x) return;
            const int t = q_ / nxc, i = q_ - t * nxc;
            const int r = tid >> 1, q4 = (tid & 1) * 4;
            const int s = r / FT_RP, n = r - s * FT_RP, b = b0 + s;
            const int c = i * FT_CC + q4;
            if (n < N && b < p.B && c < fin) {
                float* dst = reinterpret_cast<float*>(smem + FT_OFF_X + (q_ % FT_XRING) * FT_XSLOT) + r * FT_XLD + q4;
                cp_async16(dst, p.x + (size_t)t * p.xs_t + (size_t)b * p.xs_b + n * fin + c);
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(&bar_xfull[q_ % FT_XRING])) : "memory");
        };
        issue_x(0);
        issue_x(1);
        // A tile of one chunk from 8 source columns; zsrc = [rows][zld] source, c0 = first column inside it
        auto produce = [&](const float* zsrc, int zld, int c0, int cvalid, bool is_x) {
            const int sa = g & 1;
            const bool rec = (p.dbg & 4) && blockIdx.x == 0 && tid == 0 && g < 128;
            if (rec) p.dbgbuf[g * 8 + 4] = clock64();
            // stage free = the MMAs that read it are done and so is its copy to the operand image; for an x chunk
            // also wait for the streamed columns (requested two chunks ago).  One combined poll: see mbar_wait2.
            const uint32_t fpar = ((g >> 1) - 1) & 1;
            uint64_t* bst = dump ? &bar_stored[sa] : &bar_done[sa];
            if (is_x) {
                if (g >= 2) mbar_wait3(&bar_done[sa], fpar, bst, fpar, &bar_xfull[xq % FT_XRING], (xq / FT_XRING) & 1);
                else mbar_wait(&bar_xfull[xq % FT_XRING], (xq / FT_XRING) & 1);
                // every producer is past chunk g-2, so the slot x chunk xq-2 lived in can be refilled
                issue_x(xq + 2);
            } else if (g >= 2) {
                mbar_wait2(&bar_done[sa], fpar, bst, fpar);
            }
            if (rec) p.dbgbuf[g * 8 + 5] = clock64();
            float4* a_hi = reinterpret_cast<float4*>(smem + FT_OFF_A + sa * FT_A_STAGE);
            float4* a_lo = reinterpret_cast<float4*>(smem + FT_OFF_A + sa * FT_A_STAGE + FT_A_BYTES);
            const int c = c0 + 4 * half;
            float v0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
            if (4 * half < cvalid && n_ < NP) {                           // (all lanes of a warp read the same z rows: broadcast)
                const float4 own = *reinterpret_cast<const float4*>(zsrc + row * zld + c);
                v0[0] = own.x; v0[1] = own.y; v0[2] = own.z; v0[3] = own.w;
                const float* zq = zsrc + (s_ * FT_RP) * zld + c;
#pragma unroll
                for (int jb = 0; jb < NP; jb += 10) {                     // rows j >= N are zero, so are P1/P2 there
                    float4 z[10];
#pragma unroll
                    for (int j = 0; j < 10; ++j) z[j] = *reinterpret_cast<const float4*>(zq + (jb + j) * zld);
#pragma unroll
                    for (int j = 0; j < 10; ++j) {
                        a1[0] = fmaf(P1[jb + j], z[j].x, a1[0]); a1[1] = fmaf(P1[jb + j], z[j].y, a1[1]);
                        a1[2] = fmaf(P1[jb + j], z[j].z, a1[2]); a1[3] = fmaf(P1[jb + j], z[j].w, a1[3]);
                        a2[0] = fmaf(P2[jb + j], z[j].x, a2[0]); a2[1] = fmaf(P2[jb + j], z[j].y, a2[1]);
                        a2[2] = fmaf(P2[jb + j], z[j].z, a2[2]); a2[3] = fmaf(P2[jb + j], z[j].w, a2[3]);
                    }
                }
            }
            // kk order inside the quad: (c, m0) (c, m1) (c, m2) (c+1, m0) ...
            float4 f0 = make_float4(v0[0], a1[0], a2[0], v0[1]);
            float4 f1 = make_float4(a1[1], a2[1], v0[2], a1[2]);
            float4 f2 = make_float4(a2[2], v0[3], a1[3], a2[3]);
            float4 h, l;
            const int kg0 = half * 3;
            const int ai0 = ft_a_idx(kg0, row), ai1 = ft_a_idx(kg0 + 1, row), ai2 = ft_a_idx(kg0 + 2, row);
            split4(f0, h, l); a_hi[ai0] = h; a_lo[ai0] = l;
            split4(f1, h, l); a_hi[ai1] = h; a_lo[ai1] = l;
            split4(f2, h, l); a_hi[ai2] = h; a_lo[ai2] = l;
            if (rec) p.dbgbuf[g * 8 + 6] = clock64();
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_afull[sa]);
            if (rec) p.dbgbuf[g * 8 + 7] = clock64();
            ++g;
        };
        auto wait_all_mma = [&]() {                                       // MMAs of the last produced chunk (hence all)
            const unsigned gl = g - 1;
            mbar_wait(&bar_done[gl & 1], (gl >> 1) & 1);
            if (dump) {                                                   // the epilogues stage through the A stages
                mbar_wait(&bar_stored[gl & 1], (gl >> 1) & 1);
                if (gl >= 1) mbar_wait(&bar_stored[(gl - 1) & 1], ((gl - 1) >> 1) & 1);
            }
            tc_fence_after();
        };
        const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
        const int hf = warp >> 2;                                        // which column half this warp reads
        // ---- warp-private staging tile (lives in the A stages, which are idle during the epilogues) -------------
        // [32 rows][36 floats]; lane l owns row l when it exchanges with registers, and rows (l>>3)+4i, float4 l&7
        // when it exchanges with global memory
        float* stg = reinterpret_cast<float*>(smem + FT_OFF_A) + warp * (32 * 36);
        const int rq = lane >> 3, f4 = lane & 7;
        size_t grow[8];                                                   // (b*N + n) of the rows this lane moves, or ~0
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = 32 * (warp & 3) + rq + 4 * i;
            const int s = r / FT_RP, n = r - s * FT_RP, b = b0 + s;
            grow[i] = (n < N && b < p.B) ? ((size_t)b * N + n) : ~(size_t)0;
        }
        auto stage_put = [&](const float (&v)[32]) {
            __syncwarp();
            float4* d = reinterpret_cast<float4*>(stg + lane * 36);
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            __syncwarp();
        };
        auto stage_get = [&](float (&v)[32]) {
            __syncwarp();
            const float4* d = reinterpret_cast<const float4*>(stg + lane * 36);
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float4 q = d[j]; v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w; }
            __syncwarp();
        };
        auto stage_to_global = [&](float* base, int ld, int col0) {      // rows x 32 columns -> base[(b*N+n)*ld + col0 ..]
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (grow[i] != ~(size_t)0)
                    *reinterpret_cast<float4*>(base + grow[i] * ld + col0 + 4 * f4) =
                        *reinterpret_cast<const float4*>(stg + (rq + 4 * i) * 36 + 4 * f4);
        };
        auto global_to_stage = [&](const float* base, int ld, int col0) {
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                if (grow[i] != ~(size_t)0) q = __ldcg(reinterpret_cast<const float4*>(base + grow[i] * ld + col0 + 4 * f4));
                *reinterpret_cast<float4*>(stg + (rq + 4 * i) * 36 + 4 * f4) = q;
            }
        };
        {   // TMEM stash <- h0: thread = (row, column half) keeps its 32 columns of h_{t-1} on chip for the GRU update
            float hp[32];
            global_to_stage(p.h0, FT_H, hf * 32);
            stage_get(hp);
            tmem_st32(taddr + lane_base + FT_STASH + hf * 32, hp);
            producer_barrier();                       // the staging tiles alias the A stages the first chunk is built in
        }
        for (int t = 0; t < p.T; ++t) {
            float* hout = p.hseq + (size_t)t * p.B * NH;
            float* ruc = p.ruc + (size_t)t * p.B * NH * 3;
            // ---- X phase -------------------------------------------------------------------------------------
            for (int i = 0; i < nxc; ++i, ++xq) {
                produce(reinterpret_cast<const float*>(smem + FT_OFF_X + (xq % FT_XRING) * FT_XSLOT), FT_XLD, 0,
                        min(FT_CC, fin - i * FT_CC), true);
            }
            // ---- gate: recurrent part ---------------------------------------------------------------------------
            for (int i = 0; i < nhc; ++i) produce(ZH, FT_ZLD, i * FT_CC, FT_CC, false);
            const bool erec = (p.dbg & 4) && blockIdx.x == 0 && tid == 0 && t < 8;
            long long* es = p.dbgbuf + 1536 + t * 16;
            if (erec) es[0] = clock64();
            wait_all_mma();
            if (erec) es[1] = clock64();
            // epilogue 1: only the reset gate is on the critical path (the candidate needs r*h): thread = (row,
            // column half) takes 32 columns of r.  The update gate's pre-activation stays in TMEM (columns 64..127
            // are not touched by the candidate MMAs) and is turned into u by epilogue 2.
            // Global traffic goes through a warp-private staging tile so that every warp instruction moves
            // whole 128-byte lines (4 rows x 8 float4) instead of 32 scattered 16-byte pieces.
            {
                float v[32];
                tmem_ld32(taddr + lane_base + hf * 32, v);
                if (erec) es[2] = clock64();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fast_sigmoid(v[j] + sbias[hf * 32 + j]);
                if (rvalid) {                                             // ZH <- r * h
                    float4* z4 = reinterpret_cast<float4*>(ZH + row * FT_ZLD + hf * 32);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 z = z4[j];
                        z.x *= v[4 * j]; z.y *= v[4 * j + 1]; z.z *= v[4 * j + 2]; z.w *= v[4 * j + 3];
                        z4[j] = z;
                    }
                }
                if (erec) es[3] = clock64();
                stage_put(v);
                if (p.ruc) stage_to_global(ruc, 3 * FT_H, hf * 32);     // ruc = NULL: inference, nothing saved for BPTT
                if (erec) es[4] = clock64();
            }
            tc_fence_before();
            producer_barrier();                                           // r*h of every row is visible
            if (erec) es[5] = clock64();
            // ---- candidate: recurrent part ------------------------------------------------------------------------
            for (int i = 0; i < nhc; ++i) produce(ZH, FT_ZLD, i * FT_CC, FT_CC, false);
            if (erec) es[6] = clock64();
            wait_all_mma();
            if (erec) es[7] = clock64();
            {   // epilogue 2: 64 columns, each warp half takes 32.  Everything it needs is on chip: the candidate
                // and update-gate pre-activations in the accumulator, h_{t-1} in the TMEM stash (columns 192..255).
                float v[32], u[32], hp[32];
                tmem_ld32(taddr + lane_base + 128 + hf * 32, v);
                tmem_ld32(taddr + lane_base + FT_H + hf * 32, u);
                tmem_ld32(taddr + lane_base + FT_STASH + hf * 32, hp);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_epi);                     // the accumulator is free for the next step's MMAs
                if (erec) es[8] = clock64();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float pre = v[j] + sbias[2 * FT_H + hf * 32 + j];
                    const float cv = (p.act == 0) ? fast_tanh(pre) : fmaxf(pre, 0.f);
                    const float uv = fast_sigmoid(u[j] + sbias[FT_H + hf * 32 + j]);
                    v[j] = cv;
                    u[j] = uv;
                    hp[j] = uv * hp[j] + (1.f - uv) * cv;                  // h_new
                }
                if (rvalid) {                                             // critical path first: the next step reads ZH
                    float4* z4 = reinterpret_cast<float4*>(ZH + row * FT_ZLD + hf * 32);
#pragma unroll
                    for (int j = 0; j < 8; ++j) z4[j] = make_float4(hp[4 * j], hp[4 * j + 1], hp[4 * j + 2], hp[4 * j + 3]);
                }
                if (erec) es[9] = clock64();
                tmem_st32(taddr + lane_base + FT_STASH + hf * 32, hp);
                if (erec) es[10] = clock64();
                if (p.ruc) {
                    stage_put(v);
                    stage_to_global(ruc, 3 * FT_H, 2 * FT_H + hf * 32);
                    stage_put(u);
                    stage_to_global(ruc, 3 * FT_H, FT_H + hf * 32);
                }
                stage_put(hp);
                stage_to_global(hout, FT_H, hf * 32);
                if (erec) es[11] = clock64();
            }
            producer_barrier();                                           // h_t of every row is visible
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(taddr);
}

size_t seq_fwd_tc_wimg_bytes(int fin) { return ft_wimg_bytes(fin); }
// operand image geometry shared with dw_mm.cu: K groups per step, CTAs, bytes
int seq_fwd_tc_kgt(int fin) { return (ft_nxc(fin) + 2 * (FT_H / FT_CC)) * FT_KG; }
int seq_tc_nslab(int B, int T) { return ((B + FT_SB - 1) / FT_SB) * T; }
// floats per image row: the kk of a step rounded up to whole 32-float groups (128-byte TMA rows)
int seq_fwd_tc_kkp(int fin) { return (seq_fwd_tc_kgt(fin) * 4 + 31) / 32 * 32; }
size_t seq_fwd_tc_gsave_bytes(int B, int T, int fin) {
    return (size_t)seq_tc_nslab(B, T) * 2 * TC_IMG_ROWS * seq_fwd_tc_kkp(fin) * 4;
}
bool seq_fwd_tc_supported(int N, int fin, int H, int M, int smem_limit) {
    return H == FT_H && M == FT_M && N <= NP && fin % 4 == 0 && FT_SMEM + 2304 <= smem_limit;
}

cudaError_t launch_seq_fwd_tc(int B, int T, int N, int fin, int act, const float* x, long long xs_t,
                              long long xs_b, const float* h0, const float* P, const float* Wg, const float* bg,
                              const float* Wc, const float* bc, float* wimg, float* hseq, float* ruc,
                              void* gsave, cudaStream_t st) {
    const int nblk = ft_nxc(fin) + 2 * (FT_H / FT_CC);
    pack_w_fwd_kernel<<<nblk, 256, 0, st>>>(Wg, Wc, fin, wimg);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    FwdTcParams p;
    { const char* e = getenv("DCGRU_DBG"); p.dbg = e ? atoi(e) : 0; }
    p.B = B; p.T = T; p.N = N; p.fin = fin; p.act = act; p.x = x; p.xs_t = xs_t; p.xs_b = xs_b; p.h0 = h0; p.P = P;
    p.bg = bg; p.bc = bc; p.wimg = wimg; p.hseq = hseq; p.ruc = ruc;
    p.gsave = reinterpret_cast<uint8_t*>(gsave);
    p.dbgbuf = reinterpret_cast<long long*>(reinterpret_cast<uint8_t*>(wimg) + ((ft_wimg_bytes(fin) + 255) / 256) * 256);
    e = cudaFuncSetAttribute(seq_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM);
    if (e != cudaSuccess) return e;
    CUtensorMap tm;
    memset(&tm, 0, sizeof tm);
    if (gsave) {
        // 4-D view of the row-major image: (8 kk | row in group, stride = one row | K-group pair, 32 B | row group)
        const unsigned long long kkp = seq_fwd_tc_kkp(fin), rowb = kkp * 4;
        const unsigned long long dims[4] = {8, 8, kkp / 8, (unsigned long long)seq_tc_nslab(B, T) * 2 * FT_SB * TC_RG};
        const unsigned long long str[4] = {4, rowb, 32, 8 * rowb};
        const unsigned box[4] = {8, 8, FT_KG / 2, TC_RG};
        e = make_tmap_f32(&tm, gsave, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_32B);
        if (e != cudaSuccess) return e;
    }
    seq_fwd_tc_kernel<<<(B + FT_SB - 1) / FT_SB, FT_THREADS, FT_SMEM, st>>>(p, tm);
    return cudaGetLastError();
}

}  // namespace dcgru
