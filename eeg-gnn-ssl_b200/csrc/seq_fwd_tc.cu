// Persistent encoder-layer forward on the tensor cores (tcgen05, 3xTF32): all T steps of one layer for a
// group of 6 samples per CTA, no relaunch (model/model.py:93-96 x model/cell.py:182-210).
//
// Per step the three projections of the cell run as UMMA GEMMs with M = 128 rows (6 samples x 20 padded
// nodes), accumulating in one 128 x 192 fp32 TMEM tile:
//     X : D[:, 0:192] (=)  diffuse(x_t)      @ [Wg_x | Wc_x]     (K = Fin*M)
//     Hg: D[:, 0:128] (+)= diffuse(h_{t-1})  @  Wg_h             (K = H*M)      -> r, u = sigmoid(. + bg)
//     Hc: D[:,128:192](+)= diffuse(r*h_{t-1})@  Wc_h             (K = H*M)      -> c = act(. + bc), GRU update
// A operand: each thread owns one (sample, node) row and keeps that row of the diffusion polynomials P_m
// in registers for the whole sequence; for every 8-column chunk of [x | h] it forms the M diffusion terms
// of 4 columns (20 broadcast float4 reads + 40 FMAs per column quad), splits them hi/lo and writes them
// in kk = c*M + m order into a K-group-major UMMA tile (2 stages).
// B operand: the weights, pre-split and pre-tiled once per launch by pack_w_fwd_kernel, streamed chunk by
// chunk from L2 with cp.async.bulk (TMA) into a 3-slot ring, each load issued one chunk ahead at the top
// of the iteration so its latency hides behind two A-tile productions; mbarriers track "weights landed"
// and "MMAs done".  x_t is streamed the same way, 8 columns at a time, with cp.async into a 3-slot ring
// (two chunks ahead); only the hidden state stays resident in shared memory.
// The epilogues read the accumulator with tcgen05.ld (thread = row) and fuse bias, sigmoid/tanh, r*h and
// the GRU update.
#include "common.cuh"
#include "dw.cuh"
#include "tc_common.cuh"

namespace dcgru {
using namespace tc;

constexpr int FT_SB = 6;                 // samples per CTA
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
constexpr int FT_OFF_ZH = FT_OFF_X + 3 * FT_XSLOT;         // [128][64] hidden state (or r*h)
constexpr int FT_SMEM = FT_OFF_ZH + FT_ROWS * FT_ZLD * 4;

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
    int B, T, N, fin, act;
    const float* x; long long xs_t, xs_b;
    const float* h0;
    const float* P;
    const float* bg; const float* bc;
    const float* wimg;
    float* hseq;
    float* ruc;
};

__global__ void __launch_bounds__(NT, 1) seq_fwd_tc_kernel(const FwdTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar_full[3], bar_done[2];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float sbias[3 * FT_H];                      // bg (r | u) | bc
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = p.N, fin = p.fin;
    const int b0 = blockIdx.x * FT_SB;
    if (tid < 3 * FT_H) sbias[tid] = (tid < 2 * FT_H) ? p.bg[tid] : p.bc[tid - 2 * FT_H];
    float* ZH = reinterpret_cast<float*>(smem + FT_OFF_ZH);              // [128][64]
    const int row = tid & 127, half = tid >> 7;
    const int s_ = row / NP, n_ = row - s_ * NP;
    const int b_ = b0 + s_;
    const bool rvalid = (s_ < FT_SB) && (n_ < N) && (b_ < p.B);

    if (warp == 0) tmem_alloc<256>(&tmem_slot);
    if (tid == 0) {
        mbar_init(&bar_full[0], 1); mbar_init(&bar_full[1], 1); mbar_init(&bar_full[2], 1);
        mbar_init(&bar_done[0], 1); mbar_init(&bar_done[1], 1);
        mbar_fence_init();
    }
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
    // hidden state <- h0, x ring <- 0 (rows of pad nodes / missing samples stay zero for ever)
    for (int idx = tid; idx < FT_ROWS * (FT_H / 4); idx += NT) {
        const int r = idx / (FT_H / 4), c4 = (idx - r * (FT_H / 4)) * 4;
        const int s = r / NP, n = r - s * NP, b = b0 + s;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s < FT_SB && n < N && b < p.B) v = *reinterpret_cast<const float4*>(p.h0 + ((size_t)b * N + n) * FT_H + c4);
        *reinterpret_cast<float4*>(ZH + r * FT_ZLD + c4) = v;
    }
    for (int idx = tid; idx < 3 * FT_XSLOT / 16; idx += NT)
        reinterpret_cast<float4*>(smem + FT_OFF_X)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = tmem_slot;

    const int nxc = ft_nxc(fin), nhc = FT_H / FT_CC;
    const int per_step = nxc + 2 * nhc;
    const unsigned total_chunks = (unsigned)p.T * per_step;
    const uint8_t* wimg = reinterpret_cast<const uint8_t*>(p.wimg);
    const size_t NH = (size_t)N * FT_H;

    // weight block of chunk index q within a step
    auto wblock = [&](int q, const uint8_t*& src, uint32_t& bytes) {
        if (q < nxc) { src = wimg + (size_t)q * FT_BX_BYTES; bytes = FT_BX_BYTES; }
        else if (q < nxc + nhc) { src = wimg + (size_t)nxc * FT_BX_BYTES + (size_t)(q - nxc) * FT_BG_BYTES; bytes = FT_BG_BYTES; }
        else { src = wimg + (size_t)nxc * FT_BX_BYTES + (size_t)nhc * FT_BG_BYTES + (size_t)(q - nxc - nhc) * FT_BC_BYTES;
               bytes = FT_BC_BYTES; }
    };
    // x chunk xq (= t*nxc + i) -> ring slot xq % 3 ; one 16-byte piece per thread
    const unsigned total_x = (unsigned)p.T * nxc;
    auto issue_x = [&](unsigned xq) {
        if (xq < total_x) {
            const int t = xq / nxc, i = xq - t * nxc;
            const int r = tid >> 1, q4 = (tid & 1) * 4;
            const int s = r / NP, n = r - s * NP, b = b0 + s;
            const int c = i * FT_CC + q4;
            if (s < FT_SB && n < N && b < p.B && c < fin) {
                float* dst = reinterpret_cast<float*>(smem + FT_OFF_X + (xq % 3) * FT_XSLOT) + r * FT_XLD + q4;
                cp_async16(dst, p.x + (size_t)t * p.xs_t + (size_t)b * p.xs_b + n * fin + c);
            }
        }
        cp_async_commit();
    };

    unsigned g = 0;                                                      // global chunk counter
    unsigned xq = 0;                                                     // x chunk counter
    // prologue: weights of chunk 0, x chunks 0 and 1
    if (tid == 0) { const uint8_t* src; uint32_t bytes; wblock(0, src, bytes); bulk_load(smem + FT_OFF_B, src, bytes, &bar_full[0]); }
    issue_x(0);
    issue_x(1);
    cp_async_wait<1>();
    __syncthreads();

    // one chunk: prefetch the next weight block, build the A tile from 8 source columns, issue the MMAs
    // zsrc: base of the [rows][zld] source (x ring slot or hidden state), c0 = first column inside it
    auto chunk = [&](const float* zsrc, int zld, int c0, int cvalid, int ncols, uint32_t dcol, bool fresh, bool is_x) {
        const int sa = g & 1, sb = g % 3;
        if (g >= 2) mbar_wait(&bar_done[sa], ((g >> 1) - 1) & 1);          // frees A stage sa and B slot (g+1)%3
        if (tid == 0 && g + 1 < total_chunks) {
            const uint8_t* src; uint32_t bytes;
            wblock((int)((g + 1) % per_step), src, bytes);
            bulk_load(smem + FT_OFF_B + ((g + 1) % 3) * FT_BX_BYTES, src, bytes, &bar_full[(g + 1) % 3]);
        }
        if (is_x) issue_x(xq + 2);
        float4* a_hi = reinterpret_cast<float4*>(smem + FT_OFF_A + sa * FT_A_STAGE);
        float4* a_lo = reinterpret_cast<float4*>(smem + FT_OFF_A + sa * FT_A_STAGE + FT_A_BYTES);
        // ---- A tile: this thread's row, 4 source columns --------------------------------------------------
        {
            const int c = c0 + 4 * half;
            float v0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
            if (4 * half < cvalid && row < FT_SB * NP) {
                const float4 own = *reinterpret_cast<const float4*>(zsrc + row * zld + c);
                v0[0] = own.x; v0[1] = own.y; v0[2] = own.z; v0[3] = own.w;
                const float* zq = zsrc + (s_ * NP) * zld + c;
#pragma unroll
                for (int j = 0; j < NP; ++j) {                            // rows j >= N are zero, so are P1/P2 there
                    const float4 z = *reinterpret_cast<const float4*>(zq + j * zld);
                    a1[0] = fmaf(P1[j], z.x, a1[0]); a1[1] = fmaf(P1[j], z.y, a1[1]);
                    a1[2] = fmaf(P1[j], z.z, a1[2]); a1[3] = fmaf(P1[j], z.w, a1[3]);
                    a2[0] = fmaf(P2[j], z.x, a2[0]); a2[1] = fmaf(P2[j], z.y, a2[1]);
                    a2[2] = fmaf(P2[j], z.z, a2[2]); a2[3] = fmaf(P2[j], z.w, a2[3]);
                }
            }
            // kk order inside the quad: (c, m0) (c, m1) (c, m2) (c+1, m0) ...
            float4 f0 = make_float4(v0[0], a1[0], a2[0], v0[1]);
            float4 f1 = make_float4(a1[1], a2[1], v0[2], a1[2]);
            float4 f2 = make_float4(a2[2], v0[3], a1[3], a2[3]);
            float4 h, l;
            const int kg0 = half * 3;
            split4(f0, h, l); a_hi[(kg0 + 0) * FT_ROWS + row] = h; a_lo[(kg0 + 0) * FT_ROWS + row] = l;
            split4(f1, h, l); a_hi[(kg0 + 1) * FT_ROWS + row] = h; a_lo[(kg0 + 1) * FT_ROWS + row] = l;
            split4(f2, h, l); a_hi[(kg0 + 2) * FT_ROWS + row] = h; a_lo[(kg0 + 2) * FT_ROWS + row] = l;
        }
        if (is_x) { cp_async_wait<1>(); ++xq; }                           // the next x chunk has landed (this thread's part)
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            mbar_wait(&bar_full[sb], (g / 3) & 1);
            tc_fence_after();
            const uint32_t b_hi = smem_u32(smem + FT_OFF_B + sb * FT_BX_BYTES), b_lo = b_hi + FT_KG * ncols * 16;
            issue_3xtf32(taddr + dcol, smem_u32(a_hi), smem_u32(a_lo), FT_ROWS, b_hi, b_lo, ncols, FT_KK / 8,
                         make_idesc_tf32(128, ncols), !fresh);
            umma_commit(&bar_done[sa]);
        }
        ++g;
    };
    auto wait_all_mma = [&]() {                                           // MMAs of the last issued chunk (hence all)
        const unsigned gl = g - 1;
        mbar_wait(&bar_done[gl & 1], (gl >> 1) & 1);
        tc_fence_after();
    };

    const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
    const int hf = warp >> 2;                                            // which column half this warp reads
    const size_t ro = ((size_t)b_ * N + n_);
    for (int t = 0; t < p.T; ++t) {
        const float* hprev = (t == 0) ? p.h0 : p.hseq + (size_t)(t - 1) * p.B * NH;
        float* hout = p.hseq + (size_t)t * p.B * NH;
        float* ruc = p.ruc + (size_t)t * p.B * NH * 3;
        // ---- X phase -----------------------------------------------------------------------------------------
        for (int i = 0; i < nxc; ++i) {
            const float* xs = reinterpret_cast<const float*>(smem + FT_OFF_X + (xq % 3) * FT_XSLOT);
            chunk(xs, FT_XLD, 0, min(FT_CC, fin - i * FT_CC), 192, 0, i == 0, true);
        }
        // ---- gate: recurrent part -------------------------------------------------------------------------------
        for (int i = 0; i < nhc; ++i) chunk(ZH, FT_ZLD, i * FT_CC, FT_CC, 128, 0, false, false);
        wait_all_mma();
        // epilogue 1: warps 0-3 -> r (cols 0..63), warps 4-7 -> u (cols 64..127); thread = row
        for (int cb = 0; cb < FT_H; cb += 32) {
            float v[32];
            tmem_ld32(taddr + lane_base + hf * FT_H + cb, v);
            if (rvalid) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = sigmoidf_(v[j] + sbias[hf * FT_H + cb + j]);
                float4* q = reinterpret_cast<float4*>(ruc + ro * 3 * FT_H + hf * FT_H + cb);
#pragma unroll
                for (int j = 0; j < 8; ++j) q[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                if (hf == 0) {                                            // ZH <- r * h
                    float4* z4 = reinterpret_cast<float4*>(ZH + row * FT_ZLD + cb);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 z = z4[j];
                        z.x *= v[4 * j]; z.y *= v[4 * j + 1]; z.z *= v[4 * j + 2]; z.w *= v[4 * j + 3];
                        z4[j] = z;
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();
        // ---- candidate: recurrent part ----------------------------------------------------------------------------
        for (int i = 0; i < nhc; ++i) chunk(ZH, FT_ZLD, i * FT_CC, FT_CC, 64, 128, false, false);
        wait_all_mma();
        {   // epilogue 2: 64 columns, each warp half takes 32; all global loads first, then math, then stores
            float v[32], u[32], hp[32];
            if (rvalid) {
                const float4* uq = reinterpret_cast<const float4*>(ruc + ro * 3 * FT_H + FT_H + hf * 32);
                const float4* hq = reinterpret_cast<const float4*>(hprev + ro * FT_H + hf * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 a = __ldcg(uq + j), b = __ldcg(hq + j);
                    u[4 * j] = a.x; u[4 * j + 1] = a.y; u[4 * j + 2] = a.z; u[4 * j + 3] = a.w;
                    hp[4 * j] = b.x; hp[4 * j + 1] = b.y; hp[4 * j + 2] = b.z; hp[4 * j + 3] = b.w;
                }
            }
            tmem_ld32(taddr + lane_base + 128 + hf * 32, v);
            if (rvalid) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float pre = v[j] + sbias[2 * FT_H + hf * 32 + j];
                    const float cv = (p.act == 0) ? tanhf(pre) : fmaxf(pre, 0.f);
                    v[j] = cv;
                    u[j] = u[j] * hp[j] + (1.f - u[j]) * cv;               // h_new
                }
                float4* cq = reinterpret_cast<float4*>(ruc + ro * 3 * FT_H + 2 * FT_H + hf * 32);
                float4* ho = reinterpret_cast<float4*>(hout + ro * FT_H + hf * 32);
                float4* z4 = reinterpret_cast<float4*>(ZH + row * FT_ZLD + hf * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    cq[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    const float4 hn = make_float4(u[4 * j], u[4 * j + 1], u[4 * j + 2], u[4 * j + 3]);
                    ho[j] = hn;
                    z4[j] = hn;
                }
            }
        }
        tc_fence_before();
        __syncthreads();
    }
    cp_async_wait<0>();
    if (warp == 0) tmem_dealloc<256>(taddr);
}

size_t seq_fwd_tc_wimg_bytes(int fin) { return ft_wimg_bytes(fin); }
bool seq_fwd_tc_supported(int N, int fin, int H, int M, int smem_limit) {
    return H == FT_H && M == FT_M && N <= NP && fin % 4 == 0 && FT_SMEM + 1088 <= smem_limit;
}

cudaError_t launch_seq_fwd_tc(int B, int T, int N, int fin, int act, const float* x, long long xs_t,
                              long long xs_b, const float* h0, const float* P, const float* Wg, const float* bg,
                              const float* Wc, const float* bc, float* wimg, float* hseq, float* ruc,
                              cudaStream_t st) {
    const int nblk = ft_nxc(fin) + 2 * (FT_H / FT_CC);
    pack_w_fwd_kernel<<<nblk, 256, 0, st>>>(Wg, Wc, fin, wimg);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    FwdTcParams p;
    p.B = B; p.T = T; p.N = N; p.fin = fin; p.act = act; p.x = x; p.xs_t = xs_t; p.xs_b = xs_b; p.h0 = h0; p.P = P;
    p.bg = bg; p.bc = bc; p.wimg = wimg; p.hseq = hseq; p.ruc = ruc;
    e = cudaFuncSetAttribute(seq_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM);
    if (e != cudaSuccess) return e;
    seq_fwd_tc_kernel<<<(B + FT_SB - 1) / FT_SB, NT, FT_SMEM, st>>>(p);
    return cudaGetLastError();
}

}  // namespace dcgru
