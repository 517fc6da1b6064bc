// Bulk "diffuse, then project" on the tensor cores (tcgen05, 2xFP16) for every (tile, step) of a layer at once:
//     out[t][b][n][:] = ( sum_m  (P_m z_t)[n][:] @ W_m ) * scale + bias          z_t = src[t][b] (N x Cin, fp32)
// The two non-recurrent contractions of a DCGRU layer are instances of it:
//   * x-part pre-projection (model/cell.py:73-116 restricted to the input columns, legal for the whole
//     sequence because a layer's input exists up front, model/model.py:98):  z = x_t, W = [Wg_x | Wc_x] (N = 3H),
//     + biases.  The recurrent kernel (rnn_fwd.cu) starts every step from this pre-activation.  The diffused operand
//     tiles are also dumped to the operand image (columns [0, KXP)) for the weight-gradient GEMM.
//   * input gradient dX (BPTT of the same columns):  z = [dA_r | dA_u | dA_c]_t, P_m -> P_m^T, W = W_x^T (N = Fin).
// Work item = (tile of 4 samples, t); a persistent CTA owns a contiguous range of items (so the tile's polynomials
// are reloaded only when the tile changes).  Per item the K dimension is cut into chunks of 64 (f16_common.cuh); the
// K order is term-major, kk = m * Cin + c, so a chunk is a column range of at most two terms.
//   warps 0-7  workers: one warp per (sample, chunk) task -- fp32 diffusion, hi/lo split, write the A chunk (3-slot ring)
//   warp 8     MMA issuer (one thread): per chunk 3 x 4 kind::f16 MMAs into one of two TMEM accumulators
//   warp 9     loader (one thread): source tile (one bulk copy per sample) and the weight pieces (hi / lo plane of
//                       a chunk, pre-swizzled by pack_w16_kernel) from L2 into a ring, cp.async.bulk + mbarrier
//   warp 10    operand-image dump: one 2-D tensor-map TMA store per (plane, sample) and chunk
//   warps 11-14 epilogue of the PREVIOUS item while the workers diffuse the next one: TMEM -> scale / bias -> warp-private
//                       staging tile -> whole 128-byte row pieces in global memory (thread = row stores cost 32 sectors per
//                       instruction and shared the LSU with the workers' diffusion loads)
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "dw.cuh"
#include "f16_common.cuh"
#include "tmap.cuh"

namespace dcgru {
using namespace f16;

constexpr int BK_NS = 3;                  // A chunk slots
constexpr int BK_MAXQ = 16;               // chunks per item (M * Cin <= 1024)
constexpr int BK_THREADS = 480;
constexpr int BK_STG_LD = 36;                // staging tile row stride (floats)
constexpr int BK_STG = 32 * BK_STG_LD * 4;     // one warp-private staging tile: 32 rows x 32 columns
constexpr int BK_NWORK = 256;

struct BulkParams {
    int B, T, N, Cin, M, Nout, transposeP;
    int nout_valid;                       // output columns that exist (<= Nout; the rest of the MMA tile is padding)
    int img_T, img_t0, src_T, src_t0;     // slab of (tile, t) inside the dumped image / the source image: tile * X_T + X_t0 + t
    int ntile, item0_stride;              // items = ntile * T, split evenly over the grid
    int NQ, NW, piece_bytes;
    const float* src; long long ss_t, ss_b;
    const __half* src16;                  // alternative source: fp16 operand image [tile*T+t][hi|lo][96][Cin] (values scaled by *scale_ptr)
    const float* P;
    const uint8_t* wimg;
    const float* bias;
    float* out; long long os_t, os_b; int out_ld;
    float out_scale;
    const float* scale_ptr;               // optional device scalar: out *= 1 / *scale_ptr (gradient scaling, rnn_bwd.cu)
    const float* in_scale_ptr;            // optional device scalar: the fp32 source is multiplied by it before the hi/lo split
    int dump, img_col0;
    int mma_diff;                         // chunks inside one term: diffusion on mma.sync (Cin % 64 == 0, DCGRU_MMA_DIFF_BULK=1; default: FMA loop)
    int off_w, off_x, off_pt, off_id, off_stg;     // shared-memory offsets
};

__device__ __forceinline__ void named_bar_workers() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }

// EPIW: the epilogue runs on four dedicated warps (480 threads, 128 registers) instead of on the workers (352 threads)
template <bool EPIW>
__global__ void __launch_bounds__(EPIW ? BK_THREADS : BK_THREADS - 128, 1) bulk_dp_kernel(const BulkParams p, const __grid_constant__ CUtensorMap tm_img) {
    constexpr int NTHREADS = EPIW ? BK_THREADS : BK_THREADS - 128;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar_afull[BK_MAXQ], bar_aempty[BK_NS], bar_stored[BK_NS], bar_wfull[8], bar_wempty[8];
    __shared__ uint64_t bar_xfull, bar_xfree, bar_accfull[2], bar_accfree[2];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float sbias[192];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = p.N, Cin = p.Cin, M = p.M, NQ = p.NQ, NW = p.NW;
    if (tid < 192) sbias[tid] = (p.bias && tid < p.nout_valid) ? p.bias[tid] : 0.f;
    const long nitems = (long)p.ntile * p.T;
    const long it0 = nitems * blockIdx.x / gridDim.x, it1 = nitems * (blockIdx.x + 1) / gridDim.x;
    const int nloc = (int)(it1 - it0);
    uint8_t* Aslots = smem;
    uint8_t* Wring = smem + p.off_w;
    float* XT = reinterpret_cast<float*>(smem + p.off_x);
    float* PTs = reinterpret_cast<float*>(smem + p.off_pt);
    float* PTid = reinterpret_cast<float*>(smem + p.off_id);
    const int srow = N * Cin;                                   // floats per sample of the source tile
    const bool src16 = p.src16 != nullptr;
    const __half* X16 = reinterpret_cast<const __half*>(smem + p.off_x);     // [hi|lo][96][Cin] when the source is an image

    if (warp == 0) tmem_alloc<512>(&tmem_slot);
    if (tid == 0) {
        for (int i = 0; i < BK_MAXQ; ++i) mbar_init(&bar_afull[i], SB);
        for (int i = 0; i < BK_NS; ++i) { mbar_init(&bar_aempty[i], 1); mbar_init(&bar_stored[i], 1); }
        for (int i = 0; i < 8; ++i) { mbar_init(&bar_wfull[i], 1); mbar_init(&bar_wempty[i], 1); }
        mbar_init(&bar_xfull, 1);
        mbar_init(&bar_xfree, BK_NWORK / 32);
        for (int i = 0; i < 2; ++i) { mbar_init(&bar_accfull[i], 1); mbar_init(&bar_accfree[i], EPIW ? 4 : BK_NWORK / 32); }
        mbar_fence_init();
    }
    // chunk slots, source tile and polynomial blocks start as zeros: pad rows / pad nodes are never written again
    for (int i = tid; i < BK_NS * SLOT / 16; i += NTHREADS) reinterpret_cast<uint4*>(Aslots)[i] = make_uint4(0, 0, 0, 0);
    if (!src16) for (int i = tid; i < SB * srow; i += NTHREADS) XT[i] = 0.f;
    for (int i = tid; i < SB * (M - 1) * PT_STRIDE; i += NTHREADS) PTs[i] = 0.f;
    for (int i = tid; i < PT_STRIDE; i += NTHREADS) PTid[i] = (i / NPAD == i % NPAD) ? 1.f : 0.f;
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = tmem_slot;

    if (warp == 8) {
        // =================================== MMA issuer =================================================================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_f16(128, p.Nout);
            const uint32_t a_base = smem_u32(Aslots), w_base = smem_u32(Wring);
            unsigned pc = 0;                                    // weight piece counter (2 per chunk)
            for (int k = 0; k < nloc; ++k) {
                const int acc_i = k & 1;
                if (k >= 2) mbar_wait(&bar_accfree[acc_i], ((k >> 1) - 1) & 1);
                const uint32_t d = taddr + acc_i * 256;
                for (int q = 0; q < NQ; ++q) {
                    const unsigned g = (unsigned)k * NQ + q;
                    const int slot = g % BK_NS;
                    const uint32_t ah = a_base + slot * SLOT, al = ah + PLANE;
                    const int ws0 = pc % NW, ws1 = (pc + 1) % NW;
                    mbar_wait2(&bar_afull[q], k & 1, &bar_wfull[ws0], (pc / NW) & 1);
                    tc_fence_after();
                    const uint32_t bh = w_base + ws0 * p.piece_bytes;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        umma_f16(d, make_desc_k128(al + 32 * ks), make_desc_k128(bh + 32 * ks), idesc, (q | ks) ? 1u : 0u);
                        umma_f16(d, make_desc_k128(ah + 32 * ks), make_desc_k128(bh + 32 * ks), idesc, 1u);
                    }
                    umma_commit(&bar_wempty[ws0]);
                    mbar_wait(&bar_wfull[ws1], ((pc + 1) / NW) & 1);
                    tc_fence_after();
                    const uint32_t bl = w_base + ws1 * p.piece_bytes;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_f16(d, make_desc_k128(ah + 32 * ks), make_desc_k128(bl + 32 * ks), idesc, 1u);
                    umma_commit(&bar_wempty[ws1]);
                    umma_commit(&bar_aempty[slot]);
                    pc += 2;
                }
                umma_commit(&bar_accfull[acc_i]);
            }
        }
        __syncwarp();
    } else if (warp == 9) {
        // =================================== loader ======================================================================
        if (lane == 0) {
            unsigned pc = 0;
            for (int k = 0; k < nloc; ++k) {
                const long it = it0 + k;
                const int tile = (int)(it / p.T), t = (int)(it - (long)tile * p.T);
                if (k >= 1) mbar_wait(&bar_xfree, (k - 1) & 1);
                if (src16) {                                      // one slab of the image: both planes, all 96 rows
                    const uint32_t bytes = (uint32_t)(2 * IMG_ROWS * Cin * 2);
                    mbar_expect_tx(&bar_xfull, bytes);
                    bulk_g2s(XT, p.src16 + (size_t)((long)tile * p.src_T + p.src_t0 + t) * (2 * IMG_ROWS * Cin), bytes, &bar_xfull);
                } else {
                    int nvalid = p.B - tile * SB; if (nvalid > SB) nvalid = SB;
                    mbar_expect_tx(&bar_xfull, (uint32_t)(nvalid * srow * 4));
                    for (int s = 0; s < nvalid; ++s)
                        bulk_g2s(XT + s * srow, p.src + (size_t)t * p.ss_t + (size_t)(tile * SB + s) * p.ss_b, (uint32_t)(srow * 4), &bar_xfull);
                }
                for (int q = 0; q < 2 * NQ; ++q, ++pc) {
                    const int ws = pc % NW;
                    if (pc >= (unsigned)NW) mbar_wait(&bar_wempty[ws], ((pc / NW) - 1) & 1);
                    mbar_expect_tx(&bar_wfull[ws], (uint32_t)p.piece_bytes);
                    bulk_g2s(Wring + ws * p.piece_bytes, p.wimg + (size_t)q * p.piece_bytes, (uint32_t)p.piece_bytes, &bar_wfull[ws]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 10) {
        // =================================== operand-image dump ==========================================================
        if (p.dump) {
            const int plane = lane >> 2, s = lane & 3;
            if (lane == 0) tma_prefetch_desc(&tm_img);
            for (int k = 0; k < nloc; ++k) {
                const long it = it0 + k;
                const int tile = (int)(it / p.T), t = (int)(it - (long)tile * p.T);
                const long slab = (long)tile * p.img_T + p.img_t0 + t;
                for (int q = 0; q < NQ; ++q) {
                    const unsigned g = (unsigned)k * NQ + q;
                    const int slot = g % BK_NS;
                    mbar_wait(&bar_afull[q], k & 1);
                    if (lane < 2 * SB) {
                        tma_store_2d(&tm_img, p.img_col0 + 64 * q, (int)((slab * 2 + plane) * IMG_ROWS + s * (RG * 8)),
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
    } else if (EPIW && warp >= 11) {
        // =================================== epilogue warps: thread = row =================================================
        const int quad = warp & 3, row = 32 * quad + lane;              // TMEM lane quadrant = warp % 4
        float* stg = reinterpret_cast<float*>(smem + p.off_stg + (warp - 11) * BK_STG);
        const int rq = lane >> 3, f4 = lane & 7;
        float oscale = p.out_scale;
        if (p.scale_ptr) oscale *= 1.f / __ldg(p.scale_ptr);
        const uint32_t tb = taddr + ((uint32_t)(32 * quad) << 16);
        for (int k = 0; k < nloc; ++k) {
            const long it = it0 + k;
            const int tile = (int)(it / p.T), t = (int)(it - (long)tile * p.T);
            const int acc_i = k & 1;
            const int b = tile * SB + quad;                             // this warp's sample
            mbar_wait(&bar_accfull[acc_i], (k >> 1) & 1);
            tc_fence_after();
            float* obase = p.out + (size_t)t * p.os_t + (size_t)(b < p.B ? b : 0) * p.os_b;
            for (int cb = 0; cb < p.Nout; cb += 32) {
                float v[32];
                tmem_ld32(tb + acc_i * 256 + cb, v);
                if (cb >= p.nout_valid) continue;
                __syncwarp();
                float4* d = reinterpret_cast<float4*>(stg + lane * BK_STG_LD);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 bq = *reinterpret_cast<const float4*>(sbias + cb + 4 * j);
                    d[j] = make_float4(fmaf(v[4 * j], oscale, bq.x), fmaf(v[4 * j + 1], oscale, bq.y), fmaf(v[4 * j + 2], oscale, bq.z),
                                       fmaf(v[4 * j + 3], oscale, bq.w));
                }
                __syncwarp();
                if (b < p.B && cb + 4 * f4 < p.nout_valid) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int n = rq + 4 * i;
                        if (n < N)
                            *reinterpret_cast<float4*>(obase + (size_t)n * p.out_ld + cb + 4 * f4) =
                                *reinterpret_cast<const float4*>(stg + n * BK_STG_LD + 4 * f4);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_accfree[acc_i]);
        }
    } else {
        // =================================== workers =====================================================================
        const int row = 32 * (warp & 3) + lane, half = warp >> 2;
        const int es = row >> 5, en = row & 31;                 // (!EPIW) epilogue: sample / node of this thread's row
        const int ncol = p.Nout / 2;                            // columns per thread in the epilogue
        float oscale = p.out_scale;
        if (p.scale_ptr) oscale *= 1.f / __ldg(p.scale_ptr);
        const float iscale = p.in_scale_ptr ? __ldg(p.in_scale_ptr) : 1.f;
        // epilogue on the workers (!EPIW): after the tasks of item k, the accumulator of item k-1 (so its MMAs had a whole item's
        // worth of diffusion time to finish): TMEM -> scale / bias -> global, thread = (row, column half)
        auto epilogue = [&](int k) {
            const long it = it0 + k;
            const int tile = (int)(it / p.T), t = (int)(it - (long)tile * p.T);
            const int acc_i = k & 1;
            mbar_wait(&bar_accfull[acc_i], (k >> 1) & 1);
            tc_fence_after();
            const int b = tile * SB + es;
            const bool valid = en < N && b < p.B;
            float* orow = p.out + (size_t)t * p.os_t + (size_t)(valid ? b : 0) * p.os_b + (size_t)(valid ? en : 0) * p.out_ld + half * ncol;
            const uint32_t tb = taddr + ((uint32_t)(32 * (warp & 3)) << 16) + acc_i * 256 + half * ncol;
            for (int cb = 0; cb < ncol; cb += 32) {
                float v[32];
                tmem_ld32(tb + cb, v);
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 o;
                        const int c = half * ncol + cb + j;
                        if (c >= p.nout_valid) continue;
                        const float4 bq = *reinterpret_cast<const float4*>(sbias + c);
                        o.x = fmaf(v[j], oscale, bq.x); o.y = fmaf(v[j + 1], oscale, bq.y);
                        o.z = fmaf(v[j + 2], oscale, bq.z); o.w = fmaf(v[j + 3], oscale, bq.w);
                        *reinterpret_cast<float4*>(orow + cb + j) = o;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_accfree[acc_i]);
        };
        int cur_tile = -1;
        const int kkmax = M * Cin;
        for (int k = 0; k < nloc; ++k) {
            const long it = it0 + k;
            const int tile = (int)(it / p.T);
            if (tile != cur_tile) {                             // (every worker is past the tasks of the previous item)
                named_bar_workers();
                load_pt(PTs, p.P, tile * SB, p.B, N, M - 1, p.transposeP, tid, BK_NWORK);
                named_bar_workers();
                cur_tile = tile;
            }
            mbar_wait(&bar_xfull, k & 1);
            // fewer chunks than slots: nothing else keeps this item's tasks from completing a chunk barrier a second
            // time before the issuer has seen the first completion -> wait for the previous item's MMAs here
            if (NQ < BK_NS && k >= 1) mbar_wait(&bar_accfull[(k - 1) & 1], ((k - 1) >> 1) & 1);
            for (int tt = warp; tt < SB * NQ; tt += BK_NWORK / 32) {
                const int q = tt / SB, s = tt - q * SB;
                const unsigned g = (unsigned)k * NQ + q;
                const int slot = g % BK_NS;
                const unsigned use = g / BK_NS;
                if (use >= 1) {
                    if (p.dump) mbar_wait2(&bar_aempty[slot], (use - 1) & 1, &bar_stored[slot], (use - 1) & 1);
                    else mbar_wait(&bar_aempty[slot], (use - 1) & 1);
                }
                uint8_t* sl = Aslots + slot * SLOT;
                const int kk = 64 * q + 2 * lane;
                const bool kvalid = kk < kkmax && (tile * SB + s) < p.B;
                int m = 0, c = 0;
                if (kvalid) { m = kk / Cin; c = kk - m * Cin; }
                const float* z = XT + s * srow + c;
                const __half* zh = X16 + (s * (RG * 8)) * Cin + c;                  // image rows s*24 + n
                const __half* zl = zh + IMG_ROWS * Cin;
                // a chunk that lies inside one diffusion term (Cin % 64 == 0: layers >= 1, dX) goes through the warp-level tensor
                // path like the recurrent kernels: polynomial fragments in registers, the source read once (f16_common.cuh)
                const bool svalid = (tile * SB + s) < p.B;
                if (p.mma_diff && svalid && 64 * q >= Cin && 64 * q + 64 <= kkmax) {
                    const int mq = (64 * q) / Cin, c0 = 64 * q - mq * Cin;
                    PFrag pf;
                    load_pfrag(PTs + (s * (M - 1) + (mq - 1)) * PT_STRIDE, lane, pf);
                    if (src16) diffuse_mma16_rm(X16 + (s * (RG * 8)) * Cin + c0, (uint32_t)(Cin * 2), (uint32_t)(IMG_ROWS * Cin * 2), pf, sl,
                                                s * RP, lane, iscale);
                    else diffuse_mma(XT + s * srow + c0, Cin, N, pf, sl, s * RP, lane, iscale);
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_afull[q]);
                    continue;
                }
                float acc[NPAD][2];
                if (__all_sync(0xffffffffu, m == 0)) {          // identity term (or padding): plain copy
#pragma unroll
                    for (int n = 0; n < NPAD; ++n) {
                        float2 zz = make_float2(0.f, 0.f);
                        if (kvalid && n < N) zz = src16 ? ld_hilo2(zh + n * Cin, zl + n * Cin) : *reinterpret_cast<const float2*>(z + n * Cin);
                        acc[n][0] = zz.x; acc[n][1] = zz.y;
                    }
                } else {
                    const float* pt = (m == 0) ? PTid : PTs + (s * (M - 1) + (m - 1)) * PT_STRIDE;
                    if (src16) diffuse2f([&](int j) { return ld_hilo2(zh + j * Cin, zl + j * Cin); }, N, pt, acc);
                    else diffuse2(z, Cin, N, pt, acc);
                    if (!kvalid) {
#pragma unroll
                        for (int n = 0; n < NPAD; ++n) { acc[n][0] = 0.f; acc[n][1] = 0.f; }
                    }
                }
                store_cols2(sl, s * RP, lane, N, acc, iscale);
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_afull[q]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_xfree);
            if (!EPIW && k >= 1) epilogue(k - 1);
        }
        if (!EPIW && nloc > 0) epilogue(nloc - 1);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(taddr);
}

// ---- weight images ----------------------------------------------------------------------------------------------------
// piece (chunk q, plane) = [nrows][64 k] fp16 in the K-major SWIZZLE_128B byte order, so a plain bulk copy drops it into
// shared memory ready for the tensor core.  The value of (output row n, kk = 64 q + k) depends on the GEMM:
//   mode 0  x pre-projection   kk = m*fin + c;  n < 2H: Wg[(c*M+m)][n], else Wc[(c*M+m)][n-2H]                 nrows = 3H
//   mode 1  dX                 kk = m*3H + o;   n = input column c:  o < 2H ? Wg[(n*M+m)][o] : Wc[(n*M+m)][o-2H]   nrows = fin
//   mode 2  gate, h part       kk = m*H + c;    Wg[((fin+c)*M+m)][n]                                            nrows = 2H
//   mode 3  candidate, h part  kk = m*H + c;    Wc[((fin+c)*M+m)][n]                                            nrows = H
//   mode 4  BPTT B1 (d(rh))    kk = m*H + o;    n = hidden column c:  Wc[((fin+n)*M+m)][o]                      nrows = H
//   mode 5  BPTT B2 (dh)       kk = (2m+g)*H + o (g = 0: r, 1: u);  Wg[((fin+n)*M+m)][g*H + o]                  nrows = H
//   mode 6  decoder projection (nn.Linear, model/model.py:146,193): kk = c < H;  Wg = proj_w (fin = Fo rows, H): Wg[n][kk]   nrows >= Fo
//   mode 7  its input gradient:  kk = o < fin = Fo;  n = hidden column:  Wg[kk][n]                                  nrows = H
// rows / k beyond the real extents are zero
__global__ void pack_w16_kernel(const float* Wg, const float* Wc, int fin, int H, int M, int mode, int nrows, int nq, uint8_t* img) {
    const int q = blockIdx.x;
    const size_t piece = (size_t)nrows * 128;
    uint8_t* hi = img + (size_t)(2 * q) * piece;
    uint8_t* lo = hi + piece;
    const int H2 = 2 * H, H3 = 3 * H;
    // (blockIdx.y splits the rows of a piece: with one CTA per chunk the launch was a ~10 us latency chain on 3 SMs)
    for (int idx = blockIdx.y * blockDim.x + threadIdx.x; idx < nrows * 64; idx += gridDim.y * blockDim.x) {
        const int n = idx >> 6, k = idx & 63, kk = 64 * q + k;
        float w = 0.f;
        if (mode == 0) {
            if (kk < M * fin) { const int m = kk / fin, c = kk - m * fin; const size_t r = (size_t)c * M + m;
                                w = n < H2 ? Wg[r * H2 + n] : Wc[r * H + (n - H2)]; }
        } else if (mode == 1) {
            if (kk < M * H3 && n < fin) { const int m = kk / H3, o = kk - m * H3; const size_t r = (size_t)n * M + m;
                               w = o < H2 ? Wg[r * H2 + o] : Wc[r * H + (o - H2)]; }
        } else if (mode == 2 || mode == 3) {
            if (kk < M * H) { const int m = kk / H, c = kk - m * H; const size_t r = (size_t)(fin + c) * M + m;
                              w = mode == 2 ? Wg[r * H2 + n] : Wc[r * H + n]; }
        } else if (mode == 4) {
            if (kk < M * H) { const int m = kk / H, o = kk - m * H; w = Wc[((size_t)(fin + n) * M + m) * H + o]; }
        } else if (mode == 5) {
            if (kk < 2 * M * H) { const int mg = kk / H, o = kk - mg * H, m = mg >> 1, g = mg & 1;
                                  w = Wg[((size_t)(fin + n) * M + m) * H2 + g * H + o]; }
        } else if (mode == 6) {
            if (kk < H && n < fin) w = Wg[(size_t)n * H + kk];
        } else {
            if (kk < fin && n < H) w = Wg[(size_t)kk * H + n];
        }
        __half h, l;
        split1(w, h, l);
        const uint32_t off = k128_off(n, k);
        *reinterpret_cast<__half*>(hi + off) = h;
        *reinterpret_cast<__half*>(lo + off) = l;
    }
}

cudaError_t launch_pack_w16(const float* Wg, const float* Wc, int fin, int H, int M, int mode, int nrows, int nq,
                            void* img, cudaStream_t st) {
    pack_w16_kernel<<<dim3(nq, 8), 256, 0, st>>>(Wg, Wc, fin, H, M, mode, nrows, nq, reinterpret_cast<uint8_t*>(img));
    return cudaGetLastError();
}

// ---- geometry shared with rnn_fwd.cu / rnn_bwd.cu / dw_mm16.cu ----------------------------------------------------------
int g16_nq(int cin, int M) { return (M * cin + 63) / 64; }
int g16_kxp(int fin, int M) { return g16_nq(fin, M) * 64; }
int g16_kkp(int fin, int H, int M) { return g16_kxp(fin, M) + 2 * M * H; }
int g16_ntile(int B) { return (B + SB - 1) / SB; }
size_t g16_image_bytes(int B, int T, int cols) { return (size_t)g16_ntile(B) * T * 2 * IMG_ROWS * cols * 2; }
size_t bulk_wimg_bytes(int cin, int M, int nout) { return (size_t)g16_nq(cin, M) * 2 * nout * 128; }

// epiw: room for the four staging tiles of the epilogue warps
static bool bulk_layout(int N, int Cin, int M, int Nout, int smem_limit, bool src16, BulkParams* p, bool epiw = false) {
    p->piece_bytes = Nout * 128;
    const int xbytes = (((src16 ? 2 * IMG_ROWS * Cin * 2 : SB * N * Cin * 4) + 1023) / 1024) * 1024;
    const int ptbytes = ((SB * (M - 1) * PT_STRIDE * 4 + 15) / 16) * 16;
    for (int nw = 8; nw >= 2; --nw) {
        int off = BK_NS * SLOT;
        p->off_w = off; off += ((nw * p->piece_bytes + 1023) / 1024) * 1024;
        p->off_x = off; off += xbytes;
        p->off_pt = off; off += ptbytes;
        p->off_id = off; off += PT_STRIDE * 4;
        p->off_stg = off; off += epiw ? 4 * BK_STG : 0;
        if (off + 1024 + 1024 <= smem_limit) { p->NW = nw; return true; }      // + alignment slack + static shared memory
    }
    return false;
}
static int bulk_smem(const BulkParams& p, bool epiw) { return p.off_stg + (epiw ? 4 * BK_STG : 0) + 1024; }

bool bulk_dp_supported(int N, int Cin, int M, int Nout, bool src16, int smem_limit) {
    BulkParams p;
    if (N > NPAD || Cin % (src16 ? 8 : 4) || M < 1 || g16_nq(Cin, M) > BK_MAXQ) return false;
    if (Nout != 64 && Nout != 128 && Nout != 192) return false;
    return bulk_layout(N, Cin, M, Nout, smem_limit, src16, &p);
}

// out (+)= ... see the header comment.  img: operand image base or nullptr; img_cols: floats.. fp16 values per image row
cudaError_t launch_bulk_dp(int B, int T, int N, int Cin, int M, int Nout, int transposeP, const float* src, long long ss_t,
                           long long ss_b, const void* src16, const float* P, const void* wimg, const float* bias, float* out, long long os_t,
                           long long os_b, int out_ld, float out_scale, const float* scale_ptr, void* img, int img_cols,
                           int img_col0, int nsms, int smem_limit, cudaStream_t st, const BulkExtra* ex) {
    BulkParams p;
    memset(&p, 0, sizeof p);
    // dedicated epilogue warps: measured at config 2 -- x pre-projection 1.13 -> 0.99 ms (wide outputs: 3H columns per row), dX
    // 0.75 -> 0.86 ms (64 output columns: the epilogue is small and the ring loses a slot) -> used for fp32 sources with
    // Nout = 192 when the staging tiles fit beside a 2-slot weight ring; DCGRU_BULK_EPIW=0 / 1 forces it off / on
    bool epiw = src16 == nullptr && Nout == 192;
    { const char* e = getenv("DCGRU_BULK_EPIW"); if (e && (e[0] == '0' || e[0] == '1')) epiw = e[0] == '1'; }
    if (epiw && !bulk_layout(N, Cin, M, Nout, smem_limit, src16 != nullptr, &p, true)) epiw = false;
    if (!epiw && !bulk_layout(N, Cin, M, Nout, smem_limit, src16 != nullptr, &p, false)) return cudaErrorInvalidConfiguration;
    p.src16 = reinterpret_cast<const __half*>(src16);
    p.B = B; p.T = T; p.N = N; p.Cin = Cin; p.M = M; p.Nout = Nout; p.transposeP = transposeP;
    p.ntile = g16_ntile(B); p.NQ = g16_nq(Cin, M);
    p.src = src; p.ss_t = ss_t; p.ss_b = ss_b; p.P = P; p.wimg = reinterpret_cast<const uint8_t*>(wimg); p.bias = bias;
    p.out = out; p.os_t = os_t; p.os_b = os_b; p.out_ld = out_ld; p.out_scale = out_scale; p.scale_ptr = scale_ptr;
    p.dump = img != nullptr; p.img_col0 = img_col0;
    // measured at config 2 (B200): no gain for the x pre-projection (1.14 ms either way) and dX 0.78 -> 0.84 ms (the row-major
    // image tile makes the ldmatrix loads 8-way bank-conflicted), so the FMA loop stays the default here
    { const char* e = getenv("DCGRU_MMA_DIFF_BULK"); p.mma_diff = (Cin % 64 == 0) && e && e[0] == '1'; }
    p.nout_valid = Nout; p.img_T = T; p.img_t0 = 0; p.src_T = T; p.src_t0 = 0;
    if (ex) {
        if (ex->nout_valid > 0) p.nout_valid = ex->nout_valid;
        if (ex->img_T > 0) { p.img_T = ex->img_T; p.img_t0 = ex->img_t0; }
        if (ex->src_T > 0) { p.src_T = ex->src_T; p.src_t0 = ex->src_t0; }
        p.in_scale_ptr = ex->in_scale_ptr;
    }
    CUtensorMap tm;
    memset(&tm, 0, sizeof tm);
    if (img) {
        // 2-D view of the fp16 image [rows][img_cols]; box = 64 columns (128 bytes) x the 24 rows of one sample
        const unsigned long long dims[2] = {(unsigned long long)img_cols, (unsigned long long)p.ntile * p.img_T * 2 * IMG_ROWS};
        const unsigned long long str[2] = {2, (unsigned long long)img_cols * 2};
        const unsigned box[2] = {64, RG * 8};
        cudaError_t e = make_tmap(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, img, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (e != cudaSuccess) return e;
    }
    const int smem = bulk_smem(p, epiw);
    cudaError_t e = epiw ? cudaFuncSetAttribute(bulk_dp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                         : cudaFuncSetAttribute(bulk_dp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    long nitems = (long)p.ntile * T;
    int grid = nsms < nitems ? nsms : (int)nitems;
    if (epiw) bulk_dp_kernel<true><<<grid, BK_THREADS, smem, st>>>(p, tm);
    else bulk_dp_kernel<false><<<grid, BK_THREADS - 128, smem, st>>>(p, tm);
    return cudaGetLastError();
}

}  // namespace dcgru
