// Weight gradient of one cell as a pure TMA-fed tensor-core GEMM (tcgen05, 2xFP16) over the fp16 operand images left
// in HBM by the second-generation kernels:
//     dW[kk][o] = (1/s) * sum over every (tile, t, row) of  G[row][kk] * dA[row][o]
//   G  image: [slab = tile*T + t][hi|lo][96 rows][KKP]  columns = x part (kk = m*fin + c, padded to 64; bulk_dp.cu) |
//                                                                gate-h chunks (m*H + c) | candidate-h chunks (rnn_fwd.cu)
//   dA image: [slab][hi|lo][96 rows][3H]                 columns r | u | c, scaled by s (rnn_bwd.cu)
// The GEMM's K index is the image row, so both operands are MN-major; for 16-bit types that is the canonical
// SWIZZLE_128B MN-major layout (rows of 64 values = 128 B, 8-row atoms), which is exactly what a tensor-map TMA load
// with a 64-column box and CU_TENSOR_MAP_SWIZZLE_128B writes.  No producer threads: warp 0 issues two TMA loads per
// K block of 16 image rows (one 4-D box with all the G column blocks of the CTA's tile set, hi and lo; one with the dA
// columns) into a 4-stage ring, warp 1 issues 3 kind::f16 MMAs per tile and K block, warps 2-9 flush TMEM into the
// CTA's split-K partial.  Same work split as the first-generation dw_mm.cu: M (= kk) in 128-row tiles, contiguous
// tiles packed into sets of <= 512 TMEM columns, a CTA owns a set and a K range, one wave.
// Half the bytes per image row and twice the rows per MMA of the 3xTF32 version.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "dw.cuh"
#include "f16_common.cuh"
#include "tmap.cuh"

namespace dcgru {
using namespace f16;

constexpr int D16_NSTAGE = 4;
constexpr int D16_KROWS = 16;                         // image rows per K block (one MMA k-step)
constexpr int D16_KB_PER_SLAB = IMG_ROWS / D16_KROWS; // 6
constexpr int D16_FLUSH = 80;                         // K blocks between flushes (1280 rows)
constexpr int D16_THREADS = 352;                       // loader, issuer, 8 flush warps, column-sum warp
constexpr int D16_BLK = D16_KROWS * 128;              // one 64-column block of one plane: 16 rows x 128 B = 2 KB
constexpr int D16_BOXBLK = 4;                         // G column blocks per TMA box (2 tiles)
constexpr int D16_BOX = 2 * D16_BOXBLK * D16_BLK;     // [hi | lo][4 blocks] = 16 KB
constexpr int D16_A_BYTES = 2 * D16_BOX;              // up to two boxes (4 tiles) = 32 KB
constexpr int D16_B_BYTES = 2 * 3 * D16_BLK;          // [hi | lo][3 blocks] = 12 KB
constexpr int D16_STAGE = D16_A_BYTES + D16_B_BYTES;  // 44 KB
constexpr int D16_SMEM = D16_NSTAGE * D16_STAGE + 1024;

struct Dw16Tile { int soff, ncol, ob0, tcol; };       // byte offset of the tile's hi part in a stage (lo: + 4 blocks), dA columns, first dA block, TMEM column
struct Dw16Set { int ntile, ncoltot, cta0, ncta, gblk0, nbox; Dw16Tile tile[4]; };
struct Dw16Params {
    const uint8_t* G; const uint8_t* DA;
    float* part;                                      // [cta][512 columns][128 rows]
    float* dbpart;                                    // [CTA of set 0][192]: column sums of the dA image over the CTA's K range (db), or nullptr
    const float* scale_ptr;
    long nkb;
    int kkp, nset, ncta;
    Dw16Set set[8];
};

__global__ void __launch_bounds__(D16_THREADS, 1) dw_mm16_kernel(const Dw16Params p, const __grid_constant__ CUtensorMap tm_g,
                                                                 const __grid_constant__ CUtensorMap tm_d) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_full[D16_NSTAGE], bar_empty[D16_NSTAGE], bar_accfull, bar_accempty;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    int si = 0;
    while (si + 1 < p.nset && (int)blockIdx.x >= p.set[si + 1].cta0) ++si;
    const Dw16Set& S = p.set[si];
    const int ntile = S.ntile;
    const int split = blockIdx.x - S.cta0;
    const long kb0 = p.nkb * split / S.ncta, kb1 = p.nkb * (split + 1) / S.ncta;
    const int nkb = (int)(kb1 - kb0);

    if (warp == 0) tmem_alloc<512>(&tmem_slot);
    if (tid == 0) {
        // (in the CTAs of set 0 a stage has two consumers when db is fused: the MMAs and the column-sum warp)
        for (int i = 0; i < D16_NSTAGE; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], (p.dbpart && si == 0) ? 2 : 1); }
        mbar_init(&bar_accfull, 1);
        mbar_init(&bar_accempty, 8);
        mbar_fence_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = tmem_slot;
    const int nflush = (nkb + D16_FLUSH - 1) / D16_FLUSH;

    if (warp == 0) {
        // =================================== loader ================================================================
        if (lane == 0) {
            tma_prefetch_desc(&tm_g);
            tma_prefetch_desc(&tm_d);
            const uint32_t tx = S.nbox * D16_BOX + D16_B_BYTES;          // (out-of-bounds blocks are zero-filled and counted)
            for (int i = 0; i < nkb; ++i) {
                const int st = i % D16_NSTAGE;
                if (i >= D16_NSTAGE) mbar_wait(&bar_empty[st], ((i / D16_NSTAGE) - 1) & 1);
                const long kb = kb0 + i;
                const long slab = kb / D16_KB_PER_SLAB;
                const int row_hi = (int)(slab * 2 * IMG_ROWS + (kb - slab * D16_KB_PER_SLAB) * D16_KROWS);   // lo plane: 96 rows further (dim 3)
                uint8_t* sbase = smem + st * D16_STAGE;
                mbar_expect_tx(&bar_full[st], tx);
                tma_load_4d(sbase, &tm_g, 0, row_hi, S.gblk0, 0, &bar_full[st]);
                if (S.nbox > 1) tma_load_4d(sbase + D16_BOX, &tm_g, 0, row_hi, S.gblk0 + D16_BOXBLK, 0, &bar_full[st]);
                tma_load_4d(sbase + D16_A_BYTES, &tm_d, 0, row_hi, 0, 0, &bar_full[st]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // =================================== MMA issuer ============================================================
        if (lane == 0) {
            uint64_t dah[4], dal[4], dbh[4], dbl[4];
            uint32_t idesc[4], dcol[4];
            const uint32_t base = smem_u32(smem), bbase = base + D16_A_BYTES;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const Dw16Tile& t = S.tile[j < ntile ? j : 0];
                dah[j] = make_desc_mn128(base + t.soff, D16_BLK, 1024);
                dal[j] = make_desc_mn128(base + t.soff + D16_BOXBLK * D16_BLK, D16_BLK, 1024);
                dbh[j] = make_desc_mn128(bbase + t.ob0 * D16_BLK, D16_BLK, 1024);
                dbl[j] = make_desc_mn128(bbase + 3 * D16_BLK + t.ob0 * D16_BLK, D16_BLK, 1024);
                idesc[j] = make_idesc_f16_mn(128, t.ncol);
                dcol[j] = taddr + t.tcol;
            }
            int since = 0, fl = 0, st = 0, ph = 0;
            for (int i = 0; i < nkb; ++i) {
                mbar_wait(&bar_full[st], ph);
                if (since == 0 && fl > 0) mbar_wait(&bar_accempty, (fl - 1) & 1);
                tc_fence_after();
                const uint32_t acc = since > 0 ? 1u : 0u;
                const uint64_t so = (uint64_t)((st * D16_STAGE) >> 4);
#pragma unroll
                for (int j = 3; j >= 0; --j)
                    if (j < ntile) {
                        umma_f16(dcol[j], dal[j] + so, dbh[j] + so, idesc[j], acc);
                        umma_f16(dcol[j], dah[j] + so, dbl[j] + so, idesc[j], 1u);
                        umma_f16(dcol[j], dah[j] + so, dbh[j] + so, idesc[j], 1u);
                    }
                umma_commit(&bar_empty[st]);
                if (++since == D16_FLUSH || i == nkb - 1) { umma_commit(&bar_accfull); since = 0; ++fl; }
                if (++st == D16_NSTAGE) { st = 0; ph ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == 10) {
        // =================================== column sums of dA (db) ====================================================
        // The dA rows of every K block pass through shared memory anyway; the CTAs of set 0 cover every image row exactly once.
        // Lane u < 24 owns the 16-byte unit u of a row (8 columns): 16 rows x (hi + lo) per K block, fp32 partial sums folded
        // into double every D16_FLUSH blocks.  The tiles are SWIZZLE_128B: unit cu of row k sits at unit cu ^ (k & 7).
        if (p.dbpart && si == 0) {
            float a[8];
            double d[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { a[j] = 0.f; d[j] = 0.0; }
            const int blk = lane >> 3, cu = lane & 7;
            int st = 0, ph = 0, since = 0;
            for (int i = 0; i < nkb; ++i) {
                mbar_wait(&bar_full[st], ph);
                if (lane < 24) {
                    const uint8_t* b = smem + st * D16_STAGE + D16_A_BYTES + blk * D16_BLK;
#pragma unroll
                    for (int pl = 0; pl < 2; ++pl)
#pragma unroll
                        for (int k = 0; k < D16_KROWS; ++k) {
                            const uint4 q = *reinterpret_cast<const uint4*>(b + pl * 3 * D16_BLK + k * 128 + ((cu ^ (k & 7)) << 4));
                            const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
                            for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h[j]); a[2 * j] += f.x; a[2 * j + 1] += f.y; }
                        }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_empty[st]);
                if (++since == D16_FLUSH) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) { d[j] += (double)a[j]; a[j] = 0.f; }
                    since = 0;
                }
                if (++st == D16_NSTAGE) { st = 0; ph ^= 1; }
            }
            if (lane < 24) {
#pragma unroll
                for (int j = 0; j < 8; ++j) p.dbpart[(size_t)split * 192 + 8 * lane + j] = (float)(d[j] + (double)a[j]);
            }
        }
    } else {
        // =================================== flush warps ============================================================
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const int row = 32 * quad + lane;
        const int csplit = ((S.ncoltot / 64 + 1) / 2) * 64;
        const int c_begin = half ? csplit : 0, c_end = half ? S.ncoltot : csplit;
        float* part = p.part + (size_t)blockIdx.x * 512 * 128 + row;
        for (int f = 0; f < nflush; ++f) {
            mbar_wait(&bar_accfull, f & 1);
            tc_fence_after();
            for (int cb = c_begin; cb < c_end; cb += 64) {
                float o[64];
                float* q = part + (size_t)cb * 128;
                if (f > 0) {
#pragma unroll
                    for (int j = 0; j < 64; ++j) o[j] = __ldcg(q + j * 128);
                } else {
#pragma unroll
                    for (int j = 0; j < 64; ++j) o[j] = 0.f;
                }
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    float v[16];
                    tmem_ld16(taddr + ((uint32_t)(32 * quad) << 16) + cb + 16 * c4, v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) __stcg(q + (16 * c4 + j) * 128, v[j] + o[16 * c4 + j]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_accempty);
        }
        if (nflush == 0)
            for (int cb = c_begin; cb < c_end; ++cb) part[(size_t)cb * 128] = 0.f;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(taddr);
}

// dWg / dWc <- (1/s) * fixed-order sum of the partials.  One thread per (W row, column of r|u|c).
__global__ void dw_mm16_reduce_kernel(const Dw16Params p, int fin, int H, int M, float* dWg, float* dWc) {
    const int CM = (fin + H) * M, H3 = 3 * H;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= CM * H3) return;
    const int o = idx / CM, r = idx - o * CM;                          // consecutive threads -> consecutive W rows
    const int c = r / M, m = r - c * M;                                // reference row order c*M + m
    const int kxp = ((M * fin + 63) / 64) * 64;
    int gk;                                                            // column of the G image
    if (c < fin) gk = m * fin + c;
    else gk = kxp + (o < 2 * H ? 0 : M * H) + m * H + (c - fin);
    const int blk = gk >> 6, rt = gk & 127;                            // tile = blk / 2
    float acc = 0.f;
    for (int s = 0; s < p.nset; ++s) {
        const Dw16Set& S = p.set[s];
        for (int j = 0; j < S.ntile; ++j) {
            const Dw16Tile& t = S.tile[j];
            if (((S.gblk0 >> 1) + j) != (blk >> 1)) continue;
            const int col = t.tcol + o - 64 * t.ob0;
            const float* q = p.part + ((size_t)S.cta0 * 512 + col) * 128 + rt;
            for (int i = 0; i < S.ncta; ++i) acc += q[(size_t)i * 512 * 128];
        }
    }
    acc *= 1.f / p.scale_ptr[0];
    if (o < 2 * H) dWg[(size_t)r * 2 * H + o] = acc;
    else dWc[(size_t)r * H + (o - 2 * H)] = acc;
}

// ---- db: column sums of the dA image (hi + lo), two fixed-order stages --------------------------------------------------
constexpr int CS16_CTAS = 296;
// thread = (8-column group, row slot): 16-byte loads (8 fp16), 8 row slots per CTA, several rows in flight per thread
__global__ void __launch_bounds__(192) colsum16_kernel(const __half* img, long nrows, float* partial) {
    __shared__ float red[8][192];
    const int cg = threadIdx.x % 24, slot = threadIdx.x / 24;         // 24 groups of 8 columns, 8 slots
    const long r0 = nrows * blockIdx.x / gridDim.x, r1 = nrows * (blockIdx.x + 1) / gridDim.x;
    const uint4* src = reinterpret_cast<const uint4*>(img) + cg;      // row pitch = 192 halfs = 24 uint4
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 0.f;
    auto add = [&](const uint4& q) {
        const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h[j]); a[2 * j] += f.x; a[2 * j + 1] += f.y; }
    };
    long r = r0 + slot;
    for (; r + 24 < r1; r += 32) {
        const uint4 q0 = __ldcs(src + (size_t)r * 24), q1 = __ldcs(src + (size_t)(r + 8) * 24), q2 = __ldcs(src + (size_t)(r + 16) * 24),
                    q3 = __ldcs(src + (size_t)(r + 24) * 24);
        add(q0); add(q1); add(q2); add(q3);
    }
    for (; r < r1; r += 8) add(__ldcs(src + (size_t)r * 24));
#pragma unroll
    for (int j = 0; j < 8; ++j) red[slot][8 * cg + j] = a[j];
    __syncthreads();
    float s = 0.f;
    for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
    partial[(size_t)blockIdx.x * 192 + threadIdx.x] = s;
}
__global__ void colsum16_final_kernel(const float* partial, int n, int H, const float* scale_ptr, float* dbg, float* dbc) {
    const int c = threadIdx.x;
    float a = 0.f;
    for (int i = 0; i < n; ++i) a += partial[(size_t)i * 3 * H + c];
    a *= 1.f / scale_ptr[0];
    if (c < 2 * H) dbg[c] = a; else dbc[c - 2 * H] = a;
}

// ---- host side ------------------------------------------------------------------------------------------------------
bool dw_mm16_plan(int fin, int H, int M, long nslab, int nsms, Dw16Params* out) {
    if (H != 64 || nsms < 2) return false;
    Dw16Params& p = *out;
    const int kxp = g16_kxp(fin, M), kkp = g16_kkp(fin, H, M);
    const int nblk = kkp / 64, ntile = (nblk + 1) / 2;
    const int g0 = kxp / 64, c0 = g0 + M;                              // first gate-h / candidate-h block
    p.kkp = kkp;
    p.nkb = nslab * D16_KB_PER_SLAB;
    p.nset = 0;
    struct T { int ncol, ob0; } tl[64];
    if (ntile > 64) return false;
    for (int j = 0; j < ntile; ++j) {
        const int a = 2 * j, b = (2 * j + 2 < nblk) ? 2 * j + 2 : nblk;          // blocks [a, b)
        const bool ox = a < g0, og = a < c0 && b > g0, oc = b > c0;
        int lo = 3, hi = 0;                                                       // dA column blocks [lo, hi)
        if (ox) { lo = 0; hi = 3; }
        if (og) { lo = 0; if (hi < 2) hi = 2; }
        if (oc) { if (lo > 2) lo = 2; hi = 3; }
        tl[j].ob0 = lo; tl[j].ncol = 64 * (hi - lo);
    }
    for (int j = 0; j < ntile;) {                                      // contiguous tiles, <= 4 per set, <= 512 TMEM columns
        if (p.nset == 8) return false;
        Dw16Set& S = p.set[p.nset++];
        S.ntile = 0; S.ncoltot = 0; S.gblk0 = 2 * j;
        while (j < ntile && S.ntile < 4 && S.ncoltot + tl[j].ncol <= 512) {
            Dw16Tile& t = S.tile[S.ntile];
            t.soff = (S.ntile / 2) * D16_BOX + (S.ntile % 2) * 2 * D16_BLK;
            t.ncol = tl[j].ncol; t.ob0 = tl[j].ob0; t.tcol = S.ncoltot;
            S.ncoltot += t.ncol; ++S.ntile; ++j;
        }
        S.nbox = (S.ntile + 1) / 2;
    }
    int cost[8], tot = 0;
    for (int s = 0; s < p.nset; ++s) { cost[s] = p.set[s].ncoltot + 24 * p.set[s].ntile; tot += cost[s]; }
    int left = nsms, cta0 = 0;
    for (int s = 0; s < p.nset; ++s) {
        int n = (s == p.nset - 1) ? left : (int)((long)nsms * cost[s] / tot);
        if (n < 1) n = 1;
        if (n > left - (p.nset - 1 - s)) n = left - (p.nset - 1 - s);
        if ((long)n > p.nkb) n = (int)p.nkb;
        p.set[s].cta0 = cta0; p.set[s].ncta = n;
        cta0 += n; left -= n;
    }
    p.ncta = cta0;
    return p.nset <= nsms;
}

size_t dw_mm16_part_floats(int nsms) { return (size_t)nsms * 512 * 128; }
size_t colsum16_part_floats(int H) { return (size_t)CS16_CTAS * 3 * H; }
int dw_mm16_smem_bytes() { return D16_SMEM; }

// dbpart != nullptr: db is fused (column sums of the dA rows by one more warp of the set-0 CTAs; dbpart holds >= sms * 192 floats)
cudaError_t launch_dw_mm16(int fin, int H, int M, int B, int T, const void* G, const void* DA, float* part, const float* scale_ptr,
                           int nsms, float* dWg, float* dWc, cudaStream_t st, float* dbpart, float* dbg, float* dbc) {
    Dw16Params p;
    memset(&p, 0, sizeof p);
    const long nslab = (long)g16_ntile(B) * T;
    if (!dw_mm16_plan(fin, H, M, nslab, nsms, &p)) return cudaErrorInvalidConfiguration;
    p.G = reinterpret_cast<const uint8_t*>(G); p.DA = reinterpret_cast<const uint8_t*>(DA); p.part = part; p.scale_ptr = scale_ptr;
    p.dbpart = dbpart;
    // 4-D views of the row-major fp16 images: (64 columns = one 128-byte row piece | image row | column block | hi / lo plane);
    // column blocks beyond the image width are out of bounds (zero-filled)
    CUtensorMap tg, td;
    const unsigned long long rows = (unsigned long long)nslab * 2 * IMG_ROWS;
    {
        const unsigned long long kkp = p.kkp;
        const unsigned long long dims[4] = {64, rows, kkp / 64, 2};
        const unsigned long long str[4] = {2, kkp * 2, 128, (unsigned long long)IMG_ROWS * kkp * 2};
        const unsigned box[4] = {64, D16_KROWS, D16_BOXBLK, 2};
        cudaError_t e = make_tmap(&tg, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, G, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (e != cudaSuccess) return e;
    }
    {
        const unsigned long long dims[4] = {64, rows, 3, 2};
        const unsigned long long str[4] = {2, (unsigned long long)3 * H * 2, 128, (unsigned long long)IMG_ROWS * 3 * H * 2};
        const unsigned box[4] = {64, D16_KROWS, 3, 2};
        cudaError_t e = make_tmap(&td, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, DA, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = cudaFuncSetAttribute(dw_mm16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, D16_SMEM);
    if (e != cudaSuccess) return e;
    dw_mm16_kernel<<<p.ncta, D16_THREADS, D16_SMEM, st>>>(p, tg, td);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int n = (fin + H) * M * 3 * H;
    dw_mm16_reduce_kernel<<<(n + 255) / 256, 256, 0, st>>>(p, fin, H, M, dWg, dWc);
    e = cudaGetLastError();
    if (e != cudaSuccess || !dbpart) return e;
    colsum16_final_kernel<<<1, 3 * H, 0, st>>>(dbpart, p.set[0].ncta, H, scale_ptr, dbg, dbc);
    return cudaGetLastError();
}

cudaError_t launch_colsum16(const void* daimg, int B, int T, int H, float* partial, const float* scale_ptr, float* dbg, float* dbc,
                            cudaStream_t st) {
    if (H != 64) return cudaErrorInvalidValue;
    colsum16_kernel<<<CS16_CTAS, 192, 0, st>>>(reinterpret_cast<const __half*>(daimg), (long)g16_ntile(B) * T * 2 * IMG_ROWS, partial);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    colsum16_final_kernel<<<1, 3 * H, 0, st>>>(partial, CS16_CTAS, H, scale_ptr, dbg, dbc);
    return cudaGetLastError();
}

}  // namespace dcgru
