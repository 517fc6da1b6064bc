// Weight gradient as a pure TMA-fed tensor-core GEMM (tcgen05, 3xTF32) over the operand images that the
// forward and backward sequence kernels leave in HBM:
//     dW[kk][o] = sum over every (cta, t, row) of  G[row][kk] * dA[row][o]
//   G  image (seq_fwd_tc.cu):  [slab = cta*T + t][hi|lo][row group of 8][kg of the step][8 rows x 16 B]
//   dA image (seq_bwd_tc.cu):  [slab            ][hi|lo][row group of 8][o/4, r|u|c   ][8 rows x 16 B]
// One K block = one row group (8 rows) of one slab: for any range of kg / column quads it is ONE contiguous
// piece of HBM, and once it sits in shared memory it already is a canonical MN-major no-swizzle UMMA operand
// (4 consecutive kk -- or o -- in 16 bytes, the 8 rows 16 bytes apart, 128 bytes between quads).  So this kernel
// has no producer threads: one thread issues cp.async.bulk loads into a 4-stage ring, one thread issues the
// MMAs, eight warps only wake up to flush the TMEM accumulators into the CTA's split-K partial.
//
// Work split: the M dimension (kk, in K groups of 4) is cut into 128-row tiles at multiples of 32 kg; a tile
// that straddles the x / gate-h / candidate-h parts of the image simply takes the union of the dA columns
// those parts need (rows x columns that mean nothing are never read back).  Tiles are packed into "sets" of
// at most 512 TMEM columns; a CTA owns one set and a contiguous range of K blocks, so every operand byte is
// read once per set.  Sets get CTAs in proportion to their MMA cost; all CTAs run in one wave.
// The tensor core truncates when it adds into the fp32 accumulator (bias ~2e-8 per accumulation), so TMEM is
// flushed into the partial every DWMM_FLUSH K blocks (640 rows), like dw_tc.cu.  dwmm_reduce_kernel then
// sums the partials of each set in fixed order (deterministic) straight into dWg / dWc.
// db is a plain column sum of the row-major dA (colsum kernels below).
#include "common.cuh"
#include "dw.cuh"
#include "tc_common.cuh"

namespace dcgru {
using namespace tc;

constexpr int DWMM_NSTAGE = 4;
constexpr int DWMM_FLUSH = 80;                       // K blocks between flushes (640 rows)
constexpr int DWMM_THREADS = 320;                    // warp 0: loader, warp 1: MMA issuer, warps 2-9: flush
constexpr int DWMM_TILE_BYTES = 32 * 128;            // one of hi / lo of a 32-kg tile
constexpr int DWMM_STAGE_BYTES = DWMM_MAXTILE * 2 * DWMM_TILE_BYTES + 2 * 48 * 128;   // 44 KB
constexpr int DWMM_SMEM = DWMM_NSTAGE * DWMM_STAGE_BYTES;

__global__ void __launch_bounds__(DWMM_THREADS, 1) dw_mm_kernel(const DwmmParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar_full[DWMM_NSTAGE], bar_empty[DWMM_NSTAGE], bar_accfull, bar_accempty;
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t desc[DWMM_NSTAGE][DWMM_MAXTILE][4];          // [stage][tile][A hi, A lo, B hi, B lo]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // which set / split am I
    int si = 0;
    while (si + 1 < p.nset && (int)blockIdx.x >= p.set[si + 1].cta0) ++si;
    const DwmmSet& S = p.set[si];
    const int split = blockIdx.x - S.cta0;
    const long kb0 = p.nkb * split / S.ncta, kb1 = p.nkb * (split + 1) / S.ncta;
    const int nkb = (int)(kb1 - kb0);
    const int b_bytes = S.ogcnt * 128;                                // one of hi / lo of the dA block

    if (warp == 0) tmem_alloc<512>(&tmem_slot);
    if (tid == 0) {
        for (int i = 0; i < DWMM_NSTAGE; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
        mbar_init(&bar_accfull, 1);
        mbar_init(&bar_accempty, 8);
        mbar_fence_init();
        for (int st = 0; st < DWMM_NSTAGE; ++st) {
            const uint32_t base = smem_u32(smem + st * DWMM_STAGE_BYTES);
            const uint32_t bbase = base + S.ntile * 2 * DWMM_TILE_BYTES;
            for (int j = 0; j < S.ntile; ++j) {
                // MN-major, one 8-row K group per MMA: quads are 128 B apart (both offset fields carry it)
                desc[st][j][0] = make_smem_desc(base + (2 * j) * DWMM_TILE_BYTES, 128, 128);
                desc[st][j][1] = make_smem_desc(base + (2 * j + 1) * DWMM_TILE_BYTES, 128, 128);
                const uint32_t bo = (S.tile[j].og0 - S.ogmin) * 128;
                desc[st][j][2] = make_smem_desc(bbase + bo, 128, 128);
                desc[st][j][3] = make_smem_desc(bbase + b_bytes + bo, 128, 128);
            }
        }
    }
    for (int idx = tid; idx < DWMM_SMEM / 16; idx += DWMM_THREADS)
        reinterpret_cast<float4*>(smem)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = tmem_slot;
    const int nflush = (nkb + DWMM_FLUSH - 1) / DWMM_FLUSH;

    if (warp == 0) {
        // =================================== loader ================================================================
        if (lane == 0) {
            const size_t g_rg = (size_t)p.KGT * 128, g_part = 16 * g_rg;            // row group / hi-lo part of G
            const size_t d_rg = (size_t)48 * 128, d_part = 16 * d_rg;
            uint32_t tx = 2 * b_bytes;
            for (int j = 0; j < S.ntile; ++j) tx += 2 * S.tile[j].nkg * 128;
            for (int i = 0; i < nkb; ++i) {
                const int st = i % DWMM_NSTAGE;
                if (i >= DWMM_NSTAGE) mbar_wait(&bar_empty[st], ((i / DWMM_NSTAGE) - 1) & 1);
                const long kb = kb0 + i;
                const size_t slab = (size_t)(kb >> 4);
                const int rg = (int)(kb & 15);
                uint8_t* sbase = smem + st * DWMM_STAGE_BYTES;
                mbar_expect_tx(&bar_full[st], tx);
                const uint8_t* g = p.G + slab * 2 * g_part + rg * g_rg;
                for (int j = 0; j < S.ntile; ++j) {
                    const uint32_t bytes = S.tile[j].nkg * 128;
                    bulk_g2s(sbase + (2 * j) * DWMM_TILE_BYTES, g + (size_t)S.tile[j].kg0 * 128, bytes, &bar_full[st]);
                    bulk_g2s(sbase + (2 * j + 1) * DWMM_TILE_BYTES, g + g_part + (size_t)S.tile[j].kg0 * 128, bytes,
                             &bar_full[st]);
                }
                const uint8_t* d = p.DA + slab * 2 * d_part + rg * d_rg + (size_t)S.ogmin * 128;
                uint8_t* sb = sbase + S.ntile * 2 * DWMM_TILE_BYTES;
                bulk_g2s(sb, d, b_bytes, &bar_full[st]);
                bulk_g2s(sb + b_bytes, d + d_part, b_bytes, &bar_full[st]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // =================================== MMA issuer ============================================================
        if (lane == 0) {
            uint32_t idesc[DWMM_MAXTILE], dcol[DWMM_MAXTILE];
            for (int j = 0; j < DWMM_MAXTILE; ++j) {
                idesc[j] = make_idesc_tf32_mn(128, j < S.ntile ? S.tile[j].ncol : 64);
                dcol[j] = taddr + (j < S.ntile ? S.tile[j].tcol : 0);
            }
            int since = 0, fl = 0;
            for (int i = 0; i < nkb; ++i) {
                const int st = i % DWMM_NSTAGE;
                mbar_wait(&bar_full[st], (i / DWMM_NSTAGE) & 1);
                if (since == 0 && fl > 0) mbar_wait(&bar_accempty, (fl - 1) & 1);   // the flush has read TMEM
                tc_fence_after();
                const uint32_t acc = since > 0 ? 1u : 0u;
                for (int j = 0; j < S.ntile; ++j) {
                    const uint64_t ah = desc[st][j][0], al = desc[st][j][1], bh = desc[st][j][2], bl = desc[st][j][3];
                    umma_tf32(dcol[j], al, bh, idesc[j], acc);      // small terms first
                    umma_tf32(dcol[j], ah, bl, idesc[j], 1u);
                    umma_tf32(dcol[j], ah, bh, idesc[j], 1u);
                }
                umma_commit(&bar_empty[st]);
                if (++since == DWMM_FLUSH || i == nkb - 1) { umma_commit(&bar_accfull); since = 0; ++fl; }
            }
        }
        __syncwarp();
    } else {
        // =================================== flush warps ============================================================
        // TMEM lane = tile row, so thread = row; the partial is stored [column][row]: a warp store is one 128-byte line
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const int row = 32 * quad + lane;
        const int ncolh = S.ncoltot / 2;
        float* part = p.part + (size_t)blockIdx.x * 512 * 128 + row;
        for (int f = 0; f < nflush; ++f) {
            mbar_wait(&bar_accfull, f & 1);
            tc_fence_after();
            for (int cb = half * ncolh; cb < (half + 1) * ncolh; cb += 16) {
                float v[16];
                tmem_ld16(taddr + ((uint32_t)(32 * quad) << 16) + cb, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float* q = part + (size_t)(cb + j) * 128;
                    *q = (f > 0) ? *q + v[j] : v[j];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_accempty);
        }
        if (nflush == 0)                                                // CTA without work: zero partial
            for (int cb = half * ncolh; cb < (half + 1) * ncolh; ++cb) part[(size_t)cb * 128] = 0.f;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(taddr);
}

// dWg / dWc <- fixed-order sum of the partials.  One thread per (W row, column of r|u|c).
__global__ void dwmm_reduce_kernel(const DwmmParams p, int fin, int H, int M, float* dWg, float* dWc) {
    const int CM = (fin + H) * M, H3 = 3 * H;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= CM * H3) return;
    const int o = idx / CM, r = idx - o * CM;                          // consecutive threads -> consecutive rows
    const int nx = ((fin + 7) / 8) * 8 * M;                            // kk of the x part incl. chunk padding
    int gk;                                                            // kk position inside the G image
    if (r < fin * M) gk = r;
    else gk = nx + (r - fin * M) + (o < 2 * H ? 0 : H * M);
    const int kg0 = (gk >> 7) * 32, rt = gk & 127;
    float acc = 0.f;
    for (int s = 0; s < p.nset; ++s)
        for (int j = 0; j < p.set[s].ntile; ++j) {
            const DwmmTile& t = p.set[s].tile[j];
            if (t.kg0 != kg0) continue;
            const int c = t.tcol + o - 4 * t.og0;
            const float* q = p.part + ((size_t)p.set[s].cta0 * 512 + c) * 128 + rt;
            for (int i = 0; i < p.set[s].ncta; ++i) acc += q[(size_t)i * 512 * 128];
        }
    if (o < 2 * H) dWg[(size_t)r * 2 * H + o] = acc;
    else dWc[(size_t)r * H + (o - 2 * H)] = acc;
}

// ---- db: column sums of the row-major dA (rows x 3H), two fixed-order stages ------------------------------------
constexpr int CS_CTAS = 592;
__global__ void colsum_kernel(const float* dA, long rows, int H3, float* partial) {
    const int c = threadIdx.x;
    const long r0 = rows * blockIdx.x / gridDim.x, r1 = rows * (blockIdx.x + 1) / gridDim.x;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    long r = r0;
    for (; r + 3 < r1; r += 4) {
        a0 += dA[(size_t)r * H3 + c];
        a1 += dA[(size_t)(r + 1) * H3 + c];
        a2 += dA[(size_t)(r + 2) * H3 + c];
        a3 += dA[(size_t)(r + 3) * H3 + c];
    }
    for (; r < r1; ++r) a0 += dA[(size_t)r * H3 + c];
    partial[(size_t)blockIdx.x * H3 + c] = (a0 + a1) + (a2 + a3);
}
__global__ void colsum_final_kernel(const float* partial, int n, int H, float* dbg, float* dbc) {
    const int c = threadIdx.x;
    float a = 0.f;
    for (int i = 0; i < n; ++i) a += partial[(size_t)i * 3 * H + c];
    if (c < 2 * H) dbg[c] = a; else dbc[c - 2 * H] = a;
}

// ---- host side ------------------------------------------------------------------------------------------------------
// Tiles and sets for one cell (H = 64, M = 3: what seq_fwd_tc / seq_bwd_tc support).
bool dwmm_plan(int fin, int H, int M, long nslab, int nsms, DwmmParams* out) {
    if (H != 64 || M != 3 || nsms < 2) return false;
    const int nxc = (fin + 7) / 8;
    const int kgx = nxc * 6, kgt = kgx + 96;                            // x | gate h | candidate h
    DwmmParams& p = *out;
    p.KGT = kgt;
    p.nkb = nslab * 16;
    p.nset = 0;
    const int ntile = (kgt + 31) / 32;
    DwmmTile tiles[16];
    if (ntile > 16) return false;
    for (int j = 0; j < ntile; ++j) {
        DwmmTile& t = tiles[j];
        t.kg0 = 32 * j;
        t.nkg = kgt - t.kg0 < 32 ? kgt - t.kg0 : 32;
        const int a = t.kg0, b = t.kg0 + t.nkg;                           // [a, b)
        const bool ox = a < kgx, og = a < kgx + 48 && b > kgx, oc = b > kgx + 48;
        int lo = 48, hi = 0;
        if (ox) { lo = 0; hi = 48; }
        if (og) { lo = 0; if (hi < 32) hi = 32; }
        if (oc) { if (lo > 32) lo = 32; hi = 48; }
        t.og0 = lo; t.ncol = 4 * (hi - lo); t.tcol = 0;
    }
    // first-fit decreasing into sets of <= 512 TMEM columns and <= DWMM_MAXTILE tiles
    bool used[16] = {false};
    for (int placed = 0; placed < ntile;) {
        if (p.nset == DWMM_MAXSET) return false;
        DwmmSet& S = p.set[p.nset++];
        S.ntile = 0; S.ncoltot = 0;
        for (int want = 192; want >= 64; want -= 64)
            for (int j = 0; j < ntile; ++j)
                if (!used[j] && tiles[j].ncol == want && S.ncoltot + want <= 512 && S.ntile < DWMM_MAXTILE) {
                    used[j] = true; ++placed;
                    DwmmTile t = tiles[j];
                    t.tcol = S.ncoltot;
                    S.ncoltot += want;
                    S.tile[S.ntile++] = t;
                }
        int lo = 48, hi = 0;
        for (int j = 0; j < S.ntile; ++j) {
            if (S.tile[j].og0 < lo) lo = S.tile[j].og0;
            if (S.tile[j].og0 + S.tile[j].ncol / 4 > hi) hi = S.tile[j].og0 + S.tile[j].ncol / 4;
        }
        S.ogmin = lo; S.ogcnt = hi - lo;
    }
    // CTAs per set in proportion to the MMA cost (columns, plus a per-instruction overhead)
    int cost[DWMM_MAXSET], tot = 0;
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
    return true;
}

size_t dwmm_part_floats(int nsms) { return (size_t)nsms * 512 * 128; }
size_t colsum_part_floats(int H) { return (size_t)CS_CTAS * 3 * H; }

cudaError_t launch_dw_mm(const DwmmParams& p, int fin, int H, int M, float* dWg, float* dWc, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(dw_mm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DWMM_SMEM);
    if (e != cudaSuccess) return e;
    dw_mm_kernel<<<p.ncta, DWMM_THREADS, DWMM_SMEM, st>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int n = (fin + H) * M * 3 * H;
    dwmm_reduce_kernel<<<(n + 255) / 256, 256, 0, st>>>(p, fin, H, M, dWg, dWc);
    return cudaGetLastError();
}

cudaError_t launch_colsum(const float* dA, long rows, int H, float* partial, float* dbg, float* dbc, cudaStream_t st) {
    colsum_kernel<<<CS_CTAS, 3 * H, 0, st>>>(dA, rows, 3 * H, partial);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    colsum_final_kernel<<<1, 3 * H, 0, st>>>(partial, CS_CTAS, H, dbg, dbc);
    return cudaGetLastError();
}

int dw_mm_smem_bytes() { return DWMM_SMEM; }

}  // namespace dcgru
