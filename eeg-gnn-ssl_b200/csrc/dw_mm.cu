// Weight gradient as a pure TMA-fed tensor-core GEMM (tcgen05, 3xTF32) over the operand images that the
// forward and backward sequence kernels leave in HBM:
//     dW[kk][o] = sum over every (cta, t, row) of  G[row][kk] * dA[row][o]
//   G  image (seq_fwd_tc.cu):  [slab = cta*T + t][hi|lo][96 rows][KKP floats]   (row-major, kk order of the step)
//   dA image (seq_bwd_tc.cu):  [slab            ][hi|lo][96 rows][192 floats]   (row-major, columns r|u|c)
//   (96 rows = 4 samples x the 24 rows that can be non-zero: 12 K blocks of 8 rows per slab)
// Both operands of this GEMM are "MN-major" (the GEMM's K index is the image row, the contiguous index is kk / o).
// For fp32/tf32 the tensor core reads MN-major operands in exactly one layout -- 128-byte rows of 32 values whose
// 32-byte chunks are XOR-swizzled with the row index (tc_common.cuh) -- and that is what a TMA load with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B produces from a row-major image.  So this kernel has no producer threads:
// one thread issues tensor-map TMA loads into a 4-stage ring (one K block = 8 image rows), one thread issues the
// MMAs, eight warps only wake up to flush the TMEM accumulators into the CTA's split-K partial.
// A K block is fetched with at most three TMA instructions: one 4-D box (32 floats x 8 rows x 8 groups x hi|lo =
// 16 KB) per PAIR of adjacent tiles and one (12 KB) for the dA rows.  (With one box per tile and part -- ten
// instructions of ~70 issue cycles each plus a ~300-cycle barrier poll -- the single loader thread was as slow as the
// MMA issuer and the ring kept running dry: 3.8 TB/s with DRAM only 49 % busy.)
//
// Work split: the M dimension (kk) is cut into 128-row tiles; a tile that straddles the x / gate-h / candidate-h
// parts of the image simply takes the union of the dA columns those parts need (rows x columns that mean nothing
// are never read back).  Tiles are packed into "sets" of at most 512 TMEM columns; a CTA owns one set and a
// contiguous range of K blocks, so every operand byte is read once per set.  Sets get CTAs in proportion to their
// MMA cost; all CTAs run in one wave.
// The tensor core truncates when it adds into the fp32 accumulator (bias ~2e-8 per accumulation), so TMEM is
// flushed into the partial every DWMM_FLUSH K blocks (640 rows), like dw_tc.cu.  dwmm_reduce_kernel then
// sums the partials of each set in fixed order (deterministic) straight into dWg / dWc.
// db is a plain column sum of the row-major dA (colsum kernels below).
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "dw.cuh"
#include "tc_common.cuh"
#include "tmap.cuh"

namespace dcgru {
using namespace tc;

constexpr int DWMM_NSTAGE = 4;
constexpr int DWMM_KB_PER_SLAB = TC_IMG_ROWS / 8;   // K blocks (8 image rows) per (cta, t)
constexpr int DWMM_FLUSH = 80;                       // K blocks between flushes (640 rows)
constexpr int DWMM_THREADS = 320;                    // warp 0: loader, warp 1: MMA issuer, warps 2-9: flush
constexpr int DWMM_TILE_BYTES = 4 * 1024;            // one of hi / lo of a 128-kk tile: 4 groups x (8 rows x 128 B)
constexpr int DWMM_BOX_BYTES = 4 * DWMM_TILE_BYTES;  // one G box: [hi | lo][2 tiles]
constexpr int DWMM_B_BYTES = 6 * 1024;               // one of hi / lo of the dA block: 6 groups of 32 columns
constexpr int DWMM_B_OFF = 2 * DWMM_BOX_BYTES;       // stage = [box 0][box 1][dA hi | lo]
constexpr int DWMM_STAGE_BYTES = DWMM_B_OFF + 2 * DWMM_B_BYTES;   // 44 KB
constexpr int DWMM_SMEM = DWMM_NSTAGE * DWMM_STAGE_BYTES + 1024;                            // + alignment slack

// timing experiment (DCGRU_DBG & 8): clock64 stamps of CTA 0, [K block][loader: top, waited, issued | issuer: top, full, accfree, issued]
__device__ long long dwmm_dbg[256 * 8];

__global__ void __launch_bounds__(DWMM_THREADS, 1) dw_mm_kernel(const DwmmParams p, const __grid_constant__ CUtensorMap tm_g,
                                                                const __grid_constant__ CUtensorMap tm_d) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_full[DWMM_NSTAGE], bar_empty[DWMM_NSTAGE], bar_accfull, bar_accempty;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // the swizzled operand tiles need 1024-byte aligned shared memory
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    // which set / split am I
    int si = 0;
    while (si + 1 < p.nset && (int)blockIdx.x >= p.set[si + 1].cta0) ++si;
    const DwmmSet& S = p.set[si];
    const int ntile = S.ntile;
    const int split = blockIdx.x - S.cta0;
    const long kb0 = p.nkb * split / S.ncta, kb1 = p.nkb * (split + 1) / S.ncta;
    const int nkb = (int)(kb1 - kb0);

    if (warp == 0) tmem_alloc<512>(&tmem_slot);
    if (tid == 0) {
        for (int i = 0; i < DWMM_NSTAGE; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
        mbar_init(&bar_accfull, 1);
        mbar_init(&bar_accempty, 8);
        mbar_fence_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = tmem_slot;
    const int nflush = (nkb + DWMM_FLUSH - 1) / DWMM_FLUSH;

    if (warp == 0) {
        // =================================== loader ================================================================
        if (lane == 0) {
            tma_prefetch_desc(&tm_g);
            tma_prefetch_desc(&tm_d);
            const int nbox = S.nbox, grp0 = S.boxgrp[0], grp1 = S.boxgrp[1];
            const uint32_t tx = nbox * DWMM_BOX_BYTES + 2 * DWMM_B_BYTES;             // out-of-bounds groups are zero-filled and counted
            for (int i = 0; i < nkb; ++i) {
                const int st = i % DWMM_NSTAGE;
                const bool rec = p.dbg && blockIdx.x == 0 && i < 256;
                if (rec) dwmm_dbg[i * 8 + 0] = clock64();
                if (i >= DWMM_NSTAGE) mbar_wait(&bar_empty[st], ((i / DWMM_NSTAGE) - 1) & 1);
                if (rec) dwmm_dbg[i * 8 + 1] = clock64();
                const long kb = kb0 + i;
                // image row of the K block: slab = kb / 12, row group = kb % 12; the lo part is 96 rows further
                const long slab = kb / DWMM_KB_PER_SLAB;
                const int row_hi = (int)(slab * 2 * TC_IMG_ROWS + (kb - slab * DWMM_KB_PER_SLAB) * 8);
                uint8_t* sbase = smem + st * DWMM_STAGE_BYTES;
                mbar_expect_tx(&bar_full[st], tx);
                tma_load_4d(sbase, &tm_g, 0, row_hi, grp0, 0, &bar_full[st]);
                if (nbox > 1) tma_load_4d(sbase + DWMM_BOX_BYTES, &tm_g, 0, row_hi, grp1, 0, &bar_full[st]);
                tma_load_4d(sbase + DWMM_B_OFF, &tm_d, 0, row_hi, 0, 0, &bar_full[st]);
                if (rec) dwmm_dbg[i * 8 + 2] = clock64();
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // =================================== MMA issuer ============================================================
        // one thread issues every MMA: everything it needs per K block lives in registers (descriptors of stage 0;
        // a stage is an address offset in the low word)
        if (lane == 0) {
            uint64_t dah[DWMM_MAXTILE], dal[DWMM_MAXTILE], dbh[DWMM_MAXTILE], dbl[DWMM_MAXTILE];
            uint32_t idesc[DWMM_MAXTILE], dcol[DWMM_MAXTILE];
            const uint32_t base = smem_u32(smem), bbase = base + DWMM_B_OFF;
#pragma unroll
            for (int j = 0; j < DWMM_MAXTILE; ++j) {
                const int og0 = (j < ntile) ? S.tile[j].og0 : 0;
                const uint32_t soff = (j < ntile) ? S.tile[j].soff : 0;
                dah[j] = make_smem_desc_mn(base + soff, 1024, 512);
                dal[j] = make_smem_desc_mn(base + soff + 2 * DWMM_TILE_BYTES, 1024, 512);
                dbh[j] = make_smem_desc_mn(bbase + (og0 / 8) * 1024, 1024, 512);
                dbl[j] = make_smem_desc_mn(bbase + DWMM_B_BYTES + (og0 / 8) * 1024, 1024, 512);
                idesc[j] = make_idesc_tf32_mn(128, j < ntile ? S.tile[j].ncol : 64);
                dcol[j] = taddr + (j < ntile ? S.tile[j].tcol : 0);
            }
            int since = 0, fl = 0, st = 0, ph = 0;
            for (int i = 0; i < nkb; ++i) {
                const bool rec = p.dbg && blockIdx.x == 0 && i < 256;
                if (rec) dwmm_dbg[i * 8 + 3] = clock64();
                mbar_wait(&bar_full[st], ph);
                if (rec) dwmm_dbg[i * 8 + 4] = clock64();
                if (since == 0 && fl > 0) mbar_wait(&bar_accempty, (fl - 1) & 1);   // the flush has read TMEM
                if (rec) dwmm_dbg[i * 8 + 5] = clock64();
                tc_fence_after();
                const uint32_t acc = since > 0 ? 1u : 0u;
                const uint64_t so = (uint64_t)((st * DWMM_STAGE_BYTES) >> 4);
                // narrow tiles first: the wide (long-running) MMAs are still queued while this thread waits for the
                // next stage, so the tensor pipe does not drain between K blocks
#pragma unroll
                for (int j = DWMM_MAXTILE - 1; j >= 0; --j)
                    if (j < ntile) {
                        umma_tf32(dcol[j], dal[j] + so, dbh[j] + so, idesc[j], acc);      // small terms first
                        umma_tf32(dcol[j], dah[j] + so, dbl[j] + so, idesc[j], 1u);
                        umma_tf32(dcol[j], dah[j] + so, dbh[j] + so, idesc[j], 1u);
                    }
                umma_commit(&bar_empty[st]);
                if (rec) dwmm_dbg[i * 8 + 6] = clock64();
                if (++since == DWMM_FLUSH || i == nkb - 1) { umma_commit(&bar_accfull); since = 0; ++fl; }
                if (++st == DWMM_NSTAGE) { st = 0; ph ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // =================================== flush warps ============================================================
        // TMEM lane = tile row, so thread = row; the partial is stored [column][row]: a warp access is one 128-byte line
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const int row = 32 * quad + lane;
        const int csplit = ((S.ncoltot / 64 + 1) / 2) * 64;             // columns are handed out in blocks of 64
        const int c_begin = half ? csplit : 0, c_end = half ? S.ncoltot : csplit;
        float* part = p.part + (size_t)blockIdx.x * 512 * 128 + row;
        for (int f = 0; f < nflush; ++f) {
            mbar_wait(&bar_accfull, f & 1);
            tc_fence_after();
            for (int cb = c_begin; cb < c_end; cb += 64) {
                float o[64];
                float* q = part + (size_t)cb * 128;
                if (f > 0) {                                          // 64 loads in flight per thread: the partial comes back from HBM
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
        if (nflush == 0)                                                // CTA without work: zero partial
            for (int cb = c_begin; cb < c_end; ++cb) part[(size_t)cb * 128] = 0.f;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(taddr);
}

cudaError_t dwmm_read_dbg(long long* out, int n) {
    return cudaMemcpyFromSymbol(out, dwmm_dbg, sizeof(long long) * (n < 2048 ? n : 2048));
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
// thread = (column quad, row slot): float4 loads, 4 rows in flight per thread, 4 row slots per CTA; fixed summation order
__global__ void __launch_bounds__(192) colsum_kernel(const float* dA, long rows, int H3, float* partial) {
    __shared__ float4 red[4][48];
    const int nq = H3 / 4;                                             // 48 column quads (H = 64)
    const int cq = threadIdx.x % nq, slot = threadIdx.x / nq;
    const long r0 = rows * blockIdx.x / gridDim.x, r1 = rows * (blockIdx.x + 1) / gridDim.x;
    const float4* src = reinterpret_cast<const float4*>(dA) + cq;
    float4 a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    long r = r0 + slot;
    for (; r + 12 < r1; r += 16) {
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __ldcs(src + (size_t)(r + 4 * i) * nq);
#pragma unroll
        for (int i = 0; i < 4; ++i) { a[i].x += v[i].x; a[i].y += v[i].y; a[i].z += v[i].z; a[i].w += v[i].w; }
    }
    for (; r < r1; r += 4) {
        const float4 v = __ldcs(src + (size_t)r * nq);
        a[0].x += v.x; a[0].y += v.y; a[0].z += v.z; a[0].w += v.w;
    }
    float4 s;
    s.x = (a[0].x + a[1].x) + (a[2].x + a[3].x); s.y = (a[0].y + a[1].y) + (a[2].y + a[3].y);
    s.z = (a[0].z + a[1].z) + (a[2].z + a[3].z); s.w = (a[0].w + a[1].w) + (a[2].w + a[3].w);
    red[slot][cq] = s;
    __syncthreads();
    if (slot == 0) {
        const float4 b = red[1][cq], c = red[2][cq], d = red[3][cq];
        s.x = (s.x + b.x) + (c.x + d.x); s.y = (s.y + b.y) + (c.y + d.y);
        s.z = (s.z + b.z) + (c.z + d.z); s.w = (s.w + b.w) + (c.w + d.w);
        reinterpret_cast<float4*>(partial + (size_t)blockIdx.x * H3)[cq] = s;
    }
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
    p.nkb = nslab * DWMM_KB_PER_SLAB;
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
    // pairs of adjacent tiles (one TMA box each), first-fit decreasing into sets of <= 512 TMEM columns and <= 2 pairs
    const int npair = (ntile + 1) / 2;
    int pcols[8];
    for (int k = 0; k < npair; ++k) pcols[k] = tiles[2 * k].ncol + (2 * k + 1 < ntile ? tiles[2 * k + 1].ncol : 0);
    bool used[8] = {false};
    for (int placed = 0; placed < npair;) {
        if (p.nset == DWMM_MAXSET) return false;
        DwmmSet& S = p.set[p.nset++];
        S.ntile = 0; S.ncoltot = 0; S.nbox = 0; S.boxgrp[0] = S.boxgrp[1] = 0;
        for (int want = 384; want >= 64; want -= 64)
            for (int k = 0; k < npair; ++k)
                if (!used[k] && pcols[k] == want && S.ncoltot + want <= 512 && S.nbox < 2) {
                    used[k] = true; ++placed;
                    for (int j = 2 * k; j < 2 * k + 2 && j < ntile; ++j) {
                        DwmmTile t = tiles[j];
                        t.tcol = S.ncoltot;
                        t.soff = S.nbox * (4 * 4096) + (j - 2 * k) * 4096;      // [box][hi|lo][tile of the pair][4 KB]
                        S.ncoltot += t.ncol;
                        S.tile[S.ntile++] = t;
                    }
                    S.boxgrp[S.nbox++] = 8 * k;
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

cudaError_t launch_dw_mm(const DwmmParams& p_, int fin, int H, int M, float* dWg, float* dWc, cudaStream_t st) {
    DwmmParams p = p_;
    { const char* e = getenv("DCGRU_DBG"); p.dbg = e ? (atoi(e) & 8) : 0; }
    // swizzling 4-D views of the row-major images: (32 floats = one 128-byte row piece | image row | 32-float group |
    // hi / lo part, 96 rows further); the row coordinate always addresses the hi part
    CUtensorMap tg, td;
    const unsigned long long rows = (unsigned long long)(p.nkb / DWMM_KB_PER_SLAB) * 2 * TC_IMG_ROWS;
    {
        const unsigned long long kkp = seq_fwd_tc_kkp(fin);
        const unsigned long long dims[4] = {32, rows, kkp / 32, 2};
        const unsigned long long str[4] = {4, kkp * 4, 128, TC_IMG_ROWS * kkp * 4};
        const unsigned box[4] = {32, 8, 8, 2};
        cudaError_t e = make_tmap_f32(&tg, p.G, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (e != cudaSuccess) return e;
    }
    {
        const unsigned long long dims[4] = {32, rows, (unsigned long long)(3 * H / 32), 2};
        const unsigned long long str[4] = {4, (unsigned long long)3 * H * 4, 128, (unsigned long long)TC_IMG_ROWS * 3 * H * 4};
        const unsigned box[4] = {32, 8, 6, 2};
        cudaError_t e = make_tmap_f32(&td, p.DA, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = cudaFuncSetAttribute(dw_mm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DWMM_SMEM);
    if (e != cudaSuccess) return e;
    dw_mm_kernel<<<p.ncta, DWMM_THREADS, DWMM_SMEM, st>>>(p, tg, td);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int n = (fin + H) * M * 3 * H;
    dwmm_reduce_kernel<<<(n + 255) / 256, 256, 0, st>>>(p, fin, H, M, dWg, dWc);
    return cudaGetLastError();
}

cudaError_t launch_colsum(const float* dA, long rows, int H, float* partial, float* dbg, float* dbc, cudaStream_t st) {
    if (H != 64) return cudaErrorInvalidValue;                         // 48 column quads x 4 row slots = 192 threads
    colsum_kernel<<<CS_CTAS, 192, 0, st>>>(dA, rows, 3 * H, partial);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    colsum_final_kernel<<<1, 3 * H, 0, st>>>(partial, CS_CTAS, H, dbg, dbc);
    return cudaGetLastError();
}

int dw_mm_smem_bytes() { return DWMM_SMEM; }

}  // namespace dcgru
