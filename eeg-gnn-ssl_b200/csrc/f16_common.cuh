// Building blocks of the second-generation tensor-core kernels (bulk_dp.cu, rnn_fwd.cu, rnn_bwd.cu, dw_mm16.cu).
//
// Arithmetic: "2xFP16" -- every fp32 operand x is split into hi = fp16(x) (round to nearest) and lo = fp16(x - hi);
// hi + lo carries 22 significand bits (fp16 has 11), and a product a*b is issued as three kind::f16 MMAs
//     a_lo*b_hi + a_hi*b_lo + a_hi*b_hi
// whose fp16 x fp16 products are exact in the fp32 accumulator (TMEM).  The dropped a_lo*b_lo term is 2^-22 relative:
// the same accuracy class as the 3xTF32 scheme of the first-generation kernels (tc_common.cuh), at twice the tensor
// rate (kind::f16 has K = 16 per instruction, kind::tf32 K = 8) and half the operand bytes in shared memory and HBM.
// Range: fp16 holds |x| < 65504 and resolves down to 6e-8.  Forward operands (standardised inputs, gates in (0,1),
// tanh states, Xavier weights, diffusion polynomials) sit far inside; the backward kernels multiply the gradient
// by a power of two chosen from the upstream gradient's magnitude before the split and undo it after the
// accumulator is read (exact), see rnn_bwd.cu.
//
// Operand tiles ("chunks"): 64 K-values (128 bytes) per row, canonical K-major SWIZZLE_128B layout
//     byte(row r, k) = (r / 8) * 1024 + (r % 8) * 128 + (((k / 8) ^ (r % 8)) * 16) + (k % 8) * 2
// i.e. exactly what a TMA box with a 128-byte inner extent and CU_TENSOR_MAP_SWIZZLE_128B reads / writes, so the
// same bytes serve as UMMA operand and as source of the operand-image dumps.  One MMA k-step (16 values) is a
// 32-byte advance of the descriptor's start address.  A chunk slot = hi plane (16 KB for 128 rows) + lo plane.
//
// Diffusion: P_m (19 x 19 per sample) is applied by the worker warps with fp32 FMAs, one warp per (sample, chunk):
// lane = two adjacent source columns, registers = the 19 (padded 20) output rows; the rows of P^T are broadcast
// reads from shared memory.  Every lane is busy whatever the node count (the first generation mapped lane = row
// and idled 13 of 32 lanes), and the whole phase is one barrier round instead of one per 8-column chunk.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc_common.cuh"

namespace dcgru {
namespace f16 {
using namespace tc;

constexpr int SB = 4;                 // samples per 128-row tile
constexpr int RP = 32;                // rows per sample (nodes 0..N-1, rest zero)
constexpr int RG = 3;                 // 8-row groups of a sample that can be non-zero (N <= 24)
constexpr int IMG_ROWS = SB * RG * 8; // 96 image rows per (tile, t, plane)
constexpr int NPAD = 20;              // node count the FMA loops are unrolled for
constexpr int PLANE = 16 * 1024;      // one plane (hi or lo) of a 128-row chunk
constexpr int SLOT = 2 * PLANE;       // chunk slot: hi | lo
constexpr int PT_STRIDE = NPAD * NPAD;   // floats per (sample, term) block of transposed polynomials

// kind::f16 (A, B = fp16), fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// both operands MN-major (dw_mm16: the GEMM's K index is the image row)
__host__ __device__ constexpr uint32_t make_idesc_f16_mn(int M, int N) {
    return make_idesc_f16(M, N) | (1u << 15) | (1u << 16);
}
// K-major SWIZZLE_128B operand: 8-row atoms of 1024 bytes, next atom SBO = 1024 further (dense)
__device__ __forceinline__ uint64_t make_desc_k128(uint32_t saddr) {
    return make_smem_desc(saddr, 16, 1024) | ((uint64_t)2 << 61);
}
// MN-major SWIZZLE_128B operand: rows of 64 mn values (128 B), 8 k-rows = one 1024-byte atom; next 8 k-rows SBO
// further, next 64 mn values LBO further
__device__ __forceinline__ uint64_t make_desc_mn128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return make_smem_desc(saddr, lbo_bytes, sbo_bytes) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// byte offset of (row, k) inside one plane of a chunk (k in [0, 64))
__device__ __forceinline__ uint32_t k128_off(int row, int k) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) ^ (row & 7)) << 4) | ((k & 7) << 1)));
}

// 32 lanes x 8 consecutive columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
}

// the same without the wait: issue several loads back to back, then tmem_wait_ld() once (the registers are valid after it)
__device__ __forceinline__ void tmem_ld8_nw(uint32_t taddr, float (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "r"(taddr) : "memory");
}
// TMEM loads are paid per INSTRUCTION (measured: an epilogue made of x8 loads spent ~1k cycles per 8-column chunk with
// nothing else in it; eight warps share the load path), so the epilogues use the widest shapes the register budget allows
__device__ __forceinline__ void tmem_ld16_nw(uint32_t taddr, float (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]),
          "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld32_nw(uint32_t taddr, float (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]),
          "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]),
          "=f"(v[16]), "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]),
          "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
// registers -> TMEM, 32 lanes x 8 consecutive columns, no wait (tmem_wait_st() before the data is consumed elsewhere)
__device__ __forceinline__ void tmem_st8_nw(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n"
                 ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// ---- fp32 -> (hi, lo) fp16 pairs ----------------------------------------------------------------------------------
__device__ __forceinline__ float clamp_h(float x) { return fminf(fmaxf(x, -65000.f), 65000.f); }
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    a = clamp_h(a); b = clamp_h(b);
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// host + device scalar version (weight packing)
__device__ __forceinline__ void split1(float a, __half& hi, __half& lo) {
    a = clamp_h(a);
    hi = __float2half_rn(a);
    lo = __float2half_rn(a - __half2float(hi));
}
// 8 consecutive k of one row (one 16-byte swizzle unit) from 8 floats
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
    split2(v[0], v[1], hi.x, lo.x); split2(v[2], v[3], hi.y, lo.y);
    split2(v[4], v[5], hi.z, lo.z); split2(v[6], v[7], hi.w, lo.w);
}

// ---- diffusion of two columns by one lane -------------------------------------------------------------------------
// acc[n][0..1] = sum_{j < N} PT[j][n] * z[j][0..1];  z points at (node row 0, this lane's first column) of an fp32
// shared-memory tile with row stride zld floats; PT = this lane's [j][NPAD] block (rows of P^T, zero beyond N).
// The inner product runs on packed fp32 pairs (fma.rn.f32x2 -> FFMA2, sm_100): a register pair holds the outputs of
// two adjacent nodes (n, n+1) for one column, the P^T pair comes straight out of the 16-byte shared-memory read and the
// source value is the instruction's broadcast scalar operand -- 20 FFMA2 instead of 40 FFMA per source row.
template <class ZF>
__device__ __forceinline__ void diffuse2f(ZF zf, int N, const float* PT, float (&acc)[NPAD][2]) {
    unsigned long long a64[NPAD / 2][2];
#pragma unroll
    for (int i = 0; i < NPAD / 2; ++i) { a64[i][0] = 0ull; a64[i][1] = 0ull; }
#pragma unroll 2
    for (int j = 0; j < N; ++j) {
        const float2 zz = zf(j);
        unsigned long long zx, zy;
        asm("mov.b64 %0, {%1, %1};" : "=l"(zx) : "f"(zz.x));
        asm("mov.b64 %0, {%1, %1};" : "=l"(zy) : "f"(zz.y));
        const ulonglong2* pr = reinterpret_cast<const ulonglong2*>(PT + j * NPAD);
#pragma unroll
        for (int q = 0; q < NPAD / 4; ++q) {
            const ulonglong2 pv = pr[q];
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a64[2 * q][0]) : "l"(pv.x), "l"(zx));
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a64[2 * q][1]) : "l"(pv.x), "l"(zy));
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a64[2 * q + 1][0]) : "l"(pv.y), "l"(zx));
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a64[2 * q + 1][1]) : "l"(pv.y), "l"(zy));
        }
    }
#pragma unroll
    for (int i = 0; i < NPAD / 2; ++i) {
        asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[2 * i][0]), "=f"(acc[2 * i + 1][0]) : "l"(a64[i][0]));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[2 * i][1]), "=f"(acc[2 * i + 1][1]) : "l"(a64[i][1]));
    }
}
__device__ __forceinline__ void diffuse2(const float* z, int zld, int N, const float* PT, float (&acc)[NPAD][2]) {
    diffuse2f([&](int j) { return *reinterpret_cast<const float2*>(z + j * zld); }, N, PT, acc);
}
// same from an fp16 hi / lo source (rows of an operand image): z = hi + lo is the fp32 value to 2^-22
__device__ __forceinline__ float2 ld_hilo2(const __half* hi, const __half* lo) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(hi)), b = __half22float2(*reinterpret_cast<const __half2*>(lo));
    return make_float2(a.x + b.x, a.y + b.y);
}
// one column per lane (32-column half tasks: every warp works on the same term, so a term's chunk is complete -- and its
// MMAs can start -- while the next term is still being diffused)
__device__ __forceinline__ void diffuse1(const float* z, int zld, int N, const float* PT, float (&acc)[NPAD]) {
    unsigned long long a64[NPAD / 2];
#pragma unroll
    for (int i = 0; i < NPAD / 2; ++i) a64[i] = 0ull;
#pragma unroll 2
    for (int j = 0; j < N; ++j) {
        const float zv = z[j * zld];
        unsigned long long zx;
        asm("mov.b64 %0, {%1, %1};" : "=l"(zx) : "f"(zv));
        const ulonglong2* pr = reinterpret_cast<const ulonglong2*>(PT + j * NPAD);
#pragma unroll
        for (int q = 0; q < NPAD / 4; ++q) {
            const ulonglong2 pv = pr[q];
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a64[2 * q]) : "l"(pv.x), "l"(zx));
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a64[2 * q + 1]) : "l"(pv.y), "l"(zx));
        }
    }
#pragma unroll
    for (int i = 0; i < NPAD / 2; ++i)
        asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[2 * i]), "=f"(acc[2 * i + 1]) : "l"(a64[i]));
}
// rows n < N of acc (one column k per lane) -> chunk slot; rows N..nzero-1 are written as zeros
__device__ __forceinline__ void store_col1(uint8_t* slot, int row0, int k, int N, int nzero, const float (&acc)[NPAD], float scale) {
#pragma unroll
    for (int n = 0; n < RG * 8; ++n) {
        if (n < nzero) {
            const float v = (n < NPAD && n < N) ? clamp_h(acc[n < NPAD ? n : 0] * scale) : 0.f;
            const __half h = __float2half_rn(v);
            const __half l = __float2half_rn(v - __half2float(h));
            const uint32_t off = k128_off(row0 + n, k);
            *reinterpret_cast<__half*>(slot + off) = h;
            *reinterpret_cast<__half*>(slot + PLANE + off) = l;
        }
    }
}
// rows n < N of acc -> chunk slot (hi plane at `slot`, lo plane PLANE further), tile row = row0 + n, k = 2 * lane;
// rows N..nzero-1 are written as zeros (nzero = 0: leave them alone)
__device__ __forceinline__ void store_cols2(uint8_t* slot, int row0, int lane, int N, const float (&acc)[NPAD][2], float scale,
                                            int nzero = 0) {
#pragma unroll
    for (int n = 0; n < RG * 8; ++n) {
        if (n < N || n < nzero) {
            uint32_t hi = 0u, lo = 0u;
            if (n < NPAD && n < N) split2(acc[n < NPAD ? n : 0][0] * scale, acc[n < NPAD ? n : 0][1] * scale, hi, lo);
            const uint32_t off = k128_off(row0 + n, 2 * lane);
            *reinterpret_cast<uint32_t*>(slot + off) = hi;
            *reinterpret_cast<uint32_t*>(slot + PLANE + off) = lo;
        }
    }
}

// ---- diffusion of a 64-column tile on the warp-level tensor path (mma.sync m16n8k16, 2xFP16) ---------------------------
// out[n][c] = sum_j PT[j][n] * z[j][c] for one (sample, term): A = P (rows n, padded to 32; K = j, padded to 32) split hi/lo
// from the fp32 block PT, B = z split hi/lo from an fp32 shared-memory tile (row stride zld floats, zld % 32 == 4 keeps
// the fragment loads conflict-free), three MMAs per product (a_lo*b_hi + a_hi*b_lo + a_hi*b_hi) with fp32 accumulation --
// the arithmetic of the tcgen05 kernels.  The FMA version (diffuse2 above) is bound by shared-memory RETURN bandwidth:
// every broadcast LDS.128 of a P^T row delivers 512 B to the warp, 22 LSU data cycles per source row against 10 of
// FFMA2 issue (measured ~56 of 128 FMA/clk/SM, profiles/fma_rate_r02.txt); here the polynomial lives in registers for the
// whole task and shared memory is read once per source element.
// The accumulator fragment holds adjacent columns (c, c+1) of a row, so results go to the chunk slot as 4-byte hi / lo
// stores exactly like store_cols2: rows n < 24 of the sample's 32-row block are written (rows N..23 come out as exact
// zeros because P has no such rows), rows 24..31 are left alone.
__device__ __forceinline__ void mma_f16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
struct PFrag { uint32_t hi[2][2][4], lo[2][2][4]; };       // [m tile][k tile][a0..a3]
// PT: this (sample, term)'s [j][NPAD] block (zero beyond N)
__device__ __forceinline__ void load_pfrag(const float* PT, int lane, PFrag& f) {
    const int g = lane >> 2, t = lane & 3;
    auto P = [&](int n, int j) { return (n < NPAD && j < NPAD) ? PT[j * NPAD + n] : 0.f; };
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
            const int r0 = 16 * mt + g, r1 = r0 + 8, j0 = 16 * kt + 2 * t;
            split2(P(r0, j0), P(r0, j0 + 1), f.hi[mt][kt][0], f.lo[mt][kt][0]);
            split2(P(r0, j0 + 8), P(r0, j0 + 9), f.hi[mt][kt][2], f.lo[mt][kt][2]);
            if (mt == 0) {
                split2(P(r1, j0), P(r1, j0 + 1), f.hi[mt][kt][1], f.lo[mt][kt][1]);
                split2(P(r1, j0 + 8), P(r1, j0 + 9), f.hi[mt][kt][3], f.lo[mt][kt][3]);
            } else {                                                    // rows 24..31: no such nodes
                f.hi[mt][kt][1] = 0u; f.lo[mt][kt][1] = 0u; f.hi[mt][kt][3] = 0u; f.lo[mt][kt][3] = 0u;
            }
        }
}
// z: (source row 0, column 0) of the sample's fp32 tile; slot: chunk slot (hi plane, lo plane PLANE further); row0: the
// sample's first tile row; the source values are multiplied by `scale` (a power of two) before their fp16 split
__device__ __forceinline__ void diffuse_mma(const float* z, int zld, int N, const PFrag& pf, uint8_t* slot, int row0, int lane,
                                            float scale) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll 1
    for (int grp = 0; grp < 4; ++grp) {
        uint32_t bh[2][2][2], bl[2][2][2];                 // [n tile][k tile][b0, b1]
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const float* zc = z + 16 * grp + 8 * nt + g;
#pragma unroll
            for (int kt = 0; kt < 2; ++kt) {
                const int j0 = 16 * kt + 2 * t;
                const float v0 = j0 < N ? zc[j0 * zld] : 0.f, v1 = j0 + 1 < N ? zc[(j0 + 1) * zld] : 0.f;
                const float v2 = j0 + 8 < N ? zc[(j0 + 8) * zld] : 0.f, v3 = j0 + 9 < N ? zc[(j0 + 9) * zld] : 0.f;
                split2(v0 * scale, v1 * scale, bh[nt][kt][0], bl[nt][kt][0]);      // scale BEFORE the fp16 split: small gradients
                split2(v2 * scale, v3 * scale, bh[nt][kt][1], bl[nt][kt][1]);      // would land in the fp16 subnormals otherwise
            }
        }
        // four independent accumulators (m tile x n tile) per dependent step: a dependent HMMA chain costs ~33 cycles per link
        float d[2][2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) { d[mt][nt][0] = 0.f; d[mt][nt][1] = 0.f; d[mt][nt][2] = 0.f; d[mt][nt][3] = 0.f; }
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) mma_f16_16816(d[mt][nt], pf.lo[mt][kt], bh[nt][kt]);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) mma_f16_16816(d[mt][nt], pf.hi[mt][kt], bl[nt][kt]);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) mma_f16_16816(d[mt][nt], pf.hi[mt][kt], bh[nt][kt]);
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const int col = 16 * grp + 8 * nt + 2 * t;
                uint32_t hi, lo;
                split2(d[mt][nt][0], d[mt][nt][1], hi, lo);
                uint32_t off = k128_off(row0 + 16 * mt + g, col);
                *reinterpret_cast<uint32_t*>(slot + off) = hi;
                *reinterpret_cast<uint32_t*>(slot + PLANE + off) = lo;
                if (mt == 0) {                             // rows 24..31 of the block do not exist
                    split2(d[mt][nt][2], d[mt][nt][3], hi, lo);
                    off = k128_off(row0 + 8 + g, col);
                    *reinterpret_cast<uint32_t*>(slot + off) = hi;
                    *reinterpret_cast<uint32_t*>(slot + PLANE + off) = lo;
                }
            }
    }
}

// The same with the source already in fp16 hi / lo form: a K-major SWIZZLE_128B chunk (rows = tile rows, 64 columns), e.g.
// the term-0 chunk the epilogue wrote for the tensor core.  ldmatrix.trans hands every lane exactly its B fragment
// (source rows 2t, 2t+1 of column g), one instruction per (n tile, plane) for both k tiles; the 16-byte units of
// eight consecutive rows sit in eight different bank groups (that is what the swizzle is for).  Rows >= N of the
// source block must be finite (they meet zero polynomial entries); the callers keep them zero.
__device__ __forceinline__ void mma_f16_1688(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr) : "memory");
}
// Core: `addr(grp, nt)` = shared-memory address of this lane's source row (row index = lane), 16-byte unit holding columns
// [16 grp + 8 nt, +8) of the hi plane; the lo plane lies lo_off bytes further.  Column groups [grp0, grp1) of 16 columns
// each are processed (a whole 64-column chunk = [0, 4)).
template <typename AddrF>
__device__ __forceinline__ void diffuse_mma16_core(AddrF addr, uint32_t lo_off, const PFrag& pf, uint8_t* slot, int row0, int lane,
                                                   float scale, int grp0, int grp1) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll 1
    for (int grp = grp0; grp < grp1; ++grp) {
        uint32_t bh[2][4], bl[2][4];                                    // [n tile][kt0.b0, kt0.b1, kt1.b0, kt1.b1]
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const uint32_t a = addr(grp, nt);
            ldmatrix_x4_trans(bh[nt], a);
            ldmatrix_x4_trans(bl[nt], a + lo_off);
        }
        float d[2][2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) { d[mt][nt][0] = 0.f; d[mt][nt][1] = 0.f; d[mt][nt][2] = 0.f; d[mt][nt][3] = 0.f; }
        // source rows 0..15: k16; rows 16..23: k8 (N <= 24; the fragments of rows 24..31 are not used)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const uint32_t b[2] = {bh[nt][0], bh[nt][1]};
                mma_f16_16816(d[mt][nt], pf.lo[mt][0], b);
            }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const uint32_t b[2] = {bl[nt][0], bl[nt][1]};
                mma_f16_16816(d[mt][nt], pf.hi[mt][0], b);
            }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const uint32_t b[2] = {bh[nt][0], bh[nt][1]};
                mma_f16_16816(d[mt][nt], pf.hi[mt][0], b);
            }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) mma_f16_1688(d[mt][nt], pf.lo[mt][1][0], pf.lo[mt][1][1], bh[nt][2]);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) mma_f16_1688(d[mt][nt], pf.hi[mt][1][0], pf.hi[mt][1][1], bl[nt][2]);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) mma_f16_1688(d[mt][nt], pf.hi[mt][1][0], pf.hi[mt][1][1], bh[nt][2]);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const int col = 16 * grp + 8 * nt + 2 * t;
                uint32_t hi, lo;
                split2(d[mt][nt][0] * scale, d[mt][nt][1] * scale, hi, lo);
                uint32_t off = k128_off(row0 + 16 * mt + g, col);
                *reinterpret_cast<uint32_t*>(slot + off) = hi;
                *reinterpret_cast<uint32_t*>(slot + PLANE + off) = lo;
                if (mt == 0) {
                    split2(d[mt][nt][2] * scale, d[mt][nt][3] * scale, hi, lo);
                    off = k128_off(row0 + 8 + g, col);
                    *reinterpret_cast<uint32_t*>(slot + off) = hi;
                    *reinterpret_cast<uint32_t*>(slot + PLANE + off) = lo;
                }
            }
    }
}

// source = a K-major SWIZZLE_128B chunk slot (hi plane, lo plane PLANE further); srow0 = the sample's first tile row
__device__ __forceinline__ void diffuse_mma16(const uint8_t* src, int srow0, const PFrag& pf, uint8_t* slot, int row0, int lane,
                                              float scale, int grp0 = 0, int grp1 = 4) {
    const int sr = srow0 + lane;                                        // lane i supplies the address of source row i
    const uint32_t rbase = smem_u32(src) + (uint32_t)((sr >> 3) * 1024 + (sr & 7) * 128);
    diffuse_mma16_core([&](int grp, int nt) { return rbase + (uint32_t)((((2 * grp + nt) ^ (sr & 7)) << 4)); }, (uint32_t)PLANE, pf,
                       slot, row0, lane, scale, grp0, grp1);
}
// source = row-major fp16 planes (an operand-image slab copied to shared memory as it lies in HBM): hi points at (the sample's
// first row, first column of the chunk), rows ld_bytes apart, the lo plane lo_off bytes further.  Rows 24..31 of the lane
// addresses only have to be readable.  (Eight rows at a 128-byte-multiple stride share their banks: the loads are 8-way
// conflicted, ~100 cycles per group against ~800 of MMA work.)
__device__ __forceinline__ void diffuse_mma16_rm(const __half* hi, uint32_t ld_bytes, uint32_t lo_off, const PFrag& pf, uint8_t* slot,
                                                 int row0, int lane, float scale) {
    const uint32_t rbase = smem_u32(hi) + (uint32_t)lane * ld_bytes;
    diffuse_mma16_core([&](int grp, int nt) { return rbase + (uint32_t)((2 * grp + nt) << 4); }, lo_off, pf, slot, row0, lane, scale, 0, 4);
}

// transposed polynomials of one tile into shared memory: PT[(s * nterm + m) * PT_STRIDE + j * NPAD + n]
//   forward  (transpose = 0): PT[j][n] = P[b][m][n][j]   (out[n] = sum_j P[n][j] z[j])
//   backward (transpose = 1): PT[j][n] = P[b][m][j][n]   (out[n] = sum_j P[j][n] z[j] = (P^T z)[n])
// Entries beyond N stay zero (the caller zeroes the block once).  Called by `nthreads` threads with index `tid`.
__device__ __forceinline__ void load_pt(float* PTs, const float* P, int b0, int B, int N, int nterm, int transpose,
                                        int tid, int nthreads) {
    const int per = nterm * N * N;
    for (int idx = tid; idx < SB * per; idx += nthreads) {
        const int s = idx / per, r = idx - s * per;
        const int m = r / (N * N), e = r - m * N * N;
        const int a = e / N, c = e - a * N;                 // P[b][m][a][c]
        const int b = b0 + s;
        float v = 0.f;
        if (b < B) v = P[((size_t)b * nterm + m) * N * N + e];
        const int j = transpose ? a : c, n = transpose ? c : a;
        PTs[(s * nterm + m) * PT_STRIDE + j * NPAD + n] = v;
    }
}

// global (2-D box) <- shared, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void tma_store_2d(const void* tmap, int c0, int c1, const void* smem_src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];\n"
                 ::"l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(smem_src)) : "memory");
}

// Gate activations on the MUFU pipe with the flush-to-zero forms written out in PTX: __expf / __fdividef without
// -ftz expand into denormal fix-up code (FSETP / FSEL / extra FMULs: ~44 instructions per (tanh, sigmoid) pair in the
// first version of the epilogue, which made it the longest phase of a step).  ex2.approx / rcp.approx are accurate
// to ~2^-22 relative; the clamps keep every intermediate finite (sigmoid(-30) = 9e-14, tanh(15) = 1 - 2e-13).
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
constexpr float LOG2E = 1.4426950408889634f;
__device__ __forceinline__ float fast_sigmoid(float x) {
    return rcp_ftz(1.f + ex2_ftz(-LOG2E * fminf(fmaxf(x, -30.f), 30.f)));
}
__device__ __forceinline__ float fast_tanh(float x) {
    return fmaf(-2.f, rcp_ftz(1.f + ex2_ftz((2.f * LOG2E) * fminf(fmaxf(x, -15.f), 15.f))), 1.f);
}
// c = tanh(a), u = sigmoid(b) with ONE reciprocal: 1/A = B/(AB), 1/B = A/(AB)  (A = 1 + e^{2a}, B = 1 + e^{-b}; AB < 1e27)
__device__ __forceinline__ void tanh_sigmoid(float a, float b, float& c, float& u) {
    const float A = 1.f + ex2_ftz((2.f * LOG2E) * fminf(fmaxf(a, -15.f), 15.f));
    const float Bq = 1.f + ex2_ftz(-LOG2E * fminf(fmaxf(b, -30.f), 30.f));
    const float rinv = rcp_ftz(A * Bq);
    c = fmaf(-2.f * Bq, rinv, 1.f);
    u = A * rinv;
}

}  // namespace f16
}  // namespace dcgru
