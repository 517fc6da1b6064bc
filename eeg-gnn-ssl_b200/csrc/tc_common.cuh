// tcgen05 / TMEM / mbarrier building blocks (sm_100a) shared by the tensor-core kernels.
//
// Arithmetic: "3xTF32" -- every fp32 operand x is split into hi = x with the low 13 mantissa bits
// cleared (exactly representable in TF32) and lo = x - hi (exact in fp32); a product a*b is issued as
// three kind::tf32 MMAs  a_hi*b_hi + a_hi*b_lo + a_lo*b_hi  accumulating in fp32 in TMEM.  The dropped
// a_lo*b_lo term and the TF32 rounding of the lo parts are ~2^-21 relative, i.e. fp32-level accuracy,
// which is what the 1e-4 parity bar against the fp32 reference needs over 60 recurrent steps.
//
// Operand tiles in shared memory use the canonical K-major no-swizzle UMMA layout with the K-groups
// outermost:   element (row r, k)  ->  tile[(k/4)][r] (a float4 holding k%4 = 0..3)
//   core matrix = 8 rows x 16 B contiguous,  SBO (next 8 rows) = 128 B,  LBO (next 4 k) = rows*16 B
// so a thread that owns 4 consecutive k of one row writes one 16-byte word and consecutive rows are
// consecutive words (conflict-free), and a sub-range of rows is just an address offset.
#pragma once
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dcgru {
namespace tc {

// Row layout of the tensor-core sequence kernels: a CTA owns TC_SB = 4 samples, sample s sits in rows
// [32 s, 32 s + 32) of the 128-row UMMA tile (nodes 0..N-1, the rest zero).  One warp = one sample, so every
// shared-memory read of the diffusion (all lanes read the same node row of the same sample) is a single
// broadcast wavefront -- with the denser 20-row packing (6 samples) a warp straddled samples and each such
// read cost 4 wavefronts, which made the kernels shared-memory bound.  B = 512 needs 128 CTAs: still one wave.
// Only the first TC_RG = 3 row groups (24 rows) of a sample are ever non-zero; the operand images keep just those.
constexpr int TC_SB = 4;
constexpr int TC_RP = 32;
constexpr int TC_RG = 3;
constexpr int TC_IMG_ROWS = TC_SB * TC_RG * 8;      // 96 image rows per (cta, t, hi|lo)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- descriptors (cute/arch/mma_sm100_desc.hpp bit layout) -----------------------------------------
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                 // descriptor version 1 (Blackwell); layout_type 0 = no swizzle
    return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M x N
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// same, both operands MN-major (the M / N index is the contiguous one in shared memory): bits 15 and 16
__host__ __device__ constexpr uint32_t make_idesc_tf32_mn(int M, int N) {
    return make_idesc_tf32(M, N) | (1u << 15) | (1u << 16);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
                 ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM ---------------------------------------------------------------------------------------------
// one full warp; ncols power of two >= 32; the base address is written to *slot (shared memory)
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                 ::"r"(smem_u32(slot)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// 32 lanes x 32 consecutive columns: thread i of the warp gets lane (base_lane + i), v[j] = column j
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

// registers -> TMEM, 32 lanes x 32 consecutive columns (the inverse of tmem_ld32)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
        ::"r"(taddr),
          "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
          "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
          "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
          "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
          "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}

// 32 lanes x 16 consecutive columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}

// ---- mbarrier -----------------------------------------------------------------------------------------
// Stress build (-DDCGRU_JITTER=<ns>, libdcgru_b200_jitter.so, tests/test_gpu_stress.py): every mbarrier wait / arrive /
// commit is preceded, one time in four, by a pseudo-random sleep of up to DCGRU_JITTER ns, so producers, issuer, loaders
// and dump warps drift against each other by whole pipeline stages.  A wait that polls a phase a later phase can
// overtake, or a slot reused before its consumer is done, then shows up within a few hundred iterations as a trap
// (bounded spins) or as a result that differs from the un-jittered run.  The production build compiles this to nothing.
#ifdef DCGRU_JITTER
__device__ __forceinline__ void dbg_jitter() {
    const unsigned long long c = clock64();
    unsigned h = ((unsigned)c ^ (unsigned)(c >> 17)) * 2654435761u + threadIdx.x * 40503u + blockIdx.x * 9973u;
    h ^= h >> 15;
    if ((h & 3u) == 0u) __nanosleep((h >> 8) % (unsigned)(DCGRU_JITTER));
}
#else
__device__ __forceinline__ void dbg_jitter() {}
#endif
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
// bounded wait: a lost arrival traps instead of hanging the GPU.  try_wait carries a suspend-time hint (ns): the
// hardware parks the warp until the phase completes instead of returning after a short internal time-out.  Without
// the hint a dozen waiting warps re-issued try_wait back to back -- mbarrier instructions execute on the XU pipe (16
// lanes/clk/SM, shared with MUFU and the fp16 conversions): ncu showed it 100 % busy over the whole rnn_fwd kernel and
// 36 % of all executed warp instructions were SYNCS/BRA of these loops, which starved the epilogues.
constexpr uint32_t MBAR_SUSPEND_NS = 1000000u;
#ifdef DCGRU_JITTER
constexpr uint32_t MBAR_SPINS = 1u << 12, MBAR_SPINS_RELAXED = 1u << 21;     // stress build: give up after a few seconds
#else
constexpr uint32_t MBAR_SPINS = 1u << 22, MBAR_SPINS_RELAXED = 1u << 24;
#endif
// a wait that never completes kills the context; the stress build first names the wait (source line, CTA, thread)
#ifdef DCGRU_JITTER
// progress marks of the stress build: one int per (CTA < 64, warp < 16), printed by the first wait that times out
__device__ int g_dbg_prog[64 * 16];
#define DBG_PROG(v) do { if ((threadIdx.x & 31) == 0 && blockIdx.x < 64) ((volatile int*)g_dbg_prog)[blockIdx.x * 16 + (threadIdx.x >> 5)] = (v); } while (0)
static __device__ __noinline__ void mbar_timeout(int line) {
    const unsigned act = __activemask();
    if ((int)(threadIdx.x & 31) == __ffs(act) - 1) {
        const volatile int* g = (const volatile int*)g_dbg_prog + (blockIdx.x < 64 ? blockIdx.x * 16 : 0);
        printf("dcgru_b200: mbarrier wait timed out at line %d (block %d, warp %d of %d, lanes %08x) progress: %d %d %d %d %d %d %d %d | %d %d %d | %d %d %d %d\n",
               line, (int)blockIdx.x, (int)(threadIdx.x >> 5), (int)(blockDim.x >> 5), act, g[0], g[1], g[2], g[3], g[4], g[5], g[6], g[7],
               g[8], g[9], g[10], g[11], g[12], g[13], g[14]);
    }
    for (int i = 0; i < 3000; ++i) __nanosleep(1000000);      // let every other stuck warp report before the context dies
    asm volatile("trap;\n");
}
#else
#define DBG_PROG(v) do { } while (0)
__device__ __forceinline__ void mbar_timeout(int) { asm volatile("trap;\n"); }
#endif
#define mbar_wait(...) mbar_wait_l(__LINE__, __VA_ARGS__)
#define mbar_wait_relaxed(...) mbar_wait_relaxed_l(__LINE__, __VA_ARGS__)
#define mbar_wait2(...) mbar_wait2_l(__LINE__, __VA_ARGS__)
#define mbar_wait3(...) mbar_wait3_l(__LINE__, __VA_ARGS__)
__device__ __forceinline__ void mbar_wait_l(int line, uint64_t* bar, uint32_t parity) {
    dbg_jitter();
    uint32_t done = 0;
    for (uint32_t it = 0; it < MBAR_SPINS; ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity), "r"(MBAR_SUSPEND_NS) : "memory");
        if (done) return;
    }
    mbar_timeout(line);
}

// for warps that are far off the critical path (ring loaders that run steps ahead, the image dump): poll with plain
// test_wait and sleep in between, so that the waiting does not compete with the working warps for issue slots
__device__ __forceinline__ void mbar_wait_relaxed_l(int line, uint64_t* bar, uint32_t parity, unsigned ns = 256) {
    dbg_jitter();
    uint32_t done = 0;
    for (uint32_t it = 0; it < MBAR_SPINS_RELAXED; ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return;
        __nanosleep(ns);
    }
    mbar_timeout(line);
}

// ---- bulk copies (TMA, no tensor map) ----------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    dbg_jitter();
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint64_t* bar, uint32_t n) {
    dbg_jitter();
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    dbg_jitter();
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n"
                 ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
// the sources of all committed groups have been read (shared memory may be overwritten)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
// all committed groups are complete (the global writes are performed)
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }

// ---- tensor-map TMA (cuTensorMapEncodeTiled descriptors passed as __grid_constant__ kernel parameters) --------
// global (3-D box) -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}
// shared -> global (4-D box), tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void tma_store_4d(const void* tmap, int c0, int c1, int c2, int c3, const void* smem_src) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];\n"
                 ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(smem_src)) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(tmap) : "memory");
}
// K-major operand with the 32-byte swizzle (descriptor layout type 6): rows of 8 k (32 B), 8 rows = one 256-byte
// atom, next 8 rows SBO further; the two 16-byte halves of a row are swapped in rows 4-7 of every atom (address
// bit 4 ^= bit 7), which makes thread-per-row 16-byte stores conflict free and lets a TMA store with
// CU_TENSOR_MAP_SWIZZLE_32B move whole 32-byte row pieces.  The tile base must be 256-byte aligned.
__device__ __forceinline__ uint64_t make_smem_desc_k32(uint32_t saddr, uint32_t sbo_bytes) {
    return make_smem_desc(saddr, 16, sbo_bytes) | ((uint64_t)6 << 61);
}
// float4 index of (row, K group kg) inside a tile of NKP K-group pairs laid out [row group][pair][8 rows][32 B]
template <int NKP>
__device__ __forceinline__ int k32_idx(int kg, int row) {
    return (row >> 3) * (NKP * 16) + (kg >> 1) * 16 + (row & 7) * 2 + ((kg & 1) ^ ((row >> 2) & 1));
}

// MN-major fp32/tf32 operand: the only layout the tensor core accepts is "128-byte swizzle with 32-byte atoms"
// (descriptor layout type 1; pinned with dcgru_tc_probe on a B200): element (mn, k) of a 128 x 8 tile lives at
//   (mn/32)*LBO + (k/4)*SBO + (k%4)*128 + (((mn%32)/8) ^ (k%4))*32 + (mn%8)*4        [bytes]
// i.e. row-major rows of 32 mn (128 B) whose four 32-byte chunks are XOR-ed with the row index -- exactly what a
// TMA load with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B writes.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return make_smem_desc(saddr, lbo_bytes, sbo_bytes) | ((uint64_t)1 << 61);
}

// wait for two (three) barriers at once: the polls are issued back to back so their latencies (~200 cycles each,
// even when the phase completed long ago) overlap instead of adding up on the waiting thread's critical path
__device__ __forceinline__ void mbar_wait2_l(int line, uint64_t* b0, uint32_t p0, uint64_t* b1, uint32_t p1) {
    dbg_jitter();
    const uint32_t a0 = smem_u32(b0), a1 = smem_u32(b1);
    uint32_t d0 = 0, d1 = 0;
    for (uint32_t it = 0; it < MBAR_SPINS; ++it) {
        asm volatile(
            "{\n\t.reg .pred p, q;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%2], %3, %6;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 q, [%4], %5, %6;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "selp.u32 %1, 1, 0, q;\n\t}\n"
            : "=r"(d0), "=r"(d1) : "r"(a0), "r"(p0), "r"(a1), "r"(p1), "r"(MBAR_SUSPEND_NS) : "memory");
        if (d0 & d1) return;
    }
    mbar_timeout(line);
}
__device__ __forceinline__ void mbar_wait3_l(int line, uint64_t* b0, uint32_t p0, uint64_t* b1, uint32_t p1, uint64_t* b2, uint32_t p2) {
    dbg_jitter();
    const uint32_t a0 = smem_u32(b0), a1 = smem_u32(b1), a2 = smem_u32(b2);
    uint32_t d0 = 0, d1 = 0, d2 = 0;
    for (uint32_t it = 0; it < MBAR_SPINS; ++it) {
        asm volatile(
            "{\n\t.reg .pred p, q, r;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%3], %4, %9;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 q, [%5], %6, %9;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 r, [%7], %8, %9;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "selp.u32 %1, 1, 0, q;\n\t"
            "selp.u32 %2, 1, 0, r;\n\t}\n"
            : "=r"(d0), "=r"(d1), "=r"(d2) : "r"(a0), "r"(p0), "r"(a1), "r"(p1), "r"(a2), "r"(p2), "r"(MBAR_SUSPEND_NS) : "memory");
        if (d0 & d1 & d2) return;
    }
    mbar_timeout(line);
}

// ---- 3xTF32 split ---------------------------------------------------------------------------------------
// round-to-nearest on both parts keeps the representation error unbiased (a truncating split showed a
// systematic ~3e-6 relative error at K=320 on hardware; with rounding it is random-walk ~1e-7)
// The rounding is done with two integer ALU ops on the bit pattern (add half an ulp of TF32, clear the low
// 13 bits = round-half-away on the magnitude): cvt.rna.tf32.f32 would do the same but issues on the XU
// pipe (16 lanes/clk/SM), which ncu showed saturated (192 % of peak) in the first dw_tc version.
// lo = x - hi is exact and has a random sign, so the tensor core truncating it to TF32 adds no bias.
__device__ __forceinline__ float rna_tf32(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = rna_tf32(x);
    lo = x - hi;
}
__device__ __forceinline__ void split4(const float4& v, float4& hi, float4& lo) {
    split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y);
    split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
}

// D[128 x N] (+)= A[128 x 8*ksteps] * B[N x 8*ksteps]^T in 3xTF32.  a_hi/a_lo/b_hi/b_lo are shared-memory
// byte addresses of K-group-major tiles with a_rows / b_rows rows.  One thread calls this.
__device__ __forceinline__ void issue_3xtf32(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, int a_rows,
                                             uint32_t b_hi, uint32_t b_lo, int b_rows, int ksteps,
                                             uint32_t idesc, bool accumulate_first) {
    const uint32_t a_lbo = a_rows * 16, b_lbo = b_rows * 16;
    uint32_t acc = accumulate_first ? 1u : 0u;
    for (int k = 0; k < ksteps; ++k) {
        const uint32_t ao = 2 * k * a_lbo, bo = 2 * k * b_lbo;
        uint64_t ah = make_smem_desc(a_hi + ao, a_lbo, 128), al = make_smem_desc(a_lo + ao, a_lbo, 128);
        uint64_t bh = make_smem_desc(b_hi + bo, b_lbo, 128), bl = make_smem_desc(b_lo + bo, b_lbo, 128);
        umma_tf32(tmem_d, al, bh, idesc, acc);      // small terms first
        umma_tf32(tmem_d, ah, bl, idesc, 1u);
        umma_tf32(tmem_d, ah, bh, idesc, 1u);
        acc = 1u;
    }
}

}  // namespace tc
}  // namespace dcgru
