// Shared geometry and device helpers for the generic (fp32 SIMT) DCGRU kernels.
//
// Thread/tile geometry (all kernels use 256 threads):
//   * a CTA owns SB consecutive samples for the whole sequence; nodes are padded 19 -> NP = 20,
//     so the CTA's activation tile has R = 20*SB rows, row = s*NP + n
//   * GEMM thread tile = 5 rows x TN columns; rows of a thread are {i*RG + g}, i = 0..4, with
//     RG = R/5 row groups; a warp holds SPW = 32/RG column slices, the CTA NSLICE = 64/SB
//   * A operands live in shared memory K-major ([k][RLD], RLD = R+1 odd => conflict-free
//     transposing stores), B operands (weights) are streamed from L2 with cp.async
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dcgru {

constexpr int NT = 256;   // threads per CTA
constexpr int NP = 20;    // padded nodes per sample
constexpr int MAXL = 4;   // cells per decoder launch (DCGRU_MAX_LAYERS)

template <int SB>
struct Geo {
    static constexpr int R = SB * NP;
    static constexpr int RLD = R + 1;
    static constexpr int RG = R / 5;       // 4*SB
    static constexpr int SPW = 32 / RG;    // column slices per warp
    static constexpr int NSLICE = 8 * SPW; // 64/SB
    static_assert(SB == 1 || SB == 2 || SB == 4 || SB == 8, "SB");
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

template <int TN>
__device__ __forceinline__ void load_vec(float (&w)[TN], const float* p) {
    if constexpr (TN % 4 == 0) {
#pragma unroll
        for (int j = 0; j < TN / 4; ++j) {
            float4 v = *reinterpret_cast<const float4*>(p + 4 * j);
            w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
        }
    } else if constexpr (TN % 2 == 0) {
#pragma unroll
        for (int j = 0; j < TN / 2; ++j) {
            float2 v = *reinterpret_cast<const float2*>(p + 2 * j);
            w[2 * j] = v.x; w[2 * j + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < TN; ++j) w[j] = p[j];
    }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// acc[5][TN] += A[k][rows] * W[k][cols]   (A K-major with leading dim RLD, W row-major ldw)
template <int SB, int TN>
__device__ __forceinline__ void gemm_tile(float (&acc)[5][TN], const float* A, const float* W,
                                          int klen, int ldw, int g, int col0) {
    constexpr int RLD = Geo<SB>::RLD, RG = Geo<SB>::RG;
    const float* ap = A + g;
    const float* wp = W + col0;
#pragma unroll 2
    for (int k = 0; k < klen; ++k) {
        float a[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) a[i] = ap[k * RLD + i * RG];
        float w[TN];
        load_vec<TN>(w, wp + (size_t)k * ldw);
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
}

// ---- host/device shared smem layout of the forward sequence kernel (units: floats) ----------
struct FwdLayout {
    int pt, z, out, gs, wb, total;   // offsets
    int zld, old_, wbuf;             // leading dims / size of ONE weight buffer
};
__host__ __device__ inline FwdLayout fwd_layout(int SB, int H, int Cmax, int M, int KC) {
    FwdLayout L;
    const int R = SB * NP, RLD = R + 1;
    L.zld = Cmax + 4;
    L.old_ = 3 * H + 4;
    L.wbuf = KC * M * 3 * H;
    int o = 0;
    L.pt = o;  o += SB * (M - 1) * NP * NP;
    L.z = o;   o += R * L.zld;
    L.out = o; { int a = R * L.old_, b = H * RLD; o += (a > b ? a : b); }
    o = (o + 3) & ~3;
    L.gs = o;  o += KC * M * RLD;
    o = (o + 3) & ~3;
    L.wb = o;  o += 2 * L.wbuf;
    L.total = o;
    return L;
}

// ---- backward sequence kernel -----------------------------------------------------------------
constexpr int BWD_OC = 16;   // weight rows (o) per streamed piece
struct BwdLayout {
    int p, da, gs, wp, dh, drh, total;
    int kb;                          // kk columns per chunk (= NSLICE*8)
};
__host__ __device__ inline BwdLayout bwd_layout(int SB, int H, int M, int Fo) {
    BwdLayout L;
    const int R = SB * NP, RLD = R + 1;
    L.kb = 512 / SB;
    int o = 0;
    L.p = o;   o += SB * (M - 1) * NP * NP;
    L.da = o;  o += (3 * H > Fo ? 3 * H : Fo) * RLD;   // decoder stages dY (Fo rows) here too
    o = (o + 3) & ~3;
    L.gs = o;  o += L.kb * RLD;
    o = (o + 3) & ~3;
    L.wp = o;  o += 2 * BWD_OC * L.kb;
    L.dh = o;  o += R * H;
    L.drh = o; o += R * H;
    L.total = o;
    return L;
}
// z-columns per backward chunk: largest multiple of 4 with cols*M <= kb
__host__ __device__ inline int bwd_zcols(int kb, int M) { return (kb / M) & ~3; }

struct CellW {           // forward weights of one cell
    const float *Wg, *bg, *Wc, *bc;
    int fin;
};
struct CellWT {          // transposed weights of one cell (backward)
    const float *WgT;    // (2H, C*M)
    const float *WcT;    // (H,  C*M)
    int fin;
};

struct FwdParams {
    int B, T, N, H, M, act, ncell, KC, mode;   // mode 0 = encoder layer, 1 = decoder
    CellW cell[MAXL];
    const float* P;            // (B, M-1, N, N)
    const float* x;            // encoder input sequence
    long long xs_t, xs_b;
    const float* h0;           // enc (B,NH) | dec (L,B,NH)
    float* hseq;               // enc (T,B,NH) | dec h_all (T,L,B,NH)
    float* ruc;                // enc (T,B,N,3H) | dec (T,L,B,N,3H) | null
    const float* targets;      // dec (T,B,N*Fo) | null
    unsigned long long teacher_mask;
    const float* projWT;       // dec (H, FoPad) transposed, zero padded
    const float* projb;        // dec (Fo)
    const float* dropmask;     // dec (T,B,N,H) | null
    float* out;                // dec (T,B,N*Fo)
    int Fo, FoPad;
};

struct BwdParams {
    int B, T, N, H, M, act, ncell, mode;
    CellWT cell[MAXL];
    const float* P;
    const float* h0;
    const float* hseq;         // enc (T,B,NH) | dec h_all
    const float* ruc;
    const float* d_hseq;       // enc upstream (T,B,NH) | null
    const float* d_hlast;      // enc upstream (B,NH) | null
    float* dx;                 // enc (T,B,N*Fin) | null
    float* dh0;                // enc (B,NH) | dec (L,B,NH) (used as the running carry)
    float* dA;                 // enc (T,B,N,3H) | dec (T,L,B,N,3H)
    // decoder only
    const float* d_out;        // (T,B,N*Fo)
    const float* proj_w;       // (Fo,H)
    const float* dropmask;
    unsigned long long teacher_mask;
    float* dY;                 // (T,B,N*Fo) total gradient wrt every projected output
    float* scratch;            // (B, N*max(Fo,H)) hand-off buffer between cells
    int Fo;
};

}  // namespace dcgru
