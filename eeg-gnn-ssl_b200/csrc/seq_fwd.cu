// Persistent forward sequence kernel (generic fp32 path).
//
// One CTA = SB samples for the whole sequence, no relaunch between steps:
//   mode 0: one encoder layer, T steps            (model/model.py:93-96  x  model/cell.py:182-210)
//   mode 1: decoder, To steps x L cells + Linear  (model/model.py:182-202)
// Per cell step:  Z = [x | h]  ->  x-part projection of gate and candidate in one pass
// (3H columns), recurrent gate part, sigmoid / r*h, recurrent candidate part, tanh / GRU update.
// Diffusion is applied chunk-wise with the per-sample polynomial matrices P_m (shared memory)
// right before each K-chunk of the projection, so the (R, C*M) diffused matrix never exists.
#include "common.cuh"

namespace dcgru {

struct WSrc {            // weight rows [kk][n0 | n1] gathered from up to two matrices
    const float* w0; int n0;
    const float* w1; int n1;
};

__device__ __forceinline__ void load_w_chunk(float* dst, const WSrc& ws, int row0, int nrows) {
    const int nc = ws.n0 + ws.n1;
    const int q4 = nc >> 2;
    for (int idx = threadIdx.x; idx < nrows * q4; idx += NT) {
        int r = idx / q4, col = (idx - r * q4) << 2;
        const float* src = (col < ws.n0)
                               ? ws.w0 + (size_t)(row0 + r) * ws.n0 + col
                               : ws.w1 + (size_t)(row0 + r) * ws.n1 + (col - ws.n0);
        cp_async16(dst + r * nc + col, src);
    }
}

// Gs[(cc*M + m)][row] = (P_m Z)[row][c0 + cc]   for cc < kc, all m, all rows of the CTA
template <int SB>
__device__ __forceinline__ void diffuse_chunk(const float* Z, int zld, int c0, int kc,
                                              const float* PT, int M, int N, float* Gs) {
    constexpr int R = Geo<SB>::R, RLD = Geo<SB>::RLD;
    // m = 0: identity
    for (int idx = threadIdx.x; idx < R * kc; idx += NT) {
        int row = idx / kc, cc = idx - row * kc;
        Gs[(cc * M) * RLD + row] = Z[row * zld + c0 + cc];
    }
    // m >= 1: 5 nodes per task
    const int ntask = SB * (M - 1) * 4 * kc;
    for (int id = threadIdx.x; id < ntask; id += NT) {
        int cc = id % kc;
        int t1 = id / kc;
        int q = t1 & 3;
        int sm = t1 >> 2;                 // s*(M-1) + m1
        int s = sm / (M - 1);
        int m1 = sm - s * (M - 1);
        const float* zp = Z + (s * NP) * zld + c0 + cc;
        const float* pp = PT + (size_t)sm * NP * NP + 5 * q;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
        for (int j = 0; j < N; ++j) {
            float z = zp[j * zld];
            const float* pj = pp + j * NP;
            a0 = fmaf(pj[0], z, a0); a1 = fmaf(pj[1], z, a1); a2 = fmaf(pj[2], z, a2);
            a3 = fmaf(pj[3], z, a3); a4 = fmaf(pj[4], z, a4);
        }
        float* gp = Gs + (cc * M + m1 + 1) * RLD + s * NP + 5 * q;
        gp[0] = a0; gp[1] = a1; gp[2] = a2; gp[3] = a3; gp[4] = a4;   // PT pad rows are zero
    }
}

struct Ctx {
    float *PT, *Z, *OUT, *Gs, *Wb;
    int zld, old_, wbuf;
    int B, N, H, M, KC, b0;
    int g, slice;
};

// OUT[:, outcol0 : outcol0 + NC] (=|+=) diffuse(Z[:, zc0:zc1]) @ W[zc0*M : zc1*M, :]
template <int SB, int TN>
__device__ void gemm_phase(const Ctx& c, int zc0, int zc1, const WSrc& ws, int outcol0,
                           const float* bias0, const float* bias1) {
    constexpr int RG = Geo<SB>::RG;
    const int NC = ws.n0 + ws.n1;
    const int col0 = c.slice * TN;
    float acc[5][TN];
    if (bias0 != nullptr) {
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int col = col0 + j;
            float b = (col < ws.n0) ? bias0[col] : bias1[col - ws.n0];
#pragma unroll
            for (int i = 0; i < 5; ++i) acc[i][j] = b;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j)
                acc[i][j] = c.OUT[(i * RG + c.g) * c.old_ + outcol0 + col0 + j];
    }
    const int M = c.M, KC = c.KC;
    const int nchunks = (zc1 - zc0 + KC - 1) / KC;
    {
        int kc = min(KC, zc1 - zc0);
        load_w_chunk(c.Wb, ws, zc0 * M, kc * M);
        cp_async_commit();
    }
    for (int ci = 0; ci < nchunks; ++ci) {
        const int c0 = zc0 + ci * KC;
        const int kc = min(KC, zc1 - c0);
        if (ci + 1 < nchunks) {
            int c0n = c0 + KC, kcn = min(KC, zc1 - c0n);
            load_w_chunk(c.Wb + ((ci + 1) & 1) * c.wbuf, ws, c0n * M, kcn * M);
            cp_async_commit();
        }
        diffuse_chunk<SB>(c.Z, c.zld, c0, kc, c.PT, M, c.N, c.Gs);
        if (ci + 1 < nchunks) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncthreads();
        gemm_tile<SB, TN>(acc, c.Gs, c.Wb + (ci & 1) * c.wbuf, kc * M, NC, c.g, col0);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j)
            c.OUT[(i * RG + c.g) * c.old_ + outcol0 + col0 + j] = acc[i][j];
    __syncthreads();   // the next phase maps threads to other columns of OUT
}

// One DCGRU cell step for the CTA's samples.  All pointers are global; xin rows are
// (b, n) -> xin + b*xin_sb + n*fin.  Ends with a __syncthreads() after the global stores.
template <int SB, int TNC>
__device__ void cell_fwd(const Ctx& c, const CellW& cw, const float* xin, long long xin_sb,
                         const float* hprev, float* hout, float* ruc_out, int act) {
    constexpr int R = Geo<SB>::R;
    const int H = c.H, N = c.N, fin = cw.fin, B = c.B;
    // ---- Z = [x | hprev] -------------------------------------------------------------------
    {
        const int c4 = (fin + H) >> 2;
        for (int idx = threadIdx.x; idx < R * c4; idx += NT) {
            int row = idx / c4, col = (idx - row * c4) << 2;
            int s = row / NP, n = row - s * NP, b = c.b0 + s;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < N && b < B) {
                if (col < fin) {
                    if (xin != nullptr)
                        v = __ldcg(reinterpret_cast<const float4*>(xin + (size_t)b * xin_sb + n * fin + col));
                } else {
                    v = __ldcg(reinterpret_cast<const float4*>(hprev + ((size_t)b * N + n) * H + col - fin));
                }
            }
            *reinterpret_cast<float4*>(c.Z + row * c.zld + col) = v;
        }
    }
    __syncthreads();
    // ---- x part of gate and candidate: OUT[:, 0:3H] = bias + diffuse(x) @ [Wg_x | Wc_x] -------
    {
        WSrc ws{cw.Wg, 2 * H, cw.Wc, H};
        gemm_phase<SB, 3 * TNC>(c, 0, fin, ws, 0, cw.bg, cw.bc);
    }
    // ---- recurrent part of the gate: OUT[:, 0:2H] += diffuse(h) @ Wg_h --------------------------
    {
        WSrc ws{cw.Wg, 2 * H, nullptr, 0};
        gemm_phase<SB, 2 * TNC>(c, fin, fin + H, ws, 0, nullptr, nullptr);
    }
    // ---- r, u; Z_h <- r*h; OUT[:,0:H] <- h (kept for the update), OUT[:,H:2H] <- u ---------------
    for (int idx = threadIdx.x; idx < R * H; idx += NT) {
        int row = idx / H, col = idx - row * H;
        float* o = c.OUT + row * c.old_;
        float r = sigmoidf_(o[col]);
        float u = sigmoidf_(o[H + col]);
        float hp = c.Z[row * c.zld + fin + col];
        c.Z[row * c.zld + fin + col] = r * hp;
        o[col] = hp;
        o[H + col] = u;
        if (ruc_out != nullptr) {
            int s = row / NP, n = row - s * NP, b = c.b0 + s;
            if (n < N && b < B) {
                float* q = ruc_out + ((size_t)b * N + n) * 3 * H;
                q[col] = r;
                q[H + col] = u;
            }
        }
    }
    __syncthreads();
    // ---- recurrent part of the candidate: OUT[:, 2H:3H] += diffuse(r*h) @ Wc_h ------------------
    {
        WSrc ws{cw.Wc, H, nullptr, 0};
        gemm_phase<SB, TNC>(c, fin, fin + H, ws, 2 * H, nullptr, nullptr);
    }
    // ---- c = act(.), h' = u*h + (1-u)*c ----------------------------------------------------------
    for (int idx = threadIdx.x; idx < R * H; idx += NT) {
        int row = idx / H, col = idx - row * H;
        int s = row / NP, n = row - s * NP, b = c.b0 + s;
        if (n < N && b < B) {
            const float* o = c.OUT + row * c.old_;
            float pre = o[2 * H + col];
            float cv = (act == 0) ? tanhf(pre) : fmaxf(pre, 0.f);
            float hp = o[col], u = o[H + col];
            float hn = u * hp + (1.f - u) * cv;
            hout[((size_t)b * N + n) * H + col] = hn;
            if (ruc_out != nullptr) ruc_out[((size_t)b * N + n) * 3 * H + 2 * H + col] = cv;
        }
    }
    __syncthreads();
}

// y = Linear(dropout(top)) for the CTA's samples (model/model.py:192-196)
template <int SB>
__device__ void project_fwd(const Ctx& c, const FwdParams& p, const float* top,
                            const float* mask, float* yout) {
    constexpr int R = Geo<SB>::R, RLD = Geo<SB>::RLD, RG = Geo<SB>::RG, NS = Geo<SB>::NSLICE;
    const int H = c.H, N = c.N, B = c.B, Fo = p.Fo, FoPad = p.FoPad;
    float* TT = c.OUT;                                   // [H][RLD]
    for (int idx = threadIdx.x; idx < H * FoPad / 4; idx += NT)
        cp_async16(c.Wb + idx * 4, p.projWT + idx * 4);
    cp_async_commit();
    for (int idx = threadIdx.x; idx < R * H; idx += NT) {
        int row = idx / H, col = idx - row * H;
        int s = row / NP, n = row - s * NP, b = c.b0 + s;
        float v = 0.f;
        if (n < N && b < B) {
            size_t off = ((size_t)b * N + n) * H + col;
            v = __ldcg(top + off);
            if (mask != nullptr) v *= mask[off];
        }
        TT[col * RLD + row] = v;
    }
    cp_async_wait<0>();
    __syncthreads();
    const int nblk = FoPad / (NS * 4);
    for (int cb = 0; cb < nblk; ++cb) {
        float acc[5][4];
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        const int col0 = cb * NS * 4 + c.slice * 4;
        gemm_tile<SB, 4>(acc, TT, c.Wb, H, FoPad, c.g, col0);
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            int row = i * RG + c.g;
            int s = row / NP, n = row - s * NP, b = c.b0 + s;
            if (n < N && b < B) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    int col = col0 + j;
                    if (col < Fo) yout[((size_t)b * N + n) * Fo + col] = acc[i][j] + p.projb[col];
                }
            }
        }
    }
    __syncthreads();
}

template <int SB, int TNC>
__global__ void __launch_bounds__(NT, 1) seq_fwd_kernel(const FwdParams p) {
    extern __shared__ __align__(16) float smem[];
    constexpr int RG = Geo<SB>::RG;
    int cmax = 0;
    for (int l = 0; l < p.ncell; ++l) cmax = max(cmax, p.cell[l].fin + p.H);
    const FwdLayout L = fwd_layout(SB, p.H, cmax, p.M, p.KC);
    Ctx c;
    c.PT = smem + L.pt; c.Z = smem + L.z; c.OUT = smem + L.out; c.Gs = smem + L.gs; c.Wb = smem + L.wb;
    c.zld = L.zld; c.old_ = L.old_; c.wbuf = L.wbuf;
    c.B = p.B; c.N = p.N; c.H = p.H; c.M = p.M; c.KC = p.KC;
    c.b0 = blockIdx.x * SB;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    c.g = lane % RG;
    c.slice = warp * Geo<SB>::SPW + lane / RG;

    // PT[s][m1][j][n] = P[b][m1][n][j], zero padded
    const int N = p.N, M1 = p.M - 1;
    for (int idx = threadIdx.x; idx < SB * M1 * NP * NP; idx += NT) {
        int n = idx % NP, j = (idx / NP) % NP, sm = idx / (NP * NP);
        int s = sm / max(M1, 1), m1 = sm - s * max(M1, 1);
        int b = c.b0 + s;
        float v = 0.f;
        if (n < N && j < N && b < p.B) v = p.P[(((size_t)b * M1 + m1) * N + n) * N + j];
        c.PT[idx] = v;
    }
    __syncthreads();

    const size_t NH = (size_t)p.N * p.H;
    if (p.mode == 0) {
        for (int t = 0; t < p.T; ++t) {
            const float* hprev = (t == 0) ? p.h0 : p.hseq + (size_t)(t - 1) * p.B * NH;
            float* hout = p.hseq + (size_t)t * p.B * NH;
            float* ruc = p.ruc ? p.ruc + (size_t)t * p.B * NH * 3 : nullptr;
            cell_fwd<SB, TNC>(c, p.cell[0], p.x + (size_t)t * p.xs_t, p.xs_b, hprev, hout, ruc, p.act);
        }
    } else {
        const int Lc = p.ncell;
        const size_t NFo = (size_t)p.N * p.Fo;
        for (int t = 0; t < p.T; ++t) {
            for (int l = 0; l < Lc; ++l) {
                const float* xin;
                long long xsb;
                if (l == 0) {
                    xsb = (long long)NFo;
                    if (t == 0) xin = nullptr;                                        // GO symbol = zeros
                    else if ((p.teacher_mask >> (t - 1)) & 1ull) xin = p.targets + (size_t)(t - 1) * p.B * NFo;
                    else xin = p.out + (size_t)(t - 1) * p.B * NFo;
                } else {
                    xsb = (long long)NH;
                    xin = p.hseq + ((size_t)t * Lc + (l - 1)) * p.B * NH;
                }
                const float* hprev = (t == 0) ? p.h0 + (size_t)l * p.B * NH
                                              : p.hseq + ((size_t)(t - 1) * Lc + l) * p.B * NH;
                float* hout = p.hseq + ((size_t)t * Lc + l) * p.B * NH;
                float* ruc = p.ruc ? p.ruc + ((size_t)t * Lc + l) * p.B * NH * 3 : nullptr;
                cell_fwd<SB, TNC>(c, p.cell[l], xin, xsb, hprev, hout, ruc, p.act);
            }
            const float* top = p.hseq + ((size_t)t * Lc + (Lc - 1)) * p.B * NH;
            const float* mask = p.dropmask ? p.dropmask + (size_t)t * p.B * NH : nullptr;
            project_fwd<SB>(c, p, top, mask, p.out + (size_t)t * p.B * NFo);
        }
    }
}

// ---- launcher -----------------------------------------------------------------------------------
template <int SB, int TNC>
static cudaError_t launch_one(const FwdParams& p, int smem_bytes, cudaStream_t st) {
    auto k = seq_fwd_kernel<SB, TNC>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    int grid = (p.B + SB - 1) / SB;
    k<<<grid, NT, smem_bytes, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_seq_fwd(const FwdParams& p, int SB, int smem_bytes, cudaStream_t st) {
    const int tnc = p.H * SB / 64;
#define CASE(sb, tn) if (SB == sb && tnc == tn) return launch_one<sb, tn>(p, smem_bytes, st);
    CASE(1, 1) CASE(2, 2) CASE(4, 4)           // H = 64
    CASE(1, 2) CASE(2, 4)                      // H = 128
    CASE(2, 1) CASE(4, 2) CASE(8, 4)           // H = 32
#undef CASE
    return cudaErrorInvalidValue;
}

}  // namespace dcgru
