// Persistent backward (BPTT) sequence kernel (generic fp32 path) -- SURVEY A.4.
//
// Same CTA <-> samples mapping as the forward kernel, time runs backwards.  Per cell step:
//   E1  du, dc, dH*u, dA_c, dA_u                                  (elementwise)
//   B1  d(rH)  = diffT( dA_c @ Wc_h^T )                            (K = H)
//   E2  dr, dH += d(rH)*r, dA_r
//   B2  dH    += diffT( [dA_r|dA_u] @ Wg_h^T )                     (K = 2H)
//   B3  dX     = diffT( [dA_r|dA_u|dA_c] @ [Wg_x|Wc_x]^T )         (K = 3H, only if dX is needed)
// and dA = [dA_r|dA_u|dA_c] is stored for the bulk weight-gradient kernel (dw.cu), which is
// where dW = G^T dA is formed (it is not recurrent).  diffT applies sum_m P_m^T to the M column
// groups of the product.  The weights arrive pre-transposed ((out, C*M), see capi.cu).
#include "common.cuh"

namespace dcgru {

struct BCtx {
    float *P, *DA, *Gs, *Wp, *DH, *DRH;
    int kb, zcb;
    int B, N, H, M, b0;
    int g, slice;
};

struct WTSrc {            // rows o of the stacked transposed weights: o < n0 from w0, else w1
    const float* w0; int n0;
    const float* w1;
    int ld;               // C*M
};

__device__ __forceinline__ void load_wt_piece(float* dst, const WTSrc& ws, int o0, int nrows,
                                              int kk0, int kcols, int kb) {
    const int q4 = kcols >> 2;
    for (int idx = threadIdx.x; idx < nrows * q4; idx += NT) {
        int r = idx / q4, col = (idx - r * q4) << 2;
        int o = o0 + r;
        const float* src = (o < ws.n0) ? ws.w0 + (size_t)o * ws.ld : ws.w1 + (size_t)(o - ws.n0) * ws.ld;
        cp_async16(dst + r * kb + col, src + kk0 + col);
    }
}

// Gs[kkl][row] = sum_{o<K} A[o][row] * WT[o][kk0 + kkl]     for kkl < kcols
template <int SB>
__device__ void gemm_bwd(const BCtx& c, const float* A, int K, const WTSrc& ws, int kk0, int kcols) {
    constexpr int RG = Geo<SB>::RG, RLD = Geo<SB>::RLD;
    float acc[5][8];
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const int kb = c.kb;
    const int col0 = c.slice * 8;
    const int np = (K + BWD_OC - 1) / BWD_OC;
    load_wt_piece(c.Wp, ws, 0, min(BWD_OC, K), kk0, kcols, kb);
    cp_async_commit();
    for (int pi = 0; pi < np; ++pi) {
        const int o0 = pi * BWD_OC, rows = min(BWD_OC, K - o0);
        if (pi + 1 < np) {
            load_wt_piece(c.Wp + ((pi + 1) & 1) * BWD_OC * kb, ws, o0 + BWD_OC,
                          min(BWD_OC, K - o0 - BWD_OC), kk0, kcols, kb);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (col0 < kcols)
            gemm_tile<SB, 8>(acc, A + (size_t)o0 * RLD, c.Wp + (pi & 1) * BWD_OC * kb, rows, kb, c.g, col0);
        __syncthreads();
    }
    if (col0 < kcols) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (col0 + j < kcols) {
#pragma unroll
                for (int i = 0; i < 5; ++i) c.Gs[(col0 + j) * RLD + i * RG + c.g] = acc[i][j];
            }
        }
    }
    __syncthreads();
}

// dZ[(s,j)][c0+ccl] = Gs[ccl*M][(s,j)] + sum_{m>=1} sum_n P_m[s][n][j] * Gs[ccl*M+m][(s,n)]
// mode 0: DST[row*ld + cl] = v ; mode 1: DST[row*ld + cl] += v ; mode 2: global rows (b,n)*ld + cl
template <int SB, int MODE>
__device__ __forceinline__ void diff_t(const BCtx& c, int ncolz, float* dst, int ld, int cl0) {
    constexpr int RLD = Geo<SB>::RLD;
    const int M = c.M, N = c.N;
    const int ntask = SB * 4 * ncolz;
    for (int id = threadIdx.x; id < ntask; id += NT) {
        int ccl = id % ncolz;
        int t1 = id / ncolz;
        int q = t1 & 3, s = t1 >> 2;
        const float* gp = c.Gs + (ccl * M) * RLD + s * NP;
        float a[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) a[i] = gp[5 * q + i];
        for (int m1 = 0; m1 < M - 1; ++m1) {
            const float* gm = gp + (m1 + 1) * RLD;
            const float* pp = c.P + (size_t)(s * (M - 1) + m1) * NP * NP + 5 * q;
            for (int n = 0; n < N; ++n) {
                float gv = gm[n];
                const float* pr = pp + n * NP;
#pragma unroll
                for (int i = 0; i < 5; ++i) a[i] = fmaf(pr[i], gv, a[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            int j = 5 * q + i;
            if (j < N) {
                if (MODE == 0) dst[(s * NP + j) * ld + cl0 + ccl] = a[i];
                else if (MODE == 1) dst[(s * NP + j) * ld + cl0 + ccl] += a[i];
                else {
                    int b = c.b0 + s;
                    if (b < c.B) dst[((size_t)b * N + j) * ld + cl0 + ccl] = a[i];
                }
            }
        }
    }
    __syncthreads();
}

// One cell step backward.  In: c.DH = dL/dh_new (rows of this CTA).  Out: c.DH = dL/dh_prev,
// dA_out (global (B,N,3H)), dx_out (global rows (b,n)*fin, or null).
template <int SB>
__device__ void cell_bwd(const BCtx& c, const CellWT& cw, const float* hprev, const float* ruc,
                         float* dA_out, float* dx_out, int act) {
    constexpr int R = Geo<SB>::R, RLD = Geo<SB>::RLD;
    const int H = c.H, N = c.N, B = c.B, M = c.M, fin = cw.fin;
    const int CM = (fin + H) * M;
    // ---- E1 ------------------------------------------------------------------------------------
    for (int idx = threadIdx.x; idx < R * H; idx += NT) {
        int row = idx / H, col = idx - row * H;
        int s = row / NP, n = row - s * NP, b = c.b0 + s;
        float dAu = 0.f, dAc = 0.f, dhu = 0.f;
        if (n < N && b < B) {
            size_t ro = (size_t)b * N + n;
            float hp = __ldcg(hprev + ro * H + col);
            const float* q = ruc + ro * 3 * H;
            float u = q[H + col], cv = q[2 * H + col];
            float d = c.DH[row * H + col];
            float du = d * (hp - cv);
            float dc = d * (1.f - u);
            dhu = d * u;
            dAc = dc * ((act == 0) ? (1.f - cv * cv) : (cv > 0.f ? 1.f : 0.f));
            dAu = du * u * (1.f - u);
        }
        c.DH[row * H + col] = dhu;
        c.DA[(H + col) * RLD + row] = dAu;
        c.DA[(2 * H + col) * RLD + row] = dAc;
    }
    __syncthreads();
    // ---- B1: d(rH) -------------------------------------------------------------------------------
    {
        WTSrc ws{cw.WcT, H, nullptr, CM};
        for (int z0 = 0; z0 < H; z0 += c.zcb) {
            int ncolz = min(c.zcb, H - z0);
            gemm_bwd<SB>(c, c.DA + (size_t)2 * H * RLD, H, ws, (fin + z0) * M, ncolz * M);
            diff_t<SB, 0>(c, ncolz, c.DRH, H, z0);
        }
    }
    // ---- E2 ------------------------------------------------------------------------------------
    for (int idx = threadIdx.x; idx < R * H; idx += NT) {
        int row = idx / H, col = idx - row * H;
        int s = row / NP, n = row - s * NP, b = c.b0 + s;
        float dAr = 0.f;
        if (n < N && b < B) {
            size_t ro = (size_t)b * N + n;
            float hp = __ldcg(hprev + ro * H + col);
            float r = ruc[ro * 3 * H + col];
            float drh = c.DRH[row * H + col];
            c.DH[row * H + col] += drh * r;
            dAr = drh * hp * r * (1.f - r);
        }
        c.DA[col * RLD + row] = dAr;
    }
    __syncthreads();
    // ---- B2: dH += diffT([dAr|dAu] @ Wg_h^T) -------------------------------------------------------
    {
        WTSrc ws{cw.WgT, 2 * H, nullptr, CM};
        for (int z0 = 0; z0 < H; z0 += c.zcb) {
            int ncolz = min(c.zcb, H - z0);
            gemm_bwd<SB>(c, c.DA, 2 * H, ws, (fin + z0) * M, ncolz * M);
            diff_t<SB, 1>(c, ncolz, c.DH, H, z0);
        }
    }
    // ---- B3: dX ----------------------------------------------------------------------------------
    if (dx_out != nullptr) {
        WTSrc ws{cw.WgT, 2 * H, cw.WcT, CM};
        for (int z0 = 0; z0 < fin; z0 += c.zcb) {
            int ncolz = min(c.zcb, fin - z0);
            gemm_bwd<SB>(c, c.DA, 3 * H, ws, z0 * M, ncolz * M);
            diff_t<SB, 2>(c, ncolz, dx_out, fin, z0);
        }
    }
    // ---- store dA (B,N,3H) -----------------------------------------------------------------------
    {
        const int H3 = 3 * H;
        for (int idx = threadIdx.x; idx < R * H3; idx += NT) {
            int row = idx / H3, o = idx - row * H3;
            int s = row / NP, n = row - s * NP, b = c.b0 + s;
            if (n < N && b < B) dA_out[((size_t)b * N + n) * H3 + o] = c.DA[o * RLD + row];
        }
    }
    __syncthreads();
}

// DH (=|+=) global rows
template <int SB>
__device__ __forceinline__ void dh_load(const BCtx& c, const float* src, bool accumulate) {
    constexpr int R = Geo<SB>::R;
    const int H = c.H, N = c.N;
    for (int idx = threadIdx.x; idx < R * H; idx += NT) {
        int row = idx / H, col = idx - row * H;
        int s = row / NP, n = row - s * NP, b = c.b0 + s;
        float v = 0.f;
        if (src != nullptr && n < N && b < c.B) v = __ldcg(src + ((size_t)b * N + n) * H + col);
        if (accumulate) c.DH[idx] += v; else c.DH[idx] = v;
    }
}
template <int SB>
__device__ __forceinline__ void dh_store(const BCtx& c, float* dst) {
    constexpr int R = Geo<SB>::R;
    const int H = c.H, N = c.N;
    for (int idx = threadIdx.x; idx < R * H; idx += NT) {
        int row = idx / H, col = idx - row * H;
        int s = row / NP, n = row - s * NP, b = c.b0 + s;
        if (n < N && b < c.B) dst[((size_t)b * N + n) * H + col] = c.DH[idx];
    }
}

template <int SB>
__global__ void __launch_bounds__(NT, 1) seq_bwd_kernel(const BwdParams p) {
    extern __shared__ __align__(16) float smem[];
    constexpr int RG = Geo<SB>::RG, R = Geo<SB>::R, RLD = Geo<SB>::RLD;
    const BwdLayout L = bwd_layout(SB, p.H, p.M, p.mode == 1 ? p.Fo : 0);
    BCtx c;
    c.P = smem + L.p; c.DA = smem + L.da; c.Gs = smem + L.gs; c.Wp = smem + L.wp;
    c.DH = smem + L.dh; c.DRH = smem + L.drh;
    c.kb = L.kb; c.zcb = bwd_zcols(L.kb, p.M);
    c.B = p.B; c.N = p.N; c.H = p.H; c.M = p.M;
    c.b0 = blockIdx.x * SB;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    c.g = lane % RG;
    c.slice = warp * Geo<SB>::SPW + lane / RG;

    const int N = p.N, M1 = p.M - 1, H = p.H;
    // P[s][m1][n][j] (untransposed), zero padded to NP x NP
    for (int idx = threadIdx.x; idx < SB * M1 * NP * NP; idx += NT) {
        int j = idx % NP, n = (idx / NP) % NP, sm = idx / (NP * NP);
        int s = sm / max(M1, 1), m1 = sm - s * max(M1, 1);
        int b = c.b0 + s;
        float v = 0.f;
        if (n < N && j < N && b < p.B) v = p.P[(((size_t)b * M1 + m1) * N + n) * N + j];
        c.P[idx] = v;
    }
    // pad rows of DA must be finite: they are multiplied into rows that are never stored
    for (int idx = threadIdx.x; idx < L.gs - L.da; idx += NT) c.DA[idx] = 0.f;
    for (int idx = threadIdx.x; idx < R * H; idx += NT) { c.DH[idx] = 0.f; c.DRH[idx] = 0.f; }
    __syncthreads();

    const size_t NH = (size_t)N * H;
    if (p.mode == 2) {
        // bulk input gradient (not recurrent): grid.y = step.  dX[t] = diffT(dA[t] @ [Wg_x | Wc_x]^T) with the
        // dA the recurrent kernel (tensor-core or this one) stored.
        const int t = blockIdx.y;
        const int fin = p.cell[0].fin, M = p.M, CM = (fin + H) * M, H3 = 3 * H;
        const float* dA = p.dA + (size_t)t * p.B * NH * 3;
        for (int idx = threadIdx.x; idx < R * H3; idx += NT) {
            int row = idx / H3, o = idx - row * H3;
            int s = row / NP, n = row - s * NP, b = c.b0 + s;
            c.DA[o * RLD + row] = (n < N && b < p.B) ? dA[((size_t)b * N + n) * H3 + o] : 0.f;
        }
        __syncthreads();
        WTSrc ws{p.cell[0].WgT, 2 * H, p.cell[0].WcT, CM};
        float* dx = p.dx + (size_t)t * p.B * N * fin;
        for (int z0 = 0; z0 < fin; z0 += c.zcb) {
            int ncolz = min(c.zcb, fin - z0);
            gemm_bwd<SB>(c, c.DA, 3 * H, ws, z0 * M, ncolz * M);
            diff_t<SB, 2>(c, ncolz, dx, fin, z0);
        }
    } else if (p.mode == 0) {
        const int fin = p.cell[0].fin;
        for (int t = p.T - 1; t >= 0; --t) {
            if (t == p.T - 1) dh_load<SB>(c, p.d_hlast, false);
            if (p.d_hseq != nullptr) dh_load<SB>(c, p.d_hseq + (size_t)t * p.B * NH, true);
            __syncthreads();
            const float* hprev = (t == 0) ? p.h0 : p.hseq + (size_t)(t - 1) * p.B * NH;
            float* dx = p.dx ? p.dx + (size_t)t * p.B * N * fin : nullptr;
            cell_bwd<SB>(c, p.cell[0], hprev, p.ruc + (size_t)t * p.B * NH * 3,
                         p.dA + (size_t)t * p.B * NH * 3, dx, p.act);
        }
        dh_store<SB>(c, p.dh0);
    } else {
        // decoder: time-outer (reversed), layer-inner (reversed); carries live in p.dh0 (zeroed
        // by the host), hand-off between cells through p.scratch (both global, L2 resident)
        const int Lc = p.ncell, Fo = p.Fo;
        const size_t NFo = (size_t)N * Fo;
        for (int t = p.T - 1; t >= 0; --t) {
            // ---- dY_t = d_out[t] (+ feedback from step t+1 unless it was teacher forced) ----------
            // staged K-major in DA ([Fo][RLD]); every cell_bwd below rewrites DA before using it
            const bool fb = (t + 1 < p.T) && !((p.teacher_mask >> t) & 1ull);
            for (int idx = threadIdx.x; idx < R * Fo; idx += NT) {
                int row = idx / Fo, f = idx - row * Fo;
                int s = row / NP, n = row - s * NP, b = c.b0 + s;
                float v = 0.f;
                if (n < N && b < p.B) {
                    size_t off = ((size_t)b * N + n) * Fo + f;
                    v = p.d_out[(size_t)t * p.B * NFo + off];
                    if (fb) v += __ldcg(p.scratch + off);
                    p.dY[(size_t)t * p.B * NFo + off] = v;
                }
                c.DA[f * RLD + row] = v;
            }
            __syncthreads();
            // ---- dTop = (dY @ proj_w) * mask  -> staged in DRH (K = Fo) ------------------------------
            {
                // proj_w is (Fo, H) row-major == "transposed weight" layout with ld = H
                WTSrc ws{p.proj_w, Fo, nullptr, H};
                for (int h0c = 0; h0c < H; h0c += c.kb) {
                    int ncol = min(c.kb, H - h0c);
                    gemm_bwd<SB>(c, c.DA, Fo, ws, h0c, ncol);
                    const float* mask = p.dropmask ? p.dropmask + (size_t)t * p.B * NH : nullptr;
                    for (int idx = threadIdx.x; idx < R * ncol; idx += NT) {
                        int row = idx / ncol, cl = idx - row * ncol;
                        int s = row / NP, n = row - s * NP, b = c.b0 + s;
                        float v = 0.f;
                        if (n < N && b < p.B) {
                            v = c.Gs[cl * RLD + row];
                            if (mask) v *= mask[((size_t)b * N + n) * H + h0c + cl];
                        }
                        c.DRH[row * H + h0c + cl] = v;     // staged; DH is loaded per cell below
                    }
                    __syncthreads();
                }
            }
            for (int l = Lc - 1; l >= 0; --l) {
                // dH_new(l) = carry_l + (top ? dTop : dX of cell l+1)
                dh_load<SB>(c, p.dh0 + (size_t)l * p.B * NH, false);
                if (l == Lc - 1) {
                    for (int idx = threadIdx.x; idx < R * H; idx += NT) c.DH[idx] += c.DRH[idx];
                } else {
                    dh_load<SB>(c, p.scratch, true);
                }
                __syncthreads();
                const float* hprev = (t == 0) ? p.h0 + (size_t)l * p.B * NH
                                              : p.hseq + ((size_t)(t - 1) * Lc + l) * p.B * NH;
                const size_t cs = ((size_t)t * Lc + l) * p.B * NH * 3;
                // cell 0 fed by the GO symbol (t == 0) or by a teacher-forced target: dX goes nowhere
                float* dx = p.scratch;
                if (l == 0 && (t == 0 || ((p.teacher_mask >> (t - 1)) & 1ull))) dx = nullptr;
                cell_bwd<SB>(c, p.cell[l], hprev, p.ruc + cs, p.dA + cs, dx, p.act);
                dh_store<SB>(c, p.dh0 + (size_t)l * p.B * NH);
                __syncthreads();
            }
        }
    }
}

cudaError_t launch_seq_bwd(const BwdParams& p, int SB, int smem_bytes, cudaStream_t st) {
    dim3 grid((p.B + SB - 1) / SB, p.mode == 2 ? p.T : 1);
#define CASE(sb)                                                                                    \
    if (SB == sb) {                                                                                 \
        auto k = seq_bwd_kernel<sb>;                                                                \
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes); \
        if (e != cudaSuccess) return e;                                                             \
        k<<<grid, NT, smem_bytes, st>>>(p);                                                         \
        return cudaGetLastError();                                                                  \
    }
    CASE(1) CASE(2) CASE(4) CASE(8)
#undef CASE
    return cudaErrorInvalidValue;
}

}  // namespace dcgru
