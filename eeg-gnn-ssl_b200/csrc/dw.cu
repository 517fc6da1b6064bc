// Bulk (non-recurrent) part of the backward pass: weight and bias gradients.
//
//   dW[c*M+m][o] = sum over (step, sample, node) of  G[row][c*M+m] * dA[row][o]
// with G = diffuse([x | h_prev]) for the gate columns and diffuse([x | r*h_prev]) for the
// candidate columns (model/cell.py:73-116 run forward again on the saved activations) and
// dA = [dA_r | dA_u | dA_c] stored by seq_bwd.cu.  Also the decoder's Linear gradient
// (dproj_w = dY^T top, dproj_b = sum dY).
//
// grid = (jobs, splits).  A job is a 64-wide slab of kk = c*M+m rows times <= 192 output
// columns; a split is a strided subset of the 4-sample groups.  Each CTA keeps its output tile
// in registers over all its (group, step) pairs and writes one partial; reduce kernels sum the
// partials in a fixed order (deterministic).
#include "common.cuh"
#include "dw.cuh"

namespace dcgru {

constexpr int DSB = 4;               // samples per group
constexpr int DR = DSB * NP;         // 80 rows
constexpr int GLD = 68;              // Gt leading dim (64 kk + pad, 16B aligned rows)
constexpr int ZLD = 65;

__device__ __forceinline__ const float* dec_x_src(const DwParams& p, int t, long long* sb) {
    // input of decoder cell `layer` at step t (model/model.py:182-202)
    const size_t NH = (size_t)p.N * p.H;
    if (p.layer == 0) {
        *sb = (long long)p.N * p.Fo;
        if (t == 0) return nullptr;
        if ((p.teacher_mask >> (t - 1)) & 1ull) return p.targets + (size_t)(t - 1) * p.B * p.N * p.Fo;
        return p.out + (size_t)(t - 1) * p.B * p.N * p.Fo;
    }
    *sb = (long long)NH;
    return p.hseq + ((size_t)t * p.ncell + (p.layer - 1)) * p.B * NH;
}

template <int TO>
__device__ void dw_run(const DwParams& p, const DwJob& job, float* smem) {
    const int N = p.N, H = p.H, M = (job.type == 3) ? 1 : p.M, B = p.B;
    const int M1 = M - 1;
    float* PT = smem;                                   // [DSB][M1][NP][NP] (PT[j][n])
    float* Zc = PT + DSB * (p.M - 1) * NP * NP;         // [DR][ZLD]
    float* Gt = Zc + DR * ZLD;                          // [DR][GLD]
    Gt = (float*)(((uintptr_t)Gt + 15) & ~(uintptr_t)15);
    const int DLD = job.nco + 4;
    float* dAt = Gt + DR * GLD;                         // [DR][DLD]

    const int tid = threadIdx.x, ky = tid >> 4, ox = tid & 15;
    const int split = blockIdx.y;
    const int nz = job.nz, kcols = nz * M;
    const size_t NH = (size_t)N * H;
    const int H3 = 3 * H;

    float acc[4][TO];
    float dbacc[TO];
    float gsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < TO; ++j) {
        dbacc[j] = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][j] = 0.f;
    }
    const bool do_db = (job.type == 0 && job.z0 == 0 && ky == 0);
    const bool do_gsum = (job.type == 3 && job.o0 == 0 && ox == 0);

    for (int idx = tid; idx < DR * GLD; idx += NT) Gt[idx] = 0.f;
    for (int idx = tid; idx < DR * DLD; idx += NT) dAt[idx] = 0.f;
    __syncthreads();

    const int ngroups = (B + DSB - 1) / DSB;
    for (int sg = split; sg < ngroups; sg += p.nsplit) {
        const int b0 = sg * DSB;
        if (job.type != 3) {
            for (int idx = tid; idx < DSB * M1 * NP * NP; idx += NT) {
                int n = idx % NP, j = (idx / NP) % NP, sm = idx / (NP * NP);
                int s = sm / max(M1, 1), m1 = sm - s * max(M1, 1);
                int b = b0 + s;
                float v = 0.f;
                if (n < N && j < N && b < B) v = p.P[(((size_t)b * M1 + m1) * N + n) * N + j];
                PT[idx] = v;
            }
        }
        for (int t = 0; t < p.T; ++t) {
            // ---- sources of this (group, step) ----------------------------------------------------
            const float* xsrc = nullptr; long long xsb = 0;
            const float* hprev; const float* ruc_t; const float* dA_t;
            if (p.mode == 0) {
                xsrc = p.x + (size_t)t * p.xs_t; xsb = p.xs_b;
                hprev = (t == 0) ? p.h0 : p.hseq + (size_t)(t - 1) * B * NH;
                ruc_t = p.ruc + (size_t)t * B * NH * 3;
                dA_t = p.dA + (size_t)t * B * NH * 3;
            } else {
                xsrc = dec_x_src(p, t, &xsb);
                hprev = (t == 0) ? p.h0 + (size_t)p.layer * B * NH
                                 : p.hseq + ((size_t)(t - 1) * p.ncell + p.layer) * B * NH;
                size_t cs = ((size_t)t * p.ncell + p.layer) * B * NH * 3;
                ruc_t = p.ruc + cs;
                dA_t = p.dA + cs;
            }
            // ---- right operand tile: dA (or top*mask for the Linear job) ----------------------------
            {
                const int q4 = job.nco >> 2;
                for (int idx = tid; idx < DR * q4; idx += NT) {
                    int row = idx / q4, col = (idx - row * q4) << 2;
                    int s = row / NP, n = row - s * NP, b = b0 + s;
                    if (n >= N) continue;                       // pad rows stay zero
                    float* d = dAt + row * DLD + col;
                    if (b < B) {
                        if (job.type != 3) {
                            cp_async16(d, dA_t + ((size_t)b * N + n) * H3 + job.o0 + col);
                        } else {
                            size_t off = ((size_t)b * N + n) * H + job.o0 + col;
                            const float* top = p.hseq + ((size_t)t * p.ncell + (p.ncell - 1)) * B * NH;
                            float4 v = *reinterpret_cast<const float4*>(top + off);
                            if (p.dropmask != nullptr) {
                                float4 mk = *reinterpret_cast<const float4*>(p.dropmask + (size_t)t * B * NH + off);
                                v.x *= mk.x; v.y *= mk.y; v.z *= mk.z; v.w *= mk.w;
                            }
                            *reinterpret_cast<float4*>(d) = v;
                        }
                    } else {
                        *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                cp_async_commit();
            }
            // ---- left operand source columns --------------------------------------------------------
            for (int idx = tid; idx < DR * nz; idx += NT) {
                int row = idx / nz, cl = idx - row * nz;
                int s = row / NP, n = row - s * NP, b = b0 + s;
                float v = 0.f;
                if (n < N && b < B) {
                    size_t ro = (size_t)b * N + n;
                    int c = job.z0 + cl;
                    if (job.type == 0) {
                        if (xsrc != nullptr) v = xsrc[(size_t)b * xsb + n * p.fin + c];
                    } else if (job.type == 1) {
                        v = hprev[ro * H + c];
                    } else if (job.type == 2) {
                        v = hprev[ro * H + c] * ruc_t[ro * H3 + c];
                    } else {
                        v = p.dY[((size_t)t * B * N + ro) * p.Fo + c];
                    }
                }
                if (job.type == 3) Gt[row * GLD + cl] = v; else Zc[row * ZLD + cl] = v;
            }
            __syncthreads();
            // ---- G = diffuse(Z) -------------------------------------------------------------------
            if (job.type != 3) {
                for (int idx = tid; idx < DR * nz; idx += NT) {
                    int row = idx / nz, cl = idx - row * nz;
                    Gt[row * GLD + cl * M] = Zc[row * ZLD + cl];
                }
                const int ntask = DSB * M1 * 4 * nz;
                for (int id = tid; id < ntask; id += NT) {
                    int cl = id % nz;
                    int t1 = id / nz;
                    int q = t1 & 3, sm = t1 >> 2;
                    int s = sm / max(M1, 1), m1 = sm - s * max(M1, 1);
                    const float* zp = Zc + (s * NP) * ZLD + cl;
                    const float* pp = PT + (size_t)sm * NP * NP + 5 * q;
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
                    for (int j = 0; j < N; ++j) {
                        float z = zp[j * ZLD];
                        const float* pj = pp + j * NP;
                        a0 = fmaf(pj[0], z, a0); a1 = fmaf(pj[1], z, a1); a2 = fmaf(pj[2], z, a2);
                        a3 = fmaf(pj[3], z, a3); a4 = fmaf(pj[4], z, a4);
                    }
                    float* gp = Gt + (s * NP + 5 * q) * GLD + cl * M + m1 + 1;
                    gp[0] = a0; gp[GLD] = a1; gp[2 * GLD] = a2; gp[3 * GLD] = a3; gp[4 * GLD] = a4;
                }
            }
            cp_async_wait<0>();
            __syncthreads();
            // ---- acc += G^T dA over the 80 rows ---------------------------------------------------------
#pragma unroll 4
            for (int row = 0; row < DR; ++row) {
                float4 g4 = *reinterpret_cast<const float4*>(Gt + row * GLD + 4 * ky);
                float d[TO];
                load_vec<TO>(d, dAt + row * DLD + ox * TO);
                const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < TO; ++j) acc[i][j] = fmaf(gg[i], d[j], acc[i][j]);
                if (do_db) {
#pragma unroll
                    for (int j = 0; j < TO; ++j) dbacc[j] += d[j];
                }
                if (do_gsum) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) gsum[i] += gg[i];
                }
            }
            __syncthreads();
        }
    }
    // ---- partial tile out ---------------------------------------------------------------------------
    const int ldp = (job.type == 3) ? H : H3;
    const size_t psz = (job.type == 3) ? (size_t)p.Fo * H : (size_t)(p.fin + H) * p.M * H3;
    float* part = p.part + (size_t)split * psz;
    const int kk0 = job.kk0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int kl = 4 * ky + i;
        if (kl < kcols) {
#pragma unroll
            for (int j = 0; j < TO; ++j)
                part[(size_t)(kk0 + kl) * ldp + job.o0 + ox * TO + j] = acc[i][j];
        }
    }
    if (do_db) {
#pragma unroll
        for (int j = 0; j < TO; ++j) p.partb[(size_t)split * H3 + job.o0 + ox * TO + j] = dbacc[j];
    }
    if (do_gsum) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int kl = 4 * ky + i;
            if (kl < kcols) p.partb[(size_t)split * p.Fo + kk0 + kl] = gsum[i];
        }
    }
}

__global__ void __launch_bounds__(NT, 2) dw_kernel(const DwParams p) {
    extern __shared__ __align__(16) float smem[];
    const DwJob job = p.jobs[blockIdx.x];
    switch (job.nco / 16) {
        case 12: dw_run<12>(p, job, smem); break;
        case 8: dw_run<8>(p, job, smem); break;
        case 6: dw_run<6>(p, job, smem); break;
        case 4: dw_run<4>(p, job, smem); break;
        case 2: dw_run<2>(p, job, smem); break;
        default: break;
    }
}

int dw_smem_bytes(int M, int nco_max) {
    int fl = DSB * (M - 1) * NP * NP + DR * ZLD + 4 + DR * GLD + DR * (nco_max + 4);
    return fl * 4;
}

cudaError_t launch_dw(const DwParams& p, int njobs, int nco_max, cudaStream_t st) {
    int smem = dw_smem_bytes(p.M, nco_max);
    cudaError_t e = cudaFuncSetAttribute(dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    dim3 grid(njobs, p.nsplit);
    dw_kernel<<<grid, NT, smem, st>>>(p);
    return cudaGetLastError();
}

// ---- reductions of the partials ---------------------------------------------------------------------
// cell: part [nsplit][CM][3H] -> dWg (CM,2H), dWc (CM,H); partb [nsplit][3H] -> dbg, dbc
__global__ void reduce_cell_kernel(const float* part, const float* partb, int nsplit, int CM, int H,
                                   float* dWg, float* dbg, float* dWc, float* dbc) {
    const int H3 = 3 * H;
    const size_t n = (size_t)CM * H3;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        float s = 0.f;
        for (int k = 0; k < nsplit; ++k) s += part[(size_t)k * n + i];
        int kk = (int)(i / H3), o = (int)(i - (size_t)kk * H3);
        if (o < 2 * H) dWg[(size_t)kk * 2 * H + o] = s; else dWc[(size_t)kk * H + o - 2 * H] = s;
    }
    if (i < (size_t)H3) {
        float s = 0.f;
        for (int k = 0; k < nsplit; ++k) s += partb[(size_t)k * H3 + i];
        if (i < (size_t)2 * H) dbg[i] = s; else dbc[i - 2 * H] = s;
    }
}

__global__ void reduce_flat_kernel(const float* part, int nsplit, size_t n, float* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        float s = 0.f;
        for (int k = 0; k < nsplit; ++k) s += part[(size_t)k * n + i];
        out[i] = s;
    }
}

cudaError_t launch_reduce_cell(const float* part, const float* partb, int nsplit, int CM, int H,
                               float* dWg, float* dbg, float* dWc, float* dbc, cudaStream_t st) {
    size_t n = (size_t)CM * 3 * H;
    reduce_cell_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(part, partb, nsplit, CM, H, dWg, dbg, dWc, dbc);
    return cudaGetLastError();
}
cudaError_t launch_reduce_flat(const float* part, int nsplit, size_t n, float* out, cudaStream_t st) {
    reduce_flat_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(part, nsplit, n, out);
    return cudaGetLastError();
}

// ---- small layout helpers ------------------------------------------------------------------------------
// out (cols, rows_pad) <- in (rows, cols)^T, zero padded to ld_out columns
__global__ void transpose_kernel(const float* in, int rows, int cols, float* out, int ld_out) {
    __shared__ float tile[32][33];
    int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? in[(size_t)r * cols + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        int c = c0 + i, r = r0 + threadIdx.x;       // out row = c, out col = r
        if (c < cols && r < ld_out) out[(size_t)c * ld_out + r] = tile[threadIdx.x][i];
    }
}
cudaError_t launch_transpose(const float* in, int rows, int cols, float* out, int ld_out, cudaStream_t st) {
    dim3 grid((cols + 31) / 32, (ld_out + 31) / 32), block(32, 8);
    transpose_kernel<<<grid, block, 0, st>>>(in, rows, cols, out, ld_out);
    return cudaGetLastError();
}

}  // namespace dcgru
