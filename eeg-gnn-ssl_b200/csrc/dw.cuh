// Parameter blocks of the bulk weight-gradient kernel (dw.cu) and host launchers shared with capi.cu.
#pragma once
#include "common.cuh"

namespace dcgru {

struct DwJob {
    int type;   // 0: x columns (G = diffuse(x), dA cols [0,3H)), 1: gate h columns (h_prev, dA cols [0,2H)),
                // 2: candidate h columns (r*h_prev, dA cols [2H,3H)), 3: decoder Linear (dY^T top)
    int z0;     // first source column (within x, within h, or within Fo)
    int nz;     // number of source columns (nz*M <= 64)
    int kk0;    // first output row of the partial this job owns
    int o0;     // first output column (absolute within 3H, or within H for type 3)
    int nco;    // number of output columns: 32, 64, 96, 128 or 192
};

constexpr int DW_MAXJOBS = 96;

struct DwParams {
    int B, T, N, H, M, nsplit, mode, layer, ncell, fin, Fo;
    const float* P;
    const float* x; long long xs_t, xs_b;
    const float* h0;
    const float* hseq;
    const float* ruc;
    const float* dA;
    const float* targets;
    const float* out;
    unsigned long long teacher_mask;
    const float* dY;
    const float* dropmask;
    float* part;
    float* partb;
    DwJob jobs[DW_MAXJOBS];
};

int dw_smem_bytes(int M, int nco_max);
cudaError_t launch_dw(const DwParams& p, int njobs, int nco_max, cudaStream_t st);
cudaError_t launch_reduce_cell(const float* part, const float* partb, int nsplit, int CM, int H,
                               float* dWg, float* dbg, float* dWc, float* dbc, cudaStream_t st);
cudaError_t launch_reduce_flat(const float* part, int nsplit, size_t n, float* out, cudaStream_t st);
cudaError_t launch_transpose(const float* in, int rows, int cols, float* out, int ld_out, cudaStream_t st);

cudaError_t launch_seq_fwd(const FwdParams& p, int SB, int smem_bytes, cudaStream_t st);
cudaError_t launch_seq_bwd(const BwdParams& p, int SB, int smem_bytes, cudaStream_t st);

cudaError_t launch_graph_poly(int B, int N, int K, int S, const float* const* sup_dev_ptrs_host,
                              const long long* bstride, float* P, cudaStream_t st);
cudaError_t launch_corr_supports(int B, int T, int N, int F, const float* clip, long long sb, long long st_,
                                 float scale, float shift, int top_k, float* adj, float* s0, float* s1,
                                 cudaStream_t st);
int dw_tc_smem_bytes(int M, int nco_max);
size_t dw_tc_pt_floats(int B, int M);
cudaError_t launch_make_pt(const float* P, int B, int M, int N, float* PT, cudaStream_t st);
cudaError_t launch_dw_tc(const DwParams& p, int njobs, int nco_max, cudaStream_t st);
size_t seq_fwd_tc_wimg_bytes(int fin);
bool seq_fwd_tc_supported(int N, int fin, int H, int M, int smem_limit);
cudaError_t launch_seq_fwd_tc(int B, int T, int N, int fin, int act, const float* x, long long xs_t, long long xs_b,
                              const float* h0, const float* P, const float* Wg, const float* bg, const float* Wc,
                              const float* bc, float* wimg, float* hseq, float* ruc, cudaStream_t st);
size_t seq_bwd_tc_wimg_bytes();
bool seq_bwd_tc_supported(int N, int H, int M, int smem_limit);
cudaError_t launch_seq_bwd_tc(int B, int T, int N, int fin, int act, const float* h0, const float* hseq, const float* ruc,
                              const float* P, const float* Wg, const float* Wc, const float* d_hseq, const float* d_hlast,
                              float* wimg, float* dh0, float* dA, cudaStream_t st);
cudaError_t launch_dx_tc(int B, int T, int N, const float* P, const float* Wg, const float* Wc, const float* dA,
                         float* wimg, float* dx, cudaStream_t st);
cudaError_t launch_tc_selftest(const float* A, const float* B, float* C, int N, int K, cudaStream_t st);
}  // namespace dcgru
