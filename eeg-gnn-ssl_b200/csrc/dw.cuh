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

// ---- dw_mm.cu: weight gradient as a GEMM over the saved operand images ----------------------------------
constexpr int DWMM_MAXTILE = 4;
constexpr int DWMM_MAXSET = 4;
struct DwmmTile {
    int kg0, nkg;      // K groups (4 kk each) of the G image this 128-row tile covers
    int og0, ncol;     // first column quad of the dA image and number of columns (64, 128 or 192)
    int tcol;          // TMEM column base inside the set
    int soff;          // byte offset of the tile's hi part inside a pipeline stage (lo part: + 8 KB)
};
struct DwmmSet {
    int ntile, ncoltot, cta0, ncta, ogmin, ogcnt;
    int nbox, boxgrp[2];   // TMA boxes of the G image per K block: a pair of adjacent tiles, hi and lo, in one load
    DwmmTile tile[DWMM_MAXTILE];
};
struct DwmmParams {
    const uint8_t* G;
    const uint8_t* DA;
    float* part;       // [cta][512 columns][128 rows]
    long nkb;          // K blocks (8 rows each) = slabs * 16
    int KGT, nset, ncta, dbg;
    DwmmSet set[DWMM_MAXSET];
};
cudaError_t dwmm_read_dbg(long long* out, int n);
bool dwmm_plan(int fin, int H, int M, long nslab, int nsms, DwmmParams* out);
size_t dwmm_part_floats(int nsms);
size_t colsum_part_floats(int H);
int dw_mm_smem_bytes();
cudaError_t launch_dw_mm(const DwmmParams& p, int fin, int H, int M, float* dWg, float* dWc, cudaStream_t st);
cudaError_t launch_colsum(const float* dA, long rows, int H, float* partial, float* dbg, float* dbc, cudaStream_t st);
int seq_fwd_tc_kgt(int fin);
int seq_fwd_tc_kkp(int fin);
int seq_tc_nslab(int B, int T);
size_t seq_fwd_tc_gsave_bytes(int B, int T, int fin);
size_t seq_bwd_tc_daimg_bytes(int B, int T);

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
cudaError_t launch_fft_features(int B, int N, int T, const float* signal, long long sig_sb, long long sig_sn, const int* perm,
                                const float* log_scale, const float* mean, const float* stdv, int stat_len, float* raw, float* x,
                                int nsms, cudaStream_t st);
int dw_tc_smem_bytes(int M, int nco_max);
size_t dw_tc_pt_floats(int B, int M);
cudaError_t launch_make_pt(const float* P, int B, int M, int N, float* PT, cudaStream_t st);
cudaError_t launch_dw_tc(const DwParams& p, int njobs, int nco_max, cudaStream_t st);
size_t seq_fwd_tc_wimg_bytes(int fin);
bool seq_fwd_tc_supported(int N, int fin, int H, int M, int smem_limit);
cudaError_t launch_seq_fwd_tc(int B, int T, int N, int fin, int act, const float* x, long long xs_t, long long xs_b,
                              const float* h0, const float* P, const float* Wg, const float* bg, const float* Wc,
                              const float* bc, float* wimg, float* hseq, float* ruc, void* gsave, cudaStream_t st);
size_t seq_bwd_tc_wimg_bytes();
bool seq_bwd_tc_supported(int N, int H, int M, int smem_limit);
cudaError_t launch_seq_bwd_tc(int B, int T, int N, int fin, int act, const float* h0, const float* hseq, const float* ruc,
                              const float* P, const float* Wg, const float* Wc, const float* d_hseq, const float* d_hlast,
                              float* wimg, float* dh0, float* dA, void* daimg, cudaStream_t st);
cudaError_t launch_dx_tc(int B, int T, int N, const float* P, const float* Wg, const float* Wc, const float* dA,
                         float* wimg, float* dx, cudaStream_t st);
int clip_adam_npart(size_t n);
cudaError_t launch_clip_adam(float* p, float* g, float* m, float* v, size_t n, const float* lr_dev, int* step,
                             float beta1, float beta2, float eps, float wd, float max_norm, float gscale,
                             double* partial, float* norm_out, cudaStream_t st);
cudaError_t launch_tc_selftest(const float* A, const float* B, float* C, int N, int K, cudaStream_t st);
cudaError_t launch_tc_probe(const float* Aimg, int a_bytes, const float* Bimg, int b_bytes, uint32_t a_lbo, uint32_t a_sbo,
                            uint32_t b_lbo, uint32_t b_sbo, uint32_t idesc, uint32_t a_type, uint32_t b_type, float* D, int N,
                            cudaStream_t st);
// ---- second-generation tensor-core kernels (2xFP16, f16_common.cuh) ----------------------------------------------------
int g16_nq(int cin, int M);
int g16_kxp(int fin, int M);
int g16_kkp(int fin, int H, int M);
int g16_ntile(int B);
size_t g16_image_bytes(int B, int T, int cols);
size_t bulk_wimg_bytes(int cin, int M, int nout);
bool bulk_dp_supported(int N, int Cin, int M, int Nout, bool src16, int smem_limit);
cudaError_t launch_pack_w16(const float* Wg, const float* Wc, int fin, int H, int M, int mode, int nrows, int nq,
                            void* img, cudaStream_t st);
// optional extras of launch_bulk_dp (decoder orchestration): real output columns (< Nout = padded MMA N), position of this
// launch's T steps inside a larger dumped / source image (slab = tile * X_T + X_t0 + t), scale applied to the fp32 source
struct BulkExtra { int nout_valid = 0; int img_T = 0, img_t0 = 0; int src_T = 0, src_t0 = 0; const float* in_scale_ptr = nullptr; };
cudaError_t launch_bulk_dp(int B, int T, int N, int Cin, int M, int Nout, int transposeP, const float* src, long long ss_t,
                           long long ss_b, const void* src16, const float* P, const void* wimg, const float* bias, float* out, long long os_t,
                           long long os_b, int out_ld, float out_scale, const float* scale_ptr, void* img, int img_cols,
                           int img_col0, int nsms, int smem_limit, cudaStream_t st, const BulkExtra* ex = nullptr);

size_t rnn_fwd_wimg_bytes(int M);
cudaError_t rnn_fwd_read_dbg(long long* out, int n);
bool rnn_fwd_supported(int N, int H, int M, int smem_limit);
// Wg == nullptr: wimg already holds the packed weight planes (rnn_fwd_pack_weights); img_T > 0: this launch's steps are slabs
// tile * img_T + img_t0 + t of a larger image
cudaError_t rnn_fwd_pack_weights(const float* Wg, const float* Wc, int fin, int M, void* wimg, cudaStream_t st);
cudaError_t launch_rnn_fwd(int B, int T, int N, int fin, int M, int act, const float* xp, const float* h0, const float* P,
                           const float* Wg, const float* Wc, void* wimg, float* hseq, float* ruc, void* img, int img_cols,
                           int img_col0, cudaStream_t st, int img_T = 0, int img_t0 = 0);

size_t rnn_bwd_wimg_bytes(int M);
cudaError_t rnn_bwd_read_dbg(long long* out, int n);
bool rnn_bwd_supported(int N, int H, int M, int smem_limit);
cudaError_t launch_grad_scale(const float* a, size_t na, const float* b, size_t nb, const float* c, size_t nc, unsigned* scratch,
                              float* scale, cudaStream_t st);
// d_hsel (B,N*H) / sel_t (B): sparse upstream gradient -- sample b's slab belongs to step sel_t[b] (head.cu); may be nullptr
cudaError_t launch_rnn_bwd(int B, int T, int N, int fin, int M, int act, const float* h0, const float* hseq, const float* ruc,
                           const float* P, const float* Wg, const float* Wc, const float* d_hseq, const float* d_hlast,
                           const float* d_hsel, const int* sel_t, void* wimg, const float* scale_ptr, float* dh0, void* daimg,
                           cudaStream_t st, int img_T = 0, int img_t0 = 0);
cudaError_t rnn_bwd_pack_weights(const float* Wg, const float* Wc, int fin, int M, void* wimg, cudaStream_t st);
cudaError_t launch_img_to_rows(const void* img, int B, int T, int N, int cols, const float* scale_ptr, float* out, int group,
                               int stride, int off, cudaStream_t st);

size_t dw_mm16_part_floats(int nsms);
size_t colsum16_part_floats(int H);
int dw_mm16_smem_bytes();
cudaError_t launch_dw_mm16(int fin, int H, int M, int B, int T, const void* G, const void* DA, float* part, const float* scale_ptr,
                           int nsms, float* dWg, float* dWc, cudaStream_t st, float* dbpart = nullptr, float* dbg = nullptr,
                           float* dbc = nullptr);
cudaError_t launch_colsum16(const void* daimg, int B, int T, int H, float* partial, const float* scale_ptr, float* dbg, float* dbc,
                            cudaStream_t st);


// ---- fused classification head (head.cu) --------------------------------------------------------------------------------
bool cls_head_supported(int N, int H, int C);
cudaError_t launch_cls_head_fwd(int B, int T, int N, int H, int C, const float* hseq, const int* sel_t, const float* drop,
                                const float* W, const float* bias, float* logits, int* arg, cudaStream_t st);
cudaError_t launch_cls_head_bwd(int B, int T, int N, int H, int C, const float* hseq, const int* sel_t, const float* drop,
                                const float* W, const int* arg, const float* dlogits, float* d_hsel, float* dW, float* db,
                                float* dwpart, cudaStream_t st);
cudaError_t launch_ew_add_mul(const float* a, const float* b, const float* m, float* out, size_t n, cudaStream_t st);
cudaError_t launch_scatter_sel(int B, int T, int NH, const float* d_hsel, const int* sel_t, float* dense, cudaStream_t st);

}  // namespace dcgru
