/*
 * dcgru_b200 -- C ABI of the B200-native DCGRU forward/backward path.
 *
 * The reference (tsy935/eeg-gnn-ssl) has no FFI for this path: its boundary is the Python
 * class surface of model/cell.py and model/model.py.  These entry points are what
 * a binding for that surface has to call; every one cites the reference code it replaces.
 * All tensors are dense fp32, row-major, device pointers owned by the caller; the library
 * allocates no device memory, keeps no thread-local state except the last error string,
 * launches on the caller's stream and never synchronises it.  One exception, only when the
 * bias gradient is not fused into the weight-gradient GEMM (DCGRU_FUSE_DB=0): the backward
 * calls fork that small pass onto a library-owned non-blocking stream (one stream + two events
 * per host thread and device, created on first use) and join it before they return -- plain
 * event dependencies, legal under CUDA-graph capture.
 *
 * Return value: 0 on success, non-zero on error (message via dcgru_last_error()).
 *
 * Notation: B batch, T steps, N nodes (<= 20; the reference uses 19), Fin/Fo feature dims,
 * H rnn units, K max_diffusion_step, S number of supports, M = S*K + 1, C = Fin + H.
 */
#ifndef DCGRU_B200_H
#define DCGRU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCGRU_MAX_LAYERS 4
#define DCGRU_ACT_TANH 0
#define DCGRU_ACT_RELU 1

/* One DCGRU cell's hyper-parameters: the constructor arguments of
 * DCGRUCell (model/cell.py:126-175) that change the arithmetic. */
typedef struct dcgru_cell_desc {
    int32_t num_nodes;          /* N   (model/cell.py:147) */
    int32_t input_dim;          /* Fin (model/cell.py:162) */
    int32_t hid_dim;            /* H   (model/cell.py:148) */
    int32_t max_diffusion_step; /* K   (model/cell.py:149) */
    int32_t num_supports;       /* S   (model/cell.py:151-158) */
    int32_t activation;         /* DCGRU_ACT_* (model/cell.py:146: anything but 'tanh' is relu) */
} dcgru_cell_desc;

/* One cell's parameters: dconv_gate.{weight,biases}, dconv_candidate.{weight,biases}
 * (model/cell.py:40-48,160-175).  weight rows are indexed c*M + m (model/cell.py:98-114). */
typedef struct dcgru_cell_params {
    const float *Wg; /* (C*M, 2H) */
    const float *bg; /* (2H)      */
    const float *Wc; /* (C*M, H)  */
    const float *bc; /* (H)       */
} dcgru_cell_params;

typedef struct dcgru_cell_grads {
    float *dWg, *dbg, *dWc, *dbc; /* same shapes, overwritten */
} dcgru_cell_grads;

int dcgru_version(void);
const char *dcgru_last_error(void);

/* Diagnostics (no reference counterpart; used by bench.py's roofline leg): when enabled, each kernel
 * launched by this library is bracketed by cudaEvents on the launching stream.  collect() waits for
 * them and writes "name count total_ms\n" lines into buf.  enable(0/1) also clears the records.    */
int dcgru_timing_enable(int on);
int dcgru_timing_collect(char *buf, size_t cap);

/* Hardware self-test of the tcgen05/TMEM plumbing the tensor-core kernels are built from:
 * C[128 x N] = A[128 x K] * B[N x K]^T in 3xTF32 (fp32-level accuracy); A, B, C device pointers,
 * row-major, K % 32 == 0, N in {64, 128, 192, 256}.                                              */
int dcgru_tc_selftest(const float *A, const float *B, float *C, int32_t N, int32_t K, void *stream);
/* Layout probe: one kind::tf32 MMA D[128 x N] = A[128 x 8] * B[N x 8]^T over raw shared-memory images of the
 * operands (device pointers, copied verbatim into shared memory) with the given descriptor byte offsets,
 * major-ness and layout type (descriptor bits 61-63: 0 none, 1 128B swizzle with 32B base, 2 128B, 4 64B, 6 32B).  tests/test_gpu_tc.py uses it to pin the operand layouts the kernels rely on.              */
int dcgru_tc_probe(const float *a_img, int32_t a_bytes, const float *b_img, int32_t b_bytes,
                   uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo,
                   int32_t a_mn_major, int32_t b_mn_major, int32_t a_layout_type, int32_t b_layout_type,
                   float *D, int32_t N, void *stream);

/* ---- graph -> diffusion polynomials -------------------------------------------------------
 * Replaces the hop recurrence of DiffusionGraphConv.forward (model/cell.py:76-93): since
 * diffusion is linear, term_m = P_m Z with per-sample N x N matrices P_m (the
 * carried-x0 quirk across supports is reproduced).  supports[s] points at (B,N,N) with
 * batch stride support_bstride[s] elements (0 broadcasts one (N,N) matrix).
 * P out: (B, M-1, N, N).                                                                     */
int dcgru_graph_poly(int32_t batch, int32_t num_nodes, int32_t max_diffusion_step,
                     int32_t num_supports, const float *const *supports,
                     const int64_t *support_bstride, float *P, void *stream);

/* ---- correlation graph -> supports --------------------------------------------------------
 * Replaces SeizureDataset._get_indiv_graphs + _compute_supports('dual_random_walk')
 * (data/dataloader_detection.py:258-307,335-354; data/data_utils.py:174-222;
 * utils.py:220-230) for a whole batch on the device.
 * clip: (B, T, N, F) with element strides (sb, st); value used = clip*scale + shift
 * (scale=std, shift=mean undoes the scalar StandardScaler, utils.py:402-403; 1,0 for raw).
 * adj out (B,N,N) (may be NULL), support0/support1 out (B,N,N).                              */
int dcgru_corr_supports(int32_t batch, int32_t seq_len, int32_t num_nodes, int32_t feat,
                        const float *clip, int64_t stride_b, int64_t stride_t,
                        float scale, float shift, int32_t top_k,
                        float *adj, float *support0, float *support1, void *stream);

/* ---- input features ---------------------------------------------------------
 * What the reference's DataLoader workers do per clip before the model sees it:
 *   1-second windows of 200 samples -> fft(n=200) -> first 100 bins -> log|.| with 0 -> 1e-8
 *   (data/dataloader_detection.py:58-72 computeSliceMatrix(is_fft=True), data/data_utils.py:13-34
 *   computeFFT), then _random_reflect / _random_scale (data/dataloader_detection.py:233-256:
 *   swap channel pairs; += log(scale_factor)), then StandardScaler.transform (utils.py:402-403).
 * signal: resampled EEG, element (b, n, s) at signal + b*stride_b + n*stride_n + s, seq_len*200
 *         contiguous samples per channel (the h5 layout, channels x samples); 16-byte aligned,
 *         strides multiples of 4 samples (the windows are fetched with 800-byte bulk copies)
 * dest_channel: (B, N) int32 or NULL -- output channel that source channel n lands in (for the
 *         reference's pair swaps, an involution, this is the swapped index; NULL = no reflection)
 * log_scale: (B) or NULL -- log(scale_factor) of _random_scale
 * mean/std: stat_len values (1: scalar scaler, N: per-channel (1,N,1) scaler, 0: no scaling)
 * raw out (B, T, N, 100) or NULL: features before augmentation/scaling -- what
 *         _get_indiv_graphs is given (data/dataloader_detection.py:395: eeg_clip, not curr_feature),
 *         i.e. the input of dcgru_corr_supports
 * x out   (B, T, N, 100) or NULL: the model input                                              */
int dcgru_fft_features(int32_t batch, int32_t num_nodes, int32_t seq_len, int32_t window,
                       const float *signal, int64_t stride_b, int64_t stride_n,
                       const int32_t *dest_channel, const float *log_scale,
                       const float *mean, const float *std, int32_t stat_len,
                       float *raw, float *x, void *stream);

/* ---- encoder layer -------------------------------------------------------------------------
 * One launch = one RNN layer over all T steps (the inner loop of DCRNNEncoder.forward,
 * model/model.py:93-96, around DCGRUCell.forward, model/cell.py:182-210).
 * x:     input sequence, element (t,b,:) at x + t*x_stride_t + b*x_stride_b, N*Fin contiguous
 * h0:    (B, N*H)
 * P:     (B, M-1, N, N) from dcgru_graph_poly
 * h_seq: (T, B, N*H) out -- every step's hidden state (the layer's output sequence)
 * ruc:   (T, B, N, 3H) out -- r | u | c per node, saved for backward (NULL: inference)
 * gsave: optional "operand image" saved for backward (NULL: not saved).  When the tensor-core
 *        kernels serve this configuration, dcgru_encoder_layer_gsave_bytes() is non-zero and the
 *        forward kernel can leave its diffused GEMM operands [x | h | r*h] (hi/lo split, already
 *        in tensor-core tile order) in this caller-owned buffer; handing the same buffer to
 *        dcgru_encoder_layer_bwd turns the weight gradient (the G^T dA of MmBackward for
 *        model/cell.py:116) into a pure GEMM over saved operands instead of recomputing the
 *        diffusion.  Costs gsave_bytes of HBM per layer (3.7 GB at B=512, T=60, Fin=100).
 * workspace: scratch for the pre-tiled weight image of the tensor-core kernel (may be NULL:
 *        the fp32 FMA kernel is used then)                                                    */
size_t dcgru_encoder_layer_gsave_bytes(const dcgru_cell_desc *d, int32_t batch, int32_t seq_len);
size_t dcgru_encoder_layer_fwd_workspace(const dcgru_cell_desc *d, int32_t batch, int32_t seq_len);
int dcgru_encoder_layer_fwd(const dcgru_cell_desc *d, int32_t batch, int32_t seq_len,
                            const float *x, int64_t x_stride_t, int64_t x_stride_b,
                            const float *h0, const float *P, const dcgru_cell_params *w,
                            float *h_seq, float *ruc, void *gsave, size_t gsave_bytes,
                            void *workspace, size_t workspace_bytes, void *stream);

size_t dcgru_encoder_layer_bwd_workspace(const dcgru_cell_desc *d, int32_t batch,
                                         int32_t seq_len);
/* Diagnostics for the tests: byte offsets inside the backward workspace of {row-major dA (T,B,N,3H),
 * dA operand image, per-CTA partials of the weight-gradient GEMM} (0 = not used by this configuration). */
int dcgru_debug_encoder_bwd_offsets(const dcgru_cell_desc *d, int32_t batch, int32_t seq_len,
                                    size_t *out3);
/* the tile / set / CTA plan of the weight-gradient GEMM as integers (pure host logic; see capi.cu for the layout) */
int dcgru_debug_dwmm_plan(int32_t input_dim, int32_t batch, int32_t seq_len, int32_t num_sms,
                          int32_t *out, int32_t cap);
/* clock64 stamps of the weight-gradient GEMM's loader / MMA-issuer threads (recorded when DCGRU_DBG & 8) */
int dcgru_debug_dwmm_stamps(long long *out, int32_t n);

/* Backward of the above (what autograd derives for the reference).
 * d_hseq:  (T,B,N*H) upstream gradient of h_seq (NULL = zeros)
 * d_hlast: (B,N*H)  extra upstream gradient of h_seq[T-1] (the layer's output_hidden; NULL)
 * dx:      (T,B,N*Fin) out, or NULL when the input needs no gradient (layer 0)
 * dh0:     (B,N*H) out
 * gsave:   the operand image the forward call filled, or NULL (the diffusion is recomputed)   */
int dcgru_encoder_layer_bwd(const dcgru_cell_desc *d, int32_t batch, int32_t seq_len,
                            const float *x, int64_t x_stride_t, int64_t x_stride_b,
                            const float *h0, const float *P, const dcgru_cell_params *w,
                            const float *h_seq, const float *ruc,
                            const float *d_hseq, const float *d_hlast,
                            float *dx, float *dh0, const dcgru_cell_grads *g,
                            const void *gsave, size_t gsave_bytes,
                            void *workspace, size_t workspace_bytes, void *stream);

/* ---- fused classification head -------------------------------------------------
 * Replaces the tail of DCRNNModel_classification.forward (model/model.py:257-270):
 *   last = utils.last_relevant_pytorch(h_seq, seq_lengths)   (utils.py:346-357)  -> h_seq[sel_t[b], b]
 *   logits = max over nodes of fc(relu(dropout(last)))                            -> (B, num_classes)
 * h_seq:      (T,B,N*H) top-layer sequence of the encoder
 * sel_t:      (B) int32 device array, sel_t[b] = seq_lengths[b]-1 (clamped to [0,T-1]); NULL = T-1 for every sample
 * drop_mask:  (B,N,H) multiplicative dropout mask (0 or 1/(1-p), as nn.Dropout draws it) or NULL (eval / p = 0)
 * fc_w (num_classes,H), fc_b (num_classes): nn.Linear(rnn_units, num_classes)
 * logits out (B,num_classes); argmax_node out (B,num_classes) int32: the node the max came from (saved for backward)
 *
 * Backward: d_logits (B,num_classes) -> d_fc_w, d_fc_b (overwritten; batch sums in fixed order) and d_hsel (B,N*H),
 * the gradient of h_seq in SPARSE form: sample b's slab belongs to step sel_t[b]; every other step's gradient is zero.
 * dcgru_encoder_layer_bwd_sel consumes it directly, so the dense (T,B,N*H) gradient that autograd derives for the
 * reference's gather (utils.py:354) is never written or read.                                                      */
int dcgru_cls_head_fwd(int32_t batch, int32_t seq_len, int32_t num_nodes, int32_t hid_dim, int32_t num_classes,
                       const float *h_seq, const int32_t *sel_t, const float *drop_mask,
                       const float *fc_w, const float *fc_b, float *logits, int32_t *argmax_node, void *stream);
size_t dcgru_cls_head_bwd_workspace(int32_t batch, int32_t hid_dim, int32_t num_classes);
int dcgru_cls_head_bwd(int32_t batch, int32_t seq_len, int32_t num_nodes, int32_t hid_dim, int32_t num_classes,
                       const float *h_seq, const int32_t *sel_t, const float *drop_mask, const float *fc_w,
                       const int32_t *argmax_node, const float *d_logits, float *d_hsel,
                       float *d_fc_w, float *d_fc_b, void *workspace, size_t workspace_bytes, void *stream);

/* dcgru_encoder_layer_bwd with the upstream gradient of h_seq in the sparse form above (d_hsel, sel_t) instead of
 * the dense d_hseq; everything else as dcgru_encoder_layer_bwd.  The tensor-core BPTT kernel injects the slab at step
 * sel_t[b]; configurations served by the other kernels expand it into a dense buffer inside the (larger) workspace. */
size_t dcgru_encoder_layer_bwd_sel_workspace(const dcgru_cell_desc *d, int32_t batch, int32_t seq_len);
int dcgru_encoder_layer_bwd_sel(const dcgru_cell_desc *d, int32_t batch, int32_t seq_len,
                                const float *x, int64_t x_stride_t, int64_t x_stride_b,
                                const float *h0, const float *P, const dcgru_cell_params *w,
                                const float *h_seq, const float *ruc,
                                const float *d_hsel, const int32_t *sel_t, const float *d_hlast,
                                float *dx, float *dh0, const dcgru_cell_grads *g,
                                const void *gsave, size_t gsave_bytes,
                                void *workspace, size_t workspace_bytes, void *stream);

/* ---- fused optimiser step ---------------------------------------------------
 * The tail of the reference's training step on flat fp32 buffers of n elements (train.py:273-275):
 * torch.nn.utils.clip_grad_norm_(max_grad_norm) followed by torch.optim.Adam(lr, betas, eps,
 * weight_decay).step() (L2 form: grad += weight_decay * param), in two launches.
 * lr and step are DEVICE scalars: step is incremented by the call (bias correction uses the new value),
 * lr may be changed by a scheduler between calls even when the call sits inside a CUDA graph.
 * grads is overwritten with the clipped gradient, total_norm (device, may be NULL) receives the
 * pre-clip global norm (the return value of clip_grad_norm_).  max_grad_norm <= 0 disables clipping.
 * grad_scale multiplies the gradient first (1 = none; 1/world_size turns the all-reduced SUM into the
 * data-parallel average without a separate pass); the norm and the clip are those of the scaled gradient. */
size_t dcgru_clip_adam_workspace(size_t n);
int dcgru_clip_adam_step(float *params, float *grads, float *exp_avg, float *exp_avg_sq, size_t n,
                         const float *lr, int32_t *step, float beta1, float beta2, float eps,
                         float weight_decay, float max_grad_norm, float grad_scale, float *total_norm,
                         void *workspace, size_t workspace_bytes, void *stream);

/* ---- decoder -------------------------------------------------------------------------------
 * One launch = DCGRUDecoder.forward (model/model.py:149-204): To autoregressive steps x L
 * cells + Linear(H->Fo) per node.  d describes cell 0 (input_dim = Fo); cells >= 1 have
 * input_dim = H.  w[l] may alias for l >= 1 (weight tying, model/model.py:126,142-143).
 * targets:      (To,B,N*Fo) teacher inputs, or NULL
 * teacher_mask: bit t set => input of step t+1 is targets[t] (the outcome of the reference's
 *               per-step random.random() draw, model/model.py:198-202); To <= 64
 * h0:           (L,B,N*H) encoder context
 * proj_w:       (Fo,H), proj_b: (Fo)   (nn.Linear, model/model.py:146)
 * drop_mask:    (To,B,N,H) multiplicative dropout mask or NULL (model/model.py:192)
 * out:          (To,B,N*Fo)
 * h_all:        (To,L,B,N*H) out, ruc: (To,L,B,N,3H) out (saved for backward)                 */
size_t dcgru_decoder_fwd_workspace(const dcgru_cell_desc *d, int32_t num_layers,
                                   int32_t batch, int32_t seq_len);
int dcgru_decoder_fwd(const dcgru_cell_desc *d, int32_t num_layers, int32_t batch,
                      int32_t seq_len, const float *targets, uint64_t teacher_mask,
                      const float *h0, const float *P, const dcgru_cell_params *w,
                      const float *proj_w, const float *proj_b, const float *drop_mask,
                      float *out, float *h_all, float *ruc,
                      void *workspace, size_t workspace_bytes, void *stream);

size_t dcgru_decoder_bwd_workspace(const dcgru_cell_desc *d, int32_t num_layers,
                                   int32_t batch, int32_t seq_len);
/* g[l] for tied layers (w[l] aliasing w[1]) must alias too; gradients of tied cells are summed.
 * d_out: (To,B,N*Fo) upstream; dh0: (L,B,N*H) out; dproj_w (Fo,H), dproj_b (Fo) out.         */
int dcgru_decoder_bwd(const dcgru_cell_desc *d, int32_t num_layers, int32_t batch,
                      int32_t seq_len, const float *targets, uint64_t teacher_mask,
                      const float *h0, const float *P, const dcgru_cell_params *w,
                      const float *proj_w, const float *drop_mask,
                      const float *out, const float *h_all, const float *ruc,
                      const float *d_out, float *dh0, const dcgru_cell_grads *g,
                      float *dproj_w, float *dproj_b,
                      void *workspace, size_t workspace_bytes, void *stream);

/* The same two calls with a caller-owned operand image saved for backward (the decoder's counterpart of the encoder's
 * gsave): when dcgru_decoder_gsave_bytes() is non-zero the tensor-core kernels serve this configuration (H = 64) and the
 * forward pass can leave the diffused GEMM operands of every (step, layer) cell in gsave -- cell 0: T slabs, the tied
 * upper cell: T*(L-1) slabs -- so that the backward pass computes the cells' weight gradients as GEMMs over saved
 * operands (one per distinct cell: the tied cell's gradient sums over layers and steps inside the GEMM's K range,
 * model/model.py:126,142-143) instead of recomputing the diffusion.  gsave = NULL: exactly dcgru_decoder_fwd / _bwd.   */
size_t dcgru_decoder_gsave_bytes(const dcgru_cell_desc *d, int32_t num_layers, int32_t batch, int32_t seq_len);
int dcgru_decoder_fwd_saved(const dcgru_cell_desc *d, int32_t num_layers, int32_t batch,
                            int32_t seq_len, const float *targets, uint64_t teacher_mask,
                            const float *h0, const float *P, const dcgru_cell_params *w,
                            const float *proj_w, const float *proj_b, const float *drop_mask,
                            float *out, float *h_all, float *ruc, void *gsave, size_t gsave_bytes,
                            void *workspace, size_t workspace_bytes, void *stream);
int dcgru_decoder_bwd_saved(const dcgru_cell_desc *d, int32_t num_layers, int32_t batch,
                            int32_t seq_len, const float *targets, uint64_t teacher_mask,
                            const float *h0, const float *P, const dcgru_cell_params *w,
                            const float *proj_w, const float *drop_mask,
                            const float *out, const float *h_all, const float *ruc,
                            const float *d_out, float *dh0, const dcgru_cell_grads *g,
                            float *dproj_w, float *dproj_b, const void *gsave, size_t gsave_bytes,
                            void *workspace, size_t workspace_bytes, void *stream);

/* ---- diagnostics of the second-generation (2xFP16) kernels: one stage at a time ----------------
 * mode 0: out (T,B,N,3H) = sum_m (P_m x_t) [Wg_x | Wc_x]_m + bias   (the hoisted x-part of model/cell.py:73-116);
 *         img (may be NULL): fp16 operand image [tile*T+t][hi|lo][96][img_cols], columns from img_col0
 * mode 1: out (T,B,N,fin) = sum_m P_m^T (src_t W_x,m^T), src = dA (T,B,N,3H)   (its BPTT)                   */
int dcgru_debug_rnn_fwd_stamps(long long *out, int32_t n);   /* clock64 timeline of CTA 0 (DCGRU_DBG & 16) */
int dcgru_debug_rnn_bwd_stamps(long long *out, int32_t n);   /* same for the BPTT kernel (DCGRU_DBG & 32) */
size_t dcgru_debug_bulk_dp_workspace(int32_t mode, int32_t input_dim, int32_t hid_dim, int32_t num_matrices);
int dcgru_debug_bulk_dp(int32_t mode, int32_t batch, int32_t seq_len, int32_t num_nodes, int32_t input_dim,
                        int32_t hid_dim, int32_t num_matrices, const float *src, const float *P,
                        const float *Wg, const float *Wc, const float *bias, float *out, void *img,
                        int32_t img_cols, int32_t img_col0, void *workspace, size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DCGRU_B200_H */
