/*
 * metaasr_b200 -- C ABI of the B200-native meta-training hot path.
 *
 * The reference (sunprinceS/MetaASR-CrossAccent) is pure Python on stock PyTorch and has no
 * FFI of its own; every entry point below names the reference call site (file:line relative
 * to the reference root) whose arithmetic it replaces.  INTEGRATION.md shows the ctypes stub a
 * maintainer adds on the reference side.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only; no torch / C++ types cross this boundary.
 *   - every function returns 0 on success or a negative MASR_E_* code; masr_last_error()
 *     returns a thread-local message for the last failure.
 *   - all data pointers are DEVICE pointers owned by the caller; `stream` is a cudaStream_t
 *     passed as void*; calls are asynchronous on that stream.
 *   - dtype codes: MASR_F32 = 0, MASR_BF16 = 1 ("act" tensors use the given dtype; statistics,
 *     losses, logits of the CE kernel and all weight gradients are always fp32).
 *   - activations are row-major, batch-first: [rows, features]; conv tensors are NHWC
 *     (H = time, W = frequency).
 */
#ifndef METAASR_B200_H_
#define METAASR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MASR_OK 0
#define MASR_E_INVALID (-1)   /* bad argument / unsupported shape */
#define MASR_E_CUDA (-2)      /* CUDA runtime or driver error     */
#define MASR_E_NOGPU (-3)     /* no sm_100 device                 */

#define MASR_F32 0
#define MASR_BF16 1

#define MASR_ABI_VERSION 3

int masr_abi_version(void);
const char* masr_last_error(void);
/* Fails (MASR_E_NOGPU) unless `device` is an sm_100 GPU: there is no CPU fallback. */
int masr_init(int device);

/* ------------------------------------------------------------------ kernel 1: CTC
 * Replaces F.log_softmax + nn.CTCLoss(blank, reduction='mean', zero_infinity) forward+backward,
 * src/blstm_trainer.py:22,65-70.  One CTA per utterance.
 *   acts        [T, B, C] fp32 (act_is_logprob=0: un-normalised logits, log-softmax fused;
 *               act_is_logprob=1: log-probabilities as nn.CTCLoss receives them)
 *   targets     int64, concatenated; tgt_offsets [B] int64 start of each utterance's labels
 *   in_lens     [B] int64, tgt_lens [B] int64
 *   nll         [B] fp32 out (0 where infeasible and zero_infinity)
 *   loss        [1] fp32 out = mean_b(nll_b / max(tgt_len_b, 1))      (may be NULL)
 *   grad        [T, B, C] fp32 out = d loss / d acts * grad_scale; frames >= in_len get 0
 *               (may be NULL for forward only)
 */
int masr_ctc_fwd_bwd(const float* acts, int T, int B, int C, int act_is_logprob,
                     const int64_t* targets, const int64_t* tgt_offsets,
                     const int64_t* in_lens, const int64_t* tgt_lens, int max_tgt_len,
                     int blank, int zero_infinity, float grad_scale,
                     float* nll, float* loss, float* grad,
                     void* workspace, size_t workspace_bytes, void* stream);
/* Same, with explicit element strides of the activation / gradient tensors: acts(t, b, c) = acts[t*st_t + b*st_b + c].
 * (st_t, st_b) = (B*C, C) is nn.CTCLoss's [T, B, C]; (C, T*C) reads the batch-first [B, T, C] output of a CTC head
 * GEMM in place (joint CTC / attention training: the `ctc_weight` extension, no transpose pass). */
int masr_ctc_fwd_bwd_ex(const float* acts, int T, int B, int C, int64_t st_t, int64_t st_b, int act_is_logprob,
                        const int64_t* targets, const int64_t* tgt_offsets,
                        const int64_t* in_lens, const int64_t* tgt_lens, int max_tgt_len,
                        int blank, int zero_infinity, float grad_scale,
                        float* nll, float* loss, float* grad,
                        void* workspace, size_t workspace_bytes, void* stream);
/* Profiling hook: record SM-clock timestamps of CTA 0 at the phase boundaries of the following CTC launches
 * (start, setup done, emissions done, recursions done, -, end) and read them back (out6: 6 x int64). */
int masr_ctc_debug_enable(int on);
int masr_ctc_debug_read(long long* out6);
/* Bytes of global workspace masr_ctc_fwd_bwd needs for this shape (0: tables fit in shared memory). */
size_t masr_ctc_workspace_bytes(int T, int B, int C, int max_tgt_len);

/* ------------------------------------------------------------------ kernel 2: GEMM family
 * C[M,N] = op(sum_k A(m,k) * B(n,k) + bias[n]),  A(m,k) = A[m*sam + k*sak], B(n,k) = B[n*sbn + k*sbk].
 * Covers nn.Linear forward (torch/nn/functional.py linear, called from
 * mono_transformer_torch.py:62,66,120,206 and nn.Transformer*Layer), its dgrad and wgrad.
 *   flags: bit0 relu, bit1 accumulate into C (C += ...), bit2 split-K with fp32 atomics
 *          (C must be fp32 and pre-initialised)
 *   path : 0 = fp32-accumulate SIMT kernel (any dtype / stride),
 *          1 = tcgen05/TMEM/TMA kernel (bf16 operands; see masr_umma_* for constraints)
 */
#define MASR_GEMM_RELU 1
#define MASR_GEMM_ACCUM 2
#define MASR_GEMM_SPLITK 4
int masr_gemm(const void* A, int a_dtype, int64_t sam, int64_t sak,
              const void* B, int b_dtype, int64_t sbn, int64_t sbk,
              void* C, int c_dtype, int64_t ldc, const float* bias,
              int M, int N, int K, int flags, int splitk, void* stream);

/* tcgen05 GEMM, bf16 operands, fp32 accumulation in TMEM, C [M,N] bf16 or fp32 (bias / ReLU / accumulate
 * flags as above).  a_mn = 0: A is [M,K] K-major (lda); a_mn = 1: A is stored transposed, At[K,M] (M
 * contiguous).  Same for B with N.  Forward: (0,0); dgrad: (0,1) with B = w[N,K] read as Bt[K'=N, N'=K];
 * wgrad: (1,1).  Requires 16-byte aligned bases and leading dimensions that are multiples of 8. */
int masr_umma_gemm(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn,
                   void* C, int c_dtype, int64_t ldc, const float* bias,
                   int M, int N, int K, int flags, int splitk, void* stream);
/* Fused epilogue / side outputs of masr_umma_gemm_ex (zero-initialise; NULL = none):
 *   rowsum     : rowsum[M] (fp32) += sum_k A(m,k).  For a wgrad GEMM (A = dy^T) this is the bias gradient
 *                (torch: dy.sum(0)); computed on the tensor cores by a second accumulator fed with an all-ones
 *                B tile, so no separate column-sum pass over dy is needed
 *   mask       : bf16 [M,N] (row stride ldmask): C = mask > 0 ? C * mask_scale : 0.  With mask = the stored
 *                output of ReLU(+dropout) and mask_scale = 1/(1-p): the backward of both, fused in a dgrad GEMM
 *   p_drop     : dropout on C after bias / ReLU with masr_dropout's element index m*N+n under (seed, site)
 *   dot_src    : bf16 [M,N] (row stride lddot), N = dot_H * 64, M = batch * dot_L: dot_out[(b * dot_H + h) * dot_L + q]
 *                = sum over the 64 columns of head h of C(m, .) * dot_src(m, .), m = b * dot_L + q.  With C = dO
 *                (out-projection dgrad) and dot_src = O this is the D vector of the attention backward
 *                (masr_umma_attn_bwd with dsum_ready = 1), fused into the GEMM epilogue  */
typedef struct masr_gemm_epilogue {
  float* rowsum;
  const void* mask;
  int64_t ldmask;
  float mask_scale;
  float p_drop;
  uint64_t seed;
  uint32_t site;
  const void* dot_src;
  int64_t lddot;
  float* dot_out;
  int dot_L;
  int dot_H;
} masr_gemm_epilogue;
int masr_umma_gemm_ex(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn,
                      void* C, int c_dtype, int64_t ldc, const float* bias,
                      int M, int N, int K, int flags, int splitk, const masr_gemm_epilogue* epi, void* stream);
/* The same contract on the PERSISTENT CTA-PAIR kernel (tcgen05.mma.cta_group::2: two CTAs of a cluster compute one
 * 256 x bn tile, each staging its own 128 rows of A and half of B; accumulator double-buffered in TMEM so that the
 * epilogue of a tile overlaps the MMAs of the next).  masr_umma_gemm / _ex dispatch to it by problem size (encoder /
 * front-end sized problems: nn.Linear of mono_transformer_torch.py:62,74-85 at M = B T/4 rows and their backward);
 * this entry selects it explicitly.  bn: 128 | 256 | 0 (choose); splitk <= 0: choose (with MASR_GEMM_SPLITK). */
int masr_umma_gemm_pair(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn,
                        void* C, int c_dtype, int64_t ldc, const float* bias,
                        int M, int N, int K, int flags, int splitk, int bn, const masr_gemm_epilogue* epi, void* stream);
/* Upper bound of the TMA operand ring depth of masr_umma_gemm* (0 = default policy: the whole shared memory when the
 * grid fits one wave).  The lock-step meta-step sets 3 while several task lanes run their small GEMMs concurrently, so
 * that two CTAs of different lanes fit one SM. */
/* Grouped launch of small GEMMs.  Between _begin and _end every masr_umma_gemm / masr_umma_gemm_ex call that resolves to
 * the one-tile tcgen05 kernel is recorded instead of launched (its operands must stay valid and unchanged until _end);
 * _end launches the recorded problems, all those of one kernel instantiation in ONE grid (up to 24 per launch).  Made for
 * the weight gradients of a batch (the implied backward of every nn.Linear of the decoder, mono_transformer_torch.py:
 * 74-98): ~26 independent 16-64 CTA problems that otherwise cost a launch each.  Calls that resolve to other kernels
 * (CTA-pair GEMM, CUDA-core GEMM) launch immediately.  Not re-entrant; per host thread. */
int masr_gemm_group_begin(void);
int masr_gemm_group_end(void* stream);
/* how many problems the last masr_gemm_group_end of this thread had recorded and how many kernels it launched for them */
int masr_gemm_group_last(int* recorded, int* launched);
int masr_gemm_set_stage_cap(int stages);
/* mode 0: masr_umma_gemm* never use the CTA-pair kernel; 1 (default): by problem size (A/B measurements). */
int masr_gemm_set_pair_mode(int mode);
/* Convenience form of the above: A [M,K] and B [N,K] both K-major ("TN"). */
int masr_umma_gemm_tn(const void* A, int64_t lda, const void* B, int64_t ldb,
                      void* C, int c_dtype, int64_t ldc, const float* bias,
                      int M, int N, int K, int flags, void* stream);

/* ------------------------------------------------------------------ conv front end
 * nn.Conv2d(3x3, stride 1, pad 1) + ReLU + MaxPool2d(2,2), mono_transformer_torch.py:49-60,116.
 */
/* Implicit-GEMM 3x3 convolution on tcgen05 (bf16 NHWC, channels 64 or 128): the 9 shifted input boxes are
 * fetched by 4-D TMA (out-of-bounds zero fill = padding), nothing is unfolded in memory.
 * wp is the [Cout, 9*Cin] bf16 layout of masr_conv_w_prep; dwp the same layout in fp32 (accumulated). */
int masr_umma_conv3x3_fwd(const void* x, const void* wp, const float* bias, void* y,
                          int B, int H, int W, int Cin, int Cout, int relu, void* stream);
/* wpt (may be NULL): the transposed weight layout [Cin][tap][Cout] of masr_conv_w_prep_t; when given, dgrad reads
 * its B operand K-major like the forward pass (faster than reading wp MN-major). */
int masr_umma_conv3x3_dgrad(const void* dy, const void* wp, const void* wpt, void* dx, const void* relu_src,
                            int B, int H, int W, int Cin, int Cout, void* stream);
/* First convolution (Cin = 1 -> 64) on the tensor cores: 3x3 patches (K = 9 padded to 16) are built in shared memory
 * as a UMMA operand; y / dy are bf16 NHWC, x / w / bias / dw / db fp32.  Same results as masr_conv1_fwd /
 * masr_conv1_wgrad up to the bf16 rounding of x (and w in the forward). */
int masr_umma_conv1_fwd(const float* x, const float* w, const float* bias, void* y,
                        int B, int H, int W, int Cout, void* stream);
int masr_umma_conv1_wgrad(const float* x, const void* dy, float* dw, float* db,
                          int B, int H, int W, int Cout, void* stream);
/* db (may be NULL): [Cout] fp32 bias gradient += sum over pixels of dy, fused (tensor-core row sums). */
int masr_umma_conv3x3_wgrad(const void* x, const void* dy, float* dwp, float* db,
                            int B, int H, int W, int Cin, int Cout, void* stream);
/* conv1: Cin = 1.  x [B,H,W] fp32 -> y [B,H,W,Cout] act, bias+ReLU fused.  w [Cout,9] fp32. */
int masr_conv1_fwd(const float* x, const float* w, const float* bias, void* y, int y_dtype,
                   int B, int H, int W, int Cout, void* stream);
/* dw [Cout,9] += sum_p dy[p,co] * x[p+tap]; db [Cout] += sum_p dy[p,co]   (dy already ReLU-masked) */
int masr_conv1_wgrad(const float* x, const void* dy, int dy_dtype, float* dw, float* db,
                     int B, int H, int W, int Cout, void* stream);
/* im2col for the SIMT path: col[p, tap*Cin+ci] = x[b, h+dh, w+dw, ci] (0 outside). */
int masr_im2col3x3(const void* x, void* col, int dtype, int B, int H, int W, int Cin, void* stream);
/* dx[q, ci] (=|+=) sum_tap dcol[q-(dh,dw), tap*Cin+ci]; optional ReLU mask: dx *= (mask_src > 0) */
int masr_col2im3x3(const void* dcol, void* dx, int dtype, const void* relu_src,
                   int B, int H, int W, int Cin, void* stream);
/* weight layout prep: w [Cout,Cin,3,3] fp32 -> wp [Cout, 9*Cin] (k = tap*Cin+ci) of dtype */
int masr_conv_w_prep(const float* w, void* wp, int dtype, int Cout, int Cin, void* stream);
/* w [Cout,Cin,3,3] fp32 -> wpt [Cin, 9*Cout] (tap-major inside a row): the dgrad operand layout */
int masr_conv_w_prep_t(const float* w, void* wpt, int dtype, int Cout, int Cin, void* stream);
/* dw [Cout,Cin,3,3] += dwp [Cout, 9*Cin] (fp32) */
int masr_conv_w_unprep_add(const float* dwp, float* dw, int Cout, int Cin, void* stream);
/* 2x2/2 floor-mode max pool, NHWC */
int masr_maxpool2x2_fwd(const void* x, void* y, int dtype, int B, int H, int W, int C, void* stream);
/* dx = scatter(dy) to the first arg-max of each window, times (x > 0) when relu_mask != 0 */
int masr_maxpool2x2_bwd(const void* x, const void* dy, void* dx, int dtype, int relu_mask,
                        int B, int H, int W, int C, void* stream);
/* The same pool with a per-output code byte (bits 0-1: which of the four inputs won, first maximum in (h, w) scan order;
 * bit 2: the maximum is positive): the backward scatters dy from the codes and no longer re-reads the full-resolution
 * activation (C % 8 == 0, 16-byte aligned tensors; code [B, H/2, W/2, C] uint8).  nn.MaxPool2d(2, stride=2) behind the
 * ReLU of mono_transformer_torch.py:52-58. */
int masr_maxpool2x2_fwd_code(const void* x, void* y, void* code, int dtype, int B, int H, int W, int C, void* stream);
int masr_maxpool2x2_bwd_code(const void* code, const void* dy, void* dx, int dtype, int relu_mask,
                             int B, int H, int W, int C, void* stream);
/* dx = (y > 0) ? dx * scale : 0, elementwise (scale = 1/(1-p): backward of ReLU followed by dropout(p) from the stored output) */
int masr_relu_bwd(const void* y, void* dx, int dtype, int64_t n, float scale, void* stream);

/* ------------------------------------------------------------------ attention
 * softmax(Q K^T / sqrt(hd) + mask) V with the masks built from lengths inside the kernel:
 * key-padding (src_key_padding_mask / memory_key_padding_mask of make_bool_pad_mask,
 * src/nets_utils.py:85-94) via klens[B] (NULL = none) and the causal mask of
 * generate_square_subsequent_mask (:9-15) via causal != 0.  Replaces
 * torch/nn/functional.py multi_head_attention_forward's SDPA call.
 * q/k/v/out rows are [b*L + l], head h occupies columns [h*hd, (h+1)*hd); ld* are row strides
 * in elements.  lse [B,H,Lq] fp32.  Attention-probability dropout (p_drop, seed, site) uses the
 * library's counter-based generator and is replayed in backward.
 */
int masr_attn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                  void* out, int64_t ldo, float* lse, int dtype,
                  int B, int H, int Lq, int Lk, int hd, const int64_t* klens, int causal,
                  float p_drop, uint64_t seed, uint32_t site, void* stream);
int masr_attn_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                  const void* out, int64_t ldo, const void* dout, int64_t lddo, const float* lse,
                  float* dsum_ws /* [B*H*Lq] fp32 scratch */, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int dtype,
                  int B, int H, int Lq, int Lk, int hd, const int64_t* klens, int causal,
                  float p_drop, uint64_t seed, uint32_t site, void* stream);

/* tcgen05 fast path of the two calls above: bf16, head dim 64, rows 16 B aligned (ld % 8 == 0).
 * Forward handles any Lk (online soft-max over 128-key tiles).  Backward: one CTA per 128-key tile of a (batch, head);
 * with Lk <= 128 dQ is written directly, with longer memories (utterances up to max_ilen 1500 -> 375 keys) the key tiles
 * add their dQ partials into dq_ws ([B*Lq, H*64] fp32, zeroed by the call) and a cast pass rounds the sum (dq_ws may be
 * NULL when Lk <= 128). */
int masr_umma_attn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                       void* out, int64_t ldo, float* lse, int B, int H, int Lq, int Lk,
                       const int64_t* klens, int causal, float p_drop, uint64_t seed, uint32_t site, void* stream);
/* The two forward calls with K / V that hold `kv_rows` >= Lk rows per utterance, of which the first Lk are attended to:
 * incremental greedy decoding reads a key/value cache of fixed capacity in place (rows beyond Lk must be finite). */
int masr_umma_attn_fwd_cached(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                              void* out, int64_t ldo, float* lse, int B, int H, int Lq, int Lk, int kv_rows,
                              const int64_t* klens, int causal, float p_drop, uint64_t seed, uint32_t site, void* stream);
int masr_attn_fwd_cached(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                         void* out, int64_t ldo, float* lse, int dtype,
                         int B, int H, int Lq, int Lk, int kv_rows, int hd, const int64_t* klens, int causal,
                         float p_drop, uint64_t seed, uint32_t site, void* stream);
int masr_umma_attn_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                       const void* out, int64_t ldo, const void* dout, int64_t lddo, const float* lse,
                       float* dsum_ws, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                       int B, int H, int Lq, int Lk, const int64_t* klens, int causal,
                       float p_drop, uint64_t seed, uint32_t site, int dsum_ready, float* dq_ws, void* stream);
/* Query sequences of at most max_lq rows (default and upper bound 64: the decoder's tgt length, mono_transformer_torch.py:200-203)
 * are served inside masr_umma_attn_fwd* / masr_umma_attn_bwd by warp-level MMA kernels that keep S / P / dS in registers
 * (one CTA per (batch, head)); longer ones take the 128-row tcgen05 tiles.  0 sends every problem to tcgen05 (tests). */
int masr_attn_set_small_lq(int max_lq);
/* dsum_ready != 0: dsum_ws already holds D[b,h,q] = dO . O (e.g. from masr_umma_gemm_ex's dot epilogue). */

/* ------------------------------------------------------------------ kernel 3: fused elementwise
 * residual + dropout + LayerNorm (post-norm TransformerEncoder/DecoderLayer, eps 1e-5):
 *   s = res + dropout(x)   (written back over x; res may be NULL)
 *   y = LN(s) * gamma + beta ; mean/rstd [rows] fp32 saved for backward.
 */
int masr_add_layernorm_fwd(void* x_inout, const void* res, const float* gamma, const float* beta,
                           void* y, float* mean, float* rstd, int dtype, int rows, int d, float eps,
                           float p_drop, uint64_t seed, uint32_t site, void* stream);
/* ds = LN'(dy) (written to ds; if ds_accum != 0: ds += ...), dgamma/dbeta += ;
 * dx (the sub-layer output branch) = ds * dropmask/(1-p) when dx != NULL */
int masr_add_layernorm_bwd(const void* dy, const void* s, const float* mean, const float* rstd,
                           const float* gamma, void* ds, int ds_accum, void* dx,
                           float* dgamma, float* dbeta, int dtype, int rows, int d,
                           float p_drop, uint64_t seed, uint32_t site, void* stream);
/* x = dropout(x + pe[row % L]) in place (PositionalEncoding, mono_transformer_torch.py:30-32) */
int masr_add_pe_dropout(void* x, const float* pe, int dtype, int rows, int L, int d,
                        float p_drop, uint64_t seed, uint32_t site, void* stream);
/* out[r] = dropout(E[ids[r]] + pe[r % L]) (preprocess(), :131-133); ids int64 */
int masr_embed_pe_fwd(const int64_t* ids, const float* E, const float* pe, void* out, int dtype,
                      int rows, int L, int d, float p_drop, uint64_t seed, uint32_t site, void* stream);
/* dE[ids[r]] += dropmask * dout[r]  (fp32 atomics) */
int masr_embed_bwd(const int64_t* ids, const void* dout, int dtype, float* dE, int rows, int L, int d,
                   float p_drop, uint64_t seed, uint32_t site, void* stream);
/* x = dropout(x) in place / dx = dropout'(dx) in place with the same (seed, site) */
int masr_dropout(void* x, int dtype, int64_t n, float p_drop, uint64_t seed, uint32_t site, void* stream);
/* out[n] += sum_m x[m, n]  (bias gradients) */
int masr_colsum_add(const void* x, int dtype, int64_t ldx, float* out, int M, int N, void* stream);
/* dst = cast(src) */
int masr_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream);
/* Every derived weight copy the engine needs before a run-batch, in ONE launch (the nn.Module weights the reference
 * reads in place, src/model/transformer_pytorch/mono_transformer_torch.py:49-60,113-122, in the layouts the kernels
 * want): shadow (may be NULL = already fresh) = cast(params[0:n]) to `dtype`; for every job wp = masr_conv_w_prep(w) and
 * (wpt != NULL) wpt = masr_conv_w_prep_t(w); v2e_p (may be NULL) = masr_permute_cf(v2e). */
typedef struct { const float* w; void* wp; void* wpt; int Cout, Cin; } masr_conv_prep_job;
int masr_prep_weights(const float* params, void* shadow, int64_t n, const masr_conv_prep_job* jobs, int njobs,
                      const float* v2e, void* v2e_p, int v2e_rows, int C, int F, int dtype, void* stream);
/* dst[r, f*C + c] = src[r, c*F + f]   (vgg2enc column permutation, (c,f) -> (f,c)); with add != 0
 * and the roles swapped it un-permutes a gradient: dst[r, c*F+f] += src[r, f*C+c] (fp32 only) */
int masr_permute_cf(const void* src, int src_dtype, void* dst, int dst_dtype, int rows, int C, int F,
                    int inverse_add, void* stream);

/* label-smoothed cross entropy + accuracy + gradient, src/transformer_torch_trainer.py:64-92.
 *   logits [N, C] fp32, gold [N] int64 (IGNORE_ID = -1 rows are skipped)
 *   stats  [4] double: {sum of row losses, n_correct, n_non_pad, 0} (accumulated; zero it first)
 *   argmax [N] int64 out (may be NULL); dlogits [N, C] out (fp32 or bf16 per dl_dtype, row stride ld_dl >= C
 *   elements; a bf16 gradient with ld_dl % 8 == 0 feeds the tcgen05 GEMMs directly) = d(mean loss)/d logits
 *   given inv_n = 1 / n_non_pad (may be NULL)
 *   inv_n_dev (may be NULL): device-resident 1 / n_non_pad overriding inv_n (CUDA-graph replay)
 */
int masr_ls_ce_fwd_bwd(const float* logits, const int64_t* gold, int N, int C, float eps, float inv_n,
                       const float* inv_n_dev, double* stats, int64_t* argmax, void* dlogits, int dl_dtype,
                       int64_t ld_dl, void* stream);

/* CTC-weight mixing (north_star kernel 3; espnet-style joint CTC / attention objective, an EXTENSION: the reference
 * carries only dead config for it, config/transformer/mono-test.yaml:44-50):
 *   total = (1 - w) * attention LS-CE + w * CTC.  The gradient scales are folded into the producing kernels
 *   (masr_ls_ce_fwd_bwd's inv_n = (1-w)/n, masr_ctc_fwd_bwd's grad_scale = w); this entry stores the CTC term next to
 *   the CE statistics so that one device->host read returns both: stats[3] = *ctc_loss, stats[4] = w. */
int masr_loss_mix(double* stats, const float* ctc_loss, float w, void* stream);
/* dst[r, 0:cols] = src[r, 0:cols] (dtype conversion), dst[r, cols:ld_dst] = 0: pads the fp32 CTC gradient rows to the
 * 16-byte row pitch the tcgen05 GEMMs need. */
int masr_cast_pad2d(const void* src, int src_dtype, int64_t ld_src, void* dst, int dst_dtype, int64_t ld_dst,
                    int rows, int cols, void* stream);

/* Dropout seeds: every dropout site draws from (seed + *dev_ptr, site, element index).  dev_ptr (set once per
 * process; NULL = offset 0) lives in device memory so a captured CUDA graph of the step replays with fresh
 * masks; masr_seed_bump advances it on the stream. */
int masr_set_seed_ptr(const uint64_t* dev_ptr);
int masr_seed_bump(uint64_t* dev_ptr, uint64_t inc, void* stream);

/* ------------------------------------------------------------------ kernel 4: flat multi-tensor ops
 * All operate on flat fp32 arenas of n elements (parameters laid out back to back).
 */
/* out[0] (double) = sum g^2 ; must be zeroed by the caller (or zero_first != 0) */
int masr_mt_sumsq(const float* g, int64_t n, double* out, int zero_first, void* stream);
/* clip_grad_norm_(max_norm) + SGD(momentum, nesterov) in one pass, fo_meta_interface.py:242-248.
 * Reads sum-of-squares from sumsq[0]; if it is NaN the step is skipped (math.isnan guard).
 * g is scaled in place by the clip coefficient (as clip_grad_norm_ does). first_step: buf = g. */
int masr_mt_clip_sgd(float* p, float* g, float* buf, int64_t n, const double* sumsq, float max_norm,
                     float lr, float momentum, int nesterov, int first_step, void* stream);
/* Same, and in the same pass: shadow_bf16 (may be NULL) = bf16(p) -- the compute-dtype copy of the arena that
 * TransformerEngine.prep_weights would otherwise produce with a separate cast pass before the next run-batch
 * (src/fo_meta_interface.py:229-240: the step is always followed by a forward on the new weights).
 * flags & MASR_SGD_LAST_STEP: no further inner step of this task follows (meta_k reached, :229): the scaled gradient
 * and the momentum buffer are dead and are not written back. */
#define MASR_SGD_LAST_STEP 1
int masr_mt_clip_sgd_ex(float* p, float* g, float* buf, int64_t n, const double* sumsq, float max_norm,
                        float lr, float momentum, int nesterov, int first_step, void* shadow_bf16, int flags,
                        void* stream);
/* dst = src and shadow_bf16 = bf16(src) in one pass (either destination may be NULL):
 * asr_model.load_state_dict(self._original) at the head of run_task (src/fo_meta_interface.py:226) on flat arenas */
int masr_mt_copy_cast(float* dst, void* shadow_bf16, const float* src, int64_t n, void* stream);
/* g *= min(1, max_norm / (sqrt(sumsq) + 1e-6)) */
int masr_mt_clip(float* g, int64_t n, const double* sumsq, float max_norm, void* stream);
/* FOMAML: upd += g * clipcoef(sumsq)   (fo_meta_interface.py:148-149,192-196); sumsq may be NULL */
int masr_mt_accumulate(float* upd, const float* g, int64_t n, const double* sumsq, float max_norm,
                       void* stream);
/* Reptile: upd += theta - phi */
int masr_mt_reptile_delta(float* upd, const float* theta, const float* phi, int64_t n, void* stream);
/* average + Adam: g = upd / count; m,v,p updated (torch/optim/adam.py single-tensor form);
 * bias corrections are passed pre-computed in double by the host.  skip_if_nan (may be NULL):
 * sumsq whose NaN skips the step (multi-task path, multi_interface.py:111-114). */
int masr_mt_adam(float* p, float* m, float* v, const float* upd, int64_t n, float count,
                 float lr, float beta1, float beta2, float eps, double bc1, double bc2,
                 const double* skip_if_nan, const double* clip_sumsq, float max_norm, void* stream);
/* The one collective of the path (src/fo_meta_interface.py:200-221 has none: the reference is single-process) done in the
 * NVSwitch: multicast_ptr = the multicast mapping of a symmetric allocation holding every rank's flat update arena; this
 * rank sums elements [begin, end) over all ranks (multimem.ld_reduce, addition inside the switch) and writes the sums
 * into every rank's copy (multimem.st).  The caller orders it against the arena's producers / consumers on the other GPUs
 * with a cross-GPU barrier before and after. */
int masr_nvls_allreduce_f32(void* multicast_ptr, int64_t begin, int64_t end, void* stream);
/* The outer update of a multi-GPU meta-step in one kernel (src/fo_meta_interface.py:200-221: `_updates /= counter`, noam-Adam
 * step on the meta weights; plus the collective the reference does not have).  mc_upd / mc_theta: multicast mappings of
 * the symmetric allocations holding every rank's update arena / meta weights; theta, m, v: this rank's local arenas (the
 * moments are maintained only for the owned slice [begin, end)).  g = (sum over ranks of upd)[i] / count, Adam, new theta
 * written into every rank's meta weights, upd[begin, end) cleared on every rank.  Elements >= n are only cleared.
 * Cross-GPU barriers before and after are the caller's. */
int masr_nvls_reduce_adam(void* mc_upd, void* mc_theta, const float* theta, float* m, float* v,
                          int64_t begin, int64_t end, int64_t n, float count, float lr, float beta1, float beta2,
                          float eps, double bc1, double bc2, void* stream);
/* Reptile interpolation outer update: theta -= eps * upd * inv_count */
int masr_mt_axpy(float* y, const float* x, float a, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* METAASR_B200_H_ */
