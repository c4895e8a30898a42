/* mrblip_b200.h -- C ABI of libmrblip_b200.so: the sm_100a kernels behind the BLIP2_MR hot path.
 *
 * The reference (sudo-Boris/mr-Blip) is pure Python/PyTorch and has NO FFI of its own: its
 * "operator API" for this path is the LAVIS nn.Module surface.  Each entry point
 * below therefore cites the reference PyTorch call site whose device work it replaces; the Python
 * host side (mr_blip_b200/*.py) mirrors the LAVIS classes and binds these symbols through ctypes
 * (mr_blip_b200/_lib.py).  INTEGRATION.md shows the reference-side patch.
 *
 * Conventions: every pointer is a DEVICE pointer borrowed from the caller (no allocation, no
 * ownership transfer); sizes are element counts; `ld*`/`*_rs`/`*_bs` are element strides;
 * dtype codes: 0 = fp16, 1 = bf16, 2 = fp32; `stream` is a cudaStream_t; calls are asynchronous and
 * stream-ordered, re-entrant per device.  Return 0 on success, <0 on error
 * (-1 bad argument, -2 CUDA error -- text from mrb_last_error(), -3 unsupported shape).
 */
#ifndef MRBLIP_B200_H
#define MRBLIP_B200_H
#ifdef __cplusplus
extern "C" {
#endif

int mrb_abi_version(void);
const char* mrb_last_error(void);

/* D[M,N] = epi(A[M,K] . B[N,K]^T): tcgen05/TMEM/TMA GEMM.  epi: +bias[N] (fp32), exact-erf GELU, + fp32 residual
 * (may alias out), output fp16/bf16/fp32; row_group = 256 selects the ViT patch-embed row remap (+cls slot, +pos_embed).
 * Replaces F.linear / nn.Linear / Conv2d(k=s=14) at: eva_vit.py:122-125,146,54-61,196-203; Qformer.py:185-196,
 * 278-289,349-375; blip2_mr.py:491; modeling_t5.py:323-329,542-558,611,1870 and their autograd dgrads. */
int mrb_gemm(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K, int dtype,
             const float* bias, int gelu, const float* resid, long long ldr, void* out, int out_dtype, long long ldc,
             int row_group, int force_bn, void* stream);

/* mrb_gemm with a caller-owned fp32 workspace (16-byte aligned).  Problems with a single 128-row tile (decoder steps of
 * generate / of the training decoder, modeling_t5.py:542-558 at M = B x L_dec) or with 32 output columns (LoRA
 * down-projections) leave most SMs idle and are bound by what one SM can stream: they are split along K over up to
 * max_splits (<= 8) CTAs per output tile, partials in ws[splits][M][N] summed in split order by a second launch that applies
 * the epilogue.  Needs ws_bytes >= splits * M * N * 4; any other problem, or a workspace that is too small, runs exactly as
 * mrb_gemm.  ws is busy until the call's work has finished on `stream` (one workspace per issuing stream). */
int mrb_gemm_splitk(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K, int dtype,
                    const float* bias, int gelu, const float* resid, long long ldr, void* out, int out_dtype, long long ldc,
                    int row_group, int force_bn, void* ws, long long ws_bytes, int max_splits, void* stream);

/* The (tile width, splits, K blocks of 64 per split) mrb_gemm_splitk would use on a device with `sms` SMs; splits == 1 is
 * the unsplit path.  Host arithmetic only (no device work): workspace sizing and tests. */
int mrb_gemm_splitk_plan(int M, int N, int K, int sms, int force_bn, int max_splits, int* bn, int* splits, int* kb_per_split);

/* Cap the SMs the persistent 2-CTA kernel of mrb_gemm occupies for the launches the calling thread issues from now on (sms even;
 * 0 lifts the cap).  For GEMMs a caller issues on a side stream NEXT TO a chain of small dependent kernels: a CTA of that kernel
 * holds its SM's whole shared memory, so an uncapped side GEMM stalls the chain for its length.  Used around the 48 encoder-sized
 * cross-attention K/V GEMMs (projection and dgrad) of the T5 decoder, which torch's autograd runs serially with the decoder's
 * M = B x L_dec kernels (modeling_t5.py:542-558 key_value_states branch).  No reference counterpart: scheduling only. */
int mrb_gemm_sm_limit(int sms);

/* softmax(scale * Q K^T + bias[h, j - i] + mask) V, scores never written to HBM; optional log-sum-exp for backward.
 * kv_div > 1: query batch b reads K/V/mask batch b / kv_div (beams sharing one encoder output).
 * Replaces eva_vit.py:128-145, Qformer.py:198-268, modeling_t5.py:561-610. */
int mrb_attention_fwd(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                      const void* v, long long v_bs, long long v_rs, void* o, long long o_bs, long long o_rs,
                      int B, int H, int Lq, int Lk, int hd, int dtype, float scale, const float* bias, int bias_len,
                      int bias_zero, const int* kmask, int kv_div, int causal, int q_pos0, float* lse, void* stream);
/* One query row per (batch, head), no bias / mask: the 257th ViT token (257 = 2 x 128 + 1; the two full tiles use the
 * tcgen05 kernel).  q / o point at that row of batch 0. */
int mrb_attention_row(const void* q, long long q_bs, const void* k, long long k_bs, long long k_rs, const void* v,
                      long long v_bs, long long v_rs, void* o, long long o_bs, int B, int H, int Lk, int hd, int dtype,
                      float scale, void* stream);
/* Same contract, tcgen05/TMEM/TMA implementation (S and O tiles in tensor memory, K/V by TMA, P through swizzled smem)
 * for the large shapes: hd == 64 (T5) or 64 < hd <= 96 (ViT hd 88, zero-filled to 96 by the tensor map). */
int mrb_attention_fwd_tc(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                         const void* v, long long v_bs, long long v_rs, void* o, long long o_bs, long long o_rs,
                         int B, int H, int Lq, int Lk, int hd, int dtype, float scale, const float* bias, int bias_len,
                         int bias_zero, const int* kmask, int kv_div, int causal, int q_pos0, float* lse, void* stream);
/* Self-attention of the EVA ViT over every (frame, head) -- Attention.forward of lavis/models/eva_vit.py:128-145 between the qkv
 * and proj Linears (no relative position bias at eva_vit.py:416-428) -- as ONE persistent tcgen05 kernel specialised for
 * L = 257 tokens (CLS + 256 patches) and 64 < hd <= 96 (hd % 8 == 0; ViT-g: 88): per item S = Q K^T is a 128 x 256 MMA per
 * query group, P is written back into tensor memory and O += P V reads it from there, the CLS token (as key and as query) is
 * rank-1 work on CUDA cores (csrc/attention_vit.cu).  q / k / v / o point at token 0, head 0 of frame 0; strides in elements. */
int mrb_attention_vit(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                      const void* v, long long v_bs, long long v_rs, void* o, long long o_bs, long long o_rs,
                      int frames, int H, int L, int hd, int dtype, float scale, void* stream);
/* dQ, dK, dV of the above (autograd of modeling_t5.py:561-610); hd <= 64. delta_ws: fp32 [B*H*Lq] workspace. */
int mrb_attention_bwd(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                      const void* v, long long v_bs, long long v_rs, const void* o, long long o_bs, long long o_rs,
                      const void* dout, long long do_bs, long long do_rs, void* dq, void* dk, void* dv,
                      int B, int H, int Lq, int Lk, int hd, int dtype, float scale, const float* bias, int bias_len,
                      int bias_zero, const int* kmask, int causal, int q_pos0, const float* lse, float* delta_ws,
                      void* stream);

/* tcgen05 implementation of mrb_attention_bwd (hd == 64): a dK/dV kernel (keys stationary, 64-query tiles streamed)
 * and a dQ kernel (queries stationary, 64-key tiles streamed); S^T/dP^T tiles and the dK/dV/dQ accumulators live in TMEM. */
int mrb_attention_bwd_tc(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                         const void* v, long long v_bs, long long v_rs, const void* o, long long o_bs, long long o_rs,
                         const void* dout, long long do_bs, long long do_rs, void* dq, void* dk, void* dv,
                         int B, int H, int Lq, int Lk, int hd, int dtype, float scale, const float* bias, int bias_len,
                         int bias_zero, const int* kmask, int causal, int q_pos0, const float* lse, float* delta_ws,
                         void* stream);

/* mode 0: LayerNorm(x [+ add]) (eva_vit.py:175-176, blip2.py:113-119, Qformer.py:104-108,288);
 * mode 1: T5 RMSNorm (modeling_t5.py:263-277).  fp32 in; optional fp32 out, 16-bit out (ld_h), and x+add out. */
int mrb_norm(const float* x, const float* add, const float* w, const float* bias, float eps, int rows, int C, int mode,
             float* out_f32, void* out_h, int h_dtype, long long ld_h, float* sum_out, void* stream);
/* dres += d/dx RMSNorm(x; w) . dy   (autograd of modeling_t5.py:263-277, weights frozen).  dy fp32 or 16-bit with row
 * stride ld_dy; with lora_A != NULL, dy is an extended dgrad buffer [rows, C+32] and dy += dy[:, C:C+R] . A first. */
int mrb_rmsnorm_bwd(const float* x, const float* w, const void* dy, int dy_dtype, long long ld_dy, const float* lora_A,
                    int R, float eps, int rows, int C, float* dres, void* stream);

/* frames fp32 [F,3,S,S] -> patch matrix [F*(S/P)^2, ldA] (16-bit), zero padded: A operand of the patch-embed GEMM
 * (eva_vit.py:196-203). */
int mrb_patchify(const float* img, void* out, int dtype, int frames, int img_size, int patch, int ldA, void* stream);
/* The same from RAW uint8 frames with the video processors' normalisation ((x/255 - mean) / std, CLIP constants of
 * lavis/processors/blip_processors.py:61-70) fused in: bit-identical patch matrix, a quarter of the PCIe bytes. */
int mrb_patchify_u8(const unsigned char* img, void* out, int dtype, int frames, int img_size, int patch, int ldA,
                    float mean0, float mean1, float mean2, float std0, float std1, float std2, void* stream);
/* x[f, 0, :] = cls_token + pos_embed[0]  (eva_vit.py:328-331) */
int mrb_cls_pos(const float* cls, const float* pos, float* x, int frames, int tokens, int C, void* stream);

/* h = gelu(ab[:, :F]) * ab[:, F:]  and its backward  (modeling_t5.py:323-329) */
int mrb_gated_gelu_fwd(const void* ab, void* h, int M, int F, long long ldh, int dtype, void* stream);
int mrb_gated_gelu_bwd(const void* ab, const void* dh, long long lddh, void* dab, long long lddab, int M, int F, int dtype,
                       void* stream);

/* inputs_embeds assembly (blip2_mr.py:691-783 interleave + embed_tokens lookups) from an int32 row table:
 * idx >= 0 embedding row, idx < 0 frame-token row -(idx+1), INT_MIN zero row; and its backward to the frame tokens. */
int mrb_gather_rows(const int* idx, const float* emb, const float* frames, float* out, int rows, int C, void* stream);
int mrb_scatter_frames(const int* idx, const float* dout, float* dframes, int rows, int C, void* stream);
/* frame_token_aggregation == "mean" (blip2_mr.py:493-498) and backward */
int mrb_group_mean(const float* x, float* out, int groups, int n, int C, void* stream);
int mrb_group_mean_bwd(const float* dout, float* dx, int groups, int n, int C, void* stream);

/* CrossEntropyLoss(ignore_index=-100) rows + d(logits) (modeling_t5.py:1872-1875).  gscale scales loss_sum and dlogits;
 * gscale < 0 selects the mean over valid targets with the count taken on the device (no host sync: graph-capturable). */
int mrb_cross_entropy(const float* logits, const long long* labels, int rows, int V, float* row_loss, void* dlogits,
                      int d_dtype, long long ldd, float gscale, float* loss_sum, void* stream);

/* LoRA (peft lora.Linear, configured at blip2_mr.py:193-200): x_ext[:, K:K+32] = x_ext[:, :K] . A^T (R = 8/16/24 rows) */
int mrb_lora_down(void* x_ext, long long ldx, const float* A, int M, int K, int R, int dtype, void* stream);
/* LoRA input gradient: x_ext[:, :K] += x_ext[:, K:K+R] . A (in place), or acc[M,K] (fp32) += that sum when acc != NULL */
int mrb_lora_up_add(void* x_ext, long long ldx, const float* A, int R, int M, int K, int dtype, float* acc, void* stream);
/* out[C,8] (or [8,C] when transposed_out) += P[M,C]^T . Q[M,8]: LoRA A/B weight gradients */
int mrb_skinny_wgrad(const void* P, long long ldp, const void* Q, long long ldq, int M, int C, float* out,
                     int transposed_out, int dtype, void* stream);

/* tcgen05 implementation of mrb_skinny_wgrad for large M (both operands read MN-major, split-K over M, fp32 atomics);
 * Q needs >= 16 readable columns per row. */
int mrb_skinny_wgrad_tc(const void* P, long long ldp, const void* Q, long long ldq, int M, int C, float* out,
                        int transposed_out, int dtype, void* stream);
/* Same pass over P for two adjacent LoRA slots: out += P^T Q[:, 0:8], out2 += P^T Q[:, 8:16] (q|k|v, wi_0|wi_1 and the
 * cross k|v Linears share one input, so their dA reductions share the read of x). */
int mrb_skinny_wgrad_tc2(const void* P, long long ldp, const void* Q, long long ldq, int M, int C, float* out, float* out2,
                         int transposed_out, int dtype, void* stream);
/* out[m, r] = sum_k x[m,k] W[r,k], r < 32, 16-bit: LoRA down-projection when M is tiny (decoder); larger M use mrb_gemm */
int mrb_small_down(const void* x, long long ldx, const void* W, long long ldw, int M, int K, void* out, long long ldo,
                   int dtype, void* stream);

/* Re-pack of the trainable LoRA A [8,K] / B [N,8] (fp32) of n Linears into their 16-bit operand slots after an
 * optimiser step: descs = device array of n records {A, B, ext_slot, ld, bdown_slot, ld, adown_slot, ld, extb_slot, ld,
 * K, N (int64), scale (double)} (13 x 8 bytes each; see LoraPackDesc in csrc/elementwise.cu).  peft's lora.Linear keeps
 * A/B as separate Parameters (blip2_mr.py:193-237); this is the layout change that lets base(x)+B(A x) be one GEMM. */
int mrb_lora_pack(const void* descs, int n, int blocks_per_linear, int dtype, void* stream);

/* plumbing: casts, 16-bit transpose, fp32 column sums (t5_proj bias grad), y = a*x + b*y */
int mrb_cast_f32_to_h(const float* in, void* out, long long n, int dtype, void* stream);
int mrb_cast2d_f32_to_h(const float* in, long long ld_in, void* out, long long ld_out, int rows, int cols, int dtype, void* stream);
int mrb_transpose16(const void* in, long long ld_in, void* out, long long ld_out, int rows, int cols, void* stream);
int mrb_colsum(const float* in, int rows, int C, float* out, void* stream);
int mrb_axpby(const float* x, float* y, long long n, float a, float b, void* stream);

/* ---- train-mode dropout (csrc/dropmask.cuh, csrc/dropout.cu) ----------------------------------------------------------------
 * The reference applies nn.Dropout in train() inside the frozen Q-Former (0.1: Qformer.py:107,258,287,373), T5 (0.1:
 * modeling_t5.py:327,346,600,652,690,1149,1258) and on every LoRA input (0.05: blip2_mr.py:197).  Masks here are a pure function
 * keep(seed, site, row, column) (counter hash, four 8-bit draws per 32-bit word; p quantised to 1/256, kept values scaled by
 * 256 / (256 - round(256 p))): forward and backward kernels and the CPU oracle evaluate the same function, nothing is stored.
 * `seed` points to ONE device word the host rewrites before every step (a replayed CUDA graph then draws fresh masks);
 * `site` identifies the dropout call within the step; rows / columns index the 2-D operand the site masks. */
/* out = drop(x):  (dtype, out_dtype) in {(f32, f32), (f32, 16-bit), (16-bit, same)}; cols % 4 == 0; out may alias x */
int mrb_dropout(const void* x, long long ldx, void* out, long long ldo, int rows, int cols, int dtype, int out_dtype,
                const unsigned* seed, unsigned site, float p, void* stream);
/* out = resid + drop(branch), fp32 [rows, cols] contiguous: hidden + dropout(sublayer) (modeling_t5.py:346,652,690) */
int mrb_dropout_add(const float* resid, const float* branch, float* out, int rows, int cols, const unsigned* seed,
                    unsigned site, float p, void* stream);
/* sum_out = x + drop(add) and out_h = T5LayerNorm(sum_out) * w as a 16-bit operand [rows, ld_h]: the residual add under dropout
 * (modeling_t5.py:346,652,690) fused with the next sublayer's norm (:263-277); = mrb_dropout_add then mrb_norm(mode 1), bit for
 * bit.  C % 4 == 0, C <= 2048; x, add, sum_out fp32 [rows, C] contiguous (sum_out may alias neither). */
int mrb_dropout_add_norm(const float* x, const float* add, const float* w, float eps, int rows, int C, void* out_h, int h_dtype,
                         long long ld_h, float* sum_out, const unsigned* seed, unsigned site, float p, void* stream);
/* mrb_rmsnorm_bwd (frozen weight, no LoRA fold) that also emits dy_next = drop_site(dres) as a 16-bit operand [rows, ld_next]:
 * the masked gradient the next sublayer's backward starts from (autograd of hidden + dropout(branch), modeling_t5.py:346,652,690),
 * = mrb_rmsnorm_bwd then mrb_dropout(fp32 -> 16 bit), bit for bit. */
int mrb_rmsnorm_bwd_drop(const float* x, const float* w, const void* dy, int dy_dtype, long long ld_dy, float eps, int rows, int C,
                         float* dres, void* dy_next, int next_dtype, long long ld_next, const unsigned* seed, unsigned site, float p,
                         void* stream);
/* mrb_gated_gelu_fwd / _bwd with the FF-inner dropout (modeling_t5.py:327): h = drop(gelu(a) * b) */
int mrb_gated_gelu_fwd_drop(const void* ab, void* h, int M, int F, long long ldh, int dtype, const unsigned* seed,
                            unsigned site, float p, void* stream);
int mrb_gated_gelu_bwd_drop(const void* ab, const void* dh, long long lddh, void* dab, long long lddab, int M, int F,
                            int dtype, const unsigned* seed, unsigned site, float p, void* stream);
/* peft lora.Linear in train mode, y = W x + B A drop_j(x) with an independent mask per adapted Linear j (site0 + j) of a group
 * of nlin <= 3 Linears sharing the input x:
 *   out[m, 8j + r] = scale * sum_k keep_j(m, k) x[m, k] A[8j + r, k]      (out: 32 16-bit columns, zeros from 8 nlin on)
 *   dA[r, k]      += scale * sum_m keep(m, k) x[m, k] q[m, r]              (q = dy sB_j, 16-bit [M, 8])
 *   dx[m, k]      += scale * sum_j keep_j(m, k) sum_r q[m, 8j + r] A[8j + r, k]   (dx 16-bit or fp32, read-modify-write) */
int mrb_lora_down_drop(const void* x, long long ldx, const void* A, long long lda, int M, int K, int nlin, void* out,
                       long long ldo, int dtype, const unsigned* seed, unsigned site0, float p, void* stream);
int mrb_lora_wgrad_drop(const void* x, long long ldx, const void* q, long long ldq, int M, int K, float* dA, int dtype,
                        const unsigned* seed, unsigned site, float p, void* stream);
int mrb_lora_dx_drop(const void* q, long long ldq, const void* A, long long lda, int nlin, void* dx, long long lddx,
                     int dx_dtype, int M, int K, int dtype, const unsigned* seed, unsigned site0, float p, void* stream);
/* Attention with train-mode dropout of the probabilities (modeling_t5.py:600, Qformer.py:258): the contracts of
 * mrb_attention_fwd / _fwd_tc / _bwd / _bwd_tc plus (seed, site, p); O = drop(softmax(S)) V, lse is that of the undropped scores,
 * the backward recomputes the mask (row = (b H + h) Lq + i, column = key).  hd <= 64; the tcgen05 versions need bf16 (forward)
 * and an even round(256 p). */
int mrb_attention_fwd_drop(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                           const void* v, long long v_bs, long long v_rs, void* o, long long o_bs, long long o_rs,
                           int B, int H, int Lq, int Lk, int hd, int dtype, float scale, const float* bias,
                           int bias_len, int bias_zero, const int* kmask, int kv_div, int causal, int q_pos0, float* lse,
                           const unsigned* seed, unsigned site, float p, void* stream);
int mrb_attention_fwd_tc_drop(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                              const void* v, long long v_bs, long long v_rs, void* o, long long o_bs, long long o_rs,
                              int B, int H, int Lq, int Lk, int hd, int dtype, float scale, const float* bias,
                              int bias_len, int bias_zero, const int* kmask, int kv_div, int causal, int q_pos0, float* lse,
                              const unsigned* seed, unsigned site, float p, void* stream);
int mrb_attention_bwd_drop(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                           const void* v, long long v_bs, long long v_rs, const void* o, long long o_bs, long long o_rs,
                           const void* dout, long long do_bs, long long do_rs, void* dq, void* dk, void* dv,
                           int B, int H, int Lq, int Lk, int hd, int dtype, float scale, const float* bias,
                           int bias_len, int bias_zero, const int* kmask, int causal, int q_pos0, const float* lse,
                           float* delta_ws, const unsigned* seed, unsigned site, float p, void* stream);
int mrb_attention_bwd_tc_drop(const void* q, long long q_bs, long long q_rs, const void* k, long long k_bs, long long k_rs,
                              const void* v, long long v_bs, long long v_rs, const void* o, long long o_bs, long long o_rs,
                              const void* dout, long long do_bs, long long do_rs, void* dq, void* dk, void* dv,
                              int B, int H, int Lq, int Lk, int hd, int dtype, float scale, const float* bias,
                              int bias_len, int bias_zero, const int* kmask, int causal, int q_pos0, const float* lse,
                              float* delta_ws, const unsigned* seed, unsigned site, float p, void* stream);

#ifdef __cplusplus
}
#endif
#endif
