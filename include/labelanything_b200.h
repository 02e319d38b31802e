/* labelanything_b200 — C ABI of the B200-native LabelAnything hot path.
 *
 * The reference (pasqualedem/LabelAnything) is pure Python: it has no FFI of its own, its "plugin API"
 * is the nn.Module surface `label_anything.models.LabelAnything` / `Lam`.  This header is
 * the C boundary our Python host code (labelanything_b200/*.py, which mirrors that module surface) binds
 * with ctypes.  Every entry point
 *   - is stateless and stream-ordered: it only enqueues work on `stream` (a cudaStream_t passed as void*),
 *   - takes plain device pointers + sizes (no torch types), never allocates device memory,
 *   - returns 0 on success or a negative LA_ERR_* code, with a message available from la_last_error().
 * Each declaration cites the reference code it replaces (paths relative to the reference repository).
 *
 * Layout conventions: activations are token-major ("NHWC"): [rows, channels] with channels contiguous.
 * bf16 = __nv_bfloat16 bit pattern.  All pointers must be 16-byte aligned unless stated otherwise.
 */
#ifndef LABELANYTHING_B200_H
#define LABELANYTHING_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define LA_B200_VERSION 100

/* error codes */
#define LA_OK 0
#define LA_ERR_INVALID (-1)
#define LA_ERR_CUDA (-2)
#define LA_ERR_UNSUPPORTED (-3)

/* epilogue activations */
#define LA_ACT_NONE 0
#define LA_ACT_GELU 1 /* exact erf GELU: nn.GELU(), label_anything/models/common.py:24 */
#define LA_ACT_RELU 2 /* nn.ReLU: label_anything/models/transformer.py:164, mask_decoder.py:797 */

/* element types of outputs */
#define LA_DTYPE_BF16 0
#define LA_DTYPE_F32 1
#define LA_DTYPE_F16 2

/* ---- library management ------------------------------------------------------------------------- */
const char* la_last_error(void); /* thread-local message of the last failing call */
int la_version(void);
int la_device_check(void); /* LA_OK iff the current device is sm_100 */

/* ---- dense contraction -------------------------------------------------------------------------- */
/* out[M,N] = act(a[M,K] @ w[N,K]^T + bias[N]);  a, w bf16 with K contiguous (nn.Linear weight layout),
 * fp32 accumulation on tcgen05 tensor cores, out bf16, fp32 or fp16 (out_dtype = LA_DTYPE_*).  lda/ldw/ldo are row strides in elements.
 * Replaces every nn.Linear / 1x1 Conv2d / im2col'd Conv2d / stride==kernel ConvTranspose2d on the path:
 *   label_anything/models/image_encoder.py:227-228,242,253,402-410; common.py:28-37,82-85,108-110,146;
 *   build_lam.py:150-171; mask_decoder.py:206-255,776-804. */
int la_gemm_bf16(void* stream, const void* a, long long lda, const void* w, long long ldw, const float* bias,
                 void* out, long long ldo, int out_dtype, int M, int N, int K, int act);

/* out[M,N] (fp32, zeroed by the call) = a @ w^T + bias with the CONTRACTION split over the SMs: work items are (output
 * tile, K range), partial tiles are added with TMA reduce stores (fp32 adds in L2: the summation order, not the
 * operands, varies from run to run).  For products with few output tiles and a long K -- the weight gradients
 * dW[N_out, K_in] = dY^T X of the training step, whose contraction runs over up to S*T = 54 000 token rows
 * (label_anything/experiment/run.py:359-361: autograd's grad-weight GEMMs). */
int la_gemm_bf16_splitk(void* stream, const void* a, long long lda, const void* w, long long ldw, const float* bias,
                        float* out, long long ldo, int M, int N, int K);

/* The same product stored into a PADDED GRID: the M rows of `a` are the pixels (image, y, x) of M / grid^2 square
 * grid x grid token maps; `out` (bf16, row stride ldo) is [M / grid^2][padded][padded][ldo] and receives row
 * (image, y, x) at position (image, y, x) of the padded grid (4-D TMA stores), while the padded^2 - grid^2 positions
 * outside get the bias row -- what the reference's zero padding after norm1 projects to
 * (label_anything/models/image_encoder.py:183-192,271-275: F.pad, then qkv = Linear(x)).  With it the windowed ViT
 * blocks project 64 x 64 instead of 70 x 70 tokens per image and la_attention_window_bf16 (in_pad > 0) fetches each
 * window from the padded grid.  grid % 32 == 0. */
int la_gemm_bf16_to_grid(void* stream, const void* a, long long lda, const void* w, long long ldw, const float* bias,
                         void* out, long long ldo, int M, int N, int K, int grid, int padded);

/* x[M,N] (fp32, row stride ldx) += a @ w^T + bias: the residual add of the ViT blocks done in the epilogue of the
 * GEMM that produces the branch, in fp32 on the fp32 accumulator (each epilogue warp TMA-loads its chunk of x into the
 * staging buffer it stores from).  CTA-pair kernel only: M >= 2048, N >= 256.
 *   label_anything/models/image_encoder.py:181-197 (x = x + mlp(norm2(x))) */
int la_gemm_bf16_accumulate(void* stream, const void* a, long long lda, const void* w, long long ldw, const float* bias,
                            float* x, long long ldx, int M, int N, int K);

/* 3x3 convolution (stride 1, zero padding 1) as an implicit GEMM on the CTA-pair tcgen05 kernel: x is a token-major
 * bf16 feature map [n_img, H, W, C] (W = 64, H % 4 == 0, C % 64 == 0), w the bf16 weight [N, 9*C] with column
 * (ky*3 + kx)*C + ci (row stride ldw), out [n_img*H*W, N] bf16 / fp32 (row stride ldo); out = act(conv(x) + bias).
 * The im2col matrix is never written: every k-block is one 4-D TMA box at the tap's shifted coordinates and the
 * border is zero-filled by the TMA unit.  Replaces Conv2d(3x3) of the necks: label_anything/models/build_lam.py:162-168,
 * image_encoder.py:92-108. */
int la_conv3x3_bf16(void* stream, const void* x, int n_img, int H, int W, int C, const void* w, long long ldw,
                    const float* bias, void* out, long long ldo, int out_dtype, int N, int act);

/* ---- fused attention ----------------------------------------------------------------------------- */
/* Multi-head self-attention, head_dim 64, over n_seq sequences of seq_len tokens stored as consecutive rows
 * of the projection matrices q [rows_total, ld_q] and kv [rows_total, ld_kv] (bf16; they may be the same
 * packed qkv buffer).  Head h reads q at column q_off + 64*h and k / v at columns {k,v}_off + 64*h of kv.
 * softmax(scale * q k^T + bias) v -> out [.., ld_out] bf16 at columns 64*h.
 * Optional decomposed relative-position bias (bias_h/bias_w != NULL): tables [rows_total][n_heads][ldb], fp32 or
 * fp16 (bias_dtype = LA_DTYPE_F32 / LA_DTYPE_F16; the rel_w part is rounded to fp16 inside the kernel anyway)
 * with table[row][h][grid_hw-1 - q_pos + k_pos] = q_row(h) . rel_pos[q_pos - k_pos + grid_hw-1], i.e. the
 * product of the head's q rows with the REVERSED rel_pos table (computed with la_gemm_bf16); grid_hw = 64
 * (global blocks, seq_len 4096).  The 14x14 windowed blocks use la_attention_window_bf16 below.
 * out_mode 0: out row = sequence*seq_len + token.  out_mode 1: window un-partition: sequence = image*nwin^2
 * + window, token (ty,tx) of window (wy,wx) goes to image row (wy*14+ty)*img_hw + wx*14+tx, padded
 * positions are dropped.
 *   label_anything/models/image_encoder.py:239-255,282-304,340-376; transformers modeling_vit.py:199-250 */
int la_attention_bf16(void* stream, const void* q, long long ld_q, int q_off, const void* kv, long long ld_kv,
                      int k_off, int v_off, long long rows_total, int n_seq, int seq_len, int n_heads, float scale,
                      const void* bias_h, const void* bias_w, int bias_dtype, int ldb, int grid_hw, void* out,
                      long long ld_out, int out_mode, int nwin, int img_hw);

/* 14x14-window attention of the SAM ViT blocks (seq_len 196, head_dim 64) with the decomposed relative-position
 * bias computed INSIDE the kernel: rel_table is the bf16 operand [2 * rel_pad][64], rel_pad = 32, rows [0, 27) = the
 * REVERSED rel_pos_h table (row i = rel_pos_h[26 - i]), rows [rel_pad, rel_pad + 27) = the reversed rel_pos_w table,
 * all other rows zero.  Per work item one extra tcgen05.mma forms T = Q_tile x rel_table^T in tensor memory and the
 * bias of key (kh, kw) for a query at (qh, qw) is T[13 - qh + kh] + T[rel_pad + 13 - qw + kw] -- same arithmetic as
 * the fp32 tables of la_attention_bf16 without their HBM round trip.  out_mode / nwin / img_hw as above.
 * in_pad = 0: q / kv hold window-partitioned rows (sequence * 196 + token).  in_pad > 0: q / kv are padded-grid
 * tensors [image][in_pad][in_pad][ld] written by la_gemm_bf16_to_grid (padding positions = bias row) and window
 * (wy, wx) of an image is fetched as one 4-D TMA box at (x, y) = (14 wx, 14 wy); sequence = image * nwin^2 + wy * nwin + wx.
 * The output leaves through one 4-D TMA store per (window, head) -- (channel, x, y, image) for out_mode 1, whose box drops the
 * rows and columns past the image -- so `out` must be 16-byte aligned (ld_out % 8 == 0 as everywhere).
 *   label_anything/models/image_encoder.py:239-255,258-304,319-376 */
int la_attention_window_bf16(void* stream, const void* q, long long ld_q, int q_off, const void* kv, long long ld_kv,
                             int k_off, int v_off, long long rows_total, int n_seq, int n_heads, float scale,
                             const void* rel_table, int rel_pad, void* out, long long ld_out, int out_mode, int nwin,
                             int img_hw, int in_pad);

/* Diagnostics: when device_buffer != NULL, CTA (0,0,0) of every following la_attention_bf16 launch records clock64()
 * stamps into it: int64 [5 roles (MMA issuers, softmax A / B first warp, softmax A / B last warp)][192 tiles]
 * [4 events]; NULL switches it off. */
int la_attention_set_trace(void* device_buffer);   /* LA_ERR_UNSUPPORTED unless built with -DLA_ATT_TRACE */

/* Token -> image attention WITHOUT materialised k / v projections (few query rows against many image tokens):
 *   y[s, r, :] = sum_t softmax_t(scale * (u[s, r] . x[s, t] + e[s, r, t])) x[s, t, :]      r < rows <= 8
 * x bf16 [n_seq * tokens, d] (row stride ldx) is BOTH the key and the value operand and is read exactly once;
 * u bf16 [n_seq * rows, d]; e fp32 [n_seq * rows, >= tokens] (row stride lde) or NULL; y bf16 [n_seq * rows, d]; d in {64, 128, 256, 512}.
 * With u_h = W_k[h]^T q_h, e = u . pe^T and o_h = W_v[h] y_h + b_v[h] this is exactly Attention(q, k = keys + pe,
 * v = keys) of label_anything/models/transformer.py:311-318 / common.py:97-148 (the k bias is constant over t and
 * drops out of the softmax); the small projections around it are la_gemm_bf16 calls. */
int la_attention_pooled_bf16(void* stream, const void* x, long long ldx, const void* u, const float* e, long long lde,
                             float scale, void* y, long long n_seq, int tokens, int rows, int d);
/* mode 0: out[(s, h), c] = in[s, c] if c / head_dim == h else 0   (bf16 [n_seq, H*dh] -> [n_seq*H, H*dh]);
 * mode 1: out[s, h*dh + j] = in[(s, h), h*dh + j]                 (bf16 [n_seq*H, H*dh] -> [n_seq, H*dh]). */
int la_head_rows_bf16(void* stream, const void* in, void* out, long long n_seq, int heads, int head_dim, int mode);

/* ---- streaming row kernels ----------------------------------------------------------------------- */
/* x = x_in[(row % x_mod) if x_mod > 0 else row] + delta[row]  (fp32 + bf16); optionally stored to x_out (may
 * alias x_in), plus the optional addends delta2[row] (bf16) and seq_add[row / seq_rows] (fp32, one vector per
 * sequence); y = act(LayerNorm(x) * gamma + beta) (biased variance, `eps`), or a plain cast when
 * gamma == NULL.  Outputs (each optional): y_out as bf16 / fp32 (y_dtype), y2_out = y in fp32,
 * ype_out = bf16(y + pe[row % pe_mod]) (positional table folded in for the next projection).  Row remapping:
 *   map_mode 0: identity.
 *   map_mode 1: `rows` counts OUTPUT rows in window-partitioned order (image, wy, wx, ty, tx) for win x win
 *               windows, nwin per side, over an hw x hw grid; rows that fall in the zero padding are written
 *               as zeros (F.pad happens after norm1).  image_encoder.py:183-187,258-279
 *   map_mode 2: drop token 0 (CLS) of every seq_len-token sequence.  build_encoder.py:98
 *   map_mode 3: pixel shuffle of a stride-2 ConvTranspose2d computed as a GEMM: source row (img, y, x, ky, kx) over an
 *               hw x hw grid -> destination row (img, 2y+ky, 2x+kx).  mask_decoder.py:206-222
 *   image_encoder.py:181-197; common.py:42-54,183-184; transformer.py:308-327; mask_decoder.py:214-215,250-254;
 *   transformers modeling_vit.py:325-346,416 */
int la_add_layernorm(void* stream, const float* x_in, long long x_mod, const void* delta, const void* delta2,
                     const float* seq_add, long long seq_rows, float* x_out, const float* gamma, const float* beta,
                     float eps, int act, void* y_out, int y_dtype, float* y2_out, const float* pe, long long pe_mod,
                     void* ype_out, long long rows, int d, int map_mode, int seq_len, int win, int nwin, int hw);

/* out[s, :] = mean over the rows_per_seq rows of sequence s of LayerNorm(x_in + delta + delta2 + seq_add[s]): the
 * last image-token LayerNorm of the prompt encoder's two-way transformer fused with the spatial average
 * pooling (the normalised tokens are never written).  partial_ws: fp32 scratch [n_seq * slices * d];
 * deterministic.   label_anything/models/transformer.py:326-327 + prompt_encoder.py:733-735 */
int la_add_layernorm_meanpool(void* stream, const float* x_in, const void* delta, const void* delta2,
                              const float* seq_add, const float* gamma, const float* beta, float eps,
                              long long n_seq, int rows_per_seq, int d, float* partial_ws, int slices, float* out);

/* x[img, tok, :] = (tok < n_cls ? cls : patch[img, tok - n_cls, :]) + pos[tok, :]; patch bf16, x fp32.
 *   image_encoder.py:112-114; transformers modeling_vit.py:109-125 */
int la_embed_tokens(void* stream, const void* patch, const float* cls, const float* pos, float* x, long long n_img,
                    int tokens_per_img, int n_cls, int d);

/* images [n_img, channels, size, size] fp32 (NCHW) -> [n_img*(size/16)^2, channels*256] bf16 patch rows,
 * column = c*256 + ky*16 + kx (matches Conv2d weight.flatten(1)).  image_encoder.py:402-410 */
int la_im2col_patch16(void* stream, const float* images, void* out, long long n_img, int channels, int size);

/* token-major bf16 map [n_img, height, width, channels] -> [n_img*height*width, 9*channels] bf16 rows,
 * column = (ky*3+kx)*channels + c, zero padding 1.  build_lam.py:162-168; mask_decoder.py:241-247 */
int la_im2col_3x3(void* stream, const void* in, void* out, long long n_img, int height, int width, int channels);

/* ---- token <-> image attention of the two-way transformer -------------------------------------------- */
/* Multi-head attention softmax(scale * (q + q_add)(k + k_add)^T) v for n_seq sequences of nq queries and nk keys;
 * q [n_seq*nq, ld_q], k [n_seq*nk, ld_k], v [n_seq*nk, ld_v], out [n_seq*nq, ld_out], all bf16 with head h at
 * columns [h*head_dim, (h+1)*head_dim) of the given base pointers (pass base + column offset for packed buffers).
 * q_add [nq, ld_qadd] / k_add [nk, ld_kadd]: optional fp32 tables shared by all sequences (the projected image
 * positional encoding), indexed by the query / key position.  head_dim in {8, 16, 32, 64}.  No masks: the
 * reference's key_mask / query_mask are no-ops (common.py:117-139).  Few queries against many keys run
 * key-parallel and may need `workspace` (la_attention_tokens_workspace_bytes; 0 = none).
 *   label_anything/models/common.py:97-148; transformer.py:245-250,300-327 */
int la_attention_tokens_splits(long long n_seq, int nq, int nk);
long long la_attention_tokens_workspace_bytes(long long n_seq, int nq, int nk, int n_heads, int head_dim);
int la_attention_tokens(void* stream, const void* q, long long ld_q, const void* k, long long ld_k, const void* v,
                        long long ld_v, const float* q_add, long long ld_qadd, const float* k_add,
                        long long ld_kadd, void* out, long long ld_out, long long n_seq, int nq, int nk, int n_heads,
                        int head_dim, float scale, void* workspace);

/* ---- prompt encoder ------------------------------------------------------------------------------------ */
/* masks [n_seq, height, width] fp32 -> out [n_seq, height/4, width/4, 16] fp32:
 * Conv2d(1,4,2,2) -> LayerNorm2d -> GELU -> Conv2d(4,16,2,2) -> LayerNorm2d -> GELU.  The 10 weight arrays are HOST
 * pointers (352 floats, passed by value to the kernel): w0 [4,1,2,2], w3 [16,4,2,2] in PyTorch layout.
 *   label_anything/models/prompt_encoder.py:61-67,516-531 */
int la_mask_downscale(void* stream, const float* masks, float* out, long long n_seq, int height, int width,
                      const float* w0, const float* b0, const float* ln1_w, const float* ln1_b, float eps1,
                      const float* w3, const float* b3, const float* ln2_w, const float* ln2_b, float eps2);

/* token-major fp32 [n, in_h, in_w, channels] -> [n, out_h, out_w, channels], bilinear, align_corners=False
 * (F.interpolate semantics).  prompt_encoder.py:787-793 */
int la_resize_bilinear(void* stream, const float* in, float* out, long long n, int in_h, int in_w, int out_h,
                       int out_w, int channels);

/* src[s, t, :] = feat[img(s), t, :] + dense(s, t) + code[s % n_classes]  -> bf16 [n_seq*tokens, d], with
 * dense(s, t) = w6 . m16[s, t, :] + b6 (mask_downscaling[6]), or not_a_mask when mask_flags[s] == 0, or no_mask when
 * m16 == NULL (no mask prompts).  Sequence s = ((b * examples) + m) * n_classes + c reads the features of image
 * img(s) = b * (examples + feat_lead) + feat_lead + m of feat fp32 [n_img * tokens, d] (feat_lead = 1 skips the query
 * image stored in front of every episode's support images, lam.py:167-168); m16 fp32 [n_seq, tokens, 16];
 * w6 [d, 16]; code [n_classes, d] or NULL.   prompt_encoder.py:68,532-539,637-646,795-805,250-264 */
int la_build_src(void* stream, const float* feat, const float* m16, const unsigned char* mask_flags,
                 const float* w6, const float* b6, const float* not_a_mask, const float* no_mask, const float* code,
                 void* out, long long n_seq, int tokens, int d, int n_classes, int examples, int feat_lead);

/* Sparse prompt tokens [n_seq, n, d] fp32, n = (n_points + (boxes ? 0 : 1)) + 2*n_boxes: random-Fourier positional
 * encoding of point / box-corner coordinates (+0.5, normalised by the image size) plus the label dependent
 * embeddings; pe_table [4, d] = point_embeddings[0..3].   prompt_encoder.py:83-114,201-211,226-233,648-669 */
int la_embed_sparse(void* stream, const float* points, const float* point_labels, int n_points, const float* boxes,
                    const float* box_flags, int n_boxes, const float* gauss, const float* not_a_point,
                    const float* pe_table, float* out, long long n_seq, int d, int image_w, int image_h);

/* out[b, c, :] = sum_m flags[b,m,c] * emb[b,m,c,:] / max(sum_m flags[b,m,c], 1).   prompt_encoder.py:738-745 */
int la_masked_mean(void* stream, const float* emb, const unsigned char* flags, float* out, int batch, int examples,
                   int classes, int d);

/* ---- layout conversion at the module boundary ---------------------------------------------------------------- */
/* [n, channels, pixels] fp32 (NCHW) -> [n, pixels, channels] fp32 and/or bf16 (either output may be NULL).
 *   label_anything/models/lam.py:139-146 (precomputed `embeddings` input) */
int la_nchw_to_tokens(void* stream, const float* in, float* out_f32, void* out_bf16, long long n, int channels,
                      int pixels);
/* [n, pixels, channels] fp32 -> [n, channels, pixels] fp32.   image_encoder.py:119-131 (x.permute(0, 3, 1, 2)) */
int la_tokens_to_nchw(void* stream, const float* in, float* out, long long n, int channels, int pixels);
/* out[s, r, :] = in[s * in_stride_rows + in_offset_rows + r, :] for r < slab_rows: gathers equally spaced row slabs
 * (the query image of every episode, lam.py:167 `embeddings[:, 0]`) as fp32 and/or bf16. */
int la_copy_slabs(void* stream, const float* in, long long in_stride_rows, long long in_offset_rows, float* out_f32,
                  void* out_bf16, long long n_slabs, long long slab_rows, int d);

/* out[r, :] = a[r, :] + b[(r / row_div) % b_mod, :], fp32 rows of d channels: the RandomMatrixEncoder class code added
 * to every sparse token of its class.   prompt_encoder.py:250-256 */
int la_add_bcast(void* stream, const float* a, const float* b, float* out, long long rows, int d, long long row_div,
                 long long b_mod);
/* out[o, j, i, :] = in[o, i, j, :] for i < na, j < nb (fp32 rows): "b m c d -> b c m d" around example_attention.
 *   prompt_encoder.py:706-710 */
int la_permute_rows(void* stream, const float* in, float* out, long long outer, int na, int nb, int d);

/* ---- mask decoder / post-processing ---------------------------------------------------------------------- */
/* out[b, c, p] = sum_k cls[b, c, k] * x[b, p, k];  x bf16 [batch*pixels, dk], cls fp32 [batch, classes, dk],
 * out fp32 [batch, classes, pixels].   label_anything/models/mask_decoder.py:299-314 */
int la_classify(void* stream, const void* x, const float* cls, float* out, int batch, long long pixels, int classes,
                int dk);

/* logits [batch, classes, low_h, low_w] fp32 -> out [batch, classes, out_h, out_w] fp32: bilinear to
 * image_size x image_size, crop [0:ih, 0:iw], bilinear to (oh, ow), pad with -inf (class 0: 0); classes with
 * flag_gts[b, c] == 0 become -inf.  sizes int32 [batch, 4] = (oh, ow, ih, iw) (device memory).
 *   label_anything/models/lam.py:383-453,92-93 */
int la_postprocess_masks(void* stream, const float* logits, float* out, const int* sizes,
                         const unsigned char* flag_gts, int batch, int classes, int low_h, int low_w, int image_size,
                         int out_h, int out_w);

/* ---- after the logits: prediction, global labels, confusion matrix ----------------------- */
/* One pass over [batch, classes, pixels] fp32 logits (or over int64 local predictions preds_in [batch, pixels]):
 *   pred = argmax over classes (first maximum, NaN wins: torch.argmax);
 *   pred, target = label_map[b][.] applied to values in [0, map_len) (others, e.g. -100, pass through) -- the
 *     composition of to_global_multiclass's sequential substitutions, built by the host;
 *   preds_out / gt_out (int64 [batch, pixels], optional) receive the mapped labels;
 *   confmat[target * num_classes + pred] += 1 for every pixel whose target != ignore_index (int64 [G, G],
 *     ACCUMULATED, optional); pixels with target or pred outside [0, num_classes) are skipped and counted in
 *     invalid[0] (torchmetrics raises for them under validate_args).
 * Any of logits / preds_in / gt / label_map / outputs may be NULL (label remapping only, argmax only, ...).
 * Replaces  label_anything/experiment/run.py:520-541,696-704 (`outputs.argmax(dim=1)`, to_global_multiclass,
 * metric update), label_anything/data/utils.py:567-590, label_anything/utils/metrics.py:28-42 and the confusion
 * matrix update of torchmetrics 1.7.1 MulticlassJaccardIndex (third party, uv.lock:2672-2673). */
int la_label_confusion(void* stream, const float* logits, const long long* preds_in, const long long* gt,
                       const long long* label_map, long long* preds_out, long long* gt_out, long long* confmat,
                       long long* invalid, int batch, int classes, long long pixels, int map_len, int num_classes,
                       long long ignore_index);

/* ---- loss of the training / validation step (first piece) ---------------------------------- */
/* Label-frequency class weights of get_weight_matrix_from_labels (label_anything/loss/utils.py:17-42):
 * hist (int64 [classes + 2], overwritten) = counts of labels 0..classes-1, of ignore_index, of anything else;
 * class_w[c] = 1 / log(1.1 + n_c / n) for the classes that occur, 1 otherwise (the weight of ignored pixels is 0). */
int la_label_class_weights(void* stream, const long long* labels, long long n, int classes, long long ignore_index,
                           long long* hist, float* class_w);

/* Focal loss with optional class weighting over logits [batch, classes, pixels] fp32, target int64 [batch, pixels]:
 *   ce = cross_entropy(x, target) (0 where target == ignore_index), pt = exp(-ce),
 *   loss = mean|sum((1 - pt)^gamma * class_w[target] * ce)          -> loss_out[0] (deterministic summation)
 *   grad_out (optional, [batch, classes, pixels]) = grad_scale[0] * d loss / d logits (grad_scale NULL = 1)
 *   wtarget_out (optional, [batch, pixels]) = class_w[target] (0 where ignored): the reference's weight matrix.
 *     (logits may be NULL when only wtarget_out is requested.)
 * workspace: la_focal_loss_workspace_bytes() bytes of 8-byte aligned scratch, no initialisation required (the entry
 * point zeroes its own counter on the stream); one workspace per call in flight.  Replaces label_anything/loss/focal.py:8-25 and the focal branch of LabelAnythingLoss.logits_loss
 * (label_anything/loss/__init__.py:67-92). */
long long la_focal_loss_workspace_bytes(void);
int la_focal_loss(void* stream, const float* logits, const long long* target, const float* class_w,
                  const float* grad_scale, float* loss_out, float* grad_out, float* wtarget_out, void* workspace,
                  int batch, int classes, long long pixels, float gamma, long long ignore_index, int mean);

/* Corrective prompt points of the iterative-prompting loop: `generate_points_from_errors`
 * (label_anything/experiment/substitution.py:17-96) + the coordinate scaling of Substitutor.generate_new_points
 * (substitution.py:161-171).  logits [batch, classes, height, width] fp32, gt [batch, height, width] int64 (ignore_index
 * counts as background, :33).  For every (b, c): the pixels where exactly one of argmax(logits) == c, gt == c holds
 * are its errors, in row-major order; point i is error number (|rnd[b, c, i]| mod count): points[b, c, i] = (x * sx[b],
 * y * sy[b]) in fp32, labels[b, c, i] = +1 (gt == c: false negative) / -1 (false positive), 0 for class 0 and for
 * classes without errors (whose point is (0, 0)).  workspace: la_error_points_workspace_bytes(batch, classes, height). */
long long la_error_points_workspace_bytes(int batch, int classes, int height);
int la_error_points(void* stream, const float* logits, const long long* gt, int batch, int classes, int height,
                    int width, long long ignore_index, const long long* rnd, int n_points, const float* sx,
                    const float* sy, void* workspace, float* points, float* labels);

/* ---- input preprocessing (the step right before Lam.forward) ----------------------------------------------- */
/* One image: uint8 HWC [H, W, 3] -> fp32 CHW [3, S, S] = zero_pad((resize_bilinear_PIL(img, new_h, new_w) / 255 - mean)
 * / std), bit-identical to CustomResize -> ToTensor -> CustomNormalize (label_anything/data/transforms.py:14-46; the
 * non-custom Resize((S, S)) -> ToTensor -> Normalize of data/__init__.py:33-61 is new_h = new_w = S).  The resize is
 * Pillow's 8-bit antialiased triangle resample (Resample.c): bounds_* int32 [new, 2] = (first source index, count),
 * kk_* int32 [new, ksize_*] = 22-bit fixed-point coefficients, computed by the caller exactly as Pillow does
 * (labelanything_b200/transforms.py::pil_bilinear_coeffs); NULL tables for an axis whose size does not change.
 * tmp: uint8 scratch [H, new_w, 3] (horizontal pass output), required when bounds_x != NULL. */
int la_preprocess_image_u8(void* stream, const void* src, int H, int W, int new_h, int new_w, int S,
                           const int* bounds_x, const int* kk_x, int ksize_x, const int* bounds_y, const int* kk_y,
                           int ksize_y, void* tmp, float mean0, float mean1, float mean2, float std0, float std1,
                           float std2, float* out);
/* PromptsProcessor.apply_masks (transforms.py:196-224): OR of n uint8 instance masks [n, H, W], nearest resize to
 * (new_h, new_w), zero pad to long_side, nearest resize to out_side (new_h = new_w = 0: straight nearest resize, the
 * non-custom pipeline) -> out fp32 [out_side, out_side] in {0, 1}; *flag (uint8, optional) is set to 1 when any output
 * pixel is set (flag_masks, data/utils.py:218-224) and left untouched otherwise.  n = 0 -> zeros. */
int la_rasterize_masks_u8(void* stream, const void* masks, int n, int H, int W, int new_h, int new_w, int long_side,
                          int out_side, float* out, void* flag);
/* PromptsProcessor.apply_coords / apply_boxes (transforms.py:159-194): n (x, y) pairs in float64, x * sx and y * sy in
 * double, rounded once to fp32 (the reference assigns the float64 result into a float32 tensor). */
int la_scale_coords_f64(void* stream, const void* coords, long long n, double sx, double sy, float* out);

/* ---- training step (BASELINE config 4: frozen / pre-computed encoder, MAE-L-256) --------------
 * Backward passes of the ops between the image embeddings and the loss, the fp32 forward variants the training path
 * keeps its activations in, and the optimiser update.  The reference gets all of this from torch.autograd over
 * label_anything/models/{lam,prompt_encoder,transformer,mask_decoder,common}.py, driven by
 * label_anything/experiment/run.py:359-361 (accelerator.backward) and :425-590 (train loop, AdamW step).  The
 * contractions (dgrad / wgrad of every nn.Linear / Conv2d / ConvTranspose2d) are la_gemm_bf16 launches on operands
 * transposed by la_cast_transpose_bf16; everything below is elementwise / row / reduction work in fp32.
 * Outputs documented as "accumulated" are added to (+=); all others are overwritten (zeroed inside where the kernel
 * scatters with atomics). */
/* out = bf16(in), n elements (GEMM operand of an fp32 activation) */
int la_cast_bf16(void* stream, const float* in, void* out, long long n);
/* hi = bf16(in), lo = bf16(in - hi): the operand split of the fp32-accurate GEMM mode (a w^T = hi_a hi_w^T + hi_a lo_w^T
 * + lo_a hi_w^T up to 2^-16 relative, three la_gemm_bf16 launches with fp32 accumulation) -- north_star's fp32 tolerance
 * for the training path and the yardstick the bf16 gradients are checked against */
int la_split_bf16(void* stream, const float* in, void* hi, void* lo, long long n);
/* three terms hi + mid + lo = in to 2^-24: the "bf16x6" mode (the six products whose orders sum to <= 2), fp32-level
 * accuracy of every contraction -- north_star's 1e-5 fp32 tolerance on the tensor cores */
int la_split3_bf16(void* stream, const float* in, void* hi, void* mid, void* lo, long long n);
/* out = a + b, n elements (residual adds; backward is the identity) */
int la_add_f32(void* stream, const float* a, const float* b, float* out, long long n);
/* dx = dy where y > 0 else 0 (nn.ReLU behind lin1 / class_mlp: transformer.py:164, mask_decoder.py:797) */
int la_relu_bwd_f32(void* stream, const float* dy, const float* y, float* dx, long long n);
/* y = GELU(x) (exact erf form, nn.GELU()) and dx = dy * GELU'(x): the activation of AttentionMLPBlock's MLP
 * (common.py:19-37,151-184), kept apart from its GEMM so that the pre-activation is available to the backward pass */
int la_gelu_f32(void* stream, const float* x, float* y, long long n);
int la_gelu_bwd_f32(void* stream, const float* dy, const float* x, float* dx, long long n);
/* out[c][r] = bf16(in[r][c]), in fp32 or bf16 [rows, cols] (row stride ld_in), out bf16 [cols, ld_out] with columns
 * [rows, ld_out) zero: the K-major operands of the weight-gradient GEMM dW[N, K] = dY^T[N, rows] @ X^T[K, rows]^T. */
int la_cast_transpose_bf16(void* stream, const void* in, int in_dtype, long long ld_in, void* out, long long ld_out,
                           long long rows, int cols);
/* One pass over an output gradient dy fp32 [rows, cols] (row stride ld) for everything a Linear's backward needs from
 * it (each output optional): out_bf16 [rows, cols] = bf16(dy) (data-gradient GEMM operand), out_t_bf16 [cols, ld_t] =
 * its transpose, columns [rows, ld_t) zero (weight-gradient GEMM operand), colsum fp32 [cols] = column sums (bias
 * gradient; overwritten). */
int la_grad_prep_bf16(void* stream, const float* dy, long long ld, void* out_bf16, void* out_t_bf16, long long ld_t,
                      float* colsum, long long rows, int cols);
/* out[(r / row_div) % b_mod][:] (+)= dy[r][:]: bias gradients (row_div = 1, b_mod = 1 -> column sum), gradients of
 * broadcast addends (class codes, no_sparse_embedding, positional tables).  accumulate = 0 zeroes out first. */
int la_bcast_reduce_f32(void* stream, const float* dy, float* out, long long rows, int d, long long row_div,
                        long long b_mod, int accumulate);
/* y = act(LayerNorm(x) * gamma + beta) per row of fp32 [rows, d] (biased variance; gamma = beta = NULL: no affine),
 * act = LA_ACT_NONE / LA_ACT_GELU.  nn.LayerNorm and LayerNorm2d (+ the GELU behind it): common.py:42-54,
 * transformer.py:298-329, prompt_encoder.py:61-69, mask_decoder.py:206-255, build_lam.py:150-171.  d <= 1024. */
int la_layernorm_f32(void* stream, const float* x, const float* gamma, const float* beta, float eps, int act, float* y,
                     long long rows, int d);
/* its backward: dx overwritten, dgamma / dbeta ACCUMULATED (may be NULL together); statistics are recomputed from x */
int la_layernorm_f32_bwd(void* stream, const float* x, const float* gamma, const float* beta, float eps, int act,
                         const float* dy, float* dx, float* dgamma, float* dbeta, long long rows, int d);
/* out = softmax(scale q k^T) v per (sequence, head) on fp32 [n_seq * nq | nk, heads * head_dim] rows; lse fp32
 * [n_seq, heads, nq] = log-sum-exp of the scaled scores (kept for the backward pass).  common.py:97-148 after the
 * projections (the reference's masks are no-ops).  head_dim <= 64 and a multiple of 4 (16-byte row loads). */
int la_attention_f32(void* stream, const float* q, const float* k, const float* v, float* out, float* lse,
                     long long n_seq, int nq, int nk, int heads, int head_dim, float scale);
/* its backward: dq / dk / dv overwritten; delta fp32 [n_seq, heads, nq] is scratch (rowsum(dout * out)) */
int la_attention_f32_bwd(void* stream, const float* q, const float* k, const float* v, const float* out, const float* lse,
                         const float* dout, float* delta, float* dq, float* dk, float* dv, long long n_seq, int nq, int nk,
                         int heads, int head_dim, float scale);
/* adjoint of la_im2col_3x3: dcol fp32 [n_img*H*W, 9*C] -> dx fp32 [n_img*H*W, C] (spatial_convs, neck 3x3) */
int la_col2im_3x3_f32(void* stream, const float* dcol, float* dx, long long n_img, int height, int width, int channels);
/* la_mask_downscale with the weights in DEVICE memory (they change every step): weights = 332 floats, the flattened
 * parameters mask_downscaling.{0.weight, 0.bias, 1.weight, 1.bias, 3.weight, 3.bias, 4.weight, 4.bias} in this order
 * (prompt_encoder.py:61-69) -> out fp32 [n_seq, H/4, W/4, 16] */
int la_mask_downscale_dev(void* stream, const float* masks, const float* weights, float eps1, float eps2, float* out,
                          long long n_seq, int height, int width);
/* parameter gradients of the above (the masks are inputs, not differentiated): dweights 332 floats, overwritten */
int la_mask_downscale_bwd(void* stream, const float* masks, const float* weights, float eps1, float eps2,
                          const float* dout, float* dweights, long long n_seq, int height, int width);
/* adjoint of la_resize_bilinear: dout [n, out_h, out_w, c] -> din [n, in_h, in_w, c] (overwritten) */
int la_resize_bilinear_bwd(void* stream, const float* dout, float* din, long long n, int in_h, int in_w, int out_h,
                           int out_w, int channels);
/* src[s, t, :] = feat[s / n_classes, t, :] + (dense != NULL && (mask_flags == NULL || mask_flags[s]) ? dense[s, t, :]
 * : alt[:]) in fp32 -- prompt_encoder.py:783-803 with alt = not_a_mask_embed (masks given) / no_mask_embed (none) */
int la_src_combine_f32(void* stream, const float* feat, const float* dense, const unsigned char* mask_flags,
                       const float* alt, float* out, long long n_seq, int tokens, int d, int n_classes);
/* its backward: dfeat [n_seq / n_classes, tokens, d] = sum over the classes, ddense (may be NULL) = dsrc of the
 * sequences that used dense (else 0), dalt_rows (may be NULL) [n_seq / n_classes, tokens, d] = sum over the sequences
 * that used alt (column-summed into the embedding's gradient by la_bcast_reduce_f32) */
int la_src_combine_bwd(void* stream, const float* dsrc, const unsigned char* mask_flags, int has_dense, float* dfeat,
                       float* ddense, float* dalt_rows, long long n_seq, int tokens, int d, int n_classes);
/* gradients of la_embed_sparse's learned vectors: dtable [4, d] (point_embeddings.0-3), dnot_a_point [d], both
 * overwritten; dout [n_seq, n, d] with n as in la_embed_sparse (prompt_encoder.py:83-114) */
int la_embed_sparse_bwd(void* stream, const float* point_labels, int n_points, const float* box_flags, int n_boxes,
                        int has_points, const float* dout, float* dtable, float* dnot_a_point, long long n_seq, int d);
/* out[s, :] = mean over the seg_rows rows of segment s (fused.mean(dim=(2, 3)), prompt_encoder.py:733-735) + backward */
int la_segment_mean_f32(void* stream, const float* x, float* out, long long n_seg, int seg_rows, int d);
int la_segment_mean_bwd(void* stream, const float* dout, float* dx, long long n_seg, int seg_rows, int d);
/* backward of la_masked_mean: demb[b, m, c, :] = flags[b, m, c] * dout[b, c, :] / max(sum_m flags, 1) */
int la_masked_mean_bwd(void* stream, const float* dout, const unsigned char* flags, float* demb, int batch, int examples,
                       int classes, int d);
/* backward of la_classify: dx fp32 [batch*pixels, dk], dcls fp32 [batch, classes, dk] (mask_decoder.py:309) */
int la_classify_bwd(void* stream, const float* dlogits, const void* x, const float* cls, float* dx, float* dcls, int batch,
                    long long pixels, int classes, int dk);
/* adjoint of la_postprocess_masks: dout [batch, classes, out_h, out_w] -> din [batch, classes, low_h, low_w]; padded
 * pixels and flag_gts-masked classes carry no gradient (lam.py:383-453) */
int la_postprocess_masks_bwd(void* stream, const float* dout, const int* sizes, const unsigned char* flag_gts, float* din,
                             int batch, int classes, int low_h, int low_w, int image_size, int out_h, int out_w);
/* torch.optim.AdamW step t = step (>= 1) over one flat fp32 bucket: g = grads * grad_scale; decoupled weight decay
 * (experiment/run.py:172-200 builds AdamW over get_learnable_params) */
int la_adamw_f32(void* stream, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                 float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale);
/* the same update with the per-step scalars {1 - beta1^t, sqrt(1 - beta2^t), lr} read from device memory (fp32 [3]):
 * the launch parameters stay constant, so a captured CUDA graph of the whole training step can be replayed as t grows
 * and the learning-rate schedule (run.py: constant_with_warmup) moves */
int la_adamw_f32_dev(void* stream, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                     float beta1, float beta2, float eps, float weight_decay, const float* step_scalars,
                     float grad_scale);

#ifdef __cplusplus
}
#endif
#endif /* LABELANYTHING_B200_H */
