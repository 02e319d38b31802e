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

/* ---- library management ------------------------------------------------------------------------- */
const char* la_last_error(void); /* thread-local message of the last failing call */
int la_version(void);
int la_device_check(void); /* LA_OK iff the current device is sm_100 */

/* ---- dense contraction -------------------------------------------------------------------------- */
/* out[M,N] = act(a[M,K] @ w[N,K]^T + bias[N]);  a, w bf16 with K contiguous (nn.Linear weight layout),
 * fp32 accumulation on tcgen05 tensor cores, out bf16 or fp32.  lda/ldw/ldo are row strides in elements.
 * Replaces every nn.Linear / 1x1 Conv2d / im2col'd Conv2d / stride==kernel ConvTranspose2d on the path:
 *   label_anything/models/image_encoder.py:227-228,242,253,402-410; common.py:28-37,82-85,108-110,146;
 *   build_lam.py:150-171; mask_decoder.py:206-255,776-804. */
int la_gemm_bf16(void* stream, const void* a, long long lda, const void* w, long long ldw, const float* bias,
                 void* out, long long ldo, int out_dtype, int M, int N, int K, int act);

/* ---- fused attention ----------------------------------------------------------------------------- */
/* Multi-head self-attention, head_dim 64, over n_seq sequences of seq_len tokens stored as consecutive rows
 * of the projection matrices q [rows_total, ld_q] and kv [rows_total, ld_kv] (bf16; they may be the same
 * packed qkv buffer).  Head h reads q at column q_off + 64*h and k / v at columns {k,v}_off + 64*h of kv.
 * softmax(scale * q k^T + bias) v -> out [.., ld_out] bf16 at columns 64*h.
 * Optional decomposed relative-position bias (bias_h/bias_w != NULL): fp32 tables [rows_total][n_heads][ldb]
 * with table[row][h][grid_hw-1 - q_pos + k_pos] = q_row(h) . rel_pos[q_pos - k_pos + grid_hw-1], i.e. the
 * product of the head's q rows with the REVERSED rel_pos table (computed with la_gemm_bf16); grid_hw = 64
 * (global blocks, seq_len 4096) or 14 (windowed blocks, seq_len 196).
 * out_mode 0: out row = sequence*seq_len + token.  out_mode 1: window un-partition: sequence = image*nwin^2
 * + window, token (ty,tx) of window (wy,wx) goes to image row (wy*14+ty)*img_hw + wx*14+tx, padded
 * positions are dropped.
 *   label_anything/models/image_encoder.py:239-255,282-304,340-376; transformers modeling_vit.py:199-250 */
int la_attention_bf16(void* stream, const void* q, long long ld_q, int q_off, const void* kv, long long ld_kv,
                      int k_off, int v_off, long long rows_total, int n_seq, int seq_len, int n_heads, float scale,
                      const float* bias_h, const float* bias_w, int ldb, int grid_hw, void* out, long long ld_out,
                      int out_mode, int nwin, int img_hw);

/* ---- streaming row kernels ----------------------------------------------------------------------- */
/* x = x_in[(row % x_mod) if x_mod > 0 else row] + delta[row]  (fp32 + bf16); optionally stored to x_out (may
 * alias x_in); y = act(LayerNorm(x) * gamma + beta) (biased variance, `eps`), or a plain cast when
 * gamma == NULL.  Outputs (each optional): y_out as bf16 / fp32 (y_dtype), y2_out = y in fp32,
 * ype_out = bf16(y + pe[row % pe_mod]) (positional table folded in for the next projection).  Row remapping:
 *   map_mode 0: identity.
 *   map_mode 1: `rows` counts OUTPUT rows in window-partitioned order (image, wy, wx, ty, tx) for win x win
 *               windows, nwin per side, over an hw x hw grid; rows that fall in the zero padding are written
 *               as zeros (F.pad happens after norm1).  image_encoder.py:183-187,258-279
 *   map_mode 2: drop token 0 (CLS) of every seq_len-token sequence.  build_encoder.py:98
 *   image_encoder.py:181-197; common.py:42-54,183-184; transformer.py:308-327; mask_decoder.py:214-215,250-254;
 *   transformers modeling_vit.py:325-346,416 */
int la_add_layernorm(void* stream, const float* x_in, long long x_mod, const void* delta, float* x_out,
                     const float* gamma, const float* beta, float eps, int act, void* y_out, int y_dtype,
                     float* y2_out, const float* pe, long long pe_mod, void* ype_out, long long rows, int d,
                     int map_mode, int seq_len, int win, int nwin, int hw);

/* out[s, :] = mean over the rows_per_seq rows of sequence s of LayerNorm(x_in + delta): the last
 * image-token LayerNorm of the prompt encoder's two-way transformer fused with the spatial average pooling
 * (the normalised tokens are never written).  partial_ws: fp32 scratch [n_seq * slices * d]; deterministic.
 *   label_anything/models/transformer.py:326-327 + prompt_encoder.py:733-735 */
int la_add_layernorm_meanpool(void* stream, const float* x_in, const void* delta, const float* gamma,
                              const float* beta, float eps, long long n_seq, int rows_per_seq, int d,
                              float* partial_ws, int slices, float* out);

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

#ifdef __cplusplus
}
#endif
#endif /* LABELANYTHING_B200_H */
