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

#ifdef __cplusplus
}
#endif
#endif /* LABELANYTHING_B200_H */
