// labelanything_b200 — kernels of the training step: the backward passes of the ops between the
// (pre-computed) image embeddings and the loss -- Lam.neck, the visual prompt encoder, the mask decoder and
// postprocess_masks -- plus the fp32 forward variants the training path needs (activations are kept in fp32 between
// ops, GEMM operands are rounded to bf16 like the inference path) and the optimiser update.
//
// Reference: label_anything/experiment/run.py:359-361,425-590 (backward + optimiser step of the training loop) over
// label_anything/models/{prompt_encoder,transformer,mask_decoder,common,lam}.py.  The configuration this serves
// (BASELINE.json configs[3]: MAE-L-256 on pre-computed embeddings, 2-way 5-shot) is ~0.1 TFLOP per step, so every
// kernel here is a plain, HBM / latency-bound CUDA-core kernel; the contractions (dgrad / wgrad of every Linear and
// convolution) run on the tcgen05 GEMM (la_gemm_bf16) with operands transposed by la_cast_transpose_bf16.
#include <math.h>

#include "la_common.cuh"

namespace la {
namespace {

unsigned train_grid(long long items, int block = 256) {
  long long blocks = (items + block - 1) / block;
  const long long cap = 16ll * sm_count();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<unsigned>(blocks);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// exact GELU and its derivative (nn.GELU(): 0.5 x (1 + erf(x / sqrt 2)))
__host__ __device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__host__ __device__ __forceinline__ float gelu_grad(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
}

// PyTorch's align_corners=False source index (aten/src/ATen/native/UpSample.h, area_pixel_compute_source_index)
__device__ __forceinline__ void tap(int dst, float scale, int in_size, int& i0, int& i1, float& lam) {
  float src = (dst + 0.5f) * scale - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = static_cast<int>(src);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  lam = src - i0;
}

// ------------------------------------------------------------------------------------------------------------------
// elementwise
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                        long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}

// x = hi + lo + O(2^-17 |x|): the two bf16 terms of the split-operand ("bf16x3") GEMM mode
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi,
                                                         __nv_bfloat16* __restrict__ lo, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float x = in[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(x - __bfloat162float(h));
  }
}

// three bf16 terms (24 mantissa bits): the "bf16x6" mode, fp32-level products from six tensor-core GEMMs
__global__ void __launch_bounds__(256) split3_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi,
                                                          __nv_bfloat16* __restrict__ mid, __nv_bfloat16* __restrict__ lo,
                                                          long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float x = in[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16_rn(r1);
    hi[i] = h;
    mid[i] = m;
    lo[i] = __float2bfloat16_rn(r1 - __bfloat162float(m));
  }
}

__global__ void __launch_bounds__(256) add_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                  float* __restrict__ out, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = a[i] + b[i];
}

__global__ void __launch_bounds__(256) gelu_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    y[i] = gelu_f(x[i]);
}

__global__ void __launch_bounds__(256) gelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                       float* __restrict__ dx, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dx[i] = dy[i] * gelu_grad(x[i]);
}

__global__ void __launch_bounds__(256) relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                       float* __restrict__ dx, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dx[i] = y[i] > 0.f ? dy[i] : 0.f;
}

// out[c][r] = bf16(in[r][c]); columns [rows, ld_out) of every output row are zero (the contraction dimension of the
// weight-gradient GEMM must be a multiple of 8)
template <typename T>
__global__ void __launch_bounds__(256) cast_transpose_kernel(const T* __restrict__ in, long long ld_in,
                                                             __nv_bfloat16* __restrict__ out, long long ld_out,
                                                             long long rows, int cols) {
  __shared__ float tile[32][33];
  const long long r0 = static_cast<long long>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long r = r0 + ty + 8 * k;
    const int c = c0 + tx;
    float v = 0.f;
    if (r < rows && c < cols) v = static_cast<float>(in[r * ld_in + c]);
    tile[ty + 8 * k][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k;
    const long long r = r0 + tx;
    if (c < cols && r < ld_out) out[c * ld_out + r] = __float2bfloat16_rn(tile[tx][ty + 8 * k]);
  }
}

// One pass over an output gradient dy fp32 [rows, cols] for the three things a Linear's backward needs from it: the bf16
// copy (A operand of the data-gradient GEMM), the bf16 transpose padded to 8 rows (A operand of the weight-gradient
// GEMM) and the column sums (bias gradient).  Separately these were three kernels reading dy three times.
constexpr int PREP_TILES = 16;   // 32-row tiles per CTA: 512 rows -> one atomic per column per CTA
__global__ void __launch_bounds__(256)
grad_prep_kernel(const float* __restrict__ dy, long long ld, __nv_bfloat16* __restrict__ out16,
                 __nv_bfloat16* __restrict__ outT, long long ldT, float* __restrict__ colsum, long long rows, int cols) {
  __shared__ float tile[32][33];
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c0 = blockIdx.y * 32;
  const long long rbase = static_cast<long long>(blockIdx.x) * (32 * PREP_TILES);
  float csum = 0.f;
  for (int t = 0; t < PREP_TILES; ++t) {
    const long long r0 = rbase + 32ll * t;
    if (r0 >= ldT) break;                 // ldT >= rows: the padding rows of the transpose are written (as zeros) too
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long r = r0 + ty + 8 * k;
      const int c = c0 + tx;
      float v = 0.f;
      if (r < rows && c < cols) {
        v = dy[r * ld + c];
        if (out16) out16[r * cols + c] = __float2bfloat16_rn(v);
      }
      csum += v;
      tile[ty + 8 * k][tx] = v;
    }
    __syncthreads();
    if (outT) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = c0 + ty + 8 * k;
        const long long r = r0 + tx;
        if (c < cols && r < ldT) outT[c * ldT + r] = __float2bfloat16_rn(tile[tx][ty + 8 * k]);
      }
    }
    __syncthreads();
  }
  if (colsum) {
    part[ty][tx] = csum;
    __syncthreads();
    if (ty == 0 && c0 + tx < cols) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) a += part[k][tx];
      atomicAdd(colsum + c0 + tx, a);
    }
  }
}

// out[(r / row_div) % b_mod][c] += sum of dy[r][c] over the rows r of one chunk (runs of equal targets are added
// locally, one atomic per run)
constexpr int RED_ROWS = 64;
__global__ void __launch_bounds__(256) bcast_reduce_kernel(const float* __restrict__ dy, float* __restrict__ out,
                                                           long long rows, int d, long long row_div, long long b_mod) {
  const long long n_chunks = (rows + RED_ROWS - 1) / RED_ROWS;
  const long long total = n_chunks * d;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % d);
    const long long r0 = (i / d) * RED_ROWS;
    const long long r1 = r0 + RED_ROWS < rows ? r0 + RED_ROWS : rows;
    long long cur = (r0 / row_div) % b_mod;
    float acc = 0.f;
    for (long long r = r0; r < r1; ++r) {
      const long long t = (r / row_div) % b_mod;
      if (t != cur) {
        atomicAdd(out + cur * d + c, acc);
        acc = 0.f;
        cur = t;
      }
      acc += dy[r * d + c];
    }
    atomicAdd(out + cur * d + c, acc);
  }
}

// plain column sum (b_mod == 1: bias gradients over up to S*T rows): 32 columns x 8 row-lanes per CTA over a chunk of
// COLSUM_ROWS rows, shared-memory reduction over the row-lanes, one atomic per column per CTA (the generic kernel above
// issued one atomic per column per 64 rows: 844 per address for the 54 000 image-token rows of a training step)
constexpr int COLSUM_ROWS = 512;
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dy, float* __restrict__ out, long long rows,
                                                     int d) {
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const long long r0 = static_cast<long long>(blockIdx.y) * COLSUM_ROWS;
  const long long r1 = r0 + COLSUM_ROWS < rows ? r0 + COLSUM_ROWS : rows;
  float a0 = 0.f, a1 = 0.f;
  if (c < d) {
    long long r = r0 + ty;
    for (; r + 8 < r1; r += 16) {
      a0 += dy[r * d + c];
      a1 += dy[(r + 8) * d + c];
    }
    if (r < r1) a0 += dy[r * d + c];
  }
  part[ty][tx] = a0 + a1;
  __syncthreads();
  if (ty == 0 && c < d) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += part[k][tx];
    atomicAdd(out + c, s);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// LayerNorm (nn.LayerNorm / LayerNorm2d on token-major rows: biased variance) with an optional GELU behind it
// ------------------------------------------------------------------------------------------------------------------
template <int NV>   // NV = ceil(d / 32) rounded up to a power of two, <= 32
__global__ void __launch_bounds__(256) layernorm_f32_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps, int act,
                                                            float* __restrict__ y, long long rows, int d) {
  const int lane = threadIdx.x & 31;
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const float inv_d = 1.f / d;
  for (long long r = warp; r < rows; r += n_warps) {
    float v[NV];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = lane + 32 * k;
      v[k] = c < d ? x[r * d + c] : 0.f;
      s += v[k];
    }
    const float mean = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = lane + 32 * k;
      const float t = c < d ? v[k] - mean : 0.f;
      q += t * t;
    }
    const float rstd = 1.f / sqrtf(warp_sum(q) * inv_d + eps);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = lane + 32 * k;
      if (c < d) {
        float o = (v[k] - mean) * rstd;
        if (gamma) o = o * gamma[c] + beta[c];
        if (act == LA_ACT_GELU) o = gelu_f(o);
        y[r * d + c] = o;
      }
    }
  }
}

// dx = d/dx of act(gamma * xhat + beta) against dy; dgamma / dbeta accumulated (atomics, one per column per CTA)
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_f32_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                         float eps, int act, const float* __restrict__ dy, float* __restrict__ dx,
                         float* __restrict__ dgamma, float* __restrict__ dbeta, long long rows, int d) {
  extern __shared__ float s_acc[];   // [2][d]
  for (int i = threadIdx.x; i < 2 * d; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const float inv_d = 1.f / d;
  float ag[NV], ab[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) ag[k] = ab[k] = 0.f;
  for (long long r = warp; r < rows; r += n_warps) {
    float v[NV];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = lane + 32 * k;
      v[k] = c < d ? x[r * d + c] : 0.f;
      s += v[k];
    }
    const float mean = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = lane + 32 * k;
      const float t = c < d ? v[k] - mean : 0.f;
      q += t * t;
    }
    const float rstd = 1.f / sqrtf(warp_sum(q) * inv_d + eps);
    float dxh[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = lane + 32 * k;
      dxh[k] = 0.f;
      if (c < d) {
        const float xh = (v[k] - mean) * rstd;
        const float g = gamma ? gamma[c] : 1.f;
        float go = dy[r * d + c];
        if (act == LA_ACT_GELU) go *= gelu_grad(xh * g + (gamma ? beta[c] : 0.f));
        ag[k] += go * xh;
        ab[k] += go;
        dxh[k] = go * g;
        s1 += dxh[k];
        s2 += dxh[k] * xh;
        v[k] = xh;
      }
    }
    s1 = warp_sum(s1) * inv_d;
    s2 = warp_sum(s2) * inv_d;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = lane + 32 * k;
      if (c < d) dx[r * d + c] = rstd * (dxh[k] - s1 - v[k] * s2);
    }
  }
  if (dgamma) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = lane + 32 * k;
      if (c < d) {
        atomicAdd(s_acc + c, ag[k]);
        atomicAdd(s_acc + d + c, ab[k]);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
      atomicAdd(dgamma + i, s_acc[i]);
      atomicAdd(dbeta + i, s_acc[d + i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// token attention in fp32 (common.Attention after the projections): one warp per (sequence, head, query), lanes over keys
// ------------------------------------------------------------------------------------------------------------------
struct AttnParams {
  const float *q, *k, *v;     // [n_seq * nq | nk, heads * dh]
  float* out;                 // [n_seq * nq, heads * dh]
  float* lse;                 // [n_seq, heads, nq]
  const float *dout;
  float *delta;               // [n_seq, heads, nq]: rowsum(dout * out)
  float *dq, *dk, *dv;
  long long n_seq;
  int nq, nk, heads, dh;
  float scale;
};

// rows are read as 16-byte vectors (head_dim % 4 == 0, 16-byte aligned rows): with lanes over keys every lane walks its
// own row, and scalar loads made these kernels LSU-bound (32 sectors per 4-byte load instruction)
template <int DH>
__device__ __forceinline__ void load_row(float (&r)[DH], const float* __restrict__ p, int dh) {
#pragma unroll
  for (int e = 0; e < DH; e += 4) {
    if (e < dh) {
      const float4 t = *reinterpret_cast<const float4*>(p + e);
      r[e] = t.x; r[e + 1] = t.y; r[e + 2] = t.z; r[e + 3] = t.w;
    } else {
      r[e] = r[e + 1] = r[e + 2] = r[e + 3] = 0.f;
    }
  }
}
template <int DH>
__device__ __forceinline__ void store_row(float* __restrict__ p, const float (&r)[DH], int dh) {
#pragma unroll
  for (int e = 0; e < DH; e += 4)
    if (e < dh) *reinterpret_cast<float4*>(p + e) = make_float4(r[e], r[e + 1], r[e + 2], r[e + 3]);
}
template <int DH>
__device__ __forceinline__ float dot_row(const float (&a)[DH], const float (&b)[DH]) {
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < DH; ++e) s = fmaf(a[e], b[e], s);   // entries past head_dim are zero
  return s;
}

// ROW = false: one warp per row, lanes over the inner side, shuffle reductions (long inner side);
// ROW = true : one thread per row, serial inner loop (inner side <= 32: the image->token attention of the two-way
//              blocks has 900 queries against 1..30 keys per (sequence, head) -- lanes over keys would idle and every
//              output would need a shuffle reduction; the rows of a warp share their (sequence, head), so the
//              inner-side rows are broadcast loads)
template <int DH, bool ROW>
__global__ void __launch_bounds__(128) attn_f32_fwd_kernel(const AttnParams p) {
  const int lane = ROW ? 0 : (threadIdx.x & 31);
  const int step = ROW ? 1 : 32;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nthr = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long total = p.n_seq * p.heads * p.nq;
  const int ld = p.heads * p.dh;
  for (long long w = ROW ? tid : (tid >> 5); w < total; w += ROW ? nthr : (nthr >> 5)) {
    const int i = static_cast<int>(w % p.nq);
    const int h = static_cast<int>((w / p.nq) % p.heads);
    const long long s = w / (static_cast<long long>(p.nq) * p.heads);
    float q[DH], o[DH], kk[DH];
    load_row<DH>(q, p.q + (s * p.nq + i) * ld + h * p.dh, p.dh);
#pragma unroll
    for (int e = 0; e < DH; ++e) {
      q[e] *= p.scale;
      o[e] = 0.f;
    }
    const float* kb = p.k + s * p.nk * ld + h * p.dh;
    const float* vb = p.v + s * p.nk * ld + h * p.dh;
    float m = -INFINITY;
    for (int j = lane; j < p.nk; j += step) {
      load_row<DH>(kk, kb + static_cast<long long>(j) * ld, p.dh);
      m = fmaxf(m, dot_row<DH>(q, kk));
    }
    if (!ROW) m = warp_max(m);
    float l = 0.f;
    for (int j = lane; j < p.nk; j += step) {
      load_row<DH>(kk, kb + static_cast<long long>(j) * ld, p.dh);
      const float pj = expf(dot_row<DH>(q, kk) - m);
      l += pj;
      load_row<DH>(kk, vb + static_cast<long long>(j) * ld, p.dh);
#pragma unroll
      for (int e = 0; e < DH; ++e) o[e] = fmaf(pj, kk[e], o[e]);
    }
    if (!ROW) {
      l = warp_sum(l);
#pragma unroll
      for (int e = 0; e < DH; ++e)
        if (e < p.dh) o[e] = warp_sum(o[e]);
    }
    const float inv = 1.f / l;
#pragma unroll
    for (int e = 0; e < DH; ++e) o[e] *= inv;
    if (lane == 0) {
      store_row<DH>(p.out + (s * p.nq + i) * ld + h * p.dh, o, p.dh);
      p.lse[w] = m + logf(l);
    }
  }
}

// dq (and delta = rowsum(dout * out), consumed by the dk / dv kernel)
template <int DH, bool ROW>
__global__ void __launch_bounds__(128) attn_f32_bwd_q_kernel(const AttnParams p) {
  const int lane = ROW ? 0 : (threadIdx.x & 31);
  const int step = ROW ? 1 : 32;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nthr = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long total = p.n_seq * p.heads * p.nq;
  const int ld = p.heads * p.dh;
  for (long long w = ROW ? tid : (tid >> 5); w < total; w += ROW ? nthr : (nthr >> 5)) {
    const int i = static_cast<int>(w % p.nq);
    const int h = static_cast<int>((w / p.nq) % p.heads);
    const long long s = w / (static_cast<long long>(p.nq) * p.heads);
    const long long row = (s * p.nq + i) * ld + h * p.dh;
    float q[DH], go[DH], dq[DH], kk[DH];
    load_row<DH>(q, p.q + row, p.dh);
    load_row<DH>(go, p.dout + row, p.dh);
    load_row<DH>(kk, p.out + row, p.dh);
    const float dl = dot_row<DH>(go, kk);
#pragma unroll
    for (int e = 0; e < DH; ++e) {
      q[e] *= p.scale;
      dq[e] = 0.f;
    }
    const float lse = p.lse[w];
    const float* kb = p.k + s * p.nk * ld + h * p.dh;
    const float* vb = p.v + s * p.nk * ld + h * p.dh;
    for (int j = lane; j < p.nk; j += step) {
      load_row<DH>(kk, vb + static_cast<long long>(j) * ld, p.dh);
      const float dp = dot_row<DH>(go, kk);
      load_row<DH>(kk, kb + static_cast<long long>(j) * ld, p.dh);
      const float ds = expf(dot_row<DH>(q, kk) - lse) * (dp - dl) * p.scale;
#pragma unroll
      for (int e = 0; e < DH; ++e) dq[e] = fmaf(ds, kk[e], dq[e]);
    }
    if (!ROW) {
#pragma unroll
      for (int e = 0; e < DH; ++e)
        if (e < p.dh) dq[e] = warp_sum(dq[e]);
    }
    if (lane == 0) {
      store_row<DH>(p.dq + row, dq, p.dh);
      p.delta[w] = dl;
    }
  }
}

// dk, dv: rows are the keys, the inner side are the queries
template <int DH, bool ROW>
__global__ void __launch_bounds__(128) attn_f32_bwd_kv_kernel(const AttnParams p) {
  const int lane = ROW ? 0 : (threadIdx.x & 31);
  const int step = ROW ? 1 : 32;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nthr = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long total = p.n_seq * p.heads * p.nk;
  const int ld = p.heads * p.dh;
  for (long long w = ROW ? tid : (tid >> 5); w < total; w += ROW ? nthr : (nthr >> 5)) {
    const int j = static_cast<int>(w % p.nk);
    const int h = static_cast<int>((w / p.nk) % p.heads);
    const long long s = w / (static_cast<long long>(p.nk) * p.heads);
    const long long row = (s * p.nk + j) * ld + h * p.dh;
    float k[DH], v[DH], dk[DH], dv[DH], qq[DH], gg[DH];
    load_row<DH>(k, p.k + row, p.dh);
    load_row<DH>(v, p.v + row, p.dh);
#pragma unroll
    for (int e = 0; e < DH; ++e) dk[e] = dv[e] = 0.f;
    const float* qb = p.q + s * p.nq * ld + h * p.dh;
    const float* gb = p.dout + s * p.nq * ld + h * p.dh;
    const long long stat = (s * p.heads + h) * p.nq;
    for (int i = lane; i < p.nq; i += step) {
      load_row<DH>(qq, qb + static_cast<long long>(i) * ld, p.dh);
      load_row<DH>(gg, gb + static_cast<long long>(i) * ld, p.dh);
      const float pij = expf(dot_row<DH>(qq, k) * p.scale - p.lse[stat + i]);
      const float ds = pij * (dot_row<DH>(gg, v) - p.delta[stat + i]) * p.scale;
#pragma unroll
      for (int e = 0; e < DH; ++e) {
        dv[e] = fmaf(pij, gg[e], dv[e]);
        dk[e] = fmaf(ds, qq[e], dk[e]);
      }
    }
    if (!ROW) {
#pragma unroll
      for (int e = 0; e < DH; ++e) {
        if (e < p.dh) {
          dk[e] = warp_sum(dk[e]);
          dv[e] = warp_sum(dv[e]);
        }
      }
    }
    if (lane == 0) {
      store_row<DH>(p.dk + row, dk, p.dh);
      store_row<DH>(p.dv + row, dv, p.dh);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// col2im of the 3x3 / pad 1 im2col (la_im2col_3x3): dx[p, ci] = sum over taps of dcol[p - offset(tap), tap, ci]
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) col2im_3x3_kernel(const float* __restrict__ dcol, float* __restrict__ dx,
                                                         long long n_img, int h, int w, int c) {
  const long long total = n_img * h * w * c;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % c);
    long long r = i / c;
    const int x = static_cast<int>(r % w);
    r /= w;
    const int y = static_cast<int>(r % h);
    const long long img = r / h;
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y - ky + 1;     // the output pixel whose tap (ky, kx) reads (y, x)
      if (yy < 0 || yy >= h) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = x - kx + 1;
        if (xx < 0 || xx >= w) continue;
        acc += dcol[((img * h + yy) * w + xx) * (9ll * c) + (ky * 3 + kx) * c + ci];
      }
    }
    dx[i] = acc;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// mask downscaling (Conv2d(1,4,2,2) -> LayerNorm2d -> GELU -> Conv2d(4,16,2,2) -> LayerNorm2d -> GELU) with the
// weights in device memory, and its parameter gradients.  Packed weights (332 floats): w0[4][4], b0[4], g1[4], be1[4],
// w3[16][16] ([c2][c1 * 4 + ky * 2 + kx]), b3[16], g2[16], be2[16] = the flattened parameters in module order.
// ------------------------------------------------------------------------------------------------------------------
constexpr int MD_W0 = 0, MD_B0 = 16, MD_G1 = 20, MD_BE1 = 24, MD_W3 = 28, MD_B3 = 284, MD_G2 = 300, MD_BE2 = 316,
              MD_N = 332;

struct MdFwd {
  float in[4][4];
  float xh1[4][4], pre1[4][4], a1[4][4], rstd1[4];   // [sub-block][c1]
  float xh2[16], pre2[16], rstd2;
};

__device__ __forceinline__ void md_forward(const float* __restrict__ W, float eps1, float eps2, MdFwd& f, float* out16) {
#pragma unroll
  for (int sb = 0; sb < 4; ++sb) {
    const int sy = sb >> 1, sx = sb & 1;
    float t[4], mean = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float acc = W[MD_B0 + c];
#pragma unroll
      for (int k = 0; k < 4; ++k) acc = fmaf(W[MD_W0 + c * 4 + k], f.in[2 * sy + (k >> 1)][2 * sx + (k & 1)], acc);
      t[c] = acc;
      mean += acc;
    }
    mean *= 0.25f;
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) var += (t[c] - mean) * (t[c] - mean);
    const float rstd = rsqrtf(var * 0.25f + eps1);
    f.rstd1[sb] = rstd;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      f.xh1[sb][c] = (t[c] - mean) * rstd;
      f.pre1[sb][c] = f.xh1[sb][c] * W[MD_G1 + c] + W[MD_BE1 + c];
      f.a1[sb][c] = gelu_f(f.pre1[sb][c]);
    }
  }
  float t2[16], mean = 0.f;
#pragma unroll
  for (int c2 = 0; c2 < 16; ++c2) {
    float acc = W[MD_B3 + c2];
#pragma unroll
    for (int c1 = 0; c1 < 4; ++c1)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc = fmaf(W[MD_W3 + c2 * 16 + c1 * 4 + k], f.a1[k][c1], acc);
    t2[c2] = acc;
    mean += acc;
  }
  mean *= (1.f / 16.f);
  float var = 0.f;
#pragma unroll
  for (int c = 0; c < 16; ++c) var += (t2[c] - mean) * (t2[c] - mean);
  f.rstd2 = rsqrtf(var * (1.f / 16.f) + eps2);
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    f.xh2[c] = (t2[c] - mean) * f.rstd2;
    f.pre2[c] = f.xh2[c] * W[MD_G2 + c] + W[MD_BE2 + c];
    out16[c] = gelu_f(f.pre2[c]);
  }
}

__device__ __forceinline__ void md_load(const float* __restrict__ masks, long long i, int oh, int ow, int Hm, int Wm,
                                        MdFwd& f) {
  const int ox = static_cast<int>(i % ow);
  const long long r = i / ow;
  const int oy = static_cast<int>(r % oh);
  const long long s = r / oh;
  const float* src = masks + (s * Hm + 4 * oy) * static_cast<long long>(Wm) + 4 * ox;
#pragma unroll
  for (int y = 0; y < 4; ++y)
#pragma unroll
    for (int x = 0; x < 4; ++x) f.in[y][x] = src[static_cast<long long>(y) * Wm + x];
}

__global__ void __launch_bounds__(256) mask_downscale_dev_kernel(const float* __restrict__ masks,
                                                                 const float* __restrict__ weights, float eps1, float eps2,
                                                                 float* __restrict__ out, long long n_seq, int Hm, int Wm) {
  __shared__ float W[MD_N];
  for (int i = threadIdx.x; i < MD_N; i += blockDim.x) W[i] = weights[i];
  __syncthreads();
  const int oh = Hm >> 2, ow = Wm >> 2;
  const long long total = n_seq * oh * ow;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    MdFwd f;
    md_load(masks, i, oh, ow, Hm, Wm, f);
    float o[16];
    md_forward(W, eps1, eps2, f, o);
#pragma unroll
    for (int c = 0; c < 16; ++c) out[i * 16 + c] = o[c];
  }
}

// every thread back-propagates one output pixel; the 332 parameter gradients are summed over the warp with shuffles,
// over the CTA in shared memory and over the grid with one atomic per parameter per CTA.  All warps of a CTA run the
// same number of iterations (inactive lanes contribute zeros), so the shuffles are full-warp.
__global__ void __launch_bounds__(256) mask_downscale_bwd_kernel(const float* __restrict__ masks,
                                                                 const float* __restrict__ weights, float eps1, float eps2,
                                                                 const float* __restrict__ dout, float* __restrict__ dweights,
                                                                 long long n_seq, int Hm, int Wm) {
  __shared__ float W[MD_N];
  __shared__ float G[MD_N];
  for (int i = threadIdx.x; i < MD_N; i += blockDim.x) {
    W[i] = weights[i];
    G[i] = 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int oh = Hm >> 2, ow = Wm >> 2;
  const long long total = n_seq * oh * ow;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long iters = (total + stride - 1) / stride;
  for (long long it = 0; it < iters; ++it) {
    const long long i = it * stride + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const bool on = i < total;
    MdFwd f;
    float go[16];
    if (on) {
      md_load(masks, i, oh, ow, Hm, Wm, f);
#pragma unroll
      for (int c = 0; c < 16; ++c) go[c] = dout[i * 16 + c];
    } else {
#pragma unroll
      for (int y = 0; y < 4; ++y)
#pragma unroll
        for (int x = 0; x < 4; ++x) f.in[y][x] = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) go[c] = 0.f;
    }
    float o[16];
    md_forward(W, eps1, eps2, f, o);
    // ---- second LayerNorm + GELU
    float dxh[16], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float g = go[c] * gelu_grad(f.pre2[c]);
      float a = warp_sum(g * f.xh2[c]), b = warp_sum(g);
      if (lane == 0) {
        atomicAdd(G + MD_G2 + c, a);
        atomicAdd(G + MD_BE2 + c, b);
      }
      dxh[c] = g * W[MD_G2 + c];
      s1 += dxh[c];
      s2 += dxh[c] * f.xh2[c];
    }
    s1 *= (1.f / 16.f);
    s2 *= (1.f / 16.f);
    float dt2[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) dt2[c] = f.rstd2 * (dxh[c] - s1 - f.xh2[c] * s2);
    // ---- second convolution
    float da1[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int c1 = 0; c1 < 4; ++c1) da1[k][c1] = 0.f;
#pragma unroll
    for (int c2 = 0; c2 < 16; ++c2) {
      const float b = warp_sum(dt2[c2]);
      if (lane == 0) atomicAdd(G + MD_B3 + c2, b);
#pragma unroll
      for (int c1 = 0; c1 < 4; ++c1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float a = warp_sum(dt2[c2] * f.a1[k][c1]);
          if (lane == 0) atomicAdd(G + MD_W3 + c2 * 16 + c1 * 4 + k, a);
          da1[k][c1] = fmaf(W[MD_W3 + c2 * 16 + c1 * 4 + k], dt2[c2], da1[k][c1]);
        }
      }
    }
    // ---- first LayerNorm + GELU + convolution, per 2x2 sub-block
#pragma unroll
    for (int sb = 0; sb < 4; ++sb) {
      const int sy = sb >> 1, sx = sb & 1;
      float d1[4], t1 = 0.f, t2 = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float g = da1[sb][c] * gelu_grad(f.pre1[sb][c]);
        const float a = warp_sum(g * f.xh1[sb][c]), b = warp_sum(g);
        if (lane == 0) {
          atomicAdd(G + MD_G1 + c, a);
          atomicAdd(G + MD_BE1 + c, b);
        }
        d1[c] = g * W[MD_G1 + c];
        t1 += d1[c];
        t2 += d1[c] * f.xh1[sb][c];
      }
      t1 *= 0.25f;
      t2 *= 0.25f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float dt = f.rstd1[sb] * (d1[c] - t1 - f.xh1[sb][c] * t2);
        const float b = warp_sum(dt);
        if (lane == 0) atomicAdd(G + MD_B0 + c, b);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float a = warp_sum(dt * f.in[2 * sy + (k >> 1)][2 * sx + (k & 1)]);
          if (lane == 0) atomicAdd(G + MD_W0 + c * 4 + k, a);
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < MD_N; i += blockDim.x) atomicAdd(dweights + i, G[i]);
}

// ------------------------------------------------------------------------------------------------------------------
// adjoint of la_resize_bilinear (token-major fp32 [n, h, w, c])
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) resize_bilinear_bwd_kernel(const float* __restrict__ dout, float* __restrict__ din,
                                                                  long long n, int ih, int iw, int oh, int ow, int c) {
  const float sy = static_cast<float>(ih) / oh, sx = static_cast<float>(iw) / ow;
  const long long total = n * oh * ow * c;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % c);
    long long r = i / c;
    const int x = static_cast<int>(r % ow);
    r /= ow;
    const int y = static_cast<int>(r % oh);
    const long long s = r / oh;
    int y0, y1, x0, x1;
    float ly, lx;
    tap(y, sy, ih, y0, y1, ly);
    tap(x, sx, iw, x0, x1, lx);
    const float g = dout[i];
    float* base = din + s * ih * iw * c + ch;
    atomicAdd(base + (static_cast<long long>(y0) * iw + x0) * c, (1.f - ly) * (1.f - lx) * g);
    atomicAdd(base + (static_cast<long long>(y0) * iw + x1) * c, (1.f - ly) * lx * g);
    atomicAdd(base + (static_cast<long long>(y1) * iw + x0) * c, ly * (1.f - lx) * g);
    atomicAdd(base + (static_cast<long long>(y1) * iw + x1) * c, ly * lx * g);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// src = support features (one map per (episode, example), shared by its C class sequences) + dense mask embedding
// (or the not-a-mask / no-mask vector)       prompt_encoder.py:783-803
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
src_combine_kernel(const float* __restrict__ feat, const float* __restrict__ dense, const unsigned char* __restrict__ mflag,
                   const float* __restrict__ alt, float* __restrict__ out, long long n_seq, int T, int D, int C) {
  const long long total = n_seq * T * D;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(i % D);
    const long long r = i / D;
    const int t = static_cast<int>(r % T);
    const long long s = r / T;
    const bool use_dense = dense != nullptr && (mflag == nullptr || mflag[s] != 0);
    out[i] = feat[((s / C) * T + t) * D + d] + (use_dense ? dense[i] : alt[d]);
  }
}

// dfeat[img, t, :] = sum_c dsrc[(img, c), t, :];  ddense = dsrc where the sequence has a mask, else 0;
// dalt_rows[img, t, :] = sum over the sequences WITHOUT a mask (column-summed by la_bcast_reduce_f32 afterwards)
__global__ void __launch_bounds__(256)
src_combine_bwd_kernel(const float* __restrict__ dsrc, const unsigned char* __restrict__ mflag, int has_dense,
                       float* __restrict__ dfeat, float* __restrict__ ddense, float* __restrict__ dalt_rows,
                       long long n_img, int T, int D, int C) {
  const long long total = n_img * T * D;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(i % D);
    const long long r = i / D;
    const int t = static_cast<int>(r % T);
    const long long img = r / T;
    float sum = 0.f, alt = 0.f;
    for (int c = 0; c < C; ++c) {
      const long long s = img * C + c;
      const long long j = (s * T + t) * D + d;
      const float g = dsrc[j];
      sum += g;
      const bool use_dense = has_dense && (mflag == nullptr || mflag[s] != 0);
      if (ddense) ddense[j] = use_dense ? g : 0.f;
      if (!use_dense) alt += g;
    }
    dfeat[i] = sum;
    if (dalt_rows) dalt_rows[i] = alt;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// sparse prompt tokens: gradients of the four point / corner embeddings and of not_a_point_embed (la_embed_sparse)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embed_sparse_bwd_kernel(const float* __restrict__ plabels, const float* __restrict__ bflags, const float* __restrict__ dout,
                        float* __restrict__ dtab, float* __restrict__ dnap, long long n_seq, int P, int Bx, int pad, int n,
                        int D, int has_points) {
  const long long total = n_seq * n * D;
  const int n_pts = has_points ? P + pad : 0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(i % D);
    const long long r = i / D;
    const int tok = static_cast<int>(r % n);
    const long long s = r / n;
    int add = -1;
    bool null_tok = false;
    if (tok < n_pts) {
      const float label = tok < P ? plabels[s * P + tok] : -1.f;
      null_tok = label == 0.f;
      if (label == -1.f) add = 0;
      if (label == 1.f) add = 1;
    } else {
      const int cj = tok - n_pts;
      add = 2 + (cj & 1);
      null_tok = bflags[s * Bx + (cj % Bx)] == 0.f;
    }
    const float g = dout[i];
    if (null_tok)
      atomicAdd(dnap + j, g);
    else if (add >= 0)
      atomicAdd(dtab + add * D + j, g);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// segment mean (spatial mean of the fused image tokens, prompt_encoder.py:733-735) and masked mean backward
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) segment_mean_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                           long long n_seg, int seg_rows, int d) {
  const long long total = n_seg * d;
  const float inv = 1.f / seg_rows;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % d);
    const long long s = i / d;
    const float* p = x + s * seg_rows * d + c;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int t = 0;
    for (; t + 3 < seg_rows; t += 4) {
      a0 += p[static_cast<long long>(t) * d];
      a1 += p[static_cast<long long>(t + 1) * d];
      a2 += p[static_cast<long long>(t + 2) * d];
      a3 += p[static_cast<long long>(t + 3) * d];
    }
    for (; t < seg_rows; ++t) a0 += p[static_cast<long long>(t) * d];
    out[i] = ((a0 + a1) + (a2 + a3)) * inv;
  }
}

__global__ void __launch_bounds__(256) segment_mean_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dx,
                                                               long long n_seg, int seg_rows, int d) {
  const long long total = n_seg * seg_rows * d;
  const float inv = 1.f / seg_rows;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % d);
    const long long s = i / (static_cast<long long>(seg_rows) * d);
    dx[i] = dout[s * d + c] * inv;
  }
}

__global__ void __launch_bounds__(256)
masked_mean_bwd_kernel(const float* __restrict__ dout, const unsigned char* __restrict__ flag, float* __restrict__ demb,
                       int B, int M, int C, int D) {
  const long long total = static_cast<long long>(B) * M * C * D;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(i % D);
    long long r = i / D;
    const int c = static_cast<int>(r % C);
    r /= C;
    const int m = static_cast<int>(r % M);
    const long long b = r / M;
    float cnt = 0.f;
    for (int mm = 0; mm < M; ++mm) cnt += flag[(b * M + mm) * C + c] ? 1.f : 0.f;
    const float f = flag[(b * M + m) * C + c] ? 1.f : 0.f;
    demb[i] = f * dout[(b * C + c) * D + d] / (cnt == 0.f ? 1.f : cnt);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// classify backward: logits[b, c, p] = sum_k cls[b, c, k] x[b, p, k]
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
classify_bwd_x_kernel(const float* __restrict__ dl, const float* __restrict__ cls, float* __restrict__ dx, int B,
                      long long P, int C, int dk) {
  const long long total = static_cast<long long>(B) * P * dk;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % dk);
    const long long r = i / dk;
    const long long pix = r % P;
    const long long b = r / P;
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(dl[(b * C + c) * P + pix], cls[(b * C + c) * dk + k], acc);
    dx[i] = acc;
  }
}

constexpr int CLS_PIX = 512;   // pixels per CTA of the dcls reduction
__global__ void __launch_bounds__(256)
classify_bwd_cls_kernel(const float* __restrict__ dl, const __nv_bfloat16* __restrict__ x, float* __restrict__ dcls,
                        long long P, int C, int dk) {
  const long long b = blockIdx.y;
  const long long p0 = static_cast<long long>(blockIdx.x) * CLS_PIX;
  const long long p1 = p0 + CLS_PIX < P ? p0 + CLS_PIX : P;
  for (int idx = threadIdx.x; idx < C * dk; idx += blockDim.x) {
    const int k = idx % dk, c = idx / dk;
    const float* g = dl + (b * C + c) * P;
    const __nv_bfloat16* xr = x + b * P * dk + k;
    float acc = 0.f;
    for (long long pix = p0; pix < p1; ++pix) acc = fmaf(g[pix], __bfloat162float(xr[pix * dk]), acc);
    atomicAdd(dcls + (b * C + c) * dk + k, acc);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// adjoint of la_postprocess_masks (two bilinear resizes + crop; padded / flag_gts-masked outputs carry no gradient)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void scatter_stage1(float* __restrict__ plane, int lh, int lw, int S, int y, int x, float g) {
  int y0, y1, x0, x1;
  float ly, lx;
  tap(y, static_cast<float>(lh) / S, lh, y0, y1, ly);
  tap(x, static_cast<float>(lw) / S, lw, x0, x1, lx);
  atomicAdd(plane + y0 * lw + x0, (1.f - ly) * (1.f - lx) * g);
  atomicAdd(plane + y0 * lw + x1, (1.f - ly) * lx * g);
  atomicAdd(plane + y1 * lw + x0, ly * (1.f - lx) * g);
  atomicAdd(plane + y1 * lw + x1, ly * lx * g);
}

__global__ void __launch_bounds__(256)
postprocess_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ sizes,
                       const unsigned char* __restrict__ flag_gts, float* __restrict__ din, int B, int C, int lh, int lw,
                       int S, int Hmax, int Wmax) {
  const long long total = static_cast<long long>(B) * C * Hmax * Wmax;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % Wmax);
    long long r = i / Wmax;
    const int y = static_cast<int>(r % Hmax);
    r /= Hmax;
    const int c = static_cast<int>(r % C);
    const int b = static_cast<int>(r / C);
    const int oh = sizes[b * 4], ow = sizes[b * 4 + 1], ih = sizes[b * 4 + 2], iw = sizes[b * 4 + 3];
    if ((flag_gts && flag_gts[b * C + c] == 0) || y >= oh || x >= ow) continue;
    const float g = dout[i];
    if (g == 0.f) continue;
    float* plane = din + (static_cast<long long>(b) * C + c) * lh * lw;
    int y0, y1, x0, x1;
    float ly, lx;
    tap(y, static_cast<float>(ih) / oh, ih, y0, y1, ly);
    tap(x, static_cast<float>(iw) / ow, iw, x0, x1, lx);
    scatter_stage1(plane, lh, lw, S, y0, x0, (1.f - ly) * (1.f - lx) * g);
    scatter_stage1(plane, lh, lw, S, y0, x1, (1.f - ly) * lx * g);
    scatter_stage1(plane, lh, lw, S, y1, x0, ly * (1.f - lx) * g);
    scatter_stage1(plane, lh, lw, S, y1, x1, ly * lx * g);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// AdamW (torch.optim.AdamW, decoupled weight decay) over one flat fp32 parameter bucket
// ------------------------------------------------------------------------------------------------------------------
// bias corrections and learning rate either by value or (bc_dev != nullptr: {1 - beta1^t, sqrt(1 - beta2^t), lr}) from
// device memory, so that a captured CUDA graph of the step can be replayed with a growing step count and a scheduled
// learning rate
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
             float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt, float grad_scale,
             const float* __restrict__ bc_dev) {
  if (bc_dev) {
    bc1 = bc_dev[0];
    bc2_sqrt = bc_dev[1];
    lr = bc_dev[2];
  }
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float gi = g[i] * grad_scale;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

}  // namespace
}  // namespace la

extern "C" {

using namespace la;

#define ST(s) static_cast<cudaStream_t>(s)

int la_cast_bf16(void* stream, const float* in, void* out, long long n) {
  LA_CHECK_ARG(in && out && n > 0, "la_cast_bf16: bad arguments");
  cast_bf16_kernel<<<train_grid(n), 256, 0, ST(stream)>>>(in, static_cast<__nv_bfloat16*>(out), n);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_split_bf16(void* stream, const float* in, void* hi, void* lo, long long n) {
  LA_CHECK_ARG(in && hi && lo && n > 0, "la_split_bf16: bad arguments");
  split_bf16_kernel<<<train_grid(n), 256, 0, ST(stream)>>>(in, static_cast<__nv_bfloat16*>(hi),
                                                            static_cast<__nv_bfloat16*>(lo), n);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_split3_bf16(void* stream, const float* in, void* hi, void* mid, void* lo, long long n) {
  LA_CHECK_ARG(in && hi && mid && lo && n > 0, "la_split3_bf16: bad arguments");
  split3_bf16_kernel<<<train_grid(n), 256, 0, ST(stream)>>>(in, static_cast<__nv_bfloat16*>(hi),
                                                             static_cast<__nv_bfloat16*>(mid),
                                                             static_cast<__nv_bfloat16*>(lo), n);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_add_f32(void* stream, const float* a, const float* b, float* out, long long n) {
  LA_CHECK_ARG(a && b && out && n > 0, "la_add_f32: bad arguments");
  add_kernel<<<train_grid(n), 256, 0, ST(stream)>>>(a, b, out, n);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_relu_bwd_f32(void* stream, const float* dy, const float* y, float* dx, long long n) {
  LA_CHECK_ARG(dy && y && dx && n > 0, "la_relu_bwd_f32: bad arguments");
  relu_bwd_kernel<<<train_grid(n), 256, 0, ST(stream)>>>(dy, y, dx, n);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_gelu_f32(void* stream, const float* x, float* y, long long n) {
  LA_CHECK_ARG(x && y && n > 0, "la_gelu_f32: bad arguments");
  gelu_kernel<<<train_grid(n), 256, 0, ST(stream)>>>(x, y, n);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_gelu_bwd_f32(void* stream, const float* dy, const float* x, float* dx, long long n) {
  LA_CHECK_ARG(dy && x && dx && n > 0, "la_gelu_bwd_f32: bad arguments");
  gelu_bwd_kernel<<<train_grid(n), 256, 0, ST(stream)>>>(dy, x, dx, n);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_cast_transpose_bf16(void* stream, const void* in, int in_dtype, long long ld_in, void* out, long long ld_out,
                           long long rows, int cols) {
  LA_CHECK_ARG(in && out && rows > 0 && cols > 0 && ld_in >= cols && ld_out >= rows, "la_cast_transpose_bf16: bad arguments");
  LA_CHECK_ARG(in_dtype == LA_DTYPE_F32 || in_dtype == LA_DTYPE_BF16, "la_cast_transpose_bf16: fp32 or bf16 input");
  const long long gx = (ld_out + 31) / 32;
  LA_CHECK_ARG(gx < (1ll << 31) && (cols + 31) / 32 <= 65535, "la_cast_transpose_bf16: matrix too large");
  dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>((cols + 31) / 32));
  if (in_dtype == LA_DTYPE_F32)
    cast_transpose_kernel<float><<<grid, 256, 0, ST(stream)>>>(static_cast<const float*>(in), ld_in,
                                                               static_cast<__nv_bfloat16*>(out), ld_out, rows, cols);
  else
    cast_transpose_kernel<__nv_bfloat16><<<grid, 256, 0, ST(stream)>>>(static_cast<const __nv_bfloat16*>(in), ld_in,
                                                                       static_cast<__nv_bfloat16*>(out), ld_out, rows, cols);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_grad_prep_bf16(void* stream, const float* dy, long long ld, void* out_bf16, void* out_t_bf16, long long ld_t,
                      float* colsum, long long rows, int cols) {
  LA_CHECK_ARG(dy && rows > 0 && cols > 0 && ld >= cols && (out_bf16 || out_t_bf16 || colsum), "la_grad_prep_bf16: bad arguments");
  LA_CHECK_ARG(!out_t_bf16 || ld_t >= rows, "la_grad_prep_bf16: ld_t must cover the rows");
  const long long span = out_t_bf16 ? ld_t : rows;
  const long long gx = (span + 32 * PREP_TILES - 1) / (32 * PREP_TILES);
  LA_CHECK_ARG(gx < (1ll << 31) && (cols + 31) / 32 <= 65535, "la_grad_prep_bf16: matrix too large");
  if (colsum) LA_CHECK_CUDA(cudaMemsetAsync(colsum, 0, sizeof(float) * cols, ST(stream)));
  dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>((cols + 31) / 32));
  grad_prep_kernel<<<grid, 256, 0, ST(stream)>>>(dy, ld, static_cast<__nv_bfloat16*>(out_bf16),
                                                 static_cast<__nv_bfloat16*>(out_t_bf16), out_t_bf16 ? ld_t : rows, colsum,
                                                 rows, cols);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_bcast_reduce_f32(void* stream, const float* dy, float* out, long long rows, int d, long long row_div,
                        long long b_mod, int accumulate) {
  LA_CHECK_ARG(dy && out && rows > 0 && d > 0 && row_div > 0 && b_mod > 0, "la_bcast_reduce_f32: bad arguments");
  if (!accumulate) LA_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * b_mod * d, ST(stream)));
  if (b_mod == 1 && (rows + COLSUM_ROWS - 1) / COLSUM_ROWS <= 65535) {
    dim3 grid(static_cast<unsigned>((d + 31) / 32), static_cast<unsigned>((rows + COLSUM_ROWS - 1) / COLSUM_ROWS));
    colsum_kernel<<<grid, 256, 0, ST(stream)>>>(dy, out, rows, d);
  } else {
    const long long items = (rows + RED_ROWS - 1) / RED_ROWS * d;
    bcast_reduce_kernel<<<train_grid(items), 256, 0, ST(stream)>>>(dy, out, rows, d, row_div, b_mod);
  }
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

#define LA_NV_DISPATCH(d, CALL)       \
  do {                                \
    if ((d) <= 32) { CALL(1); }       \
    else if ((d) <= 64) { CALL(2); }  \
    else if ((d) <= 128) { CALL(4); } \
    else if ((d) <= 256) { CALL(8); } \
    else if ((d) <= 512) { CALL(16); } \
    else { CALL(32); }                \
  } while (0)

int la_layernorm_f32(void* stream, const float* x, const float* gamma, const float* beta, float eps, int act, float* y,
                     long long rows, int d) {
  LA_CHECK_ARG(x && y && rows > 0 && d > 0 && d <= 1024 && (gamma == nullptr) == (beta == nullptr),
               "la_layernorm_f32: bad arguments (d <= 1024)");
  LA_CHECK_ARG(act == LA_ACT_NONE || act == LA_ACT_GELU, "la_layernorm_f32: act must be none or GELU");
  const unsigned grid = train_grid(rows * 32);
#define CALL(NV) layernorm_f32_kernel<NV><<<grid, 256, 0, ST(stream)>>>(x, gamma, beta, eps, act, y, rows, d)
  LA_NV_DISPATCH(d, CALL);
#undef CALL
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_layernorm_f32_bwd(void* stream, const float* x, const float* gamma, const float* beta, float eps, int act,
                         const float* dy, float* dx, float* dgamma, float* dbeta, long long rows, int d) {
  LA_CHECK_ARG(x && dy && dx && rows > 0 && d > 0 && d <= 1024 && (gamma == nullptr) == (beta == nullptr) &&
                   (dgamma == nullptr) == (dbeta == nullptr),
               "la_layernorm_f32_bwd: bad arguments (d <= 1024)");
  LA_CHECK_ARG(act == LA_ACT_NONE || act == LA_ACT_GELU, "la_layernorm_f32_bwd: act must be none or GELU");
  long long blocks = (rows + 7) / 8;
  const long long cap = 4ll * sm_count();
  if (blocks > cap) blocks = cap;
  const unsigned grid = static_cast<unsigned>(blocks);
  const size_t smem = sizeof(float) * 2 * d;
#define CALL(NV) \
  layernorm_f32_bwd_kernel<NV><<<grid, 256, smem, ST(stream)>>>(x, gamma, beta, eps, act, dy, dx, dgamma, dbeta, rows, d)
  LA_NV_DISPATCH(d, CALL);
#undef CALL
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

#define LA_DH_DISPATCH(dh, CALL)     \
  do {                               \
    if ((dh) <= 8) { CALL(8); }      \
    else if ((dh) <= 16) { CALL(16); } \
    else if ((dh) <= 32) { CALL(32); } \
    else { CALL(64); }               \
  } while (0)

int la_attention_f32(void* stream, const float* q, const float* k, const float* v, float* out, float* lse,
                     long long n_seq, int nq, int nk, int heads, int head_dim, float scale) {
  LA_CHECK_ARG(q && k && v && out && lse && n_seq > 0 && nq > 0 && nk > 0 && heads > 0 && head_dim > 0 && head_dim <= 64 &&
                   head_dim % 4 == 0,
               "la_attention_f32: bad arguments (head_dim <= 64, a multiple of 4)");
  AttnParams p{};
  p.q = q; p.k = k; p.v = v; p.out = out; p.lse = lse;
  p.n_seq = n_seq; p.nq = nq; p.nk = nk; p.heads = heads; p.dh = head_dim; p.scale = scale;
  if (nk <= 32) {
    const unsigned grid = train_grid(n_seq * heads * nq, 128);
#define CALL(DH) attn_f32_fwd_kernel<DH, true><<<grid, 128, 0, ST(stream)>>>(p)
    LA_DH_DISPATCH(head_dim, CALL);
#undef CALL
  } else {
    const unsigned grid = train_grid(n_seq * heads * nq * 32, 128);
#define CALL(DH) attn_f32_fwd_kernel<DH, false><<<grid, 128, 0, ST(stream)>>>(p)
    LA_DH_DISPATCH(head_dim, CALL);
#undef CALL
  }
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_attention_f32_bwd(void* stream, const float* q, const float* k, const float* v, const float* out, const float* lse,
                         const float* dout, float* delta, float* dq, float* dk, float* dv, long long n_seq, int nq, int nk,
                         int heads, int head_dim, float scale) {
  LA_CHECK_ARG(q && k && v && out && lse && dout && delta && dq && dk && dv && n_seq > 0 && nq > 0 && nk > 0 &&
                   heads > 0 && head_dim > 0 && head_dim <= 64 && head_dim % 4 == 0,
               "la_attention_f32_bwd: bad arguments (head_dim <= 64, a multiple of 4)");
  AttnParams p{};
  p.q = q; p.k = k; p.v = v; p.out = const_cast<float*>(out); p.lse = const_cast<float*>(lse);
  p.dout = dout; p.delta = delta; p.dq = dq; p.dk = dk; p.dv = dv;
  p.n_seq = n_seq; p.nq = nq; p.nk = nk; p.heads = heads; p.dh = head_dim; p.scale = scale;
  // dq (+ delta) first: the dk / dv kernels read delta
  if (nk <= 32) {
    const unsigned gq = train_grid(n_seq * heads * nq, 128);
#define CALL(DH) attn_f32_bwd_q_kernel<DH, true><<<gq, 128, 0, ST(stream)>>>(p)
    LA_DH_DISPATCH(head_dim, CALL);
#undef CALL
  } else {
    const unsigned gq = train_grid(n_seq * heads * nq * 32, 128);
#define CALL(DH) attn_f32_bwd_q_kernel<DH, false><<<gq, 128, 0, ST(stream)>>>(p)
    LA_DH_DISPATCH(head_dim, CALL);
#undef CALL
  }
  if (nq <= 32) {
    const unsigned gk = train_grid(n_seq * heads * nk, 128);
#define CALL(DH) attn_f32_bwd_kv_kernel<DH, true><<<gk, 128, 0, ST(stream)>>>(p)
    LA_DH_DISPATCH(head_dim, CALL);
#undef CALL
  } else {
    const unsigned gk = train_grid(n_seq * heads * nk * 32, 128);
#define CALL(DH) attn_f32_bwd_kv_kernel<DH, false><<<gk, 128, 0, ST(stream)>>>(p)
    LA_DH_DISPATCH(head_dim, CALL);
#undef CALL
  }
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_col2im_3x3_f32(void* stream, const float* dcol, float* dx, long long n_img, int height, int width, int channels) {
  LA_CHECK_ARG(dcol && dx && n_img > 0 && height > 0 && width > 0 && channels > 0, "la_col2im_3x3_f32: bad arguments");
  col2im_3x3_kernel<<<train_grid(n_img * height * width * channels), 256, 0, ST(stream)>>>(dcol, dx, n_img, height, width,
                                                                                         channels);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_mask_downscale_dev(void* stream, const float* masks, const float* weights, float eps1, float eps2, float* out,
                          long long n_seq, int height, int width) {
  LA_CHECK_ARG(masks && weights && out && n_seq > 0 && height > 0 && width > 0 && height % 4 == 0 && width % 4 == 0,
               "la_mask_downscale_dev: bad arguments");
  mask_downscale_dev_kernel<<<train_grid(n_seq * (height / 4) * (width / 4)), 256, 0, ST(stream)>>>(
      masks, weights, eps1, eps2, out, n_seq, height, width);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_mask_downscale_bwd(void* stream, const float* masks, const float* weights, float eps1, float eps2,
                          const float* dout, float* dweights, long long n_seq, int height, int width) {
  LA_CHECK_ARG(masks && weights && dout && dweights && n_seq > 0 && height > 0 && width > 0 && height % 4 == 0 &&
                   width % 4 == 0,
               "la_mask_downscale_bwd: bad arguments");
  LA_CHECK_CUDA(cudaMemsetAsync(dweights, 0, sizeof(float) * MD_N, ST(stream)));
  long long blocks = (n_seq * (height / 4) * (width / 4) + 255) / 256;
  const long long cap = 4ll * sm_count();
  if (blocks > cap) blocks = cap;
  mask_downscale_bwd_kernel<<<static_cast<unsigned>(blocks), 256, 0, ST(stream)>>>(masks, weights, eps1, eps2, dout,
                                                                                   dweights, n_seq, height, width);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_resize_bilinear_bwd(void* stream, const float* dout, float* din, long long n, int in_h, int in_w, int out_h,
                           int out_w, int channels) {
  LA_CHECK_ARG(dout && din && n > 0 && in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0 && channels > 0,
               "la_resize_bilinear_bwd: bad arguments");
  LA_CHECK_CUDA(cudaMemsetAsync(din, 0, sizeof(float) * n * in_h * in_w * channels, ST(stream)));
  resize_bilinear_bwd_kernel<<<train_grid(n * out_h * out_w * channels), 256, 0, ST(stream)>>>(dout, din, n, in_h, in_w,
                                                                                             out_h, out_w, channels);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_src_combine_f32(void* stream, const float* feat, const float* dense, const unsigned char* mask_flags,
                       const float* alt, float* out, long long n_seq, int tokens, int d, int n_classes) {
  LA_CHECK_ARG(feat && alt && out && n_seq > 0 && tokens > 0 && d > 0 && n_classes > 0 && n_seq % n_classes == 0,
               "la_src_combine_f32: bad arguments");
  src_combine_kernel<<<train_grid(n_seq * tokens * d), 256, 0, ST(stream)>>>(feat, dense, mask_flags, alt, out, n_seq,
                                                                            tokens, d, n_classes);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_src_combine_bwd(void* stream, const float* dsrc, const unsigned char* mask_flags, int has_dense, float* dfeat,
                       float* ddense, float* dalt_rows, long long n_seq, int tokens, int d, int n_classes) {
  LA_CHECK_ARG(dsrc && dfeat && n_seq > 0 && tokens > 0 && d > 0 && n_classes > 0 && n_seq % n_classes == 0,
               "la_src_combine_bwd: bad arguments");
  const long long n_img = n_seq / n_classes;
  src_combine_bwd_kernel<<<train_grid(n_img * tokens * d), 256, 0, ST(stream)>>>(dsrc, mask_flags, has_dense, dfeat,
                                                                                ddense, dalt_rows, n_img, tokens, d,
                                                                                n_classes);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_embed_sparse_bwd(void* stream, const float* point_labels, int n_points, const float* box_flags, int n_boxes,
                        int has_points, const float* dout, float* dtable, float* dnot_a_point, long long n_seq, int d) {
  LA_CHECK_ARG(dout && dtable && dnot_a_point && n_seq > 0 && d > 0 && (has_points || n_boxes > 0),
               "la_embed_sparse_bwd: bad arguments");
  LA_CHECK_ARG((!has_points || point_labels) && (n_boxes == 0 || box_flags), "la_embed_sparse_bwd: missing labels / flags");
  const int pad = (has_points && n_boxes == 0) ? 1 : 0;
  const int n = (has_points ? n_points + pad : 0) + 2 * n_boxes;
  LA_CHECK_CUDA(cudaMemsetAsync(dtable, 0, sizeof(float) * 4 * d, ST(stream)));
  LA_CHECK_CUDA(cudaMemsetAsync(dnot_a_point, 0, sizeof(float) * d, ST(stream)));
  embed_sparse_bwd_kernel<<<train_grid(n_seq * n * d), 256, 0, ST(stream)>>>(point_labels, box_flags, dout, dtable,
                                                                            dnot_a_point, n_seq, n_points, n_boxes, pad, n,
                                                                            d, has_points);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_segment_mean_f32(void* stream, const float* x, float* out, long long n_seg, int seg_rows, int d) {
  LA_CHECK_ARG(x && out && n_seg > 0 && seg_rows > 0 && d > 0, "la_segment_mean_f32: bad arguments");
  segment_mean_kernel<<<train_grid(n_seg * d), 256, 0, ST(stream)>>>(x, out, n_seg, seg_rows, d);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_segment_mean_bwd(void* stream, const float* dout, float* dx, long long n_seg, int seg_rows, int d) {
  LA_CHECK_ARG(dout && dx && n_seg > 0 && seg_rows > 0 && d > 0, "la_segment_mean_bwd: bad arguments");
  segment_mean_bwd_kernel<<<train_grid(n_seg * seg_rows * d), 256, 0, ST(stream)>>>(dout, dx, n_seg, seg_rows, d);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_masked_mean_bwd(void* stream, const float* dout, const unsigned char* flags, float* demb, int batch, int examples,
                       int classes, int d) {
  LA_CHECK_ARG(dout && flags && demb && batch > 0 && examples > 0 && classes > 0 && d > 0, "la_masked_mean_bwd: bad arguments");
  masked_mean_bwd_kernel<<<train_grid(static_cast<long long>(batch) * examples * classes * d), 256, 0, ST(stream)>>>(
      dout, flags, demb, batch, examples, classes, d);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_classify_bwd(void* stream, const float* dlogits, const void* x, const float* cls, float* dx, float* dcls, int batch,
                    long long pixels, int classes, int dk) {
  LA_CHECK_ARG(dlogits && x && cls && dx && dcls && batch > 0 && batch <= 65535 && pixels > 0 && classes > 0 && dk > 0,
               "la_classify_bwd: bad arguments");
  classify_bwd_x_kernel<<<train_grid(batch * pixels * dk), 256, 0, ST(stream)>>>(dlogits, cls, dx, batch, pixels, classes,
                                                                                dk);
  LA_CHECK_CUDA(cudaGetLastError());
  LA_CHECK_CUDA(cudaMemsetAsync(dcls, 0, sizeof(float) * batch * classes * dk, ST(stream)));
  dim3 grid(static_cast<unsigned>((pixels + CLS_PIX - 1) / CLS_PIX), static_cast<unsigned>(batch));
  classify_bwd_cls_kernel<<<grid, 256, 0, ST(stream)>>>(dlogits, static_cast<const __nv_bfloat16*>(x), dcls, pixels,
                                                        classes, dk);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_postprocess_masks_bwd(void* stream, const float* dout, const int* sizes, const unsigned char* flag_gts, float* din,
                             int batch, int classes, int low_h, int low_w, int image_size, int out_h, int out_w) {
  LA_CHECK_ARG(dout && sizes && din && batch > 0 && classes > 0 && low_h > 0 && low_w > 0 && image_size > 0 && out_h > 0 &&
                   out_w > 0,
               "la_postprocess_masks_bwd: bad arguments");
  LA_CHECK_CUDA(cudaMemsetAsync(din, 0, sizeof(float) * batch * classes * low_h * low_w, ST(stream)));
  postprocess_bwd_kernel<<<train_grid(static_cast<long long>(batch) * classes * out_h * out_w), 256, 0, ST(stream)>>>(
      dout, sizes, flag_gts, din, batch, classes, low_h, low_w, image_size, out_h, out_w);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_adamw_f32(void* stream, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                 float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale) {
  LA_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && n > 0 && step >= 1, "la_adamw_f32: bad arguments");
  const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.f - powf(beta2, static_cast<float>(step));
  adamw_kernel<<<train_grid(n), 256, 0, ST(stream)>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                      weight_decay, bc1, sqrtf(bc2), grad_scale, nullptr);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_adamw_f32_dev(void* stream, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                     float beta1, float beta2, float eps, float weight_decay, const float* step_scalars,
                     float grad_scale) {
  LA_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && step_scalars && n > 0, "la_adamw_f32_dev: bad arguments");
  adamw_kernel<<<train_grid(n), 256, 0, ST(stream)>>>(params, grads, exp_avg, exp_avg_sq, n, 0.f, beta1, beta2, eps,
                                                      weight_decay, 1.f, 1.f, grad_scale, step_scalars);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

}  // extern "C"
