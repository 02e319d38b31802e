// labelanything_b200 — layout conversion kernels at the module boundary (sm_100a, HBM-bound).
//
// The reference hands feature maps around as NCHW fp32 (image_encoder.py:119-131 `x.permute(0, 3, 1, 2)`,
// lam.py:139-170); the native kernels work token-major ([rows, channels], channels contiguous).  These kernels
// are the only places where a layout changes:
//   la_nchw_to_tokens : [n, C, P] fp32 -> [n, P, C] fp32 and/or bf16 (precomputed `embeddings` input, lam.py:139-146)
//   la_tokens_to_nchw : [n, P, C] fp32 -> [n, C, P] fp32 (encoder output handed back to reference-style callers,
//                       preprocess.py:65-73)
//   la_copy_slabs     : gather equally spaced row slabs (e.g. the query image of every episode, lam.py:167-168
//                       `embeddings[:, 0]`) into one contiguous fp32 and/or bf16 buffer.
// Both transposes go through a 32x33 shared-memory tile so that reads and writes are coalesced.
#include "la_common.cuh"

namespace la {

template <bool kToTokens>
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ in, float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16,
                 int C, int P) {
  // kToTokens: in [n, C, P] -> out [n, P, C];  else in [n, P, C] -> out [n, C, P]
  __shared__ float tile[32][33];
  const int R = kToTokens ? C : P;      // rows of the input matrix
  const int Q = kToTokens ? P : C;      // columns of the input matrix (contiguous)
  const long long img = blockIdx.z;
  const float* src = in + img * static_cast<long long>(R) * Q;
  const int q0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int r = r0 + ty + j, q = q0 + tx;
    tile[ty + j][tx] = (r < R && q < Q) ? __ldg(src + static_cast<long long>(r) * Q + q) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int q = q0 + ty + j, r = r0 + tx;  // output row q, column r
    if (q < Q && r < R) {
      const float v = tile[tx][ty + j];
      const long long o = img * static_cast<long long>(R) * Q + static_cast<long long>(q) * R + r;
      if (out_f32) out_f32[o] = v;
      if (out_bf16) out_bf16[o] = __float2bfloat16_rn(v);
    }
  }
}

__global__ void __launch_bounds__(256)
copy_slabs_kernel(const float* __restrict__ in, long long in_stride_rows, long long in_offset_rows,
                  float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, long long n_slabs,
                  long long slab_rows, int d4) {
  const long long per_slab = slab_rows * d4;
  const long long total = n_slabs * per_slab;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long s = i / per_slab, r = i % per_slab;
    const float4 v = __ldg(reinterpret_cast<const float4*>(in) + (s * in_stride_rows + in_offset_rows) * d4 + r);
    if (out_f32) reinterpret_cast<float4*>(out_f32)[i] = v;
    if (out_bf16) {
      uint2 pk;
      pk.x = pack_bf16(v.x, v.y);
      pk.y = pack_bf16(v.z, v.w);
      reinterpret_cast<uint2*>(out_bf16)[i] = pk;
    }
  }
}

// out[r, :] = a[r, :] + b[(r / row_div) % b_mod, :]   (fp32; the per-class code added to every sparse token)
__global__ void __launch_bounds__(256)
add_bcast_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long rows,
                 int d4, long long row_div, long long b_mod) {
  const long long total = rows * d4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / d4;
    const int c = static_cast<int>(i % d4);
    float4 v = __ldg(reinterpret_cast<const float4*>(a) + i);
    const float4 w = __ldg(reinterpret_cast<const float4*>(b) + ((r / row_div) % b_mod) * d4 + c);
    v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    reinterpret_cast<float4*>(out)[i] = v;
  }
}

// out[o, j, i, :] = in[o, i, j, :]   (rows of d fp32 channels; "b m c d -> b c m d")
__global__ void __launch_bounds__(256)
permute_rows_kernel(const float* __restrict__ in, float* __restrict__ out, long long outer, int na, int nb, int d4) {
  const long long total = outer * na * nb * d4;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % d4);
    long long r = idx / d4;
    const int i = static_cast<int>(r % na);
    r /= na;
    const int j = static_cast<int>(r % nb);
    const long long o = r / nb;
    reinterpret_cast<float4*>(out)[idx] =
        __ldg(reinterpret_cast<const float4*>(in) + ((o * na + i) * nb + j) * d4 + c);
  }
}

static unsigned stream_grid(long long items) {
  long long blocks = (items + 255) / 256;
  const long long cap = 8ll * sm_count();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<unsigned>(blocks);
}

}  // namespace la

extern "C" {

int la_add_bcast(void* stream, const float* a, const float* b, float* out, long long rows, int d, long long row_div,
                 long long b_mod) {
  using namespace la;
  LA_CHECK_ARG(a && b && out && rows > 0 && d > 0 && d % 4 == 0 && row_div > 0 && b_mod > 0, "la_add_bcast: bad arguments");
  add_bcast_kernel<<<stream_grid(rows * (d / 4)), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, out, rows, d / 4,
                                                                                             row_div, b_mod);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_permute_rows(void* stream, const float* in, float* out, long long outer, int na, int nb, int d) {
  using namespace la;
  LA_CHECK_ARG(in && out && outer > 0 && na > 0 && nb > 0 && d > 0 && d % 4 == 0, "la_permute_rows: bad arguments");
  permute_rows_kernel<<<stream_grid(outer * na * nb * (d / 4)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, out, outer, na, nb, d / 4);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_nchw_to_tokens(void* stream, const float* in, float* out_f32, void* out_bf16, long long n, int channels,
                      int pixels) {
  using namespace la;
  LA_CHECK_ARG(in && (out_f32 || out_bf16) && n > 0 && channels > 0 && pixels > 0, "la_nchw_to_tokens: bad arguments");
  LA_CHECK_ARG(n <= 65535, "la_nchw_to_tokens: more than 65535 images per call");
  dim3 grid((pixels + 31) / 32, (channels + 31) / 32, static_cast<unsigned>(n));
  transpose_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, out_f32, static_cast<__nv_bfloat16*>(out_bf16), channels, pixels);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_tokens_to_nchw(void* stream, const float* in, float* out, long long n, int channels, int pixels) {
  using namespace la;
  LA_CHECK_ARG(in && out && n > 0 && channels > 0 && pixels > 0, "la_tokens_to_nchw: bad arguments");
  LA_CHECK_ARG(n <= 65535, "la_tokens_to_nchw: more than 65535 images per call");
  dim3 grid((channels + 31) / 32, (pixels + 31) / 32, static_cast<unsigned>(n));
  transpose_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, out, nullptr, channels, pixels);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_copy_slabs(void* stream, const float* in, long long in_stride_rows, long long in_offset_rows, float* out_f32,
                  void* out_bf16, long long n_slabs, long long slab_rows, int d) {
  using namespace la;
  LA_CHECK_ARG(in && (out_f32 || out_bf16) && n_slabs > 0 && slab_rows > 0 && d > 0 && d % 4 == 0,
               "la_copy_slabs: bad arguments");
  const long long total = n_slabs * slab_rows * (d / 4);
  long long blocks = (total + 255) / 256;
  const long long cap = 8ll * sm_count();
  if (blocks > cap) blocks = cap;
  copy_slabs_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, in_stride_rows, in_offset_rows, out_f32, static_cast<__nv_bfloat16*>(out_bf16), n_slabs, slab_rows, d / 4);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

}  // extern "C"
