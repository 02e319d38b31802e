// labelanything_b200 — prompt-side and output-side streaming kernels (sm_100a, CUDA cores, HBM-bound).
//
//   la_mask_downscale  : prompt masks -> 16-channel map at 1/4 resolution: Conv2d(1,4,2,2) -> LayerNorm2d -> GELU ->
//                        Conv2d(4,16,2,2) -> LayerNorm2d -> GELU, one thread per output pixel.
//                        reference: label_anything/models/prompt_encoder.py:61-69 (mask_downscaling[0..5]), 516-540
//   la_resize_bilinear : token-major [S, h, w, c] fp32 bilinear resize (align_corners=False); used when the mask grid
//                        (64x64) differs from the feature grid (30x30 for 480 px models).  prompt_encoder.py:787-793
//   la_build_src       : src[s, t, :] = feat[(b,m), t, :] + W6 . m16[s, t, :] + b6 (+ class code[c]) -> bf16 rows:
//                        the last 1x1 conv of mask_downscaling fused with the support-feature broadcast and the
//                        RandomMatrixEncoder class code, written once (never materialised in fp32).
//                        prompt_encoder.py:68 (mask_downscaling[6]), 532-539 (null masks), 795-805, 250-264
//   la_embed_sparse    : Fourier positional encoding of points / box corners + label dependent embeddings.
//                        prompt_encoder.py:83-114, 201-211, 226-233, 648-669
//   la_masked_mean     : class embeddings = flag-masked mean over the examples.  prompt_encoder.py:738-745
//   la_classify        : logits[b, c, p] = <class_mlp(class token)[b, c, :], upscaled[b, p, :]>.  mask_decoder.py:309
//   la_postprocess_masks : bilinear to image_size, crop the un-padded region, bilinear to the query's original size,
//                        pad to the batch maximum with -inf (background channel: 0), absent classes -> -inf.
//                        label_anything/models/lam.py:383-453, 92-93
#include "la_common.cuh"
#ifndef LA_BUILD_SRC_FMA
#define LA_BUILD_SRC_FMA 0      // experiment builds: 1 = FFMA2 kernel for every width
#endif
#include <cstdlib>

namespace la {

static int grid_for(long long work_items, int block, int per_sm) {
  long long blocks = (work_items + block - 1) / block;
  const long long cap = static_cast<long long>(sm_count()) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

// ------------------------------------------------------------------------------------------------------
// mask downscaling
// ------------------------------------------------------------------------------------------------------
struct MaskDownWeights {
  float w0[4][4];     // [c1][ky*2+kx]
  float b0[4], g1[4], be1[4];
  float w3[16][16];   // [c2][c1*4 + ky*2 + kx]
  float b3[16], g2[16], be2[16];
  float eps1, eps2;
};

__global__ void __launch_bounds__(256)
mask_downscale_kernel(const float* __restrict__ masks, float* __restrict__ out, long long n_seq, int Hm, int Wm,
                      const MaskDownWeights W) {
  const int oh = Hm >> 2, ow = Wm >> 2;
  const long long total = n_seq * oh * ow;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(i % ow);
    const long long r = i / ow;
    const int oy = static_cast<int>(r % oh);
    const long long s = r / oh;
    const float* src = masks + (s * Hm + 4 * oy) * static_cast<long long>(Wm) + 4 * ox;
    float in[4][4];
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src + static_cast<long long>(y) * Wm));
      in[y][0] = v.x; in[y][1] = v.y; in[y][2] = v.z; in[y][3] = v.w;
    }
    float a1[4][4];  // [sub-block sy*2+sx][c1]
#pragma unroll
    for (int sb = 0; sb < 4; ++sb) {
      const int sy = sb >> 1, sx = sb & 1;
      float t[4], mean = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float acc = W.b0[c];
#pragma unroll
        for (int k = 0; k < 4; ++k) acc = fmaf(W.w0[c][k], in[2 * sy + (k >> 1)][2 * sx + (k & 1)], acc);
        t[c] = acc;
        mean += acc;
      }
      mean *= 0.25f;
      float var = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) var += (t[c] - mean) * (t[c] - mean);
      const float rstd = rsqrtf(var * 0.25f + W.eps1);
#pragma unroll
      for (int c = 0; c < 4; ++c) a1[sb][c] = gelu_erf((t[c] - mean) * rstd * W.g1[c] + W.be1[c]);
    }
    float t2[16], mean = 0.f;
#pragma unroll
    for (int c2 = 0; c2 < 16; ++c2) {
      float acc = W.b3[c2];
#pragma unroll
      for (int c1 = 0; c1 < 4; ++c1)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc = fmaf(W.w3[c2][c1 * 4 + k], a1[k][c1], acc);
      t2[c2] = acc;
      mean += acc;
    }
    mean *= (1.f / 16.f);
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) var += (t2[c] - mean) * (t2[c] - mean);
    const float rstd = rsqrtf(var * (1.f / 16.f) + W.eps2);
    float4* dst = reinterpret_cast<float4*>(out + i * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 o;
      o.x = gelu_erf((t2[4 * q + 0] - mean) * rstd * W.g2[4 * q + 0] + W.be2[4 * q + 0]);
      o.y = gelu_erf((t2[4 * q + 1] - mean) * rstd * W.g2[4 * q + 1] + W.be2[4 * q + 1]);
      o.z = gelu_erf((t2[4 * q + 2] - mean) * rstd * W.g2[4 * q + 2] + W.be2[4 * q + 2]);
      o.w = gelu_erf((t2[4 * q + 3] - mean) * rstd * W.g2[4 * q + 3] + W.be2[4 * q + 3]);
      dst[q] = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// bilinear resize, token-major (align_corners = False, PyTorch semantics)
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bilinear_tap(int dst, float scale, int in_size, int& i0, int& i1, float& lam) {
  float src = (dst + 0.5f) * scale - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = static_cast<int>(src);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  lam = src - i0;
}

__global__ void __launch_bounds__(256)
resize_bilinear_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, int ih, int iw, int oh,
                       int ow, int c4) {
  const float sy = static_cast<float>(ih) / oh, sx = static_cast<float>(iw) / ow;
  const long long total = n * oh * ow * c4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % c4);
    long long r = i / c4;
    const int x = static_cast<int>(r % ow);
    r /= ow;
    const int y = static_cast<int>(r % oh);
    const long long s = r / oh;
    int y0, y1, x0, x1;
    float ly, lx;
    bilinear_tap(y, sy, ih, y0, y1, ly);
    bilinear_tap(x, sx, iw, x0, x1, lx);
    const float4* base = reinterpret_cast<const float4*>(in) + s * ih * iw * c4 + c;
    const float4 a = __ldg(base + (static_cast<long long>(y0) * iw + x0) * c4);
    const float4 b = __ldg(base + (static_cast<long long>(y0) * iw + x1) * c4);
    const float4 cc = __ldg(base + (static_cast<long long>(y1) * iw + x0) * c4);
    const float4 d = __ldg(base + (static_cast<long long>(y1) * iw + x1) * c4);
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    float4 o;
    o.x = w00 * a.x + w01 * b.x + w10 * cc.x + w11 * d.x;
    o.y = w00 * a.y + w01 * b.y + w10 * cc.y + w11 * d.y;
    o.z = w00 * a.z + w01 * b.z + w10 * cc.z + w11 * d.z;
    o.w = w00 * a.w + w01 * b.w + w10 * cc.w + w11 * d.w;
    reinterpret_cast<float4*>(out)[i] = o;
  }
}

// ------------------------------------------------------------------------------------------------------
// src construction
// ------------------------------------------------------------------------------------------------------
struct BuildSrcParams {
  const float* feat;           // [n_img * T, D] fp32 support features (token-major)
  const float* m16;            // [S, T, 16] fp32 or nullptr (no mask prompts at all)
  const unsigned char* mflag;  // [S] or nullptr: 0 -> null mask -> not_a_mask vector
  const float* w6;             // [D, 16]
  const float* b6;             // [D]
  const float* not_a_mask;     // [D]
  const float* no_mask;        // [D]   (m16 == nullptr)
  const float* code;           // [C, D] or nullptr
  __nv_bfloat16* out;          // [S * T, D]
  long long n_seq;
  int T, D, C;
  int M, lead;                 // feature image of sequence s: (bm / M) * (M + lead) + lead + bm % M, bm = s / C
};

// One thread owns 4 channels and keeps its 4x16 slice of W6 in registers; a CTA walks the rows of one sequence.
// (d0, d1) += (a0, a1) * (b, b)   (sm_100 FFMA2)
__device__ __forceinline__ void fma2_pair(float& d0, float& d1, float a0, float a1, float b) {
  asm("{\n\t"
      ".reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %4};\n\t"
      "mov.b64 rd, {%0, %1};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rd;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t"
      "}"
      : "+f"(d0), "+f"(d1)
      : "f"(a0), "f"(a1), "f"(b));
}

__global__ void __launch_bounds__(256) build_src_kernel(const BuildSrcParams p) {
  const int tpr = p.D >> 2;                 // threads per row
  const int rows_per_pass = blockDim.x / tpr;
  const int lr = threadIdx.x / tpr;         // row slot inside the pass
  const int c0 = (threadIdx.x % tpr) * 4;
  const long long s = blockIdx.y;           // sequence index
  const int chunk = blockIdx.x;             // row chunk inside the sequence (grid.x)
  const int chunks = gridDim.x;
  const int per = (p.T + chunks - 1) / chunks;
  const int t_begin = chunk * per;
  const int t_end = min(p.T, t_begin + per);
  const int c = static_cast<int>(s % p.C);
  const long long bm_seq = s / p.C;
  const long long bm = (bm_seq / p.M) * (p.M + p.lead) + p.lead + bm_seq % p.M;

  const bool has_map = p.m16 != nullptr && (p.mflag == nullptr || p.mflag[s] != 0);
  float base[4];
  {
    const float* vec = p.m16 == nullptr ? p.no_mask : (has_map ? p.b6 : p.not_a_mask);
    const float4 v = __ldg(reinterpret_cast<const float4*>(vec + c0));
    base[0] = v.x; base[1] = v.y; base[2] = v.z; base[3] = v.w;
    if (p.code) {
      const float4 e = __ldg(reinterpret_cast<const float4*>(p.code + static_cast<long long>(c) * p.D + c0));
      base[0] += e.x; base[1] += e.y; base[2] += e.z; base[3] += e.w;
    }
  }
  float w[4][16];
  if (has_map) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p.w6 + static_cast<long long>(c0 + a) * 16) + q);
        w[a][4 * q] = v.x; w[a][4 * q + 1] = v.y; w[a][4 * q + 2] = v.z; w[a][4 * q + 3] = v.w;
      }
  }
  // Rows are processed in blocks of SRC_BLK: the block's 16-channel mask rows are staged in shared memory by a
  // coalesced cooperative load (every thread of a row needs all 16 values), and each thread keeps 4 independent
  // feature loads in flight -- the kernel is latency bound otherwise (one 16-byte load per thread per row).
  constexpr int SRC_BLK = 64;
  constexpr int SRC_INFLIGHT = 8;   // independent 16-byte feature loads per thread
  __shared__ float4 s_m16[SRC_BLK][4];
  const float* feat_seq = p.feat + bm * p.T * static_cast<long long>(p.D) + c0;
  __nv_bfloat16* out_seq = p.out + s * p.T * static_cast<long long>(p.D) + c0;
  for (int blk = t_begin; blk < t_end; blk += SRC_BLK) {
    if (has_map) {
      __syncthreads();
      if (threadIdx.x < SRC_BLK * 4) {
        const int rr = threadIdx.x >> 2, q = threadIdx.x & 3;
        if (blk + rr < t_end)
          s_m16[rr][q] = __ldg(reinterpret_cast<const float4*>(p.m16 + (s * p.T + blk + rr) * 16) + q);
      }
      __syncthreads();
    }
    for (int r0 = lr; r0 < SRC_BLK; r0 += SRC_INFLIGHT * rows_per_pass) {
      float4 f[SRC_INFLIGHT];
#pragma unroll
      for (int u = 0; u < SRC_INFLIGHT; ++u) {
        const int t = blk + r0 + u * rows_per_pass;
        if (r0 + u * rows_per_pass < SRC_BLK && t < t_end)
          f[u] = __ldg(reinterpret_cast<const float4*>(feat_seq + static_cast<long long>(t) * p.D));
      }
#pragma unroll
      for (int u = 0; u < SRC_INFLIGHT; ++u) {
        const int rl = r0 + u * rows_per_pass;
        const int t = blk + rl;
        if (rl < SRC_BLK && t < t_end) {
          float o[4] = {f[u].x + base[0], f[u].y + base[1], f[u].z + base[2], f[u].w + base[3]};
          if (has_map) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 m = s_m16[rl][q];
#pragma unroll
              for (int a = 0; a < 4; a += 2) {   // packed fp32x2: two output channels per instruction
                fma2_pair(o[a], o[a + 1], w[a][4 * q], w[a + 1][4 * q], m.x);
                fma2_pair(o[a], o[a + 1], w[a][4 * q + 1], w[a + 1][4 * q + 1], m.y);
                fma2_pair(o[a], o[a + 1], w[a][4 * q + 2], w[a + 1][4 * q + 2], m.z);
                fma2_pair(o[a], o[a + 1], w[a][4 * q + 3], w[a + 1][4 * q + 3], m.w);
              }
            }
          }
          uint2 pk;
          pk.x = pack_bf16(o[0], o[1]);
          pk.y = pack_bf16(o[2], o[3]);
          *reinterpret_cast<uint2*>(out_seq + static_cast<long long>(t) * p.D) = pk;
        }
      }
    }
  }
}

// Tensor-core variant (D % 64 == 0, D <= 512).  The 16 -> D projection is a K = 16 matrix product: as FFMA2 it costs 85
// instructions per warp-row and the kernel is issue bound at 1.8 TB/s (profiles/r01_ncu_prompt_v4.txt); as
// mma.m16n8k8 TF32 (10-bit mantissas, fp32 accumulation: error 2^-11 per product, below the bf16 rounding of the
// output) it is 16 MMAs per 16 rows x 64 channels.  A warp owns 64 channels, a CTA's D / 64 warps walk the same
// 16-row tiles.  The column index n of an MMA tile is free, so tile j's column n stands for channel
// 64 w + 16 (n >> 1) + 2 j + (n & 1): a lane's accumulators (columns 2t, 2t+1 of 8 tiles) are then 16 CONTIGUOUS
// channels -- the fp32 feature rows are loaded straight into the accumulators as four 16-byte loads and the bf16 row
// leaves as two 16-byte stores.
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256) build_src_mma_kernel(const BuildSrcParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const long long s = blockIdx.y;
  const int chunks = gridDim.x;
  const int per = (((p.T + chunks - 1) / chunks) + 15) & ~15;   // whole 16-row tiles per chunk
  const int t_begin = blockIdx.x * per;
  const int t_end = min(p.T, t_begin + per);
  if (t_begin >= t_end) return;
  const int c = static_cast<int>(s % p.C);
  const long long bm_seq = s / p.C;
  const long long bm = (bm_seq / p.M) * (p.M + p.lead) + p.lead + bm_seq % p.M;
  const bool has_map = p.m16 != nullptr && (p.mflag == nullptr || p.mflag[s] != 0);
  const int col0 = warp * 64 + t * 16;   // this lane's 16 contiguous channels

  float base[16];
  {
    const float* vec = p.m16 == nullptr ? p.no_mask : (has_map ? p.b6 : p.not_a_mask);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 v = __ldg(reinterpret_cast<const float4*>(vec + col0) + i);
      if (p.code) {
        const float4 e = __ldg(reinterpret_cast<const float4*>(p.code + static_cast<long long>(c) * p.D + col0) + i);
        v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
      }
      base[4 * i] = v.x; base[4 * i + 1] = v.y; base[4 * i + 2] = v.z; base[4 * i + 3] = v.w;
    }
  }
  // B fragments: tile j, k-step q: b0 = W[k = 8q + t][n = g], b1 = W[k = 8q + t + 4][n = g], W[k][n] = w6[channel(j, n)][k]
  uint32_t bf[8][2][2];
  if (has_map) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float* wrow = p.w6 + static_cast<long long>(warp * 64 + (g >> 1) * 16 + 2 * j + (g & 1)) * 16;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        bf[j][q][0] = to_tf32(__ldg(wrow + 8 * q + t));
        bf[j][q][1] = to_tf32(__ldg(wrow + 8 * q + t + 4));
      }
    }
  }
  const float* feat_seq = p.feat + bm * p.T * static_cast<long long>(p.D) + col0;
  __nv_bfloat16* out_seq = p.out + s * p.T * static_cast<long long>(p.D) + col0;
  const float* m16_seq = p.m16 != nullptr ? p.m16 + s * p.T * 16 : nullptr;

  for (int tile0 = t_begin; tile0 < t_end; tile0 += 16) {
    const int r_lo = tile0 + g, r_hi = tile0 + g + 8;
    const bool v_lo = r_lo < t_end, v_hi = r_hi < t_end;
    // acc[j][e] = row g, channel col0 + 2j + e;  acc[j][2 + e] = row g + 8, same channel
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = make_float4(0.f, 0.f, 0.f, 0.f);
      if (v_lo) lo = __ldg(reinterpret_cast<const float4*>(feat_seq + static_cast<long long>(r_lo) * p.D) + i);
      if (v_hi) hi = __ldg(reinterpret_cast<const float4*>(feat_seq + static_cast<long long>(r_hi) * p.D) + i);
      acc[2 * i][0] = lo.x + base[4 * i];
      acc[2 * i][1] = lo.y + base[4 * i + 1];
      acc[2 * i + 1][0] = lo.z + base[4 * i + 2];
      acc[2 * i + 1][1] = lo.w + base[4 * i + 3];
      acc[2 * i][2] = hi.x + base[4 * i];
      acc[2 * i][3] = hi.y + base[4 * i + 1];
      acc[2 * i + 1][2] = hi.z + base[4 * i + 2];
      acc[2 * i + 1][3] = hi.w + base[4 * i + 3];
    }
    if (has_map) {
      // A fragments: a0 = (row g, k = t), a1 = (row g + 8, k = t), a2 = (row g, k = t + 4), a3 = (row g + 8, k = t + 4)
      uint32_t a[2][4];
      const float* m_lo = m16_seq + static_cast<long long>(r_lo) * 16;
      const float* m_hi = m16_seq + static_cast<long long>(r_hi) * 16;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        a[q][0] = v_lo ? to_tf32(__ldg(m_lo + 8 * q + t)) : 0u;
        a[q][1] = v_hi ? to_tf32(__ldg(m_hi + 8 * q + t)) : 0u;
        a[q][2] = v_lo ? to_tf32(__ldg(m_lo + 8 * q + t + 4)) : 0u;
        a[q][3] = v_hi ? to_tf32(__ldg(m_hi + 8 * q + t + 4)) : 0u;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        mma_tf32_16x8x8(acc[j], a[0], bf[j][0][0], bf[j][0][1]);
        mma_tf32_16x8x8(acc[j], a[1], bf[j][1][0], bf[j][1][1]);
      }
    }
    if (v_lo) {
      uint4* dst = reinterpret_cast<uint4*>(out_seq + static_cast<long long>(r_lo) * p.D);
      dst[0] = make_uint4(pack_bf16(acc[0][0], acc[0][1]), pack_bf16(acc[1][0], acc[1][1]),
                          pack_bf16(acc[2][0], acc[2][1]), pack_bf16(acc[3][0], acc[3][1]));
      dst[1] = make_uint4(pack_bf16(acc[4][0], acc[4][1]), pack_bf16(acc[5][0], acc[5][1]),
                          pack_bf16(acc[6][0], acc[6][1]), pack_bf16(acc[7][0], acc[7][1]));
    }
    if (v_hi) {
      uint4* dst = reinterpret_cast<uint4*>(out_seq + static_cast<long long>(r_hi) * p.D);
      dst[0] = make_uint4(pack_bf16(acc[0][2], acc[0][3]), pack_bf16(acc[1][2], acc[1][3]),
                          pack_bf16(acc[2][2], acc[2][3]), pack_bf16(acc[3][2], acc[3][3]));
      dst[1] = make_uint4(pack_bf16(acc[4][2], acc[4][3]), pack_bf16(acc[5][2], acc[5][3]),
                          pack_bf16(acc[6][2], acc[6][3]), pack_bf16(acc[7][2], acc[7][3]));
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// sparse prompt tokens
// ------------------------------------------------------------------------------------------------------
struct SparseParams {
  const float* points;   // [S, P, 2] (x, y) pixels or nullptr
  const float* plabels;  // [S, P] in {1, 0, -1}
  const float* boxes;    // [S, Bx, 4] (x1, y1, x2, y2) or nullptr
  const float* bflags;   // [S, Bx]
  const float* gauss;    // [2, D/2]
  const float* not_a_point;  // [D]
  const float* pe_tab;   // [4, D]: negative point, positive point, box corner 0, box corner 1
  float* out;            // [S, n, D]
  long long n_seq;
  int P, Bx, pad, n, D;
  float inv_w, inv_h;
};

__global__ void __launch_bounds__(256) embed_sparse_kernel(const SparseParams p) {
  const int half = p.D >> 1;
  const long long total = p.n_seq * p.n * half;
  const int n_pts = p.points ? p.P + p.pad : 0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(i % half);
    const long long r = i / half;
    const int tok = static_cast<int>(r % p.n);
    const long long s = r / p.n;
    float x, y, label;
    const float* add = nullptr;
    bool null_tok = false;
    if (tok < n_pts) {
      if (tok < p.P) {
        x = p.points[(s * p.P + tok) * 2] + 0.5f;
        y = p.points[(s * p.P + tok) * 2 + 1] + 0.5f;
        label = p.plabels[s * p.P + tok];
      } else {  // the padding point appended when there are no boxes: (0, 0), label -1
        x = 0.f;
        y = 0.f;
        label = -1.f;
      }
      null_tok = label == 0.f;
      if (label == -1.f) add = p.pe_tab;
      if (label == 1.f) add = p.pe_tab + p.D;
    } else {
      const int cj = tok - n_pts;  // corner slot in [0, 2*Bx)
      const int bx = cj >> 1, corner = cj & 1;
      x = p.boxes[(s * p.Bx + bx) * 4 + corner * 2] + 0.5f;
      y = p.boxes[(s * p.Bx + bx) * 4 + corner * 2 + 1] + 0.5f;
      add = p.pe_tab + (2 + corner) * p.D;
      // the reference indexes the (box, corner)-flattened axis with flags.repeat(2): slot j <- flags[j % Bx]
      null_tok = p.bflags[s * p.Bx + (cj % p.Bx)] == 0.f;
    }
    float o_sin, o_cos;
    if (null_tok) {
      o_sin = p.not_a_point[j];
      o_cos = p.not_a_point[half + j];
    } else {
      const float cx = 2.f * (x * p.inv_w) - 1.f, cy = 2.f * (y * p.inv_h) - 1.f;
      const float proj = 6.283185307179586f * (cx * p.gauss[j] + cy * p.gauss[half + j]);
      sincosf(proj, &o_sin, &o_cos);
      if (add) {
        o_sin += add[j];
        o_cos += add[half + j];
      }
    }
    p.out[r * p.D + j] = o_sin;
    p.out[r * p.D + half + j] = o_cos;
  }
}

// out[b, c, :] = sum_m flag[b,m,c] * emb[b,m,c,:] / max(sum_m flag[b,m,c], 1)
__global__ void __launch_bounds__(256)
masked_mean_kernel(const float* __restrict__ emb, const unsigned char* __restrict__ flag, float* __restrict__ out,
                   int B, int M, int C, int D) {
  const long long total = static_cast<long long>(B) * C * D;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(i % D);
    const long long r = i / D;
    const int c = static_cast<int>(r % C);
    const long long b = r / C;
    float acc = 0.f, cnt = 0.f;
    for (int m = 0; m < M; ++m) {
      const float f = flag[(b * M + m) * C + c] ? 1.f : 0.f;
      acc = fmaf(f, emb[((b * M + m) * C + c) * static_cast<long long>(D) + d], acc);
      cnt += f;
    }
    out[i] = acc / (cnt == 0.f ? 1.f : cnt);
  }
}

// ------------------------------------------------------------------------------------------------------
// classify: logits[b, c, p] = sum_d cls[b, c, d] * x[b, p, d]
// ------------------------------------------------------------------------------------------------------
constexpr int CLS_MAX_C = 32;

__global__ void __launch_bounds__(256)
classify_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ cls, float* __restrict__ out,
                long long P, int C, int dk, int c_total, int c_off) {
  extern __shared__ float s_cls[];  // [C][dk]
  const long long b = blockIdx.y;
  for (int i = threadIdx.x; i < C * dk; i += blockDim.x) s_cls[i] = cls[(b * c_total + c_off) * dk + i];
  __syncthreads();
  for (long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; pix < P;
       pix += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4* row = reinterpret_cast<const uint4*>(x + (b * P + pix) * dk);
    float acc[CLS_MAX_C];
#pragma unroll
    for (int c = 0; c < CLS_MAX_C; ++c) acc[c] = 0.f;
    for (int v = 0; v < dk / 8; ++v) {
      const uint4 u = __ldg(row + v);
      const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
      float f[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        f[2 * q] = __low2float(hh[q]);
        f[2 * q + 1] = __high2float(hh[q]);
      }
#pragma unroll
      for (int c = 0; c < CLS_MAX_C; ++c) {
        if (c < C) {
          const float* w = s_cls + c * dk + v * 8;
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[c] = fmaf(f[q], w[q], acc[c]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < CLS_MAX_C; ++c)
      if (c < C) out[(b * c_total + c_off + c) * P + pix] = acc[c];
  }
}

// ------------------------------------------------------------------------------------------------------
// postprocess
// ------------------------------------------------------------------------------------------------------
struct PostParams {
  const float* in;          // [B, C, lh, lw]
  float* out;               // [B, C, Hmax, Wmax]
  const int* sizes;         // [B, 4]: original (oh, ow), un-padded model input (ih, iw)
  const unsigned char* flag_gts;  // [B, C] or nullptr
  int B, C, lh, lw, S, Hmax, Wmax;
};

__device__ __forceinline__ float sample_stage1(const float* __restrict__ plane, int lh, int lw, int S, int y, int x) {
  // value of the (lh, lw) -> (S, S) bilinear upsampling at integer position (y, x)
  int y0, y1, x0, x1;
  float ly, lx;
  bilinear_tap(y, static_cast<float>(lh) / S, lh, y0, y1, ly);
  bilinear_tap(x, static_cast<float>(lw) / S, lw, x0, x1, lx);
  const float a = __ldg(plane + y0 * lw + x0), b = __ldg(plane + y0 * lw + x1);
  const float c = __ldg(plane + y1 * lw + x0), d = __ldg(plane + y1 * lw + x1);
  // PyTorch upsample_bilinear2d: w00*a + w01*b + w10*c + w11*d with w = (1-ly)(1-lx) ...
  return (1.f - ly) * ((1.f - lx) * a + lx * b) + ly * ((1.f - lx) * c + lx * d);
}

__global__ void __launch_bounds__(256) postprocess_kernel(const PostParams p) {
  const long long plane_out = static_cast<long long>(p.Hmax) * p.Wmax;
  const long long total = static_cast<long long>(p.B) * p.C * plane_out;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % p.Wmax);
    long long r = i / p.Wmax;
    const int y = static_cast<int>(r % p.Hmax);
    r /= p.Hmax;
    const int c = static_cast<int>(r % p.C);
    const int b = static_cast<int>(r / p.C);
    const int oh = p.sizes[b * 4], ow = p.sizes[b * 4 + 1], ih = p.sizes[b * 4 + 2], iw = p.sizes[b * 4 + 3];
    float v;
    if (p.flag_gts && p.flag_gts[b * p.C + c] == 0) {
      v = -INFINITY;
    } else if (y >= oh || x >= ow) {
      v = (c == 0) ? 0.f : -INFINITY;
    } else {
      const float* plane = p.in + (static_cast<long long>(b) * p.C + c) * p.lh * p.lw;
      int y0, y1, x0, x1;
      float ly, lx;
      bilinear_tap(y, static_cast<float>(ih) / oh, ih, y0, y1, ly);
      bilinear_tap(x, static_cast<float>(iw) / ow, iw, x0, x1, lx);
      const float a = sample_stage1(plane, p.lh, p.lw, p.S, y0, x0);
      const float bb = sample_stage1(plane, p.lh, p.lw, p.S, y0, x1);
      const float cc = sample_stage1(plane, p.lh, p.lw, p.S, y1, x0);
      const float d = sample_stage1(plane, p.lh, p.lw, p.S, y1, x1);
      v = (1.f - ly) * ((1.f - lx) * a + lx * bb) + ly * ((1.f - lx) * cc + lx * d);
    }
    p.out[i] = v;
  }
}

// Four horizontally adjacent output pixels per thread: index arithmetic by multiply-high divisions, the row taps and
// the per-episode scale factors once per thread, one 16-byte store.  Pixel arithmetic is the scalar kernel's, call for
// call, so the two give identical bits.  (The scalar kernel spends ~600 instructions per pixel, four 64-bit divisions
// among them: 0.91 ms for 8 x 6 x 1024^2 outputs = 0.23 TB/s; `out_w % 4 == 0` and < 2^31 quads take this one.)
struct PostFast {
  FastDiv w4, h, c;
};
__global__ void __launch_bounds__(256) postprocess_quad_kernel(const PostParams p, const PostFast f) {
  const uint32_t w4 = static_cast<uint32_t>(p.Wmax) >> 2;
  const uint32_t total = static_cast<uint32_t>(p.B) * p.C * p.Hmax * w4;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t r1 = fast_div(i, f.w4);
    const int x4 = static_cast<int>(i - r1 * w4) * 4;
    const uint32_t r2 = fast_div(r1, f.h);
    const int y = static_cast<int>(r1 - r2 * p.Hmax);
    const uint32_t bq = fast_div(r2, f.c);
    const int c = static_cast<int>(r2 - bq * p.C), b = static_cast<int>(bq);
    const int4 sz = __ldg(reinterpret_cast<const int4*>(p.sizes) + b);
    const int oh = sz.x, ow = sz.y, ih = sz.z, iw = sz.w;
    float4 v;
    float* vv = reinterpret_cast<float*>(&v);
    const float pad = (c == 0) ? 0.f : -INFINITY;
    if (p.flag_gts && p.flag_gts[b * p.C + c] == 0) {
      v = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    } else if (y >= oh || x4 >= ow) {
      v = make_float4(pad, pad, pad, pad);
    } else {
      const float* plane = p.in + (static_cast<long long>(b) * p.C + c) * p.lh * p.lw;
      int y0, y1;
      float ly;
      bilinear_tap(y, static_cast<float>(ih) / oh, ih, y0, y1, ly);
      const float sx = static_cast<float>(iw) / ow;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int x = x4 + k;
        if (x >= ow) {
          vv[k] = pad;
        } else {
          int x0, x1;
          float lx;
          bilinear_tap(x, sx, iw, x0, x1, lx);
          const float a = sample_stage1(plane, p.lh, p.lw, p.S, y0, x0);
          const float bb = sample_stage1(plane, p.lh, p.lw, p.S, y0, x1);
          const float cc = sample_stage1(plane, p.lh, p.lw, p.S, y1, x0);
          const float d = sample_stage1(plane, p.lh, p.lw, p.S, y1, x1);
          vv[k] = (1.f - ly) * ((1.f - lx) * a + lx * bb) + ly * ((1.f - lx) * cc + lx * d);
        }
      }
    }
    reinterpret_cast<float4*>(p.out)[i] = v;
  }
}

}  // namespace la

extern "C" {

int la_mask_downscale(void* stream, const float* masks, float* out, long long n_seq, int height, int width,
                      const float* w0, const float* b0, const float* ln1_w, const float* ln1_b, float eps1,
                      const float* w3, const float* b3, const float* ln2_w, const float* ln2_b, float eps2) {
  using namespace la;
  LA_CHECK_ARG(masks && out && w0 && b0 && ln1_w && ln1_b && w3 && b3 && ln2_w && ln2_b, "la_mask_downscale: null pointer");
  LA_CHECK_ARG(n_seq > 0 && height % 4 == 0 && width % 4 == 0 && height > 0 && width > 0,
               "la_mask_downscale: mask size must be a positive multiple of 4");
  // the 352 weight floats live in HOST memory here (they are packed once per model by the caller)
  MaskDownWeights W;
  for (int c = 0; c < 4; ++c) {
    for (int k = 0; k < 4; ++k) W.w0[c][k] = w0[c * 4 + k];
    W.b0[c] = b0[c];
    W.g1[c] = ln1_w[c];
    W.be1[c] = ln1_b[c];
  }
  for (int c = 0; c < 16; ++c) {
    for (int k = 0; k < 16; ++k) W.w3[c][k] = w3[c * 16 + k];
    W.b3[c] = b3[c];
    W.g2[c] = ln2_w[c];
    W.be2[c] = ln2_b[c];
  }
  W.eps1 = eps1;
  W.eps2 = eps2;
  const long long total = n_seq * (height / 4) * (width / 4);
  mask_downscale_kernel<<<grid_for(total, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(masks, out, n_seq,
                                                                                              height, width, W);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_resize_bilinear(void* stream, const float* in, float* out, long long n, int in_h, int in_w, int out_h,
                       int out_w, int channels) {
  using namespace la;
  LA_CHECK_ARG(in && out && n > 0 && in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0 && channels % 4 == 0,
               "la_resize_bilinear: bad arguments");
  const long long total = n * out_h * out_w * (channels / 4);
  resize_bilinear_kernel<<<grid_for(total, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, out, n, in_h, in_w, out_h, out_w, channels / 4);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_build_src(void* stream, const float* feat, const float* m16, const unsigned char* mask_flags,
                 const float* w6, const float* b6, const float* not_a_mask, const float* no_mask, const float* code,
                 void* out, long long n_seq, int tokens, int d, int n_classes, int examples, int feat_lead) {
  using namespace la;
  LA_CHECK_ARG(feat && out && n_seq > 0 && tokens > 0 && n_classes > 0, "la_build_src: bad arguments");
  LA_CHECK_ARG(examples > 0 && feat_lead >= 0, "la_build_src: bad examples / feat_lead");
  LA_CHECK_ARG(d % 4 == 0 && d >= 32 && d <= 1024, "la_build_src: d=%d unsupported (multiple of 4 in [32, 1024])", d);
  LA_CHECK_ARG(m16 ? (w6 && b6 && not_a_mask) : (no_mask != nullptr), "la_build_src: missing dense-embedding weights");
  LA_CHECK_ARG(n_seq <= 65535ll * 65535ll, "la_build_src: too many sequences");
  BuildSrcParams p;
  p.feat = feat;
  p.m16 = m16;
  p.mflag = mask_flags;
  p.w6 = w6;
  p.b6 = b6;
  p.not_a_mask = not_a_mask;
  p.no_mask = no_mask;
  p.code = code;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.n_seq = n_seq;
  p.T = tokens;
  p.D = d;
  p.C = n_classes;
  p.M = examples;
  p.lead = feat_lead;
  const int tpr = d / 4;
  int threads = tpr >= 256 ? tpr : (256 / tpr) * tpr;
  // row chunks per sequence: enough CTAs for >= 16 waves of the 2 resident CTAs per SM, so that the ragged last
  // wave costs a few percent (one CTA per sequence left a 5th wave of 16 CTAs behind 4 full ones: 20 %)
  long long chunks = (32ll * sm_count() + n_seq - 1) / n_seq;
  const long long max_chunks = (tokens + 63) / 64;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  LA_CHECK_ARG(n_seq <= 65535, "la_build_src: more than 65535 sequences per call (chunk the call)");
  dim3 grid(static_cast<unsigned>(chunks), static_cast<unsigned>(n_seq));
  // tensor-core variant: one warp per 64 channels (-DLA_BUILD_SRC_FMA=1 builds keep the FFMA2 kernel)
  if (d % 64 == 0 && d <= 512 && !LA_BUILD_SRC_FMA) {
    build_src_mma_kernel<<<grid, d / 2, 0, static_cast<cudaStream_t>(stream)>>>(p);
    LA_CHECK_CUDA(cudaGetLastError());
    return LA_OK;
  }
  build_src_kernel<<<grid, threads, 0, static_cast<cudaStream_t>(stream)>>>(p);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_embed_sparse(void* stream, const float* points, const float* point_labels, int n_points, const float* boxes,
                    const float* box_flags, int n_boxes, const float* gauss, const float* not_a_point,
                    const float* pe_table, float* out, long long n_seq, int d, int image_w, int image_h) {
  using namespace la;
  LA_CHECK_ARG(out && gauss && not_a_point && pe_table && n_seq > 0 && d % 2 == 0, "la_embed_sparse: bad arguments");
  LA_CHECK_ARG((points && point_labels && n_points > 0) || (boxes && box_flags && n_boxes > 0),
               "la_embed_sparse: no prompts");
  SparseParams p;
  p.points = points;
  p.plabels = point_labels;
  p.boxes = boxes;
  p.bflags = box_flags;
  p.gauss = gauss;
  p.not_a_point = not_a_point;
  p.pe_tab = pe_table;
  p.out = out;
  p.n_seq = n_seq;
  p.P = points ? n_points : 0;
  p.Bx = boxes ? n_boxes : 0;
  p.pad = (points && !boxes) ? 1 : 0;
  p.n = (points ? p.P + p.pad : 0) + 2 * p.Bx;
  p.D = d;
  p.inv_w = 1.0f / image_w;
  p.inv_h = 1.0f / image_h;
  const long long total = n_seq * p.n * (d / 2);
  embed_sparse_kernel<<<grid_for(total, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_masked_mean(void* stream, const float* emb, const unsigned char* flags, float* out, int batch, int examples,
                   int classes, int d) {
  using namespace la;
  LA_CHECK_ARG(emb && flags && out && batch > 0 && examples > 0 && classes > 0 && d > 0, "la_masked_mean: bad arguments");
  const long long total = static_cast<long long>(batch) * classes * d;
  masked_mean_kernel<<<grid_for(total, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(emb, flags, out, batch,
                                                                                            examples, classes, d);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_classify(void* stream, const void* x, const float* cls, float* out, int batch, long long pixels, int classes,
                int dk) {
  using namespace la;
  LA_CHECK_ARG(x && cls && out && batch > 0 && pixels > 0, "la_classify: bad arguments");
  LA_CHECK_ARG(classes > 0, "la_classify: no classes");
  LA_CHECK_ARG(dk % 8 == 0 && dk > 0 && static_cast<size_t>(CLS_MAX_C) * dk * 4 <= 48 * 1024, "la_classify: bad dk=%d", dk);
  LA_CHECK_ARG(batch <= 65535, "la_classify: batch too large");
  long long bx = (pixels + 255) / 256;
  const long long cap = (8ll * sm_count() + batch - 1) / batch;
  if (bx > cap) bx = cap;
  dim3 grid(static_cast<unsigned>(bx), static_cast<unsigned>(batch));
  for (int c_off = 0; c_off < classes; c_off += CLS_MAX_C) {  // 32 classes per pass
    const int cc = classes - c_off < CLS_MAX_C ? classes - c_off : CLS_MAX_C;
    classify_kernel<<<grid, 256, static_cast<size_t>(cc) * dk * 4, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), cls, out, pixels, cc, dk, classes, c_off);
    LA_CHECK_CUDA(cudaGetLastError());
  }
  return LA_OK;
}

int la_postprocess_masks(void* stream, const float* logits, float* out, const int* sizes,
                         const unsigned char* flag_gts, int batch, int classes, int low_h, int low_w, int image_size,
                         int out_h, int out_w) {
  using namespace la;
  LA_CHECK_ARG(logits && out && sizes && batch > 0 && classes > 0 && low_h > 0 && low_w > 0 && image_size > 0 &&
                   out_h > 0 && out_w > 0,
               "la_postprocess_masks: bad arguments");
  PostParams p;
  p.in = logits;
  p.out = out;
  p.sizes = sizes;
  p.flag_gts = flag_gts;
  p.B = batch;
  p.C = classes;
  p.lh = low_h;
  p.lw = low_w;
  p.S = image_size;
  p.Hmax = out_h;
  p.Wmax = out_w;
  const long long total = static_cast<long long>(batch) * classes * out_h * out_w;
  if (out_w % 4 == 0 && total / 4 < (1ll << 31) && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(sizes) & 15) == 0) {
    PostFast f;
    f.w4 = make_fastdiv(static_cast<uint32_t>(out_w / 4));
    f.h = make_fastdiv(static_cast<uint32_t>(out_h));
    f.c = make_fastdiv(static_cast<uint32_t>(classes));
    postprocess_quad_kernel<<<grid_for(total / 4, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, f);
    LA_CHECK_CUDA(cudaGetLastError());
    return LA_OK;
  }
  postprocess_kernel<<<grid_for(total, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

}  // extern "C"
