// labelanything_b200 — token -> image attention of the two-way transformer WITHOUT materialised K / V projections
// (sm_100a).
//
// Reference arithmetic (label_anything/models/transformer.py:311-318, common.py:97-148): per prompt sequence
//     k_t = W_k (x_t + pe_t) + b_k,  v_t = W_v x_t + b_v            for every image token t (T = 4096, D = 512)
//     o   = sum_t softmax_t(q_h . k_{h,t} / sqrt(dh)) v_{h,t}        per head h of the few query tokens
// The native path used to run the k / v projections as one GEMM over all S*T image tokens (2.6 TFLOP and 10 GB of HBM
// traffic per layer for 1200 sequences) only to reduce them against ONE query token per sequence.  By associativity
//     q_h . k_{h,t} = u_h . x_t + u_h . pe_t + const_h,   u_h = W_k[h]^T q_h  (a D-vector per (query, head))
//     o_h           = W_v[h] y_h + b_v[h],                 y_h = sum_t p_{h,t} x_t
// (const_h drops out of the softmax), so the image tokens only have to be read once, as they are: this kernel computes
//     y[s, r, :] = sum_t softmax_t(scale * (u[s, r] . x[s, t] + e[s, r, t])) x[s, t, :]      r < rows <= 8
// for every sequence s with x (bf16 [n_seq * tokens, d]) as BOTH the key and the value operand: a FlashAttention-style
// single pass with "head_dim" d and `rows` queries.  u ([n_seq * rows, d] bf16), the positional scores e = u . pe^T
// ([n_seq * rows, tokens] fp32) and the final W_v / out projections are small GEMMs on la_gemm_bf16
// (labelanything_b200/transformer.py::run_two_way).
//
// One CTA per sequence, one warp per 64-channel slab of x: every warp TMA-loads ITS slab of a 64-token tile (one
// 128B-swizzled [64 tokens x 64 channels] box, 3-deep ring, the warp is its own producer), forms its partial scores with
// mma.sync m16n8k16 (M = 16 query rows of which `rows` are real -- an 8-row problem has no use for a 128-row tcgen05
// tile), the partials are summed through shared memory, every warp runs the same online softmax and accumulates
// y[:, its 64 channels] with a second set of mma.sync (P from the score fragments, x^T through ldmatrix.trans).
// HBM-bound: x is read exactly once (algorithmic bytes = 2 * tokens * d per sequence).
#include "la_common.cuh"

namespace la {

constexpr int PA_TILE = 64;                 // tokens per tile
constexpr int PA_STAGES = 3;
constexpr int PA_SLAB = PA_TILE * 128;      // one [64 tokens x 64 channels] bf16 box
constexpr int PA_RED_STRIDE = 68;           // floats per (warp, row) of the partial-score exchange

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// D (rows g / g+8) += A (16 x 16) B (16 x 8); only rows 0..7 of A are non-zero here: a1 = a3 = 0
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float pa_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

struct PoolAttnParams {
  const __nv_bfloat16* u;   // [n_seq * rows, d]
  const float* e;           // [n_seq * rows, lde] or nullptr
  long long lde;            // row stride of e (>= tokens)
  __nv_bfloat16* y;         // [n_seq * rows, d]
  long long n_seq;
  int tokens, rows, d;
  float scale_log2;         // scale * log2(e)
};

template <int NW>
__global__ void __launch_bounds__(NW * 32)
pooled_attention_kernel(const __grid_constant__ CUtensorMap tm_x, const PoolAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // [stage][warp] slabs, then the partial-score exchange, then one mbarrier per (stage, warp)
  float* red = reinterpret_cast<float*>(smem + PA_STAGES * NW * PA_SLAB);
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + NW * 8 * PA_RED_STRIDE);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;           // fragment row / column pair
  const long long seq = blockIdx.x;
  const int n_tiles = (p.tokens + PA_TILE - 1) / PA_TILE;
  const long long row0 = seq * p.tokens;

  if (lane == 0) {
    for (int s = 0; s < PA_STAGES; ++s) mbar_init(&bars[s * NW + warp], 1);
    fence_barrier_init();
  }
  if (threadIdx.x == 0) tma_prefetch_desc(&tm_x);
  __syncthreads();

  auto issue = [&](int tile) {   // lane 0 of every warp loads the warp's own slab
    const int st = tile % PA_STAGES;
    uint64_t* bar = &bars[st * NW + warp];
    mbar_arrive_expect_tx(bar, PA_SLAB);
    tma_load_2d(smem + (st * NW + warp) * PA_SLAB, &tm_x, bar, warp * 64, static_cast<int32_t>(row0 + tile * PA_TILE));
  };
  if (lane == 0) {
    for (int t = 0; t < PA_STAGES && t < n_tiles; ++t) issue(t);
  }

  // query-side operand: A fragments of u[seq, g, 64 * warp + 16 kk + ...] (rows >= p.rows are zero)
  uint32_t ua0[4], ua2[4];
  {
    const bool live = g < p.rows;
    const __nv_bfloat16* ur = p.u + (seq * p.rows + (live ? g : 0)) * p.d + warp * 64 + 2 * q;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      ua0[kk] = live ? __ldg(reinterpret_cast<const uint32_t*>(ur + 16 * kk)) : 0u;
      ua2[kk] = live ? __ldg(reinterpret_cast<const uint32_t*>(ur + 16 * kk + 8)) : 0u;
    }
  }
  const float* erow = (p.e != nullptr && g < p.rows) ? p.e + (seq * p.rows + g) * p.lde : nullptr;

  float yacc[8][4];
#pragma unroll
  for (int c = 0; c < 8; ++c) yacc[c][0] = yacc[c][1] = yacc[c][2] = yacc[c][3] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;

  for (int i = 0; i < n_tiles; ++i) {
    const int st = i % PA_STAGES;
    mbar_wait(&bars[st * NW + warp], (i / PA_STAGES) & 1);
    const uint32_t slab = smem_u32(smem + (st * NW + warp) * PA_SLAB);

    // ---- partial scores of this warp's 64 channels: S^T[16 x 64 tokens] += U_w[16 x 64] X_w^T ----
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        // matrices: (tile j, k lo) (tile j, k hi) (tile j+1, k lo) (tile j+1, k hi); lane -> (matrix lane/8, row lane%8)
        const int mi = lane >> 3, rr = lane & 7;
        const int tok = 8 * (j + (mi >> 1)) + rr;
        const int chunk = 2 * kk + (mi & 1);
        uint32_t b[4];
        ldmatrix_x4(b, slab + tok * 128 + ((chunk ^ (tok & 7)) << 4));
        mma_16816(s[j], ua0[kk], ua2[kk], b[0], b[1]);
        mma_16816(s[j + 1], ua0[kk], ua2[kk], b[2], b[3]);
      }
    }
    // ---- sum the partials of all warps (rows 0..7 only: rows 8..15 of the A operand are zero) ----
    {
      float* mine = red + (warp * 8 + g) * PA_RED_STRIDE + 2 * q;
#pragma unroll
      for (int j = 0; j < 8; ++j) *reinterpret_cast<float2*>(mine + 8 * j) = make_float2(s[j][0], s[j][1]);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < NW; ++w2) {
        const float2 v = *reinterpret_cast<const float2*>(red + (w2 * 8 + g) * PA_RED_STRIDE + 8 * j + 2 * q);
        a0 += v.x;
        a1 += v.y;
      }
      s[j][0] = a0;
      s[j][1] = a1;
    }
    __syncthreads();   // `red` is rewritten by the next tile

    // ---- positional scores, scale, tail mask, online softmax (every warp computes the same statistics) ----
    const int t0 = i * PA_TILE + 2 * q;
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int t = t0 + 8 * j;
      float e0 = 0.f, e1 = 0.f;
      if (erow != nullptr) {
        if (t < p.tokens) e0 = __ldg(erow + t);
        if (t + 1 < p.tokens) e1 = __ldg(erow + t + 1);
      }
      s[j][0] = t < p.tokens ? (s[j][0] + e0) * p.scale_log2 : -INFINITY;
      s[j][1] = t + 1 < p.tokens ? (s[j][1] + e1) * p.scale_log2 : -INFINITY;
      mx = fmaxf(mx, fmaxf(s[j][0], s[j][1]));
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float m_new = fmaxf(m_run, mx);          // finite: every tile holds at least one real token
    const float alpha = pa_ex2(m_run - m_new);      // 0 on the first tile (m_run = -inf)
    float lsum = 0.f;
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = pa_ex2(s[j][0] - m_new), p1 = pa_ex2(s[j][1] - m_new);
      lsum += p0 + p1;
      pk[j] = pack_bf16(p0, p1);
    }
    lsum += __shfl_xor_sync(0xffffffffu, lsum, 1);
    lsum += __shfl_xor_sync(0xffffffffu, lsum, 2);
    l_run = l_run * alpha + lsum;
    m_run = m_new;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      yacc[c][0] *= alpha;
      yacc[c][1] *= alpha;
    }

    // ---- y[:, this warp's 64 channels] += P[16 x 64 tokens] X_w[64 tokens x 64 channels] ----
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        // matrices: (tokens 16kt.., chunk c) (tokens 16kt+8.., chunk c) (tokens 16kt.., chunk c+1) (tokens 16kt+8.., chunk c+1)
        const int mi = lane >> 3, rr = lane & 7;
        const int tok = 16 * kt + 8 * (mi & 1) + rr;
        const int chunk = c + (mi >> 1);
        uint32_t b[4];
        ldmatrix_x4_trans(b, slab + tok * 128 + ((chunk ^ (tok & 7)) << 4));
        mma_16816(yacc[c], pk[2 * kt], pk[2 * kt + 1], b[0], b[1]);
        mma_16816(yacc[c + 1], pk[2 * kt], pk[2 * kt + 1], b[2], b[3]);
      }
    }

    // ---- this warp is done with its slab: refill it with tile i + STAGES ----
    __syncwarp();
    if (lane == 0 && i + PA_STAGES < n_tiles) {
      fence_proxy_async_smem();
      issue(i + PA_STAGES);
    }
  }

  if (g < p.rows) {
    const float inv = 1.0f / l_run;
    __nv_bfloat16* dst = p.y + (seq * p.rows + g) * p.d + warp * 64 + 2 * q;
#pragma unroll
    for (int c = 0; c < 8; ++c)
      *reinterpret_cast<uint32_t*>(dst + 8 * c) = pack_bf16(yacc[c][0] * inv, yacc[c][1] * inv);
  }
}

template <int NW>
static int launch_pooled(cudaStream_t st, const CUtensorMap& tm, const PoolAttnParams& p) {
  const int smem = PA_STAGES * NW * PA_SLAB + NW * 8 * PA_RED_STRIDE * static_cast<int>(sizeof(float)) +
                   PA_STAGES * NW * 8 + 1024;
  auto kern = pooled_attention_kernel<NW>;
  LA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<static_cast<unsigned>(p.n_seq), NW * 32, smem, st>>>(tm, p);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

}  // namespace la

extern "C" int la_attention_pooled_bf16(void* stream, const void* x, long long ldx, const void* u, const float* e,
                                        long long lde, float scale, void* y, long long n_seq, int tokens, int rows,
                                        int d) {
  using namespace la;
  LA_CHECK_ARG(x && u && y, "la_attention_pooled_bf16: null pointer");
  LA_CHECK_ARG(n_seq > 0 && n_seq < (1ll << 31) && tokens > 0, "la_attention_pooled_bf16: empty problem");
  LA_CHECK_ARG(rows >= 1 && rows <= 8, "la_attention_pooled_bf16: 1..8 query rows per sequence (got %d)", rows);
  LA_CHECK_ARG(d % 64 == 0 && d >= 64 && d <= 512 && (d / 64 == 1 || d / 64 == 2 || d / 64 == 4 || d / 64 == 8),
               "la_attention_pooled_bf16: d must be 64, 128, 256 or 512 (got %d)", d);
  LA_CHECK_ARG(ldx >= d && ldx % 8 == 0, "la_attention_pooled_bf16: bad row stride");
  LA_CHECK_ARG(e == nullptr || lde >= tokens, "la_attention_pooled_bf16: lde must be >= tokens");
  LA_CHECK_ARG(n_seq * tokens < (1ll << 31), "la_attention_pooled_bf16: too many rows for TMA coordinates");
  LA_CHECK_ARG(scale > 0.f, "la_attention_pooled_bf16: the softmax scale must be positive");
  CUtensorMap tm;
  int rc = make_tensor_map_2d(&tm, x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<uint64_t>(d),
                              static_cast<uint64_t>(n_seq * tokens), static_cast<uint64_t>(ldx) * 2, 64, PA_TILE,
                              Swizzle::B128);
  if (rc) return rc;
  PoolAttnParams p;
  p.u = static_cast<const __nv_bfloat16*>(u);
  p.e = e;
  p.lde = lde;
  p.y = static_cast<__nv_bfloat16*>(y);
  p.n_seq = n_seq;
  p.tokens = tokens;
  p.rows = rows;
  p.d = d;
  p.scale_log2 = scale * 1.4426950408889634f;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (d / 64) {
    case 1: return launch_pooled<1>(st, tm, p);
    case 2: return launch_pooled<2>(st, tm, p);
    case 4: return launch_pooled<4>(st, tm, p);
    default: return launch_pooled<8>(st, tm, p);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// head-block expansion / gathering around the pooled attention: the per-head vectors u_h = W_k[h]^T q_h and
// o_h = W_v[h] y_h are formed by ordinary GEMMs over rows indexed by (sequence, head):
//   mode 0 (expand): out[(s, h), c] = in[s, c] if c / head_dim == h else 0       [n_seq, H*dh] -> [n_seq*H, H*dh]
//   mode 1 (gather): out[s, h*dh + j] = in[(s, h), h*dh + j]                       [n_seq*H, H*dh] -> [n_seq, H*dh]
// ------------------------------------------------------------------------------------------------------------------
namespace la {
__global__ void __launch_bounds__(256)
head_rows_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n_seq, int heads,
                 int head_dim, int mode) {
  const int w = heads * head_dim;
  const long long total = mode == 0 ? n_seq * heads * w : n_seq * w;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    if (mode == 0) {
      const int c = static_cast<int>(i % w);
      const long long sh = i / w;
      const int h = static_cast<int>(sh % heads);
      out[i] = (c / head_dim == h) ? in[(sh / heads) * w + c] : __float2bfloat16_rn(0.f);
    } else {
      const int c = static_cast<int>(i % w);
      const long long s = i / w;
      out[i] = in[(s * heads + c / head_dim) * w + c];
    }
  }
}
}  // namespace la

extern "C" int la_head_rows_bf16(void* stream, const void* in, void* out, long long n_seq, int heads, int head_dim,
                                 int mode) {
  using namespace la;
  LA_CHECK_ARG(in && out && n_seq > 0 && heads > 0 && head_dim > 0 && (mode == 0 || mode == 1),
               "la_head_rows_bf16: bad arguments");
  const long long total = (mode == 0 ? n_seq * heads : n_seq) * static_cast<long long>(heads) * head_dim;
  long long grid = (total + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (grid > cap) grid = cap;
  head_rows_kernel<<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), n_seq, heads, head_dim, mode);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}
