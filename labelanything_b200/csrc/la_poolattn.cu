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
// mma.sync m16n8k16 (16 tokens x 8 query rows per instruction -- an 8-row problem has no use for a 128-row tcgen05
// tile), the partials are summed through shared memory, every warp runs the same online softmax and accumulates
// y[:, its 64 channels] with a second set of mma.sync (x^T through ldmatrix.trans, P through a 1 KB per-warp buffer).
// HBM-bound: x is read exactly once (algorithmic bytes = 2 * tokens * d per sequence).
#include "la_common.cuh"

namespace la {

constexpr int PA_TILE = 64;                 // tokens per tile
constexpr int PA_STAGES = 3;
constexpr int PA_SLAB = PA_TILE * 128;      // one [64 tokens x 64 channels] bf16 box
constexpr int PA_P_STRIDE = 72;             // bf16 per head row of a warp's P buffer (64 tokens + pad)

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// D[16 x 8] += A[16 x 16] B[16 x 8]  (bf16 operands, fp32 accumulate)
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
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

// The legacy tensor path (HMMA) runs at an eighth of the tcgen05 rate on sm_100 (measured: 32 cycles per m16n8k16 per
// scheduler), so the MMA shapes carry no padding: tokens are the M dimension of the score product (16 tokens x 8 query
// rows per instruction), channels the M dimension of the value product (16 channels x 8 query rows, K = 16 tokens).
template <int NW>
__global__ void __launch_bounds__(NW * 32)
pooled_attention_kernel(const __grid_constant__ CUtensorMap tm_x, const PoolAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // [stage][warp] slabs | partial scores [warp][64 tokens][8 rows] fp32 | totals [64][8] | P [warp][8 rows][72] bf16 | mbarriers
  float* red = reinterpret_cast<float*>(smem + PA_STAGES * NW * PA_SLAB);
  float* tot = red + NW * PA_TILE * 8;                                   // summed scores [64 tokens][8 rows]
  __nv_bfloat16* pbuf_all = reinterpret_cast<__nv_bfloat16*>(tot + PA_TILE * 8);
  uint64_t* bars = reinterpret_cast<uint64_t*>(pbuf_all + NW * 8 * PA_P_STRIDE);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const long long seq = blockIdx.x;
  const int n_tiles = (p.tokens + PA_TILE - 1) / PA_TILE;
  const long long row0 = seq * p.tokens;
  __nv_bfloat16* pbuf = pbuf_all + warp * 8 * PA_P_STRIDE;

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

  // query-side operand as B fragments: b0 = u[row g][64 warp + 16 kk + 2q, +1], b1 = ... + 8 (rows >= p.rows: zero)
  uint32_t ub0[4], ub1[4];
  {
    const bool live = g < p.rows;
    const __nv_bfloat16* ur = p.u + (seq * p.rows + (live ? g : 0)) * p.d + warp * 64 + 2 * q;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      ub0[kk] = live ? __ldg(reinterpret_cast<const uint32_t*>(ur + 16 * kk)) : 0u;
      ub1[kk] = live ? __ldg(reinterpret_cast<const uint32_t*>(ur + 16 * kk + 8)) : 0u;
    }
  }
  // this lane's score columns are query rows 2q and 2q + 1; its score rows tokens 16 mt + g and 16 mt + 8 + g
  const int h0 = 2 * q, h1 = 2 * q + 1;
  const float* e0row = (p.e != nullptr && h0 < p.rows) ? p.e + (seq * p.rows + h0) * p.lde : nullptr;
  const float* e1row = (p.e != nullptr && h1 < p.rows) ? p.e + (seq * p.rows + h1) * p.lde : nullptr;
  auto load_e = [&](int tile, float (&ev)[16]) {
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int t = tile * PA_TILE + 16 * mt + 8 * hh + g;
        const bool in = t < p.tokens;
        ev[4 * mt + 2 * hh] = (in && e0row != nullptr) ? __ldg(e0row + t) : 0.f;
        ev[4 * mt + 2 * hh + 1] = (in && e1row != nullptr) ? __ldg(e1row + t) : 0.f;
      }
    }
  };
  float e_next[16];
  load_e(0, e_next);

  float yacc[4][4];
#pragma unroll
  for (int c = 0; c < 4; ++c) yacc[c][0] = yacc[c][1] = yacc[c][2] = yacc[c][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  // partial scores of tile `tile` over this warp's 64 channels: S[64 tokens x 8 rows] = X_w[64 x 64] U_w^T
  auto score_mmas = [&](int tile, float (&sp)[4][4]) {
    const int st = tile % PA_STAGES;
    mbar_wait(&bars[st * NW + warp], (tile / PA_STAGES) & 1);
    const uint32_t slab = smem_u32(smem + (st * NW + warp) * PA_SLAB);
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      sp[mt][0] = sp[mt][1] = sp[mt][2] = sp[mt][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        // A = X: matrices (tokens 0-7, k lo) (tokens 8-15, k lo) (tokens 0-7, k hi) (tokens 8-15, k hi)
        const int mi = lane >> 3, rr = lane & 7;
        const int tok = 16 * mt + 8 * (mi & 1) + rr;
        const int chunk = 2 * kk + (mi >> 1);
        uint32_t a[4];
        ldmatrix_x4(a, slab + tok * 128 + ((chunk ^ (tok & 7)) << 4));
        mma_16816(sp[mt], a, ub0[kk], ub1[kk]);
      }
    }
  };
  auto store_partials = [&](const float (&sp)[4][4]) {
    float* mine = red + (warp * PA_TILE + g) * 8 + 2 * q;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      *reinterpret_cast<float2*>(mine + (16 * mt) * 8) = make_float2(sp[mt][0], sp[mt][1]);
      *reinterpret_cast<float2*>(mine + (16 * mt + 8) * 8) = make_float2(sp[mt][2], sp[mt][3]);
    }
  };

  // Software pipeline: the score MMAs of tile i + 1 are issued BEFORE the softmax of tile i, so the tensor pipe works
  // through them (and then through the value MMAs of tile i) while the ALUs do the reduction and the exponentials.
  float s_next[4][4];
  score_mmas(0, s_next);
  store_partials(s_next);
  __syncthreads();

  for (int i = 0; i < n_tiles; ++i) {
    float ev[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) ev[k] = e_next[k];
    if (i + 1 < n_tiles) load_e(i + 1, e_next);   // consumed a tile later: the load latency hides under this tile

    // ---- (a) every warp sums the partials of ITS 64 / NW tokens (all 8 rows) and publishes the totals ----
    {
      constexpr int TPW = PA_TILE / NW;            // tokens per warp
#pragma unroll
      for (int unit = lane; unit < TPW * 4; unit += 32) {
        const int tok = warp * TPW + (unit >> 2), rp = unit & 3;
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < NW; ++w2) {
          const float2 v = *reinterpret_cast<const float2*>(red + (w2 * PA_TILE + tok) * 8 + 2 * rp);
          a0 += v.x;
          a1 += v.y;
        }
        *reinterpret_cast<float2*>(tot + tok * 8 + 2 * rp) = make_float2(a0, a1);
      }
    }
    __syncthreads();   // totals complete; `red` may be rewritten

    // ---- (b) score MMAs of the next tile (results are only needed at the end of this iteration) ----
    if (i + 1 < n_tiles) score_mmas(i + 1, s_next);

    // ---- (c) totals + positional scores, scale, tail mask, online softmax per query row ----
    float s[4][4];
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int tok = 16 * mt + 8 * hh + g;
        const float2 v = *reinterpret_cast<const float2*>(tot + tok * 8 + 2 * q);
        const bool in = i * PA_TILE + tok < p.tokens;
        const float v0 = in ? (v.x + ev[4 * mt + 2 * hh]) * p.scale_log2 : -INFINITY;
        const float v1 = in ? (v.y + ev[4 * mt + 2 * hh + 1]) * p.scale_log2 : -INFINITY;
        s[mt][2 * hh] = v0;
        s[mt][2 * hh + 1] = v1;
        mx0 = fmaxf(mx0, v0);
        mx1 = fmaxf(mx1, v1);
      }
    }
#pragma unroll
    for (int o = 4; o <= 16; o <<= 1) {
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, o));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, o));
    }
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);   // finite: every tile holds at least one real token
    const float al0 = pa_ex2(m0 - mn0), al1 = pa_ex2(m1 - mn1);
    float ls0 = 0.f, ls1 = 0.f;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const float p0 = pa_ex2(s[mt][2 * hh] - mn0), p1 = pa_ex2(s[mt][2 * hh + 1] - mn1);
        ls0 += p0;
        ls1 += p1;
        const int tok = 16 * mt + 8 * hh + g;
        pbuf[h0 * PA_P_STRIDE + tok] = __float2bfloat16_rn(p0);
        pbuf[h1 * PA_P_STRIDE + tok] = __float2bfloat16_rn(p1);
      }
    }
#pragma unroll
    for (int o = 4; o <= 16; o <<= 1) {
      ls0 += __shfl_xor_sync(0xffffffffu, ls0, o);
      ls1 += __shfl_xor_sync(0xffffffffu, ls1, o);
    }
    l0 = l0 * al0 + ls0;
    l1 = l1 * al1 + ls1;
    m0 = mn0;
    m1 = mn1;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      yacc[c][0] *= al0;
      yacc[c][1] *= al1;
      yacc[c][2] *= al0;
      yacc[c][3] *= al1;
    }
    __syncwarp();   // P of this warp is complete in its buffer

    // ---- (d) Y^T[64 channels x 8 rows] += X_w^T[64 channels x 64 tokens] P[64 tokens x 8 rows] ----
    const uint32_t slab = smem_u32(smem + ((i % PA_STAGES) * NW + warp) * PA_SLAB);
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
      // B = P: b0 = P[tokens 16kt + 2q, +1][row g], b1 = tokens + 8
      const uint32_t pb0 = *reinterpret_cast<const uint32_t*>(pbuf + g * PA_P_STRIDE + 16 * kt + 2 * q);
      const uint32_t pb1 = *reinterpret_cast<const uint32_t*>(pbuf + g * PA_P_STRIDE + 16 * kt + 8 + 2 * q);
#pragma unroll
      for (int ct = 0; ct < 4; ++ct) {
        // A = X^T through ldmatrix.trans: matrices (tokens lo, channels 0-7) (tokens lo, channels 8-15)
        //                                          (tokens hi, channels 0-7) (tokens hi, channels 8-15)
        const int mi = lane >> 3, rr = lane & 7;
        const int tok = 16 * kt + 8 * (mi >> 1) + rr;
        const int chunk = 2 * ct + (mi & 1);
        uint32_t a[4];
        ldmatrix_x4_trans(a, slab + tok * 128 + ((chunk ^ (tok & 7)) << 4));
        mma_16816(yacc[ct], a, pb0, pb1);
      }
    }

    // ---- (e) this warp is done with the slab of tile i (and its P buffer): refill it with tile i + STAGES ----
    __syncwarp();
    if (lane == 0 && i + PA_STAGES < n_tiles) {
      fence_proxy_async_smem();
      issue(i + PA_STAGES);
    }
    // ---- (f) publish the next tile's partial scores ----
    if (i + 1 < n_tiles) store_partials(s_next);
    __syncthreads();   // partials of tile i + 1 complete; every warp has read the totals of tile i
  }

  // y[seq, row, 64 warp + 16 ct + g (+8)] for rows 2q, 2q + 1
  const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
  __nv_bfloat16* y0 = p.y + (seq * p.rows + h0) * p.d + warp * 64 + g;
  __nv_bfloat16* y1 = p.y + (seq * p.rows + h1) * p.d + warp * 64 + g;
#pragma unroll
  for (int ct = 0; ct < 4; ++ct) {
    if (h0 < p.rows) {
      y0[16 * ct] = __float2bfloat16_rn(yacc[ct][0] * inv0);
      y0[16 * ct + 8] = __float2bfloat16_rn(yacc[ct][2] * inv0);
    }
    if (h1 < p.rows) {
      y1[16 * ct] = __float2bfloat16_rn(yacc[ct][1] * inv1);
      y1[16 * ct + 8] = __float2bfloat16_rn(yacc[ct][3] * inv1);
    }
  }
}

template <int NW>
static int launch_pooled(cudaStream_t st, const CUtensorMap& tm, const PoolAttnParams& p) {
  const int smem = PA_STAGES * NW * PA_SLAB + (NW + 1) * PA_TILE * 8 * static_cast<int>(sizeof(float)) +
                   NW * 8 * PA_P_STRIDE * 2 + PA_STAGES * NW * 8 + 1024;
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
