// labelanything_b200 — HBM-bound row kernels of the ViT encoder and neck (sm_100a).
//
//   la_add_layernorm : residual add (fp32 stream + bf16 branch output) fused with LayerNorm and with the
//                      window-partition / zero-pad / CLS-drop row remapping, one warp per token row.
//                      reference: image_encoder.py:181-197 (x = shortcut + attn; x + mlp(norm2(x))),
//                      258-279 (window_partition incl. F.pad), common.py:42-54 (LayerNorm2d == per-token LN
//                      in token-major layout), HF modeling_vit.py:325-346,416.
//   la_embed_tokens  : x[img, tok] = (cls | patch-GEMM row) + pos_embed[tok]   image_encoder.py:112-114,
//                      HF modeling_vit.py:109-125.
//   la_im2col_patch  : NCHW fp32 image -> [tokens, 3*P*P] bf16 rows for the patch-embed GEMM
//                      (stride == kernel conv, image_encoder.py:402-410).
//   la_im2col_3x3    : token-major bf16 feature map -> [tokens, 9*C] rows (zero padded) for 3x3 convs
//                      (build_lam.py:162-168, mask_decoder.py:241-247).
// All are pure streaming kernels: 16-byte vectorised, coalesced along the channel dimension, grid sized in
// multiples of the SM count; the roofline that bounds them is HBM bandwidth.
#include "la_common.cuh"
#ifndef LA_LN_FULL_MAXV
#define LA_LN_FULL_MAXV 6       // widest row (in 128-channel slots) that takes the lean staged row loop
#endif
#ifndef LA_LN_UNSTAGED
#define LA_LN_UNSTAGED 0        // experiment builds: 1 = register-load kernels only (tools/diag_ln.py)
#endif
#include <cstdlib>

namespace la {

constexpr int ROW_MAXV = 10;  // float4 per lane -> rows up to 1280 channels

struct AddLnParams {
  const float* x_in;            // [x_rows, d] fp32 or nullptr
  long long x_mod;              // >0: x_in row = src_row % x_mod (broadcast table)
  const __nv_bfloat16* delta;   // [rows, d] or nullptr
  const __nv_bfloat16* delta2;  // [rows, d] or nullptr (second bf16 addend)
  const float* seq_add;         // [rows / seq_rows, d] or nullptr: per-sequence vector added to every row
  long long seq_rows;
  float* x_out;                 // [rows, d] or nullptr
  const float* gamma;           // nullptr -> no normalisation (plain cast)
  const float* beta;
  float eps;
  void* y_out;                  // nullptr -> skip
  int y_f32;                    // 0 bf16, 1 fp32
  float* y2_out;                // optional second copy of y in fp32 (same row mapping)
  const float* pe;              // optional [pe_mod, d] table: ype = y + pe[dst_row % pe_mod]
  long long pe_mod;
  __nv_bfloat16* ype_out;       // optional bf16 (y + pe)
  int act;                      // LA_ACT_* applied to y after the affine
  float* pool_out;              // pooling variant: [n_seq, pool_slices, d] partial sums of y
  int pool_rows;                // rows per sequence (pooling variant)
  int pool_slices;
  long long rows;               // source rows (map 0, 2) ; output rows (map 1)
  int d;
  int map_mode;                 // 0 identity, 1 window partition with zero pad, 2 drop first token per sequence,
                                // 3 pixel shuffle: src row = (img, y, x, ky, kx) -> dst row (img, 2y+ky, 2x+kx)
  int seq_len;                  // map 2
  int win, nwin, hw;            // map 1
  FastDiv fd_seq;               // division by seq_rows (staged kernels: rows < 2^31)
  FastDiv fd_w2, fd_per_img, fd_nwin, fd_win;   // map 1: divisions by win^2, nwin^2, nwin, win (rows < 2^31)
};

// NVT: float4 slots per lane (compile time, >= ceil(d / 128)); POOL: the mean-pooling variant (keeps per-lane column sums).
// Small NVT keeps the register count low enough for 4+ resident CTAs per SM, which is what hides the HBM latency.
template <int NVT, bool POOL>
__global__ void __launch_bounds__(256, NVT <= 4 ? (POOL ? 3 : 4) : ((NVT <= 8 && !POOL) ? 3 : 2))
add_layernorm_kernel(const AddLnParams p) {
  const int lane = threadIdx.x & 31;
  const int warp_in_cta = threadIdx.x >> 5;
  const int nv = p.d >> 7;            // full float4 groups per lane
  const int tail = (p.d & 127) >> 2;  // leftover float4s (< 32)

  long long row_begin, row_end, row_step;
  if (p.pool_out) {
    // pooling variant: CTA = (sequence, slice); its 8 warps stride over the slice's rows
    const long long seq = blockIdx.x / p.pool_slices;
    const int sl = blockIdx.x % p.pool_slices;
    const int per = (p.pool_rows + p.pool_slices - 1) / p.pool_slices;
    row_begin = seq * p.pool_rows + static_cast<long long>(sl) * per + warp_in_cta;
    long long e = seq * p.pool_rows + static_cast<long long>(sl + 1) * per;
    const long long seq_end = (seq + 1) * p.pool_rows;
    row_end = e < seq_end ? e : seq_end;
    row_step = blockDim.x >> 5;
  } else {
    row_begin = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    row_end = p.rows;
    row_step = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  }
  const bool any_y = p.y_out || p.y2_out || p.ype_out || p.pool_out;
  float4 acc[POOL ? NVT : 1];
#pragma unroll
  for (int i = 0; i < (POOL ? NVT : 1); ++i) acc[i] = make_float4(0, 0, 0, 0);

  for (long long row = row_begin; row < row_end; row += row_step) {
    long long src = row, dst = row;
    bool pad = false, write_y = any_y;
    if (p.map_mode == 1) {
      const int w2 = p.win * p.win;
      const long long widx = row / w2;
      const int tin = static_cast<int>(row % w2);
      const int per_img = p.nwin * p.nwin;
      const long long img = widx / per_img;
      const int wi = static_cast<int>(widx % per_img);
      const int y = (wi / p.nwin) * p.win + tin / p.win;
      const int x = (wi % p.nwin) * p.win + tin % p.win;
      pad = (y >= p.hw) || (x >= p.hw);
      src = (img * p.hw + y) * p.hw + x;
    } else if (p.map_mode == 2) {
      const long long s = row / p.seq_len;
      const int tin = static_cast<int>(row % p.seq_len);
      write_y = write_y && tin > 0;
      dst = row - s - 1;
    } else if (p.map_mode == 3) {
      // hw = input grid side (square), rows enumerate (img, y, x, ky, kx)
      const int g4 = static_cast<int>(row & 3);
      const long long pix = row >> 2;
      const int x = static_cast<int>(pix % p.hw);
      const long long r2 = pix / p.hw;
      const int y = static_cast<int>(r2 % p.hw);
      const long long img = r2 / p.hw;
      dst = (img * (2 * p.hw) + 2 * y + (g4 >> 1)) * (2 * p.hw) + 2 * x + (g4 & 1);
    }
    const size_t ebytes = p.y_f32 ? 4 : 2;
    uint8_t* yrow = static_cast<uint8_t*>(p.y_out) + static_cast<size_t>(dst) * p.d * ebytes;
    if (pad) {
      // zero row (padding token of a window): F.pad after norm1, image_encoder.py:271-275
      if (p.y_out) {
        if (p.y_f32) {
          for (int i = lane; i < p.d / 4; i += 32) reinterpret_cast<float4*>(yrow)[i] = make_float4(0, 0, 0, 0);
        } else {
          for (int i = lane; i < p.d / 8; i += 32) reinterpret_cast<uint4*>(yrow)[i] = make_uint4(0, 0, 0, 0);
        }
      }
      continue;
    }
    float4 v[NVT];
    const long long xrow = p.x_mod > 0 ? src % p.x_mod : src;
    const float4* xin = p.x_in ? reinterpret_cast<const float4*>(p.x_in + xrow * p.d) : nullptr;
    const uint2* din = p.delta ? reinterpret_cast<const uint2*>(p.delta + src * p.d) : nullptr;
    const uint2* din2 = p.delta2 ? reinterpret_cast<const uint2*>(p.delta2 + src * p.d) : nullptr;
    const float4* sadd = p.seq_add ? reinterpret_cast<const float4*>(p.seq_add + (src / p.seq_rows) * p.d) : nullptr;
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NVT; ++i) {
      const bool on = (i < nv) || (i == nv && lane < tail);
      float4 a = make_float4(0, 0, 0, 0);
      if (on) {
        const int idx = i * 32 + lane;
        if (xin) a = xin[idx];
        if (din) {
          const uint2 dv = din[idx];
          const __nv_bfloat162 d01 = *reinterpret_cast<const __nv_bfloat162*>(&dv.x);
          const __nv_bfloat162 d23 = *reinterpret_cast<const __nv_bfloat162*>(&dv.y);
          a.x += __low2float(d01);
          a.y += __high2float(d01);
          a.z += __low2float(d23);
          a.w += __high2float(d23);
        }
        if (din2) {
          const uint2 dv = din2[idx];
          const __nv_bfloat162 d01 = *reinterpret_cast<const __nv_bfloat162*>(&dv.x);
          const __nv_bfloat162 d23 = *reinterpret_cast<const __nv_bfloat162*>(&dv.y);
          a.x += __low2float(d01);
          a.y += __high2float(d01);
          a.z += __low2float(d23);
          a.w += __high2float(d23);
        }
        if (sadd) {
          const float4 e = __ldg(sadd + idx);
          a.x += e.x;
          a.y += e.y;
          a.z += e.z;
          a.w += e.w;
        }
        if (p.x_out) reinterpret_cast<float4*>(p.x_out + src * p.d)[idx] = a;
      }
      v[i] = a;
      sum += a.x + a.y + a.z + a.w;
    }
    if (!write_y) continue;
    float mean = 0.f, rstd = 1.f;
    if (p.gamma) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      mean = sum / p.d;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < NVT; ++i) {
        const bool on = (i < nv) || (i == nv && lane < tail);
        if (on) {
          const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
          sq += dx * dx + dy * dy + dz * dz + dw * dw;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      rstd = rsqrtf(sq / p.d + p.eps);
    }
#pragma unroll
    for (int i = 0; i < NVT; ++i) {
      const bool on = (i < nv) || (i == nv && lane < tail);
      if (on) {
        const int idx = i * 32 + lane;
        float4 o = v[i];
        if (p.gamma) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma) + idx);
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.beta) + idx);
          o.x = (o.x - mean) * rstd * g.x + b.x;
          o.y = (o.y - mean) * rstd * g.y + b.y;
          o.z = (o.z - mean) * rstd * g.z + b.z;
          o.w = (o.w - mean) * rstd * g.w + b.w;
        }
        if (p.act == LA_ACT_GELU) {
          o.x = gelu_erf(o.x);
          o.y = gelu_erf(o.y);
          o.z = gelu_erf(o.z);
          o.w = gelu_erf(o.w);
        } else if (p.act == LA_ACT_RELU) {
          o.x = fmaxf(o.x, 0.f);
          o.y = fmaxf(o.y, 0.f);
          o.z = fmaxf(o.z, 0.f);
          o.w = fmaxf(o.w, 0.f);
        }
        if (p.y_out) {
          if (p.y_f32) {
            reinterpret_cast<float4*>(yrow)[idx] = o;
          } else {
            uint2 pk;
            pk.x = pack_bf16(o.x, o.y);
            pk.y = pack_bf16(o.z, o.w);
            reinterpret_cast<uint2*>(yrow)[idx] = pk;
          }
        }
        if (p.y2_out) reinterpret_cast<float4*>(p.y2_out + dst * p.d)[idx] = o;
        if (p.ype_out) {
          const float4 e = __ldg(reinterpret_cast<const float4*>(p.pe + (dst % p.pe_mod) * p.d) + idx);
          uint2 pk;
          pk.x = pack_bf16(o.x + e.x, o.y + e.y);
          pk.y = pack_bf16(o.z + e.z, o.w + e.w);
          reinterpret_cast<uint2*>(p.ype_out + dst * p.d)[idx] = pk;
        }
        if constexpr (POOL) {
          acc[i].x += o.x;
          acc[i].y += o.y;
          acc[i].z += o.z;
          acc[i].w += o.w;
        }
      }
    }
  }
  if constexpr (POOL) {
    // deterministic reduction: lanes own disjoint channels; the CTA's warps are summed in a fixed order
    // through shared memory; one partial row per (sequence, slice).
    extern __shared__ float4 red[];  // [warps][d/4]
    const int dv = p.d >> 2;
#pragma unroll
    for (int i = 0; i < NVT; ++i) {
      const bool on = (i < nv) || (i == nv && lane < tail);
      if (on) red[warp_in_cta * dv + i * 32 + lane] = acc[i];
    }
    __syncthreads();
    const int nw = blockDim.x >> 5;
    float4* dstp = reinterpret_cast<float4*>(p.pool_out + static_cast<long long>(blockIdx.x) * p.d);
    for (int c = threadIdx.x; c < dv; c += blockDim.x) {
      float4 t = red[c];
      for (int w = 1; w < nw; ++w) {
        const float4 u = red[w * dv + c];
        t.x += u.x;
        t.y += u.y;
        t.z += u.z;
        t.w += u.w;
      }
      dstp[c] = t;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Staged variant for the ViT block LayerNorms (the hot case: x += delta; y = LN(x) as bf16; identity or
// window-partition row mapping).  Same arithmetic as add_layernorm_kernel, different data movement: every warp
// owns a 3-deep ring of row buffers in shared memory that is filled by bulk async copies (cp.async.bulk, completion
// on an mbarrier), so each warp keeps three rows of reads in flight while it reduces / normalises / stores the
// current one -- with one-row-at-a-time register loads the kernel sat at ~4 TB/s because the reads of a warp stop
// during its compute-and-store phase.  2 CTAs x 8 warps x 3 rows x 4.5 KB = 216 KB of reads in flight per SM.
// ---------------------------------------------------------------------------------------------------------
constexpr int LN_MAX_STAGES = 8;

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Every warp owns a ring of `stages` slots; a slot holds a CHUNK of `rpc` consecutive rows (fp32 part, then the bf16
// part) fetched by ONE bulk copy per operand -- the bulk-copy engine has a per-operation cost of the order of 100
// cycles, so 1 KB rows (the prompt encoder's bf16 image tokens) must be moved several at a time to reach HBM speed.
// rpc = 1 when rows are remapped (window partition).  POOL: the mean-pool variant -- CTA = (sequence, slice), nothing
// but per-CTA column sums of y is written (prompt_encoder.py:733-735).
// FULL: d == NVT * 128 (every lane slot is live): the lean row loop -- no slot predicates, the per-sequence vector
// looked up with a multiply-high division (or once per CTA when pooling), pointers advanced instead of recomputed.
// The general loop issued 348 warp instructions per 512-channel row in the mean-pool case, 68 % issue-active at
// 2.3 TB/s (profiles/r01_ncu_meanpool_v4.txt): instruction-bound on an HBM-bound operation.
template <int NVT, bool POOL, bool HAS_X, bool HAS_D, bool HAS_D2 = false, bool FULL = false>
__global__ void __launch_bounds__(256, 2) add_layernorm_staged_kernel(const AddLnParams p, const int stages, const int rpc) {
  extern __shared__ __align__(128) uint8_t ln_smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nv = p.d >> 7;
  const int tail = (p.d & 127) >> 2;
  const uint32_t xbytes = HAS_X ? static_cast<uint32_t>(p.d) * 4 : 0;
  const uint32_t dbytes = HAS_D ? static_cast<uint32_t>(p.d) * 2 : 0;
  const uint32_t d2bytes = HAS_D2 ? static_cast<uint32_t>(p.d) * 2 : 0;   // second bf16 stream (delta2)
  const uint32_t slot_bytes = (static_cast<uint32_t>(rpc) * (xbytes + dbytes + d2bytes) + 127) & ~127u;
  uint8_t* ring = ln_smem + static_cast<size_t>(warp) * stages * slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ln_smem + static_cast<size_t>(8) * stages * slot_bytes) + warp * LN_MAX_STAGES;
  if (lane == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  __syncwarp();

  // this warp's chunks: rows [base + c * rpc, min(base + (c + 1) * rpc, row_end)) for c = c_first, c_first + c_step, ...
  long long base, row_end, c_first, c_step;
  if constexpr (POOL) {
    const long long seq = blockIdx.x / p.pool_slices;
    const int sl = blockIdx.x % p.pool_slices;
    const int per = (p.pool_rows + p.pool_slices - 1) / p.pool_slices;
    base = seq * p.pool_rows + static_cast<long long>(sl) * per;
    const long long e = base + per;
    const long long seq_end = (seq + 1) * p.pool_rows;
    row_end = e < seq_end ? e : seq_end;
    c_first = warp;
    c_step = 8;
  } else {
    base = 0;
    row_end = p.rows;
    c_first = static_cast<long long>(blockIdx.x) * 8 + warp;
    c_step = static_cast<long long>(gridDim.x) * 8;
  }
  const long long n_chunks = row_end > base ? (row_end - base + rpc - 1) / rpc : 0;

  // row mapping (map_mode 0: identity; 1: window partition with zero padding, image_encoder.py:258-279)
  auto map_row = [&](long long row, long long& src, bool& pad) {
    src = row;
    pad = false;
    if (p.map_mode == 1) {
      // 32-bit index arithmetic (the launcher guarantees rows < 2^31): these divisions sit on the refill path
      // (multiply-high divisions: the plain ones -- four MUFU.RCP sequences per chunk on the refill path and again
      //  in the consumer -- held this mode at 0.60 of the copy peak against 0.82 for the identity mapping)
      const uint32_t r32 = static_cast<uint32_t>(row);
      const uint32_t w2 = p.win * p.win;
      const uint32_t widx = fast_div(r32, p.fd_w2);
      const uint32_t tin = r32 - widx * w2;
      const uint32_t per_img = p.nwin * p.nwin;
      const uint32_t img = fast_div(widx, p.fd_per_img);
      const uint32_t wi = widx - img * per_img;
      const uint32_t wy = fast_div(wi, p.fd_nwin), ty = fast_div(tin, p.fd_win);
      const uint32_t y = wy * p.win + ty;
      const uint32_t x = (wi - wy * p.nwin) * p.win + (tin - ty * p.win);
      pad = (y >= static_cast<uint32_t>(p.hw)) || (x >= static_cast<uint32_t>(p.hw));
      src = (static_cast<long long>(img) * p.hw + y) * p.hw + x;
    }
  };
  auto issue = [&](long long chunk, int slot) {   // lane 0
    const long long row0 = base + chunk * rpc;
    long long src;
    bool pad;
    map_row(row0, src, pad);   // remapped rows: the launcher only allows chunks that share src contiguity and padding
    if (pad) return;
    const uint32_t n = static_cast<uint32_t>(row_end - row0 < rpc ? row_end - row0 : rpc);
    uint8_t* dst = ring + static_cast<size_t>(slot) * slot_bytes;
    mbar_arrive_expect_tx(&bars[slot], n * (xbytes + dbytes + d2bytes));
    if constexpr (HAS_X) bulk_load(dst, p.x_in + src * p.d, n * xbytes, &bars[slot]);
    if constexpr (HAS_D) bulk_load(dst + rpc * xbytes, p.delta + src * p.d, n * dbytes, &bars[slot]);
    if constexpr (HAS_D2) bulk_load(dst + rpc * (xbytes + dbytes), p.delta2 + src * p.d, n * d2bytes, &bars[slot]);
  };

  if (lane == 0) {
    for (int s = 0; s < stages; ++s) {
      const long long c = c_first + s * c_step;
      if (c < n_chunks) issue(c, s);
    }
  }
  uint32_t phases = 0;   // bit s: parity of the next completion of slot s
  const float inv_d = 1.0f / static_cast<float>(p.d);
  const size_t ebytes = p.y_f32 ? 4 : 2;
  // gamma / beta (and the current sequence's seq_add vector) live in registers: re-reading them per row through L1
  // costs more load bandwidth than the row itself when rows are 1 KB of bf16
  float4 gmv[NVT], btv[NVT], sav[NVT];
#pragma unroll
  for (int i = 0; i < NVT; ++i) {
    const bool on = (i < nv) || (i == nv && lane < tail);
    gmv[i] = on ? __ldg(reinterpret_cast<const float4*>(p.gamma) + i * 32 + lane) : make_float4(0, 0, 0, 0);
    btv[i] = on ? __ldg(reinterpret_cast<const float4*>(p.beta) + i * 32 + lane) : make_float4(0, 0, 0, 0);
    sav[i] = make_float4(0, 0, 0, 0);
  }
  long long sa_seq = -1;
  if constexpr (FULL && POOL) {
    if (p.seq_add) {   // one sequence per CTA
      sa_seq = blockIdx.x / p.pool_slices;
      const float4* sadd = reinterpret_cast<const float4*>(p.seq_add + sa_seq * p.d);
#pragma unroll
      for (int i = 0; i < NVT; ++i) sav[i] = __ldg(sadd + i * 32 + lane);
    }
  }
  float4 acc[POOL ? NVT : 1];
#pragma unroll
  for (int i = 0; i < (POOL ? NVT : 1); ++i) acc[i] = make_float4(0, 0, 0, 0);
  [[maybe_unused]] float pool_c = 0.f, pool_n = 0.f;   // POOL: sum of rstd_r * mean_r, rows seen by this warp
  int slot = 0;
  for (long long chunk = c_first; chunk < n_chunks; chunk += c_step) {
    const long long row0 = base + chunk * rpc;
    const int n = static_cast<int>(row_end - row0 < rpc ? row_end - row0 : rpc);
    long long src0;
    bool pad;
    map_row(row0, src0, pad);
    if (pad) {
      uint8_t* yrow = static_cast<uint8_t*>(p.y_out) + static_cast<size_t>(row0) * p.d * ebytes;   // n consecutive rows
      if (p.y_f32) {
        for (int i = lane; i < n * (p.d / 4); i += 32) reinterpret_cast<float4*>(yrow)[i] = make_float4(0, 0, 0, 0);
      } else {
        for (int i = lane; i < n * (p.d / 8); i += 32) reinterpret_cast<uint4*>(yrow)[i] = make_uint4(0, 0, 0, 0);
      }
    } else {
      const uint8_t* buf = ring + static_cast<size_t>(slot) * slot_bytes;
      mbar_wait(&bars[slot], (phases >> slot) & 1);
      phases ^= 1u << slot;
      if constexpr (FULL) {
        const uint8_t* xrow = buf + lane * 16;                                  // + i * 512 + rr * xbytes
        const uint8_t* drow = buf + static_cast<size_t>(rpc) * xbytes + lane * 8;   // + i * 256 + rr * dbytes
        const uint8_t* d2row = buf + static_cast<size_t>(rpc) * (xbytes + dbytes) + lane * 8;
        float* xo = p.x_out ? p.x_out + src0 * p.d + lane * 4 : nullptr;
        uint8_t* yrow = POOL ? nullptr : static_cast<uint8_t*>(p.y_out) + static_cast<size_t>(row0) * p.d * ebytes;
        uint32_t sq_cur = 0, sq_left = 0;   // sequence of the current row, rows left in it
        if (!POOL && p.seq_add) {
          sq_cur = fast_div(static_cast<uint32_t>(src0), p.fd_seq);
          const long long left = (static_cast<long long>(sq_cur) + 1) * p.seq_rows - src0;
          sq_left = left < 0xffffffffll ? static_cast<uint32_t>(left) : 0xffffffffu;
        }
#pragma unroll 1
        for (int rr = 0; rr < n; ++rr) {
          if (!POOL && p.seq_add) {
            if (static_cast<long long>(sq_cur) != sa_seq) {   // warp-uniform; once per sequence
              sa_seq = sq_cur;
              const float4* sadd = reinterpret_cast<const float4*>(p.seq_add + static_cast<long long>(sq_cur) * p.d);
#pragma unroll
              for (int i = 0; i < NVT; ++i) sav[i] = __ldg(sadd + i * 32 + lane);
            }
            if (--sq_left == 0) {
              ++sq_cur;
              sq_left = p.seq_rows < 0xffffffffll ? static_cast<uint32_t>(p.seq_rows) : 0xffffffffu;
            }
          }
          float4 v[NVT];
#pragma unroll
          for (int i = 0; i < NVT; ++i) {
            float4 a;
            if constexpr (HAS_X) a = *reinterpret_cast<const float4*>(xrow + i * 512);
            if constexpr (HAS_D) {
              const uint2 dv = *reinterpret_cast<const uint2*>(drow + i * 256);
              const float d0 = __uint_as_float(dv.x << 16), d1 = __uint_as_float(dv.x & 0xffff0000u);
              const float d2 = __uint_as_float(dv.y << 16), d3 = __uint_as_float(dv.y & 0xffff0000u);
              if constexpr (HAS_X) {
                a.x += d0, a.y += d1, a.z += d2, a.w += d3;
              } else {
                a = make_float4(d0, d1, d2, d3);
              }
            }
            if constexpr (HAS_D2) {
              const uint2 dv = *reinterpret_cast<const uint2*>(d2row + i * 256);
              a.x += __uint_as_float(dv.x << 16), a.y += __uint_as_float(dv.x & 0xffff0000u);
              a.z += __uint_as_float(dv.y << 16), a.w += __uint_as_float(dv.y & 0xffff0000u);
            }
            if (p.seq_add) a.x += sav[i].x, a.y += sav[i].y, a.z += sav[i].z, a.w += sav[i].w;
            if (xo) *reinterpret_cast<float4*>(xo + i * 128) = a;
            v[i] = a;
          }
          float sum = 0.f;
#pragma unroll
          for (int i = 0; i < NVT; ++i) sum += v[i].x + v[i].y + v[i].z + v[i].w;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
          const float mean = sum * inv_d;
          float sq = 0.f;
#pragma unroll
          for (int i = 0; i < NVT; ++i) {
            const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
            sq += dx * dx + dy * dy + dz * dz + dw * dw;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
          const float rstd = rsqrtf(sq * inv_d + p.eps);
          if constexpr (POOL) {
            pool_c = fmaf(rstd, mean, pool_c);
            pool_n += 1.0f;
#pragma unroll
            for (int i = 0; i < NVT; ++i) {
              acc[i].x = fmaf(v[i].x, rstd, acc[i].x);
              acc[i].y = fmaf(v[i].y, rstd, acc[i].y);
              acc[i].z = fmaf(v[i].z, rstd, acc[i].z);
              acc[i].w = fmaf(v[i].w, rstd, acc[i].w);
            }
          } else {
#pragma unroll
            for (int i = 0; i < NVT; ++i) {
              const float4 g = gmv[i], b = btv[i];
              float4 o;   // same arithmetic as the general loop: the two paths give identical bits
              o.x = (v[i].x - mean) * rstd * g.x + b.x;
              o.y = (v[i].y - mean) * rstd * g.y + b.y;
              o.z = (v[i].z - mean) * rstd * g.z + b.z;
              o.w = (v[i].w - mean) * rstd * g.w + b.w;
              if (p.y_f32) {
                reinterpret_cast<float4*>(yrow)[i * 32 + lane] = o;
              } else {
                reinterpret_cast<uint2*>(yrow)[i * 32 + lane] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
              }
            }
            yrow += static_cast<size_t>(p.d) * ebytes;
          }
          xrow += xbytes;
          drow += dbytes;
          d2row += d2bytes;
          if (xo) xo += p.d;
        }
      } else
      for (int rr = 0; rr < n; ++rr) {
        const long long row = row0 + rr, src = src0 + rr;
        const float4* xs = reinterpret_cast<const float4*>(buf + static_cast<size_t>(rr) * xbytes);
        const uint2* ds = reinterpret_cast<const uint2*>(buf + static_cast<size_t>(rpc) * xbytes + static_cast<size_t>(rr) * dbytes);
        [[maybe_unused]] const uint2* ds2 = reinterpret_cast<const uint2*>(buf + static_cast<size_t>(rpc) * (xbytes + dbytes) + static_cast<size_t>(rr) * d2bytes);
        if (p.seq_add) {
          const long long sq_ = src / p.seq_rows;
          if (sq_ != sa_seq) {   // warp-uniform
            sa_seq = sq_;
            const float4* sadd = reinterpret_cast<const float4*>(p.seq_add + sq_ * p.d);
#pragma unroll
            for (int i = 0; i < NVT; ++i)
              if ((i < nv) || (i == nv && lane < tail)) sav[i] = __ldg(sadd + i * 32 + lane);
          }
        }
        float4 v[NVT];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < NVT; ++i) {
          const bool on = (i < nv) || (i == nv && lane < tail);
          float4 a = make_float4(0, 0, 0, 0);
          if (on) {
            const int idx = i * 32 + lane;
            if constexpr (HAS_X) a = xs[idx];
            if constexpr (HAS_D) {
              const uint2 dv = ds[idx];
              const __nv_bfloat162 d01 = *reinterpret_cast<const __nv_bfloat162*>(&dv.x);
              const __nv_bfloat162 d23 = *reinterpret_cast<const __nv_bfloat162*>(&dv.y);
              a.x += __low2float(d01);
              a.y += __high2float(d01);
              a.z += __low2float(d23);
              a.w += __high2float(d23);
            }
            if constexpr (HAS_D2) {
              const uint2 dv = ds2[idx];
              const __nv_bfloat162 d01 = *reinterpret_cast<const __nv_bfloat162*>(&dv.x);
              const __nv_bfloat162 d23 = *reinterpret_cast<const __nv_bfloat162*>(&dv.y);
              a.x += __low2float(d01);
              a.y += __high2float(d01);
              a.z += __low2float(d23);
              a.w += __high2float(d23);
            }
            a.x += sav[i].x;
            a.y += sav[i].y;
            a.z += sav[i].z;
            a.w += sav[i].w;
            if (p.x_out) reinterpret_cast<float4*>(p.x_out + src * p.d)[idx] = a;
          }
          v[i] = a;
          sum += a.x + a.y + a.z + a.w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum * inv_d;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < NVT; ++i) {
          const bool on = (i < nv) || (i == nv && lane < tail);
          if (on) {
            const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
            sq += dx * dx + dy * dy + dz * dz + dw * dw;
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const float rstd = rsqrtf(sq * inv_d + p.eps);
        if constexpr (POOL) {
          // mean over rows of (x - mean_r) rstd_r gamma + beta = gamma (sum_r rstd_r x_r - sum_r rstd_r mean_r) / T + beta:
          // one FMA per element here, gamma / beta once per warp after the loop
          pool_c = fmaf(rstd, mean, pool_c);
          pool_n += 1.0f;
#pragma unroll
          for (int i = 0; i < NVT; ++i) {
            acc[i].x = fmaf(v[i].x, rstd, acc[i].x);
            acc[i].y = fmaf(v[i].y, rstd, acc[i].y);
            acc[i].z = fmaf(v[i].z, rstd, acc[i].z);
            acc[i].w = fmaf(v[i].w, rstd, acc[i].w);
          }
          continue;
        }
        uint8_t* yrow = static_cast<uint8_t*>(p.y_out) + static_cast<size_t>(row) * p.d * ebytes;
#pragma unroll
        for (int i = 0; i < NVT; ++i) {
          const bool on = (i < nv) || (i == nv && lane < tail);
          if (on) {
            const int idx = i * 32 + lane;
            const float4 g = gmv[i], b = btv[i];
            float4 o;
            o.x = (v[i].x - mean) * rstd * g.x + b.x;
            o.y = (v[i].y - mean) * rstd * g.y + b.y;
            o.z = (v[i].z - mean) * rstd * g.z + b.z;
            o.w = (v[i].w - mean) * rstd * g.w + b.w;
            if constexpr (POOL) {
              acc[i].x += o.x;
              acc[i].y += o.y;
              acc[i].z += o.z;
              acc[i].w += o.w;
            } else if (p.y_f32) {
              reinterpret_cast<float4*>(yrow)[idx] = o;
            } else {
              uint2 pk;
              pk.x = pack_bf16(o.x, o.y);
              pk.y = pack_bf16(o.z, o.w);
              reinterpret_cast<uint2*>(yrow)[idx] = pk;
            }
          }
        }
      }
    }
    // every lane has read its part of the slot: refill it with the chunk `stages` ahead.  Chunks that are pure padding
    // neither load nor wait, but they still advance the slot so that issue order == consume order.
    __syncwarp();
    const long long nxt = chunk + stages * c_step;
    if (lane == 0 && nxt < n_chunks) issue(nxt, slot);
    slot = slot + 1 == stages ? 0 : slot + 1;
  }
  if constexpr (POOL) {
    // deterministic reduction: lanes own disjoint channels; the CTA's warps are summed in a fixed order through
    // shared memory; one partial row per (sequence, slice).  The scratch aliases the ring: every issued chunk has been
    // consumed by now, the barrier makes sure all warps are done with theirs.
    __syncthreads();
    float4* red = reinterpret_cast<float4*>(ln_smem);
    const int dv = p.d >> 2;
#pragma unroll
    for (int i = 0; i < NVT; ++i) {
      const bool on = (i < nv) || (i == nv && lane < tail);
      if (on) {
        const float4 g = gmv[i], b = btv[i];
        float4 a = acc[i];
        a.x = fmaf(g.x, a.x - pool_c, b.x * pool_n);
        a.y = fmaf(g.y, a.y - pool_c, b.y * pool_n);
        a.z = fmaf(g.z, a.z - pool_c, b.z * pool_n);
        a.w = fmaf(g.w, a.w - pool_c, b.w * pool_n);
        red[warp * dv + i * 32 + lane] = a;
      }
    }
    __syncthreads();
    float4* dstp = reinterpret_cast<float4*>(p.pool_out + static_cast<long long>(blockIdx.x) * p.d);
    for (int c = threadIdx.x; c < dv; c += blockDim.x) {
      float4 t = red[c];
      for (int w = 1; w < 8; ++w) {
        const float4 u = red[w * dv + c];
        t.x += u.x;
        t.y += u.y;
        t.z += u.z;
        t.w += u.w;
      }
      dstp[c] = t;
    }
  }
}

// ring geometry: chunks of up to 6 KB -- the bulk-copy engine sustains roughly one operation per 200 cycles per SM, so
// 3 KB rows moved one at a time cap the kernel near 3.7 TB/s -- and as many stages as fit in 112 KB per CTA (2 CTAs
// per SM), at most 8.  max_rpc: 1 or 2 when rows are remapped (window partition: a chunk must stay inside one
// window row and be all-valid or all-padding).
struct LnRing {
  int stages, rpc, smem;
};
static inline LnRing ln_ring_for(uint32_t row_bytes, int max_rpc, size_t extra, int min_stages = 3) {
  LnRing r;
  r.rpc = static_cast<int>(6144 / row_bytes);
  if (r.rpc > max_rpc) r.rpc = max_rpc;
  if (r.rpc < 1) r.rpc = 1;
  if (r.rpc > 8) r.rpc = 8;
  uint32_t slot;
  for (;; --r.rpc) {   // at least 3 chunks in flight per warp (2 if even a single row is that large)
    slot = (static_cast<uint32_t>(r.rpc) * row_bytes + 127) & ~127u;
    r.stages = static_cast<int>((112 * 1024 - extra - 8 * LN_MAX_STAGES * 8) / (8 * static_cast<size_t>(slot)));
    if (r.stages >= min_stages || r.rpc == 1) break;
  }
  if (r.stages > LN_MAX_STAGES) r.stages = LN_MAX_STAGES;
  r.smem = 8 * r.stages * static_cast<int>(slot) + 8 * LN_MAX_STAGES * 8 + static_cast<int>(extra);
  return r;
}

template <int NVT, bool POOL>
static int launch_staged(cudaStream_t st, const AddLnParams& p, int grid, const LnRing& ring) {
  auto go = [&](auto kern) -> int {
    LA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ring.smem));
    kern<<<grid, 256, ring.smem, st>>>(p, ring.stages, ring.rpc);
    LA_CHECK_CUDA(cudaGetLastError());
    return LA_OK;
  };
  // Every lane slot live: the lean row loop -- where the general one is instruction-bound, i.e. for bf16-only rows
  // (norm4 of the prompt encoder 4.2 -> 5.7 TB/s, its mean-pool 2.5 -> 4.2 TB/s) and the window partition of an fp32
  // stream (5.3 -> 5.8 TB/s).  With an fp32 stream in identity order the general loop, which the compiler unrolls
  // across rows, is already at 0.89-0.95 of the copy peak and the lean one is 6-12 % slower (same-box A/B).
  if (p.d == NVT * 128 && p.rows < (1ll << 31) && NVT <= LA_LN_FULL_MAXV) {
    if constexpr (POOL) {
      if (p.delta2) return go(add_layernorm_staged_kernel<NVT, POOL, false, true, true, true>);
    }
    if (!p.x_in) return go(add_layernorm_staged_kernel<NVT, POOL, false, true, false, true>);
    if (p.map_mode == 1 && !p.delta) return go(add_layernorm_staged_kernel<NVT, POOL, true, false, false, true>);
  }
  if constexpr (POOL) {   // the mean-pool of the prompt encoder adds two bf16 streams (image tokens + last MLP output)
    if (p.delta2) return go(add_layernorm_staged_kernel<NVT, POOL, false, true, true>);
  }
  if (p.x_in && p.delta) return go(add_layernorm_staged_kernel<NVT, POOL, true, true>);
  if (p.x_in) return go(add_layernorm_staged_kernel<NVT, POOL, true, false>);
  return go(add_layernorm_staged_kernel<NVT, POOL, false, true>);
}

// out[s, :] = scale * sum_p partial[s, p, :]
__global__ void __launch_bounds__(256)
pool_finish_kernel(const float* __restrict__ partial, float* __restrict__ out, long long n_seq, int slices, int d,
                   float scale) {
  const long long total = n_seq * d;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long s = i / d;
    const int c = static_cast<int>(i % d);
    float t = 0.f;
    for (int q = 0; q < slices; ++q) t += partial[(s * slices + q) * d + c];
    out[i] = t * scale;
  }
}

// x[img*T + tok, :] = (tok < n_cls ? cls : patch[img*(T-n_cls) + tok - n_cls, :]) + pos[tok, :]
__global__ void __launch_bounds__(256)
embed_tokens_kernel(const __nv_bfloat16* __restrict__ patch, const float* __restrict__ cls,
                    const float* __restrict__ pos, float* __restrict__ x, long long n_img, int T, int n_cls, int d) {
  const int dv = d >> 2;
  const long long total = n_img * T * dv;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c4 = static_cast<int>(i % dv);
    const long long rt = i / dv;
    const int tok = static_cast<int>(rt % T);
    const long long img = rt / T;
    float4 a;
    if (tok < n_cls) {
      a = __ldg(reinterpret_cast<const float4*>(cls) + c4);
    } else {
      const uint2 pv =
          reinterpret_cast<const uint2*>(patch + (img * (T - n_cls) + tok - n_cls) * static_cast<long long>(d))[c4];
      const __nv_bfloat162 p01 = *reinterpret_cast<const __nv_bfloat162*>(&pv.x);
      const __nv_bfloat162 p23 = *reinterpret_cast<const __nv_bfloat162*>(&pv.y);
      a = make_float4(__low2float(p01), __high2float(p01), __low2float(p23), __high2float(p23));
    }
    if (pos) {
      const float4 pe = __ldg(reinterpret_cast<const float4*>(pos + static_cast<long long>(tok) * d) + c4);
      a.x += pe.x;
      a.y += pe.y;
      a.z += pe.z;
      a.w += pe.w;
    }
    reinterpret_cast<float4*>(x)[i] = a;
  }
}

// the same with 32-bit indices and multiply-high divisions (two 64-bit divisions per 16 bytes in the kernel above)
__global__ void __launch_bounds__(256)
embed_tokens_fast_kernel(const __nv_bfloat16* __restrict__ patch, const float* __restrict__ cls,
                         const float* __restrict__ pos, float* __restrict__ x, uint32_t total, int T, int n_cls, int d,
                         FastDiv f_dv, FastDiv f_t) {
  const uint32_t dv = static_cast<uint32_t>(d) >> 2;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t rt = fast_div(i, f_dv);
    const uint32_t c4 = i - rt * dv;
    const uint32_t img = fast_div(rt, f_t);
    const int tok = static_cast<int>(rt - img * T);
    float4 a;
    if (tok < n_cls) {
      a = __ldg(reinterpret_cast<const float4*>(cls) + c4);
    } else {
      const uint2 pv = reinterpret_cast<const uint2*>(
          patch + (static_cast<long long>(img) * (T - n_cls) + tok - n_cls) * static_cast<long long>(d))[c4];
      a = make_float4(__uint_as_float(pv.x << 16), __uint_as_float(pv.x & 0xffff0000u), __uint_as_float(pv.y << 16),
                      __uint_as_float(pv.y & 0xffff0000u));
    }
    if (pos) {
      const float4 pe = __ldg(reinterpret_cast<const float4*>(pos + static_cast<long long>(tok) * d) + c4);
      a.x += pe.x;
      a.y += pe.y;
      a.z += pe.z;
      a.w += pe.w;
    }
    reinterpret_cast<float4*>(x)[i] = a;
  }
}

// images [I, C, S, S] fp32 -> rows [I * (S/P)^2, C*P*P] bf16, column = c*P*P + ky*P + kx   (P == 16)
__global__ void __launch_bounds__(256)
im2col_patch16_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, long long n_img, int C, int S) {
  const int g = S / 16;
  // one thread handles 8 consecutive kx (half a patch row): 32 B in, 16 B out
  const long long total = n_img * C * S * (S / 8);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int xs = static_cast<int>(i % (S / 8));
    long long r = i / (S / 8);
    const int y = static_cast<int>(r % S);
    r /= S;
    const int c = static_cast<int>(r % C);
    const long long im = r / C;
    const float4* src = reinterpret_cast<const float4*>(img + ((im * C + c) * S + y) * S + xs * 8);
    const float4 a = __ldg(src), b = __ldg(src + 1);
    const int px = xs >> 1, half = xs & 1, py = y >> 4, ky = y & 15;
    const long long row = (im * g + py) * g + px;
    uint4 pk;
    pk.x = pack_bf16(a.x, a.y);
    pk.y = pack_bf16(a.z, a.w);
    pk.z = pack_bf16(b.x, b.y);
    pk.w = pack_bf16(b.z, b.w);
    *reinterpret_cast<uint4*>(out + row * (C * 256) + c * 256 + ky * 16 + half * 8) = pk;
  }
}

// in [I, H, W, C] bf16 -> out [I*H*W, 9*C] bf16, column = (ky*3 + kx)*C + c, zero padded borders
__global__ void __launch_bounds__(256)
im2col_3x3_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n_img, int H,
                  int W, int C) {
  const int cv = C >> 3;  // uint4 per pixel
  const long long total = n_img * H * W * 9 * cv;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % cv);
    long long r = i / cv;
    const int tap = static_cast<int>(r % 9);
    r /= 9;
    const int x = static_cast<int>(r % W);
    r /= W;
    const int y = static_cast<int>(r % H);
    const long long im = r / H;
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (yy >= 0 && yy < H && xx >= 0 && xx < W)
      v = __ldg(reinterpret_cast<const uint4*>(in + ((im * H + yy) * W + xx) * static_cast<long long>(C)) + c8);
    reinterpret_cast<uint4*>(out)[i] = v;
  }
}

// the same with 32-bit indices and multiply-high divisions (five 64-bit divisions per 16 bytes above)
__global__ void __launch_bounds__(256)
im2col_3x3_fast_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, uint32_t total, int H,
                       int W, int C, FastDiv f_cv, FastDiv f_w, FastDiv f_h) {
  const uint32_t cv = static_cast<uint32_t>(C) >> 3;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t r0 = fast_div(i, f_cv);
    const uint32_t c8 = i - r0 * cv;
    const uint32_t r1 = r0 / 9u;
    const int tap = static_cast<int>(r0 - r1 * 9u);
    const uint32_t r2 = fast_div(r1, f_w);
    const int x = static_cast<int>(r1 - r2 * W);
    const uint32_t im = fast_div(r2, f_h);
    const int y = static_cast<int>(r2 - im * H);
    const int ty = tap / 3;
    const int yy = y + ty - 1, xx = x + (tap - ty * 3) - 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (yy >= 0 && yy < H && xx >= 0 && xx < W)
      v = __ldg(reinterpret_cast<const uint4*>(in + ((static_cast<long long>(im) * H + yy) * W + xx) * static_cast<long long>(C)) + c8);
    reinterpret_cast<uint4*>(out)[i] = v;
  }
}

static int grid_for(long long work_items, int block, int per_sm) {
  long long blocks = (work_items + block - 1) / block;
  const long long cap = static_cast<long long>(sm_count()) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace la

extern "C" {

int la_add_layernorm(void* stream, const float* x_in, long long x_mod, const void* delta, const void* delta2,
                     const float* seq_add, long long seq_rows, float* x_out, const float* gamma, const float* beta,
                     float eps, int act, void* y_out, int y_dtype, float* y2_out, const float* pe, long long pe_mod,
                     void* ype_out, long long rows, int d, int map_mode, int seq_len, int win, int nwin, int hw) {
  using namespace la;
  LA_CHECK_ARG(rows > 0 && d > 0, "la_add_layernorm: empty problem");
  LA_CHECK_ARG(d % 8 == 0 && d <= 128 * ROW_MAXV + 124, "la_add_layernorm: d=%d unsupported (multiple of 8, <= 1404)", d);
  LA_CHECK_ARG(x_in || delta, "la_add_layernorm: need x_in and/or delta");
  LA_CHECK_ARG(!seq_add || seq_rows > 0, "la_add_layernorm: seq_add needs seq_rows");
  LA_CHECK_ARG((gamma == nullptr) == (beta == nullptr), "la_add_layernorm: gamma/beta must come together");
  LA_CHECK_ARG(map_mode >= 0 && map_mode <= 3, "la_add_layernorm: bad map_mode");
  LA_CHECK_ARG(map_mode != 3 || (hw > 0 && rows % (4ll * hw * hw) == 0), "la_add_layernorm: bad pixel-shuffle grid");
  LA_CHECK_ARG(map_mode != 1 || (win > 0 && nwin > 0 && hw > 0 && rows % (static_cast<long long>(win) * win * nwin * nwin) == 0),
               "la_add_layernorm: bad window parameters");
  LA_CHECK_ARG(map_mode != 2 || (seq_len > 1 && rows % seq_len == 0), "la_add_layernorm: bad seq_len");
  LA_CHECK_ARG(!ype_out || (pe && pe_mod > 0), "la_add_layernorm: ype_out needs the pe table");
  LA_CHECK_ARG(act >= LA_ACT_NONE && act <= LA_ACT_RELU, "la_add_layernorm: bad act");
  AddLnParams p = {};
  p.x_in = x_in;
  p.x_mod = x_mod;
  p.delta = static_cast<const __nv_bfloat16*>(delta);
  p.delta2 = static_cast<const __nv_bfloat16*>(delta2);
  p.seq_add = seq_add;
  p.seq_rows = seq_rows;
  p.x_out = x_out;
  p.gamma = gamma;
  p.beta = beta;
  p.eps = eps;
  p.y_out = y_out;
  p.y_f32 = y_dtype == LA_DTYPE_F32;
  p.y2_out = y2_out;
  p.pe = pe;
  p.pe_mod = pe_mod;
  p.ype_out = static_cast<__nv_bfloat16*>(ype_out);
  p.act = act;
  p.rows = rows;
  p.d = d;
  p.map_mode = map_mode;
  p.seq_len = seq_len;
  p.win = win;
  p.nwin = nwin;
  p.hw = hw;
  // (a sequence longer than the staged kernels' 2^31 rows never ends inside them: any divisor above the row count does)
  p.fd_seq = make_fastdiv(seq_rows <= 0 ? 1u : (seq_rows >= (1ll << 31) ? 0x7fffffffu : static_cast<uint32_t>(seq_rows)));
  if (map_mode == 1) {
    p.fd_w2 = make_fastdiv(static_cast<uint32_t>(win) * win);
    p.fd_per_img = make_fastdiv(static_cast<uint32_t>(nwin) * nwin);
    p.fd_nwin = make_fastdiv(static_cast<uint32_t>(nwin));
    p.fd_win = make_fastdiv(static_cast<uint32_t>(win));
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int need = (d + 127) / 128;
  // the ViT block case goes through the staged kernel (bulk async row copies, 3 rows in flight per warp)
  const bool staged = x_mod == 0 && !delta2 && !y2_out && !ype_out && act == LA_ACT_NONE && gamma && y_out &&
                      rows < (1ll << 31) &&
                      (map_mode == 0 || map_mode == 1) && need <= 6 && d % 32 == 0 && rows >= 4096 && !LA_LN_UNSTAGED;
  if (staged) {
    // window partition: pairs of rows (tx, tx + 1) with tx even share a window row and their padding status when the
    // window side and the grid side are even
    const int max_rpc = map_mode == 0 ? 8 : ((win % 2 == 0 && hw % 2 == 0) ? 2 : 1);
    const LnRing ring = ln_ring_for(static_cast<uint32_t>(d) * ((x_in ? 4 : 0) + (delta ? 2 : 0)), max_rpc, 0);
    if (ring.stages >= 2) {
      const int sgrid = grid_for((rows + ring.rpc - 1) / ring.rpc * 32, 256, 2);
      if (need <= 2) return launch_staged<2, false>(st, p, sgrid, ring);
      if (need <= 4) return launch_staged<4, false>(st, p, sgrid, ring);
      return launch_staged<6, false>(st, p, sgrid, ring);   // wider rows (ViT-L) keep the register-load kernel
    }
  }
  const int grid = grid_for(rows * 32, 256, need <= 4 ? 4 : (need <= 8 ? 3 : 2));
  if (need <= 1) add_layernorm_kernel<1, false><<<grid, 256, 0, st>>>(p);
  else if (need <= 2) add_layernorm_kernel<2, false><<<grid, 256, 0, st>>>(p);
  else if (need <= 4) add_layernorm_kernel<4, false><<<grid, 256, 0, st>>>(p);
  else if (need <= 6) add_layernorm_kernel<6, false><<<grid, 256, 0, st>>>(p);
  else if (need <= 8) add_layernorm_kernel<8, false><<<grid, 256, 0, st>>>(p);
  else add_layernorm_kernel<ROW_MAXV + 1, false><<<grid, 256, 0, st>>>(p);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_add_layernorm_meanpool(void* stream, const float* x_in, const void* delta, const void* delta2,
                              const float* seq_add, const float* gamma, const float* beta, float eps,
                              long long n_seq, int rows_per_seq, int d, float* partial_ws, int slices, float* out) {
  using namespace la;
  LA_CHECK_ARG(n_seq > 0 && rows_per_seq > 0 && d > 0 && slices > 0, "la_add_layernorm_meanpool: empty problem");
  LA_CHECK_ARG(d % 8 == 0 && d <= 128 * ROW_MAXV + 124, "la_add_layernorm_meanpool: d=%d unsupported", d);
  LA_CHECK_ARG((x_in || delta) && gamma && beta && partial_ws && out, "la_add_layernorm_meanpool: null pointer");
  LA_CHECK_ARG(n_seq * slices < (1ll << 31), "la_add_layernorm_meanpool: grid too large");
  AddLnParams p = {};
  p.x_in = x_in;
  p.delta = static_cast<const __nv_bfloat16*>(delta);
  p.delta2 = static_cast<const __nv_bfloat16*>(delta2);
  p.seq_add = seq_add;
  p.seq_rows = rows_per_seq;
  p.gamma = gamma;
  p.beta = beta;
  p.eps = eps;
  p.pool_out = partial_ws;
  p.pool_rows = rows_per_seq;
  p.pool_slices = slices;
  p.rows = n_seq * rows_per_seq;
  p.d = d;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = 8 * static_cast<size_t>(d) * sizeof(float);
  const int need = (d + 127) / 128;
  const int grid = static_cast<int>(n_seq * slices);
  // big problems: the staged (bulk-copy ring) variant
  if ((!delta2 || (!x_in && delta)) && need <= 4 && d % 32 == 0 && p.rows >= 4096 && !LA_LN_UNSTAGED) {
    // two chunks in flight per warp are enough here (16 warps x 2 x 6 KB per SM), so the ring prefers large bulk
    // copies over depth; the reduction scratch aliases the ring
    LnRing ring = ln_ring_for(static_cast<uint32_t>(d) * ((x_in ? 4 : 0) + (delta ? 2 : 0) + (delta2 ? 2 : 0)), 8, 0, 2);
    if (ring.stages >= 2) {
      int rc;
      if (ring.smem < static_cast<int>(smem)) ring.smem = static_cast<int>(smem);   // reduction scratch (aliases the ring)
      if (need <= 2) rc = launch_staged<2, true>(st, p, grid, ring);
      else rc = launch_staged<4, true>(st, p, grid, ring);
      if (rc) return rc;
      pool_finish_kernel<<<grid_for(n_seq * d, 256, 8), 256, 0, st>>>(partial_ws, out, n_seq, slices, d,
                                                                     1.0f / rows_per_seq);
      LA_CHECK_CUDA(cudaGetLastError());
      return LA_OK;
    }
  }
  if (need <= 2) add_layernorm_kernel<2, true><<<grid, 256, smem, st>>>(p);
  else if (need <= 4) add_layernorm_kernel<4, true><<<grid, 256, smem, st>>>(p);
  else if (need <= 8) add_layernorm_kernel<8, true><<<grid, 256, smem, st>>>(p);
  else add_layernorm_kernel<ROW_MAXV + 1, true><<<grid, 256, smem, st>>>(p);
  LA_CHECK_CUDA(cudaGetLastError());
  pool_finish_kernel<<<grid_for(n_seq * d, 256, 8), 256, 0, st>>>(partial_ws, out, n_seq, slices, d,
                                                                 1.0f / rows_per_seq);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_embed_tokens(void* stream, const void* patch, const float* cls, const float* pos, float* x, long long n_img,
                    int tokens_per_img, int n_cls, int d) {
  using namespace la;
  LA_CHECK_ARG(patch && x && n_img > 0 && tokens_per_img > n_cls && d % 4 == 0, "la_embed_tokens: bad arguments");
  LA_CHECK_ARG(n_cls == 0 || cls, "la_embed_tokens: cls token missing");
  const long long total = n_img * tokens_per_img * (d / 4);
  if (total < (1ll << 31)) {
    embed_tokens_fast_kernel<<<grid_for(total, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(patch), cls, pos, x, static_cast<uint32_t>(total), tokens_per_img, n_cls, d,
        make_fastdiv(static_cast<uint32_t>(d / 4)), make_fastdiv(static_cast<uint32_t>(tokens_per_img)));
    LA_CHECK_CUDA(cudaGetLastError());
    return LA_OK;
  }
  embed_tokens_kernel<<<grid_for(total, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(patch), cls, pos, x, n_img, tokens_per_img, n_cls, d);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_im2col_patch16(void* stream, const float* images, void* out, long long n_img, int channels, int size) {
  using namespace la;
  LA_CHECK_ARG(images && out && n_img > 0 && channels > 0 && size % 16 == 0, "la_im2col_patch16: bad arguments");
  const long long total = n_img * channels * size * (size / 8);
  im2col_patch16_kernel<<<grid_for(total, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      images, static_cast<__nv_bfloat16*>(out), n_img, channels, size);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

int la_im2col_3x3(void* stream, const void* in, void* out, long long n_img, int height, int width, int channels) {
  using namespace la;
  LA_CHECK_ARG(in && out && n_img > 0 && height > 0 && width > 0 && channels % 8 == 0, "la_im2col_3x3: bad arguments");
  const long long total = n_img * height * width * 9 * (channels / 8);
  if (total < (1ll << 31)) {
    im2col_3x3_fast_kernel<<<grid_for(total, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), static_cast<uint32_t>(total), height,
        width, channels, make_fastdiv(static_cast<uint32_t>(channels / 8)), make_fastdiv(static_cast<uint32_t>(width)),
        make_fastdiv(static_cast<uint32_t>(height)));
    LA_CHECK_CUDA(cudaGetLastError());
    return LA_OK;
  }
  im2col_3x3_kernel<<<grid_for(total, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), n_img, height, width, channels);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

}  // extern "C"
