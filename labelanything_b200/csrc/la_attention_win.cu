// labelanything_b200 — 14x14-window attention of the SAM ViT blocks, second generation (sm_100a).
//
// Replaces label_anything/models/image_encoder.py:239-255 (Attention.forward), 258-304 (window partition /
// unpartition, folded into the output row mapping) and 319-376 (decomposed rel-pos bias, formed in-kernel) for the
// windowed blocks: 196 queries x 196 keys x head_dim 64 per (window, head) item.
//
// The first-generation window mode (la_attention.cu, 112-key tiles) walks every item through a chain of four tensor-
// pipe round trips (T -> S0 -> [P0 -> PV0, S1] -> [P1 -> PV1] -> O), ~9.6 K cycles per item even with no exponential at
// all against an MUFU floor of 3.6 K (profiles/r02_exp_attention.txt).  Here ALL 196 keys of a Q tile are one score
// accumulator (one M128 x N208 product per Q tile, 208 = 196 rounded up to the MMA's N granularity), so an item is
//     [T, S] -> softmax over the whole key row -> P -> PV (K = 208, issued in two parts: keys 0..127 while the last
//     two score chunks are still being exponentiated) -> O
// i.e. two round trips, no running maximum and no O rescale:
//   * TMEM (512 columns): S_A [0, 208), S_B [208, 416), T [416, 480) shared by the two Q tiles in turn (A(i), B(i),
//     A(i+1), ...).  P (bf16) overlays columns [0, 104) of its score region, O columns [104, 168) -- free once the
//     softmax has consumed the scores, and PV is only issued after the whole P row is delivered.
//   * softmax: one thread per query row, two passes over its 196 scores in 28-column chunks (two key-grid rows each),
//     the next chunk's TMEM load in flight while the current one is processed: pass 1 takes the row maximum of
//     scale * s + rel_w + rel_h, pass 2 forms P = exp2(. - max) -> bf16 -> TMEM.  The decomposed rel-pos products
//     T = Q x rel^T come from one extra MMA per Q tile; every thread parks its T row (fp16, 128 B) in a staging row
//     in shared memory and reads the 2 x 14 entries at its (qh, qw) shift back as aligned 32-bit words (+ one byte
//     permute each for an odd shift).  The prologue only needs T, so it runs under the S product.
//   * output: every thread writes its normalised bf16 row (128 B) over its -- long consumed -- Q row in shared memory
//     (the 128B-swizzled image of a [196 tokens][64 channels] box) and the producer warp sends the item out with ONE
//     4-D TMA store (channels, x, y, image; box 64 x 14 x 14 x 1: the window un-partition is the tensor map, rows and
//     columns past the image are dropped by the copy engine) before it refills the stage.  The first version stored
//     from the softmax threads: 1200-1900 cycles per item between O ready and the next prologue.
//   * row sums on the tensor core: next to every K step of PV the issuer queues  L += P x ones  (M128 x N16, 8 cycles)
//     into 16 spare TMEM columns per Q tile, so the softmax threads do no fp32 accumulation of their own (98 packed
//     adds per row: a fifth of the dispatch cycles of the exponential pass, which is what bounds an item) and the
//     normaliser is the sum of exactly the bf16 P values the numerator uses.
//   * padded-grid input (in_pad > 0): Q, K and V of a window are one 4-D TMA box each out of [image][70][70][C]
//     projections whose padding positions hold the bias row, so the windowed blocks project 64 x 64 tokens per image
//     instead of the partitioned 70 x 70 and no window-partition copy exists anywhere.
//   * loads: Q (2 x 128 rows), K and V (208 rows each: the 12 rows past the window are the next window's, finite,
//     and are masked / multiplied by P = 0) through a 2-stage ring.
//   warp 0: TMA producer; warps 1 / 3: MMA issuers of Q tile A / B; warp 2: TMEM allocation;
//   warps 4-7 / 8-11: softmax + epilogue of Q tile A / B.
#include "la_common.cuh"
#include "la_attn_math.cuh"

namespace la {

constexpr int WA_THREADS = 384;
constexpr int WA_KEYS = 196, WA_N = 208, WA_GW = 14;
constexpr int WA_Q_BYTES = 2 * 128 * 128;         // two Q tiles
constexpr int WA_KV_BYTES = WA_N * 128;           // one K or V tile (208 rows)
constexpr int WA_STAGE = WA_Q_BYTES + 2 * WA_KV_BYTES;
constexpr int WA_OFF_REL = 2 * WA_STAGE;
// staging rows of the rel-pos products T (64 fp16 entries per query row; 144-byte stride: conflict-free 16-byte stores)
constexpr int WA_STG_STRIDE = 144;
constexpr int WA_OFF_STG = WA_OFF_REL + 8192;
// sixteen 128-byte rows of bf16 ones: the B operand of the row-sum product (every K step reads the same block)
constexpr int WA_OFF_ONES = WA_OFF_STG + 256 * WA_STG_STRIDE;
constexpr int WA_OFF_BAR = WA_OFF_ONES + 2048;
constexpr int WA_SMEM = WA_OFF_BAR + 512 + 1024;
constexpr int WA_SOFTMAX_REGS = 208, WA_CONTROL_REGS = 88;
// O = P V is issued in two parts: K steps [0, WA_PV_SPLIT) (keys 0..127) after score chunk WA_P_EARLY_CHUNK
constexpr int WA_PV_SPLIT = 8, WA_P_EARLY_CHUNK = 4;
static_assert(28 * (WA_P_EARLY_CHUNK + 1) >= 16 * WA_PV_SPLIT, "the first PV part only reads delivered P columns");
static_assert(28 * (WA_P_EARLY_CHUNK + 2) >= 168, "the score columns under O are in registers when the first PV part may start");
static_assert(256 * WA_SOFTMAX_REGS + 128 * WA_CONTROL_REGS <= WA_THREADS * 168, "setmaxnreg budget");
static_assert(WA_SMEM <= 232448, "shared memory budget");
// share of the exponentials on the FMA pipe: WA_POLY_NUM of every 7 score pairs
#ifndef WA_POLY_NUM
#define WA_POLY_NUM 2
#endif

struct WinParams {
  int n_seq, n_heads;
  int q_off, k_off, v_off;
  float scale_log2;
  int rel_pad;
  __nv_bfloat16* out;
  long long ld_out;
  int out_mode, nwin, img_hw;
  int in_pad;         // > 0: q / kv are padded-grid tensors [image][in_pad][in_pad][ld] (la_gemm_bf16_to_grid): a window is
                      // one 4-D box (64 channels, 14 x, 14 y, 1 image); 0: window-partitioned rows (196 per window)
  long long* trace;   // -DLA_ATT_TRACE builds: clock64 stamps of CTA 0, [role][item][event]; else nullptr
};

// trace roles: 0 / 1 = issuer of Q tile A / B (loads + TMEM free, T/S issued, P seen, PV issued);
// 2 / 3 = first softmax warp of tile A / B (T+S ready, prologue done, pass 1 done, P delivered, O ready, stores issued)
constexpr int WA_TRACE_ITEMS = 64, WA_TRACE_EVENTS = 6;
__device__ __forceinline__ void wa_trace([[maybe_unused]] const WinParams& p, [[maybe_unused]] bool on,
                                         [[maybe_unused]] int role, [[maybe_unused]] int item, [[maybe_unused]] int ev) {
#ifdef LA_ATT_TRACE
  if (on && item < WA_TRACE_ITEMS) p.trace[(role * WA_TRACE_ITEMS + item) * WA_TRACE_EVENTS + ev] = clock64();
#endif
}

// 28 consecutive TMEM columns of this thread's lane -> registers (x16 + x8 + x4)
__device__ __forceinline__ void tmem_ld_28(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23])
               : "r"(taddr + 16) : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]) : "r"(taddr + 24) : "memory");
}
// 14 registers -> 14 consecutive TMEM columns (x8 + x4 + x2)
__device__ __forceinline__ void tmem_st_14(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr + 8), "r"(r[8]), "r"(r[9]),
               "r"(r[10]), "r"(r[11]) : "memory");
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr + 12), "r"(r[12]), "r"(r[13])
               : "memory");
}

__global__ void __launch_bounds__(WA_THREADS, 1)
window_attention_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                        const __grid_constant__ CUtensorMap tm_rel, const __grid_constant__ CUtensorMap tm_out,
                        const WinParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WA_OFF_BAR);
  uint64_t* full = bars;              // [stage]  Q + K + V landed
  uint64_t* stage_free = full + 2;    // [stage]  both PV products of the item completed (two commits)
  uint64_t* bar_t = stage_free + 2;   // [tile]   T ready
  uint64_t* bar_s = bar_t + 2;        // [tile]   S ready
  uint64_t* bar_p = bar_s + 2;        // [tile]   P of keys 0..139 delivered and score columns < 168 consumed (4 warps)
  uint64_t* bar_p2 = bar_p + 2;       // [tile]   whole P row delivered (4 warps)
  uint64_t* bar_o = bar_p2 + 2;       // [tile]   O ready
  uint64_t* o_free = bar_o + 2;       // [tile]   O read by the epilogue (4 warps): the score region may be overwritten
  uint64_t* t_done = o_free + 2;      // [tile]   T read by the prologue (4 warps): the T columns may be overwritten
  uint64_t* out_full = t_done + 2;    // [stage]  the item's output rows are written over its Q rows (8 warps)
  uint64_t* rel_full = out_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rel_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.n_seq * p.n_heads;          // item = (window sequence, head), head fastest
  const int n_my = (n_items > static_cast<int>(blockIdx.x))
                       ? (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                       : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    tma_prefetch_desc(&tm_rel);
    tma_prefetch_desc(&tm_out);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&stage_free[s], 2);
      mbar_init(&bar_t[s], 1);
      mbar_init(&bar_s[s], 1);
      mbar_init(&bar_p[s], 4);
      mbar_init(&bar_p2[s], 4);
      mbar_init(&bar_o[s], 1);
      mbar_init(&o_free[s], 4);
      mbar_init(&t_done[s], 4);
      mbar_init(&out_full[s], 8);
    }
    mbar_init(rel_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 2048 / 4; i += WA_THREADS) reinterpret_cast<uint32_t*>(smem + WA_OFF_ONES)[i] = 0x3f803f80u;
  if (p.in_pad > 0) {
    // the 4-D boxes bring 196 rows: rows 196..207 of every K / V tile are never written -- make them zero once
    constexpr int TAIL = (WA_N - WA_KEYS) * 128 / 16;   // uint4 per tile
    for (int i = threadIdx.x; i < 4 * TAIL; i += WA_THREADS) {
      const int tile = i / TAIL;   // (stage, K | V)
      uint8_t* base = smem + (tile >> 1) * WA_STAGE + WA_Q_BYTES + (tile & 1) * WA_KV_BYTES + WA_KEYS * 128;
      reinterpret_cast<uint4*>(base)[i - tile * TAIL] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t TM_T = 416, TM_O_IN_S = 104, TM_L = 480;   // L_A [480, 496), L_B [496, 512)

  if (warp < 4) {
    setmaxnreg_dec<WA_CONTROL_REGS>();
    if (warp == 0) {
      // ------------------------------------ TMA producer ------------------------------------
      if (lane == 0) {
        mbar_arrive_expect_tx(rel_full, 8192);
        tma_load_2d(smem + WA_OFF_REL, &tm_rel, rel_full, 0, 0);
        // item j's output rows (written over its Q rows by the softmax warps) -> global memory
        auto store_item = [&](const int j) {
          const int w = blockIdx.x + j * gridDim.x;
          const int head = w % p.n_heads, seq = w / p.n_heads;
          mbar_wait(&out_full[j & 1], (j >> 1) & 1);
          const uint8_t* src = smem + (j & 1) * WA_STAGE;
          if (p.out_mode == 0) {
            tma_store_4d(&tm_out, src, head * 64, 0, seq, 0);
          } else {
            const int per_img = p.nwin * p.nwin;
            const int img = seq / per_img, wi = seq - img * per_img;
            const int wy = wi / p.nwin;
            tma_store_4d(&tm_out, src, head * 64, (wi - wy * p.nwin) * WA_GW, wy * WA_GW, img);
          }
          tma_store_commit();
        };
        for (int it = 0; it < n_my; ++it) {
          const int w = blockIdx.x + it * gridDim.x;
          const int head = w % p.n_heads, seq = w / p.n_heads;
          const int row0 = seq * WA_KEYS;
          const int st = it & 1;
          mbar_wait(&stage_free[st], ((it >> 1) & 1) ^ 1);
          if (it >= 2) {
            store_item(it - 2);
            tma_store_wait_read<0>();   // the copy engine has read the stage: it may be refilled
          }
          uint8_t* base = smem + st * WA_STAGE;
          if (p.in_pad > 0) {
            const int per_img = p.nwin * p.nwin;
            const int img = seq / per_img, wi = seq - img * per_img;
            const int wy = wi / p.nwin;
            const int x0 = (wi - wy * p.nwin) * WA_GW, y0 = wy * WA_GW;
            mbar_arrive_expect_tx(&full[st], 3 * WA_KEYS * 128);
            tma_load_4d(base, &tm_q, &full[st], p.q_off + head * 64, x0, y0, img);
            tma_load_4d(base + WA_Q_BYTES, &tm_kv, &full[st], p.k_off + head * 64, x0, y0, img);
            tma_load_4d(base + WA_Q_BYTES + WA_KV_BYTES, &tm_kv, &full[st], p.v_off + head * 64, x0, y0, img);
          } else {
            mbar_arrive_expect_tx(&full[st], WA_STAGE);
            tma_load_2d(base, &tm_q, &full[st], p.q_off + head * 64, row0);
            tma_load_2d(base + 16384, &tm_q, &full[st], p.q_off + head * 64, row0 + 128);
            tma_load_2d(base + WA_Q_BYTES, &tm_kv, &full[st], p.k_off + head * 64, row0);
            tma_load_2d(base + WA_Q_BYTES + WA_KV_BYTES, &tm_kv, &full[st], p.v_off + head * 64, row0);
          }
        }
        for (int j = n_my > 2 ? n_my - 2 : 0; j < n_my; ++j) store_item(j);
        tma_store_wait_read<0>();
      }
    } else if (warp == 1 || warp == 3) {
      // ------------------------------------ MMA issuers -------------------------------------
      // Per item and Q tile x: T_x = Q_x rel^T and S_x = Q_x K^T as soon as the loads are there, the T columns have
      // been released by the other tile's prologue and the previous item's O_x (which overlays S_x) has been read;
      // then, when the softmax warps have delivered the whole P row,  O_x = P_x V  (K = 208).
      constexpr uint32_t idesc_t = umma_idesc_bf16(128, 64, 0, 0);
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, WA_N, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);     // P from TMEM, V MN-major
      constexpr uint32_t idesc_l = umma_idesc_bf16(128, 16, 0, 1);     // P from TMEM, ones (MN-major, any 16 rows)
      const int x = warp == 1 ? 0 : 1;
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t lo_rel = desc_lo(smem_base + WA_OFF_REL);
      const uint32_t tm_s = tm + x * WA_N;
      const uint32_t tm_l = tm + TM_L + x * 16;
      const uint32_t lo_ones = desc_lo(smem_base + WA_OFF_ONES);
      const bool tri = p.trace != nullptr && blockIdx.x == 0 && lane == 0;
      for (int it = 0; it < n_my; ++it) {
        const int st = it & 1;
        const uint32_t lo_q = desc_lo(smem_base + st * WA_STAGE + x * 16384);
        const uint32_t lo_k = desc_lo(smem_base + st * WA_STAGE + WA_Q_BYTES);
        const uint32_t lo_v = desc_lo(smem_base + st * WA_STAGE + WA_Q_BYTES + WA_KV_BYTES);
        mbar_wait(&full[st], (it >> 1) & 1);
        if (it == 0) mbar_wait(rel_full, 0);
        // T columns: users in the order A(0), B(0), A(1), B(1), ...
        if (x == 0) {
          if (it > 0) mbar_wait(&t_done[1], (it - 1) & 1);
        } else {
          mbar_wait(&t_done[0], it & 1);
        }
        if (it > 0) mbar_wait(&o_free[x], (it - 1) & 1);
        tc_fence_after();
        wa_trace(p, tri, x, it, 0);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_ss_lo(tm + TM_T, lo_q + 2 * ks, lo_rel + 2 * ks, idesc_t, ks > 0);
          umma_commit(&bar_t[x]);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_ss_lo(tm_s, lo_q + 2 * ks, lo_k + 2 * ks, idesc_s, ks > 0);
          umma_commit(&bar_s[x]);
        }
        __syncwarp();
        wa_trace(p, tri, x, it, 1);
        // O = P V in two parts: keys 0..127 as soon as the softmax warps have delivered them (and hold the score
        // columns O overlays in registers), the rest with the end of the row -- the first part runs under the last
        // two score chunks
        mbar_wait(&bar_p[x], it & 1);
        tc_fence_after();
        wa_trace(p, tri, x, it, 2);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < WA_PV_SPLIT; ++ks) {
            umma_ts_lo(tm_s + TM_O_IN_S, tm_s + ks * 8, lo_v + ks * (2048 >> 4), idesc_o, ks > 0);
            umma_ts_lo(tm_l, tm_s + ks * 8, lo_ones, idesc_l, ks > 0);
          }
        }
        __syncwarp();
        mbar_wait(&bar_p2[x], it & 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = WA_PV_SPLIT; ks < WA_N / 16; ++ks) {
            umma_ts_lo(tm_s + TM_O_IN_S, tm_s + ks * 8, lo_v + ks * (2048 >> 4), idesc_o, true);
            umma_ts_lo(tm_l, tm_s + ks * 8, lo_ones, idesc_l, true);
          }
          umma_commit(&bar_o[x]);
          umma_commit(&stage_free[st]);
        }
        __syncwarp();
        wa_trace(p, tri, x, it, 3);
      }
    }
  } else {
    // ===================================== softmax warpgroups =====================================
    setmaxnreg_inc<WA_SOFTMAX_REGS>();
    const int x = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                 // row inside the Q tile
    const int t = x * 128 + r;                         // token inside the window
    const bool row_valid = t < WA_KEYS;
    const int ty = t / WA_GW, tx = t - ty * WA_GW;
    const int qh = row_valid ? ty : 0, qw = row_valid ? tx : 0;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t t_s = tmem_base + lane_addr + x * WA_N;
    const uint32_t t_t = tmem_base + lane_addr + TM_T;
    const uint32_t t_l = tmem_base + lane_addr + TM_L + x * 16;
    const float sl2 = p.scale_log2;
    constexpr float LOG2E = 1.4426950408889634f;
    // a whole warp without a valid row (rows 224..255 of tile B) only keeps the barriers moving
    const bool warp_live = x * 128 + quarter * 32 < WA_KEYS;
    const bool trs = p.trace != nullptr && blockIdx.x == 0 && quarter == 0 && lane == 0;

    // item (seq, head) of this CTA's it-th item, advanced without divisions
    int head = static_cast<int>(blockIdx.x) % p.n_heads, seq = static_cast<int>(blockIdx.x) / p.n_heads;
    const int d_head = static_cast<int>(gridDim.x) % p.n_heads, d_seq = static_cast<int>(gridDim.x) / p.n_heads;
    for (int it = 0; it < n_my; ++it) {
      const int st = it & 1;
      const uint32_t par = it & 1;

      // ---- rel-pos prologue: T row -> fp16 -> staging row in shared memory -> the 2 x 14 shifted entries ----
      // (trace build of the first version, one scalar LDS + conversion + address arithmetic per entry through the
      //  swizzled Q row: 1824 of the 8715 cycles of an item, gpurun_out/r2_trace_window.log)
      mbar_wait(&bar_t[x], par);
      tc_fence_after();
      wa_trace(p, trs, 2 + x, it, 0);
      float rw2[WA_GW], rh2[WA_GW];
      {
        uint32_t tb[64];
        tmem_ld_x32(t_t, tb);
        tmem_ld_x32(t_t + 32, tb + 32);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_done[x]);
        uint8_t* srow = smem + WA_OFF_STG + (x * 128 + r) * WA_STG_STRIDE;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4 v;
          v.x = pack_f16(__uint_as_float(tb[8 * c + 0]), __uint_as_float(tb[8 * c + 1]));
          v.y = pack_f16(__uint_as_float(tb[8 * c + 2]), __uint_as_float(tb[8 * c + 3]));
          v.z = pack_f16(__uint_as_float(tb[8 * c + 4]), __uint_as_float(tb[8 * c + 5]));
          v.w = pack_f16(__uint_as_float(tb[8 * c + 6]), __uint_as_float(tb[8 * c + 7]));
          *reinterpret_cast<uint4*>(srow + 16 * c) = v;
        }
        __syncwarp();   // (own row only: this is the compiler-level ordering of the stores above and the loads below)
        // bias of key (kh, kw) for a query at (qh, qw): T[13 - qh + kh] + T[rel_pad + 13 - qw + kw]: 14 consecutive
        // fp16 entries from entry s = 13 - qh (rel_pad + 13 - qw): the eight aligned words from word s / 2
        auto shifted = [&](const int s0, float* dst) {
          const uint32_t* wsrc = reinterpret_cast<const uint32_t*>(srow) + (s0 >> 1);
          const uint32_t sel = (s0 & 1) ? 0x5432u : 0x3210u;
          uint32_t wv[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) wv[k] = wsrc[k];
#pragma unroll
          for (int i = 0; i < 7; ++i) {
            const uint32_t pr = __byte_perm(wv[i], wv[i + 1], sel);
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&pr));
            fmul2s(dst[2 * i], dst[2 * i + 1], f.x, f.y, LOG2E);
          }
        };
        shifted(WA_GW - 1 - qh, rh2);
        shifted(p.rel_pad + WA_GW - 1 - qw, rw2);
      }
      mbar_wait(&bar_s[x], par);
      tc_fence_after();
      wa_trace(p, trs, 2 + x, it, 1);
      if (warp_live) {
        // ---- pass 1: row maximum of  scale * s + rel_w + rel_h  over the 196 keys, 28 columns (2 key rows) at a time ----
        uint32_t ca[28], cb[28];
        float mx = -INFINITY;
        tmem_ld_28(t_s, ca);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 7; ++c) {
          uint32_t* cur = (c & 1) ? cb : ca;
          uint32_t* nxt = (c & 1) ? ca : cb;
          if (c + 1 < 7) tmem_ld_28(t_s + 28 * (c + 1), nxt);
#pragma unroll
          for (int gi = 0; gi < 2; ++gi) {
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int i = 0; i < WA_GW; i += 2) {
              float a0, a1;
              ffma2v(a0, a1, __uint_as_float(cur[gi * WA_GW + i]), __uint_as_float(cur[gi * WA_GW + i + 1]), sl2, rw2[i],
                     rw2[i + 1]);
              if (i & 2) m1 = max3(m1, a0, a1);
              else m0 = max3(m0, a0, a1);
            }
            mx = fmaxf(mx, fmaxf(m0, m1) + rh2[2 * c + gi]);
          }
          if (c + 1 < 7) tmem_ld_wait();
        }
        wa_trace(p, trs, 2 + x, it, 2);
        // ---- pass 2: P = exp2(. - max) -> bf16 -> TMEM over the consumed score columns (row sum: tensor core) ----
        tmem_ld_28(t_s, ca);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 7; ++c) {
          uint32_t* cur = (c & 1) ? cb : ca;
          uint32_t* nxt = (c & 1) ? ca : cb;
          if (c + 1 < 7) tmem_ld_28(t_s + 28 * (c + 1), nxt);
          uint32_t pk[14];
#pragma unroll
          for (int gi = 0; gi < 2; ++gi) {
            const float off = rh2[2 * c + gi] - mx;
#pragma unroll
            for (int i = 0; i < WA_GW; i += 2) {
              float a0, a1;
              ffma2v(a0, a1, __uint_as_float(cur[gi * WA_GW + i]), __uint_as_float(cur[gi * WA_GW + i + 1]), sl2, rw2[i],
                     rw2[i + 1]);
              fadd2s(a0, a1, a0, a1, off);
              float e0, e1;
              if ((i >> 1) < WA_POLY_NUM) {
                e0 = a0;
                e1 = a1;
                exp2_poly_x2(e0, e1);
              } else {
                e0 = ex2_approx(a0);
                e1 = ex2_approx(a1);
              }
              pk[(gi * WA_GW + i) >> 1] = pack_bf16(e0, e1);
            }
          }
          if (c + 1 < 7) tmem_ld_wait();        // the next chunk is in registers before its columns are overwritten below
          tmem_st_14(t_s + 14 * c, pk);          // columns [14c, 14c+14) <= 28c: already consumed
          if (c == WA_P_EARLY_CHUNK) {
            // P of keys 0..139 is stored and score chunk c + 1 (columns up to 168) sits in registers
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_p[x]);
          }
        }
      } else {
        if (lane == 0) mbar_arrive(&bar_p[x]);   // a warp without rows keeps the barriers moving
      }
      {
        // keys 196..207 (the next window's rows) get P = 0: columns 98..103.  (Rows without a query -- a whole warp of
        // tile B -- keep whatever their lanes hold: rows are independent in the PV product and theirs are never stored.)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(t_s + 98), "r"(0u), "r"(0u),
                     "r"(0u), "r"(0u) : "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(t_s + 102), "r"(0u), "r"(0u) : "memory");
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_p2[x]);
      wa_trace(p, trs, 2 + x, it, 3);

      // ---- epilogue: O / l -> bf16 -> global (window un-partition folded into the row mapping) ----
      mbar_wait(&bar_o[x], par);
      tc_fence_after();
      wa_trace(p, trs, 2 + x, it, 4);
      uint32_t ov[64];
      uint32_t lsum;
      tmem_ld_x32(t_s + TM_O_IN_S, ov);
      tmem_ld_x32(t_s + TM_O_IN_S + 32, ov + 32);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(lsum) : "r"(t_l) : "memory");
      tmem_ld_wait();
      const float inv_l = 1.0f / __uint_as_float(lsum);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[x]);
      // my row, normalised, as bf16 over my Q row of this item's stage (128B-swizzled like the load that filled it):
      // Q was last read by the S product, and the stage is only refilled after the producer has stored these rows
      {
        uint8_t* orow = smem + st * WA_STAGE + x * 16384 + r * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float o0, o1, o2, o3, o4, o5, o6, o7;
          fmul2s(o0, o1, __uint_as_float(ov[8 * c + 0]), __uint_as_float(ov[8 * c + 1]), inv_l);
          fmul2s(o2, o3, __uint_as_float(ov[8 * c + 2]), __uint_as_float(ov[8 * c + 3]), inv_l);
          fmul2s(o4, o5, __uint_as_float(ov[8 * c + 4]), __uint_as_float(ov[8 * c + 5]), inv_l);
          fmul2s(o6, o7, __uint_as_float(ov[8 * c + 6]), __uint_as_float(ov[8 * c + 7]), inv_l);
          *reinterpret_cast<uint4*>(orow + ((c ^ (r & 7)) << 4)) =
              make_uint4(pack_bf16(o0, o1), pack_bf16(o2, o3), pack_bf16(o4, o5), pack_bf16(o6, o7));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&out_full[st]);
      }
      wa_trace(p, trs, 2 + x, it, 5);
      head += d_head;
      seq += d_seq;
      if (head >= p.n_heads) {
        head -= p.n_heads;
        ++seq;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace la

// Second-generation entry behind la_attention_window_bf16 (same contract, see the public header).
int la_attention_window_v2(void* stream, const void* q, long long ld_q, int q_off, const void* kv, long long ld_kv,
                           int k_off, int v_off, long long rows_total, int n_seq, int n_heads, float scale,
                           const void* rel_table, int rel_pad, void* out, long long ld_out, int out_mode, int nwin,
                           int img_hw, int in_pad, long long* trace) {
  using namespace la;
  CUtensorMap tm_q, tm_kv, tm_rel;
  int rc;
  if (in_pad > 0) {
    // (channel, x, y, image) over the padded grids; one window = one box
    const uint64_t pd = static_cast<uint64_t>(in_pad), n_img = static_cast<uint64_t>(n_seq / (nwin * nwin));
    const uint32_t box[4] = {64, WA_GW, WA_GW, 1};
    const uint64_t dq[4] = {static_cast<uint64_t>(ld_q), pd, pd, n_img};
    const uint64_t sq[3] = {static_cast<uint64_t>(ld_q) * 2, static_cast<uint64_t>(ld_q) * 2 * pd,
                            static_cast<uint64_t>(ld_q) * 2 * pd * pd};
    rc = make_tensor_map_4d(&tm_q, q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dq, sq, box, Swizzle::B128);
    if (rc) return rc;
    const uint64_t dk[4] = {static_cast<uint64_t>(ld_kv), pd, pd, n_img};
    const uint64_t sk[3] = {static_cast<uint64_t>(ld_kv) * 2, static_cast<uint64_t>(ld_kv) * 2 * pd,
                            static_cast<uint64_t>(ld_kv) * 2 * pd * pd};
    rc = make_tensor_map_4d(&tm_kv, kv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dk, sk, box, Swizzle::B128);
    if (rc) return rc;
  } else {
    rc = make_tensor_map_2d(&tm_q, q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (uint64_t)ld_q, (uint64_t)rows_total,
                            (uint64_t)ld_q * 2, 64, 128, Swizzle::B128);
    if (rc) return rc;
    rc = make_tensor_map_2d(&tm_kv, kv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (uint64_t)ld_kv, (uint64_t)rows_total,
                            (uint64_t)ld_kv * 2, 64, WA_N, Swizzle::B128);
    if (rc) return rc;
  }
  rc = make_tensor_map_2d(&tm_rel, rel_table, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 64, 64, 128, 64, 64, Swizzle::B128);
  if (rc) return rc;
  WinParams p;
  p.n_seq = n_seq;
  p.n_heads = n_heads;
  p.q_off = q_off;
  p.k_off = k_off;
  p.v_off = v_off;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.rel_pad = rel_pad;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ld_out = ld_out;
  p.out_mode = out_mode;
  p.nwin = nwin;
  p.img_hw = img_hw;
  p.in_pad = in_pad;
  p.trace = trace;
  // the output as (channel, x, y, image) [window un-partition] or (channel, token, window sequence, 1): one box per item
  CUtensorMap tm_out;
  {
    const uint64_t row_bytes = static_cast<uint64_t>(ld_out) * 2;
    uint64_t dims[4], strides[3];
    uint32_t box[4];
    if (out_mode == 0) {
      dims[0] = static_cast<uint64_t>(ld_out), dims[1] = WA_KEYS, dims[2] = static_cast<uint64_t>(n_seq), dims[3] = 1;
      strides[0] = row_bytes, strides[1] = row_bytes * WA_KEYS, strides[2] = row_bytes * WA_KEYS * n_seq;
      box[0] = 64, box[1] = WA_KEYS, box[2] = 1, box[3] = 1;
    } else {
      const uint64_t hw = static_cast<uint64_t>(img_hw);
      dims[0] = static_cast<uint64_t>(ld_out), dims[1] = hw, dims[2] = hw, dims[3] = static_cast<uint64_t>(n_seq / (nwin * nwin));
      strides[0] = row_bytes, strides[1] = row_bytes * hw, strides[2] = row_bytes * hw * hw;
      box[0] = 64, box[1] = WA_GW, box[2] = WA_GW, box[3] = 1;
    }
    rc = make_tensor_map_4d(&tm_out, out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dims, strides, box, Swizzle::B128);
    if (rc) return rc;
  }
  LA_CHECK_CUDA(cudaFuncSetAttribute(window_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WA_SMEM));
  const long long items = static_cast<long long>(n_seq) * n_heads;
  const int grid = items < sm_count() ? static_cast<int>(items) : sm_count();
  window_attention_kernel<<<grid, WA_THREADS, WA_SMEM, static_cast<cudaStream_t>(stream)>>>(tm_q, tm_kv, tm_rel, tm_out,
                                                                                            p);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}
