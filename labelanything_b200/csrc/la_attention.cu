// labelanything_b200 — fused multi-head self-attention (QK^T -> [+ decomposed rel-pos bias] -> softmax -> PV)
// on tcgen05 tensor cores, head_dim = 64 (sm_100a).
//
// Replaces the materialised attention of the reference's ViT blocks:
//   label_anything/models/image_encoder.py:239-255 (Attention.forward), 340-376 (add_decomposed_rel_pos),
//   258-304 (window partition / unpartition — folded into the output row mapping), and the HF ViT
//   self-attention used by the MAE encoders (transformers/models/vit/modeling_vit.py:199-250).
//
// Persistent kernel: one CTA per SM walks a static round-robin list of work items; an item is 256 consecutive
// query rows of one (sequence, head): two 128-row Q tiles ("A", "B") that ping-pong on the tensor core while
// their softmax warpgroups run on the CUDA cores (FlashAttention-style online softmax, accumulator O kept in
// TMEM and rescaled only when the running max grows by > 2^8).  All pipelines (Q double buffer, K / V rings,
// score buffers) keep running across item boundaries, so the loads and the first QK^T of item i+1 overlap the
// last softmax tiles and the epilogue of item i -- this is what keeps the 14x14-window blocks (2 key tiles per
// item) from being dominated by per-CTA set-up and load latency.
//   warp 0      : TMA producer (Q once; K and V tiles through two independent smem rings)
//   warps 1, 3  : MMA issuers for Q tile A / B (S = Q K^T into TMEM;  O += P V with P read from TMEM -- it overlays
//                 the scores it was computed from -- and V as MN-major smem operand)
//   warp 2      : TMEM allocator; 64x64 rel-pos mode: identity operand + rel_w operand builder
//   warps 4-7   : softmax + epilogue for Q tile A     warps 8-11: same for Q tile B   (one thread per query row,
//                 the whole score row of a key tile held in registers: setmaxnreg gives these warps 216 registers)
// The decomposed relative-position bias  rel_h[q, kh] + rel_w[q, kw]  is read from fp32 tables produced
// by la_gemm_bf16 (q_head @ reversed_table^T), already shifted so that entry (gh-1 - qh + kh) is the bias
// of key row kh for a query in grid row qh.
//   * 14x14 windows: both terms are added by the softmax threads (14 + 8 registers per row).
//   * 64x64 global blocks: a key tile is one key-grid row, so rel_h is ONE scalar per (query, tile) folded into the
//     softmax offset, and rel_w[q, 0..63] is the same for every tile: warps 2/3 write it once per item as an fp16
//     A operand and the MMA warp accumulates  A_w x (I / scale)  on top of  Q K^T  in TMEM (kind::f16, fp16 inputs),
//     so the softmax threads see scores that already contain it -- no per-element bias arithmetic, no 64-register
//     rel_w row (the MUFU/issue-bound softmax loop is the limiter of this kernel, the tensor pipe has slack).
#include "la_common.cuh"
#include "la_attn_math.cuh"
#include <type_traits>
#include <cuda_fp16.h>

namespace la {

constexpr int ATT_D = 64;
// Threads per query row of the 64-key modes (template parameter SPLIT of the kernel): 1 = one softmax thread per row
// (8 softmax warps, 2 per scheduler), 2 = the 64 score columns of a tile are split between two threads of different
// warps (16 softmax warps, 4 per scheduler).  The softmax warps bound the 64x64 mode (in-order issue at IPC ~0.55 with
// two warps per scheduler, profiles/r02_ncu_attention_global_v1.txt); twice the warps hide each other's MUFU / TMEM /
// barrier latencies.  See the SPLIT == 2 softmax block for how the two threads of a row agree on the running maximum.
#ifndef ATT_SPLIT
#define ATT_SPLIT 1
#endif
// SPLIT == 2, 64x64 rel-pos mode: 1 = rel_w through the extra score MMAs (as with SPLIT == 1), 0 = added by the threads
#ifndef ATT_SPLIT_FOLD
#define ATT_SPLIT_FOLD 0
#endif
constexpr int att_threads(int split) { return 128 + 256 * split; }
// ATT_POLY_NUM of every ATT_POLY_DEN pairs of scores of the 64-key modes take the polynomial exp2 (evenly spread)
#ifndef ATT_POLY_NUM
#define ATT_POLY_NUM 1
#endif
#ifndef ATT_POLY_DEN
#define ATT_POLY_DEN 4
#endif
// ... and of the 112-key window mode (0: measured no gain there -- that kernel is bound by its per-item latency chain)
#ifndef ATT_POLY_NUM_WIN
#define ATT_POLY_NUM_WIN 0
#endif
#ifndef ATT_TS_OPERANDS
#define ATT_TS_OPERANDS 1
#endif
// 64-key modes: the softmax threads fetch score tile g+1 into a second register set right after the exponentials of
// tile g: the barrier wait and the TMEM load latency run under the P store / fence / arrive of tile g.
#ifndef ATT_PIPE
#define ATT_PIPE 0
#endif
// 64-key modes: the exponential passes of the two Q tiles' softmax warps take turns (FlashAttention-3/4 style
// ping-pong).  Warp 4+q (tile A) and warp 8+q (tile B) share scheduler q and its MUFU unit; left alone they run in
// lockstep (both wait, both take the maximum, both compete for the MUFU: ~940 cycles of exponentials per tile for
// 2 x 384 cycles of MUFU work, with the unit idle during the other ~500 cycles of every tile period).  A pair of
// named barriers per scheduler hands the exponential pass back and forth, so one warp's barrier waits, TMEM loads
// and maximum run under the other's exponentials.
// element type of P (the A operand of the PV product): 1 = fp16, 0 = bf16
#ifndef ATT_P_F16
#define ATT_P_F16 0
#endif
// timing diagnostics (WRONG results; experiment builds only): skip the rel_w fold MMAs / the MUFU exponentials
#ifndef ATT_DIAG_NOFOLD
#define ATT_DIAG_NOFOLD 0
#endif
#ifndef ATT_DIAG_NOEXP
#define ATT_DIAG_NOEXP 0
#endif
#ifndef ATT_PINGPONG
#define ATT_PINGPONG 2
#endif
// 64-key modes: the tile maximum is taken on every ATT_MAX_EVERY-th key tile (see pass 1 of the softmax warps)
#ifndef ATT_EPI_FMUL2
#define ATT_EPI_FMUL2 1
#endif
#ifndef ATT_EPI_WIDE
#define ATT_EPI_WIDE 1
#endif
#ifndef ATT_MAX_EVERY
#define ATT_MAX_EVERY 1
#endif
constexpr int ATT_STG_STRIDE = 272;     // bytes per row of the table staging area (68 floats: conflict-free STS.128)
// setmaxnreg budget: 256 softmax threads + 128 control threads share 384 x 168 = 64512 registers (launch allocation).
// 64-key tiles keep 64 score registers per thread and leave the control warps 104; the 112-key window tiles need
// everything the softmax threads can get.
template <int KV_TILE, int SPLIT = 1>
struct AttRegs {
  // launch allocation: 168 registers x 384 threads, 96 x 640 threads (SPLIT == 2: 32 score registers per thread)
  static constexpr int LAUNCH = SPLIT == 2 ? 96 : 168;
  static constexpr int SOFTMAX = SPLIT == 2 ? 104 : 216;
  static constexpr int CONTROL = SPLIT == 2 ? 56 : 72;
  static_assert(256 * SPLIT * SOFTMAX + 128 * CONTROL <= att_threads(SPLIT) * LAUNCH,
                "setmaxnreg.inc can only hand out what the CTA was launched with");
};

enum AttBias : int { ATT_BIAS_NONE = 0, ATT_BIAS_GLOBAL64 = 1, ATT_BIAS_WINDOW14 = 2 };

struct AttParams {
  int n_seq, seq_len, n_heads;
  int q_off, k_off, v_off;  // column (element) offsets of head 0 inside a qkv row
  long long rows_total;     // rows in the qkv matrix
  float scale_log2;         // softmax scale * log2(e)
  float inv_scale;          // 1 / softmax scale
  // rel-pos bias tables: [rows_total][n_heads][ldb] fp32 (nullptr for ATT_BIAS_NONE)
  const void* bias_h;   // fp32 or fp16 tables (template parameter TF16)
  const void* bias_w;
  int ldb;
  int rel_pad;              // window mode: rows of each (h / w) half of the reversed rel-pos operand (32)
  // output
  __nv_bfloat16* out;
  long long ld_out;
  int out_mode;  // 0: row = seq*seq_len + t ; 1: window unpartition
  int win, nwin, img_hw;  // out_mode 1: window size, windows per side, un-padded grid side
  int wide_store;         // output rows are 32-byte aligned: 256-bit stores in the epilogue
  long long* trace;       // -DLA_ATT_TRACE builds: clock64 stamps of CTA (0,0,0) (la_attention_set_trace), else nullptr
};

// trace layout: [role][tile][event] int64; roles: 0 = MMA issuers (P seen / next S issued, per Q tile), 1 / 2 = softmax
// A / B first warp (wait S, got S, max done, P delivered), 3 / 4 = the same warps' item boundaries, indexed by item
// (O ready, epilogue stores issued, window tables ready, prologue done)
constexpr int ATT_TRACE_TILES = 192, ATT_TRACE_EVENTS = 4;   // the first three items of a 64-tile-per-item run
// Compiled in only with -DLA_ATT_TRACE (experiment builds, labelanything_b200/build.py::build_variant): the product
// library has no trace state and no environment lookups on the launch path.
__device__ __forceinline__ void att_trace([[maybe_unused]] const AttParams& p, [[maybe_unused]] bool on,
                                          [[maybe_unused]] int role, [[maybe_unused]] int tile,
                                          [[maybe_unused]] int ev) {
#ifdef LA_ATT_TRACE
  if (on && tile < ATT_TRACE_TILES) p.trace[(role * ATT_TRACE_TILES + tile) * ATT_TRACE_EVENTS + ev] = clock64();
#endif
}
// diagnostics (compile time): bit0 always rescale, bit1 never skip the O wait
#ifndef LA_ATT_DEBUG_FLAGS
#define LA_ATT_DEBUG_FLAGS 0
#endif

template <int KV_TILE, int BIAS, int SPLIT = 1>
struct AttSmem {
  static constexpr int STAGES = KV_TILE <= 64 ? (BIAS == 1 ? (SPLIT == 2 ? 4 : 5) : 6) : 3;   // K / V ring depth
  static constexpr int Q_BYTES = 2 * 128 * 128;             // two Q tiles, 128 rows x 128 B
  static constexpr int KV_BYTES = KV_TILE * 128;            // one K or V tile
  static constexpr int KV_SLOT = ((KV_BYTES + 1023) / 1024) * 1024;
  static constexpr int OFF_K = 2 * Q_BYTES;                 // Q is double-buffered across work items
  static constexpr int OFF_V = OFF_K + STAGES * KV_SLOT;
  // 64x64 rel-pos mode: rel_w A operands [item parity][Q tile] (128 rows x 128 B, 128B-swizzled) + 64x64 identity
  static constexpr int OFF_AW = OFF_V + STAGES * KV_SLOT;
  static constexpr int AW_BYTES = BIAS == 1 ? 4 * 16384 : 0;
  static constexpr int OFF_ID = OFF_AW + AW_BYTES;
  static constexpr int ID_BYTES = BIAS == 1 ? 8192 : 0;
  // 14x14 window mode: the reversed rel_pos tables as a B operand ([64 entries][64 channels] bf16, 128B-swizzled) and a
  // per-row staging area for the table products T = Q x rel^T ([256 rows][68] fp32)
  static constexpr int OFF_REL = OFF_ID + ID_BYTES;
  static constexpr int REL_BYTES = BIAS == 2 ? 8192 : 0;
  static constexpr int OFF_STG = OFF_REL + REL_BYTES;
  static constexpr int STG_BYTES = BIAS == 2 ? 256 * ATT_STG_STRIDE : 0;
  // SPLIT == 2: per-row half-tile maxima [4 tile slots][Q tile][half][128 rows] and row-sum exchange [Q tile][half][128]
  static constexpr int OFF_HM = OFF_STG + STG_BYTES;
  static constexpr int HM_BYTES = SPLIT == 2 ? 4 * 2 * 2 * 128 * 4 : 0;
  static constexpr int OFF_LX = OFF_HM + HM_BYTES;
  static constexpr int LX_BYTES = SPLIT == 2 ? 2 * 2 * 128 * 4 : 0;
  static constexpr int OFF_BAR = OFF_LX + LX_BYTES;
  static constexpr int TOTAL = OFF_BAR + 512 + 1024;
  static_assert(TOTAL <= 232448, "shared memory budget (227 KB per CTA)");
};

template <int KV_TILE, int BIAS, bool TF16 = false, int SPLIT = 1>
__global__ void __launch_bounds__(att_threads(SPLIT), 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                     const __grid_constant__ CUtensorMap tm_rel, const AttParams p) {
  using S = AttSmem<KV_TILE, BIAS, SPLIT>;
  static_assert(SPLIT == 1 || (SPLIT == 2 && KV_TILE == 64 && ATT_TS_OPERANDS),
                "two threads per row: 64-key tiles with two score buffers per Q tile");
  constexpr int ST = S::STAGES;
  // bias(q, k) = rel_w[q][k % GW] + rel_h[q][k / GW]: a KV tile holds NG key-grid rows of GW keys
  constexpr int GW = BIAS == ATT_BIAS_GLOBAL64 ? 64 : (BIAS == ATT_BIAS_WINDOW14 ? 14 : KV_TILE);
  constexpr int NG = KV_TILE / GW;
  static_assert(NG * GW == KV_TILE && KV_TILE % 16 == 0 && KV_TILE <= 128, "tile must hold whole key-grid rows");
  // Tiles of <= 64 keys double-buffer the score tile in TMEM: S(j+1) is issued BEFORE the MMA warp waits for P(j), so
  // the softmax warps never wait for the tensor pipe once the pipeline is full.
  constexpr bool DB = KV_TILE <= 64;
  constexpr bool TSQ = KV_TILE == 64 && ATT_TS_OPERANDS;
  // Score buffers per Q tile in TMEM.  With three (no TMEM-resident A operands) a score tile is issued THREE tiles
  // ahead of its consumer, which takes the P -> PV -> next-S issue latency of the shared tensor pipe (~1200 cycles with
  // both chains queued) off the softmax warps' critical path; two buffers leave room for Q / A_w in tensor memory.
  constexpr uint32_t NBUF = DB ? (TSQ ? 2 : 3) : 1;
  // 64x64 rel-pos mode: rel_w enters the scores through an extra MMA (see the header comment)
  // SPLIT == 2 (ATT_SPLIT_FOLD 0): the softmax threads add rel_w themselves -- with four softmax warps per scheduler they
  // have the issue slots, and the tensor pipe (12 instead of 8 MMAs per Q tile and key tile) is what is short
  constexpr bool FOLD_W = BIAS == ATT_BIAS_GLOBAL64 && (SPLIT == 1 || ATT_SPLIT_FOLD);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* q_full = bars;                         // [Q buffer]
  uint64_t* q_empty = q_full + 2;                  // [Q buffer]: every QK^T of the item has read it
  uint64_t* full_k = q_empty + 2;                  // [stage]
  uint64_t* empty_k = full_k + ST;
  uint64_t* full_v = empty_k + ST;
  uint64_t* empty_v = full_v + ST;
  uint64_t* bar_s = empty_v + ST;                  // [Q tile][score buffer]: S ready
  uint64_t* bar_p = bar_s + 6;                     // [Q tile][score buffer]: P written (one arrival per warp).  Per
                                                   // buffer, because with double buffering a fast warp may finish
                                                   // tile j+1 before a slow one has delivered its rows of tile j.
  uint64_t* bar_pv = bar_p + 6;                    // [Q tile][score buffer]: O += P V of a tile completed.  Per buffer
                                                   // as well: with double buffering a softmax warp may run one tile
                                                   // ahead of its siblings (S(g+1) is issued before P(g) is awaited), so
                                                   // with ONE barrier flipping every tile a wait for PV(g) could be
                                                   // satisfied by the parity of PV(g-2) while PV(g-1) is still pending.
  uint64_t* o_empty = bar_pv + 6;                  // [Q tile]: the epilogue has read O (one arrival per warp)
  uint64_t* aw_full = o_empty + 2;                 // [item parity][Q tile]: rel_w A operand written (64x64 rel-pos mode)
  uint64_t* rel_full = aw_full + 4;                // window mode: rel-pos operand loaded (once)
  uint64_t* bar_t = rel_full + 1;                  // [Q tile] window mode: table product T of the item ready in TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_t + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int NT = (p.seq_len + KV_TILE - 1) / KV_TILE;   // key tiles per item
  const int n_qp = (p.seq_len + 255) >> 8;              // 256-row query pairs per sequence
  const int n_items = p.n_seq * p.n_heads * n_qp;       // item = (seq, head, qpair), qpair fastest
  const bool tr0 = p.trace != nullptr && blockIdx.x == 0 && lane == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    if constexpr (BIAS == ATT_BIAS_WINDOW14) tma_prefetch_desc(&tm_rel);
  }
  if (warp == 1 && lane == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&q_full[b], 1);
      mbar_init(&q_empty[b], 2);   // one commit per Q tile
    }
    for (int s = 0; s < ST; ++s) {
      mbar_init(&full_k[s], 1);
      mbar_init(&empty_k[s], 2);   // the score MMAs of both Q tiles
      mbar_init(&full_v[s], 1);
      mbar_init(&empty_v[s], 2);   // the PV MMAs of both Q tiles
    }
    for (int x = 0; x < 2; ++x) {
      for (int b = 0; b < 3; ++b) {
        mbar_init(&bar_s[3 * x + b], 1);
        mbar_init(&bar_p[3 * x + b], 4 * SPLIT);
        mbar_init(&bar_pv[3 * x + b], 1);
      }
      mbar_init(&o_empty[x], 4 * SPLIT);
      mbar_init(&aw_full[x], 1);
      mbar_init(&aw_full[2 + x], 1);
      mbar_init(&bar_t[x], 1);
    }
    mbar_init(rel_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if constexpr (FOLD_W) {
    if (warp == 2) {
      // B operand of the bias MMA: I[n][k] / scale (64 x 64 fp16, K-major, 128B swizzle: 16-byte chunk c of row n
      // sits at chunk position c ^ (n & 7))
      uint8_t* idm = smem + S::OFF_ID;
      for (int i = lane; i < S::ID_BYTES / 16; i += 32) reinterpret_cast<uint4*>(idm)[i] = make_uint4(0u, 0u, 0u, 0u);
      __syncwarp();
      for (int n = lane; n < 64; n += 32)
        *reinterpret_cast<__half*>(idm + n * 128 + ((((n >> 3) ^ (n & 7))) << 4) + (n & 7) * 2) =
            __float2half_rn(p.inv_scale);
      fence_proxy_async_smem();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384); with double buffering each S region holds
  // two KV_TILE-wide score buffers.  P (bf16) overlays the first KV_TILE/2 columns of the score buffer it came from.
  // Window mode: table products T_A [384,448) T_B [448,512).  64-key modes: the A operands of the score MMAs live in
  // tensor memory too (copied from shared memory once per item with tcgen05.cp): Q_A [384,416) A_w_A [416,448)
  // Q_B [448,480) A_w_B [480,512) -- an M128 x N64 x K16 MMA costs 32 cycles with A in TMEM against 48 with both
  // operands in shared memory (profiles/r01_micro_tcgen05.txt), and the tensor pipe is what bounds the 64x64 mode.
  // With three score buffers per Q tile (no TMEM-resident operands): S_A [0,192) S_B [192,384) O_A [384,448) O_B [448,512).
  const uint32_t TM_S = 0, TM_O = NBUF == 3 ? 384 : 256, TM_T = 384;
  constexpr uint32_t S_SPAN = NBUF == 3 ? 192 : 128;   // TMEM columns between the score regions of the two Q tiles

  if (warp < 4) {
    // ===================================== control warpgroup =====================================
    setmaxnreg_dec<AttRegs<KV_TILE, SPLIT>::CONTROL>();
    if (warp == 0) {
      // ------------------------------------ TMA producer ------------------------------------
      if (lane == 0) {
        if constexpr (BIAS == ATT_BIAS_WINDOW14) {
          mbar_arrive_expect_tx(rel_full, S::REL_BYTES);
          tma_load_2d(smem + S::OFF_REL, &tm_rel, rel_full, 0, 0);
        }
        int it = 0;
        uint32_t g = 0;   // running key-tile counter: ring slot g % ST, phase (g / ST) & 1
        for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
          const int qpair = w % n_qp;
          const int head = (w / n_qp) % p.n_heads;
          const int seq = w / (n_qp * p.n_heads);
          const int row0 = seq * p.seq_len;
          const int qb = it & 1;
          mbar_wait(&q_empty[qb], ((it >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&q_full[qb], S::Q_BYTES);
          tma_load_2d(smem + qb * S::Q_BYTES, &tm_q, &q_full[qb], p.q_off + head * ATT_D, row0 + qpair * 256);
          tma_load_2d(smem + qb * S::Q_BYTES + 16384, &tm_q, &q_full[qb], p.q_off + head * ATT_D,
                      row0 + qpair * 256 + 128);
          for (int j = 0; j < NT; ++j, ++g) {
            const int stage = g % ST;
            const uint32_t phase = (g / ST) & 1;
            const int kv_row = row0 + j * KV_TILE;
            mbar_wait(&empty_k[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_k[stage], S::KV_BYTES);
            tma_load_2d(smem + S::OFF_K + stage * S::KV_SLOT, &tm_kv, &full_k[stage], p.k_off + head * ATT_D, kv_row);
            mbar_wait(&empty_v[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_v[stage], S::KV_BYTES);
            tma_load_2d(smem + S::OFF_V + stage * S::KV_SLOT, &tm_kv, &full_v[stage], p.v_off + head * ATT_D, kv_row);
          }
        }
      }
    } else if (warp == 1 || warp == 3) {
      // ------------------------------------ MMA issuers -------------------------------------
      // One issuer warp per Q tile (warp 1: tile A, warp 3: tile B), each running its own chain: whenever the softmax
      // warps of its tile deliver P(g), it issues  O += P(g) V(g)  and right behind it  S(g + LA)  into the score
      // buffer that PV has just been queued to read (in-order tensor pipe), LA = 2 with double-buffered scores.  So a
      // score tile is always issued a full tile ahead of its consumer, and a slow Q tile never delays the other one.
      // The two chains only meet in the K / V rings, whose slots are released by the commits of both (mbarrier count
      // 2).  Tiles are numbered g = 0, 1, ... across the items of this CTA.
      // The issue path is kept short on purpose: one thread feeds the tensor pipe with 32..48-cycle instructions
      // (M128 x N64 x K16), so descriptor arithmetic between them is what starves it.  All descriptors are
      // base + small offset on pre-shifted 32-bit low words (the high word is a constant), one elect per tile.
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, KV_TILE, 0, 0);  // S = Q K^T   (both K-major, smem)
      // O += P V  (P in TMEM, V MN-major).  ATT_P_F16: P travels as fp16 (A format field 0) against the bf16 V --
      // P <= 2^8 by the lazy rescale, so fp16's 11 significant bits apply (bf16: 8) and the product is exact in fp32
      constexpr uint32_t idesc_o = ATT_P_F16 ? (umma_idesc_bf16(128, ATT_D, 0, 1) & ~(7u << 7))
                                             : umma_idesc_bf16(128, ATT_D, 0, 1);
      // S += A_w I/scale: fp16 operands (A/B format fields 0) -- 11 significant bits for the bias instead of 8
      constexpr uint32_t idesc_w = umma_idesc_bf16(128, KV_TILE, 0, 0) & ~((7u << 7) | (7u << 10));
      constexpr uint32_t LA = NBUF;
      const int x = warp == 1 ? 0 : 1;
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t lo_q = desc_lo(smem_base + x * 16384);                  // + qb * (Q_BYTES >> 4) + 2 * ks
      const uint32_t lo_k = desc_lo(smem_base + S::OFF_K);                   // + slot * (KV_SLOT >> 4) + 2 * ks
      const uint32_t lo_v = desc_lo(smem_base + S::OFF_V);                   // + slot * (KV_SLOT >> 4) + 128 * ks
      const uint32_t lo_aw = desc_lo(smem_base + S::OFF_AW + x * 16384);     // + qb * (32768 >> 4) + 2 * ks
      const uint32_t lo_id = desc_lo(smem_base + S::OFF_ID);                 // + 2 * ks
      const uint32_t lo_rel = desc_lo(smem_base + S::OFF_REL);               // + 2 * ks
      const uint32_t tm_s = tm + TM_S + x * S_SPAN;                          // + buf * KV_TILE
      const uint32_t tm_o = tm + TM_O + x * 64;
      const uint32_t n_my = (static_cast<uint32_t>(n_items) > blockIdx.x)
                                ? (static_cast<uint32_t>(n_items) - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
      const uint32_t total = n_my * static_cast<uint32_t>(NT);

      struct Cursor { uint32_t g, it, j; };   // tile counter, item, tile inside the item
      auto advance = [&](Cursor& c) {
        ++c.g;
        if (++c.j == static_cast<uint32_t>(NT)) { c.j = 0; ++c.it; }
      };
      // (elected thread) S(c) into its score buffer; signals bar_s, releases the K slot (and Q / A_w after the last tile)
      auto mma_s = [&](const Cursor& c) {
        const uint32_t qb = c.it & 1;
        const uint32_t slot = c.g % ST;
        const uint32_t buf = c.g % NBUF;
        const uint32_t d = tm_s + buf * KV_TILE;
        const uint32_t aq = lo_q + qb * (S::Q_BYTES >> 4);
        const uint32_t bk = lo_k + slot * (S::KV_SLOT >> 4);
        if constexpr (BIAS == ATT_BIAS_WINDOW14) {
          // first tile of an item: T = Q x rel^T (the decomposed rel-pos products of the tile's 128 queries with the
          // 2 x 27 table rows) for the softmax warps' prologue; overwrites the previous item's T, which they have
          // consumed before delivering the P tile that triggered this issue
          if (c.j == 0) {
            constexpr uint32_t idesc_t = umma_idesc_bf16(128, 64, 0, 0);
#pragma unroll
            for (int ks = 0; ks < ATT_D / 16; ++ks)
              umma_ss_lo(tm + TM_T + x * 64, aq + 2 * ks, lo_rel + 2 * ks, idesc_t, ks > 0);
            umma_commit(&bar_t[x]);
          }
        }
        if constexpr (TSQ) {
          const uint32_t aw = lo_aw + qb * (32768 >> 4);
          const uint32_t tq = tm + TM_T + x * 64;
          if (c.j == 0) {
            // first tile of an item: Q (and A_w) tiles -> tensor memory, one 128-row x 32-byte K slice per copy.
            // tcgen05.cp and tcgen05.mma execute in issue order, so the copies queue up behind the previous item's
            // score MMAs that still read the old operands, and the shared-memory tiles are free once they are done.
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) tmem_cp_128x256b(tq + ks * 8, aq + 2 * ks);
            if constexpr (FOLD_W) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) tmem_cp_128x256b(tq + 32 + ks * 8, aw + 2 * ks);
            }
            umma_commit(&q_empty[qb]);
          }
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_ts_lo(d, tq + ks * 8, bk + 2 * ks, idesc_s, ks > 0);
          if constexpr (FOLD_W && !ATT_DIAG_NOFOLD) {
            // S += A_w x I / scale : adds rel_w[q, kw] to column kw of every key tile (KV_TILE == 64 == grid width)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_ts_lo(d, tq + 32 + ks * 8, lo_id + 2 * ks, idesc_w, true);
          }
        } else {
#pragma unroll
          for (int ks = 0; ks < ATT_D / 16; ++ks) umma_ss_lo(d, aq + 2 * ks, bk + 2 * ks, idesc_s, ks > 0);
          if constexpr (FOLD_W) {
            const uint32_t aw = lo_aw + qb * (32768 >> 4);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_ss_lo(d, aw + 2 * ks, lo_id + 2 * ks, idesc_w, true);
          }
        }
        umma_commit(&bar_s[3 * x + buf]);
        umma_commit(&empty_k[slot]);
        if constexpr (!TSQ) {
          if (c.j + 1 == static_cast<uint32_t>(NT)) umma_commit(&q_empty[qb]);
        }
      };
      auto wait_s_inputs = [&](const Cursor& c) {
        const int qb = c.it & 1;
        if (c.j == 0) {
          mbar_wait(&q_full[qb], (c.it >> 1) & 1);
          if constexpr (FOLD_W) mbar_wait(&aw_full[qb * 2 + x], (c.it >> 1) & 1);
          if constexpr (BIAS == ATT_BIAS_WINDOW14) {
            if (c.it == 0) mbar_wait(rel_full, 0);
          }
        }
        mbar_wait(&full_k[c.g % ST], (c.g / ST) & 1);
      };

      Cursor sc{0, 0, 0}, pc{0, 0, 0};   // next score tile to issue / next P tile to consume
      for (uint32_t i = 0; i < LA && sc.g < total; ++i) {
        wait_s_inputs(sc);
        tc_fence_after();
        if (elect_one()) mma_s(sc);
        __syncwarp();
        advance(sc);
      }
      while (pc.g < total) {
        const uint32_t buf = pc.g % NBUF;
        const uint32_t vslot = pc.g % ST;
        const bool more = sc.g < total;
        if (more) wait_s_inputs(sc);   // long since there: K is loaded ST tiles ahead
        mbar_wait(&full_v[vslot], (pc.g / ST) & 1);
        if (pc.j == 0 && pc.it > 0) mbar_wait(&o_empty[x], (pc.it - 1) & 1);   // previous item's epilogue has read O
        mbar_wait(&bar_p[3 * x + buf], (pc.g / NBUF) & 1);
        tc_fence_after();
        att_trace(p, tr0, 0, pc.g, 2 * x);
        if (elect_one()) {
          const uint32_t a_p = tm_s + buf * KV_TILE;
          const uint32_t bv = lo_v + vslot * (S::KV_SLOT >> 4);
#pragma unroll
          for (int ks = 0; ks < KV_TILE / 16; ++ks) {
            // SPLIT == 2: each half of the row delivers its 32 keys of P over the first 16 of its OWN 32 score columns
            const uint32_t pa = SPLIT == 2 ? a_p + (ks >> 1) * 32 + (ks & 1) * 8 : a_p + ks * 8;
            umma_ts_lo(tm_o, pa, bv + ks * (2048 >> 4), idesc_o, pc.j > 0 || ks > 0);
          }
          umma_commit(&bar_pv[3 * x + buf]);
          umma_commit(&empty_v[vslot]);
          if (more) mma_s(sc);
        }
        __syncwarp();
        att_trace(p, tr0, 0, pc.g, 2 * x + 1);
        advance(pc);
        if (more) advance(sc);
      }
    } else if constexpr (FOLD_W) {
      // ------------------------------- rel_w operand builder (warp 2, both Q tiles) -------------------------------
      // A_w[r][kw] = fp16(rel_w[q_r, kw]), r = row of the Q tile, read from the fp32 table (entry 63 - qw + kw of
      // the query's row) with coalesced 256-byte requests and written in the 128B-swizzled K-major operand layout.
      int it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const int qpair = w % n_qp;
        const int head = (w / n_qp) % p.n_heads;
        const int seq = w / (n_qp * p.n_heads);
        const long long seq_row0 = static_cast<long long>(seq) * p.seq_len;
        const int qb = it & 1;
        // the buffer was last read by the score MMAs of item it-2 (same completion that frees the Q buffer)
        mbar_wait(&q_empty[qb], ((it >> 1) & 1) ^ 1);
#pragma unroll 1
        for (int x = 0; x < 2; ++x) {
        uint8_t* dst = smem + S::OFF_AW + (qb * 2 + x) * 16384;
        const int t0 = qpair * 256 + x * 128;
#pragma unroll 1
        for (int r0 = 0; r0 < 128; r0 += 16) {
          if constexpr (TF16) {
            // fp16 table: the 64 entries of a row start at a 2-byte aligned offset; fetch the aligned 32-bit words
            // around this lane's pair and funnel them (the shift is the same for the whole row)
            uint32_t w0[16], w1[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const int t = (t0 + r0 + u < p.seq_len) ? t0 + r0 + u : 0;
              const long long e = ((seq_row0 + t) * p.n_heads + head) * p.ldb + (63 - (t & 63));
              const uint32_t* src = reinterpret_cast<const uint32_t*>(static_cast<const __half*>(p.bias_w) + (e & ~1ll)) + lane;
              w0[u] = __ldg(src);
              w1[u] = __ldg(src + 1);
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const int r = r0 + u;
              const int t = (t0 + r < p.seq_len) ? t0 + r : 0;
              const bool odd = ((63 - (t & 63)) & 1) != 0;   // ldb and the view offset are even
              const uint32_t pr = odd ? __byte_perm(w0[u], w1[u], 0x5432) : w0[u];
              *reinterpret_cast<uint32_t*>(dst + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4) = pr;
            }
          } else {
          float v0[16], v1[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int t = (t0 + r0 + u < p.seq_len) ? t0 + r0 + u : 0;
            const float* src = static_cast<const float*>(p.bias_w) + ((seq_row0 + t) * p.n_heads + head) * p.ldb + (63 - (t & 63)) + 2 * lane;
            v0[u] = __ldg(src);
            v1[u] = __ldg(src + 1);
          }
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int r = r0 + u;
            *reinterpret_cast<uint32_t*>(dst + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4) =
                pack_f16(v0[u], v1[u]);
          }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&aw_full[qb * 2 + x]);
        }
      }
    }
  } else if constexpr (SPLIT == 2) {
    // ============================ softmax warps, two threads per query row ============================
    // Warp 4 + 8 * half + 4 * x + quarter owns rows [32 * quarter, +32) of Q tile x (TMEM lane quarter = warp % 4)
    // and score columns [32 * half, +32) of every key tile; it writes its 32 keys of P over the first 16 of its own
    // columns and accumulates the row sum of its half.  The two threads of a row must scale a tile by the SAME
    // running maximum m (they feed one PV product), but m need not be the exact maximum: P is bf16 (8 exponent
    // bits) and O / l are fp32, so any m within ~2^100 of the true one costs no precision.  Protocol:
    //   * every tile each thread publishes the maximum of its half (log2 units, rel_h included) in shared memory,
    //     slot g % 4, BEFORE it arrives on the tile's P barrier;
    //   * tile 0 of an item: exact -- the pair meets at a named barrier and takes the larger half maximum;
    //   * tile j >= 2: candidate = larger half maximum of tile j - 2.  It is visible without any extra
    //     synchronisation: S(j) is only issued once all eight warps have delivered P(j - 2).  m moves to the
    //     candidate when it exceeds m by more than 2^8 (lazy rescale, as before); both threads see the same two
    //     numbers and take the same decision.
    // The half maximum of a tile is off the critical path (nothing in the tile depends on it), so the maximum pass,
    // the exponentials and the P stores of the four warps of a scheduler interleave freely.
    setmaxnreg_inc<AttRegs<KV_TILE, SPLIT>::SOFTMAX>();
    const int sw = warp - 4;
    const int quarter = warp & 3;
    const int x = (sw >> 2) & 1;
    const int half = sw >> 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t t_s0 = tmem_base + lane_addr + TM_S + x * S_SPAN + half * 32;   // my columns of score buffer 0
    const uint32_t t_o = tmem_base + lane_addr + TM_O + x * 64 + half * 32;        // my 32 channels of O
    const int pair_bar = 1 + x * 4 + quarter;                                       // named barrier of the row pair
    float* const hm = reinterpret_cast<float*>(smem + S::OFF_HM);                   // [slot][x][half][row]
    float* const lx = reinterpret_cast<float*>(smem + S::OFF_LX);                   // [x][half][row]
    auto hm_at = [&](uint32_t slot, int hf) -> float* { return hm + (((slot * 2 + x) * 2 + hf) * 128 + r); };
    const float sl2 = p.scale_log2;
    constexpr float LOG2E = 1.4426950408889634f;
    using BT = typename std::conditional<TF16, __half, float>::type;
    auto tab_f32 = [](BT v) -> float {
      if constexpr (TF16) return __half2float(v); else return v;
    };
    auto bh_row_of = [&](int w2) -> const BT* {
      const int qpair2 = w2 % n_qp;
      const int head2 = (w2 / n_qp) % p.n_heads;
      const long long row0 = static_cast<long long>(w2 / (n_qp * p.n_heads)) * p.seq_len;
      const int t2 = qpair2 * 256 + x * 128 + r;
      const int tt = t2 < p.seq_len ? t2 : 0;
      return static_cast<const BT*>(p.bias_h) + ((row0 + tt) * p.n_heads + head2) * p.ldb + (GW - 1 - tt / GW);
    };
    constexpr bool REL64 = BIAS == ATT_BIAS_GLOBAL64;
    constexpr bool RWR = REL64 && !FOLD_W;   // rel_w of my 32 key columns in registers
    const BT* bh_row_pre = nullptr;
    BT rh_pre = BT(0.0f);
    if constexpr (REL64) {
      if (static_cast<int>(blockIdx.x) < n_items) {
        bh_row_pre = bh_row_of(blockIdx.x);
        rh_pre = __ldg(bh_row_pre);
      }
    }
    const bool tr = tr0 && quarter == 0 && half == 0;

    int it = 0;
    uint32_t g0 = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it, g0 += NT) {
      const int qpair = w % n_qp;
      const int head = (w / n_qp) % p.n_heads;
      const int seq = w / (n_qp * p.n_heads);
      const int t = qpair * 256 + x * 128 + r;
      const bool row_valid = t < p.seq_len;
      const BT* bh_row = bh_row_pre;
      BT rh_next = rh_pre;
      // rel_w[q, kw] of my columns kw = 32 * half + i (log2 units): entry (63 - qw + kw) of the query's table row
      float rw[RWR ? 32 : 2];
      if constexpr (RWR) {
        const int tt = row_valid ? t : 0;
#ifdef ATT_DIAG_RW0
        const long long e = (63 - (tt & 63)) + half * 32;   // timing diagnostic (WRONG results): every row reads table row 0
#else
        const long long e = ((static_cast<long long>(seq) * p.seq_len + tt) * p.n_heads + head) * p.ldb + (63 - (tt & 63)) +
                            half * 32;
#endif
        if constexpr (TF16) {
          // the 32 entries start at a 2-byte aligned offset: fetch the aligned 32-bit words around them and funnel
          const uint32_t* src = reinterpret_cast<const uint32_t*>(static_cast<const __half*>(p.bias_w) + (e & ~1ll));
          const bool odd = (e & 1) != 0;
          uint32_t wv[17];
#pragma unroll
          for (int i = 0; i < 17; ++i) wv[i] = __ldg(src + i);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint32_t pr = odd ? __byte_perm(wv[i], wv[i + 1], 0x5432) : wv[i];
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&pr));
            rw[2 * i] = f.x * LOG2E;
            rw[2 * i + 1] = f.y * LOG2E;
          }
        } else {
          const float* src = static_cast<const float*>(p.bias_w) + e;
#pragma unroll
          for (int i = 0; i < 32; ++i) rw[i] = __ldg(src + i) * LOG2E;
        }
      }
      float m_used = 0.0f;
      float l0 = 0.0f, l1 = 0.0f, l2 = 0.0f, l3 = 0.0f;

#pragma unroll 1
      for (int j = 0; j < NT; ++j) {
        const uint32_t g = g0 + j;
        const uint32_t buf = g % NBUF;
        float rh2 = 0.0f;
        if constexpr (REL64) {
          rh2 = tab_f32(rh_next) * LOG2E;
          if (j + 1 < NT) rh_next = __ldg(bh_row + j + 1);
        }
        att_trace(p, tr, 1 + x, g, 0);
        mbar_wait(&bar_s[3 * x + buf], (g / NBUF) & 1);
        tc_fence_after();
        const uint32_t t_s = t_s0 + buf * KV_TILE;
        // My 32 score columns travel in two chunks of 16 (one register set: with rel_w in registers a whole half row
        // does not fit next to it, and a spilled rel_h prefetch stalls every tile for a full global-load latency).
        // P of chunk c goes to columns [8c, 8c + 8) of my region: score columns consumed by then.
        uint32_t sv[16];
        auto load_chunk = [&](const int c) {
          tmem_ld_x16(t_s + 16 * c, sv);
          tmem_ld_wait();
          if constexpr (BIAS == ATT_BIAS_NONE) {
            const int valid = p.seq_len - j * KV_TILE - half * 32 - 16 * c;   // keys of this chunk that exist
            if (valid < 16) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (i >= valid) sv[i] = 0xff800000u;
            }
          }
        };
        att_trace(p, tr, 1 + x, g, 1);

        // ---- the tile's m ----
        float alpha = 1.0f;
        bool need = false;
        if (j == 0) {
          // exact: one extra pass over the scores (once per item)
          float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            load_chunk(c);
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              if constexpr (RWR) {
                float a0, a1;
                ffma2v(a0, a1, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]), sl2, rw[16 * c + i], rw[16 * c + i + 1]);
                if (i & 2) m1 = max3(m1, a0, a1);
                else m0 = max3(m0, a0, a1);
              } else {
                if (i & 2) m1 = max3(m1, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
                else m0 = max3(m0, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
              }
            }
          }
          const float mine = RWR ? fmaxf(m0, m1) + rh2 : fmaf(fmaxf(m0, m1), sl2, rh2);
          *hm_at(g & 3, half) = mine;
          named_bar_sync(pair_bar, 64);
          m_used = fmaxf(mine, *hm_at(g & 3, half ^ 1));
        } else if (j >= 2) {
          const float cand = fmaxf(*hm_at((g - 2) & 3, 0), *hm_at((g - 2) & 3, 1));
          if (cand > m_used + 8.0f || ((LA_ATT_DEBUG_FLAGS & 1) && cand > m_used)) {
            alpha = ex2_approx(m_used - cand);
            m_used = cand;
            need = true;
          }
        }
        att_trace(p, tr, 1 + x, g, 2);
        if (__any_sync(0xffffffffu, need)) {
          // O must hold everything up to the previous tile before it is rescaled (my 32 channels)
          mbar_wait(&bar_pv[3 * x + (g - 1) % NBUF], ((g - 1) / NBUF) & 1);
          tc_fence_after();
#pragma unroll
          for (int hs = 0; hs < 2; ++hs) {
            uint32_t ov[16];
            tmem_ld_x16(t_o + 16 * hs, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            tmem_st_32x32b_x16(t_o + 16 * hs, ov);
          }
          l0 *= alpha;
          l1 *= alpha;
          l2 *= alpha;
          l3 *= alpha;
        }

        // ---- P = exp2(s * scale + rel_h - m) -> bf16 -> TMEM; the half maximum of THIS tile for tile j + 2 ----
        const float off = rh2 - m_used;
        float x0 = -INFINITY, x1 = -INFINITY;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          load_chunk(c);
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float a0, a1;
            if constexpr (RWR) {
              // a = s * scale + rel_w + (rel_h - m); the half maximum is taken on a (it differs from the logit by -m)
              float c0f, c1f;
              fadd2s(c0f, c1f, rw[16 * c + i], rw[16 * c + i + 1], off);
              ffma2v(a0, a1, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]), sl2, c0f, c1f);
              if (i & 2) x1 = max3(x1, a0, a1);
              else x0 = max3(x0, a0, a1);
            } else {
              ffma2(a0, a1, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]), sl2, off);
              if (i & 2) x1 = max3(x1, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
              else x0 = max3(x0, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
            }
            float e0, e1;
            if ((((16 * c + i) >> 1) * ATT_POLY_NUM) % ATT_POLY_DEN < ATT_POLY_NUM) {
              e0 = a0;
              e1 = a1;
              exp2_poly_x2(e0, e1);
            } else if (ATT_DIAG_NOEXP) {
              e0 = a0;
              e1 = a1;
            } else {
              e0 = ex2_approx(a0);
              e1 = ex2_approx(a1);
            }
            if ((i & 2) == 0) fadd2_acc(l0, l1, e0, e1);
            else fadd2_acc(l2, l3, e0, e1);
            pk[i >> 1] = pack_bf16(e0, e1);
          }
          tmem_st_32x32b_x8(t_s + 8 * c, pk);
        }
        *hm_at(g & 3, half) = RWR ? fmaxf(x0, x1) + m_used : fmaf(fmaxf(x0, x1), sl2, rh2);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_p[3 * x + buf]);
        att_trace(p, tr, 1 + x, g, 3);
      }

      if constexpr (REL64) {
        if (w + static_cast<int>(gridDim.x) < n_items) {
          bh_row_pre = bh_row_of(w + gridDim.x);
          rh_pre = __ldg(bh_row_pre);
        }
      }

      // ---- epilogue: the pair adds its row sums, each thread normalises and stores its 32 channels ----
      const float l_mine = (l0 + l1) + (l2 + l3);
      lx[(x * 2 + half) * 128 + r] = l_mine;
      {
        const uint32_t gl = g0 + NT - 1;   // last tile of the item
        mbar_wait(&bar_pv[3 * x + gl % NBUF], (gl / NBUF) & 1);
      }
      tc_fence_after();
      att_trace(p, tr, 3 + x, it, 0);
      uint32_t ov[32];
      tmem_ld_x32(t_o, ov);
      named_bar_sync(pair_bar, 64);
      const float inv_l = 1.0f / (l_mine + lx[(x * 2 + (half ^ 1)) * 128 + r]);
      tmem_ld_wait();
      // O is in registers: hand the accumulator back to the MMA warp before the global stores
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[x]);
      if (row_valid) {
        __nv_bfloat16* dst = p.out + (static_cast<long long>(seq) * p.seq_len + t) * p.ld_out + head * ATT_D + half * 32;
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          uint4 pk4;
          pk4.x = pack_bf16(__uint_as_float(ov[gq * 8 + 0]) * inv_l, __uint_as_float(ov[gq * 8 + 1]) * inv_l);
          pk4.y = pack_bf16(__uint_as_float(ov[gq * 8 + 2]) * inv_l, __uint_as_float(ov[gq * 8 + 3]) * inv_l);
          pk4.z = pack_bf16(__uint_as_float(ov[gq * 8 + 4]) * inv_l, __uint_as_float(ov[gq * 8 + 5]) * inv_l);
          pk4.w = pack_bf16(__uint_as_float(ov[gq * 8 + 6]) * inv_l, __uint_as_float(ov[gq * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + gq * 8) = pk4;
        }
      }
      att_trace(p, tr, 3 + x, it, 1);
    }
  } else {
    // ===================================== softmax warpgroups =====================================
    setmaxnreg_inc<AttRegs<KV_TILE, SPLIT>::SOFTMAX>();
    const int x = (warp - 4) >> 2;     // Q tile: 0 = A, 1 = B
    const int quarter = warp & 3;      // TMEM lane quarter
    const int r = quarter * 32 + lane;  // row inside the Q tile
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t t_s0 = tmem_base + lane_addr + TM_S + x * S_SPAN;
    const uint32_t t_o = tmem_base + lane_addr + TM_O + x * 64;
    const float sl2 = p.scale_log2;
    constexpr float LOG2E = 1.4426950408889634f;
    // bias arithmetic in the softmax threads only for the 14x14 windows (see the header comment)
    constexpr bool WIN = BIAS == ATT_BIAS_WINDOW14;

    // rel_h table row of this thread's query in work item w2 (pointer to the entry of key-grid row 0)
    using BT = typename std::conditional<TF16, __half, float>::type;   // rel-pos table element
    // table entries travel in their storage type and are converted where they are consumed: a conversion placed
    // right behind the load would stall the thread for the full load latency once per key tile
    auto tab_f32 = [](BT v) -> float {
      if constexpr (TF16) return __half2float(v); else return v;
    };
    auto bh_row_of = [&](int w2) -> const BT* {
      const int qpair2 = w2 % n_qp;
      const int head2 = (w2 / n_qp) % p.n_heads;
      const long long row0 = static_cast<long long>(w2 / (n_qp * p.n_heads)) * p.seq_len;
      const int t2 = qpair2 * 256 + x * 128 + r;
      const int tt = t2 < p.seq_len ? t2 : 0;
      return static_cast<const BT*>(p.bias_h) + ((row0 + tt) * p.n_heads + head2) * p.ldb + (GW - 1 - tt / GW);
    };
    // the next item's row pointer and (64x64 mode) its first rel_h term are fetched before the epilogue of the current
    // item, so the global-load latency is off the item-to-item critical path
    const BT* bh_row_pre = nullptr;
    BT rh_pre = BT(0.0f);
    if constexpr (FOLD_W) {
      if (static_cast<int>(blockIdx.x) < n_items) {
        bh_row_pre = bh_row_of(blockIdx.x);
        if constexpr (FOLD_W) rh_pre = __ldg(bh_row_pre);
      }
    }

    // score tile gq of this CTA's tile sequence (TMEM -> registers); the caller awaits it with tmem_ld_wait()
    auto fetch_scores = [&](const uint32_t gq, uint32_t* dst) {
      const uint32_t bq = gq % NBUF;
      mbar_wait(&bar_s[3 * x + bq], (gq / NBUF) & 1);
      tc_fence_after();
      const uint32_t ts = t_s0 + bq * KV_TILE;
#pragma unroll
      for (int c = 0; c + 32 <= KV_TILE; c += 32) tmem_ld_x32(ts + c, dst + c);
      if constexpr (KV_TILE % 32 != 0) tmem_ld_x16(ts + (KV_TILE / 32) * 32, dst + (KV_TILE / 32) * 32);
    };
    constexpr bool PF = DB && (ATT_PIPE != 0);
    // measured (profiles/r02_exp_attention.txt): +14 % on the HF ViT shape (901 tokens, no bias), -4 % on the 64x64
    // rel-pos mode, whose tile period is set by the two exponential passes either way -> 2 = bias-free mode only
    constexpr bool PP = DB && (ATT_PINGPONG == 1 || (ATT_PINGPONG == 2 && BIAS == ATT_BIAS_NONE));
    if constexpr (PP) {
      // tile A goes first: tile B's warp pre-arrives on A's barrier (only if there is work at all)
      if (x == 1 && static_cast<int>(blockIdx.x) < n_items) named_bar_arrive(1 + 2 * quarter, 64);
    }
    uint32_t sva[KV_TILE];
    if constexpr (PF) {
      if (static_cast<int>(blockIdx.x) < n_items) {
        fetch_scores(0, sva);
        tmem_ld_wait();
      }
    }

    int it = 0;
    uint32_t g0 = 0;
    // window mode: one 256-row pair per sequence, so a thread's token and its position in the window never change
    [[maybe_unused]] const int w_t = x * 128 + r;
    [[maybe_unused]] const int w_ty = w_t / GW, w_tx = w_t - (w_t / GW) * GW;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it, g0 += NT) {
      // (an incremental, division-free decode of the item index was measured 5 % SLOWER in the 64x64 mode -- 2.63
      //  against 2.50 ms -- although it removes ~100 instructions per item; the divisions stay)
      const int qpair = WIN ? 0 : w % n_qp;
      const int head = WIN ? w % p.n_heads : (w / n_qp) % p.n_heads;
      const int seq = WIN ? w / p.n_heads : w / (n_qp * p.n_heads);
      const long long seq_row0 = static_cast<long long>(seq) * p.seq_len;
      const int t = WIN ? w_t : qpair * 256 + x * 128 + r;  // token index inside the sequence
      const bool row_valid = t < p.seq_len;
      const bool tr = tr0 && quarter == 0;
      const int tr_role = 1 + x;

      // ---- rel-pos bias prologue (log2 units) ----
      float rw2[WIN ? GW : 1];
      float rh_all[WIN ? 2 * NG : 1];   // window mode: rel_h terms of all 14 key-grid rows (+ 2 zeros), log2 units
      const BT* bh_row = bh_row_pre;
      if constexpr (WIN) {
        // T[row][e] = q_row . rel_rev[e] sits in TMEM (issued with the item's first score tile): entries [0, 27) are
        // the rel_h products, [rel_pad, rel_pad + 27) the rel_w products, and the bias of key (kh, kw) for a query at
        // (qh, qw) is T[13 - qh + kh] + T[rel_pad + 13 - qw + kw].  The shift differs per row, so each thread parks
        // its row in shared memory and reads the 2 x 14 entries it needs back.
        static_assert(!WIN || 2 * NG >= GW, "two key tiles cover the 14 key-grid rows");
        float* srow = reinterpret_cast<float*>(smem + S::OFF_STG + (x * 128 + r) * ATT_STG_STRIDE);
        mbar_wait(&bar_t[x], it & 1);
        tc_fence_after();
        att_trace(p, tr, 3 + x, it, 2);
        {
          uint32_t tb[64];
          tmem_ld_x32(tmem_base + lane_addr + TM_T + x * 64, tb);
          tmem_ld_x32(tmem_base + lane_addr + TM_T + x * 64 + 32, tb + 32);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i)
            reinterpret_cast<uint4*>(srow)[i] = make_uint4(tb[4 * i], tb[4 * i + 1], tb[4 * i + 2], tb[4 * i + 3]);
        }
        __syncwarp();
        const int qh = row_valid ? w_ty : 0, qw = row_valid ? w_tx : 0;
        const float* th = srow + (GW - 1 - qh);
        const float* tw = srow + p.rel_pad + (GW - 1 - qw);
#pragma unroll
        for (int i = 0; i < GW; ++i) rw2[i] = tw[i] * LOG2E;
#pragma unroll
        for (int i = 0; i < 2 * NG; ++i) rh_all[i] = i < GW ? th[i] * LOG2E : 0.0f;
      } else {
        rw2[0] = 0.0f;
        rh_all[0] = 0.0f;
      }

      att_trace(p, tr, 3 + x, it, 3);
      float m_used = -INFINITY;
      float l_sum = 0.0f;
      BT rh_next = BT(0.0f);   // 64x64 mode: rel_h term of the next key tile (= key-grid row), fetched one tile ahead
      if constexpr (FOLD_W) rh_next = rh_pre;

      // one key tile: `sv` holds (PF) or receives (!PF) the scores of tile j; with PF the next tile of this CTA -- of
      // this item or the first one of the next item -- is fetched into `svn` while tile j is being processed
      auto do_tile = [&](const int j, uint32_t* sv, uint32_t* svn) {
        const uint32_t g = g0 + j;
        const int valid = p.seq_len - j * KV_TILE;  // keys of this tile that exist (may exceed KV_TILE)
        // rel_h terms of the NG key-grid rows of this tile
        float rh2[NG];
        if constexpr (FOLD_W) {
          static_assert(!FOLD_W || NG == 1, "one key-grid row per tile");
          rh2[0] = tab_f32(rh_next) * LOG2E;
          if (j + 1 < NT) rh_next = __ldg(bh_row + j + 1);
        } else if constexpr (WIN) {
#pragma unroll
          for (int i = 0; i < NG; ++i) rh2[i] = j == 0 ? rh_all[i] : rh_all[NG + i];
        } else {
          rh2[0] = 0.0f;
        }

        const uint32_t buf = g % NBUF;
        const uint32_t t_s = t_s0 + buf * KV_TILE;
        att_trace(p, tr, tr_role, g, 0);
        [[maybe_unused]] bool has_next = false;
        if constexpr (PF) {
          has_next = (j + 1 < NT) || (w + static_cast<int>(gridDim.x) < n_items);
        } else {
          // ---- the whole score row into registers: ONE pass over TMEM ----
          fetch_scores(g, sv);
          tmem_ld_wait();
        }
        att_trace(p, tr, tr_role, g, 1);
        if constexpr (WIN) {
          // 196 keys = 112 + 84: the second tile's last 28 columns are beyond the window (compile-time positions)
          if (j == 1) {
#pragma unroll
            for (int i = 196 - KV_TILE; i < KV_TILE; ++i) sv[i] = 0xff800000u;
          }
        } else if (valid < KV_TILE) {   // ragged last tile: keys beyond the sequence get -inf
#pragma unroll
          for (int i = 0; i < KV_TILE; ++i)
            if (i >= valid) sv[i] = 0xff800000u;
        }

        // ---- pass 1: tile max in log2 units ----
        float mx;
        if constexpr (WIN) {
          // t = s * scale + rel_w (kept in place), max incl. rel_h
          mx = -INFINITY;
#pragma unroll
          for (int gi = 0; gi < NG; ++gi) {
            static_assert(!WIN || GW % 2 == 0, "score pairs stay inside one key-grid row");
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int i = 0; i < GW; i += 2) {
              const int col = gi * GW + i;
              float t0, t1;
              ffma2v(t0, t1, __uint_as_float(sv[col]), __uint_as_float(sv[col + 1]), sl2, rw2[i], rw2[i + 1]);
              sv[col] = __float_as_uint(t0);
              sv[col + 1] = __float_as_uint(t1);
              if (i & 2) m1 = max3(m1, t0, t1);
              else m0 = max3(m0, t0, t1);
            }
            mx = fmaxf(mx, fmaxf(m0, m1) + rh2[gi]);
          }
        } else {
          // raw scores (rel_w already inside them in the 64x64 mode): max first, scale once (scale > 0)
          static_assert(WIN || KV_TILE % 8 == 0, "max tree works on groups of 8");
          // The running maximum is only a scaling reference: P is bf16 (8 exponent bits) and O / l are fp32, so a
          // stale m costs no precision as long as the scores stay within ~2^100 of it.  ATT_MAX_EVERY > 1 takes the
          // tile maximum on every n-th tile only (always on tile 0).  Measured (profiles/r02_exp_attention.txt): no
          // gain (756 / 765 TF with n = 4 / 8 against 777 with n = 1), so the product takes it on every tile.
          if (j % ATT_MAX_EVERY == 0) {
            float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
            for (int i = 0; i < KV_TILE; i += 8) {
              m0 = max3(m0, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
              m1 = max3(m1, __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
              m2 = max3(m2, __uint_as_float(sv[i + 4]), __uint_as_float(sv[i + 5]));
              m3 = max3(m3, __uint_as_float(sv[i + 6]), __uint_as_float(sv[i + 7]));
            }
            mx = fmaf(fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)), sl2, rh2[0]);
          } else {
            mx = m_used;
          }
        }

        att_trace(p, tr, tr_role, g, 2);
        // ---- running max with lazy rescale (threshold 8 in log2 units => P <= 256) ----
        float alpha = 1.0f;
        bool need = false;
        if (j == 0) {
          m_used = mx;
        } else if (mx > m_used + 8.0f || (LA_ATT_DEBUG_FLAGS & 1)) {
          alpha = ex2_approx(m_used - mx);
          m_used = mx;
          need = true;
        }
        if (__any_sync(0xffffffffu, need)) {
          // O must hold everything up to the previous tile before it is rescaled
          mbar_wait(&bar_pv[3 * x + (g - 1) % NBUF], ((g - 1) / NBUF) & 1);
          tc_fence_after();
#pragma unroll
          for (int hseg = 0; hseg < 2; ++hseg) {
            uint32_t ov[32];
            tmem_ld_x32(t_o + hseg * 32, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            tmem_st_x32(t_o + hseg * 32, ov);
          }
          l_sum *= alpha;
        }

        // ---- pass 2: P = exp2(t - m) -> bf16 pairs -> TMEM (A operand of the PV MMA) over the score columns just
        //      consumed, 32 columns (16 words) at a time so that the stores overlap the remaining exponentials ----
        float offg[NG];
#pragma unroll
        for (int gi = 0; gi < NG; ++gi) offg[gi] = rh2[gi] - m_used;
        float l0 = 0.0f, l1 = 0.0f, l2 = 0.0f, l3 = 0.0f;
        if constexpr (PP) named_bar_sync(1 + 2 * quarter + x, 64);   // my turn on this scheduler's MUFU
#pragma unroll
        for (int c0 = 0; c0 < KV_TILE; c0 += 32) {
          const int width = (KV_TILE - c0 >= 32) ? 32 : 16;
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            if (i < width) {
              const int col = c0 + i;
              float a0, a1;
              if constexpr (WIN) {
                fadd2s(a0, a1, __uint_as_float(sv[col]), __uint_as_float(sv[col + 1]), offg[col / GW]);
              } else {
                ffma2(a0, a1, __uint_as_float(sv[col]), __uint_as_float(sv[col + 1]), sl2, offg[0]);
              }
              float e0, e1;
              constexpr int PNUM = WIN ? ATT_POLY_NUM_WIN : ATT_POLY_NUM;
              if (((col >> 1) * PNUM) % ATT_POLY_DEN < PNUM) {
                e0 = a0;
                e1 = a1;
                exp2_poly_x2(e0, e1);
              } else if (ATT_DIAG_NOEXP) {
                e0 = a0;
                e1 = a1;
              } else {
                e0 = ex2_approx(a0);
                e1 = ex2_approx(a1);
              }
              if ((i & 2) == 0) fadd2_acc(l0, l1, e0, e1);
              else fadd2_acc(l2, l3, e0, e1);
              pk[i >> 1] = ATT_P_F16 ? pack_f16(e0, e1) : pack_bf16(e0, e1);
            }
          }
          if (width == 32) tmem_st_32x32b_x16(t_s + (c0 >> 1), pk);
          else tmem_st_32x32b_x8(t_s + (c0 >> 1), pk);
        }
        if constexpr (PP) {
          // hand the MUFU to the other Q tile's warp (not after this CTA's very last tile: nobody would wait for it)
          if (x == 0 || j + 1 < NT || w + static_cast<int>(gridDim.x) < n_items)
            named_bar_arrive(1 + 2 * quarter + (x ^ 1), 64);
        }
        l_sum += (l0 + l1) + (l2 + l3);
        if constexpr (PF) {
          // S(g+1) was issued when P(g-1) arrived, a whole tile ago: fetch it now, so that its barrier wait and TMEM
          // load latency run under the P store / fence / arrive of this tile instead of opening the next one.
          // (Any earlier and the score tile is not there yet: its buffer held P(g-1) until PV(g-1) had read it.)
          if (has_next) fetch_scores(g + 1, svn);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_p[3 * x + buf]);
        if constexpr (PF) {
          if (has_next) tmem_ld_wait();   // svn is valid from here on
        }
        att_trace(p, tr, tr_role, g, 3);
      };
      // PF: all scores of tile j are consumed when the next tile is fetched, so one register set serves both
      for (int j = 0; j < NT; ++j) do_tile(j, sva, sva);

      if constexpr (FOLD_W) {
        if (w + static_cast<int>(gridDim.x) < n_items) {
          bh_row_pre = bh_row_of(w + gridDim.x);
          if constexpr (FOLD_W) rh_pre = __ldg(bh_row_pre);
        }
      }

      // ---- epilogue: O / l -> bf16 -> global (with the window-unpartition row mapping) ----
      {
        const uint32_t gl = g0 + NT - 1;   // last tile of the item
        mbar_wait(&bar_pv[3 * x + gl % NBUF], (gl / NBUF) & 1);
      }
      tc_fence_after();
      att_trace(p, tr, 3 + x, it, 0);
      long long out_row = -1;
      if (row_valid) {
        if (p.out_mode == 0) {
          out_row = seq_row0 + t;
        } else {
          const int per_img = p.nwin * p.nwin;
          const int img = seq / per_img, wi = seq - img * per_img;
          const int wy = wi / p.nwin;
          const int y = wy * p.win + (WIN ? w_ty : t / p.win);
          const int xx = (wi - wy * p.nwin) * p.win + (WIN ? w_tx : t % p.win);
          if (y < p.img_hw && xx < p.img_hw)
            out_row = (static_cast<long long>(img) * p.img_hw + y) * p.img_hw + xx;
        }
      }
      const float inv_l = 1.0f / l_sum;
      uint32_t ov[64];
      tmem_ld_x32(t_o, ov);
      tmem_ld_x32(t_o + 32, ov + 32);
      tmem_ld_wait();
      // O is in registers: hand the accumulator back to the MMA warp before the global stores
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[x]);
      if (out_row >= 0) {
        __nv_bfloat16* dst = p.out + out_row * p.ld_out + head * ATT_D;
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
#if ATT_EPI_FMUL2
          float o0, o1;
          fmul2s(o0, o1, __uint_as_float(ov[2 * i]), __uint_as_float(ov[2 * i + 1]), inv_l);
          pk[i] = pack_bf16(o0, o1);
#else
          pk[i] = pack_bf16(__uint_as_float(ov[2 * i]) * inv_l, __uint_as_float(ov[2 * i + 1]) * inv_l);
#endif
        }
        if (ATT_EPI_WIDE && p.wide_store) {
#pragma unroll
          for (int gq = 0; gq < 4; ++gq)
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + gq * 16), "r"(pk[8 * gq]),
                         "r"(pk[8 * gq + 1]), "r"(pk[8 * gq + 2]), "r"(pk[8 * gq + 3]), "r"(pk[8 * gq + 4]),
                         "r"(pk[8 * gq + 5]), "r"(pk[8 * gq + 6]), "r"(pk[8 * gq + 7]) : "memory");
        } else {
#pragma unroll
          for (int gq = 0; gq < 8; ++gq)
            *reinterpret_cast<uint4*>(dst + gq * 8) = make_uint4(pk[4 * gq], pk[4 * gq + 1], pk[4 * gq + 2], pk[4 * gq + 3]);
        }
      }
      att_trace(p, tr, 3 + x, it, 1);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int KV_TILE, int BIAS, bool TF16 = false, int SPLIT = 1>
static int launch_attention(cudaStream_t stream, const void* q, long long ld_q, const void* kv, long long ld_kv,
                            const AttParams& p, const void* rel = nullptr) {
  using S = AttSmem<KV_TILE, BIAS, SPLIT>;
  CUtensorMap tm_q, tm_kv, tm_rel;
  int rc = make_tensor_map_2d(&tm_q, q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (uint64_t)ld_q, (uint64_t)p.rows_total,
                              (uint64_t)ld_q * 2, 64, 128, Swizzle::B128);
  if (rc) return rc;
  rc = make_tensor_map_2d(&tm_kv, kv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (uint64_t)ld_kv, (uint64_t)p.rows_total,
                          (uint64_t)ld_kv * 2, 64, KV_TILE, Swizzle::B128);
  if (rc) return rc;
  auto kern = attention_fwd_kernel<KV_TILE, BIAS, TF16, SPLIT>;
  LA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
  const long long items = static_cast<long long>((p.seq_len + 255) / 256) * p.n_heads * p.n_seq;
  const int grid = items < sm_count() ? static_cast<int>(items) : sm_count();
  tm_rel = tm_q;
  if (rel != nullptr) {   // [2 * rel_pad = 64 rows][64 channels] bf16
    rc = make_tensor_map_2d(&tm_rel, rel, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 64, 64, 128, 64, 64, Swizzle::B128);
    if (rc) return rc;
  }
  kern<<<grid, att_threads(SPLIT), S::TOTAL, stream>>>(tm_q, tm_kv, tm_rel, p);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

}  // namespace la

#ifdef LA_ATT_TRACE
static long long* g_att_trace = nullptr;   // experiment builds only
#endif

extern "C" int la_attention_set_trace(void* device_buffer) {
#ifdef LA_ATT_TRACE
  g_att_trace = static_cast<long long*>(device_buffer);
  return LA_OK;
#else
  (void)device_buffer;
  la::set_last_error("la_attention_set_trace: this library was built without -DLA_ATT_TRACE (the product library keeps "
                     "no trace state)");
  return LA_ERR_UNSUPPORTED;
#endif
}

static int attention_dispatch(const char* fn, void* stream, const void* q, long long ld_q, int q_off, const void* kv,
                              long long ld_kv, int k_off, int v_off, long long rows_total, int n_seq, int seq_len,
                              int n_heads, float scale, const void* bias_h, const void* bias_w, int bias_dtype,
                              int ldb, const void* rel_table, int rel_pad, int grid_hw, void* out, long long ld_out,
                              int out_mode, int nwin, int img_hw) {
  using namespace la;
  LA_CHECK_ARG(q && kv && out, "%s: null pointer", fn);
  LA_CHECK_ARG(n_seq > 0 && seq_len > 0 && n_heads > 0, "%s: empty problem", fn);
  LA_CHECK_ARG(scale > 0.0f, "%s: the softmax scale must be positive", fn);
  LA_CHECK_ARG(static_cast<long long>(n_seq) * n_heads * ((seq_len + 255) / 256) < (1ll << 31),
               "%s: too many work items", fn);
  LA_CHECK_ARG(ld_q % 8 == 0 && ld_kv % 8 == 0 && ld_out % 8 == 0 && q_off % 8 == 0 && k_off % 8 == 0 && v_off % 8 == 0,
               "%s: strides/offsets must be multiples of 8 elements", fn);
  LA_CHECK_ARG(rows_total >= static_cast<long long>(n_seq) * seq_len, "%s: rows_total too small", fn);
  LA_CHECK_ARG(rows_total < (1ll << 31), "%s: rows_total exceeds TMA coordinate range", fn);
  const bool has_bias = bias_h != nullptr || bias_w != nullptr;
  LA_CHECK_ARG(!has_bias || (bias_h && bias_w), "%s: bias_h and bias_w must come together", fn);
  AttParams p;
  p.n_seq = n_seq;
  p.seq_len = seq_len;
  p.n_heads = n_heads;
  p.q_off = q_off;
  p.k_off = k_off;
  p.v_off = v_off;
  p.rows_total = rows_total;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.inv_scale = 1.0f / scale;
  p.bias_h = bias_h;
  p.bias_w = bias_w;
  p.ldb = ldb;
  p.rel_pad = rel_pad;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ld_out = ld_out;
  p.out_mode = out_mode;
  p.win = grid_hw;
  p.nwin = nwin;
  p.img_hw = img_hw;
  p.wide_store = (ld_out % 16 == 0 && (reinterpret_cast<uintptr_t>(out) & 31) == 0) ? 1 : 0;
#ifdef LA_ATT_TRACE
  p.trace = g_att_trace;
#else
  p.trace = nullptr;
#endif
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (rel_table != nullptr) {
    LA_CHECK_ARG(grid_hw == 14 && seq_len == 196 && rel_pad == 32,
                 "%s: in-kernel rel-pos tables are built for 14x14 windows (seq_len 196, rel_pad 32)", fn);
    LA_CHECK_ARG((reinterpret_cast<uintptr_t>(rel_table) & 15) == 0, "%s: rel_table must be 16-byte aligned", fn);
    LA_CHECK_ARG(out_mode == 0 || (nwin > 0 && img_hw > 0 && n_seq % (nwin * nwin) == 0),
                 "%s: bad window-unpartition parameters", fn);
    return launch_attention<112, ATT_BIAS_WINDOW14>(st, q, ld_q, kv, ld_kv, p, rel_table);
  }
  if (!has_bias) {
    LA_CHECK_ARG(out_mode == 0, "%s: window output mapping needs the window mode", fn);
    return launch_attention<64, ATT_BIAS_NONE, false, ATT_SPLIT>(st, q, ld_q, kv, ld_kv, p);
  }
  if (grid_hw == 64) {
    LA_CHECK_ARG(seq_len == 4096 && ldb >= 127 && out_mode == 0, "%s: 64x64 rel-pos mode expects seq_len 4096, ldb >= 127",
                 fn);
    LA_CHECK_ARG(bias_dtype == LA_DTYPE_F32 || bias_dtype == LA_DTYPE_F16, "%s: tables are fp32 or fp16", fn);
    if (bias_dtype == LA_DTYPE_F16) {
      LA_CHECK_ARG(ldb % 2 == 0 && (reinterpret_cast<uintptr_t>(bias_w) & 3) == 0 && ldb >= 128,
                   "%s: fp16 tables need an even ldb >= 128 and 4-byte aligned rows", fn);
      return launch_attention<64, ATT_BIAS_GLOBAL64, true, ATT_SPLIT>(st, q, ld_q, kv, ld_kv, p);
    }
    return launch_attention<64, ATT_BIAS_GLOBAL64, false, ATT_SPLIT>(st, q, ld_q, kv, ld_kv, p);
  }
  set_last_error("%s: unsupported rel-pos grid %d (fp32 tables: 64; 14x14 windows go through la_attention_window_bf16)",
                 fn, grid_hw);
  return LA_ERR_UNSUPPORTED;
}

extern "C" int la_attention_bf16(void* stream, const void* q, long long ld_q, int q_off, const void* kv,
                                 long long ld_kv, int k_off, int v_off, long long rows_total, int n_seq,
                                 int seq_len, int n_heads, float scale, const void* bias_h, const void* bias_w,
                                 int bias_dtype, int ldb, int grid_hw, void* out, long long ld_out, int out_mode, int nwin,
                                 int img_hw) {
  return attention_dispatch("la_attention_bf16", stream, q, ld_q, q_off, kv, ld_kv, k_off, v_off, rows_total, n_seq,
                            seq_len, n_heads, scale, bias_h, bias_w, bias_dtype, ldb, nullptr, 0, grid_hw, out, ld_out, out_mode,
                            nwin, img_hw);
}

// second-generation window kernel (la_attention_win.cu): one N = 208 score accumulator per Q tile
int la_attention_window_v2(void* stream, const void* q, long long ld_q, int q_off, const void* kv, long long ld_kv,
                           int k_off, int v_off, long long rows_total, int n_seq, int n_heads, float scale,
                           const void* rel_table, int rel_pad, void* out, long long ld_out, int out_mode, int nwin,
                           int img_hw, int in_pad, long long* trace);
#ifndef LA_WINDOW_V1
#define LA_WINDOW_V1 0      // 1: the first-generation 112-key-tile mode of attention_fwd_kernel (experiment builds)
#endif

extern "C" int la_attention_window_bf16(void* stream, const void* q, long long ld_q, int q_off, const void* kv,
                                        long long ld_kv, int k_off, int v_off, long long rows_total, int n_seq,
                                        int n_heads, float scale, const void* rel_table, int rel_pad, void* out,
                                        long long ld_out, int out_mode, int nwin, int img_hw, int in_pad) {
  using namespace la;
  LA_CHECK_ARG(rel_table != nullptr, "la_attention_window_bf16: rel_table is required");
  if (!LA_WINDOW_V1) {
    LA_CHECK_ARG(q && kv && out && n_seq > 0 && n_heads > 0 && scale > 0.0f, "la_attention_window_bf16: bad arguments");
    LA_CHECK_ARG(ld_q % 8 == 0 && ld_kv % 8 == 0 && ld_out % 8 == 0 && q_off % 8 == 0 && k_off % 8 == 0 && v_off % 8 == 0,
                 "la_attention_window_bf16: strides/offsets must be multiples of 8 elements");
    if (in_pad > 0) {
      LA_CHECK_ARG(nwin > 0 && n_seq % (nwin * nwin) == 0 && in_pad >= nwin * 14 &&
                       rows_total >= static_cast<long long>(n_seq / (nwin * nwin)) * in_pad * in_pad,
                   "la_attention_window_bf16: padded-grid input needs nwin, in_pad >= 14 nwin and "
                   "rows_total >= images * in_pad^2");
    } else {
      LA_CHECK_ARG(rows_total >= static_cast<long long>(n_seq) * 196, "la_attention_window_bf16: rows_total too small");
    }
    LA_CHECK_ARG(rows_total < (1ll << 31), "la_attention_window_bf16: rows_total out of range");
    LA_CHECK_ARG(rel_pad == 32 && (reinterpret_cast<uintptr_t>(rel_table) & 15) == 0,
                 "la_attention_window_bf16: rel_table is the 16-byte aligned [64][64] operand with rel_pad 32");
    LA_CHECK_ARG(out_mode == 0 || (nwin > 0 && img_hw > 0 && n_seq % (nwin * nwin) == 0),
                 "la_attention_window_bf16: bad window-unpartition parameters");
    LA_CHECK_ARG(static_cast<long long>(n_seq) * n_heads < (1ll << 31), "la_attention_window_bf16: too many work items");
#ifdef LA_ATT_TRACE
    long long* const trace = g_att_trace;
#else
    long long* const trace = nullptr;
#endif
    return la_attention_window_v2(stream, q, ld_q, q_off, kv, ld_kv, k_off, v_off, rows_total, n_seq, n_heads, scale,
                                  rel_table, rel_pad, out, ld_out, out_mode, nwin, img_hw, in_pad, trace);
  }
  LA_CHECK_ARG(in_pad == 0, "la_attention_window_bf16: the first-generation kernel reads window-partitioned rows only");
  return attention_dispatch("la_attention_window_bf16", stream, q, ld_q, q_off, kv, ld_kv, k_off, v_off, rows_total,
                            n_seq, 196, n_heads, scale, nullptr, nullptr, LA_DTYPE_F32, 0, rel_table, rel_pad, 14, out, ld_out,
                            out_mode, nwin, img_hw);
}
