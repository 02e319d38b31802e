// labelanything_b200 — tcgen05 GEMM with fused bias / activation epilogue (sm_100a).
//
//   out[M, N] = act( A[M, K] @ W[N, K]^T + bias[N] )          A, W bf16 (K contiguous); fp32 accumulate
//
// This is the one dense-contraction kernel behind every nn.Linear / 1x1 conv / (im2col'd) conv /
// stride==kernel ConvTranspose on the hot path:
//   reference call sites: label_anything/models/image_encoder.py:227-228,242,253 (qkv / proj),
//   common.py:28-37 (MLPBlock), common.py:82-85,108-110,146 (Attention projections),
//   build_lam.py:150-171 (neck), mask_decoder.py:206-255 (upscaling / spatial convs).
//
// Structure (persistent, warp-specialised, one CTA per SM):
//   warp 0   : TMA producer   — A and W tiles (128B-swizzled boxes) into a STAGES-deep smem ring
//   warp 1   : MMA issuer     — one thread issues tcgen05.mma (M=128, N=BLOCK_N, K=16) into TMEM
//   warp 2   : TMEM allocator
//   warps 4-11: epilogue      — tcgen05.ld -> bias/activation -> swizzled smem staging -> TMA store; two warps per
//                               TMEM lane quarter take alternate 128-byte column chunks, so the exact-GELU
//                               epilogue of the MLP GEMM (K = 768: 6144 MMA cycles per tile) stays off the critical path
// The accumulator is double-buffered in TMEM (2 x BLOCK_N columns) so the epilogue of tile i overlaps
// the mainloop of tile i+1.
#include "la_common.cuh"
// experiment switches (compile time, labelanything_b200/build.py::build_variant): the product library reads no
// environment variables on the launch path
#ifndef LA_GEMM_1CTA
#define LA_GEMM_1CTA 0          // 1: never use CTA pairs
#endif
#ifndef LA_GEMM_ACC_MODE
#define LA_GEMM_ACC_MODE 0      // 1 / 2: force the in-place / prefetching residual epilogue
#endif
#include <cstdlib>
#include <type_traits>
#include <cuda_fp16.h>
#include "../../include/labelanything_b200.h"

namespace la {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle atom
constexpr int GEMM_UMMA_K = 16;
constexpr int GEMM_THREADS = 384;

// CTA2: the CTA-pair variant (cta_group::2): a 256 x BLOCK_N tile per cluster of two CTAs, each CTA holding its 128
// rows of A and HALF of the W tile, so a stage is 32 KB instead of 48 KB (6 stages instead of 4) and each MMA reads
// 8 KB instead of 12 KB of operands per CTA.  With one CTA per tile the kernel is shared-memory-bandwidth bound:
// 96 B/clk of operand reads plus 96 B/clk of TMA fills against 128 B/clk.
template <int BLOCK_N, bool CTA2 = false, bool RES2 = false>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  static constexpr int B_ROWS = CTA2 ? BLOCK_N / 2 : BLOCK_N;
  static constexpr int B_BYTES = B_ROWS * GEMM_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = CTA2 ? (RES2 ? 5 : 6) : (BLOCK_N >= 256 ? 4 : (BLOCK_N >= 128 ? 6 : 8));
  static constexpr int EPI_WARP_BYTES = 4096;  // one 32-row x 128-byte staging buffer per epilogue warp
  // RES2: a second buffer per epilogue warp receives the residual chunk one chunk ahead of its use
  static constexpr int EPI_BYTES = (RES2 ? 16 : 8) * EPI_WARP_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
};

// CONV (CTA-pair variant only): implicit-GEMM 3x3 convolution, stride 1, zero padding 1, over a token-major bf16
// feature map [n_img, H, W = 64, C]: A is never materialised -- k-block (tap, channel chunk) of the 128 pixels of a
// CTA (two image rows) is ONE 4-D TMA box [64 ch, 64 w, 2 h, 1 img] fetched at the tap's shifted coordinates, and the
// out-of-bounds rows / columns of the border are zero-filled by the TMA unit.  conv_c = C, conv_h = H; K = 9 * C.
// two fp32 -> one packed 16-bit pair of the output type
template <typename OutT>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  if constexpr (std::is_same<OutT, __half>::value) {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  } else {
    return pack_bf16(lo, hi);
  }
}

// RESID (fp32 output only): out += A W^T + bias, i.e. the output tensor is also the residual operand.  Each epilogue
// warp TMA-loads its 32 x 32 chunk of `out` into the staging buffer it will store from, adds the accumulator in place
// and stores it back -- the residual stream's read and write ride under the tensor-core mainloop of the next tile.
// RESID == 2: the residual chunk is fetched into a buffer of its own, ONE CHUNK AHEAD (also across tiles), so the
// epilogue never waits for a load -- what the short-K proj GEMM needs (its tile lasts 3 us, four exposed L2/HBM
// round trips per tile would make it epilogue bound).
// SPLITK (one-CTA variant, fp32 output, no activation): the work items are (output tile, K split); every CTA contracts
// its range of k-blocks and ADDS its partial tile to the zeroed output with TMA reduce stores (split 0 adds the bias).
// The weight-gradient GEMMs of the training step have one or two output tiles and K = 54 000 image-token rows: without
// the split ONE CTA walks 844 k-blocks (330 us per launch, a quarter of the whole step).
template <int BLOCK_N, typename OutT, bool CTA2 = false, bool CONV = false, int RESID = 0, bool SPLITK = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
                         const __grid_constant__ CUtensorMap tm_out, const float* __restrict__ bias, int M, int N,
                         int K, int act, int conv_c = 0, int conv_h = 0, int out_grid = 0, int splits = 1) {
  static_assert(!SPLITK || (!CTA2 && !CONV && RESID == 0 && sizeof(OutT) == 4), "split-K: one-CTA fp32 kernel only");
  // out_grid > 0: the output rows are the pixels (image, y, x) of a square out_grid x out_grid grid and `tm_out` is a
  // 4-D map (channel, x, y, image) of a LARGER (padded) grid: every 32-row store box lies inside one grid row
  static_assert(!CONV || CTA2, "the implicit-GEMM convolution is built on the CTA-pair kernel");
  static_assert(RESID == 0 || sizeof(OutT) == 4, "the residual epilogue accumulates into an fp32 tensor");
  using S = GemmSmem<BLOCK_N, CTA2, RESID == 2>;
  // CTA pair: rank 0 (leader) issues the MMAs of both; every CTA loads and stores its own 128 rows
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0;
  constexpr int TILE_M = CTA2 ? 2 * GEMM_BLOCK_M : GEMM_BLOCK_M;
  const int tile0 = CTA2 ? blockIdx.x >> 1 : blockIdx.x;
  const int tile_step = CTA2 ? gridDim.x >> 1 : gridDim.x;
  constexpr int CHUNK = 128 / (int)sizeof(OutT);  // output columns per staging row (128 bytes)
  constexpr int NCHUNK = BLOCK_N / CHUNK;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + S::STAGES * S::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + S::EPI_BYTES);
  uint64_t* empty_bar = full_bar + S::STAGES;
  uint64_t* tmem_full = empty_bar + S::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* resid_bar = tmem_empty + 2;   // [epilogue warp] RESID: the warp's chunk of `out` has landed in its staging buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(resid_bar + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int n_blks = (N + BLOCK_N - 1) / BLOCK_N;
  const int m_blks = (M + TILE_M - 1) / TILE_M;
  const int num_tiles = n_blks * m_blks;
  const int k_blks = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  const int act_code = act;
  const int num_work = SPLITK ? num_tiles * splits : num_tiles;     // work item = split * num_tiles + tile
  auto kb_first = [&](int sp) { return SPLITK ? static_cast<int>(static_cast<long long>(sp) * k_blks / splits) : 0; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_out);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], CTA2 ? 16 : 8);   // the epilogue warps of both CTAs release the leader's accumulator
    }
    for (int w8 = 0; w8 < 8; ++w8) mbar_init(&resid_bar[w8], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (CTA2) {
      tmem_alloc_2cta(tmem_slot, S::TMEM_COLS);
      tmem_relinquish_2cta();
    } else {
      tmem_alloc(tmem_slot, S::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) cluster_sync_all();   // the peer's barriers exist before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int wi = tile0; wi < num_work; wi += tile_step) {
        const int tile = SPLITK ? wi % num_tiles : wi;
        const int kb_lo = kb_first(SPLITK ? wi / num_tiles : 0);
        const int kb_n = SPLITK ? kb_first(wi / num_tiles + 1) - kb_lo : k_blks;
        const int m0 = (tile / n_blks) * TILE_M + rank * GEMM_BLOCK_M;
        const int n0 = (tile % n_blks) * BLOCK_N + rank * S::B_ROWS * (CTA2 ? 1 : 0);
        for (int kb = 0; kb < kb_n; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = smem + stage * S::STAGE_BYTES;
          uint8_t* b_dst = a_dst + S::A_BYTES;
          if constexpr (CTA2) {
            // both CTAs' boxes are counted on the leader's barrier
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * S::STAGE_BYTES);
            if constexpr (CONV) {
              const int cpb = conv_c / GEMM_BLOCK_K;            // channel chunks per tap
              const int tap = kb / cpb, c0 = (kb - tap * cpb) * GEMM_BLOCK_K;
              const int img = m0 / (conv_h * 64), y0 = (m0 - img * conv_h * 64) / 64;
              tma_load_4d_2cta(a_dst, &tm_a, &full_bar[stage], c0, tap % 3 - 1, y0 + tap / 3 - 1, img);
            } else {
              tma_load_2d_2cta(a_dst, &tm_a, &full_bar[stage], kb * GEMM_BLOCK_K, m0);
            }
            tma_load_2d_2cta(b_dst, &tm_w, &full_bar[stage], kb * GEMM_BLOCK_K, n0);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], S::STAGE_BYTES);
            tma_load_2d(a_dst, &tm_a, &full_bar[stage], (kb_lo + kb) * GEMM_BLOCK_K, m0);
            tma_load_2d(b_dst, &tm_w, &full_bar[stage], (kb_lo + kb) * GEMM_BLOCK_K, n0);
          }
          if (++stage == S::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    // The whole warp walks the loop converged so that descriptors / addresses stay in uniform registers; one
    // elected lane issues the tcgen05 instructions of each k-block.  CTA pair: only the leader issues.
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, BLOCK_N);
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int wi = tile0; wi < num_work; wi += tile_step, ++it) {
        const int kb_n = SPLITK ? kb_first(wi / num_tiles + 1) - kb_first(wi / num_tiles) : k_blks;
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tm + acc * BLOCK_N;
        for (int kb = 0; kb < kb_n; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * S::STAGE_BYTES);
          const uint32_t b_base = a_base + S::A_BYTES;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < GEMM_BLOCK_K / GEMM_UMMA_K; ++k) {
              const uint64_t a_desc = umma_smem_desc_sw128(a_base + k * GEMM_UMMA_K * 2);
              const uint64_t b_desc = umma_smem_desc_sw128(b_base + k * GEMM_UMMA_K * 2);
              if constexpr (CTA2) umma_bf16_ss_2cta(d_tmem, a_desc, b_desc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
              else umma_bf16_ss(d_tmem, a_desc, b_desc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            if constexpr (CTA2) {
              umma_commit_2cta_mc(&empty_bar[stage], 3);                         // frees the slot in both CTAs
              if (kb == kb_n - 1) umma_commit_2cta_mc(&tmem_full[acc], 3);       // both epilogues
            } else {
              umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
              if (kb == kb_n - 1) umma_commit(&tmem_full[acc]);    // accumulator complete -> epilogue
            }
          }
          __syncwarp();
          if (++stage == S::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------- epilogue -----------------------------------
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;  // which of the quarter's two warps: takes chunks half, half + 2, ...
    uint8_t* buf = epi_smem + (half * 4 + q) * S::EPI_WARP_BYTES;
    constexpr int LAST0 = ((NCHUNK - 1) / 2) * 2;                 // last chunk of the even warp
    constexpr int LAST1 = NCHUNK >= 2 ? ((NCHUNK - 2) / 2) * 2 + 1 : -1;  // last chunk of the odd warp (none if 1 chunk)
    const int last_c = half == 0 ? LAST0 : LAST1;
    [[maybe_unused]] uint32_t rphase = 0;
    [[maybe_unused]] uint8_t* rbuf = epi_smem + (8 + half * 4 + q) * S::EPI_WARP_BYTES;   // RESID == 2 only
    [[maybe_unused]] uint64_t* rbar2 = &resid_bar[half * 4 + q];
    // RESID == 2: fetch the 32 x CHUNK piece of `out` at (tile t, chunk c) into rbuf (lane 0)
    auto prefetch_resid = [&](int t, int c) {
      const int pm0 = (t / n_blks) * TILE_M + rank * GEMM_BLOCK_M;
      const int pn0 = (t % n_blks) * BLOCK_N;
      mbar_arrive_expect_tx(rbar2, S::EPI_WARP_BYTES);
      tma_load_2d(rbuf, &tm_out, rbar2, pn0 + c * CHUNK, pm0 + q * 32);
    };
    if constexpr (RESID == 2) {
      if (lane == 0 && half < NCHUNK && tile0 < num_tiles) prefetch_resid(tile0, half);
    }
    int it = 0;
    for (int wi = tile0; wi < num_work; wi += tile_step, ++it) {
      const int tile = SPLITK ? wi % num_tiles : wi;
      const bool add_bias = bias != nullptr && (!SPLITK || wi < num_tiles);     // split 0 carries the bias
      const int m0 = (tile / n_blks) * TILE_M + rank * GEMM_BLOCK_M;
      const int n0 = (tile % n_blks) * BLOCK_N;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      [[maybe_unused]] uint64_t* rbar = &resid_bar[half * 4 + q];
      auto load_resid = [&](int c) {   // lane 0: this warp's 32 x CHUNK piece of `out` -> staging buffer
        tma_store_wait_read<0>();      // the previous store from the buffer has read it
        mbar_arrive_expect_tx(rbar, S::EPI_WARP_BYTES);
        tma_load_2d(buf, &tm_out, rbar, n0 + c * CHUNK, m0 + q * 32);
      };
      if constexpr (RESID == 1) {
        if (lane == 0 && half < NCHUNK) load_resid(half);
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N;
      if (last_c < 0) {   // nothing to read for this warp: release its share of the accumulator at once
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CTA2) mbar_arrive_leader(&tmem_empty[acc]);
          else mbar_arrive(&tmem_empty[acc]);
        }
      }
#pragma unroll 1
      for (int c = half; c < NCHUNK; c += 2) {
        const int col0 = c * CHUNK;
        float v[CHUNK];
        {
          uint32_t r[CHUNK];
#pragma unroll
          for (int j = 0; j < CHUNK / 32; ++j) tmem_ld_x32(t_row + col0 + j * 32, r + j * 32);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < CHUNK; ++i) v[i] = __uint_as_float(r[i]);
        }
        if (c == last_c) {
          // all TMEM reads of this warp for this accumulator are done -> hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
          if constexpr (CTA2) mbar_arrive_leader(&tmem_empty[acc]);
          else mbar_arrive(&tmem_empty[acc]);
        }
        }
        if (add_bias) {
#pragma unroll
          for (int i = 0; i < CHUNK; i += 4) {
            const int col = n0 + col0 + i;
            if (col + 3 < N) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col));
              v[i] += b.x;
              v[i + 1] += b.y;
              v[i + 2] += b.z;
              v[i + 3] += b.w;
            }
          }
        }
        if (act_code == LA_ACT_GELU) {
#pragma unroll
          for (int i = 0; i < CHUNK; i += 2) {
            if constexpr (sizeof(OutT) == 2) gelu_tanh_erf_x2(v[i], v[i + 1]);   // error far below the bf16 rounding
            else gelu_erf_x2(v[i], v[i + 1]);
          }
        } else if (act_code == LA_ACT_RELU) {
#pragma unroll
          for (int i = 0; i < CHUNK; ++i) v[i] = fmaxf(v[i], 0.0f);
        }
        uint8_t* row_ptr = buf + lane * 128;
        if constexpr (RESID == 2) {
          // the chunk arrived in rbuf (requested one chunk ago): add it, hand rbuf to the next request, then wait for
          // the store buffer as in the plain path
          mbar_wait(rbar2, rphase);
          rphase ^= 1;
          const uint8_t* rrow = rbuf + lane * 128;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 xr = *reinterpret_cast<const float4*>(rrow + ((g ^ (lane & 7)) << 4));
            v[g * 4 + 0] += xr.x;
            v[g * 4 + 1] += xr.y;
            v[g * 4 + 2] += xr.z;
            v[g * 4 + 3] += xr.w;
          }
          __syncwarp();
          if (lane == 0) {
            if (c + 2 < NCHUNK) prefetch_resid(tile, c + 2);
            else if (tile + tile_step < num_tiles) prefetch_resid(tile + tile_step, half);
            tma_store_wait_read<0>();
          }
          __syncwarp();
        } else if constexpr (RESID == 1) {
          // out chunk (loaded by TMA, 128B-swizzled like the store layout) + accumulator, in place
          mbar_wait(rbar, rphase);
          rphase ^= 1;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 xr = *reinterpret_cast<const float4*>(row_ptr + ((g ^ (lane & 7)) << 4));
            v[g * 4 + 0] += xr.x;
            v[g * 4 + 1] += xr.y;
            v[g * 4 + 2] += xr.z;
            v[g * 4 + 3] += xr.w;
          }
        } else {
          // staging buffer: make sure the TMA store of this warp's previous chunk has read it
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
        }
        if constexpr (sizeof(OutT) == 2) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            uint4 pk;
            pk.x = pack2<OutT>(v[g * 8 + 0], v[g * 8 + 1]);
            pk.y = pack2<OutT>(v[g * 8 + 2], v[g * 8 + 3]);
            pk.z = pack2<OutT>(v[g * 8 + 4], v[g * 8 + 5]);
            pk.w = pack2<OutT>(v[g * 8 + 6], v[g * 8 + 7]);
            *reinterpret_cast<uint4*>(row_ptr + ((g ^ (lane & 7)) << 4)) = pk;
          }
        } else {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float4 pk = make_float4(v[g * 4 + 0], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
            *reinterpret_cast<float4*>(row_ptr + ((g ^ (lane & 7)) << 4)) = pk;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (out_grid > 0) {
            const int row = m0 + q * 32;
            const int img = row / (out_grid * out_grid);
            const int rem = row - img * out_grid * out_grid;
            const int gy = rem / out_grid;
            tma_store_4d(&tm_out, buf, n0 + col0, rem - gy * out_grid, gy, img);
          } else if constexpr (SPLITK) {
            tma_reduce_add_2d(&tm_out, buf, n0 + col0, m0 + q * 32);
          } else {
            tma_store_2d(&tm_out, buf, n0 + col0, m0 + q * 32);
          }
          tma_store_commit();
          if constexpr (RESID == 1) {
            if (c + 2 < NCHUNK) load_resid(c + 2);   // next chunk of this tile (after the store has read the buffer)
          }
        }
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) cluster_sync_all();   // neither CTA leaves (or frees tensor memory) while its peer still works
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CTA2) tmem_dealloc_2cta(tmem_base, S::TMEM_COLS);
    else tmem_dealloc(tmem_base, S::TMEM_COLS);
  }
}

// CTA-pair launch: clusters of 2 CTAs, 256 x 256 tiles
template <typename OutT, bool CONV = false, int RESID = 0>
static int launch_gemm_2cta(cudaStream_t stream, const CUtensorMap& tm_a, const CUtensorMap& tm_w,
                            const CUtensorMap& tm_out, const float* bias, int M, int N, int K, int act,
                            int conv_c = 0, int conv_h = 0, int out_grid = 0) {
  using S = GemmSmem<256, true, RESID == 2>;
  auto kern = gemm_bf16_tcgen05_kernel<256, OutT, true, CONV, RESID>;
  LA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
  const int tiles = ((N + 255) / 256) * ((M + 255) / 256);
  const int pairs = sm_count() / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * (tiles < pairs ? tiles : pairs));
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = S::TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LA_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tm_a, tm_w, tm_out, bias, M, N, K, act, conv_c, conv_h, out_grid, 1));
  return LA_OK;
}

template <int BLOCK_N, typename OutT>
static int launch_gemm(cudaStream_t stream, const CUtensorMap& tm_a, const CUtensorMap& tm_w,
                       const CUtensorMap& tm_out, const float* bias, int M, int N, int K, int act, int out_grid = 0) {
  using S = GemmSmem<BLOCK_N>;
  auto kern = gemm_bf16_tcgen05_kernel<BLOCK_N, OutT>;
  LA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
  const int n_blks = (N + BLOCK_N - 1) / BLOCK_N;
  const int m_blks = (M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  const int tiles = n_blks * m_blks;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  kern<<<grid, GEMM_THREADS, S::TOTAL, stream>>>(tm_a, tm_w, tm_out, bias, M, N, K, act, 0, 0, out_grid, 1);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

// split-K launch (one-CTA kernel, fp32 output zeroed here): tiles x splits work items over the SMs
template <int BLOCK_N>
static int launch_gemm_splitk(cudaStream_t stream, const CUtensorMap& tm_a, const CUtensorMap& tm_w,
                              const CUtensorMap& tm_out, const float* bias, int M, int N, int K, int splits) {
  using S = GemmSmem<BLOCK_N>;
  auto kern = gemm_bf16_tcgen05_kernel<BLOCK_N, float, false, false, 0, true>;
  LA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
  const int tiles = ((N + BLOCK_N - 1) / BLOCK_N) * ((M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M);
  const int work = tiles * splits;
  const int grid = work < sm_count() ? work : sm_count();
  kern<<<grid, GEMM_THREADS, S::TOTAL, stream>>>(tm_a, tm_w, tm_out, bias, M, N, K, LA_ACT_NONE, 0, 0, 0, splits);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

}  // namespace la

extern "C" int la_gemm_bf16_splitk(void* stream, const void* a, long long lda, const void* w, long long ldw,
                                   const float* bias, float* out, long long ldo, int M, int N, int K) {
  using namespace la;
  LA_CHECK_ARG(a && w && out && M > 0 && N > 0 && K > 0, "la_gemm_bf16_splitk: bad arguments");
  LA_CHECK_ARG(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && N % 8 == 0 && ldo % 4 == 0 && ldo >= N,
               "la_gemm_bf16_splitk: K / lda / ldw / N must be multiples of 8, ldo of 4");
  LA_CHECK_ARG((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(out)) %
                       16 == 0, "la_gemm_bf16_splitk: pointers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int block_n = N >= 256 ? 256 : (N > 64 ? 128 : 64);
  const int tiles = ((N + block_n - 1) / block_n) * ((M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M);
  const int k_blks = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  // enough splits to fill the SMs once, at least 8 k-blocks per split
  int splits = (sm_count() + tiles - 1) / tiles;
  if (splits > k_blks / 8) splits = k_blks / 8;
  if (splits < 1) splits = 1;
  LA_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * static_cast<size_t>(M - 1) * ldo + sizeof(float) * N, st));
  CUtensorMap tm_a, tm_w, tm_out;
  int rc = make_tensor_map_2d(&tm_a, a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2,
                              GEMM_BLOCK_K, GEMM_BLOCK_M, Swizzle::B128);
  if (rc) return rc;
  rc = make_tensor_map_2d(&tm_w, w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (uint64_t)K, (uint64_t)N, (uint64_t)ldw * 2,
                          GEMM_BLOCK_K, (uint32_t)block_n, Swizzle::B128);
  if (rc) return rc;
  rc = make_tensor_map_2d(&tm_out, out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (uint64_t)N, (uint64_t)M, (uint64_t)ldo * 4, 32,
                          32, Swizzle::B128);
  if (rc) return rc;
  if (block_n == 256) return launch_gemm_splitk<256>(st, tm_a, tm_w, tm_out, bias, M, N, K, splits);
  if (block_n == 128) return launch_gemm_splitk<128>(st, tm_a, tm_w, tm_out, bias, M, N, K, splits);
  return launch_gemm_splitk<64>(st, tm_a, tm_w, tm_out, bias, M, N, K, splits);
}

extern "C" int la_gemm_bf16(void* stream, const void* a, long long lda, const void* w, long long ldw,
                            const float* bias, void* out, long long ldo, int out_dtype, int M, int N, int K,
                            int act) {
  using namespace la;
  LA_CHECK_ARG(a && w && out, "la_gemm_bf16: null pointer");
  LA_CHECK_ARG(M > 0 && N > 0 && K > 0, "la_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
  LA_CHECK_ARG(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "la_gemm_bf16: K/lda/ldw must be multiples of 8");
  LA_CHECK_ARG(N % 8 == 0, "la_gemm_bf16: N must be a multiple of 8 (got %d)", N);
  LA_CHECK_ARG(out_dtype == LA_DTYPE_BF16 || out_dtype == LA_DTYPE_F32 || out_dtype == LA_DTYPE_F16,
               "la_gemm_bf16: bad out_dtype %d", out_dtype);
  LA_CHECK_ARG((ldo * (out_dtype == LA_DTYPE_F32 ? 4 : 2)) % 16 == 0, "la_gemm_bf16: ldo must give 16B rows");
  LA_CHECK_ARG((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(out)) %
                       16 ==
                   0,
               "la_gemm_bf16: pointers must be 16-byte aligned");
  LA_CHECK_ARG(act >= LA_ACT_NONE && act <= LA_ACT_RELU, "la_gemm_bf16: bad act %d", act);
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  const int block_n = N >= 256 ? 256 : (N > 64 ? 128 : 64);
  // big problems run on CTA pairs (W boxes of 128 rows: each CTA loads half of the 256-wide tile)
  const bool pair = block_n == 256 && M >= 2048 && !LA_GEMM_1CTA;
  CUtensorMap tm_a, tm_w, tm_out;
  int rc = make_tensor_map_2d(&tm_a, a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (uint64_t)K, (uint64_t)M,
                              (uint64_t)lda * 2, GEMM_BLOCK_K, GEMM_BLOCK_M, Swizzle::B128);
  if (rc) return rc;
  rc = make_tensor_map_2d(&tm_w, w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (uint64_t)K, (uint64_t)N, (uint64_t)ldw * 2,
                          GEMM_BLOCK_K, (uint32_t)(pair ? 128 : block_n), Swizzle::B128);
  if (rc) return rc;
  if (out_dtype == LA_DTYPE_F16) {   // rel-pos tables (no activation: the GELU forms are tuned for bf16 / fp32)
    LA_CHECK_ARG(act == LA_ACT_NONE, "la_gemm_bf16: fp16 output has no activation epilogue");
    rc = make_tensor_map_2d(&tm_out, out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (uint64_t)N, (uint64_t)M, (uint64_t)ldo * 2,
                            64, 32, Swizzle::B128);
    if (rc) return rc;
    if (pair) return launch_gemm_2cta<__half>(st, tm_a, tm_w, tm_out, bias, M, N, K, act);
    if (block_n == 256) return launch_gemm<256, __half>(st, tm_a, tm_w, tm_out, bias, M, N, K, act);
    if (block_n == 128) return launch_gemm<128, __half>(st, tm_a, tm_w, tm_out, bias, M, N, K, act);
    return launch_gemm<64, __half>(st, tm_a, tm_w, tm_out, bias, M, N, K, act);
  } else if (out_dtype == LA_DTYPE_BF16) {
    rc = make_tensor_map_2d(&tm_out, out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (uint64_t)N, (uint64_t)M,
                            (uint64_t)ldo * 2, 64, 32, Swizzle::B128);
    if (rc) return rc;
    if (pair) return launch_gemm_2cta<__nv_bfloat16>(st, tm_a, tm_w, tm_out, bias, M, N, K, act);
    if (block_n == 256) return launch_gemm<256, __nv_bfloat16>(st, tm_a, tm_w, tm_out, bias, M, N, K, act);
    if (block_n == 128) return launch_gemm<128, __nv_bfloat16>(st, tm_a, tm_w, tm_out, bias, M, N, K, act);
    return launch_gemm<64, __nv_bfloat16>(st, tm_a, tm_w, tm_out, bias, M, N, K, act);
  } else {
    rc = make_tensor_map_2d(&tm_out, out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (uint64_t)N, (uint64_t)M,
                            (uint64_t)ldo * 4, 32, 32, Swizzle::B128);
    if (rc) return rc;
    if (pair) return launch_gemm_2cta<float>(st, tm_a, tm_w, tm_out, bias, M, N, K, act);
    if (block_n == 256) return launch_gemm<256, float>(st, tm_a, tm_w, tm_out, bias, M, N, K, act);
    if (block_n == 128) return launch_gemm<128, float>(st, tm_a, tm_w, tm_out, bias, M, N, K, act);
    return launch_gemm<64, float>(st, tm_a, tm_w, tm_out, bias, M, N, K, act);
  }
}

namespace la {
// positions of the padded grid outside the projected grid x grid part <- the bias row (what a zero input row projects to)
__global__ void __launch_bounds__(256)
fill_grid_padding_kernel(__nv_bfloat16* __restrict__ out, long long ldo, const float* __restrict__ bias, int n_img,
                         int grid, int padded, int n8) {
  const int per_img = padded * padded - grid * grid;     // padding positions per image
  const long long total = static_cast<long long>(n_img) * per_img * n8;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % n8);
    const long long r = i / n8;
    const int k = static_cast<int>(r % per_img);
    const long long img = r / per_img;
    // k enumerates the right strip (grid rows x (padded - grid) columns), then the bottom strip (full rows)
    const int strip = grid * (padded - grid);
    int y, x;
    if (k < strip) {
      y = k / (padded - grid);
      x = grid + k - y * (padded - grid);
    } else {
      const int k2 = k - strip;
      y = grid + k2 / padded;
      x = k2 - (k2 / padded) * padded;
    }
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (bias != nullptr) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias) + 2 * c8);
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias) + 2 * c8 + 1);
      v = make_uint4(pack_bf16(b0.x, b0.y), pack_bf16(b0.z, b0.w), pack_bf16(b1.x, b1.y), pack_bf16(b1.z, b1.w));
    }
    reinterpret_cast<uint4*>(out + ((img * padded + y) * padded + x) * ldo)[c8] = v;
  }
}
}  // namespace la

extern "C" int la_gemm_bf16_to_grid(void* stream, const void* a, long long lda, const void* w, long long ldw,
                                    const float* bias, void* out, long long ldo, int M, int N, int K, int grid,
                                    int padded) {
  using namespace la;
  LA_CHECK_ARG(a && w && out, "la_gemm_bf16_to_grid: null pointer");
  LA_CHECK_ARG(M > 0 && N > 0 && K > 0, "la_gemm_bf16_to_grid: empty problem M=%d N=%d K=%d", M, N, K);
  LA_CHECK_ARG(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && N % 8 == 0 && ldo % 8 == 0 && ldo >= N,
               "la_gemm_bf16_to_grid: K/lda/ldw/N/ldo must be multiples of 8");
  LA_CHECK_ARG(grid > 0 && grid % 32 == 0 && padded >= grid && M % (grid * grid) == 0,
               "la_gemm_bf16_to_grid: grid must be a multiple of 32 (32-row store boxes stay inside a grid row), "
               "padded >= grid and M a whole number of grid x grid images");
  LA_CHECK_ARG((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(out)) % 16 == 0,
               "la_gemm_bf16_to_grid: pointers must be 16-byte aligned");
  LA_CHECK_ARG(!bias || (reinterpret_cast<uintptr_t>(bias) % 16 == 0), "la_gemm_bf16_to_grid: bias must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n_img = M / (grid * grid);
  const int block_n = N >= 256 ? 256 : (N > 64 ? 128 : 64);
  const bool pair = block_n == 256 && M >= 2048 && !LA_GEMM_1CTA;
  CUtensorMap tm_a, tm_w, tm_out;
  int rc = make_tensor_map_2d(&tm_a, a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (uint64_t)K, (uint64_t)M,
                              (uint64_t)lda * 2, GEMM_BLOCK_K, GEMM_BLOCK_M, Swizzle::B128);
  if (rc) return rc;
  rc = make_tensor_map_2d(&tm_w, w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (uint64_t)K, (uint64_t)N, (uint64_t)ldw * 2,
                          GEMM_BLOCK_K, (uint32_t)(pair ? 128 : block_n), Swizzle::B128);
  if (rc) return rc;
  {
    const uint64_t row_bytes = static_cast<uint64_t>(ldo) * 2, pd = static_cast<uint64_t>(padded);
    const uint64_t dims[4] = {static_cast<uint64_t>(N), static_cast<uint64_t>(grid), static_cast<uint64_t>(grid),
                              static_cast<uint64_t>(n_img)};   // only the projected part is addressable
    const uint64_t strides[3] = {row_bytes, row_bytes * pd, row_bytes * pd * pd};
    const uint32_t box[4] = {64, 32, 1, 1};
    rc = make_tensor_map_4d(&tm_out, out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dims, strides, box, Swizzle::B128);
    if (rc) return rc;
  }
  if (pair) rc = launch_gemm_2cta<__nv_bfloat16>(st, tm_a, tm_w, tm_out, bias, M, N, K, LA_ACT_NONE, 0, 0, grid);
  else if (block_n == 256) rc = launch_gemm<256, __nv_bfloat16>(st, tm_a, tm_w, tm_out, bias, M, N, K, LA_ACT_NONE, grid);
  else if (block_n == 128) rc = launch_gemm<128, __nv_bfloat16>(st, tm_a, tm_w, tm_out, bias, M, N, K, LA_ACT_NONE, grid);
  else rc = launch_gemm<64, __nv_bfloat16>(st, tm_a, tm_w, tm_out, bias, M, N, K, LA_ACT_NONE, grid);
  if (rc) return rc;
  if (padded > grid) {
    const long long total = static_cast<long long>(n_img) * (padded * padded - grid * grid) * (N / 8);
    long long blocks = (total + 255) / 256;
    const long long cap = 8ll * sm_count();
    if (blocks > cap) blocks = cap;
    fill_grid_padding_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(static_cast<__nv_bfloat16*>(out), ldo, bias, n_img,
                                                                     grid, padded, N / 8);
    LA_CHECK_CUDA(cudaGetLastError());
  }
  return LA_OK;
}

extern "C" int la_conv3x3_bf16(void* stream, const void* x, int n_img, int H, int W, int C, const void* w,
                               long long ldw, const float* bias, void* out, long long ldo, int out_dtype, int N,
                               int act) {
  using namespace la;
  LA_CHECK_ARG(x && w && out, "la_conv3x3_bf16: null pointer");
  LA_CHECK_ARG(n_img > 0 && H > 0 && N > 0, "la_conv3x3_bf16: empty problem");
  LA_CHECK_ARG(W == 64 && H % 4 == 0, "la_conv3x3_bf16: built for 64-wide feature maps with H %% 4 == 0 (got %dx%d)", H, W);
  LA_CHECK_ARG(C % 64 == 0 && N >= 256 && N % 8 == 0, "la_conv3x3_bf16: C must be a multiple of 64 and N >= 256");
  LA_CHECK_ARG(ldw % 8 == 0 && ldw >= 9ll * C, "la_conv3x3_bf16: ldw must cover the 9*C taps");
  LA_CHECK_ARG(out_dtype == LA_DTYPE_BF16 || out_dtype == LA_DTYPE_F32, "la_conv3x3_bf16: bad out_dtype %d", out_dtype);
  LA_CHECK_ARG((ldo * (out_dtype == LA_DTYPE_BF16 ? 2 : 4)) % 16 == 0, "la_conv3x3_bf16: ldo must give 16B rows");
  LA_CHECK_ARG((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(out)) % 16 == 0,
               "la_conv3x3_bf16: pointers must be 16-byte aligned");
  LA_CHECK_ARG(act >= LA_ACT_NONE && act <= LA_ACT_RELU, "la_conv3x3_bf16: bad act %d", act);
  const long long Mll = static_cast<long long>(n_img) * H * W;
  LA_CHECK_ARG(Mll < (1ll << 31), "la_conv3x3_bf16: too many pixels");
  const int M = static_cast<int>(Mll), K = 9 * C;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUtensorMap tm_a, tm_w, tm_out;
  const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)n_img};
  const uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
  const uint32_t box[4] = {64, 64, 2, 1};
  int rc = make_tensor_map_4d(&tm_a, x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dims, strides, box, Swizzle::B128);
  if (rc) return rc;
  rc = make_tensor_map_2d(&tm_w, w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (uint64_t)K, (uint64_t)N, (uint64_t)ldw * 2,
                          GEMM_BLOCK_K, 128, Swizzle::B128);
  if (rc) return rc;
  if (out_dtype == LA_DTYPE_BF16) {
    rc = make_tensor_map_2d(&tm_out, out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (uint64_t)N, (uint64_t)M, (uint64_t)ldo * 2,
                            64, 32, Swizzle::B128);
    if (rc) return rc;
    return launch_gemm_2cta<__nv_bfloat16, true>(st, tm_a, tm_w, tm_out, bias, M, N, K, act, C, H);
  }
  rc = make_tensor_map_2d(&tm_out, out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (uint64_t)N, (uint64_t)M, (uint64_t)ldo * 4, 32,
                          32, Swizzle::B128);
  if (rc) return rc;
  return launch_gemm_2cta<float, true>(st, tm_a, tm_w, tm_out, bias, M, N, K, act, C, H);
}

extern "C" int la_gemm_bf16_accumulate(void* stream, const void* a, long long lda, const void* w, long long ldw,
                                       const float* bias, float* x, long long ldx, int M, int N, int K) {
  using namespace la;
  LA_CHECK_ARG(a && w && x, "la_gemm_bf16_accumulate: null pointer");
  LA_CHECK_ARG(M >= 2048 && N >= 256 && K > 0, "la_gemm_bf16_accumulate: built for the CTA-pair kernel (M >= 2048, N >= 256)");
  LA_CHECK_ARG(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && N % 8 == 0 && ldx % 4 == 0,
               "la_gemm_bf16_accumulate: K/lda/ldw/N must be multiples of 8, ldx of 4");
  LA_CHECK_ARG((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(x)) % 16 == 0,
               "la_gemm_bf16_accumulate: pointers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUtensorMap tm_a, tm_w, tm_out;
  int rc = make_tensor_map_2d(&tm_a, a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2,
                              GEMM_BLOCK_K, GEMM_BLOCK_M, Swizzle::B128);
  if (rc) return rc;
  rc = make_tensor_map_2d(&tm_w, w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (uint64_t)K, (uint64_t)N, (uint64_t)ldw * 2,
                          GEMM_BLOCK_K, 128, Swizzle::B128);
  if (rc) return rc;
  rc = make_tensor_map_2d(&tm_out, x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (uint64_t)N, (uint64_t)M, (uint64_t)ldx * 4, 32, 32,
                          Swizzle::B128);
  if (rc) return rc;
  // long-K GEMMs hide the residual chunk's load behind their own tile; short-K ones get the prefetching epilogue
  // (one pipeline stage less, a second staging buffer per epilogue warp)
  const int mode = LA_GEMM_ACC_MODE ? LA_GEMM_ACC_MODE : (K >= 2048 ? 1 : 2);
  if (mode == 2) return launch_gemm_2cta<float, false, 2>(st, tm_a, tm_w, tm_out, bias, M, N, K, LA_ACT_NONE);
  return launch_gemm_2cta<float, false, 1>(st, tm_a, tm_w, tm_out, bias, M, N, K, LA_ACT_NONE);
}
