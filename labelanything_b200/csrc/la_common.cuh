// labelanything_b200 — shared device-side helpers for the sm_100a kernels.
//
// Thin inline-PTX wrappers for the Blackwell primitives the kernels in this directory are
// built from: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM alloc / ld / st /
// commit) and the shared-memory + instruction descriptors tcgen05.mma consumes.
// Everything here is sm_100a-only; there is no fallback path.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/labelanything_b200.h"

namespace la {

// ----------------------------------------------------------------------------------------
// error plumbing shared by the C-ABI entry points
// ----------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);

// (error codes LA_OK / LA_ERR_* come from the public header)

#define LA_CHECK_ARG(cond, ...)                      \
  do {                                               \
    if (!(cond)) {                                   \
      ::la::set_last_error(__VA_ARGS__);             \
      return LA_ERR_INVALID;                   \
    }                                                \
  } while (0)

#define LA_CHECK_CUDA(expr)                                                          \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      ::la::set_last_error("%s failed: %s%s", #expr, cudaGetErrorString(_e),         \
                           _e == cudaErrorMemoryAllocation ? " (out of memory)" : ""); \
      return LA_ERR_CUDA;                                                      \
    }                                                                                \
  } while (0)

// Host: build a 2D/3D tiled tensor map (driver entry point resolved at run time, so the
// library carries no link-time dependency on libcuda).
enum class Swizzle : int { None = 0, B32 = 1, B64 = 2, B128 = 3 };
int make_tensor_map_2d(CUtensorMap* map, const void* base, CUtensorMapDataType dtype, uint64_t inner,
                       uint64_t outer, uint64_t outer_stride_bytes, uint32_t box_inner, uint32_t box_outer,
                       Swizzle swizzle);
int make_tensor_map_3d(CUtensorMap* map, const void* base, CUtensorMapDataType dtype, uint64_t d0, uint64_t d1,
                       uint64_t d2, uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0, uint32_t box1,
                       uint32_t box2, Swizzle swizzle);
int make_tensor_map_4d(CUtensorMap* map, const void* base, CUtensorMapDataType dtype, const uint64_t dims[4],
                       const uint64_t strides_bytes[3], const uint32_t box[4], Swizzle swizzle);
int sm_count();

#ifdef __CUDACC__

// ----------------------------------------------------------------------------------------
// small utilities
// ----------------------------------------------------------------------------------------
// n / d for n < 2^31 and a run-time d without the ~25-instruction division sequence (MUFU.RCP + conversions):
// mul = ceil(2^(31 + s) / d), s = ceil(log2 d); n / d = umulhi(n, mul) >> (s - 1)   (d = 1: identity)
struct FastDiv {
  uint32_t mul, shift, d;
};
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  f.mul = 0;
  f.shift = 0;
  if (d > 1) {
    uint32_t s = 0;
    while ((1ull << s) < d) ++s;
    f.mul = static_cast<uint32_t>(((1ull << (31 + s)) + d - 1) / d);
    f.shift = s - 1;
  }
  return f;
}
__device__ __forceinline__ uint32_t fast_div(uint32_t n, const FastDiv& f) {
  return f.d == 1 ? n : (__umulhi(n, f.mul) >> f.shift);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, 0xFFFFFFFF;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// ----------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "LA_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra LA_DONE;\n\t"
      "bra LA_WAIT;\n\t"
      "LA_DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// generic-proxy writes to smem -> visible to the async proxy (TMA store, tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ----------------------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0,
                                            int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0,
                                            int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0,
                                            int32_t c1, int32_t c2, int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
// the same box ADDED to global memory (element type of the tensor map: fp32 here) -- the split-K epilogue
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* smem_src, int32_t c0, int32_t c1,
                                             int32_t c2, int32_t c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA, commit, ld/st, fences
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: A (M x 16 bf16 per instruction) is read from tensor memory, row = lane,
// two consecutive K elements per 32-bit column (what tcgen05.st.32x32b of packed bf16x2 words writes).
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Lean forms for issue loops that feed the tensor pipe with short (N = 64) instructions: the caller keeps the
// pre-shifted low word of each 128B-swizzle descriptor ((addr >> 4) & 0x3FFF, see umma_smem_desc_sw128) and adds byte
// offsets >> 4 to it; the high word (SBO 1024, version 1, SWIZZLE_128B) is the constant 0x40004040.
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return (smem_addr >> 4) & 0x3FFF; }
__device__ __forceinline__ void umma_ss_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                           bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate ? 1u : 0u), "r"(0x40004040u)
      : "memory");
}
__device__ __forceinline__ void umma_ts_lo(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc,
                                           bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(accumulate ? 1u : 0u), "r"(0x40004040u)
      : "memory");
}
// shared memory -> tensor memory: 128 rows x 32 bytes of the tile described by the (pre-shifted) descriptor low word,
// written to 8 columns starting at taddr (row r -> lane r); ordered with tcgen05.mma in issue order.
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint32_t src_lo) {
  asm volatile(
      "{\n\t"
      ".reg .b64 ds;\n\t"
      "mov.b64 ds, {%1, %2};\n\t"
      "tcgen05.cp.cta_group::1.128x256b [%0], ds;\n\t"
      "}\n" ::"r"(taddr),
      "r"(src_lo), "r"(0x40004040u)
      : "memory");
}
// ---- CTA-pair (cta_group::2) forms: one MMA spans the tensor cores, shared and tensor memory of the two CTAs of a
// cluster; the even ("leader") CTA issues it.  Shared-window addresses carry the CTA rank of the pair in bit 24:
// masking it addresses the leader's copy of a barrier from either CTA.
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_dst, uint32_t ncols) {  // one warp in EACH CTA
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// executed by both CTAs: the box lands in the executing CTA's shared memory, the bytes are counted on the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_2cta(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0,
                                                 int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_2cta(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0,
                                                 int32_t c1, int32_t c2, int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread -> arrive(1) on the barrier at this offset in every CTA of cta_mask
__device__ __forceinline__ void umma_commit_2cta_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
// arrive on the leader CTA's copy of a barrier (local for the leader, remote for its peer)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_BIT_MASK) : "memory");
}

// all previously issued tcgen05.mma of this thread -> arrive(1) on the mbarrier when they complete
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// TMEM -> registers: warp w may touch lanes [32*(w%4), +32); thread t gets row (lane) t, 32 consecutive columns.
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// pointer forms (the array must be fully unrolled / register resident)
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// register re-allocation between warpgroups (all warps of the warpgroup must execute it)
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------------------
// descriptors
// ----------------------------------------------------------------------------------------
// Shared-memory matrix descriptor for a 128B-swizzled operand tile whose rows are 128 bytes
// (64 bf16) and whose 8-row groups sit 1024 bytes apart -- exactly what a TMA box with
// CU_TENSOR_MAP_SWIZZLE_128B and a 128-byte inner extent writes.  Used both for K-major
// operands (row = M/N index, 128B = 64 K elements) and MN-major operands (row = K index,
// 128B = 64 M/N elements); the major-ness is selected in the instruction descriptor.
//   bits [0,14)  start address >> 4         bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4    bits [46,48) descriptor version (1 on sm_100)
//   bits [61,64) layout type (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes = 0,
                                                         uint32_t sbo_bytes = 1024) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

// Instruction descriptor for kind::f16 with bf16 A/B and fp32 accumulation.
//   bits [4,6) D format (1 = f32)   bits [7,10) A format (1 = bf16)   bits [10,13) B format
//   bit 15 A major (0 = K, 1 = MN)  bit 16 B major                    bits [17,23) N >> 3
//   bits [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major = 0, int b_mn_major = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// exact-erf GELU (reference: nn.GELU() default, label_anything/models/common.py:24).
// erf(z), z >= 0, via Abramowitz-Stegun 7.1.28: 1 - (1 + a1 z + ... + a6 z^6)^-16, |err| <= 3e-7 --
// far below bf16 resolution; one MUFU (rcp.approx) and ~15 FMA-pipe instructions per element.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float p = fmaf(0.0000430638f, z, 0.0002765672f);
  p = fmaf(p, z, 0.0001520143f);
  p = fmaf(p, z, 0.0092705272f);
  p = fmaf(p, z, 0.0422820123f);
  p = fmaf(p, z, 0.0705230784f);
  p = fmaf(p, z, 1.0f);
  p = p * p;
  p = p * p;
  p = p * p;
  p = p * p;                       // (..)^16 ; overflows to +inf for huge |x| -> rcp gives 0 -> erf = 1
  const float e = 1.0f - rcp_approx(p);
  const float h = 0.5f * x;
  return fmaf(h, copysignf(e, x), h);
}

// Two exact-erf GELUs at once on the sm_100 packed fp32x2 forms (FFMA2 / FMUL2): the same A&S 7.1.28 evaluation as
// gelu_erf with half the issue slots -- the GELU epilogue of the MLP GEMM is instruction-issue bound
// (128 x 256 outputs per tile against 6144 tensor cycles at K = 768).
//   gelu(x) = h + |h| * erf(|x| / sqrt 2),  h = x / 2
__device__ __forceinline__ void gelu_erf_x2(float& x0, float& x1) {
  const float z0 = fabsf(x0) * 0.70710678118654752f, z1 = fabsf(x1) * 0.70710678118654752f;
  float r0, r1;
  asm("{\n\t"
      ".reg .b64 z, p, c, one;\n\t"
      ".reg .f32 q0, q1;\n\t"
      "mov.b64 z, {%2, %3};\n\t"
      "mov.b64 c, {%4, %4};\n\t"
      "mov.b64 p, {%5, %5};\n\t"
      "fma.rn.f32x2 p, c, z, p;\n\t"      // a6 z + a5
      "mov.b64 c, {%6, %6};\n\t"
      "fma.rn.f32x2 p, p, z, c;\n\t"      // .. + a4
      "mov.b64 c, {%7, %7};\n\t"
      "fma.rn.f32x2 p, p, z, c;\n\t"      // .. + a3
      "mov.b64 c, {%8, %8};\n\t"
      "fma.rn.f32x2 p, p, z, c;\n\t"      // .. + a2
      "mov.b64 c, {%9, %9};\n\t"
      "fma.rn.f32x2 p, p, z, c;\n\t"      // .. + a1
      "mov.b64 one, {%10, %10};\n\t"
      "fma.rn.f32x2 p, p, z, one;\n\t"    // .. + 1
      "mul.rn.f32x2 p, p, p;\n\t"
      "mul.rn.f32x2 p, p, p;\n\t"
      "mul.rn.f32x2 p, p, p;\n\t"
      "mul.rn.f32x2 p, p, p;\n\t"         // (..)^16 ; +inf for huge |x| -> rcp 0 -> erf 1
      "mov.b64 {q0, q1}, p;\n\t"
      "rcp.approx.ftz.f32 q0, q0;\n\t"
      "rcp.approx.ftz.f32 q1, q1;\n\t"
      "mov.b64 p, {q0, q1};\n\t"
      "mov.b64 c, {%11, %11};\n\t"
      "fma.rn.f32x2 p, p, c, one;\n\t"    // erf = 1 - 1/(..)^16
      "mov.b64 {%0, %1}, p;\n\t"
      "}"
      : "=f"(r0), "=f"(r1)
      : "f"(z0), "f"(z1), "f"(0.0000430638f), "f"(0.0002765672f), "f"(0.0001520143f), "f"(0.0092705272f),
        "f"(0.0422820123f), "f"(0.0705230784f), "f"(1.0f), "f"(-1.0f));
  const float h0 = 0.5f * x0, h1 = 0.5f * x1;
  x0 = fmaf(fabsf(h0), r0, h0);
  x1 = fmaf(fabsf(h1), r1, h1);
}

// GELU for bf16 outputs: erf(z) ~ tanh(z (a1 + a3 z^2 + a5 z^4)) (|err| <= 3.7e-5, odd, so no sign handling) with the
// hardware tanh (MUFU.TANH, relative error 2^-11): |gelu error| <= 2.5e-4 |x|, an order of magnitude below the bf16
// rounding of the result, at 4.5 issue slots per element instead of 9.5 -- the epilogue of the MLP's first GEMM is
// issue bound (128 x 256 GELUs per 6144 tensor cycles).  fp32 outputs keep the 3e-7 form above.
__device__ __forceinline__ void gelu_tanh_erf_x2(float& x0, float& x1) {
  float t0, t1, h0, h1;
  asm("{\n\t"
      ".reg .b64 x, z, z2, p, c, h;\n\t"
      ".reg .f32 q0, q1;\n\t"
      "mov.b64 x, {%4, %5};\n\t"
      "mov.b64 c, {%6, %6};\n\t"
      "mul.rn.f32x2 z, x, c;\n\t"          // z = x / sqrt 2
      "mul.rn.f32x2 z2, z, z;\n\t"
      "mov.b64 p, {%7, %7};\n\t"
      "mov.b64 c, {%8, %8};\n\t"
      "fma.rn.f32x2 p, p, z2, c;\n\t"      // a5 z^2 + a3
      "mov.b64 c, {%9, %9};\n\t"
      "fma.rn.f32x2 p, p, z2, c;\n\t"      // .. z^2 + a1
      "mul.rn.f32x2 p, p, z;\n\t"
      "mov.b64 {q0, q1}, p;\n\t"
      "tanh.approx.f32 q0, q0;\n\t"
      "tanh.approx.f32 q1, q1;\n\t"
      "mov.f32 %0, q0;\n\t"
      "mov.f32 %1, q1;\n\t"
      "mov.b64 c, {%10, %10};\n\t"
      "mul.rn.f32x2 h, x, c;\n\t"          // h = x / 2
      "mov.b64 {%2, %3}, h;\n\t"
      "}"
      : "=f"(t0), "=f"(t1), "=f"(h0), "=f"(h1)
      : "f"(x0), "f"(x1), "f"(0.70710678118654752f), "f"(-0.0017864745f), "f"(0.1040811837f), "f"(1.1281434298f),
        "f"(0.5f));
  x0 = fmaf(h0, t0, h0);
  x1 = fmaf(h1, t1, h1);
}

#endif  // __CUDACC__

}  // namespace la
