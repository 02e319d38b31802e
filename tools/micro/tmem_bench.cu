// Micro-benchmark: TMEM load/store throughput and its interference with tcgen05.mma.
#include <cstdio>
#include "la_common.cuh"
using namespace la;

// warps 0..3: control (warp 0 alloc, warp 1 MMA issuer); warps 4..4+NW-1: TMEM readers/writers
// mma_mode: 0 none, 1 SS N=64, 2 TS N=64;  tm_mode: 0 none, 1 LDTM 64 cols, 2 STTM 32 cols, 3 both
__global__ void __launch_bounds__(384, 1) bench(long long* out, int mma_mode, int tm_mode, int n_iter) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 384) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1 && mma_mode) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem) + 32768;
    long long t0 = clock64();
    if (elect_one()) {
#pragma unroll 16
      for (int i = 0; i < n_iter * 8; ++i) {
        const uint32_t d = tm + 256 + (i & 1) * 64;
        const int ks = i & 3;
        if (mma_mode == 2) umma_bf16_ts(d, tm + 448 + ks * 8, umma_smem_desc_sw128(b_base + ks * 32), idesc, 1u);
        else umma_bf16_ss(d, umma_smem_desc_sw128(a_base + ks * 32), umma_smem_desc_sw128(b_base + ks * 32), idesc, 1u);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  if (warp >= 4 && tm_mode) {
    const uint32_t t_s = tm + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp - 4) >> 2) * 128;
    uint32_t v[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = i;
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < n_iter; ++it) {
      if (tm_mode & 1) {
        tmem_ld_x32(t_s, v);
        tmem_ld_x32(t_s + 32, v + 32);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 64; ++i) acc += v[i];
      }
      if (tm_mode & 2) {
        tmem_st_32x32b_x16(t_s, v);
        tmem_st_32x32b_x16(t_s + 16, v + 16);
        tmem_st_wait();
      }
    }
    long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) { out[1 + warp] = t1 - t0; out[20] = acc; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 256);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int n_iter = 512;
  const char* mm[] = {"no MMA", "SS N=64 MMA", "TS N=64 MMA"};
  const char* tmn[] = {"no TMEM traffic", "LDTM 64 cols", "STTM 32 cols", "LDTM 64 + STTM 32"};
  for (int mma_mode = 0; mma_mode < 3; ++mma_mode)
    for (int tm_mode = 0; tm_mode < 4; ++tm_mode) {
      if (!mma_mode && !tm_mode) continue;
      cudaMemset(d_out, 0, 256);
      bench<<<148, 384, 100 * 1024>>>(d_out, mma_mode, tm_mode, n_iter);
      long long h[32];
      cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
      cudaError_t e = cudaGetLastError();
      printf("%-14s + %-20s: %6.1f cyc/mma ; per-warp cyc per tile-iteration (8 warps):", mm[mma_mode], tmn[tm_mode],
             (double)h[0] / (n_iter * 8));
      for (int w = 4; w < 12; ++w) printf(" %5.0f", (double)h[1 + w] / n_iter);
      printf(" %s\n", e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
