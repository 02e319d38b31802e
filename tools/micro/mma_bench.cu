// Micro-benchmark: issue cost of tcgen05.mma (M=128, bf16) vs N, operand source and accumulator dependence.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I labelanything_b200/csrc -o gpurun_out/mma_bench tools/micro/mma_bench.cu
#include <cstdio>
#include "la_common.cuh"
using namespace la;

// mode: 0 = SS, 1 = TS (A from TMEM), 2 = SS with MN-major B
template <int N, int MODE, int NACC>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int n_mma) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, MODE == 2 ? 1 : 0);
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem) + 32768;
    long long t0 = 0, t1 = 0, t2 = 0;
    for (int rep = 0; rep < 3; ++rep) {
      __syncwarp();
      t0 = clock64();
      if (elect_one()) {
#pragma unroll 16
        for (int i = 0; i < n_mma; ++i) {
          const uint32_t d = tm + (i % NACC) * N;           // accumulator tile
          const int ks = i & 3;
          if (MODE == 1) umma_bf16_ts(d, tm + 448 + ks * 8, umma_smem_desc_sw128(b_base + ks * 32), idesc, 1u);
          else if (MODE == 2) umma_bf16_ss(d, umma_smem_desc_sw128(a_base + ks * 32), umma_smem_desc_sw128(b_base + ks * 2048), idesc, 1u);
          else umma_bf16_ss(d, umma_smem_desc_sw128(a_base + ks * 32), umma_smem_desc_sw128(b_base + ks * 32), idesc, 1u);
        }
        umma_commit(&bar);
      }
      __syncwarp();
      t1 = clock64();
      mbar_wait(&bar, rep & 1);
      t2 = clock64();
    }
    if (lane == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int N, int MODE, int NACC>
void run(const char* name, long long* d_out) {
  const int n_acc = NACC;
  auto k = bench<N, MODE, NACC>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int n_mma = 8192;
  k<<<148, 128, 100 * 1024>>>(d_out, n_mma);
  long long h[2];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  printf("%-28s N=%3d acc=%d : issue %6.1f cyc/mma, complete %6.1f cyc/mma  (floor M*N/256 = %d) %s\n", name, N, n_acc,
         (double)h[0] / n_mma, (double)h[1] / n_mma, 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64);
  run<64, 0, 1>("SS", d_out);
  run<64, 0, 2>("SS", d_out);
  run<64, 0, 4>("SS", d_out);
  run<128, 0, 1>("SS", d_out);
  run<128, 0, 2>("SS", d_out);
  run<256, 0, 1>("SS", d_out);
  run<64, 1, 1>("TS (A in TMEM)", d_out);
  run<64, 1, 2>("TS (A in TMEM)", d_out);
  run<64, 1, 4>("TS (A in TMEM)", d_out);
  run<128, 1, 1>("TS (A in TMEM)", d_out);
  run<64, 2, 1>("SS, B MN-major", d_out);
  run<64, 2, 2>("SS, B MN-major", d_out);
  run<32, 0, 1>("SS", d_out);
  run<32, 0, 4>("SS", d_out);
  return 0;
}
