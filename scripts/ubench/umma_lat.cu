// Microbenchmark: tcgen05.mma issue->commit latency and chain throughput on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../ferrum-infer-rs_b200/csrc -o umma_lat umma_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace fb2;

template <int N, bool kTS>
__global__ void __launch_bounds__(128) bench(long long *out, int n_mma, int n_acc, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;  // 1.0
  ptx::fence_proxy_async_smem();
  if (warp == 0) {
    if (lane == 0) { ptx::mbar_init(ptx::smem_u32(&s_bar), 1); ptx::mbar_fence_init(); }
    __syncwarp();
    ptx::tmem_alloc<512>(ptx::smem_u32(&s_tmem));
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = s_tmem;
  uint32_t z[16];
  for (int i = 0; i < 16; i++) z[i] = 0x3c003c00u;
  for (int c = 0; c < 4; c++) ptx::tmem_st_x16(tm + ((uint32_t)(warp * 32) << 16) + 448 + c * 16, z);
  ptx::tmem_wait_st();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 0 && ptx::elect_one()) {
    constexpr uint32_t idesc = ptx::umma_idesc_f16(128, N);
    const uint32_t bar = ptx::smem_u32(&s_bar);
    const uint32_t sb = ptx::smem_u32(smem);
    uint32_t phase = 0;
    for (int r = 0; r < reps; r++) {
      long long t0 = clock64();
      const uint64_t bd0 = ptx::umma_desc_kmajor_sw128(sb);
      const uint64_t ad0 = ptx::umma_desc_kmajor_sw128(sb + 32768);
      for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const uint32_t d = tm + (uint32_t)((j % 4 < n_acc ? j % 4 : 0) * N);
          if (kTS) ptx::umma_f16_ts(d, tm + 448 + j * 8, bd0 + (uint64_t)((j & 3) * 2), idesc, 1u);
          else ptx::umma_f16_ss(d, ad0 + (uint64_t)((j & 3) * 2), bd0 + (uint64_t)((j & 3) * 2), idesc, 1u);
        }
      }
      long long t1 = clock64();
      ptx::tc_commit(bar);
      ptx::mbar_wait(bar, phase);
      phase ^= 1;
      long long t2 = clock64();
      out[r * 2] = t1 - t0;
      out[r * 2 + 1] = t2 - t0;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<512>(tm);
}

template <int N, bool kTS>
void run(const char *name) {
  long long *d; cudaMalloc(&d, 64 * 8);
  cudaFuncSetAttribute(bench<N, kTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
  for (int n_acc : {1, 4}) {
    if (n_acc * N > 448) continue;
    for (int n_mma : {8, 32, 128}) {
      bench<N, kTS><<<1, 128, 65536 + 1024>>>(d, n_mma, n_acc, 4);
      long long h[8];
      cudaError_t e = cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
      printf("%s N=%3d n_acc=%d n_mma=%3d  issue=%6lld cyc  issue+complete=%6lld cyc  (%.1f cyc/mma)\n", name, N, n_acc,
             n_mma, h[6], h[7], (double)h[7] / n_mma);
    }
  }
  cudaFree(d);
}

int main() {
  run<16, true>("TS"); run<32, true>("TS"); run<64, true>("TS"); run<128, true>("TS"); run<256, true>("TS");
  run<16, false>("SS"); run<64, false>("SS"); run<256, false>("SS");
  return 0;
}
