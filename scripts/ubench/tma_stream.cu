// Microbenchmark: how fast can persistent CTAs stream a contiguous buffer with cp.async.bulk?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../ferrum-infer-rs_b200/csrc -o tma_stream tma_stream.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace fb2;

__global__ void __launch_bounds__(64) stream_kernel(const uint8_t *src, size_t total_stages, int stage_bytes, int stages,
                                                    int hint, unsigned long long *sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sb = ptx::smem_u32(smem);
  const uint32_t bars = sb + stages * stage_bytes;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; i++) { ptx::mbar_init(bars + 8 * i, 1); ptx::mbar_init(bars + 8 * (16 + i), 1); }
    ptx::mbar_fence_init();
  }
  __syncthreads();
  const size_t s0 = total_stages * blockIdx.x / gridDim.x, s1 = total_stages * (blockIdx.x + 1) / gridDim.x;
  if (warp == 0) {
    if (lane == 0) {
      const uint64_t pol = ptx::policy_evict_first();
      int st = 0; uint32_t par = 0;
      for (size_t s = s0; s < s1; s++) {
        ptx::mbar_wait(bars + 8 * (16 + st), par ^ 1);
        ptx::mbar_arrive_expect_tx(bars + 8 * st, stage_bytes);
        if (hint) ptx::bulk_g2s_hint(sb + st * stage_bytes, src + s * stage_bytes, stage_bytes, bars + 8 * st, pol);
        else ptx::bulk_g2s(sb + st * stage_bytes, src + s * stage_bytes, stage_bytes, bars + 8 * st);
        if (++st == stages) { st = 0; par ^= 1; }
      }
    }
  } else {
    int st = 0; uint32_t par = 0;
    unsigned long long acc = 0;
    for (size_t s = s0; s < s1; s++) {
      ptx::mbar_wait(bars + 8 * st, par);
      acc += reinterpret_cast<const uint32_t *>(smem + st * stage_bytes)[lane];
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bars + 8 * (16 + st));
      if (++st == stages) { st = 0; par ^= 1; }
    }
    if (acc == 0x1234567) *sink = acc;
  }
}

int main() {
  const size_t bytes = 1ull << 30;   // 16 x 62 MB chain buffers fit
  uint8_t *d; cudaMalloc(&d, bytes); cudaMemset(d, 1, bytes);
  unsigned long long *sink; cudaMalloc(&sink, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  struct Cfg { int stage_bytes, stages, ctas_per_sm, hint; };
  Cfg cfgs[] = {{8704, 8, 2, 1}, {8704, 8, 2, 0}, {8704, 12, 2, 1}, {8704, 4, 4, 1}, {8704, 8, 3, 1}, {17408, 4, 2, 1}, {17408, 6, 2, 1},
                {32768, 3, 2, 1}, {32768, 6, 1, 1}, {8192, 12, 2, 0}, {4352, 16, 2, 1}};
  for (auto c : cfgs) {
    const size_t total_stages = bytes / c.stage_bytes;
    const int smem = c.stage_bytes * c.stages + 512;
    const int grid = 148 * c.ctas_per_sm;
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      stream_kernel<<<grid, 64, smem>>>(d, total_stages, c.stage_bytes, c.stages, c.hint, sink);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) { printf("error %s\n", cudaGetErrorString(err)); return 1; }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep == 1)
        printf("stage=%6d B stages=%2d ctas/sm=%d hint=%d in-flight/SM=%4d KB : %.0f GB/s\n", c.stage_bytes, c.stages,
               c.ctas_per_sm, c.hint, c.stage_bytes * c.stages * c.ctas_per_sm / 1024, total_stages * (double)c.stage_bytes / ms / 1e6);
    }
  }
  // small-buffer case: 61 MB (one gate_up GEMM), single launch, cold-ish
  for (int rep = 0; rep < 3; rep++) {
    const size_t total_stages = 7168;
    cudaEventRecord(e0);
    stream_kernel<<<296, 64, 8704 * 8 + 512>>>(d + (size_t)rep * (128 << 20), total_stages, 8704, 8, 1, sink);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("61MB single launch: %.1f us -> %.0f GB/s\n", ms * 1e3, total_stages * 8704.0 / ms / 1e6);
  }
  // chain of back-to-back 61 MB launches over DISTINCT buffers (what a sequence of gate_up GEMMs
  // streams), timed as a whole: the per-launch cost of a do-nothing streamer
  {
    const int chain = 16;
    const size_t total_stages = 7168;
    for (int stages : {8, 10}) {
      for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        for (int i = 0; i < chain; i++)
          stream_kernel<<<296, 64, 8704 * stages + 512>>>(d + (size_t)i * (62 << 20), total_stages, 8704, stages, 1, sink);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep == 1)
          printf("chain of %d x 61MB launches, %d stages: %.1f us per launch -> %.0f GB/s\n", chain, stages, ms * 1e3 / chain,
                 chain * total_stages * 8704.0 / ms / 1e6);
      }
    }
    // same, 148 CTAs (one per SM) with 16 stages, and 592 CTAs with 4
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      for (int i = 0; i < chain; i++)
        stream_kernel<<<148, 64, 8704 * 16 + 512>>>(d + (size_t)i * (62 << 20), total_stages, 8704, 16, 1, sink);
      cudaEventRecord(e1);
      cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep == 1) printf("chain, 148 CTAs x 16 stages: %.1f us per launch\n", ms * 1e3 / chain);
    }
  }
  return 0;
}
