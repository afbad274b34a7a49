// occupancy probe: which resource stops two 320-thread CTAs with ~105 KB of shared memory each?
#include <cstdio>
#include <cuda_runtime.h>
template <int R> __global__ void __launch_bounds__(384, 2) k(float *o) {
  extern __shared__ float s[];
  float acc[R];
  for (int i = 0; i < R; i++) acc[i] = s[(threadIdx.x + i) & 1023];
  __syncthreads();
  float t = 0; for (int i = 0; i < R; i++) t += acc[i] * acc[(i + 1) % R];
  o[threadIdx.x] = t;
}
template <typename K> void probe(const char *name, K kern) {
  cudaFuncAttributes fa{}; cudaFuncGetAttributes(&fa, kern);
  printf("%s regs %d\n", name, fa.numRegs);
  for (int threads : {256, 320, 384}) {
    for (int smem : {48 * 1024, 90 * 1024, 100 * 1024, 107520, 110 * 1024, 113 * 1024}) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      int n = -1; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, threads, smem);
      size_t av = 0; cudaOccupancyAvailableDynamicSMemPerBlock(&av, kern, 2, threads);
      printf("  threads %d smem %6d -> blocks/SM %d  (avail dyn smem @2 blocks: %zu)\n", threads, smem, n, av);
    }
  }
}
int main() {
  int v;
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerMultiprocessor, 0); printf("smem/SM %d\n", v);
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, 0); printf("smem/block optin %d\n", v);
  cudaDeviceGetAttribute(&v, cudaDevAttrReservedSharedMemoryPerBlock, 0); printf("reserved/block %d\n", v);
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxRegistersPerMultiprocessor, 0); printf("regs/SM %d\n", v);
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxThreadsPerMultiProcessor, 0); printf("threads/SM %d\n", v);
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxBlocksPerMultiprocessor, 0); printf("blocks/SM %d\n", v);
  probe("k<8>", k<8>); probe("k<60>", k<60>);
  return 0;
}
