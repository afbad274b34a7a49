// Does a kernel that allocates tensor memory get two co-resident CTAs per SM?
// (occupancy query + an empirical all-CTAs-arrive barrier with a timeout)
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
template <int COLS> __global__ void __launch_bounds__(384, 2) k(int *arrive, int *ok) {
  extern __shared__ __align__(1024) uint8_t s[];
  __shared__ uint32_t taddr;
  if (COLS > 0 && threadIdx.x < 32) {
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(&taddr);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "n"(COLS > 0 ? COLS : 32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    s[0] = 1;
    atomicAdd(arrive, 1);
    long long t0 = clock64();
    bool good = true;
    while (*(volatile int *)arrive < (int)gridDim.x) { if (clock64() - t0 > 400000000LL) { good = false; break; } }
    if (!good) atomicAdd(ok, 1);
  }
  __syncthreads();
  if (COLS > 0 && threadIdx.x < 32) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS > 0 ? COLS : 32) : "memory");
  }
}
template <typename K> void run(const char *name, K kern, int smem) {
  int *d; cudaMalloc(&d, 8); cudaMemset(d, 0, 8);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int n = -1; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, 320, smem);
  kern<<<296, 320, smem>>>(d, d + 1);
  cudaError_t e = cudaDeviceSynchronize();
  int h[2]; cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-12s smem %6d: occupancy query %d; 296-CTA barrier: %d arrived, %d timed out (%s)\n", name, smem, n, h[0], h[1], cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  run("no tmem", k<0>, 107520);
  run("tmem 256", k<256>, 107520);
  run("tmem 128", k<128>, 107520);
  run("tmem 256", k<256>, 100 * 1024);
  run("tmem 256", k<256>, 90 * 1024);
  run("tmem 256", k<256>, 64 * 1024);
  run("tmem 32", k<32>, 107520);
  return 0;
}
