// Experiment: can tcgen05.ld.32x32b read with a lane offset that is not a multiple of 32 (a free lane shift)?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128, 1) k(uint32_t* out, int lane_off) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  const uint32_t taddr = base + ((uint32_t)(warp * 32) << 16);
  uint32_t v[8];
  for (int i = 0; i < 8; ++i) v[i] = (uint32_t)((warp * 32 + lane) * 100 + i);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[8];
  const uint32_t raddr = base + ((uint32_t)(warp * 32 + lane_off) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(raddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 8; ++i) out[threadIdx.x * 8 + i] = r[i];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(32) : "memory");
}
int main() {
  uint32_t* d; cudaMalloc(&d, 128 * 8 * 4);
  static uint32_t h[128 * 8];
  for (int off : {0, 1, 2, 16, 31}) {
    cudaMemset(d, 0xff, sizeof(h));
    k<<<1, 128>>>(d, off);
    cudaError_t e = cudaDeviceSynchronize();
    printf("lane offset %d: %s\n", off, cudaGetErrorString(e));
    if (e != cudaSuccess) return 0;
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int t : {0, 1, 30, 31, 32, 33, 63, 64, 127}) printf("  thread %3d got %u (col0), %u (col7)\n", t, h[t * 8], h[t * 8 + 7]);
  }
  return 0;
}
