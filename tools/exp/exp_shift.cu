// Experiment: semantics and cost of tcgen05.shift.down on sm_100a (no public description in this sandbox).
// Fill 128 lanes x 64 columns of tensor memory with lane*1000 + column, shift some strips, read back.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) k(uint32_t* out, long long* cyc, int nshift, int strips) {
  __shared__ uint32_t slot;
  __shared__ __align__(8) unsigned long long bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  const uint32_t taddr = base + ((uint32_t)(warp * 32) << 16);
  // write: lane L (global row 32*warp+lane), column c -> L*1000 + c
  for (int c0 = 0; c0 < 64; c0 += 8) {
    uint32_t v[8];
    for (int i = 0; i < 8; ++i) v[i] = (uint32_t)((warp * 32 + lane) * 1000 + c0 + i);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr + c0), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int s = 0; s < nshift; ++s)
      for (int st = 0; st < strips; ++st)
        asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(base + 8 * st) : "memory");
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    }
    const long long t2 = clock64();
    cyc[0] = t1 - t0; cyc[1] = t2 - t1;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < 64; c0 += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr + c0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) out[(warp * 32 + lane) * 64 + c0 + i] = v[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(64) : "memory");
}

int main() {
  uint32_t* d; long long* c;
  cudaMalloc(&d, 128 * 64 * 4); cudaMalloc(&c, 16);
  static uint32_t h[128 * 64]; long long hc[2];
  for (int cfg = 0; cfg < 4; ++cfg) {
    const int nshift = cfg == 0 ? 1 : (cfg == 1 ? 2 : (cfg == 2 ? 1 : 16)), strips = cfg == 2 ? 4 : (cfg == 3 ? 8 : 1);
    k<<<1, 128>>>(d, c, nshift, strips);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cfg %d: CUDA error %s\n", cfg, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(hc, c, 16, cudaMemcpyDeviceToHost);
    printf("cfg %d: nshift %d strips %d  issue %lld cyc, wait %lld cyc\n", cfg, nshift, strips, hc[0], hc[1]);
    const int rows[] = {0, 1, 2, 3, 30, 31, 32, 33, 34, 63, 64, 65, 126, 127};
    for (int r : rows) {
      printf("  row %3d:", r);
      for (int col : {0, 1, 7, 8, 9, 15, 16, 31, 32, 63}) printf(" c%-2d=%-7u", col, h[r * 64 + col]);
      printf("\n");
    }
  }
  return 0;
}
