// Experiment: issue throughput of the instructions the tap-sharing epilogue is made of (SHFL, packed FADD2, FMNMX3), per SM,
// as a function of the resident warps -- is the 60-SHFL lane exchange bound by a per-SM pipe or by per-scheduler dispatch?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o exp_shfl.bin exp_shfl.cu && ./exp_shfl.bin
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kIters = 2000;

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(float* out, long long* cyc, float seed) {
  float x[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) x[i] = seed * (float)(threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < kIters; ++it) {
    if (MODE == 0) {          // 60 independent shuffles (30 up, 30 down), as in tap_sums
      float up[32], dn[32];
#pragma unroll
      for (int i = 1; i < 31; ++i) { up[i] = __shfl_up_sync(0xffffffffu, x[i - 1], 1); dn[i] = __shfl_down_sync(0xffffffffu, x[i + 1], 1); }
#pragma unroll
      for (int i = 1; i < 31; ++i) x[i] = up[i] + dn[i];   // 30 FADD to keep the values live (subtracted in the report)
    } else if (MODE == 1) {   // the 30 FADD alone
      float y[32];
#pragma unroll
      for (int i = 1; i < 31; ++i) y[i] = x[i - 1] + x[i + 1];
#pragma unroll
      for (int i = 1; i < 31; ++i) x[i] = y[i];
    } else if (MODE == 2) {   // 32 packed adds
#pragma unroll
      for (int i = 0; i < 32; i += 2)
        asm volatile("{\n.reg .b64 ra, rb, rc;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nadd.rn.f32x2 rc, ra, rb;\nmov.b64 {%0, %1}, rc;\n}"
                     : "=f"(x[i]), "=f"(x[i + 1]) : "f"(x[i]), "f"(x[i + 1]), "f"(x[(i + 2) & 31]), "f"(x[(i + 3) & 31]));
#pragma unroll
      for (int i = 0; i < 32; i += 2)
        asm volatile("{\n.reg .b64 ra, rb, rc;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nadd.rn.f32x2 rc, ra, rb;\nmov.b64 {%0, %1}, rc;\n}"
                     : "=f"(x[i]), "=f"(x[i + 1]) : "f"(x[i]), "f"(x[i + 1]), "f"(x[(i + 4) & 31]), "f"(x[(i + 5) & 31]));
    } else if (MODE == 3) {   // 32 three-input maxima, 32 independent chains, no moves
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = fmaxf(fmaxf(x[i], seed), x[(i + 7) & 31]);
    } else if (MODE == 4) {   // 32 two-input maxima
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = fmaxf(x[i], x[(i + 7) & 31]);
    } else if (MODE == 5) {   // 32 scalar adds
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = x[i] + x[(i + 7) & 31];
    } else if (MODE == 7) {   // the same exchange through shared memory: 8 STS.128 (own row) + 16 LDS.128 (rows m-1, m+1), 144-byte pitch
      extern __shared__ float4 xs[];
      float4* mine = xs + (threadIdx.x >> 5) * (32 * 9) + (threadIdx.x & 31) * 9;
      const int lane = threadIdx.x & 31;
      const float4* upr = mine - (lane > 0 ? 9 : 0);
      const float4* dnr = mine + (lane < 31 ? 9 : 0);
#pragma unroll
      for (int i = 0; i < 8; ++i) mine[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 a = upr[i], b = dnr[i];
        x[4 * i] = a.x + b.x; x[4 * i + 1] = a.y + b.y; x[4 * i + 2] = a.z + b.z; x[4 * i + 3] = a.w + b.w;
      }
      __syncwarp();
    } else if (MODE == 6) {   // 32 shuffles whose results feed nothing but the next shuffle of the same register
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = __shfl_up_sync(0xffffffffu, x[i], 1);
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
static double run(int warps, float* d, long long* c) {
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  k<MODE><<<148, warps * 32, warps * 32 * 144>>>(d, c, 1e-30f);
  cudaDeviceSynchronize();
  k<MODE><<<148, warps * 32, warps * 32 * 144>>>(d, c, 1e-30f);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return -1; }
  long long hc; cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
  return (double)hc / kIters;
}

int main() {
  float* d; long long* c;
  cudaMalloc(&d, 148 * 1024 * 4); cudaMalloc(&c, 8);
  printf("cycles per loop iteration per SM (all warps resident on one SM run the same loop)\n");
  printf("warps  60xSHFL+30xFADD  30xFADD+MOV  32xFADD2  32xFMNMX3  32xFMNMX  32xFADD  32xSHFL  -> per clk per SM: SHFL  FADD2  FMNMX3  FMNMX  FADD\n");
  for (int w : {1, 2, 4, 8, 12, 16, 32}) {
    const double a = run<0>(w, d, c), b = run<1>(w, d, c), p = run<2>(w, d, c), m = run<3>(w, d, c), m2 = run<4>(w, d, c), f = run<5>(w, d, c),
                 sh = run<6>(w, d, c), sm = run<7>(w, d, c);
    printf("%5d  %15.1f  %11.1f  %8.1f  %9.1f  %8.1f  %7.1f  %7.1f  -> %6.3f %6.3f %6.3f %6.3f %6.3f   smem exchange (8 STS.128 + 16 LDS.128 + 32 FADD): %7.1f\n", w, a, b, p, m, m2, f, sh, 32.0 * w / sh,
           32.0 * w / p, 32.0 * w / m, 32.0 * w / m2, 32.0 * w / f, sm);
  }
  return 0;
}
