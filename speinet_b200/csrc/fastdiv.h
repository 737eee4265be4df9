// Division of a non-negative int (< 2^31) by a launch-invariant positive divisor with one multiply-high and one shift
// (exact: m = ceil(2^k / d), k = 31 + ceil(log2 d); the error term n * (m d - 2^k) stays below 2^k for n < 2^31).  The compiler's
// generic 32-bit division is ~20 instructions; the rescoring kernels decode several indices per exact score.
// Plain C++ (no CUDA headers) so that tests/test_host.py can compile it with g++ and compare with `/` exhaustively.
#pragma once

struct FastDiv {
  unsigned m, sh;
  int d;
};
inline FastDiv make_fastdiv(int d) {
  FastDiv f{0u, 0u, d};
  if (d > 1) {
    int lg = 0;
    while ((1ll << lg) < (long long)d) ++lg;              // ceil(log2 d)
    const int k = 31 + lg;
    f.m = (unsigned)(((1ull << k) + (unsigned long long)d - 1) / (unsigned long long)d);
    f.sh = (unsigned)(k - 32);
  }
  return f;
}
#ifdef __CUDACC__
__device__ __forceinline__ int fast_div(int n, const FastDiv f) { return f.d == 1 ? n : (int)(__umulhi((unsigned)n, f.m) >> f.sh); }
#else
inline int fast_div(int n, const FastDiv f) {
  return f.d == 1 ? n : (int)((unsigned)(((unsigned long long)(unsigned)n * f.m) >> 32) >> f.sh);
}
#endif
