#include <cstdio>
#include <cstdint>
struct FastDiv { unsigned m, sh; int d; };
inline FastDiv make_fastdiv(int d) {
  FastDiv f{0u, 0u, d};
  if (d > 1) { int lg = 0; while ((1ll << lg) < (long long)d) ++lg; const int k = 31 + lg;
    f.m = (unsigned)(((1ull << k) + (unsigned long long)d - 1) / (unsigned long long)d); f.sh = (unsigned)(k - 32); }
  return f; }
static int fast_div(int n, FastDiv f) { return f.d == 1 ? n : (int)((((unsigned long long)(unsigned)n * f.m) >> 32) >> f.sh); }
int main() {
  long long bad = 0;
  int ds[] = {1,2,3,5,7,8,9,17,40,45,80,320,321,57600,57601,65536,100003,1<<20,(1<<30)+1,2147483647};
  for (int d : ds) { FastDiv f = make_fastdiv(d);
    for (long long n = 0; n < (1ll<<31); n += 9973) if (fast_div((int)n, f) != (int)(n / d)) ++bad;
    for (long long n = (1ll<<31) - 100000; n < (1ll<<31); ++n) if (fast_div((int)n, f) != (int)(n / d)) ++bad;
    for (long long k = 1; k * d < (1ll<<31) && k < 2000000; k += 37) { long long n = k * d; if (fast_div((int)n, f) != (int)(n/d)) ++bad; if (fast_div((int)(n-1), f) != (int)((n-1)/d)) ++bad; }
  }
  printf("bad %lld\n", bad); return bad != 0; }
