#include <cstdio>
#include <cstdint>
#include "../../speinet_b200/csrc/fastdiv.h"
int main() {
  long long bad = 0;
  int ds[] = {1,2,3,5,7,8,9,17,40,45,80,320,321,57600,57601,65536,100003,1<<20,(1<<30)+1,2147483647};
  for (int d : ds) { FastDiv f = make_fastdiv(d);
    for (long long n = 0; n < (1ll<<31); n += 9973) if (fast_div((int)n, f) != (int)(n / d)) ++bad;
    for (long long n = (1ll<<31) - 100000; n < (1ll<<31); ++n) if (fast_div((int)n, f) != (int)(n / d)) ++bad;
    for (long long k = 1; k * d < (1ll<<31) && k < 2000000; k += 37) { long long n = k * d; if (fast_div((int)n, f) != (int)(n/d)) ++bad; if (fast_div((int)(n-1), f) != (int)((n-1)/d)) ++bad; }
  }
  printf("bad %lld\n", bad); return bad != 0; }
