#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
static uint64_t s = 88172645463325252ULL;
static inline uint64_t rnd(void){ s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static inline float asf(uint32_t u){ float f; memcpy(&f,&u,4); return f; }
int main(int argc, char** argv){
  long N = atol(argv[1]); long bad1=0, bad2=0;
  for (long i=0;i<N;i++){
    uint32_t ma = rnd() & 0x7fffff, mb = rnd() & 0x7fffff;
    int ea = 127 + (int)(rnd()%40) - 20, eb = 127 + (int)(rnd()%12) - 6;
    float a = asf(((rnd()&1)<<31) | (ea<<23) | ma), b = asf((eb<<23)|mb);
    float r = 1.0f/b;
    float q = a/b;
    float q0 = a*r;
    float e = fmaf(-q0, b, a);
    float q1 = fmaf(e, r, q0);
    if (q1 != q) { bad1++; if (bad1<5) printf("1-iter mismatch a=%a b=%a q=%a q1=%a\n", a,b,q,q1); }
    float e2 = fmaf(-q1, b, a);
    float q2 = fmaf(e2, r, q1);
    if (q2 != q) bad2++;
  }
  printf("N=%ld bad1=%ld bad2=%ld\n", N, bad1, bad2);
}
