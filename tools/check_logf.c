/* C twin of nc_logf (nanocall_b200/csrc/nc_device.cuh) checked against this box's libm logf on a dense sample of
 * all positive floats.  logf_data.inc = the 36 doubles of glibc 2.39's __logf_data (16 x {invc, logc}, ln2, poly[3]),
 * dumped from libm-2.39.a's e_logf_data.o:  objcopy -O binary --only-section=.rodata e_logf_data.o  */
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
static const uint64_t D[36] = {
#include "logf_data.inc"
};
static inline double dd(int i){ double d; memcpy(&d,&D[i],8); return d; }
static float my_logf(float x){
  uint32_t ix; memcpy(&ix,&x,4);
  if (ix == 0x3f800000) return 0;
  if (ix - 0x00800000 >= 0x7f800000 - 0x00800000) {
    if (ix * 2 == 0) return -INFINITY;
    if (ix == 0x7f800000) return x;
    if ((ix & 0x80000000) || ix * 2 >= 0xff000000) return NAN;
    float xs = x * 0x1p23f; memcpy(&ix,&xs,4); ix -= 23 << 23;
  }
  uint32_t tmp = ix - 0x3f330000;
  int i = (tmp >> 19) % 16;
  int k = (int32_t)tmp >> 23;
  uint32_t iz = ix - (tmp & 0xff800000);
  double invc = dd(2*i), logc = dd(2*i+1);
  float zf; memcpy(&zf,&iz,4);
  double z = zf;
  double r = z * invc - 1;
  double y0 = logc + (double)k * dd(32);
  double r2 = r * r;
  double y = dd(34) * r + dd(35);
  y = dd(33) * r2 + y;
  y = y * r2 + (y0 + r);
  return (float)y;
}
int main(){
  long bad=0, tot=0;
  for (uint32_t u = 1; u < 0x7f800000u; u += 61) { float x; memcpy(&x,&u,4);
    float a = logf(x), b = my_logf(x); tot++; if (memcmp(&a,&b,4)) { bad++; if (bad<10) printf("x=%a libm=%a mine=%a\n", x,a,b);} }
  printf("tot=%ld bad=%ld\n", tot, bad);
  return bad != 0;
}
