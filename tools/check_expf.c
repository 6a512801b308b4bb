#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
static const uint64_t T[32] = {
#include "exp2f_tab.inc" /* python tools/gen_exp2f_table.py > tools/exp2f_tab.inc */
};
static inline uint64_t asu64(double d){uint64_t u; memcpy(&u,&d,8); return u;}
static inline double asd(uint64_t u){double d; memcpy(&d,&u,8); return d;}
static float my_expf(float x){
  const double InvLn2N = 0x1.71547652b82fep+0 * 32;
  const double SHIFT = 0x1.8p+52;
  const double C0 = 0x1.c6af84b912394p-5/32/32/32, C1 = 0x1.ebfce50fac4f3p-3/32/32, C2 = 0x1.62e42ff0c52d6p-1/32;
  double xd = x;
  if (!(x > -0x1.9fe368p6f)) return (x!=x)? x : 0.0f;   /* underflow to 0 */
  if (x > 0x1.62e42ep6f) return INFINITY;
  double z = InvLn2N * xd;
  double kd = z + SHIFT; uint64_t ki = asu64(kd); kd -= SHIFT;
  double r = z - kd;
  uint64_t t = T[ki % 32]; t += ki << (52 - 5);
  double s = asd(t);
  z = C0 * r + C1; double r2 = r*r; double y = C2 * r + 1; y = z * r2 + y; y = y * s;
  return (float)y;
}
int main(){
  long bad=0, tot=0; 
  for (uint32_t u = 0; u < 0xffffffffu - 1000; u += 97) { float x; memcpy(&x,&u,4); if (x!=x) continue; if (fabsf(x) > 120) continue;
    float a = expf(x), b = my_expf(x); tot++; if (memcmp(&a,&b,4)) { bad++; if (bad < 10) printf("x=%a libm=%a mine=%a\n", x, a, b);} }
  printf("tot=%ld bad=%ld\n", tot, bad);
}
