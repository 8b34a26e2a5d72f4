#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
static inline double exp_nonpos(double x) {
  const double L2E = 1.4426950408889634074, LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
  const double SHIFT = 6755399441055744.0;
  if (x < -708.0) return 0.0;
  double t = fma(x, L2E, SHIFT);
  int64_t tb; memcpy(&tb, &t, 8);
  int n = (int)(int32_t)(tb & 0xffffffff);
  double nf = t - SHIFT;
  double r = fma(nf, -LN2_HI, x);
  r = fma(nf, -LN2_LO, r);
  double p = 0x1.af4134720f354p-26;
  p = fma(p, r, 0x1.289876a2dbdc0p-22);
  p = fma(p, r, 0x1.71de0a0471800p-19);
  p = fma(p, r, 0x1.a019b31890abfp-16);
  p = fma(p, r, 0x1.a01a01a8ba744p-13);
  p = fma(p, r, 0x1.6c16c17a1c437p-10);
  p = fma(p, r, 0x1.1111111110871p-7);
  p = fma(p, r, 0x1.555555555394cp-5);
  p = fma(p, r, 0x1.5555555555556p-3);
  p = fma(p, r, 0x1.0000000000001p-1);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  int64_t pb; memcpy(&pb, &p, 8);
  pb += (int64_t)n << 52;
  memcpy(&p, &pb, 8);
  return p;
}
int main() {
  double worst = 0; double wx = 0;
  srand(1);
  for (long i = 0; i < 20000000; ++i) {
    double u = (double)rand() / RAND_MAX;
    double x = (i % 3 == 0) ? -u * 708.0 : (i % 3 == 1 ? -u * 40.0 : -u * 1.0);
    double a = exp_nonpos(x);
    long double b = expl((long double)x);
    double ulp = fabs((double)((long double)a - b)) / (double)(b * 1.1102230246251565e-16L);
    if (ulp > worst) { worst = ulp; wx = x; }
  }
  printf("max error %.3f ulp-units (of 2^-53 relative) at x=%.6f; exp_nonpos(0)=%.17g exp_nonpos(-1e-300)=%.17g\n", worst, wx, exp_nonpos(0.0), exp_nonpos(-1e-300));
  return 0;
}
