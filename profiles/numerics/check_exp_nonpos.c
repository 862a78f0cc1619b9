#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
static double exp_nonpos(double x){
  const double L2E = 1.4426950408889634074, LN2HI = 6.93147180369123816490e-01, LN2LO = 1.90821492927058770002e-10;
  double xc = x < -708.0 ? -708.0 : x;
  double n = rint(xc * L2E);
  double r = fma(n, -LN2HI, xc);
  r = fma(n, -LN2LO, r);
  double p = 1.0/6227020800.0;           /* 1/13! */
  p = fma(p, r, 1.0/479001600.0);
  p = fma(p, r, 1.0/39916800.0);
  p = fma(p, r, 1.0/3628800.0);
  p = fma(p, r, 1.0/362880.0);
  p = fma(p, r, 1.0/40320.0);
  p = fma(p, r, 1.0/5040.0);
  p = fma(p, r, 1.0/720.0);
  p = fma(p, r, 1.0/120.0);
  p = fma(p, r, 1.0/24.0);
  p = fma(p, r, 1.0/6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  int64_t bits = ((int64_t)n + 1023) << 52; double s; memcpy(&s, &bits, 8);
  double v = p * s;
  return x < -708.0 ? 0.0 : v;
}
int main(){ double worst=0, wref=0; srand(2);
 for (long i=0;i<30000000;i++){ double r = rand()/(double)RAND_MAX; double x;
   switch(i%4){case 0: x=-r; break; case 1: x=-708*r; break; case 2: x=-pow(10,-16*r); break; default: x=-40*r;}
   long double t = expl((long double)x); double got=exp_nonpos(x), ref=exp(x);
   double ulp = nextafter((double)t,INFINITY)-(double)t;
   double err=fabsl((long double)got-t)/ulp, e2=fabsl((long double)ref-t)/ulp;
   if(err>worst)worst=err; if(e2>wref)wref=e2; }
 printf("worst ulp: custom %.3f glibc %.3f; f(0)=%.17g f(-709)=%g f(-708)=%g\n",worst,wref,exp_nonpos(0.0),exp_nonpos(-709.0),exp_nonpos(-708.0)); }
