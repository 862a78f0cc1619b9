#include <math.h>
#include <stdio.h>
#include <stdlib.h>
static double f(double e){ volatile double u = 1.0 + e; volatile double um1 = u - 1.0; double c = e - um1; return log(u) + c * (double)(1.0f/(float)u); }
int main(){ double worst=0, worst_ref=0; srand(1);
 for (long i=0;i<20000000;i++){ double r = rand()/(double)RAND_MAX; double e;
   switch(i%4){case 0: e=r; break; case 1: e=pow(10,-16*r); break; case 2: e = ldexp(r,-30); break; default: e = exp(-fabs(40*(r-0.5))); }
   long double t = log1pl((long double)e); double got=f(e), ref=log1p(e);
   double ulp = nextafter(fabs((double)t),INFINITY)-fabs((double)t); if (ulp==0) continue;
   double err = fabsl((long double)got - t)/ulp, err2 = fabsl((long double)ref - t)/ulp;
   if (err>worst) worst=err; if (err2>worst_ref) worst_ref=err2; }
 printf("worst ulp err: fast %.3f, glibc log1p %.3f; f(0)=%g\n", worst, worst_ref, f(0.0)); return 0; }
