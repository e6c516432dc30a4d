// Check of ddiv_by_const (ptl_physics.cuh): x / b by RN(1/b) and two fused residual steps equals the IEEE quotient.
// gcc -O2 -ffp-contract=off -o ddiv_check scripts/ddiv_check.c -lm && ./ddiv_check   (4e8 random arguments, 4 divisors: 0 mismatches)
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
static uint64_t s=88172645463325252ULL; static uint64_t rnd(){ s^=s<<13; s^=s>>7; s^=s<<17; return s; }
int main(){
  double bs[4]={262144000.0*1.602176634e-19, 3.0, 1.9999999999999998, 4.199999e-11};
  long bad1=0,bad2=0,n=0;
  for(int bi=0;bi<4;bi++){ double b=bs[bi], y=1.0/b;
   for(long i=0;i<100000000L;i++){
    uint64_t u=rnd(); double m=1.0+(double)(u>>12)*(1.0/4503599627370496.0); int e=(int)(rnd()%80)-70; double x=ldexp(m,e)*b*0.999;
    double q0=x*y; double r0=fma(-q0,b,x); double q1=fma(r0,y,q0); double r1=fma(-q1,b,x); double q2=fma(r1,y,q1);
    double q=x/b; n++; if(q1!=q) bad1++; if(q2!=q) bad2++; }
  }
  printf("n=%ld one-step mismatches=%ld two-step mismatches=%ld\n",n,bad1,bad2); return 0; }
