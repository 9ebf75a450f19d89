#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
static void rsqrt_dd(double x,double*hi,double*lo){
  double r=1.0/sqrt(x); double t=r*r, tl=fma(r,r,-t); double s=x*t, sl=fma(x,t,-s)+x*tl;
  double e=(1.0-s)-sl; *hi=r; *lo=0.5*r*e; }
static double pm05(double x){double h,l; rsqrt_dd(x,&h,&l); return h+l;}
static double pm15(double x){double h,l; rsqrt_dd(x,&h,&l); double q0=h/x; double rem=fma(-q0,x,h); double c=(rem+l)/x; return q0+c;}
int main(){
  uint64_t s=88172645463325252ULL; long n=20000000, bad1=0,bad2=0, naive1=0;
  for(long i=0;i<n;i++){
    s^=s<<13; s^=s>>7; s^=s<<17;
    double u=(double)(s>>11)/9007199254740992.0;
    double x=exp((u-0.5)*40.0);   /* 2e-9 .. 5e8 */
    s^=s<<13; s^=s>>7; s^=s<<17;
    x*= 1.0+ (double)(s>>11)/9007199254740992.0;
    double a=pow(x,-.5), b=pm05(x); if(a!=b) bad1++;
    if (a!=1.0/sqrt(x)) naive1++;
    double c=pow(x,-1.5), d=pm15(x); if(c!=d) bad2++;
  }
  printf("n=%ld mismatches pow(x,-.5): %ld  pow(x,-1.5): %ld  (naive 1/sqrt mismatches %ld)\n",n,bad1,bad2,naive1);
  return 0;}
