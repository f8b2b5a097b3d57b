#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
static uint64_t s=0x9E3779B97F4A7C15ull;
static inline uint64_t nx(){ uint64_t z=(s+=0x9E3779B97F4A7C15ull); z=(z^(z>>30))*0xBF58476D1CE4E5B9ull; z=(z^(z>>27))*0x94D049BB133111EBull; return z^(z>>31);}
static inline double mk(int emin,int emax){ uint64_t m=nx()&0xfffffffffffffull; int e=emin+(int)(nx()%(uint64_t)(emax-emin+1)); uint64_t b=((uint64_t)(e+1023)<<52)|m; if(nx()&1) b|=1ull<<63; double d; memcpy(&d,&b,8); return d;}
int main(){
  long bad1=0,bad2=0,n=0;
  for(long i=0;i<400000000L;++i){
    double a=mk(-60,60), b=fabs(mk(-40,40));
    if((i&7)==0){ /* adversarial: mantissas near all-ones / near powers of two */ uint64_t bb; memcpy(&bb,&b,8); bb|=0xffffffffff000ull; if(i&8) bb&=~0xfffffffffff00ull; memcpy(&b,&bb,8);}        
    double y=1.0/b;
    double q0=a*y; double r0=fma(-b,q0,a); double q1=fma(r0,y,q0);
    double r1=fma(-b,q1,a); double q2=fma(r1,y,q1);
    double q=a/b;
    if(q1!=q) ++bad1; if(q2!=q) ++bad2; ++n;
  }
  printf("n=%ld one-step mismatches=%ld two-step mismatches=%ld\n",n,bad1,bad2);
  return 0;
}
