// Host check of rf_div_rn_fixed (csrc/rf_common.cuh): the 5-step FMA quotient with y = RN(1/b) equals the IEEE
// division a / b bit for bit over the guarded ranges (|a| in [2^-20, 2^20], |b| in [2^-10, 2^10]), including
// divisors with all-ones / all-zeros significands.  usage: check_fixed_div [n_divisors] (250 000 dividends each)
#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
static uint64_t s=88172645463325252ull;
static inline uint64_t rnd(){s^=s<<13;s^=s>>7;s^=s<<17;return s;}
static inline float mk(uint32_t u){float f;memcpy(&f,&u,4);return f;}
int main(int argc,char**argv){ int nb = argc>1?atoi(argv[1]):4000;
  long bad=0,tot=0;
  for(int ib=0; ib<nb; ++ib){
    // b: random mantissa, exponent in [-10,10]
    uint32_t mb = rnd()&0x7fffff; int eb = (int)(rnd()%21)-10;
    if (ib==0){mb=0x7fffff;} if(ib==1){mb=0;} if(ib==2){mb=1;} if(ib==3){mb=0x7ffffe;}
    float b = mk(((uint32_t)(eb+127)<<23)|mb); if (rnd()&1) b=-b;
    volatile float y = 1.0f/b;
    for(int ia=0; ia<250000; ++ia){
      uint32_t ma = rnd()&0x7fffff; int ea=(int)(rnd()%41)-20;
      float a = mk(((uint32_t)(ea+127)<<23)|ma); if(rnd()&1) a=-a;
      if (fabsf(a) > 1048576.f) continue;
      float q0=a*y; float r0=fmaf(-b,q0,a); float q1=fmaf(r0,y,q0); float r1=fmaf(-b,q1,a); float q2=fmaf(r1,y,q1);
      volatile float ref=a/b; tot++;
      if (memcmp(&q2,(float*)&ref,4)) { if(bad<10) printf("bad a=%a b=%a got %a want %a\n",a,b,q2,ref); bad++; }
    }
  }
  printf("tot %ld bad %ld\n",tot,bad); return bad!=0;
}
