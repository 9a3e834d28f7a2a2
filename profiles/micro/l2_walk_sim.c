#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
static inline uint64_t ld64(const uint8_t*s,int p){uint64_t v;memcpy(&v,s+p,8);return v;}
static inline uint32_t ld32(const uint8_t*s,int p){uint32_t v;memcpy(&v,s+p,4);return v;}
static inline uint32_t hash7(uint64_t u,int h){return (uint32_t)(((u<<8)*58295818150454627ull)>>(64-h));}
static inline uint32_t hash4(uint64_t u,int h){return ((uint32_t)u*2654435761u)>>(32-h);}
#define MAXOFF ((2<<20)+65535)
static int ext_tail(const uint8_t*src,int n,int s,int c){ while(s<=n-8){uint64_t d=ld64(src,s)^ld64(src,c); if(d){return s+(__builtin_ctzll(d)>>3);} s+=8;c+=8;} while(s<n&&src[s]==src[c]){s++;c++;} return s;}
static uint32_t lT[1<<17], sT[1<<14];
uint64_t steps,miss,m_long8,m_rep,m_long4,m_short,m_longp1, fwdhist[5], backhist[4], inserts, far_drop;
uint64_t windows, w_steps; int W=16;
int main(int argc,char**argv){ FILE*f=fopen(argv[1],"rb"); int n=1<<20; uint8_t*src=malloc(n+64); if(argc>2)W=atoi(argv[2]);
 int nb=0;
 for(int b=0;b<4;b++){ if(fread(src,1,n,f)!=(size_t)n)break; nb++; memset(src+n,0,64); memset(lT,0,sizeof lT); memset(sT,0,sizeof sT);
  int sLimit=n-8,nextEmit=0,s=1,repeat=1; uint64_t cv=ld64(src,s);
  int wbase=s; windows++;
  for(;;){ int cand=0,nextS=0; int probe=0; int src_kind=0;
    for(;;){ nextS=s+((s-nextEmit)>>7)+1; if(nextS>sLimit) goto done; steps++;
      if(s-wbase+1>=W){wbase=s;windows++;} w_steps++;
      int minp=s-MAXOFF+1; uint32_t hL=hash7(cv,17),hS=hash4(cv,14); cand=lT[hL]; int cS=sT[hS]; lT[hL]=s; sT[hS]=s; inserts+=2;
      uint64_t vL=ld64(src,cand),vS=ld64(src,cS);
      if(cand>minp&&cv==vL){m_long8++;src_kind=1;break;}
      uint64_t rm=0xffffffffull<<8;
      if(repeat>0&&(cv&rm)==(ld64(src,s-repeat)&rm)){ m_rep++; int base=s+1; for(int i=base-repeat;base>nextEmit&&i>0&&src[i-1]==src[base-1];){i--;base--;}
         int c=s-repeat+5; s+=5; s=ext_tail(src,n,s,c); nextEmit=s; if(s>=sLimit) goto done;
         int i0=base+1,i1=s-2; while(i0<i1){ inserts+=4; uint64_t c0=ld64(src,i0),c1=ld64(src,i1); lT[hash7(c0,17)]=i0; sT[hash4(c0>>8,14)]=i0+1; lT[hash7(c1,17)]=i1; sT[hash4(c1>>8,14)]=i1+1; i0+=2;i1-=2;}
         cv=ld64(src,s); if(s-wbase+1>=W){wbase=s;windows++;} continue; }
      if(cand>=minp&&(uint32_t)cv==(uint32_t)vL){m_long4++;src_kind=2;break;}
      if(cS>=minp&&(uint32_t)cv==(uint32_t)vS){ hL=hash7(cv>>8,17); cand=lT[hL]; lT[hL]=s+1; inserts++;
         if(cand>minp&&(uint32_t)(cv>>8)==ld32(src,cand)){s++;m_longp1++;src_kind=4;break;} cand=cS; m_short++;src_kind=3;break;}
      miss++; cv=ld64(src,nextS); s=nextS; }
    probe=s; int back=0; while(cand>0&&s>nextEmit&&src[cand-1]==src[s-1]){cand--;s--;back++;}
    backhist[back>3?3:back]++;
    int base=s,offset=base-cand; s+=4;cand+=4; s=ext_tail(src,n,s,cand);
    int fwd=s-probe; fwdhist[fwd<8?0:fwd<12?1:fwd<16?2:fwd<28?3:4]++;
    if(offset>65535&&s-base<=4&&repeat!=offset){ far_drop++; s=nextS+1; if(s>=sLimit)goto done; cv=ld64(src,s); if(s-wbase+1>=W){wbase=s;windows++;} continue;}
    repeat=offset; nextEmit=s; if(s>=sLimit) goto done;
    { int i0=base+1,i1=s-2; uint64_t c0=ld64(src,i0),c1=ld64(src,i1); lT[hash7(c0,17)]=i0; sT[hash4(c0>>8,14)]=i0+1; lT[hash7(c1,17)]=i1; sT[hash4(c1>>8,14)]=i1+1; inserts+=4; i0++;i1--; cv=ld64(src,s);
      int i2=(i0+i1+1)>>1; while(i2<i1){ lT[hash7(ld64(src,i0),17)]=i0; lT[hash7(ld64(src,i2),17)]=i2; inserts+=2; i0+=2;i2+=2;} }
    if(s-wbase+1>=W){wbase=s;windows++;}
  }
  done:;
 }
 double k=nb;
 printf("per block: steps %.0f miss %.0f | long8 %.0f rep %.0f long4 %.0f short %.0f long+1 %.0f far_drop %.0f | inserts %.0f\n",steps/k,miss/k,m_long8/k,m_rep/k,m_long4/k,m_short/k,m_longp1/k,far_drop/k,inserts/k);
 printf("fwd len from probe: <8 %.0f  8-11 %.0f  12-15 %.0f 16-27 %.0f  >=28 %.0f\n",fwdhist[0]/k,fwdhist[1]/k,fwdhist[2]/k,fwdhist[3]/k,fwdhist[4]/k);
 printf("back: 0 %.0f 1 %.0f 2 %.0f 3+ %.0f\n",backhist[0]/k,backhist[1]/k,backhist[2]/k,backhist[3]/k);
 printf("windows(W=%d) per block %.0f  steps/window %.2f\n",W,windows/k,(double)w_steps/windows);
}
