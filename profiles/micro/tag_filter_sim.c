// window-replay simulator for encode_l1 (Go flavour walk), counts probe statistics
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
static inline uint64_t ld64(const uint8_t*s,int p){uint64_t v;memcpy(&v,s+p,8);return v;}
static inline uint32_t ld32(const uint8_t*s,int p){uint32_t v;memcpy(&v,s+p,4);return v;}
static inline uint32_t hash6(uint64_t u){return (uint32_t)(((u<<16)*227718039650203ull)>>(64-15));}
static int extend8(const uint8_t*src,int s,int cand,int limit){
  while(s<=limit){uint64_t d=ld64(src,s)^ld64(src,cand); if(d){s+=__builtin_ctzll(d)>>3;break;} s+=8;cand+=8;} return s;}
#define MAXOFF ((2<<20)+65535)
static uint64_t st_batches, st_lanes, st_eq, st_tag[9], st_consumed, st_steps, st_never, st_fetch_now;
static uint64_t st_end_re, st_end_search, st_end_rep, st_cons_hit;
static uint32_t table[1<<15];
static int W=32;
static inline uint32_t tagof(uint32_t v){ return (v*2654435761u)>>24; }
static int wbase; 
static void begin_batch(const uint8_t*src,int n,int s,int rematch){
  wbase = rematch? s-2 : s; st_batches++;
  for(int l=0;l<W;l++){ int p=wbase+l; if(p+8>n||p<0) continue;
    int never = rematch? l<2 : (l==3);
    st_lanes++; if(never){st_never++;continue;}
    st_fetch_now++;
    uint32_t h=hash6(ld64(src,p)); int c=table[h];
    uint32_t a=ld32(src,p), b=ld32(src,c);
    if(a==b) st_eq++;
    uint32_t ta=tagof(a), tb=tagof(b);
    for(int t=1;t<=8;t++) if((ta>>(8-t))==(tb>>(8-t))) st_tag[t]++;
  }
}
int main(int argc,char**argv){
  FILE*f=fopen(argv[1],"rb"); int n=1<<20; if(argc>2) W=atoi(argv[2]);
  uint8_t*src=malloc(n+64); 
  for(int blk=0;blk<4;blk++){
  if(fread(src,1,n,f)!=(size_t)n) break; memset(src+n,0,64);
  memset(table,0,sizeof table);
  int sLimit=n-8; int nextEmit=0,s=1,repeat=1; int rematch=0; 
  int rep_snap=0,Rps=0;
  begin_batch(src,n,s,0);
  for(;;){
    if(rematch){
      for(;;){
        nextEmit=s; if(s>=sLimit) goto done;
        int L=s-wbase; if(L>=W){ st_end_re++; begin_batch(src,n,s,1); rep_snap=0; L=s-wbase; }
        st_steps++; st_consumed++;
        uint64_t x=ld64(src,s-2); uint32_t m2=hash6(x); x>>=16; uint32_t ch=hash6(x);
        int cand=table[ch]; table[m2]=s-2; table[ch]=s;
        if(s-cand>MAXOFF || (uint32_t)x!=ld32(src,cand)){ s++; rematch=0; break; }
        st_cons_hit++;
        repeat=s-cand; int base=s; s+=4; cand+=4; s=extend8(src,s,cand,n-8); rep_snap=1; Rps=base;
      }
    }
    // search step
    int t=s; int nextS=t+((t-nextEmit)>>6)+4; if(nextS>sLimit) goto done;
    int L=t-wbase; if(L+2>=W){ st_end_search++; begin_batch(src,n,s,0); rep_snap=0; L=0; }
    if(rep_snap && t+1-Rps>20){ st_end_rep++; begin_batch(src,n,s,0); rep_snap=0; L=0;}
    st_steps++;
    uint64_t cv=ld64(src,t);
    uint32_t h0=hash6(cv),h1=hash6(cv>>8),h2=hash6(cv>>16);
    int c0=table[h0],c1=table[h1]; table[h0]=t; table[h1]=t+1;
    if((uint32_t)(cv>>8)==ld32(src,t-repeat+1)){
      int base=t+1; for(int i=base-repeat; base>nextEmit&&i>0&&src[i-1]==src[base-1];){i--;base--;}
      int cand=t-repeat+5; s=t+5; s=extend8(src,s,cand,sLimit); nextEmit=s; 
      if(s>=sLimit) goto done; 
      // after repeat, kernel continues in the same batch with rep_snap unchanged
      continue;
    }
    int cand; int minp=t-MAXOFF;
    st_consumed++;
    if(c0>=minp && (uint32_t)cv==ld32(src,c0)){cand=c0; st_cons_hit++;}
    else { int c2=table[h2]; st_consumed++;
      if(c1>=minp && (uint32_t)(cv>>8)==ld32(src,c1)){ table[h2]=t+2; cand=c1; s=t+1; st_cons_hit++;}
      else { table[h2]=t+2; st_consumed++; if(c2>=minp && (uint32_t)(cv>>16)==ld32(src,c2)){cand=c2;s=t+2; st_cons_hit++;}
        else { s=nextS; continue; } } }
    int mps=s;
    while(cand>0&&s>nextEmit&&src[cand-1]==src[s-1]){cand--;s--;}
    int base=s; repeat=base-cand; s+=4;cand+=4; s=extend8(src,s,cand,n-8);
    rep_snap=1; Rps=mps; rematch=1;
  }
  done:;
  rematch=0;
  }
  printf("W=%d batches/blk %.0f steps/blk %.0f steps/batch %.2f\n",W,st_batches/4.0,st_steps/4.0,(double)st_steps/st_batches);
  printf("lanes/batch %.2f never %.2f fetched-now %.2f consumed %.2f (hits %.2f)\n",(double)st_lanes/st_batches,(double)st_never/st_batches,(double)st_fetch_now/st_batches,(double)st_consumed/st_batches,(double)st_cons_hit/st_batches);
  printf("eq4 %.2f/batch (%.1f%% of fetched)\n",(double)st_eq/st_batches,100.0*st_eq/st_fetch_now);
  for(int t=1;t<=8;t++) printf("  tag %d bits: fetch %.2f/batch (%.1f%%)\n",t,(double)st_tag[t]/st_batches,100.0*st_tag[t]/st_fetch_now);
  printf("batch ends: rematch-out %.1f%% search-out %.1f%% rep-snap %.1f%%\n",100.0*st_end_re/st_batches,100.0*st_end_search/st_batches,100.0*st_end_rep/st_batches);
}
