#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <string>
#include "parse_swar.cuh"
static uint64_t ref_fnv(const unsigned char* s,int n){uint64_t h=1469598103934665603ULL;for(int i=0;i<n;i++){h^=s[i];h*=1099511628211ULL;}return h;}
int main(){
  alignas(16) unsigned char buf[256]; uint32_t* w=(uint32_t*)buf; int fails=0; srand(1);
  for(int it=0;it<2000000;it++){
    memset(buf,'x',sizeof buf);
    int off=rand()%40; int len=1+rand()%10; char s[16]; long long val=0; bool bad=false;
    for(int i=0;i<len;i++){ int r=rand()%200; char c= r<196? '0'+rand()%10 : (r==196?'/':(r==197?':':(r==198?'\t':'a'))); s[i]=c; if(c<'0'||c>'9') bad=true; val=val*10+(c-'0'); }
    memcpy(buf+off,s,len);
    int32_t out=-1; bool ok=swar_parse_uint(w,off,len,&out);
    bool want = !bad && val<=2147483647LL;
    if(ok!=want || (ok && out!=(int32_t)val)){ if(fails<10) printf("uint fail off=%d len=%d s=%.*s ok=%d want=%d out=%d val=%lld\n",off,len,len,s,ok,want,out,val); fails++; }
  }
  const char* L="ACTG";
  for(int it=0;it<1000000;it++){
    int off=rand()%40; unsigned char c4[4]; bool good=true; unsigned code=0;
    for(int i=0;i<4;i++){ int r=rand()%50; unsigned char c = r<46? (unsigned char)"ACGT"[rand()%4] : (unsigned char)(32+rand()%90); c4[i]=c; const char* q=strchr(L,c); if(!q||c==0) good=false; else code=(code<<2)|(unsigned)(q-L); }
    memcpy(buf+off,c4,4);
    bool ok=true; unsigned got=swar_pack4(swar_get4(w,off),&ok);
    if(ok!=good || (good && got!=code)){ if(fails<20) printf("pack fail %c%c%c%c ok=%d good=%d got=%u code=%u\n",c4[0],c4[1],c4[2],c4[3],ok,good,got,code); fails++; }
  }
  for(int it=0;it<200000;it++){
    int off=rand()%40,len=1+rand()%30; for(int i=0;i<len;i++) buf[off+i]=33+rand()%90;
    if(swar_fnv1a(w,off,len)!=ref_fnv(buf+off,len)){ fails++; if(fails<30) printf("fnv fail\n"); }
  }
  for(int it=0;it<500000;it++){
    for(int i=0;i<16;i++) buf[i]= (rand()%5==0)? '\t' : ((rand()%7==0)?'\n':(unsigned char)(rand()%256));
    unsigned t=swar_eqmask16(w[0],w[1],w[2],w[3],0x09090909u), n=swar_eqmask16(w[0],w[1],w[2],w[3],0x0a0a0a0au);
    unsigned wt=0,wn=0; for(int i=0;i<16;i++){ if(buf[i]=='\t') wt|=1u<<i; if(buf[i]=='\n') wn|=1u<<i; }
    if(t!=wt||n!=wn){ fails++; if(fails<40) printf("mask fail %x %x %x %x\n",t,wt,n,wn); }
  }
  // contig keys: equal names <=> equal keys for short names; the bytes after the name must not leak in
  for(int it=0;it<500000;it++){
    for(int i=0;i<64;i++) buf[i]=(unsigned char)(33+rand()%90);
    int off=rand()%20,len=rand()%13;
    uint64_t k1=swar_contig_key(w,off,len);
    unsigned char save[64]; memcpy(save,buf,64);
    for(int i=off+len;i<64;i++) buf[i]=(unsigned char)rand();           // different context after the name
    for(int i=0;i<off;i++) buf[i]=(unsigned char)rand();
    uint64_t k2=swar_contig_key(w,off,len);
    if(k1!=k2){ fails++; if(fails<50) printf("contig key context leak off=%d len=%d\n",off,len); }
    if(len>=1&&len<=8){
      uint64_t want=0; for(int i=0;i<len;i++) want|=(uint64_t)buf[off+i]<<(8*i);
      if(k1!=want){ fails++; if(fails<50) printf("contig key value off=%d len=%d\n",off,len); }
      buf[off+rand()%len]^=1; if(swar_contig_key(w,off,len)==k1){ fails++; if(fails<50) printf("contig key collision\n"); }
    }
  }
  printf("fails=%d\n",fails); return fails!=0;
}
