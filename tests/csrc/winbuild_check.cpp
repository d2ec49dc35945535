// Host check of the enumeration kernel's lane-window build for lite plans (gm_dfs_kernel, rnamotif_b200/csrc/gm_machine.cuh):
// eight positions at a time through expand8 against the per-nucleotide loop it replaces -- window bytes and base bitsets
// over random records, strands, contexts and window sizes (record edges, odd alignments, windows beyond the record).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <algorithm>
#include "gm_tilebits.h"
using namespace gm;
using std::min; using std::max;
static uint32_t funnelshift_r(uint32_t lo, uint32_t hi, unsigned sh){ sh&=31; return sh? (lo>>sh)|(hi<<(32-sh)) : lo; }
static uint32_t brev(uint32_t v){ uint32_t r=0; for(int i=0;i<32;i++) if(v>>i&1) r|=1u<<(31-i); return r; }
static uint8_t expand_code(unsigned c){ unsigned b = c == 1 ? 0 : c == 2 ? 1 : c == 4 ? 2 : c == 8 ? 3 : 4; return (uint8_t)(c | (b << 4)); }
static uint8_t complement_byte(uint8_t v){ unsigned b = v >> 4; if (b > 3) return (uint8_t)(15 | (4 << 4)); unsigned nb = 3 - b; return (uint8_t)((1u << nb) | (nb << 4)); }
static int bcode_of(int v){ return v>>4; }

int main(){
  srand(3);
  const int64_t total_nt = 100000;
  uint8_t *packed=(uint8_t*)calloc(total_nt/2+2048,1);
  for(int64_t i=0;i<total_nt;i++){ int r=rand()%100; unsigned c = r<96 ? (1u<<(rand()%4)) : (rand()%16); packed[i>>1] |= c<<((i&1)*4); }
  // garbage in the slack
  for(int i=0;i<1024;i++) packed[(total_nt+1)/2+i]=rand();
  long fails=0;
  for(int it=0; it<400000 && fails<5; it++){
    int W = 5 + rand()%200, Lc = rand()%12, Wtot = W+2*Lc;
    int nwl=((Wtot+31)>>5)+3;
    int64_t roff = (rand()%4==0)?0: rand()%(total_nt-10);
    int slen = 1 + rand()%std::min<int64_t>(3000, total_nt-roff);
    if (rand()%5==0) slen = std::min<int64_t>(total_nt-roff, 1+rand()%40);
    int comp = rand()&1;
    int szero = rand()%slen; if(rand()%4==0) szero = rand()%std::min(slen,8); if (rand()%4==0) szero = std::max(0, slen-1-rand()%8);
    int c00 = szero - Lc;
    int words=(Wtot+3)/4; if(!(words&1)) words++; int wstride=words*4;
    static uint8_t winA[2048], winB[2048]; static uint32_t bitA[4*64], bitB[4*64];
    memset(winA,0xee,sizeof winA); memset(winB,0xee,sizeof winB); memset(bitA,0,sizeof bitA); memset(bitB,0,sizeof bitB);
    // old
    for (int w = 0; w < nwl; w++) {
      uint32_t b0 = 0, b1 = 0, b2 = 0, b3 = 0; const int i0 = w << 5;
      for (int t = 0; t < 32 && i0 + t < Wtot; t++) {
        const int c = c00 + i0 + t; uint8_t v = (uint8_t)(4 << 4);
        if (c >= 0 && c < slen) { const int64_t gf = roff + (comp ? slen - 1 - c : c); const unsigned byte = packed[gf >> 1]; v = expand_code((byte >> ((gf & 1) * 4)) & 15); if (comp) v = complement_byte(v); }
        winA[i0 + t] = v; const int bc = bcode_of(v);
        b0 |= (uint32_t)(bc == 0) << t; b1 |= (uint32_t)(bc == 1) << t; b2 |= (uint32_t)(bc == 2) << t; b3 |= (uint32_t)(bc == 3) << t;
      }
      bitA[0*nwl+w]=b0; bitA[1*nwl+w]=b1; bitA[2*nwl+w]=b2; bitA[3*nwl+w]=b3;
    }
    // new
    const uint32_t *pw=(const uint32_t*)packed; const int64_t wmax=(total_nt>>3)+1; uint32_t *win32=(uint32_t*)winB;
    for (int w = 0; w < nwl; w++) {
      uint32_t b0 = 0, b1 = 0, b2 = 0, b3 = 0; const int i0 = w << 5;
      for (int g = 0; g < 4; g++) {
        const int t0 = i0 + 8 * g;
        if (t0 < Wtot) {
          const int c0 = c00 + t0;
          const int jlo = max(0, -c0), jhi = min(min(8, slen - c0), Wtot - t0);
          uint32_t x = 0;
          if (jhi > jlo) {
            const int64_t f0 = comp ? roff + (slen - 1 - (c0 + 7)) : roff + c0;
            const int64_t wi = f0 >> 3; const int sh = (int)(f0 & 7) * 4;
            x = funnelshift_r(pw[max((int64_t)0, min(wi, wmax))], pw[max((int64_t)0, min(wi + 1, wmax))], sh);
            if (jlo > 0 || jhi < 8) { const int nlo = comp ? 8 - jhi : jlo, nhi = comp ? 8 - jlo : jhi; x &= (nhi >= 8 ? ~0u : ((1u << (4 * nhi)) - 1u)) & (~0u << (4 * nlo)); }
          }
          uint32_t f0_, f1_, r0_, r1_, bits; expand8(x, f0_, f1_, r0_, r1_, bits);
          if (comp) { f0_ = r0_; f1_ = r1_; bits = brev(bits);
            if (jlo > 0 || jhi < 8) for (int jj = 0; jj < 8; jj++) if (jj < jlo || jj >= jhi) { uint32_t &wd = jj < 4 ? f0_ : f1_; wd = (wd & ~(0xffu << (8 * (jj & 3)))) | (0x40u << (8 * (jj & 3))); } }
          win32[t0 >> 2] = f0_; if (t0 + 4 < Wtot) win32[(t0 >> 2) + 1] = f1_;
          b0 |= (bits & 0xffu) << (8 * g); b1 |= ((bits >> 8) & 0xffu) << (8 * g); b2 |= ((bits >> 16) & 0xffu) << (8 * g); b3 |= (bits >> 24) << (8 * g);
        }
      }
      bitB[0*nwl+w]=b0; bitB[1*nwl+w]=b1; bitB[2*nwl+w]=b2; bitB[3*nwl+w]=b3;
    }
    bool bad = memcmp(winA,winB,Wtot)!=0 || memcmp(bitA,bitB,sizeof(uint32_t)*4*nwl)!=0;
    for(int i=wstride;i<wstride+16;i++) if(winB[i]!=0xee) bad=true;
    if(bad){ fails++; printf("MISMATCH it %d W %d Lc %d roff %lld slen %d comp %d szero %d\n",it,W,Lc,(long long)roff,slen,comp,szero);
      for(int i=0;i<Wtot;i++) if(winA[i]!=winB[i]){ printf("  byte %d: old %02x new %02x\n",i,winA[i],winB[i]); break; }
      for(int i=0;i<4*nwl;i++) if(bitA[i]!=bitB[i]){ printf("  bits word %d (set %d w %d): old %08x new %08x\n",i,i/nwl,i%nwl,bitA[i],bitB[i]); break; } }
  }
  printf(fails? "FAILED\n":"ok\n");
}
