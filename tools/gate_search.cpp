// gate_search.cpp -- design-space exploration behind the DP-cell variants of apm_sliced.cuh (not product code).
// Exhaustive search for a circuit computing one bit-sliced DP cell (inputs eq, vertical delta a, horizontal delta b;
// outputs v = x - b, h = x - a with x = min(1 - eq, a + 1, b + 1)) from L three-input boolean gates (LOP3) and F
// bitwise-exact ADD/SUB operations (FMA pipe), over all injective 2-bit encodings of the deltas.
//   g++ -O2 -o gate_search gate_search.cpp && ./gate_search Lmax Fmax Imax [part nparts]
// Findings (round 1): (L,F) = (4,0) and (4,1): no circuit; (5,0): the shipped 5-LOP3 cell; (4,2): circuits exist on
// the encodings a = (plus, zero) / b = (minus, nonzero) and their mirror images -- the mirrored one (row recurrence
// through v- only) is CELL 2 of apm_sliced.cuh; (4,3) on the (plus, minus) encoding is CELL 1 (found by hand);
// (3,3) with <= 2 intermediates: none found in 3 x 12 CPU-minutes per third of the encoding space (not exhaustive).
#include <cstdio>
#include <cstdint>
#include <vector>
#include <array>
#include <string>
#include <algorithm>
#include <cstring>
#include <cstdlib>
using namespace std;
typedef uint32_t u32;
static int R; static u32 ALL;
struct Op { int type; int a,b,c; u32 res; int out; }; // type 0=L,1=ADD,2=SUB ; out = output index or -1
static int Lmax, Fmax, Imax;
static vector<u32> sig; static vector<Op> ops; static u32 tgt[4];
static long long nodes=0; static bool found=false; static vector<Op> best;
static const int ST[3]={-1,0,1};

static bool lop_ok(u32 s1,u32 s2,u32 s3,u32 T){
  for(int p=0;p<8;p++){ u32 M=ALL; M&=(p&4)?s1:~s1; M&=(p&2)?s2:~s2; M&=(p&1)?s3:~s3; M&=ALL; u32 x=T&M; if(x!=0&&x!=M) return false;} return true; }
static bool find_F(u32 T,int&ty,int&ia,int&ib){ int n=sig.size();
  for(int i=0;i<n;i++)for(int j=0;j<n;j++){ if(i==j)continue; u32 a=sig[i],b=sig[j];
    if(i<j && (a&b)==0 && (a|b)==T){ty=1;ia=i;ib=j;return true;}
    if((b&~a)==0 && (a&~b)==T){ty=2;ia=i;ib=j;return true;} }
  return false; }
static bool find_L(u32 T,int&ia,int&ib,int&ic){ int n=sig.size();
  for(int i=0;i<n;i++)for(int j=i+1;j<n;j++)for(int k=j+1;k<n;k++) if(lop_ok(sig[i],sig[j],sig[k],T)){ia=i;ib=j;ic=k;return true;}
  return false; }
static bool has(u32 s){ for(u32 x:sig) if(x==s) return true; return false; }

static void dfs(int done,int L,int F,int I){
  if(found) return; nodes++;
  if(done==15){ found=true; best=ops; return; }
  int rem=4-__builtin_popcount(done);
  if((Lmax-L)+(Fmax-F)<rem) return;
  // outputs
  for(int o=0;o<4;o++) if(!(done>>o&1)){
    if(has(tgt[o])){ dfs(done|1<<o,L,F,I); return; }
  }
  for(int o=0;o<4;o++) if(!(done>>o&1)){
    int ty,a,b,c;
    if(F<Fmax && find_F(tgt[o],ty,a,b)){ sig.push_back(tgt[o]); ops.push_back({ty,a,b,-1,tgt[o],o}); dfs(done|1<<o,L,F+1,I); ops.pop_back(); sig.pop_back(); if(found) return; }
    else if(L<Lmax && find_L(tgt[o],a,b,c)){ sig.push_back(tgt[o]); ops.push_back({0,a,b,c,tgt[o],o}); dfs(done|1<<o,L+1,F,I); ops.pop_back(); sig.pop_back(); if(found) return; }
  }
  if(I>=Imax) return;
  if((Lmax-L)+(Fmax-F)<=rem) return;
  int n=sig.size();
  // F intermediates
  if(F<Fmax){
    for(int i=0;i<n;i++)for(int j=0;j<n;j++){ if(i==j)continue; u32 a=sig[i],b=sig[j];
      if(i<j && (a&b)==0){ u32 r=a|b; if(r!=ALL && !has(r)){ sig.push_back(r); ops.push_back({1,i,j,-1,r,-1}); dfs(done,L,F+1,I+1); ops.pop_back(); sig.pop_back(); if(found) return; } }
      if((b&~a)==0){ u32 r=a&~b; if(r!=0 && !has(r)){ sig.push_back(r); ops.push_back({2,i,j,-1,r,-1}); dfs(done,L,F+1,I+1); ops.pop_back(); sig.pop_back(); if(found) return; } } }
  }
  if(L<Lmax){
    static thread_local vector<uint8_t> seen; // per level dedupe is expensive for 2^18; use local vector
    vector<u32> cand;
    for(int i=0;i<n;i++)for(int j=i+1;j<n;j++)for(int k=j+1;k<n;k++){
      u32 pm[8]; int np=0;
      for(int p=0;p<8;p++){ u32 M=ALL; M&=(p&4)?sig[i]:~sig[i]; M&=(p&2)?sig[j]:~sig[j]; M&=(p&1)?sig[k]:~sig[k]; M&=ALL; if(M) pm[np++]=M; }
      for(int f=1;f<(1<<np)-1;f++){ u32 r=0; for(int q=0;q<np;q++) if(f>>q&1) r|=pm[q]; cand.push_back(r); }
    }
    sort(cand.begin(),cand.end()); cand.erase(unique(cand.begin(),cand.end()),cand.end());
    for(u32 r:cand){ if(has(r)) continue; sig.push_back(r); ops.push_back({0,-1,-1,-1,r,-1}); dfs(done,L+1,F,I+1); ops.pop_back(); sig.pop_back(); if(found) return; }
  }
}

int main(int argc,char**argv){
  Lmax=atoi(argv[1]); Fmax=atoi(argv[2]); Imax=atoi(argv[3]);
  int part=argc>4?atoi(argv[4]):0, nparts=argc>5?atoi(argv[5]):1;
  int perms[24][3]; int np=0; { int c[4]={0,1,2,3}; // all injective maps state->code
    for(int x=0;x<4;x++)for(int y=0;y<4;y++)for(int z=0;z<4;z++) if(x!=y&&y!=z&&x!=z){perms[np][0]=x;perms[np][1]=y;perms[np][2]=z;np++;} }
  R=18; ALL=(1u<<R)-1;
  int idx=0;
  for(int ea=0;ea<24;ea++)for(int eb=0;eb<24;eb++){
    if((idx++)%nparts!=part) continue;
    // rows
    u32 in[5]={0,0,0,0,0}; u32 out[4]={0,0,0,0}; int r=0;
    for(int eq=0;eq<2;eq++)for(int a=0;a<3;a++)for(int b=0;b<3;b++,r++){
      int A=ST[a],B=ST[b]; int x=min(1-eq,min(A+1,B+1)); int v=x-B,h=x-A;
      int ca=perms[ea][a], cb=perms[eb][b], cv=perms[ea][v+1], ch=perms[eb][h+1];
      if(eq) in[0]|=1u<<r; if(ca&2) in[1]|=1u<<r; if(ca&1) in[2]|=1u<<r; if(cb&2) in[3]|=1u<<r; if(cb&1) in[4]|=1u<<r;
      if(cv&2) out[0]|=1u<<r; if(cv&1) out[1]|=1u<<r; if(ch&2) out[2]|=1u<<r; if(ch&1) out[3]|=1u<<r;
    }
    sig.assign(in,in+5); sig.push_back(ALL); for(int o=0;o<4;o++) tgt[o]=out[o];
    ops.clear(); found=false; nodes=0;
    dfs(0,0,0,0);
    if(found){
      printf("FOUND ea=[%d %d %d] eb=[%d %d %d] L<=%d F<=%d nodes=%lld\n",perms[ea][0],perms[ea][1],perms[ea][2],perms[eb][0],perms[eb][1],perms[eb][2],Lmax,Fmax,nodes);
      for(auto&o:best) printf("   type=%d a=%d b=%d c=%d res=%05x out=%d\n",o.type,o.a,o.b,o.c,o.res,o.out);
      fflush(stdout);
    }
  }
  fprintf(stderr,"part %d done\n",part);
}
