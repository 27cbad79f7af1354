#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32; typedef uint64_t u64;
__device__ __forceinline__ u32 csub(u32 x, u32 c){u32 y=x-c; return y<x?y:x;}
__device__ __forceinline__ u32 add_alu(u32 a,u32 b){return __viaddmax_u32(a,b,0u);}
__device__ __forceinline__ u32 mulw_i(u32 x, uint2 w, u32 p){ return x*w.x - __umulhi(x,w.y)*p; }
// DFMA quotient: c = k*2^-50 ~ w/p, K = 2^52 - 2^52 c
__device__ __forceinline__ u32 mulw_d(u32 x, u32 w, double c, double K, u32 p, u32 negp){
  double xd = __hiloint2double(0x43300000, (int)x);
  double Q = __fma_rn(xd, c, K);
  u32 q = (u32)__double2loint(Q);
  return q*negp + (x*w + p);
}
__device__ __forceinline__ u32 mulw_f(u32 x, u32 w, double c, u32 p, u32 negp){
  double xd = __uint2double_rn(x);
  double Q = __fma_rn(xd, c, 4503599627370496.0);
  u32 q = (u32)__double2loint(Q);
  return q*negp + (x*w + p);
}
template<int KIND>
__global__ void __launch_bounds__(512,1) k_bf(u32* out, u32 p, u32 w, u32 wq, double c, double K, int iters){
  u32 x[8];
  #pragma unroll
  for(int k=0;k<8;++k) x[k]=(threadIdx.x*8+k+blockIdx.x*977)%p;
  const u32 p2=2*p; u32 negp=0u-p; asm volatile("mov.u32 %0, %0;" : "+r"(negp));
  uint2 W=make_uint2(w,wq);
  for(int it=0;it<iters;++it){
    #pragma unroll
    for(int s=0;s<3;++s){
      const int st = 4>>s;
      #pragma unroll
      for(int j=0;j<8;++j){
        if(!(j&st)){
          u32 X=x[j],Y=x[j+st];
          u32 s_=add_alu(X,Y), d_=X+p2-Y;
          x[j]=csub(s_,p2);
          if(KIND==0) x[j+st]=mulw_i(d_,W,p);
          else if(KIND==1) x[j+st]=mulw_d(d_,w,c,K,p,negp);
          else if(KIND==2) x[j+st]=mulw_f(d_,w,c,p,negp);
          else x[j+st]= (j&1) ? mulw_f(d_,w,c,p,negp) : mulw_i(d_,W,p);
        }
      }
    }
  }
  u32 s=0;
  #pragma unroll
  for(int k=0;k<8;++k) s^=x[k];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void k_check_f(u32 p, u32 w, double c, u32* bad, u32* mx, u32* mn){
  u32 x = blockIdx.x*blockDim.x+threadIdx.x; u32 negp=0u-p;
  for(u64 v=x; v<(1ull<<32); v+= (u64)gridDim.x*blockDim.x){
    u32 r = mulw_f((u32)v,w,c,p,negp);
    u32 e = (u32)(((u64)(u32)v*w)%p);
    if(r>=2*p || (r%p)!=e) atomicAdd(bad,1);
    atomicMax(mx,r); atomicMin(mn,r);
  }
}
__global__ void k_check(u32 p, u32 w, u32 wq, double c, double K, u32* bad, u32* mx){
  u32 x = blockIdx.x*blockDim.x+threadIdx.x; // all 2^32 via grid-stride
  u32 negp=0u-p;
  for(u64 v=x; v<(1ull<<32); v+= (u64)gridDim.x*blockDim.x){
    u32 r = mulw_d((u32)v,w,c,K,p,negp);
    u32 e = (u32)(((u64)(u32)v*w)%p);
    if(r>=2*p || (r%p)!=e) atomicAdd(bad,1);
    atomicMax(mx,r);
  }
}
int main(){
  u32 p=1073479681u; // some 30-bit prime-like (only modulus needs odd); use real prime
  // find prime = 1 mod 1024 below 2^30
  auto isp=[](u64 n){ if(n%2==0) return false; for(u64 d=3; d*d<=n; d+=2) if(n%d==0) return false; return true;};
  p = ((1u<<30)-1)/1024*1024+1; while(!isp(p)) p-=1024;
  u32 w = 0x2468ace1u % p; u32 wq=(u32)(((u64)w<<32)/p);
  unsigned __int128 kk = ((unsigned __int128)w<<50)/p; double c = (double)(u64)kk * ldexp(1.0,-50); double K = ldexp(1.0,52) - 4.0*(double)(u64)kk;
  printf("p=%u w=%u c=%.17g K=%.17g\n",p,w,c,K);
  u32 *out,*bad,*mx; cudaMalloc(&out, 148*8*512*4); cudaMalloc(&bad,4); cudaMalloc(&mx,4); cudaMemset(bad,0,4); cudaMemset(mx,0,4);
  k_check<<<148*8,256>>>(p,w,wq,c,K,bad,mx);
  u32 hb,hm; cudaMemcpy(&hb,bad,4,cudaMemcpyDeviceToHost); cudaMemcpy(&hm,mx,4,cudaMemcpyDeviceToHost);
  printf("check all 2^32 x: bad=%u max_r=%u (2p=%u) %s\n",hb,hm,2*p, cudaGetErrorString(cudaGetLastError()));
  // second w near p-1 and w=1
  for(u32 ww : {1u, p-1, p/2, 3u}){ u32 q2=(u32)(((u64)ww<<32)/p); unsigned __int128 k2=((unsigned __int128)ww<<50)/p; double c2=(double)(u64)k2*ldexp(1.0,-50), K2=ldexp(1.0,52)-4.0*(double)(u64)k2; cudaMemset(bad,0,4); cudaMemset(mx,0,4); k_check<<<148*8,256>>>(p,ww,q2,c2,K2,bad,mx); cudaMemcpy(&hb,bad,4,cudaMemcpyDeviceToHost); cudaMemcpy(&hm,mx,4,cudaMemcpyDeviceToHost); printf("w=%u bad=%u max_r=%u\n",ww,hb,hm);}  
  u32* mn; cudaMalloc(&mn,4);
  for(u32 ww : {1u, p-1, p/2, 3u, w, 0u}){ double c2=(double)ww/(double)p; cudaMemset(bad,0,4); cudaMemset(mx,0,4); cudaMemset(mn,0xff,4); k_check_f<<<148*8,256>>>(p,ww,c2,bad,mx,mn); u32 hn; cudaMemcpy(&hb,bad,4,cudaMemcpyDeviceToHost); cudaMemcpy(&hm,mx,4,cudaMemcpyDeviceToHost); cudaMemcpy(&hn,mn,4,cudaMemcpyDeviceToHost); printf("mulw_f w=%u bad=%u min_r=%u max_r=%u (p=%u)\n",ww,hb,hn,hm,p);}
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters=2000;
  for(int kind=0;kind<4;++kind) for(int thr : {128,256,512,768,1024}){
    float best=1e9;
    for(int rep=0;rep<4;++rep){
      cudaEventRecord(e0);
      if(kind==0) k_bf<0><<<148,thr>>>(out,p,w,wq,c,K,iters); else if(kind==1) k_bf<1><<<148,thr>>>(out,p,w,wq,c,K,iters); else if(kind==2) k_bf<2><<<148,thr>>>(out,p,w,wq,c,K,iters); else k_bf<3><<<148,thr>>>(out,p,w,wq,c,K,iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(rep&&ms<best)best=ms;
    }
    double bfly = 148.0*thr*12.0*iters; // 12 butterflies per iter per thread
    printf("kind=%d threads/SM=%d : %.3f ms  %.1f Gbutterfly/s  (%s)\n",kind,thr,best,bfly/best*1e-6,cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
