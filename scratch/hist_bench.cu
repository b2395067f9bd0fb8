// Microbenchmark: which scatter-add mechanism sustains the Cox pass-1 histogram on B200?
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <random>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
constexpr int NB = 4096;
__device__ __forceinline__ float4 ld4(const float* p){ float4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];":"=f"(r.x),"=f"(r.y),"=f"(r.z),"=f"(r.w):"l"(p)); return r;}
__device__ __forceinline__ uint32_t ld1(const void* p){ uint32_t r; asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];":"=r"(r):"l"(p)); return r;}

template<int V> __device__ __forceinline__ void row(float eta, float t, unsigned ev, float* hf, unsigned* hu, unsigned long long* h64, float* gf, unsigned* gu, float& acc){
  float w = __expf(eta); int b = (int)t; b = min(max(b,0),NB-1);
  if (V==0) { acc += w + (ev? 1.f:0.f) + b; }
  if (V==1) { // CAS-loop float (+ packed count for events)
    if (ev) { unsigned long long* p = h64 + b; unsigned long long old=*p, as; do { as=old; float s=__uint_as_float((unsigned)as)+w; unsigned long long m=(as>>32)+1ull; old=atomicCAS(p, as, (m<<32)|(unsigned long long)__float_as_uint(s)); } while(old!=as); }
    else atomicAdd(hf + b, w);
  }
  if (V==2) { // native u32 add, one per row (throughput probe; fixed point 2^20)
    unsigned q = (unsigned)(w * 1048576.f); atomicAdd(hu + (ev? NB:0) + b, q); if (ev) atomicAdd(hu + 2*NB + b, 1u);
  }
  if (V==3) { // 64-bit fixed point as two native u32 adds with carry
    unsigned long long q = (unsigned long long)((double)w * 1099511627776.0); unsigned lo=(unsigned)q, hi=(unsigned)(q>>32);
    unsigned* base = hu + (ev? 2*NB:0) + 2*b; unsigned old = atomicAdd(base, lo); hi += (old + lo < old); if (hi) atomicAdd(base+1, hi); if (ev) atomicAdd(hu + 4*NB + b, 1u);
  }
  if (V==4) { // global float RED
    atomicAdd(gf + (ev? NB:0) + b, w); if (ev) atomicAdd(gu + b, 1u);
  }
  if (V==5) { // hybrid: events -> global RED, censored -> smem CAS float
    if (ev) { atomicAdd(gf + NB + b, w); atomicAdd(gu + b, 1u);} else atomicAdd(hf + b, w);
  }
  if (V==6) { // hybrid: events -> smem native (fixed32 + count), censored -> global RED
    if (ev) { unsigned q=(unsigned)(w*1048576.f); atomicAdd(hu+b,q); atomicAdd(hu+NB+b,1u);} else atomicAdd(gf+b, w);
  }
  if (V==7) { // non-atomic racy RMW in smem (upper bound on plain LDS/STS speed; WRONG results)
    float* p = hf + (ev? NB:0) + b; *p = *p + w;
  }
}
template<int V> __global__ void __launch_bounds__(1024,1) k(const float* __restrict__ lh, const float* __restrict__ tm, const uint8_t* __restrict__ evp, long long n, float* gf, unsigned* gu, float* out){
  extern __shared__ __align__(16) unsigned char sm[];
  float* hf = (float*)sm; unsigned* hu=(unsigned*)sm; unsigned long long* h64=(unsigned long long*)(sm + 4*NB*2);
  for (int i=threadIdx.x;i<NB*6;i+=blockDim.x) hu[i]=0;
  __syncthreads();
  float acc=0;
  long long ng=n/4, stride=(long long)gridDim.x*blockDim.x, g=(long long)blockIdx.x*blockDim.x+threadIdx.x;
  for (; g+stride<ng; g+=2*stride){ long long g2=g+stride;
    float4 e0=ld4(lh+4*g), e1=ld4(lh+4*g2), t0=ld4(tm+4*g), t1=ld4(tm+4*g2); uint32_t v0=ld1(evp+4*g), v1=ld1(evp+4*g2);
    row<V>(e0.x,t0.x,v0&0xff,hf,hu,h64,gf,gu,acc); row<V>(e0.y,t0.y,v0&0xff00,hf,hu,h64,gf,gu,acc); row<V>(e0.z,t0.z,v0&0xff0000,hf,hu,h64,gf,gu,acc); row<V>(e0.w,t0.w,v0&0xff000000,hf,hu,h64,gf,gu,acc);
    row<V>(e1.x,t1.x,v1&0xff,hf,hu,h64,gf,gu,acc); row<V>(e1.y,t1.y,v1&0xff00,hf,hu,h64,gf,gu,acc); row<V>(e1.z,t1.z,v1&0xff0000,hf,hu,h64,gf,gu,acc); row<V>(e1.w,t1.w,v1&0xff000000,hf,hu,h64,gf,gu,acc);
  }
  if (g<ng){ float4 e0=ld4(lh+4*g), t0=ld4(tm+4*g); uint32_t v0=ld1(evp+4*g);
    row<V>(e0.x,t0.x,v0&0xff,hf,hu,h64,gf,gu,acc); row<V>(e0.y,t0.y,v0&0xff00,hf,hu,h64,gf,gu,acc); row<V>(e0.z,t0.z,v0&0xff0000,hf,hu,h64,gf,gu,acc); row<V>(e0.w,t0.w,v0&0xff000000,hf,hu,h64,gf,gu,acc);}
  __syncthreads();
  float s=acc; for (int i=threadIdx.x;i<NB*6;i+=blockDim.x) s+= (float)hu[i];
  if (s==123.456f) out[0]=s;  // keep alive
  // flush cost is part of the real kernel: plain stores of 3*NB words
  for (int i=threadIdx.x;i<NB*3;i+=blockDim.x) out[(size_t)blockIdx.x*NB*3+i]=hf[i];
}
template<int V> void run(const char* name,const float* lh,const float* tm,const uint8_t* ev,long long n,float* gf,unsigned* gu,float* out,int grid){
  CK(cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, NB*24));
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  for(int i=0;i<2;i++) k<V><<<grid,1024,NB*24>>>(lh,tm,ev,n,gf,gu,out);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a); for(int i=0;i<5;i++) k<V><<<grid,1024,NB*24>>>(lh,tm,ev,n,gf,gu,out); cudaEventRecord(b); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms,a,b); ms/=5; printf("%-44s grid %4d: %8.1f us  %7.1f GB/s (9B/row)\n",name,grid,ms*1e3, 9.0*n/ms/1e6);
}
int main(){
  long long n=1<<24; std::vector<float> lh(n),tm(n); std::vector<uint8_t> ev(n); std::mt19937 r(1); std::normal_distribution<float> nd; std::exponential_distribution<float> ed(1.f/1000); std::uniform_real_distribution<float> ud;
  for(long long i=0;i<n;i++){ lh[i]=nd(r); float t=floorf(ed(r)); tm[i]=fminf(fmaxf(t,1),4000); ev[i]=ud(r)<0.3f; }
  float *dl,*dt,*gf,*out; uint8_t* de; unsigned* gu; CK(cudaMalloc(&dl,n*4)); CK(cudaMalloc(&dt,n*4)); CK(cudaMalloc(&de,n)); CK(cudaMalloc(&gf,NB*8*4)); CK(cudaMalloc(&gu,NB*4*4)); CK(cudaMalloc(&out,(size_t)1024*NB*3*4));
  cudaMemcpy(dl,lh.data(),n*4,cudaMemcpyHostToDevice); cudaMemcpy(dt,tm.data(),n*4,cudaMemcpyHostToDevice); cudaMemcpy(de,ev.data(),n,cudaMemcpyHostToDevice); cudaMemset(gf,0,NB*8*4); cudaMemset(gu,0,NB*4*4);
  for (int grid : {148, 296}) {
    if (grid==296) printf("-- (grid 296 only runs 1 CTA/SM at a time with 1024 threads + 96KB smem)\n");
    run<0>("V0 stream only",dl,dt,de,n,gf,gu,out,grid);
    run<1>("V1 smem CAS float (+packed count)",dl,dt,de,n,gf,gu,out,grid);
    run<2>("V2 smem native u32 add",dl,dt,de,n,gf,gu,out,grid);
    run<3>("V3 smem 2x native u32 (64b fixed)",dl,dt,de,n,gf,gu,out,grid);
    run<4>("V4 global RED float",dl,dt,de,n,gf,gu,out,grid);
    run<5>("V5 events->global RED, cens->smem CAS",dl,dt,de,n,gf,gu,out,grid);
    run<6>("V6 events->smem native, cens->global RED",dl,dt,de,n,gf,gu,out,grid);
    run<7>("V7 racy non-atomic smem RMW (bound)",dl,dt,de,n,gf,gu,out,grid);
  }
  return 0;
}
