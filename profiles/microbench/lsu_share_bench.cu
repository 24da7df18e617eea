// Micro-benchmark 3: do global REDs and shared-memory loads share one pipe on sm_100a?
// Every thread issues R vector REDs (2x2-quad pattern, L2-resident image) interleaved with S conflict-free 16-byte
// shared-memory loads.  If the two kinds of access went through independent pipes the time would be max(t_red, t_lds);
// if they share the SM's LSU/L1TEX data path it is their sum.  Backs the "K1 time ~ shared wavefronts + 0.75 x RED
// lanes" reading of profiles/r01/prof_k1_c4_v8.summary.txt.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lsu_share_bench lsu_share_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__device__ __forceinline__ uint32_t hash32(uint32_t x){ x^=x>>16; x*=0x7feb352dU; x^=x>>15; x*=0x846ca68bU; x^=x>>16; return x; }

template<int S> __global__ void mix(float4* img,int R,int iters,float* sink){
  __shared__ float4 sm[1024];
  for(int i=threadIdx.x;i<1024;i+=blockDim.x) sm[i]=make_float4(1.f,2.f,3.f,4.f);
  __syncthreads();
  uint32_t tid=blockIdx.x*blockDim.x+threadIdx.x;
  uint32_t g=tid>>2, l=tid&3;
  uint32_t s=hash32(g+1);
  float acc=0.f;
  for(int it=0;it<iters;++it){
    s=hash32(s+it);
    int x=s%(R-2)+(l&1), y=(s>>12)%(R-2)+(l>>1);
#pragma unroll
    for(int k=0;k<S;++k){ const float4 v=sm[(threadIdx.x+97*k+it)&1023]; acc+=v.x+v.w; }   // conflict-free LDS.128 = 4 wavefronts
    atomicAdd(img+(size_t)y*R+x, make_float4(1.f,2.f,3.f,acc*0.f+4.f));
  }
  if(acc==-1.f) sink[0]=acc;
}

template<int S> __global__ void lds_only(int iters,float* sink){
  __shared__ float4 sm[1024];
  for(int i=threadIdx.x;i<1024;i+=blockDim.x) sm[i]=make_float4(1.f,2.f,3.f,4.f);
  __syncthreads();
  float acc=0.f;
  for(int it=0;it<iters;++it){
#pragma unroll
    for(int k=0;k<S;++k){ const float4 v=sm[(threadIdx.x+97*k+it)&1023]; acc+=v.x+v.w; }
  }
  if(acc==-1.f) sink[0]=acc;
}

int main(){
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr,0));
  printf("device %s SMs %d\n",pr.name,pr.multiProcessorCount);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int R=2048, nsm=pr.multiProcessorCount, iters=200, blocks=nsm*8, threads=256;
  float4* img; CK(cudaMalloc(&img,(size_t)R*R*16)); CK(cudaMemset(img,0,(size_t)R*R*16));
  float* sink; CK(cudaMalloc(&sink,4));
  float ms;
  const double lanes=(double)blocks*threads*iters, clk=1.965e9;
#define RUN(S) { for(int rep=0;rep<2;++rep){ cudaEventRecord(e0); mix<S><<<blocks,threads>>>(img,R,iters,sink); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms,e0,e1);} \
  float ms2=0; if(S>0){ for(int rep=0;rep<2;++rep){ cudaEventRecord(e0); lds_only<(S>0?S:1)><<<blocks,threads>>>(iters,sink); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms2,e0,e1);} } \
  printf("REDs + %d LDS.128 per RED: %.3f ms  = %.2f clk per RED lane per SM;  the LDS alone: %.3f ms (%.2f clk per lane-iteration)\n", S, ms, ms*1e-3*clk*nsm/lanes, ms2, ms2*1e-3*clk*nsm/lanes); }
  RUN(0) RUN(1) RUN(2) RUN(4) RUN(8)
  return 0;
}
