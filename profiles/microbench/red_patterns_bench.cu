// Micro-benchmark 2: does the per-lane cost of REDG.E.ADD.F32x4 drop when lanes of a warp share a sector / line,
// and what does the TMA bulk reduce (cp.reduce.async.bulk ... add.f32) sustain for small row-sized payloads?
// Decides whether K1's RED ceiling (profiles/r01/atomics_bench_b200.txt: 1.9e11 lanes/s) can be lifted by lane
// arrangement or by routing footprint rows through the TMA engine.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_patterns_bench red_patterns_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__device__ __forceinline__ uint32_t hash32(uint32_t x){ x^=x>>16; x*=0x7feb352dU; x^=x>>15; x*=0x846ca68bU; x^=x>>16; return x; }

// GROUP lanes share one random base pixel and take consecutive pixels (16 B each) from it; ROWS>1 folds the group
// into a ROWS-row block (2x2 quads etc).
template<int GROUP,int ROWS> __global__ void red_group(float4* img,int R,int iters){
  uint32_t tid=blockIdx.x*blockDim.x+threadIdx.x;
  uint32_t g=tid/GROUP, l=tid%GROUP;
  uint32_t s=hash32(g+1);
  const int cols=GROUP/ROWS;
  for(int it=0;it<iters;++it){
    s=hash32(s+it);
    int x=s%(R-GROUP), y=(s>>12)%(R-ROWS);
    x+= l%cols; y+= l/cols;
    atomicAdd(img+(size_t)y*R+x, make_float4(1.f,2.f,3.f,4.f));
  }
}

// TMA bulk reduce: every thread owns BYTES of shared memory and adds it to a random BYTES-aligned image location.
template<int BYTES> __global__ void tma_reduce(float* img,int R,int iters){
  extern __shared__ __align__(128) float sm[];
  for(int i=threadIdx.x;i<blockDim.x*BYTES/4;i+=blockDim.x) sm[i]=1.0f;
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;");
  uint32_t tid=blockIdx.x*blockDim.x+threadIdx.x;
  uint32_t s=hash32(tid+1);
  uint32_t src=(uint32_t)__cvta_generic_to_shared(sm+threadIdx.x*(BYTES/4));
  size_t nchunks=(size_t)R*R*16/BYTES;
  for(int it=0;it<iters;++it){
    s=hash32(s+it);
    float* dst=img+(size_t)(s%nchunks)*(BYTES/4);
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"::"l"(dst),"r"(src),"n"(BYTES):"memory");
    asm volatile("cp.async.bulk.commit_group;");
    if((it&7)==7) asm volatile("cp.async.bulk.wait_group.read 0;":::"memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;":::"memory");
}

int main(){
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr,0));
  printf("device %s SMs %d\n",pr.name,pr.multiProcessorCount);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int nsm=pr.multiProcessorCount; float ms;
  const int R=2048;
  float4* img; CK(cudaMalloc(&img,(size_t)R*R*16)); CK(cudaMemset(img,0,(size_t)R*R*16));
  int iters=200, blocks=nsm*8, threads=256;
#define RUN(G,RW,name) for(int rep=0;rep<2;++rep){ cudaEventRecord(e0); red_group<G,RW><<<blocks,threads>>>(img,R,iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms,e0,e1);} \
  { double n=(double)blocks*threads*iters; printf("red.v4 %-28s: %.3f ms  %.3e lanes/s  (%.2f lanes/clk/SM @1.9GHz)\n",name,ms,n/ms*1e3,n/ms*1e3/nsm/1.9e9); }
  RUN(1,1,"scattered (1 px / lane)")
  RUN(2,1,"pairs (2 px row = 1 sector)")
  RUN(4,2,"quads 2x2")
  RUN(4,1,"4 px row (2 sectors)")
  RUN(8,1,"8 px row (1 line)")
  RUN(32,1,"32 px row (4 lines)")
  RUN(16,4,"4x4 block")
#define TRUN(B) { CK(cudaFuncSetAttribute(tma_reduce<B>,cudaFuncAttributeMaxDynamicSharedMemorySize,256*B)); \
  for(int rep=0;rep<2;++rep){ cudaEventRecord(e0); tma_reduce<B><<<nsm*2,256,256*B>>>((float*)img,R,iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms,e0,e1);} \
  double n=(double)nsm*2*256*iters; printf("tma bulk reduce %4d B          : %.3f ms  %.3e ops/s  %.3e px/s (%.2f ops/clk/SM)\n",B,ms,n/ms*1e3,n*(B/16)/ms*1e3,n/ms*1e3/nsm/1.9e9); }
  TRUN(16) TRUN(32) TRUN(64) TRUN(128) TRUN(512)
  return 0;
}
