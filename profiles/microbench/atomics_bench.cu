// Micro-benchmark: fp32 accumulation throughput on B200 (sm_100a).
// Decides the accumulation strategy of the splat kernels.
//   smem_int   : ATOMS.ADD (native int) random addresses in a 32 KB tile
//   smem_f32   : atomicAdd(float) on shared = LDS+FADD+ATOMS.CAST.SPIN loop
//   red_f32/x2/x4 : REDG.E.ADD.F32{,x2,x4} random pixel in an R*R image (L2 resident)
//   red_*_local: same but the 32 lanes of a warp hit a 2-D neighbourhood (what spatially coherent particles do)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomics_bench atomics_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__device__ __forceinline__ uint32_t hash32(uint32_t x){ x^=x>>16; x*=0x7feb352dU; x^=x>>15; x*=0x846ca68bU; x^=x>>16; return x; }

template<int KIND> __global__ void smem_kernel(float* out, int iters){
  __shared__ float sf[8192];
  int* si=(int*)sf;
  for(int i=threadIdx.x;i<8192;i+=blockDim.x) sf[i]=0.f;
  __syncthreads();
  uint32_t s=hash32(blockIdx.x*blockDim.x+threadIdx.x+1);
  for(int it=0;it<iters;++it){
    s=hash32(s+it);
    int a=s&8191;
    if(KIND==0) atomicAdd(&si[a],1);
    else atomicAdd(&sf[a],1.0f);
  }
  __syncthreads();
  if(threadIdx.x==0) out[blockIdx.x]=sf[0];
}

template<int VEC,int LOCAL> __global__ void red_kernel(float* img,int R,int iters){
  uint32_t tid=blockIdx.x*blockDim.x+threadIdx.x;
  uint32_t s=hash32(tid+1);
  uint32_t ws=hash32((tid>>5)+77);
  for(int it=0;it<iters;++it){
    s=hash32(s+it);
    int x,y;
    if(LOCAL){ ws=hash32(ws+it); int bx=ws%(R-16), by=(ws>>12)%(R-16); x=bx+(s&15); y=by+((s>>4)&15);} 
    else { x=s%R; y=(s>>12)%R; }
    size_t p=(size_t)y*R+x;
    if(VEC==1) atomicAdd(&img[p],1.0f);
    else if(VEC==2) atomicAdd(((float2*)img)+p, make_float2(1.f,2.f));
    else atomicAdd(((float4*)img)+p, make_float4(1.f,2.f,3.f,4.f));
  }
}

int main(){
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr,0));
  printf("device %s SMs %d\n",pr.name,pr.multiProcessorCount);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float* out; CK(cudaMalloc(&out,1<<20));
  int nsm=pr.multiProcessorCount;
  float ms;
  for(int kind=0;kind<2;++kind){
    int iters=2000, blocks=nsm*2, threads=1024;
    for(int rep=0;rep<2;++rep){
      cudaEventRecord(e0);
      if(kind==0) smem_kernel<0><<<blocks,threads>>>(out,iters); else smem_kernel<1><<<blocks,threads>>>(out,iters);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms,e0,e1);
    }
    double n=(double)blocks*threads*iters;
    printf("%s: %.3f ms  %.3e updates/s  (%.2f /clk/SM @1.9GHz)\n",kind?"smem_f32_cas":"smem_int_add",ms,n/ms*1e3,n/ms*1e3/nsm/1.9e9);
  }
  int Rs[3]={1024,2048,4096};
  for(int ri=0;ri<3;++ri){
    int R=Rs[ri];
    float* img; CK(cudaMalloc(&img,(size_t)R*R*16)); CK(cudaMemset(img,0,(size_t)R*R*16));
    for(int local=0;local<2;++local) for(int vec=1;vec<=4;vec*=2){
      int iters=200, blocks=nsm*8, threads=256;
      for(int rep=0;rep<2;++rep){
        cudaEventRecord(e0);
        #define L(V,LC) red_kernel<V,LC><<<blocks,threads>>>(img,R,iters)
        if(local==0){ if(vec==1)L(1,0); else if(vec==2)L(2,0); else L(4,0);} else { if(vec==1)L(1,1); else if(vec==2)L(2,1); else L(4,1);} 
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms,e0,e1);
      }
      double n=(double)blocks*threads*iters;
      printf("red R=%d vec=%d %s: %.3f ms  %.3e ops/s  %.3e floats/s\n",R,vec,local?"local16x16":"random",ms,n/ms*1e3,n*vec/ms*1e3);
    }
    cudaFree(img);
  }
  return 0;
}
