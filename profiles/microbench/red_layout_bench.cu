// Micro-benchmark 4 (round 2): does the image LAYOUT change the per-lane cost of REDG.E.ADD.F32x4 for 2x2-pixel
// footprints at random alignment?  VERDICT r01 item 5: a block-linear image (BW x BH pixel blocks contiguous) puts a
// 2x2 footprint into one 64 B / 128 B block when aligned.  Also: is the RED limit per SM (L1TEX/LSU) or chip-wide (L2)?
// -> run the scattered pattern on 1/4, 1/2 and all SMs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_layout_bench red_layout_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__device__ __forceinline__ uint32_t hash32(uint32_t x){ x^=x>>16; x*=0x7feb352dU; x^=x>>15; x*=0x846ca68bU; x^=x>>16; return x; }

// pixel (x, y) -> float4 index in an image of R x R pixels stored as BW x BH blocks (row-major blocks, row-major inside)
template<int BW,int BH> __device__ __forceinline__ size_t pix_index(int x,int y,int R){
  if (BW==1 && BH==1) return (size_t)y*R+x;
  const int bx=x/BW, by=y/BH, ix=x%BW, iy=y%BH;
  return ((size_t)by*(R/BW)+bx)*(BW*BH)+iy*BW+ix;
}

// MODE 0: 4 lanes = one 2x2 quad in ONE instruction.  MODE 1: lane pair = two adjacent columns, the two rows in two
// consecutive instructions (what K1's (cell column, row pair) work items do).  MODE 2: one lane = whole quad, 4 instructions.
template<int BW,int BH,int MODE> __global__ void red_quads(float4* img,int R,int iters){
  const uint32_t tid=blockIdx.x*blockDim.x+threadIdx.x;
  const int G = MODE==0?4:MODE==1?2:1;
  const uint32_t g=tid/G, l=tid%G;
  uint32_t s=hash32(g+1);
  const float4 v=make_float4(1.f,2.f,3.f,4.f);
  for(int it=0;it<iters;++it){
    s=hash32(s+it);
    const int x=s%(R-2), y=(s>>12)%(R-2);
    if(MODE==0){ atomicAdd(img+pix_index<BW,BH>(x+(l&1),y+(l>>1),R),v); }
    else if(MODE==1){ atomicAdd(img+pix_index<BW,BH>(x+l,y,R),v); atomicAdd(img+pix_index<BW,BH>(x+l,y+1,R),v); }
    else { atomicAdd(img+pix_index<BW,BH>(x,y,R),v); atomicAdd(img+pix_index<BW,BH>(x+1,y,R),v);
           atomicAdd(img+pix_index<BW,BH>(x,y+1,R),v); atomicAdd(img+pix_index<BW,BH>(x+1,y+1,R),v); }
  }
}

__global__ void red_scattered(float4* img,int R,int iters){
  const uint32_t tid=blockIdx.x*blockDim.x+threadIdx.x;
  uint32_t s=hash32(tid+1);
  for(int it=0;it<iters;++it){ s=hash32(s+it); atomicAdd(img+(size_t)((s>>12)%R)*R+(s%R), make_float4(1.f,2.f,3.f,4.f)); }
}

int main(){
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr,0));
  printf("device %s SMs %d\n",pr.name,pr.multiProcessorCount);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int nsm=pr.multiProcessorCount; float ms;
  const int R=2048;
  float4* img; CK(cudaMalloc(&img,(size_t)R*R*16)); CK(cudaMemset(img,0,(size_t)R*R*16));
  const int iters=200, threads=256;
  for(int frac=4; frac>=1; frac/=2){
    const int blocks=nsm/frac;        // ONE CTA per SM at most -> `blocks` SMs busy (8 warps each)
    for(int rep=0;rep<2;++rep){ cudaEventRecord(e0); red_scattered<<<blocks,1024>>>(img,R,iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms,e0,e1);}
    const double n=(double)blocks*1024*iters;
    printf("scattered on %3d SMs (1 CTA of 1024 each): %.3f ms %.3e lanes/s (%.2f lanes/clk/busy SM)\n",blocks,ms,n/ms*1e3,n/ms*1e3/blocks/1.9e9);
  }
  const int blocks=nsm*8;
#define RUN(BW,BH,MODE,lanes_per_thread,name) { for(int rep=0;rep<2;++rep){ cudaEventRecord(e0); red_quads<BW,BH,MODE><<<blocks,threads>>>(img,R,iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms,e0,e1);} \
  const double n=(double)blocks*threads*iters*lanes_per_thread; printf("%-58s: %.3f ms  %.3e lanes/s (%.2f lanes/clk/SM)\n",name,ms,n/ms*1e3,n/ms*1e3/nsm/1.9e9); }
  RUN(1,1,0,1,"row-major      quad in one instr (4 lanes)")
  RUN(2,2,0,1,"2x2 blocks 64B quad in one instr (4 lanes)")
  RUN(4,2,0,1,"4x2 blocks 128B quad in one instr (4 lanes)")
  RUN(4,4,0,1,"4x4 blocks 256B quad in one instr (4 lanes)")
  RUN(1,1,1,2,"row-major      lane pair, rows in 2 instr (K1 pattern)")
  RUN(2,2,1,2,"2x2 blocks 64B lane pair, rows in 2 instr")
  RUN(4,2,1,2,"4x2 blocks 128B lane pair, rows in 2 instr")
  RUN(4,4,1,2,"4x4 blocks 256B lane pair, rows in 2 instr")
  RUN(1,1,2,4,"row-major      one lane per quad, 4 instr")
  RUN(2,2,2,4,"2x2 blocks 64B one lane per quad, 4 instr")
  RUN(4,2,2,4,"4x2 blocks 128B one lane per quad, 4 instr")
  return 0;
}
