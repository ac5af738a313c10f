// Microbenchmark: peak rate of the lazy modular MAC inner loops on B200 (operands from shared memory, no HBM traffic).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mac_rate mac_rate.cu ; run on the GPU box.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
struct AccW { uint32_t e0,e1,e2,e3,o0,o1,o2; };
__device__ __forceinline__ void macw(AccW& A, uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1){
  asm("mad.lo.cc.u32 %0, %7, %9, %0;\n\tmadc.hi.cc.u32 %1, %7, %9, %1;\n\tmadc.lo.cc.u32 %2, %8, %10, %2;\n\tmadc.hi.u32 %3, %8, %10, %3;\n\t"
      "mad.lo.cc.u32 %4, %7, %10, %4;\n\tmadc.hi.cc.u32 %5, %7, %10, %5;\n\taddc.u32 %6, %6, 0;\n\t"
      "mad.lo.cc.u32 %4, %8, %9, %4;\n\tmadc.hi.cc.u32 %5, %8, %9, %5;\n\taddc.u32 %6, %6, 0;\n\t"
      : "+r"(A.e0),"+r"(A.e1),"+r"(A.e2),"+r"(A.e3),"+r"(A.o0),"+r"(A.o1),"+r"(A.o2) : "r"(a0),"r"(a1),"r"(b0),"r"(b1));
}
struct AccN { uint32_t e0,e1,e2; };
__device__ __forceinline__ void macn(AccN& A, uint32_t a, uint32_t b){
  asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;\n\t" : "+r"(A.e0),"+r"(A.e1),"+r"(A.e2) : "r"(a),"r"(b));
}
// narrow without the third word: 64-bit accumulate only (what a reduce-every-few-steps scheme would cost)
__device__ __forceinline__ void macn2(uint64_t& A, uint32_t a, uint32_t b){ A += (uint64_t)a*b; }

template<int TR,int TC,int MODE>
__global__ void __launch_bounds__(512,1) k(uint32_t* out, int K){
  __shared__ uint64_t sa[4][32][32];   // [stage][row][lane]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for(int i=threadIdx.x;i<4*32*32;i+=blockDim.x) ((uint64_t*)sa)[i] = (uint64_t)i*0x9E3779B97F4A7C15ull;
  __syncthreads();
  AccW aw[MODE==0?TR:1][MODE==0?TC:1]; AccN an[MODE==1?TR:1][MODE==1?TC:1]; uint64_t a2[MODE==2?TR:1][MODE==2?TC:1];
  for(auto& r:aw)for(auto& x:r)x=AccW{0,0,0,0,0,0,0};
  for(auto& r:an)for(auto& x:r)x=AccN{0,0,0};
  for(auto& r:a2)for(auto& x:r)x=0;
  for(int k=0;k<K;k++){
    const uint64_t (*st)[32] = sa[k&3];
    uint64_t a[TR], b[TC];
    #pragma unroll
    for(int r=0;r<TR;r++) a[r]=st[(r+w)&31][lane];
    #pragma unroll
    for(int c=0;c<TC;c++) b[c]=st[(16+c+w)&31][lane];
    #pragma unroll
    for(int r=0;r<TR;r++)
    #pragma unroll
    for(int c=0;c<TC;c++){
      if(MODE==0) macw(aw[r][c],(uint32_t)a[r],(uint32_t)(a[r]>>32),(uint32_t)b[c],(uint32_t)(b[c]>>32));
      if(MODE==1) macn(an[r][c],(uint32_t)a[r],(uint32_t)b[c]);
      if(MODE==2) macn2(a2[r][c],(uint32_t)a[r],(uint32_t)b[c]);
    }
  }
  uint32_t s=0;
  for(auto& r:aw)for(auto& A:r) s^=A.e0^A.e1^A.e2^A.e3^A.o0^A.o1^A.o2;
  for(auto& r:an)for(auto& A:r) s^=A.e0^A.e1^A.e2;
  for(auto& r:a2)for(auto& A:r) s^=(uint32_t)A^(uint32_t)(A>>32);
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int TR,int TC,int MODE> void run(const char* name, int threads){
  uint32_t* out; cudaMalloc(&out, 148*8*1024*4);
  int K=4096; int blocks=148*2;
  k<TR,TC,MODE><<<blocks,threads>>>(out,64); cudaDeviceSynchronize();
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<TR,TC,MODE><<<blocks,threads>>>(out,K); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  double macs=(double)blocks*threads*K*TR*TC;
  int occ=0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ,k<TR,TC,MODE>,threads,0);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa,k<TR,TC,MODE>);
  printf("%-28s threads=%4d regs=%3d occ=%d  %.3f ms  %.2f TMAC/s  (%.1f MAC/clk/SM @1.965GHz)\n",name,threads,fa.numRegs,occ,ms,macs/ms/1e9,macs/ms/1e3/148/1.965e6);
  cudaFree(out);
}
int main(){
  run<5,2,0>("wide 5x2",512); run<4,2,0>("wide 4x2",512); run<10,1,0>("wide 10x1",512); run<5,4,0>("wide 5x4",256);
  run<10,4,1>("narrow96 10x4",512); run<10,2,1>("narrow96 10x2",512); run<8,4,1>("narrow96 8x4",512); run<5,4,1>("narrow96 5x4",1024); run<10,4,1>("narrow96 10x4 t256",256);
  run<10,4,2>("narrow64 10x4",512); run<10,2,2>("narrow64 10x2",512); run<8,8,2>("narrow64 8x8",256);
  return 0;
}
