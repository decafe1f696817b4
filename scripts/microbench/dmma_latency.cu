// DMMA.8x8x4 dependent-chain latency / throughput vs number of independent accumulators and warps per SM.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template<int NCH> __global__ void k(double* out, int iters){
  double acc[NCH][2]; for(int i=0;i<NCH;i++){acc[i][0]=0;acc[i][1]=0;}
  double a = threadIdx.x*1e-3, b = threadIdx.x*2e-3;
  long long t0=clock64();
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int r=0;r<16/NCH;r++)
    #pragma unroll
    for(int i=0;i<NCH;i++) dmma884(acc[i][0],acc[i][1],a,b);
  }
  long long t1=clock64();
  double s=0; for(int i=0;i<NCH;i++) s+=acc[i][0]+acc[i][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
  if(threadIdx.x==0 && blockIdx.x==0) out[1<<20]=(double)(t1-t0)/(iters*16.0);
}
template<int NCH> void run(double* out, int warps){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms; int iters=4000;
  k<NCH><<<148,warps*32>>>(out,100); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<NCH><<<148,warps*32>>>(out,iters); cudaEventRecord(e1); cudaDeviceSynchronize(); cudaEventElapsedTime(&ms,e0,e1);
  double cyc; cudaMemcpy(&cyc,out+(1<<20),8,cudaMemcpyDeviceToHost);
  printf("chains=%d warps/SM=%2d : %.2f TFLOP/s, %.1f cycles per DMMA per warp\n",NCH,warps,148.0*warps*iters*16*512/ms/1e9,cyc);
}
int main(){ double* out; cudaMalloc(&out,8*((1<<20)+8));
  for(int w: {1,4,8,16}){ run<1>(out,w); run<2>(out,w); run<4>(out,w); run<8>(out,w); run<16>(out,w);} return 0; }
