// Microbenchmarks that fix the design constants of the ALS streaming kernels on B200:
// DMMA.8x8x4 rate, DFMA rate, LDG.256 streaming rate, and streaming + DMMA at the
// ALS k=16 ratio (8 DMMA per 32 B loaded per lane).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
   : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void ld256(const double* p, double (&v)[4]){
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];\n" : "=d"(v[0]),"=d"(v[1]),"=d"(v[2]),"=d"(v[3]) : "l"(p));
}
__global__ void k_dmma(double* out, int iters){
  double acc[8][2]; for(int i=0;i<8;i++){acc[i][0]=0;acc[i][1]=0;}
  double a = threadIdx.x*1e-3, b = threadIdx.x*2e-3;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<8;i++) dmma884(acc[i][0],acc[i][1],a,b);
  }
  double s=0; for(int i=0;i<8;i++) s+=acc[i][0]+acc[i][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void k_dfma(double* out, int iters){
  double acc[16]; for(int i=0;i<16;i++) acc[i]=i;
  double a = 1.0000001, b = threadIdx.x*1e-9;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<16;i++) acc[i]=fma(acc[i],a,b);
  }
  double s=0; for(int i=0;i<16;i++) s+=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// stream: each warp reads contiguous 1 KB chunks grid-strided; optional DMMAs per load
template<int NDMMA, int UNROLL>
__global__ void k_stream(const double* __restrict__ A, size_t n4, double* out){
  size_t tid = blockIdx.x*(size_t)blockDim.x+threadIdx.x;
  size_t stride = (size_t)gridDim.x*blockDim.x;
  double acc[8][2]; for(int i=0;i<8;i++){acc[i][0]=0;acc[i][1]=0;}
  double s=0;
  size_t i=tid;
  for(; i + (UNROLL-1)*stride < n4; i += UNROLL*stride){
    double v[UNROLL][4];
    #pragma unroll
    for(int u=0;u<UNROLL;u++) ld256(A + 4*(i+u*stride), v[u]);
    #pragma unroll
    for(int u=0;u<UNROLL;u++){
      if (NDMMA==0) s += v[u][0]+v[u][1]+v[u][2]+v[u][3];
      else {
        #pragma unroll
        for(int d=0; d<NDMMA; ++d) dmma884(acc[d&7][0],acc[d&7][1],v[u][d&3],v[u][(d+1)&3]);
      }
    }
  }
  for(int q=0;q<8;q++) s+=acc[q][0]+acc[q][1];
  if (s==123.456) out[0]=s;
}
int main(){
  int dev=0; CK(cudaSetDevice(dev)); cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,dev));
  int sms=p.multiProcessorCount; printf("device %s SMs %d clock %d kHz\n",p.name,sms,p.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double)*sms*32*1024));
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
  for(int warps=4; warps<=32; warps*=2){
    int iters=20000; dim3 g(sms), b(warps*32);
    k_dmma<<<g,b>>>(out,100); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); k_dmma<<<g,b>>>(out,iters); cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms,e0,e1);
    double flops = (double)sms*warps*iters*8*512; printf("DMMA.8x8x4 warps/SM=%d: %.2f TFLOP/s (%.3f ms)\n",warps,flops/ms/1e9,ms);
    k_dfma<<<g,b>>>(out,100); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); k_dfma<<<g,b>>>(out,iters); cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms,e0,e1);
    flops = (double)sms*warps*32*iters*16*2; printf("DFMA       warps/SM=%d: %.2f TFLOP/s (%.3f ms)\n",warps,flops/ms/1e9,ms);
  }
  size_t bytes = (size_t)1600*1000*1000; size_t n4 = bytes/32; double* A; CK(cudaMalloc(&A, bytes)); CK(cudaMemset(A,0,bytes));
  #define RUN(ND,UN,ctas,thr) { k_stream<ND,UN><<<sms*ctas,thr>>>(A,n4,out); CK(cudaDeviceSynchronize()); float best=1e9; for(int r=0;r<5;r++){cudaEventRecord(e0); k_stream<ND,UN><<<sms*ctas,thr>>>(A,n4,out); cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms;} printf("stream NDMMA=%d unroll=%d ctas/SM=%d thr=%d: %.1f GB/s (%.3f ms)\n",ND,UN,ctas,thr,bytes/best/1e6,best); }
  RUN(0,1,8,256) RUN(0,2,4,256) RUN(0,4,4,256) RUN(0,4,2,256) RUN(0,8,2,256)
  RUN(8,1,8,256) RUN(8,2,4,256) RUN(8,4,4,256) RUN(8,4,3,128) RUN(8,4,2,256) RUN(8,8,2,128)
  RUN(4,4,4,256) RUN(16,4,4,256) RUN(32,4,4,256)
  return 0;
}
