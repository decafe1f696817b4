// Does DMMA.8x8x4 throughput depend on operand register variety? (operand reuse cache / RF bandwidth)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template<int MODE> __global__ void k(double* out, const double* in, int iters){
  double acc[4][2]; for(int i=0;i<4;i++){acc[i][0]=0;acc[i][1]=0;}
  double a[8], b[8];
  for(int i=0;i<8;i++){ a[i]=in[threadIdx.x+32*i]; b[i]=in[threadIdx.x+32*i+512]; }
  for(int it=0; it<iters; ++it){
    if (MODE==0){
      #pragma unroll
      for(int r=0;r<16;r++) dmma884(acc[r&3][0],acc[r&3][1],a[0],b[0]);
    } else if (MODE==1){   // like the H-step: f[cb][h].{x,y} x a[v][h].{x,y}
      #pragma unroll
      for(int h=0;h<2;h++){
        #pragma unroll
        for(int cb=0;cb<2;cb++)
        #pragma unroll
        for(int v=0;v<2;v++) dmma884(acc[cb*2+v][0],acc[cb*2+v][1],a[cb*2+h],b[v*2+h]);
        #pragma unroll
        for(int cb=0;cb<2;cb++)
        #pragma unroll
        for(int v=0;v<2;v++) dmma884(acc[cb*2+v][0],acc[cb*2+v][1],a[4+cb*2+h],b[4+v*2+h]);
      }
    } else {              // all operands distinct each time
      #pragma unroll
      for(int r=0;r<16;r++) dmma884(acc[r&3][0],acc[r&3][1],a[r&7],b[(r*3)&7]);
    }
  }
  double s=0; for(int i=0;i<4;i++) s+=acc[i][0]+acc[i][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int MODE> void run(double*out,double*in,int warps){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms; int iters=4000;
  k<MODE><<<148,warps*32>>>(out,in,100); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<148,warps*32>>>(out,in,iters); cudaEventRecord(e1); cudaDeviceSynchronize(); cudaEventElapsedTime(&ms,e0,e1);
  printf("mode %d warps/SM=%2d : %.2f TFLOP/s\n",MODE,warps,148.0*warps*iters*16*512/ms/1e9);
}
int main(){ double*out,*in; cudaMalloc(&out,8<<20); cudaMalloc(&in,8<<16); cudaMemset(in,0,8<<16);
  for(int w:{4,8,12}){ run<0>(out,in,w); run<1>(out,in,w); run<2>(out,in,w);} return 0; }
