// Microbenchmark: raw fp64 pipe rates on B200 (DMMA.8x8x4 vs DFMA), used to set
// the fp64 roofline denominator (MEASURED_PEAKS.json has no fp64 entry).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__);exit(1);}}while(0)

template<int NACC>
__global__ void __launch_bounds__(256) dmma_peak(double* out, int iters, double av, double bv){
  double c[NACC][2];
  #pragma unroll
  for(int i=0;i<NACC;i++){c[i][0]=0;c[i][1]=0;}
  double a = av + threadIdx.x*1e-9, b = bv;
  for(int it=0; it<iters; it++){
    #pragma unroll
    for(int i=0;i<NACC;i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i][0]+c[i][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int NACC>
__global__ void __launch_bounds__(256) dfma_peak(double* out, int iters, double av, double bv){
  double c[NACC];
  #pragma unroll
  for(int i=0;i<NACC;i++) c[i]=i;
  double a = av + threadIdx.x*1e-9, b = bv;
  for(int it=0; it<iters; it++){
    #pragma unroll
    for(int i=0;i<NACC;i++) c[i]=fma(a,c[i],b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
int main(){
  int dev=0; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,dev));
  printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  double* out; CK(cudaMalloc(&out, 148*8*1024*8));
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for(int warps=1; warps<=8; warps*=2){
   for(int ctas=1; ctas<=4; ctas*=2){
    int iters=20000; int grid=p.multiProcessorCount*ctas; int threads=32*warps;
    dmma_peak<8><<<grid,threads>>>(out,100,1.0,1.0);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); dmma_peak<8><<<grid,threads>>>(out,iters,1.0,1e-3); cudaEventRecord(e1);
    CK(cudaDeviceSynchronize()); float ms; cudaEventElapsedTime(&ms,e0,e1);
    double flops = 2.0*256*8*(double)iters*warps*grid;
    printf("DMMA  warps/CTA %d CTAs/SM %d : %.2f ms  %.2f TF/s\n", warps, ctas, ms, flops/ms*1e-9);
   }
  }
  for(int warps=4; warps<=8; warps*=2){
    int iters=20000; int grid=p.multiProcessorCount*2; int threads=32*warps;
    dfma_peak<16><<<grid,threads>>>(out,100,1.0,1.0);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); dfma_peak<16><<<grid,threads>>>(out,iters,1.0000001,1e-3); cudaEventRecord(e1);
    CK(cudaDeviceSynchronize()); float ms; cudaEventElapsedTime(&ms,e0,e1);
    double flops = 2.0*32*16*(double)iters*warps*grid;
    printf("DFMA  warps/CTA %d CTAs/SM 2 : %.2f ms  %.2f TF/s\n", warps, ms, flops/ms*1e-9);
  }
  // sustained: 3 s of DMMA, report rate of the last launch
  {
    int grid=p.multiProcessorCount*2, threads=256, iters=200000; float ms=0;
    for(int rep=0; rep<12; rep++){
      cudaEventRecord(e0); dmma_peak<8><<<grid,threads>>>(out,iters,1.0,1e-3); cudaEventRecord(e1);
      CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms,e0,e1);
      double flops = 2.0*256*8*(double)iters*8*grid;
      printf("DMMA sustained rep %d: %.1f ms %.2f TF/s\n", rep, ms, flops/ms*1e-9);
    }
  }
  return 0;
}
