#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]),"d"(a[1]),"d"(a[2]),"d"(a[3]),"d"(a[4]),"d"(a[5]),"d"(a[6]),"d"(a[7]),
                 "d"(b[0]),"d"(b[1]),"d"(b[2]),"d"(b[3]));
}
template<int ILP>
__global__ void k_dmma884(double *out, int iters) {
  double c0[ILP], c1[ILP];
  for (int i=0;i<ILP;i++){c0[i]=threadIdx.x*1e-3+i; c1[i]=i*0.5;}
  double a = threadIdx.x*1e-6, b = 1.0+threadIdx.x*1e-7;
  for (int it=0; it<iters; ++it) {
    #pragma unroll
    for (int i=0;i<ILP;i++) dmma884(c0[i], c1[i], a, b);
  }
  double s=0; for (int i=0;i<ILP;i++) s+=c0[i]+c1[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int ILP>
__global__ void k_dmma16816(double *out, int iters) {
  double c[ILP][4];
  for (int i=0;i<ILP;i++) for(int j=0;j<4;j++) c[i][j]=threadIdx.x*1e-3+i+j;
  double a[8], b[4];
  for (int j=0;j<8;j++) a[j]=threadIdx.x*1e-6+j; for(int j=0;j<4;j++) b[j]=1.0+threadIdx.x*1e-7*j;
  for (int it=0; it<iters; ++it) {
    #pragma unroll
    for (int i=0;i<ILP;i++) dmma16816(c[i], a, b);
  }
  double s=0; for (int i=0;i<ILP;i++) for(int j=0;j<4;j++) s+=c[i][j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int ILP>
__global__ void k_dfma(double *out, int iters) {
  double c[ILP];
  for (int i=0;i<ILP;i++) c[i]=threadIdx.x*1e-3+i;
  double a = 1.0+threadIdx.x*1e-9, b = threadIdx.x*1e-7;
  for (int it=0; it<iters; ++it) {
    #pragma unroll
    for (int i=0;i<ILP;i++) c[i]=fma(c[i],a,b);
  }
  double s=0; for (int i=0;i<ILP;i++) s+=c[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<typename F> float timeit(F f) {
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1); return ms;
}
int main(){
  double *out; cudaMalloc(&out, 148*8*1024*sizeof(double));
  int iters=20000;
  for (int warps : {4,8,16,32}) {
    int thr=warps*32; int blocks=148*(2048/thr > 2 ? 2 : 2048/thr);
    { float ms=timeit([&]{k_dmma884<8><<<blocks,thr>>>(out,iters);});
      double fl=(double)blocks*warps*iters*8*512; printf("dmma884   warps/blk=%d blocks=%d: %.2f TF\n",warps,blocks,fl/ms*1e-9);}
    { float ms=timeit([&]{k_dmma16816<8><<<blocks,thr>>>(out,iters);});
      double fl=(double)blocks*warps*iters*8*(2.0*16*8*16); printf("dmma16816 warps/blk=%d blocks=%d: %.2f TF\n",warps,blocks,fl/ms*1e-9);}
    { float ms=timeit([&]{k_dfma<8><<<blocks,thr>>>(out,iters);});
      double fl=(double)blocks*thr*iters*8*2.0; printf("dfma      warps/blk=%d blocks=%d: %.2f TF\n",warps,blocks,fl/ms*1e-9);}
  }
  cudaError_t e=cudaDeviceSynchronize(); printf("status %s\n", cudaGetErrorString(e));
}
