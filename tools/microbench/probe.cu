// Early hardware probe (dev tool, not product): strided column-tile copy bandwidth for
// W = 1,2,4,8 adjacent complex128 columns, contiguous copy baseline, FP64 FMA / sincos / exp rates.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__global__ void copy_contig(const double2* __restrict__ in, double2* __restrict__ out, size_t n){
  size_t i = blockIdx.x*(size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x*blockDim.x;
  for(; i<n; i+=stride) out[i]=in[i];
}
// tile of W columns x NY rows, E elements per thread held in registers, in place
template<int W, int E>
__global__ void copy_cols(double2* __restrict__ a, int nx, int ny, int planes_tiles){
  int tile = blockIdx.x;              // over planes * nx/W
  int tiles_per_plane = nx / W;
  int plane = tile / tiles_per_plane;
  int col0 = (tile % tiles_per_plane) * W;
  int c = threadIdx.x % W, j = threadIdx.x / W;
  int rows_per_m = ny / E;
  double2* base = a + (size_t)plane*nx*ny + col0 + c;
  double2 v[E];
#pragma unroll
  for(int m=0;m<E;m++) v[m] = base[(size_t)(j + m*rows_per_m)*nx];
#pragma unroll
  for(int m=0;m<E;m++){ v[m].x += 1.0; }
#pragma unroll
  for(int m=0;m<E;m++) base[(size_t)(j + m*rows_per_m)*nx] = v[m];
}
__global__ void fma_peak(double* out, int iters){
  double a0=threadIdx.x*1e-9, a1=a0+1,a2=a0+2,a3=a0+3,a4=a0+4,a5=a0+5,a6=a0+6,a7=a0+7;
  double b=1.0000001, c=1e-9;
  for(int i=0;i<iters;i++){
    a0=fma(a0,b,c);a1=fma(a1,b,c);a2=fma(a2,b,c);a3=fma(a3,b,c);
    a4=fma(a4,b,c);a5=fma(a5,b,c);a6=fma(a6,b,c);a7=fma(a7,b,c);
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=a0+a1+a2+a3+a4+a5+a6+a7;
}
template<int MODE>
__global__ void trans_rate(double* out, int iters){
  double x = 0.37 + threadIdx.x*1e-3, acc=0;
  for(int i=0;i<iters;i++){
    if(MODE==0){ double s,c; sincos(x,&s,&c); acc+=s*c; }
    else if(MODE==1){ acc+=exp(-x); }
    else { double s,c; sincospi(x,&s,&c); acc+=s*c; }
    x+=1e-3;
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc;
}
template<typename F> float timeit(F f, int reps){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); for(int i=0;i<reps;i++) f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1); return ms/reps;
}
template<int W,int E> void run_cols(double2* a,int nx,int ny,int planes){
  int tiles = planes*nx/W; int threads = W*ny/E;
  float ms = timeit([&]{ copy_cols<W,E><<<tiles,threads>>>(a,nx,ny,tiles); }, 10);
  CK(cudaGetLastError());
  double bytes = 2.0*planes*(double)nx*ny*16;
  printf("col-tile copy W=%d E=%d threads=%d tiles=%d : %.3f ms  %.1f GB/s\n",W,E,threads,tiles,ms,bytes/ms*1e-6);
}
int main(){
  int nx=2048, ny=2048, planes=8; // 8 planes x 64 MiB = 512 MiB
  size_t n=(size_t)planes*nx*ny;
  double2 *a,*b; CK(cudaMalloc(&a,n*16)); CK(cudaMalloc(&b,n*16));
  CK(cudaMemset(a,0,n*16)); CK(cudaMemset(b,0,n*16));
  float ms = timeit([&]{ copy_contig<<<148*8,512>>>(a,b,n); },10);
  printf("contig copy: %.3f ms %.1f GB/s\n",ms,2.0*n*16/ms*1e-6);
  ms = timeit([&]{ cudaMemcpyAsync(b,a,n*16,cudaMemcpyDeviceToDevice); },10);
  printf("cudaMemcpy D2D: %.3f ms %.1f GB/s\n",ms,2.0*n*16/ms*1e-6);
  run_cols<1,16>(a,nx,ny,planes);
  run_cols<2,16>(a,nx,ny,planes);
  run_cols<4,16>(a,nx,ny,planes);

  run_cols<4,8>(a,nx,ny,planes);

  run_cols<2,8>(a,nx,ny,planes);
  double* o; CK(cudaMalloc(&o, 148*8*1024*8));
  int iters=20000;
  ms = timeit([&]{ fma_peak<<<148*2,1024>>>(o,iters); },3);
  printf("FP64 FMA: %.3f ms  %.2f TFLOP/s\n",ms, 2.0*8*iters*148.0*2*1024/ms*1e-9);
  int it2=2000;
  ms = timeit([&]{ trans_rate<0><<<148*2,1024>>>(o,it2); },3);
  printf("sincos f64: %.3f ms  %.2f G/s\n",ms, (double)it2*148*2*1024/ms*1e-6);
  ms = timeit([&]{ trans_rate<1><<<148*2,1024>>>(o,it2); },3);
  printf("exp f64: %.3f ms  %.2f G/s\n",ms, (double)it2*148*2*1024/ms*1e-6);
  ms = timeit([&]{ trans_rate<2><<<148*2,1024>>>(o,it2); },3);
  printf("sincospi f64: %.3f ms  %.2f G/s\n",ms, (double)it2*148*2*1024/ms*1e-6);
  return 0;
}
