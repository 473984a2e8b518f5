// microbenchmark: random 16/32-byte reads over a large buffer; DRAM bytes per access under different L2 fetch settings
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t z){z=(z^(z>>30))*0xBF58476D1CE4E5B9ull;z=(z^(z>>27))*0x94D049BB133111EBull;return z^(z>>31);}
template<int MODE>
__global__ void rnd_kernel(const ulonglong2* __restrict__ buf, uint64_t mask, int iters, unsigned long long* out){
  uint64_t tid=(uint64_t)blockIdx.x*blockDim.x+threadIdx.x; unsigned long long acc=0;
  for(int i=0;i<iters;i+=4){
    ulonglong2 v[4];
    #pragma unroll
    for(int q=0;q<4;q++){
      uint64_t idx=mix(tid*1315423911ull+(uint64_t)(i+q))&mask;   // 16-byte units
      const ulonglong2* p=buf+idx;
      if(MODE==0) v[q]=__ldg(p);
      else if(MODE==1) v[q]=__ldcg(p);
      else if(MODE==2) { unsigned long long c,d; const ulonglong2* p2=buf+(idx&~1ull); asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.b64 {%0,%1,%2,%3}, [%4];":"=l"(v[q].x),"=l"(v[q].y),"=l"(c),"=l"(d):"l"(p2)); v[q].x^=c; v[q].y^=d; }
      else if(MODE==3) { asm volatile("ld.global.cs.v2.u64 {%0,%1}, [%2];":"=l"(v[q].x),"=l"(v[q].y):"l"(p)); }
      else if(MODE==4) { asm volatile("ld.global.cv.v2.u64 {%0,%1}, [%2];":"=l"(v[q].x),"=l"(v[q].y):"l"(p)); }
    }
    #pragma unroll
    for(int q=0;q<4;q++) acc+=v[q].x^v[q].y;
  }
  if(acc==0x1234567) *out=acc;
}
int main(int argc,char**argv){
  size_t gb=argc>1?atoi(argv[1]):8; int gran=argc>2?atoi(argv[2]):0;
  if(gran){ cudaError_t e=cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity,gran); printf("set gran %d -> %s\n",gran,cudaGetErrorString(e)); }
  size_t g=0; cudaDeviceGetLimit(&g,cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity limit = %zu\n",g);
  size_t bytes=gb<<30; ulonglong2* buf; cudaMalloc(&buf,bytes); cudaMemset(buf,1,bytes);
  unsigned long long* out; cudaMalloc(&out,8);
  uint64_t mask=(bytes/16)-1; int iters=64; int blocks=148*8, threads=256;
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  for(int mode=0;mode<5;mode++){
    for(int rep=0;rep<2;rep++){
      cudaEventRecord(a);
      switch(mode){case 0:rnd_kernel<0><<<blocks*8,threads>>>(buf,mask,iters,out);break;case 1:rnd_kernel<1><<<blocks*8,threads>>>(buf,mask,iters,out);break;
        case 2:rnd_kernel<2><<<blocks*8,threads>>>(buf,mask,iters,out);break;case 3:rnd_kernel<3><<<blocks*8,threads>>>(buf,mask,iters,out);break;case 4:rnd_kernel<4><<<blocks*8,threads>>>(buf,mask,iters,out);break;}
      cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b);
      double n=(double)blocks*8*threads*iters;
      if(rep) printf("mode %d: %.2f ms, %.2f G loads/s (%s)\n",mode,ms,n/ms/1e6,cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
