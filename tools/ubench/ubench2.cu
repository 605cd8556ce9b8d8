// L2 RED throughput on small histograms + scattered-store cost (development aid).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix64(uint64_t x){x^=x>>33;x*=0xff51afd7ed558ccdull;x^=x>>33;x*=0xc4ceb9fe1a85ec53ull;x^=x>>33;return x;}
__global__ void k_red(uint32_t* hist, uint32_t mask, uint64_t n){
  uint64_t stride=(uint64_t)gridDim.x*blockDim.x;
  for(uint64_t i=blockIdx.x*(uint64_t)blockDim.x+threadIdx.x;i<n;i+=stride){ atomicAdd(&hist[mix64(i)&mask],1u); }
}
// same but locality like sorted-ish tiles: consecutive threads hit bins that share high bits
__global__ void k_red_local(uint32_t* hist, uint32_t mask, uint64_t n){
  uint64_t stride=(uint64_t)gridDim.x*blockDim.x;
  for(uint64_t i=blockIdx.x*(uint64_t)blockDim.x+threadIdx.x;i<n;i+=stride){ uint32_t hi=(uint32_t)(mix64(i>>12)&mask)&~511u; atomicAdd(&hist[hi|(mix64(i)&511)],1u); }
}
template<class F> float timeit(F f,int rep=3){ cudaEvent_t a,b; cudaEventCreate(&a);cudaEventCreate(&b); f(); cudaDeviceSynchronize(); float best=1e30f; for(int r=0;r<rep;r++){cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); if(ms<best)best=ms;} return best;}
int main(){
  uint64_t n=1ull<<29;
  for(int lg: {8,12,16,18,20,24}){ uint32_t* h; cudaMalloc(&h,(4ull<<lg)); cudaMemset(h,0,4ull<<lg);
    float t=timeit([&]{k_red<<<148*16,256>>>(h,(1u<<lg)-1,n);}); float t2=timeit([&]{k_red_local<<<148*16,256>>>(h,(1u<<lg)-1,n);});
    printf("global RED, %2d-bit histogram: random %.2f ms %.1f Gops/s | tile-local %.2f ms %.1f Gops/s\n",lg,t,n/t/1e6,t2,n/t2/1e6); cudaFree(h);}
  return 0;
}
