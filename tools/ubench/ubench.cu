// Microbenchmarks that decide the counting design (development aid; not shipped in the library).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__device__ __forceinline__ uint64_t mix64(uint64_t x){x^=x>>33;x*=0xff51afd7ed558ccdull;x^=x>>33;x*=0xc4ceb9fe1a85ec53ull;x^=x>>33;return x;}

template<int MODE> __global__ void k_match(uint32_t* out, int iters){
  uint32_t x = mix64(threadIdx.x + blockIdx.x*blockDim.x), acc=0;
  for(int i=0;i<iters;i++){
    uint32_t d = (x>>8)&0xFF;
    uint32_t peers;
    if (MODE==0){ peers=0xffffffffu;
      #pragma unroll
      for(int b=0;b<8;b++){ bool bit=(d>>b)&1; uint32_t bal=__ballot_sync(0xffffffffu,bit); peers &= bit?bal:~bal; } }
    else if (MODE==1){ peers=0xffffffffu;
      #pragma unroll
      for(int b=0;b<8;b++){ uint32_t bal=__ballot_sync(0xffffffffu,(d>>b)&1); int s=((int)(d<<(31-b)))>>31; peers &= ~(bal ^ (uint32_t)s);} }
    else { peers=__match_any_sync(0xffffffffu,d); }
    acc += __popc(peers); x = x*1664525u+1013904223u + peers;
  }
  out[threadIdx.x + blockIdx.x*blockDim.x]=acc;
}

// shared atomics: MODE 0 = atomicAdd no return, 1 = with return, 2 = CAS64
template<int MODE> __global__ void k_satom(uint32_t* out, int iters){
  __shared__ unsigned long long sh64[4096];
  uint32_t* sh=(uint32_t*)sh64;
  for(int i=threadIdx.x;i<8192;i+=blockDim.x) sh[i]=0;
  __syncthreads();
  uint32_t x = mix64(threadIdx.x + blockIdx.x*blockDim.x), acc=0;
  for(int i=0;i<iters;i++){
    x = x*1664525u+1013904223u;
    uint32_t d=(x>>10)&0xFF;
    if(MODE==0) atomicAdd(&sh[(threadIdx.x>>5)*256+d],1u);
    else if(MODE==1) acc+=atomicAdd(&sh[(threadIdx.x>>5)*256+d],1u);
    else { unsigned long long k=(x>>4)|1ull; acc+=(uint32_t)atomicCAS(&sh64[(x>>12)&4095],0ull,k); }
  }
  __syncthreads();
  out[threadIdx.x + blockIdx.x*blockDim.x]=acc+sh[threadIdx.x];
}

// global table inserts: CAS key + RED count on 16B slots, table of `cap` slots
__global__ void k_gins(ulonglong2* tab, uint64_t mask, uint64_t nkeys, uint64_t distinct){
  uint64_t stride=(uint64_t)gridDim.x*blockDim.x;
  for(uint64_t i=blockIdx.x*(uint64_t)blockDim.x+threadIdx.x;i<nkeys;i+=stride){
    uint64_t k = mix64(i % distinct + 12345)|1ull;   // ~5x duplicates when distinct = nkeys/5
    uint64_t h = mix64(k)&mask;
    for(int p=0;p<1000;p++){
      unsigned long long* kp=(unsigned long long*)&tab[h];
      unsigned long long cur=*(volatile unsigned long long*)kp;
      if(cur==0) cur=atomicCAS(kp,0ull,k);
      if(cur==0||cur==k){ atomicAdd((uint32_t*)(kp+1),1u); break; }
      h=(h+1)&mask;
    }
  }
}
// plain streaming copy for reference
__global__ void k_copy(const ulonglong2* a, ulonglong2* b, uint64_t n){
  uint64_t stride=(uint64_t)gridDim.x*blockDim.x;
  for(uint64_t i=blockIdx.x*(uint64_t)blockDim.x+threadIdx.x;i<n;i+=stride) b[i]=a[i];
}

template<class F> float timeit(F f,int rep=3){ cudaEvent_t a,b; cudaEventCreate(&a);cudaEventCreate(&b); f(); cudaDeviceSynchronize(); float best=1e30f; for(int r=0;r<rep;r++){cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); if(ms<best)best=ms;} return best;}

int main(){
  uint32_t* out; CK(cudaMalloc(&out, 148*32*256*4));
  const int iters=4096; const double thr=148.0*8*256; // 8 CTAs of 256 per SM
  float t;
  t=timeit([&]{k_match<0><<<148*8,256>>>(out,iters);}); printf("match ballot-sel   : %.3f ms  %.1f Gkeys/s\n",t,thr*iters/t/1e6);
  t=timeit([&]{k_match<1><<<148*8,256>>>(out,iters);}); printf("match ballot-xnor  : %.3f ms  %.1f Gkeys/s\n",t,thr*iters/t/1e6);
  t=timeit([&]{k_match<2><<<148*8,256>>>(out,iters);}); printf("match MATCH.ANY    : %.3f ms  %.1f Gkeys/s\n",t,thr*iters/t/1e6);
  const double thr2=148.0*4*256;
  t=timeit([&]{k_satom<0><<<148*4,256>>>(out,iters);}); printf("smem RED (no ret)  : %.3f ms  %.1f Gops/s\n",t,thr2*iters/t/1e6);
  t=timeit([&]{k_satom<1><<<148*4,256>>>(out,iters);}); printf("smem ATOM (ret)    : %.3f ms  %.1f Gops/s\n",t,thr2*iters/t/1e6);
  t=timeit([&]{k_satom<2><<<148*4,256>>>(out,iters);}); printf("smem CAS64         : %.3f ms  %.1f Gops/s\n",t,thr2*iters/t/1e6);
  for(int lg : {21,23,25,28}){
    uint64_t cap=1ull<<lg; ulonglong2* tab; CK(cudaMalloc(&tab,cap*16));
    uint64_t nkeys = 1ull<<28; uint64_t distinct = cap/4;   // load 0.25
    float best=1e30f;
    for(int r=0;r<2;r++){ cudaMemset(tab,0,cap*16); cudaDeviceSynchronize(); float ms=timeit([&]{k_gins<<<148*16,256>>>(tab,cap-1,nkeys,distinct);},1); if(ms<best)best=ms; }
    printf("global insert table %4llu MB, %llu keys (%llu distinct): %.2f ms  %.1f Ginserts/s\n",(unsigned long long)(cap*16>>20),(unsigned long long)nkeys,(unsigned long long)distinct,best,nkeys/best/1e6);
    cudaFree(tab);
  }
  { uint64_t n=1ull<<28; ulonglong2 *a,*b; CK(cudaMalloc(&a,n*16)); CK(cudaMalloc(&b,n*16)); t=timeit([&]{k_copy<<<148*16,256>>>(a,b,n);}); printf("copy 4GB->4GB      : %.3f ms  %.0f GB/s\n",t,2.0*n*16/t/1e6); }
  return 0;
}
