// Micro-benchmark: what bounds the bucket scatter?  (A) returning atomics on random counters,
// (B) random 16-byte stores, (C) both, (D) non-returning reductions.  nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t hash(uint32_t x){ x^=x>>16; x*=0x7feb352dU; x^=x>>15; x*=0x846ca68bU; x^=x>>16; return x; }
template<int MODE, int PER>
__global__ void k(uint32_t* counts, float4* out, uint32_t ntiles, int64_t n, uint32_t cap_per_tile){
  for (int64_t g=(int64_t)blockIdx.x*blockDim.x+threadIdx.x; g*PER<n; g+=(int64_t)gridDim.x*blockDim.x){
    uint32_t t[PER], s[PER];
#pragma unroll
    for(int q=0;q<PER;q++) t[q]=hash((uint32_t)(g*PER+q))%ntiles;
    if (MODE==0 || MODE==2) {
#pragma unroll
      for(int q=0;q<PER;q++) s[q]=atomicAdd(&counts[t[q]],1u);
    } else if (MODE==3) {
#pragma unroll
      for(int q=0;q<PER;q++) { atomicAdd(&counts[t[q]],1u); s[q]=0; }
    } else {
#pragma unroll
      for(int q=0;q<PER;q++) s[q]=hash(t[q]+q+(uint32_t)g);
    }
    if (MODE==1 || MODE==2) {
#pragma unroll
      for(int q=0;q<PER;q++) out[(uint64_t)t[q]*cap_per_tile + (s[q]%cap_per_tile)] = make_float4(1,2,3,(float)s[q]);
    } else if (MODE==0) {
      uint32_t acc=0;
#pragma unroll
      for(int q=0;q<PER;q++) acc+=s[q];
      if (acc==0xdeadbeef) counts[0]=acc;
    }
  }
}
template<int MODE,int PER> float run(uint32_t* c, float4* o, uint32_t nt, int64_t n, uint32_t cap, int blocks){
  cudaMemset(c,0,nt*4); cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE,PER><<<blocks,256>>>(c,o,nt,n,cap); cudaDeviceSynchronize(); cudaMemset(c,0,nt*4);
  cudaEventRecord(a); k<MODE,PER><<<blocks,256>>>(c,o,nt,n,cap); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms,a,b); return ms; }
int main(){
  const int64_t n=71428572; 
  for (uint32_t nt : {65536u, 524288u}) {
    uint32_t cap = (uint32_t)(n/nt*1.3)+8; uint32_t* c; float4* o; cudaMalloc(&c,nt*4); cudaMalloc(&o,(size_t)nt*cap*16);
    for (int blocks : {148*5, 148*8}) {
      printf("ntiles %u blocks %d: atom_ret %.3f  store16 %.3f  both %.3f  red %.3f | PER8: atom %.3f both %.3f  ms\n", nt, blocks,
        run<0,4>(c,o,nt,n,cap,blocks), run<1,4>(c,o,nt,n,cap,blocks), run<2,4>(c,o,nt,n,cap,blocks), run<3,4>(c,o,nt,n,cap,blocks),
        run<0,8>(c,o,nt,n,cap,blocks), run<2,8>(c,o,nt,n,cap,blocks));
    }
    cudaFree(c); cudaFree(o);
  }
  return 0; }
