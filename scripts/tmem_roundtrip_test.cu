// Developer check (run on B200): every thread of a 512-thread CTA stores 32 words in tensor memory with tcgen05.st.32x32b and
// reads them back with tcgen05.ld -- the per-thread constant store the map kernel uses.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/tmem_roundtrip_test.cu
#include <cstdio>
#include "../joxsz_b200/csrc/jx_tmem.cuh"
__global__ void k(uint32_t* o){ __shared__ uint32_t slot; if (threadIdx.x<32) tmem_alloc(&slot,512); tmem_fence_before_sync(); __syncthreads(); tmem_fence_after_sync(); uint32_t base=slot; uint32_t r[32]; for(int i=0;i<32;++i) r[i]=threadIdx.x*100+i; uint32_t ta = base + ((uint32_t)(32*((threadIdx.x>>5)&3))<<16) + 128*(threadIdx.x>>7); tmem_st32(ta,r); tmem_wait_st(); uint32_t q[32]; tmem_ld32(ta,q); tmem_wait_ld(); uint32_t s=0; for(int i=0;i<32;++i) s+= (q[i]==r[i]); o[threadIdx.x]=s; __syncthreads(); if (threadIdx.x<32) tmem_dealloc(base,512);} 
int main(){uint32_t* d; cudaMalloc(&d,512*4); k<<<1,512>>>(d); uint32_t h[512]; cudaMemcpy(h,d,sizeof h,cudaMemcpyDeviceToHost); int bad=0; for(int i=0;i<512;++i) bad+= h[i]!=32; printf("err=%s bad=%d\n", cudaGetErrorString(cudaGetLastError()), bad); return bad;}
