// Microbenchmark behind the partition design: throughput of the primitives a
// hash partition can be built from, on random addresses.  nvcc -arch=sm_100a -O3 atomics_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix64(uint64_t x){x^=x>>33;x*=0xff51afd7ed558ccdULL;x^=x>>33;x*=0xc4ceb9fe1a85ec53ULL;x^=x>>33;return x;}

template<int MODE> // 0 RED global, 1 ATOM global (return used), 2 ATOM global + 32B store, 3 store only (random), 4 ATOMS smem return
__global__ void __launch_bounds__(256) k(uint32_t* ctr, uint32_t nb_mask, uint4* out, uint64_t n, uint32_t* sink)
{
    __shared__ uint32_t sh[4096];
    if (MODE==4) { for (int i=threadIdx.x;i<4096;i+=256) sh[i]=0; __syncthreads(); }
    uint32_t acc=0;
    for (uint64_t i=(uint64_t)blockIdx.x*256+threadIdx.x; i<n; i+=(uint64_t)gridDim.x*256*8) {
        uint32_t b[8], p[8];
        #pragma unroll
        for (int j=0;j<8;j++){ uint64_t r=i+(uint64_t)j*gridDim.x*256; b[j]=(uint32_t)(mix64(r)>>32)&nb_mask; p[j]=(uint32_t)(r<n? r:0); }
        #pragma unroll
        for (int j=0;j<8;j++){
            if (MODE==0) atomicAdd(&ctr[b[j]],1u);
            if (MODE==1||MODE==2) p[j]=atomicAdd(&ctr[b[j]],1u);
            if (MODE==4) p[j]=atomicAdd(&sh[b[j]&4095],1u);
        }
        #pragma unroll
        for (int j=0;j<8;j++){
            if (MODE==1||MODE==4) acc+=p[j];
            if (MODE==2) { uint64_t pos=((uint64_t)b[j]*((n/(nb_mask+1))+64)+ (p[j]% ((n/(nb_mask+1))+64))); out[2*pos]=make_uint4(p[j],1,2,3); out[2*pos+1]=make_uint4(4,5,6,7);}  
            if (MODE==3) { uint64_t pos=mix64(i+j*977)%n; out[2*pos]=make_uint4(p[j],1,2,3); out[2*pos+1]=make_uint4(4,5,6,7);}  
        }
    }
    if (acc==0xdeadbeef) *sink=acc;
}
int main(){
    const uint64_t n=100000000ull; 
    uint32_t *ctr,*sink; uint4* out; 
    cudaMalloc(&ctr,(1<<20)*4); cudaMalloc(&sink,4); cudaMalloc(&out,(n+ (1<<20)*64)*32);
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[]={"RED global","ATOM global (return)","ATOM global + 32B row store","random 32B row store only","ATOM shared (return)"};
    for (int logb : {10, 17}) for (int mode=0; mode<5; mode++){
        uint32_t mask=(1u<<logb)-1; float best=1e9;
        for (int rep=0;rep<3;rep++){
            cudaMemset(ctr,0,(1<<20)*4);
            cudaEventRecord(e0);
            int grid=148*8;
            switch(mode){case 0:k<0><<<grid,256>>>(ctr,mask,out,n,sink);break;case 1:k<1><<<grid,256>>>(ctr,mask,out,n,sink);break;case 2:k<2><<<grid,256>>>(ctr,mask,out,n,sink);break;case 3:k<3><<<grid,256>>>(ctr,mask,out,n,sink);break;case 4:k<4><<<grid,256>>>(ctr,mask,out,n,sink);break;}
            cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms;
        }
        printf("buckets=2^%d  %-30s %8.3f ms  %6.1f Gops/s  err=%s\n",logb,names[mode],best,n/best/1e6,cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
