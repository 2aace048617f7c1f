// Microbenchmark: HBM write bandwidth of a bucketed scatter of 32-byte rows as a function of
// (a) the number of open buckets and (b) the contiguous chunk (rows) each visit writes, plus the
// read bandwidth of gathering fixed-size segments from random places.  No atomics involved:
// row r -> bucket (r*ODD) & mask, slot r >> logb (each bucket visited round-robin).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix64(uint64_t x){x^=x>>33;x*=0xff51afd7ed558ccdULL;x^=x>>33;x*=0xc4ceb9fe1a85ec53ULL;x^=x>>33;return x;}

template<int CHUNK>   // rows written contiguously per visit (1,2,4,8,16)
__global__ void __launch_bounds__(256) scatter(uint4* out, uint64_t nvisits, int logb, uint64_t cap_rows)
{
    const uint32_t mask=(1u<<logb)-1;
    for (uint64_t v=(uint64_t)blockIdx.x*256+threadIdx.x; v<nvisits; v+=(uint64_t)gridDim.x*256) {
        const uint32_t b=(uint32_t)(v*2654435761ull)&mask;
        const uint64_t slot=(v>>logb)*CHUNK;
        uint4* dst=out+2*((uint64_t)b*cap_rows+slot);
        #pragma unroll
        for (int j=0;j<CHUNK;j++){ dst[2*j]=make_uint4((uint32_t)v,1,2,3); dst[2*j+1]=make_uint4(4,5,6,j);}  
    }
}
// warp-cooperative variant: a warp writes CHUNK rows of one bucket with consecutive lanes (coalesced)
template<int CHUNK>
__global__ void __launch_bounds__(256) scatter_warp(uint4* out, uint64_t nvisits, int logb, uint64_t cap_rows)
{
    const uint32_t mask=(1u<<logb)-1; const int lane=threadIdx.x&31; constexpr int VPW=64/CHUNK; // visits per warp-iteration (64 uint4 = 32 rows per iter.. use 2 stores)
    const uint64_t warp=((uint64_t)blockIdx.x*256+threadIdx.x)>>5, nwarps=((uint64_t)gridDim.x*256)>>5;
    for (uint64_t v0=warp*VPW; v0<nvisits; v0+=nwarps*VPW) {
        #pragma unroll
        for (int h=0;h<2;h++){
            const int q=h*32+lane;              // uint4 index within the 64-uint4 (32 rows) group
            const uint64_t v=v0+q/(2*CHUNK); const int within=q%(2*CHUNK);
            if (v<nvisits){ const uint32_t b=(uint32_t)(v*2654435761ull)&mask; const uint64_t slot=(v>>logb)*CHUNK;
                out[2*((uint64_t)b*cap_rows+slot)+within]=make_uint4((uint32_t)v,1,2,within); }
        }
    }
}
template<int SEGROWS>
__global__ void __launch_bounds__(256) gather(const uint4* in, uint64_t nseg, uint64_t total_rows, uint32_t* sink)
{
    uint32_t acc=0; const int lane=threadIdx.x&31;
    const uint64_t warp=((uint64_t)blockIdx.x*256+threadIdx.x)>>5, nwarps=((uint64_t)gridDim.x*256)>>5;
    for (uint64_t s=warp; s<nseg; s+=nwarps){
        const uint64_t start=(mix64(s)%(total_rows/SEGROWS))*SEGROWS;
        for (int q=lane;q<2*SEGROWS;q+=32){ uint4 x=in[2*start+q]; acc+=x.x^x.w; }
    }
    if (acc==0xdeadbeef)*sink=acc;
}
int main(){
    const uint64_t n=100000000ull; uint4* out; uint32_t* sink; cudaMalloc(&sink,4);
    const uint64_t slack=(1ull<<20)*64; cudaMalloc(&out,(n+slack)*32); cudaMemset(out,0,(n+slack)*32);
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); const int grid=148*8;
    auto time=[&](auto f){ float best=1e9; for(int r=0;r<3;r++){ cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms;} return best; };
    printf("thread-per-visit scatter of 1e8 32B rows (3.2 GB): ms / TB/s\n");
    for (int logb : {8,10,12,14,16,17,18,20}) {
        const uint64_t cap=(n>>logb)+64; float t1=time([&]{scatter<1><<<grid,256>>>(out,n,logb,cap);});
        float t2=time([&]{scatter<2><<<grid,256>>>(out,n/2,logb,cap);}); float t4=time([&]{scatter<4><<<grid,256>>>(out,n/4,logb,cap);});
        float t8=time([&]{scatter<8><<<grid,256>>>(out,n/8,logb,cap);});
        printf(" buckets=2^%-2d chunk1 %6.2f ms %5.2f | chunk2 %6.2f %5.2f | chunk4 %6.2f %5.2f | chunk8 %6.2f %5.2f  %s\n",logb,t1,3.2/t1,t2,3.2/t2,t4,3.2/t4,t8,3.2/t8,cudaGetErrorString(cudaGetLastError()));
    }
    printf("warp-cooperative (coalesced) scatter: \n");
    for (int logb : {12,17,20}) {
        const uint64_t cap=(n>>logb)+64;
        float t4=time([&]{scatter_warp<4><<<grid,256>>>(out,n/4,logb,cap);}); float t8=time([&]{scatter_warp<8><<<grid,256>>>(out,n/8,logb,cap);}); float t16=time([&]{scatter_warp<16><<<grid,256>>>(out,n/16,logb,cap);}); float t32=time([&]{scatter_warp<32><<<grid,256>>>(out,n/32,logb,cap);});
        printf(" buckets=2^%-2d chunk4 %6.2f ms %5.2f | chunk8 %6.2f %5.2f | chunk16 %6.2f %5.2f | chunk32 %6.2f %5.2f TB/s %s\n",logb,t4,3.2/t4,t8,3.2/t8,t16,3.2/t16,t32,3.2/t32,cudaGetErrorString(cudaGetLastError()));
    }
    printf("gather of random segments (3.2 GB total):\n");
    { float a=time([&]{gather<1><<<grid,256>>>(out,n,n,sink);}); float b=time([&]{gather<4><<<grid,256>>>(out,n/4,n,sink);}); float c=time([&]{gather<16><<<grid,256>>>(out,n/16,n,sink);}); float d=time([&]{gather<64><<<grid,256>>>(out,n/64,n,sink);});
      printf(" seg1 %6.2f ms %5.2f | seg4 %6.2f %5.2f | seg16 %6.2f %5.2f | seg64 %6.2f %5.2f TB/s %s\n",a,3.2/a,b,3.2/b,c,3.2/c,d,3.2/d,cudaGetErrorString(cudaGetLastError())); }
    return 0;
}
