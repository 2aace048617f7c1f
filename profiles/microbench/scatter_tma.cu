// Are scattered 32-byte row stores faster as TMA bulk copies (shared -> global, 32 bytes each) than as STG.E.256?
// Both variants write N rows of 32 bytes to pseudo-random 32-byte slots of a 3.2 GB buffer.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scatter_tma scatter_tma.cu && ./scatter_tma
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

__global__ void __launch_bounds__(256) scatter_stg(uint4 *out, uint64_t n, uint64_t slots)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t s = mix64(i) % slots;
        const uint32_t a = (uint32_t)i;
        asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" :: "l"(out + 2 * s), "r"(a) : "memory");
    }
}

// every thread parks its row in shared memory and issues one 32-byte bulk copy; a slot is reused after the
// thread's previous bulk group has been read out (cp.async.bulk.wait_group.read)
__global__ void __launch_bounds__(256) scatter_bulk(uint4 *out, uint64_t n, uint64_t slots)
{
    __shared__ __align__(128) uint4 ring[2][256][2];
    int phase = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t s = mix64(i) % slots;
        const uint32_t a = (uint32_t)i;
        ring[phase][threadIdx.x][0] = make_uint4(a, a, a, a);
        ring[phase][threadIdx.x][1] = make_uint4(a, a, a, a);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const uint32_t src = (uint32_t)__cvta_generic_to_shared(&ring[phase][threadIdx.x][0]);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 32;" :: "l"(out + 2 * s), "r"(src) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");      // the other phase's slot is free again
        phase ^= 1;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main()
{
    const uint64_t n = 100000000ull, slots = n;
    uint4 *out;
    cudaMalloc(&out, slots * 32);
    cudaMemset(out, 0, slots * 32);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int variant = 0; variant < 2; variant++) {
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            if (variant == 0) scatter_stg<<<148 * 8, 256>>>(out, n, slots);
            else scatter_bulk<<<148 * 8, 256>>>(out, n, slots);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            printf("%s rep %d: %.3f ms  %.2f TB/s  %.1f G rows/s  (%s)\n", variant ? "bulk32" : "stg256", rep, ms,
                   n * 32 / ms * 1e-9, n / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
