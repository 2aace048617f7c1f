// Oversized buckets: buckets with more rows than the group kernel's shared-memory capacity
// (a connection with thousands of points, or an unlucky hash bucket).  Rare by construction
// (pick_logb leaves ~2.7x head-room), so this path favours simplicity: the rows of all such
// buckets are gathered, sorted device-wide by (hash, key, flowEndSeconds) with the CUB radix
// sort that ships with the CUDA toolkit, and turned into exactly the artefacts the group kernel
// produces: per-series arrays at the bucket's own offsets, series entries in place over the
// bucket's dead part[] region, series/points per bucket.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda/std/tuple>

#include "tad_kernels.h"

namespace tad {

struct SpillRow {
    uint64_t h, a, b, value;
    uint32_t proto, t;
};
static_assert(sizeof(SpillRow) == 40, "SpillRow layout");

struct SpillDecomposer {
    __host__ __device__ ::cuda::std::tuple<uint64_t &, uint64_t &, uint64_t &, uint32_t &, uint32_t &> operator()(SpillRow &r) const
    {
        return {r.h, r.a, r.b, r.proto, r.t};
    }
};

// part[] rows carry hash-tag bits above the protocol byte (see hash_tag in tad_kernels.cu): strip them here
__device__ __forceinline__ SpillRow to_spill(const Row32 &r)
{
    SpillRow s;
    s.proto = r.proto & 0xffu;
    s.h = key_hash(r.a, r.b, s.proto);
    s.a = r.a; s.b = r.b; s.value = r.value; s.t = r.t;
    return s;
}

__global__ void __launch_bounds__(256) spill_gather_kernel(const SegDesc seg, const uint32_t *__restrict__ big_list,
                                                           const uint32_t *__restrict__ big_base, SpillRow *__restrict__ in)
{
    const uint32_t j = blockIdx.x;
    const uint32_t b = big_list[j];
    uint32_t base = big_base[j];
    if (seg.stride && seg.nseg == 1) {
        // optimistic partition, one GPU: the slot holds the first `stride` rows; they are packed at j * stride (the
        // overflow rows follow after all slots; the sort that comes next does not care about input order)
        const Row32 *rows = seg.base[0] + (size_t)b * seg.stride;
        for (uint32_t i = threadIdx.x; i < seg.stride; i += blockDim.x) {
            in[(size_t)j * seg.stride + i] = to_spill(rows[i]);
        }
        return;
    }
    // exact partition, or the multi-GPU optimistic one (every source's slot holds all of its rows: an overflow anywhere
    // sends the whole job down the exact path): the bucket is the concatenation of its per-source pieces
    for (int sg = 0; sg < seg.nseg; sg++) {
        uint32_t off, n;
        seg_span(seg, sg, b, 0u, off, n);
        const Row32 *rows = seg.base[sg] + off;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            in[base + i] = to_spill(rows[i]);
        }
        base += n;
    }
}

__global__ void __launch_bounds__(256) spill_append_kernel(const Row32 *__restrict__ ovf, uint32_t n_ovf, SpillRow *__restrict__ in)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_ovf) return;
    in[i] = to_spill(ovf[i]);
}

// flags[i] = key_head << 32 | point_head
__global__ void __launch_bounds__(256) spill_flags_kernel(const SpillRow *__restrict__ rows, uint32_t M, uint64_t *__restrict__ flags)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const SpillRow r = rows[i];
    bool key_head = true, pt_head = true;
    if (i > 0) {
        const SpillRow p = rows[i - 1];
        key_head = !(p.a == r.a && p.b == r.b && p.proto == r.proto);
        pt_head = key_head || p.t != r.t;
    }
    flags[i] = ((uint64_t)(key_head ? 1u : 0u) << 32) | (pt_head ? 1u : 0u);
}

__device__ __forceinline__ uint32_t find_big(const uint32_t *__restrict__ big_base, uint32_t n_big, uint32_t i)
{
    uint32_t lo = 0, hi = n_big;          // largest j with big_base[j] <= i
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (big_base[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// PASS 0: series entries (n = 0) for key heads.  PASS 1: points (value reduced over duplicates),
// per-series counts, per-bucket totals.
template <int PASS>
__global__ void __launch_bounds__(256) spill_place_kernel(const SpillRow *__restrict__ rows, const uint64_t *__restrict__ scan,
                                                          uint32_t M, const uint32_t *__restrict__ offsets,
                                                          const uint32_t *__restrict__ big_list,
                                                          const uint32_t *__restrict__ big_base, uint32_t n_big,
                                                          SeriesEntry *__restrict__ entries, uint64_t *__restrict__ csr_v,
                                                          uint32_t *__restrict__ csr_t, uint32_t *__restrict__ nsb,
                                                          uint32_t *__restrict__ npb, int reducer)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const uint64_t sc = scan[i], prev = i ? scan[i - 1] : 0ull;
    const bool key_head = (sc >> 32) != (prev >> 32);
    const bool pt_head = (uint32_t)sc != (uint32_t)prev;
    const uint32_t j = find_big(big_base, n_big, i);
    const uint32_t bstart = big_base[j];
    const uint64_t bsc = bstart ? scan[bstart - 1] : 0ull;
    const uint32_t b = big_list[j];
    const uint32_t off_b = offsets[b];
    const uint32_t k = (uint32_t)(sc >> 32) - (uint32_t)(bsc >> 32) - 1u;       // series index inside the bucket
    const uint32_t p = (uint32_t)sc - (uint32_t)bsc - 1u;                        // point index inside the bucket
    SeriesEntry *ent = entries + off_b;
    if (PASS == 0) {
        if (key_head) {
            const SpillRow r = rows[i];
            SeriesEntry e;
            e.a = r.a; e.b = r.b; e.proto = r.proto; e.n = 0; e.off = off_b + p; e.pad = 1;   // pad = 1: no value permutation
            ent[k] = e;
        }
    } else {
        if (pt_head) {
            const SpillRow r = rows[i];
            unsigned long long val = r.value;
            for (uint32_t q = i + 1; q < M && (uint32_t)scan[q] == (uint32_t)sc; q++) {
                const unsigned long long y = rows[q].value;
                val = reducer == 0 ? (val > y ? val : y) : (val + y);
            }
            csr_t[off_b + p] = r.t;
            csr_v[off_b + p] = val;
            atomicAdd(&ent[k].n, 1u);
        }
        if (i + 1 == big_base[j + 1]) {
            nsb[b] = k + 1;
            npb[b] = p + 1;
        }
    }
}

static size_t cub_temp_bytes(uint64_t M)
{
    size_t sort_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, (const SpillRow *)nullptr, (SpillRow *)nullptr, (uint32_t)M,
                                   SpillDecomposer{});
    cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr, (uint32_t)M);
    return (sort_bytes > scan_bytes ? sort_bytes : scan_bytes) + 256;
}

static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

size_t spill_scratch_bytes(uint64_t M)
{
    if (M == 0) M = 1;
    return 2 * align256(M * sizeof(SpillRow)) + 2 * align256(M * 8) + align256(cub_temp_bytes(M)) + 1024;
}

cudaError_t run_spill(cudaStream_t st, const SegDesc &seg, SeriesEntry *entries, const uint32_t *offsets, const uint32_t *big_list,
                      const uint32_t *big_base, uint32_t n_big, uint64_t big_rows, void *scratch, size_t scratch_bytes,
                      uint64_t *csr_v, uint32_t *csr_t, uint32_t *nsb, uint32_t *npb, int reducer, int *launches,
                      const Row32 *ovf, uint32_t n_ovf)
{
    *launches = 0;
    if (n_big == 0 || big_rows == 0) return cudaSuccess;
    if (scratch_bytes < spill_scratch_bytes(big_rows)) return cudaErrorInvalidValue;
    const uint32_t M = (uint32_t)big_rows;
    char *base = static_cast<char *>(scratch);
    SpillRow *in = reinterpret_cast<SpillRow *>(base);
    base += align256((size_t)M * sizeof(SpillRow));
    SpillRow *out = reinterpret_cast<SpillRow *>(base);
    base += align256((size_t)M * sizeof(SpillRow));
    uint64_t *flags = reinterpret_cast<uint64_t *>(base);
    base += align256((size_t)M * 8);
    uint64_t *scan = reinterpret_cast<uint64_t *>(base);
    base += align256((size_t)M * 8);
    void *temp = base;
    size_t temp_bytes = cub_temp_bytes(M);

    spill_gather_kernel<<<n_big, 256, 0, st>>>(seg, big_list, big_base, in);
    if (seg.stride && seg.nseg == 1) {
        if ((uint64_t)n_big * seg.stride + n_ovf != big_rows) return cudaErrorInvalidValue;
        if (n_ovf) spill_append_kernel<<<(n_ovf + 255) / 256, 256, 0, st>>>(ovf, n_ovf, in + (size_t)n_big * seg.stride);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    size_t tb = temp_bytes;
    e = cub::DeviceRadixSort::SortKeys(temp, tb, (const SpillRow *)in, out, M, SpillDecomposer{}, st);
    if (e != cudaSuccess) return e;
    const uint32_t grid = (M + 255) / 256;
    spill_flags_kernel<<<grid, 256, 0, st>>>(out, M, flags);
    tb = temp_bytes;
    e = cub::DeviceScan::InclusiveSum(temp, tb, (const uint64_t *)flags, scan, M, st);
    if (e != cudaSuccess) return e;
    spill_place_kernel<0><<<grid, 256, 0, st>>>(out, scan, M, offsets, big_list, big_base, n_big, entries, csr_v, csr_t, nsb, npb,
                                               reducer);
    spill_place_kernel<1><<<grid, 256, 0, st>>>(out, scan, M, offsets, big_list, big_base, n_big, entries, csr_v, csr_t, nsb, npb,
                                               reducer);
    *launches = 4 + 8;        // ours + CUB's sort/scan passes (approximate)
    return cudaGetLastError();
}

// ----------------------------------------------------------------------------------------------------------------
// Capacity-class bucket lists in ascending bucket order.  The bucket scan appends to them with atomics, i.e. in no particular
// order; sorted, the CTAs that run at the same time work on neighbouring buckets -- neighbouring slots of the partition buffer
// (local, and over NVLink the peers') and neighbouring stretches of the csr arrays -- instead of 600 random places.
// ----------------------------------------------------------------------------------------------------------------
size_t sort_lists_scratch_bytes(uint32_t max_items)
{
    size_t bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)max_items, 0, 32);
    return align256(bytes) + align256((size_t)max_items * 4) + 256;
}

cudaError_t sort_bucket_list(cudaStream_t st, uint32_t *list, uint32_t n, int key_bits, void *scratch, size_t scratch_bytes)
{
    if (n < 2) return cudaSuccess;
    size_t temp_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, temp_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n, 0, key_bits);
    if (align256(temp_bytes) + (size_t)n * 4 > scratch_bytes) return cudaErrorInvalidValue;
    uint32_t *tmp = reinterpret_cast<uint32_t *>(static_cast<char *>(scratch) + align256(temp_bytes));
    cudaError_t e = cub::DeviceRadixSort::SortKeys(scratch, temp_bytes, (const uint32_t *)list, tmp, (int)n, 0, key_bits, st);
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(list, tmp, (size_t)n * 4, cudaMemcpyDeviceToDevice, st);
}

}  // namespace tad
