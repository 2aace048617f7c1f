// NCCL plumbing for the multi-GPU exchange step, loaded with dlopen so that the library has
// no link-time NCCL dependency (a process that already loaded torch's bundled libnccl.so.2
// resolves to that copy; a Go host gets the system one).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace tad {

struct NcclComm {
    void *lib = nullptr;
    void *comm = nullptr;
    int world = 1, rank = 0;
    void *fn[8] = {nullptr};
};

int nccl_get_unique_id(void *out, size_t bytes);                 // 128 bytes
int nccl_comm_init(NcclComm *c, int world, int rank, const void *unique_id, size_t bytes);
void nccl_comm_destroy(NcclComm *c);
// byte-granular all-to-all: send_off/send_bytes/recv_off/recv_bytes have `world` entries
int nccl_alltoallv(NcclComm *c, const void *send, const uint64_t *send_off, const uint64_t *send_bytes, void *recv,
                   const uint64_t *recv_off, const uint64_t *recv_bytes, cudaStream_t st);
int nccl_allgather(NcclComm *c, const void *send, void *recv, size_t bytes_per_rank, cudaStream_t st);
const char *nccl_last_error();

}  // namespace tad
