// placeholder until the oversized-bucket path lands (next commit)
#include "tad_kernels.h"
namespace tad {
size_t spill_scratch_bytes(uint64_t) { return 256; }
cudaError_t run_spill(cudaStream_t, Row32 *, const uint32_t *, const uint32_t *, uint32_t, uint64_t, void *, size_t, uint64_t *,
                      uint32_t *, uint32_t *, uint32_t *, int, int *) { return cudaErrorNotSupported; }
cudaError_t launch_detect_dbscan(cudaStream_t, const Row32 *, const uint32_t *, const uint32_t *, uint32_t, uint32_t,
                                 const uint64_t *, const uint32_t *, double *, uint32_t *, const OutCols &, uint32_t, uint32_t *,
                                 int) { return cudaErrorNotSupported; }
}
