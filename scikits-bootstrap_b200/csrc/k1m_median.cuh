// K1m / K3m — column-batched resample median (stub until the rank-space kernels land).
#pragma once
#include "common.cuh"

namespace b200boot {

inline size_t median_ws_bytes(int64_t, int64_t, int64_t) { return 0; }

inline int median_resample(const double*, int64_t, int64_t, int64_t, uint64_t, int64_t, int64_t, double*,
                           void*, size_t, cudaStream_t) {
  return fail(B200BOOT_EUNSUPPORTED, "median kernels not built yet");
}
inline int median_jackknife(const double*, int64_t, int64_t, int64_t, double*, void*, size_t,
                            cudaStream_t) {
  return fail(B200BOOT_EUNSUPPORTED, "median kernels not built yet");
}

}  // namespace b200boot
