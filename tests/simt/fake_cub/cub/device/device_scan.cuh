// TEST INFRASTRUCTURE ONLY -- host stand-in for cub::DeviceScan::InclusiveSum (see device_radix_sort.cuh).
#pragma once
#include <cstddef>

#include "cuda_runtime_fake.h"

namespace cub {
struct DeviceScan {
  template <class In, class Out, class N>
  static cudaError_t InclusiveSum(void* tmp, size_t& tmp_bytes, const In* in, Out* out, N n, cudaStream_t = nullptr) {
    if (tmp == nullptr) { tmp_bytes = 256; return cudaSuccess; }
    Out acc = 0;
    for (size_t i = 0; i < (size_t) n; ++i) { acc += in[i]; out[i] = acc; }
    return cudaSuccess;
  }
};
}  // namespace cub
