// TEST INFRASTRUCTURE ONLY -- host stand-in for the one cub::DeviceRadixSort entry point ec_kernels.cu calls (library
// code on the GPU; here a stable sort), with CUB's two-phase "query the temporary size, then run" protocol.
#pragma once
#include <algorithm>
#include <cstddef>
#include <numeric>
#include <vector>

#include "cuda_runtime_fake.h"

namespace cub {
struct DeviceRadixSort {
  template <class K, class V, class N>
  static cudaError_t SortPairs(void* tmp, size_t& tmp_bytes, const K* keys_in, K* keys_out, const V* vals_in, V* vals_out,
                               N n, int begin_bit = 0, int end_bit = sizeof(K) * 8, cudaStream_t = nullptr) {
    if (tmp == nullptr) { tmp_bytes = 256; return cudaSuccess; }
    const K mask = end_bit - begin_bit >= (int) (sizeof(K) * 8) ? ~K(0) : (((K(1) << (end_bit - begin_bit)) - 1) << begin_bit);
    std::vector<size_t> order((size_t) n);
    std::iota(order.begin(), order.end(), size_t(0));
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return (keys_in[a] & mask) < (keys_in[b] & mask); });
    for (size_t i = 0; i < (size_t) n; ++i) { keys_out[i] = keys_in[order[i]]; vals_out[i] = vals_in[order[i]]; }
    return cudaSuccess;
  }
};
}  // namespace cub
