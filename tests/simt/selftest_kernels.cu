// TEST INFRASTRUCTURE ONLY -- small kernels in plain CUDA that exercise every primitive of simt_shim.h.  The CPU
// test-suite runs them through the shim and compares with numpy; `nvcc -c` of this same file (sm_100a) keeps the shim's
// signatures honest.  Nothing here is part of the product.
#ifndef GBRS_SIMT_EMULATION
#include <cuda_runtime.h>
#endif
#include <cstdint>

// block sum: xor-shuffle tree inside the warp, shared memory across warps; one atomicAdd(double) per block
__global__ void k_block_sum(const double* __restrict__ x, int64_t n, double* __restrict__ total, double* __restrict__ per_block) {
  __shared__ double part[32];
  double v = 0.0;
  for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x) v += x[i];
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    double w = threadIdx.x < (blockDim.x + 31) / 32 ? part[threadIdx.x] : 0.0;
    for (int m = 16; m > 0; m >>= 1) w += __shfl_down_sync(0xffffffffu, w, m);
    if (threadIdx.x == 0) {
      per_block[blockIdx.x] = w;
      atomicAdd(total, w);
    }
  }
}

// stream compaction of the positive entries: ballot + popc for the position inside the warp, one integer atomic per
// warp for its base, shfl to broadcast it
__global__ void k_compact_positive(const int32_t* __restrict__ x, int32_t n, int32_t* __restrict__ out, uint32_t* __restrict__ counter) {
  const int32_t rounds = (n + (int32_t) (gridDim.x * blockDim.x) - 1) / (int32_t) (gridDim.x * blockDim.x);
  for (int32_t r = 0; r < rounds; ++r) {  // every lane takes part in every round: full-mask collectives
    const int32_t i = (r * (int32_t) gridDim.x + (int32_t) blockIdx.x) * (int32_t) blockDim.x + (int32_t) threadIdx.x;
    const int keep = i < n && x[i] > 0;
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == 0 && mask) base = atomicAdd(counter, (uint32_t) __popc(mask));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) out[base + __popc(mask & ((1u << lane) - 1u))] = x[i];
  }
}

// inclusive scan inside groups of 8 lanes (shfl_up with width), max over the warp, any / all votes
__global__ void k_group_scan(const uint32_t* __restrict__ x, uint32_t* __restrict__ scan8, uint32_t* __restrict__ warp_max,
                             int32_t* __restrict__ votes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t v = x[i];
  for (int d = 1; d < 8; d <<= 1) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, v, d, 8);
    if ((threadIdx.x & 7) >= (unsigned) d) v += up;
  }
  scan8[i] = v;
  warp_max[i] = __reduce_max_sync(0xffffffffu, x[i]);
  votes[i] = (__any_sync(0xffffffffu, x[i] > 90u) ? 1 : 0) | (__all_sync(0xffffffffu, x[i] < 100u) ? 2 : 0);
  __syncwarp();
}

// histogram with double atomics on shared memory, flushed with global atomics; vector loads; hi/lo conversions
__global__ void k_histogram(const double2* __restrict__ xy, int32_t n, double* __restrict__ hist /* [16] */,
                            unsigned long long* __restrict__ checksum) {
  __shared__ double local[16];
  if (threadIdx.x < 16) local[threadIdx.x] = 0.0;
  __syncthreads();
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double2 p = __ldg(xy + i);
    const int bin = ((int) p.x) & 15;
    atomicAdd(&local[bin], p.y);
    const double back = __hiloint2double(__double2hiint(p.y), __double2loint(p.y));
    if (back != p.y) atomicAdd(checksum, 1ull << 40);
    atomicAdd(checksum, (unsigned long long) __popc((unsigned) bin));
  }
  __syncthreads();
  if (threadIdx.x < 16) atomicAdd(&hist[threadIdx.x], local[threadIdx.x]);
}
