// TEST INFRASTRUCTURE ONLY -- a minimal host stand-in for the CUDA execution model, enough to run a kernel that uses
// threadIdx / blockIdx, static shared memory, __syncthreads and the cp.async pipeline primitives: one OS thread per
// CUDA thread of a block (blocks run one after the other), a pthread barrier for __syncthreads, and cp.async copies that
// are DEFERRED until __pipeline_wait_prior (a kernel that forgets the wait reads stale shared memory here too).
// Built with -fsanitize=thread it also reports shared / global memory races between barrier phases.
#pragma once
#include <pthread.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static

struct simt_dim3 { unsigned x = 1, y = 1, z = 1; };
static thread_local simt_dim3 threadIdx;
static simt_dim3 blockIdx, blockDim, gridDim;
static pthread_barrier_t simt_barrier;

struct simt_copy { void* dst; const void* src; size_t n; };
static thread_local std::vector<simt_copy> simt_pending;

inline void __syncthreads() { pthread_barrier_wait(&simt_barrier); }
inline void __pipeline_memcpy_async(void* dst, const void* src, size_t n) { simt_pending.push_back({dst, src, n}); }
inline void __pipeline_commit() {}
inline void __pipeline_wait_prior(int) {
  for (const simt_copy& c : simt_pending) std::memcpy(c.dst, c.src, c.n);
  simt_pending.clear();
}
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, sizeof d); return d; }

template <class F>
void simt_launch(unsigned grid, unsigned block, F kernel) {
  gridDim.x = grid;
  blockDim.x = block;
  for (unsigned b = 0; b < grid; ++b) {
    blockIdx.x = b;
    pthread_barrier_init(&simt_barrier, nullptr, block);
    std::vector<std::thread> ts;
    for (unsigned t = 0; t < block; ++t)
      ts.emplace_back([=] { threadIdx.x = t; simt_pending.clear(); kernel(); });
    for (auto& t : ts) t.join();
    pthread_barrier_destroy(&simt_barrier);
  }
}
