// TEST INFRASTRUCTURE ONLY -- a minimal host stand-in for the CUDA execution model, enough to run the kernels of this
// repository unchanged on a machine without a GPU:
//   * one OS thread per CUDA thread of a block (blocks run one after the other), threadIdx / blockIdx / blockDim /
//     gridDim, static `__shared__` memory (a function-local static: one copy, shared by the threads of the running block);
//   * __syncthreads = a pthread barrier over the block; warp-level primitives (__shfl*_sync, __ballot_sync, __any / __all,
//     __reduce_*_sync, __syncwarp) = an exchange slot per lane + a pthread barrier over the 32 threads of the warp --
//     FULL-mask semantics only: every lane of the warp must reach the call (the kernels here use 0xffffffff);
//   * cp.async (__pipeline_memcpy_async) copies are DEFERRED until __pipeline_wait_prior: a kernel that forgets the wait
//     reads stale shared memory here too;
//   * atomicAdd / atomicMax / atomicCAS / atomicExch on 32- and 64-bit integers and double (GCC __atomic builtins; the
//     double add is a CAS loop, like the hardware's shared-memory path), __threadfence* = a sequentially consistent fence;
//   * read-only / streaming load intrinsics (__ldg, __ldcs, __ldcg, __ldca) = plain loads; the vector types the kernels
//     use (double2, uint2, uint4, ulonglong2, int2, int4) with their make_* constructors; bit and conversion helpers.
// Built with -fsanitize=thread it reports shared / global memory accesses of a kernel that no barrier orders.
// Not modelled: warp divergence around partial-mask collectives, memory-model subtleties weaker than x86-TSO, timing.
#pragma once
#include <pthread.h>
#include <sched.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __grid_constant__

struct simt_dim3 { unsigned x = 1, y = 1, z = 1; };
static thread_local simt_dim3 threadIdx;
static simt_dim3 blockIdx, blockDim, gridDim;
static constexpr int warpSize = 32;

// ---- vector types -------------------------------------------------------------------------------------------------
struct alignas(16) double2 { double x, y; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
inline double2 make_double2(double x, double y) { return {x, y}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }
inline int2 make_int2(int x, int y) { return {x, y}; }
inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }
inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return {x, y}; }

// ---- block barrier, cp.async ----------------------------------------------------------------------------------------
static pthread_barrier_t simt_barrier;
static std::vector<pthread_barrier_t> simt_warp_barrier;   // one per warp of the running block
static std::vector<unsigned long long> simt_slot;          // one exchange slot per thread of the running block

struct simt_copy { void* dst; const void* src; size_t n; };
static thread_local std::vector<simt_copy> simt_pending;

inline void __syncthreads() { pthread_barrier_wait(&simt_barrier); }
inline void __pipeline_memcpy_async(void* dst, const void* src, size_t n) { simt_pending.push_back({dst, src, n}); }
inline void __pipeline_commit() {}
inline void __pipeline_wait_prior(int) {
  for (const simt_copy& c : simt_pending) std::memcpy(c.dst, c.src, c.n);
  simt_pending.clear();
}

// ---- warp collectives (full mask) -----------------------------------------------------------------------------------
inline int simt_lane() { return (int) (threadIdx.x & 31u); }
inline int simt_warp() { return (int) (threadIdx.x >> 5); }
inline int simt_warp_lanes() {  // the last warp of a block may be partial
  const unsigned first = threadIdx.x & ~31u;
  const unsigned left = blockDim.x - first;
  return (int) (left < 32u ? left : 32u);
}
inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&simt_warp_barrier[simt_warp()]); }

template <class T>
inline T simt_exchange(T v, int src_lane) {  // every lane publishes v, then reads the value of `src_lane`
  static_assert(sizeof(T) <= 8, "shuffles move up to 64 bits");
  unsigned long long bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  const unsigned base = threadIdx.x & ~31u;
  simt_slot[threadIdx.x] = bits;
  __syncwarp();
  T out = v;
  if (src_lane >= 0 && src_lane < simt_warp_lanes()) {
    const unsigned long long got = simt_slot[base + (unsigned) src_lane];
    std::memcpy(&out, &got, sizeof(T));
  }
  __syncwarp();  // slots may be overwritten by the next collective
  return out;
}
template <class T> inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  const int lane = simt_lane();
  return simt_exchange(v, (lane / width) * width + (src % width));
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
  const int lane = simt_lane(), src = lane ^ m;
  return simt_exchange(v, (src / width == lane / width) ? src : lane);
}
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
  const int lane = simt_lane(), src = lane + (int) d;
  return simt_exchange(v, (src / width == lane / width) ? src : lane);
}
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
  const int lane = simt_lane(), src = lane - (int) d;
  return simt_exchange(v, (src >= 0 && src / width == lane / width) ? src : lane);
}
inline unsigned __ballot_sync(unsigned, int pred) {
  const unsigned base = threadIdx.x & ~31u;
  simt_slot[threadIdx.x] = pred ? 1ull : 0ull;
  __syncwarp();
  unsigned out = 0;
  for (int l = 0; l < simt_warp_lanes(); ++l) out |= (unsigned) (simt_slot[base + (unsigned) l] & 1ull) << l;
  __syncwarp();
  return out;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }
inline int __all_sync(unsigned m, int pred) {
  const unsigned full = simt_warp_lanes() == 32 ? 0xffffffffu : ((1u << simt_warp_lanes()) - 1u);
  return __ballot_sync(m, pred) == full;
}
template <class T, class F> inline T simt_warp_reduce(T v, F f) {
  const unsigned base = threadIdx.x & ~31u;
  unsigned long long bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  simt_slot[threadIdx.x] = bits;
  __syncwarp();
  T acc{};
  for (int l = 0; l < simt_warp_lanes(); ++l) {
    T x;
    const unsigned long long got = simt_slot[base + (unsigned) l];
    std::memcpy(&x, &got, sizeof(T));
    acc = l == 0 ? x : f(acc, x);
  }
  __syncwarp();
  return acc;
}
inline unsigned __reduce_max_sync(unsigned, unsigned v) { return simt_warp_reduce(v, [](unsigned a, unsigned b) { return a > b ? a : b; }); }
inline unsigned __reduce_min_sync(unsigned, unsigned v) { return simt_warp_reduce(v, [](unsigned a, unsigned b) { return a < b ? a : b; }); }
inline unsigned __reduce_add_sync(unsigned, unsigned v) { return simt_warp_reduce(v, [](unsigned a, unsigned b) { return a + b; }); }
inline int __reduce_max_sync(unsigned, int v) { return simt_warp_reduce(v, [](int a, int b) { return a > b ? a : b; }); }
inline int __reduce_add_sync(unsigned, int v) { return simt_warp_reduce(v, [](int a, int b) { return a + b; }); }

// ---- atomics, fences ------------------------------------------------------------------------------------------------
template <class T> inline T simt_atomic_add_int(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicAdd(int* p, int v) { return simt_atomic_add_int(p, v); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return simt_atomic_add_int(p, v); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return simt_atomic_add_int(p, v); }
inline double atomicAdd(double* p, double v) {
  unsigned long long* q = reinterpret_cast<unsigned long long*>(p);
  unsigned long long old = __atomic_load_n(q, __ATOMIC_SEQ_CST), want;
  double cur;
  do {
    std::memcpy(&cur, &old, 8);
    const double next = cur + v;
    std::memcpy(&want, &next, 8);
  } while (!__atomic_compare_exchange_n(q, &old, want, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
  return cur;
}
template <class T> inline T atomicCAS(T* p, T cmp, T val) {
  __atomic_compare_exchange_n(p, &cmp, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return cmp;
}
template <class T> inline T atomicExch(T* p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <class T> inline T atomicMax(T* p, T v) {
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
template <class T> inline T atomicMin(T* p, T v) {
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
template <class T> inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __nanosleep(unsigned) { sched_yield(); }

// ---- loads, bit helpers, conversions ----------------------------------------------------------------------------------
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcs(const T* p) { return *p; }
template <class T> inline T __ldcg(const T* p) { return *p; }
template <class T> inline T __ldca(const T* p) { return *p; }
template <class T> inline void __stcs(T* p, T v) { *p = v; }
template <class T> inline void __stcg(T* p, T v) { *p = v; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned) v) : 32; }
inline unsigned __brev(unsigned v) {
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
  return r;
}
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, sizeof d); return d; }
inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, sizeof v); return v; }
inline double __hiloint2double(int hi, int lo) {
  const unsigned long long v = ((unsigned long long) (unsigned) hi << 32) | (unsigned) lo;
  double d; std::memcpy(&d, &v, sizeof d); return d;
}
inline int __double2hiint(double d) { unsigned long long v; std::memcpy(&v, &d, 8); return (int) (v >> 32); }
inline int __double2loint(double d) { unsigned long long v; std::memcpy(&v, &d, 8); return (int) (v & 0xffffffffull); }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __drcp_rn(double a) { return 1.0 / a; }

// ---- launch -------------------------------------------------------------------------------------------------------------
template <class F>
void simt_launch(unsigned grid, unsigned block, F kernel) {
  gridDim.x = grid;
  blockDim.x = block;
  const unsigned warps = (block + 31u) / 32u;
  simt_slot.assign(block, 0ull);
  for (unsigned b = 0; b < grid; ++b) {
    blockIdx.x = b;
    pthread_barrier_init(&simt_barrier, nullptr, block);
    simt_warp_barrier.resize(warps);
    for (unsigned w = 0; w < warps; ++w) {
      const unsigned lanes = (w + 1u) * 32u <= block ? 32u : block - w * 32u;
      pthread_barrier_init(&simt_warp_barrier[w], nullptr, lanes);
    }
    std::vector<std::thread> ts;
    for (unsigned t = 0; t < block; ++t)
      ts.emplace_back([=] { threadIdx.x = t; simt_pending.clear(); kernel(); });
    for (auto& t : ts) t.join();
    pthread_barrier_destroy(&simt_barrier);
    for (unsigned w = 0; w < warps; ++w) pthread_barrier_destroy(&simt_warp_barrier[w]);
  }
}
