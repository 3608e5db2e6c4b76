// TEST INFRASTRUCTURE ONLY -- the slice of the CUDA runtime API that gbrs_b200/csrc/em_kernels.cu's host code calls,
// implemented for the host SIMT shim: "device" memory is host memory, streams are synchronous, graph capture reports
// "not supported" (the library then falls back to plain launches, as it does under GBRS_NO_GRAPH), occupancy queries
// return a small constant so that the persistent grids stay a handful of blocks.
#pragma once
#include <chrono>
#include <cstdlib>
#include <cstring>

#include "simt_shim.h"

#define GBRS_SIMT_EMULATION 1  // em_kernels.cu: host stand-ins for the bulk-copy / mbarrier PTX, static "dynamic" smem

using std::isfinite;
using std::isnan;
using std::isinf;

typedef void* cudaStream_t;
typedef int cudaError_t;
typedef void* cudaGraph_t;
typedef void* cudaGraphExec_t;
struct simt_event { std::chrono::steady_clock::time_point t; };
typedef simt_event* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorNotSupported = 801 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaStreamCaptureModeThreadLocal = 1, cudaDevAttrMultiProcessorCount = 16 };
#define cudaStreamLegacy ((cudaStream_t) 0x1)

inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "not supported by the host SIMT shim"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* dev) { *dev = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) {  // the SM count: 2, or $GBRS_SIMT_SMS
  const char* e = std::getenv("GBRS_SIMT_SMS");
  *v = e && std::atoi(e) > 0 ? std::atoi(e) : 2;
  return cudaSuccess;
}
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class K> inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }
template <class K> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, K, int, size_t) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(dst, src, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* dst, int v, size_t n, cudaStream_t) { std::memset(dst, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t) 0x2; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, int) { return cudaErrorNotSupported; }
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) { *g = nullptr; return cudaErrorNotSupported; }
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t, unsigned long long) { *e = nullptr; return cudaErrorNotSupported; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new simt_event(); return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }

// `kernel<<<grid, block, 0, stream>>>(args)` is rewritten (tests/simt_em.py) to SIMT_LAUNCH(grid, block, kernel(args))
// (variadic: template argument lists of the kernel name carry top-level commas)
#define SIMT_LAUNCH(grid, block, ...) simt_launch((unsigned) (grid), (unsigned) (block), [=] { __VA_ARGS__; })
// Inline PTX of the multi-GPU exchange.  The system-scope flag accesses become C++ atomics (the flag store is issued
// after a system fence in the kernel: a release store here), the system-scope f64x2 accesses plain accesses (the flags
// order them), so that ranks emulated as concurrently running library instances exchange through shared host memory;
// the NVSwitch multimem instructions have no host meaning and trap.
#define SIMT_PTX_UNSUPPORTED() __builtin_trap()
inline unsigned simt_ld_acquire_u32(const unsigned* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline unsigned long long simt_globaltimer_ns() {
  return (unsigned long long) std::chrono::duration_cast<std::chrono::nanoseconds>(
             std::chrono::steady_clock::now().time_since_epoch()).count();
}
inline void simt_st_release_u32(unsigned* p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
// tag form of the exchange: a double and its "arrived" bit travel in one indivisible 8-byte access
inline unsigned long long simt_ld_acquire_u64(const double* p) {
  return __atomic_load_n(reinterpret_cast<const unsigned long long*>(p), __ATOMIC_ACQUIRE);
}
inline void simt_st_release_u64(double* p, unsigned long long v) {
  __atomic_store_n(reinterpret_cast<unsigned long long*>(p), v, __ATOMIC_RELEASE);
}
// rcp.approx.ftz.f64: a reciprocal good to ~20 bits (the seed of fast_div's Newton steps) -- modelled as the exact
// reciprocal with the low 32 mantissa bits cleared, so that the refinement steps have real work to do
inline double simt_rcp_approx(double s) {
  double r = 1.0 / s;
  unsigned long long b;
  std::memcpy(&b, &r, 8);
  b &= 0xffffffff00000000ull;
  std::memcpy(&r, &b, 8);
  return r;
}
