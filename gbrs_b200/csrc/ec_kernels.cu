// Equivalence-class construction of `gbrs compress` on the GPU: reads with identical alignment patterns are merged
// into one class (count = sum of the read counts), classes numbered in order of first appearance.
// reference: src/gbrs/gbrs/emase_utils.py:46-72 -- a per-read Python loop that builds a string key from the sorted
// locus ids of every haplotype and sums the counts in a dict.
//
// Here a read is a row of (locus | hapmask << 24) pair words in ascending locus order, so "same key" is "same word
// sequence".  Pipeline (all on the caller's stream, device buffers from the caller):
//   k_ec_hash        64-bit hash of every read's word sequence                              (thread per read)
//   radix sort       reads by hash, stable => equal patterns adjacent, ascending read id    (cub::DeviceRadixSort)
//   k_ec_heads       exact comparison with the predecessor in sorted order: run heads; two different patterns with
//                    the same hash are counted as a collision (the caller retries with another seed)
//   prefix sum       run number of every sorted position                                     (cub::DeviceScan)
//   radix sort       runs by their first read id => class id = order of first appearance    (cub::DeviceRadixSort)
//   k_ec_assign      class of every read, class counts (integer-valued => the atomic sum is exact and order-free)
// The sorts and the scan are CUB (library code shipped with the toolkit); the hashing, the exact grouping and the
// class numbering are ours.
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "gbrs_em.h"

void gbrs_set_error(const std::string& s);  // em_kernels.cu

namespace {

constexpr int kEcThreads = 256;

#define EC_CUDA(call)                                                                                \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) {                                                                         \
      gbrs_set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                            \
      return GBRS_E_CUDA;                                                                            \
    }                                                                                                \
  } while (0)

inline int ec_grid(int64_t n) {
  int64_t b = (n + kEcThreads - 1) / kEcThreads;
  return (int) (b < 1 ? 1 : (b > (1 << 20) ? (1 << 20) : b));
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}

__global__ void __launch_bounds__(kEcThreads) k_ec_hash(int64_t n, const uint32_t* __restrict__ rowptr,
                                                        const uint32_t* __restrict__ pairs, uint64_t seed,
                                                        uint64_t* __restrict__ key, uint32_t* __restrict__ idx) {
  const int64_t stride = (int64_t) gridDim.x * blockDim.x;
  for (int64_t r = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    const uint32_t b = rowptr[r], e = rowptr[r + 1];
    uint64_t h = mix64(seed ^ (uint64_t) (e - b));
    for (uint32_t p = b; p < e; ++p) h = mix64(h ^ ((uint64_t) pairs[p] + 0x9E3779B97F4A7C15ull * (p - b + 1)));
    key[r] = h;
    idx[r] = (uint32_t) r;
  }
}

// head[i] = 1 if the read at sorted position i starts a new run of identical patterns
__global__ void __launch_bounds__(kEcThreads) k_ec_heads(int64_t n, const uint32_t* __restrict__ rowptr,
                                                         const uint32_t* __restrict__ pairs, const uint64_t* __restrict__ key,
                                                         const uint32_t* __restrict__ idx, uint32_t* __restrict__ head,
                                                         unsigned long long* __restrict__ collisions) {
  const int64_t stride = (int64_t) gridDim.x * blockDim.x;
  for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint32_t hd = 1;
    if (i > 0 && key[i] == key[i - 1]) {
      const uint32_t r = idx[i], q = idx[i - 1];
      const uint32_t br = rowptr[r], er = rowptr[r + 1], bq = rowptr[q], eq = rowptr[q + 1];
      bool same = (er - br) == (eq - bq);
      for (uint32_t k = 0; same && k < er - br; ++k) same = pairs[br + k] == pairs[bq + k];
      hd = same ? 0u : 1u;
      if (!same) atomicAdd(collisions, 1ull);  // equal hash, different pattern: grouping by adjacency is unsafe
    }
    head[i] = hd;
  }
}

// run[i] = inclusive prefix sum of head (1-based run number).  First read id of every run (the head's read: the sort
// is stable and started from ascending read ids), and the identity list of run ids for the second sort.
__global__ void __launch_bounds__(kEcThreads) k_ec_reps(int64_t n, const uint32_t* __restrict__ head,
                                                        const uint32_t* __restrict__ run, const uint32_t* __restrict__ idx,
                                                        uint32_t* __restrict__ rep, uint32_t* __restrict__ run_id) {
  const int64_t stride = (int64_t) gridDim.x * blockDim.x;
  for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    if (head[i]) {
      rep[run[i] - 1] = idx[i];
      run_id[run[i] - 1] = run[i] - 1;
    }
}

__global__ void __launch_bounds__(kEcThreads) k_ec_rank(int64_t n_runs, const uint32_t* __restrict__ run_sorted,
                                                        uint32_t* __restrict__ rank_of_run) {
  const int64_t stride = (int64_t) gridDim.x * blockDim.x;
  for (int64_t k = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; k < n_runs; k += stride) rank_of_run[run_sorted[k]] = (uint32_t) k;
}

__global__ void __launch_bounds__(kEcThreads) k_ec_assign(int64_t n, const uint32_t* __restrict__ run,
                                                          const uint32_t* __restrict__ idx,
                                                          const uint32_t* __restrict__ rank_of_run,
                                                          const double* __restrict__ count, uint32_t* __restrict__ class_of_read,
                                                          double* __restrict__ class_count) {
  const int64_t stride = (int64_t) gridDim.x * blockDim.x;
  for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t r = idx[i], c = rank_of_run[run[i] - 1];
    class_of_read[r] = c;
    atomicAdd(class_count + c, count ? count[r] : 1.0);  // emase_utils.py:58-59: a missing count vector means ones
  }
}

struct EcLayout {
  size_t key_a, key_b, idx_a, idx_b, head, run, rep_a, rep_b, rid_a, rid_b, rank, coll, cub, total;
  size_t cub_bytes;
};

inline size_t align_up(size_t x) { return (x + 255) / 256 * 256; }

int ec_layout(int64_t n, EcLayout* L) {
  size_t s1 = 0, s2 = 0, s3 = 0;
  EC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, s1, (const uint64_t*) nullptr, (uint64_t*) nullptr, (const uint32_t*) nullptr,
                                          (uint32_t*) nullptr, n));
  EC_CUDA(cub::DeviceScan::InclusiveSum(nullptr, s2, (const uint32_t*) nullptr, (uint32_t*) nullptr, n));
  EC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, s3, (const uint32_t*) nullptr, (uint32_t*) nullptr, (const uint32_t*) nullptr,
                                          (uint32_t*) nullptr, n));
  L->cub_bytes = s1 > s2 ? (s1 > s3 ? s1 : s3) : (s2 > s3 ? s2 : s3);
  const size_t m = (size_t) (n > 0 ? n : 1);
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t at = o; o += align_up(bytes); return at; };
  L->key_a = take(8 * m); L->key_b = take(8 * m);
  L->idx_a = take(4 * m); L->idx_b = take(4 * m);
  L->head = take(4 * m);  L->run = take(4 * m);
  L->rep_a = take(4 * m); L->rep_b = take(4 * m);
  L->rid_a = take(4 * m); L->rid_b = take(4 * m);
  L->rank = take(4 * m);  L->coll = take(8);
  L->cub = take(L->cub_bytes);
  L->total = o;
  return GBRS_OK;
}

}  // namespace

extern "C" int gbrs_ec_workspace_bytes(int64_t n_reads, int64_t* bytes) {
  if (!bytes || n_reads < 0) { gbrs_set_error("gbrs_ec_workspace_bytes: bad argument"); return GBRS_E_ARG; }
  EcLayout L;
  if (int rc = ec_layout(n_reads, &L)) return rc;
  *bytes = (int64_t) L.total;
  return GBRS_OK;
}

extern "C" int gbrs_ec_build(int64_t n_reads, const uint32_t* rowptr_dev, const uint32_t* pairs_dev, const double* count_dev,
                             uint64_t seed, uint32_t* class_of_read_dev, uint32_t* first_read_dev, double* class_count_dev,
                             void* workspace_dev, int64_t workspace_bytes, void* stream, int64_t* n_classes_out,
                             int64_t* collisions_out) {
  if (n_reads < 0 || n_reads >= (int64_t(1) << 32) || !rowptr_dev || !class_of_read_dev || !first_read_dev ||
      !class_count_dev || !workspace_dev || !n_classes_out || !collisions_out) {
    gbrs_set_error("gbrs_ec_build: bad argument"); return GBRS_E_ARG;
  }
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { gbrs_set_error("gbrs_ec_build: no CUDA device (there is no CPU fallback)"); return GBRS_E_CUDA; }
  *n_classes_out = 0;
  *collisions_out = 0;
  if (n_reads == 0) return GBRS_OK;
  if (!pairs_dev) { gbrs_set_error("gbrs_ec_build: null pair words"); return GBRS_E_ARG; }
  EcLayout L;
  if (int rc = ec_layout(n_reads, &L)) return rc;
  if ((int64_t) L.total > workspace_bytes) { gbrs_set_error("gbrs_ec_build: workspace too small"); return GBRS_E_ARG; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  char* w = static_cast<char*>(workspace_dev);
  auto* key_a = reinterpret_cast<uint64_t*>(w + L.key_a); auto* key_b = reinterpret_cast<uint64_t*>(w + L.key_b);
  auto* idx_a = reinterpret_cast<uint32_t*>(w + L.idx_a); auto* idx_b = reinterpret_cast<uint32_t*>(w + L.idx_b);
  auto* head = reinterpret_cast<uint32_t*>(w + L.head);   auto* run = reinterpret_cast<uint32_t*>(w + L.run);
  auto* rep_a = reinterpret_cast<uint32_t*>(w + L.rep_a); auto* rep_b = reinterpret_cast<uint32_t*>(w + L.rep_b);
  auto* rid_a = reinterpret_cast<uint32_t*>(w + L.rid_a); auto* rid_b = reinterpret_cast<uint32_t*>(w + L.rid_b);
  auto* rank = reinterpret_cast<uint32_t*>(w + L.rank);
  auto* coll = reinterpret_cast<unsigned long long*>(w + L.coll);
  void* cub_tmp = w + L.cub;
  size_t cub_bytes = L.cub_bytes;
  const int g = ec_grid(n_reads);

  EC_CUDA(cudaMemsetAsync(coll, 0, 8, s));
  EC_CUDA(cudaMemsetAsync(class_count_dev, 0, sizeof(double) * (size_t) n_reads, s));
  k_ec_hash<<<g, kEcThreads, 0, s>>>(n_reads, rowptr_dev, pairs_dev, seed, key_a, idx_a);
  EC_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, key_a, key_b, idx_a, idx_b, n_reads, 0, 64, s));
  k_ec_heads<<<g, kEcThreads, 0, s>>>(n_reads, rowptr_dev, pairs_dev, key_b, idx_b, head, coll);
  cub_bytes = L.cub_bytes;
  EC_CUDA(cub::DeviceScan::InclusiveSum(cub_tmp, cub_bytes, head, run, n_reads, s));
  uint32_t n_runs = 0;
  unsigned long long n_coll = 0;
  EC_CUDA(cudaMemcpyAsync(&n_runs, run + (n_reads - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  EC_CUDA(cudaMemcpyAsync(&n_coll, coll, sizeof(n_coll), cudaMemcpyDeviceToHost, s));
  EC_CUDA(cudaStreamSynchronize(s));
  *collisions_out = (int64_t) n_coll;
  if (n_coll) return GBRS_OK;  // the caller retries with another seed
  k_ec_reps<<<g, kEcThreads, 0, s>>>(n_reads, head, run, idx_b, rep_a, rid_a);
  cub_bytes = L.cub_bytes;
  EC_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, rep_a, rep_b, rid_a, rid_b, (int64_t) n_runs, 0, 32, s));
  k_ec_rank<<<ec_grid(n_runs), kEcThreads, 0, s>>>((int64_t) n_runs, rid_b, rank);
  k_ec_assign<<<g, kEcThreads, 0, s>>>(n_reads, run, idx_b, rank, count_dev, class_of_read_dev, class_count_dev);
  EC_CUDA(cudaMemcpyAsync(first_read_dev, rep_b, sizeof(uint32_t) * (size_t) n_runs, cudaMemcpyDeviceToDevice, s));
  EC_CUDA(cudaGetLastError());
  EC_CUDA(cudaStreamSynchronize(s));
  *n_classes_out = (int64_t) n_runs;
  return GBRS_OK;
}
