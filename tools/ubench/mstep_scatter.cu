// Micro-benchmark: the SCATTER formulation of the M-step (single pass, atomics) on the real C2 locus distribution,
// against which the atomic-free gather formulations of gbrs_b200/csrc/em_kernels.cu are judged.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mstep_scatter tools/ubench/mstep_scatter.cu
//   ./mstep_scatter [n_classes = 5000000] [T = 80000]
//
// Workload: the canonical generator of SURVEY.md 8(d) (Lomax(1.2) base locus scaled by T/50, 1 + Poisson(1.5) loci per
// class within +0..5 of the base, mask 0xFF w.p. 0.5 else U{1..255}), classes sorted by their smallest locus -- the order
// the packer gives them.  The E-step side (the normaliser gather) is left out on purpose: every class scatters the same
// weight 1.0, so the kernels time the M-step accumulation alone.
//
//   A  hap_red        one RED.F64 per (pair, set haplotype bit) into acc[T][8]            (~nnz atomics; the naive form)
//   B  nibble_red     two RED.F64 per pair into the nibble-bucket table B[T][32]           (2 * pairs atomics)
//   C  nibble_red_agg B with warp pre-aggregation: lanes of a warp that hit the same bucket (classes are sorted, so
//                     neighbours mostly do) are matched with __match_any_sync, summed by shuffles, one RED per group
//   D  smem_tile      per-CTA privatised accumulators: a CTA takes a contiguous class range, accumulates the nibble
//                     buckets of a window of loci in shared memory (fp64 shared atomics = CAS loops on sm_100),
//                     out-of-window pairs go to global REDs, the window is flushed with one RED per touched bucket
//   E  smem_tile_agg  D with the warp pre-aggregation of C in front of the shared-memory atomics
// Every variant is checked against the exact per-bucket totals computed on the host (integer-valued, so the sums are
// exact in any order).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>
#include <vector>

#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("%s: %s\n", #x, cudaGetErrorString(e_)); std::exit(1); } } while (0)

__device__ __forceinline__ void red_add(double* p, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// classes as CSR over pair words (locus | mask << 24)
__global__ void __launch_bounds__(256) k_hap_red(const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ pairs,
                                                 int64_t n, double* __restrict__ acc) {
  for (int64_t c = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (int64_t) gridDim.x * blockDim.x) {
    const uint32_t b = rowptr[c], e = rowptr[c + 1];
    for (uint32_t p = b; p < e; ++p) {
      const uint32_t w = pairs[p], t = w & 0xFFFFFFu, m = w >> 24;
      for (int h = 0; h < 8; ++h)
        if ((m >> h) & 1u) red_add(acc + (size_t) t * 8 + h, 1.0);
    }
  }
}

__global__ void __launch_bounds__(256) k_nibble_red(const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ pairs,
                                                    int64_t n, double* __restrict__ B) {
  for (int64_t c = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (int64_t) gridDim.x * blockDim.x) {
    const uint32_t b = rowptr[c], e = rowptr[c + 1];
    for (uint32_t p = b; p < e; ++p) {
      const uint32_t w = pairs[p], t = w & 0xFFFFFFu, m = w >> 24;
      if (m & 15u) red_add(B + (size_t) t * 32 + (m & 15u), 1.0);
      if (m >> 4) red_add(B + (size_t) t * 32 + 16 + (m >> 4), 1.0);
    }
  }
}

// all lanes call; lanes with the same key are summed, the lowest lane of each group gets the total (others get 0 and
// `leader` false).  key == 0xFFFFFFFF = nothing to add.
__device__ __forceinline__ double warp_agg(uint32_t key, double v, bool& leader) {
  const unsigned peers = __match_any_sync(0xFFFFFFFFu, key);
  const int lane = threadIdx.x & 31;
  leader = (__ffs(peers) - 1) == lane && key != 0xFFFFFFFFu;
  double s = 0.0;
  // every lane walks the peer list of ITS group; groups are disjoint, so 32 rounds of a uniform shuffle do all of them
  unsigned rest = peers;
#pragma unroll 1
  while (__any_sync(0xFFFFFFFFu, rest != 0u)) {
    const int src = rest ? __ffs(rest) - 1 : lane;
    const double x = __shfl_sync(0xFFFFFFFFu, v, src);
    if (rest) { s += x; rest &= rest - 1; }
  }
  return s;
}

__global__ void __launch_bounds__(256) k_nibble_red_agg(const uint32_t* __restrict__ rowptr,
                                                        const uint32_t* __restrict__ pairs, int64_t n,
                                                        double* __restrict__ B) {
  const int64_t stride = (int64_t) gridDim.x * blockDim.x;
  const int64_t rounds = (n + stride - 1) / stride;
  int64_t c = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t r = 0; r < rounds; ++r, c += stride) {
    uint32_t b = 0, e = 0;
    if (c < n) { b = rowptr[c]; e = rowptr[c + 1]; }
    const uint32_t kmax = __reduce_max_sync(0xFFFFFFFFu, e - b);
    for (uint32_t j = 0; j < kmax; ++j) {
      uint32_t klo = 0xFFFFFFFFu, khi = 0xFFFFFFFFu;
      if (b + j < e) {
        const uint32_t w = pairs[b + j], t = w & 0xFFFFFFu, m = w >> 24;
        if (m & 15u) klo = t * 32 + (m & 15u);
        if (m >> 4) khi = t * 32 + 16 + (m >> 4);
      }
      bool lead;
      double s = warp_agg(klo, 1.0, lead);
      if (lead) red_add(B + klo, s);
      s = warp_agg(khi, 1.0, lead);
      if (lead) red_add(B + khi, s);
    }
  }
}

// per-CTA privatised window.  class range of CTA i: [i * cpb, (i + 1) * cpb); window = [locus of the first pair of the
// first class, + WIN)
template <int WIN, bool AGG>
__global__ void __launch_bounds__(256) k_smem_tile(const uint32_t* __restrict__ rowptr, const uint32_t* __restrict__ pairs,
                                                   int64_t n, int cpb, int T, double* __restrict__ B) {
  __shared__ double win[WIN * 32];
  for (int64_t c0 = (int64_t) blockIdx.x * cpb; c0 < n; c0 += (int64_t) gridDim.x * cpb) {
    for (int i = threadIdx.x; i < WIN * 32; i += blockDim.x) win[i] = 0.0;
    const uint32_t t0 = pairs[rowptr[c0]] & 0xFFFFFFu;  // classes are sorted by smallest locus; pairs ascending in a class
    __syncthreads();
    const int64_t c1 = c0 + cpb < n ? c0 + cpb : n;
    for (int64_t cb = c0; cb < c1; cb += blockDim.x) {
      const int64_t c = cb + threadIdx.x;
      uint32_t b = 0, e = 0;
      if (c < c1) { b = rowptr[c]; e = rowptr[c + 1]; }
      const uint32_t kmax = AGG ? __reduce_max_sync(0xFFFFFFFFu, e - b) : e - b;
      for (uint32_t j = 0; j < kmax; ++j) {
        uint32_t klo = 0xFFFFFFFFu, khi = 0xFFFFFFFFu;
        if (b + j < e) {
          const uint32_t w = pairs[b + j], t = w & 0xFFFFFFu, m = w >> 24;
          if (m & 15u) klo = t * 32 + (m & 15u);
          if (m >> 4) khi = t * 32 + 16 + (m >> 4);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t key = half ? khi : klo;
          double v = 1.0;
          bool go = key != 0xFFFFFFFFu;
          if (AGG) v = warp_agg(key, 1.0, go);
          if (go) {
            const uint32_t rel = key - t0 * 32;  // unsigned: loci below the window wrap to huge values
            if (rel < (uint32_t) WIN * 32) atomicAdd(win + rel, v);
            else red_add(B + key, v);
          }
        }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < WIN * 32; i += blockDim.x)
      if (win[i] != 0.0 && (size_t) t0 * 32 + i < (size_t) T * 32) red_add(B + (size_t) t0 * 32 + i, win[i]);
    __syncthreads();
  }
}

int main(int argc, char** argv) {
  const int64_t N = argc > 1 ? std::atoll(argv[1]) : 5000000;
  const int T = argc > 2 ? std::atoi(argv[2]) : 80000;
  std::mt19937_64 rng(20261018);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  std::poisson_distribution<int> pois(1.5);
  const double scale = std::max(T / 50.0, 1.0);
  struct Cls { uint32_t minloc; std::vector<uint32_t> words; };
  std::vector<uint32_t> minloc(N);
  std::vector<std::vector<uint32_t>> rows(N);
  int64_t P = 0, nnz = 0;
  for (int64_t c = 0; c < N; ++c) {
    const int k = std::min(1 + pois(rng), 8);
    const double lomax = std::pow(1.0 - U(rng), -1.0 / 1.2) - 1.0;
    const int64_t base = std::min<int64_t>((int64_t) (lomax * scale), T - 1);
    std::vector<uint32_t> loci{(uint32_t) base};
    for (int j = 1; j < k; ++j) loci.push_back((uint32_t) ((base + (int64_t) (U(rng) * 6)) % T));
    std::sort(loci.begin(), loci.end());
    loci.erase(std::unique(loci.begin(), loci.end()), loci.end());
    for (uint32_t t : loci) {
      const uint32_t m = U(rng) < 0.5 ? 0xFFu : 1u + (uint32_t) (U(rng) * 255);
      rows[c].push_back(t | (std::min(m, 255u) << 24));
      nnz += __builtin_popcount(std::min(m, 255u));
    }
    minloc[c] = loci[0];
    P += (int64_t) loci.size();
  }
  std::vector<int64_t> order(N);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return minloc[a] < minloc[b]; });
  std::vector<uint32_t> rowptr(N + 1, 0), pairs;
  pairs.reserve(P);
  std::vector<double> expect((size_t) T * 32, 0.0), expect_h((size_t) T * 8, 0.0);
  for (int64_t i = 0; i < N; ++i) {
    for (uint32_t w : rows[order[i]]) {
      pairs.push_back(w);
      const uint32_t t = w & 0xFFFFFFu, m = w >> 24;
      if (m & 15u) expect[(size_t) t * 32 + (m & 15u)] += 1.0;
      if (m >> 4) expect[(size_t) t * 32 + 16 + (m >> 4)] += 1.0;
      for (int h = 0; h < 8; ++h) if ((m >> h) & 1u) expect_h[(size_t) t * 8 + h] += 1.0;
    }
    rowptr[i + 1] = (uint32_t) pairs.size();
  }
  std::printf("classes %lld  loci %d  pairs %lld  nnz %lld\n", (long long) N, T, (long long) P, (long long) nnz);

  uint32_t *d_rowptr, *d_pairs;
  double* d_B;
  CK(cudaMalloc(&d_rowptr, sizeof(uint32_t) * (N + 1)));
  CK(cudaMalloc(&d_pairs, sizeof(uint32_t) * P));
  CK(cudaMalloc(&d_B, sizeof(double) * (size_t) T * 32));
  CK(cudaMemcpy(d_rowptr, rowptr.data(), sizeof(uint32_t) * (N + 1), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_pairs, pairs.data(), sizeof(uint32_t) * P, cudaMemcpyHostToDevice));
  int sms = 148;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  std::vector<double> got((size_t) T * 32);

  auto run = [&](const char* name, int64_t atomics, const std::vector<double>& want, size_t n_out, auto launch) {
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
      CK(cudaMemset(d_B, 0, sizeof(double) * (size_t) T * 32));
      CK(cudaEventRecord(e0));
      launch();
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaGetLastError());
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0) best = std::min(best, ms);
    }
    CK(cudaMemcpy(got.data(), d_B, sizeof(double) * n_out, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < n_out; ++i) bad += got[i] != want[i];
    std::printf("%-16s %9.1f us   %6.1f G atomic-adds/s issued   %s\n", name, best * 1e3,
                atomics > 0 ? atomics / (best * 1e-3) / 1e9 : 0.0, bad ? "MISMATCH" : "exact");
  };
  const int grid = sms * 8;
  run("A hap_red", nnz, expect_h, (size_t) T * 8, [&] { k_hap_red<<<grid, 256>>>(d_rowptr, d_pairs, N, d_B); });
  run("B nibble_red", 0, expect, (size_t) T * 32, [&] { k_nibble_red<<<grid, 256>>>(d_rowptr, d_pairs, N, d_B); });
  run("C nibble_red_agg", 0, expect, (size_t) T * 32, [&] { k_nibble_red_agg<<<grid, 256>>>(d_rowptr, d_pairs, N, d_B); });
  for (int cpb : {1024, 4096}) {
    char nm[64];
    std::snprintf(nm, sizeof nm, "D smem_tile/%d", cpb);
    run(nm, 0, expect, (size_t) T * 32, [&] { k_smem_tile<64, false><<<sms * 4, 256>>>(d_rowptr, d_pairs, N, cpb, T, d_B); });
    std::snprintf(nm, sizeof nm, "E smem_agg/%d", cpb);
    run(nm, 0, expect, (size_t) T * 32, [&] { k_smem_tile<64, true><<<sms * 4, 256>>>(d_rowptr, d_pairs, N, cpb, T, d_B); });
  }
  return 0;
}
