// Microbenchmark (2 GPUs of one NVSwitch domain, one process): how long do SM-issued stores of a few MB into a PEER's memory
// take, by store flavour?  The exchange step of the row-sharded EM update pushes 64 T bytes per rank per update this way.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o p2p_store p2p_store.cu && ./p2p_store
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void k_st_sys8(double* dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(dst + i), "d"((double) i) : "memory");
}
__global__ void k_st_sys16(double* dst, size_t n) {
  for (size_t i = 2 * (blockIdx.x * (size_t) blockDim.x + threadIdx.x); i < n; i += 2 * (size_t) gridDim.x * blockDim.x)
    asm volatile("st.relaxed.sys.global.v2.f64 [%0], {%1, %2};" ::"l"(dst + i), "d"((double) i), "d"(1.0) : "memory");
}
__global__ void k_st_weak16(double* dst, size_t n) {
  for (size_t i = 2 * (blockIdx.x * (size_t) blockDim.x + threadIdx.x); i < n; i += 2 * (size_t) gridDim.x * blockDim.x)
    *reinterpret_cast<double2*>(dst + i) = make_double2((double) i, 1.0);
}
// 64-byte rows at scattered positions (the numerator push of phase A: 8 lanes = one locus row, loci in a permuted order)
__global__ void k_st_sys8_rows(double* dst, size_t n) {
  const size_t rows = n / 8;
  for (size_t g = (blockIdx.x * (size_t) blockDim.x + threadIdx.x) / 8; g < rows; g += (size_t) gridDim.x * blockDim.x / 8) {
    const size_t r = (g * 2654435761ull) % rows;  // permuted row
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(dst + r * 8 + (threadIdx.x & 7)), "d"((double) g) : "memory");
  }
}
// bulk copy engine path: stage CHUNK bytes in shared memory, one thread issues cp.async.bulk shared -> (peer) global
template <int CHUNK>
__global__ void k_bulk(double* dst, size_t n) {
  extern __shared__ __align__(128) double sm[];
  const size_t per = CHUNK / 8, chunks = n / per;
  for (size_t c = blockIdx.x; c < chunks; c += gridDim.x) {
    for (int i = threadIdx.x; i < (int) per; i += blockDim.x) sm[i] = (double) (c * per + i);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c * per),
                   "r"((unsigned) __cvta_generic_to_shared(sm)), "r"(CHUNK) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the staging buffer may be refilled
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <class F>
float time_it(F launch, int reps = 20) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  float best = 1e9f, tot = 0.f;
  for (int i = 0; i < reps; ++i) {
    CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    best = ms < best ? ms : best; tot += ms;
  }
  printf("   best %7.2f us  mean %7.2f us", best * 1e3f, tot / reps * 1e3f);
  return best;
}

int main() {
  int nd = 0; CK(cudaGetDeviceCount(&nd));
  if (nd < 2) { printf("needs 2 GPUs\n"); return 0; }
  CK(cudaSetDevice(0));
  int can = 0; CK(cudaDeviceCanAccessPeer(&can, 0, 1));
  printf("peer access 0->1: %d\n", can);
  CK(cudaDeviceEnablePeerAccess(1, 0));
  double *local, *peer;
  const size_t max_n = (size_t) 8 << 20;
  CK(cudaMalloc(&local, max_n * 8));
  CK(cudaSetDevice(1)); CK(cudaMalloc(&peer, max_n * 8)); CK(cudaSetDevice(0));
  CK(cudaFuncSetAttribute(k_bulk<8192>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192));
  CK(cudaFuncSetAttribute(k_bulk<32768>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  const int grid = 148 * 4;
  for (size_t mb10 : {6, 26, 51, 205}) {  // 0.64 MB (1/8 slice), 2.56 MB (1/2), 5.12 MB (whole numerator), 20.5 MB
    const size_t n = mb10 * 100000 / 8 / 4096 * 4096;
    for (int where = 0; where < 2; ++where) {
      double* dst = where ? peer : local;
      printf("%5.2f MB -> %s\n", n * 8 / 1e6, where ? "PEER" : "local");
      auto rep = [&](const char* name, float ms) { printf("  %-28s %8.1f GB/s\n", name, n * 8 / (ms * 1e-3) / 1e9); };
      rep("st.relaxed.sys 8B/lane", time_it([&] { k_st_sys8<<<grid, 256>>>(dst, n); }));
      rep("st.relaxed.sys 16B/lane", time_it([&] { k_st_sys16<<<grid, 256>>>(dst, n); }));
      rep("st (weak) 16B/lane", time_it([&] { k_st_weak16<<<grid, 256>>>(dst, n); }));
      rep("st.relaxed.sys 64B rows perm", time_it([&] { k_st_sys8_rows<<<grid, 256>>>(dst, n); }));
      rep("cp.async.bulk 8 KB chunks", time_it([&] { k_bulk<8192><<<grid, 128, 8192>>>(dst, n); }));
      rep("cp.async.bulk 32 KB chunks", time_it([&] { k_bulk<32768><<<148, 256, 32768>>>(dst, n); }));
      rep("empty-ish (1 block)", time_it([&] { k_st_sys8<<<1, 32>>>(dst, 32); }));
    }
  }
  return 0;
}
