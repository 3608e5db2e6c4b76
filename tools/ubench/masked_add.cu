// Micro-benchmark: cost of "a[h] += w for the haplotypes in an 8-bit mask" per (entry, haplotype), isolated from memory.
// Variants of the inner operation of the column pass (gbrs_b200/csrc/em_kernels.cu, masked_add8).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o masked_add tools/ubench/masked_add.cu && ./masked_add
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int BIT>
__device__ __forceinline__ int select_hi(int whi, uint32_t m) {
  int r;
  asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tand.b32 t, %2, %3;\n\tsetp.ne.u32 p, t, 0;\n\tselp.b32 %0, %1, 0, p;\n\t}"
      : "=r"(r) : "r"(whi), "r"(m), "n"(BIT));
  return r;
}

template <int V>
__device__ __forceinline__ void madd(double (&a)[8], double w, uint32_t m) {
  if (V == 0) {  // 0.0 / 1.0 multiplier built from the bit (the round-1 kernel)
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      const uint32_t hi = (m & (1u << h)) * (0x3FF00000u >> h);
      a[h] = fma(w, __hiloint2double((int) hi, 0), a[h]);
    }
  } else if (V == 1) {  // select the high word only (masked-out term = denormal residue)
    const int whi = __double2hiint(w), wlo = __double2loint(w);
    a[0] += __hiloint2double(select_hi<1>(whi, m), wlo);
    a[1] += __hiloint2double(select_hi<2>(whi, m), wlo);
    a[2] += __hiloint2double(select_hi<4>(whi, m), wlo);
    a[3] += __hiloint2double(select_hi<8>(whi, m), wlo);
    a[4] += __hiloint2double(select_hi<16>(whi, m), wlo);
    a[5] += __hiloint2double(select_hi<32>(whi, m), wlo);
    a[6] += __hiloint2double(select_hi<64>(whi, m), wlo);
    a[7] += __hiloint2double(select_hi<128>(whi, m), wlo);
  } else if (V == 2) {  // select both words (exact zero), predicates from setp
    const int whi = __double2hiint(w), wlo = __double2loint(w);
    a[0] += __hiloint2double(select_hi<1>(whi, m), select_hi<1>(wlo, m));
    a[1] += __hiloint2double(select_hi<2>(whi, m), select_hi<2>(wlo, m));
    a[2] += __hiloint2double(select_hi<4>(whi, m), select_hi<4>(wlo, m));
    a[3] += __hiloint2double(select_hi<8>(whi, m), select_hi<8>(wlo, m));
    a[4] += __hiloint2double(select_hi<16>(whi, m), select_hi<16>(wlo, m));
    a[5] += __hiloint2double(select_hi<32>(whi, m), select_hi<32>(wlo, m));
    a[6] += __hiloint2double(select_hi<64>(whi, m), select_hi<64>(wlo, m));
    a[7] += __hiloint2double(select_hi<128>(whi, m), select_hi<128>(wlo, m));
  } else if (V == 3) {  // plain C: a += bit ? w : 0
#pragma unroll
    for (int h = 0; h < 8; ++h) a[h] += ((m >> h) & 1u) ? w : 0.0;
  } else if (V == 4) {  // high word AND-ed with a sign-extended bit (no predicates), low word kept
    const int whi = __double2hiint(w), wlo = __double2loint(w);
#pragma unroll
    for (int h = 0; h < 8; ++h) a[h] += __hiloint2double(whi & ((int) (m << (31 - h)) >> 31), wlo);
  } else {  // V == 5: unmasked reference (lower bound: 8 DADD)
#pragma unroll
    for (int h = 0; h < 8; ++h) a[h] += w;
  }
}

template <int V>
__global__ void __launch_bounds__(256) k(const uint32_t* __restrict__ ent, const double* __restrict__ wts, double* out, int per_thread) {
  double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t* e = ent + (size_t) (tid & 1023) * 4;  // small, cache-resident input: the loop is compute only
  for (int i = 0; i < per_thread; i += 4) {
    const uint4 v = *reinterpret_cast<const uint4*>(e + (size_t) ((i * 257) & 4095) * 1024);
    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) madd<V>(a, wts[w4[j] & 0xFFFFu], w4[j] >> 24);
  }
#pragma unroll
  for (int h = 0; h < 8; ++h) out[(size_t) tid * 8 + h] = a[h];
}

template <int V>
float run(const uint32_t* ent, const double* wts, double* out, int blocks, int per_thread, double* check) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<V><<<blocks, 256>>>(ent, wts, out, per_thread);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) k<V><<<blocks, 256>>>(ent, wts, out, per_thread);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  double h[8];
  cudaMemcpy(h, out + 8 * 12345, sizeof(h), cudaMemcpyDeviceToHost);
  *check = h[0] + h[3] + h[7];
  return ms / 5;
}

int main() {
  const int blocks = 148 * 4, per_thread = 4096;
  const size_t n_ent = (size_t) 4096 * 1024 + 4096;
  uint32_t* h_ent = new uint32_t[n_ent];
  uint64_t s = 88172645463325252ull;
  for (size_t i = 0; i < n_ent; ++i) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    const uint32_t mask = (s >> 40) & 1 ? 0xFFu : (uint32_t) ((s >> 20) % 255 + 1);
    h_ent[i] = (uint32_t) (s & 0xFFFF) | (mask << 24);
  }
  double* h_w = new double[65536];
  for (int i = 0; i < 65536; ++i) h_w[i] = 1e-3 + i * 1e-7;
  uint32_t* ent; double *wts, *out;
  cudaMalloc(&ent, n_ent * 4); cudaMalloc(&wts, 65536 * 8); cudaMalloc(&out, (size_t) blocks * 256 * 8 * 8);
  cudaMemcpy(ent, h_ent, n_ent * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(wts, h_w, 65536 * 8, cudaMemcpyHostToDevice);
  const double entries = (double) blocks * 256 * per_thread;
  double c;
  float t;
  t = run<0>(ent, wts, out, blocks, per_thread, &c); printf("V0 fma 0/1 multiplier       %8.3f ms  %6.2f ps/entry  check %.9g\n", t, t * 1e9 / entries, c);
  t = run<1>(ent, wts, out, blocks, per_thread, &c); printf("V1 select high word (R2P)   %8.3f ms  %6.2f ps/entry  check %.9g\n", t, t * 1e9 / entries, c);
  t = run<2>(ent, wts, out, blocks, per_thread, &c); printf("V2 select both words        %8.3f ms  %6.2f ps/entry  check %.9g\n", t, t * 1e9 / entries, c);
  t = run<3>(ent, wts, out, blocks, per_thread, &c); printf("V3 plain C ternary          %8.3f ms  %6.2f ps/entry  check %.9g\n", t, t * 1e9 / entries, c);
  t = run<4>(ent, wts, out, blocks, per_thread, &c); printf("V4 AND with sign-extended   %8.3f ms  %6.2f ps/entry  check %.9g\n", t, t * 1e9 / entries, c);
  t = run<5>(ent, wts, out, blocks, per_thread, &c); printf("V5 unmasked (8 DADD)        %8.3f ms  %6.2f ps/entry  check %.9g\n", t, t * 1e9 / entries, c);
  return 0;
}
