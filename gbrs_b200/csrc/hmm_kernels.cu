// `gbrs reconstruct` on the GPU: emission log-probabilities per gene and the diplotype HMM along every chromosome
// (scaled forward / backward passes, posterior, Viterbi scores and back-trace).
// reference: src/gbrs/gbrs/gbrs_utils.py:382-609 (reconstruct), :63-67 (unit_vector), :80-100 (get_genotype_probability)
// -- per chromosome a Python loop over genes with numpy operations on S x S matrices, S = H(H+1)/2 <= 36 diplotypes.
//
// A chain (one chromosome of one sample) is sequential in the gene index, so the parallelism is over chains (20
// chromosomes x the samples of a cohort) and over the S target states inside a step:
//   k_hmm_emission<H>   one thread per gene: squared distances between the unit expression vector and the unit
//                        specificity vector of every diplotype, gaussian kernel, normalisation, log
//   k_hmm_exp           transition matrices into the probability domain, once per matrix (they are shared by the
//                        samples of a cohort): the forward / backward recursions then are multiply-adds
//   k_hmm_chain<H>      two 64-thread blocks per chain, thread k owns diplotype k.  Block 0: scaled forward pass,
//                        backward pass and posterior on the probability-domain matrices -- per gene S multiply-adds,
//                        one log and one exp per thread and two block barriers (the reference's formulation costs S
//                        exp per state and gene).  Block 1: Viterbi scores, byte back-pointers and the back-trace on
//                        the log matrices (additions and comparisons only).  The S x S matrix of the next step is
//                        staged into double-buffered, odd-stride shared memory with cp.async while the current step
//                        computes.
// HBM traffic per chain = its matrices three times (forward, backward, Viterbi): 3 * 8 * S^2 bytes per gene (31 KB at
// S = 36); the chains of all samples on one chromosome are launched next to each other, so they share it through L2.
// Sums run in the reference's order where it is defined (python `sum` = left to right, numpy reductions over the
// first axis = row by row); exp / log are CUDA's (<= 1 ulp), so values agree with numpy to ~1e-14, not bitwise.  Given
// the same emission values the Viterbi scores and path are bit-exact (additions and comparisons only).
//
// GBRS_SIMT_EMULATION: tests/simt/ compiles the two kernels of this file unchanged with g++ against a small host shim
// (one OS thread per CUDA thread, pthread barriers, deferred cp.async) so that the CPU test-suite -- including a
// ThreadSanitizer build -- exercises the very code the GPU runs.  That build is test infrastructure only: it has no
// launchers, is not part of libgbrs_em.so and nothing in the package can reach it.
#ifndef GBRS_SIMT_EMULATION
#include <cuda_pipeline.h>
#include <cuda_runtime.h>
#endif

#include <cstdint>
#include <string>

#include "gbrs_em.h"

#ifndef GBRS_SIMT_EMULATION
void gbrs_set_error(const std::string& s);  // em_kernels.cu
#endif

namespace {

constexpr int kChainThreads = 64;   // >= 36 states; two warps
constexpr int kEmitThreads = 128;
constexpr int kTraceRows = 256;     // back-pointer rows staged per round of the back-trace

#ifndef GBRS_SIMT_EMULATION
#define HMM_CUDA(call)                                                                               \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) {                                                                         \
      gbrs_set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                            \
      return GBRS_E_CUDA;                                                                            \
    }                                                                                                \
  } while (0)
#endif

__device__ __forceinline__ double tiny() { return __longlong_as_double(1ll); }  // np.nextafter(0, 1)

// unit_vector (gbrs_utils.py:63-67): v / |v| if the plain sum of v exceeds 1e-6, else v unchanged
template <int H>
__device__ __forceinline__ void unit_vec(double (&v)[H]) {
  double s = v[0], q = v[0] * v[0];
#pragma unroll
  for (int h = 1; h < H; ++h) { s += v[h]; q += v[h] * v[h]; }
  if (s > 1e-6) {
    const double nrm = sqrt(q);
#pragma unroll
    for (int h = 0; h < H; ++h) v[h] = v[h] / nrm;
  }
}

template <int H>
__device__ __forceinline__ void load_spec_row(const double* __restrict__ spec, int i, double (&u)[H]) {
  if (spec) {
#pragma unroll
    for (int h = 0; h < H; ++h) u[h] = spec[i * H + h];
  } else {  // naive specificity: identity + 1e-4 elsewhere (gbrs_utils.py:472-475)
#pragma unroll
    for (int h = 0; h < H; ++h) u[h] = (h == i) ? 1.0 : (0.0 + 1.0 * 0.0001);
  }
  unit_vec<H>(u);
}

// Emission log-probabilities of one gene (gbrs_utils.py:476-488 with get_genotype_probability :80-100).
template <int H>
__global__ void __launch_bounds__(kEmitThreads) k_hmm_emission(int64_t n_genes, const double* __restrict__ expr,
                                                               const double* __restrict__ avec,
                                                               const int32_t* __restrict__ avec_index,
                                                               const double* __restrict__ init, double expr_threshold,
                                                               double sigma, double* __restrict__ eprob) {
  constexpr int S = H * (H + 1) / 2;
  const int64_t stride = (int64_t) gridDim.x * blockDim.x;
  for (int64_t g = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; g < n_genes; g += stride) {
    double aln[H];
    double total = 0.0;
#pragma unroll
    for (int h = 0; h < H; ++h) { aln[h] = expr[g * H + h]; total += aln[h]; }
    double* out = eprob + g * S;
    if (total < expr_threshold) {  // not expressed: the null model
#pragma unroll
      for (int k = 0; k < S; ++k) out[k] = init[k];
      continue;
    }
    const int32_t a = avec_index[g];
    const double* spec = a >= 0 ? avec + (int64_t) a * H * H : nullptr;
    const double sg = a >= 0 ? sigma : 0.450;
    const double denom = -2 * sg * sg;
    unit_vec<H>(aln);
    double psum = 0.0;
    int k = 0;
#pragma unroll
    for (int i = 0; i < H; ++i) {
      double v1[H];
      load_spec_row<H>(spec, i, v1);
#pragma unroll
      for (int j = i; j < H; ++j) {
        double gv[H];
        if (j == i) {
#pragma unroll
          for (int h = 0; h < H; ++h) gv[h] = v1[h];
        } else {
          double v2[H];
          load_spec_row<H>(spec, j, v2);
#pragma unroll
          for (int h = 0; h < H; ++h) gv[h] = v1[h] + v2[h];
          unit_vec<H>(gv);
        }
        double d = 0.0;
#pragma unroll
        for (int h = 0; h < H; ++h) { const double x = aln[h] - gv[h]; d += x * x; }
        const double p = exp(d / denom);
        out[k] = p;
        psum += p;
        ++k;
      }
    }
#pragma unroll
    for (int q = 0; q < S; ++q) out[q] = log(out[q] / psum + tiny());
  }
}

// Stage one S x S matrix (row-major, contiguous) into shared memory rows of SP doubles with 8-byte cp.async copies.
template <int S, int SP>
__device__ __forceinline__ void stage_matrix(double* __restrict__ dst, const double* __restrict__ src, int tid) {
  for (int e = tid; e < S * S; e += kChainThreads) {
    const int r = e / S, c = e - r * S;
    __pipeline_memcpy_async(dst + r * SP + c, src + e, sizeof(double));
  }
  __pipeline_commit();
}

// Transition matrices in the probability domain: out = exp(in), once per matrix, shared by every chain that walks it.
__global__ void __launch_bounds__(kEmitThreads) k_hmm_exp(int64_t n, const double* __restrict__ in,
                                                          double* __restrict__ out) {
  const int64_t stride = (int64_t) gridDim.x * blockDim.x;
  for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = exp(in[i]);
}

// One block per (chain, role).  Role 0: scaled forward pass, backward pass and posterior, on the probability-domain
// matrices (a step is a 36-term multiply-add per thread plus one exp and one log, instead of 36 exp).  Role 1: Viterbi
// scores, back-pointers and back-trace on the log matrices (additions and comparisons only: bit-exact).  The two roles
// of a chain are independent and run side by side.
template <int H>
__global__ void __launch_bounds__(kChainThreads) k_hmm_chain(int32_t n_chains, const gbrs_hmm_chain* __restrict__ chains,
                                                             const double* __restrict__ init,
                                                             const double* __restrict__ eprob,
                                                             const double* __restrict__ tprob,
                                                             const double* __restrict__ tlin,
                                                             double* __restrict__ alpha, double* __restrict__ scaler,
                                                             double* __restrict__ gamma, double* __restrict__ delta,
                                                             uint8_t* __restrict__ backptr, int32_t* __restrict__ states) {
  constexpr int S = H * (H + 1) / 2;
  constexpr int SP = S | 1;  // odd row stride: row- and column-wise walks of the staged matrix are bank-conflict free
  __shared__ double tp[2][S * SP];          // the staged matrix of the current / the next step
  __shared__ double a_s[kChainThreads];     // role 0: exp(alpha) of the previous gene / backward terms of the next gene
  __shared__ double x_s[kChainThreads];     // role 0: terms of the column being normalised
  __shared__ double d_s[2][kChainThreads];  // role 1: Viterbi scores of the previous / the current gene
  __shared__ uint8_t bp_s[kTraceRows * S];  // role 1: back-pointer rows of one back-trace round

  const int k = threadIdx.x;
  const bool on = k < S;
  for (int work = blockIdx.x; work < 2 * n_chains; work += gridDim.x) {
    const gbrs_hmm_chain ch = chains[work >> 1];
    const int n = ch.n_genes, n_steps = ch.n_steps;
    const double* e_c = eprob + ch.gene0 * S;
    __syncthreads();  // shared memory of the previous work item is no longer read

    if ((work & 1) == 0) {
      // ---------------- forward (gbrs_utils.py:498-523) ----------------
      // alpha[k, i] = log(sum_j exp(alpha[j, i-1] + tprob[i-1][k][j]) + tiny) + eprob[i][k], then normalised;
      // here sum_j P[k][j] * a[j] with P = exp(tprob) and a = exp(alpha[:, i-1]) (= the normalised terms kept from the
      // previous step).
      const double* p_c = tlin + ch.tprob0 * (int64_t) (S * S);
      double* al_c = alpha + ch.gene0 * S;
      double* sc_c = scaler + ch.gene0;
      double* ga_c = gamma + ch.gene0 * S;
      if (n > 1) stage_matrix<S, SP>(tp[0], p_c, k);
      {
        const double v = on ? init[k] + e_c[k] : 0.0;  // alpha[:, 0] = init_vec + eprob[first gene]
        const double x = on ? exp(v) : 0.0;
        x_s[k] = x;
        __syncthreads();
        double sum = 0.0;
        for (int j = 0; j < S; ++j) sum += x_s[j];
        const double norm = log(sum);
        if (on) al_c[k] = v - norm;
        if (k == 0) sc_c[0] = -norm;
        a_s[k] = x / sum;
      }
      int cur = 0;
      for (int i = 1; i < n; ++i) {
        __pipeline_wait_prior(0);
        __syncthreads();  // matrix i-1 is in tp[cur]; a_s of gene i-1 is complete; x_s may be rewritten
        if (i + 1 < n) stage_matrix<S, SP>(tp[cur ^ 1], p_c + (int64_t) i * (S * S), k);
        double raw = 0.0, x = 0.0;
        if (on) {
          const double e = e_c[(int64_t) i * S + k];
          const double* row = tp[cur] + k * SP;  // P[i-1][k][:]
          double acc = 0.0;
          for (int j = 0; j < S; ++j) acc += row[j] * a_s[j];
          raw = log(acc + tiny()) + e;
          x = exp(raw);
        }
        x_s[k] = x;
        __syncthreads();  // all reads of a_s done, x_s complete
        double sum = 0.0;
        for (int j = 0; j < S; ++j) sum += x_s[j];
        const double norm = log(sum);
        a_s[k] = x / sum;
        if (on) al_c[(int64_t) i * S + k] = raw - norm;
        if (k == 0) sc_c[i] = -norm;
        cur ^= 1;
      }

      // ---------------- backward + posterior (gbrs_utils.py:527-558) ----------------
      // beta[k, i] = log(sum_j exp(tprob[i][j][k] + beta[j, i+1] + eprob[i+1][j] + scaler[i])) = log(sum_j P[j][k] b[j])
      // with b[j] = exp((beta[j, i+1] + eprob[i+1][j]) + scaler[i]), computed by thread j as soon as it has beta[j, i+1].
      __syncthreads();  // forward reads of tp / a_s are over; scaler and alpha written by this block are visible
      if (n > 1) stage_matrix<S, SP>(tp[0], p_c + (int64_t) (n - 2) * (S * S), k);
      {
        const double b = sc_c[n - 1];  // beta[:, -1] = alpha_scaler[-1]
        const double g = on ? exp(al_c[(int64_t) (n - 1) * S + k] + b) : 0.0;
        x_s[k] = g;
        a_s[k] = (on && n > 1) ? exp((b + e_c[(int64_t) (n - 1) * S + k]) + sc_c[n - 2]) : 0.0;
        __syncthreads();
        double sum = 0.0;
        for (int j = 0; j < S; ++j) sum += x_s[j];
        if (on) ga_c[(int64_t) (n - 1) * S + k] = g / sum;
      }
      cur = 0;
      for (int i = n - 2; i >= 0; --i) {
        __pipeline_wait_prior(0);
        __syncthreads();  // matrix i in tp[cur]; a_s = b terms of gene i+1; x_s may be rewritten
        if (i > 0) stage_matrix<S, SP>(tp[cur ^ 1], p_c + (int64_t) (i - 1) * (S * S), k);
        double bnew = 0.0, g = 0.0;
        if (on) {
          const double* col = tp[cur] + k;  // P[i][:, k]
          double acc = 0.0;
          for (int j = 0; j < S; ++j) acc += col[j * SP] * a_s[j];
          bnew = log(acc);
          g = exp(al_c[(int64_t) i * S + k] + bnew);
        }
        x_s[k] = g;
        __syncthreads();  // all reads of a_s done, x_s complete
        double sum = 0.0;
        for (int j = 0; j < S; ++j) sum += x_s[j];
        if (on) {
          ga_c[(int64_t) i * S + k] = g / sum;
          if (i > 0) a_s[k] = exp((bnew + e_c[(int64_t) i * S + k]) + sc_c[i - 1]);
        }
        cur ^= 1;
      }
    } else {
      // ---------------- Viterbi scores and back-pointers (gbrs_utils.py:565-575, :589) ----------------
      const double* t_c = tprob + ch.tprob0 * (int64_t) (S * S);
      double* de_c = delta + ch.gene0 * S;
      uint8_t* bp_c = backptr + ch.gene0 * S;
      if (n > 1) stage_matrix<S, SP>(tp[0], t_c, k);
      {
        const double v = on ? init[k] + e_c[k] : 0.0;  // delta[:, 0] = init_vec + eprob[first gene]
        d_s[0][k] = v;
        if (on) de_c[k] = v;
      }
      int cur = 0, dc = 0;
      for (int i = 1; i < n; ++i) {
        __pipeline_wait_prior(0);
        __syncthreads();  // matrix i-1 is in tp[cur]; d_s[dc] of gene i-1 is complete; d_s[dc ^ 1] is free
        if (i + 1 < n) stage_matrix<S, SP>(tp[cur ^ 1], t_c + (int64_t) i * (S * S), k);
        if (on) {
          const double e = e_c[(int64_t) i * S + k];
          const double* row = tp[cur] + k * SP;  // tprob[i-1][k][:]
          double best = 0.0;
          int arg = 0;
          for (int j = 0; j < S; ++j) {
            const double v = d_s[dc][j] + row[j];
            if (j == 0 || v > best) { best = v; arg = j; }  // first maximum, as numpy's max / argmax
          }
          const double dnew = best + e;
          d_s[dc ^ 1][k] = dnew;
          de_c[(int64_t) i * S + k] = dnew;
          bp_c[(int64_t) (i - 1) * S + k] = (uint8_t) arg;  // = argmax_j(delta[j, i-1] + tprob[i-1][k][j])
        }
        cur ^= 1;
        dc ^= 1;
      }
      __syncthreads();  // d_s[dc] holds the scores of the last gene
      // Legacy transition files carry one matrix per gene: the reference's back-trace then starts at the last gene
      // with that extra matrix (gbrs_utils.py:585-590).
      const int n_called = n < n_steps ? n : n_steps;
      if (n_called == n && on) {
        const double* row = t_c + (int64_t) (n - 1) * (S * S) + k * S;
        double best = 0.0;
        int arg = 0;
        for (int j = 0; j < S; ++j) {
          const double v = d_s[dc][j] + row[j];
          if (j == 0 || v > best) { best = v; arg = j; }
        }
        bp_c[(int64_t) (n - 1) * S + k] = (uint8_t) arg;
      }
      // ---------------- back-trace (gbrs_utils.py:578-594) ----------------
      // states[n_called] = arg-max of the last gene, states[i] = backptr[i][states[i + 1]] for i = n_called-1 .. 0
      int32_t* st_c = states + ch.state0;
      int sid = 0;
      if (k == 0) {
        double best = d_s[dc][0];
        for (int j = 1; j < S; ++j)
          if (d_s[dc][j] > best) { best = d_s[dc][j]; sid = j; }
        st_c[n_called] = sid;
      }
      for (int hi = n_called; hi > 0; hi -= kTraceRows) {
        const int lo = hi > kTraceRows ? hi - kTraceRows : 0;
        __syncthreads();  // back-pointers written by this block are visible; bp_s of the previous round is done with
        for (int e = k; e < (hi - lo) * S; e += kChainThreads) bp_s[e] = bp_c[(int64_t) lo * S + e];
        __syncthreads();
        if (k == 0) {
          for (int i = hi - 1; i >= lo; --i) {
            sid = bp_s[(i - lo) * S + sid];
            st_c[i] = sid;
          }
        }
      }
    }
  }
}

#ifndef GBRS_SIMT_EMULATION
template <int H>
int launch_emission(int64_t n, const double* expr, const double* avec, const int32_t* avec_index, const double* init,
                    double thr, double sigma, double* eprob, cudaStream_t s) {
  int64_t b = (n + kEmitThreads - 1) / kEmitThreads;
  if (b > 65535) b = 65535;
  k_hmm_emission<H><<<(int) b, kEmitThreads, 0, s>>>(n, expr, avec, avec_index, init, thr, sigma, eprob);
  HMM_CUDA(cudaGetLastError());
  return GBRS_OK;
}

template <int H>
int launch_chain(int32_t n_chains, const gbrs_hmm_chain* chains, const double* init, const double* eprob,
                 const double* tprob, const double* tlin, double* alpha, double* scaler, double* gamma, double* delta,
                 uint8_t* backptr, int32_t* states, cudaStream_t s) {
  k_hmm_chain<H><<<2 * n_chains, kChainThreads, 0, s>>>(n_chains, chains, init, eprob, tprob, tlin, alpha, scaler, gamma,
                                                        delta, backptr, states);
  HMM_CUDA(cudaGetLastError());
  return GBRS_OK;
}

int hmm_device(const char* who) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    gbrs_set_error(std::string(who) + ": no CUDA device (there is no CPU fallback)");
    return GBRS_E_CUDA;
  }
  return GBRS_OK;
}

#endif  // GBRS_SIMT_EMULATION

}  // namespace

#ifndef GBRS_SIMT_EMULATION
#define HMM_DISPATCH(H, fn, ...)               \
  switch (H) {                                 \
    case 1: return fn<1>(__VA_ARGS__);         \
    case 2: return fn<2>(__VA_ARGS__);         \
    case 3: return fn<3>(__VA_ARGS__);         \
    case 4: return fn<4>(__VA_ARGS__);         \
    case 5: return fn<5>(__VA_ARGS__);         \
    case 6: return fn<6>(__VA_ARGS__);         \
    case 7: return fn<7>(__VA_ARGS__);         \
    default: return fn<8>(__VA_ARGS__);        \
  }

extern "C" int gbrs_hmm_emission(int64_t n_genes, int32_t H, const double* expr_dev, const double* avec_dev,
                                 const int32_t* avec_index_dev, const double* init_dev, double expr_threshold,
                                 double sigma, double* eprob_dev, void* stream) {
  if (n_genes < 0 || H < 1 || H > GBRS_HPAD || !(sigma > 0.0)) { gbrs_set_error("gbrs_hmm_emission: bad argument"); return GBRS_E_ARG; }
  if (int rc = hmm_device("gbrs_hmm_emission")) return rc;
  if (n_genes == 0) return GBRS_OK;
  if (!expr_dev || !avec_index_dev || !init_dev || !eprob_dev) { gbrs_set_error("gbrs_hmm_emission: null device buffer"); return GBRS_E_ARG; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  HMM_DISPATCH(H, launch_emission, n_genes, expr_dev, avec_dev, avec_index_dev, init_dev, expr_threshold, sigma, eprob_dev, s);
}

extern "C" int gbrs_hmm_run(int32_t n_chains, const gbrs_hmm_chain* chains_dev, int32_t H, const double* init_dev,
                            const double* eprob_dev, const double* tprob_dev, int64_t n_matrices, double* tprob_lin_dev,
                            double* alpha_dev, double* scaler_dev, double* gamma_dev, double* delta_dev,
                            uint8_t* backptr_dev, int32_t* states_dev, void* stream) {
  if (n_chains < 0 || n_matrices < 0 || H < 1 || H > GBRS_HPAD) { gbrs_set_error("gbrs_hmm_run: bad argument"); return GBRS_E_ARG; }
  if (int rc = hmm_device("gbrs_hmm_run")) return rc;
  if (n_chains == 0) return GBRS_OK;
  if (!chains_dev || !init_dev || !eprob_dev || !alpha_dev || !scaler_dev || !gamma_dev || !delta_dev || !backptr_dev ||
      !states_dev || (n_matrices > 0 && (!tprob_dev || !tprob_lin_dev))) {
    gbrs_set_error("gbrs_hmm_run: null device buffer"); return GBRS_E_ARG;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t n_elem = n_matrices * (int64_t) (H * (H + 1) / 2) * (H * (H + 1) / 2);
  if (n_elem > 0) {
    int64_t b = (n_elem + kEmitThreads - 1) / kEmitThreads;
    if (b > 148 * 16) b = 148 * 16;
    k_hmm_exp<<<(int) b, kEmitThreads, 0, s>>>(n_elem, tprob_dev, tprob_lin_dev);
    HMM_CUDA(cudaGetLastError());
  }
  HMM_DISPATCH(H, launch_chain, n_chains, chains_dev, init_dev, eprob_dev, tprob_dev, tprob_lin_dev, alpha_dev, scaler_dev,
               gamma_dev, delta_dev, backptr_dev, states_dev, s);
}
#endif  // GBRS_SIMT_EMULATION
