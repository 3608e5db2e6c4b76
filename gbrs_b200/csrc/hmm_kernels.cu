// `gbrs reconstruct` on the GPU: emission log-probabilities per gene and the diplotype HMM along every chromosome
// (scaled forward / backward passes, posterior, Viterbi scores and back-trace).
// reference: src/gbrs/gbrs/gbrs_utils.py:382-609 (reconstruct), :63-67 (unit_vector), :80-100 (get_genotype_probability)
// -- per chromosome a Python loop over genes with numpy operations on S x S matrices, S = H(H+1)/2 <= 36 diplotypes.
//
// A chain (one chromosome of one sample) is sequential in the gene index, so the parallelism is over chains (20
// chromosomes x the samples of a cohort) and over the S target states inside a step:
//   k_hmm_emission<H>   one thread per gene: squared distances between the unit expression vector and the unit
//                        specificity vector of every diplotype, gaussian kernel, normalisation, log
//   k_hmm_chain<H>      one 64-thread block per chain, thread k owns diplotype k.  The S x S log transition matrix of
//                        the next step is staged into (double-buffered, odd-stride) shared memory with cp.async while
//                        the current step computes, so a step costs S exp() per thread plus two block barriers; the
//                        Viterbi scores and back-pointers share the forward pass' matrix, the posterior is formed in
//                        the backward pass, the back-trace walks byte back-pointers staged through shared memory.
// HBM traffic per chain = its transition matrices twice (forward, backward): 2 * 8 * S^2 bytes per gene (20.7 KB at
// S = 36; matrices are shared by the samples of a cohort and stay L2-resident across them).
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

template <int H>
__global__ void __launch_bounds__(kChainThreads) k_hmm_chain(int32_t n_chains, const gbrs_hmm_chain* __restrict__ chains,
                                                             const double* __restrict__ init,
                                                             const double* __restrict__ eprob,
                                                             const double* __restrict__ tprob,
                                                             double* __restrict__ alpha, double* __restrict__ scaler,
                                                             double* __restrict__ gamma, double* __restrict__ delta,
                                                             uint8_t* __restrict__ backptr, int32_t* __restrict__ states) {
  constexpr int S = H * (H + 1) / 2;
  constexpr int SP = S | 1;  // odd row stride: row- and column-wise walks of the staged matrix are bank-conflict free
  __shared__ double tp[2][S * SP];
  __shared__ double a_s[kChainThreads];   // normalised forward values of the previous gene / beta of the next gene
  __shared__ double d_s[kChainThreads];   // Viterbi scores of the previous gene / emission of the next gene
  __shared__ double x_s[kChainThreads];   // exp() terms of the column being normalised
  __shared__ uint8_t bp_s[kTraceRows * S];

  const int k = threadIdx.x;
  const bool on = k < S;
  for (int chain = blockIdx.x; chain < n_chains; chain += gridDim.x) {
    const gbrs_hmm_chain ch = chains[chain];
    const int n = ch.n_genes, n_steps = ch.n_steps;
    const double* e_c = eprob + ch.gene0 * S;
    const double* t_c = tprob + ch.tprob0 * (int64_t) (S * S);
    double* al_c = alpha + ch.gene0 * S;
    double* sc_c = scaler + ch.gene0;
    double* ga_c = gamma + ch.gene0 * S;
    double* de_c = delta + ch.gene0 * S;
    uint8_t* bp_c = backptr + ch.gene0 * S;
    __syncthreads();  // shared memory of the previous chain is no longer read

    // ---------------- forward + Viterbi scores (gbrs_utils.py:498-523, :565-575) ----------------
    if (n > 1) stage_matrix<S, SP>(tp[0], t_c, k);
    {
      const double v = on ? init[k] + e_c[k] : 0.0;  // alpha[:, 0] = delta[:, 0] = init_vec + eprob[first gene]
      x_s[k] = on ? exp(v) : 0.0;
      __syncthreads();
      double sum = 0.0;
      for (int j = 0; j < S; ++j) sum += x_s[j];
      const double norm = log(sum);
      if (on) {
        a_s[k] = v - norm;
        d_s[k] = v;
        al_c[k] = v - norm;
        de_c[k] = v;
      }
      if (k == 0) sc_c[0] = -norm;
    }
    int cur = 0;
    for (int i = 1; i < n; ++i) {
      __pipeline_wait_prior(0);
      __syncthreads();  // matrix i-1 is in tp[cur]; a_s / d_s of gene i-1 are complete; x_s may be rewritten
      if (i + 1 < n) stage_matrix<S, SP>(tp[cur ^ 1], t_c + (int64_t) i * (S * S), k);
      double raw = 0.0, dnew = 0.0;
      if (on) {
        const double e = e_c[(int64_t) i * S + k];
        const double* row = tp[cur] + k * SP;  // tprob[i-1][k][:]
        double acc = 0.0, best = 0.0;
        int arg = 0;
        for (int j = 0; j < S; ++j) {
          const double t = row[j];
          acc += exp(a_s[j] + t);
          const double v = d_s[j] + t;
          if (j == 0 || v > best) { best = v; arg = j; }  // first maximum, as numpy's max / argmax
        }
        raw = log(acc + tiny()) + e;
        dnew = best + e;
        bp_c[(int64_t) (i - 1) * S + k] = (uint8_t) arg;  // = argmax_j(delta[j, i-1] + tprob[i-1][k][j]) (:589)
        x_s[k] = exp(raw);
      }
      __syncthreads();  // all reads of a_s / d_s done, x_s complete
      double sum = 0.0;
      for (int j = 0; j < S; ++j) sum += x_s[j];
      const double norm = log(sum);
      if (on) {
        a_s[k] = raw - norm;
        d_s[k] = dnew;
        al_c[(int64_t) i * S + k] = raw - norm;
        de_c[(int64_t) i * S + k] = dnew;
      }
      if (k == 0) sc_c[i] = -norm;
      cur ^= 1;
    }
    __syncthreads();  // d_s holds the scores of the last gene
    // Legacy transition files carry one matrix per gene: the reference's back-trace then starts at the last gene with
    // that extra matrix (gbrs_utils.py:585-590).
    const int n_called = n < n_steps ? n : n_steps;
    if (n_called == n && on) {
      const double* row = t_c + (int64_t) (n - 1) * (S * S) + k * S;
      double best = 0.0;
      int arg = 0;
      for (int j = 0; j < S; ++j) {
        const double v = d_s[j] + row[j];
        if (j == 0 || v > best) { best = v; arg = j; }
      }
      bp_c[(int64_t) (n - 1) * S + k] = (uint8_t) arg;
    }

    // ---------------- backward + posterior (gbrs_utils.py:527-558) ----------------
    __syncthreads();  // d_s is reused below: the extra step above has read it
    int last_state = 0;
    if (k == 0) {  // arg-max of the last gene's scores: where the back-trace starts (:581)
      double best = d_s[0];
      for (int j = 1; j < S; ++j)
        if (d_s[j] > best) { best = d_s[j]; last_state = j; }
    }
    __syncthreads();
    if (n > 1) stage_matrix<S, SP>(tp[0], t_c + (int64_t) (n - 2) * (S * S), k);
    {
      const double b = sc_c[n - 1];  // beta[:, -1] = alpha_scaler[-1]
      const double g = on ? exp(al_c[(int64_t) (n - 1) * S + k] + b) : 0.0;
      x_s[k] = g;
      if (on) {
        a_s[k] = b;
        d_s[k] = e_c[(int64_t) (n - 1) * S + k];
      }
      __syncthreads();
      double sum = 0.0;
      for (int j = 0; j < S; ++j) sum += x_s[j];
      if (on) ga_c[(int64_t) (n - 1) * S + k] = g / sum;
    }
    cur = 0;
    for (int i = n - 2; i >= 0; --i) {
      __pipeline_wait_prior(0);
      __syncthreads();  // matrix i in tp[cur]; a_s = beta of gene i+1, d_s = emission of gene i+1; x_s may be rewritten
      if (i > 0) stage_matrix<S, SP>(tp[cur ^ 1], t_c + (int64_t) (i - 1) * (S * S), k);
      double bnew = 0.0, g = 0.0, enew = 0.0;
      if (on) {
        const double sc = sc_c[i];
        const double al = al_c[(int64_t) i * S + k];
        enew = e_c[(int64_t) i * S + k];
        const double* col = tp[cur] + k;  // tprob[i][:, k]
        double acc = 0.0;
        for (int j = 0; j < S; ++j) acc += exp(((col[j * SP] + a_s[j]) + d_s[j]) + sc);
        bnew = log(acc);
        g = exp(al + bnew);
        x_s[k] = g;
      }
      __syncthreads();
      double sum = 0.0;
      for (int j = 0; j < S; ++j) sum += x_s[j];
      if (on) {
        ga_c[(int64_t) i * S + k] = g / sum;
        a_s[k] = bnew;
        d_s[k] = enew;
      }
      cur ^= 1;
    }

    // ---------------- back-trace (gbrs_utils.py:578-594) ----------------
    // states[n_called] = arg-max of the last gene, states[i] = backptr[i][states[i + 1]] for i = n_called-1 .. 0
    int32_t* st_c = states + ch.state0;
    int sid = last_state;
    if (k == 0) st_c[n_called] = sid;
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
                 const double* tprob, double* alpha, double* scaler, double* gamma, double* delta, uint8_t* backptr,
                 int32_t* states, cudaStream_t s) {
  k_hmm_chain<H><<<n_chains, kChainThreads, 0, s>>>(n_chains, chains, init, eprob, tprob, alpha, scaler, gamma, delta,
                                                    backptr, states);
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
                            const double* eprob_dev, const double* tprob_dev, double* alpha_dev, double* scaler_dev,
                            double* gamma_dev, double* delta_dev, uint8_t* backptr_dev, int32_t* states_dev,
                            void* stream) {
  if (n_chains < 0 || H < 1 || H > GBRS_HPAD) { gbrs_set_error("gbrs_hmm_run: bad argument"); return GBRS_E_ARG; }
  if (int rc = hmm_device("gbrs_hmm_run")) return rc;
  if (n_chains == 0) return GBRS_OK;
  if (!chains_dev || !init_dev || !eprob_dev || !alpha_dev || !scaler_dev || !gamma_dev || !delta_dev || !backptr_dev ||
      !states_dev) {
    gbrs_set_error("gbrs_hmm_run: null device buffer"); return GBRS_E_ARG;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  HMM_DISPATCH(H, launch_chain, n_chains, chains_dev, init_dev, eprob_dev, tprob_dev, alpha_dev, scaler_dev, gamma_dev,
                                delta_dev, backptr_dev, states_dev, s);
}
#endif  // GBRS_SIMT_EMULATION
