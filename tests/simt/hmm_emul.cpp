// TEST INFRASTRUCTURE ONLY -- gbrs_b200/csrc/hmm_kernels.cu compiled for the host SIMT shim (see simt_shim.h).  Exposes
// the two kernels with the argument lists of gbrs_hmm_emission / gbrs_hmm_run (host pointers, no stream) so that
// tests/test_reconstruct_simt.py can compare the kernel code itself with the oracle and the reference's golden vectors
// on a machine without a GPU.  Never loaded by the package.
#include "simt_shim.h"

#define GBRS_SIMT_EMULATION 1
#include "../../gbrs_b200/csrc/hmm_kernels.cu"

#define EMUL_DISPATCH(H, body)                      \
  switch (H) {                                      \
    case 1: { constexpr int HH = 1; body; } break;  \
    case 2: { constexpr int HH = 2; body; } break;  \
    case 3: { constexpr int HH = 3; body; } break;  \
    case 4: { constexpr int HH = 4; body; } break;  \
    case 5: { constexpr int HH = 5; body; } break;  \
    case 6: { constexpr int HH = 6; body; } break;  \
    case 7: { constexpr int HH = 7; body; } break;  \
    default: { constexpr int HH = 8; body; } break; \
  }

extern "C" int emul_hmm_emission(int64_t n_genes, int32_t H, const double* expr, const double* avec,
                                 const int32_t* avec_index, const double* init, double expr_threshold, double sigma,
                                 double* eprob, int32_t grid) {
  EMUL_DISPATCH(H, simt_launch(grid, kEmitThreads, [=] {
                  k_hmm_emission<HH>(n_genes, expr, avec, avec_index, init, expr_threshold, sigma, eprob);
                }));
  return 0;
}

extern "C" int emul_hmm_run(int32_t n_chains, const gbrs_hmm_chain* chains, int32_t H, const double* init,
                            const double* eprob, const double* tprob, int64_t n_matrices, double* tlin, double* alpha,
                            double* scaler, double* gamma, double* delta, uint8_t* backptr, int32_t* states,
                            int32_t grid) {
  const int64_t n_elem = n_matrices * (int64_t) (H * (H + 1) / 2) * (H * (H + 1) / 2);
  if (n_elem > 0) simt_launch(3, kEmitThreads, [=] { k_hmm_exp(n_elem, tprob, tlin); });
  EMUL_DISPATCH(H, simt_launch(grid, kChainThreads, [=] {
                  k_hmm_chain<HH>(n_chains, chains, init, eprob, tprob, tlin, alpha, scaler, gamma, delta, backptr,
                                  states);
                }));
  return 0;
}
