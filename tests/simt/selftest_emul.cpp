// TEST INFRASTRUCTURE ONLY -- selftest_kernels.cu on the host SIMT shim (see tests/test_simt_shim.py).
#include "simt_shim.h"

#define GBRS_SIMT_EMULATION 1
#include "selftest_kernels.cu"

extern "C" void emul_block_sum(const double* x, int64_t n, double* total, double* per_block, int grid, int block) {
  simt_launch(grid, block, [=] { k_block_sum(x, n, total, per_block); });
}
extern "C" void emul_compact_positive(const int32_t* x, int32_t n, int32_t* out, uint32_t* counter, int grid, int block) {
  simt_launch(grid, block, [=] { k_compact_positive(x, n, out, counter); });
}
extern "C" void emul_group_scan(const uint32_t* x, uint32_t* scan8, uint32_t* warp_max, int32_t* votes, int grid, int block) {
  simt_launch(grid, block, [=] { k_group_scan(x, scan8, warp_max, votes); });
}
extern "C" void emul_histogram(const double* xy, int32_t n, double* hist, unsigned long long* checksum, int grid, int block) {
  simt_launch(grid, block, [=] { k_histogram(reinterpret_cast<const double2*>(xy), n, hist, checksum); });
}
