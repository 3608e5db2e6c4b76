// TEST INFRASTRUCTURE ONLY -- the three symbols the host-only sources (pack.cpp, report.cpp) expect from em_kernels.cu,
// so that they can be built without nvcc into an AddressSanitizer / UBSan library (tests/test_host_sanitizers.py).
#include <string>

#include "gbrs_em.h"

static thread_local std::string g_err;
void gbrs_set_error(const std::string& s) { g_err = s; }
extern "C" const char* gbrs_last_error(void) { return g_err.c_str(); }
extern "C" int gbrs_abi_version(void) { return GBRS_EM_ABI_VERSION; }
