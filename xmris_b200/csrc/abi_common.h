// Shared error plumbing of the C ABI translation units (thread-local message behind xmr_last_error()).
#pragma once
#include <cuda_runtime.h>

namespace xmr_abi {
int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
}  // namespace xmr_abi
