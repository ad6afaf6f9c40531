// Launch entry points of K2 (per-voxel chain), one translation unit per transform length.
#pragma once
#include <cuda_runtime.h>

#include "k2_pervoxel.cuh"

namespace xmr {
#define XMR_DECL_K2(NN) cudaError_t k2_launch_##NN(const K2Params& p, int method, cudaStream_t st);
XMR_DECL_K2(512) XMR_DECL_K2(1024) XMR_DECL_K2(2048) XMR_DECL_K2(4096) XMR_DECL_K2(8192)
#undef XMR_DECL_K2
}  // namespace xmr
