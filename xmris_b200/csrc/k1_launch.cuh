// Launch entry points of K1, one translation unit per transform length (k1_inst.cu compiled with -DXMR_N=...).
#pragma once
#include <cuda_runtime.h>

#include "k1_fft.cuh"

namespace xmr {

// win: 1 table, 2 separable.  Returns the launch status.  max_ctas <= 0: one full persistent wave.
#define XMR_DECL_K1(NN) \
    cudaError_t k1_launch_##NN(const K1Params& p, bool inverse, int win, bool tma, int max_ctas, cudaStream_t st);
XMR_DECL_K1(16) XMR_DECL_K1(32) XMR_DECL_K1(64) XMR_DECL_K1(128) XMR_DECL_K1(256) XMR_DECL_K1(512)
XMR_DECL_K1(1024) XMR_DECL_K1(2048) XMR_DECL_K1(4096) XMR_DECL_K1(8192)
#undef XMR_DECL_K1

}  // namespace xmr
