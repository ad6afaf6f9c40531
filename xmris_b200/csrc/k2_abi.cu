// C ABI: the per-voxel chain (autophase mode="all"), see include/xmris_b200.h.
#include <cstring>

#include "../../include/xmris_b200.h"
#include "abi_common.h"
#include "k2_launch.cuh"

namespace xmr_abi {
int get_twiddles(int n, const float2** out);   // xmris_abi.cu
}

extern "C" int xmr_chain_each_c64(const void* in_dev, void* out_dev, int64_t batch, int n_in, int n_out, int pad_left,
                                  int input_is_spectrum, int window_mode, const float* window_dev,
                                  const float* win_rows_host, float scale, int method, double du, int fixed_pivot,
                                  double u0_fixed, int fixed_target, int index_width, int p0_only, double* p0_dev,
                                  double* p1_dev, int* pivot_dev, float* fun_dev, void* stream) {
    using namespace xmr;
    if (!(n_out == 512 || n_out == 1024 || n_out == 2048 || n_out == 4096 || n_out == 8192))
        return xmr_abi::fail(XMR_ERR_UNSUPPORTED_N,
                             "per-spectrum autophase: n_out=%d must be a power of two in [512, 8192]", n_out);
    if (batch < 0 || n_in < 1 || n_in > n_out || pad_left < 0 || pad_left + n_in > n_out)
        return xmr_abi::fail(XMR_ERR_BAD_ARG, "bad sizes: batch=%lld n_in=%d n_out=%d pad_left=%d", (long long)batch, n_in,
                             n_out, pad_left);
    if (batch == 0) return XMR_OK;
    if (!in_dev || !out_dev || !p0_dev || !p1_dev || !pivot_dev || !fun_dev)
        return xmr_abi::fail(XMR_ERR_BAD_ARG, "NULL pointer");
    if (input_is_spectrum && (n_in != n_out || pad_left != 0))
        return xmr_abi::fail(XMR_ERR_BAD_ARG, "spectrum input needs n_in == n_out and pad_left == 0");
    if (method < XMR_METHOD_ACME || method > XMR_METHOD_POSITIVITY) return xmr_abi::fail(XMR_ERR_BAD_ARG, "method=%d", method);
    if (index_width < 1) return xmr_abi::fail(XMR_ERR_BAD_ARG, "index_width=%d", index_width);
    if (fixed_pivot && (fixed_target < 0 || fixed_target >= n_out))
        return xmr_abi::fail(XMR_ERR_BAD_ARG, "fixed_target=%d", fixed_target);
    if (window_mode < XMR_WIN_NONE || window_mode > XMR_WIN_SEPARABLE) return xmr_abi::fail(XMR_ERR_BAD_ARG, "window_mode");
    if (window_mode != XMR_WIN_NONE && !window_dev) return xmr_abi::fail(XMR_ERR_BAD_ARG, "window_dev is NULL");
    if (window_mode == XMR_WIN_SEPARABLE && !win_rows_host) return xmr_abi::fail(XMR_ERR_BAD_ARG, "win_rows_host is NULL");

    K2Params p;
    std::memset(&p, 0, sizeof(p));
    int rc = xmr_abi::get_twiddles(n_out, &p.twN);
    if (rc != XMR_OK) return rc;
    p.in = static_cast<const float2*>(in_dev);
    p.out = static_cast<float2*>(out_dev);
    p.batch = batch;
    p.n_in = n_in;
    p.pad_left = pad_left;
    p.out_shift = n_out / 2;
    p.use_tma = ((reinterpret_cast<uintptr_t>(in_dev) & 15) == 0) && ((n_in & 1) == 0);
    p.scale = scale;
    p.spec_in = input_is_spectrum ? 1 : 0;
    for (int i = 0; i < 32; ++i) p.win_rows[i] = 1.0f;
    const int r0 = n_out >= 256 ? n_out / 256 : 1;
    if (window_mode == XMR_WIN_TABLE) {
        p.win = window_dev;
        p.win_table = 1;
    } else if (window_mode == XMR_WIN_SEPARABLE) {
        p.win = window_dev;
        for (int i = 0; i < r0; ++i) p.win_rows[i] = win_rows_host[i];
    }
    p.du = du;
    p.fixed_pivot = fixed_pivot ? 1 : 0;
    p.u0_fixed = u0_fixed;
    p.fixed_target = fixed_target;
    p.index_width = index_width;
    p.p0_only = p0_only ? 1 : 0;
    p.p0_out = p0_dev;
    p.p1_out = p1_dev;
    p.pivot_out = pivot_dev;
    p.fun_out = fun_dev;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaSuccess;
    switch (n_out) {
        case 512: e = k2_launch_512(p, method, st); break;
        case 1024: e = k2_launch_1024(p, method, st); break;
        case 2048: e = k2_launch_2048(p, method, st); break;
        case 4096: e = k2_launch_4096(p, method, st); break;
        default: e = k2_launch_8192(p, method, st); break;
    }
    if (e != cudaSuccess) return xmr_abi::cuda_fail(e, "k2 launch");
    return XMR_OK;
}
