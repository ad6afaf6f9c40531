// C ABI of libxmris_b200.so (see include/xmris_b200.h) + the small elementwise / reduction kernels.
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "../../include/xmris_b200.h"
#include "abi_common.h"
#include "k1_launch.cuh"

namespace xmr_abi {
thread_local char g_err[512] = "";
int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    return fail(XMR_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}
}  // namespace xmr_abi

namespace {
using xmr_abi::cuda_fail;
using xmr_abi::fail;

bool supported_n(int n) { return n >= 16 && n <= 8192 && (n & (n - 1)) == 0; }

// ---- per-(device, N) twiddle tables: exp(-2 pi i k / N) rounded from float64 -----------------------------
}  // namespace

namespace xmr_abi {
std::mutex g_tw_mutex;
std::map<std::pair<int, int>, float2*> g_tw;

int get_twiddles(int n, const float2** out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    std::lock_guard<std::mutex> lock(g_tw_mutex);
    auto key = std::make_pair(dev, n);
    auto it = g_tw.find(key);
    if (it != g_tw.end()) {
        *out = it->second;
        return XMR_OK;
    }
    std::vector<float2> h(n);
    for (int k = 0; k < n; ++k) {
        // exact octant symmetries keep the table bit-symmetric; plain cos/sin in double is enough for float32
        const double a = -2.0 * M_PI * double(k) / double(n);
        h[k] = make_float2(float(std::cos(a)), float(std::sin(a)));
    }
    float2* d = nullptr;
    e = cudaMalloc(&d, sizeof(float2) * n);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(twiddles)");
    e = cudaMemcpy(d, h.data(), sizeof(float2) * n, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(d);
        return cuda_fail(e, "cudaMemcpy(twiddles)");
    }
    g_tw[key] = d;
    *out = d;
    return XMR_OK;
}
}  // namespace xmr_abi

namespace {
using xmr_abi::get_twiddles;

// ---- elementwise kernels --------------------------------------------------------------------------------
__global__ void zero_fill_kernel(const float2* __restrict__ in, float2* __restrict__ out, long long batch, int n_in,
                                 int n_out, int pad_left) {
    const long long total = batch * n_out;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / n_out;
        const int k = int(i - b * n_out) - pad_left;
        out[i] = (k >= 0 && k < n_in) ? in[b * n_in + k] : make_float2(0.f, 0.f);
    }
}
// out[b, (k + shift) mod n] = in[b, k]  (np.roll along the last axis; fftshift / ifftshift, fourier.py:10-58)
__global__ void roll_rows_kernel(const float2* __restrict__ in, float2* __restrict__ out, long long batch, int n, int shift) {
    const long long total = batch * n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / n;
        int k = int(i - b * n) - shift;       // destination index i <- source index k
        if (k < 0) k += n;
        out[i] = in[b * n + k];
    }
}
__global__ void scale_rows_kernel(const float2* __restrict__ in, float2* __restrict__ out, long long batch, int n,
                                  const float* __restrict__ w) {
    const long long total = batch * n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const float s = w[i % n];
        const float2 x = in[i];
        out[i] = make_float2(x.x * s, x.y * s);
    }
}
__global__ void rotate_rows_kernel(const float2* __restrict__ in, float2* __restrict__ out, long long batch, int n,
                                   const float2* __restrict__ rot) {
    const long long total = batch * n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        out[i] = xmr::cmul(in[i], rot[i % n]);
    }
}
// out[b, (j + out_shift) mod n] = in[b * in_stride + (j + in_shift) mod n] * rot[j]: the pre- and post-chirp of the chirp-z
// transform with the input rotation / fftshift folded in and the input rows taken out of longer ones
__global__ void rotate_rows_shift_kernel(const float2* __restrict__ in, long long in_stride, float2* __restrict__ out,
                                         long long batch, int n, const float2* __restrict__ rot, int in_shift, int out_shift) {
    const long long total = batch * n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / n;
        const int j = int(i - b * n);
        int js = j + in_shift, jd = j + out_shift;
        js -= js >= n ? n : 0;
        jd -= jd >= n ? n : 0;
        out[b * n + jd] = xmr::cmul(in[b * in_stride + js], rot[j]);
    }
}
// one block per spectrum row chunk; phase in turns reduced in double once per (row, 64-point anchor)
__global__ void phase_each_kernel(const float2* __restrict__ in, float2* __restrict__ out, long long batch, int n,
                                  const double* __restrict__ a_turns, const double* __restrict__ b_turns) {
    for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
        const double a = a_turns[b], s = b_turns[b];
        for (int m = threadIdx.x; m < n; m += blockDim.x) {
            double turns = a + s * double(m);
            turns -= floor(turns);
            float sn, cs;
            sincospif(2.0f * float(turns), &sn, &cs);
            out[b * n + m] = xmr::cmul(in[b * n + m], make_float2(cs, sn));
        }
    }
}

// Two-level first-occurrence argmax: stage 1 reduces strided slices to per-block candidates, stage 2 (one block)
// merges them.  Ties resolve to the lowest row index (numpy argmax order, phasing.py:229-231).
__device__ __forceinline__ void argmax_merge(float& v, long long& i, float ov, long long oi) {
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}
__device__ __forceinline__ void argmax_block(float& best, long long& besti) {
    __shared__ float sv[32];
    __shared__ long long si[32];
    for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, off);
        const long long oi = __shfl_xor_sync(0xffffffffu, besti, off);
        argmax_merge(best, besti, ov, oi);
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = besti; }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int w = 1; w < int(blockDim.x >> 5); ++w) argmax_merge(best, besti, sv[w], si[w]);
}
constexpr int ARGMAX_BLOCKS = 256;
__device__ float g_part_v[8][ARGMAX_BLOCKS];          // indexed by a small stream-slot to keep concurrent calls apart
__device__ long long g_part_i[8][ARGMAX_BLOCKS];

__global__ void global_argmax_stage1(const float* __restrict__ absmax, long long batch, int slot) {
    float best = -1.f;
    long long besti = 0x7fffffffffffffffLL;
    for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < batch; b += (long long)gridDim.x * blockDim.x)
        argmax_merge(best, besti, absmax[b], b);
    argmax_block(best, besti);
    if (threadIdx.x == 0) { g_part_v[slot][blockIdx.x] = best; g_part_i[slot][blockIdx.x] = besti; }
}
__global__ void global_argmax_kernel(const float* __restrict__ absmax, const int* __restrict__ argmax, long long batch,
                                     int n, unsigned char* out, int nparts, int slot) {
    float best = -1.f;
    long long besti = 0x7fffffffffffffffLL;
    for (int b = threadIdx.x; b < nparts; b += blockDim.x) argmax_merge(best, besti, g_part_v[slot][b], g_part_i[slot][b]);
    argmax_block(best, besti);
    if (threadIdx.x == 0) {
        *reinterpret_cast<float*>(out) = best;
        *reinterpret_cast<long long*>(out + 8) = (batch > 0) ? besti * n + (argmax ? argmax[besti] : 0) : -1;
    }
}

// ---- multi-GPU exchange of mode="single" without the host (sharding.py): every rank packs its best row + record, ONE
// all-gather moves them, every rank selects the global winner on the device and searches it redundantly ----------------------
// slot layout (complex64 elements, slot_elems = even(n_in) + 2): [0, n_in) the FID row | element slot_elems-2: {float max |S|, 0}
// | element slot_elems-1: int64 global row
__global__ void pack_winner_kernel(const float2* __restrict__ fid, long long batch, int n_in, int n_out, int slot_elems,
                                   const unsigned char* __restrict__ arg_record, long long row_offset, float2* __restrict__ slot) {
    const float vmax = batch > 0 ? *reinterpret_cast<const float*>(arg_record) : -1.f;
    const long long row = batch > 0 ? *reinterpret_cast<const long long*>(arg_record + 8) / n_out : 0;
    for (int k = threadIdx.x; k < n_in; k += blockDim.x) slot[k] = batch > 0 ? fid[row * n_in + k] : make_float2(0.f, 0.f);
    if (threadIdx.x == 0) {
        slot[slot_elems - 2] = make_float2(vmax, 0.f);
        *reinterpret_cast<long long*>(slot + slot_elems - 1) = batch > 0 ? row + row_offset : 0x7fffffffffffffffLL;
    }
}
// ctl: {float max @0, int64 winning slot @8, int64 global row @16}; ties go to the lowest global row (numpy argmax order)
__global__ void select_winner_kernel(const float2* __restrict__ gathered, int world, int slot_elems, unsigned char* ctl) {
    if (threadIdx.x == 0) {
        float best = -2.f;
        long long brow = 0x7fffffffffffffffLL, bslot = 0;
        for (int r = 0; r < world; ++r) {
            const float2* s = gathered + (long long)r * slot_elems;
            const float v = s[slot_elems - 2].x;
            const long long row = *reinterpret_cast<const long long*>(s + slot_elems - 1);
            if (v > best || (v == best && row < brow)) { best = v; brow = row; bslot = r; }
        }
        *reinterpret_cast<float*>(ctl) = best;
        *reinterpret_cast<long long*>(ctl + 8) = bslot;
        *reinterpret_cast<long long*>(ctl + 16) = brow;
    }
}

int grid_for(long long total, int block) {
    long long g = (total + block - 1) / block;
    const long long cap = 148LL * 16;
    return int(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" {

int xmr_version(void) { return 100; }
const char* xmr_last_error(void) { return xmr_abi::g_err; }

}  // extern "C"

namespace xmr_abi {
// row_slot_dev (optional): transform the row at fid_dev + (*row_slot_dev) * row_stride elements (batch must be 1);
// ph_dev (optional, phase_mode = XMR_PHASE_UNIFORM): xmr::K1PhaseDev in device memory instead of ph_a_turns / ph_b_turns
int k1_dispatch(const void* fid_dev, void* spec_dev, int64_t batch, int n_in, int n_out, int pad_left,
                int window_mode, const float* window_dev, const float* win_rows_host, float scale,
                int inverse, int in_shift, int out_shift, float* absmax_dev, int* argmax_dev,
                int phase_mode, double ph_a_turns, double ph_b_turns, float* run_max2, const long long* row_slot_dev,
                int row_stride, const void* ph_dev, void* stream) {
    if (!supported_n(n_out))
        return fail(XMR_ERR_UNSUPPORTED_N, "n_out=%d: transform length must be a power of two in [16, 8192]", n_out);
    if (batch < 0 || n_in < 1 || n_in > n_out || pad_left < 0 || pad_left + n_in > n_out)
        return fail(XMR_ERR_BAD_ARG, "bad sizes: batch=%lld n_in=%d n_out=%d pad_left=%d", (long long)batch, n_in, n_out,
                    pad_left);
    if (batch == 0) return XMR_OK;
    if (!fid_dev) return fail(XMR_ERR_BAD_ARG, "fid_dev is NULL");
    if (!spec_dev && !absmax_dev) return fail(XMR_ERR_BAD_ARG, "nothing to do: spec_dev and absmax_dev are both NULL");
    if (argmax_dev != nullptr && absmax_dev == nullptr)
        return fail(XMR_ERR_BAD_ARG, "argmax_dev needs absmax_dev");
    if (in_shift != 0 && (n_in != n_out || pad_left != 0))
        return fail(XMR_ERR_BAD_ARG, "in_shift requires n_in == n_out and pad_left == 0");
    if (in_shift < 0 || in_shift >= n_out || out_shift < 0 || out_shift >= n_out)
        return fail(XMR_ERR_BAD_ARG, "shifts must lie in [0, n_out)");
    if (window_mode < XMR_WIN_NONE || window_mode > XMR_WIN_SEPARABLE)
        return fail(XMR_ERR_BAD_ARG, "window_mode=%d", window_mode);
    if (window_mode != XMR_WIN_NONE && !window_dev) return fail(XMR_ERR_BAD_ARG, "window_dev is NULL");
    if (window_mode == XMR_WIN_SEPARABLE && !win_rows_host) return fail(XMR_ERR_BAD_ARG, "win_rows_host is NULL");
    if (inverse && window_mode != XMR_WIN_NONE) return fail(XMR_ERR_BAD_ARG, "inverse transform takes no window");
    if (phase_mode != XMR_PHASE_NONE && phase_mode != XMR_PHASE_UNIFORM)
        return fail(XMR_ERR_BAD_ARG, "phase_mode=%d", phase_mode);
    const int q = n_out / 16;
    if (phase_mode == XMR_PHASE_UNIFORM && (out_shift % q) != 0)
        return fail(XMR_ERR_BAD_ARG, "fused phase needs out_shift to be a multiple of n_out/16");
    if ((reinterpret_cast<uintptr_t>(fid_dev) & 7) || (reinterpret_cast<uintptr_t>(spec_dev) & 7))
        return fail(XMR_ERR_BAD_ARG, "complex64 rows must be 8-byte aligned");

    xmr::K1Params p;
    std::memset(&p, 0, sizeof(p));
    int rc = get_twiddles(n_out, &p.twN);
    if (rc != XMR_OK) return rc;
    p.in = static_cast<const float2*>(fid_dev);
    p.out = static_cast<float2*>(spec_dev);
    p.batch = batch;
    p.n_in = n_in;
    p.pad_left = pad_left;
    p.in_shift = in_shift;
    p.out_shift = out_shift;
    p.scale = scale;
    p.absmax = absmax_dev;
    p.argmax = argmax_dev;
    p.run_max2 = run_max2;
    p.row_slot = row_slot_dev;
    p.row_stride = row_stride;
    p.ph_dev = static_cast<const xmr::K1PhaseDev*>(ph_dev);
    if (row_slot_dev != nullptr && (batch != 1 || (row_stride & 1))) return fail(XMR_ERR_BAD_ARG, "row_slot_dev needs batch == 1 and an even stride");
    const int r0 = n_out >= 256 ? n_out / 256 : 1;
    for (int i = 0; i < 32; ++i) p.win_rows[i] = 1.0f;
    int win = 2;
    if (window_mode == XMR_WIN_TABLE) {
        win = 1;
        p.win = window_dev;
    } else if (window_mode == XMR_WIN_SEPARABLE) {
        p.win = window_dev;
        for (int i = 0; i < r0; ++i) p.win_rows[i] = win_rows_host[i];
    } else {
        p.win = nullptr;   // scale only
    }
    p.phase_on = (phase_mode == XMR_PHASE_UNIFORM);
    p.ph_a_turns = ph_a_turns;
    p.ph_b_turns = ph_b_turns;
    for (int d = 0; d < 16; ++d) {
        double turns = ph_b_turns * double(q) * double(d);
        turns -= std::floor(turns);
        p.ph_step[d] = make_float2(float(std::cos(2.0 * M_PI * turns)), float(std::sin(2.0 * M_PI * turns)));
        // folded-phase fast variants (out_shift = n_out/2): the whole rotation of stored index m0(d) = (q*d + N/2) mod N
        const long long m0 = (static_cast<long long>(q) * d + n_out / 2) % n_out;
        double tf = ph_a_turns + ph_b_turns * double(m0);
        tf -= std::floor(tf);
        p.ph_fold[d] = make_float2(float(std::cos(2.0 * M_PI * tf)), float(std::sin(2.0 * M_PI * tf)));
    }
    // TMA bulk copies need 16-byte aligned rows
    const bool tma = ((reinterpret_cast<uintptr_t>(fid_dev) & 15) == 0) && ((n_in & 1) == 0);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaSuccess;
    switch (n_out) {
#define XMR_CASE(NN) case NN: e = xmr::k1_launch_##NN(p, inverse != 0, win, tma, 0, st); break;
        XMR_CASE(16) XMR_CASE(32) XMR_CASE(64) XMR_CASE(128) XMR_CASE(256) XMR_CASE(512) XMR_CASE(1024)
        XMR_CASE(2048) XMR_CASE(4096) XMR_CASE(8192)
#undef XMR_CASE
        default: return fail(XMR_ERR_UNSUPPORTED_N, "n_out=%d", n_out);
    }
    if (e != cudaSuccess) return cuda_fail(e, "k1 launch");
    return XMR_OK;
}
}  // namespace xmr_abi

namespace xmr_abi {
int pack_winner(const void* fid_dev, int64_t batch, int n_in, int n_out, const void* arg_record_dev, int64_t row_offset,
                void* slot_dev, void* stream) {
    pack_winner_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const float2*>(fid_dev), batch, n_in, n_out,
                                                                        ((n_in + 1) & ~1) + 2,
                                                                        static_cast<const unsigned char*>(arg_record_dev), row_offset,
                                                                        static_cast<float2*>(slot_dev));
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? XMR_OK : cuda_fail(e, "pack_winner launch");
}
int select_winner(const void* gathered_dev, int world, int n_in, void* ctl_dev, void* stream) {
    select_winner_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const float2*>(gathered_dev), world, ((n_in + 1) & ~1) + 2,
                                                                         static_cast<unsigned char*>(ctl_dev));
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? XMR_OK : cuda_fail(e, "select_winner launch");
}
}  // namespace xmr_abi

extern "C" {
using xmr_abi::k1_dispatch;

int xmr_fid_to_spectrum_c64(const void* fid_dev, void* spec_dev, int64_t batch, int n_in, int n_out, int pad_left,
                            int window_mode, const float* window_dev, const float* win_rows_host, float scale,
                            int inverse, int in_shift, int out_shift, float* absmax_dev, int* argmax_dev,
                            int phase_mode, double ph_a_turns, double ph_b_turns, void* stream) {
    return k1_dispatch(fid_dev, spec_dev, batch, n_in, n_out, pad_left, window_mode, window_dev, win_rows_host, scale, inverse,
                       in_shift, out_shift, absmax_dev, argmax_dev, phase_mode, ph_a_turns, ph_b_turns, nullptr, nullptr, 0, nullptr, stream);
}

int xmr_fid_absmax_pruned_c64(const void* fid_dev, int64_t batch, int n_in, int n_out, int pad_left, int window_mode,
                              const float* window_dev, const float* win_rows_host, float scale, float* absmax_dev,
                              float* running_max2_dev, int reset_running_max, void* stream) {
    if (!absmax_dev || !running_max2_dev) return fail(XMR_ERR_BAD_ARG, "NULL pointer");
    if (reset_running_max) {
        cudaError_t e = cudaMemsetAsync(running_max2_dev, 0, sizeof(float), static_cast<cudaStream_t>(stream));
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(running max)");
    }
    return k1_dispatch(fid_dev, nullptr, batch, n_in, n_out, pad_left, window_mode, window_dev, win_rows_host, scale, 0, 0,
                       n_out / 2, absmax_dev, nullptr, XMR_PHASE_NONE, 0.0, 0.0, running_max2_dev, nullptr, 0, nullptr, stream);
}

int xmr_zero_fill_c64(const void* in_dev, void* out_dev, int64_t batch, int n_in, int n_out, int pad_left,
                      void* stream) {
    if (batch < 0 || n_in < 0 || n_out < n_in || pad_left < 0 || pad_left + n_in > n_out)
        return fail(XMR_ERR_BAD_ARG, "bad sizes: batch=%lld n_in=%d n_out=%d pad_left=%d", (long long)batch, n_in, n_out,
                    pad_left);
    if (batch == 0 || n_out == 0) return XMR_OK;
    if (!in_dev || !out_dev) return fail(XMR_ERR_BAD_ARG, "NULL pointer");
    zero_fill_kernel<<<grid_for(batch * n_out, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float2*>(in_dev), static_cast<float2*>(out_dev), batch, n_in, n_out, pad_left);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? XMR_OK : cuda_fail(e, "zero_fill launch");
}

int xmr_roll_rows_c64(const void* in_dev, void* out_dev, int64_t batch, int n, int shift, void* stream) {
    if (batch < 0 || n < 0) return fail(XMR_ERR_BAD_ARG, "bad sizes");
    if (batch == 0 || n == 0) return XMR_OK;
    if (!in_dev || !out_dev || in_dev == out_dev) return fail(XMR_ERR_BAD_ARG, "NULL or aliased pointer");
    shift %= n;
    if (shift < 0) shift += n;
    roll_rows_kernel<<<grid_for(batch * n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float2*>(in_dev), static_cast<float2*>(out_dev), batch, n, shift);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? XMR_OK : cuda_fail(e, "roll_rows launch");
}

int xmr_scale_rows_c64(const void* in_dev, void* out_dev, int64_t batch, int n, const float* w_dev, void* stream) {
    if (batch < 0 || n < 0) return fail(XMR_ERR_BAD_ARG, "bad sizes");
    if (batch == 0 || n == 0) return XMR_OK;
    if (!in_dev || !out_dev || !w_dev) return fail(XMR_ERR_BAD_ARG, "NULL pointer");
    scale_rows_kernel<<<grid_for(batch * n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float2*>(in_dev), static_cast<float2*>(out_dev), batch, n, w_dev);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? XMR_OK : cuda_fail(e, "scale_rows launch");
}

int xmr_rotate_rows_c64(const void* in_dev, void* out_dev, int64_t batch, int n, const void* rot_dev, void* stream) {
    if (batch < 0 || n < 0) return fail(XMR_ERR_BAD_ARG, "bad sizes");
    if (batch == 0 || n == 0) return XMR_OK;
    if (!in_dev || !out_dev || !rot_dev) return fail(XMR_ERR_BAD_ARG, "NULL pointer");
    rotate_rows_kernel<<<grid_for(batch * n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float2*>(in_dev), static_cast<float2*>(out_dev), batch, n, static_cast<const float2*>(rot_dev));
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? XMR_OK : cuda_fail(e, "rotate_rows launch");
}

int xmr_rotate_rows_shift_c64(const void* in_dev, int64_t in_stride, void* out_dev, int64_t batch, int n, const void* rot_dev,
                              int in_shift, int out_shift, void* stream) {
    if (batch < 0 || n < 0 || in_stride < n) return fail(XMR_ERR_BAD_ARG, "bad sizes");
    if (batch == 0 || n == 0) return XMR_OK;
    if (!in_dev || !out_dev || !rot_dev) return fail(XMR_ERR_BAD_ARG, "NULL pointer");
    if (in_dev == out_dev && (in_shift != out_shift || in_stride != n)) return fail(XMR_ERR_BAD_ARG, "in place only without a shift");
    in_shift = ((in_shift % n) + n) % n;
    out_shift = ((out_shift % n) + n) % n;
    rotate_rows_shift_kernel<<<grid_for(batch * n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float2*>(in_dev), in_stride, static_cast<float2*>(out_dev), batch, n,
        static_cast<const float2*>(rot_dev), in_shift, out_shift);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? XMR_OK : cuda_fail(e, "rotate_rows_shift launch");
}

int xmr_phase_each_c64(const void* in_dev, void* out_dev, int64_t batch, int n, const double* a_turns_dev,
                       const double* b_turns_dev, void* stream) {
    if (batch < 0 || n < 0) return fail(XMR_ERR_BAD_ARG, "bad sizes");
    if (batch == 0 || n == 0) return XMR_OK;
    if (!in_dev || !out_dev || !a_turns_dev || !b_turns_dev) return fail(XMR_ERR_BAD_ARG, "NULL pointer");
    const int grid = int(batch < 148LL * 16 ? batch : 148LL * 16);
    phase_each_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float2*>(in_dev), static_cast<float2*>(out_dev), batch, n, a_turns_dev, b_turns_dev);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? XMR_OK : cuda_fail(e, "phase_each launch");
}

int xmr_global_argmax(const float* absmax_dev, const int* argmax_dev, int64_t batch, int n, void* out_dev,
                      void* stream) {
    if (batch < 0 || n < 1) return fail(XMR_ERR_BAD_ARG, "bad sizes");
    if (!absmax_dev || !out_dev) return fail(XMR_ERR_BAD_ARG, "NULL pointer");
    static std::atomic<unsigned> counter{0};
    const int slot = int(counter.fetch_add(1) & 7u);   // partials of up to 8 in-flight calls never collide
    long long blocks = (batch + 1023) / 1024;
    if (blocks > ARGMAX_BLOCKS) blocks = ARGMAX_BLOCKS;
    if (blocks < 1) blocks = 1;
    global_argmax_stage1<<<int(blocks), 1024, 0, static_cast<cudaStream_t>(stream)>>>(absmax_dev, batch, slot);
    global_argmax_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(absmax_dev, argmax_dev, batch, n,
                                                                           static_cast<unsigned char*>(out_dev), int(blocks), slot);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? XMR_OK : cuda_fail(e, "global_argmax launch");
}

}  // extern "C"
