// C ABI: the whole chain on HOST buffers in one call -- what a numpy-side binding of the reference would use.
//
//   xmr_chain_host_c64: pinned or pageable host FIDs -> (H2D || pass 1) -> argmax + search -> (pass 2 || D2H) -> host spectra
//
// Copies and kernels are pipelined over voxel chunks on three internal streams (created once per thread); device
// scratch is a per-thread workspace that grows on demand and is released by xmr_host_workspace_release().
// Pageable buffers are page-locked for the duration of the call (cudaHostRegister) so that the copies are truly
// asynchronous; already pinned buffers are used as they are.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/xmris_b200.h"
#include "abi_common.h"

namespace xmr_abi {
int k1_dispatch(const void* fid_dev, void* spec_dev, int64_t batch, int n_in, int n_out, int pad_left, int window_mode,
                const float* window_dev, const float* win_rows_host, float scale, int inverse, int in_shift, int out_shift,
                float* absmax_dev, int* argmax_dev, int phase_mode, double ph_a_turns, double ph_b_turns, float* run_max2,
                const long long* row_slot_dev, int row_stride, const void* ph_dev, void* stream);        // xmris_abi.cu
int autophase_search_dev(const void* spec_dev, int n, double du, int method, const int* pivot_dev, int fixed_pivot, double u0_fixed,
                         int fixed_target, int index_width, int p0_only, double* result_dev, void* workspace_dev, void* ph_out,
                         void* stream);                                                                  // autophase_abi.cu
int pack_winner(const void* fid_dev, int64_t batch, int n_in, int n_out, const void* arg_record_dev, int64_t row_offset,
                void* slot_dev, void* stream);                                                           // xmris_abi.cu
int select_winner(const void* gathered_dev, int world, int n_in, void* ctl_dev, void* stream);
}  // namespace xmr_abi

namespace {

struct Workspace {
    cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
    void* d_in = nullptr;
    size_t d_in_bytes = 0;
    void* d_out[2] = {nullptr, nullptr};
    size_t d_out_bytes = 0;
    void* d_small = nullptr;   // absmax / argmax / per-voxel results / window / search result + workspace
    size_t d_small_bytes = 0;
    int device = -1;
};
thread_local Workspace g_ws;

#define XMR_CU(call)                                                   \
    do {                                                               \
        cudaError_t e_ = (call);                                       \
        if (e_ != cudaSuccess) return xmr_abi::cuda_fail(e_, #call);   \
    } while (0)

int grow(void** p, size_t* have, size_t need) {
    if (*have >= need) return XMR_OK;
    if (*p) XMR_CU(cudaFree(*p));
    *p = nullptr;
    *have = 0;
    XMR_CU(cudaMalloc(p, need));
    *have = need;
    return XMR_OK;
}

int ensure_workspace(size_t in_bytes, size_t out_chunk_bytes, size_t small_bytes) {
    int dev = 0;
    XMR_CU(cudaGetDevice(&dev));
    Workspace& w = g_ws;
    if (w.device != dev) {
        if (w.device >= 0) return xmr_abi::fail(XMR_ERR_BAD_ARG, "host workspace belongs to device %d; call xmr_host_workspace_release() first", w.device);
        XMR_CU(cudaStreamCreateWithFlags(&w.s_in, cudaStreamNonBlocking));
        XMR_CU(cudaStreamCreateWithFlags(&w.s_cmp, cudaStreamNonBlocking));
        XMR_CU(cudaStreamCreateWithFlags(&w.s_out, cudaStreamNonBlocking));
        w.device = dev;
    }
    int rc = grow(&w.d_in, &w.d_in_bytes, in_bytes);
    if (rc != XMR_OK) return rc;
    if (w.d_out_bytes < out_chunk_bytes) {
        for (int i = 0; i < 2; ++i) {
            if (w.d_out[i]) XMR_CU(cudaFree(w.d_out[i]));
            w.d_out[i] = nullptr;
        }
        w.d_out_bytes = 0;
        for (int i = 0; i < 2; ++i) XMR_CU(cudaMalloc(&w.d_out[i], out_chunk_bytes));
        w.d_out_bytes = out_chunk_bytes;
    }
    return grow(&w.d_small, &w.d_small_bytes, small_bytes);
}

bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

struct HostPin {   // page-lock a pageable range for the duration of the call
    void* p = nullptr;
    bool registered = false;
    int lock(const void* ptr, size_t bytes) {
        if (bytes == 0 || is_pinned(ptr)) return XMR_OK;
        cudaError_t e = cudaHostRegister(const_cast<void*>(ptr), bytes, cudaHostRegisterDefault);
        if (e != cudaSuccess) {
            cudaGetLastError();   // fall back to staged (synchronous) copies: still correct
            return XMR_OK;
        }
        p = const_cast<void*>(ptr);
        registered = true;
        return XMR_OK;
    }
    ~HostPin() {
        if (registered) cudaHostUnregister(p);
    }
};

// factor w[M*n1 + n2] = cols[n2] * rows[n1] when the window allows it (exponential on a uniform axis)
bool split_window(const double* w, int n, std::vector<float>& cols, std::vector<float>& rows) {
    const int m = n < 256 ? n : 256, r0 = n / m;
    cols.resize(m);
    rows.assign(32, 1.0f);
    if (r0 == 1) {
        for (int i = 0; i < m; ++i) cols[i] = float(w[i]);
        return true;
    }
    if (w[0] == 0.0 || !std::isfinite(w[0])) return false;
    for (int n1 = 0; n1 < r0; ++n1) {
        const double r = w[size_t(n1) * m] / w[0];
        for (int n2 = 0; n2 < m; ++n2) {
            const double want = w[size_t(n1) * m + n2], got = r * w[n2];
            if (!(std::fabs(want - got) <= 1e-9 * std::fabs(want) + 1e-300)) return false;
        }
        rows[n1] = float(r);
    }
    for (int i = 0; i < m; ++i) cols[i] = float(w[i]);
    return true;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

std::atomic<long long> g_graph_launches{0};
std::atomic<long long> g_resident_limit{0};    // xmr_host_chain_resident_limit: bytes of FIDs mode=single may keep on the device

}  // namespace

extern "C" {

int xmr_host_chain_resident_limit(int64_t bytes) {
    if (bytes < 0) return xmr_abi::fail(XMR_ERR_BAD_ARG, "resident limit must be >= 0 (0: whatever is free)");
    g_resident_limit.store(bytes);
    return XMR_OK;
}

int xmr_host_workspace_release(void) {
    Workspace& w = g_ws;
    if (w.device < 0) return XMR_OK;
    cudaDeviceSynchronize();
    if (w.d_in) cudaFree(w.d_in);
    for (int i = 0; i < 2; ++i)
        if (w.d_out[i]) cudaFree(w.d_out[i]);
    if (w.d_small) cudaFree(w.d_small);
    if (w.s_in) cudaStreamDestroy(w.s_in);
    if (w.s_cmp) cudaStreamDestroy(w.s_cmp);
    if (w.s_out) cudaStreamDestroy(w.s_out);
    w = Workspace();
    return XMR_OK;
}

int xmr_chain_host_c64(const xmr_host_chain_desc* d, const void* fid_host, void* out_host, int64_t batch,
                       double* result_host, double* p0_host, double* p1_host, int* pivot_host, float* fun_host) {
    if (!d) return xmr_abi::fail(XMR_ERR_BAD_ARG, "descriptor is NULL");
    const int n_in = d->n_in, n_out = d->n_out;
    if (batch < 0 || n_in < 1 || n_out < n_in || d->pad_left < 0 || d->pad_left + n_in > n_out)
        return xmr_abi::fail(XMR_ERR_BAD_ARG, "bad sizes: batch=%lld n_in=%d n_out=%d pad_left=%d", (long long)batch, n_in, n_out, d->pad_left);
    if (!(n_out >= 16 && n_out <= 8192 && (n_out & (n_out - 1)) == 0))
        return xmr_abi::fail(XMR_ERR_UNSUPPORTED_N, "n_out=%d: the host chain needs a power-of-two length in [16, 8192]", n_out);
    if (d->autophase_mode < 0 || d->autophase_mode > 2) return xmr_abi::fail(XMR_ERR_BAD_ARG, "autophase_mode=%d", d->autophase_mode);
    if (d->autophase_mode == 2 && n_out < 512) return xmr_abi::fail(XMR_ERR_UNSUPPORTED_N, "per-spectrum autophase needs n_out >= 512");
    if (batch == 0) return XMR_OK;
    if (!fid_host || !out_host) return xmr_abi::fail(XMR_ERR_BAD_ARG, "NULL host buffer");
    if (d->autophase_mode == 1 && !result_host) return xmr_abi::fail(XMR_ERR_BAD_ARG, "result_host is NULL");
    if (d->autophase_mode == 2 && (!p0_host || !p1_host || !pivot_host))
        return xmr_abi::fail(XMR_ERR_BAD_ARG, "per-spectrum outputs are NULL");

    const int64_t chunk = std::min<int64_t>(batch, d->chunk > 0 ? d->chunk : 8192);
    const size_t row_in = size_t(n_in) * 8, row_out = size_t(n_out) * 8;
    // small device area: absmax[batch] | p0[batch] p1[batch] (double) | pivot[batch] fun[batch] | window | result | search ws
    const size_t o_absmax = 0;
    const size_t o_p0 = align_up(o_absmax + size_t(batch) * 4, 256);
    const size_t o_p1 = o_p0 + size_t(batch) * 8;
    const size_t o_piv = o_p1 + size_t(batch) * 8;
    const size_t o_fun = o_piv + size_t(batch) * 4;
    const size_t o_win = align_up(o_fun + size_t(batch) * 4, 256);
    const size_t o_res = align_up(o_win + size_t(n_out) * 4, 256);
    const size_t o_arg = o_res + 64;
    const size_t o_row = align_up(o_arg + 64, 256);                       // one spectrum + its stats
    const size_t o_sws = align_up(o_row + row_out + 64, 256);
    const size_t small = o_sws + size_t(xmr_autophase_workspace_bytes());
    bool keep_all_in = (d->autophase_mode == 1);                           // the FIDs are read twice
    if (keep_all_in) {
        // mode="single" keeps the whole FID batch on the device between its two passes when it fits (and stays under the
        // caller's limit, xmr_host_chain_resident_limit); a larger data set is streamed through twice instead -- pass 1 over
        // double-buffered chunks, the winning row fetched again, pass 2 over the re-uploaded chunks
        size_t free_b = 0, total_b = 0;
        XMR_CU(cudaMemGetInfo(&free_b, &total_b));
        const size_t need = size_t(batch) * row_in + 2 * size_t(chunk) * row_out + small;
        const long long limit = g_resident_limit.load();
        if (need > free_b + g_ws.d_in_bytes + 2 * g_ws.d_out_bytes + g_ws.d_small_bytes ||
            (limit > 0 && size_t(batch) * row_in > size_t(limit)))
            keep_all_in = false;
    }
    const bool restream = (d->autophase_mode == 1) && !keep_all_in;
    int rc = ensure_workspace(keep_all_in ? size_t(batch) * row_in : 2 * size_t(chunk) * row_in, size_t(chunk) * row_out, small);
    if (rc != XMR_OK) return rc;
    Workspace& w = g_ws;
    unsigned char* sm = static_cast<unsigned char*>(w.d_small);
    float* absmax = reinterpret_cast<float*>(sm + o_absmax);

    HostPin pin_in, pin_out;
    pin_in.lock(fid_host, size_t(batch) * row_in);
    pin_out.lock(out_host, size_t(batch) * row_out);

    // window
    int win_mode = XMR_WIN_NONE;
    float* win_dev = nullptr;
    std::vector<float> cols, rows(32, 1.0f), table;
    if (d->window_host) {
        win_dev = reinterpret_cast<float*>(sm + o_win);
        if (split_window(d->window_host, n_out, cols, rows)) {
            win_mode = XMR_WIN_SEPARABLE;
            XMR_CU(cudaMemcpyAsync(win_dev, cols.data(), cols.size() * 4, cudaMemcpyHostToDevice, w.s_cmp));
        } else {
            win_mode = XMR_WIN_TABLE;
            table.resize(n_out);
            for (int i = 0; i < n_out; ++i) table[i] = float(d->window_host[i]);
            XMR_CU(cudaMemcpyAsync(win_dev, table.data(), table.size() * 4, cudaMemcpyHostToDevice, w.s_cmp));
        }
        XMR_CU(cudaStreamSynchronize(w.s_cmp));   // cols/table are stack-lifetime host vectors
    }
    const float scale = d->scale != 0.f ? d->scale : 1.0f / std::sqrt(float(n_out));
    const int64_t nchunks = (batch + chunk - 1) / chunk;
    std::vector<cudaEvent_t> in_done(nchunks, nullptr);
    cudaEvent_t out_free[2] = {nullptr, nullptr}, cmp_done[2] = {nullptr, nullptr}, in_free[2] = {nullptr, nullptr};
    auto cleanup = [&]() {          // (also reached when an event creation below fails: destroys what exists)
        for (auto& e : in_done)
            if (e) cudaEventDestroy(e);
        for (int i = 0; i < 2; ++i) {
            if (out_free[i]) cudaEventDestroy(out_free[i]);
            if (cmp_done[i]) cudaEventDestroy(cmp_done[i]);
            if (in_free[i]) cudaEventDestroy(in_free[i]);
        }
    };
    {
        cudaError_t ee = cudaSuccess;
        for (auto& e : in_done)
            if (ee == cudaSuccess) ee = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        for (int i = 0; i < 2 && ee == cudaSuccess; ++i) {
            ee = cudaEventCreateWithFlags(&out_free[i], cudaEventDisableTiming);
            if (ee == cudaSuccess) ee = cudaEventCreateWithFlags(&cmp_done[i], cudaEventDisableTiming);
            if (ee == cudaSuccess) ee = cudaEventCreateWithFlags(&in_free[i], cudaEventDisableTiming);
        }
        if (ee != cudaSuccess) {
            cleanup();
            return xmr_abi::cuda_fail(ee, "cudaEventCreateWithFlags");
        }
    }
    const unsigned char* h_in = static_cast<const unsigned char*>(fid_host);
    unsigned char* h_out = static_cast<unsigned char*>(out_host);
    unsigned char* d_in = static_cast<unsigned char*>(w.d_in);
    auto d_in_chunk = [&](int64_t c) { return keep_all_in ? d_in + size_t(c) * chunk * row_in : d_in + size_t(c % 2) * chunk * row_in; };

#define XMR_RC(call)            \
    do {                        \
        rc = (call);            \
        if (rc != XMR_OK) {     \
            cudaDeviceSynchronize(); \
            cleanup();          \
            return rc;          \
        }                       \
    } while (0)
#define XMR_CUC(call)                                      \
    do {                                                   \
        cudaError_t e_ = (call);                           \
        if (e_ != cudaSuccess) {                           \
            cudaDeviceSynchronize();                       \
            cleanup();                                     \
            return xmr_abi::cuda_fail(e_, #call);          \
        }                                                  \
    } while (0)

    double ph_a = 0.0, ph_b = 0.0;
    if (d->autophase_mode == 1) {
        // ---- pass 1 trails the uploads chunk by chunk ---------------------------------------------------------------
        for (int64_t c = 0; c < nchunks; ++c) {
            const int64_t lo = c * chunk, nb = std::min(chunk, batch - lo);
            if (restream && c >= 2) XMR_CUC(cudaStreamWaitEvent(w.s_in, in_free[c % 2], 0));   // slot last read by chunk c-2
            XMR_CUC(cudaMemcpyAsync(d_in_chunk(c), h_in + size_t(lo) * row_in, size_t(nb) * row_in, cudaMemcpyHostToDevice, w.s_in));
            XMR_CUC(cudaEventRecord(in_done[c], w.s_in));
            XMR_CUC(cudaStreamWaitEvent(w.s_cmp, in_done[c], 0));
            // branch-and-bound statistics: the running maximum (one float at o_arg + 32) is shared by the chunks of this call
            XMR_RC(xmr_fid_absmax_pruned_c64(d_in_chunk(c), nb, n_in, n_out, d->pad_left, win_mode, win_dev, rows.data(), scale,
                                             absmax + lo, reinterpret_cast<float*>(sm + o_arg + 32), c == 0 ? 1 : 0, w.s_cmp));
            if (restream) XMR_CUC(cudaEventRecord(in_free[c % 2], w.s_cmp));
        }
        // ---- winner, its spectrum, the search ------------------------------------------------------------------------
        unsigned char* arg = sm + o_arg;
        XMR_RC(xmr_global_argmax(absmax, nullptr, batch, n_out, arg, w.s_cmp));
        unsigned char arg_h[16];
        XMR_CUC(cudaMemcpyAsync(arg_h, arg, 16, cudaMemcpyDeviceToHost, w.s_cmp));
        XMR_CUC(cudaStreamSynchronize(w.s_cmp));
        long long flat;
        std::memcpy(&flat, arg_h + 8, 8);
        const long long row = flat / n_out;
        unsigned char* rowbuf = sm + o_row;
        float* row_abs = reinterpret_cast<float*>(rowbuf + row_out);
        int* row_arg = reinterpret_cast<int*>(rowbuf + row_out + 16);
        const unsigned char* row_dev = d_in + size_t(row) * row_in;
        if (restream) {      // the winning FID is no longer on the device (s_cmp is idle here: both ring slots are free)
            XMR_CUC(cudaMemcpyAsync(d_in, h_in + size_t(row) * row_in, row_in, cudaMemcpyHostToDevice, w.s_cmp));
            row_dev = d_in;
        }
        XMR_RC(xmr_fid_to_spectrum_c64(row_dev, rowbuf, 1, n_in, n_out, d->pad_left, win_mode, win_dev, rows.data(),
                                       scale, 0, 0, n_out / 2, row_abs, row_arg, XMR_PHASE_NONE, 0.0, 0.0, w.s_cmp));
        int idx = 0;
        XMR_CUC(cudaMemcpyAsync(&idx, row_arg, 4, cudaMemcpyDeviceToHost, w.s_cmp));
        XMR_CUC(cudaStreamSynchronize(w.s_cmp));
        const int target = d->fixed_pivot ? d->fixed_target : idx;
        const double u0 = d->fixed_pivot ? d->u0_fixed : -d->du * double(idx);
        double* res = reinterpret_cast<double*>(sm + o_res);
        XMR_RC(xmr_autophase_search_c64(rowbuf, n_out, u0, d->du, d->method, target, d->index_width > 0 ? d->index_width : 1, d->p0_only,
                                        res, sm + o_sws, w.s_cmp));
        double res_h[4];
        XMR_CUC(cudaMemcpyAsync(res_h, res, 32, cudaMemcpyDeviceToHost, w.s_cmp));
        XMR_CUC(cudaStreamSynchronize(w.s_cmp));
        const double p0 = res_h[0], p1 = d->p0_only ? 0.0 : res_h[1];
        result_host[0] = p0;
        result_host[1] = p1;
        result_host[2] = double(idx);
        result_host[3] = res_h[2];
        ph_a = p0 / 360.0 + (p1 / 360.0) * u0;
        ph_b = (p1 / 360.0) * d->du;
    }
    // ---- output pass per chunk with the write-back trailing it ----------------------------------------------------------
    for (int64_t c = 0; c < nchunks; ++c) {
        const int64_t lo = c * chunk, nb = std::min(chunk, batch - lo);
        const int k = int(c % 2);
        if (!keep_all_in) {
            if (c >= 2) XMR_CUC(cudaStreamWaitEvent(w.s_in, in_free[k], 0));   // input slot k was last read by chunk c-2
            XMR_CUC(cudaMemcpyAsync(d_in_chunk(c), h_in + size_t(lo) * row_in, size_t(nb) * row_in, cudaMemcpyHostToDevice, w.s_in));
            XMR_CUC(cudaEventRecord(in_done[c], w.s_in));
        }
        XMR_CUC(cudaStreamWaitEvent(w.s_cmp, in_done[c], 0));
        if (c >= 2) XMR_CUC(cudaStreamWaitEvent(w.s_cmp, out_free[k], 0));
        if (d->autophase_mode == 2) {
            XMR_RC(xmr_chain_each_c64(d_in_chunk(c), w.d_out[k], nb, n_in, n_out, d->pad_left, 0, win_mode, win_dev, rows.data(), scale,
                                      d->method, d->du, d->fixed_pivot, d->u0_fixed, d->fixed_target, d->index_width > 0 ? d->index_width : 1,
                                      d->p0_only, reinterpret_cast<double*>(sm + o_p0) + lo, reinterpret_cast<double*>(sm + o_p1) + lo,
                                      reinterpret_cast<int*>(sm + o_piv) + lo, reinterpret_cast<float*>(sm + o_fun) + lo, w.s_cmp));
        } else {
            XMR_RC(xmr_fid_to_spectrum_c64(d_in_chunk(c), w.d_out[k], nb, n_in, n_out, d->pad_left, win_mode, win_dev, rows.data(), scale, 0, 0,
                                           n_out / 2, nullptr, nullptr, d->autophase_mode == 1 ? XMR_PHASE_UNIFORM : XMR_PHASE_NONE, ph_a, ph_b,
                                           w.s_cmp));
        }
        XMR_CUC(cudaEventRecord(cmp_done[k], w.s_cmp));
        if (!keep_all_in) XMR_CUC(cudaEventRecord(in_free[k], w.s_cmp));
        XMR_CUC(cudaStreamWaitEvent(w.s_out, cmp_done[k], 0));
        XMR_CUC(cudaMemcpyAsync(h_out + size_t(lo) * row_out, w.d_out[k], size_t(nb) * row_out, cudaMemcpyDeviceToHost, w.s_out));
        XMR_CUC(cudaEventRecord(out_free[k], w.s_out));
    }
    if (d->autophase_mode == 2) {
        XMR_CUC(cudaMemcpyAsync(p0_host, sm + o_p0, size_t(batch) * 8, cudaMemcpyDeviceToHost, w.s_cmp));
        XMR_CUC(cudaMemcpyAsync(p1_host, sm + o_p1, size_t(batch) * 8, cudaMemcpyDeviceToHost, w.s_cmp));
        XMR_CUC(cudaMemcpyAsync(pivot_host, sm + o_piv, size_t(batch) * 4, cudaMemcpyDeviceToHost, w.s_cmp));
        if (fun_host) XMR_CUC(cudaMemcpyAsync(fun_host, sm + o_fun, size_t(batch) * 4, cudaMemcpyDeviceToHost, w.s_cmp));
    }
    XMR_CUC(cudaStreamSynchronize(w.s_cmp));
    XMR_CUC(cudaStreamSynchronize(w.s_out));
    XMR_CUC(cudaStreamSynchronize(w.s_in));
    cleanup();
    return XMR_OK;
#undef XMR_RC
#undef XMR_CUC
}

// ---- the same chain on DEVICE-resident data: autophase(mode="single") end to end in one call ---------------------------
// (what the survey's proposed `xmr_autophase_c64(fid, out, ..., mode=single)` asks for: pass 1 with branch and bound ->
// global argmax -> the winning spectrum -> (p0, p1) search -> pass 2 with the fused phase; three small device->host reads,
// no host code between the launches but the winner bookkeeping.)
int64_t xmr_chain_single_graph_launches(void) { return g_graph_launches.load(); }

int64_t xmr_chain_single_slot_bytes(int n_in) { return n_in > 0 ? int64_t(((n_in + 1) & ~1) + 2) * 8 : 0; }

int64_t xmr_chain_single_workspace_bytes(int64_t batch, int n_out) {
    if (batch < 0 || n_out < 1) return 0;
    return int64_t(align_up(size_t(batch) * 4, 256) + 256 + 512 + align_up(size_t(n_out) * 8 + 64, 256) +
                   align_up(size_t(n_out + 4) * 8, 256) + size_t(xmr_autophase_workspace_bytes()));
}

}  // extern "C"

namespace {

struct ChainLayout {
    size_t o_arg, o_ph, o_row, o_slot, o_sws;
    ChainLayout(int64_t batch, int n_out) {
        o_arg = align_up(size_t(batch) * 4, 256);     // control block (256 B), see chain_back
        o_ph = o_arg + 256;                            // K1PhaseDev of pass 2
        o_row = o_ph + 512;                            // the winning spectrum
        o_slot = o_row + align_up(size_t(n_out) * 8 + 64, 256);   // this rank's candidate slot (single-GPU: the gathered buffer)
        o_sws = o_slot + align_up(size_t(n_out + 4) * 8, 256);    // search scratch
    }
};

int check_chain_args(const xmr_host_chain_desc* d, int64_t batch) {
    if (!d) return xmr_abi::fail(XMR_ERR_BAD_ARG, "descriptor is NULL");
    const int n_in = d->n_in, n_out = d->n_out;
    if (batch < 0 || n_in < 1 || n_out < n_in || d->pad_left < 0 || d->pad_left + n_in > n_out)
        return xmr_abi::fail(XMR_ERR_BAD_ARG, "bad sizes: batch=%lld n_in=%d n_out=%d pad_left=%d", (long long)batch, n_in, n_out, d->pad_left);
    if (!(n_out >= 16 && n_out <= 8192 && (n_out & (n_out - 1)) == 0))
        return xmr_abi::fail(XMR_ERR_UNSUPPORTED_N, "n_out=%d: the device chain needs a power-of-two length in [16, 8192]", n_out);
    return XMR_OK;
}

#define XMR_RC(call)                  \
    do {                              \
        rc = (call);                  \
        if (rc != XMR_OK) return rc;  \
    } while (0)

// pass 1 (branch and bound) -> this shard's argmax record -> candidate slot {best FID row, max |S|, global row}
int chain_front(const xmr_host_chain_desc* d, const void* fid_dev, int64_t batch, int window_mode, const float* window_dev,
                const float* rows, float scale, unsigned char* sm, int64_t row_offset, void* slot_dev, cudaStream_t st) {
    const ChainLayout L(batch, d->n_out);
    float* absmax = reinterpret_cast<float*>(sm);
    int rc;
    if (batch > 0) {
        XMR_RC(xmr_fid_absmax_pruned_c64(fid_dev, batch, d->n_in, d->n_out, d->pad_left, window_mode, window_dev, rows, scale, absmax,
                                         reinterpret_cast<float*>(sm + L.o_arg + 32), 1, st));
        XMR_RC(xmr_global_argmax(absmax, nullptr, batch, d->n_out, sm + L.o_arg, st));
    }
    return xmr_abi::pack_winner(fid_dev, batch, d->n_in, d->n_out, sm + L.o_arg, row_offset, slot_dev, st);
}

// global winner among `world` gathered slots -> its spectrum + pivot -> (p0, p1) search -> phase parameters (device) ->
// 256-byte control block to pinned host memory.   Control block: {float max @0, int64 winning slot @8, int64 global row @16,
// running max @32, search result double[4] @64, winner's |S| max float @128, its argmax int @144}
int chain_mid(const xmr_host_chain_desc* d, int64_t batch, int window_mode, const float* window_dev, const float* rows, float scale,
              unsigned char* sm, const void* gathered_dev, int world, unsigned char* ctl_pinned, cudaStream_t st) {
    const ChainLayout L(batch, d->n_out);
    const int n_in = d->n_in, n_out = d->n_out;
    int rc;
    XMR_RC(xmr_abi::select_winner(gathered_dev, world, n_in, sm + L.o_arg, st));
    unsigned char* rowbuf = sm + L.o_row;
    float* row_abs = reinterpret_cast<float*>(sm + L.o_arg + 128);
    int* row_arg = reinterpret_cast<int*>(sm + L.o_arg + 144);
    XMR_RC(xmr_abi::k1_dispatch(gathered_dev, rowbuf, 1, n_in, n_out, d->pad_left, window_mode, window_dev, rows, scale, 0, 0, n_out / 2,
                                row_abs, row_arg, XMR_PHASE_NONE, 0.0, 0.0, nullptr,
                                reinterpret_cast<const long long*>(sm + L.o_arg + 8), ((n_in + 1) & ~1) + 2, nullptr, st));
    double* res = reinterpret_cast<double*>(sm + L.o_arg + 64);
    XMR_RC(xmr_abi::autophase_search_dev(rowbuf, n_out, d->du, d->method, row_arg, d->fixed_pivot, d->u0_fixed, d->fixed_target,
                                         d->index_width > 0 ? d->index_width : 1, d->p0_only, res, sm + L.o_sws, sm + L.o_ph, st));
    XMR_CU(cudaMemcpyAsync(ctl_pinned, sm + L.o_arg, 256, cudaMemcpyDeviceToHost, st));
    return XMR_OK;
}

// per-thread state of the device chain: pinned landing buffer, events, capture stream, a small graph cache
struct GraphEntry {
    unsigned char key[160];
    int seen = 0;
    cudaGraphExec_t exec = nullptr;
};
struct ChainState {
    unsigned char* ctl_pinned = nullptr;
    cudaEvent_t t_begin = nullptr, searched = nullptr, t_end = nullptr;
    cudaStream_t cap = nullptr;
    int dev = -1;
    GraphEntry cache[4];
    int cache_next = 0;
    bool timed = false;
};
thread_local ChainState g_cs;

int chain_state(ChainState** out) {
    int dev = 0;
    XMR_CU(cudaGetDevice(&dev));
    ChainState& s = g_cs;
    if (s.dev != dev) {
        if (s.ctl_pinned) {
            cudaFreeHost(s.ctl_pinned);
            cudaEventDestroy(s.t_begin);
            cudaEventDestroy(s.searched);
            cudaEventDestroy(s.t_end);
            cudaStreamDestroy(s.cap);
        }
        for (GraphEntry& e : s.cache) {
            if (e.exec) cudaGraphExecDestroy(e.exec);
            e = GraphEntry();
        }
        XMR_CU(cudaMallocHost(reinterpret_cast<void**>(&s.ctl_pinned), 256));
        XMR_CU(cudaEventCreate(&s.t_begin));
        XMR_CU(cudaEventCreate(&s.searched));
        XMR_CU(cudaEventCreate(&s.t_end));
        XMR_CU(cudaStreamCreateWithFlags(&s.cap, cudaStreamNonBlocking));
        s.dev = dev;
        s.timed = false;
    }
    *out = &s;
    return XMR_OK;
}

// pass 2 with the phase parameters in device memory, then wait for the control block (pass 2 keeps running)
int chain_finish(const xmr_host_chain_desc* d, const void* fid_dev, void* spec_dev, int64_t batch, int window_mode,
                 const float* window_dev, const float* rows, float scale, unsigned char* sm, ChainState& cs, double* result_host,
                 cudaStream_t st) {
    const ChainLayout L(batch, d->n_out);
    int rc;
    XMR_CU(cudaEventRecord(cs.searched, st));
    if (batch > 0)
        XMR_RC(xmr_abi::k1_dispatch(fid_dev, spec_dev, batch, d->n_in, d->n_out, d->pad_left, window_mode, window_dev, rows, scale, 0, 0,
                                    d->n_out / 2, nullptr, nullptr, XMR_PHASE_UNIFORM, 0.0, 0.0, nullptr, nullptr, 0, sm + L.o_ph, st));
    XMR_CU(cudaEventRecord(cs.t_end, st));
    cs.timed = true;
    XMR_CU(cudaEventSynchronize(cs.searched));     // the control block has landed; pass 2 keeps running
    unsigned char ctl[256];
    std::memcpy(ctl, cs.ctl_pinned, 256);
    float vmax;
    long long grow;
    double res_h[4];
    int idx;
    std::memcpy(&vmax, ctl, 4);
    std::memcpy(&grow, ctl + 16, 8);
    std::memcpy(res_h, ctl + 64, 32);
    std::memcpy(&idx, ctl + 144, 4);
    result_host[0] = res_h[0];
    result_host[1] = d->p0_only ? 0.0 : res_h[1];
    result_host[2] = double(d->fixed_pivot ? d->fixed_target : idx);
    result_host[3] = res_h[2];
    result_host[4] = double(vmax);
    result_host[5] = double(grow);
    return XMR_OK;
}

}  // namespace

extern "C" {

// {front part ms (pass 1 ... search), pass 2 ms} of this thread's last device chain (waits for its pass 2)
int xmr_chain_single_last_timing(double* ms_out) {
    ChainState& cs = g_cs;
    if (!ms_out) return xmr_abi::fail(XMR_ERR_BAD_ARG, "NULL pointer");
    if (!cs.timed) return xmr_abi::fail(XMR_ERR_BAD_ARG, "no device chain has run on this thread");
    float a = 0.f, b = 0.f;
    XMR_CU(cudaEventSynchronize(cs.t_end));
    XMR_CU(cudaEventElapsedTime(&a, cs.t_begin, cs.searched));
    XMR_CU(cudaEventElapsedTime(&b, cs.searched, cs.t_end));
    ms_out[0] = double(a);
    ms_out[1] = double(b);
    return XMR_OK;
}

int xmr_chain_single_front_c64(const xmr_host_chain_desc* d, const void* fid_dev, int64_t batch, int window_mode,
                               const float* window_dev, const float* win_rows_host, void* workspace_dev, int64_t row_offset,
                               void* slot_dev, void* stream) {
    int rc = check_chain_args(d, batch);
    if (rc != XMR_OK) return rc;
    if ((batch > 0 && !fid_dev) || !workspace_dev || !slot_dev) return xmr_abi::fail(XMR_ERR_BAD_ARG, "NULL pointer");
    ChainState* cs = nullptr;
    XMR_RC(chain_state(&cs));
    const float scale = d->scale != 0.f ? d->scale : 1.0f / std::sqrt(float(d->n_out));
    float ones[32];
    for (float& r : ones) r = 1.0f;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    XMR_CU(cudaEventRecord(cs->t_begin, st));
    return chain_front(d, fid_dev, batch, window_mode, window_dev, win_rows_host ? win_rows_host : ones, scale,
                       static_cast<unsigned char*>(workspace_dev), row_offset, slot_dev, st);
}

int xmr_chain_single_back_c64(const xmr_host_chain_desc* d, const void* fid_dev, void* spec_dev, int64_t batch, int window_mode,
                              const float* window_dev, const float* win_rows_host, void* workspace_dev, const void* gathered_dev,
                              int world, double* result_host, void* stream) {
    int rc = check_chain_args(d, batch);
    if (rc != XMR_OK) return rc;
    if ((batch > 0 && (!fid_dev || !spec_dev)) || !workspace_dev || !gathered_dev || !result_host || world < 1)
        return xmr_abi::fail(XMR_ERR_BAD_ARG, "NULL pointer / world < 1");
    ChainState* cs = nullptr;
    XMR_RC(chain_state(&cs));
    const float scale = d->scale != 0.f ? d->scale : 1.0f / std::sqrt(float(d->n_out));
    float ones[32];
    for (float& r : ones) r = 1.0f;
    const float* rows = win_rows_host ? win_rows_host : ones;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned char* sm = static_cast<unsigned char*>(workspace_dev);
    XMR_RC(chain_mid(d, batch, window_mode, window_dev, rows, scale, sm, gathered_dev, world, cs->ctl_pinned, st));
    return chain_finish(d, fid_dev, spec_dev, batch, window_mode, window_dev, rows, scale, sm, *cs, result_host, st);
}

int xmr_chain_single_dev_c64(const xmr_host_chain_desc* d, const void* fid_dev, void* spec_dev, int64_t batch, int window_mode,
                             const float* window_dev, const float* win_rows_host, void* workspace_dev, double* result_host,
                             void* stream) {
    int rc = check_chain_args(d, batch);
    if (rc != XMR_OK) return rc;
    if (batch == 0) return XMR_OK;
    if (!fid_dev || !spec_dev || !workspace_dev || !result_host) return xmr_abi::fail(XMR_ERR_BAD_ARG, "NULL pointer");
    const int n_in = d->n_in, n_out = d->n_out;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned char* sm = static_cast<unsigned char*>(workspace_dev);
    const ChainLayout L(batch, n_out);
    const float scale = d->scale != 0.f ? d->scale : 1.0f / std::sqrt(float(n_out));
    float ones[32];
    for (float& r : ones) r = 1.0f;
    const float* rows = win_rows_host ? win_rows_host : ones;
    ChainState* csp = nullptr;
    XMR_RC(chain_state(&csp));
    ChainState& cs = *csp;
    // Everything between the two passes takes its inputs from device memory: the winning row (candidate slot + winner record),
    // the pivot (the winner's own argmax) and the phase parameters of pass 2 (written by the search's final kernel).  Nothing is
    // read back before pass 2 is enqueued.  The ~14 launches up to the search's final kernel (+ the 256-byte control block's
    // copy to pinned host memory) are captured ONCE per argument set into a CUDA graph and replayed: small batches are
    // launch-bound otherwise (C2: 4096 x 2048 spends more time between its kernels than in them).
    auto enqueue_front = [&](cudaStream_t s) -> int {
        int r = chain_front(d, fid_dev, batch, window_mode, window_dev, rows, scale, sm, 0, sm + L.o_slot, s);
        if (r != XMR_OK) return r;
        return chain_mid(d, batch, window_mode, window_dev, rows, scale, sm, sm + L.o_slot, 1, cs.ctl_pinned, s);
    };
    unsigned char key[160];
    std::memset(key, 0, sizeof(key));
    {
        size_t o = 0;
        auto put = [&](const void* v, size_t nbytes) { std::memcpy(key + o, v, nbytes); o += nbytes; };
        put(&fid_dev, 8); put(&workspace_dev, 8); put(&window_dev, 8); put(&batch, 8);
        put(&n_in, 4); put(&n_out, 4); put(&d->pad_left, 4); put(&window_mode, 4); put(&scale, 4);
        put(&d->method, 4); put(&d->index_width, 4); put(&d->p0_only, 4); put(&d->fixed_pivot, 4); put(&d->fixed_target, 4);
        put(&d->u0_fixed, 8); put(&d->du, 8);
        float rsum[2] = {0.f, 0.f};                              // the row factors travel by value into the kernel parameters
        for (int i = 0; i < 32; ++i) { rsum[0] += rows[i] * float(i + 1); rsum[1] += rows[i] * rows[i]; }
        put(rsum, 8);
    }
    GraphEntry* ent = nullptr;
    for (GraphEntry& e : cs.cache)
        if (e.seen && std::memcmp(e.key, key, sizeof(key)) == 0) ent = &e;
    bool launched = false;
    if (ent != nullptr && ent->exec == nullptr && ent->seen == 1) {
        // second call with these arguments (tables and kernel attributes exist now): capture
        ent->seen = 2;
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(cs.cap, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            const int crc = enqueue_front(cs.cap);
            const cudaError_t ce = cudaStreamEndCapture(cs.cap, &graph);
            if (crc == XMR_OK && ce == cudaSuccess && graph != nullptr) {
                if (cudaGraphInstantiate(&ent->exec, graph, 0) != cudaSuccess) ent->exec = nullptr;
            }
            if (graph) cudaGraphDestroy(graph);
        }
        cudaGetLastError();     // a failed capture leaves the eager path below (same kernels, same stream order)
    }
    XMR_CU(cudaEventRecord(cs.t_begin, st));
    if (ent != nullptr && ent->exec != nullptr) {
        XMR_CU(cudaGraphLaunch(ent->exec, st));
        launched = true;
        ++g_graph_launches;
    }
    if (!launched) {
        XMR_RC(enqueue_front(st));
        if (ent == nullptr) {
            GraphEntry& e = cs.cache[cs.cache_next];
            cs.cache_next = (cs.cache_next + 1) % 4;
            if (e.exec) cudaGraphExecDestroy(e.exec);
            e = GraphEntry();
            std::memcpy(e.key, key, sizeof(key));
            e.seen = 1;
        }
    }
    return chain_finish(d, fid_dev, spec_dev, batch, window_mode, window_dev, rows, scale, sm, cs, result_host, st);
}
#undef XMR_RC

}  // extern "C"
