// C ABI: baseline_als -- asymmetric least squares baseline (the step after autophase in the reference's pipeline,
// src/xmris/processing/baseline.py:10-119), batched over spectra.
//
// Per spectrum and iteration the reference solves (W + lam * D'D) z = W y with scipy's sparse LU, D the second-difference
// operator, then re-weights w = p*(y > z) + (1-p)*(y < z).  The matrix is symmetric positive definite and pentadiagonal:
// here ONE THREAD owns one spectrum and runs a banded LDL^T factorisation fused with the forward substitution, then the
// backward substitution fused with the re-weighting (float64 throughout: lam*|D'D| / min w ~ 1e9).  The factor rows
// (l1_{i+1}, l2_{i+2}, u_i/d_i) cannot stay on chip (24 B x n per spectrum): they stream through a per-CTA scratch area in
// HBM laid out [point][thread], so that every access of a warp is one contiguous 256-byte segment; the spectra themselves
// are transposed into / out of that layout through a padded shared-memory tile.
//
// Bound: HBM.  Bytes per point and iteration: forward 4 (y) + 1 (weight code) read, 24 written; backward 24 + 4 read,
// 1 written = 58 B, i.e. 58 * n * n_iter (+ 12 n in/out) per spectrum -- 2.4 MB at n = 4096, 10 iterations.
#include <cstdint>

#include "../../include/xmris_b200.h"
#include "abi_common.h"

namespace {

constexpr int ALS_TPB = 128;             // spectra per CTA (one per thread)
constexpr int ALS_MAX_CTAS_PER_SM = 4;
constexpr int ALS_U = 8;                 // points whose loads are in flight ahead of the recurrences

struct AlsParams {
    const void* in;
    float* out;
    long long batch;
    int n;
    double lam, p;
    int n_iter;
    unsigned char* scratch;
    size_t per_cta;
};

__host__ __device__ inline size_t als_per_cta(int n) {
    // Q, P, V double [n][TPB] | Y float [n][TPB] | W uint8 [n][TPB], rounded up to 256 bytes
    const size_t b = size_t(n) * ALS_TPB * (3 * sizeof(double) + sizeof(float) + 1);
    return (b + 255) / 256 * 256;
}

template <bool CPLX>
__global__ void __launch_bounds__(ALS_TPB) als_kernel(const __grid_constant__ AlsParams a) {
    __shared__ float tile[32][ALS_TPB + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n;
    unsigned char* base = a.scratch + size_t(blockIdx.x) * a.per_cta;
    double* Q = reinterpret_cast<double*>(base);
    double* P = Q + size_t(n) * ALS_TPB;
    double* V = P + size_t(n) * ALS_TPB;
    float* Y = reinterpret_cast<float*>(V + size_t(n) * ALS_TPB);
    unsigned char* W = reinterpret_cast<unsigned char*>(Y + size_t(n) * ALS_TPB);
    const long long groups = (a.batch + ALS_TPB - 1) / ALS_TPB;
    const double lam = a.lam, pw = a.p, qw = 1.0 - a.p;

    for (long long grp = blockIdx.x; grp < groups; grp += gridDim.x) {
        const long long s0 = grp * ALS_TPB;
        const int nrows = int((a.batch - s0) < ALS_TPB ? (a.batch - s0) : ALS_TPB);
        // ---- spectra (real part) -> Y[point][thread] --------------------------------------------------------------------
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int pt = i0 + lane;
#pragma unroll 4
            for (int r = 0; r < 32; ++r) {
                const int row = warp * 32 + r;
                float v = 0.f;
                if (row < nrows && pt < n) {
                    const size_t idx = size_t(s0 + row) * n + pt;
                    v = CPLX ? static_cast<const float2*>(a.in)[idx].x : static_cast<const float*>(a.in)[idx];
                }
                tile[lane][row] = v;
            }
            __syncthreads();
#pragma unroll 4
            for (int k = 0; k < 32; ++k)
                if (i0 + k < n) Y[size_t(i0 + k) * ALS_TPB + tid] = tile[k][tid];
            __syncthreads();
        }
        // ---- n_iter x (banded LDL^T + forward substitution, backward substitution + re-weighting) ------------------------
        if (tid < nrows) {
            for (int it = 0; it < a.n_iter; ++it) {
                const bool last = (it == a.n_iter - 1);
                double d1 = 0.0, d2 = 0.0, l1 = 0.0, l2 = 0.0, l2n = 0.0, u1 = 0.0, u2 = 0.0;   // l1 = l1_i, l2 = l2_i, l2n = l2_{i+1}
                // Both sweeps are latency chains per thread; their loads do not depend on the chain, so they are issued
                // ALS_U points ahead of it (the kernel is bound by HBM latency x bytes in flight).
                for (int ib = 0; ib < n; ib += ALS_U) {
                    float yv[ALS_U];
                    unsigned char cv[ALS_U];
#pragma unroll
                    for (int k = 0; k < ALS_U; ++k) {
                        const int i = ib + k;
                        const size_t o = size_t(i < n ? i : n - 1) * ALS_TPB + tid;
                        yv[k] = Y[o];
                        cv[k] = it > 0 ? W[o] : (unsigned char)3;
                    }
#pragma unroll
                    for (int k = 0; k < ALS_U; ++k) {
                        const int i = ib + k;
                        if (i < n) {
                            const size_t o = size_t(i) * ALS_TPB + tid;
                            const double y = double(yv[k]);
                            const unsigned char c = cv[k];
                            const double w = c == 3 ? 1.0 : (c == 1 ? pw : (c == 2 ? qw : 0.0));   // baseline.py:24: first solve unweighted
                            // bands of D'D (D = second differences, (n-2) x n): rows k = i, i-1, i-2 of D touch column i
                            const int v0 = (i <= n - 3), v1 = (i >= 1 && i <= n - 2), v2 = (i >= 2);
                            const double dg = double(v0 + 4 * v1 + v2);
                            const double o1 = (i <= n - 2) ? -2.0 * double(v0 + v1) : 0.0;      // (i, i+1)
                            const double o2 = v0 ? 1.0 : 0.0;                                    // (i, i+2)
                            const double d = (w + lam * dg) - l1 * l1 * d1 - l2 * l2 * d2;
                            const double u = w * y - l1 * u1 - l2 * u2;
                            const double dinv = 1.0 / d;
                            const double q = (lam * o1 - l2n * l1 * d1) * dinv;                  // l1_{i+1}
                            const double pp = (lam * o2) * dinv;                                 // l2_{i+2}
                            Q[o] = q;
                            P[o] = pp;
                            V[o] = u * dinv;
                            d2 = d1; d1 = d; u2 = u1; u1 = u;
                            l2 = l2n; l1 = q; l2n = pp;
                        }
                    }
                }
                double z1 = 0.0, z2 = 0.0;
                for (int ib = n - 1; ib >= 0; ib -= ALS_U) {
                    double vv[ALS_U], qv[ALS_U], pv[ALS_U];
                    float yv[ALS_U];
#pragma unroll
                    for (int k = 0; k < ALS_U; ++k) {
                        const int i = ib - k;
                        const size_t o = size_t(i >= 0 ? i : 0) * ALS_TPB + tid;
                        vv[k] = V[o];
                        qv[k] = Q[o];
                        pv[k] = P[o];
                        yv[k] = Y[o];
                    }
#pragma unroll
                    for (int k = 0; k < ALS_U; ++k) {
                        const int i = ib - k;
                        if (i >= 0) {
                            const size_t o = size_t(i) * ALS_TPB + tid;
                            const double z = vv[k] - qv[k] * z1 - pv[k] * z2;
                            const double y = double(yv[k]);
                            if (last) Y[o] = float(y - z);                                       // baseline.py:99: corrected = real - baseline
                            else W[o] = y > z ? 1 : (y < z ? 2 : 0);                             // baseline.py:37
                            z2 = z1;
                            z1 = z;
                        }
                    }
                }
            }
        }
        __syncthreads();
        // ---- corrected spectra back to [spectrum][point] -------------------------------------------------------------------
        for (int i0 = 0; i0 < n; i0 += 32) {
#pragma unroll 4
            for (int k = 0; k < 32; ++k)
                tile[k][tid] = (i0 + k < n) ? Y[size_t(i0 + k) * ALS_TPB + tid] : 0.f;
            __syncthreads();
            const int pt = i0 + lane;
#pragma unroll 4
            for (int r = 0; r < 32; ++r) {
                const int row = warp * 32 + r;
                if (row < nrows && pt < n) a.out[size_t(s0 + row) * n + pt] = tile[lane][row];
            }
            __syncthreads();
        }
    }
}

}  // namespace

extern "C" {

int64_t xmr_baseline_als_workspace_bytes(int64_t batch, int n) {
    if (batch <= 0 || n < 1) return 0;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    else cudaGetLastError();
    const int64_t groups = (batch + ALS_TPB - 1) / ALS_TPB;
    const int64_t ctas = groups < int64_t(sms) * ALS_MAX_CTAS_PER_SM ? groups : int64_t(sms) * ALS_MAX_CTAS_PER_SM;
    return ctas * int64_t(als_per_cta(n));
}

int xmr_baseline_als(const void* in_dev, int in_is_complex, float* out_dev, int64_t batch, int n, double lam, double p,
                     int n_iter, void* workspace_dev, int64_t workspace_bytes, void* stream) {
    if (batch < 0 || n < 3) return xmr_abi::fail(XMR_ERR_BAD_ARG, "baseline_als: batch=%lld, n=%d (needs n >= 3)", (long long)batch, n);
    if (n_iter < 1) return xmr_abi::fail(XMR_ERR_BAD_ARG, "baseline_als: n_iter=%d (needs at least one solve)", n_iter);
    if (!(lam >= 0.0)) return xmr_abi::fail(XMR_ERR_BAD_ARG, "baseline_als: lam=%g", lam);
    if (batch == 0) return XMR_OK;
    if (!in_dev || !out_dev || !workspace_dev) return xmr_abi::fail(XMR_ERR_BAD_ARG, "NULL pointer");
    const size_t per_cta = als_per_cta(n);
    const int64_t groups = (batch + ALS_TPB - 1) / ALS_TPB;
    int64_t ctas = workspace_bytes / int64_t(per_cta);
    if (ctas < 1) return xmr_abi::fail(XMR_ERR_BAD_ARG, "baseline_als: workspace of %lld bytes is smaller than one CTA's %lld",
                                        (long long)workspace_bytes, (long long)per_cta);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (ctas > int64_t(sms) * ALS_MAX_CTAS_PER_SM) ctas = int64_t(sms) * ALS_MAX_CTAS_PER_SM;
    if (ctas > groups) ctas = groups;
    AlsParams a;
    a.in = in_dev;
    a.out = out_dev;
    a.batch = batch;
    a.n = n;
    a.lam = lam;
    a.p = p;
    a.n_iter = n_iter;
    a.scratch = static_cast<unsigned char*>(workspace_dev);
    a.per_cta = per_cta;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (in_is_complex) als_kernel<true><<<unsigned(ctas), ALS_TPB, 0, st>>>(a);
    else als_kernel<false><<<unsigned(ctas), ALS_TPB, 0, st>>>(a);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? XMR_OK : xmr_abi::cuda_fail(e, "baseline_als launch");
}

}  // extern "C"
