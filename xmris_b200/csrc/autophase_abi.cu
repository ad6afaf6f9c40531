// C ABI: autophase(mode="single") building blocks -- per-row |S| statistics and the one-spectrum (p0, p1) search.
#include <cstdarg>
#include <cstdio>

#include "../../include/xmris_b200.h"
#include "abi_common.h"
#include "autophase_search.cuh"

namespace {

using namespace xmr;

// one warp per row: max |S| and its first index (numpy argmax semantics, phasing.py:229-231)
__global__ void row_absmax_kernel(const float2* __restrict__ spec, long long batch, int n, float* __restrict__ absmax,
                                  int* __restrict__ argmax) {
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < batch; row += warps) {
        const float2* r = spec + row * n;
        float best = -1.f;
        int besti = 0x7fffffff;
        for (int m = lane; m < n; m += 32) {
            const float2 x = r[m];
            const float v = x.x * x.x + x.y * x.y;
            if (v > best) { best = v; besti = m; }
        }
        for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, besti, off);
            if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
        }
        if (lane == 0) { absmax[row] = sqrtf(best); argmax[row] = besti; }
    }
}

// ---- the objective at given (p0, p1): one CTA per candidate, the 8 warps split the spectrum ---------------------------------
// R = double: what the last search levels use; R = float: what the coarse levels use.  Partial sums combine in double.
template <int METHOD, typename R>
__global__ void __launch_bounds__(SEARCH_THREADS) score_kernel(const float2* __restrict__ spec, int n, double u0, double du,
                                                               ScoreGeom geom, const double* __restrict__ p0,
                                                               const double* __restrict__ p1, double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sp = reinterpret_cast<float2*>(smem_raw);
    __shared__ double part[SEARCH_THREADS / 32][4];
    constexpr int NW = SEARCH_THREADS / 32;
    const int per_warp = (n + NW - 1) / NW;
    const int padshift = ilog2_ceil((per_warp + 31) / 32);
    load_padded(sp, spec, n, padshift);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    R c0[1], s0[1];
    RealOps<R>::sincospi2(R(p0[blockIdx.x] / 360.0), &s0[0], &c0[0]);
    const int w0 = min(warp * per_warp, n), w1 = min(w0 + per_warp, n);
    const int m0 = min(w0 + (lane << padshift), w1), m1 = min(m0 + (1 << padshift), w1);
    Acc<R, METHOD, 1> acc;
    acc.init();
    lane_accumulate_rt<R, METHOD, 1>(sp, padshift, m0, m1, geom, R(p1[blockIdx.x] / 360.0), R(u0), R(du), c0, s0, acc);
    acc.warp_reduce();
    if (lane == 0)
        for (int s = 0; s < 4; ++s) part[warp][s] = double(acc.a[0][s]);
    __syncthreads();
    if (threadIdx.x == 0) {
        Acc<double, METHOD, 1> tot;
        for (int s = 0; s < 4; ++s) {
            double v = part[0][s];
            for (int w = 1; w < NW; ++w) v = Acc<double, METHOD, 1>::comb(s, v, part[w][s]);
            tot.a[0][s] = v;
        }
        out[blockIdx.x] = tot.template score<true>(0, geom);
    }
}

template <int METHOD>
int run_score(const float2* spec, int n, double u0, double du, ScoreGeom geom, const double* p0, const double* p1, int k,
              int use_f64, double* out, cudaStream_t st) {
    const int per_warp = (n + 7) / 8, Lz = (per_warp + 31) / 32;
    int ps = 0;
    while ((1 << ps) < Lz) ++ps;
    const size_t smem = sizeof(float2) * (size_t(n) + (size_t(n) >> ps) + 2);
    cudaError_t e;
    if (use_f64) {
        e = cudaFuncSetAttribute(score_kernel<METHOD, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return xmr_abi::cuda_fail(e, "cudaFuncSetAttribute(score f64)");
        score_kernel<METHOD, double><<<k, SEARCH_THREADS, smem, st>>>(spec, n, u0, du, geom, p0, p1, out);
    } else {
        e = cudaFuncSetAttribute(score_kernel<METHOD, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return xmr_abi::cuda_fail(e, "cudaFuncSetAttribute(score f32)");
        score_kernel<METHOD, float><<<k, SEARCH_THREADS, smem, st>>>(spec, n, u0, du, geom, p0, p1, out);
    }
    e = cudaGetLastError();
    return e == cudaSuccess ? XMR_OK : xmr_abi::cuda_fail(e, "score launch");
}

constexpr int WS_LIST = 8192;   // candidates per ping-pong list

// Search geometry (xmr_autophase_search_tuning): coarse grid steps in degrees, number of distinct coarse cells refined
// side by side, nested zoom levels (each shrinks the window by 5).
struct SearchTuning {
    double p0_step = 6.0, p1_step = 15.0;
    int starts = 4, levels = 4;
    int f32_levels = 2;      // leading zoom levels evaluated in float32 (all `starts` basins)
    int late_starts = 2;     // basins kept for the float64 levels
    double first_ratio = 2.5;   // window shrink factor from zoom level 0 to level 1 (5 between all later levels)
    int polish_starts = 3;      // ACME: basins finished by the float64 Newton polish after the float32 levels (0: float64 zoom levels)
    int fine_b = 1;             // ACME: second fine float64 level (+-0.06 x +-0.18 deg, spacing 0.006 x 0.018) after level A
    int polish_f32_levels = 1;  // ACME: float32 zoom levels before the polish (its trust region covers the second one)
};
SearchTuning g_tuning;

// pivot_dev (optional): the pivot index lives in device memory (u0 = -du*idx, ROI around idx): the geometry is resolved on the
// device (search_geom_kernel) and `u0` / `geom.target_idx` of the arguments are ignored.  ph_out (optional): K1PhaseDev for pass 2.
template <int METHOD>
int run_search(const float2* spec, int n, double u0, double du, ScoreGeom geom, int p0_only, double* result, Cand* ws,
               cudaStream_t st, const int* pivot_dev = nullptr, int index_width = 1, void* ph_out = nullptr) {
    SearchGeomDev* gd = nullptr;
    if (pivot_dev != nullptr) {
        gd = reinterpret_cast<SearchGeomDev*>(ws + 2 * WS_LIST);
        search_geom_kernel<<<1, 32, 0, st>>>(pivot_dev, du, n, index_width, gd);
    }
    const int Lc = (n + 31) / 32;
    int psc = 0;
    while ((1 << psc) < Lc) ++psc;
    const size_t smem_coarse = sizeof(float2) * (size_t(n) + (size_t(n) >> psc) + 2);
    const int per_warp = (n + 7) / 8, Lz = (per_warp + 31) / 32;
    int psz = 0;
    while ((1 << psz) < Lz) ++psz;
    const size_t smem_zoom = sizeof(float2) * (size_t(n) + (size_t(n) >> psz) + 2);
    cudaError_t e;
    e = cudaFuncSetAttribute(search_coarse_kernel<METHOD>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_coarse));
    if (e != cudaSuccess) return xmr_abi::cuda_fail(e, "cudaFuncSetAttribute(search_coarse)");
    e = cudaFuncSetAttribute(search_zoom_kernel<METHOD, double, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_zoom));
    if (e != cudaSuccess) return xmr_abi::cuda_fail(e, "cudaFuncSetAttribute(search_zoom)");
    e = cudaFuncSetAttribute(search_zoom_kernel<METHOD, float, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_zoom));
    if (e != cudaSuccess) return xmr_abi::cuda_fail(e, "cudaFuncSetAttribute(search_zoom f32)");

    Cand* listA = ws;
    Cand* listB = ws + WS_LIST;

    SearchParams sp;
    sp.spec = spec;
    sp.n = n;
    sp.u0 = u0;
    sp.du = du;
    sp.geom = geom;
    sp.gd = gd;
    sp.p0_lo = -180.0;
    sp.p0_hi = 180.0;
    sp.p1_lo = p0_only ? 0.0 : -4000.0;
    sp.p1_hi = p0_only ? 0.0 : 4000.0;
    sp.p0_step = p0_only ? 0.05 : g_tuning.p0_step;
    sp.p1_step = g_tuning.p1_step;
    sp.n_p0 = int((sp.p0_hi - sp.p0_lo) / sp.p0_step + 0.5) + 1;
    sp.n_p1 = p0_only ? 1 : int((sp.p1_hi - sp.p1_lo) / sp.p1_step + 0.5) + 1;
    sp.out = listA;
    const long long items = (long long)sp.n_p1 * ((sp.n_p0 + SEARCH_K - 1) / SEARCH_K);
    int sms = 148;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long grid = (items + 7) / 8;
    const long long cap = (long long)sms * 2;
    if (grid > cap) grid = cap;
    if (grid * (SEARCH_THREADS / 32) > WS_LIST) grid = WS_LIST / (SEARCH_THREADS / 32);
    search_coarse_kernel<METHOD><<<int(grid), SEARCH_THREADS, smem_coarse, st>>>(sp);
    e = cudaGetLastError();
    if (e != cudaSuccess) return xmr_abi::cuda_fail(e, "search_coarse launch");

    ZoomParams zp;
    zp.spec = spec;
    zp.n = n;
    zp.u0 = u0;
    zp.du = du;
    zp.geom = geom;
    zp.gd = gd;
    zp.only_flagged = 0;
    zp.p0_lo = sp.p0_lo;
    zp.p0_hi = sp.p0_hi;
    zp.p1_lo = sp.p1_lo;
    zp.p1_hi = sp.p1_hi;
    zp.rows = p0_only ? 1 : ZOOM_SIDE;
    const Cand* prev = listA;
    int n_prev = int(grid) * (SEARCH_THREADS / 32);      // the coarse list: one best cell per warp
    Cand* cur = listB;
    double h0 = sp.p0_step, h1 = p0_only ? 0.0 : sp.p1_step;
    zp.sep0 = 2.0 * sp.p0_step;
    zp.sep1 = p0_only ? 0.0 : 2.0 * sp.p1_step;
    // ACME: the float32 zoom levels localise the basins, the float64 Newton polish (search_polish_kernel) finishes them;
    // the ROI methods (piecewise-constant minima) keep the float64 zoom levels
    const bool polish = (METHOD == METHOD_ACME) && g_tuning.polish_starts > 0;
    const int levels = polish ? g_tuning.polish_f32_levels : g_tuning.levels;
    int starts_prev = 0;
    double ratio_prev = 1.0;
    for (int lvl = 0; lvl < levels; ++lvl) {
        const bool f32 = lvl < g_tuning.f32_levels;
        const int n_starts = f32 ? g_tuning.starts : (g_tuning.late_starts < g_tuning.starts ? g_tuning.late_starts : g_tuning.starts);
        zp.n_starts = n_starts;
        zp.prev = prev;
        zp.cur = cur;
        zp.h0 = h0;
        zp.h1 = h1;
        if (lvl == 0) {
            zp.first_level = 1;
            zp.n_prev = n_prev;
        } else if (n_starts != starts_prev) {
            // fewer basins from here on: the best mutually distinct candidates over ALL blocks of the previous level
            // (two candidates further apart than that level's window belong to different basins)
            zp.first_level = 1;
            zp.n_prev = n_prev * starts_prev;
            zp.sep0 = 2.0 * h0 * ratio_prev;
            zp.sep1 = 2.0 * h1 * ratio_prev;
        } else {
            zp.first_level = 0;
            zp.n_prev = n_prev;
        }
        if (f32) search_zoom_kernel<METHOD, float, 8><<<zp.rows * (ZOOM_SPAN / 8) * n_starts, SEARCH_THREADS, smem_zoom, st>>>(zp);
        else search_zoom_kernel<METHOD, double, 4><<<zp.rows * (ZOOM_SPAN / 4) * n_starts, SEARCH_THREADS, smem_zoom, st>>>(zp);
        e = cudaGetLastError();
        if (e != cudaSuccess) return xmr_abi::cuda_fail(e, "search_zoom launch");
        n_prev = zp.rows * ZOOM_SPAN;                  // candidates per start
        starts_prev = n_starts;
        prev = cur;
        cur = (cur == listB) ? listA : listB;
        ratio_prev = (lvl == 0) ? g_tuning.first_ratio : 5.0;
        h0 /= ratio_prev;
        h1 /= ratio_prev;
    }
    int n_starts = starts_prev;
    if (polish) {
        PolishParams pp;
        pp.spec = spec;
        pp.n = n;
        pp.u0 = u0;
        pp.gd = gd;
        pp.du = du;
        pp.prev = prev;
        pp.n_prev = (levels == 0) ? n_prev : n_prev * n_starts;    // no zoom level ran: the coarse grid's list
        pp.sep0 = 2.0 * h0 * ratio_prev;       // further apart than the last level's window: another basin
        pp.sep1 = p0_only ? 0.0 : 2.0 * h1 * ratio_prev;
        pp.p1_lo = sp.p1_lo;
        pp.p1_hi = sp.p1_hi;
        pp.p0_only = p0_only;
        pp.out = cur;
        const int chunk = (n + SEARCH_THREADS - 1) / SEARCH_THREADS;
        int psp = 0;
        while ((1 << psp) < chunk) ++psp;
        const size_t smem_polish = sizeof(float2) * (size_t(n) + (size_t(n) >> psp) + 2);
        e = cudaFuncSetAttribute(search_polish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_polish));
        if (e != cudaSuccess) return xmr_abi::cuda_fail(e, "cudaFuncSetAttribute(search_polish)");
        search_polish_kernel<<<g_tuning.polish_starts * POLISH_ROLES, SEARCH_THREADS, smem_polish, st>>>(pp);
        e = cudaGetLastError();
        if (e != cudaSuccess) return xmr_abi::cuda_fail(e, "search_polish launch");
        // Two fine float64 zoom levels finish the job by direct search.  Where the penalty term vanishes at the optimum (clean,
        // all-positive spectra; data of amplitude ~1e9 like the reference's Bruker fixture) the minimum sits against the wall
        // 1000 P > 0 of the feasible region -- a constrained optimum where the gradient does not vanish -- and where the
        // entropy term dominates the objective is rough below ~0.3 deg (one kink per spectral point).
        //   A: +-0.3 x +-0.9 deg (spacing 0.03 x 0.09) around every polished point;  B: +-0.06 x +-0.18 deg around the best
        zp.first_level = 0;
        zp.prev = cur;
        zp.n_prev = 1;
        zp.n_starts = g_tuning.polish_starts * POLISH_ROLES;      // (every role's result is a start of level A)
        zp.only_flagged = 1;                                       // smooth optima: the polished point stands
        zp.h0 = 0.3;
        zp.h1 = p0_only ? 0.0 : 0.9;
        Cand* fin = (cur == listB) ? listA : listB;
        zp.cur = fin;
        search_zoom_kernel<METHOD, double, 4><<<zp.rows * (ZOOM_SPAN / 4) * zp.n_starts, SEARCH_THREADS, smem_zoom, st>>>(zp);
        e = cudaGetLastError();
        if (e != cudaSuccess) return xmr_abi::cuda_fail(e, "search_zoom (fine A) launch");
        if (g_tuning.fine_b) {
            zp.first_level = 1;                           // the single best candidate over all starts of level A
            zp.prev = fin;
            zp.n_prev = zp.rows * ZOOM_SPAN * g_tuning.polish_starts * POLISH_ROLES;
            zp.n_starts = 1;
            zp.h0 = 0.06;
            zp.h1 = p0_only ? 0.0 : 0.18;
            zp.cur = cur;
            zp.only_flagged = 1;                          // level B only where the optimum sits against the wall
            search_zoom_kernel<METHOD, double, 4><<<zp.rows * (ZOOM_SPAN / 4), SEARCH_THREADS, smem_zoom, st>>>(zp);
            e = cudaGetLastError();
            if (e != cudaSuccess) return xmr_abi::cuda_fail(e, "search_zoom (fine B) launch");
            // the final pick sees level B (which contains A's best point: its window centre)
            prev = cur;
            n_prev = zp.rows * ZOOM_SPAN;
            n_starts = 1;
        } else {
            prev = fin;
            n_prev = zp.rows * ZOOM_SPAN;
            n_starts = g_tuning.polish_starts * POLISH_ROLES;
        }
    }
    FinalizePhase fp;
    fp.ph_out = ph_out;
    fp.gd = gd;
    fp.u0 = u0;
    fp.du = du;
    fp.n = n;
    fp.p0_only = p0_only;
    search_finalize_kernel<<<1, SEARCH_THREADS, 0, st>>>(prev, n_prev * n_starts, result, fp);
    e = cudaGetLastError();
    if (e != cudaSuccess) return xmr_abi::cuda_fail(e, "search_finalize launch");
    return XMR_OK;
}

}  // namespace

extern "C" {

int xmr_row_absmax_c64(const void* spec_dev, int64_t batch, int n, float* absmax_dev, int* argmax_dev, void* stream) {
    if (batch < 0 || n < 1) return xmr_abi::fail(XMR_ERR_BAD_ARG, "bad sizes");
    if (batch == 0) return XMR_OK;
    if (!spec_dev || !absmax_dev || !argmax_dev) return xmr_abi::fail(XMR_ERR_BAD_ARG, "NULL pointer");
    const long long blocks = (batch + 7) / 8;
    const int grid = int(blocks < 148LL * 8 ? blocks : 148LL * 8);
    row_absmax_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const float2*>(spec_dev), batch, n,
                                                                         absmax_dev, argmax_dev);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? XMR_OK : xmr_abi::cuda_fail(e, "row_absmax launch");
}

int xmr_autophase_search_tuning(double p0_step_deg, double p1_step_deg, int starts, int levels, int f32_levels,
                                int late_starts, double first_ratio) {
    if (!(p0_step_deg > 0.0) || !(p1_step_deg > 0.0) || p0_step_deg > 45.0 || p1_step_deg > 500.0 || starts < 1 ||
        starts > ZOOM_MAX_STARTS || levels < 1 || levels > 12 || starts * ZOOM_SIDE * ZOOM_SPAN > WS_LIST ||
        f32_levels < 0 || late_starts < 1 || !(first_ratio >= 1.0) || first_ratio > 10.0)
        return xmr_abi::fail(XMR_ERR_BAD_ARG, "search tuning: p0_step=%g p1_step=%g starts=%d levels=%d f32_levels=%d "
                             "late_starts=%d", p0_step_deg, p1_step_deg, starts, levels, f32_levels, late_starts);
    g_tuning.f32_levels = f32_levels;
    g_tuning.first_ratio = first_ratio;
    g_tuning.late_starts = late_starts;
    g_tuning.p0_step = p0_step_deg;
    g_tuning.p1_step = p1_step_deg;
    g_tuning.starts = starts;
    g_tuning.levels = levels;
    return XMR_OK;
}

int xmr_autophase_search_polish(int starts, int f32_levels, int fine_b) {
    if (starts < 0 || starts * POLISH_ROLES > ZOOM_MAX_STARTS + 1 || f32_levels < 0 || f32_levels > 4)
        return xmr_abi::fail(XMR_ERR_BAD_ARG, "polish starts=%d must lie in [0, %d], f32_levels=%d in [0, 4]", starts,
                             (ZOOM_MAX_STARTS + 1) / POLISH_ROLES, f32_levels);
    g_tuning.polish_starts = starts;
    g_tuning.polish_f32_levels = f32_levels;
    g_tuning.fine_b = fine_b ? 1 : 0;
    return XMR_OK;
}

int64_t xmr_autophase_workspace_bytes(void) { return int64_t(sizeof(Cand)) * 2 * WS_LIST + 256; }

int xmr_autophase_score_c64(const void* spec_dev, int n, double u0, double du, int method, int target_idx, int index_width,
                            const double* p0_dev, const double* p1_dev, int k, int use_f64, double* out_dev, void* stream) {
    if (n < 2 || n > 8192) return xmr_abi::fail(XMR_ERR_UNSUPPORTED_N, "autophase score: n=%d must lie in [2, 8192]", n);
    if (k < 0) return xmr_abi::fail(XMR_ERR_BAD_ARG, "k=%d", k);
    if (k == 0) return XMR_OK;
    if (!spec_dev || !p0_dev || !p1_dev || !out_dev) return xmr_abi::fail(XMR_ERR_BAD_ARG, "NULL pointer");
    if (target_idx < 0 || target_idx >= n || index_width < 1)
        return xmr_abi::fail(XMR_ERR_BAD_ARG, "bad target_idx=%d / index_width=%d", target_idx, index_width);
    ScoreGeom g;
    g.n = n;
    g.target_idx = target_idx;
    g.roi_start = target_idx - index_width > 0 ? target_idx - index_width : 0;
    g.roi_end = target_idx + index_width < n ? target_idx + index_width : n;
    const float2* s = static_cast<const float2*>(spec_dev);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (method) {
        case XMR_METHOD_ACME: return run_score<METHOD_ACME>(s, n, u0, du, g, p0_dev, p1_dev, k, use_f64, out_dev, st);
        case XMR_METHOD_PEAK_MINIMA: return run_score<METHOD_PEAK_MINIMA>(s, n, u0, du, g, p0_dev, p1_dev, k, use_f64, out_dev, st);
        case XMR_METHOD_POSITIVITY: return run_score<METHOD_POSITIVITY>(s, n, u0, du, g, p0_dev, p1_dev, k, use_f64, out_dev, st);
        default: return xmr_abi::fail(XMR_ERR_BAD_ARG, "method=%d", method);
    }
}

}  // extern "C"

namespace xmr_abi {
// the search with its geometry resolved on the device (pivot index in device memory) and the phase parameters of pass 2
// written to device memory: the mode="single" chain without host read-backs (host_chain.cu)
int autophase_search_dev(const void* spec_dev, int n, double du, int method, const int* pivot_dev, int fixed_pivot, double u0_fixed,
                         int fixed_target, int index_width, int p0_only, double* result_dev, void* workspace_dev, void* ph_out,
                         void* stream) {
    ScoreGeom g;
    g.n = n;
    g.target_idx = fixed_pivot ? fixed_target : 0;
    g.roi_start = g.target_idx - index_width > 0 ? g.target_idx - index_width : 0;
    g.roi_end = g.target_idx + index_width < n ? g.target_idx + index_width : n;
    const float2* s = static_cast<const float2*>(spec_dev);
    Cand* ws = static_cast<Cand*>(workspace_dev);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int* piv = fixed_pivot ? nullptr : pivot_dev;
    switch (method) {
        case XMR_METHOD_ACME: return run_search<METHOD_ACME>(s, n, u0_fixed, du, g, p0_only, result_dev, ws, st, piv, index_width, ph_out);
        case XMR_METHOD_PEAK_MINIMA: return run_search<METHOD_PEAK_MINIMA>(s, n, u0_fixed, du, g, p0_only, result_dev, ws, st, piv, index_width, ph_out);
        case XMR_METHOD_POSITIVITY: return run_search<METHOD_POSITIVITY>(s, n, u0_fixed, du, g, p0_only, result_dev, ws, st, piv, index_width, ph_out);
        default: return fail(XMR_ERR_BAD_ARG, "method=%d", method);
    }
}
}  // namespace xmr_abi

extern "C" {

int xmr_autophase_search_c64(const void* spec_dev, int n, double u0, double du, int method, int target_idx,
                             int index_width, int p0_only, double* result_dev, void* workspace_dev, void* stream) {
    if (n < 2 || n > 8192) return xmr_abi::fail(XMR_ERR_UNSUPPORTED_N, "autophase search: n=%d must lie in [2, 8192]", n);
    if (!spec_dev || !result_dev || !workspace_dev) return xmr_abi::fail(XMR_ERR_BAD_ARG, "NULL pointer");
    if (target_idx < 0 || target_idx >= n || index_width < 1)
        return xmr_abi::fail(XMR_ERR_BAD_ARG, "bad target_idx=%d / index_width=%d", target_idx, index_width);
    ScoreGeom g;
    g.n = n;
    g.target_idx = target_idx;
    g.roi_start = target_idx - index_width > 0 ? target_idx - index_width : 0;
    g.roi_end = target_idx + index_width < n ? target_idx + index_width : n;
    const float2* s = static_cast<const float2*>(spec_dev);
    Cand* ws = static_cast<Cand*>(workspace_dev);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (method) {
        case XMR_METHOD_ACME: return run_search<METHOD_ACME>(s, n, u0, du, g, p0_only, result_dev, ws, st);
        case XMR_METHOD_PEAK_MINIMA: return run_search<METHOD_PEAK_MINIMA>(s, n, u0, du, g, p0_only, result_dev, ws, st);
        case XMR_METHOD_POSITIVITY: return run_search<METHOD_POSITIVITY>(s, n, u0, du, g, p0_only, result_dev, ws, st);
        default: return xmr_abi::fail(XMR_ERR_BAD_ARG, "method=%d", method);
    }
}

}  // extern "C"
