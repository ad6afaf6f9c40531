// Global (p0, p1) search on ONE spectrum -- the reference's autophase(mode="single") optimisation step
// (src/xmris/processing/phasing.py:270-287: differential evolution over p0 in [-180,180], p1 in [-4000,4000],
// polished by L-BFGS-B).  The GPU replaces the stochastic population search by a deterministic dense grid
// (float32, every SM) followed by nested 21x21 zoom refinements in float64 around the best cell, honouring the
// same closed box by clamping.  SURVEY.md Appendix C / tests/test_autophase_gpu.py: this lands on the reference's
// optimum to a few 1e-3 degrees.
#pragma once
#include "acme_grad.cuh"
#include "autophase_eval.cuh"

namespace xmr {

struct Cand {
    double f, p0, p1, pad;
};

// Search geometry in DEVICE memory (the pivot is the winning spectrum's own |S| argmax, known only on the device when the
// chain runs without host read-backs): written by search_geom_kernel, read by every search kernel when `gd` is set.
struct SearchGeomDev {
    double u0;
    int target_idx, roi_start, roi_end;
};
__global__ void search_geom_kernel(const int* pivot_idx, double du, int n, int index_width, SearchGeomDev* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const int idx = *pivot_idx;
        out->u0 = -du * double(idx);
        out->target_idx = idx;
        out->roi_start = idx - index_width > 0 ? idx - index_width : 0;
        out->roi_end = idx + index_width < n ? idx + index_width : n;
    }
}
__device__ __forceinline__ void resolve_geom(const SearchGeomDev* gd, double& u0, ScoreGeom& g) {
    if (gd != nullptr) {
        u0 = gd->u0;
        g.target_idx = gd->target_idx;
        g.roi_start = gd->roi_start;
        g.roi_end = gd->roi_end;
    }
}

struct SearchParams {
    const float2* spec;   // one spectrum, n points (global memory)
    int n;
    double u0, du;        // u_m = u0 + du*m = (x_m - pivot)/(x_max - x_min)
    ScoreGeom geom;
    const SearchGeomDev* gd;   // optional: u0 / ROI from device memory
    double p0_lo, p0_hi, p0_step;
    int n_p0;
    double p1_lo, p1_hi, p1_step;
    int n_p1;
    Cand* out;
};

constexpr int SEARCH_K = 8;          // zero-order candidates evaluated per pass over the spectrum
constexpr int SEARCH_THREADS = 256;

__device__ __forceinline__ int ilog2_ceil(int v) {
    int s = 0;
    while ((1 << s) < v) ++s;
    return s;
}

// cooperative copy of the spectrum into the padded shared layout idx(m) = m + (m >> padshift)
__device__ __forceinline__ void load_padded(float2* sp, const float2* __restrict__ g, int n, int padshift) {
    for (int m = threadIdx.x; m < n; m += blockDim.x) sp[m + (m >> padshift)] = g[m];
}



// ---- coarse grid: one warp per (p1, chunk of K p0 values), float32 -----------------------------------------
template <int METHOD>
__global__ void __launch_bounds__(SEARCH_THREADS) search_coarse_kernel(const __grid_constant__ SearchParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sp = reinterpret_cast<float2*>(smem_raw);
    const int n = p.n;
    const int L = (n + 31) / 32;
    const int padshift = ilog2_ceil(L);
    load_padded(sp, p.spec, n, padshift);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double u0g = p.u0;
    ScoreGeom geom = p.geom;
    resolve_geom(p.gd, u0g, geom);
    const int nchunks = (p.n_p0 + SEARCH_K - 1) / SEARCH_K;
    const long long items = (long long)p.n_p1 * nchunks;
    const int m0 = min(lane << padshift, n), m1 = min((lane + 1) << padshift, n);

    double best_f = CUDART_INF, best_p0 = p.p0_lo, best_p1 = p.p1_lo;
    for (long long item = (long long)blockIdx.x * (SEARCH_THREADS / 32) + warp; item < items;
         item += (long long)gridDim.x * (SEARCH_THREADS / 32)) {
        const int i1 = int(item / nchunks), ch = int(item - (long long)i1 * nchunks);
        const double p1 = fmin(p.p1_lo + i1 * p.p1_step, p.p1_hi);
        float c0[SEARCH_K], s0[SEARCH_K];
        double p0k[SEARCH_K];
#pragma unroll
        for (int k = 0; k < SEARCH_K; ++k) {
            p0k[k] = fmin(p.p0_lo + (ch * SEARCH_K + k) * p.p0_step, p.p0_hi);
            sincospif(float(p0k[k] / 180.0), &s0[k], &c0[k]);
        }
        Acc<float, METHOD, SEARCH_K> acc;
        acc.init();
        lane_accumulate_rt<float, METHOD, SEARCH_K>(sp, padshift, m0, m1, geom, float(p1 / 360.0), float(u0g),
                                                    float(p.du), c0, s0, acc);
        acc.warp_reduce();
#pragma unroll
        for (int k = 0; k < SEARCH_K; ++k) {
            const float f = acc.score(k, geom);
            if (double(f) < best_f) { best_f = double(f); best_p0 = p0k[k]; best_p1 = p1; }
        }
    }
    // every warp reports its own best cell: the zoom stage picks several mutually distinct starts from this list
    if (lane == 0) {
        Cand b;
        b.f = best_f; b.p0 = best_p0; b.p1 = best_p1; b.pad = 0;
        p.out[blockIdx.x * (SEARCH_THREADS / 32) + warp] = b;
    }
}

struct ZoomParams {
    const float2* spec;
    int n;
    double u0, du;
    ScoreGeom geom;
    const Cand* prev;   // candidates of the previous level (or the per-CTA bests of the coarse grid)
    int n_prev;
    double h0, h1;      // half-widths of this level's window around the best previous candidate
    double p0_lo, p0_hi, p1_lo, p1_hi;
    int rows;           // 21 (two parameters) or 1 (p0 only)
    Cand* cur;          // n_starts * rows * ZOOM_SPAN candidates (one contiguous block per start)
    int n_starts;       // basins refined in parallel (<= ZOOM_MAX_STARTS)
    int first_level;    // 1: prev is ONE list of n_prev candidates shared by all starts (start s takes the s-th best
                        //    DISTINCT one): the coarse grid's list, or all blocks of a level that ran more starts;
                        // 0: prev holds one block of n_prev candidates per start
    double sep0, sep1;  // first level: two cells are distinct when they differ by more than this in p0 or in p1
    const SearchGeomDev* gd;   // optional: u0 / ROI from device memory
    int only_flagged;   // 1: evaluate the window only if the centre carries the wall flag (Cand::pad != 0); otherwise the
                        //    centre itself is the only candidate reported
};

constexpr int ZOOM_MAX_STARTS = 8;  // the best distinct coarse cells are refined side by side (63 CTAs each)
constexpr int ZOOM_SIDE = 21;   // 21 x 21 points per level, spacing h/10
constexpr int ZOOM_SPAN = 24;   // p0 slots per row (>= ZOOM_SIDE), split into chunks of K candidates

__device__ __forceinline__ bool cand_distinct(const Cand& a, const Cand& b, double sep0, double sep1) {
    double d0 = fabs(a.p0 - b.p0);
    d0 = fmin(d0, 360.0 - d0);                       // p0 is periodic
    return d0 > sep0 || fabs(a.p1 - b.p1) > sep1;
}

// block-wide argmin of a candidate list (lowest index wins ties) -> broadcast.  Candidates that are not distinct from
// every one of the `n_excl` cells in `excl` are skipped.
__device__ __forceinline__ Cand block_argmin(const Cand* list, int n_list, const Cand* excl = nullptr, int n_excl = 0,
                                             double sep0 = 0.0, double sep1 = 0.0) {
    __shared__ double sf[SEARCH_THREADS / 32];
    __shared__ int si[SEARCH_THREADS / 32];
    __shared__ int s_best;
    double f = CUDART_INF;
    int idx = 0x7fffffff;
    for (int i = threadIdx.x; i < n_list; i += blockDim.x) {
        const double v = list[i].f;
        bool ok = true;
        for (int e = 0; e < n_excl; ++e) ok = ok && cand_distinct(list[i], excl[e], sep0, sep1);
        if (!ok) continue;
        if (v < f || (v == f && i < idx)) { f = v; idx = i; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double of = __shfl_xor_sync(0xffffffffu, f, off);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, off);
        if (of < f || (of == f && oi < idx)) { f = of; idx = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sf[threadIdx.x >> 5] = f; si[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < int(blockDim.x >> 5); ++w)
            if (sf[w] < f || (sf[w] == f && si[w] < idx)) { f = sf[w]; idx = si[w]; }
        s_best = (idx == 0x7fffffff || !(f < CUDART_INF)) ? -1 : idx;
    }
    __syncthreads();
    const int b = s_best;
    __syncthreads();                                  // s_best may be rewritten by a second call
    if (b < 0) {                                      // nothing (finite / distinct) found
        Cand none = list[0];
        none.f = CUDART_INF;
        return none;
    }
    return list[b];
}

// ---- zoom level: one CTA per (p1 row, chunk of K p0 values), the 8 warps split the spectrum -----------------------
// R = float for the wide early levels (spacing >= 0.1 deg: objective differences between neighbouring candidates are far
// above float32 summation noise), double for the final ones.  Per-warp partial sums are combined and scored in double.
// K: zero-order candidates per CTA (24/K chunks per p1 row): 8 for the float32 levels, 4 for the float64 ones, whose long
// dependent chains per candidate profit more from twice the CTAs than from sharing the first-order rotation.
template <int METHOD, typename R, int K>
__global__ void __launch_bounds__(SEARCH_THREADS) search_zoom_kernel(const __grid_constant__ ZoomParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sp = reinterpret_cast<float2*>(smem_raw);
    __shared__ double part[SEARCH_THREADS / 32][K][4];
    const int n = p.n;
    constexpr int NW = SEARCH_THREADS / 32;
    const int per_warp = (n + NW - 1) / NW;
    const int L = (per_warp + 31) / 32;
    const int padshift = ilog2_ceil(L);
    if (!p.only_flagged) load_padded(sp, p.spec, n, padshift);   // (flagged-only levels decide first: most CTAs only pass their centre on)
    const int per_start = p.rows * (ZOOM_SPAN / K);
    const int start = blockIdx.x / per_start, local = blockIdx.x % per_start;
    Cand centre;                                          // (block_argmin contains the barriers that also publish `sp`)
    if (p.first_level) {
        Cand chosen[ZOOM_MAX_STARTS];
        centre = block_argmin(p.prev, p.n_prev);
        for (int s = 1; s <= start; ++s) {                // the (start+1)-th best cell, distinct from all better starts
            chosen[s - 1] = centre;
            centre = block_argmin(p.prev, p.n_prev, chosen, s, p.sep0, p.sep1);
        }
    } else {
        centre = block_argmin(p.prev + (size_t)start * p.n_prev, p.n_prev);
    }
    const bool dead = !(centre.f < CUDART_INF);           // no second basin: this start reports +inf
    if (p.only_flagged && centre.pad == 0.0) {            // (CTA-uniform) smooth optimum: the polished point stands
        if (threadIdx.x < K) {
            Cand c = centre;
            if (!(local == 0 && threadIdx.x == 0)) c.f = CUDART_INF;
            p.cur[blockIdx.x * K + threadIdx.x] = c;
        }
        return;
    }
    if (p.only_flagged) {
        load_padded(sp, p.spec, n, padshift);
        __syncthreads();
    }

    const int row = local / (ZOOM_SPAN / K), ch = local % (ZOOM_SPAN / K);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double u0g = p.u0;
    ScoreGeom geom = p.geom;
    resolve_geom(p.gd, u0g, geom);
    const double p1 = (p.rows == 1) ? centre.p1
                                    : fmin(fmax(centre.p1 + (row - ZOOM_SIDE / 2) * (p.h1 / (ZOOM_SIDE / 2)), p.p1_lo), p.p1_hi);
    R c0[K], s0[K];
    double p0k[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int i = min(ch * K + k, ZOOM_SIDE - 1);
        // p0 is periodic: a window that leaves the closed box [-180, 180] re-enters on the other side
        double q0 = centre.p0 + (i - ZOOM_SIDE / 2) * (p.h0 / (ZOOM_SIDE / 2));
        q0 = q0 > p.p0_hi ? q0 - 360.0 : (q0 < p.p0_lo ? q0 + 360.0 : q0);
        p0k[k] = q0;
        RealOps<R>::sincospi2(R(p0k[k] / 360.0), &s0[k], &c0[k]);
    }
    const int w0 = min(warp * per_warp, n), w1 = min(w0 + per_warp, n);
    const int m0 = min(w0 + (lane << padshift), w1), m1 = min(m0 + (1 << padshift), w1);
    Acc<R, METHOD, K> acc;
    acc.init();
    lane_accumulate_rt<R, METHOD, K>(sp, padshift, m0, m1, geom, R(p1 / 360.0), R(u0g), R(p.du), c0, s0, acc);
    acc.warp_reduce();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int s = 0; s < 4; ++s) part[warp][k][s] = double(acc.a[k][s]);
    }
    __syncthreads();
    if (threadIdx.x < K) {
        const int k = threadIdx.x;
        Acc<double, METHOD, 1> tot;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            double v = part[0][k][s];
            for (int w = 1; w < NW; ++w) v = Acc<double, METHOD, 1>::comb(s, v, part[w][k][s]);
            tot.a[0][s] = v;
        }
        Cand c;
        c.f = tot.score(0, geom);
        if (!(c.f == c.f) || dead) c.f = CUDART_INF;   // NaN never wins
        // select p0k[k] without dynamic register indexing
        double myp0 = p0k[0];
#pragma unroll
        for (int kk = 1; kk < K; ++kk)
            if (kk == k) myp0 = p0k[kk];
        c.p0 = myp0;
        c.p1 = p1;
        c.pad = centre.pad;
        p.cur[blockIdx.x * K + k] = c;
    }
}

// ---- ACME polish: bounded Newton on the analytic gradient, float64, one CTA per start ---------------------------------------
// Replaces the two float64 zoom levels of round 1: from the best mutually distinct candidates of the last
// float32 zoom level, iterate x <- x - H^-1 g with the analytic gradient (acme_grad.cuh) and a secant Hessian until the step
// is below 1e-3 deg, then try the two neighbouring branches of the max(d) envelope ("kink hop").  Converges to the local
// minimum to ~1e-4 deg in 4-6 iterations -- the point the reference's own optimiser converges to when it is run with a tight
// tolerance (profiles/parity_r2.json).
struct PolishParams {
    const float2* spec;
    int n;
    double u0, du;
    const Cand* prev;     // all candidates of the previous level (one list)
    int n_prev;
    double sep0, sep1;    // two candidates further apart than this are different basins
    double p1_lo, p1_hi;
    int p0_only;
    Cand* out;            // one result per CTA
    const SearchGeomDev* gd;   // optional: u0 from device memory
};

constexpr int POLISH_MAXIT = 8;

struct PolishShared {
    double part[SEARCH_THREADS / 32][3][GRAD_NSUMS];
    FG fg[3];                     // the three evaluation points' results (threads 0..2 -> thread 0)
    double base[GRAD_NSUMS];      // the combined sums at the base point
    double y0, y1;        // trial point of this round
    double H, pen;        // entropy term and 1000 * penalty at the last accepted point
    int frozen;
    int go;
};

// all threads: evaluate (f, g) at (y0, y1), (y0 + h0, y1), (y0, y1 + h1); thread 0 receives the three results
__device__ __forceinline__ void polish_eval(const float2* sp, int padshift, int n, double u0, double du, int m0, int m1,
                                            PolishShared& sh, FG (&out)[3], GradSums<double>* base_sums) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double y0 = sh.y0, y1 = sh.y1;
    const int frozen = sh.frozen;
    // per-point arithmetic in float32 (random 1e-7 relative errors average out over the sums; the anchor phase of every
    // thread's short chunk is reduced in float64), cross-warp combination and everything after it in float64.  The three
    // points are walked in ONE interleaved loop (lane_grad_multi): the iteration is bound by the latency of these short
    // dependent chains, not by their instruction count.
    GradSums<float> a[3];
    float t0f[3], tpuf[3];
    const double um0 = u0 + du * double(m0);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        a[k].init();
        const double q0 = y0 + (k == 1 ? NEWTON_H0 : 0.0), q1 = y1 + (k == 2 ? NEWTON_H1 : 0.0);
        // fold the first-order phase at the chunk start into the zero-order term in float64: the float32 reduction of the
        // walk then only sees |turns| < 1 plus the small in-chunk ramp
        double t0 = q0 / 360.0 + (q1 / 360.0) * um0;
        t0 -= floor(t0);
        t0f[k] = float(t0);
        tpuf[k] = float(q1 / 360.0);
    }
    lane_grad_multi<float, 3>(sp, padshift, m0, m1, n, t0f, tpuf, float(-du * double(m0)), float(du), frozen, a);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        // (u restarts at 0 inside the chunk: shift the u-weighted sums back to the global ramp)
        a[k].gP1 += float(um0) * a[k].gP0;
        a[k].As1 += float(um0) * a[k].As0;
        a[k].Al1 += float(um0) * a[k].Al0;
        a[k].umax += float(um0);
        a[k].warp_reduce();
        if (lane == 0) grad_store(a[k], sh.part[warp][k]);
    }
    __syncthreads();
    // threads 0..2 combine the warps' partial sums of one point each; thread 0 then collects the three results
    if (threadIdx.x < 3) {
        const int k = threadIdx.x;
        GradSums<double> tot = grad_load(sh.part[0][k]);
        for (int w = 1; w < SEARCH_THREADS / 32; ++w) tot.merge(grad_load(sh.part[w][k]));
        sh.fg[k] = acme_finish(tot, n);
        if (k == 0) grad_store(tot, sh.base);
    }
    if (threadIdx.x < 32) __syncwarp();
    if (threadIdx.x == 0) {
        for (int k = 0; k < 3; ++k) out[k] = sh.fg[k];
        if (base_sums) *base_sums = grad_load(sh.base);
    }
    __syncthreads();
}

// CTA-wide Newton iteration from (x0, x1) in the given branch mode; returns the final point, its value and the sums there
__device__ __forceinline__ void polish_run(const float2* sp, int padshift, int n, double u0, double du, int m0, int m1,
                                           const PolishParams& p, PolishShared& sh, int frozen, int maxit, double& x0, double& x1,
                                           double& fx, GradSums<double>& sums_x) {
    NewtonState st;
    st.x0 = x0; st.x1 = x1; st.f = CUDART_INF; st.g0 = st.g1 = 0.0; st.done = 0; st.iters = 0;
    FG ax0, ax1;                 // offset gradients at the accepted point
    double step0 = 0.0, step1 = 0.0;
    int rejects = 0;
    if (threadIdx.x == 0) { sh.y0 = x0; sh.y1 = x1; sh.frozen = frozen; sh.go = 1; }
    __syncthreads();
    for (int it = 0; it < maxit; ++it) {
        FG r[3];
        GradSums<double> bs;
        polish_eval(sp, padshift, n, u0, du, m0, m1, sh, r, &bs);
        if (threadIdx.x == 0) {
            const bool first = (it == 0);
            const bool small = fabs(step0) < 0.05 && fabs(step1) < 0.15;      // inside the quadratic bowl: trust the gradient
            if (first || r[0].f < st.f || (small && r[0].f <= st.f * (1.0 + 1e-6))) {
                st.x0 = sh.y0; st.x1 = sh.y1; st.f = r[0].f; st.g0 = r[0].g0; st.g1 = r[0].g1;
                sh.H = r[0].H; sh.pen = r[0].pen;
                ax0 = r[1]; ax1 = r[2];
                sums_x = bs;
                rejects = 0;
                // an offset point beyond the wall of the feasible region (its value explodes) says nothing about curvature
                // on this side: creep towards the wall along the gradient instead (rejections shrink the step)
                if (!(ax0.f < 4.0 * st.f)) { ax0 = r[0]; ax0.g0 = st.g0 + (st.g0 < 0.0 ? -1.0 : 1.0) * fabs(st.g0) * NEWTON_H0 / 0.05; }
                if (!(ax1.f < 4.0 * st.f)) { ax1 = r[0]; ax1.g1 = st.g1 + (st.g1 < 0.0 ? -1.0 : 1.0) * fabs(st.g1) * NEWTON_H1 / 0.15; }
                double t0, t1;
                newton_step(st, ax0, ax1, p.p0_only, p.p1_lo, p.p1_hi, &t0, &t1);
                step0 = t0 - st.x0; step1 = t1 - st.x1;
                if (!(st.f < CUDART_INF) || (!first && fabs(step0) < NEWTON_TOL0 && fabs(step1) < NEWTON_TOL1)) sh.go = 0;
                sh.y0 = t0; sh.y1 = t1;
            } else {
                ++rejects;
                step0 /= 3.0; step1 /= 3.0;
                sh.y0 = st.x0 + step0; sh.y1 = st.x1 + step1;
                if (rejects > 5) sh.go = 0;
            }
        }
        __syncthreads();
        if (!sh.go) break;
    }
    __syncthreads();
    if (threadIdx.x == 0) { x0 = st.x0; x1 = st.x1; fx = st.f; }
}

// grid = starts x POLISH_ROLES.  Role 0 minimises the objective itself; roles 1, 2 minimise the smooth branch functions
// f_k = A / (N d_k) of the two neighbours k = kmax -+ 1 of the |S| maximum (f = min_k f_k is their lower envelope: when the
// real-part maximum hops between the points of the peak top, two minima a fraction of a degree apart exist and the lower one
// may belong to a neighbouring branch) and report the TRUE objective at their branch minimum.  All roles run in parallel.
constexpr int POLISH_ROLES = 3;

__global__ void __launch_bounds__(SEARCH_THREADS) search_polish_kernel(const __grid_constant__ PolishParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sp = reinterpret_cast<float2*>(smem_raw);
    __shared__ PolishShared sh;
    __shared__ float amax_v[SEARCH_THREADS / 32];
    __shared__ int amax_i[SEARCH_THREADS / 32];
    const int n = p.n;
    const int chunk = (n + SEARCH_THREADS - 1) / SEARCH_THREADS;
    const int padshift = ilog2_ceil(chunk);
    load_padded(sp, p.spec, n, padshift);
    const double pu0 = p.gd != nullptr ? p.gd->u0 : p.u0;
    const int start = int(blockIdx.x) / POLISH_ROLES, role = int(blockIdx.x) % POLISH_ROLES;
    // the (start+1)-th best candidate that is distinct from all better ones (block_argmin's barriers also publish `sp`)
    Cand chosen[ZOOM_MAX_STARTS];
    Cand centre = block_argmin(p.prev, p.n_prev);
    for (int s = 1; s <= start; ++s) {
        chosen[s - 1] = centre;
        centre = block_argmin(p.prev, p.n_prev, chosen, s, p.sep0, p.sep1);
    }
    // first index of the |S| maximum (the peak top the real-part maximum lives on)
    int kmax = 0;
    if (role != 0) {
        float bv = -1.f;
        int bi = 0x7fffffff;
        for (int m = threadIdx.x; m < n; m += SEARCH_THREADS) {
            const float2 x = sp[m + (m >> padshift)];
            const float v = x.x * x.x + x.y * x.y;
            if (v > bv) { bv = v; bi = m; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if ((threadIdx.x & 31) == 0) { amax_v[threadIdx.x >> 5] = bv; amax_i[threadIdx.x >> 5] = bi; }
        __syncthreads();
        bv = amax_v[0]; bi = amax_i[0];
        for (int w = 1; w < SEARCH_THREADS / 32; ++w)
            if (amax_v[w] > bv || (amax_v[w] == bv && amax_i[w] < bi)) { bv = amax_v[w]; bi = amax_i[w]; }
        kmax = bi;
    }
    const int frozen = role == 0 ? -1 : kmax + (role == 1 ? -1 : 1);
    Cand res;
    res.f = CUDART_INF; res.p0 = centre.p0; res.p1 = centre.p1; res.pad = 0;
    if (centre.f < CUDART_INF && (role == 0 || (frozen >= 0 && frozen < n))) {             // (uniform over the CTA)
        const int m0 = min(int(threadIdx.x) * chunk, n), m1 = min(m0 + chunk, n);
        double x0 = centre.p0, x1 = centre.p1, fx = CUDART_INF;
        GradSums<double> sx;
        polish_run(sp, padshift, n, pu0, p.du, m0, m1, p, sh, frozen, POLISH_MAXIT, x0, x1, fx, sx);
        __shared__ double bx0, bx1, bf;
        __shared__ int wall;
        if (threadIdx.x == 0) { bx0 = x0; bx1 = x1; bf = fx; sh.y0 = x0; sh.y1 = x1; sh.frozen = -1; }
        __syncthreads();
        if (role != 0 && bf < CUDART_INF) {
            // the true objective (free maximum) at the branch minimum
            FG r[3];
            polish_eval(sp, padshift, n, pu0, p.du, m0, m1, sh, r, nullptr);
            if (threadIdx.x == 0) { bf = r[0].f; sh.H = r[0].H; sh.pen = r[0].pen; }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            // the penalty (almost) vanishes here: the optimum is CONSTRAINED by the wall 1000 P > 0 (or the entropy term's
            // roughness decides) and the direct-search levels locate it (flag travels in Cand::pad)
            wall = (bf < CUDART_INF && sh.pen < 4.0 * sh.H) ? 1 : 0;
            double w = fmod(bx0 + 180.0, 360.0);
            if (w < 0) w += 360.0;
            res.f = bf; res.p0 = w - 180.0; res.p1 = bx1; res.pad = double(wall);
            // never return something worse than the grid point the polish started from
            if (!(res.f <= centre.f)) { res.f = centre.f; res.p0 = centre.p0; res.p1 = centre.p1; res.pad = 1.0; }
        }
    }
    if (threadIdx.x == 0) p.out[blockIdx.x] = res;
}

// final pick: result = {p0, p1, f, 0}; optionally the phase parameters of pass 2 (K1PhaseDev) so that the chain needs no host
// read-back: turns(m) = a + b*m with a = p0/360 + (p1/360) u0, b = (p1/360) du  (phasing.py:56-69 on a uniform axis)
struct FinalizePhase {
    void* ph_out;              // K1PhaseDev* or null
    const SearchGeomDev* gd;   // u0 from device memory (or null: u0 below)
    double u0, du;
    int n;                     // transform length (fold / step tables are per n/16 block)
    int p0_only;
};
struct K1PhaseDevLayout {      // == xmr::K1PhaseDev (k1_fft.cuh); repeated here so that the search does not include the FFT
    double ph_a_turns, ph_b_turns;
    float2 ph_step[16];
    float2 ph_fold[16];
};
__global__ void search_finalize_kernel(const Cand* list, int n_list, double* result, FinalizePhase fp) {
    const Cand b = block_argmin(list, n_list);
    if (threadIdx.x == 0) {
        result[0] = b.p0;
        result[1] = b.p1;
        result[2] = b.f;
        result[3] = 0.0;
    }
    if (fp.ph_out != nullptr && threadIdx.x < 16) {
        K1PhaseDevLayout* ph = static_cast<K1PhaseDevLayout*>(fp.ph_out);
        const double u0 = fp.gd != nullptr ? fp.gd->u0 : fp.u0;
        const double p1 = fp.p0_only ? 0.0 : b.p1;
        const double a = b.p0 / 360.0 + (p1 / 360.0) * u0, bt = (p1 / 360.0) * fp.du;
        const int d = threadIdx.x, q = fp.n / 16;
        double turns = bt * double(q) * double(d);
        turns -= floor(turns);
        double sn, cs;
        sincospi(2.0 * turns, &sn, &cs);
        ph->ph_step[d] = make_float2(float(cs), float(sn));
        const long long m0 = ((long long)q * d + fp.n / 2) % fp.n;
        double tf = a + bt * double(m0);
        tf -= floor(tf);
        sincospi(2.0 * tf, &sn, &cs);
        ph->ph_fold[d] = make_float2(float(cs), float(sn));
        if (d == 0) { ph->ph_a_turns = a; ph->ph_b_turns = bt; }
    }
}

}  // namespace xmr
