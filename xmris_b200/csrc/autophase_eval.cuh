// Objective evaluation primitives of the autophase search (reference: src/xmris/processing/phasing.py:100-157).
//
// A spectrum S[0..N) sits in shared memory in a padded layout, idx(m) = m + (m >> padshift), so that 32 lanes that
// each walk a contiguous chunk of 2^padshift points hit distinct banks.  A lane evaluates K zero-order candidates p0[k] at one
// first-order value p1 while walking its chunk once:
//     w_m  = S_m * exp(i*2*pi*turns_per_u*u_m),  u_m = u0 + du*m          (first-order rotation, shared by all k)
//     d_mk = Re(w_m * exp(i*p0_k)) = w.x*c_k - w.y*s_k                     (phasing.py:69-73, real part)
// and accumulates, per candidate, what the three objectives need:
//   ACME (phasing.py:100-122):   A0 = sum |d_{m+1}-d_m|,  A1 = sum |D| log2|D|,  A2 = sum min(d,0)^2,  A3 = max d
//        score = (H + 1000*P)/N/max d,  g = |D|/2, H = ln G - (sum g ln g)/G, G = A0/2, P = A2
//        (the reference's `if sum(d-|d|) < 0` switch is redundant: P = 0 exactly when no d is negative)
//   POSITIVITY (phasing.py:142-157): over ROI  A0 = sum_{d<0} |d|,  A1 = sum_{d>0} d;   score = 5*A0 - A1
//   PEAK_MINIMA (phasing.py:125-139): A0 = min d over [start,target), A1 = min d over [target,end), A2 = d[target]
//        score = |mina - minb| with the reference's empty-range fallbacks
// The incremental phasor is re-anchored (exact sincospi of the reduced turns) at the start of every chunk.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace xmr {

enum : int { METHOD_ACME = 0, METHOD_PEAK_MINIMA = 1, METHOD_POSITIVITY = 2 };

template <typename R> struct RealOps;
template <> struct RealOps<float> {
    static __device__ __forceinline__ void sincospi2(float turns, float* s, float* c) { sincospif(2.0f * turns, s, c); }
    static __device__ __forceinline__ float log2r(float x) {   // one MUFU.LG2, denormals flushed (x >= tiny() here)
        float y;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
        return y;
    }
    static __device__ __forceinline__ float mn(float a, float b) { return fminf(a, b); }
    static __device__ __forceinline__ float mx(float a, float b) { return fmaxf(a, b); }
    static __device__ __forceinline__ float ab(float a) { return fabsf(a); }
    static __device__ __forceinline__ float tiny() { return 1.17549435e-38f; }
    static __device__ __forceinline__ float inf() { return CUDART_INF_F; }
    static __device__ __forceinline__ float lnr(float x) { return __logf(x); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdividef(a, b); }
};
template <> struct RealOps<double> {
    static __device__ __forceinline__ void sincospi2(double turns, double* s, double* c) { sincospi(2.0 * turns, s, c); }
    static __device__ __forceinline__ double log2r(double x) { return log2(x); }
    static __device__ __forceinline__ double mn(double a, double b) { return fmin(a, b); }
    static __device__ __forceinline__ double mx(double a, double b) { return fmax(a, b); }
    static __device__ __forceinline__ double ab(double a) { return fabs(a); }
    static __device__ __forceinline__ double tiny() { return 2.2250738585072014e-308; }
    static __device__ __forceinline__ double inf() { return CUDART_INF; }
    static __device__ __forceinline__ double lnr(double x) { return log(x); }
    static __device__ __forceinline__ double div(double a, double b) { return a / b; }
};

struct ScoreGeom {
    int n;            // spectrum length
    int target_idx;   // ROI methods
    int roi_start;    // max(0, target - width)
    int roi_end;      // min(n, target + width)
};

template <typename R, int METHOD, int K>
struct Acc {
    R a[K][4];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (METHOD == METHOD_ACME) { a[k][0] = 0; a[k][1] = 0; a[k][2] = 0; a[k][3] = -RealOps<R>::inf(); }
            else if (METHOD == METHOD_POSITIVITY) { a[k][0] = 0; a[k][1] = 0; a[k][2] = 0; a[k][3] = 0; }
            else { a[k][0] = RealOps<R>::inf(); a[k][1] = RealOps<R>::inf(); a[k][2] = -RealOps<R>::inf(); a[k][3] = 0; }
        }
    }
    // associative combine (used across lanes and across warps)
    __device__ __forceinline__ static R comb(int slot, R x, R y) {
        if (METHOD == METHOD_ACME) return slot == 3 ? (x > y ? x : y) : x + y;
        if (METHOD == METHOD_POSITIVITY) return x + y;
        return slot == 2 ? (x > y ? x : y) : (slot == 3 ? x + y : (x < y ? x : y));
    }
    __device__ __forceinline__ void warp_reduce() {
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                if (METHOD == METHOD_POSITIVITY && s >= 2) continue;
                if (METHOD == METHOD_PEAK_MINIMA && s == 3) continue;
                R v = a[k][s];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v = comb(s, v, __shfl_xor_sync(0xffffffffu, v, off));
                a[k][s] = v;
            }
    }
    // RAW: the reference's formula as written, including the negative branch max(d) <= 0 that the searches reject
    // (xmr_autophase_score_c64: the evaluator checked against the reference's own score functions).
    template <bool RAW = false>
    __device__ __forceinline__ R score(int k, const ScoreGeom& g) const {
        if (METHOD == METHOD_ACME) {
            const R LN2 = R(0.69314718055994530942);
            const R A0 = a[k][0], A1 = a[k][1];
            const R G = R(0.5) * A0;
            // sum g ln g = 0.5 * (ln2 * A1 - A0 * ln2);   H = ln G - (sum g ln g) / G
            const R sglg = R(0.5) * LN2 * (A1 - A0);
            const R H = RealOps<R>::lnr(G) - RealOps<R>::div(sglg, G);
            // The reference divides by the SIGNED max(d) (phasing.py:122): where the whole real part is negative the
            // objective is negative with a pole at max(d) -> 0- (SURVEY finding 5).  Such upside-down candidates are
            // rejected: the search minimises over the region max(d) > 0, which is where the reference's optimiser
            // lands on well-posed data.
            if (!RAW && !(a[k][3] > R(0))) return RealOps<R>::inf();
            return RealOps<R>::div(H + R(1000) * a[k][2], R(g.n) * a[k][3]);
        }
        if (METHOD == METHOD_POSITIVITY) return R(5) * a[k][0] - a[k][1];
        const R mina = (g.roi_start < g.target_idx) ? a[k][0] : a[k][2];
        const R minb = (g.roi_end > g.target_idx) ? a[k][1] : a[k][2];
        const R df = mina - minb;
        return df < 0 ? -df : df;
    }
};

// One lane walks points [m0, m1) of the padded spectrum (idx(m) = m + (m >> padshift)) and folds them into `acc`
// for K candidates.  For ACME it also forms the differences D_m = d_{m+1} - d_m for m in [m0, m1) (reading point m1
// when m1 < n; that extra point only closes the last difference).
template <typename R, int METHOD, int K>
__device__ __forceinline__ void lane_accumulate_rt(const float2* sp, int padshift, int m0, int m1, const ScoreGeom& g,
                                                   R turns_per_u, R u0, R du, const R (&c0)[K], const R (&s0)[K],
                                                   Acc<R, METHOD, K>& acc) {
    if (METHOD != METHOD_ACME) {
        m0 = m0 > g.roi_start ? m0 : g.roi_start;
        m1 = m1 < g.roi_end ? m1 : g.roi_end;
    }
    if (m0 >= m1) return;
    R sr, cr, si, ci;
    {
        R t = turns_per_u * (u0 + du * R(m0));
        t -= floor(t);
        RealOps<R>::sincospi2(t, &sr, &cr);
        R ti = turns_per_u * du;
        ti -= floor(ti);
        RealOps<R>::sincospi2(ti, &si, &ci);
    }
    using O = RealOps<R>;
    if constexpr (METHOD == METHOD_ACME) {
        // point m0: no difference yet.  points m0+1 .. m1-1: full body.  point m1 (if it exists): closes the last
        // difference only (it belongs to the next lane).  No predicates inside the hot loop.
        R dprev[K];
        {
            const float2 S = sp[m0 + (m0 >> padshift)];
            const R wx = R(S.x) * cr - R(S.y) * sr, wy = R(S.x) * sr + R(S.y) * cr;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const R d = wx * c0[k] - wy * s0[k];
                dprev[k] = d;
                const R neg = O::mn(d, R(0));
                acc.a[k][2] += neg * neg;
                acc.a[k][3] = O::mx(acc.a[k][3], d);
            }
            const R ncr = cr * ci - sr * si;
            sr = cr * si + sr * ci;
            cr = ncr;
        }
#pragma unroll 2
        for (int m = m0 + 1; m < m1; ++m) {
            const float2 S = sp[m + (m >> padshift)];
            const R wx = R(S.x) * cr - R(S.y) * sr, wy = R(S.x) * sr + R(S.y) * cr;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const R d = wx * c0[k] - wy * s0[k];
                const R D = O::ab(d - dprev[k]);
                dprev[k] = d;
                acc.a[k][0] += D;
                acc.a[k][1] += D * O::log2r(O::mx(D, O::tiny()));
                const R neg = O::mn(d, R(0));
                acc.a[k][2] += neg * neg;
                acc.a[k][3] = O::mx(acc.a[k][3], d);
            }
            const R ncr = cr * ci - sr * si;
            sr = cr * si + sr * ci;
            cr = ncr;
        }
        if (m1 < g.n) {
            const float2 S = sp[m1 + (m1 >> padshift)];
            const R wx = R(S.x) * cr - R(S.y) * sr, wy = R(S.x) * sr + R(S.y) * cr;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const R D = O::ab((wx * c0[k] - wy * s0[k]) - dprev[k]);
                acc.a[k][0] += D;
                acc.a[k][1] += D * O::log2r(O::mx(D, O::tiny()));
            }
        }
    } else {
    for (int m = m0; m < m1; ++m) {
        const float2 S = sp[m + (m >> padshift)];
        const R wx = R(S.x) * cr - R(S.y) * sr;
        const R wy = R(S.x) * sr + R(S.y) * cr;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const R d = wx * c0[k] - wy * s0[k];
            if (METHOD == METHOD_POSITIVITY) {
                acc.a[k][0] -= O::mn(d, R(0));
                acc.a[k][1] += O::mx(d, R(0));
            } else {
                if (m < g.target_idx) acc.a[k][0] = O::mn(acc.a[k][0], d);
                else acc.a[k][1] = O::mn(acc.a[k][1], d);
                if (m == g.target_idx) acc.a[k][2] = d;
            }
        }
        const R ncr = cr * ci - sr * si;
        sr = cr * si + sr * ci;
        cr = ncr;
    }
    }
}

// PEAK_MINIMA needs d[target_idx] even when the target lies outside every lane's ROI slice (it cannot: the target is
// always inside [roi_start, roi_end) unless the ROI is empty on one side, and then it is still >= roi_start).

}  // namespace xmr
