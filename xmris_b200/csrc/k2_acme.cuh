// K2-ACME: the whole chain for one voxel in one pass with a DERIVATIVE-BASED per-voxel (p0, p1) search  [autophase mode="all"]
//
//   zero-fill -> window -> FFT -> fftshift -> per-spectrum ACME search on the shared-memory resident spectrum -> phase -> store
//
// Round 1 spent ~2000 full-spectrum objective evaluations per voxel (dense grid + 7 zoom rounds); this kernel spends ~150
// (VERDICT r1 item 2 / SURVEY H1 (iv)):
//   L1  coarse localisation WITHOUT evaluating the objective: where the penalty term dominates (1000 P >> H, true whenever the
//       data are not scaled to ~1e-3), f ~ 1000 P / (N max d) with P = sum_m min(d_m, 0)^2 = sum_m r_m^2 h(theta_m + phi_m),
//       h(x) = min(cos x, 0)^2.  h has the Fourier series c0 + 2 sum_k c_k cos(k x) with c = (1/4, -2/(3 pi), 1/8, -2/(15 pi), 0,
//       2/(105 pi), 0, ...), so for one p1 the WHOLE p0 profile of P follows from four complex moments
//       F_k(p1) = sum_m r_m^2 e^{i k (theta_m + p1 u_m)}, k = 1, 2, 3, 5, accumulated on a 256-point decimation of the
//       spectrum: 179 p1 rows (45 deg apart) x 64 p0 values cost ~9 full evaluations instead of ~540.
//   L2  the 6 best mutually separated rows are re-evaluated with the TRUE objective on the full spectrum: 3 p1 rows x 8 p0
//       candidates around each cell (18 walks, lane_accumulate_rt).
//   L3  the 3 best cells are refined by a bounded quasi-Newton iteration on the ANALYTIC gradient (acme_grad.cuh): secant
//       Hessian at the first iterate, BFGS updates afterwards, ~10 gradient evaluations per start, converged to ~1e-3 deg.
//   F   out[m] = S[m] * exp(i (p0 + p1 u_m)) and (p0, p1, pivot, objective) per voxel
// Parity: tools/angle_parity.py / profiles/parity_r2.json (>= 1024 spectra per shape against the reference's optimiser and its
// tight-tolerance adjudicator).  The ROI methods (peak_minima, positivity) keep the round-1 kernel (k2_pervoxel.cuh).
#pragma once
#include <cuda_runtime.h>

#include "acme_grad.cuh"
#include "k2_pervoxel.cuh"

namespace xmr {

constexpr int K2A_NDEC = 256;      // blocks of the L1 stage (block moments of N / 256 points each)
constexpr int K2A_GSTRIDE = K2A_NDEC + K2A_NDEC / 8 + 8;   // padded length of one moment array
constexpr int K2A_NP0 = 64;        // L1 p0 grid: -180 + 5.625 k
#ifndef XMR_K2A_T
#define XMR_K2A_T 6
#endif
constexpr int K2A_T = XMR_K2A_T;   // L1 cells handed to L2 (6: 3 rows for the best four + 2 rows for the others = 16 walks; 4: 12 walks)
constexpr int K2A_NS = 3;          // L2 cells handed to the Newton refinement
constexpr int K2A_MAXIT = 16;
constexpr float K2A_ROWSTEP = 45.f;
constexpr float K2A_L2_DP1 = 15.f;   // L2 rows: cell.p1 + {-15, 0, 15}
constexpr float K2A_L2_DP0 = 2.5f;   // L2 p0 candidates: cell.p0 + (k - 3.5) * 2.5
constexpr float K2A_SIBLING = 20.f;  // sibling starts: result.p1 -+ 20 deg

struct K2aShared {      // lives in the MISC area
    uint64_t bar;
    float redv[32];
    int redi[32];
    float cell_p0[K2A_T], cell_p1[K2A_T];
    float l2f[K2A_T * 3], l2p0[K2A_T * 3], l2p1[K2A_T * 3];
    double part[K2A_NS][8][3][GRAD_NSUMS];    // [start][segment][evaluation point][sum] (per-warp partial sums)
    int gs;                                   // segments (warps) per active start in this iteration
    double y0[K2A_NS], y1[K2A_NS];            // trial point per start
    int np[K2A_NS];                           // evaluation points wanted at the trial point: 1 (gradient) or 3 (+ secant Hessian)
    int active[K2A_NS];
    int any_active;
    double res_f[K2A_NS], res_p0[K2A_NS], res_p1[K2A_NS];
    double res_h[K2A_NS], res_pen[K2A_NS];    // entropy term H and penalty 1000 P at the result
    double l2_f, l2_p0, l2_p1;
    float zf[8], zp0[8], zp1[8];              // direct-search rounds: per-row results
    float zc0, zc1, zcf;
    int wall;
};

template <int N>
struct K2aSmem {
    using C = FftCfg<N>;
    static constexpr int PADSHIFT = ilog2(N / 32 > 0 ? N / 32 : 1);
    // N >= 8192: ONE buffer is TMA landing slot, both FFT exchanges (in place, as K1 does at this length) and then the padded
    // spectrum the search walks: 92 KB per CTA instead of 156 KB, so that TWO CTAs fit on an SM (four instead of two warps per
    // SM sub-partition for a latency-bound search); the next voxel's FID is fetched at the end of the voxel.
    static constexpr bool IPB = (N >= 8192);
    static constexpr size_t SLOT = IPB ? 0 : size_t(C::N) * sizeof(float2);
    static constexpr size_t SPN = (size_t(C::N) + (size_t(C::N) >> PADSHIFT) + 2);
    static constexpr size_t B = (C::SIZE_B > SPN ? size_t(C::SIZE_B) : SPN) * sizeof(float2);
    static constexpr size_t DEC = size_t(4) * K2A_GSTRIDE * sizeof(float2);     // block moments G_1, G_2, G_3, G_5
    static constexpr size_t ROWS = size_t(192) * 2 * sizeof(float);
    static constexpr size_t MISC = (sizeof(K2aShared) + 255) & ~size_t(255);
    static constexpr size_t TOTAL = SLOT + B + DEC + ROWS + MISC;
};


template <int N>
__global__ void __launch_bounds__(FftCfg<N>::T, (FftCfg<N>::T >= 256 ? 2 : (FftCfg<N>::T >= 128 ? 4 : 8)))
k2_acme_kernel(const __grid_constant__ K2Params p) {
    using C = FftCfg<N>;
    using SM = K2aSmem<N>;
    constexpr bool TW_PERSIST = false;
    constexpr int WPS = C::T / 32;
    constexpr int PADSHIFT = SM::PADSHIFT;
    constexpr int L = N / 32;                      // points per lane when one warp walks the whole spectrum
    constexpr int GSMAX = WPS;                     // warps per Newton start: all of them once a single start is left
    constexpr int DEC = N / K2A_NDEC;
    static_assert(C::T >= 32 && N >= 512, "per-voxel kernel needs N >= 512");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr bool IPB = SM::IPB;
    float2* Bbuf = reinterpret_cast<float2*>(smem_raw + SM::SLOT);
    float2* slot = IPB ? Bbuf : reinterpret_cast<float2*>(smem_raw);
    float2* sp = Bbuf;                                                     // padded spectrum reuses exchange B
    float2* gdec = reinterpret_cast<float2*>(smem_raw + SM::SLOT + SM::B);
    float* rowf = reinterpret_cast<float*>(smem_raw + SM::SLOT + SM::B + SM::DEC);
    float* rowp0 = rowf + 192;
    K2aShared& sh = *reinterpret_cast<K2aShared*>(smem_raw + SM::SLOT + SM::B + SM::DEC + SM::ROWS);

    const int t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;
    const int n_in = p.n_in;
    const bool need_load_barrier = (p.pad_left != 0);

    float2 tw_persist[1];
    float2 tw0_base[C::C0 * 2], tw1_base[C::C1 * 2];
    init_twiddles<C, false>(t, p.twN, nullptr, tw0_base, tw1_base);
    float wcol[C::C0];
#pragma unroll
    for (int j = 0; j < C::C0; ++j) wcol[j] = (p.win && !p.win_table) ? p.win[t + C::T * j] : p.scale;

    auto issue = [&](long long v) {
        const uint32_t row_bytes = uint32_t(n_in) * 8u;
        mbar_arrive_expect_tx(&sh.bar, row_bytes);
        bulk_g2s(slot, p.in + v * n_in, row_bytes, &sh.bar);
    };
    if (p.use_tma) {
        if (t == 0) {
            mbar_init(&sh.bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (t == 0 && (long long)blockIdx.x < p.batch) issue(blockIdx.x);
    }

    ScoreGeom geom;
    geom.n = N;
    geom.target_idx = 0;
    geom.roi_start = 0;
    geom.roi_end = N;
    const float duf = float(p.du);
    // Fourier coefficients of h(x) = min(cos x, 0)^2 (k = 0, 1, 2, 3, 5; c4 = c6 = 0)
    const float HC0 = 0.25f, HC1 = -0.21220659078919378f, HC2 = 0.125f, HC3 = -0.042441318157838755f, HC5 = 0.006063045451119822f;

    int it = 0;
    for (long long vox = blockIdx.x; vox < p.batch; vox += gridDim.x, ++it) {
        // ---- A: FFT (as K1) ------------------------------------------------------------------------------------------
        if (p.use_tma) {
            mbar_wait(&sh.bar, it & 1);
        } else {
            __syncthreads();
            for (int k = t; k < n_in; k += C::T) slot[k] = p.in[vox * n_in + k];
            __syncthreads();
        }
        int mstar;
        if (!p.spec_in) {
            float2 v[C::E];
            if (p.win_table) stage0_load<C, 1>(t, slot, n_in, p.pad_left, 0, p.scale, p.win, wcol, p.win_rows, v);
            else stage0_load<C, 2>(t, slot, n_in, p.pad_left, 0, p.scale, p.win, wcol, p.win_rows, v);
            if (need_load_barrier) __syncthreads();
            stage0_store<C, false, TW_PERSIST>(t, slot, v, tw_persist, tw0_base);
            __syncthreads();
            if (IPB) {
                stage1_load<C>(t, slot, v);
                __syncthreads();                               // exchange B overwrites exchange A
                stage1_store<C, false>(t, Bbuf, v, tw1_base);
            } else {
                stage1<C, false>(t, slot, Bbuf, tw1_base);
            }
            __syncthreads();
            if (!IPB && p.use_tma && t == 0) {
                const long long nv = vox + gridDim.x;
                if (nv < p.batch) {
                    fence_proxy_async_smem();
                    issue(nv);
                }
            }
            stage2<C, false>(t, Bbuf, v);
            // ---- B: |S| argmax of this spectrum (the voxel's pivot, phasing.py:229-238) -----------------------------------
            constexpr int Q = C::R0 * C::R1;
            float best = -1.f;
            int besti = 0x7fffffff;
#pragma unroll
            for (int j = 0; j < C::C2; ++j)
#pragma unroll
                for (int d = 0; d < C::R2; ++d) {
                    const float2 x = v[j * C::R2 + d];
                    amax_combine(best, besti, x.x * x.x + x.y * x.y, (t + C::T * j + Q * d + p.out_shift) & (N - 1));
                }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, off);
                const int oi = __shfl_xor_sync(0xffffffffu, besti, off);
                amax_combine(best, besti, ov, oi);
            }
            if (lane == 0) { sh.redv[warp] = best; sh.redi[warp] = besti; }
            __syncthreads();                                   // also: every thread is done reading exchange B
            // ---- C: spectrum -> padded shared layout ---------------------------------------------------------------------
#pragma unroll
            for (int j = 0; j < C::C2; ++j)
#pragma unroll
                for (int d = 0; d < C::R2; ++d) {
                    const int m = (t + C::T * j + Q * d + p.out_shift) & (N - 1);
                    sp[m + (m >> PADSHIFT)] = v[j * C::R2 + d];
                }
            best = sh.redv[0];
            besti = sh.redi[0];
            for (int w = 1; w < WPS; ++w) amax_combine(best, besti, sh.redv[w], sh.redi[w]);
            mstar = besti;
        } else {
            float best = -1.f;
            int besti = 0x7fffffff;
            if (IPB) {
                // the padded layout is written over the buffer the spectrum landed in: every thread reads its points first
                float2 tmp[N / C::T];
#pragma unroll
                for (int i = 0; i < N / C::T; ++i) tmp[i] = slot[t + C::T * i];
                __syncthreads();
#pragma unroll
                for (int i = 0; i < N / C::T; ++i) {
                    const int m = t + C::T * i;
                    sp[m + (m >> PADSHIFT)] = tmp[i];
                    amax_combine(best, besti, tmp[i].x * tmp[i].x + tmp[i].y * tmp[i].y, m);
                }
            } else
            for (int m = t; m < N; m += C::T) {
                const float2 x = slot[m];
                sp[m + (m >> PADSHIFT)] = x;
                amax_combine(best, besti, x.x * x.x + x.y * x.y, m);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, off);
                const int oi = __shfl_xor_sync(0xffffffffu, besti, off);
                amax_combine(best, besti, ov, oi);
            }
            if (lane == 0) { sh.redv[warp] = best; sh.redi[warp] = besti; }
            __syncthreads();
            if (!IPB && p.use_tma && t == 0) {
                const long long nv = vox + gridDim.x;
                if (nv < p.batch) {
                    fence_proxy_async_smem();
                    issue(nv);
                }
            }
            best = sh.redv[0];
            besti = sh.redi[0];
            for (int w = 1; w < WPS; ++w) amax_combine(best, besti, sh.redv[w], sh.redi[w]);
            mstar = besti;
        }
        __syncthreads();                                       // sp complete
        const double u0d = p.fixed_pivot ? p.u0_fixed : -p.du * double(mstar);
        const float u0 = float(u0d);
        const float2 Spiv = sp[mstar + (mstar >> PADSHIFT)];
        const float upiv = u0 + duf * float(mstar);
        const float apiv = sqrtf(Spiv.x * Spiv.x + Spiv.y * Spiv.y);

        // ---- D: L1, series surrogate of the penalty term from BLOCK MOMENTS of the whole spectrum ------------------------------
        // G_k[b] = sum over the DEC points m of block b of r_m^2 e^{i k theta_m}, k = 1, 2, 3, 5 (every point contributes: narrow
        // peaks are not stepped over), F_0 = sum r^2.  Per row, F_k(p1) = sum_b G_k[b] e^{i k p1 u_b} with u at the block centre
        // (the first-order phase varies by < p1*du*DEC/2 <= 8 deg inside a block).
        {
            constexpr int LG = ilog2(DEC);
            float f0acc = 0.f;
            for (int base = warp * 32; base < N; base += C::T) {           // 32 consecutive points per warp step: conflict-free
                const int m = base + lane;
                const float2 S = sp[m + (m >> PADSHIFT)];
                const float r2 = S.x * S.x + S.y * S.y;
                const float inv = r2 > 0.f ? rsqrtf(r2) : 0.f;
                const float zx = S.x * inv, zy = S.y * inv;                    // e^{i theta}
                const float bx = zx * zx - zy * zy, by = 2.f * zx * zy;        // ^2
                const float cx = bx * zx - by * zy, cy = bx * zy + by * zx;    // ^3
                const float ex = bx * cx - by * cy, ey = bx * cy + by * cx;    // ^5
                float g[8] = {r2 * zx, r2 * zy, r2 * bx, r2 * by, r2 * cx, r2 * cy, r2 * ex, r2 * ey};
                f0acc += r2;
#pragma unroll
                for (int s = 0; s < (LG < 5 ? LG : 5); ++s)                     // segmented sum over the DEC lanes of a block
#pragma unroll
                    for (int q = 0; q < 8; ++q) g[q] += __shfl_xor_sync(0xffffffffu, g[q], 1 << s);
                if (DEC >= 32) {
                    // a block spans DEC/32 warp steps: accumulate in shared memory (one warp owns a block at a time)
                    const int b = m / DEC;
                    if (lane == 0) {
                        const bool first = (m % DEC) == 0;
                        float2* gp = gdec + (b + (b >> 3));
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float2 cur = first ? make_float2(0.f, 0.f) : gp[q * K2A_GSTRIDE];
                            gp[q * K2A_GSTRIDE] = make_float2(cur.x + g[2 * q], cur.y + g[2 * q + 1]);
                        }
                    }
                } else if ((lane & (DEC - 1)) == 0) {
                    const int b = m / DEC;
                    float2* gp = gdec + (b + (b >> 3));
#pragma unroll
                    for (int q = 0; q < 4; ++q) gp[q * K2A_GSTRIDE] = make_float2(g[2 * q], g[2 * q + 1]);
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) f0acc += __shfl_xor_sync(0xffffffffu, f0acc, off);
            if (lane == 0) sh.redv[warp] = f0acc;
        }
        __syncthreads();
        constexpr int PPL = K2A_NDEC / 32;                     // blocks per lane (contiguous)
        static_assert(PPL == 8, "the padding rule b + (b >> 3) assumes 8 blocks per lane");
        float F0 = 0.f;
#pragma unroll
        for (int w = 0; w < WPS; ++w) F0 += sh.redv[w];
        const int nrows = p.p0_only ? 1 : K2_NP1;
        for (int r = warp; r < nrows; r += WPS) {
            const float p1 = p.p0_only ? 0.f : fminf(-4000.f + K2A_ROWSTEP * float(r), 4000.f);
            const float tpu = p1 * (1.0f / 360.0f);
            float sr, cr, ss, cs;
            {
                float ta = tpu * (u0 + duf * (float(DEC * lane * PPL) + 0.5f * float(DEC - 1)));      // centre of the lane's first block
                ta -= floorf(ta);
                sincospif(2.0f * ta, &sr, &cr);
                float ts = tpu * duf * float(DEC);
                ts -= floorf(ts);
                sincospif(2.0f * ts, &ss, &cs);
            }
            float f1x = 0.f, f1y = 0.f, f2x = 0.f, f2y = 0.f, f3x = 0.f, f3y = 0.f, f5x = 0.f, f5y = 0.f;
#pragma unroll
            for (int i = 0; i < PPL; ++i) {
                const float2* gp = gdec + lane * (PPL + 1) + i;
                const float2 g1 = gp[0], g2 = gp[K2A_GSTRIDE], g3 = gp[2 * K2A_GSTRIDE], g5 = gp[3 * K2A_GSTRIDE];
                const float bx = cr * cr - sr * sr, by = 2.f * cr * sr;                // e^{2 i p1 u_b}
                const float cx = bx * cr - by * sr, cy = bx * sr + by * cr;            // ^3
                const float ex = bx * cx - by * cy, ey = bx * cy + by * cx;            // ^5
                f1x = fmaf(g1.x, cr, fmaf(-g1.y, sr, f1x)); f1y = fmaf(g1.x, sr, fmaf(g1.y, cr, f1y));
                f2x = fmaf(g2.x, bx, fmaf(-g2.y, by, f2x)); f2y = fmaf(g2.x, by, fmaf(g2.y, bx, f2y));
                f3x = fmaf(g3.x, cx, fmaf(-g3.y, cy, f3x)); f3y = fmaf(g3.x, cy, fmaf(g3.y, cx, f3y));
                f5x = fmaf(g5.x, ex, fmaf(-g5.y, ey, f5x)); f5y = fmaf(g5.x, ey, fmaf(g5.y, ex, f5y));
                const float ncr = cr * cs - sr * ss;
                sr = cr * ss + sr * cs;
                cr = ncr;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                f1x += __shfl_xor_sync(0xffffffffu, f1x, off); f1y += __shfl_xor_sync(0xffffffffu, f1y, off);
                f2x += __shfl_xor_sync(0xffffffffu, f2x, off); f2y += __shfl_xor_sync(0xffffffffu, f2y, off);
                f3x += __shfl_xor_sync(0xffffffffu, f3x, off); f3y += __shfl_xor_sync(0xffffffffu, f3y, off);
                f5x += __shfl_xor_sync(0xffffffffu, f5x, off); f5y += __shfl_xor_sync(0xffffffffu, f5y, off);
            }
            // the pivot point (global |S| maximum) stands in for max(d): it dominates wherever the candidate is any good
            float px, py;
            {
                float tp = tpu * upiv;
                tp -= floorf(tp);
                float sn, c;
                sincospif(2.0f * tp, &sn, &c);
                px = Spiv.x * c - Spiv.y * sn;
                py = Spiv.x * sn + Spiv.y * c;
            }
            float bf = CUDART_INF_F, bp0 = 0.f;
#pragma unroll
            for (int i = 0; i < K2A_NP0 / 32; ++i) {
                const float p0 = -180.f + (360.f / K2A_NP0) * float(lane + 32 * i);
                float s1, c1;
                sincospif(p0 * (1.0f / 180.0f), &s1, &c1);
                const float c2 = c1 * c1 - s1 * s1, s2 = 2.f * c1 * s1;
                const float c3 = c2 * c1 - s2 * s1, s3 = c2 * s1 + s2 * c1;
                const float c5 = c2 * c3 - s2 * s3, s5 = c2 * s3 + s2 * c3;
                float P = HC0 * F0;
                P = fmaf(2.f * HC1, c1 * f1x - s1 * f1y, P);
                P = fmaf(2.f * HC2, c2 * f2x - s2 * f2y, P);
                P = fmaf(2.f * HC3, c3 * f3x - s3 * f3y, P);
                P = fmaf(2.f * HC5, c5 * f5x - s5 * f5y, P);
                const float dm = px * c1 - py * s1;
                const float f = (dm > 0.05f * apiv) ? __fdividef(fmaxf(P, 1e-9f * F0), dm) : CUDART_INF_F;
                if (p.p0_only) rowf[lane + 32 * i] = f;            // one row: keep the whole p0 profile
                if (f < bf) { bf = f; bp0 = p0; }
            }
            if (!p.p0_only) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const float of = __shfl_xor_sync(0xffffffffu, bf, off);
                    const float op = __shfl_xor_sync(0xffffffffu, bp0, off);
                    if (of < bf || (of == bf && op < bp0)) { bf = of; bp0 = op; }
                }
                if (lane == 0) { rowf[r] = bf; rowp0[r] = bp0; }
            }
        }
        __syncthreads();
        // the K2A_T best rows, +-1 row suppressed around each pick (p0_only: the best p0 values, +-2 grid points suppressed)
        if (warp == 0) {
            const int nitem = p.p0_only ? K2A_NP0 : nrows;
            const int sup = p.p0_only ? 2 : 1;
            for (int s = 0; s < K2A_T; ++s) {
                float bf = CUDART_INF_F;
                int bc = 0x7fffffff;
                for (int c = lane; c < nitem; c += 32) {
                    const float f = rowf[c];
                    if (f < bf || (f == bf && c < bc)) { bf = f; bc = c; }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const float of = __shfl_xor_sync(0xffffffffu, bf, off);
                    const int oc = __shfl_xor_sync(0xffffffffu, bc, off);
                    if (of < bf || (of == bf && oc < bc)) { bf = of; bc = oc; }
                }
                const bool none = !(bf < CUDART_INF_F);
                if (lane == 0) {
                    if (none) {
                        sh.cell_p0[s] = CUDART_NAN_F;             // fewer usable cells than K2A_T
                        sh.cell_p1[s] = 0.f;
                    } else if (p.p0_only) {
                        sh.cell_p0[s] = -180.f + (360.f / K2A_NP0) * float(bc);
                        sh.cell_p1[s] = 0.f;
                    } else {
                        sh.cell_p0[s] = rowp0[bc];
                        sh.cell_p1[s] = fminf(-4000.f + K2A_ROWSTEP * float(bc), 4000.f);
                    }
                }
                __syncwarp();
                if (!none) {
                    for (int c = lane; c < nitem; c += 32) {
                        int dc = abs(c - bc);
                        if (p.p0_only) dc = min(dc, nitem - dc);
                        if (dc <= sup) rowf[c] = CUDART_INF_F;
                    }
                }
                __syncwarp();
            }
        }
        __syncthreads();

        // ---- E: L2, the true objective around the K2A_T cells (full spectrum, 3 p1 rows x 8 p0 per cell) ----------------
        // 3 rows (cell.p1 - 15, 0, + 15) for the four best cells, 2 rows (-+ 7.5) for the other two: 16 walks = two full rounds
        // of 8 warps
        const int l2items = p.p0_only ? K2A_T : (K2A_T >= 6 ? 16 : 3 * K2A_T);
        for (int item = warp; item < l2items; item += WPS) {
            const int cell = p.p0_only ? item : (item < 12 ? item / 3 : 4 + (item - 12) / 2);
            const int rr = p.p0_only ? 0 : (item < 12 ? item % 3 : (item - 12) % 2);
            const float dp1 = p.p0_only ? 0.f : (item < 12 ? K2A_L2_DP1 * float(rr - 1) : K2A_L2_DP1 * (float(rr) - 0.5f));
            const float c0c = sh.cell_p0[cell], c1c = sh.cell_p1[cell];
            float rbf = CUDART_INF_F, rb0 = 0.f, rb1 = 0.f;
            if (c0c == c0c) {                                       // (warp-uniform)
                const float p1 = p.p0_only ? 0.f : fminf(fmaxf(c1c + dp1, -4000.f), 4000.f);
                const float tpu = p1 * (1.0f / 360.0f);
                float c0[K2_K], s0[K2_K], p0k[K2_K];
#pragma unroll
                for (int k = 0; k < K2_K; ++k) {
                    float q0 = c0c + (float(k) - 3.5f) * K2A_L2_DP0;
                    q0 = q0 > 180.f ? q0 - 360.f : (q0 < -180.f ? q0 + 360.f : q0);
                    p0k[k] = q0;
                    sincospif(q0 * (1.0f / 180.0f), &s0[k], &c0[k]);
                }
                Acc<float, METHOD_ACME, K2_K> acc;
                acc.init();
                lane_accumulate_rt<float, METHOD_ACME, K2_K>(sp, PADSHIFT, lane * L, (lane + 1) * L, geom, tpu, u0, duf, c0, s0, acc);
                acc.warp_reduce();
                rb1 = p1;
#pragma unroll
                for (int k = 0; k < K2_K; ++k) {
                    const float f = acc.score(k, geom);
                    if (f < rbf) { rbf = f; rb0 = p0k[k]; }
                }
            }
            if (lane == 0) { sh.l2f[item] = rbf; sh.l2p0[item] = rb0; sh.l2p1[item] = rb1; }
        }
        __syncthreads();
        if (t == 0) {
            // the K2A_NS best row results that are mutually distinct (a valley can hold two minima ~20 deg of p1 apart where
            // two branches of max(d) meet: rows of the same cell may both become starts)
            const int nitems = l2items;
            int any = 0;
            for (int s = 0; s < K2A_NS; ++s) {
                int b = -1;
                for (int i = 0; i < nitems; ++i) {
                    if (!(sh.l2f[i] < CUDART_INF_F)) continue;
                    bool distinct = true;
                    for (int q = 0; q < s; ++q) {
                        float d0 = fabsf(sh.l2p0[i] - float(sh.res_p0[q]));
                        d0 = fminf(d0, 360.f - d0);
                        if (sh.active[q] && d0 < 4.f && fabsf(sh.l2p1[i] - float(sh.res_p1[q])) < 10.f) distinct = false;
                    }
                    if (distinct && (b < 0 || sh.l2f[i] < sh.l2f[b])) b = i;
                }
                const bool ok = b >= 0;
                sh.y0[s] = ok ? double(sh.l2p0[b]) : 0.0;
                sh.y1[s] = ok ? double(sh.l2p1[b]) : 0.0;
                sh.np[s] = 3;
                sh.active[s] = ok ? 1 : 0;
                sh.res_f[s] = ok ? double(sh.l2f[b]) : CUDART_INF;
                sh.res_p0[s] = sh.y0[s];
                sh.res_p1[s] = sh.y1[s];
                any |= sh.active[s];
                if (ok) sh.l2f[b] = CUDART_INF_F;
            }
            sh.any_active = any;
            {
                int nact = 0;
                for (int s = 0; s < K2A_NS; ++s) nact += sh.active[s];
                sh.gs = nact <= 1 ? GSMAX : (nact == 2 ? (GSMAX >= 2 ? GSMAX / 2 : 1) : (GSMAX >= 4 ? GSMAX / 4 : 1));
            }
            sh.l2_f = sh.res_f[0];          // the best sampled point: where the direct search starts if the optimum is a wall
            sh.l2_p0 = sh.res_p0[0];
            sh.l2_p1 = sh.res_p1[0];
        }
        __syncthreads();

        // ---- F: L3, bounded quasi-Newton on the analytic gradient ---------------------------------------------------------
        // Newton state of start s lives in thread s (t < K2A_NS); the walks are spread over (start, segment) slots.
        // Phase 0 refines the L2 starts.  Phase 1 ("siblings"): a valley often holds a second minimum 10-30 deg of p1 away
        // (two branches of max(d) meet); two more starts, 20 deg of p1 either side of the phase-0 result with p0 re-scanned
        // there, either fall back into it (and are merged after an iteration or two) or find the sibling.
        const double p1_lo = p.p0_only ? 0.0 : -4000.0, p1_hi = p.p0_only ? 0.0 : 4000.0;
#ifndef XMR_K2A_PHASES
#define XMR_K2A_PHASES 2
#endif
        for (int phase = 0; phase < (p.p0_only ? 1 : XMR_K2A_PHASES); ++phase) {
        if (phase == 1) {
            if (t == 0) {
                int bs = 0;
                for (int s = 1; s < K2A_NS; ++s)
                    if (sh.res_f[s] < sh.res_f[bs]) bs = s;
                sh.res_f[0] = sh.res_f[bs]; sh.res_p0[0] = sh.res_p0[bs]; sh.res_p1[0] = sh.res_p1[bs];
                sh.res_h[0] = sh.res_h[bs]; sh.res_pen[0] = sh.res_pen[bs];
            }
            __syncthreads();
            for (int w = warp; w < 2; w += WPS) {
                const float c0c = float(sh.res_p0[0]);
                const float p1 = fminf(fmaxf(float(sh.res_p1[0]) + (w == 0 ? -K2A_SIBLING : K2A_SIBLING), -4000.f), 4000.f);
                float c0[K2_K], s0[K2_K], p0k[K2_K];
#pragma unroll
                for (int k = 0; k < K2_K; ++k) {
                    float q0 = c0c + (float(k) - 3.5f) * K2A_L2_DP0;
                    q0 = q0 > 180.f ? q0 - 360.f : (q0 < -180.f ? q0 + 360.f : q0);
                    p0k[k] = q0;
                    sincospif(q0 * (1.0f / 180.0f), &s0[k], &c0[k]);
                }
                Acc<float, METHOD_ACME, K2_K> acc;
                acc.init();
                lane_accumulate_rt<float, METHOD_ACME, K2_K>(sp, PADSHIFT, lane * L, (lane + 1) * L, geom, p1 * (1.0f / 360.0f), u0, duf, c0, s0, acc);
                acc.warp_reduce();
                float rbf = CUDART_INF_F, rb0 = c0c;
#pragma unroll
                for (int k = 0; k < K2_K; ++k) {
                    const float f = acc.score(k, geom);
                    if (f < rbf) { rbf = f; rb0 = p0k[k]; }
                }
                if (lane == 0) { sh.zf[w] = rbf; sh.zp0[w] = rb0; sh.zp1[w] = p1; }
            }
            __syncthreads();
            if (t == 0) {
                int any = 0;
                sh.active[0] = 0;
                for (int s = 1; s < K2A_NS; ++s) {
                    const int w = s - 1;
                    const bool ok = w < 2 && sh.res_f[0] < CUDART_INF && sh.zf[w] < CUDART_INF_F &&
                                    fabs(double(sh.zp1[w]) - sh.res_p1[0]) > 5.0;      // (clamped at the box edge: no sibling there)
                    sh.y0[s] = double(sh.zp0[w]);
                    sh.y1[s] = double(sh.zp1[w]);
                    sh.np[s] = 3;
                    sh.active[s] = ok ? 1 : 0;
                    sh.res_f[s] = ok ? double(sh.zf[w]) : CUDART_INF;
                    sh.res_p0[s] = sh.y0[s];
                    sh.res_p1[s] = sh.y1[s];
                    any |= sh.active[s];
                }
                sh.any_active = any;
                int nact = 0;
                for (int s = 0; s < K2A_NS; ++s) nact += sh.active[s];
                sh.gs = nact <= 1 ? GSMAX : (nact == 2 ? (GSMAX >= 2 ? GSMAX / 2 : 1) : (GSMAX >= 4 ? GSMAX / 4 : 1));
            }
            __syncthreads();
        }
        NewtonState st;
        st.x0 = st.x1 = 0.0; st.f = CUDART_INF; st.g0 = st.g1 = 0.0; st.done = 0; st.iters = 0;
        double h00 = 0.0, h01 = 0.0, h11 = 0.0;         // current Hessian model (per degree^2)
        double stp0 = 0.0, stp1 = 0.0;                  // last proposed step
        int rejects = 0, refreshed = 0;
        double res_h = 1.0, res_pen = 0.0;
        for (int nit = 0; nit < K2A_MAXIT && sh.any_active; ++nit) {
            // the warps are shared out among the ACTIVE starts: 8 warps -> 2 per start while three run, 4 while two, all 8 for
            // the last one (most voxels: the starts merge after two or three iterations)
            const int GS = sh.gs;
            int act[K2A_NS], nact = 0;
#pragma unroll
            for (int s = 0; s < K2A_NS; ++s)
                if (sh.active[s]) act[nact++] = s;
            for (int slotid = warp; slotid < nact * GS; slotid += WPS) {
                const int ai = slotid / GS, g = slotid - ai * GS;
                int s = act[0];
#pragma unroll
                for (int q = 1; q < K2A_NS; ++q)
                    if (q == ai) s = act[q];
                // lane chunk: the lane's L points split between the GS warps of this start (keeps the padded layout
                // conflict-free: lanes stay L points apart)
                const int m0 = lane * L + g * (L / GS), m1 = m0 + L / GS;
                const double y0 = sh.y0[s], y1 = sh.y1[s];
                const int np = sh.np[s];
                for (int k = 0; k < np; ++k) {
                    GradSums<float> a;
                    a.init();
                    const double q0 = y0 + (k == 1 ? NEWTON_H0 : 0.0), q1 = y1 + (k == 2 ? NEWTON_H1 : 0.0);
                    const double um0 = u0d + p.du * double(m0);
                    double ta = q0 / 360.0 + (q1 / 360.0) * um0;
                    ta -= floor(ta);
                    lane_grad<float, 16>(sp, PADSHIFT, m0, m1, N, float(ta), float(q1 / 360.0), float(-p.du * double(m0)), duf, -1, a);
                    // float32 only WITHIN the lane's short chunk: the chunk sums go to float64 before they are combined (where
                    // the penalty is large the gradient is a small difference of large sums; float32 lane-to-lane adds left the
                    // minimum 0.1-0.2 deg short along the flat valley)
                    GradSums<double> ad;
                    ad.P = a.P; ad.gP0 = a.gP0; ad.G2 = a.G2; ad.T2 = a.T2; ad.As0 = a.As0; ad.Al0 = a.Al0;
                    ad.gP1 = double(a.gP1) + um0 * double(a.gP0);      // u restarts at 0 inside the chunk: shift the u-weighted sums back
                    ad.As1 = double(a.As1) + um0 * double(a.As0);
                    ad.Al1 = double(a.Al1) + um0 * double(a.Al0);
                    ad.dmax = a.dmax; ad.qmax = a.qmax; ad.umax = double(a.umax) + um0;
                    ad.warp_reduce();
                    if (lane == 0) grad_store(ad, sh.part[s][g][k]);
                }
            }
            __syncthreads();
            if (t < K2A_NS && sh.active[t]) {
                const int s = t;
                const int np = sh.np[s];
                FG r[3];
                for (int k = 0; k < np; ++k) {
                    GradSums<double> tot;
                    tot.init();
                    for (int g = 0; g < sh.gs; ++g) tot.merge(grad_load(sh.part[s][g][k]));
                    r[k] = acme_finish(tot, N);
                }
                const double y0 = sh.y0[s], y1 = sh.y1[s];
                const bool first = (st.iters == 0);
                const bool small = fabs(stp0) < 0.05 && fabs(stp1) < 0.15;
                bool accept = false;
                if (np == 3 && !first && y0 == st.x0 && y1 == st.x1) {
                    // Hessian refresh at the current point
                    st.f = r[0].f; st.g0 = r[0].g0; st.g1 = r[0].g1;
                    accept = true;
                } else if (first || r[0].f < st.f || (small && r[0].f <= st.f * (1.0 + 1e-6))) {
                    if (!first && np == 1 && (fabs(y0 - st.x0) >= 0.5 || fabs(y1 - st.x1) >= 1.5)) {
                        // BFGS update of the Hessian model with s = y - x, v = g(y) - g(x)
                        const double s0 = y0 - st.x0, s1 = y1 - st.x1, v0 = r[0].g0 - st.g0, v1 = r[0].g1 - st.g1;
                        const double sv = s0 * v0 + s1 * v1;
                        const double hs0 = h00 * s0 + h01 * s1, hs1 = h01 * s0 + h11 * s1;
                        const double shs = s0 * hs0 + s1 * hs1;
                        if (sv > 0.0 && shs > 0.0) {
                            h00 += v0 * v0 / sv - hs0 * hs0 / shs;
                            h01 += v0 * v1 / sv - hs0 * hs1 / shs;
                            h11 += v1 * v1 / sv - hs1 * hs1 / shs;
                        }
                    }
                    st.x0 = y0; st.x1 = y1; st.f = r[0].f; st.g0 = r[0].g0; st.g1 = r[0].g1;
                    accept = true;
                }
                if (accept) {
                    if (np == 3) {
                        h00 = (r[1].g0 - r[0].g0) / NEWTON_H0;
                        h11 = (r[2].g1 - r[0].g1) / NEWTON_H1;
                        h01 = 0.5 * ((r[1].g1 - r[0].g1) / NEWTON_H0 + (r[2].g0 - r[0].g0) / NEWTON_H1);
                    }
                    ++st.iters;
                    rejects = 0;
                    // step from the model (acme_grad.cuh: newton_step takes offset gradients; here the model is explicit)
                    double s0, s1;
                    if (p.p0_only) {
                        s0 = h00 > 0.0 ? -st.g0 / h00 : (st.g0 > 0.0 ? -NEWTON_CAP0 : NEWTON_CAP0);
                        s1 = 0.0;
                    } else {
                        const double det = h00 * h11 - h01 * h01;
                        if (h00 > 0.0 && det > 1e-12 * h00 * h11) {
                            s0 = -(h11 * st.g0 - h01 * st.g1) / det;
                            s1 = -(h00 * st.g1 - h01 * st.g0) / det;
                        } else {
                            s0 = -st.g0 / fmax(fabs(h00), 1e-300);
                            s1 = -st.g1 / fmax(fabs(h11), 1e-300);
                        }
                    }
                    if (!(s0 == s0) || !(s1 == s1)) { s0 = 0.0; s1 = 0.0; }
                    const double sc = fmax(fmax(fabs(s0) / NEWTON_CAP0, fabs(s1) / NEWTON_CAP1), 1.0);
                    stp0 = s0 / sc;
                    stp1 = s1 / sc;
                    const double t1 = fmin(fmax(st.x1 + stp1, p1_lo), p1_hi);
                    stp1 = t1 - st.x1;
                    // gradient differences over small steps are float32 noise: BFGS models are only trusted for large steps;
                    // once the step is small every iterate gets its own secant Hessian (offsets 0.2 / 0.8 deg), and only a
                    // step computed from one may declare convergence
                    const bool smallstep = fabs(stp0) < 0.5 && fabs(stp1) < 1.5;
                    const bool conv = np == 3 && st.iters > 1 && fabs(stp0) < NEWTON_TOL0 && fabs(stp1) < NEWTON_TOL1;
                    // a Newton step from a fresh secant Hessian that is already tiny is taken WITHOUT re-evaluating there: the
                    // error after it is second order in the step (quadratic convergence), far below the 0.1 deg tolerance
                    const bool last = np == 3 && st.iters > 1 && !conv && fabs(stp0) < 0.02 && fabs(stp1) < 0.06;
                    res_h = r[0].H; res_pen = r[0].pen;
                    if (last) {
                        st.x0 += stp0;
                        st.x1 = t1;
                        sh.active[s] = 0;
                    } else if (!(st.f < CUDART_INF) || conv) {
                        sh.active[s] = 0;
                    } else {
                        sh.y0[s] = st.x0 + stp0;
                        sh.y1[s] = t1;
                        sh.np[s] = smallstep ? 3 : 1;
                    }
                } else {
                    ++rejects;
                    if (rejects <= 2) {
                        stp0 /= 3.0; stp1 /= 3.0;
                        sh.y0[s] = st.x0 + stp0;
                        sh.y1[s] = st.x1 + stp1;
                        sh.np[s] = 1;
                    } else if (!refreshed) {
                        refreshed = 1;                  // the model is off (a kink was crossed): one fresh secant Hessian at x
                        rejects = 0;
                        sh.y0[s] = st.x0;
                        sh.y1[s] = st.x1;
                        sh.np[s] = 3;
                    } else {
                        sh.active[s] = 0;
                    }
                }
                // the start's result is its ACCEPTED iterate: near convergence the objective changes by less than its float32
                // noise, so "record only if f decreased" would freeze the answer one or two (0.1 deg) steps early
                if (st.iters >= 1) {
                    sh.res_f[s] = st.f; sh.res_p0[s] = st.x0; sh.res_p1[s] = st.x1;
                    sh.res_h[s] = res_h; sh.res_pen[s] = res_pen;
                }
            }
            __syncthreads();
            if (t == 0) {
                // two starts that have walked into the same minimum: the worse one stops (most voxels have ONE valley and all
                // three starts end in it)
                for (int a = 0; a < K2A_NS; ++a)
                    for (int b = a + 1; b < K2A_NS; ++b) {
                        if (!(sh.res_f[a] < CUDART_INF) || !(sh.res_f[b] < CUDART_INF)) continue;
                        if (!sh.active[a] && !sh.active[b]) continue;
                        double d0 = fabs(sh.res_p0[a] - sh.res_p0[b]);
                        d0 = fmin(d0, fabs(360.0 - d0));
                        if (d0 < 0.5 && fabs(sh.res_p1[a] - sh.res_p1[b]) < 2.0) {
                            const int worse = (sh.res_f[a] <= sh.res_f[b]) ? b : a;
                            if (sh.active[worse]) sh.active[worse] = 0;
                            else sh.active[worse == a ? b : a] = sh.active[worse == a ? b : a];   // the finished one is the better: keep going
                        }
                    }
                int any = 0, nact = 0;
                for (int s = 0; s < K2A_NS; ++s) { any |= sh.active[s]; nact += sh.active[s]; }
                sh.any_active = any;
                sh.gs = nact <= 1 ? GSMAX : (nact == 2 ? (GSMAX >= 2 ? GSMAX / 2 : 1) : (GSMAX >= 4 ? GSMAX / 4 : 1));
            }
            __syncthreads();
        }
        }   // phase
        double fin_f = CUDART_INF, fin_p0 = 0.0, fin_p1 = 0.0;
        {
            int bs = 0;
            for (int s = 1; s < K2A_NS; ++s)
                if (sh.res_f[s] < sh.res_f[bs]) bs = s;
            fin_f = sh.res_f[bs]; fin_p0 = sh.res_p0[bs]; fin_p1 = sh.res_p1[bs];
            // Where the penalty vanishes at the optimum (clean all-positive spectra, or data of amplitude >> 1 where 1000 P is
            // a wall) the minimum is CONSTRAINED: the entropy term is minimised against the wall P > 0 and the gradient does
            // not vanish there.  Newton stalls at the wall; these voxels (a few per cent) are finished by direct search.
            if (t == 0) sh.wall = (fin_f < CUDART_INF && sh.res_pen[bs] < 4.0 * sh.res_h[bs]) ? 1 : 0;
        }
        __syncthreads();
        if (sh.wall) {
            // 8 x 8 zoom rounds on the full spectrum (the round-1 search) from the best sampled point of L2
            if (t == 0) { sh.zc0 = float(sh.l2_p0); sh.zc1 = float(sh.l2_p1); sh.zcf = CUDART_INF_F; }
            __syncthreads();
            float zh0 = 10.f, zh1 = p.p0_only ? 0.f : 22.5f;
            for (int round = 0; round < 11; ++round) {
                const float c0c = sh.zc0, c1c = sh.zc1;
                for (int row = warp; row < 8; row += WPS) {
                    if (p.p0_only && row > 0) { if (lane == 0) sh.zf[row] = CUDART_INF_F; continue; }
                    const float p1 = p.p0_only ? 0.f : fminf(fmaxf(c1c + (float(2 * row) - 7.f) * (1.f / 7.f) * zh1, -4000.f), 4000.f);
                    float c0[K2_K], s0[K2_K], p0k[K2_K];
#pragma unroll
                    for (int k = 0; k < K2_K; ++k) {
                        float q0 = c0c + (float(2 * k) - 7.f) * (1.f / 7.f) * zh0;
                        q0 = q0 > 180.f ? q0 - 360.f : (q0 < -180.f ? q0 + 360.f : q0);
                        p0k[k] = q0;
                        sincospif(q0 * (1.0f / 180.0f), &s0[k], &c0[k]);
                    }
                    Acc<float, METHOD_ACME, K2_K> acc;
                    acc.init();
                    lane_accumulate_rt<float, METHOD_ACME, K2_K>(sp, PADSHIFT, lane * L, (lane + 1) * L, geom, p1 * (1.0f / 360.0f), u0, duf, c0, s0, acc);
                    acc.warp_reduce();
                    float rbf = CUDART_INF_F, rb0 = c0c;
#pragma unroll
                    for (int k = 0; k < K2_K; ++k) {
                        const float f = acc.score(k, geom);
                        if (f < rbf) { rbf = f; rb0 = p0k[k]; }
                    }
                    if (lane == 0) { sh.zf[row] = rbf; sh.zp0[row] = rb0; sh.zp1[row] = p1; }
                }
                __syncthreads();
                if (t == 0) {
                    for (int row = 0; row < 8; ++row)
                        if (sh.zf[row] < sh.zcf) { sh.zcf = sh.zf[row]; sh.zc0 = sh.zp0[row]; sh.zc1 = sh.zp1[row]; }
                }
                zh0 *= 0.45f;
                zh1 *= 0.45f;
                __syncthreads();
            }
            if (double(sh.zcf) < fin_f) { fin_f = double(sh.zcf); fin_p0 = double(sh.zc0); fin_p1 = double(sh.zc1); }
        }
        {
            double w = fmod(fin_p0 + 180.0, 360.0);
            if (w < 0) w += 360.0;
            fin_p0 = w - 180.0;
        }

        // ---- G: apply the phase and store ----------------------------------------------------------------------------------
        {
            const double a_turns = fin_p0 / 360.0 + (fin_p1 / 360.0) * u0d;
            const double b_turns = (fin_p1 / 360.0) * p.du;
            float2* dst = p.out + vox * (long long)N;
            for (int m = t; m < N; m += C::T) {
                double turns = a_turns + b_turns * double(m);
                turns -= floor(turns);
                float sn, cs;
                sincospif(2.0f * float(turns), &sn, &cs);
                st_stream(dst + m, cmul(sp[m + (m >> PADSHIFT)], make_float2(cs, sn)));
            }
            if (t == 0) {
                p.p0_out[vox] = fin_p0;
                p.p1_out[vox] = p.p0_only ? 0.0 : fin_p1;
                p.pivot_out[vox] = mstar;
                p.fun_out[vox] = float(fin_f);
            }
        }
        __syncthreads();   // sp (exchange B) and the shared search state are rewritten by the next voxel
        if (IPB && p.use_tma && t == 0) {                      // one buffer: the next FID can only land now
            const long long nv = vox + gridDim.x;
            if (nv < p.batch) {
                fence_proxy_async_smem();
                issue(nv);
            }
        }
    }
}

}  // namespace xmr
