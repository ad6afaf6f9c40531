// K2: the whole chain for one voxel in one pass -- zero-fill -> window -> FFT -> fftshift -> per-spectrum
// (p0, p1) search on the shared-memory resident spectrum -> phase -> store.   [autophase mode="all"]
//
// One CTA (T = N/E threads) owns one spectrum at a time (persistent loop over voxels):
//   A  FFT exactly as K1 (TMA bulk load of the FID row, three register/shared stages)
//   B  |S| argmax -> the voxel's pivot (the reference, called on a 1-D spectrum, pivots on that spectrum's own
//      maximum: phasing.py:229-238)
//   C  spectrum -> padded shared layout (reusing exchange buffer B) + a stride-4 subsample of point pairs
//   D  coarse grid (p0 step 15 deg x p1 step 45 deg over the reference's box) of the reference's objective, ACME on
//      the pair subsample (each warp: one p1, eight p0 per walk), local methods on their ROI
//   E  the NSTART best, mutually separated cells are refined by NROUND rounds of an 8x8 zoom on the FULL
//      spectrum (each warp one p1 row, eight p0), shrinking the window ~3.1x per round, pruning starts as it goes
//   F  out[m] = S[m] * exp(i*(p0 + p1*u_m)) and (p0, p1, pivot, objective) per voxel
// HBM traffic per voxel is still 8*n_in + 8*n_out (+ 24 B of results); the search makes this kernel SFU/FP32-bound.
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

#include "autophase_eval.cuh"
#include "fft_stages.cuh"
#include "k1_fft.cuh"
#include "ptx_sm100.cuh"

namespace xmr {

constexpr int K2_SUB = 4;        // coarse stage: every 4th point pair
constexpr int K2_K = 8;          // p0 candidates per walk
constexpr int K2_NP0 = 24;       // coarse p0: -180 + 15*k
constexpr int K2_NP1 = 179;      // coarse p1: -4000 + 45*k (last clamped to 4000)
#ifndef XMR_K2_NSTART
#define XMR_K2_NSTART 4
#endif
constexpr int K2_NSTART = XMR_K2_NSTART;
constexpr int K2_NROUND = 7;
#ifndef XMR_K2_SHRINK
#define XMR_K2_SHRINK 0.32f
#endif
constexpr float K2_SHRINK = XMR_K2_SHRINK;   // window half-width ratio between zoom rounds (8 points per axis: spacing = 2h/7)
constexpr int K2_NSHORT = 64;     // coarse cells re-evaluated at the finer subsample
constexpr int K2_NP0_ONLY = 121; // p0_only coarse: -180 + 3*k
constexpr int K2_PPL_SHIFT = 5;   // log2(pairs per lane) for N = 4096; other N: any padding works, 4096 is conflict-free

struct K2Params {
    const float2* in;
    float2* out;
    long long batch;
    int n_in;
    int pad_left;
    int out_shift;
    int use_tma;
    float scale;
    const float2* twN;
    const float* win;
    float win_rows[32];
    int win_table;         // 1: p.win is a full table; 0: separable / scale only
    int spec_in;           // 1: `in` already holds spectra [batch, N] (stored order): skip the transform
    // search geometry: u_m = u0 + du*m over the STORED index m
    double du;
    int fixed_pivot;       // target_coord given: u0 = u0_fixed, ROI centre = fixed_target for every voxel
    double u0_fixed;
    int fixed_target;
    int index_width;
    int p0_only;
    double* p0_out;
    double* p1_out;
    int* pivot_out;
    float* fun_out;
};

template <int N>
struct K2Smem {
    using C = FftCfg<N>;
    static constexpr int PADSHIFT = ilog2(N / 32 > 0 ? N / 32 : 1);
    static constexpr size_t SLOT = size_t(C::N) * sizeof(float2);
    static constexpr size_t SPN = (size_t(C::N) + (size_t(C::N) >> PADSHIFT) + 2);
    static constexpr size_t B = (C::SIZE_B > SPN ? size_t(C::SIZE_B) : SPN) * sizeof(float2);
    static constexpr size_t SUBP = size_t(C::N / K2_SUB + (C::N / K2_SUB >> K2_PPL_SHIFT) + 2) * sizeof(float4);
    static constexpr size_t CELLS = size_t(K2_NP1) * (K2_NP0 / K2_K);         // (p1, chunk) cells
    static constexpr size_t F = ((CELLS > 32 ? CELLS : 32) * 8 + 15) / 16 * 16; // {float f, int k} per cell
    static constexpr size_t MISC = 2048;
    static constexpr size_t TOTAL = SLOT + B + SUBP + F + MISC;
};

struct K2Start {
    float f;
    float p0, p1;
};

// ACME on the stride-SUB pair subsample: pair j holds (S[SUB*j], S[SUB*j+1]).
template <int K, int STRIDE = 1>
__device__ __forceinline__ void lane_accumulate_pairs(const float4* __restrict__ sub, int j0, int j1, float tpu,
                                                      float u0, float du, const float (&c0)[K], const float (&s0)[K],
                                                      Acc<float, METHOD_ACME, K>& acc) {
    if (j0 >= j1) return;
    float sr, cr, s1, c1, ss, cs;
    {
        float t = tpu * (u0 + du * float(K2_SUB * j0));
        t -= floorf(t);
        sincospif(2.0f * t, &sr, &cr);
        float ti = tpu * du;
        ti -= floorf(ti);
        sincospif(2.0f * ti, &s1, &c1);
        float ts = tpu * du * float(K2_SUB * STRIDE);
        ts -= floorf(ts);
        sincospif(2.0f * ts, &ss, &cs);
    }
#pragma unroll 2
    for (int j = j0; j < j1; j += STRIDE) {
        const float4 q = sub[j + (j >> K2_PPL_SHIFT)];                           // one pad element per lane chunk
        const float ax = q.x * cr - q.y * sr, ay = q.x * sr + q.y * cr;           // w at m
        const float r1c = cr * c1 - sr * s1, r1s = cr * s1 + sr * c1;            // rotation at m+1
        const float ex = (q.z * r1c - q.w * r1s) - ax, ey = (q.z * r1s + q.w * r1c) - ay;   // w(m+1) - w(m)
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float d0 = ax * c0[k] - ay * s0[k];
            const float D = fabsf(ex * c0[k] - ey * s0[k]);                      // |d(m+1) - d(m)|
            acc.a[k][0] += D;
            acc.a[k][1] += D * RealOps<float>::log2r(fmaxf(D, 1.17549435e-38f));
            const float neg = fminf(d0, 0.f);
            acc.a[k][2] += neg * neg;
            acc.a[k][3] = fmaxf(acc.a[k][3], d0);
        }
        const float ncr = cr * cs - sr * ss;
        sr = cr * ss + sr * cs;
        cr = ncr;
    }
}

template <int N, int METHOD>
__global__ void __launch_bounds__(FftCfg<N>::T, (N >= 8192 ? 1 : (FftCfg<N>::T >= 256 ? 2 : (FftCfg<N>::T >= 128 ? 4 : 8))))
k2_kernel(const __grid_constant__ K2Params p) {
    using C = FftCfg<N>;
    using SM = K2Smem<N>;
    constexpr bool TW_PERSIST = false;   // registers are better spent on the search accumulators here
    constexpr int NTW = 1;
    constexpr int WPS = C::T / 32;                    // warps per spectrum
    constexpr int PADSHIFT = SM::PADSHIFT;
    constexpr int L = N / 32;                         // full-resolution points per lane
    static_assert(C::T >= 32, "per-voxel kernel needs N >= 512");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* slot = reinterpret_cast<float2*>(smem_raw);
    float2* Bbuf = reinterpret_cast<float2*>(smem_raw + SM::SLOT);
    float2* sp = Bbuf;                                                     // padded spectrum reuses exchange B
    float4* subp = reinterpret_cast<float4*>(smem_raw + SM::SLOT + SM::B);
    float* cellf = reinterpret_cast<float*>(smem_raw + SM::SLOT + SM::B + SM::SUBP);
    int* cellk = reinterpret_cast<int*>(cellf + (SM::CELLS > 32 ? SM::CELLS : 32));
    unsigned char* misc = smem_raw + SM::SLOT + SM::B + SM::SUBP + SM::F;
    uint64_t* bar = reinterpret_cast<uint64_t*>(misc);                     // 8 B
    float* redv = reinterpret_cast<float*>(misc + 16);                     // [32]
    int* redi = reinterpret_cast<int*>(misc + 16 + 128);                   // [32]
    K2Start* starts = reinterpret_cast<K2Start*>(misc + 16 + 256);         // [NSTART]
    float* rowf = reinterpret_cast<float*>(misc + 16 + 256 + 128);         // scratch base
    float* st_p0 = rowf + 32;                                              // [NSTART] per-start centre / value / flag
    float* st_p1 = st_p0 + 8;
    float* st_f = st_p0 + 16;
    int* st_on = reinterpret_cast<int*>(st_p0 + 24);
    float* rf = st_p0 + 32;                                                // [NSTART*8] per-(start,row) results
    float* rp0 = rf + K2_NSTART * 8;
    float* rp1 = rp0 + K2_NSTART * 8;
    int* shortlist = reinterpret_cast<int*>(rp1 + K2_NSTART * 8);           // [K2_NSHORT]

    const int t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;
    const int n_in = p.n_in;
    const bool need_load_barrier = (p.pad_left != 0);

    float2 tw_persist[TW_PERSIST ? NTW : 1];
    float2 tw0_base[C::C0 * 2], tw1_base[C::C1 * 2];
    init_twiddles<C, false>(t, p.twN, TW_PERSIST ? tw_persist : nullptr, tw0_base, tw1_base);
    float wcol[C::C0];
#pragma unroll
    for (int j = 0; j < C::C0; ++j) wcol[j] = (p.win && !p.win_table) ? p.win[t + C::T * j] : p.scale;

    auto issue = [&](long long v) {
        const uint32_t row_bytes = uint32_t(n_in) * 8u;
        mbar_arrive_expect_tx(bar, row_bytes);
        bulk_g2s(slot, p.in + v * n_in, row_bytes, bar);
    };
    if (p.use_tma) {
        if (t == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (t == 0 && (long long)blockIdx.x < p.batch) issue(blockIdx.x);
    }

    ScoreGeom geom;
    geom.n = N;
    const float duf = float(p.du);

    int it = 0;
    for (long long vox = blockIdx.x; vox < p.batch; vox += gridDim.x, ++it) {
        // ---- A: FFT --------------------------------------------------------------------------------------------
        if (p.use_tma) {
            mbar_wait(bar, it & 1);
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
            stage1<C, false>(t, slot, Bbuf, tw1_base);
            __syncthreads();
            if (p.use_tma && t == 0) {
                const long long nv = vox + gridDim.x;
                if (nv < p.batch) {
                    fence_proxy_async_smem();
                    issue(nv);
                }
            }
            stage2<C, false>(t, Bbuf, v);

            // ---- B: |S| argmax of this spectrum ----------------------------------------------------------------
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
            if (lane == 0) { redv[warp] = best; redi[warp] = besti; }
            __syncthreads();                                   // also: every thread is done reading exchange B
            // ---- C: spectrum -> padded shared layout -----------------------------------------------------------
#pragma unroll
            for (int j = 0; j < C::C2; ++j)
#pragma unroll
                for (int d = 0; d < C::R2; ++d) {
                    const int m = (t + C::T * j + Q * d + p.out_shift) & (N - 1);
                    sp[m + (m >> PADSHIFT)] = v[j * C::R2 + d];
                }
            best = redv[0];
            besti = redi[0];
            for (int w = 1; w < WPS; ++w) amax_combine(best, besti, redv[w], redi[w]);
            mstar = besti;
        } else {
            // spectra given: copy the row into the padded layout and find its maximum
            float best = -1.f;
            int besti = 0x7fffffff;
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
            if (lane == 0) { redv[warp] = best; redi[warp] = besti; }
            __syncthreads();
            if (p.use_tma && t == 0) {
                const long long nv = vox + gridDim.x;
                if (nv < p.batch) {
                    fence_proxy_async_smem();
                    issue(nv);
                }
            }
            best = redv[0];
            besti = redi[0];
            for (int w = 1; w < WPS; ++w) amax_combine(best, besti, redv[w], redi[w]);
            mstar = besti;
        }
        __syncthreads();
        if (METHOD == METHOD_ACME) {
            for (int j = t; j < N / K2_SUB; j += C::T) {
                const int m = K2_SUB * j;
                const float2 a = sp[m + (m >> PADSHIFT)], b = sp[m + 1 + ((m + 1) >> PADSHIFT)];
                subp[j + (j >> K2_PPL_SHIFT)] = make_float4(a.x, a.y, b.x, b.y);
            }
        }
        const int target = p.fixed_pivot ? p.fixed_target : mstar;
        const float u0 = p.fixed_pivot ? float(p.u0_fixed) : float(-p.du * double(mstar));
        geom.target_idx = target;
        geom.roi_start = max(0, target - p.index_width);
        geom.roi_end = min(N, target + p.index_width);
        const float2 Spiv = sp[mstar + (mstar >> PADSHIFT)];   // global max point: candidate for max(d) in the coarse stage
        const float upiv = u0 + duf * float(mstar);
        __syncthreads();

        // ---- D: coarse grid ------------------------------------------------------------------------------------------
        const int np0 = p.p0_only ? K2_NP0_ONLY : K2_NP0;
        const float p0step = p.p0_only ? 3.0f : 15.0f;
        const int nchunk = (np0 + K2_K - 1) / K2_K;
        const int np1 = p.p0_only ? 1 : K2_NP1;
        const int ncell = np1 * nchunk;
        // evaluates one (p1, chunk of 8 p0) cell on the pair subsample (ACME; every STRIDE-th pair) or on the ROI
        auto eval_cell = [&](int cell, auto stride_tag) {
            constexpr int STRIDE = decltype(stride_tag)::value;
            const int i1 = cell / nchunk, ch = cell - i1 * nchunk;
            const float p1 = p.p0_only ? 0.f : fminf(-4000.f + 45.f * float(i1), 4000.f);
            const float tpu = p1 * (1.0f / 360.0f);
            float c0[K2_K], s0[K2_K];
#pragma unroll
            for (int k = 0; k < K2_K; ++k) {
                const float p0 = fminf(-180.f + p0step * float(ch * K2_K + k), 180.f);
                sincospif(p0 * (1.0f / 180.0f), &s0[k], &c0[k]);
            }
            Acc<float, METHOD, K2_K> acc;
            acc.init();
            if constexpr (METHOD == METHOD_ACME) {
                constexpr int PPL = (N / K2_SUB) / 32;   // pairs per lane
                lane_accumulate_pairs<K2_K, STRIDE>(subp, lane * PPL, (lane + 1) * PPL, tpu, u0, duf, c0, s0, acc);
            } else {
                lane_accumulate_rt<float, METHOD, K2_K>(sp, PADSHIFT, lane * L, (lane + 1) * L, geom, tpu, u0, duf, c0, s0, acc);
            }
            acc.warp_reduce();
            float bf = CUDART_INF_F;
            int bk = 0;
            if (METHOD == METHOD_ACME) {
                // the global-maximum point always takes part in max(d): the subsample may step over the peak
                float tp = tpu * upiv;
                tp -= floorf(tp);
                float sn, cs;
                sincospif(2.0f * tp, &sn, &cs);
                const float wx = Spiv.x * cs - Spiv.y * sn, wy = Spiv.x * sn + Spiv.y * cs;
#pragma unroll
                for (int k = 0; k < K2_K; ++k) {
                    acc.a[k][0] *= float(K2_SUB * STRIDE);
                    acc.a[k][1] *= float(K2_SUB * STRIDE);
                    acc.a[k][2] *= float(K2_SUB * STRIDE);
                    acc.a[k][3] = fmaxf(acc.a[k][3], wx * c0[k] - wy * s0[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < K2_K; ++k) {
                const float f = acc.score(k, geom);
                if (ch * K2_K + k < np0 && f < bf) { bf = f; bk = k; }
            }
            if (lane == 0) { cellf[cell] = bf; cellk[cell] = bk; }
        };
        const bool two_level = (METHOD == METHOD_ACME) && (ncell > K2_NSHORT);
        if (two_level) {
            // level 1: every 16th point pair on all cells; level 2: the K2_NSHORT best cells again on every 4th pair
            for (int cell = warp; cell < ncell; cell += WPS) eval_cell(cell, std::integral_constant<int, 4>{});
            __syncthreads();
            if (warp == 0) {
                for (int i = 0; i < K2_NSHORT; ++i) {
                    float bf = CUDART_INF_F;
                    int bc = 0x7fffffff;
                    for (int c = lane; c < ncell; c += 32) {
                        const float f = cellf[c];
                        if (f < bf || (f == bf && c < bc)) { bf = f; bc = c; }
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        const float of = __shfl_xor_sync(0xffffffffu, bf, off);
                        const int oc = __shfl_xor_sync(0xffffffffu, bc, off);
                        if (of < bf || (of == bf && oc < bc)) { bf = of; bc = oc; }
                    }
                    if (bc == 0x7fffffff) bc = -1;      // fewer finite cells than the short list
                    if (lane == 0) {
                        shortlist[i] = bc;
                        if (bc >= 0) cellf[bc] = CUDART_INF_F;
                    }
                    __syncwarp();
                }
                for (int c = lane; c < ncell; c += 32) cellf[c] = CUDART_INF_F;   // only re-evaluated cells compete
            }
            __syncthreads();
            for (int i = warp; i < K2_NSHORT; i += WPS) {
                const int cell = shortlist[i];
                if (cell >= 0) eval_cell(cell, std::integral_constant<int, 1>{});
            }
        } else {
            for (int cell = warp; cell < ncell; cell += WPS) eval_cell(cell, std::integral_constant<int, 1>{});
        }
        __syncthreads();

        // ---- E: NSTART separated minima (warp 0), then zoom refinement on the full spectrum -----------------------------
        if (warp == 0) {
            for (int s = 0; s < K2_NSTART; ++s) {
                float bf = CUDART_INF_F;
                int bc = 0x7fffffff;
                for (int c = lane; c < ncell; c += 32) {
                    const float f = cellf[c];
                    if (f < bf || (f == bf && c < bc)) { bf = f; bc = c; }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const float of = __shfl_xor_sync(0xffffffffu, bf, off);
                    const int oc = __shfl_xor_sync(0xffffffffu, bc, off);
                    if (of < bf || (of == bf && oc < bc)) { bf = of; bc = oc; }
                }
                if (bc == 0x7fffffff) bc = 0;
                const int i1 = bc / nchunk, ch = bc - i1 * nchunk;
                if (lane == 0) {
                    starts[s].f = bf;
                    starts[s].p0 = fminf(-180.f + p0step * float(ch * K2_K + cellk[bc]), 180.f);
                    starts[s].p1 = p.p0_only ? 0.f : fminf(-4000.f + 45.f * float(i1), 4000.f);
                }
                __syncwarp();
                // suppress the neighbourhood: +-2 cells in p1, the chunk itself and its neighbours (p0 is periodic)
                for (int c = lane; c < ncell; c += 32) {
                    const int j1 = c / nchunk, jc = c - j1 * nchunk;
                    int dc = abs(jc - ch);
                    dc = min(dc, nchunk - dc);
                    if (abs(j1 - i1) <= 2 && dc <= 1) cellf[c] = CUDART_INF_F;
                }
                __syncwarp();
            }
        }
        __syncthreads();

        // Zoom refinement with successive pruning.  Every round evaluates, for each active start, an 8x8 window
        // (8 p1 rows x 8 p0) around its current centre; rows of all active starts are spread over the warps.
        //   rounds 0-1: all NSTART (4) starts     (window +-15 x +-45 deg -> +-1.5 x +-4.6 deg)
        //   rounds 2-3: the best 2 distinct starts
        //   rounds 4-6: the best one              (final spacing 0.005 x 0.014 deg)
        // (tools/validate_pervoxel.py: 6/3/2 starts give the same quality within noise at 1.3x the cost)
        // All rounds use the FULL spectrum: the pair subsample only localises basins -- its noise-induced fine structure
        // (local minima every ~25 deg of p1) differs from the full objective's, so it must not steer the refinement.
        if (t < K2_NSTART) {
            st_p0[t] = starts[t].p0;
            st_p1[t] = starts[t].p1;
            st_f[t] = CUDART_INF_F;
            st_on[t] = (starts[t].f < CUDART_INF_F) ? 1 : 0;
        }
        __syncthreads();
        float h0 = p0step, h1 = p.p0_only ? 0.f : 45.f;
        for (int round = 0; round < K2_NROUND; ++round) {
            const bool use_pairs = false;
            if (round == 2 || round == 4) {
                // prune: keep the best `keep` starts that are not duplicates of a better one
                if (t == 0) {
#ifndef XMR_K2_KEEP_A
#define XMR_K2_KEEP_A 2
#endif
#ifndef XMR_K2_KEEP_B
#define XMR_K2_KEEP_B 1
#endif
                    const int keep = (round == 2) ? XMR_K2_KEEP_A : XMR_K2_KEEP_B;
                    int kept = 0;
                    bool used[K2_NSTART];
                    for (int i = 0; i < K2_NSTART; ++i) used[i] = false;
                    int order[K2_NSTART];
                    for (int r = 0; r < K2_NSTART; ++r) {
                        int bi = -1;
                        for (int i = 0; i < K2_NSTART; ++i)
                            if (!used[i] && st_on[i] && (bi < 0 || st_f[i] < st_f[bi])) bi = i;
                        order[r] = bi;
                        if (bi >= 0) used[bi] = true;
                    }
                    for (int r = 0; r < K2_NSTART; ++r) {
                        const int i = order[r];
                        if (i < 0) continue;
                        bool dup = false;
                        for (int q = 0; q < r; ++q) {
                            const int j = order[q];
                            if (j >= 0 && st_on[j] && fabsf(st_p0[i] - st_p0[j]) < 4.f * h0 && fabsf(st_p1[i] - st_p1[j]) < 4.f * h1)
                                dup = true;
                        }
                        if (dup || kept >= keep) st_on[i] = 0;
                        else ++kept;
                    }
                }
                __syncthreads();
            }
            for (int r = warp; r < K2_NSTART * 8; r += WPS) {
                const int s = r >> 3, row = r & 7;
                if (!st_on[s]) continue;
                if (p.p0_only && row > 0) { if (lane == 0) rf[r] = CUDART_INF_F; continue; }
                const float c0c = st_p0[s], c1c = st_p1[s];
                const float p1 = p.p0_only ? 0.f : fminf(fmaxf(c1c + (float(2 * row) - 7.f) * (1.f / 7.f) * h1, -4000.f), 4000.f);
                const float tpu = p1 * (1.0f / 360.0f);
                float c0[K2_K], s0[K2_K], p0k[K2_K];
#pragma unroll
                for (int k = 0; k < K2_K; ++k) {
                    // p0 is periodic: candidates leaving the closed box [-180, 180] re-enter on the other side
                    float q0 = c0c + (float(2 * k) - 7.f) * (1.f / 7.f) * h0;
                    q0 = q0 > 180.f ? q0 - 360.f : (q0 < -180.f ? q0 + 360.f : q0);
                    p0k[k] = q0;
                    sincospif(q0 * (1.0f / 180.0f), &s0[k], &c0[k]);
                }
                Acc<float, METHOD, K2_K> acc;
                acc.init();
                if constexpr (METHOD == METHOD_ACME) {
                    if (use_pairs) {
                        constexpr int PPL = (N / K2_SUB) / 32;
                        lane_accumulate_pairs<K2_K>(subp, lane * PPL, (lane + 1) * PPL, tpu, u0, duf, c0, s0, acc);
                    } else {
                        lane_accumulate_rt<float, METHOD, K2_K>(sp, PADSHIFT, lane * L, (lane + 1) * L, geom, tpu, u0, duf, c0, s0, acc);
                    }
                } else {
                    lane_accumulate_rt<float, METHOD, K2_K>(sp, PADSHIFT, lane * L, (lane + 1) * L, geom, tpu, u0, duf, c0, s0, acc);
                }
                acc.warp_reduce();
                if constexpr (METHOD == METHOD_ACME) {
                    if (use_pairs) {
                        float tp = tpu * upiv;
                        tp -= floorf(tp);
                        float sn, cs;
                        sincospif(2.0f * tp, &sn, &cs);
                        const float wx = Spiv.x * cs - Spiv.y * sn, wy = Spiv.x * sn + Spiv.y * cs;
#pragma unroll
                        for (int k = 0; k < K2_K; ++k) {
                            acc.a[k][0] *= float(K2_SUB);
                            acc.a[k][1] *= float(K2_SUB);
                            acc.a[k][2] *= float(K2_SUB);
                            acc.a[k][3] = fmaxf(acc.a[k][3], wx * c0[k] - wy * s0[k]);
                        }
                    }
                }
                float rbf = CUDART_INF_F, rb0 = c0c;
#pragma unroll
                for (int k = 0; k < K2_K; ++k) {
                    const float f = acc.score(k, geom);
                    if (f < rbf) { rbf = f; rb0 = p0k[k]; }
                }
                if (lane == 0) { rf[r] = rbf; rp0[r] = rb0; rp1[r] = p1; }
            }
            __syncthreads();
            if (t < K2_NSTART && st_on[t]) {
                float bf = st_f[t], b0 = st_p0[t], b1 = st_p1[t];
                for (int row = 0; row < 8; ++row) {
                    const int r = t * 8 + row;
                    if (rf[r] < bf) { bf = rf[r]; b0 = rp0[r]; b1 = rp1[r]; }
                }
                st_f[t] = bf; st_p0[t] = b0; st_p1[t] = b1;
            }
            h0 *= K2_SHRINK;
            h1 *= K2_SHRINK;
            __syncthreads();
        }
        float fin_f = CUDART_INF_F, fin_p0 = 0.f, fin_p1 = 0.f;
        for (int s = 0; s < K2_NSTART; ++s)
            if (st_on[s] && st_f[s] < fin_f) { fin_f = st_f[s]; fin_p0 = st_p0[s]; fin_p1 = st_p1[s]; }

        // ---- F: apply the phase and store --------------------------------------------------------------------------------
        {
            const double a_turns = double(fin_p0) / 360.0 + (double(fin_p1) / 360.0) * double(u0);
            const double b_turns = (double(fin_p1) / 360.0) * p.du;
            float2* dst = p.out + vox * (long long)N;
            for (int m = t; m < N; m += C::T) {
                double turns = a_turns + b_turns * double(m);
                turns -= floor(turns);
                float sn, cs;
                sincospif(2.0f * float(turns), &sn, &cs);
                st_stream(dst + m, cmul(sp[m + (m >> PADSHIFT)], make_float2(cs, sn)));
            }
            if (t == 0) {
                p.p0_out[vox] = double(fin_p0);
                p.p1_out[vox] = p.p0_only ? 0.0 : double(fin_p1);
                p.pivot_out[vox] = mstar;
                p.fun_out[vox] = fin_f;
            }
        }
        __syncthreads();   // sp (exchange B) is rewritten by the next voxel's stage 1
    }
}

}  // namespace xmr
