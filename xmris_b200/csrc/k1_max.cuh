// K1-max: pass 1 of autophase(mode="single") -- per-spectrum max |S| for the global argmax (phasing.py:229-231), with
// branch and bound.  Specialised for full-length input or input zero-filled 2x / 4x at the end, separable window,
// fftshift-free statistics, N in [512, 8192] (N = 8192: 32 points per thread, one CTA per SM, stage-0 twiddles from a power chain).
//
// Two upper bounds on every output of a spectrum, each far cheaper than finishing the transform:
//   level 0 (no transform):                          |X| <= sum_n |x_n| |w_n|                      (triangle inequality)
//   level 1 (after two FFT stages, 16 terms left):   |X|^2 <= 16 * sum_b max_c |Z[k1][c][b]|^2     (Cauchy-Schwarz)
// On decaying multi-line FIDs level 0 is within ~1.2x of the spectrum's own maximum (a Cauchy-Schwarz bound after one
// radix-16 stage: ~1.5x, at ten times the instructions) and level 1 within ~1.02x.  If a bound is below the running
// global maximum the spectrum cannot hold the global argmax and the rest of its work is skipped: ~3/4 of the voxels of
// MRSI-like data cost one pass over their samples (4 instructions per point, HBM-bound), most of the others two FFT
// stages, < 1 % a full transform.  The bounds never under-estimate (float32 rounding is covered by the 1.0001 margin)
// and the running maximum never exceeds the true one, so the global maximum and its first row are exact and
// run-independent.
//
// Exchange B aliases the landing slot (no separate buffer), which pays for a 3-deep TMA ring at 2 CTAs/SM.
#pragma once
#include "k1_fft.cuh"

namespace xmr {

constexpr int K1MAX_STAGES = 3;

template <int N>
struct K1MaxSmem {
    using C = FftCfg<N>;
    static constexpr size_t SLOT = (C::SIZE_B > C::N ? C::SIZE_B : C::N);   // complex elements: holds A, then B
    static constexpr size_t RING = size_t(K1MAX_STAGES) * C::SPB * SLOT * sizeof(float2);
    // level-0 sums x2 (iteration parity) + level-1 maxima (their OWN buffer: when every group is pruned at level 1 there is no
    // trailing barrier, and a warp racing into the next iteration writes its level-0 sum while slower warps still read the
    // level-1 values) + the CTA's snapshot of the running maximum
    static constexpr size_t RED = size_t(3) * C::SPB * 32 * sizeof(float) + 64;
    static constexpr size_t BAR = 64;
    static constexpr size_t TW1 = size_t(15 * 16) * sizeof(float2);
    static constexpr size_t TOTAL = RING + RED + BAR + TW1;
};

// ZF: the input holds N/ZF points and is zero-filled at the end (zero_fill's default geometry): only the first R0/ZF rows
// of a stage-0 column are loaded, bounded and -- for survivors -- transformed (degenerate first butterfly layers).
template <int N, int ZF = 1>
__global__ void __launch_bounds__(FftCfg<N>::THREADS, (N >= 8192 ? 1 : 2)) k1_max_kernel(const __grid_constant__ K1Params p) {
    using C = FftCfg<N>;
    using SM = K1MaxSmem<N>;
    static_assert((C::E == 16 || C::E == 32) && C::R1 == 16 && C::R2 == 16 && C::T >= 32, "k1_max_kernel: N in [512, 8192]");
    constexpr bool TWP = (N <= 4096);    // persistent stage-0 twiddles (N = 8192: 31 of them do not fit beside 32 points)
    static_assert(ZF >= 1 && C::R0 >= ZF && (ZF & (ZF - 1)) == 0, "zero-fill factor: a power of two <= R0");
    constexpr int NR = C::R0 / ZF;       // non-zero rows of a stage-0 column
    constexpr int NTW = C::C0 * (C::R0 - 1);
    constexpr int WPG = C::T / 32;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* ring = reinterpret_cast<float2*>(smem_raw);
    float* red = reinterpret_cast<float*>(smem_raw + SM::RING);
    float* red1 = red + 2 * C::SPB * 32;    // level-1 maxima
    float* run_s = red + 3 * C::SPB * 32;   // [2]: thread 0's read of the running maximum, double buffered by iteration parity
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + SM::RING + SM::RED);
    float2* tw1_tab = reinterpret_cast<float2*>(smem_raw + SM::RING + SM::RED + SM::BAR);

    const int tid = threadIdx.x, g = tid / C::T, t = tid % C::T;
    const long long ntiles = (p.batch + C::SPB - 1) / C::SPB;

    float2 tw_persist[TWP ? NTW : 1];
    float2 tw0_base[C::C0 * 2], tw1_base[C::C1 * 2];
    init_twiddles<C, false>(t, p.twN, TWP ? tw_persist : nullptr, tw0_base, tw1_base);
    float wcol[C::C0];
#pragma unroll
    for (int j = 0; j < C::C0; ++j) wcol[j] = p.win ? p.win[t + C::T * j] : p.scale;
    for (int i = tid; i < 15 * 16; i += C::THREADS) tw1_tab[i] = p.twN[((i % 16) * (i / 16 + 1) * C::R0) % C::N];

    auto issue = [&](long long tile, int slot) {
        const long long s0 = tile * C::SPB;
        const int nvalid = int((p.batch - s0) < C::SPB ? (p.batch - s0) : C::SPB);
        constexpr uint32_t row_bytes = uint32_t(C::N / ZF) * 8u;
        mbar_arrive_expect_tx(&bars[slot], row_bytes * nvalid);
        float2* dst = ring + size_t(slot) * C::SPB * SM::SLOT;
        for (int r = 0; r < nvalid; ++r) bulk_g2s(dst + size_t(r) * SM::SLOT, p.in + (s0 + r) * (C::N / ZF), row_bytes, &bars[slot]);
    };
    if (tid == 0) {
        for (int s = 0; s < K1MAX_STAGES; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < K1MAX_STAGES; ++s) {
            const long long tile = blockIdx.x + (long long)s * gridDim.x;
            if (tile < ntiles) issue(tile, s);
        }
    }

    int it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int slot = it % K1MAX_STAGES;
        const long long spec = tile * C::SPB + g;
        const bool valid = spec < p.batch;
        float2* my_slot = ring + (size_t(slot) * C::SPB + g) * SM::SLOT;
        // ONE thread samples the (concurrently growing) running maximum and publishes it through shared memory: the skip
        // decision gates block barriers, so every thread of the CTA must see the same value.
        if (tid == 0) run_s[it & 1] = *reinterpret_cast<volatile float*>(p.run_max2);
        mbar_wait(&bars[slot], (it / K1MAX_STAGES) & 1);

        // ---- level-0 bound (no transform at all): |X_k| = |sum_n x_n w_n e^{..}| <= sum_n |x_n| w_n  (triangle inequality).
        // On decaying multi-line FIDs this L1 norm is within ~1.2x of the spectrum's own maximum -- tighter than the
        // Cauchy-Schwarz bound after a first radix-16 stage (~1.5x) at a tenth of the instructions.
        float2 v[C::E];
        float l1 = 0.f;
#pragma unroll
        for (int j = 0; j < C::C0; ++j) {
            float cj = 0.f;
#pragma unroll
            for (int n1 = 0; n1 < C::R0; ++n1) {
                if (n1 < NR) {
                    const float2 x = valid ? my_slot[C::M * n1 + t + C::T * j] : make_float2(0.f, 0.f);
                    v[j * C::R0 + n1] = x;
                    cj = fmaf(sqrt_approx(fmaf(x.x, x.x, x.y * x.y)), fabsf(p.win_rows[n1]), cj);
                } else {
                    v[j * C::R0 + n1] = make_float2(0.f, 0.f);
                }
            }
            l1 = fmaf(cj, fabsf(wcol[j]), l1);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) l1 += __shfl_xor_sync(0xffffffffu, l1, off);
        // (level-0 sums are double buffered by iteration parity: a warp that races ahead into the next iteration must not
        //  overwrite sums a slower warp is still reading -- the decision below has to be CTA-uniform)
        float* red0 = red + (it & 1) * C::SPB * 32;
        if ((t & 31) == 0) red0[g * 32 + (t >> 5)] = l1;
        __syncthreads();                       // also: every thread has read its part of the landing slot
        const float run2 = run_s[it & 1];      // written before this barrier
        bool skip_mine, skip_all = true;
        {
            float mine = 0.f;
#pragma unroll
            for (int gg = 0; gg < C::SPB; ++gg) {
                float m = red0[gg * 32];
#pragma unroll
                for (int w = 1; w < WPG; ++w) m += red0[gg * 32 + w];
                const bool sk = (m * m * 1.0001f < run2);
                skip_all = skip_all && sk;
                if (gg == g) mine = sk ? 1.f : 0.f;
            }
            skip_mine = mine != 0.f;
        }
        if (!skip_all) {
            // ---- survivors: exchange A, stage 1, level-1 bound |X|^2 <= 16 * sum_b max_c |Z[k1][c][b]|^2 -----------------
            if (!skip_mine) {
#pragma unroll
                for (int j = 0; j < C::C0; ++j)
#pragma unroll
                    for (int n1 = 0; n1 < NR; ++n1) v[j * C::R0 + n1] = cscale(v[j * C::R0 + n1], wcol[j] * p.win_rows[n1]);
                stage0_compute<C, false, TWP, ZF>(t, v, tw_persist, tw0_base);
                stage0_write<C>(t, my_slot, v);
            }
            __syncthreads();
            float mq = 0.f;
            if (!skip_mine) {
                stage1_load<C>(t, my_slot, v);
                stage1_compute<C, false, true>(t, v, tw1_base, tw1_tab);
#pragma unroll
                for (int j = 0; j < C::C1; ++j) {                                                      // this thread's k1 values
                    float mj = 0.f;
#pragma unroll
                    for (int c = 0; c < C::R1; ++c) {
                        const float2 z = v[j * C::R1 + c];
                        mj = fmaxf(mj, z.x * z.x + z.y * z.y);
                    }
#pragma unroll
                    for (int off = 1; off < 16; off <<= 1) mj += __shfl_xor_sync(0xffffffffu, mj, off);   // sum over b
                    mq = fmaxf(mq, mj);
                }
                mq = fmaxf(mq, __shfl_xor_sync(0xffffffffu, mq, 16));                                  // max over k1
            }
            if ((t & 31) == 0) red1[g * 32 + (t >> 5)] = mq;
            __syncthreads();                   // every survivor holds its exchange-A inputs: the slot may be reused
            bool skip1_all = true, skip1_mine = true;
#pragma unroll
            for (int gg = 0; gg < C::SPB; ++gg) {
                float m = red1[gg * 32];
#pragma unroll
                for (int w = 1; w < WPG; ++w) m = fmaxf(m, red1[gg * 32 + w]);
                const bool sk = (16.0f * m * 1.0001f < run2);     // groups pruned at level 0 wrote 0 -> skipped here too
                skip1_all = skip1_all && sk;
                if (gg == g) skip1_mine = sk;
            }
            if (!skip1_all) {
                if (!skip1_mine) stage1_write<C>(t, my_slot, v);
                __syncthreads();
                if (!skip1_mine) {
                    stage2<C, false>(t, my_slot, v);
                    float best = 0.f;
#pragma unroll
                    for (int i = 0; i < C::E; ++i) best = fmaxf(best, v[i].x * v[i].x + v[i].y * v[i].y);
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, off));
                    if ((t & 31) == 0 && valid) {
                        atomicMax(reinterpret_cast<int*>(p.absmax + spec), __float_as_int(sqrtf(best)));
                        atomicMax(reinterpret_cast<int*>(p.run_max2), __float_as_int(best));
                    }
                }
                __syncthreads();               // exchange B fully consumed before the slot is re-armed
            }
        }
        if (tid == 0) {
            const long long nt = tile + (long long)K1MAX_STAGES * gridDim.x;
            if (nt < ntiles) {
                fence_proxy_async_smem();
                issue(nt, slot);
            }
        }
    }
}

// ---- zero-filled input, ZF tiles per iteration -------------------------------------------------------------------------
// With input rows of N/ZF points the per-tile work of a pruned spectrum is so small that the loop above is bound by its
// latency chain (mbarrier wait -> loads -> reduction -> block barrier -> re-arm) per 8*N/ZF bytes.  This variant bounds ZF
// consecutive tiles per iteration -- the same bytes and the same registers per iteration as the full-length kernel -- with
// ONE block barrier, on a ring of 2 groups x ZF short landing slots; survivors run the two-level pipeline one tile at a time
// through a separate exchange buffer, so the landing slots are re-armed right after the bound, survivors or not.
template <int N, int ZF>
struct K1MaxZfSmem {
    using C = FftCfg<N>;
    static constexpr size_t ROW = size_t(C::N / ZF);                                   // complex elements per input row
    static constexpr size_t RING = size_t(2) * ZF * C::SPB * ROW * sizeof(float2);     // = 2 * SPB * N * 8 bytes
    static constexpr size_t XSLOT = (C::SIZE_B > C::N ? C::SIZE_B : C::N);             // exchange A, then B
    static constexpr size_t X = size_t(C::SPB) * XSLOT * sizeof(float2);
    static constexpr size_t RED = size_t(2) * ZF * C::SPB * 32 * sizeof(float) + 64;   // level-0 sums x2 (iteration parity)
    static constexpr size_t BAR = 64;
    static constexpr size_t TW1 = size_t(15 * 16) * sizeof(float2);
    static constexpr size_t TOTAL = RING + X + RED + BAR + TW1;
};

template <int N, int ZF>
__global__ void __launch_bounds__(FftCfg<N>::THREADS, (N >= 8192 ? 1 : 2)) k1_max_zf_kernel(const __grid_constant__ K1Params p) {
    using C = FftCfg<N>;
    using SM = K1MaxZfSmem<N, ZF>;
    static_assert((C::E == 16 || C::E == 32) && C::R1 == 16 && C::R2 == 16 && C::T >= 32, "k1_max_zf_kernel: N in [512, 8192]");
    constexpr bool TWP = (N <= 4096);
    static_assert(ZF >= 2 && C::R0 >= ZF && (ZF & (ZF - 1)) == 0, "zero-fill factor: a power of two in [2, R0]");
    constexpr int NR = C::R0 / ZF;       // non-zero rows of a stage-0 column
    constexpr int NV = C::C0 * NR;       // samples per thread and tile
    constexpr int NTW = C::C0 * (C::R0 - 1);
    constexpr int WPG = C::T / 32;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* ring = reinterpret_cast<float2*>(smem_raw);
    float2* xbuf = reinterpret_cast<float2*>(smem_raw + SM::RING);
    float* red = reinterpret_cast<float*>(smem_raw + SM::RING + SM::X);
    float* run_s = red + 2 * ZF * C::SPB * 32;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + SM::RING + SM::X + SM::RED);
    float2* tw1_tab = reinterpret_cast<float2*>(smem_raw + SM::RING + SM::X + SM::RED + SM::BAR);

    const int tid = threadIdx.x, g = tid / C::T, t = tid % C::T;
    constexpr long long SPT = (long long)ZF * C::SPB;                 // spectra per super-tile
    const long long ntiles = (p.batch + SPT - 1) / SPT;

    float2 tw_persist[TWP ? NTW : 1];
    float2 tw0_base[C::C0 * 2], tw1_base[C::C1 * 2];
    init_twiddles<C, false>(t, p.twN, TWP ? tw_persist : nullptr, tw0_base, tw1_base);
    float wcol[C::C0];
#pragma unroll
    for (int j = 0; j < C::C0; ++j) wcol[j] = p.win ? p.win[t + C::T * j] : p.scale;
    for (int i = tid; i < 15 * 16; i += C::THREADS) tw1_tab[i] = p.twN[((i % 16) * (i / 16 + 1) * C::R0) % C::N];

    auto issue = [&](long long tile, int grp) {
        const long long s0 = tile * SPT;
        const long long left = p.batch - s0;
        const int nvalid = int(left < SPT ? left : SPT);
        constexpr uint32_t row_bytes = uint32_t(SM::ROW) * 8u;
        mbar_arrive_expect_tx(&bars[grp], row_bytes * nvalid);
        // rows of a super-tile are consecutive in HBM and in the ring group: one bulk copy
        bulk_g2s(ring + size_t(grp) * SPT * SM::ROW, p.in + s0 * (long long)SM::ROW, row_bytes * nvalid, &bars[grp]);
    };
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            const long long tile = blockIdx.x + (long long)s * gridDim.x;
            if (tile < ntiles) issue(tile, s);
        }
    }

    int it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int grp = it & 1;
        const long long spec0 = tile * SPT + g;                       // this thread's spectrum in sub-tile b: spec0 + b*SPB
        if (tid == 0) run_s[it & 1] = *reinterpret_cast<volatile float*>(p.run_max2);
        mbar_wait(&bars[grp], (it >> 1) & 1);

        // ---- level-0 bound of ZF sub-tiles: |X| <= sum_n |x_n| |w_n| ---------------------------------------------------
        float* red0 = red + (it & 1) * ZF * C::SPB * 32;               // double buffered by iteration parity (see above)
        float2 raw[ZF][NV];
#pragma unroll
        for (int b = 0; b < ZF; ++b) {
            const bool valid = spec0 + (long long)b * C::SPB < p.batch;
            const float2* row = ring + (size_t(grp) * SPT + size_t(b) * C::SPB + g) * SM::ROW;
            float l1 = 0.f;
#pragma unroll
            for (int j = 0; j < C::C0; ++j) {
                float cj = 0.f;
#pragma unroll
                for (int n1 = 0; n1 < NR; ++n1) {
                    const float2 x = valid ? row[C::M * n1 + t + C::T * j] : make_float2(0.f, 0.f);
                    raw[b][j * NR + n1] = x;
                    cj = fmaf(sqrt_approx(fmaf(x.x, x.x, x.y * x.y)), fabsf(p.win_rows[n1]), cj);
                }
                l1 = fmaf(cj, fabsf(wcol[j]), l1);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) l1 += __shfl_xor_sync(0xffffffffu, l1, off);
            if ((t & 31) == 0) red0[(b * C::SPB + g) * 32 + (t >> 5)] = l1;
        }
        __syncthreads();                       // also: every thread holds its samples, the landing group is free
        if (tid == 0) {
            const long long nt = tile + 2LL * gridDim.x;
            if (nt < ntiles) {
                fence_proxy_async_smem();
                issue(nt, grp);
            }
        }
        const float run2 = run_s[it & 1];
        unsigned surv = 0;                     // bit (b*SPB + gg): that spectrum survives level 0 (CTA-uniform)
#pragma unroll
        for (int i = 0; i < ZF * C::SPB; ++i) {
            float m = red0[i * 32];
#pragma unroll
            for (int w = 1; w < WPG; ++w) m += red0[i * 32 + w];
            if (!(m * m * 1.0001f < run2)) surv |= 1u << i;
        }
        if (surv == 0) continue;

        // ---- survivors, one sub-tile at a time through the exchange buffer --------------------------------------------
        float2* my_x = xbuf + size_t(g) * SM::XSLOT;
#pragma unroll
        for (int b = 0; b < ZF; ++b) {
            const unsigned sub = (surv >> (b * C::SPB)) & ((1u << C::SPB) - 1u);
            if (sub == 0) continue;            // CTA-uniform
            const bool mine = (sub >> g) & 1u;
            const long long spec = spec0 + (long long)b * C::SPB;
            float2 v[C::E];
            if (mine) {
#pragma unroll
                for (int j = 0; j < C::C0; ++j)
#pragma unroll
                    for (int n1 = 0; n1 < C::R0; ++n1)
                        v[j * C::R0 + n1] = n1 < NR ? cscale(raw[b][j * NR + n1], wcol[j] * p.win_rows[n1]) : make_float2(0.f, 0.f);
                stage0_compute<C, false, TWP, ZF>(t, v, tw_persist, tw0_base);
                stage0_write<C>(t, my_x, v);
            }
            __syncthreads();
            float mq = 0.f;
            if (mine) {
                stage1_load<C>(t, my_x, v);
                stage1_compute<C, false, true>(t, v, tw1_base, tw1_tab);
#pragma unroll
                for (int j = 0; j < C::C1; ++j) {                                                      // this thread's k1 values
                    float mj = 0.f;
#pragma unroll
                    for (int c = 0; c < C::R1; ++c) {
                        const float2 z = v[j * C::R1 + c];
                        mj = fmaxf(mj, z.x * z.x + z.y * z.y);
                    }
#pragma unroll
                    for (int off = 1; off < 16; off <<= 1) mj += __shfl_xor_sync(0xffffffffu, mj, off);   // sum over b
                    mq = fmaxf(mq, mj);
                }
                mq = fmaxf(mq, __shfl_xor_sync(0xffffffffu, mq, 16));                                  // max over k1
            }
            if ((t & 31) == 0) red[g * 32 + (t >> 5)] = mq;
            __syncthreads();                   // every survivor holds its exchange-A inputs: the buffer may be reused
            bool skip1_all = true, skip1_mine = true;
#pragma unroll
            for (int gg = 0; gg < C::SPB; ++gg) {
                float m = red[gg * 32];
#pragma unroll
                for (int w = 1; w < WPG; ++w) m = fmaxf(m, red[gg * 32 + w]);
                const bool sk = (16.0f * m * 1.0001f < run2);     // groups pruned at level 0 wrote 0 -> skipped here too
                skip1_all = skip1_all && sk;
                if (gg == g) skip1_mine = sk;
            }
            if (!skip1_all) {
                if (!skip1_mine) stage1_write<C>(t, my_x, v);
                __syncthreads();
                if (!skip1_mine) {
                    stage2<C, false>(t, my_x, v);
                    float best = 0.f;
#pragma unroll
                    for (int i = 0; i < C::E; ++i) best = fmaxf(best, v[i].x * v[i].x + v[i].y * v[i].y);
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, off));
                    if ((t & 31) == 0 && spec < p.batch) {
                        atomicMax(reinterpret_cast<int*>(p.absmax + spec), __float_as_int(sqrtf(best)));
                        atomicMax(reinterpret_cast<int*>(p.run_max2), __float_as_int(best));
                    }
                }
            }
            __syncthreads();                   // exchange buffer and `red` free for the next sub-tile / iteration
        }
    }
}

}  // namespace xmr
