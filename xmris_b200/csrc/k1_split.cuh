// K1-split: the fused zero-fill -> window -> FFT -> fftshift [-> uniform phase] -> store kernel for N = 8192 as two interleaved
// 4096-point transforms (fft_split.cuh) -- 512 threads, 16 points per thread, the register budget of the 4096-point kernel.
//
// One persistent CTA per SM loops over spectra:
//   thread 0            cp.async.bulk of the raw FID row (n_in = N / ZF points) into ring slot it % 2 (mbarrier completion)
//   group g = tid / 256 stage 0 on samples 2*i + g of the slot, IN PLACE  -> bar -> stage 1 into exchange B_g -> bar
//                       -> [slot re-armed for spectrum it + 2] -> stage 2 -> E (g = 0) / W_N^k O (g = 1) in registers
//                       -> bar -> exchange through the B region -> bar -> X[k] = E + W O (g = 0), X[k + N/2] = E - W O (g = 1)
//                       -> phase -> streaming stores (fftshift folded into the index)
// Geometry: input at the start of the row (pad_left = 0), separable window, out_shift = N/2 -- zero_fill's default ("end")
// followed by apodize_exp, i.e. config C3.  Everything else at N = 8192 stays on the generic three-stage kernel.
#pragma once
#include <cuda_runtime.h>

#include "fft_split.cuh"
#include "k1_fft.cuh"

namespace xmr {

template <int NH>
struct K1SplitSmem {
    using C = FftCfg<NH>;
    static constexpr int N = 2 * NH;
    static constexpr int STAGES = 2;
    static constexpr size_t RING = size_t(STAGES) * N * sizeof(float2);
    static constexpr size_t B = size_t(2) * C::SIZE_B * sizeof(float2);          // exchange B of both groups; later E | W O
    static constexpr size_t TW1 = size_t(15 * 16) * sizeof(float2);
    static constexpr size_t PH = size_t(16) * sizeof(float2);
    static constexpr size_t BAR = 64;
    static constexpr size_t TOTAL = RING + B + TW1 + PH + BAR;
    static_assert(B >= size_t(N) * sizeof(float2), "the E / W O exchange reuses the B region");
};

template <int NH, int ZF, bool PHASE, bool PHDEV>
__global__ void __launch_bounds__(2 * FftCfg<NH>::T, 1) k1_split_kernel(const __grid_constant__ K1Params p) {
    using C = FftCfg<NH>;
    using SM = K1SplitSmem<NH>;
    constexpr int N = 2 * NH;
    constexpr int Q = C::R0 * C::R1;
    static_assert(C::C0 == 1 && C::C1 == 1 && C::C2 == 1 && C::E == 16, "written for 16 points per thread");
    static_assert(C::R0 >= ZF, "zero-fill factor <= first radix");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* ring = reinterpret_cast<float2*>(smem_raw);
    float2* Bbuf = reinterpret_cast<float2*>(smem_raw + SM::RING);
    float2* tw1_tab = reinterpret_cast<float2*>(smem_raw + SM::RING + SM::B);
    float2* ph_tab = reinterpret_cast<float2*>(smem_raw + SM::RING + SM::B + SM::TW1);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + SM::RING + SM::B + SM::TW1 + SM::PH);

    const int tid = threadIdx.x;
    const int g = tid / C::T, t = tid % C::T;
    constexpr int n_in = N / ZF;

    // ---- per-thread persistent state ---------------------------------------------------------------------------------------
    float2 tw_persist[C::R0 - 1];
    float2 tw0_base[2], tw1_base[2];
    init_twiddles<C, false>(t, p.twH, tw_persist, tw0_base, tw1_base);
    const float2 wq = p.twN[t];                                // W_N^q, q = t (combine twiddle of the odd group)
    int n2p, carry;
    split_window_index<C>(t, 0, g, &n2p, &carry);
    const float wcol = p.win ? p.win[n2p] : p.scale;
    double ph_a = p.ph_a_turns, ph_b = p.ph_b_turns;
    if (PHASE && PHDEV) {
        ph_a = p.ph_dev->ph_a_turns;
        ph_b = p.ph_dev->ph_b_turns;
    }
    float2 ph_base = make_float2(1.f, 0.f);
    if (PHASE) {
        // stored index m = t + Q*d + (g == 0 ? NH : 0):  exp(2 pi i (a + b m)) = ph_base * ph_tab[d]
        double turns = ph_a + ph_b * double(t + (g == 0 ? NH : 0));
        turns -= floor(turns);
        double s, c;
        sincospi(2.0 * turns, &s, &c);
        ph_base = make_float2(float(c), float(s));
        if (tid < 16) {
            double td = ph_b * double(Q) * double(tid);
            td -= floor(td);
            sincospi(2.0 * td, &s, &c);
            ph_tab[tid] = make_float2(float(c), float(s));
        }
    }
    for (int i = tid; i < 15 * 16; i += 2 * C::T) {
        const int c = i / 16 + 1, b = i % 16;
        tw1_tab[i] = p.twH[(b * c * C::R0) % NH];             // W_M^(b c) of the NH-point transform
    }

    auto issue = [&](long long spec, int slot) {
        constexpr uint32_t row_bytes = uint32_t(n_in) * 8u;
        mbar_arrive_expect_tx(&bars[slot], row_bytes);
        bulk_g2s(ring + size_t(slot) * N, p.in + spec * n_in, row_bytes, &bars[slot]);
    };
    if (tid == 0) {
        for (int s = 0; s < SM::STAGES; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < SM::STAGES; ++s) {
            const long long spec = blockIdx.x + (long long)s * gridDim.x;
            if (spec < p.batch) issue(spec, s);
        }
    }

    float2* my_B = Bbuf + size_t(g) * C::SIZE_B;
    float2* my_X = Bbuf + size_t(g) * NH;                      // E (g = 0) | W O (g = 1) after stage 2
    const float2* other_X = Bbuf + size_t(1 - g) * NH;

    int it = 0;
    for (long long spec = blockIdx.x; spec < p.batch; spec += gridDim.x, ++it) {
        const int slot = it % SM::STAGES;
        float2* my_slot = ring + size_t(slot) * N;
        mbar_wait(&bars[slot], (it / SM::STAGES) & 1);

        // ---- stage 0 on this group's samples (in place: every thread rewrites only its own column) ---------------------
        float2 v[C::E];
        {
            constexpr int NR = C::R0 / ZF;
#pragma unroll
            for (int n1 = 0; n1 < C::R0; ++n1) {
                if (n1 < NR) {
                    const float wr = carry ? p.win_rows[2 * n1 + 1] : p.win_rows[2 * n1];
                    v[n1] = cscale(my_slot[2 * (C::M * n1 + t) + g], wcol * wr);
                } else {
                    v[n1] = make_float2(0.f, 0.f);
                }
            }
        }
        stage0_compute<C, false, true, ZF>(t, v, tw_persist, tw0_base);
        split_stage0_write<C>(t, g, my_slot, v);
        __syncthreads();
        // ---- stage 1 ------------------------------------------------------------------------------------------------------
        split_stage1_load<C>(t, g, my_slot, v);
        stage1_store<C, false, true>(t, my_B, v, tw1_base, tw1_tab);
        __syncthreads();
        if (tid == 0) {                                         // the landing slot is free: prefetch spectrum it + 2
            const long long ns = spec + (long long)SM::STAGES * gridDim.x;
            if (ns < p.batch) {
                fence_proxy_async_smem();
                issue(ns, slot);
            }
        }
        // ---- stage 2, combine ------------------------------------------------------------------------------------------
        stage2<C, false>(t, my_B, v);
        if (g == 1) split_twiddle_odd<C>(v, &wq);
        __syncthreads();                                        // exchange B fully read: its memory now carries E | W O
#pragma unroll
        for (int d = 0; d < C::R2; ++d) my_X[t + Q * d] = v[d];
        __syncthreads();
        float2* dst = p.out + spec * (long long)N + (g == 0 ? NH : 0);
#pragma unroll
        for (int d = 0; d < C::R2; ++d) {
            const float2 o = other_X[t + Q * d];
            float2 x = g == 0 ? cadd(v[d], o) : csub(o, v[d]);  // X[k] = E + W O ;  X[k + NH] = E - W O
            if (PHASE) x = cmul(x, cmul(ph_base, ph_tab[d]));
            st_stream(dst + t + Q * d, x);
        }
        // (the next iteration's first write into the B region comes after its stage-0 barrier: every thread has left this
        //  loop's reads by then)
    }
}

}  // namespace xmr
