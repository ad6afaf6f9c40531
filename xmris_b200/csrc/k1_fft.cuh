// K1: fused  zero-fill -> window -> FFT -> fftshift [-> |S| statistics] [-> uniform phase] -> store.
//
// One persistent CTA loops over "tiles" of SPB consecutive spectra.  Per tile:
//   producer (thread 0)   cp.async.bulk global->shared of the raw FID rows into ring slot (it % STAGES),
//                         completion on an mbarrier (TMA bulk-copy engine; SASS UBLKCP)
//   all threads           wait(mbarrier) -> stage0 (in-place exchange A) -> bar -> stage1 (exchange B) -> bar
//                         -> [producer re-arms the slot for tile it+STAGES] -> stage2 -> epilogue from registers
// Padded points are never read or written in HBM; the shifted store index implements fftshift.
// HBM traffic per spectrum = 8*n_in + 8*n_out bytes (+8 B of statistics), the algorithmic minimum.
#pragma once
#include <cuda_runtime.h>

#include "fft_stages.cuh"
#include "ptx_sm100.cuh"

namespace xmr {

constexpr int K1_STAGES = 2;


// Phase parameters of the fused epilogue in DEVICE memory (written by search_finalize_kernel): lets pass 2 of the
// mode="single" chain be enqueued before the search has finished -- no host read-back between the passes.
struct K1PhaseDev {
    double ph_a_turns, ph_b_turns;
    float2 ph_step[16];
    float2 ph_fold[16];
};

struct K1Params {
    const float2* in;
    float2* out;
    long long batch;
    int n_in;
    int pad_left;
    int in_shift;
    int out_shift;
    float scale;
    const float2* twN;     // exp(-2 pi i k / N), k < N
    const float* win;      // table [N] (WIN==1) or column factors [min(N,256)] (WIN==2)
    float win_rows[32];    // row factors (WIN==2)
    float* absmax;
    int* argmax;
    int phase_on;          // epilogue: multiply by exp(2 pi i (a + b*m)), m = stored index
    double ph_a_turns;     // uniform phase: turns(m) = a + b*m
    double ph_b_turns;
    float2 ph_step[16];    // exp(2 pi i * b * R0*R1 * d), d < 16
    float2 ph_fold[16];    // folded-phase variants: exp(2 pi i (a + b * ((R0*R1*d + N/2) mod N))), d < 16
    float* run_max2;       // k1_max_kernel: running global max of |S|^2 (device scalar, zeroed by the launcher)
    const long long* row_slot;   // generic kernel, optional: transform the row starting at in + (*row_slot) * row_stride
    int row_stride;              //   (the device-resident winner among the gathered candidate rows; elements)
    const K1PhaseDev* ph_dev;    // optional: phase parameters from device memory instead of ph_* above
};

// N >= 8192: one shared buffer per spectrum serves as TMA landing slot, exchange A and exchange B ("in-place B"), one
// ring stage, so that two CTAs fit on an SM (2 x 68 KiB) instead of one (196 KiB with separate buffers).
// GROUPS == 2 (N >= 8192, TMA): ONE CTA per SM made of two independent 256-thread groups (named barriers) that share a
// ring of GROUPS + 1 such buffers: the group that finishes tile q re-arms its buffer with tile q + 3, which the OTHER group
// picks up after its own current tile -- every load has half a tile period to land, and both groups compute all the time
// (two separate CTAs with one buffer each leave 8 warps computing whenever one of them waits for its load).
template <int N, int GROUPS = 1>
struct K1Smem {
    using C = FftCfg<N>;
    static constexpr bool INPLACE_B = (N >= 8192);
    static constexpr int STAGES = GROUPS > 1 ? GROUPS + 1 : (INPLACE_B ? 1 : K1_STAGES);
    static constexpr size_t SLOT = INPLACE_B ? (C::SIZE_B > C::N ? C::SIZE_B : C::N) : C::N;   // complex elements
    static constexpr size_t RING = size_t(STAGES) * C::SPB * SLOT * sizeof(float2);
    static constexpr size_t B = INPLACE_B ? 0 : size_t(C::SPB) * C::SIZE_B * sizeof(float2);
    static constexpr size_t RED = size_t(C::SPB) * GROUPS * 32 * 8;  // per spectrum slot: up to 32 warps x (float, int)
    static constexpr size_t BAR = 64;
    static constexpr size_t TW1 = (C::R1 == 16 && C::R2 == 16) ? size_t(15 * 16) * sizeof(float2) : 0;   // stage-1 twiddle table
    static constexpr size_t TOTAL = RING + B + RED + BAR + TW1;
};

template <int N, int GROUPS = 1>
constexpr int k1_min_blocks() {
    // shared-memory limited residency on a 227 KB SM, capped at 2048 threads and at 3 CTAs
    int by_smem = int((227u * 1024u) / (K1Smem<N, GROUPS>::TOTAL + 1024));
    int by_thr = 2048 / (FftCfg<N>::THREADS * GROUPS);
    int m = by_smem < by_thr ? by_smem : by_thr;
    if (m > 2) m = 2;   // persistent per-thread twiddles / 32 points per thread want <= 128 registers at 256 threads
    return m < 1 ? 1 : m;
}

// (value, index) max-reduction with first-occurrence tie-break
__device__ __forceinline__ void amax_combine(float& v, int& i, float ov, int oi) {
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

// WIN: 1 = full window table p.win[N]; 2 = separable p.win[column] * p.win_rows[row] (also "scale only").
// FAST == 0: generic kernel; geometry (n_in, pad_left, shifts) and the epilogue options (statistics, phase, store) are
//            uniform run-time values.
// FAST != 0: specialised hot path for n_in == N (or, with K1_FAST_ZF2 / ZF4, n_in == N/2 / N/4 zero-filled at the end: the
//            zero rows of every stage-0 column are never loaded and the first butterfly layers degenerate), no left
//            padding, no input rotation, out_shift == N/2, with the epilogue fixed at compile time (bit mask of
//            K1_FAST_*): no per-element predicates or index arithmetic.
enum : int { K1_FAST_ON = 1, K1_FAST_STORE = 2, K1_FAST_STATS = 4, K1_FAST_PHASE = 8, K1_FAST_ZF2 = 16, K1_FAST_ZF4 = 32,
             K1_FAST_PHDEV = 64 /* store+phase variants: phase parameters read from p.ph_dev (shared-memory table) */,
             K1_FAST_BULKST = 128 /* in-place-B store variants: results staged in the (free) shared buffer and written by ONE
                                     cp.async.bulk per spectrum instead of 32 STG per thread (p.out 16-byte aligned) */ };

// PRUNE (generic statistics-only launches with p.run_max2 set): branch and bound on the level-0 bound of k1_max.cuh,
//            |X| <= sum_n |x_n w_n|, for ANY geometry (zero-filled input, N = 8192, table windows): a tile whose spectra
//            all fall below the running global maximum skips its transform; survivors are transformed in full.
template <int N, bool INVERSE, int WIN, bool TMA, int FAST = 0, bool PRUNE = false, int GROUPS = 1>
__global__ void __launch_bounds__(FftCfg<N>::THREADS * GROUPS, k1_min_blocks<N, GROUPS>()) k1_kernel(const __grid_constant__ K1Params p) {
    static_assert(!PRUNE || (FAST == 0 && !INVERSE), "PRUNE is a variant of the generic forward statistics pass");
    using KS = K1Smem<N, GROUPS>;
    constexpr bool GRP = GROUPS > 1;
    static_assert(!GRP || (KS::INPLACE_B && TMA), "grouped CTAs: in-place exchange B, TMA loads");
    __shared__ float run_s[2 * GROUPS];   // PRUNE: a group's thread 0 samples the running maximum, double buffered by iteration parity
    __shared__ float2 ph_tab[32];   // generic / PHDEV variants: ph_step[16] | ph_fold[16] (from p.ph_dev when given)
    using C = FftCfg<N>;
    constexpr bool F = (FAST != 0);
    constexpr int ZF = (FAST & K1_FAST_ZF4) ? 4 : ((FAST & K1_FAST_ZF2) ? 2 : 1);   // zero-fill factor of the fast variants
    static_assert(FftCfg<N>::R0 >= ZF, "the zero-fill fast variants need a first radix >= the zero-fill factor");
    constexpr bool TW_PERSIST = (N <= 4096);
    constexpr int NTW = (C::R0 > 1) ? C::C0 * (C::R0 - 1) : 1;
    // Stage-1 twiddles W_M^(b*c) from a shared table (15 loads per 16-point butterfly) or from a power chain (14 multiplies).
    // Measured, alternating runs on one box (the mix also decides the SM clock under the power cap): the table wins for
    // full-length input up to 4096 points (C5 pass 2: 11.4-11.5 ms against 11.5-11.7 ms); the chain wins at N = 8192, where the
    // shared-memory instruction queue is the first stall reason (profiles/k1_8192_full_r2.md: 8192 -> 8192 0.84 -> 0.88 of the
    // roofline), and in the zero-filled store+phase variants of every length (chain 2048 -> 4096: 3.32 -> 3.21 ms,
    // 1024 -> 2048: 3.17 -> 3.08 ms, 512 -> 1024: 3.20 -> 3.12 ms).  The folded phase works with either (see FOLD).
    constexpr bool TW1_TAB_ = (KS::TW1 != 0) && N < 8192 && !(F && (FAST & K1_FAST_PHASE) != 0 && ZF > 1);
    // Folded phase (fast store+phase variants): with the stored index m = k1 + R0*c + m0(d), m0(d) = (R0*R1*d + N/2) mod N,
    // the rotation exp(2 pi i (a + b*m)) factors into E1(k1) * E2(c) * step(d).  E1 rides on the persistent stage-0
    // twiddles (N = 8192: on the two bases W^n2, W^(4 n2) of their power chain w[k] = w1^(k&3) * w4^(k>>2), which then carries
    // E1(k) = E1(1)^k), E2 on the shared stage-1 twiddle table (both are 1 at index 0, where no multiply exists), so the
    // epilogue is ONE complex multiply per point by a kernel-parameter constant instead of two.
    // Without the table E2 rides on the two bases of the stage-1 power chain in the same way.
    constexpr bool FOLD = F && ((FAST & K1_FAST_PHASE) != 0) && C::R0 > 1 && C::R1 > 1;
    constexpr bool IPB = KS::INPLACE_B;
    constexpr bool BULKST = F && (FAST & K1_FAST_BULKST) != 0;
    static_assert(!BULKST || (IPB && TMA && (FAST & K1_FAST_STORE) != 0 && C::SPB == 1), "bulk-store variants: one spectrum per tile, in-place B");
    constexpr int STAGES = KS::STAGES;
    constexpr size_t SLOT = KS::SLOT;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* ring = reinterpret_cast<float2*>(smem_raw);
    float2* Bbuf = reinterpret_cast<float2*>(smem_raw + KS::RING);
    float* red = reinterpret_cast<float*>(smem_raw + KS::RING + KS::B);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + KS::RING + KS::B + KS::RED);
    constexpr bool TW1_TAB = TW1_TAB_;
    float2* tw1_tab = TW1_TAB ? reinterpret_cast<float2*>(smem_raw + KS::RING + KS::B + KS::RED + KS::BAR) : nullptr;

    const int grp = GRP ? int(threadIdx.x) / C::THREADS : 0;      // grouped CTAs: which of the independent groups
    const int tid = GRP ? int(threadIdx.x) % C::THREADS : int(threadIdx.x);
    // block barrier of one group (named barrier 1 + grp) -- the whole CTA when there is one group
    auto bsync = [&]() {
        if (GRP) asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(C::THREADS) : "memory");
        else __syncthreads();
    };
    auto bsync_and = [&](bool pred) -> bool {
        if (GRP) {
            int r;
            asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %1, 0;\n\tbar.red.and.pred q, %2, %3, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                         : "=r"(r) : "r"(int(pred)), "r"(1 + grp), "n"(C::THREADS) : "memory");
            return r != 0;
        }
        return __syncthreads_and(pred) != 0;
    };
    const int g = tid / C::T;     // spectrum slot within the tile
    const int t = tid % C::T;     // thread within the spectrum
    constexpr bool PHDEV = F && ((FAST & K1_FAST_PHDEV) != 0);
    constexpr bool PH_TABLE = !F || PHDEV;          // the epilogue reads its step / fold phasors from shared memory
    const float2* in_base = p.in;
    if (!F && p.row_slot != nullptr) in_base += (*p.row_slot) * (long long)p.row_stride;
    double ph_a_turns = p.ph_a_turns, ph_b_turns = p.ph_b_turns;
    if (PH_TABLE) {
        if (p.ph_dev != nullptr) {
            ph_a_turns = p.ph_dev->ph_a_turns;
            ph_b_turns = p.ph_dev->ph_b_turns;
        }
        if (tid < 32) {
            ph_tab[tid] = p.ph_dev != nullptr ? (tid < 16 ? p.ph_dev->ph_step[tid] : p.ph_dev->ph_fold[tid - 16])
                                              : (tid < 16 ? p.ph_step[tid] : p.ph_fold[tid - 16]);
        }
        // (published by the block barriers that precede the first epilogue)
    }

    const long long ntiles = (p.batch + C::SPB - 1) / C::SPB;
    const int n_in = F ? C::N / ZF : p.n_in;
    const int pad_left = F ? 0 : p.pad_left;
    const int in_shift = F ? 0 : p.in_shift;
    const int out_shift = F ? C::N / 2 : p.out_shift;
    const bool do_stats = F ? ((FAST & K1_FAST_STATS) != 0) : (p.absmax != nullptr);
    const bool do_index = F ? false : (p.argmax != nullptr);      // the fast statistics pass records maxima only
    const bool do_store = F ? ((FAST & K1_FAST_STORE) != 0) : (p.out != nullptr);
    const bool do_phase = F ? ((FAST & K1_FAST_PHASE) != 0) : (p.phase_on != 0);
    const bool need_load_barrier = (pad_left != 0) || ((in_shift % C::M) != 0);

    // ---- per-thread persistent state -------------------------------------------------------------------
    float2 tw_persist[TW_PERSIST ? NTW : 1];
    float2 tw0_base[C::C0 * 2], tw1_base[C::C1 * 2];
    init_twiddles<C, INVERSE>(t, p.twN, TW_PERSIST ? tw_persist : nullptr, tw0_base, tw1_base);
    float wcol[C::C0];
    if (WIN == 2) {
#pragma unroll
        for (int j = 0; j < C::C0; ++j) wcol[j] = p.win ? p.win[t + C::T * j] : p.scale;   // null: scale only
    }
    if (FOLD) {
#pragma unroll
        for (int k1 = 1; k1 < C::R0; ++k1) {
            double turns = ph_b_turns * double(k1);
            turns -= floor(turns);
            double s, c;
            sincospi(2.0 * turns, &s, &c);
            if (!TW_PERSIST && k1 != 1 && k1 != 4) continue;     // power chain: only its two bases carry the phase
#pragma unroll
            for (int j = 0; j < C::C0; ++j) {
                float2& w = TW_PERSIST ? tw_persist[(TW_PERSIST ? j * (C::R0 - 1) + k1 - 1 : 0)] : tw0_base[2 * j + (k1 == 4 ? 1 : 0)];
                const double wr = double(w.x) * c - double(w.y) * s, wi = double(w.x) * s + double(w.y) * c;
                w = make_float2(float(wr), float(wi));
            }
        }
    }
    if (FOLD && !TW1_TAB_) {
        // stage-1 power chain w[c] = w1^(c&3) * w4^(c>>2): E2(c) = E2(1)^c rides on its bases (E2(4) computed exactly)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            double turns = ph_b_turns * double(C::R0 * (e == 0 ? 1 : 4));
            turns -= floor(turns);
            double s, c;
            sincospi(2.0 * turns, &s, &c);
#pragma unroll
            for (int j = 0; j < C::C1; ++j) {
                float2& w = tw1_base[2 * j + e];
                const double wr = double(w.x) * c - double(w.y) * s, wi = double(w.x) * s + double(w.y) * c;
                w = make_float2(float(wr), float(wi));
            }
        }
    }
    float2 ph_base[C::C2];
#pragma unroll
    for (int j = 0; j < C::C2; ++j) ph_base[j] = make_float2(1.f, 0.f);
    if (do_phase && !FOLD) {
#pragma unroll
        for (int j = 0; j < C::C2; ++j) {
            const int q = t + C::T * j;
            double turns = ph_a_turns + ph_b_turns * double(q);
            turns -= floor(turns);
            double s, c;
            sincospi(2.0 * turns, &s, &c);
            ph_base[j] = make_float2(float(c), float(s));
        }
    }

    // Barriers: one per ring slot -- and per consuming group in grouped CTAs (tile `it` uses barrier it % NBAR, slot it % STAGES;
    // GROUPS and STAGES are coprime): every barrier is then waited on by ONE group in phase order, so the 1-bit phase parity
    // can never alias (with one barrier per slot the other group's tile sits between two of mine: a group running two
    // phases ahead of a late load would see "its" parity already flipped).
    constexpr int NBAR = GRP ? GROUPS * STAGES : STAGES;
    static_assert(!GRP || (STAGES % GROUPS != 0 && NBAR * 8 <= int(KS::BAR)), "grouped CTAs: coprime ring / group counts");
    auto issue = [&](long long tile, int slot, int bar) {
        // producer: arm the barrier with the tile's byte count, then one bulk copy per spectrum row
        const long long s0 = tile * C::SPB;
        const int nvalid = int((p.batch - s0) < C::SPB ? (p.batch - s0) : C::SPB);
        const uint32_t row_bytes = uint32_t(n_in) * 8u;
        mbar_arrive_expect_tx(&bars[bar], row_bytes * nvalid);
        float2* dst = ring + size_t(slot) * C::SPB * SLOT;
        if (n_in == C::N && SLOT == size_t(C::N)) {
            bulk_g2s(dst, in_base + s0 * n_in, row_bytes * nvalid, &bars[bar]);
        } else {
            for (int r = 0; r < nvalid; ++r) bulk_g2s(dst + size_t(r) * SLOT, in_base + (s0 + r) * n_in, row_bytes, &bars[bar]);
        }
    };

    if (TW1_TAB) {
        // W_M^(b*c) = W_N^(b*c*R0), c = 1..15, b = 0..15 (published by the barrier below / the first in-loop barrier)
        for (int i = tid; i < 15 * 16; i += C::THREADS) {
            const int c = i / 16 + 1, b = i % 16;
            float2 w = p.twN[(b * c * C::R0) % C::N];
            if (INVERSE) w.y = -w.y;
            if (FOLD) {
                double turns = ph_b_turns * double(C::R0 * c);
                turns -= floor(turns);
                double sn, cs;
                sincospi(2.0 * turns, &sn, &cs);
                w = make_float2(float(double(w.x) * cs - double(w.y) * sn), float(double(w.x) * sn + double(w.y) * cs));
            }
            tw1_tab[i] = w;
        }
    }
    if (TMA) {
        if (threadIdx.x == 0) {
            for (int s = 0; s < NBAR; ++s) mbar_init(&bars[s], 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int s = 0; s < STAGES; ++s) {
                const long long tile = blockIdx.x + (long long)s * gridDim.x;
                if (tile < ntiles) issue(tile, s, s);
            }
        }
    }

    int it = grp;   // the CTA's tile sequence number: group `grp` takes it = grp, grp + GROUPS, ...
    for (long long tile = blockIdx.x + (long long)grp * gridDim.x; tile < ntiles; tile += (long long)GROUPS * gridDim.x, it += GROUPS) {
        const int slot = it % STAGES;
        const long long spec = tile * C::SPB + g;
        const bool valid = spec < p.batch;
        float2* my_slot = ring + (size_t(slot) * C::SPB + g) * SLOT;
        float2* my_B = IPB ? my_slot : Bbuf + size_t(g) * C::SIZE_B;

        const int rs = grp * 2 + ((it / GROUPS) & 1);
        if (PRUNE && tid == 0) run_s[rs] = *reinterpret_cast<volatile float*>(p.run_max2);
        if (TMA) {
            mbar_wait(&bars[it % NBAR], (it / NBAR) & 1);
        } else {
            // plain-load path (unaligned base or odd n_in): cooperative coalesced copy of the tile's rows
            const long long s0 = tile * C::SPB;
            const int nvalid = int((p.batch - s0) < C::SPB ? (p.batch - s0) : C::SPB);
            bsync();   // previous tile's readers of this slot are done
            for (int idx = tid; idx < nvalid * n_in; idx += C::THREADS) {
                const int r = idx / n_in, k = idx - r * n_in;
                ring[(size_t(slot) * C::SPB + r) * SLOT + k] = in_base[(s0 + r) * n_in + k];
            }
            bsync();
        }

        // ---- stage 0: load (zero-fill + window), R0-point DFTs, in-place exchange A --------------------
        float2 v[C::E];
        stage0_load<C, WIN>(t, my_slot, (F || valid) ? n_in : 0, pad_left, in_shift, p.scale, p.win, wcol, p.win_rows, v);
        if (need_load_barrier) bsync();
        if (PRUNE) {
            // level-0 bound on the windowed, zero-filled samples this thread already holds
            float l1 = 0.f;
#pragma unroll
            for (int i = 0; i < C::E; ++i) l1 += sqrt_approx(fmaf(v[i].x, v[i].x, v[i].y * v[i].y));
            constexpr int LANES0 = C::T < 32 ? C::T : 32;
#pragma unroll
            for (int off = LANES0 / 2; off > 0; off >>= 1) l1 += __shfl_xor_sync(0xffffffffu, l1, off);
            if (C::T > 32) {
                constexpr int WPG0 = C::T / 32;
                float* rv = red + size_t(g + grp * C::SPB) * 64;
                if ((t & 31) == 0) rv[t >> 5] = l1;
                bsync();
                l1 = rv[0];
#pragma unroll
                for (int w = 1; w < WPG0; ++w) l1 += rv[w];
            }
            const bool skip_mine = !valid || (l1 * l1 * 1.0001f < run_s[rs]);
            // (the barrier also orders thread 0's write of run_s and every thread's reads of the landing slot)
            if (bsync_and(skip_mine)) {
                if (t == 0 && valid) p.absmax[spec] = 0.f;
                if (TMA && tid == 0) {
                    const long long nt = tile + (long long)STAGES * gridDim.x;
                    if (nt < ntiles) {
                        fence_proxy_async_smem();
                        issue(nt, slot, (it + STAGES) % NBAR);
                    }
                }
                continue;
            }
        }
        stage0_store<C, INVERSE, TW_PERSIST, ZF>(t, my_slot, v, tw_persist, tw0_base);
        bsync();
        // ---- stage 1: R1-point DFTs, exchange B ----------------------------------------------------------
        if (IPB) {
            stage1_load<C>(t, my_slot, v);
            bsync();                       // exchange B overwrites exchange A: every thread has its inputs
            stage1_store<C, INVERSE, TW1_TAB>(t, my_B, v, tw1_base, tw1_tab);
        } else {
            stage1<C, INVERSE, TW1_TAB>(t, my_slot, my_B, tw1_base, tw1_tab);
        }
        bsync();
        // separate buffers: the input slot is free again -> prefetch tile it+STAGES while stage 2 and the epilogue run
        if (!IPB && TMA && tid == 0) {
            const long long nt = tile + (long long)STAGES * gridDim.x;
            if (nt < ntiles) {
                fence_proxy_async_smem();
                issue(nt, slot, (it + STAGES) % NBAR);
            }
        }
        // ---- stage 2: R2-point DFTs, results in registers ------------------------------------------------
        stage2<C, INVERSE>(t, my_B, v);
        if (IPB) {
            bsync();                       // the shared buffer is free once every thread holds its results
            if (!BULKST && TMA && tid == 0) {
                const long long nt = tile + (long long)STAGES * gridDim.x;
                if (nt < ntiles) {
                    fence_proxy_async_smem();
                    issue(nt, slot, (it + STAGES) % NBAR);
                }
            }
        }

        // ---- epilogue --------------------------------------------------------------------------------------
        constexpr int Q = C::R0 * C::R1;
        if (do_stats) {
            float best = -1.f;
            int besti = 0x7fffffff;
            if (do_index) {
#pragma unroll
                for (int j = 0; j < C::C2; ++j)
#pragma unroll
                    for (int d = 0; d < C::R2; ++d) {
                        const float2 x = v[j * C::R2 + d];
                        const float m2 = x.x * x.x + x.y * x.y;
                        const int m = (t + C::T * j + Q * d + out_shift) & (C::N - 1);
                        amax_combine(best, besti, m2, m);
                    }
            } else {
#pragma unroll
                for (int i = 0; i < C::E; ++i) best = fmaxf(best, v[i].x * v[i].x + v[i].y * v[i].y);
            }
            constexpr int LANES = C::T < 32 ? C::T : 32;
#pragma unroll
            for (int off = LANES / 2; off > 0; off >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, off);
                const int oi = __shfl_xor_sync(0xffffffffu, besti, off);
                amax_combine(best, besti, ov, oi);
            }
            if (F && !do_index) {
                // max-only fast pass: one atomicMax per warp on the per-spectrum slot (zeroed by the launcher) instead of a
                // block-wide reduction + barrier; non-negative floats order like their bit patterns.
                if ((t & 31) == 0 && valid) {
                    atomicMax(reinterpret_cast<int*>(p.absmax + spec), __float_as_int(sqrtf(best)));
                }
            } else if (C::T > 32) {
                constexpr int WPG = C::T / 32;
                float* rv = red + size_t(g + grp * C::SPB) * 64;
                int* ri = reinterpret_cast<int*>(rv) + 32;
                if ((t & 31) == 0) { rv[t >> 5] = best; ri[t >> 5] = besti; }
                bsync();
                if (t == 0) {
                    for (int w = 1; w < WPG; ++w) amax_combine(best, besti, rv[w], ri[w]);
                }
            }
            if (!(F && !do_index) && t == 0 && valid) {
                p.absmax[spec] = sqrtf(best);
                if (do_index) p.argmax[spec] = besti;
                if (PRUNE) atomicMax(reinterpret_cast<int*>(p.run_max2), __float_as_int(best));
            }
        }
        if (BULKST) {
            // stage the spectrum (final order) in the free shared buffer; ONE bulk copy writes it (32 STG per thread in a burst
            // were 18 % of this kernel's stall samples, profiles/k1_8192_full_r2.md), and the buffer is re-armed with the next
            // load as soon as the copy engine has read it
            float2* stg = my_slot;
#pragma unroll
            for (int j = 0; j < C::C2; ++j)
#pragma unroll
                for (int d = 0; d < C::R2; ++d) {
                    constexpr int NH = C::N / 2;
                    const int kc = C::T * j + Q * d;
                    const int m = t + (kc >= NH ? kc - NH : kc + NH);
                    float2 x = v[j * C::R2 + d];
                    if (FOLD) {
                        // explicit rounding order: the host-parameter and the device-parameter variants must agree bit for bit
                        // (left to the compiler, the two instantiations contract this product differently)
                        const float2 w = PHDEV ? ph_tab[16 + d] : p.ph_fold[d];
                        x = make_float2(fmaf(x.x, w.x, -__fmul_rn(x.y, w.y)), fmaf(x.x, w.y, __fmul_rn(x.y, w.x)));
                    }
                    stg[m] = x;
                }
            bsync();
            if (tid == 0) {
                fence_proxy_async_smem();
                bulk_s2g(p.out + spec * (long long)C::N, stg, uint32_t(C::N) * 8u);
                bulk_commit();
                bulk_wait_read<0>();
                const long long nt = tile + (long long)STAGES * gridDim.x;
                if (nt < ntiles) issue(nt, slot, (it + STAGES) % NBAR);
            }
        } else if (do_store && valid) {
            float2* dst = p.out + spec * (long long)C::N;
#pragma unroll
            for (int j = 0; j < C::C2; ++j)
#pragma unroll
                for (int d = 0; d < C::R2; ++d) {
                    const int k = t + C::T * j + Q * d;
                    // fast variants (out_shift == N/2): T*j + Q*d is a compile-time multiple of T and t < T never crosses
                    // N/2, so the wrap is decided at compile time and every store address is `dst + t + immediate`
                    constexpr int NH = C::N / 2;
                    const int kc = C::T * j + Q * d;
                    const int m = F ? t + (kc >= NH ? kc - NH : kc + NH) : ((k + out_shift) & (C::N - 1));
                    float2 x = v[j * C::R2 + d];
                    if (FOLD) {
                        x = cmul(x, PHDEV ? ph_tab[16 + d] : p.ph_fold[d]);
                    } else if (do_phase) {
                        // m = q + Q*d' with d' = m / Q: rot = base(q) * step(d')
                        const float2 r = cmul(ph_base[j], PH_TABLE ? ph_tab[(m / Q) & 15] : p.ph_step[(m / Q) & 15]);
                        x = cmul(x, r);
                    }
                    st_stream(dst + m, x);
                }
        }
    }
    if (BULKST && tid == 0) bulk_wait_all<0>();
}

}  // namespace xmr
