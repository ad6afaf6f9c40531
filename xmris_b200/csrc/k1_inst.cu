// One instantiation set of K1 per transform length: compile with -DXMR_N=<N>.
#include "k1_launch.cuh"
#include "k1_max.cuh"
#include <cstdlib>



#ifndef XMR_N
#error "compile with -DXMR_N=<transform length>"
#endif

namespace xmr {

template <int N, bool INVERSE, int WIN, bool TMA, int FAST, bool PRUNE, int GROUPS>
static cudaError_t launch_impl(const K1Params& p, int max_ctas, cudaStream_t st) {
    using C = FftCfg<N>;
    auto kern = k1_kernel<N, INVERSE, WIN, TMA, FAST, PRUNE, GROUPS>;
    constexpr size_t smem = K1Smem<N, GROUPS>::TOTAL;
    constexpr int threads = C::THREADS * GROUPS;
    static thread_local int cached_dev = -1;
    static thread_local int ctas_per_wave = 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev != cached_dev) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return e;
        int per_sm = 0, sms = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        ctas_per_wave = per_sm * sms;
        cached_dev = dev;
    }
    const long long ntiles = (p.batch + C::SPB - 1) / C::SPB;
    long long grid = ntiles < ctas_per_wave ? ntiles : ctas_per_wave;
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
    if (grid < 1) return cudaSuccess;
    kern<<<dim3((unsigned)grid), dim3(threads), smem, st>>>(p);
    return cudaGetLastError();
}

template <int N, bool INVERSE, int WIN, bool TMA, int FAST = 0, bool PRUNE = false>
static cudaError_t launch_one(const K1Params& p, int max_ctas, cudaStream_t st) {
    if constexpr (N >= 8192 && TMA) {
        // store variants: results staged in the free shared buffer and written by one bulk copy per spectrum (16-byte aligned
        // output; XMR_K1_BULKST=0 keeps the per-thread stores)
        constexpr bool CAN_BULK = (FAST & K1_FAST_STORE) != 0 && (FAST & K1_FAST_STATS) == 0;
        static const bool bulk_on = !(getenv("XMR_K1_BULKST") != nullptr && getenv("XMR_K1_BULKST")[0] == '0');
        const bool bulk = CAN_BULK && bulk_on && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
        // two 256-thread groups per CTA sharing a three-buffer ring (K1Smem); XMR_K1_GROUPS=1 selects the two-CTA form
        static const bool grouped = !(getenv("XMR_K1_GROUPS") != nullptr && getenv("XMR_K1_GROUPS")[0] == '1');
        if (grouped) {
            // (ntiles here counts spectra: SPB == 1; one CTA takes two tiles at a time)
            const int mc = max_ctas > 0 ? (max_ctas + 1) / 2 : 0;
            if constexpr (CAN_BULK) {
                if (bulk) return launch_impl<N, INVERSE, WIN, TMA, FAST | K1_FAST_BULKST, PRUNE, 2>(p, mc, st);
            }
            // per-thread stores with input a quarter of the transform: nothing to prefetch, and two separate CTAs drift into
            // opposite phases -- measured faster than the grouped form
            if constexpr ((FAST & K1_FAST_ZF4) == 0) return launch_impl<N, INVERSE, WIN, TMA, FAST, PRUNE, 2>(p, mc, st);
        }
    }
    // (up to 4096 points -- separate exchange buffers, two barriers per tile -- the bulk store needs two more barriers and
    //  measured slower: C5 pass 2 12.1-12.5 ms against 11.4-11.9 ms; profiles/rejected_r2.md)
    return launch_impl<N, INVERSE, WIN, TMA, FAST, PRUNE, 1>(p, max_ctas, st);
}

#if XMR_N >= 512
template <int ZF>
struct K1MaxPick {
    static constexpr auto kern = k1_max_zf_kernel<XMR_N, ZF>;
    static constexpr size_t smem = K1MaxZfSmem<XMR_N, ZF>::TOTAL;
    static constexpr long long spt = (long long)ZF * FftCfg<XMR_N>::SPB;     // spectra per loop iteration
};
template <>
struct K1MaxPick<1> {
    static constexpr auto kern = k1_max_kernel<XMR_N, 1>;
    static constexpr size_t smem = K1MaxSmem<XMR_N>::TOTAL;
    static constexpr long long spt = FftCfg<XMR_N>::SPB;
};

template <int ZF = 1>
static cudaError_t launch_max(const K1Params& p, cudaStream_t st) {
    using C = FftCfg<XMR_N>;
    auto kern = K1MaxPick<ZF>::kern;
    constexpr size_t smem = K1MaxPick<ZF>::smem;
    static thread_local int cached_dev = -1;
    static thread_local int ctas_per_wave = 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev != cached_dev) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return e;
        int per_sm = 0, sms = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, C::THREADS, smem);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        ctas_per_wave = (per_sm < 1 ? 1 : per_sm) * sms;
        cached_dev = dev;
    }
    const long long ntiles = (p.batch + K1MaxPick<ZF>::spt - 1) / K1MaxPick<ZF>::spt;
    const long long grid = ntiles < ctas_per_wave ? ntiles : ctas_per_wave;
    if (grid < 1) return cudaSuccess;
    kern<<<dim3((unsigned)grid), dim3(C::THREADS), smem, st>>>(p);
    return cudaGetLastError();
}
#endif



#define XMR_CAT2(a, b) a##b
#define XMR_CAT(a, b) XMR_CAT2(a, b)

cudaError_t XMR_CAT(k1_launch_, XMR_N)(const K1Params& p, bool inverse, int win, bool tma, int max_ctas,
                                       cudaStream_t st) {
    if (inverse) {
        // to_fid: no window (scale only -> separable mode with unit rows)
        return tma ? launch_one<XMR_N, true, 2, true>(p, max_ctas, st) : launch_one<XMR_N, true, 2, false>(p, max_ctas, st);
    }
    // hot path: full-length input, separable window, TMA, fftshift store -> compile-time epilogue variants
    const bool fast_geom = p.row_slot == nullptr && tma && win == 2 && p.n_in == XMR_N && p.pad_left == 0 && p.in_shift == 0 &&
                           p.out_shift == XMR_N / 2 && XMR_N >= 512;
    if (fast_geom) {
        const bool st_ = p.out != nullptr, stats = p.absmax != nullptr && p.argmax == nullptr, ph = p.phase_on != 0;
        if (st_ && !stats && !ph && p.absmax == nullptr)
            return launch_one<XMR_N, false, 2, true, K1_FAST_ON | K1_FAST_STORE>(p, max_ctas, st);
        if (st_ && ph && p.absmax == nullptr) {
            if (p.ph_dev != nullptr)
                return launch_one<XMR_N, false, 2, true, K1_FAST_ON | K1_FAST_STORE | K1_FAST_PHASE | K1_FAST_PHDEV>(p, max_ctas, st);
            return launch_one<XMR_N, false, 2, true, K1_FAST_ON | K1_FAST_STORE | K1_FAST_PHASE>(p, max_ctas, st);
        }
#if XMR_N >= 512
        constexpr bool HAS_MAX_KERNEL = true;
#else
        constexpr bool HAS_MAX_KERNEL = false;
#endif
        if (!st_ && stats && (p.run_max2 == nullptr || HAS_MAX_KERNEL)) {
            cudaError_t e = cudaMemsetAsync(p.absmax, 0, sizeof(float) * size_t(p.batch), st);   // atomicMax accumulators
            if (e != cudaSuccess) return e;
#if XMR_N >= 512
            if (p.run_max2 != nullptr) return launch_max(p, st);
#endif
            return launch_one<XMR_N, false, 2, true, K1_FAST_ON | K1_FAST_STATS>(p, max_ctas, st);
        }
    }
#if XMR_N >= 512
    // the same geometry, statistics only with a running maximum: the branch-and-bound kernel on the short rows
    if (tma && win == 2 && p.pad_left == 0 && p.in_shift == 0 && p.out_shift == XMR_N / 2 && p.out == nullptr &&
        p.absmax != nullptr && p.argmax == nullptr && p.run_max2 != nullptr && (2 * p.n_in == XMR_N || 4 * p.n_in == XMR_N)) {
        cudaError_t e = cudaMemsetAsync(p.absmax, 0, sizeof(float) * size_t(p.batch), st);
        if (e != cudaSuccess) return e;
        if (2 * p.n_in == XMR_N) return launch_max<2>(p, st);
#if XMR_N >= 1024
        return launch_max<4>(p, st);
#endif
    }
#endif
#if XMR_N >= 512
    // input zero-filled at the end to 2x / 4x its length (zero_fill's default geometry): store variants
    if (tma && win == 2 && p.pad_left == 0 && p.in_shift == 0 && p.out_shift == XMR_N / 2 && p.out != nullptr &&
        p.absmax == nullptr) {
        if (2 * p.n_in == XMR_N) {
            if (p.phase_on != 0 && p.ph_dev != nullptr)
                return launch_one<XMR_N, false, 2, true, K1_FAST_ON | K1_FAST_STORE | K1_FAST_PHASE | K1_FAST_ZF2 | K1_FAST_PHDEV>(p, max_ctas, st);
            if (p.phase_on != 0)
                return launch_one<XMR_N, false, 2, true, K1_FAST_ON | K1_FAST_STORE | K1_FAST_PHASE | K1_FAST_ZF2>(p, max_ctas, st);
            return launch_one<XMR_N, false, 2, true, K1_FAST_ON | K1_FAST_STORE | K1_FAST_ZF2>(p, max_ctas, st);
        }
#if XMR_N >= 1024
        if (4 * p.n_in == XMR_N) {
            if (p.phase_on != 0 && p.ph_dev != nullptr)
                return launch_one<XMR_N, false, 2, true, K1_FAST_ON | K1_FAST_STORE | K1_FAST_PHASE | K1_FAST_ZF4 | K1_FAST_PHDEV>(p, max_ctas, st);
            if (p.phase_on != 0)
                return launch_one<XMR_N, false, 2, true, K1_FAST_ON | K1_FAST_STORE | K1_FAST_PHASE | K1_FAST_ZF4>(p, max_ctas, st);
            return launch_one<XMR_N, false, 2, true, K1_FAST_ON | K1_FAST_STORE | K1_FAST_ZF4>(p, max_ctas, st);
        }
#endif
    }
#endif
    if (p.run_max2 != nullptr && p.out == nullptr && p.absmax != nullptr && p.argmax == nullptr) {
        // statistics-only pass of any other geometry with a running maximum: prune on the level-0 bound
        if (win == 1)
            return tma ? launch_one<XMR_N, false, 1, true, 0, true>(p, max_ctas, st)
                       : launch_one<XMR_N, false, 1, false, 0, true>(p, max_ctas, st);
        return tma ? launch_one<XMR_N, false, 2, true, 0, true>(p, max_ctas, st)
                   : launch_one<XMR_N, false, 2, false, 0, true>(p, max_ctas, st);
    }
    if (win == 1)
        return tma ? launch_one<XMR_N, false, 1, true>(p, max_ctas, st) : launch_one<XMR_N, false, 1, false>(p, max_ctas, st);
    return tma ? launch_one<XMR_N, false, 2, true>(p, max_ctas, st) : launch_one<XMR_N, false, 2, false>(p, max_ctas, st);
}

}  // namespace xmr
