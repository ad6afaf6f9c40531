// One instantiation set of K2 (per-voxel chain) per transform length: compile with -DXMR_N=<N>.
#include "k2_launch.cuh"
#include "k2_acme.cuh"

#ifndef XMR_N
#error "compile with -DXMR_N=<transform length>"
#endif

namespace xmr {

template <int N, int METHOD> struct K2Pick {          // ROI methods: the round-1 grid + zoom kernel
    static constexpr auto kern = k2_kernel<N, METHOD>;
    static constexpr size_t smem = K2Smem<N>::TOTAL;
};
template <int N> struct K2Pick<N, METHOD_ACME> {       // ACME: series localisation + quasi-Newton on the analytic gradient
    static constexpr auto kern = k2_acme_kernel<N>;
    static constexpr size_t smem = K2aSmem<N>::TOTAL;
};

template <int N, int METHOD>
static cudaError_t launch_k2(const K2Params& p, cudaStream_t st) {
    using C = FftCfg<N>;
    auto kern = K2Pick<N, METHOD>::kern;
    constexpr size_t smem = K2Pick<N, METHOD>::smem;
    static thread_local int cached_dev = -1;
    static thread_local int ctas_per_wave = 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev != cached_dev) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return e;
        int per_sm = 0, sms = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, C::T, smem);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        ctas_per_wave = per_sm * sms;
        cached_dev = dev;
    }
    long long grid = p.batch < ctas_per_wave ? p.batch : ctas_per_wave;
    if (grid < 1) return cudaSuccess;
    kern<<<dim3((unsigned)grid), dim3(C::T), smem, st>>>(p);
    return cudaGetLastError();
}

#define XMR_CAT2(a, b) a##b
#define XMR_CAT(a, b) XMR_CAT2(a, b)

cudaError_t XMR_CAT(k2_launch_, XMR_N)(const K2Params& p, int method, cudaStream_t st) {
    switch (method) {
        case METHOD_ACME: return launch_k2<XMR_N, METHOD_ACME>(p, st);
        case METHOD_PEAK_MINIMA: return launch_k2<XMR_N, METHOD_PEAK_MINIMA>(p, st);
        default: return launch_k2<XMR_N, METHOD_POSITIVITY>(p, st);
    }
}

}  // namespace xmr
