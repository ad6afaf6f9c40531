// CPU emulation of the K1 thread choreography -- TEST SUPPORT ONLY, never on a product path.
//
// The build container has no GPU, so the index algebra of the CUDA kernel (fft_stages.cuh: thread -> element
// maps, shared-memory layouts, twiddle conventions, in-register DFTs, fftshift-on-store) is exercised here
// thread by thread, phase by phase (each phase boundary is a __syncthreads in the kernel), in float32 exactly
// as the device does.  tests/test_host_emul.py compares it with the oracle.  Built by `make emul` into
// xmris_b200/csrc/libxmris_emul.so; the python package never loads it.
#include <cstdint>
#include <cstring>
#include <vector>

#include "fft_stages.cuh"

using namespace xmr;

template <int N, bool INVERSE, int ZF = 1>
static void emul_one(const float2* in, float2* out, int n_in, int pad_left, int in_shift, const float* wtab, float scale,
                     const float2* twN, int shift, bool tw_persist_mode) {
    using C = FftCfg<N>;
    std::vector<float2> slot(C::N > n_in ? C::N : n_in), B(C::SIZE_B);
    std::memcpy(slot.data(), in, sizeof(float2) * n_in);
    std::vector<float2> regs(size_t(C::T) * C::E);
    std::vector<float2> twp(size_t(C::T) * C::C0 * (C::R0 > 1 ? C::R0 - 1 : 1)), tw0(size_t(C::T) * C::C0 * 2),
        tw1(size_t(C::T) * C::C1 * 2);
    for (int t = 0; t < C::T; ++t)
        init_twiddles<C, INVERSE>(t, twN, &twp[size_t(t) * C::C0 * (C::R0 > 1 ? C::R0 - 1 : 1)], &tw0[size_t(t) * C::C0 * 2],
                                  &tw1[size_t(t) * C::C1 * 2]);
    // phase: stage0 loads (barrier when pad_left != 0 -- emulated by finishing all loads first)
    for (int t = 0; t < C::T; ++t) {
        if (wtab)
            stage0_load<C, 1>(t, slot.data(), n_in, pad_left, in_shift, scale, wtab, nullptr, nullptr, &regs[size_t(t) * C::E]);
        else
            stage0_load<C, 0>(t, slot.data(), n_in, pad_left, in_shift, scale, nullptr, nullptr, nullptr, &regs[size_t(t) * C::E]);
    }
    for (int t = 0; t < C::T; ++t) {
        if (tw_persist_mode)
            stage0_store<C, INVERSE, true, ZF>(t, slot.data(), &regs[size_t(t) * C::E],
                                           &twp[size_t(t) * C::C0 * (C::R0 > 1 ? C::R0 - 1 : 1)], &tw0[size_t(t) * C::C0 * 2]);
        else
            stage0_store<C, INVERSE, false, ZF>(t, slot.data(), &regs[size_t(t) * C::E],
                                            &twp[size_t(t) * C::C0 * (C::R0 > 1 ? C::R0 - 1 : 1)], &tw0[size_t(t) * C::C0 * 2]);
    }
    // __syncthreads
    for (int t = 0; t < C::T; ++t) stage1<C, INVERSE>(t, slot.data(), B.data(), &tw1[size_t(t) * C::C1 * 2]);
    // __syncthreads
    for (int t = 0; t < C::T; ++t) {
        float2 x[C::E];
        stage2<C, INVERSE>(t, B.data(), x);
        for (int j = 0; j < C::C2; ++j)
            for (int d = 0; d < C::R2; ++d) {
                const int k = t + C::T * j + C::R0 * C::R1 * d;
                out[(k + shift) % C::N] = x[j * C::R2 + d];
            }
    }
}

extern "C" int xmr_emul_fft_c64(const float* in, float* out, long long batch, int n_in, int n_out, int pad_left,
                                const float* wtab, float scale, const float* twN, int inverse, int in_shift, int shift,
                                int tw_persist) {
    const float2* fin = reinterpret_cast<const float2*>(in);
    float2* fout = reinterpret_cast<float2*>(out);
    const float2* tw = reinterpret_cast<const float2*>(twN);
    for (long long b = 0; b < batch; ++b) {
        const float2* src = fin + b * n_in;
        float2* dst = fout + b * n_out;
#define XMR_CASE(NN)                                                                                         \
    case NN:                                                                                                 \
        if (inverse) emul_one<NN, true>(src, dst, n_in, pad_left, in_shift, wtab, scale, tw, shift, tw_persist != 0);  \
        else emul_one<NN, false>(src, dst, n_in, pad_left, in_shift, wtab, scale, tw, shift, tw_persist != 0);         \
        break;
        switch (n_out) {
            XMR_CASE(16) XMR_CASE(32) XMR_CASE(64) XMR_CASE(128) XMR_CASE(256) XMR_CASE(512) XMR_CASE(1024)
            XMR_CASE(2048) XMR_CASE(4096) XMR_CASE(8192)
            default: return 2;
        }
#undef XMR_CASE
    }
    return 0;
}

// The zero-fill fast variants (K1_FAST_ZF2 / ZF4): input of n_out/zf points zero-filled at the end; the degenerate first
// butterfly layers of dft_dif<R0, ..., NZ = R0/zf> run in place of the full ones.
extern "C" int xmr_emul_fft_zf_c64(const float* in, float* out, long long batch, int n_out, int zf, const float* wtab,
                                   float scale, const float* twN, int tw_persist) {
    const float2* fin = reinterpret_cast<const float2*>(in);
    float2* fout = reinterpret_cast<float2*>(out);
    const float2* tw = reinterpret_cast<const float2*>(twN);
    const int n_in = n_out / zf;
    for (long long b = 0; b < batch; ++b) {
        const float2* src = fin + b * n_in;
        float2* dst = fout + b * n_out;
#define XMR_CASE(NN, ZZ)                                                                                     \
    if (n_out == NN && zf == ZZ) {                                                                           \
        emul_one<NN, false, ZZ>(src, dst, n_in, 0, 0, wtab, scale, tw, NN / 2, tw_persist != 0);             \
        continue;                                                                                            \
    }
        XMR_CASE(512, 2) XMR_CASE(1024, 2) XMR_CASE(2048, 2) XMR_CASE(4096, 2) XMR_CASE(8192, 2)
        XMR_CASE(1024, 4) XMR_CASE(2048, 4) XMR_CASE(4096, 4) XMR_CASE(8192, 4)
#undef XMR_CASE
        return 2;
    }
    return 0;
}
