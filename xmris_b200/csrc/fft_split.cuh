// N = 2*NH transforms as TWO interleaved NH-point transforms (even / odd samples) on two thread groups of one CTA.
//
// Why: at N = 8192 the three-stage kernel needs 32 points per thread (radix 32 x 16 x 16) at the 128-register cap and is
// compute-bound at 0.64-0.70 of the HBM roofline.  Decimation in time,
//     X[k]      = E[k] + W_N^k O[k],      X[k + NH] = E[k] - W_N^k O[k],      E = FFT_NH(x[2n]),  O = FFT_NH(x[2n+1]),
// lets group 0 (threads 0 .. T-1) run the 16 x 16 x 16 transform of the even samples and group 1 that of the odd samples with
// the register budget, occupancy and code of the 4096-point kernel (16 points per thread), plus one exchange of E / W O
// through shared memory.  The landing slot holds the raw row; group g works on its elements 2*i + g IN PLACE (each thread
// rewrites only addresses of its own column), so exchange A needs no extra buffer.
//
// Everything here is __host__ __device__: csrc/host_emul.cpp runs the same index algebra thread by thread on the CPU.
#pragma once
#include "fft_stages.cuh"

namespace xmr {

// window index bookkeeping: sub-sequence sample (n1, n2) of group g is original sample n' = 2*(M*n1 + n2) + g of the N-point
// row; with the separable window of the N-point layout, w[M*n1' + n2'] = cols[n2'] * rows[n1'] (M = 256):
//     n2' = (2*n2 + g) mod M,   n1' = 2*n1 + carry,   carry = (2*n2 + g) / M
template <class C>
XMR_HD void split_window_index(int t, int j, int g, int* n2p, int* carry) {
    const int n2 = t + C::T * j;
    const int v = 2 * n2 + g;
    *n2p = v % C::M;
    *carry = v / C::M;
}

// stage 0 load of group g: v[j*R0 + n1] = slot[2*(M*n1 + n2) + g] * wcol[j] * wrows[2*n1 + carry[j]] for the non-zero rows
// n1 < R0/ZF (input zero-filled at the end by ZF), zeros elsewhere.
template <class C, int ZF>
XMR_HD void split_stage0_load(int t, int g, const float2* slot, const float* wcol /*[C0]*/, const float* wrows /*[2*R0]*/,
                              const int* carry /*[C0]*/, float2* v /*[E]*/) {
    constexpr int NR = C::R0 / ZF;
    XMR_UNROLL
    for (int j = 0; j < C::C0; ++j) {
        const int n2 = t + C::T * j;
        XMR_UNROLL
        for (int n1 = 0; n1 < C::R0; ++n1) {
            if (n1 < NR) {
                const float2 x = slot[2 * (C::M * n1 + n2) + g];
                v[j * C::R0 + n1] = cscale(x, wcol[j] * wrows[2 * n1 + carry[j]]);
            } else {
                v[j * C::R0 + n1] = make_float2(0.f, 0.f);
            }
        }
    }
}
// in-place exchange A of group g:  A_g[k1*M + n2] lives at slot[2*(k1*M + n2) + g]
template <class C>
XMR_HD void split_stage0_write(int t, int g, float2* slot, const float2* v /*[E]*/) {
    XMR_UNROLL
    for (int j = 0; j < C::C0; ++j) {
        const int n2 = t + C::T * j;
        XMR_UNROLL
        for (int k1 = 0; k1 < C::R0; ++k1) slot[2 * (k1 * C::M + n2) + g] = v[j * C::R0 + k1];
    }
}
template <class C>
XMR_HD void split_stage1_load(int t, int g, const float2* slot, float2* v /*[E]*/) {
    XMR_UNROLL
    for (int j = 0; j < C::C1; ++j) {
        const int beta = t + C::T * j, b = beta % C::R2, k1 = beta / C::R2;
        XMR_UNROLL
        for (int a = 0; a < C::R1; ++a) v[j * C::R1 + a] = slot[2 * (k1 * C::M + C::R2 * a + b) + g];
    }
}

// after stage 2 the thread holds Xsub[q + Q*d] at x[j*R2 + d], q = t + T*j, Q = R0*R1.  Group 1 turns O into W_N^k O:
// W_N^(q + Q d) = wq[j] * W_(N/Q)^d with N/Q = 2*NH/(R0*R1) = 32 for NH = 4096 (an immediate twiddle of fft_regs.cuh).
template <class C>
XMR_HD void split_twiddle_odd(float2* x /*[E]*/, const float2* wq /*[C2]: W_N^q*/) {
    static_assert(2 * C::N / (C::R0 * C::R1) == 32, "the combine twiddle W_(N/Q)^d must be a 32nd root of unity");
    XMR_UNROLL
    for (int j = 0; j < C::C2; ++j)
        XMR_UNROLL
        for (int d = 0; d < C::R2; ++d) x[j * C::R2 + d] = mul_w32<false>(cmul(x[j * C::R2 + d], wq[j]), d);
}

}  // namespace xmr
