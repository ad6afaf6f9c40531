// Thread-level stages of the three-pass shared-memory FFT, N = R0 * R1 * R2 (R1 = R2 = 16 for N >= 256).
//
// Index algebra (decimation in time, derived in DESIGN.md "K1"):
//   input  n  = M*n1 + n2            n1 < R0, n2 < M = R1*R2
//   stage0: for every column n2:      Y[k1][n2]   = W_N^(n2*k1) * sum_n1 x[M*n1+n2] W_R0^(n1*k1)
//           n2 = R2*a + b             a < R1, b < R2
//   stage1: for every (k1, b):        Z[k1][c][b] = W_M^(b*c)   * sum_a  Y[k1][R2*a+b] W_R1^(a*c)
//   stage2: for every (k1, c):        X[k1 + R0*(c + R1*d)] = sum_b Z[k1][c][b] W_R2^(b*d)
// Each of the T = N/E threads owns E points: E/R0 columns in stage0, E/R1 butterflies in stage1, E/R2 in stage2.
//   stage0 thread t, column j : n2   = t + T*j      -> global/shared reads are contiguous across threads
//   stage1 thread t, item   j : beta = t + T*j = b + R2*k1
//   stage2 thread t, item   j : q    = t + T*j = k1 + R0*c ; outputs X[q + R0*R1*d] contiguous across threads
// Shared-memory layouts (complex64 elements):
//   exchange A (in place in the input slot):  A[k1*M + n2]
//   exchange B:                               B[c*PC + b*PB + k1],  PB odd >= R0, PC = R2*PB padded to == R0 (mod 16)
//   both are bank-conflict free for 64-bit accesses (a half-warp touches 16 distinct 8-byte bank pairs).
//
// All functions are __host__ __device__: csrc/host_emul.cpp runs them thread by thread on the CPU.
#pragma once
#include "fft_regs.cuh"

namespace xmr {

template <int N_>
struct FftCfg {
    static_assert(N_ >= 16 && N_ <= 8192 && (N_ & (N_ - 1)) == 0, "N must be a power of two in [16, 8192]");
    static constexpr int N = N_;
    static constexpr int R2 = 16;
    static constexpr int R1 = (N >= 256) ? 16 : N / 16;
    static constexpr int R0 = (N >= 256) ? N / 256 : 1;
    static constexpr int M = R1 * R2;
    static constexpr int E = (R0 > 16) ? R0 : 16;   // points per thread
    static constexpr int T = N / E;                  // threads per spectrum
    static constexpr int PB = R0 | 1;                // odd, >= R0 (R0 is a power of two)
    static constexpr int PC = ((R2 * PB - R0 + 15) / 16) * 16 + R0;  // >= R2*PB and == R0 (mod 16)
    static constexpr int SIZE_B = R1 * PC;
    static constexpr int C0 = E / R0, C1 = E / R1, C2 = E / R2;      // items per thread per stage
    static constexpr int SPB = (T >= 256) ? 1 : ((256 / T) > 32 ? 32 : (256 / T));  // spectra per block
    static constexpr int THREADS = T * SPB;
};

// Powers w^1 .. w^(R-1) of a unit complex number from its exact first and fourth powers (depth <= 4 products).
template <int R>
XMR_HD void twiddle_powers(float2 w1, float2 w4, float2* w /* [R], w[0] unused */) {
    if (R > 1) w[1] = w1;
    XMR_UNROLL
    for (int k = 2; k < R; ++k) {
        if (k == 4) w[k] = w4;
        else if (k < 4) w[k] = cmul(w[k - 1], w1);
        else if ((k & 3) == 0) w[k] = cmul(w[k - 4], w4);
        else w[k] = cmul(w[k & ~3], w[k & 3]);
    }
}

// ---- stage 0 -------------------------------------------------------------------------------------------
// Loads the thread's columns from the input slot (implicit zero fill, window), R0-point DFTs, inter-stage
// twiddle, in-place store as exchange A.  `slot` holds the raw FID at [0, n_in); `wcol[j]` is the column
// part of the window (incl. 1/sqrt(N)), `wrow[n1]` its row part (separable window), or `wtab` a full table.
template <class C, int WIN /*0 scalar, 1 table, 2 separable*/>
XMR_HD void stage0_load(int t, const float2* slot, int n_in, int pad_left, int in_shift, float scale,
                        const float* wtab, const float* wcol, const float* wrow, float2* v /* [E] */) {
    XMR_UNROLL
    for (int j = 0; j < C::C0; ++j) {
        const int n2 = t + C::T * j;
        XMR_UNROLL
        for (int n1 = 0; n1 < C::R0; ++n1) {
            const int n = C::M * n1 + n2;
            const int src = n - pad_left;
            float2 x = make_float2(0.f, 0.f);
            if (src >= 0 && src < n_in) {
                x = slot[(src + in_shift) & (C::N - 1)];   // in_shift != 0 only with n_in == N, pad_left == 0
                float w;
                if (WIN == 1) w = wtab[n];
                else if (WIN == 2) w = wcol[j] * wrow[n1];
                else w = scale;
                x = cscale(x, w);
            }
            v[j * C::R0 + n1] = x;
        }
    }
}

// When pad_left != 0 (or in_shift is not a multiple of M) the loads above touch other threads' columns: the caller
// must put a block barrier between stage0_load and stage0_store.  Otherwise each thread reads and writes the same
// set of addresses and exchange A happens in place without a barrier.
// stage0_compute leaves Y[k1][n2] (inter-stage twiddle applied) at v[j*R0 + k1]; stage0_write stores it as exchange A.
// ZF: the input is zero-filled at the end by this factor (1, 2, 4 ...; <= R0): rows n1 >= R0/ZF of every column are zero.
template <class C, bool INVERSE, bool TW_PERSIST, int ZF = 1>
XMR_HD void stage0_compute(int t, float2* v /* [E] */, const float2* tw_persist /* [C0][R0-1] */,
                           const float2* tw_base /* [C0][2] */) {
    XMR_UNROLL
    for (int j = 0; j < C::C0; ++j) {
        dft_dif<C::R0, INVERSE, 1, (C::R0 / ZF >= 1 ? C::R0 / ZF : 1)>(v + j * C::R0);
        float2 w[C::R0 > 1 ? C::R0 : 2];
        if (!TW_PERSIST && C::R0 > 1) twiddle_powers<C::R0>(tw_base[2 * j], tw_base[2 * j + 1], w);
        float2 y[C::R0];
        XMR_UNROLL
        for (int k1 = 0; k1 < C::R0; ++k1) {
            y[k1] = v[j * C::R0 + bitrev(k1, ilog2(C::R0))];
            if (k1 > 0) y[k1] = cmul(y[k1], TW_PERSIST ? tw_persist[j * (C::R0 - 1) + k1 - 1] : w[k1]);
        }
        XMR_UNROLL
        for (int k1 = 0; k1 < C::R0; ++k1) v[j * C::R0 + k1] = y[k1];
    }
}
template <class C>
XMR_HD void stage0_write(int t, float2* slot, const float2* v /* [E] */) {
    XMR_UNROLL
    for (int j = 0; j < C::C0; ++j) {
        const int n2 = t + C::T * j;
        XMR_UNROLL
        for (int k1 = 0; k1 < C::R0; ++k1) slot[k1 * C::M + n2] = v[j * C::R0 + k1];
    }
}
template <class C, bool INVERSE, bool TW_PERSIST, int ZF = 1>
XMR_HD void stage0_store(int t, float2* slot, float2* v /* [E] */, const float2* tw_persist /* [C0][R0-1] */,
                         const float2* tw_base /* [C0][2] */) {
    stage0_compute<C, INVERSE, TW_PERSIST, ZF>(t, v, tw_persist, tw_base);
    stage0_write<C>(t, slot, v);
}

// ---- stage 1 -------------------------------------------------------------------------------------------
// tw1_base[j][0..1] = W_M^b, W_M^(4b) for this thread's item j (b = (t + T*j) % R2).
// Split in two so that exchange B can reuse the memory of exchange A (large N: one shared buffer per spectrum):
// the caller puts a block barrier between stage1_load and stage1_store when A and B alias.
template <class C>
XMR_HD void stage1_load(int t, const float2* A, float2* v /* [E] */) {
    XMR_UNROLL
    for (int j = 0; j < C::C1; ++j) {
        const int beta = t + C::T * j, b = beta % C::R2, k1 = beta / C::R2;
        XMR_UNROLL
        for (int a = 0; a < C::R1; ++a) v[j * C::R1 + a] = A[k1 * C::M + C::R2 * a + b];
    }
}
// tw1_tab (optional, shared memory): W_M^(b*c) at [(c-1)*R2 + b] -- 15 conflict-free 8-byte loads per butterfly instead of
// a 14-multiply power chain (the FFT is issue-bound, the LSU has headroom).  nullptr: powers of tw1_base.
// stage1_compute leaves Z[k1][c][b] at v[j*R1 + c] (natural order in c); stage1_write scatters it into exchange B.
template <class C, bool INVERSE, bool TAB = false>
XMR_HD void stage1_compute(int t, float2* v /* [E] */, const float2* tw1_base /* [C1][2] */, const float2* tw1_tab = nullptr) {
    XMR_UNROLL
    for (int j = 0; j < C::C1; ++j) {
        const int b = (t + C::T * j) % C::R2;
        dft_dif<C::R1, INVERSE>(v + j * C::R1);
        float2 w[C::R1 > 1 ? C::R1 : 2];
        if (C::R1 > 1 && !TAB) twiddle_powers<C::R1>(tw1_base[2 * j], tw1_base[2 * j + 1], w);
        float2 z[C::R1];
        XMR_UNROLL
        for (int c = 0; c < C::R1; ++c) {
            z[c] = v[j * C::R1 + bitrev(c, ilog2(C::R1))];
            if (c > 0) z[c] = cmul(z[c], TAB ? tw1_tab[(c - 1) * C::R2 + b] : w[c]);
        }
        XMR_UNROLL
        for (int c = 0; c < C::R1; ++c) v[j * C::R1 + c] = z[c];
    }
}
template <class C>
XMR_HD void stage1_write(int t, float2* B, const float2* v /* [E] */) {
    XMR_UNROLL
    for (int j = 0; j < C::C1; ++j) {
        const int beta = t + C::T * j, b = beta % C::R2, k1 = beta / C::R2;
        XMR_UNROLL
        for (int c = 0; c < C::R1; ++c) B[c * C::PC + b * C::PB + k1] = v[j * C::R1 + c];
    }
}
// maxq (optional): receives max over this thread's exchange-B values of |z|^2 (branch-and-bound pruning in K1).
template <class C, bool INVERSE, bool TAB = false, bool MAXQ = false>
XMR_HD void stage1_store(int t, float2* B, float2* v /* [E] */, const float2* tw1_base /* [C1][2] */,
                         const float2* tw1_tab = nullptr, float* maxq = nullptr) {
    stage1_compute<C, INVERSE, TAB>(t, v, tw1_base, tw1_tab);
    if (MAXQ) {
        float mq = 0.f;
        XMR_UNROLL
        for (int i = 0; i < C::E; ++i) {
            const float q = v[i].x * v[i].x + v[i].y * v[i].y;
            mq = q > mq ? q : mq;
        }
        *maxq = mq;
    }
    stage1_write<C>(t, B, v);
}
template <class C, bool INVERSE, bool TAB = false>
XMR_HD void stage1(int t, const float2* A, float2* B, const float2* tw1_base /* [C1][2] */, const float2* tw1_tab = nullptr) {
    float2 v[C::E];
    stage1_load<C>(t, A, v);
    stage1_store<C, INVERSE, TAB>(t, B, v, tw1_base, tw1_tab);
}

// ---- stage 2 -------------------------------------------------------------------------------------------
// Leaves X[q + R0*R1*d] in out[j*R2 + d]  (natural order in d), q = t + T*j.
template <class C, bool INVERSE>
XMR_HD void stage2(int t, const float2* B, float2* out /* [E] */) {
    float2 v[C::E];
    XMR_UNROLL
    for (int j = 0; j < C::C2; ++j) {
        const int q = t + C::T * j, k1 = q % C::R0, c = q / C::R0;
        XMR_UNROLL
        for (int b = 0; b < C::R2; ++b) v[j * C::R2 + b] = B[c * C::PC + b * C::PB + k1];
    }
    XMR_UNROLL
    for (int j = 0; j < C::C2; ++j) {
        dft_dif<C::R2, INVERSE>(v + j * C::R2);
        XMR_UNROLL
        for (int d = 0; d < C::R2; ++d) out[j * C::R2 + d] = v[j * C::R2 + bitrev(d, ilog2(C::R2))];
    }
}

// Per-thread twiddle set-up from the global table twN[k] = exp(-2*pi*i*k/N) (conjugated by the caller for inverse).
template <class C, bool INVERSE>
XMR_HD void init_twiddles(int t, const float2* twN, float2* tw_persist /* [C0][R0-1] or null */,
                          float2* tw0_base /* [C0][2] */, float2* tw1_base /* [C1][2] */) {
    XMR_UNROLL
    for (int j = 0; j < C::C0; ++j) {
        const int n2 = t + C::T * j;
        if (C::R0 > 1) {
            float2 a = twN[n2 % C::N], b4 = twN[(4 * n2) % C::N];
            if (INVERSE) { a.y = -a.y; b4.y = -b4.y; }
            tw0_base[2 * j] = a;
            tw0_base[2 * j + 1] = b4;
            if (tw_persist) {
                XMR_UNROLL
                for (int k1 = 1; k1 < C::R0; ++k1) {
                    float2 w = twN[(n2 * k1) % C::N];
                    if (INVERSE) w.y = -w.y;
                    tw_persist[j * (C::R0 - 1) + k1 - 1] = w;
                }
            }
        }
    }
    XMR_UNROLL
    for (int j = 0; j < C::C1; ++j) {
        const int b = (t + C::T * j) % C::R2;
        float2 a = twN[(b * C::R0) % C::N], b4 = twN[(4 * b * C::R0) % C::N];   // W_M^b = W_N^(b*R0)
        if (INVERSE) { a.y = -a.y; b4.y = -b4.y; }
        tw1_base[2 * j] = a;
        tw1_base[2 * j + 1] = b4;
    }
}

}  // namespace xmr
