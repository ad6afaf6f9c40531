// Register-resident small DFTs (radix 1..32) used by every stage of the shared-memory FFT.
//
// Everything here is __host__ __device__ so that the same index algebra can be executed on the CPU by
// csrc/host_emul.cpp (a thread-by-thread emulation used by the CPU test-suite; there is no GPU in the
// build container).  All loops have compile-time bounds and are fully unrolled: the arrays live in
// registers and every twiddle is an immediate.
//
// Convention: forward transform, X[k] = sum_n x[n] * exp(-2*pi*i*n*k/R)   (numpy.fft.fft; the reference calls
// np.fft.fftn(..., norm="ortho") at processing/fourier.py:153 -- the 1/sqrt(N) is folded into the window).
// INVERSE=true conjugates the twiddles (np.fft.ifft without its 1/N).
#pragma once

#if defined(__CUDACC__)
#define XMR_HD __host__ __device__ __forceinline__
#define XMR_UNROLL _Pragma("unroll")
#else
#define XMR_HD inline
#define XMR_UNROLL
#include <cmath>
struct float2 { float x, y; };
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
#endif

namespace xmr {

// cos(2*pi*k/32), k = 0..8 (first octant+); everything else by symmetry.  Rounded from float64.
XMR_HD float cos32_oct(int k) {
    switch (k) {
        case 0: return 1.0f;
        case 1: return 0.98078528040323043f;
        case 2: return 0.92387953251128674f;
        case 3: return 0.83146961230254524f;
        case 4: return 0.70710678118654757f;
        case 5: return 0.55557023301960218f;
        case 6: return 0.38268343236508978f;
        case 7: return 0.19509032201612825f;
        default: return 0.0f;  // k == 8
    }
}
// cos(2*pi*k/32) for any k in [0, 32)
XMR_HD float cos32(int k) {
    k &= 31;
    if (k > 16) k = 32 - k;          // cos is even
    if (k > 8) return -cos32_oct(16 - k);
    return cos32_oct(k);
}
XMR_HD float sin32(int k) { return cos32((k + 24) & 31); }  // sin(x) = cos(x - pi/2)

#if defined(__CUDA_ARCH__) && defined(XMR_PACKED_F32X2)
// Blackwell packed FP32: one FADD2 adds both components of a complex number (same FP32 pipe time as two FADDs -- measured
// tools/ubench/packed_fp32.cu: 36.9 vs 36.1 T results/s -- but ONE issue slot).
__device__ __forceinline__ unsigned long long pack2(float2 a) {
    unsigned long long u;
    asm("mov.b64 %0, {%1, %2};" : "=l"(u) : "f"(a.x), "f"(a.y));
    return u;
}
__device__ __forceinline__ float2 unpack2(unsigned long long u) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(u));
    return r;
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pack2(a)), "l"(pack2(b)));
    return unpack2(r);
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pack2(a)), "l"(pack2(b)));
    return unpack2(r);
}
#else
XMR_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
XMR_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#endif
XMR_HD float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
XMR_HD float2 cmul_conj(float2 a, float2 b) {  // a * conj(b)
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
XMR_HD float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }

// a * exp(-/+ 2*pi*i*k/32) with k a compile-time constant after unrolling; trivial angles cost no multiply.
template <bool INVERSE>
XMR_HD float2 mul_w32(float2 a, int k) {
    k &= 31;
    if (k == 0) return a;
    if (k == 16) return make_float2(-a.x, -a.y);
    if (k == 8) return INVERSE ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);    // * -i  (forward)
    if (k == 24) return INVERSE ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);   // * +i  (forward)
    const float c = cos32(k);
    const float s = INVERSE ? sin32(k) : -sin32(k);   // forward: exp(-i th) = c - i s
    if (k == 4 || k == 12 || k == 20 || k == 28) {
        // |c| == |s| == sqrt(1/2): two adds and two multiplies
        const float h = 0.70710678118654757f;
        const float sc = (c > 0.f) ? 1.f : -1.f, ss = (s > 0.f) ? 1.f : -1.f;
        // (a.x + i a.y) * h * (sc + i ss)
        return make_float2(h * (sc * a.x - ss * a.y), h * (ss * a.x + sc * a.y));
    }
    return make_float2(a.x * c - a.y * s, a.x * s + a.y * c);
}

XMR_HD constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v >> 1); }
XMR_HD constexpr int bitrev(int v, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1) << (bits - 1 - i);
    return r;
}

// In-place radix-2 decimation-in-frequency DFT of R points held at v[off + i*stride], i = 0..R-1.
// Output k is left at position bitrev(k): callers read v[off + bitrev(k)*stride]  (free: indices are immediates).
// NZ < R: only inputs 0 .. NZ-1 can be non-zero (input zero-filled at the end by a factor R/NZ): as long as a butterfly's
// partner lies in the zero part the layer degenerates to a copy and a twiddle multiply (a + 0, (a - 0) * w), and within
// every half-block again only the first NZ entries are non-zero.
template <int R, bool INVERSE, int STRIDE = 1, int NZ = R>
XMR_HD void dft_dif(float2* v, int off = 0) {
    static_assert(R >= 1 && R <= 32 && (R & (R - 1)) == 0, "radix must be a power of two <= 32");
    static_assert(NZ >= 1 && NZ <= R && (NZ & (NZ - 1)) == 0, "NZ must be a power of two <= R");
    XMR_UNROLL
    for (int len = R; len >= 2; len >>= 1) {
        const int half = len >> 1;
        XMR_UNROLL
        for (int blk = 0; blk < R; blk += len) {
            if (half >= NZ) {
                XMR_UNROLL
                for (int j = 0; j < NZ; ++j)
                    v[off + (blk + j + half) * STRIDE] = mul_w32<INVERSE>(v[off + (blk + j) * STRIDE], j * (32 / len));
            } else {
                XMR_UNROLL
                for (int j = 0; j < half; ++j) {
                    const int i0 = off + (blk + j) * STRIDE, i1 = off + (blk + j + half) * STRIDE;
                    const float2 a = v[i0], b = v[i1];
                    v[i0] = cadd(a, b);
                    v[i1] = mul_w32<INVERSE>(csub(a, b), j * (32 / len));
                }
            }
        }
    }
}

}  // namespace xmr
