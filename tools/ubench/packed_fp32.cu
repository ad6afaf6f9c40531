// Microbenchmark: issue rate of scalar FADD/FFMA against the packed add.f32x2 / fma.rn.f32x2 (SASS FADD2 / FFMA2) on sm_100a.
// Decides whether packed butterflies can lower K1's issue pressure.  Build: nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 16

template <int MODE>
__global__ void __launch_bounds__(256) bench(float* out, float seed) {
    // CHAINS independent accumulators per thread (ILP), ITERS dependent steps each
    float a[2 * CHAINS];
#pragma unroll
    for (int i = 0; i < 2 * CHAINS; ++i) a[i] = seed + threadIdx.x * 1e-3f + i;
    const float b = seed * 0.5f, c = seed * 0.25f;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (MODE == 0) {            // 2 scalar FADD
                a[2 * i] += b;
                a[2 * i + 1] += c;
            } else if (MODE == 1) {     // 1 FADD2
                unsigned long long x, y;
                asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
                asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b), "f"(c));
                asm("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(y));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(x));
            } else if (MODE == 2) {     // 2 scalar FFMA
                a[2 * i] = fmaf(a[2 * i], b, c);
                a[2 * i + 1] = fmaf(a[2 * i + 1], c, b);
            } else if (MODE == 3) {     // 1 FFMA2
                unsigned long long x, y, z;
                asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
                asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b), "f"(c));
                asm("mov.b64 %0, {%1, %2};" : "=l"(z) : "f"(c), "f"(b));
                asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(y), "l"(z));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(x));
            } else if (MODE == 4) {     // radix-2 butterfly, scalar: 4 FADD
                const float x0 = a[2 * i], y0 = a[2 * i + 1];
                const int j = (i + 1) % CHAINS;
                const float x1 = a[2 * j], y1 = a[2 * j + 1];
                a[2 * i] = x0 + x1;
                a[2 * i + 1] = y0 + y1;
                a[2 * j] = x0 - x1;
                a[2 * j + 1] = y0 - y1;
                ++i;
            } else if (MODE == 5) {     // radix-2 butterfly, packed: FADD2 + FADD2 (sub = add of negated: fma with -1)
                const int j = (i + 1) % CHAINS;
                unsigned long long x, y, s, d;
                asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
                asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(a[2 * j]), "f"(a[2 * j + 1]));
                asm("add.rn.f32x2 %0, %1, %2;" : "=l"(s) : "l"(x), "l"(y));
                asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(x), "l"(y));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(s));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * j]), "=f"(a[2 * j + 1]) : "l"(d));
                ++i;
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 2 * CHAINS; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double flops_per_iter_thread, int ctas_per_sm) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * ctas_per_sm;
    float* out;
    cudaMalloc(&out, sizeof(float) * grid * 256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    bench<MODE><<<grid, 256>>>(out, 1.0f);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        bench<MODE><<<grid, 256>>>(out, 1.0f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double flop = flops_per_iter_thread * ITERS * double(grid) * 256.0;
    printf("%-28s ctas/sm=%d  %.3f ms  %.2f TFLOP/s (fp32 results/s)\n", name, ctas_per_sm, best, flop / best * 1e-9);
    cudaFree(out);
}

int main() {
    for (int c : {2, 4, 8}) {
        run<0>("scalar FADD", 2.0 * CHAINS, c);
        run<1>("packed add.f32x2", 2.0 * CHAINS, c);
        run<2>("scalar FFMA", 4.0 * CHAINS, c);
        run<3>("packed fma.f32x2", 4.0 * CHAINS, c);
        run<4>("butterfly scalar (4 FADD)", 2.0 * CHAINS, c);
        run<5>("butterfly packed (2 FADD2)", 2.0 * CHAINS, c);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
