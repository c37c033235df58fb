// Which pipe does cvt.rn.f16x2.f32 (F2FP.PACK_AB) use on B200, and does it contend with MUFU.EX2?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
template <int NEX, int NCVT, int NMAX>
__global__ void __launch_bounds__(256) mix(int iters, float* sink, long long* cyc) {
    float x[8]; uint32_t h[8]; float m[8];
    for (int i = 0; i < 8; ++i) { x[i] = -0.001f * (threadIdx.x + i); h[i] = i; m[i] = i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NEX; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
#pragma unroll
        for (int i = 0; i < NCVT; ++i) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(x[i]), "f"(m[i]));
#pragma unroll
        for (int i = 0; i < NMAX; ++i) asm volatile("max.f32 %0, %0, %1;" : "+f"(m[i & 7]) : "f"(x[i & 7]));
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += x[i] + __uint_as_float(h[i]) + m[i];
    if (s == 12345.f) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int NEX, int NCVT, int NMAX> void run(float* sink, long long* cyc, const char* name) {
    const int iters = 4000;
    mix<NEX, NCVT, NMAX><<<148, 256>>>(iters, sink, cyc);
    mix<NEX, NCVT, NMAX><<<148, 256>>>(iters, sink, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s %6.1f clk per iteration (2 warps/SMSP)\n", name, (double)c / iters);
}
int main() {
    float* sink; long long* cyc; cudaMalloc(&sink, 16); cudaMalloc(&cyc, 8);
    run<8, 0, 0>(sink, cyc, "8 ex2");
    run<0, 8, 0>(sink, cyc, "8 cvt.f16x2");
    run<8, 4, 0>(sink, cyc, "8 ex2 + 4 cvt.f16x2");
    run<8, 8, 0>(sink, cyc, "8 ex2 + 8 cvt.f16x2");
    run<0, 0, 8>(sink, cyc, "8 max");
    run<8, 0, 8>(sink, cyc, "8 ex2 + 8 max");
    return 0;
}
