// Does MUFU.EX2 block the issue port?  Per loop iteration: 8 independent ex2 + 8*K independent FFMA per thread.
// If the pipes overlap, cycles/iter ~ max(64, 8 + 8K) per warp on an SMSP (2 warps/SMSP here); if MUFU holds the
// dispatch port, cycles/iter ~ 64 + 8K.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/probe_mufu_mix tools/probe_mufu_mix.cu
#include <cuda_runtime.h>
#include <cstdio>
template <int K>
__global__ void __launch_bounds__(256) mix(int iters, float* sink, long long* cyc) {
    float x[8], y[8 * (K > 0 ? K : 1)];
    for (int i = 0; i < 8; ++i) x[i] = -0.001f * (threadIdx.x + i);
    for (int i = 0; i < 8 * (K > 0 ? K : 1); ++i) y[i] = 0.5f + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
#pragma unroll
        for (int i = 0; i < 8 * K; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(y[i]) : "f"(1.0001f), "f"(0.0001f));
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += x[i];
    for (int i = 0; i < 8 * (K > 0 ? K : 1); ++i) s += y[i];
    if (s == 12345.f) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int K> void run(float* sink, long long* cyc) {
    const int iters = 4000;
    mix<K><<<148, 256>>>(iters, sink, cyc);   // 8 warps/SM = 2 warps per SMSP
    mix<K><<<148, 256>>>(iters, sink, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("K=%d: %6.1f clk per iteration (2 warps/SMSP; MUFU-only bound 128, issue-only bound %d, serialized %d)\n", K, (double)c / iters, 2 * (8 + 8 * K), 128 + 2 * 8 * K);
}
int main() {
    float* sink; long long* cyc; cudaMalloc(&sink, 16); cudaMalloc(&cyc, 8);
    run<0>(sink, cyc); run<1>(sink, cyc); run<2>(sink, cyc); run<3>(sink, cyc); run<4>(sink, cyc); run<6>(sink, cyc); run<8>(sink, cyc);
    return 0;
}
