// tcgen05.mma latency/throughput for the small shapes of the attention kernel (single CTA, operands resident in smem).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/probe_mma_lat tools/probe_mma_lat.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
// mode 0: `n` MMAs accumulating into the SAME TMEM columns; mode 1: round-robin over 8 different accumulators
__global__ void __launch_bounds__(128) lat(int N, int n, int mode, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_s;
    __shared__ __align__(8) uint64_t bar_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    const uint32_t bar = smem_u32(&bar_s);
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(bar)); asm volatile("fence.mbarrier_init.release.cluster;\n"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(&tmem_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_s;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t sa = smem_u32(smem), sb = sa + 32768;
        for (int rep = 0; rep < 3; ++rep) {
            long long t0 = clock64();
            for (int i = 0; i < n; ++i) {
                const int k = i & 3;
                const uint32_t d = tmem + (mode == 1 ? (i & 7) * (N <= 64 ? 64 : 0) : 0);
                umma(d, make_desc(sa + k * 4096, 2048, 128), make_desc(sb + k * 2 * (N / 8) * 128, (N / 8) * 128, 128), idesc, 1);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(bar) : "memory");
            long long t1 = clock64();
            wait_bar(bar, rep & 1);
            long long t2 = clock64();
            if (rep == 2) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "r"(512) : "memory");
}
int main() {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(lat, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const int Ns[] = {32, 32, 32, 32, 96, 96, 160, 256, 256};
    const int ns[] = {1, 6, 10, 10, 2, 10, 2, 1, 10};
    const int ms[] = {0, 0, 0, 1, 0, 0, 0, 0, 0};
    for (int i = 0; i < 9; ++i) {
        lat<<<1, 128, 65536>>>(Ns[i], ns[i], ms[i], d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("M=128 N=%3d K=16  x%2d %s: issue %5lld clk, complete %5lld clk (%.0f clk/MMA)\n", Ns[i], ns[i], ms[i] ? "independent accumulators" : "same accumulator        ", h[0], h[1], (double)h[1] / ns[i]);
    }
    return 0;
}
