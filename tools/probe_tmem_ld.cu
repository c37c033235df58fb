// TMEM read-port rate on B200 (sm_100a): W warps per CTA (one CTA per SM) issue back-to-back tcgen05.ld 32x32b.xN on
// TMEM columns allocated by the CTA, nothing else running.  Reports bytes per clock per SM for W = 4, 8, 16 warps
// (1, 2, 4 warps per scheduler / TMEM lane quarter) and xN = 16, 32, and the single-warp load latency.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/probe_tmem_ld tools/probe_tmem_ld.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&u)[32]) {
    if (X == 32) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
                     : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]), "=r"(u[10]),
                       "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]),
                       "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
                     : "r"(taddr) : "memory");
    } else {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                     : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]), "=r"(u[10]),
                       "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                     : "r"(taddr) : "memory");
    }
}

// mode 0: throughput (two loads in flight per warp, wait, repeat); mode 1: latency (one load, wait, dependent address)
template <int X>
__global__ void __launch_bounds__(512) tmem_ld_probe(int iters, int mode, long long* cycles, uint32_t* sink) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" :: "r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t a[32], b[32], acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (mode == 0) {
        for (int i = 0; i < iters; ++i) {
            ld<X>(tmem + ((i * 2 * X) & 255), a);
            ld<X>(tmem + ((i * 2 * X + X) & 255) + 256 * ((warp >> 2) & 1), b);
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            acc += a[0] ^ b[X - 1];
        }
    } else {
        uint32_t off = 0;
        for (int i = 0; i < iters; ++i) {
            ld<X>(tmem + off, a);
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            off = (a[0] & 1) ? 0 : 0;            // dependent (always 0)
            acc += a[1];
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" :: "r"(tmem_slot) : "memory");
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    long long* cyc; uint32_t* sink;
    CK(cudaMalloc(&cyc, sms * sizeof(long long)));
    CK(cudaMalloc(&sink, 4));
    const int iters = 4096;
    long long h[256];
    for (int x = 16; x <= 32; x += 16)
        for (int warps = 4; warps <= 16; warps *= 2) {
            for (int rep = 0; rep < 2; ++rep) {
                if (x == 16) tmem_ld_probe<16><<<sms, warps * 32>>>(iters, 0, cyc, sink);
                else tmem_ld_probe<32><<<sms, warps * 32>>>(iters, 0, cyc, sink);
                CK(cudaDeviceSynchronize());
            }
            CK(cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
            const double bytes = (double)iters * 2 * x * 4 * 32 * warps;
            printf("tcgen05.ld 32x32b.x%d, %2d warps/SM (2 loads in flight per warp): %.1f B/clk/SM (%.0f clk per load per warp)\n", x, warps,
                   bytes / (double)h[0], (double)h[0] / (iters * 2));
        }
    for (int x = 16; x <= 32; x += 16) {
        if (x == 16) tmem_ld_probe<16><<<1, 32>>>(iters, 1, cyc, sink);
        else tmem_ld_probe<32><<<1, 32>>>(iters, 1, cyc, sink);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, cyc, sizeof(long long), cudaMemcpyDeviceToHost));
        printf("tcgen05.ld 32x32b.x%d single warp, ld + wait::ld round trip: %.0f clk\n", x, (double)h[0] / iters);
    }
    printf("probe_tmem_ld done\n");
    return 0;
}
