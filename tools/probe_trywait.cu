// How does mbarrier.try_wait behave on B200?  Warp 1 arrives on an mbarrier D cycles after the start; thread 0 of warp 0 waits
// for it with (a) try_wait without a suspend-time hint in a loop, (b) try_wait with a hint of H ns, (c) test_wait + nanosleep.
// Reports: loop iterations until completion, wake-up latency (cycles from the arrive to the waiter's exit).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/probe_trywait tools/probe_trywait.cu && /tmp/probe_trywait
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void probe(long long delay, uint32_t hint, long long* out) {
    __shared__ __align__(8) unsigned long long bar;
    __shared__ long long t_arrive;
    const uint32_t b = smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;\n");
    }
    __syncthreads();
    const long long t0 = clock64();
    if (threadIdx.x == 32) {
        while (clock64() - t0 < delay) {}
        t_arrive = clock64();
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(b) : "memory");
    } else if (threadIdx.x == 0) {
        uint32_t done = 0;
        long long iters = 0;
        while (!done) {
            if (MODE == 0) {
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(b), "r"(0u) : "memory");
            } else if (MODE == 1) {
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(b), "r"(0u), "r"(hint) : "memory");
            } else {
                asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(b), "r"(0u) : "memory");
                if (!done) asm volatile("nanosleep.u32 %0;\n" :: "r"(hint));
            }
            ++iters;
        }
        const long long t1 = clock64();
        out[0] = iters;
        out[1] = t1 - t_arrive;
        out[2] = t1 - t0;
    }
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 64);
    const long long delays[] = {2000, 20000, 200000};
    for (long long d : delays) {
        probe<0><<<1, 64>>>(d, 0, out); cudaDeviceSynchronize();
        printf("delay %7lld  try_wait (no hint):        iterations %6lld  wake latency %5lld cycles  (%lld cycles per iteration)\n", d, out[0], out[1], out[2] / out[0]);
        for (uint32_t h : {100u, 1000u, 10000u, 1000000u}) {
            probe<1><<<1, 64>>>(d, h, out); cudaDeviceSynchronize();
            printf("delay %7lld  try_wait hint %7u ns:  iterations %6lld  wake latency %5lld cycles\n", d, h, out[0], out[1]);
        }
        for (uint32_t h : {32u, 200u, 1000u}) {
            probe<2><<<1, 64>>>(d, h, out); cudaDeviceSynchronize();
            printf("delay %7lld  test_wait + nanosleep %4u: iterations %6lld  wake latency %5lld cycles\n", d, h, out[0], out[1]);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
