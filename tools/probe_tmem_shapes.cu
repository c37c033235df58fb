// Register <-> (lane, column) mapping of tcgen05.ld / tcgen05.st .16x256b on B200 (sm_100a): TMEM is filled through the
// 32x32b shape with value = lane * 1000 + column, read back with 16x256b.x4 and the mapping of every thread is printed;
// then the reverse (st.16x256b, ld.32x32b) is checked.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/probe_tmem_shapes tools/probe_tmem_shapes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128) probe(float* out, int* bad) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;\n" :: "r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tq = tmem_slot + ((uint32_t)(warp * 32) << 16);
    uint32_t v[16];
    for (int c0 = 0; c0 < 64; c0 += 16) {
        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint((float)((warp * 32 + lane) * 1000 + c0 + j));
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
                     :: "r"(tq + c0), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    // read: lanes 16 h .. 16 h + 15 of this warp's quarter, columns 32 b .. 32 b + 31
    for (int h = 0; h < 2; ++h)
        for (int b = 0; b < 2; ++b) {
            asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(tq + ((uint32_t)(16 * h) << 16) + 32 * b) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            for (int j = 0; j < 16; ++j) {
                const float f = __uint_as_float(v[j]);
                out[((threadIdx.x * 2 + h) * 2 + b) * 16 + j] = f;
                // expected: register 4 g + 2 rr + e <-> lane 32 warp + 16 h + 8 rr + lane / 4, column 32 b + 8 g + 2 (lane % 4) + e
                const int g = j >> 2, rr = (j >> 1) & 1, e = j & 1;
                const float want = (float)((warp * 32 + 16 * h + 8 * rr + lane / 4) * 1000 + 32 * b + 8 * g + 2 * (lane % 4) + e);
                if (f != want) atomicAdd(bad, 1);
            }
        }
    // reverse: st.16x256b.x4 of value = 7 + the (lane, column) it should land on, checked with ld.32x32b
    for (int h = 0; h < 2; ++h)
        for (int b = 0; b < 2; ++b) {
            for (int j = 0; j < 16; ++j) {
                const int g = j >> 2, rr = (j >> 1) & 1, e = j & 1;
                v[j] = __float_as_uint(7.f + (float)((warp * 32 + 16 * h + 8 * rr + lane / 4) * 1000 + 32 * b + 8 * g + 2 * (lane % 4) + e));
            }
            asm volatile("tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
                         :: "r"(tq + ((uint32_t)(16 * h) << 16) + 32 * b), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                            "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
        }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    for (int c0 = 0; c0 < 64; c0 += 16) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(tq + c0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        for (int j = 0; j < 16; ++j)
            if (__uint_as_float(v[j]) != 7.f + (float)((warp * 32 + lane) * 1000 + c0 + j)) atomicAdd(bad + 1, 1);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;\n" :: "r"(tmem_slot) : "memory");
}
int main() {
    float* d; int* bad;
    cudaMalloc(&d, 128 * 4 * 16 * 4); cudaMalloc(&bad, 8); cudaMemset(bad, 0, 8);
    probe<<<1, 128>>>(d, bad);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    static float h[128 * 4 * 16]; int hb[2];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(hb, bad, 8, cudaMemcpyDeviceToHost);
    printf("ld.16x256b.x4 mismatches against the expected mapping: %d; st.16x256b.x4 mismatches: %d\n", hb[0], hb[1]);
    for (int t : {0, 1, 4, 5, 33}) {
        printf("thread %d, lanes +0, columns 0..31:", t);
        for (int j = 0; j < 16; ++j) printf(" %g", h[((t * 2 + 0) * 2 + 0) * 16 + j]);
        printf("\n");
    }
    return 0;
}
