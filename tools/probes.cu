// Hardware probes for B200 (sm_100a): validates the tcgen05 / TMEM programming model used by the
// v2 kernels (shared-memory descriptor semantics, TMEM load mapping) and measures the pipe rates the
// design decisions rest on (mma.sync fp16, tcgen05 kind::f16, MUFU ex2 f32 vs f16x2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/probes tools/probes.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------------------------------ tcgen05 helpers
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;     // version = 1 (Blackwell)
    return d;                   // layout_type = 0 (no swizzle), base_offset = 0
}
__device__ __forceinline__ uint32_t make_idesc_f16(int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;                      // c_format = F32
    d |= 0u << 7;                      // a_format = F16
    d |= 0u << 10;                     // b_format = F16
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    long long t0 = clock64();
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && clock64() - t0 > 2000000000LL) return false;
    }
    return true;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// ------------------------------------------------------------------------------------------ probe 1: UMMA GEMM correctness
// C[128][N] = A[128][K] . B[N][K]^T, fp16 in, fp32 out, one CTA of 128 threads.
// Canonical no-swizzle K-major layout: element (r,k) at (k/8)*kstride + (r/8)*rstride + (r%8)*16 + (k%8)*2 bytes
// (core matrix = 8 rows x 16 B, contiguous).  `swap` exchanges which descriptor field gets which stride.
template <int N, int K>
__global__ void __launch_bounds__(128) umma_gemm_probe(const __half* A, const __half* B, float* C, int swap, int* status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr int A_BYTES = 128 * K * 2, B_BYTES = N * K * 2;
    uint8_t* sa = smem;
    uint8_t* sb = smem + A_BYTES;
    constexpr uint32_t RS = 128;                 // stride between 8-row groups
    constexpr uint32_t KS_A = (128 / 8) * 128;   // stride between K chunks (A: 128 rows)
    constexpr uint32_t KS_B = (N / 8) * 128;
    for (int i = tid; i < 128 * (K / 8); i += 128) {
        const int r = i / (K / 8), c = i % (K / 8);
        *reinterpret_cast<uint4*>(sa + c * KS_A + (r / 8) * RS + (r % 8) * 16) = *reinterpret_cast<const uint4*>(A + r * K + c * 8);
    }
    for (int i = tid; i < N * (K / 8); i += 128) {
        const int r = i / (K / 8), c = i % (K / 8);
        *reinterpret_cast<uint4*>(sb + c * KS_B + (r / 8) * RS + (r % 8) * 16) = *reinterpret_cast<const uint4*>(B + r * K + c * 8);
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    const uint32_t bar = smem_u32(&bar_s);
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(&tmem_base_s)), "r"(N < 32 ? 32 : N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_f16(128, N);
        for (int k = 0; k < K / 16; ++k) {
            uint64_t ad, bd;
            if (!swap) {
                ad = make_desc(smem_u32(sa) + k * 2 * KS_A, KS_A, RS);
                bd = make_desc(smem_u32(sb) + k * 2 * KS_B, KS_B, RS);
            } else {
                ad = make_desc(smem_u32(sa) + k * 2 * KS_A, RS, KS_A);
                bd = make_desc(smem_u32(sb) + k * 2 * KS_B, RS, KS_B);
            }
            umma_f16(tmem, ad, bd, idesc, k > 0);
        }
        umma_commit(bar);
    }
    const bool ok = mbar_wait_bounded(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (!ok) { if (tid == 0) *status = -1; }
    else {
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
            for (int j = 0; j < 32; ++j) C[tid * N + c0 + j] = __uint_as_float(v[j]);
        }
        if (tid == 0) *status = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "r"(N < 32 ? 32 : N) : "memory");
}

template <int N, int K>
void run_umma_probe() {
    std::vector<__half> ha(128 * K), hb(N * K);
    std::vector<float> fa(128 * K), fb(N * K), ref(128 * N), out(128 * N);
    srand(1);
    for (int i = 0; i < 128 * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; ha[i] = __float2half(v); fa[i] = __half2float(ha[i]); }
    for (int i = 0; i < N * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hb[i] = __float2half(v); fb[i] = __half2float(hb[i]); }
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)fa[m * K + k] * fb[n * K + k]; ref[m * N + n] = (float)s; }
    __half *dA, *dB; float* dC; int* dS;
    CK(cudaMalloc(&dA, ha.size() * 2)); CK(cudaMalloc(&dB, hb.size() * 2)); CK(cudaMalloc(&dC, out.size() * 4)); CK(cudaMalloc(&dS, 4));
    CK(cudaMemcpy(dA, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
    const int smem = 128 * K * 2 + N * K * 2;
    CK(cudaFuncSetAttribute(umma_gemm_probe<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int swap = 0; swap < 2; ++swap) {
        CK(cudaMemset(dC, 0, out.size() * 4)); CK(cudaMemset(dS, 0, 4));
        umma_gemm_probe<N, K><<<1, 128, smem>>>(dA, dB, dC, swap, dS);
        cudaError_t e = cudaDeviceSynchronize();
        int st = 0;
        if (e != cudaSuccess) { printf("umma_probe N=%d K=%d swap=%d: CUDA error %s\n", N, K, swap, cudaGetErrorString(e)); exit(2); }
        CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(out.data(), dC, out.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0;
        for (size_t i = 0; i < out.size(); ++i) maxerr = fmax(maxerr, fabs((double)out[i] - ref[i]));
        printf("umma_probe M=128 N=%d K=%d desc{LBO=%s,SBO=%s}: status=%d max_abs_err=%.3e %s\n", N, K, swap ? "row-group" : "k-chunk",
               swap ? "k-chunk" : "row-group", st, maxerr, maxerr < 1e-3 ? "MATCH" : "mismatch");
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dS);
}

// ------------------------------------------------------------------------------------------ probe 2: tcgen05 throughput
// Every CTA issues `iters` x (128 x 256 x 16) fp16 MMAs back to back on resident smem operands.
__global__ void __launch_bounds__(128) umma_rate_probe(int iters, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 * 64 * 2 + 256 * 64 * 2) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    const uint32_t bar = smem_u32(&bar_s);
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    long long t0 = clock64();
    if (tid == 0) {
        const uint32_t idesc = make_idesc_f16(128, 256);
        const uint32_t sa = smem_u32(smem), sb = sa + 128 * 64 * 2;
        for (int it = 0; it < iters; ++it) {
            const int k = it & 3;
            const uint64_t ad = make_desc(sa + k * 2 * 2048, 2048, 128);
            const uint64_t bd = make_desc(sb + k * 2 * 4096, 4096, 128);
            umma_f16(tmem + (it & 1) * 256, ad, bd, idesc, 1);
        }
        umma_commit(bar);
    }
    mbar_wait_bounded(bar, 0);
    long long t1 = clock64();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "r"(512) : "memory");
}

// ------------------------------------------------------------------------------------------ probe 3: mma.sync fp16 rate
__global__ void __launch_bounds__(512) mma_sync_rate(int iters, float* sink) {
    float c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    uint32_t a0 = 0x3c003c00u + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 0x3c003c00u, b1 = b0 + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 12345.f) sink[0] = s;
}

// ------------------------------------------------------------------------------------------ probe 4: MUFU ex2 rates
__global__ void __launch_bounds__(512) ex2_rate(int iters, int mode, float* sink) {
    float x[8];
    uint32_t h[8];
    for (int i = 0; i < 8; ++i) { x[i] = -0.001f * (threadIdx.x + i); h[i] = 0xb800b800u + i; }
    for (int it = 0; it < iters; ++it) {
        if (mode == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += x[i] + __uint_as_float(h[i]);
    if (s == 12345.f) sink[0] = s;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    printf("device %s sm_%d%d SMs=%d clock=%d MHz\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount, clk_khz / 1000);
    float* sink; CK(cudaMalloc(&sink, 16));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    const int sms = prop.multiProcessorCount;

    // mma.sync rate
    for (int warm = 0; warm < 2; ++warm) {
        const int iters = 20000;
        CK(cudaEventRecord(e0));
        mma_sync_rate<<<sms * 2, 512>>>(iters, sink);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        if (warm) {
            double flop = (double)sms * 2 * 16 * iters * 8 * 4096.0;
            printf("mma.sync m16n8k16 f16: %.1f TFLOP/s (%.0f flop/clk/SM at %d MHz nominal)\n", flop / ms / 1e9, flop / (ms * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1000);
        }
    }
    // ex2 rates
    for (int mode = 0; mode < 2; ++mode) for (int warm = 0; warm < 2; ++warm) {
        const int iters = 20000;
        CK(cudaEventRecord(e0));
        ex2_rate<<<sms * 2, 512>>>(iters, mode, sink);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        if (warm) {
            double n = (double)sms * 2 * 512 * iters * 8 * (mode ? 2 : 1);
            printf("ex2.approx %s: %.2f Texp/s (%.1f exp/clk/SM)\n", mode ? "f16x2" : "f32", n / ms / 1e9, n / (ms * 1e-3) / sms / (clk_khz * 1e3));
        }
    }
    // tcgen05 correctness
    run_umma_probe<128, 128>();
    run_umma_probe<256, 64>();
    run_umma_probe<64, 256>();
    // tcgen05 rate
    {
        long long* cyc; CK(cudaMalloc(&cyc, sms * 8));
        const int smem = 128 * 64 * 2 + 256 * 64 * 2;
        CK(cudaFuncSetAttribute(umma_rate_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        for (int warm = 0; warm < 2; ++warm) {
            const int iters = 8192;
            CK(cudaEventRecord(e0));
            umma_rate_probe<<<sms, 128, smem>>>(iters, cyc);
            CK(cudaEventRecord(e1));
            cudaError_t e = cudaEventSynchronize(e1);
            if (e != cudaSuccess) { printf("umma_rate_probe: CUDA error %s\n", cudaGetErrorString(e)); return 3; }
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (warm) {
                long long c0; CK(cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost));
                double flop = (double)sms * iters * 2.0 * 128 * 256 * 16;
                printf("tcgen05.mma kind::f16 128x256x16 cta_group::1: %.1f TFLOP/s; CTA0 %.1f clk per MMA (%.0f flop/clk/SM)\n",
                       flop / ms / 1e9, (double)c0 / iters, 2.0 * 128 * 256 * 16 / ((double)c0 / iters));
            }
        }
    }
    printf("probes done\n");
    return 0;
}
