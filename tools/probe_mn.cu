// Probe for the attention kernel's tcgen05 operand forms on B200:
//   variant 0/1: D[128][32] = P[128][K] . V[K][32]  with B = V in MN-major (N = d contiguous) no-swizzle layout
//                [key/8][d/8][key%8][d%8]; variant 1 swaps the LBO/SBO assignment.
//   variant 2:   S[128][160] = Q[128][32] . Kc[160][32]^T with B a 160-key chunk inside a 480-key K-major image.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/probe_mn tools/probe_mn.cu ; tools/bin/probe_mn <variant>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(2); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool wait_bar(uint32_t bar) {
    uint32_t done = 0; long long t0 = clock64();
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(0) : "memory");
        if (!done && clock64() - t0 > 2000000000LL) return false;
    }
    return true;
}

// A: [128][K] fp16 row-major -> K-major canonical image; B per variant; out C [128][N] fp32
template <int N, int K>
__global__ void __launch_bounds__(128) probe(const __half* A, const __half* B, float* C, int variant, int* status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_s;
    __shared__ __align__(8) uint64_t bar_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint8_t* sa = smem;                      // A image: [k/8][row/8 (16)][row%8][8]  (chunk stride 2048)
    uint8_t* sb = smem + 128 * K * 2;
    for (int i = tid; i < 128 * (K / 8); i += 128) {
        const int r = i / (K / 8), c = i % (K / 8);
        *reinterpret_cast<uint4*>(sa + c * 2048 + r * 16) = *reinterpret_cast<const uint4*>(A + r * K + c * 8);
    }
    if (variant < 2) {
        // B = V [K keys][32 d] row-major -> MN-major image [key/8][d/8][key%8][d%8]
        for (int i = tid; i < K * 4; i += 128) {
            const int key = i / 4, dc = i % 4;
            *reinterpret_cast<uint4*>(sb + (key / 8) * 512 + dc * 128 + (key % 8) * 16) = *reinterpret_cast<const uint4*>(B + key * 32 + dc * 8);
        }
    } else {
        // B = Kmat [480 keys][32 d] row-major -> K-major image [d/8 (4)][key/8 (60)][key%8][8]
        for (int i = tid; i < 480 * 4; i += 128) {
            const int key = i / 4, dc = i % 4;
            *reinterpret_cast<uint4*>(sb + dc * 7680 + key * 16) = *reinterpret_cast<const uint4*>(B + key * 32 + dc * 8);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    const uint32_t bar = smem_u32(&bar_s);
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(bar)); asm volatile("fence.mbarrier_init.release.cluster;\n"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(&tmem_s)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_s;
    if (tid == 0) {
        uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        if (variant < 2) idesc |= 1u << 16;                                   // B is MN-major
        for (int k = 0; k < K / 16; ++k) {
            const uint64_t ad = make_desc(smem_u32(sa) + k * 2 * 2048, 2048, 128);
            uint64_t bd;
            if (variant == 0) bd = make_desc(smem_u32(sb) + k * 2 * 512, 512 /*LBO = K-group stride*/, 128 /*SBO = MN-group stride*/);
            else if (variant == 1) bd = make_desc(smem_u32(sb) + k * 2 * 512, 128, 512);
            else bd = make_desc(smem_u32(sb) + 160 * 16 /*chunk 1: keys 160..319*/ + k * 2 * 7680, 7680, 128);
            umma(tmem, ad, bd, idesc, k > 0);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(bar) : "memory");
    }
    const bool ok = wait_bar(bar);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (ok) {
        for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t v[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            for (int j = 0; j < 16; ++j) C[tid * N + c0 + j] = __uint_as_float(v[j]);
        }
    }
    if (tid == 0) *status = ok ? 1 : -1;
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "r"(256) : "memory");
}

template <int N, int K>
int run(int variant) {
    const int brows = variant < 2 ? K : 480;
    std::vector<__half> ha(128 * K), hb(brows * 32);
    std::vector<float> fa(128 * K), fb(brows * 32), ref(128 * N), out(128 * N);
    srand(7);
    for (size_t i = 0; i < ha.size(); ++i) { ha[i] = __float2half((rand() % 2001 - 1000) / 1000.f); fa[i] = __half2float(ha[i]); }
    for (size_t i = 0; i < hb.size(); ++i) { hb[i] = __float2half((rand() % 2001 - 1000) / 1000.f); fb[i] = __half2float(hb[i]); }
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            if (variant < 2) for (int k = 0; k < K; ++k) s += (double)fa[m * K + k] * fb[k * 32 + n];          // P . V
            else for (int k = 0; k < K; ++k) s += (double)fa[m * K + k] * fb[(160 + n) * 32 + k];              // Q . K_chunk1^T
            ref[m * N + n] = (float)s;
        }
    __half *dA, *dB; float* dC; int* dS;
    CK(cudaMalloc(&dA, ha.size() * 2)); CK(cudaMalloc(&dB, hb.size() * 2)); CK(cudaMalloc(&dC, out.size() * 4)); CK(cudaMalloc(&dS, 4));
    CK(cudaMemcpy(dA, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0, out.size() * 4)); CK(cudaMemset(dS, 0, 4));
    const int smem = 128 * K * 2 + brows * 64;
    CK(cudaFuncSetAttribute(probe<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe<N, K><<<1, 128, smem>>>(dA, dB, dC, variant, dS);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 2; }
    int st = 0;
    CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out.data(), dC, out.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (size_t i = 0; i < out.size(); ++i) maxerr = fmax(maxerr, fabs((double)out[i] - ref[i]));
    printf("variant %d (M=128 N=%d K=%d): status=%d max_abs_err=%.3e %s\n", variant, N, K, st, maxerr, maxerr < 2e-3 ? "MATCH" : "mismatch");
    return 0;
}

int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    if (variant < 2) return run<32, 64>(variant);
    return run<160, 32>(variant);
}
