// Training-step kernels of the T2S-DiT (reference: train.py:66-87 around model/denoiser/transformer.py:158-193).
//
// The training path is modular: activations live in plain row-major fp32 [tokens][features] buffers
// (token row = seq * 480 + n), every Linear (forward, input-gradient and weight-gradient form) goes through one
// generic tcgen05 GEMM (kind::tf32, fp32 operands straight from the parameter / activation buffers, fp32
// accumulation in TMEM), the token-local maths (LayerNorm + modulate, gates, GELU and their backward forms, the
// per-sequence adaLN reductions, bias column sums, loss) runs in warp-per-row kernels, and attention (forward and
// backward) is fused on tcgen05 without materialised scores (train_attn.cuh).
#pragma once
#include "common.cuh"

namespace t2s {

// =================================================================================== generic GEMM (tcgen05, tf32)
// C[b][m][n] (=, +=, atomic +=)  alpha * sum_k A_b(m,k) * B_b(n,k)  (+ bias[n])
//   operand "K-major":  element (m,k) at base + m*ld + k     (nn.Linear weights, activations as the left operand)
//   operand "MN-major": element (m,k) at base + k*ld + m     (transposed views: X^T, W^T), fed to the tensor core
//                       through the MN-major shared-memory descriptor form, no data transposition
//   batch b -> offset (b / bdiv) * s_hi + (b % bdiv) * s_lo  per operand ((sequence, head) views of [T][384] buffers)
//   split-K: blockIdx.z = b * ksplit + ks, partial sums are accumulated with atomics (mode GEMM_ATOMIC)
// Requirements: K-major operands: ld % 4 == 0, K % 4 == 0;  MN-major operands: ld % 4 == 0, rows % 4 == 0;
// bases 16-byte aligned.  Rows / K beyond the matrix are zero-filled (cp.async src-size).
// CTA = 160 threads: warps 0-3 load (cp.async 16 B into the canonical layouts: no-swizzle core matrices for K-major
// operands, the 128B/32B-atom swizzle for MN-major ones; 3 stages) and
// then run the epilogue (thread = accumulator row = TMEM lane), warp 4 issues tcgen05.mma (M128 x bn x 8).
enum { GEMM_STORE = 0, GEMM_ADD = 1, GEMM_ATOMIC = 2 };
struct GemmArgs {
    const float* A; const float* B; float* C; const float* bias;
    int M, N, K;
    int lda, ldb, ldc;
    long long sa_hi, sa_lo, sb_hi, sb_lo, sc_hi, sc_lo;
    int bdiv;
    int a_mn, b_mn;     // 1 = MN-major operand
    int bn;             // N tile: 32, 64, 96 or 128
    int ksplit;
    int mode;
    float alpha;
    // persistent form only: when set, C is not written; the tile (+ bias) goes out as the fp16 T8 operand images of the
    // fused attention (train_attn.cuh): N = 384 = q | k | v, each 4 heads x 32 features; rows = sequence * 480 + token
    __half* qkv_img;
    int img_ntok;       // tokens per sequence of those images (480 | 800 | 1024)
};
constexpr int G_BM = 128, G_BK = 32, G_STAGES = 3, G_THREADS = 160;
constexpr int G_STAGE_BYTES = 2 * G_BM * G_BK * 4;                     // A tile + B tile (B sized for bn = 128)
constexpr int G_SM_BAR = G_STAGES * G_STAGE_BYTES;
constexpr int G_SMEM_BYTES = G_SM_BAR + 16 * 8;

__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// MN-major 32-bit operands exist only in the "128-byte swizzle with 32-byte atomicity" layout (layout type 1):
// per block of 32 MN elements, rows = k (128 B each, 32-byte chunks XOR-ed with k % 4), 4-row atoms of 512 B.
//   element (mn, k) at  (mn/32)*4096 + k*128 + ((((mn%32)/8) ^ (k%4)) * 32) + (mn%8)*4      [32 k rows per stage]
// LBO = stride between 32-element MN blocks (4096 B), SBO = stride between 4-row k groups (512 B).
__device__ __forceinline__ uint64_t umma_desc_mn32(uint32_t saddr) {
    return umma_desc(saddr, 4096, 512) | ((uint64_t)1 << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major fp32 operand image of one 32-wide K step: the 128-byte-swizzle layout (descriptor layout type 2): row r is 128
// contiguous bytes (32 floats) at r * 128, its 16-byte chunk j stored at position j ^ (r % 8); 8-row groups 1024 B apart
// (SBO).  16-byte chunk c of the tile: 8 consecutive lanes copy one row's 128 contiguous global bytes (full lines) into
// 128 contiguous (permuted) bytes of shared memory, a warp instruction 4 rows = 512 contiguous bytes: conflict-free.
// [The first version used the no-swizzle core-matrix image [k/4][row][4]: the 8 lanes of a row wrote 2048 B apart, an
// 8-way bank conflict per cp.async.]  A K = 8 MMA step advances the descriptor start address by 32 B inside the atom.
__device__ __forceinline__ uint32_t kmajor_chunk(int c, int& row, int& kc) {
    kc = c & 7;
    row = c >> 3;
    return (uint32_t)(row * 128 + ((kc ^ (row & 7)) << 4));
}
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t saddr) {
    return umma_desc(saddr, 16, 1024) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(dst), "l"(src), "r"(bytes) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }
__device__ __forceinline__ float round_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__global__ void __launch_bounds__(G_THREADS, 2) gemm_tf32_kernel(const GemmArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int mt = blockIdx.x, nt = blockIdx.y, batch = blockIdx.z / p.ksplit, ks = blockIdx.z % p.ksplit;
    const int ksteps = (p.K + G_BK - 1) / G_BK, per = (ksteps + p.ksplit - 1) / p.ksplit;
    const int k_begin = ks * per, nk = min(ksteps, k_begin + per) - k_begin;
    if (nk <= 0) return;                                                   // uniform across the CTA
    const int bn = p.bn;
    const uint32_t sb = smem_u32(smem), bar0 = sb + G_SM_BAR;
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (G_STAGES + s); };
    const uint32_t ACC = bar0 + 8u * (2 * G_STAGES);
    if (tid == 0) {
        for (int s = 0; s < G_STAGES; ++s) { mbar_init(FULL(s), 128); mbar_init(EMPTY(s), 1); }
        mbar_init(ACC, 1);
        mbar_fence_init();
    }
    const uint32_t tcols = bn <= 32 ? 32u : (bn <= 64 ? 64u : 128u);
    if (warp == 4) tmem_alloc(smem_u32(&tmem_slot), tcols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);

    const float* A = p.A + (long long)(batch / p.bdiv) * p.sa_hi + (long long)(batch % p.bdiv) * p.sa_lo;
    const float* B = p.B + (long long)(batch / p.bdiv) * p.sb_hi + (long long)(batch % p.bdiv) * p.sb_lo;
    float* C = p.C + (long long)(batch / p.bdiv) * p.sc_hi + (long long)(batch % p.bdiv) * p.sc_lo;
    const int m0 = mt * G_BM, n0 = nt * bn;

    if (warp < 4) {
        // ------------------------------------------------------------------ loaders
        auto load_stage = [&](int it) {
            const int s = it % G_STAGES, k0 = (k_begin + it) * G_BK;
            const uint32_t sa = sb + s * G_STAGE_BYTES, sbb = sa + G_BM * G_BK * 4;
            if (!p.a_mn) {                                   // K-major image (kmajor_chunk): conflict-free 512-byte warp writes
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    int row, kc;
                    const uint32_t off = kmajor_chunk(j * 128 + tid, row, kc);
                    const int gm = m0 + row, gk = k0 + kc * 4;
                    const bool ok = gm < p.M && gk < p.K;
                    cp_async16_zfill(sa + off, ok ? A + (long long)gm * p.lda + gk : A, ok ? 16u : 0u);
                }
            } else {                                         // MN-major image (see gemm header): lanes = 8 chunks of one k row
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = j * 128 + tid, ql = c & 7, k = (c >> 3) & 31, blk = c >> 8;
                    const int gm = m0 + (blk * 8 + ql) * 4, gk = k0 + k;
                    const bool ok = gm < p.M && gk < p.K;
                    cp_async16_zfill(sa + blk * 4096 + k * 128 + ((((ql >> 1) ^ k) & 3) << 5) + (ql & 1) * 16,
                                     ok ? A + (long long)gk * p.lda + gm : A, ok ? (uint32_t)min(16, (p.M - gm) * 4) : 0u);
                }
            }
            const int nchunk = bn * 8;
            if (!p.b_mn) {
                for (int c = tid; c < nchunk; c += 128) {
                    int row, kc;
                    const uint32_t off = kmajor_chunk(c, row, kc);
                    const int gn = n0 + row, gk = k0 + kc * 4;
                    const bool ok = gn < p.N && gk < p.K;
                    cp_async16_zfill(sbb + off, ok ? B + (long long)gn * p.ldb + gk : B, ok ? 16u : 0u);
                }
            } else {
                for (int c = tid; c < nchunk; c += 128) {
                    const int ql = c & 7, k = (c >> 3) & 31, blk = c >> 8;
                    const int gn = n0 + (blk * 8 + ql) * 4, gk = k0 + k;
                    const bool ok = gn < p.N && gk < p.K;
                    cp_async16_zfill(sbb + blk * 4096 + k * 128 + ((((ql >> 1) ^ k) & 3) << 5) + (ql & 1) * 16,
                                     ok ? B + (long long)gk * p.ldb + gn : B, ok ? (uint32_t)min(16, (p.N - gn) * 4) : 0u);
                }
            }
            cp_async_commit();
        };
        auto publish = [&](int it) {                         // this thread's copies of stage `it` have landed
            fence_async_smem();
            mbar_arrive(FULL(it % G_STAGES));
        };
#pragma unroll 1
        for (int it = 0; it < nk; ++it) {
            const int use = it / G_STAGES;
            if (use > 0) mbar_wait(EMPTY(it % G_STAGES), (use - 1) & 1);
            load_stage(it);
            if (it >= 2) { cp_async_wait<2>(); publish(it - 2); }
        }
        if (nk >= 2) { cp_async_wait<1>(); publish(nk - 2); }
        cp_async_wait<0>();
        publish(nk - 1);

        // ------------------------------------------------------------------ epilogue
        // thread = accumulator row (TMEM lane): rows are staged in this warp's slice of the (now idle) operand
        // stages and written out row by row, a lane owning 4 consecutive columns (coalesced 16-byte accesses)
        mbar_wait(ACC, 0);
        tc_fence_after();
        const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
        const int pitch = bn + 4;                                          // floats; +4 keeps the float4 stores conflict-free
        float* stg = reinterpret_cast<float*>(smem) + warp * 32 * (G_BM + 4);
#pragma unroll 1
        for (int cb = 0; cb < bn / 16; ++cb) {
            float v[16];
            tmem_ld16(trow + cb * 16, v);
            tmem_wait_ld();
#pragma unroll
            for (int q = 0; q < 4; ++q) st4(stg + lane * pitch + cb * 16 + q * 4, make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]));
        }
        __syncwarp();
        const int lpr = bn >> 2;                                           // lanes per row
        const int rpi = 32 / lpr > 0 ? 32 / lpr : 1;                       // rows per warp instruction
        const int cl = (lane % lpr) * 4, rsub = lane / lpr;
        const int gn = n0 + cl;
        const bool lane_ok = lane < lpr * rpi && gn < p.N;
        const bool vec_ok = (p.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && gn + 3 < p.N;
        float bv[4] = {0.f, 0.f, 0.f, 0.f};
        if (p.bias != nullptr && lane_ok) {
#pragma unroll
            for (int u = 0; u < 4; ++u) if (gn + u < p.N) bv[u] = p.bias[gn + u];
        }
#pragma unroll 1
        for (int r0 = 0; r0 < 32; r0 += rpi) {
            const int r = r0 + rsub, gm = m0 + warp * 32 + r;
            if (!lane_ok || gm >= p.M) continue;
            const float4 a = ld4(stg + r * pitch + cl);
            float rr[4] = {fmaf(a.x, p.alpha, bv[0]), fmaf(a.y, p.alpha, bv[1]), fmaf(a.z, p.alpha, bv[2]), fmaf(a.w, p.alpha, bv[3])};
            float* dst = C + (long long)gm * p.ldc + gn;
            if (p.mode == GEMM_ATOMIC) {
#pragma unroll
                for (int u = 0; u < 4; ++u) if (gn + u < p.N) atomicAdd(dst + u, rr[u]);
            } else if (vec_ok) {
                if (p.mode == GEMM_ADD) { const float4 o = ld4(dst); rr[0] += o.x; rr[1] += o.y; rr[2] += o.z; rr[3] += o.w; }
                st4(dst, make_float4(rr[0], rr[1], rr[2], rr[3]));
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) if (gn + u < p.N) dst[u] = (p.mode == GEMM_ADD ? dst[u] : 0.f) + rr[u];
            }
        }
    } else {
        // ------------------------------------------------------------------ MMA issuer (whole warp converged; lane 0 issues)
        const bool lead = lane == 0;
        const uint32_t idesc = umma_idesc_tf32(G_BM, bn, p.a_mn, p.b_mn);
#pragma unroll 1
        for (int it = 0; it < nk; ++it) {
            const int s = it % G_STAGES;
            mbar_wait(FULL(s), (it / G_STAGES) & 1);
            tc_fence_after();
            const uint32_t sa = sb + s * G_STAGE_BYTES, sbb = sa + G_BM * G_BK * 4;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {                 // K = 8 per instruction
                const uint64_t ad = p.a_mn ? umma_desc_mn32(sa + kk * 1024) : umma_desc_k128(sa + kk * 32);
                const uint64_t bd = p.b_mn ? umma_desc_mn32(sbb + kk * 1024) : umma_desc_k128(sbb + kk * 32);
                if (lead) umma_tf32(tmem, ad, bd, idesc, (it > 0 || kk > 0) ? 1u : 0u);
            }
            if (lead) umma_commit(EMPTY(s));
            __syncwarp();
        }
        if (lead) umma_commit(ACC);
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem, tcols);
}

// ----------------------------------------------------------------------------------- persistent form
// The same GEMM for the shapes the forward and the input-gradient passes use (A K-major, one batch, no split-K,
// N a multiple of 128): short K (128..384) makes a one-tile-per-CTA kernel latency-bound (prologue, a 4-step main
// loop and a 64 KB epilogue in series: ncu shows 15 % warps active, 20 % DRAM throughput), so here ONE CTA per SM
// walks over the output tiles with three decoupled roles:
//   warps 0-3  loaders: cp.async into a 4-stage ring that runs ahead across tile boundaries
//   warp  4    MMA issuer: accumulators double-buffered in TMEM (2 x 128 columns)
//   warps 5-8  epilogue: TMEM -> padded shared staging -> coalesced row stores (+ bias), overlapping the next tile's
//              loads and MMAs
constexpr int P_STAGES = 4, P_THREADS = 288;
constexpr int P_SM_STG = P_STAGES * G_STAGE_BYTES;                         // epilogue staging [4 warps][32 rows][132]
constexpr int P_SM_BAR = P_SM_STG + 4 * 32 * (G_BM + 4) * 4;
constexpr int P_SMEM_BYTES = P_SM_BAR + 16 * 8;
static_assert(P_SMEM_BYTES <= 232448, "persistent GEMM shared memory");

__global__ void __launch_bounds__(P_THREADS, 1) gemm_tf32_persistent_kernel(const GemmArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int ksteps = (p.K + G_BK - 1) / G_BK;
    const int mtiles = (p.M + G_BM - 1) / G_BM, ntiles = p.N / 128, items = mtiles * ntiles;
    const uint32_t sb = smem_u32(smem), bar0 = sb + P_SM_BAR;
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (P_STAGES + s); };
    auto ACCFULL = [&](int b) { return bar0 + 8u * (2 * P_STAGES + b); };
    auto ACCEMPTY = [&](int b) { return bar0 + 8u * (2 * P_STAGES + 2 + b); };
    if (tid == 0) {
        for (int s = 0; s < P_STAGES; ++s) { mbar_init(FULL(s), 128); mbar_init(EMPTY(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(ACCFULL(b), 1); mbar_init(ACCEMPTY(b), 128); }
        mbar_fence_init();
    }
    if (warp == 4) tmem_alloc(smem_u32(&tmem_slot), 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);
    const float* A = p.A;
    const float* B = p.B;
    // profiling knobs of tools/gemm_bench.py (never set by the library's own launches): skip the global stores / the loads
    const bool dbg_nostore = (p.mode & 16) != 0, dbg_noload = (p.mode & 32) != 0;
    const int mode = p.mode & 15;

    if (warp < 4) {
        // ------------------------------------------------------------------ loaders
        int g = 0;                                           // global stage counter (ring position)
        auto load_stage = [&](int m0, int n0, int k0) {
            const int s = g % P_STAGES;
            const uint32_t sa = sb + s * G_STAGE_BYTES, sbb = sa + G_BM * G_BK * 4;
#pragma unroll
            for (int j = 0; j < 8; ++j) {                    // A K-major image (kmajor_chunk)
                int row, kc;
                const uint32_t off = kmajor_chunk(j * 128 + tid, row, kc);
                const int gm = m0 + row, gk = k0 + kc * 4;
                const bool ok = gm < p.M && gk < p.K;
                cp_async16_zfill(sa + off, ok ? A + (long long)gm * p.lda + gk : A, ok ? 16u : 0u);
            }
            if (!p.b_mn) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    int row, kc;
                    const uint32_t off = kmajor_chunk(j * 128 + tid, row, kc);
                    const int gn = n0 + row, gk = k0 + kc * 4;
                    const bool ok = gk < p.K;
                    cp_async16_zfill(sbb + off, ok ? B + (long long)gn * p.ldb + gk : B, ok ? 16u : 0u);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {                // MN-major image (see the header comment of this file)
                    const int c = j * 128 + tid, ql = c & 7, k = (c >> 3) & 31, blk = c >> 8;
                    const int gn = n0 + (blk * 8 + ql) * 4, gk = k0 + k;
                    const bool ok = gk < p.K;
                    cp_async16_zfill(sbb + blk * 4096 + k * 128 + ((((ql >> 1) ^ k) & 3) << 5) + (ql & 1) * 16,
                                     ok ? B + (long long)gk * p.ldb + gn : B, ok ? 16u : 0u);
                }
            }
            cp_async_commit();
        };
        auto publish = [&](int gg) {                          // this thread's copies of ring position gg have landed
            fence_async_smem();
            mbar_arrive(FULL(gg % P_STAGES));
        };
#pragma unroll 1
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int m0 = (item / ntiles) * G_BM, n0 = (item % ntiles) * 128;
#pragma unroll 1
            for (int ks = 0; ks < ksteps; ++ks, ++g) {
                const int use = g / P_STAGES;
                if (use > 0) mbar_wait(EMPTY(g % P_STAGES), (use - 1) & 1);
                if (!dbg_noload) load_stage(m0, n0, ks * G_BK); else cp_async_commit();
                if (g >= 2) { cp_async_wait<2>(); publish(g - 2); }
            }
        }
        if (g >= 2) { cp_async_wait<1>(); publish(g - 2); }
        cp_async_wait<0>();
        if (g >= 1) publish(g - 1);
    } else if (warp == 4) {
        // ------------------------------------------------------------------ MMA issuer (whole warp converged; lane 0 issues)
        const bool lead = lane == 0;
        const uint32_t idesc = umma_idesc_tf32(G_BM, 128, 0, p.b_mn);
        int g = 0, it = 0;
#pragma unroll 1
        for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
            const int ab = it & 1;
            if (it >= 2) {                                       // the epilogue has drained this accumulator
                mbar_wait(ACCEMPTY(ab), ((it >> 1) - 1) & 1);
                tc_fence_after();
            }
#pragma unroll 1
            for (int ks = 0; ks < ksteps; ++ks, ++g) {
                const int s = g % P_STAGES;
                mbar_wait(FULL(s), (g / P_STAGES) & 1);
                tc_fence_after();
                const uint32_t sa = sb + s * G_STAGE_BYTES, sbb = sa + G_BM * G_BK * 4;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {                 // K = 8 per instruction
                    const uint64_t ad = umma_desc_k128(sa + kk * 32);
                    const uint64_t bd = p.b_mn ? umma_desc_mn32(sbb + kk * 1024) : umma_desc_k128(sbb + kk * 32);
                    if (lead) umma_tf32(tmem + ab * 128, ad, bd, idesc, (ks > 0 || kk > 0) ? 1u : 0u);
                }
                if (lead) umma_commit(EMPTY(s));
                __syncwarp();
            }
            if (lead) umma_commit(ACCFULL(ab));
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ epilogue: thread = accumulator row (TMEM lane)
        const int wq = warp & 3;                                 // TMEM lane group of this warp
        const uint32_t trow = tmem + ((uint32_t)(wq * 32) << 16);
        constexpr int pitch = G_BM + 4;
        float* stg = reinterpret_cast<float*>(smem + P_SM_STG) + wq * 32 * pitch;
        const int cl = lane * 4;
        int it = 0;
#pragma unroll 1
        for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
            const int ab = it & 1;
            const int m0 = (item / ntiles) * G_BM, n0 = (item % ntiles) * 128;
            mbar_wait(ACCFULL(ab), (it >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int cb = 0; cb < 8; ++cb) {
                float v[16];
                tmem_ld16(trow + ab * 128 + cb * 16, v);
                tmem_wait_ld();
#pragma unroll
                for (int q = 0; q < 4; ++q) st4(stg + lane * pitch + cb * 16 + q * 4, make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]));
            }
            tc_fence_before();
            mbar_arrive(ACCEMPTY(ab));                           // the MMA warp may overwrite this accumulator
            __syncwarp();
            const int gn = n0 + cl;
            if (p.qkv_img != nullptr) {
                // thread = row: [sequence][head][q | k | v][token / 8][feature / 8][token % 8][8] fp16; for one (head, chunk) a
                // warp writes 4 x 128 contiguous bytes
                const int gm = m0 + wq * 32 + lane;
                if (gm < p.M) {
                    const int ntok = p.img_ntok, sq = gm / ntok, tok = gm - sq * ntok, which = n0 >> 7;
                    __half* base = p.qkv_img + ((size_t)sq * NHEAD * 3 + which) * ((size_t)ntok * HD) + (tok >> 3) * 256 + (tok & 7) * 8;
                    const float* srow = stg + lane * pitch;
#pragma unroll 4
                    for (int hc = 0; hc < 16; ++hc) {            // head hc / 4, feature chunk hc % 4
                        const float4 a = ld4(srow + hc * 8), b = ld4(srow + hc * 8 + 4);
                        float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
                        if (p.bias != nullptr) { b0 = ld4(p.bias + n0 + hc * 8); b1 = ld4(p.bias + n0 + hc * 8 + 4); }
                        *reinterpret_cast<uint4*>(base + (size_t)(hc >> 2) * 3 * ((size_t)ntok * HD) + (hc & 3) * 64) =
                            make_uint4(pack_h2(fmaf(a.x, p.alpha, b0.x), fmaf(a.y, p.alpha, b0.y)), pack_h2(fmaf(a.z, p.alpha, b0.z), fmaf(a.w, p.alpha, b0.w)),
                                       pack_h2(fmaf(b.x, p.alpha, b1.x), fmaf(b.y, p.alpha, b1.y)), pack_h2(fmaf(b.z, p.alpha, b1.z), fmaf(b.w, p.alpha, b1.w)));
                    }
                }
                __syncwarp();
                continue;
            }
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias != nullptr) bv = ld4(p.bias + gn);
#pragma unroll 4
            for (int r = 0; r < 32; ++r) {
                const int gm = m0 + wq * 32 + r;
                if (gm >= p.M) break;
                const float4 a = ld4(stg + r * pitch + cl);
                float4 rr = make_float4(fmaf(a.x, p.alpha, bv.x), fmaf(a.y, p.alpha, bv.y), fmaf(a.z, p.alpha, bv.z), fmaf(a.w, p.alpha, bv.w));
                float* dst = p.C + (long long)gm * p.ldc + gn;
                if (mode == GEMM_ADD) { const float4 o = ld4(dst); rr.x += o.x; rr.y += o.y; rr.z += o.z; rr.w += o.w; }
                if (!dbg_nostore) st4(dst, rr);
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem, 256);
}

// =================================================================================== warp helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float4 round4(float4 v) { return make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w)); }

// Row kernels over [T][128] buffers use: grid (nseq, tokens / rc), block 256 = 8 warps; a CTA owns rc tokens of one sequence
// (rc = 60 | 50 | 64 for the 480 | 800 | 1024-token shapes; the token count is gridDim.y * rc),
// a warp walks rows, a lane owns 4 consecutive features (float4).  Per-sequence / per-feature sums are reduced
// across the CTA's warps in shared memory and added to the global accumulators with one atomic per feature.
constexpr int ROW_THREADS = 256;
__device__ __forceinline__ void cta_feature_atomic(float4 acc, float* __restrict__ dst, float (*sm)[D]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    st4(&sm[warp][lane * 4], acc);
    __syncthreads();
    if (threadIdx.x < D) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < ROW_THREADS / 32; ++w) s += sm[w][threadIdx.x];
        atomicAdd(dst + threadIdx.x, s);
    }
    __syncthreads();
}

// ---- time embedding + text conditioning + SiLU: sc[seq][128] = SiLU(temb(t) (+ emb))   transformer.py:30-40,106,176-178
__global__ void cond_act_kernel(float* __restrict__ sc, const float* __restrict__ t100, const float* __restrict__ emb,
                                const float* __restrict__ freqs, int nseq) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nseq * D) return;
    const int seq = i >> 7, f = i & 127;
    const float arg = __fdiv_rn(t100[seq], freqs[f & 63]);
    float c = (f < 64) ? sinf(arg) : cosf(arg);
    if (emb != nullptr) c = c + emb[i];
    sc[i] = c / (1.0f + expf(-c));
}

// ---- patchify + conv + patch_emb + pos_embed (transformer.py:166-172): h0 [T][128], xp [T][4] (patch pixels)
__global__ void __launch_bounds__(ROW_THREADS) embed_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w_embed,
                                                               const float* __restrict__ b_embed, const float* __restrict__ pos,
                                                               float* __restrict__ h0, float* __restrict__ xp, int nseq, int rc) {
    [[maybe_unused]] const int ntok = gridDim.y * rc, latp = ntok >> 4, lat = LATC * latp;   // tokens per sequence, latent width

    const int seq = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* xs = x + (size_t)seq * lat;
    for (int tl = warp; tl < rc; tl += ROW_THREADS / 32) {
        const int n = blockIdx.y * rc + tl, i = n >> 5, j = n & 31;
        float xv[4];
#pragma unroll
        for (int pq = 0; pq < 4; ++pq) xv[pq] = xs[(2 * j + (pq & 1)) * latp + 2 * i + (pq >> 1)];
        const size_t row = (size_t)seq * ntok + n;
        if (lane == 0) st4(xp + row * 4, make_float4(xv[0], xv[1], xv[2], xv[3]));
        float r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int c = lane * 4 + u;
            const float4 w = ld4(w_embed + c * 4);
            r[u] = w.x * xv[0] + w.y * xv[1] + w.z * xv[2] + w.w * xv[3] + b_embed[c] + pos[n * D + c];
        }
        st4(h0 + row * D + lane * 4, make_float4(r[0], r[1], r[2], r[3]));
    }
}

// ---- a = LN(h; eps, no affine) * (1 + scale) + shift    (transformer.py:7-8,102-103,116-117)
__global__ void __launch_bounds__(ROW_THREADS) ln_mod_fwd_kernel(const float* __restrict__ h, const float* __restrict__ mod, int mod_stride,
                                                                int shift_off, float* __restrict__ a, float eps, int rc) {
    [[maybe_unused]] const int ntok = gridDim.y * rc, latp = ntok >> 4, lat = LATC * latp;   // tokens per sequence, latent width

    const int seq = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float4 sh = ld4(mod + (size_t)seq * mod_stride + shift_off + lane * 4);
    const float4 sc = ld4(mod + (size_t)seq * mod_stride + shift_off + D + lane * 4);
    for (int tl = warp; tl < rc; tl += ROW_THREADS / 32) {
        const size_t row = (size_t)seq * ntok + blockIdx.y * rc + tl;
        const float4 v = ld4(h + row * D + lane * 4);
        const float mean = warp_sum(v.x + v.y + v.z + v.w) * (1.f / D);
        const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
        const float rstd = rsqrtf(warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.f / D) + eps);
        st4(a + row * D + lane * 4, round4(make_float4(fmaf(dx * rstd, 1.f + sc.x, sh.x), fmaf(dy * rstd, 1.f + sc.y, sh.y),
                                                       fmaf(dz * rstd, 1.f + sc.z, sh.z), fmaf(dw * rstd, 1.f + sc.w, sh.w))));
    }
}

// ---- backward of the above: da -> dshift, dscale (per sequence), dh_out = dh_in + LN'(da * (1 + scale))
__global__ void __launch_bounds__(ROW_THREADS) ln_mod_bwd_kernel(const float* __restrict__ da, const float* __restrict__ h,
                                                                const float* __restrict__ mod, float* __restrict__ dmod, int mod_stride,
                                                                int shift_off, const float* __restrict__ dh_in, float* __restrict__ dh_out, float eps, int rc) {
    [[maybe_unused]] const int ntok = gridDim.y * rc, latp = ntok >> 4, lat = LATC * latp;   // tokens per sequence, latent width

    __shared__ float sm[ROW_THREADS / 32][D];
    const int seq = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float4 sc = ld4(mod + (size_t)seq * mod_stride + shift_off + D + lane * 4);
    float4 dsh = make_float4(0.f, 0.f, 0.f, 0.f), dsc = dsh;
    for (int tl = warp; tl < rc; tl += ROW_THREADS / 32) {
        const size_t row = (size_t)seq * ntok + blockIdx.y * rc + tl;
        const float4 v = ld4(h + row * D + lane * 4), g = ld4(da + row * D + lane * 4);
        const float mean = warp_sum(v.x + v.y + v.z + v.w) * (1.f / D);
        const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
        const float rstd = rsqrtf(warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.f / D) + eps);
        const float4 xh = make_float4(dx * rstd, dy * rstd, dz * rstd, dw * rstd);
        dsh.x += g.x; dsh.y += g.y; dsh.z += g.z; dsh.w += g.w;
        dsc.x = fmaf(g.x, xh.x, dsc.x); dsc.y = fmaf(g.y, xh.y, dsc.y); dsc.z = fmaf(g.z, xh.z, dsc.z); dsc.w = fmaf(g.w, xh.w, dsc.w);
        const float4 gx = make_float4(g.x * (1.f + sc.x), g.y * (1.f + sc.y), g.z * (1.f + sc.z), g.w * (1.f + sc.w));
        const float m1 = warp_sum(gx.x + gx.y + gx.z + gx.w) * (1.f / D);
        const float m2 = warp_sum(gx.x * xh.x + gx.y * xh.y + gx.z * xh.z + gx.w * xh.w) * (1.f / D);
        float4 o = make_float4(rstd * (gx.x - m1 - xh.x * m2), rstd * (gx.y - m1 - xh.y * m2),
                               rstd * (gx.z - m1 - xh.z * m2), rstd * (gx.w - m1 - xh.w * m2));
        if (dh_in != nullptr) { const float4 r = ld4(dh_in + row * D + lane * 4); o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }
        st4(dh_out + row * D + lane * 4, o);
    }
    float* dm = dmod + (size_t)seq * mod_stride + shift_off;
    cta_feature_atomic(dsh, dm, sm);
    cta_feature_atomic(dsc, dm + D, sm);
}

// ---- h_out = h_in + gate * y     (transformer.py:116-117)
__global__ void __launch_bounds__(ROW_THREADS) gate_res_fwd_kernel(const float* __restrict__ h_in, const float* __restrict__ y,
                                                                  const float* __restrict__ mod, int mod_stride, int gate_off,
                                                                  float* __restrict__ h_out, int rc) {
    [[maybe_unused]] const int ntok = gridDim.y * rc, latp = ntok >> 4, lat = LATC * latp;   // tokens per sequence, latent width

    const int seq = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float4 g = ld4(mod + (size_t)seq * mod_stride + gate_off + lane * 4);
    for (int tl = warp; tl < rc; tl += ROW_THREADS / 32) {
        const size_t o = ((size_t)seq * ntok + blockIdx.y * rc + tl) * D + lane * 4;
        const float4 a = ld4(h_in + o), b = ld4(y + o);
        st4(h_out + o, make_float4(fmaf(g.x, b.x, a.x), fmaf(g.y, b.y, a.y), fmaf(g.z, b.z, a.z), fmaf(g.w, b.w, a.w)));
    }
}
// backward: dgate[seq] += sum_n dh * y ; dy = dh * gate ; dbias += sum_rows dy
__global__ void __launch_bounds__(ROW_THREADS) gate_res_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ y,
                                                                  const float* __restrict__ mod, float* __restrict__ dmod, int mod_stride,
                                                                  int gate_off, float* __restrict__ dy, float* __restrict__ dbias, int rc) {
    [[maybe_unused]] const int ntok = gridDim.y * rc, latp = ntok >> 4, lat = LATC * latp;   // tokens per sequence, latent width

    __shared__ float sm[ROW_THREADS / 32][D];
    const int seq = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float4 g = ld4(mod + (size_t)seq * mod_stride + gate_off + lane * 4);
    float4 dg = make_float4(0.f, 0.f, 0.f, 0.f), db = dg;
    for (int tl = warp; tl < rc; tl += ROW_THREADS / 32) {
        const size_t o = ((size_t)seq * ntok + blockIdx.y * rc + tl) * D + lane * 4;
        const float4 a = ld4(dh + o), b = ld4(y + o);
        dg.x = fmaf(a.x, b.x, dg.x); dg.y = fmaf(a.y, b.y, dg.y); dg.z = fmaf(a.z, b.z, dg.z); dg.w = fmaf(a.w, b.w, dg.w);
        const float4 r = make_float4(a.x * g.x, a.y * g.y, a.z * g.z, a.w * g.w);
        db.x += r.x; db.y += r.y; db.z += r.z; db.w += r.w;
        st4(dy + o, round4(r));
    }
    cta_feature_atomic(dg, dmod + (size_t)seq * mod_stride + gate_off, sm);
    cta_feature_atomic(db, dbias, sm);
}

// ---- GELU(tanh) forward / backward over [T][256]  (timm Mlp act, transformer.py:99)
__global__ void gelu_fwd_kernel(const float* __restrict__ z, float* __restrict__ hid, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = ld4(z + i * 4);
        auto f = [](float x) { const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x); return 0.5f * x * (1.f + tanhf(u)); };
        st4(hid + i * 4, round4(make_float4(f(v.x), f(v.y), f(v.z), f(v.w))));
    }
}
// dz = dhid * gelu'(z) in place over dhid
__global__ void gelu_bwd_kernel(const float* __restrict__ z, float* __restrict__ dhid, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = ld4(z + i * 4), g = ld4(dhid + i * 4);
        auto f = [](float x, float gy) {
            const float k0 = 0.7978845608028654f, k1 = 0.044715f;
            const float t = tanhf(k0 * (x + k1 * x * x * x));
            const float d = 0.5f * (1.f + t) + 0.5f * x * (1.f - t * t) * k0 * (1.f + 3.f * k1 * x * x);
            return gy * d;
        };
        st4(dhid + i * 4, round4(make_float4(f(v.x, g.x), f(v.y, g.y), f(v.z, g.z), f(v.w, g.w))));
    }
}

// ---- out[c] += sum_rows x[row*ld + c], c < cols   (bias gradients); grid (ceil(cols/128), nblocks), block 256
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, size_t rows, int ld, int cols, float* __restrict__ out) {
    __shared__ float sm[8][D];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * D + lane * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < cols)
        for (size_t r = (size_t)blockIdx.y * 8 + warp; r < rows; r += (size_t)gridDim.y * 8) {
            const float4 v = ld4(x + r * ld + c);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    st4(&sm[warp][lane * 4], acc);
    __syncthreads();
    if (threadIdx.x < D && blockIdx.x * D + threadIdx.x < cols) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += sm[w][threadIdx.x];
        atomicAdd(out + blockIdx.x * D + threadIdx.x, s);
    }
}

// ---- final LN (affine, eps 1e-5) + Linear(128 -> 4) + unpatchify (transformer.py:182-190), fused with the MSE loss
// (rectified_flow.py:13-16 / DDPM.py:37-38) and its backward down to dh4.
//   pred != nullptr: the prediction is written.  target != nullptr: loss_sum += sum (pred - target)^2.
//   dh != nullptr: backward, driven by dpred [nseq][64][30] when given, else by the fused MSE gradient
//   (pred - target) * dscale with dscale = 2 / numel; produces dh [T][128] and += dlnw / dlnb [128], dwf [4][128], dbf [4].
__global__ void __launch_bounds__(ROW_THREADS) final_kernel(const float* __restrict__ h, const float* __restrict__ lnw, const float* __restrict__ lnb,
                                                           const float* __restrict__ wf, const float* __restrict__ bf, float* __restrict__ pred,
                                                           const float* __restrict__ target, const float* __restrict__ dpred, float dscale,
                                                           float* __restrict__ loss_sum,
                                                           float* __restrict__ dh, float* __restrict__ dlnw, float* __restrict__ dlnb,
                                                           float* __restrict__ dwf, float* __restrict__ dbf, int rc) {
    [[maybe_unused]] const int ntok = gridDim.y * rc, latp = ntok >> 4, lat = LATC * latp;   // tokens per sequence, latent width

    __shared__ float sm[ROW_THREADS / 32][D];
    const int seq = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float4 w4 = ld4(lnw + lane * 4), b4 = ld4(lnb + lane * 4);
    float4 wfr[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) wfr[c] = ld4(wf + c * D + lane * 4);
    float4 a_lnw = make_float4(0.f, 0.f, 0.f, 0.f), a_lnb = a_lnw, a_wf[4] = {a_lnw, a_lnw, a_lnw, a_lnw};
    float a_bf[4] = {0.f, 0.f, 0.f, 0.f}, a_loss = 0.f;
    for (int tl = warp; tl < rc; tl += ROW_THREADS / 32) {
        const int n = blockIdx.y * rc + tl, i = n >> 5, j = n & 31;
        const size_t row = (size_t)seq * ntok + n;
        const float4 v = ld4(h + row * D + lane * 4);
        const float mean = warp_sum(v.x + v.y + v.z + v.w) * (1.f / D);
        const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
        const float rstd = rsqrtf(warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.f / D) + 1e-5f);
        const float4 xh = make_float4(dx * rstd, dy * rstd, dz * rstd, dw * rstd);
        const float4 y = make_float4(fmaf(xh.x, w4.x, b4.x), fmaf(xh.y, w4.y, b4.y), fmaf(xh.z, w4.z, b4.z), fmaf(xh.w, w4.w, b4.w));
        float o4[4], d4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            o4[c] = warp_sum(y.x * wfr[c].x + y.y * wfr[c].y + y.z * wfr[c].z + y.w * wfr[c].w) + bf[c];
            const size_t xi = (size_t)seq * lat + (2 * j + (c & 1)) * latp + 2 * i + (c >> 1);
            if (lane == 0 && pred != nullptr) pred[xi] = o4[c];
            d4[c] = 0.f;
            if (target != nullptr) {
                const float e = o4[c] - target[xi];
                a_loss = fmaf(e, e, a_loss);
                d4[c] = e * dscale;
            }
            if (dpred != nullptr) d4[c] = dpred[xi];
        }
        if (dh != nullptr) {
            float4 gy = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                gy.x = fmaf(d4[c], wfr[c].x, gy.x); gy.y = fmaf(d4[c], wfr[c].y, gy.y);
                gy.z = fmaf(d4[c], wfr[c].z, gy.z); gy.w = fmaf(d4[c], wfr[c].w, gy.w);
                a_wf[c].x = fmaf(d4[c], y.x, a_wf[c].x); a_wf[c].y = fmaf(d4[c], y.y, a_wf[c].y);
                a_wf[c].z = fmaf(d4[c], y.z, a_wf[c].z); a_wf[c].w = fmaf(d4[c], y.w, a_wf[c].w);
                a_bf[c] += d4[c];
            }
            a_lnw.x = fmaf(gy.x, xh.x, a_lnw.x); a_lnw.y = fmaf(gy.y, xh.y, a_lnw.y);
            a_lnw.z = fmaf(gy.z, xh.z, a_lnw.z); a_lnw.w = fmaf(gy.w, xh.w, a_lnw.w);
            a_lnb.x += gy.x; a_lnb.y += gy.y; a_lnb.z += gy.z; a_lnb.w += gy.w;
            const float4 gx = make_float4(gy.x * w4.x, gy.y * w4.y, gy.z * w4.z, gy.w * w4.w);
            const float m1 = warp_sum(gx.x + gx.y + gx.z + gx.w) * (1.f / D);
            const float m2 = warp_sum(gx.x * xh.x + gx.y * xh.y + gx.z * xh.z + gx.w * xh.w) * (1.f / D);
            st4(dh + row * D + lane * 4, make_float4(rstd * (gx.x - m1 - xh.x * m2), rstd * (gx.y - m1 - xh.y * m2),
                                                     rstd * (gx.z - m1 - xh.z * m2), rstd * (gx.w - m1 - xh.w * m2)));
        }
    }
    if (dh != nullptr) {
        cta_feature_atomic(a_lnw, dlnw, sm);
        cta_feature_atomic(a_lnb, dlnb, sm);
#pragma unroll
        for (int c = 0; c < 4; ++c) cta_feature_atomic(a_wf[c], dwf + c * D, sm);
    }
    if (dh != nullptr || target != nullptr) {
        // the loss and dbf are uniform across a warp's lanes (taken from lane 0)
        if (lane == 0) {
            sm[warp][0] = a_loss;
#pragma unroll
            for (int c = 0; c < 4; ++c) sm[warp][1 + c] = a_bf[c];
        }
        __syncthreads();
        if (threadIdx.x == 0 && target != nullptr && loss_sum != nullptr) {
            float s = 0.f;
            for (int w = 0; w < ROW_THREADS / 32; ++w) s += sm[w][0];
            atomicAdd(loss_sum, s);
        } else if (threadIdx.x >= 1 && threadIdx.x < 5 && dh != nullptr) {
            float s = 0.f;
            for (int w = 0; w < ROW_THREADS / 32; ++w) s += sm[w][threadIdx.x];
            atomicAdd(dbf + threadIdx.x - 1, s);
        }
    }
}

// ---- patch-embedding backward reductions: m4[c][pq] += sum_rows dh0[row][c] * xp[row][pq] ; s[c] += sum_rows dh0[row][c]
// out: [128][5] (m4 | s)
__global__ void __launch_bounds__(ROW_THREADS) embed_bwd_kernel(const float* __restrict__ dh0, const float* __restrict__ xp, float* __restrict__ out, int rc) {
    [[maybe_unused]] const int ntok = gridDim.y * rc, latp = ntok >> 4, lat = LATC * latp;   // tokens per sequence, latent width

    __shared__ float sm[ROW_THREADS / 32][D];
    const int seq = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 acc[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int tl = warp; tl < rc; tl += ROW_THREADS / 32) {
        const size_t row = (size_t)seq * ntok + blockIdx.y * rc + tl;
        const float4 g = ld4(dh0 + row * D + lane * 4), x4 = ld4(xp + row * 4);
        const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            acc[q].x = fmaf(g.x, xv[q], acc[q].x); acc[q].y = fmaf(g.y, xv[q], acc[q].y);
            acc[q].z = fmaf(g.z, xv[q], acc[q].z); acc[q].w = fmaf(g.w, xv[q], acc[q].w);
        }
        acc[4].x += g.x; acc[4].y += g.y; acc[4].z += g.z; acc[4].w += g.w;
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        st4(&sm[warp][lane * 4], acc[q]);
        __syncthreads();
        if (threadIdx.x < D) {
            float s = 0.f;
            for (int w = 0; w < ROW_THREADS / 32; ++w) s += sm[w][threadIdx.x];
            atomicAdd(out + threadIdx.x * 5 + q, s);
        }
        __syncthreads();
    }
}
// one block of 128 threads: red [128][5] -> dWpe [128][4], dbpe [128], dconv_w [4][4], dconv_b [4]   (all +=)
__global__ void embed_bwd_finish_kernel(const float* __restrict__ red, const float* __restrict__ wpe, const float* __restrict__ cw,
                                        const float* __restrict__ cb, float* __restrict__ dwpe, float* __restrict__ dbpe,
                                        float* __restrict__ dcw, float* __restrict__ dcb) {
    __shared__ float sred[D][5];
    const int c = threadIdx.x;
    for (int q = 0; q < 5; ++q) sred[c][q] = red[c * 5 + q];
    __syncthreads();
    // tok[oc] = sum_pq cw[oc][pq] xp[pq] + cb[oc] ; h = Wpe tok + ...  =>  dWpe[c][oc] = sum_pq m4[c][pq] cw[oc][pq] + s[c] cb[oc]
    for (int oc = 0; oc < 4; ++oc) {
        float v = sred[c][4] * cb[oc];
        for (int pq = 0; pq < 4; ++pq) v = fmaf(sred[c][pq], cw[oc * 4 + pq], v);
        dwpe[c * 4 + oc] += v;
    }
    dbpe[c] += sred[c][4];
    if (c < 16) {                         // dcw[oc][pq] = sum_c Wpe[c][oc] m4[c][pq]
        const int oc = c >> 2, pq = c & 3;
        float v = 0.f;
        for (int k = 0; k < D; ++k) v = fmaf(wpe[k * 4 + oc], sred[k][pq], v);
        dcw[c] += v;
    } else if (c < 20) {                  // dcb[oc] = sum_c Wpe[c][oc] s[c]
        const int oc = c - 16;
        float v = 0.f;
        for (int k = 0; k < D; ++k) v = fmaf(wpe[k * 4 + oc], sred[k][4], v);
        dcb[oc] += v;
    }
}

// ---- training inputs: rectified-flow interpolation (rectified_flow.py:8-12, train.py:69-71) and DDPM q_sample
// (DDPM.py:19-27, train.py:73-75).  kind 0: x_t = t x1 + (1-t) x0, target = x1 - x0;  kind 1: x_t = ca[b] x1 + cb[b] eps, target = eps
__global__ void make_train_inputs_kernel(int kind, const float* __restrict__ x1, const float* __restrict__ nz, const float* __restrict__ ca,
                                         const float* __restrict__ cb, float* __restrict__ xt, float* __restrict__ target, size_t n, int lat) {

    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / lat;
        const float a = x1[i], z = nz[i];
        if (kind == 0) {
            const float t = ca[b];
            xt[i] = t * a + (1.f - t) * z;
            target[i] = a - z;
        } else {
            xt[i] = ca[b] * a + cb[b] * z;
            target[i] = z;
        }
    }
}

// ---- fused AdamW over a flat fp32 parameter buffer (torch.optim.AdamW as configured at train.py:37)
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n,
                             float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt, float grad_scale) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float gi = g[i] * grad_scale;
        float pi = p[i] * (1.f - lr * wd);
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        pi -= (lr / bc1) * mi / denom;
        p[i] = pi; m[i] = mi; v[i] = vi;
    }
}

}  // namespace t2s

namespace t2s {
// y[c][pq] = sum_oc Wpe[c][oc] cw[oc][pq] ; b[c] = sum_oc Wpe[c][oc] cb[oc] + bpe[c]   (conv folded into patch_emb)
__global__ void embed_fold_kernel(const float* __restrict__ wpe, const float* __restrict__ bpe, const float* __restrict__ cw,
                                  const float* __restrict__ cb, float* __restrict__ w_embed, float* __restrict__ b_embed) {
    const int c = threadIdx.x;
    if (c >= D) return;
    float b = bpe[c];
    float w[4] = {0.f, 0.f, 0.f, 0.f};
    for (int oc = 0; oc < 4; ++oc) {
        const float a = wpe[c * 4 + oc];
        b = fmaf(a, cb[oc], b);
        for (int pq = 0; pq < 4; ++pq) w[pq] = fmaf(a, cw[oc * 4 + pq], w[pq]);
    }
    for (int pq = 0; pq < 4; ++pq) w_embed[c * 4 + pq] = w[pq];
    b_embed[c] = b;
}
}  // namespace t2s
