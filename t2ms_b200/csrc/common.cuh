// Shared device helpers for the T2S sm_100a kernels: PTX wrappers (tcgen05 / TMEM, bulk async copy + mbarrier,
// cp.async, packed fp32 arithmetic) and fast math.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace t2s {

// Model constants (reference: model/denoiser/transformer.py:94-105,127-149)
constexpr int D = 128;          // d_model
constexpr int NHEAD = 4;
constexpr int HD = 32;          // head dim
constexpr int NLAYER = 4;
constexpr int DMLP = 256;
constexpr int LATC = 64;        // latent channels (axis 1 of the (B,64,30) latent)
constexpr int MOD = 6 * D;      // adaLN chunk: shift/scale/gate (msa) + shift/scale/gate (mlp)

// Token-local kernels work on "pair tiles": TILE_TOK tokens of two consecutive sequences
// (uncond/cond of one sample in CFG sampling) = rows 0..TILE_TOK-1 and 64..64+TILE_TOK-1 of a 128-row tile.
constexpr int TILE_ROWS = 128;
constexpr int STAGE_BYTES = 32768;                // one weight stage: [128][128] fp16

// Shape of the denoiser for a latent of H positions (axis 2 of the (B,64,H) latent): H = 30 is the T2S model
// (model/denoiser/transformer.py:132), H = 50 / 64 the fork's Transformer(dim) (model/denoiser/mytransformer.py:128-136,
// config.yaml:46,91).  Tokens = (H/2) x 32 patches; token n = i*32 + j covers latent rows 2j, 2j+1 and positions 2i, 2i+1.
template <int H_>
struct DitShape {
    static_assert(H_ == 30 || H_ == 50 || H_ == 64, "supported latent widths: 30 (T2S), 50 and 64 (fork configs)");
    static constexpr int H = H_;
    static constexpr int NTOK = 16 * H_;                                              // 480 | 800 | 1024
    static constexpr int LAT = LATC * H_;
    static constexpr int TILE_TOK = H_ == 30 ? 60 : (H_ == 50 ? 50 : 64);             // tokens per pair tile
    static constexpr int TILES_PER_PAIR = NTOK / TILE_TOK;                            // 8 | 16 | 16 (even: work item = two tiles)
    // attention: NQT query tiles of QT_ROWS valid rows (M = 128), NCH key chunks of KC keys
    static constexpr int QT_ROWS = H_ == 30 ? 120 : (H_ == 50 ? 100 : 128);
    static constexpr int NQT = NTOK / QT_ROWS;                                        // 4 | 8 | 8
    static constexpr int KC = H_ == 30 ? 48 : 32;        // two S buffers + O in 128 TMEM columns: four softmax warpgroups per SM
    static constexpr int NCH = NTOK / KC;                                             // 10 | 25 | 32
    // q|k|v scratch: per (sequence, head) three tcgen05 operand images (layouts at attn_kernel)
    static constexpr int Q_HALVES = NQT * 4 * 128 * 8;
    static constexpr int K_HALVES = 4 * NTOK * 8;
    static constexpr int V_HALVES = NTOK * HD;
    static constexpr int HEAD_HALVES = Q_HALVES + K_HALVES + V_HALVES;
    static_assert(TILES_PER_PAIR * TILE_TOK == NTOK && TILES_PER_PAIR % 2 == 0 && NQT * QT_ROWS == NTOK && NCH * KC == NTOK && KC % 16 == 0, "shape");
};

// The T2S shape (H = 30), used by the training path
constexpr int NTOK = DitShape<30>::NTOK;          // (30/2)*(64/2) patches
constexpr int LATP = 30;                          // latent positions (axis 2)
constexpr int LAT = LATC * LATP;
constexpr int TILE_TOK = DitShape<30>::TILE_TOK;
constexpr int TILES_PER_PAIR = DitShape<30>::TILES_PER_PAIR;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// fp32 pair -> packed fp16, round-to-nearest, SATURATING: a value beyond the fp16 range becomes +-65504 instead of inf,
// so an out-of-range activation (LN-modulated input, GELU output, q|k|v, attention output) degrades accuracy instead of
// poisoning the sequence with inf / NaN (tests/test_gpu_config_parity.py: fp16 range stress).  One F2FP either way.
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// ---------------------------------------------------------------- fast math (MUFU)
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// Packed fp32 arithmetic (SASS FFMA2 / FADD2): two independent fp32 operations per issue slot.
__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
    unsigned long long a, b, c, d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(c0), "f"(c1));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}
__device__ __forceinline__ void add2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    unsigned long long a, b, d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}
__device__ __forceinline__ void mul2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    unsigned long long a, b, d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// GELU(tanh), timm Mlp act (model/denoiser/transformer.py:99): two values with packed fp32 arithmetic,
// 0.5 x (1 + tanh(x (k0 + k0 k1 x^2)))
__device__ __forceinline__ void gelu_tanh2(float& y0, float& y1, float x0, float x1) {
    const float k0 = 0.7978845608028654f, k01 = 0.7978845608028654f * 0.044715f;
    float s0, s1, p0, p1, u0, u1, h0, h1;
    mul2(s0, s1, x0, x1, x0, x1);
    fma2(p0, p1, s0, s1, k01, k01, k0, k0);
    mul2(u0, u1, p0, p1, x0, x1);
#ifdef T2S_PRECISE_GELU
    const float t0 = tanhf(u0), t1 = tanhf(u1);
#else
#ifdef T2S_KO_MUFU
    const float t0 = u0, t1 = u1;                       // timing knock-out (wrong results)
#else
    const float t0 = tanh_approx(u0), t1 = tanh_approx(u1);
#endif
#endif
    mul2(h0, h1, x0, x1, 0.5f, 0.5f);
    fma2(y0, y1, h0, h1, t0, t1, h0, h1);
}

// ---------------------------------------------------------------- counter-based Gaussian noise
// Philox4x32-10 (Salmon et al., SC'11) keyed by a 64-bit seed on the counter (element index, step), followed by one
// Box-Muller transform: a standard normal per (seed, step, element), stateless and order-independent, so the DDPM
// ancestral update (DDPM.py:35 draws torch.randn inside p_sample) needs no noise tensor and no RNG kernel between steps.
// The parity tests restate it in numpy and feed those values to the CPU reference loop as the step noise.
__device__ __forceinline__ float philox_normal(unsigned long long seed, unsigned int step, unsigned long long element) {
    uint32_t c0 = (uint32_t)element, c1 = (uint32_t)(element >> 32), c2 = step, c3 = 0u;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    const float u1 = ((float)(c0 >> 8) + 1.0f) * 5.9604644775390625e-8f;     // (0, 1]   (24 bits)
    const float u2 = (float)(c1 >> 8) * 5.9604644775390625e-8f;              // [0, 1)
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// ---------------------------------------------------------------- async copies
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }

// mbarrier + bulk async copy (the TMA engine without a tensor map: SASS UBLKCP)
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// Bounded wait: a lost completion traps (kernel error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    long long t0 = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile("{\n.reg .pred p;\n"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                     "selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
#ifdef T2S_WAIT_NS
        if (!done) asm volatile("nanosleep.u32 %0;\n" :: "r"((uint32_t)T2S_WAIT_NS));   // A/B: back-off between polls
#endif
        if (!done && (spins & 15) == 15) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000LL) __trap();     // ~2 s at 2 GHz
        }
    }
}


__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(bar) : "memory");
}
// One arrival per WARP instead of one per thread: every arrival on an mbarrier wakes its suspended waiters (the MMA issuer
// spun ~550 times per 128-arrival phase, each spin an MIO instruction queued in front of the compute warps' TMEM traffic).
// Call with the whole warp converged, after each lane's own fences; the barrier's count is the number of WARPS.
__device__ __forceinline__ void mbar_arrive_warp(uint32_t bar) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// ---------------------------------------------------------------- tcgen05 / TMEM (5th-gen tensor cores)
// Shared-memory matrix descriptor, no-swizzle K-major canonical layout: core matrix = 8 rows x 16 B stored
// contiguously (128 B); LBO = byte stride between core matrices adjacent in K, SBO = byte stride between
// 8-row groups (semantics verified on B200 by tools/probes.cu, profiles/r01_probes.log).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                       // descriptor version 1 (sm_100)
    return d;                                     // layout_type 0 = no swizzle, base_offset 0
}
// Instruction descriptor for kind::f16: fp16 A/B (K-major), fp32 accumulate, shape M x N x 16.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Arrives on the mbarrier once every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {      // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {       // the allocating warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(taddr), "r"(ncols) : "memory");
}
// TMEM -> registers: lane (32*(warp%4) + laneid), 32 consecutive fp32 columns. Follow with tmem_wait_ld().
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t* u = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                   "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
                   "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]),
                   "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t* u = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                   "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    const uint32_t* u = reinterpret_cast<const uint32_t*>(v);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
                 :: "r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]),
                    "r"(u[8]), "r"(u[9]), "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15]) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
// L2 prefetch of a contiguous global range (one thread)
__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" :: "l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// D[tmem] (+)= A[tmem] . B[smem]: the A operand (M128 x K16 fp16, lane = row, column c = elements 2c, 2c+1) is read
// from tensor memory
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                 :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&u)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n"
                 :: "r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]) : "memory");
}
// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// The kernels of a sampling step are launched with programmatic stream serialization: a kernel's CTAs may start (barrier
// set-up, TMEM allocation) while the previous kernel's last CTAs are still running.  pdl_launch_dependents() lets the next
// grid start launching; pdl_wait() blocks until the previous grid has completed and its writes are visible — it must precede
// the first read of anything the previous kernel produced.  Both are no-ops in a launch without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }

// ---------------------------------------------------------------- gpu-scope flags (dataflow between CTAs of one launch)
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;\n" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_gpu(int* p, int v) {
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;\n" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int atom_add_acq_rel_gpu(int* p, int v) {
    int old;
    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], %2;\n" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ int atom_cas_relaxed_gpu(int* p, int cmp, int val) {
    int old;
    asm volatile("atom.relaxed.gpu.global.cas.b32 %0, [%1], %2, %3;\n" : "=r"(old) : "l"(p), "r"(cmp), "r"(val) : "memory");
    return old;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;\n" ::: "memory"); }
// orders this thread's earlier generic-proxy accesses (the acquire of a flag) before its later async-proxy accesses
// (bulk copies that read what another SM wrote with ordinary stores)
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
__device__ __forceinline__ void nanosleep(uint32_t ns) { asm volatile("nanosleep.u32 %0;\n" :: "r"(ns) : "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" :: "r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ float max3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}


}  // namespace t2s
