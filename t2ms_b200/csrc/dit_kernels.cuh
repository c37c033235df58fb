// T2S-DiT denoiser kernels for sm_100a (reference: model/denoiser/transformer.py:94-193 and the
// timm Attention / Mlp it calls).  Arithmetic: fp16 operands (10-bit mantissa, = tf32 operand
// precision) with fp32 accumulation on the tensor cores; residual stream, LayerNorm, softmax
// statistics, modulation and the sampler update in fp32.
//
// Kernels (one sampling step = cond + embed_qkv + 4 x (attention + token)):
//   cond_kernel       time embedding + text conditioning + SiLU + adaLN Linear for all 4 blocks
//   token_kernel<EMBED>  patch-embed + pos  -> h ; LN1+modulate -> QKV(l=0)
//   attn_kernel       softmax(q k^T / sqrt(32)) v per (sequence, head), flash-style, P kept in registers
//   token_kernel<MID>    proj+gate+residual, LN2+modulate, fc1+GELU, fc2+gate+residual -> h ;
//                        LN1+modulate -> QKV(l+1)
//   token_kernel<FINAL>  ... + final LN + Linear(128->4) + unpatchify + CFG mix + Euler / DDPM update
#pragma once
#include "common.cuh"

namespace t2s {

struct DitWeights {                 // device pointers; mirrors t2s_dit_weights (include/t2s_b200.h)
    const __half* w_qkv[NLAYER];    // 3 stages  [128 n][128 k] fp16, 16B-chunk XOR swizzle (q | k | v)
    const __half* w_post[NLAYER];   // 5 stages: proj, then 4 x { fc1 rows c*64.. [64][128] | fc2 cols c*64.. [128][64] }
    const float* b_qkv[NLAYER];     // [384]
    const float* b_proj[NLAYER];    // [128]
    const float* b_fc1[NLAYER];     // [256]
    const float* b_fc2[NLAYER];     // [128]
    const float* w_ada_t;           // [4][128 k][768 o]  (adaLN Linear weight, transposed)
    const float* b_ada;             // [4][768]
    const float* w_embed;           // [128][4]   patch_emb.weight @ conv.weight  (folded)
    const float* b_embed;           // [128]      patch_emb.weight @ conv.bias + patch_emb.bias
    const float* pos;               // [480][128]
    const float* w_final;           // [4][128]   linear_emb_to_patch.weight * ln.weight
    const float* b_final;           // [4]        linear_emb_to_patch.weight @ ln.bias + bias
    const float* freqs;             // [64]       10000 ** linspace(0,1,64)
};

enum TokenMode { TOK_EMBED = 0, TOK_MID = 1, TOK_FINAL = 2 };
enum OutMode { OUT_FWD = 0, OUT_RF = 1, OUT_DDPM = 2 };

struct TokArgs {
    DitWeights w;
    const float* x;        // latents [(nseq >> x_shift)][64][30]
    int x_shift;           // 1 when the two sequences of a pair share one latent (CFG), else 0
    float* h;              // residual stream  [nseq][480][128] fp32
    __half* qkv;           // [nseq][4 heads][3][480][32] fp16, 16B chunks XOR-swizzled by (tok>>1)&3
    const __half* o;       // attention output [nseq][480][128] fp16
    const float* mod;      // adaLN modulation [nseq][4][768] fp32
    int nseq;
    int layer;             // block whose post-attention half runs here (MID / FINAL)
    // FINAL only
    int out_mode;
    float* out;            // OUT_FWD: [nseq][64][30]; OUT_RF / OUT_DDPM: optional guided prediction [npair][64][30]
    float* x_upd;          // OUT_RF / OUT_DDPM: latent updated in place [npair][64][30]
    const float* noise;    // OUT_DDPM: [npair][64][30] for this step
    float cfg, c1, c2, c3; // RF: x += pred*c1 ; DDPM: x = c1*(x - c2*pred) + c3*noise
};

// =================================================================================== cond
// mod[seq][l][:] = Linear_l( SiLU( temb(t) (+ text) ) )      transformer.py:30-40,106-109,115,174-178
// grid (ceil(nseq/8), 4), block 256
__global__ void __launch_bounds__(256) cond_kernel(float* __restrict__ mod, const float* __restrict__ t100, int t_stride,
                                                   const float* __restrict__ emb, int emb_shift, int cfg_pairs,
                                                   const float* __restrict__ freqs, const float* __restrict__ w_ada_t,
                                                   const float* __restrict__ b_ada, int nseq) {
    __shared__ float sc[8][D];
    const int s0 = blockIdx.x * 8, l = blockIdx.y, tid = threadIdx.x;
    for (int i = tid; i < 8 * D; i += 256) {
        const int si = i >> 7, f = i & 127, seq = s0 + si;
        float v = 0.f;
        if (seq < nseq) {
            const float arg = __fdiv_rn(t100[(size_t)seq * t_stride], freqs[f & 63]);
            float c = (f < 64) ? sinf(arg) : cosf(arg);
            if (emb != nullptr && (!cfg_pairs || (seq & 1))) c = c + emb[(size_t)(seq >> emb_shift) * D + f];
            v = c / (1.0f + expf(-c));
        }
        sc[si][f] = v;
    }
    __syncthreads();
    float acc[3][8];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int s = 0; s < 8; ++s) acc[a][s] = 0.f;
    const float* w = w_ada_t + (size_t)l * D * MOD;
#pragma unroll 4
    for (int k = 0; k < D; ++k) {
        const float w0 = w[k * MOD + tid], w1 = w[k * MOD + 256 + tid], w2 = w[k * MOD + 512 + tid];
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const float c = sc[s][k];
            acc[0][s] = fmaf(w0, c, acc[0][s]);
            acc[1][s] = fmaf(w1, c, acc[1][s]);
            acc[2][s] = fmaf(w2, c, acc[2][s]);
        }
    }
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        const int seq = s0 + s;
        if (seq < nseq) {
            float* dst = mod + ((size_t)seq * NLAYER + l) * MOD;
#pragma unroll
            for (int a = 0; a < 3; ++a) dst[a * 256 + tid] = acc[a][s] + b_ada[l * MOD + a * 256 + tid];
        }
    }
}

// =================================================================================== warp GEMM helpers
// acc[NB n8-blocks][4] += A(16 rows x 16*KS) . W^T, A in registers (mma A fragments), W stage in smem as
// [n][k] fp16 rows of ROWB bytes with 16B chunks XOR-swizzled by (n & 7).
template <int NB, int KS, int ROWB>
__device__ __forceinline__ void warp_gemm_rega(float (&acc)[NB][4], const uint32_t (&a)[KS][4], uint32_t wbase, int lane) {
    const int l7 = lane & 7;
    const int kc = (lane >> 3) & 1;
    const uint32_t lane_base = wbase + (uint32_t)(l7 + ((lane >> 4) << 3)) * ROWB;
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
        const uint32_t koff = (uint32_t)(((2 * kk + kc) ^ l7) << 4);
#pragma unroll
        for (int jj = 0; jj < NB / 2; ++jj) {
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4(b0, b1, b2, b3, lane_base + jj * 16 * ROWB + koff);
            mma_f16(acc[2 * jj], a[kk][0], a[kk][1], a[kk][2], a[kk][3], b0, b1);
            mma_f16(acc[2 * jj + 1], a[kk][0], a[kk][1], a[kk][2], a[kk][3], b2, b3);
        }
    }
}

// Same with A read from a swizzled smem tile [rows][128 fp16] (256 B rows, chunk ^ (row & 7)).
template <int NB>
__device__ __forceinline__ void warp_gemm_smema(float (&acc)[NB][4], uint32_t abase, int row0, uint32_t wbase, int lane) {
    const int l7 = lane & 7;
    const int kc = (lane >> 3) & 1;
    const uint32_t lane_base = wbase + (uint32_t)(l7 + ((lane >> 4) << 3)) * 256;
    const int arow = row0 + (lane & 15);
    const uint32_t a_lane = abase + arow * 256;
    const int a7 = arow & 7, ah = lane >> 4;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
        uint32_t a0, a1, a2, a3;
        ldmatrix_x4(a0, a1, a2, a3, a_lane + (uint32_t)(((2 * kk + ah) ^ a7) << 4));
        const uint32_t koff = (uint32_t)(((2 * kk + kc) ^ l7) << 4);
#pragma unroll
        for (int jj = 0; jj < NB / 2; ++jj) {
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4(b0, b1, b2, b3, lane_base + jj * 16 * 256 + koff);
            mma_f16(acc[2 * jj], a0, a1, a2, a3, b0, b1);
            mma_f16(acc[2 * jj + 1], a0, a1, a2, a3, b2, b3);
        }
    }
}

// LayerNorm (no affine) over the 128 features of the two rows a thread quad holds in accumulator
// layout, then modulate x*(1+scale)+shift (transformer.py:7-8,102-103,116-117) and repack as fp16
// mma A fragments for the next GEMM.  Row statistics: warp-shuffle reduction over the quad.
__device__ __forceinline__ void ln_mod_afrag(const float (&x)[16][4], uint32_t (&a)[8][4], const float* __restrict__ shift,
                                             const float* __restrict__ scale, float eps, int t) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) { s0 += x[j][0] + x[j][1]; s1 += x[j][2] + x[j][3]; }
    const float m0 = quad_sum(s0) * (1.f / D), m1 = quad_sum(s1) * (1.f / D);
    float v0 = 0.f, v1 = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        float d;
        d = x[j][0] - m0; v0 = fmaf(d, d, v0);
        d = x[j][1] - m0; v0 = fmaf(d, d, v0);
        d = x[j][2] - m1; v1 = fmaf(d, d, v1);
        d = x[j][3] - m1; v1 = fmaf(d, d, v1);
    }
    const float r0 = rsqrtf(quad_sum(v0) * (1.f / D) + eps), r1 = rsqrtf(quad_sum(v1) * (1.f / D) + eps);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int f = 8 * j + 2 * t;
        const float2 sc = *reinterpret_cast<const float2*>(scale + f);
        const float2 sh = *reinterpret_cast<const float2*>(shift + f);
        const float y00 = fmaf((x[j][0] - m0) * r0, 1.f + sc.x, sh.x);
        const float y01 = fmaf((x[j][1] - m0) * r0, 1.f + sc.y, sh.y);
        const float y10 = fmaf((x[j][2] - m1) * r1, 1.f + sc.x, sh.x);
        const float y11 = fmaf((x[j][3] - m1) * r1, 1.f + sc.y, sh.y);
        a[j >> 1][(j & 1) * 2 + 0] = pack_h2(y00, y01);
        a[j >> 1][(j & 1) * 2 + 1] = pack_h2(y10, y11);
    }
}

// Weight-stage ring: 2 x 32 KB smem buffers filled by bulk async copies (TMA engine, UBLKCP),
// completion on mbarriers; the consumer side is the whole CTA (a __syncthreads releases a buffer).
struct StageRing {
    uint32_t buf, bar;                // smem addresses: 2 buffers, 2 mbarriers
    const char* src_a; int n_a;       // first n_a stages come from src_a, the rest from src_b
    const char* src_b; int n_total;
    __device__ __forceinline__ const char* stage_src(int s) const {
        return s < n_a ? src_a + (size_t)s * STAGE_BYTES : src_b + (size_t)(s - n_a) * STAGE_BYTES;
    }
    __device__ __forceinline__ void issue(int s) const {     // one thread
        const uint32_t b = bar + (s & 1) * 8;
        mbar_expect_tx(b, STAGE_BYTES);
        bulk_g2s(buf + (s & 1) * STAGE_BYTES, stage_src(s), STAGE_BYTES, b);
    }
    __device__ __forceinline__ uint32_t wait(int s) const {  // all threads; returns the stage's smem address
        mbar_wait(bar + (s & 1) * 8, (s >> 1) & 1);
        return buf + (s & 1) * STAGE_BYTES;
    }
    __device__ __forceinline__ void release(int s, int tid) const {  // all threads
        __syncthreads();
        if (tid == 0 && s + 2 < n_total) issue(s + 2);
    }
};

// LN1+modulate of block `l` -> QKV GEMM (3 stages starting at ring stage s0) -> q|k|v stored fp16 in the
// attention kernel's smem image layout.
__device__ __forceinline__ void qkv_phase(const float (&hreg)[16][4], const TokArgs& p, const StageRing& ring, int s0, int l,
                                          int seq, bool v0, bool v1, int tok0, int tok1, int lane, int tid) {
    const int t = lane & 3;
    const float* mod = p.mod + ((size_t)seq * NLAYER + l) * MOD;
    uint32_t a[8][4];
    ln_mod_afrag(hreg, a, mod /*shift_msa*/, mod + D /*scale_msa*/, 1e-6f, t);
    const float* bq = p.w.b_qkv[l];
#pragma unroll 1
    for (int which = 0; which < 3; ++which) {
        const uint32_t wb = ring.wait(s0 + which);
        float acc[16][4];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
        warp_gemm_rega<16, 8, 256>(acc, a, wb, lane);
        ring.release(s0 + which, tid);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int f = 8 * j + 2 * t, head = j >> 2, chunk = j & 3;
            const float2 b = *reinterpret_cast<const float2*>(bq + which * D + f);
            __half* base = p.qkv + (((size_t)seq * NHEAD + head) * 3 + which) * (NTOK * HD) + 2 * t;
            if (v0) *reinterpret_cast<uint32_t*>(base + tok0 * HD + ((chunk ^ ((tok0 >> 1) & 3)) << 3)) =
                        pack_h2(acc[j][0] + b.x, acc[j][1] + b.y);
            if (v1) *reinterpret_cast<uint32_t*>(base + tok1 * HD + ((chunk ^ ((tok1 >> 1) & 3)) << 3)) =
                        pack_h2(acc[j][2] + b.x, acc[j][3] + b.y);
        }
    }
}

// smem carve-up of token_kernel
constexpr int TOK_SMEM_W = 0;                                   // 2 x 32 KB weight stages
constexpr int TOK_SMEM_O = 2 * STAGE_BYTES;                     // attention-output tile, fp16 [128][128] swizzled
constexpr int HS_LD = 136;                                      // fp32 row stride of the h tile (bank-conflict-free float2)
constexpr int TOK_SMEM_H = TOK_SMEM_O + TILE_ROWS * D * 2;      // residual tile fp32 [128][136]
constexpr int TOK_SMEM_BAR = TOK_SMEM_H + TILE_ROWS * HS_LD * 4;
constexpr int TOK_SMEM_V = TOK_SMEM_BAR + 64;                   // [128][4] fp32 final projection exchange
constexpr int TOK_SMEM_BYTES = TOK_SMEM_V + TILE_ROWS * 4 * 4;

// grid = npair * 8 tiles, block = 256 (8 warps x 16 rows)
template <int MODE>
__global__ void __launch_bounds__(256, 1) token_kernel(const TokArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int pair = blockIdx.x / TILES_PER_PAIR, tt = blockIdx.x % TILES_PER_PAIR;
    const uint32_t s_base = smem_u32(smem);

    // rows owned by this thread: r0 = warp*16 + g and r0 + 8 (same sequence: branch = warp >> 2)
    const int branch = warp >> 2;
    const int seq = 2 * pair + branch;
    const bool seq_ok = seq < p.nseq;
    const int tl0 = (warp & 3) * 16 + g, tl1 = tl0 + 8;
    const bool v0 = seq_ok && tl0 < TILE_TOK, v1 = seq_ok && tl1 < TILE_TOK;
    const int tok0 = tt * TILE_TOK + tl0, tok1 = tt * TILE_TOK + tl1;
    const int seq_c = seq_ok ? seq : 0;                           // clamped for address formation
    const int l = p.layer;

    StageRing ring;
    ring.buf = s_base + TOK_SMEM_W;
    ring.bar = s_base + TOK_SMEM_BAR;
    if (MODE == TOK_EMBED) {
        ring.src_a = reinterpret_cast<const char*>(p.w.w_qkv[0]); ring.n_a = 3; ring.src_b = nullptr; ring.n_total = 3;
    } else if (MODE == TOK_MID) {
        ring.src_a = reinterpret_cast<const char*>(p.w.w_post[l]); ring.n_a = 5;
        ring.src_b = reinterpret_cast<const char*>(p.w.w_qkv[l + 1]); ring.n_total = 8;
    } else {
        ring.src_a = reinterpret_cast<const char*>(p.w.w_post[l]); ring.n_a = 5; ring.src_b = nullptr; ring.n_total = 5;
    }
    if (tid == 0) {
        mbar_init(ring.bar, 1);
        mbar_init(ring.bar + 8, 1);
        mbar_fence_init();
        ring.issue(0);
        ring.issue(1);
    }

    float hreg[16][4];   // residual rows in accumulator layout: [j][0..1] row r0, [j][2..3] row r0+8, cols 8j+2t,+1

    if (MODE == TOK_EMBED) {
        // ---- patchify + patch_emb + pos_embed (transformer.py:166-172), conv folded into the Linear
        float xa[4] = {0.f, 0.f, 0.f, 0.f}, xb[4] = {0.f, 0.f, 0.f, 0.f};
        const float* xs = p.x + (size_t)(seq_c >> p.x_shift) * LAT;
        if (v0) {
            const int i = tok0 >> 5, j = tok0 & 31;
#pragma unroll
            for (int pq = 0; pq < 4; ++pq) xa[pq] = xs[(2 * j + (pq & 1)) * LATP + 2 * i + (pq >> 1)];
        }
        if (v1) {
            const int i = tok1 >> 5, j = tok1 & 31;
#pragma unroll
            for (int pq = 0; pq < 4; ++pq) xb[pq] = xs[(2 * j + (pq & 1)) * LATP + 2 * i + (pq >> 1)];
        }
        float* hg = p.h + (size_t)seq_c * NTOK * D;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int f = 8 * j + 2 * t;
            const float4 wa = *reinterpret_cast<const float4*>(p.w.w_embed + f * 4);
            const float4 wb = *reinterpret_cast<const float4*>(p.w.w_embed + f * 4 + 4);
            const float2 be = *reinterpret_cast<const float2*>(p.w.b_embed + f);
            float2 h0 = make_float2(0.f, 0.f), h1 = make_float2(0.f, 0.f);
            if (v0) {
                const float2 pe = *reinterpret_cast<const float2*>(p.w.pos + tok0 * D + f);
                h0.x = wa.x * xa[0] + wa.y * xa[1] + wa.z * xa[2] + wa.w * xa[3] + be.x + pe.x;
                h0.y = wb.x * xa[0] + wb.y * xa[1] + wb.z * xa[2] + wb.w * xa[3] + be.y + pe.y;
                *reinterpret_cast<float2*>(hg + tok0 * D + f) = h0;
            }
            if (v1) {
                const float2 pe = *reinterpret_cast<const float2*>(p.w.pos + tok1 * D + f);
                h1.x = wa.x * xb[0] + wa.y * xb[1] + wa.z * xb[2] + wa.w * xb[3] + be.x + pe.x;
                h1.y = wb.x * xb[0] + wb.y * xb[1] + wb.z * xb[2] + wb.w * xb[3] + be.y + pe.y;
                *reinterpret_cast<float2*>(hg + tok1 * D + f) = h1;
            }
            hreg[j][0] = h0.x; hreg[j][1] = h0.y; hreg[j][2] = h1.x; hreg[j][3] = h1.y;
        }
        __syncthreads();    // mbarrier init visible to all waiters
        qkv_phase(hreg, p, ring, 0, 0, seq_c, v0, v1, tok0, tok1, lane, tid);
        return;
    }

    // ---- MID / FINAL: stage the attention-output tile (fp16) and the residual tile (fp32) in smem
    {
        const uint32_t so = s_base + TOK_SMEM_O;
        for (int q = tid; q < TILE_ROWS * 16; q += 256) {
            const int r = q >> 4, c = q & 15, sq = 2 * pair + (r >> 6), tl = r & 63;
            const uint32_t dst = so + r * 256 + ((c ^ (r & 7)) << 4);
            if (tl < TILE_TOK && sq < p.nseq)
                cp_async16(dst, p.o + ((size_t)sq * NTOK + tt * TILE_TOK + tl) * D + c * 8);
            else
                *reinterpret_cast<uint4*>(smem + TOK_SMEM_O + r * 256 + ((c ^ (r & 7)) << 4)) = make_uint4(0, 0, 0, 0);
        }
        const uint32_t sh = s_base + TOK_SMEM_H;
        for (int q = tid; q < TILE_ROWS * 32; q += 256) {
            const int r = q >> 5, c = q & 31, sq = 2 * pair + (r >> 6), tl = r & 63;
            if (tl < TILE_TOK && sq < p.nseq)
                cp_async16(sh + (r * HS_LD + c * 4) * 4, p.h + ((size_t)sq * NTOK + tt * TILE_TOK + tl) * D + c * 4);
            else
                *reinterpret_cast<float4*>(smem + TOK_SMEM_H + (r * HS_LD + c * 4) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();
    }
    float* hs = reinterpret_cast<float*>(smem + TOK_SMEM_H);
    const int r0 = warp * 16 + g, r1 = r0 + 8;
    const float* mod = p.mod + ((size_t)seq_c * NLAYER + l) * MOD;

    // ---- attention out-projection + gate + residual   x = x + gate_msa * proj(o)   (transformer.py:116)
    {
        const uint32_t wb = ring.wait(0);
#pragma unroll
        for (int j = 0; j < 16; ++j) hreg[j][0] = hreg[j][1] = hreg[j][2] = hreg[j][3] = 0.f;
        warp_gemm_smema<16>(hreg, s_base + TOK_SMEM_O, warp * 16, wb, lane);
        ring.release(0, tid);
        const float* gate = mod + 2 * D;
        const float* bp = p.w.b_proj[l];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int f = 8 * j + 2 * t;
            const float2 gv = *reinterpret_cast<const float2*>(gate + f);
            const float2 bv = *reinterpret_cast<const float2*>(bp + f);
            float2 h0 = *reinterpret_cast<float2*>(hs + r0 * HS_LD + f);
            float2 h1 = *reinterpret_cast<float2*>(hs + r1 * HS_LD + f);
            h0.x = fmaf(gv.x, hreg[j][0] + bv.x, h0.x);
            h0.y = fmaf(gv.y, hreg[j][1] + bv.y, h0.y);
            h1.x = fmaf(gv.x, hreg[j][2] + bv.x, h1.x);
            h1.y = fmaf(gv.y, hreg[j][3] + bv.y, h1.y);
            *reinterpret_cast<float2*>(hs + r0 * HS_LD + f) = h0;
            *reinterpret_cast<float2*>(hs + r1 * HS_LD + f) = h1;
            hreg[j][0] = h0.x; hreg[j][1] = h0.y; hreg[j][2] = h1.x; hreg[j][3] = h1.y;
        }
    }
    // ---- MLP: x = x + gate_mlp * fc2(GELU(fc1(modulate(LN2(x)))))   (transformer.py:117), hidden in 4 chunks of 64
    {
        uint32_t a2[8][4];
        ln_mod_afrag(hreg, a2, mod + 3 * D, mod + 4 * D, 1e-6f, t);
        float acc2[16][4];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc2[j][0] = acc2[j][1] = acc2[j][2] = acc2[j][3] = 0.f;
        const float* b1 = p.w.b_fc1[l];
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            const uint32_t wb = ring.wait(1 + c);
            float acc1[8][4];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc1[j][0] = acc1[j][1] = acc1[j][2] = acc1[j][3] = 0.f;
            warp_gemm_rega<8, 8, 256>(acc1, a2, wb, lane);
            uint32_t hf[4][4];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 b = *reinterpret_cast<const float2*>(b1 + c * 64 + 8 * j + 2 * t);
                hf[j >> 1][(j & 1) * 2 + 0] = pack_h2(gelu_tanh(acc1[j][0] + b.x), gelu_tanh(acc1[j][1] + b.y));
                hf[j >> 1][(j & 1) * 2 + 1] = pack_h2(gelu_tanh(acc1[j][2] + b.x), gelu_tanh(acc1[j][3] + b.y));
            }
            warp_gemm_rega<16, 4, 128>(acc2, hf, wb + 16384, lane);
            ring.release(1 + c, tid);
        }
        const float* gate = mod + 5 * D;
        const float* b2 = p.w.b_fc2[l];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int f = 8 * j + 2 * t;
            const float2 gv = *reinterpret_cast<const float2*>(gate + f);
            const float2 bv = *reinterpret_cast<const float2*>(b2 + f);
            const float2 h0 = *reinterpret_cast<const float2*>(hs + r0 * HS_LD + f);
            const float2 h1 = *reinterpret_cast<const float2*>(hs + r1 * HS_LD + f);
            hreg[j][0] = fmaf(gv.x, acc2[j][0] + bv.x, h0.x);
            hreg[j][1] = fmaf(gv.y, acc2[j][1] + bv.y, h0.y);
            hreg[j][2] = fmaf(gv.x, acc2[j][2] + bv.x, h1.x);
            hreg[j][3] = fmaf(gv.y, acc2[j][3] + bv.y, h1.y);
        }
    }

    if (MODE == TOK_MID) {
        float* hg = p.h + (size_t)seq_c * NTOK * D;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int f = 8 * j + 2 * t;
            if (v0) *reinterpret_cast<float2*>(hg + tok0 * D + f) = make_float2(hreg[j][0], hreg[j][1]);
            if (v1) *reinterpret_cast<float2*>(hg + tok1 * D + f) = make_float2(hreg[j][2], hreg[j][3]);
        }
        qkv_phase(hreg, p, ring, 5, l + 1, seq_c, v0, v1, tok0, tok1, lane, tid);
        return;
    }

    // ---- FINAL: LN(eps 1e-5, affine folded) + Linear(128->4) + unpatchify (transformer.py:182-190)
    {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) { s0 += hreg[j][0] + hreg[j][1]; s1 += hreg[j][2] + hreg[j][3]; }
        const float m0 = quad_sum(s0) * (1.f / D), m1 = quad_sum(s1) * (1.f / D);
        float q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float d;
            d = hreg[j][0] - m0; q0 = fmaf(d, d, q0);
            d = hreg[j][1] - m0; q0 = fmaf(d, d, q0);
            d = hreg[j][2] - m1; q1 = fmaf(d, d, q1);
            d = hreg[j][3] - m1; q1 = fmaf(d, d, q1);
        }
        const float rs0 = rsqrtf(quad_sum(q0) * (1.f / D) + 1e-5f), rs1 = rsqrtf(quad_sum(q1) * (1.f / D) + 1e-5f);
        float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int f = 8 * j + 2 * t;
            const float y00 = (hreg[j][0] - m0) * rs0, y01 = (hreg[j][1] - m0) * rs0;
            const float y10 = (hreg[j][2] - m1) * rs1, y11 = (hreg[j][3] - m1) * rs1;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const float2 w = *reinterpret_cast<const float2*>(p.w.w_final + c4 * D + f);
                d0[c4] = fmaf(y00, w.x, fmaf(y01, w.y, d0[c4]));
                d1[c4] = fmaf(y10, w.x, fmaf(y11, w.y, d1[c4]));
            }
        }
        float* vb = reinterpret_cast<float*>(smem + TOK_SMEM_V);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
            const float e0 = quad_sum(d0[c4]) + p.w.b_final[c4];
            const float e1 = quad_sum(d1[c4]) + p.w.b_final[c4];
            if (t == c4) { vb[r0 * 4 + c4] = e0; vb[r1 * 4 + c4] = e1; }
        }
        __syncthreads();
        if (p.out_mode == OUT_FWD) {
            for (int idx = tid; idx < 2 * TILE_TOK * 4; idx += 256) {
                const int br = idx / (TILE_TOK * 4), rem = idx - br * (TILE_TOK * 4), tl = rem >> 2, c4 = rem & 3;
                const int sq = 2 * pair + br;
                if (sq < p.nseq) {
                    const int n = tt * TILE_TOK + tl, i = n >> 5, jx = n & 31;
                    p.out[(size_t)sq * LAT + (2 * jx + (c4 & 1)) * LATP + 2 * i + (c4 >> 1)] = vb[(br * 64 + tl) * 4 + c4];
                }
            }
        } else if (tid < TILE_TOK * 4) {
            // classifier-free guidance mix (infer.py:81/:87) + Euler (rectified_flow.py:5-7) or
            // DDPM ancestral update (DDPM.py:28-36), latent updated in place
            const int tl = tid >> 2, c4 = tid & 3;
            const int n = tt * TILE_TOK + tl, i = n >> 5, jx = n & 31;
            const size_t xi = (size_t)pair * LAT + (2 * jx + (c4 & 1)) * LATP + 2 * i + (c4 >> 1);
            const float u = vb[tl * 4 + c4], c = vb[(64 + tl) * 4 + c4];
            const float pred = u + p.cfg * (c - u);
            if (p.out != nullptr) p.out[xi] = pred;
            const float xo = p.x_upd[xi];
            float xn;
            if (p.out_mode == OUT_RF) {
                xn = xo + pred * p.c1;
            } else {
                const float mean = p.c1 * (xo - p.c2 * pred);
                xn = mean + p.c3 * p.noise[xi];
            }
            p.x_upd[xi] = xn;
        }
    }
}

// =================================================================================== attention
// softmax(q k^T / sqrt(32)) v for one (sequence, head) per CTA; 15 warps x 32 query rows.
// K/V/Q arrive with three bulk async copies (the global layout is already the swizzled smem image).
// S and P never leave registers: the S accumulator fragments are re-packed as the A operand of P.V.
constexpr int ATT_THREADS = 480;
constexpr int ATT_SMEM_Q = 0, ATT_SMEM_K = NTOK * HD * 2, ATT_SMEM_V = 2 * NTOK * HD * 2;
constexpr int ATT_SMEM_BAR = 3 * NTOK * HD * 2;
constexpr int ATT_SMEM_BYTES = ATT_SMEM_BAR + 32;
constexpr int ATT_KC = 32;          // keys per online-softmax chunk

__global__ void __launch_bounds__(ATT_THREADS, 1) attn_kernel(const __half* __restrict__ qkv, __half* __restrict__ o) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int seq = blockIdx.x >> 2, head = blockIdx.x & 3;
    const uint32_t sb = smem_u32(smem);
    const uint32_t bar = sb + ATT_SMEM_BAR;
    constexpr uint32_t PIECE = NTOK * HD * 2;   // 30720 B
    if (tid == 0) {
        mbar_init(bar, 1); mbar_init(bar + 8, 1); mbar_init(bar + 16, 1);
        mbar_fence_init();
        const char* src = reinterpret_cast<const char*>(qkv) + (size_t)blockIdx.x * 3 * PIECE;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            mbar_expect_tx(bar + 8 * i, PIECE);
            bulk_g2s(sb + i * PIECE, src + (size_t)i * PIECE, PIECE, bar + 8 * i);
        }
    }
    __syncthreads();

    // Q fragments: 2 m-tiles x 2 k-steps
    const int qrow0 = warp * 32;
    uint32_t qf[2][2][4];
    mbar_wait(bar, 0);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const int row = qrow0 + mt * 16 + (lane & 15);
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
            ldmatrix_x4(qf[mt][kk][0], qf[mt][kk][1], qf[mt][kk][2], qf[mt][kk][3],
                        sb + ATT_SMEM_Q + row * 64 + (((2 * kk + (lane >> 4)) ^ ((row >> 1) & 3)) << 4));
    }
    float oacc[2][4][4];
    float mx[2][2], ls[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        mx[mt][0] = mx[mt][1] = -INFINITY;
        ls[mt][0] = ls[mt][1] = 0.f;
#pragma unroll
        for (int d = 0; d < 4; ++d) oacc[mt][d][0] = oacc[mt][d][1] = oacc[mt][d][2] = oacc[mt][d][3] = 0.f;
    }
    const float sc = 0.25503486f;   // log2(e) / sqrt(32)
    mbar_wait(bar + 8, 0);
    mbar_wait(bar + 16, 0);

#pragma unroll 1
    for (int c0 = 0; c0 < NTOK; c0 += ATT_KC) {
        float s[2][4][4];
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
            const int key = c0 + nb * 8 + (lane & 7);
            uint32_t k0, k1, k2, k3;
            ldmatrix_x4(k0, k1, k2, k3, sb + ATT_SMEM_K + key * 64 + (((lane >> 3) ^ ((key >> 1) & 3)) << 4));
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                s[mt][nb][0] = s[mt][nb][1] = s[mt][nb][2] = s[mt][nb][3] = 0.f;
                mma_f16(s[mt][nb], qf[mt][0][0], qf[mt][0][1], qf[mt][0][2], qf[mt][0][3], k0, k1);
                mma_f16(s[mt][nb], qf[mt][1][0], qf[mt][1][1], qf[mt][1][2], qf[mt][1][3], k2, k3);
            }
        }
        uint32_t pf[2][2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            float a0 = fmaxf(fmaxf(s[mt][0][0], s[mt][0][1]), fmaxf(s[mt][1][0], s[mt][1][1]));
            float a1 = fmaxf(fmaxf(s[mt][0][2], s[mt][0][3]), fmaxf(s[mt][1][2], s[mt][1][3]));
            a0 = fmaxf(a0, fmaxf(fmaxf(s[mt][2][0], s[mt][2][1]), fmaxf(s[mt][3][0], s[mt][3][1])));
            a1 = fmaxf(a1, fmaxf(fmaxf(s[mt][2][2], s[mt][2][3]), fmaxf(s[mt][3][2], s[mt][3][3])));
            const float n0 = fmaxf(mx[mt][0], quad_max(a0)), n1 = fmaxf(mx[mt][1], quad_max(a1));
            const float cr0 = ex2_approx((mx[mt][0] - n0) * sc), cr1 = ex2_approx((mx[mt][1] - n1) * sc);
            mx[mt][0] = n0; mx[mt][1] = n1;
            const float b0 = -n0 * sc, b1 = -n1 * sc;
            float l0 = ls[mt][0] * cr0, l1 = ls[mt][1] * cr1;
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                oacc[mt][d][0] *= cr0; oacc[mt][d][1] *= cr0; oacc[mt][d][2] *= cr1; oacc[mt][d][3] *= cr1;
            }
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                const float p0 = ex2_approx(fmaf(s[mt][nb][0], sc, b0)), p1 = ex2_approx(fmaf(s[mt][nb][1], sc, b0));
                const float p2 = ex2_approx(fmaf(s[mt][nb][2], sc, b1)), p3 = ex2_approx(fmaf(s[mt][nb][3], sc, b1));
                l0 += p0 + p1; l1 += p2 + p3;
                pf[mt][nb >> 1][(nb & 1) * 2 + 0] = pack_h2(p0, p1);
                pf[mt][nb >> 1][(nb & 1) * 2 + 1] = pack_h2(p2, p3);
            }
            ls[mt][0] = l0; ls[mt][1] = l1;
        }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const int key = c0 + ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
#pragma unroll
            for (int dc = 0; dc < 4; dc += 2) {
                uint32_t v0, v1, v2, v3;
                ldmatrix_x4_trans(v0, v1, v2, v3, sb + ATT_SMEM_V + key * 64 + (((dc + (lane >> 4)) ^ ((key >> 1) & 3)) << 4));
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    mma_f16(oacc[mt][dc], pf[mt][ks][0], pf[mt][ks][1], pf[mt][ks][2], pf[mt][ks][3], v0, v1);
                    mma_f16(oacc[mt][dc + 1], pf[mt][ks][0], pf[mt][ks][1], pf[mt][ks][2], pf[mt][ks][3], v2, v3);
                }
            }
        }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const float i0 = 1.f / quad_sum(ls[mt][0]), i1 = 1.f / quad_sum(ls[mt][1]);
        const int r0 = qrow0 + mt * 16 + g;
        __half* dst = o + ((size_t)seq * NTOK + r0) * D + head * HD + 2 * t;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            *reinterpret_cast<uint32_t*>(dst + d * 8) = pack_h2(oacc[mt][d][0] * i0, oacc[mt][d][1] * i0);
            *reinterpret_cast<uint32_t*>(dst + 8 * D + d * 8) = pack_h2(oacc[mt][d][2] * i1, oacc[mt][d][3] * i1);
        }
    }
}

}  // namespace t2s
