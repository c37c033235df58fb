// T2S-DiT denoiser kernels for sm_100a (reference: model/denoiser/transformer.py:94-193 and the
// timm Attention / Mlp it calls).  Arithmetic: fp16 operands (10-bit mantissa, = tf32 operand
// precision) with fp32 accumulation on the tensor cores; residual stream, LayerNorm, softmax
// statistics, modulation and the sampler update in fp32.
//
// Kernels (one sampling step = cond + embed_qkv + 4 x (attention + token)):
//   cond_kernel       time embedding + text conditioning + SiLU + adaLN Linear for all 4 blocks
//   token_kernel<EMBED>  patch-embed + pos  -> h ; LN1+modulate -> QKV(l=0)
//   attn_kernel       softmax(q k^T / sqrt(32)) v per (sequence, head): tcgen05 QK^T / PV, thread-per-row softmax
//   token_kernel<MID>    proj+gate+residual, LN2+modulate, fc1+GELU, fc2+gate+residual -> h ;
//                        LN1+modulate -> QKV(l+1)
//   token_kernel<FINAL>  ... + final LN + Linear(128->4) + unpatchify + CFG mix + Euler / DDPM update
#pragma once
#include "common.cuh"

namespace t2s {

struct DitWeights {                 // device pointers; mirrors t2s_dit_weights (include/t2s_b200.h)
    // weight stages of WSTAGE_BYTES: [128 n][128 k] fp16 tcgen05 operand image (32 KB) + the Linear's bias as a [128 n][16 k]
    // operand block (4 KB: k = 0 / 1 hold the fp16 high / low parts of the fp32 bias, the rest is zero), see tc_gemm
    const __half* w_qkv[NLAYER];    // 3 stages: q | k | v
    const __half* w_post[NLAYER];   // 5 stages: proj | fc1 rows 0..127 | fc1 rows 128..255 | fc2 k 0..127 | fc2 k 128..255 (zero bias block)
    const float* b_qkv[NLAYER];     // [384]
    const float* b_proj[NLAYER];    // [128]
    const float* b_fc1[NLAYER];     // [256]
    const float* b_fc2[NLAYER];     // [128]
    const float* w_ada_t;           // [4][128 k][768 o]  (adaLN Linear weight, transposed)
    const float* b_ada;             // [4][768]
    const float* w_embed;           // [4][128]   (patch_emb.weight @ conv.weight)^T  (folded, pixel-major)
    const float* b_embed;           // [128]      patch_emb.weight @ conv.bias + patch_emb.bias
    const float* pos;               // [tiles per pair][32 col chunks][64 rows][4]  pos_embed in the residual tile layout
    const float* w_final;           // [4][128]   linear_emb_to_patch.weight * ln.weight
    const float* b_final;           // [4]        linear_emb_to_patch.weight @ ln.bias + bias
    const float* freqs;             // [64]       10000 ** linspace(0,1,64)
    int latent_h;                   // latent width H: 0 / 30 (T2S), 50 or 64 (fork); selects the DitShape instantiation
    // the same fp16 weights as 16 KB HALF stages [64 n][128 k] (K-chunk stride 1024 B) for the fused step kernel, or NULL:
    const __half* w_qkv_half[NLAYER];   // 6 halves:  q n 0..63 | q n 64..127 | k .. | k .. | v .. | v ..
    const __half* w_post_half[NLAYER];  // 10 halves: proj h0 h1 | fc1[0:128] h0 h1 | fc1[128:256] h0 h1 | fc2 (k0,h0) (k1,h0) (k0,h1) (k1,h1)
};

enum TokenMode { TOK_EMBED = 0, TOK_MID = 1, TOK_FINAL = 2 };
enum OutMode { OUT_FWD = 0, OUT_RF = 1, OUT_DDPM = 2 };

struct TokArgs {
    DitWeights w;
    const float* x;        // latents [(nseq >> x_shift)][64][H]
    int x_shift;           // 1 when the two sequences of a pair share one latent (CFG), else 0
    float* h;              // residual stream, tiled: [npair][tiles per pair][32 col chunks][128 rows][4] fp32
    __half* qkv;           // [nseq][4 heads] x {Q, K, V tcgen05 operand images} fp16 (see attn_kernel)
    const __half* o;       // attention output, tiled A-operand images: [npair][tiles per pair][16 K chunks][16 row groups][8][8] fp16
    const float* mod;      // adaLN modulation [nseq][4][768] fp32
    int nseq;
    int layer;             // block whose post-attention half runs here (MID / FINAL)
    // FINAL only
    int out_mode;
    float* out;            // OUT_FWD: [nseq][64][H]; OUT_RF / OUT_DDPM: optional guided prediction [npair][64][H]
    float* x_upd;          // OUT_RF / OUT_DDPM: latent updated in place [npair][64][H]
    const float* noise;    // OUT_DDPM: [npair][64][H] for this step, or NULL = in-kernel Philox noise (seed, step)
    unsigned long long seed;
    unsigned int step;
    float cfg, c1, c2, c3; // RF: x += pred*c1 ; DDPM: x = c1*(x - c2*pred) + c3*noise
    long long* trace;      // optional phase trace [grid][32] of clock64 stamps (tile 0, row 0), NULL = off
    // The residual stream entering block 0 is a 4-pixel affine map of the latent plus pos_embed: the fused loops do not
    // materialise it.  EMBED skips its store (skip_h_store) and the MID kernel of block 0 recomputes its rows from x with the
    // same instruction sequence (recompute_h0; needs x, x_shift): 240 KB per sequence less written and read per step.
    int skip_h_store, recompute_h0;
    // guided sampling loops: the unconditional branch's modulation depends on the step only, so cond_kernel writes ONE row
    // (sequence 0) instead of one per sample and every pair reads that row for its even (unconditional) sequence
    int uncond_shared;
};

// =================================================================================== cond
// The modulation table holds, per (sequence, block), shift_msa | 1 + scale_msa | gate_msa | shift_mlp | 1 + scale_mlp |
// gate_mlp: the token kernel's normalise-modulate is then two fused multiply-adds per element.
__device__ __forceinline__ bool mod_is_scale(int col) { return (col >> 7) == 1 || (col >> 7) == 4; }
// mod[seq][l][:] = Linear_l( SiLU( temb(t) (+ text) ) )      transformer.py:30-40,106-109,115,174-178
// grid (ceil(nseq/8), 4), block 256
// uncond_shared (guided loops, cfg_pairs): the unconditional modulation row depends on the step only: CTA x = gridDim.x - 1
// computes it once (sequence 0), the others 8 CONDITIONAL (odd) sequences each; grid (ceil(npair/8) + 1, 4).
__device__ __forceinline__ int cond_seq_of(int slot, int uncond_shared) {
    if (!uncond_shared) return blockIdx.x * 8 + slot;
    if (blockIdx.x == gridDim.x - 1) return slot == 0 ? 0 : 0x3fffffff;
    return 2 * (blockIdx.x * 8 + slot) + 1;
}
__global__ void __launch_bounds__(256) cond_kernel(float* __restrict__ mod, const float* __restrict__ t100, int t_stride,
                                                   const float* __restrict__ emb, int emb_shift, int cfg_pairs,
                                                   const float* __restrict__ freqs, const float* __restrict__ w_ada_t,
                                                   const float* __restrict__ b_ada, int nseq, int uncond_shared) {
    __shared__ float sc[8][D];
    const int l = blockIdx.y, tid = threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();                  // the previous step's kernels still read the modulation table this kernel overwrites
    for (int i = tid; i < 8 * D; i += 256) {
        const int si = i >> 7, f = i & 127, seq = cond_seq_of(si, uncond_shared);
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
        const int seq = cond_seq_of(s, uncond_shared);
        if (seq < nseq) {
            float* dst = mod + ((size_t)seq * NLAYER + l) * MOD;
#pragma unroll
            for (int a = 0; a < 3; ++a)                      // the scale slices hold 1 + scale (see ln_mod_store)
                dst[a * 256 + tid] = acc[a][s] + (b_ada[l * MOD + a * 256 + tid] + (mod_is_scale(a * 256 + tid) ? 1.f : 0.f));
        }
    }
}

// ----- small batches: the same computation with the 768 outputs of a block split over three CTAs
// grid (ceil(nseq/8), 4 blocks x 3 column groups), block 256: a thread owns one of the 768 outputs of one block for 8
// sequences (8 weight loads in flight).  Three times the CTAs and a third of the serial work per thread: at batch 1 the
// one-CTA-per-block form was a quarter of a sampling step; at large batches the form above is faster (the prologue is
// not repeated per column group).
__global__ void __launch_bounds__(256) cond_split_kernel(float* __restrict__ mod, const float* __restrict__ t100, int t_stride,
                                                   const float* __restrict__ emb, int emb_shift, int cfg_pairs,
                                                   const float* __restrict__ freqs, const float* __restrict__ w_ada_t,
                                                   const float* __restrict__ b_ada, int nseq) {
    __shared__ float sc[8][D];
    const int s0 = blockIdx.x * 8, l = blockIdx.y / 3, col = (blockIdx.y % 3) * 256 + threadIdx.x, tid = threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
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
    float acc[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) acc[s] = 0.f;
    const float* w = w_ada_t + (size_t)l * D * MOD + col;
#pragma unroll 8
    for (int k = 0; k < D; ++k) {
        const float wv = w[k * MOD];
#pragma unroll
        for (int s = 0; s < 8; ++s) acc[s] = fmaf(wv, sc[s][k], acc[s]);
    }
    const float bv = b_ada[l * MOD + col] + (mod_is_scale(col) ? 1.f : 0.f);
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        const int seq = s0 + s;
        if (seq < nseq) mod[((size_t)seq * NLAYER + l) * MOD + col] = acc[s] + bv;
    }
}

// =================================================================================== token block (tcgen05)
// One CTA = TWO consecutive pair tiles (each 128 rows = 60 tokens x {sequence 2p, sequence 2p+1} + 8 padding
// rows) that march through the same weight stages in lock-step, so every 32 KB weight stage fetched from L2
// feeds two 128x128x128 GEMM chunks.  Everything that is local to a token runs here for one DiT block boundary:
//   MID  (block l):  x += gate_msa*proj(o) ; a2 = mod(LN2(x)) ; x += gate_mlp*fc2(GELU(fc1(a2))) ; store x ;
//                    a' = mod(LN1(x)) of block l+1 ; q|k|v = a' Wqkv^T + b   (transformer.py:114-117, timm Attention/Mlp)
//   EMBED:           x = patch-embed + pos ; a' of block 0 ; q|k|v
//   FINAL (block 3): ... ; final LN + Linear(128->4) + unpatchify + CFG mix + Euler/DDPM update
// Warp roles (576 threads): warps 0-7 / 8-15 = epilogue of tile 0 / 1, TWO threads per tile row (= TMEM lane), each
// owning 64 of the 128 columns of the current region (warp w: lanes 32 (w % 4).., column half (w / 4) % 2);
// warp 16 = producer (bulk async copies: weight stages, attention-output
// tiles, per-tile vectors; completion on mbarriers); warp 17 = MMA issuer (one lane issues
// tcgen05.mma.kind::f16 128x128x16, accumulators in TMEM, completion via tcgen05.commit).
// Each tile owns two 128-column TMEM regions X, Y:
//   proj->X (the updated residual is parked in X until the MLP branch adds to it: it never leaves the SM)
//   fc1[0:128]->Y | fc1[128:256]->Y | fc2 (two K halves)->Y (each after the previous occupant has been drained)
//   q->X | k->Y | v->X
// While tile 0's epilogue warps work on a chunk, the tensor pipe runs tile 1's chunk and vice versa.
// Persistent: grid = min(#work items, #SMs), a work item = two consecutive pair tiles; every per-tile barrier
// completes once per item (wait parity = item iteration & 1), the weight ring and the two vector buffers run on
// their own counters.  While an item is in its second half the producer already fetches the next item's vectors,
// attention-output tiles (into the HA buffers, free once fc2's first K half has been consumed) and L2-prefetches its
// residual tiles, so the next item starts with everything on the SM.
#ifndef T2S_TOK_MMA_WARPS
#define T2S_TOK_MMA_WARPS 1      // MMA issuer warps of the token kernel: 1 = one in-order issuer for both tiles (default), 2 = one per tile (A/B build)
#endif
constexpr int TC_MMA_WARPS = T2S_TOK_MMA_WARPS;
constexpr int TC_THREADS = 544 + 32 * TC_MMA_WARPS;               // 16 epilogue warps + producer + MMA issuer(s)
constexpr int TC_NSTAGE = 2;
constexpr int WBIAS_BYTES = 4096;                                // bias operand block [128 n][16 k] fp16 behind every weight image
constexpr int WSTAGE_BYTES = STAGE_BYTES + WBIAS_BYTES;          // one weight stage in global memory and in the ring
constexpr int TC_SM_A = 0;                                       // [2 tiles] 32 KB A operand: o tile / a2 / hidden-b / a'
constexpr int TC_SM_HA = 2 * STAGE_BYTES;                        // [2 tiles] 32 KB A operand: hidden-a
constexpr int TC_SM_W = 4 * STAGE_BYTES;                         // weight ring
constexpr int TC_SM_VEC = TC_SM_W + TC_NSTAGE * WSTAGE_BYTES;    // [2 buffers] per-pair vectors (fp32), shared by both tiles
constexpr int V_MOD = 0;        // [2 branches][768]  adaLN chunk of block l
constexpr int V_MODN = 1536;    // [2][256]           shift_msa | scale_msa of the next block
constexpr int V_WEMB = 0;       // [4][128]   EMBED only: aliases the V_MOD area it does not use
constexpr int V_BEMB = 512;     // [128]
constexpr int V_WFIN = 1536;    // [4][128]   FINAL only: aliases the V_MODN area it does not use
constexpr int V_BFIN = 2048;    // [4]        FINAL only
constexpr int V_FLOATS = 2064;                                   // one vector buffer (the Linear biases travel inside the weight stages)
constexpr int TC_SM_ONES = TC_SM_VEC + 2 * V_FLOATS * 4;         // [128 rows][16 k] fp16 A operand block: k = 0, 1 are 1.0, the rest 0 (bias MMA)
constexpr int TC_SM_ST = TC_SM_ONES + WBIAS_BYTES;               // [2 tiles][2 halves][128] float2 LayerNorm statistics exchange
constexpr int TC_SM_BAR = TC_SM_ST + 2 * 2 * TILE_ROWS * 8;
constexpr int TC_SM_TMEM = TC_SM_BAR + 40 * 8;
constexpr int TC_SM_EMB = TC_SM_TMEM + 16;                        // [4][128] folded patch-embed weight + [128] bias (recompute_h0)
constexpr int TOK_SMEM_BYTES = TC_SM_EMB + 5 * D * 4;
// FINAL's [2 tiles][128][4] fp32 projection exchange lives in the tile's A operand buffer: its last reader (fc2's second K
// half) has completed before the final pass starts, and its next writer (the LN-modulate of the tile's next item) comes after
// a 256-thread barrier (merge_stats) that every reader of the exchange has to reach first
static_assert(TOK_SMEM_BYTES <= 232448, "token kernel shared memory exceeds 227 KB");
// barrier ids
enum { B_WFULL = 0, B_WEMPTY = 2, B_VFULL = 4, B_VFREE = 6, B_TILE = 8 };
enum { T_OFULL = 0, T_A2 = 1, T_HA = 2, T_HB = 3, T_A3 = 4, T_XFREE = 5, T_DONE = 6, T_HAFREE = 7, T_ACC = 8 /* ..14 */, T_YFREE = 15, T_COUNT = 16 };
static_assert(B_TILE + 2 * T_COUNT <= 40, "barrier slots");
constexpr uint32_t TC_IDESC = umma_idesc_f16(128, 128);
constexpr uint32_t KCH = 2048;   // byte stride between K chunks (16 row groups x 128 B) in a [128][128] operand image

// ---- timing knock-outs (A/B builds only, WRONG results): which resource paces the token kernel?  (tools/build_variant.py,
// tools/ab_tok.sh; measured in round 2, MID at 2048 sequences, base 0.489 ms: profiles/r02_token_knockouts.log)
//   -DT2S_KO_LDS         the per-column constants (gate / scale / shift) come from a register instead of shared memory   -1.6 %
//   -DT2S_KO_MUFU        GELU without the tanh                                                                           -1.6 %
//   -DT2S_KO_STG         no global stores (q|k|v images, residual tile)                                                  -15 %
//   -DT2S_KO_HLD         the residual tile is not loaded                                                                 -8 %
//   -DT2S_EXP_SKIP_WLOAD weight stages fetched from L2 for a CTA's first item only                                       -4 %
//   -DT2S_KO_MMA         one of the eight MMAs of a chunk (sustained-clock / power experiment, tools/power_by_kernel.py)
// i.e. what is left of the kernel's time is spread over everything; the path to L2 / HBM (2.2 GB per launch) is the largest share.
// ko_never is false at run time but unknown to the compiler (g_ko_zero is never written), so the arithmetic feeding a knocked-out
// store survives
__device__ int g_ko_zero;
__device__ __forceinline__ bool ko_never() { return *reinterpret_cast<volatile int*>(&g_ko_zero) != 0; }
#ifdef T2S_KO_LDS
#define KO_LD4(ptr) make_float4(__int_as_float(0x3f800000 + (int)(size_t)(ptr)), 1.f, 1.f, 1.f)
#else
#define KO_LD4(ptr) (*reinterpret_cast<const float4*>(ptr))
#endif

// global accesses of the epilogues: plain, or (-DT2S_TOK_CS, A/B build) with the streaming cache hint
__device__ __forceinline__ void stg16(void* p, uint4 v) {
#ifdef T2S_TOK_CS
    __stcs(reinterpret_cast<uint4*>(p), v);
#else
    *reinterpret_cast<uint4*>(p) = v;
#endif
}
__device__ __forceinline__ void stg16f(float* p, float4 v) {
#ifdef T2S_TOK_CS
    __stcs(reinterpret_cast<float4*>(p), v);
#else
    *reinterpret_cast<float4*>(p) = v;
#endif
}
__device__ __forceinline__ float4 ldg16f(const float* p) {
#ifdef T2S_TOK_CS
    return __ldcs(reinterpret_cast<const float4*>(p));
#else
    return *reinterpret_cast<const float4*>(p);
#endif
}

// one 128x128x128 GEMM chunk: 8 x tcgen05.mma (K = 16 each); operands in the canonical no-swizzle K-major image
// Called by the whole (converged) MMA warp so that descriptors live in uniform registers; only `lead` issues.
// ones_smem != 0: a ninth MMA adds the Linear's bias: A = the constant block whose k = 0, 1 columns are 1.0 (every row), B =
// the stage's bias block (k = 0 / 1: fp16 high / low part of the fp32 bias) — the epilogues then neither load nor add a bias
// (the per-column constants every row thread has to load through the shared-memory crossbar are what paces the passes).
__device__ __forceinline__ void tc_gemm(uint32_t a_smem, uint32_t w_smem, uint32_t d_tmem, bool accumulate, bool lead, uint32_t ones_smem = 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint64_t ad = umma_desc(a_smem + k * 2 * KCH, KCH, 128), bd = umma_desc(w_smem + k * 2 * KCH, KCH, 128);
#ifdef T2S_KO_MMA              // timing / power knock-out: only the first of the eight MMAs of a chunk is issued
        if (lead && k == 0) umma_f16(d_tmem, ad, bd, TC_IDESC, accumulate ? 1u : 0u);
#else
        if (lead) umma_f16(d_tmem, ad, bd, TC_IDESC, (accumulate || k > 0) ? 1u : 0u);
#endif
    }
    if (ones_smem != 0) {
        const uint64_t ad = umma_desc(ones_smem, KCH, 128), bd = umma_desc(w_smem + STAGE_BYTES, KCH, 128);
        if (lead) umma_f16(d_tmem, ad, bd, TC_IDESC, 1u);
    }
}

// ---- epilogue building blocks.  TWO threads own one tile row (= one TMEM lane): thread (r, hh) works on the 64
// columns [64 hh, 64 hh + 64) of whatever 128-column region is current, in 16-column blocks inside rolled loops
// (compact code, no large register arrays).  The updated residual row is written back over the consumed accumulator
// in TMEM, which serves as the row buffer for the LayerNorm's second pass.  LayerNorm statistics of the two halves
// are merged through shared memory (Chan's parallel-variance formula) behind a 256-thread named barrier.
struct RowStats { float mean, rstd; };
// per-column constants (biases): from shared memory (token_kernel stages them per item) or, GB = true, straight from global
// memory through the read-only path (the fused step kernel: a warp-uniform address, one L1 sector per load)
template <bool GB>
__device__ __forceinline__ float4 ldv4(const float* p) {
    if constexpr (GB) return __ldg(reinterpret_cast<const float4*>(p));
    else return *reinterpret_cast<const float4*>(p);
}
struct HalfStats { float mean, m2; };

// pipelined walk over NB 16-column blocks of a TMEM region (NB even)
template <int NB, class F>
__device__ __forceinline__ void for_blocks16(uint32_t taddr, F&& body) {
    float a[16], b[16];
    tmem_ld16(taddr, a);
#pragma unroll 1
    for (int cb = 0; cb < NB; cb += 2) {
        tmem_wait_ld();
        tmem_ld16(taddr + (cb + 1) * 16, b);
        body(cb, a);
        tmem_wait_ld();
        if (cb + 2 < NB) tmem_ld16(taddr + (cb + 2) * 16, a);
        body(cb + 1, b);
    }
}


// pass 1 over this thread's 64 columns: hn = hin + gate * (acc + bias), written back over the accumulator (TMEM),
// returning the half-row mean and centred sum of squares.  All pointers / addresses are already offset to the
// thread's column half.
//   resid_pass_regs: hin was prefetched from the residual tile into registers (before the accumulator wait);
//   resid_pass_tmem: hin is the row parked in another TMEM region (the residual after the attention branch never
//                    leaves the SM); optionally stores hn to the residual tile in global memory
//                    (element (col chunk c4, this row) at +c4*TILE_ROWS*4 floats).
__device__ __forceinline__ void block_stats(const float (&a)[16], float shift, float& sum, float& sq) {
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
    const float ns = -shift;
#pragma unroll
    for (int j = 0; j < 16; j += 2) {                       // packed fp32: three issue slots per pair
        float d0, d1;
        add2(d0, d1, a[j], a[j + 1], ns, ns);
        add2(s0, s1, s0, s1, d0, d1);
        fma2(q0, q1, d0, d1, d0, d1, q0, q1);
    }
    sum += s0 + s1; sq += q0 + q1;
}
__device__ __forceinline__ HalfStats half_stats(float shift, float sum, float sq) {
    HalfStats st;
    st.mean = shift + sum * (1.f / 64);
    st.m2 = fmaxf(sq - sum * sum * (1.f / 64), 0.f);
    return st;
}
// HASB = false: the GEMM already added the bias (token_kernel folds every Linear bias into its GEMM, see tc_gemm)
template <bool GB = false, bool HASB = true>
__device__ __forceinline__ HalfStats resid_pass_regs(uint32_t tacc, const float* __restrict__ gate, const float* __restrict__ bias,
                                                     const float4 (&hq)[16]) {
    float sum = 0.f, sq = 0.f, shift = 0.f;
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) {
        float a[16];
        tmem_ld16(tacc + cb * 16, a);
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 g4 = KO_LD4(gate + cb * 16 + q * 4);
            float t0 = a[q * 4 + 0], t1 = a[q * 4 + 1], t2 = a[q * 4 + 2], t3 = a[q * 4 + 3];
            if constexpr (HASB) {
                const float4 b4 = ldv4<GB>(bias + cb * 16 + q * 4);
                add2(t0, t1, t0, t1, b4.x, b4.y);
                add2(t2, t3, t2, t3, b4.z, b4.w);
            }
            fma2(a[q * 4 + 0], a[q * 4 + 1], g4.x, g4.y, t0, t1, hq[cb * 4 + q].x, hq[cb * 4 + q].y);
            fma2(a[q * 4 + 2], a[q * 4 + 3], g4.z, g4.w, t2, t3, hq[cb * 4 + q].z, hq[cb * 4 + q].w);
        }
        if (cb == 0) shift = a[0];
        block_stats(a, shift, sum, sq);
        tmem_st16(tacc + cb * 16, a);
    }
    tmem_wait_st();
    return half_stats(shift, sum, sq);
}
template <bool STORE, bool GB = false, bool HASB = true>
__device__ __forceinline__ HalfStats resid_pass_tmem(uint32_t tacc, uint32_t thin, const float* __restrict__ gate, const float* __restrict__ bias,
                                                     float* __restrict__ hdst, bool valid) {
    float sum = 0.f, sq = 0.f, shift = 0.f;
    float a[16], h[16];
    tmem_ld16(tacc, a);
    tmem_ld16(thin, h);
#pragma unroll 1
    for (int cb = 0; cb < 4; ++cb) {
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 g4 = KO_LD4(gate + cb * 16 + q * 4);
            float t0 = a[q * 4 + 0], t1 = a[q * 4 + 1], t2 = a[q * 4 + 2], t3 = a[q * 4 + 3];
            if constexpr (HASB) {
                const float4 b4 = ldv4<GB>(bias + cb * 16 + q * 4);
                add2(t0, t1, t0, t1, b4.x, b4.y);
                add2(t2, t3, t2, t3, b4.z, b4.w);
            }
            fma2(a[q * 4 + 0], a[q * 4 + 1], g4.x, g4.y, t0, t1, h[q * 4 + 0], h[q * 4 + 1]);
            fma2(a[q * 4 + 2], a[q * 4 + 3], g4.z, g4.w, t2, t3, h[q * 4 + 2], h[q * 4 + 3]);
        }
        if (cb < 3) tmem_ld16(thin + (cb + 1) * 16, h);
        if (cb == 0) shift = a[0];
        block_stats(a, shift, sum, sq);
        tmem_st16(tacc + cb * 16, a);
        if (STORE && valid) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                *reinterpret_cast<float4*>(hdst + (cb * 4 + q) * TILE_ROWS * 4) = make_float4(a[q * 4], a[q * 4 + 1], a[q * 4 + 2], a[q * 4 + 3]);
        }
        if (cb < 3) tmem_ld16(tacc + (cb + 1) * 16, a);
    }
    tmem_wait_st();
    return half_stats(shift, sum, sq);
}

// merge the two half-row statistics of a row: slot = exchange buffer [2 halves][128 rows] float2 of this tile
__device__ __forceinline__ RowStats merge_stats(HalfStats hs, float2* slot, int r, int hh, int bar_id, float eps) {
    slot[hh * TILE_ROWS + r] = make_float2(hs.mean, hs.m2);
    asm volatile("bar.sync %0, 256;\n" :: "r"(bar_id) : "memory");
    const float2 o = slot[(hh ^ 1) * TILE_ROWS + r];
    const float dm = hs.mean - o.x;
    RowStats st;
    st.mean = 0.5f * (hs.mean + o.x);
    st.rstd = rsqrtf((hs.m2 + o.y + 32.f * dm * dm) * (1.f / D) + eps);
    return st;
}

// pass 2 over this thread's 64 columns: LayerNorm (no affine) + modulate x*(1+scale)+shift (transformer.py:7-8,102-103),
// packed to fp16 and stored as the next GEMM's A operand image.  kc0 = first 8-column K chunk of the half (8 hh).
__device__ __forceinline__ void ln_mod_store(uint32_t trow, RowStats st, const float* __restrict__ shift, const float* __restrict__ scale1,
                                             uint8_t* abuf, int r, int kc0) {
    const float rs = st.rstd, nm = -st.mean * st.rstd;
    for_blocks16<4>(trow, [&](int cb, float (&a)[16]) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 sc = KO_LD4(scale1 + cb * 16 + q * 4);   // 1 + scale
            const float4 sh = KO_LD4(shift + cb * 16 + q * 4);
            // ((a - mean) * rstd) * (1 + scale) + shift as two packed fused multiply-adds per pair
            float n0, n1, n2, n3;
            fma2(n0, n1, a[q * 4 + 0], a[q * 4 + 1], rs, rs, nm, nm);
            fma2(n2, n3, a[q * 4 + 2], a[q * 4 + 3], rs, rs, nm, nm);
            fma2(a[q * 4 + 0], a[q * 4 + 1], n0, n1, sc.x, sc.y, sh.x, sh.y);
            fma2(a[q * 4 + 2], a[q * 4 + 3], n2, n3, sc.z, sc.w, sh.z, sh.w);
        }
#pragma unroll
        for (int c8 = 0; c8 < 2; ++c8)
            *reinterpret_cast<uint4*>(abuf + (kc0 + cb * 2 + c8) * KCH + r * 16) =
                make_uint4(pack_h2(a[c8 * 8 + 0], a[c8 * 8 + 1]), pack_h2(a[c8 * 8 + 2], a[c8 * 8 + 3]),
                           pack_h2(a[c8 * 8 + 4], a[c8 * 8 + 5]), pack_h2(a[c8 * 8 + 6], a[c8 * 8 + 7]));
    });
}

// hidden = GELU_tanh(acc + b1) over this thread's 64 columns, packed to fp16 into an A operand image (timm Mlp,
// transformer.py:99,105)
template <bool GB = false, bool HASB = true>
__device__ __forceinline__ void gelu_store(uint32_t taddr, const float* __restrict__ bias, uint8_t* abuf, int r, int kc0) {
    for_blocks16<4>(taddr, [&](int cb, float (&v)[16]) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float x0 = v[q * 4 + 0], x1 = v[q * 4 + 1], x2 = v[q * 4 + 2], x3 = v[q * 4 + 3];
            if constexpr (HASB) {
                const float4 b4 = ldv4<GB>(bias + cb * 16 + q * 4);
                add2(x0, x1, x0, x1, b4.x, b4.y);
                add2(x2, x3, x2, x3, b4.z, b4.w);
            }
            gelu_tanh2(v[q * 4 + 0], v[q * 4 + 1], x0, x1);
            gelu_tanh2(v[q * 4 + 2], v[q * 4 + 3], x2, x3);
        }
#pragma unroll
        for (int c8 = 0; c8 < 2; ++c8)
            *reinterpret_cast<uint4*>(abuf + (kc0 + cb * 2 + c8) * KCH + r * 16) =
                make_uint4(pack_h2(v[c8 * 8 + 0], v[c8 * 8 + 1]), pack_h2(v[c8 * 8 + 2], v[c8 * 8 + 3]),
                           pack_h2(v[c8 * 8 + 4], v[c8 * 8 + 5]), pack_h2(v[c8 * 8 + 6], v[c8 * 8 + 7]));
    });
}

// ---- quad-layout epilogue passes: MEASURED, NOT ADOPTED (run by tools/probe_pass.cu only; token_kernel uses the row-per-thread
// passes above).  LN pass 1729 -> 1573 clk for both tiles (2478 -> 1838 beside a tensor-core stream): the 4-byte stores and 8-byte
// loads cost as many crossbar slots as the 16-byte ones they replace, and once the biases had moved into the GEMMs the constant
// loads were 1.6 % of the kernel (knock-out map above).  Kept as the reference for the 16x256b TMEM shape.
// The row-per-thread passes make every thread load every per-column constant
// of its 64 columns (scale, shift, gate, bias: warp-uniform 16-byte LDS, 3.3 KB per thread and work item): measured with
// tools/probe_pass.cu, those broadcast loads alone cost 2.7-3.6 clk of the SM's shared-memory crossbar each and, together with
// the tensor core's operand reads from the same crossbar, they — not issue slots, TMEM or HBM — paced every pass.  These passes
// read the accumulators through tcgen05.ld.16x256b instead (mapping verified by tools/probe_tmem_shapes.cu): thread T of
// a warp then holds FOUR rows x 16 columns of its 32-lane x 64-column block,
//   rows    R(rho) = 32 q + 8 rho + (T >> 2),  rho = 2 h + rr = 0..3      (q = warp & 3: TMEM lane quarter)
//   columns c(g,e) = 64 hh + 8 g + 2 (T & 3) + e,  g = 0..7, e = 0..1     (hh: column half of the warp)
// so a per-column constant is loaded once for four rows (a quarter of the bytes, 8-byte LDS), the packed fp32 operations pair
// the two adjacent columns of a row, an fp16 pair of a row is one 4-byte store (a warp's store = 8 rows x 16 B contiguous, in
// shared memory conflict-free) and row statistics take two xor-shuffles across the four threads of a row.
// One tcgen05.ld.16x256b.x4 = block (b, h): lanes 16 h .. 16 h + 15 of the quarter, columns 32 b .. 32 b + 31 of the half;
// register 4 g' + 2 rr + e <-> row rho = 2 h + rr, column group g = 4 b + g'.
__device__ __forceinline__ void tmem_ld_q(uint32_t taddr, float (&v)[16]) {
    uint32_t* u = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                   "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_q(uint32_t taddr, const float (&v)[16]) {
    const uint32_t* u = reinterpret_cast<const uint32_t*>(v);
    asm volatile("tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
                 :: "r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]),
                    "r"(u[8]), "r"(u[9]), "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15]) : "memory");
}
__device__ __forceinline__ uint32_t q_blk(uint32_t tq, int b, int h) { return tq + ((uint32_t)(16 * h) << 16) + 32 * b; }
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }

// pipelined walk over the four blocks of a region, (b, h) = (0,0) (0,1) (1,0) (1,1): prep(b) once per column block (loads
// the column constants of its four groups), body(b, h, v) per block
template <class P, class F>
__device__ __forceinline__ void for_blocks_q(uint32_t tq, P&& prep, F&& body) {
    float a[16], c[16];
    tmem_ld_q(tq, a);
#pragma unroll 1
    for (int b = 0; b < 2; ++b) {
        prep(b);
        tmem_wait_ld();
        tmem_ld_q(q_blk(tq, b, 1), c);
        body(b, 0, a);
        tmem_wait_ld();
        if (b == 0) tmem_ld_q(q_blk(tq, 1, 0), a);
        body(b, 1, c);
    }
}

// per-thread accumulators of the shifted row sums of its four rows (packed over the two columns of a pair)
struct QStats {
    float shift[4], s0[4], s1[4], q0[4], q1[4];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < 4; ++i) shift[i] = s0[i] = s1[i] = q0[i] = q1[i] = 0.f;
    }
    // one block (b, h): b == 0 also fixes the shift of rows 2 h, 2 h + 1 (the row's first column of this half, held by T & 3 == 0)
    __device__ __forceinline__ void add(int b, int h, const float (&v)[16]) {
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int rho = 2 * h + rr;
            if (b == 0) shift[rho] = __shfl_sync(0xffffffffu, v[2 * rr], (threadIdx.x & 31) & ~3);
            const float ns = -shift[rho];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float d0, d1;
                add2(d0, d1, v[4 * g + 2 * rr], v[4 * g + 2 * rr + 1], ns, ns);
                add2(s0[rho], s1[rho], s0[rho], s1[rho], d0, d1);
                fma2(q0[rho], q1[rho], d0, d1, d0, d1, q0[rho], q1[rho]);
            }
        }
    }
    // half-row statistics of row rho, identical in the four threads of the row
    __device__ __forceinline__ HalfStats finish(int rho) const {
        float sum = s0[rho] + s1[rho], sq = q0[rho] + q1[rho];
        sum += __shfl_xor_sync(0xffffffffu, sum, 1); sq += __shfl_xor_sync(0xffffffffu, sq, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2); sq += __shfl_xor_sync(0xffffffffu, sq, 2);
        return half_stats(shift[rho], sum, sq);
    }
};
// LayerNorm scalars of a thread's four rows: rs = rstd, nm = -mean * rstd
struct QRows { float rs[4], nm[4]; };

// merge the two half-row statistics of the thread's four rows (Chan's formula): thread T & 3 == rho publishes row rho
// slot = exchange buffer [2 halves][128 rows] float2 of this tile; r0 = 32 q + (T >> 2), row rho = r0 + 8 rho
__device__ __forceinline__ QRows merge_stats_q(const QStats& qs, float2* slot, int r0, int t, int hh, int bar_id, float eps) {
    HalfStats hs[4];
#pragma unroll
    for (int rho = 0; rho < 4; ++rho) hs[rho] = qs.finish(rho);
    {
        const float m = t == 0 ? hs[0].mean : (t == 1 ? hs[1].mean : (t == 2 ? hs[2].mean : hs[3].mean));
        const float v = t == 0 ? hs[0].m2 : (t == 1 ? hs[1].m2 : (t == 2 ? hs[2].m2 : hs[3].m2));
        slot[hh * TILE_ROWS + r0 + 8 * t] = make_float2(m, v);
    }
    asm volatile("bar.sync %0, 256;\n" :: "r"(bar_id) : "memory");
    QRows o;
#pragma unroll
    for (int rho = 0; rho < 4; ++rho) {
        const float2 p = slot[(hh ^ 1) * TILE_ROWS + r0 + 8 * rho];
        const float dm = hs[rho].mean - p.x;
        const float mean = 0.5f * (hs[rho].mean + p.x);
        o.rs[rho] = rsqrtf((hs[rho].m2 + p.y + 32.f * dm * dm) * (1.f / D) + eps);
        o.nm[rho] = -mean * o.rs[rho];
    }
    return o;
}

// hn = hin + gate * (acc + bias) over the thread's 4 x 16 elements, written back over the accumulator; hin prefetched into
// registers: hq[8 rho + g] = columns c(g, 0..1) of row rho.  gate / bias already offset to the column half + 2 (T & 3).
__device__ __forceinline__ void resid_pass_regs_q(uint32_t tq, const float* __restrict__ gate, const float* __restrict__ bias,
                                                  const float2 (&hq)[32], QStats& qs) {
    qs.init();
    float2 g2[4], b2[4];
    for_blocks_q(tq, [&](int b) {
#pragma unroll
        for (int g = 0; g < 4; ++g) { g2[g] = ld2(gate + 32 * b + 8 * g); b2[g] = ld2(bias + 32 * b + 8 * g); }
    }, [&](int b, int h, float (&v)[16]) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                float t0, t1;
                const float2 hin = hq[8 * (2 * h + rr) + 4 * b + g];
                add2(t0, t1, v[4 * g + 2 * rr], v[4 * g + 2 * rr + 1], b2[g].x, b2[g].y);
                fma2(v[4 * g + 2 * rr], v[4 * g + 2 * rr + 1], g2[g].x, g2[g].y, t0, t1, hin.x, hin.y);
            }
        qs.add(b, h, v);
        tmem_st_q(q_blk(tq, b, h), v);
    });
    tmem_wait_st();
}
// the same with hin = the row parked in another TMEM region; STORE: hn also goes to the residual tile in global memory
// (hdst = tile + 2 (T & 3) % 4 ... : element (row R, column c) at (c / 4) * TILE_ROWS * 4 + R * 4 + c % 4; see the caller)
template <bool STORE>
__device__ __forceinline__ void resid_pass_tmem_q(uint32_t tacc, uint32_t thin, const float* __restrict__ gate, const float* __restrict__ bias,
                                                  float* __restrict__ hdst, const bool (&valid)[4], QStats& qs) {
    qs.init();
    float2 g2[4], b2[4];
    float a[16], x[16];
#pragma unroll 1
    for (int b = 0; b < 2; ++b) {
#pragma unroll
        for (int g = 0; g < 4; ++g) { g2[g] = ld2(gate + 32 * b + 8 * g); b2[g] = ld2(bias + 32 * b + 8 * g); }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            tmem_ld_q(q_blk(tacc, b, h), a);
            tmem_ld_q(q_blk(thin, b, h), x);
            tmem_wait_ld();
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    float t0, t1;
                    add2(t0, t1, a[4 * g + 2 * rr], a[4 * g + 2 * rr + 1], b2[g].x, b2[g].y);
                    fma2(a[4 * g + 2 * rr], a[4 * g + 2 * rr + 1], g2[g].x, g2[g].y, t0, t1, x[4 * g + 2 * rr], x[4 * g + 2 * rr + 1]);
                }
            qs.add(b, h, a);
            tmem_st_q(q_blk(tacc, b, h), a);
            if (STORE) {
#pragma unroll
                for (int rr = 0; rr < 2; ++rr)
                    if (valid[2 * h + rr]) {
#pragma unroll
                        for (int g = 0; g < 4; ++g)       // column 32 b + 8 g (+ 2 (T & 3), in hdst): col chunk + 8 b + 2 g
                            *reinterpret_cast<float2*>(hdst + (8 * b + 2 * g) * TILE_ROWS * 4 + (16 * h + 8 * rr) * 4) =
                                make_float2(a[4 * g + 2 * rr], a[4 * g + 2 * rr + 1]);
                    }
            }
        }
    }
    tmem_wait_st();
}

// LayerNorm (no affine) + modulate, packed to fp16 into the next GEMM's A operand image.  shift / scale1 offset to the
// column half + 2 (T & 3); arow = abuf + kc0 * KCH + r0 * 16 + 4 (T & 3) (kc0 = first K chunk of the half)
__device__ __forceinline__ void ln_mod_store_q(uint32_t tq, const QRows& st, const float* __restrict__ shift, const float* __restrict__ scale1, uint8_t* arow) {
    float2 sc[4], sh[4];
    for_blocks_q(tq, [&](int b) {
#pragma unroll
        for (int g = 0; g < 4; ++g) { sc[g] = ld2(scale1 + 32 * b + 8 * g); sh[g] = ld2(shift + 32 * b + 8 * g); }
    }, [&](int b, int h, float (&v)[16]) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int rho = 2 * h + rr;
                float n0, n1, y0, y1;
                fma2(n0, n1, v[4 * g + 2 * rr], v[4 * g + 2 * rr + 1], st.rs[rho], st.rs[rho], st.nm[rho], st.nm[rho]);
                fma2(y0, y1, n0, n1, sc[g].x, sc[g].y, sh[g].x, sh[g].y);
                *reinterpret_cast<uint32_t*>(arow + (4 * b + g) * KCH + (8 * rho) * 16) = pack_h2(y0, y1);
            }
    });
}

// hidden = GELU_tanh(acc + b1), packed to fp16 into an A operand image
__device__ __forceinline__ void gelu_store_q(uint32_t tq, const float* __restrict__ bias, uint8_t* arow) {
    float2 b2[4];
    for_blocks_q(tq, [&](int b) {
#pragma unroll
        for (int g = 0; g < 4; ++g) b2[g] = ld2(bias + 32 * b + 8 * g);
    }, [&](int b, int h, float (&v)[16]) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                float x0, x1, y0, y1;
                add2(x0, x1, v[4 * g + 2 * rr], v[4 * g + 2 * rr + 1], b2[g].x, b2[g].y);
                gelu_tanh2(y0, y1, x0, x1);
                *reinterpret_cast<uint32_t*>(arow + (4 * b + g) * KCH + (8 * (2 * h + rr)) * 16) = pack_h2(y0, y1);
            }
    });
}

// grid = min(#work items, #SMs), block = 576;  H = latent width (DitShape)
// NE = pair tiles per work item: 2 (tiles alternate on the tensor pipe: throughput) or 1 (small batches: twice the CTAs,
// shorter items; the second tile's epilogue warps idle)
template <int MODE, int H, int NE = 2>
__global__ void __launch_bounds__(TC_THREADS, 1) token_kernel(const TokArgs p) {
    using S = DitShape<H>;
    [[maybe_unused]] constexpr int NTOK = S::NTOK, TILE_TOK = S::TILE_TOK, TILES_PER_PAIR = S::TILES_PER_PAIR, LATP = S::H, LAT = S::LAT;
    [[maybe_unused]] constexpr int QT_ROWS = S::QT_ROWS, QKV_Q_HALVES = S::Q_HALVES, QKV_K_HALVES = S::K_HALVES, QKV_HEAD_HALVES = S::HEAD_HALVES;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform
    const uint32_t sb = smem_u32(smem);
    const uint32_t bar0 = sb + TC_SM_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    auto TBAR = [&](int e, int i) { return bar0 + 8u * (B_TILE + e * T_COUNT + i); };
    const int l = p.layer;                                         // block whose second half runs here (MID / FINAL)
    const int ln = (MODE == TOK_EMBED) ? 0 : l + 1;                // block whose QKV is produced here (EMBED / MID)
    constexpr int N_STAGES = (MODE == TOK_EMBED) ? 3 : (MODE == TOK_MID ? 8 : 5);
    const int n_items = ((p.nseq + 1) / 2) * (TILES_PER_PAIR / NE);  // work item = NE consecutive pair tiles

    if (tid == 0) {
        for (int i = 0; i < B_VFREE; ++i)                                         // WFULL, WEMPTY (one commit per issuer warp), VFULL
            mbar_init(BAR(i), (i >= B_WEMPTY && i < B_VFULL && TC_MMA_WARPS == 2) ? NE : 1);
        for (int i = 0; i < 2; ++i) mbar_init(BAR(B_VFREE + i), 8 * NE);          // one arrival per epilogue WARP (mbar_arrive_warp)
        for (int e = 0; e < NE; ++e) {
            mbar_init(TBAR(e, T_OFULL), 1);
            for (int i = T_A2; i <= T_DONE; ++i) mbar_init(TBAR(e, i), 8);
            mbar_init(TBAR(e, T_YFREE), 8);
            mbar_init(TBAR(e, T_HAFREE), 1);
            for (int i = T_ACC; i < T_ACC + 7; ++i) mbar_init(TBAR(e, i), 1);
        }
        mbar_fence_init();
    }
    if (warp == 16) tmem_alloc(sb + TC_SM_TMEM, 512);
    if (MODE == TOK_MID && p.recompute_h0) {
        float* semb = reinterpret_cast<float*>(smem + TC_SM_EMB);
        for (int i = tid; i < 5 * D; i += TC_THREADS) semb[i] = i < 4 * D ? p.w.w_embed[i] : p.w.b_embed[i - 4 * D];
    }
    // constant A block of the bias MMA: K chunk 0 = rows of {1, 1, 0, 0, 0, 0, 0, 0}, K chunk 1 = zeros
    for (int i = tid; i < WBIAS_BYTES / 16; i += TC_THREADS)
        reinterpret_cast<uint4*>(smem + TC_SM_ONES)[i] = make_uint4(i < KCH / 16 ? 0x3c003c00u : 0u, 0u, 0u, 0u);
    fence_async_smem();
    pdl_launch_dependents();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t*>(smem + TC_SM_TMEM), 0);
    pdl_wait();                  // set-up above overlapped the previous kernel's tail; everything below reads what it wrote

    if (warp == 16) {
        // ================================================================= producer (whole warp converged; lane 0 issues)
        const bool lead = lane == 0;
        const char* src_a = reinterpret_cast<const char*>(MODE == TOK_EMBED ? p.w.w_qkv[0] : p.w.w_post[l]);
        const char* src_b = reinterpret_cast<const char*>(MODE == TOK_MID ? p.w.w_qkv[l + 1] : nullptr);
        constexpr int N_A = (MODE == TOK_EMBED) ? 3 : 5;
        // inputs of work item `item` (iteration it): per-pair vectors into vector buffer it & 1, the two
        // attention-output tiles into the HA buffers, the residual tiles towards L2
        auto fetch_inputs = [&](int it, int item) {
            const int pair = item / (TILES_PER_PAIR / NE), vb = it & 1;
            if (it >= 2) mbar_wait(BAR(B_VFREE + vb), ((it >> 1) - 1) & 1);     // the item two back has released this buffer
            const int sq0 = p.uncond_shared ? 0 : min(2 * pair, p.nseq - 1), sq1 = min(2 * pair + 1, p.nseq - 1);
            const uint32_t vdst = sb + TC_SM_VEC + vb * (V_FLOATS * 4), vbar = BAR(B_VFULL + vb);
            uint32_t bytes = 0;
            auto cp = [&](int voff, const float* src, uint32_t n) {
                if (lead) bulk_g2s(vdst + voff * 4, src, n * 4, vbar);
                bytes += n * 4;
            };
            constexpr uint32_t VBYTES = (MODE == TOK_EMBED ? 0u : (2 * MOD) * 4u) + (MODE == TOK_FINAL ? 0u : 512 * 4u) +
                                        (MODE == TOK_EMBED ? (5 * D) * 4u : 0u) + (MODE == TOK_FINAL ? (4 * D + 4) * 4u : 0u);
            if (lead) mbar_expect_tx(vbar, VBYTES);
            if (MODE != TOK_EMBED) {
                cp(V_MOD, p.mod + ((size_t)sq0 * NLAYER + l) * MOD, MOD);
                cp(V_MOD + MOD, p.mod + ((size_t)sq1 * NLAYER + l) * MOD, MOD);
            }
            if (MODE != TOK_FINAL) {
                cp(V_MODN, p.mod + ((size_t)sq0 * NLAYER + ln) * MOD, 256);
                cp(V_MODN + 256, p.mod + ((size_t)sq1 * NLAYER + ln) * MOD, 256);
            }
            if (MODE == TOK_EMBED) { cp(V_WEMB, p.w.w_embed, 4 * D); cp(V_BEMB, p.w.b_embed, D); }
            if (MODE == TOK_FINAL) { cp(V_WFIN, p.w.w_final, 4 * D); cp(V_BFIN, p.w.b_final, 4); }
            if (bytes != VBYTES) __trap();
            if (MODE != TOK_EMBED) {
                if (lead && !(MODE == TOK_MID && p.recompute_h0)) prefetch_l2(p.h + (size_t)item * NE * (TILE_ROWS * D), NE * TILE_ROWS * D * 4);
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    if (it >= 1) mbar_wait(TBAR(e, T_HAFREE), (it - 1) & 1);     // fc2's first K half of the previous item has read HA
                    if (lead) {
                        mbar_expect_tx(TBAR(e, T_OFULL), STAGE_BYTES);
                        bulk_g2s(sb + TC_SM_HA + e * STAGE_BYTES, reinterpret_cast<const char*>(p.o) + ((size_t)item * NE + e) * STAGE_BYTES,
                                 STAGE_BYTES, TBAR(e, T_OFULL));
                    }
                }
            }
            __syncwarp();
        };
        int gs = 0;                                                     // global weight-stage counter (ring of TC_NSTAGE slots)
        if ((int)blockIdx.x < n_items) fetch_inputs(0, blockIdx.x);
#pragma unroll 1
        for (int it = 0, item = blockIdx.x; item < n_items; ++it, item += gridDim.x) {
#pragma unroll 1
            for (int s = 0; s < N_STAGES; ++s, ++gs) {
                const int slot = gs % TC_NSTAGE, use = gs / TC_NSTAGE;
                if (use > 0) mbar_wait(BAR(B_WEMPTY + slot), (use - 1) & 1);
                if (lead) {
#ifdef T2S_EXP_SKIP_WLOAD      // timing experiment only (wrong results): weight stages are fetched for a CTA's first item only
                    if (it > 0) mbar_arrive(BAR(B_WFULL + slot)); else {
#endif
                    mbar_expect_tx(BAR(B_WFULL + slot), WSTAGE_BYTES);
                    bulk_g2s(sb + TC_SM_W + slot * WSTAGE_BYTES, s < N_A ? src_a + (size_t)s * WSTAGE_BYTES : src_b + (size_t)(s - N_A) * WSTAGE_BYTES,
                             WSTAGE_BYTES, BAR(B_WFULL + slot));
#ifdef T2S_EXP_SKIP_WLOAD
                    }
#endif
                }
                __syncwarp();
                // the next item's inputs: as early as its buffers can be free (MID / FINAL: after the fc2 stages have been queued)
                if (s == (MODE == TOK_EMBED ? 0 : 4) && item + (int)gridDim.x < n_items) fetch_inputs(it + 1, item + gridDim.x);
            }
        }
        __syncwarp();
    } else if (warp >= 17) {
        // ================================================================= MMA issuer (whole warp converged; lane 0 issues)
        // -DT2S_TOK_MMA_WARPS=2 gives every tile its own issuer warp (17 + e; the ring slots are then released by one commit per
        // issuer).  Measured (round 2): the leading tile no longer waits behind the other tile's pending event (its item 31.2 k ->
        // 27.5 k cycles in the phase trace) but the step gets 2 % SLOWER (4.65 -> 4.75 ms): the tiles share the SM's load/store,
        // FMA and MUFU throughput, which is what paces an item, so the single in-order issuer stays the default.
        // E0 / E1 = range of tiles this warp serves.
        const bool lead = lane == 0;
        const int E0 = TC_MMA_WARPS == 2 ? warp - 17 : 0, E1 = TC_MMA_WARPS == 2 ? min(warp - 16, NE) : NE;
        int gs = 0;                                                     // global weight-stage counter
        auto wfull = [&](int g) { mbar_wait(BAR(B_WFULL + g % TC_NSTAGE), (g / TC_NSTAGE) & 1); };
        auto wslot = [&](int g) { return sb + TC_SM_W + (g % TC_NSTAGE) * WSTAGE_BYTES; };
        const uint32_t ones = sb + TC_SM_ONES;
        auto wdone = [&](int g) { if (lead) umma_commit(BAR(B_WEMPTY + g % TC_NSTAGE)); __syncwarp(); };
#pragma unroll 1
        for (int it = 0, item = blockIdx.x; item < n_items && E0 < E1; ++it, item += gridDim.x) {
            const uint32_t par = it & 1;
            const uint32_t X = (it & 1) * 128, Y = 128 - X;             // the two TMEM regions swap roles every item
            // one weight stage feeds the same GEMM chunk of both tiles; chunks that wait on the same epilogue event
            // (fc2's two K halves; q and k) are issued tile by tile so that a tile never waits for its neighbour
            auto A_ = [&](int e) { return sb + TC_SM_A + e * STAGE_BYTES; };
            auto HA_ = [&](int e) { return sb + TC_SM_HA + e * STAGE_BYTES; };
            auto D_ = [&](int e, uint32_t col) { return tmem + e * 256 + col; };
            auto acc = [&](int e, int i) { if (lead) umma_commit(TBAR(e, T_ACC + i)); };
            if (MODE != TOK_EMBED) {
                wfull(gs);                                               // proj (o tile sits in HA) -> X
#pragma unroll 1
                for (int e = E0; e < E1; ++e) {
                    if (it > 0) mbar_wait(TBAR(e, T_DONE), (it - 1) & 1);   // the previous item has drained this region
                    mbar_wait(TBAR(e, T_OFULL), par);
                    tc_fence_after();
                    tc_gemm(HA_(e), wslot(gs), D_(e, X), false, lead, ones);
                    acc(e, 0);
                }
                wdone(gs); ++gs;
                wfull(gs);                                               // fc1 cols 0..127 -> Y
#pragma unroll 1
                for (int e = E0; e < E1; ++e) {
                    mbar_wait(TBAR(e, T_A2), par);
                    tc_fence_after();
                    tc_gemm(A_(e), wslot(gs), D_(e, Y), false, lead, ones);
                    acc(e, 1);
                }
                wdone(gs); ++gs;
                wfull(gs);                                               // fc1 cols 128..255 -> Y (hidden-a done: Y drained)
#pragma unroll 1
                for (int e = E0; e < E1; ++e) {
                    mbar_wait(TBAR(e, T_HA), par);
                    tc_fence_after();
                    tc_gemm(A_(e), wslot(gs), D_(e, Y), false, lead, ones);
                    acc(e, 2);
                }
                wdone(gs); ++gs;
                wfull(gs); wfull(gs + 1);                                // fc2, both K halves -> Y (hidden-b done: Y drained)
#pragma unroll 1
                for (int e = E0; e < E1; ++e) {
                    mbar_wait(TBAR(e, T_HB), par);
                    tc_fence_after();
                    tc_gemm(HA_(e), wslot(gs), D_(e, Y), false, lead, ones);
                    if (lead) umma_commit(TBAR(e, T_HAFREE));
                    tc_gemm(A_(e), wslot(gs + 1), D_(e, Y), true, lead);
                    acc(e, 3);
                }
                wdone(gs); wdone(gs + 1); gs += 2;
            }
            if (MODE != TOK_FINAL) {
#ifndef T2S_TOK_HSTORE_EARLY
                // MID writes the residual tile to global memory AFTER the LN pass (see the epilogue): the k chunk may only
                // overwrite region Y once that store pass has drained it, so q is issued for both tiles first and its ring slot
                // is released before k: the v stage is then on its way while the q / k epilogues run
                wfull(gs);                                               // q -> X
#pragma unroll 1
                for (int e = E0; e < E1; ++e) {
                    mbar_wait(TBAR(e, T_A3), par);
                    tc_fence_after();
                    tc_gemm(A_(e), wslot(gs), D_(e, X), false, lead, ones);
                    acc(e, 4);
                }
                wdone(gs); ++gs;
                wfull(gs);                                               // k -> Y
#pragma unroll 1
                for (int e = E0; e < E1; ++e) {
                    if (MODE == TOK_MID) { mbar_wait(TBAR(e, T_YFREE), par); tc_fence_after(); }
                    tc_gemm(A_(e), wslot(gs), D_(e, Y), false, lead, ones);
                    acc(e, 5);
                }
                wdone(gs); ++gs;
#else
                wfull(gs); wfull(gs + 1);                                // q -> X, k -> Y
#pragma unroll 1
                for (int e = E0; e < E1; ++e) {
                    mbar_wait(TBAR(e, T_A3), par);
                    tc_fence_after();
                    tc_gemm(A_(e), wslot(gs), D_(e, X), false, lead, ones);
                    acc(e, 4);
                    tc_gemm(A_(e), wslot(gs + 1), D_(e, Y), false, lead, ones);
                    acc(e, 5);
                }
                wdone(gs); wdone(gs + 1); gs += 2;
#endif
                wfull(gs);                                               // v -> X (after the q epilogue has drained X)
#pragma unroll 1
                for (int e = E0; e < E1; ++e) {
                    mbar_wait(TBAR(e, T_XFREE), par);
                    tc_fence_after();
                    tc_gemm(A_(e), wslot(gs), D_(e, X), false, lead, ones);
                    acc(e, 6);
                }
                wdone(gs); ++gs;
            }
        }
        __syncwarp();
    } else {
        // ================================================================= epilogue: threads (r, hh) <-> tile row r, columns 64 hh ..
        const int e = warp >> 3;                                        // tile handled by these eight warps
        const int hh = (warp >> 2) & 1;                                 // column half
        const int r = (warp & 3) * 32 + lane;
        const int branch = r >> 6, tl = r & 63;
        const int c0 = hh * 64, kc0 = hh * 8;                           // first column / first 8-column K chunk of the half
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + e * 256 + c0;
        uint8_t* abuf = smem + TC_SM_A + e * STAGE_BYTES;
        uint8_t* habuf = smem + TC_SM_HA + e * STAGE_BYTES;
        float2* stx = reinterpret_cast<float2*>(smem + TC_SM_ST) + e * (2 * TILE_ROWS);
        const bool tr = p.trace != nullptr && e == 0 && r == 0 && hh == 0;
#define STAMP(i) do { if (tr) p.trace[(size_t)blockIdx.x * 32 + (i)] = clock64(); } while (0)
#pragma unroll 1
        for (int it = 0, item = blockIdx.x; item < n_items && e < NE; ++it, item += gridDim.x) {
        const uint32_t par = it & 1;
        const uint32_t X = (it & 1) * 128, Y = 128 - X;                 // the two TMEM regions swap roles every item
        const int pair = item / (TILES_PER_PAIR / NE), tt = (item % (TILES_PER_PAIR / NE)) * NE + e;
        const int seq = 2 * pair + branch;
        const bool valid = tl < TILE_TOK && seq < p.nseq;
        const int tok = tt * TILE_TOK + tl;
        const float* vec = reinterpret_cast<const float*>(smem + TC_SM_VEC) + (it & 1) * V_FLOATS;
        const float* modb = vec + V_MOD + branch * MOD;
        float* htile = p.h + ((size_t)item * NE + e) * (TILE_ROWS * D);  // [32 col chunks][128 rows][4]
        const float* hrow_c = htile + (c0 / 4) * TILE_ROWS * 4 + r * 4; // + c4 * TILE_ROWS * 4
        float* hrow = htile + (c0 / 4) * TILE_ROWS * 4 + r * 4;
        STAMP(0);
        mbar_wait(BAR(B_VFULL + (it & 1)), (it >> 1) & 1);
        STAMP(2);
        RowStats st;
        if (MODE == TOK_EMBED) {
            // patchify + patch_emb + pos_embed (transformer.py:166-172), conv folded into the Linear; the row is
            // parked in TMEM region X (not yet an accumulator) for the LayerNorm pass
            float xv[4] = {0.f, 0.f, 0.f, 0.f};
            if (valid) {
                const float* xs = p.x + (size_t)(seq >> p.x_shift) * LAT;
                const int i = tok >> 5, j = tok & 31;
#pragma unroll
                for (int pq = 0; pq < 4; ++pq) xv[pq] = xs[(2 * j + (pq & 1)) * LATP + 2 * i + (pq >> 1)];
            }
            const float* pos = p.w.pos + ((size_t)tt * 32 * 64 + tl) * 4 + (c0 / 4) * 64 * 4;   // [8 tiles][32 chunks][64 rows][4]
            float sum = 0.f, sq = 0.f, shift = 0.f;
#pragma unroll 1
            for (int cb = 0; cb < 4; ++cb) {
                float a[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {                    // four columns: bias + pos, then the 4-pixel dot products, packed fp32
                    const int c = c0 + cb * 16 + q * 4;
                    const float4 pe = valid ? *reinterpret_cast<const float4*>(pos + (cb * 4 + q) * 64 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 b4 = *reinterpret_cast<const float4*>(vec + V_BEMB + c);
                    float y0, y1, y2, y3;
                    add2(y0, y1, b4.x, b4.y, pe.x, pe.y);
                    add2(y2, y3, b4.z, b4.w, pe.z, pe.w);
#pragma unroll
                    for (int pq = 0; pq < 4; ++pq) {
                        const float4 w4 = *reinterpret_cast<const float4*>(vec + V_WEMB + pq * D + c);   // [4 pixels][128 columns]
                        fma2(y0, y1, w4.x, w4.y, xv[pq], xv[pq], y0, y1);
                        fma2(y2, y3, w4.z, w4.w, xv[pq], xv[pq], y2, y3);
                    }
                    a[q * 4 + 0] = valid ? y0 : 0.f; a[q * 4 + 1] = valid ? y1 : 0.f;
                    a[q * 4 + 2] = valid ? y2 : 0.f; a[q * 4 + 3] = valid ? y3 : 0.f;
                }
                if (cb == 0) shift = a[0];
                block_stats(a, shift, sum, sq);
                tmem_st16(trow + X + cb * 16, a);
                if (valid && !p.skip_h_store) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<float4*>(hrow + (cb * 4 + q) * TILE_ROWS * 4) = make_float4(a[q * 4], a[q * 4 + 1], a[q * 4 + 2], a[q * 4 + 3]);
                }
            }
            tmem_wait_st();
            HalfStats hs;
            hs.mean = shift + sum * (1.f / 64);
            hs.m2 = fmaxf(sq - sum * sum * (1.f / 64), 0.f);
            st = merge_stats(hs, stx, r, hh, 1 + e, 1e-6f);
        } else {
            // x = x + gate_msa * (o Wproj^T + b)        (transformer.py:116); x stays parked in X until the MLP branch adds to it
            float4 hq[16];                                               // this thread's 64 residual values, in flight during the wait
#pragma unroll
            for (int c4 = 0; c4 < 16; ++c4) hq[c4] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (MODE == TOK_MID && p.recompute_h0) {
                // block 0: the residual input is patch-embed(x) + pos (transformer.py:166-172), recomputed exactly as EMBED formed it
                if (valid) {
                    const float* xs = p.x + (size_t)(seq >> p.x_shift) * LAT;
                    const int i = tok >> 5, j = tok & 31;
                    float xv[4];
#pragma unroll
                    for (int pq = 0; pq < 4; ++pq) xv[pq] = xs[(2 * j + (pq & 1)) * LATP + 2 * i + (pq >> 1)];
                    const float* pos = p.w.pos + ((size_t)tt * 32 * 64 + tl) * 4 + (c0 / 4) * 64 * 4;
                    const float* semb = reinterpret_cast<const float*>(smem + TC_SM_EMB);
#pragma unroll
                    for (int c4 = 0; c4 < 16; ++c4) {
                        const int c = c0 + c4 * 4;
                        const float4 pe = *reinterpret_cast<const float4*>(pos + c4 * 64 * 4);
                        const float4 b4 = *reinterpret_cast<const float4*>(semb + 4 * D + c);
                        float y0, y1, y2, y3;
                        add2(y0, y1, b4.x, b4.y, pe.x, pe.y);
                        add2(y2, y3, b4.z, b4.w, pe.z, pe.w);
#pragma unroll
                        for (int pq = 0; pq < 4; ++pq) {
                            const float4 w4 = *reinterpret_cast<const float4*>(semb + pq * D + c);
                            fma2(y0, y1, w4.x, w4.y, xv[pq], xv[pq], y0, y1);
                            fma2(y2, y3, w4.z, w4.w, xv[pq], xv[pq], y2, y3);
                        }
                        hq[c4] = make_float4(y0, y1, y2, y3);
                    }
                }
#ifdef T2S_KO_HLD
            } else if (valid && ko_never()) {
#else
            } else if (valid) {
#endif
#pragma unroll
                for (int c4 = 0; c4 < 16; ++c4) hq[c4] = ldg16f(hrow_c + c4 * TILE_ROWS * 4);
            }
            mbar_wait(TBAR(e, T_ACC + 0), par);
            tc_fence_after();
            STAMP(3);
            HalfStats hs = resid_pass_regs<false, false>(trow + X, modb + 2 * D + c0, nullptr, hq);
            st = merge_stats(hs, stx, r, hh, 1 + e, 1e-6f);
            ln_mod_store(trow + X, st, modb + 3 * D + c0, modb + 4 * D + c0, abuf, r, kc0);
            fence_async_smem();
            tc_fence_before();
            mbar_arrive_warp(TBAR(e, T_A2));
            STAMP(4);
            // hidden = GELU(fc1)                         (transformer.py:117)
            mbar_wait(TBAR(e, T_ACC + 1), par);
            tc_fence_after();
            STAMP(5);
            gelu_store<false, false>(trow + Y, nullptr, habuf, r, kc0);
            fence_async_smem();
            tc_fence_before();
            mbar_arrive_warp(TBAR(e, T_HA));
            STAMP(6);
            mbar_wait(TBAR(e, T_ACC + 2), par);
            tc_fence_after();
            STAMP(7);
            gelu_store<false, false>(trow + Y, nullptr, abuf, r, kc0);
            fence_async_smem();
            tc_fence_before();
            mbar_arrive_warp(TBAR(e, T_HB));
            STAMP(8);
            // x = x + gate_mlp * (hidden W2^T + b): X (parked x) + gate * Y -> Y
            mbar_wait(TBAR(e, T_ACC + 3), par);
            tc_fence_after();
            STAMP(9);
#ifndef T2S_TOK_HSTORE_EARLY
            hs = resid_pass_tmem<false, false, false>(trow + Y, trow + X, modb + 5 * D + c0, nullptr, hrow, valid);
#else
            hs = resid_pass_tmem<MODE == TOK_MID, false, false>(trow + Y, trow + X, modb + 5 * D + c0, nullptr, hrow, valid);
#endif
            st = merge_stats(hs, stx, r, hh, 1 + e, MODE == TOK_FINAL ? 1e-5f : 1e-6f);
            STAMP(10);
        }
        const uint32_t HREG = (MODE == TOK_EMBED) ? X : Y;              // TMEM region holding the residual row now

        if (MODE != TOK_FINAL) {
            const float* modn = vec + V_MODN + branch * 256;
            ln_mod_store(trow + HREG, st, modn + c0, modn + D + c0, abuf, r, kc0);
            fence_async_smem();
            tc_fence_before();
            mbar_arrive_warp(TBAR(e, T_A3));
            STAMP(11);
#ifndef T2S_TOK_HSTORE_EARLY
            if (MODE == TOK_MID) {
                // the updated residual tile goes to global memory HERE, while the tensor pipe computes q: a store pass is paced by
                // the SM's path to L2 (tools/probe_pass.cu: +1.7 k cycles per tile inside the residual pass), which this slot
                // hides; region Y belongs to the k chunk afterwards
                for_blocks16<4>(trow + Y, [&](int cb, float (&a)[16]) {
#ifdef T2S_KO_STG
                    if (valid && ko_never()) {
#else
                    if (valid) {
#endif
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            stg16f(hrow + (cb * 4 + q) * TILE_ROWS * 4, make_float4(a[q * 4], a[q * 4 + 1], a[q * 4 + 2], a[q * 4 + 3]));
                    }
                });
                tc_fence_before();
                mbar_arrive_warp(TBAR(e, T_YFREE));
            }
#endif
            // q | k | v = a' W^T + b, stored fp16 in the attention kernel's smem image layout
#pragma unroll 1
            for (int which = 0; which < 3; ++which) {
                mbar_wait(TBAR(e, T_ACC + 4 + which), par);
                tc_fence_after();
                STAMP(12 + 2 * which);
                const uint32_t tcol = which == 1 ? Y : X;
                // tcgen05 operand images read by attn_kernel (a warp's 32 rows are contiguous): element offset of this
                // token inside a (sequence, head) image and the stride between 8-wide d chunks
                const int off = which == 0 ? (tok / QT_ROWS) * 4096 + (tok % QT_ROWS) * 8
                              : which == 1 ? QKV_Q_HALVES + tok * 8
                                           : QKV_Q_HALVES + QKV_K_HALVES + (tok >> 3) * 256 + (tok & 7) * 8;
                const int dstride = which == 0 ? 1024 : (which == 1 ? NTOK * 8 : 64);
                __half* hb0 = p.qkv + ((size_t)seq * NHEAD + hh * 2) * QKV_HEAD_HALVES + off;
                for_blocks16<4>(trow + tcol, [&](int cb, float (&v)[16]) {
#ifdef T2S_KO_STG
                    if (valid && ko_never()) {
#else
                    if (valid) {
#endif
                        __half* hb = hb0 + (cb >> 1) * QKV_HEAD_HALVES + (cb & 1) * 2 * dstride;
#pragma unroll
                        for (int c = 0; c < 2; ++c) {                     // the bias is already in the accumulator (tc_gemm)
                            const float* x = v + c * 8;
                            stg16(hb + c * dstride, make_uint4(pack_h2(x[0], x[1]), pack_h2(x[2], x[3]), pack_h2(x[4], x[5]), pack_h2(x[6], x[7])));
                        }
                    }
                });
                STAMP(13 + 2 * which);
                if (which == 0) {                                        // X drained: the v chunk may overwrite it
                    tc_fence_before();
                    mbar_arrive_warp(TBAR(e, T_XFREE));
                } else if (which == 1 && MODE == TOK_MID) {              // Y drained: the next item's proj may overwrite it
                    tc_fence_before();
                    mbar_arrive_warp(TBAR(e, T_DONE));
                }
            }
        } else {
            // final LN (eps 1e-5, affine folded) + Linear(128->4) + unpatchify (transformer.py:182-190)
            float d4[4] = {0.f, 0.f, 0.f, 0.f};
            for_blocks16<4>(trow + HREG, [&](int cb, float (&a)[16]) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    float y0, y1, y2, y3;
                    add2(y0, y1, a[j], a[j + 1], -st.mean, -st.mean); add2(y2, y3, a[j + 2], a[j + 3], -st.mean, -st.mean);
                    mul2(y0, y1, y0, y1, st.rstd, st.rstd); mul2(y2, y3, y2, y3, st.rstd, st.rstd);
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const float4 w = *reinterpret_cast<const float4*>(vec + V_WFIN + c4 * D + c0 + cb * 16 + j);
                        d4[c4] = fmaf(y0, w.x, fmaf(y1, w.y, fmaf(y2, w.z, fmaf(y3, w.w, d4[c4]))));
                    }
                }
            });
            tc_fence_before();
            mbar_arrive_warp(TBAR(e, T_DONE));                                // Y drained: the next item's proj may overwrite it
            float* vb = reinterpret_cast<float*>(abuf);                     // [128][4] projection exchange (lives in the A buffer: see the note at TOK_SMEM_BYTES)
            float4* px = reinterpret_cast<float4*>(stx);                 // the statistics exchange is idle now: partial sums of half 1
            asm volatile("bar.sync %0, 256;\n" :: "r"(1 + e) : "memory");  // ... once every thread has read its merge partner
            if (hh == 1) px[r] = make_float4(d4[0], d4[1], d4[2], d4[3]);
            asm volatile("bar.sync %0, 256;\n" :: "r"(1 + e) : "memory");
            if (hh == 0) {
                const float4 o4 = px[r];
                *reinterpret_cast<float4*>(vb + r * 4) = make_float4(d4[0] + o4.x + vec[V_BFIN], d4[1] + o4.y + vec[V_BFIN + 1],
                                                                     d4[2] + o4.z + vec[V_BFIN + 2], d4[3] + o4.w + vec[V_BFIN + 3]);
            }
            asm volatile("bar.sync %0, 256;\n" :: "r"(1 + e) : "memory");
            const int t256 = hh * TILE_ROWS + r;
            if (p.out_mode == OUT_FWD) {
                for (int idx = t256; idx < 2 * TILE_TOK * 4; idx += 256) {
                    const int br = idx / (TILE_TOK * 4), rem = idx - br * (TILE_TOK * 4), t2 = rem >> 2, c4 = rem & 3;
                    const int sq = 2 * pair + br;
                    if (sq < p.nseq) {
                        const int n = tt * TILE_TOK + t2, i = n >> 5, jx = n & 31;
                        p.out[(size_t)sq * LAT + (2 * jx + (c4 & 1)) * LATP + 2 * i + (c4 >> 1)] = vb[(br * 64 + t2) * 4 + c4];
                    }
                }
            } else {
                // classifier-free guidance mix (infer.py:81/:87) + Euler (rectified_flow.py:5-7) or
                // DDPM ancestral update (DDPM.py:28-36); the latent is updated in place
                for (int idx = t256; idx < TILE_TOK * 4; idx += 256) {
                    const int t2 = idx >> 2, c4 = idx & 3;
                    const int n = tt * TILE_TOK + t2, i = n >> 5, jx = n & 31;
                    const size_t xi = (size_t)pair * LAT + (2 * jx + (c4 & 1)) * LATP + 2 * i + (c4 >> 1);
                    const float u = vb[t2 * 4 + c4], c = vb[(64 + t2) * 4 + c4];
                    const float pred = u + p.cfg * (c - u);
                    if (p.out != nullptr) p.out[xi] = pred;
                    const float xo = p.x_upd[xi];
                    float xn;
                    if (p.out_mode == OUT_RF) {
                        xn = xo + pred * p.c1;
                    } else {
                        const float mean2 = p.c1 * (xo - p.c2 * pred);
                        xn = mean2 + p.c3 * (p.noise != nullptr ? p.noise[xi] : philox_normal(p.seed, p.step, xi));
                    }
                    p.x_upd[xi] = xn;
                }
            }
        }
        // item done: this vector buffer may be reused (the producer's copies for the item after next wait on it)
        mbar_arrive_warp(BAR(B_VFREE + (it & 1)));
        STAMP(18);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) tmem_dealloc(tmem, 512);
}

// =================================================================================== attention (tcgen05)
// softmax(q k^T / sqrt(32)) v for one (sequence, head) per CTA (timm Attention -> F.scaled_dot_product_attention).
// The 480 queries form four 120-row q-tiles (M = 128 with 8 padding rows), the 480 keys ten 48-key chunks.
// Single pass, thread per query row (= TMEM lane), a 48-key score chunk held in registers:
//   S_j = Q_tile K_j^T (tcgen05.mma M128 N48 K16 x2, double-buffered in TMEM) is read ONCE (a second pass over S
//   costs the softmax warps another dependent round of loads); the row maximum of the chunk is taken with FMNMX3;
//   P_j = exp2((S_j - m) * log2e/sqrt(32)) is packed to fp16 and written back over the first 24 columns of its own
//   S buffer (tcgen05.st); O += P_j V_j is a tcgen05.mma with the A operand read from TMEM (M128 N32 K16 x3, V as
//   MN-major B operand) accumulating over the chunks in one 32-column accumulator; O / rowsum is written straight
//   into the out-projection's A-operand tile.
//   The reference point m is the maximum of the first chunk and is only moved when a later chunk exceeds it by more
//   than 2^8 in the exp2 domain (P <= 256 stays exact in fp16, sums are fp32): then the row sum and the O row are
//   rescaled (rare; taken warp-uniformly after the previous P.V has completed).  The result is the exact softmax.
// CTA = TWO softmax warpgroups (thread = query row), each with its own MMA warp, barriers and 128 TMEM columns
// (two 48-column S buffers + the 32-column O), on alternate q-tiles over the shared Q / K / V images; 92 KB of shared
// memory, so TWO CTAs share an SM = FOUR softmax warps per scheduler.  The kernel sits on the MUFU; a softmax warp's
// chunk is a serial load -> max -> exp -> store -> arrive chain, and with only two such warps per scheduler (one
// warpgroup per CTA, 96-key chunks: the previous form) the MUFU idled whenever both were outside their exp loop
// (0.636 -> 0.597 ms per 2048-sequence launch).  The wide shapes (K / V of one head no longer fit twice) run ONE CTA
// per SM with FOUR warpgroups on 32-key chunks (H = 64: 8.36 -> 7.55 ms per guided step at batch 512).
// Q, K, V arrive by bulk async copies: the token kernel stores them directly as tcgen05 operand images
//   Q: [q-tile][d/8][row 0..127][8]     (A, K-major)         32768 B
//   K: [d/8][key 0..479][8]             (B, K-major)         30720 B
//   V: [key/8][d/8][key%8][8]           (B, MN-major)        30720 B
constexpr int ATT_THREADS = 160;
// barriers: Q / K / V loaded (shared), then per warpgroup a block of 3 NBUF + 2: S full [NBUF] | P full [NBUF] | P.V done [NBUF] | O full | O free
enum { AB_QFULL = 0, AB_KFULL = 1, AB_VFULL = 2, AB_WG = 3 };
#ifndef T2S_ATT_NBUF
#define T2S_ATT_NBUF 2          // score buffers per warpgroup of the T2S shape: 2 x 48 keys (default) or 3 x 32 keys (A/B build)
#endif
// LAT: the small-batch form of the T2S shape (one CTA per q-tile, see attn_kernel): ONE warpgroup and 96-key chunks, the
// shortest serial chain for a single q-tile when there is nothing to overlap it with
template <int H, bool LAT = false>
struct AttShape {
    using S = DitShape<H>;
    static constexpr bool ONE_WG = LAT && H == 30;
    static constexpr int NBUF = (!ONE_WG && H == 30) ? T2S_ATT_NBUF : 2;                   // S buffers per warpgroup (double / triple buffering)
    static constexpr int KC = ONE_WG ? 96 : (NBUF == 3 ? 32 : S::KC), NCH = S::NTOK / KC, NQT = S::NQT;      // keys per chunk, chunks, q-tiles
    static constexpr int AB_SFULL = 0, AB_PFULL = NBUF, AB_PVDONE = 2 * NBUF, AB_OFULL = 3 * NBUF, AB_OFREE = 3 * NBUF + 1, AB_BLOCK = 3 * NBUF + 2;
    static constexpr int SM_Q = 0, SM_K = S::Q_HALVES * 2, SM_V = SM_K + S::K_HALVES * 2;
    static constexpr int SM_BAR = SM_V + S::V_HALVES * 2;
    static constexpr int SM_TMEM = SM_BAR + 48 * 8;
    static constexpr int SMEM_BYTES = SM_TMEM + 16;
    static constexpr int CTAS_PER_SM = 2 * (SMEM_BYTES + 1024) <= 233472 ? 2 : 1;   // H = 30: two CTAs share an SM
    // TWO softmax warpgroups (+ their MMA warps) per CTA on alternate q-tiles over the shared K / V images, so one group's
    // MMA / TMEM phases hide under the other's exps; FOUR (32-key chunks) where only one CTA fits an SM (wide shapes)
    static constexpr int NWG = ONE_WG ? 1 : (CTAS_PER_SM == 2 ? 2 : 4);   // four softmax warpgroups per SM either way
    static constexpr int THREADS = NWG * ATT_THREADS;
    static constexpr uint32_t IDESC_S = umma_idesc_f16(128, KC);
    static constexpr uint32_t TCOLS_WG = NBUF * KC + HD <= 128 ? 128 : 256;   // per warpgroup: NBUF S buffers (NBUF x KC) + O (32)
    static constexpr uint32_t TCOLS = NWG * TCOLS_WG;
    static constexpr uint32_t T_S = 0, T_O = NBUF * KC;
    static_assert(SMEM_BYTES <= 232448 && NBUF * KC + HD <= 256 && NQT % NWG == 0 && AB_WG + NWG * AB_BLOCK <= 48, "attention kernel resources");
};
static_assert(AttShape<30>::CTAS_PER_SM == 2, "two attention CTAs must fit one SM for the T2S shape");
constexpr uint32_t ATT_IDESC_PV = umma_idesc_f16(128, HD) | (1u << 16);   // B (V) is MN-major

// grid = nseq * 4, block = 160 per warpgroup: warps 0-3 = softmax (thread = query row = TMEM lane), warp 4 = (loads +) MMA issue
template <int H, bool LAT = false>
__global__ void __launch_bounds__(AttShape<H, LAT>::THREADS, AttShape<H, LAT>::CTAS_PER_SM) attn_kernel(const __half* __restrict__ qkv, __half* __restrict__ o, long long* trace, int npart) {
    using S = DitShape<H>;
    using AS = AttShape<H, LAT>;
    constexpr int NTOK = S::NTOK, TILE_TOK = S::TILE_TOK, TILES_PER_PAIR = S::TILES_PER_PAIR, QT_ROWS = S::QT_ROWS;
    constexpr int QKV_Q_HALVES = S::Q_HALVES, QKV_K_HALVES = S::K_HALVES, QKV_V_HALVES = S::V_HALVES, QKV_HEAD_HALVES = S::HEAD_HALVES;
    constexpr int ATT_KC = AS::KC, ATT_NCH = AS::NCH, ATT_NQT = AS::NQT;
    constexpr int ATT_SM_Q = AS::SM_Q, ATT_SM_K = AS::SM_K, ATT_SM_V = AS::SM_V, ATT_SM_BAR = AS::SM_BAR, ATT_SM_TMEM = AS::SM_TMEM;
    constexpr uint32_t ATT_IDESC_S = AS::IDESC_S, ATT_TCOLS = AS::TCOLS, ATT_T_S = AS::T_S, ATT_T_O = AS::T_O;
    constexpr int NWG = AS::NWG, NB = AS::NBUF;
    constexpr int AB_SFULL = AB_WG + AS::AB_SFULL, AB_PFULL = AB_WG + AS::AB_PFULL, AB_PVDONE = AB_WG + AS::AB_PVDONE, AB_OFULL = AB_WG + AS::AB_OFULL,
                  AB_OFREE = AB_WG + AS::AB_OFREE;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp_cta = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform
    const int wg = warp_cta / 5, warp = warp_cta % 5;                 // warpgroup; role inside it (0-3 softmax, 4 MMA)
    // npart = 1: one CTA per (sequence, head) walks over all q-tiles.  npart = NQT / NWG (small batches: too few (sequence,
    // head) pairs to fill the GPU): one CTA per q-tile (group), K / V are re-read from L2 by every part.
    const int sh = blockIdx.x / npart, part = blockIdx.x - sh * npart;
    const int seq = sh >> 2, head = sh & 3;
    const uint32_t sb = smem_u32(smem);
    const uint32_t bar0 = sb + ATT_SM_BAR;
    // barriers 0-2 (Q / K / V loaded) are shared; every warpgroup owns a block of ten (AB_SFULL .. AB_OFREE)
    auto BAR = [&](int i) { return bar0 + 8u * (i < 3 ? i : i + wg * AS::AB_BLOCK); };
    if (warp_cta == 4 && lane == 0) {
        // the operand loads go out first: their latency overlaps the TMEM allocation, the remaining barrier set-up and
        // the CTA-wide synchronisation
        for (int i = 0; i < 3; ++i) mbar_init(bar0 + 8u * i, 1);
        mbar_fence_init();
        pdl_launch_dependents();
        pdl_wait();              // q|k|v come from the previous kernel
        const char* src = reinterpret_cast<const char*>(qkv + (size_t)sh * QKV_HEAD_HALVES);
        if (npart == 1) {
            mbar_expect_tx(bar0 + 8u * AB_QFULL, QKV_Q_HALVES * 2);
            bulk_g2s(sb + ATT_SM_Q, src, QKV_Q_HALVES * 2, bar0 + 8u * AB_QFULL);
        } else {                                                 // only this part's NWG q-tiles (adjacent 8 KB images)
            mbar_expect_tx(bar0 + 8u * AB_QFULL, NWG * 8192);
            bulk_g2s(sb + ATT_SM_Q + part * NWG * 8192, src + part * NWG * 8192, NWG * 8192, bar0 + 8u * AB_QFULL);
        }
        mbar_expect_tx(bar0 + 8u * AB_KFULL, QKV_K_HALVES * 2);
        bulk_g2s(sb + ATT_SM_K, src + QKV_Q_HALVES * 2, QKV_K_HALVES * 2, bar0 + 8u * AB_KFULL);
        mbar_expect_tx(bar0 + 8u * AB_VFULL, QKV_V_HALVES * 2);
        bulk_g2s(sb + ATT_SM_V, src + (QKV_Q_HALVES + QKV_K_HALVES) * 2, QKV_V_HALVES * 2, bar0 + 8u * AB_VFULL);
    }
    if (tid == 0) {
        for (int g = 0; g < NWG; ++g) {
            const uint32_t bg = bar0 + 8u * (g * AS::AB_BLOCK);
            for (int b = 0; b < NB; ++b) {
                mbar_init(bg + 8u * (AB_SFULL + b), 1);
                mbar_init(bg + 8u * (AB_PFULL + b), 4);            // one arrival per softmax WARP (mbar_arrive_warp)
                mbar_init(bg + 8u * (AB_PVDONE + b), 1);
            }
            mbar_init(bg + 8u * AB_OFULL, 1);
            mbar_init(bg + 8u * AB_OFREE, 4);
        }
        mbar_fence_init();
    }
    if (warp_cta == 4) { __syncwarp(); tmem_alloc(sb + ATT_SM_TMEM, ATT_TCOLS); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();                  // (the attention-output tiles this kernel overwrites were read by the previous kernel)
    const uint32_t tmem = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t*>(smem + ATT_SM_TMEM), 0) + wg * AS::TCOLS_WG;
    const int nql = ATT_NQT / NWG / npart;                              // q-tiles of this warpgroup in this CTA
    const int NG = nql * ATT_NCH;                                       // its score chunks: (local q-tile, key chunk)
    auto q_tile = [&](int ql) { return (ql * npart + part) * NWG + wg; };

    if (warp == 4) {
        // ================================================================= loads + MMA issue; whole warp converged
        const bool lead = lane == 0;
        // score chunk G: q-tile (G / NCH) * NWG + wg, key chunk G % NCH -> S buffer G % NB
        auto issue_s = [&](int G) {
            const int qt = q_tile(G / ATT_NCH), j = G % ATT_NCH;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                const uint64_t ad = umma_desc(sb + ATT_SM_Q + qt * 8192 + kk * 2 * 2048, 2048, 128);
                const uint64_t bd = umma_desc(sb + ATT_SM_K + j * (ATT_KC * 16) + kk * 2 * (NTOK * 16), NTOK * 16, 128);
                if (lead) umma_f16(tmem + ATT_T_S + (G % NB) * ATT_KC, ad, bd, ATT_IDESC_S, kk > 0);
            }
            if (lead) umma_commit(BAR(AB_SFULL + (G % NB)));
            __syncwarp();
        };
        mbar_wait(BAR(AB_QFULL), 0);
        mbar_wait(BAR(AB_KFULL), 0);
        tc_fence_after();
        for (int g0 = 0; g0 < NB && g0 < NG; ++g0) issue_s(g0);
        mbar_wait(BAR(AB_VFULL), 0);
#pragma unroll 1
        for (int G = 0; G < NG; ++G) {
            const int b = G % NB, qt = G / ATT_NCH, j = G % ATT_NCH;
            const uint32_t par = (G / NB) & 1;
            mbar_wait(BAR(AB_PFULL + b), par);                   // P_j is in TMEM over S_j
            if (j == 0 && qt > 0) mbar_wait(BAR(AB_OFREE), (qt - 1) & 1);   // previous q-tile's O has been read
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < ATT_KC / 16; ++ks) {
                const uint64_t bd = umma_desc(sb + ATT_SM_V + (j * (ATT_KC / 8) + 2 * ks) * 512, 512, 128);
                if (lead) umma_f16_ts(tmem + ATT_T_O, tmem + ATT_T_S + b * ATT_KC + ks * 8, bd, ATT_IDESC_PV, (j > 0 || ks > 0) ? 1u : 0u);
            }
            if (lead) {
                umma_commit(BAR(AB_PVDONE + b));
                if (j == ATT_NCH - 1) umma_commit(BAR(AB_OFULL));
            }
            __syncwarp();
            if (G + NB < NG) {                                   // the next S into this buffer overwrites P_j
#ifndef T2S_ATT_NO_PVWAIT
                mbar_wait(BAR(AB_PVDONE + b), par);
                tc_fence_after();
#endif
                issue_s(G + NB);
            }
        }
        __syncwarp();
    } else {
        // ================================================================= softmax: thread = query row
        const int r = (warp_cta & 3) * 32 + lane;           // TMEM lane group of a warp = its index in the CTA mod 4
        const uint32_t trow = tmem + ((uint32_t)((warp_cta & 3) * 32) << 16);
        const float sc = 0.25503486f;                        // log2(e) / sqrt(32)
        const bool tr = trace != nullptr && tid == 0 && npart == 1;
#define ASTAMP(i) do { if (tr) trace[(size_t)blockIdx.x * 32 + (i)] = clock64(); } while (0)
        ASTAMP(0);
        // O / rowsum of q-tile qt -> the out-projection A-operand tile of the token kernel
        auto finish = [&](int ql, float lsum) {                 // ql = local q-tile index of this warpgroup
            const float inv = 1.f / lsum;
            mbar_wait(BAR(AB_OFULL), ql & 1);
            tc_fence_after();
            const int tok = q_tile(ql) * QT_ROWS + r;
            const int tt = tok / TILE_TOK, tilerow = (seq & 1) * 64 + (tok - tt * TILE_TOK);
            __half* dst = o + ((size_t)(seq >> 1) * TILES_PER_PAIR + tt) * (TILE_ROWS * D) + head * 4 * 1024 + tilerow * 8;
            float a0[32];
            tmem_ld32(trow + ATT_T_O, a0);
            tmem_wait_ld();
            tc_fence_before();
            mbar_arrive_warp(BAR(AB_OFREE));
            if (r < QT_ROWS) {
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8)
                {
                    float y[8];
#pragma unroll
                    for (int u = 0; u < 8; u += 2) mul2(y[u], y[u + 1], a0[c8 * 8 + u], a0[c8 * 8 + u + 1], inv, inv);
                    *reinterpret_cast<uint4*>(dst + c8 * 1024) = make_uint4(pack_h2(y[0], y[1]), pack_h2(y[2], y[3]), pack_h2(y[4], y[5]), pack_h2(y[6], y[7]));
                }
            }
        };
        float lprev = 1.f;
        float v[ATT_KC];                                         // the score chunk: 32-column loads + a 16-column remainder
        auto load_chunk = [&](int Gl) {                          // wait for S of chunk Gl and start its TMEM -> register load
            const int bl = Gl % NB;
            mbar_wait(BAR(AB_SFULL + bl), (Gl / NB) & 1);
            tc_fence_after();
            const uint32_t tl = trow + ATT_T_S + bl * ATT_KC;
#pragma unroll
            for (int c = 0; c + 32 <= ATT_KC; c += 32) tmem_ld32(tl + c, *reinterpret_cast<float (*)[32]>(&v[c]));
            if constexpr (ATT_KC % 32 == 16) tmem_ld16(tl + ATT_KC - 16, *reinterpret_cast<float (*)[16]>(&v[ATT_KC - 16]));
        };
#pragma unroll 1
        for (int qt = 0; qt < nql; ++qt) {                       // local q-tile index (q-tile q_tile(qt))
            float mref = 0.f, l0 = 0.f, l1 = 0.f;
#pragma unroll 1
            for (int j = 0; j < ATT_NCH; ++j) {
                const int G = qt * ATT_NCH + j, b = G % NB;
                const uint32_t ts = trow + ATT_T_S + b * ATT_KC;
#ifndef T2S_ATT_EARLY_LOAD
                load_chunk(G);
#else
                if (G == 0) load_chunk(0);                       // later chunks were requested at the end of their predecessor
#endif
                tmem_wait_ld();
                // exponentials of the chunk against the reference point `mr`: P packed to fp16 over the S buffer, row sums into l0 / l1
                auto exps = [&](float mr) {
                    const float nb = -mr * sc;
#pragma unroll
                    for (int hb = 0; hb < ATT_KC / 16; ++hb) {   // 16 scores -> 8 packed P columns
                        uint32_t pk[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) {             // packed fp32 (FFMA2 / FADD2): one issue slot per pair
                            float t0, t1;
                            fma2(t0, t1, v[hb * 16 + 2 * q], v[hb * 16 + 2 * q + 1], sc, sc, nb, nb);
#ifdef T2S_ATT_POLY
                            // A/B build: T2S_ATT_POLY of every 8 pairs take their exponentials on the FMA / ALU pipes instead of the MUFU:
                            // 2^t = 2^n p(f), n = round(t) by the magic-number add, f = t - n in [-0.5, 0.5], p = degree-3 minimax polynomial
                            // (7.5e-5 relative; P is rounded to fp16 afterwards), the exponent added to the bits of p.  t is clamped at -30
                            // (2^-30 of a row sum >= 1)
                            float e0, e1;
                            if (q >= 8 - T2S_ATT_POLY) {
                                const float MAGIC = 12582912.f;
                                const float c0 = fmaxf(t0, -30.f), c1 = fmaxf(t1, -30.f);
                                float z0, z1, n0, n1, f0, f1, p0, p1;
                                add2(z0, z1, c0, c1, MAGIC, MAGIC);
                                add2(n0, n1, z0, z1, -MAGIC, -MAGIC);
                                fma2(f0, f1, n0, n1, -1.f, -1.f, c0, c1);
                                fma2(p0, p1, f0, f1, 0.05517164617776871f, 0.05517164617776871f, 0.2426111251115799f, 0.2426111251115799f);
                                fma2(p0, p1, p0, p1, f0, f1, 0.6932609677314758f, 0.6932609677314758f);
                                fma2(p0, p1, p0, p1, f0, f1, 0.9999280571937561f, 0.9999280571937561f);
                                e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(z0) << 23));
                                e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(z1) << 23));
                            } else {
                                e0 = ex2_approx(t0); e1 = ex2_approx(t1);
                            }
#else
                            const float e0 = ex2_approx(t0), e1 = ex2_approx(t1);
#endif
                            add2(l0, l1, l0, l1, e0, e1);
                            pk[q] = pack_h2(e0, e1);
                        }
                        tmem_st8(ts + hb * 8, pk);
                    }
                };
                auto chunk_max = [&]() {
                    float c0 = -INFINITY, c1 = -INFINITY;
#pragma unroll
                    for (int q = 0; q < ATT_KC; q += 4) { c0 = max3(c0, v[q], v[q + 1]); c1 = max3(c1, v[q + 2], v[q + 3]); }
                    return fmaxf(c0, c1);
                };
                // the reference point moves (rare): row sum and O row are rescaled once every earlier P.V has landed in O
                auto move_reference = [&](float cm, bool need, float& la, float& lb) {
                    const float alpha = need ? ex2_approx((mref - cm) * sc) : 1.f;
                    if (need) mref = cm;
                    la *= alpha; lb *= alpha;
                    mbar_wait(BAR(AB_PVDONE + (G - 1) % NB), ((G - 1) / NB) & 1);
                    tc_fence_after();
                    float a0[32];
                    tmem_ld32(trow + ATT_T_O, a0);
                    tmem_wait_ld();
#pragma unroll
                    for (int q = 0; q < 32; ++q) a0[q] *= alpha;
                    tmem_st16(trow + ATT_T_O, *reinterpret_cast<float (*)[16]>(&a0[0]));
                    tmem_st16(trow + ATT_T_O + 16, *reinterpret_cast<float (*)[16]>(&a0[16]));
                };
#ifndef T2S_ATT_NO_SPECULATE
                // Later chunks exponentiate against the CURRENT reference point straight away and take the chunk maximum beside the
                // MUFU stream instead of in front of it; only if a row turns out to exceed the reference point by more than 2^8 (rare)
                // are the row sums restored, the reference point moved and the chunk redone.
                if (j == 0) {
                    mref = chunk_max();
                    exps(mref);
                } else {
                    const float s0 = l0, s1 = l1;
                    exps(mref);
                    const float cm = chunk_max();
                    const bool need = (cm - mref) * sc > 8.f;
                    if (__any_sync(0xffffffffu, need)) {
                        l0 = s0; l1 = s1;
                        move_reference(cm, need, l0, l1);
                        exps(mref);
                    }
                }
#else
                const float cm = chunk_max();
                if (j == 0) {
                    mref = cm;
                } else {
                    const bool need = (cm - mref) * sc > 8.f;    // P would exceed 2^8: move the reference point
                    if (__any_sync(0xffffffffu, need)) move_reference(cm, need, l0, l1);
                }
                exps(mref);
#endif
                if (j == 0 && qt > 0) finish(qt - 1, lprev);  // previous q-tile's O -> global before its accumulator is reused
#ifdef T2S_ATT_EARLY_LOAD                                      // A/B build: measured slower (the chunk registers spill), off by default
                if (G + 1 < NG) load_chunk(G + 1);            // the next chunk's TMEM load flies while this chunk's P store completes
#endif
                tmem_wait_st();
                tc_fence_before();
                mbar_arrive_warp(BAR(AB_PFULL + b));
            }
            ASTAMP(1 + qt);
            lprev = l0 + l1;
        }
        finish(nql - 1, lprev);
        ASTAMP(20);
    }
    tc_fence_before();
    __syncthreads();
    if (warp_cta == 4) tmem_dealloc(tmem, ATT_TCOLS);
}

}  // namespace t2s
