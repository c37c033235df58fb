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
    const float* pos;               // [8 tiles][32 col chunks][64 rows][4]  pos_embed in the residual tile layout
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
    float* h;              // residual stream, tiled: [npair][8 tiles][32 col chunks][128 rows][4] fp32
    __half* qkv;           // [nseq][4 heads][3][480][32] fp16, 16B chunks XOR-swizzled by (tok>>1)&3
    const __half* o;       // attention output, tiled A-operand images: [npair][8 tiles][16 K chunks][16 row groups][8][8] fp16
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

// =================================================================================== token block (tcgen05)
// One CTA = one pair tile (128 rows = 60 tokens x {sequence 2p, sequence 2p+1}, 4+4 padding rows).
// Everything that is local to a token runs here for one DiT block boundary:
//   MID  (block l):  x += gate_msa*proj(o) ; a2 = mod(LN2(x)) ; x += gate_mlp*fc2(GELU(fc1(a2))) ; store x ;
//                    a' = mod(LN1(x)) of block l+1 ; q|k|v = a' Wqkv^T + b   (transformer.py:114-117, timm Attention/Mlp)
//   EMBED:           x = patch-embed + pos ; a' of block 0 ; q|k|v
//   FINAL (block 3): ... ; final LN + Linear(128->4) + unpatchify + CFG mix + Euler/DDPM update
// Warp roles (192 threads): warps 0-3 = epilogue, one thread per tile row (= TMEM lane), the residual row
// lives in 128 registers; warp 4 = producer (bulk async copies of 32 KB weight stages and of the
// attention-output tile, completion on mbarriers); warp 5 = MMA issuer (one lane issues
// tcgen05.mma.kind::f16 128x128x16, accumulators in TMEM, completion via tcgen05.commit).
// The seven GEMM chunks of a tile (proj | fc1 a,b | fc2 K-halves | q,k,v), each 128x128x128, rotate over
// three 128-column TMEM regions so that the MMA of chunk i+1 overlaps the epilogue of chunk i.
constexpr int TC_THREADS = 192;
constexpr int TC_NSTAGE = 3;
constexpr int TC_SM_A = 0;                                       // 32 KB A operand: o tile / a2 / hidden-b / a'
constexpr int TC_SM_HA = STAGE_BYTES;                            // 32 KB A operand: hidden-a
constexpr int TC_SM_W = 2 * STAGE_BYTES;                         // weight ring
constexpr int TC_SM_VEC = TC_SM_W + TC_NSTAGE * STAGE_BYTES;     // per-tile vectors (fp32)
constexpr int V_MOD = 0;        // [2 branches][768]  adaLN chunk of block l
constexpr int V_MODN = 1536;    // [2][256]           shift_msa | scale_msa of the next block
constexpr int V_BPROJ = 2048, V_B1 = 2176, V_B2 = 2432, V_BQKV = 2560;
constexpr int V_WEMB = 2944;    // [128][4]
constexpr int V_BEMB = 3456;    // [128]
constexpr int V_WFIN = 3584;    // [4][128]
constexpr int V_BFIN = 4096;    // [4]
constexpr int V_END = 4104;
constexpr int TC_SM_VB = TC_SM_VEC + V_END * 4;                  // [128][4] fp32 final-projection exchange
constexpr int TC_SM_BAR = TC_SM_VB + TILE_ROWS * 4 * 4;
constexpr int TC_SM_TMEM = TC_SM_BAR + 32 * 8;
constexpr int TOK_SMEM_BYTES = TC_SM_TMEM + 16;
enum { B_WFULL = 0, B_WEMPTY = 3, B_OFULL = 6, B_A2 = 7, B_HA = 8, B_HB = 9, B_A3 = 10, B_ACC = 11 /* ..17 */, B_COUNT = 18 };
constexpr uint32_t TC_IDESC = umma_idesc_f16(128, 128);
constexpr uint32_t KCH = 2048;   // byte stride between K chunks (16 row groups x 128 B) in a [128][128] operand image

// one 128x128x128 GEMM chunk: 8 x tcgen05.mma (K = 16 each); operands in the canonical no-swizzle K-major image
__device__ __forceinline__ void tc_gemm(uint32_t a_smem, uint32_t w_smem, uint32_t d_tmem, bool accumulate) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
        umma_f16(d_tmem, umma_desc(a_smem + k * 2 * KCH, KCH, 128), umma_desc(w_smem + k * 2 * KCH, KCH, 128), TC_IDESC,
                 (accumulate || k > 0) ? 1u : 0u);
}

// thread-per-row LayerNorm (no affine) + modulate, packed to fp16 and stored as the A operand image.
__device__ __forceinline__ void ln_mod_store(const float (&h)[D], const float* __restrict__ shift, const float* __restrict__ scale,
                                             float eps, uint8_t* abuf, int r) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < D; ++c) s += h[c];
    const float mean = s * (1.f / D);
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < D; ++c) { const float d = h[c] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(q * (1.f / D) + eps);
#pragma unroll
    for (int c8 = 0; c8 < 16; ++c8) {
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; j += 4) {
            const float4 sc = *reinterpret_cast<const float4*>(scale + c8 * 8 + j);
            const float4 sh = *reinterpret_cast<const float4*>(shift + c8 * 8 + j);
            y[j + 0] = fmaf((h[c8 * 8 + j + 0] - mean) * rstd, 1.f + sc.x, sh.x);
            y[j + 1] = fmaf((h[c8 * 8 + j + 1] - mean) * rstd, 1.f + sc.y, sh.y);
            y[j + 2] = fmaf((h[c8 * 8 + j + 2] - mean) * rstd, 1.f + sc.z, sh.z);
            y[j + 3] = fmaf((h[c8 * 8 + j + 3] - mean) * rstd, 1.f + sc.w, sh.w);
        }
        *reinterpret_cast<uint4*>(abuf + c8 * KCH + r * 16) =
            make_uint4(pack_h2(y[0], y[1]), pack_h2(y[2], y[3]), pack_h2(y[4], y[5]), pack_h2(y[6], y[7]));
    }
}

// x += gate * (acc + bias) over the 128 columns of a TMEM region
__device__ __forceinline__ void residual_update(float (&h)[D], uint32_t taddr, const float* __restrict__ gate,
                                                const float* __restrict__ bias) {
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) {
        float v[32];
        tmem_ld32(taddr + cb * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 g4 = *reinterpret_cast<const float4*>(gate + cb * 32 + j);
            const float4 b4 = *reinterpret_cast<const float4*>(bias + cb * 32 + j);
            h[cb * 32 + j + 0] = fmaf(g4.x, v[j + 0] + b4.x, h[cb * 32 + j + 0]);
            h[cb * 32 + j + 1] = fmaf(g4.y, v[j + 1] + b4.y, h[cb * 32 + j + 1]);
            h[cb * 32 + j + 2] = fmaf(g4.z, v[j + 2] + b4.z, h[cb * 32 + j + 2]);
            h[cb * 32 + j + 3] = fmaf(g4.w, v[j + 3] + b4.w, h[cb * 32 + j + 3]);
        }
    }
}

// hidden = GELU_tanh(acc + b1) packed to fp16 into an A operand image (timm Mlp, transformer.py:99,105)
__device__ __forceinline__ void gelu_store(uint32_t taddr, const float* __restrict__ bias, uint8_t* abuf, int r) {
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) {
        float v[32];
        tmem_ld32(taddr + cb * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
            const float4 b0 = *reinterpret_cast<const float4*>(bias + cb * 32 + c8 * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(bias + cb * 32 + c8 * 8 + 4);
            const float* x = v + c8 * 8;
            *reinterpret_cast<uint4*>(abuf + (cb * 4 + c8) * KCH + r * 16) =
                make_uint4(pack_h2(gelu_tanh(x[0] + b0.x), gelu_tanh(x[1] + b0.y)), pack_h2(gelu_tanh(x[2] + b0.z), gelu_tanh(x[3] + b0.w)),
                           pack_h2(gelu_tanh(x[4] + b1.x), gelu_tanh(x[5] + b1.y)), pack_h2(gelu_tanh(x[6] + b1.z), gelu_tanh(x[7] + b1.w)));
        }
    }
}

// grid = npair * 8 tiles, block = 192
template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1) token_kernel(const TokArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pair = blockIdx.x / TILES_PER_PAIR, tt = blockIdx.x % TILES_PER_PAIR;
    const uint32_t sb = smem_u32(smem);
    const uint32_t bar0 = sb + TC_SM_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    float* vec = reinterpret_cast<float*>(smem + TC_SM_VEC);
    const int l = p.layer;                                         // block whose second half runs here (MID / FINAL)
    const int ln = (MODE == TOK_EMBED) ? 0 : l + 1;                // block whose QKV is produced here (EMBED / MID)
    constexpr int N_STAGES = (MODE == TOK_EMBED) ? 3 : (MODE == TOK_MID ? 8 : 5);

    if (tid == 0) {
        for (int i = 0; i < 7; ++i) mbar_init(BAR(i), 1);
        for (int i = B_A2; i <= B_A3; ++i) mbar_init(BAR(i), 128);
        for (int i = B_ACC; i < B_COUNT; ++i) mbar_init(BAR(i), 1);
        mbar_fence_init();
    }
    if (warp == 4) tmem_alloc(sb + TC_SM_TMEM, 512);
    // stage the per-tile vectors
    {
        const int sq0 = min(2 * pair, p.nseq - 1), sq1 = min(2 * pair + 1, p.nseq - 1);
        if (MODE != TOK_EMBED) {
            for (int i = tid; i < 2 * MOD; i += TC_THREADS)
                vec[V_MOD + i] = p.mod[((size_t)(i < MOD ? sq0 : sq1) * NLAYER + l) * MOD + (i < MOD ? i : i - MOD)];
            for (int i = tid; i < D; i += TC_THREADS) { vec[V_BPROJ + i] = p.w.b_proj[l][i]; vec[V_B2 + i] = p.w.b_fc2[l][i]; }
            for (int i = tid; i < DMLP; i += TC_THREADS) vec[V_B1 + i] = p.w.b_fc1[l][i];
        }
        if (MODE != TOK_FINAL) {
            for (int i = tid; i < 512; i += TC_THREADS)
                vec[V_MODN + i] = p.mod[((size_t)(i < 256 ? sq0 : sq1) * NLAYER + ln) * MOD + (i & 255)];
            for (int i = tid; i < 3 * D; i += TC_THREADS) vec[V_BQKV + i] = p.w.b_qkv[ln][i];
        }
        if (MODE == TOK_EMBED) {
            for (int i = tid; i < 4 * D; i += TC_THREADS) vec[V_WEMB + i] = p.w.w_embed[i];
            for (int i = tid; i < D; i += TC_THREADS) vec[V_BEMB + i] = p.w.b_embed[i];
        }
        if (MODE == TOK_FINAL) {
            for (int i = tid; i < 4 * D; i += TC_THREADS) vec[V_WFIN + i] = p.w.w_final[i];
            if (tid < 4) vec[V_BFIN + tid] = p.w.b_final[tid];
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + TC_SM_TMEM);
    const size_t tile = (size_t)pair * TILES_PER_PAIR + tt;

    if (warp == 4) {
        // ================================================================= producer
        if (lane == 0) {
            if (MODE != TOK_EMBED) {
                mbar_expect_tx(BAR(B_OFULL), STAGE_BYTES);
                bulk_g2s(sb + TC_SM_A, reinterpret_cast<const char*>(p.o) + tile * STAGE_BYTES, STAGE_BYTES, BAR(B_OFULL));
            }
            const char* src_a = reinterpret_cast<const char*>(MODE == TOK_EMBED ? p.w.w_qkv[0] : p.w.w_post[l]);
            const char* src_b = reinterpret_cast<const char*>(MODE == TOK_MID ? p.w.w_qkv[l + 1] : nullptr);
            constexpr int N_A = (MODE == TOK_EMBED) ? 3 : 5;
#pragma unroll 1
            for (int s = 0; s < N_STAGES; ++s) {
                const int slot = s % TC_NSTAGE, use = s / TC_NSTAGE;
                if (use > 0) mbar_wait(BAR(B_WEMPTY + slot), (use - 1) & 1);
                mbar_expect_tx(BAR(B_WFULL + slot), STAGE_BYTES);
                bulk_g2s(sb + TC_SM_W + slot * STAGE_BYTES, s < N_A ? src_a + (size_t)s * STAGE_BYTES : src_b + (size_t)(s - N_A) * STAGE_BYTES,
                         STAGE_BYTES, BAR(B_WFULL + slot));
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ================================================================= MMA issuer
        if (lane == 0) {
            int s = 0;
            auto chunk = [&](uint32_t a_smem, uint32_t d_col, bool accumulate, int acc_bar) {
                const int slot = s % TC_NSTAGE;
                mbar_wait(BAR(B_WFULL + slot), (s / TC_NSTAGE) & 1);
                tc_fence_after();
                tc_gemm(a_smem, sb + TC_SM_W + slot * STAGE_BYTES, tmem + d_col, accumulate);
                umma_commit(BAR(B_WEMPTY + slot));
                if (acc_bar >= 0) umma_commit(BAR(B_ACC + acc_bar));
                ++s;
            };
            if (MODE != TOK_EMBED) {
                mbar_wait(BAR(B_OFULL), 0);
                chunk(sb + TC_SM_A, 0, false, 0);                       // proj            -> R0
                mbar_wait(BAR(B_A2), 0);
                chunk(sb + TC_SM_A, 128, false, 1);                     // fc1 cols 0..127   -> R1
                chunk(sb + TC_SM_A, 256, false, 2);                     // fc1 cols 128..255 -> R2
                mbar_wait(BAR(B_HA), 0);
                chunk(sb + TC_SM_HA, 0, false, -1);                     // fc2, K half 0   -> R0
                mbar_wait(BAR(B_HB), 0);
                chunk(sb + TC_SM_A, 0, true, 3);                        // fc2, K half 1   -> R0
            }
            if (MODE != TOK_FINAL) {
                mbar_wait(BAR(B_A3), 0);
                chunk(sb + TC_SM_A, 128, false, 4);                     // q -> R1
                chunk(sb + TC_SM_A, 256, false, 5);                     // k -> R2
                chunk(sb + TC_SM_A, 0, false, 6);                       // v -> R0
            }
        }
        __syncwarp();
    } else {
        // ================================================================= epilogue: thread r <-> tile row r <-> TMEM lane r
        const int r = tid;
        const int branch = r >> 6, tl = r & 63;
        const int seq = 2 * pair + branch;
        const bool valid = tl < TILE_TOK && seq < p.nseq;
        const int tok = tt * TILE_TOK + tl;
        const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
        const float* modb = vec + V_MOD + branch * MOD;
        float* htile = p.h + tile * (TILE_ROWS * D);                    // [32 col chunks][128 rows][4]
        float h[D];
        if (MODE == TOK_EMBED) {
            // patchify + patch_emb + pos_embed (transformer.py:166-172), conv folded into the Linear
            float xv[4] = {0.f, 0.f, 0.f, 0.f};
            if (valid) {
                const float* xs = p.x + (size_t)(seq >> p.x_shift) * LAT;
                const int i = tok >> 5, j = tok & 31;
#pragma unroll
                for (int pq = 0; pq < 4; ++pq) xv[pq] = xs[(2 * j + (pq & 1)) * LATP + 2 * i + (pq >> 1)];
            }
            const float* pos = p.w.pos + ((size_t)tt * 32 * 64 + tl) * 4;   // [8 tiles][32 chunks][64 rows][4]
#pragma unroll
            for (int c4 = 0; c4 < 32; ++c4) {
                float4 pe = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) pe = *reinterpret_cast<const float4*>(pos + c4 * 64 * 4);
                const float pev[4] = {pe.x, pe.y, pe.z, pe.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 w4 = *reinterpret_cast<const float4*>(vec + V_WEMB + (c4 * 4 + e) * 4);
                    h[c4 * 4 + e] = valid ? (w4.x * xv[0] + w4.y * xv[1] + w4.z * xv[2] + w4.w * xv[3] + vec[V_BEMB + c4 * 4 + e] + pev[e]) : 0.f;
                }
            }
        } else {
#pragma unroll
            for (int c4 = 0; c4 < 32; ++c4) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) v = *reinterpret_cast<const float4*>(htile + (c4 * TILE_ROWS + r) * 4);
                h[c4 * 4 + 0] = v.x; h[c4 * 4 + 1] = v.y; h[c4 * 4 + 2] = v.z; h[c4 * 4 + 3] = v.w;
            }
            // x = x + gate_msa * (o Wproj^T + b)        (transformer.py:116)
            mbar_wait(BAR(B_ACC + 0), 0);
            tc_fence_after();
            residual_update(h, trow + 0, modb + 2 * D, vec + V_BPROJ);
            ln_mod_store(h, modb + 3 * D, modb + 4 * D, 1e-6f, smem + TC_SM_A, r);
            fence_async_smem();
            tc_fence_before();
            mbar_arrive(BAR(B_A2));
            // hidden = GELU(fc1)                         (transformer.py:117)
            mbar_wait(BAR(B_ACC + 1), 0);
            tc_fence_after();
            gelu_store(trow + 128, vec + V_B1, smem + TC_SM_HA, r);
            fence_async_smem();
            tc_fence_before();
            mbar_arrive(BAR(B_HA));
            mbar_wait(BAR(B_ACC + 2), 0);
            tc_fence_after();
            gelu_store(trow + 256, vec + V_B1 + D, smem + TC_SM_A, r);
            fence_async_smem();
            tc_fence_before();
            mbar_arrive(BAR(B_HB));
            // x = x + gate_mlp * (hidden W2^T + b)
            mbar_wait(BAR(B_ACC + 3), 0);
            tc_fence_after();
            residual_update(h, trow + 0, modb + 5 * D, vec + V_B2);
        }

        if (MODE != TOK_FINAL) {
            if (valid) {
#pragma unroll
                for (int c4 = 0; c4 < 32; ++c4)
                    *reinterpret_cast<float4*>(htile + (c4 * TILE_ROWS + r) * 4) = make_float4(h[c4 * 4], h[c4 * 4 + 1], h[c4 * 4 + 2], h[c4 * 4 + 3]);
            }
            const float* modn = vec + V_MODN + branch * 256;
            ln_mod_store(h, modn, modn + D, 1e-6f, smem + TC_SM_A, r);
            fence_async_smem();
            tc_fence_before();
            mbar_arrive(BAR(B_A3));
            // q | k | v = a' W^T + b, stored fp16 in the attention kernel's smem image layout
#pragma unroll 1
            for (int which = 0; which < 3; ++which) {
                mbar_wait(BAR(B_ACC + 4 + which), 0);
                tc_fence_after();
                const uint32_t tcol = which == 0 ? 128u : (which == 1 ? 256u : 0u);
                const float* bq = vec + V_BQKV + which * D;
                const int swz = (tok >> 1) & 3;
#pragma unroll
                for (int head = 0; head < 4; ++head) {
                    float v[32];
                    tmem_ld32(trow + tcol + head * 32, v);
                    tmem_wait_ld();
                    if (valid) {
                        __half* dst = p.qkv + ((((size_t)seq * NHEAD + head) * 3 + which) * NTOK + tok) * HD;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float4 b0 = *reinterpret_cast<const float4*>(bq + head * 32 + c * 8);
                            const float4 b1 = *reinterpret_cast<const float4*>(bq + head * 32 + c * 8 + 4);
                            const float* x = v + c * 8;
                            *reinterpret_cast<uint4*>(dst + ((c ^ swz) << 3)) =
                                make_uint4(pack_h2(x[0] + b0.x, x[1] + b0.y), pack_h2(x[2] + b0.z, x[3] + b0.w),
                                           pack_h2(x[4] + b1.x, x[5] + b1.y), pack_h2(x[6] + b1.z, x[7] + b1.w));
                        }
                    }
                }
            }
        } else {
            // final LN (eps 1e-5, affine folded) + Linear(128->4) + unpatchify (transformer.py:182-190)
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < D; ++c) s += h[c];
            const float mean = s * (1.f / D);
            float q = 0.f;
#pragma unroll
            for (int c = 0; c < D; ++c) { const float d = h[c] - mean; q = fmaf(d, d, q); }
            const float rstd = rsqrtf(q * (1.f / D) + 1e-5f);
            float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < D; c += 4) {
                const float y0 = (h[c] - mean) * rstd, y1 = (h[c + 1] - mean) * rstd, y2 = (h[c + 2] - mean) * rstd, y3 = (h[c + 3] - mean) * rstd;
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const float4 w = *reinterpret_cast<const float4*>(vec + V_WFIN + c4 * D + c);
                    d4[c4] = fmaf(y0, w.x, fmaf(y1, w.y, fmaf(y2, w.z, fmaf(y3, w.w, d4[c4]))));
                }
            }
            float* vb = reinterpret_cast<float*>(smem + TC_SM_VB);
            *reinterpret_cast<float4*>(vb + r * 4) =
                make_float4(d4[0] + vec[V_BFIN], d4[1] + vec[V_BFIN + 1], d4[2] + vec[V_BFIN + 2], d4[3] + vec[V_BFIN + 3]);
            asm volatile("bar.sync 1, 128;\n" ::: "memory");
            if (p.out_mode == OUT_FWD) {
                for (int idx = tid; idx < 2 * TILE_TOK * 4; idx += 128) {
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
                for (int idx = tid; idx < TILE_TOK * 4; idx += 128) {
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
                        xn = mean2 + p.c3 * p.noise[xi];
                    }
                    p.x_upd[xi] = xn;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem, 512);
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
        // output goes straight into the token kernel's A-operand image: [pair][tile][16 K chunks][16 row groups][8 rows][8 halves]
        const int ra = qrow0 + mt * 16 + g, rb = ra + 8;
        const int ta = ra / TILE_TOK, tb = rb / TILE_TOK;
        const int rowa = (seq & 1) * 64 + (ra - ta * TILE_TOK), rowb = (seq & 1) * 64 + (rb - tb * TILE_TOK);
        __half* da = o + ((size_t)(seq >> 1) * TILES_PER_PAIR + ta) * (TILE_ROWS * D) + (rowa >> 3) * 64 + (rowa & 7) * 8 + 2 * t;
        __half* db = o + ((size_t)(seq >> 1) * TILES_PER_PAIR + tb) * (TILE_ROWS * D) + (rowb >> 3) * 64 + (rowb & 7) * 8 + 2 * t;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            *reinterpret_cast<uint32_t*>(da + (head * 4 + d) * 1024) = pack_h2(oacc[mt][d][0] * i0, oacc[mt][d][1] * i0);
            *reinterpret_cast<uint32_t*>(db + (head * 4 + d) * 1024) = pack_h2(oacc[mt][d][2] * i1, oacc[mt][d][3] * i1);
        }
    }
}

}  // namespace t2s
