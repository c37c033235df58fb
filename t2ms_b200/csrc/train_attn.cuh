// Fused attention of the training step on tcgen05 / TMEM: forward with saved log-sum-exp, and the exact backward
// (dQ, dK, dV) without ever materialising the 480 x 480 score matrices (reference: timm Attention ->
// F.scaled_dot_product_attention inside model/denoiser/transformer.py:116, differentiated by autograd at
// train.py:83-85).
//
// Operand precision: fp16 operands / fp32 accumulation (10-bit mantissa, the same operand precision as the tf32
// GEMMs of the other training Linears).  Range is handled explicitly: probabilities are kept relative to a reference
// point (forward: P <= 2^8, backward: P' = 16 P <= 16), dO is scaled per (sequence, head) by a power of two so that
// its largest element lies in [0.5, 1); all scale factors are powers of two and are undone in the fp32 epilogue.
//
// Data layout.  Every operand (Q, K, V, dO of one (sequence, head): 480 tokens x 32 features) is ONE 30 KB fp16
// "T8 image" [token/8][feature/8][token%8][8]: 128-byte core matrices of 8 tokens x 8 features.  The same bytes serve
//   * as a K-major tcgen05 operand (rows = tokens, contraction over the 32 features): LBO 128 B, SBO 512 B; any
//     8-aligned window of tokens is a contiguous slice (a 120-token tile = 7680 B = one bulk copy);
//   * as an MN-major B operand (N = 32 features, contraction over tokens): LBO 512 B, SBO 128 B.
// So no transposed copies exist anywhere: the backward recomputes the scores once per orientation instead
// (S = Q K^T with thread = query row for dQ; S^T = K Q^T with thread = key row for dK / dV), which keeps every
// accumulation a plain "A from TMEM x B from shared memory" tcgen05.mma.
//
// One kernel, three modes; CTA = (sequence, head, 120-token tile) with M = 128 (8 padding rows), looping over chunks of
// the other side (FWD: five of 96 tokens, DQ / DKV: ten of 48), double-buffered in TMEM; warps 0-3: thread = tile row =
// TMEM lane; warp 4: loads + MMA issue; 80 KB of shared memory and 256 TMEM columns, so two CTAs share an SM.
//   FWD: S = Q_t K_j^T ; P = exp2((S - m) c) -> TMEM (over S) ; O += P V_j ; O / l and the log-sum-exp are stored
//   DQ : S = Q_t K_j^T, dP = dO_t V_j^T ; dS' = P' (dP - D) -> TMEM (over dP) ; dQ += dS' K_j
//   DKV: S^T = K_t Q_j^T, dP^T = V_t dO_j^T ; P'^T -> TMEM (over S^T), dS'^T -> TMEM (over dP^T) ;
//        dV += P'^T dO_j ; dK += dS'^T Q_j
// with P' = 16 exp2(S c - lse2), c = log2(e) / sqrt(32), D = rowsum(dO . O).
#pragma once
#include "common.cuh"

namespace t2s {

constexpr int TA_THREADS = 160;
constexpr float TA_PSHIFT = 4.f;                         // backward probabilities are kept as 2^4 P
enum TaMode { TA_FWD = 0, TA_DQ = 1, TA_DKV = 2 };

// Shape of the training attention for a latent of H positions (16 H tokens; H = 30: T2S, 50 / 64: the fork's
// Transformer(dim)): tile rows, chunk widths and the shared-memory map.  H = 30 fits two CTAs per SM; the wide shapes keep
// two full images of 51 / 64 KB resident and run one CTA per SM.
template <int H>
struct TaShape {
    static constexpr int NTOK = 16 * H;
    static constexpr int IMG_HALVES = NTOK * HD;             // one T8 image
    static constexpr int IMG_BYTES = IMG_HALVES * 2;
    static constexpr int ROWS = H == 64 ? 128 : 120;         // valid rows per tile (a multiple of 8); the last tile may be partial
    static constexpr int NTILE = (NTOK + ROWS - 1) / ROWS;   // 4 | 7 | 8
    static constexpr int KC_FWD = H == 30 ? 96 : (H == 50 ? 80 : 64);
    static constexpr int KC_BWD = H == 30 ? 48 : 32;
    static constexpr int SM_A0 = 0, SM_A1 = 8192;            // two 128-row operand tiles
    static constexpr int SM_B0 = 16384, SM_B1 = SM_B0 + IMG_BYTES;   // two full images
    static constexpr int SM_VEC = SM_B1 + IMG_BYTES;         // DKV: [2][NTOK] fp32 (4 - lse2 | D) of the query side
    static constexpr int SM_BAR = SM_VEC + 2 * NTOK * 4;
    static constexpr int SM_TMEM = SM_BAR + 8 * 8;
    static constexpr int SMEM_BYTES = SM_TMEM + 16;
    static constexpr int CTAS_PER_SM = 2 * (SMEM_BYTES + 1024) <= 233472 ? 2 : 1;
    static_assert(SMEM_BYTES <= 232448 && NTOK % KC_FWD == 0 && NTOK % KC_BWD == 0 && ROWS % 8 == 0 && (NTOK % ROWS) % 8 == 0, "training attention shape");
};
static_assert(TaShape<30>::CTAS_PER_SM == 2, "two training-attention CTAs must fit one SM for the T2S shape");
enum { TB_LOADED = 0, TB_SFULL = 1, TB_PFULL = 3, TB_ACC = 5 };       // SFULL / PFULL / ACC: one barrier per score buffer
constexpr uint32_t TA_BUF = 96, TA_T_ACC0 = 192, TA_T_ACC1 = 224, TA_TCOLS = 256;   // two score buffers of 96 columns + accumulators
constexpr uint32_t TA_IDESC_ACC = umma_idesc_f16(128, HD) | (1u << 16);     // B operand MN-major

struct TaArgs {
    const __half* img;     // [nseq][4 heads][3: q, k, v] T8 images
    const __half* doimg;   // [nseq][4 heads] T8 image of the scaled dO                 (DQ, DKV)
    float* o;              // FWD out: attention output [T][128] fp32, head h in columns 32 h ..
    float* nlse;           // FWD out / DQ, DKV in: [nseq][4][480]  4 - log2-sum-exp of the scaled scores
    const float* dvec;     // [nseq][4][480] rowsum(dO_scaled . O)                        (DQ, DKV)
    const float* dinv;     // [nseq][4] 1 / (dO scale)                                   (DQ, DKV)
    float* dqkv;           // DQ, DKV out: [T][384] fp32 gradient of the q | k | v rows
};

__device__ __forceinline__ uint32_t pack_h2_sat(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// ----------------------------------------------------------------------------------------------------------------
// q | k | v rows [T][384] fp32 -> T8 fp16 images.  One warp = 8 consecutive tokens of one (which, head): lane =
// (feature chunk, token % 8) reads 32 contiguous bytes and writes one 16-byte core-matrix row; a warp writes 512
// contiguous bytes.
__global__ void __launch_bounds__(256) ta_pack_qkv_kernel(const float* __restrict__ qkv, __half* __restrict__ img, int nseq, int ntok) {
    const int NTOK = ntok, TA_IMG_HALVES = ntok * HD;
    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long long nw = (long long)nseq * (NTOK / 8) * 12;
    if (w >= nw) return;
    const int sel = (int)(w % 12), which = sel >> 2, head = sel & 3;
    const long long tg = w / 12;                               // global 8-token group
    const int seq = (int)(tg / (NTOK / 8)), g = (int)(tg % (NTOK / 8));
    const int dc = lane >> 3, t8 = lane & 7;
    const float* src = qkv + ((size_t)seq * NTOK + g * 8 + t8) * (3 * D) + which * D + head * HD + dc * 8;
    const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
    __half* dst = img + (((size_t)seq * NHEAD + head) * 3 + which) * TA_IMG_HALVES + g * 256 + lane * 8;
    *reinterpret_cast<uint4*>(dst) = make_uint4(pack_h2(a.x, a.y), pack_h2(a.z, a.w), pack_h2(b.x, b.y), pack_h2(b.z, b.w));
}

// dO [T][128] fp32 (+ O) -> per (sequence, head): power-of-two scale s with max |s dO| in [0.5, 1), the T8 image of
// s dO, D = rowsum(s dO . O) and 1 / s.   grid = nseq * 4, block = 512 (16 warps x up to NG 8-token groups each).
template <int NG>
__global__ void __launch_bounds__(512) ta_pack_do_kernel(const float* __restrict__ dout, const float* __restrict__ o,
                                                         __half* __restrict__ doimg, float* __restrict__ dvec, float* __restrict__ dinv, int ntok) {
    const int NTOK = ntok, TA_IMG_HALVES = ntok * HD;
    __shared__ float red[16];
    const int seq = blockIdx.x >> 2, head = blockIdx.x & 3;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int dc = lane >> 3, t8 = lane & 7;
    float4 v[NG][2];
    float dsum[NG];
    float amax = 0.f;
#pragma unroll
    for (int i = 0; i < NG; ++i) {
        const int g = warp + 16 * i;
        dsum[i] = 0.f;
        v[i][0] = v[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g < NTOK / 8) {
            const size_t off = ((size_t)seq * NTOK + g * 8 + t8) * D + head * HD + dc * 8;
            v[i][0] = *reinterpret_cast<const float4*>(dout + off);
            v[i][1] = *reinterpret_cast<const float4*>(dout + off + 4);
            const float4 o0 = *reinterpret_cast<const float4*>(o + off), o1 = *reinterpret_cast<const float4*>(o + off + 4);
            dsum[i] = v[i][0].x * o0.x + v[i][0].y * o0.y + v[i][0].z * o0.z + v[i][0].w * o0.w +
                      v[i][1].x * o1.x + v[i][1].y * o1.y + v[i][1].z * o1.z + v[i][1].w * o1.w;
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v[i][0].x), fabsf(v[i][0].y)), fmaxf(fabsf(v[i][0].z), fabsf(v[i][0].w))));
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v[i][1].x), fabsf(v[i][1].y)), fmaxf(fabsf(v[i][1].z), fabsf(v[i][1].w))));
        }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, m));
    if (lane == 0) red[warp] = amax;
    __syncthreads();
    amax = red[0];
#pragma unroll
    for (int i = 1; i < 16; ++i) amax = fmaxf(amax, red[i]);
    int e = 0;
    if (amax > 0.f && amax < INFINITY) frexpf(amax, &e);       // amax = m 2^e, m in [0.5, 1)
    const float s = ldexpf(1.f, -e);
    if (threadIdx.x == 0) dinv[blockIdx.x] = ldexpf(1.f, e);
#pragma unroll
    for (int i = 0; i < NG; ++i) {
        const int g = warp + 16 * i;
        float dsm = dsum[i];
        dsm += __shfl_xor_sync(0xffffffffu, dsm, 8);
        dsm += __shfl_xor_sync(0xffffffffu, dsm, 16);
        if (g < NTOK / 8) {
            __half* dst = doimg + (size_t)blockIdx.x * TA_IMG_HALVES + g * 256 + lane * 8;
            *reinterpret_cast<uint4*>(dst) = make_uint4(pack_h2(v[i][0].x * s, v[i][0].y * s), pack_h2(v[i][0].z * s, v[i][0].w * s),
                                                        pack_h2(v[i][1].x * s, v[i][1].y * s), pack_h2(v[i][1].z * s, v[i][1].w * s));
            if (dc == 0) dvec[(size_t)blockIdx.x * NTOK + g * 8 + t8] = dsm * s;
        }
    }
}

// ----------------------------------------------------------------------------------------------------------------
// grid = nseq * 4 heads * NTILE tiles, block = 160.  Score chunks are double-buffered in TMEM: while the row threads work on
// chunk j (buffer j & 1) the tensor pipe has already produced the scores of chunk j + 1 and runs the accumulation of
// chunk j - 1, so the row threads (MUFU / issue bound) never wait for an MMA in steady state.
//   FWD     : chunks of KC = 96 | 80 | 64 keys (H = 30 | 50 | 64);  TMEM  S0 [0,KC) | S1 [96,96+KC) | O [192,224)
//   DQ / DKV: chunks of KC = 48 | 32 | 32;   TMEM  buffer b: S [96 b, +KC) | dP [96 b + KC, +KC) ; acc0 [192,224) | acc1 [224,256)
template <int MODE, int H = 30>
__global__ void __launch_bounds__(TA_THREADS, TaShape<H>::CTAS_PER_SM) ta_attn_kernel(const TaArgs p) {
    using TS = TaShape<H>;
    constexpr int NTOK = TS::NTOK, TA_IMG_BYTES = TS::IMG_BYTES, TA_ROWS = TS::ROWS, TA_NTILE = TS::NTILE;
    constexpr int TA_SM_A0 = TS::SM_A0, TA_SM_A1 = TS::SM_A1, TA_SM_B0 = TS::SM_B0, TA_SM_B1 = TS::SM_B1, TA_SM_VEC = TS::SM_VEC;
    constexpr int TA_SM_BAR = TS::SM_BAR, TA_SM_TMEM = TS::SM_TMEM;
    constexpr int KC = MODE == TA_FWD ? TS::KC_FWD : TS::KC_BWD;    // tokens per chunk
    constexpr int NCH = NTOK / KC;
    constexpr uint32_t IDESC_S = umma_idesc_f16(128, KC);
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform
    const int sh = blockIdx.x / TA_NTILE, tile = blockIdx.x - sh * TA_NTILE;   // (sequence, head) index; tile of the row side
    const int seq = sh >> 2, head = sh & 3;
    const int nrow = min(TA_ROWS, NTOK - tile * TA_ROWS);           // valid rows of this tile (the last one may be partial)
    const uint32_t TA_TILE_BYTES = (uint32_t)nrow * HD * 2;
    const uint32_t sb = smem_u32(smem);
    const uint32_t bar0 = sb + TA_SM_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    if (tid == 0) {
        mbar_init(BAR(TB_LOADED), 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(BAR(TB_SFULL + b), 1);
            mbar_init(BAR(TB_PFULL + b), 128);
            mbar_init(BAR(TB_ACC + b), 1);
        }
        mbar_fence_init();
    }
    if (warp == 4) tmem_alloc(sb + TA_SM_TMEM, TA_TCOLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t*>(smem + TA_SM_TMEM), 0);
    const float sc = 0.25503486f;                                   // log2(e) / sqrt(32)

    if (warp == 4) {
        // ================================================================= loads + MMA issue; whole warp converged
        const bool lead = lane == 0;
        if (lead) {
            const char* q = reinterpret_cast<const char*>(p.img) + (size_t)sh * 3 * TA_IMG_BYTES;
            const char* k = q + TA_IMG_BYTES;
            const char* v = k + TA_IMG_BYTES;
            const size_t toff = (size_t)tile * TA_ROWS * HD * 2;
            if (MODE == TA_FWD) {
                mbar_expect_tx(BAR(TB_LOADED), TA_TILE_BYTES + 2 * TA_IMG_BYTES);
                bulk_g2s(sb + TA_SM_A0, q + toff, TA_TILE_BYTES, BAR(TB_LOADED));
                bulk_g2s(sb + TA_SM_B0, k, TA_IMG_BYTES, BAR(TB_LOADED));
                bulk_g2s(sb + TA_SM_B1, v, TA_IMG_BYTES, BAR(TB_LOADED));
            } else {
                const char* d = reinterpret_cast<const char*>(p.doimg) + (size_t)sh * TA_IMG_BYTES;
                if (MODE == TA_DQ) {
                    mbar_expect_tx(BAR(TB_LOADED), 2 * TA_TILE_BYTES + 2 * TA_IMG_BYTES);
                    bulk_g2s(sb + TA_SM_A0, q + toff, TA_TILE_BYTES, BAR(TB_LOADED));
                    bulk_g2s(sb + TA_SM_A1, d + toff, TA_TILE_BYTES, BAR(TB_LOADED));
                    bulk_g2s(sb + TA_SM_B0, k, TA_IMG_BYTES, BAR(TB_LOADED));
                    bulk_g2s(sb + TA_SM_B1, v, TA_IMG_BYTES, BAR(TB_LOADED));
                } else {
                    mbar_expect_tx(BAR(TB_LOADED), 2 * TA_TILE_BYTES + 2 * TA_IMG_BYTES + 2 * NTOK * 4);
                    bulk_g2s(sb + TA_SM_A0, k + toff, TA_TILE_BYTES, BAR(TB_LOADED));
                    bulk_g2s(sb + TA_SM_A1, v + toff, TA_TILE_BYTES, BAR(TB_LOADED));
                    bulk_g2s(sb + TA_SM_B0, q, TA_IMG_BYTES, BAR(TB_LOADED));
                    bulk_g2s(sb + TA_SM_B1, d, TA_IMG_BYTES, BAR(TB_LOADED));
                    bulk_g2s(sb + TA_SM_VEC, p.nlse + (size_t)sh * NTOK, NTOK * 4, BAR(TB_LOADED));
                    bulk_g2s(sb + TA_SM_VEC + NTOK * 4, p.dvec + (size_t)sh * NTOK, NTOK * 4, BAR(TB_LOADED));
                }
            }
        }
        __syncwarp();
        // scores of chunk G into buffer G & 1 (contraction over the 32 features: two K = 16 steps)
        auto issue_scores = [&](int G) {
            const uint32_t tb = tmem + (G & 1) * TA_BUF;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                const uint64_t a0 = umma_desc(sb + TA_SM_A0 + kk * 256, 128, 512);
                const uint64_t b0 = umma_desc(sb + TA_SM_B0 + G * (KC / 8) * 512 + kk * 256, 128, 512);
                if (lead) umma_f16(tb, a0, b0, IDESC_S, kk > 0);
            }
            if (MODE != TA_FWD) {
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const uint64_t a1 = umma_desc(sb + TA_SM_A1 + kk * 256, 128, 512);
                    const uint64_t b1 = umma_desc(sb + TA_SM_B1 + G * (KC / 8) * 512 + kk * 256, 128, 512);
                    if (lead) umma_f16(tb + KC, a1, b1, IDESC_S, kk > 0);
                }
            }
            if (lead) umma_commit(BAR(TB_SFULL + (G & 1)));
            __syncwarp();
        };
        mbar_wait(BAR(TB_LOADED), 0);
        tc_fence_after();
        issue_scores(0);
        issue_scores(1);
#pragma unroll 1
        for (int j = 0; j < NCH; ++j) {
            const int b = j & 1;
            const uint32_t par = (j >> 1) & 1, tb = tmem + b * TA_BUF;
            mbar_wait(BAR(TB_PFULL + b), par);                       // the fp16 A operands of chunk j are in TMEM
            tc_fence_after();
            // accumulations over the KC tokens of the chunk (K = 16 per step), B operands MN-major
#pragma unroll
            for (int ks = 0; ks < KC / 16; ++ks) {
                const uint32_t boff = (j * (KC / 8) + 2 * ks) * 512;
                const uint32_t acc = (j > 0 || ks > 0) ? 1u : 0u;
                if (MODE == TA_FWD) {                                // O += P V
                    const uint64_t bd = umma_desc(sb + TA_SM_B1 + boff, 512, 128);
                    if (lead) umma_f16_ts(tmem + TA_T_ACC0, tb + ks * 8, bd, TA_IDESC_ACC, acc);
                } else if (MODE == TA_DQ) {                          // dQ += dS K
                    const uint64_t bd = umma_desc(sb + TA_SM_B0 + boff, 512, 128);
                    if (lead) umma_f16_ts(tmem + TA_T_ACC0, tb + KC + ks * 8, bd, TA_IDESC_ACC, acc);
                } else {                                             // dV += P^T dO ; dK += dS^T Q
                    const uint64_t bd = umma_desc(sb + TA_SM_B1 + boff, 512, 128), bq = umma_desc(sb + TA_SM_B0 + boff, 512, 128);
                    if (lead) {
                        umma_f16_ts(tmem + TA_T_ACC0, tb + ks * 8, bd, TA_IDESC_ACC, acc);
                        umma_f16_ts(tmem + TA_T_ACC1, tb + KC + ks * 8, bq, TA_IDESC_ACC, acc);
                    }
                }
            }
            if (lead) umma_commit(BAR(TB_ACC + b));
            __syncwarp();
            if (j + 2 < NCH) {                                       // the next scores into this buffer overwrite its A operands
                mbar_wait(BAR(TB_ACC + b), par);
                tc_fence_after();
                issue_scores(j + 2);
            }
        }
    } else {
        // ================================================================= thread = tile row
        const int r = tid;
        const int tok = tile * TA_ROWS + r;
        const bool valid = r < nrow;
        const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
        constexpr int LAST = NCH - 1;                                // the final accumulation: barrier ACC[LAST & 1], phase (LAST >> 1) & 1
        if (MODE == TA_FWD) {
            float mref = 0.f, l0 = 0.f, l1 = 0.f;
#pragma unroll 1
            for (int j = 0; j < NCH; ++j) {
                const int b = j & 1;
                const uint32_t ts = trow + b * TA_BUF;
                mbar_wait(BAR(TB_SFULL + b), (j >> 1) & 1);
                tc_fence_after();
                constexpr int ZN = KC - 64;                      // scores beyond the first two 32-column blocks: 32, 16 or 0
                float x[32], y[32], z[ZN > 0 ? ZN : 1];
                tmem_ld32(ts, x); tmem_ld32(ts + 32, y);
                if constexpr (ZN == 32) tmem_ld32(ts + 64, *reinterpret_cast<float (*)[32]>(&z[0]));
                if constexpr (ZN == 16) tmem_ld16(ts + 64, *reinterpret_cast<float (*)[16]>(&z[0]));
                tmem_wait_ld();
                float c0 = -INFINITY, c1 = -INFINITY;
#pragma unroll
                for (int q = 0; q < 32; q += 4) {
                    c0 = max3(c0, x[q], x[q + 1]); c1 = max3(c1, x[q + 2], x[q + 3]);
                    c0 = max3(c0, y[q], y[q + 1]); c1 = max3(c1, y[q + 2], y[q + 3]);
                    if (q < ZN) { c0 = max3(c0, z[q], z[q + 1]); c1 = max3(c1, z[q + 2], z[q + 3]); }
                }
                const float cm = fmaxf(c0, c1);
                if (j == 0) {
                    mref = cm;
                } else {
                    // move the reference point only when P would exceed 2^8 (exact in fp16 below that)
                    const bool need = (cm - mref) * sc > 8.f;
                    if (__any_sync(0xffffffffu, need)) {
                        const float alpha = need ? ex2_approx((mref - cm) * sc) : 1.f;
                        if (need) mref = cm;
                        l0 *= alpha; l1 *= alpha;
                        mbar_wait(BAR(TB_ACC + (b ^ 1)), ((j - 1) >> 1) & 1);     // every earlier P.V has landed in O
                        tc_fence_after();
                        float a0[32];
                        tmem_ld32(trow + TA_T_ACC0, a0);
                        tmem_wait_ld();
#pragma unroll
                        for (int q = 0; q < 32; ++q) a0[q] *= alpha;
                        tmem_st16(trow + TA_T_ACC0, *reinterpret_cast<float (*)[16]>(&a0[0]));
                        tmem_st16(trow + TA_T_ACC0 + 16, *reinterpret_cast<float (*)[16]>(&a0[16]));
                    }
                }
                const float nb = -mref * sc;
                auto half = [&](const float* v, int pcol) {   // 16 scores -> 8 packed P columns
                    uint32_t pk[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {                 // packed fp32 (FFMA2 / FADD2): one issue slot per pair
                        float t0, t1;
                        fma2(t0, t1, v[2 * q], v[2 * q + 1], sc, sc, nb, nb);
                        const float e0 = ex2_approx(t0), e1 = ex2_approx(t1);
                        add2(l0, l1, l0, l1, e0, e1);
                        pk[q] = pack_h2(e0, e1);
                    }
                    tmem_st8(ts + pcol, pk);
                };
                half(x, 0); half(x + 16, 8); half(y, 16); half(y + 16, 24);
                if constexpr (ZN >= 16) half(z, 32);
                if constexpr (ZN == 32) half(z + 16, 40);
                tmem_wait_st();
                tc_fence_before();
                mbar_arrive(BAR(TB_PFULL + b));
            }
            mbar_wait(BAR(TB_ACC + (LAST & 1)), (LAST >> 1) & 1);
            tc_fence_after();
            const float l = l0 + l1, inv = 1.f / l;
            float a0[32];
            tmem_ld32(trow + TA_T_ACC0, a0);
            tmem_wait_ld();
            if (valid) {
                float* dst = p.o + ((size_t)seq * NTOK + tok) * D + head * HD;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    *reinterpret_cast<float4*>(dst + q * 4) = make_float4(a0[q * 4] * inv, a0[q * 4 + 1] * inv, a0[q * 4 + 2] * inv, a0[q * 4 + 3] * inv);
                p.nlse[(size_t)sh * NTOK + tok] = TA_PSHIFT - (mref * sc + log2f(l));
            }
        } else {
            const float nl_r = (MODE == TA_DQ && valid) ? p.nlse[(size_t)sh * NTOK + tok] : 0.f;
            const float d_r = (MODE == TA_DQ && valid) ? p.dvec[(size_t)sh * NTOK + tok] : 0.f;
            const float* vnl = reinterpret_cast<const float*>(smem + TA_SM_VEC);
            const float* vd = vnl + NTOK;
            if (MODE == TA_DKV) mbar_wait(BAR(TB_LOADED), 0);           // the per-query vectors are in shared memory
#pragma unroll 1
            for (int j = 0; j < NCH; ++j) {
                const int b = j & 1;
                const uint32_t ts = trow + b * TA_BUF, tdp = ts + KC;
                mbar_wait(BAR(TB_SFULL + b), (j >> 1) & 1);
                tc_fence_after();
                float s[KC], g[KC];
                tmem_ld32(ts, *reinterpret_cast<float (*)[32]>(&s[0]));
                if constexpr (KC == 48) tmem_ld16(ts + 32, *reinterpret_cast<float (*)[16]>(&s[32]));
                tmem_ld32(tdp, *reinterpret_cast<float (*)[32]>(&g[0]));
                if constexpr (KC == 48) tmem_ld16(tdp + 32, *reinterpret_cast<float (*)[16]>(&g[32]));
                tmem_wait_ld();
#pragma unroll
                for (int h = 0; h < KC / 16; ++h) {                     // 16 scores -> 8 packed columns
                    uint32_t pp[8], pd[8];
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        float nl[4], dd[4];
                        if (MODE == TA_DKV) {
                            const int c = j * KC + h * 16 + q4 * 4;
                            const float4 n4 = *reinterpret_cast<const float4*>(vnl + c), d4 = *reinterpret_cast<const float4*>(vd + c);
                            nl[0] = n4.x; nl[1] = n4.y; nl[2] = n4.z; nl[3] = n4.w;
                            dd[0] = d4.x; dd[1] = d4.y; dd[2] = d4.z; dd[3] = d4.w;
                        } else {
                            nl[0] = nl[1] = nl[2] = nl[3] = nl_r;
                            dd[0] = dd[1] = dd[2] = dd[3] = d_r;
                        }
                        float pv[4], dv[4];
#pragma unroll
                        for (int u = 0; u < 4; u += 2) {             // packed fp32: scale, difference and product per pair
                            const int i = h * 16 + q4 * 4 + u;
                            float t0, t1, d0, d1;
                            fma2(t0, t1, s[i], s[i + 1], sc, sc, nl[u], nl[u + 1]);
                            pv[u] = ex2_approx(t0); pv[u + 1] = ex2_approx(t1);
                            add2(d0, d1, g[i], g[i + 1], -dd[u], -dd[u + 1]);
                            mul2(dv[u], dv[u + 1], pv[u], pv[u + 1], d0, d1);
                        }
                        pp[q4 * 2] = pack_h2(pv[0], pv[1]); pp[q4 * 2 + 1] = pack_h2(pv[2], pv[3]);
                        pd[q4 * 2] = pack_h2_sat(dv[0], dv[1]); pd[q4 * 2 + 1] = pack_h2_sat(dv[2], dv[3]);
                    }
                    if (MODE == TA_DKV) tmem_st8(ts + h * 8, pp);
                    tmem_st8(tdp + h * 8, pd);
                }
                tmem_wait_st();
                tc_fence_before();
                mbar_arrive(BAR(TB_PFULL + b));
            }
            mbar_wait(BAR(TB_ACC + (LAST & 1)), (LAST >> 1) & 1);
            tc_fence_after();
            const float kscale = 0.17677669529663687f;                  // 1 / sqrt(32)
            const float base = p.dinv[sh] * 0.0625f;                     // undo the dO scale and the 2^4 of P'
            float a0[32];
            tmem_ld32(trow + TA_T_ACC0, a0);
            tmem_wait_ld();
            float* drow = p.dqkv + ((size_t)seq * NTOK + tok) * (3 * D) + head * HD;
            if (valid) {
                const float f = MODE == TA_DQ ? base * kscale : base;
                float* dst = drow + (MODE == TA_DQ ? 0 : 2 * D);
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    *reinterpret_cast<float4*>(dst + q * 4) = make_float4(a0[q * 4] * f, a0[q * 4 + 1] * f, a0[q * 4 + 2] * f, a0[q * 4 + 3] * f);
            }
            if (MODE == TA_DKV) {
                tmem_ld32(trow + TA_T_ACC1, a0);
                tmem_wait_ld();
                if (valid) {
                    const float f = base * kscale;
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        *reinterpret_cast<float4*>(drow + D + q * 4) = make_float4(a0[q * 4] * f, a0[q * 4 + 1] * f, a0[q * 4 + 2] * f, a0[q * 4 + 3] * f);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem, TA_TCOLS);
}

}  // namespace t2s
