// One guided denoising step of the T2S-DiT (reference: model/denoiser/transformer.py:158-193 x {uncond, cond}, the guidance
// mix of infer.py:81/87 and the Euler / DDPM update of rectified_flow.py:5-7 / DDPM.py:28-36) as ONE persistent kernel.
//
// Why: the attention phase of a block sits on the MUFU (softmax exponentials: XU pipe 80 % busy, tensor pipe 20 %), the
// token phase (out-projection, LayerNorm, MLP, next QKV) on dependent-issue latency, HBM and the tensor pipe (XU 12 %).
// As separate kernels they ran back to back; here they run CONCURRENTLY ON EVERY SM, on different sequence pairs:
//
//   one CTA per SM (cooperative launch, 640 threads), two independent halves that never synchronise with each other:
//     token half      warps 0-7 epilogue (two threads per tile row), warp 8 lane 0 producer + scheduler, warp 9 MMA issuer:
//                     one pair tile (60 tokens x {uncond, cond}) per work item, EMBED | MID(l) | FINAL per item
//     attention half  warps 10-13 / 15-18 two softmax warpgroups (thread = query row), warps 14 / 19 their MMA issuers,
//                     warp 8 lane 16 loader + scheduler: one (sequence, head) per work unit
//     (the two scheduler threads share a warp as independently scheduled lanes: neither uses a warp-level primitive)
//   TMEM: token X | Y (2 x 128 columns), two attention warpgroups (2 x 128: two 48-column S buffers + the 32-column O)
//   smem: token A | HA operand tiles, a three-slot ring of 16 KB weight HALF stages, per-item adaLN rows; attention Q | K | V
//
// The halves are coupled through a dataflow scheduler in global memory (a few KB, L2 resident): the 8 tiles of a pair in
// token phase ph make the pair's 8 attention units of block ph runnable, the 8 units make its tiles of phase ph + 1
// runnable.  Each half pops the runnable work of the LATEST phase first (depth first), so a pair races through its four
// blocks while it is still in L2 and only `inflight` pairs are admitted at a time: q|k|v, attention output and the
// residual stream are written and re-read through L2 by neighbouring SMs instead of crossing HBM once per block, and the
// whole step is one launch (+ the conditioning kernel) instead of ten.
#pragma once
#include "dit_kernels.cuh"

namespace t2s {

// ------------------------------------------------------------------------------------------------ scheduler state (ints)
constexpr int SC_TOK_HEAD = 0;      // [5]  items claimed per token phase (E, M0, M1, M2, F)
constexpr int SC_TOK_TAIL = 8;      // [4]  pairs published to token phases 1..4
constexpr int SC_TOK_RESV = 12;     // [4]
constexpr int SC_ATT_HEAD = 16;     // [4]  units claimed per block
constexpr int SC_ATT_TAIL = 20;     // [4]  pairs published to the attention queue of block l
constexpr int SC_ATT_RESV = 24;     // [4]
constexpr int SC_DONE_TILES = 28;   //      FINAL tiles finished (admission gate)
constexpr int SC_HDR = 64;          // then tokq[4][npair] | attq[4][npair] | tok_done[npair][4] | att_done[npair][4]
__host__ __device__ inline size_t fused_sched_ints(int npair) { return SC_HDR + (size_t)16 * npair; }

struct FusedArgs {
    DitWeights w;
    const float* x;        // latents [(nseq >> x_shift)][64][30]
    int x_shift;           // 1: the two sequences of a pair share one latent (CFG)
    float* h;              // residual stream tiles [npair][8][32 col chunks][128 rows][4] fp32
    __half* qkv;           // [nseq][4 heads] x {Q, K, V operand images} fp16
    __half* o;             // attention output tiles [npair][8][16 K chunks][16 row groups][8][8] fp16
    const float* mod;      // adaLN modulation [nseq][4][768] fp32 (cond_kernel)
    int nseq;
    int out_mode;          // OUT_FWD | OUT_RF | OUT_DDPM
    float* out;            // OUT_FWD: [nseq][64][30]; else optional guided prediction [npair][64][30]
    float* x_upd;          // OUT_RF / OUT_DDPM: latent updated in place
    const float* noise;    // OUT_DDPM: step noise or NULL (in-kernel Philox)
    unsigned long long seed;
    unsigned int step;
    float cfg, c1, c2, c3;
    int* sched;            // fused_sched_ints(npair) ints, zeroed before the launch
    int inflight;          // pairs admitted and not yet finished (0 = no limit)
    long long* stats;      // optional [grid][8]: token items, token starved cycles, attention units, attention starved cycles, total
};

constexpr int FS_THREADS = 640;                          // 20 warps (register allocation is per 4 warps: a 21st would cost 16 registers per thread)
constexpr int FS_W_EPI = 0, FS_W_PROD = 8, FS_W_MMA = 9, FS_W_SM0 = 10;
constexpr int FS_HALF = 16384;                           // one weight half stage [64 n][128 k] fp16
constexpr int FS_NSLOT = 3;
constexpr int FS_SM_A = 0;                               // 32 KB A operand: a2 / hidden-b / a'
constexpr int FS_SM_HA = STAGE_BYTES;                    // 32 KB A operand: o tile / hidden-a
constexpr int FS_SM_W = 2 * STAGE_BYTES;                 // weight ring
constexpr int FS_SM_VEC = FS_SM_W + FS_NSLOT * FS_HALF;  // [2 buffers] adaLN rows of the item's pair
constexpr int FV_MOD = 0;                                // [2 branches][768] block l
constexpr int FV_MODN = 1536;                            // [2][256] shift_msa | 1 + scale_msa of the next block
constexpr int FV_FLOATS = 2048;
constexpr int FS_SM_VB = FS_SM_VEC + 2 * FV_FLOATS * 4;  // [128][4] final-projection exchange
constexpr int FS_SM_ST = FS_SM_VB + TILE_ROWS * 4 * 4;   // [2 halves][128] float2 LayerNorm statistics exchange
using FSS = DitShape<30>;
constexpr int FS_SM_Q = FS_SM_ST + 2 * TILE_ROWS * 8;
constexpr int FS_SM_K = FS_SM_Q + FSS::Q_HALVES * 2;
constexpr int FS_SM_V = FS_SM_K + FSS::K_HALVES * 2;
constexpr int FS_SM_BAR = FS_SM_V + FSS::V_HALVES * 2;
constexpr int FS_NBAR = 80;
constexpr int FS_SM_DESC = FS_SM_BAR + FS_NBAR * 8;      // token item ring [4] int4 | attention unit ring [4] int4
constexpr int FS_SM_TMEM = FS_SM_DESC + 128;
constexpr int FS_SMEM_BYTES = FS_SM_TMEM + 16;
static_assert(FS_SMEM_BYTES <= 232448, "fused step kernel shared memory exceeds 227 KB");
static_assert(FS_SM_Q % 128 == 0 && FS_SM_BAR % 8 == 0, "alignment");

// barriers
enum {
    FB_WFULL = 0, FB_WEMPTY = 3, FB_VFULL = 6, FB_VFREE = 8, FB_DFULL = 10,
    FB_OFULL = 12, FB_A2 = 13, FB_HA = 14 /* 2 */, FB_HB = 16, FB_A3 = 17, FB_XFREE = 18 /* 2 */, FB_DONE = 20, FB_HAFREE = 21,
    FB_ACC = 22,   /* 7 chunks x 2 halves */
    FA_QFULL = 40, FA_KFULL = 41, FA_VFULL = 42, FA_QKFREE = 43, FA_VFREE = 44, FA_DFULL = 45 /* 2 */,
    FA_WG = 48     /* per warpgroup, 10 apart: SFULL 0,1 | PFULL 2,3 | PVDONE 4,5 | OFULL 6 | OFREE 7 */
};
enum { FW_SFULL = 0, FW_PFULL = 2, FW_PVDONE = 4, FW_OFULL = 6, FW_OFREE = 7 };
constexpr int FS_KC = 48, FS_NCH = 10;                   // keys per score chunk, chunks per q-tile
constexpr int FS_NG = 2 * FS_NCH;                        // score chunks of one warpgroup in one unit (two q-tiles)
constexpr uint32_t FS_IDESC_H = umma_idesc_f16(128, 64);
constexpr uint32_t FS_IDESC_S = umma_idesc_f16(128, FS_KC);
constexpr uint32_t FS_T_S = 0, FS_T_O = 2 * FS_KC;

// one half chunk: D[128 x 64] (+)= A[128 x 128] . W_half[64 x 128]^T : 8 x tcgen05.mma M128 N64 K16
__device__ __forceinline__ void fs_gemm_half(uint32_t a_smem, uint32_t w_smem, uint32_t d_tmem, bool accumulate, bool lead) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint64_t ad = umma_desc(a_smem + k * 2 * KCH, KCH, 128), bd = umma_desc(w_smem + k * 2 * 1024, 1024, 128);
        if (lead) umma_f16(d_tmem, ad, bd, FS_IDESC_H, (accumulate || k > 0) ? 1u : 0u);
    }
}

// ------------------------------------------------------------------------------------------------ scheduler (one thread)
__device__ __forceinline__ void fs_push(int* resv, int* tail, int* q, int pair) {
    const int slot = atomicAdd(resv, 1);
    st_relaxed_gpu(q + slot, pair);
    long long t0 = 0;
    while (ld_relaxed_gpu(tail) != slot) {               // publish in order: the previous pusher is a few instructions away
        nanosleep(20);
        const long long now = clock64();
        if (t0 == 0) t0 = now; else if (now - t0 > 4000000000LL) __trap();
    }
    st_release_gpu(tail, slot + 1);
}
// 1 = item claimed, 0 = nothing runnable now, -1 = every token item of this launch has been claimed
__device__ __forceinline__ int fs_tok_try_pop(int* sc, int npair, int inflight, int4& d) {
    const int total = 8 * npair;
    bool all = true;
#pragma unroll 1
    for (int ph = 4; ph >= 0; --ph) {
        int* head = sc + SC_TOK_HEAD + ph;
        int h = ld_relaxed_gpu(head);
        if (h >= total) continue;
        all = false;
        int avail;
        if (ph == 0) avail = inflight > 0 ? min(total, ld_relaxed_gpu(sc + SC_DONE_TILES) + 8 * inflight) : total;
        else avail = 8 * ld_acquire_gpu(sc + SC_TOK_TAIL + ph - 1);
        while (h < avail) {
            const int old = atom_cas_relaxed_gpu(head, h, h + 1);
            if (old == h) {
                const int pair = ph == 0 ? (h >> 3) : ld_relaxed_gpu(sc + SC_HDR + (ph - 1) * npair + (h >> 3));
                d = make_int4(ph == 0 ? TOK_EMBED : (ph == 4 ? TOK_FINAL : TOK_MID), ph - 1, pair, h & 7);
                fence_acq_rel_gpu();
                return 1;
            }
            h = old;
        }
    }
    return all ? -1 : 0;
}
__device__ __forceinline__ int fs_att_try_pop(int* sc, int npair, int4& d) {
    const int total = 8 * npair;
    bool all = true;
#pragma unroll 1
    for (int l = NLAYER - 1; l >= 0; --l) {
        int* head = sc + SC_ATT_HEAD + l;
        int h = ld_relaxed_gpu(head);
        if (h >= total) continue;
        all = false;
        const int avail = 8 * ld_acquire_gpu(sc + SC_ATT_TAIL + l);
        while (h < avail) {
            const int old = atom_cas_relaxed_gpu(head, h, h + 1);
            if (old == h) {
                const int pair = ld_relaxed_gpu(sc + SC_HDR + (4 + l) * npair + (h >> 3));
                d = make_int4(0, l, 2 * pair + ((h >> 2) & 1), h & 3);
                fence_acq_rel_gpu();
                return 1;
            }
            h = old;
        }
    }
    return all ? -1 : 0;
}
// a token item of phase ph (0..4) of `pair` has written everything it produces
__device__ __forceinline__ void fs_tok_done(int* sc, int npair, int ph, int pair) {
    fence_acq_rel_gpu();
    if (ph == 4) { atomicAdd(sc + SC_DONE_TILES, 1); return; }
    if (atom_add_acq_rel_gpu(sc + SC_HDR + 8 * npair + pair * 4 + ph, 1) == 7) {
        fence_acq_rel_gpu();
        fs_push(sc + SC_ATT_RESV + ph, sc + SC_ATT_TAIL + ph, sc + SC_HDR + (4 + ph) * npair, pair);
    }
}
// `inc` of the 16 warpgroup-halves (8 units x 2 warpgroups) of block l of `pair` are done
__device__ __forceinline__ void fs_att_done(int* sc, int npair, int l, int pair, int inc) {
    fence_acq_rel_gpu();
    if (atom_add_acq_rel_gpu(sc + SC_HDR + 12 * npair + pair * 4 + l, inc) + inc == 16) {
        fence_acq_rel_gpu();
        fs_push(sc + SC_TOK_RESV + l, sc + SC_TOK_TAIL + l, sc + SC_HDR + l * npair, pair);
    }
}

// the residual stream entering block 0 (transformer.py:166-172: patchify, conv folded into patch_emb, + pos_embed): this
// thread's 64 columns [c0, c0 + 64) of token `tok`.  EMBED and the MID item of block 0 both call it: identical values.
__device__ __forceinline__ void fs_embed_row(const FusedArgs& p, int seq, int tok, int tt, int tl, int c0, float4 (&hq)[16]) {
    constexpr int LATP = 30, LAT = FSS::LAT;
    const float* xs = p.x + (size_t)(seq >> p.x_shift) * LAT;
    const int i = tok >> 5, j = tok & 31;
    float xv[4];
#pragma unroll
    for (int pq = 0; pq < 4; ++pq) xv[pq] = __ldcg(xs + (2 * j + (pq & 1)) * LATP + 2 * i + (pq >> 1));
    const float* pos = p.w.pos + ((size_t)tt * 32 * 64 + tl) * 4 + (c0 / 4) * 64 * 4;
#pragma unroll
    for (int c4 = 0; c4 < 16; ++c4) {
        const int c = c0 + c4 * 4;
        const float4 pe = __ldg(reinterpret_cast<const float4*>(pos + c4 * 64 * 4));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.w.b_embed + c));
        float y0, y1, y2, y3;
        add2(y0, y1, b4.x, b4.y, pe.x, pe.y);
        add2(y2, y3, b4.z, b4.w, pe.z, pe.w);
#pragma unroll
        for (int pq = 0; pq < 4; ++pq) {
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.w.w_embed + pq * D + c));
            fma2(y0, y1, w4.x, w4.y, xv[pq], xv[pq], y0, y1);
            fma2(y2, y3, w4.z, w4.w, xv[pq], xv[pq], y2, y3);
        }
        hq[c4] = make_float4(y0, y1, y2, y3);
    }
}

// per-thread record of the phase parity of every barrier this thread waits on (each completion is consumed exactly once)
struct BarPhases {
    unsigned long long bits = 0ull;
    __device__ __forceinline__ void wait(uint32_t bar0, int id) {
        mbar_wait(bar0 + 8u * id, (uint32_t)((bits >> id) & 1ull));
        bits ^= 1ull << id;
    }
};

__global__ void __launch_bounds__(FS_THREADS, 1) fused_step_kernel(const FusedArgs p) {
    constexpr int NTOK = FSS::NTOK, TILE_TOK = FSS::TILE_TOK, TILES_PER_PAIR = FSS::TILES_PER_PAIR, LATP = FSS::H, LAT = FSS::LAT;
    constexpr int QT_ROWS = FSS::QT_ROWS, Q_HALVES = FSS::Q_HALVES, K_HALVES = FSS::K_HALVES, V_HALVES = FSS::V_HALVES, HEAD_HALVES = FSS::HEAD_HALVES;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t sb = smem_u32(smem);
    const uint32_t bar0 = sb + FS_SM_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    const int npair = (p.nseq + 1) / 2;
    int* const sc = p.sched;
    volatile int4* tdesc = reinterpret_cast<volatile int4*>(smem + FS_SM_DESC);
    volatile int4* adesc = tdesc + 4;
    const long long t_start = clock64();

    if (tid == 0) {
        for (int i = 0; i < FB_VFREE; ++i) mbar_init(BAR(i), 1);                       // WFULL, WEMPTY, VFULL
        mbar_init(BAR(FB_VFREE), 256); mbar_init(BAR(FB_VFREE + 1), 256);
        mbar_init(BAR(FB_DFULL), 1); mbar_init(BAR(FB_DFULL + 1), 1);
        mbar_init(BAR(FB_OFULL), 1);
        mbar_init(BAR(FB_A2), 256);
        mbar_init(BAR(FB_HA), 128); mbar_init(BAR(FB_HA + 1), 128);
        mbar_init(BAR(FB_HB), 256);
        mbar_init(BAR(FB_A3), 256);
        mbar_init(BAR(FB_XFREE), 128); mbar_init(BAR(FB_XFREE + 1), 128);
        mbar_init(BAR(FB_DONE), 256);
        mbar_init(BAR(FB_HAFREE), 1);
        for (int i = 0; i < 14; ++i) mbar_init(BAR(FB_ACC + i), 1);
        mbar_init(BAR(FA_QFULL), 1); mbar_init(BAR(FA_KFULL), 1); mbar_init(BAR(FA_VFULL), 1);
        mbar_init(BAR(FA_QKFREE), 2); mbar_init(BAR(FA_VFREE), 2);
        mbar_init(BAR(FA_DFULL), 1); mbar_init(BAR(FA_DFULL + 1), 1);
        for (int g = 0; g < 2; ++g) {
            const int b = FA_WG + g * 10;
            mbar_init(BAR(b + FW_SFULL), 1); mbar_init(BAR(b + FW_SFULL + 1), 1);
            mbar_init(BAR(b + FW_PFULL), 128); mbar_init(BAR(b + FW_PFULL + 1), 128);
            mbar_init(BAR(b + FW_PVDONE), 1); mbar_init(BAR(b + FW_PVDONE + 1), 1);
            mbar_init(BAR(b + FW_OFULL), 1);
            mbar_init(BAR(b + FW_OFREE), 128);
        }
        mbar_fence_init();
    }
    if (warp == FS_W_MMA) tmem_alloc(sb + FS_SM_TMEM, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t*>(smem + FS_SM_TMEM), 0);

    if (warp == FS_W_PROD && lane == 0) {
        // ================================================================= token producer + scheduler (one thread)
        BarPhases ph;
        long long starved = 0;
        int n_items = 0;
        auto pop_blocking = [&](int4& d) {
            const long long t0 = clock64();
            int r;
            while ((r = fs_tok_try_pop(sc, npair, p.inflight, d)) == 0) {
                nanosleep(100);
                if (clock64() - t0 > 4000000000LL) __trap();
            }
            if (r < 0) d = make_int4(-1, 0, 0, 0);
            starved += clock64() - t0;
        };
        auto publish = [&](int it, const int4& d) {
            tdesc[it & 3].x = d.x; tdesc[it & 3].y = d.y; tdesc[it & 3].z = d.z; tdesc[it & 3].w = d.w;
            mbar_arrive(BAR(FB_DFULL + (it & 1)));
        };
        // inputs of item `it`: the adaLN rows of its pair into vector buffer it & 1, its attention-output tile into HA
        // (the caller has made sure HA is free), its residual tile towards L2
        auto fetch_inputs = [&](int it, const int4& d) {
            const int mode = d.x, l = d.y, pair = d.z, tile = d.w, vb = it & 1;
            if (it >= 2) mbar_wait(BAR(FB_VFREE + vb), ((it >> 1) - 1) & 1);
            fence_proxy_async_all();
            const int sq0 = min(2 * pair, p.nseq - 1), sq1 = min(2 * pair + 1, p.nseq - 1);
            const int ln = mode == TOK_EMBED ? 0 : l + 1;
            const uint32_t vdst = sb + FS_SM_VEC + vb * (FV_FLOATS * 4), vbar = BAR(FB_VFULL + vb);
            const uint32_t bytes = (mode != TOK_EMBED ? 2u * MOD * 4u : 0u) + (mode != TOK_FINAL ? 2u * 256u * 4u : 0u);
            mbar_expect_tx(vbar, bytes);
            if (mode != TOK_EMBED) {
                bulk_g2s(vdst + FV_MOD * 4, p.mod + ((size_t)sq0 * NLAYER + l) * MOD, MOD * 4, vbar);
                bulk_g2s(vdst + (FV_MOD + MOD) * 4, p.mod + ((size_t)sq1 * NLAYER + l) * MOD, MOD * 4, vbar);
            }
            if (mode != TOK_FINAL) {
                bulk_g2s(vdst + FV_MODN * 4, p.mod + ((size_t)sq0 * NLAYER + ln) * MOD, 256 * 4, vbar);
                bulk_g2s(vdst + (FV_MODN + 256) * 4, p.mod + ((size_t)sq1 * NLAYER + ln) * MOD, 256 * 4, vbar);
            }
            if (mode != TOK_EMBED) {
                const size_t t = (size_t)pair * TILES_PER_PAIR + tile;
                if (!(mode == TOK_MID && l == 0)) prefetch_l2(p.h + t * (TILE_ROWS * D), TILE_ROWS * D * 4);
                mbar_expect_tx(BAR(FB_OFULL), STAGE_BYTES);
                bulk_g2s(sb + FS_SM_HA, reinterpret_cast<const char*>(p.o) + t * STAGE_BYTES, STAGE_BYTES, BAR(FB_OFULL));
            }
        };
        int4 cur;
        pop_blocking(cur);
        publish(0, cur);
        int gs = 0;                                             // global half-stage counter (ring of FS_NSLOT slots)
        if (cur.x >= 0) {
            fetch_inputs(0, cur);
#pragma unroll 1
            for (int it = 0;; ++it) {
                ++n_items;
                const int mode = cur.x, l = cur.y;
                const int n_st = mode == TOK_EMBED ? 6 : (mode == TOK_MID ? 16 : 10);
                const int look = mode == TOK_EMBED ? 0 : 9;     // MID / FINAL: once every fc2 half stage has been queued
                const char* src_a = reinterpret_cast<const char*>(mode == TOK_EMBED ? p.w.w_qkv_half[0] : p.w.w_post_half[l]);
                const char* src_b = reinterpret_cast<const char*>(mode == TOK_MID ? p.w.w_qkv_half[l + 1] : nullptr);
                const int n_a = mode == TOK_EMBED ? 6 : 10;
                bool have_next = false;
                int4 nxt = make_int4(-1, 0, 0, 0);
#pragma unroll 1
                for (int s = 0; s < n_st; ++s, ++gs) {
                    const int slot = gs % FS_NSLOT, use = gs / FS_NSLOT;
                    if (use > 0) mbar_wait(BAR(FB_WEMPTY + slot), (use - 1) & 1);
                    mbar_expect_tx(BAR(FB_WFULL + slot), FS_HALF);
                    bulk_g2s(sb + FS_SM_W + slot * FS_HALF, s < n_a ? src_a + (size_t)s * FS_HALF : src_b + (size_t)(s - n_a) * FS_HALF,
                             FS_HALF, BAR(FB_WFULL + slot));
                    if (s == look) {
                        if (mode != TOK_EMBED) ph.wait(bar0, FB_HAFREE);      // fc2 has read hidden-a: HA may take the next o tile
                        if (fs_tok_try_pop(sc, npair, p.inflight, nxt) == 1) {
                            have_next = true;
                            publish(it + 1, nxt);
                            fetch_inputs(it + 1, nxt);
                        }
                    }
                }
                if (!have_next) {
                    pop_blocking(nxt);
                    publish(it + 1, nxt);
                    if (nxt.x >= 0) fetch_inputs(it + 1, nxt);
                }
                if (nxt.x < 0) break;
                cur = nxt;
            }
        }
        if (p.stats) { p.stats[blockIdx.x * 8 + 0] = n_items; p.stats[blockIdx.x * 8 + 1] = starved; }
    } else if (warp == FS_W_MMA) {
        // ================================================================= token MMA issuer (lane 0 issues)
        const bool lead = lane == 0;
        BarPhases ph;
        int gs = 0;
        auto wfull = [&]() { mbar_wait(BAR(FB_WFULL + gs % FS_NSLOT), (gs / FS_NSLOT) & 1); tc_fence_after(); };
        auto wslot = [&]() { return sb + FS_SM_W + (gs % FS_NSLOT) * FS_HALF; };
        auto wdone = [&]() { if (lead) umma_commit(BAR(FB_WEMPTY + gs % FS_NSLOT)); __syncwarp(); ++gs; };
        auto acc = [&](int chunk, int half) { if (lead) umma_commit(BAR(FB_ACC + 2 * chunk + half)); };
        const uint32_t A_ = sb + FS_SM_A, HA_ = sb + FS_SM_HA;
#pragma unroll 1
        for (int it = 0;; ++it) {
            mbar_wait(BAR(FB_DFULL + (it & 1)), (it >> 1) & 1);
            const int mode = tdesc[it & 3].x;
            if (mode < 0) break;
            const uint32_t X = tmem + (it & 1) * 128, Y = tmem + 128 - (it & 1) * 128;   // the two regions swap roles every item
            if (it > 0) ph.wait(bar0, FB_DONE);                 // the previous item has drained its Y = this item's X
            if (mode != TOK_EMBED) {
                ph.wait(bar0, FB_OFULL);                        // proj (o tile sits in HA) -> X
                tc_fence_after();
#pragma unroll 1
                for (int hf = 0; hf < 2; ++hf) { wfull(); fs_gemm_half(HA_, wslot(), X + 64 * hf, false, lead); acc(0, hf); wdone(); }
                ph.wait(bar0, FB_A2);                           // fc1 cols 0..127 -> Y
                tc_fence_after();
#pragma unroll 1
                for (int hf = 0; hf < 2; ++hf) { wfull(); fs_gemm_half(A_, wslot(), Y + 64 * hf, false, lead); acc(1, hf); wdone(); }
#pragma unroll 1
                for (int hf = 0; hf < 2; ++hf) {                // fc1 cols 128..255 -> Y, half by half as hidden-a drains it
                    ph.wait(bar0, FB_HA + hf);
                    tc_fence_after();
                    wfull(); fs_gemm_half(A_, wslot(), Y + 64 * hf, false, lead);
                    if (hf == 1) acc(2, 1);
                    wdone();
                }
                ph.wait(bar0, FB_HB);                           // fc2, both K halves per output half -> Y
                tc_fence_after();
#pragma unroll 1
                for (int hf = 0; hf < 2; ++hf) {
                    wfull(); fs_gemm_half(HA_, wslot(), Y + 64 * hf, false, lead);
                    if (hf == 1 && lead) umma_commit(BAR(FB_HAFREE));
                    wdone();
                    wfull(); fs_gemm_half(A_, wslot(), Y + 64 * hf, true, lead); acc(3, hf); wdone();
                }
            }
            if (mode != TOK_FINAL) {
                ph.wait(bar0, FB_A3);                           // q -> X, k -> Y
                tc_fence_after();
#pragma unroll 1
                for (int hf = 0; hf < 2; ++hf) { wfull(); fs_gemm_half(A_, wslot(), X + 64 * hf, false, lead); acc(4, hf); wdone(); }
#pragma unroll 1
                for (int hf = 0; hf < 2; ++hf) { wfull(); fs_gemm_half(A_, wslot(), Y + 64 * hf, false, lead); acc(5, hf); wdone(); }
#pragma unroll 1
                for (int hf = 0; hf < 2; ++hf) {                // v -> X, half by half as the q epilogue drains it
                    ph.wait(bar0, FB_XFREE + hf);
                    tc_fence_after();
                    wfull(); fs_gemm_half(A_, wslot(), X + 64 * hf, false, lead); acc(6, hf); wdone();
                }
            }
        }
        __syncwarp();
    } else if (warp < FS_W_PROD) {
        // ================================================================= token epilogue: thread (r, hh) <-> tile row r, columns 64 hh ..
        const int hh = (warp >> 2) & 1;
        const int r = (warp & 3) * 32 + lane;
        const int branch = r >> 6, tl = r & 63;
        const int c0 = hh * 64, kc0 = hh * 8;
        const uint32_t trow0 = tmem + ((uint32_t)((warp & 3) * 32) << 16) + c0;
        uint8_t* abuf = smem + FS_SM_A;
        uint8_t* habuf = smem + FS_SM_HA;
        float2* stx = reinterpret_cast<float2*>(smem + FS_SM_ST);
        BarPhases ph;
        auto accw = [&](int chunk, int half) { ph.wait(bar0, FB_ACC + 2 * chunk + half); tc_fence_after(); };
#pragma unroll 1
        for (int it = 0;; ++it) {
            mbar_wait(BAR(FB_DFULL + (it & 1)), (it >> 1) & 1);
            const int mode = tdesc[it & 3].x, l = tdesc[it & 3].y, pair = tdesc[it & 3].z, tt = tdesc[it & 3].w;
            if (mode < 0) break;
            const uint32_t X = (it & 1) * 128, Y = 128 - X;
            const int seq = 2 * pair + branch;
            const bool valid = tl < TILE_TOK && seq < p.nseq;
            const int tok = tt * TILE_TOK + tl;
            const float* vec = reinterpret_cast<const float*>(smem + FS_SM_VEC) + (it & 1) * FV_FLOATS;
            const float* modb = vec + FV_MOD + branch * MOD;
            float* htile = p.h + ((size_t)pair * TILES_PER_PAIR + tt) * (TILE_ROWS * D);
            float* hrow = htile + (c0 / 4) * TILE_ROWS * 4 + r * 4;
            mbar_wait(BAR(FB_VFULL + (it & 1)), (it >> 1) & 1);
            RowStats st;
            float4 hq[16];
#pragma unroll
            for (int c4 = 0; c4 < 16; ++c4) hq[c4] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
                if (mode == TOK_EMBED || (mode == TOK_MID && l == 0)) {
                    fs_embed_row(p, seq, tok, tt, tl, c0, hq);
                } else {
#pragma unroll
                    for (int c4 = 0; c4 < 16; ++c4) hq[c4] = __ldcg(reinterpret_cast<const float4*>(hrow + c4 * TILE_ROWS * 4));
                }
            }
            uint32_t HREG;
            if (mode == TOK_EMBED) {
                // the row is parked in TMEM region X (not yet an accumulator) for the LayerNorm pass
                float sum = 0.f, sq = 0.f;
                const float shift = hq[0].x;
#pragma unroll
                for (int cb = 0; cb < 4; ++cb) {
                    float a[16];
#pragma unroll
                    for (int q = 0; q < 4; ++q) { a[q * 4] = hq[cb * 4 + q].x; a[q * 4 + 1] = hq[cb * 4 + q].y; a[q * 4 + 2] = hq[cb * 4 + q].z; a[q * 4 + 3] = hq[cb * 4 + q].w; }
                    block_stats(a, shift, sum, sq);
                    tmem_st16(trow0 + X + cb * 16, a);
                }
                tmem_wait_st();
                st = merge_stats(half_stats(shift, sum, sq), stx, r, hh, 1, 1e-6f);
                HREG = X;
            } else {
                // x = x + gate_msa * (o Wproj^T + b)   (transformer.py:116); x stays parked in X until the MLP branch adds to it
                accw(0, hh);
                HalfStats hs = resid_pass_regs<true>(trow0 + X, modb + 2 * D + c0, p.w.b_proj[l] + c0, hq);
                st = merge_stats(hs, stx, r, hh, 1, 1e-6f);
                ln_mod_store(trow0 + X, st, modb + 3 * D + c0, modb + 4 * D + c0, abuf, r, kc0);
                fence_async_smem();
                tc_fence_before();
                mbar_arrive(BAR(FB_A2));
                // hidden = GELU(fc1)   (transformer.py:117)
                accw(1, hh);
                gelu_store<true>(trow0 + Y, p.w.b_fc1[l] + c0, habuf, r, kc0);
                fence_async_smem();
                tc_fence_before();
                mbar_arrive(BAR(FB_HA + hh));
                accw(2, 1);                                     // BOTH halves of fc1[128:256] have read a2 before hidden-b overwrites it
                gelu_store<true>(trow0 + Y, p.w.b_fc1[l] + D + c0, abuf, r, kc0);
                fence_async_smem();
                tc_fence_before();
                mbar_arrive(BAR(FB_HB));
                // x = x + gate_mlp * (hidden W2^T + b): X (parked x) + gate * Y -> Y
                accw(3, hh);
                hs = resid_pass_tmem<true, true>(trow0 + Y, trow0 + X, modb + 5 * D + c0, p.w.b_fc2[l] + c0, hrow, valid && mode == TOK_MID);
                st = merge_stats(hs, stx, r, hh, 1, mode == TOK_FINAL ? 1e-5f : 1e-6f);   // the barrier inside also orders fc2's last read of A
                HREG = Y;
            }
            if (mode != TOK_FINAL) {
                const int ln = mode == TOK_EMBED ? 0 : l + 1;
                const float* modn = vec + FV_MODN + branch * 256;
                ln_mod_store(trow0 + HREG, st, modn + c0, modn + D + c0, abuf, r, kc0);
                fence_async_smem();
                tc_fence_before();
                mbar_arrive(BAR(FB_A3));
                // q | k | v = a' W^T + b, stored fp16 as the attention half's operand images
#pragma unroll 1
                for (int which = 0; which < 3; ++which) {
                    accw(4 + which, hh);
                    const uint32_t tcol = which == 1 ? Y : X;
                    const float* bq = p.w.b_qkv[ln] + which * D + c0;
                    const int off = which == 0 ? (tok / QT_ROWS) * 4096 + (tok % QT_ROWS) * 8
                                  : which == 1 ? Q_HALVES + tok * 8
                                               : Q_HALVES + K_HALVES + (tok >> 3) * 256 + (tok & 7) * 8;
                    const int dstride = which == 0 ? 1024 : (which == 1 ? NTOK * 8 : 64);
                    __half* hb0 = p.qkv + ((size_t)seq * NHEAD + hh * 2) * HEAD_HALVES + off;
                    for_blocks16<4>(trow0 + tcol, [&](int cb, float (&v)[16]) {
                        if (valid) {
                            __half* hb = hb0 + (cb >> 1) * HEAD_HALVES + (cb & 1) * 2 * dstride;
#pragma unroll
                            for (int c = 0; c < 2; ++c) {
                                const float4 b0 = __ldg(reinterpret_cast<const float4*>(bq + cb * 16 + c * 8));
                                const float4 b1 = __ldg(reinterpret_cast<const float4*>(bq + cb * 16 + c * 8 + 4));
                                const float* x = v + c * 8;
                                float y0, y1, y2, y3, y4, y5, y6, y7;
                                add2(y0, y1, x[0], x[1], b0.x, b0.y); add2(y2, y3, x[2], x[3], b0.z, b0.w);
                                add2(y4, y5, x[4], x[5], b1.x, b1.y); add2(y6, y7, x[6], x[7], b1.z, b1.w);
                                *reinterpret_cast<uint4*>(hb + c * dstride) = make_uint4(pack_h2(y0, y1), pack_h2(y2, y3), pack_h2(y4, y5), pack_h2(y6, y7));
                            }
                        }
                    });
                    if (which == 0) {                            // this half of X drained: the v half chunk may overwrite it
                        tc_fence_before();
                        mbar_arrive(BAR(FB_XFREE + hh));
                    } else if (which == 1) {                     // Y drained: the next item's first chunk may overwrite it
                        tc_fence_before();
                        mbar_arrive(BAR(FB_DONE));
                    }
                }
            } else {
                // final LN (eps 1e-5, affine folded) + Linear(128->4) + unpatchify (transformer.py:182-190)
                float d4[4] = {0.f, 0.f, 0.f, 0.f};
                for_blocks16<4>(trow0 + HREG, [&](int cb, float (&a)[16]) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        float y0, y1, y2, y3;
                        add2(y0, y1, a[j], a[j + 1], -st.mean, -st.mean); add2(y2, y3, a[j + 2], a[j + 3], -st.mean, -st.mean);
                        mul2(y0, y1, y0, y1, st.rstd, st.rstd); mul2(y2, y3, y2, y3, st.rstd, st.rstd);
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) {
                            const float4 w = __ldg(reinterpret_cast<const float4*>(p.w.w_final + c4 * D + c0 + cb * 16 + j));
                            d4[c4] = fmaf(y0, w.x, fmaf(y1, w.y, fmaf(y2, w.z, fmaf(y3, w.w, d4[c4]))));
                        }
                    }
                });
                tc_fence_before();
                mbar_arrive(BAR(FB_DONE));                       // Y drained
                float* vb = reinterpret_cast<float*>(smem + FS_SM_VB);
                float4* px = reinterpret_cast<float4*>(stx);     // the statistics exchange is idle now: partial sums of half 1
                named_bar_sync(1, 256);                          // ... once every thread has read its merge partner
                if (hh == 1) px[r] = make_float4(d4[0], d4[1], d4[2], d4[3]);
                named_bar_sync(1, 256);
                if (hh == 0) {
                    const float4 o4 = px[r];
                    const float4 bf = __ldg(reinterpret_cast<const float4*>(p.w.b_final));
                    *reinterpret_cast<float4*>(vb + r * 4) = make_float4(d4[0] + o4.x + bf.x, d4[1] + o4.y + bf.y, d4[2] + o4.z + bf.z, d4[3] + o4.w + bf.w);
                }
                named_bar_sync(1, 256);
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
                        const float xo = __ldcg(p.x_upd + xi);
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
            // everything this item produces is written: tell the scheduler, release the vector buffer
            named_bar_sync(1, 256);
            if (warp == 0 && lane == 0) fs_tok_done(sc, npair, mode == TOK_EMBED ? 0 : (mode == TOK_FINAL ? 4 : l + 1), pair);
            mbar_arrive(BAR(FB_VFREE + (it & 1)));
        }
    } else if (warp == FS_W_PROD) {
        // ================================================================= attention loader + scheduler (lane 16 of the producer warp)
        if (lane == 16) {
            long long starved = 0;
            int n_units = 0;
            int4 d;
            auto pop_blocking = [&]() {
                const long long t0 = clock64();
                for (;;) {
                    const int r = fs_att_try_pop(sc, npair, d);
                    if (r < 0) { d = make_int4(-1, 0, 0, 0); break; }
                    if (r == 1) {
                        if (d.z < p.nseq) break;
                        fs_att_done(sc, npair, d.y, d.z >> 1, 2);   // the missing second sequence of an odd batch: nothing to compute
                        continue;
                    }
                    nanosleep(100);
                    if (clock64() - t0 > 4000000000LL) __trap();
                }
                starved += clock64() - t0;
            };
            pop_blocking();
#pragma unroll 1
            for (int u = 0;; ++u) {
                adesc[u & 3].x = d.x; adesc[u & 3].y = d.y; adesc[u & 3].z = d.z; adesc[u & 3].w = d.w;
                mbar_arrive(BAR(FA_DFULL + (u & 1)));
                if (d.x < 0) break;
                ++n_units;
                const char* src = reinterpret_cast<const char*>(p.qkv + ((size_t)d.z * NHEAD + d.w) * HEAD_HALVES);
                if (u > 0) mbar_wait(BAR(FA_QKFREE), (u - 1) & 1);      // both warpgroups' last score MMAs of the previous unit are done
                fence_proxy_async_all();
                mbar_expect_tx(BAR(FA_QFULL), Q_HALVES * 2);
                bulk_g2s(sb + FS_SM_Q, src, Q_HALVES * 2, BAR(FA_QFULL));
                mbar_expect_tx(BAR(FA_KFULL), K_HALVES * 2);
                bulk_g2s(sb + FS_SM_K, src + Q_HALVES * 2, K_HALVES * 2, BAR(FA_KFULL));
                if (u > 0) mbar_wait(BAR(FA_VFREE), (u - 1) & 1);       // ... and their last P.V MMAs
                mbar_expect_tx(BAR(FA_VFULL), V_HALVES * 2);
                bulk_g2s(sb + FS_SM_V, src + (Q_HALVES + K_HALVES) * 2, V_HALVES * 2, BAR(FA_VFULL));
                pop_blocking();                                         // the next unit, while this one runs
            }
            if (p.stats) { p.stats[blockIdx.x * 8 + 2] = n_units; p.stats[blockIdx.x * 8 + 3] = starved; }
        }
    } else {
        // ================================================================= attention: two warpgroups of 4 softmax warps + 1 MMA warp
        const int wg = (warp - FS_W_SM0) / 5, role = (warp - FS_W_SM0) % 5;      // role 0-3 softmax, 4 MMA issuer
        const uint32_t tw = tmem + 256 + wg * 128;
        const int wb = FA_WG + wg * 10;
        if (role == 4) {
            const bool lead = lane == 0;
            auto issue_s = [&](int u, int G) {                   // score chunk G of unit u: q-tile (G / NCH) * 2 + wg, key chunk G % NCH
                const int qt = (G / FS_NCH) * 2 + wg, j = G % FS_NCH;
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const uint64_t ad = umma_desc(sb + FS_SM_Q + qt * 8192 + kk * 2 * 2048, 2048, 128);
                    const uint64_t bd = umma_desc(sb + FS_SM_K + j * (FS_KC * 16) + kk * 2 * (NTOK * 16), NTOK * 16, 128);
                    if (lead) umma_f16(tw + FS_T_S + (G & 1) * FS_KC, ad, bd, FS_IDESC_S, kk > 0);
                }
                if (lead) umma_commit(BAR(wb + FW_SFULL + (G & 1)));
                __syncwarp();
            };
#pragma unroll 1
            for (int u = 0;; ++u) {
                mbar_wait(BAR(FA_DFULL + (u & 1)), (u >> 1) & 1);
                if (adesc[u & 3].x < 0) break;
                mbar_wait(BAR(FA_QFULL), u & 1);
                mbar_wait(BAR(FA_KFULL), u & 1);
                if (u > 0) {                                     // the last two P.V of the previous unit have consumed the S buffers
                    mbar_wait(BAR(wb + FW_PVDONE), 1);
                    mbar_wait(BAR(wb + FW_PVDONE + 1), 1);
                }
                tc_fence_after();
                issue_s(u, 0);
                issue_s(u, 1);
                mbar_wait(BAR(FA_VFULL), u & 1);
#pragma unroll 1
                for (int G = 0; G < FS_NG; ++G) {
                    const int GG = u * FS_NG + G, b = G & 1, j = G % FS_NCH, QQ = u * 2 + G / FS_NCH;
                    const uint32_t par = (GG >> 1) & 1;
                    mbar_wait(BAR(wb + FW_PFULL + b), par);                  // P_j is in TMEM over S_j
                    if (j == 0 && QQ > 0) mbar_wait(BAR(wb + FW_OFREE), (QQ - 1) & 1);   // the previous q-tile's O has been read
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < FS_KC / 16; ++ks) {
                        const uint64_t bd = umma_desc(sb + FS_SM_V + (j * (FS_KC / 8) + 2 * ks) * 512, 512, 128);
                        if (lead) umma_f16_ts(tw + FS_T_O, tw + FS_T_S + b * FS_KC + ks * 8, bd, ATT_IDESC_PV, (j > 0 || ks > 0) ? 1u : 0u);
                    }
                    if (lead) {
                        umma_commit(BAR(wb + FW_PVDONE + b));
                        if (j == FS_NCH - 1) umma_commit(BAR(wb + FW_OFULL));
                        if (G == FS_NG - 1) umma_commit(BAR(FA_VFREE));
                    }
                    __syncwarp();
                    if (G + 2 < FS_NG) {                         // the next S into this buffer overwrites P_j
                        mbar_wait(BAR(wb + FW_PVDONE + b), par);
                        tc_fence_after();
                        issue_s(u, G + 2);
                        if (G + 2 == FS_NG - 1 && lead) umma_commit(BAR(FA_QKFREE));
                        __syncwarp();
                    }
                }
            }
            __syncwarp();
        } else {
            // ---- softmax: thread = query row (attn_kernel's single-pass scheme, see dit_kernels.cuh)
            const int r = (warp & 3) * 32 + lane;
            const uint32_t trow = tw + ((uint32_t)((warp & 3) * 32) << 16);
            const float scl = 0.25503486f;                       // log2(e) / sqrt(32)
#pragma unroll 1
            for (int u = 0;; ++u) {
                mbar_wait(BAR(FA_DFULL + (u & 1)), (u >> 1) & 1);
                if (adesc[u & 3].x < 0) break;
                const int l = adesc[u & 3].y, seq = adesc[u & 3].z, head = adesc[u & 3].w;
                auto finish = [&](int ql, float lsum) {          // O / rowsum of local q-tile ql -> the out-projection A-operand tile
                    const float inv = 1.f / lsum;
                    mbar_wait(BAR(wb + FW_OFULL), ql & 1);
                    tc_fence_after();
                    const int tok = (ql * 2 + wg) * QT_ROWS + r;
                    const int tt = tok / TILE_TOK, tilerow = (seq & 1) * 64 + (tok - tt * TILE_TOK);
                    __half* dst = p.o + ((size_t)(seq >> 1) * TILES_PER_PAIR + tt) * (TILE_ROWS * D) + head * 4 * 1024 + tilerow * 8;
                    float a0[32];
                    tmem_ld32(trow + FS_T_O, a0);
                    tmem_wait_ld();
                    tc_fence_before();
                    mbar_arrive(BAR(wb + FW_OFREE));
                    if (r < QT_ROWS) {
#pragma unroll
                        for (int c8 = 0; c8 < 4; ++c8) {
                            float y[8];
#pragma unroll
                            for (int q = 0; q < 8; q += 2) mul2(y[q], y[q + 1], a0[c8 * 8 + q], a0[c8 * 8 + q + 1], inv, inv);
                            *reinterpret_cast<uint4*>(dst + c8 * 1024) = make_uint4(pack_h2(y[0], y[1]), pack_h2(y[2], y[3]), pack_h2(y[4], y[5]), pack_h2(y[6], y[7]));
                        }
                    }
                };
                float lprev = 1.f;
#pragma unroll 1
                for (int ql = 0; ql < 2; ++ql) {
                    float mref = 0.f, l0 = 0.f, l1 = 0.f;
#pragma unroll 1
                    for (int j = 0; j < FS_NCH; ++j) {
                        const int G = ql * FS_NCH + j, GG = u * FS_NG + G, b = G & 1;
                        mbar_wait(BAR(wb + FW_SFULL + b), (GG >> 1) & 1);
                        tc_fence_after();
                        const uint32_t ts = trow + FS_T_S + b * FS_KC;
                        float v[FS_KC];
                        tmem_ld32(ts, *reinterpret_cast<float (*)[32]>(&v[0]));
                        tmem_ld16(ts + 32, *reinterpret_cast<float (*)[16]>(&v[32]));
                        tmem_wait_ld();
                        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
                        for (int q = 0; q < FS_KC; q += 4) { m0 = max3(m0, v[q], v[q + 1]); m1 = max3(m1, v[q + 2], v[q + 3]); }
                        const float cm = fmaxf(m0, m1);
                        if (j == 0) {
                            mref = cm;
                        } else {
                            const bool need = (cm - mref) * scl > 8.f;   // P would exceed 2^8: move the reference point
                            if (__any_sync(0xffffffffu, need)) {
                                const float alpha = need ? ex2_approx((mref - cm) * scl) : 1.f;
                                if (need) mref = cm;
                                l0 *= alpha; l1 *= alpha;
                                mbar_wait(BAR(wb + FW_PVDONE + (b ^ 1)), ((GG - 1) >> 1) & 1);   // every earlier P.V has landed in O
                                tc_fence_after();
                                float a0[32];
                                tmem_ld32(trow + FS_T_O, a0);
                                tmem_wait_ld();
#pragma unroll
                                for (int q = 0; q < 32; ++q) a0[q] *= alpha;
                                tmem_st16(trow + FS_T_O, *reinterpret_cast<float (*)[16]>(&a0[0]));
                                tmem_st16(trow + FS_T_O + 16, *reinterpret_cast<float (*)[16]>(&a0[16]));
                            }
                        }
                        const float nb = -mref * scl;
#pragma unroll
                        for (int hb = 0; hb < FS_KC / 16; ++hb) {
                            uint32_t pk[8];
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                float t0, t1;
                                fma2(t0, t1, v[hb * 16 + 2 * q], v[hb * 16 + 2 * q + 1], scl, scl, nb, nb);
                                const float e0 = ex2_approx(t0), e1 = ex2_approx(t1);
                                add2(l0, l1, l0, l1, e0, e1);
                                pk[q] = pack_h2(e0, e1);
                            }
                            tmem_st8(ts + hb * 8, pk);
                        }
                        if (j == 0 && ql > 0) finish(ql - 1, lprev);
                        tmem_wait_st();
                        tc_fence_before();
                        mbar_arrive(BAR(wb + FW_PFULL + b));
                    }
                    lprev = l0 + l1;
                }
                finish(1, lprev);
                // this warpgroup's rows of the unit are written: tell the scheduler
                named_bar_sync(2 + wg, 128);
                if (role == 0 && lane == 0) fs_att_done(sc, npair, l, seq >> 1, 1);
            }
        }
    }
    if (tid == 0 && p.stats) p.stats[blockIdx.x * 8 + 4] = clock64() - t_start;
    tc_fence_before();
    __syncthreads();
    if (warp == FS_W_MMA) tmem_dealloc(tmem, 512);
}

}  // namespace t2s
