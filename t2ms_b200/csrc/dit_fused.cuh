// One guided denoising step of the T2S-DiT (reference: model/denoiser/transformer.py:158-193 x {uncond, cond}, the guidance
// mix of infer.py:81/87 and the Euler / DDPM update of rectified_flow.py:5-7 / DDPM.py:28-36) as ONE persistent kernel.
//
// Why: the attention phase of a block sits on the MUFU (softmax exponentials: XU pipe 80 % busy, tensor pipe 20 %), the
// token phase (out-projection, LayerNorm, MLP, next QKV) on dependent-issue latency, HBM and the tensor pipe (XU 12 %).
// As separate kernels they ran back to back; here they run CONCURRENTLY ON EVERY SM, on different sequence pairs:
//
//   one CTA per SM (cooperative launch, 640 threads), two independent halves that never synchronise with each other:
//     token half      warps 0-7 epilogue (two threads per tile row); warp 8 (one thread): MMA issuer that, instead of
//                     sleeping at a barrier, pumps the producer state machine (weight half stages, item inputs);
//                     one pair tile (60 tokens x {uncond, cond}) per work item, EMBED | MID(l) | FINAL per item
//     attention half  warps 10-13 / 15-18 two softmax warpgroups (thread = query row), warps 14 / 19 their MMA issuers;
//                     one (sequence, head) per work unit
//     control         warp 9 (one thread, a non-blocking state machine): token scheduler, attention scheduler, Q/K/V loader
//     (20 warps is what the register file holds at 96 registers; two spinning roles must not share a warp as divergent
//      lanes — a suspended try_wait of one holds the other up — hence the merged single-thread state machines)
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
    int uncond_shared;     // guided loops: every pair's unconditional sequence reads modulation row 0 (cond_kernel writes it once)
    long long* stats;      // optional [grid][8]: token items, token starved cycles, attention units, attention starved cycles, total
    long long* trace;      // optional [grid][FS_TRACE_ITEMS][32]: clock64 stamps of the token epilogue's phases (row 0, half 0)
};
constexpr int FS_TRACE_ITEMS = 16;

constexpr int FS_THREADS = 640;                          // 20 warps (register allocation is per 4 warps: a 21st would cost 16 registers per thread)
constexpr int FS_W_TOK = 8, FS_W_CTL = 9, FS_W_SM0 = 10;
constexpr int FS_HALF = 16384;                           // one weight half stage [64 n][128 k] fp16
constexpr int FS_NSLOT = 3;
constexpr int FS_SM_A = 0;                               // 32 KB A operand: a2 / hidden-b / a'
constexpr int FS_SM_HA = STAGE_BYTES;                    // 32 KB A operand: o tile / hidden-a
constexpr int FS_SM_W = 2 * STAGE_BYTES;                 // weight ring
constexpr int FS_SM_VEC = FS_SM_W + FS_NSLOT * FS_HALF;  // [2 buffers] adaLN rows of the item's pair
constexpr int FV_MOD = 0;                                // [2 branches][768] block l
constexpr int FV_MODN = 1536;                            // [2][256] shift_msa | 1 + scale_msa of the next block; FINAL: [128][4] projection exchange
constexpr int FV_FLOATS = 2048;
constexpr int FS_SM_ST = FS_SM_VEC + 2 * FV_FLOATS * 4;  // [2 halves][128] float2 LayerNorm statistics exchange
constexpr int FS_SM_EMB = FS_SM_ST + 2 * TILE_ROWS * 8;  // [4][128] folded patch-embed weight + [128] bias
using FSS = DitShape<30>;
constexpr int FS_SM_Q = FS_SM_EMB + 5 * D * 4;
constexpr int FS_SM_K = FS_SM_Q + FSS::Q_HALVES * 2;
constexpr int FS_SM_V = FS_SM_K + FSS::K_HALVES * 2;
constexpr int FS_SM_BAR = FS_SM_V + FSS::V_HALVES * 2;
constexpr int FS_NBAR = 92;
constexpr int FS_SM_DESC = FS_SM_BAR + FS_NBAR * 8;      // token item ring [4] int4 | attention unit ring [4] int4
constexpr int FS_SM_TMEM = FS_SM_DESC + 128;
constexpr int FS_SMEM_BYTES = FS_SM_TMEM + 16;
static_assert(FS_SMEM_BYTES <= 232448, "fused step kernel shared memory exceeds 227 KB");
static_assert(FS_SM_Q % 128 == 0 && FS_SM_BAR % 8 == 0, "alignment");

// barriers
enum {
    FB_WFULL = 0, FB_WEMPTY = 3, FB_VFULL = 6, FB_VFREE = 8,
    FB_OFULL = 12, FB_A2 = 13, FB_HA = 14 /* 2 */, FB_HB = 16, FB_A3 = 17, FB_XFREE = 18 /* 2 */, FB_DONE = 20, FB_HAFREE = 21,
    FB_ACC = 22,   /* 7 chunks x 2 halves */
    FA_QFULL = 40, FA_KFULL = 41, FA_VFULL = 42, FA_QKFREE = 43, FA_VFREE = 44,
    FA_WG = 48,    /* per warpgroup, 10 apart: SFULL 0,1 | PFULL 2,3 | PVDONE 4,5 | OFULL 6 | OFREE 7 */
    /* rings of FS_RING: descriptor published / item (unit) completely written, indexed by item (unit) % FS_RING */
    FB_DFULL = 72, FB_IDONE = 76, FA_DFULL = 80, FA_UDONE = 84 /* + 4 wg */,
};
enum { FW_SFULL = 0, FW_PFULL = 2, FW_PVDONE = 4, FW_OFULL = 6, FW_OFREE = 7 };
constexpr int FS_RING = 4;            // descriptor ring; the schedulers publish up to FS_RING - 1 items ahead of the last reported completion
constexpr int FS_KC = 48, FS_NCH = 10;                   // keys per score chunk, chunks per q-tile
constexpr int FS_NG = 2 * FS_NCH;                        // score chunks of one warpgroup in one unit (two q-tiles)
constexpr uint32_t FS_IDESC_H = umma_idesc_f16(128, 64);
constexpr uint32_t FS_IDESC_S = umma_idesc_f16(128, FS_KC);
constexpr uint32_t FS_T_S = 0, FS_T_O = 2 * FS_KC;

// one half chunk: D[128 x 64] (+)= A[128 x 128] . W_half[64 x 128]^T : 8 x tcgen05.mma M128 N64 K16 (one thread)
__device__ __forceinline__ void fs_gemm_half(uint32_t a_smem, uint32_t w_smem, uint32_t d_tmem, bool accumulate) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint64_t ad = umma_desc(a_smem + k * 2 * KCH, KCH, 128), bd = umma_desc(w_smem + k * 2 * 1024, 1024, 128);
        umma_f16(d_tmem, ad, bd, FS_IDESC_H, (accumulate || k > 0) ? 1u : 0u);
    }
}
// waits for the phase with this parity, but gives up after about `ns` nanoseconds of hardware-suspended waiting (the thread
// does not occupy issue slots while suspended, unlike a test_wait spin)
__device__ __forceinline__ bool mbar_try_wait_ns(uint32_t bar, uint32_t parity, uint32_t ns) {
    uint32_t done;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(bar), "r"(parity), "r"(ns) : "memory");
    return done != 0;
}
// wait without the per-site watchdog code of mbar_wait (the control thread of the kernel is the watchdog): a three-instruction loop
__device__ __forceinline__ void fs_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait_ns(bar, parity, 4000)) {}
}
// non-blocking: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
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
// 1 = item claimed, 0 = nothing runnable now, -1 = every token item of this launch has been claimed.
// All counters are read first (one L2 round trip), the latest runnable phase is claimed with a CAS.
__device__ __forceinline__ int fs_tok_try_pop(int* sc, int npair, int inflight, int4& d) {
    const int total = 8 * npair;
    int head[5], avail[5];
#pragma unroll
    for (int ph = 0; ph < 5; ++ph) head[ph] = ld_relaxed_gpu(sc + SC_TOK_HEAD + ph);
#pragma unroll
    for (int ph = 1; ph < 5; ++ph) avail[ph] = ld_relaxed_gpu(sc + SC_TOK_TAIL + ph - 1);
    avail[0] = inflight > 0 ? ld_relaxed_gpu(sc + SC_DONE_TILES) : 0;
#pragma unroll
    for (int ph = 1; ph < 5; ++ph) avail[ph] *= 8;
    avail[0] = inflight > 0 ? min(total, avail[0] + 8 * inflight) : total;
    bool all = true;
#pragma unroll 1
    for (int ph = 4; ph >= 0; --ph) {
        int h = head[ph];
        if (h >= total) continue;
        all = false;
        while (h < avail[ph]) {
            const int old = atom_cas_relaxed_gpu(sc + SC_TOK_HEAD + ph, h, h + 1);
            if (old == h) {
                fence_acq_rel_gpu();                      // acquire: the queue entry and everything its pusher had observed
                const int pair = ph == 0 ? (h >> 3) : ld_relaxed_gpu(sc + SC_HDR + (ph - 1) * npair + (h >> 3));
                d = make_int4(ph == 0 ? TOK_EMBED : (ph == 4 ? TOK_FINAL : TOK_MID), ph - 1, pair, h & 7);
                return 1;
            }
            h = old;
        }
    }
    return all ? -1 : 0;
}
__device__ __forceinline__ int fs_att_try_pop(int* sc, int npair, int4& d) {
    const int total = 8 * npair;
    int head[NLAYER], avail[NLAYER];
#pragma unroll
    for (int l = 0; l < NLAYER; ++l) head[l] = ld_relaxed_gpu(sc + SC_ATT_HEAD + l);
#pragma unroll
    for (int l = 0; l < NLAYER; ++l) avail[l] = ld_relaxed_gpu(sc + SC_ATT_TAIL + l);
    bool all = true;
#pragma unroll 1
    for (int l = NLAYER - 1; l >= 0; --l) {
        int h = head[l];
        if (h >= total) continue;
        all = false;
        while (h < 8 * avail[l]) {
            const int old = atom_cas_relaxed_gpu(sc + SC_ATT_HEAD + l, h, h + 1);
            if (old == h) {
                fence_acq_rel_gpu();
                const int pair = ld_relaxed_gpu(sc + SC_HDR + (4 + l) * npair + (h >> 3));
                d = make_int4(0, l, 2 * pair + ((h >> 2) & 1), h & 3);
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
// semb = [4][128] folded weight + [128] bias in shared memory.
__device__ __forceinline__ void fs_embed_row(const FusedArgs& p, const float* semb, int seq, int tok, int tt, int tl, int c0, float4 (&hq)[16]) {
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

// per-thread record of the phase parity of every barrier this thread waits on (each completion is consumed exactly once)
struct BarPhases {
    unsigned long long bits = 0ull;
    __device__ __forceinline__ void wait(uint32_t bar0, int id) {
        fs_wait(bar0 + 8u * id, (uint32_t)((bits >> id) & 1ull));
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
        mbar_init(BAR(FB_VFREE), 8); mbar_init(BAR(FB_VFREE + 1), 8);     // epilogue barriers: one arrival per WARP
        for (int i = 0; i < FS_RING; ++i) { mbar_init(BAR(FB_DFULL + i), 1); mbar_init(BAR(FB_IDONE + i), 8); mbar_init(BAR(FA_DFULL + i), 1); }
        for (int i = 0; i < 2 * FS_RING; ++i) mbar_init(BAR(FA_UDONE + i), 4);
        mbar_init(BAR(FB_OFULL), 1);
        mbar_init(BAR(FB_A2), 8);
        mbar_init(BAR(FB_HA), 4); mbar_init(BAR(FB_HA + 1), 4);
        mbar_init(BAR(FB_HB), 8);
        mbar_init(BAR(FB_A3), 8);
        mbar_init(BAR(FB_XFREE), 4); mbar_init(BAR(FB_XFREE + 1), 4);
        mbar_init(BAR(FB_DONE), 8);
        mbar_init(BAR(FB_HAFREE), 1);
        for (int i = 0; i < 14; ++i) mbar_init(BAR(FB_ACC + i), 1);
        mbar_init(BAR(FA_QFULL), 1); mbar_init(BAR(FA_KFULL), 1); mbar_init(BAR(FA_VFULL), 1);
        mbar_init(BAR(FA_QKFREE), 2); mbar_init(BAR(FA_VFREE), 2);
        for (int g = 0; g < 2; ++g) {
            const int b = FA_WG + g * 10;
            mbar_init(BAR(b + FW_SFULL), 1); mbar_init(BAR(b + FW_SFULL + 1), 1);
            mbar_init(BAR(b + FW_PFULL), 4); mbar_init(BAR(b + FW_PFULL + 1), 4);
            mbar_init(BAR(b + FW_PVDONE), 1); mbar_init(BAR(b + FW_PVDONE + 1), 1);
            mbar_init(BAR(b + FW_OFULL), 1);
            mbar_init(BAR(b + FW_OFREE), 4);
        }
        mbar_fence_init();
    }
    if (warp == FS_W_CTL) tmem_alloc(sb + FS_SM_TMEM, 512);
    {
        float* semb = reinterpret_cast<float*>(smem + FS_SM_EMB);
        for (int i = tid; i < 5 * D; i += FS_THREADS) semb[i] = i < 4 * D ? p.w.w_embed[i] : p.w.b_embed[i - 4 * D];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t*>(smem + FS_SM_TMEM), 0);

    if (warp == FS_W_TOK) {
        if (lane == 0) {
            // ============================================================= token MMA issuer + producer (ONE thread)
            // The thread walks the MMA sequence of each item; whenever it has to wait for a barrier it runs the producer state
            // machine instead of sleeping (pump): weight half stages into free ring slots, the adaLN rows / attention-output
            // tile / residual prefetch of the items the scheduler has published.  Nothing here blocks, so no other role of
            // this warp can be held up (a second spinning lane in the same warp was: try_wait suspends the whole warp).
            int p_it = 0, p_s = 0, p_gs = 0, p_state = 0;       // producer: item, stage in item, global half stage, 0 desc | 1 inputs | 2 stages | 3 end
            int p_mode = 0, p_l = 0, p_nst = 0, p_na = 0;
            const char* p_src_a = nullptr;
            const char* p_src_b = nullptr;
            int ha_expected = 0, ha_consumed = 0;               // HAFREE completions due / consumed (one per non-EMBED item)
            auto pump = [&]() {
#pragma unroll 1
                for (int guard = 0; guard < 4; ++guard) {
                    if (p_state == 0) {
                        if (!mbar_test(BAR(FB_DFULL + (p_it & 3)), (p_it >> 2) & 1)) return;
                        p_mode = tdesc[p_it & 3].x; p_l = tdesc[p_it & 3].y;
                        if (p_mode < 0) { p_state = 3; return; }
                        p_nst = p_mode == TOK_EMBED ? 6 : (p_mode == TOK_MID ? 16 : 10);
                        p_na = p_mode == TOK_EMBED ? 6 : 10;
                        p_src_a = reinterpret_cast<const char*>(p_mode == TOK_EMBED ? p.w.w_qkv_half[0] : p.w.w_post_half[p_l]);
                        p_src_b = reinterpret_cast<const char*>(p_mode == TOK_MID ? p.w.w_qkv_half[p_l + 1] : nullptr);
                        p_state = 1;
                    }
                    if (p_state == 1) {
                        // inputs of item p_it: adaLN rows into vector buffer p_it & 1 (free once the item two back has ended), the
                        // attention-output tile into HA (free once every earlier item's fc2 has read hidden-a), residual tile -> L2
                        const int vb = p_it & 1;
                        if (p_it >= 2 && !mbar_test(BAR(FB_VFREE + vb), ((p_it >> 1) - 1) & 1)) return;
                        if (p_mode != TOK_EMBED) {
                            while (ha_consumed < ha_expected && mbar_test(BAR(FB_HAFREE), ha_consumed & 1)) ++ha_consumed;
                            if (ha_consumed < ha_expected) return;
                        }
                        const int pair = tdesc[p_it & 3].z, tile = tdesc[p_it & 3].w;
                        fence_proxy_async_all();
                        const int sq0 = p.uncond_shared ? 0 : min(2 * pair, p.nseq - 1), sq1 = min(2 * pair + 1, p.nseq - 1);
                        const int ln = p_mode == TOK_EMBED ? 0 : p_l + 1;
                        const uint32_t vdst = sb + FS_SM_VEC + vb * (FV_FLOATS * 4), vbar = BAR(FB_VFULL + vb);
                        const uint32_t bytes = (p_mode != TOK_EMBED ? 2u * MOD * 4u : 0u) + (p_mode != TOK_FINAL ? 2u * 256u * 4u : 0u);
                        mbar_expect_tx(vbar, bytes);
                        if (p_mode != TOK_EMBED) {
                            bulk_g2s(vdst + FV_MOD * 4, p.mod + ((size_t)sq0 * NLAYER + p_l) * MOD, MOD * 4, vbar);
                            bulk_g2s(vdst + (FV_MOD + MOD) * 4, p.mod + ((size_t)sq1 * NLAYER + p_l) * MOD, MOD * 4, vbar);
                        }
                        if (p_mode != TOK_FINAL) {
                            bulk_g2s(vdst + FV_MODN * 4, p.mod + ((size_t)sq0 * NLAYER + ln) * MOD, 256 * 4, vbar);
                            bulk_g2s(vdst + (FV_MODN + 256) * 4, p.mod + ((size_t)sq1 * NLAYER + ln) * MOD, 256 * 4, vbar);
                        }
                        if (p_mode != TOK_EMBED) {
                            const size_t t = (size_t)pair * TILES_PER_PAIR + tile;
                            if (!(p_mode == TOK_MID && p_l == 0)) prefetch_l2(p.h + t * (TILE_ROWS * D), TILE_ROWS * D * 4);
                            mbar_expect_tx(BAR(FB_OFULL), STAGE_BYTES);
                            bulk_g2s(sb + FS_SM_HA, reinterpret_cast<const char*>(p.o) + t * STAGE_BYTES, STAGE_BYTES, BAR(FB_OFULL));
                        }
                        p_s = 0;
                        p_state = 2;
                    }
                    if (p_state == 2) {
                        const int slot = p_gs % FS_NSLOT, use = p_gs / FS_NSLOT;
                        if (use > 0 && !mbar_test(BAR(FB_WEMPTY + slot), (use - 1) & 1)) return;
                        mbar_expect_tx(BAR(FB_WFULL + slot), FS_HALF);
                        bulk_g2s(sb + FS_SM_W + slot * FS_HALF,
                                 p_s < p_na ? p_src_a + (size_t)p_s * FS_HALF : p_src_b + (size_t)(p_s - p_na) * FS_HALF, FS_HALF, BAR(FB_WFULL + slot));
                        ++p_gs;
                        if (++p_s == p_nst) {
                            if (p_mode != TOK_EMBED) ++ha_expected;
                            ++p_it;
                            p_state = 0;
                        }
                    }
                    if (p_state == 3) return;
                }
            };
            // The MMA sequence of an item as a table walked by ONE loop with ONE wait site, so that the producer state machine is
            // instantiated once (inlined at every wait it made this thread stream ~160 KB of code per item through the
            // instruction cache the compute warps share: ncu showed a 74 % hit rate and no-instruction stalls on top).
            // step = wait barrier (63: none) | A operand (0: A, 1: HA) << 6 | region (0: X, 1: Y) << 7 | column half << 8 |
            //        accumulate << 9 | accumulator barrier to commit (63: none) << 10 | commit HAFREE << 16
#define FS_STEP(wait, a, reg, half, accum, commit, hafree) \
    ((uint32_t)(wait) | ((uint32_t)(a) << 6) | ((uint32_t)(reg) << 7) | ((uint32_t)(half) << 8) | ((uint32_t)(accum) << 9) | ((uint32_t)(commit) << 10) | ((uint32_t)(hafree) << 16))
            auto step_of = [](int s) -> uint32_t {
                switch (s) {
                    case 0: return FS_STEP(FB_OFULL, 1, 0, 0, 0, FB_ACC + 0, 0);       // proj (o tile in HA) -> X
                    case 1: return FS_STEP(63, 1, 0, 1, 0, FB_ACC + 1, 0);
                    case 2: return FS_STEP(FB_A2, 0, 1, 0, 0, FB_ACC + 2, 0);          // fc1 cols 0..127 -> Y
                    case 3: return FS_STEP(63, 0, 1, 1, 0, FB_ACC + 3, 0);
                    case 4: return FS_STEP(FB_HA, 0, 1, 0, 0, 63, 0);                  // fc1 cols 128..255 -> Y, half by half as hidden-a drains it
                    case 5: return FS_STEP(FB_HA + 1, 0, 1, 1, 0, FB_ACC + 5, 0);
                    case 6: return FS_STEP(FB_HB, 1, 1, 0, 0, 63, 0);                  // fc2: both K halves per output half -> Y
                    case 7: return FS_STEP(63, 0, 1, 0, 1, FB_ACC + 6, 0);
                    case 8: return FS_STEP(63, 1, 1, 1, 0, 63, 1);                     //   ... HA has been read: it may take the next o tile
                    case 9: return FS_STEP(63, 0, 1, 1, 1, FB_ACC + 7, 0);
                    case 10: return FS_STEP(FB_A3, 0, 0, 0, 0, FB_ACC + 8, 0);         // q -> X
                    case 11: return FS_STEP(63, 0, 0, 1, 0, FB_ACC + 9, 0);
                    case 12: return FS_STEP(63, 0, 1, 0, 0, FB_ACC + 10, 0);           // k -> Y
                    case 13: return FS_STEP(63, 0, 1, 1, 0, FB_ACC + 11, 0);
                    case 14: return FS_STEP(FB_XFREE, 0, 0, 0, 0, FB_ACC + 12, 0);     // v -> X, half by half as the q epilogue drains it
                    default: return FS_STEP(FB_XFREE + 1, 0, 0, 1, 0, FB_ACC + 13, 0);
                }
            };
#undef FS_STEP
            unsigned long long bits = 0ull;                     // phase parity of the barriers this thread consumes in order
            int gs = 0, it = 0, s = 0, s_end = 0, state = 0;    // state: 0 item descriptor | 1 previous item drained | 2 step's barrier | 3 weights
            uint32_t X = 0, Y = 0, step = 0;
            long long t_last = clock64();
#pragma unroll 1
            for (;;) {
                int id;
                uint32_t par;
                if (state == 0) { id = FB_DFULL + (it & 3); par = (it >> 2) & 1; }
                else if (state == 3) { id = FB_WFULL + gs % FS_NSLOT; par = (gs / FS_NSLOT) & 1; }
                else { id = state == 1 ? FB_DONE : (int)(step & 63u); par = (uint32_t)((bits >> id) & 1ull); }
#pragma unroll 1
                for (;;) {                                      // the one wait site: asleep in try_wait, awake to keep the producer going
                    pump();
                    if (mbar_try_wait_ns(BAR(id), par, 2000)) break;
                    if (clock64() - t_last > 8000000000LL) __trap();
                }
                t_last = clock64();
                if (state == 0) {
                    const int mode = tdesc[it & 3].x;
                    if (mode < 0) break;
                    X = tmem + (it & 1) * 128; Y = tmem + 128 - (it & 1) * 128;       // the two regions swap roles every item
                    s = mode == TOK_EMBED ? 10 : 0;
                    s_end = mode == TOK_FINAL ? 10 : 16;
                    step = step_of(s);
                    state = it > 0 ? 1 : ((step & 63u) != 63u ? 2 : 3);                // item > 0: its X (= the previous Y) must be drained
                    continue;
                }
                if (state == 1 || state == 2) {
                    bits ^= 1ull << id;
                    tc_fence_after();
                    state = (state == 1 && (step & 63u) != 63u) ? 2 : 3;
                    continue;
                }
                // state 3: the half stage is in its ring slot: issue the half chunk
                tc_fence_after();
                {
                    const uint32_t a_smem = sb + (((step >> 6) & 1u) ? FS_SM_HA : FS_SM_A);
                    const uint32_t d_tmem = (((step >> 7) & 1u) ? Y : X) + 64u * ((step >> 8) & 1u);
                    fs_gemm_half(a_smem, sb + FS_SM_W + (gs % FS_NSLOT) * FS_HALF, d_tmem, ((step >> 9) & 1u) != 0u);
                    const uint32_t cm = (step >> 10) & 63u;
                    if (cm != 63u) umma_commit(BAR((int)cm));
                    if ((step >> 16) & 1u) umma_commit(BAR(FB_HAFREE));
                    umma_commit(BAR(FB_WEMPTY + gs % FS_NSLOT));
                    ++gs;
                }
                if (++s == s_end) { ++it; state = 0; }
                else { step = step_of(s); state = (step & 63u) != 63u ? 2 : 3; }
            }
        }
    } else if (warp == FS_W_CTL) {
        if (lane == 0) {
            // ============================================================= control (ONE thread, nothing blocks): the token
            // scheduler (pops up to two items ahead of the epilogue, reports finished items), the attention scheduler (same for
            // units and the two warpgroups) and the attention loader (Q | K | V of published units as the buffers free up)
            long long tok_starved = 0, att_starved = 0, tok_idle = 0, att_idle = 0;
            int tk = 0, td = 0;                                 // token items published / completions reported
            bool tok_closed = false;
            int ak = 0, ad[2] = {0, 0};                         // attention units published / completions reported per warpgroup
            bool att_closed = false;
            int lu = 0, lstate = 0;                             // loader: next unit, 0 = Q | K pending, 1 = V pending
            long long t_prog = clock64();
#pragma unroll 1
            for (;;) {
                bool progress = false;
                // ---- token scheduler
                if (!tok_closed && tk - td < FS_RING - 1) {               // desc ring / barrier phases: never two items ahead of the epilogue
                    int4 it4;
                    const int r = fs_tok_try_pop(sc, npair, p.inflight, it4);
                    if (r != 0) {
                        if (r < 0) { it4 = make_int4(-1, 0, 0, 0); tok_closed = true; }
                        tdesc[tk & 3].x = it4.x; tdesc[tk & 3].y = it4.y; tdesc[tk & 3].z = it4.z; tdesc[tk & 3].w = it4.w;
                        mbar_arrive(BAR(FB_DFULL + (tk & 3)));
                        if (r > 0) ++tk;
                        progress = true;
                        if (tok_idle) { tok_starved += clock64() - tok_idle; tok_idle = 0; }
                    } else if (tk == td && tok_idle == 0) {
                        tok_idle = clock64();                   // nothing in flight, nothing runnable: the token half is starved
                    }
                }
                if (td < tk && mbar_test(BAR(FB_IDONE + (td & 3)), (td >> 2) & 1)) {
                    const int mode = tdesc[td & 3].x, l = tdesc[td & 3].y, pair = tdesc[td & 3].z;
                    fs_tok_done(sc, npair, mode == TOK_EMBED ? 0 : (mode == TOK_FINAL ? 4 : l + 1), pair);
                    ++td;
                    progress = true;
                }
                // ---- attention scheduler
                if (!att_closed && ak - min(ad[0], ad[1]) < FS_RING - 1) {
                    int4 u4;
                    const int r = fs_att_try_pop(sc, npair, u4);
                    if (r > 0 && u4.z >= p.nseq) {              // the missing second sequence of an odd batch: nothing to compute
                        fs_att_done(sc, npair, u4.y, u4.z >> 1, 2);
                        progress = true;
                    } else if (r != 0) {
                        if (r < 0) { u4 = make_int4(-1, 0, 0, 0); att_closed = true; }
                        adesc[ak & 3].x = u4.x; adesc[ak & 3].y = u4.y; adesc[ak & 3].z = u4.z; adesc[ak & 3].w = u4.w;
                        mbar_arrive(BAR(FA_DFULL + (ak & 3)));
                        if (r > 0) ++ak;
                        progress = true;
                        if (att_idle) { att_starved += clock64() - att_idle; att_idle = 0; }
                    } else if (ak == ad[0] && ak == ad[1] && att_idle == 0) {
                        att_idle = clock64();
                    }
                }
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    if (ad[g] < ak && mbar_test(BAR(FA_UDONE + g * FS_RING + (ad[g] & 3)), (ad[g] >> 2) & 1)) {
                        fs_att_done(sc, npair, adesc[ad[g] & 3].y, adesc[ad[g] & 3].z >> 1, 1);
                        ++ad[g];
                        progress = true;
                    }
                }
                // ---- attention loader
                if (lu < ak) {
                    const int seq = adesc[lu & 3].z, head = adesc[lu & 3].w;
                    const char* src = reinterpret_cast<const char*>(p.qkv + ((size_t)seq * NHEAD + head) * HEAD_HALVES);
                    if (lstate == 0 && (lu == 0 || mbar_test(BAR(FA_QKFREE), (lu - 1) & 1))) {    // both warpgroups' last score MMAs of the previous unit are done
                        fence_proxy_async_all();
                        mbar_expect_tx(BAR(FA_QFULL), Q_HALVES * 2);
                        bulk_g2s(sb + FS_SM_Q, src, Q_HALVES * 2, BAR(FA_QFULL));
                        mbar_expect_tx(BAR(FA_KFULL), K_HALVES * 2);
                        bulk_g2s(sb + FS_SM_K, src + Q_HALVES * 2, K_HALVES * 2, BAR(FA_KFULL));
                        lstate = 1;
                        progress = true;
                    }
                    if (lstate == 1 && (lu == 0 || mbar_test(BAR(FA_VFREE), (lu - 1) & 1))) {     // ... and their last P.V MMAs
                        mbar_expect_tx(BAR(FA_VFULL), V_HALVES * 2);
                        bulk_g2s(sb + FS_SM_V, src + (Q_HALVES + K_HALVES) * 2, V_HALVES * 2, BAR(FA_VFULL));
                        lstate = 0;
                        ++lu;
                        progress = true;
                    }
                }
                if (tok_closed && td == tk && att_closed && ad[0] == ak && ad[1] == ak && lu == ak) break;
                if (progress) {
                    t_prog = clock64();
                } else {
                    nanosleep(200);                             // nothing to do: stay off the issue slots of this SM partition
                    if (clock64() - t_prog > 8000000000LL) __trap();
                }
            }
            if (p.stats) {
                p.stats[blockIdx.x * 8 + 0] = tk; p.stats[blockIdx.x * 8 + 1] = tok_starved;
                p.stats[blockIdx.x * 8 + 2] = ak; p.stats[blockIdx.x * 8 + 3] = att_starved;
            }
        }
    } else if (warp < FS_W_TOK) {
        // ================================================================= token epilogue: thread (r, hh) <-> tile row r, columns 64 hh ..
        const int hh = (warp >> 2) & 1;
        const int r = (warp & 3) * 32 + lane;
        const int branch = r >> 6, tl = r & 63;
        const int c0 = hh * 64, kc0 = hh * 8;
        const uint32_t trow0 = tmem + ((uint32_t)((warp & 3) * 32) << 16) + c0;
        uint8_t* abuf = smem + FS_SM_A;
        uint8_t* habuf = smem + FS_SM_HA;
        float2* stx = reinterpret_cast<float2*>(smem + FS_SM_ST);
        const float* semb = reinterpret_cast<const float*>(smem + FS_SM_EMB);
        BarPhases ph;
        auto accw = [&](int chunk, int half) { ph.wait(bar0, FB_ACC + 2 * chunk + half); tc_fence_after(); };
        const bool tr = p.trace != nullptr && warp == 0 && lane == 0;
#define FSTAMP(i) do { if (tr && it < FS_TRACE_ITEMS) p.trace[((size_t)blockIdx.x * FS_TRACE_ITEMS + it) * 32 + (i)] = clock64(); } while (0)
#pragma unroll 1
        for (int it = 0;; ++it) {
            FSTAMP(0);
            fs_wait(BAR(FB_DFULL + (it & 3)), (it >> 2) & 1);
            const int mode = tdesc[it & 3].x, l = tdesc[it & 3].y, pair = tdesc[it & 3].z, tt = tdesc[it & 3].w;
            if (mode < 0) break;
            if (tr && it < FS_TRACE_ITEMS) p.trace[((size_t)blockIdx.x * FS_TRACE_ITEMS + it) * 32 + 31] = mode * 16 + l + 1;
            FSTAMP(1);
            const uint32_t X = (it & 1) * 128, Y = 128 - X;
            const int seq = 2 * pair + branch;
            const bool valid = tl < TILE_TOK && seq < p.nseq;
            const int tok = tt * TILE_TOK + tl;
            float* vec = reinterpret_cast<float*>(smem + FS_SM_VEC) + (it & 1) * FV_FLOATS;
            const float* modb = vec + FV_MOD + branch * MOD;
            float* htile = p.h + ((size_t)pair * TILES_PER_PAIR + tt) * (TILE_ROWS * D);
            float* hrow = htile + (c0 / 4) * TILE_ROWS * 4 + r * 4;
            float4 hq[16];
#pragma unroll
            for (int c4 = 0; c4 < 16; ++c4) hq[c4] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
                if (mode == TOK_EMBED || (mode == TOK_MID && l == 0)) {
                    fs_embed_row(p, semb, seq, tok, tt, tl, c0, hq);
                } else {
#pragma unroll
                    for (int c4 = 0; c4 < 16; ++c4) hq[c4] = __ldcg(reinterpret_cast<const float4*>(hrow + c4 * TILE_ROWS * 4));
                }
            }
            fs_wait(BAR(FB_VFULL + (it & 1)), (it >> 1) & 1);
            FSTAMP(2);
            RowStats st;
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
                FSTAMP(3);
                HalfStats hs = resid_pass_regs<true>(trow0 + X, modb + 2 * D + c0, p.w.b_proj[l] + c0, hq);
                st = merge_stats(hs, stx, r, hh, 1, 1e-6f);
                ln_mod_store(trow0 + X, st, modb + 3 * D + c0, modb + 4 * D + c0, abuf, r, kc0);
                fence_async_smem();
                tc_fence_before();
                mbar_arrive_warp(BAR(FB_A2));
                FSTAMP(4);
                // hidden = GELU(fc1)   (transformer.py:117)
                accw(1, hh);
                FSTAMP(5);
                gelu_store<true>(trow0 + Y, p.w.b_fc1[l] + c0, habuf, r, kc0);
                fence_async_smem();
                tc_fence_before();
                mbar_arrive_warp(BAR(FB_HA + hh));
                FSTAMP(6);
                accw(2, 1);                                     // BOTH halves of fc1[128:256] have read a2 before hidden-b overwrites it
                FSTAMP(7);
                gelu_store<true>(trow0 + Y, p.w.b_fc1[l] + D + c0, abuf, r, kc0);
                fence_async_smem();
                tc_fence_before();
                mbar_arrive_warp(BAR(FB_HB));
                FSTAMP(8);
                // x = x + gate_mlp * (hidden W2^T + b): X (parked x) + gate * Y -> Y
                accw(3, hh);
                FSTAMP(9);
                hs = resid_pass_tmem<true, true>(trow0 + Y, trow0 + X, modb + 5 * D + c0, p.w.b_fc2[l] + c0, hrow, valid && mode == TOK_MID);
                st = merge_stats(hs, stx, r, hh, 1, mode == TOK_FINAL ? 1e-5f : 1e-6f);   // the barrier inside also orders fc2's last read of A
                HREG = Y;
                FSTAMP(10);
            }
            if (mode != TOK_FINAL) {
                const int ln = mode == TOK_EMBED ? 0 : l + 1;
                const float* modn = vec + FV_MODN + branch * 256;
                ln_mod_store(trow0 + HREG, st, modn + c0, modn + D + c0, abuf, r, kc0);
                fence_async_smem();
                tc_fence_before();
                mbar_arrive_warp(BAR(FB_A3));
                FSTAMP(11);
                // q | k | v = a' W^T + b, stored fp16 as the attention half's operand images
#pragma unroll 1
                for (int which = 0; which < 3; ++which) {
                    accw(4 + which, hh);
                    FSTAMP(12 + 2 * which);
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
                    FSTAMP(13 + 2 * which);
                    if (which == 0) {                            // this half of X drained: the v half chunk may overwrite it
                        tc_fence_before();
                        mbar_arrive_warp(BAR(FB_XFREE + hh));
                    } else if (which == 1) {                     // Y drained: the next item's first chunk may overwrite it
                        tc_fence_before();
                        mbar_arrive_warp(BAR(FB_DONE));
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
                mbar_arrive_warp(BAR(FB_DONE));                       // Y drained
                float* vb = vec + FV_MODN;                       // FINAL has no next block: its rows' slot is the [128][4] exchange
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
            // everything this thread produces for the item is written: the token scheduler reports the item once all 256 have
            // arrived; the vector buffer is free for the item after next
            mbar_arrive_warp(BAR(FB_IDONE + (it & 3)));
            mbar_arrive_warp(BAR(FB_VFREE + (it & 1)));
            FSTAMP(18);
        }
    } else {
        // ================================================================= attention: two warpgroups of 4 softmax warps + 1 MMA warp
        const int wg = (warp - FS_W_SM0) / 5, role = (warp - FS_W_SM0) % 5;      // role 0-3 softmax, 4 MMA issuer
        const uint32_t tw = tmem + 256 + wg * 128;
        const int wb = FA_WG + wg * 10;
        if (role == 4) {
            const bool lead = lane == 0;
            auto issue_s = [&](int G) {                          // score chunk G of the unit: q-tile (G / NCH) * 2 + wg, key chunk G % NCH
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
                fs_wait(BAR(FA_DFULL + (u & 3)), (u >> 2) & 1);
                if (adesc[u & 3].x < 0) break;
                fs_wait(BAR(FA_QFULL), u & 1);
                fs_wait(BAR(FA_KFULL), u & 1);
                if (u > 0) {                                     // the last two P.V of the previous unit have consumed the S buffers
                    fs_wait(BAR(wb + FW_PVDONE), 1);
                    fs_wait(BAR(wb + FW_PVDONE + 1), 1);
                }
                tc_fence_after();
                issue_s(0);
                issue_s(1);
                fs_wait(BAR(FA_VFULL), u & 1);
#pragma unroll 1
                for (int G = 0; G < FS_NG; ++G) {
                    const int GG = u * FS_NG + G, b = G & 1, j = G % FS_NCH, QQ = u * 2 + G / FS_NCH;
                    const uint32_t par = (GG >> 1) & 1;
                    fs_wait(BAR(wb + FW_PFULL + b), par);                  // P_j is in TMEM over S_j
                    if (j == 0 && QQ > 0) fs_wait(BAR(wb + FW_OFREE), (QQ - 1) & 1);   // the previous q-tile's O has been read
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
                        fs_wait(BAR(wb + FW_PVDONE + b), par);
                        tc_fence_after();
                        issue_s(G + 2);
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
                fs_wait(BAR(FA_DFULL + (u & 3)), (u >> 2) & 1);
                if (adesc[u & 3].x < 0) break;
                const int seq = adesc[u & 3].z, head = adesc[u & 3].w;
                auto finish = [&](int ql, float lsum) {          // O / rowsum of local q-tile ql -> the out-projection A-operand tile
                    const float inv = 1.f / lsum;
                    fs_wait(BAR(wb + FW_OFULL), ql & 1);
                    tc_fence_after();
                    const int tok = (ql * 2 + wg) * QT_ROWS + r;
                    const int tt = tok / TILE_TOK, tilerow = (seq & 1) * 64 + (tok - tt * TILE_TOK);
                    __half* dst = p.o + ((size_t)(seq >> 1) * TILES_PER_PAIR + tt) * (TILE_ROWS * D) + head * 4 * 1024 + tilerow * 8;
                    float a0[32];
                    tmem_ld32(trow + FS_T_O, a0);
                    tmem_wait_ld();
                    tc_fence_before();
                    mbar_arrive_warp(BAR(wb + FW_OFREE));
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
                        fs_wait(BAR(wb + FW_SFULL + b), (GG >> 1) & 1);
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
                                fs_wait(BAR(wb + FW_PVDONE + (b ^ 1)), ((GG - 1) >> 1) & 1);   // every earlier P.V has landed in O
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
                        mbar_arrive_warp(BAR(wb + FW_PFULL + b));
                    }
                    lprev = l0 + l1;
                }
                finish(1, lprev);
                mbar_arrive_warp(BAR(FA_UDONE + wg * FS_RING + (u & 3)));       // this thread's rows of the unit are written (the attention scheduler reports it)
            }
        }
    }
    if (tid == 0 && p.stats) p.stats[blockIdx.x * 8 + 4] = clock64() - t_start;
    tc_fence_before();
    __syncthreads();
    if (warp == FS_W_CTL) { __syncwarp(); tmem_dealloc(tmem, 512); }
}

}  // namespace t2s
