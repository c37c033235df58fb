// C ABI of the T2S B200 hot path: host-side launch logic (see include/t2s_b200.h).
#include <cstdio>
#include <cstring>

#include "../../include/t2s_b200.h"
#include "dit_kernels.cuh"
#include "dit_fused.cuh"
#include "vae_kernels.cuh"
#include "eval_kernels.cuh"
#include "backbone_kernels.cuh"

using namespace t2s;

static_assert(sizeof(DitWeights) == sizeof(t2s_dit_weights), "DitWeights must mirror t2s_dit_weights");
static_assert(sizeof(VaeDecWeights) == sizeof(t2s_vae_dec_weights), "VaeDecWeights must mirror t2s_vae_dec_weights");
static_assert(sizeof(VaeEncWeights) == sizeof(t2s_vae_enc_weights), "VaeEncWeights must mirror t2s_vae_enc_weights");

#include "api_common.h"

namespace t2s_api {
thread_local char g_err[512] = "";
int fail(int code, const char* fmt, const char* a, const char* b) {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}
}  // namespace t2s_api
using namespace t2s_api;

namespace {
long long* g_trace = nullptr;   // t2s_debug_set_phase_trace
long long* g_fused_stats = nullptr;   // t2s_debug_set_fused_stats
long long* g_fused_trace = nullptr;   // t2s_debug_set_fused_trace
// fused per-step kernel (dit_fused.cuh): used for the T2S shape from this many sequence pairs on; -1 = never (the default:
// measured slower than the per-phase kernels at every batch size; t2s_set_fused switches it on)
int g_fused_min_pairs = -1;
int g_fused_inflight = 0;       // pairs admitted and not yet finished (0 = no limit)
constexpr int MAX_DEV = 64;
bool g_inited[MAX_DEV] = {};
int g_sms[MAX_DEV] = {};
}  // namespace

int t2s_api::ensure_init() {
    int dev = 0;
    CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEV) return fail(T2S_EINVAL, "device index out of range%s%s");
    if (g_inited[dev]) return T2S_OK;
    int major = 0;
    CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) return fail(T2S_EARCH, "t2s_b200 is built for sm_100a only%s%s");
#define T2S_SET_ATTRS(HH)                                                                                                          \
    CUDA_OK(cudaFuncSetAttribute(token_kernel<TOK_EMBED, HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, TOK_SMEM_BYTES));       \
    CUDA_OK(cudaFuncSetAttribute(token_kernel<TOK_MID, HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, TOK_SMEM_BYTES));         \
    CUDA_OK(cudaFuncSetAttribute(token_kernel<TOK_FINAL, HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, TOK_SMEM_BYTES));       \
    CUDA_OK(cudaFuncSetAttribute(token_kernel<TOK_EMBED, HH, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TOK_SMEM_BYTES));    \
    CUDA_OK(cudaFuncSetAttribute(token_kernel<TOK_MID, HH, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TOK_SMEM_BYTES));      \
    CUDA_OK(cudaFuncSetAttribute(token_kernel<TOK_FINAL, HH, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TOK_SMEM_BYTES));    \
    CUDA_OK(cudaFuncSetAttribute(attn_kernel<HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttShape<HH>::SMEM_BYTES));       \
    CUDA_OK(cudaFuncSetAttribute(attn_kernel<HH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttShape<HH, true>::SMEM_BYTES));
    T2S_SET_ATTRS(30)
    T2S_SET_ATTRS(50)
    T2S_SET_ATTRS(64)
#undef T2S_SET_ATTRS
    CUDA_OK(cudaFuncSetAttribute(fused_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM_BYTES));
    const int dec = VAE_DEC_SMEM_FLOATS * 4, enc = VAE_ENC_SMEM_FLOATS * 4;
    CUDA_OK(cudaFuncSetAttribute(vae_decode_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, dec));
    CUDA_OK(cudaFuncSetAttribute(vae_decode_kernel<6, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, dec));
    CUDA_OK(cudaFuncSetAttribute(vae_decode_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, dec));
    CUDA_OK(cudaFuncSetAttribute(vae_decode_kernel<12, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, dec));
    CUDA_OK(cudaFuncSetAttribute(vae_decode_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, dec));
    CUDA_OK(cudaFuncSetAttribute(vae_decode_kernel<24, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, dec));
    CUDA_OK(cudaFuncSetAttribute(vae_encode_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, enc));
    CUDA_OK(cudaFuncSetAttribute(vae_encode_kernel<6, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, enc));
    CUDA_OK(cudaFuncSetAttribute(vae_encode_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, enc));
    CUDA_OK(cudaFuncSetAttribute(vae_encode_kernel<12, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, enc));
    CUDA_OK(cudaFuncSetAttribute(vae_encode_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, enc));
    CUDA_OK(cudaFuncSetAttribute(vae_encode_kernel<24, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, enc));
    CUDA_OK(cudaDeviceGetAttribute(&g_sms[dev], cudaDevAttrMultiProcessorCount, dev));
    g_inited[dev] = true;
    return T2S_OK;
}

namespace {

// Launch with programmatic stream serialization (PDL, common.cuh: pdl_wait): the kernels of a step overlap their set-up with
// the previous kernel's tail.  g_pdl = 0 launches plainly (A/B switch: t2s_set_pdl).
int g_pdl = 1;
template <typename... KArgs, typename... Args>
cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

struct Workspace {
    float* h; __half* qkv; __half* o; float* mod; int* sched;
};
// run-time view of DitShape<H>
struct Shape { int H, ntok, tiles_per_pair, head_halves, lat; };
int get_shape(int latent_h, Shape* s) {
    const int H = latent_h == 0 ? 30 : latent_h;
    switch (H) {
        case 30: *s = Shape{30, DitShape<30>::NTOK, DitShape<30>::TILES_PER_PAIR, DitShape<30>::HEAD_HALVES, DitShape<30>::LAT}; return T2S_OK;
        case 50: *s = Shape{50, DitShape<50>::NTOK, DitShape<50>::TILES_PER_PAIR, DitShape<50>::HEAD_HALVES, DitShape<50>::LAT}; return T2S_OK;
        case 64: *s = Shape{64, DitShape<64>::NTOK, DitShape<64>::TILES_PER_PAIR, DitShape<64>::HEAD_HALVES, DitShape<64>::LAT}; return T2S_OK;
    }
    return fail(T2S_EINVAL, "latent width must be 30 (T2S), 50 or 64 (fork configs)%s%s");
}
#define T2S_DISPATCH_H(H, STMT)                                     \
    switch (H) {                                                    \
        case 30: { constexpr int HH = 30; STMT; } break;            \
        case 50: { constexpr int HH = 50; STMT; } break;            \
        case 64: { constexpr int HH = 64; STMT; } break;            \
        default: return fail(T2S_EINVAL, "unsupported latent width%s%s"); \
    }

void ws_offsets(int nseq, const Shape& sh, size_t off[5], size_t* total) {
    size_t p = 0;
    const size_t ntile = (size_t)((nseq + 1) / 2) * sh.tiles_per_pair;         // pair tiles of 128 rows
    off[0] = p; p = align256(p + ntile * TILE_ROWS * D * 4);                   // residual stream tiles, fp32
    off[1] = p; p = align256(p + (size_t)nseq * NHEAD * sh.head_halves * 2);   // q|k|v operand images, fp16
    off[2] = p; p = align256(p + ntile * TILE_ROWS * D * 2);                   // attention-output tiles, fp16
    off[3] = p; p = align256(p + (size_t)nseq * NLAYER * MOD * 4);
    off[4] = p; p = align256(p + fused_sched_ints((nseq + 1) / 2) * 4);          // dataflow scheduler state of the fused step kernel
    if (total) *total = p;
}
Workspace ws_view(void* base, int nseq, const Shape& sh) {
    size_t off[5];
    ws_offsets(nseq, sh, off, nullptr);
    char* b = static_cast<char*>(base);
    return Workspace{reinterpret_cast<float*>(b + off[0]), reinterpret_cast<__half*>(b + off[1]),
                     reinterpret_cast<__half*>(b + off[2]), reinterpret_cast<float*>(b + off[3]), reinterpret_cast<int*>(b + off[4])};
}

int check_ws(const void* ws, size_t bytes, int nseq, const Shape& sh) {
    if (ws == nullptr) return fail(T2S_EINVAL, "workspace is NULL%s%s");
    if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return fail(T2S_EINVAL, "workspace must be 256-byte aligned%s%s");
    if (bytes < t2s_dit_workspace_bytes_h(nseq, sh.H)) return fail(T2S_EWORKSPACE, "workspace too small%s%s");
    return T2S_OK;
}

// persistent: one CTA per SM, each looping over work items of `ne` pair tiles; small batches (all single-tile items fit
// one wave of CTAs) use single-tile items: twice the CTAs, shorter items
int token_ne(int nseq, const Shape& sh) {
    int dev = 0;
    cudaGetDevice(&dev);
    const int sms = g_sms[dev] > 0 ? g_sms[dev] : 148;
    return ((nseq + 1) / 2) * sh.tiles_per_pair <= sms ? 1 : 2;
}
int token_grid(int nseq, const Shape& sh, int ne) {
    int dev = 0;
    cudaGetDevice(&dev);
    const int items = ((nseq + 1) / 2) * (sh.tiles_per_pair / ne), sms = g_sms[dev] > 0 ? g_sms[dev] : 148;
    return items < sms ? items : sms;
}
#define T2S_TOKEN_LAUNCH(MODE_, HH_)                                                                                  \
    {                                                                                                                 \
        const int ne = token_ne(nseq, sh);                                                                            \
        if (ne == 2) CUDA_OK(launch_k(token_kernel<MODE_, HH_, 2>, token_grid(nseq, sh, 2), TC_THREADS, TOK_SMEM_BYTES, st, a)); \
        else CUDA_OK(launch_k(token_kernel<MODE_, HH_, 1>, token_grid(nseq, sh, 1), TC_THREADS, TOK_SMEM_BYTES, st, a));         \
    }

// set by sample_impl for the launches of a guided loop whose cond kernel writes one shared unconditional modulation row
thread_local bool g_uncond_shared = false;
TokArgs base_args(const t2s_dit_weights* w, const Workspace& ws, int nseq, bool uncond_shared = g_uncond_shared) {
    TokArgs a;
    memset(&a, 0, sizeof(a));
    memcpy(&a.w, w, sizeof(DitWeights));
    a.h = ws.h; a.qkv = ws.qkv; a.o = ws.o; a.mod = ws.mod; a.nseq = nseq;
    a.uncond_shared = uncond_shared ? 1 : 0;
    a.trace = g_trace;
    return a;
}

// uncond_shared (guided loops at throughput batch sizes): ONE unconditional modulation row per step (sequence 0) instead of one
// per sample; the token kernels of the same loop read it for every pair (TokArgs / FusedArgs uncond_shared)
#ifndef T2S_COND_SPLIT_MAX
// up to this many sequences the conditioning runs as cond_split_kernel (three CTAs per block and 8 sequences): the one-CTA form
// has too few CTAs to fill the GPU there.  Same-box A/B of 64 vs 256 (profiles/r02_ab_cond_split.log): batch 64 0.397 -> 0.386 ms per
// guided step, batch 128 0.677 -> 0.667; at batch 256 (512 sequences) the one-CTA form is the faster one (1.240 vs 1.252 ms)
#define T2S_COND_SPLIT_MAX 256
#endif
bool cond_uncond_shared(int nseq, int cfg_pairs) { return cfg_pairs && nseq > 64; }
int launch_cond(const t2s_dit_weights* w, const float* t100, int t_stride, const float* emb, int emb_shift, int cfg_pairs,
                int nseq, const Workspace& ws, cudaStream_t st, bool uncond_shared = false) {
    if (nseq <= T2S_COND_SPLIT_MAX)
        CUDA_OK(launch_k(cond_split_kernel, dim3((nseq + 7) / 8, NLAYER * 3), 256, 0, st, ws.mod, t100, t_stride, emb, emb_shift, cfg_pairs, w->freqs,
                         w->w_ada_t, w->b_ada, nseq));
    else if (uncond_shared)
        CUDA_OK(launch_k(cond_kernel, dim3((nseq / 2 + 7) / 8 + 1, NLAYER), 256, 0, st, ws.mod, t100, t_stride, emb, emb_shift, cfg_pairs, w->freqs,
                         w->w_ada_t, w->b_ada, nseq, 1));
    else
        CUDA_OK(launch_k(cond_kernel, dim3((nseq + 7) / 8, NLAYER), 256, 0, st, ws.mod, t100, t_stride, emb, emb_shift, cfg_pairs, w->freqs,
                         w->w_ada_t, w->b_ada, nseq, 0));
    CUDA_OK(cudaGetLastError());
    return T2S_OK;
}
// fused = inside t2s_dit_forward / t2s_sample: the residual stream entering block 0 is not materialised (EMBED skips the
// store, the MID kernel of block 0 recomputes it from x); the stage-wise test entries keep it in the workspace
int launch_embed(const t2s_dit_weights* w, const float* x, int x_shift, int nseq, const Shape& sh, const Workspace& ws, cudaStream_t st,
                 bool fused = false) {
    TokArgs a = base_args(w, ws, nseq);
    a.x = x; a.x_shift = x_shift; a.skip_h_store = fused ? 1 : 0;
    T2S_DISPATCH_H(sh.H, T2S_TOKEN_LAUNCH(TOK_EMBED, HH));
    CUDA_OK(cudaGetLastError());
    return T2S_OK;
}
int launch_attn(int nseq, const Shape& sh, const Workspace& ws, cudaStream_t st) {
    int dev = 0;
    cudaGetDevice(&dev);
    const int sms = g_sms[dev] > 0 ? g_sms[dev] : 148;
    // fewer (sequence, head) CTAs than resident slots: one CTA per q-tile (group) instead, for latency
#define T2S_ATTN_LAUNCH(HH_)                                                                                                     \
    {                                                                                                                            \
        using LS = AttShape<HH_, true>;                                                                                          \
        if (nseq * NHEAD < sms * LS::CTAS_PER_SM / 2) {                                                                          \
            const int npart = LS::NQT / LS::NWG;                                                                                 \
            CUDA_OK(launch_k(attn_kernel<HH_, true>, nseq * NHEAD * npart, LS::THREADS, LS::SMEM_BYTES, st, ws.qkv, ws.o, g_trace, npart));    \
        } else {                                                                                                                 \
            CUDA_OK(launch_k(attn_kernel<HH_, false>, nseq * NHEAD, AttShape<HH_>::THREADS, AttShape<HH_>::SMEM_BYTES, st, ws.qkv, ws.o, g_trace, 1)); \
        }                                                                                                                        \
    }
    T2S_DISPATCH_H(sh.H, T2S_ATTN_LAUNCH(HH));
#undef T2S_ATTN_LAUNCH
    CUDA_OK(cudaGetLastError());
    return T2S_OK;
}
int launch_mid(const t2s_dit_weights* w, int layer, int nseq, const Shape& sh, const Workspace& ws, cudaStream_t st,
               const float* x = nullptr, int x_shift = 0) {
    TokArgs a = base_args(w, ws, nseq);
    a.layer = layer;
    if (layer == 0 && x != nullptr) { a.recompute_h0 = 1; a.x = x; a.x_shift = x_shift; }
    T2S_DISPATCH_H(sh.H, T2S_TOKEN_LAUNCH(TOK_MID, HH));
    CUDA_OK(cudaGetLastError());
    return T2S_OK;
}
int launch_final(const t2s_dit_weights* w, int nseq, const Shape& sh, const Workspace& ws, int out_mode, float* out, float* x_upd,
                 const float* noise, float cfg, float c1, float c2, float c3, cudaStream_t st, unsigned long long seed = 0, unsigned int step = 0) {
    TokArgs a = base_args(w, ws, nseq);
    a.layer = NLAYER - 1;
    a.out_mode = out_mode; a.out = out; a.x_upd = x_upd; a.noise = noise; a.seed = seed; a.step = step;
    a.cfg = cfg; a.c1 = c1; a.c2 = c2; a.c3 = c3;
    T2S_DISPATCH_H(sh.H, T2S_TOKEN_LAUNCH(TOK_FINAL, HH));
    CUDA_OK(cudaGetLastError());
    return T2S_OK;
}

// ---- the fused per-step kernel (dit_fused.cuh): one cooperative persistent launch for everything after the conditioning
bool use_fused(const t2s_dit_weights* w, int nseq, const Shape& sh) {
    return sh.H == 30 && g_fused_min_pairs >= 0 && (nseq + 1) / 2 >= g_fused_min_pairs && w->w_qkv_half[0] != nullptr &&
           w->w_post_half[0] != nullptr;
}
int launch_fused(const t2s_dit_weights* w, const float* x, int x_shift, int nseq, const Workspace& ws, int out_mode, float* out,
                 float* x_upd, const float* noise, float cfg, float c1, float c2, float c3, cudaStream_t st, unsigned long long seed = 0,
                 unsigned int step = 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    FusedArgs a;
    memset(&a, 0, sizeof(a));
    memcpy(&a.w, w, sizeof(DitWeights));
    a.x = x; a.x_shift = x_shift; a.h = ws.h; a.qkv = ws.qkv; a.o = ws.o; a.mod = ws.mod; a.nseq = nseq;
    a.out_mode = out_mode; a.out = out; a.x_upd = x_upd; a.noise = noise; a.seed = seed; a.step = step;
    a.cfg = cfg; a.c1 = c1; a.c2 = c2; a.c3 = c3;
    a.sched = ws.sched; a.inflight = g_fused_inflight; a.stats = g_fused_stats; a.trace = g_fused_trace;
    a.uncond_shared = g_uncond_shared ? 1 : 0;
    CUDA_OK(cudaMemsetAsync(ws.sched, 0, fused_sched_ints((nseq + 1) / 2) * 4, st));
    // the CTAs wait on one another through the scheduler flags: a cooperative launch guarantees that all of them are resident
    cudaLaunchConfig_t cfgl;
    memset(&cfgl, 0, sizeof(cfgl));
    cfgl.gridDim = dim3(g_sms[dev] > 0 ? g_sms[dev] : 148);
    cfgl.blockDim = dim3(FS_THREADS);
    cfgl.dynamicSmemBytes = FS_SMEM_BYTES;
    cfgl.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfgl.attrs = attr;
    cfgl.numAttrs = 1;
    CUDA_OK(cudaLaunchKernelEx(&cfgl, fused_step_kernel, a));
    return T2S_OK;
}

}  // namespace

extern "C" {

int t2s_version(void) { return 200; }
void t2s_set_fused(int min_pairs, int inflight) { g_fused_min_pairs = min_pairs; g_fused_inflight = inflight; }
void t2s_set_pdl(int on) { g_pdl = on ? 1 : 0; }
void t2s_debug_set_fused_stats(long long* device_buf) { g_fused_stats = device_buf; }
void t2s_debug_set_fused_trace(long long* device_buf) { g_fused_trace = device_buf; }
const char* t2s_last_error(void) { return g_err; }
int t2s_init(void) { return ensure_init(); }
void t2s_debug_set_phase_trace(long long* device_buf) { g_trace = device_buf; }

size_t t2s_dit_workspace_bytes_h(int nseq, int latent_h) {
    Shape sh;
    if (get_shape(latent_h, &sh) != T2S_OK) return 0;
    size_t off[5], total = 0;
    ws_offsets(nseq > 0 ? nseq : 0, sh, off, &total);
    return total;
}
size_t t2s_dit_workspace_bytes(int nseq) { return t2s_dit_workspace_bytes_h(nseq, 30); }
int t2s_dit_workspace_offsets_h(int nseq, int latent_h, size_t offsets[4]) {
    Shape sh;
    TRY(get_shape(latent_h, &sh));
    size_t off[5];
    ws_offsets(nseq, sh, off, nullptr);
    for (int i = 0; i < 4; ++i) offsets[i] = off[i];
    return T2S_OK;
}
void t2s_dit_workspace_offsets(int nseq, size_t offsets[4]) { t2s_dit_workspace_offsets_h(nseq, 30, offsets); }

int t2s_dit_cond(const t2s_dit_weights* w, const float* t100, int t_stride, const float* emb, int cfg_pairs, int nseq,
                 void* workspace, t2s_stream_t stream) {
    if (!w || !t100 || nseq <= 0 || !workspace) return fail(T2S_EINVAL, "t2s_dit_cond: bad argument%s%s");
    Shape sh;
    TRY(get_shape(w->latent_h, &sh));
    TRY(ensure_init());
    return launch_cond(w, t100, t_stride, emb, cfg_pairs ? 1 : 0, cfg_pairs, nseq, ws_view(workspace, nseq, sh), (cudaStream_t)stream);
}
int t2s_dit_embed_qkv(const t2s_dit_weights* w, const float* x, int x_shared, int nseq, void* workspace, t2s_stream_t stream) {
    if (!w || !x || nseq <= 0 || !workspace) return fail(T2S_EINVAL, "t2s_dit_embed_qkv: bad argument%s%s");
    Shape sh;
    TRY(get_shape(w->latent_h, &sh));
    TRY(ensure_init());
    return launch_embed(w, x, x_shared ? 1 : 0, nseq, sh, ws_view(workspace, nseq, sh), (cudaStream_t)stream);
}
int t2s_dit_attention_h(int nseq, int latent_h, void* workspace, t2s_stream_t stream) {
    if (nseq <= 0 || !workspace) return fail(T2S_EINVAL, "t2s_dit_attention: bad argument%s%s");
    Shape sh;
    TRY(get_shape(latent_h, &sh));
    TRY(ensure_init());
    return launch_attn(nseq, sh, ws_view(workspace, nseq, sh), (cudaStream_t)stream);
}
int t2s_dit_attention(int nseq, void* workspace, t2s_stream_t stream) { return t2s_dit_attention_h(nseq, 30, workspace, stream); }
int t2s_dit_block_post(const t2s_dit_weights* w, int layer, int nseq, void* workspace, t2s_stream_t stream) {
    if (!w || layer < 0 || layer >= NLAYER - 1 || nseq <= 0 || !workspace)
        return fail(T2S_EINVAL, "t2s_dit_block_post: bad argument (layer must be 0..2)%s%s");
    Shape sh;
    TRY(get_shape(w->latent_h, &sh));
    TRY(ensure_init());
    return launch_mid(w, layer, nseq, sh, ws_view(workspace, nseq, sh), (cudaStream_t)stream);
}
int t2s_dit_final(const t2s_dit_weights* w, float* out, int nseq, void* workspace, t2s_stream_t stream) {
    if (!w || !out || nseq <= 0 || !workspace) return fail(T2S_EINVAL, "t2s_dit_final: bad argument%s%s");
    Shape sh;
    TRY(get_shape(w->latent_h, &sh));
    TRY(ensure_init());
    return launch_final(w, nseq, sh, ws_view(workspace, nseq, sh), OUT_FWD, out, nullptr, nullptr, 0.f, 0.f, 0.f, 0.f, (cudaStream_t)stream);
}

int t2s_dit_forward(const t2s_dit_weights* w, const float* x, const float* t100, const float* emb, float* out, int nseq,
                    void* workspace, size_t workspace_bytes, t2s_stream_t stream) {
    if (!w || !x || !t100 || !out || nseq <= 0) return fail(T2S_EINVAL, "t2s_dit_forward: bad argument%s%s");
    Shape sh;
    TRY(get_shape(w->latent_h, &sh));
    TRY(check_ws(workspace, workspace_bytes, nseq, sh));
    TRY(ensure_init());
    cudaStream_t st = (cudaStream_t)stream;
    const Workspace ws = ws_view(workspace, nseq, sh);
    TRY(launch_cond(w, t100, 1, emb, 0, 0, nseq, ws, st));
    if (use_fused(w, nseq, sh)) return launch_fused(w, x, 0, nseq, ws, OUT_FWD, out, nullptr, nullptr, 0.f, 0.f, 0.f, 0.f, st);
    TRY(launch_embed(w, x, 0, nseq, sh, ws, st, true));
    for (int l = 0; l < NLAYER; ++l) {
        TRY(launch_attn(nseq, sh, ws, st));
        if (l < NLAYER - 1) TRY(launch_mid(w, l, nseq, sh, ws, st, x, 0));
    }
    return launch_final(w, nseq, sh, ws, OUT_FWD, out, nullptr, nullptr, 0.f, 0.f, 0.f, 0.f, st);
}

namespace {
int sample_impl(const t2s_dit_weights* w, int kind, float* x, const float* emb, const float* t100, const float* coef,
                const float* step_noise, unsigned long long seed, unsigned int step0, float* pred_trace, int batch, int steps, float cfg_scale,
                void* workspace, size_t workspace_bytes, t2s_stream_t stream) {
    const int nseq = 2 * batch;
    Shape sh;
    TRY(get_shape(w->latent_h, &sh));
    TRY(check_ws(workspace, workspace_bytes, nseq, sh));
    TRY(ensure_init());
    cudaStream_t st = (cudaStream_t)stream;
    const Workspace ws = ws_view(workspace, nseq, sh);
    const size_t lat = (size_t)batch * sh.lat;
    const bool fused = use_fused(w, nseq, sh);
    struct Scope { Scope(bool v) { g_uncond_shared = v; } ~Scope() { g_uncond_shared = false; } } scope(cond_uncond_shared(nseq, 1));
    for (int j = 0; j < steps; ++j) {
        TRY(launch_cond(w, t100 + j, 0, emb, 1, 1, nseq, ws, st, g_uncond_shared));
        if (fused) {
            // one persistent launch per guided step: patch-embed, the four blocks (attention and token work of different
            // sequence pairs overlapped on every SM), final projection, guidance mix and the Euler / ancestral update
            TRY(launch_fused(w, x, 1, nseq, ws, kind == 0 ? OUT_RF : OUT_DDPM, pred_trace ? pred_trace + j * lat : nullptr, x,
                             (kind == 1 && step_noise) ? step_noise + j * lat : nullptr, cfg_scale, coef[3 * j], coef[3 * j + 1],
                             coef[3 * j + 2], st, seed, step0 + (unsigned int)j));
            continue;
        }
        TRY(launch_embed(w, x, 1, nseq, sh, ws, st, true));
        for (int l = 0; l < NLAYER; ++l) {
            TRY(launch_attn(nseq, sh, ws, st));
            if (l < NLAYER - 1) TRY(launch_mid(w, l, nseq, sh, ws, st, x, 1));
        }
        TRY(launch_final(w, nseq, sh, ws, kind == 0 ? OUT_RF : OUT_DDPM, pred_trace ? pred_trace + j * lat : nullptr, x,
                         (kind == 1 && step_noise) ? step_noise + j * lat : nullptr, cfg_scale, coef[3 * j], coef[3 * j + 1], coef[3 * j + 2], st, seed,
                         step0 + (unsigned int)j));
    }
    return T2S_OK;
}
}  // namespace

int t2s_sample(const t2s_dit_weights* w, int kind, float* x, const float* emb, const float* t100, const float* coef,
               const float* step_noise, float* pred_trace, int batch, int steps, float cfg_scale, void* workspace,
               size_t workspace_bytes, t2s_stream_t stream) {
    if (!w || !x || !emb || !t100 || !coef || batch <= 0 || steps <= 0 || (kind != 0 && kind != 1))
        return fail(T2S_EINVAL, "t2s_sample: bad argument%s%s");
    if (kind == 1 && !step_noise) return fail(T2S_EINVAL, "t2s_sample: DDPM needs step_noise (or t2s_sample_ddpm_seeded)%s%s");
    return sample_impl(w, kind, x, emb, t100, coef, step_noise, 0ull, 0u, pred_trace, batch, steps, cfg_scale, workspace, workspace_bytes, stream);
}

int t2s_sample_ddpm_seeded(const t2s_dit_weights* w, float* x, const float* emb, const float* t100, const float* coef, unsigned long long seed,
                           float* pred_trace, int batch, int steps, float cfg_scale, void* workspace, size_t workspace_bytes,
                           t2s_stream_t stream) {
    if (!w || !x || !emb || !t100 || !coef || batch <= 0 || steps <= 0) return fail(T2S_EINVAL, "t2s_sample_ddpm_seeded: bad argument%s%s");
    return sample_impl(w, 1, x, emb, t100, coef, nullptr, seed, 0u, pred_trace, batch, steps, cfg_scale, workspace, workspace_bytes, stream);
}

int t2s_vae_decode(const t2s_vae_dec_weights* w, const float* z, float* series, float* after, int batch, int length,
                   t2s_stream_t stream) {
    if (!w || !z || !series || batch <= 0) return fail(T2S_EINVAL, "t2s_vae_decode: bad argument%s%s");
    TRY(ensure_init());
    VaeDecWeights dw;
    memcpy(&dw, w, sizeof(dw));
    cudaStream_t st = (cudaStream_t)stream;
    const int smem = VAE_DEC_SMEM_FLOATS * 4;
    const bool wide = batch <= 148;                    // fewer series than SMs: 1024 threads per series (latency), else 256
#define T2S_VAE_LAUNCH(KERNEL, L4_, ...)                                                       \
    if (wide) KERNEL<L4_, 1024><<<batch, 1024, smem, st>>>(__VA_ARGS__);                       \
    else KERNEL<L4_, 256><<<batch, 256, smem, st>>>(__VA_ARGS__)
    switch (length) {
        case 24: T2S_VAE_LAUNCH(vae_decode_kernel, 6, dw, z, series, after); break;
        case 48: T2S_VAE_LAUNCH(vae_decode_kernel, 12, dw, z, series, after); break;
        case 96: T2S_VAE_LAUNCH(vae_decode_kernel, 24, dw, z, series, after); break;
        default: return fail(T2S_EINVAL, "t2s_vae_decode: length must be 24, 48 or 96%s%s");
    }
    CUDA_OK(cudaGetLastError());
    return T2S_OK;
}

int t2s_vae_encode(const t2s_vae_enc_weights* w, const float* x, float* z, float* before, int batch, int length,
                   t2s_stream_t stream) {
    if (!w || !x || !z || batch <= 0) return fail(T2S_EINVAL, "t2s_vae_encode: bad argument%s%s");
    TRY(ensure_init());
    VaeEncWeights ew;
    memcpy(&ew, w, sizeof(ew));
    cudaStream_t st = (cudaStream_t)stream;
    const int smem = VAE_ENC_SMEM_FLOATS * 4;
    const bool wide = batch <= 148;
    switch (length) {
        case 24: T2S_VAE_LAUNCH(vae_encode_kernel, 6, ew, x, z, before); break;
        case 48: T2S_VAE_LAUNCH(vae_encode_kernel, 12, ew, x, z, before); break;
        case 96: T2S_VAE_LAUNCH(vae_encode_kernel, 24, ew, x, z, before); break;
        default: return fail(T2S_EINVAL, "t2s_vae_encode: length must be 24, 48 or 96%s%s");
    }
    CUDA_OK(cudaGetLastError());
    return T2S_OK;
}

int t2s_rf_euler(const float* x, const float* v, float dt, float* out, size_t n, t2s_stream_t stream) {
    if (!x || !v || !out || n == 0) return fail(T2S_EINVAL, "t2s_rf_euler: bad argument%s%s");
    TRY(ensure_init());
    const size_t blocks = (n + 255) / 256;
    rf_euler_kernel<<<(unsigned)(blocks > 4736 ? 4736 : blocks), 256, 0, (cudaStream_t)stream>>>(x, v, dt, out, n);
    CUDA_OK(cudaGetLastError());
    return T2S_OK;
}

int t2s_ddpm_p_sample(const float* xt, const float* eps_hat, const float* noise, const float* c1, const float* c2, const float* c3,
                      float* out, int batch, int elems_per_sample, t2s_stream_t stream) {
    if (!xt || !eps_hat || !noise || !c1 || !c2 || !c3 || !out || batch <= 0 || elems_per_sample <= 0)
        return fail(T2S_EINVAL, "t2s_ddpm_p_sample: bad argument%s%s");
    TRY(ensure_init());
    const size_t n = (size_t)batch * elems_per_sample, blocks = (n + 255) / 256;
    ddpm_p_sample_kernel<<<(unsigned)(blocks > 4736 ? 4736 : blocks), 256, 0, (cudaStream_t)stream>>>(xt, eps_hat, noise, c1, c2, c3, out, n,
                                                                                                   elems_per_sample);
    CUDA_OK(cudaGetLastError());
    return T2S_OK;
}

int t2s_series_metrics(const float* ori, const float* gen, int n, int length, float* per_sample, double* out, t2s_stream_t stream) {
    if (!ori || !gen || !per_sample || !out || n <= 0 || length <= 0) return fail(T2S_EINVAL, "t2s_series_metrics: bad argument%s%s");
    TRY(ensure_init());
    cudaStream_t st = (cudaStream_t)stream;
    series_sums_kernel<<<(n + 7) / 8, 256, 0, st>>>(ori, gen, n, length, per_sample);
    series_metrics_finish_kernel<<<1, 1024, 0, st>>>(per_sample, n, length, out);
    CUDA_OK(cudaGetLastError());
    return T2S_OK;
}

}  // extern "C"
