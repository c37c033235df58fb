// C ABI of the training step (include/t2s_b200.h: t2s_dit_train_step and friends): host-side orchestration of the
// modular forward / backward in train_kernels.cuh.  Enqueue-only on the caller's stream; caller-owned workspace.
#include <cmath>
#include <cstring>

#include "api_common.h"
#include "train_kernels.cuh"
#include "train_attn.cuh"

using namespace t2s;
using namespace t2s_api;

namespace {

constexpr int MAX_DEV = 64;
bool g_train_inited[MAX_DEV] = {};
int g_sms[MAX_DEV] = {};

int train_init() {
    TRY(ensure_init());
    int dev = 0;
    CUDA_OK(cudaGetDevice(&dev));
    if (g_train_inited[dev]) return T2S_OK;
    CUDA_OK(cudaFuncSetAttribute(gemm_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(gemm_tf32_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES));
#define T2S_TA_ATTRS(HH)                                                                                                              \
    CUDA_OK(cudaFuncSetAttribute(ta_attn_kernel<TA_FWD, HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, TaShape<HH>::SMEM_BYTES));  \
    CUDA_OK(cudaFuncSetAttribute(ta_attn_kernel<TA_DQ, HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, TaShape<HH>::SMEM_BYTES));   \
    CUDA_OK(cudaFuncSetAttribute(ta_attn_kernel<TA_DKV, HH>, cudaFuncAttributeMaxDynamicSharedMemorySize, TaShape<HH>::SMEM_BYTES));
    T2S_TA_ATTRS(30)
    T2S_TA_ATTRS(50)
    T2S_TA_ATTRS(64)
#undef T2S_TA_ATTRS
    CUDA_OK(cudaDeviceGetAttribute(&g_sms[dev], cudaDevAttrMultiProcessorCount, dev));
    g_train_inited[dev] = true;
    return T2S_OK;
}
int sm_count() {
    int dev = 0;
    cudaGetDevice(&dev);
    return g_sms[dev] > 0 ? g_sms[dev] : 148;
}

// latent width -> token geometry of the training path (0 = the T2S default 30)
struct Geom { int H, ntok, rc, lat; };
int get_geom(int latent_h, Geom* g) {
    const int H = latent_h == 0 ? 30 : latent_h;
    if (H != 30 && H != 50 && H != 64) return fail(T2S_EINVAL, "latent width must be 30 (T2S), 50 or 64 (fork configs)%s%s");
    *g = Geom{H, 16 * H, H == 30 ? 60 : (H == 50 ? 50 : 64), LATC * H};
    return T2S_OK;
}
#define T2S_TRAIN_DISPATCH_H(H, STMT)                               \
    switch (H) {                                                    \
        case 30: { constexpr int HH = 30; STMT; } break;            \
        case 50: { constexpr int HH = 50; STMT; } break;            \
        case 64: { constexpr int HH = 64; STMT; } break;            \
        default: return fail(T2S_EINVAL, "unsupported latent width%s%s"); \
    }

struct Gemm {
    GemmArgs a;
    int batch = 1;
    Gemm(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc) {
        memset(&a, 0, sizeof(a));
        a.A = A; a.B = B; a.C = C; a.M = M; a.N = N; a.K = K; a.lda = lda; a.ldb = ldb; a.ldc = ldc;
        a.bdiv = 1; a.ksplit = 1; a.mode = GEMM_STORE; a.alpha = 1.f;
        a.bn = N <= 32 ? 32 : (N <= 64 ? 64 : (N % 128 != 0 && N % 96 == 0 ? 96 : 128));
    }
    Gemm& bias(const float* b) { a.bias = b; return *this; }
    Gemm& amn() { a.a_mn = 1; return *this; }
    Gemm& bmn() { a.b_mn = 1; return *this; }
    Gemm& mode(int m) { a.mode = m; return *this; }
    Gemm& qkv_images(__half* img, int ntok) { a.qkv_img = img; a.img_ntok = ntok; return *this; }
    Gemm& alpha(float v) { a.alpha = v; return *this; }
    Gemm& ksplit(int k) { a.ksplit = k < 1 ? 1 : k; return *this; }
    // weight-gradient form: few output tiles, long K -> split K so that about two waves of CTAs run
    Gemm& wgrad() {
        const int tiles = ((a.M + G_BM - 1) / G_BM) * ((a.N + a.bn - 1) / a.bn), ksteps = (a.K + G_BK - 1) / G_BK;
        int ks = (2 * sm_count() + tiles - 1) / tiles;
        if (ks > ksteps) ks = ksteps;
        a.ksplit = ks < 1 ? 1 : ks;
        a.mode = GEMM_ATOMIC;
        return *this;
    }
    int launch(cudaStream_t st) const {
        if ((a.lda & 3) || (a.ldb & 3) || (!a.a_mn && (a.K & 3)) || (!a.b_mn && (a.K & 3)) || (a.a_mn && (a.M & 3)) || (a.b_mn && (a.N & 3)))
            return fail(T2S_EINVAL, "gemm_tf32: leading dimensions / extents must be multiples of 4%s%s");
        if (((uintptr_t)a.A | (uintptr_t)a.B) & 15) return fail(T2S_EINVAL, "gemm_tf32: operands must be 16-byte aligned%s%s");
        if ((a.mode & ~15) != 0 && (a.a_mn || batch != 1 || a.ksplit != 1 || a.N % 128 != 0 || a.M < 4 * G_BM))
            return fail(T2S_EINVAL, "gemm_tf32: profiling mode bits apply to the persistent form only%s%s");
        if (a.qkv_img != nullptr && (a.a_mn || batch != 1 || a.ksplit != 1 || a.mode != GEMM_STORE || a.N != 3 * D || a.M < 4 * G_BM || a.img_ntok <= 0 || a.M % a.img_ntok != 0))
            return fail(T2S_EINVAL, "gemm_tf32: the q|k|v image epilogue needs the persistent form with N = 384%s%s");
        if (a.mode == GEMM_ATOMIC && a.bias != nullptr) return fail(T2S_EINVAL, "gemm_tf32: bias with atomic accumulation%s%s");
        if (a.ksplit > 1 && a.mode != GEMM_ATOMIC) return fail(T2S_EINVAL, "gemm_tf32: split-K needs atomic accumulation%s%s");
        // forward / input-gradient shapes: the persistent form (decoupled load / MMA / epilogue roles, one CTA per SM)
        const bool vec_c = (a.qkv_img != nullptr || ((a.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(a.C) & 15) == 0)) && (a.bias == nullptr || (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0);
        if (!a.a_mn && batch == 1 && a.ksplit == 1 && (a.mode & 15) != GEMM_ATOMIC && a.N % 128 == 0 && a.M >= 4 * G_BM && vec_c) {
            const int items = ((a.M + G_BM - 1) / G_BM) * (a.N / 128), sms = sm_count();
            gemm_tf32_persistent_kernel<<<items < sms ? items : sms, P_THREADS, P_SMEM_BYTES, st>>>(a);
            CUDA_OK(cudaGetLastError());
            return T2S_OK;
        }
        dim3 grid((a.M + G_BM - 1) / G_BM, (a.N + a.bn - 1) / a.bn, batch * a.ksplit);
        gemm_tf32_kernel<<<grid, G_THREADS, G_SMEM_BYTES, st>>>(a);
        CUDA_OK(cudaGetLastError());
        return T2S_OK;
    }
};

size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

// scratch of the fused attention (train_attn.cuh): fp16 operand images of q | k | v and of the scaled dO, the per-row
// D = rowsum(dO . O) and the per-(sequence, head) inverse dO scale
struct AttnScratch {
    __half *img, *doimg;
    float *dvec, *dinv;
    size_t total;
};
AttnScratch attn_scratch(void* base, int nseq, int ntok) {
    const size_t TA_IMG_BYTES = (size_t)ntok * HD * 2;
    const int NTOK = ntok;
    AttnScratch a;
    char* b = static_cast<char*>(base);
    size_t p = 0;
    auto take = [&](size_t bytes) { char* r = b + p; p = align256(p + bytes); return r; };
    a.img = reinterpret_cast<__half*>(take((size_t)nseq * NHEAD * 3 * TA_IMG_BYTES + 1024));
    a.doimg = reinterpret_cast<__half*>(take((size_t)nseq * NHEAD * TA_IMG_BYTES + 1024));
    a.dvec = reinterpret_cast<float*>(take((size_t)nseq * NHEAD * NTOK * 4));
    a.dinv = reinterpret_cast<float*>(take((size_t)nseq * NHEAD * 4));
    a.total = p;
    return a;
}

struct TrainWs {
    float *sc, *mod, *dmod, *xp, *wemb, *bemb, *red;
    float* h[NLAYER + 1];
    __half* img[NLAYER];                                // q | k | v operand images of every block (written by the QKV GEMM epilogue)
    float *a1[NLAYER], *o[NLAYER], *y1[NLAYER], *hm[NLAYER], *a2[NLAYER], *z1[NLAYER], *hid[NLAYER], *y2[NLAYER];
    float* nlse[NLAYER];
    float *g, *g2, *d1, *d2, *dqkv, *dob;
    AttnScratch att;
    size_t total;
};
TrainWs train_ws(void* base, int nseq, int ntok) {
    const int NTOK = ntok;
    const size_t TA_IMG_HALVES = (size_t)ntok * HD;
    TrainWs w;
    char* b = static_cast<char*>(base);
    size_t p = 0;
    const size_t T = (size_t)nseq * NTOK;
    auto take = [&](size_t floats) { float* r = reinterpret_cast<float*>(b + p); p = align256(p + floats * 4); return r; };
    w.sc = take((size_t)nseq * D); w.mod = take((size_t)nseq * NLAYER * MOD); w.dmod = take((size_t)nseq * NLAYER * MOD);
    w.xp = take(T * 4); w.wemb = take(D * 4); w.bemb = take(D); w.red = take(D * 5);
    for (int l = 0; l <= NLAYER; ++l) w.h[l] = take(T * D);
    for (int l = 0; l < NLAYER; ++l) {
        w.a1[l] = take(T * D); w.img[l] = reinterpret_cast<__half*>(take((size_t)nseq * NHEAD * 3 * TA_IMG_HALVES / 2 + 256)); w.o[l] = take(T * D); w.y1[l] = take(T * D); w.hm[l] = take(T * D);
        w.a2[l] = take(T * D); w.z1[l] = take(T * DMLP); w.hid[l] = take(T * DMLP); w.y2[l] = take(T * D);
    }
    w.g = take(T * D); w.g2 = take(T * D); w.d1 = take(T * D); w.d2 = take(T * DMLP); w.dqkv = take(T * 3 * D); w.dob = take(T * D);
    for (int l = 0; l < NLAYER; ++l) w.nlse[l] = take((size_t)nseq * NHEAD * NTOK);
    w.att = attn_scratch(b + p, nseq, ntok);
    p = align256(p + w.att.total);
    w.total = p;
    return w;
}

// q | k | v rows [T][384] fp32 -> fp16 operand images (only the exported test entries and single-sequence batches need
// it: the training step's QKV GEMM writes the images in its epilogue)
int pack_qkv(const float* qkv, __half* img, int nseq, const Geom& g, cudaStream_t st) {
    const long long warps = (long long)nseq * (g.ntok / 8) * 12;
    ta_pack_qkv_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(qkv, img, nseq, g.ntok);
    CUDA_OK(cudaGetLastError());
    return T2S_OK;
}
// softmax(q k^T / sqrt(32)) v of every (sequence, head), keeping 4 - log2-sum-exp per row for the backward
// (timm Attention -> F.scaled_dot_product_attention, transformer.py:116)
int attn_forward_images(const __half* img, float* o, float* nlse, int nseq, const Geom& g, cudaStream_t st) {
    TaArgs p{};
    p.img = img; p.o = o; p.nlse = nlse;
    T2S_TRAIN_DISPATCH_H(g.H, (ta_attn_kernel<TA_FWD, HH><<<nseq * NHEAD * TaShape<HH>::NTILE, TA_THREADS, TaShape<HH>::SMEM_BYTES, st>>>(p)));
    CUDA_OK(cudaGetLastError());
    return T2S_OK;
}
// d(q | k | v) from dO, recomputing the probabilities from the q, k images and the saved log-sum-exp
int attn_backward_images(const __half* img, const float* o, const float* nlse, const float* dout, float* dqkv, const AttnScratch& a, int nseq,
                         const Geom& g, cudaStream_t st) {
    switch (g.H) {
        case 30: ta_pack_do_kernel<4><<<nseq * NHEAD, 512, 0, st>>>(dout, o, a.doimg, a.dvec, a.dinv, g.ntok); break;
        case 50: ta_pack_do_kernel<7><<<nseq * NHEAD, 512, 0, st>>>(dout, o, a.doimg, a.dvec, a.dinv, g.ntok); break;
        default: ta_pack_do_kernel<8><<<nseq * NHEAD, 512, 0, st>>>(dout, o, a.doimg, a.dvec, a.dinv, g.ntok); break;
    }
    TaArgs p{};
    p.img = img; p.doimg = a.doimg; p.nlse = const_cast<float*>(nlse); p.dvec = a.dvec; p.dinv = a.dinv; p.dqkv = dqkv;
    T2S_TRAIN_DISPATCH_H(g.H, (ta_attn_kernel<TA_DQ, HH><<<nseq * NHEAD * TaShape<HH>::NTILE, TA_THREADS, TaShape<HH>::SMEM_BYTES, st>>>(p)));
    T2S_TRAIN_DISPATCH_H(g.H, (ta_attn_kernel<TA_DKV, HH><<<nseq * NHEAD * TaShape<HH>::NTILE, TA_THREADS, TaShape<HH>::SMEM_BYTES, st>>>(p)));
    CUDA_OK(cudaGetLastError());
    return T2S_OK;
}
int attn_forward(const float* qkv, float* o, float* nlse, const AttnScratch& a, int nseq, const Geom& g, cudaStream_t st) {
    TRY(pack_qkv(qkv, a.img, nseq, g, st));
    return attn_forward_images(a.img, o, nlse, nseq, g, st);
}
int attn_backward(const float* qkv, const float* o, const float* nlse, const float* dout, float* dqkv, const AttnScratch& a, int nseq,
                  const Geom& g, cudaStream_t st) {
    TRY(pack_qkv(qkv, a.img, nseq, g, st));
    return attn_backward_images(a.img, o, nlse, dout, dqkv, a, nseq, g, st);
}

int colsum(const float* x, size_t rows, int ld, int cols, float* out, cudaStream_t st) {
    const int gy = (int)(rows / 64 < 1 ? 1 : (rows / 64 > 592 ? 592 : rows / 64));
    colsum_kernel<<<dim3((cols + D - 1) / D, gy), 256, 0, st>>>(x, rows, ld, cols, out);
    CUDA_OK(cudaGetLastError());
    return T2S_OK;
}

}  // namespace

extern "C" {

size_t t2s_train_workspace_bytes_h(int nseq, int latent_h) {
    Geom g;
    if (get_geom(latent_h, &g) != T2S_OK) return 0;
    return train_ws(nullptr, nseq > 0 ? nseq : 0, g.ntok).total;
}
size_t t2s_train_workspace_bytes(int nseq) { return t2s_train_workspace_bytes_h(nseq, 30); }

int t2s_gemm_tf32(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int lda, int ldb, int ldc,
                  int a_mn, int b_mn, int mode, float alpha, int ksplit, t2s_stream_t stream) {
    if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0) return fail(T2S_EINVAL, "t2s_gemm_tf32: bad argument%s%s");
    TRY(train_init());
    Gemm g(A, B, C, M, N, K, lda, ldb, ldc);
    g.bias(bias).mode(mode).alpha(alpha).ksplit(ksplit);
    if (a_mn) g.amn();
    if (b_mn) g.bmn();
    return g.launch((cudaStream_t)stream);
}

size_t t2s_train_attention_scratch_bytes(int nseq) { return attn_scratch(nullptr, nseq > 0 ? nseq : 0, NTOK).total; }

int t2s_train_attention_forward(const float* qkv, float* o, float* nlse, int nseq, void* scratch, size_t scratch_bytes, t2s_stream_t stream) {
    if (!qkv || !o || !nlse || nseq <= 0 || !scratch) return fail(T2S_EINVAL, "t2s_train_attention_forward: bad argument%s%s");
    if ((reinterpret_cast<uintptr_t>(scratch) & 255) != 0) return fail(T2S_EINVAL, "scratch must be 256-byte aligned%s%s");
    if (scratch_bytes < t2s_train_attention_scratch_bytes(nseq)) return fail(T2S_EWORKSPACE, "attention scratch too small%s%s");
    TRY(train_init());
    Geom g;
    TRY(get_geom(30, &g));
    return attn_forward(qkv, o, nlse, attn_scratch(scratch, nseq, g.ntok), nseq, g, (cudaStream_t)stream);
}

int t2s_train_attention_backward(const float* qkv, const float* o, const float* nlse, const float* dout, float* dqkv, int nseq,
                                 void* scratch, size_t scratch_bytes, t2s_stream_t stream) {
    if (!qkv || !o || !nlse || !dout || !dqkv || nseq <= 0 || !scratch) return fail(T2S_EINVAL, "t2s_train_attention_backward: bad argument%s%s");
    if ((reinterpret_cast<uintptr_t>(scratch) & 255) != 0) return fail(T2S_EINVAL, "scratch must be 256-byte aligned%s%s");
    if (scratch_bytes < t2s_train_attention_scratch_bytes(nseq)) return fail(T2S_EWORKSPACE, "attention scratch too small%s%s");
    TRY(train_init());
    Geom g;
    TRY(get_geom(30, &g));
    return attn_backward(qkv, o, nlse, dout, dqkv, attn_scratch(scratch, nseq, g.ntok), nseq, g, (cudaStream_t)stream);
}

int t2s_train_make_inputs_h(int kind, const float* x1, const float* noise, const float* ca, const float* cb, float* x_t, float* target,
                            int batch, int latent_h, t2s_stream_t stream) {
    if (!x1 || !noise || !ca || !x_t || !target || batch <= 0 || (kind != 0 && kind != 1) || (kind == 1 && !cb))
        return fail(T2S_EINVAL, "t2s_train_make_inputs: bad argument%s%s");
    Geom g;
    TRY(get_geom(latent_h, &g));
    TRY(train_init());
    const size_t n = (size_t)batch * g.lat;
    make_train_inputs_kernel<<<(unsigned)((n + 255) / 256 > 4736 ? 4736 : (n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(kind, x1, noise, ca, cb, x_t, target, n,
                                                                                                                        g.lat);
    CUDA_OK(cudaGetLastError());
    return T2S_OK;
}
int t2s_train_make_inputs(int kind, const float* x1, const float* noise, const float* ca, const float* cb, float* x_t, float* target,
                          int batch, t2s_stream_t stream) {
    return t2s_train_make_inputs_h(kind, x1, noise, ca, cb, x_t, target, batch, 30, stream);
}

int t2s_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, int step, float lr, float beta1,
                   float beta2, float eps, float weight_decay, float grad_scale, t2s_stream_t stream) {
    if (!param || !grad || !exp_avg || !exp_avg_sq || n == 0 || step < 1) return fail(T2S_EINVAL, "t2s_adamw_step: bad argument%s%s");
    TRY(train_init());
    const float bc1 = (float)(1.0 - std::pow((double)beta1, step)), bc2s = (float)std::sqrt(1.0 - std::pow((double)beta2, step));
    adamw_kernel<<<(unsigned)((n + 255) / 256 > 2368 ? 2368 : (n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2s, grad_scale);
    CUDA_OK(cudaGetLastError());
    return T2S_OK;
}

}  // extern "C"

namespace {
struct Dims {
    int nseq, T, MS; dim3 rgrid; size_t n256; unsigned egrid; Geom g;
    Dims(int n, const Geom& gg) : nseq(n), T(n * gg.ntok), MS(NLAYER * MOD), rgrid(n, gg.ntok / gg.rc), n256((size_t)n * gg.ntok * DMLP / 4), g(gg) {
        egrid = (unsigned)((n256 + 255) / 256 > 4736 ? 4736 : (n256 + 255) / 256);
    }
};

int train_forward(const t2s_dit_params* P, const float* x_t, const float* t100, const float* emb, const TrainWs& w, const Dims& d, cudaStream_t st) {
    const int nseq = d.nseq, T = d.T, MS = d.MS;
    const dim3 rgrid = d.rgrid;
    const size_t n256 = d.n256;
    const unsigned egrid = d.egrid;
    // ------------------------------------------------------------------ forward (transformer.py:158-193)
    embed_fold_kernel<<<1, D, 0, st>>>(P->pe_w, P->pe_b, P->conv_w, P->conv_b, w.wemb, w.bemb);
    cond_act_kernel<<<(nseq * D + 255) / 256, 256, 0, st>>>(w.sc, t100, emb, P->freqs, nseq);
    CUDA_OK(cudaGetLastError());
    for (int l = 0; l < NLAYER; ++l)                                      // adaLN_modulation (:106-109,115)
        TRY(Gemm(w.sc, P->ada_w[l], w.mod + l * MOD, nseq, MOD, D, D, D, MS).bias(P->ada_b[l]).launch(st));
    embed_fwd_kernel<<<rgrid, ROW_THREADS, 0, st>>>(x_t, w.wemb, w.bemb, P->pos, w.h[0], w.xp, nseq, d.g.rc);
    CUDA_OK(cudaGetLastError());
    for (int l = 0; l < NLAYER; ++l) {
        const int mo = l * MOD;
        ln_mod_fwd_kernel<<<rgrid, ROW_THREADS, 0, st>>>(w.h[l], w.mod, MS, mo, w.a1[l], 1e-6f, d.g.rc);
        if (T >= 4 * G_BM) {                                               // q | k | v go straight out as the attention's operand images
            TRY(Gemm(w.a1[l], P->qkv_w[l], w.dqkv, T, 3 * D, D, D, D, 3 * D).bias(P->qkv_b[l]).qkv_images(w.img[l], d.g.ntok).launch(st));
        } else {                                                           // a single sequence: fp32 rows (dqkv is free here), then the pack kernel
            TRY(Gemm(w.a1[l], P->qkv_w[l], w.dqkv, T, 3 * D, D, D, D, 3 * D).bias(P->qkv_b[l]).launch(st));
            TRY(pack_qkv(w.dqkv, w.img[l], nseq, d.g, st));
        }
        TRY(attn_forward_images(w.img[l], w.o[l], w.nlse[l], nseq, d.g, st));
        TRY(Gemm(w.o[l], P->proj_w[l], w.y1[l], T, D, D, D, D, D).bias(P->proj_b[l]).launch(st));
        gate_res_fwd_kernel<<<rgrid, ROW_THREADS, 0, st>>>(w.h[l], w.y1[l], w.mod, MS, mo + 2 * D, w.hm[l], d.g.rc);
        ln_mod_fwd_kernel<<<rgrid, ROW_THREADS, 0, st>>>(w.hm[l], w.mod, MS, mo + 3 * D, w.a2[l], 1e-6f, d.g.rc);
        TRY(Gemm(w.a2[l], P->fc1_w[l], w.z1[l], T, DMLP, D, D, D, DMLP).bias(P->fc1_b[l]).launch(st));
        gelu_fwd_kernel<<<egrid, 256, 0, st>>>(w.z1[l], w.hid[l], n256);
        TRY(Gemm(w.hid[l], P->fc2_w[l], w.y2[l], T, D, DMLP, DMLP, DMLP, D).bias(P->fc2_b[l]).launch(st));
        gate_res_fwd_kernel<<<rgrid, ROW_THREADS, 0, st>>>(w.hm[l], w.y2[l], w.mod, MS, mo + 5 * D, w.h[l + 1], d.g.rc);
        CUDA_OK(cudaGetLastError());
    }
    return T2S_OK;
}

int train_backward(const t2s_dit_params* P, const t2s_dit_params* Gp, const TrainWs& w, const Dims& d, cudaStream_t st) {
    const int nseq = d.nseq, T = d.T, MS = d.MS;
    const dim3 rgrid = d.rgrid;
    const size_t n256 = d.n256;
    const unsigned egrid = d.egrid;
    // ------------------------------------------------------------------ backward
    CUDA_OK(cudaMemsetAsync(w.dmod, 0, (size_t)nseq * MS * 4, st));
    for (int l = NLAYER - 1; l >= 0; --l) {
        const int mo = l * MOD;
        // MLP branch: h' = hm + gate_mlp * fc2(GELU(fc1(a2)))
        gate_res_bwd_kernel<<<rgrid, ROW_THREADS, 0, st>>>(w.g, w.y2[l], w.mod, w.dmod, MS, mo + 5 * D, w.d1, Gp->fc2_b[l], d.g.rc);
        TRY(Gemm(w.d1, w.hid[l], Gp->fc2_w[l], D, DMLP, T, D, DMLP, DMLP).amn().bmn().wgrad().launch(st));
        TRY(Gemm(w.d1, P->fc2_w[l], w.d2, T, DMLP, D, D, DMLP, DMLP).bmn().launch(st));
        gelu_bwd_kernel<<<egrid, 256, 0, st>>>(w.z1[l], w.d2, n256);
        TRY(colsum(w.d2, T, DMLP, DMLP, Gp->fc1_b[l], st));
        TRY(Gemm(w.d2, w.a2[l], Gp->fc1_w[l], DMLP, D, T, DMLP, D, D).amn().bmn().wgrad().launch(st));
        TRY(Gemm(w.d2, P->fc1_w[l], w.d1, T, D, DMLP, DMLP, D, D).bmn().launch(st));
        ln_mod_bwd_kernel<<<rgrid, ROW_THREADS, 0, st>>>(w.d1, w.hm[l], w.mod, w.dmod, MS, mo + 3 * D, w.g, w.g2, 1e-6f, d.g.rc);
        // attention branch: hm = h + gate_msa * proj(attn(a1))
        gate_res_bwd_kernel<<<rgrid, ROW_THREADS, 0, st>>>(w.g2, w.y1[l], w.mod, w.dmod, MS, mo + 2 * D, w.d1, Gp->proj_b[l], d.g.rc);
        TRY(Gemm(w.d1, w.o[l], Gp->proj_w[l], D, D, T, D, D, D).amn().bmn().wgrad().launch(st));
        TRY(Gemm(w.d1, P->proj_w[l], w.dob, T, D, D, D, D, D).bmn().launch(st));
        CUDA_OK(cudaGetLastError());
        TRY(attn_backward_images(w.img[l], w.o[l], w.nlse[l], w.dob, w.dqkv, w.att, nseq, d.g, st));
        TRY(colsum(w.dqkv, T, 3 * D, 3 * D, Gp->qkv_b[l], st));
        TRY(Gemm(w.dqkv, w.a1[l], Gp->qkv_w[l], 3 * D, D, T, 3 * D, D, D).amn().bmn().wgrad().launch(st));
        TRY(Gemm(w.dqkv, P->qkv_w[l], w.d1, T, D, 3 * D, 3 * D, D, D).bmn().launch(st));
        ln_mod_bwd_kernel<<<rgrid, ROW_THREADS, 0, st>>>(w.d1, w.h[l], w.mod, w.dmod, MS, mo, w.g2, w.g, 1e-6f, d.g.rc);
        CUDA_OK(cudaGetLastError());
    }
    // patch embedding (transformer.py:166-172)
    CUDA_OK(cudaMemsetAsync(w.red, 0, D * 5 * 4, st));
    embed_bwd_kernel<<<rgrid, ROW_THREADS, 0, st>>>(w.g, w.xp, w.red, d.g.rc);
    embed_bwd_finish_kernel<<<1, D, 0, st>>>(w.red, P->pe_w, P->conv_w, P->conv_b, Gp->pe_w, Gp->pe_b, Gp->conv_w, Gp->conv_b);
    CUDA_OK(cudaGetLastError());
    // adaLN Linear: mod_l = SiLU(c) W_l^T + b_l  (c = time embedding + text has no trainable ancestors)
    for (int l = 0; l < NLAYER; ++l) {
        TRY(Gemm(w.dmod + l * MOD, w.sc, Gp->ada_w[l], MOD, D, nseq, MS, D, D).amn().bmn().wgrad().launch(st));
        TRY(colsum(w.dmod + l * MOD, nseq, MS, MOD, Gp->ada_b[l], st));
    }
    return T2S_OK;
}

int check_train_args(const void* workspace, size_t workspace_bytes, int nseq, const Geom& g) {
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return fail(T2S_EINVAL, "workspace must be 256-byte aligned%s%s");
    if (workspace_bytes < t2s_train_workspace_bytes_h(nseq, g.H)) return fail(T2S_EWORKSPACE, "training workspace too small%s%s");
    return train_init();
}
}  // namespace

extern "C" {

int t2s_dit_train_step(const t2s_dit_params* P, const t2s_dit_params* Gp, const float* x_t, const float* t100, const float* emb,
                       const float* target, float* loss_sum, float* pred, int nseq, double loss_numel, void* workspace,
                       size_t workspace_bytes, t2s_stream_t stream) {
    if (!P || !x_t || !t100 || nseq <= 0 || !workspace) return fail(T2S_EINVAL, "t2s_dit_train_step: bad argument%s%s");
    if (Gp != nullptr && (!target || !loss_sum || !(loss_numel > 0))) return fail(T2S_EINVAL, "t2s_dit_train_step: backward needs target / loss_sum / loss_numel%s%s");
    if (target != nullptr && loss_sum == nullptr) return fail(T2S_EINVAL, "t2s_dit_train_step: target without loss_sum%s%s");
    Geom geom;
    TRY(get_geom(P->latent_h, &geom));
    TRY(check_train_args(workspace, workspace_bytes, nseq, geom));
    cudaStream_t st = (cudaStream_t)stream;
    const TrainWs w = train_ws(workspace, nseq, geom.ntok);
    const Dims d(nseq, geom);
    TRY(train_forward(P, x_t, t100, emb, w, d, st));
    const bool bwd = Gp != nullptr;
    final_kernel<<<d.rgrid, ROW_THREADS, 0, st>>>(w.h[NLAYER], P->ln_w, P->ln_b, P->lf_w, P->lf_b, pred, target, nullptr,
                                                  bwd ? (float)(2.0 / loss_numel) : 0.f, loss_sum, bwd ? w.g : nullptr,
                                                  bwd ? Gp->ln_w : nullptr, bwd ? Gp->ln_b : nullptr, bwd ? Gp->lf_w : nullptr, bwd ? Gp->lf_b : nullptr, d.g.rc);
    CUDA_OK(cudaGetLastError());
    return bwd ? train_backward(P, Gp, w, d, st) : T2S_OK;
}

int t2s_dit_train_forward(const t2s_dit_params* P, const float* x_t, const float* t100, const float* emb, float* pred, int nseq,
                          void* workspace, size_t workspace_bytes, t2s_stream_t stream) {
    if (!P || !x_t || !t100 || !pred || nseq <= 0 || !workspace) return fail(T2S_EINVAL, "t2s_dit_train_forward: bad argument%s%s");
    Geom geom;
    TRY(get_geom(P->latent_h, &geom));
    TRY(check_train_args(workspace, workspace_bytes, nseq, geom));
    cudaStream_t st = (cudaStream_t)stream;
    const TrainWs w = train_ws(workspace, nseq, geom.ntok);
    const Dims d(nseq, geom);
    TRY(train_forward(P, x_t, t100, emb, w, d, st));
    final_kernel<<<d.rgrid, ROW_THREADS, 0, st>>>(w.h[NLAYER], P->ln_w, P->ln_b, P->lf_w, P->lf_b, pred, nullptr, nullptr, 0.f, nullptr,
                                                  nullptr, nullptr, nullptr, nullptr, nullptr, d.g.rc);
    CUDA_OK(cudaGetLastError());
    return T2S_OK;
}

int t2s_dit_train_backward(const t2s_dit_params* P, const t2s_dit_params* Gp, const float* dpred, int nseq, void* workspace,
                           size_t workspace_bytes, t2s_stream_t stream) {
    if (!P || !Gp || !dpred || nseq <= 0 || !workspace) return fail(T2S_EINVAL, "t2s_dit_train_backward: bad argument%s%s");
    Geom geom;
    TRY(get_geom(P->latent_h, &geom));
    TRY(check_train_args(workspace, workspace_bytes, nseq, geom));
    cudaStream_t st = (cudaStream_t)stream;
    const TrainWs w = train_ws(workspace, nseq, geom.ntok);
    const Dims d(nseq, geom);
    final_kernel<<<d.rgrid, ROW_THREADS, 0, st>>>(w.h[NLAYER], P->ln_w, P->ln_b, P->lf_w, P->lf_b, nullptr, nullptr, dpred, 0.f, nullptr,
                                                  w.g, Gp->ln_w, Gp->ln_b, Gp->lf_w, Gp->lf_b, d.g.rc);
    CUDA_OK(cudaGetLastError());
    return train_backward(P, Gp, w, d, st);
}

}  // extern "C"
