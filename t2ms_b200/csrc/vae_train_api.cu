// C ABI of the generic LA-VAE path (include/t2s_b200.h: t2s_lavae_*): layer-wise forward with saved activations and the
// exact backward of model/pretrained/vqvae.py:36-135 / model/pretrained/myvqvae.py:24-136.
// Enqueue-only on the caller's stream; caller-owned workspace.
#include <cstring>

#include "api_common.h"
#include "vae_train_kernels.cuh"

using namespace t2s;
using namespace t2s_api;

namespace {

size_t align256(size_t v) { return (v + 255) & ~size_t(255); }
constexpr int MAX_RES = 4;
constexpr size_t LAVAE_SMEM_LIMIT = 200 * 1024;
bool g_lavae_inited[64] = {};
int lavae_init() {
    TRY(ensure_init());
    int dev = 0;
    CUDA_OK(cudaGetDevice(&dev));
    if (g_lavae_inited[dev]) return T2S_OK;
    CUDA_OK(cudaFuncSetAttribute(conv_gather_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LAVAE_SMEM_LIMIT));
    CUDA_OK(cudaFuncSetAttribute(conv_gather_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LAVAE_SMEM_LIMIT));
    CUDA_OK(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LAVAE_SMEM_LIMIT));
    g_lavae_inited[dev] = true;
    return T2S_OK;
}

struct Dims {
    int B, C, L, H, R, E, F, NR;     // batch, series channels, length, hidden, residual hidden, embedding, latent positions, residual layers
    int T1, n, nd, Lr;               // encoder resolutions L -> T1 -> n ; decoder nd = int(L / 4) -> 2 nd -> Lr = 4 nd (-> L)
};
int make_dims(const t2s_lavae_params* P, int batch, int length, Dims* d) {
    if (!P || batch <= 0 || length < 4) return fail(T2S_EINVAL, "lavae: bad batch / length%s%s");
    if (P->in_channels < 1 || P->hidden < 2 || (P->hidden & 1) || P->res_hidden < 1 || P->emb < 1 || P->n_res < 0 || P->n_res > MAX_RES || P->flow_dim < 1)
        return fail(T2S_EINVAL, "lavae: bad architecture fields%s%s");
    d->B = batch; d->C = P->in_channels; d->L = length; d->H = P->hidden; d->R = P->res_hidden; d->E = P->emb; d->F = P->flow_dim; d->NR = P->n_res;
    d->T1 = (length + 2 - 4) / 2 + 1;
    d->n = (d->T1 + 2 - 4) / 2 + 1;
    d->nd = length / 4;
    d->Lr = 4 * d->nd;
    if (d->n < 1 || d->nd < 1) return fail(T2S_EINVAL, "lavae: series too short%s%s");
    // per-thread accumulator budget of conv_wgrad_kernel: 8 * Cx * k <= 4096
    const int worst = d->H * 4 > d->R ? d->H * 4 : d->R;
    if (worst > 512 || d->C * 4 > 512) return fail(T2S_EINVAL, "lavae: channel counts beyond the weight-gradient kernel's budget (Cx * k <= 512)%s%s");
    // the gather kernels stage 32 output channels' weights + one sample's layer input in shared memory
    const int nm = d->n > d->nd ? d->n : d->nd, t1m = d->T1 > 2 * d->nd ? d->T1 : 2 * d->nd, lm = d->L > d->Lr ? d->L : d->Lr;
    const size_t worst_n = (size_t)(d->H > d->R ? d->H : d->R) * (3 * CG_OT + nm), worst_1 = (size_t)(d->H / 2) * (4 * CG_OT + t1m),
                 worst_0 = (size_t)d->C * (4 * CG_OT + lm);
    if (worst_n * 4 > LAVAE_SMEM_LIMIT || worst_1 * 4 > LAVAE_SMEM_LIMIT || worst_0 * 4 > LAVAE_SMEM_LIMIT || 8 * lm > 256 * CG_MAXU)
        return fail(T2S_EINVAL, "lavae: series too long for the shared-memory staging of the layer kernels%s%s");
    return T2S_OK;
}

struct Stack { float *rin[MAX_RES], *h[MAX_RES], *x[MAX_RES + 1], *out; };    // x[0] = the stack's input (owned by the caller)
struct Acts {
    float *c1, *c2, *c3, *before, *z, *after, *d1, *t1, *rec0, *rec;
    Stack es, ds;
    // gradient temporaries
    float *gH[2], *gR, *gE, *gEn, *gZ, *gT1, *gRec, *gC1, *gRec0;
    size_t total;
};
Acts make_acts(void* base, const Dims& d) {
    Acts a;
    char* b = static_cast<char*>(base);
    size_t p = 0;
    auto take = [&](size_t floats) { float* r = reinterpret_cast<float*>(b + p); p = align256(p + floats * 4); return r; };
    const size_t B = d.B;
    const int nm = d.n > d.nd ? d.n : d.nd;
    a.c1 = take(B * (d.H / 2) * d.T1); a.c2 = take(B * d.H * d.n); a.c3 = take(B * d.H * d.n);
    for (int i = 0; i < d.NR; ++i) { a.es.rin[i] = take(B * d.H * d.n); a.es.h[i] = take(B * d.R * d.n); a.es.x[i + 1] = take(B * d.H * d.n); }
    a.es.x[0] = a.c3; a.es.out = take(B * d.H * d.n);
    a.before = take(B * d.E * d.n); a.z = take(B * d.E * d.F);
    a.after = take(B * d.E * d.nd); a.d1 = take(B * d.H * d.nd);
    for (int i = 0; i < d.NR; ++i) { a.ds.rin[i] = take(B * d.H * d.nd); a.ds.h[i] = take(B * d.R * d.nd); a.ds.x[i + 1] = take(B * d.H * d.nd); }
    a.ds.x[0] = a.d1; a.ds.out = take(B * d.H * d.nd);
    a.t1 = take(B * (d.H / 2) * 2 * d.nd); a.rec0 = take(B * d.C * d.Lr); a.rec = take(B * d.C * d.L);
    a.gH[0] = take(B * d.H * nm); a.gH[1] = take(B * d.H * nm); a.gR = take(B * d.R * nm);
    a.gE = take(B * d.E * nm); a.gEn = take(B * d.E * nm); a.gZ = take(B * d.E * d.F);
    const size_t t1max = (size_t)(d.H / 2) * (2 * d.nd > d.T1 ? 2 * d.nd : d.T1);
    a.gT1 = take(B * t1max); a.gRec = take(B * d.C * d.L); a.gRec0 = take(B * d.C * d.Lr); a.gC1 = take(B * (d.H / 2) * d.T1);
    a.total = p;
    return a;
}

struct Layer { int Cin, Tin, Cout, Tout, k, s, p; };
unsigned ew_grid(size_t n) { const size_t g = (n + 255) / 256; return (unsigned)(g > 2368 ? 2368 : (g < 1 ? 1 : g)); }

// one gather launch: CTAs = (sample chunk, 32 output channels); samples per CTA chosen for about two waves of CTAs
template <bool FORM_B>
int gather_launch(const float* in, const float* w, const float* bias, const float* res, float* out, int B, int Cin, int Tin, int Cout, int Tout,
                  int k, int s, int p, int so, int si, int relu_out, cudaStream_t st) {
    if ((CG_OT / 4) * Tout > 256 * CG_MAXU) return fail(T2S_EINVAL, "lavae: layer too long for the gather kernel%s%s");
    const int tiles = (Cout + CG_OT - 1) / CG_OT;
    int chunks = (296 + tiles - 1) / tiles;
    if (chunks > B) chunks = B;
    const int bchunk = (B + chunks - 1) / chunks;
    ConvGeom g{Cin, Tin, Cout, Tout, k, s, p, so, si, B, bchunk};
    const size_t smem = ((size_t)Cin * k * CG_OT + (size_t)Cin * Tin) * 4;
    conv_gather_kernel<FORM_B><<<dim3((B + bchunk - 1) / bchunk, tiles), 256, smem, st>>>(in, w, bias, res, out, g, 0, relu_out);
    CUDA_OK(cudaGetLastError());
    return T2S_OK;
}
// Conv1d forward (form A)
int conv_fwd(const float* in, const float* w, const float* bias, const float* res, float* out, int B, const Layer& l, int relu_out, cudaStream_t st) {
    return gather_launch<false>(in, w, bias, res, out, B, l.Cin, l.Tin, l.Cout, l.Tout, l.k, l.s, l.p, l.Cin * l.k, l.k, relu_out, st);
}
// Conv1d input gradient (form B): din [Cin][Tin] from dout [Cout][Tout]
int conv_dx(const float* dout, const float* w, float* din, int B, const Layer& l, cudaStream_t st) {
    return gather_launch<true>(dout, w, nullptr, nullptr, din, B, l.Cout, l.Tout, l.Cin, l.Tin, l.k, l.s, l.p, l.k, l.Cin * l.k, 0, st);
}
// ConvTranspose1d forward (form B): weight [Cin][Cout][k]
int convT_fwd(const float* in, const float* w, const float* bias, float* out, int B, const Layer& l, int relu_out, cudaStream_t st) {
    return gather_launch<true>(in, w, bias, nullptr, out, B, l.Cin, l.Tin, l.Cout, l.Tout, l.k, l.s, l.p, l.k, l.Cout * l.k, relu_out, st);
}
// ConvTranspose1d input gradient (form A)
int convT_dx(const float* dout, const float* w, float* din, int B, const Layer& l, cudaStream_t st) {
    return gather_launch<false>(dout, w, nullptr, nullptr, din, B, l.Cout, l.Tout, l.Cin, l.Tin, l.k, l.s, l.p, l.Cout * l.k, l.k, 0, st);
}
int wgrad_launch(const float* Y, const float* X, float* dw, int B, int Cy, int Ty, int Cx, int Tx, int k, int s, int p, int sy, int sx, cudaStream_t st) {
    int chunks = (592 * 8 + Cy - 1) / Cy;                      // about four waves of CTAs
    if (chunks > B) chunks = B;
    if (chunks < 1) chunks = 1;
    const int bchunk = (B + chunks - 1) / chunks;
    WgradGeom g{Cy, Ty, Cx, Tx, k, s, p, sy, sx, bchunk, B};
    const size_t smem = ((size_t)8 * Ty + (size_t)Cx * (Tx + 1)) * 4;
    conv_wgrad_kernel<<<dim3((B + bchunk - 1) / bchunk, (Cy + 7) / 8), 256, smem, st>>>(Y, X, dw, g, 0);
    CUDA_OK(cudaGetLastError());
    return T2S_OK;
}
// Conv1d parameter gradients: dW [Cout][Cin][k], db [Cout]
int conv_dw(const float* dout, const float* in, float* dw, float* db, int B, const Layer& l, cudaStream_t st) {
    TRY(wgrad_launch(dout, in, dw, B, l.Cout, l.Tout, l.Cin, l.Tin, l.k, l.s, l.p, l.Cin * l.k, l.k, st));
    if (db != nullptr) chan_sum_kernel<<<l.Cout, 256, 0, st>>>(dout, db, B, l.Cout, l.Tout);
    CUDA_OK(cudaGetLastError());
    return T2S_OK;
}
// ConvTranspose1d parameter gradients: dW [Cin][Cout][k], db [Cout]
int convT_dw(const float* dout, const float* in, float* dw, float* db, int B, const Layer& l, cudaStream_t st) {
    TRY(wgrad_launch(in, dout, dw, B, l.Cin, l.Tin, l.Cout, l.Tout, l.k, l.s, l.p, l.Cout * l.k, l.k, st));
    if (db != nullptr) chan_sum_kernel<<<l.Cout, 256, 0, st>>>(dout, db, B, l.Cout, l.Tout);
    CUDA_OK(cudaGetLastError());
    return T2S_OK;
}

// ResidualStack.forward (vqvae.py:7-33): the in-place ReLU makes every skip connection carry relu(x)
int stack_fwd(float* const* w3, float* const* w1, const Stack& s, const Dims& d, int T, cudaStream_t st) {
    const size_t nH = (size_t)d.B * d.H * T;
    const Layer l3{d.H, T, d.R, T, 3, 1, 1}, l1{d.R, T, d.H, T, 1, 1, 0};
    for (int i = 0; i < d.NR; ++i) {
        relu_kernel<<<ew_grid(nH), 256, 0, st>>>(s.x[i], s.rin[i], nH);
        TRY(conv_fwd(s.rin[i], w3[i], nullptr, nullptr, s.h[i], d.B, l3, 1, st));
        TRY(conv_fwd(s.h[i], w1[i], nullptr, s.rin[i], s.x[i + 1], d.B, l1, 0, st));
    }
    relu_kernel<<<ew_grid(nH), 256, 0, st>>>(s.x[d.NR], s.out, nH);
    CUDA_OK(cudaGetLastError());
    return T2S_OK;
}
// g (in: d out, out: d x[0]); tmp, gr: scratch of the same size / of R channels
int stack_bwd(float* const* w3, float* const* w1, float* const* dw3, float* const* dw1, const Stack& s, const Dims& d, int T, float*& g, float*& tmp,
              float* gr, cudaStream_t st) {
    const size_t nH = (size_t)d.B * d.H * T, nR = (size_t)d.B * d.R * T;
    const Layer l3{d.H, T, d.R, T, 3, 1, 1}, l1{d.R, T, d.H, T, 1, 1, 0};
    relu_bwd_kernel<<<ew_grid(nH), 256, 0, st>>>(g, s.out, nullptr, nH);
    for (int i = d.NR - 1; i >= 0; --i) {
        TRY(conv_dw(g, s.h[i], dw1[i], nullptr, d.B, l1, st));
        TRY(conv_dx(g, w1[i], gr, d.B, l1, st));
        relu_bwd_kernel<<<ew_grid(nR), 256, 0, st>>>(gr, s.h[i], nullptr, nR);
        TRY(conv_dw(gr, s.rin[i], dw3[i], nullptr, d.B, l3, st));
        TRY(conv_dx(gr, w3[i], tmp, d.B, l3, st));
        relu_bwd_kernel<<<ew_grid(nH), 256, 0, st>>>(tmp, s.rin[i], g, nH);       // d x_i = (rin > 0) (conv path + skip)
        float* t = g; g = tmp; tmp = t;
    }
    CUDA_OK(cudaGetLastError());
    return T2S_OK;
}

struct Layers { Layer e1, e2, e3, pre, dc1, ct1, ct2; };
Layers make_layers(const Dims& d) {
    Layers l;
    l.e1 = Layer{d.C, d.L, d.H / 2, d.T1, 4, 2, 1};
    l.e2 = Layer{d.H / 2, d.T1, d.H, d.n, 4, 2, 1};
    l.e3 = Layer{d.H, d.n, d.H, d.n, 3, 1, 1};
    l.pre = Layer{d.H, d.n, d.E, d.n, 1, 1, 0};
    l.dc1 = Layer{d.E, d.nd, d.H, d.nd, 3, 1, 1};
    l.ct1 = Layer{d.H, d.nd, d.H / 2, 2 * d.nd, 4, 2, 1};
    l.ct2 = Layer{d.H / 2, 2 * d.nd, d.C, d.Lr, 4, 2, 1};
    return l;
}

// Encoder.forward (vqvae.py:57-71 / myvqvae.py:49-61)
int enc_forward(const t2s_lavae_params* P, const float* x, const Acts& a, const Dims& d, const Layers& l, cudaStream_t st) {
    TRY(conv_fwd(x, P->enc_conv1_w, P->enc_conv1_b, nullptr, a.c1, d.B, l.e1, 1, st));
    TRY(conv_fwd(a.c1, P->enc_conv2_w, P->enc_conv2_b, nullptr, a.c2, d.B, l.e2, 1, st));
    TRY(conv_fwd(a.c2, P->enc_conv3_w, P->enc_conv3_b, nullptr, a.c3, d.B, l.e3, 0, st));
    TRY(stack_fwd(P->enc_res_w3, P->enc_res_w1, a.es, d, d.n, st));
    TRY(conv_fwd(a.es.out, P->enc_pre_w, P->enc_pre_b, nullptr, a.before, d.B, l.pre, 0, st));
    const size_t rows = (size_t)d.B * d.E;
    interp_fwd_kernel<<<ew_grid(rows * d.F), 256, 0, st>>>(a.before, a.z, rows, d.n, d.F);
    CUDA_OK(cudaGetLastError());
    return T2S_OK;
}
// Decoder.forward (vqvae.py:97-105 / myvqvae.py:76-86) from z [B][E][F]
int dec_forward(const t2s_lavae_params* P, const float* z, const Acts& a, const Dims& d, const Layers& l, cudaStream_t st) {
    const size_t rows = (size_t)d.B * d.E;
    interp_fwd_kernel<<<ew_grid(rows * d.nd), 256, 0, st>>>(z, a.after, rows, d.F, d.nd);
    TRY(conv_fwd(a.after, P->dec_conv1_w, P->dec_conv1_b, nullptr, a.d1, d.B, l.dc1, 0, st));
    TRY(stack_fwd(P->dec_res_w3, P->dec_res_w1, a.ds, d, d.nd, st));
    TRY(convT_fwd(a.ds.out, P->dec_ct1_w, P->dec_ct1_b, a.t1, d.B, l.ct1, 1, st));
    TRY(convT_fwd(a.t1, P->dec_ct2_w, P->dec_ct2_b, a.rec0, d.B, l.ct2, 0, st));
    if (d.Lr != d.L) {                                           // myvqvae.py:85 (identity when 4 int(L/4) == L)
        const size_t r2 = (size_t)d.B * d.C;
        interp_fwd_kernel<<<ew_grid(r2 * d.L), 256, 0, st>>>(a.rec0, a.rec, r2, d.Lr, d.L);
    }
    CUDA_OK(cudaGetLastError());
    return T2S_OK;
}

int check_ws(const void* ws, size_t bytes, size_t need) {
    if (ws == nullptr) return fail(T2S_EINVAL, "workspace is NULL%s%s");
    if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return fail(T2S_EINVAL, "workspace must be 256-byte aligned%s%s");
    if (bytes < need) return fail(T2S_EWORKSPACE, "lavae workspace too small%s%s");
    return T2S_OK;
}
int copy_out(float* dst, const float* src, size_t floats, cudaStream_t st) {
    if (dst != nullptr) CUDA_OK(cudaMemcpyAsync(dst, src, floats * 4, cudaMemcpyDeviceToDevice, st));
    return T2S_OK;
}

}  // namespace

extern "C" {

size_t t2s_lavae_workspace_bytes(const t2s_lavae_params* P, int batch, int length) {
    Dims d;
    if (make_dims(P, batch, length, &d) != T2S_OK) return 0;
    return make_acts(nullptr, d).total;
}

int t2s_lavae_encode(const t2s_lavae_params* P, const float* x, float* z, float* before, int batch, int length, void* workspace,
                     size_t workspace_bytes, t2s_stream_t stream) {
    Dims d;
    TRY(make_dims(P, batch, length, &d));
    if (!x || !z) return fail(T2S_EINVAL, "t2s_lavae_encode: bad argument%s%s");
    TRY(check_ws(workspace, workspace_bytes, make_acts(nullptr, d).total));
    TRY(lavae_init());
    cudaStream_t st = (cudaStream_t)stream;
    const Acts a = make_acts(workspace, d);
    TRY(enc_forward(P, x, a, d, make_layers(d), st));
    TRY(copy_out(z, a.z, (size_t)d.B * d.E * d.F, st));
    return copy_out(before, a.before, (size_t)d.B * d.E * d.n, st);
}

int t2s_lavae_decode(const t2s_lavae_params* P, const float* z, float* recon, float* after, int batch, int length, void* workspace,
                     size_t workspace_bytes, t2s_stream_t stream) {
    Dims d;
    TRY(make_dims(P, batch, length, &d));
    if (!z || !recon) return fail(T2S_EINVAL, "t2s_lavae_decode: bad argument%s%s");
    TRY(check_ws(workspace, workspace_bytes, make_acts(nullptr, d).total));
    TRY(lavae_init());
    cudaStream_t st = (cudaStream_t)stream;
    const Acts a = make_acts(workspace, d);
    TRY(dec_forward(P, z, a, d, make_layers(d), st));
    TRY(copy_out(recon, d.Lr != d.L ? a.rec : a.rec0, (size_t)d.B * d.C * d.L, st));
    return copy_out(after, a.after, (size_t)d.B * d.E * d.nd, st);
}

int t2s_lavae_train_step(const t2s_lavae_params* P, const t2s_lavae_params* G, const float* x, float* recon, float* z, float* loss_sums,
                         int batch, int length, void* workspace, size_t workspace_bytes, t2s_stream_t stream) {
    Dims d;
    TRY(make_dims(P, batch, length, &d));
    if (!x || !loss_sums) return fail(T2S_EINVAL, "t2s_lavae_train_step: bad argument%s%s");
    if (d.n != d.nd) return fail(T2S_EINVAL, "t2s_lavae_train_step: the cross loss needs equal encoder / decoder resolutions (length % 4 == 0)%s%s");
    TRY(check_ws(workspace, workspace_bytes, make_acts(nullptr, d).total));
    TRY(lavae_init());
    cudaStream_t st = (cudaStream_t)stream;
    const Acts a = make_acts(workspace, d);
    const Layers l = make_layers(d);
    // ---- forward (vqvae.py:121-125)
    TRY(enc_forward(P, x, a, d, l, st));
    TRY(dec_forward(P, a.z, a, d, l, st));
    const float* rec = d.Lr != d.L ? a.rec : a.rec0;
    const size_t n_rec = (size_t)d.B * d.C * d.L, n_cross = (size_t)d.B * d.E * d.n;
    const bool bwd = G != nullptr;
    mse_kernel<<<ew_grid(n_rec) > 296 ? 296 : ew_grid(n_rec), 256, 0, st>>>(rec, x, n_rec, (float)(2.0 / (double)n_rec), loss_sums, bwd ? a.gRec : nullptr, nullptr);
    mse_kernel<<<ew_grid(n_cross) > 296 ? 296 : ew_grid(n_cross), 256, 0, st>>>(a.before, a.after, n_cross, (float)(2.0 / (double)n_cross), loss_sums + 1,
                                                                              bwd ? a.gE : nullptr, bwd ? a.gEn : nullptr);
    CUDA_OK(cudaGetLastError());
    TRY(copy_out(recon, rec, n_rec, st));
    TRY(copy_out(z, a.z, (size_t)d.B * d.E * d.F, st));
    if (!bwd) return T2S_OK;
    // ---- backward of loss = mse(recon, x) + mse(before, after)     (a.gE = dL/dbefore, a.gEn = dL/dafter from the cross term)
    const size_t rowsE = (size_t)d.B * d.E, rowsC = (size_t)d.B * d.C;
    float* grec0 = a.gRec;
    if (d.Lr != d.L) {
        interp_bwd_kernel<<<ew_grid(rowsC * d.Lr), 256, 0, st>>>(a.gRec, a.gRec0, rowsC, d.Lr, d.L, 0);
        grec0 = a.gRec0;
    }
    TRY(convT_dw(grec0, a.t1, G->dec_ct2_w, G->dec_ct2_b, d.B, l.ct2, st));
    TRY(convT_dx(grec0, P->dec_ct2_w, a.gT1, d.B, l.ct2, st));
    relu_bwd_kernel<<<ew_grid((size_t)d.B * (d.H / 2) * 2 * d.nd), 256, 0, st>>>(a.gT1, a.t1, nullptr, (size_t)d.B * (d.H / 2) * 2 * d.nd);
    TRY(convT_dw(a.gT1, a.ds.out, G->dec_ct1_w, G->dec_ct1_b, d.B, l.ct1, st));
    float *g = a.gH[0], *tmp = a.gH[1];
    TRY(convT_dx(a.gT1, P->dec_ct1_w, g, d.B, l.ct1, st));
    TRY(stack_bwd(P->dec_res_w3, P->dec_res_w1, G->dec_res_w3, G->dec_res_w1, a.ds, d, d.nd, g, tmp, a.gR, st));
    TRY(conv_dw(g, a.after, G->dec_conv1_w, G->dec_conv1_b, d.B, l.dc1, st));
    // d after = cross term + conv path (res = out = gEn: every element is read and written by the same thread)
    TRY(gather_launch<true>(g, P->dec_conv1_w, nullptr, a.gEn, a.gEn, d.B, l.dc1.Cout, l.dc1.Tout, l.dc1.Cin, l.dc1.Tin, 3, 1, 1, 3, l.dc1.Cin * 3, 0, st));
    interp_bwd_kernel<<<ew_grid(rowsE * d.F), 256, 0, st>>>(a.gEn, a.gZ, rowsE, d.F, d.nd, 0);
    // encoder: d before = cross term + interp^T(d z)
    interp_bwd_kernel<<<ew_grid(rowsE * d.n), 256, 0, st>>>(a.gZ, a.gE, rowsE, d.n, d.F, 1);
    CUDA_OK(cudaGetLastError());
    TRY(conv_dw(a.gE, a.es.out, G->enc_pre_w, G->enc_pre_b, d.B, l.pre, st));
    g = a.gH[0]; tmp = a.gH[1];
    TRY(conv_dx(a.gE, P->enc_pre_w, g, d.B, l.pre, st));
    TRY(stack_bwd(P->enc_res_w3, P->enc_res_w1, G->enc_res_w3, G->enc_res_w1, a.es, d, d.n, g, tmp, a.gR, st));
    TRY(conv_dw(g, a.c2, G->enc_conv3_w, G->enc_conv3_b, d.B, l.e3, st));
    TRY(conv_dx(g, P->enc_conv3_w, tmp, d.B, l.e3, st));
    relu_bwd_kernel<<<ew_grid((size_t)d.B * d.H * d.n), 256, 0, st>>>(tmp, a.c2, nullptr, (size_t)d.B * d.H * d.n);
    TRY(conv_dw(tmp, a.c1, G->enc_conv2_w, G->enc_conv2_b, d.B, l.e2, st));
    TRY(conv_dx(tmp, P->enc_conv2_w, a.gC1, d.B, l.e2, st));
    relu_bwd_kernel<<<ew_grid((size_t)d.B * (d.H / 2) * d.T1), 256, 0, st>>>(a.gC1, a.c1, nullptr, (size_t)d.B * (d.H / 2) * d.T1);
    TRY(conv_dw(a.gC1, x, G->enc_conv1_w, G->enc_conv1_b, d.B, l.e1, st));
    CUDA_OK(cudaGetLastError());
    return T2S_OK;
}

}  // extern "C"
