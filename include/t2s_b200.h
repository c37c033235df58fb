/* t2s_b200 — C ABI of the B200-native T2S generation hot path.
 *
 * Every entry point is enqueue-only on the caller's CUDA stream (no host synchronisation, no
 * allocation, no global state besides per-device kernel attributes), takes plain device pointers
 * and sizes, and returns 0 on success or a negative T2S_E* code (t2s_last_error() gives the text).
 * The caller owns every buffer.  Citations name the reference interface each entry replaces
 * (paths relative to the Bill9125/T2MS repo root).
 */
#ifndef T2S_B200_H
#define T2S_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* t2s_stream_t; /* cudaStream_t */

enum {
    T2S_OK = 0,
    T2S_EINVAL = -1,   /* bad shape / null pointer / misaligned buffer */
    T2S_EWORKSPACE = -2, /* workspace too small */
    T2S_ECUDA = -3,    /* CUDA launch / runtime failure */
    T2S_EARCH = -4     /* device is not sm_100 */
};

/* DiT weights, packed by the host (t2ms_b200/packing.py) from the reference state dict
 * (model/denoiser/transformer.py:127-154; key names in SURVEY.md §8b).  All device pointers. */
typedef struct {
    /* fp16 weight stages of 36,864 B each: a [128 n][128 k] operand image in the tcgen05 no-swizzle K-major canonical layout
     * (element (n,k) at byte (k/8)*2048 + (n/8)*128 + (n%8)*16 + (k%8)*2; 32,768 B) followed by the Linear's bias as a
     * [128 n][16 k] operand block in the same layout (4,096 B): k = 0 holds fp16(bias[n]), k = 1 holds
     * fp16(bias[n] - fp16(bias[n])), k = 2..15 are zero (the GEMM adds the bias with one extra K = 16 MMA against a block
     * of ones; mlp.fc2's second K half carries a zero block) */
    const void* w_qkv[4];   /* 3 stages: q | k | v rows of layers.{l}.attn.qkv.weight (+ attn.qkv.bias) */
    const void* w_post[4];  /* 5 stages: attn.proj | mlp.fc1 rows 0..127 | rows 128..255 | mlp.fc2 cols 0..127 | cols 128..255 */
    /* the same biases in fp32 (the fused step kernel and the stage-wise debugging entries read them) */
    const float* b_qkv[4];  /* [384] */
    const float* b_proj[4]; /* [128] */
    const float* b_fc1[4];  /* [256] */
    const float* b_fc2[4];  /* [128] */
    const float* w_ada_t;   /* [4][128][768]  adaLN_modulation.1.weight transposed */
    const float* b_ada;     /* [4][768] */
    const float* w_embed;   /* [4][128]  (patch_emb.weight @ conv.weight.view(4,4))^T: pixel-major */
    const float* b_embed;   /* [128]     patch_emb.weight @ conv.bias + patch_emb.bias */
    const float* pos;       /* [tiles][32 col chunks][64 rows][4] pos_embed in the residual tile layout (8 tiles of 60 tokens for H = 30) */
    const float* w_final;   /* [4][128]  linear_emb_to_patch.weight * ln.weight */
    const float* b_final;   /* [4]       linear_emb_to_patch.weight @ ln.bias + linear_emb_to_patch.bias */
    const float* freqs;     /* [64]      10000 ** linspace(0,1,64)   (TimeEmbedding, transformer.py:34) */
    /* latent width H (axis 2 of the (B,64,H) latent): 0 or 30 = T2S (transformer.py:132); 50 / 64 = the fork's
     * Transformer(dim) (model/denoiser/mytransformer.py:128-136 with config.yaml:46,91 flow_dim).  Tokens = 16 H;
     * pos is [16 H / T tiles][32][64][4] with T = 60 | 50 | 64 tokens per tile; every latent buffer of the calls
     * below is [..][64][H]. */
    int latent_h;
    /* The same fp16 weights cut into 16 KB HALF stages [64 n][128 k] — element (n,k) at byte
     * (k/8)*1024 + (n/8)*128 + (n%8)*16 + (k%8)*2 — for the fused per-step kernel (t2s_dit_fused_step, H = 30), whose
     * three-slot weight ring holds half stages.  NULL = not packed: the fused kernel is not used.
     *   w_qkv_half[l]:  6 halves: q rows 0..63 | q rows 64..127 | k 0..63 | k 64..127 | v 0..63 | v 64..127
     *   w_post_half[l]: 10 halves: proj h0 h1 | fc1 rows 0..127 h0 h1 | fc1 rows 128..255 h0 h1 |
     *                   fc2 (cols 0..127, h0) (cols 128..255, h0) (cols 0..127, h1) (cols 128..255, h1)      */
    const void* w_qkv_half[4];
    const void* w_post_half[4];
} t2s_dit_weights;

/* LA-VAE decoder / encoder weights (model/pretrained/vqvae.py:36-105), fp32, [ic][k][oc] layouts. */
typedef struct {
    const float* conv1_w;   /* [64][3][128] */
    const float* conv1_b;   /* [128] */
    const float* res_w3[2]; /* [128][3][256] */
    const float* res_w1[2]; /* [256][128] */
    const float* ct1_w;     /* [128][4][64] */
    const float* ct1_b;     /* [64] */
    const float* ct2_w;     /* [64][4] */
    const float* ct2_b;     /* [1] */
} t2s_vae_dec_weights;

typedef struct {
    const float* conv1_w;   /* [1][4][64] */
    const float* conv1_b;   /* [64] */
    const float* conv2_w;   /* [64][4][128] */
    const float* conv2_b;   /* [128] */
    const float* conv3_w;   /* [128][3][128] */
    const float* conv3_b;   /* [128] */
    const float* res_w3[2]; /* [128][3][256] */
    const float* res_w1[2]; /* [256][128] */
    const float* pre_w;     /* [128][64] */
    const float* pre_b;     /* [64] */
} t2s_vae_enc_weights;

int t2s_version(void);
const char* t2s_last_error(void);

/* Sets kernel attributes for the current device; call once per device before stream capture. */
int t2s_init(void);

/* The fused per-step kernel (one persistent cooperative launch per guided step: attention and token work of different
 * sequence pairs co-resident on every SM, coupled by a dataflow scheduler; csrc/dit_fused.cuh) serves the T2S shape from
 * `min_pairs` sequence pairs on; -1 (the default) = never: the per-phase kernels run, which measured faster (DESIGN.md
 * §4.9).  `inflight` bounds the pairs admitted and not yet finished (0 = no limit).  Process-wide switches; results are
 * the same either way (tests/test_gpu_fused.py). */
void t2s_set_fused(int min_pairs, int inflight);
/* The kernels of a denoiser evaluation are launched with programmatic stream serialization (PDL): each kernel's set-up runs
 * under the previous kernel's tail and waits (griddepcontrol.wait) before touching its data.  0 launches them plainly. */
void t2s_set_pdl(int on);
/* Profiling aid: when non-NULL, every CTA of the fused kernel writes device_buf[blockIdx.x*8 + {0: token items, 1: token
 * scheduler-starved cycles, 2: attention units, 3: attention starved cycles, 4: total cycles}]. */
void t2s_debug_set_fused_stats(long long* device_buf);
/* ... and device_buf[(blockIdx.x*16 + item)*32 + i] = clock64() at phase boundary i of its first 16 token items (slot 31: mode*16 + layer + 1). */
void t2s_debug_set_fused_trace(long long* device_buf);

/* Profiling aid: when non-NULL, every token-block CTA writes clock64() stamps of its phase boundaries to
 * device_buf[blockIdx.x*32 + i] (see tools/phase_trace.py).  NULL (default) switches it off. */
void t2s_debug_set_phase_trace(long long* device_buf);

/* Bytes of scratch for `nseq` sequences (residual stream, q|k|v, attention output, modulation).
 * Contents need no initialisation; rows of partially filled tiles are never read back. */
size_t t2s_dit_workspace_bytes(int nseq);                       /* H = 30 */
size_t t2s_dit_workspace_bytes_h(int nseq, int latent_h);       /* 0 when latent_h is unsupported */

/* Transformer.forward (model/denoiser/transformer.py:158-193).
 *   x    [nseq][64][30] fp32      latent
 *   t100 [nseq] fp32              100 * t  (TimeEmbedding scales t by 100, transformer.py:31)
 *   emb  [nseq][128] fp32 or NULL text embedding (NULL = the text_input=None branch)
 *   out  [nseq][64][30] fp32      predicted velocity / noise                                      */
int t2s_dit_forward(const t2s_dit_weights* w, const float* x, const float* t100, const float* emb, float* out,
                    int nseq, void* workspace, size_t workspace_bytes, t2s_stream_t stream);

/* The classifier-free-guided sampling loop of infer.py:76-88 for `batch` samples, fully enqueued:
 * per step 2 x batch denoiser sequences (uncond, cond), guidance mix, Euler (RectifiedFlow.euler,
 * model/backbone/rectified_flow.py:5-7) or ancestral (DDPM.p_sample, model/backbone/DDPM.py:28-36) update.
 *   x          [batch][64][30] fp32   in: initial noise (infer.py:75), out: final latent
 *   emb        [batch][128] fp32
 *   t100       [steps] fp32 DEVICE    100 * t_j  (RF: t_j = j/steps; DDPM: t_j = steps-1-j)
 *   coef       [steps][3] fp32 HOST   RF: {dt, 0, 0};  DDPM: {1/sqrt(alpha_t), (1-alpha_t)/sqrt(1-alpha_bar_t), sqrt(beta_t)}
 *   step_noise [steps][batch][64][30] fp32 DEVICE, DDPM only (replaces torch.randn of DDPM.py:35); NULL for RF
 *   pred_trace [steps][batch][64][30] fp32 or NULL: guided prediction of every step (parity tests)
 *   kind       0 = rectified flow, 1 = DDPM                                                         */
int t2s_sample(const t2s_dit_weights* w, int kind, float* x, const float* emb, const float* t100, const float* coef,
               const float* step_noise, float* pred_trace, int batch, int steps, float cfg_scale,
               void* workspace, size_t workspace_bytes, t2s_stream_t stream);

/* The DDPM loop of t2s_sample with the Gaussian noise of DDPM.p_sample (model/backbone/DDPM.py:35) generated inside the
 * update kernel: Philox4x32-10 keyed by `seed` on the counter (element index in [batch][64][H], step j, 0, 0), first two
 * words -> u1 = ((w0 >> 8) + 1) 2^-24, u2 = (w1 >> 8) 2^-24 -> sqrt(-2 ln u1) cos(2 pi u2).  No noise tensor, no RNG
 * kernel between steps.  The element index is local to the call: callers that split a batch over calls / ranks give
 * every part its own seed. */
int t2s_sample_ddpm_seeded(const t2s_dit_weights* w, float* x, const float* emb, const float* t100, const float* coef,
                           unsigned long long seed, float* pred_trace, int batch, int steps, float cfg_scale,
                           void* workspace, size_t workspace_bytes, t2s_stream_t stream);

/* Decoder.forward (model/pretrained/vqvae.py:97-105): z [batch][64][30] -> series [batch][length]
 * (length in {24,48,96}), after [batch][64][length/4] or NULL. */
int t2s_vae_decode(const t2s_vae_dec_weights* w, const float* z, float* series, float* after, int batch, int length,
                   t2s_stream_t stream);

/* Encoder.forward (model/pretrained/vqvae.py:57-71): x [batch][length] -> z [batch][64][30],
 * before [batch][64][length/4] or NULL. */
int t2s_vae_encode(const t2s_vae_enc_weights* w, const float* x, float* z, float* before, int batch, int length,
                   t2s_stream_t stream);

/* Step-wise process maths for loops that call the backbone classes once per step (infer.py:82,88), element-wise, every
 * operation rounded like the reference's torch expression:
 *   t2s_rf_euler       RectifiedFlow.euler (model/backbone/rectified_flow.py:5-7): out = x + v * dt, n elements
 *   t2s_ddpm_p_sample  DDPM.p_sample (model/backbone/DDPM.py:28-36): out = c1[b] * (xt - c2[b] * eps_hat) + c3[b] * noise with
 *                      c1 = 1/sqrt(alpha_t), c2 = (1-alpha_t)/sqrt(1-alpha_bar_t), c3 = sqrt(beta_t) gathered per sample
 *                      ([batch] device arrays), noise = the torch.randn of DDPM.py:35; [batch][elems_per_sample] tensors.
 * (RectifiedFlow.create_flow and DDPM.q_sample are t2s_train_make_inputs.) */
int t2s_rf_euler(const float* x, const float* v, float dt, float* out, size_t n, t2s_stream_t stream);
int t2s_ddpm_p_sample(const float* xt, const float* eps_hat, const float* noise, const float* c1, const float* c2, const float* c3,
                      float* out, int batch, int elems_per_sample, t2s_stream_t stream);

/* Evaluation of generated series on the device (the step after the path): evaluation.py:166-181 calculate_mse and
 * evaluation.py:184-206 calculate_wape on univariate series (the (N, L, 1) arrays written at infer.py:117-121).
 *   ori, gen    [n][length] fp32
 *   per_sample  [n][3] fp32 out: sum (ori-gen)^2, sum |ori-gen|, sum |ori| of every sample
 *   out         [3] fp64 DEVICE: MSE = mean_i mean_t (ori-gen)^2 ; WAPE = nanmean_i (sum|ori-gen| / sum|ori|) ; number of
 *               samples with a nonzero denominator (WAPE is NaN when there is none, like np.nanmean)                  */
int t2s_series_metrics(const float* ori, const float* gen, int n, int length, float* per_sample, double* out,
                       t2s_stream_t stream);

/* Single stages of the denoiser, exported for stage-wise parity tests.  Workspace layout:
 * t2s_dit_workspace_offsets() fills {h, qkv, o, mod} byte offsets. */
void t2s_dit_workspace_offsets(int nseq, size_t offsets[4]);
int t2s_dit_workspace_offsets_h(int nseq, int latent_h, size_t offsets[4]);
int t2s_dit_cond(const t2s_dit_weights* w, const float* t100, int t_stride, const float* emb, int cfg_pairs,
                 int nseq, void* workspace, t2s_stream_t stream);
int t2s_dit_embed_qkv(const t2s_dit_weights* w, const float* x, int x_shared, int nseq, void* workspace, t2s_stream_t stream);
int t2s_dit_attention(int nseq, void* workspace, t2s_stream_t stream);
int t2s_dit_attention_h(int nseq, int latent_h, void* workspace, t2s_stream_t stream);
int t2s_dit_block_post(const t2s_dit_weights* w, int layer, int nseq, void* workspace, t2s_stream_t stream);
int t2s_dit_final(const t2s_dit_weights* w, float* out, int nseq, void* workspace, t2s_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Training step (train.py:66-87): forward + MSE loss + backward of the DiT, and the AdamW update.
 * ------------------------------------------------------------------------------------------------ */

/* The DiT parameters as the reference state dict holds them (fp32, nn.Linear / nn.Conv2d layouts; key names in
 * SURVEY.md 8b).  The same struct type carries the gradient pointers (pos and freqs are ignored there). */
typedef struct {
    float* conv_w;     /* conv.weight (4,1,2,2) */
    float* conv_b;     /* conv.bias (4) */
    float* pe_w;       /* patch_emb.weight (128,4) */
    float* pe_b;       /* patch_emb.bias (128) */
    float* pos;        /* pos_embed (1,480,128), not trained (transformer.py:139-140) */
    float* ln_w;       /* ln.weight (128) */
    float* ln_b;       /* ln.bias (128) */
    float* lf_w;       /* linear_emb_to_patch.weight (4,128) */
    float* lf_b;       /* linear_emb_to_patch.bias (4) */
    float* freqs;      /* [64] 10000 ** linspace(0,1,64) (TimeEmbedding, transformer.py:34); no gradient */
    float* qkv_w[4];   /* layers.{l}.attn.qkv.weight (384,128) */
    float* qkv_b[4];   /* (384) */
    float* proj_w[4];  /* layers.{l}.attn.proj.weight (128,128) */
    float* proj_b[4];  /* (128) */
    float* fc1_w[4];   /* layers.{l}.mlp.fc1.weight (256,128) */
    float* fc1_b[4];   /* (256) */
    float* fc2_w[4];   /* layers.{l}.mlp.fc2.weight (128,256) */
    float* fc2_b[4];   /* (128) */
    float* ada_w[4];   /* layers.{l}.adaLN_modulation.1.weight (768,128) */
    float* ada_b[4];   /* (768) */
    int latent_h;      /* latent width H: 0 / 30 = T2S (pos_embed (1,480,128), latents (B,64,30)); 50 / 64 = the fork's
                        * Transformer(dim) trained by mytrain.py:23 (pos_embed (1,16 H,128), latents (B,64,H)) */
} t2s_dit_params;

/* Scratch for a training step over `nseq` sequences (saved activations of every block + backward temporaries). */
size_t t2s_train_workspace_bytes(int nseq);                        /* H = 30 */
size_t t2s_train_workspace_bytes_h(int nseq, int latent_h);        /* 0 when latent_h is unsupported */

/* pred = Transformer.forward(x_t, t, emb | None) (transformer.py:158-193), loss = F.mse_loss(pred, target)
 * (rectified_flow.py:13-16 / DDPM.py:37-38), backward to every trainable parameter (train.py:83-85).
 *   x_t, target [nseq][64][30]; t100 [nseq] = 100 * t; emb [nseq][128] or NULL (the CFG-dropped batch, train.py:80-82)
 *   loss_sum  device float, += sum (pred - target)^2   (caller zeroes; loss = loss_sum / (nseq * 1920))
 *   pred      [nseq][64][30] or NULL
 *   grads     += dL/dparam for loss = mean over loss_numel elements (pass nseq*1920 for a single batch; the global
 *             element count when the batch is split into micro-batches / data-parallel ranks).  Caller zeroes.
 * With grads == NULL only the forward and the loss are evaluated. */
int t2s_dit_train_step(const t2s_dit_params* params, const t2s_dit_params* grads, const float* x_t, const float* t100,
                       const float* emb, const float* target, float* loss_sum, float* pred, int nseq, double loss_numel,
                       void* workspace, size_t workspace_bytes, t2s_stream_t stream);

/* The same step split in two for autograd-style callers (Transformer.forward in training mode followed by
 * loss.backward(), train.py:83-85): the forward keeps every activation in `workspace`; the backward consumes
 * dL/dpred [nseq][64][30] and the SAME workspace contents, and accumulates into grads. */
int t2s_dit_train_forward(const t2s_dit_params* params, const float* x_t, const float* t100, const float* emb, float* pred,
                          int nseq, void* workspace, size_t workspace_bytes, t2s_stream_t stream);
int t2s_dit_train_backward(const t2s_dit_params* params, const t2s_dit_params* grads, const float* dpred, int nseq,
                           void* workspace, size_t workspace_bytes, t2s_stream_t stream);

/* Training inputs.  kind 0: RectifiedFlow.create_flow (rectified_flow.py:8-12) + target of train.py:71:
 *   x_t = t x1 + (1 - t) x0, target = x1 - x0, ca = t [batch].
 * kind 1: DDPM.q_sample (DDPM.py:19-27), target = eps (train.py:74-76): x_t = ca x1 + cb eps with
 *   ca = sqrt(alpha_bar_t), cb = sqrt(1 - alpha_bar_t) per sample.   x1, noise, x_t, target: [batch][64][30]. */
int t2s_train_make_inputs(int kind, const float* x1, const float* noise, const float* ca, const float* cb, float* x_t,
                          float* target, int batch, t2s_stream_t stream);                       /* latents (B,64,30) */
int t2s_train_make_inputs_h(int kind, const float* x1, const float* noise, const float* ca, const float* cb, float* x_t,
                            float* target, int batch, int latent_h, t2s_stream_t stream);       /* latents (B,64,H) */

/* torch.optim.AdamW single step over a flat fp32 buffer (train.py:37: lr 1e-4, betas (0.9, 0.999), eps 1e-8, wd 0):
 * g is multiplied by grad_scale first (1/world_size after a SUM all-reduce).  step counts from 1. */
int t2s_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, int step, float lr,
                   float beta1, float beta2, float eps, float weight_decay, float grad_scale, t2s_stream_t stream);

/* The fused attention of the training step (timm Attention -> F.scaled_dot_product_attention at transformer.py:116
 * and its autograd backward at train.py:83-85), exported for unit tests.  qkv / dqkv: [nseq*480][384] fp32 rows
 * (q | k | v, head h in columns 32 h .. of each part); o / dout: [nseq*480][128]; nlse: [nseq][4][480] written by the
 * forward (4 - log2-sum-exp of the scaled scores) and consumed by the backward.  scratch: caller-owned, 256-byte
 * aligned, t2s_train_attention_scratch_bytes(nseq). */
size_t t2s_train_attention_scratch_bytes(int nseq);
int t2s_train_attention_forward(const float* qkv, float* o, float* nlse, int nseq, void* scratch, size_t scratch_bytes,
                                t2s_stream_t stream);
int t2s_train_attention_backward(const float* qkv, const float* o, const float* nlse, const float* dout, float* dqkv,
                                 int nseq, void* scratch, size_t scratch_bytes, t2s_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Generic LA-VAE path (the stage before the hot path): layer-wise forward with saved activations and
 * the exact backward, fp32, reference layouts.  Covers model/pretrained/vqvae.py (in_channels 1, flow_dim 30) and the
 * fork's multivariate model/pretrained/myvqvae.py (in_channels = input_dim, flow_dim = args.flow_dim, any length).
 * The same struct type carries the gradient pointers (the int fields are ignored there).
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    int in_channels;      /* series channels: 1 (vqvae.py:114) or args.input_dim (myvqvae.py:100) */
    int hidden;           /* args.block_hidden_size */
    int res_hidden;       /* args.res_hidden_size */
    int emb;              /* args.embedding_dim */
    int n_res;            /* args.num_residual_layers (<= 4) */
    int flow_dim;         /* latent positions: 30 (vqvae.py:70) or args.flow_dim (myvqvae.py:60) */
    float* enc_conv1_w;   /* encoder._conv_1.weight (hidden/2, in_channels, 4) */
    float* enc_conv1_b;
    float* enc_conv2_w;   /* encoder._conv_2.weight (hidden, hidden/2, 4) */
    float* enc_conv2_b;
    float* enc_conv3_w;   /* encoder._conv_3.weight (hidden, hidden, 3) */
    float* enc_conv3_b;
    float* enc_res_w3[4]; /* encoder._residual_stack._layers.{i}._block.1.weight (res_hidden, hidden, 3) */
    float* enc_res_w1[4]; /* encoder._residual_stack._layers.{i}._block.3.weight (hidden, res_hidden, 1) */
    float* enc_pre_w;     /* encoder._pre_vq_conv.weight (emb, hidden, 1) */
    float* enc_pre_b;
    float* dec_conv1_w;   /* decoder._conv_1.weight (hidden, emb, 3) */
    float* dec_conv1_b;
    float* dec_res_w3[4];
    float* dec_res_w1[4];
    float* dec_ct1_w;     /* decoder._conv_trans_1.weight (hidden, hidden/2, 4) */
    float* dec_ct1_b;
    float* dec_ct2_w;     /* decoder._conv_trans_2.weight (hidden/2, in_channels, 4) */
    float* dec_ct2_b;
} t2s_lavae_params;

/* Scratch (activations of every layer + gradient temporaries) for `batch` series of `length`; 0 on a bad argument. */
size_t t2s_lavae_workspace_bytes(const t2s_lavae_params* params, int batch, int length);

/* Encoder.forward (vqvae.py:57-71, myvqvae.py:49-61): x [batch][in_channels][length] -> z [batch][emb][flow_dim],
 * before [batch][emb][n] or NULL (n = length after the two stride-2 convolutions). */
int t2s_lavae_encode(const t2s_lavae_params* params, const float* x, float* z, float* before, int batch, int length,
                     void* workspace, size_t workspace_bytes, t2s_stream_t stream);
/* Decoder.forward (vqvae.py:97-105, myvqvae.py:76-86): z [batch][emb][flow_dim] -> recon [batch][in_channels][length]
 * (incl. the final interpolation of myvqvae.py:85), after [batch][emb][length/4] or NULL. */
int t2s_lavae_decode(const t2s_lavae_params* params, const float* z, float* recon, float* after, int batch, int length,
                     void* workspace, size_t workspace_bytes, t2s_stream_t stream);
/* vqvae.shared_eval (vqvae.py:118-135, myvqvae.py:116-136): encoder, decoder, recon_error = mse(recon, x),
 * cross_loss = mse(before, after), and — when grads != NULL — the backward of loss = recon_error + cross_loss to every
 * parameter (loss.backward(), vqvae.py:127 / myvqvae.py:127).
 *   loss_sums [2] device floats, += { sum (recon - x)^2, sum (before - after)^2 }  (caller zeroes; the two means divide by
 *             batch*in_channels*length and batch*emb*(length/4))
 *   recon [batch][in_channels][length], z [batch][emb][flow_dim]: outputs, either may be NULL
 *   grads += dL/dparam (caller zeroes); the optimizer update stays with the caller (optimizer.step(), :128). */
int t2s_lavae_train_step(const t2s_lavae_params* params, const t2s_lavae_params* grads, const float* x, float* recon, float* z,
                         float* loss_sums, int batch, int length, void* workspace, size_t workspace_bytes,
                         t2s_stream_t stream);

/* The tcgen05 tf32 GEMM every training Linear runs on, exported for unit tests:
 * C[m][n] (mode 0: =, 1: +=, 2: atomic +=) alpha * sum_k A(m,k) B(n,k) (+ bias[n]);  a_mn / b_mn = 1: the operand is
 * stored transposed (element (m,k) at k*ld + m). */
int t2s_gemm_tf32(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int lda, int ldb,
                  int ldc, int a_mn, int b_mn, int mode, float alpha, int ksplit, t2s_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* T2S_B200_H */
