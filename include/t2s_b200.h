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
    /* fp16 weight stages, each a [128 n][128 k] operand image in the tcgen05 no-swizzle K-major canonical layout:
     * element (n,k) at byte (k/8)*2048 + (n/8)*128 + (n%8)*16 + (k%8)*2 */
    const void* w_qkv[4];   /* 3 stages: q | k | v rows of layers.{l}.attn.qkv.weight */
    const void* w_post[4];  /* 5 stages: attn.proj | mlp.fc1 rows 0..127 | rows 128..255 | mlp.fc2 cols 0..127 | cols 128..255 */
    const float* b_qkv[4];  /* [384] */
    const float* b_proj[4]; /* [128] */
    const float* b_fc1[4];  /* [256] */
    const float* b_fc2[4];  /* [128] */
    const float* w_ada_t;   /* [4][128][768]  adaLN_modulation.1.weight transposed */
    const float* b_ada;     /* [4][768] */
    const float* w_embed;   /* [128][4]  patch_emb.weight @ conv.weight.view(4,4) */
    const float* b_embed;   /* [128]     patch_emb.weight @ conv.bias + patch_emb.bias */
    const float* pos;       /* [8 tiles][32 col chunks][64 rows][4] pos_embed in the residual tile layout */
    const float* w_final;   /* [4][128]  linear_emb_to_patch.weight * ln.weight */
    const float* b_final;   /* [4]       linear_emb_to_patch.weight @ ln.bias + linear_emb_to_patch.bias */
    const float* freqs;     /* [64]      10000 ** linspace(0,1,64)   (TimeEmbedding, transformer.py:34) */
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

/* Profiling aid: when non-NULL, every token-block CTA writes clock64() stamps of its phase boundaries to
 * device_buf[blockIdx.x*32 + i] (see tools/phase_trace.py).  NULL (default) switches it off. */
void t2s_debug_set_phase_trace(long long* device_buf);

/* Bytes of scratch for `nseq` sequences (residual stream, q|k|v, attention output, modulation).
 * Contents need no initialisation; rows of partially filled tiles are never read back. */
size_t t2s_dit_workspace_bytes(int nseq);

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

/* Decoder.forward (model/pretrained/vqvae.py:97-105): z [batch][64][30] -> series [batch][length]
 * (length in {24,48,96}), after [batch][64][length/4] or NULL. */
int t2s_vae_decode(const t2s_vae_dec_weights* w, const float* z, float* series, float* after, int batch, int length,
                   t2s_stream_t stream);

/* Encoder.forward (model/pretrained/vqvae.py:57-71): x [batch][length] -> z [batch][64][30],
 * before [batch][64][length/4] or NULL. */
int t2s_vae_encode(const t2s_vae_enc_weights* w, const float* x, float* z, float* before, int batch, int length,
                   t2s_stream_t stream);

/* Single stages of the denoiser, exported for stage-wise parity tests.  Workspace layout:
 * t2s_dit_workspace_offsets() fills {h, qkv, o, mod} byte offsets. */
void t2s_dit_workspace_offsets(int nseq, size_t offsets[4]);
int t2s_dit_cond(const t2s_dit_weights* w, const float* t100, int t_stride, const float* emb, int cfg_pairs,
                 int nseq, void* workspace, t2s_stream_t stream);
int t2s_dit_embed_qkv(const t2s_dit_weights* w, const float* x, int x_shared, int nseq, void* workspace, t2s_stream_t stream);
int t2s_dit_attention(int nseq, void* workspace, t2s_stream_t stream);
int t2s_dit_block_post(const t2s_dit_weights* w, int layer, int nseq, void* workspace, t2s_stream_t stream);
int t2s_dit_final(const t2s_dit_weights* w, float* out, int nseq, void* workspace, t2s_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* T2S_B200_H */
