// LA-VAE encoder / decoder forward for sm_100a (reference: model/pretrained/vqvae.py:7-105).
// One CTA per series; every activation of the chain stays in shared memory, weights are read from
// L2 through coalesced [ic][k][oc] layouts (packed on the host).  fp32 throughout (the chain is
// 0.008 % of the sampling FLOPs; it is bounded by weight/L2 and HBM traffic, not by math).
#pragma once
#include "common.cuh"

namespace t2s {

constexpr int VH = 128;    // block_hidden_size
constexpr int VR = 256;    // res_hidden_size
constexpr int VE = 64;     // embedding_dim (= LATC)

struct VaeDecWeights {     // mirrors t2s_vae_dec_weights
    const float* conv1_w;  // [64 ic][3][128 oc]
    const float* conv1_b;  // [128]
    const float* res_w3[2];  // [128 ic][3][256 oc]   (no bias)
    const float* res_w1[2];  // [256 ic][128 oc]      (no bias)
    const float* ct1_w;    // [128 ic][4][64 oc]
    const float* ct1_b;    // [64]
    const float* ct2_w;    // [64 ic][4]
    const float* ct2_b;    // [1]
};
struct VaeEncWeights {     // mirrors t2s_vae_enc_weights
    const float* conv1_w;  // [1][4][64 oc]
    const float* conv1_b;  // [64]
    const float* conv2_w;  // [64 ic][4][128 oc]
    const float* conv2_b;  // [128]
    const float* conv3_w;  // [128 ic][3][128 oc]
    const float* conv3_b;  // [128]
    const float* res_w3[2];
    const float* res_w1[2];
    const float* pre_w;    // [128 ic][64 oc]
    const float* pre_b;    // [64]
};

// out[oc][p] = act( bias[oc] + sum_{ic,k} wt[(ic*K+k)*OC+oc] * in[ic*ldin + p*S + k] (+ res[oc][p]) )
// `in` rows carry a left halo equal to the conv padding, so tap k of output p reads index p*S+k.
// Work item = (oc, group of NP consecutive outputs); P must be a multiple of NP.
template <int K, int S, int NP>
__device__ __forceinline__ void conv_rows(const float* __restrict__ in, int ldin, int IC, const float* __restrict__ wt,
                                          const float* __restrict__ bias, int OC, int P, float* __restrict__ out, int ldout,
                                          int out_off, bool relu_out, const float* __restrict__ res, int ldres, int res_off) {
    constexpr int RW = (NP - 1) * S + K;
    const int groups = P / NP;
    for (int item = threadIdx.x; item < OC * groups; item += blockDim.x) {
        const int oc = item % OC, p0 = (item / OC) * NP;
        float acc[NP];
        const float b = bias ? bias[oc] : 0.f;
#pragma unroll
        for (int i = 0; i < NP; ++i) acc[i] = b;
        for (int ic = 0; ic < IC; ++ic) {
            float row[RW];
            const float* r = in + ic * ldin + p0 * S;
#pragma unroll
            for (int i = 0; i < RW; ++i) row[i] = r[i];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float w = wt[(ic * K + k) * OC + oc];
#pragma unroll
                for (int i = 0; i < NP; ++i) acc[i] = fmaf(w, row[i * S + k], acc[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            float v = acc[i];
            if (res) v += res[oc * ldres + res_off + p0 + i];
            if (relu_out) v = fmaxf(v, 0.f);
            out[oc * ldout + out_off + p0 + i] = v;
        }
    }
}

// ConvTranspose1d k4 s2 p1 (vqvae.py:85-93): out[2m] = w1*in[m] + w3*in[m-1], out[2m+1] = w0*in[m+1] + w2*in[m].
// `in` rows carry a 1-column halo on both sides; wt is [ic][4][OC].
template <int NP>
__device__ __forceinline__ void convT_rows(const float* __restrict__ in, int ldin, int IC, const float* __restrict__ wt,
                                           const float* __restrict__ bias, int OC, int Pin, float* __restrict__ out, int ldout,
                                           int out_off, bool relu_out) {
    const int groups = Pin / NP;
    for (int item = threadIdx.x; item < OC * groups; item += blockDim.x) {
        const int oc = item % OC, m0 = (item / OC) * NP;
        float ev[NP], od[NP];
        const float b = bias[oc];
#pragma unroll
        for (int i = 0; i < NP; ++i) ev[i] = od[i] = b;
        for (int ic = 0; ic < IC; ++ic) {
            float row[NP + 2];                       // row[q] = in position m0 - 1 + q
            const float* r = in + ic * ldin + m0;
#pragma unroll
            for (int i = 0; i < NP + 2; ++i) row[i] = r[i];
            const float w0 = wt[(ic * 4 + 0) * OC + oc], w1 = wt[(ic * 4 + 1) * OC + oc];
            const float w2 = wt[(ic * 4 + 2) * OC + oc], w3 = wt[(ic * 4 + 3) * OC + oc];
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                ev[i] = fmaf(w1, row[i + 1], fmaf(w3, row[i], ev[i]));
                od[i] = fmaf(w0, row[i + 2], fmaf(w2, row[i + 1], od[i]));
            }
        }
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            float e = ev[i], o = od[i];
            if (relu_out) { e = fmaxf(e, 0.f); o = fmaxf(o, 0.f); }
            out[oc * ldout + out_off + 2 * (m0 + i)] = e;
            out[oc * ldout + out_off + 2 * (m0 + i) + 1] = o;
        }
    }
}

// F.interpolate(mode='linear', align_corners=True) along the last axis (vqvae.py:70,98).
__device__ __forceinline__ void interp_rows(const float* __restrict__ in, int ldin, int Win, float* __restrict__ out, int ldout,
                                            int out_off, int Wout, int C) {
    const float scale = Wout > 1 ? (float)(Win - 1) / (float)(Wout - 1) : 0.f;
    for (int idx = threadIdx.x; idx < C * Wout; idx += blockDim.x) {
        const int c = idx / Wout, i = idx - c * Wout;
        const float src = scale * (float)i;
        int i0 = (int)floorf(src);
        i0 = min(i0, Win - 1);
        const float l1 = fminf(fmaxf(src - (float)i0, 0.f), 1.f), l0 = 1.f - l1;
        const int i1 = i0 + (i0 < Win - 1 ? 1 : 0);
        out[c * ldout + out_off + i] = l0 * in[c * ldin + i0] + l1 * in[c * ldin + i1];
    }
}

// ResidualStack (vqvae.py:7-33) on x = buf [128][ld] (halo 1), scratch hid [256][ld].
// Residual's first op is an in-place ReLU, so each layer is r = relu(x); x = r + W1x1 . relu(W3 * r).
template <int NP>
__device__ __forceinline__ void residual_stack(float* x, float* hid, int ld, int P, const float* const* w3, const float* const* w1) {
    for (int layer = 0; layer < 2; ++layer) {
        for (int idx = threadIdx.x; idx < VH * P; idx += blockDim.x) {
            const int c = idx / P, i = idx - c * P;
            x[c * ld + 1 + i] = fmaxf(x[c * ld + 1 + i], 0.f);
        }
        __syncthreads();
        conv_rows<3, 1, NP>(x, ld, VH, w3[layer], nullptr, VR, P, hid, ld, 1, true, nullptr, 0, 0);
        __syncthreads();
        // 1x1 conv reads hid at index p + 1 (halo offset): pass in + 1 with K = 1
        conv_rows<1, 1, NP>(hid + 1, ld, VR, w1[layer], nullptr, VH, P, x, ld, 1, layer == 1, x, ld, 1);
        __syncthreads();
    }
}

constexpr int VAE_LD_MAX = 24 + 2;
constexpr int VAE_DEC_SMEM_FLOATS = LATC * LATP + VH * VAE_LD_MAX + VR * VAE_LD_MAX + VE * (2 * 24 + 2);
constexpr int VAE_ENC_SMEM_FLOATS = (96 + 2) + VE * (2 * 24 + 2) + VH * VAE_LD_MAX + VR * VAE_LD_MAX + VH * VAE_LD_MAX;

// Decoder.forward (vqvae.py:97-105).  grid = B, block = 256 (1024 for small batches: one series per CTA, latency-bound).  z [B][64][30] -> series [B][4*L4], after [B][64][L4]
template <int L4, int NT = 256>
__global__ void __launch_bounds__(NT) vae_decode_kernel(const VaeDecWeights w, const float* __restrict__ z, float* __restrict__ series,
                                                         float* __restrict__ after) {
    extern __shared__ __align__(16) float sm[];
    constexpr int LD = L4 + 2, NP = 6;
    float* zs = sm;                          // [64][30]
    float* xa = zs + LATC * LATP;            // [128][LD]
    float* hid = xa + VH * LD;               // [256][LD]
    float* ct = hid + VR * LD;               // [64][2*L4+2]
    const int b = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < LATC * LATP; i += blockDim.x) zs[i] = z[(size_t)b * LAT + i];
    for (int i = tid; i < VR * LD; i += blockDim.x) hid[i] = 0.f;
    for (int i = tid; i < VH * LD; i += blockDim.x) xa[i] = 0.f;
    for (int i = tid; i < VE * (2 * L4 + 2); i += blockDim.x) ct[i] = 0.f;
    __syncthreads();
    // interpolate 30 -> L4 into hid[0..63] (halo 1) and emit `after`
    interp_rows(zs, LATP, LATP, hid, LD, 1, L4, LATC);
    __syncthreads();
    if (after)
        for (int i = tid; i < LATC * L4; i += blockDim.x) after[(size_t)b * LATC * L4 + i] = hid[(i / L4) * LD + 1 + (i % L4)];
    conv_rows<3, 1, NP>(hid, LD, LATC, w.conv1_w, w.conv1_b, VH, L4, xa, LD, 1, false, nullptr, 0, 0);
    __syncthreads();
    for (int i = tid; i < LATC * LD; i += blockDim.x) hid[i] = 0.f;   // restore halo/zero state of the scratch rows
    __syncthreads();
    residual_stack<NP>(xa, hid, LD, L4, w.res_w3, w.res_w1);
    convT_rows<NP>(xa, LD, VH, w.ct1_w, w.ct1_b, VE, L4, ct, 2 * L4 + 2, 1, true);
    __syncthreads();
    // conv_trans_2: 64 -> 1, out length 4*L4
    for (int o = tid; o < 4 * L4; o += blockDim.x) {
        const int m = o >> 1;
        float acc = w.ct2_b[0];
        if ((o & 1) == 0) {
            for (int ic = 0; ic < VE; ++ic)
                acc = fmaf(w.ct2_w[ic * 4 + 1], ct[ic * (2 * L4 + 2) + 1 + m], fmaf(w.ct2_w[ic * 4 + 3], ct[ic * (2 * L4 + 2) + m], acc));
        } else {
            for (int ic = 0; ic < VE; ++ic)
                acc = fmaf(w.ct2_w[ic * 4 + 0], ct[ic * (2 * L4 + 2) + 2 + m], fmaf(w.ct2_w[ic * 4 + 2], ct[ic * (2 * L4 + 2) + 1 + m], acc));
        }
        series[(size_t)b * 4 * L4 + o] = acc;
    }
}

// Encoder.forward (vqvae.py:57-71).  x [B][4*L4] -> z [B][64][30], before [B][64][L4]
template <int L4, int NT = 256>
__global__ void __launch_bounds__(NT) vae_encode_kernel(const VaeEncWeights w, const float* __restrict__ x, float* __restrict__ z,
                                                         float* __restrict__ before) {
    extern __shared__ __align__(16) float sm[];
    constexpr int L = 4 * L4, L2 = 2 * L4, LD = L4 + 2, NP = 6;
    float* xs = sm;                         // [L+2]
    float* c1 = xs + (96 + 2);              // [64][L2+2]
    float* xa = c1 + VE * (2 * 24 + 2);     // [128][LD]
    float* hid = xa + VH * VAE_LD_MAX;      // [256][LD]
    float* xb = hid + VR * VAE_LD_MAX;      // [128][LD]
    const int b = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < VAE_ENC_SMEM_FLOATS; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < L; i += blockDim.x) xs[1 + i] = x[(size_t)b * L + i];
    __syncthreads();
    conv_rows<4, 2, NP>(xs, 0, 1, w.conv1_w, w.conv1_b, VE, L2, c1, L2 + 2, 1, true, nullptr, 0, 0);
    __syncthreads();
    conv_rows<4, 2, NP>(c1, L2 + 2, VE, w.conv2_w, w.conv2_b, VH, L4, xb, LD, 1, true, nullptr, 0, 0);
    __syncthreads();
    conv_rows<3, 1, NP>(xb, LD, VH, w.conv3_w, w.conv3_b, VH, L4, xa, LD, 1, false, nullptr, 0, 0);
    __syncthreads();
    residual_stack<NP>(xa, hid, LD, L4, w.res_w3, w.res_w1);
    // pre_vq 1x1 conv 128 -> 64 into xb rows 0..63 (index p, no halo)
    conv_rows<1, 1, NP>(xa + 1, LD, VH, w.pre_w, w.pre_b, VE, L4, xb, LD, 0, false, nullptr, 0, 0);
    __syncthreads();
    if (before)
        for (int i = tid; i < VE * L4; i += blockDim.x) before[(size_t)b * VE * L4 + i] = xb[(i / L4) * LD + (i % L4)];
    interp_rows(xb, LD, L4, z + (size_t)b * LAT, LATP, 0, LATP, VE);
}

}  // namespace t2s
