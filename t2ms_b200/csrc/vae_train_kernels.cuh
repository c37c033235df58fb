// Generic LA-VAE layer kernels for the stage before the hot path: forward with saved activations and the
// exact backward of model/pretrained/vqvae.py:36-135 (univariate T2S LA-VAE) and of the fork's multivariate
// model/pretrained/myvqvae.py:24-136 (in_channels = input_dim, arbitrary length, flow_dim latent positions).
//
// Everything is fp32 in the reference layouts: activations [B][C][T], Conv1d weights [Cout][Cin][k], ConvTranspose1d
// weights [Cin][Cout][k].  The network is tiny (0.67 M parameters, <= 75 positions per sample at the inner resolution),
// so the layers are direct convolutions: one CTA stages the input rows of one sample in shared memory and every
// thread owns output elements (consecutive threads = consecutive positions: shared-memory reads are conflict-free,
// weight reads are warp broadcasts served by L1).  Two gather forms with generic weight strides cover all four
// data paths:
//   form A  out[b][oc][t] = sum_{ic,kk} W(oc,ic,kk) in[b][ic][t s - p + kk]            Conv1d forward, ConvTranspose1d dX
//   form B  out[b][oc][u] = sum_{ic,kk : (u + p - kk) % s == 0} W(oc,ic,kk) in[b][ic][(u + p - kk) / s]
//                                                                                      ConvTranspose1d forward, Conv1d dX
// and one kernel forms every weight gradient  dW(yc,xc,kk) = sum_{b,t} Y[b][yc][t] X[b][xc][t s - p + kk].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace t2s {

struct ConvGeom {
    int Cin, Tin, Cout, Tout, k, stride, pad;
    int w_so, w_si;        // weight element (oc, ic, kk) at oc * w_so + ic * w_si + kk
};

// grid (B, nsplit): split y handles output channels [y * Cout / nsplit, ...); dynamic smem Cin * Tin floats
template <bool FORM_B>
__global__ void __launch_bounds__(256) conv_gather_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                                          const float* __restrict__ res, float* __restrict__ out, ConvGeom g, int relu_in, int relu_out) {
    extern __shared__ float s_in[];
    const int b = blockIdx.x;
    const float* src = in + (size_t)b * g.Cin * g.Tin;
    for (int i = threadIdx.x; i < g.Cin * g.Tin; i += 256) {
        const float v = src[i];
        s_in[i] = relu_in ? fmaxf(v, 0.f) : v;
    }
    __syncthreads();
    const int per = (g.Cout + gridDim.y - 1) / gridDim.y, oc0 = blockIdx.y * per, oc1 = min(g.Cout, oc0 + per);
    const int n = (oc1 - oc0) * g.Tout;
    for (int e = threadIdx.x; e < n; e += 256) {
        const int oc = oc0 + e / g.Tout, t = e % g.Tout;
        float acc = bias != nullptr ? bias[oc] : 0.f;
        const float* wr = w + (size_t)oc * g.w_so;
        if (!FORM_B) {
            const int base = t * g.stride - g.pad;
            const int k0 = max(0, -base), k1 = min(g.k, g.Tin - base);
            for (int ic = 0; ic < g.Cin; ++ic) {
                const float* wi = wr + (size_t)ic * g.w_si;
                const float* si = s_in + ic * g.Tin + base;
                for (int kk = k0; kk < k1; ++kk) acc = fmaf(wi[kk], si[kk], acc);
            }
        } else {
            // valid taps of this output position (at most ceil(k / stride)): kk = (t + p) mod s, + s, ...
            for (int kk = (t + g.pad) % g.stride; kk < g.k; kk += g.stride) {
                const int i = (t + g.pad - kk) / g.stride;
                if (t + g.pad - kk < 0 || i >= g.Tin) continue;
                for (int ic = 0; ic < g.Cin; ++ic) acc = fmaf(wr[(size_t)ic * g.w_si + kk], s_in[ic * g.Tin + i], acc);
            }
        }
        const size_t o = ((size_t)b * g.Cout + oc) * g.Tout + t;
        if (res != nullptr) acc += res[o];
        out[o] = relu_out ? fmaxf(acc, 0.f) : acc;
    }
}

// dW(yc, xc, kk) += sum_{b in chunk} sum_t Y[b][yc][t] * X[b][xc][t * stride - pad + kk]     (atomic accumulation)
//   Y: [B][Cy][Ty]   X: [B][Cx][Tx]   element stored at dw[yc * w_sy + xc * w_sx + kk]
// grid (ceil(B / bchunk), ceil(Cy / 8)); dynamic smem: (8 * Ty + Cx * (Tx + 1)) floats
struct WgradGeom { int Cy, Ty, Cx, Tx, k, stride, pad, w_sy, w_sx, bchunk, B; };
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const float* __restrict__ Y, const float* __restrict__ X, float* __restrict__ dw, WgradGeom g,
                                                         int relu_x) {
    extern __shared__ float sm[];
    float* sy = sm;                          // [8][Ty]
    float* sx = sm + 8 * g.Ty;               // [Cx][Tx + 1]  (odd pitch: consecutive channels hit different banks)
    const int yc0 = blockIdx.y * 8, ny = min(8, g.Cy - yc0);
    const int nel = ny * g.Cx * g.k;         // gradient elements of this CTA
    constexpr int MAXE = 16;                 // per-thread accumulators: 8 * Cx * k / 256 <= 16 for Cx * k <= 512
    float acc[MAXE];
#pragma unroll
    for (int i = 0; i < MAXE; ++i) acc[i] = 0.f;
    const int b0 = blockIdx.x * g.bchunk, b1 = min(g.B, b0 + g.bchunk);
    const int px = g.Tx + 1;
    for (int b = b0; b < b1; ++b) {
        __syncthreads();
        for (int i = threadIdx.x; i < ny * g.Ty; i += 256) sy[i] = Y[((size_t)b * g.Cy + yc0) * g.Ty + i];
        for (int i = threadIdx.x; i < g.Cx * g.Tx; i += 256) {
            const float v = X[(size_t)b * g.Cx * g.Tx + i];
            sx[(i / g.Tx) * px + i % g.Tx] = relu_x ? fmaxf(v, 0.f) : v;
        }
        __syncthreads();
#pragma unroll
        for (int a = 0; a < MAXE; ++a) {
            const int e = threadIdx.x + a * 256;
            if (e < nel) {
                const int kk = e % g.k, xc = (e / g.k) % g.Cx, yl = e / (g.k * g.Cx);
                const float* yr = sy + yl * g.Ty;
                const float* xr = sx + xc * px;
                float s = 0.f;
                for (int t = 0; t < g.Ty; ++t) {
                    const int i = t * g.stride - g.pad + kk;
                    if (i >= 0 && i < g.Tx) s = fmaf(yr[t], xr[i], s);
                }
                acc[a] += s;
            }
        }
    }
#pragma unroll
    for (int a = 0; a < MAXE; ++a) {
        const int e = threadIdx.x + a * 256;
        if (e < nel) {
            const int kk = e % g.k, xc = (e / g.k) % g.Cx, yl = e / (g.k * g.Cx);
            atomicAdd(dw + (size_t)(yc0 + yl) * g.w_sy + (size_t)xc * g.w_sx + kk, acc[a]);
        }
    }
}

// db[c] += sum_{b,t} d[b][c][t]      grid (C), block 256
__global__ void __launch_bounds__(256) chan_sum_kernel(const float* __restrict__ d, float* __restrict__ db, int B, int C, int T) {
    __shared__ float red[8];
    const int c = blockIdx.x;
    float s = 0.f;
    for (int i = threadIdx.x; i < B * T; i += 256) s += d[((size_t)(i / T) * C + c) * T + i % T];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        atomicAdd(db + c, t);
    }
}

// F.interpolate(mode='linear', align_corners=True) along the last axis: rows = B * C
__device__ __forceinline__ void interp_coef(int j, int Tin, int Tout, int& i0, int& i1, float& l1) {
    const float scale = Tout > 1 ? (float)(Tin - 1) / (float)(Tout - 1) : 0.f;
    const float src = scale * j;
    i0 = (int)src;
    i1 = i0 + (i0 < Tin - 1 ? 1 : 0);
    l1 = src - i0;
}
__global__ void interp_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, size_t rows, int Tin, int Tout) {
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < rows * Tout; e += (size_t)gridDim.x * blockDim.x) {
        const size_t r = e / Tout;
        const int j = (int)(e % Tout);
        int i0, i1; float l1;
        interp_coef(j, Tin, Tout, i0, i1, l1);
        const float* s = in + r * Tin;
        out[e] = (1.f - l1) * s[i0] + l1 * s[i1];
    }
}
// din[r][i] (+)= sum_j coef(j -> i) dout[r][j]     (gather form: deterministic)
__global__ void interp_bwd_kernel(const float* __restrict__ dout, float* __restrict__ din, size_t rows, int Tin, int Tout, int accumulate) {
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < rows * Tin; e += (size_t)gridDim.x * blockDim.x) {
        const size_t r = e / Tin;
        const int i = (int)(e % Tin);
        const float* d = dout + r * Tout;
        float s = 0.f;
        for (int j = 0; j < Tout; ++j) {
            int i0, i1; float l1;
            interp_coef(j, Tin, Tout, i0, i1, l1);
            if (i0 == i) s += (1.f - l1) * d[j];
            if (i1 == i) s += l1 * d[j];
        }
        din[e] = accumulate ? din[e] + s : s;
    }
}

__global__ void relu_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n) {
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) out[e] = fmaxf(in[e], 0.f);
}
// d = (y > 0) ? d (+ add) : 0     (backward of an (in-place) ReLU whose output is y)
__global__ void relu_bwd_kernel(float* __restrict__ d, const float* __restrict__ y, const float* __restrict__ add, size_t n) {
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const float v = d[e] + (add != nullptr ? add[e] : 0.f);
        d[e] = y[e] > 0.f ? v : 0.f;
    }
}
// F.mse_loss(a, b): loss_sum += sum (a - b)^2 ; optional gradients da = scale (a - b), db = -da   (scale = 2 / numel)
__global__ void __launch_bounds__(256) mse_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n, float scale,
                                                  float* __restrict__ loss_sum, float* __restrict__ da, float* __restrict__ db) {
    __shared__ float red[8];
    float s = 0.f;
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (size_t)gridDim.x * 256) {
        const float d = a[e] - b[e];
        s = fmaf(d, d, s);
        if (da != nullptr) da[e] = scale * d;
        if (db != nullptr) db[e] = -scale * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        atomicAdd(loss_sum, t);
    }
}

}  // namespace t2s
