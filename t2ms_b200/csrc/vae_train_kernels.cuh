// Generic LA-VAE layer kernels for the stage before the hot path: forward with saved activations and the
// exact backward of model/pretrained/vqvae.py:36-135 (univariate T2S LA-VAE) and of the fork's multivariate
// model/pretrained/myvqvae.py:24-136 (in_channels = input_dim, arbitrary length, flow_dim latent positions).
//
// Everything is fp32 in the reference layouts: activations [B][C][T], Conv1d weights [Cout][Cin][k], ConvTranspose1d
// weights [Cin][Cout][k].  The network is tiny (0.67 M parameters, <= 75 positions per sample at the inner resolution),
// so the layers are direct convolutions with shared-memory staging and register tiling (see conv_gather_kernel).  Two
// gather forms with generic weight strides cover all four data paths:
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
    int B, bchunk;         // batch; samples per CTA
};
constexpr int CG_OT = 32;                  // output channels per CTA
constexpr int CG_MAXU = 4;                 // work units (4 output channels x 1 position) per thread: 8 * Tout <= 1024

// grid (ceil(B / bchunk), ceil(Cout / 32)), block 256; dynamic smem: (Cin * k * 32 + Cin * Tin) floats.
// The CTA keeps the weights of its 32 output channels in shared memory (transposed: [ic * k + kk][32], so a thread reads
// the 4 output channels it owns with one 16-byte load) and walks over its samples, staging each sample's input rows.
// A thread owns 4 output channels x one position: per (input channel, tap) one weight vector load + one input load feed
// 4 FMAs.  The taps that are valid for a thread's position (bounds, stride phase) are resolved once, outside the loops.
template <bool FORM_B>
__global__ void __launch_bounds__(256) conv_gather_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                                          const float* __restrict__ res, float* __restrict__ out, ConvGeom g, int relu_in, int relu_out) {
    extern __shared__ __align__(16) float cg_sm[];
    float* s_w = cg_sm;                                  // [Cin * k][32]
    float* s_in = cg_sm + g.Cin * g.k * CG_OT;           // [Cin][Tin]
    const int oc0 = blockIdx.y * CG_OT, noc = min(CG_OT, g.Cout - oc0);
    const int nj = g.Cin * g.k;
    for (int i = threadIdx.x; i < nj * CG_OT; i += 256) {
        const int ocl = i % CG_OT, j = i / CG_OT, ic = j / g.k, kk = j % g.k;
        s_w[i] = ocl < noc ? w[(size_t)(oc0 + ocl) * g.w_so + (size_t)ic * g.w_si + kk] : 0.f;
    }
    // work units of this thread: u = tid + 256 a  ->  (output-channel quad u / Tout, position u % Tout)
    const int nunit = (CG_OT / 4) * g.Tout;
    // tap slot x of a unit: kernel tap u_tk and input position u_ti (-1: tap not valid for this output position)
    int u_pos[CG_MAXU], u_q[CG_MAXU], u_tk[CG_MAXU][4], u_ti[CG_MAXU][4];
#pragma unroll
    for (int a = 0; a < CG_MAXU; ++a) {
        const int u = threadIdx.x + a * 256;
        const bool live = u < nunit;
        const int t = live ? u % g.Tout : 0;
        u_pos[a] = t; u_q[a] = live ? u / g.Tout : 0;
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            const int kk = FORM_B ? (t + g.pad) % g.stride + x * g.stride : x;
            int idx = -1;
            if (live && kk < g.k) {
                if (!FORM_B) {
                    const int i = t * g.stride - g.pad + kk;
                    if (i >= 0 && i < g.Tin) idx = i;
                } else {
                    const int num = t + g.pad - kk;
                    if (num >= 0 && num / g.stride < g.Tin) idx = num / g.stride;
                }
            }
            u_tk[a][x] = min(kk, g.k - 1);
            u_ti[a][x] = idx;
        }
    }
    const int b0 = blockIdx.x * g.bchunk, b1 = min(g.B, b0 + g.bchunk);
    for (int b = b0; b < b1; ++b) {
        __syncthreads();                                  // previous sample consumed (first pass: weights staged)
        const float* src = in + (size_t)b * g.Cin * g.Tin;
        for (int i = threadIdx.x; i < g.Cin * g.Tin; i += 256) {
            const float v = src[i];
            s_in[i] = relu_in ? fmaxf(v, 0.f) : v;
        }
        __syncthreads();
#pragma unroll
        for (int a = 0; a < CG_MAXU; ++a) {
            if (threadIdx.x + a * 256 >= nunit) continue;
            const int q = u_q[a], t = u_pos[a];
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int ic = 0; ic < g.Cin; ++ic) {
                const float* wi = s_w + (size_t)(ic * g.k) * CG_OT + q * 4;
                const float* si = s_in + ic * g.Tin;
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    if (u_ti[a][x] >= 0) {
                        const float4 w4 = *reinterpret_cast<const float4*>(wi + u_tk[a][x] * CG_OT);
                        const float v = si[u_ti[a][x]];
                        acc[0] = fmaf(w4.x, v, acc[0]); acc[1] = fmaf(w4.y, v, acc[1]);
                        acc[2] = fmaf(w4.z, v, acc[2]); acc[3] = fmaf(w4.w, v, acc[3]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int oc = oc0 + q * 4 + j;
                if (oc < g.Cout) {
                    float r = acc[j] + (bias != nullptr ? bias[oc] : 0.f);
                    const size_t o = ((size_t)b * g.Cout + oc) * g.Tout + t;
                    if (res != nullptr) r += res[o];
                    out[o] = relu_out ? fmaxf(r, 0.f) : r;
                }
            }
        }
    }
}

// dW(yc, xc, kk) += sum_{b in chunk} sum_t Y[b][yc][t] * X[b][xc][t * stride - pad + kk]     (atomic accumulation)
//   Y: [B][Cy][Ty]   X: [B][Cx][Tx]   element stored at dw[yc * w_sy + xc * w_sx + kk]
// grid (ceil(B / bchunk), ceil(Cy / 8)); dynamic smem: (8 * Ty + Cx * (Tx + 1)) floats.
// A thread owns one (xc, kk) and all 8 output channels of the CTA: per position one X load and two 16-byte Y loads
// ([t][8] layout) feed 8 FMAs; the batch chunk is accumulated in registers, then one atomic per element.
struct WgradGeom { int Cy, Ty, Cx, Tx, k, stride, pad, w_sy, w_sx, bchunk, B; };
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const float* __restrict__ Y, const float* __restrict__ X, float* __restrict__ dw, WgradGeom g,
                                                         int relu_x) {
    extern __shared__ __align__(16) float sm[];
    float* sy = sm;                          // [Ty][8]
    float* sx = sm + 8 * g.Ty;               // [Cx][Tx + 1]  (odd pitch: consecutive channels hit different banks)
    const int yc0 = blockIdx.y * 8, ny = min(8, g.Cy - yc0);
    const int nunit = g.Cx * g.k;            // (xc, kk) pairs: <= 512 -> at most two per thread
    float acc[2][8];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[a][j] = 0.f;
    const int b0 = blockIdx.x * g.bchunk, b1 = min(g.B, b0 + g.bchunk);
    const int px = g.Tx + 1;
    for (int b = b0; b < b1; ++b) {
        __syncthreads();
        for (int i = threadIdx.x; i < 8 * g.Ty; i += 256) {
            const int yl = i / g.Ty, t = i % g.Ty;
            sy[t * 8 + yl] = yl < ny ? Y[((size_t)b * g.Cy + yc0 + yl) * g.Ty + t] : 0.f;
        }
        for (int i = threadIdx.x; i < g.Cx * g.Tx; i += 256) {
            const float v = X[(size_t)b * g.Cx * g.Tx + i];
            sx[(i / g.Tx) * px + i % g.Tx] = relu_x ? fmaxf(v, 0.f) : v;
        }
        __syncthreads();
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int e = threadIdx.x + a * 256;
            if (e < nunit) {
                const int kk = e % g.k, xc = e / g.k;
                const float* xr = sx + xc * px;
                const int off = kk - g.pad;
                // positions t with 0 <= t * stride + off < Tx
                const int t_lo = off >= 0 ? 0 : (-off + g.stride - 1) / g.stride;
                const int t_hi = g.Tx - 1 - off < 0 ? 0 : min(g.Ty, (g.Tx - 1 - off) / g.stride + 1);
                for (int t = t_lo; t < t_hi; ++t) {
                    const float xv = xr[t * g.stride + off];
                    const float4 y0 = *reinterpret_cast<const float4*>(sy + t * 8), y1 = *reinterpret_cast<const float4*>(sy + t * 8 + 4);
                    acc[a][0] = fmaf(y0.x, xv, acc[a][0]); acc[a][1] = fmaf(y0.y, xv, acc[a][1]);
                    acc[a][2] = fmaf(y0.z, xv, acc[a][2]); acc[a][3] = fmaf(y0.w, xv, acc[a][3]);
                    acc[a][4] = fmaf(y1.x, xv, acc[a][4]); acc[a][5] = fmaf(y1.y, xv, acc[a][5]);
                    acc[a][6] = fmaf(y1.z, xv, acc[a][6]); acc[a][7] = fmaf(y1.w, xv, acc[a][7]);
                }
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const int e = threadIdx.x + a * 256;
        if (e < nunit) {
            const int kk = e % g.k, xc = e / g.k;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < ny) atomicAdd(dw + (size_t)(yc0 + j) * g.w_sy + (size_t)xc * g.w_sx + kk, acc[a][j]);
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
