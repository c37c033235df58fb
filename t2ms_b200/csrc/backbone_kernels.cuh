// Step-wise diffusion-process maths of the reference's backbone classes as device kernels (the fused sampler applies the
// same updates inside the last DiT kernel; these serve an unmodified infer.py-style loop that calls the classes per step).
// Every operation is rounded separately, in the order of the reference's torch expressions, so results match them bit for bit.
#pragma once
#include <cuda_runtime.h>

namespace t2s {

// RectifiedFlow.euler (model/backbone/rectified_flow.py:5-7): x_t + v * dt
__global__ void rf_euler_kernel(const float* __restrict__ x, const float* __restrict__ v, float dt, float* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = __fadd_rn(x[i], __fmul_rn(v[i], dt));
}

// DDPM.p_sample (model/backbone/DDPM.py:28-36): mean = 1/sqrt(alpha_t) * (x_t - eps_coef_t * eps_hat); mean + sqrt(beta_t) * noise.
// c1 / c2 / c3 [batch] hold 1/sqrt(alpha_t), (1 - alpha_t)/sqrt(1 - alpha_bar_t), sqrt(beta_t) of each sample's timestep.
__global__ void ddpm_p_sample_kernel(const float* __restrict__ xt, const float* __restrict__ eps_hat, const float* __restrict__ noise,
                                     const float* __restrict__ c1, const float* __restrict__ c2, const float* __restrict__ c3,
                                     float* __restrict__ out, size_t n, int lat) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / lat;
        const float mean = __fmul_rn(c1[b], __fsub_rn(xt[i], __fmul_rn(c2[b], eps_hat[i])));
        out[i] = __fadd_rn(mean, __fmul_rn(c3[b], noise[i]));
    }
}

}  // namespace t2s
