// On-device evaluation of generated series (the step after the hot path): evaluation.py:166-206 calculate_mse /
// calculate_wape on the (N, L, 1) arrays infer.py:117-121 saves.  HBM-bound: 8 bytes read per (sample, time step).
#pragma once
#include "common.cuh"

namespace t2s {

// one warp per sample: per_sample[i] = { sum_t (ori - gen)^2, sum_t |ori - gen|, sum_t |ori| }
__global__ void __launch_bounds__(256) series_sums_kernel(const float* __restrict__ ori, const float* __restrict__ gen, int n, int length,
                                                          float* __restrict__ per_sample) {
    const int lane = threadIdx.x & 31;
    const long long i = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    const float* a = ori + i * length;
    const float* b = gen + i * length;
    float sq = 0.f, ab = 0.f, den = 0.f;
    for (int t = lane; t < length; t += 32) {
        const float x = a[t], d = x - b[t];
        sq = fmaf(d, d, sq); ab += fabsf(d); den += fabsf(x);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
        ab += __shfl_xor_sync(0xffffffffu, ab, o);
        den += __shfl_xor_sync(0xffffffffu, den, o);
    }
    if (lane == 0) { per_sample[i * 3] = sq; per_sample[i * 3 + 1] = ab; per_sample[i * 3 + 2] = den; }
}

// out[0] = mean_i (sq_i / L)   out[1] = nanmean_i (ab_i / den_i)   out[2] = #samples with den_i != 0     (one CTA, fp64 sums)
__global__ void __launch_bounds__(1024) series_metrics_finish_kernel(const float* __restrict__ per_sample, int n, int length, double* __restrict__ out) {
    __shared__ double sm[3][32];
    double mse = 0.0, wape = 0.0, cnt = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) {
        mse += (double)per_sample[i * 3] / length;
        const float den = per_sample[i * 3 + 2];
        if (den != 0.f) { wape += (double)per_sample[i * 3 + 1] / (double)den; cnt += 1.0; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mse += __shfl_xor_sync(0xffffffffu, mse, o);
        wape += __shfl_xor_sync(0xffffffffu, wape, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sm[0][warp] = mse; sm[1][warp] = wape; sm[2][warp] = cnt; }
    __syncthreads();
    if (warp == 0) {
        mse = sm[0][lane]; wape = sm[1][lane]; cnt = sm[2][lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mse += __shfl_xor_sync(0xffffffffu, mse, o);
            wape += __shfl_xor_sync(0xffffffffu, wape, o);
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        }
        if (lane == 0) {
            out[0] = mse / n;
            out[1] = cnt > 0.0 ? wape / cnt : __longlong_as_double(0x7ff8000000000000LL);   // np.nanmean of an all-NaN list
            out[2] = cnt;
        }
    }
}

}  // namespace t2s
