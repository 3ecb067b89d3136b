// Fused optimiser step over the flat parameter / gradient arena.
//
// Replaces avr_runner.py:192-200: clip_grad_norm_(max_norm) -> NaN/Inf scrub of every .grad -> Adam.step(),
// which the reference runs as a Python loop over parameters (3 passes + ~10 small kernels per tensor).
// Here: one deterministic two-stage sum of squares, then ONE pass that scales, scrubs and applies Adam:
// 28 bytes of HBM traffic per parameter (read g,p,m,v; write p,m,v).
// Order of operations follows torch: the clip coefficient is computed from the UNscrubbed gradient, so a NaN
// anywhere makes the coefficient NaN, every gradient NaN, and the scrub then zeroes the whole step -- exactly
// what the reference does.
#include <math.h>
#include "common.cuh"

namespace avr {

constexpr int SQ_BLOCKS = 1184;   // 148 SMs x 8

__global__ void __launch_bounds__(256)
sumsq_partial_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ partial) {
    __shared__ double red[8];
    double acc = 0.0;
    const int64_t n4 = n / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
        acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int64_t i = n4 * 4; i < n; ++i) acc += (double)g[i] * g[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; ++k) t += red[k];
        partial[blockIdx.x] = t;
    }
}

// norm_out[0] = total L2 norm, norm_out[1] = clip coefficient min(1, max_norm / (norm + 1e-6))
__global__ void sumsq_final_kernel(const double* __restrict__ partial, int n_blocks, float max_norm, float* __restrict__ norm_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double t = 0.0;
    for (int k = 0; k < n_blocks; ++k) t += partial[k];       // fixed order
    const float norm = (float)sqrt(t);
    float coef = 1.0f;
    if (max_norm > 0.f) {
        coef = max_norm / (norm + 1e-6f);
        coef = coef > 1.0f ? 1.0f : coef;                      // NaN stays NaN (torch.clamp(max=1))
    }
    norm_out[0] = norm;
    norm_out[1] = coef;
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
            const float* __restrict__ norm, float lr, float beta1, float beta2, float eps, float weight_decay,
            float bias_corr1, float bias_corr2_sqrt, int write_back_grad) {
    const float coef = __ldg(norm + 1);
    const float step_size = lr / bias_corr1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float gi = g[i] * coef;
        if (gi != gi || isinf(gi)) gi = 0.f;                   // avr_runner.py:193-197
        if (write_back_grad) g[i] = gi;
        const float pi = p[i];
        if (weight_decay != 0.f) gi = fmaf(weight_decay, pi, gi);
        const float mi = m[i] + (gi - m[i]) * (1.0f - beta1);  // exp_avg.lerp_(grad, 1 - beta1)
        const float vi = fmaf(gi * gi, 1.0f - beta2, v[i] * beta2);
        const float denom = sqrtf(vi) / bias_corr2_sqrt + eps;
        m[i] = mi;
        v[i] = vi;
        p[i] = pi - step_size * (mi / denom);
    }
}

}  // namespace avr

using namespace avr;

extern "C" {

AVR_API int64_t avr_adam_workspace_bytes(void) { return SQ_BLOCKS * (int64_t)sizeof(double) + 64; }

// One training update on flat fp32 arenas of n elements.  norm_out: 2 floats (total gradient norm, clip coefficient).
// max_norm <= 0 disables clipping.  step: 1-based Adam step count.
AVR_API int avr_fused_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                                float beta1, float beta2, float eps, float weight_decay, float max_norm, int64_t step,
                                int write_back_grad, float* norm_out, void* workspace, int64_t workspace_bytes, int device,
                                void* stream) {
    AVR_REQUIRE(params && grads && exp_avg && exp_avg_sq && norm_out && workspace, "null pointer");
    AVR_REQUIRE(step >= 1 && n >= 0, "bad step / size");
    AVR_REQUIRE(workspace_bytes >= avr_adam_workspace_bytes() && aligned16(grads) && aligned16(workspace), "workspace / alignment");
    AVR_ENTER(device);
    if (n == 0) return AVR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    double* partial = (double*)workspace;
    sumsq_partial_kernel<<<SQ_BLOCKS, 256, 0, st>>>(grads, n, partial);
    AVR_LAUNCH_CHECK();
    sumsq_final_kernel<<<1, 32, 0, st>>>(partial, SQ_BLOCKS, max_norm, norm_out);
    AVR_LAUNCH_CHECK();
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    adam_kernel<<<SQ_BLOCKS, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, norm_out, lr, beta1, beta2, eps, weight_decay,
                                           (float)bc1, (float)sqrt(bc2), write_back_grad);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

}  // extern "C"
