// Shared helpers for the avr_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/avr_b200.h"

namespace avr {

int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* where);
void count_launch(int n = 1);

// Makes `device` current for the duration of one ABI call (autograd threads / DataParallel
// threads may arrive with another device current).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && device >= 0 && prev != device) {
            err = cudaSetDevice(device);
            switched = (err == cudaSuccess);
        }
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};

#define AVR_ENTER(device)                                         \
    avr::DeviceGuard _guard(device);                              \
    if (_guard.err != cudaSuccess) return avr::cuda_fail(_guard.err, __func__)

#define AVR_LAUNCH_CHECK()                                        \
    do {                                                          \
        avr::count_launch();                                      \
        cudaError_t _e = cudaGetLastError();                      \
        if (_e != cudaSuccess) return avr::cuda_fail(_e, __func__); \
    } while (0)

#define AVR_CUDA(call)                                            \
    do {                                                          \
        cudaError_t _e = (call);                                  \
        if (_e != cudaSuccess) return avr::cuda_fail(_e, __func__); \
    } while (0)

#define AVR_REQUIRE(cond, msg)                                    \
    do {                                                          \
        if (!(cond)) return avr::fail(AVR_ERR_INVALID, "%s: %s", __func__, msg); \
    } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// Bit-exact sample geometry.  Every operation is an explicitly rounded fp32
// intrinsic so that nvcc can neither contract a*b+c into an FMA nor replace the true division
// by a reciprocal multiply: the results equal renderer_cpu.py's torch-CPU arithmetic bit for bit.
// ---------------------------------------------------------------------------------------------
struct Geom {
    int bs, R, S, T;
    float lo, span, fs, speed;
};

__host__ inline Geom make_geom(const avr_render_geom* g) {
    Geom r;
    r.bs = g->bs; r.R = g->R; r.S = g->S; r.T = g->T;
    r.lo = g->xyz_min; r.span = g->xyz_span; r.fs = g->fs; r.speed = g->speed;
    return r;
}

// normalize_points (renderer_cpu.py:105-106): 2*(p - lo)/(hi - lo) - 1
__device__ __forceinline__ float to_unit(float p, float lo, float span) {
    return __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fsub_rn(p, lo)), span), 1.0f);
}
// ray point (renderer_cpu.py:47): o + dir*d, product rounded before the sum
__device__ __forceinline__ float ray_point(float o, float dir, float d) {
    return __fadd_rn(o, __fmul_rn(dir, d));
}
// model.py:187: (x + 1)/2
__device__ __forceinline__ float to_cube(float n) { return __fmul_rn(__fadd_rn(n, 1.0f), 0.5f); }

// tx -> point delay in samples (renderer_cpu.py:76-77,108-109), offset quirk included.
__device__ __forceinline__ int source_delay(float ntx_x, float ntx_y, float ntx_z, float nx, float ny, float nz,
                                            const Geom& g) {
    float qx = __fadd_rn(__fmul_rn(__fmul_rn(__fadd_rn(__fsub_rn(ntx_x, nx), 1.0f), 0.5f), g.span), g.lo);
    float qy = __fadd_rn(__fmul_rn(__fmul_rn(__fadd_rn(__fsub_rn(ntx_y, ny), 1.0f), 0.5f), g.span), g.lo);
    float qz = __fadd_rn(__fmul_rn(__fmul_rn(__fadd_rn(__fsub_rn(ntx_z, nz), 1.0f), 0.5f), g.span), g.lo);
    float acc = __fmul_rn(qx, qx);
    acc = __fmaf_rn(qy, qy, acc);
    acc = __fmaf_rn(qz, qz, acc);
    float dist = __fsqrt_rn(acc);
    float v = rintf(__fdiv_rn(__fmul_rn(dist, g.fs), g.speed));      // round half to even
    v = fminf(fmaxf(v, 0.0f), (float)(g.T - 1));
    return (int)v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace avr
