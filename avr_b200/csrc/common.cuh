// Shared helpers for the avr_b200 kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
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

// ---------------------------------------------------------------------------------------------
// Plane sets (tensor-core operands, include/avr_b200.h "plane-set kinds"): 16-bit planes [n][rows][ld].
//   AVR_PLANES_BF16x2 / x3: x = hi + mid (+ lo), all bf16                      (16 / 24 mantissa bits, fp32 range)
//   AVR_PLANES_F16x2:       x = hi + lo' * 2^-11, hi = fp16(x), lo' = fp16((x - hi) * 2^11)   (24 bits, fp16 range)
// ---------------------------------------------------------------------------------------------
constexpr float F16_LO_SCALE = 2048.0f, F16_LO_INV = 1.0f / 2048.0f;
__host__ __device__ inline int planes_count(int kind) { return kind & 15; }
__host__ __device__ inline bool planes_f16(int kind) { return (kind >> 4) == 1; }
__host__ inline bool planes_kind_ok(int kind) { return kind == AVR_PLANES_BF16x2 || kind == AVR_PLANES_BF16x3 || kind == AVR_PLANES_F16x2; }

__device__ __forceinline__ void planes_store(void* base, int64_t idx, int64_t plane, int kind, float v) {
    if (planes_f16(kind)) {
        __half* q = reinterpret_cast<__half*>(base) + idx;
        const __half hi = __float2half_rn(v);
        q[0] = hi;
        q[plane] = __float2half_rn((v - __half2float(hi)) * F16_LO_SCALE);
    } else {
        __nv_bfloat16* q = reinterpret_cast<__nv_bfloat16*>(base) + idx;
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const float r1 = v - __bfloat162float(hi);
        const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
        q[0] = hi;
        q[plane] = mid;
        if (planes_count(kind) == 3) q[2 * plane] = __float2bfloat16_rn(r1 - __bfloat162float(mid));
    }
}
__device__ __forceinline__ float planes_load(const void* base, int64_t idx, int64_t plane, int kind) {
    if (planes_f16(kind)) {
        const __half* q = reinterpret_cast<const __half*>(base) + idx;
        return fmaf(__half2float(q[plane]), F16_LO_INV, __half2float(q[0]));
    }
    const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(base) + idx;
    float tail = __bfloat162float(q[plane]);
    if (planes_count(kind) == 3) tail += __bfloat162float(q[2 * plane]);
    return __bfloat162float(q[0]) + tail;
}
// two fp32 -> packed 16-bit pair (round to nearest even); low half = first argument
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t cvt_f16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float2 f16x2_to_float2(uint32_t w) {
    return __half22float2(*reinterpret_cast<const __half2*>(&w));
}
// packed pair of a value pair (a, b) in plane-set `kind`: planes 0, 1 (, 2)
__device__ __forceinline__ void planes_pack2(int kind, float a, float b, uint32_t& p0, uint32_t& p1, uint32_t& p2) {
    if (planes_f16(kind)) {
        p0 = cvt_f16x2(a, b);
        const float2 h = f16x2_to_float2(p0);
        p1 = cvt_f16x2((a - h.x) * F16_LO_SCALE, (b - h.y) * F16_LO_SCALE);
        p2 = 0u;
    } else {
        p0 = cvt_bf16x2(a, b);
        const float ar = a - __uint_as_float(p0 << 16), br = b - __uint_as_float(p0 & 0xFFFF0000u);
        p1 = cvt_bf16x2(ar, br);
        p2 = planes_count(kind) == 3 ? cvt_bf16x2(ar - __uint_as_float(p1 << 16), br - __uint_as_float(p1 & 0xFFFF0000u)) : 0u;
    }
}
// value pair of packed planes 0, 1 (bf16x3's third plane is the caller's business)
__device__ __forceinline__ float2 planes_unpack2(bool f16, uint32_t p0, uint32_t p1) {
    if (f16) {
        const float2 h = f16x2_to_float2(p0), l = f16x2_to_float2(p1);
        return make_float2(fmaf(l.x, F16_LO_INV, h.x), fmaf(l.y, F16_LO_INV, h.y));
    }
    return make_float2(__uint_as_float(p0 << 16) + __uint_as_float(p1 << 16),
                       __uint_as_float(p0 & 0xFFFF0000u) + __uint_as_float(p1 & 0xFFFF0000u));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace avr
