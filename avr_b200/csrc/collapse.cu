// Output layer of the signal network fused with the time-domain ray reduction ("collapse").
//
// Reference: signal = H @ W_out^T (model.py:231, the 512 -> T layer, 215 GFLOP / receiver at simu), then
// renderer.py:86-90,115-118 masks every row with (t >= delay) and sums the rays with the compositing weights:
//
//      y[b,s,t] = sum_r w[b,r,s] * [t >= delay[b,r,s]] * ( H[b,r,s,:] . W_out[t,:] )
//
// Sorting the rays of one (b,s) by delay makes the masked sum a PREFIX sum:
//
//      G[b,s,t,:] = sum_{r : delay <= t} w * H[b,r,s,:]          y[b,s,t] = G[b,s,t,:] . W_out[t,:]
//
// i.e. S*T*width MACs per receiver instead of R*S*T*width (R = 2050 x fewer), and the [bs,R,S,T] signal
// tensor (0.84 GB / receiver) and its gradient never exist.  Exact in real arithmetic; in fp32 only the
// summation order differs from the reference.  The backward uses the mirrored SUFFIX sum
//      g[b,s,d,:] = sum_{t >= d} d_y[b,s,t] * W_out[t,:]         d_H = w * g[delay],   d_w = H . g[delay]
// and  d_W_out[t,:] = sum_{b,s} d_y[b,s,t] * G[b,s,t,:].
//
// One CTA per (b,s); each thread owns 4 hidden columns; the delay-sorted ray list lives in shared memory;
// activation rows (bf16 hi/lo plane pairs) are fetched 8 rays at a time to keep loads in flight.
#include <cuda_bf16.h>
#include "common.cuh"

namespace avr {

constexpr int COL_PER_THREAD = 4;
constexpr int RAY_BATCH = 8;

struct Planes {
    const __nv_bfloat16* p;
    long long ld, plane;
};

__device__ __forceinline__ void ld_row4(const Planes& a, long long row, int c, uint2& hi, uint2& lo) {
    const __nv_bfloat16* q = a.p + row * a.ld + c;
    hi = __ldg(reinterpret_cast<const uint2*>(q));
    lo = __ldg(reinterpret_cast<const uint2*>(q + a.plane));
}
__device__ __forceinline__ void unpack4(uint2 hi, uint2 lo, float (&x)[4]) {
    x[0] = __uint_as_float(hi.x << 16) + __uint_as_float(lo.x << 16);
    x[1] = __uint_as_float(hi.x & 0xFFFF0000u) + __uint_as_float(lo.x & 0xFFFF0000u);
    x[2] = __uint_as_float(hi.y << 16) + __uint_as_float(lo.y << 16);
    x[3] = __uint_as_float(hi.y & 0xFFFF0000u) + __uint_as_float(lo.y & 0xFFFF0000u);
}
__device__ __forceinline__ uint32_t pack_bf2(float a, float b, uint32_t& lo_out) {
    const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
    const __nv_bfloat16 al = __float2bfloat16_rn(a - __bfloat162float(ah)), bl = __float2bfloat16_rn(b - __bfloat162float(bh));
    lo_out = (uint32_t)__bfloat16_as_ushort(al) | ((uint32_t)__bfloat16_as_ushort(bl) << 16);
    return (uint32_t)__bfloat16_as_ushort(ah) | ((uint32_t)__bfloat16_as_ushort(bh) << 16);
}

// ------------------------------------------------------------------------------------------------
// stable counting sort of the R rays of every (b,s) by delay: one warp per (b,s)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
delay_sort_kernel(const Geom geo, const int* __restrict__ delay, const float* __restrict__ w, int* __restrict__ order,
                  int* __restrict__ sdelay, float* __restrict__ sw) {
    extern __shared__ int hist[];
    const int lane = threadIdx.x;
    const int bsi = blockIdx.x, b = bsi / geo.S, s = bsi - b * geo.S;
    const int R = geo.R, T = geo.T;
    for (int t = lane; t < T; t += 32) hist[t] = 0;
    __syncwarp();
    for (int r0 = 0; r0 < R; r0 += 32) {
        const int r = r0 + lane;
        const int d = r < R ? __ldg(delay + ((long long)b * R + r) * geo.S + s) : -1;
        const unsigned m = __match_any_sync(0xffffffffu, d);
        if (d >= 0 && lane == __ffs(m) - 1) hist[d] += __popc(m);
        __syncwarp();
    }
    int carry = 0;
    for (int t0 = 0; t0 < T; t0 += 32) {
        const int t = t0 + lane;
        const int v = t < T ? hist[t] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (t < T) hist[t] = carry + incl - v;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
    const long long out0 = (long long)bsi * R;
    for (int r0 = 0; r0 < R; r0 += 32) {
        const int r = r0 + lane;
        const long long n = ((long long)b * R + r) * geo.S + s;
        const int d = r < R ? __ldg(delay + n) : -1;
        const unsigned m = __match_any_sync(0xffffffffu, d);
        const int rank = __popc(m & ((1u << lane) - 1u));
        const int base = d >= 0 ? hist[d] : 0;
        __syncwarp();
        if (d >= 0) {
            const int pos = base + rank;                       // stable: rays of one bucket keep ray order
            order[out0 + pos] = r;
            sdelay[out0 + pos] = d;
            sw[out0 + pos] = __ldg(w + n);
            if (lane == __ffs(m) - 1) hist[d] += __popc(m);
        }
        __syncwarp();
    }
}

// shared-memory staging of one (b,s)'s sorted ray list
struct RayList {
    int* ord;
    int* del;
    float* wgt;
};
__device__ __forceinline__ RayList stage_rays(unsigned char* smem, int R, long long base, const int* order, const int* sdelay,
                                              const float* sw) {
    RayList L;
    L.ord = reinterpret_cast<int*>(smem);
    L.del = L.ord + R;
    L.wgt = reinterpret_cast<float*>(L.del + R);
    for (int k = threadIdx.x; k < R; k += blockDim.x) {
        L.ord[k] = __ldg(order + base + k);
        L.del[k] = __ldg(sdelay + base + k);
        L.wgt[k] = __ldg(sw + base + k);
    }
    return L;
}

// ------------------------------------------------------------------------------------------------
// forward: y[b,s,t] = G[b,s,t,:] . W_out[t,:]
// MODE 0: write y.   MODE 1 (weight gradient, first pass): write P[t - dmin, :] = d_y[t] * G[t,:] for
// t in [dmin, dmax), G_tot and the (dmin, dmax) range instead.
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256)
collapse_fwd_kernel(const Geom geo, const Planes act, int width, const int* __restrict__ order,
                    const int* __restrict__ sdelay, const float* __restrict__ sw, const float* __restrict__ w_out,
                    long long ldw, float* __restrict__ y, const float* __restrict__ d_y, float* __restrict__ P,
                    float* __restrict__ gtot, int* __restrict__ range, int tspan) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int bsi = blockIdx.x, b = bsi / geo.S, s = bsi - b * geo.S;
    const int R = geo.R, T = geo.T;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int c = threadIdx.x * COL_PER_THREAD;
    const bool col_ok = c < width;
    RayList L = stage_rays(smem, R, (long long)bsi * R, order, sdelay, sw);
    float* partial = reinterpret_cast<float*>(smem + (size_t)R * 12);       // MODE 0: [n_warps][T]
    float* dy_s = partial;                                                   // MODE 1: [T]
    if (MODE == 0) {
        for (int i = threadIdx.x; i < n_warps * T; i += blockDim.x) partial[i] = 0.f;
    } else {
        for (int i = threadIdx.x; i < T; i += blockDim.x) dy_s[i] = __ldg(d_y + (long long)bsi * T + i);
    }
    __syncthreads();
    const int dmin = L.del[0], dmax = L.del[R - 1];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int t = dmin;
    bool overflow = false;
    if (MODE == 1) overflow = (dmax - dmin) > tspan;

    auto emit = [&](int tt) {
        if (MODE == 0) {
            float p = 0.f;
            if (col_ok) {
                const float4 wv = __ldg(reinterpret_cast<const float4*>(w_out + (long long)tt * ldw + c));
                p = acc[0] * wv.x + acc[1] * wv.y + acc[2] * wv.z + acc[3] * wv.w;
            }
            p = warp_sum(p);
            if (lane == 0) partial[warp * T + tt] = p;
        } else if (col_ok && !overflow) {
            const float g = dy_s[tt];
            *reinterpret_cast<float4*>(P + ((long long)bsi * tspan + (tt - dmin)) * width + c) =
                make_float4(g * acc[0], g * acc[1], g * acc[2], g * acc[3]);
        }
    };

    for (int k0 = 0; k0 < R; k0 += RAY_BATCH) {
        uint2 hi[RAY_BATCH], lo[RAY_BATCH];
#pragma unroll
        for (int j = 0; j < RAY_BATCH; ++j) {
            hi[j] = make_uint2(0u, 0u); lo[j] = make_uint2(0u, 0u);
            if (k0 + j < R && col_ok) ld_row4(act, ((long long)b * R + L.ord[k0 + j]) * geo.S + s, c, hi[j], lo[j]);
        }
#pragma unroll
        for (int j = 0; j < RAY_BATCH; ++j) {
            if (k0 + j >= R) break;
            const int d = L.del[k0 + j];
            while (t < d) { emit(t); ++t; }
            float x[4];
            unpack4(hi[j], lo[j], x);
            const float wk = L.wgt[k0 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = fmaf(wk, x[i], acc[i]);
        }
    }
    if (MODE == 0) {
        for (; t < T; ++t) emit(t);                                          // t >= dmax: every ray contributes
        __syncthreads();
        for (int i = threadIdx.x; i < T; i += blockDim.x) {
            float v = 0.f;
            for (int wq = 0; wq < n_warps; ++wq) v += partial[wq * T + i];
            y[(long long)bsi * T + i] = v;
        }
    } else {
        if (col_ok) {
            const float poison = overflow ? __uint_as_float(0x7fc00000u) : 0.f;
            *reinterpret_cast<float4*>(gtot + (long long)bsi * width + c) =
                make_float4(acc[0] + poison, acc[1] + poison, acc[2] + poison, acc[3] + poison);
        }
        if (threadIdx.x == 0) { range[2 * bsi] = dmin; range[2 * bsi + 1] = dmax; }
    }
}

// d_W_out[t, c] (+)= sum_{b,s} ( t < dmin ? 0 : t < dmax ? P[b,s,t-dmin,c] : d_y[b,s,t] * G_tot[b,s,c] )
__global__ void collapse_dw_reduce_kernel(const Geom geo, int width, const float* __restrict__ d_y, const float* __restrict__ P,
                                          const float* __restrict__ gtot, const int* __restrict__ range, int tspan,
                                          float* __restrict__ d_wout, long long ldw, int accumulate) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int wq = width / 4;
    if (q >= geo.T * wq) return;
    const int t = q / wq, c = (q - t * wq) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const int n = geo.bs * geo.S;
    for (int i = 0; i < n; ++i) {                                            // fixed order: deterministic
        const int dmin = __ldg(range + 2 * i), dmax = __ldg(range + 2 * i + 1);
        if (t < dmin) continue;
        float4 a;
        if (t < dmax) {
            a = *reinterpret_cast<const float4*>(P + ((long long)i * tspan + (t - dmin)) * width + c);
        } else {
            const float g = __ldg(d_y + (long long)i * geo.T + t);
            const float4 gt = *reinterpret_cast<const float4*>(gtot + (long long)i * width + c);
            a = make_float4(g * gt.x, g * gt.y, g * gt.z, g * gt.w);
        }
        v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    float* dst = d_wout + (long long)t * ldw + c;
    if (accumulate) {
        const float4 o = *reinterpret_cast<const float4*>(dst);
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
    }
    *reinterpret_cast<float4*>(dst) = v;
}

// ------------------------------------------------------------------------------------------------
// backward (data): d_H[row,:] = (H > 0) * w * g[delay],  d_w = H . g[delay],  g = suffix sum over t
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
collapse_bwd_data_kernel(const Geom geo, const Planes act, int width, const int* __restrict__ order,
                         const int* __restrict__ sdelay, const float* __restrict__ sw, const float* __restrict__ w_out,
                         long long ldw, const float* __restrict__ d_y, __nv_bfloat16* __restrict__ d_act, long long ld_d,
                         long long d_plane, float* __restrict__ d_w) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int bsi = blockIdx.x, b = bsi / geo.S, s = bsi - b * geo.S;
    const int R = geo.R, T = geo.T;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int c = threadIdx.x * COL_PER_THREAD;
    const bool col_ok = c < width;
    RayList L = stage_rays(smem, R, (long long)bsi * R, order, sdelay, sw);
    float* dy_s = reinterpret_cast<float*>(smem + (size_t)R * 12);          // [T]
    float* dwp = dy_s + T;                                                   // [n_warps][R]
    for (int i = threadIdx.x; i < T; i += blockDim.x) dy_s[i] = __ldg(d_y + (long long)bsi * T + i);
    __syncthreads();
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    int t = T - 1;
    for (int k1 = R; k1 > 0; k1 -= RAY_BATCH) {
        uint2 hi[RAY_BATCH], lo[RAY_BATCH];
#pragma unroll
        for (int j = 0; j < RAY_BATCH; ++j) {
            const int k = k1 - 1 - j;
            hi[j] = make_uint2(0u, 0u); lo[j] = make_uint2(0u, 0u);
            if (k >= 0 && col_ok) ld_row4(act, ((long long)b * R + L.ord[k]) * geo.S + s, c, hi[j], lo[j]);
        }
#pragma unroll
        for (int j = 0; j < RAY_BATCH; ++j) {
            const int k = k1 - 1 - j;
            if (k < 0) break;
            const int d = L.del[k];
            while (t >= d) {                                                 // include every t >= delay
                if (col_ok) {
                    const float gy = dy_s[t];
                    const float4 wv = __ldg(reinterpret_cast<const float4*>(w_out + (long long)t * ldw + c));
                    g[0] = fmaf(gy, wv.x, g[0]); g[1] = fmaf(gy, wv.y, g[1]);
                    g[2] = fmaf(gy, wv.z, g[2]); g[3] = fmaf(gy, wv.w, g[3]);
                }
                --t;
            }
            float x[4];
            unpack4(hi[j], lo[j], x);
            const float wk = L.wgt[k];
            float dot = x[0] * g[0] + x[1] * g[1] + x[2] * g[2] + x[3] * g[3];
            dot = warp_sum(dot);
            if (lane == 0) dwp[warp * R + k] = dot;
            if (col_ok) {
                const long long row = ((long long)b * R + L.ord[k]) * geo.S + s;
                uint32_t l0, l1;
                const uint32_t h0 = pack_bf2(x[0] > 0.f ? wk * g[0] : 0.f, x[1] > 0.f ? wk * g[1] : 0.f, l0);
                const uint32_t h1 = pack_bf2(x[2] > 0.f ? wk * g[2] : 0.f, x[3] > 0.f ? wk * g[3] : 0.f, l1);
                *reinterpret_cast<uint2*>(d_act + row * ld_d + c) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(d_act + row * ld_d + c + d_plane) = make_uint2(l0, l1);
            }
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < R; k += blockDim.x) {
        float v = 0.f;
        for (int wq = 0; wq < n_warps; ++wq) v += dwp[wq * R + k];
        d_w[((long long)b * R + L.ord[k]) * geo.S + s] = v;
    }
}

static int collapse_threads(int width) {
    int th = (width / COL_PER_THREAD + 31) / 32 * 32;
    return th < 32 ? 32 : th;
}

static int check_collapse(const Geom& geo, int width, const void* act, long long ld, long long plane) {
    if (width % 8 != 0 || width <= 0 || width > 1024) return fail(AVR_ERR_UNSUPPORTED, "collapse: hidden width %d must be a multiple of 8 and <= 1024", width);
    if (geo.R < 1 || geo.T < 1) return fail(AVR_ERR_INVALID, "collapse: empty geometry");
    if (!act || ld % 4 != 0 || plane % 4 != 0 || (reinterpret_cast<uintptr_t>(act) & 7u)) return fail(AVR_ERR_INVALID, "collapse: activation planes must be 8-byte aligned");
    return AVR_OK;
}

}  // namespace avr

using namespace avr;

extern "C" {

AVR_API int avr_delay_sort(const avr_render_geom* geom, const int32_t* delay, const float* w, int32_t* order,
                           int32_t* sdelay, float* sw, int device, void* stream) {
    AVR_REQUIRE(geom && delay && w && order && sdelay && sw, "null pointer");
    AVR_ENTER(device);
    const Geom geo = make_geom(geom);
    const int n = geo.bs * geo.S;
    if (n == 0 || geo.R == 0) return AVR_OK;
    const size_t smem = (size_t)geo.T * sizeof(int);
    AVR_CUDA(cudaFuncSetAttribute(delay_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    delay_sort_kernel<<<n, 32, smem, (cudaStream_t)stream>>>(geo, delay, w, order, sdelay, sw);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

AVR_API int avr_collapse_fwd(const avr_render_geom* geom, const void* act_planes, int64_t ld_act, int64_t act_plane,
                             int32_t width, const int32_t* order, const int32_t* sdelay, const float* sw,
                             const float* w_out, int64_t ldw, float* y, int device, void* stream) {
    AVR_REQUIRE(geom && order && sdelay && sw && w_out && y, "null pointer");
    AVR_ENTER(device);
    const Geom geo = make_geom(geom);
    if (int rc = check_collapse(geo, width, act_planes, ld_act, act_plane)) return rc;
    AVR_REQUIRE(ldw % 4 == 0 && aligned16(w_out), "W_out must be 16-byte aligned");
    const int n = geo.bs * geo.S;
    if (n == 0) return AVR_OK;
    const int threads = collapse_threads(width);
    const size_t smem = (size_t)geo.R * 12 + (size_t)(threads / 32) * geo.T * sizeof(float);
    AVR_CUDA(cudaFuncSetAttribute(collapse_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const Planes act = {(const __nv_bfloat16*)act_planes, ld_act, act_plane};
    collapse_fwd_kernel<0><<<n, threads, smem, (cudaStream_t)stream>>>(geo, act, width, order, sdelay, sw, w_out, ldw, y,
                                                                     nullptr, nullptr, nullptr, nullptr, 0);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

AVR_API int avr_collapse_bwd_data(const avr_render_geom* geom, const void* act_planes, int64_t ld_act, int64_t act_plane,
                                  int32_t width, const int32_t* order, const int32_t* sdelay, const float* sw,
                                  const float* w_out, int64_t ldw, const float* d_y, void* d_act_planes, int64_t ld_d,
                                  int64_t d_plane, float* d_w, int device, void* stream) {
    AVR_REQUIRE(geom && order && sdelay && sw && w_out && d_y && d_act_planes && d_w, "null pointer");
    AVR_ENTER(device);
    const Geom geo = make_geom(geom);
    if (int rc = check_collapse(geo, width, act_planes, ld_act, act_plane)) return rc;
    AVR_REQUIRE(ldw % 4 == 0 && aligned16(w_out), "W_out must be 16-byte aligned");
    AVR_REQUIRE(ld_d % 4 == 0 && d_plane % 4 == 0 && (reinterpret_cast<uintptr_t>(d_act_planes) & 7u) == 0, "d_act misaligned");
    const int n = geo.bs * geo.S;
    if (n == 0) return AVR_OK;
    const int threads = collapse_threads(width);
    const size_t smem = (size_t)geo.R * 12 + (size_t)geo.T * 4 + (size_t)(threads / 32) * geo.R * 4;
    AVR_CUDA(cudaFuncSetAttribute(collapse_bwd_data_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const Planes act = {(const __nv_bfloat16*)act_planes, ld_act, act_plane};
    collapse_bwd_data_kernel<<<n, threads, smem, (cudaStream_t)stream>>>(geo, act, width, order, sdelay, sw, w_out, ldw, d_y,
                                                                       (__nv_bfloat16*)d_act_planes, ld_d, d_plane, d_w);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

AVR_API int64_t avr_collapse_bwd_weight_workspace_bytes(const avr_render_geom* geom, int32_t width, int32_t tspan) {
    if (!geom) return 0;
    const int64_t n = (int64_t)geom->bs * geom->S;
    return (n * tspan * width + n * width) * (int64_t)sizeof(float) + n * 2 * (int64_t)sizeof(int) + 64;
}

// tspan: static bound on (max delay - min delay) within one (b,s); a violation poisons the result with NaN.
AVR_API int avr_collapse_bwd_weight(const avr_render_geom* geom, const void* act_planes, int64_t ld_act, int64_t act_plane,
                                    int32_t width, const int32_t* order, const int32_t* sdelay, const float* sw,
                                    const float* d_y, float* d_wout, int64_t ldw, int accumulate, int32_t tspan,
                                    void* workspace, int64_t workspace_bytes, int device, void* stream) {
    AVR_REQUIRE(geom && order && sdelay && sw && d_y && d_wout && workspace, "null pointer");
    AVR_ENTER(device);
    const Geom geo = make_geom(geom);
    if (int rc = check_collapse(geo, width, act_planes, ld_act, act_plane)) return rc;
    AVR_REQUIRE(tspan > 0 && workspace_bytes >= avr_collapse_bwd_weight_workspace_bytes(geom, width, tspan), "workspace too small");
    AVR_REQUIRE(ldw % 4 == 0 && aligned16(d_wout) && aligned16(workspace), "buffers must be 16-byte aligned");
    const int64_t n = (int64_t)geo.bs * geo.S;
    if (n == 0) return AVR_OK;
    float* P = (float*)workspace;
    float* gtot = P + n * tspan * width;
    int* range = (int*)(gtot + n * width);
    const int threads = collapse_threads(width);
    const size_t smem = (size_t)geo.R * 12 + (size_t)geo.T * 4;
    cudaStream_t st = (cudaStream_t)stream;
    AVR_CUDA(cudaFuncSetAttribute(collapse_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const Planes act = {(const __nv_bfloat16*)act_planes, ld_act, act_plane};
    collapse_fwd_kernel<1><<<(unsigned)n, threads, smem, st>>>(geo, act, width, order, sdelay, sw, nullptr, 0, nullptr, d_y, P,
                                                             gtot, range, tspan);
    AVR_LAUNCH_CHECK();
    const int total = geo.T * (width / 4);
    collapse_dw_reduce_kernel<<<(total + 127) / 128, 128, 0, st>>>(geo, width, d_y, P, gtot, range, tspan, d_wout, ldw,
                                                                 accumulate);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

}  // extern "C"
