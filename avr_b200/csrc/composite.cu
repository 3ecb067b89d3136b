// Density -> ray weights, time-domain ray reduction (compositing), spectrum (mask * path loss, DFT,
// per-sample phase, sum over samples), and the broadcast / reduce helpers of the signal-network inputs.
//
// Replaces renderer.py:79-121 and 167-193.  Everything after the network is
// linear in the signal and the phase / path loss depend only on the sample index s, so the ray sum is
// taken in the time domain first (SURVEY App. A "reordered"): the [ray, sample, freq] tensor of the
// reference never exists, `sig` is streamed exactly once, and only bs*S rows reach the DFT.
#include <cuda_bf16.h>
#include <math.h>
#include "common.cuh"

namespace avr {

int gemm_impl(int la, int lb, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb,
              float* C, int64_t ldc, int flags, const float* aux, int64_t ldaux, void* workspace, int64_t workspace_bytes,
              cudaStream_t st);

// ------------------------------------------------------------------------------------------------
// ray weights: one warp per ray, warp-shuffle scans along the samples
// ------------------------------------------------------------------------------------------------
constexpr int RW_MAX_CHUNKS = 8;   // S <= 256

// attn = |leaky_relu(raw, slope)| (model.py:233,329); slope < 0 selects the identity (raw already is attn)
__device__ __forceinline__ float attn_of(float raw, float slope) {
    if (slope < 0.f) return raw;
    return fabsf(raw > 0.f ? raw : __fmul_rn(raw, slope));
}

__device__ __forceinline__ float warp_incl_prod(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = __fmul_rn(v, u);
    }
    return v;
}

__device__ __forceinline__ float warp_incl_sum_rev(float v, int lane) {   // suffix sums within the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float u = __shfl_down_sync(0xffffffffu, v, o);
        if (lane + o < 32) v += u;
    }
    return v;
}

template <bool BACKWARD>
__global__ void __launch_bounds__(256)
ray_weights_kernel(const Geom geo, const float* __restrict__ raw, int64_t ld_raw, const float* __restrict__ delta,
                   float slope, float* __restrict__ attn_out, float* __restrict__ w_out, const float* __restrict__ d_w,
                   float* __restrict__ d_raw, int64_t ld_draw) {
    const int lane = threadIdx.x & 31;
    const int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ray >= (int64_t)geo.bs * geo.R) return;
    const int S = geo.S;
    const int64_t base = ray * S;
    const int n_chunks = (S + 31) >> 5;

    float alpha[RW_MAX_CHUNKS], trans[RW_MAX_CHUNKS], att[RW_MAX_CHUNKS];
    float carry = 1.0f;
#pragma unroll
    for (int c = 0; c < RW_MAX_CHUNKS; ++c) {
        if (c >= n_chunks) break;
        const int s = c * 32 + lane;
        float a = 0.f, at = 0.f;
        if (s < S) {
            at = attn_of(__ldg(raw + (base + s) * ld_raw), slope);
            a = __fsub_rn(1.0f, expf(__fmul_rn(-at, __ldg(delta + s))));
        }
        const float q = (s < S) ? __fadd_rn(__fsub_rn(1.0f, a), 1e-6f) : 1.0f;
        const float incl = warp_incl_prod(q, lane);
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.0f;
        const float tr = __fmul_rn(carry, excl);
        carry = __fmul_rn(carry, __shfl_sync(0xffffffffu, incl, 31));
        alpha[c] = a; trans[c] = tr; att[c] = at;
        if (!BACKWARD && s < S) {
            w_out[base + s] = __fmul_rn(tr, a);
            if (attn_out) attn_out[base + s] = at;
        }
    }
    if (!BACKWARD) return;

    float tail = 0.f;                                  // sum_{k > s} d_w[k] * w[k], carried across chunks
#pragma unroll
    for (int c = RW_MAX_CHUNKS - 1; c >= 0; --c) {
        if (c >= n_chunks) continue;
        const int s = c * 32 + lane;
        const float dw = (s < S) ? __ldg(d_w + base + s) : 0.f;
        const float p = dw * trans[c] * alpha[c];
        const float incl = warp_incl_sum_rev(p, lane);
        const float after = incl - p + tail;           // strictly-after sum
        tail += __shfl_sync(0xffffffffu, incl, 0);
        if (s < S) {
            const float q = __fadd_rn(__fsub_rn(1.0f, alpha[c]), 1e-6f);
            const float d_alpha = dw * trans[c] - after / q;
            const float dl = __ldg(delta + s);
            const float d_att = dl * expf(__fmul_rn(-att[c], dl)) * d_alpha;
            const float x = __ldg(raw + (base + s) * ld_raw);
            const float g = slope < 0.f ? 1.0f : (x > 0.f ? 1.0f : (x < 0.f ? -slope : 0.0f));
            d_raw[(base + s) * ld_draw] = d_att * g;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// compositing forward: y[b,s,t] = sum_r w * (t >= delay) * sig      (sig streamed once)
// grid (bs*S, n_chunks): a CTA owns one (b,s) and a chunk of rays; threads own float4 columns of t.
// ------------------------------------------------------------------------------------------------
constexpr int CF_THREADS = 256;
constexpr int CF_MAXQ = 4;        // float4 columns per thread -> T <= 4096
constexpr int CF_UNROLL = 4;      // rays in flight

__device__ __forceinline__ float4 ld_stream(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }

template <int NQ>
__global__ void __launch_bounds__(CF_THREADS)
composite_fwd_kernel(const Geom geo, const float* __restrict__ sig, const float* __restrict__ w,
                     const int* __restrict__ delay, float* __restrict__ partial, int rays_per_chunk) {
    const int bsi = blockIdx.x;                        // b*S + s
    const int b = bsi / geo.S, s = bsi - b * geo.S;
    const int r_beg = blockIdx.y * rays_per_chunk;
    const int r_end = min(geo.R, r_beg + rays_per_chunk);
    const int T = geo.T, TQ = T >> 2;
    float4 acc[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int r0 = r_beg; r0 < r_end; r0 += CF_UNROLL) {
        float wr[CF_UNROLL];
        int dr[CF_UNROLL];
        float4 v[CF_UNROLL][NQ];
#pragma unroll
        for (int u = 0; u < CF_UNROLL; ++u) {
            const int r = r0 + u;
            const int64_t n = ((int64_t)b * geo.R + r) * geo.S + s;
            wr[u] = (r < r_end) ? __ldg(w + n) : 0.f;
            dr[u] = (r < r_end) ? __ldg(delay + n) : T;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int tq = threadIdx.x + q * CF_THREADS;
                v[u][q] = make_float4(0.f, 0.f, 0.f, 0.f);
                // skip quads that are masked out entirely (t < delay) and rays that carry no weight
                if (tq < TQ && wr[u] != 0.f && tq * 4 + 3 >= dr[u]) v[u][q] = ld_stream(sig + n * T + tq * 4);
            }
        }
#pragma unroll
        for (int u = 0; u < CF_UNROLL; ++u) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int t = (threadIdx.x + q * CF_THREADS) * 4;
                const float ww = wr[u];
                const int d = dr[u];
                acc[q].x = fmaf(t + 0 >= d ? ww : 0.f, v[u][q].x, acc[q].x);
                acc[q].y = fmaf(t + 1 >= d ? ww : 0.f, v[u][q].y, acc[q].y);
                acc[q].z = fmaf(t + 2 >= d ? ww : 0.f, v[u][q].z, acc[q].z);
                acc[q].w = fmaf(t + 3 >= d ? ww : 0.f, v[u][q].w, acc[q].w);
            }
        }
    }
    float* out = partial + ((int64_t)blockIdx.y * gridDim.x + bsi) * T;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const int tq = threadIdx.x + q * CF_THREADS;
        if (tq < TQ) *reinterpret_cast<float4*>(out + tq * 4) = acc[q];
    }
}

// y = sum over ray chunks (fixed order), optionally scaled by gain[s,t]
__global__ void chunk_reduce_kernel(const float* __restrict__ partial, int n_chunks, int64_t n_rows, int T, int S,
                                    const float* __restrict__ gain, float* __restrict__ y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // float4 index
    const int TQ = T >> 2;
    if (i >= n_rows * TQ) return;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = 0; c < n_chunks; ++c) {
        const float4 p = *reinterpret_cast<const float4*>(partial + (int64_t)c * n_rows * T + i * 4);
        v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    }
    if (gain) {
        const int64_t row = i / TQ;
        const int tq = (int)(i - row * TQ);
        const float4 g = *reinterpret_cast<const float4*>(gain + (int64_t)(row % S) * T + tq * 4);
        v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w;
    }
    *reinterpret_cast<float4*>(y + i * 4) = v;
}

// ------------------------------------------------------------------------------------------------
// compositing backward: d_sig = w*(t>=delay)*d_y ; d_w = sum_t (t>=delay)*sig*d_y
// grid (bs*S, n_chunks): d_y[b,s,:] is staged in shared memory once, each warp then owns whole rays.
// ------------------------------------------------------------------------------------------------
constexpr int CB_THREADS = 256;

__global__ void __launch_bounds__(CB_THREADS)
composite_bwd_kernel(const Geom geo, const float* __restrict__ sig, const float* __restrict__ w,
                     const int* __restrict__ delay, const float* __restrict__ d_y, float* __restrict__ d_sig,
                     float* __restrict__ d_w, int rays_per_chunk) {
    extern __shared__ __align__(16) float dy_s[];
    const int bsi = blockIdx.x;
    const int b = bsi / geo.S, s = bsi - b * geo.S;
    const int T = geo.T, TQ = T >> 2;
    for (int q = threadIdx.x; q < TQ; q += CB_THREADS)
        reinterpret_cast<float4*>(dy_s)[q] = *reinterpret_cast<const float4*>(d_y + (int64_t)bsi * T + q * 4);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r_beg = blockIdx.y * rays_per_chunk;
    const int r_end = min(geo.R, r_beg + rays_per_chunk);
    for (int r = r_beg + warp; r < r_end; r += CB_THREADS / 32) {
        const int64_t n = ((int64_t)b * geo.R + r) * geo.S + s;
        const float ww = __ldg(w + n);
        const int d = __ldg(delay + n);
        float dot = 0.f;
        for (int q0 = 0; q0 < TQ; q0 += 128) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int q = q0 + u * 32 + lane;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (sig && q < TQ && q * 4 + 3 >= d) v[u] = ld_stream(sig + n * T + q * 4);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int q = q0 + u * 32 + lane;
                if (q >= TQ) continue;
                const int t = q * 4;
                float4 g = reinterpret_cast<const float4*>(dy_s)[q];
                g.x = t + 0 >= d ? g.x : 0.f; g.y = t + 1 >= d ? g.y : 0.f;
                g.z = t + 2 >= d ? g.z : 0.f; g.w = t + 3 >= d ? g.w : 0.f;
                dot = fmaf(v[u].x, g.x, dot); dot = fmaf(v[u].y, g.y, dot);
                dot = fmaf(v[u].z, g.z, dot); dot = fmaf(v[u].w, g.w, dot);
                if (d_sig) __stcs(reinterpret_cast<float4*>(d_sig + n * T + q * 4),
                                  make_float4(ww * g.x, ww * g.y, ww * g.z, ww * g.w));
            }
        }
        if (d_w) {
            dot = warp_sum(dot);
            if (lane == 0) d_w[n] = dot;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// spectrum helpers
// ------------------------------------------------------------------------------------------------
// z[row, t] = y[row, t] * gain[row % S, t]
__device__ __forceinline__ void split3_store(__nv_bfloat16* dst, int64_t plane, int np, float v) {
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(hi);
    const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
    dst[0] = hi;
    dst[plane] = mid;
    if (np == 3) dst[2 * plane] = __float2bfloat16_rn(r1 - __bfloat162float(mid));
}

// z = y * gain written as a bf16 plane set (operand of the tensor-core DFT)
__global__ void gain_planes_kernel(const float* __restrict__ y, const float* __restrict__ gain, int64_t n_rows, int T, int S,
                                   __nv_bfloat16* __restrict__ z, int64_t ldz, int64_t plane, int np) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows * T) return;
    const int64_t row = i / T;
    const int t = (int)(i - row * T);
    split3_store(z + row * ldz + t, plane, np, y[i] * __ldg(gain + (int64_t)(row % S) * T + t));
}

__global__ void gain_kernel(const float* y, const float* __restrict__ gain, int64_t n_rows, int T, int S, float* z) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int TQ = T >> 2;
    if (i >= n_rows * TQ) return;
    const int64_t row = i / TQ;
    const int tq = (int)(i - row * TQ);
    float4 v = *reinterpret_cast<const float4*>(y + i * 4);
    const float4 g = *reinterpret_cast<const float4*>(gain + (int64_t)(row % S) * T + tq * 4);
    v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w;
    *reinterpret_cast<float4*>(z + i * 4) = v;
}

// out[b,f] = sum_s X[b,s,f] * phase[s,f]   (complex, s ascending)
__global__ void phase_sum_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ phase, int bs, int S,
                                 int F, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= bs * F) return;
    const int b = i / F, f = i - b * F;
    float re = 0.f, im = 0.f;
    for (int s = 0; s < S; ++s) {
        const float2 xv = *reinterpret_cast<const float2*>(x + ((int64_t)b * S + s) * ldx + 2 * f);
        const float2 p = *reinterpret_cast<const float2*>(phase + ((int64_t)s * F + f) * 2);
        re = fmaf(xv.x, p.x, re); re = fmaf(-xv.y, p.y, re);
        im = fmaf(xv.x, p.y, im); im = fmaf(xv.y, p.x, im);
    }
    out[2 * i] = re;
    out[2 * i + 1] = im;
}

// d_X[b,s,f] = d_out[b,f] * conj(phase[s,f]) ; pad columns [2F, ldx) are zeroed.  dx: fp32 or bf16 plane set.
__global__ void phase_bwd_kernel(const float* __restrict__ d_out, const float* __restrict__ phase, int bs, int S, int F,
                                 int64_t ldx, float* __restrict__ dx, __nv_bfloat16* __restrict__ dxp, int64_t plane, int np) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t half = ldx / 2;
    if (i >= (int64_t)bs * S * half) return;
    const int64_t row = i / half;
    const int f = (int)(i - row * half);
    float2 o = make_float2(0.f, 0.f);
    if (f < F) {
        const int b = (int)(row / S), s = (int)(row - (int64_t)b * S);
        const float2 g = *reinterpret_cast<const float2*>(d_out + ((int64_t)b * F + f) * 2);
        const float2 p = *reinterpret_cast<const float2*>(phase + ((int64_t)s * F + f) * 2);
        o.x = fmaf(g.x, p.x, g.y * p.y);
        o.y = fmaf(g.y, p.x, -g.x * p.y);
    }
    if (dxp) {
        split3_store(dxp + row * ldx + 2 * f, plane, np, o.x);
        split3_store(dxp + row * ldx + 2 * f + 1, plane, np, o.y);
    } else {
        *reinterpret_cast<float2*>(dx + row * ldx + 2 * f) = o;
    }
}

// ------------------------------------------------------------------------------------------------
// broadcast rows into / reduce rows out of the signal-network input buffer
// ------------------------------------------------------------------------------------------------
__global__ void rows_broadcast_kernel(const Geom geo, const float* __restrict__ src, int w, int per_receiver,
                                      void* __restrict__ dst_v, int64_t ld_dst, int64_t dst_plane, int dst_np, int col0) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n_pts = (int64_t)geo.bs * geo.R * geo.S;
    if (i >= n_pts * w) return;
    const int64_t n = i / w;
    const int c = (int)(i - n * w);
    const int64_t row = per_receiver ? n / ((int64_t)geo.R * geo.S) : (n / geo.S) % geo.R;
    const float v = __ldg(src + row * w + c);
    const int64_t o = n * ld_dst + col0 + c;
    if (dst_plane == 0) reinterpret_cast<float*>(dst_v)[o] = v;
    else planes_store(dst_v, o, dst_plane, dst_np, v);
}

// plane-set destination, 8 columns (one 16-byte store per plane) per thread.  Two adjacent blocks (src, then src2) can be
// written by one launch: a 40-column block is 2.5 sectors per plane row, two of them side by side are whole sectors.
__global__ void rows_broadcast_planes8_kernel(const Geom geo, const float* __restrict__ src, int w, int per_receiver,
                                              const float* __restrict__ src2, int w2, int per_receiver2,
                                              __nv_bfloat16* __restrict__ db, int64_t ld_dst, int64_t dst_plane, int dst_np,
                                              int col0) {
    const int groups = (w + w2) >> 3, g1 = w >> 3;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n_pts = (int64_t)geo.bs * geo.R * geo.S;
    if (i >= n_pts * groups) return;
    const int64_t n = i / groups;
    const int g = (int)(i - n * groups);
    const bool second = g >= g1;
    const bool rcv = second ? per_receiver2 != 0 : per_receiver != 0;
    const int64_t row = rcv ? n / ((int64_t)geo.R * geo.S) : (n / geo.S) % geo.R;
    const float* sp = second ? src2 + row * w2 + 8 * (g - g1) : src + row * w + 8 * g;
    const float4 a = __ldg(reinterpret_cast<const float4*>(sp));
    const float4 b = __ldg(reinterpret_cast<const float4*>(sp) + 1);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t h[4], m[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) planes_pack2(dst_np, v[2 * k], v[2 * k + 1], h[k], m[k], l[k]);
    __nv_bfloat16* o = db + n * ld_dst + col0 + 8 * g;
    *reinterpret_cast<uint4*>(o) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(o + dst_plane) = make_uint4(m[0], m[1], m[2], m[3]);
    if (planes_count(dst_np) == 3) *reinterpret_cast<uint4*>(o + 2 * dst_plane) = make_uint4(l[0], l[1], l[2], l[3]);
}

// plane-set source, 8 columns (one 16-byte load per plane) per thread; same partial layout and a fixed summation order
__global__ void __launch_bounds__(256)
rows_reduce_planes8_kernel(const Geom geo, const __nv_bfloat16* __restrict__ db, int64_t ld_dst, int64_t d_plane, int col0,
                           int w, int per_receiver, int items_per_chunk, float* __restrict__ partial) {
    extern __shared__ float red8[];                               // [items per pass][w + 1]
    const int row = blockIdx.x, chunk = blockIdx.y;
    const int groups = w >> 3, ipp = 256 / groups;
    const int g = threadIdx.x % groups, y = threadIdx.x / groups;
    const int64_t n_items = per_receiver ? (int64_t)geo.R * geo.S : (int64_t)geo.bs * geo.S;
    const int64_t j_beg = (int64_t)chunk * items_per_chunk;
    const int64_t j_end = min(n_items, j_beg + items_per_chunk);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (y < ipp) {
        for (int64_t j = j_beg + y; j < j_end; j += ipp) {
            int64_t n;
            if (per_receiver) n = (int64_t)row * geo.R * geo.S + j;
            else { const int64_t b = j / geo.S, sx = j - b * geo.S; n = (b * geo.R + row) * geo.S + sx; }
            const __nv_bfloat16* src = db + n * ld_dst + col0 + 8 * g;
            const uint4 h = __ldg(reinterpret_cast<const uint4*>(src));
            const uint4 m = __ldg(reinterpret_cast<const uint4*>(src + d_plane));
            const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                acc[2 * k] += __uint_as_float(hw[k] << 16) + __uint_as_float(mw[k] << 16);
                acc[2 * k + 1] += __uint_as_float(hw[k] & 0xffff0000u) + __uint_as_float(mw[k] & 0xffff0000u);
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) red8[y * (w + 1) + 8 * g + k] = acc[k];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < w; c += 256) {
        float t = 0.f;
        for (int k = 0; k < ipp; ++k) t += red8[k * (w + 1) + c];
        partial[((int64_t)row * gridDim.y + chunk) * w + c] = t;
    }
}

// stage 1: partial[row, chunk, c] = sum over the chunk's contributing points (fixed order)
__global__ void __launch_bounds__(256)
rows_reduce_kernel(const Geom geo, const void* __restrict__ d_dst_v, int64_t ld_dst, int64_t d_plane, int col0, int w,
                   int per_receiver, int items_per_chunk, float* __restrict__ partial) {
    __shared__ float red[8][33];
    const int row = blockIdx.x, chunk = blockIdx.y;
    const int lane = threadIdx.x & 31, y = threadIdx.x >> 5;
    const int64_t n_items = per_receiver ? (int64_t)geo.R * geo.S : (int64_t)geo.bs * geo.S;
    const int64_t j_beg = (int64_t)chunk * items_per_chunk;
    const int64_t j_end = min(n_items, j_beg + items_per_chunk);
    for (int c0 = 0; c0 < w; c0 += 32) {
        const int c = c0 + lane;
        float acc = 0.f;
        if (c < w) {
            for (int64_t j = j_beg + y; j < j_end; j += 8) {
                int64_t n;
                if (per_receiver) n = (int64_t)row * geo.R * geo.S + j;
                else { const int64_t b = j / geo.S, s = j - b * geo.S; n = (b * geo.R + row) * geo.S + s; }
                const int64_t o = n * ld_dst + col0 + c;
                if (d_plane == 0) acc += __ldg(reinterpret_cast<const float*>(d_dst_v) + o);
                else {
                    const __nv_bfloat16* db = reinterpret_cast<const __nv_bfloat16*>(d_dst_v);
                    acc += __bfloat162float(db[o]) + __bfloat162float(db[o + d_plane]);
                }
            }
        }
        red[y][lane] = acc;
        __syncthreads();
        if (y == 0 && c < w) {
            float t = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) t += red[k][lane];
            partial[((int64_t)row * gridDim.y + chunk) * w + c] = t;
        }
        __syncthreads();
    }
}

__global__ void rows_reduce_final_kernel(const float* __restrict__ partial, int n_rows, int n_chunks, int w,
                                         float* __restrict__ d_src) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows * w) return;
    const int row = i / w, c = i - row * w;
    float t = 0.f;
    for (int k = 0; k < n_chunks; ++k) t += partial[((int64_t)row * n_chunks + k) * w + c];
    d_src[i] = t;
}

static int composite_chunks(const Geom& g, int* rays_per_chunk) {
    const int64_t rows = (int64_t)g.bs * g.S;
    int64_t chunks = rows > 0 ? ceil_div(148 * 8, rows) : 1;
    if (chunks > 32) chunks = 32;
    if (chunks > g.R) chunks = g.R;
    if (chunks < 1) chunks = 1;
    int rpc = (int)ceil_div(g.R, chunks);
    rpc = (int)(ceil_div(rpc, 8) * 8);
    *rays_per_chunk = rpc;
    return (int)ceil_div(g.R, rpc);
}

static int rows_reduce_plan(const Geom& g, int per_receiver, int* items_per_chunk) {
    const int64_t n_items = per_receiver ? (int64_t)g.R * g.S : (int64_t)g.bs * g.S;
    const int64_t ipc = 2048;
    *items_per_chunk = (int)ipc;
    return (int)(n_items > 0 ? ceil_div(n_items, ipc) : 1);
}

}  // namespace avr

using namespace avr;

extern "C" int avr_ray_weights_fwd(const avr_render_geom* geom, const float* raw, int64_t ld_raw, const float* delta,
                                   float slope, float* attn, float* w, int device, void* stream) {
    AVR_REQUIRE(geom && raw && delta && w, "null pointer");
    AVR_ENTER(device);
    const Geom geo = make_geom(geom);
    if (geo.S > 32 * RW_MAX_CHUNKS) return fail(AVR_ERR_UNSUPPORTED, "n_samples=%d > %d", geo.S, 32 * RW_MAX_CHUNKS);
    const int64_t rays = (int64_t)geo.bs * geo.R;
    if (rays == 0 || geo.S == 0) return AVR_OK;
    ray_weights_kernel<false><<<(unsigned)ceil_div(rays, 8), 256, 0, (cudaStream_t)stream>>>(
        geo, raw, ld_raw, delta, slope, attn, w, nullptr, nullptr, 0);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

extern "C" int avr_ray_weights_bwd(const avr_render_geom* geom, const float* raw, int64_t ld_raw, const float* delta,
                                   float slope, const float* d_w, float* d_raw, int64_t ld_draw, int device,
                                   void* stream) {
    AVR_REQUIRE(geom && raw && delta && d_w && d_raw, "null pointer");
    AVR_ENTER(device);
    const Geom geo = make_geom(geom);
    if (geo.S > 32 * RW_MAX_CHUNKS) return fail(AVR_ERR_UNSUPPORTED, "n_samples=%d > %d", geo.S, 32 * RW_MAX_CHUNKS);
    const int64_t rays = (int64_t)geo.bs * geo.R;
    if (rays == 0 || geo.S == 0) return AVR_OK;
    ray_weights_kernel<true><<<(unsigned)ceil_div(rays, 8), 256, 0, (cudaStream_t)stream>>>(
        geo, raw, ld_raw, delta, slope, nullptr, nullptr, d_w, d_raw, ld_draw);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

extern "C" int64_t avr_composite_workspace_bytes(const avr_render_geom* geom) {
    if (!geom) return 0;
    const Geom geo = make_geom(geom);
    int rpc;
    const int chunks = composite_chunks(geo, &rpc);
    return (int64_t)chunks * geo.bs * geo.S * geo.T * (int64_t)sizeof(float);
}

extern "C" int avr_composite_fwd(const avr_render_geom* geom, const float* sig, const float* w, const int32_t* delay,
                                 float* y, void* workspace, int64_t workspace_bytes, int device, void* stream) {
    AVR_REQUIRE(geom && sig && w && delay && y && workspace, "null pointer");
    AVR_ENTER(device);
    const Geom geo = make_geom(geom);
    if (geo.T % 4 != 0 || geo.T > 4 * CF_THREADS * CF_MAXQ)
        return fail(AVR_ERR_UNSUPPORTED, "IR length T=%d must be a multiple of 4 and <= %d", geo.T, 4 * CF_THREADS * CF_MAXQ);
    AVR_REQUIRE(aligned16(sig) && aligned16(y) && aligned16(workspace), "buffers must be 16-byte aligned");
    AVR_REQUIRE(workspace_bytes >= avr_composite_workspace_bytes(geom), "workspace too small");
    const int64_t rows = (int64_t)geo.bs * geo.S;
    if (rows == 0 || geo.T == 0) return AVR_OK;
    int rpc;
    const int chunks = composite_chunks(geo, &rpc);
    const dim3 grid((unsigned)rows, (unsigned)chunks);
    const int nq = (int)ceil_div(geo.T / 4, CF_THREADS);
    float* partial = (float*)workspace;
    cudaStream_t st = (cudaStream_t)stream;
    switch (nq) {
        case 1: composite_fwd_kernel<1><<<grid, CF_THREADS, 0, st>>>(geo, sig, w, delay, partial, rpc); break;
        case 2: composite_fwd_kernel<2><<<grid, CF_THREADS, 0, st>>>(geo, sig, w, delay, partial, rpc); break;
        case 3: composite_fwd_kernel<3><<<grid, CF_THREADS, 0, st>>>(geo, sig, w, delay, partial, rpc); break;
        default: composite_fwd_kernel<4><<<grid, CF_THREADS, 0, st>>>(geo, sig, w, delay, partial, rpc); break;
    }
    AVR_LAUNCH_CHECK();
    const int64_t nq4 = rows * (geo.T / 4);
    chunk_reduce_kernel<<<(unsigned)ceil_div(nq4, 256), 256, 0, st>>>(partial, chunks, rows, geo.T, geo.S, nullptr, y);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

extern "C" int avr_composite_bwd(const avr_render_geom* geom, const float* sig, const float* w, const int32_t* delay,
                                 const float* d_y, float* d_sig, float* d_w, int device, void* stream) {
    AVR_REQUIRE(geom && w && delay && d_y, "null pointer");
    AVR_REQUIRE(d_sig || d_w, "nothing to compute");
    AVR_REQUIRE(d_w == nullptr || sig != nullptr, "d_w needs sig");
    AVR_ENTER(device);
    const Geom geo = make_geom(geom);
    if (geo.T % 4 != 0) return fail(AVR_ERR_UNSUPPORTED, "IR length T=%d must be a multiple of 4", geo.T);
    AVR_REQUIRE(aligned16(d_y) && (!sig || aligned16(sig)) && (!d_sig || aligned16(d_sig)), "16-byte alignment");
    const int64_t rows = (int64_t)geo.bs * geo.S;
    if (rows == 0 || geo.T == 0) return AVR_OK;
    int rpc;
    const int chunks = composite_chunks(geo, &rpc);
    const dim3 grid((unsigned)rows, (unsigned)chunks);
    const size_t smem = (size_t)geo.T * sizeof(float);
    AVR_CUDA(cudaFuncSetAttribute(composite_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    composite_bwd_kernel<<<grid, CB_THREADS, smem, (cudaStream_t)stream>>>(geo, sig, w, delay, d_y, d_sig, d_w, rpc);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

extern "C" int avr_spectrum_fwd(const avr_render_geom* geom, const float* y, const float* gain, const float* phase,
                                const float* dft, int64_t ldd, float* zbuf, float* xbuf, float* out, int device,
                                void* stream) {
    AVR_REQUIRE(geom && y && gain && phase && dft && zbuf && xbuf && out, "null pointer");
    AVR_ENTER(device);
    const Geom geo = make_geom(geom);
    const int F = geo.T / 2 + 1;
    AVR_REQUIRE(geo.T % 4 == 0, "T must be a multiple of 4");
    AVR_REQUIRE(ldd >= 2 * F && ldd % 4 == 0, "ldd must be >= 2F and a multiple of 4");
    const int64_t rows = (int64_t)geo.bs * geo.S;
    if (rows == 0) return AVR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nq4 = rows * (geo.T / 4);
    gain_kernel<<<(unsigned)ceil_div(nq4, 256), 256, 0, st>>>(y, gain, rows, geo.T, geo.S, zbuf);
    AVR_LAUNCH_CHECK();
    if (int rc = gemm_impl(AVR_K_CONTIG, AVR_I_CONTIG, rows, ldd, geo.T, zbuf, geo.T, dft, ldd, xbuf, ldd, 0, nullptr, 0,
                           nullptr, 0, st))
        return rc;
    phase_sum_kernel<<<(unsigned)ceil_div((int64_t)geo.bs * F, 128), 128, 0, st>>>(xbuf, ldd, phase, geo.bs, geo.S, F, out);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

extern "C" int avr_spectrum_bwd(const avr_render_geom* geom, const float* d_out, const float* gain, const float* phase,
                                const float* dft, int64_t ldd, float* xbuf, float* d_y, int device, void* stream) {
    AVR_REQUIRE(geom && d_out && gain && phase && dft && xbuf && d_y, "null pointer");
    AVR_ENTER(device);
    const Geom geo = make_geom(geom);
    const int F = geo.T / 2 + 1;
    AVR_REQUIRE(geo.T % 4 == 0, "T must be a multiple of 4");
    AVR_REQUIRE(ldd >= 2 * F && ldd % 4 == 0, "ldd must be >= 2F and a multiple of 4");
    const int64_t rows = (int64_t)geo.bs * geo.S;
    if (rows == 0) return AVR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    phase_bwd_kernel<<<(unsigned)ceil_div(rows * (ldd / 2), 256), 256, 0, st>>>(d_out, phase, geo.bs, geo.S, F, ldd, xbuf,
                                                                              nullptr, 0, 0);
    AVR_LAUNCH_CHECK();
    if (int rc = gemm_impl(AVR_K_CONTIG, AVR_K_CONTIG, rows, geo.T, ldd, xbuf, ldd, dft, ldd, d_y, geo.T, 0, nullptr, 0,
                           nullptr, 0, st))
        return rc;
    const int64_t nq4 = rows * (geo.T / 4);
    gain_kernel<<<(unsigned)ceil_div(nq4, 256), 256, 0, st>>>(d_y, gain, rows, geo.T, geo.S, d_y);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

extern "C" int avr_rows_broadcast(const avr_render_geom* geom, const float* src, int32_t w, int per_receiver,
                                  void* dst, int64_t ld_dst, int64_t dst_plane, int32_t dst_nplanes, int32_t col0, int device,
                                  void* stream) {
    AVR_REQUIRE(geom && src && dst, "null pointer");
    AVR_REQUIRE(w > 0 && col0 >= 0 && ld_dst >= col0 + w, "bad column window");
    AVR_ENTER(device);
    const Geom geo = make_geom(geom);
    const int64_t total = (int64_t)geo.bs * geo.R * geo.S * w;
    if (total == 0) return AVR_OK;
    if (dst_plane != 0 && w % 8 == 0 && col0 % 8 == 0 && ld_dst % 8 == 0 && dst_plane % 8 == 0 && aligned16(dst) && aligned16(src))
        rows_broadcast_planes8_kernel<<<(unsigned)ceil_div(total / 8, 256), 256, 0, (cudaStream_t)stream>>>(
            geo, src, w, per_receiver, nullptr, 0, 0, (__nv_bfloat16*)dst, ld_dst, dst_plane, dst_nplanes, col0);
    else
        rows_broadcast_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(geo, src, w, per_receiver, dst,
                                                                                             ld_dst, dst_plane, dst_nplanes, col0);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

// two adjacent column blocks in one launch: dst[n, col0:col0+w] = src[row(n)], dst[n, col0+w:col0+w+w2] = src2[row2(n)]
extern "C" int avr_rows_broadcast2(const avr_render_geom* geom, const float* src, int32_t w, int per_receiver,
                                   const float* src2, int32_t w2, int per_receiver2, void* dst, int64_t ld_dst,
                                   int64_t dst_plane, int32_t dst_nplanes, int32_t col0, int device, void* stream) {
    AVR_REQUIRE(geom && src && src2 && dst, "null pointer");
    AVR_REQUIRE(w > 0 && w2 > 0 && col0 >= 0 && ld_dst >= col0 + w + w2, "bad column window");
    AVR_REQUIRE(dst_plane != 0 && w % 8 == 0 && w2 % 8 == 0 && col0 % 8 == 0 && ld_dst % 8 == 0 && dst_plane % 8 == 0 &&
                aligned16(dst) && aligned16(src) && aligned16(src2) && planes_kind_ok(dst_nplanes),
                "the two-block broadcast writes plane sets in 8-column groups");
    AVR_ENTER(device);
    const Geom geo = make_geom(geom);
    const int64_t total = (int64_t)geo.bs * geo.R * geo.S * (w + w2);
    if (total == 0) return AVR_OK;
    rows_broadcast_planes8_kernel<<<(unsigned)ceil_div(total / 8, 256), 256, 0, (cudaStream_t)stream>>>(
        geo, src, w, per_receiver, src2, w2, per_receiver2, (__nv_bfloat16*)dst, ld_dst, dst_plane, dst_nplanes, col0);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

extern "C" int64_t avr_rows_reduce_workspace_bytes(const avr_render_geom* geom, int32_t w, int per_receiver) {
    if (!geom) return 0;
    const Geom geo = make_geom(geom);
    int ipc;
    const int chunks = rows_reduce_plan(geo, per_receiver, &ipc);
    const int64_t rows = per_receiver ? geo.bs : geo.R;
    return rows * chunks * (int64_t)w * (int64_t)sizeof(float);
}

extern "C" int avr_rows_reduce(const avr_render_geom* geom, const void* d_dst, int64_t ld_dst, int64_t d_plane,
                               int32_t col0, int32_t w, int per_receiver, float* d_src, float* workspace,
                               int64_t workspace_bytes, int device, void* stream) {
    AVR_REQUIRE(geom && d_dst && d_src && workspace, "null pointer");
    AVR_REQUIRE(w > 0 && col0 >= 0 && ld_dst >= col0 + w, "bad column window");
    AVR_REQUIRE(workspace_bytes >= avr_rows_reduce_workspace_bytes(geom, w, per_receiver), "workspace too small");
    AVR_ENTER(device);
    const Geom geo = make_geom(geom);
    const int rows = per_receiver ? geo.bs : geo.R;
    if (rows == 0) return AVR_OK;
    int ipc;
    const int chunks = rows_reduce_plan(geo, per_receiver, &ipc);
    AVR_REQUIRE(chunks <= 65535, "too many reduce chunks");
    cudaStream_t st = (cudaStream_t)stream;
    if (d_plane != 0 && w % 8 == 0 && w <= 2048 && col0 % 8 == 0 && ld_dst % 8 == 0 && d_plane % 8 == 0 && aligned16(d_dst)) {
        const int ipp = 256 / (w / 8);
        const size_t smem = (size_t)ipp * (w + 1) * sizeof(float);          // <= 32 x 257 floats
        rows_reduce_planes8_kernel<<<dim3((unsigned)rows, (unsigned)chunks), 256, smem, st>>>(
            geo, (const __nv_bfloat16*)d_dst, ld_dst, d_plane, col0, w, per_receiver, ipc, workspace);
    } else {
        rows_reduce_kernel<<<dim3((unsigned)rows, (unsigned)chunks), 256, 0, st>>>(geo, d_dst, ld_dst, d_plane, col0, w,
                                                                                 per_receiver, ipc, workspace);
    }
    AVR_LAUNCH_CHECK();
    rows_reduce_final_kernel<<<(unsigned)ceil_div((int64_t)rows * w, 256), 256, 0, st>>>(workspace, rows, chunks, w, d_src);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}


// ---- pieces of the spectrum stage for the tensor-core DFT (the GEMM itself is avr_umma_gemm_nt) -------------
extern "C" int avr_spectrum_gain(const avr_render_geom* geom, const float* y, const float* gain, void* z, int64_t ldz,
                                 int64_t z_plane, int32_t z_nplanes, int device, void* stream) {
    AVR_REQUIRE(geom && y && gain && z, "null pointer");
    AVR_ENTER(device);
    const Geom geo = make_geom(geom);
    const int64_t rows = (int64_t)geo.bs * geo.S;
    if (rows == 0) return AVR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (z_plane == 0) {
        AVR_REQUIRE(geo.T % 4 == 0 && ldz == geo.T, "fp32 gain output must be dense with T % 4 == 0");
        gain_kernel<<<(unsigned)ceil_div(rows * (geo.T / 4), 256), 256, 0, st>>>(y, gain, rows, geo.T, geo.S, (float*)z);
    } else {
        AVR_REQUIRE(z_nplanes == 2 || z_nplanes == 3, "plane count must be 2 or 3");
        gain_planes_kernel<<<(unsigned)ceil_div(rows * geo.T, 256), 256, 0, st>>>(y, gain, rows, geo.T, geo.S, (__nv_bfloat16*)z,
                                                                               ldz, z_plane, z_nplanes);
    }
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

extern "C" int avr_spectrum_phase_sum(const avr_render_geom* geom, const float* x, int64_t ldx, const float* phase, float* out,
                                      int device, void* stream) {
    AVR_REQUIRE(geom && x && phase && out, "null pointer");
    AVR_ENTER(device);
    const Geom geo = make_geom(geom);
    const int F = geo.T / 2 + 1;
    AVR_REQUIRE(ldx >= 2 * F && ldx % 2 == 0, "ldx must be even and >= 2F");
    if (geo.bs == 0) return AVR_OK;
    phase_sum_kernel<<<(unsigned)ceil_div((int64_t)geo.bs * F, 128), 128, 0, (cudaStream_t)stream>>>(x, ldx, phase, geo.bs, geo.S, F, out);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

extern "C" int avr_spectrum_phase_bwd(const avr_render_geom* geom, const float* d_out, const float* phase, void* dx, int64_t ldx,
                                      int64_t dx_plane, int32_t dx_nplanes, int device, void* stream) {
    AVR_REQUIRE(geom && d_out && phase && dx, "null pointer");
    AVR_ENTER(device);
    const Geom geo = make_geom(geom);
    const int F = geo.T / 2 + 1;
    AVR_REQUIRE(ldx >= 2 * F && ldx % 2 == 0, "ldx must be even and >= 2F");
    const int64_t rows = (int64_t)geo.bs * geo.S;
    if (rows == 0) return AVR_OK;
    if (dx_plane) AVR_REQUIRE(dx_nplanes == 2 || dx_nplanes == 3, "plane count must be 2 or 3");
    phase_bwd_kernel<<<(unsigned)ceil_div(rows * (ldx / 2), 256), 256, 0, (cudaStream_t)stream>>>(
        d_out, phase, geo.bs, geo.S, F, ldx, dx_plane ? nullptr : (float*)dx, dx_plane ? (__nv_bfloat16*)dx : nullptr, dx_plane,
        dx_nplanes);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}


// partial[(b*R + r), c] = sum_s x[(b*R + r)*S + s, c]   (x: fp32 or bf16 plane set, first two planes)
__global__ void __launch_bounds__(256)
rows_block_sum_kernel(const Geom geo, const void* __restrict__ x_v, int64_t ldx, int64_t plane, int w, float* __restrict__ partial) {
    const int64_t blk = blockIdx.x;                        // (b, r)
    const int c = threadIdx.x * 4;
    if (c >= w) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t row0 = blk * geo.S;
    if (plane == 0) {
        const float* x = reinterpret_cast<const float*>(x_v);
        for (int s = 0; s < geo.S; ++s) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(x + (row0 + s) * ldx + c));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    } else {
        const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(x_v);
#pragma unroll 4
        for (int s = 0; s < geo.S; ++s) {
            const __nv_bfloat16* q = x + (row0 + s) * ldx + c;
            const uint2 hi = __ldg(reinterpret_cast<const uint2*>(q)), lo = __ldg(reinterpret_cast<const uint2*>(q + plane));
            acc.x += __uint_as_float(hi.x << 16) + __uint_as_float(lo.x << 16);
            acc.y += __uint_as_float(hi.x & 0xFFFF0000u) + __uint_as_float(lo.x & 0xFFFF0000u);
            acc.z += __uint_as_float(hi.y << 16) + __uint_as_float(lo.y << 16);
            acc.w += __uint_as_float(hi.y & 0xFFFF0000u) + __uint_as_float(lo.y & 0xFFFF0000u);
        }
    }
    *reinterpret_cast<float4*>(partial + blk * w + c) = acc;
}

// Sum of the S consecutive sample rows of every (receiver, ray): the adjoint of adding a per-ray / per-receiver
// row to every sample point.  partial: fp32 [bs*R, w] dense.
extern "C" int avr_rows_block_sum(const avr_render_geom* geom, const void* x, int64_t ldx, int64_t x_plane, int32_t w,
                                  float* partial, int device, void* stream) {
    AVR_REQUIRE(geom && x && partial, "null pointer");
    AVR_REQUIRE(w > 0 && w % 4 == 0 && w <= 1024 && ldx % 4 == 0, "width must be a multiple of 4 and <= 1024");
    AVR_ENTER(device);
    const Geom geo = make_geom(geom);
    const int64_t blocks = (int64_t)geo.bs * geo.R;
    if (blocks == 0 || geo.S == 0) return AVR_OK;
    const int threads = ((w / 4 + 31) / 32) * 32;
    rows_block_sum_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(geo, x, ldx, x_plane, w, partial);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}
