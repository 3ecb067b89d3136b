// Output layer of the signal network fused with the time-domain ray reduction ("collapse").
//
// Reference: signal = H @ W_out^T (model.py:231, the width -> T layer, 215 GFLOP / receiver at simu), then
// renderer.py:86-90,115-118 masks every row with (t >= delay) and sums the rays with the compositing weights:
//
//      y[b,s,t] = sum_r w[b,r,s] * [t >= delay[b,r,s]] * ( H[b,r,s,:] . W_out[t,:] )
//
// Sorting the rays of one cell (b,s) by delay makes the masked sum a PREFIX sum:
//
//      G[b,s,t,:] = sum_{r : delay <= t} w * H[b,r,s,:]          y[b,s,t] = G[b,s,t,:] . W_out[t,:]
//
// i.e. S*T*width MACs per receiver instead of R*S*T*width (R = 2050 x fewer), and the [bs,R,S,T] signal
// tensor (0.84 GB / receiver) and its gradient never exist.  Exact in real arithmetic; in fp32 only the
// summation order differs from the reference.  G only changes for t in [dmin, dmax) of the cell -- at most
// `tspan` <= 2*far*fs/speed steps -- so its snapshots fit in a small workspace, which the backward reuses:
//      d_W_out[t,:] = sum_cells d_y[t] * G[t,:]
//      g[d,:]       = sum_{t >= d} d_y[t] * W_out[t,:]      (suffix sum)      d_H = w * g[delay],  d_w = H . g[delay]
//
// Kernels
//   delay_sort      stable counting sort of the rays of every cell by delay (one warp per cell)
//   prefix_walk     one CTA per cell walks its sorted rays; activation rows (bf16 hi/mid planes) are prefetched
//                   32 rays deep with cp.async; every thread owns 4 hidden columns; snapshots of G are written
//                   whenever t advances (stores only: no load latency on the critical path)
//   prefix_dot      y = G . W_out rows, one warp per (cell, t)
//   suffix_walk     g snapshots for d in [dmin, dmax], thread per 4 columns, W_out rows streamed with unrolled loads
//   ray_backward    parallel over the sorted rays: d_H planes and d_w
//   dwout_reduce    d_W_out from the saved G snapshots, fixed summation order over the cells
#include <cuda_bf16.h>
#include "common.cuh"

namespace avr {

struct Planes {
    const __nv_bfloat16* p;     // 16-bit planes: bf16 (hi, mid[, lo]) or an fp16 (hi, lo'*2^11) pair
    long long ld, plane;
    bool f16;
};

// four values from planes 0 and 1 (the first 16 bits of a bf16 set, all 24 of an fp16 pair)
__device__ __forceinline__ void unpack4(bool f16, uint2 hi, uint2 lo, float (&x)[4]) {
    const float2 a = planes_unpack2(f16, hi.x, lo.x), b = planes_unpack2(f16, hi.y, lo.y);
    x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y;
}
__device__ __forceinline__ uint32_t pack_bf2(float a, float b, uint32_t& lo_out) {
    const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
    const __nv_bfloat16 al = __float2bfloat16_rn(a - __bfloat162float(ah)), bl = __float2bfloat16_rn(b - __bfloat162float(bh));
    lo_out = (uint32_t)__bfloat16_as_ushort(al) | ((uint32_t)__bfloat16_as_ushort(bl) << 16);
    return (uint32_t)__bfloat16_as_ushort(ah) | ((uint32_t)__bfloat16_as_ushort(bh) << 16);
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// workspace kept from forward to backward.  The sorted ray list of a cell is walked in PW_SEG independent segments:
//   gp    [cells][tspan][width]   slot t - dmin: the prefix sum WITHIN the segment that owns t (rays of that segment with
//                                 delay <= t); segment k owns t in [first delay of segment k, first delay of segment k+1)
//   cum   [cells][PW_SEG][width]  sum of all rays of the segments before k  ->  G(t) = cum[k(t)] + gp[t - dmin]
//   gtot  [cells][width]          sum of all rays (G(t) for t >= dmax), NaN-poisoned when the cell's delay spread > tspan
//   segfirst [cells][PW_SEG], range [cells][2] = (dmin, dmax);  segtot [cells][PW_SEG][width] = per-segment totals
constexpr int PW_SEG = 4;
struct PrefixWs {
    float* gp;
    float* cum;
    float* gtot;
    float* segtot;
    int* segfirst;
    int* range;
};
__host__ __device__ inline PrefixWs carve_prefix(void* ws, long long cells, int tspan, int width) {
    PrefixWs w;
    w.gp = reinterpret_cast<float*>(ws);
    w.cum = w.gp + cells * tspan * width;
    w.gtot = w.cum + cells * PW_SEG * width;
    w.segtot = w.gtot + cells * width;
    w.segfirst = reinterpret_cast<int*>(w.segtot + cells * PW_SEG * width);
    w.range = w.segfirst + cells * PW_SEG;
    return w;
}
// segment of the cell that owns time step t (segfirst is non-decreasing; empty segments carry dmax)
__device__ __forceinline__ int owner_segment(const int* __restrict__ sf, int t) {
    int k = 0;
#pragma unroll
    for (int j = 1; j < PW_SEG; ++j) k += (__ldg(sf + j) <= t) ? 1 : 0;
    return k;
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

// ------------------------------------------------------------------------------------------------
// prefix walk: G snapshots of one cell
// ------------------------------------------------------------------------------------------------
constexpr int PW_GROUP = 4;        // rays per cp.async group
constexpr int PW_NGROUPS = 4;      // groups in flight  -> 16 rays deep (seven CTAs per SM: the segments supply the parallelism)
constexpr int PW_MAX_THREADS = 256;

// grid (cells, PW_SEG): segment k of the cell's delay-sorted ray list
__global__ void __launch_bounds__(PW_MAX_THREADS)
prefix_walk_kernel(const Geom geo, const Planes act, int width, const int* __restrict__ order,
                   const int* __restrict__ sdelay, const float* __restrict__ sw, PrefixWs ws, int tspan) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int cell = blockIdx.x, seg = blockIdx.y, b = cell / geo.S, s = cell - b * geo.S;
    const int R = geo.R;
    const int L = (R + PW_SEG - 1) / PW_SEG;
    const int k_beg = min(R, seg * L), k_end = min(R, k_beg + L), n_here = k_end - k_beg;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int c = tid * 4;
    const bool col_ok = c < width;
    int* l_ord = reinterpret_cast<int*>(smem);
    int* l_del = l_ord + L;
    float* l_w = reinterpret_cast<float*>(l_del + L);
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(smem + (((size_t)L * 12 + 15) & ~(size_t)15));
    const long long lbase = (long long)cell * R;
    for (int k = tid; k < n_here; k += nthr) {
        l_ord[k] = __ldg(order + lbase + k_beg + k);
        l_del[k] = __ldg(sdelay + lbase + k_beg + k);
        l_w[k] = __ldg(sw + lbase + k_beg + k);
    }
    __syncthreads();
    const int dmin = __ldg(sdelay + lbase), dmax = __ldg(sdelay + lbase + R - 1);
    const bool overflow = (dmax - dmin) > tspan;
    // this segment owns the time steps from its first delay up to the next segment's first delay
    const int t_first = n_here > 0 ? l_del[0] : dmax;
    const int t_next = k_end < R ? __ldg(sdelay + lbase + k_end) : dmax;
    const int n_groups = (n_here + PW_GROUP - 1) / PW_GROUP;

    auto issue = [&](int g) {
        if (g < n_groups && col_ok) {
#pragma unroll
            for (int j = 0; j < PW_GROUP; ++j) {
                const int k = g * PW_GROUP + j;
                if (k < n_here) {
                    const long long row = ((long long)b * R + l_ord[k]) * geo.S + s;
                    const __nv_bfloat16* q = act.p + row * act.ld + c;
                    const uint32_t dst = ring + (uint32_t)((((g % PW_NGROUPS) * PW_GROUP + j) * nthr + tid) * 16);
                    cp_async8(dst, q);
                    cp_async8(dst + 8, q + act.plane);
                }
            }
        }
        cp_async_commit();
    };
    for (int g = 0; g < PW_NGROUPS; ++g) issue(g);

    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int t = t_first;
    float* gp_cell = ws.gp + (long long)cell * tspan * width;
    for (int g = 0; g < n_groups; ++g) {
        cp_async_wait<PW_NGROUPS - 1>();
#pragma unroll
        for (int j = 0; j < PW_GROUP; ++j) {
            const int k = g * PW_GROUP + j;
            if (k >= n_here) break;
            const int d = l_del[k];
            while (t < d) {                                    // snapshot for every t the prefix is constant over
                if (col_ok && !overflow)
                    *reinterpret_cast<float4*>(gp_cell + (long long)(t - dmin) * width + c) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                ++t;
            }
            if (col_ok) {
                const uint32_t src = ring + (uint32_t)((((g % PW_NGROUPS) * PW_GROUP + j) * nthr + tid) * 16);
                uint2 hi, lo;
                asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(hi.x), "=r"(hi.y) : "r"(src));
                asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(lo.x), "=r"(lo.y) : "r"(src + 8));
                float x[4];
                unpack4(act.f16, hi, lo, x);
                const float wk = l_w[k];
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] = fmaf(wk, x[i], acc[i]);
            }
        }
        issue(g + PW_NGROUPS);
    }
    cp_async_wait<0>();
    while (t < t_next) {                                       // ... up to the next segment's first delay: the segment total
        if (col_ok && !overflow)
            *reinterpret_cast<float4*>(gp_cell + (long long)(t - dmin) * width + c) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        ++t;
    }
    if (col_ok)
        *reinterpret_cast<float4*>(ws.segtot + ((long long)cell * PW_SEG + seg) * width + c) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    if (tid == 0) {
        ws.segfirst[cell * PW_SEG + seg] = t_first;
        if (seg == 0) { ws.range[2 * cell] = dmin; ws.range[2 * cell + 1] = dmax; }
    }
}

// cum[cell][k] = sum of the totals of the segments before k (fixed order); gtot = the sum of all, NaN when the cell's
// delay spread exceeds tspan (its snapshots do not exist: consumers read the poisoned total instead)
__global__ void prefix_offsets_kernel(const Geom geo, int width, PrefixWs ws, int tspan) {
    const int cell = blockIdx.x;
    const int c = threadIdx.x * 4;
    if (c >= width) return;
    const bool overflow = (ws.range[2 * cell + 1] - ws.range[2 * cell]) > tspan;
    float4 run = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < PW_SEG; ++k) {
        *reinterpret_cast<float4*>(ws.cum + ((long long)cell * PW_SEG + k) * width + c) = run;
        const float4 t = *reinterpret_cast<const float4*>(ws.segtot + ((long long)cell * PW_SEG + k) * width + c);
        run.x += t.x; run.y += t.y; run.z += t.z; run.w += t.w;
    }
    const float poison = overflow ? __uint_as_float(0x7fc00000u) : 0.f;
    *reinterpret_cast<float4*>(ws.gtot + (long long)cell * width + c) = make_float4(run.x + poison, run.y + poison, run.z + poison, run.w + poison);
}

// y[cell, t] = G(cell, t) . W_out[t, :]      grid (cells, t-chunks), one warp per t
__global__ void __launch_bounds__(128)
prefix_dot_kernel(const Geom geo, int width, PrefixWs ws, int tspan, const float* __restrict__ w_out, long long ldw,
                  float* __restrict__ y, int t_per_block) {
    const int cell = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int dmin = ws.range[2 * cell], dmax = ws.range[2 * cell + 1];
    const int t_beg = blockIdx.y * t_per_block, t_end = min(geo.T, t_beg + t_per_block);
    const float* gp_cell = ws.gp + (long long)cell * tspan * width;
    const float* gt = ws.gtot + (long long)cell * width;
    const int* sf = ws.segfirst + cell * PW_SEG;
    for (int t = t_beg + warp; t < t_end; t += n_warps) {
        float p = 0.f;
        if (t >= dmin) {
            // a cell whose delay spread exceeds tspan has no snapshots (they would lie outside its slice of the
            // workspace): it reads the poisoned total instead -> NaN, never an out-of-bounds access
            const bool snap = t < dmax && dmax - dmin <= tspan;
            const float* g = snap ? gp_cell + (long long)(t - dmin) * width : gt;
            const float* off = snap ? ws.cum + ((long long)cell * PW_SEG + owner_segment(sf, t)) * width : nullptr;
            const float* wr = w_out + (long long)t * ldw;
            for (int c = lane * 4; c < width; c += 128) {
                float4 a = *reinterpret_cast<const float4*>(g + c);
                if (off) { const float4 o = *reinterpret_cast<const float4*>(off + c); a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w; }
                const float4 wv = __ldg(reinterpret_cast<const float4*>(wr + c));
                p = fmaf(a.x, wv.x, p); p = fmaf(a.y, wv.y, p); p = fmaf(a.z, wv.z, p); p = fmaf(a.w, wv.w, p);
            }
        }
        p = warp_sum(p);
        if (lane == 0) y[(long long)cell * geo.T + t] = p;
    }
}

// d_W_out[t, c] (+)= sum_cells d_y[cell, t] * G(cell, t)[c]      (fixed order over the cells)
// block = (32 float4 columns of one t) x DW_GROUPS cell groups: every group sums its share of the cells (four times the
// loads in flight of one thread walking all cells), the groups are combined through shared memory in a fixed order
constexpr int DW_GROUPS = 4;
__global__ void __launch_bounds__(32 * DW_GROUPS)
dwout_reduce_kernel(const Geom geo, int width, const float* __restrict__ d_y, PrefixWs ws, int tspan,
                    float* __restrict__ d_wout, long long ldw, int accumulate) {
    __shared__ float4 part[DW_GROUPS][32];
    const int wq = width / 4;
    const int q = blockIdx.x * 32 + threadIdx.x;
    const bool ok = q < geo.T * wq;
    const int t = ok ? q / wq : 0, c = ok ? (q - t * wq) * 4 : 0;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const int n = geo.bs * geo.S;
    const int per = (n + DW_GROUPS - 1) / DW_GROUPS;
    const int i_beg = threadIdx.y * per, i_end = min(n, i_beg + per);
    if (ok) {
        for (int i = i_beg; i < i_end; ++i) {
            const int dmin = __ldg(ws.range + 2 * i), dmax = __ldg(ws.range + 2 * i + 1);
            if (t < dmin) continue;
            const float g = __ldg(d_y + (long long)i * geo.T + t);
            const bool snap = t < dmax && dmax - dmin <= tspan;
            const float* src = snap ? ws.gp + ((long long)i * tspan + (t - dmin)) * width + c : ws.gtot + (long long)i * width + c;
            float4 a = *reinterpret_cast<const float4*>(src);
            if (snap) {
                const float4 o = *reinterpret_cast<const float4*>(ws.cum + ((long long)i * PW_SEG + owner_segment(ws.segfirst + i * PW_SEG, t)) * width + c);
                a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
            }
            v.x = fmaf(g, a.x, v.x); v.y = fmaf(g, a.y, v.y); v.z = fmaf(g, a.z, v.z); v.w = fmaf(g, a.w, v.w);
        }
    }
    part[threadIdx.y][threadIdx.x] = v;
    __syncthreads();
    if (threadIdx.y != 0 || !ok) return;
#pragma unroll
    for (int k = 1; k < DW_GROUPS; ++k) {
        const float4 o = part[k][threadIdx.x];
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
    }
    float* dst = d_wout + (long long)t * ldw + c;
    if (accumulate) {
        const float4 o = *reinterpret_cast<const float4*>(dst);
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
    }
    *reinterpret_cast<float4*>(dst) = v;
}

// g snapshots: gs[cell][d - dmin][c] = sum_{t >= d} d_y[cell, t] * W_out[t, c],  d in [dmin, dmax]
// grid (cells, column groups of 128), 32 threads x 4 columns
__global__ void __launch_bounds__(32)
suffix_walk_kernel(const Geom geo, int width, const int* __restrict__ range, int tspan, const float* __restrict__ w_out,
                   long long ldw, const float* __restrict__ d_y, float* __restrict__ gs) {
    extern __shared__ float dy_s[];
    const int cell = blockIdx.x;
    const int c = (blockIdx.y * 32 + threadIdx.x) * 4;
    const int T = geo.T;
    for (int i = threadIdx.x; i < T; i += 32) dy_s[i] = __ldg(d_y + (long long)cell * T + i);
    __syncwarp();
    if (c >= width) return;
    const int dmin = range[2 * cell], dmax = range[2 * cell + 1];
    if (dmax - dmin > tspan) return;                           // poisoned in the forward pass already
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    const float* wc = w_out + c;
    int t = T - 1;
#pragma unroll 8
    for (; t > dmax; --t) {
        const float gy = dy_s[t];
        const float4 wv = __ldg(reinterpret_cast<const float4*>(wc + (long long)t * ldw));
        g[0] = fmaf(gy, wv.x, g[0]); g[1] = fmaf(gy, wv.y, g[1]); g[2] = fmaf(gy, wv.z, g[2]); g[3] = fmaf(gy, wv.w, g[3]);
    }
    float* gs_cell = gs + (long long)cell * (tspan + 1) * width + c;
#pragma unroll 4
    for (; t >= dmin; --t) {
        const float gy = dy_s[t];
        const float4 wv = __ldg(reinterpret_cast<const float4*>(wc + (long long)t * ldw));
        g[0] = fmaf(gy, wv.x, g[0]); g[1] = fmaf(gy, wv.y, g[1]); g[2] = fmaf(gy, wv.z, g[2]); g[3] = fmaf(gy, wv.w, g[3]);
        *reinterpret_cast<float4*>(gs_cell + (long long)(t - dmin) * width) = make_float4(g[0], g[1], g[2], g[3]);
    }
}

// d_H[row,:] = (H > 0) * w * g[delay],  d_w[row] = H[row,:] . g[delay]; one warp per 32 consecutive sorted rays
constexpr int RB_RAYS = 32;
__global__ void __launch_bounds__(128)
ray_backward_kernel(const Geom geo, const Planes act, int width, const int* __restrict__ order,
                    const int* __restrict__ sdelay, const float* __restrict__ sw, const int* __restrict__ range, int tspan,
                    const float* __restrict__ gs, __nv_bfloat16* __restrict__ d_act, long long ld_d, long long d_plane,
                    float* __restrict__ d_w) {
    const int cell = blockIdx.x, b = cell / geo.S, s = cell - b * geo.S;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k0 = (blockIdx.y * (blockDim.x >> 5) + warp) * RB_RAYS;
    const int R = geo.R;
    if (k0 >= R) return;
    const int dmin = range[2 * cell];
    const bool poisoned = (range[2 * cell + 1] - dmin) > tspan;
    const long long lbase = (long long)cell * R;
    const int my_k = k0 + lane;
    const int my_ord = my_k < R ? __ldg(order + lbase + my_k) : 0;
    const int my_del = my_k < R ? __ldg(sdelay + lbase + my_k) : 0;
    const float my_w = my_k < R ? __ldg(sw + lbase + my_k) : 0.f;
    const float* gs_cell = gs + (long long)cell * (tspan + 1) * width;
    const int n_here = min(RB_RAYS, R - k0);
    for (int c0 = 0; c0 < width; c0 += 128) {                  // 128 columns per pass: 4 per lane
        const int c = c0 + lane * 4;
        const bool col_ok = c < width;
        int d_prev = -1;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < n_here; ++j) {
            const int r = __shfl_sync(0xffffffffu, my_ord, j);
            const int d = __shfl_sync(0xffffffffu, my_del, j);
            const float wk = __shfl_sync(0xffffffffu, my_w, j);
            if (d != d_prev && col_ok) {
                g = poisoned ? make_float4(__uint_as_float(0x7fc00000u), 0.f, 0.f, 0.f)
                             : *reinterpret_cast<const float4*>(gs_cell + (long long)(d - dmin) * width + c);
                d_prev = d;
            }
            const long long row = ((long long)b * R + r) * geo.S + s;
            float dot = 0.f;
            if (col_ok) {
                const __nv_bfloat16* q = act.p + row * act.ld + c;
                const uint2 hi = __ldg(reinterpret_cast<const uint2*>(q)), lo = __ldg(reinterpret_cast<const uint2*>(q + act.plane));
                float x[4];
                unpack4(act.f16, hi, lo, x);
                dot = x[0] * g.x + x[1] * g.y + x[2] * g.z + x[3] * g.w;
                uint32_t l0, l1;
                const uint32_t h0 = pack_bf2(x[0] > 0.f ? wk * g.x : 0.f, x[1] > 0.f ? wk * g.y : 0.f, l0);
                const uint32_t h1 = pack_bf2(x[2] > 0.f ? wk * g.z : 0.f, x[3] > 0.f ? wk * g.w : 0.f, l1);
                *reinterpret_cast<uint2*>(d_act + row * ld_d + c) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(d_act + row * ld_d + c + d_plane) = make_uint2(l0, l1);
            }
            dot = warp_sum(dot);
            if (lane == 0) {
                if (c0 == 0) d_w[row] = dot;
                else d_w[row] += dot;                          // same thread, fixed order: deterministic
            }
        }
    }
}

static int check_collapse(const Geom& geo, int width, const void* act, long long ld, long long plane) {
    if (width % 8 != 0 || width <= 0 || width > 4 * PW_MAX_THREADS)
        return fail(AVR_ERR_UNSUPPORTED, "collapse: hidden width %d must be a multiple of 8 and <= %d", width, 4 * PW_MAX_THREADS);
    if (geo.R < 1 || geo.T < 1) return fail(AVR_ERR_INVALID, "collapse: empty geometry");
    if (!act || ld % 4 != 0 || plane % 4 != 0 || (reinterpret_cast<uintptr_t>(act) & 7u))
        return fail(AVR_ERR_INVALID, "collapse: activation planes must be 8-byte aligned");
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

AVR_API int64_t avr_collapse_prefix_bytes(const avr_render_geom* geom, int32_t width, int32_t tspan) {
    if (!geom) return 0;
    const int64_t n = (int64_t)geom->bs * geom->S;
    return (n * tspan * width + n * width + 2 * n * PW_SEG * width) * (int64_t)sizeof(float) + n * (2 + PW_SEG) * (int64_t)sizeof(int) + 64;
}

AVR_API int64_t avr_collapse_suffix_bytes(const avr_render_geom* geom, int32_t width, int32_t tspan) {
    if (!geom) return 0;
    return (int64_t)geom->bs * geom->S * (tspan + 1) * width * (int64_t)sizeof(float) + 64;
}

// y[bs,S,T] and the prefix workspace (kept by the caller for the backward pass).
// tspan: static bound on (max delay - min delay) within one (b,s); a violation poisons the results with NaN.
AVR_API int avr_collapse_fwd(const avr_render_geom* geom, const void* act_planes, int64_t ld_act, int64_t act_plane,
                             int32_t act_kind, int32_t width, const int32_t* order, const int32_t* sdelay, const float* sw,
                             const float* w_out, int64_t ldw, int32_t tspan, void* prefix_ws, int64_t prefix_bytes, float* y,
                             int device, void* stream) {
    AVR_REQUIRE(geom && order && sdelay && sw && w_out && y && prefix_ws, "null pointer");
    AVR_ENTER(device);
    const Geom geo = make_geom(geom);
    if (int rc = check_collapse(geo, width, act_planes, ld_act, act_plane)) return rc;
    AVR_REQUIRE(planes_kind_ok(act_kind), "unknown plane-set kind");
    AVR_REQUIRE(ldw % 4 == 0 && aligned16(w_out) && aligned16(prefix_ws), "W_out / workspace must be 16-byte aligned");
    AVR_REQUIRE(tspan > 0 && prefix_bytes >= avr_collapse_prefix_bytes(geom, width, tspan), "prefix workspace too small");
    const int cells = geo.bs * geo.S;
    if (cells == 0) return AVR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const PrefixWs ws = carve_prefix(prefix_ws, cells, tspan, width);
    const int threads = ((width / 4 + 31) / 32) * 32;
    const size_t seg_len = (size_t)(geo.R + PW_SEG - 1) / PW_SEG;
    const size_t smem = ((seg_len * 12 + 15) & ~(size_t)15) + (size_t)PW_GROUP * PW_NGROUPS * threads * 16;
    AVR_CUDA(cudaFuncSetAttribute(prefix_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const Planes act = {(const __nv_bfloat16*)act_planes, ld_act, act_plane, planes_f16(act_kind)};
    prefix_walk_kernel<<<dim3(cells, PW_SEG), threads, smem, st>>>(geo, act, width, order, sdelay, sw, ws, tspan);
    AVR_LAUNCH_CHECK();
    prefix_offsets_kernel<<<cells, threads, 0, st>>>(geo, width, ws, tspan);
    AVR_LAUNCH_CHECK();
    const int t_chunks = 8;
    const int t_per_block = (int)ceil_div(geo.T, t_chunks);
    prefix_dot_kernel<<<dim3(cells, t_chunks), 128, 0, st>>>(geo, width, ws, tspan, w_out, ldw, y, t_per_block);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

// d_act (plane pair, gradient w.r.t. the pre-activation), d_w[bs,R,S] and d_W_out[T,width] (+)=
AVR_API int avr_collapse_bwd(const avr_render_geom* geom, const void* act_planes, int64_t ld_act, int64_t act_plane,
                             int32_t act_kind, int32_t width, const int32_t* order, const int32_t* sdelay, const float* sw,
                             const float* w_out, int64_t ldw, const float* d_y, int32_t tspan, const void* prefix_ws,
                             void* suffix_ws, int64_t suffix_bytes, void* d_act_planes, int64_t ld_d, int64_t d_plane,
                             float* d_w, float* d_wout, int64_t ld_dw, int accumulate, int device, void* stream) {
    AVR_REQUIRE(geom && order && sdelay && sw && w_out && d_y && prefix_ws && suffix_ws && d_act_planes && d_w && d_wout,
                "null pointer");
    AVR_ENTER(device);
    const Geom geo = make_geom(geom);
    if (int rc = check_collapse(geo, width, act_planes, ld_act, act_plane)) return rc;
    AVR_REQUIRE(planes_kind_ok(act_kind), "unknown plane-set kind");
    AVR_REQUIRE(ldw % 4 == 0 && ld_dw % 4 == 0 && aligned16(w_out) && aligned16(d_wout) && aligned16(suffix_ws), "16-byte alignment");
    AVR_REQUIRE(ld_d % 4 == 0 && d_plane % 4 == 0 && (reinterpret_cast<uintptr_t>(d_act_planes) & 7u) == 0, "d_act misaligned");
    AVR_REQUIRE(tspan > 0 && suffix_bytes >= avr_collapse_suffix_bytes(geom, width, tspan), "suffix workspace too small");
    const int cells = geo.bs * geo.S;
    if (cells == 0) return AVR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const PrefixWs ws = carve_prefix(const_cast<void*>(prefix_ws), cells, tspan, width);
    float* gs = (float*)suffix_ws;
    const int total = geo.T * (width / 4);
    dwout_reduce_kernel<<<(total + 31) / 32, dim3(32, DW_GROUPS), 0, st>>>(geo, width, d_y, ws, tspan, d_wout, ld_dw, accumulate);
    AVR_LAUNCH_CHECK();
    const size_t smem = (size_t)geo.T * sizeof(float);
    AVR_CUDA(cudaFuncSetAttribute(suffix_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    suffix_walk_kernel<<<dim3(cells, (unsigned)ceil_div(width, 128)), 32, smem, st>>>(geo, width, ws.range, tspan, w_out, ldw, d_y, gs);
    AVR_LAUNCH_CHECK();
    const Planes act = {(const __nv_bfloat16*)act_planes, ld_act, act_plane, planes_f16(act_kind)};
    const int rays_per_block = 4 * RB_RAYS;
    ray_backward_kernel<<<dim3(cells, (unsigned)ceil_div(geo.R, rays_per_block)), 128, 0, st>>>(
        geo, act, width, order, sdelay, sw, ws.range, tspan, gs, (__nv_bfloat16*)d_act_planes, ld_d, d_plane, d_w);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

}  // extern "C"
