// Ray generation + sampling + multiresolution hash-grid encode, forward and deterministic backward.
//
// Replaces: renderer.py:54-62,86-87 (ray points, normalisation, tx delays) and the tiny-cuda-nn
// `kernel_grid` / `kernel_grid_backward` launched from model.py:191,219-220.
//
// Layout: a CTA owns 128 consecutive sample points n=(b,r,s) and all levels; threadIdx.y strides the
// levels so 4 x 8 gathers are in flight per point.  Encoded rows are staged in shared memory and
// written as contiguous rows.  The backward recomputes the sample positions (nothing is saved) and
// accumulates 2^e-scaled int64 contributions with red.global.add.u64 -- integer addition is
// associative, so the result is bit-identical from run to run regardless of scheduling.
#include <cuda_bf16.h>
#include <math.h>
#include "common.cuh"

namespace avr {

struct GridDev {
    int n_levels;
    float scale[AVR_MAX_LEVELS];
    uint32_t res[AVR_MAX_LEVELS];
    uint32_t size[AVR_MAX_LEVELS];
    uint32_t offset[AVR_MAX_LEVELS];
    // 0: generic (strides below, optional hash, % size)   1: dense level (res^3 <= size)
    // 2: hashed, size = 2^k (mask)                          3: NOT hashed, size = 2^k: (x + y*s1 + z*s2) & (size - 1)
    uint32_t kind[AVR_MAX_LEVELS];
    uint32_t s1[AVR_MAX_LEVELS], s2[AVR_MAX_LEVELS];   // dense-index strides of y and z (0: dimension not reached / wrapped)
    uint32_t hashed[AVR_MAX_LEVELS];
    int pair;                          // 16-byte accesses to x-neighbour entry pairs are aligned (even offsets, aligned base)
};

// tiny-cuda-nn's grid_index: `stride = 1; for d < 3 while stride <= size: index += pos[d] * stride;
// stride *= res`, then the hash if size < stride, then % size.  tcnn keeps `stride` in a uint32: for res >= 2^16 the
// second multiplication wraps to 0, the loop runs on with stride 0 and `size < stride` is false -- the level is NOT
// hashed although res^3 > size (levels 12..14 of the 2^18 grids, 12..16 of MeshRIR's 2^20 one: index = (x + y*res) %
// size).  stride32 = 1 reproduces that (what a tcnn-trained checkpoint expects); stride32 = 0 keeps the stride exact.
// The per-level outcome (strides reached, hashed or not) is decided once here on the host.
static GridDev make_grid(const avr_grid_meta* g, const void* table_base) {
    GridDev d;
    d.n_levels = g->n_levels;
    d.pair = aligned16(table_base) ? 1 : 0;
    for (int l = 0; l < g->n_levels; ++l)
        if (g->offset[l] & 1u) d.pair = 0;
    for (int l = 0; l < AVR_MAX_LEVELS; ++l) {
        d.scale[l] = g->scale[l]; d.res[l] = g->res[l]; d.size[l] = g->size[l]; d.offset[l] = g->offset[l];
        const uint64_t res = g->res[l], size = g->size[l];
        uint64_t stride = 1, st[3] = {0, 0, 0};
        for (int dim = 0; dim < 3 && stride <= size; ++dim) {
            st[dim] = stride;
            stride *= res;                                     // <= 2^32 * 2^32: no 64-bit overflow
            if (g->stride32) stride &= 0xFFFFFFFFull;
        }
        d.s1[l] = (uint32_t)st[1]; d.s2[l] = (uint32_t)st[2];
        d.hashed[l] = size < stride ? 1u : 0u;
        d.kind[l] = 0;
        const bool pow2 = size > 0 && (size & (size - 1)) == 0;
        if (size > 0 && res > 0 && res < (1u << 20) && res * res * res <= size) d.kind[l] = 1;
        else if (pow2) d.kind[l] = d.hashed[l] ? 2 : 3;
    }
    return d;
}

constexpr int ENC_PTS = 128;   // points per CTA
constexpr int ENC_LG = 4;      // level groups (threadIdx.y)

// generic form: strides and the hashed / not hashed outcome come from make_grid
__device__ __forceinline__ uint32_t grid_index(uint32_t cx, uint32_t cy, uint32_t cz, uint32_t s1, uint32_t s2,
                                               uint32_t hashed, uint32_t size) {
    uint32_t index = cx + cy * s1 + cz * s2;                 // uint32 wrap-around, as in tcnn
    if (hashed) index = cx ^ (cy * 2654435761u) ^ (cz * 805459861u);
    return index % size;
}

// Same index, without the runtime division in the common cases (the level kind is warp-uniform): a power-of-two
// table masks; a dense level's index exceeds its table only for corners outside the grid.
template <int KIND>
__device__ __forceinline__ uint32_t grid_index_k(uint32_t cx, uint32_t cy, uint32_t cz, uint32_t s1, uint32_t s2,
                                                 uint32_t hashed, uint32_t size) {
    if (KIND == 2) return (cx ^ (cy * 2654435761u) ^ (cz * 805459861u)) & (size - 1u);
    if (KIND == 3) return (cx + cy * s1 + cz * s2) & (size - 1u);
    if (KIND == 1) {
        uint32_t index = cx + cy * s1 + cz * s2;
        if (index >= size) index %= size;
        return index;
    }
    return grid_index(cx, cy, cz, s1, s2, hashed, size);
}

// level-kind dispatch (warp-uniform)
__device__ __forceinline__ uint32_t grid_index_any(uint32_t kind, uint32_t cx, uint32_t cy, uint32_t cz, uint32_t s1,
                                                   uint32_t s2, uint32_t hashed, uint32_t size) {
    return kind == 2 ? grid_index_k<2>(cx, cy, cz, s1, s2, hashed, size)
         : kind == 1 ? grid_index_k<1>(cx, cy, cz, s1, s2, hashed, size)
         : kind == 3 ? grid_index_k<3>(cx, cy, cz, s1, s2, hashed, size)
                     : grid_index(cx, cy, cz, s1, s2, hashed, size);
}

struct Cell {
    uint32_t gx, gy, gz;
    float wx, wy, wz;
};

__device__ __forceinline__ Cell locate(float scale, float ux, float uy, float uz) {
    Cell c;
    float px = __fmaf_rn(scale, ux, 0.5f), py = __fmaf_rn(scale, uy, 0.5f), pz = __fmaf_rn(scale, uz, 0.5f);
    float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    c.gx = (uint32_t)(int)fx; c.gy = (uint32_t)(int)fy; c.gz = (uint32_t)(int)fz;
    c.wx = px - fx; c.wy = py - fy; c.wz = pz - fz;
    return c;
}

__device__ __forceinline__ float corner_weight(const Cell& c, int corner) {
    float w = (corner & 1) ? c.wx : (1.0f - c.wx);
    w = __fmul_rn(w, (corner & 2) ? c.wy : (1.0f - c.wy));
    w = __fmul_rn(w, (corner & 4) ? c.wz : (1.0f - c.wz));
    return w;
}

__device__ __forceinline__ float2 encode_level(const GridDev& g, int l, const float2* __restrict__ table,
                                               float ux, float uy, float uz) {
    const Cell c = locate(g.scale[l], ux, uy, uz);
    const uint32_t s1 = g.s1[l], s2 = g.s2[l], hashed = g.hashed[l], size = g.size[l];
    const float2* base = table + g.offset[l];
    float2 v[8];
    const uint32_t kind = g.kind[l];
    // (pairing the x-neighbour corners into one 16-byte load when they share a slot was measured: no gain, 0.25 ms)
#define AVR_GATHER(KIND)                                                                                              \
    _Pragma("unroll") for (int k = 0; k < 8; ++k)                                                                      \
        v[k] = __ldg(base + grid_index_k<KIND>(c.gx + (k & 1), c.gy + ((k >> 1) & 1), c.gz + ((k >> 2) & 1), s1, s2, hashed, size));
    if (kind == 2) { AVR_GATHER(2) } else if (kind == 1) { AVR_GATHER(1) } else if (kind == 3) { AVR_GATHER(3) } else { AVR_GATHER(0) }
#undef AVR_GATHER
    float r0 = 0.f, r1 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float w = corner_weight(c, k);
        r0 = fmaf(w, v[k].x, r0);
        r1 = fmaf(w, v[k].y, r1);
    }
    return make_float2(r0, r1);
}

// exponent e of the fixed-point scale 2^e: gmax * 2^e < 2^(62 - headroom).  ok=false -> skip.
__device__ __forceinline__ int fixed_exponent(uint32_t gmax_bits, int headroom, bool& ok, bool& poisoned) {
    float gmax = __uint_as_float(gmax_bits);
    poisoned = !(gmax <= 3.0e38f);                 // inf or nan reached the gradient
    ok = (gmax > 0.f) && !poisoned;
    int ex = 0;
    if (ok) frexpf(gmax, &ex);                     // gmax = m * 2^ex, m in [0.5, 1)
    int e = (62 - headroom) - ex;
    return max(-120, min(120, e));
}

__device__ __forceinline__ void scatter_level(const GridDev& g, int l, float ux, float uy, float uz, float g0,
                                              float g1, float sc, unsigned long long* __restrict__ acc) {
    const Cell c = locate(g.scale[l], ux, uy, uz);
    const uint32_t s1 = g.s1[l], s2 = g.s2[l], hashed = g.hashed[l], size = g.size[l];
    unsigned long long* base = acc + 2ull * g.offset[l];
    // all sixteen addends first, then sixteen back-to-back reductions from distinct registers: a RED holds its
    // source registers until the LSU has taken them, so interleaving address math with REDs serialises on that
    uint32_t idx[8];
    long long q0[8], q1[8];
    const uint32_t kind = g.kind[l];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t cx = c.gx + (k & 1), cy = c.gy + ((k >> 1) & 1), cz = c.gz + ((k >> 2) & 1);
        idx[k] = grid_index_any(kind, cx, cy, cz, s1, s2, hashed, size);
        const float w = corner_weight(c, k);
        q0[k] = __float2ll_rn(__fmul_rn(__fmul_rn(w, g0), sc));
        q1[k] = __float2ll_rn(__fmul_rn(__fmul_rn(w, g1), sc));
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        unsigned long long* dst = base + 2ull * idx[k];
        asm volatile("red.global.add.u64 [%0], %1;" ::"l"(dst), "l"(q0[k]) : "memory");
        asm volatile("red.global.add.u64 [%0], %1;" ::"l"(dst + 1), "l"(q1[k]) : "memory");
    }
}

// fp32 variant: one vector reduction per corner (both features of an entry), straight into the gradient.  The cost of
// a RED is per instruction, not per byte (measured: 2 x u64 1.57 ms, 2 x f32 1.75 ms, 1 x v2.f32 0.92 ms per simu
// step), so this halves the scatter -- at the price of a summation order that changes from run to run.
__device__ __forceinline__ void scatter_level_f32(const GridDev& g, int l, float ux, float uy, float uz, float g0,
                                                  float g1, float* __restrict__ grad) {
    const Cell c = locate(g.scale[l], ux, uy, uz);
    const uint32_t s1 = g.s1[l], s2 = g.s2[l], hashed = g.hashed[l], size = g.size[l];
    float* base = grad + 2ull * g.offset[l];
    const uint32_t kind = g.kind[l];
    uint32_t idx[8];
    float wv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t cx = c.gx + (k & 1), cy = c.gy + ((k >> 1) & 1), cz = c.gz + ((k >> 2) & 1);
        idx[k] = grid_index_any(kind, cx, cy, cz, s1, s2, hashed, size);
        wv[k] = corner_weight(c, k);
    }
    // x-neighbour corners that share a 16-byte slot (half of the pairs) go out as ONE red.global.add.v4.f32: the LSU
    // takes a reduction one lane-packet at a time, so this is a quarter fewer packets
#pragma unroll
    for (int p2 = 0; p2 < 4; ++p2) {
        const uint32_t i0 = idx[2 * p2], i1 = idx[2 * p2 + 1];
        const float a0 = __fmul_rn(wv[2 * p2], g0), a1 = __fmul_rn(wv[2 * p2], g1);
        const float b0 = __fmul_rn(wv[2 * p2 + 1], g0), b1 = __fmul_rn(wv[2 * p2 + 1], g1);
        if (g.pair && (i0 ^ i1) == 1u) {
            const bool odd = i0 & 1u;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(base + 2ull * (i0 & ~1u)), "f"(odd ? b0 : a0),
                         "f"(odd ? b1 : a1), "f"(odd ? a0 : b0), "f"(odd ? a1 : b1) : "memory");
        } else {
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(base + 2ull * i0), "f"(a0), "f"(a1) : "memory");
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(base + 2ull * i1), "f"(b0), "f"(b1) : "memory");
        }
    }
}

// Coarse levels: consecutive samples of a ray (= consecutive lanes) stay in one cell for several steps, so runs of
// adjacent lanes add to the SAME table entry.  A segmented shuffle reduction sums each run and only its first lane issues
// the RED: the LSU takes a RED one lane-packet at a time (the kernel's bound), a run of 13 lanes at level 0 costs one.
// Every lane of the warp must call this (inactive lanes pass ok = false).
__device__ __forceinline__ void scatter_level_f32_merged(const GridDev& g, int l, float ux, float uy, float uz, float g0,
                                                         float g1, bool ok, float* __restrict__ grad) {
    const Cell c = locate(g.scale[l], ux, uy, uz);
    const uint32_t s1 = g.s1[l], s2 = g.s2[l], hashed = g.hashed[l], size = g.size[l];
    float* base = grad + 2ull * g.offset[l];
    const uint32_t kind = g.kind[l];
    const int lane = (threadIdx.x + threadIdx.y * blockDim.x) & 31;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t cx = c.gx + (k & 1), cy = c.gy + ((k >> 1) & 1), cz = c.gz + ((k >> 2) & 1);
        uint32_t idx = grid_index_any(kind, cx, cy, cz, s1, s2, hashed, size);
        if (!ok) idx = 0xFFFFFFFFu;                                       // never equal to a valid neighbour's index
        const float w = corner_weight(c, k);
        float v0 = ok ? __fmul_rn(w, g0) : 0.f, v1 = ok ? __fmul_rn(w, g1) : 0.f;
        const uint32_t prev = __shfl_up_sync(0xffffffffu, idx, 1);
        const bool head = lane == 0 || prev != idx;
        const uint32_t heads = __ballot_sync(0xffffffffu, head);
        const uint32_t after = lane == 31 ? 0xffffffffu : (heads >> (lane + 1));   // bit j: lane + 1 + j starts a new run
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const float o0 = __shfl_down_sync(0xffffffffu, v0, d), o1 = __shfl_down_sync(0xffffffffu, v1, d);
            if (lane + d < 32 && (after & ((1u << d) - 1u)) == 0u) { v0 += o0; v1 += o1; }
        }
        if (head && ok)
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(base + 2ull * idx), "f"(v0), "f"(v1) : "memory");
    }
}

// the same for the deterministic mode: 2^e-scaled int64 addends (integer sums are exact, so merging keeps the result
// bit-identical to the unmerged scatter)
__device__ __forceinline__ void scatter_level_i64_merged(const GridDev& g, int l, float ux, float uy, float uz, float g0,
                                                         float g1, float sc, bool ok, unsigned long long* __restrict__ acc) {
    const Cell c = locate(g.scale[l], ux, uy, uz);
    const uint32_t s1 = g.s1[l], s2 = g.s2[l], hashed = g.hashed[l], size = g.size[l];
    unsigned long long* base = acc + 2ull * g.offset[l];
    const uint32_t kind = g.kind[l];
    const int lane = (threadIdx.x + threadIdx.y * blockDim.x) & 31;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t cx = c.gx + (k & 1), cy = c.gy + ((k >> 1) & 1), cz = c.gz + ((k >> 2) & 1);
        uint32_t idx = grid_index_any(kind, cx, cy, cz, s1, s2, hashed, size);
        if (!ok) idx = 0xFFFFFFFFu;
        const float w = corner_weight(c, k);
        long long q0 = ok ? __float2ll_rn(__fmul_rn(__fmul_rn(w, g0), sc)) : 0ll;
        long long q1 = ok ? __float2ll_rn(__fmul_rn(__fmul_rn(w, g1), sc)) : 0ll;
        const uint32_t prev = __shfl_up_sync(0xffffffffu, idx, 1);
        const bool head = lane == 0 || prev != idx;
        const uint32_t heads = __ballot_sync(0xffffffffu, head);
        const uint32_t after = lane == 31 ? 0xffffffffu : (heads >> (lane + 1));
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long o0 = __shfl_down_sync(0xffffffffu, q0, d), o1 = __shfl_down_sync(0xffffffffu, q1, d);
            if (lane + d < 32 && (after & ((1u << d) - 1u)) == 0u) { q0 += o0; q1 += o1; }
        }
        if (head && ok) {
            unsigned long long* dst = base + 2ull * idx;
            asm volatile("red.global.add.u64 [%0], %1;" ::"l"(dst), "l"(q0) : "memory");
            asm volatile("red.global.add.u64 [%0], %1;" ::"l"(dst + 1), "l"(q1) : "memory");
        }
    }
}

// unit-cube position of sample n=(b,r,s); optionally its tx delay
__device__ __forceinline__ void sample_unit(const Geom& geo, int64_t n, const float* __restrict__ rays_o,
                                            const float* __restrict__ dirs, const float* __restrict__ d_vals,
                                            float& ux, float& uy, float& uz, float& nx, float& ny, float& nz, int& b) {
    const int64_t P = (int64_t)geo.R * geo.S;
    b = (int)(n / P);
    const int rem = (int)(n - (int64_t)b * P);
    const int r = rem / geo.S, s = rem - r * geo.S;
    const float d = __ldg(d_vals + s);
    nx = to_unit(ray_point(__ldg(rays_o + 3 * b + 0), __ldg(dirs + 3 * r + 0), d), geo.lo, geo.span);
    ny = to_unit(ray_point(__ldg(rays_o + 3 * b + 1), __ldg(dirs + 3 * r + 1), d), geo.lo, geo.span);
    nz = to_unit(ray_point(__ldg(rays_o + 3 * b + 2), __ldg(dirs + 3 * r + 2), d), geo.lo, geo.span);
    ux = to_cube(nx); uy = to_cube(ny); uz = to_cube(nz);
}

template <bool RAYGEN>
__global__ void __launch_bounds__(ENC_PTS* ENC_LG)
encode_fwd_kernel(const Geom geo, const __grid_constant__ GridDev grid, int64_t n_pts, const float* __restrict__ rays_o,
                  const float* __restrict__ pos_tx, const float* __restrict__ dirs, const float* __restrict__ d_vals,
                  const float* __restrict__ u_in, const float2* __restrict__ table, void* __restrict__ out_v,
                  int64_t ld_out, int64_t out_plane, int out_np, int col0, int n_ones, int* __restrict__ delay) {
    extern __shared__ float tile[];
    const int W = 2 * grid.n_levels + n_ones;
    const int Wp = W | 1;                                   // odd stride: conflict-free column writes
    const int p = threadIdx.x, lg = threadIdx.y;
    const int64_t n = (int64_t)blockIdx.x * ENC_PTS + p;
    if (n < n_pts) {
        float ux, uy, uz;
        if (RAYGEN) {
            float nx, ny, nz;
            int b;
            sample_unit(geo, n, rays_o, dirs, d_vals, ux, uy, uz, nx, ny, nz, b);
            if (lg == 0 && delay != nullptr) {
                float tx0 = to_unit(__ldg(pos_tx + 3 * b + 0), geo.lo, geo.span);
                float tx1 = to_unit(__ldg(pos_tx + 3 * b + 1), geo.lo, geo.span);
                float tx2 = to_unit(__ldg(pos_tx + 3 * b + 2), geo.lo, geo.span);
                delay[n] = source_delay(tx0, tx1, tx2, nx, ny, nz, geo);
            }
        } else {
            ux = __ldg(u_in + 3 * n + 0); uy = __ldg(u_in + 3 * n + 1); uz = __ldg(u_in + 3 * n + 2);
        }
        for (int l = lg; l < grid.n_levels; l += ENC_LG) {
            float2 v = encode_level(grid, l, table, ux, uy, uz);
            tile[p * Wp + 2 * l] = v.x;
            tile[p * Wp + 2 * l + 1] = v.y;
        }
        if (lg == 0)
            for (int c = 0; c < n_ones; ++c) tile[p * Wp + 2 * grid.n_levels + c] = 1.0f;
    }
    __syncthreads();
    const int tid = lg * ENC_PTS + p;
    const int64_t n0 = (int64_t)blockIdx.x * ENC_PTS;
    for (int idx = tid; idx < ENC_PTS * W; idx += ENC_PTS * ENC_LG) {
        const int pp = idx / W, c = idx - pp * W;
        if (n0 + pp >= n_pts) continue;
        const float v = tile[pp * Wp + c];
        const int64_t o = (n0 + pp) * ld_out + col0 + c;
        if (out_plane == 0) reinterpret_cast<float*>(out_v)[o] = v;
        else planes_store(out_v, o, out_plane, out_np, v);  // tensor-core operand: plane set of kind out_np
    }
}

template <bool RAYGEN, bool F32ACC>
__global__ void __launch_bounds__(ENC_PTS* ENC_LG)
encode_bwd_kernel(const Geom geo, const __grid_constant__ GridDev grid, int64_t n_pts, const float* __restrict__ rays_o,
                  const float* __restrict__ dirs, const float* __restrict__ d_vals, const float* __restrict__ u_in,
                  const void* __restrict__ d_out_v, int64_t ld_out, int64_t d_plane, int col0,
                  const uint32_t* __restrict__ gmax_bits, int headroom, unsigned long long* __restrict__ acc,
                  int merge_levels) {
    extern __shared__ float tile[];
    float sc = 1.f;
    if (!F32ACC) {
        bool ok, poisoned;
        const int e = fixed_exponent(__ldg(gmax_bits), headroom, ok, poisoned);
        if (!ok) return;                                    // zero (or poisoned) gradient: nothing to add
        sc = ldexpf(1.0f, e);
    }
    const int W = 2 * grid.n_levels;
    const int Wp = W | 1;
    const int p = threadIdx.x, lg = threadIdx.y;
    const int tid = lg * ENC_PTS + p;
    const int64_t n0 = (int64_t)blockIdx.x * ENC_PTS;
    for (int idx = tid; idx < ENC_PTS * W; idx += ENC_PTS * ENC_LG) {
        const int pp = idx / W, c = idx - pp * W;
        float v = 0.f;
        if (n0 + pp < n_pts) {
            const int64_t o = (n0 + pp) * ld_out + col0 + c;
            if (d_plane == 0) v = __ldg(reinterpret_cast<const float*>(d_out_v) + o);
            else {
                const __nv_bfloat16* db = reinterpret_cast<const __nv_bfloat16*>(d_out_v);
                v = __bfloat162float(db[o]) + __bfloat162float(db[o + d_plane]);
            }
        }
        tile[pp * Wp + c] = v;
    }
    __syncthreads();
    const int64_t n = n0 + p;
    const bool active = n < n_pts;
    if (!active && merge_levels == 0) return;                 // the merged levels need whole warps
    float ux = 0.f, uy = 0.f, uz = 0.f;
    if (active) {
        if (RAYGEN) {
            float nx, ny, nz;
            int b;
            sample_unit(geo, n, rays_o, dirs, d_vals, ux, uy, uz, nx, ny, nz, b);
        } else {
            ux = __ldg(u_in + 3 * n + 0); uy = __ldg(u_in + 3 * n + 1); uz = __ldg(u_in + 3 * n + 2);
        }
    }
    for (int l = lg; l < grid.n_levels; l += ENC_LG) {
        const float g0 = active ? tile[p * Wp + 2 * l] : 0.f, g1 = active ? tile[p * Wp + 2 * l + 1] : 0.f;
        const bool ok = active && !(g0 == 0.f && g1 == 0.f);
        if (l < merge_levels) {                               // warp-uniform: l depends on threadIdx.y only
            if (F32ACC) scatter_level_f32_merged(grid, l, ux, uy, uz, g0, g1, ok, reinterpret_cast<float*>(acc));
            else scatter_level_i64_merged(grid, l, ux, uy, uz, g0, g1, sc, ok, acc);
            continue;
        }
        if (!ok) continue;
        if (F32ACC) scatter_level_f32(grid, l, ux, uy, uz, g0, g1, reinterpret_cast<float*>(acc));
        else scatter_level(grid, l, ux, uy, uz, g0, g1, sc, acc);
    }
}

__global__ void absmax_kernel(const void* __restrict__ x_v, int64_t rows, int64_t ld, int64_t plane, int col0, int ncols,
                              uint32_t* __restrict__ gmax_bits) {
    uint32_t m = 0;
    const int64_t total = rows * ncols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / ncols;
        const int c = (int)(i - r * ncols);
        float v;
        if (plane == 0) v = __ldg(reinterpret_cast<const float*>(x_v) + r * ld + col0 + c);
        else {
            const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x_v);
            v = __bfloat162float(xb[r * ld + col0 + c]) + __bfloat162float(xb[r * ld + col0 + c + plane]);
        }
        m = max(m, __float_as_uint(fabsf(v)));   // |x| bit patterns order like values; NaN sorts last
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m != 0) atomicMax(gmax_bits, m);
}

// the same for (hi, mid) plane pairs whose window is made of whole 16-byte groups: one thread per 8 columns of a row,
// two 16-byte loads (the element-wise kernel above spends most of its time on index arithmetic: 95 -> ~25 us at simu)
__global__ void absmax_planes8_kernel(const __nv_bfloat16* __restrict__ x, int64_t rows, int64_t ld, int64_t plane, int col0,
                                      int groups, uint32_t* __restrict__ gmax_bits) {
    uint32_t m = 0;
    const int64_t total = rows * groups;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / groups;
        const int g = (int)(i - r * groups);
        const __nv_bfloat16* q = x + r * ld + col0 + 8 * g;
        const uint4 h = __ldg(reinterpret_cast<const uint4*>(q)), l = __ldg(reinterpret_cast<const uint4*>(q + plane));
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float a = __uint_as_float(hw[k] << 16) + __uint_as_float(lw[k] << 16);
            const float b = __uint_as_float(hw[k] & 0xFFFF0000u) + __uint_as_float(lw[k] & 0xFFFF0000u);
            m = max(m, max(__float_as_uint(fabsf(a)), __float_as_uint(fabsf(b))));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m != 0) atomicMax(gmax_bits, m);
}

__global__ void grad_finalize_kernel(const long long* __restrict__ acc, int64_t n, const uint32_t* __restrict__ gmax_bits,
                                     int headroom, float* __restrict__ grad, int accumulate) {
    bool ok, poisoned;
    const int e = fixed_exponent(__ldg(gmax_bits), headroom, ok, poisoned);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float v = 0.f;
        if (poisoned) v = __uint_as_float(0x7fc00000u);
        else if (ok) v = (float)ldexp((double)acc[i], -e);
        grad[i] = accumulate ? grad[i] + v : v;
    }
}

__global__ void sample_points_kernel(const Geom geo, const float* __restrict__ rays_o, const float* __restrict__ pos_tx,
                                     const float* __restrict__ dirs, const float* __restrict__ d_vals,
                                     float* __restrict__ pts_n, float* __restrict__ view, float* __restrict__ tx_n,
                                     int* __restrict__ delay) {
    const int64_t n_pts = (int64_t)geo.bs * geo.R * geo.S;
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_pts) return;
    float ux, uy, uz, nx, ny, nz;
    int b;
    sample_unit(geo, n, rays_o, dirs, d_vals, ux, uy, uz, nx, ny, nz, b);
    const float tx0 = to_unit(__ldg(pos_tx + 3 * b + 0), geo.lo, geo.span);
    const float tx1 = to_unit(__ldg(pos_tx + 3 * b + 1), geo.lo, geo.span);
    const float tx2 = to_unit(__ldg(pos_tx + 3 * b + 2), geo.lo, geo.span);
    if (pts_n) { pts_n[3 * n] = nx; pts_n[3 * n + 1] = ny; pts_n[3 * n + 2] = nz; }
    if (tx_n) { tx_n[3 * n] = tx0; tx_n[3 * n + 1] = tx1; tx_n[3 * n + 2] = tx2; }
    if (view) {
        const int r = (int)((n / geo.S) % geo.R);
        view[3 * n] = -__ldg(dirs + 3 * r); view[3 * n + 1] = -__ldg(dirs + 3 * r + 1); view[3 * n + 2] = -__ldg(dirs + 3 * r + 2);
    }
    if (delay) delay[n] = source_delay(tx0, tx1, tx2, nx, ny, nz, geo);
}

// unit-cube inputs of the per-ray / per-receiver encodings: (-dir+1)/2 and (normalize(x)+1)/2
__global__ void aux_inputs_kernel(const Geom geo, const float* __restrict__ pos_tx, const float* __restrict__ dirs,
                                  const float* __restrict__ dir_tx, float* __restrict__ u_view, float* __restrict__ u_tx,
                                  float* __restrict__ u_dir_tx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (u_view && i < geo.R * 3) u_view[i] = to_cube(-__ldg(dirs + i));
    if (u_tx && i < geo.bs * 3) u_tx[i] = to_cube(to_unit(__ldg(pos_tx + i), geo.lo, geo.span));
    if (u_dir_tx && dir_tx && i < geo.bs * 3) u_dir_tx[i] = to_cube(__ldg(dir_tx + i));
}

static int check_grid(const avr_grid_meta* g) {
    if (!g) return fail(AVR_ERR_INVALID, "grid meta is null");
    if (g->n_feat != 2) return fail(AVR_ERR_UNSUPPORTED, "n_features_per_level=%d (only 2 is built)", g->n_feat);
    if (g->n_levels < 1 || g->n_levels > AVR_MAX_LEVELS) return fail(AVR_ERR_INVALID, "n_levels=%d out of range", g->n_levels);
    return AVR_OK;
}

}  // namespace avr

using namespace avr;

extern "C" int avr_sample_points(const avr_render_geom* geom, const float* rays_o, const float* pos_tx,
                                 const float* dirs, const float* d_vals, float* pts_n, float* view, float* tx_n,
                                 int32_t* delay, int device, void* stream) {
    AVR_REQUIRE(geom, "null geometry");
    if ((int64_t)geom->bs * geom->R * geom->S == 0) return AVR_OK;
    AVR_REQUIRE(rays_o && pos_tx && dirs && d_vals, "null input");
    AVR_ENTER(device);
    const Geom geo = make_geom(geom);
    const int64_t n = (int64_t)geo.bs * geo.R * geo.S;
    if (n == 0) return AVR_OK;
    sample_points_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(geo, rays_o, pos_tx, dirs, d_vals,
                                                                                      pts_n, view, tx_n, delay);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

extern "C" int avr_aux_inputs(const avr_render_geom* geom, const float* pos_tx, const float* dirs, const float* dir_tx,
                              float* u_view, float* u_tx, float* u_dir_tx, int device, void* stream) {
    AVR_REQUIRE(geom && pos_tx && dirs, "null input");
    AVR_ENTER(device);
    const Geom geo = make_geom(geom);
    const int n = 3 * (geo.R > geo.bs ? geo.R : geo.bs);
    if (n == 0) return AVR_OK;
    aux_inputs_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(geo, pos_tx, dirs, dir_tx, u_view, u_tx, u_dir_tx);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

static int encode_fwd_common(bool raygen, const Geom& geo, const avr_grid_meta* grid, int64_t n_pts, const float* rays_o,
                             const float* pos_tx, const float* dirs, const float* d_vals, const float* u,
                             const float* table, void* out, int64_t ld_out, int64_t out_plane, int out_np,
                             int32_t col0, int32_t n_ones, int32_t* delay, void* stream) {
    if (int rc = check_grid(grid)) return rc;
    AVR_REQUIRE(table && out, "null table/out");
    AVR_REQUIRE(out_plane == 0 || planes_kind_ok(out_np), "unknown plane-set kind");
    AVR_REQUIRE(n_ones >= 0 && col0 >= 0 && ld_out >= col0 + 2 * grid->n_levels + n_ones, "bad output window");
    AVR_REQUIRE((reinterpret_cast<uintptr_t>(table) & 7u) == 0, "table must be 8-byte aligned");
    if (n_pts == 0) return AVR_OK;
    const GridDev gd = make_grid(grid, table);
    const int W = 2 * grid->n_levels + n_ones;
    const size_t smem = (size_t)ENC_PTS * (W | 1) * sizeof(float);
    const dim3 block(ENC_PTS, ENC_LG);
    const unsigned blocks = (unsigned)ceil_div(n_pts, ENC_PTS);
    if (raygen) {
        AVR_CUDA(cudaFuncSetAttribute(encode_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        encode_fwd_kernel<true><<<blocks, block, smem, (cudaStream_t)stream>>>(
            geo, gd, n_pts, rays_o, pos_tx, dirs, d_vals, nullptr, (const float2*)table, out, ld_out, out_plane, out_np, col0,
            n_ones, delay);
    } else {
        AVR_CUDA(cudaFuncSetAttribute(encode_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        encode_fwd_kernel<false><<<blocks, block, smem, (cudaStream_t)stream>>>(
            geo, gd, n_pts, nullptr, nullptr, nullptr, nullptr, u, (const float2*)table, out, ld_out, out_plane, out_np, col0,
            n_ones, nullptr);
    }
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

extern "C" int avr_raygen_encode_fwd(const avr_render_geom* geom, const avr_grid_meta* grid, const float* rays_o,
                                     const float* pos_tx, const float* dirs, const float* d_vals, const float* table,
                                     void* out, int64_t ld_out, int64_t out_plane, int32_t out_nplanes, int32_t col0,
                                     int32_t n_ones, int32_t* delay, int device, void* stream) {
    AVR_REQUIRE(geom && rays_o && dirs && d_vals, "null input");
    AVR_REQUIRE(delay == nullptr || pos_tx != nullptr, "delay requested without pos_tx");
    AVR_ENTER(device);
    const Geom geo = make_geom(geom);
    return encode_fwd_common(true, geo, grid, (int64_t)geo.bs * geo.R * geo.S, rays_o, pos_tx, dirs, d_vals, nullptr,
                             table, out, ld_out, out_plane, out_nplanes, col0, n_ones, delay, stream);
}

extern "C" int avr_grid_encode_fwd(const avr_grid_meta* grid, const float* u, int64_t n_pts, const float* table,
                                   void* out, int64_t ld_out, int64_t out_plane, int32_t out_nplanes, int32_t col0,
                                   int32_t n_ones, int device, void* stream) {
    AVR_REQUIRE(u != nullptr || n_pts == 0, "null input");
    AVR_ENTER(device);
    Geom geo = {};
    return encode_fwd_common(false, geo, grid, n_pts, nullptr, nullptr, nullptr, nullptr, u, table, out, ld_out, out_plane,
                             out_nplanes, col0, n_ones, nullptr, stream);
}

static int encode_bwd_common(bool raygen, const Geom& geo, const avr_grid_meta* grid, int64_t n_pts, const float* rays_o,
                             const float* dirs, const float* d_vals, const float* u, const void* d_out, int64_t ld_out,
                             int64_t d_plane, int32_t col0, const uint32_t* gmax_bits, int32_t headroom, void* acc,
                             float sample_step, void* stream) {
    if (int rc = check_grid(grid)) return rc;
    const bool f32acc = headroom == AVR_GRID_GRAD_F32;
    AVR_REQUIRE(d_out && acc && (gmax_bits || f32acc), "null d_out/gmax/acc");
    AVR_REQUIRE(col0 >= 0 && ld_out >= col0 + 2 * grid->n_levels, "bad gradient window");
    AVR_REQUIRE(f32acc || (headroom >= 0 && headroom <= 56), "log2_headroom out of range");
    if (n_pts == 0) return AVR_OK;
    const GridDev gd = make_grid(grid, f32acc ? acc : nullptr);          // paired 16-byte reductions: fp32 mode only
    const int W = 2 * grid->n_levels;
    const size_t smem = (size_t)ENC_PTS * (W | 1) * sizeof(float);
    const dim3 block(ENC_PTS, ENC_LG);
    const unsigned blocks = (unsigned)ceil_div(n_pts, ENC_PTS);
    // Leading levels whose cells hold >= 2 consecutive samples of a ray on average get the run-merging scatter
    // (ray-generation mode: consecutive points are consecutive samples, `sample_step` apart in unit-cube coordinates).
    int merge_levels = 0;
    if (raygen && sample_step > 0.f) {
        for (int l = 0; l < grid->n_levels && grid->scale[l] * sample_step < 0.5f; ++l) merge_levels = l + 1;
#ifdef AVR_EXPERIMENTS
        if (const char* e = getenv("AVR_SCATTER_MERGE_LEVELS")) merge_levels = atoi(e);      // A/B measurements
#endif
        if (merge_levels > grid->n_levels) merge_levels = grid->n_levels;
        if (merge_levels < 0) merge_levels = 0;
    }
    auto launch = [&](auto kernel) -> int {
        AVR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kernel<<<blocks, block, smem, (cudaStream_t)stream>>>(geo, gd, n_pts, rays_o, dirs, d_vals, u, d_out, ld_out, d_plane,
                                                             col0, gmax_bits, headroom, (unsigned long long*)acc, merge_levels);
        return AVR_OK;
    };
    int rc;
    if (raygen) rc = f32acc ? launch(encode_bwd_kernel<true, true>) : launch(encode_bwd_kernel<true, false>);
    else rc = f32acc ? launch(encode_bwd_kernel<false, true>) : launch(encode_bwd_kernel<false, false>);
    if (rc) return rc;
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

extern "C" int avr_raygen_encode_bwd(const avr_render_geom* geom, const avr_grid_meta* grid, const float* rays_o,
                                     const float* dirs, const float* d_vals, const void* d_out, int64_t ld_out,
                                     int64_t d_plane, int32_t col0, const uint32_t* gmax_bits, int32_t log2_headroom,
                                     void* acc, float sample_step, int device, void* stream) {
    AVR_REQUIRE(geom && rays_o && dirs && d_vals, "null input");
    AVR_ENTER(device);
    const Geom geo = make_geom(geom);
    return encode_bwd_common(true, geo, grid, (int64_t)geo.bs * geo.R * geo.S, rays_o, dirs, d_vals, nullptr, d_out,
                             ld_out, d_plane, col0, gmax_bits, log2_headroom, acc, sample_step, stream);
}

extern "C" int avr_grid_encode_bwd(const avr_grid_meta* grid, const float* u, int64_t n_pts, const void* d_out,
                                   int64_t ld_out, int64_t d_plane, int32_t col0, const uint32_t* gmax_bits,
                                   int32_t log2_headroom, void* acc, int device, void* stream) {
    AVR_REQUIRE(u != nullptr || n_pts == 0, "null input");
    AVR_ENTER(device);
    Geom geo = {};
    return encode_bwd_common(false, geo, grid, n_pts, nullptr, nullptr, nullptr, u, d_out, ld_out, d_plane, col0,
                             gmax_bits, log2_headroom, acc, 0.f, stream);
}

extern "C" int avr_absmax_bits(const void* x, int64_t rows, int64_t ld, int64_t plane, int32_t col0, int32_t ncols,
                               uint32_t* gmax_bits, int device, void* stream) {
    AVR_REQUIRE(gmax_bits, "null gmax_bits");
    AVR_REQUIRE(x != nullptr || rows == 0, "null input");
    AVR_ENTER(device);
    if (rows * ncols == 0) return AVR_OK;
    if (plane != 0 && col0 % 8 == 0 && ncols % 8 == 0 && ld % 8 == 0 && plane % 8 == 0 && aligned16(x)) {
        int64_t blocks = ceil_div(rows * (ncols / 8), 256 * 2);
        if (blocks > 148 * 16) blocks = 148 * 16;
        absmax_planes8_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, rows, ld, plane, col0,
                                                                                 ncols / 8, gmax_bits);
        AVR_LAUNCH_CHECK();
        return AVR_OK;
    }
    int64_t blocks = ceil_div(rows * ncols, 256 * 8);
    if (blocks > 148 * 16) blocks = 148 * 16;
    absmax_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, ld, plane, col0, ncols, gmax_bits);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

extern "C" int avr_grid_grad_finalize(const int64_t* acc, int64_t n, const uint32_t* gmax_bits, int32_t log2_headroom,
                                      float* grad, int accumulate, int device, void* stream) {
    AVR_REQUIRE(acc && gmax_bits && grad, "null pointer");
    AVR_ENTER(device);
    if (n == 0) return AVR_OK;
    int64_t blocks = ceil_div(n, 256 * 4);
    if (blocks > 148 * 16) blocks = 148 * 16;
    grad_finalize_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const long long*)acc, n, gmax_bits,
                                                                            log2_headroom, grad, accumulate);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}
