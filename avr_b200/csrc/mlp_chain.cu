// Fused chains of 128-wide dense layers: the sigma encoder -> sigma decoder stack of model.py:206-216 in ONE launch.
//
// Replaces what the reference runs as tiny-cuda-nn FullyFusedMLP kernels (model.py:117,146: weights in shared memory,
// activations never leaving the SM) and what round 1 of this library ran as one tcgen05 GEMM launch per layer, every
// activation making a round trip through HBM as three bf16 planes.  Here a persistent CTA owns a 128-row tile of sample
// points and walks it through all layers:
//
//   A (shared memory, 96 KB)   the tile's current activation as bf16 (hi, mid, lo) planes in the canonical K-major
//                              128-byte-swizzled layout -- exactly what TMA would have delivered -- so it is directly the
//                              A operand of the next layer's tcgen05.mma AND the source of a TMA store of the planes
//                              the backward pass needs (no staging copy, no second conversion)
//   W (shared memory, 2x48 KB) ring of 64-deep k-blocks of the layers' weight planes, streamed from L2 by TMA while the
//                              previous block's MMAs and the previous layer's epilogue run
//   D (tensor memory, 256 col) main + small-products accumulators (the two-accumulator six-product scheme of
//                              umma_gemm.cu) -- results are bit-identical to the per-layer kernel
//
//   warp 0      TMA producer (weights for every layer of every tile; the tile's input planes)
//   warp 1      TMEM allocator + single-thread MMA issuer
//   warps 2-9   epilogue: tcgen05.ld both accumulators, add, ReLU, ReLU bitmask, split into planes, st.shared into A
//               (swizzled), proxy fence, then ONE thread issues the bulk stores of the saved planes and releases the
//               next layer.  A layer whose raw output is needed too (sigma_feat feeds the signal network un-rectified,
//               model.py:219) drains its accumulator twice: raw planes -> store -> rectified planes -> store.
//
// HBM traffic per sample point (simu: 48 -> 128 x4 -> [relu] -> 128 x3 -> 16): 288 B in, 5.5 KB out (what the backward
// pass reads) against 11.9 KB for the layer-by-layer chain.
#include <cuda.h>
#include <cuda_bf16.h>
#include <atomic>
#include <stdlib.h>
#include "common.cuh"
#include "umma_ptx.cuh"

namespace avr {

constexpr int CH_MAX_LAYERS = AVR_CHAIN_MAX_LAYERS;
constexpr int CH_THREADS = 320;
constexpr uint32_t CH_A_KB = 3 * A_PLANE_BYTES;          // one 64-column k-block of the activation tile: 3 planes x 16 KB
constexpr uint32_t CH_A_BYTES = 2 * CH_A_KB;             // 128 columns
constexpr uint32_t CH_W_STAGE = 3 * 128 * 128;           // 3 planes x 128 rows x 128 B
constexpr uint32_t CH_SMEM = 1024 + CH_A_BYTES + 2 * CH_W_STAGE + 256;
constexpr uint32_t CH_TMEM_COLS = 512;                   // two accumulator sets (main + small products, 128 columns each)

struct ChainLayer {
    int K, N;                        // reduction (multiple of 16, <= 128), outputs (multiple of 16, <= 128)
    int products;                    // 6: bf16x3 . bf16x3 (24-bit operands), 3: the (hi, mid) planes only (16 bits)
    int w_planes;                    // planes of the weight set in shared memory (2 or 3)
    int relu;                        // rectify what the next layer (and `save`) sees
    int a_planes;                    // planes of the output written into the activation tile (what the next layer multiplies)
    int save_planes, raw_planes;     // planes of the (rectified / masked) output and of the raw output stored to HBM (0: none)
    __nv_bfloat16* raw; long long ld_raw, raw_plane;   // the raw output's plane set (written by the epilogue threads), or null
    uint32_t* bits; long long ldbits;        // out: bit (col % 32) of word [row][col / 32] = (raw output > 0), or null
    const uint32_t* mask; long long ldmask;  // in: the output is multiplied by this bitmask (ReLU backward), or null
    const __nv_bfloat16* accum; long long ld_acc, acc_plane; int acc_planes;   // in: planes added to the output, or null
    float* out_f32; long long ld_f32;        // fp32 output instead of planes (no activation; the chain's last layer)
    const float* bias; long long ld_bias, bias_group;   // in: fp32 row (point / bias_group) added before the activation, or null
};
struct ChainParams {
    int M, n_layers, tiles, x_planes;
    int debug;                       // -DAVR_EXPERIMENTS builds only (AVR_CHAIN_DEBUG): 1 epilogue without pack / stores to the tile,
                                     // 2 no bulk stores, 4 one product per k-step, 8 no proxy fence, 16 no tensor-memory reads,
                                     // 32 no st.shared, 64 no ReLU bitmask -- timing, results are garbage
    long long* trace;                // -DAVR_EXPERIMENTS builds only (AVR_CHAIN_TRACE_PTR): clock64() stamps of CTA 0's roles
    ChainLayer L[CH_MAX_LAYERS];
};
#ifdef AVR_EXPERIMENTS
#define CH_TRACE(role, n) do { if (p.trace && blockIdx.x == 0 && (n) < 4000) p.trace[(role) * 4096 + (n)++] = clock64(); } while (0)
#else
#define CH_TRACE(role, n) do { } while (0)
#endif
struct ChainMaps {
    CUtensorMap x0;
    CUtensorMap w[CH_MAX_LAYERS];
    CUtensorMap save[CH_MAX_LAYERS];
};


// Schedule of one 128-row tile (steady state).  Accumulators alternate between two TMEM sets from layer to layer and
// the epilogue hands the activation tile over one 64-column k-block at a time, so the next layer's first k-block runs on
// the tensor pipe while the epilogue is still converting the second half of this layer's output:
//
//   tensor pipe   ... L: kb1 (all 128 cols) |                        | L+1: kb0                  | L+1: kb1 | ...
//   epilogue      ...                       | L: cols 0-63 -> A[kb0] | L: cols 64-127 -> A[kb1]  |          | L+1: ...
//   bulk stores                                                      | saved planes of A[kb0]    | of A[kb1]
//
// Synchronisation is by mbarriers only: each epilogue warp announces (one lane) "my chunk of k-block h of the next input
// is written"; the MMA thread waits for the eight of them, issues the bulk store of the saved planes of that k-block and
// then the MMAs that read it; one tcgen05.commit per layer tells the epilogue that the accumulator is final.  The tile
// is updated in place: the commit implies that every MMA reading it has retired, and the MMA thread waits for the bulk
// stores that read it before it issues that commit.
__global__ void __launch_bounds__(CH_THREADS, 1)
mlp_chain_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t a_base = smem_u32(smem);
    const uint32_t w_base = a_base + CH_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CH_A_BYTES + 2 * CH_W_STAGE);
    const uint32_t bar_wfull = smem_u32(bars), bar_wempty = bar_wfull + 16;
    const uint32_t bar_x0full = bar_wfull + 32, bar_afree = bar_wfull + 40, bar_dfull = bar_wfull + 48, bar_aready = bar_wfull + 64;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_wfull + 8 * s, 1); mbar_init(bar_wempty + 8 * s, 1);
            mbar_init(bar_dfull + 8 * s, 1);         // accumulator columns [64 s, 64 s + 64) of the current layer are final
            mbar_init(bar_aready + 8 * s, 8);        // the eight epilogue warps are done with half s (tile k-block s written)
        }
        mbar_init(bar_x0full, 1);
        mbar_init(bar_afree, 1);             // the last layer's MMAs have retired and the saved planes have left the tile
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.x0) : "memory");
        for (int l = 0; l < p.n_layers; ++l) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w[l]) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(CH_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================== TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t wphase = 0, afree_phase = 0;
            auto load_w = [&](int l) {
                const ChainLayer& L = p.L[l];
                for (int k0 = 0; k0 < L.K; k0 += UBK) {
                    mbar_wait(bar_wempty + 8 * stage, wphase ^ 1);
                    mbar_expect_tx(bar_wfull + 8 * stage, (uint32_t)L.w_planes * (uint32_t)L.N * 128u);
                    tma_load_3d(w_base + stage * CH_W_STAGE, &maps.w[l], bar_wfull + 8 * stage, k0, 0, 0);
                    if (++stage == 2) { stage = 0; wphase ^= 1; }
                }
            };
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
                load_w(0);                                                    // first layer's weights before the tile is free
                mbar_wait(bar_afree, afree_phase ^ 1);
                afree_phase ^= 1;
                const int nkb0 = (p.L[0].K + UBK - 1) / UBK;
                mbar_expect_tx(bar_x0full, (uint32_t)(nkb0 * p.x_planes) * A_PLANE_BYTES);
                for (int kb = 0; kb < nkb0; ++kb) tma_load_3d(a_base + kb * CH_A_KB, &maps.x0, bar_x0full, kb * UBK, tile * UM, 0);
                for (int l = 1; l < p.n_layers; ++l) load_w(l);
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer (also issues the bulk stores of the saved planes)
        // The whole warp walks the schedule in lock step; ONE elected lane executes the tcgen05 / bulk-store instructions.
        // Descriptors and addresses are then warp-uniform (uniform registers, UTCHMMAs back to back); inside an
        // `if (lane == 0)` branch every MMA was wrapped in an ELECT / R2UR.BROADCAST loop of ~15 dependent instructions --
        // 89-120 clk per 128 x 128 x 16 MMA against the 64 clk it occupies the tensor pipe.
        {
            const bool elected = elect_one() != 0u;
            int stage = 0;
            uint32_t wphase = 0, x0phase = 0, ar_phase[2] = {0u, 0u}, gl = 0;
            int tn = 0; (void)tn;
            // half h of layer `l`'s output is in the tile: wait for the eight epilogue warps, then store the saved planes
            auto take_half = [&](int l, int h, int m0) {
                mbar_wait(bar_aready + 8 * h, ar_phase[h]);
                ar_phase[h] ^= 1u;
                const ChainLayer& P = p.L[l];
                int planes = (P.out_f32 || 64 * h >= P.N) ? 0 : P.save_planes;
#ifdef AVR_EXPERIMENTS
                if (p.debug & 2) planes = 0;
#endif
                if (planes && elected) { tma_store_3d(&maps.save[l], a_base + h * CH_A_KB, h * UBK, m0, 0); tma_store_commit(); }
                __syncwarp();
            };
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
                const int m0 = tile * UM;
                for (int l = 0; l < p.n_layers; ++l, ++gl) {
                    // layer constants into registers ONCE: p.L[l] is an indexed constant-bank load, and every asm volatile
                    // below is a compiler memory barrier that would otherwise re-issue it between two MMAs
                    const int LK = p.L[l].K, LN = p.L[l].N;
                    const bool six = p.L[l].products == 6;
                    const uint32_t d_main = tmem_base + 256u * (gl & 1u), d_small = d_main + 128u;
                    const uint32_t b_plane = (uint32_t)LN * 128u;
                    const int nkb = (LK + UBK - 1) / UBK;
                    for (int kb = 0; kb < nkb; ++kb) {
                        // this k-block of the layer's input: the tile load (first layer) or the previous layer's epilogue
                        // (which has then also drained the accumulator set that is about to be overwritten)
                        if (l == 0) { if (kb == 0) { mbar_wait(bar_x0full, x0phase); x0phase ^= 1; } }
                        else take_half(l - 1, kb, m0);
                        CH_TRACE(0, tn);                                         // [4 per layer: k-block ready / issued]
                        mbar_wait(bar_wfull + 8 * stage, wphase);
                        tc_fence_after();
                        const uint32_t sa = a_base + kb * CH_A_KB, sb = w_base + stage * CH_W_STAGE;
                        const int k_steps = min(UBK / 16, (LK - kb * UBK + 15) / 16);
                        const bool last_kb = kb == nkb - 1;
                        const uint64_t a0 = smem_desc(sa, 16, 1024), a1 = smem_desc(sa + A_PLANE_BYTES, 16, 1024),
                                       a2 = smem_desc(sa + 2 * A_PLANE_BYTES, 16, 1024);
                        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(LN >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);
                        const uint64_t b0 = smem_desc(sb, 16, 1024), b1 = smem_desc(sb + b_plane, 16, 1024),
                                       b2 = smem_desc(sb + 2 * b_plane, 16, 1024);
                        uint32_t acc = kb == 0 ? 0u : 1u;
                        if (elected) {
#pragma unroll
                        for (int j = 0; j < UBK / 16; ++j) {
                            if (j >= k_steps) break;
                            const uint64_t o = 2u * j;                          // 32 bytes per k16 step, in 16-byte units
#ifdef AVR_EXPERIMENTS
                            if (p.debug & 4) { umma_bf16(d_main, a0 + o, b0 + o, idesc, acc); acc = 1u; continue; }
#endif
                            if (six) {                                          // smallest products first (umma_gemm.cu, mode 1)
                                umma_bf16(d_small, a2 + o, b0 + o, idesc, acc);
                                umma_bf16(d_small, a0 + o, b2 + o, idesc, 1u);
                                umma_bf16(d_small, a1 + o, b1 + o, idesc, 1u);
                                umma_bf16(d_small, a1 + o, b0 + o, idesc, 1u);
                                umma_bf16(d_small, a0 + o, b1 + o, idesc, 1u);
                                umma_bf16(d_main, a0 + o, b0 + o, idesc, acc);
                            } else {                                            // (hi, mid) . (hi, mid): three products, one accumulator
                                umma_bf16(d_main, a1 + o, b0 + o, idesc, acc);
                                umma_bf16(d_main, a0 + o, b1 + o, idesc, 1u);
                                umma_bf16(d_main, a0 + o, b0 + o, idesc, 1u);
                            }
                            acc = 1u;
                        }
                        }
                        CH_TRACE(0, tn);
                        if (last_kb && elected) {
                            // The accumulator is final -> the epilogue may rewrite the tile.  Before that is announced the
                            // bulk stores that read it (issued when this layer started on each k-block) must be through.
                            // (Issuing the last k-block as two 64-column halves with a commit each was measured slower: a
                            // 128 x 64 x 16 MMA occupies the tensor pipe as long as a 128 x 128 x 16 one.)
                            tma_store_wait_read();
                            umma_commit(bar_dfull);
                            umma_commit(bar_dfull + 8);
                        }
                        if (elected) umma_commit(bar_wempty + 8 * stage);
                        __syncwarp();
                        if (++stage == 2) { stage = 0; wphase ^= 1; }
                    }
                }
                // the last layer's output: wait for its epilogue (the accumulators are then drained, too), store what it
                // saves, and release the tile for the next load once the MMAs have retired and the stores have read it
                take_half(p.n_layers - 1, 0, m0);
                take_half(p.n_layers - 1, 1, m0);
                if (elected) {
                    tma_store_wait_read();
                    umma_commit(bar_afree);
                }
                __syncwarp();
            }
            if (elected) tma_store_wait_all();                                 // shared memory must outlive the bulk stores
        }
    } else {
        // ===================================== epilogue: warps 2..9 -> TMEM lane groups 2,3,0,1; two warps per group,
        // each taking one 32-column chunk of the 64-column half in flight.  No block-level barrier: a warp announces its
        // chunk on the k-block's mbarrier and moves on.
        const int lane_grp = warp & 3, sub = (warp - 2) >> 2;
        const int r = lane_grp * 32 + lane;                                    // row of the tile
        const uint32_t row_off = (uint32_t)r * 128u, sw = (uint32_t)(r & 7);
        uint32_t dphase[2] = {0u, 0u}, gl = 0;
        int tn = 0; (void)tn;
        const int trole = (lane == 0 && lane_grp == 2) ? 1 + sub : -1; (void)trole;   // warps 2 and 6
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            const int m0 = tile * UM;
            const long long row = (long long)m0 + r;
            const bool row_ok = row < p.M;
            for (int l = 0; l < p.n_layers; ++l, ++gl) {
                const ChainLayer L = p.L[l];                                    // by value: registers, not re-read after every asm
                const uint32_t taddr = tmem_base + 256u * (gl & 1u) + ((uint32_t)(lane_grp * 32) << 16);
                for (int h = 0; h < 2; ++h) {
                    const int c0 = 64 * h + EPI_COLS * sub;
                    const bool active = c0 < L.N;                                // (a 16 / 48-wide last layer leaves chunks idle)
                    // operands of the epilogue that do not depend on the accumulator: fetched before the wait
                    uint32_t mword = 0xffffffffu;
                    if (L.mask && active && row_ok) mword = __ldg(L.mask + row * L.ldmask + (c0 >> 5));
                    const float* bias_row = (L.bias && active) ? L.bias + (row_ok ? row / L.bias_group : 0) * L.ld_bias + c0 : nullptr;
                    mbar_wait(bar_dfull + 8 * h, dphase[h]);
                    dphase[h] ^= 1u;
                    tc_fence_after();
                    if (trole > 0) CH_TRACE(trole, tn);                          // [4 per layer: accumulator seen / chunk handed over]
                    if (active && L.out_f32) {
                        for (int c = c0; c < c0 + EPI_COLS && c < L.N; c += 16) {
                            float a[16], b[16];
                            tmem_ld16x2(taddr + c, taddr + 128u + c, a, b);
                            if (!row_ok) continue;
                            if (L.products != 6) {
#pragma unroll
                                for (int i = 0; i < 16; ++i) b[i] = 0.f;          // one accumulator only
                            }
                            float* dst = L.out_f32 + row * L.ld_f32 + c;
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(a[4 * q] + b[4 * q], a[4 * q + 1] + b[4 * q + 1],
                                                                                      a[4 * q + 2] + b[4 * q + 2], a[4 * q + 3] + b[4 * q + 3]);
                        }
                    } else if (active) {
                        float v[32];
#ifdef AVR_EXPERIMENTS
                        if (p.debug & 16) {                                      // timing: no tensor-memory reads
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = __int_as_float(0x3f800000 + i + lane + (int)gl);
                        } else
#endif
                        if (L.products == 6) tmem_ld32_dual(taddr + c0, taddr + 128u + c0, 1.0f, v);
                        else {
                            float t0[16], t1[16];
                            tmem_ld16x2(taddr + c0, taddr + c0 + 16, t0, t1);
#pragma unroll
                            for (int i = 0; i < 16; ++i) { v[i] = t0[i]; v[16 + i] = t1[i]; }
                        }
#ifdef AVR_EXPERIMENTS
                        if (p.debug & 1) { if (v[0] == 123.456f) L.bits[0] = 1u; goto chunk_done; }
#endif
                        if (bias_row) {                                          // per-receiver embedding row (same order as the
#pragma unroll                                                                       // per-layer kernel: after the accumulators are summed)
                            for (int q = 0; q < 8; ++q) {
                                const float4 t = __ldg(reinterpret_cast<const float4*>(bias_row) + q);
                                v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
                            }
                        }
                        if (L.mask) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = (mword >> i) & 1u ? v[i] : 0.f;
                        }
                        if (L.accum && row_ok) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const long long col = c0 + 8 * q;
                                if (col >= L.N) continue;
                                const __nv_bfloat16* src = L.accum + row * L.ld_acc + col;
                                const uint4 oh = *reinterpret_cast<const uint4*>(src), om = *reinterpret_cast<const uint4*>(src + L.acc_plane);
                                uint4 ol = make_uint4(0u, 0u, 0u, 0u);
                                if (L.acc_planes == 3) ol = *reinterpret_cast<const uint4*>(src + 2 * L.acc_plane);
                                const uint32_t hw[4] = {oh.x, oh.y, oh.z, oh.w}, mw[4] = {om.x, om.y, om.z, om.w}, lw[4] = {ol.x, ol.y, ol.z, ol.w};
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    v[8 * q + 2 * i] += bf_lo(hw[i]) + (bf_lo(mw[i]) + bf_lo(lw[i]));
                                    v[8 * q + 2 * i + 1] += bf_hi(hw[i]) + (bf_hi(mw[i]) + bf_hi(lw[i]));
                                }
                            }
                        }
                        if (L.bits
#ifdef AVR_EXPERIMENTS
                            && !(p.debug & 64)
#endif
                        ) {
                            uint32_t word = 0;
#pragma unroll
                            for (int i = 0; i < 32; ++i) word |= (v[i] > 0.f ? 1u : 0u) << i;
                            if (row_ok) L.bits[row * L.ldbits + (c0 >> 5)] = word;
                        }
                        uint32_t ph[16], pm[16], pl[16];
                        if (L.raw) {
                            // the un-rectified output goes straight to global memory: 64 contiguous bytes per plane and row
                            // (whole sectors), so the tile itself only ever holds what the next layer reads
                            if (L.raw_planes == 3) pack_planes32_k<AVR_PLANES_BF16x3>(v, false, ph, pm, pl);
                            else pack_planes32_k<AVR_PLANES_BF16x2>(v, false, ph, pm, pl);
                            if (row_ok) {
                                __nv_bfloat16* dst = L.raw + row * L.ld_raw + c0;
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    if (c0 + 8 * q >= L.N) continue;
                                    *reinterpret_cast<uint4*>(dst + 8 * q) = make_uint4(ph[4 * q], ph[4 * q + 1], ph[4 * q + 2], ph[4 * q + 3]);
                                    *reinterpret_cast<uint4*>(dst + 8 * q + L.raw_plane) = make_uint4(pm[4 * q], pm[4 * q + 1], pm[4 * q + 2], pm[4 * q + 3]);
                                    if (L.raw_planes == 3)
                                        *reinterpret_cast<uint4*>(dst + 8 * q + 2 * L.raw_plane) = make_uint4(pl[4 * q], pl[4 * q + 1], pl[4 * q + 2], pl[4 * q + 3]);
                                }
                            }
                        }
                        if (L.a_planes == 3) pack_planes32_k<AVR_PLANES_BF16x3>(v, L.relu != 0, ph, pm, pl);
                        else pack_planes32_k<AVR_PLANES_BF16x2>(v, L.relu != 0, ph, pm, pl);
                        // 16-byte chunk c of row r of a k-block plane lives at chunk c ^ (r & 7)  (128-byte swizzle)
                        const uint32_t blk = a_base + (uint32_t)h * CH_A_KB + row_off;
                        const uint32_t c16 = (uint32_t)(c0 & 63) >> 3;
#ifdef AVR_EXPERIMENTS
                        if (p.debug & 32) { if (ph[0] == 0x12345678u && pm[3] == 7u && pl[5] == 9u) L.bits[0] = 1u; goto chunk_done; }   // timing: no st.shared
#endif
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint32_t off = blk + (((c16 + q) ^ sw) << 4);
                            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(off), "r"(ph[4 * q]), "r"(ph[4 * q + 1]), "r"(ph[4 * q + 2]), "r"(ph[4 * q + 3]) : "memory");
                            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(off + A_PLANE_BYTES), "r"(pm[4 * q]), "r"(pm[4 * q + 1]), "r"(pm[4 * q + 2]), "r"(pm[4 * q + 3]) : "memory");
                            if (L.a_planes == 3)
                                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(off + 2 * A_PLANE_BYTES), "r"(pl[4 * q]), "r"(pl[4 * q + 1]), "r"(pl[4 * q + 2]), "r"(pl[4 * q + 3]) : "memory");
                        }
                    }
#ifdef AVR_EXPERIMENTS
                chunk_done:
#endif
                    fence_async_smem();                                        // generic-proxy writes -> visible to UMMA and TMA
                    tc_fence_before();                                         // ... and this warp's TMEM reads are complete
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_aready + 8 * h);
                    if (trole > 0) CH_TRACE(trole, tn);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(CH_TMEM_COLS) : "memory");
    }
}

}  // namespace avr

using namespace avr;

extern "C" {

// See include/avr_b200.h ("fused chain of 128-wide dense layers").
AVR_API int avr_mlp_chain(int64_t M, const void* x0, int64_t ldx, int64_t x_plane, int32_t x_kind, int32_t k0,
                          const avr_chain_layer* layers, int32_t n_layers, int device, void* stream) {
    AVR_REQUIRE(x0 && layers, "null pointer");
    AVR_REQUIRE(n_layers >= 1 && n_layers <= CH_MAX_LAYERS, "1..8 layers");
    AVR_REQUIRE(M >= 0 && M < (1ll << 31), "bad row count");
    AVR_REQUIRE(k0 > 0 && k0 <= 128 && k0 % 16 == 0, "the chain input must be 16..128 columns wide, a multiple of 16");
    AVR_REQUIRE(x_kind == AVR_PLANES_BF16x2 || x_kind == AVR_PLANES_BF16x3, "the chain input is a bf16 pair or triple");
    AVR_ENTER(device);
    if (M == 0) return AVR_OK;
    static_assert(sizeof(ChainMaps) + sizeof(ChainParams) < 16000, "kernel parameters (CUDA 12.1+: up to 32 KB on sm_70 and later)");
    ChainMaps maps;
    ChainParams p = {};
    p.M = (int)M; p.n_layers = n_layers; p.tiles = (int)ceil_div(M, UM); p.x_planes = planes_count(x_kind);
#ifdef AVR_EXPERIMENTS
    if (const char* e = getenv("AVR_CHAIN_DEBUG")) p.debug = atoi(e);
    if (const char* e = getenv("AVR_CHAIN_TRACE_PTR")) p.trace = (long long*)strtoull(e, nullptr, 0);
#endif
    if (int rc = make_map(&maps.x0, x0, M, k0, ldx, x_plane, UM, p.x_planes)) return rc;
    int k_in = k0, in_planes = p.x_planes;
    for (int l = 0; l < n_layers; ++l) {
        const avr_chain_layer& a = layers[l];
        ChainLayer& L = p.L[l];
        const bool last = l == n_layers - 1;
        AVR_REQUIRE(a.w && a.k_in == k_in, "layer input width does not match the previous layer's output");
        AVR_REQUIRE(a.w_kind == AVR_PLANES_BF16x2 || a.w_kind == AVR_PLANES_BF16x3, "weights are bf16 pairs or triples");
        AVR_REQUIRE(a.n_out % 16 == 0 && a.n_out >= 16 && a.n_out <= 128, "layer widths must be multiples of 16 up to 128");
        AVR_REQUIRE(last || a.n_out == 128, "hidden layers of a fused chain are 128 wide");
        AVR_REQUIRE(last || !a.out_f32, "only the last layer writes fp32");
        AVR_REQUIRE(!a.out_f32 || (!a.save && !a.save_raw && !a.bits && !a.relu && !a.mask && !a.accumulate && a.ld_f32 % 4 == 0 &&
                                   aligned16(a.out_f32) && a.n_out <= 64),
                    "an fp32 output layer is at most 64 wide and has no activation, mask, planes or bitmask");
        AVR_REQUIRE(!a.bits || a.ldbits * 32 >= a.n_out, "bitmask rows too short");
        AVR_REQUIRE(!a.mask || a.ldmask * 32 >= a.n_out, "mask rows too short");
        AVR_REQUIRE(!(a.mask && a.relu) && !(a.mask && a.save_raw), "a masked (backward) layer has no activation and no raw output");
        AVR_REQUIRE(!a.bias || (!a.out_f32 && !a.mask && a.n_out == 128 && a.bias_group_rows > 0 && a.ld_bias % 4 == 0 && aligned16(a.bias)),
                    "a bias belongs to a 128-wide forward layer: 16-byte aligned fp32 rows, one per bias_group_rows points");
        L.bias = a.bias; L.ld_bias = a.ld_bias; L.bias_group = a.bias_group_rows;
        L.K = a.k_in; L.N = a.n_out; L.relu = a.relu ? 1 : 0;
        L.products = (in_planes == 3 && planes_count(a.w_kind) == 3) ? 6 : 3;   // six products need 24 bits on both sides
        L.bits = a.bits; L.ldbits = a.ldbits; L.mask = a.mask; L.ldmask = a.ldmask; L.out_f32 = a.out_f32; L.ld_f32 = a.ld_f32;
        L.w_planes = planes_count(a.w_kind);
        if (int rc = make_map(&maps.w[l], a.w, a.n_out, a.k_in, a.ldw, a.w_plane, a.n_out, L.w_planes)) return rc;
        maps.save[l] = maps.x0;
        const void* save = a.save; int64_t ld_save = a.ld_save, save_plane = a.save_plane; int save_kind = a.save_kind;
        const void* raw = a.save_raw;
        if (raw && !a.relu) {                                                   // linear layer: raw and activated outputs coincide
            AVR_REQUIRE(!save, "a linear layer has one output");
            save = raw; ld_save = a.ld_raw; save_plane = a.raw_plane; save_kind = a.raw_kind; raw = nullptr;
        }
        if (save) {
            AVR_REQUIRE(save_kind == AVR_PLANES_BF16x2 || save_kind == AVR_PLANES_BF16x3, "saved planes are bf16 pairs or triples");
            L.save_planes = planes_count(save_kind);
            if (int rc = make_map(&maps.save[l], save, M, a.n_out, ld_save, save_plane, UM, L.save_planes)) return rc;
        }
        if (a.accumulate) {
            AVR_REQUIRE(save && !raw && ld_save % 8 == 0 && save_plane % 8 == 0, "accumulate adds to the planes already in `save`");
            L.accum = (const __nv_bfloat16*)save; L.ld_acc = ld_save; L.acc_plane = save_plane; L.acc_planes = L.save_planes;
        }
        if (raw) {
            AVR_REQUIRE(a.raw_kind == AVR_PLANES_BF16x2 || a.raw_kind == AVR_PLANES_BF16x3, "saved planes are bf16 pairs or triples");
            AVR_REQUIRE(aligned16(raw) && a.ld_raw % 8 == 0 && a.raw_plane % 8 == 0, "raw planes must be 16-byte aligned");
            L.raw_planes = planes_count(a.raw_kind);
            L.raw = (__nv_bfloat16*)raw; L.ld_raw = a.ld_raw; L.raw_plane = a.raw_plane;
        }
        // planes written into the tile: what the next layer multiplies (3 when its weights carry 24 bits) and what is stored
        int need = L.save_planes > L.raw_planes ? L.save_planes : L.raw_planes;
        if (!last && planes_count(layers[l + 1].w_kind) == 3 && need < 3) need = 3;
        if (need < 2) need = 2;
        L.a_planes = need;
        AVR_REQUIRE(a.out_f32 || last || need >= 2, "internal: plane count");
        k_in = a.n_out; in_planes = L.a_planes;
    }
    static std::atomic<uint64_t> attr_done{0};
    const uint64_t bit = 1ull << (device & 63);
    if (!(attr_done.load(std::memory_order_acquire) & bit)) {
        AVR_CUDA(cudaFuncSetAttribute(mlp_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CH_SMEM));
        attr_done.fetch_or(bit, std::memory_order_release);
    }
    const int grid = p.tiles < num_sms(device) ? p.tiles : num_sms(device);
    mlp_chain_kernel<<<grid, CH_THREADS, CH_SMEM, (cudaStream_t)stream>>>(maps, p);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

}  // extern "C"
