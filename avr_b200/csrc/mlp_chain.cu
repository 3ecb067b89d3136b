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
#include "common.cuh"
#include "umma_ptx.cuh"

namespace avr {

constexpr int CH_MAX_LAYERS = AVR_CHAIN_MAX_LAYERS;
constexpr int CH_THREADS = 320;
constexpr uint32_t CH_A_KB = 3 * A_PLANE_BYTES;          // one 64-column k-block of the activation tile: 3 planes x 16 KB
constexpr uint32_t CH_A_BYTES = 2 * CH_A_KB;             // 128 columns
constexpr uint32_t CH_W_STAGE = 3 * 128 * 128;           // 3 planes x 128 rows x 128 B
constexpr uint32_t CH_SMEM = 1024 + CH_A_BYTES + 2 * CH_W_STAGE + 256;

struct ChainLayer {
    int K, N;                        // reduction (multiple of 16, <= 128), outputs (multiple of 16, <= 128)
    int relu;                        // rectify what the next layer (and `save`) sees
    int save_planes, raw_planes;     // planes of the (rectified) output / of the raw output stored to HBM (0: none)
    uint32_t* bits; long long ldbits;    // bit (col % 32) of word [row][col / 32] = (raw output > 0), or null
    float* out_f32; long long ld_f32;    // fp32 output instead of planes (no activation; the chain's last layer)
};
struct ChainParams {
    int M, n_layers, tiles;
    ChainLayer L[CH_MAX_LAYERS];
};
struct ChainMaps {
    CUtensorMap x0;
    CUtensorMap w[CH_MAX_LAYERS];
    CUtensorMap save[CH_MAX_LAYERS];
    CUtensorMap raw[CH_MAX_LAYERS];
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(CH_THREADS, 1)
mlp_chain_fwd_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t a_base = smem_u32(smem);
    const uint32_t w_base = a_base + CH_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CH_A_BYTES + 2 * CH_W_STAGE);
    const uint32_t bar_wfull = smem_u32(bars), bar_wempty = bar_wfull + 16;
    const uint32_t bar_x0full = bar_wfull + 32, bar_afree = bar_wfull + 40, bar_dfull = bar_wfull + 48, bar_aready = bar_wfull + 56;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(bar_wfull + 8 * s, 1); mbar_init(bar_wempty + 8 * s, 1); }
        mbar_init(bar_x0full, 1);
        mbar_init(bar_afree, 2);             // the last layer's MMAs have retired + the saved planes have left the tile
        mbar_init(bar_dfull, 1);
        mbar_init(bar_aready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.x0) : "memory");
        for (int l = 0; l < p.n_layers; ++l) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w[l]) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
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
                    mbar_expect_tx(bar_wfull + 8 * stage, 3u * (uint32_t)L.N * 128u);
                    tma_load_3d(w_base + stage * CH_W_STAGE, &maps.w[l], bar_wfull + 8 * stage, k0, 0, 0);
                    if (++stage == 2) { stage = 0; wphase ^= 1; }
                }
            };
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
                load_w(0);                                                    // first layer's weights before the tile is free
                mbar_wait(bar_afree, afree_phase ^ 1);
                afree_phase ^= 1;
                const int nkb0 = (p.L[0].K + UBK - 1) / UBK;
                mbar_expect_tx(bar_x0full, (uint32_t)nkb0 * CH_A_KB);
                for (int kb = 0; kb < nkb0; ++kb) tma_load_3d(a_base + kb * CH_A_KB, &maps.x0, bar_x0full, kb * UBK, tile * UM, 0);
                for (int l = 1; l < p.n_layers; ++l) load_w(l);
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer
        if (lane == 0) {
            int stage = 0;
            uint32_t wphase = 0, x0phase = 0, ar_phase = 0;
            const uint32_t d_main = tmem_base, d_small = tmem_base + 128u;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
                for (int l = 0; l < p.n_layers; ++l) {
                    const ChainLayer& L = p.L[l];
                    // the previous layer's epilogue has drained D and written this layer's A (passes at once the first time)
                    mbar_wait(bar_aready, ar_phase ^ 1);
                    ar_phase ^= 1;
                    if (l == 0) { mbar_wait(bar_x0full, x0phase); x0phase ^= 1; }
                    tc_fence_after();
                    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(L.N >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);
                    const uint32_t b_plane = (uint32_t)L.N * 128u;
                    uint32_t acc = 0u;
                    int kb = 0;
                    for (int k0 = 0; k0 < L.K; k0 += UBK, ++kb) {
                        mbar_wait(bar_wfull + 8 * stage, wphase);
                        tc_fence_after();
                        const uint32_t sa = a_base + kb * CH_A_KB, sb = w_base + stage * CH_W_STAGE;
                        const int k_steps = min(UBK / 16, (L.K - k0 + 15) / 16);
                        const uint64_t a0 = smem_desc(sa, 16, 1024), a1 = smem_desc(sa + A_PLANE_BYTES, 16, 1024),
                                       a2 = smem_desc(sa + 2 * A_PLANE_BYTES, 16, 1024);
                        const uint64_t b0 = smem_desc(sb, 16, 1024), b1 = smem_desc(sb + b_plane, 16, 1024),
                                       b2 = smem_desc(sb + 2 * b_plane, 16, 1024);
#pragma unroll
                        for (int j = 0; j < UBK / 16; ++j) {                    // six products, smallest first (umma_gemm.cu, mode 1)
                            if (j >= k_steps) break;
                            const uint64_t o = 2u * j;                          // 32 bytes per k16 step, in 16-byte units
                            umma_bf16(d_small, a2 + o, b0 + o, idesc, acc);
                            umma_bf16(d_small, a0 + o, b2 + o, idesc, 1u);
                            umma_bf16(d_small, a1 + o, b1 + o, idesc, 1u);
                            umma_bf16(d_small, a1 + o, b0 + o, idesc, 1u);
                            umma_bf16(d_small, a0 + o, b1 + o, idesc, 1u);
                            umma_bf16(d_main, a0 + o, b0 + o, idesc, acc);
                            acc = 1u;
                        }
                        umma_commit(bar_wempty + 8 * stage);
                        if (++stage == 2) { stage = 0; wphase ^= 1; }
                    }
                    umma_commit(bar_dfull);
                    if (l == p.n_layers - 1) umma_commit(bar_afree);
                }
            }
        }
    } else {
        // ===================================== epilogue: warps 2..9 -> TMEM lane groups 2,3,0,1; two warps per group
        const int lane_grp = warp & 3, half = (warp - 2) >> 2;
        const bool leader = threadIdx.x == 64;
        const int r = lane_grp * 32 + lane;                                    // row of the tile
        const uint32_t row_off = (uint32_t)r * 128u, sw = (uint32_t)(r & 7);
        const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16);
        uint32_t dphase = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            const int m0 = tile * UM;
            const long long row = (long long)m0 + r;
            const bool row_ok = row < p.M;
            for (int l = 0; l < p.n_layers; ++l) {
                const ChainLayer& L = p.L[l];
                mbar_wait(bar_dfull, dphase);
                dphase ^= 1;
                tc_fence_after();
                if (L.out_f32) {
                    if (half == 0) {
                        for (int c0 = 0; c0 < L.N; c0 += 16) {
                            float a[16], b[16];
                            tmem_ld16x2(taddr + c0, taddr + 128u + c0, a, b);
                            if (!row_ok) continue;
                            float* dst = L.out_f32 + row * L.ld_f32 + c0;
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(a[4 * q] + b[4 * q], a[4 * q + 1] + b[4 * q + 1],
                                                                                      a[4 * q + 2] + b[4 * q + 2], a[4 * q + 3] + b[4 * q + 3]);
                        }
                    }
                } else {
                    const int passes = L.raw_planes ? 2 : 1;
                    for (int pass = 0; pass < passes; ++pass) {
                        const bool raw_pass = L.raw_planes && pass == 0;
                        if (leader) tma_store_wait_read();                     // earlier bulk stores have finished reading the tile
                        epi_bar_sync();
                        for (int c0 = EPI_COLS * half; c0 < L.N; c0 += 2 * EPI_COLS) {
                            float v[32];
                            tmem_ld32_dual(taddr + c0, taddr + 128u + c0, 1.0f, v);
                            if (pass == 0 && L.bits) {
                                uint32_t word = 0;
#pragma unroll
                                for (int i = 0; i < 32; ++i) word |= (v[i] > 0.f ? 1u : 0u) << i;
                                if (row_ok) L.bits[row * L.ldbits + (c0 >> 5)] = word;
                            }
                            uint32_t ph[16], pm[16], pl[16];
                            pack_planes32_k<AVR_PLANES_BF16x3>(v, L.relu && !raw_pass, ph, pm, pl);
                            // 16-byte chunk c of row r of a k-block plane lives at chunk c ^ (r & 7)  (128-byte swizzle)
                            const uint32_t blk = a_base + (uint32_t)(c0 >> 6) * CH_A_KB + row_off;
                            const uint32_t c16 = (uint32_t)(c0 & 63) >> 3;
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint32_t off = blk + (((c16 + q) ^ sw) << 4);
                                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(off), "r"(ph[4 * q]), "r"(ph[4 * q + 1]), "r"(ph[4 * q + 2]), "r"(ph[4 * q + 3]) : "memory");
                                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(off + A_PLANE_BYTES), "r"(pm[4 * q]), "r"(pm[4 * q + 1]), "r"(pm[4 * q + 2]), "r"(pm[4 * q + 3]) : "memory");
                                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(off + 2 * A_PLANE_BYTES), "r"(pl[4 * q]), "r"(pl[4 * q + 1]), "r"(pl[4 * q + 2]), "r"(pl[4 * q + 3]) : "memory");
                            }
                        }
                        fence_async_smem();                                    // generic-proxy writes -> visible to UMMA and TMA
                        tc_fence_before();
                        epi_bar_sync();
                        if (leader) {
                            const int planes = raw_pass ? L.raw_planes : L.save_planes;
                            if (planes) {
                                const CUtensorMap* map = raw_pass ? &maps.raw[l] : &maps.save[l];
                                for (int kb = 0; kb * UBK < L.N; ++kb) tma_store_3d(map, a_base + kb * CH_A_KB, kb * UBK, m0, 0);
                                tma_store_commit();
                            }
                        }
                    }
                }
                if (l == p.n_layers - 1) {
                    // the tile may be refilled once nothing reads it any more: the MMA side commits on its own, here the
                    // bulk stores of the saved planes
                    if (L.out_f32) { tc_fence_before(); epi_bar_sync(); }
                    if (leader) { tma_store_wait_read(); mbar_arrive(bar_afree); mbar_arrive(bar_aready); }
                } else {
                    if (L.out_f32) { tc_fence_before(); epi_bar_sync(); }
                    if (leader) mbar_arrive(bar_aready);
                }
            }
        }
        if (leader) tma_store_wait_all();                                      // shared memory must outlive the bulk stores
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
    }
}

}  // namespace avr

using namespace avr;

extern "C" {

// x0: bf16 (hi, mid, lo) planes [3][M][ldx] of the chain's input (k0 columns); every layer's weights: bf16x3 planes of
// W[n_out, k_in] (K-major).  See include/avr_b200.h.
AVR_API int avr_mlp_chain_fwd(int64_t M, const void* x0, int64_t ldx, int64_t x_plane, int32_t k0,
                              const avr_chain_layer* layers, int32_t n_layers, int device, void* stream) {
    AVR_REQUIRE(x0 && layers, "null pointer");
    AVR_REQUIRE(n_layers >= 1 && n_layers <= CH_MAX_LAYERS, "1..8 layers");
    AVR_REQUIRE(M >= 0 && M < (1ll << 31), "bad row count");
    AVR_REQUIRE(k0 > 0 && k0 <= 128 && k0 % 16 == 0, "the chain input must be 16..128 columns wide, a multiple of 16");
    AVR_ENTER(device);
    if (M == 0) return AVR_OK;
    static_assert(sizeof(ChainMaps) + sizeof(ChainParams) < 4000, "kernel parameters");
    ChainMaps maps;
    ChainParams p = {};
    p.M = (int)M; p.n_layers = n_layers; p.tiles = (int)ceil_div(M, UM);
    if (int rc = make_map(&maps.x0, x0, M, k0, ldx, x_plane, UM, 3)) return rc;
    int k_in = k0;
    for (int l = 0; l < n_layers; ++l) {
        const avr_chain_layer& a = layers[l];
        ChainLayer& L = p.L[l];
        const bool last = l == n_layers - 1;
        AVR_REQUIRE(a.w && a.k_in == k_in, "layer input width does not match the previous layer's output");
        AVR_REQUIRE(a.n_out % 16 == 0 && a.n_out >= 16 && a.n_out <= 128, "layer widths must be multiples of 16 up to 128");
        AVR_REQUIRE(last || a.n_out == 128, "hidden layers of a fused chain are 128 wide");
        AVR_REQUIRE(last || !a.out_f32, "only the last layer writes fp32");
        AVR_REQUIRE(!a.out_f32 || (!a.save && !a.save_raw && !a.bits && !a.relu && a.ld_f32 % 4 == 0 && aligned16(a.out_f32)),
                    "an fp32 output layer has no activation, planes or bitmask");
        AVR_REQUIRE(a.out_f32 || a.n_out % 64 == 0, "plane outputs are stored in 64-column blocks");
        AVR_REQUIRE(!a.bits || a.ldbits * 32 >= a.n_out, "bitmask rows too short");
        L.K = a.k_in; L.N = a.n_out; L.relu = a.relu ? 1 : 0;
        L.bits = a.bits; L.ldbits = a.ldbits; L.out_f32 = a.out_f32; L.ld_f32 = a.ld_f32;
        if (int rc = make_map(&maps.w[l], a.w, a.n_out, a.k_in, a.ldw, a.w_plane, a.n_out, 3)) return rc;
        maps.save[l] = maps.x0; maps.raw[l] = maps.x0;
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
        if (raw) {
            AVR_REQUIRE(a.raw_kind == AVR_PLANES_BF16x2 || a.raw_kind == AVR_PLANES_BF16x3, "saved planes are bf16 pairs or triples");
            L.raw_planes = planes_count(a.raw_kind);
            if (int rc = make_map(&maps.raw[l], raw, M, a.n_out, a.ld_raw, a.raw_plane, UM, L.raw_planes)) return rc;
        }
        k_in = a.n_out;
    }
    static std::atomic<uint64_t> attr_done{0};
    const uint64_t bit = 1ull << (device & 63);
    if (!(attr_done.load(std::memory_order_acquire) & bit)) {
        AVR_CUDA(cudaFuncSetAttribute(mlp_chain_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CH_SMEM));
        attr_done.fetch_or(bit, std::memory_order_release);
    }
    const int grid = p.tiles < num_sms(device) ? p.tiles : num_sms(device);
    mlp_chain_fwd_kernel<<<grid, CH_THREADS, CH_SMEM, (cudaStream_t)stream>>>(maps, p);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

}  // extern "C"
