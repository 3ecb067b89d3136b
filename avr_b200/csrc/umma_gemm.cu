// Dense layers on the 5th-generation tensor cores: tcgen05.mma + TMEM accumulators + TMA operand staging.
//
// Replaces the tiny-cuda-nn FullyFusedMLP / CutlassMLP kernels of model.py:117,146,176 (and their backward).
// tcnn computes in fp16 with fp16 accumulators; the parity target here is the fp32 oracle (1e-4 rel-L2 on
// the IR and on every parameter gradient), which single-pass bf16/tf32 products cannot meet.  Operands are
// therefore error-compensated SETS of 16-bit planes (include/avr_b200.h, plane-set kinds):
//   bf16 pair    x = hi + mid                  A*B ~= hi*hi + hi*mid + mid*hi                 (16 bits: gradients)
//   bf16 triple  x = hi + mid + lo             + hi*lo + lo*hi + mid*mid                      (24 bits: forward)
//   fp16 pair    x = hi + lo' * 2^-11          hi*hi  +  2^-11 * (hi*lo' + lo'*hi)            (24 bits in fp16's range)
// Each tcgen05.mma truncates the fp32 accumulator once, so the 24-bit modes keep TWO accumulators per tile in TMEM --
// hi*hi in one, the small products in the other -- and the epilogue adds them (scaled for fp16 pairs).
//
// Kernel anatomy (one persistent CTA per SM, 320 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor (128B swizzle) of every plane of the A and B tiles into
//               a ring of shared-memory stages, completion on mbarriers
//   warp 1      TMEM allocator + MMA issuer: one elected lane issues the 3 or 6 tcgen05.mma of every k16 step
//               (descriptors built once per 64-wide k-block), tcgen05.commit releases the stage / publishes the accumulator
//   warps 2-9   epilogue, one or two warps per TMEM lane group: tcgen05.ld the accumulators (double-buffered in TMEM
//               so the next tile's MMAs overlap), apply bias / mask / accumulate / ReLU, split into planes of the
//               output kind (or write fp32), stage in swizzled shared memory, TMA-store
// Two operand modes: K-major x K-major (forward, backward-data with a pre-transposed weight) and
// MN-major x MN-major with split-K over the sample points (weight gradients, deterministic second pass).
// Third template flavour, PAIR (both operand modes): the CTAs of a 2x1x1 thread-block cluster work as one tcgen05 CTA pair
// (cta_group::2) on two adjacent row tiles x one column tile -- one 256-row MMA issued by the leader, each CTA stages its
// own 128 rows of A and HALF of the B tile, the leader's "full" barriers count both CTAs' TMA bytes, its commits are
// multicast to both CTAs' "stage free" / "accumulator final" barriers, and the peer's epilogue arrives remotely on the
// leader's "accumulator drained" barrier.  Used for every wide product over many row tiles (see avr_umma_gemm_nt) and for
// weight gradients of at least two row tiles (avr_umma_gemm_tn; same split over k, bit-identical partial sums).
// In the MN-major flavour the epilogue warps can also convert an fp16-pair B tile to bf16 in shared memory (conv_b).
#include <cuda.h>
#include <cuda_bf16.h>
#include <atomic>
#include <mutex>
#include <unordered_map>
#include <stdlib.h>
#include <type_traits>
#include "common.cuh"
#include "umma_ptx.cuh"

namespace avr {

enum { UF_RELU = 1, UF_ACCUM = 2, UF_MASK = 4, UF_OUT_F32 = 8, UF_DUAL_RELU = 16, UF_BITS = 32, UF_BIAS = 64, UF_DUAL_COPY = 2048, UF_DEBUG_NOWAIT = 128, UF_DEBUG_NOSTORE = 256, UF_DEBUG_NOSTAGE = 512, UF_DEBUG_NOFENCE = 1024 };
// Timing experiments (epilogue stages switched off, tile shapes forced through environment variables) exist only in
// builds made with -DAVR_EXPERIMENTS (python -m avr_b200.build --experiments); the shipped library reads no
// environment variable and its numerics cannot be changed from outside.
#ifdef AVR_EXPERIMENTS
#define UF_DBG(flags, f) ((flags) & (f))
#define AVR_EXP_ENV(name) getenv(name)
#else
#define UF_DBG(flags, f) 0
#define AVR_EXP_ENV(name) ((const char*)nullptr)
#endif

struct UmmaParams {
    int M, N, K;            // K-major: rows, cols, reduction.  MN-major: A extent, B extent, reduction (points)
    int BN;                 // tile columns (multiple of 16, <= 256)
    int tiles_m, tiles_n, k_splits, k_per_split;
    int stages, tmem_cols;
    int dual_acc;           // six-product mode: the five small products go to a second TMEM accumulator (see the MMA issuer)
    int acc_bufs;           // tiles buffered in TMEM: 2 (the epilogue of one overlaps the MMAs of the next) or 1 (256-wide
                            // tiles with two accumulators fill the 512 columns: less operand traffic, no overlap)
    int na, nb, nc;         // planes used of A / B (2: hi,mid  3: hi,mid,lo) and written to C
    int fa, fb, kc, kc2;    // A / B planes are fp16 (hi, lo'*2^11) pairs; plane-set kinds of C and C2 (AVR_PLANES_*)
    int mode;               // product schedule, see the MMA issuer
    float small_scale;      // factor of the small-products accumulator in the epilogue (2^-11 with an fp16 operand)
    int b_resident, nkb;    // K-major, single column tile, short K: the whole B operand stays in shared memory
    int conv_b;             // MN-major only: the B tile arrives as an fp16 (hi, lo') pair and the otherwise idle epilogue warps
                            // turn it into a bf16 (hi, mid) pair in shared memory before the MMAs read it (one element format
                            // per MMA; the gradient operand needs bf16's range)
    int cluster;            // 2: CTA pairs (thread-block cluster 2x1x1, tcgen05 cta_group::2) multiply two adjacent row tiles
                            // by one column tile as ONE 256-row MMA; each CTA stages only half of the B tile
    int flags;
    __nv_bfloat16* c; long long ldc, c_plane;          // plane-pair output
    __nv_bfloat16* c2; long long ldc2, c2_plane;       // second (ReLU'd) plane-pair output
    const uint32_t* mask; long long ldmask;            // ReLU bitmask of the gating activation: bit (col % 32) of word [row][col / 32]
    uint32_t* bits; long long ldbits;                  // bitmask (value > 0) written by the forward epilogue (UF_BITS)
    const float* bias_ray; const float* bias_rcv;      // UF_BIAS: out[row,:] += bias_ray[ray(row),:] + bias_rcv[receiver(row),:]
    long long ld_bias_ray, ld_bias_rcv;
    int geo_R, geo_S;                                  // row = (b*R + r)*S + s
    float* c32; long long ldc32;                       // fp32 output / split-K partials
    int epi_split;                                     // 1: warps 2..5 drain every chunk; 2: warps 6..9 take the odd chunks
    uint32_t epi_warp_bytes;                           // staging bytes per epilogue warp and buffer (nc planes x 2 KB)
    int epi_bufs;                                      // staging buffers per epilogue warp: 2 = a chunk is converted and staged
                                                       // while the bulk store of the previous one still reads its buffer
    // near-zero guard (forward layers): elements with |out| < near_tau * (mean |out| of their row chunk) are listed
    // and re-evaluated by umma_fixup_kernel with fp32 FMAs, so that sign decisions (ReLU, |leaky_relu|) taken on the
    // output are as good as an fp32 GEMM's.  The tensor core truncates once per MMA: ~6e-9*K of the row scale.
    uint2* near_list; uint32_t near_cap; uint32_t* near_count; float near_tau;
    const __nv_bfloat16* a_raw; long long lda, a_plane;
    const __nv_bfloat16* b_raw; long long ldb, b_plane;
};

// ... and write them into the warp's staging tiles ([plane][32 rows][64 B], 64-byte swizzle: 16-byte chunk c of row r
// lives at chunk c ^ ((r >> 1) & 3)).  Kept apart so that the conversions run BEFORE the warp waits for the previous
// chunk's bulk store to release the staging tiles.
__device__ __forceinline__ void stage_packed32(uint32_t stage, int lane, int nplanes, const uint32_t (&h)[16],
                                               const uint32_t (&m)[16], const uint32_t (&l)[16]) {
    const uint32_t row_base = stage + (uint32_t)lane * 64u;
    const uint32_t sw = (uint32_t)((lane >> 1) & 3);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t off = row_base + (((uint32_t)q ^ sw) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(off), "r"(h[4 * q]), "r"(h[4 * q + 1]), "r"(h[4 * q + 2]), "r"(h[4 * q + 3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(off + EPI_PLANE_BYTES), "r"(m[4 * q]), "r"(m[4 * q + 1]), "r"(m[4 * q + 2]), "r"(m[4 * q + 3]) : "memory");
        if (nplanes == 3)
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(off + 2 * EPI_PLANE_BYTES), "r"(l[4 * q]), "r"(l[4 * q + 1]), "r"(l[4 * q + 2]), "r"(l[4 * q + 3]) : "memory");
    }
}

// near-zero guard of one row chunk held in registers: v[0..NV) are columns col0.. of `row`.  Cheap common path: one
// FMNMX per element for the minimum magnitude plus a strided sample of the magnitudes for the scale.
template <int NV>
__device__ __forceinline__ void near_zero_guard(const UmmaParams& p, const float* v, long long row, bool row_ok, long long col0) {
    float m = fabsf(v[0]), s = 0.f;
#pragma unroll
    for (int i = 1; i < NV; ++i) m = fminf(m, fabsf(v[i]));
#pragma unroll
    for (int i = 0; i < NV; i += NV / 8) s += fabsf(v[i]);                  // eight samples of the row chunk
    const float thr = p.near_tau * s * 0.125f;
    if (m < thr && row_ok) {                                                  // rare
        uint32_t near = 0;
#pragma unroll
        for (int i = 0; i < NV; ++i) near |= (fabsf(v[i]) < thr ? 1u : 0u) << i;
        while (near) {
            const int i = __ffs(near) - 1;
            near &= near - 1;
            if (col0 + i >= p.N) break;
            const uint32_t slot = atomicAdd(p.near_count, 1u);
            if (slot < p.near_cap) p.near_list[slot] = make_uint2((uint32_t)row, (uint32_t)(col0 + i));
        }
    }
}

// ---------------------------------------------------------------------------------------------- kernel
template <bool MN_MAJOR, bool PAIR>
__global__ void __launch_bounds__(UTHREADS, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2,
                 const __grid_constant__ CUtensorMap tmBh, const UmmaParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // (a separate instantiation: a kernel that contains cta_group::2 instructions cannot be launched without a cluster)
    constexpr bool pair = PAIR;
    uint32_t crank = 0u;
    if constexpr (PAIR) crank = cluster_ctarank();
    const bool leader = crank == 0u;
    // B rows (K-major; pair mode: this CTA's half of the column tile) / MN extent (MN-major) in smem
    const int bn_rows = MN_MAJOR ? (((pair ? p.BN >> 1 : p.BN) + 63) / 64) * 64 : (pair ? p.BN >> 1 : p.BN);
    const uint32_t b_plane_bytes = MN_MAJOR ? 8192u : (uint32_t)bn_rows * 128u;
    const uint32_t a_tile_bytes = (uint32_t)p.na * A_PLANE_BYTES;
    const uint32_t b_tile_bytes = MN_MAJOR ? (uint32_t)p.nb * (uint32_t)bn_rows * 128u : (uint32_t)p.nb * b_plane_bytes;
    const uint32_t mn_blk_a = (uint32_t)p.na * 8192u;     // MN-major: one TMA box = 64 (mn) x 64 (k) x planes
    const uint32_t mn_blk_b = (uint32_t)p.nb * 8192u;
    const bool bres = !MN_MAJOR && p.b_resident;
    const uint32_t bres_bytes = bres ? (uint32_t)p.nkb * b_tile_bytes : 0u;    // resident B: [k-block][plane][rows][128 B]
    const uint32_t stage_bytes = bres ? a_tile_bytes : a_tile_bytes + b_tile_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + bres_bytes + (size_t)p.stages * stage_bytes);
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * p.stages;
    const uint32_t bar_tfull = bar_empty + 8 * p.stages, bar_tempty = bar_tfull + 16, bar_bres = bar_tempty + 16;
    const uint32_t bar_conv = bar_bres + 8;                  // [stages]: the B tile of a stage has been converted (conv_b)
    const uint32_t bar_bload = bar_conv + 8 * p.stages;      // [stages]: pair mode + conv_b: THIS CTA's half of the B tile has landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * p.stages + 5);
    const uint32_t bres_base = smem_u32(smem);
    const uint32_t smem_base = bres_base + bres_bytes;
    // epilogue staging: (4 or 8) warps x nc planes x 2 KB, 1024-byte aligned, after the barrier block
    const uint32_t epi_base = (smem_base + (uint32_t)p.stages * stage_bytes + 256u + 1023u) & ~1023u;

    if (warp == 0 && lane == 0) {
        // pair mode: the leader's "full" barrier counts both producers (its own arrive + expect_tx of both CTAs' bytes, the
        // peer's remote arrive) and its "accumulator drained" barrier both CTAs' epilogue threads; "stage free" and
        // "accumulator final" are announced to each CTA's own barrier by the leader's multicast commits
        for (int s = 0; s < p.stages; ++s) { mbar_init(bar_full + 8 * s, pair ? 2 : 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, (pair ? 256 : 128) * p.epi_split); }
        mbar_init(bar_bres, 1);
        // one arrival per converting warp (pair mode: of both CTAs, on the leader's barrier)
        for (int s = 0; s < p.stages; ++s) { mbar_init(bar_conv + 8 * s, 4 * p.epi_split * (pair ? 2 : 1)); mbar_init(bar_bload + 8 * s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1) {
        if constexpr (pair) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if constexpr (pair) cluster_sync_all(); else __syncthreads();      // the peer's barriers must exist before anything is sent to them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // work list: tiles (row tile, column tile, K slice) strided over the CTAs, or -- pair mode -- units of two adjacent
    // row tiles x one column tile strided over the clusters, rank r of the pair taking row tile 2 u + r
    const int n_tiles = (pair ? ((p.tiles_m + 1) / 2) * p.tiles_n : p.tiles_m * p.tiles_n) * p.k_splits;
    const int t_begin = pair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int t_stride = pair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    auto row_tile = [&](int mn) { const int mt = mn / p.tiles_n; return pair ? 2 * mt + (int)crank : mt; };
    if (warp == 0) {
        // ===================================== TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            if (bres && t_begin < n_tiles) {                                  // B is the same for every tile of this CTA
                mbar_expect_tx(bar_bres, bres_bytes);
                for (int kb = 0; kb < p.nkb; ++kb) tma_load_3d(bres_base + kb * b_tile_bytes, &tmB, bar_bres, kb * UBK, 0, 0);
            }
            for (int tile = t_begin; tile < n_tiles; tile += t_stride) {
                const int split = tile % p.k_splits;
                const int mn = tile / p.k_splits;
                const int m0 = row_tile(mn) * UM, n0 = (mn % p.tiles_n) * p.BN;
                // split-K slices are interleaved (k-block j of slice s is block j*k_splits + s) so that the CTAs
                // sharing operand columns walk the same rows at the same time and the re-reads hit L2
                for (int k0 = split * UBK; k0 < p.K; k0 += p.k_splits * UBK) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    const uint32_t full = bar_full + 8 * stage;
                    const uint32_t sa = smem_base + stage * stage_bytes, sb = sa + a_tile_bytes;
                    // pair mode with the in-kernel conversion: the B halves complete on each CTA's OWN barrier (its converting
                    // warps wait there); the leader's "full" barrier then counts the A bytes only
                    const bool b_local = pair && MN_MAJOR && p.conv_b;
                    if constexpr (!pair) mbar_expect_tx(full, stage_bytes);
                    else if (leader) mbar_expect_tx(full, 2u * (b_local ? a_tile_bytes : stage_bytes));
                    else mbar_arrive_cluster(full, 0u);
                    if constexpr (pair && MN_MAJOR) {
                        // my 128 columns of A (two 64-wide boxes) and my half of the B columns
                        tma_load_3d_2sm(sa, &tmA, full, m0, k0, 0);
                        tma_load_3d_2sm(sa + mn_blk_a, &tmA, full, m0 + 64, k0, 0);
                        // (my half starts BN / 2 columns into the tile -- not bn_rows, which is BN / 2 rounded up to whole
                        // 64-column boxes; the columns a box holds beyond the half are never read by the MMAs)
                        const int nb0 = n0 + (int)crank * (p.BN >> 1);
                        if (b_local) mbar_expect_tx(bar_bload + 8 * stage, b_tile_bytes);
                        for (int i = 0; i < bn_rows / 64; ++i) {
                            if (b_local) tma_load_3d(sb + mn_blk_b * i, &tmB, bar_bload + 8 * stage, nb0 + 64 * i, k0, 0);
                            else tma_load_3d_2sm(sb + mn_blk_b * i, &tmB, full, nb0 + 64 * i, k0, 0);
                        }
                    } else if constexpr (pair) {
                        // my row tile and my half of the rows of the column tile; the bytes count on the leader's barrier
                        tma_load_3d_2sm(sa, &tmA, full, k0, m0, 0);
                        tma_load_3d_2sm(sb, &tmBh, full, k0, n0 + (int)crank * bn_rows, 0);
                    } else if (!MN_MAJOR) {
                        tma_load_3d(sa, &tmA, full, k0, m0, 0);
                        if (!bres) tma_load_3d(sb, &tmB, full, k0, n0, 0);
                    } else {
                        tma_load_3d(sa, &tmA, full, m0, k0, 0);
                        tma_load_3d(sa + mn_blk_a, &tmA, full, m0 + 64, k0, 0);
                        for (int i = 0; i < bn_rows / 64; ++i) tma_load_3d(sb + mn_blk_b * i, &tmB, full, n0 + 64 * i, k0, 0);
                    }
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer
        // (pair mode: the leader's thread issues for both CTAs -- M = 256, each CTA's tensor core takes its 128 rows of A and
        // its half of B from its own shared memory at the same offsets, and accumulates into its own tensor memory)
        // The whole warp runs this loop in lock step and ONE elected lane executes the tcgen05 instructions: everything an
        // MMA needs (descriptors, accumulator addresses, flags) is then warp-uniform and lives in uniform registers.  Inside
        // an `if (lane == 0)` branch the compiler cannot prove that and wraps every UTCHMMA in an ELECT / R2UR.BROADCAST /
        // BRA.U.ANY loop (~15 dependent instructions per MMA, 90-140 clk): longer than the 64 clk a 128-column MMA occupies
        // the tensor pipe, so 128-column tiles were bound by the issuing thread.
        auto issue_tiles = [&](auto cta2_tag) {
            constexpr bool CTA2 = decltype(cta2_tag)::value;
            const bool elected = elect_one() != 0u;
            // instruction descriptor: fp32 accumulate, A / B element formats (0 = f16, 1 = bf16), N, M
            uint32_t idesc = (1u << 4) | ((p.fa ? 0u : 1u) << 7) | ((p.fb ? 0u : 1u) << 10) | ((uint32_t)(p.BN >> 3) << 17) |
                             ((uint32_t)((CTA2 ? 2 * UM : UM) >> 4) << 24);
            if (MN_MAJOR) idesc |= (1u << 15) | (1u << 16);
            int stage = 0, iter = 0;
            uint32_t phase = 0;
            if (bres && t_begin < n_tiles) mbar_wait(bar_bres, 0);
            for (int tile = t_begin; tile < n_tiles; tile += t_stride, ++iter) {
                const int split = tile % p.k_splits;
                const int acc = p.acc_bufs == 2 ? (iter & 1) : 0;
                const uint32_t acc_phase = (uint32_t)(p.acc_bufs == 2 ? (iter >> 1) : iter) & 1u;
                mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1u);
                tc_fence_after();
                // Each tcgen05.mma truncates the fp32 accumulator once (measured: ~1.5e-8 of its magnitude per MMA, toward
                // zero).  With one accumulator all 6*K/16 MMAs of the six-product mode truncate at full magnitude; with
                // dual_acc only the K/16 hi*hi MMAs do -- the other five products (<= 2^-8 of it) collect in a second
                // accumulator whose truncations are 2^-8 smaller, and the epilogue adds the two in fp32.
                const uint32_t d_main = tmem_base + (uint32_t)(acc * p.BN * (p.dual_acc ? 2 : 1));
                const uint32_t d_small = p.dual_acc ? d_main + (uint32_t)p.BN : d_main;
                uint32_t acc_main = 0u, acc_small = 0u;                           // "accumulate" flags (one accumulator when !dual_acc)
                const uint32_t shared_acc = p.dual_acc ? 0u : 0xffffffffu;
                for (int k0 = split * UBK; k0 < p.K; k0 += p.k_splits * UBK) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    if (MN_MAJOR && p.conv_b) mbar_wait(bar_conv + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * stage_bytes;
                    const uint32_t sb = bres ? bres_base + (uint32_t)(k0 / UBK) * b_tile_bytes : sa + a_tile_bytes;
                    const int k_steps = min(UBK / 16, (p.K - k0 + 15) / 16);    // the zero-filled tail of the last block is skipped
                    // The single issuing thread must stay ahead of the tensor pipe (a 128 x 80 x 16 MMA retires in ~40 clk):
                    // plane descriptors are built once per k-block, a k16 step only adds a constant to their address
                    // field (32 B K-major, 2 KB MN-major, in 16-byte units), and each schedule is straight-line code.
                    uint64_t a0, a1, a2 = 0, b0, b1, b2 = 0;                    // planes 0 (hi), 1, 2 of A and B
                    if (!MN_MAJOR) {
                        a0 = smem_desc(sa, 16, 1024); a1 = smem_desc(sa + A_PLANE_BYTES, 16, 1024);
                        b0 = smem_desc(sb, 16, 1024); b1 = smem_desc(sb + b_plane_bytes, 16, 1024);
                        if (p.mode == 1) { a2 = smem_desc(sa + 2 * A_PLANE_BYTES, 16, 1024); b2 = smem_desc(sb + 2 * b_plane_bytes, 16, 1024); }
                    } else {
                        a0 = smem_desc(sa, mn_blk_a, 1024); a1 = smem_desc(sa + 8192, mn_blk_a, 1024);
                        b0 = smem_desc(sb, mn_blk_b, 1024); b1 = smem_desc(sb + 8192, mn_blk_b, 1024);
                        if (p.mode == 1) { a2 = smem_desc(sa + 16384, mn_blk_a, 1024); b2 = smem_desc(sb + 16384, mn_blk_b, 1024); }
                    }
                    const uint64_t step = MN_MAJOR ? (2048u >> 4) : (32u >> 4);
                    // Product schedules, smallest terms first.  With fp16 pairs the products that involve a lo' plane carry
                    // a factor 2^11 and MUST go to the small accumulator (scaled back in the epilogue).  A and B share one
                    // element format: tcgen05.mma raises an illegal-instruction fault on a bf16 x f16 mix.
                    if (!elected) {
                    } else if (p.mode == 1) {   // bf16x3 . bf16x3: six products (fp32-grade pre-activations / ill-conditioned sums)
#pragma unroll
                        for (int j = 0; j < UBK / 16; ++j) {
                            if (j >= k_steps) break;
                            const uint64_t o = step * j;
                            umma_issue<CTA2>(d_small, a2 + o, b0 + o, idesc, acc_small | (acc_main & shared_acc));
                            umma_issue<CTA2>(d_small, a0 + o, b2 + o, idesc, 1u);
                            umma_issue<CTA2>(d_small, a1 + o, b1 + o, idesc, 1u);
                            umma_issue<CTA2>(d_small, a1 + o, b0 + o, idesc, 1u);
                            umma_issue<CTA2>(d_small, a0 + o, b1 + o, idesc, 1u);
                            umma_issue<CTA2>(d_main, a0 + o, b0 + o, idesc, p.dual_acc ? acc_main : 1u);
                            acc_small = acc_main = 1u;
                        }
                    } else {                    // bf16x2 . bf16x2 (16 bits; gradients) and f16x2 . f16x2 (24 bits): three products
#pragma unroll
                        for (int j = 0; j < UBK / 16; ++j) {
                            if (j >= k_steps) break;
                            const uint64_t o = step * j;
                            umma_issue<CTA2>(d_small, a1 + o, b0 + o, idesc, acc_small | (acc_main & shared_acc));
                            umma_issue<CTA2>(d_small, a0 + o, b1 + o, idesc, 1u);
                            umma_issue<CTA2>(d_main, a0 + o, b0 + o, idesc, p.dual_acc ? acc_main : 1u);
                            acc_small = acc_main = 1u;
                        }
                    }
                    acc_small = acc_main = 1u;                                   // (all lanes: the flags stay warp-uniform)
                    if (elected) {
                        if constexpr (CTA2) umma_commit_2sm(bar_empty + 8 * stage);      // frees the stage (in both CTAs) when the MMAs retire
                        else umma_commit(bar_empty + 8 * stage);
                    }
                    __syncwarp();
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                if (elected) {
                    if constexpr (CTA2) umma_commit_2sm(bar_tfull + 8 * acc);            // accumulator complete (both CTAs' epilogues)
                    else umma_commit(bar_tfull + 8 * acc);
                }
                __syncwarp();
            }
        };
        if constexpr (!pair) issue_tiles(std::false_type{});
        else if (leader) issue_tiles(std::true_type{});
    } else if (((warp - 2) >> 2) < p.epi_split) {
        // ===================================== epilogue: warps 2..5 (and 6..9) -> TMEM lane groups 2,3,0,1.
        // One chunk of a tile is a latency chain (tcgen05.ld -> convert -> st.shared -> proxy fence -> bulk store ->
        // wait for the store to have read the staging tile); with epi_split = 2 the two warps of a lane group drain
        // alternate chunks through their own staging tiles, so two such chains are in flight per lane group.
        const int lane_grp = warp & 3;
        const int half = (warp - 2) >> 2;
        int iter = 0, cstage = 0;
        uint32_t cphase = 0;
        for (int tile = t_begin; tile < n_tiles; tile += t_stride, ++iter) {
            const int split = tile % p.k_splits;
            const int mn = tile / p.k_splits;
            const int m0 = row_tile(mn) * UM, n0 = (mn % p.tiles_n) * p.BN;
            const int acc = p.acc_bufs == 2 ? (iter & 1) : 0;
            const uint32_t acc_phase = (uint32_t)(p.acc_bufs == 2 ? (iter >> 1) : iter) & 1u;
            const long long row = m0 + lane_grp * 32 + lane;
            const bool row_ok = row < p.M;
            if (MN_MAJOR && p.conv_b) {
                // Weight gradients from an fp16-pair activation: while the MMAs of this work item run, the epilogue warps
                // (idle until the accumulator is final) rewrite every B tile in place, element by element -- the two
                // planes share one layout -- from x = hi + lo' * 2^-11 (exact in fp32) to bf16 (hi, mid), exactly what
                // avr_planes_split would have written; ~700 issue cycles per k-block against ~1 500 clk of MMAs.
                const int n_conv = 128 * p.epi_split, t_conv = (warp - 2) * 32 + lane;
                const uint32_t chunks = (uint32_t)(bn_rows / 64) * 512u;          // 16-byte chunks per plane of the B tile
                for (int k0 = split * UBK; k0 < p.K; k0 += p.k_splits * UBK) {
                    mbar_wait((pair ? bar_bload : bar_full) + 8 * cstage, cphase);
                    const uint32_t sb = smem_base + cstage * stage_bytes + a_tile_bytes;
                    for (uint32_t c = t_conv; c < chunks; c += n_conv) {
                        const uint32_t a0 = sb + (c >> 9) * mn_blk_b + (c & 511u) * 16u, a1 = a0 + 8192u;
                        uint32_t h[4], l[4];
                        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(h[0]), "=r"(h[1]), "=r"(h[2]), "=r"(h[3]) : "r"(a0));
                        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(l[0]), "=r"(l[1]), "=r"(l[2]), "=r"(l[3]) : "r"(a1));
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float2 x = planes_unpack2(true, h[i], l[i]);
                            uint32_t unused;
                            planes_pack2(AVR_PLANES_BF16x2, x.x, x.y, h[i], l[i], unused);
                        }
                        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a0), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a1), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
                    }
                    fence_async_smem();                                          // generic-proxy writes -> visible to the MMAs
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (pair) mbar_arrive_cluster(bar_conv + 8 * cstage, 0u);     // the leader's MMA warp waits for both CTAs
                        else mbar_arrive(bar_conv + 8 * cstage);
                    }
                    if (++cstage == p.stages) { cstage = 0; cphase ^= 1u; }
                }
            }
            // per-tile operands of the epilogue are fetched BEFORE waiting for the accumulator, so that their
            // (row-strided / L2) load latency overlaps the tile's MMAs:
            //   mw[j]   ReLU gating word of columns [32j, 32j+32) of this row            (UF_MASK)
            //   bt*[j]  value of column 32j + lane of the warp's shared bias row           (UF_BIAS, uniform case)
            uint32_t mw[8];
            int ray_row = 0, rcv_row = 0, ray_uniform = 0, rcv_uniform = 0;
            const uint32_t bias_s = epi_base + 4u * (uint32_t)(p.epi_split * p.epi_bufs) * p.epi_warp_bytes + (uint32_t)lane_grp * 1024u;   // 256 floats (UF_BIAS implies epi_split = 1)
            if (!MN_MAJOR && !(p.flags & UF_OUT_F32)) {
                if (p.flags & UF_BIAS) {
                    const long long rrow = row_ok ? row : (long long)p.M - 1;
                    ray_row = (int)((rrow / p.geo_S) % p.geo_R);
                    rcv_row = (int)(rrow / ((long long)p.geo_R * p.geo_S));
                    __match_all_sync(0xffffffffu, ray_row, &ray_uniform);
                    __match_all_sync(0xffffffffu, rcv_row, &rcv_uniform);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    mw[j] = 0xffffffffu;
                    const long long col = n0 + 32 * j;
                    if (32 * j >= p.BN || col >= p.N) continue;
                    if ((p.flags & UF_MASK) && row_ok) mw[j] = __ldg(p.mask + row * p.ldmask + (col >> 5));
                    if (p.flags & UF_BIAS) {                                     // the warp's bias row segment -> smem
                        float tr = 0.f, tb = 0.f;
                        if (col + lane < p.N) {
                            if (p.bias_ray && ray_uniform) tr = __ldg(p.bias_ray + (long long)ray_row * p.ld_bias_ray + col + lane);
                            if (p.bias_rcv && rcv_uniform) tb = __ldg(p.bias_rcv + (long long)rcv_row * p.ld_bias_rcv + col + lane);
                        }
                        asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_s + (uint32_t)(32 * j + lane) * 4u), "f"(tr + tb) : "memory");
                    }
                }
                if (p.flags & UF_BIAS) __syncwarp();
            }
            mbar_wait(bar_tfull + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(acc * p.BN * (p.dual_acc ? 2 : 1));
            if (MN_MAJOR || (p.flags & UF_OUT_F32)) {
                // fp32 outputs (split-K partials, the 16-wide density head): direct stores
                for (int c0 = 16 * half; c0 < p.BN; c0 += 16 * p.epi_split) {
                    float v[16];
                    tmem_ld16(taddr + c0, v);                                   // warp-collective: outside the row guard
                    if (p.dual_acc) {
                        float u[16];
                        tmem_ld16(taddr + p.BN + c0, u);
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = fmaf(u[i], p.small_scale, v[i]);
                    }
                    if (!row_ok) continue;
                    if (!MN_MAJOR && p.near_list) near_zero_guard<16>(p, v, row, row_ok, n0 + c0);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const long long col = n0 + c0 + 8 * h;
                        if (col >= p.N) continue;
                        float* x = v + 8 * h;
                        float* dst = (MN_MAJOR || p.k_splits > 1) ? p.c32 + ((long long)split * p.M + row) * p.ldc32 + col
                                                                  : p.c32 + row * p.ldc32 + col;
                        if (!MN_MAJOR && (p.flags & UF_ACCUM)) {
                            const float4 o0 = *reinterpret_cast<const float4*>(dst), o1 = *reinterpret_cast<const float4*>(dst + 4);
                            x[0] += o0.x; x[1] += o0.y; x[2] += o0.z; x[3] += o0.w; x[4] += o1.x; x[5] += o1.y; x[6] += o1.z; x[7] += o1.w;
                        }
                        if (p.flags & UF_RELU) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) x[i] = fmaxf(x[i], 0.f);
                        }
                        *reinterpret_cast<float4*>(dst) = make_float4(x[0], x[1], x[2], x[3]);
                        *reinterpret_cast<float4*>(dst + 4) = make_float4(x[4], x[5], x[6], x[7]);
                    }
                }
            } else {
                // plane outputs: registers -> swizzled smem staging -> TMA bulk tensor store (full lines, rows and
                // columns outside the output window are clipped by the tensor map)
                const uint32_t stage0 = epi_base + (uint32_t)((4 * half + lane_grp) * p.epi_bufs) * p.epi_warp_bytes;
                uint32_t sbuf = 0;
                for (int c0 = EPI_COLS * half; c0 < p.BN && n0 + c0 < p.N; c0 += EPI_COLS * p.epi_split) {
                    float v[32];
                    {
                        float t0[16], t1[16];
                        const bool second = c0 + 16 < p.BN;                     // ragged last chunk (single-tile N only): the
                        if (p.dual_acc && second) {
                            tmem_ld32_dual(taddr + c0, taddr + p.BN + c0, p.small_scale, v);
                        } else if (p.dual_acc) {                                // missing columns lie outside the output
                            tmem_ld16x2(taddr + c0, taddr + p.BN + c0, t0, t1); // window and are clipped by TMA
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = fmaf(t1[i], p.small_scale, t0[i]);   // + the small-products accumulator
                            if (second) {
                                tmem_ld16x2(taddr + c0 + 16, taddr + p.BN + c0 + 16, t0, t1);
#pragma unroll
                                for (int i = 0; i < 16; ++i) v[16 + i] = fmaf(t1[i], p.small_scale, t0[i]);
                            } else {
#pragma unroll
                                for (int i = 0; i < 16; ++i) v[16 + i] = 0.f;
                            }
                        } else {
                            if (second) {
                                tmem_ld16x2(taddr + c0, taddr + c0 + 16, t0, t1);
                            } else {
                                tmem_ld16(taddr + c0, t0);
#pragma unroll
                                for (int i = 0; i < 16; ++i) t1[i] = 0.f;
                            }
#pragma unroll
                            for (int i = 0; i < 16; ++i) { v[i] = t0[i]; v[16 + i] = t1[i]; }
                        }
                    }
                    if (p.flags & UF_BIAS) {
                        // per-ray / per-receiver additive terms of the first signal layer (the broadcast inputs of
                        // renderer.py:59-60 folded through W0).  The 32 rows of a warp almost always share their ray and
                        // receiver: the prefetched lane value is then broadcast with shuffles.
                        // (uniform tables were summed into the warp's smem row above: broadcast reads, no shuffles)
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            float4 t;
                            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                                         : "r"(bias_s + (uint32_t)(c0 + 4 * q) * 4u));
                            v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
                        }
#pragma unroll
                        for (int tsel = 0; tsel < 2; ++tsel) {                   // rare: rows of the warp straddle rays / receivers
                            const float* tab = tsel == 0 ? p.bias_ray : p.bias_rcv;
                            if (tab == nullptr || (tsel == 0 ? ray_uniform : rcv_uniform)) continue;
                            const long long ldt = tsel == 0 ? p.ld_bias_ray : p.ld_bias_rcv;
                            const int my = tsel == 0 ? ray_row : rcv_row;
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (n0 + c0 + i < p.N) v[i] += __ldg(tab + (long long)my * ldt + n0 + c0 + i);
                        }
                    }
                    if (p.near_list) near_zero_guard<32>(p, v, row, row_ok, n0 + c0);
                    if (p.flags & UF_MASK) {
                        uint32_t word = 0;
#pragma unroll
                        for (int j = 0; j < 8; ++j) word = (c0 >> 5) == j ? mw[j] : word;
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = (word >> i) & 1u ? v[i] : 0.f;
                    }
                    if (p.flags & UF_BITS) {
                        uint32_t word = 0;
#pragma unroll
                        for (int i = 0; i < 32; ++i) word |= (v[i] > 0.f ? 1u : 0u) << i;
                        if (row_ok) p.bits[row * p.ldbits + ((n0 + c0) >> 5)] = word;
                    }
                    if (p.flags & UF_ACCUM) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const long long col = n0 + c0 + 8 * q;
                            float* x = v + 8 * q;
                            if (!row_ok || col >= p.N) continue;
                            {
                                const __nv_bfloat16* dst = p.c + row * p.ldc + col;
                                const uint4 oh = *reinterpret_cast<const uint4*>(dst), ol = *reinterpret_cast<const uint4*>(dst + p.c_plane);
                                uint4 o3 = make_uint4(0u, 0u, 0u, 0u);
                                if (p.nc == 3) o3 = *reinterpret_cast<const uint4*>(dst + 2 * p.c_plane);
                                const uint32_t hw[4] = {oh.x, oh.y, oh.z, oh.w}, lw[4] = {ol.x, ol.y, ol.z, ol.w};
                                const uint32_t tw[4] = {o3.x, o3.y, o3.z, o3.w};
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    x[2 * i] += bf_lo(hw[i]) + (bf_lo(lw[i]) + bf_lo(tw[i]));
                                    x[2 * i + 1] += bf_hi(hw[i]) + (bf_hi(lw[i]) + bf_hi(tw[i]));
                                }
                            }
                        }
                    }
                    // second output: max(out, 0) next to the raw values (DUAL_RELU), or the same values as a plane set of
                    // another kind (DUAL_COPY: fp16 pair for the next forward layer + bf16 pair for the weight gradient)
                    const int n_out = (p.flags & (UF_DUAL_RELU | UF_DUAL_COPY)) ? 2 : 1;
                    for (int o = 0; o < n_out; ++o) {
                        uint32_t ph[16], pm[16], pl[16];
                        const int kind_o = o == 0 ? p.kc : p.kc2;
                        const int n_o = planes_count(kind_o);
                        pack_planes32(v, (o == 1 && (p.flags & UF_DUAL_RELU)) || (p.flags & UF_RELU), kind_o, ph, pm, pl);
                        // staging buffer free again?  Bulk-store groups complete in order: with two buffers, at most ONE
                        // group still in flight means that it is the store of the other buffer
                        if (lane == 0 && !UF_DBG(p.flags, UF_DEBUG_NOWAIT)) { if (p.epi_bufs == 2) tma_store_wait_read1(); else tma_store_wait_read(); }
                        __syncwarp();
                        const uint32_t stage = stage0 + sbuf * p.epi_warp_bytes;
                        sbuf ^= (uint32_t)(p.epi_bufs - 1);
                        if (!UF_DBG(p.flags, UF_DEBUG_NOSTAGE)) stage_packed32(stage, lane, n_o, ph, pm, pl);
                        if (!UF_DBG(p.flags, UF_DEBUG_NOFENCE)) fence_async_smem();
                        __syncwarp();
                        if (lane == 0 && !UF_DBG(p.flags, UF_DEBUG_NOSTORE)) {
                            const CUtensorMap* map = o == 0 ? &tmC : &tmC2;
                            for (int q = 0; q < n_o; ++q)
                                tma_store_3d(map, stage + q * EPI_PLANE_BYTES, n0 + c0, m0 + lane_grp * 32, q);
                            tma_store_commit();
                        }
                    }
                }
            }
            tc_fence_before();
            if constexpr (pair) mbar_arrive_cluster(bar_tempty + 8 * acc, 0u);             // the leader's MMA thread waits for both CTAs
            else mbar_arrive(bar_tempty + 8 * acc);
        }
        if (lane == 0) tma_store_wait_all();                                   // smem must outlive the bulk stores
    }
    tc_fence_before();
    if constexpr (pair) cluster_sync_all(); else __syncthreads();      // neither CTA of a pair leaves while the other may still signal it
    if (warp == 1) {
        tc_fence_after();
        if constexpr (pair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// Second pass of the near-zero guard: one warp per listed element re-evaluates the dot product with fp32 FMAs on the
// exact fp32 values of the operands (hi + mid + lo), applies the layer's epilogue and patches every output of the
// element (planes, second planes, fp32, ReLU bit).  A few thousand elements per layer at most.
__device__ __forceinline__ float planes_value(const __nv_bfloat16* q, long long plane, int n) {
    float tail = __bfloat162float(q[plane]);
    if (n == 3) tail += __bfloat162float(q[2 * plane]);
    return __bfloat162float(q[0]) + tail;
}
__device__ __forceinline__ void planes_patch(__nv_bfloat16* q, long long plane, int n, float v) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const float r = v - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16_rn(r);
    q[0] = h;
    q[plane] = m;
    if (n == 3) q[2 * plane] = __float2bfloat16_rn(r - __bfloat162float(m));
}
__global__ void __launch_bounds__(256) umma_fixup_kernel(const UmmaParams p) {
    const uint32_t total = *p.near_count;
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    if (total > p.near_cap && warp == 0 && lane == 0) {                       // list overflow: poison, never hide
        const float nan = __int_as_float(0x7fc00000);
        if (p.flags & UF_OUT_F32) p.c32[0] = nan; else planes_patch(p.c, p.c_plane, p.nc, nan);
    }
    const uint32_t n = total < p.near_cap ? total : p.near_cap;
    for (uint32_t e = warp; e < n; e += n_warps) {
        const uint2 rc = p.near_list[e];
        const long long row = rc.x, col = rc.y;
        const __nv_bfloat16* a = p.a_raw + row * p.lda;
        const __nv_bfloat16* b = p.b_raw + col * p.ldb;
        float acc = 0.f;
        for (int k = lane; k < p.K; k += 32) acc = fmaf(planes_value(a + k, p.a_plane, p.na), planes_value(b + k, p.b_plane, p.nb), acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane != 0) continue;
        if (p.flags & UF_BIAS) {
            float tr = 0.f, tb = 0.f;
            if (p.bias_ray) tr = p.bias_ray[((row / p.geo_S) % p.geo_R) * p.ld_bias_ray + col];
            if (p.bias_rcv) tb = p.bias_rcv[(row / ((long long)p.geo_R * p.geo_S)) * p.ld_bias_rcv + col];
            acc += tr + tb;
        }
        const float pos = fmaxf(acc, 0.f);
        if (p.flags & UF_OUT_F32) {
            p.c32[row * p.ldc32 + col] = (p.flags & UF_RELU) ? pos : acc;
            continue;
        }
        planes_patch(p.c + row * p.ldc + col, p.c_plane, p.nc, (p.flags & UF_RELU) ? pos : acc);
        if (p.flags & UF_DUAL_RELU) planes_patch(p.c2 + row * p.ldc2 + col, p.c2_plane, p.nc, pos);
        if (p.flags & UF_BITS) {
            uint32_t* w = p.bits + row * p.ldbits + (col >> 5);
            const uint32_t bit = 1u << (col & 31);
            if (acc > 0.f) atomicOr(w, bit); else atomicAnd(w, ~bit);
        }
    }
}

// fp32 [rows, cols] (ld) -> plane set [n][rows][ldp] of kind `kind`; optionally transposed (out[c][r] = in[r][c])
__global__ void planes_split_kernel(const float* __restrict__ x, long long rows, long long cols, long long ld,
                                    void* __restrict__ out, long long ldp, long long plane, int kind,
                                    int transpose, int relu) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const long long r = i / cols, c = i - r * cols;
    float v = x[r * ld + c];
    if (relu) v = fmaxf(v, 0.f);
    planes_store(out, transpose ? c * ldp + r : r * ldp + c, plane, kind, v);
}

// several matrices in one launch (the weight matrices of a pass: 11 forward, 12 transposed backward at simu)
struct SplitBatch {
    avr_split_desc d[AVR_SPLIT_BATCH_MAX];
};
__global__ void planes_split_batch_kernel(const __grid_constant__ SplitBatch b) {
    const avr_split_desc& d = b.d[blockIdx.y];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.rows * d.cols) return;
    const long long r = i / d.cols, c = i - r * d.cols;
    const float v = d.x[r * d.ld + c];
    planes_store(d.planes, d.transpose ? c * d.ldp + r : r * d.ldp + c, d.plane_stride, d.kind, v);
}

__global__ void planes_merge_kernel(const void* __restrict__ in, long long rows, long long cols, long long ldp,
                                    long long plane, int kind, float* __restrict__ out, long long ld) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const long long r = i / cols, c = i - r * cols;
    out[r * ld + c] = planes_load(in, r * ldp + c, plane, kind);
}

// dW[m, n] (+)= sum over splits of partial[split][m][n].  One warp per float4 of the output: lane l adds the
// splits l, l+32, ... in order, then a fixed butterfly combines the lanes -> deterministic and latency-tolerant.
__global__ void __launch_bounds__(256)
umma_splitk_reduce_kernel(const float* __restrict__ partial, long long ldp, int splits, long long M, long long N,
                          float* __restrict__ C, long long ldc, int accumulate) {
    const long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long nq = N / 4;
    if (q >= M * nq) return;
    const long long r = q / nq, c = (q - r * nq) * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int z = lane; z < splits; z += 32) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(partial + ((long long)z * M + r) * ldp + c));
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s.x += __shfl_xor_sync(0xffffffffu, s.x, o); s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
        s.z += __shfl_xor_sync(0xffffffffu, s.z, o); s.w += __shfl_xor_sync(0xffffffffu, s.w, o);
    }
    if (lane == 0) {
        float* dst = C + r * ldc + c;
        if (accumulate) { const float4 o = *reinterpret_cast<const float4*>(dst); s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w; }
        *reinterpret_cast<float4*>(dst) = s;
    }
}

// ---------------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    });
    return fn;
}

// plane-pair tensor [2][rows][ld] of bf16, logical width `cols`; box = (64 cols, box_rows, 2 planes), 128B swizzle
static int encode_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, long long plane,
                      int box_rows, int nplanes, bool store_map, int box_planes) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(AVR_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from the driver");
    if ((reinterpret_cast<uintptr_t>(base) & 15u) || (ld * 2) % 16 || (plane * 2) % 16)
        return fail(AVR_ERR_INVALID, "plane tensors need 16-byte aligned base, row pitch and plane pitch");
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)nplanes};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)plane * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, (cuuint32_t)box_planes};
    if (store_map) { box[0] = EPI_COLS; box[1] = 32; box[2] = 1; }          // epilogue staging tile: 32 rows x 64 B
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, store_map ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                     store_map ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return fail(AVR_ERR_INVALID, "cuTensorMapEncodeTiled failed with CUresult %d", (int)rc);
    return AVR_OK;
}

// A tensor map is a pure function of (base, shape, pitches, box): the caller's allocator hands the same buffers back
// step after step, so the four driver calls per GEMM (and their ~1.5 us each on the launching thread) are paid once per
// distinct operand.  Guarded by a mutex (concurrent callers: nn.DataParallel threads, autograd's device threads);
// bounded: the cache is dropped when it reaches 4096 entries.
struct MapKey {
    const void* base; long long rows, cols, ld, plane; int box_rows, nplanes, store, box_planes;
    bool operator==(const MapKey& o) const {
        return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && plane == o.plane && box_rows == o.box_rows &&
               nplanes == o.nplanes && store == o.store && box_planes == o.box_planes;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        uint64_t h = reinterpret_cast<uintptr_t>(k.base) * 0x9E3779B97F4A7C15ull;
        auto mix = [&h](uint64_t v) { h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); };
        mix((uint64_t)k.rows); mix((uint64_t)k.cols); mix((uint64_t)k.ld); mix((uint64_t)k.plane);
        mix(((uint64_t)k.box_rows << 12) | ((uint64_t)k.box_planes << 8) | ((uint64_t)k.nplanes << 1) | (uint64_t)k.store);
        return (size_t)h;
    }
};
int make_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, long long plane,
             int box_rows, int nplanes, bool store_map, int box_planes) {
    if (box_planes <= 0) box_planes = nplanes;
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    const MapKey key{base, rows, cols, ld, plane, box_rows, nplanes, store_map ? 1 : 0, box_planes};
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(key);
        if (it != cache.end()) { *map = it->second; return AVR_OK; }
    }
    if (int rc = encode_map(map, base, rows, cols, ld, plane, box_rows, nplanes, store_map, box_planes)) return rc;
    std::lock_guard<std::mutex> lock(mu);
    if (cache.size() >= 4096) cache.clear();
    cache.emplace(key, *map);
    return AVR_OK;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is per (function, device) state: raised to the 227 KB opt-in limit once
template <bool MN_MAJOR, bool PAIR>
static int allow_max_smem(int device) {
    static std::atomic<uint64_t> done{0};                        // one flag word per kernel instantiation
    auto kernel = umma_gemm_kernel<MN_MAJOR, PAIR>;
    const uint64_t bit = 1ull << (device & 63);
    if (done.load(std::memory_order_acquire) & bit) return AVR_OK;
    AVR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    done.fetch_or(bit, std::memory_order_release);
    return AVR_OK;
}

static int pick_bn(long long N, int max_bn = 256) {
    if (N <= max_bn) return (int)(ceil_div(N, 16) * 16);
    if (N % max_bn == 0) return max_bn;
    if (max_bn == 256 && N % 128 == 0) return 128;
    // e.g. 1600: prefer the multiple of 16 <= max_bn that wastes least
    int best = max_bn;
    long long best_waste = ceil_div(N, max_bn) * max_bn - N;
    for (int bn = max_bn - 32; bn >= max_bn / 2; bn -= 32) {      // multiples of 32: staged epilogue chunks never straddle tiles
        long long waste = ceil_div(N, bn) * bn - N;
        if (waste < best_waste) { best = bn; best_waste = waste; }
    }
    return best;
}

static int tmem_cols_for(int bn, int dual = 0, int bufs = 2) {
    int need = (dual ? 2 : 1) * bufs * bn, c = 32;
    while (c < need) c <<= 1;
    return c;
}

// How many 2-CTA clusters of the K-major kernel can be resident at once (one CTA per SM; a GPC with an odd number of
// free SMs leaves one idle).  Asked once per (device, dynamic shared memory size).
template <bool MN_MAJOR>
static int max_resident_pairs_of(int device) {
    static std::mutex mu;
    static std::unordered_map<uint64_t, int> cache;
    const size_t smem = 232448;
    const uint64_t key = (uint64_t)device;
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(key);
        if (it != cache.end()) return it->second;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2u * 148u);
    cfg.blockDim = dim3(UTHREADS);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    if (allow_max_smem<MN_MAJOR, true>(device) != AVR_OK) return 0;
    if (cudaOccupancyMaxActiveClusters(&n, umma_gemm_kernel<MN_MAJOR, true>, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
    std::lock_guard<std::mutex> lock(mu);
    cache[key] = n;
    return n;
}

static int max_resident_pairs(int device) { return max_resident_pairs_of<false>(device); }

// which launches run as CTA pairs: every wide K-major product over many row tiles.  Measured on the 524800 x 512 x 512
// layers of the simu signal network (profiles/r2/pair_probe.jsonl): fp16 pairs 0.82 -> 0.66 ms, with the second bf16 copy
// of the output 1.16 -> 0.80 ms, bf16 triples at K = 208 0.93 -> 0.83 ms, masked bf16-pair backward products 0.66 -> 0.60 ms.
static bool pair_mode_wanted(int a_f16, int a_nplanes) {
    (void)a_f16; (void)a_nplanes;
    if (const char* e = AVR_EXP_ENV("AVR_UMMA_CLUSTER")) return atoi(e) != 0;
    return true;
}

int num_sms(int device) {
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
    return n > 0 ? n : 148;
}

}  // namespace avr

using namespace avr;

extern "C" {

// x fp32 [rows, cols] (ld) -> plane pair; transpose != 0 writes out[c, r]
AVR_API int avr_planes_split(const float* x, int64_t rows, int64_t cols, int64_t ld, void* planes, int64_t ldp,
                             int64_t plane_stride, int nplanes, int transpose, int relu, int device, void* stream) {
    AVR_REQUIRE(planes && (x || rows * cols == 0), "null pointer");
    AVR_REQUIRE(planes_kind_ok(nplanes), "unknown plane-set kind");
    AVR_ENTER(device);
    if (rows * cols == 0) return AVR_OK;
    planes_split_kernel<<<(unsigned)ceil_div(rows * cols, 256), 256, 0, (cudaStream_t)stream>>>(
        x, rows, cols, ld, planes, ldp, plane_stride, nplanes, transpose, relu);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

AVR_API int avr_planes_split_batch(const avr_split_desc* descs, int32_t n, int device, void* stream) {
    AVR_REQUIRE(descs || n == 0, "null pointer");
    AVR_REQUIRE(n >= 0 && n <= AVR_SPLIT_BATCH_MAX, "too many matrices for one batch");
    AVR_ENTER(device);
    if (n == 0) return AVR_OK;
    SplitBatch b = {};
    long long most = 0;
    for (int k = 0; k < n; ++k) {
        AVR_REQUIRE(descs[k].x && descs[k].planes && planes_kind_ok(descs[k].kind) && descs[k].rows >= 0 && descs[k].cols >= 0,
                    "bad matrix descriptor");
        b.d[k] = descs[k];
        const long long e = (long long)descs[k].rows * descs[k].cols;
        most = e > most ? e : most;
    }
    if (most == 0) return AVR_OK;
    planes_split_batch_kernel<<<dim3((unsigned)ceil_div(most, 256), (unsigned)n), 256, 0, (cudaStream_t)stream>>>(b);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

AVR_API int avr_planes_merge(const void* planes, int64_t rows, int64_t cols, int64_t ldp, int64_t plane_stride, int nplanes,
                             float* out, int64_t ld, int device, void* stream) {
    AVR_REQUIRE(planes && out, "null pointer");
    AVR_REQUIRE(planes_kind_ok(nplanes), "unknown plane-set kind");
    AVR_ENTER(device);
    if (rows * cols == 0) return AVR_OK;
    planes_merge_kernel<<<(unsigned)ceil_div(rows * cols, 256), 256, 0, (cudaStream_t)stream>>>(
        planes, rows, cols, ldp, plane_stride, nplanes, out, ld);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

// number of interleaved K slices of the split-K mode of avr_umma_gemm_nt: 256-deep slices (1 when K <= 512)
AVR_API int64_t avr_umma_gemm_nt_splitk_slices(int64_t K) { return K <= 512 ? 1 : ceil_div(K, 256); }

// C[M,N] = A[M,K] . B[N,K]^T on plane pairs (both K-major).  Output: plane pair (default) or fp32 (UF_OUT_F32).
AVR_API int avr_umma_gemm_nt(int64_t M, int64_t N, int64_t K, const void* a_planes, int64_t lda, int64_t a_plane,
                             int a_nplanes, const void* b_planes, int64_t ldb, int64_t b_plane, int b_nplanes, int flags,
                             void* c_planes, int64_t ldc, int64_t c_plane, int c_nplanes, void* c2_planes, int64_t ldc2,
                             int64_t c2_plane, int c2_nplanes, const uint32_t* mask_bits, int64_t ldmask, uint32_t* bits_out,
                             int64_t ldbits, const float* bias_ray, int64_t ld_bias_ray, const float* bias_rcv,
                             int64_t ld_bias_rcv, int32_t geo_R, int32_t geo_S, float* c_f32, int64_t ldc32,
                             uint32_t* near_list, int64_t near_cap, uint32_t* near_count, float near_tau,
                             void* splitk_workspace, int64_t splitk_workspace_bytes, int device, void* stream) {
    AVR_REQUIRE(a_planes && b_planes, "null operand");
    if (near_list) {
        AVR_REQUIRE(near_count && near_cap > 0 && near_cap < (1ll << 31) && near_tau >= 0.f && aligned16(near_list),
                    "near-zero guard needs a list, its capacity and a zeroed counter");
        AVR_REQUIRE(!(flags & (UF_MASK | UF_ACCUM)), "the near-zero guard applies to forward layers (no MASK / ACCUM)");
    }
    AVR_REQUIRE(planes_kind_ok(a_nplanes) && planes_kind_ok(b_nplanes) && planes_kind_ok(c_nplanes), "unknown plane-set kind");
    const int a_f16 = planes_f16(a_nplanes), b_f16 = planes_f16(b_nplanes), c_kind = c_nplanes;
    AVR_REQUIRE(a_f16 == b_f16, "A and B must both be bf16 plane sets or both fp16 pairs");
    AVR_REQUIRE(!near_list || (!a_f16 && !planes_f16(c_kind)), "the near-zero guard works on bf16 plane sets");
    a_nplanes = planes_count(a_nplanes); b_nplanes = planes_count(b_nplanes); c_nplanes = planes_count(c_nplanes);
    if (!(a_nplanes == 3 && b_nplanes == 3)) a_nplanes = b_nplanes = 2;     // six products need 24 bits on both sides
    AVR_REQUIRE(M >= 0 && N > 0 && K > 0, "bad dimensions");
    AVR_REQUIRE((flags & UF_OUT_F32) ? N % 8 == 0 : N % 4 == 0, "N must be a multiple of 8 (fp32 output) / 4 (plane output)");
    AVR_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "dimension overflow");
    if (flags & UF_OUT_F32) AVR_REQUIRE(c_f32 && ldc32 % 4 == 0 && aligned16(c_f32), "fp32 output must be 16-byte aligned");
    else AVR_REQUIRE(c_planes && ldc % 8 == 0 && c_plane % 8 == 0 && aligned16(c_planes), "plane output must be 16-byte aligned");
    if (flags & UF_MASK) AVR_REQUIRE(mask_bits && ldmask * 32 >= N, "MASK needs a bitmask with >= N/32 words per row");
    if (flags & UF_BITS) AVR_REQUIRE(bits_out && ldbits * 32 >= N && !(flags & UF_OUT_F32), "BITS needs a bitmask output");
    const bool dual = (flags & (UF_DUAL_RELU | UF_DUAL_COPY)) != 0;
    AVR_REQUIRE((flags & (UF_DUAL_RELU | UF_DUAL_COPY)) != (UF_DUAL_RELU | UF_DUAL_COPY), "DUAL_RELU and DUAL_COPY exclude each other");
    if (dual) AVR_REQUIRE(c2_planes && ldc2 % 8 == 0 && c2_plane % 8 == 0 && aligned16(c2_planes) && planes_kind_ok(c2_nplanes) &&
                          !(flags & (UF_OUT_F32 | UF_ACCUM)), "second output missing / misaligned / unknown kind");
    if (!dual) c2_nplanes = c_kind;
    AVR_ENTER(device);
    if (M == 0) return AVR_OK;
    UmmaParams p = {};
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.na = a_nplanes; p.nb = b_nplanes; p.nc = c_nplanes;
    p.fa = a_f16; p.fb = b_f16; p.kc = c_kind; p.kc2 = c2_nplanes;
    p.mode = a_f16 ? 2 : (a_nplanes == 3 ? 1 : 0);
    p.small_scale = a_f16 ? F16_LO_INV : 1.0f;
    p.BN = pick_bn(N, (a_nplanes == 3 || a_f16) ? 128 : 256);                // two accumulators per tile: 4 * BN <= 512
    p.acc_bufs = 2;
    // The 24-bit forward modes keep two accumulators per tile, so a double-buffered tile is at most 128 columns wide --
    // and a 128 x 128 x 16 MMA reads 8 KB of operands from shared memory in the ~64 clk it occupies the tensor pipe: the
    // main loop is bound by the shared-memory operand bandwidth (128 B/clk), not by the pipe or by L2 (measured: the
    // 524800 x 512 x 512 fp16-pair layer WITHOUT epilogue takes 0.80-0.92 ms in 128-column tiles against 0.57 ms for the same
    // number of 256-column MMAs; sharing the B tile between two CTAs by TMA multicast, which cuts the L2 traffic by a
    // quarter, changes nothing).  CTA pairs (cta_group::2) multiply two adjacent row tiles by one column tile as ONE
    // 256-row MMA: each SM reads its 128 rows of A and only HALF of the B tile (6 KB per MMA), the pipe is the bound
    // again, and tiles stay double-buffered in tensor memory.
    if (int rc = allow_max_smem<false, false>(device)) return rc;
    p.cluster = 1;
    int pairs = 0;
    // (a single column tile whose whole B operand fits in shared memory keeps the B-resident single-CTA schedule)
    const bool b_fits = N <= p.BN && (uint64_t)ceil_div(K, UBK) * (uint64_t)b_nplanes * (uint64_t)p.BN * 128u +
                                         2ull * (uint64_t)a_nplanes * A_PLANE_BYTES <= 232448u - 2048u - 65536u;
    if (!splitk_workspace && !(flags & UF_OUT_F32) && !b_fits && p.BN % 16 == 0 && M >= 2ll * UM * num_sms(device) &&
        pair_mode_wanted(a_f16, a_nplanes)) {
        pairs = max_resident_pairs(device);
        if (pairs > 0) p.cluster = 2;
    }
    // Without pairs (no cluster launch possible): fp16 pairs in 128 x 256 tiles, whose two accumulators fill tensor memory
    // (no overlap of a tile's epilogue with the next tile's MMAs), but whose MMAs are not operand-bound.
    if (p.cluster == 1 && a_f16 && N % 256 == 0 && K >= 256 && !(flags & UF_BIAS) && !AVR_EXP_ENV("AVR_UMMA_F16_BN128")) { p.BN = 256; p.acc_bufs = 1; }
    p.tiles_m = (int)ceil_div(M, UM); p.tiles_n = (int)ceil_div(N, p.BN);
    p.k_splits = 1; p.k_per_split = (int)(ceil_div(K, UBK) * UBK);
    // long reductions into an fp32 output (the DFT and its adjoint, K = T or 2F): the accumulator is truncated once per
    // MMA, a bias that grows with the chain length.  Short interleaved K slices into separate accumulators, summed in
    // fp32 by the deterministic reduce below, cut it by ~(number of slices)^1.5.
    const int64_t splitk_ld = ceil_div(N, 8) * 8;
    if (splitk_workspace) {
        AVR_REQUIRE((flags & UF_OUT_F32) && !(flags & (UF_RELU | UF_ACCUM)) && !near_list && aligned16(splitk_workspace),
                    "split-K applies to plain fp32 outputs");
        const int64_t splits = avr_umma_gemm_nt_splitk_slices(K);
        AVR_REQUIRE(splitk_workspace_bytes >= splits * M * splitk_ld * (int64_t)sizeof(float), "split-K workspace too small");
        p.k_splits = (int)splits;
    }
    p.dual_acc = (a_f16 || (p.na == 3 && p.nb == 3 && 4 * p.BN <= 512 && !AVR_EXP_ENV("AVR_UMMA_SINGLE_ACC"))) ? 1 : 0;
    p.tmem_cols = tmem_cols_for(p.BN, p.dual_acc, p.acc_bufs);
    p.flags = flags & ~(UF_DEBUG_NOWAIT | UF_DEBUG_NOSTORE | UF_DEBUG_NOSTAGE | UF_DEBUG_NOFENCE);
    p.c = (__nv_bfloat16*)c_planes; p.ldc = ldc; p.c_plane = c_plane;
    p.c2 = (__nv_bfloat16*)c2_planes; p.ldc2 = ldc2; p.c2_plane = c2_plane;
    p.mask = mask_bits; p.ldmask = ldmask;
    p.bits = bits_out; p.ldbits = ldbits;
    if (flags & UF_BIAS) {
        AVR_REQUIRE((bias_ray || bias_rcv) && geo_R > 0 && geo_S > 0 && !(flags & UF_OUT_F32), "BIAS needs a table and the ray geometry");
        AVR_REQUIRE((!bias_ray || (aligned16(bias_ray) && ld_bias_ray % 4 == 0)) && (!bias_rcv || (aligned16(bias_rcv) && ld_bias_rcv % 4 == 0)),
                    "bias tables must be 16-byte aligned");
    }
    p.bias_ray = bias_ray; p.bias_rcv = bias_rcv; p.ld_bias_ray = ld_bias_ray; p.ld_bias_rcv = ld_bias_rcv;
    p.geo_R = geo_R; p.geo_S = geo_S;
    if (AVR_EXP_ENV("AVR_UMMA_NOWAIT")) p.flags |= UF_DEBUG_NOWAIT;       // timing experiments only: results are garbage
    if (const char* dbg = AVR_EXP_ENV("AVR_UMMA_DEBUG")) p.flags |= (atoi(dbg) & (UF_DEBUG_NOWAIT | UF_DEBUG_NOSTORE | UF_DEBUG_NOSTAGE | UF_DEBUG_NOFENCE));
    p.c32 = c_f32; p.ldc32 = ldc32;
    if (p.k_splits > 1) { p.c32 = (float*)splitk_workspace; p.ldc32 = splitk_ld; }
    p.near_list = (uint2*)near_list; p.near_cap = (uint32_t)near_cap; p.near_count = near_count; p.near_tau = near_tau;
    p.a_raw = (const __nv_bfloat16*)a_planes; p.lda = lda; p.a_plane = a_plane;
    p.b_raw = (const __nv_bfloat16*)b_planes; p.ldb = ldb; p.b_plane = b_plane;
    const uint32_t a_tile = (uint32_t)p.na * A_PLANE_BYTES;
    const uint32_t b_tile = (uint32_t)p.nb * (uint32_t)(p.cluster == 2 ? p.BN / 2 : p.BN) * 128u;    // pair mode: this CTA's half
    p.nkb = (int)ceil_div(K, UBK);
    const int nc2 = planes_count(c2_nplanes);
    p.epi_warp_bytes = (flags & UF_OUT_F32) ? 0u : (uint32_t)(p.nc > nc2 ? p.nc : nc2) * EPI_PLANE_BYTES;
    const uint32_t bias_bytes = (flags & UF_BIAS) ? 4u * 1024u : 0u;
    uint32_t epi_bytes = 0, bres_total = 0, stage_bytes = 0;
    // two epilogue warps per lane group when their staging tiles fit next to a two-stage pipeline (two-plane outputs,
    // fp32 outputs); the per-receiver bias rows are staged per lane group, so that mode keeps one warp per group
    // ... and two staging buffers per warp when three pipeline stages still fit next to them
    const int cand[4][2] = {{2, 2}, {2, 1}, {1, 2}, {1, 1}};
    const bool one_buf = AVR_EXP_ENV("AVR_UMMA_EPI_BUFS") && atoi(AVR_EXP_ENV("AVR_UMMA_EPI_BUFS")) == 1;
    for (int ci = (flags & UF_BIAS) ? 2 : 0; ci < 4; ++ci) {
        p.epi_split = cand[ci][0]; p.epi_bufs = cand[ci][1];
        if (one_buf && p.epi_bufs == 2) continue;
        // 227 KB per CTA = alignment slack (1 KB) + operands + barrier block padded to 1 KB + staging (+ bias rows)
        epi_bytes = 4u * (uint32_t)(p.epi_split * p.epi_bufs) * p.epi_warp_bytes + bias_bytes;
        const uint32_t budget = 232448u - 2048u - epi_bytes;
        p.b_resident = (p.cluster == 1 && p.tiles_n == 1 && p.tiles_m > 1 && (uint64_t)p.nkb * b_tile + 2ull * a_tile <= budget) ? 1 : 0;
        bres_total = p.b_resident ? (uint32_t)p.nkb * b_tile : 0u;
        stage_bytes = p.b_resident ? a_tile : a_tile + b_tile;
        p.stages = (int)((budget - bres_total) / stage_bytes);
        if (p.stages > 6) p.stages = 6;
        if (p.stages >= (p.epi_bufs == 2 ? 3 : 2)) break;
    }
    if (p.stages < 2) return fail(AVR_ERR_UNSUPPORTED, "tile does not fit two pipeline stages");
    if (const char* e = AVR_EXP_ENV("AVR_UMMA_EPI_SPLIT")) {                   // A/B experiments: force one warp per lane group
        if (atoi(e) == 1 && p.epi_split == 2) p.epi_split = 1;
    }
    const size_t smem = (size_t)bres_total + (size_t)p.stages * stage_bytes + 2048 + epi_bytes;
    CUtensorMap ta, tb, tc, tc2;
    if (int rc = make_map(&ta, a_planes, M, K, lda, a_plane, UM, p.na)) return rc;
    if (int rc = make_map(&tb, b_planes, N, K, ldb, b_plane, p.BN, p.nb)) return rc;
    tc = ta; tc2 = ta;
    if (!(flags & UF_OUT_F32)) {
        if (int rc = make_map(&tc, c_planes, M, N, ldc, c_plane, 32, p.nc, true)) return rc;
        tc2 = tc;
        if (dual)
            if (int rc = make_map(&tc2, c2_planes, M, N, ldc2, c2_plane, 32, nc2, true)) return rc;
    }
    CUtensorMap tbh = tb;
    if (p.cluster == 2) {
        if (int rc = make_map(&tbh, b_planes, N, K, ldb, b_plane, p.BN / 2, p.nb)) return rc;    // this CTA's half of a column tile

        const int units = ((p.tiles_m + 1) / 2) * p.tiles_n;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2u * (unsigned)(units < pairs ? units : pairs));
        cfg.blockDim = dim3(UTHREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        AVR_CUDA(cudaLaunchKernelEx(&cfg, umma_gemm_kernel<false, true>, ta, tb, tc, tc2, tbh, p));
    } else {
        const int tiles = p.tiles_m * p.tiles_n * p.k_splits;
        const int grid = tiles < num_sms(device) ? tiles : num_sms(device);
        umma_gemm_kernel<false, false><<<grid, UTHREADS, smem, (cudaStream_t)stream>>>(ta, tb, tc, tc2, tbh, p);
    }
    AVR_LAUNCH_CHECK();
    if (p.k_splits > 1) {
        umma_splitk_reduce_kernel<<<(unsigned)ceil_div(M * (N / 4) * 32, 256), 256, 0, (cudaStream_t)stream>>>(
            (const float*)splitk_workspace, splitk_ld, p.k_splits, M, N, c_f32, ldc32, 0);
        AVR_LAUNCH_CHECK();
    }
    if (near_list) {
        umma_fixup_kernel<<<num_sms(device), 256, 0, (cudaStream_t)stream>>>(p);
        AVR_LAUNCH_CHECK();
    }
    return AVR_OK;
}

// split-K slices of the weight-gradient GEMM for a given tile width
static int64_t tn_splits(int64_t M, int64_t N, int64_t K, int bn) {
    const int64_t tiles = ceil_div(M, UM) * ceil_div(N, bn);
    int64_t splits = 148 / tiles;                  // one wave: tiles * splits <= number of SMs
    const int64_t max_by_k = ceil_div(K, 1024);
    if (splits > max_by_k) splits = max_by_k;
    return splits < 1 ? 1 : splits;
}

AVR_API int64_t avr_umma_gemm_tn_workspace_bytes(int64_t M, int64_t N, int64_t K) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    const int64_t s2 = tn_splits(M, N, K, pick_bn(N)), s3 = tn_splits(M, N, K, pick_bn(N, 128));   // either plane count
    return (s2 > s3 ? s2 : s3) * M * (ceil_div(N, 8) * 8) * (int64_t)sizeof(float);
}

// C[M,N] (+)= sum_k A[k,M] * B[k,N] on plane pairs stored [K, M] and [K, N] (both MN-major); fp32 output,
// deterministic split-K over k (the sample points).
AVR_API int avr_umma_gemm_tn(int64_t M, int64_t N, int64_t K, const void* a_planes, int64_t lda, int64_t a_plane,
                             int a_kind, const void* b_planes, int64_t ldb, int64_t b_plane, int b_kind, float* c,
                             int64_t ldc, int accumulate, void* workspace, int64_t workspace_bytes, int device,
                             void* stream) {
    AVR_REQUIRE(a_planes && b_planes && c && workspace, "null pointer");
    AVR_REQUIRE(planes_kind_ok(a_kind) && planes_kind_ok(b_kind) && !planes_f16(a_kind),
                "the gradient operand is a bf16 plane set (gradients need bf16's range)");
    // an fp16-pair activation is converted to a bf16 (hi, mid) pair in shared memory (conv_b): the MMAs see bf16 on both sides
    const int conv_b = planes_f16(b_kind) ? 1 : 0;
    const int b_f16 = 0;
    int na = planes_count(a_kind), nb = planes_count(b_kind);
    if (!b_f16 && !(na == 3 && nb == 3)) na = nb = 2;                       // bf16 . bf16: six products need 24 bits on both sides
    const int nplanes = (na == 3 || b_f16) ? 3 : 2;                         // "wide" mode: 128-column tiles, two accumulators
    AVR_REQUIRE(M > 0 && N > 0 && K >= 0, "bad dimensions");
    AVR_REQUIRE(M % 4 == 0 && N % 4 == 0, "M and N must be multiples of 4");
    AVR_REQUIRE(K < (1ll << 31), "dimension overflow");
    AVR_ENTER(device);
    UmmaParams p = {};
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.BN = pick_bn(N, nplanes == 3 ? 128 : 256);                // two stages and two accumulators must fit
    p.na = na; p.nb = nb; p.nc = 2;
    p.fa = 0; p.fb = b_f16; p.kc = p.kc2 = AVR_PLANES_BF16x2;
    p.conv_b = conv_b;
    p.acc_bufs = 2;
    p.mode = b_f16 ? (na == 3 ? 4 : 3) : (na == 3 ? 1 : 0);
    p.small_scale = b_f16 ? F16_LO_INV : 1.0f;
    p.tiles_m = (int)ceil_div(M, UM); p.tiles_n = (int)ceil_div(N, p.BN);
    const int64_t ldp = ceil_div(N, 8) * 8;                    // partial rows padded to whole 8-column store groups
    const int64_t splits = K > 0 ? tn_splits(M, N, K, p.BN) : 1;
    AVR_REQUIRE(workspace_bytes >= splits * M * ldp * (int64_t)sizeof(float) && aligned16(workspace),
                "workspace too small or misaligned");
    p.k_per_split = (int)(ceil_div(ceil_div(K, splits), UBK) * UBK);
    if (p.k_per_split < UBK) p.k_per_split = UBK;
    p.k_splits = (int)(K > 0 ? ceil_div(K, p.k_per_split) : 1);
    p.dual_acc = (b_f16 || (nplanes == 3 && 4 * p.BN <= 512 && !AVR_EXP_ENV("AVR_UMMA_SINGLE_ACC"))) ? 1 : 0;
    p.tmem_cols = tmem_cols_for(p.BN, p.dual_acc);
    p.c32 = (float*)workspace; p.ldc32 = ldp;
    p.epi_split = 2; p.epi_warp_bytes = 0; p.epi_bufs = 1;
    if (const char* e = AVR_EXP_ENV("AVR_UMMA_EPI_SPLIT_TN")) p.epi_split = atoi(e) == 1 ? 1 : 2;
    // CTA pairs (cta_group::2) for gradients of at least two row tiles: one 256-row MMA per pair, each CTA stages its 128
    // columns of A and half of the B columns (a third pipeline stage fits).  The split over k is the single-CTA schedule's,
    // so the partial sums -- and the result -- are bit-identical to it.
    p.cluster = 1;
    int pairs = 0;
    if (K >= 4096 && p.tiles_m >= 2 && p.BN % 32 == 0 && pair_mode_wanted(0, na)) {
        pairs = max_resident_pairs_of<true>(device);
        if (pairs > 0) p.cluster = 2;
    }
    const int bn_rows = ((p.cluster == 2 ? p.BN / 2 : p.BN) + 63) / 64 * 64;
    const uint32_t stage_bytes = (uint32_t)na * A_PLANE_BYTES + (uint32_t)nb * (uint32_t)bn_rows * 128u;
    p.stages = (int)((220 * 1024) / stage_bytes);
    if (p.stages > 6) p.stages = 6;
    if (p.stages < 2) return fail(AVR_ERR_UNSUPPORTED, "tile does not fit two pipeline stages");
    const size_t smem = (size_t)p.stages * stage_bytes + 1024 + 256 + 2048;
    cudaStream_t st = (cudaStream_t)stream;
    if (K > 0) {
        CUtensorMap ta, tb;
        if (int rc = make_map(&ta, a_planes, K, M, lda, a_plane, 64, na)) return rc;
        if (int rc = make_map(&tb, b_planes, K, N, ldb, b_plane, 64, nb)) return rc;
        if (p.cluster == 2) {
            const int items = ((p.tiles_m + 1) / 2) * p.tiles_n * p.k_splits;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(2u * (unsigned)(items < pairs ? items : pairs));
            cfg.blockDim = dim3(UTHREADS);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            AVR_CUDA(cudaLaunchKernelEx(&cfg, umma_gemm_kernel<true, true>, ta, tb, ta, ta, ta, p));
        } else {
            if (int rc = allow_max_smem<true, false>(device)) return rc;
            const int tiles = p.tiles_m * p.tiles_n * p.k_splits;
            const int grid = tiles < num_sms(device) ? tiles : num_sms(device);
            umma_gemm_kernel<true, false><<<grid, UTHREADS, smem, st>>>(ta, tb, ta, ta, ta, p);
        }
        AVR_LAUNCH_CHECK();
    } else {
        p.k_splits = 0;
    }
    AVR_REQUIRE(ldc % 4 == 0 && aligned16(c), "C must be 16-byte aligned");
    umma_splitk_reduce_kernel<<<(unsigned)ceil_div(M * (N / 4) * 32, 256), 256, 0, st>>>((const float*)workspace, ldp, p.k_splits,
                                                                                       M, N, c, ldc, accumulate);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

}  // extern "C"
