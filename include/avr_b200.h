/*
 * avr_b200 -- C-ABI of the B200-native (sm_100a) AVR acoustic volume-rendering hot path.
 *
 * This is the drop-in boundary.  The reference has no native code of
 * its own: the functions below replace what `renderer.py` / `model.py` execute inside
 * tiny-cuda-nn and ATen.  Each entry cites the reference lines it replaces
 * (paths relative to /root/reference).
 *
 * Conventions
 *   - every function returns 0 on success, else an AVR_ERR_* / cudaError_t value;
 *     avr_last_error() gives a thread-local message.  Nothing throws, nothing exits.
 *   - everything is asynchronous on `stream` (a cudaStream_t passed as void*) of CUDA
 *     device `device`; no call synchronises, allocates or retains pointers.
 *   - all buffers are device pointers owned by the caller (PyTorch allocates outputs and
 *     workspaces); fp32 unless said otherwise; matrices row-major with an explicit leading
 *     dimension (`ld*`, in elements) so column blocks of wider buffers can be addressed.
 *   - point / row order is the reference's: n = (b*R + r)*S + s   (renderer.py:55-58).
 *   - no global mutable state: safe under concurrent calls from several host threads
 *     (nn.DataParallel, the autograd engine's device threads).
 */
#ifndef AVR_B200_H
#define AVR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AVR_B200_ABI_VERSION 2
#if defined(__GNUC__)
#define AVR_API __attribute__((visibility("default")))
#else
#define AVR_API
#endif
#define AVR_MAX_LEVELS 32

enum {
    AVR_OK = 0,
    AVR_ERR_INVALID = 10001,      /* bad argument (null pointer, misaligned, bad shape) */
    AVR_ERR_UNSUPPORTED = 10002   /* valid request this build does not implement */
};

/* Render geometry: the per-(b,r,s) arithmetic of renderer.py:54-62,86-87 / renderer_cpu.py:46-52,76-77. */
typedef struct avr_render_geom {
    int32_t bs, R, S, T;   /* receivers, rays (n_azi*n_ele+2), samples per ray, IR length */
    float xyz_min;         /* (float)xyz_min                                   */
    float xyz_span;        /* (float)((double)xyz_max - (double)xyz_min)       */
    float fs, speed;       /* (float)fs, (float)speed                          */
} avr_render_geom;

/* Multiresolution hash grid (tiny-cuda-nn "HashGrid", 3-D input, linear interpolation,
 * coherent-prime hash) -- what tcnn.Encoding(3, cfg) at model.py:66-68,258-264 builds. */
typedef struct avr_grid_meta {
    int32_t n_levels, n_feat;             /* n_feat must be 2 in this build */
    float scale[AVR_MAX_LEVELS];          /* exp2f(l*log2(pls))*base - 1                 */
    uint32_t res[AVR_MAX_LEVELS];         /* ceil(scale)+1                               */
    uint32_t size[AVR_MAX_LEVELS];        /* entries in level l                          */
    uint32_t offset[AVR_MAX_LEVELS];      /* first entry of level l in the flat table    */
    uint32_t total;                       /* sum(size)                                   */
    int32_t stride32;                     /* 1: tiny-cuda-nn's grid_index arithmetic -- the dense-index stride is a uint32
                                           * that wraps (levels with 2^16 <= res <= size are then NOT hashed:
                                           * index = (x + y*res mod 2^32) % size); 0: exact (64-bit) stride, such levels
                                           * are hashed.  See DESIGN.md 2 "grid_index". */
} avr_grid_meta;

/* Epilogue / operand flags of avr_gemm */
enum {
    AVR_GEMM_RELU = 1,        /* C = max(C, 0)                                             */
    AVR_GEMM_ACCUM = 2,       /* C += product (otherwise C = product)                      */
    AVR_GEMM_MASK = 4,        /* product *= (aux[i,j] > 0)  (ReLU backward)                */
    AVR_GEMM_RELU_A = 8,      /* A := max(A, 0) on load (consumer-side ReLU)               */
    AVR_GEMM_RELU_B = 16      /* B := max(B, 0) on load                                    */
};
/* Operand layouts of avr_gemm: element (i,k) of an operand P with leading dimension ld */
enum {
    AVR_K_CONTIG = 0,         /* P[i*ld + k]  (reduction index contiguous)                  */
    AVR_I_CONTIG = 1          /* P[k*ld + i]  (output index contiguous)                     */
};

AVR_API int avr_abi_version(void);
AVR_API const char* avr_last_error(void);
/* number of kernels launched by this library (process-wide) since the last reset */
AVR_API int64_t avr_launch_count(int reset);

/* ---- ray generation + sampling (+ delays) -------------------------------------------------
 * renderer.py:54-62 (ray_pts, normalize_points, network_view/tx) and :86-87 (tx->point delay).
 * Writes the network inputs the reference would materialise -- used by the generic
 * `networks_fn` path and by the bit-exactness tests.  Any output pointer may be NULL.
 *   rays_o[bs,3], pos_tx[bs,3], dirs[R,3], d_vals[S]
 *   pts_n[bs,P,3], view[bs,P,3], tx_n[bs,P,3]  (normalised to [-1,1]; P = R*S)
 *   delay[bs,R,S] int32 in [0,T-1] */
AVR_API int avr_sample_points(const avr_render_geom* geom, const float* rays_o, const float* pos_tx,
                      const float* dirs, const float* d_vals, float* pts_n, float* view, float* tx_n,
                      int32_t* delay, int device, void* stream);

/* Unit-cube inputs of the per-ray / per-receiver encodings (renderer.py:59-60 + model.py:188-189,
 * 309-311): u_view[R,3] = (-dirs+1)/2, u_tx[bs,3] = (normalize(pos_tx)+1)/2,
 * u_dir_tx[bs,3] = (dir_tx+1)/2.  dir_tx and any output may be NULL. */
AVR_API int avr_aux_inputs(const avr_render_geom* geom, const float* pos_tx, const float* dirs, const float* dir_tx,
                   float* u_view, float* u_tx, float* u_dir_tx, int device, void* stream);

/* ---- fused ray generation + sampling + hash-grid encode --------------------------------------
 * renderer.py:54-58 + model.py:187,191 (tcnn kernel_grid forward) in one kernel: sample positions
 * are produced in registers and encoded; nothing of shape [bs,P,3] reaches HBM.
 *   table[total*2]; out rows n=(b,r,s): out[n*ld_out + col0 + 2*l + f];
 *   columns [col0+2L, col0+2L+n_ones) are set to 1 (tcnn input padding);
 *   delay (optional, int32[bs,R,S]) as in avr_sample_points.
 *   out_plane == 0: `out` is fp32.  out_plane > 0: `out` is a bf16 plane set (see "dense layers on the
 *   tensor cores") of out_nplanes (2 or 3) planes with that plane stride, ld_out in bf16 elements; readers
 *   of plane sets (d_out, d_dst, x) use the first two planes.  The same convention holds for the
 *   d_out / x / dst / d_dst arguments below that carry a `*_plane` companion. */
AVR_API int avr_raygen_encode_fwd(const avr_render_geom* geom, const avr_grid_meta* grid, const float* rays_o,
                          const float* pos_tx, const float* dirs, const float* d_vals, const float* table,
                          void* out, int64_t ld_out, int64_t out_plane, int32_t out_nplanes, int32_t col0,
                          int32_t n_ones, int32_t* delay, int device, void* stream);

/* Backward of the above w.r.t. the table (tcnn kernel_grid_backward).  Two accumulation modes:
 *   log2_headroom in [0,56]  DETERMINISTIC: contributions are accumulated as 2^e-scaled int64 (integer addition is
 *     associative), `e` derived on the device from absmax(d_out) and `log2_headroom` >= log2(max contributions
 *     per entry).  acc[total*2] is int64, zeroed by the caller (or holding a previous partial sum taken with the
 *     same gmax_bits); gmax_bits is a device uint32 filled by avr_absmax_bits; avr_grid_grad_finalize converts.
 *   log2_headroom == AVR_GRID_GRAD_F32  acc[total*2] is the fp32 gradient itself (zeroed, or a partial sum); one
 *     red.global.add.v2.f32 per cell corner (half the reductions of the int64 mode, summation order -- and
 *     so the last bits -- vary from run to run, as in tcnn); gmax_bits is ignored and may be NULL.
 * sample_step (ray-generation entry): distance between consecutive samples of a ray in unit-cube
 * coordinates, (far - near) / (S - 1) / (xyz_max - xyz_min), or 0 if unknown.  On the leading levels whose cells hold
 * two or more consecutive samples, runs of adjacent points that add to the same entry are summed with warp shuffles
 * and issue one reduction. */
#define AVR_GRID_GRAD_F32 (-1)
AVR_API int avr_raygen_encode_bwd(const avr_render_geom* geom, const avr_grid_meta* grid, const float* rays_o,
                          const float* dirs, const float* d_vals, const void* d_out, int64_t ld_out,
                          int64_t d_plane, int32_t col0, const uint32_t* gmax_bits, int32_t log2_headroom, void* acc,
                          float sample_step, int device, void* stream);

/* Encode explicit unit-cube points u[N,3] (model.py:191,219-220 on arbitrary inputs). */
AVR_API int avr_grid_encode_fwd(const avr_grid_meta* grid, const float* u, int64_t n_pts, const float* table,
                        void* out, int64_t ld_out, int64_t out_plane, int32_t out_nplanes, int32_t col0, int32_t n_ones,
                        int device, void* stream);
AVR_API int avr_grid_encode_bwd(const avr_grid_meta* grid, const float* u, int64_t n_pts, const void* d_out,
                        int64_t ld_out, int64_t d_plane, int32_t col0, const uint32_t* gmax_bits, int32_t log2_headroom,
                        void* acc, int device, void* stream);
/* gmax_bits = max(gmax_bits, bit pattern of max |x[i, col0:col0+ncols]|)  (caller zeroes it first) */
AVR_API int avr_absmax_bits(const void* x, int64_t rows, int64_t ld, int64_t plane, int32_t col0, int32_t ncols,
                    uint32_t* gmax_bits, int device, void* stream);
/* grad[i] (+)= (float)(acc[i] * 2^-e) */
AVR_API int avr_grid_grad_finalize(const int64_t* acc, int64_t n, const uint32_t* gmax_bits, int32_t log2_headroom,
                           float* grad, int accumulate, int device, void* stream);

/* ---- dense layers (tcnn FullyFusedMLP / CutlassMLP, model.py:117,146,176) ---------------------
 * C[i,j] (+)= sum_k A(i,k) * B(j,k),  i<M, j<N, k<K, fp32 FMA accumulation.
 * `workspace` (>= avr_gemm_workspace_bytes) enables deterministic split-K for long reductions;
 * pass NULL/0 to force a single pass.  aux (MASK) has C's shape with leading dimension ldaux. */
AVR_API int64_t avr_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K);
AVR_API int avr_gemm(int layout_a, int layout_b, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
             const float* B, int64_t ldb, float* C, int64_t ldc, int flags, const float* aux, int64_t ldaux,
             void* workspace, int64_t workspace_bytes, int device, void* stream);

/* ---- dense layers on the tensor cores (tcgen05.mma / TMEM / TMA), fp32-grade accuracy ----------------
 * Operands are error-compensated SETS of bf16 planes, x = hi + mid (+ lo) with hi = bf16(x),
 * mid = bf16(x - hi), lo = bf16(x - hi - mid): a bf16 buffer [nplanes][rows][ld] addressed by
 * (base, ld, plane_stride), all in elements, with 16-byte aligned base / row pitch / plane pitch.
 *   2 planes x 2 planes: hi*hi + hi*mid + mid*hi            (~2^-17 per product; gradients: errors enter linearly)
 *   3 planes x 3 planes: + hi*lo + lo*hi + mid*mid          (~2^-24: fp32-grade pre-activations, needed in the
 *                        forward pass so that ReLU decisions agree with an fp32 evaluation -- DESIGN.md 5)
 * all accumulated in one fp32 TMEM accumulator. */
/* Plane-set kinds (every `*_nplanes` / `*_kind` argument of this header):
 *   AVR_PLANES_BF16x2 / AVR_PLANES_BF16x3   x = hi + mid (+ lo), bf16 planes: 16 / 24 mantissa bits, fp32 range
 *   AVR_PLANES_F16x2                        x = hi + lo' * 2^-11 with hi = fp16(x), lo' = fp16((x - hi) * 2^11):
 *                                           24 mantissa bits in TWO planes for |x| in fp16's normal range
 *                                           (6.1e-5 .. 65504; absolute error <= 1.5e-11 below it, inf above --
 *                                           the range tiny-cuda-nn's own fp16 activations live in).
 * A product of two F16x2 operands needs three tcgen05 MMAs for fp32 grade (hi*hi into one accumulator, hi*lo' +
 * lo'*hi into a second one that the epilogue scales by 2^-11) where two BF16x3 operands need six. */
enum { AVR_PLANES_BF16x2 = 2, AVR_PLANES_BF16x3 = 3, AVR_PLANES_F16x2 = 18 };
enum {
    AVR_UMMA_RELU = 1,        /* out = max(out, 0)                                                   */
    AVR_UMMA_ACCUM = 2,       /* out += previous contents of the output                              */
    AVR_UMMA_MASK = 4,        /* product *= bit (j % 32) of mask_bits[i][j / 32]  (ReLU backward; before ACCUM) */
    AVR_UMMA_OUT_F32 = 8,     /* write fp32 c_f32 instead of a plane pair                            */
    AVR_UMMA_DUAL_RELU = 16,  /* additionally write max(out,0) as a second plane pair (c2)           */
    AVR_UMMA_BITS = 32,       /* additionally write the bitmask (out > 0) to bits_out (1 bit / element) */
    AVR_UMMA_BIAS = 64,       /* out[row,:] += bias_ray[ray(row),:] + bias_rcv[receiver(row),:], row = (b*R + r)*S + s */
    AVR_UMMA_DUAL_COPY = 2048 /* additionally write the same values (after RELU) as a second plane set (c2) of kind c2_nplanes:
                                 an fp16 pair feeds the next forward layer, a bf16 pair the weight-gradient GEMM */
};
/* fp32 [rows, cols] (ld) <-> plane pair; transpose != 0 writes planes[c, r] = x[r, c]; relu != 0 clamps */
AVR_API int avr_planes_split(const float* x, int64_t rows, int64_t cols, int64_t ld, void* planes, int64_t ldp,
                             int64_t plane_stride, int nplanes, int transpose, int relu, int device, void* stream);
/* Several avr_planes_split calls in ONE launch (all weight matrices of a pass). */
#define AVR_SPLIT_BATCH_MAX 32
typedef struct avr_split_desc {
    const float* x; int64_t rows, cols, ld;      /* fp32 source [rows, cols]                              */
    void* planes; int64_t ldp, plane_stride;     /* destination plane set; out[c, r] when transpose != 0 */
    int32_t kind, transpose;
} avr_split_desc;
AVR_API int avr_planes_split_batch(const avr_split_desc* descs, int32_t n, int device, void* stream);
AVR_API int avr_planes_merge(const void* planes, int64_t rows, int64_t cols, int64_t ldp, int64_t plane_stride,
                             int nplanes, float* out, int64_t ld, int device, void* stream);
/* C[M,N] = A[M,K] . B[N,K]^T   (A, B plane pairs with the reduction index contiguous; N % 8 == 0).
 * Forward layers (B = W[out,in]) and backward-data (B = transposed weight planes W^T[in,out]).
 * Near-zero guard (forward layers; near_list != NULL): the tensor core truncates its accumulator once per MMA
 * (~6e-9*K of the row scale, biased), enough to flip the sign of pre-activations that an fp32 evaluation puts within
 * that distance of zero -- and a flipped ReLU / |leaky_relu| decision changes that unit's gradient by 100 %.  Elements
 * with |out| < near_tau * (mean |out| of their 32-column row chunk) are appended to near_list (near_cap entries of
 * {row, col}; *near_count must be 0 on entry and holds the number listed afterwards) and a second kernel of the same
 * call re-evaluates them with fp32 FMAs over the exact operand values and patches c / c2 / c_f32 / bits_out.  The
 * listed set depends only on the values, so results stay bit-identical run to run; a list overflow poisons the
 * output with NaN. */
AVR_API int avr_umma_gemm_nt(int64_t M, int64_t N, int64_t K, const void* a_planes, int64_t lda, int64_t a_plane,
                             int a_nplanes, const void* b_planes, int64_t ldb, int64_t b_plane, int b_nplanes, int flags,
                             void* c_planes, int64_t ldc, int64_t c_plane, int c_nplanes, void* c2_planes, int64_t ldc2,
                             int64_t c2_plane, int c2_nplanes, const uint32_t* mask_bits, int64_t ldmask, uint32_t* bits_out,
                             int64_t ldbits, const float* bias_ray, int64_t ld_bias_ray, const float* bias_rcv,
                             int64_t ld_bias_rcv, int32_t geo_R, int32_t geo_S, float* c_f32, int64_t ldc32,
                             uint32_t* near_list, int64_t near_cap, uint32_t* near_count, float near_tau,
                             void* splitk_workspace, int64_t splitk_workspace_bytes, int device, void* stream);
/* Split-K mode (plain AVR_UMMA_OUT_F32 outputs, splitk_workspace != NULL): the reduction runs as
 * avr_umma_gemm_nt_splitk_slices(K) interleaved 256-deep slices into separate accumulators, summed in fp32 in a fixed
 * order -- for the long reductions of the DFT and its adjoint (K = T, 2F), where the per-MMA truncation of one long
 * accumulator chain is a visible bias (~6e-9*K).  Workspace: slices * M * roundup(N, 8) * 4 bytes. */
AVR_API int64_t avr_umma_gemm_nt_splitk_slices(int64_t K);
/* ---- fused chain of 128-wide dense layers ------------------------------------------------------------------
 * model.py:206-216 -- the sigma encoder and sigma decoder, which the reference evaluates as tiny-cuda-nn FullyFusedMLP
 * kernels (model.py:117,146) -- and their backward-data pass: y_0 = x0, y_{l+1} = act_l(y_l . W_l^T), all layers of a
 * 128-row tile of sample points in ONE launch.  The activation tile stays in shared memory as bf16 planes -- it is the
 * next layer's tcgen05 operand and the source of the bulk stores below -- accumulators live in tensor memory (two sets,
 * so the next layer's MMAs overlap this layer's epilogue), weights stream from L2.  Same arithmetic as avr_umma_gemm_nt
 * (six products / two accumulators when both operands are BF16x3, else three products on the (hi, mid) planes): results
 * are bit-identical to running the layers one by one.
 *   w          BF16x2 / BF16x3 planes (w_kind) of W[n_out, k_in] (K-major), leading dimension ldw, plane stride w_plane
 *   relu       forward: rectify the output before the next layer (and before `save`)
 *   mask       backward: bitmask [M][ldmask] multiplied into the output (the ReLU decisions of the forward pass)
 *   accumulate backward: add the planes already in `save` to the output before it is stored back (two consumers of one
 *              activation: d_feat = signal path + density path)
 *   save       optional: the (rectified / masked) output as a BF16x2 / BF16x3 plane set [*][M][ld_save]  (what the
 *              weight-gradient GEMMs read); save_raw: optional, the un-rectified output likewise (sigma_feat feeds the
 *              signal network raw, model.py:219, and the decoder rectified, model.py:209)
 *   bits       optional ReLU bitmask out: bit (col % 32) of word [row][col / 32] = (raw output > 0)
 *   out_f32    last layer only: fp32 output [M][ld_f32] instead of planes (the 16-wide density head); no activation
 *   bias       optional, forward: fp32 rows [*][ld_bias] added to the output before the activation, row (point /
 *              bias_group_rows) for every point -- the per-receiver channel-embedding row of model.py:44-47
 *              (`h + layer_embeddings[l][ch_id]`), bias_group_rows = R * S points per receiver
 * Hidden layers are 128 wide; k_in of layer 0 (= k0, the width of x0) is a multiple of 16 up to 128; the last layer may
 * be any multiple of 16 up to 128 (up to 64 for fp32).  x0: BF16x2 / BF16x3 planes [*][M][ldx]. */
#define AVR_CHAIN_MAX_LAYERS 8
typedef struct avr_chain_layer {
    const void* w; int64_t ldw, w_plane; int32_t w_kind; int32_t n_out, k_in;
    int32_t relu;
    void* save; int64_t ld_save, save_plane; int32_t save_kind;
    void* save_raw; int64_t ld_raw, raw_plane; int32_t raw_kind;
    uint32_t* bits; int64_t ldbits;
    const uint32_t* mask; int64_t ldmask;
    int32_t accumulate;
    float* out_f32; int64_t ld_f32;
    const float* bias; int64_t ld_bias, bias_group_rows;
} avr_chain_layer;
AVR_API int avr_mlp_chain(int64_t M, const void* x0, int64_t ldx, int64_t x_plane, int32_t x_kind, int32_t k0,
                          const avr_chain_layer* layers, int32_t n_layers, int device, void* stream);

/* C[M,N] (+)= sum_k A[k,M] * B[k,N]   (A = dY[points,out], B = X[points,in] plane sets; weight gradients).
 * fp32 output; deterministic split-K over the points through `workspace`.  A (the gradient) is a bf16 plane set, B
 * (the activation) a bf16 set or an fp16 pair:
 *   BF16x2 . BF16x2 (or mixed counts): three products on the (hi, mid) planes
 *   BF16x3 . BF16x3: six products on 24-bit operands -- for the ill-conditioned sums of the density path, whose
 *                    terms cancel to ~1/50 of their magnitude
 *   BF16x* . F16x2: the activation tile is converted in shared memory to the bf16 (hi, mid) pair avr_planes_split would
 *                    write for the same values (the kernel's epilogue warps do it while the MMAs of earlier tiles run),
 *                    then three products -- bit-identical to passing that bf16 pair
 * The gradient operand itself is never fp16: gradients need bf16's exponent range, and tcgen05.mma faults on a bf16 x f16
 * mix. */
AVR_API int64_t avr_umma_gemm_tn_workspace_bytes(int64_t M, int64_t N, int64_t K);
AVR_API int avr_umma_gemm_tn(int64_t M, int64_t N, int64_t K, const void* a_planes, int64_t lda, int64_t a_plane,
                             int a_kind, const void* b_planes, int64_t ldb, int64_t b_plane, int b_kind, float* c,
                             int64_t ldc, int accumulate, void* workspace, int64_t workspace_bytes, int device,
                             void* stream);

/* ---- output layer fused with the ray reduction ("collapse"; model.py:231 + renderer.py:86-90,115-118) -----
 * y[b,s,t] = sum_r w[b,r,s] [t >= delay[b,r,s]] (H[b,r,s,:] . W_out[t,:]) evaluated as a prefix sum over the
 * rays of each (b,s) sorted by delay, so neither signal[bs,R,S,T] nor its gradient is ever formed.
 *   H: plane pair [bs*R*S, width] (post-ReLU last hidden activation);  W_out: fp32 [T, width] (ld = ldw).
 *   avr_delay_sort: stable counting sort by delay -> order / sdelay / sw, each [bs,S,R]. */
AVR_API int avr_delay_sort(const avr_render_geom* geom, const int32_t* delay, const float* w, int32_t* order,
                           int32_t* sdelay, float* sw, int device, void* stream);
/* tspan: static bound on (max delay - min delay) inside one (b,s) (<= 2*far*fs/speed); a violation poisons the
 * results with NaN instead of corrupting them.  The prefix workspace written by avr_collapse_fwd holds the
 * G snapshots and is handed back to avr_collapse_bwd. */
AVR_API int64_t avr_collapse_prefix_bytes(const avr_render_geom* geom, int32_t width, int32_t tspan);
AVR_API int64_t avr_collapse_suffix_bytes(const avr_render_geom* geom, int32_t width, int32_t tspan);
AVR_API int avr_collapse_fwd(const avr_render_geom* geom, const void* act_planes, int64_t ld_act, int64_t act_plane,
                             int32_t act_kind, int32_t width, const int32_t* order, const int32_t* sdelay, const float* sw,
                             const float* w_out, int64_t ldw, int32_t tspan, void* prefix_ws, int64_t prefix_bytes, float* y,
                             int device, void* stream);
/* d_act = (H > 0) * w * g[delay] as a plane pair (gradient w.r.t. the pre-activation), d_w[bs,R,S] = H . g[delay],
 * d_W_out[T, width] (+)= sum_{b,s} d_y[b,s,t] G[b,s,t,:] */
AVR_API int avr_collapse_bwd(const avr_render_geom* geom, const void* act_planes, int64_t ld_act, int64_t act_plane,
                             int32_t act_kind, int32_t width, const int32_t* order, const int32_t* sdelay, const float* sw,
                             const float* w_out, int64_t ldw, const float* d_y, int32_t tspan, const void* prefix_ws,
                             void* suffix_ws, int64_t suffix_bytes, void* d_act_planes, int64_t ld_d, int64_t d_plane,
                             float* d_w, float* d_wout, int64_t ld_dw, int accumulate, int device, void* stream);

/* ---- broadcast inputs of the signal network (renderer.py:59-60; model.py:219-221) ------------
 * dst[n, col0:col0+w] = src[row(n), 0:w] with row(n) = r (per_receiver=0) or b (per_receiver=1). */
AVR_API int avr_rows_broadcast(const avr_render_geom* geom, const float* src, int32_t w, int per_receiver,
                       void* dst, int64_t ld_dst, int64_t dst_plane, int32_t dst_nplanes, int32_t col0, int device,
                       void* stream);
/* two adjacent column blocks in one launch (plane-set destinations; whole 32-byte sectors where one 40-column block
 * alone leaves half-written ones): dst[n, col0:col0+w] = src[row(n)], dst[n, col0+w:col0+w+w2] = src2[row2(n)] */
AVR_API int avr_rows_broadcast2(const avr_render_geom* geom, const float* src, int32_t w, int per_receiver,
                        const float* src2, int32_t w2, int per_receiver2, void* dst, int64_t ld_dst, int64_t dst_plane,
                        int32_t dst_nplanes, int32_t col0, int device, void* stream);
/* transpose of the above: d_src[row, :] = sum over the points mapped to `row` (fixed order). */
/* partial[(b*R + r), 0:w] = sum over the S sample rows of (b, r) of x (fp32, or plane set: first two planes) */
AVR_API int avr_rows_block_sum(const avr_render_geom* geom, const void* x, int64_t ldx, int64_t x_plane, int32_t w,
                               float* partial, int device, void* stream);
AVR_API int64_t avr_rows_reduce_workspace_bytes(const avr_render_geom* geom, int32_t w, int per_receiver);
AVR_API int avr_rows_reduce(const avr_render_geom* geom, const void* d_dst, int64_t ld_dst, int64_t d_plane, int32_t col0,
                    int32_t w, int per_receiver, float* d_src, float* workspace, int64_t workspace_bytes,
                    int device, void* stream);

/* ---- density -> alpha -> transmittance -> ray weights (renderer.py:167-190; model.py:233) ----
 * attn = |leaky_relu(raw, slope)|; alpha = 1-exp(-attn*delta[s]); w = alpha*prod_{j<s}(1-alpha_j+1e-6)
 * raw is read as raw[n*ld_raw]; attn / w are [bs,R,S] dense (attn may be NULL). One warp per ray.
 * slope < 0: raw already is the density (generic networks_fn path), no activation is applied. */
AVR_API int avr_ray_weights_fwd(const avr_render_geom* geom, const float* raw, int64_t ld_raw, const float* delta,
                        float slope, float* attn, float* w, int device, void* stream);
/* d_raw[n*ld_draw] = dL/draw given d_w[bs,R,S] (recomputes alpha / transmittance from raw) */
AVR_API int avr_ray_weights_bwd(const avr_render_geom* geom, const float* raw, int64_t ld_raw, const float* delta,
                        float slope, const float* d_w, float* d_raw, int64_t ld_draw, int device, void* stream);

/* ---- compositing (renderer.py:79-118), ray reduction done in the time domain ----------------
 * y[b,s,t] = sum_r w[b,r,s] * (t >= delay[b,r,s]) * sig[b,r,s,t]          (masks :82-90, sums :118,192)
 * sig[bs,R,S,T] is streamed exactly once; `workspace` holds per-ray-chunk partial sums. */
AVR_API int64_t avr_composite_workspace_bytes(const avr_render_geom* geom);
AVR_API int avr_composite_fwd(const avr_render_geom* geom, const float* sig, const float* w, const int32_t* delay,
                      float* y, void* workspace, int64_t workspace_bytes, int device, void* stream);
/* d_sig[b,r,s,t] = w*(t>=delay)*d_y[b,s,t];  d_w[b,r,s] = sum_t (t>=delay)*sig*d_y[b,s,t]
 * (d_sig may be NULL when only d_w is wanted, sig may be NULL when only d_sig is wanted) */
AVR_API int avr_composite_bwd(const avr_render_geom* geom, const float* sig, const float* w, const int32_t* delay,
                      const float* d_y, float* d_sig, float* d_w, int device, void* stream);

/* ---- spectrum: tail mask * path loss, real DFT, per-sample phase, sum over samples ----------
 * (renderer.py:82-83,96-109,118-121).  Small tables come from the host glue:
 *   gain[S,T]      = (t < T-1-shift[s]) ? pl[shift[s]+t] : 0
 *   phase[S,F,2]   = (cos ang[s,f], -sin ang[s,f])              (exp(-j*ang), renderer.py:108)
 *   dft[T,ldd]     = columns (cos(2 pi f t/T), -sin(2 pi f t/T)) interleaved, ldd >= 2F, ldd%4==0
 * fwd:  z = y*gain;  X = z @ dft;  out[b,f,:] = sum_s X[b,s,f] * phase[s,f]        -> out[bs,F,2]
 * bwd:  d_y = gain * ( (conj-phase-weighted d_out) @ dft^T )
 * xbuf: scratch [bs*S, ldd] floats (z is formed in place in zbuf [bs*S,T]). */
AVR_API int avr_spectrum_fwd(const avr_render_geom* geom, const float* y, const float* gain, const float* phase,
                     const float* dft, int64_t ldd, float* zbuf, float* xbuf, float* out,
                     int device, void* stream);
AVR_API int avr_spectrum_bwd(const avr_render_geom* geom, const float* d_out, const float* gain, const float* phase,
                     const float* dft, int64_t ldd, float* xbuf, float* d_y, int device, void* stream);

/* The three elementwise pieces of the spectrum stage, for running the DFT on the tensor cores
 * (avr_umma_gemm_nt against a plane set of the DFT matrix): z = y*gain, the phase-weighted sum over samples,
 * and its adjoint.  z / dx: fp32 (plane == 0) or bf16 plane set. */
AVR_API int avr_spectrum_gain(const avr_render_geom* geom, const float* y, const float* gain, void* z, int64_t ldz,
                              int64_t z_plane, int32_t z_nplanes, int device, void* stream);
AVR_API int avr_spectrum_phase_sum(const avr_render_geom* geom, const float* x, int64_t ldx, const float* phase, float* out,
                                   int device, void* stream);
AVR_API int avr_spectrum_phase_bwd(const avr_render_geom* geom, const float* d_out, const float* phase, void* dx, int64_t ldx,
                                   int64_t dx_plane, int32_t dx_nplanes, int device, void* stream);

/* ---- training loss on rendered spectra (utils/criterion.py:69-100) ---------------------------
 * Spectra are interleaved (re, im) fp32 rows of F = T/2+1 bins; time signals are fp32 rows of T samples.  Every term
 * writes fixed-order partial sums (the caller adds them and applies weight / count) and, when the gradient pointer is
 * non-NULL, d(scale * sum)/d(pred) in the same pass.  `dft` is the [T, ldd] (cos, -sin) table of avr_spectrum_*. */
/* x[r,:] = irfft(spec[r,:])  (criterion.py:71-72) and its adjoint out[r,f] = d<x, d_x>/d spec[r,f] */
AVR_API int avr_crit_irfft(const float* spec, int32_t n_rows, int32_t T, const float* dft, int64_t ldd, float* x, int device,
                   void* stream);
AVR_API int avr_crit_irfft_adjoint(const float* d_x, int32_t n_rows, int32_t T, const float* dft, int64_t ldd, float* out,
                           int device, void* stream);
/* criterion.py:86-93: partial[bs,3] = (sum |d re| + |d im|, sum ||X|-|Y||, sum |d cos| + |d sin| of the angles);
 * grad[3,bs,F,2] (may be NULL) = scale_k * d(sum_k)/d pred */
AVR_API int avr_crit_freq_terms(const float* pred, const float* ori, int32_t bs, int32_t F, float scale_spec, float scale_amp,
                        float scale_angle, float* partial, float* grad, int device, void* stream);
/* criterion.py:95: partial[bs] = sum_t |ori - pred|; d_x[bs,T] (may be NULL) = scale * d(sum)/d x_pred */
AVR_API int avr_crit_time_l1(const float* x_pred, const float* x_ori, int32_t bs, int32_t T, float scale, float* partial,
                     float* d_x, int device, void* stream);
/* criterion.py:74-84,97: energy-decay curves of the rectangular-window STFT (n_fft, hop; center, reflect padding);
 * partial[bs] = sum_m |E_ori - E_pred|; d_x[bs,T] (may be NULL) = scale * d(sum)/d x_pred */
AVR_API int64_t avr_crit_energy_workspace_bytes(int32_t bs, int32_t T, int32_t hop);
AVR_API int avr_crit_energy(const float* x_pred, const float* x_ori, int32_t bs, int32_t T, int32_t n_fft, int32_t hop,
                    float scale, float* partial, float* d_x, float* workspace, int64_t workspace_bytes, int device,
                    void* stream);
/* One resolution of the multi-resolution STFT term (criterion.py:33,99; auraloss STFTLoss): window[win] centred in
 * n_fft (a power of two), M = 1 + T/hop frames, K = n_fft/2+1 bins, magnitudes sqrt(max(re^2+im^2, eps)).
 * fwd: S_pred[bs,M,K,2], mag_ori[bs,M,K], partial[bs*M,4] = sums of (P-O)^2, P^2, |log O - log P|, |O - P|.
 * bwd: with sums[4] (device) = the column sums of `partial`, adds to d_x[bs,T] the gradient of
 *      scale * (sqrt(A)/sqrt(B) + C/n + w_lin * D/n), n = bs*M*K; frames[bs,M,win] is scratch. */
AVR_API int avr_crit_stft_fwd(const float* x_pred, const float* x_ori, int32_t bs, int32_t T, int32_t n_fft, int32_t hop,
                      int32_t win, const float* window, float eps, float* S_pred, float* mag_ori, float* partial,
                      int device, void* stream);
AVR_API int avr_crit_stft_bwd(const float* S_pred, const float* mag_ori, const float* sums, int32_t bs, int32_t T, int32_t n_fft,
                      int32_t hop, int32_t win, const float* window, float eps, float scale, float w_lin, float* frames,
                      float* d_x, int device, void* stream);

/* ---- fused optimiser step over flat fp32 arenas (avr_runner.py:192-200: clip_grad_norm_ -> NaN/Inf scrub -> Adam) ----
 * norm_out[0] = total gradient L2 norm (before clipping), norm_out[1] = clip coefficient.  max_norm <= 0: no
 * clipping.  step is the 1-based Adam step.  write_back_grad != 0 stores the clipped / scrubbed gradient. */
AVR_API int64_t avr_adam_workspace_bytes(void);
AVR_API int avr_fused_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                                float beta1, float beta2, float eps, float weight_decay, float max_norm, int64_t step,
                                int write_back_grad, float* norm_out, void* workspace, int64_t workspace_bytes, int device,
                                void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AVR_B200_H */
