// fp32 dense layers for the density / signal MLPs (bias-free, ReLU) -- replaces the tiny-cuda-nn
// FullyFusedMLP / CutlassMLP launched from model.py:117,146,176 and their backward passes.
//
// C[i,j] (+)= sum_k A(i,k) * B(j,k) in exact fp32 FMA arithmetic: the 1e-4 rel-L2 parity bar against
// the fp32 oracle rules out single-pass bf16/tf32 tensor-core products (SURVEY section 7 "hard parts").
// 128x128x16 CTA tile, 256 threads, 8x8 register tile, double-buffered shared memory with register
// prefetch of the next k-slab.  Long reductions (weight gradients: K = number of sample points) use a
// deterministic split-K: every slice writes its partial tile, a second kernel adds them in slice order.
#include "common.cuh"

namespace avr {

constexpr int BM = 128, BN = 128, BK = 16, GEMM_THREADS = 256;
constexpr int LDS = BM + 4;   // smem row stride (floats), keeps float4 alignment

template <int LAYOUT>
__device__ __forceinline__ void load_tile(const float* __restrict__ P, int64_t ld, int64_t dim, int64_t i0, int64_t k0,
                                          int64_t k_end, int tid, bool relu_in, float4 (&st)[2]) {
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int f = tid + t * GEMM_THREADS;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (LAYOUT == AVR_K_CONTIG) {
            const int row = f >> 2, kq = (f & 3) * 4;
            if (i0 + row < dim && k0 + kq < k_end) v = *reinterpret_cast<const float4*>(P + (i0 + row) * ld + k0 + kq);
        } else {
            const int k = f >> 5, iq = (f & 31) * 4;
            if (k0 + k < k_end && i0 + iq < dim) v = *reinterpret_cast<const float4*>(P + (k0 + k) * ld + i0 + iq);
        }
        if (relu_in) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        st[t] = v;
    }
}

template <int LAYOUT>
__device__ __forceinline__ void store_tile(float (*S)[LDS], int tid, const float4 (&st)[2]) {
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int f = tid + t * GEMM_THREADS;
        if (LAYOUT == AVR_K_CONTIG) {
            const int row = f >> 2, kq = (f & 3) * 4;
            S[kq + 0][row] = st[t].x; S[kq + 1][row] = st[t].y; S[kq + 2][row] = st[t].z; S[kq + 3][row] = st[t].w;
        } else {
            const int k = f >> 5, iq = (f & 31) * 4;
            *reinterpret_cast<float4*>(&S[k][iq]) = st[t];
        }
    }
}

__device__ __forceinline__ float4 epilogue4(float4 v, int flags, const float* aux_p, const float* c_p) {
    if (flags & AVR_GEMM_MASK) {
        const float4 m = *reinterpret_cast<const float4*>(aux_p);
        v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f; v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
    }
    if (flags & AVR_GEMM_ACCUM) {
        const float4 c = *reinterpret_cast<const float4*>(c_p);
        v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w;
    }
    if (flags & AVR_GEMM_RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    return v;
}

// grid: (tiles_m, tiles_n, splits).  With splits > 1 the raw partial tile goes to `partial`
// ([splits, M, N] dense) and splitk_reduce_kernel applies the epilogue.
template <int LA, int LB>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
sgemm_kernel(int64_t M, int64_t N, int64_t K, const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
             int64_t ldb, float* __restrict__ C, int64_t ldc, int flags, const float* __restrict__ aux, int64_t ldaux,
             int64_t k_chunk, float* __restrict__ partial) {
    __shared__ __align__(16) float As[2][BK][LDS];
    __shared__ __align__(16) float Bs[2][BK][LDS];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t i0 = (int64_t)blockIdx.x * BM, j0 = (int64_t)blockIdx.y * BN;
    const int64_t k_beg = (int64_t)blockIdx.z * k_chunk;
    const int64_t k_end = (k_beg + k_chunk < K) ? k_beg + k_chunk : K;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float4 sa[2], sb[2];
    const bool relu_a = (flags & AVR_GEMM_RELU_A) != 0, relu_b = (flags & AVR_GEMM_RELU_B) != 0;
    const int n_iter = (int)((k_end - k_beg + BK - 1) / BK);
    if (n_iter > 0) {
        load_tile<LA>(A, lda, M, i0, k_beg, k_end, tid, relu_a, sa);
        load_tile<LB>(B, ldb, N, j0, k_beg, k_end, tid, relu_b, sb);
        store_tile<LA>(As[0], tid, sa);
        store_tile<LB>(Bs[0], tid, sb);
    }
    __syncthreads();
    for (int it = 0; it < n_iter; ++it) {
        const int buf = it & 1;
        if (it + 1 < n_iter) {
            const int64_t k0 = k_beg + (int64_t)(it + 1) * BK;
            load_tile<LA>(A, lda, M, i0, k0, k_end, tid, relu_a, sa);
            load_tile<LB>(B, ldb, N, j0, k0, k_end, tid, relu_b, sb);
        }
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (it + 1 < n_iter) {
            store_tile<LA>(As[buf ^ 1], tid, sa);
            store_tile<LB>(Bs[buf ^ 1], tid, sb);
        }
        __syncthreads();
    }

    const bool split = gridDim.z > 1;
    float* out = split ? partial + (int64_t)blockIdx.z * M * N : C;
    const int64_t ldo = split ? N : ldc;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t row = i0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (row >= M) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t col = j0 + h * 64 + tx * 4;
            if (col >= N) continue;
            float4 v = make_float4(acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
            if (!split) v = epilogue4(v, flags, aux ? aux + row * ldaux + col : nullptr, C + row * ldc + col);
            *reinterpret_cast<float4*>(out + row * ldo + col) = v;
        }
    }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, int64_t M, int64_t N,
                                     float* __restrict__ C, int64_t ldc, int flags, const float* __restrict__ aux,
                                     int64_t ldaux) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // float4 index
    const int64_t nq = N / 4;
    if (q >= M * nq) return;
    const int64_t row = q / nq, col = (q - row * nq) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int z = 0; z < splits; ++z) {                                   // fixed order: deterministic
        const float4 p = *reinterpret_cast<const float4*>(partial + ((int64_t)z * M + row) * N + col);
        v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    }
    v = epilogue4(v, flags, aux ? aux + row * ldaux + col : nullptr, C + row * ldc + col);
    *reinterpret_cast<float4*>(C + row * ldc + col) = v;
}

static int plan_splits(int64_t M, int64_t N, int64_t K, int64_t* k_chunk) {
    const int64_t tiles = ceil_div(M, BM) * ceil_div(N, BN);
    int64_t splits = 1;
    if (tiles < 148 && K >= 1024) {
        splits = ceil_div(2 * 148, tiles);
        const int64_t max_by_k = K / 512;
        if (splits > max_by_k) splits = max_by_k;
        if (splits < 1) splits = 1;
    }
    int64_t chunk = ceil_div(ceil_div(K, splits), BK) * BK;
    splits = ceil_div(K, chunk);
    *k_chunk = chunk;
    return (int)splits;
}

template <int LA, int LB>
static void launch_sgemm(dim3 grid, cudaStream_t st, int64_t M, int64_t N, int64_t K, const float* A,
                         int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int flags, const float* aux,
                         int64_t ldaux, int64_t k_chunk, float* partial) {
    sgemm_kernel<LA, LB><<<grid, GEMM_THREADS, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, flags, aux, ldaux, k_chunk, partial);
}

int gemm_impl(int la, int lb, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb,
              float* C, int64_t ldc, int flags, const float* aux, int64_t ldaux, void* workspace, int64_t workspace_bytes,
              cudaStream_t st) {
    AVR_REQUIRE(A && B && C, "null operand");
    AVR_REQUIRE(M >= 0 && N >= 0 && K >= 0, "negative dimension");
    if (M == 0 || N == 0) return AVR_OK;
    AVR_REQUIRE(N % 4 == 0 && ldc % 4 == 0 && aligned16(C), "C: N, ldc must be multiples of 4 and C 16-byte aligned");
    AVR_REQUIRE(lda % 4 == 0 && ldb % 4 == 0 && aligned16(A) && aligned16(B), "A/B: ld multiple of 4, 16-byte aligned");
    AVR_REQUIRE(la == AVR_K_CONTIG ? (K % 4 == 0) : (M % 4 == 0), "A: contiguous extent must be a multiple of 4");
    AVR_REQUIRE(lb == AVR_K_CONTIG ? (K % 4 == 0) : (N % 4 == 0), "B: contiguous extent must be a multiple of 4");
    if (flags & AVR_GEMM_MASK) AVR_REQUIRE(aux && ldaux % 4 == 0 && aligned16(aux), "MASK needs an aligned aux");
    if (!(la == AVR_K_CONTIG && lb == AVR_K_CONTIG) && !(la == AVR_K_CONTIG && lb == AVR_I_CONTIG) &&
        !(la == AVR_I_CONTIG && lb == AVR_I_CONTIG))
        return fail(AVR_ERR_UNSUPPORTED, "avr_gemm: layout pair (%d,%d) not built", la, lb);
    AVR_REQUIRE(ceil_div(N, BN) <= 65535, "N too large for grid.y");

    int64_t k_chunk = K;
    int splits = 1;
    if (workspace && workspace_bytes > 0 && K > 0) {
        splits = plan_splits(M, N, K, &k_chunk);
        if (splits > 1 && (int64_t)splits * M * N * (int64_t)sizeof(float) > workspace_bytes) { splits = 1; k_chunk = K; }
        if (splits > 1) AVR_REQUIRE(aligned16(workspace), "workspace must be 16-byte aligned");
    }
    if (k_chunk <= 0) k_chunk = BK;
    const dim3 grid((unsigned)ceil_div(M, BM), (unsigned)ceil_div(N, BN), (unsigned)splits);
    float* partial = splits > 1 ? (float*)workspace : nullptr;
    if (la == AVR_K_CONTIG && lb == AVR_K_CONTIG)
        launch_sgemm<AVR_K_CONTIG, AVR_K_CONTIG>(grid, st, M, N, K, A, lda, B, ldb, C, ldc, flags, aux, ldaux, k_chunk, partial);
    else if (la == AVR_K_CONTIG)
        launch_sgemm<AVR_K_CONTIG, AVR_I_CONTIG>(grid, st, M, N, K, A, lda, B, ldb, C, ldc, flags, aux, ldaux, k_chunk, partial);
    else
        launch_sgemm<AVR_I_CONTIG, AVR_I_CONTIG>(grid, st, M, N, K, A, lda, B, ldb, C, ldc, flags, aux, ldaux, k_chunk, partial);
    AVR_LAUNCH_CHECK();
    if (splits > 1) {
        const int64_t nq = M * (N / 4);
        splitk_reduce_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, st>>>(partial, splits, M, N, C, ldc, flags, aux, ldaux);
        AVR_LAUNCH_CHECK();
    }
    return AVR_OK;
}

}  // namespace avr

using namespace avr;

extern "C" int64_t avr_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    int64_t k_chunk;
    const int splits = plan_splits(M, N, K, &k_chunk);
    return splits > 1 ? (int64_t)splits * M * N * (int64_t)sizeof(float) : 0;
}

extern "C" int avr_gemm(int layout_a, int layout_b, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                        const float* B, int64_t ldb, float* C, int64_t ldc, int flags, const float* aux, int64_t ldaux,
                        void* workspace, int64_t workspace_bytes, int device, void* stream) {
    AVR_ENTER(device);
    return gemm_impl(layout_a, layout_b, M, N, K, A, lda, B, ldb, C, ldc, flags, aux, ldaux, workspace, workspace_bytes,
                     (cudaStream_t)stream);
}
