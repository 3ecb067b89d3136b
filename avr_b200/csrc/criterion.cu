// The training loss of the reference (utils/criterion.py:69-100) on rendered spectra, SURVEY 8f rank 1.
//
// The data is tiny (bs x 801 complex bins, bs x 1600 samples) and the reference spends ~40 launches plus four
// torch.stft / auraloss passes on it; here every term is a couple of small kernels that produce the loss partials
// AND d(term)/d(pred time signal or spectrum) in the same pass, so that autograd's backward is one weighted sum.
//   irfft / its adjoint        : dense real DFT against the renderer's cached [T, 2F] (cos, -sin) table
//   spectral / amplitude / angle L1 : one pass over [bs, F]
//   time L1                    : one pass over [bs, T]
//   energy-decay term          : frame energies of the rectangular n_fft STFT by Parseval (no DFT), EDC on one warp
//   multi-resolution STFT term : direct windowed DFT per frame (only win_length samples are non-zero), magnitudes,
//                                spectral-convergence / log / linear sums; the backward pass re-uses the stored pred
//                                STFT and folds the frame gradients back through the reflect padding
// All reductions have a fixed order (per-CTA partials summed by the caller): results are bit-reproducible.
#include "common.cuh"

namespace avr {

namespace {

constexpr float kLn10 = 2.302585092994046f;

__device__ __forceinline__ float sgnf(float v) { return (v > 0.f) - (v < 0.f); }

__device__ __forceinline__ int reflect_index(int j, int T) {       // torch.stft center=True, pad_mode="reflect"
    if (j < 0) j = -j;
    if (j >= T) j = 2 * (T - 1) - j;
    return j;
}

template <int N>
__device__ __forceinline__ void block_sum(float (&v)[N], float* smem /* >= N * 32 */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
    __syncthreads();
    if (lane == 0)
        for (int i = 0; i < N; ++i) smem[i * 32 + warp] = v[i];
    __syncthreads();
    for (int i = 0; i < N; ++i) {
        float t = 0.f;
        for (int w = 0; w < n_warps; ++w) t += smem[i * 32 + w];           // same order in every thread
        v[i] = t;
    }
}

// x[r, t] = (1/T) sum_f c_f (a_f cos(2 pi f t / T) - b_f sin(2 pi f t / T)),  c_0 = c_{F-1} = 1 else 2; the imaginary
// parts of the DC and Nyquist bins are ignored (torch.fft.irfft).  The table row f is read at column t' = min(t, T - t)
// (cos is even, sin odd about T/2) so that a warp's loads are contiguous.
__global__ void __launch_bounds__(128)
irfft_kernel(const float2* __restrict__ spec, int n_rows, int T, const float* __restrict__ dft, int64_t ldd, float* __restrict__ x) {
    extern __shared__ float2 row[];
    const int F = T / 2 + 1, r = blockIdx.y;
    for (int f = threadIdx.x; f < F; f += blockDim.x) row[f] = spec[(int64_t)r * F + f];
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int tp = t <= T / 2 ? t : T - t;
    const float sgn = t <= T / 2 ? 1.f : -1.f;
    // double accumulator: the STFT log-magnitude term downstream is sensitive to the rounding noise of quiet samples
    double acc = (double)row[0].x + (double)row[F - 1].x * (double)__ldg(dft + (int64_t)(F - 1) * ldd + 2 * tp);
    for (int f = 1; f < F - 1; ++f) {
        const float2 cs = __ldg(reinterpret_cast<const float2*>(dft + (int64_t)f * ldd) + tp);     // (cos, -sin)
        acc += 2.0 * ((double)row[f].x * (double)cs.x + (double)(sgn * row[f].y) * (double)cs.y);
    }
    x[(int64_t)r * T + t] = (float)(acc / (double)T);
}

// out[r, f] = (c_f / T) * sum_t d_x[r, t] * (cos, -sin)(2 pi f t / T), imaginary parts of DC / Nyquist = 0
__global__ void __launch_bounds__(128)
irfft_adjoint_kernel(const float* __restrict__ d_x, int n_rows, int T, const float* __restrict__ dft, int64_t ldd,
                     float2* __restrict__ out) {
    extern __shared__ float xs[];
    const int F = T / 2 + 1, r = blockIdx.y;
    for (int t = threadIdx.x; t < T; t += blockDim.x) xs[t] = d_x[(int64_t)r * T + t];
    __syncthreads();
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    float re = 0.f, im = 0.f;
    for (int t = 0; t < T; ++t) {
        const float2 cs = __ldg(reinterpret_cast<const float2*>(dft + (int64_t)t * ldd) + f);
        re = fmaf(xs[t], cs.x, re);
        im = fmaf(xs[t], cs.y, im);
    }
    const bool edge = f == 0 || f == F - 1;
    const float c = (edge ? 1.f : 2.f) / (float)T;
    out[(int64_t)r * F + f] = make_float2(c * re, edge ? 0.f : c * im);
}

// criterion.py:86-93.  partial[b] = (sum |da| + |db|, sum ||X| - |Y||, sum |dcos| + |dsin|); grad[k][b][f] = scale[k] * d(sum_k)/dX
__global__ void __launch_bounds__(256)
freq_terms_kernel(const float2* __restrict__ pred, const float2* __restrict__ ori, int F, float s_spec, float s_amp, float s_ang,
                  float* __restrict__ partial, float2* __restrict__ grad, int64_t term_stride) {
    __shared__ float red[3 * 32];
    const int b = blockIdx.x;
    float acc[3] = {0.f, 0.f, 0.f};
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        const int64_t i = (int64_t)b * F + f;
        const float2 X = pred[i], Y = ori[i];
        const float da = X.x - Y.x, db = X.y - Y.y;
        acc[0] += fabsf(da) + fabsf(db);
        const float m = hypotf(X.x, X.y), my = hypotf(Y.x, Y.y);
        acc[1] += fabsf(m - my);
        // cos / sin of torch.angle(): formed through atan2 like the reference, so that exactly-real bins behave the same
        // (sin(atan2(0, a < 0)) = sin(fl(pi)) = -8.7e-8, not 0, and the L1 sign sees it); angle(0) = 0 with no gradient
        float cx, sx, cy, sy;
        sincosf(atan2f(X.y, X.x), &sx, &cx);
        sincosf(atan2f(Y.y, Y.x), &sy, &cy);
        acc[2] += fabsf(cx - cy) + fabsf(sx - sy);
        if (grad != nullptr) {
            grad[i] = make_float2(s_spec * sgnf(da), s_spec * sgnf(db));
            const float ga = m > 0.f ? s_amp * sgnf(m - my) / m : 0.f;
            grad[term_stride + i] = make_float2(ga * X.x, ga * X.y);
            float gx = 0.f, gy = 0.f;
            if (m > 0.f) {
                // d angle = (-b da + a db) / |X|^2;  d cos = -sin d angle, d sin = cos d angle
                const float q = s_ang * (-sgnf(cx - cy) * sx + sgnf(sx - sy) * cx) / (m * m);
                gx = -X.y * q;
                gy = X.x * q;
            }
            grad[2 * term_stride + i] = make_float2(gx, gy);
        }
    }
    block_sum(acc, red);
    if (threadIdx.x == 0) { partial[3 * b] = acc[0]; partial[3 * b + 1] = acc[1]; partial[3 * b + 2] = acc[2]; }
}

// criterion.py:95.  partial[b] = sum_t |y - x|;  d_x = -scale * sign(y - x)
__global__ void __launch_bounds__(256)
time_l1_kernel(const float* __restrict__ xp, const float* __restrict__ xo, int T, float scale, float* __restrict__ partial,
               float* __restrict__ d_x) {
    __shared__ float red[32];
    const int b = blockIdx.x;
    float acc[1] = {0.f};
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        const float d = xo[(int64_t)b * T + t] - xp[(int64_t)b * T + t];
        acc[0] += fabsf(d);
        if (d_x != nullptr) d_x[(int64_t)b * T + t] = -scale * sgnf(d);
    }
    block_sum(acc, red);
    if (threadIdx.x == 0) partial[b] = acc[0];
}

// ---- energy-decay term (criterion.py:74-84,97) ------------------------------------------------------------------
// Frame m of the rectangular-window STFT covers padded samples [m*hop, m*hop + n_fft).  Its one-sided energy
//   e_m = sum_{k=0}^{n_fft/2} |S_k|^2 = (n_fft * sum x^2 + S_0^2 + S_{n/2}^2) / 2      (Parseval + Hermitian symmetry)
// with S_0 = sum x, S_{n/2} = sum (-1)^n x.  stats[sig][b][m] = (e, S_0, S_{n/2}).
__global__ void __launch_bounds__(256)
frame_energy_kernel(const float* __restrict__ xp, const float* __restrict__ xo, int bs, int T, int n_fft, int hop, int M,
                    float* __restrict__ stats) {
    __shared__ float red[3 * 32];
    const int m = blockIdx.x, sb = blockIdx.y;                     // sb = sig * bs + b
    const float* x = (sb < bs ? xp : xo) + (int64_t)(sb % bs) * T;
    float acc[3] = {0.f, 0.f, 0.f};
    for (int n = threadIdx.x; n < n_fft; n += blockDim.x) {
        const float v = x[reflect_index(m * hop + n - n_fft / 2, T)];
        acc[0] = fmaf(v, v, acc[0]);
        acc[1] += v;
        acc[2] += (n & 1) ? -v : v;
    }
    block_sum(acc, red);
    if (threadIdx.x == 0) {
        float* o = stats + ((int64_t)sb * M + m) * 3;
        o[0] = 0.5f * ((float)n_fft * acc[0] + acc[1] * acc[1] + acc[2] * acc[2]);
        o[1] = acc[1];
        o[2] = acc[2];
    }
}

// one warp per receiver: E_m = log10(sum_{m' >= m} e_m'^2 + 1e-9) - (same at m = 0) for both signals, the L1 partial and
// de[b][m] = d(scale * sum_m |E^o_m - E^p_m|) / d e^p_m
__global__ void edc_kernel(const float* __restrict__ stats, int bs, int M, float scale, float* __restrict__ partial,
                           float* __restrict__ de, float* __restrict__ scratch /* [bs][3][M] */) {
    const int b = blockIdx.x;
    if (threadIdx.x != 0) return;
    float* Rp = scratch + (int64_t)b * 3 * M;
    float* Ro = Rp + M;
    float* h = Ro + M;
    const float* sp = stats + (int64_t)b * M * 3;
    const float* so = stats + (int64_t)(bs + b) * M * 3;
    float cp = 0.f, co = 0.f;
    for (int m = M - 1; m >= 0; --m) {                             // reversed cumulative sums of e^2
        cp += sp[3 * m] * sp[3 * m];
        co += so[3 * m] * so[3 * m];
        Rp[m] = cp;
        Ro[m] = co;
    }
    const float Ep0 = log10f(Rp[0] + 1e-9f), Eo0 = log10f(Ro[0] + 1e-9f);
    float sum = 0.f, h_total = 0.f;
    for (int m = 0; m < M; ++m) {
        const float d = (log10f(Ro[m] + 1e-9f) - Eo0) - (log10f(Rp[m] + 1e-9f) - Ep0);
        sum += fabsf(d);
        h[m] = -scale * sgnf(d);                                  // dL / dE^p_m
        h_total += h[m];
    }
    partial[b] = sum;
    if (de == nullptr) return;
    float run = 0.f;                                              // dL/dq_m = sum_{j <= m} dL/dR_j
    for (int m = 0; m < M; ++m) {
        float dR = h[m] / ((Rp[m] + 1e-9f) * kLn10);
        if (m == 0) dR -= h_total / ((Rp[0] + 1e-9f) * kLn10);
        run += dR;
        de[(int64_t)b * M + m] = run * 2.f * sp[3 * m];           // q = e^2
    }
}

// gradient w.r.t. one padded sample p of one receiver: frames covering p, d e_m / d x = n_fft x + S_0 + (-1)^n S_{n/2}
__device__ __forceinline__ float energy_dpad(int p, float xv, const float* __restrict__ sp, const float* __restrict__ de,
                                             int n_fft, int hop, int M) {
    float g = 0.f;
    const int first = p - n_fft + 1;                               // smallest m with m*hop + n_fft > p
    const int m_hi = min(M - 1, p / hop), m_lo = first > 0 ? (first + hop - 1) / hop : 0;
    for (int m = m_lo; m <= m_hi; ++m) {
        const int n = p - m * hop;
        g += de[m] * ((float)n_fft * xv + sp[3 * m + 1] + ((n & 1) ? -sp[3 * m + 2] : sp[3 * m + 2]));
    }
    return g;
}

__global__ void __launch_bounds__(256)
energy_bwd_kernel(const float* __restrict__ xp, const float* __restrict__ stats, const float* __restrict__ de, int T, int n_fft,
                  int hop, int M, float* __restrict__ d_x) {
    const int b = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int pad = n_fft / 2;
    const float* sp = stats + (int64_t)b * M * 3;
    const float* deb = de + (int64_t)b * M;
    const float xv = xp[(int64_t)b * T + t];
    float g = energy_dpad(t + pad, xv, sp, deb, n_fft, hop, M);
    if (t >= 1 && t <= pad) g += energy_dpad(pad - t, xv, sp, deb, n_fft, hop, M);                       // left mirror
    if (t <= T - 2 && t >= T - 1 - pad) g += energy_dpad(pad + 2 * (T - 1) - t, xv, sp, deb, n_fft, hop, M);   // right mirror
    d_x[(int64_t)b * T + t] = g;
}

// ---- multi-resolution STFT term (criterion.py:33,99; auraloss STFTLoss) ---------------------------------------------
// One CTA per (frame, receiver): windowed DFT of the pred and the ori frame, magnitudes sqrt(max(re^2+im^2, eps)),
// and the four sums the loss needs: sum (P-O)^2, sum P^2, sum |log O - log P|, sum |O - P|.
__global__ void __launch_bounds__(128)
stft_fwd_kernel(const float* __restrict__ xp, const float* __restrict__ xo, int T, int n_fft, int hop, int win,
                const float* __restrict__ window, int M, float eps, float2* __restrict__ S_pred, float* __restrict__ mag_ori,
                float* __restrict__ partial) {
    extern __shared__ float sm[];
    float2* tw = reinterpret_cast<float2*>(sm);                    // [n_fft] (cos, sin)(2 pi j / n_fft)
    float* fp = sm + 2 * n_fft;                                    // [win] windowed pred frame
    float* fo = fp + win;                                          // [win] windowed ori frame
    float* red = fo + win;                                         // [4 * 32]
    const int m = blockIdx.x, b = blockIdx.y, K = n_fft / 2 + 1, left = (n_fft - win) / 2, pad = n_fft / 2;
    for (int j = threadIdx.x; j < n_fft; j += blockDim.x) {
        float s, c;
        sincospif(2.f * (float)j / (float)n_fft, &s, &c);
        tw[j] = make_float2(c, s);
    }
    for (int j = threadIdx.x; j < win; j += blockDim.x) {
        const int t = reflect_index(m * hop + left + j - pad, T);
        const float w = window[j];
        fp[j] = w * xp[(int64_t)b * T + t];
        fo[j] = w * xo[(int64_t)b * T + t];
    }
    __syncthreads();
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        // double accumulators: the log-magnitude term divides by |S|^2, so bins where the frame nearly cancels need the
        // sum itself to be exact to fp32 rounding of the inputs (the whole term is ~5 M MACs per receiver)
        double dpr = 0.0, dpi = 0.0, dor = 0.0, doi = 0.0;
        int idx = (k * left) & (n_fft - 1);                        // n_fft is a power of two
        for (int j = 0; j < win; ++j) {
            const float2 cs = tw[idx];
            dpr += (double)fp[j] * (double)cs.x; dpi -= (double)fp[j] * (double)cs.y;
            dor += (double)fo[j] * (double)cs.x; doi -= (double)fo[j] * (double)cs.y;
            idx = (idx + k) & (n_fft - 1);
        }
        const float pr = (float)dpr, pi = (float)dpi, orr = (float)dor, oi = (float)doi;
        const float P = sqrtf(fmaxf(pr * pr + pi * pi, eps)), O = sqrtf(fmaxf(orr * orr + oi * oi, eps));
        const int64_t o = ((int64_t)b * M + m) * K + k;
        S_pred[o] = make_float2(pr, pi);
        mag_ori[o] = O;
        acc[0] += (P - O) * (P - O);
        acc[1] += P * P;
        acc[2] += fabsf(logf(O) - logf(P));
        acc[3] += fabsf(O - P);
    }
    block_sum(acc, red);
    if (threadIdx.x == 0) {
        float* o = partial + ((int64_t)b * M + m) * 4;
        o[0] = acc[0]; o[1] = acc[1]; o[2] = acc[2]; o[3] = acc[3];
    }
}

// d(loss_r)/d(windowed pred frame sample), times the window: frames[b][m][j].
// loss_r = scale * ( sqrt(A)/sqrt(B) + C/n + w_lin * D/n ),  sums = (A, B, C, D) on the device.
__global__ void __launch_bounds__(128)
stft_bwd_kernel(const float2* __restrict__ S_pred, const float* __restrict__ mag_ori, const float* __restrict__ sums, int n_fft,
                int win, const float* __restrict__ window, int M, float eps, float scale, float w_lin, float inv_n,
                float* __restrict__ frames) {
    extern __shared__ float sm[];
    float2* tw = reinterpret_cast<float2*>(sm);                    // [n_fft]
    float2* g = tw + n_fft;                                        // [K] dL/d(re, im)
    const int m = blockIdx.x, b = blockIdx.y, K = n_fft / 2 + 1, left = (n_fft - win) / 2;
    for (int j = threadIdx.x; j < n_fft; j += blockDim.x) {
        float s, c;
        sincospif(2.f * (float)j / (float)n_fft, &s, &c);
        tw[j] = make_float2(c, s);
    }
    const float A = sums[0], B = sums[1];
    const float rA = sqrtf(A), rB = sqrtf(B);
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const int64_t o = ((int64_t)b * M + m) * K + k;
        const float2 S = S_pred[o];
        const float p2 = S.x * S.x + S.y * S.y, O = mag_ori[o];
        float2 gk = make_float2(0.f, 0.f);
        if (p2 >= eps) {                                          // below the clamp the magnitude is a constant
            const float P = sqrtf(p2);
            float dP = 0.f;
            if (A > 0.f) dP += (P - O) / (rA * rB);
            dP -= rA * P / (B * rB);
            dP += inv_n * (sgnf(logf(P) - logf(O)) / P + w_lin * sgnf(P - O));
            dP *= scale / P;
            gk = make_float2(dP * S.x, dP * S.y);
        }
        g[k] = gk;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < win; j += blockDim.x) {
        float acc = 0.f;
        const int step = (left + j) & (n_fft - 1);
        int idx = 0;
        for (int k = 0; k < K; ++k) {                              // re = sum x cos, im = -sum x sin
            const float2 cs = tw[idx];
            acc = fmaf(g[k].x, cs.x, acc);
            acc = fmaf(-g[k].y, cs.y, acc);
            idx = (idx + step) & (n_fft - 1);
        }
        frames[((int64_t)b * M + m) * win + j] = window[j] * acc;
    }
}

__device__ __forceinline__ float frames_at(const float* __restrict__ fr, int p, int hop, int left, int win, int M) {
    float g = 0.f;
    const int q = p - left;                                       // frame m holds padded samples [m*hop+left, m*hop+left+win)
    if (q < 0) return 0.f;
    const int first = q - win + 1;                                // smallest m with m*hop + win > q
    const int m_hi = min(M - 1, q / hop), m_lo = first > 0 ? (first + hop - 1) / hop : 0;
    for (int m = m_lo; m <= m_hi; ++m) g += fr[(int64_t)m * win + (q - m * hop)];
    return g;
}

// overlap-add of the frame gradients through the reflect padding, accumulated onto d_x
__global__ void __launch_bounds__(256)
stft_fold_kernel(const float* __restrict__ frames, int T, int n_fft, int hop, int win, int M, float* __restrict__ d_x) {
    const int b = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int pad = n_fft / 2, left = (n_fft - win) / 2;
    const float* fr = frames + (int64_t)b * M * win;
    float g = frames_at(fr, t + pad, hop, left, win, M);
    if (t >= 1 && t <= pad) g += frames_at(fr, pad - t, hop, left, win, M);
    if (t <= T - 2 && t >= T - 1 - pad) g += frames_at(fr, pad + 2 * (T - 1) - t, hop, left, win, M);
    d_x[(int64_t)b * T + t] += g;
}

}  // namespace

}  // namespace avr

using namespace avr;

extern "C" int avr_crit_irfft(const float* spec, int32_t n_rows, int32_t T, const float* dft, int64_t ldd, float* x, int device,
                              void* stream) {
    AVR_REQUIRE(spec && dft && x, "null pointer");
    AVR_REQUIRE(T >= 4 && T % 2 == 0 && ldd >= T + 2 && ldd % 2 == 0, "T must be even, ldd >= 2F and even");
    AVR_ENTER(device);
    if (n_rows == 0) return AVR_OK;
    const int F = T / 2 + 1;
    irfft_kernel<<<dim3((unsigned)ceil_div(T, 128), (unsigned)n_rows), 128, F * sizeof(float2), (cudaStream_t)stream>>>(
        reinterpret_cast<const float2*>(spec), n_rows, T, dft, ldd, x);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

extern "C" int avr_crit_irfft_adjoint(const float* d_x, int32_t n_rows, int32_t T, const float* dft, int64_t ldd, float* out,
                                      int device, void* stream) {
    AVR_REQUIRE(d_x && dft && out, "null pointer");
    AVR_REQUIRE(T >= 4 && T % 2 == 0 && ldd >= T + 2 && ldd % 2 == 0, "T must be even, ldd >= 2F and even");
    AVR_ENTER(device);
    if (n_rows == 0) return AVR_OK;
    const int F = T / 2 + 1;
    irfft_adjoint_kernel<<<dim3((unsigned)ceil_div(F, 128), (unsigned)n_rows), 128, T * sizeof(float), (cudaStream_t)stream>>>(
        d_x, n_rows, T, dft, ldd, reinterpret_cast<float2*>(out));
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

extern "C" int avr_crit_freq_terms(const float* pred, const float* ori, int32_t bs, int32_t F, float scale_spec, float scale_amp,
                                   float scale_angle, float* partial, float* grad, int device, void* stream) {
    AVR_REQUIRE(pred && ori && partial, "null pointer");
    AVR_ENTER(device);
    if (bs == 0) return AVR_OK;
    freq_terms_kernel<<<(unsigned)bs, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(pred),
                                                                     reinterpret_cast<const float2*>(ori), F, scale_spec, scale_amp,
                                                                     scale_angle, partial, reinterpret_cast<float2*>(grad),
                                                                     (int64_t)bs * F);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

extern "C" int avr_crit_time_l1(const float* x_pred, const float* x_ori, int32_t bs, int32_t T, float scale, float* partial,
                                float* d_x, int device, void* stream) {
    AVR_REQUIRE(x_pred && x_ori && partial, "null pointer");
    AVR_ENTER(device);
    if (bs == 0) return AVR_OK;
    time_l1_kernel<<<(unsigned)bs, 256, 0, (cudaStream_t)stream>>>(x_pred, x_ori, T, scale, partial, d_x);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

extern "C" int64_t avr_crit_energy_workspace_bytes(int32_t bs, int32_t T, int32_t hop) {
    if (bs <= 0 || T <= 0 || hop <= 0) return 0;
    const int64_t M = 1 + T / hop;
    return (int64_t)sizeof(float) * (2 * bs * M * 3 + bs * M + bs * 3 * M);
}

extern "C" int avr_crit_energy(const float* x_pred, const float* x_ori, int32_t bs, int32_t T, int32_t n_fft, int32_t hop,
                               float scale, float* partial, float* d_x, float* workspace, int64_t workspace_bytes, int device,
                               void* stream) {
    AVR_REQUIRE(x_pred && x_ori && partial && workspace, "null pointer");
    AVR_REQUIRE(n_fft >= 2 && n_fft % 2 == 0 && hop >= 1 && n_fft / 2 < T, "bad STFT geometry (reflect padding needs n_fft/2 < T)");
    AVR_REQUIRE(workspace_bytes >= avr_crit_energy_workspace_bytes(bs, T, hop), "workspace too small");
    AVR_ENTER(device);
    if (bs == 0) return AVR_OK;
    const int M = 1 + T / hop;
    float* stats = workspace;
    float* de = stats + (int64_t)2 * bs * M * 3;
    float* scratch = de + (int64_t)bs * M;
    cudaStream_t st = (cudaStream_t)stream;
    frame_energy_kernel<<<dim3((unsigned)M, (unsigned)(2 * bs)), 256, 0, st>>>(x_pred, x_ori, bs, T, n_fft, hop, M, stats);
    AVR_LAUNCH_CHECK();
    edc_kernel<<<(unsigned)bs, 32, 0, st>>>(stats, bs, M, scale, partial, d_x ? de : nullptr, scratch);
    AVR_LAUNCH_CHECK();
    if (d_x) {
        energy_bwd_kernel<<<dim3((unsigned)ceil_div(T, 256), (unsigned)bs), 256, 0, st>>>(x_pred, stats, de, T, n_fft, hop, M, d_x);
        AVR_LAUNCH_CHECK();
    }
    return AVR_OK;
}

static int stft_check(int32_t T, int32_t n_fft, int32_t hop, int32_t win) {
    AVR_REQUIRE(n_fft >= 4 && (n_fft & (n_fft - 1)) == 0 && n_fft <= 4096, "n_fft must be a power of two <= 4096");
    AVR_REQUIRE(win >= 1 && win <= n_fft && hop >= 1 && n_fft / 2 < T, "bad STFT geometry (reflect padding needs n_fft/2 < T)");
    return AVR_OK;
}

extern "C" int avr_crit_stft_fwd(const float* x_pred, const float* x_ori, int32_t bs, int32_t T, int32_t n_fft, int32_t hop,
                                 int32_t win, const float* window, float eps, float* S_pred, float* mag_ori, float* partial,
                                 int device, void* stream) {
    AVR_REQUIRE(x_pred && x_ori && window && S_pred && mag_ori && partial, "null pointer");
    if (int rc = stft_check(T, n_fft, hop, win)) return rc;
    AVR_ENTER(device);
    if (bs == 0) return AVR_OK;
    const int M = 1 + T / hop;
    const size_t smem = sizeof(float) * (2 * (size_t)n_fft + 2 * (size_t)win + 4 * 32);
    AVR_CUDA(cudaFuncSetAttribute(stft_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stft_fwd_kernel<<<dim3((unsigned)M, (unsigned)bs), 128, smem, (cudaStream_t)stream>>>(
        x_pred, x_ori, T, n_fft, hop, win, window, M, eps, reinterpret_cast<float2*>(S_pred), mag_ori, partial);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}

extern "C" int avr_crit_stft_bwd(const float* S_pred, const float* mag_ori, const float* sums, int32_t bs, int32_t T, int32_t n_fft,
                                 int32_t hop, int32_t win, const float* window, float eps, float scale, float w_lin, float* frames,
                                 float* d_x, int device, void* stream) {
    AVR_REQUIRE(S_pred && mag_ori && sums && window && frames && d_x, "null pointer");
    if (int rc = stft_check(T, n_fft, hop, win)) return rc;
    AVR_ENTER(device);
    if (bs == 0) return AVR_OK;
    const int M = 1 + T / hop, K = n_fft / 2 + 1;
    const float inv_n = 1.0f / ((float)bs * (float)M * (float)K);
    const size_t smem = sizeof(float2) * ((size_t)n_fft + (size_t)K);
    cudaStream_t st = (cudaStream_t)stream;
    AVR_CUDA(cudaFuncSetAttribute(stft_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stft_bwd_kernel<<<dim3((unsigned)M, (unsigned)bs), 128, smem, st>>>(reinterpret_cast<const float2*>(S_pred), mag_ori, sums, n_fft,
                                                                      win, window, M, eps, scale, w_lin, inv_n, frames);
    AVR_LAUNCH_CHECK();
    stft_fold_kernel<<<dim3((unsigned)ceil_div(T, 256), (unsigned)bs), 256, 0, st>>>(frames, T, n_fft, hop, win, M, d_x);
    AVR_LAUNCH_CHECK();
    return AVR_OK;
}
