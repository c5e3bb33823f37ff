// STFT / iSTFT on the device (Enhancer.stft / .istft, pb_chime5/core.py:305-321 ->
// nara_wpe.utils.stft / istft, third party; semantics per SURVEY.md appendix A,
// framing pinned by the doctest pb_chime5/database/chime5/database.py:417-453).
//   analysis : periodic Blackman window, zero padding of (size - shift) samples on
//              both ends when fading, zero padding of the last frame, rfft.
//   synthesis: irfft, biorthogonal window w / sum_i w[n mod shift + i shift]^2,
//              overlap-add, fading samples dropped.
// The FFT is a shared-memory Stockham radix-2 transform in float64 (input float32,
// output complex64); a frame tile per CTA keeps the bin-major stores in 64 B runs.
#include "common.cuh"

namespace gss {

constexpr int FFT_NT = 256;
constexpr int FFT_TILE = 8;       // frames per CTA

__device__ __forceinline__ double blackman_periodic(int n, int N) {
    const double a = 2.0 * n / (double)N;      // in units of pi
    return 0.42 - 0.5 * cospi(a) + 0.08 * cospi(2.0 * a);
}

// In-place (ping-pong) complex FFT of length N = 1 << logN held in xa; xb scratch.
// Returns the buffer that holds the result.  tw[q] = exp(-+ 2 pi i q / N), q < N/2.
__device__ inline cd* block_fft(cd* xa, cd* xb, const cd* tw, int N, int tid) {
    cd* x = xa; cd* y = xb;
    for (int l = N / 2, m = 1; l >= 1; l >>= 1, m <<= 1) {
        for (int i = tid; i < N / 2; i += FFT_NT) {
            const int j = i / m, k = i - j * m;
            const cd c0 = x[k + j * m], c1 = x[k + j * m + N / 2];
            const cd w = tw[j * m];
            y[k + 2 * j * m] = cadd(c0, c1);
            y[k + 2 * j * m + m] = cmul(w, csub(c0, c1));
        }
        __syncthreads();
        cd* t = x; x = y; y = t;
    }
    return x;
}

// grid (B*D, ceil(T / FFT_TILE)).  dynamic smem: 2 N cd + N/2 cd + FFT_TILE * (N/2+1) float2 + N doubles
__global__ void __launch_bounds__(FFT_NT) stft_kernel(const float* __restrict__ x, float2* __restrict__ Y,
                                                      int D, int Ns, int T, int size, int shift, int pad) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = size, F = size / 2 + 1;
    cd* xa = reinterpret_cast<cd*>(smem_raw);
    cd* xb = xa + N;
    cd* tw = xb + N;
    double* win = reinterpret_cast<double*>(tw + N / 2);
    float2* stage = reinterpret_cast<float2*>(win + N);            // [FFT_TILE][F]
    const int tid = threadIdx.x;
    const int bd = blockIdx.x;
    const int b = bd / D, d = bd - b * D;
    const int t0 = blockIdx.y * FFT_TILE;
    const float* __restrict__ xs = x + (size_t)bd * Ns;
    for (int q = tid; q < N / 2; q += FFT_NT) {
        double s, c;
        sincospi(-2.0 * q / (double)N, &s, &c);
        tw[q] = cmake(c, s);
    }
    for (int n = tid; n < N; n += FFT_NT) win[n] = blackman_periodic(n, N);
    __syncthreads();
    const int nt = min(FFT_TILE, T - t0);
    for (int tt = 0; tt < nt; ++tt) {
        const long start = (long)(t0 + tt) * shift - pad;
        for (int n = tid; n < N; n += FFT_NT) {
            const long sidx = start + n;
            const double v = (sidx >= 0 && sidx < Ns) ? (double)xs[sidx] : 0.0;
            xa[n] = cmake(v * win[n], 0.0);
        }
        __syncthreads();
        const cd* r = block_fft(xa, xb, tw, N, tid);
        for (int f = tid; f < F; f += FFT_NT) stage[tt * F + f] = make_float2((float)r[f].x, (float)r[f].y);
        __syncthreads();
    }
    // Y[b][f][d][t0 + tt]
    for (int i = tid; i < F * nt; i += FFT_NT) {
        const int f = i / nt, tt = i - f * nt;
        Y[(((size_t)b * F + f) * D + d) * T + t0 + tt] = stage[tt * F + f];
    }
}

// frames[b][t][n] = window_syn[n] * irfft(X[b, :, t])[n]
__global__ void __launch_bounds__(FFT_NT) istft_frames_kernel(const float2* __restrict__ X, float* __restrict__ frames,
                                                              int T, int size, int shift) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = size, F = size / 2 + 1;
    cd* xa = reinterpret_cast<cd*>(smem_raw);
    cd* xb = xa + N;
    cd* tw = xb + N;
    double* win = reinterpret_cast<double*>(tw + N / 2);
    float2* stage = reinterpret_cast<float2*>(win + N);            // [FFT_TILE][F]
    const int tid = threadIdx.x;
    const int b = blockIdx.x;
    const int t0 = blockIdx.y * FFT_TILE;
    for (int q = tid; q < N / 2; q += FFT_NT) {
        double s, c;
        sincospi(2.0 * q / (double)N, &s, &c);
        tw[q] = cmake(c, s);
    }
    for (int n = tid; n < N; n += FFT_NT) win[n] = blackman_periodic(n, N);
    __syncthreads();
    // biorthogonal synthesis window (in place, needs the squared sums per residue)
    double ssq_mine[16];
    {
        int c = 0;
        for (int n = tid; n < N; n += FFT_NT, ++c) {
            double s = 0.0;
            for (int r = n % shift; r < N; r += shift) s += win[r] * win[r];
            ssq_mine[c] = s;
        }
    }
    __syncthreads();
    {
        int c = 0;
        for (int n = tid; n < N; n += FFT_NT, ++c) win[n] = win[n] / ssq_mine[c] / (double)N;   // 1/N of the inverse FFT
    }
    const int nt = min(FFT_TILE, T - t0);
    for (int i = tid; i < F * nt; i += FFT_NT) {
        const int f = i / nt, tt = i - f * nt;
        stage[tt * F + f] = X[((size_t)b * F + f) * T + t0 + tt];
    }
    __syncthreads();
    for (int tt = 0; tt < nt; ++tt) {
        // Hermitian extension (numpy irfft ignores the imaginary part of bin 0 and N/2)
        for (int n = tid; n < N; n += FFT_NT) {
            cd v;
            if (n == 0 || n == N / 2) v = cmake((double)stage[tt * F + n].x, 0.0);
            else if (n < N / 2) v = cmake((double)stage[tt * F + n].x, (double)stage[tt * F + n].y);
            else v = cmake((double)stage[tt * F + (N - n)].x, -(double)stage[tt * F + (N - n)].y);
            xa[n] = v;
        }
        __syncthreads();
        const cd* r = block_fft(xa, xb, tw, N, tid);
        float* out = frames + ((size_t)b * T + t0 + tt) * N;
        for (int n = tid; n < N; n += FFT_NT) out[n] = (float)(r[n].x * win[n]);
        __syncthreads();
    }
}

// out[b][n] = sum over the frames covering sample n + drop (deterministic gather)
__global__ void istft_ola_kernel(const float* __restrict__ frames, float* __restrict__ out,
                                 int T, int size, int shift, int drop, int Nout) {
    const int b = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= Nout) return;
    const int p = n + drop;                       // position in the un-cropped signal
    const int thi = min(p / shift, T - 1);
    double s = 0.0;
    for (int t = thi; t >= 0 && p - t * shift < size; --t)
        s += (double)frames[((size_t)b * T + t) * size + (p - t * shift)];
    out[(size_t)b * Nout + n] = (float)s;
}

static size_t fft_smem(int size) {
    return (size_t)(2 * size + size / 2) * sizeof(cd) + (size_t)size * sizeof(double)
           + (size_t)FFT_TILE * (size / 2 + 1) * sizeof(float2);
}

static bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

size_t istft_ws_bytes(int B, int T, int size) { return align_up((size_t)B * T * size * sizeof(float)); }

}  // namespace gss

extern "C" {

int gss_stft_f32(const float* x, gss_c64* Y, int B, int D, int N, int size, int shift, int fading,
                 void* ws, size_t ws_bytes, void* stream) {
    using namespace gss;
    (void)ws; (void)ws_bytes;
    GSS_REQUIRE(x && Y, GSS_ERR_ARG, "gss_stft_f32: null pointer");
    GSS_REQUIRE(B >= 0 && D > 0 && N > 0, GSS_ERR_ARG, "gss_stft_f32: bad dims");
    GSS_REQUIRE(pow2(size) && size >= 64 && size <= 4096, GSS_ERR_UNSUPPORTED, "gss_stft_f32: size=%d (power of two in [64, 4096])", size);
    GSS_REQUIRE(shift > 0 && size % shift == 0, GSS_ERR_ARG, "gss_stft_f32: shift=%d must divide size=%d", shift, size);
    if (B == 0) return GSS_OK;
    const int pad = fading ? size - shift : 0;
    const long total = (long)N + 2L * pad;
    const int T = total <= size ? 1 : (int)((total - size + shift - 1) / shift + 1);
    const size_t smem = fft_smem(size);
    GSS_CUDA(cudaFuncSetAttribute(stft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(B * D, (T + FFT_TILE - 1) / FFT_TILE);
    stft_kernel<<<grid, FFT_NT, smem, (cudaStream_t)stream>>>(x, (float2*)Y, D, N, T, size, shift, pad);
    GSS_LAUNCH_CHECK("stft_kernel");
    return GSS_OK;
}

int gss_istft_f32(const gss_c64* X, float* x, int B, int T, int size, int shift, int fading,
                  void* ws, size_t ws_bytes, void* stream) {
    using namespace gss;
    GSS_REQUIRE(X && x, GSS_ERR_ARG, "gss_istft_f32: null pointer");
    GSS_REQUIRE(B >= 0 && T > 0, GSS_ERR_ARG, "gss_istft_f32: bad dims");
    GSS_REQUIRE(pow2(size) && size >= 64 && size <= 4096, GSS_ERR_UNSUPPORTED, "gss_istft_f32: size=%d", size);
    GSS_REQUIRE(shift > 0 && size % shift == 0 && size / shift <= 16, GSS_ERR_ARG, "gss_istft_f32: shift=%d", shift);
    if (B == 0) return GSS_OK;
    const size_t need = istft_ws_bytes(B, T, size);
    GSS_REQUIRE(ws && ws_bytes >= need, GSS_ERR_WORKSPACE, "gss_istft_f32: workspace %zu < %zu", ws_bytes, need);
    const int drop = fading ? size - shift : 0;
    const int Nout = T * shift + size - shift - 2 * drop;
    GSS_REQUIRE(Nout > 0, GSS_ERR_ARG, "gss_istft_f32: empty output");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = fft_smem(size);
    GSS_CUDA(cudaFuncSetAttribute(istft_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(B, (T + FFT_TILE - 1) / FFT_TILE);
    istft_frames_kernel<<<grid, FFT_NT, smem, st>>>((const float2*)X, (float*)ws, T, size, shift);
    GSS_LAUNCH_CHECK("istft_frames_kernel");
    dim3 g2((Nout + 255) / 256, B);
    istft_ola_kernel<<<g2, 256, 0, st>>>((const float*)ws, x, T, size, shift, drop, Nout);
    GSS_LAUNCH_CHECK("istft_ola_kernel");
    return GSS_OK;
}

}  // extern "C"
