// Mask-based beamforming on the device.
//   weighted_cov_kernel : get_power_spectral_density_matrix, beamformer.py:61-145
//                         (also the covariance einsum of cACG._fit, cACG.py:293-300)
//   bf_weights_kernel   : get_mvdr_vector_souden :546-617 (+ stable_solve,
//                         math/solve.py:20-114) / get_gev_vector :267-348
//                         (zhegvd, cythonized/get_gev_vector.pyx:42-150)
//   bf_refchan_kernel   : get_optimal_reference_channel :524-543 (reduction over F)
//   bf_apply_kernel     : blind_analytic_normalization :396-418 +
//                         apply_beamforming_vector :502-510 + 'mask_mul' core.py:270-271
// Observations are complex64, all arithmetic is float64.
#include "common.cuh"
#include "smallmat.cuh"

#ifndef GSS_DP_LIST
#define GSS_DP_LIST GSS_CASE(2) GSS_CASE(4) GSS_CASE(6) GSS_CASE(8) GSS_CASE(10) GSS_CASE(12) GSS_CASE(16) GSS_CASE(20) GSS_CASE(24) GSS_CASE(28) GSS_CASE(32)
#endif

namespace gss {

// Source of the per-frame weights of the two (or K) classes.
struct WeightSrc {
    int mode;                  // 0: w (B,F,K,T) f32 ; 1: two masks (B,F,T) ; 2: posterior (B,F,K,T) + target/context
    const float* w;            // mode 0 / 2
    const float* m0;           // mode 1: target mask
    const float* m1;           // mode 1: distortion mask
    const int* target_index;   // mode 2 (B)
    const int* start_ctx;      // mode 2 (B) frames, may be null
    const int* end_ctx;        // mode 2 (B) frames, may be null
    int K;                     // classes in w / posterior
    const int* Tper;           // (B) valid frames per utterance or null
};

__device__ __forceinline__ int valid_frames(const WeightSrc& s, int b, int T) {
    return s.Tper ? min(max(s.Tper[b], 0), T) : T;
}

// weights of class pair (c0, c0+1) for frame t of bin (b,f); mode 1/2 always give (target, distortion)
__device__ __forceinline__ void frame_weights(const WeightSrc& s, int b, size_t bf, int T, int t, int c0,
                                              double& w0, double& w1, int Tv = -1) {
    if (Tv < 0) Tv = T;
    if (s.mode == 0) {
        const float* base = s.w + (bf * s.K) * T;
        w0 = (double)base[(size_t)c0 * T + t];
        w1 = (c0 + 1 < s.K) ? (double)base[(size_t)(c0 + 1) * T + t] : 0.0;
    } else if (s.mode == 1) {
        w0 = (double)s.m0[bf * T + t];
        w1 = (double)s.m1[bf * T + t];
    } else {
        const int sc = s.start_ctx ? s.start_ctx[b] : 0;
        const int ec = s.end_ctx ? s.end_ctx[b] : 0;
        w0 = 0.0; w1 = 0.0;
        if (t >= sc && t < Tv - ec) {           // masks[:, :sc] = 0 ; masks[:, -ec:] = 0 (core.py:545-547)
            const float* base = s.w + (bf * s.K) * T;
            const int ti = s.target_index[b];
            for (int k = 0; k < s.K; ++k) {
                const double v = (double)base[(size_t)k * T + t];
                if (k == ti) w0 = v; else w1 += v;
            }
        }
    }
}

// One CTA per (bin, class pair).  Phi (.., K, D, D) complex128 full Hermitian.
template <int DP, int NT, int TM>
__global__ void __launch_bounds__(NT) weighted_cov_kernel(const float2* __restrict__ Y, WeightSrc src,
                                                          cd* __restrict__ Phi, double* __restrict__ wsum_out,
                                                          int F, int D, int T, int Kout, int normalize) {
    constexpr int NB = DP / 2, G = NB * (NB + 1) / 2, NG = NT / G, YLD = DP + 1;
    static_assert(NG >= 1, "block too small");
    __shared__ __align__(16) cd ysm[TM * YLD];
    __shared__ double wsm[TM * 2];
    __shared__ __align__(16) cd acc_sm[2 * DP * (DP + 1) / 2];
    __shared__ double wred[2 * (NT / 32)];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t bf = blockIdx.x;
    const int b = (int)(bf / F);
    const int c0 = blockIdx.y * 2;
    const float2* __restrict__ Yg = Y + bf * D * T;
    const int Tv = valid_frames(src, b, T);
    const int m_g = tid / G, m_l = tid - m_g * G;
    const bool m_active = m_g < NG;
    int bi = 0;
    while ((bi + 1) * (bi + 2) / 2 <= m_l) ++bi;
    const int bj = m_l - bi * (bi + 1) / 2;
    const int r0 = 2 * bi, cc0 = 2 * bj;
    cd macc[4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a) { macc[a][0] = cmake(0.0, 0.0); macc[a][1] = cmake(0.0, 0.0); }
    double ws0 = 0.0, ws1 = 0.0;
    for (int t0 = 0; t0 < Tv; t0 += TM) {
        for (int i = tid; i < DP * TM; i += NT) {
            const int d = i / TM, t = i - d * TM;
            float2 v = make_float2(0.f, 0.f);
            if (d < D && t0 + t < Tv) v = __ldg(&Yg[(size_t)d * T + t0 + t]);
            ysm[t * YLD + d] = cmake((double)v.x, (double)v.y);
        }
        for (int t = tid; t < TM; t += NT) {
            double w0 = 0.0, w1 = 0.0;
            if (t0 + t < Tv) frame_weights(src, b, bf, T, t0 + t, c0, w0, w1, Tv);
            wsm[2 * t] = w0; wsm[2 * t + 1] = w1;
            ws0 += w0; ws1 += w1;
        }
        __syncthreads();
        if (m_active) {
            const int tn = min(TM, Tv - t0);
            for (int t = m_g; t < tn; t += NG) {
                const cd* yrow = ysm + t * YLD;
                const cd a0 = yrow[r0], a1 = yrow[r0 + 1], b0 = yrow[cc0], b1 = yrow[cc0 + 1];
                cd P[4];
                P[0] = cmulc(a0, b0); P[1] = cmulc(a0, b1);
                P[2] = cmulc(a1, b0); P[3] = cmulc(a1, b1);
                const double w0 = wsm[2 * t], w1 = wsm[2 * t + 1];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    macc[a][0].x = fma(w0, P[a].x, macc[a][0].x); macc[a][0].y = fma(w0, P[a].y, macc[a][0].y);
                    macc[a][1].x = fma(w1, P[a].x, macc[a][1].x); macc[a][1].y = fma(w1, P[a].y, macc[a][1].y);
                }
            }
        }
        __syncthreads();
    }
    constexpr int NP = DP * (DP + 1) / 2;
    for (int g = 0; g < NG; ++g) {
        if (m_active && m_g == g) {
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int d = r0 + (a >> 1), e = cc0 + (a & 1);
                if (e <= d) {
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        cd* dst = &acc_sm[k * NP + tri(d, e)];
                        if (g == 0) *dst = macc[a][k];
                        else { cd v = *dst; v.x += macc[a][k].x; v.y += macc[a][k].y; *dst = v; }
                    }
                }
            }
        }
        __syncthreads();
    }
    ws0 = warp_sum(ws0); ws1 = warp_sum(ws1);
    if (lane == 0) { wred[2 * warp] = ws0; wred[2 * warp + 1] = ws1; }
    __syncthreads();
    double tot[2] = {0.0, 0.0};
    for (int w = 0; w < NT / 32; ++w) { tot[0] += wred[2 * w]; tot[1] += wred[2 * w + 1]; }
    for (int k = 0; k < 2; ++k) {
        if (c0 + k >= Kout) break;
        const double sc = normalize ? 1.0 / fmax(tot[k], 1e-10) : 1.0;     // beamformer.py:124
        cd* out = Phi + (bf * Kout + c0 + k) * D * D;
        for (int i = tid; i < D * D; i += NT) {
            const int r = i / D, c = i - r * D;
            cd v;
            if (r == c) v = cmake(acc_sm[k * NP + tri(r, r)].x, 0.0);
            else if (r > c) v = acc_sm[k * NP + tri(r, c)];
            else v = cconj(acc_sm[k * NP + tri(c, r)]);
            out[i] = cscale(v, sc);
        }
        if (wsum_out && tid == 0) wsum_out[bf * Kout + c0 + k] = tot[k];
    }
}

// ---------------------------------------------------------------------------
// Small channel counts (D <= 8): the covariance is HBM bound, not FP64 bound.  One warp per
// bin, lanes stride the frames (every load is a coalesced 256 B / 128 B row segment), each
// lane keeps the whole packed Hermitian accumulator of KC classes in registers, one
// shuffle-tree reduction per bin.  Grid = ceil(bins / 4) CTAs of 4 warps.
// ---------------------------------------------------------------------------
template <int D, int KC>
__global__ void __launch_bounds__(128) weighted_cov_small_kernel(const float2* __restrict__ Y, WeightSrc src,
                                                                 cd* __restrict__ Phi, int nbins, int F, int T,
                                                                 int Kout, int normalize) {
    constexpr int NPAIR = D * (D + 1) / 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t bf = (size_t)blockIdx.x * 4 + warp;
    if (bf >= (size_t)nbins) return;
    const int b = (int)(bf / F);
    const int c0 = blockIdx.y * KC;
    const float2* __restrict__ Yg = Y + bf * D * T;
    const int Tv = valid_frames(src, b, T);
    double are[KC][NPAIR], aim[KC][NPAIR], wsum[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        wsum[k] = 0.0;
#pragma unroll
        for (int p = 0; p < NPAIR; ++p) { are[k][p] = 0.0; aim[k][p] = 0.0; }
    }
    for (int t = lane; t < Tv; t += 32) {
        double yr[D], yi[D], w[KC];
#pragma unroll
        for (int d = 0; d < D; ++d) { const float2 v = __ldg(&Yg[(size_t)d * T + t]); yr[d] = (double)v.x; yi[d] = (double)v.y; }
        if (KC == 2 && src.mode != 0) frame_weights(src, b, bf, T, t, c0, w[0], w[KC - 1], Tv);
        else {
#pragma unroll
            for (int k = 0; k < KC; ++k)
                w[k] = (c0 + k < src.K) ? (double)src.w[(bf * src.K + c0 + k) * T + t] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < KC; ++k) wsum[k] += w[k];
#pragma unroll
        for (int d = 0; d < D; ++d)
#pragma unroll
            for (int e = 0; e <= d; ++e) {
                const double pre = fma(yr[d], yr[e], yi[d] * yi[e]);
                const double pim = fma(yi[d], yr[e], -(yr[d] * yi[e]));
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    are[k][tri(d, e)] = fma(w[k], pre, are[k][tri(d, e)]);
                    aim[k][tri(d, e)] = fma(w[k], pim, aim[k][tri(d, e)]);
                }
            }
    }
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        wsum[k] = warp_sum(wsum[k]);
#pragma unroll
        for (int p = 0; p < NPAIR; ++p) { are[k][p] = warp_sum(are[k][p]); aim[k][p] = warp_sum(aim[k][p]); }
    }
    // lane p writes row/column entries of pair p (all lanes hold the full sums after the xor tree)
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        if (c0 + k >= Kout) break;
        const double sc = normalize ? 1.0 / fmax(wsum[k], 1e-10) : 1.0;
        cd* out = Phi + (bf * Kout + c0 + k) * D * D;
#pragma unroll
        for (int d = 0; d < D; ++d)
#pragma unroll
            for (int e = 0; e <= d; ++e) {
                if (lane == (tri(d, e) & 31)) {
                    const double re = are[k][tri(d, e)] * sc, im = (d == e) ? 0.0 : aim[k][tri(d, e)] * sc;
                    out[d * D + e] = cmake(re, im);
                    if (d != e) out[e * D + d] = cmake(re, -im);
                }
            }
    }
}

template <int D, int KC>
static int launch_weighted_cov_small(const float2* Y, const WeightSrc& src, cd* Phi, int B, int F, int T, int Kout,
                                     int normalize, cudaStream_t st) {
    const int nbins = B * F;
    dim3 grid((nbins + 3) / 4, (Kout + KC - 1) / KC);
    weighted_cov_small_kernel<D, KC><<<grid, 128, 0, st>>>(Y, src, Phi, nbins, F, T, Kout, normalize);
    GSS_LAUNCH_CHECK("weighted_cov_small_kernel");
    return GSS_OK;
}

template <int DP>
static int launch_weighted_cov(const float2* Y, const WeightSrc& src, cd* Phi, double* wsum,
                               int B, int F, int D, int T, int Kout, int normalize, cudaStream_t st) {
    constexpr int NT = 256, TM = DP > 28 ? 32 : 64;          // static shared memory stays below 48 KB
    dim3 grid(B * F, (Kout + 1) / 2);
    weighted_cov_kernel<DP, NT, TM><<<grid, NT, 0, st>>>(Y, src, Phi, wsum, F, D, T, Kout, normalize);
    GSS_LAUNCH_CHECK("weighted_cov_kernel");
    return GSS_OK;
}

int weighted_cov_dispatch(const float2* Y, const WeightSrc& src, cd* Phi, double* wsum,
                          int B, int F, int D, int T, int Kout, int normalize, cudaStream_t st) {
    if (wsum == nullptr && (D <= 6 || (D <= 8 && src.mode == 0))) {
        // HBM-bound regime: classes per pass limited by the register budget (2 * KC * D(D+1)/2 doubles)
        const bool two = src.mode != 0;          // mask pair (target, distortion)
        switch (D) {
            case 1: return two ? launch_weighted_cov_small<1, 2>(Y, src, Phi, B, F, T, Kout, normalize, st) : launch_weighted_cov_small<1, 4>(Y, src, Phi, B, F, T, Kout, normalize, st);
            case 2: return two ? launch_weighted_cov_small<2, 2>(Y, src, Phi, B, F, T, Kout, normalize, st) : launch_weighted_cov_small<2, 4>(Y, src, Phi, B, F, T, Kout, normalize, st);
            case 3: return two ? launch_weighted_cov_small<3, 2>(Y, src, Phi, B, F, T, Kout, normalize, st) : launch_weighted_cov_small<3, 3>(Y, src, Phi, B, F, T, Kout, normalize, st);
            case 4: return two ? launch_weighted_cov_small<4, 2>(Y, src, Phi, B, F, T, Kout, normalize, st) : launch_weighted_cov_small<4, 3>(Y, src, Phi, B, F, T, Kout, normalize, st);
            case 5: return launch_weighted_cov_small<5, 2>(Y, src, Phi, B, F, T, Kout, normalize, st);
            case 6: return launch_weighted_cov_small<6, 2>(Y, src, Phi, B, F, T, Kout, normalize, st);
            case 7: return launch_weighted_cov_small<7, 1>(Y, src, Phi, B, F, T, Kout, normalize, st);
            case 8: return launch_weighted_cov_small<8, 1>(Y, src, Phi, B, F, T, Kout, normalize, st);
        }
    }
    // any channel count runs on the next built padded size (padded channels are zero rows of the tile)
#define GSS_CASE(dp) if (D <= dp) return launch_weighted_cov<dp>(Y, src, Phi, wsum, B, F, D, T, Kout, normalize, st);
    GSS_DP_LIST
#undef GSS_CASE
    return fail(GSS_ERR_UNSUPPORTED, "weighted covariance: D=%d not built", D);
}

__global__ void c128_to_c64_kernel(const cd* __restrict__ src, float2* __restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = make_float2((float)src[i].x, (float)src[i].y);
}

// ---------------------------------------------------------------------------
// Per-bin beamforming matrices.  One CTA (128 threads) per (b, f).
//   MVDR:  mat = Phi_N^{-1} Phi_X / max(Re tr, eps)   (D x D, all candidate
//          reference columns) + per-column SNR numerators / denominators.
//   GEV:   principal generalised eigenvector (phase-normalised) in column 0.
// ---------------------------------------------------------------------------
constexpr int BFW_NT = 128;

struct BfwSmem {
    static __host__ __device__ size_t bytes(int D) {
        const int ld = D + 1;
        return size_t(4) * D * ld * sizeof(cd) + 64 * sizeof(double) + 16 * sizeof(JacobiRot) + 64 * sizeof(double) + 64
               + 32 * sizeof(cd);
    }
};

// What to compute per bin: the `get_bf_vector` DSL of pb_bss (beamformer_wrapper.py:108-227),
// "[rank1_pca+|rank1_gev+]core[+ban]".  The reference's Beamformer block uses only
// mvdr_souden(+ban) (core.py:263-268) and the wrapper-level gev(+ban) (beamforming_wrapper.py:192-208).
struct BfProgram {
    int rank1;          // 0 none, 1 rank1_pca, 2 rank1_gev  (trace-preserving rank-1 model of Phi_X first)
    int core;           // GSS_BFCORE_*
    int pca_scaling;    // pca: 0 none, 1 'trace', 2 'eigenvalue'   (beamformer.py:183-201)
    int chan;           // chN
    double mu;          // wmwf distortion_weight; < 0: 'frequency_dependent'   (beamformer.py:658-663)
    double eps;         // mvdr_souden: floor of Re tr(phi) and of the SNR denominators
};

struct BfwCtx {
    cd *PX, *PN, *A, *Bm;
    double* red; JacobiRot* rot; double* lam; int* iscr;
    int D, ld, tid, lane, warp;
};

// Bm <- solve(Phi_N, Phi_X) with LAPACK-like partial pivoting; exact-zero pivot -> lstsq fallback
// (solve.py:108-113): minimum-norm solution through eigh of the Hermitian Phi_N, rcond = eps * D
__device__ __forceinline__ void bfw_solve_phi(const BfwCtx& c) {
    const int D = c.D, ld = c.ld, tid = c.tid;
    cd *PX = c.PX, *PN = c.PN, *A = c.A, *Bm = c.Bm;
    for (int i = tid; i < D * D; i += BFW_NT) {
        const int r = i / D, cc = i - r * D;
        A[r * ld + cc] = PN[r * ld + cc]; Bm[r * ld + cc] = PX[r * ld + cc];
    }
    __syncthreads();
    const bool ok = block_lu_solve(A, ld, Bm, ld, D, D, c.iscr, tid, BFW_NT);
    if (!ok) {
        for (int i = tid; i < D * D; i += BFW_NT) {
            const int r = i / D, cc = i - r * D;
            A[r * ld + cc] = PN[r * ld + cc];
        }
        __syncthreads();
        cd* V = Bm;
        block_jacobi_eigh(A, V, D, ld, c.rot, c.red, tid, BFW_NT);
        if (tid == 0) {
            double mx = 0.0;
            for (int i = 0; i < D; ++i) mx = fmax(mx, fabs(A[i * ld + i].x));
            const double cut = mx * 2.220446049250313e-16 * D;
            for (int i = 0; i < D; ++i) {
                const double l = A[i * ld + i].x;
                c.lam[i] = (fabs(l) > cut) ? 1.0 / l : 0.0;
            }
        }
        __syncthreads();
        // A <- V diag(lam) V^H  (pseudo inverse), then Bm <- A Phi_X
        for (int i = tid; i < D * D; i += BFW_NT) {
            const int r = i / D, cc = i - r * D;
            cd sacc = cmake(0.0, 0.0);
            for (int j = 0; j < D; ++j) cfmac(sacc, cscale(V[r * ld + j], c.lam[j]), V[cc * ld + j]);
            A[r * ld + cc] = sacc;
        }
        __syncthreads();
        for (int i = tid; i < D * D; i += BFW_NT) {
            const int r = i / D, cc = i - r * D;
            cd sacc = cmake(0.0, 0.0);
            for (int j = 0; j < D; ++j) cfma(sacc, A[r * ld + j], PX[j * ld + cc]);
            Bm[r * ld + cc] = sacc;
        }
        __syncthreads();
    }
}

// per candidate reference r:  w_r^H Phi_X w_r  and  w_r^H Phi_N w_r   (beamformer.py:535-541); W = Bm
__device__ __forceinline__ void bfw_snr_terms(const BfwCtx& c, double* __restrict__ numden_bf) {
    const int D = c.D, ld = c.ld;
    for (int i = c.tid; i < 2 * D; i += BFW_NT) {
        const int which = i / D, r = i - which * D;
        const cd* Pm = which ? c.PN : c.PX;
        cd sacc = cmake(0.0, 0.0);
        for (int d = 0; d < D; ++d) {
            cd u = cmake(0.0, 0.0);
            for (int e = 0; e < D; ++e) cfma(u, Pm[d * ld + e], c.Bm[e * ld + r]);
            cfma(sacc, cconj(c.Bm[d * ld + r]), u);
        }
        numden_bf[which * D + r] = sacc.x;
    }
}

// GEV: Cholesky Phi_N = L L^H ; C = L^-1 Phi_X L^-H ; eigh(C) ; v = L^-H u, canonical phase
// ((Phi_N v)[0] real, non-negative) -> vv[0..D).  Destroys PX, A, Bm.  false: Phi_N not positive definite.
__device__ __forceinline__ bool bfw_gev(const BfwCtx& c, cd* __restrict__ vv, int* __restrict__ info, int b, int f) {
    const int D = c.D, ld = c.ld, tid = c.tid;
    cd *PX = c.PX, *PN = c.PN, *Bm = c.Bm;
    cd* Lp = c.A;                                  // packed lower
    for (int i = tid; i < D * D; i += BFW_NT) {
        const int r = i / D, cc = i - r * D;
        if (cc <= r) Lp[tri(r, cc)] = (r == cc) ? cmake(PN[r * ld + r].x, 0.0) : PN[r * ld + cc];
    }
    __syncthreads();
    if (c.warp == 0) {
        const bool ok = warp_cholesky_packed(Lp, D, c.lane);
        if (ok) warp_tri_inverse_inplace(Lp, D, c.lane);
        if (c.lane == 0) c.iscr[1] = ok ? 1 : 0;
    }
    __syncthreads();
    if (!c.iscr[1]) {
        if (tid == 0 && info) atomicMax(&info[b], GSS_INFO_NOT_POSDEF | (f << 8));
        return false;
    }
    // Bm <- M Phi_X  (M lower triangular packed in Lp)
    for (int i = tid; i < D * D; i += BFW_NT) {
        const int r = i / D, cc = i - r * D;
        cd sacc = cmake(0.0, 0.0);
        for (int j = 0; j <= r; ++j) cfma(sacc, Lp[tri(r, j)], PX[j * ld + cc]);
        Bm[r * ld + cc] = sacc;
    }
    __syncthreads();
    // PX <- Bm M^H  (Hermitian C), reuse PX storage after a barrier
    cd cval[ (32 * 32 + BFW_NT - 1) / BFW_NT ];
    {
        int n = 0;
        for (int i = tid; i < D * D; i += BFW_NT, ++n) {
            const int r = i / D, cc = i - r * D;
            cd sacc = cmake(0.0, 0.0);
            for (int j = 0; j <= cc; ++j) cfmac(sacc, Bm[r * ld + j], Lp[tri(cc, j)]);
            cval[n] = sacc;
        }
    }
    __syncthreads();
    {
        int n = 0;
        for (int i = tid; i < D * D; i += BFW_NT, ++n) {
            const int r = i / D, cc = i - r * D;
            PX[r * ld + cc] = cval[n];
        }
    }
    __syncthreads();
    // hermitise C exactly
    for (int i = tid; i < D * D; i += BFW_NT) {
        const int r = i / D, cc = i - r * D;
        if (r == cc) Bm[r * ld + cc] = cmake(PX[r * ld + r].x, 0.0);
        else {
            const cd u = PX[r * ld + cc], v = PX[cc * ld + r];
            Bm[r * ld + cc] = cmake(0.5 * (u.x + v.x), 0.5 * (u.y - v.y));
        }
    }
    __syncthreads();
    cd* V = PX;
    const int sweeps = block_jacobi_eigh(Bm, V, D, ld, c.rot, c.red, tid, BFW_NT);
    if (sweeps < 0 && tid == 0 && info) atomicMax(&info[b], GSS_INFO_NO_CONVERGE | (f << 8));
    if (tid == 0) {
        int best = 0; double bv = Bm[0].x;
        for (int i = 1; i < D; ++i) if (Bm[i * ld + i].x > bv) { bv = Bm[i * ld + i].x; best = i; }
        c.iscr[2] = best;
    }
    __syncthreads();
    const int best = c.iscr[2];
    // v = M^H u ; then canonical phase: (Phi_N v)[0] real, non-negative
    if (tid < D) {
        cd sacc = cmake(0.0, 0.0);
        for (int j = tid; j < D; ++j) cfma(sacc, cconj(Lp[tri(j, tid)]), V[j * ld + best]);
        vv[tid] = sacc;
    }
    __syncthreads();
    if (tid == 0) {
        cd z = cmake(0.0, 0.0);
        for (int e = 0; e < D; ++e) cfma(z, PN[0 * ld + e], vv[e]);
        const double az = sqrt(cabs2(z));
        c.red[0] = az > 0.0 ? z.x / az : 1.0;
        c.red[1] = az > 0.0 ? -z.y / az : 0.0;        // conj(phase)
    }
    __syncthreads();
    const cd ph = cmake(c.red[0], c.red[1]);
    __syncthreads();
    if (tid < D) vv[tid] = cmul(vv[tid], ph);
    __syncthreads();
    return true;
}

// principal eigenvector (largest eigenvalue, np.linalg.eigh semantics up to the phase, which LAPACK
// leaves unspecified: here the largest component is made real and positive) of the Hermitian PX.
// Destroys A, Bm; PX stays.  vv[0..D) <- unit vector, returns the eigenvalue.
__device__ __forceinline__ double bfw_principal(const BfwCtx& c, cd* __restrict__ vv) {
    const int D = c.D, ld = c.ld, tid = c.tid;
    for (int i = tid; i < D * D; i += BFW_NT) {
        const int r = i / D, cc = i - r * D;
        cd v;
        if (r == cc) v = cmake(c.PX[r * ld + r].x, 0.0);
        else if (r > cc) v = c.PX[r * ld + cc];
        else v = cconj(c.PX[cc * ld + r]);             // eigh reads the lower triangle
        c.A[r * ld + cc] = v;
    }
    __syncthreads();
    block_jacobi_eigh(c.A, c.Bm, D, ld, c.rot, c.red, tid, BFW_NT);
    if (tid == 0) {
        int best = 0; double bv = c.A[0].x;
        for (int i = 1; i < D; ++i) if (c.A[i * ld + i].x > bv) { bv = c.A[i * ld + i].x; best = i; }
        int big = 0; double bm = 0.0;
        for (int i = 0; i < D; ++i) { const double m = cabs2(c.Bm[i * ld + best]); if (m > bm) { bm = m; big = i; } }
        const cd z = c.Bm[big * ld + best];
        const double az = sqrt(cabs2(z));
        c.iscr[2] = best;
        c.red[0] = az > 0.0 ? z.x / az : 1.0; c.red[1] = az > 0.0 ? -z.y / az : 0.0; c.red[2] = bv;
    }
    __syncthreads();
    const int best = c.iscr[2];
    const cd ph = cmake(c.red[0], c.red[1]);
    const double lam = c.red[2];
    __syncthreads();
    if (tid < D) vv[tid] = cmul(c.Bm[tid * ld + best], ph);
    __syncthreads();
    return lam;
}

__global__ void __launch_bounds__(BFW_NT) bf_weights_kernel(const cd* __restrict__ PhiX /*(B,F,[2,]D,D)*/,
                                                            const cd* __restrict__ PhiN,
                                                            size_t bin_stride /* elements between bins */,
                                                            cd* __restrict__ mat /*(B,F,D,D)*/,
                                                            double* __restrict__ numden /*(B,F,2,D)*/,
                                                            int* __restrict__ info, int F, int D, BfProgram prog) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ld = D + 1;
    BfwCtx c;
    c.PX = reinterpret_cast<cd*>(smem_raw);         // Phi_X
    c.PN = c.PX + D * ld;                            // Phi_N
    c.A = c.PN + D * ld;                             // work
    c.Bm = c.A + D * ld;                             // work / result
    c.red = reinterpret_cast<double*>(c.Bm + D * ld);           // [64]
    c.rot = reinterpret_cast<JacobiRot*>(c.red + 64);           // [16]
    c.lam = reinterpret_cast<double*>(c.rot + 16);              // [64]
    c.iscr = reinterpret_cast<int*>(c.lam + 64);                // [16]
    cd* vv = reinterpret_cast<cd*>(c.iscr + 16);                // [32] vector scratch
    c.D = D; c.ld = ld; c.tid = threadIdx.x; c.lane = c.tid & 31; c.warp = c.tid >> 5;
    const int tid = c.tid;
    cd *PX = c.PX, *PN = c.PN, *Bm = c.Bm;
    const size_t bf = blockIdx.x;
    const int b = (int)(bf / F), f = (int)(bf - (size_t)b * F);
    const cd* gx = PhiX + bf * bin_stride;
    const cd* gn = PhiN ? PhiN + bf * bin_stride : nullptr;
    for (int i = tid; i < D * D; i += BFW_NT) {
        const int r = i / D, cc = i - r * D;
        PX[r * ld + cc] = gx[i]; PN[r * ld + cc] = gn ? gn[i] : cmake(r == cc ? 1.0 : 0.0, 0.0);
    }
    __syncthreads();
    cd* out = mat + bf * D * D;
    double* nd = numden + bf * 2 * D;

    // ---- optional rank-1 model of Phi_X: scale a a^H with the trace of Phi_X (beamformer_wrapper.py:11-62) ----
    if (prog.rank1 != 0) {
        cd trx = cmake(0.0, 0.0);
        if (tid < D) trx = PX[tid * ld + tid];
        const double trr = block_sum(trx.x, c.red, tid, BFW_NT), tri_ = block_sum(trx.y, c.red, tid, BFW_NT);
        bool ok = true;
        if (prog.rank1 == 1) bfw_principal(c, vv);
        else {
            ok = bfw_gev(c, vv, info, b, f);           // w_gev; ATF estimate a = Phi_N w   (beamformer_wrapper.py:24-44)
            if (ok) {
                cd a = cmake(0.0, 0.0);
                if (tid < D) for (int e = 0; e < D; ++e) cfma(a, PN[tid * ld + e], vv[e]);
                __syncthreads();
                if (tid < D) vv[tid] = a;
                __syncthreads();
            }
        }
        if (!ok) {
            for (int i = tid; i < D * D; i += BFW_NT) out[i] = cmake(0.0, 0.0);
            for (int i = tid; i < 2 * D; i += BFW_NT) nd[i] = 0.0;
            return;
        }
        double n2 = tid < D ? cabs2(vv[tid]) : 0.0;
        n2 = block_sum(n2, c.red, tid, BFW_NT);
        const cd scale = cmake(trr / n2, tri_ / n2);
        for (int i = tid; i < D * D; i += BFW_NT) {
            const int r = i / D, cc = i - r * D;
            PX[r * ld + cc] = cmul(scale, cmulc(vv[r], vv[cc]));
        }
        __syncthreads();
    }

    if (prog.core == GSS_BFCORE_MVDR_SOUDEN || prog.core == GSS_BFCORE_WMWF) {
        bfw_solve_phi(c);
        cd tr = cmake(0.0, 0.0);
        if (tid < D) tr = Bm[tid * ld + tid];
        const double trr = block_sum(tr.x, c.red, tid, BFW_NT);
        cd sc;
        if (prog.core == GSS_BFCORE_MVDR_SOUDEN) {
            sc = cmake(1.0 / fmax(trr, prog.eps), 0.0);            // mat = phi / max(Re tr(phi), eps)   (beamformer.py:604-607)
        } else {
            const double tri_ = block_sum(tr.y, c.red, tid, BFW_NT);
            cd den;
            if (prog.mu < 0.0) {                                   // 'frequency_dependent': sqrt(Phi_X[0,0] * lambda)
                const cd z = cmul(PX[0], cmake(trr, tri_));
                const double az = sqrt(cabs2(z));
                const double re = sqrt(0.5 * (az + z.x)), im = sqrt(fmax(0.5 * (az - z.x), 0.0));
                den = cmake(re, z.y < 0.0 ? -im : im);             // principal complex square root
            } else {
                den = cmake(prog.mu + trr, tri_);                  // distortion_weight + lambda   (beamformer.py:663)
            }
            const double d2 = 1.0 / cabs2(den);
            sc = cmake(den.x * d2, -den.y * d2);
        }
        for (int i = tid; i < D * D; i += BFW_NT) {
            const int r = i / D, cc = i - r * D;
            const cd v = cmul(Bm[r * ld + cc], sc);
            Bm[r * ld + cc] = v;
            out[i] = v;
        }
        __syncthreads();
        bfw_snr_terms(c, nd);
        return;
    }

    // ---- cores that yield one vector: it goes to column 0 of `mat` ----
    bool ok = true;
    if (prog.core == GSS_BFCORE_GEV) {
        ok = bfw_gev(c, vv, info, b, f);
    } else if (prog.core == GSS_BFCORE_PCA) {
        const double lam = bfw_principal(c, vv);
        if (prog.pca_scaling != 0) {                               // beamformer.py:183-201 (the eigenvector has unit norm)
            double trx = tid < D ? PX[tid * ld + tid].x : 0.0;
            trx = block_sum(trx, c.red, tid, BFW_NT);
            const double s = prog.pca_scaling == 1 ? sqrt(trx) : lam;
            if (tid < D) vv[tid] = cscale(vv[tid], s);
            __syncthreads();
        }
    } else if (prog.core == GSS_BFCORE_PCA_MVDR || prog.core == GSS_BFCORE_GEVATF_MVDR) {
        // ATF estimate, then Phi_N^-1 a / (a^H Phi_N^-1 a) with Phi_N hermitised (beamformer.py:205-235)
        if (prog.core == GSS_BFCORE_PCA_MVDR) bfw_principal(c, vv);
        else {
            ok = bfw_gev(c, vv, info, b, f);
            if (ok) {
                cd a = cmake(0.0, 0.0);
                if (tid < D) for (int e = 0; e < D; ++e) cfma(a, PN[tid * ld + e], vv[e]);
                __syncthreads();
                if (tid < D) vv[tid] = a;
                __syncthreads();
            }
        }
        if (ok) {
            for (int i = tid; i < D * D; i += BFW_NT) {
                const int r = i / D, cc = i - r * D;
                const cd u = PN[r * ld + cc], v = PN[cc * ld + r];
                c.A[r * ld + cc] = cmake(0.5 * (u.x + v.x), 0.5 * (u.y - v.y));
            }
            if (tid < D) Bm[tid * ld] = vv[tid];                   // one right-hand side in column 0
            __syncthreads();
            block_lu_solve(c.A, ld, Bm, ld, D, 1, c.iscr, tid, BFW_NT);
            cd den = cmake(0.0, 0.0);
            if (tid < D) den = ccmul(vv[tid], Bm[tid * ld]);
            const double dr = block_sum(den.x, c.red, tid, BFW_NT), di = block_sum(den.y, c.red, tid, BFW_NT);
            const double d2 = 1.0 / (dr * dr + di * di);
            const cd inv = cmake(dr * d2, -di * d2);
            __syncthreads();
            if (tid < D) vv[tid] = cmul(Bm[tid * ld], inv);
            __syncthreads();
        }
    } else {                                                       // chN
        if (tid < D) vv[tid] = cmake(tid == prog.chan ? 1.0 : 0.0, 0.0);
        __syncthreads();
    }
    for (int i = tid; i < D * D; i += BFW_NT) {
        const int r = i / D, cc = i - r * D;
        out[i] = (ok && cc == 0) ? vv[r] : cmake(0.0, 0.0);
    }
    for (int i = tid; i < 2 * D; i += BFW_NT) nd[i] = 0.0;
}

// One CTA per utterance: SNR_r = sum_f num / max(sum_f den, eps); argmax (first max wins, np.argmax).
__global__ void bf_refchan_kernel(const double* __restrict__ numden, int* __restrict__ ref, int* __restrict__ info,
                                  int F, int D, double eps) {
    __shared__ double snr[64];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nw = blockDim.x / 32;
    for (int r = warp; r < D; r += nw) {
        double n = 0.0, d = 0.0;
        for (int f = lane; f < F; f += 32) {
            n += numden[(((size_t)b * F + f) * 2 + 0) * D + r];
            d += numden[(((size_t)b * F + f) * 2 + 1) * D + r];
        }
        n = warp_sum(n); d = warp_sum(d);
        if (lane == 0) snr[r] = n / fmax(d, eps);
    }
    __syncthreads();
    if (tid == 0) {
        int best = 0; bool finite = true;
        for (int r = 0; r < D; ++r) {
            if (!isfinite(snr[r])) finite = false;
            if (snr[r] > snr[best]) best = r;
        }
        if (!finite && info) atomicMax(&info[b], GSS_INFO_NONFINITE);
        ref[b] = best;
    }
}

// One CTA per (b, f): pick column, BAN, apply, postfilter.
__global__ void __launch_bounds__(256) bf_apply_kernel(const float2* __restrict__ Y, const cd* __restrict__ Phi,
                                                       const cd* __restrict__ mat, const int* __restrict__ ref,
                                                       WeightSrc src, float2* __restrict__ Xhat,
                                                       cd* __restrict__ weights_out,
                                                       int F, int D, int T, int fixed_col, int ban, int postfilter) {
    __shared__ cd w[32];
    __shared__ cd pw[32];
    __shared__ double scal[2];
    const int tid = threadIdx.x;
    const size_t bf = blockIdx.x;
    const int b = (int)(bf / F);
    const int col = fixed_col >= 0 ? fixed_col : ref[b];
    if (tid < D) w[tid] = mat[bf * D * D + (size_t)tid * D + col];
    __syncthreads();
    if (ban) {
        const cd* PN = Phi + (bf * 2 + 1) * D * D;
        if (tid < D) {
            cd s = cmake(0.0, 0.0);
            for (int e = 0; e < D; ++e) cfma(s, PN[tid * D + e], w[e]);
            pw[tid] = s;                                 // (Phi_N w)[tid]
        }
        __syncthreads();
        if (tid == 0) {
            double nom = 0.0; cd den = cmake(0.0, 0.0);
            for (int d = 0; d < D; ++d) { nom += cabs2(pw[d]); cfma(den, cconj(w[d]), pw[d]); }
            const double ad = sqrt(cabs2(den));
            scal[0] = ad != 0.0 ? sqrt(nom) / ad : 0.0;  // beamformer.py:406-417
        }
        __syncthreads();
        if (tid < D) w[tid] = cscale(w[tid], scal[0]);
        __syncthreads();
    }
    if (weights_out && tid < D) weights_out[bf * D + tid] = w[tid];
    const float2* __restrict__ Yg = Y + bf * D * T;
    const int Tv = valid_frames(src, b, T);
    for (int t = tid; t < T; t += blockDim.x) {
        if (t >= Tv) { Xhat[bf * T + t] = make_float2(0.f, 0.f); continue; }
        cd s = cmake(0.0, 0.0);
        for (int d = 0; d < D; ++d) {
            const float2 v = __ldg(&Yg[(size_t)d * T + t]);
            cfma(s, cconj(w[d]), cmake((double)v.x, (double)v.y));
        }
        if (postfilter == GSS_POSTFILTER_MASK_MUL) {
            double w0, w1;
            frame_weights(src, b, bf, T, t, 0, w0, w1, Tv);
            s = cscale(s, w0);
        }
        Xhat[bf * T + t] = make_float2((float)s.x, (float)s.y);
    }
}

// 'ch' / 'sum' (core.py:259-262)
__global__ void bf_simple_kernel(const float2* __restrict__ Y, WeightSrc src, float2* __restrict__ Xhat,
                                 int F, int D, int T, int ch, int postfilter) {
    const size_t bf = blockIdx.x;
    const int b = (int)(bf / F);
    const float2* __restrict__ Yg = Y + bf * D * T;
    const int Tv = valid_frames(src, b, T);
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        if (t >= Tv) { Xhat[bf * T + t] = make_float2(0.f, 0.f); continue; }
        double re = 0.0, im = 0.0;
        if (ch >= 0) { const float2 v = Yg[(size_t)ch * T + t]; re = v.x; im = v.y; }
        else for (int d = 0; d < D; ++d) { const float2 v = Yg[(size_t)d * T + t]; re += v.x; im += v.y; }
        if (postfilter == GSS_POSTFILTER_MASK_MUL) {
            double w0, w1;
            frame_weights(src, b, bf, T, t, 0, w0, w1, Tv);
            re *= w0; im *= w0;
        }
        Xhat[bf * T + t] = make_float2((float)re, (float)im);
    }
}

static size_t bf_ws_layout(int B, int F, int D, cd** Phi, cd** mat, double** numden, int** ref, void* ws) {
    Arena a(ws, ~size_t(0));
    cd* p1 = a.take<cd>((size_t)B * F * 2 * D * D);
    cd* p2 = a.take<cd>((size_t)B * F * D * D);
    double* p3 = a.take<double>((size_t)B * F * 2 * D);
    int* p4 = a.take<int>(B);
    if (Phi) *Phi = p1; if (mat) *mat = p2; if (numden) *numden = p3; if (ref) *ref = p4;
    return a.off;
}

size_t beamform_ws_bytes(int B, int F, int D) { return bf_ws_layout(B, F, D, nullptr, nullptr, nullptr, nullptr, nullptr); }
size_t weighted_cov_ws_bytes(int B, int F, int D, int K) { return align_up((size_t)B * F * K * D * D * sizeof(cd)); }

static int beamform_impl(const float2* Y, const WeightSrc& src, float2* Xhat, int bf_type, int bf_arg,
                         int postfilter, int B, int F, int D, int T, int* ref_out, double* weights_out,
                         int* info, void* ws, size_t ws_bytes, cudaStream_t st) {
    GSS_REQUIRE(Y && Xhat, GSS_ERR_ARG, "beamform: null pointer");
    GSS_REQUIRE(B >= 0 && F >= 0 && D > 0 && T > 0, GSS_ERR_ARG, "beamform: bad dims");
    GSS_REQUIRE(postfilter == GSS_POSTFILTER_NONE || postfilter == GSS_POSTFILTER_MASK_MUL, GSS_ERR_UNSUPPORTED,
                "postfilter %d (core.py:272-273)", postfilter);
    if (B == 0 || F == 0) return GSS_OK;
    if (bf_type == GSS_BF_CH || bf_type == GSS_BF_SUM) {
        if (bf_type == GSS_BF_CH) GSS_REQUIRE(bf_arg >= 0 && bf_arg < D, GSS_ERR_ARG, "channel %d out of range for D=%d", bf_arg, D);
        bf_simple_kernel<<<B * F, 256, 0, st>>>(Y, src, Xhat, F, D, T, bf_type == GSS_BF_CH ? bf_arg : -1, postfilter);
        GSS_LAUNCH_CHECK("bf_simple_kernel");
        return GSS_OK;
    }
    BfProgram prog{0, GSS_BFCORE_MVDR_SOUDEN, 0, 0, 1.0, 1e-10};      // eps = 1e-10: beamforming_wrapper.py:75-83
    bool ban;
    if (bf_type & GSS_BF_PROGRAM_FLAG) {
        // a `get_bf_vector` program (beamformer_wrapper.py:108-227) with the reference's default keyword arguments
        prog.core = bf_type & 0xF; prog.rank1 = (bf_type >> 4) & 3; ban = (bf_type >> 6) & 1;
        prog.chan = bf_arg; prog.eps = GSS_F64_TINY;
        GSS_REQUIRE(prog.core <= GSS_BFCORE_CH && prog.rank1 <= 2, GSS_ERR_UNSUPPORTED, "beamformer program 0x%x", bf_type);
        if (prog.core == GSS_BFCORE_CH) GSS_REQUIRE(bf_arg >= 0 && bf_arg < D, GSS_ERR_ARG, "channel %d out of range for D=%d", bf_arg, D);
    } else {
        const bool gev = bf_type == GSS_BF_GEV_BAN || bf_type == GSS_BF_GEV;
        const bool mvdr = bf_type == GSS_BF_MVDR_SOUDEN_BAN || bf_type == GSS_BF_MVDR_SOUDEN;
        GSS_REQUIRE(gev || mvdr, GSS_ERR_UNSUPPORTED, "beamformer type %d (core.py:263-264)", bf_type);
        ban = bf_type == GSS_BF_MVDR_SOUDEN_BAN || bf_type == GSS_BF_GEV_BAN;
        if (gev) prog.core = GSS_BFCORE_GEV;
    }
    const bool matrix_core = prog.core == GSS_BFCORE_MVDR_SOUDEN || prog.core == GSS_BFCORE_WMWF;
    GSS_REQUIRE(D < 30, GSS_ERR_ARG, "D=%d: beamformer needs D < 30 (beamforming_wrapper.py:44)", D);
    cd *Phi, *mat; double* numden; int* ref;
    const size_t need = bf_ws_layout(B, F, D, &Phi, &mat, &numden, &ref, ws);
    GSS_REQUIRE(ws && ws_bytes >= need, GSS_ERR_WORKSPACE, "beamform: workspace %zu < %zu", ws_bytes, need);
    int rc = weighted_cov_dispatch(Y, src, Phi, nullptr, B, F, D, T, 2, 1, st);
    if (rc) return rc;
    const size_t smem = BfwSmem::bytes(D);
    GSS_CUDA(cudaFuncSetAttribute(bf_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bf_weights_kernel<<<B * F, BFW_NT, smem, st>>>(Phi, Phi + (size_t)D * D, (size_t)2 * D * D, mat, numden, info, F, D, prog);
    GSS_LAUNCH_CHECK("bf_weights_kernel");
    if (matrix_core) {
        bf_refchan_kernel<<<B, 256, 0, st>>>(numden, ref, info, F, D, prog.eps);
        GSS_LAUNCH_CHECK("bf_refchan_kernel");
        if (ref_out) GSS_CUDA(cudaMemcpyAsync(ref_out, ref, sizeof(int) * B, cudaMemcpyDeviceToDevice, st));
    }
    bf_apply_kernel<<<B * F, 256, 0, st>>>(Y, Phi, mat, ref, src, Xhat, reinterpret_cast<cd*>(weights_out),
                                           F, D, T, matrix_core ? -1 : 0, ban ? 1 : 0, postfilter);
    GSS_LAUNCH_CHECK("bf_apply_kernel");
    return GSS_OK;
}

// column `ref` of mat (or a fixed column), optional blind analytic normalisation -> w (B,F,D)
__global__ void bf_vector_finish_kernel(const cd* __restrict__ mat, const cd* __restrict__ PhiN, const int* __restrict__ ref,
                                        cd* __restrict__ w_out, int F, int D, int fixed_col, int ban) {
    __shared__ cd w[32];
    __shared__ cd pw[32];
    __shared__ double scal;
    const int tid = threadIdx.x;
    const size_t bf = blockIdx.x;
    const int b = (int)(bf / F);
    const int col = fixed_col >= 0 ? fixed_col : ref[b];
    if (tid < D) w[tid] = mat[bf * D * D + (size_t)tid * D + col];
    __syncthreads();
    if (ban) {
        const cd* PN = PhiN + bf * D * D;
        if (tid < D) {
            cd s = cmake(0.0, 0.0);
            for (int e = 0; e < D; ++e) cfma(s, PN[tid * D + e], w[e]);
            pw[tid] = s;
        }
        __syncthreads();
        if (tid == 0) {
            double nom = 0.0; cd den = cmake(0.0, 0.0);
            for (int d = 0; d < D; ++d) { nom += cabs2(pw[d]); cfma(den, cconj(w[d]), pw[d]); }
            const double ad = sqrt(cabs2(den));
            scal = ad != 0.0 ? sqrt(nom) / ad : 0.0;       // beamformer.py:406-417
        }
        __syncthreads();
        if (tid < D) w[tid] = cscale(w[tid], scal);
    }
    if (tid < D) w_out[bf * D + tid] = w[tid];
}

size_t bf_vector_ws_bytes(int B, int F, int D) {
    return align_up((size_t)B * F * D * D * sizeof(cd)) + align_up((size_t)B * F * 2 * D * sizeof(double)) + align_up((size_t)B * sizeof(int));
}

}  // namespace gss

extern "C" {

int gss_bf_vector_c128(const double* Phi_X, const double* Phi_N, double* w_out,
                       int core, int rank1, int ban, int ref_channel, double distortion_weight,
                       int pca_scaling, int channel, int B, int F, int D,
                       int* ref_channel_out, int* info, void* ws, size_t ws_bytes, void* stream) {
    using namespace gss;
    GSS_REQUIRE(Phi_X && w_out, GSS_ERR_ARG, "gss_bf_vector_c128: null pointer");
    GSS_REQUIRE(B >= 0 && F >= 0 && D > 0, GSS_ERR_ARG, "gss_bf_vector_c128: bad dims");
    GSS_REQUIRE(D <= 32, GSS_ERR_UNSUPPORTED, "gss_bf_vector_c128: D=%d > 32 not built", D);
    GSS_REQUIRE(core >= GSS_BFCORE_MVDR_SOUDEN && core <= GSS_BFCORE_CH, GSS_ERR_ARG, "gss_bf_vector_c128: core %d", core);
    GSS_REQUIRE(rank1 >= 0 && rank1 <= 2 && pca_scaling >= 0 && pca_scaling <= 2, GSS_ERR_ARG, "gss_bf_vector_c128: rank1 / scaling");
    const bool needs_noise = ban || rank1 == 2 || (core != GSS_BFCORE_PCA && core != GSS_BFCORE_CH);
    GSS_REQUIRE(Phi_N || !needs_noise, GSS_ERR_ARG, "gss_bf_vector_c128: this beamformer needs the noise PSD matrix");
    if (core == GSS_BFCORE_CH) GSS_REQUIRE(channel >= 0 && channel < D, GSS_ERR_ARG, "channel %d out of range for D=%d", channel, D);
    GSS_REQUIRE(ref_channel < D, GSS_ERR_ARG, "ref_channel %d out of range for D=%d", ref_channel, D);
    if (B == 0 || F == 0) return GSS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    GSS_REQUIRE(ws && ws_bytes >= bf_vector_ws_bytes(B, F, D), GSS_ERR_WORKSPACE, "gss_bf_vector_c128: workspace %zu < %zu",
                ws_bytes, bf_vector_ws_bytes(B, F, D));
    Arena a(ws, ws_bytes);
    cd* mat = a.take<cd>((size_t)B * F * D * D);
    double* numden = a.take<double>((size_t)B * F * 2 * D);
    int* ref = a.take<int>(B);
    // reference defaults: eps = smallest positive float64 (beamformer.py:602-603, 524-526)
    BfProgram prog{rank1, core, pca_scaling, channel, distortion_weight, GSS_F64_TINY};
    const size_t smem = BfwSmem::bytes(D);
    GSS_CUDA(cudaFuncSetAttribute(bf_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bf_weights_kernel<<<B * F, BFW_NT, smem, st>>>(reinterpret_cast<const cd*>(Phi_X), reinterpret_cast<const cd*>(Phi_N),
                                                   (size_t)D * D, mat, numden, info, F, D, prog);
    GSS_LAUNCH_CHECK("bf_weights_kernel");
    const bool matrix_core = core == GSS_BFCORE_MVDR_SOUDEN || core == GSS_BFCORE_WMWF;
    int fixed_col = 0;
    if (matrix_core) {
        fixed_col = ref_channel;                       // -1: SNR-optimal reference channel over all bins of the utterance
        if (ref_channel < 0) {
            bf_refchan_kernel<<<B, 256, 0, st>>>(numden, ref, info, F, D, GSS_F64_TINY);
            GSS_LAUNCH_CHECK("bf_refchan_kernel");
            if (ref_channel_out) GSS_CUDA(cudaMemcpyAsync(ref_channel_out, ref, sizeof(int) * B, cudaMemcpyDeviceToDevice, st));
        }
    }
    bf_vector_finish_kernel<<<B * F, 32, 0, st>>>(mat, reinterpret_cast<const cd*>(Phi_N), ref, reinterpret_cast<cd*>(w_out),
                                                  F, D, fixed_col, ban ? 1 : 0);
    GSS_LAUNCH_CHECK("bf_vector_finish_kernel");
    return GSS_OK;
}

int gss_weighted_cov_c64(const gss_c64* Y, const float* w, gss_c64* Phi, int normalize_mode,
                         int B, int F, int D, int T, int K, const int* T_per_utt, void* ws, size_t ws_bytes, void* stream) {
    using namespace gss;
    GSS_REQUIRE(Y && w && Phi, GSS_ERR_ARG, "gss_weighted_cov_c64: null pointer");
    GSS_REQUIRE(B >= 0 && F >= 0 && D > 0 && T > 0 && K > 0, GSS_ERR_ARG, "gss_weighted_cov_c64: bad dims");
    GSS_REQUIRE(D <= 32, GSS_ERR_UNSUPPORTED, "gss_weighted_cov_c64: D=%d > 32 not built", D);
    GSS_REQUIRE(normalize_mode == 0 || normalize_mode == 1, GSS_ERR_ARG, "normalize_mode %d", normalize_mode);
    if (B == 0 || F == 0) return GSS_OK;
    const size_t need = weighted_cov_ws_bytes(B, F, D, K);
    GSS_REQUIRE(ws && ws_bytes >= need, GSS_ERR_WORKSPACE, "gss_weighted_cov_c64: workspace %zu < %zu", ws_bytes, need);
    WeightSrc src{}; src.mode = 0; src.w = w; src.K = K; src.Tper = T_per_utt;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = weighted_cov_dispatch((const float2*)Y, src, (cd*)ws, nullptr, B, F, D, T, K, normalize_mode, st);
    if (rc) return rc;
    const size_t n = (size_t)B * F * K * D * D;
    c128_to_c64_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>((const cd*)ws, (float2*)Phi, n);
    GSS_LAUNCH_CHECK("c128_to_c64_kernel");
    return GSS_OK;
}

int gss_beamform_c64(const gss_c64* Y, const float* target_mask, const float* distortion_mask, gss_c64* X_hat,
                     int bf_type, int bf_arg, int postfilter, int B, int F, int D, int T, const int* T_per_utt,
                     int* ref_channel_out, double* weights_out, int* info, void* ws, size_t ws_bytes, void* stream) {
    using namespace gss;
    GSS_REQUIRE(target_mask && distortion_mask, GSS_ERR_ARG, "gss_beamform_c64: null mask");
    WeightSrc src{}; src.mode = 1; src.m0 = target_mask; src.m1 = distortion_mask; src.K = 2; src.Tper = T_per_utt;
    return beamform_impl((const float2*)Y, src, (float2*)X_hat, bf_type, bf_arg, postfilter, B, F, D, T,
                         ref_channel_out, weights_out, info, ws, ws_bytes, (cudaStream_t)stream);
}

int gss_beamform_from_posterior_c64(const gss_c64* Y, const float* posterior, const int* target_index,
                                    const int* start_ctx, const int* end_ctx, gss_c64* X_hat,
                                    int bf_type, int bf_arg, int postfilter, int B, int F, int D, int T, int K,
                                    const int* T_per_utt, int* ref_channel_out, double* weights_out, int* info,
                                    void* ws, size_t ws_bytes, void* stream) {
    using namespace gss;
    GSS_REQUIRE(posterior && target_index, GSS_ERR_ARG, "gss_beamform_from_posterior_c64: null pointer");
    GSS_REQUIRE(K > 1 && K < 20, GSS_ERR_ARG, "K=%d", K);
    WeightSrc src{}; src.mode = 2; src.w = posterior; src.target_index = target_index;
    src.start_ctx = start_ctx; src.end_ctx = end_ctx; src.K = K; src.Tper = T_per_utt;
    return beamform_impl((const float2*)Y, src, (float2*)X_hat, bf_type, bf_arg, postfilter, B, F, D, T,
                         ref_channel_out, weights_out, info, ws, ws_bytes, (cudaStream_t)stream);
}

}  // extern "C"
