// Guided CACGMM EM for the shapes the fused fast kernel (cacgmm.cu) is not instantiated for:
// any D < 35 and any K < 20 (the reference's own limits, cacgmm.py:247-248), runtime D and K.
// Same entry point, same results; reached when K > 6 (RTTM diarisation output with more than five
// speakers + noise) or D > 24.  Off the fast path: one CTA per (utterance, bin), frames tiled through
// shared memory, classes handled four at a time in the E phase and one entry at a time in the M
// phase; the class matrices ALWAYS take the reference's literal route
// (eigh -> normalise by the largest eigenvalue -> floor -> V diag(1/lambda) V^H,
// complex_angular_central_gaussian.py:81-131,185-201) through the CTA-wide Jacobi solver, warm
// started from the previous pass.  Everything float64 on complex64 observations, like cacgmm.cu.
#include "common.cuh"
#include "smallmat.cuh"

namespace gss {

struct CacgmmGenericParams {
    const void* Y;            // (B,F,D,T) complex64 or complex128 (template parameter of the kernel)
    const uint8_t* activity;  // (B,K,T_act)
    const int* Tper;
    float* posterior;         // (B,F,K,T)
    double* weight_out;
    double* logdet_out;
    double* cov_out;
    int* info;
    cd* Bmat;                 // (B*F, K, NP) packed B' = Phi^-1 (off-diagonals doubled)
    cd* Phi;                  // (B*F, K, NP) packed accumulators
    cd* Vwarm;                // (B*F, K, D, D+1) eigenvectors of the previous pass
    int B, F, D, K, T, T_act;
    int iterations, iterations_post;
    double eps, floor_;
};

constexpr int GEN_NT = 256;
constexpr int GEN_TT = 256;          // frames per tile = threads (E phase: thread owns a frame)
constexpr int GEN_KC = 4;            // classes per E-phase sweep over the pairs
constexpr int GEN_KMAX = 19;

__host__ __device__ inline size_t gen_smem_bytes(int D, int K, size_t elem) {
    const size_t tile = (size_t)D * (GEN_TT + 1) * elem;                      // y tile [D][TT+1]
    const size_t wt = (size_t)K * GEN_TT * sizeof(double);                    // w tile [K][TT]
    const size_t jac = (size_t)3 * D * (D + 1) * sizeof(cd);                  // A, V, T of the Jacobi phase
    const size_t a = tile + wt;
    return (a > jac ? a : jac) + (size_t)(4 * 32 + 8 + 8 * GEN_KMAX + 64) * sizeof(double) + 20 * sizeof(JacobiRot) + 64;
}

// InT = float2 (complex64 observations, the normal hand-off between the blocks) or double2
// (complex128: the float64 hand-off of gss_enhance_c64_ex / gss_cacgmm_c128)
template <typename InT>
__global__ void __launch_bounds__(GEN_NT, sizeof(InT) == 8 ? 2 : 1) cacgmm_em_generic_kernel(const CacgmmGenericParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = p.D, K = p.K, Ts = p.T;
    const int NP = D * (D + 1) / 2;
    const int JLD = D + 1;
    const int YP = GEN_TT + 1;                                                 // tile row pitch (elements)
    const size_t tile_b = (size_t)D * YP * sizeof(InT), wt_b = (size_t)K * GEN_TT * sizeof(double);
    const size_t jac_b = (size_t)3 * D * JLD * sizeof(cd);
    InT* ytile = reinterpret_cast<InT*>(smem_raw);                            // [D][YP]
    double* wtile = reinterpret_cast<double*>(smem_raw + tile_b);             // [K][TT]
    unsigned char* sp = smem_raw + (tile_b + wt_b > jac_b ? tile_b + wt_b : jac_b);
    double* logdet_s = reinterpret_cast<double*>(sp);                         // [32]
    double* pi_s = logdet_s + 32;                                             // [32]
    double* tr_s = logdet_s + 64;                                             // [32]
    double* lam_s = logdet_s + 96;                                            // [32+] inverse floored eigenvalues (D <= 34)
    double* gred = logdet_s + 4 * 32 + 8;                                     // [8 warps][KMAX]
    double* jred = gred + 8 * GEN_KMAX;                                       // [64]
    JacobiRot* jrot = reinterpret_cast<JacobiRot*>(jred + 64);                // [20] (D <= 34: 17 rotations per round)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bf = blockIdx.x;
    const int b = bf / p.F, f = bf - b * p.F;
    const int T = p.Tper ? min(max(p.Tper[b], 0), Ts) : Ts;
    const InT* __restrict__ Yg = reinterpret_cast<const InT*>(p.Y) + (size_t)bf * D * Ts;
    if (T <= 0) {
        for (int i = tid; i < K * Ts; i += GEN_NT) p.posterior[(size_t)bf * K * Ts + i] = 0.f;
        return;
    }
    const uint8_t* __restrict__ act = p.activity + (size_t)b * K * p.T_act;
    cd* __restrict__ Bm = p.Bmat + (size_t)bf * K * NP;
    cd* __restrict__ Ph = p.Phi + (size_t)bf * K * NP;
    const int total_iters = p.iterations + (p.iterations_post - 1);

    for (int pass = 0; pass <= total_iters; ++pass) {
        const bool is_final = pass == total_iters;
        const bool guided = pass < p.iterations;
        const double eps = is_final ? 0.0 : p.eps;
        double gsum[GEN_KMAX];
        for (int k = 0; k < K; ++k) gsum[k] = 0.0;
        for (int e = tid; e < K * NP; e += GEN_NT) Ph[e] = cmake(0.0, 0.0);

        for (int s0 = 0; s0 < T; s0 += GEN_TT) {
            const int tn = min(GEN_TT, T - s0);
            for (int i = tid; i < D * GEN_TT; i += GEN_NT) {
                const int d = i / GEN_TT, t = i - d * GEN_TT;
                InT v; v.x = 0; v.y = 0;
                if (t < tn) v = __ldg(&Yg[(size_t)d * Ts + s0 + t]);
                ytile[d * YP + t] = v;
            }
            __syncthreads();
            // ---------------- E step: thread owns frame s0 + tid ----------------
            if (tid < tn) {
                const int t = s0 + tid;
                double n2 = 0.0;
                for (int d = 0; d < D; ++d) {
                    const InT v = ytile[d * YP + tid];
                    n2 = fma((double)v.x, (double)v.x, fma((double)v.y, (double)v.y, n2));
                }
                const double s = n2 > 0.0 ? 1.0 / n2 : 0.0;
                double g[GEN_KMAX], qn[GEN_KMAX];
                if (pass == 0) {
                    double tot = 0.0;
                    for (int k = 0; k < K; ++k) { g[k] = act[(size_t)k * p.T_act + t] ? 1.0 : 1e-10; tot += g[k]; qn[k] = 1.0; }
                    for (int k = 0; k < K; ++k) g[k] /= tot;                  // core.py:156-160; q == 1
                } else {
                    for (int k0 = 0; k0 < K; k0 += GEN_KC) {
                        double qa[GEN_KC];
#pragma unroll
                        for (int c = 0; c < GEN_KC; ++c) qa[c] = 0.0;
                        int pr = 0;
                        for (int d = 0; d < D; ++d) {
                            const InT yd = ytile[d * YP + tid];
                            const double dr = yd.x, di = yd.y;
                            for (int e = 0; e <= d; ++e, ++pr) {
                                const InT ye = ytile[e * YP + tid];
                                const double pre = fma(dr, (double)ye.x, di * (double)ye.y);
                                const double pim = fma(di, (double)ye.x, -(dr * (double)ye.y));
#pragma unroll
                                for (int c = 0; c < GEN_KC; ++c) {
                                    if (k0 + c < K) {
                                        const cd bk = Bm[(size_t)(k0 + c) * NP + pr];     // warp-uniform address
                                        qa[c] = fma(pre, bk.x, fma(pim, bk.y, qa[c]));
                                    }
                                }
                            }
                        }
#pragma unroll
                        for (int c = 0; c < GEN_KC; ++c)
                            if (k0 + c < K) qn[k0 + c] = fmax(fabs(qa[c]) * s, GSS_F64_TINY);
                    }
                    // log_pdf = -D log q - log det; masked softmax with the mixture weights
                    // (complex_angular_central_gaussian.py:166-203, mixture_model_utils.py:32-53)
                    double mx = -INFINITY;
                    for (int k = 0; k < K; ++k) { g[k] = -(double)D * log(qn[k]) - logdet_s[k]; mx = fmax(mx, g[k]); }
                    double den = 0.0;
                    for (int k = 0; k < K; ++k) {
                        double a = exp(g[k] - mx) * pi_s[k];
                        if (guided && !act[(size_t)k * p.T_act + t]) a = 0.0;
                        g[k] = a; den += a;
                    }
                    den = fmax(den, GSS_F64_TINY);
                    for (int k = 0; k < K; ++k) {
                        double v = g[k] / den;
                        if (eps != 0.0) v = fmin(fmax(v, eps), 1.0 - eps);
                        g[k] = v;
                    }
                }
                for (int k = 0; k < K; ++k) {
                    gsum[k] += g[k];
                    wtile[k * GEN_TT + tid] = g[k] * s / qn[k];
                    if (is_final) p.posterior[((size_t)bf * K + k) * Ts + t] = (float)g[k];
                }
            } else {
                for (int k = 0; k < K; ++k) wtile[k * GEN_TT + tid] = 0.0;
            }
            __syncthreads();
            // ---------------- M step: thread owns entries (k, d >= e) ----------------
            if (!is_final) {
                for (int en = tid; en < K * NP; en += GEN_NT) {
                    const int k = en / NP, pr = en - k * NP;
                    int d = 0;
                    while ((d + 1) * (d + 2) / 2 <= pr) ++d;
                    const int e = pr - d * (d + 1) / 2;
                    const InT* yd = ytile + d * YP;
                    const InT* ye = ytile + e * YP;
                    const double* wk = wtile + k * GEN_TT;
                    double ar = 0.0, ai = 0.0, br = 0.0, bi = 0.0;                        // two chains
                    int t = 0;
                    for (; t + 1 < tn; t += 2) {
                        const InT a0 = yd[t], c0 = ye[t], a1 = yd[t + 1], c1 = ye[t + 1];
                        const double w0 = wk[t], w1 = wk[t + 1];
                        ar = fma(w0, fma((double)a0.x, (double)c0.x, (double)a0.y * (double)c0.y), ar);
                        ai = fma(w0, fma((double)a0.y, (double)c0.x, -((double)a0.x * (double)c0.y)), ai);
                        br = fma(w1, fma((double)a1.x, (double)c1.x, (double)a1.y * (double)c1.y), br);
                        bi = fma(w1, fma((double)a1.y, (double)c1.x, -((double)a1.x * (double)c1.y)), bi);
                    }
                    if (t < tn) {
                        const InT a0 = yd[t], c0 = ye[t];
                        const double w0 = wk[t];
                        ar = fma(w0, fma((double)a0.x, (double)c0.x, (double)a0.y * (double)c0.y), ar);
                        ai = fma(w0, fma((double)a0.y, (double)c0.x, -((double)a0.x * (double)c0.y)), ai);
                    }
                    cd v = Ph[en];
                    v.x += ar + br; v.y += ai + bi;
                    Ph[en] = v;
                }
            }
            __syncthreads();
        }
        if (is_final) {
            for (int i = tid; i < K * (Ts - T); i += GEN_NT) {
                const int k = i / (Ts - T), t = T + i - k * (Ts - T);
                p.posterior[((size_t)bf * K + k) * Ts + t] = 0.f;
            }
            break;
        }
        // ---- mixture weights pi_k = mean_t gamma_kt (mixture_model_utils.py:187), fixed summation order ----
        for (int k = 0; k < K; ++k) {
            const double v = warp_sum(gsum[k]);
            if (lane == 0) gred[warp * GEN_KMAX + k] = v;
        }
        __syncthreads();
        if (tid < K) {
            double v = 0.0;
            for (int w = 0; w < GEN_NT / 32; ++w) v += gred[w * GEN_KMAX + tid];
            pi_s[tid] = v / (double)T;
            double tr = 0.0;
            for (int d = 0; d < D; ++d) tr += Ph[(size_t)tid * NP + tri(d, d)].x;
            tr_s[tid] = tr;
        }
        __syncthreads();
        // ---- class matrices: eigh -> normalise by the largest eigenvalue -> floor -> V diag(1/lambda) V^H ----
        for (int k = 0; k < K; ++k) {
            cd* A = reinterpret_cast<cd*>(smem_raw);
            cd* V = A + D * JLD;
            cd* Tm = V + D * JLD;
            const cd* src = Ph + (size_t)k * NP;
            const double tr = tr_s[k];
            const double itr = (tr > 0.0 && isfinite(tr)) ? 1.0 / tr : 1.0;     // scale only: eigenvalues are re-normalised below
            for (int i = tid; i < D * D; i += GEN_NT) {
                const int r = i / D, c = i - r * D;
                cd v;
                if (r == c) v = cmake(src[tri(r, r)].x * itr, 0.0);              // force_hermitian (utils.py:323-334)
                else if (r > c) v = cscale(src[tri(r, c)], itr);
                else v = cscale(cconj(src[tri(c, r)]), itr);
                A[r * JLD + c] = v;
            }
            __syncthreads();
            cd* Vg = p.Vwarm + ((size_t)bf * K + k) * D * JLD;
            const bool warm = pass > 0;
            if (warm) {
                for (int i = tid; i < D * JLD; i += GEN_NT) V[i] = Vg[i];
                __syncthreads();
                for (int i = tid; i < D * D; i += GEN_NT) {                      // T = A V
                    const int r = i / D, c = i - r * D;
                    cd acc = cmake(0.0, 0.0);
                    for (int j = 0; j < D; ++j) cfma(acc, A[r * JLD + j], V[j * JLD + c]);
                    Tm[r * JLD + c] = acc;
                }
                __syncthreads();
                for (int i = tid; i < D * D; i += GEN_NT) {                      // A' = V^H T
                    const int r = i / D, c = i - r * D;
                    if (c > r) continue;
                    cd acc = cmake(0.0, 0.0);
                    for (int j = 0; j < D; ++j) cfma(acc, cconj(V[j * JLD + r]), Tm[j * JLD + c]);
                    if (r == c) acc.y = 0.0;
                    A[r * JLD + c] = acc;
                    if (r != c) A[c * JLD + r] = cconj(acc);
                }
                __syncthreads();
            }
            const int sweeps = block_jacobi_eigh(A, V, D, JLD, jrot, jred, tid, GEN_NT, !warm);
            if (sweeps < 0 && tid == 0 && p.info) atomicMax(&p.info[b], GSS_INFO_NO_CONVERGE | (f << 8));
            for (int i = tid; i < D * JLD; i += GEN_NT) Vg[i] = V[i];
            __syncthreads();
            if (tid == 0) {
                double mxl = -INFINITY;
                for (int i = 0; i < D; ++i) mxl = fmax(mxl, A[i * JLD + i].x);
                const double den = fmax(mxl, GSS_F64_TINY);
                double ld = 0.0;
                for (int i = 0; i < D; ++i) {
                    const double l = fmax(A[i * JLD + i].x / den, p.floor_);
                    ld += log(l);
                    lam_s[i] = 1.0 / l;
                }
                logdet_s[k] = ld;
            }
            __syncthreads();
            for (int pr = tid; pr < NP; pr += GEN_NT) {
                int d = 0;
                while ((d + 1) * (d + 2) / 2 <= pr) ++d;
                const int e = pr - d * (d + 1) / 2;
                cd sacc = cmake(0.0, 0.0);
                for (int j = 0; j < D; ++j) {
                    const cd t1 = cscale(V[d * JLD + j], lam_s[j]);
                    cfmac(sacc, t1, V[e * JLD + j]);
                }
                Bm[(size_t)k * NP + pr] = (d == e) ? cmake(sacc.x, 0.0) : cmake(2.0 * sacc.x, 2.0 * sacc.y);
            }
            __syncthreads();
        }
        if (pass == total_iters - 1) {
            if (p.weight_out && tid < K) p.weight_out[(size_t)bf * K + tid] = pi_s[tid];
            if (p.logdet_out && tid < K) p.logdet_out[(size_t)bf * K + tid] = logdet_s[tid];
            if (p.cov_out) {
                cd* out = reinterpret_cast<cd*>(p.cov_out) + (size_t)bf * K * D * D;
                for (int i = tid; i < K * D * D; i += GEN_NT) {
                    const int k = i / (D * D), rc = i - k * D * D, r = rc / D, c = rc - r * D;
                    const double tr = tr_s[k];
                    const double itr = (tr > 0.0 && isfinite(tr)) ? 1.0 / tr : 1.0;
                    const cd* src = Ph + (size_t)k * NP;
                    cd v;
                    if (r == c) v = cmake(src[tri(r, r)].x * itr, 0.0);
                    else if (r > c) v = cscale(src[tri(r, c)], itr);
                    else v = cscale(cconj(src[tri(c, r)]), itr);
                    out[i] = v;
                }
            }
        }
        __syncthreads();
    }
}

size_t cacgmm_generic_ws_bytes(int B, int F, int D, int K) {
    const size_t NP = (size_t)D * (D + 1) / 2, BF = (size_t)B * F;
    return 2 * align_up(BF * K * NP * sizeof(cd)) + align_up(BF * K * D * (D + 1) * sizeof(cd));
}

int cacgmm_generic_launch(const void* Y, int y_is_c128, const uint8_t* activity, const int* Tper, float* posterior,
                          double* weight_out, double* logdet_out, double* cov_out, int* info,
                          int B, int F, int D, int T, int K, int T_act, int iterations, int iterations_post,
                          double eps, double floor_, void* ws, size_t ws_bytes, cudaStream_t st) {
    GSS_REQUIRE(D <= 34 && K <= GEN_KMAX, GSS_ERR_UNSUPPORTED, "gss_cacgmm_c64 (generic): D=%d K=%d", D, K);
    const size_t need = cacgmm_generic_ws_bytes(B, F, D, K);
    GSS_REQUIRE(ws && ws_bytes >= need, GSS_ERR_WORKSPACE,
                "gss_cacgmm_c64: D=%d, K=%d run on the generic kernel, which needs %zu bytes of workspace (got %zu)",
                D, K, need, ws_bytes);
    Arena a(ws, ws_bytes);
    const size_t NP = (size_t)D * (D + 1) / 2, BF = (size_t)B * F;
    CacgmmGenericParams p;
    p.Y = Y; p.activity = activity; p.Tper = Tper; p.posterior = posterior;
    p.weight_out = weight_out; p.logdet_out = logdet_out; p.cov_out = cov_out; p.info = info;
    p.Bmat = a.take<cd>(BF * K * NP);
    p.Phi = a.take<cd>(BF * K * NP);
    p.Vwarm = a.take<cd>(BF * K * D * (D + 1));
    p.B = B; p.F = F; p.D = D; p.K = K; p.T = T; p.T_act = T_act;
    p.iterations = iterations; p.iterations_post = iterations_post; p.eps = eps; p.floor_ = floor_;
    if (y_is_c128) {
        const size_t smem = gen_smem_bytes(D, K, sizeof(double2));
        GSS_REQUIRE(smem <= 227 * 1024, GSS_ERR_UNSUPPORTED, "gss_cacgmm_c128: D=%d K=%d need %zu bytes of shared memory", D, K, smem);
        GSS_CUDA(cudaFuncSetAttribute(cacgmm_em_generic_kernel<double2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cacgmm_em_generic_kernel<double2><<<B * F, GEN_NT, smem, st>>>(p);
    } else {
        const size_t smem = gen_smem_bytes(D, K, sizeof(float2));
        GSS_CUDA(cudaFuncSetAttribute(cacgmm_em_generic_kernel<float2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cacgmm_em_generic_kernel<float2><<<B * F, GEN_NT, smem, st>>>(p);
    }
    GSS_LAUNCH_CHECK("cacgmm_em_generic_kernel");
    return GSS_OK;
}

}  // namespace gss
