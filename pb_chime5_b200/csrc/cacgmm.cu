// Guided CACGMM EM, one CTA per (utterance, frequency bin), all EM iterations
// fused in one launch.  Restates (for the device) GSS.__call__ core.py:154-214,
// CACGMMTrainer.fit cacgmm.py:141-278, _m_step :313-343, cACG._fit
// complex_angular_central_gaussian.py:253-310, from_covariance :81-131,
// _log_pdf :166-203, log_pdf_to_affiliation mixture_model_utils.py:7-55,
// estimate_mixture_weight :187-190, CACGMM.predict cacgmm.py:63-94.
//
// Arithmetic: complex64 observations (exact), everything else float64.
// Formulation (see DESIGN.md "CACGMM kernel"):
//   P_t      = y_t y_t^H  (raw, un-normalised frame; products of float32 are
//              exact in float64), frame norm folded into a scalar s_t = 1/|y_t|^2
//   E-phase  q_kt = s_t * sum_{d>=e} B'_k[d,e] . P_t[d,e]      (thread owns frame)
//   M-phase  Phi_k = sum_t w_kt P_t,  w_kt = gamma_kt s_t / q_kt (lane owns 2x2 block)
//   matrix   B_k = Phi_k^{-1} via Cholesky when no eigenvalue can be floored
//            (trace bound), otherwise Jacobi eigh with the reference's
//            normalise-by-max + floor semantics.
#include "common.cuh"
#include "smallmat.cuh"

#ifndef GSS_DP_LIST
#define GSS_DP_LIST GSS_CASE(2) GSS_CASE(4) GSS_CASE(6) GSS_CASE(8) GSS_CASE(12) GSS_CASE(16) GSS_CASE(24)
#endif

namespace gss {

struct CacgmmParams {
    const float2* Y;          // (B,F,D,T)
    const uint8_t* activity;  // (B,K,T_act)
    float* posterior;         // (B,F,K,T)
    double* weight_out;       // (B,F,K) or null
    double* logdet_out;       // (B,F,K) or null
    double* cov_out;          // (B,F,K,D,D) complex128 or null
    int* info;                // (B) or null
    int* slow_count;          // (1) or null : number of slow-path (Jacobi) class updates
    int B, F, D, T, T_act;
    int iterations, iterations_post;
    double eps, floor_;
};

constexpr size_t cmax(size_t a, size_t b) { return a > b ? a : b; }

// E-phase sub-blocking of the packed lower triangle: SB x SB blocks so that only
// 2*SB complex values of the frame are live in registers at a time.
__host__ __device__ constexpr int sub_block(int DP) {
    return DP < 12 ? DP : (DP % 6 == 0 ? 6 : 4);
}
__host__ __device__ constexpr int sb_pairs(int SB, int I, int J) { return I == J ? SB * (SB + 1) / 2 : SB * SB; }
// Which of the ES threads of a frame handles sub-block idx.  Part 0 also runs the
// softmax, so it gets a smaller share (~40 % for ES = 2).
__host__ __device__ constexpr int sb_part(int DP, int ES, int idx) {
    if (ES == 1) return 0;
    const int SB = sub_block(DP), NS = DP / SB, NPAIR = DP * (DP + 1) / 2;
    const double share0 = 0.8 / (0.8 + 1.2 * (ES - 1));
    int cum = 0, i = 0;
    for (int I = 0; I < NS; ++I)
        for (int J = 0; J <= I; ++J, ++i) {
            if (i == idx) {
                const double mid = cum + 0.5 * sb_pairs(SB, I, J);
                if (mid < share0 * NPAIR) return 0;
                const double rest = (mid - share0 * NPAIR) / ((1.0 - share0) * NPAIR);
                int h = 1 + int(rest * (ES - 1));
                return h > ES - 1 ? ES - 1 : h;
            }
            cum += sb_pairs(SB, I, J);
        }
    return 0;
}

template <int DP, int K, int NT, int ES>
struct CacgmmCfg {
    static constexpr int NP = DP * (DP + 1) / 2;          // packed Hermitian pairs
    static constexpr int NB = DP / 2;                     // 2x2 block rows
    static constexpr int G = NB * (NB + 1) / 2;           // lanes per M-phase group
    static constexpr int NG = NT / G;                     // M-phase groups
    static constexpr int YLD = DP + 1;                    // padded smem row (elements)
    static constexpr int KP = (K + 1) & ~1;               // padded class count (16 B rows)
    static constexpr int NW = NT / 32;
    static constexpr int JLD = DP + 1;                    // leading dim of Jacobi matrices
    static constexpr int TE = NT / ES;                    // frames per E step (complex64 tile)
    static constexpr int TM = TE / 2;                     // frames per M step (complex128 tile)
    static constexpr int ST = 4 * TE;                     // frames per super tile (w buffer)
    static constexpr int SB = sub_block(DP);
    static constexpr int NS = DP / SB;
    static constexpr int NSB = NS * (NS + 1) / 2;
    static constexpr size_t Y_BYTES = size_t(TE) * YLD * sizeof(float2);   // == TM * YLD * sizeof(cd)
    static constexpr size_t CHOL_BYTES = size_t(K) * NP * sizeof(cd);
    static constexpr size_t JAC_BYTES = size_t(2) * DP * JLD * sizeof(cd);
    static constexpr size_t YS_BYTES = cmax(Y_BYTES, cmax(CHOL_BYTES, JAC_BYTES));
    static constexpr size_t W_BYTES = size_t(ST) * KP * sizeof(double);
    static constexpr size_t B_BYTES = size_t(NP) * K * sizeof(cd);
    static constexpr size_t ACC_BYTES = size_t(K) * NP * sizeof(cd);
    static constexpr size_t QP_BYTES = (ES > 1) ? size_t(ES - 1) * K * TE * sizeof(double) : 0;
    static constexpr int MISC_DOUBLES = 4 * 32 + NW * K + 64;
    static constexpr size_t MISC_BYTES = size_t(MISC_DOUBLES) * 8 + 16 * sizeof(JacobiRot) + 64 * sizeof(int);
    static constexpr size_t SMEM = YS_BYTES + W_BYTES + B_BYTES + ACC_BYTES + QP_BYTES + MISC_BYTES;
    static_assert(DP % 2 == 0 && DP <= 32, "padded channel count must be even and <= 32");
    static_assert(DP % SB == 0, "sub-block must divide DP");
    static_assert(NG >= 1, "block too small for the M-phase mapping");
    static_assert(NT % (32 * ES) == 0, "E parts must be warp aligned");
    static_assert(K < 20, "cacgmm.py:247");
};

// q_k += sum over the pairs of sub-block (I,J) of B'_k[d,e] . P[d,e]
template <int DP, int K, int SB, int I, int J>
__device__ __forceinline__ void quad_subblock(const float2* __restrict__ yrow, const cd* __restrict__ Bsm,
                                              double (&qa)[K], double (&qb)[K]) {
    double rr[SB], ri[SB], cr[SB], ci[SB];
#pragma unroll
    for (int a = 0; a < SB; ++a) { const float2 v = yrow[I * SB + a]; rr[a] = (double)v.x; ri[a] = (double)v.y; }
#pragma unroll
    for (int c = 0; c < SB; ++c) {
        if (I == J) { cr[c] = rr[c]; ci[c] = ri[c]; }
        else { const float2 v = yrow[J * SB + c]; cr[c] = (double)v.x; ci[c] = (double)v.y; }
    }
#pragma unroll
    for (int a = 0; a < SB; ++a) {
#pragma unroll
        for (int c = 0; c < SB; ++c) {
            if (I == J && c > a) continue;
            const double pre = fma(rr[a], cr[c], ri[a] * ci[c]);
            const double pim = fma(ri[a], cr[c], -(rr[a] * ci[c]));
            const cd* b = Bsm + tri(I * SB + a, J * SB + c) * K;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const cd bk = b[k];
                qa[k] = fma(pre, bk.x, qa[k]);
                qb[k] = fma(pim, bk.y, qb[k]);
            }
        }
    }
}

template <int DP, int K, int ES, int IDX, int NSB>
struct QuadAll {
    static __device__ __forceinline__ void run(int h, const float2* yrow, const cd* Bsm,
                                               double (&qa)[K], double (&qb)[K]) {
        constexpr int SB = sub_block(DP);
        // idx -> (I, J), row-major lower triangle
        constexpr int I = [] { int r = 0; while ((r + 1) * (r + 2) / 2 <= IDX) ++r; return r; }();
        constexpr int J = IDX - I * (I + 1) / 2;
        if (h == sb_part(DP, ES, IDX)) quad_subblock<DP, K, SB, I, J>(yrow, Bsm, qa, qb);
        QuadAll<DP, K, ES, IDX + 1, NSB>::run(h, yrow, Bsm, qa, qb);
    }
};
template <int DP, int K, int ES, int NSB>
struct QuadAll<DP, K, ES, NSB, NSB> {
    static __device__ __forceinline__ void run(int, const float2*, const cd*, double (&)[K], double (&)[K]) {}
};

template <int DP, int K, int NT, int ES, int MINB>
__global__ void __launch_bounds__(NT, MINB) cacgmm_em_kernel(const CacgmmParams p) {
    using C = CacgmmCfg<DP, K, NT, ES>;
    constexpr int TE = C::TE, TM = C::TM, ST = C::ST;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char* sp = smem_raw;
    unsigned char* ys_raw = sp;                     sp += C::YS_BYTES;   // frame tile / matrix scratch
    double* wsm = reinterpret_cast<double*>(sp);    sp += C::W_BYTES;    // [ST][KP]
    cd* Bsm = reinterpret_cast<cd*>(sp);            sp += C::B_BYTES;    // [NP][K]
    cd* acc_sm = reinterpret_cast<cd*>(sp);         sp += C::ACC_BYTES;  // [K][NP]
    double* qpart = reinterpret_cast<double*>(sp);  sp += C::QP_BYTES;   // [ES-1][K][TE]
    double* logdet_s = reinterpret_cast<double*>(sp);                    // [32]
    double* pi_s = logdet_s + 32;                                        // [32]
    double* tr_s = logdet_s + 64;                                        // [32]
    double* gred = logdet_s + 128;                                       // [NW][K]
    double* jred = gred + C::NW * K;                                     // [64]
    JacobiRot* jrot = reinterpret_cast<JacobiRot*>(jred + 64);           // [16]
    int* flags_s = reinterpret_cast<int*>(jrot + 16);                    // [K] slow-path flags
    float2* yf = reinterpret_cast<float2*>(ys_raw);                      // E tile [TE][YLD] complex64
    cd* yd = reinterpret_cast<cd*>(ys_raw);                              // M tile [TM][YLD] complex128

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int bf = blockIdx.x;
    const int b = bf / p.F, f = bf - b * p.F;
    const int D = p.D, T = p.T;
    const float2* __restrict__ Yg = p.Y + (size_t)bf * D * T;
    const uint8_t* __restrict__ act = p.activity + (size_t)b * K * p.T_act;

    // E-phase role: thread (frame e_t, part e_h); e_h is warp-uniform
    const int e_h = tid / TE;
    const int e_t = tid - e_h * TE;
    // M-phase role: group m_g, lane m_l -> 2x2 block (bi >= bj)
    const int m_g = tid / C::G;
    const int m_l = tid - m_g * C::G;
    const bool m_active = m_g < C::NG;
    int bi = 0;
    while ((bi + 1) * (bi + 2) / 2 <= m_l) ++bi;
    const int bj = m_l - bi * (bi + 1) / 2;
    const int r0 = 2 * bi, c0 = 2 * bj;

    const int total_iters = p.iterations + (p.iterations_post - 1);

    for (int i = tid; i < C::NP * K; i += NT) Bsm[i] = cmake(0.0, 0.0);
    __syncthreads();

    // pass i < total_iters : (E-step if i > 0) + M-step.   pass == total_iters :
    // final predict = E-step only, unguided, unclipped (cacgmm.py:63-70).
    for (int pass = 0; pass <= total_iters; ++pass) {
        const bool is_final = pass == total_iters;
        const bool guided = pass < p.iterations;
        const double eps = is_final ? 0.0 : p.eps;
        double gsum[K];
#pragma unroll
        for (int k = 0; k < K; ++k) gsum[k] = 0.0;

        for (int s0 = 0; s0 < T; s0 += ST) {
            const int s1 = min(s0 + ST, T);
            // ================= E sweep over the super tile =================
            for (int t0 = s0; t0 < s1; t0 += TE) {
                for (int i = tid; i < DP * TE; i += NT) {
                    const int d = i / TE, t = i - d * TE;
                    float2 v = make_float2(0.f, 0.f);
                    if (d < D && t0 + t < T) v = __ldg(&Yg[(size_t)d * T + t0 + t]);
                    yf[t * C::YLD + d] = v;
                }
                __syncthreads();
                const float2* yrow = yf + e_t * C::YLD;
                double qa[K], qb[K];
#pragma unroll
                for (int k = 0; k < K; ++k) { qa[k] = 0.0; qb[k] = 0.0; }
                if (pass > 0) {
                    QuadAll<DP, K, ES, 0, C::NSB>::run(e_h, yrow, Bsm, qa, qb);
                    if (ES > 1 && e_h > 0) {
#pragma unroll
                        for (int k = 0; k < K; ++k) qpart[((e_h - 1) * K + k) * TE + e_t] = qa[k] + qb[k];
                    }
                }
                if (ES > 1) __syncthreads();
                if (e_h == 0) {
                    const int t = t0 + e_t;
                    if (t < s1) {
                        double n2 = 0.0;
#pragma unroll
                        for (int d = 0; d < DP; ++d) {
                            const float2 v = yrow[d];
                            n2 = fma((double)v.x, (double)v.x, fma((double)v.y, (double)v.y, n2));
                        }
                        const double s = n2 > 0.0 ? 1.0 / n2 : 0.0;
                        double g[K], w[K];
                        if (pass == 0) {
                            // initialisation from the activity (core.py:156-160); q == 1
                            double tot = 0.0;
#pragma unroll
                            for (int k = 0; k < K; ++k) { g[k] = act[(size_t)k * p.T_act + t] ? 1.0 : 1e-10; tot += g[k]; }
#pragma unroll
                            for (int k = 0; k < K; ++k) { g[k] /= tot; w[k] = g[k] * s; }
                        } else {
                            double lp[K], qn[K];
                            double mx = -INFINITY;
#pragma unroll
                            for (int k = 0; k < K; ++k) {
                                double q = qa[k] + qb[k];
                                if (ES > 1) {
#pragma unroll
                                    for (int h = 1; h < ES; ++h) q += qpart[((h - 1) * K + k) * TE + e_t];
                                }
                                q = fmax(fabs(q) * s, GSS_F64_TINY);
                                qn[k] = q;
                                lp[k] = -(double)D * log(q) - logdet_s[k];
                                mx = fmax(mx, lp[k]);
                            }
                            double den = 0.0;
#pragma unroll
                            for (int k = 0; k < K; ++k) {
                                double a = exp(lp[k] - mx) * pi_s[k];
                                if (guided && !act[(size_t)k * p.T_act + t]) a = 0.0;
                                g[k] = a; den += a;
                            }
                            den = fmax(den, GSS_F64_TINY);
#pragma unroll
                            for (int k = 0; k < K; ++k) {
                                double v = g[k] / den;
                                if (eps != 0.0) v = fmin(fmax(v, eps), 1.0 - eps);
                                g[k] = v;
                                w[k] = v * s / qn[k];
                            }
                        }
                        if (is_final) {
#pragma unroll
                            for (int k = 0; k < K; ++k)
                                p.posterior[((size_t)bf * K + k) * T + t] = (float)g[k];
                        }
#pragma unroll
                        for (int k = 0; k < K; ++k) { gsum[k] += g[k]; wsm[(t - s0) * C::KP + k] = w[k]; }
                    }
                }
                __syncthreads();
            }
            if (is_final) continue;

            // ================= M sweep over the super tile =================
            cd macc[4][K];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int k = 0; k < K; ++k) macc[a][k] = cmake(0.0, 0.0);
            for (int t0 = s0; t0 < s1; t0 += TM) {
                for (int i = tid; i < DP * TM; i += NT) {
                    const int d = i / TM, t = i - d * TM;
                    float2 v = make_float2(0.f, 0.f);
                    if (d < D && t0 + t < T) v = __ldg(&Yg[(size_t)d * T + t0 + t]);
                    yd[t * C::YLD + d] = cmake((double)v.x, (double)v.y);
                }
                __syncthreads();
                if (m_active) {
                    const int tn = min(TM, s1 - t0);
                    for (int t = m_g; t < tn; t += C::NG) {
                        const cd* yrow = yd + t * C::YLD;
                        const cd a0 = yrow[r0], a1 = yrow[r0 + 1], b0 = yrow[c0], b1 = yrow[c0 + 1];
                        cd P[4];
                        P[0] = cmulc(a0, b0); P[1] = cmulc(a0, b1);
                        P[2] = cmulc(a1, b0); P[3] = cmulc(a1, b1);
                        const double* wr = wsm + (t0 - s0 + t) * C::KP;
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            const double wk = wr[k];
#pragma unroll
                            for (int a = 0; a < 4; ++a) {
                                macc[a][k].x = fma(wk, P[a].x, macc[a][k].x);
                                macc[a][k].y = fma(wk, P[a].y, macc[a][k].y);
                            }
                        }
                    }
                }
                __syncthreads();
            }
            // ---- flush: reduce over the M-phase groups in fixed order (deterministic) ----
            for (int g = 0; g < C::NG; ++g) {
                if (m_active && m_g == g) {
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const int d = r0 + (a >> 1), e = c0 + (a & 1);
                        if (e <= d) {
#pragma unroll
                            for (int k = 0; k < K; ++k) {
                                cd* dst = &acc_sm[k * C::NP + tri(d, e)];
                                if (g == 0 && s0 == 0) *dst = macc[a][k];
                                else { cd v = *dst; v.x += macc[a][k].x; v.y += macc[a][k].y; *dst = v; }
                            }
                        }
                    }
                }
                __syncthreads();
            }
        }
        if (is_final) break;

        // ---- mixture weights  pi_k = mean_t gamma_kt  (mixture_model_utils.py:187) ----
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double v = warp_sum(gsum[k]);
            if (lane == 0) gred[warp * K + k] = v;
        }
        __syncthreads();
        if (tid < K) {
            double v = 0.0;
            for (int w = 0; w < C::NW; ++w) v += gred[w * K + tid];
            pi_s[tid] = v / (double)T;
        }

        // ---- class matrices: fast path = Cholesky inverse, one warp per class ----
        cd* chol = reinterpret_cast<cd*>(ys_raw);      // [K][NP] scaled copy -> L -> L^{-1}
        const int NPD = D * (D + 1) / 2;
        for (int k = warp; k < K; k += C::NW) {
            const cd* src = acc_sm + k * C::NP;
            double tr = 0.0;
            if (lane < D) tr = src[tri(lane, lane)].x;
            tr = warp_sum(tr);
            bool slow = !(tr > 0.0) || !isfinite(tr);
            cd* L = chol + k * C::NP;
            if (!slow) {
                const double itr = 1.0 / tr;
                for (int i = lane; i < NPD; i += 32) L[i] = cscale(src[i], itr);
                __syncwarp();
                if (lane < D) L[tri(lane, lane)].y = 0.0;   // force_hermitian (utils.py:323-334)
                __syncwarp();
                slow = !warp_cholesky_packed(L, D, lane);
            }
            double ld = 0.0, trb = 0.0;
            if (!slow) {
                if (lane < D) ld = 2.0 * log(L[tri(lane, lane)].x);
                ld = warp_sum(ld);
                warp_tri_inverse_inplace(L, D, lane);
                // B = M^H M, row `lane`
                if (lane < D) {
                    const int d = lane;
                    for (int e = 0; e <= d; ++e) {
                        cd v = mhm_entry(L, D, d, e);
                        if (e == d) { trb = v.x; Bsm[tri(d, e) * K + k] = cmake(v.x, 0.0); }
                        else Bsm[tri(d, e) * K + k] = cmake(2.0 * v.x, 2.0 * v.y);
                    }
                }
                trb = warp_sum(trb);
                // lambda_min >= 1/tr(B), lambda_max <= tr(Phi)=1  => no eigenvalue is floored
                // if 1/tr(B) >= floor  (complex_angular_central_gaussian.py:118-121)
                if (!(trb * p.floor_ <= 1.0) || !isfinite(trb) || !isfinite(ld)) slow = true;
            }
            if (lane == 0) { flags_s[k] = slow ? 1 : 0; logdet_s[k] = ld; tr_s[k] = tr; }
        }
        __syncthreads();

        // ---- slow path: Jacobi eigh with normalise-by-max + floor, whole CTA per class ----
        for (int k = 0; k < K; ++k) {
            if (!flags_s[k]) continue;           // uniform (shared memory)
            cd* A = reinterpret_cast<cd*>(ys_raw);
            cd* V = A + DP * C::JLD;
            const cd* src = acc_sm + k * C::NP;
            const double tr = tr_s[k];
            const double itr = (tr > 0.0 && isfinite(tr)) ? 1.0 / tr : 1.0;
            for (int i = tid; i < D * D; i += NT) {
                const int r = i / D, c = i - r * D;
                cd v;
                if (r == c) v = cmake(src[tri(r, r)].x * itr, 0.0);
                else if (r > c) v = cscale(src[tri(r, c)], itr);
                else v = cscale(cconj(src[tri(c, r)]), itr);
                A[r * C::JLD + c] = v;
            }
            __syncthreads();
            const int sweeps = block_jacobi_eigh(A, V, D, C::JLD, jrot, jred, tid, NT);
            if (sweeps < 0 && tid == 0 && p.info) atomicMax(&p.info[b], GSS_INFO_NO_CONVERGE | (f << 8));
            if (tid == 0 && p.slow_count) atomicAdd(p.slow_count, 1);
            double* lam = jred;                  // [D] inverse floored eigenvalues
            __syncthreads();
            if (tid == 0) {
                double mxl = -INFINITY;
                for (int i = 0; i < D; ++i) mxl = fmax(mxl, A[i * C::JLD + i].x);
                const double den = fmax(mxl, GSS_F64_TINY);
                double ld = 0.0;
                for (int i = 0; i < D; ++i) {
                    double l = fmax(A[i * C::JLD + i].x / den, p.floor_);
                    ld += log(l);
                    lam[i] = 1.0 / l;
                }
                logdet_s[k] = ld;
            }
            __syncthreads();
            for (int i = tid; i < NPD; i += NT) {
                int d = 0;
                while ((d + 1) * (d + 2) / 2 <= i) ++d;
                const int e = i - d * (d + 1) / 2;
                cd s = cmake(0.0, 0.0);
                for (int j = 0; j < D; ++j) {
                    cd t1 = cscale(V[d * C::JLD + j], lam[j]);
                    cfmac(s, t1, V[e * C::JLD + j]);
                }
                Bsm[tri(d, e) * K + k] = (d == e) ? cmake(s.x, 0.0) : cmake(2.0 * s.x, 2.0 * s.y);
            }
            __syncthreads();
        }

        // ---- optional model outputs after the last M-step ----
        if (pass == total_iters - 1) {
            if (p.weight_out && tid < K) p.weight_out[(size_t)bf * K + tid] = pi_s[tid];
            if (p.logdet_out && tid < K) p.logdet_out[(size_t)bf * K + tid] = logdet_s[tid];
            if (p.cov_out) {
                cd* out = reinterpret_cast<cd*>(p.cov_out) + (size_t)bf * K * D * D;
                for (int i = tid; i < K * D * D; i += NT) {
                    const int k = i / (D * D), rc = i - k * D * D, r = rc / D, c = rc - r * D;
                    const double tr = tr_s[k];
                    const double itr = (tr > 0.0 && isfinite(tr)) ? 1.0 / tr : 1.0;
                    const cd* src = acc_sm + k * C::NP;
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

// ---------------------------------------------------------------------------
// Host dispatch
// ---------------------------------------------------------------------------
template <int DP, int K, int NT, int ES, int MINB>
static int launch_cacgmm(const CacgmmParams& p, cudaStream_t st) {
    using C = CacgmmCfg<DP, K, NT, ES>;
    auto kern = cacgmm_em_kernel<DP, K, NT, ES, MINB>;
    GSS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    kern<<<p.B * p.F, NT, C::SMEM, st>>>(p);
    GSS_LAUNCH_CHECK("cacgmm_em_kernel");
    return GSS_OK;
}

template <int DP, int K>
static int launch_cacgmm_dk(const CacgmmParams& p, cudaStream_t st) {
    if constexpr (DP >= 12) return launch_cacgmm<DP, K, 256, 2, 2>(p, st);
    else                    return launch_cacgmm<DP, K, 256, 1, 2>(p, st);
}

template <int DP>
static int launch_cacgmm_d(const CacgmmParams& p, int K, cudaStream_t st) {
    switch (K) {
        case 2: return launch_cacgmm_dk<DP, 2>(p, st);
        case 3: return launch_cacgmm_dk<DP, 3>(p, st);
        case 4: return launch_cacgmm_dk<DP, 4>(p, st);
        case 5: return launch_cacgmm_dk<DP, 5>(p, st);
        case 6: return launch_cacgmm_dk<DP, 6>(p, st);
        default: return fail(GSS_ERR_UNSUPPORTED, "gss_cacgmm_c64: K=%d not built (built: 2..6)", K);
    }
}

int cacgmm_dispatch(const CacgmmParams& p, int K, cudaStream_t st) {
    const int DP = (p.D + 1) & ~1;
    switch (DP) {
#define GSS_CASE(dp) case dp: return launch_cacgmm_d<dp>(p, K, st);
        GSS_DP_LIST
#undef GSS_CASE
        default: return fail(GSS_ERR_UNSUPPORTED, "gss_cacgmm_c64: D=%d not built", p.D);
    }
}

}  // namespace gss

extern "C" int gss_cacgmm_c64(const gss_c64* Y, const uint8_t* activity, float* posterior,
                              int iterations, int iterations_post,
                              double affiliation_eps, double eigenvalue_floor,
                              int B, int F, int D, int T, int K, int T_act,
                              double* weight_out, double* logdet_out, double* covariance_out,
                              int* info, void* ws, size_t ws_bytes, void* stream) {
    using namespace gss;
    (void)ws; (void)ws_bytes;
    GSS_REQUIRE(Y && activity && posterior, GSS_ERR_ARG, "gss_cacgmm_c64: null pointer");
    GSS_REQUIRE(B >= 0 && F >= 0 && T > 0, GSS_ERR_ARG, "gss_cacgmm_c64: bad dims B=%d F=%d T=%d", B, F, T);
    GSS_REQUIRE(D > 1, GSS_ERR_ARG, "gss_cacgmm_c64: D=%d, need D > 1 (cacgmm.py:196)", D);
    GSS_REQUIRE(D < 35, GSS_ERR_ARG, "Channels: %d, sure? (cacgmm.py:248)", D);
    GSS_REQUIRE(K > 1, GSS_ERR_ARG, "num_classes: %d, need > 1 (cacgmm.py:212)", K);
    GSS_REQUIRE(K < 20, GSS_ERR_ARG, "num_classes: %d, sure? (cacgmm.py:247)", K);
    GSS_REQUIRE(iterations > 0, GSS_ERR_ARG, "iterations=%d must be > 0 (cacgmm.py:199)", iterations);
    GSS_REQUIRE(iterations_post >= 1, GSS_ERR_UNSUPPORTED,
                "iterations_post=%d: the reference raises TypeError for 0 (core.py:198-202)", iterations_post);
    GSS_REQUIRE(T_act >= T, GSS_ERR_ARG, "activity has %d frames, observation %d (cacgmm.py:216-218)", T_act, T);
    GSS_REQUIRE(D <= 32, GSS_ERR_UNSUPPORTED, "gss_cacgmm_c64: D=%d > 32 not built", D);
    if (B == 0 || F == 0) return GSS_OK;
    CacgmmParams p;
    p.Y = (const float2*)Y; p.activity = activity; p.posterior = posterior;
    p.weight_out = weight_out; p.logdet_out = logdet_out; p.cov_out = covariance_out;
    p.info = info; p.slow_count = nullptr;
    p.B = B; p.F = F; p.D = D; p.T = T; p.T_act = T_act;
    p.iterations = iterations; p.iterations_post = iterations_post;
    p.eps = affiliation_eps; p.floor_ = eigenvalue_floor;
    return cacgmm_dispatch(p, K, (cudaStream_t)stream);
}
