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

// The padded channel counts are instantiated in three translation units so that the build is
// parallel (cacgmm.cu = part 0 with the dispatch and the C entry point; cacgmm_part1.cu and
// cacgmm_part2.cu include this file with GSS_EM_PART set).  -DGSS_DP_LIST=... (developer builds)
// puts everything into part 0.
#ifndef GSS_EM_PART
#define GSS_EM_PART 0
#endif
#ifdef GSS_DP_LIST
#define GSS_DP_PART0 GSS_DP_LIST
#define GSS_DP_PART1
#define GSS_DP_PART2
#else
#define GSS_DP_PART0 GSS_CASE(2) GSS_CASE(4) GSS_CASE(6) GSS_CASE(8) GSS_CASE(10) GSS_CASE(12)
#define GSS_DP_PART1 GSS_CASE(16) GSS_CASE(20)
#define GSS_DP_PART2 GSS_CASE(24)
#define GSS_DP_LIST GSS_DP_PART0 GSS_DP_PART1 GSS_DP_PART2
#endif

namespace gss {

struct CacgmmParams {
    const float2* Y;          // (B,F,D,T)
    const uint8_t* activity;  // (B,K,T_act)
    const int* Tper;          // (B) valid frames per utterance (<= T) or null
    float* posterior;         // (B,F,K,T)
    double* weight_out;       // (B,F,K) or null
    double* logdet_out;       // (B,F,K) or null
    double* cov_out;          // (B,F,K,D,D) complex128 or null
    int* info;                // (B) or null
    int* slow_count;          // (2) or null : [0] slow-path (Jacobi) class updates, [1] Jacobi sweeps
    double* jac_v;            // (B*F*K, D, D+1) complex128 or null: eigenvectors of the previous pass (warm start)
    int* jac_has_v;           // (B*F*K) zeroed per launch
    int B, F, D, T, T_act;
    int iterations, iterations_post;
    double eps, floor_;
};

constexpr size_t cmax(size_t a, size_t b) { return a > b ? a : b; }

// cacgmm_generic.cu: any D < 35, K < 20 (runtime shapes)
size_t cacgmm_generic_ws_bytes(int B, int F, int D, int K);
int cacgmm_generic_launch(const void* Y, int y_is_c128, const uint8_t* activity, const int* Tper, float* posterior,
                          double* weight_out, double* logdet_out, double* cov_out, int* info,
                          int B, int F, int D, int T, int K, int T_act, int iterations, int iterations_post,
                          double eps, double floor_, void* ws, size_t ws_bytes, cudaStream_t st);
bool cacgmm_fast_path(int D, int K);

// E-phase sub-blocking of the packed lower triangle: SB x SB blocks so that only
// 2*SB complex values of the frame are live in registers at a time.
#ifndef GSS_SB_LARGE
#define GSS_SB_LARGE 6          // developer knob: sub-block edge for DP % 6 == 0 (A/B builds)
#endif
__host__ __device__ constexpr int sub_block(int DP) {
    return DP < 12 ? DP : (DP % GSS_SB_LARGE == 0 ? GSS_SB_LARGE : 4);
}
template <int DP, int K, int NT>
struct CacgmmCfg {
    static constexpr int NP = DP * (DP + 1) / 2;          // packed Hermitian pairs
    static constexpr int NB = DP / 2;                     // 2x2 block rows
    static constexpr int G = NB * (NB + 1) / 2;           // lanes per M-phase group
    static constexpr int NG = NT / G;                     // M-phase groups
    static constexpr int YLD = DP + 1;                    // padded smem row (elements)
    static constexpr int KP = (K + 1) & ~1;               // padded class count (16 B rows)
    static constexpr int NW = NT / 32;
    static constexpr int JLD = DP + 1;                    // leading dim of Jacobi matrices
    static constexpr int TE = NT;                         // frames per E step = super tile (complex64 tile)
#ifndef GSS_SPLIT_E
#define GSS_SPLIT_E 0           // measured alternative, see quad2_half: 44.8 vs 44.0 ms per utterance at cfg2
#endif
    static constexpr bool SPLIT_E = GSS_SPLIT_E && DP >= 12 && (NT / 2) % 32 == 0;   // E phase with two frames per thread
    static constexpr int SB = sub_block(DP);
    static constexpr int NS = DP / SB;
    static constexpr int NSB = NS * (NS + 1) / 2;
    // sweep: a thread owns a chunk of up to CH consecutive columns of one row of one class (registers);
    // CH = the smallest width for which the K * NCH chunks fit the block
    static constexpr int chunks_per_class(int ch) { int n = 0; for (int i = 0; i < DP; ++i) n += (i + ch) / ch; return n; }
    static constexpr int pick_ch() { int ch = 1; while (K * chunks_per_class(ch) > NT) ++ch; return ch; }
    static constexpr int CH = pick_ch();
    static constexpr int NCH = chunks_per_class(CH);
    static constexpr size_t E_BYTES = size_t(TE) * YLD * sizeof(float2);       // ONE resident super tile (E and M phase)
    static constexpr size_t SWEEP_BYTES = size_t(2) * K * DP * sizeof(cd) + 4 * K * 8 + size_t(K) * DP * 8 + 64;
    static constexpr size_t JAC_BYTES = size_t(3) * DP * JLD * sizeof(cd);
    static constexpr size_t YS_BYTES = cmax(E_BYTES, cmax(SWEEP_BYTES, JAC_BYTES));
    static constexpr size_t W_BYTES = size_t(TE) * KP * sizeof(double);
    static constexpr size_t B_BYTES = size_t(NP) * K * sizeof(cd);
    static constexpr size_t ACC_BYTES = size_t(K) * NP * sizeof(cd);
    // misc: logdet[32] pi[32] tr[32] apri[32] | gred[NW*K] | jred[64] | col[K*DP] cd | piv[K*DP] | rot[16] | flags[32] | table[NP] u16
    static constexpr size_t MISC_BYTES = (128 + NW * K + 64) * 8 + size_t(K) * DP * 8
                                         + 16 * sizeof(JacobiRot) + 32 * 4 + ((NP * 2 + 15) / 16) * 16;
    static constexpr size_t SMEM = YS_BYTES + W_BYTES + B_BYTES + ACC_BYTES + MISC_BYTES;
    static_assert(DP % 2 == 0 && DP <= 32, "padded channel count must be even and <= 32");
    static_assert(DP % SB == 0, "sub-block must divide DP");
    static_assert(NG >= 1, "block too small for the M-phase mapping");
    static_assert(K < 20, "cacgmm.py:247");
};

// Position of channel d inside a frame row of the resident tile: even channels first, then the odd
// ones ("planar"), so that the 2x2 blocks of consecutive M-phase lanes read consecutive 8 B chunks.
// The E phase indexes the row with compile-time channel numbers, so the permutation is free there.
template <int DP>
__host__ __device__ constexpr int ypos(int d) { return (d & 1) * (DP / 2) + (d >> 1); }

// q_k += sum over the pairs of sub-block (I,J) of B'_k[d,e] . P[d,e]
template <int DP, int K, int SB, int I, int J>
__device__ __forceinline__ void quad_subblock(const float2* __restrict__ yrow, const cd* __restrict__ Bsm,
                                              double (&qa)[K], double (&qb)[K]) {
    double rr[SB], ri[SB], cr[SB], ci[SB];
#pragma unroll
    for (int a = 0; a < SB; ++a) { const float2 v = yrow[ypos<DP>(I * SB + a)]; rr[a] = (double)v.x; ri[a] = (double)v.y; }
#pragma unroll
    for (int c = 0; c < SB; ++c) {
        if (I == J) { cr[c] = rr[c]; ci[c] = ri[c]; }
        else { const float2 v = yrow[ypos<DP>(J * SB + c)]; cr[c] = (double)v.x; ci[c] = (double)v.y; }
    }
#pragma unroll
    for (int a = 0; a < SB; ++a) {
        // all products of this row first (independent -> hides the FP64 latency), then the FMAs
        double pre[SB], pim[SB];
#pragma unroll
        for (int c = 0; c < SB; ++c) {
            if (I == J && c > a) continue;
            pre[c] = fma(rr[a], cr[c], ri[a] * ci[c]);
            pim[c] = fma(ri[a], cr[c], -(rr[a] * ci[c]));
        }
#pragma unroll
        for (int c = 0; c < SB; ++c) {
            if (I == J && c > a) continue;
            const cd* b = Bsm + tri(I * SB + a, J * SB + c) * K;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const cd bk = b[k];
#ifdef GSS_E_ONEACC      // developer knob: one accumulator chain per class (10 registers less)
                qa[k] = fma(pim[c], bk.y, fma(pre[c], bk.x, qa[k]));
#else
                qa[k] = fma(pre[c], bk.x, qa[k]);
                qb[k] = fma(pim[c], bk.y, qb[k]);
#endif
            }
        }
    }
}

template <int DP, int K, int IDX, int NSB>
struct QuadAll {
    static __device__ __forceinline__ void run(const float2* yrow, const cd* Bsm, double (&qa)[K], double (&qb)[K]) {
        constexpr int SB = sub_block(DP);
        constexpr int I = [] { int r = 0; while ((r + 1) * (r + 2) / 2 <= IDX) ++r; return r; }();
        constexpr int J = IDX - I * (I + 1) / 2;
        quad_subblock<DP, K, SB, I, J>(yrow, Bsm, qa, qb);
        QuadAll<DP, K, IDX + 1, NSB>::run(yrow, Bsm, qa, qb);
    }
};
template <int DP, int K, int NSB>
struct QuadAll<DP, K, NSB, NSB> {
    static __device__ __forceinline__ void run(const float2*, const cd*, double (&)[K], double (&)[K]) {}
};

// ---- E phase, two frames per thread (DP >= 12): MEASURED ALTERNATIVE, off by default (-DGSS_SPLIT_E=1)
// The thread-owns-a-frame E phase reads every B'_k entry once per warp as a warp-uniform LDS.128,
// which costs two shared-memory wavefronts: 10 wavefronts per pair against 7 SM cycles of FP64 work
// -- the phase runs at 91 % shared-memory pipe utilisation (ncu, profiles/r2_em_kernel_v7_*).  Here a
// thread accumulates TWO frames (tid mod NT/2 and that + NT/2) per B' load, and the two halves of the
// block (warp-uniform) split the sub-blocks of the packed triangle between them; the partial sums of
// the frame a thread does not own travel through the weight tile.  Rolled loops over SB2 x SB2
// sub-blocks (runtime sub-block coordinates): ~0.6 k instructions instead of 7 k.
// Result (B200, cfg2, profiles/r2_em_kernel_v8_split_e_*): shared-memory wavefronts of the kernel -23 %,
// E phase -8 %, but the extra barrier of the exchange and the per-sub-block conversions give it back:
// 44.8 ms per utterance against 44.0 for the unrolled form.  The E phase is then throttled by the FP64
// pipe itself (math_pipe_throttle is the top stall), which is where it should be.
__host__ __device__ constexpr int sub_block2(int DP) { return DP % 3 == 0 ? 3 : (DP % 4 == 0 ? 4 : 2); }

template <int DP>
__device__ __forceinline__ int ypos_rt(int d) { return (d & 1) * (DP / 2) + (d >> 1); }

template <int DP, int K, int SB, bool DIAG>
__device__ __forceinline__ void quad2_subblock(const float2* __restrict__ yA, const float2* __restrict__ yB,
                                               const cd* __restrict__ Bsm, const int I, const int J,
                                               double (&pA)[K], double (&pB)[K]) {
    double rA[SB], iA[SB], rB[SB], iB[SB], crA[SB], ciA[SB], crB[SB], ciB[SB];
#pragma unroll
    for (int a = 0; a < SB; ++a) {
        const int pos = ypos_rt<DP>(I * SB + a);
        const float2 u = yA[pos], v = yB[pos];
        rA[a] = (double)u.x; iA[a] = (double)u.y; rB[a] = (double)v.x; iB[a] = (double)v.y;
    }
#pragma unroll
    for (int c = 0; c < SB; ++c) {
        if (DIAG) { crA[c] = rA[c]; ciA[c] = iA[c]; crB[c] = rB[c]; ciB[c] = iB[c]; }
        else {
            const int pos = ypos_rt<DP>(J * SB + c);
            const float2 u = yA[pos], v = yB[pos];
            crA[c] = (double)u.x; ciA[c] = (double)u.y; crB[c] = (double)v.x; ciB[c] = (double)v.y;
        }
    }
#pragma unroll
    for (int a = 0; a < SB; ++a) {
        const int row = I * SB + a;
        const cd* __restrict__ brow = Bsm + (size_t)(row * (row + 1) / 2 + J * SB) * K;
#pragma unroll
        for (int c = 0; c < SB; ++c) {
            if (DIAG && c > a) continue;
            const double preA = fma(rA[a], crA[c], iA[a] * ciA[c]);
            const double pimA = fma(iA[a], crA[c], -(rA[a] * ciA[c]));
            const double preB = fma(rB[a], crB[c], iB[a] * ciB[c]);
            const double pimB = fma(iB[a], crB[c], -(rB[a] * ciB[c]));
            const cd* __restrict__ b = brow + c * K;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const cd bk = b[k];
                pA[k] = fma(pimA, bk.y, fma(preA, bk.x, pA[k]));
                pB[k] = fma(pimB, bk.y, fma(preB, bk.x, pB[k]));
            }
        }
    }
}

// sub-blocks of half `half` (0 / 1): every second diagonal block and every second off-diagonal block
template <int DP, int K>
__device__ __forceinline__ void quad2_half(const float2* __restrict__ yA, const float2* __restrict__ yB,
                                           const cd* __restrict__ Bsm, const int half,
                                           double (&pA)[K], double (&pB)[K]) {
    constexpr int SB = sub_block2(DP), NS = DP / SB;
#pragma unroll 1
    for (int I = half; I < NS; I += 2) quad2_subblock<DP, K, SB, true>(yA, yB, Bsm, I, I, pA, pB);
    int I = 1, J = 0;
    if (half) { ++J; if (J == I) { ++I; J = 0; } }
#pragma unroll 1
    while (I < NS) {
        quad2_subblock<DP, K, SB, false>(yA, yB, Bsm, I, J, pA, pB);
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) { ++J; if (J == I) { ++I; J = 0; } }
    }
}

// Symmetric sweep operator on the packed lower triangle of K class matrices at once (in place:
// -Phi^-1; pivots = squared Cholesky pivots).  The chunk of a thread (up to CH consecutive columns of
// one row of one class) stays in registers over all D sweep steps; per step only the pivot column u
// travels through shared memory (one barrier per step: the column of the next pivot is forwarded while
// the current step is applied).
// Branch-free step: with d = 1 / pivot and the pivot entry of the published column replaced by
// u_j := pivot - 1, the rank-1 update  S_ic - (u_i d) conj(u_c)  is ALSO right for the entries of row
// and column j  (S_ij - u_i d (p - 1) = u_i d;  conj(u_c) - (p - 1) d conj(u_c) = conj(u_c) d);  only the
// pivot entry itself is fixed up (the formula gives 2 - d, the sweep wants -d).
// Not inlined: the fused kernel around it is at its register limit and would keep the chunk in local memory.
template <int DP, int K, int CH>
__device__ __noinline__ void sweep_invert(cd (&sv_out)[CH], const cd* __restrict__ acc_k, const double* __restrict__ tr_s,
                                          cd* __restrict__ colb, double* __restrict__ pivb, double* __restrict__ dinvb,
                                          double* __restrict__ pivbuf, const bool sw_on, const int sw_k, const int sw_i,
                                          const int sw_c0, const int sw_n, const int D) {
    cd sv[CH];                                                          // registers for the whole function
#pragma unroll
    for (int n = 0; n < CH; ++n) sv[n] = cmake(0.0, 0.0);
    if (sw_on) {
        const double tr = tr_s[sw_k];
        const double itr = (tr > 0.0 && isfinite(tr)) ? 1.0 / tr : 0.0;
#pragma unroll
        for (int n = 0; n < CH; ++n) {
            const int c = sw_c0 + n;
            if (n < sw_n) {
                cd v = cscale(acc_k[tri(sw_i, c)], itr);
                if (sw_i == c) v.y = 0.0;                               // force_hermitian (utils.py:323-334)
                sv[n] = v;
            }
        }
        if (sw_c0 == 0) {                                               // column of the first pivot
            const cd v = sv[0];
            if (sw_i == 0) {
                pivb[sw_k] = v.x; dinvb[sw_k] = (v.x > 0.0 && isfinite(v.x)) ? 1.0 / v.x : 0.0;
                colb[sw_k * DP] = cmake(v.x - 1.0, 0.0);
            } else {
                colb[sw_k * DP + sw_i] = v;
            }
        }
    }
    __syncthreads();
    for (int j = 0; j < D; ++j) {
        const cd* col = colb + (j & 1) * K * DP;
        cd* coln = colb + ((j + 1) & 1) * K * DP;
        if (sw_on) {
            const double d = dinvb[(j & 1) * K + sw_k];
            const cd uid = cscale(col[sw_k * DP + sw_i], d);
            const cd* colk = col + sw_k * DP + sw_c0;
#pragma unroll
            for (int n = 0; n < CH; ++n) {
                if (n < sw_n) {
                    cd v = sv[n];
                    cfmsc(v, uid, colk[n]);
                    sv[n] = v;
                }
            }
            const int nj = j - sw_c0;                                   // position of column j in this chunk
            if (sw_i == j && nj >= 0 && nj < sw_n) {                    // the pivot entry: 2 - d  ->  -d
                pivbuf[sw_k * DP + j] = pivb[(j & 1) * K + sw_k];
#pragma unroll
                for (int n = 0; n < CH; ++n)
                    if (n == nj) { sv[n].x -= 2.0; sv[n].y = 0.0; }
            }
            // forward the column of the next pivot: entry (i, j + 1) of this chunk, or, for row j + 1,
            // the conjugates of its entries (j + 1, c < j + 1)
            const int nf = nj + 1;
            if (nf >= 0 && nf < sw_n) {
#pragma unroll
                for (int n = 0; n < CH; ++n) {
                    if (n == nf) {
                        if (sw_i == j + 1) {
                            const double pv = sv[n].x;
                            pivb[((j + 1) & 1) * K + sw_k] = pv;
                            dinvb[((j + 1) & 1) * K + sw_k] = (pv > 0.0 && isfinite(pv)) ? 1.0 / pv : 0.0;
                            coln[sw_k * DP + sw_i] = cmake(pv - 1.0, 0.0);
                        } else {
                            coln[sw_k * DP + sw_i] = sv[n];
                        }
                    }
                }
            }
            if (sw_i == j + 1) {
#pragma unroll
                for (int n = 0; n < CH; ++n)
                    if (n < sw_n && n != nf) coln[sw_k * DP + sw_c0 + n] = cconj(sv[n]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int n = 0; n < CH; ++n) sv_out[n] = sv[n];
}

template <int DP, int K, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) cacgmm_em_kernel(const CacgmmParams p) {
    using C = CacgmmCfg<DP, K, NT>;
    constexpr int TE = C::TE;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char* sp = smem_raw;
    unsigned char* ys_raw = sp;                     sp += C::YS_BYTES;   // frame tiles / matrix scratch
    double* wsm = reinterpret_cast<double*>(sp);    sp += C::W_BYTES;    // [TE][KP]
    cd* Bsm = reinterpret_cast<cd*>(sp);            sp += C::B_BYTES;    // [NP][K]
    cd* acc_sm = reinterpret_cast<cd*>(sp);         sp += C::ACC_BYTES;  // [K][NP]
    double* logdet_s = reinterpret_cast<double*>(sp);                    // [32]
    double* pi_s = logdet_s + 32;                                        // [32]
    double* tr_s = logdet_s + 64;                                        // [32]
    double* apri_s = logdet_s + 96;                                      // [32] pi_k exp(min_j logdet_j - logdet_k)
    double* gred = logdet_s + 128;                                       // [NW][K]
    double* jred = gred + C::NW * K;                                     // [64]
    double* pivbuf = jred + 64;                                          // [K][DP] sweep pivots
    JacobiRot* jrot = reinterpret_cast<JacobiRot*>(pivbuf + K * DP);     // [16]
    int* flags_s = reinterpret_cast<int*>(jrot + 16);                    // [K] slow-path flags, [31] exact flag
    unsigned short* tri_tab = reinterpret_cast<unsigned short*>(flags_s + 32);   // [NP] (i << 8 | c)
    float2* yf = reinterpret_cast<float2*>(ys_raw);                      // resident super tile [TE][YLD] complex64, planar rows

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int bf = blockIdx.x;
    const int b = bf / p.F, f = bf - b * p.F;
    const int D = p.D, Ts = p.T;                               // Ts: frame stride of the batch
    const int T = p.Tper ? min(max(p.Tper[b], 0), Ts) : Ts;    // valid frames of this utterance
    const float2* __restrict__ Yg = p.Y + (size_t)bf * D * Ts;
    if (T <= 0) {
        for (int i = tid; i < K * Ts; i += NT) p.posterior[(size_t)bf * K * Ts + i] = 0.f;
        return;
    }
    const uint8_t* __restrict__ act = p.activity + (size_t)b * K * p.T_act;

    // M-phase role: group m_g, lane m_l -> 2x2 block (bi >= bj)
    const int m_g = tid / C::G;
    const int m_l = tid - m_g * C::G;
    const bool m_active = m_g < C::NG;
    int bi = 0;
    while ((bi + 1) * (bi + 2) / 2 <= m_l) ++bi;
    const int bj = m_l - bi * (bi + 1) / 2;
    const int r0 = 2 * bi, c0 = 2 * bj;

    // sweep role: chunk q = tid -> class sw_k, row sw_i, columns [sw_c0, sw_c0 + sw_n)
    int sw_k = -1, sw_i = 0, sw_c0 = 0, sw_n = 0;
    if (tid < K * C::NCH) {
        sw_k = tid / C::NCH;
        int r = tid - sw_k * C::NCH;
        for (sw_i = 0; sw_i < DP; ++sw_i) {
            const int nc = (sw_i + C::CH) / C::CH;
            if (r < nc) break;
            r -= nc;
        }
        sw_c0 = r * C::CH;
        sw_n = min(C::CH, sw_i + 1 - sw_c0);
    }

    const int total_iters = p.iterations + (p.iterations_post - 1);
    const int NPD = D * (D + 1) / 2;

    for (int i = tid; i < C::NP * K; i += NT) Bsm[i] = cmake(0.0, 0.0);
    for (int i = tid; i < NPD; i += NT) {
        int d = 0;
        while ((d + 1) * (d + 2) / 2 <= i) ++d;
        tri_tab[i] = (unsigned short)((d << 8) | (i - d * (d + 1) / 2));
    }
    if (tid == 0) flags_s[31] = 0;
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
        int zero_frames = 0;

        for (int s0 = 0; s0 < T; s0 += TE) {
            const int s1 = min(s0 + TE, T);
            // ================= E step on the super tile (thread owns a frame) =================
            for (int i = tid; i < DP * TE; i += NT) {
                const int d = i / TE, t = i - d * TE;
                float2 v = make_float2(0.f, 0.f);
                if (d < D && s0 + t < T) v = __ldg(&Yg[(size_t)d * Ts + s0 + t]);
                yf[t * C::YLD + ypos<DP>(d)] = v;
            }
            __syncthreads();
            {
                double qa[K], qb[K];
#pragma unroll
                for (int k = 0; k < K; ++k) { qa[k] = 0.0; qb[k] = 0.0; }
                const float2* yrow = yf + tid * C::YLD;
                if constexpr (C::SPLIT_E) {
                    if (pass > 0) {
                        // two frames per thread, the halves of the block split the sub-blocks (see quad2_half)
                        const int half = tid >= NT / 2 ? 1 : 0, lt = tid - half * (NT / 2);
                        double pB[K];
#pragma unroll
                        for (int k = 0; k < K; ++k) pB[k] = 0.0;
                        quad2_half<DP, K>(yf + lt * C::YLD, yf + (lt + NT / 2) * C::YLD, Bsm, half, qa, pB);
                        // qa: frame lt, pB: frame lt + NT/2.  Hand the partial sums of the frame this thread
                        // does not own to its owner (thread = frame) through the weight tile.
                        const int other = half ? lt : lt + NT / 2;
#pragma unroll
                        for (int k = 0; k < K; ++k) wsm[other * C::KP + k] = half ? qa[k] : pB[k];
                        __syncthreads();
#pragma unroll
                        for (int k = 0; k < K; ++k) qa[k] = (half ? pB[k] : qa[k]) + wsm[tid * C::KP + k];
                    }
                } else {
                    if (pass > 0) QuadAll<DP, K, 0, C::NSB>::run(yrow, Bsm, qa, qb);
                }
                const int t = s0 + tid;
                if (t < s1) {
                    double n2 = 0.0;
#pragma unroll
                    for (int d = 0; d < DP; ++d) {
                        const float2 v = yrow[d];
                        n2 = fma((double)v.x, (double)v.x, fma((double)v.y, (double)v.y, n2));
                    }
                    const double s = n2 > 0.0 ? 1.0 / n2 : 0.0;
                    if (!(n2 > 0.0)) ++zero_frames;
                    double g[K], w[K];
                    if (pass == 0) {
                        // initialisation from the activity (core.py:156-160); q == 1
                        double tot = 0.0;
#pragma unroll
                        for (int k = 0; k < K; ++k) { g[k] = act[(size_t)k * p.T_act + t] ? 1.0 : 1e-10; tot += g[k]; }
#pragma unroll
                        for (int k = 0; k < K; ++k) { g[k] /= tot; w[k] = g[k] * s; }
                    } else {
                        // gamma_k ~ pi_k exp(-logdet_k) q_k^-D  (mixture_model_utils.py:32-53 with
                        // log_pdf = -D log q - logdet).  The softmax is invariant to a common factor, so
                        // instead of K logarithms and K exponentials: r_k = q_min / q_k <= 1, r_k^D by
                        // repeated squaring, times the per-pass class constants apri_k <= 1 -- no overflow,
                        // relative error ~D eps.  Frames where every active term is below 1e-200 (the
                        // reference's own max-subtraction keeps up to e^(logdet spread) more range there)
                        // take the reference's log / exp form.
                        double qn[K], iq[K];
                        double qmin = INFINITY;
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            const double q = fmax(fabs(qa[k] + qb[k]) * s, GSS_F64_TINY);
                            qn[k] = q;
                            qmin = fmin(qmin, q);
                        }
                        double u[K], bp[K];
#pragma unroll
                        for (int k = 0; k < K; ++k) { iq[k] = 1.0 / qn[k]; bp[k] = qmin * iq[k]; u[k] = 1.0; }
                        for (int e = D; e; e >>= 1) {
#pragma unroll
                            for (int k = 0; k < K; ++k) { if (e & 1) u[k] *= bp[k]; bp[k] *= bp[k]; }
                        }
                        double den = 0.0;
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            double a = u[k] * apri_s[k];
                            if (guided && !act[(size_t)k * p.T_act + t]) a = 0.0;
                            g[k] = a; den += a;
                        }
                        if (!(den > 1e-200)) {
                            double lp[K];
                            double mx = -INFINITY;
#pragma unroll
                            for (int k = 0; k < K; ++k) {
                                lp[k] = -(double)D * log(qn[k]) - logdet_s[k];
                                mx = fmax(mx, lp[k]);
                            }
                            den = 0.0;
#pragma unroll
                            for (int k = 0; k < K; ++k) {
                                double a = exp(lp[k] - mx) * pi_s[k];
                                if (guided && !act[(size_t)k * p.T_act + t]) a = 0.0;
                                g[k] = a; den += a;
                            }
                        }
                        const double iden = 1.0 / fmax(den, GSS_F64_TINY);
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            double v = g[k] * iden;
                            if (eps != 0.0) v = fmin(fmax(v, eps), 1.0 - eps);
                            g[k] = v;
                            w[k] = v * s * iq[k];
                        }
                    }
                    if (is_final) {
#pragma unroll
                        for (int k = 0; k < K; ++k)
                            p.posterior[((size_t)bf * K + k) * Ts + t] = (float)g[k];
                    }
#pragma unroll
                    for (int k = 0; k < K; ++k) { gsum[k] += g[k]; wsm[tid * C::KP + k] = w[k]; }
                }
            }
            __syncthreads();
            if (is_final) continue;

            // ================= M step on the super tile (lane owns a 2x2 block) =================
            cd macc[4][K];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int k = 0; k < K; ++k) macc[a][k] = cmake(0.0, 0.0);
            // The M phase reads the SAME resident complex64 tile as the E phase (planar rows: the 2x2
            // blocks of consecutive lanes read consecutive 8 B chunks); float32 -> float64 happens on
            // read.  Group m_g takes frames m_g, m_g + NG, ... of the super tile; operands of the next
            // frame are loaded ahead of the FMAs of the current one.
            if (m_active) {
                const int tn = s1 - s0;
                const float2* yrow = yf + m_g * C::YLD;
                const double* wr = wsm + m_g * C::KP;
                // no software prefetch: with 40 accumulators live it only produced spills; the other
                // warps of the two resident CTAs cover the shared-memory latency
#pragma unroll 1
                for (int t = m_g; t < tn; t += C::NG, yrow += C::NG * C::YLD, wr += C::NG * C::KP) {
                    const float2 fa0 = yrow[bi], fa1 = yrow[DP / 2 + bi], fb0 = yrow[bj], fb1 = yrow[DP / 2 + bj];
                    const cd a0 = cmake((double)fa0.x, (double)fa0.y), a1 = cmake((double)fa1.x, (double)fa1.y);
                    const cd b0 = cmake((double)fb0.x, (double)fb0.y), b1 = cmake((double)fb1.x, (double)fb1.y);
                    cd P[4];
                    P[0] = cmulc(a0, b0); P[1] = cmulc(a0, b1);
                    P[2] = cmulc(a1, b0); P[3] = cmulc(a1, b1);
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const double w = wr[k];
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            macc[a][k].x = fma(w, P[a].x, macc[a][k].x);
                            macc[a][k].y = fma(w, P[a].y, macc[a][k].y);
                        }
                    }
                }
            }
            __syncthreads();
            // ---- flush: reduce over the M-phase groups in fixed order (deterministic) ----
            if constexpr (C::NG >= 2 && C::NG <= 4 && size_t(C::NG - 1) * K * C::NP * sizeof(cd) <= C::YS_BYTES) {
                // groups 1.. park their partial blocks in slabs inside the (now dead) frame tile, group 0
                // goes straight to the accumulator; then every thread sums its entries in fixed order
                cd* slab = reinterpret_cast<cd*>(ys_raw);
                if (m_active) {
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const int d = r0 + (a >> 1), e = c0 + (a & 1);
                        if (e <= d) {
#pragma unroll
                            for (int k = 0; k < K; ++k) {
                                if (m_g == 0) {
                                    cd* dst = &acc_sm[k * C::NP + tri(d, e)];
                                    if (s0 == 0) *dst = macc[a][k];
                                    else { cd v = *dst; v.x += macc[a][k].x; v.y += macc[a][k].y; *dst = v; }
                                } else {
                                    slab[((m_g - 1) * K + k) * C::NP + tri(d, e)] = macc[a][k];
                                }
                            }
                        }
                    }
                }
                __syncthreads();
                for (int e = tid; e < K * C::NP; e += NT) {
                    cd v = acc_sm[e];
#pragma unroll
                    for (int g = 1; g < C::NG; ++g) { const cd u = slab[(g - 1) * K * C::NP + e]; v.x += u.x; v.y += u.y; }
                    acc_sm[e] = v;
                }
                __syncthreads();
            } else if constexpr (C::NG <= 4) {
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
            } else {
                // many small groups (few channels): dump the partial blocks of a chunk of groups into a
                // shared slab (the frame tiles are dead here) and let one thread sum each entry.
                constexpr int ENT = 4 * K;                        // entries per lane
                constexpr int PER_GROUP = C::G * ENT;             // entries per group
                constexpr int CHUNK = (int)(C::YS_BYTES / (PER_GROUP * sizeof(cd)));
                static_assert(CHUNK >= 1, "slab too small");
                cd* slab = reinterpret_cast<cd*>(ys_raw);
                for (int g0 = 0; g0 < C::NG; g0 += CHUNK) {
                    const int cnt = min(CHUNK, C::NG - g0);
                    if (m_active && m_g >= g0 && m_g < g0 + cnt) {
                        cd* dst = slab + (size_t)(m_g - g0) * PER_GROUP + m_l * ENT;
#pragma unroll
                        for (int a = 0; a < 4; ++a)
#pragma unroll
                            for (int k = 0; k < K; ++k) dst[a * K + k] = macc[a][k];
                    }
                    __syncthreads();
                    for (int e = tid; e < PER_GROUP; e += NT) {
                        const int l = e / ENT, ak = e - l * ENT, a = ak / K, k = ak - a * K;
                        int lb = 0;
                        while ((lb + 1) * (lb + 2) / 2 <= l) ++lb;
                        const int d = 2 * lb + (a >> 1), ee = 2 * (l - lb * (lb + 1) / 2) + (a & 1);
                        if (ee <= d) {
                            cd sum = cmake(0.0, 0.0);
                            for (int g = 0; g < cnt; ++g) { const cd v = slab[(size_t)g * PER_GROUP + e]; sum.x += v.x; sum.y += v.y; }
                            cd* dst = &acc_sm[k * C::NP + tri(d, ee)];
                            if (g0 == 0 && s0 == 0) *dst = sum;
                            else { cd v = *dst; v.x += sum.x; v.y += sum.y; *dst = v; }
                        }
                    }
                    __syncthreads();
                }
            }
        }
        if (is_final) {
            for (int i = tid; i < K * (Ts - T); i += NT) {
                const int k = i / (Ts - T), t = T + i - k * (Ts - T);
                p.posterior[((size_t)bf * K + k) * Ts + t] = 0.f;
            }
            break;
        }
        if (pass == 0) {
            // all-zero frames make q = tiny, where the eigenvalue normalisation of the reference
            // does not cancel any more -> this bin needs the exact (eigh) class matrices.
            const int any_zero = __syncthreads_or(zero_frames);
            if (tid == 0) flags_s[31] = any_zero ? 1 : 0;
        }

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
        // ---- traces ----
        for (int k = warp; k < K; k += C::NW) {
            double tr = 0.0;
            if (lane < D) tr = acc_sm[k * C::NP + tri(lane, lane)].x;
            tr = warp_sum(tr);
            if (lane == 0) { tr_s[k] = tr; flags_s[k] = (!(tr > 0.0) || !isfinite(tr) || flags_s[31]) ? 1 : 0; }
        }
        __syncthreads();

        // ---- fast path: B_k = (Phi_k / tr)^-1 for all classes at once with the symmetric sweep
        //      operator (in place, packed lower; pivots = Cholesky pivots^2, logdet = sum log pivots).
        //      Thread-private entry metadata is decoded once per pass; one barrier per sweep step
        //      (the column of the next pivot is forwarded while the current step is applied).
        cd* colb = reinterpret_cast<cd*>(ys_raw);                               // [2][K][DP] column of the pivot
        double* pivb = reinterpret_cast<double*>(colb + 2 * K * DP);            // [2][K] pivot
        double* dinvb = pivb + 2 * K;                                           // [2][K] 1 / pivot (0 if not positive): one division per class and step
        double* diagb = dinvb + 2 * K;                                          // [K][DP] diagonal of the result
        const bool need_exact = flags_s[31] != 0;
        if (!need_exact) {
            constexpr int CH = C::CH;
            cd sv[CH];
            const bool sw_on = sw_k >= 0 && sw_i < D;
            sweep_invert<DP, K, CH>(sv, acc_sm + (sw_on ? sw_k * C::NP : 0), tr_s, colb, pivb, dinvb, pivbuf,
                                    sw_on, sw_k, sw_i, sw_c0, sw_n, D);
            if (sw_on) {
#pragma unroll
                for (int n = 0; n < CH; ++n)
                    if (n < sw_n && sw_c0 + n == sw_i) diagb[sw_k * DP + sw_i] = -sv[n].x;
            }
            __syncthreads();
            // logdet, trace of the inverse, validity of the fast path
            for (int k = warp; k < K; k += C::NW) {
                double ld = 0.0, trb = 0.0;
                bool bad = false;
                if (lane < D) {
                    const double pv = pivbuf[k * DP + lane];
                    bad = !(pv > 0.0) || !isfinite(pv);
                    ld = bad ? 0.0 : log(pv);
                    trb = diagb[k * DP + lane];
                }
                ld = warp_sum(ld); trb = warp_sum(trb);
                bad = __any_sync(0xffffffffu, bad);
                // lambda_min >= 1/tr(B), lambda_max <= tr(Phi) = 1  =>  nothing is floored if
                // 1/tr(B) >= floor  (complex_angular_central_gaussian.py:118-121)
                if (bad || !(trb * p.floor_ <= 1.0) || !isfinite(trb) || !isfinite(ld)) { if (lane == 0) flags_s[k] = 1; }
                else if (lane == 0) logdet_s[k] = ld;
            }
            __syncthreads();
            if (sw_on && !flags_s[sw_k]) {
#pragma unroll
                for (int n = 0; n < CH; ++n) {
                    if (n >= sw_n) continue;
                    const int c = sw_c0 + n;
                    const cd v = sv[n];
                    Bsm[tri(sw_i, c) * K + sw_k] = (sw_i == c) ? cmake(-v.x, 0.0) : cmake(-2.0 * v.x, -2.0 * v.y);
                }
            }
            __syncthreads();
        }

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
            // Warm start: EM changes the covariance slowly, so the eigenvectors of the previous pass
            // almost diagonalise it: A' = V^H A V needs 1-2 sweeps instead of ~8.
            bool warm = false;
            cd* Vg = nullptr;
            if (p.jac_v != nullptr) {
                Vg = reinterpret_cast<cd*>(p.jac_v) + ((size_t)bf * K + k) * D * C::JLD;
                warm = p.jac_has_v[(size_t)bf * K + k] != 0;
            }
            if (warm) {
                cd* Tm = V + DP * C::JLD;                                  // third scratch matrix
                for (int i = tid; i < D * C::JLD; i += NT) V[i] = Vg[i];
                __syncthreads();
                for (int i = tid; i < D * D; i += NT) {                    // T = A V
                    const int r = i / D, c = i - r * D;
                    cd acc = cmake(0.0, 0.0);
                    for (int j = 0; j < D; ++j) cfma(acc, A[r * C::JLD + j], V[j * C::JLD + c]);
                    Tm[r * C::JLD + c] = acc;
                }
                __syncthreads();
                for (int i = tid; i < D * D; i += NT) {                    // A' = V^H T (lower, then mirrored)
                    const int r = i / D, c = i - r * D;
                    if (c > r) continue;
                    cd acc = cmake(0.0, 0.0);
                    for (int j = 0; j < D; ++j) cfma(acc, cconj(V[j * C::JLD + r]), Tm[j * C::JLD + c]);
                    if (r == c) acc.y = 0.0;
                    A[r * C::JLD + c] = acc;
                    if (r != c) A[c * C::JLD + r] = cconj(acc);
                }
                __syncthreads();
            }
            const int sweeps = block_jacobi_eigh(A, V, D, C::JLD, jrot, jred, tid, NT, !warm);
            if (sweeps < 0 && tid == 0 && p.info) atomicMax(&p.info[b], GSS_INFO_NO_CONVERGE | (f << 8));
            if (tid == 0 && p.slow_count) { atomicAdd(p.slow_count, 1); atomicAdd(p.slow_count + 1, sweeps < 0 ? 40 : sweeps); }
            if (Vg != nullptr) {
                for (int i = tid; i < D * C::JLD; i += NT) Vg[i] = V[i];
                if (tid == 0) p.jac_has_v[(size_t)bf * K + k] = 1;
            }
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
                const unsigned ic = tri_tab[i];
                const int d = ic >> 8, e = ic & 255;
                cd s = cmake(0.0, 0.0);
                for (int j = 0; j < D; ++j) {
                    cd t1 = cscale(V[d * C::JLD + j], lam[j]);
                    cfmac(s, t1, V[e * C::JLD + j]);
                }
                Bsm[tri(d, e) * K + k] = (d == e) ? cmake(s.x, 0.0) : cmake(2.0 * s.x, 2.0 * s.y);
            }
            __syncthreads();
        }

        // ---- per-pass class constants of the softmax: pi_k exp(min_j logdet_j - logdet_k) <= 1 ----
        if (tid < K) {
            double ldmin = logdet_s[0];
#pragma unroll
            for (int k = 1; k < K; ++k) ldmin = fmin(ldmin, logdet_s[k]);
            apri_s[tid] = pi_s[tid] * exp(ldmin - logdet_s[tid]);
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
template <int DP, int K>
static int launch_cacgmm_dk(const CacgmmParams& p, cudaStream_t st) {
#ifndef GSS_EM_NT
#define GSS_EM_NT 256
#endif
#ifndef GSS_EM_MINB
#define GSS_EM_MINB 2
#endif
    constexpr int NT = GSS_EM_NT;
    using C = CacgmmCfg<DP, K, NT>;
    // two CTAs per SM (128 registers) while two of them fit the shared memory; the K = 7, 8 variants of the
    // large channel counts need one CTA per SM (and get its 255 registers)
    constexpr int MINB = (2 * (C::SMEM + 1024) <= 228 * 1024) ? GSS_EM_MINB : 1;      // 228 KB per SM, 1 KB reserved per CTA
    auto kern = cacgmm_em_kernel<DP, K, NT, MINB>;
    GSS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    kern<<<p.B * p.F, NT, C::SMEM, st>>>(p);
    GSS_LAUNCH_CHECK("cacgmm_em_kernel");
    return GSS_OK;
}

template <int DP>
int launch_cacgmm_d(const CacgmmParams& p, int K, cudaStream_t st) {
    switch (K) {
#ifndef GSS_K_ONLY_5          // developer builds: -DGSS_K_ONLY_5 instantiates K = 5 only
        case 2: return launch_cacgmm_dk<DP, 2>(p, st);
        case 3: return launch_cacgmm_dk<DP, 3>(p, st);
        case 4: return launch_cacgmm_dk<DP, 4>(p, st);
        case 6: return launch_cacgmm_dk<DP, 6>(p, st);
        case 7: return launch_cacgmm_dk<DP, 7>(p, st);
        case 8: return launch_cacgmm_dk<DP, 8>(p, st);
#endif
        case 5: return launch_cacgmm_dk<DP, 5>(p, st);
        default: return fail(GSS_ERR_UNSUPPORTED, "gss_cacgmm_c64: K=%d not built (built: 2..8)", K);
    }
}

// explicit instantiations of this part, declarations of the others
#if GSS_EM_PART == 0
#define GSS_CASE(dp) template int launch_cacgmm_d<dp>(const CacgmmParams&, int, cudaStream_t);
GSS_DP_PART0
#undef GSS_CASE
#define GSS_CASE(dp) extern template int launch_cacgmm_d<dp>(const CacgmmParams&, int, cudaStream_t);
GSS_DP_PART1 GSS_DP_PART2
#undef GSS_CASE
#elif GSS_EM_PART == 1
#define GSS_CASE(dp) template int launch_cacgmm_d<dp>(const CacgmmParams&, int, cudaStream_t);
GSS_DP_PART1
#undef GSS_CASE
#else
#define GSS_CASE(dp) template int launch_cacgmm_d<dp>(const CacgmmParams&, int, cudaStream_t);
GSS_DP_PART2
#undef GSS_CASE
#endif

#if GSS_EM_PART == 0
// the fused kernel is built for these padded channel counts and K = 2..8; any D runs on the next
// built size (padded channels are zero rows; the matrix phase uses the true D)
bool cacgmm_fast_path(int D, int K) {
    if (K < 2 || K > 8) return false;
#define GSS_CASE(dp) if (D <= dp) return true;
    GSS_DP_LIST
#undef GSS_CASE
    return false;
}

int cacgmm_dispatch(const CacgmmParams& p, int K, cudaStream_t st) {
#define GSS_CASE(dp) if (p.D <= dp) return launch_cacgmm_d<dp>(p, K, st);
    GSS_DP_LIST
#undef GSS_CASE
    return fail(GSS_ERR_UNSUPPORTED, "gss_cacgmm_c64: D=%d not built", p.D);
}
#endif

}  // namespace gss

#if GSS_EM_PART == 0
extern "C" int gss_cacgmm_c64(const gss_c64* Y, const uint8_t* activity, float* posterior,
                              int iterations, int iterations_post,
                              double affiliation_eps, double eigenvalue_floor,
                              int B, int F, int D, int T, int K, int T_act, const int* T_per_utt,
                              double* weight_out, double* logdet_out, double* covariance_out,
                              int* info, void* ws, size_t ws_bytes, void* stream) {
    using namespace gss;
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
    if (B == 0 || F == 0) return GSS_OK;
    if (!cacgmm_fast_path(D, K))          // K > 6 or D > 24: runtime-shape kernel (cacgmm_generic.cu), same results
        return cacgmm_generic_launch(Y, 0, activity, T_per_utt, posterior, weight_out, logdet_out,
                                     covariance_out, info, B, F, D, T, K, T_act, iterations, iterations_post,
                                     affiliation_eps, eigenvalue_floor, ws, ws_bytes, (cudaStream_t)stream);
    CacgmmParams p;
    p.Y = (const float2*)Y; p.activity = activity; p.posterior = posterior; p.Tper = T_per_utt;
    p.weight_out = weight_out; p.logdet_out = logdet_out; p.cov_out = covariance_out;
    p.info = info; p.slow_count = nullptr;
    p.jac_v = nullptr; p.jac_has_v = nullptr;
    {   // optional warm-start store for the exact (Jacobi) path; without workspace the path starts cold
        Arena a(ws, ws_bytes);
        int* has = a.take<int>((size_t)B * F * K + 2);
        double* jv = a.take<double>((size_t)B * F * K * D * (D + 2) * 2);
        if (ws != nullptr && a.ok()) {
            GSS_CUDA(cudaMemsetAsync(has, 0, ((size_t)B * F * K + 2) * sizeof(int), (cudaStream_t)stream));
            p.jac_has_v = has; p.jac_v = jv; p.slow_count = has + (size_t)B * F * K;
        }
    }
    p.B = B; p.F = F; p.D = D; p.T = T; p.T_act = T_act;
    p.iterations = iterations; p.iterations_post = iterations_post;
    p.eps = affiliation_eps; p.floor_ = eigenvalue_floor;
    return cacgmm_dispatch(p, K, (cudaStream_t)stream);
}

// complex128 observations (the float64 hand-off from WPE): always the runtime-shape kernel
extern "C" int gss_cacgmm_c128(const double* Y, const uint8_t* activity, float* posterior,
                               int iterations, int iterations_post,
                               double affiliation_eps, double eigenvalue_floor,
                               int B, int F, int D, int T, int K, int T_act, const int* T_per_utt,
                               double* weight_out, double* logdet_out, double* covariance_out,
                               int* info, void* ws, size_t ws_bytes, void* stream) {
    using namespace gss;
    GSS_REQUIRE(Y && activity && posterior, GSS_ERR_ARG, "gss_cacgmm_c128: null pointer");
    GSS_REQUIRE(B >= 0 && F >= 0 && T > 0, GSS_ERR_ARG, "gss_cacgmm_c128: bad dims B=%d F=%d T=%d", B, F, T);
    GSS_REQUIRE(D > 1 && D < 35, GSS_ERR_ARG, "Channels: %d, sure? (cacgmm.py:196,248)", D);
    GSS_REQUIRE(K > 1 && K < 20, GSS_ERR_ARG, "num_classes: %d, sure? (cacgmm.py:212,247)", K);
    GSS_REQUIRE(iterations > 0, GSS_ERR_ARG, "iterations=%d must be > 0 (cacgmm.py:199)", iterations);
    GSS_REQUIRE(iterations_post >= 1, GSS_ERR_UNSUPPORTED,
                "iterations_post=%d: the reference raises TypeError for 0 (core.py:198-202)", iterations_post);
    GSS_REQUIRE(T_act >= T, GSS_ERR_ARG, "activity has %d frames, observation %d (cacgmm.py:216-218)", T_act, T);
    if (B == 0 || F == 0) return GSS_OK;
    return cacgmm_generic_launch(Y, 1, activity, T_per_utt, posterior, weight_out, logdet_out, covariance_out, info,
                                 B, F, D, T, K, T_act, iterations, iterations_post, affiliation_eps, eigenvalue_floor,
                                 ws, ws_bytes, (cudaStream_t)stream);
}
#endif  // GSS_EM_PART == 0
