// WPE correlation build on the Blackwell INT8 tensor cores (tcgen05.mma kind::i8),
// exact integer arithmetic ("Ozaki" digit splitting), float64 recombination.
//
// What it computes (same contract as wpe_corr_kernel in wpe.cu): the lower trapezoid of
//   C[i][j] = sum_t inv_t a_i(t) conj(a_j(t)),   a = [Yt ; Y]  (nara_wpe.wpe.wpe_v6:
//   get_correlations / R = Yt Lambda^-1 Yt^H, P = Yt Lambda^-1 Y^H; call site
//   pb_chime5/core.py:52-58, algorithm per SURVEY.md appendix A).
//
// How.  u_r(t) = a_r(t) sqrt(inv_t) (power-whitened rows: bounded dynamic range), every
// complex row gets a power-of-two scale 2^e_r so that |u 2^e| < 2^38, the scaled value is
// rounded ONCE to a 40-bit integer x (one FP64 FMA with a magic constant) and split into
// five balanced base-256 digits d_p in [-128, 127] (x = sum_p d_p 256^(4-p)).  The real
// Gram matrix of the 2(LD+D) real rows (re/im interleaved) is then
//   sum_t x_a x_b = sum_{p,q} 256^(8-p-q) sum_t d_p(a,t) d_q(b,t)
// where every inner sum is an INT8 x INT8 -> INT32 tensor-core GEMM (exact).  Digit pairs
// with p + q <= 4 are kept (15 MMAs per k-step, five INT32 accumulators per tile in TMEM,
// one per order p + q); the dropped pairs are below 2^-38 of the row scales.  The epilogue
// recombines the five accumulators in float64 (Horner in base 256, relative error 2^-53), forms
// Re = rr + ii and Im = ir - ri between the two lanes of a complex row and undoes the row
// scales (powers of two).  Measured error vs the float64 Gram matrix:
// <= 6e-10 sqrt(R_ii R_jj) on adversarial envelopes, ~1e-11 typical (tests/test_gpu_wpe_i8.py);
// bins whose normal equations are too ill-conditioned for that are re-done by the FP64
// (DMMA) path, see wpe.cu.
//
// Kernels: wpe_i8_scale_kernel (row maxima -> exponents, sqrt(inv)), wpe_i8_slice_kernel
// (digit planes in the canonical no-swizzle K-major UMMA layout, so tiles are plain 1-D bulk
// copies), wpe_gram_i8_kernel (persistent; TMA producer warp / MMA issuer warp / 8 epilogue
// warps, 6-stage mbarrier pipeline, A planes and accumulators in TMEM).
#include "wpe_i8.cuh"
#include "tc_i8.cuh"
#include <algorithm>

namespace gss {

constexpr int GI_NMAX = 80;                              // real columns per tile: 5 accumulators x 80 + 2 x 40 columns of A planes <= 512 TMEM columns
constexpr int GI_TMEM_A = GI_NS * GI_NMAX;               // first TMEM column of the A-plane buffers (two, 8 columns per plane)
constexpr int GI_STAGES = 6;
constexpr int GI_EPI_WARPS = 8;
constexpr int GI_NT = 64 + 32 * GI_EPI_WARPS;            // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2..9 epilogue
constexpr int GI_A_STAGE = GI_BLK_BYTES * GI_BM / 8;     // 20480
constexpr int GI_B_STAGE = GI_BLK_BYTES * GI_NMAX / 8;   // 15360
constexpr int GI_STAGE_BYTES = GI_A_STAGE + GI_B_STAGE;  // 35840
constexpr int GI_SMEM = GI_STAGES * GI_STAGE_BYTES;      // 179200 (forces one CTA per SM: TMEM is allocated whole)
constexpr int GI_MAX_ITEMS = 96;

struct GiItem { short r0, c0, n; };                      // first real row of the A tile, first real column, width
struct GiPlan { int n_items; GiItem items[GI_MAX_ITEMS]; };
struct GiDims { int DP, NRc, NRp, KB; };                 // padded channel rows, complex rows DP + LD, padded real rows, 16-frame blocks (even)

// complex row rc of the digit planes -> (channel d, shift s); d < 0: padding row (zeros)
__device__ __forceinline__ void gi_row(const WpeDims& m, const GiDims& g, int rc, int& d, int& s) {
    if (rc < m.D) { d = rc; s = 0; }
    else if (rc < g.DP || rc >= g.NRc) { d = -1; s = 0; }
    else { const int i = rc - g.DP, k = i / m.D; d = i - k * m.D; s = m.delay + k; }
}

static GiDims gi_dims(int D, int T, int LD) {
    GiDims g;
    g.DP = (D + 7) / 8 * 8;                                // tap rows start on a 16-real-row boundary
    g.NRc = g.DP + LD;
    g.NRp = 2 * g.DP + (2 * LD + GI_BM - 1) / GI_BM * GI_BM;
    g.KB = (T + 31) / 32 * 2;
    return g;
}
__host__ __device__ static inline size_t gi_slice_bytes_per_bin(const GiDims& g) { return (size_t)GI_NS * g.KB * g.NRp * 16; }

bool wpe_i8_applicable(int D, int T, int L) {
    const int LD = L * D;
    // the digit sums must fit INT32: 2^14 per product, 5 pairs per order
    if ((long long)(T + 32) * 5 * 16384 >= 2147483647LL) return false;
    const int tiles = (2 * LD + GI_BM - 1) / GI_BM, DP2 = (D + 7) / 8 * 16;
    int items = 0;
    for (int i = 0; i < tiles; ++i) {
        int C = std::min(DP2 + GI_BM * (i + 1), DP2 + 2 * LD);
        C = (C + 15) / 16 * 16;
        items += (C + GI_NMAX - 1) / GI_NMAX;
    }
    return LD >= 48 && items <= GI_MAX_ITEMS;
}

static int gi_chunk_bins(int F, int D, int T, int L) {
    const GiDims g = gi_dims(D, T, L * D);
    const size_t per = gi_slice_bytes_per_bin(g);
    int n = (int)((size_t)104 << 20) / (int)per;         // keep one chunk of digit planes inside L2 (126 MB)
    return std::max(1, std::min(n, 64));
}

size_t wpe_i8_ws_layout(void* p, int F, int D, int T, int L, WpeI8Ws* out) {
    const GiDims g = gi_dims(D, T, L * D);
    const int cb = gi_chunk_bins(F, D, T, L);
    Arena a(p, ~size_t(0));
    out->slices = a.take<int8_t>((size_t)cb * gi_slice_bytes_per_bin(g));
    out->mu = a.take<double>((size_t)cb * T);
    out->ex = a.take<int>((size_t)cb * g.NRc);
    out->chunk_bins = cb;
    return a.off;
}

size_t wpe_i8_ws_bytes(int F, int D, int T, int L) {
    if (!wpe_i8_applicable(D, T, L)) return 0;
    WpeI8Ws w;
    return wpe_i8_ws_layout(nullptr, F, D, T, L, &w);
}

// ---------------------------------------------------------------------------------------------
// row exponents and sqrt(inv)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) wpe_i8_scale_kernel(const float2* __restrict__ Y, const double* __restrict__ inv,
                                                           double* __restrict__ mu, int* __restrict__ ex,
                                                           WpeDims m, GiDims g, size_t bf0, const int* __restrict__ skip) {
    extern __shared__ float mus[];                         // [T]
    const size_t bf = bf0 + blockIdx.x;
    if (skip && skip[bf]) return;
    const int T = m.T, Tv = wpe_valid_frames(m, bf);
    const float2* __restrict__ Yg = Y + bf * (size_t)m.D * T;
    const double* __restrict__ iv = inv + bf * (size_t)T;
    double* muo = mu + (size_t)blockIdx.x * T;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        const double v = t < Tv ? sqrt(iv[t]) : 0.0;
        if (blockIdx.y == 0) muo[t] = v;
        mus[t] = (float)v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int rc = warp + nw * blockIdx.y; rc < g.NRc; rc += nw * gridDim.y) {
        int d, s;
        gi_row(m, g, rc, d, s);
        float mx = 0.f;
        for (int t = s + lane; d >= 0 && t < Tv; t += 32) {
            const float2 v = __ldg(&Yg[(size_t)d * T + t - s]);
            mx = fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y)) * mus[t]);
        }
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) {
            int e = 0;
            if (mx > 0.f && isfinite(mx)) e = GI_HEADROOM - ilogbf(mx);
            ex[(size_t)blockIdx.x * g.NRc + rc] = e;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// digit planes.  Layout per bin: [k-step = t / 32][8-row block][plane p][half = (t / 16) % 2][row % 8][t % 16]
// int8: core matrices (8 rows x 16 B = 128 contiguous bytes) of the canonical K-major no-swizzle
// UMMA layout, the two 16-frame halves of a k-step 128 B apart (LBO), 8-row groups 1280 B apart
// (SBO), planes 256 B apart -- so ALL planes of any row range of one k-step are one contiguous
// run: a stage is two bulk copies (the TMA unit is issue-bound on many small copies: 20 copies
// of 2 KB per stage ran at 7 % tensor activity).  Real row 2 rc = Re, 2 rc + 1 = Im of complex row
// rc; rows [0, D) = Y (unshifted), rows [D, D + LD) = the taps.  Thread = (kb, complex row).
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256, 3) wpe_i8_slice_kernel(const float2* __restrict__ Y, const double* __restrict__ mu,
                                                           const int* __restrict__ ex, int8_t* __restrict__ slices,
                                                           WpeDims m, GiDims g, size_t bf0, const int* __restrict__ skip) {
    const int half = g.NRp >> 1;
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= half * g.KB) return;
    const int kb = idx / half, rc = idx - kb * half;
    const size_t bl = blockIdx.y, bf = bf0 + bl;
    if (skip && skip[bf]) return;
    const int T = m.T, Tv = wpe_valid_frames(m, bf);
    int8_t* out = slices + bl * gi_slice_bytes_per_bin(g);
    unsigned lo_re[16], hi_re[16], lo_im[16], hi_im[16];
    int d = -1, s = 0;
    gi_row(m, g, rc, d, s);
    const bool live = d >= 0;
    if (live) {
        const float2* __restrict__ Yg = Y + bf * (size_t)m.D * T + (size_t)d * T;
        const double* __restrict__ mub = mu + bl * (size_t)T;
        const int e = ex[bl * g.NRc + rc];
        const double sc = __longlong_as_double((long long)(1023 + e) << 52);
        const double magic = GI_MAGIC;
#pragma unroll
        for (int tt = 0; tt < 16; ++tt) {
            const int t = kb * 16 + tt, ts = t - s;
            float2 v = make_float2(0.f, 0.f);
            double w = 0.0;
            if (t < Tv && ts >= 0) { v = __ldg(&Yg[ts]); w = mub[t] * sc; }
            const double zr = fma((double)v.x, w, magic), zi = fma((double)v.y, w, magic);
            lo_re[tt] = (unsigned)__double2loint(zr); hi_re[tt] = (unsigned)__double2hiint(zr);
            lo_im[tt] = (unsigned)__double2loint(zi); hi_im[tt] = (unsigned)__double2hiint(zi);
        }
    }
#pragma unroll
    for (int p = 0; p < GI_NS; ++p) {
        // plane p = digit of weight 256^(4-p): byte 4-p of the 40-bit value (byte 4 = low byte of the high word)
        uint4 re4 = make_uint4(0u, 0u, 0u, 0u), im4 = make_uint4(0u, 0u, 0u, 0u);
        if (live) {
            unsigned wr[4], wi[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (p == 0) {
                    wr[q] = gi_pack4(hi_re[4 * q], hi_re[4 * q + 1], hi_re[4 * q + 2], hi_re[4 * q + 3], 0);
                    wi[q] = gi_pack4(hi_im[4 * q], hi_im[4 * q + 1], hi_im[4 * q + 2], hi_im[4 * q + 3], 0);
                } else {
                    wr[q] = gi_pack4(lo_re[4 * q], lo_re[4 * q + 1], lo_re[4 * q + 2], lo_re[4 * q + 3], 4 - p);
                    wi[q] = gi_pack4(lo_im[4 * q], lo_im[4 * q + 1], lo_im[4 * q + 2], lo_im[4 * q + 3], 4 - p);
                }
            }
            re4 = make_uint4(wr[0], wr[1], wr[2], wr[3]);
            im4 = make_uint4(wi[0], wi[1], wi[2], wi[3]);
        }
        const size_t blk = (size_t)(kb >> 1) * (g.NRp >> 3) + (rc >> 2);          // (k-step, 8-row block)
        uint4* dst = reinterpret_cast<uint4*>(out + ((blk * GI_NS + p) * 2 + (kb & 1)) * 128 + (2 * rc & 7) * 16);
        dst[0] = re4;
        dst[1] = im4;
    }
}


// ---------------------------------------------------------------------------------------------
// the GEMM.  Persistent CTAs (one per SM) walk the (bin, tile) work items of the chunk; a tile is
// 128 real rows x n <= 80 real columns, all five accumulators in TMEM (5 n <= 400 columns) plus two
// buffers of A planes (2 x 40 columns).
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2..9 = epilogue (two
// warps per TMEM lane quarter, alternating 16-column blocks).  The producer runs ahead into the
// next tile while the epilogue drains TMEM; the MMA warp waits for the drain (bar_drain).
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(GI_NT, 1) wpe_gram_i8_kernel(const int8_t* __restrict__ slices, const int* __restrict__ ex,
                                                               cd* __restrict__ Raug, double* __restrict__ rdiag,
                                                               WpeDims m, GiDims g, GiPlan plan, size_t bf0, int nbins,
                                                               const int* __restrict__ skip) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar_full[GI_STAGES], bar_empty[GI_STAGES], bar_acc, bar_drain;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem_base = smem_u32(smem);
    const int n_work = nbins * plan.n_items;
    const size_t bin_bytes = (size_t)GI_NS * g.KB * g.NRp * 16;
    const size_t nrb = (size_t)(g.NRp >> 3);

    if (tid == 0) {
        for (int s = 0; s < GI_STAGES; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
        mbar_init(smem_u32(&bar_acc), 1);
        mbar_init(smem_u32(&bar_drain), GI_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        // ===== TMA producer: two bulk copies per stage (A: 20 KB, B: n * 160 B) =====
        if (lane == 0) {
            int kg = 0;                                    // k-steps issued so far (all tiles)
            for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
                const int bl = w / plan.n_items;
                const GiItem it = plan.items[w - bl * plan.n_items];
                const int Tv = wpe_valid_frames(m, bf0 + bl);
                const int nk = (skip && skip[bf0 + bl]) ? 0 : min(g.KB >> 1, (Tv + 31) >> 5);
                const int8_t* __restrict__ sl = slices + (size_t)bl * bin_bytes;
                const uint32_t stage_tx = (uint32_t)(GI_BLK_BYTES * (GI_BM / 8 + it.n / 8));
                for (int ks = 0; ks < nk; ++ks, ++kg) {
                    const int st = kg % GI_STAGES;
                    if (kg >= GI_STAGES) mbar_wait(smem_u32(&bar_empty[st]), ((kg / GI_STAGES) - 1) & 1);
                    const uint32_t full = smem_u32(&bar_full[st]);
                    mbar_expect_tx(full, stage_tx);
                    const uint32_t a_dst = smem_base + st * GI_STAGE_BYTES, b_dst = a_dst + GI_A_STAGE;
                    bulk_g2s(a_dst, sl + ((size_t)ks * nrb + (it.r0 >> 3)) * GI_BLK_BYTES, GI_BLK_BYTES * (GI_BM / 8), full);
                    bulk_g2s(b_dst, sl + ((size_t)ks * nrb + (it.c0 >> 3)) * GI_BLK_BYTES, GI_BLK_BYTES * (it.n / 8), full);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer.  Per stage: the five A planes go shared memory -> TMEM once (tcgen05.cp),
        // then the 15 digit-pair MMAs read A from TMEM and only B from shared memory (with A in shared
        // memory every MMA re-read 4 KB of it and the kernel was shared-memory-bandwidth bound).
        // Accumulator of order p + q at TMEM column (p + q) n; A planes double buffered at GI_TMEM_A.
        // The whole warp walks the loop (waits), one elected lane issues.
        {
            // descriptor without the address: LBO = 128 B (k halves), SBO = 1280 B (8-row groups), version 1
            const uint64_t desc_hi = ((uint64_t)((128u >> 4) & 0x3FFFu) << 16) | ((uint64_t)(((uint32_t)GI_BLK_BYTES >> 4) & 0x3FFFu) << 32) |
                                     ((uint64_t)1 << 46);
            int kg = 0, tile = 0;
            for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
                const int bl = w / plan.n_items;
                const GiItem it = plan.items[w - bl * plan.n_items];
                const int n = it.n;
                const int Tv = wpe_valid_frames(m, bf0 + bl);
                const int nk = (skip && skip[bf0 + bl]) ? 0 : min(g.KB >> 1, (Tv + 31) >> 5);
                if (nk == 0) continue;
                if (tile > 0) { mbar_wait(smem_u32(&bar_drain), (tile - 1) & 1); tc_fence_after(); }   // TMEM drained
                const uint32_t idesc = umma_idesc_i8(n);
                for (int ks = 0; ks < nk; ++ks, ++kg) {
                    const int st = kg % GI_STAGES;
                    mbar_wait(smem_u32(&bar_full[st]), (kg / GI_STAGES) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a_src = smem_base + st * GI_STAGE_BYTES, b_src = a_src + GI_A_STAGE;
                        const uint64_t a0 = desc_hi | (uint64_t)((a_src >> 4) & 0x3FFFu), b0 = desc_hi | (uint64_t)((b_src >> 4) & 0x3FFFu);
                        const uint32_t ta = tmem + (uint32_t)(GI_TMEM_A + (kg & 1) * (GI_NS * 8));
                        const uint32_t acc0 = ks > 0 ? 1u : 0u;
#pragma unroll
                        for (int p = 0; p < GI_NS; ++p) tc_cp_128x256b(ta + (uint32_t)(p * 8), a0 + (uint64_t)(p * 16));
#pragma unroll
                        for (int p = 0; p < GI_NS; ++p)
#pragma unroll
                            for (int q = 0; q + p < GI_NS; ++q)   // planes are 256 B apart: +16 in the address field
                                tc_mma_i8_ts(tmem + (uint32_t)((p + q) * n), ta + (uint32_t)(p * 8), b0 + (uint64_t)(q * 16), idesc,
                                             p > 0 ? 1u : acc0);
                        tc_commit(smem_u32(&bar_empty[st]));   // stage free once the copies and MMAs have read it
                    }
                    __syncwarp();
                }
                if (elect_one()) tc_commit(smem_u32(&bar_acc));      // accumulators complete
                __syncwarp();
                ++tile;
            }
        }
    } else {
        // ===== epilogue: TMEM -> float64 recombination -> complex -> Raug =====
        const int ew = warp - 2;
        const int q = warp & 3;                            // TMEM lane quarter this warp may access
        const int sub = ew >> 2;                           // which of the two warps of the quarter
        int tile = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
            const int bl = w / plan.n_items;
            const GiItem it = plan.items[w - bl * plan.n_items];
            const int n = it.n;
            const size_t bf = bf0 + bl;
            if (skip && skip[bf]) continue;                // float64 list: nothing was produced, nothing to store
            const int Tv = wpe_valid_frames(m, bf);
            const int nk = min(g.KB >> 1, (Tv + 31) >> 5);
            const int a = it.r0 + 32 * q + lane;           // real row (digit-plane numbering)
            const int rca = a >> 1, i = rca - g.DP;
            const bool odd = a & 1;
            const bool row_ok = i >= 0 && i < m.LD;
            cd* __restrict__ Rb = Raug + bf * (size_t)(m.LD + m.D) * m.LD;
            if (nk == 0) {
                // no valid frames: the trapezoid is zero
                if (row_ok && !odd && sub == 0) {
                    for (int cc = it.c0 / 2; cc < (it.c0 + n) / 2 && cc < g.NRc; ++cc) {
                        const int j = cc - g.DP;
                        if (cc < m.D) Rb[(size_t)(m.LD + cc) * m.LD + i] = cmake(0.0, 0.0);
                        else if (j >= 0 && j <= i) { Rb[(size_t)i * m.LD + j] = cmake(0.0, 0.0); if (j == i && rdiag) rdiag[bf * (size_t)m.LD + i] = 0.0; }
                    }
                }
                continue;
            }
            const int* __restrict__ exb = ex + (size_t)bl * g.NRc;
            const int ea = (rca < g.NRc) ? exb[rca] : 0;
            const double row_scale = __longlong_as_double((long long)(1023 + 32 - ea) << 52);
            mbar_wait(smem_u32(&bar_acc), tile & 1);
            tc_fence_after();
            for (int cb = 16 * sub; cb < n; cb += 32) {
                int acc[GI_NS][16];
#pragma unroll
                for (int o = 0; o < GI_NS; ++o) tc_ld16(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(o * n + cb), acc[o]);
                tc_ld_wait();
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    // sum_o acc_o 256^(4-o): |.| < 2^57, float64 Horner (relative error 2^-53)
                    double v0 = (double)acc[0][2 * jj], v1 = (double)acc[0][2 * jj + 1];
#pragma unroll
                    for (int o = 1; o < GI_NS; ++o) {
                        v0 = fma(v0, 256.0, (double)acc[o][2 * jj]);
                        v1 = fma(v1, 256.0, (double)acc[o][2 * jj + 1]);
                    }
                    // even lane holds (rr, ri), odd lane (ir, ii):  Re = rr + ii,  Im = ir - ri
                    const double other = __shfl_xor_sync(0xffffffffu, v1, 1);
                    const double comb = odd ? v0 - other : v0 + other;
                    const int cc = (it.c0 + cb) / 2 + jj;  // complex column (digit-plane numbering)
                    if (!row_ok || cc >= g.NRc || (cc >= m.D && cc < g.DP)) continue;
                    const double val = comb * row_scale * __longlong_as_double((long long)(1023 - exb[cc]) << 52);
                    if (cc < m.D) {
                        // (tap row i, channel cc): conj goes to the P^H rows of Raug
                        double* dst = reinterpret_cast<double*>(&Rb[(size_t)(m.LD + cc) * m.LD + i]);
                        if (odd) dst[1] = -val; else dst[0] = val;
                    } else {
                        const int j = cc - g.DP;
                        if (j <= i) {
                            double* dst = reinterpret_cast<double*>(&Rb[(size_t)i * m.LD + j]);
                            if (odd) dst[1] = (i == j) ? 0.0 : val;
                            else {
                                dst[0] = val;
                                if (i == j && rdiag) rdiag[bf * (size_t)m.LD + i] = val;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_drain));   // this warp has read its part of TMEM
            ++tile;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
static GiPlan gi_plan(int D, int LD) {
    GiPlan pl;
    pl.n_items = 0;
    const int tiles = (2 * LD + GI_BM - 1) / GI_BM, DP2 = (D + 7) / 8 * 16;
    // heavy (wide) row tiles first: better tail behaviour
    for (int i = tiles - 1; i >= 0; --i) {
        int C = std::min(DP2 + GI_BM * (i + 1), DP2 + 2 * LD);
        C = (C + 15) / 16;                                 // 16-column units
        const int nch = (C * 16 + GI_NMAX - 1) / GI_NMAX;
        int c0 = 0;
        for (int c = 0; c < nch; ++c) {
            const int w = C / nch + (c < C % nch ? 1 : 0);
            GiItem itm;
            itm.r0 = (short)(DP2 + GI_BM * i); itm.c0 = (short)(c0 * 16); itm.n = (short)(w * 16);
            pl.items[pl.n_items++] = itm;
            c0 += w;
        }
    }
    return pl;
}

int wpe_gram_i8_run(const float2* Y, const double* inv, cd* Raug, double* rdiag, const WpeDims& m, int BF,
                    const WpeI8Ws& ws, const int* skip, cudaStream_t st) {
    const GiDims g = gi_dims(m.D, m.T, m.LD);
    const GiPlan plan = gi_plan(m.D, m.LD);
    GSS_CUDA(cudaFuncSetAttribute(wpe_gram_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GI_SMEM));   // per device
    const int half = g.NRp >> 1;
    // bins per chunk: at most what the scratch holds; the count whose work items fill whole waves of
    // the persistent CTAs best (ties: the larger chunk)
    int cbins = 1;
    {
        double best = 0.0;
        for (int c = 1; c <= ws.chunk_bins; ++c) {
            const int items = c * plan.n_items, waves = (items + num_sms() - 1) / num_sms();
            const double util = (double)items / ((double)waves * num_sms());
            if (util >= best - 1e-9) { best = util; cbins = c; }
        }
    }
    for (int b0 = 0; b0 < BF; b0 += cbins) {
        const int nb = std::min(cbins, BF - b0);
        wpe_i8_scale_kernel<<<dim3(nb, 8), 256, (size_t)m.T * sizeof(float), st>>>(Y, inv, ws.mu, ws.ex, m, g, (size_t)b0, skip);
        GSS_LAUNCH_CHECK("wpe_i8_scale_kernel");
        dim3 sg((half * g.KB + 255) / 256, nb);
        wpe_i8_slice_kernel<<<sg, 256, 0, st>>>(Y, ws.mu, ws.ex, ws.slices, m, g, (size_t)b0, skip);
        GSS_LAUNCH_CHECK("wpe_i8_slice_kernel");
        const int grid = std::min(plan.n_items * nb, num_sms());
        wpe_gram_i8_kernel<<<grid, GI_NT, GI_SMEM, st>>>(ws.slices, ws.ex, Raug, rdiag, m, g, plan, (size_t)b0, nb, skip);
        GSS_LAUNCH_CHECK("wpe_gram_i8_kernel");
    }
    return GSS_OK;
}

}  // namespace gss
