// EXPERIMENTAL, libgss_dev.so only (not in the product library; reached only through gss_debug_mstep_i8):
// the CACGMM M-step covariance  Phi_k = sum_t w_kt y_t y_t^H  (ComplexAngularCentralGaussianTrainer._fit,
// pb_bss/distribution/complex_angular_central_gaussian.py:293-300) on the INT8 tensor cores with the
// exact digit-split arithmetic of wpe_gram_i8.cu.  It is the building block of the tensor-core EM
// iteration planned in DESIGN.md section 7: the weights change every iteration, so the A operand
// (rows (k, d, re/im) = w_kt y_t[d], 240 rows at D = 24, K = 5) cannot be pre-split in memory (1.1 MB of
// digit planes per bin and iteration); eight warps of the CTA generate its planes straight into the
// shared-memory pipeline stages (one FP64 FMA against the magic constant + byte permutes per value)
// while the tensor core consumes the previous stages.  The B operand (the 2 D real rows of Y) is
// static over the EM iterations: its planes are split once (mstep_i8_yplanes_kernel) and streamed
// with bulk copies.
//
// One CTA per (utterance, bin): warp 0 = TMA producer (B planes), warp 1 = MMA issuer + TMEM,
// warps 2..17 = A-plane generators during the main loop (two threads per row), warps 2..9 the epilogue
// afterwards.  Two M tiles of 128
// rows x N = 2 D columns x 5 accumulators = 480 TMEM columns.  Row scales: one power of two per
// class (max_t w_kt max_d |y_td|) and one per channel.
#include "tc_i8.cuh"
#include "../../include/gss_dev.h"
#include <algorithm>

namespace gss {

constexpr int MS_STAGES = 4;
constexpr int MS_GEN_WARPS = 16;
constexpr int MS_NT = 64 + 32 * MS_GEN_WARPS;
constexpr int MS_A_STAGE = 2 * (GI_BM / 8) * GI_BLK_BYTES;      // two M tiles: 40960
constexpr int MS_YLD = 33;                                       // row stride (float2) of the staged raw frames

struct MsDims { int F, D, T, K, KB; const int* Tper; };         // KB: 16-frame blocks (even)

__device__ __forceinline__ int ms_valid_frames(const MsDims& m, size_t bf) {
    return m.Tper ? min(max(m.Tper[bf / m.F], 0), m.T) : m.T;
}

// ---- per bin: channel exponents (max_t |y_d|), class exponents (max_t w_kt max_d |y_td|) -----------
__global__ void __launch_bounds__(256) mstep_i8_scale_kernel(const float2* __restrict__ Y, const double* __restrict__ w,
                                                             int* __restrict__ ey, int* __restrict__ ek, MsDims m) {
    extern __shared__ float ymax[];                               // [T]
    __shared__ float red[8];
    const size_t bf = blockIdx.x;
    const int T = m.T, Tv = ms_valid_frames(m, bf);
    const float2* __restrict__ Yg = Y + bf * (size_t)m.D * T;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int t = threadIdx.x; t < T; t += blockDim.x) ymax[t] = 0.f;
    __syncthreads();
    for (int d = warp; d < m.D; d += 8) {
        float mx = 0.f;
        for (int t = lane; t < Tv; t += 32) {
            const float2 v = __ldg(&Yg[(size_t)d * T + t]);
            const float a = fmaxf(fabsf(v.x), fabsf(v.y));
            mx = fmaxf(mx, a);
            atomicMax(reinterpret_cast<int*>(&ymax[t]), __float_as_int(a));   // non-negative floats order like ints
        }
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) ey[bf * m.D + d] = (mx > 0.f && isfinite(mx)) ? GI_HEADROOM - ilogbf(mx) : 0;
    }
    __syncthreads();
    for (int k = 0; k < m.K; ++k) {
        const double* __restrict__ wk = w + (bf * m.K + k) * (size_t)T;
        float mx = 0.f;
        for (int t = threadIdx.x; t < Tv; t += blockDim.x) mx = fmaxf(mx, (float)fabs(wk[t]) * ymax[t] * 1.0000002f);
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) red[warp] = mx;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
            ek[bf * m.K + k] = (mx > 0.f && isfinite(mx)) ? GI_HEADROOM - ilogbf(mx) : 0;
        }
        __syncthreads();
    }
}

// ---- digit planes of the 2 D real rows of Y: [k-step][8-row block][plane][half][row % 8][16] -------
__global__ void __launch_bounds__(256) mstep_i8_yplanes_kernel(const float2* __restrict__ Y, const int* __restrict__ ey,
                                                               int8_t* __restrict__ planes, MsDims m) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= m.D * m.KB) return;
    const int kb = idx / m.D, d = idx - kb * m.D;
    const size_t bf = blockIdx.y;
    const int T = m.T, Tv = ms_valid_frames(m, bf);
    const float2* __restrict__ Yg = Y + bf * (size_t)m.D * T + (size_t)d * T;
    const double sc = __longlong_as_double((long long)(1023 + ey[bf * m.D + d]) << 52);
    unsigned lo_re[16], hi_re[16], lo_im[16], hi_im[16];
#pragma unroll
    for (int tt = 0; tt < 16; ++tt) {
        const int t = kb * 16 + tt;
        float2 v = make_float2(0.f, 0.f);
        if (t < Tv) v = __ldg(&Yg[t]);
        const double zr = fma((double)v.x, sc, GI_MAGIC), zi = fma((double)v.y, sc, GI_MAGIC);
        lo_re[tt] = (unsigned)__double2loint(zr); hi_re[tt] = (unsigned)__double2hiint(zr);
        lo_im[tt] = (unsigned)__double2loint(zi); hi_im[tt] = (unsigned)__double2hiint(zi);
    }
    const int nrb = (2 * m.D) >> 3;
    int8_t* out = planes + bf * ((size_t)(m.KB >> 1) * nrb * GI_BLK_BYTES);
#pragma unroll
    for (int p = 0; p < GI_NS; ++p) {
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
        const size_t blk = (size_t)(kb >> 1) * nrb + (d >> 2);
        uint4* dst = reinterpret_cast<uint4*>(out + ((blk * GI_NS + p) * 2 + (kb & 1)) * 128 + (2 * d & 7) * 16);
        dst[0] = make_uint4(wr[0], wr[1], wr[2], wr[3]);
        dst[1] = make_uint4(wi[0], wi[1], wi[2], wi[3]);
    }
}

// ---- the M-step GEMM -------------------------------------------------------------------------------
__global__ void __launch_bounds__(MS_NT, 1) mstep_i8_kernel(const float2* __restrict__ Y, const double* __restrict__ w,
                                                            const int8_t* __restrict__ planes, const int* __restrict__ ey,
                                                            const int* __restrict__ ek, cd* __restrict__ Phi, MsDims m) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar_a[MS_STAGES], bar_b[MS_STAGES], bar_empty[MS_STAGES], bar_acc;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t bf = blockIdx.x;
    const int D = m.D, K = m.K, T = m.T, N = 2 * D, rows = K * N;
    const int Tv = ms_valid_frames(m, bf);
    const int nk = min(m.KB >> 1, (Tv + 31) >> 5);
    const int nrb = N >> 3;
    const uint32_t b_stage = (uint32_t)(nrb * GI_BLK_BYTES), stage_bytes = MS_A_STAGE + b_stage;
    const uint32_t smem_base = smem_u32(smem);
    cd* __restrict__ out = Phi + bf * (size_t)K * D * D;

    if (tid == 0) {
        for (int s = 0; s < MS_STAGES; ++s) {
            mbar_init(smem_u32(&bar_a[s]), MS_GEN_WARPS);
            mbar_init(smem_u32(&bar_b[s]), 1);
            mbar_init(smem_u32(&bar_empty[s]), 1);
        }
        mbar_init(smem_u32(&bar_acc), 1);
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

    if (nk == 0) {
        for (int i = tid; i < K * D * D; i += MS_NT) out[i] = cmake(0.0, 0.0);
    } else if (warp == 0) {
        // ===== TMA producer: the static B planes (2 D rows x 32 frames x 5 planes per stage) =====
        if (lane == 0) {
            const int8_t* __restrict__ pl = planes + bf * ((size_t)(m.KB >> 1) * nrb * GI_BLK_BYTES);
            for (int ks = 0; ks < nk; ++ks) {
                const int st = ks % MS_STAGES;
                if (ks >= MS_STAGES) mbar_wait(smem_u32(&bar_empty[st]), ((ks / MS_STAGES) - 1) & 1);
                const uint32_t full = smem_u32(&bar_b[st]);
                mbar_expect_tx(full, b_stage);
                bulk_g2s(smem_base + st * stage_bytes + MS_A_STAGE, pl + (size_t)ks * b_stage, b_stage, full);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: 2 M tiles x 15 digit pairs per stage, accumulator (tile, order) at column (5 tile + order) N =====
        const uint64_t desc_hi = ((uint64_t)((128u >> 4) & 0x3FFFu) << 16) | ((uint64_t)(((uint32_t)GI_BLK_BYTES >> 4) & 0x3FFFu) << 32) |
                                 ((uint64_t)1 << 46);
        const uint32_t idesc = umma_idesc_i8(N);
        for (int ks = 0; ks < nk; ++ks) {
            const int st = ks % MS_STAGES;
            mbar_wait(smem_u32(&bar_a[st]), (ks / MS_STAGES) & 1);
            mbar_wait(smem_u32(&bar_b[st]), (ks / MS_STAGES) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a_src = smem_base + st * stage_bytes, b_src = a_src + MS_A_STAGE;
                const uint64_t b0 = desc_hi | (uint64_t)((b_src >> 4) & 0x3FFFu);
                const uint32_t acc0 = ks > 0 ? 1u : 0u;
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const uint64_t a0 = desc_hi | (uint64_t)(((a_src + mt * (MS_A_STAGE / 2)) >> 4) & 0x3FFFu);
#pragma unroll
                    for (int p = 0; p < GI_NS; ++p)
#pragma unroll
                        for (int q = 0; q + p < GI_NS; ++q)
                            tc_mma_i8_ss(tmem + (uint32_t)((mt * GI_NS + p + q) * N), a0 + (uint64_t)(p * 16), b0 + (uint64_t)(q * 16),
                                         idesc, p > 0 ? 1u : acc0);
                }
                tc_commit(smem_u32(&bar_empty[st]));
            }
            __syncwarp();
        }
        if (elect_one()) tc_commit(smem_u32(&bar_acc));
        __syncwarp();
    } else {
        // ===== A-plane generators: thread = (row (k, d, re/im) of the 256-row A tile pair, 16-frame half).
        // The raw frames (D x 32 complex64) and the scaled weights (K x 32) of a k-step are staged in
        // shared memory by the same 512 threads (coalesced, prefetched one k-step ahead into registers). =====
        const int gid = tid - 64;                                 // 0..511
        const int r = gid >> 1, h = gid & 1;
        const bool live = r < rows;
        const int k = live ? r / N : 0, rem = r - k * N, d = rem >> 1, c = rem & 1;
        const float2* __restrict__ Yb = Y + bf * (size_t)D * T;
        const double* __restrict__ wb = w + bf * (size_t)K * T;
        float2* yraw = reinterpret_cast<float2*>(smem + MS_STAGES * stage_bytes);          // [2][D][MS_YLD] (odd row stride: the channels a warp reads hit distinct bank pairs)
        double* wraw = reinterpret_cast<double*>(yraw + 2 * D * MS_YLD);                   // [2][K][32]
        const uint32_t row_off = (uint32_t)(((r >> 7) * (GI_BM / 8) + ((r & 127) >> 3)) * GI_BLK_BYTES + (r & 7) * 16);
        // staging role: elements e = gid + 512 j of the (D x 32) frame tile, element gid of the (K x 32) weight tile
        float2 py[2];
        double pw = 0.0;
        const int wk_k = gid >> 5, wk_t = gid & 31;
        const double wk_sc = (wk_k < K) ? __longlong_as_double((long long)(1023 + ek[bf * K + wk_k]) << 52) : 0.0;
        auto prefetch = [&](int ks) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int e = gid + 512 * j, dd = e >> 5, t = ks * 32 + (e & 31);
                py[j] = (dd < D && t < Tv) ? __ldg(&Yb[(size_t)dd * T + t]) : make_float2(0.f, 0.f);
            }
            const int t = ks * 32 + wk_t;
            pw = (wk_k < K && t < Tv) ? wb[(size_t)wk_k * T + t] * wk_sc : 0.0;
        };
        auto stash = [&](int buf) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int e = gid + 512 * j;
                if (e < D * 32) yraw[buf * D * MS_YLD + (e >> 5) * MS_YLD + (e & 31)] = py[j];
            }
            if (wk_k < K) wraw[buf * K * 32 + gid] = pw;
        };
        prefetch(0);
        stash(0);
        asm volatile("bar.sync 1, 512;" ::: "memory");
        for (int ks = 0; ks < nk; ++ks) {
            const int st = ks % MS_STAGES, buf = ks & 1;
            if (ks + 1 < nk) prefetch(ks + 1);
            if (ks >= MS_STAGES) mbar_wait(smem_u32(&bar_empty[st]), ((ks / MS_STAGES) - 1) & 1);
            unsigned char* a_dst = smem + st * stage_bytes + row_off;
            const float2* yr = yraw + buf * D * MS_YLD + d * MS_YLD + h * 16;
            const double* wr = wraw + buf * K * 32 + k * 32 + h * 16;
            unsigned lo[16], hi[16];
#pragma unroll
            for (int tt = 0; tt < 16; ++tt) {
                const float2 v = yr[tt];
                const double z = live ? fma((double)(c ? v.y : v.x), wr[tt], GI_MAGIC) : GI_MAGIC;
                lo[tt] = (unsigned)__double2loint(z); hi[tt] = (unsigned)__double2hiint(z);
            }
#pragma unroll
            for (int p = 0; p < GI_NS; ++p) {
                unsigned wd[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    wd[q] = p == 0 ? gi_pack4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3], 0)
                                   : gi_pack4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3], 4 - p);
                *reinterpret_cast<uint4*>(a_dst + (p * 2 + h) * 128) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_a[st]));
            if (ks + 1 < nk) stash(buf ^ 1);                      // the other raw buffer: last read in k-step ks - 1
            asm volatile("bar.sync 1, 512;" ::: "memory");
        }
        if (warp >= 10) goto done;                                // the epilogue needs 8 warps (2 M tiles x 4 lane quarters)
        // ===== epilogue: the same warps; warp -> (M tile, TMEM lane quarter) =====
        const int q = warp & 3, mt = (warp - 2) >> 2;
        const int a = mt * GI_BM + 32 * q + lane;                 // A row = (class, channel, re/im)
        const bool row_ok = a < rows;
        const int ka = row_ok ? a / N : 0, da = (a - ka * N) >> 1;
        const bool odd = a & 1;
        const double row_scale = __longlong_as_double((long long)(1023 + 32 - (row_ok ? ek[bf * K + ka] : 0)) << 52);
        mbar_wait(smem_u32(&bar_acc), 0);
        tc_fence_after();
        for (int cb = 0; cb < N; cb += 8) {
            int acc[GI_NS][8];
#pragma unroll
            for (int o = 0; o < GI_NS; ++o) tc_ld8(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)((mt * GI_NS + o) * N + cb), acc[o]);
            tc_ld_wait();
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                double v0 = (double)acc[0][2 * jj], v1 = (double)acc[0][2 * jj + 1];
#pragma unroll
                for (int o = 1; o < GI_NS; ++o) {
                    v0 = fma(v0, 256.0, (double)acc[o][2 * jj]);
                    v1 = fma(v1, 256.0, (double)acc[o][2 * jj + 1]);
                }
                // even lane (re row): (rr, ri); odd lane (im row): (ir, ii).  (w y_d) conj(y_e): Re = rr + ii, Im = ir - ri
                const double other = __shfl_xor_sync(0xffffffffu, v1, 1);
                const double comb = odd ? v0 - other : v0 + other;
                const int e = (cb >> 1) + jj;                     // channel of the column pair
                if (!row_ok || e >= D) continue;
                const double val = comb * row_scale * __longlong_as_double((long long)(1023 - ey[bf * D + e]) << 52);
                double* dst = reinterpret_cast<double*>(&out[((size_t)ka * D + da) * D + e]);
                if (odd) dst[1] = (da == e) ? 0.0 : val; else dst[0] = val;
            }
        }
    }
done:
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    }
}

}  // namespace gss

// Phi (B,F,K,D,D) c128 = sum_t w[b,f,k,t] y y^H  with Y (B,F,D,T) c64, w (B,F,K,T) f64 >= 0.
// Built for D in {4, 8, 16, 24}, K * 2 D <= 256.  Workspace: B F (ceil(T/32) 2 D 160 + 4 (D + K)) bytes.
extern "C" int gss_debug_mstep_i8(const gss_c64* Y, const double* w, double* Phi, int B, int F, int D, int T, int K,
                                  const int* T_per_utt, void* ws, size_t ws_bytes, void* stream) {
    using namespace gss;
    GSS_REQUIRE(Y && w && Phi, GSS_ERR_ARG, "gss_debug_mstep_i8: null pointer");
    GSS_REQUIRE(B > 0 && F > 0 && T > 0 && K > 0, GSS_ERR_ARG, "gss_debug_mstep_i8: bad dims");
    GSS_REQUIRE((D == 4 || D == 8 || D == 16 || D == 24) && K * 2 * D <= 256, GSS_ERR_UNSUPPORTED,
                "gss_debug_mstep_i8: built for D in {4, 8, 16, 24} (UMMA N = 2 D), K * 2 D <= 256 (D=%d K=%d)", D, K);
    GSS_REQUIRE((long long)(T + 32) * 5 * 16384 < 2147483647LL, GSS_ERR_UNSUPPORTED, "gss_debug_mstep_i8: T=%d too long for INT32 sums", T);
    cudaStream_t st = (cudaStream_t)stream;
    MsDims m{F, D, T, K, (T + 31) / 32 * 2, T_per_utt};
    const size_t BF = (size_t)B * F;
    const size_t plane_bytes = (size_t)(m.KB >> 1) * ((2 * D) >> 3) * GI_BLK_BYTES;
    Arena a(ws, ws_bytes);
    int8_t* planes = a.take<int8_t>(BF * plane_bytes);
    int* ey = a.take<int>(BF * D);
    int* ek = a.take<int>(BF * K);
    GSS_REQUIRE(ws && a.ok(), GSS_ERR_WORKSPACE, "gss_debug_mstep_i8: workspace %zu < %zu", ws_bytes, a.off);
    mstep_i8_scale_kernel<<<(unsigned)BF, 256, (size_t)T * sizeof(float), st>>>((const float2*)Y, w, ey, ek, m);
    GSS_LAUNCH_CHECK("mstep_i8_scale_kernel");
    dim3 pg((D * m.KB + 255) / 256, (unsigned)BF);
    mstep_i8_yplanes_kernel<<<pg, 256, 0, st>>>((const float2*)Y, ey, planes, m);
    GSS_LAUNCH_CHECK("mstep_i8_yplanes_kernel");
    const size_t smem = (size_t)MS_STAGES * (MS_A_STAGE + ((2 * D) >> 3) * GI_BLK_BYTES) + (size_t)2 * (D * MS_YLD * sizeof(float2) + 32 * K * sizeof(double)) + 16;
    GSS_CUDA(cudaFuncSetAttribute(mstep_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mstep_i8_kernel<<<(unsigned)BF, MS_NT, smem, st>>>((const float2*)Y, w, planes, ey, ek, reinterpret_cast<cd*>(Phi), m);
    GSS_LAUNCH_CHECK("mstep_i8_kernel");
    return GSS_OK;
}
