/*
 * libgss -- C ABI of the B200-native guided-source-separation hot path.
 *
 * The reference (fgnt/pb_chime5) has NO FFI for this path: the seam is the set
 * of Python callables in pb_chime5/core.py (WPE 41-88, GSS 144-214, Beamformer
 * 241-278, Enhancer 281-571).  Each entry point below replaces the numeric body
 * of one of those callables; the ctypes binding a maintainer would add on the
 * reference side is shown in INTEGRATION.md.
 *
 * Conventions
 *  - All data pointers are DEVICE pointers (cudaMalloc'd / torch tensors'
 *    data_ptr()); the library never owns or frees caller memory.
 *  - Complex values are interleaved (re, im) float32 pairs ("c64").
 *  - "Bin-major" layout: the time axis is contiguous, one (D,T) slab per
 *    (utterance b, frequency bin f):   Y[b][f][d][t].  This is the layout the
 *    reference itself uses inside all three blocks (core.py:53,
 *    complex_angular_central_gaussian.py:55, beamforming_wrapper.py:24).  The
 *    reference's outer layouts (D,T,F) / (K,T,F) / (T,F) are produced/consumed
 *    by the gss_pack_* / gss_unpack_* entry points.
 *  - `stream` is a cudaStream_t passed as void*.  Calls are asynchronous with
 *    respect to the host; pass stream == NULL to use the legacy default stream.
 *  - Return value: 0 ok; < 0 argument / configuration error (nothing was
 *    launched); the message is available from gss_last_error() (thread local).
 *    No exception or abort crosses the ABI.
 *  - Numerical failures are detected on the device and reported through the
 *    optional `info` array (one int32 per utterance, device memory, may be
 *    NULL): 0 ok, otherwise GSS_INFO_* | (first failing bin << 8).
 *  - `ws` is caller-provided device scratch of at least gss_workspace_bytes().
 *  - Ragged batches: `T` is the frame stride of the batch; `T_per_utt` (device, one int32 per
 *    utterance, may be NULL = all T) gives the valid frames of each utterance.  Frames beyond
 *    are ignored on input and zero on output; results are identical to running the utterance alone.
 */
#ifndef GSS_H_
#define GSS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float re, im; } gss_c64;

#define GSS_OK                0
#define GSS_ERR_ARG          -1   /* bad pointer / shape / limit (D<35, K<20 ...) -> AssertionError */
#define GSS_ERR_UNSUPPORTED  -2   /* valid in the reference but not built here -> NotImplementedError */
#define GSS_ERR_WORKSPACE    -3   /* ws too small */
#define GSS_ERR_CUDA         -4   /* CUDA runtime error at launch */

#define GSS_INFO_OK           0
#define GSS_INFO_NOT_POSDEF   1   /* GEV: Phi_NN not positive definite (zhegvd INFO>N, get_gev_vector.pyx:139-147) */
#define GSS_INFO_NONFINITE    2   /* non-finite SNR in reference-channel search (beamformer.py:542) */
#define GSS_INFO_NO_CONVERGE  3   /* Jacobi eigensolver hit the sweep limit */
#define GSS_INFO_SINGULAR     4   /* WPE normal equations not positive definite */

#define GSS_OP_WEIGHTED_COV   0
#define GSS_OP_CACGMM         1
#define GSS_OP_BEAMFORM       2
#define GSS_OP_WPE            3
#define GSS_OP_STFT           4
#define GSS_OP_ISTFT          5
#define GSS_OP_ENHANCE        6
#define GSS_OP_BF_VECTOR      7
#define GSS_OP_CACGMM_C128     8
#define GSS_OP_ENHANCE_F64     9

#define GSS_BF_MVDR_SOUDEN_BAN 0  /* core.py:249-258 */
#define GSS_BF_GEV_BAN         1  /* beamforming_wrapper.py:192-208 */
#define GSS_BF_CH              2  /* core.py:259-260, channel index = bf_arg */
#define GSS_BF_SUM             3  /* core.py:261-262 */
#define GSS_BF_MVDR_SOUDEN     4  /* ban=False */
#define GSS_BF_GEV             5  /* ban=False */

#define GSS_POSTFILTER_NONE     0
#define GSS_POSTFILTER_MASK_MUL 1 /* core.py:270-271 */

int gss_version(void);
const char* gss_last_error(void);
/* Number of libgss kernels launched by this process so far (bench bookkeeping). */
long long gss_launch_count(void);

/* Scratch size for one call of `op` with these dimensions (L = WPE taps). */
int gss_workspace_bytes(int op, int B, int F, int D, int T, int K, int L, size_t* out);

/* ---- layout glue (pure permutations, replaces the morph()/transpose calls at
 *      core.py:53,58,181,208 and beamforming_wrapper.py:21-34) -------------- */
/* (B,D,T,F) -> (B,F,D,T) */
int gss_pack_dtf_to_fdt_c64(const gss_c64* src, gss_c64* dst, int B, int D, int T, int F, void* stream);
/* (B,F,D,T) -> (B,D,T,F) */
int gss_unpack_fdt_to_dtf_c64(const gss_c64* src, gss_c64* dst, int B, int D, int T, int F, void* stream);
/* float (B,F,K,T) -> (B,K,T,F) and back */
int gss_unpack_fkt_to_ktf_f32(const float* src, float* dst, int B, int K, int T, int F, void* stream);
int gss_pack_ktf_to_fkt_f32(const float* src, float* dst, int B, int K, int T, int F, void* stream);
/* complex (B,F,T) -> (B,T,F) */
int gss_unpack_ft_to_tf_c64(const gss_c64* src, gss_c64* dst, int B, int T, int F, void* stream);

/* ---- weighted spatial covariance  Phi[b,f,k] = sum_t w'[b,f,k,t] y y^H ------
 * Replaces get_power_spectral_density_matrix (pb_bss/extraction/beamformer.py:
 * 61-145) and the covariance einsum of ComplexAngularCentralGaussianTrainer._fit
 * (complex_angular_central_gaussian.py:293-300).
 * Y (B,F,D,T) c64;  w (B,F,K,T) f32;  Phi (B,F,K,D,D) c64, row-major, full
 * Hermitian matrix.
 * normalize_mode 0: w' = w;  1: w' = w / max(sum_t w, 1e-10) (beamformer.py:124) */
int gss_weighted_cov_c64(const gss_c64* Y, const float* w, gss_c64* Phi,
                         int normalize_mode, int B, int F, int D, int T, int K,
                         const int* T_per_utt, void* ws, size_t ws_bytes, void* stream);

/* ---- guided CACGMM EM (GSS.__call__, core.py:154-214; CACGMMTrainer.fit,
 * pb_bss/distribution/cacgmm.py:141-278; CACGMM.predict :63-94) --------------
 * Y (B,F,D,T) c64;  activity (B,K,T_act) u8 with T_act >= T (sliced [:T] as
 * core.py:182-184);  posterior out (B,F,K,T) f32.
 * iterations >= 1 guided EM iterations; iterations_post >= 1 (iterations_post-1
 * unguided EM iterations, then the final unguided, unclipped E-step).
 * Optional model outputs (may be NULL): weight (B,F,K) f64, logdet (B,F,K) f64,
 * covariance (B,F,K,D,D) c128 (normalised class covariance after the last
 * M-step, scaled to unit trace). */
int gss_cacgmm_c64(const gss_c64* Y, const uint8_t* activity, float* posterior,
                   int iterations, int iterations_post,
                   double affiliation_eps, double eigenvalue_floor,
                   int B, int F, int D, int T, int K, int T_act, const int* T_per_utt,
                   double* weight_out, double* logdet_out, double* covariance_out,
                   int* info, void* ws, size_t ws_bytes, void* stream);

/* ---- mask based beamforming (Beamformer.__call__, core.py:246-278;
 * beamforming_wrapper.py:11-124,192-208; beamformer.py:396-418,502-617) -------
 * Y (B,F,D,T) c64; target_mask, distortion_mask (B,F,T) f32; X_hat (B,F,T) c64.
 * bf_type GSS_BF_*; bf_arg = channel for GSS_BF_CH.  ref_channel_out (B) int32
 * device (may be NULL).  weights_out (B,F,D) c128 device (may be NULL). */
int gss_beamform_c64(const gss_c64* Y, const float* target_mask,
                     const float* distortion_mask, gss_c64* X_hat,
                     int bf_type, int bf_arg, int postfilter,
                     int B, int F, int D, int T, const int* T_per_utt,
                     int* ref_channel_out, double* weights_out,
                     int* info, void* ws, size_t ws_bytes, void* stream);

/* Variant fed directly by the posterior of gss_cacgmm_c64 (fuses core.py:537-554:
 * context-frame zeroing, target / distortion split).  posterior (B,F,K,T) f32;
 * target_index (B) int32 device; start_ctx/end_ctx (B) int32 device, frames. */
int gss_beamform_from_posterior_c64(const gss_c64* Y, const float* posterior,
                     const int* target_index, const int* start_ctx, const int* end_ctx,
                     gss_c64* X_hat, int bf_type, int bf_arg, int postfilter,
                     int B, int F, int D, int T, int K, const int* T_per_utt,
                     int* ref_channel_out, double* weights_out,
                     int* info, void* ws, size_t ws_bytes, void* stream);

/* ---- beamforming vectors from PSD matrices: the `get_bf_vector` DSL of pb_bss
 * (pb_bss/extraction/beamformer_wrapper.py:108-227): "[rank1_pca+|rank1_gev+]core[+ban]".
 * Phi_X, Phi_N (B,F,D,D) complex128 (re,im interleaved doubles), Phi_N may be NULL for 'pca' / 'chN';
 * w_out (B,F,D) complex128.  core = GSS_BFCORE_*; rank1: 0 none, 1 rank1_pca, 2 rank1_gev;
 * ref_channel: -1 = SNR-optimal over the F bins of each utterance (mvdr_souden / wmwf,
 * beamformer.py:524-543), else fixed; distortion_weight: wmwf mu (< 0: 'frequency_dependent');
 * pca_scaling: 0 none, 1 'trace', 2 'eigenvalue'.  Eigenvector phases (pca, gev), which LAPACK leaves
 * unspecified, are fixed deterministically (pca: largest component real positive; gev: (Phi_N w)[0] real
 * non-negative).  Workspace: gss_workspace_bytes(GSS_OP_BF_VECTOR, B, F, D, ...). */
#define GSS_BFCORE_MVDR_SOUDEN   0
#define GSS_BFCORE_GEV           1
#define GSS_BFCORE_WMWF          2
#define GSS_BFCORE_PCA           3
#define GSS_BFCORE_PCA_MVDR      4   /* 'pca+mvdr' */
#define GSS_BFCORE_GEVATF_MVDR   5   /* 'scaled_gev_atf+mvdr' */
#define GSS_BFCORE_CH            6   /* 'chN' */
/* gss_beamform_c64 / gss_beamform_from_posterior_c64 also accept a DSL program as `bf_type` (the
 * reference's default keyword arguments; `bf_arg` = N of 'chN'): */
#define GSS_BF_PROGRAM_FLAG      0x100
#define GSS_BF_PROGRAM(core, rank1, ban)  (GSS_BF_PROGRAM_FLAG | (core) | ((rank1) << 4) | ((ban) << 6))
int gss_bf_vector_c128(const double* Phi_X, const double* Phi_N, double* w_out,
                       int core, int rank1, int ban, int ref_channel, double distortion_weight,
                       int pca_scaling, int channel, int B, int F, int D,
                       int* ref_channel_out, int* info, void* ws, size_t ws_bytes, void* stream);

/* ---- WPE dereverberation (WPE.__call__, core.py:48-88 -> nara_wpe.wpe.wpe_v8,
 * third party) ---------------------------------------------------------------
 * Y, X (B,F,D,T) c64 (X may not alias Y).
 * The correlation build R = Yt L^-1 Yt^H, P = Yt L^-1 Y^H runs on the INT8 tensor cores
 * (tcgen05.mma kind::i8, exact digit-split integer arithmetic, float64 recombination) when
 * taps * D >= 48; bins whose normal equations are too ill conditioned for its 2^-38 truncation are
 * detected after the factorisation (a-posteriori check on the Cholesky pivots) and re-done with the
 * float64 (FP64 MMA) build inside the same call, still asynchronously.  A bin that was flagged once
 * stays on the float64 list for the remaining iterations of the call (no INT8 build for it). */
int gss_wpe_c64(const gss_c64* Y, gss_c64* X, int taps, int delay, int iterations,
                int psd_context, int B, int F, int D, int T, const int* T_per_utt,
                int* info, void* ws, size_t ws_bytes, void* stream);

/* Same with per-call options (no process-global switches; calls on different streams may differ):
 * gram_mode  GSS_WPE_GRAM_AUTO (what gss_wpe_c64 does), _F64 (float64 build for every bin),
 *            _I8 (INT8 build only, no re-do: diagnostics), _I8_REDO (INT8 + float64 re-do);
 * i8_tau     threshold of the a-posteriori check (< 0: default 1e-3);
 * stats      device int32[4] or NULL, ACCUMULATED (caller zeroes): [0] bins processed, [1] bins that
 *            ended the call on the float64 list, [2] float64 re-do builds (bins x iterations).
 * A caller that sees stats[1] / stats[0] > 1/2 (reverberant, low-noise recordings) should pass
 * GSS_WPE_GRAM_F64 for the following batches: pb_chime5_b200.core.WPE does exactly that.
 * X_c128   optional (B,F,D,T) complex128 copy of the result before it is rounded to complex64: the
 *            "float64 hand-off" to gss_cacgmm_c128 (the reference hands complex128 from block to block). */
#define GSS_WPE_GRAM_AUTO    -1
#define GSS_WPE_GRAM_F64      0
#define GSS_WPE_GRAM_I8       1
#define GSS_WPE_GRAM_I8_REDO  2
int gss_wpe_c64_ex(const gss_c64* Y, gss_c64* X, int taps, int delay, int iterations,
                   int psd_context, int B, int F, int D, int T, const int* T_per_utt,
                   int gram_mode, double i8_tau, int* stats, double* X_c128,
                   int* info, void* ws, size_t ws_bytes, void* stream);

/* gss_cacgmm_c64 on complex128 observations (B,F,D,T) -- the float64 hand-off from gss_wpe_c64_ex
 * (X_c128).  Always the runtime-shape kernel (csrc/cacgmm_generic.cu); same results as gss_cacgmm_c64
 * would give on unrounded input.  Workspace: gss_workspace_bytes(GSS_OP_CACGMM_C128, ...). */
int gss_cacgmm_c128(const double* Y, const uint8_t* activity, float* posterior,
                    int iterations, int iterations_post,
                    double affiliation_eps, double eigenvalue_floor,
                    int B, int F, int D, int T, int K, int T_act, const int* T_per_utt,
                    double* weight_out, double* logdet_out, double* covariance_out,
                    int* info, void* ws, size_t ws_bytes, void* stream);

/* ---- whole STFT-domain hot path in one call (Enhancer.enhance_observation without the
 * transforms, core.py:524-564), reference layouts on both sides ---------------------
 * Obs (B,D,T,F) c64; activity (B,K,T_act) u8; target_index / start_ctx / end_ctx (B) int32
 * device (context in frames, may be NULL = no context drop); T_per_utt (B) or NULL.
 * X_hat (B,T,F) c64; posterior (B,K,T,F) f32 or NULL (as GSS returns it: context NOT zeroed).
 * wpe_taps == 0 or wpe_iterations == 0 skips WPE (wpe_block is None).  Workspace:
 * gss_workspace_bytes(GSS_OP_ENHANCE, B, F, D, T, K, wpe_taps). */
int gss_enhance_c64(const gss_c64* Obs, const uint8_t* activity, const int* target_index,
                    const int* start_ctx, const int* end_ctx, const int* T_per_utt,
                    gss_c64* X_hat, float* posterior,
                    int wpe_taps, int wpe_delay, int wpe_iterations, int wpe_psd_context,
                    int em_iterations, int em_iterations_post,
                    int bf_type, int bf_arg, int postfilter,
                    int B, int F, int D, int T, int K, int T_act,
                    int* info, void* ws, size_t ws_bytes, void* stream);

/* The same with options.  flags: GSS_ENHANCE_F64_HANDOFF -- the dereverberated spectrum goes to the
 * EM in complex128 (as in the reference) instead of complex64: removes the only end-to-end difference
 * to the float64 chain that is not rounding of an output at the price of the
 * slower runtime-shape EM kernel.  Workspace: gss_workspace_bytes(GSS_OP_ENHANCE_F64, ...). */
#define GSS_ENHANCE_F64_HANDOFF  1
int gss_enhance_c64_ex(const gss_c64* Obs, const uint8_t* activity, const int* target_index,
                       const int* start_ctx, const int* end_ctx, const int* T_per_utt,
                       gss_c64* X_hat, float* posterior,
                       int wpe_taps, int wpe_delay, int wpe_iterations, int wpe_psd_context,
                       int em_iterations, int em_iterations_post,
                       int bf_type, int bf_arg, int postfilter, int flags,
                       int B, int F, int D, int T, int K, int T_act,
                       int* info, void* ws, size_t ws_bytes, void* stream);

/* ---- STFT / iSTFT (Enhancer.stft / .istft, core.py:305-321 -> nara_wpe.utils)
 * x (B,D,N) f32 -> Y (B,F,D,T) c64 bin-major, F = size/2+1,
 * T = ceil((N + 2*(size-shift)*fading - size + shift)/shift); Blackman window. */
int gss_stft_f32(const float* x, gss_c64* Y, int B, int D, int N,
                 int size, int shift, int fading, void* ws, size_t ws_bytes, void* stream);
/* X (B,F,T) c64 -> x (B, T*shift + size - shift - 2*(size-shift)*fading) f32 */
int gss_istft_f32(const gss_c64* X, float* x, int B, int T,
                  int size, int shift, int fading, void* ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* GSS_H_ */
