// gss_enhance_c64: the whole STFT-domain hot path of Enhancer.enhance_observation
// (pb_chime5/core.py:524-564) behind one C call, reference layouts in and out:
//   Obs (B,D,T,F) -> pack -> [WPE] -> guided CACGMM EM -> context drop + target/distortion
//   split + beamformer (+ postfilter) -> X_hat (B,T,F) [, posterior (B,K,T,F)]
// Pure orchestration of the other entry points on one stream; no extra kernels.
#include "common.cuh"

namespace gss {
size_t beamform_ws_bytes(int B, int F, int D);
size_t wpe_ws_bytes(int Bc, int F, int D, int T, int L);

struct EnhanceWs { float2* Yf; float2* Yw; double* Yw64; float* post; float2* Xf; int* stage_info; void* sub; size_t sub_bytes; size_t bytes; };

// One status word per stage and utterance inside the call; the caller's word gets the most severe
// one (a failure code of any stage wins over the WPE's "singular, minimum-norm solution used").
__global__ void enhance_merge_info_kernel(int* __restrict__ out, const int* __restrict__ st, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int best = 0, rank = 0;
    for (int s = 0; s < 3; ++s) {
        const int v = st[s * B + b], code = v & 0xFF;
        const int r = code == 0 ? 0 : (code == GSS_INFO_SINGULAR ? 1 : 2);
        if (r > rank) { rank = r; best = v; }
    }
    if (best) atomicMax(&out[b], best);
}

size_t cacgmm_ws_bytes(int B, int F, int D, int K);

size_t cacgmm_generic_ws_bytes(int B, int F, int D, int K);

static EnhanceWs enhance_layout(void* ws, int B, int F, int D, int T, int K, int L, bool wpe, bool f64 = false) {
    Arena a(ws, ~size_t(0));
    EnhanceWs w;
    w.Yf = a.take<float2>((size_t)B * F * D * T);
    w.Yw = wpe ? a.take<float2>((size_t)B * F * D * T) : nullptr;
    w.Yw64 = (wpe && f64) ? a.take<double>((size_t)2 * B * F * D * T) : nullptr;
    w.post = a.take<float>((size_t)B * F * K * T);
    w.Xf = a.take<float2>((size_t)B * F * T);
    w.stage_info = a.take<int>((size_t)3 * B);
    size_t sub = std::max(cacgmm_ws_bytes(B, F, D, K), beamform_ws_bytes(B, F, D));
    if (f64) sub = std::max(sub, cacgmm_generic_ws_bytes(B, F, D, K));
    if (wpe) sub = std::max(sub, wpe_ws_bytes(B < 8 ? B : 8, F, D, T, L));
    w.sub = a.take<char>(sub);
    w.sub_bytes = sub;
    w.bytes = a.off;
    return w;
}

size_t enhance_ws_bytes(int B, int F, int D, int T, int K, int L) {
    return enhance_layout(nullptr, B, F, D, T, K, L, L > 0).bytes;
}
size_t enhance_f64_ws_bytes(int B, int F, int D, int T, int K, int L) {
    return enhance_layout(nullptr, B, F, D, T, K, L, L > 0, true).bytes;
}
}  // namespace gss

extern "C" int gss_enhance_c64(const gss_c64* Obs, const uint8_t* activity, const int* target_index,
                               const int* start_ctx, const int* end_ctx, const int* T_per_utt,
                               gss_c64* X_hat, float* posterior,
                               int wpe_taps, int wpe_delay, int wpe_iterations, int wpe_psd_context,
                               int em_iterations, int em_iterations_post,
                               int bf_type, int bf_arg, int postfilter,
                               int B, int F, int D, int T, int K, int T_act,
                               int* info, void* ws, size_t ws_bytes, void* stream) {
    return gss_enhance_c64_ex(Obs, activity, target_index, start_ctx, end_ctx, T_per_utt, X_hat, posterior,
                              wpe_taps, wpe_delay, wpe_iterations, wpe_psd_context, em_iterations, em_iterations_post,
                              bf_type, bf_arg, postfilter, 0, B, F, D, T, K, T_act, info, ws, ws_bytes, stream);
}

extern "C" int gss_enhance_c64_ex(const gss_c64* Obs, const uint8_t* activity, const int* target_index,
                                  const int* start_ctx, const int* end_ctx, const int* T_per_utt,
                                  gss_c64* X_hat, float* posterior,
                                  int wpe_taps, int wpe_delay, int wpe_iterations, int wpe_psd_context,
                                  int em_iterations, int em_iterations_post,
                                  int bf_type, int bf_arg, int postfilter, int flags,
                                  int B, int F, int D, int T, int K, int T_act,
                                  int* info, void* ws, size_t ws_bytes, void* stream) {
    using namespace gss;
    GSS_REQUIRE((flags & ~GSS_ENHANCE_F64_HANDOFF) == 0, GSS_ERR_ARG, "gss_enhance_c64_ex: flags 0x%x", flags);
    GSS_REQUIRE(Obs && activity && target_index && X_hat, GSS_ERR_ARG, "gss_enhance_c64: null pointer");
    GSS_REQUIRE(B >= 0 && F > 0 && D > 0 && T > 0 && K > 1, GSS_ERR_ARG, "gss_enhance_c64: bad dims");
    if (B == 0) return GSS_OK;
    const bool wpe = wpe_taps > 0 && wpe_iterations > 0;
    const bool f64 = wpe && (flags & GSS_ENHANCE_F64_HANDOFF);       // without WPE the EM input is the exact complex64 STFT
    EnhanceWs w = enhance_layout(ws, B, F, D, T, K, wpe ? wpe_taps : 0, wpe, f64);
    GSS_REQUIRE(ws && ws_bytes >= w.bytes, GSS_ERR_WORKSPACE, "gss_enhance_c64: workspace %zu < %zu", ws_bytes, w.bytes);
    int* si = info ? w.stage_info : nullptr;           // [3][B]: WPE, EM, beamformer
    if (si) GSS_CUDA(cudaMemsetAsync(si, 0, sizeof(int) * 3 * (size_t)B, (cudaStream_t)stream));
    int rc = gss_pack_dtf_to_fdt_c64(Obs, (gss_c64*)w.Yf, B, D, T, F, stream);
    if (rc) return rc;
    const gss_c64* Y = (const gss_c64*)w.Yf;
    if (wpe) {
        rc = gss_wpe_c64_ex(Y, (gss_c64*)w.Yw, wpe_taps, wpe_delay, wpe_iterations, wpe_psd_context, B, F, D, T,
                            T_per_utt, GSS_WPE_GRAM_AUTO, -1.0, nullptr, w.Yw64, si, w.sub, w.sub_bytes, stream);
        if (rc) return rc;
        Y = (const gss_c64*)w.Yw;
    }
    if (f64)
        rc = gss_cacgmm_c128(w.Yw64, activity, w.post, em_iterations, em_iterations_post, 1e-10, 1e-10, B, F, D, T, K, T_act,
                             T_per_utt, nullptr, nullptr, nullptr, si ? si + B : nullptr, w.sub, w.sub_bytes, stream);
    else
        rc = gss_cacgmm_c64(Y, activity, w.post, em_iterations, em_iterations_post, 1e-10, 1e-10, B, F, D, T, K, T_act,
                            T_per_utt, nullptr, nullptr, nullptr, si ? si + B : nullptr, w.sub, w.sub_bytes, stream);
    if (rc) return rc;
    rc = gss_beamform_from_posterior_c64(Y, w.post, target_index, start_ctx, end_ctx, (gss_c64*)w.Xf, bf_type, bf_arg,
                                         postfilter, B, F, D, T, K, T_per_utt, nullptr, nullptr, si ? si + 2 * B : nullptr,
                                         w.sub, w.sub_bytes, stream);
    if (rc) return rc;
    if (si) {
        enhance_merge_info_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(info, si, B);
        GSS_LAUNCH_CHECK("enhance_merge_info_kernel");
    }
    rc = gss_unpack_ft_to_tf_c64((const gss_c64*)w.Xf, X_hat, B, T, F, stream);
    if (rc) return rc;
    if (posterior) rc = gss_unpack_fkt_to_ktf_f32(w.post, posterior, B, K, T, F, stream);
    return rc;
}
