// C ABI of libgss (include/gss.h): argument validation, workspace carving and
// dispatch.  No torch types, no exceptions across the boundary.
#include "common.cuh"
#include <cstring>
#include <atomic>

namespace gss {

static thread_local char g_err[512] = "";
char* last_error_buf() { return g_err; }

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return GSS_OK;
    return fail(GSS_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches() { return g_launches.load(); }

int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace gss

namespace gss {
size_t beamform_ws_bytes(int B, int F, int D);
size_t weighted_cov_ws_bytes(int B, int F, int D, int K);
size_t wpe_ws_bytes(int Bc, int F, int D, int T, int L);
size_t istft_ws_bytes(int B, int T, int size);
size_t enhance_ws_bytes(int B, int F, int D, int T, int K, int L);
size_t cacgmm_generic_ws_bytes(int B, int F, int D, int K);
size_t bf_vector_ws_bytes(int B, int F, int D);
size_t enhance_f64_ws_bytes(int B, int F, int D, int T, int K, int L);
bool cacgmm_fast_path(int D, int K);
size_t cacgmm_ws_bytes(int B, int F, int D, int K) {
    if (!cacgmm_fast_path(D, K)) return cacgmm_generic_ws_bytes(B, F, D, K);
    // warm-start eigenvectors of the exact path + flags
    return align_up(((size_t)B * F * K + 2) * sizeof(int)) + align_up((size_t)B * F * K * D * (D + 2) * 16);
}
}

extern "C" {

int gss_workspace_bytes(int op, int B, int F, int D, int T, int K, int L, size_t* out) {
    using namespace gss;
    GSS_REQUIRE(out, GSS_ERR_ARG, "gss_workspace_bytes: null out");
    GSS_REQUIRE(B >= 0 && F >= 0 && D >= 0 && T >= 0 && K >= 0 && L >= 0, GSS_ERR_ARG, "gss_workspace_bytes: negative dim");
    size_t n = 256;
    switch (op) {
        case GSS_OP_WEIGHTED_COV: n = weighted_cov_ws_bytes(B, F, D, K); break;
        case GSS_OP_CACGMM: n = cacgmm_ws_bytes(B, F, D, K); break;
        case GSS_OP_BEAMFORM: n = beamform_ws_bytes(B, F, D); break;
        case GSS_OP_WPE: n = wpe_ws_bytes(B < 8 ? B : 8, F, D, T, L); break;   // utterances are processed in chunks of up to 8
        case GSS_OP_STFT: n = 256; break;
        case GSS_OP_ISTFT: n = istft_ws_bytes(B, T, F > 1 ? 2 * (F - 1) : 2); break;   // F = size/2 + 1
        case GSS_OP_BF_VECTOR: n = bf_vector_ws_bytes(B, F, D); break;
        case GSS_OP_CACGMM_C128: n = cacgmm_generic_ws_bytes(B, F, D, K); break;
        case GSS_OP_ENHANCE_F64: n = enhance_f64_ws_bytes(B, F, D, T, K, L); break;
        case GSS_OP_ENHANCE: n = enhance_ws_bytes(B, F, D, T, K, L); break;   // L = WPE taps (0: no WPE)
        default: return fail(GSS_ERR_ARG, "gss_workspace_bytes: unknown op %d", op);
    }
    *out = n < 256 ? 256 : n;
    return GSS_OK;
}

int gss_version(void) { return 100; }   // 0.1.0
long long gss_launch_count(void) { return gss::launches(); }
const char* gss_last_error(void) { return gss::last_error_buf(); }

}  // extern "C"
