// Pure permutations between the reference's outer layouts and the bin-major
// device layout (replaces morph()/transpose at core.py:53,58,181,208 and
// beamforming_wrapper.py:21-34).  Tiled through shared memory so that both the
// read and the write side are coalesced.
#include "common.cuh"

namespace gss {

// Batched 2-D transpose: src [batch][R][C] -> dst [batch][C][R], element type E.
template <typename E>
__global__ void transpose_kernel(const E* __restrict__ src, E* __restrict__ dst, int R, int C) {
    __shared__ E tile[32][33];
    const size_t base = (size_t)blockIdx.z * R * C;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        if (r < R && c < C) tile[i][threadIdx.x] = src[base + (size_t)r * C + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < R && c < C) dst[base + (size_t)c * R + r] = tile[threadIdx.x][i];
    }
}

template <typename E>
int transpose_batched(const E* src, E* dst, int batch, int R, int C, cudaStream_t st) {
    if (batch == 0 || R == 0 || C == 0) return GSS_OK;
    GSS_REQUIRE(batch <= 65535, GSS_ERR_ARG, "transpose: batch %d too large", batch);
    dim3 grid((C + 31) / 32, (R + 31) / 32, batch), block(32, 8);
    transpose_kernel<E><<<grid, block, 0, st>>>(src, dst, R, C);
    GSS_LAUNCH_CHECK("transpose_kernel");
    return GSS_OK;
}

template int transpose_batched<float2>(const float2*, float2*, int, int, int, cudaStream_t);
template int transpose_batched<float>(const float*, float*, int, int, int, cudaStream_t);

}  // namespace gss

extern "C" {

#define GSS_PTRS(a, b) GSS_REQUIRE((a) && (b), GSS_ERR_ARG, "%s: null pointer", __func__)

// (B, D*T, F) -> (B, F, D*T)
int gss_pack_dtf_to_fdt_c64(const gss_c64* src, gss_c64* dst, int B, int D, int T, int F, void* stream) {
    GSS_PTRS(src, dst);
    return gss::transpose_batched<float2>((const float2*)src, (float2*)dst, B, D * T, F, (cudaStream_t)stream);
}
int gss_unpack_fdt_to_dtf_c64(const gss_c64* src, gss_c64* dst, int B, int D, int T, int F, void* stream) {
    GSS_PTRS(src, dst);
    return gss::transpose_batched<float2>((const float2*)src, (float2*)dst, B, F, D * T, (cudaStream_t)stream);
}
int gss_unpack_fkt_to_ktf_f32(const float* src, float* dst, int B, int K, int T, int F, void* stream) {
    GSS_PTRS(src, dst);
    return gss::transpose_batched<float>(src, dst, B, F, K * T, (cudaStream_t)stream);
}
int gss_pack_ktf_to_fkt_f32(const float* src, float* dst, int B, int K, int T, int F, void* stream) {
    GSS_PTRS(src, dst);
    return gss::transpose_batched<float>(src, dst, B, K * T, F, (cudaStream_t)stream);
}
int gss_unpack_ft_to_tf_c64(const gss_c64* src, gss_c64* dst, int B, int T, int F, void* stream) {
    GSS_PTRS(src, dst);
    return gss::transpose_batched<float2>((const float2*)src, (float2*)dst, B, F, T, (cudaStream_t)stream);
}

}  // extern "C"
